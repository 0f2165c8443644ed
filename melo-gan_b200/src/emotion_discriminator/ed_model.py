"""Emotion discriminator (classifier over note tensors), B200-native.

Drop-in for the reference's src/emotion_discriminator/ed_model.py: ConvBlock1D, NotesEncoder,
MLPClassifier, EmotionDiscriminator with the same constructors, cfg keys and state_dict keys.  On the
GAN hot path the module is frozen and in eval mode (src/gan/train_gan.py:131-133): BatchNorm is folded
into the conv epilogue and the forward + input-gradient run behind mg_emotion_forward /
mg_emotion_backward_input.  In train mode (BASELINE config #3: BatchNorm batch statistics, dropout, all parameter
gradients) it runs behind mg_emotion_train_forward / mg_emotion_train_backward.
"""
from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F

from melogan import blocks as B_
from melogan import engine as E
from melogan import runtime as R


class ConvBlock1D(nn.Module):
    def __init__(self, in_ch, out_ch, kernel_size=3, stride=1, padding=1, use_sn=False):
        super().__init__()
        conv = nn.Conv1d(in_ch, out_ch, kernel_size, stride, padding)
        if use_sn:                           # same wrapper, same state_dict keys (weight_orig, weight_u, weight_v) as the reference
            from torch.nn.utils import spectral_norm
            conv = spectral_norm(conv)
        self.net = nn.Sequential(conv, nn.BatchNorm1d(out_ch), nn.GELU())

    def forward(self, x):
        """Stand-alone call (the emotion discriminator runs its blocks fused): (B, C_in, T) -> (B, C_out, T) on the native
        conv unit of melogan.blocks (Conv1d + BatchNorm1d + GELU, train or eval, full backward)."""
        return B_.run_conv_stack(self.net, x.permute(0, 2, 1)).permute(0, 2, 1)


class NotesEncoder(nn.Module):
    """(B, max_notes, note_dim) -> (B, hidden_dim): Conv1d blocks over the note axis, mean-pool, Linear."""

    def __init__(self, note_dim: int = 4, hidden_dim: int = 256, num_blocks: int = 4, use_sn: bool = False):
        super().__init__()
        blocks, c_in, c_out = [], note_dim, 64
        for i in range(num_blocks):
            first = i == 0
            blocks.append(ConvBlock1D(c_in, c_out, kernel_size=5 if first else 3, padding=2 if first else 1, use_sn=use_sn))
            c_in, c_out = c_out, min(c_out * 2, hidden_dim)
        self.conv = nn.Sequential(*blocks)
        self.pool = nn.AdaptiveAvgPool1d(1)
        self.project = nn.Linear(c_in, hidden_dim)

    def forward(self, notes):
        """Stand-alone call: (B, T, note_dim) -> (B, hidden_dim).  The notes are already channels-last, so the reference's
        permute (ed_model.py:63-64) disappears; conv units, mean over the note axis, projection."""
        x = B_.run_conv_stack(self.conv, notes)
        return B_.linear(x.mean(dim=1), self.project.weight, self.project.bias)


class MLPClassifier(nn.Module):
    def __init__(self, in_dim: int, hidden_dims=(256, 128), n_classes: int = 4, dropout: float = 0.2, use_sn: bool = False):
        super().__init__()
        stack, width = [], in_dim
        for h in hidden_dims:
            lin = nn.Linear(width, h)
            if use_sn:
                from torch.nn.utils import spectral_norm
                lin = spectral_norm(lin)
            stack += [lin, nn.GELU(), nn.Dropout(dropout)]
            width = h
        self.net = nn.Sequential(*stack)
        self.head = nn.Linear(width, n_classes)

    def forward(self, x, masks=None):
        """Stand-alone call, and the whole model when input_mode is 'latent' (ed_model.py:128-136): Linear -> GELU ->
        Dropout per hidden layer, then the head, on the native operators of melogan.blocks with full backward."""
        h = B_.run_mlp(self.net, x, self.training, masks)
        return B_.linear(h, B_.effective_weight(self.head, h), self.head.bias)


class _EmotionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, notes):
        eng = module._engine(notes)
        eng.bind(E.MOD_ED, R.params_of(module, E.ED_KEYS), None)
        logits = eng.emotion_forward(R.as_f32c(notes))
        ctx.module = module
        ctx.save_for_backward(notes)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        module = ctx.module
        notes, = ctx.saved_tensors
        eng = module._engine(notes)
        eng.bind(E.MOD_ED, R.params_of(module, E.ED_KEYS), None)
        eng.emotion_forward(R.as_f32c(notes))      # recompute, then the input gradient
        return None, eng.emotion_backward_input(R.as_f32c(dlogits))


class _EmotionTrainFn(torch.autograd.Function):
    """Train-mode forward/backward (BASELINE config #3, reference train_ed.py:61-74): BatchNorm batch statistics,
    MLP dropout with keep-masks drawn like nn.Dropout would, full parameter gradients."""

    @staticmethod
    def forward(ctx, module, notes, mask1, mask2, *params):
        eng = module._engine(notes)
        P = R.params_of(module, E.ED_KEYS)
        eng.bind(E.MOD_ED, P, None)
        logits = eng.emotion_train_forward(R.as_f32c(notes), mask1, mask2, dropout_p=module.dropout)
        for blk in module.encoder.conv:
            blk.net[1].num_batches_tracked += 1
        ctx.module = module
        ctx.save_for_backward(notes, mask1, mask2)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        module = ctx.module
        notes, mask1, mask2 = ctx.saved_tensors
        eng = module._engine(notes)
        P = dict(R.params_of(module, E.ED_KEYS))
        for k in E.ED_KEYS:                      # recompute with batch statistics WITHOUT advancing the running stats again
            if k.endswith(("running_mean", "running_var")):
                P[k] = P[k].clone()
        G = R.fresh_grads(P, E.ED_GRAD_KEYS)
        eng.bind(E.MOD_ED, P, G)
        eng.emotion_train_forward(R.as_f32c(notes), mask1, mask2, dropout_p=module.dropout)
        dnotes = eng.emotion_train_backward(R.as_f32c(dlogits), want_dnotes=ctx.needs_input_grad[1])
        named = dict(module.named_parameters())
        grads = tuple(G[k] if named[k].requires_grad else None for k in E.ED_GRAD_KEYS)
        return (None, dnotes, None, None) + grads


class EmotionDiscriminator(nn.Module):
    """cfg keys: input_mode, n_classes, use_spectral_norm, dropout, latent_dim, note_dim, notes_hidden,
    notes_blocks, mlp_hidden (reference ed_model.py:115-145)."""

    def __init__(self, cfg: Dict):
        super().__init__()
        self.cfg = cfg.copy()
        self.input_mode = cfg.get('input_mode', 'latent')
        self.n_classes = cfg.get('n_classes', 4)
        self.use_sn = cfg.get('use_spectral_norm', False)
        self.dropout = cfg.get('dropout', 0.2)
        hidden = tuple(cfg.get('mlp_hidden', (256, 128)))
        if self.input_mode == 'latent':
            self.encoder = None
            in_dim = cfg.get('latent_dim', 128)
        elif self.input_mode == 'notes':
            in_dim = cfg.get('notes_hidden', 256)
            self.encoder = NotesEncoder(note_dim=cfg.get('note_dim', 4), hidden_dim=in_dim,
                                        num_blocks=cfg.get('notes_blocks', 4), use_sn=self.use_sn)
        else:
            raise ValueError("input_mode must be 'latent' or 'notes'")
        self.classifier = MLPClassifier(in_dim=in_dim, hidden_dims=hidden, n_classes=self.n_classes,
                                        dropout=self.dropout, use_sn=self.use_sn)

    def _engine(self, notes):
        c = self.cfg
        if (c.get('notes_hidden', 256), c.get('notes_blocks', 4), tuple(c.get('mlp_hidden', (256, 128))),
                c.get('note_dim', 4)) != (256, 4, (256, 128), 4):
            raise NotImplementedError("native emotion discriminator implements the config/ed_config.yaml shape "
                                      "(notes_hidden 256, 4 blocks, mlp_hidden [256,128], note_dim 4)")
        return R.engine_for(notes.device, notes.shape[0], max_notes=notes.shape[1], n_classes=self.n_classes)

    def forward(self, x: torch.Tensor, masks=None) -> torch.Tensor:
        if self.input_mode == 'latent':
            if x.dim() != 2:
                raise ValueError(f"Expected latent input shape (B, latent_dim), got {x.shape}")
            return self.classifier(x, masks)          # 'latent' mode IS the MLP classifier (ed_model.py:156-160)
        if x.dim() != 3:
            raise ValueError(f"Expected notes input shape (B, T, note_dim), got {x.shape}")
        if self.use_sn:
            # spectral norm (off in config/ed_config.yaml): the fused kernels bind raw weights, so this variant runs unfused on the
            # stand-alone block operators (conv units + Linears), whose weights come through the spectral-norm pre-hooks
            return self.classifier(self.encoder(x), masks)
        if self.training:
            keep = 1.0 - self.dropout
            B, dev = x.shape[0], x.device
            if masks is None:      # the two bernoulli draws nn.Dropout would make, in order
                masks = (torch.bernoulli(torch.full((B, 256), keep, device=dev)),
                         torch.bernoulli(torch.full((B, 128), keep, device=dev)))
            named = dict(self.named_parameters())
            return _EmotionTrainFn.apply(self, x, masks[0], masks[1], *[named[k] for k in E.ED_GRAD_KEYS])
        if any(p.requires_grad for p in self.parameters()) and torch.is_grad_enabled():
            raise NotImplementedError("the native emotion discriminator is frozen: set requires_grad=False on its "
                                      "parameters (train_gan.py:131-132) to get the input gradient")
        return _EmotionFn.apply(self, x)

    def predict_proba(self, x: torch.Tensor) -> torch.Tensor:
        return F.softmax(self.forward(x), dim=-1)

    def predict(self, x: torch.Tensor) -> torch.Tensor:
        return self.forward(x).argmax(dim=-1)

    def freeze_encoder(self):
        if self.encoder is not None:
            for p in self.encoder.parameters():
                p.requires_grad = False

    def unfreeze_encoder(self):
        if self.encoder is not None:
            for p in self.encoder.parameters():
                p.requires_grad = True
