"""Generator and WGAN-GP critic of MELO-GAN, B200-native.

Drop-in for the reference's src/gan/models.py: same class names, constructor signatures, attributes and
state_dict keys (so checkpoints and `weights_init`, which matches on class names containing 'Conv' /
'Linear', keep working).  The nn.Linear / nn.ConvTranspose1d / nn.Conv1d / nn.BatchNorm1d children are
parameter containers created in the reference's order (same RNG consumption for a given seed); the
arithmetic of forward and backward runs in the sm_100a kernels behind the mg_generator_* and
mg_discriminator_* entry points, in channels-last layout (the permutes at models.py:73,159 vanish).
"""
import torch
import torch.nn as nn

from melogan import blocks as B_
from melogan import engine as E
from melogan import runtime as R


class NoiseToLatent(nn.Module):
    """MLP (noise ++ conditioning) -> decoder latent.  reference models.py:20-29"""

    def __init__(self, noise_dim, out_dim, hidden=512):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(noise_dim, hidden), nn.ReLU(True), nn.Linear(hidden, out_dim))

    def forward(self, z):
        """Stand-alone call (the Generator runs this block fused): Linear -> ReLU -> Linear on the native operators of
        melogan.blocks, with full backward."""
        return B_.run_mlp(self.net, z, self.training)


class GeneratorDecoder(nn.Module):
    """latent -> (B, max_notes, out_channels) through two Linears and three stride-2 transposed convs.
    reference models.py:32-83"""

    def __init__(self, latent_dim=128, max_notes=512, out_channels=4):
        super().__init__()
        self.latent_dim = latent_dim
        self.max_notes = max_notes
        self.reduced_len = max(1, max_notes // 8)
        self.pre = nn.Sequential(nn.Linear(latent_dim, 512), nn.ReLU(True),
                                 nn.Linear(512, 256 * self.reduced_len), nn.ReLU(True))
        up = dict(kernel_size=5, stride=2, padding=2, output_padding=1)
        self.deconv = nn.Sequential(nn.ConvTranspose1d(256, 128, **up), nn.BatchNorm1d(128), nn.ReLU(True),
                                    nn.ConvTranspose1d(128, 64, **up), nn.BatchNorm1d(64), nn.ReLU(True),
                                    nn.ConvTranspose1d(64, out_channels, **up))

    def pre_forward(self, latent):
        """The Linear stack in front of the transposed convs (models.py:46-51,66-69), stand-alone: (B, latent) ->
        (B, 256, reduced_len) on the native operators of melogan.blocks."""
        return B_.run_mlp(self.pre, latent, self.training).view(-1, 256, self.reduced_len)

    def forward(self, latent):
        """Stand-alone call (the Generator runs this block fused): (B, latent) -> (B, max_notes, out_channels) on the native
        operators of melogan.blocks -- the Linear stack, then the three transposed-conv units on channels-last activations
        (models.py:66-83; the reference's final permute back to (B, T, C) is the layout the units already produce)."""
        x = self.pre_forward(latent).permute(0, 2, 1)             # (B, reduced_len, 256) channels-last
        out = B_.run_conv_stack(self.deconv, x)
        T = out.shape[1]
        if T > self.max_notes:
            out = out[:, :self.max_notes]
        elif T < self.max_notes:
            out = torch.nn.functional.pad(out, (0, 0, 0, self.max_notes - T))
        return out


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, noise, emb, cond, *params):
        eng = module._engine(noise)
        P = R.params_of(module, E.G_PARAM_KEYS + E.G_BUFFER_KEYS)
        eng.bind(E.MOD_G, P, None)
        if cond is not None:
            eng.set_condition(R.as_f32c(cond))
        notes, latent = eng.generator_forward(R.as_f32c(noise), R.as_f32c(emb), train=module.training)
        if module.training:
            for bn in (module.decoder.deconv[1], module.decoder.deconv[4]):
                bn.num_batches_tracked += 1
        ctx.module, ctx.train, ctx.has_cond = module, module.training, cond is not None
        ctx.save_for_backward(*((noise, emb) if cond is None else (noise, emb, cond)))
        return notes, latent

    @staticmethod
    def backward(ctx, dnotes, dlatent):
        module = ctx.module
        noise, emb = ctx.saved_tensors[:2]
        eng = module._engine(noise)
        if ctx.has_cond:
            eng.set_condition(R.as_f32c(ctx.saved_tensors[2]))
        P = R.params_of(module, E.G_PARAM_KEYS + E.G_BUFFER_KEYS)
        G = R.fresh_grads(P, E.G_PARAM_KEYS)
        if ctx.train:      # recompute with batch statistics but WITHOUT advancing the running stats twice
            P = dict(P)
            for k in E.G_BUFFER_KEYS:
                P[k] = P[k].clone()
        eng.bind(E.MOD_G, P, G)
        eng.generator_forward(R.as_f32c(noise), R.as_f32c(emb), train=ctx.train)
        dn = (R.as_f32c(dnotes) if dnotes is not None else
              torch.zeros((noise.shape[0], module.max_notes, module.note_dim), device=noise.device))
        demb = eng.generator_backward(dn, R.as_f32c(dlatent) if dlatent is not None else None)
        named = dict(module.named_parameters())
        grads = tuple(G[k] if named[k].requires_grad else None for k in E.G_PARAM_KEYS)
        return (None, None, demb, None) + grads       # the AE latent is data: no gradient (reference feeds it detached)


class Generator(nn.Module):
    def __init__(self, noise_dim=128, latent_dim=128, mode="conditioning", hidden=512, max_notes=512, note_dim=4,
                 numeric_embed_dim=0):
        super().__init__()
        assert mode in ("conditioning", "warm_start")
        self.mode = mode
        self.noise_dim = noise_dim
        self.latent_dim = latent_dim
        self.max_notes = max_notes
        self.note_dim = note_dim
        self.numeric_embed_dim = numeric_embed_dim
        self.hidden = hidden
        self.input_dim = noise_dim + numeric_embed_dim + (latent_dim if mode == "conditioning" else 0)
        print(f"[G] Init Generator. Mode: {self.mode}. Input MLP dim: {self.input_dim}")
        self.noise_to_latent = NoiseToLatent(self.input_dim, latent_dim, hidden=hidden)
        self.decoder = GeneratorDecoder(latent_dim=latent_dim, max_notes=max_notes, out_channels=note_dim)

    def _engine(self, noise):
        return R.engine_for(noise.device, noise.shape[0], max_notes=self.max_notes, note_dim=self.note_dim,
                            noise_dim=self.noise_dim, latent_dim=self.latent_dim, gen_hidden=self.hidden,
                            embed_dim=self.numeric_embed_dim, cond_dim=self.latent_dim if self.mode == "conditioning" else 0)

    def forward(self, noise, encoder_latent=None, numeric_embedding=None):
        """noise (B, noise_dim); encoder_latent (B, latent_dim) only in 'conditioning' mode;
        numeric_embedding (B, numeric_embed_dim).  Returns (notes (B, max_notes, note_dim), latent)."""
        if self.numeric_embed_dim > 0:
            assert numeric_embedding is not None, "numeric_embedding is required"
        if self.mode == "conditioning":
            assert encoder_latent is not None, "conditioning mode requires encoder latent input"
        if self.numeric_embed_dim <= 0:
            raise NotImplementedError("the native generator expects the numeric embedding of gan_config.yaml "
                                      "(numeric_embed_dim > 0)")
        if self.max_notes % 8 != 0 or self.note_dim != 4:
            raise NotImplementedError("native generator: max_notes must be a multiple of 8 and note_dim 4")
        names = E.G_PARAM_KEYS
        named = dict(self.named_parameters())
        cond = encoder_latent.detach() if self.mode == "conditioning" else None
        return _GeneratorFn.apply(self, noise, numeric_embedding, cond, *[named[k] for k in names])


class _CriticFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, notes, emb, *params):
        eng = module._engine(notes)
        eng.bind(E.MOD_D, R.params_of(module, E.D_KEYS), None)
        score = eng.critic_forward(R.as_f32c(notes), R.as_f32c(emb) if emb is not None else None)
        ctx.module, ctx.has_emb = module, emb is not None
        ctx.save_for_backward(notes, emb)
        return score

    @staticmethod
    def backward(ctx, dscore):
        module = ctx.module
        notes, emb = ctx.saved_tensors
        eng = module._engine(notes)
        P = R.params_of(module, E.D_KEYS)
        G = R.fresh_grads(P, E.D_KEYS)
        eng.bind(E.MOD_D, P, G)
        eng.critic_forward(R.as_f32c(notes), R.as_f32c(emb) if emb is not None else None)   # recompute
        need_params = any(p.requires_grad for p in module.parameters())
        dnotes, demb = eng.critic_backward(R.as_f32c(dscore), param_grads=need_params,
                                           want_dnotes=ctx.needs_input_grad[1],
                                           want_demb=ctx.has_emb and ctx.needs_input_grad[2])
        named = dict(module.named_parameters())
        grads = tuple(G[k] if (need_params and named[k].requires_grad) else None for k in E.D_KEYS)
        return (None, dnotes, demb) + grads


class Discriminator(nn.Module):
    """WGAN-GP critic: real-vs-fake score only, no BatchNorm, raw (unbounded) output.
    reference models.py:132-169"""

    def __init__(self, max_notes=512, note_dim=4, emb_dim=256, numeric_embed_dim=0):
        super().__init__()
        down = dict(kernel_size=5, stride=2, padding=2)
        self.conv = nn.Sequential(nn.Conv1d(note_dim, 64, **down), nn.LeakyReLU(0.2, inplace=True),
                                  nn.Conv1d(64, 128, **down), nn.LeakyReLU(0.2, inplace=True),
                                  nn.Conv1d(128, 256, **down), nn.LeakyReLU(0.2, inplace=True))
        self.pool = nn.AdaptiveAvgPool1d(1)
        self.fc = nn.Sequential(nn.Flatten(), nn.Linear(256, emb_dim), nn.LeakyReLU(0.2, inplace=True))
        self.combined_dim = emb_dim + numeric_embed_dim
        self.real_fake = nn.Linear(self.combined_dim, 1)
        self.max_notes, self.note_dim, self.emb_dim, self.numeric_embed_dim = max_notes, note_dim, emb_dim, numeric_embed_dim

    def _engine(self, notes):
        if self.emb_dim != 256 or self.note_dim != 4:
            raise NotImplementedError("native critic: emb_dim must be 256 and note_dim 4 (gan_config.yaml shapes)")
        return R.engine_for(notes.device, notes.shape[0], max_notes=notes.shape[1], note_dim=self.note_dim,
                            embed_dim=max(self.numeric_embed_dim, 1))

    def forward(self, notes, numeric_embedding=None):
        if notes.dim() != 3 or notes.shape[2] != self.note_dim:
            raise ValueError(f"expected notes of shape (B, T, {self.note_dim}), got {tuple(notes.shape)}")
        if numeric_embedding is None and self.numeric_embed_dim > 0:
            # reference: Linear(combined_dim) would fail on the un-concatenated feature vector
            raise RuntimeError("numeric_embedding is required by this critic (numeric_embed_dim > 0)")
        named = dict(self.named_parameters())
        return _CriticFn.apply(self, notes, numeric_embedding, *[named[k] for k in E.D_KEYS])
