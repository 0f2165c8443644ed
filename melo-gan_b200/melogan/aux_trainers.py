"""Fast-path trainers of BASELINE configs #2 and #3 (the VAE and the emotion discriminator), shaped like GanTrainer.

Reference loops: src/ae/train_ae.py:100-122 (forward -> vae_loss -> zero_grad -> backward -> clip_grad_norm_(1.0) ->
AdamW) and src/emotion_discriminator/train_ed.py:61-74 (forward -> CrossEntropyLoss -> backward -> AdamW).  Here one
training step is a fixed sequence of native launches over flat parameter / gradient buffers

    VAE:  draw eps (device RNG) | zero grads | mg_vae_loss_step (forward + vae_loss + backward)
          | mg_adam_step_clipped (grad-norm reduction + clip + AdamW in the same launch sequence) | loss accumulation
    ED :  draw dropout masks | zero grads | mg_emotion_train_forward | mg_cross_entropy (loss, accuracy, dlogits)
          | mg_emotion_train_backward | mg_adam_step (AdamW form) | loss accumulation

with no host synchronisation inside, so the whole step replays as ONE CUDA graph over a static input batch.  The
drop-in modules (src.ae.model.VAE, src.emotion_discriminator.ed_model.EmotionDiscriminator) stay the owners of the
parameters: state_dict() / checkpoints keep working because the parameters are views of the flat buffers.
"""
import torch
import torch.nn as nn

from . import _native
from . import dist as D_
from . import engine as E
from .optim import FlatParams, FusedAdam

_U64 = (1 << 64) - 1


def _rng(t, kind, p, seed, ctr, stream, device):
    with torch.cuda.device(device):
        _native.check(_native.lib().mg_rng_fill_counter(t.data_ptr(), t.numel(), kind, p, seed & _U64, ctr, 1 << 22, stream))


class _StepGraph:
    """Shared plumbing: device loss accumulators, one-graph replay of `_step_body` over static inputs."""

    def _init_common(self, device, n_acc, seed, seed_offset, process_group):
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.loss_acc = torch.zeros(n_acc, device=self.device)
        self.rng_seed = (int(seed) * 0x9E3779B97F4A7C15 + seed_offset * 0xD1B54A32D192ED03) & _U64
        self.rng_counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._graph = None

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _bump_counter(self):
        with torch.cuda.device(self.device):
            _native.check(_native.lib().mg_counter_add(self.rng_counter.data_ptr(), 1, self._stream()))

    def _allreduce(self, flat):
        if self.world > 1:
            D_.allreduce_sum_(flat.grad, self.pg)          # 1/world is folded into Adam's grad_scale

    def replay(self):
        if self._graph is None:
            raise RuntimeError("capture() first")
        self._graph.replay()


class ReduceOnPlateau:
    """torch.optim.lr_scheduler.ReduceLROnPlateau (mode 'min', threshold_mode 'rel', cooldown 0) for an optimizer whose
    learning rate is a plain attribute (FusedAdam.lr); used as train_ae.py:82 / train_ed.py's build_scheduler use it.
    step() returns True when the learning rate changed (a captured step graph has the old value baked in)."""

    def __init__(self, opt, factor=0.5, patience=5, threshold=1e-4, min_lr=0.0):
        self.opt, self.factor, self.patience, self.threshold, self.min_lr = opt, factor, patience, threshold, min_lr
        self.best, self.bad = float("inf"), 0

    def step(self, metric):
        if metric < self.best * (1.0 - self.threshold):
            self.best, self.bad = metric, 0
        else:
            self.bad += 1
        if self.bad > self.patience:
            self.bad = 0
            new = max(self.opt.lr * self.factor, self.min_lr)
            if self.opt.lr - new > 1e-8:
                self.opt.lr = new
                return True
        return False


def find_split_dir(splits_dir, split):
    """<SPLITS_DIR>/<split> of the pre-saved arrays.  The reference's GAN path names the directory after the split CSV's stem
    (train_gan.py:45-48 -> data/splits/train_split/), its encode.py and configs after the bare split (data/splits/train/):
    both are accepted everywhere, first match wins."""
    import os
    stem = split[:-6] if split.endswith("_split") else split
    for name in (split, stem, stem + "_split"):
        d = os.path.join(splits_dir, name)
        if os.path.isdir(d):
            return d
    return os.path.join(splits_dir, split)


class VaeTrainer(_StepGraph):
    """Config #2.  cfg: config/ae_config.yaml (LATENT_DIM, MAX_NOTES, BATCH_SIZE, LR, WEIGHT_DECAY, BETA, ...)."""

    def __init__(self, cfg, batch=None, precision="fp32", device=None, model=None, max_norm=1.0, process_group=None,
                 seed_offset=0):
        from src.ae.model import VAE
        if not torch.cuda.is_available():
            raise RuntimeError("VaeTrainer needs a CUDA (sm_100a) device; there is no CPU fallback")
        self._init_common(device, 4, cfg.get('SEED', 0), seed_offset, process_group)
        self.cfg, self.precision = cfg, precision
        self.B, self.T, self.latent = int(batch or cfg['BATCH_SIZE']), int(cfg['MAX_NOTES']), int(cfg['LATENT_DIM'])
        self.model = (model if model is not None else VAE(cfg)).to(self.device)
        if self.model.encoder._linear is None:       # the reference's dummy pass (train_ae.py:75-77): materialises the lazy
            self.model.train()                       # Linear and, in train mode, updates the BatchNorm running statistics
            with torch.no_grad():
                self.model.encoder(torch.zeros(1, self.T, 4, device=self.device))
            self.model.to(self.device)
        self.model.train()
        named = dict(self.model.named_parameters())
        self.flat = FlatParams([named[k] for k in E.VAE_PARAM_KEYS])
        self.opt = FusedAdam(self.flat, lr=float(cfg.get('LR', 1e-4)), weight_decay=float(cfg.get('WEIGHT_DECAY', 1e-5)),
                             decoupled=True, eps=1e-8, max_norm=max_norm)
        self.opt.grad_scale = 1.0 / self.world
        self.engine = E.VaeEngine(self.B, max_notes=self.T, latent_dim=self.latent, precision=precision, device=self.device)
        self.rebind()
        self.eps = torch.empty((self.B, self.latent), device=self.device)
        self.metrics = torch.empty(3, device=self.device)          # [total, recon MSE, KLD] of the last step
        self.beta_dev = None
        self.beta = float(cfg.get('BETA', 1.0))

    def rebind(self):
        named = dict(self.model.named_parameters()); named.update(dict(self.model.named_buffers()))
        P = {k: named[k].data for k in E.VAE_PARAM_KEYS + E.VAE_BUFFER_KEYS}
        G = {k: named[k].grad for k in E.VAE_PARAM_KEYS}
        self.engine.bind(P, G)

    def step(self, x, eps=None, beta=None):
        """One training step of train_ae.py:107-122 on the CUDA batch x (B, MAX_NOTES, 4); returns the device tensor
        [loss, recon, kld].  eps: inject the reparameterisation noise (tests); default: device RNG."""
        if eps is None:
            _rng(self.eps, 0, 0.0, self.rng_seed + 1, self.rng_counter.data_ptr(), self._stream(), self.device)
            self._bump_counter()
            eps = self.eps
        self.opt.zero_grad()
        self.engine.loss_step(x, eps, self.beta if beta is None else beta, metrics=self.metrics)
        for bn in self.model._batchnorms():
            bn.num_batches_tracked += 1
        self._allreduce(self.flat)
        self.opt.step()
        self.loss_acc[0:3] += self.metrics
        self.loss_acc[3] += 1
        return self.metrics

    def capture(self):
        """Captures step() over a static input buffer (returned).  Run one eager step first.  beta is baked into the
        graph: re-capture when the KLD warm-up changes it (once per epoch in train_ae.py:104)."""
        if self.world > 1:
            raise NotImplementedError("graph capture of the data-parallel VAE step (NCCL stays outside graphs here)")
        self.s_x = torch.zeros((self.B, self.T, 4), device=self.device)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.step(self.s_x)
        self._graph = g
        return self.s_x

    def epoch_means(self):
        """(loss, recon, kld) means since the last call; one host sync."""
        if self.world > 1:
            D_.allreduce_sum_(self.loss_acc, self.pg)
        a = self.loss_acc.cpu()
        self.loss_acc.zero_()
        n = max(a[3].item(), 1.0)
        return a[0].item() / n, a[1].item() / n, a[2].item() / n

    @torch.no_grad()
    def evaluate(self, x, eps=None):
        """Eval-mode forward + vae_loss(beta=1) of the validation loop (train_ae.py:132-141); device tensor [3]."""
        self.model.eval()
        try:
            recon, z, mu, log_var = self.model(x)
        finally:
            self.model.train()
        recon_loss = torch.mean((recon - x) ** 2)
        kld = -0.5 * torch.mean(1 + log_var - mu.pow(2) - log_var.exp())
        return torch.stack([recon_loss + kld, recon_loss, kld])


class EdTrainer(_StepGraph):
    """Config #3.  cfg: config/ed_config.yaml (batch_size, dropout, optimizer {name, lr, betas, weight_decay}, ...)."""

    def __init__(self, cfg, batch=None, precision="fp32", device=None, model=None, process_group=None, seed_offset=0):
        from src.emotion_discriminator.ed_model import EmotionDiscriminator
        if not torch.cuda.is_available():
            raise RuntimeError("EdTrainer needs a CUDA (sm_100a) device; there is no CPU fallback")
        self._init_common(device, 3, cfg.get('seed', 42), seed_offset, process_group)
        self.cfg, self.precision = cfg, precision
        self.B, self.T = int(batch or cfg['batch_size']), int(cfg.get('max_notes', 512))
        self.n_classes = int(cfg.get('n_classes', 4))
        self.model = (model if model is not None else EmotionDiscriminator(cfg)).to(self.device)
        if self.model.input_mode != 'notes':
            raise NotImplementedError("EdTrainer: input_mode 'notes' (config/ed_config.yaml)")
        if getattr(self.model, "use_sn", False):
            raise NotImplementedError("EdTrainer: the fused step binds raw weights; a spectral-norm model trains through "
                                      "model(x) (unfused block operators) and a torch optimizer")
        self.model.train()
        named = dict(self.model.named_parameters())
        self.flat = FlatParams([named[k] for k in E.ED_GRAD_KEYS])
        o = cfg.get("optimizer", {})
        name = str(o.get("name", "AdamW")).lower()
        if name not in ("adamw", "adam"):
            raise ValueError(f"Unsupported optimizer {name}")
        self.opt = FusedAdam(self.flat, lr=float(o.get("lr", 2e-4)), betas=tuple(o.get("betas", [0.9, 0.999])),
                             weight_decay=float(o.get("weight_decay", 0.0)), decoupled=(name == "adamw"))
        self.opt.grad_scale = 1.0 / self.world
        self.dropout = float(self.model.dropout)
        self.engine = E.GanEngine(self.B, precision=precision, max_notes=self.T, n_classes=self.n_classes, device=self.device)
        self.rebind()
        self.mask1 = torch.empty((self.B, 256), device=self.device)
        self.mask2 = torch.empty((self.B, 128), device=self.device)
        self.dlogits = torch.empty((self.B, self.n_classes), device=self.device)
        self.metrics = torch.empty(2, device=self.device)          # [mean CE, accuracy] of the last step

    def rebind(self):
        named = dict(self.model.named_parameters()); named.update(dict(self.model.named_buffers()))
        P = {k: named[k].data for k in E.ED_KEYS}
        G = {k: named[k].grad for k in E.ED_GRAD_KEYS}
        self.engine.bind(E.MOD_ED, P, G)

    @staticmethod
    def check_labels(y, n_classes):
        """nn.CrossEntropyLoss raises on targets outside [0, n_classes); the fused kernel indexes the logits row with
        them, so they are validated on the host once per data set (emotion_to_index returns -1 for unknown moods)."""
        if y.numel() and (int(y.min()) < 0 or int(y.max()) >= n_classes):
            raise IndexError(f"Target out of bounds: labels must be in [0, {n_classes})")

    def step(self, x, y, masks=None):
        """One training step of train_ed.py:61-74 on CUDA tensors x (B, max_notes, 4), y (B,) int64 (validated by the
        caller with check_labels); returns the device tensor [loss, accuracy]."""
        if masks is None:
            keep, ctr, st = 1.0 - self.dropout, self.rng_counter.data_ptr(), self._stream()
            _rng(self.mask1, 2, keep, self.rng_seed + 2, ctr, st, self.device)
            _rng(self.mask2, 2, keep, self.rng_seed + 3, ctr, st, self.device)
            self._bump_counter()
            masks = (self.mask1, self.mask2)
        self.opt.zero_grad()
        logits = self.engine.emotion_train_forward(x, masks[0], masks[1], dropout_p=self.dropout)
        for blk in self.model.encoder.conv:
            blk.net[1].num_batches_tracked += 1
        with torch.cuda.device(self.device):
            _native.call("mg_cross_entropy", logits.data_ptr(), y.data_ptr(), self.B, self.n_classes,
                         self.dlogits.data_ptr(), self.metrics.data_ptr(), self._stream())
        self.engine.emotion_train_backward(self.dlogits, want_dnotes=False)
        self._allreduce(self.flat)
        self.opt.step()
        self.loss_acc[0:2] += self.metrics
        self.loss_acc[2] += 1
        return self.metrics

    def capture(self):
        """Captures step() over static (x, y) buffers (returned).  Run one eager step first."""
        if self.world > 1:
            raise NotImplementedError("graph capture of the data-parallel ED step (NCCL stays outside graphs here)")
        self.s_x = torch.zeros((self.B, self.T, 4), device=self.device)
        self.s_y = torch.zeros(self.B, dtype=torch.int64, device=self.device)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.step(self.s_x, self.s_y)
        self._graph = g
        return self.s_x, self.s_y

    def epoch_means(self):
        """(loss, accuracy) means since the last call; one host sync."""
        if self.world > 1:
            D_.allreduce_sum_(self.loss_acc, self.pg)
        a = self.loss_acc.cpu()
        self.loss_acc.zero_()
        n = max(a[2].item(), 1.0)
        return a[0].item() / n, a[1].item() / n

    @torch.no_grad()
    def evaluate(self, x, y):
        """Eval-mode forward + CE / accuracy of the validation loop; device tensor [loss, accuracy] for this batch
        (any batch size: the eval engine is cached per size)."""
        self.model.eval()
        try:
            logits = self.model(x)
        finally:
            self.model.train()
        out = torch.empty(2, device=self.device)
        with torch.cuda.device(self.device):
            _native.call("mg_cross_entropy", logits.contiguous().data_ptr(), y.data_ptr(), int(x.shape[0]), self.n_classes,
                         None, out.data_ptr(), self._stream())
        return out
