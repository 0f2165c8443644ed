"""Convolutional VAE over note tensors (BASELINE config #2), B200-native.

Drop-in for the reference's src/ae/model.py: ConvEncoder, ConvDecoder, VAE(cfg) with the same constructors, cfg keys
(LATENT_DIM, MAX_NOTES) and state_dict keys (encoder.conv.*, encoder._linear.1.*, fc_mu, fc_log_var, decoder.pre.*,
decoder.deconv.*).  VAE.forward(x) -> (recon, z, mu, log_var) runs as ONE native call (mg_vae_forward) and its
backward as one (mg_vae_backward); the reparameterisation noise is drawn with torch.randn_like exactly where the
reference draws it (model.py:127-133), so the torch RNG stream is the reference's.
"""
import torch
import torch.nn as nn

from melogan import blocks as B_
from melogan import engine as E
from melogan import runtime as R

_VAE_ENGINES = {}


def _engine_for(device, batch, max_notes, latent):
    if torch.device(device).type != "cuda":
        raise RuntimeError("melogan_b200 modules run on CUDA (sm_100a) only; there is no CPU fallback. "
                           "Move the module and its inputs to a CUDA device.")
    key = (str(device), int(batch), int(max_notes), int(latent), R.PRECISION)
    eng = _VAE_ENGINES.get(key)
    if eng is None:
        eng = _VAE_ENGINES[key] = E.VaeEngine(batch, max_notes, latent, precision=R.PRECISION, device=device)
    return eng


class ConvEncoder(nn.Module):
    """(B, MAX_NOTES, 4) -> (B, hidden_dim): three stride-2 Conv1d(k=5)+BatchNorm+ReLU, Flatten, Linear, ReLU.
    `_linear` is created on first use, as in the reference (model.py:27-36), because its width depends on MAX_NOTES."""

    def __init__(self, in_channels=4, latent_dim=128, hidden_dim=512):
        super().__init__()
        widths = (in_channels, 32, 64, 128)
        layers = []
        for ci, co in zip(widths[:-1], widths[1:]):
            layers += [nn.Conv1d(ci, co, kernel_size=5, stride=2, padding=2), nn.BatchNorm1d(co), nn.ReLU(inplace=True)]
        self.conv = nn.Sequential(*layers)
        self.latent_dim, self.hidden_dim = latent_dim, hidden_dim
        self._linear = None

    def build_linear(self, seq_len):
        length = seq_len
        for _ in range(3):
            length = (length + 2 * 2 - 5) // 2 + 1
        device = self.conv[0].weight.device
        self._linear = nn.Sequential(nn.Flatten(), nn.Linear(128 * length, self.hidden_dim), nn.ReLU(inplace=True)).to(device)

    def forward(self, x):
        """Stand-alone call (VAE.forward computes the hidden state inside its single native call): (B, MAX_NOTES, 4) ->
        (B, hidden_dim) on the native conv units / Linear of melogan.blocks, with full backward.  The first call also
        materialises `_linear`; in train mode build_linear's own pass over zeros (reference model.py:27-36) updates the
        BatchNorm running statistics once more, as the reference's does.  CPU tensors: only the reference's zero dummy pass
        (train_ae.py:75-77) is supported -- it creates `_linear`, updates the running statistics and returns None."""
        first = self._linear is None
        if first:
            self.build_linear(x.shape[1])
        if not x.is_cuda:
            if bool(x.any()) or (torch.is_grad_enabled() and x.requires_grad):
                raise RuntimeError("ConvEncoder: CUDA tensors only (the B200 path has no CPU fallback)")
            if self.training:
                self._dummy_pass_running_stats(2 if first else 1)
            return None
        if first and self.training:
            self._dummy_pass_running_stats(1)
        y = B_.run_conv_stack(self.conv, x)                          # channels-last (B, L, 128): no permute needed
        flat = y.permute(0, 2, 1).reshape(y.shape[0], -1)            # channel-major, like y.view(B, -1) on (B, C, L)
        return B_.run_mlp(nn.Sequential(*list(self._linear)[1:]), flat, self.training)

    @torch.no_grad()
    def _dummy_pass_running_stats(self, passes):
        """The reference's dummy pass (train_ae.py:75-77: model.encoder(zeros) in train mode, under no_grad) runs self.conv once
        in forward and once more inside build_linear (model.py:27-44), and BatchNorm updates its running statistics both
        times.  On an all-zero input every conv output is its bias at every position (the next layer sees ReLU(beta)), so
        each update is running_mean <- (1-m) running_mean + m * mean, running_var <- (1-m) running_var + m * 0 -- done here
        on the buffers, so that state_dict() after construction matches the reference's."""
        const = torch.zeros(self.conv[0].in_channels, device=self.conv[0].weight.device)
        layers = [(self.conv[i], self.conv[i + 1]) for i in range(0, len(self.conv), 3)]
        for _ in range(passes):
            c = const
            for conv, bn in layers:
                if bool(c.any()):          # a non-zero constant is not constant after zero padding: not the dummy pass
                    return
                mean = conv.bias.detach().clone() if conv.bias is not None else torch.zeros_like(bn.running_mean)
                m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked + 1)
                bn.running_mean.mul_(1.0 - m).add_(m * mean)
                bn.running_var.mul_(1.0 - m)
                bn.num_batches_tracked += 1
                c = torch.relu(bn.bias.detach())       # x_hat = 0 -> BatchNorm output = beta


class ConvDecoder(nn.Module):
    """(B, latent) -> (B, max_notes, 4): Linear, ReLU, Linear, ReLU, view(B,128,L0), two ConvTranspose1d(k=5,s=2)+BN+ReLU,
    ConvTranspose1d(k=5,s=2), Tanh (model.py:50-98)."""

    def __init__(self, out_channels=4, max_notes=512, latent_dim=128, hidden_dim=512):
        super().__init__()
        self.max_notes = max_notes
        self.reduced_len = max(1, max_notes // 8)
        self.pre = nn.Sequential(nn.Linear(latent_dim, hidden_dim), nn.ReLU(inplace=True),
                                 nn.Linear(hidden_dim, 128 * self.reduced_len), nn.ReLU(inplace=True))
        up = dict(kernel_size=5, stride=2, padding=2, output_padding=1)
        self.deconv = nn.Sequential(nn.ConvTranspose1d(128, 64, **up), nn.BatchNorm1d(64), nn.ReLU(inplace=True),
                                    nn.ConvTranspose1d(64, 32, **up), nn.BatchNorm1d(32), nn.ReLU(inplace=True),
                                    nn.ConvTranspose1d(32, out_channels, **up), nn.Tanh())

    def forward(self, z):
        """Stand-alone call (VAE.forward runs this block fused): (B, latent) -> (B, max_notes, 4) on the native operators of
        melogan.blocks: the Linear stack, view (B, 128, L0), the transposed-conv units and Tanh on channels-last
        activations, trimmed / zero-padded to max_notes like the reference (model.py:76-98)."""
        y = B_.run_mlp(self.pre, z, self.training)
        L0 = y.shape[1] // 128
        x = y[:, :128 * L0].reshape(y.shape[0], 128, L0).permute(0, 2, 1)
        out = B_.run_conv_stack(self.deconv, x)
        T = out.shape[1]
        if T > self.max_notes:
            out = out[:, :self.max_notes]
        elif T < self.max_notes:
            out = nn.functional.pad(out, (0, 0, 0, self.max_notes - T))
        return out


class _VaeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, eps, *params):
        eng = module._engine(x)
        eng.bind(module._tensors(), None)
        out = eng.forward(R.as_f32c(x), eps, train=True)
        for bn in module._batchnorms():
            bn.num_batches_tracked += 1
        ctx.module = module
        ctx.save_for_backward(x, eps)
        return out

    @staticmethod
    def backward(ctx, drecon, dz, dmu, dlv):
        module = ctx.module
        x, eps = ctx.saved_tensors
        eng = module._engine(x)
        P = dict(module._tensors())
        for k in E.VAE_BUFFER_KEYS:          # recompute with batch statistics WITHOUT advancing the running stats again
            P[k] = P[k].clone()
        G = R.fresh_grads(P, E.VAE_PARAM_KEYS)
        eng.bind(P, G)
        eng.forward(R.as_f32c(x), eps, train=True)
        f = lambda t: None if t is None else R.as_f32c(t)
        eng.backward(R.as_f32c(drecon) if drecon is not None else torch.zeros_like(x), f(dz), f(dmu), f(dlv))
        named = dict(module.named_parameters())
        return (None, None, None) + tuple(G[k] if named[k].requires_grad else None for k in E.VAE_PARAM_KEYS)


class VAE(nn.Module):
    """cfg keys: LATENT_DIM, MAX_NOTES (model.py:105-125)."""

    def __init__(self, cfg):
        super().__init__()
        latent_dim, hidden_dim = cfg['LATENT_DIM'], 512
        self.max_notes = cfg['MAX_NOTES']
        if self.max_notes % 8 != 0:
            raise NotImplementedError("MAX_NOTES must be a multiple of 8 on the CUDA path (the reference pads/crops otherwise)")
        self.encoder = ConvEncoder(in_channels=4, latent_dim=latent_dim, hidden_dim=hidden_dim)
        self.fc_mu = nn.Linear(hidden_dim, latent_dim)
        self.fc_log_var = nn.Linear(hidden_dim, latent_dim)
        self.decoder = ConvDecoder(out_channels=4, max_notes=self.max_notes, latent_dim=latent_dim, hidden_dim=hidden_dim)

    def _engine(self, x):
        return _engine_for(x.device, x.shape[0], self.max_notes, self.fc_mu.out_features)

    def _tensors(self):
        return R.params_of(self, E.VAE_PARAM_KEYS + E.VAE_BUFFER_KEYS)

    def _batchnorms(self):
        return [m for m in self.modules() if isinstance(m, nn.BatchNorm1d)]

    def reparameterize(self, mu, log_var):
        std = torch.exp(0.5 * log_var)
        return mu + torch.randn_like(std) * std

    def forward(self, x):
        if x.dim() != 3 or x.shape[1] != self.max_notes or x.shape[2] != 4:
            raise ValueError(f"expected (B, {self.max_notes}, 4) notes, got {tuple(x.shape)}")
        if self.encoder._linear is None:
            self.encoder.build_linear(self.max_notes)
        eps = torch.randn_like(self.fc_mu.bias.new_empty((x.shape[0], self.fc_mu.out_features)))
        if self.training:
            params = [dict(self.named_parameters())[k] for k in E.VAE_PARAM_KEYS]
            return _VaeFn.apply(self, x, eps, *params)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and x.requires_grad:
            raise NotImplementedError("eval-mode VAE has no backward on the CUDA path")
        eng = self._engine(x)
        eng.bind(self._tensors(), None)
        return eng.forward(R.as_f32c(x), eps, train=False)
