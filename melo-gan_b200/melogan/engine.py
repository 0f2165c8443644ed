"""Thin host wrapper of the mg_gan_* C entry points (include/melogan_b200.h).

The engine owns only the native context (workspaces); parameters, gradients and optimizer state are
torch CUDA tensors owned by the caller (the drop-in nn.Modules of src/gan/models.py etc.) and are
bound by pointer in the reference's state_dict order.
"""
import ctypes

import torch

from . import _native

MOD_E, MOD_G, MOD_D, MOD_ED = 0, 1, 2, 3

E_KEYS = ["net.0.weight", "net.0.bias", "net.1.weight", "net.1.bias", "net.4.weight", "net.4.bias",
          "net.7.weight", "net.7.bias"]
G_PARAM_KEYS = ["noise_to_latent.net.0.weight", "noise_to_latent.net.0.bias", "noise_to_latent.net.2.weight",
                "noise_to_latent.net.2.bias", "decoder.pre.0.weight", "decoder.pre.0.bias", "decoder.pre.2.weight",
                "decoder.pre.2.bias", "decoder.deconv.0.weight", "decoder.deconv.0.bias", "decoder.deconv.1.weight",
                "decoder.deconv.1.bias", "decoder.deconv.3.weight", "decoder.deconv.3.bias", "decoder.deconv.4.weight",
                "decoder.deconv.4.bias", "decoder.deconv.6.weight", "decoder.deconv.6.bias"]
G_BUFFER_KEYS = ["decoder.deconv.1.running_mean", "decoder.deconv.1.running_var", "decoder.deconv.4.running_mean",
                 "decoder.deconv.4.running_var"]
D_KEYS = ["conv.0.weight", "conv.0.bias", "conv.2.weight", "conv.2.bias", "conv.4.weight", "conv.4.bias",
          "fc.1.weight", "fc.1.bias", "real_fake.weight", "real_fake.bias"]
ED_KEYS = []
for _i in range(4):
    _p = f"encoder.conv.{_i}.net."
    ED_KEYS += [_p + "0.weight", _p + "0.bias", _p + "1.weight", _p + "1.bias", _p + "1.running_mean",
                _p + "1.running_var"]
ED_KEYS += ["encoder.project.weight", "encoder.project.bias", "classifier.net.0.weight", "classifier.net.0.bias",
            "classifier.net.3.weight", "classifier.net.3.bias", "classifier.head.weight", "classifier.head.bias"]

ED_GRAD_KEYS = [k for k in ED_KEYS if not k.endswith(("running_mean", "running_var"))]

PARAM_KEYS = {MOD_E: E_KEYS, MOD_G: G_PARAM_KEYS + G_BUFFER_KEYS, MOD_D: D_KEYS, MOD_ED: ED_KEYS}
GRAD_KEYS = {MOD_E: E_KEYS, MOD_G: G_PARAM_KEYS, MOD_D: D_KEYS, MOD_ED: ED_GRAD_KEYS}


class _Config(ctypes.Structure):
    _fields_ = [("batch", ctypes.c_int), ("precision", ctypes.c_int), ("max_notes", ctypes.c_int),
                ("note_dim", ctypes.c_int), ("noise_dim", ctypes.c_int), ("latent_dim", ctypes.c_int),
                ("gen_hidden", ctypes.c_int), ("numeric_dim", ctypes.c_int), ("enc_hidden1", ctypes.c_int),
                ("enc_hidden2", ctypes.c_int), ("embed_dim", ctypes.c_int), ("n_classes", ctypes.c_int),
                ("enc_dropout", ctypes.c_double), ("lambda_gp", ctypes.c_double), ("lambda_emotion", ctypes.c_double),
                ("bn_momentum", ctypes.c_double), ("bn_eps", ctypes.c_double), ("cond_dim", ctypes.c_int)]


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _check_f32_cuda(t, name, shape=None):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise ValueError(f"{name}: expected a contiguous float32 CUDA tensor")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t


class GanEngine:
    """One native context for a fixed per-rank batch size and precision ('fp32' | 'bf16')."""

    def __init__(self, batch, precision="fp32", max_notes=512, note_dim=4, noise_dim=128, latent_dim=64,
                 gen_hidden=512, numeric_dim=6, enc_hidden=(256, 128), embed_dim=128, n_classes=4, enc_dropout=0.2,
                 lambda_gp=10.0, lambda_emotion=5.0, bn_momentum=0.1, bn_eps=1e-5, device=None, cond_dim=0):
        if not torch.cuda.is_available():
            raise RuntimeError("melogan_b200 needs a CUDA (sm_100a) device; there is no CPU path")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.B, self.T, self.note_dim = int(batch), int(max_notes), int(note_dim)
        self.noise_dim, self.latent_dim, self.embed_dim = int(noise_dim), int(latent_dim), int(embed_dim)
        self.numeric_dim, self.enc_hidden, self.n_classes = int(numeric_dim), tuple(enc_hidden), int(n_classes)
        self.precision = precision
        self.cond_dim = int(cond_dim)
        cfg = _Config(self.B, {"fp32": 0, "bf16": 1}[precision], self.T, self.note_dim, self.noise_dim, self.latent_dim,
                      int(gen_hidden), self.numeric_dim, int(enc_hidden[0]), int(enc_hidden[1]), self.embed_dim,
                      self.n_classes, float(enc_dropout), float(lambda_gp), float(lambda_emotion), float(bn_momentum),
                      float(bn_eps), self.cond_dim)
        self._h = ctypes.c_void_p()
        L = _native.lib()
        for name, (argtypes, restype) in _GAN_SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes, fn.restype = argtypes, restype
        with torch.cuda.device(self.device):
            _native.check(L.mg_gan_create(ctypes.byref(cfg), ctypes.byref(self._h)))
        self._keep = {}

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _native.lib().mg_gan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing ----
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _call(self, name, *args):
        with torch.cuda.device(self.device):
            _native.check(getattr(_native.lib(), name)(self._h, *args))

    def weight_cache(self, on=True):
        """Keep packed tensor-core weights between launches; only for callers whose parameters change through
        FusedAdam.step alone (include/melogan_b200.h: mg_gan_weight_cache)."""
        return bool(_native.lib().mg_gan_weight_cache(self._h, int(on)))

    def workspace_bytes(self):
        return int(_native.lib().mg_gan_workspace_bytes(self._h))

    def buffer(self, name, dtype=torch.float32):
        """Debug view of a named workspace buffer (copied out)."""
        p, n = ctypes.c_void_p(), ctypes.c_longlong()
        self._call("mg_gan_buffer", name.encode(), ctypes.byref(p), ctypes.byref(n))
        out = torch.empty(n.value // torch.empty((), dtype=dtype).element_size(), dtype=dtype, device=self.device)
        with torch.cuda.device(self.device):
            _native.call("mg_device_copy", out.data_ptr(), p.value, n.value, self._stream())
        return out

    def bind(self, module, params, grads=None):
        """params/grads: dict key -> CUDA float32 tensor using the reference's state_dict keys."""
        pk, gk = PARAM_KEYS[module], GRAD_KEYS[module]
        pt = [_check_f32_cuda(params[k], k) for k in pk]
        gt = [_check_f32_cuda(grads[k], "grad " + k, params[k].shape) for k in gk] if grads is not None else None
        parr = (ctypes.c_void_p * len(pt))(*[t.data_ptr() for t in pt])
        garr = (ctypes.c_void_p * len(gt))(*[t.data_ptr() for t in gt]) if gt else None
        self._call("mg_gan_bind", module, parr, len(pt), garr, len(gt) if gt else 0)
        self._keep[module] = (pt, gt)

    # ---- A-1 ----
    def encoder_forward(self, numeric, mask1=None, mask2=None, train=True, out=None):
        _check_f32_cuda(numeric, "numeric", (self.B, self.numeric_dim))
        out = out if out is not None else torch.empty((self.B, self.embed_dim), device=self.device)
        if train:
            _check_f32_cuda(mask1, "mask1", (self.B, self.enc_hidden[0]))
            _check_f32_cuda(mask2, "mask2", (self.B, self.enc_hidden[1]))
        self._keep["e_in"] = (numeric, mask1, mask2)
        self._call("mg_feature_encoder_forward", _ptr(numeric), _ptr(mask1) if train else None,
                   _ptr(mask2) if train else None, int(train), _ptr(out), self._stream())
        return out

    def encoder_backward(self, demb):
        _check_f32_cuda(demb, "demb", (self.B, self.embed_dim))
        self._call("mg_feature_encoder_backward", _ptr(demb), self._stream())

    # ---- SyncBatchNorm over NVLink peer memory ----
    def sync_bn_connect(self, process_group=None):
        """Collective: from here on the generator's two BatchNorm layers use batch statistics over ALL ranks of the group (one
        node, <= 8 GPUs, one process per GPU), exchanged by a one-CTA kernel through CUDA-IPC peer buffers -- no NCCL call in
        the step, so the captured step graphs stay valid.  The handles travel once, through torch.distributed."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        h = (ctypes.c_ubyte * 64)()
        self._call("mg_gan_sync_bn_export", rank, world, ctypes.cast(h, ctypes.c_void_p))
        mine = torch.tensor(list(bytes(h)), dtype=torch.uint8, device=self.device)
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine, group=process_group)
        blob = bytes(torch.cat(allh).cpu().tolist())
        buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._call("mg_gan_sync_bn_connect", ctypes.cast(buf, ctypes.c_void_p))
        dist.barrier(group=process_group)          # every rank has opened every buffer before the first sync point
        self.sync_bn = True

    # ---- A-2..A-4 ----
    def set_condition(self, encoder_latent):
        """'conditioning' mode: the AE latent block of G's input row, read by every later generator forward / fused step."""
        _check_f32_cuda(encoder_latent, "encoder_latent", (self.B, self.cond_dim))
        self._keep["cond"] = encoder_latent
        self._call("mg_generator_set_condition", _ptr(encoder_latent))

    def generator_forward(self, noise, emb, train=True):
        _check_f32_cuda(noise, "noise", (self.B, self.noise_dim))
        _check_f32_cuda(emb, "numeric_embedding", (self.B, self.embed_dim))
        notes = torch.empty((self.B, self.T, self.note_dim), device=self.device)
        latent = torch.empty((self.B, self.latent_dim), device=self.device)
        self._call("mg_generator_forward", _ptr(noise), _ptr(emb), int(train), _ptr(notes), _ptr(latent), self._stream())
        return notes, latent

    def generator_backward(self, dnotes, dlatent=None):
        _check_f32_cuda(dnotes, "dnotes", (self.B, self.T, self.note_dim))
        if dlatent is not None:
            _check_f32_cuda(dlatent, "dlatent", (self.B, self.latent_dim))
        demb = torch.empty((self.B, self.embed_dim), device=self.device)
        self._call("mg_generator_backward", _ptr(dnotes), _ptr(dlatent), _ptr(demb), self._stream())
        return demb

    # ---- A-5 ----
    def critic_forward(self, notes, emb=None):
        R = notes.shape[0]
        _check_f32_cuda(notes, "notes", (R, self.T, self.note_dim))
        if emb is not None:
            _check_f32_cuda(emb, "numeric_embedding", (self.B, self.embed_dim))
        score = torch.empty(R, device=self.device)
        self._keep["d_in"] = (notes, emb)
        self._call("mg_discriminator_forward", _ptr(notes), _ptr(emb), R, _ptr(score), self._stream())
        return score

    def critic_backward(self, dscore, param_grads=False, want_dnotes=True, want_demb=False):
        notes, emb = self._keep["d_in"]
        R = notes.shape[0]
        _check_f32_cuda(dscore, "dscore", (R,))
        dnotes = torch.empty_like(notes) if want_dnotes else None
        demb = torch.empty((R, self.embed_dim), device=self.device) if (want_demb and emb is not None) else None
        if param_grads:
            self._call("mg_discriminator_backward_ex", _ptr(notes), _ptr(dscore), _ptr(dnotes), _ptr(demb), self._stream())
        else:
            self._call("mg_discriminator_backward", _ptr(dscore), 0, _ptr(dnotes), _ptr(demb), self._stream())
        return dnotes, demb

    # ---- A-6/A-7 ----
    def critic_loss_backward(self, real, fake, emb, alpha, metrics=None):
        _check_f32_cuda(real, "real", (self.B, self.T, self.note_dim))
        _check_f32_cuda(fake, "fake", (self.B, self.T, self.note_dim))
        _check_f32_cuda(alpha, "alpha", (self.B,))
        metrics = metrics if metrics is not None else torch.empty(4, device=self.device)
        self._call("mg_critic_loss_backward", _ptr(real), _ptr(fake), _ptr(emb), _ptr(alpha), _ptr(metrics), self._stream())
        return metrics

    def gradient_penalty(self, real, fake, emb, alpha, metrics=None):
        """GP value (metrics[1]) and d(GP)/d(theta_D) added into the bound critic grads."""
        _check_f32_cuda(real, "real", (self.B, self.T, self.note_dim))
        _check_f32_cuda(fake, "fake", (self.B, self.T, self.note_dim))
        _check_f32_cuda(alpha, "alpha", (self.B,))
        metrics = metrics if metrics is not None else torch.empty(4, device=self.device)
        self._call("mg_gradient_penalty", _ptr(real), _ptr(fake), _ptr(emb), _ptr(alpha), _ptr(metrics), self._stream())
        return metrics

    # ---- A-8 ----
    def emotion_forward(self, notes):
        _check_f32_cuda(notes, "notes", (self.B, self.T, self.note_dim))
        logits = torch.empty((self.B, self.n_classes), device=self.device)
        self._call("mg_emotion_forward", _ptr(notes), _ptr(logits), self._stream())
        return logits

    def emotion_backward_input(self, dlogits, out=None, accumulate=False):
        _check_f32_cuda(dlogits, "dlogits", (self.B, self.n_classes))
        out = out if out is not None else torch.empty((self.B, self.T, self.note_dim), device=self.device)
        self._call("mg_emotion_backward_input", _ptr(dlogits), _ptr(out), int(accumulate), self._stream())
        return out

    # ---- A-13 ----
    def emotion_train_forward(self, notes, mask1, mask2, dropout_p=0.2):
        _check_f32_cuda(notes, "notes", (self.B, self.T, self.note_dim))
        logits = torch.empty((self.B, self.n_classes), device=self.device)
        self._keep["ed_train"] = (notes, mask1, mask2)
        self._call("mg_emotion_train_forward", _ptr(notes), _ptr(mask1), _ptr(mask2), float(dropout_p), _ptr(logits),
                   self._stream())
        return logits

    def emotion_train_backward(self, dlogits, want_dnotes=False):
        notes = self._keep["ed_train"][0]
        _check_f32_cuda(dlogits, "dlogits", (self.B, self.n_classes))
        dnotes = torch.empty_like(notes) if want_dnotes else None
        self._call("mg_emotion_train_backward", _ptr(notes), _ptr(dlogits), _ptr(dnotes), self._stream())
        return dnotes

    # ---- composites ----
    def critic_step(self, real, numeric, noise, alpha, mask1, mask2, metrics=None):
        _check_f32_cuda(real, "real", (self.B, self.T, self.note_dim))
        _check_f32_cuda(numeric, "numeric", (self.B, self.numeric_dim))
        _check_f32_cuda(noise, "noise", (self.B, self.noise_dim))
        _check_f32_cuda(alpha, "alpha", (self.B,))
        _check_f32_cuda(mask1, "mask1", (self.B, self.enc_hidden[0]))
        _check_f32_cuda(mask2, "mask2", (self.B, self.enc_hidden[1]))
        metrics = metrics if metrics is not None else torch.empty(4, device=self.device)
        self._call("mg_critic_step", _ptr(real), _ptr(numeric), _ptr(noise), _ptr(alpha), _ptr(mask1), _ptr(mask2),
                   _ptr(metrics), self._stream())
        return metrics

    def generator_step(self, numeric, noise, labels, mask1, mask2, metrics=None):
        _check_f32_cuda(numeric, "numeric", (self.B, self.numeric_dim))
        _check_f32_cuda(noise, "noise", (self.B, self.noise_dim))
        if not (labels.is_cuda and labels.dtype == torch.int64 and tuple(labels.shape) == (self.B,)):
            raise ValueError("labels: expected int64 CUDA tensor of shape (B,)")
        metrics = metrics if metrics is not None else torch.empty(2, device=self.device)
        self._call("mg_generator_step", _ptr(numeric), _ptr(noise), _ptr(labels), _ptr(mask1), _ptr(mask2), _ptr(metrics),
                   self._stream())
        return metrics


def rng_fill(out, kind, seed, offset=0, p=0.0):
    """kind: 'normal' | 'uniform' | 'bernoulli' (keep probability p); counter-based, reproducible."""
    _check_f32_cuda(out, "out")
    L = _native.lib()
    L.mg_rng_fill.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_float, ctypes.c_ulonglong,
                              ctypes.c_ulonglong, ctypes.c_void_p]
    L.mg_rng_fill.restype = ctypes.c_int
    with torch.cuda.device(out.device):
        _native.check(L.mg_rng_fill(_ptr(out), out.numel(), {"normal": 0, "uniform": 1, "bernoulli": 2}[kind], float(p),
                                    int(seed), int(offset), ctypes.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)))
    return out


_vp, _i, _ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
_GAN_SIGNATURES = {
    "mg_gan_create": ([ctypes.POINTER(_Config), ctypes.POINTER(ctypes.c_void_p)], _i),
    "mg_gan_destroy": ([_vp], None),
    "mg_gan_workspace_bytes": ([_vp], _ll),
    "mg_gan_buffer": ([_vp, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(_ll)], _i),
    "mg_gan_bind": ([_vp, _i, _vp, _i, _vp, _i], _i),
    "mg_feature_encoder_forward": ([_vp, _vp, _vp, _vp, _i, _vp, _vp], _i),
    "mg_feature_encoder_backward": ([_vp, _vp, _vp], _i),
    "mg_generator_forward": ([_vp, _vp, _vp, _i, _vp, _vp, _vp], _i),
    "mg_generator_set_condition": ([_vp, _vp], _i),
    "mg_generator_backward": ([_vp, _vp, _vp, _vp, _vp], _i),
    "mg_discriminator_forward": ([_vp, _vp, _vp, _i, _vp, _vp], _i),
    "mg_discriminator_backward": ([_vp, _vp, _i, _vp, _vp, _vp], _i),
    "mg_discriminator_backward_ex": ([_vp, _vp, _vp, _vp, _vp, _vp], _i),
    "mg_critic_loss_backward": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "mg_gradient_penalty": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "mg_emotion_forward": ([_vp, _vp, _vp, _vp], _i),
    "mg_emotion_backward_input": ([_vp, _vp, _vp, _i, _vp], _i),
    "mg_emotion_train_forward": ([_vp, _vp, _vp, _vp, ctypes.c_double, _vp, _vp], _i),
    "mg_emotion_train_backward": ([_vp, _vp, _vp, _vp, _vp], _i),
    "mg_critic_step": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "mg_generator_step": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp], _i),
}


# ------------------------------------------------------------------------------------------------
# A-12: VAE context (BASELINE config #2)
# ------------------------------------------------------------------------------------------------
VAE_PARAM_KEYS = []
for _c, _b in ((0, 1), (3, 4), (6, 7)):
    VAE_PARAM_KEYS += [f"encoder.conv.{_c}.weight", f"encoder.conv.{_c}.bias", f"encoder.conv.{_b}.weight",
                       f"encoder.conv.{_b}.bias"]
VAE_PARAM_KEYS += ["encoder._linear.1.weight", "encoder._linear.1.bias", "fc_mu.weight", "fc_mu.bias",
                   "fc_log_var.weight", "fc_log_var.bias", "decoder.pre.0.weight", "decoder.pre.0.bias",
                   "decoder.pre.2.weight", "decoder.pre.2.bias", "decoder.deconv.0.weight", "decoder.deconv.0.bias",
                   "decoder.deconv.1.weight", "decoder.deconv.1.bias", "decoder.deconv.3.weight", "decoder.deconv.3.bias",
                   "decoder.deconv.4.weight", "decoder.deconv.4.bias", "decoder.deconv.6.weight", "decoder.deconv.6.bias"]
VAE_BUFFER_KEYS = []
for _m in ("encoder.conv.1", "encoder.conv.4", "encoder.conv.7", "decoder.deconv.1", "decoder.deconv.4"):
    VAE_BUFFER_KEYS += [_m + ".running_mean", _m + ".running_var"]


class VaeEngine:
    def __init__(self, batch, max_notes=512, latent_dim=8, precision="fp32", device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("melogan_b200 needs a CUDA (sm_100a) device; there is no CPU path")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.B, self.T, self.latent = int(batch), int(max_notes), int(latent_dim)
        self._h = ctypes.c_void_p()
        L = _native.lib()
        with torch.cuda.device(self.device):
            _native.check(L.mg_vae_create(self.B, self.T, self.latent, {"fp32": 0, "bf16": 1}[precision], ctypes.byref(self._h)))
        self._keep = {}

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _native.lib().mg_vae_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _call(self, name, *args):
        with torch.cuda.device(self.device):
            _native.check(getattr(_native.lib(), name)(self._h, *args))

    def bind(self, params, grads=None):
        pt = [_check_f32_cuda(params[k], k) for k in VAE_PARAM_KEYS + VAE_BUFFER_KEYS]
        gt = [_check_f32_cuda(grads[k], "grad " + k, params[k].shape) for k in VAE_PARAM_KEYS] if grads is not None else None
        parr = (ctypes.c_void_p * len(pt))(*[t.data_ptr() for t in pt])
        garr = (ctypes.c_void_p * len(gt))(*[t.data_ptr() for t in gt]) if gt else None
        self._call("mg_vae_bind", parr, len(pt), garr, len(gt) if gt else 0)
        self._keep["bind"] = (pt, gt)

    def forward(self, x, eps, train=True):
        _check_f32_cuda(x, "x", (self.B, self.T, 4))
        _check_f32_cuda(eps, "eps", (self.B, self.latent))
        recon = torch.empty_like(x)
        z, mu, lv = (torch.empty((self.B, self.latent), device=self.device) for _ in range(3))
        self._keep["fwd"] = (x, eps)
        self._call("mg_vae_forward", _ptr(x), _ptr(eps), int(train), _ptr(recon), _ptr(z), _ptr(mu), _ptr(lv), self._stream())
        return recon, z, mu, lv

    def backward(self, drecon, dz=None, dmu=None, dlogvar=None):
        x = self._keep["fwd"][0]
        _check_f32_cuda(drecon, "drecon", (self.B, self.T, 4))
        self._call("mg_vae_backward", _ptr(x), _ptr(drecon), _ptr(dz), _ptr(dmu), _ptr(dlogvar), self._stream())

    def buffer(self, name, dtype=torch.float32):
        """Debug view of a named workspace buffer (copied out)."""
        p, n = ctypes.c_void_p(), ctypes.c_longlong()
        self._call("mg_vae_buffer", name.encode(), ctypes.byref(p), ctypes.byref(n))
        out = torch.empty(n.value // torch.empty((), dtype=dtype).element_size(), dtype=dtype, device=self.device)
        with torch.cuda.device(self.device):
            _native.call("mg_device_copy", out.data_ptr(), p.value, n.value, self._stream())
        return out

    def loss_step(self, x, eps, beta, metrics=None):
        _check_f32_cuda(x, "x", (self.B, self.T, 4))
        _check_f32_cuda(eps, "eps", (self.B, self.latent))
        metrics = metrics if metrics is not None else torch.empty(3, device=self.device)
        self._keep["fwd"] = (x, eps)
        self._call("mg_vae_loss_step", _ptr(x), _ptr(eps), float(beta), _ptr(metrics), self._stream())
        return metrics
