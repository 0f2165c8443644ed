"""Process-wide glue between the drop-in nn.Modules and the native engine.

A module forward needs a native context sized for its batch; contexts are cached per
(device, batch, precision, shape signature).  Modules bind their parameter tensors before every call
(a pointer table; cheap), so load_state_dict / .to() / optimizer steps are always seen.

Autograd: every module forward is one torch.autograd.Function.  Backward RE-RUNS the forward into the
context's workspace (activation recomputation) before the native backward, so interleaved calls such
as D(real), D(fake), D(interp) followed by one loss.backward() -- the reference's critic step at
src/gan/train_gan.py:191-203 -- are always correct.  The fused step entry points used by
melogan.trainer avoid that recomputation; this path is the faithful drop-in one.
"""
import os

import torch

from . import engine as E

_ENGINES = {}
PRECISION = os.environ.get("MELOGAN_PRECISION", "fp32")   # 'fp32' (parity) | 'bf16' (tensor cores)


def set_precision(p):
    global PRECISION
    if p not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    PRECISION = p


def set_fp32_tensor_cores(on):
    """fp32 mode's contractions as six bf16 tensor-core terms per float32 product (operands split exactly into three bf16
    parts) instead of CUDA-core FFMA.  True / False, or None to follow MELOGAN_FP32_TC (default: off).
    Process-wide, like the kernel-selection switches of mg_debug_set."""
    import ctypes
    from . import _native
    lib = _native.lib()
    lib.mg_debug_set.argtypes = [ctypes.c_char_p, ctypes.c_int]
    lib.mg_debug_set.restype = ctypes.c_int
    _native.check(lib.mg_debug_set(b"fp32_tc", -1 if on is None else (1 if on else 0)))


def engine_for(device, batch, **shape):
    if torch.device(device).type != "cuda":
        raise RuntimeError("melogan_b200 modules run on CUDA (sm_100a) only; there is no CPU fallback. "
                           "Move the module and its inputs to a CUDA device.")
    key = (str(device), int(batch), PRECISION, tuple(sorted(shape.items())))
    eng = _ENGINES.get(key)
    if eng is None:
        eng = E.GanEngine(batch, precision=PRECISION, device=device, **shape)
        _ENGINES[key] = eng
    return eng


def clear_engines():
    for e in _ENGINES.values():
        e.close()
    _ENGINES.clear()


def params_of(module, keys):
    sd = dict(module.named_parameters())
    sd.update(dict(module.named_buffers()))
    return {k: sd[k].data for k in keys}


def fresh_grads(params, keys):
    return {k: torch.zeros_like(params[k]) for k in keys}


def as_f32c(t):
    return t.detach().to(torch.float32).contiguous()
