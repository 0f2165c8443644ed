"""ctypes binding of libmelogan_b200.so (the C ABI declared in include/melogan_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the
caller gets an exception -- nothing on the product path is computed on the CPU.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmelogan_b200.so")

MG_OK, MG_ERR_INVALID, MG_ERR_CUDA, MG_ERR_NONFINITE, MG_ERR_STATE = 0, -1, -2, -3, -4


class MeloGanNativeError(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"melogan_b200 native call failed (status {status}): {text}")
        self.status = status


_lib = None
_vp, _ll, _i, _d, _f, _u32 = (ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_double, ctypes.c_float,
                              ctypes.c_uint32)

_SIGNATURES = {
    "mg_last_error": ([], ctypes.c_char_p),
    "mg_abi_version": ([], _i),
    "mg_build_info": ([], ctypes.c_char_p),
    "mg_scale_mask": ([ctypes.c_char_p, _i], _u32),
    "mg_extract_notes_gan": ([_vp, _ll, _i, _d, _u32, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "mg_extract_notes_gan_host": ([_vp, _ll, _i, _d, _u32, _vp, _vp, _vp, _vp, _vp], _i),
    "mg_extract_notes_abs": ([_vp, _ll, _i, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "mg_extract_notes_abs_host": ([_vp, _ll, _i, _vp, _vp, _vp, _vp], _i),
    "mg_ae_normalize": ([_vp, _vp, _ll, _i, _d, _d, _vp], _i),
    "mg_adam_step": ([_vp, _vp, _vp, _vp, _ll, _d, _d, _d, _d, _d, _i, _f, _ll, _vp, _vp, _vp], _i),
    "mg_adam_step_clipped": ([_vp, _vp, _vp, _vp, _ll, _d, _d, _d, _d, _d, _i, _f, _f, _vp, _ll, _vp, _vp, _vp], _i),
    "mg_launch_count": ([], _ll),
    "mg_probe_begin": ([_i], _i),
    "mg_probe_end": ([_vp], _i),
    "mg_tc_enable": ([_i], _i),
    "mg_gan_weight_cache": ([_vp, _i], _i),
    "mg_weight_cache_invalidate": ([], _i),
    "mg_device_copy": ([_vp, _vp, _ll, _vp], _i),
    "mg_rng_fill": ([_vp, _ll, _i, _f, ctypes.c_ulonglong, ctypes.c_ulonglong, _vp], _i),
    "mg_rng_fill_counter": ([_vp, _ll, _i, _f, ctypes.c_ulonglong, _vp, ctypes.c_ulonglong, _vp], _i),
    "mg_counter_add": ([_vp, ctypes.c_ulonglong, _vp], _i),
    # mg_gan_* (argtypes are refined by melogan.engine, which owns the config struct)
    "mg_gan_create": ([_vp, _vp], _i),
    "mg_gan_destroy": ([_vp], None),
    "mg_gan_workspace_bytes": ([_vp], _ll),
    "mg_gan_buffer": ([_vp, ctypes.c_char_p, _vp, _vp], _i),
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
    "mg_emotion_train_forward": ([_vp, _vp, _vp, _vp, _d, _vp, _vp], _i),
    "mg_emotion_train_backward": ([_vp, _vp, _vp, _vp, _vp], _i),
    "mg_cross_entropy": ([_vp, _vp, _i, _i, _vp, _vp, _vp], _i),
    "mg_vae_create": ([_i, _i, _i, _i, _vp], _i),
    "mg_vae_destroy": ([_vp], None),
    "mg_vae_bind": ([_vp, _vp, _i, _vp, _i], _i),
    "mg_vae_forward": ([_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp], _i),
    "mg_vae_backward": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "mg_vae_buffer": ([_vp, ctypes.c_char_p, _vp, _vp], _i),
    "mg_vae_loss_step": ([_vp, _vp, _vp, _d, _vp, _vp], _i),
    "mg_peer_create": ([_i, _i, _ll, _vp, _vp], _i),
    "mg_peer_connect": ([_vp, _vp], _i),
    "mg_peer_allreduce_sum": ([_vp, _vp, _ll, _vp], _i),
    "mg_peer_destroy": ([_vp], None),
    "mg_gan_sync_bn_export": ([_vp, _i, _i, _vp], _i),
    "mg_gan_sync_bn_connect": ([_vp, _vp], _i),
    "mg_convunit_create": ([_vp], _i),
    "mg_convunit_destroy": ([_vp], None),
    "mg_convunit_forward": ([_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp,
                             _vp, _vp], _i),
    "mg_convunit_backward": ([_vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                              _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "mg_linear_forward": ([_vp, _vp, _vp, _vp, _i, _i, _i, _vp], _i),
    "mg_linear_backward": ([_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp], _i),
    "mg_act_dropout_forward": ([_vp, _vp, _f, _i, _vp, _ll, _vp], _i),
    "mg_act_dropout_backward": ([_vp, _vp, _vp, _f, _i, _vp, _ll, _vp], _i),
    # per-layer harness (tests/tc_layers.py owns the mg_debug_layer struct)
    "mg_debug_layer_run": ([_vp, _vp], _i),
    "mg_debug_set": ([ctypes.c_char_p, _i], _i),
    "mg_debug_last_launch": ([], ctypes.c_char_p),
    "mg_critic_step": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "mg_generator_step": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp], _i),
}


def lib():
    """Loads the shared library once; raises if it has not been built (see __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MeloGanNativeError(MG_ERR_STATE, f"{LIB_PATH} not built; run `make -C melo-gan_b200/csrc`")
        L = ctypes.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes, fn.restype = argtypes, restype
        _lib = L
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(status):
    if status != MG_OK:
        raise MeloGanNativeError(status, lib().mg_last_error().decode())


def call(name, *args):
    check(getattr(lib(), name)(*args))
