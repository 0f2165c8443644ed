"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/melogan_b200.h declares (no compute calls: there is no GPU in the build container)."""
import ctypes
import os
import re

from melogan import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "melogan_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(_native.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 9
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/melogan_b200.h but not exported"


def test_python_binding_covers_the_header():
    assert set(_declared_symbols()) == set(_native.exported_symbols())


def test_invalid_arguments_fail_loudly_without_a_gpu():
    L = _native.lib()
    assert L.mg_abi_version() == 1
    assert b"sm_100a" in L.mg_build_info()
    st = L.mg_extract_notes_gan(None, 4, 512, 120.0, 0xFFF, None, None, None, None, None, None)
    assert st == _native.MG_ERR_INVALID and b"null" in L.mg_last_error()
    st = L.mg_extract_notes_gan(None, 4, 513, 120.0, 0xFFF, None, None, None, None, None, None)
    assert st == _native.MG_ERR_INVALID


def test_scale_mask_matches_reference_tables():
    from melogan.notes import SCALES, scale_mask
    for name, iv in SCALES.items():
        for root in range(12):
            want = 0
            for i in iv:
                want |= 1 << ((i + root) % 12)
            assert scale_mask(name, root) == want
    assert scale_mask("no_such_scale", 3) == 0xFFF


def test_kernel_selection_switches_are_host_state():
    """mg_debug_set only records a choice (no device needed): known keys succeed, unknown keys fail loudly; the fp32-on-
    tensor-cores switch is reachable from the Python runtime."""
    from melogan import runtime
    L = _native.lib()
    L.mg_debug_set.argtypes = [ctypes.c_char_p, ctypes.c_int]
    L.mg_debug_set.restype = ctypes.c_int
    for on in (True, False, None):
        runtime.set_fp32_tensor_cores(on)
    assert L.mg_debug_set(b"no_such_switch", 1) == _native.MG_ERR_INVALID and b"unknown key" in L.mg_last_error()
    assert L.mg_debug_set(b"reset", 0) == 0
