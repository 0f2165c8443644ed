"""ctypes front-end of oracle/notes_oracle.c (TEST INFRASTRUCTURE ONLY -- see the C file's header)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# utils.py:14-26 of the reference (interval tables are music-theory constants)
SCALE_INTERVALS = {
    "major": (0, 2, 4, 5, 7, 9, 11), "minor": (0, 2, 3, 5, 7, 8, 10), "chromatic": tuple(range(12)),
    "dorian": (0, 2, 3, 5, 7, 9, 10), "phrygian": (0, 1, 3, 5, 7, 8, 10), "lydian": (0, 2, 4, 6, 7, 9, 11),
    "mixolydian": (0, 2, 4, 5, 7, 9, 10), "locrian": (0, 1, 3, 5, 6, 8, 10),
    "major_pentatonic": (0, 2, 4, 7, 9), "minor_pentatonic": (0, 3, 5, 7, 10), "blues": (0, 3, 5, 6, 7, 10),
}


def allowed_mask(scale, root_key):
    iv = SCALE_INTERVALS.get(scale, SCALE_INTERVALS["chromatic"])
    m = 0
    for i in iv:
        m |= 1 << ((i + root_key) % 12)
    return m


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libnotes_oracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        f32p, i32p, f64p = (ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32),
                            ctypes.POINTER(ctypes.c_double))
        L.orc_extract_notes_gan_batch.argtypes = [f32p, ctypes.c_longlong, ctypes.c_int, ctypes.c_double,
                                                  ctypes.c_uint32, i32p, i32p, i32p, f64p, f64p]
        L.orc_extract_notes_gan_batch.restype = ctypes.c_int
        L.orc_extract_notes_abs_batch.argtypes = [f32p, ctypes.c_longlong, ctypes.c_int, i32p, i32p, f64p, f64p]
        L.orc_extract_notes_abs_batch.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def extract_notes_gan(rolls, bpm=120.0, scale="major", root_key=0):
    """rolls (R, T, 4) float32 -> (bad, counts (R,), pitch, velocity (R,T) int32, start, end (R,T) float64)."""
    rolls = np.ascontiguousarray(rolls, dtype=np.float32)
    R, T, _ = rolls.shape
    counts = np.zeros(R, np.int32)
    pitch = np.zeros((R, T), np.int32); vel = np.zeros((R, T), np.int32)
    start = np.zeros((R, T), np.float64); end = np.zeros((R, T), np.float64)
    bad = lib().orc_extract_notes_gan_batch(_p(rolls, ctypes.c_float), R, T, float(bpm),
                                            allowed_mask(scale, root_key), _p(counts, ctypes.c_int32),
                                            _p(pitch, ctypes.c_int32), _p(vel, ctypes.c_int32),
                                            _p(start, ctypes.c_double), _p(end, ctypes.c_double))
    return bad, counts, pitch, vel, start, end


def extract_notes_abs(rolls):
    rolls = np.ascontiguousarray(rolls, dtype=np.float32)
    R, T, _ = rolls.shape
    pitch = np.zeros((R, T), np.int32); vel = np.zeros((R, T), np.int32)
    start = np.zeros((R, T), np.float64); end = np.zeros((R, T), np.float64)
    bad = lib().orc_extract_notes_abs_batch(_p(rolls, ctypes.c_float), R, T, _p(pitch, ctypes.c_int32),
                                            _p(vel, ctypes.c_int32), _p(start, ctypes.c_double),
                                            _p(end, ctypes.c_double))
    return bad, pitch, vel, start, end


def digest(counts, pitch, vel, start, end):
    """sha256 over the valid prefix of every roll (the bit-exactness checksum of checksums)."""
    import hashlib
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(counts, np.int32).tobytes())
    for r, c in enumerate(counts):
        c = max(int(c), 0)
        h.update(np.ascontiguousarray(pitch[r, :c], np.int32).tobytes())
        h.update(np.ascontiguousarray(vel[r, :c], np.int32).tobytes())
        h.update(np.ascontiguousarray(start[r, :c], np.float64).tobytes())
        h.update(np.ascontiguousarray(end[r, :c], np.float64).tobytes())
    return h.hexdigest()


def ae_normalize(notes, max_start_beat=100.0, max_duration_beat=20.0):
    """numpy restatement of MIDIDataset.__getitem__ without augmentation (reference src/ae/dataset.py:68-89,105-106);
    notes (..., T, 4) raw rows (pitch, start, duration, velocity), float32 arithmetic as numpy does it there."""
    notes = np.array(notes, dtype=np.float32, copy=True)
    flat = notes.reshape(-1, 4)
    mask = flat[:, 0] != -1
    flat[mask, 0] = (flat[mask, 0] / 128.0) * 2.0 - 1.0
    flat[mask, 3] = np.clip(flat[mask, 3], 0, 127)
    flat[mask, 3] = (flat[mask, 3] / 128.0) * 2.0 - 1.0
    flat[mask, 1] = flat[mask, 1] / max_start_beat
    flat[mask, 2] = flat[mask, 2] / max_duration_beat
    return np.nan_to_num(flat, nan=0.0, posinf=0.0, neginf=0.0).astype(np.float32).reshape(notes.shape)
