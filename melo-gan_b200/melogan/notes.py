"""Batched note extraction on the GPU (host side of mg_extract_notes_*).

Replaces the per-row Python loops of the reference:
  N-1  save_piano_roll_to_midi     src/gan/utils.py:130-155   -> extract_notes_gan
  N-2  tools/roll_to_midi.py:10-21                             -> extract_notes_abs
Column convention (pitch, velocity, duration, step|start) as in src/gan/utils.py:131.
"""
import ctypes
from dataclasses import dataclass

import numpy as np
import torch

from . import _native

# reference src/gan/utils.py:14-26 (interval tables; the C side holds the same table for mg_scale_mask)
SCALES = {
    "major": [0, 2, 4, 5, 7, 9, 11], "minor": [0, 2, 3, 5, 7, 8, 10], "chromatic": list(range(12)),
    "dorian": [0, 2, 3, 5, 7, 9, 10], "phrygian": [0, 1, 3, 5, 7, 8, 10], "lydian": [0, 2, 4, 6, 7, 9, 11],
    "mixolydian": [0, 2, 4, 5, 7, 9, 10], "locrian": [0, 1, 3, 5, 6, 8, 10],
    "major_pentatonic": [0, 2, 4, 7, 9], "minor_pentatonic": [0, 3, 5, 7, 10], "blues": [0, 3, 5, 6, 7, 10],
}


@dataclass
class NoteBatch:
    """First counts[r] entries of row r are that roll's notes in emission order."""
    counts: object
    pitch: object
    velocity: object
    start: object
    end: object

    def notes_of(self, r):
        n = int(self.counts[r])
        return [(int(self.velocity[r, i]), int(self.pitch[r, i]), float(self.start[r, i]), float(self.end[r, i]))
                for i in range(n)]


def scale_mask(scale, root_key):
    return int(_native.lib().mg_scale_mask(str(scale).encode(), int(root_key)))


def _require_cuda_rolls(rolls):
    if not (isinstance(rolls, torch.Tensor) and rolls.is_cuda):
        raise ValueError("rolls must be a CUDA tensor (there is no CPU path; use *_host for numpy input)")
    if rolls.dtype != torch.float32 or rolls.dim() != 3 or rolls.size(2) != 4:
        raise ValueError(f"rolls must be float32 (R, T, 4), got {rolls.dtype} {tuple(rolls.shape)}")
    return rolls.contiguous()


def extract_notes_gan(rolls, bpm=120.0, scale="major", root_key=0, check=True):
    rolls = _require_cuda_rolls(rolls)
    R, T, _ = rolls.shape
    dev = rolls.device
    out = NoteBatch(torch.empty(R, dtype=torch.int32, device=dev),
                    torch.empty((R, T), dtype=torch.uint8, device=dev),
                    torch.empty((R, T), dtype=torch.uint8, device=dev),
                    torch.empty((R, T), dtype=torch.float64, device=dev),
                    torch.empty((R, T), dtype=torch.float64, device=dev))
    with torch.cuda.device(dev):
        _native.call("mg_extract_notes_gan", rolls.data_ptr(), R, T, float(bpm), scale_mask(scale, root_key),
                     out.counts.data_ptr(), out.pitch.data_ptr(), out.velocity.data_ptr(), out.start.data_ptr(),
                     out.end.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    if check and R and bool((out.counts < 0).any()):
        raise ValueError("cannot convert float NaN/inf to integer (non-finite pitch or velocity in an ungated row)")
    return out


def extract_notes_abs(rolls, check=True):
    rolls = _require_cuda_rolls(rolls)
    R, T, _ = rolls.shape
    dev = rolls.device
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    out = NoteBatch(torch.full((R,), T, dtype=torch.int32, device=dev),
                    torch.empty((R, T), dtype=torch.uint8, device=dev),
                    torch.empty((R, T), dtype=torch.uint8, device=dev),
                    torch.empty((R, T), dtype=torch.float64, device=dev),
                    torch.empty((R, T), dtype=torch.float64, device=dev))
    with torch.cuda.device(dev):
        _native.call("mg_extract_notes_abs", rolls.data_ptr(), R, T, out.pitch.data_ptr(), out.velocity.data_ptr(),
                     out.start.data_ptr(), out.end.data_ptr(), status.data_ptr(),
                     torch.cuda.current_stream(dev).cuda_stream)
    if check and int(status.item()):
        raise ValueError("cannot convert float NaN to integer")
    return out


def ae_normalize(notes, max_start_beat=100.0, max_duration_beat=20.0, out=None):
    """MIDIDataset.__getitem__ normalisation (reference src/ae/dataset.py:72-89,105) of a whole device-resident data set
    at once: notes (R, T, 4) float32 CUDA tensor of raw rows (pitch, start, duration, velocity); -1 padding rows are kept."""
    notes = _require_cuda_rolls(notes)
    out = torch.empty_like(notes) if out is None else out
    if not (out.is_cuda and out.dtype == torch.float32 and out.shape == notes.shape and out.is_contiguous()):
        raise ValueError("out must be a contiguous float32 CUDA tensor shaped like notes")
    with torch.cuda.device(notes.device):
        _native.call("mg_ae_normalize", notes.data_ptr(), out.data_ptr(), notes.shape[0], notes.shape[1],
                     float(max_start_beat), float(max_duration_beat), torch.cuda.current_stream(notes.device).cuda_stream)
    return out


def _np_ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def extract_notes_gan_host(rolls, bpm=120.0, scale="major", root_key=0):
    """numpy in / numpy out through the host-buffer C entry point (H2D + kernel + D2H inside the call)."""
    rolls = np.ascontiguousarray(rolls, dtype=np.float32)
    R, T, _ = rolls.shape
    out = NoteBatch(np.empty(R, np.int32), np.empty((R, T), np.uint8), np.empty((R, T), np.uint8),
                    np.empty((R, T), np.float64), np.empty((R, T), np.float64))
    st = _native.lib().mg_extract_notes_gan_host(_np_ptr(rolls), R, T, float(bpm), scale_mask(scale, root_key),
                                                 _np_ptr(out.counts), _np_ptr(out.pitch), _np_ptr(out.velocity),
                                                 _np_ptr(out.start), _np_ptr(out.end))
    if st == _native.MG_ERR_NONFINITE:
        raise ValueError(_native.lib().mg_last_error().decode())
    _native.check(st)
    return out


def extract_notes_abs_host(rolls):
    rolls = np.ascontiguousarray(rolls, dtype=np.float32)
    R, T, _ = rolls.shape
    out = NoteBatch(np.full(R, T, np.int32), np.empty((R, T), np.uint8), np.empty((R, T), np.uint8),
                    np.empty((R, T), np.float64), np.empty((R, T), np.float64))
    st = _native.lib().mg_extract_notes_abs_host(_np_ptr(rolls), R, T, _np_ptr(out.pitch), _np_ptr(out.velocity),
                                                 _np_ptr(out.start), _np_ptr(out.end))
    if st == _native.MG_ERR_NONFINITE:
        raise ValueError(_native.lib().mg_last_error().decode())
    _native.check(st)
    return out
