"""Generates tests/golden/ae_norm_golden.npz by running the REFERENCE's MIDIDataset.__getitem__ (src/ae/dataset.py:66-106,
augment=False) on small .npz files written to a temporary directory.

    python oracle/make_golden_ae_norm.py        (needs /root/reference; not available on the GPU box)

Stored: the raw rolls (input) and what the reference's dataset returned for them.
"""
import importlib.util
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("ref_ae_dataset", "/root/reference/src/ae/dataset.py")
ref_ds = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(ref_ds)


def raw_rolls(seed, n, T):
    """Raw rows (pitch, start beat, duration beat, velocity) as the preprocessing writes them, with the cases the reference's
    comments name: -1 padding rows, velocities above 127 ('your 194.0'), and a few non-finite values."""
    rng = np.random.default_rng(seed)
    x = np.empty((n, T, 4), np.float32)
    x[..., 0] = rng.integers(0, 128, (n, T))
    x[..., 1] = np.cumsum(rng.random((n, T), dtype=np.float32) * 0.7, axis=1)
    x[..., 2] = rng.random((n, T), dtype=np.float32) * 6.0
    x[..., 3] = rng.integers(0, 200, (n, T))
    for r in range(n):
        k = int(rng.integers(T // 2, T))
        x[r, k:, :] = -1.0                                   # padding
    x[0, 1, 3] = np.nan; x[0, 2, 0] = np.inf; x[1, 0, 1] = -np.inf; x[1, 3, 2] = np.nan; x[2, 5, 3] = -7.0
    x[3, 4, 0] = np.nan                                      # NaN pitch: `!= -1` holds, the row is normalised then zeroed
    x[2, -1, 2] = np.inf                                     # non-finite value inside a padding row
    return x


def main():
    T, n = 64, 6
    x = raw_rolls(5, n, T)
    cfg = {"MAX_NOTES": T, "AUGMENT": {"tempo_jitter": 0.0, "pitch_shift": 0, "note_dropout": 0.0, "velocity_jitter": 0.0,
                                       "timing_jitter": 0.0}}
    out = {"raw": x}
    with tempfile.TemporaryDirectory() as d:
        files = []
        for r in range(n):
            f = os.path.join(d, f"roll{r}.npz")
            np.savez(f, notes=x[r], tempo=120.0, filename=f"roll{r}")
            files.append(f)
        ds = ref_ds.MIDIDataset(files, cfg, augment=False)
        out["default"] = np.stack([ds[r][0] for r in range(n)])
        cfg2 = dict(cfg, MAX_START_BEAT=64.0, MAX_DURATION_BEAT=0.3)
        ds2 = ref_ds.MIDIDataset(files, cfg2, augment=False)
        out["start64_dur0p3"] = np.stack([ds2[r][0] for r in range(n)])
    path = os.path.join(ROOT, "tests", "golden", "ae_norm_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
