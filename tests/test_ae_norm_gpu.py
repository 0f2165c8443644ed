"""mg_ae_normalize (8f-3, reference src/ae/dataset.py:72-89,105) through the C ABI: bit-exact against the golden output
of the reference's MIDIDataset and against the oracle on a large seeded set, incl. in-place use."""
import os

import numpy as np
import pytest
import torch

from melogan import notes as N
from oracle import notes_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ae_norm_golden.npz")


def test_golden_of_the_reference_dataset():
    g = np.load(GOLD)
    x = torch.from_numpy(g["raw"]).cuda()
    assert N.ae_normalize(x).cpu().numpy().tobytes() == g["default"].tobytes()
    assert N.ae_normalize(x, 64.0, 0.3).cpu().numpy().tobytes() == g["start64_dur0p3"].tobytes()


def test_large_seeded_set_matches_oracle_and_in_place():
    rng = np.random.default_rng(11)
    R, T = 4096, 512
    x = np.empty((R, T, 4), np.float32)
    x[..., 0] = rng.integers(-1, 130, (R, T))
    x[..., 1] = rng.random((R, T), dtype=np.float32) * 300.0
    x[..., 2] = rng.standard_normal((R, T), dtype=np.float32) * 8.0
    x[..., 3] = rng.standard_normal((R, T), dtype=np.float32) * 90.0 + 60.0
    bad = rng.random((R, T)) < 1e-3
    x[bad, rng.integers(0, 4, bad.sum())] = rng.choice([np.nan, np.inf, -np.inf], bad.sum())
    want = notes_oracle.ae_normalize(x)
    d = torch.from_numpy(x).cuda()
    got = N.ae_normalize(d)
    assert got.cpu().numpy().tobytes() == want.tobytes()
    N.ae_normalize(d, out=d)                                  # aliasing is allowed
    assert d.cpu().numpy().tobytes() == want.tobytes()


def test_empty_and_errors():
    assert N.ae_normalize(torch.empty((0, 512, 4), device="cuda")).shape == (0, 512, 4)
    with pytest.raises(ValueError):
        N.ae_normalize(torch.zeros(2, 8, 4))                  # CPU tensor: no CPU path
