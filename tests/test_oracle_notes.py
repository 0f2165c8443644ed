"""Pins oracle/notes_oracle.c against what the REFERENCE produced (tests/golden, made by
oracle/make_golden_notes.py from src/gan/utils.py:95-161 and tools/roll_to_midi.py:10-21)."""
import json
import os

import numpy as np

from melogan import synth
from oracle import notes_oracle


def _digests(golden_dir):
    with open(os.path.join(golden_dir, "notes_digests.json")) as f:
        return json.load(f)


def test_gan_extraction_matches_reference_digests(golden_dir):
    for case in _digests(golden_dir)["gan"]:
        rolls = synth.rolls(case["seed"], case["n"])
        bad, c, p, v, s, e = notes_oracle.extract_notes_gan(rolls, case["bpm"], case["scale"], case["root_key"])
        assert bad == 0
        assert int(c.sum()) == case["notes"], case
        assert notes_oracle.digest(c, p, v, s, e) == case["sha256"], case


def test_gan_extraction_adversarial_full_outputs(golden_dir):
    g = np.load(os.path.join(golden_dir, "notes_golden.npz"))
    adv = synth.adversarial_rolls()
    bad, c, p, v, s, e = notes_oracle.extract_notes_gan(adv, 140.0, "major", 0)
    assert bad == 0
    np.testing.assert_array_equal(c, g["adv_counts"])
    for r, n in enumerate(c):
        np.testing.assert_array_equal(p[r, :n], g["adv_pitch"][r, :n])
        np.testing.assert_array_equal(v[r, :n], g["adv_vel"][r, :n])
        assert s[r, :n].tobytes() == g["adv_start"][r, :n].tobytes(), r   # bit-exact, not allclose
        assert e[r, :n].tobytes() == g["adv_end"][r, :n].tobytes(), r
    bad, c, p, v, s, e = notes_oracle.extract_notes_gan(adv, 70, "minor", 5)
    assert notes_oracle.digest(c, p, v, s, e) == _digests(golden_dir)["gan_adv_minor_root5_bpm70"]


def test_gan_extraction_float64_clock_prefix_is_exercised(golden_dir):
    """The all-floor-steps roll keeps the reference's clock in float64: starts are k*0.1*spb in double."""
    g = np.load(os.path.join(golden_dir, "notes_golden.npz"))
    spb = 60.0 / 140.0
    t, ks = 0.0, []
    for _ in range(8):
        ks.append(t * spb); t += 0.1
    adv = synth.adversarial_rolls()
    _, c, p, v, s, e = notes_oracle.extract_notes_gan(adv[3:4], 140.0, "major", 0)
    ungated = np.nonzero(~(adv[3, :8, 1] < np.float32(-0.2)))[0]
    assert list(s[0, :len(ungated)]) == [ks[i] for i in ungated]


def test_abs_extraction_matches_reference(golden_dir):
    for case in _digests(golden_dir)["abs"]:
        rolls = (synth.rolls(case["seed"], case["n"]) * np.array(case["scale_vec"], np.float32)).astype(np.float32)
        bad, p, v, s, e = notes_oracle.extract_notes_abs(rolls)
        assert bad == 0
        cnt = np.full(case["n"], synth.MAX_NOTES, np.int32)
        assert notes_oracle.digest(cnt, p, v, s, e) == case["sha256"], case
    g = np.load(os.path.join(golden_dir, "notes_golden.npz"))
    bad, p, v, s, e = notes_oracle.extract_notes_abs(g["abs_edge_in"])
    assert bad == 0
    np.testing.assert_array_equal(p, g["abs_edge_pitch"])
    np.testing.assert_array_equal(v, g["abs_edge_vel"])
    assert s.tobytes() == g["abs_edge_start"].tobytes() and e.tobytes() == g["abs_edge_end"].tobytes()


def test_nonfinite_is_reported_like_the_reference_raises():
    r = synth.rolls(5, 2)
    r[1, 10, 0] = np.nan
    r[1, 10, 1] = 0.5
    bad, c, *_ = notes_oracle.extract_notes_gan(r)
    assert bad == 1 and c[0] >= 0 and c[1] == -1
    r[1, 10, 1] = -0.9     # gated row: the reference never touches the NaN pitch
    bad, c, *_ = notes_oracle.extract_notes_gan(r)
    assert bad == 0
