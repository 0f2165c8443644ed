"""GPU parity of the note-extraction kernels (through the C ABI) against the oracle and the goldens."""
import json
import os

import numpy as np
import pytest
import torch

from melogan import synth

pytestmark = pytest.mark.gpu


def _digest_of(batch):
    from oracle import notes_oracle
    c = batch.counts.cpu().numpy() if hasattr(batch.counts, "cpu") else batch.counts
    f = lambda t: (t.cpu().numpy() if hasattr(t, "cpu") else t)
    return notes_oracle.digest(c, f(batch.pitch).astype(np.int32), f(batch.velocity).astype(np.int32),
                               f(batch.start), f(batch.end))


def test_gan_extraction_matches_reference_digests(golden_dir):
    from melogan import notes
    with open(os.path.join(golden_dir, "notes_digests.json")) as fh:
        dg = json.load(fh)
    for case in dg["gan"]:
        rolls = torch.from_numpy(synth.rolls(case["seed"], case["n"])).cuda()
        out = notes.extract_notes_gan(rolls, case["bpm"], case["scale"], case["root_key"])
        assert int(out.counts.sum()) == case["notes"]
        assert _digest_of(out) == case["sha256"], case
    adv = torch.from_numpy(synth.adversarial_rolls()).cuda()
    assert _digest_of(notes.extract_notes_gan(adv, 70, "minor", 5)) == dg["gan_adv_minor_root5_bpm70"]


def test_gan_extraction_adversarial_bitexact_vs_golden_and_oracle(golden_dir):
    from melogan import notes
    from oracle import notes_oracle
    g = np.load(os.path.join(golden_dir, "notes_golden.npz"))
    adv_np = synth.adversarial_rolls()
    out = notes.extract_notes_gan(torch.from_numpy(adv_np).cuda(), 140.0, "major", 0)
    c = out.counts.cpu().numpy()
    np.testing.assert_array_equal(c, g["adv_counts"])
    p, v, s, e = (t.cpu().numpy() for t in (out.pitch, out.velocity, out.start, out.end))
    for r, n in enumerate(c):
        np.testing.assert_array_equal(p[r, :n], g["adv_pitch"][r, :n])
        np.testing.assert_array_equal(v[r, :n], g["adv_vel"][r, :n])
        assert s[r, :n].tobytes() == g["adv_start"][r, :n].tobytes(), r
        assert e[r, :n].tobytes() == g["adv_end"][r, :n].tobytes(), r
    _, oc, op, ov, os_, oe = notes_oracle.extract_notes_gan(adv_np, 140.0, "major", 0)
    assert notes_oracle.digest(oc, op, ov, os_, oe) == _digest_of(out)


@pytest.mark.parametrize("R,T", [(0, 512), (1, 512), (17, 512), (33, 100), (5, 1), (3, 0), (40, 511)])
def test_gan_extraction_ragged_shapes_vs_oracle(R, T):
    from melogan import notes
    from oracle import notes_oracle
    rolls = synth.uniform(77 + R + T, (R, max(T, 1), 4), -1.15, 1.15)[:, :T, :].copy()
    out = notes.extract_notes_gan(torch.from_numpy(rolls).cuda().reshape(R, T, 4), 97.3, "dorian", 4)
    if R == 0 or T == 0:
        assert out.counts.numel() == R and (R == 0 or int(out.counts.abs().sum()) == 0)
        return
    _, oc, op, ov, os_, oe = notes_oracle.extract_notes_gan(rolls, 97.3, "dorian", 4)
    assert notes_oracle.digest(oc, op, ov, os_, oe) == _digest_of(out)


def test_gan_extraction_nonfinite_raises_like_reference():
    from melogan import notes
    r = synth.rolls(5, 3)
    r[1, 10, 0] = np.nan
    r[1, 10, 1] = 0.5
    with pytest.raises(ValueError):
        notes.extract_notes_gan(torch.from_numpy(r).cuda())
    out = notes.extract_notes_gan(torch.from_numpy(r).cuda(), check=False)
    assert out.counts.cpu().tolist()[1] == -1 and out.counts.cpu().tolist()[0] >= 0
    r[1, 10, 1] = -0.9  # gated: NaN pitch is never read
    notes.extract_notes_gan(torch.from_numpy(r).cuda())
    with pytest.raises(ValueError):
        notes.extract_notes_gan(torch.from_numpy(r))  # CPU tensor: no fallback


def test_abs_extraction_matches_reference(golden_dir):
    from melogan import notes
    with open(os.path.join(golden_dir, "notes_digests.json")) as fh:
        dg = json.load(fh)
    for case in dg["abs"]:
        rolls = (synth.rolls(case["seed"], case["n"]) * np.array(case["scale_vec"], np.float32)).astype(np.float32)
        out = notes.extract_notes_abs(torch.from_numpy(rolls).cuda())
        assert _digest_of(out) == case["sha256"], case
    g = np.load(os.path.join(golden_dir, "notes_golden.npz"))
    out = notes.extract_notes_abs(torch.from_numpy(g["abs_edge_in"]).cuda())
    np.testing.assert_array_equal(out.pitch.cpu().numpy(), g["abs_edge_pitch"])
    np.testing.assert_array_equal(out.velocity.cpu().numpy(), g["abs_edge_vel"])
    assert out.start.cpu().numpy().tobytes() == g["abs_edge_start"].tobytes()
    assert out.end.cpu().numpy().tobytes() == g["abs_edge_end"].tobytes()
    bad = g["abs_edge_in"].copy(); bad[0, 3, 0] = np.nan
    with pytest.raises(ValueError):
        notes.extract_notes_abs(torch.from_numpy(bad).cuda())


def test_host_buffer_entry_points_match_device_path():
    from melogan import notes
    rolls = synth.rolls(4242, 300)
    dev = notes.extract_notes_gan(torch.from_numpy(rolls).cuda(), 160.0, "minor", 0)
    host = notes.extract_notes_gan_host(rolls, 160.0, "minor", 0)
    assert _digest_of(dev) == _digest_of(host)
    d2 = notes.extract_notes_abs(torch.from_numpy(rolls).cuda())
    h2 = notes.extract_notes_abs_host(rolls)
    assert _digest_of(d2) == _digest_of(h2)


def test_full_size_properties_one_million_bars_class():
    """BASELINE config #5 scale (chunk of the 1M-bar run): size-independent properties + sampled oracle check."""
    from melogan import notes
    from oracle import notes_oracle
    R = 131072
    g = torch.Generator(device="cuda").manual_seed(9)
    rolls = torch.rand((R, 512, 4), generator=g, device="cuda") * 2.3 - 1.15
    out = notes.extract_notes_gan(rolls, 140.0, "major", 0)
    # counts == number of ungated rows (the gate is `v < float32(-0.2)`)
    want = (~(rolls[:, :, 1] < np.float32(-0.2))).sum(dim=1).to(torch.int32)
    assert torch.equal(out.counts, want)
    # onsets never decrease inside a roll, offsets are after onsets, quantised ranges hold
    idx = torch.arange(512, device="cuda")[None, :]
    valid = idx < out.counts[:, None]
    d = out.start[:, 1:] - out.start[:, :-1]
    assert bool(((d >= 0) | ~valid[:, 1:]).all())
    assert bool(((out.end > out.start) | ~valid).all())
    assert bool((((out.pitch >= 36) & (out.pitch <= 96)) | ~valid).all())
    assert bool(((out.velocity >= 60) | ~valid).all())
    allowed = torch.tensor([1 if k in (0, 2, 4, 5, 7, 9, 11) else 0 for k in range(12)], device="cuda", dtype=torch.bool)
    assert bool((allowed[(out.pitch % 12).long()] | ~valid).all())
    sample = torch.randint(0, R, (256,), generator=torch.Generator().manual_seed(1))
    sub = rolls[sample.cuda()].cpu().numpy()
    _, oc, op, ov, os_, oe = notes_oracle.extract_notes_gan(sub, 140.0, "major", 0)
    from melogan.notes import NoteBatch
    pick = NoteBatch(*(t[sample.cuda()] for t in (out.counts, out.pitch, out.velocity, out.start, out.end)))
    assert notes_oracle.digest(oc, op, ov, os_, oe) == _digest_of(pick)
