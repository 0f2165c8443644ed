"""melogan.midi against the files the reference commits (generated_tests/*.mid, good_gens1/*.mid: written by
pretty_midi.PrettyMIDI(initial_tempo=bpm).write at src/gan/utils.py:105-158).  Two of them are kept as golden fixtures
under tests/golden/midi/ (outputs of the reference, 3 KB each); where /root/reference is present all sixteen are used.
Parse -> (bpm, program, notes in seconds) -> write_midi must reproduce the file byte for byte."""
import glob
import os

import pytest

from melogan import midi

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "midi", "*.mid")))
REF = sorted(glob.glob("/root/reference/generated_tests/*.mid") + glob.glob("/root/reference/good_gens1/*.mid"))


def _reemit(path, tmp_path):
    m = midi.read_midi(path)
    assert m["resolution"] == 220 and m["tempo"] and m["program"] is not None and len(m["notes"]) > 50
    bpm = round(6e7 / m["tempo"])                               # the reference's emotion presets use integer tempi
    assert midi.tempo_us(bpm) == m["tempo"]
    scale = 60.0 / (bpm * 220)
    notes = [(v, p, on * scale, off * scale) for v, p, on, off in m["notes"]]
    out = str(tmp_path / "re.mid")
    n = midi.write_midi(out, notes, bpm=bpm, program=m["program"])
    want, got = open(path, "rb").read(), open(out, "rb").read()
    assert n == len(got)
    assert got == want, f"{os.path.basename(path)}: first differing byte {next(i for i, (a, b) in enumerate(zip(got, want)) if a != b) if len(got) == len(want) else (len(got), len(want))}"


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_golden_reference_files_reemit_byte_identical(path, tmp_path):
    _reemit(path, tmp_path)


@pytest.mark.skipif(not REF, reason="reference checkout not present (GPU box)")
def test_all_reference_files_reemit_byte_identical(tmp_path):
    for path in REF:
        _reemit(path, tmp_path)


def test_fixtures_present():
    assert len(FILES) >= 2


def test_layout_small_case(tmp_path):
    """Two overlapping notes: same-tick off-before-on ordering, running status, vel-0 offs, EOT at delta 1."""
    out = str(tmp_path / "s.mid")
    scale = 60.0 / (120.0 * 220)
    midi.write_midi(out, [(100, 60, 0.0, 220 * scale), (90, 60, 220 * scale, 440 * scale), (80, 64, 0.0, 110 * scale)],
                    bpm=120.0, program=5)
    d = open(out, "rb").read()
    assert d[:14] == b"MThd\x00\x00\x00\x06\x00\x01\x00\x02\x00\xdc"
    assert d[14:22] == b"MTrk\x00\x00\x00\x13" and d[22:41] == b"\x00\xff\x51\x03\x07\xa1\x20\x00\xff\x58\x04\x04\x02\x18\x08\x01\xff\x2f\x00"
    body = d[49:]
    assert body == (b"\x00\xc0\x05" b"\x00\x90\x3c\x64" b"\x00\x40\x50" b"\x6e\x40\x00" b"\x6e\x3c\x00" b"\x00\x3c\x5a"
                    b"\x81\x5c\x3c\x00" b"\x01\xff\x2f\x00")
    m = midi.read_midi(out)
    assert m["notes"] == [(100, 60, 0, 220), (80, 64, 0, 110), (90, 60, 220, 440)] and m["program"] == 5


def test_out_of_range_note_raises(tmp_path):
    with pytest.raises(ValueError):
        midi.write_midi(str(tmp_path / "x.mid"), [(128, 60, 0.0, 1.0)])
