"""Generates tests/golden/notes_golden.npz + notes_digests.json by running the REFERENCE's own
note extraction (src/gan/utils.py:95-161, tools/roll_to_midi.py:10-21) in the build container.

    python oracle/make_golden_notes.py          (needs /root/reference; not available on the GPU box)

Inputs come from melogan.synth (seed-reproducible on any machine); what is stored is what the
reference passed to pretty_midi.Note(velocity, pitch, start, end).
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "stubs"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "melo-gan_b200"))
sys.path.insert(0, ROOT)

import pretty_midi  # the recording stub  # noqa: E402
from melogan import synth  # noqa: E402
from oracle import notes_oracle  # noqa: E402

import importlib.util  # noqa: E402
_spec = importlib.util.spec_from_file_location("ref_gan_utils", "/root/reference/src/gan/utils.py")
ref_utils = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(ref_utils)


def ref_gan(rolls, bpm, scale, root_key):
    R, T, _ = rolls.shape
    counts = np.zeros(R, np.int32)
    pitch = np.zeros((R, T), np.int32); vel = np.zeros((R, T), np.int32)
    start = np.zeros((R, T), np.float64); end = np.zeros((R, T), np.float64)
    for r in range(R):
        with contextlib.redirect_stdout(io.StringIO()):
            ref_utils.save_piano_roll_to_midi(rolls[r], "unused.mid", bpm=bpm, scale=scale, root_key=root_key)
        notes = pretty_midi.LAST_WRITTEN[0].notes
        counts[r] = len(notes)
        for i, n in enumerate(notes):
            pitch[r, i], vel[r, i], start[r, i], end[r, i] = int(n.pitch), int(n.velocity), float(n.start), float(n.end)
    return counts, pitch, vel, start, end


def ref_abs(rolls):
    """tools/roll_to_midi.py is a module-level script: its row loop (lines 10-21) is executed here
    verbatim by compiling exactly those source lines against a roll we supply."""
    src = open("/root/reference/tools/roll_to_midi.py").read().splitlines()
    body = "\n".join(src[9:21])  # lines 10-21: `for row in roll:` ... `))`
    code = compile(body, "roll_to_midi.py[10:21]", "exec")
    R, T, _ = rolls.shape
    pitch = np.zeros((R, T), np.int32); vel = np.zeros((R, T), np.int32)
    start = np.zeros((R, T), np.float64); end = np.zeros((R, T), np.float64)
    for r in range(R):
        inst = pretty_midi.Instrument(program=0)
        env = {"np": np, "pretty_midi": pretty_midi, "roll": rolls[r], "instrument": inst}
        exec(code, env)
        assert len(inst.notes) == T
        for i, n in enumerate(inst.notes):
            pitch[r, i], vel[r, i], start[r, i], end[r, i] = n.pitch, n.velocity, n.start, n.end
    return pitch, vel, start, end


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    digests = {"gan": [], "abs": []}
    npz = {}

    # (1) seeded bulk cases -> digests only.  One case per emotion render setting (app.py:109-110).
    for k, (emo, (bpm, scale)) in enumerate(synth.EMOTION_RENDER.items()):
        seed, n = 100 + k, 192
        rolls = synth.rolls(seed, n)
        c, p, v, s, e = ref_gan(rolls, bpm, scale, 0)
        digests["gan"].append({"seed": seed, "n": n, "bpm": bpm, "scale": scale, "root_key": 0, "emotion": emo,
                               "notes": int(c.sum()), "sha256": notes_oracle.digest(c, p, v, s, e)})
    # scale / root / bpm-clamp sweep on a smaller set
    sweep = [(30, "blues", 3), (500, "dorian", 11), (97.3, "locrian", 7), (120, "no_such_scale", 5),
             (60, "minor_pentatonic", 1), (180, "major_pentatonic", 9), (120.0, "phrygian", 4),
             (133, "lydian", 6), (101, "mixolydian", 2), (75, "chromatic", 0)]
    for k, (bpm, scale, root) in enumerate(sweep):
        seed, n = 200 + k, 24
        rolls = synth.rolls(seed, n)
        c, p, v, s, e = ref_gan(rolls, bpm, scale, root)
        digests["gan"].append({"seed": seed, "n": n, "bpm": bpm, "scale": scale, "root_key": root,
                               "notes": int(c.sum()), "sha256": notes_oracle.digest(c, p, v, s, e)})

    # (2) adversarial rolls -> full expected outputs
    adv = synth.adversarial_rolls()
    c, p, v, s, e = ref_gan(adv, 140.0, "major", 0)
    npz.update(adv_counts=c, adv_pitch=p.astype(np.uint8), adv_vel=v.astype(np.uint8), adv_start=s, adv_end=e)
    c2, p2, v2, s2, e2 = ref_gan(adv, 70, "minor", 5)
    digests["gan_adv_minor_root5_bpm70"] = notes_oracle.digest(c2, p2, v2, s2, e2)

    # (3) tools/roll_to_midi.py on GAN-range rolls and on MIDI-range rolls (pitch 0..127, seconds)
    for k, (seed, n, scale_vec) in enumerate([(300, 64, (1, 1, 1, 1)), (301, 64, (90.0, 140.0, 3.0, 40.0))]):
        rolls = (synth.rolls(seed, n) * np.array(scale_vec, np.float32)).astype(np.float32)
        p, v, s, e = ref_abs(rolls)
        cnt = np.full(n, synth.MAX_NOTES, np.int32)
        digests["abs"].append({"seed": seed, "n": n, "scale_vec": list(scale_vec),
                               "sha256": notes_oracle.digest(cnt, p, v, s, e)})
    edge = np.zeros((1, synth.MAX_NOTES, 4), np.float32)
    vals = np.array([-5, -0.0, 0.0, 0.5, 0.999, 1.0, 1.5, 126.9, 127.0, 127.5, 500, 0.05, 0.049999, np.nan, np.inf, -np.inf],
                    np.float32)
    edge[0, :, 0] = np.resize(np.nan_to_num(vals, nan=3.0, posinf=1e9, neginf=-1e9), synth.MAX_NOTES)
    edge[0, :, 1] = np.resize(vals, synth.MAX_NOTES)       # velocity column may hold NaN/inf (min/max swallow it)
    edge[0, :, 2] = np.resize(np.roll(vals, 3), synth.MAX_NOTES)
    edge[0, :, 3] = np.resize(np.roll(vals, 7), synth.MAX_NOTES)
    edge[0, :, 1] = np.nan_to_num(edge[0, :, 1], posinf=1e9, neginf=-1e9)  # int(inf) would raise; NaN stays
    p, v, s, e = ref_abs(edge)
    npz.update(abs_edge_in=edge, abs_edge_pitch=p.astype(np.int32), abs_edge_vel=v.astype(np.int32),
               abs_edge_start=s, abs_edge_end=e)

    np.savez_compressed(os.path.join(out_dir, "notes_golden.npz"), **npz)
    with open(os.path.join(out_dir, "notes_digests.json"), "w") as f:
        json.dump(digests, f, indent=1)
    print("wrote", out_dir)


if __name__ == "__main__":
    main()
