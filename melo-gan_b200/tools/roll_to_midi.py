"""python tools/roll_to_midi.py <roll.npy>  ->  generated_sample.mid in the current directory.

Drop-in for the reference's tools/roll_to_midi.py: rows are (pitch, velocity, duration_seconds, start_seconds),
pitch clipped to [0,127], velocity to [1,127], duration floored at 0.05 s, start at 0 s.  The row mapping runs
in the mg_extract_notes_abs CUDA kernel (bit-exact with the reference loop); the file is written by melogan.midi.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def rows_to_notes(roll):
    import torch
    from melogan import notes
    if not torch.cuda.is_available():
        raise RuntimeError("roll_to_midi runs its row mapping on a CUDA (sm_100a) device; there is no CPU fallback")
    x = torch.from_numpy(np.ascontiguousarray(roll, dtype=np.float32)).reshape(1, -1, 4).cuda()
    b = notes.extract_notes_abs(x)
    return list(zip(b.velocity[0].cpu().tolist(), b.pitch[0].cpu().tolist(), b.start[0].cpu().tolist(),
                    b.end[0].cpu().tolist()))


def main(argv):
    from melogan import midi
    roll = np.load(argv[1])
    midi.write_midi("generated_sample.mid", rows_to_notes(roll), bpm=120.0, program=0)
    print("Wrote generated_sample.mid")


if __name__ == "__main__":
    main(sys.argv)
