"""Minimal Standard MIDI File writer for the note lists produced by melogan.notes.

The reference hands its notes to pretty_midi (PrettyMIDI(initial_tempo=bpm) / Instrument(program) /
Note(velocity, pitch, start, end) / .write(path) at src/gan/utils.py:105-158, tools/roll_to_midi.py:7-24),
which is not a dependency this repository can assume.  This writer follows pretty_midi's file layout:
format-1 SMF, resolution 220 ticks per beat, a conductor track with the tempo and a 4/4 time signature,
one instrument track with a program change followed by the note events, seconds converted to ticks at the
initial tempo and rounded to the nearest tick, note-offs sorted before note-ons at equal ticks.
(bit-exactness is defined at the Note(...) boundary, not at the byte level.)
"""
import struct

RESOLUTION = 220


def _vlq(n):
    n = int(n)
    out = [n & 0x7F]
    n >>= 7
    while n:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    return bytes(reversed(out))


def _track(events):
    """events: list of (tick, sort_key, bytes); returns an MTrk chunk with delta times."""
    events = sorted(events, key=lambda e: (e[0], e[1]))
    body, last = bytearray(), 0
    for tick, _, data in events:
        body += _vlq(tick - last) + data
        last = tick
    body += b"\x00\xff\x2f\x00"
    return b"MTrk" + struct.pack(">I", len(body)) + bytes(body)


def write_midi(path, notes, bpm=120.0, program=0, resolution=RESOLUTION):
    """notes: iterable of (velocity, pitch, start_seconds, end_seconds)."""
    tick_scale = 60.0 / (float(bpm) * resolution)           # seconds per tick
    tempo = int(round(60_000_000.0 / float(bpm)))           # microseconds per quarter note
    conductor = [(0, 0, b"\xff\x51\x03" + struct.pack(">I", tempo)[1:]),
                 (0, 1, b"\xff\x58\x04\x04\x02\x18\x08")]
    ev = [(0, 0, bytes([0xC0, int(program) & 0x7F]))]
    for velocity, pitch, start, end in notes:
        on, off = int(round(float(start) / tick_scale)), int(round(float(end) / tick_scale))
        p, v = max(0, min(127, int(pitch))), max(0, min(127, int(velocity)))
        ev.append((on, 2, bytes([0x90, p, v])))
        ev.append((max(off, on), 1, bytes([0x80, p, 0])))
    data = b"MThd" + struct.pack(">IHHH", 6, 1, 2, resolution) + _track(conductor) + _track(ev)
    with open(path, "wb") as f:
        f.write(data)
    return len(data)


def read_notes(path):
    """Parses a file written by write_midi back into (velocity, pitch, on_tick, off_tick) tuples (test helper)."""
    d = open(path, "rb").read()
    assert d[:4] == b"MThd"
    ntrk, res = struct.unpack(">HH", d[10:14])
    pos, notes, tempo = 14, [], None
    for _ in range(ntrk):
        assert d[pos:pos + 4] == b"MTrk"
        ln = struct.unpack(">I", d[pos + 4:pos + 8])[0]
        p, end, tick, open_notes = pos + 8, pos + 8 + ln, 0, {}
        while p < end:
            delta = 0
            while True:
                b = d[p]
                p += 1
                delta = (delta << 7) | (b & 0x7F)
                if not b & 0x80:
                    break
            tick += delta
            st = d[p]
            if st == 0xFF:
                typ, ln2 = d[p + 1], d[p + 2]
                if typ == 0x51:
                    tempo = int.from_bytes(d[p + 3:p + 6], "big")
                p += 3 + ln2
            elif st & 0xF0 == 0xC0:
                p += 2
            elif st & 0xF0 == 0x90:
                open_notes.setdefault(d[p + 1], []).append((d[p + 2], tick))
                p += 3
            elif st & 0xF0 == 0x80:
                v, on = open_notes[d[p + 1]].pop(0)
                notes.append((v, d[p + 1], on, tick))
                p += 3
            else:
                raise ValueError("unexpected status byte")
        pos = end
    return res, tempo, sorted(notes, key=lambda n: (n[2], n[1]))
