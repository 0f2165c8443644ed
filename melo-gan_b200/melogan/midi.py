"""Standard MIDI File writer / reader for the note lists produced by melogan.notes.

The reference hands its notes to pretty_midi (PrettyMIDI(initial_tempo=bpm) / Instrument(program) /
Note(velocity, pitch, start, end) / .write(path) at src/gan/utils.py:105-158, tools/roll_to_midi.py:7-24),
which is not a dependency this repository can assume.  write_midi() emits the bytes pretty_midi 0.2.x + mido produce
for that call sequence, so a file written here is byte-identical to the reference's (pinned against the files the
reference commits under generated_tests/ and good_gens1/, see tests/test_midi_cpu.py):

* format-1 header, two tracks, resolution 220 ticks per beat;
* conductor track: set_tempo then a 4/4 time signature at tick 0 (pretty_midi sorts same-tick events by type),
  tempo = int(6e7 / (60 / (tick_scale * resolution))) with tick_scale = 60 / (bpm * resolution) (truncation, as
  PrettyMIDI.write computes it), end-of-track one tick after the last event;
* instrument track on channel 0: program change, then per note a note-on and a note-on with velocity 0 as the
  note-off; tick = int(round(seconds / tick_scale)) (round-half-even, PrettyMIDI.time_to_tick for a single tempo);
  same-tick events ordered by (pitch, velocity), i.e. the off of a pitch before its on; mido's running status (the
  status byte is dropped while it repeats, reset by a meta event); end-of-track at delta 1.
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


def _chunk(body):
    return b"MTrk" + struct.pack(">I", len(body)) + bytes(body)


def tempo_us(bpm, resolution=RESOLUTION):
    """Microseconds per quarter note exactly as PrettyMIDI.write derives them from the tick scale."""
    tick_scale = 60.0 / (float(bpm) * resolution)
    return int(6e7 / (60.0 / (tick_scale * resolution)))


def time_to_tick(seconds, tick_scale):
    """PrettyMIDI.time_to_tick for a file with one tempo: times <= 0 map to tick 0."""
    seconds = float(seconds)
    if not seconds > 0.0:
        return 0
    return int(round(seconds / tick_scale))


def write_midi(path, notes, bpm=120.0, program=0, resolution=RESOLUTION):
    """notes: iterable of (velocity, pitch, start_seconds, end_seconds).  Returns the number of bytes written."""
    tick_scale = 60.0 / (float(bpm) * resolution)           # seconds per tick
    conductor = (b"\x00\xff\x51\x03" + struct.pack(">I", tempo_us(bpm, resolution))[1:]
                 + b"\x00\xff\x58\x04\x04\x02\x18\x08" + b"\x01\xff\x2f\x00")
    ev = []                                                   # (tick, note * 256 + velocity, data bytes)
    for velocity, pitch, start, end in notes:
        p, v = int(pitch), int(velocity)
        if not (0 <= p <= 127 and 0 <= v <= 127):
            raise ValueError(f"note out of the MIDI data-byte range: pitch {p}, velocity {v}")   # mido raises too
        ev.append((time_to_tick(start, tick_scale), p * 256 + v, bytes([p, v])))
        ev.append((time_to_tick(end, tick_scale), p * 256, bytes([p, 0])))
    ev.sort(key=lambda e: (e[0], e[1]))                       # stable, like sorted(cmp_to_key(event_compare))
    body, last, running = bytearray(b"\x00" + bytes([0xC0, int(program) & 0x7F])), 0, 0xC0
    for tick, _, data in ev:
        body += _vlq(tick - last)
        if running != 0x90:
            body.append(0x90)
            running = 0x90
        body += data
        last = tick
    body += b"\x01\xff\x2f\x00"
    data = b"MThd" + struct.pack(">IHHH", 6, 1, 2, resolution) + _chunk(conductor) + _chunk(body)
    with open(path, "wb") as f:
        f.write(data)
    return len(data)


def read_midi(path):
    """Parses a format-0/1 file of note / program / meta events (running status, explicit note-offs and velocity-0
    note-ons) -> dict(resolution, tempo, program, notes) with notes = [(velocity, pitch, on_tick, off_tick)] sorted by
    (on_tick, pitch).  Reads the reference's own generated files as well as write_midi's."""
    d = open(path, "rb").read()
    if d[:4] != b"MThd":
        raise ValueError("not a Standard MIDI File")
    ntrk, res = struct.unpack(">HH", d[10:14])
    pos, notes, tempo, program = 14, [], None, None
    for _ in range(ntrk):
        if d[pos:pos + 4] != b"MTrk":
            raise ValueError("missing MTrk chunk")
        ln = struct.unpack(">I", d[pos + 4:pos + 8])[0]
        p, end, tick, open_notes, status = pos + 8, pos + 8 + ln, 0, {}, None
        while p < end:
            delta = 0
            while True:
                b = d[p]
                p += 1
                delta = (delta << 7) | (b & 0x7F)
                if not b & 0x80:
                    break
            tick += delta
            if d[p] == 0xFF:
                typ, p = d[p + 1], p + 2
                ln2 = 0
                while True:
                    b = d[p]
                    p += 1
                    ln2 = (ln2 << 7) | (b & 0x7F)
                    if not b & 0x80:
                        break
                if typ == 0x51 and tempo is None:
                    tempo = int.from_bytes(d[p:p + 3], "big")
                p += ln2
                status = None
                continue
            if d[p] & 0x80:
                status, p = d[p], p + 1
            elif status is None:
                raise ValueError("data byte without a running status")
            kind = status & 0xF0
            if kind in (0xC0, 0xD0):
                if kind == 0xC0 and program is None:
                    program = d[p]
                p += 1
            elif kind in (0x80, 0x90):
                pitch, vel = d[p], d[p + 1]
                p += 2
                if kind == 0x90 and vel > 0:
                    open_notes.setdefault(pitch, []).append((vel, tick))
                elif open_notes.get(pitch):
                    v, on = open_notes[pitch].pop(0)
                    notes.append((v, pitch, on, tick))
            elif kind in (0xA0, 0xB0, 0xE0):
                p += 2
            else:
                raise ValueError(f"unexpected status byte 0x{status:02x}")
        pos = end
    return {"resolution": res, "tempo": tempo, "program": program,
            "notes": sorted(notes, key=lambda n: (n[2], n[1]))}


def read_notes(path):
    """(resolution, tempo, notes) of read_midi (kept for the round-trip tests)."""
    m = read_midi(path)
    return m["resolution"], m["tempo"], m["notes"]
