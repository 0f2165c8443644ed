"""Recording stand-in for the `pretty_midi` package (TEST INFRASTRUCTURE ONLY).

The reference imports pretty_midi at module top (src/gan/utils.py:11,
tools/roll_to_midi.py:2) and this image does not ship it.  The stub keeps the
constructor surface the reference touches and records every Note(...) so the
oracle can be pinned at the Note-constructor boundary.
"""

LAST_WRITTEN = []


class Note:
    def __init__(self, velocity, pitch, start, end):
        self.velocity, self.pitch, self.start, self.end = velocity, pitch, start, end

    def astuple(self):
        return (self.velocity, self.pitch, self.start, self.end)


class Instrument:
    def __init__(self, program=0, is_drum=False, name=""):
        self.program, self.is_drum, self.name = program, is_drum, name
        self.notes = []


class PrettyMIDI:
    def __init__(self, midi_file=None, resolution=220, initial_tempo=120.0):
        self.resolution, self.initial_tempo = resolution, initial_tempo
        self.instruments = []

    def write(self, path):
        LAST_WRITTEN.clear()
        LAST_WRITTEN.extend(self.instruments)


def instrument_name_to_program(name):
    if name == "Acoustic Grand Piano":
        return 0
    raise ValueError(name)
