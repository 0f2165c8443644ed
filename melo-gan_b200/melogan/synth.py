"""Deterministic synthetic inputs shared by tests, bench.py and oracle/make_golden.py.

Everything is derived from numpy's PCG64 `random()` doubles with exact IEEE
arithmetic, so the same seed gives the same bits on every machine (the GPU box
has no /root/reference and must regenerate the inputs the goldens were made on).
Shapes follow config/gan_config.yaml: a "roll" is a (MAX_NOTES=512, NOTE_DIM=4)
float32 note-event tensor with columns (pitch, velocity, duration, step).
"""
import numpy as np

MAX_NOTES = 512
NOTE_DIM = 4

# app.py:109-110 of the reference: per-emotion tempo and scale of the demo
EMOTION_RENDER = {
    "happy": (140.0, "major"),
    "sad": (70.0, "minor"),
    "angry": (160.0, "minor"),
    "calm": (90.0, "major"),
}


def uniform(seed, shape, lo=-1.0, hi=1.0):
    rng = np.random.Generator(np.random.PCG64(seed))
    u = rng.random(int(np.prod(shape)), dtype=np.float64)
    return (lo + (hi - lo) * u).astype(np.float32).reshape(shape)


def pseudo_normal(seed, shape, std=1.0):
    """Zero-mean, unit-variance-ish values from a sum of 4 uniforms (exact arithmetic only)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = int(np.prod(shape))
    u = rng.random(4 * n, dtype=np.float64).reshape(4, n)
    z = (u.sum(axis=0) - 2.0) * np.sqrt(3.0)  # var of sum of 4 U(0,1) = 1/3
    return (z * std).astype(np.float32).reshape(shape)


def rolls(seed, n, spread=1.15):
    """n synthetic rolls ~U(-spread, spread): slightly wider than the normalised range so the
    clip/gate/floor branches of the extractors are all exercised."""
    return uniform(seed, (n, MAX_NOTES, NOTE_DIM), -spread, spread)


def adversarial_rolls():
    """Rolls that sit on every threshold of save_piano_roll_to_midi (src/gan/utils.py:130-155)."""
    f32 = np.float32
    out = []
    base = uniform(991, (MAX_NOTES, NOTE_DIM))
    thr = f32(-0.2)

    r = base.copy(); r[:, 1] = -1.0; out.append(r)                      # every row gated
    r = base.copy(); r[:, 1] = 0.7; out.append(r)                       # no row gated
    r = base.copy(); r[0::3, 1] = thr                                   # exactly at the gate
    r[1::3, 1] = np.nextafter(thr, f32(-1)); r[2::3, 1] = np.nextafter(thr, f32(1)); out.append(r)
    r = base.copy(); r[:, 3] = -1.0; out.append(r)                      # only floor steps: float64 clock
    for k in (1, 2, 7, 100, 511):                                       # k leading floor steps
        r = base.copy(); r[:k, 3] = -1.0; r[k:, 3] = np.abs(r[k:, 3]); out.append(r)
    r = base.copy()                                                     # step values around the 0.1 floor
    s0 = f32(-0.95)
    cand = [s0]
    for _ in range(6):
        cand.append(np.nextafter(cand[-1], f32(-1)))
    lo = cand[-1]
    for _ in range(12):
        lo = np.nextafter(lo, f32(-1)); cand.append(lo)
    r[:, 3] = np.resize(np.array(cand, dtype=f32), MAX_NOTES); out.append(r)
    r = base.copy()                                                     # duration values around 0.25
    d0 = f32(-0.875)
    cand = [d0, np.nextafter(d0, f32(-1)), np.nextafter(d0, f32(1)), f32(-1.0), f32(-0.8749)]
    r[:, 2] = np.resize(np.array(cand, dtype=f32), MAX_NOTES); r[:, 1] = 0.3; out.append(r)
    r = base.copy()                                                     # pitch edges and |x| > 1
    pe = [(35.0 / 63.5) - 1, (36.0 / 63.5) - 1, (36.99 / 63.5) - 1, (96.0 / 63.5) - 1, (97.0 / 63.5) - 1,
          (96.999 / 63.5) - 1, 2.0, -3.0, 1e30, -1e30, 0.0, -1.0, 1.0]
    r[:, 0] = np.resize(np.array(pe, dtype=f32), MAX_NOTES); r[:, 1] = 0.1; out.append(r)
    r = base.copy()                                                     # velocity range incl. > 1 and huge
    ve = [-0.2, -0.19999, 1.0, 1.0001, 5.0, 1e30, 0.0, 0.4, 0.99999]
    r[:, 1] = np.resize(np.array(ve, dtype=f32), MAX_NOTES); out.append(r)
    r = np.zeros((MAX_NOTES, NOTE_DIM), f32); out.append(r)             # all zeros
    r = np.full((MAX_NOTES, NOTE_DIM), -1.0, f32); out.append(r)        # padded tail like real data
    return np.stack(out).astype(f32)


def numeric_features(seed, n):
    """(n, 6) standardised numeric features; column 5 is constant 0 (scaler.joblib var_=0)."""
    x = pseudo_normal(seed, (n, 6))
    x[:, 5] = 0.0
    return x


def emotion_labels(n):
    return (np.arange(n) % 4).astype(np.int64)  # happy, sad, angry, calm = 0..3
