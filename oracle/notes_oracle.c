/*
 * notes_oracle.c -- CPU restatement of the reference's note extraction.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under melo-gan_b200/ may call, link or
 * import this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker (never as the
 * thing shipped).
 *
 * Parity: PINNED.  tests/test_oracle_notes.py checks this restatement against
 * tests/golden/notes_*.npz, which oracle/make_golden.py generated in the build
 * container by running the reference's own functions
 *   N-1  save_piano_roll_to_midi   /root/reference/src/gan/utils.py:95-161
 *   N-2  row loop of                /root/reference/tools/roll_to_midi.py:10-21
 * with a recording pretty_midi stand-in (oracle/stubs/pretty_midi.py), i.e. at
 * the pretty_midi.Note(velocity, pitch, start, end) constructor boundary.
 *
 * Arithmetic model (numpy >= 2, NEP 50; measured against numpy 2.3.5):
 *   - a roll row yields np.float32 scalars; python float/int literals are
 *     "weak", so every per-row expression is evaluated in float32 with one
 *     rounding per operation (no FMA contraction: build with -ffp-contract=off);
 *   - python max(c, x) returns the python constant c unless x > c, so floor
 *     values (0.1 beat step, 0.25 beat duration) stay python floats (float64);
 *   - current_time_beats starts as python float 0.0 and stays float64 while
 *     only floor steps were added (t = 0.1 added k times in float64); the first
 *     float32 step turns it into float32(t) + step, float32 from then on;
 *   - start/end are float64 products when every operand is a python float and
 *     float32 products (seconds_per_beat rounded to float32) otherwise.
 * Outputs are returned as float64; a float32 result is widened exactly.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define ORC_OK 0
#define ORC_NONFINITE 1 /* the reference raises ValueError / OverflowError here */

/* snap_to_scale lookup, utils.py:119-128: allowed pitch classes sorted
 * ascending, nearest by |x - note_in_octave| with no octave wrap, first wins. */
void orc_snap_lut(uint32_t allowed_mask, uint8_t lut[12]) {
    for (int nio = 0; nio < 12; ++nio) {
        int best = -1, bestd = 1 << 30;
        for (int x = 0; x < 12; ++x) {
            if (!((allowed_mask >> x) & 1u)) continue;
            int d = x > nio ? x - nio : nio - x;
            if (d < bestd) { bestd = d; best = x; }
        }
        lut[nio] = (uint8_t)(best < 0 ? nio : best);
    }
}

static int trunc_to_i64(float x, long long* out) {
    if (!isfinite(x)) return ORC_NONFINITE; /* int(nan): ValueError, int(inf): OverflowError */
    /* int() is exact for any finite float and np.clip then saturates (numpy 2.3.5:
     * np.clip(int(1e30), 0, 127) == 127), so saturating before the conversion is the same. */
    if (x > 4.0e18f) x = 4.0e18f;
    if (x < -4.0e18f) x = -4.0e18f;
    *out = (long long)x; /* C conversion truncates toward zero like int() */
    return ORC_OK;
}

/* N-1.  roll: nrows x 4 float32, columns (pitch, velocity, duration, step).
 * Returns the note count (<= nrows) or -1 when the reference would raise. */
int orc_extract_notes_gan(const float* roll, int nrows, double bpm, uint32_t allowed_mask,
                          int32_t* pitch, int32_t* velocity, double* start, double* end) {
    /* utils.py:102-103 */
    if (bpm > 180.0) bpm = 180.0;
    if (!(bpm > 60.0)) bpm = 60.0; /* max(60, x): 60 unless x > 60 */
    const double spb64 = 60.0 / bpm;
    const float spb32 = (float)spb64;
    const float thr32 = (float)-0.2;            /* VELOCITY_THRESHOLD as weak scalar */
    const float vrange32 = (float)(1.0 - -0.2); /* utils.py:143 */
    uint8_t lut[12];
    orc_snap_lut(allowed_mask, lut);

    int t_is_f64 = 1;
    double t64 = 0.0;
    float t32 = 0.0f;
    int n = 0;
    for (int i = 0; i < nrows; ++i) {
        const float p = roll[4 * i + 0], v = roll[4 * i + 1], d = roll[4 * i + 2], s = roll[4 * i + 3];
        /* utils.py:133  step_beats = max(0.1, ((s + 1.0) / 2.0) * 4.0) */
        float s32 = ((s + 1.0f) / 2.0f) * 4.0f;
        const int step_is_floor = !(s32 > (float)0.1);

        if (!(v < thr32)) { /* utils.py:135 gate is `v < -0.2`: NaN is NOT gated */
            /* utils.py:139-141 */
            float pf = (p + 1.0f) * 63.5f;
            long long pi;
            if (trunc_to_i64(pf, &pi)) return -1;
            if (pi < 36) pi = 36;
            if (pi > 96) pi = 96;
            int pit = (int)(pi / 12) * 12 + lut[pi % 12];
            /* utils.py:143-146 */
            float voff = v - thr32;
            float vf = 60.0f + (voff / vrange32) * 67.0f;
            long long vi;
            if (trunc_to_i64(vf, &vi)) return -1;
            if (vi < 0) vi = 0;
            if (vi > 127) vi = 127;
            /* utils.py:148 */
            float d32 = ((d + 1.0f) / 2.0f) * 4.0f;
            const int dur_is_floor = !(d32 > (float)0.25);
            double st, en;
            if (t_is_f64) {
                st = t64 * spb64;
                if (dur_is_floor) en = (t64 + 0.25) * spb64;
                else en = (double)(((float)t64 + d32) * spb32);
            } else {
                st = (double)(t32 * spb32);
                float e = dur_is_floor ? t32 + (float)0.25 : t32 + d32;
                en = (double)(e * spb32);
            }
            pitch[n] = pit; velocity[n] = (int)vi; start[n] = st; end[n] = en;
            ++n;
        }
        /* utils.py:136 / :155  current_time_beats += step_beats */
        if (t_is_f64) {
            if (step_is_floor) t64 = t64 + 0.1;
            else { t32 = (float)t64 + s32; t_is_f64 = 0; }
        } else {
            t32 = t32 + (step_is_floor ? (float)0.1 : s32);
        }
    }
    return n;
}

/* N-2.  tools/roll_to_midi.py:10-21 on a float32 roll; one note per row.
 * Returns nrows, or -1 when int(nan) would raise. */
int orc_extract_notes_abs(const float* roll, int nrows, int32_t* pitch, int32_t* velocity,
                          double* start, double* end) {
    for (int i = 0; i < nrows; ++i) {
        const float r0 = roll[4 * i + 0], r1 = roll[4 * i + 1], r2 = roll[4 * i + 2], r3 = roll[4 * i + 3];
        if (isnan(r0)) return -1;
        float pc = r0 < 0.0f ? 0.0f : (r0 > 127.0f ? 127.0f : r0); /* np.clip */
        pitch[i] = (int)pc;
        float m = (r1 < 127.0f) ? r1 : 127.0f; /* min(127, r1) */
        float w = (m > 1.0f) ? m : 1.0f;       /* max(1, .)    */
        velocity[i] = (int)w;
        double dd = (double)r2, ss = (double)r3;
        double dur = (dd > 0.05) ? dd : 0.05;
        double st = (ss > 0.0) ? ss : 0.0;
        start[i] = st;
        end[i] = st + dur;
    }
    return nrows;
}

/* Batched drivers used by the tests and by bench.py's CPU baseline. */
int orc_extract_notes_gan_batch(const float* rolls, long long nrolls, int nrows, double bpm,
                                uint32_t allowed_mask, int32_t* counts, int32_t* pitch,
                                int32_t* velocity, double* start, double* end) {
    int bad = 0;
    for (long long r = 0; r < nrolls; ++r) {
        long long o = r * nrows;
        int c = orc_extract_notes_gan(rolls + o * 4, nrows, bpm, allowed_mask, pitch + o, velocity + o,
                                      start + o, end + o);
        counts[r] = c;
        if (c < 0) bad = 1;
    }
    return bad;
}

int orc_extract_notes_abs_batch(const float* rolls, long long nrolls, int nrows, int32_t* pitch,
                                int32_t* velocity, double* start, double* end) {
    int bad = 0;
    for (long long r = 0; r < nrolls; ++r) {
        long long o = r * nrows;
        if (orc_extract_notes_abs(rolls + o * 4, nrows, pitch + o, velocity + o, start + o, end + o) < 0) bad = 1;
    }
    return bad;
}
