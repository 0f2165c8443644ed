// notes.cu -- batched note extraction (velocity gate + onset clock + quantisation) on sm_100a.
//
// N-1  save_piano_roll_to_midi row loop   reference src/gan/utils.py:130-155
// N-2  tools/roll_to_midi.py row loop     reference tools/roll_to_midi.py:10-21
//
// HBM-bound byte/float work: per roll 8 KiB in, <= 512*(1+1+8+8) B + 4 B out.  The only
// sequential part is the onset clock: the reference accumulates `current_time_beats` one row
// at a time in float32 (float64 while only floor steps were added), and float addition is not
// associative, so a tree/warp scan is not bit-exact.  Layout of the work (details at the kernel):
//   P1     row -> note code, floored duration, floored step, in shared memory (bulk-copied tiles);
//   clock  one LANE PER ROLL walks that roll's 512 steps in order (serial over rows, parallel over rolls);
//   P3     __ballot_sync + popc compaction gives every surviving row its output slot (integer prefix, exact),
//          coalesced stores of the note fields.
// All float32 arithmetic uses __f*_rn intrinsics so nvcc cannot contract mul+add into FMA
// (numpy rounds after every operation).
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kMaxRows = 512;

// t = 0.0; t += 0.1 (float64) k times: the reference's clock while every step so far was the floor.
__constant__ double c_floor_clock[kMaxRows + 1];

struct GanParams {
    const float4* rolls;
    long long nrolls;
    int nrows;
    double spb64;
    float spb32;
    uint8_t lut[12];
    int32_t* counts;
    uint8_t* pitch;
    uint8_t* velocity;
    double* start;
    double* end;
};

__device__ __forceinline__ float unit_to_beats(float x) {
    // ((x + 1.0) / 2.0) * MAX_BEAT_TIME   (utils.py:133,148), one rounding per op
    // x / 2.0 == x * 0.5 bit for bit (both are the correctly rounded value of the same real number, subnormals included);
    // written as a multiplication because nvcc expands __fdiv_rn into its full division sequence even for 2.0f
    return __fmul_rn(__fmul_rn(__fadd_rn(x, 1.0f), 0.5f), 4.0f);
}

// ---------------------------------------------------------------------------------------------
// N-1.  One persistent CTA per SM runs the three phases of three DIFFERENT tiles at the same time (round 1 ran them one
// after the other per tile: one lane per roll busy for the whole clock phase while every other thread waited at a
// barrier, and the rolls were read twice -- 0.41 of the HBM roof; this form: 0.65):
//
//   stage j:   bulk copy (cp.async.bulk, no LSU work) of tile j+1: 8 rolls = 64 KB of HBM -> shared memory
//              worker warps (2 per roll)    P3(tile j-2)  then  P1(tile j)
//              clock warp (1 lane per roll) the serial float32 onset clock of tile j-1
//
//   P1  row -> (pitch | velocity << 8) code (2 B), floored duration (4 B), floored step (4 B): 10 B per row stay in
//       shared memory, the 16-byte row is dead after P1 (its buffer is refilled by the next bulk copy);
//   clock  per roll: 512 dependent __fadd_rn, register-blocked four steps per LDS.128 / STS.128, so that only the
//       FADD chain is serial;  P3  ballot/popc compaction and the float64 widening of onset / offset, coalesced stores.
// Same arithmetic, same operation order as the reference (src/gan/utils.py:130-155): bit-exact.
// ---------------------------------------------------------------------------------------------
constexpr int kTile = 8;                         // rolls per tile
constexpr int kWorkerWarps = 2 * kTile;          // two warps per roll (first / second half of its rows)
constexpr int kThreadsV2 = (kWorkerWarps + 1) * 32;
constexpr int kClkStride = kMaxRows + 36;        // floats; 548 % 32 = 4: the clock lanes' LDS.128 hit disjoint banks, and the
                                                 // software-pipelined clock loop may read 12 floats past row 511

struct alignas(16) NotesSmem {
    float4 raw[2][kTile * kMaxRows];             // 2 x 64 KB, written by bulk copies
    float clk[2][kTile][kClkStride];             // P1: max(0.1, step); clock warp: exclusive onset clock
    float dur[2][kTile][kMaxRows];               // max(0.25, duration)
    unsigned short code[2][kTile][kMaxRows];     // pitch | velocity << 8; 0xFFFF = gated; 0 = non-finite
    double floor_clock[kMaxRows + 1];
    unsigned long long full[2];
    int first_nf[4][kTile][2];                   // per half: first row whose step is above the floor (or nrows)
    int nkeep[4][kTile][2];
    int bad[4][kTile][2];
    int ndbl[4][kTile];
    unsigned char snap[64];                      // pitch 36..96 -> scale-snapped pitch
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void nb_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t a = smem_addr(bar);
    uint32_t ok = 0;
    for (uint32_t spins = 0; !ok; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (!ok && spins > (1u << 26)) asm volatile("trap;");
    }
}

__global__ void __launch_bounds__(kThreadsV2, 1) extract_notes_gan_pipe_kernel(GanParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NotesSmem& S = *reinterpret_cast<NotesSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = P.nrows;
    const float thr32 = (float)-0.2, vrange32 = (float)(1.0 - -0.2);
    const float floor_step32 = (float)0.1, floor_dur32 = (float)0.25;
    const long long ntiles = (P.nrolls + kTile - 1) / kTile;
    const int ntl = blockIdx.x < ntiles ? (int)((ntiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;   // tiles of this CTA

    for (int i = tid; i <= kMaxRows; i += kThreadsV2) S.floor_clock[i] = c_floor_clock[i];
    if (tid < 61) { const int pc = 36 + tid; S.snap[tid] = (unsigned char)((pc / 12) * 12 + P.lut[pc % 12]); }
    if (tid == 0) {
        for (int b = 0; b < 2; ++b)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&S.full[b])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue_load = [&](int j) {               // one thread: tile j of this CTA -> raw[j & 1]
        const long long roll0 = ((long long)blockIdx.x + (long long)j * gridDim.x) * kTile;
        const int nr = (int)min((long long)kTile, P.nrolls - roll0);
        const uint32_t bytes = (uint32_t)nr * (uint32_t)T * 16u;
        const uint32_t bar = smem_addr(&S.full[j & 1]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_addr(S.raw[j & 1])), "l"(P.rolls + roll0 * T), "r"(bytes), "r"(bar) : "memory");
    };
    if (tid == 0 && ntl > 0) issue_load(0);

    const int Th = ((T + 63) / 64) * 32;         // rows of the first half (a multiple of 32)
    for (int j = 0; j < ntl + 2; ++j) {
        if (tid == 0 && j + 1 < ntl) issue_load(j + 1);      // raw[(j+1)&1] was last read by P1(j-1), before the barrier
        if (warp < kWorkerWarps) {
            const int r = warp >> 1, h = warp & 1;
            const int row_lo = h ? Th : 0, row_hi = h ? T : min(T, Th);
            // ---- P3 of tile j-2 ----
            if (j >= 2) {
                const int jt = j - 2, cb = jt & 1, mb = jt & 3;
                const long long roll0 = ((long long)blockIdx.x + (long long)jt * gridDim.x) * kTile;
                const int nr = (int)min((long long)kTile, P.nrolls - roll0);
                if (r < nr) {
                    const long long obase = (roll0 + r) * (long long)T;
                    const int ndbl = S.ndbl[mb][r];
                    unsigned base = h ? (unsigned)S.nkeep[mb][r][0] : 0u;
                    uint8_t* __restrict__ op = P.pitch + obase;
                    uint8_t* __restrict__ ov = P.velocity + obase;
                    double* __restrict__ os = P.start + obase;
                    double* __restrict__ oe = P.end + obase;
                    const unsigned short* pcode = &S.code[cb][r][row_lo + lane];
                    const float* pclk = &S.clk[cb][r][row_lo + lane];
                    const float* pdur = &S.dur[cb][r][row_lo + lane];
                    const unsigned lt = (1u << lane) - 1u;
#pragma unroll 2
                    for (int i = row_lo + lane; i < row_hi + lane; i += 32, pcode += 32, pclk += 32, pdur += 32) {
                        const unsigned code = i < row_hi ? (unsigned)*pcode : 0xFFFFu;
                        const bool keep = code != 0xFFFFu;
                        const unsigned ball = __ballot_sync(0xffffffffu, keep);
                        if (keep) {
                            const unsigned slot = base + (unsigned)__popc(ball & lt);   // unsigned: one IMAD.WIDE.U32 per address
                            const float d = *pdur;
                            double st, en;
                            if (i < ndbl) {                  // rows that still see the float64 clock
                                const double t64 = S.floor_clock[i];
                                st = __dmul_rn(t64, P.spb64);
                                if (d == floor_dur32) en = __dmul_rn(__dadd_rn(t64, 0.25), P.spb64);
                                else en = (double)__fmul_rn(__fadd_rn((float)t64, d), P.spb32);
                            } else {
                                const float t32 = *pclk;
                                st = (double)__fmul_rn(t32, P.spb32);
                                en = (double)__fmul_rn(__fadd_rn(t32, d), P.spb32);
                            }
                            op[slot] = (uint8_t)(code & 0xFFu);
                            ov[slot] = (uint8_t)(code >> 8);
                            os[slot] = st;
                            oe[slot] = en;
                        }
                        base += (unsigned)__popc(ball);
                    }
                    if (h == 0 && lane == 0)
                        P.counts[roll0 + r] = (S.bad[mb][r][0] | S.bad[mb][r][1]) ? -1 : S.nkeep[mb][r][0] + S.nkeep[mb][r][1];
                }
            }
            // ---- P1 of tile j ----  (branch-free: a row past the end is a gated row on both floors)
            if (j < ntl) {
                const int cb = j & 1, mb = j & 3;
                const long long roll0 = ((long long)blockIdx.x + (long long)j * gridDim.x) * kTile;
                const int nr = (int)min((long long)kTile, P.nrolls - roll0);
                nb_wait(&S.full[cb], (uint32_t)(j >> 1) & 1u);
                if (r < nr) {
                    const float inf32 = __int_as_float(0x7f800000);
                    const float4* __restrict__ src = S.raw[cb] + r * T + row_lo + lane;
                    float* pclk = &S.clk[cb][r][row_lo + lane];
                    float* pdur = &S.dur[cb][r][row_lo + lane];
                    unsigned short* pcode = &S.code[cb][r][row_lo + lane];
                    int first = T, nk = 0;
                    bool anybad = false;
                    const int blk_hi = h ? ((T + 31) / 32) * 32 : Th;        // whole 32-row blocks: pads the clock row
#pragma unroll 2
                    for (int i = row_lo + lane; i < blk_hi; i += 32, src += 32, pclk += 32, pdur += 32, pcode += 32) {
                        float4 q = make_float4(0.0f, -1.0f, -1.0f, -1.0f);   // (pitch, velocity, duration, step)
                        if (i < row_hi) q = *src;
                        const float s32 = unit_to_beats(q.w);
                        const bool nonfloor = s32 > floor_step32;            // max(0.1, .) keeps the python float 0.1
                        const float dv = fmaxf(unit_to_beats(q.z), floor_dur32);   // NaN -> the floor, as `dd > 0.25` is false
                        const bool gate = !(q.y < thr32);                    // utils.py:135; NaN velocity is not gated
                        const float pf = __fmul_rn(__fadd_rn(q.x, 1.0f), 63.5f);
                        const float vf = __fadd_rn(60.0f, __fmul_rn(__fdiv_rn(__fsub_rn(q.y, thr32), vrange32), 67.0f));
                        const bool fin = (fabsf(pf) < inf32) && (fabsf(vf) < inf32);   // int(nan) / int(inf) raises in the reference
                        // clip(int(x), lo, hi) == int(clamp(x, lo, hi)) for truncation toward zero
                        const int pc = __float2int_rz(fminf(fmaxf(pf, 36.0f), 96.0f));
                        const int vel = __float2int_rz(fminf(fmaxf(vf, 0.0f), 127.0f));
                        unsigned code = (unsigned)S.snap[pc - 36] | ((unsigned)vel << 8);
                        code = fin ? code : 0u;
                        code = gate ? code : 0xFFFFu;
                        *pclk = nonfloor ? s32 : floor_step32;               // i < 512 always; rows >= T read as floor steps
                        *pdur = dv;
                        *pcode = (unsigned short)code;
                        first = min(first, nonfloor ? i : T);
                        nk += gate ? 1 : 0;
                        anybad |= gate && !fin;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
                        nk += __shfl_xor_sync(0xffffffffu, nk, o);
                    }
                    const unsigned badball = __ballot_sync(0xffffffffu, anybad);
                    if (lane == 0) { S.first_nf[mb][r][h] = first; S.nkeep[mb][r][h] = nk; S.bad[mb][r][h] = badball != 0; }
                }
            }
        } else if (j >= 1 && j - 1 < ntl) {
            // ---- the onset clock of tile j-1: one lane per roll, strictly in row order ----
            const int jt = j - 1, cb = jt & 1, mb = jt & 3;
            const long long roll0 = ((long long)blockIdx.x + (long long)jt * gridDim.x) * kTile;
            const int nr = (int)min((long long)kTile, P.nrolls - roll0);
            if (lane < nr) {
                float* tex = S.clk[cb][lane];
                const int i0 = min(S.first_nf[mb][lane][0], S.first_nf[mb][lane][1]);   // halves: [0,Th) and [Th,T)
                int ndbl = T;
                if (i0 < T) {
                    ndbl = i0 + 1;                       // rows 0..i0 still see the float64 clock
                    float t32 = __fadd_rn((float)S.floor_clock[i0], tex[i0]);
                    int i = i0 + 1;
                    for (; (i & 7) && i < T; ++i) {
                        const float s = tex[i];
                        tex[i] = t32;
                        t32 = __fadd_rn(t32, s);
                    }
                    // Register-blocked and software-pipelined by hand: one warp issues at most one instruction every other
                    // cycle, so the loop is kept at LDS.128 / 4 FADD / STS.128 per four steps (P1 already resolved the
                    // floors), two blocks per iteration with ping-pong registers (no moves), the loads of the NEXT eight
                    // steps in flight while the eight dependent FADDs of this block run.  Reads run up to 12 floats past
                    // the last row: that is the pad of the clock row.
                    if (i < T) {
                        float* p = tex + i;
                        float4 a0 = *reinterpret_cast<const float4*>(p), a1 = *reinterpret_cast<const float4*>(p + 4);
#define MG_CLOCK8(A0, A1, N0, N1, OFF)                                              \
    {                                                                               \
        N0 = *reinterpret_cast<const float4*>(p + (OFF) + 8);                       \
        N1 = *reinterpret_cast<const float4*>(p + (OFF) + 12);                      \
        float4 o0, o1;                                                              \
        o0.x = t32;                                                                 \
        o0.y = __fadd_rn(o0.x, A0.x);                                               \
        o0.z = __fadd_rn(o0.y, A0.y);                                               \
        o0.w = __fadd_rn(o0.z, A0.z);                                               \
        o1.x = __fadd_rn(o0.w, A0.w);                                               \
        o1.y = __fadd_rn(o1.x, A1.x);                                               \
        o1.z = __fadd_rn(o1.y, A1.y);                                               \
        o1.w = __fadd_rn(o1.z, A1.z);                                               \
        t32 = __fadd_rn(o1.w, A1.w);                                                \
        *reinterpret_cast<float4*>(p + (OFF)) = o0;                                 \
        *reinterpret_cast<float4*>(p + (OFF) + 4) = o1;                             \
    }
                        float4 b0, b1;
#pragma unroll 1
                        for (; i < T; i += 16, p += 16) {      // rows past T (up to 15 of them) are pad: computed, never read
                            MG_CLOCK8(a0, a1, b0, b1, 0)
                            MG_CLOCK8(b0, b1, a0, a1, 8)
                        }
#undef MG_CLOCK8
                    }
                }
                S.ndbl[mb][lane] = ndbl;
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) extract_notes_abs_kernel(const float4* __restrict__ rolls, long long nrowsTotal,
                                                                uint8_t* __restrict__ pitch,
                                                                uint8_t* __restrict__ velocity,
                                                                double* __restrict__ start, double* __restrict__ end,
                                                                int32_t* __restrict__ status) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nrowsTotal; i += stride) {
        const float4 q = __ldg(rolls + i);
        float pc = 0.0f;
        if (isnan(q.x)) {
            if (status) *status = 1;  // int(nan) raises in the reference
        } else {
            pc = q.x < 0.0f ? 0.0f : (q.x > 127.0f ? 127.0f : q.x);  // np.clip(row[0], 0, 127)
        }
        const float m = (q.y < 127.0f) ? q.y : 127.0f;  // min(127, row[1])
        const float w = (m > 1.0f) ? m : 1.0f;          // max(1, .)
        const double dd = (double)q.z, ss = (double)q.w;
        const double dur = (dd > 0.05) ? dd : 0.05;
        const double st = (ss > 0.0) ? ss : 0.0;
        pitch[i] = (uint8_t)__float2int_rz(pc);
        velocity[i] = (uint8_t)__float2int_rz(w);
        start[i] = st;
        end[i] = __dadd_rn(st, dur);
    }
}

// 8(f)-3  MIDIDataset.__getitem__ normalisation   reference src/ae/dataset.py:72-89,105
// Rows are (pitch, start, duration, velocity) in raw units; rows whose pitch is exactly -1 are padding and stay untouched.
// numpy float32 arithmetic, one rounding per operation (weak python scalars keep the array dtype):
//   pitch    = (p / 128) * 2 - 1;   velocity = (clip(v, 0, 127) / 128) * 2 - 1   (np.clip propagates NaN)
//   start    = s / MAX_START_BEAT;  duration = d / MAX_DURATION_BEAT
// then np.nan_to_num(nan=0, posinf=0, neginf=0) over the whole roll (padding rows included).
__device__ __forceinline__ float nan_inf_to_zero(float x) { return fabsf(x) < __int_as_float(0x7f800000) ? x : 0.0f; }
__device__ __forceinline__ float to_unit128(float x) { return __fsub_rn(__fmul_rn(__fmul_rn(x, 0.0078125f), 2.0f), 1.0f); }

__global__ void __launch_bounds__(256) ae_normalize_kernel(const float4* __restrict__ in, float4* __restrict__ out,
                                                           long long nrowsTotal, float max_start, float max_dur) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nrowsTotal; i += stride) {
        float4 q = __ldg(in + i);
        if (q.x != -1.0f) {                                   // NaN != -1: such a row is normalised (then zeroed below)
            const float v = q.w != q.w ? q.w : fminf(fmaxf(q.w, 0.0f), 127.0f);
            q.x = to_unit128(q.x);                            // x / 128.0 == x * 2^-7 bit for bit
            q.y = __fdiv_rn(q.y, max_start);
            q.z = __fdiv_rn(q.z, max_dur);
            q.w = to_unit128(v);
        }
        out[i] = make_float4(nan_inf_to_zero(q.x), nan_inf_to_zero(q.y), nan_inf_to_zero(q.z), nan_inf_to_zero(q.w));
    }
}

int init_once() {
    static int done = 0;  // 0 = not yet, 1 = ok
    if (done) return MG_OK;
    double tab[kMaxRows + 1];
    double t = 0.0;
    for (int k = 0; k <= kMaxRows; ++k) { tab[k] = t; t = t + 0.1; }
    MG_CUDA_OK(cudaMemcpyToSymbol(c_floor_clock, tab, sizeof(tab)));
    MG_CUDA_OK(cudaFuncSetAttribute(extract_notes_gan_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(NotesSmem)));
    done = 1;
    return MG_OK;
}

void snap_lut(uint32_t mask, uint8_t lut[12]) {
    // utils.py:124-128: nearest allowed pitch class by |x - note|, no octave wrap, lowest wins ties
    for (int nio = 0; nio < 12; ++nio) {
        int best = nio, bestd = 1 << 30;
        for (int x = 0; x < 12; ++x) {
            if (!((mask >> x) & 1u)) continue;
            const int d = x > nio ? x - nio : nio - x;
            if (d < bestd) { bestd = d; best = x; }
        }
        lut[nio] = (uint8_t)best;
    }
}

}  // namespace

extern "C" int mg_extract_notes_gan(const float* rolls, long long nrolls, int nrows, double bpm,
                                    uint32_t allowed_mask, int32_t* counts, uint8_t* pitch, uint8_t* velocity,
                                    double* start, double* end, void* stream) {
    MG_REQUIRE(nrolls >= 0 && nrows >= 0 && nrows <= kMaxRows, "extract_notes_gan: nrows must be in [0,%d]", kMaxRows);
    MG_REQUIRE((allowed_mask & 0xFFFu) != 0, "extract_notes_gan: empty scale mask");
    if (nrolls == 0) return MG_OK;
    if (nrows == 0) {  // empty rolls: zero notes each
        MG_REQUIRE(counts, "extract_notes_gan: null pointer");
        MG_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * nrolls, mg::as_stream(stream)));
        return MG_OK;
    }
    MG_REQUIRE(rolls && counts && pitch && velocity && start && end, "extract_notes_gan: null pointer");
    int rc = init_once();
    if (rc != MG_OK) return rc;
    GanParams P;
    P.rolls = reinterpret_cast<const float4*>(rolls);
    P.nrolls = nrolls;
    P.nrows = nrows;
    if (bpm > 180.0) bpm = 180.0;   // bpm = max(60, min(bpm, 180))   utils.py:102
    if (!(bpm > 60.0)) bpm = 60.0;
    P.spb64 = 60.0 / bpm;
    P.spb32 = (float)P.spb64;
    snap_lut(allowed_mask, P.lut);
    P.counts = counts; P.pitch = pitch; P.velocity = velocity; P.start = start; P.end = end;
    mg::ProbeScope probe(mg::PROBE_NOTES, 0.0, (double)nrolls * (nrows * 16.0 + 4.0), mg::as_stream(stream));
    const long long ntiles = (nrolls + kTile - 1) / kTile;
    const int grid = (int)((ntiles < (long long)mg::num_sms()) ? ntiles : (long long)mg::num_sms());
    extract_notes_gan_pipe_kernel<<<grid, kThreadsV2, sizeof(NotesSmem), mg::as_stream(stream)>>>(P);
    MG_LAUNCH_OK();
    return MG_OK;
}

extern "C" int mg_extract_notes_abs(const float* rolls, long long nrolls, int nrows, uint8_t* pitch,
                                    uint8_t* velocity, double* start, double* end, int32_t* status_dev,
                                    void* stream) {
    MG_REQUIRE(nrolls >= 0 && nrows >= 0, "extract_notes_abs: negative size");
    const long long n = nrolls * nrows;
    if (n == 0) return MG_OK;
    MG_REQUIRE(rolls && pitch && velocity && start && end, "extract_notes_abs: null pointer");
    if (status_dev) MG_CUDA_OK(cudaMemsetAsync(status_dev, 0, sizeof(int32_t), mg::as_stream(stream)));
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)mg::num_sms() * 16;
    if (blocks > cap) blocks = cap;
    extract_notes_abs_kernel<<<(int)blocks, 256, 0, mg::as_stream(stream)>>>(
        reinterpret_cast<const float4*>(rolls), n, pitch, velocity, start, end, status_dev);
    MG_LAUNCH_OK();
    return MG_OK;
}

extern "C" int mg_ae_normalize(const float* notes, float* out, long long nrolls, int nrows, double max_start_beat,
                               double max_duration_beat, void* stream) {
    MG_REQUIRE(nrolls >= 0 && nrows >= 0, "ae_normalize: negative size");
    const long long n = nrolls * nrows;
    if (n == 0) return MG_OK;
    MG_REQUIRE(notes && out, "ae_normalize: null pointer");
    MG_REQUIRE(((uintptr_t)notes % 16) == 0 && ((uintptr_t)out % 16) == 0, "ae_normalize: buffers must be 16-byte aligned");
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)mg::num_sms() * 16;
    if (blocks > cap) blocks = cap;
    mg::ProbeScope probe(mg::PROBE_NOTES, 0.0, (double)n * 32.0, mg::as_stream(stream));
    ae_normalize_kernel<<<(int)blocks, 256, 0, mg::as_stream(stream)>>>(reinterpret_cast<const float4*>(notes),
                                                                       reinterpret_cast<float4*>(out), n,
                                                                       (float)max_start_beat, (float)max_duration_beat);
    MG_LAUNCH_OK();
    return MG_OK;
}

// ---- host-buffer entry points: chunked, double-buffered H2D -> kernel -> D2H ----
namespace {
struct HostPipe {
    cudaStream_t st[2] = {nullptr, nullptr};
    void* dev[2] = {nullptr, nullptr};
    size_t cap = 0;
    int ensure(size_t bytes) {
        for (int k = 0; k < 2; ++k)
            if (!st[k]) MG_CUDA_OK(cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking));
        if (bytes > cap) {
            for (int k = 0; k < 2; ++k) {
                if (dev[k]) MG_CUDA_OK(cudaFree(dev[k]));
                dev[k] = nullptr;
                MG_CUDA_OK(cudaMalloc(&dev[k], bytes));
            }
            cap = bytes;
        }
        return MG_OK;
    }
};
HostPipe g_pipe;
constexpr long long kChunkRolls = 16384;  // 128 MiB of rolls per chunk
}  // namespace

extern "C" int mg_extract_notes_gan_host(const float* rolls_host, long long nrolls, int nrows, double bpm,
                                         uint32_t allowed_mask, int32_t* counts_host, uint8_t* pitch_host,
                                         uint8_t* velocity_host, double* start_host, double* end_host) {
    MG_REQUIRE(nrolls >= 0 && nrows >= 0 && nrows <= kMaxRows, "extract_notes_gan_host: bad size");
    if (nrolls == 0) return MG_OK;
    const long long chunk = nrolls < kChunkRolls ? nrolls : kChunkRolls;
    const size_t rowsC = (size_t)chunk * nrows;
    // per-chunk device layout: rolls | start | end | counts | pitch | velocity  (8-byte aligned first)
    const size_t off_start = rowsC * 16, off_end = off_start + rowsC * 8, off_counts = off_end + rowsC * 8;
    const size_t off_pitch = off_counts + (((size_t)chunk * 4 + 15) / 16) * 16, off_vel = off_pitch + rowsC;
    const size_t bytes = off_vel + rowsC;
    int rc = g_pipe.ensure(bytes);
    if (rc != MG_OK) return rc;
    int k = 0;
    for (long long r0 = 0; r0 < nrolls; r0 += chunk, k ^= 1) {
        const long long nr = (nrolls - r0 < chunk) ? nrolls - r0 : chunk;
        const size_t rows = (size_t)nr * nrows;
        char* d = static_cast<char*>(g_pipe.dev[k]);
        cudaStream_t s = g_pipe.st[k];
        MG_CUDA_OK(cudaMemcpyAsync(d, rolls_host + (size_t)r0 * nrows * 4, rows * 16, cudaMemcpyHostToDevice, s));
        rc = mg_extract_notes_gan(reinterpret_cast<float*>(d), nr, nrows, bpm, allowed_mask,
                                  reinterpret_cast<int32_t*>(d + off_counts), reinterpret_cast<uint8_t*>(d + off_pitch),
                                  reinterpret_cast<uint8_t*>(d + off_vel), reinterpret_cast<double*>(d + off_start),
                                  reinterpret_cast<double*>(d + off_end), s);
        if (rc != MG_OK) return rc;
        const size_t o = (size_t)r0 * nrows;
        MG_CUDA_OK(cudaMemcpyAsync(counts_host + r0, d + off_counts, (size_t)nr * 4, cudaMemcpyDeviceToHost, s));
        MG_CUDA_OK(cudaMemcpyAsync(pitch_host + o, d + off_pitch, rows, cudaMemcpyDeviceToHost, s));
        MG_CUDA_OK(cudaMemcpyAsync(velocity_host + o, d + off_vel, rows, cudaMemcpyDeviceToHost, s));
        MG_CUDA_OK(cudaMemcpyAsync(start_host + o, d + off_start, rows * 8, cudaMemcpyDeviceToHost, s));
        MG_CUDA_OK(cudaMemcpyAsync(end_host + o, d + off_end, rows * 8, cudaMemcpyDeviceToHost, s));
    }
    MG_CUDA_OK(cudaStreamSynchronize(g_pipe.st[0]));
    MG_CUDA_OK(cudaStreamSynchronize(g_pipe.st[1]));
    for (long long r = 0; r < nrolls; ++r)
        if (counts_host[r] < 0) {
            mg::set_error("extract_notes_gan: non-finite pitch/velocity in roll %lld (reference raises)", r);
            return MG_ERR_NONFINITE;
        }
    return MG_OK;
}

extern "C" int mg_extract_notes_abs_host(const float* rolls_host, long long nrolls, int nrows, uint8_t* pitch_host,
                                         uint8_t* velocity_host, double* start_host, double* end_host) {
    MG_REQUIRE(nrolls >= 0 && nrows > 0 && nrows <= kMaxRows, "extract_notes_abs_host: bad size");
    if (nrolls == 0) return MG_OK;
    const long long chunk = nrolls < kChunkRolls ? nrolls : kChunkRolls;
    const size_t rowsC = (size_t)chunk * nrows;
    const size_t off_start = rowsC * 16, off_end = off_start + rowsC * 8, off_status = off_end + rowsC * 8;
    const size_t off_pitch = off_status + 16, off_vel = off_pitch + rowsC;
    int rc = g_pipe.ensure(off_vel + rowsC);
    if (rc != MG_OK) return rc;
    int32_t status[2] = {0, 0};
    int k = 0, bad = 0;
    for (long long r0 = 0; r0 < nrolls; r0 += chunk, k ^= 1) {
        const long long nr = (nrolls - r0 < chunk) ? nrolls - r0 : chunk;
        const size_t rows = (size_t)nr * nrows;
        char* d = static_cast<char*>(g_pipe.dev[k]);
        cudaStream_t s = g_pipe.st[k];
        MG_CUDA_OK(cudaStreamSynchronize(s));  // status[k] of the previous use of this buffer is final
        bad |= status[k];
        MG_CUDA_OK(cudaMemcpyAsync(d, rolls_host + (size_t)r0 * nrows * 4, rows * 16, cudaMemcpyHostToDevice, s));
        rc = mg_extract_notes_abs(reinterpret_cast<float*>(d), nr, nrows, reinterpret_cast<uint8_t*>(d + off_pitch),
                                  reinterpret_cast<uint8_t*>(d + off_vel), reinterpret_cast<double*>(d + off_start),
                                  reinterpret_cast<double*>(d + off_end), reinterpret_cast<int32_t*>(d + off_status), s);
        if (rc != MG_OK) return rc;
        const size_t o = (size_t)r0 * nrows;
        MG_CUDA_OK(cudaMemcpyAsync(&status[k], d + off_status, 4, cudaMemcpyDeviceToHost, s));
        MG_CUDA_OK(cudaMemcpyAsync(pitch_host + o, d + off_pitch, rows, cudaMemcpyDeviceToHost, s));
        MG_CUDA_OK(cudaMemcpyAsync(velocity_host + o, d + off_vel, rows, cudaMemcpyDeviceToHost, s));
        MG_CUDA_OK(cudaMemcpyAsync(start_host + o, d + off_start, rows * 8, cudaMemcpyDeviceToHost, s));
        MG_CUDA_OK(cudaMemcpyAsync(end_host + o, d + off_end, rows * 8, cudaMemcpyDeviceToHost, s));
    }
    MG_CUDA_OK(cudaStreamSynchronize(g_pipe.st[0]));
    MG_CUDA_OK(cudaStreamSynchronize(g_pipe.st[1]));
    bad |= status[0] | status[1];
    if (bad) {
        mg::set_error("extract_notes_abs: NaN pitch (reference raises ValueError)");
        return MG_ERR_NONFINITE;
    }
    return MG_OK;
}
