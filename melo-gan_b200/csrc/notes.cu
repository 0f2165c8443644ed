// notes.cu -- batched note extraction (velocity gate + onset clock + quantisation) on sm_100a.
//
// N-1  save_piano_roll_to_midi row loop   reference src/gan/utils.py:130-155
// N-2  tools/roll_to_midi.py row loop     reference tools/roll_to_midi.py:10-21
//
// HBM-bound byte/float work: per roll 8 KiB in, <= 512*(1+1+8+8) B + 4 B out.  The only
// sequential part is the onset clock: the reference accumulates `current_time_beats` one row
// at a time in float32 (float64 while only floor steps were added), and float addition is not
// associative, so a tree/warp scan is not bit-exact.  Layout of the work:
//   phase 1  thread-per-row, coalesced float4 loads: gate, pitch, velocity, step, duration
//            -> shared memory (the roll never goes back to HBM);
//   phase 2  the clock: one LANE PER ROLL walks that roll's 512 steps from shared memory in
//            order (serial over rows, parallel over rolls: 16 chains per warp instruction);
//   phase 3  warp-per-roll compaction: __ballot_sync + popc gives every surviving row its
//            output slot (integer prefix, exact), then coalesced stores of the note fields.
// All float32 arithmetic uses __f*_rn intrinsics so nvcc cannot contract mul+add into FMA
// (numpy rounds after every operation).
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kMaxRows = 512;
constexpr int kRollsPerCta = 16;
constexpr int kThreads = 256;
constexpr int kTexStride = kMaxRows + 1;  // +1: lanes of phase 2 (one per roll) hit distinct banks

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
    return __fmul_rn(__fdiv_rn(__fadd_rn(x, 1.0f), 2.0f), 4.0f);
}

__device__ __forceinline__ int trunc_clip(float x, int lo, int hi) {
    // int(x) then np.clip(., lo, hi); x is finite here.  Saturate first: the clip makes it equivalent.
    x = fminf(fmaxf(x, -1.0e9f), 1.0e9f);
    int v = __float2int_rz(x);
    return min(max(v, lo), hi);
}

__global__ void __launch_bounds__(kThreads, 6) extract_notes_gan_kernel(GanParams P) {
    // Only the onset clock lives in shared memory (2 KB per roll): the serial phase 2 keeps one lane per roll busy, so its
    // throughput is the number of rolls resident per SM; pitch / velocity / duration are recomputed in phase 3 from the
    // rolls (an L2 hit: the tile was read a few microseconds earlier) instead of being parked in 6 more bytes per row.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_tex = reinterpret_cast<float*>(smem_raw);                       // [16][513] step -> exclusive clock
    int* s_ndbl = reinterpret_cast<int*>(s_tex + kRollsPerCta * kTexStride); // [16] rows on the float64 clock

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = P.nrows;
    const float thr32 = (float)-0.2;            // VELOCITY_THRESHOLD as a weak scalar
    const float vrange32 = (float)(1.0 - -0.2); // utils.py:143
    const float floor_step32 = (float)0.1, floor_dur32 = (float)0.25;
    const long long ntiles = (P.nrolls + kRollsPerCta - 1) / kRollsPerCta;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long roll0 = tile * kRollsPerCta;
        const int nr = (int)min((long long)kRollsPerCta, P.nrolls - roll0);

        // ---- phase 1: the step column, rows of the tile are contiguous in HBM ----
        const int total = nr * T;
        const float4* src = P.rolls + roll0 * T;
#pragma unroll 4
        for (int idx = tid; idx < total; idx += kThreads) {
            const float4 q = __ldg(src + idx);  // (pitch, velocity, duration, step)
            const int r = idx / T, i = idx - r * T;
            const float s32 = unit_to_beats(q.w);
            s_tex[r * kTexStride + i] = (s32 > floor_step32) ? s32 : -1.0f;  // max(0.1, .) keeps python 0.1
        }
        __syncthreads();

        // ---- phase 2: the onset clock, one lane per roll, strictly in row order ----
        if (warp == 0 && lane < nr) {
            float* tex = s_tex + lane * kTexStride;
            int i = 0;
            // float64 prefix: t is floor_clock[i] while every step so far was the python-float floor
            while (i < T && tex[i] < 0.0f) ++i;
            int ndbl = T;
            if (i < T) {
                ndbl = i + 1;  // rows 0..i still see the float64 clock
                float t32 = __fadd_rn((float)c_floor_clock[i], tex[i]);
                for (++i; i < T; ++i) {
                    const float s = tex[i];
                    tex[i] = t32;
                    t32 = __fadd_rn(t32, s < 0.0f ? floor_step32 : s);
                }
            }
            s_ndbl[lane] = ndbl;
        }
        __syncthreads();

        // ---- phase 3: per-row pitch / velocity / duration, compaction, onset / offset; one warp per roll ----
        for (int r = warp; r < nr; r += kThreads / 32) {
            const long long obase = (roll0 + r) * (long long)T;
            const int ndbl = s_ndbl[r];
            int base = 0;
            unsigned bad_any = 0;
            for (int i0 = 0; i0 < T; i0 += 32) {
                const int i = i0 + lane;
                unsigned short code = 0xFFFFu;  // gated (or past the end)
                float d32 = -1.0f;
                bool bad = false;
                if (i < T) {
                    const float4 q = __ldg(src + r * T + i);
                    const float dd = unit_to_beats(q.z);
                    d32 = (dd > floor_dur32) ? dd : -1.0f;
                    if (!(q.y < thr32)) {           // utils.py:135; NaN velocity is not gated
                        const float pf = __fmul_rn(__fadd_rn(q.x, 1.0f), 63.5f);
                        const float vf = __fadd_rn(60.0f, __fmul_rn(__fdiv_rn(__fsub_rn(q.y, thr32), vrange32), 67.0f));
                        if (!isfinite(pf) || !isfinite(vf)) {
                            bad = true;  // int(nan)/int(inf) raises in the reference
                            code = 0;
                        } else {
                            const int pc = trunc_clip(pf, 36, 96);
                            const int pit = (pc / 12) * 12 + P.lut[pc % 12];
                            const int vel = trunc_clip(vf, 0, 127);
                            code = (unsigned short)(pit | (vel << 8));
                        }
                    }
                }
                bad_any |= __ballot_sync(0xffffffffu, bad);
                const bool keep = code != 0xFFFFu;
                const unsigned ball = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const int slot = base + __popc(ball & ((1u << lane) - 1u));
                    double st, en;
                    if (i < ndbl) {
                        const double t64 = c_floor_clock[i];
                        st = __dmul_rn(t64, P.spb64);
                        if (d32 < 0.0f) en = __dmul_rn(__dadd_rn(t64, 0.25), P.spb64);
                        else en = (double)__fmul_rn(__fadd_rn((float)t64, d32), P.spb32);
                    } else {
                        const float t32 = s_tex[r * kTexStride + i];
                        st = (double)__fmul_rn(t32, P.spb32);
                        en = (double)__fmul_rn(__fadd_rn(t32, d32 < 0.0f ? floor_dur32 : d32), P.spb32);
                    }
                    P.pitch[obase + slot] = (uint8_t)(code & 0xFF);
                    P.velocity[obase + slot] = (uint8_t)(code >> 8);
                    P.start[obase + slot] = st;
                    P.end[obase + slot] = en;
                }
                base += __popc(ball);
            }
            if (lane == 0) P.counts[roll0 + r] = bad_any ? -1 : base;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// N-1, pipelined form (the one mg_extract_notes_gan launches).
//
// The three-phase kernel above keeps one lane per roll busy for the whole of phase 2 while every other thread of the
// CTA waits at a barrier, and reads the rolls twice (HBM, then L2).  Here one persistent CTA per SM runs the three
// phases of three DIFFERENT tiles at the same time:
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
constexpr int kClkStride = kMaxRows + 4;         // floats; 516 % 32 = 4: the clock lanes' LDS.128 hit disjoint banks

struct alignas(16) NotesSmem {
    float4 raw[2][kTile * kMaxRows];             // 2 x 64 KB, written by bulk copies
    float clk[2][kTile][kClkStride];             // P1: step (or -1 = floor); clock warp: exclusive onset clock
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
                    int base = h ? S.nkeep[mb][r][0] : 0;
                    uint8_t* __restrict__ op = P.pitch + obase;
                    uint8_t* __restrict__ ov = P.velocity + obase;
                    double* __restrict__ os = P.start + obase;
                    double* __restrict__ oe = P.end + obase;
                    for (int i0 = row_lo; i0 < row_hi; i0 += 32) {
                        const int i = i0 + lane;
                        const unsigned code = i < row_hi ? S.code[cb][r][i] : 0xFFFFu;
                        const bool keep = code != 0xFFFFu;
                        const unsigned ball = __ballot_sync(0xffffffffu, keep);
                        if (keep) {
                            const int slot = base + __popc(ball & ((1u << lane) - 1u));
                            const float d = S.dur[cb][r][i];
                            double st, en;
                            if (i < ndbl) {                  // rows that still see the float64 clock
                                const double t64 = S.floor_clock[i];
                                st = __dmul_rn(t64, P.spb64);
                                if (d == floor_dur32) en = __dmul_rn(__dadd_rn(t64, 0.25), P.spb64);
                                else en = (double)__fmul_rn(__fadd_rn((float)t64, d), P.spb32);
                            } else {
                                const float t32 = S.clk[cb][r][i];
                                st = (double)__fmul_rn(t32, P.spb32);
                                en = (double)__fmul_rn(__fadd_rn(t32, d), P.spb32);
                            }
                            op[slot] = (uint8_t)(code & 0xFFu);
                            ov[slot] = (uint8_t)(code >> 8);
                            os[slot] = st;
                            oe[slot] = en;
                        }
                        base += __popc(ball);
                    }
                    if (h == 0 && lane == 0)
                        P.counts[roll0 + r] = (S.bad[mb][r][0] | S.bad[mb][r][1]) ? -1 : S.nkeep[mb][r][0] + S.nkeep[mb][r][1];
                }
            }
            // ---- P1 of tile j ----
            if (j < ntl) {
                const int cb = j & 1, mb = j & 3;
                const long long roll0 = ((long long)blockIdx.x + (long long)j * gridDim.x) * kTile;
                const int nr = (int)min((long long)kTile, P.nrolls - roll0);
                nb_wait(&S.full[cb], (uint32_t)(j >> 1) & 1u);
                if (r < nr) {
                    const float4* __restrict__ src = S.raw[cb] + r * T;
                    int first = T, nk = 0;
                    unsigned anybad = 0;
                    const int blk_hi = h ? ((T + 31) / 32) * 32 : Th;        // whole 32-row blocks: pads the clock row
                    for (int i0 = row_lo; i0 < blk_hi; i0 += 32) {
                        const int i = i0 + lane;
                        unsigned code = 0xFFFFu;
                        float sv = -1.0f, dv = floor_dur32;
                        bool bad = false;
                        if (i < row_hi) {
                            const float4 q = src[i];                         // (pitch, velocity, duration, step)
                            const float s32 = unit_to_beats(q.w);
                            sv = (s32 > floor_step32) ? s32 : -1.0f;         // max(0.1, .) keeps the python float 0.1
                            const float dd = unit_to_beats(q.z);
                            dv = (dd > floor_dur32) ? dd : floor_dur32;
                            if (!(q.y < thr32)) {                            // utils.py:135; NaN velocity is not gated
                                const float pf = __fmul_rn(__fadd_rn(q.x, 1.0f), 63.5f);
                                const float vf = __fadd_rn(60.0f, __fmul_rn(__fdiv_rn(__fsub_rn(q.y, thr32), vrange32), 67.0f));
                                if (!(fabsf(pf) < __int_as_float(0x7f800000)) || !(fabsf(vf) < __int_as_float(0x7f800000))) {
                                    bad = true;                              // int(nan) / int(inf) raises in the reference
                                    code = 0;
                                } else {
                                    // clip(int(x), lo, hi) == int(clamp(x, lo, hi)) for truncation toward zero
                                    const int pc = __float2int_rz(fminf(fmaxf(pf, 36.0f), 96.0f));
                                    const int vel = __float2int_rz(fminf(fmaxf(vf, 0.0f), 127.0f));
                                    code = (unsigned)S.snap[pc - 36] | ((unsigned)vel << 8);
                                }
                            }
                            S.dur[cb][r][i] = dv;
                            S.code[cb][r][i] = (unsigned short)code;
                        }
                        S.clk[cb][r][i] = sv;                                // i < 512 always; rows >= T read as floor steps
                        const unsigned nf = __ballot_sync(0xffffffffu, sv >= 0.0f);
                        if (first == T && nf) first = i0 + __ffs(nf) - 1;
                        nk += __popc(__ballot_sync(0xffffffffu, code != 0xFFFFu));
                        anybad |= __ballot_sync(0xffffffffu, bad);
                    }
                    if (lane == 0) { S.first_nf[mb][r][h] = first; S.nkeep[mb][r][h] = nk; S.bad[mb][r][h] = anybad != 0; }
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
                    for (; (i & 3) && i < T; ++i) {
                        const float s = tex[i];
                        tex[i] = t32;
                        t32 = __fadd_rn(t32, s < 0.0f ? floor_step32 : s);
                    }
#pragma unroll 2
                    for (; i < T; i += 4) {              // rows T..T+3 of the padded clock row hold floor steps
                        const float4 s4 = *reinterpret_cast<const float4*>(tex + i);
                        const float a0 = s4.x < 0.0f ? floor_step32 : s4.x, a1 = s4.y < 0.0f ? floor_step32 : s4.y;
                        const float a2 = s4.z < 0.0f ? floor_step32 : s4.z, a3 = s4.w < 0.0f ? floor_step32 : s4.w;
                        float4 o;
                        o.x = t32;
                        o.y = __fadd_rn(o.x, a0);
                        o.z = __fadd_rn(o.y, a1);
                        o.w = __fadd_rn(o.z, a2);
                        t32 = __fadd_rn(o.w, a3);
                        *reinterpret_cast<float4*>(tex + i) = o;
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

constexpr size_t kGanSmem = sizeof(float) * kRollsPerCta * kTexStride + sizeof(int) * kRollsPerCta;

int init_once() {
    static int done = 0;  // 0 = not yet, 1 = ok
    if (done) return MG_OK;
    double tab[kMaxRows + 1];
    double t = 0.0;
    for (int k = 0; k <= kMaxRows; ++k) { tab[k] = t; t = t + 0.1; }
    MG_CUDA_OK(cudaMemcpyToSymbol(c_floor_clock, tab, sizeof(tab)));
    MG_CUDA_OK(cudaFuncSetAttribute(extract_notes_gan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)kGanSmem));
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
    static const bool three_phase = getenv("MELOGAN_NOTES_3PHASE") != nullptr;     // A/B: the unpipelined kernel
    if (three_phase) {
        const long long ntiles = (nrolls + kRollsPerCta - 1) / kRollsPerCta;
        const int grid = (int)((ntiles < (long long)mg::num_sms() * 6) ? ntiles : (long long)mg::num_sms() * 6);
        extract_notes_gan_kernel<<<grid, kThreads, kGanSmem, mg::as_stream(stream)>>>(P);
    } else {
        const long long ntiles = (nrolls + kTile - 1) / kTile;
        const int grid = (int)((ntiles < (long long)mg::num_sms()) ? ntiles : (long long)mg::num_sms());
        extract_notes_gan_pipe_kernel<<<grid, kThreadsV2, sizeof(NotesSmem), mg::as_stream(stream)>>>(P);
    }
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
