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
    const long long ntiles = (nrolls + kRollsPerCta - 1) / kRollsPerCta;
    const int grid = (int)((ntiles < (long long)mg::num_sms() * 6) ? ntiles : (long long)mg::num_sms() * 6);
    mg::ProbeScope probe(mg::PROBE_NOTES, 0.0, (double)nrolls * (nrows * 16.0 + 4.0), mg::as_stream(stream));
    extract_notes_gan_kernel<<<grid, kThreads, kGanSmem, mg::as_stream(stream)>>>(P);
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
