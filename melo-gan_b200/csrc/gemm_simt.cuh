// gemm_simt.cuh -- CUDA-core (FFMA, fp32 accumulate) implicit-GEMM kernels.
//
// Two kernel families cover every contraction of the hot path in channels-last
// layout X[b, l, c]:
//
//   tap-GEMM   Out[b, m, n] = epi( alpha * sum_t sum_k A[b, m*a_mstride + a_toff[t] + k] * W[t][n][k] )
//              linear fwd/dgrad, conv1d fwd (stride 1|2), conv-transpose1d as two sub-pixel phases,
//              and every dgrad (a dgrad of a strided conv IS a two-phase transposed conv and vice versa).
//              "A" is addressed by a flat per-sample element index; indices outside [0, a_valid) read
//              as zero, which is the conv zero padding.
//   wgrad      dW[t][n][k] += alpha * sum_{b,m} G[b, m, n] * A[b, m*a_mstride + a_toff[t] + k]
//              reduction over all positions, split across CTAs, accumulated with fp32 atomics straight
//              into the (reference-layout) gradient buffer.
//
// Weights are read/written through generic (tap, n, k) strides, so the reference's own layouts
// (Linear [N][K], Conv1d [Co][Ci][k], ConvTranspose1d [Ci][Co][k]) are used in place: no packing.
// These kernels are the fp32 "parity mode" engine and, in bf16 mode, still run the layers that are
// not tensor-core shaped (C_in = 4, C_out = 4, K = 6, vectors); the big contractions then go to
// the tcgen05 kernels in gemm_tc.cuh.
#pragma once
#include "common.cuh"

namespace mg {

constexpr int kMaxTaps = 5;

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_GELU = 3 };
// multiply the result by f'(.) reconstructed from a saved tensor with the output's indexing
enum MulMode { MUL_NONE = 0, MUL_LRELU_SIGN = 1, MUL_RELU_SIGN = 2, MUL_VALUE = 3 };

struct TapGemmArgs {
    const void* A; long long a_bstride; int a_mstride; int a_valid; int a_toff[kMaxTaps];
    int ntaps; int K;
    const float* W; int w_toff[kMaxTaps]; int w_nstride; int w_kstride;
    int n_perm_q; int n_perm_p;   // physical n = (n % q) * p + n / q   (q == 0: identity)
    int k_perm_q; int k_perm_p;   // same for the reduction index of W only (A stays logical; ntaps == 1)
    void* Out; long long o_bstride; int o_mstride; int o_off;
    int B; int Mper; int N;
    const float* bias;            // indexed by physical n
    const float* col_scale;       // optional per-n multiplier applied to the accumulator (folded BatchNorm)
    int act;
    const void* mul_src; int mul_mode;   // same indexing as Out (element type TMSK, by default Out's)
    void* aux;                    // ACT_GELU only: gelu'(z) written with Out's indexing/type
    const float* row_scale;       // optional per-row (b*Mper+m) multiplier applied before everything else
    float alpha;
    int accumulate;               // Out += result (float outputs only)
    // optional fused AdaptiveAvgPool1d(1): pool_out[b, n] = pool_scale * sum_m Out[b, m, n] (sums of the STORED values).  Only
    // the weight-stationary tensor-core kernels do it (from their staging tile); *pool_done is set to 1 when they did, else the
    // caller runs pool_rows_kernel over Out as before.
    float* pool_out; float pool_scale; int* pool_done;
    int pool_only;                // with pool_out: if the pooling is fused, Out itself need not be stored (nothing else reads it)
    // optional fused column sums (bias gradients): colsum_out[n] += sum of the STORED Out[b, m, n] over the flattened rows
    // r = b * Mper + m < colsum_rows (a multiple of 128), by atomicAdd; same contract as pool_done for *colsum_done
    float* colsum_out; long long colsum_rows; int* colsum_done;
    // optional fused BatchNorm batch statistics of a float32 output: stats_out[n] += sum Out[., ., n], stats_out[N + n] += sum
    // Out[., ., n]^2 over all rows, by atomicAdd into a zeroed [2][N] buffer; same contract for *stats_done
    float* stats_out; int* stats_done;
    // set by the fp32-on-tensor-cores path only (gemm_tc.cuh, try_split_tapgemm): K is 6 x the layer's K, A holds the six
    // bf16 operand parts per row, and the pack kernel writes the matching six parts of every fp32 weight row
    int w_split;
};

struct WgradArgs {
    const void* G; long long g_bstride; int g_mstride; int g_off;    // G[b, m, n], n contiguous
    const void* A; long long a_bstride; int a_mstride; int a_valid; int a_toff[kMaxTaps];
    int ntaps; int K;
    float* dW; int w_toff[kMaxTaps]; int w_nstride; int w_kstride;
    int n_perm_q; int n_perm_p;
    int B; int Mper; int N;
    float alpha;
    int row_begin, row_end;       // reduce over flattened rows r = b*Mper + m in [row_begin, row_end)
};

__device__ __forceinline__ float ld_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_from_float(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_from_float(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
__device__ __forceinline__ void ld4(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&q.x);
    const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&q.y);
    v[0] = __low2float(lo); v[1] = __high2float(lo); v[2] = __low2float(hi); v[3] = __high2float(hi);
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
    uint2 q;
    q.x = *reinterpret_cast<unsigned*>(&lo);
    q.y = *reinterpret_cast<unsigned*>(&hi);
    *reinterpret_cast<uint2*>(p) = q;
}

__device__ __forceinline__ int perm_index(int n, int q, int p) { return q ? (n % q) * p + n / q : n; }

// exact (erf) GELU and its derivative, as nn.GELU() default
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// bf16-mode epilogues: GELU and its derivative from ONE exp (erf by Abramowitz-Stegun 7.1.26, |err| < 1.5e-7,
// which is far below the bf16 storage rounding of the result); about 4x fewer instructions than erff + expf.
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void gelu_fast(float x, float& y, float& dy) {
    // flush-to-zero MUFU forms: the default __expf / __fdividef wrap each MUFU in a denormal range fix-up (4 extra
    // instructions each) that cannot trigger here -- the exponent argument is <= 0 and 1 + 0.33 |x| >= 1
    // constants folded so that the chain is 17 instructions: exp(-x^2/2) = ex2(x * x * -log2(e)/2), the Abramowitz-Stegun
    // argument |x|/sqrt(2) inside the FFMA (0.3275911 / sqrt 2), the 0.5 of Phi inside the polynomial coefficients
    const float e = ex2_ftz(x * x * -0.72134752044448170f);    // exp(-x^2 / 2)
    const float t = rcp_ftz(fmaf(0.23164189445300537f, fabsf(x), 1.0f));
    const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 0.5307027145f, -0.7265760135f), 0.7107068705f), -0.142248368f),
                                0.127414796f);
    const float h = poly * e;                                  // 0.5 * erfc(|x| / sqrt 2) = Phi(-|x|)
    const float cdf = x > 0.0f ? 1.0f - h : h;
    y = x * cdf;
    dy = fmaf(x * 0.3989422804014327f, e, cdf);
}

// ------------------------------------------------------------------------------------------------
// tap-GEMM
// ------------------------------------------------------------------------------------------------
template <typename TA, typename TO, typename TMSK, int BM, int BN, int TM, int TN, int NBUF, bool VEC>
__global__ void __launch_bounds__(256) tapgemm_kernel(const TapGemmArgs P) {
    constexpr int BK = 16;
    static_assert((BM / TM) * (BN / TN) == 256, "256 threads per CTA");
    static_assert(TM % 4 == 0 && TN == 4, "thread tile");
    static_assert(NBUF == 1 || NBUF == 2, "single or double buffered");
    __shared__ __align__(16) float As[NBUF][BK][BM + 4];
    __shared__ __align__(16) float Bs[NBUF][BK][BN + 4];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const long long rows = (long long)P.B * P.Mper;
    const long long row0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int Ktot = P.ntaps * P.K;
    const TA* __restrict__ Abase = static_cast<const TA*>(P.A);

    // ---- A loader: thread owns (BM*BK)/256 elements: ROWS_PER_THREAD rows x A_KPT consecutive kk ----
    constexpr int A_ROWS_PT = BM >= 128 ? BM / 128 : 1;   // BM=64/128 -> 1 row, BM=512 -> 4 rows
    constexpr int A_RSPAN = BM >= 128 ? 128 : 64;         // rows covered by one pass of the CTA's threads
    constexpr int A_KPT = BM >= 128 ? 8 : 4;              // consecutive kk per thread
    const int a_kk = (tid / A_RSPAN) * A_KPT;
    long long a_rowoff[A_ROWS_PT];                 // base element offset (b*a_bstride), or -1 if row is out of range
    int a_midx[A_ROWS_PT];
#pragma unroll
    for (int i = 0; i < A_ROWS_PT; ++i) {
        const long long r = row0 + (tid % A_RSPAN) + i * A_RSPAN;
        if (r < rows) {
            const long long b = r / P.Mper;
            a_rowoff[i] = b * P.a_bstride;
            a_midx[i] = (int)(r - b * P.Mper) * P.a_mstride;
        } else {
            a_rowoff[i] = -1; a_midx[i] = 0;
        }
    }
    // ---- B loader: n = tid % BN..., each thread loads (BK*BN)/256 elements ----
    constexpr int B_PT = (BK * BN) / 256;          // BN=64 -> 4, BN=8 -> 0.5 (handled below)

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    auto load_tiles = [&](int kk0, int buf) {
        // A tile
#pragma unroll
        for (int i = 0; i < A_ROWS_PT; ++i) {
            const int rloc = (tid % A_RSPAN) + i * A_RSPAN;
            float v[A_KPT];
#pragma unroll
            for (int h = 0; h < A_KPT / 4; ++h) {
                const int kk = kk0 + a_kk + h * 4;
                float q[4] = {0.f, 0.f, 0.f, 0.f};
                if (a_rowoff[i] >= 0 && kk < Ktot) {
                    if (VEC) {
                        const int t = kk / P.K, k = kk - t * P.K;
                        const int idx = a_midx[i] + P.a_toff[t] + k;
                        if (idx >= 0 && idx < P.a_valid) ld4(Abase + a_rowoff[i] + idx, q);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int kke = kk + e;
                            if (kke < Ktot) {
                                const int t = kke / P.K, k = kke - t * P.K;
                                const int idx = a_midx[i] + P.a_toff[t] + k;
                                if (idx >= 0 && idx < P.a_valid) q[e] = ld_as_float(Abase + a_rowoff[i] + idx);
                            }
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) v[h * 4 + e] = q[e];
            }
#pragma unroll
            for (int e = 0; e < A_KPT; ++e) As[buf][a_kk + e][rloc] = v[e];
        }
        // B tile: element (kk, n) = W[w_toff[t] + nphys*w_nstride + kphys*w_kstride]
        if (B_PT >= 1) {
            const int n = tid % BN;
            const int kb = (tid / BN) * (B_PT >= 1 ? B_PT : 1);
            const int ng = n0 + n;
            const int nphys = perm_index(ng, P.n_perm_q, P.n_perm_p);
#pragma unroll
            for (int e = 0; e < (B_PT >= 1 ? B_PT : 1); ++e) {
                const int kk = kk0 + kb + e;
                float w = 0.0f;
                if (ng < P.N && kk < Ktot) {
                    const int t = kk / P.K, k = kk - t * P.K;
                    w = __ldg(P.W + P.w_toff[t] + (long long)nphys * P.w_nstride +
                              (long long)perm_index(k, P.k_perm_q, P.k_perm_p) * P.w_kstride);
                }
                Bs[buf][kb + e][n] = w;
            }
        } else {
            if (tid < BK * BN) {
                const int n = tid % BN, kk = kk0 + tid / BN;
                const int ng = n0 + n;
                float w = 0.0f;
                if (ng < P.N && kk < Ktot) {
                    const int t = kk / P.K, k = kk - t * P.K;
                    w = __ldg(P.W + P.w_toff[t] + (long long)perm_index(ng, P.n_perm_q, P.n_perm_p) * P.w_nstride +
                              (long long)perm_index(k, P.k_perm_q, P.k_perm_p) * P.w_kstride);
                }
                Bs[buf][tid / BN][n] = w;
            }
        }
    };

    const int nchunks = (Ktot + BK - 1) / BK;
    load_tiles(0, 0);
    __syncthreads();
    for (int c = 0; c < nchunks; ++c) {
        const int buf = (NBUF == 2) ? (c & 1) : 0;
        if (NBUF == 2 && c + 1 < nchunks) load_tiles((c + 1) * BK, buf ^ 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                const float4 q = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + i]);
                a[i] = q.x; a[i + 1] = q.y; a[i + 2] = q.z; a[i + 3] = q.w;
            }
            {
                const float4 q = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * TN]);
                b[0] = q.x; b[1] = q.y; b[2] = q.z; b[3] = q.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
        if (NBUF == 1 && c + 1 < nchunks) {
            load_tiles((c + 1) * BK, 0);
            __syncthreads();
        }
    }

    // ---- epilogue ----
    TO* __restrict__ Obase = static_cast<TO*>(P.Out);
    const TMSK* __restrict__ Mbase = static_cast<const TMSK*>(P.mul_src);
    TO* __restrict__ Xbase = static_cast<TO*>(P.aux);
    const int nb = n0 + tx * TN;
    float bias[TN], cscale[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        bias[j] = (P.bias && nb + j < P.N) ? __ldg(P.bias + perm_index(nb + j, P.n_perm_q, P.n_perm_p)) : 0.0f;
        cscale[j] = (P.col_scale && nb + j < P.N) ? __ldg(P.col_scale + nb + j) : 1.0f;
    }
    const bool vec_out = VEC && (nb + TN <= P.N);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const long long r = row0 + ty * TM + i;
        if (r >= rows) continue;
        const long long b = r / P.Mper;
        const int m = (int)(r - b * P.Mper);
        const long long o = b * P.o_bstride + (long long)m * P.o_mstride + P.o_off + nb;
        const float rs = P.row_scale ? __ldg(P.row_scale + r) : 1.0f;
        float v[TN], g[TN], ms[TN];
        if (P.mul_mode != MUL_NONE) {
            if (vec_out) ld4(Mbase + o, ms);
            else
#pragma unroll
                for (int j = 0; j < TN; ++j) ms[j] = (nb + j < P.N) ? ld_as_float(Mbase + o + j) : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            float x = acc[i][j] * (rs * P.alpha * cscale[j]) + bias[j];
            g[j] = 0.0f;
            if (P.act == ACT_RELU) x = fmaxf(x, 0.0f);
            else if (P.act == ACT_LRELU) x = x > 0.0f ? x : 0.2f * x;
            else if (P.act == ACT_GELU) { g[j] = gelu_grad_f(x); x = gelu_f(x); }
            if (P.mul_mode == MUL_LRELU_SIGN) x *= (ms[j] > 0.0f ? 1.0f : 0.2f);
            else if (P.mul_mode == MUL_RELU_SIGN) x *= (ms[j] > 0.0f ? 1.0f : 0.0f);
            else if (P.mul_mode == MUL_VALUE) x *= ms[j];
            v[j] = x;
        }
        if (vec_out) {
            if (P.accumulate) {
                float old[TN];
                ld4(Obase + o, old);
#pragma unroll
                for (int j = 0; j < TN; ++j) v[j] += old[j];
            }
            st4(Obase + o, v);
            if (P.aux) st4(Xbase + o, g);
        } else {
#pragma unroll
            for (int j = 0; j < TN; ++j)
                if (nb + j < P.N) {
                    float x = v[j];
                    if (P.accumulate) x += ld_as_float(Obase + o + j);
                    st_from_float(Obase + o + j, x);
                    if (P.aux) st_from_float(Xbase + o + j, g[j]);
                }
        }
    }
}

namespace tc {   // tensor-core paths (gemm_tc.cuh); return 1 = launched, 0 = shape does not qualify, <0 = error
template <typename TA, typename TO, typename TMSK>
int try_tc_tapgemm(const TapGemmArgs& P, cudaStream_t st);
template <typename TG, typename TA>
int try_tc_wgrad(const WgradArgs& P, cudaStream_t st);
}  // namespace tc
namespace thin {   // bandwidth-class kernels for the 4-channel layers (thin.cuh); same return convention
template <typename TA, typename TO, typename TMSK>
int try_thin_tapgemm(const TapGemmArgs& P, cudaStream_t st);
template <typename TG, typename TA>
int try_thin_wgrad(const WgradArgs& P, cudaStream_t st);
}  // namespace thin

template <typename TA, typename TO, typename TMSK = TO>
int launch_tapgemm(const TapGemmArgs& P, cudaStream_t st) {
    const long long rows = (long long)P.B * P.Mper;
    if (rows == 0 || P.N == 0) return MG_OK;
    {
        int r = tc::try_tc_tapgemm<TA, TO, TMSK>(P, st);
        if (r == 0) r = thin::try_thin_tapgemm<TA, TO, TMSK>(P, st);
        if (r != 0) return r < 0 ? r : MG_OK;
    }
    const size_t ea = sizeof(TA), eo = sizeof(TO);
    bool vec = (P.K % 4 == 0) && (P.a_mstride % 4 == 0) && (P.a_bstride % 4 == 0) && (P.a_valid % 4 == 0) &&
               (P.N % 4 == 0) && (P.o_mstride % 4 == 0) && (P.o_bstride % 4 == 0) && (P.o_off % 4 == 0) &&
               (((uintptr_t)P.A) % (4 * ea) == 0) && (((uintptr_t)P.Out) % (4 * eo) == 0);
    for (int t = 0; t < P.ntaps; ++t) vec = vec && (P.a_toff[t] % 4 == 0);
    if (P.mul_src) vec = vec && (((uintptr_t)P.mul_src) % (4 * sizeof(TMSK)) == 0);
    if (P.aux) vec = vec && (((uintptr_t)P.aux) % (4 * eo) == 0);
    ProbeScope probe(PROBE_TAPGEMM, 2.0 * (double)rows * P.N * P.ntaps * P.K,
                     (double)rows * (P.K * ea + P.N * eo), st);
    if (P.N <= 8) {
        dim3 grid((unsigned)((rows + 511) / 512), (unsigned)((P.N + 7) / 8));
        if (vec) tapgemm_kernel<TA, TO, TMSK, 512, 8, 4, 4, 1, true><<<grid, 256, 0, st>>>(P);
        else tapgemm_kernel<TA, TO, TMSK, 512, 8, 4, 4, 1, false><<<grid, 256, 0, st>>>(P);
    } else if (((rows + 127) / 128) * ((P.N + 63) / 64) < 2LL * num_sms()) {
        // small problem (the MLPs of E_num / G / the critic and classifier heads): 64-row tiles for twice the CTAs
        dim3 grid((unsigned)((rows + 63) / 64), (unsigned)((P.N + 63) / 64));
        if (vec) tapgemm_kernel<TA, TO, TMSK, 64, 64, 4, 4, 2, true><<<grid, 256, 0, st>>>(P);
        else tapgemm_kernel<TA, TO, TMSK, 64, 64, 4, 4, 2, false><<<grid, 256, 0, st>>>(P);
    } else {
        dim3 grid((unsigned)((rows + 127) / 128), (unsigned)((P.N + 63) / 64));
        if (vec) tapgemm_kernel<TA, TO, TMSK, 128, 64, 8, 4, 2, true><<<grid, 256, 0, st>>>(P);
        else tapgemm_kernel<TA, TO, TMSK, 128, 64, 8, 4, 2, false><<<grid, 256, 0, st>>>(P);
    }
    MG_LAUNCH_OK();
    return MG_OK;
}

// ------------------------------------------------------------------------------------------------
// wgrad
// ------------------------------------------------------------------------------------------------
template <typename TG, typename TA, bool VEC>
__global__ void __launch_bounds__(256) wgrad_kernel(const WgradArgs P, int rows_per_split) {
    constexpr int RK = 16, BNW = 64, BKW = 64;
    __shared__ __align__(16) float Gs[2][RK][BNW + 4];
    __shared__ __align__(16) float As[2][RK][BKW + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.x * BNW, kk0 = blockIdx.y * BKW;
    const int Ktot = P.ntaps * P.K;
    const long long r_begin = P.row_begin + (long long)blockIdx.z * rows_per_split;
    long long r_end = r_begin + rows_per_split;
    if (r_end > P.row_end) r_end = P.row_end;
    const TG* __restrict__ Gbase = static_cast<const TG*>(P.G);
    const TA* __restrict__ Abase = static_cast<const TA*>(P.A);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    const int lp = tid >> 4;          // position within chunk
    const int lc = (tid & 15) * 4;    // 4 consecutive columns
    // tap / k of this thread's 4 A columns (fixed for the whole kernel)
    int a_t[4], a_k[4];
    bool a_ok[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int kk = kk0 + lc + e;
        a_ok[e] = kk < Ktot;
        a_t[e] = a_ok[e] ? kk / P.K : 0;
        a_k[e] = a_ok[e] ? kk - a_t[e] * P.K : 0;
    }

    auto load_tiles = [&](long long rbase, int buf) {
        const long long r = rbase + lp;
        float g[4] = {0.f, 0.f, 0.f, 0.f}, a[4] = {0.f, 0.f, 0.f, 0.f};
        if (r < r_end) {
            const long long b = r / P.Mper;
            const int m = (int)(r - b * P.Mper);
            const long long go = b * P.g_bstride + (long long)m * P.g_mstride + P.g_off + n0 + lc;
            const int am = m * P.a_mstride;
            if (VEC) {
                if (n0 + lc < P.N) ld4(Gbase + go, g);
                if (a_ok[0]) {
                    const int idx = am + P.a_toff[a_t[0]] + a_k[0];
                    if (idx >= 0 && idx < P.a_valid) ld4(Abase + b * P.a_bstride + idx, a);
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (n0 + lc + e < P.N) g[e] = ld_as_float(Gbase + go + e);
                    if (a_ok[e]) {
                        const int idx = am + P.a_toff[a_t[e]] + a_k[e];
                        if (idx >= 0 && idx < P.a_valid) a[e] = ld_as_float(Abase + b * P.a_bstride + idx);
                    }
                }
            }
        }
        *reinterpret_cast<float4*>(&Gs[buf][lp][lc]) = make_float4(g[0], g[1], g[2], g[3]);
        *reinterpret_cast<float4*>(&As[buf][lp][lc]) = make_float4(a[0], a[1], a[2], a[3]);
    };

    const long long nrows = r_end - r_begin;
    if (nrows > 0) {
        const int nchunks = (int)((nrows + RK - 1) / RK);
        load_tiles(r_begin, 0);
        __syncthreads();
        for (int c = 0; c < nchunks; ++c) {
            const int buf = c & 1;
            if (c + 1 < nchunks) load_tiles(r_begin + (long long)(c + 1) * RK, buf ^ 1);
#pragma unroll
            for (int p = 0; p < RK; ++p) {
                const float4 g = *reinterpret_cast<const float4*>(&Gs[buf][p][ty * 4]);
                const float4 a = *reinterpret_cast<const float4*>(&As[buf][p][tx * 4]);
                const float gv[4] = {g.x, g.y, g.z, g.w}, av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gv[i], av[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= P.N) continue;
        const long long nphys = perm_index(n, P.n_perm_q, P.n_perm_p);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kk = kk0 + tx * 4 + j;
            if (kk >= Ktot) continue;
            const int t = kk / P.K, k = kk - t * P.K;
            atomicAdd(P.dW + P.w_toff[t] + nphys * P.w_nstride + (long long)k * P.w_kstride, P.alpha * acc[i][j]);
        }
    }
}

template <typename TG, typename TA>
int launch_wgrad(const WgradArgs& P, cudaStream_t st) {
    const long long nrows = (long long)P.row_end - P.row_begin;
    if (nrows <= 0 || P.N == 0) return MG_OK;
    {
        int r = tc::try_tc_wgrad<TG, TA>(P, st);
        if (r == 0) r = thin::try_thin_wgrad<TG, TA>(P, st);
        if (r != 0) return r < 0 ? r : MG_OK;
    }
    const int Ktot = P.ntaps * P.K;
    const int tn = (P.N + 63) / 64, tk = (Ktot + 63) / 64;
    // enough splits for ~4 CTAs per SM, at least 64 positions per split
    long long want = ((long long)num_sms() * 4 + tn * tk - 1) / (tn * tk);
    long long maxs = (nrows + 63) / 64;
    long long splits = want < 1 ? 1 : (want > maxs ? maxs : want);
    if (splits > 65535) splits = 65535;
    int rps = (int)((nrows + splits - 1) / splits);
    rps = (rps + 15) / 16 * 16;
    splits = (nrows + rps - 1) / rps;
    bool vec = (P.K % 4 == 0) && (P.N % 4 == 0) && (P.a_mstride % 4 == 0) && (P.a_bstride % 4 == 0) &&
               (P.a_valid % 4 == 0) && (P.g_mstride % 4 == 0) && (P.g_bstride % 4 == 0) && (P.g_off % 4 == 0) &&
               (((uintptr_t)P.A) % (4 * sizeof(TA)) == 0) && (((uintptr_t)P.G) % (4 * sizeof(TG)) == 0);
    for (int t = 0; t < P.ntaps; ++t) vec = vec && (P.a_toff[t] % 4 == 0);
    ProbeScope probe(PROBE_WGRAD, 2.0 * (double)nrows * P.N * Ktot, (double)nrows * (P.N * sizeof(TG) + P.K * sizeof(TA)), st);
    dim3 grid(tn, tk, (unsigned)splits);
    if (vec) wgrad_kernel<TG, TA, true><<<grid, 256, 0, st>>>(P, rps);
    else wgrad_kernel<TG, TA, false><<<grid, 256, 0, st>>>(P, rps);
    MG_LAUNCH_OK();
    return MG_OK;
}

}  // namespace mg
