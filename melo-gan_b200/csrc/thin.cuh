// thin.cuh -- bandwidth-class kernels for the layers with a 4-channel side (hard part 5).
//
// The note tensor has 4 channels, so three layer shapes are not MMA tiles but skinny products against a
// 64-channel activation; as generic 128x64 GEMM tiles they waste 3-16x of their FMAs.  Dedicated forms:
//   thin_k_fwd    out[row, 0:64] = epi( W[64][Ktot<=32] . window(row) )        conv.0 of D and of ED (K = 5 taps x 4 = 20)
//                 the taps of a C_in=4 conv are ONE contiguous window of the channels-last input
//   thin_n_fwd    out[row, 0:4]  = sum_t W_t[4][64] . A[row + shift_t, 0:64]   deconv.6 of G, conv.0 dgrads
//   thin_wgrad    dW[64][Ktot<=32] += sum_rows G[row, 0:64] (x) window(row)    conv.0 / deconv.6 weight gradients
// All read each activation element once, coalesced (8 threads x 16 B per 64-channel bf16 row).
#pragma once
#include <type_traits>

#include "gemm_simt.cuh"

namespace mg {
namespace thin {

constexpr int kMaxWin = 32;

__device__ __forceinline__ float apply_act_mask(float x, int act, int mul_mode, float ms, float& gd) {
    gd = 0.0f;
    if (act == ACT_RELU) x = fmaxf(x, 0.0f);
    else if (act == ACT_LRELU) x = x > 0.0f ? x : 0.2f * x;
    else if (act == ACT_GELU) { gd = gelu_grad_f(x); x = gelu_f(x); }
    if (mul_mode == MUL_LRELU_SIGN) x *= (ms > 0.0f ? 1.0f : 0.2f);
    else if (mul_mode == MUL_RELU_SIGN) x *= (ms > 0.0f ? 1.0f : 0.0f);
    else if (mul_mode == MUL_VALUE) x *= ms;
    return x;
}

// ---- thin_k_fwd: thread = (row group of 4 rows, 8 of the 64 output channels); weights from smem as float4 ------
template <typename TO, typename TMSK>
__global__ void __launch_bounds__(256) thin_k_fwd_kernel(const TapGemmArgs P, int win_off, int Ktot) {
    __shared__ __align__(16) float Ws[kMaxWin][64];     // [kk][n]
    __shared__ float bs[64], cs[64];
    for (int i = threadIdx.x; i < Ktot * 64; i += 256) {
        const int n = i % 64, kk = i / 64, t = kk / P.K, k = kk - t * P.K;
        Ws[kk][n] = __ldg(P.W + P.w_toff[t] + (long long)n * P.w_nstride + (long long)k * P.w_kstride);
    }
    if (threadIdx.x < 64) {
        bs[threadIdx.x] = P.bias ? __ldg(P.bias + threadIdx.x) : 0.0f;
        cs[threadIdx.x] = (P.col_scale ? __ldg(P.col_scale + threadIdx.x) : 1.0f) * P.alpha;
    }
    __syncthreads();
    const long long rows = (long long)P.B * P.Mper;
    const int sub = threadIdx.x & 7, rg = threadIdx.x >> 3;
    const float* __restrict__ A = static_cast<const float*>(P.A);
    TO* __restrict__ Ob = static_cast<TO*>(P.Out);
    const TMSK* __restrict__ Mb = static_cast<const TMSK*>(P.mul_src);
    TO* __restrict__ Xb = static_cast<TO*>(P.aux);
    for (long long r0 = ((long long)blockIdx.x * 32 + rg) * 4; r0 < rows; r0 += (long long)gridDim.x * 128) {
        const float* ap[4];
        int base[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long r = r0 + i < rows ? r0 + i : rows - 1;
            const long long b = r / P.Mper;
            base[i] = (int)(r - b * P.Mper) * P.a_mstride + win_off;
            ap[i] = A + b * P.a_bstride;
        }
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
        for (int kk = 0; kk < Ktot; kk += 4) {
            float av[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = base[i] + kk;
                float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx >= 0 && idx < P.a_valid) q = __ldg(reinterpret_cast<const float4*>(ap[i] + idx));
                av[i][0] = q.x; av[i][1] = q.y; av[i][2] = q.z; av[i][3] = q.w;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float4 w0 = *reinterpret_cast<const float4*>(&Ws[kk + e][sub * 8]);
                const float4 w1 = *reinterpret_cast<const float4*>(&Ws[kk + e][sub * 8 + 4]);
                const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i][e], w[j], acc[i][j]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long r = r0 + i;
            if (r >= rows) break;
            const long long b = r / P.Mper;
            const int m = (int)(r - b * P.Mper);
            const long long o = b * P.o_bstride + (long long)m * P.o_mstride + P.o_off + sub * 8;
            float ms[8], x[8], gd[8];
            if (P.mul_mode != MUL_NONE) {
                float t4[4];
                ld4(Mb + o, t4); ms[0] = t4[0]; ms[1] = t4[1]; ms[2] = t4[2]; ms[3] = t4[3];
                ld4(Mb + o + 4, t4); ms[4] = t4[0]; ms[5] = t4[1]; ms[6] = t4[2]; ms[7] = t4[3];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
                x[j] = apply_act_mask(fmaf(acc[i][j], cs[sub * 8 + j], bs[sub * 8 + j]), P.act, P.mul_mode, ms[j], gd[j]);
            const float x0[4] = {x[0], x[1], x[2], x[3]}, x1[4] = {x[4], x[5], x[6], x[7]};
            st4(Ob + o, x0); st4(Ob + o + 4, x1);
            if (P.aux) {
                const float g0[4] = {gd[0], gd[1], gd[2], gd[3]}, g1[4] = {gd[4], gd[5], gd[6], gd[7]};
                st4(Xb + o, g0); st4(Xb + o + 4, g1);
            }
        }
    }
}

// ---- thin_n_fwd: thread = (group of 4 output rows, 8 of the 64 input channels); shuffle-reduced over the 8 ----
template <typename TA>
__global__ void __launch_bounds__(256) thin_n_fwd_kernel(const TapGemmArgs P) {
    __shared__ __align__(16) float Ws[kMaxTaps][4][64];     // [tap][n][k]
    for (int i = threadIdx.x; i < P.ntaps * 4 * 64; i += 256) {
        const int k = i % 64, n = (i / 64) % 4, t = i / 256;
        Ws[t][n][k] = __ldg(P.W + P.w_toff[t] + (long long)n * P.w_nstride + (long long)k * P.w_kstride);
    }
    __syncthreads();
    const long long rows = (long long)P.B * P.Mper;
    const int sub = threadIdx.x & 7, rg = threadIdx.x >> 3;
    const TA* __restrict__ A = static_cast<const TA*>(P.A);
    float* __restrict__ Ob = static_cast<float*>(P.Out);
    const long long rows_pad = (rows + 127) / 128 * 128;    // whole warps stay in the loop for the shuffles
    for (long long r0 = ((long long)blockIdx.x * 32 + rg) * 4; r0 < rows_pad; r0 += (long long)gridDim.x * 128) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int n = 0; n < 4; ++n) acc[i][n] = 0.0f;
        const TA* ap[4];
        int mi[4];
        bool ok[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ok[i] = r0 + i < rows;
            const long long r = ok[i] ? r0 + i : 0;
            const long long b = r / P.Mper;
            mi[i] = (int)(r - b * P.Mper);
            ap[i] = A + b * P.a_bstride;
        }
        for (int t = 0; t < P.ntaps; ++t) {
            float w[4][8];
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const float4 w0 = *reinterpret_cast<const float4*>(&Ws[t][n][sub * 8]);
                const float4 w1 = *reinterpret_cast<const float4*>(&Ws[t][n][sub * 8 + 4]);
                w[n][0] = w0.x; w[n][1] = w0.y; w[n][2] = w0.z; w[n][3] = w0.w;
                w[n][4] = w1.x; w[n][5] = w1.y; w[n][6] = w1.z; w[n][7] = w1.w;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = mi[i] * P.a_mstride + P.a_toff[t] + sub * 8;
                if (!ok[i] || idx < 0 || idx >= P.a_valid) continue;
                float av[8];
                { float t4[4]; ld4(ap[i] + idx, t4); av[0] = t4[0]; av[1] = t4[1]; av[2] = t4[2]; av[3] = t4[3];
                  ld4(ap[i] + idx + 4, t4); av[4] = t4[0]; av[5] = t4[1]; av[6] = t4[2]; av[7] = t4[3]; }
#pragma unroll
                for (int n = 0; n < 4; ++n)
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[i][n] = fmaf(av[e], w[n][e], acc[i][n]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                acc[i][n] += __shfl_xor_sync(0xffffffffu, acc[i][n], 1);
                acc[i][n] += __shfl_xor_sync(0xffffffffu, acc[i][n], 2);
                acc[i][n] += __shfl_xor_sync(0xffffffffu, acc[i][n], 4);
            }
        // lane `sub` (0..3) of the group writes row i = sub
        if (sub < 4 && ok[sub & 3]) {
            float v4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (sub == i) { v4[0] = acc[i][0]; v4[1] = acc[i][1]; v4[2] = acc[i][2]; v4[3] = acc[i][3]; }
            const long long r = r0 + sub;
            const long long b = r / P.Mper;
            const int m = (int)(r - b * P.Mper);
            const long long o = b * P.o_bstride + (long long)m * P.o_mstride + P.o_off;
            float4 v;
            v.x = v4[0] * P.alpha + (P.bias ? __ldg(P.bias + 0) : 0.f);
            v.y = v4[1] * P.alpha + (P.bias ? __ldg(P.bias + 1) : 0.f);
            v.z = v4[2] * P.alpha + (P.bias ? __ldg(P.bias + 2) : 0.f);
            v.w = v4[3] * P.alpha + (P.bias ? __ldg(P.bias + 3) : 0.f);
            if (P.accumulate) {
                const float4 old = *reinterpret_cast<const float4*>(Ob + o);
                v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
            }
            *reinterpret_cast<float4*>(Ob + o) = v;
        }
    }
}

// ---- thin_wgrad: dW[n][kk] += sum_rows G[row, n] * window(row)[kk] ------------------------------------------
// One warp per row: lane = (n8 = lane & 7 -> channels 8*n8..8*n8+7, kq = lane >> 3 -> window columns 8*kq..8*kq+7),
// 64 accumulators per thread, 64 FMAs per (one 16-byte G load + two 16-byte window loads).  The 8 warps of a CTA
// take rows round-robin; their partial [64][32] blocks are combined with shared-memory atomics, then one global
// atomicAdd per weight and CTA.
template <typename TG>
__global__ void __launch_bounds__(256) thin_wgrad_kernel(const WgradArgs P, int win_off, int Ktot, int rows_per_cta) {
    __shared__ float red[64][kMaxWin + 1];
    for (int i = threadIdx.x; i < 64 * (kMaxWin + 1); i += 256) (&red[0][0])[i] = 0.0f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n8 = lane & 7, kq = lane >> 3;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    const TG* __restrict__ G = static_cast<const TG*>(P.G);
    const float* __restrict__ A = static_cast<const float*>(P.A);
    const long long r0 = P.row_begin + (long long)blockIdx.x * rows_per_cta;
    long long r1 = r0 + rows_per_cta;
    if (r1 > P.row_end) r1 = P.row_end;
#pragma unroll 2
    for (long long r = r0 + warp; r < r1; r += 8) {
        const long long b = r / P.Mper;
        const int m = (int)(r - b * P.Mper);
        float g[8], a[8];
        { float t4[4];
          const TG* gp = G + b * P.g_bstride + (long long)m * P.g_mstride + P.g_off + n8 * 8;
          ld4(gp, t4); g[0] = t4[0]; g[1] = t4[1]; g[2] = t4[2]; g[3] = t4[3];
          ld4(gp + 4, t4); g[4] = t4[0]; g[5] = t4[1]; g[6] = t4[2]; g[7] = t4[3]; }
        const int base = m * P.a_mstride + win_off + kq * 8;
        const float* ap = A + b * P.a_bstride;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int idx = base + h * 4;
            float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kq * 8 + h * 4 < Ktot && idx >= 0 && idx < P.a_valid) q4 = __ldg(reinterpret_cast<const float4*>(ap + idx));
            a[h * 4] = q4.x; a[h * 4 + 1] = q4.y; a[h * 4 + 2] = q4.z; a[h * 4 + 3] = q4.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(g[i], a[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (kq * 8 + j < Ktot) atomicAdd(&red[n8 * 8 + i][kq * 8 + j], acc[i][j]);
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * Ktot; i += 256) {
        const int n = i / Ktot, kk = i - n * Ktot;
        const int t = kk / P.K, k = kk - t * P.K;
        atomicAdd(P.dW + P.w_toff[t] + (long long)n * P.w_nstride + (long long)k * P.w_kstride, P.alpha * red[n][kk]);
    }
}

// ---- dispatch helpers: return 1 if launched, 0 if the pattern does not match ------------------------
inline bool contiguous_window(const int* a_toff, int ntaps, int K) {
    for (int t = 1; t < ntaps; ++t)
        if (a_toff[t] != a_toff[0] + t * K) return false;
    return true;
}

template <typename TA, typename TO, typename TMSK>
int try_thin_tapgemm(const TapGemmArgs& P, cudaStream_t st) {
    const long long rows = (long long)P.B * P.Mper;
    const int Ktot = P.ntaps * P.K;
    if (P.row_scale || P.n_perm_q || P.k_perm_q) return 0;
    // thin K: float window input, 64 outputs
    if (std::is_same<TA, float>::value && P.N == 64 && Ktot <= kMaxWin && Ktot % 4 == 0 && P.K % 4 == 0 &&
        contiguous_window(P.a_toff, P.ntaps, P.K) && P.a_mstride % 4 == 0 && P.a_toff[0] % 4 == 0 && P.a_valid % 4 == 0 &&
        P.a_bstride % 4 == 0 && !P.accumulate && P.o_off % 8 == 0 && P.o_mstride % 8 == 0 && P.o_bstride % 8 == 0 &&
        ((uintptr_t)P.A) % 16 == 0 && ((uintptr_t)P.Out) % 16 == 0) {
        long long blocks = (rows + 127) / 128;
        const long long cap = (long long)num_sms() * 8;
        if (blocks > cap) blocks = cap;
        ProbeScope probe(PROBE_TAPGEMM, 2.0 * (double)rows * P.N * Ktot, (double)rows * (Ktot * 4.0 / 2.5 + P.N * sizeof(TO)), st);
        thin_k_fwd_kernel<TO, TMSK><<<(int)blocks, 256, 0, st>>>(P, P.a_toff[0], Ktot);
        MG_LAUNCH_OK();
        return 1;
    }
    // thin N: 64-channel input rows, 4 float outputs
    if (std::is_same<TO, float>::value && P.N == 4 && P.K == 64 && P.ntaps <= kMaxTaps && P.act == ACT_NONE &&
        P.mul_mode == MUL_NONE && !P.aux && !P.col_scale && P.a_mstride % 8 == 0 && P.a_valid % 8 == 0 &&
        P.a_bstride % 8 == 0 && P.o_off % 4 == 0 && P.o_mstride % 4 == 0 && P.o_bstride % 4 == 0 &&
        ((uintptr_t)P.A) % 16 == 0 && ((uintptr_t)P.Out) % 16 == 0) {
        for (int t = 0; t < P.ntaps; ++t)
            if (P.a_toff[t] % 8) return 0;
        long long blocks = (rows + 127) / 128;
        const long long cap = (long long)num_sms() * 8;
        if (blocks > cap) blocks = cap;
        ProbeScope probe(PROBE_TAPGEMM, 2.0 * (double)rows * P.N * Ktot, (double)rows * (P.K * sizeof(TA) + 16.0), st);
        thin_n_fwd_kernel<TA><<<(int)blocks, 256, 0, st>>>(P);
        MG_LAUNCH_OK();
        return 1;
    }
    return 0;
}

template <typename TG, typename TA>
int try_thin_wgrad(const WgradArgs& P, cudaStream_t st) {
    const int Ktot = P.ntaps * P.K;
    if (!std::is_same<TA, float>::value || P.N != 64 || Ktot > kMaxWin || Ktot % 4 || P.n_perm_q) return 0;
    if (!contiguous_window(P.a_toff, P.ntaps, P.K)) return 0;
    if (P.a_mstride % 4 || P.a_toff[0] % 4 || P.a_valid % 4 || P.a_bstride % 4 || ((uintptr_t)P.A) % 16) return 0;
    if (P.g_mstride % 8 || P.g_bstride % 8 || P.g_off % 8 || ((uintptr_t)P.G) % 16) return 0;
    const long long nrows = (long long)P.row_end - P.row_begin;
    if (nrows <= 0) return 1;
    long long ctas = (long long)num_sms() * 4;
    long long rpc = (nrows + ctas - 1) / ctas;
    rpc = (rpc + 7) / 8 * 8;
    if (rpc < 64) rpc = 64;
    ctas = (nrows + rpc - 1) / rpc;
    ProbeScope probe(PROBE_WGRAD, 2.0 * (double)nrows * P.N * Ktot, (double)nrows * (P.N * sizeof(TG) + Ktot * 4.0 / 2.5), st);
    thin_wgrad_kernel<TG><<<(int)ctas, 256, 0, st>>>(P, P.a_toff[0], Ktot, (int)rpc);
    MG_LAUNCH_OK();
    return 1;
}

}  // namespace thin
}  // namespace mg
