// banded.cuh -- the 4-channel layers on the tensor cores (bf16 mode).
//
// conv.0 of the critic and of the emotion discriminator (C_in = 4), G's deconv.6 (C_out = 4) and their
// dgrads / wgrads are not MMA tiles as written: K = 20 or N = 4.  Grouping consecutive positions turns each into a
// dense-enough GEMM with a BANDED weight matrix that the tcgen05 kernels of gemm_tc.cuh run unchanged:
//
//   thin-K  (4 -> 64, k5):  J = 4 output positions x 64 channels = 256 outputs per group read ONE 64-element window
//           of the zero-padded bf16 note tensor; consecutive windows overlap (TMA row stride 16 or 32 elements
//           < box width 64).  Wb[j*64+co][(s*j+t)*4+ci] = W[co][ci][t]  (s = conv stride), GEMM K = 64, N = 256.
//   thin-N  (64 -> 4):      16 output positions x 4 channels = 64 outputs per group of 8 (or 16) input rows;
//           input row (group*P + rho) is tap rho of a row-mod-P plane view; 10 (or 20) taps of K = 64, N = 64.
//   wgrad   of thin-K:      dWb[256][64] = sum_groups G[group][256]^T window[group][64] on the MN-major tcgen05 wgrad,
//           then a tiny fold of the 4 diagonal bands back into dW[co][ci][t].
//
// 2.5-4x of the MACs are structural zeros, which is free next to the 3-16x the CUDA-core forms wasted and the fp32
// FMA rate they ran at.
#pragma once
#include "gemm_tc.cuh"

namespace mg {
namespace banded {

constexpr int kPadFront = 32;                 // zero elements in front of every padded note row
constexpr int kPadTotal = 96;                 // LP = T*4 + 96: 32 zeros in front, 64 behind (a window never leaves its row)

// ---- padded bf16 copy of a (R, T, 4) float note tensor: dst[r][32 + i] = bf16(src[r][i]) ----
static __global__ void __launch_bounds__(256) pad_convert_kernel(const float4* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                                 long long R, int per4, int LP) {
    const long long n = R * per4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long r = i / per4;
        const int j = (int)(i - r * per4);
        const float4 v = __ldg(src + i);
        const float a[4] = {v.x, v.y, v.z, v.w};
        st4(dst + r * LP + kPadFront + j * 4, a);
    }
}

// ---- banded weight packs (bf16, [tap][N][64]) ----
// thin-K: Wb[j*64+co][(s*j+t)*4+ci] = W[w_co*co + w_ci*ci + w_t*t]
static __global__ void pack_band_k_kernel(const float* __restrict__ W, int w_co, int w_ci, int w_t, int s, __nv_bfloat16* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // over 256 x 64
    if (i >= 256 * 64) return;
    const int n = i / 64, w = i % 64, j = n / 64, co = n % 64, ro = w / 4, ci = w % 4, t = ro - s * j;
    float v = 0.0f;
    if (t >= 0 && t < 5) v = __ldg(W + (long long)w_co * co + (long long)w_ci * ci + (long long)w_t * t);
    out[i] = __float2bfloat16_rn(v);
}
// thin-N, x2 up-sampling form (deconv.6 forward, dgrad of a stride-2 conv): 10 taps, tau = mu + d(t) + 1
//   out[2*mu+ph][n4] += W(t, n4, k) * in[mu + d][k],  d = 1 - t/2, ph = t % 2
static __global__ void pack_band_up_kernel(const float* __restrict__ W, int w_n, int w_k, int w_t, __nv_bfloat16* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // over 10 x 64 x 64
    if (i >= 10 * 64 * 64) return;
    const int k = i % 64, n = (i / 64) % 64, tau = i / 4096;
    const int jo = n / 4, n4 = n % 4, mu = jo / 2, ph = jo % 2, d = tau - mu - 1;
    const int t = ph == 0 ? 2 - 2 * d : 3 - 2 * d;
    float v = 0.0f;
    if (d >= -1 && d <= 1 && t >= 0 && t < 5 && (t % 2) == ph)
        v = __ldg(W + (long long)w_n * n4 + (long long)w_k * k + (long long)w_t * t);
    out[i] = __float2bfloat16_rn(v);
}
// thin-N, stride-1 k5 p2 dgrad form (ED conv.0 dgrad): 20 taps, tau = jo + 4 - t
static __global__ void pack_band_s1_kernel(const float* __restrict__ W, int w_n, int w_k, int w_t, __nv_bfloat16* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // over 20 x 64 x 64
    if (i >= 20 * 64 * 64) return;
    const int k = i % 64, n = (i / 64) % 64, tau = i / 4096;
    const int jo = n / 4, n4 = n % 4, t = jo + 4 - tau;
    float v = 0.0f;
    if (t >= 0 && t < 5) v = __ldg(W + (long long)w_n * n4 + (long long)w_k * k + (long long)w_t * t);
    out[i] = __float2bfloat16_rn(v);
}
// out[i] = src[i % period]   (bias / folded-BN vectors replicated over the J positions of a group)
static __global__ void replicate_kernel(const float* __restrict__ src, int period, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src ? src[i % period] : 0.0f;
}
// dW[w_co*co + w_ci*ci + w_t*t] += sum_j dWb[j*64+co][(s*j+t)*4+ci]
static __global__ void fold_band_k_kernel(const float* __restrict__ dWb, int s, float* dW, int w_co, int w_ci, int w_t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // over 64 x 4 x 5
    if (i >= 64 * 4 * 5) return;
    const int t = i % 5, ci = (i / 5) % 4, co = i / 20;
    float a = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) a += dWb[(j * 64 + co) * 64 + (s * j + t) * 4 + ci];
    dW[(long long)w_co * co + (long long)w_ci * ci + (long long)w_t * t] += a;
}

struct Scratch2 {            // per-context scratch of the banded forms (allocated from the arena)
    __nv_bfloat16* wb;       // packed banded weights, up to 20*64*64
    float* vec_a;            // replicated bias   [256]
    float* vec_b;            // replicated scale  [256]
    float* dwb;              // banded wgrad accumulator [256][64]
};

inline int pad_convert(const float* src, __nv_bfloat16* dst, long long R, int T, cudaStream_t st) {
    const int per4 = T, LP = T * 4 + kPadTotal;
    long long blocks = (R * per4 + 255) / 256;
    if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
    pad_convert_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(src), dst, R, per4, LP);
    MG_LAUNCH_OK();
    return MG_OK;
}

// thin-K forward: xp (R, LP) padded bf16 notes -> Out (R, T/s, 64) of type TO.   W is [co][ci][t] by strides.
template <typename TO, typename TMSK>
int conv_k_fwd(const Scratch2& sc, const __nv_bfloat16* xp, int R, int T, int s, const float* W, int w_co, int w_ci,
               int w_t, const float* bias, const float* col_scale, int act, void* aux, const void* mul_src, int mul_mode,
               TO* Out, cudaStream_t st) {
    const int LP = T * 4 + kPadTotal;
    const int Lout = T / s, groups = Lout / 4;                 // groups of 4 output positions per sample
    pack_band_k_kernel<<<64, 256, 0, st>>>(W, w_co, w_ci, w_t, s, sc.wb);
    MG_LAUNCH_OK();
    replicate_kernel<<<1, 256, 0, st>>>(bias, 64, sc.vec_a, 256);
    MG_LAUNCH_OK();
    if (col_scale) { replicate_kernel<<<1, 256, 0, st>>>(col_scale, 64, sc.vec_b, 256); MG_LAUNCH_OK(); }
    tc::TcTapArgs a{};
    a.ntaps = 1; a.kblocks = 1; a.a_p[0] = 0; a.a_dm[0] = 0; a.b_row[0] = 0;
    a.mpt = groups >= 128 ? 128 : groups; a.bpt = 128 / a.mpt;
    a.Mper = groups; a.B = R; a.N = 256;
    a.Out = Out; a.o_bstride = (long long)Lout * 64; a.o_mstride = 256; a.o_off = 0;
    a.bias = bias ? sc.vec_a : nullptr; a.col_scale = col_scale ? sc.vec_b : nullptr; a.act = act;
    a.mul_src = mul_src; a.mul_mode = mul_mode; a.aux = aux; a.alpha = 1.0f; a.accumulate = 0;
    CUtensorMap am, bm;
    // window of group g starts at padded element 32 + 4*s*4*g - 8 = 24 + 16*s*g
    MG_TRY(tc::make_view_map(&am, xp + (kPadFront - 8), 64, 1, 16 * s, groups, 16 * s, R, LP, a.mpt, a.bpt));
    MG_TRY(tc::make_weight_map(&bm, sc.wb, 64, 256, 128));
    return tc::run_tc_tap<TO, TMSK>(am, bm, a, 128, 64, st);
}

// thin-N, x2 up-sampling: in (R, Lin, 64) bf16 -> out (R, 2*Lin, 4) float32.   W(t, n4, k) by strides.
inline int up_n_fwd(const Scratch2& sc, const __nv_bfloat16* in, int R, int Lin, const float* W, int w_n, int w_k, int w_t,
                    const float* bias, float* out, int accumulate, cudaStream_t st) {
    const int groups = Lin / 8;
    pack_band_up_kernel<<<160, 256, 0, st>>>(W, w_n, w_k, w_t, sc.wb);
    MG_LAUNCH_OK();
    replicate_kernel<<<1, 64, 0, st>>>(bias, 4, sc.vec_a, 64);
    MG_LAUNCH_OK();
    tc::TcTapArgs a{};
    a.ntaps = 10; a.kblocks = 1;
    for (int tau = 0; tau < 10; ++tau) {
        const int rho = tau - 1;                                  // input row = 8*group + rho
        a.a_p[tau] = ((rho % 8) + 8) % 8;
        a.a_dm[tau] = (rho - a.a_p[tau]) / 8;
        a.b_row[tau] = tau * 64;
    }
    a.mpt = groups >= 128 ? 128 : groups; a.bpt = 128 / a.mpt;
    a.Mper = groups; a.B = R; a.N = 64;
    a.Out = out; a.o_bstride = (long long)2 * Lin * 4; a.o_mstride = 64; a.o_off = 0;
    a.bias = bias ? sc.vec_a : nullptr; a.act = ACT_NONE; a.mul_mode = MUL_NONE; a.alpha = 1.0f; a.accumulate = accumulate;
    CUtensorMap am, bm;
    MG_TRY(tc::make_view_map(&am, in, 64, 8, 64, groups, 8 * 64, R, (long long)Lin * 64, a.mpt, a.bpt));
    MG_TRY(tc::make_weight_map(&bm, sc.wb, 64, 10 * 64, 64));
    return tc::run_tc_tap<float, float>(am, bm, a, 64, 64, st);
}

// thin-N, stride-1 k5 p2 dgrad: dZ (R, L, 64) bf16 -> dX (R, L, 4) float32 (+=).   W(t, n4 = ci, k = co) by strides.
inline int s1_n_dgrad(const Scratch2& sc, const __nv_bfloat16* dZ, int R, int L, const float* W, int w_n, int w_k, int w_t,
                      float* out, int accumulate, cudaStream_t st) {
    const int groups = L / 16;
    pack_band_s1_kernel<<<320, 256, 0, st>>>(W, w_n, w_k, w_t, sc.wb);
    MG_LAUNCH_OK();
    tc::TcTapArgs a{};
    a.ntaps = 20; a.kblocks = 1;
    for (int tau = 0; tau < 20; ++tau) {
        const int rho = tau - 2;                                  // input row = 16*group + rho
        a.a_p[tau] = ((rho % 16) + 16) % 16;
        a.a_dm[tau] = (rho - a.a_p[tau]) / 16;
        a.b_row[tau] = tau * 64;
    }
    a.mpt = groups >= 128 ? 128 : groups; a.bpt = 128 / a.mpt;
    a.Mper = groups; a.B = R; a.N = 64;
    a.Out = out; a.o_bstride = (long long)L * 4; a.o_mstride = 64; a.o_off = 0;
    a.act = ACT_NONE; a.mul_mode = MUL_NONE; a.alpha = 1.0f; a.accumulate = accumulate;
    CUtensorMap am, bm;
    MG_TRY(tc::make_view_map(&am, dZ, 64, 16, 64, groups, 16 * 64, R, (long long)L * 64, a.mpt, a.bpt));
    MG_TRY(tc::make_weight_map(&bm, sc.wb, 64, 20 * 64, 64));
    return tc::run_tc_tap<float, float>(am, bm, a, 64, 64, st);
}

// wgrad of a thin-K conv (stride s): dW[co][ci][t] += sum_{r,l} G[r, l, co] * x[r, s*l + t - 2, ci]
//   G (R, Lg, 64) bf16 with Lg = T/s positions, xp (R, LP) padded bf16 notes; rows [0, R) all reduced
inline int conv_k_wgrad(const Scratch2& sc, const __nv_bfloat16* G, const __nv_bfloat16* xp, int R, int T, int s, float* dW,
                        int w_co, int w_ci, int w_t, cudaStream_t st) {
    const int LP = T * 4 + kPadTotal, Lg = T / s, groups = Lg / 4;
    MG_CUDA_OK(cudaMemsetAsync(sc.dwb, 0, sizeof(float) * 256 * 64, st));
    tc::TcWgradArgs a{};
    a.ntaps = 1; a.K = 64; a.N = 256; a.a_p[0] = 0; a.a_dm[0] = 0;
    a.rpt = groups >= 64 ? 64 : groups; a.spt = 64 / a.rpt; a.Mper = groups;
    a.row_begin = 0; a.row_end = (long long)R * groups;
    a.dW = sc.dwb; a.w_toff[0] = 0; a.w_nstride = 64; a.w_kstride = 1; a.alpha = 1.0f;
    const long long nrows = a.row_end;
    const int tiles = 2;
    long long splits = ((long long)num_sms() * 2 + tiles - 1) / tiles;
    const long long maxs = (nrows + 255) / 256;
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
    long long rps = (nrows + splits - 1) / splits;
    rps = (rps + 63) / 64 * 64;
    splits = (nrows + rps - 1) / rps;
    a.rows_per_split = (int)rps;
    CUtensorMap gm, am;
    MG_TRY(tc::make_act_map(&gm, G, 256, groups, R, 1, a.rpt, a.spt));
    MG_TRY(tc::make_view_map(&am, xp + (kPadFront - 8), 64, 1, 16 * s, groups, 16 * s, R, LP, a.rpt, a.spt));
    {
        ProbeScope probe(PROBE_TC_WGRAD, 2.0 * (double)nrows * 256 * 64, (double)nrows * (256 + 64) * 2.0, st);
        MG_TRY(tc::launch_tc_wgrad<64>(gm, am, a, (int)splits, st));
    }
    fold_band_k_kernel<<<5, 256, 0, st>>>(sc.dwb, s, dW, w_co, w_ci, w_t);
    MG_LAUNCH_OK();
    return MG_OK;
}

}  // namespace banded
}  // namespace mg
