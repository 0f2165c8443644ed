// elem.cuh -- HBM-bound elementwise / reduction kernels of the GAN step (sm_100a).
// Activation tensors are channels-last [rows, C] with C contiguous; T is float (fp32 parity mode) or
// __nv_bfloat16 (bf16 mode).  Reductions over rows write per-CTA partials and a second tiny kernel
// finishes them in a fixed order (deterministic, no float atomics) unless noted.
#pragma once
#include "gemm_simt.cuh"

namespace mg {

// ---------------------------------------------------------------------------------------------
// column reductions over rows:  out[k][c] = sum_{r in [r0,r1)} f_k(row r, col c)
// one CTA = 32 (cols) x 8 (row lanes) threads striding over a row chunk; partials [nchunk][K][C]
// ---------------------------------------------------------------------------------------------
enum ColOp {
    COL_SUM = 0,        // f0 = x
    COL_SUM_SQ = 1,     // f0 = x, f1 = x*x                      (BatchNorm batch statistics)
    COL_BN_BWD = 2,     // f0 = dy, f1 = dy * xhat                (xhat from x, mean, invstd)
    COL_WSUM = 3        // f0 = w[r] * x                          (row-weighted sum)
};

struct ColReduceArgs {
    const void* x; int ldx;            // x[r*ldx + c]
    const void* y; int ldy;            // second operand (COL_BN_BWD: dy; x is the pre-BN activation)
    const float* mean; const float* invstd;
    const float* roww;                 // COL_WSUM: weight of row r is roww[r / roww_div]
    int roww_div;
    long long r0, r1; int C;
    float* partial;                    // [nchunk][nout][C]
    int rows_per_chunk;
};

template <typename T, typename TY, int OP>
static __global__ void __launch_bounds__(256) colreduce_kernel(const ColReduceArgs P) {
    constexpr int NOUT = (OP == COL_SUM_SQ || OP == COL_BN_BWD) ? 2 : 1;
    __shared__ float red[NOUT][8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    const long long rb = P.r0 + (long long)blockIdx.y * P.rows_per_chunk;
    long long re = rb + P.rows_per_chunk;
    if (re > P.r1) re = P.r1;
    float a0 = 0.f, a1 = 0.f;
    if (c < P.C) {
        const T* x = static_cast<const T*>(P.x);
        const TY* y = static_cast<const TY*>(P.y);
        float mu = 0.f, is = 0.f;
        if (OP == COL_BN_BWD) { mu = P.mean[c]; is = P.invstd[c]; }
        for (long long r = rb + ry; r < re; r += 8) {
            const float xv = ld_as_float(x + r * P.ldx + c);
            if (OP == COL_SUM) a0 += xv;
            else if (OP == COL_SUM_SQ) { a0 += xv; a1 = fmaf(xv, xv, a1); }
            else if (OP == COL_BN_BWD) {
                const float dy = ld_as_float(y + r * P.ldy + c);
                a0 += dy; a1 = fmaf(dy, (xv - mu) * is, a1);
            } else a0 = fmaf(P.roww[r / P.roww_div], xv, a0);
        }
    }
    red[0][ry][cx] = a0;
    if (NOUT == 2) red[1][ry][cx] = a1;
    __syncthreads();
    if (ry == 0 && c < P.C) {
#pragma unroll
        for (int k = 0; k < NOUT; ++k) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) s += red[k][j][cx];
            P.partial[((long long)blockIdx.y * NOUT + k) * P.C + c] = s;
        }
    }
}

// vectorised form: every thread owns 4 consecutive columns (one 8/16-byte load per row), 8 row lanes per CTA
template <typename T, typename TY, int OP>
static __global__ void __launch_bounds__(256) colreduce_vec4_kernel(const ColReduceArgs P) {
    constexpr int NOUT = (OP == COL_SUM_SQ || OP == COL_BN_BWD) ? 2 : 1;
    __shared__ float red[NOUT][8][32][4 + 1];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + cx) * 4;
    const long long rb = P.r0 + (long long)blockIdx.y * P.rows_per_chunk;
    long long re = rb + P.rows_per_chunk;
    if (re > P.r1) re = P.r1;
    float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < P.C) {
        const T* x = static_cast<const T*>(P.x);
        const TY* y = static_cast<const TY*>(P.y);
        float mu[4] = {0.f, 0.f, 0.f, 0.f}, is[4] = {0.f, 0.f, 0.f, 0.f};
        if (OP == COL_BN_BWD) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { mu[e] = P.mean[c + e]; is[e] = P.invstd[c + e]; }
        }
#pragma unroll 4
        for (long long r = rb + ry; r < re; r += 8) {
            float xv[4];
            ld4(x + r * P.ldx + c, xv);
            if (OP == COL_SUM) {
#pragma unroll
                for (int e = 0; e < 4; ++e) a0[e] += xv[e];
            } else if (OP == COL_SUM_SQ) {
#pragma unroll
                for (int e = 0; e < 4; ++e) { a0[e] += xv[e]; a1[e] = fmaf(xv[e], xv[e], a1[e]); }
            } else if (OP == COL_BN_BWD) {
                float dy[4];
                ld4(y + r * P.ldy + c, dy);
#pragma unroll
                for (int e = 0; e < 4; ++e) { a0[e] += dy[e]; a1[e] = fmaf(dy[e], (xv[e] - mu[e]) * is[e], a1[e]); }
            } else {
                const float w = P.roww[r / P.roww_div];
#pragma unroll
                for (int e = 0; e < 4; ++e) a0[e] = fmaf(w, xv[e], a0[e]);
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        red[0][ry][cx][e] = a0[e];
        if (NOUT == 2) red[1][ry][cx][e] = a1[e];
    }
    __syncthreads();
    // 32 column groups x 4 columns x NOUT outputs = up to 256 sums: one per thread
    {
        const int k = threadIdx.x / 128, rem = threadIdx.x % 128, gx = rem / 4, e = rem % 4;
        const int cc = (blockIdx.x * 32 + gx) * 4 + e;
        if (k < NOUT && cc < P.C) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) s += red[k][j][gx][e];
            P.partial[((long long)blockIdx.y * NOUT + k) * P.C + cc] = s;
        }
    }
}

// Narrow matrices (C <= 1024, C/4 a power of two): the 256 threads of a CTA tile (row lane) x (4-column group) so that
// every lane loads, whatever C is; one CTA per row chunk, rows unrolled 8 deep to keep ~16 KB per CTA in flight.
template <typename T, typename TY, int OP>
static __global__ void __launch_bounds__(256) colreduce_flat_kernel(const ColReduceArgs P) {
    constexpr int NOUT = (OP == COL_SUM_SQ || OP == COL_BN_BWD) ? 2 : 1;
    __shared__ float red[NOUT][4][256];
    const int cgs = P.C >> 2, lanes = 256 / cgs;
    const int cg = threadIdx.x % cgs, rl = threadIdx.x / cgs, c = cg * 4;
    const long long rb = P.r0 + (long long)blockIdx.y * P.rows_per_chunk;
    long long re = rb + P.rows_per_chunk;
    if (re > P.r1) re = P.r1;
    float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
    const T* x = static_cast<const T*>(P.x);
    const TY* y = static_cast<const TY*>(P.y);
    float mu[4] = {0.f, 0.f, 0.f, 0.f}, is[4] = {0.f, 0.f, 0.f, 0.f};
    if (OP == COL_BN_BWD) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { mu[e] = P.mean[c + e]; is[e] = P.invstd[c + e]; }
    }
#pragma unroll 8
    for (long long r = rb + rl; r < re; r += lanes) {
        float xv[4];
        ld4(x + r * P.ldx + c, xv);
        if (OP == COL_SUM) {
#pragma unroll
            for (int e = 0; e < 4; ++e) a0[e] += xv[e];
        } else if (OP == COL_SUM_SQ) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { a0[e] += xv[e]; a1[e] = fmaf(xv[e], xv[e], a1[e]); }
        } else if (OP == COL_BN_BWD) {
            float dy[4];
            ld4(y + r * P.ldy + c, dy);
#pragma unroll
            for (int e = 0; e < 4; ++e) { a0[e] += dy[e]; a1[e] = fmaf(dy[e], (xv[e] - mu[e]) * is[e], a1[e]); }
        } else {
            const float w = P.roww[r / P.roww_div];
#pragma unroll
            for (int e = 0; e < 4; ++e) a0[e] = fmaf(w, xv[e], a0[e]);
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        red[0][e][threadIdx.x] = a0[e];
        if (NOUT == 2) red[1][e][threadIdx.x] = a1[e];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NOUT * P.C; i += 256) {
        const int k = i / P.C, cc = i - k * P.C, g = cc >> 2, e = cc & 3;
        float sum = 0.f;
        for (int j = 0; j < lanes; ++j) sum += red[k][e][j * cgs + g];
        P.partial[((long long)blockIdx.y * NOUT + k) * P.C + cc] = sum;
    }
}

// out[k*out_kstride + perm(c)] (+)= alpha * sum_chunks partial[chunk][k][c]   (float64 accumulation)
// 8 outputs per CTA, 32 chunk lanes each, combined in a fixed order (deterministic)
static __global__ void __launch_bounds__(256) colreduce_finish_kernel(const float* __restrict__ partial, int nchunk, int nout,
                                                               int C, float* out, int out_kstride, int perm_q,
                                                               int perm_p, float alpha, int accumulate) {
    __shared__ double red[32][9];
    const int cx = threadIdx.x & 7, jl = threadIdx.x >> 3;
    const int i = blockIdx.x * 8 + cx;
    double s = 0.0;
    if (i < nout * C)
        for (int j = jl; j < nchunk; j += 32) s += (double)partial[(long long)j * nout * C + i];
    red[jl][cx] = s;
    __syncthreads();
    if (jl == 0 && i < nout * C) {
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < 32; ++j) t += red[j][cx];
        const int k = i / C, c = i - k * C;
        float* o = out + (long long)k * out_kstride + perm_index(c, perm_q, perm_p);
        const float v = (float)(t * (double)alpha);
        *o = accumulate ? *o + v : v;
    }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm1d (training) helpers for G's two BN layers        reference models.py:57,60
// ---------------------------------------------------------------------------------------------
// stats[0][c] = sum x, stats[1][c] = sum x^2 over R rows -> mean, invstd (biased var), running update
static __global__ void bn_finalize_kernel(const float* __restrict__ stats, int C, long long R, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* running_mean,
                                   float* running_var, int update_running) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double n = (double)R;
    const double mu = (double)stats[c] / n;
    double var = (double)stats[C + c] / n - mu * mu;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)mu;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (update_running) {
        const double unbiased = R > 1 ? var * n / (n - 1.0) : var;
        running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mu);
        running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
    }
}

// y = relu((x - mean) * invstd * gamma + beta)      (eval mode: mean/invstd come from running stats)
template <typename TX, typename T>
static __global__ void __launch_bounds__(256) bn_relu_apply_kernel(const TX* __restrict__ x, T* __restrict__ y, long long n4,
                                                            int C, const float* __restrict__ mean,
                                                            const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (sizeof(TX) == 4 && sizeof(T) == 2 && (n4 & 1) == 0 && (stride * 8) % C == 0 && C % 8 == 0) {
        // float32 in, bf16 out: two 16-byte loads and ONE 16-byte store per step (8 channels, constant per thread)
        const int c = (int)((i0 * 8) % C);
        float sc[8], sh[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            sc[e] = invstd[c + e] * gamma[c + e];
            sh[e] = fmaf(-mean[c + e], sc[e], beta[c + e]);
        }
        const long long n8 = n4 >> 1;
        const float4* x4 = reinterpret_cast<const float4*>(x);
        uint4* y4 = reinterpret_cast<uint4*>(y);
#pragma unroll 4
        for (long long i = i0; i < n8; i += stride) {
            const float4 a = __ldcs(x4 + 2 * i), b = __ldcs(x4 + 2 * i + 1);
            const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(fmaxf(fmaf(v[2 * e], sc[2 * e], sh[2 * e]), 0.0f),
                                                               fmaxf(fmaf(v[2 * e + 1], sc[2 * e + 1], sh[2 * e + 1]), 0.0f));
                w[e] = *reinterpret_cast<const uint32_t*>(&h);
            }
            y4[i] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        return;
    }
    if ((stride * 4) % C == 0) {      // this thread's 4 channels never change: fold the statistics once
        const int c = (int)((i0 * 4) % C);
        float sc[4], sh[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            sc[e] = invstd[c + e] * gamma[c + e];
            sh[e] = fmaf(-mean[c + e], sc[e], beta[c + e]);
        }
#pragma unroll 4
        for (long long i = i0; i < n4; i += stride) {
            float v[4];
            ld4(x + i * 4, v);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = fmaxf(fmaf(v[e], sc[e], sh[e]), 0.0f);
            st4(y + i * 4, v);
        }
        return;
    }
    for (long long i = i0; i < n4; i += stride) {
        const int c = (int)((i * 4) % C);
        float v[4];
        ld4(x + i * 4, v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float sc = invstd[c + e] * gamma[c + e];
            v[e] = fmaxf(fmaf(v[e] - mean[c + e], sc, beta[c + e]), 0.0f);
        }
        st4(y + i * 4, v);
    }
}

// y = gelu(v), g = gelu'(v) with v = (x - mean) * invstd * gamma + beta      (train-mode ConvBlock1D, ed_model.py:35-42)
template <typename TX, typename T>
static __global__ void __launch_bounds__(256) bn_gelu_apply_kernel(const TX* __restrict__ x, T* __restrict__ y, T* __restrict__ g,
                                                            long long n4, int C, const float* __restrict__ mean,
                                                            const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int c = (int)((i * 4) % C);
        float v[4], o[4], d[4];
        ld4(x + i * 4, v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float u = fmaf(v[e] - mean[c + e], invstd[c + e] * gamma[c + e], beta[c + e]);
            o[e] = gelu_f(u);
            d[e] = gelu_grad_f(u);
        }
        st4(y + i * 4, o);
        st4(g + i * 4, d);
    }
}

// dx = gamma*invstd*(dy - s1/R - xhat*s2/R), sums = [s1 | s2] (dy already carries the ReLU mask)
template <typename TX, typename T, typename TD>
static __global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const TX* __restrict__ x, const TD* __restrict__ dy,
                                                           T* __restrict__ dx, long long n4, int C, float invR,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ invstd,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ sums) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    if ((stride * 4) % C == 0) {      // this thread's 4 channels never change: dx = A*dy + Bx*x + Cc with folded coefficients
        const int c = (int)((((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4) % C);
        float A[4], Bx[4], Cc[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float is = invstd[c + e], gi = gamma[c + e] * is, s2 = sums[C + c + e] * invR;
            A[e] = gi;
            Bx[e] = -gi * is * s2;
            Cc[e] = -gi * sums[c + e] * invR - Bx[e] * mean[c + e];
        }
#pragma unroll 4
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            float xv[4], dv[4], o[4];
            ld4(x + i * 4, xv);
            ld4(dy + i * 4, dv);
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = fmaf(A[e], dv[e], fmaf(Bx[e], xv[e], Cc[e]));
            st4(dx + i * 4, o);
        }
        return;
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int c = (int)((i * 4) % C);
        float xv[4], dv[4], o[4];
        ld4(x + i * 4, xv);
        ld4(dy + i * 4, dv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float xh = (xv[e] - mean[c + e]) * invstd[c + e];
            o[e] = gamma[c + e] * invstd[c + e] * (dv[e] - sums[c + e] * invR - xh * sums[C + c + e] * invR);
        }
        st4(dx + i * 4, o);
    }
}

// ---------------------------------------------------------------------------------------------
// pooling and broadcast
// ---------------------------------------------------------------------------------------------
// out[s, c] = scale * sum_l x[s, l, c]        (AdaptiveAvgPool1d(1): scale = 1/L)
// thread = (sample, 4 channels, row lane); C/4 * RL threads per sample, RL row lanes combined through shared memory
template <typename T, typename TO>
static __global__ void __launch_bounds__(256) pool_rows_kernel(const T* __restrict__ x, TO* __restrict__ out, int S, int L,
                                                        int C, float scale) {
    __shared__ float red[256][4 + 1];
    const int c4n = C / 4;                       // channel groups (C % 4 == 0, c4n <= 256)
    const int RL = 256 / c4n;                    // row lanes
    const int cg = threadIdx.x % c4n, rl = threadIdx.x / c4n;
    const int s = blockIdx.x;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    if (rl < RL) {
        const T* p = x + ((long long)s * L) * C + cg * 4;
#pragma unroll 4
        for (int l = rl; l < L; l += RL) {
            float v[4];
            ld4(p + (long long)l * C, v);
            a[0] += v[0]; a[1] += v[1]; a[2] += v[2]; a[3] += v[3];
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) red[threadIdx.x][e] = a[e];
    __syncthreads();
    if (rl == 0) {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < RL; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) t[e] += red[j * c4n + cg][e];
#pragma unroll
        for (int e = 0; e < 4; ++e) st_from_float(out + (long long)s * C + cg * 4 + e, t[e] * scale);
    }
}

// out[s, l, c] = src[s, c] * scale * colscale[c] * f'(ref[s, l, c])
//   mode MUL_LRELU_SIGN / MUL_RELU_SIGN: derivative from the sign of the saved activation; MUL_VALUE: ref holds f'
// grid = (chunks, samples): all index arithmetic is 32-bit and per-sample
template <typename TS, typename T, typename TOUT = T>
static __global__ void __launch_bounds__(256) bcast_rows_mul_kernel(const TS* __restrict__ src, const T* __restrict__ ref,
                                                             TOUT* __restrict__ out, int L, int C, float scale,
                                                             const float* __restrict__ colscale, int mode) {
    const int s = blockIdx.y;
    const int n4 = L * C / 4;
    const T* r0 = ref + (long long)s * L * C;
    TOUT* o0 = out + (long long)s * L * C;
    const TS* sp = src + (long long)s * C;
    if (sizeof(T) == 2 && sizeof(TOUT) == 2 && (n4 & 1) == 0 && C % 8 == 0 && (gridDim.x * 2048) % C == 0) {
        // bf16 in, bf16 out: one 16-byte load and one 16-byte store per step (8 channels, constant per thread)
        const int c = ((blockIdx.x * 256 + threadIdx.x) * 8) % C;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = ld_as_float(sp + c + e) * (scale * (colscale ? colscale[c + e] : 1.0f));
        const uint4* r4 = reinterpret_cast<const uint4*>(r0);
        uint4* o4 = reinterpret_cast<uint4*>(o0);
        const int n8 = n4 >> 1;
#pragma unroll 4
        for (int i = blockIdx.x * 256 + threadIdx.x; i < n8; i += gridDim.x * 256) {
            const uint4 q = __ldcs(r4 + i);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float r_lo = __uint_as_float(w[e] << 16), r_hi = __uint_as_float(w[e] & 0xFFFF0000u);
                const float d_lo = (mode == MUL_LRELU_SIGN) ? (r_lo > 0.f ? 1.f : 0.2f) : (mode == MUL_RELU_SIGN) ? (r_lo > 0.f ? 1.f : 0.f) : r_lo;
                const float d_hi = (mode == MUL_LRELU_SIGN) ? (r_hi > 0.f ? 1.f : 0.2f) : (mode == MUL_RELU_SIGN) ? (r_hi > 0.f ? 1.f : 0.f) : r_hi;
                const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * e] * d_lo, f[2 * e + 1] * d_hi);
                o[e] = *reinterpret_cast<const uint32_t*>(&h);
            }
            o4[i] = make_uint4(o[0], o[1], o[2], o[3]);
        }
        return;
    }
    if ((gridDim.x * 1024) % C == 0) {      // this thread's 4 channels never change: fold the row factor once
        const int c = ((blockIdx.x * 256 + threadIdx.x) * 4) % C;
        float f[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) f[e] = ld_as_float(sp + c + e) * (scale * (colscale ? colscale[c + e] : 1.0f));
#pragma unroll 4
        for (int i = blockIdx.x * 256 + threadIdx.x; i < n4; i += gridDim.x * 256) {
            float r[4], o[4];
            ld4(r0 + i * 4, r);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float d = (mode == MUL_LRELU_SIGN) ? (r[e] > 0.f ? 1.f : 0.2f)
                                : (mode == MUL_RELU_SIGN) ? (r[e] > 0.f ? 1.f : 0.f) : r[e];
                o[e] = f[e] * d;
            }
            st4(o0 + i * 4, o);
        }
        return;
    }
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n4; i += gridDim.x * 256) {
        const int c = (i * 4) % C;
        float r[4], o[4];
        ld4(r0 + i * 4, r);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float d = (mode == MUL_LRELU_SIGN) ? (r[e] > 0.f ? 1.f : 0.2f)
                            : (mode == MUL_RELU_SIGN) ? (r[e] > 0.f ? 1.f : 0.f) : r[e];
            const float cs = colscale ? colscale[c + e] : 1.0f;
            o[e] = ld_as_float(sp + c + e) * (scale * cs) * d;
        }
        st4(o0 + i * 4, o);
    }
}

// The same for bf16 tensors as a persistent kernel: a CTA walks whole samples (grid-strided), 16-byte loads and stores, the
// thread's 8 channels and their row factors fixed per sample (2048 % C == 0).  Optionally adds the column sums of the STORED
// values over the samples s < colsum_samples to colsum_out (the bias gradient of the layer whose output gradient this is):
// per-thread running sums, one shared-memory pass and one atomicAdd per channel and CTA at the end.
// (stand-alone at the critic's shape, scripts/probes/elem_probe.cu: 317 us for the (chunks, samples) grid above, 295 us here)
template <typename TS>
static __global__ void __launch_bounds__(256) bcast_rows_mul_bf16_kernel(const TS* __restrict__ src, const __nv_bfloat16* __restrict__ ref,
                                                                   __nv_bfloat16* __restrict__ out, int S, int L, int C, float scale,
                                                                   const float* __restrict__ colscale, int mode,
                                                                   float* __restrict__ colsum_out, int colsum_samples) {
    __shared__ float red[256][8 + 1];
    const int per8 = L * C / 8;
    const int c = (threadIdx.x * 8) % C;
    float cs[8], acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { cs[e] = scale * (colscale ? colscale[c + e] : 1.0f); acc[e] = 0.0f; }
    for (int s = blockIdx.x; s < S; s += gridDim.x) {
        const uint4* r4 = reinterpret_cast<const uint4*>(ref + (long long)s * L * C);
        uint4* o4 = reinterpret_cast<uint4*>(out + (long long)s * L * C);
        const bool summed = colsum_out != nullptr && s < colsum_samples;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = ld_as_float(src + (long long)s * C + c + e) * cs[e];
#pragma unroll 4
        for (int i = threadIdx.x; i < per8; i += 256) {
            const uint4 q = __ldcs(r4 + i);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xFFFF0000u);
                const float dl = (mode == MUL_LRELU_SIGN) ? (lo > 0.f ? 1.f : 0.2f) : (mode == MUL_RELU_SIGN) ? (lo > 0.f ? 1.f : 0.f) : lo;
                const float dh = (mode == MUL_LRELU_SIGN) ? (hi > 0.f ? 1.f : 0.2f) : (mode == MUL_RELU_SIGN) ? (hi > 0.f ? 1.f : 0.f) : hi;
                const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * e] * dl, f[2 * e + 1] * dh);
                o[e] = *reinterpret_cast<const uint32_t*>(&h);
                if (summed) { acc[2 * e] += __low2float(h); acc[2 * e + 1] += __high2float(h); }
            }
            o4[i] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
    if (colsum_out) {                             // threads t and t + C/8 (mod 256) hold the same channels
#pragma unroll
        for (int e = 0; e < 8; ++e) red[threadIdx.x][e] = acc[e];
        __syncthreads();
        const int groups = C / 8;                 // distinct channel groups among the 256 threads
        for (int i = threadIdx.x; i < C; i += 256) {
            const int g = i >> 3, e = i & 7;
            float t = 0.0f;
            for (int j = g; j < 256; j += groups) t += red[j][e];
            atomicAdd(colsum_out + i, t);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// FeatureEncoder pieces (tiny tensors, float only)          reference feature_encoder.py:17-41
// ---------------------------------------------------------------------------------------------
// LayerNorm over D (<= 32) features per row; writes y and xhat
static __global__ void layernorm_small_kernel(const float* __restrict__ x, int B, int D, const float* __restrict__ w,
                                       const float* __restrict__ b, float eps, float* __restrict__ y,
                                       float* __restrict__ xhat) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B) return;
    float mu = 0.f;
    for (int j = 0; j < D; ++j) mu += x[r * D + j];
    mu /= (float)D;
    float var = 0.f;
    for (int j = 0; j < D; ++j) { const float d = x[r * D + j] - mu; var = fmaf(d, d, var); }
    var /= (float)D;
    const float is = rsqrtf(var + eps);
    for (int j = 0; j < D; ++j) {
        const float xh = (x[r * D + j] - mu) * is;
        if (xhat) xhat[r * D + j] = xh;
        y[r * D + j] = fmaf(xh, w[j], b[j]);
    }
}

// h = gelu(z) * mask * scale   (mask == nullptr: eval mode, no dropout)
static __global__ void gelu_dropout_fwd_kernel(const float* __restrict__ z, const float* __restrict__ mask, float scale,
                                        float* __restrict__ h, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g = gelu_f(z[i]);
    h[i] = mask ? g * (mask[i] * scale) : g;
}
// dz = dh * mask * scale * gelu'(z)
static __global__ void gelu_dropout_bwd_kernel(const float* dh, const float* __restrict__ z,
                                        const float* __restrict__ mask, float scale, float* dz, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float m = mask ? mask[i] * scale : 1.0f;
    dz[i] = dh[i] * m * gelu_grad_f(z[i]);
}

// xcat[b] = [a[b, 0:Da] | c[b, 0:Dc]]
static __global__ void concat2_kernel(const float* __restrict__ a, int Da, const float* __restrict__ c, int Dc,
                               float* __restrict__ out, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int D = Da + Dc;
    if (i >= B * D) return;
    const int b = i / D, j = i - b * D;
    out[i] = j < Da ? a[b * Da + j] : c[b * Dc + (j - Da)];
}

// out[b, off + j] = src[b, j]  for j < D   (block of a wider row)
static __global__ void copy_cols_kernel(const float* __restrict__ src, int D, float* __restrict__ out, int ldo, int off, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * D) return;
    const int b = i / D, j = i - b * D;
    out[(long long)b * ldo + off + j] = src[i];
}

// out[b, j] (+)= src[b, off + j]  for j < D   (slice of a wider row)
static __global__ void slice_add_kernel(const float* __restrict__ src, int lds, int off, float* __restrict__ out, int D,
                                 int B, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * D) return;
    const int b = i / D, j = i - b * D;
    const float v = src[(long long)b * lds + off + j];
    out[i] = accumulate ? out[i] + v : v;
}

// ---------------------------------------------------------------------------------------------
// critic head, gradient penalty, losses
// ---------------------------------------------------------------------------------------------
// score[r] = hf[r,:] . w[0:F] + emb[r % B,:] . w[F:F+E] + bias           reference models.py:164-169
static __global__ void critic_score_kernel(const float* __restrict__ hf, const float* __restrict__ emb,
                                    const float* __restrict__ w, const float* __restrict__ bias, int R, int B, int F,
                                    int E, float* __restrict__ score) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    float a = 0.f;
    for (int j = lane; j < F; j += 32) a = fmaf(hf[(long long)r * F + j], w[j], a);
    if (emb)
        for (int j = lane; j < E; j += 32) a = fmaf(emb[(long long)(r % B) * E + j], w[F + j], a);
#pragma unroll
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) score[r] = a + bias[0];
}

// dzf[r, j] = seed[r] * w[j] * lrelu'(hf[r, j])
static __global__ void critic_head_bwd_kernel(const float* __restrict__ hf, const float* __restrict__ w,
                                       const float* __restrict__ seed, int R, int F, float* __restrict__ dzf) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)R * F) return;
    const int r = (int)(i / F), j = (int)(i - (long long)r * F);
    dzf[i] = seed[r] * w[j] * (hf[i] > 0.f ? 1.f : 0.2f);
}

// X3 = [real | fake | alpha*real + (1-alpha)*fake]  (each segment B*per floats)   utils.py:76-78
static __global__ void __launch_bounds__(256) assemble_critic_input_kernel(const float4* __restrict__ real,
                                                                    const float4* __restrict__ fake,
                                                                    const float* __restrict__ alpha,
                                                                    float4* __restrict__ x3, int B, int per4) {
    const long long n = (long long)B * per4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float4 r = __ldg(real + i), f = __ldg(fake + i);
        const float a = alpha[i / per4];
        const float na = 1.0f - a;
        float4 m;
        m.x = __fadd_rn(__fmul_rn(a, r.x), __fmul_rn(na, f.x));
        m.y = __fadd_rn(__fmul_rn(a, r.y), __fmul_rn(na, f.y));
        m.z = __fadd_rn(__fmul_rn(a, r.z), __fmul_rn(na, f.z));
        m.w = __fadd_rn(__fmul_rn(a, r.w), __fmul_rn(na, f.w));
        x3[i] = r; x3[n + i] = f; x3[2 * n + i] = m;
    }
}

// per sample: n = ||g||_2 over `per` floats; gp_b = (n-1)^2; u = lambda * 2(n-1)/(B n) * g written to `u`
static __global__ void __launch_bounds__(256) gp_norm_kernel(const float* __restrict__ g, float* __restrict__ u, int per,
                                                      float lambda_over_B, float* __restrict__ gp_per_sample,
                                                      float* __restrict__ norm_per_sample) {
    __shared__ float red[8];
    __shared__ float s_coef;
    const int b = blockIdx.x;
    const float* gb = g + (long long)b * per;
    float a = 0.f;
    for (int i = threadIdx.x; i < per; i += blockDim.x) a = fmaf(gb[i], gb[i], a);
#pragma unroll
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
        const float nrm = sqrtf(s);
        gp_per_sample[b] = (nrm - 1.0f) * (nrm - 1.0f);
        if (norm_per_sample) norm_per_sample[b] = nrm;
        s_coef = nrm > 0.f ? lambda_over_B * 2.0f * (nrm - 1.0f) / nrm : 0.f;
    }
    __syncthreads();
    const float coef = s_coef;
    float* ub = u + (long long)b * per;
    for (int i = threadIdx.x; i < per; i += blockDim.x) ub[i] = coef * gb[i];
}

// metrics of the critic step, one thread block: means over B of score segments and gp
// out: [0]=loss_d [1]=gp [2]=mean d_real [3]=mean d_fake
static __global__ void critic_loss_kernel(const float* __restrict__ score, const float* __restrict__ gp_ps, int B,
                                   float lambda_gp, float* __restrict__ out) {
    __shared__ double red[3][32];
    double a = 0, f = 0, g = 0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) { a += score[i]; f += score[B + i]; g += gp_ps[i]; }
    for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o); f += __shfl_xor_sync(0xffffffffu, f, o);
        g += __shfl_xor_sync(0xffffffffu, g, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = f; red[2][threadIdx.x >> 5] = g; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = f = g = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += red[0][i]; f += red[1][i]; g += red[2][i]; }
        const float dr = (float)(a / B), df = (float)(f / B), gp = (float)(g / B);
        out[0] = df - dr + lambda_gp * gp; out[1] = gp; out[2] = dr; out[3] = df;
    }
}

// generator losses: adv = -mean(score); CE over 4-way logits; dlogits = weight * (softmax - onehot)/B
// out: [0]=loss_g_adv [1]=loss_g_emo
static __global__ void generator_loss_kernel(const float* __restrict__ score, const float* __restrict__ logits,
                                      const long long* __restrict__ labels, int B, int NC, float emo_weight,
                                      float* __restrict__ dlogits, float* __restrict__ out) {
    __shared__ double red[2][32];
    double a = 0, ce = 0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        a += score[i];
        const float* l = logits + (long long)i * NC;
        float mx = l[0];
        for (int k = 1; k < NC; ++k) mx = fmaxf(mx, l[k]);
        float se = 0.f;
        for (int k = 0; k < NC; ++k) se += expf(l[k] - mx);
        const float lse = mx + logf(se);
        // nn.CrossEntropyLoss raises on a target outside [0, NC) (emotion_to_index returns -1 for unknown moods); a kernel
        // cannot raise, so such a row poisons the loss and its gradient with NaN instead of reading outside the logits row
        const long long y64 = labels[i];
        const bool y_ok = y64 >= 0 && y64 < NC;
        const int y = y_ok ? (int)y64 : 0;
        ce += y_ok ? (double)(lse - l[y]) : (double)__int_as_float(0x7fc00000);
        for (int k = 0; k < NC; ++k)
            dlogits[(long long)i * NC + k] =
                y_ok ? emo_weight * (expf(l[k] - lse) - (k == y ? 1.f : 0.f)) / (float)B : __int_as_float(0x7fc00000);
    }
    for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); ce += __shfl_xor_sync(0xffffffffu, ce, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = ce; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = ce = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += red[0][i]; ce += red[1][i]; }
        out[0] = (float)(-a / B); out[1] = (float)(ce / B);
    }
}

// a[i] *= b[i]
static __global__ void mul_inplace_kernel(float* a, const float* b, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] *= b[i];
}
// a[i] += b[i]
static __global__ void axpy_kernel(float* a, const float* b, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += b[i];
}
// a[i] += x[i]; b[i] += y[i]   (accumulate the two BatchNorm affine gradients)
static __global__ void add2_kernel(float* a, const float* x, float* b, const float* y, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { a[i] += x[i]; b[i] += y[i]; }
}
// hf[i] = lrelu'(hf[i]) * q[i]
static __global__ void lrelu_mask_mul_inplace_kernel(float* hf, const float* q, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) hf[i] = (hf[i] > 0.f ? 1.f : 0.2f) * q[i];
}
// eval-mode BatchNorm: mean/invstd from the running statistics
static __global__ void bn_eval_stats_kernel(const float* rm, const float* rv, float eps, int C, float* mean, float* invstd) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) { mean[c] = rm[c]; invstd[c] = 1.0f / sqrtf(rv[c] + eps); }
}
static __global__ void const_fill_kernel(float* p, int n, float v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// demb[b, j] (+)= seed[b] * w[F + j]     (critic conditioning term, models.py:165-168)
static __global__ void critic_demb_kernel(const float* seed, const float* w, int B, int F, int E, float* demb, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * E) return;
    const float v = seed[i / E] * w[F + i % E];
    demb[i] = accumulate ? demb[i] + v : v;
}
static __global__ void critic_seed_kernel(float* seed, int B, float w_real, float w_fake) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * B) return;
    seed[i] = i < B ? w_real / (float)B : (i < 2 * B ? w_fake / (float)B : 1.0f);
}

// classifier training loss: mean CE over B, accuracy, and dlogits = (softmax - onehot)/B    (train_ed.py:66-80)
// out: [0] = loss, [1] = accuracy
static __global__ void ce_loss_acc_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int B, int NC,
                                   float* __restrict__ dlogits, float* __restrict__ out) {
    __shared__ double red[2][32];
    double ce = 0, hit = 0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const float* l = logits + (long long)i * NC;
        float mx = l[0];
        int am = 0;
        for (int k = 1; k < NC; ++k) if (l[k] > mx) { mx = l[k]; am = k; }
        float se = 0.f;
        for (int k = 0; k < NC; ++k) se += expf(l[k] - mx);
        const float lse = mx + logf(se);
        const long long y64 = labels[i];                 // out-of-range targets: NaN loss and gradient (see above)
        const bool y_ok = y64 >= 0 && y64 < NC;
        const int y = y_ok ? (int)y64 : 0;
        ce += y_ok ? (double)(lse - l[y]) : (double)__int_as_float(0x7fc00000);
        hit += (y_ok && am == y) ? 1.0 : 0.0;
        if (dlogits)
            for (int k = 0; k < NC; ++k)
                dlogits[(long long)i * NC + k] =
                    y_ok ? (expf(l[k] - lse) - (k == y ? 1.f : 0.f)) / (float)B : __int_as_float(0x7fc00000);
    }
    for (int o = 16; o; o >>= 1) { ce += __shfl_xor_sync(0xffffffffu, ce, o); hit += __shfl_xor_sync(0xffffffffu, hit, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ce; red[1][threadIdx.x >> 5] = hit; }
    __syncthreads();
    if (threadIdx.x == 0) {
        ce = hit = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { ce += red[0][i]; hit += red[1][i]; }
        out[0] = (float)(ce / B); out[1] = (float)(hit / B);
    }
}

static __global__ void fill_kernel(float* p, long long n, float v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// eval-mode BatchNorm folded into a per-channel scale/shift around a conv bias:
//   y = (conv + bias - rm) * g/sqrt(rv+eps) + b  =  conv*scale + shift
static __global__ void bn_fold_kernel(const float* __restrict__ conv_bias, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ rm,
                               const float* __restrict__ rv, float eps, int C, float* __restrict__ scale,
                               float* __restrict__ shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float s = gamma[c] / sqrtf(rv[c] + eps);
    scale[c] = s;
    shift[c] = fmaf((conv_bias ? conv_bias[c] : 0.f) - rm[c], s, beta[c]);
}

}  // namespace mg
