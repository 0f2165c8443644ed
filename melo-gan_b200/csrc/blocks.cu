// blocks.cu -- the MLP-type inner blocks of the reference as stand-alone operators: NoiseToLatent
// (src/gan/models.py:20-29), MLPClassifier (src/emotion_discriminator/ed_model.py:74-101; with input_mode 'latent' it IS
// the emotion discriminator, ed_model.py:128-136,156-160) and the Linear stack of GeneratorDecoder.pre (models.py:46-51).
// On the hot path these run fused inside mg_generator_* / mg_emotion_*; called on their own (module.forward of the inner
// block) they are chains of the four operators below, float32, on the same contraction kernels as the fp32 parity mode.
#include "gan_ctx.cuh"

using namespace mg;

namespace {

__global__ void act_dropout_fwd_kernel(const float* __restrict__ z, const float* __restrict__ mask, float scale, int act,
                                       float* __restrict__ h, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = z[i];
    if (act == ACT_RELU) v = fmaxf(v, 0.0f);
    else if (act == ACT_LRELU) v = v > 0.0f ? v : 0.2f * v;
    else if (act == ACT_GELU) v = gelu_f(v);
    h[i] = mask ? v * (mask[i] * scale) : v;
}

__global__ void act_dropout_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ z, const float* __restrict__ mask,
                                       float scale, int act, float* __restrict__ dz, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = z[i];
    float d = 1.0f;
    if (act == ACT_RELU) d = v > 0.0f ? 1.0f : 0.0f;
    else if (act == ACT_LRELU) d = v > 0.0f ? 1.0f : 0.2f;
    else if (act == ACT_GELU) d = gelu_grad_f(v);
    dz[i] = dh[i] * d * (mask ? mask[i] * scale : 1.0f);
}

// db[c] += sum_r dz[r, c]: 32 columns per block, 8 row lanes, fixed summation order per column
__global__ void bias_grad_kernel(const float* __restrict__ dz, float* __restrict__ db, int rows, int N) {
    __shared__ float part[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float s = 0.0f;
    if (c < N)
        for (int r = threadIdx.y; r < rows; r += 8) s += dz[(long long)r * N + c];
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < N) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x];
        db[c] += t;
    }
}

}  // namespace

extern "C" int mg_linear_forward(const float* x, const float* W, const float* bias, float* z, int rows, int K, int N,
                                 void* stream) {
    MG_REQUIRE(x && W && z && rows > 0 && K > 0 && N > 0, "linear_forward: null pointer or empty shape");
    mg::tc::set_tf32(false); mg::tc::set_cache_mode(false);          // float32 parity arithmetic, weights packed per call
    return linear_fwd<float, float>(x, z, W, bias, rows, K, N, ACT_NONE, nullptr, as_stream(stream));
}

extern "C" int mg_linear_backward(const float* x, const float* W, const float* dz, float* dx, float* dW, float* db, int rows,
                                  int K, int N, void* stream) {
    MG_REQUIRE(W && dz && rows > 0 && K > 0 && N > 0, "linear_backward: null pointer or empty shape");
    cudaStream_t st = as_stream(stream);
    mg::tc::set_tf32(false); mg::tc::set_cache_mode(false);
    if (dx) {
        const int rc = linear_dgrad<float, float>(dz, dx, W, rows, K, N, nullptr, MUL_NONE, st);
        if (rc != MG_OK) return rc;
    }
    if (dW) {
        MG_REQUIRE(x, "linear_backward: the weight gradient needs the layer input");
        const int rc = linear_wgrad<float, float>(dz, x, dW, 0, rows, K, N, st);
        if (rc != MG_OK) return rc;
    }
    if (db) {
        bias_grad_kernel<<<(N + 31) / 32, dim3(32, 8), 0, st>>>(dz, db, rows, N);
        MG_LAUNCH_OK();
    }
    return MG_OK;
}

extern "C" int mg_act_dropout_forward(const float* z, const float* mask, float scale, int act, float* h, long long n,
                                      void* stream) {
    MG_REQUIRE(z && h && n > 0 && act >= ACT_NONE && act <= ACT_GELU, "act_dropout_forward: bad arguments");
    act_dropout_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(z, mask, scale, act, h, n);
    MG_LAUNCH_OK();
    return MG_OK;
}

extern "C" int mg_act_dropout_backward(const float* dh, const float* z, const float* mask, float scale, int act, float* dz,
                                       long long n, void* stream) {
    MG_REQUIRE(dh && z && dz && n > 0 && act >= ACT_NONE && act <= ACT_GELU, "act_dropout_backward: bad arguments");
    act_dropout_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(dh, z, mask, scale, act, dz, n);
    MG_LAUNCH_OK();
    return MG_OK;
}

// ================================================================================================
// Conv-type inner blocks as stand-alone operators: one "unit" = Conv1d (stride 1 or 2) or ConvTranspose1d (k5 s2 p2 op1)
// [+ BatchNorm1d, train or eval] + activation, float32, channels-last activations, on the contraction / reduction kernels
// of the fp32 parity mode.  Covers ConvBlock1D and NotesEncoder (ed_model.py:24-69: Conv1d k5/k3 s1 + BN + GELU),
// GeneratorDecoder's stack (models.py:52-83: ConvTranspose1d + BN + ReLU x2, ConvTranspose1d), the VAE's ConvEncoder
// (src/ae/model.py:9-25: Conv1d k5 s2 + BN + ReLU x3) and ConvDecoder (model.py:64-98: ... + Tanh).
// ================================================================================================
struct mg_convunit {
    mg_gan g;                       // reduction scratch only (partial, g_bn_sums) + bn_eps / bn_momentum
    float* stats = nullptr;         // [2][C] of the last forward
    char* arena = nullptr;
};

namespace {
enum { UNIT_CONV = 0, UNIT_CONVT = 1 };
enum { UACT_NONE = 0, UACT_RELU = 1, UACT_GELU = 2, UACT_TANH = 3 };

// y = act(v), v = bn ? (z - mean) * invstd * gamma + beta : z;  gelu' tile for GELU
__global__ void unit_apply_kernel(const float* __restrict__ z, float* __restrict__ y, float* __restrict__ gd, long long n, int C,
                                  const float* __restrict__ mean, const float* __restrict__ invstd,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, int act) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(i % C);
    float v = z[i];
    if (mean) v = fmaf((v - mean[c]) * invstd[c], gamma[c], beta[c]);
    float o = v;
    if (act == UACT_RELU) o = fmaxf(v, 0.0f);
    else if (act == UACT_GELU) { o = gelu_f(v); if (gd) gd[i] = gelu_grad_f(v); }
    else if (act == UACT_TANH) o = tanhf(v);
    y[i] = o;
}
// d(pre-activation) = dy * act'(.)   (ReLU / Tanh from the output y, GELU from the saved derivative tile)
__global__ void unit_dact_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ gd,
                                 float* __restrict__ dv, long long n, int act) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float d = dy[i];
    if (act == UACT_RELU) d = y[i] > 0.0f ? d : 0.0f;
    else if (act == UACT_GELU) d *= gd[i];
    else if (act == UACT_TANH) d *= 1.0f - y[i] * y[i];
    dv[i] = d;
}
// eval-mode BatchNorm backward: the statistics are constants, dz = dv * gamma * invstd
__global__ void unit_bn_eval_bwd_kernel(const float* __restrict__ dv, float* __restrict__ dz, long long n, int C,
                                        const float* __restrict__ invstd, const float* __restrict__ gamma) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dz[i] = dv[i] * gamma[(int)(i % C)] * invstd[(int)(i % C)];
}
inline unsigned blocks_for(long long n) { return (unsigned)((n + 255) / 256); }
}  // namespace

extern "C" int mg_convunit_create(mg_convunit** out) {
    MG_REQUIRE(out, "convunit_create: null argument");
    mg_convunit* u = new mg_convunit();
    u->g.cfg.bn_eps = 1e-5; u->g.cfg.bn_momentum = 0.1;
    u->g.partial_floats = (size_t)2 << 20;
    const size_t bytes = (u->g.partial_floats + 2048 + 4096) * sizeof(float);
    if (cudaMalloc(&u->arena, bytes) != cudaSuccess) { delete u; set_error("convunit_create: out of memory"); return MG_ERR_CUDA; }
    cudaMemset(u->arena, 0, bytes);
    float* f = reinterpret_cast<float*>(u->arena);
    u->g.partial = f; f += u->g.partial_floats;
    u->g.g_bn_sums = f; f += 2048;
    u->stats = f;
    *out = u;
    return MG_OK;
}
extern "C" void mg_convunit_destroy(mg_convunit* u) {
    if (!u) return;
    if (u->arena) cudaFree(u->arena);
    delete u;
}

// x [R][Lin][Cin] -> z (pre-BatchNorm) and y, both [R][Lout][Cout]; Lout = Lin / stride (conv) or 2 Lin (transposed conv).
// bn_* null: no BatchNorm.  train: batch statistics + running-statistics update; else the running statistics.
// mean / invstd [Cout] and gd (GELU' tile, GELU only) are outputs the backward needs.
extern "C" int mg_convunit_forward(mg_convunit* u, int kind, const float* x, const float* W, const float* bias, int R, int Lin,
                                   int Cin, int Cout, int ks, int stride, int pad, const float* gamma, const float* beta,
                                   float* running_mean, float* running_var, int train, int act, float* z, float* y, float* gd,
                                   float* mean, float* invstd, void* stream) {
    MG_REQUIRE(u && x && W && z && y && R > 0 && Lin > 0 && Cin > 0 && Cout > 0, "convunit_forward: null pointer or empty shape");
    MG_REQUIRE(kind == UNIT_CONV ? (stride == 1 || (stride == 2 && ks == 5 && pad == 2 && Lin % 2 == 0))
                                 : (kind == UNIT_CONVT && ks == 5 && stride == 2 && pad == 2),
               "convunit_forward: Conv1d stride 1, Conv1d k5 s2 p2, or ConvTranspose1d k5 s2 p2 op1");
    MG_REQUIRE(Cout <= 1024 && act >= UACT_NONE && act <= UACT_TANH, "convunit_forward: bad channel count / activation");
    MG_REQUIRE(!gamma || (beta && running_mean && running_var && mean && invstd), "convunit_forward: BatchNorm needs all its tensors");
    MG_REQUIRE(act != UACT_GELU || gd, "convunit_forward: GELU needs the derivative tile");
    cudaStream_t st = as_stream(stream);
    mg::tc::set_tf32(false); mg::tc::set_cache_mode(false);
    const int Lout = kind == UNIT_CONV ? Lin / stride : 2 * Lin;
    if (kind == UNIT_CONV)
        MG_TRY((conv_fwd<float, float>(x, z, W, bias, R, Lin, Cin, Cout, ks, stride, pad, ACT_NONE, nullptr, nullptr, nullptr,
                                       MUL_NONE, st)));
    else
        MG_TRY((upsample2_fwd<float, float>(x, z, W, bias, R, Lin, Cin, Cout, 5, Cout * 5, ACT_NONE, nullptr, MUL_NONE, 0, st)));
    const long long rows = (long long)R * Lout, n = rows * Cout;
    if (gamma) {
        if (train) {
            MG_TRY((colreduce<float, COL_SUM_SQ>(&u->g, z, Cout, nullptr, 0, nullptr, nullptr, nullptr, 1, 0, rows, Cout, u->stats,
                                                 Cout, 0, 0, 1.0f, 0, st)));
            bn_finalize_kernel<<<(Cout + 127) / 128, 128, 0, st>>>(u->stats, Cout, rows, (float)u->g.cfg.bn_eps,
                                                                  (float)u->g.cfg.bn_momentum, mean, invstd, running_mean,
                                                                  running_var, 1);
        } else {
            bn_eval_stats_kernel<<<(Cout + 127) / 128, 128, 0, st>>>(running_mean, running_var, (float)u->g.cfg.bn_eps, Cout, mean,
                                                                    invstd);
        }
        MG_LAUNCH_OK();
    }
    unit_apply_kernel<<<blocks_for(n), 256, 0, st>>>(z, y, gd, n, Cout, gamma ? mean : nullptr, invstd, gamma, beta, act);
    MG_LAUNCH_OK();
    return MG_OK;
}

// dy [R][Lout][Cout] -> dx (optional), dW +=, db +=, dgamma +=, dbeta += (caller zeroes); scratch: two [R][Lout][Cout] buffers
extern "C" int mg_convunit_backward(mg_convunit* u, int kind, const float* x, const float* W, int R, int Lin, int Cin, int Cout,
                                    int ks, int stride, int pad, const float* gamma, int train, int act, const float* z,
                                    const float* y, const float* gd, const float* mean, const float* invstd, const float* dy,
                                    float* scratch0, float* scratch1, float* dx, float* dW, float* dbias, float* dgamma,
                                    float* dbeta, void* stream) {
    MG_REQUIRE(u && x && W && z && y && dy && scratch0 && scratch1 && dW, "convunit_backward: null pointer");
    MG_REQUIRE(!gamma || (mean && invstd && dgamma && dbeta), "convunit_backward: BatchNorm needs its statistics and gradient buffers");
    cudaStream_t st = as_stream(stream);
    mg::tc::set_tf32(false); mg::tc::set_cache_mode(false);
    const int Lout = kind == UNIT_CONV ? Lin / stride : 2 * Lin;
    const long long rows = (long long)R * Lout, n = rows * Cout;
    unit_dact_kernel<<<blocks_for(n), 256, 0, st>>>(dy, y, gd, scratch0, n, act);            // d(BatchNorm output)
    MG_LAUNCH_OK();
    const float* dz = scratch0;
    if (gamma) {
        if (train) {
            MG_TRY((bn_backward<float>(&u->g, z, scratch0, scratch1, rows, Cout, mean, invstd, gamma, dgamma, dbeta, st)));
        } else {
            MG_TRY((colreduce<float, COL_BN_BWD, float>(&u->g, z, Cout, scratch0, Cout, mean, invstd, nullptr, 1, 0, rows, Cout,
                                                        u->g.g_bn_sums, Cout, 0, 0, 1.0f, 0, st)));
            add2_kernel<<<(Cout + 127) / 128, 128, 0, st>>>(dbeta, u->g.g_bn_sums, dgamma, u->g.g_bn_sums + Cout, Cout);
            MG_LAUNCH_OK();
            unit_bn_eval_bwd_kernel<<<blocks_for(n), 256, 0, st>>>(scratch0, scratch1, n, Cout, invstd, gamma);
            MG_LAUNCH_OK();
        }
        dz = scratch1;
    }
    if (dbias)
        MG_TRY((colreduce<float, COL_SUM>(&u->g, dz, Cout, nullptr, 0, nullptr, nullptr, nullptr, 1, 0, rows, Cout, dbias, 0, 0, 0,
                                          1.0f, 1, st)));
    if (kind == UNIT_CONV) {
        MG_TRY((conv_wgrad<float, float>(dz, x, dW, 0, rows, Lin, Cin, Cout, ks, stride, pad, st)));
        if (dx) {
            if (stride == 1)
                MG_TRY((conv_s1_dgrad<float, float>(dz, dx, W, R, Lin, Cin, Cout, ks, pad, nullptr, nullptr, MUL_NONE, 0, st)));
            else
                MG_TRY((upsample2_fwd<float, float>(dz, dx, W, nullptr, R, Lout, Cout, Cin, 5, Cin * 5, ACT_NONE, nullptr, MUL_NONE,
                                                    0, st)));
        }
    } else {
        MG_TRY((convT_wgrad<float, float>(x, dz, dW, R, Lin, Cin, Cout, st)));
        if (dx)
            MG_TRY((conv_fwd<float, float>(dz, dx, W, nullptr, R, 2 * Lin, Cout, Cin, 5, 2, 2, ACT_NONE, nullptr, nullptr, nullptr,
                                           MUL_NONE, st, Cout * 5, 5)));
    }
    return MG_OK;
}
