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
