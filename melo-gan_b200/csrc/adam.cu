// adam.cu -- fused, vectorised Adam / AdamW over a flat float32 parameter segment (sm_100a).
//
// A-10  torch.optim.Adam.step as called at reference src/gan/train_gan.py:136-145,204,248
//       (AdamW form: src/ae/train_ae.py:79, src/emotion_discriminator/train_ed.py:97).
// HBM-bound: 16 B read (p, g, m, v) + 12 B written (p, m, v) per parameter = 28 B/param;
// one launch covers a whole optimizer group because the host keeps each group's parameters,
// gradients and moments in flat buffers (the same buffers the gradient all-reduce uses).
// Operation order follows torch's single-tensor Adam so that a teacher-forced step agrees to
// float32 rounding:  m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, 1-b2);
//                    denom = sqrt(v)/sqrt(bc2) + eps; p.addcdiv_(m, denom, -lr/bc1).
#include "common.cuh"

namespace {

struct AdamScalars {
    float neg_step_size;  // -(lr / (1 - beta1^t))
    float bc2_sqrt;       // sqrt(1 - beta2^t)
    float clip;           // clip_grad_norm_ coefficient min(1, max_norm / (||g|| + 1e-6)); 1 without clipping
    float pad;
};

// sum of squares of the flat gradient, one double per CTA (clip_grad_norm_: reference src/ae/train_ae.py:121)
__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ grad, long long n, double* __restrict__ parts) {
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const float4* g4 = reinterpret_cast<const float4*>(grad);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (long long i = tid; i < n4; i += stride) {
        const float4 g = __ldg(g4 + i);
        a0 = fmaf(g.x, g.x, a0); a1 = fmaf(g.y, g.y, a1); a2 = fmaf(g.z, g.z, a2); a3 = fmaf(g.w, g.w, a3);
    }
    for (long long i = (n4 << 2) + tid; i < n; i += stride) a0 = fmaf(grad[i], grad[i], a0);
    double acc = (double)a0 + (double)a1 + (double)a2 + (double)a3;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double warp_sum[8];
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += warp_sum[w];
        parts[blockIdx.x] = t;
    }
}

// One thread: advance the device step counter and derive the bias corrections in float64
// (python computes them with float64 `**`).  Keeps the whole update CUDA-graph capturable.
__global__ void adam_prepare_kernel(long long* step_dev, long long step_host, double lr, double beta1, double beta2,
                                    AdamScalars* out, const double* parts, int nparts, float grad_scale, float max_norm,
                                    float* norm_out) {
    double ss = 0.0;                                   // one warp: lanes share the partial sums
    if (parts) {
        for (int i = threadIdx.x; i < nparts; i += 32) ss += parts[i];
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if (threadIdx.x != 0) return;
    long long t = step_host;
    if (step_dev) { t = *step_dev + 1; *step_dev = t; }
    const double bc1 = 1.0 - pow(beta1, (double)t);
    const double bc2 = 1.0 - pow(beta2, (double)t);
    out->neg_step_size = (float)(-(lr / bc1));
    out->bc2_sqrt = (float)sqrt(bc2);
    float clip = 1.0f;
    if (parts) {   // torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to <= 1
        const float total = (float)(sqrt(ss) * (double)grad_scale);
        if (norm_out) *norm_out = total;
        const float coef = max_norm / (total + 1e-6f);
        clip = coef < 1.0f ? coef : 1.0f;
    }
    out->clip = clip;
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float w_lerp, float one_minus_w,
                                         float beta2, float one_minus_b2, float eps, float decay_mul,
                                         const AdamScalars s) {
    p = p * decay_mul;  // AdamW: p.mul_(1 - lr*wd); multiplies by exactly 1.0f otherwise
    // torch lerp (ATen Lerp.h, vectorised form): fma(coeff, end - start, base) with
    // (coeff, base) = weight < 0.5 ? (w, start) : (w - 1, end)
    const float diff = __fsub_rn(g, m);
    m = (w_lerp < 0.5f) ? __fmaf_rn(w_lerp, diff, m) : __fmaf_rn(-one_minus_w, diff, g);
    v = __fadd_rn(__fmul_rn(v, beta2), __fmul_rn(__fmul_rn(one_minus_b2, g), g));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), s.bc2_sqrt), eps);
    p = __fadd_rn(p, __fdiv_rn(__fmul_rn(s.neg_step_size, m), denom));
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ param, const float* __restrict__ grad,
                                                   float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                   long long n, float w, float beta2, float omb2, float eps,
                                                   float decay_mul,
                                                   float grad_scale, const AdamScalars* __restrict__ sc,
                                                   __nv_bfloat16* __restrict__ bf16_copy) {
    const AdamScalars s = *sc;
    grad_scale *= s.clip;
    const float omw = 1.0f - w;  // torch evaluates (1 - weight) in the tensor dtype
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float4* p4 = reinterpret_cast<float4*>(param);
    const float4* g4 = reinterpret_cast<const float4*>(grad);
    float4* m4 = reinterpret_cast<float4*>(exp_avg);
    float4* v4 = reinterpret_cast<float4*>(exp_avg_sq);
    for (long long i = tid; i < n4; i += stride) {
        float4 p = p4[i], g = __ldg(g4 + i), m = m4[i], v = v4[i];
        if (grad_scale != 1.0f) { g.x *= grad_scale; g.y *= grad_scale; g.z *= grad_scale; g.w *= grad_scale; }
        adam_one(p.x, g.x, m.x, v.x, w, omw, beta2, omb2, eps, decay_mul, s);
        adam_one(p.y, g.y, m.y, v.y, w, omw, beta2, omb2, eps, decay_mul, s);
        adam_one(p.z, g.z, m.z, v.z, w, omw, beta2, omb2, eps, decay_mul, s);
        adam_one(p.w, g.w, m.w, v.w, w, omw, beta2, omb2, eps, decay_mul, s);
        p4[i] = p; m4[i] = m; v4[i] = v;
        if (bf16_copy) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
            uint2 pk;
            pk.x = *reinterpret_cast<unsigned*>(&lo);
            pk.y = *reinterpret_cast<unsigned*>(&hi);
            reinterpret_cast<uint2*>(bf16_copy)[i] = pk;
        }
    }
    // tail (n not a multiple of 4)
    for (long long i = (n4 << 2) + tid; i < n; i += stride) {
        float p = param[i], g = grad[i] * grad_scale, m = exp_avg[i], v = exp_avg_sq[i];
        adam_one(p, g, m, v, w, omw, beta2, omb2, eps, decay_mul, s);
        param[i] = p; exp_avg[i] = m; exp_avg_sq[i] = v;
        if (bf16_copy) bf16_copy[i] = __float2bfloat16_rn(p);
    }
}

AdamScalars* g_scalars[16] = {nullptr};  // one slot per device
double* g_parts[16] = {nullptr};         // per device: 4 rings of per-CTA partial sums of squares
constexpr int kMaxParts = 2048;

int adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, double lr, double beta1,
              double beta2, double eps, double weight_decay, int decoupled, float grad_scale, float max_norm,
              float* norm_out_dev, long long step, long long* step_dev, uint16_t* bf16_copy, void* stream);

}  // namespace

extern "C" int mg_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                            double lr, double beta1, double beta2, double eps, double weight_decay,
                            int decoupled, float grad_scale, long long step, long long* step_dev,
                            uint16_t* bf16_copy, void* stream) {
    return adam_step(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, decoupled, grad_scale, 0.0f,
                     nullptr, step, step_dev, bf16_copy, stream);
}

extern "C" int mg_adam_step_clipped(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                                    double lr, double beta1, double beta2, double eps, double weight_decay, int decoupled,
                                    float grad_scale, float max_norm, float* norm_out_dev, long long step,
                                    long long* step_dev, uint16_t* bf16_copy, void* stream) {
    MG_REQUIRE(max_norm > 0.0f, "adam_step_clipped: max_norm must be positive");
    return adam_step(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, decoupled, grad_scale, max_norm,
                     norm_out_dev, step, step_dev, bf16_copy, stream);
}

namespace {
int adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, double lr, double beta1,
              double beta2, double eps, double weight_decay, int decoupled, float grad_scale, float max_norm,
              float* norm_out_dev, long long step, long long* step_dev, uint16_t* bf16_copy, void* stream) {
    MG_REQUIRE(n >= 0, "adam: negative n");
    if (n == 0) return MG_OK;
    MG_REQUIRE(param && grad && exp_avg && exp_avg_sq, "adam: null pointer");
    MG_REQUIRE(step_dev || step >= 1, "adam: step must be >= 1");
    MG_REQUIRE(weight_decay == 0.0 || decoupled, "adam: only decoupled (AdamW) weight decay is implemented");
    MG_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
               "adam: buffers must be 16-byte aligned");
    int dev = 0;
    MG_CUDA_OK(cudaGetDevice(&dev));
    MG_REQUIRE(dev < 16, "adam: device index too large");
    if (!g_scalars[dev]) MG_CUDA_OK(cudaMalloc(&g_scalars[dev], sizeof(AdamScalars) * 64));
    // a distinct scalar slot per (stream-ordered) call so that back-to-back groups do not race
    static thread_local unsigned slot = 0;
    AdamScalars* sc = g_scalars[dev] + (slot++ & 63);
    cudaStream_t st = mg::as_stream(stream);
    long long blocks = ((n >> 2) + 255) / 256;
    const long long cap = (long long)mg::num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const double* parts = nullptr;
    int nparts = 0;
    if (max_norm > 0.0f) {   // fused clip_grad_norm_: one extra read of the gradient (4 B/param), no host round trip
        if (!g_parts[dev]) MG_CUDA_OK(cudaMalloc(&g_parts[dev], sizeof(double) * kMaxParts * 4));
        static thread_local unsigned pslot = 0;
        double* p = g_parts[dev] + (size_t)(pslot++ & 3) * kMaxParts;
        nparts = (int)(blocks < kMaxParts ? blocks : kMaxParts);
        mg::ProbeScope probe(mg::PROBE_ADAM, 0.0, 4.0 * (double)n, st);
        grad_sumsq_kernel<<<nparts, 256, 0, st>>>(grad, n, p);
        MG_LAUNCH_OK();
        parts = p;
    }
    adam_prepare_kernel<<<1, 32, 0, st>>>(step_dev, step, lr, beta1, beta2, sc, parts, nparts, grad_scale, max_norm, norm_out_dev);
    MG_LAUNCH_OK();
    const float decay_mul = (weight_decay != 0.0) ? (float)(1.0 - lr * weight_decay) : 1.0f;
    mg::tc_weights_changed(param, n);     // packed tensor-core copies of these parameters are stale from here on
    mg::ProbeScope probe(mg::PROBE_ADAM, 0.0, 28.0 * (double)n + (bf16_copy ? 2.0 * (double)n : 0.0), st);
    adam_kernel<<<(int)blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, (float)(1.0 - beta1), (float)beta2,
                                             (float)(1.0 - beta2), (float)eps, decay_mul,
                                             grad_scale, sc, reinterpret_cast<__nv_bfloat16*>(bf16_copy));
    MG_LAUNCH_OK();
    return MG_OK;
}
}  // namespace
