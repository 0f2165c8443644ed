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
};

// One thread: advance the device step counter and derive the bias corrections in float64
// (python computes them with float64 `**`).  Keeps the whole update CUDA-graph capturable.
__global__ void adam_prepare_kernel(long long* step_dev, long long step_host, double lr, double beta1, double beta2,
                                    AdamScalars* out) {
    long long t = step_host;
    if (step_dev) { t = *step_dev + 1; *step_dev = t; }
    const double bc1 = 1.0 - pow(beta1, (double)t);
    const double bc2 = 1.0 - pow(beta2, (double)t);
    out->neg_step_size = (float)(-(lr / bc1));
    out->bc2_sqrt = (float)sqrt(bc2);
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

}  // namespace

extern "C" int mg_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                            double lr, double beta1, double beta2, double eps, double weight_decay,
                            int decoupled, float grad_scale, long long step, long long* step_dev,
                            uint16_t* bf16_copy, void* stream) {
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
    adam_prepare_kernel<<<1, 1, 0, st>>>(step_dev, step, lr, beta1, beta2, sc);
    MG_LAUNCH_OK();
    const float decay_mul = (weight_decay != 0.0) ? (float)(1.0 - lr * weight_decay) : 1.0f;
    long long blocks = ((n >> 2) + 255) / 256;
    const long long cap = (long long)mg::num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    mg::ProbeScope probe(mg::PROBE_ADAM, 0.0, 28.0 * (double)n + (bf16_copy ? 2.0 * (double)n : 0.0), st);
    adam_kernel<<<(int)blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, (float)(1.0 - beta1), (float)beta2,
                                             (float)(1.0 - beta2), (float)eps, decay_mul,
                                             grad_scale, sc, reinterpret_cast<__nv_bfloat16*>(bf16_copy));
    MG_LAUNCH_OK();
    return MG_OK;
}
