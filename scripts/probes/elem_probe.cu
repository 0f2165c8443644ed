// Stand-alone timing of the element-wise kernels of csrc/elem.cuh at the bench shapes (B = 8192), with grid variants, against
// a plain device-to-device copy of the same bytes.  Build + run (GPU box):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I melo-gan_b200/csrc -I include scripts/probes/elem_probe.cu -o /tmp/elem_probe && /tmp/elem_probe
#include <cstdio>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
namespace mg { void set_error(const char*, ...) {} void count_launch() {} int num_sms() { return 148; }
void tc_weights_changed(const float*, long long) {} }
#include "elem.cuh"
using namespace mg;
using bf = __nv_bfloat16;

template <typename F> float timed(F f, int n = 10) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    static char* flush = nullptr; if (!flush) cudaMalloc(&flush, 512 << 20);
    float tot = 0;
    for (int i = 0; i < n + 2; ++i) {
        cudaMemsetAsync(flush, i, 512 << 20);
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (i >= 2) tot += ms;
    }
    return tot / n;
}

// candidate: persistent row broadcast (block loops over samples, 16-byte accesses)
__global__ void __launch_bounds__(256) bcast8_persist(const float* __restrict__ src, const bf* __restrict__ ref, bf* __restrict__ out,
                                                      int S, int L, int C, float scale) {
    const int per8 = L * C / 8;
    for (int s = blockIdx.x; s < S; s += gridDim.x) {
        const uint4* r4 = reinterpret_cast<const uint4*>(ref + (long long)s * L * C);
        uint4* o4 = reinterpret_cast<uint4*>(out + (long long)s * L * C);
        const int c = (threadIdx.x * 8) % C;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = src[(long long)s * C + c + e] * scale;
#pragma unroll 8
        for (int i = threadIdx.x; i < per8; i += 256) {
            const uint4 q = __ldcs(r4 + i);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xFFFF0000u);
                const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * e] * (lo > 0.f ? 1.f : 0.2f), f[2 * e + 1] * (hi > 0.f ? 1.f : 0.2f));
                o[e] = *reinterpret_cast<const uint32_t*>(&h);
            }
            __stcs(o4 + i, make_uint4(o[0], o[1], o[2], o[3]));
        }
    }
}

int main() {
    const double peak = 6546.6;
    {   // row broadcast x mask: critic, R = 3 * 8192 samples x 64 positions x 256 channels
        const int R = 24576, L = 64, C = 256; const size_t n = (size_t)R * L * C;
        float* src; bf *ref, *out; cudaMalloc(&src, R * C * 4); cudaMalloc(&ref, n * 2); cudaMalloc(&out, n * 2);
        cudaMemset(src, 0, R * C * 4); cudaMemset(ref, 0, n * 2);
        const double gb = n * 4 / 1e9;
        float ms = timed([&] { cudaMemcpyAsync(out, ref, n * 2, cudaMemcpyDeviceToDevice); });
        printf("bcast  copy (memcpy D2D)              %7.1f us  %6.0f GB/s (%.2f)\n", ms * 1e3, gb / ms * 1e3, gb / ms * 1e3 / peak);
        const int chunks = (int)((L * C / 4 + 1023) / 1024);
        ms = timed([&] { bcast_rows_mul_kernel<float, bf><<<dim3(chunks, R), 256>>>(src, ref, out, L, C, 1.0f / L, nullptr, MUL_LRELU_SIGN); });
        printf("bcast  product kernel grid (%d,%d)    %7.1f us  %6.0f GB/s (%.2f)\n", chunks, R, ms * 1e3, gb / ms * 1e3, gb / ms * 1e3 / peak);
        for (int g : {148 * 2, 148 * 4, 148 * 8, 148 * 16, 24576}) {
            ms = timed([&] { bcast8_persist<<<g, 256>>>(src, ref, out, R, L, C, 1.0f / L); });
            printf("bcast  persistent grid %6d           %7.1f us  %6.0f GB/s (%.2f)\n", g, ms * 1e3, gb / ms * 1e3, gb / ms * 1e3 / peak);
        }
    }
    {   // BatchNorm apply: 524288 rows x 128 channels fp32 -> bf16
        const long long rows = 524288; const int C = 128; const size_t n = rows * C;
        float *x, *mean, *is, *g, *b; bf* y; cudaMalloc(&x, n * 4); cudaMalloc(&y, n * 2);
        cudaMalloc(&mean, C * 4); cudaMalloc(&is, C * 4); cudaMalloc(&g, C * 4); cudaMalloc(&b, C * 4);
        cudaMemset(x, 0, n * 4); cudaMemset(mean, 0, C * 4); cudaMemset(is, 0, C * 4); cudaMemset(g, 0, C * 4); cudaMemset(b, 0, C * 4);
        const double gb = n * 6 / 1e9;
        for (int per_sm : {2, 4, 8, 16, 32}) {
            float ms = timed([&] { bn_relu_apply_kernel<float, bf><<<148 * per_sm, 256>>>(x, y, (long long)n / 4, C, mean, is, g, b); });
            printf("bn_relu_apply grid 148x%-2d               %7.1f us  %6.0f GB/s (%.2f)\n", per_sm, ms * 1e3, gb / ms * 1e3, gb / ms * 1e3 / peak);
        }
        ColReduceArgs a{}; float* partial; cudaMalloc(&partial, 4 << 20);
        a.x = x; a.ldx = C; a.r0 = 0; a.r1 = rows; a.C = C; a.partial = partial; a.roww_div = 1;
        for (int nch : {148 * 2, 148 * 4, 148 * 8, 148 * 16}) {
            a.rows_per_chunk = (int)(((rows + nch - 1) / nch + 7) / 8 * 8);
            const int nchunk = (int)((rows + a.rows_per_chunk - 1) / a.rows_per_chunk);
            float ms = timed([&] { colreduce_flat_kernel<float, float, COL_SUM_SQ><<<dim3(1, nchunk), 256>>>(a); });
            printf("colreduce SUM_SQ fp32 chunks %5d       %7.1f us  %6.0f GB/s (%.2f)\n", nchunk, ms * 1e3, n * 4 / 1e9 / ms * 1e3, n * 4 / 1e9 / ms * 1e3 / peak);
        }
    }
    return 0;
}
