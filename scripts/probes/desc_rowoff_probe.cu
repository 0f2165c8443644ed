// Probe: can a K-major SWIZZLE_128B shared-memory operand descriptor start at a row that is NOT a multiple of 8
// (i.e. not 1024-byte aligned)?  A holds 144 rows x 64 bf16 laid out the way TMA would write them (16-byte piece index
// XOR (absolute row & 7)); for each row shift d the MMA reads rows d..d+127 through a descriptor whose start address is
// base + d*128, with the descriptor's base_offset field either 0 or ((addr >> 7) & 7).  Prints which variant matches.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../melo-gan_b200/csrc/gemm_tc.cuh"
using namespace mg::tc;

__device__ __forceinline__ uint64_t desc_bo(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
    return make_smem_desc(saddr, lbo, sbo) | ((uint64_t)(base_off & 7) << 49);
}

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int d, int mode) {
    extern __shared__ unsigned char raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sa = base;                 // 144 rows x 128 B
    unsigned char* sb = base + 144 * 128;     // 128 rows x 128 B (18432 is 1024-aligned)
    uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 128 * 128);
    uint32_t* tbase = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 144 * 8; i += 128) {
        const int r = i / 8, c = i % 8;
        *reinterpret_cast<uint4*>(sa + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 64 + c * 8);
    }
    for (int i = threadIdx.x; i < 128 * 8; i += 128) {
        const int r = i / 8, c = i % 8;
        *reinterpret_cast<uint4*>(sb + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 64 + c * 8);
    }
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(tbase, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tacc = *tbase;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = make_idesc(128, 0, 0);
        const uint32_t a_addr = smem_u32(sa) + d * 128, b_addr = smem_u32(sb);
        for (int k = 0; k < 4; ++k) {
            const uint32_t aa = a_addr + k * 32;
            const uint32_t bo = mode == 0 ? 0u : ((aa >> 7) & 7);
            umma_f16(tacc, desc_bo(aa, 16, 1024, bo), make_smem_desc(b_addr + k * 32, 16, 1024), idesc, k ? 1u : 0u);
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < 128; c0 += 16) {
        float v[16];
        tmem_ld16(tacc + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * 128 + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tacc, 128); }
}

int main() {
    std::vector<__nv_bfloat16> hA(144 * 64), hB(128 * 64);
    std::vector<float> fA(144 * 64), fB(128 * 64);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)(rand() % 17 - 8); hA[i] = __float2bfloat16(fA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)(rand() % 13 - 6); hB[i] = __float2bfloat16(fB[i]); }
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 128 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    const int smem = 144 * 128 + 128 * 128 + 64 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> hD(128 * 128);
    for (int mode = 0; mode < 2; ++mode)
        for (int d = 0; d < 16; ++d) {
            cudaMemset(dD, 0, 128 * 128 * 4);
            probe<<<1, 128, smem>>>(dA, dB, dD, d, mode);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d d %d: CUDA error %s\n", mode, d, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hD.data(), dD, 128 * 128 * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 128; ++n) {
                    float ref = 0;
                    for (int k = 0; k < 64; ++k) ref += fA[(m + d) * 64 + k] * fB[n * 64 + k];
                    if (ref != hD[m * 128 + n]) ++bad;
                }
            printf("mode %s  row shift %2d : %s (%d mismatches)\n", mode ? "base_offset=(addr>>7)&7" : "base_offset=0", d,
                   bad ? "WRONG" : "exact", bad);
        }
    return 0;
}
