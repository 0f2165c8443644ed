// gemm_tc.cuh -- tcgen05 / TMEM / TMA implicit-GEMM kernels for sm_100a (bf16 operands, fp32 accumulate).
//
// Same two contractions as gemm_simt.cuh, for the layers that are tensor-core shaped:
//
//   tc_tapgemm   Out[b, m, n] = epi( sum_t sum_k A[b, s*m + r_t, k] * Wp[t][n][k] )        (s = 1 or 2)
//     A tiles come straight from the channels-last activation through ONE 4-D TMA tensor map
//     (k, parity, row/stride, sample): a tap is a coordinate shift, the conv zero padding and the
//     sample boundaries are TMA out-of-bounds zero fill, and a stride-2 conv reads the even/odd row
//     plane it needs.  128 rows x 64 k bf16 tiles land in shared memory 128B-swizzled (K-major), the
//     packed weight tile [BN n][64 k] likewise; one elected thread issues tcgen05.mma (M=128, N=BN,
//     K=16) into a TMEM accumulator; four epilogue warps read it back with tcgen05.ld and apply
//     bias / folded-BN scale / ReLU|LeakyReLU|GELU / derivative masks before the global store.
//   tc_wgrad     dW[t][n][k] += sum_{b,m} G[b, m, n] * A[b, s*m + r_t, k]
//     the reduction runs over positions, so BOTH operands are MN-major in shared memory (rows of
//     128 B = 64 channels, one row per position); split over CTAs along the positions, fp32 atomics
//     into the reference-layout gradient.
//
// Two kernels run tc_tapgemm.  tc_tapgemm_kernel: one tile per CTA (192 threads: warp 0 = TMA producer, warp 1 = TMEM
// allocator + MMA issuer, warps 2..5 = epilogue, TMEM lane quarter = warp_id % 4), 3-stage smem ring (full/empty
// mbarriers), tcgen05.commit releases a stage back to the producer and finally signals the epilogue.
// tc_tapgemm_ws_kernel (the one that carries the training cycle): weight-stationary and persistent -- the packed weights
// of ALL taps of one 64/128-wide output slab stay in shared memory, activation tiles stream through a ring, taps of one
// plane share a tile with a halo (row-offset descriptors), two TMEM accumulators, one epilogue warpgroup per 32
// accumulator columns, results leave through a 128B-swizzled staging tile and TMA bulk stores, mask tiles arrive by TMA.
// The same kernels run the float32 Linears of bf16 mode as kind::tf32 (TF32 template flag: fp32 operands straight from
// HBM, 32 elements per 128-byte k-block).  MELOGAN_TRACE=1 prints one line per launch; MELOGAN_TC_DEBUG is an ablation
// switch for profiling (bits: 1 no stores, 2 no MMA, 4 no activation loads, 8 no mask loads).
#pragma once
#include <cuda.h>

#include <stdlib.h>
#include <cstdio>

#include <algorithm>
#include <type_traits>

#include "gemm_simt.cuh"

namespace mg {
namespace tc {

constexpr int kStages = 3;
constexpr int kTcMaxTaps = 20;  // banded forms of the 4-channel layers use up to 20 row taps (banded.cuh)
constexpr int kTileM = 128;
constexpr int kTileK = 64;      // bf16 elements = 128 bytes = one swizzle row

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok = 0;
    for (uint32_t spins = 0; !ok; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!ok && spins > (1u << 26)) asm volatile("trap;");   // a protocol bug must not hang the GPU
    }
}
// exactly one lane of a converged warp gets true; unlike `lane == 0` the compiler keeps the surrounding values in the
// uniform datapath, which tcgen05.mma / tcgen05.commit operands need (otherwise every issue is wrapped in a vote loop)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// ---- CTA-pair (cta_group::2) forms: two CTAs of one cluster (the two SMs of a TPC) run ONE 256-row MMA; each holds its own
// 128 activation rows and HALF of the weight columns in its shared memory, and its own 128 accumulator lanes in its TMEM ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {          // every thread of both CTAs
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this kernel's layout) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
// Remote arrive with the default (.release.cta) semantics: what is handed over is TMEM, ordered by tcgen05.wait::ld +
// tcgen05.fence, not generic memory.  The .release.cluster form compiles to MEMBAR.ALL.GPU + CCTL.IVALL (an L1 invalidate)
// per arrive -- with it every epilogue warp stalled once per tile and the pair kernel lost 1.7x on short tiles.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a pair: data into THIS CTA's shared memory, complete_tx on a barrier of either CTA of the pair
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0, int c1,
                                                 int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {   // one warp of EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// arrives (once the MMAs issued so far have retired) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// fp32 operands in shared memory, read as TF32 (10-bit mantissa), fp32 accumulate: K = 8 per instruction = 32 bytes
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
template <bool TF32>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    if (TF32) umma_tf32(d_tmem, a_desc, b_desc, idesc, acc);
    else umma_f16(d_tmem, a_desc, b_desc, idesc, acc);
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread = its warp's lane = TMEM lane)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 32 consecutive fp32 columns, asynchronous: call tmem_wait_ld() before using the registers
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
          "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
          "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_async(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32_async(taddr, r); }
__device__ __forceinline__ void tmem_ld_async(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16_async(taddr, r); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B, version 1
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;        // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;        // LayoutType::SWIZZLE_128B
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): bf16 x bf16 -> f32, M = 128
// operand format: 1 = BF16 (kind::f16), 2 = TF32 (kind::tf32)
__host__ __device__ constexpr uint32_t make_idesc(int N, int a_mn_major, int b_mn_major, uint32_t fmt = 1) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}
// the same for a CTA pair: M = 256 (128 rows in each CTA's TMEM), N = the pair's slab width
__host__ __device__ constexpr uint32_t make_idesc_pair(int N, uint32_t fmt = 1) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// tap-GEMM
// ---------------------------------------------------------------------------------------------
struct TcTapArgs {
    int ntaps, kblocks;                 // kblocks = K / ktile per tap
    int ktile;                          // elements per 128-byte k-block: 64 (bf16) or 32 (fp32 read as TF32)
    int a_p[kTcMaxTaps], a_dm[kTcMaxTaps];  // per tap: plane (dim 1) and row shift (dim 2) of the A map
    int b_row[kTcMaxTaps];                  // per tap: first row of that tap's [N][K] block in the packed weight
    // tap groups (weight-stationary kernel): the taps of a group read the same plane at row shifts within `halo` rows of
    // g_dmin, so ONE activation tile of 128 + halo rows per k-block serves all of them -- each tap's MMA operand is the
    // same shared-memory tile entered `a_dm - g_dmin` rows further down (a SWIZZLE_128B K-major descriptor may start at any
    // 128-byte row: the swizzle is a function of the absolute shared-memory address, scripts/probes/desc_rowoff_probe.cu).
    int ngroups, halo;
    int g_p[kTcMaxTaps], g_dmin[kTcMaxTaps], g_first[kTcMaxTaps], g_count[kTcMaxTaps], g_tap[kTcMaxTaps];
    int mpt, bpt;                       // tile = bpt samples x mpt rows, mpt * bpt = 128
    int Mper, B, N;
    int mper_shift, mpt_shift;          // log2(Mper), log2(mpt): both are powers of two (set by run_tc_tap)
    int il, il_shift;                   // halo tiles of bpt > 1 samples are loaded ROW-INTERLEAVED ([row][sample][k], il = bpt):
                                        // a tap's row shift d is then the byte offset d * il * 128 for every sample of the tile
                                        // at once (sample-major boxes cannot be entered at a common shift).  Accumulator lane r
                                        // holds tile row (r % il) * mpt + r / il.  il = 1: plain layout, lane = row.
    int n_perm_q, n_perm_p;             // bias index permutation (the packed weight is already permuted)
    void* Out; long long o_bstride; int o_mstride; int o_off;
    const float* bias; const float* col_scale; int act;
    const void* mul_src; int mul_mode; void* aux;
    float alpha; int accumulate;
    int kb_outer;                       // one-tile kernel: walk k-blocks outermost, taps innermost (fp32_tc: the five small-term
                                        // segments of ALL taps before the full-size one, see kSplitA)
    int tma_mask;                       // ... and the mask tile (f' of the saved activation) arrives by TMA as well
    int tma_store;                      // weight-stationary kernel: tiles leave through shared memory + TMA bulk stores
    int nsb;                            // ... through a ring of nsb (1 or 2) staging tiles: the bulk store of one tile drains
                                        // while the next tile is read from TMEM, computed and staged
    int nmb;                            // TMA mask tiles: ring of nmb (1 or 2) buffers; with 2 the tile after next is in flight
                                        // while this one is drained (a single buffer exposes one L2/HBM latency per tile)
    int reverse;                        // walk the M tiles from the last to the first (see run_tc_tap)
    float* pool_out; float pool_scale;  // fused mean over the rows of a sample (see ws_pool_*): pool_out[b * N + n]
    int pool_atomic;                    // a sample spans several tiles (Mper > 128): atomicAdd into a zeroed pool_out
    int* pool_done;                     // host only
    int skip_out;                       // the caller only wants the fused pooling (and the derivative tile): Out is not stored
    float* colsum_out; int colsum_tiles; int* colsum_done;   // fused column sums over the row tiles < colsum_tiles (bias gradients)
    float* stats_out; int* stats_done;  // fused BatchNorm statistics of a float32 output: [sum | sum of squares][N]
    int rot_step;                       // weight-stationary kernels: slab y starts its walk y * rot_step tiles further (see ws_row_tile)
    int dbg;                            // MELOGAN_TC_DEBUG bits (profiling only): 1 = no epilogue stores, 2 = no MMA, 4 = no A loads
};

// Row tile a CTA works on at step `tile` of its grid-strided walk: optionally from the last tile to the first (snake
// order between successive launches), and rotated by rot_step tiles per output slab -- when many slabs stream the SAME
// activation (decoder.pre.2: 128 slabs over 64 row tiles) a common order has every SM asking the same few L2 slices for the
// same lines at the same time (measured: 6500 clocks per tile pair against 2048 of MMA); staggered starts spread them.
__device__ __forceinline__ int ws_row_tile(const TcTapArgs& P, int tile, int mtiles) {
    int t = P.reverse ? mtiles - 1 - tile : tile;
    if (P.rot_step) { t += (int)blockIdx.y * P.rot_step; t %= mtiles; }
    return t;
}

__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&v)[8]) {      // one 16-byte store
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = make_uint4(*reinterpret_cast<unsigned*>(&a), *reinterpret_cast<unsigned*>(&b),
                                              *reinterpret_cast<unsigned*>(&c), *reinterpret_cast<unsigned*>(&d));
}

// Column loop of the drain for one thread's row.  ACT/MUL/AFF/AUX are compile-time unless GEN (then read from P).
template <int NC, int ACT, int MUL, bool AFF, bool AUX, bool GEN, bool kPrefetchMask, int KMV, typename TO, typename TMSK>
__device__ __forceinline__ void drain_cols(const TcTapArgs& P, uint32_t taddr, int c_begin, bool row_ok, TO* __restrict__ orow,
                                           const TMSK* __restrict__ mrow, TO* __restrict__ xrow, const uint4 (&mreg)[KMV],
                                           const float4* sc4, const float4* bi4) {
    static_assert(NC % 32 == 0, "column range must be whole groups of 32");
    const int act = GEN ? P.act : ACT, mul = GEN ? P.mul_mode : MUL;
    const bool aff = GEN ? true : AFF, aux = GEN ? (P.aux != nullptr) : AUX, acc = GEN ? (P.accumulate != 0) : false;
#pragma unroll 1
    for (int c0 = c_begin; c0 < c_begin + NC; c0 += 32) {
        uint32_t raw[32];
        tmem_ld32_async(taddr + (uint32_t)c0, raw);
        tmem_wait_ld();
        if (!row_ok) continue;
#pragma unroll
        for (int g8 = 0; g8 < 32; g8 += 8) {          // 8 columns: one 16-byte (bf16) or two 16-byte (float) stores
            float ms[8], old[8], x[8], gd[8];
            if (mul != MUL_NONE) {
                if (kPrefetchMask) {
                    const uint4 mv = mreg[((c0 - c_begin + g8) >> 3) < KMV ? ((c0 - c_begin + g8) >> 3) : 0];
                    const uint32_t w[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        ms[2 * e] = __uint_as_float(w[e] << 16);
                        ms[2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u);
                    }
                } else {
#pragma unroll
                    for (int h = 0; h < 8; h += 4) {
                        float t4[4];
                        ld4(mrow + c0 + g8 + h, t4);
                        ms[h] = t4[0]; ms[h + 1] = t4[1]; ms[h + 2] = t4[2]; ms[h + 3] = t4[3];
                    }
                }
            }
            if (acc) {
#pragma unroll
                for (int h = 0; h < 8; h += 4) {
                    float t4[4];
                    ld4(orow + c0 + g8 + h, t4);
                    old[h] = t4[0]; old[h + 1] = t4[1]; old[h + 2] = t4[2]; old[h + 3] = t4[3];
                }
            }
#pragma unroll
            for (int h = 0; h < 8; h += 4) {
                float scv[4] = {1.f, 1.f, 1.f, 1.f}, biv[4] = {0.f, 0.f, 0.f, 0.f};
                if (aff) {
                    const float4 sc = sc4[(c0 + g8 + h) >> 2], bi = bi4[(c0 + g8 + h) >> 2];
                    scv[0] = sc.x; scv[1] = sc.y; scv[2] = sc.z; scv[3] = sc.w;
                    biv[0] = bi.x; biv[1] = bi.y; biv[2] = bi.z; biv[3] = bi.w;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float y = __uint_as_float(raw[g8 + h + j]);
                    if (aff) y = fmaf(y, scv[j], biv[j]);
                    gd[h + j] = 0.0f;
                    if (act == ACT_RELU) y = fmaxf(y, 0.0f);
                    else if (act == ACT_LRELU) y = fmaxf(y, 0.2f * y);
                    else if (act == ACT_GELU) { float yy; gelu_fast(y, yy, gd[h + j]); y = yy; }
                    if (mul == MUL_LRELU_SIGN) y *= (ms[h + j] > 0.0f ? 1.0f : 0.2f);
                    else if (mul == MUL_RELU_SIGN) y = ms[h + j] > 0.0f ? y : 0.0f;
                    else if (mul == MUL_VALUE) y *= ms[h + j];
                    if (acc) y += old[h + j];
                    x[h + j] = y;
                }
            }
            st8(orow + c0 + g8, x);
            if (aux) st8(xrow + c0 + g8, gd);
        }
    }
}

// Drain one 128 x BN accumulator tile: TMEM -> registers -> bias/scale/activation/mask -> global.
// `wait_bar`/`wait_parity`: the MMA->epilogue barrier of this accumulator.  The mask row is prefetched BEFORE the wait.
// NC = number of accumulator columns this warp drains, starting at column c_begin (two epilogue warpgroups split a tile).
template <int BN, int NC, typename TO, typename TMSK>
__device__ __forceinline__ void drain_tile(const TcTapArgs& P, const float* s_bias, const float* s_scale, uint32_t tmem_acc,
                                           int b0, int m0, int n0, int warp, int lane, uint64_t* wait_bar,
                                           uint32_t wait_parity, int c_begin = 0) {
        const int q = warp & 3;                  // TMEM lane quarter this warp may read
        const int lane_row = q * 32 + lane;
        const int r = ((lane_row & (P.il - 1)) << P.mpt_shift) + (lane_row >> P.il_shift);   // tile row (see TcTapArgs::il)
        const int bb = b0 + r / P.mpt, mm = m0 + r % P.mpt;
        const bool row_ok = bb < P.B && !(P.dbg & 1);
        TO* __restrict__ Ob = static_cast<TO*>(P.Out);
        const TMSK* __restrict__ Mb = static_cast<const TMSK*>(P.mul_src);
        TO* __restrict__ Xb = static_cast<TO*>(P.aux);
        const long long o = (long long)bb * P.o_bstride + (long long)mm * P.o_mstride + P.o_off + n0;
        constexpr bool kPrefetchMask = sizeof(TMSK) == 2;
        constexpr int kMaskVecs = kPrefetchMask ? NC / 8 : 1;
        uint4 mreg[kMaskVecs];
        if (kPrefetchMask && P.mul_mode != MUL_NONE && row_ok && !(P.dbg & 8)) {
#pragma unroll
            for (int i = 0; i < kMaskVecs; ++i) mreg[i] = __ldg(reinterpret_cast<const uint4*>(Mb + o + c_begin) + i);
        }
        mbar_wait(wait_bar, wait_parity);
        tc_fence_after();
        const float4* sc4 = reinterpret_cast<const float4*>(s_scale);
        const float4* bi4 = reinterpret_cast<const float4*>(s_bias);
        const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
        // The drain is bound by instruction issue (a 128x128 tile is 16 K elements against as few as 4 MMAs), and a
        // predicated-off instruction still costs its issue slot: pick a loop compiled for exactly this epilogue.
        const bool affine = P.bias != nullptr || P.col_scale != nullptr || P.alpha != 1.0f;
        int variant = 0;                                  // generic
        if (!P.accumulate) {
            if (P.mul_mode == MUL_NONE && !P.aux && P.act != ACT_GELU) variant = 1 + P.act;              // 1,2,3
            else if (P.mul_mode == MUL_NONE && P.aux && P.act == ACT_GELU) variant = 4;
            else if (P.act == ACT_NONE && !P.aux && !affine && P.mul_mode != MUL_NONE) variant = 4 + P.mul_mode;  // 5,6,7
        }
#define MG_DRAIN(ACT_, MUL_, AFF_, AUX_, GEN_)                                                                        \
    drain_cols<NC, ACT_, MUL_, AFF_, AUX_, GEN_, kPrefetchMask, kMaskVecs, TO, TMSK>(P, taddr, c_begin, row_ok, Ob + o, \
                                                                                    Mb + o, Xb + o, mreg, sc4, bi4)
        switch (variant) {
            case 1: MG_DRAIN(ACT_NONE, MUL_NONE, true, false, false); break;
            case 2: MG_DRAIN(ACT_RELU, MUL_NONE, true, false, false); break;
            case 3: MG_DRAIN(ACT_LRELU, MUL_NONE, true, false, false); break;
            case 4: MG_DRAIN(ACT_GELU, MUL_NONE, true, true, false); break;
            case 5: MG_DRAIN(ACT_NONE, MUL_LRELU_SIGN, false, false, false); break;
            case 6: MG_DRAIN(ACT_NONE, MUL_RELU_SIGN, false, false, false); break;
            case 7: MG_DRAIN(ACT_NONE, MUL_VALUE, false, false, false); break;
            default: MG_DRAIN(ACT_NONE, MUL_NONE, true, true, true); break;
        }
#undef MG_DRAIN
}

// ---- drain through shared memory + TMA store (weight-stationary kernel) ----
// Row-per-thread global stores hand L2 thirty-two half-sector writes per instruction; measured on the whole training
// cycle they cost 11 of 37 ms.  Here every epilogue thread writes its 32 columns into a 128B-swizzled staging tile
// (conflict-free: 8 consecutive rows hit 8 different 16-byte pieces) and ONE thread issues cp.async.bulk.tensor stores:
// full 128-byte lines, no LSU work, asynchronous to the next tile's drain.
// Epilogue math of 8 accumulator columns of one row: y = act(acc * scale + bias) [* mask], gd = act'(.) for GELU.
// Working in chunks of 8 (one 16-byte staging store) keeps the live state at the 32 raw accumulators plus one chunk:
// holding a whole row of results and derivatives (the first form of this drain) spilled at the 96-register cap of the
// 576-thread CTA and the spills, not the arithmetic, were what the GELU layers ran at.
template <int ACT, int MUL, bool AFF, bool AUX, bool GEN, bool SCL = true>
__device__ __forceinline__ void epi_chunk8(const TcTapArgs& P, const uint32_t* raw8, const float (&ms)[8], const float4* sc4,
                                           const float4* bi4, int col, float (&y8)[8], float (&gd)[8]) {
    const int act = GEN ? P.act : ACT, mul = GEN ? P.mul_mode : MUL;
    const bool aff = GEN ? true : AFF;
#pragma unroll
    for (int h = 0; h < 8; h += 4) {
        float scv[4] = {1.f, 1.f, 1.f, 1.f}, biv[4] = {0.f, 0.f, 0.f, 0.f};
        if (aff) {
            const float4 bi = bi4[(col + h) >> 2];
            biv[0] = bi.x; biv[1] = bi.y; biv[2] = bi.z; biv[3] = bi.w;
            if (SCL) {                        // SCL = false: no folded-BN scale and alpha == 1 -- one FADD, half the LDS
                const float4 sc = sc4[(col + h) >> 2];
                scv[0] = sc.x; scv[1] = sc.y; scv[2] = sc.z; scv[3] = sc.w;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float y = __uint_as_float(raw8[h + j]);
            if (aff) y = SCL ? fmaf(y, scv[j], biv[j]) : y + biv[j];
            gd[h + j] = 0.0f;
            if (act == ACT_RELU) y = fmaxf(y, 0.0f);
            else if (act == ACT_LRELU) y = fmaxf(y, 0.2f * y);
            else if (act == ACT_GELU) { float yy; gelu_fast(y, yy, gd[h + j]); y = yy; }
            if (mul == MUL_LRELU_SIGN) y *= (ms[h + j] > 0.0f ? 1.0f : 0.2f);
            else if (mul == MUL_RELU_SIGN) y = ms[h + j] > 0.0f ? y : 0.0f;
            else if (mul == MUL_VALUE) y *= ms[h + j];
            y8[h + j] = y;
        }
    }
}

// staging tile: boxes of [128 rows][128 bytes], piece p of row r at (p ^ (r & 7)) * 16
__device__ __forceinline__ uint4 pack8(const float (&x)[8]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(x[0], x[1]), b = __floats2bfloat162_rn(x[2], x[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(x[4], x[5]), d = __floats2bfloat162_rn(x[6], x[7]);
    return make_uint4(*reinterpret_cast<unsigned*>(&a), *reinterpret_cast<unsigned*>(&b), *reinterpret_cast<unsigned*>(&c),
                      *reinterpret_cast<unsigned*>(&d));
}
__device__ __forceinline__ void stage8(unsigned char* staging, int r, int cc, const float (&x)[8], __nv_bfloat16) {
    const int box = cc >> 6, pidx = (cc & 63) >> 3;
    *reinterpret_cast<uint4*>(staging + box * 16384 + r * 128 + ((pidx ^ (r & 7)) << 4)) = pack8(x);
}
__device__ __forceinline__ void stage8(unsigned char* staging, int r, int cc, const float (&x)[8], float) {
#pragma unroll
    for (int g4 = 0; g4 < 8; g4 += 4) {
        const int c = cc + g4, box = c >> 5, pidx = (c & 31) >> 2;
        *reinterpret_cast<uint4*>(staging + box * 16384 + r * 128 + ((pidx ^ (r & 7)) << 4)) =
            make_uint4(__float_as_uint(x[g4]), __float_as_uint(x[g4 + 1]), __float_as_uint(x[g4 + 2]), __float_as_uint(x[g4 + 3]));
    }
}

// One epilogue thread = one accumulator row x CPT (32 or 16) columns: math in chunks of 8, results (and the GELU derivative tile)
// straight into the staging tile(s).
template <int ACT, int MUL, bool AFF, bool AUX, bool GEN, typename TO, typename TMSK, bool SCL = true, int CPT = 32>
__device__ __forceinline__ void epi_row32(const TcTapArgs& P, const uint32_t (&raw)[CPT], const TMSK* __restrict__ mrow, bool row_ok,
                                          const unsigned char* maskrow, const uint4 (&mreg)[CPT / 8], const float4* sc4,
                                          const float4* bi4, int c_begin, int r, unsigned char* stage_out,
                                          unsigned char* stage_aux) {
    const int mul = GEN ? P.mul_mode : MUL;
    const bool aux = GEN ? (P.aux != nullptr) : AUX;
#pragma unroll
    for (int g8 = 0; g8 < CPT; g8 += 8) {
        float ms[8], y8[8], gd[8];
        if (mul != MUL_NONE) {
            if (sizeof(TMSK) == 2) {
                uint4 mv = mreg[g8 >> 3];      // prefetched from global memory before the accumulator wait ...
                if (maskrow) {                 // ... or read from the TMA-loaded mask tile (same swizzle as the staging tile)
                    const int cc = c_begin + g8, box = cc >> 6, pidx = (cc & 63) >> 3;
                    mv = *reinterpret_cast<const uint4*>(maskrow + box * 16384 + ((pidx ^ (r & 7)) << 4));
                }
                const uint32_t w[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    ms[2 * e] = __uint_as_float(w[e] << 16);
                    ms[2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u);
                }
            } else {
#pragma unroll
                for (int h = 0; h < 8; h += 4) {
                    float t4[4] = {0.f, 0.f, 0.f, 0.f};
                    if (row_ok) ld4(mrow + c_begin + g8 + h, t4);
                    ms[h] = t4[0]; ms[h + 1] = t4[1]; ms[h + 2] = t4[2]; ms[h + 3] = t4[3];
                }
            }
        }
        epi_chunk8<ACT, MUL, AFF, AUX, GEN, SCL>(P, &raw[g8], ms, sc4, bi4, c_begin + g8, y8, gd);
        stage8(stage_out, r, c_begin + g8, y8, TO());
        if (aux) stage8(stage_aux, r, c_begin + g8, gd, TO());
    }
}

// Fused AdaptiveAvgPool1d(1) of the weight-stationary kernels (bf16 outputs, 128-wide slabs, 1 or 2 samples per tile): the
// finished staging tile [128 rows][128 columns] is summed column-wise by the 512 epilogue threads -- thread = (column pair,
// group of 16 rows), conflict-free 4-byte reads of the swizzled rows -- into a double-buffered [8 row groups][128] float
// buffer; the partials of tile t are combined and written out after the barrier of tile t + 1 (no extra barrier per tile).
// The sums are sums of the STORED bf16 values, in float32, like pool_rows_kernel's.
template <int BN>
__device__ __forceinline__ void ws_pool_partial(const unsigned char* stage_out, float* poolbuf, int buf, int et) {
    constexpr int CP = BN / 2, RG = 512 / CP, ROWS = 128 / RG;     // column pairs, row groups, rows per group (512 threads)
    const int cp = et % CP, rg = et / CP;
    const unsigned char* col = stage_out + (cp >> 5) * 16384 + (cp & 3) * 4;
    const int pidx = (cp & 31) >> 2;
    float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
        const int r = rg * ROWS + rr;
        const uint32_t w = *reinterpret_cast<const uint32_t*>(col + r * 128 + ((pidx ^ (r & 7)) << 4));
        s0 += __uint_as_float(w << 16);
        s1 += __uint_as_float(w & 0xFFFF0000u);
    }
    *reinterpret_cast<float2*>(poolbuf + (buf * RG + rg) * BN + 2 * cp) = make_float2(s0, s1);
}
// BatchNorm batch statistics of a float32 staging tile [BN / 32 boxes][128 rows][32 floats]: thread = (column, group of rows),
// sum and sum of squares; partials [buf][row group][2][BN]
template <int BN>
__device__ __forceinline__ void ws_stats_partial(const unsigned char* stage_out, float* poolbuf, int buf, int et) {
    constexpr int RG = 512 / BN, ROWS = 128 / RG;
    const int col = et % BN, rg = et / BN;
    const unsigned char* p = stage_out + (col >> 5) * 16384 + (col & 3) * 4;
    const int pidx = (col & 31) >> 2;
    float s = 0.0f, q = 0.0f;
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
        const int r = rg * ROWS + rr;
        const float v = *reinterpret_cast<const float*>(p + r * 128 + ((pidx ^ (r & 7)) << 4));
        s += v;
        q = fmaf(v, v, q);
    }
    poolbuf[(buf * RG + rg) * 2 * BN + col] = s;
    poolbuf[(buf * RG + rg) * 2 * BN + BN + col] = q;
}
template <int BN>
__device__ __forceinline__ float ws_stats_take(const float* poolbuf, int buf, int et) {      // et < 2 * BN: (statistic, column)
    constexpr int RG = 512 / BN;
    float s = 0.0f;
#pragma unroll
    for (int g = 0; g < RG; ++g) s += poolbuf[(buf * RG + g) * 2 * BN + et];
    return s;
}
// column sums over ALL rows (bias gradients): each thread et < BN keeps a running sum of its column over this CTA's tiles
template <int BN>
__device__ __forceinline__ float ws_colsum_take(const float* poolbuf, int buf, int et) {
    constexpr int RG = 512 / (BN / 2);
    float s = 0.0f;
#pragma unroll
    for (int g = 0; g < RG; ++g) s += poolbuf[(buf * RG + g) * BN + et];
    return s;
}
__device__ __forceinline__ void ws_pool_flush(const TcTapArgs& P, const float* poolbuf, int buf, int row0, int n0, int et) {
    const int j = et >> 7, col = et & 127;                    // sample inside the tile, column of the slab
    if (j >= P.bpt) return;
    const int per = 8 / P.bpt;                                // row groups per sample
    float s = 0.0f;
    for (int g = j * per; g < (j + 1) * per; ++g) s += poolbuf[(buf * 8 + g) * 128 + col];
    const int b = (row0 >> P.mper_shift) + j;
    if (b >= P.B) return;
    float* dst = P.pool_out + (long long)b * P.N + n0 + col;
    if (P.pool_atomic) atomicAdd(dst, s * P.pool_scale);
    else *dst = s * P.pool_scale;
}

// The whole epilogue of the weight-stationary kernel for one compile-time epilogue variant: the tile loop lives INSIDE
// the variant (ncu, round 2: with the variant switch, three integer divisions for the tile coordinates and two
// 512-thread barriers per tile the drain executed ~470 instructions per warp and tile of which ~150 were the epilogue
// math; issue-bound at 4 warps per scheduler, it -- not HBM, not the tensor pipe -- set the pace of every layer).
//   * tile coordinates by shift/mask (Mper and mpt are powers of two), everything tile-invariant hoisted;
//   * TMEM hand-back by one mbarrier arrive per WARP right after its tcgen05.ld (no CTA-wide barrier);
//   * ONE named barrier per tile (staging tile complete -> TMA store); the staging slot is recycled by waiting, before
//     that barrier, for the bulk store issued one tile earlier (it has had a whole tile of math to drain).
template <int BN, int kEpi, int CPT, int ACT, int MUL, bool AFF, bool AUX, bool GEN, bool SCL, typename TO, typename TMSK, typename HDR,
          bool PAIR = false>
__device__ __forceinline__ void ws_epilogue_loop(const TcTapArgs& P, const CUtensorMap* o_map, const CUtensorMap* x_map,
                                                 const CUtensorMap* m_map, HDR& H, uint32_t tmem0, int n0, int mtiles,
                                                 unsigned char* staging0, unsigned char* maskbuf, float* poolbuf = nullptr) {
    constexpr int EPB = 128 / (int)sizeof(TO);       // elements per staging box row
    constexpr size_t kStageTile = (size_t)128 * BN * sizeof(TO);
    const int et = threadIdx.x - 64, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, lane_row = q * 32 + lane, c_begin = (et >> 7) * CPT;
    const int r = ((lane_row & (P.il - 1)) << P.mpt_shift) + (lane_row >> P.il_shift);   // tile row this TMEM lane holds
    const bool has_aux = GEN ? (P.aux != nullptr) : AUX;
    const int per_tile = has_aux ? 2 : 1;
    const bool ring2 = P.nsb >= 2 * per_tile;        // a free staging slot while the previous tile's store drains
    const int mul = GEN ? P.mul_mode : MUL;
    const bool tma_mask = mul != MUL_NONE && P.tma_mask;
    const bool reg_mask = mul != MUL_NONE && !P.tma_mask;
    const int rdiv = r >> P.mpt_shift, rmod = r & (P.mpt - 1);
    const TMSK* __restrict__ Mb = static_cast<const TMSK*>(P.mul_src);
    const float4* sc4 = reinterpret_cast<const float4*>(H.scale);
    const float4* bi4 = reinterpret_cast<const float4*>(H.bias);
    const uint32_t tmem_lane = tmem0 + ((uint32_t)(q * 32) << 16) + (uint32_t)c_begin;
    constexpr uint32_t kMaskTile = (uint32_t)(BN / 64) * 16384u;
    const int gstep = (int)gridDim.x;
    int sbuf = 0, tcount = 0, pool_row0 = 0;
    float colsum_acc = 0.0f;
    double stats_acc = 0.0;
    auto load_mask = [&](int tile, int buf) {        // one thread: the mask tile of `tile` -> ring slot buf
        const int mrow0 = ws_row_tile(P, tile, mtiles) * 128;
        mbar_expect_tx(&H.mask_full[buf], kMaskTile);
#pragma unroll
        for (int bx = 0; bx < BN / 64; ++bx)
            tma_load_2d(m_map, &H.mask_full[buf], maskbuf + (size_t)buf * kMaskTile + bx * 16384, n0 + bx * 64, mrow0);
    };
    if (tma_mask && et == 0) {
        for (int k = 0; k < P.nmb; ++k)
            if ((int)blockIdx.x + k * gstep < mtiles) load_mask((int)blockIdx.x + k * gstep, k);
    }
    int mbuf = 0;
    uint32_t mpar = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < mtiles; tile += gstep, ++tcount) {
        const int acc = tcount & 1;
        const int row0 = ws_row_tile(P, tile, mtiles) * 128;
        const int bb = (row0 >> P.mper_shift) + rdiv;
        const bool row_ok = bb < P.B;
        uint4 mreg[CPT / 8] = {};
        const TMSK* mrow = nullptr;
        if (reg_mask) {                                          // per-thread mask loads, in flight during the accumulator wait
            const long long o = (long long)bb * P.o_bstride + (long long)((row0 & (P.Mper - 1)) + rmod) * P.o_mstride + P.o_off + n0;
            mrow = Mb + o;
            if (sizeof(TMSK) == 2 && row_ok && !(P.dbg & 8)) {
#pragma unroll
                for (int i = 0; i < CPT / 8; ++i) mreg[i] = __ldg(reinterpret_cast<const uint4*>(mrow + c_begin) + i);
            }
        }
        mbar_wait(&H.tmem_full[acc], (uint32_t)(tcount >> 1) & 1u);
        tc_fence_after();
        uint32_t raw[CPT];
        tmem_ld_async(tmem_lane + (uint32_t)(acc * BN), raw);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {                                         // this warp's lanes of the accumulator are free again
            if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&H.tmem_empty[acc]), 0));   // the pair's MMA warp lives in CTA 0
            else mbar_arrive(&H.tmem_empty[acc]);
        }
        if (!ring2) {                                            // single staging slot: the previous store must have left it
            if (et == 0) bulk_wait_read0();
            asm volatile("bar.sync 1, %0;" ::"n"(kEpi) : "memory");
        }
        unsigned char* stage_out = staging0 + (size_t)sbuf * kStageTile;
        unsigned char* stage_aux = staging0 + (size_t)(sbuf + 1) * kStageTile;
        const unsigned char* maskrow = nullptr;
        if (tma_mask) {
            mbar_wait(&H.mask_full[mbuf], mpar);
            maskrow = maskbuf + (size_t)mbuf * kMaskTile + r * 128;
        }
        epi_row32<ACT, MUL, AFF, AUX, GEN, TO, TMSK, SCL, CPT>(P, raw, mrow, row_ok && !(P.dbg & 8), maskrow, mreg, sc4, bi4, c_begin,
                                                               r, stage_out, stage_aux);
        fence_proxy_async();
        if (ring2 && et == 0) bulk_wait_read0();                 // the store issued one tile ago has drained its slot
        asm volatile("bar.sync 1, %0;" ::"n"(kEpi) : "memory");
        if (et == 0) {
            if (!(P.dbg & 1)) {
                if (!P.skip_out) {
#pragma unroll
                    for (int bx = 0; bx < BN / EPB; ++bx) tma_store_2d(o_map, stage_out + bx * 16384, n0 + bx * EPB, row0);
                }
                if (has_aux) {
#pragma unroll
                    for (int bx = 0; bx < BN / EPB; ++bx) tma_store_2d(x_map, stage_aux + bx * 16384, n0 + bx * EPB, row0);
                }
                bulk_commit();
            }
            // every thread has read this tile's mask: its ring slot takes the tile nmb steps ahead
            if (tma_mask && tile + P.nmb * gstep < mtiles) load_mask(tile + P.nmb * gstep, mbuf);
        }
        if (sizeof(TO) == 4 && poolbuf && P.stats_out) {         // fused BatchNorm statistics (float32 staging tile)
            if (tcount > 0 && et < 2 * BN) stats_acc += (double)ws_stats_take<BN>(poolbuf, (tcount - 1) & 1, et);
            ws_stats_partial<BN>(stage_out, poolbuf, tcount & 1, et);
        }
        if (sizeof(TO) == 2 && poolbuf) {
            if (BN == 128 && P.pool_out) {                       // fused mean over the rows of a sample (ws_pool_*)
                if (tcount > 0) ws_pool_flush(P, poolbuf, (tcount - 1) & 1, pool_row0, n0, et);
                ws_pool_partial<BN>(stage_out, poolbuf, tcount & 1, et);
                pool_row0 = row0;
            } else if (P.colsum_out) {                           // fused column sums over the first colsum_tiles row tiles
                if (tcount > 0 && pool_row0 && et < BN) colsum_acc += ws_colsum_take<BN>(poolbuf, (tcount - 1) & 1, et);
                pool_row0 = (row0 >> 7) < P.colsum_tiles;        // (reused as "the previous tile contributed")
                if (pool_row0) ws_pool_partial<BN>(stage_out, poolbuf, tcount & 1, et);
            }
        }
        if (++mbuf >= P.nmb) { mbuf = 0; mpar ^= 1u; }
        sbuf += per_tile;
        if (sbuf >= P.nsb) sbuf = 0;
    }
    if (sizeof(TO) == 4 && poolbuf && P.stats_out && tcount > 0) {
        asm volatile("bar.sync 1, %0;" ::"n"(kEpi) : "memory");
        if (et < 2 * BN) {
            stats_acc += (double)ws_stats_take<BN>(poolbuf, (tcount - 1) & 1, et);
            atomicAdd(P.stats_out + (et / BN) * P.N + n0 + (et % BN), (float)stats_acc);
        }
    }
    if (sizeof(TO) == 2 && poolbuf && tcount > 0) {              // the last tile's partials
        asm volatile("bar.sync 1, %0;" ::"n"(kEpi) : "memory");
        if (BN == 128 && P.pool_out) ws_pool_flush(P, poolbuf, (tcount - 1) & 1, pool_row0, n0, et);
        else if (P.colsum_out && et < BN) {
            if (pool_row0) colsum_acc += ws_colsum_take<BN>(poolbuf, (tcount - 1) & 1, et);
            atomicAdd(P.colsum_out + n0 + et, colsum_acc);
        }
    }
    if (et == 0) bulk_wait0();                                   // all bulk stores complete before the CTA retires
}

template <int BN>
struct TapSmem {
    __nv_bfloat16 a[kStages][kTileM * kTileK];
    __nv_bfloat16 b[kStages][BN * kTileK];
    uint64_t full[kStages], empty[kStages], tmem_full;
    uint32_t tmem_base;
    alignas(16) float bias[BN], scale[BN];          // this tile's bias (permuted) and alpha * folded-BN scale
};

template <int BN, typename TO, typename TMSK, bool TF32 = false>
__global__ void __launch_bounds__(192) tc_tapgemm_kernel(const __grid_constant__ CUtensorMap a_map,
                                                         const __grid_constant__ CUtensorMap b_map, const TcTapArgs P) {
    extern __shared__ unsigned char smem_raw[];
    TapSmem<BN>& S = *reinterpret_cast<TapSmem<BN>*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
    constexpr uint32_t kStageBytes = (kTileM + BN) * kTileK * 2;

    // tile coordinates
    int b0, m0;
    if (P.mpt == kTileM) {
        const int tiles_m = P.Mper / kTileM;
        b0 = blockIdx.x / tiles_m;
        m0 = (blockIdx.x % tiles_m) * kTileM;
    } else {
        b0 = blockIdx.x * P.bpt;
        m0 = 0;
    }
    const int n0 = blockIdx.y * BN;
    const int nkb = P.ntaps * P.kblocks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
        mbar_init(&S.tmem_full, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc(&S.tmem_base, kTmemCols);
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&a_map); tma_prefetch_desc(&b_map); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = S.tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages, it = kb / kStages;
                mbar_wait(&S.empty[s], (it & 1) ^ 1);
                const int t = P.kb_outer ? kb % P.ntaps : kb / P.kblocks, kc = P.kb_outer ? kb / P.ntaps : kb - t * P.kblocks;
                mbar_expect_tx(&S.full[s], kStageBytes);
                tma_load_4d(&a_map, &S.full[s], S.a[s], kc * P.ktile, P.a_p[t], m0 + P.a_dm[t], b0);
                tma_load_2d(&b_map, &S.full[s], S.b[s], kc * P.ktile, P.b_row[t] + n0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = make_idesc(BN, 0, 0, TF32 ? 2u : 1u);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kStages, it = kb / kStages;
            mbar_wait(&S.full[s], it & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_addr = smem_u32(S.a[s]), b_addr = smem_u32(S.b[s]);
#pragma unroll
                for (int k = 0; k < kTileK / 16; ++k) {
                    const uint64_t ad = make_smem_desc(a_addr + k * 32, 16, 1024);
                    const uint64_t bd = make_smem_desc(b_addr + k * 32, 16, 1024);
                    umma<TF32>(tmem_acc, ad, bd, idesc, (kb | k) ? 1u : 0u);
                }
                umma_commit(&S.empty[s]);                       // frees the smem stage when these MMAs retire
                if (kb == nkb - 1) umma_commit(&S.tmem_full);   // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: TMEM -> registers -> global =====
        // While the main loop runs: stage this tile's bias / BN scale in shared memory and prefetch this thread's
        // row of the mask tensor (f'(saved activation)) into registers, so that nothing in the drain loop waits
        // on a dependent global load.
        const int et = threadIdx.x - 64;             // 0..127
        for (int i = et; i < BN; i += 128) {
            S.bias[i] = P.bias ? __ldg(P.bias + perm_index(n0 + i, P.n_perm_q, P.n_perm_p)) : 0.0f;
            S.scale[i] = (P.col_scale ? __ldg(P.col_scale + n0 + i) : 1.0f) * P.alpha;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");    // bias/scale staged (epilogue warps only)
        drain_tile<BN, BN, TO, TMSK>(P, S.bias, S.scale, tmem_acc, b0, m0, n0, warp, lane, &S.tmem_full, 0);
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_acc, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------
// weight-stationary persistent tap-GEMM
// ---------------------------------------------------------------------------------------------
// The conv layers here are skinny (C = 64..256): a 128 x BN output tile re-fetching its weight tile per k-block moves
// as many bytes from L2 as its activation tile and the kernel is L2-bound at ~10 % tensor utilisation.  In this form
// a CTA owns one BN-wide slab of output channels for its whole life: the packed weights of ALL taps for that slab are
// TMA-loaded into shared memory once (<= 192 KB), the CTA then walks M tiles (grid-strided), streaming only activation
// tiles through a small ring, with two TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
constexpr int kWsMaxStages = 8;
constexpr int kWsMaxLoads = 80;      // taps x k-blocks per tile
constexpr int kWsHeaderBytes = 4096;
// TMA warp, MMA warp, then BN / 32 epilogue warpgroups of 4 warps; each warpgroup drains 32 accumulator columns.  The
// drain is a chain of dependent latencies (tcgen05.ld -> math -> store), so it is hidden by warps, not by ILP.
// 16 epilogue warps for both slab widths: 32 columns per thread at BN = 128, 16 at BN = 64 (eight warps left the GELU /
// mask epilogues of the 64-wide slabs latency-bound at two warps per scheduler)
template <int BN> struct WsCfg {
    static constexpr int kCpt = BN >= 128 ? 32 : 16, kGroups = BN / kCpt, kEpiThreads = kGroups * 128, kThreads = 64 + kEpiThreads;
};
template <int BN>
struct WsHeader {
    uint64_t full[kWsMaxStages], empty[kWsMaxStages], wfull, tmem_full[2], tmem_empty[2], mask_full[2];
    uint32_t tmem_base;
    alignas(16) float bias[BN], scale[BN];
    // per-tile schedules, built once per CTA so that the single-thread producer / MMA loops carry no index arithmetic:
    // ldtab[j]  = (k coordinate, plane, row shift, -) of activation load j;  mmatab[i] = (byte offset of the tap's first row
    // inside the stage, byte offset of its weight block, 1 if last MMA block of its stage, -)
    alignas(16) int4 ldtab[kWsMaxLoads];
    alignas(16) int4 mmatab[kWsMaxLoads];
};

template <int BN, typename TO, typename TMSK, bool TF32 = false>
__global__ void __launch_bounds__(WsCfg<BN>::kThreads) tc_tapgemm_ws_kernel(const __grid_constant__ CUtensorMap a_map,
                                                            const __grid_constant__ CUtensorMap b_map,
                                                            const __grid_constant__ CUtensorMap o_map,
                                                            const __grid_constant__ CUtensorMap x_map,
                                                            const __grid_constant__ CUtensorMap m_map, const TcTapArgs P,
                                                            int mtiles, int nstages) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    WsHeader<BN>& H = *reinterpret_cast<WsHeader<BN>*>(base);
    const int nkb = P.ntaps * P.kblocks;
    static_assert(sizeof(WsHeader<BN>) <= kWsHeaderBytes, "header does not fit its reservation");
    __nv_bfloat16* wsm = reinterpret_cast<__nv_bfloat16*>(base + kWsHeaderBytes);             // [nkb][BN][64]
    unsigned char* asm_ = reinterpret_cast<unsigned char*>(wsm + (size_t)nkb * BN * kTileK);  // [nstages][128 + halo][64]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = 2 * BN;
    constexpr uint32_t kWBytes = BN * kTileK * 2;
    const uint32_t a_tx = (uint32_t)(kTileM + P.halo * P.il) * kTileK * 2;                    // bytes one A box delivers
    const uint32_t a_stage = (a_tx + 1023u) & ~1023u;
    unsigned char* staging = asm_ + (size_t)nstages * a_stage;                                 // [BN*sizeof(TO)/128][128][128 B]
    unsigned char* maskbuf = staging + (size_t)P.nsb * 128 * BN * sizeof(TO);                  // [BN/64][128][128 B] (bf16 masks)
    const int n0 = blockIdx.y * BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < nstages; ++s) { mbar_init(&H.full[s], 1); mbar_init(&H.empty[s], 1); }
        mbar_init(&H.wfull, 1);
        // TMA-store drain: every epilogue WARP hands its TMEM lanes back; row-per-thread drain: one arrive after a barrier
        for (int a = 0; a < 2; ++a) {
            mbar_init(&H.tmem_full[a], 1);
            mbar_init(&H.tmem_empty[a], P.tma_store ? (uint32_t)(WsCfg<BN>::kEpiThreads / 32) : 1u);
        }
        mbar_init(&H.mask_full[0], 1); mbar_init(&H.mask_full[1], 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc(&H.tmem_base, kTmemCols);
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&a_map); tma_prefetch_desc(&b_map); }
    const int npt = P.ngroups * P.kblocks;            // activation loads (= stages consumed) per tile
    if (threadIdx.x == 64) {
        int i = 0;
        for (int g = 0; g < P.ngroups; ++g)
            for (int kc = 0; kc < P.kblocks; ++kc) {
                H.ldtab[g * P.kblocks + kc] = make_int4(kc * P.ktile, P.g_p[g], P.g_dmin[g], 0);
                for (int j = 0; j < P.g_count[g]; ++j, ++i) {
                    const int t = P.g_tap[P.g_first[g] + j];
                    H.mmatab[i] = make_int4((P.a_dm[t] - P.g_dmin[g]) * P.il * 128, (t * P.kblocks + kc) * (int)kWBytes,
                                            j == P.g_count[g] - 1, 0);
                }
            }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = H.tmem_base;

    auto tile_coords = [&](int tile, int& b0, int& m0) {      // tile -> first sample, first row inside it (powers of two)
        tile = ws_row_tile(P, tile, mtiles);
        const int row0 = tile * kTileM;
        b0 = row0 >> P.mper_shift;
        m0 = row0 & (P.Mper - 1);
    };

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(&H.wfull, (uint32_t)nkb * kWBytes);
            for (int kb = 0; kb < nkb; ++kb) {
                const int t = kb / P.kblocks, kc = kb - t * P.kblocks;
                tma_load_2d(&b_map, &H.wfull, wsm + (size_t)kb * BN * kTileK, kc * P.ktile, P.b_row[t] + n0);
            }
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < mtiles; tile += gridDim.x) {
                int b0, m0;
                tile_coords(tile, b0, m0);
                for (int j = 0; j < npt; ++j) {
                    mbar_wait(&H.empty[s], ph ^ 1);
                    const int4 L = H.ldtab[j];
                    if (P.dbg & 4) { mbar_expect_tx(&H.full[s], 0); }
                    else {
                        mbar_expect_tx(&H.full[s], a_tx);
                        if (P.il > 1) tma_load_4d(&a_map, &H.full[s], asm_ + (size_t)s * a_stage, L.x, b0, L.y, m0 + L.z);
                        else tma_load_4d(&a_map, &H.full[s], asm_ + (size_t)s * a_stage, L.x, L.y, m0 + L.z, b0);
                    }
                    if (++s == nstages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(BN, 0, 0, TF32 ? 2u : 1u);
        mbar_wait(&H.wfull, 0);
        int s = 0, tcount = 0;
        uint32_t ph = 0;
        const uint64_t dbase = make_smem_desc(0, 16, 1024);
        const uint32_t a_base = smem_u32(asm_), w_base = smem_u32(wsm);
        for (int tile = blockIdx.x; tile < mtiles; tile += gridDim.x, ++tcount) {
            const int acc = tcount & 1;
            mbar_wait(&H.tmem_empty[acc], ((tcount >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tacc = tmem0 + (uint32_t)(acc * BN);
            uint32_t started = 0;
            int i = 0;
            for (int j = 0; j < npt; ++j) {
                mbar_wait(&H.full[s], ph);
                tc_fence_after();
                const uint32_t stage_addr = a_base + (uint32_t)s * a_stage;
                int last;
                do {                                           // warp-uniform: every lane walks the schedule
                    const int4 M = H.mmatab[i++];
                    const uint64_t ad = dbase | (uint64_t)(((stage_addr + (uint32_t)M.x) >> 4) & 0x3FFF);
                    const uint64_t bd = dbase | (uint64_t)(((w_base + (uint32_t)M.y) >> 4) & 0x3FFF);
                    if (!(P.dbg & 2) && elect_one()) {
#pragma unroll
                        for (int k = 0; k < kTileK / 16; ++k) umma<TF32>(tacc, ad + 2u * k, bd + 2u * k, idesc, k ? 1u : started);
                    }
                    started = 1;
                    last = M.z;
                } while (!last);
                if (elect_one()) {
                    umma_commit(&H.empty[s]);
                    if (j == npt - 1) umma_commit(&H.tmem_full[acc]);
                }
                __syncwarp();
                if (++s == nstages) { s = 0; ph ^= 1; }
            }
        }
    } else {
        const int et = threadIdx.x - 64;                 // 0..255
        constexpr int kEpi = WsCfg<BN>::kEpiThreads;
        const int grp = et >> 7;                         // which 32 columns of the slab this warpgroup drains
        for (int i = et; i < BN; i += kEpi) {
            H.bias[i] = P.bias ? __ldg(P.bias + perm_index(n0 + i, P.n_perm_q, P.n_perm_p)) : 0.0f;
            H.scale[i] = (P.col_scale ? __ldg(P.col_scale + n0 + i) : 1.0f) * P.alpha;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpi) : "memory");
        if (P.tma_store) {
            // one tile loop per compile-time epilogue variant (see ws_epilogue_loop)
            const bool scaled = P.col_scale != nullptr || P.alpha != 1.0f;
            const bool affine = P.bias != nullptr || scaled;
            int variant = 0;
            if (P.mul_mode == MUL_NONE && !P.aux && P.act != ACT_GELU) variant = (scaled ? 1 : 8) + P.act;     // 1..3 / 8..10
            else if (P.mul_mode == MUL_NONE && P.aux && P.act == ACT_GELU) variant = 4;
            else if (P.act == ACT_NONE && !P.aux && !affine && P.mul_mode != MUL_NONE) variant = 4 + P.mul_mode;
            float* poolbuf = (P.pool_out || P.colsum_out || P.stats_out) ? reinterpret_cast<float*>(maskbuf + (P.tma_mask ? (size_t)P.nmb * (BN / 64) * 16384 : 0)) : nullptr;
#define MG_LOOP(ACT_, MUL_, AFF_, AUX_, GEN_, SCL_)                                                                  \
    ws_epilogue_loop<BN, kEpi, WsCfg<BN>::kCpt, ACT_, MUL_, AFF_, AUX_, GEN_, SCL_, TO, TMSK>(P, &o_map, &x_map, &m_map, H, tmem0, n0, mtiles, \
                                                                             staging, maskbuf, poolbuf)
            switch (variant) {
                case 1: MG_LOOP(ACT_NONE, MUL_NONE, true, false, false, true); break;
                case 2: MG_LOOP(ACT_RELU, MUL_NONE, true, false, false, true); break;
                case 3: MG_LOOP(ACT_LRELU, MUL_NONE, true, false, false, true); break;
                case 4: MG_LOOP(ACT_GELU, MUL_NONE, true, true, false, true); break;
                case 5: MG_LOOP(ACT_NONE, MUL_LRELU_SIGN, false, false, false, true); break;
                case 6: MG_LOOP(ACT_NONE, MUL_RELU_SIGN, false, false, false, true); break;
                case 7: MG_LOOP(ACT_NONE, MUL_VALUE, false, false, false, true); break;
                case 8: MG_LOOP(ACT_NONE, MUL_NONE, true, false, false, false); break;
                case 9: MG_LOOP(ACT_RELU, MUL_NONE, true, false, false, false); break;
                case 10: MG_LOOP(ACT_LRELU, MUL_NONE, true, false, false, false); break;
                default: MG_LOOP(ACT_NONE, MUL_NONE, true, true, true, true); break;
            }
#undef MG_LOOP
        } else {
            int tcount = 0;
            for (int tile = blockIdx.x; tile < mtiles; tile += gridDim.x, ++tcount) {
                const int acc = tcount & 1;
                int b0, m0;
                tile_coords(tile, b0, m0);
                if (grp * 32 < BN)      // row-per-thread drain: 32 columns per warpgroup (the other warpgroups only keep the barrier)
                    drain_tile<BN, 32, TO, TMSK>(P, H.bias, H.scale, tmem0 + (uint32_t)(acc * BN), b0, m0, n0, warp, lane,
                                                 &H.tmem_full[acc], (tcount >> 1) & 1, grp * 32);
                tc_fence_before();
                asm volatile("bar.sync 1, %0;" ::"n"(kEpi) : "memory");   // every epilogue thread has read its TMEM lanes
                if (et == 0) mbar_arrive(&H.tmem_empty[acc]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem0, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------
// weight-stationary persistent tap-GEMM on CTA PAIRS (tcgen05 cta_group::2)
// ---------------------------------------------------------------------------------------------
// Two CTAs of a cluster (the two SMs of a TPC) work on adjacent 128-row tiles with ONE tcgen05.mma of M = 256, N = 128: each
// CTA streams its own activation tiles, keeps only HALF of the slab's weight columns (64) resident and drains its own 128
// accumulator lanes.  Against the one-CTA kernel this halves (a) the resident weight bytes per SM -- 256-channel layers whose
// 128-wide slab did not fit next to a pipeline (K = 3 x 256, 5 x 128, 2 x 256) keep BN = 128, deeper activation rings, double
// staging and TMA mask tiles -- and (b) the shared-memory operand bytes the tensor core reads per FLOP (A 4 KB + B 2 KB per
// CTA and 64 clocks instead of A 4 KB + B 4 KB).  Roles as in tc_tapgemm_ws_kernel; differences: only CTA 0's warp 1 issues
// MMAs; every `full` / `wfull` barrier that gates them lives in CTA 0 and counts the bytes of BOTH CTAs' TMA loads (the
// peer's loads name CTA 0's barrier, cta_group::2 form); tcgen05.commit multicasts its arrive to the barrier at the same
// offset in both CTAs (each producer's `empty`, each epilogue's `tmem_full`); both epilogues hand their TMEM lanes back
// to CTA 0's `tmem_empty`.  Epilogue: the TMA-store staging form only (ws_epilogue_loop, PAIR = true).
template <typename TO, typename TMSK>
__global__ void __launch_bounds__(WsCfg<128>::kThreads) tc_tapgemm_ws2_kernel(const __grid_constant__ CUtensorMap a_map,
                                                            const __grid_constant__ CUtensorMap b_map,
                                                            const __grid_constant__ CUtensorMap o_map,
                                                            const __grid_constant__ CUtensorMap x_map,
                                                            const __grid_constant__ CUtensorMap m_map, const TcTapArgs P,
                                                            int mtiles, int nstages) {
    constexpr int BN = 128, BH = 64;                  // the pair's slab width, this CTA's resident half
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    WsHeader<BN>& H = *reinterpret_cast<WsHeader<BN>*>(base);
    const int nkb = P.ntaps * P.kblocks;
    __nv_bfloat16* wsm = reinterpret_cast<__nv_bfloat16*>(base + kWsHeaderBytes);             // [nkb][BH][64]
    unsigned char* asm_ = reinterpret_cast<unsigned char*>(wsm + (size_t)nkb * BH * kTileK);  // [nstages][128 + halo][64]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = 2 * BN;
    constexpr uint32_t kWBytes = BH * kTileK * 2;
    const uint32_t a_tx = (uint32_t)(kTileM + P.halo * P.il) * kTileK * 2;
    const uint32_t a_stage = (a_tx + 1023u) & ~1023u;
    unsigned char* staging = asm_ + (size_t)nstages * a_stage;
    unsigned char* maskbuf = staging + (size_t)P.nsb * 128 * BN * sizeof(TO);
    const int n0 = blockIdx.y * BN;
    const uint32_t rank = cluster_ctarank();

    if (threadIdx.x == 0) {
        for (int s = 0; s < nstages; ++s) { mbar_init(&H.full[s], 1); mbar_init(&H.empty[s], 1); }
        mbar_init(&H.wfull, 1);
        for (int a = 0; a < 2; ++a) {
            mbar_init(&H.tmem_full[a], 1);
            mbar_init(&H.tmem_empty[a], 2u * (uint32_t)(WsCfg<BN>::kEpiThreads / 32));    // every epilogue warp of both CTAs
        }
        mbar_init(&H.mask_full[0], 1); mbar_init(&H.mask_full[1], 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc_pair(&H.tmem_base, kTmemCols);
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&a_map); tma_prefetch_desc(&b_map); }
    const int npt = P.ngroups * P.kblocks;
    if (threadIdx.x == 64) {
        int i = 0;
        for (int g = 0; g < P.ngroups; ++g)
            for (int kc = 0; kc < P.kblocks; ++kc) {
                H.ldtab[g * P.kblocks + kc] = make_int4(kc * P.ktile, P.g_p[g], P.g_dmin[g], 0);
                for (int j = 0; j < P.g_count[g]; ++j, ++i) {
                    const int t = P.g_tap[P.g_first[g] + j];
                    H.mmatab[i] = make_int4((P.a_dm[t] - P.g_dmin[g]) * P.il * 128, (t * P.kblocks + kc) * (int)kWBytes,
                                            j == P.g_count[g] - 1, 0);
                }
            }
    }
    tc_fence_before();
    cluster_sync_all();                               // barriers of BOTH CTAs initialised before any remote arrive / TMA
    tc_fence_after();
    const uint32_t tmem0 = H.tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t wfull0 = mapa_u32(smem_u32(&H.wfull), 0);
            if (rank == 0) mbar_expect_tx(&H.wfull, 2u * (uint32_t)nkb * kWBytes);
            for (int kb = 0; kb < nkb; ++kb) {
                const int t = kb / P.kblocks, kc = kb - t * P.kblocks;
                tma_load_2d_pair(&b_map, wfull0, wsm + (size_t)kb * BH * kTileK, kc * P.ktile, P.b_row[t] + n0 + (int)rank * BH);
            }
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < mtiles; tile += gridDim.x) {
                const int row0 = ws_row_tile(P, tile, mtiles) * kTileM;
                const int b0 = row0 >> P.mper_shift, m0 = row0 & (P.Mper - 1);
                for (int j = 0; j < npt; ++j) {
                    mbar_wait(&H.empty[s], ph ^ 1);
                    const int4 L = H.ldtab[j];
                    if (rank == 0) mbar_expect_tx(&H.full[s], 2u * a_tx);
                    const uint32_t full0 = mapa_u32(smem_u32(&H.full[s]), 0);
                    if (P.il > 1) tma_load_4d_pair(&a_map, full0, asm_ + (size_t)s * a_stage, L.x, b0, L.y, m0 + L.z);
                    else tma_load_4d_pair(&a_map, full0, asm_ + (size_t)s * a_stage, L.x, L.y, m0 + L.z, b0);
                    if (++s == nstages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            constexpr uint32_t idesc = make_idesc_pair(BN);
            mbar_wait(&H.wfull, 0);
            int s = 0, tcount = 0;
            uint32_t ph = 0;
            const uint64_t dbase = make_smem_desc(0, 16, 1024);
            const uint32_t a_base = smem_u32(asm_), w_base = smem_u32(wsm);
            for (int tile = blockIdx.x; tile < mtiles; tile += gridDim.x, ++tcount) {
                const int acc = tcount & 1;
                mbar_wait(&H.tmem_empty[acc], ((tcount >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tacc = tmem0 + (uint32_t)(acc * BN);
                uint32_t started = 0;
                int i = 0;
                for (int j = 0; j < npt; ++j) {
                    mbar_wait(&H.full[s], ph);
                    tc_fence_after();
                    const uint32_t stage_addr = a_base + (uint32_t)s * a_stage;
                    int last;
                    do {
                        const int4 M = H.mmatab[i++];
                        const uint64_t ad = dbase | (uint64_t)(((stage_addr + (uint32_t)M.x) >> 4) & 0x3FFF);
                        const uint64_t bd = dbase | (uint64_t)(((w_base + (uint32_t)M.y) >> 4) & 0x3FFF);
                        if (!(P.dbg & 2) && elect_one()) {
#pragma unroll
                            for (int k = 0; k < kTileK / 16; ++k) umma_f16_pair(tacc, ad + 2u * k, bd + 2u * k, idesc, k ? 1u : started);
                        }
                        started = 1;
                        last = M.z;
                    } while (!last);
                    if (elect_one()) {
                        umma_commit_pair(&H.empty[s]);
                        if (j == npt - 1) umma_commit_pair(&H.tmem_full[acc]);
                    }
                    __syncwarp();
                    if (++s == nstages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else {
        const int et = threadIdx.x - 64;
        constexpr int kEpi = WsCfg<BN>::kEpiThreads;
        for (int i = et; i < BN; i += kEpi) {
            H.bias[i] = P.bias ? __ldg(P.bias + perm_index(n0 + i, P.n_perm_q, P.n_perm_p)) : 0.0f;
            H.scale[i] = (P.col_scale ? __ldg(P.col_scale + n0 + i) : 1.0f) * P.alpha;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpi) : "memory");
        const bool scaled = P.col_scale != nullptr || P.alpha != 1.0f;
        const bool affine = P.bias != nullptr || scaled;
        int variant = 0;
        if (P.mul_mode == MUL_NONE && !P.aux && P.act != ACT_GELU) variant = (scaled ? 1 : 8) + P.act;
        else if (P.mul_mode == MUL_NONE && P.aux && P.act == ACT_GELU) variant = 4;
        else if (P.act == ACT_NONE && !P.aux && !affine && P.mul_mode != MUL_NONE) variant = 4 + P.mul_mode;
        float* poolbuf = (P.pool_out || P.colsum_out || P.stats_out) ? reinterpret_cast<float*>(maskbuf + (P.tma_mask ? (size_t)P.nmb * (BN / 64) * 16384 : 0)) : nullptr;
#define MG_LOOP(ACT_, MUL_, AFF_, AUX_, GEN_, SCL_)                                                                  \
    ws_epilogue_loop<BN, kEpi, WsCfg<BN>::kCpt, ACT_, MUL_, AFF_, AUX_, GEN_, SCL_, TO, TMSK, WsHeader<BN>, true>(   \
        P, &o_map, &x_map, &m_map, H, tmem0, n0, mtiles, staging, maskbuf, poolbuf)
        switch (variant) {
            case 1: MG_LOOP(ACT_NONE, MUL_NONE, true, false, false, true); break;
            case 2: MG_LOOP(ACT_RELU, MUL_NONE, true, false, false, true); break;
            case 3: MG_LOOP(ACT_LRELU, MUL_NONE, true, false, false, true); break;
            case 4: MG_LOOP(ACT_GELU, MUL_NONE, true, true, false, true); break;
            case 5: MG_LOOP(ACT_NONE, MUL_LRELU_SIGN, false, false, false, true); break;
            case 6: MG_LOOP(ACT_NONE, MUL_RELU_SIGN, false, false, false, true); break;
            case 7: MG_LOOP(ACT_NONE, MUL_VALUE, false, false, false, true); break;
            case 8: MG_LOOP(ACT_NONE, MUL_NONE, true, false, false, false); break;
            case 9: MG_LOOP(ACT_RELU, MUL_NONE, true, false, false, false); break;
            case 10: MG_LOOP(ACT_LRELU, MUL_NONE, true, false, false, false); break;
            default: MG_LOOP(ACT_NONE, MUL_NONE, true, true, true, true); break;
        }
#undef MG_LOOP
    }
    tc_fence_before();
    cluster_sync_all();                               // nobody leaves while the peer may still signal it or read its operands
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem0, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------
// wgrad
// ---------------------------------------------------------------------------------------------
struct TcWgradArgs {
    int ntaps, K, N;                    // K, N: channel counts of A and G (multiples of 64 / 128)
    int a_p[kMaxTaps], a_dm[kMaxTaps];
    int rpt, spt;                       // reduction chunk = spt samples x rpt rows, rpt * spt = 64
    int Mper;
    long long row_begin, row_end;       // flattened (b, m) rows to reduce over
    int rows_per_split;                 // multiple of 64
    float* dW; int w_toff[kMaxTaps]; int w_nstride; int w_kstride; int n_perm_q, n_perm_p;
    float alpha;
};

template <int BNK>
struct WgradSmem {
    __nv_bfloat16 g[kStages][2 * 64 * 64];          // two 64-channel boxes x 64 positions
    __nv_bfloat16 a[kStages][(BNK / 64) * 64 * 64];
    uint64_t full[kStages], empty[kStages], tmem_full;
    uint32_t tmem_base;
};

template <int BNK>
__global__ void __launch_bounds__(192) tc_wgrad_kernel(const __grid_constant__ CUtensorMap g_map,
                                                       const __grid_constant__ CUtensorMap a_map, const TcWgradArgs P) {
    extern __shared__ unsigned char smem_raw[];
    WgradSmem<BNK>& S = *reinterpret_cast<WgradSmem<BNK>*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = BNK < 32 ? 32 : BNK;
    constexpr uint32_t kStageBytes = (2 + BNK / 64) * 64 * 64 * 2;
    constexpr uint32_t kBoxBytes = 64 * 64 * 2;     // one TMA box: 64 positions x 128 B

    const int n0 = blockIdx.x * kTileM;
    const int kpt = P.K / BNK;                      // column tiles per tap
    const int t = blockIdx.y / kpt, k0 = (blockIdx.y % kpt) * BNK;
    const long long r_begin = P.row_begin + (long long)blockIdx.z * P.rows_per_split;
    long long r_end = r_begin + P.rows_per_split;
    if (r_end > P.row_end) r_end = P.row_end;
    const int nchunks = r_end > r_begin ? (int)((r_end - r_begin + 63) / 64) : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
        mbar_init(&S.tmem_full, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc(&S.tmem_base, kTmemCols);
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&g_map); tma_prefetch_desc(&a_map); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = S.tmem_base;

    if (nchunks > 0) {
        if (warp == 0) {
            if (lane == 0) {
                for (int c = 0; c < nchunks; ++c) {
                    const int s = c % kStages, it = c / kStages;
                    mbar_wait(&S.empty[s], (it & 1) ^ 1);
                    const long long r = r_begin + (long long)c * 64;
                    int bq, mq;
                    if (P.rpt == 64) { bq = (int)(r / P.Mper); mq = (int)(r % P.Mper); }
                    else { bq = (int)(r / P.Mper); mq = 0; }
                    mbar_expect_tx(&S.full[s], kStageBytes);
                    tma_load_4d(&g_map, &S.full[s], S.g[s], n0, 0, mq, bq);
                    tma_load_4d(&g_map, &S.full[s], S.g[s] + 64 * 64, n0 + 64, 0, mq, bq);
#pragma unroll
                    for (int j = 0; j < BNK / 64; ++j)
                        tma_load_4d(&a_map, &S.full[s], S.a[s] + j * 64 * 64, k0 + j * 64, P.a_p[t], mq + P.a_dm[t], bq);
                }
            }
        } else if (warp == 1) {
            constexpr uint32_t idesc = make_idesc(BNK, 1, 1);
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % kStages, it = c / kStages;
                mbar_wait(&S.full[s], it & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t g_addr = smem_u32(S.g[s]), a_addr = smem_u32(S.a[s]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {       // 16 positions (= 16 rows of 128 B) per MMA
                        const uint64_t gd = make_smem_desc(g_addr + k * 16 * 128, kBoxBytes, 1024);
                        const uint64_t ad = make_smem_desc(a_addr + k * 16 * 128, kBoxBytes, 1024);
                        umma_f16(tmem_acc, gd, ad, idesc, (c | k) ? 1u : 0u);
                    }
                    umma_commit(&S.empty[s]);
                    if (c == nchunks - 1) umma_commit(&S.tmem_full);
                }
                __syncwarp();
            }
        } else {
            mbar_wait(&S.tmem_full, 0);
            tc_fence_after();
            const int q = warp & 3;
            const int n = n0 + q * 32 + lane;
            const long long nphys = perm_index(n, P.n_perm_q, P.n_perm_p);
            float* __restrict__ dst = P.dW + P.w_toff[t] + nphys * P.w_nstride;
#pragma unroll 1
            for (int c0 = 0; c0 < BNK; c0 += 16) {
                float v[16];
                tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                if (n < P.N) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        atomicAdd(dst + (long long)(k0 + c0 + j) * P.w_kstride, P.alpha * v[j]);
                }
            }
            tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_acc, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------
// weight packing: Wp[(t*N + n)*K + k] (bf16) = W[w_toff[t] + nphys(n)*w_nstride + kphys(k)*w_kstride]
// ---------------------------------------------------------------------------------------------
struct PackArgs {
    const float* W; int w_toff[kMaxTaps]; int w_nstride, w_kstride, n_perm_q, n_perm_p, k_perm_q, k_perm_p;
    int ntaps, N, K;
    void* out; int f32;          // packed [tap][n][k] as bf16, or as fp32 for the TF32 path
    int split;                   // fp32 on the tensor cores: K = 6 x the layer's K, segment s of a row holds part kSplitW[s]
};
// ---- fp32 contractions on the bf16 tensor cores (MELOGAN_FP32_TC / mg_debug_set("fp32_tc")) ----
// A float32 x is EXACTLY h + m + l with h = bf16(x), m = bf16(x - h), l = bf16(x - h - m) (8 + 8 + 8 significand bits, each
// difference is exact in float32).  x * w = hh + hm + mh + hl + lh + mm + O(2^-26 |x w|): six bf16 products, each exact in
// the fp32 accumulator.  The six terms are laid out along the reduction index -- row k-segments of A: h h m h l m, of W:
// h m h l h m -- so that the UNCHANGED bf16 tap-GEMM kernels compute the float32 contraction with K' = 6 K; the weight
// gradient stacks the same parts along the samples instead (its reduction runs over the rows).
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ void split3(float x, float (&p)[3]) {
    p[0] = bf16_round(x);
    const float r = x - p[0];
    p[1] = bf16_round(r);
    p[2] = bf16_round(r - p[1]);
}
// Smallest terms first (mm, hl, lh, hm, mh, hh): tcgen05 adds into its fp32 accumulator with truncation, an error of up to one
// ulp OF THE ACCUMULATOR per instruction -- while only small terms have been added the accumulator, and with it that error, is
// 2^-8 of its final size.  The one-tile kernel therefore walks the k-blocks outermost (TcTapArgs::kb_outer): the small-term
// segments of every tap come before the first full-size product; see DESIGN.md 5.
constexpr unsigned kSplitA = 0x010201u;   // part of segment s = (order >> 4 s) & 3:  m h l h m h
constexpr unsigned kSplitW = 0x001021u;   //                                           m l h m h h
static __global__ void __launch_bounds__(256) pack_weight_kernel(const PackArgs P) {
    const long long total = (long long)P.ntaps * P.N * P.K;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int K0 = P.split ? P.K / 6 : P.K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int kk = (int)(i % P.K);
        const int seg = kk / K0, k = kk - seg * K0;
        const long long tn = i / P.K;
        const int n = (int)(tn % P.N), t = (int)(tn / P.N);
        float w = __ldg(P.W + P.w_toff[t] + (long long)perm_index(n, P.n_perm_q, P.n_perm_p) * P.w_nstride +
                        (long long)perm_index(k, P.k_perm_q, P.k_perm_p) * P.w_kstride);
        if (P.split) { float p[3]; split3(w, p); w = p[(kSplitW >> (4 * seg)) & 3u]; }
        if (P.f32) static_cast<float*>(P.out)[i] = w;
        else static_cast<__nv_bfloat16*>(P.out)[i] = __float2bfloat16_rn(w);
    }
}
// y[(i / inner) * 6 * inner + s * inner + i % inner] = part order[s] of x[i]   (4 elements per thread; inner % 4 == 0).
// inner = K: the six k-segments of every activation row (tap-GEMM); inner = n: six stacked copies of the tensor (wgrad).
template <unsigned ORDER>
static __global__ void __launch_bounds__(256) split3_kernel(const float4* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                            long long n4, long long inner) {
    for (long long i4 = (long long)blockIdx.x * 256 + threadIdx.x; i4 < n4; i4 += (long long)gridDim.x * 256) {
        const float4 v = __ldg(x + i4);
        float p[4][3];
        split3(v.x, p[0]); split3(v.y, p[1]); split3(v.z, p[2]); split3(v.w, p[3]);
        const long long i = i4 * 4, outer = i / inner, in = i - outer * inner;
        __nv_bfloat16* row = y + outer * 6 * inner + in;
#pragma unroll
        for (int s = 0; s < 6; ++s) {
            const int q = (int)((ORDER >> (4 * s)) & 3u);      // a compile-time constant once the loop is unrolled
            const __nv_bfloat162 a = __floats2bfloat162_rn(p[0][q], p[1][q]), b = __floats2bfloat162_rn(p[2][q], p[3][q]);
            *reinterpret_cast<uint2*>(row + s * inner) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static inline bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }
struct Scratch {             // ring of packed-weight slots (stream-ordered reuse)
    __nv_bfloat16* slot[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t slot_elems = 0;
    int next = 0;
};
Scratch& scratch();
int ensure_scratch(size_t elems);   // allocate the ring (not capturable: call before any CUDA-graph capture)
// Packed-weight cache (mg_gan_weight_cache): one persistent packed copy per (weight tensor, layout); an entry is
// re-packed only after mg_adam_step touched its source (weights_changed) or after mg_weight_cache_invalidate.  Returns
// the entry's buffer and whether it already holds the current weights; nullptr when the cache is off (or cannot
// allocate, e.g. a new layout during CUDA-graph capture) -- the caller then packs into the scratch ring as before.
void* cache_lookup(const PackArgs& key, size_t bytes, bool* is_current);
void weights_changed(const float* param, long long n);
void set_cache_mode(bool on);       // per API call, from the context's mg_gan_weight_cache setting (like set_tf32)
bool enabled();              // MELOGAN_DISABLE_TC=1 forces the CUDA-core kernels (A/B testing)
bool ws_enabled();           // MELOGAN_DISABLE_WS=1 keeps the non-persistent tensor-core kernel (A/B profiling)

// 4-D view (k, parity, row, sample) of a channels-last activation [B][L][C] (bf16):
//   stride 1: dims (C, 1, L, B);  stride 2: dims (C, 2, L/2, B)
int make_act_map(CUtensorMap* map, const void* base, int C, int L, long long B, int stride, int box_rows,
                 int box_samples, int elem_bytes = 2);
// the same activation as dims (k, sample, parity, row): a box of box_samples x box_rows lands in shared memory as
// [row][sample][k] (see TcTapArgs::il); coordinates (k, first sample, plane, first row)
int make_act_map_interleaved(CUtensorMap* map, const void* base, int C, int L, long long B, int stride, int box_rows,
                             int box_samples, int elem_bytes = 2);
// general 4-D bf16 view: dims (64-wide inner box over `inner` elements, planes, rows, samples) with explicit
// element strides; used for the overlapping note windows and the row-mod-P planes of the banded layers
int make_view_map(CUtensorMap* map, const void* base, long long inner, long long planes, long long plane_stride,
                  long long rows, long long row_stride, long long samples, long long sample_stride, int box_rows,
                  int box_samples);
// 2-D view (column, flat row) of an output tensor whose rows are row_stride elements apart; boxes of 128 bytes x 128 rows
int make_out_map(CUtensorMap* map, const void* base, int elem_bytes, long long cols, long long rows, long long row_stride);
bool mask_tma_enabled();     // MELOGAN_DISABLE_TMA_MASK=1 keeps per-thread mask loads (A/B profiling)
bool tma_store_enabled();    // MELOGAN_DISABLE_TMA_STORE=1 keeps the row-per-thread epilogue stores (A/B profiling)
// 2-D view (k, rows) of a packed weight [rows][K]
int make_weight_map(CUtensorMap* map, const void* base, int K, long long rows, int box_rows, int elem_bytes = 2);
bool fp32_tc_enabled();      // tuning().fp32_tc, else MELOGAN_FP32_TC=1
// grow-only device scratch of the fp32_tc path (0 = split activations, 1 = packed split weights, 2 = split gradients); like
// wgrad_cast_scratch it never frees (captured graphs keep the addresses) and refuses to grow inside a stream capture
void* split_scratch(int kind, size_t bytes, cudaStream_t st);
bool tf32_enabled();         // set per API call from the context's precision (fp32 parity mode never uses TF32)
void set_tf32(bool on);

template <int BN, typename TO, typename TMSK, bool TF32 = false>
int launch_tc_tap(const CUtensorMap& am, const CUtensorMap& bm, const TcTapArgs& a, int mtiles, cudaStream_t st) {
    static bool attr_done = false;
    const size_t smem = sizeof(TapSmem<BN>) + 1024;
    if (!attr_done) {
        MG_CUDA_OK(cudaFuncSetAttribute(tc_tapgemm_kernel<BN, TO, TMSK, TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    dim3 grid(mtiles, a.N / BN);
    tc_tapgemm_kernel<BN, TO, TMSK, TF32><<<grid, 192, smem, st>>>(am, bm, a);
    MG_LAUNCH_OK();
    return MG_OK;
}

template <int BN, typename TO, typename TMSK, bool TF32 = false>
int launch_tc_tap_ws(const CUtensorMap& am, const CUtensorMap& bm, const CUtensorMap& om, const CUtensorMap& xm,
                     const CUtensorMap& mm, const TcTapArgs& a, int mtiles, int nstages, int ctas_x, size_t smem, cudaStream_t st) {
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        MG_CUDA_OK(cudaFuncSetAttribute(tc_tapgemm_ws_kernel<BN, TO, TMSK, TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(227 * 1024)));
        attr_smem = 227 * 1024;
    }
    dim3 grid(ctas_x, a.N / BN);
    tc_tapgemm_ws_kernel<BN, TO, TMSK, TF32><<<grid, WsCfg<BN>::kThreads, smem, st>>>(am, bm, om, xm, mm, a, mtiles, nstages);
    MG_LAUNCH_OK();
    return MG_OK;
}

// CTA-pair form: cluster (2, 1, 1), even grid.x
template <typename TO, typename TMSK>
int launch_tc_tap_ws2(const CUtensorMap& am, const CUtensorMap& bm, const CUtensorMap& om, const CUtensorMap& xm,
                      const CUtensorMap& mm, const TcTapArgs& a, int mtiles, int nstages, int ctas_x, size_t smem, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        MG_CUDA_OK(cudaFuncSetAttribute(tc_tapgemm_ws2_kernel<TO, TMSK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas_x, a.N / 128);
    cfg.blockDim = dim3(WsCfg<128>::kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MG_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_tapgemm_ws2_kernel<TO, TMSK>, am, bm, om, xm, mm, a, mtiles, nstages));
    MG_LAUNCH_OK();
    return MG_OK;
}

template <int BNK>
int launch_tc_wgrad(const CUtensorMap& gm, const CUtensorMap& am, const TcWgradArgs& a, int splits, cudaStream_t st) {
    static bool attr_done = false;
    const size_t smem = sizeof(WgradSmem<BNK>) + 1024;
    if (!attr_done) {
        MG_CUDA_OK(cudaFuncSetAttribute(tc_wgrad_kernel<BNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    dim3 grid(a.N / kTileM, a.ntaps * (a.K / BNK), splits);
    tc_wgrad_kernel<BNK><<<grid, 192, smem, st>>>(gm, am, a);
    MG_LAUNCH_OK();
    return MG_OK;
}

// Tap groups for the weight-stationary kernel.  halo_max = 0: every tap is its own group (one 128-row tile per tap).
// Otherwise taps of the same plane whose row shifts lie within halo_max rows share one tile with a halo.
inline void build_tap_groups(TcTapArgs& a, int halo_max) {
    int order[kTcMaxTaps];
    for (int t = 0; t < a.ntaps; ++t) order[t] = t;
    if (halo_max > 0)
        for (int i = 1; i < a.ntaps; ++i)          // insertion sort by (plane, row shift)
            for (int j = i; j > 0; --j) {
                const int x = order[j - 1], y = order[j];
                if (a.a_p[x] < a.a_p[y] || (a.a_p[x] == a.a_p[y] && a.a_dm[x] <= a.a_dm[y])) break;
                order[j - 1] = y; order[j] = x;
            }
    a.ngroups = 0; a.halo = 0;
    for (int i = 0; i < a.ntaps; ++i) {
        const int t = order[i];
        a.g_tap[i] = t;
        const int g = a.ngroups - 1;
        if (g >= 0 && halo_max > 0 && a.g_p[g] == a.a_p[t] && a.a_dm[t] - a.g_dmin[g] <= halo_max) {
            a.g_count[g]++;
            if (a.a_dm[t] - a.g_dmin[g] > a.halo) a.halo = a.a_dm[t] - a.g_dmin[g];
        } else {
            a.g_p[a.ngroups] = a.a_p[t]; a.g_dmin[a.ngroups] = a.a_dm[t]; a.g_first[a.ngroups] = i; a.g_count[a.ngroups] = 1;
            a.ngroups++;
        }
    }
}

bool reuse_enabled();        // MELOGAN_DISABLE_TAP_REUSE=1: one activation tile per tap (A/B profiling)
bool pair_enabled();         // MELOGAN_DISABLE_PAIR=1: never the CTA-pair (cta_group::2) kernel (A/B profiling)

// Kernel-selection overrides for the per-layer harness (mg_debug_set, csrc/debug.cu); all zero = the product heuristics.
struct Tuning {
    int force_bn = 0;        // 64 / 128: slab width of the weight-stationary kernel
    int max_stages = 0;      // cap of the activation ring
    int staging_bufs = 0;    // 1 / 2 staging tiles (0 = two when they fit)
    int mask_bufs = 0;       // 1 = single TMA mask tile (0 = two when they fit)
    int no_ws = 0;           // one tile per CTA
    int dbg = -1;            // MELOGAN_TC_DEBUG ablation bits (-1 = environment)
    int reverse = -1;        // fixed tile order (-1 = alternate)
    int no_tma_store = 0, no_tma_mask = 0, no_reuse = 0;
    int no_pair = 0;         // 1 = never the CTA-pair (cta_group::2) kernel
    int no_rot = 0;          // 1 = every slab walks the row tiles in the same order
    int ring2_stages = 0;    // activation stages a second staging set must leave (0 = 4)
    int fp32_tc = -1;        // float32 contractions as six bf16 tensor-core terms (-1 = MELOGAN_FP32_TC, default off)
    int no_fuse = 0;         // bits: 1 = no pooling, 2 = no column sums, 4 = no BatchNorm statistics in the epilogues (the callers
                             // then run their own passes)
};
Tuning& tuning();
// What the last tensor-core launch on this thread looked like (tests assert that the variant they mean to pin ran).
struct LaunchInfo {
    int kind = 0;            // 0 none, 1 tap-GEMM, 2 wgrad
    long long rows = 0;
    int N = 0, K = 0, taps = 0, groups = 0, halo = 0, BN = 0, out_bytes = 0, ws = 0, stages = 0, act = 0, mul = 0, aux = 0;
    int tma_store = 0, tma_mask = 0, nsb = 0, reverse = 0, ctas_x = 0, slabs = 0, tf32 = 0, splits = 0, pair = 0, pool = 0;
    int fp32x6 = 0;          // the launch was the six-term bf16 form of a float32 contraction
    double flops = 0, bytes = 0;
};
LaunchInfo& last_launch();

// Launch a tap-GEMM whose tensor maps and TcTapArgs are ready: weight-stationary persistent form when the slab's
// weights fit in shared memory next to >= 3 activation stages and every CTA gets >= 4 tiles, else one tile per CTA.
// am_halo (optional): the same activation view with boxes of 128 + a.halo rows, a.ngroups/g_* describing the tap groups.
template <typename TO, typename TMSK, bool TF32 = false>
int run_tc_tap(const CUtensorMap& am, const CUtensorMap& bm, TcTapArgs a, int BN, int K, cudaStream_t st,
               const CUtensorMap* am_halo = nullptr, const CUtensorMap* bm_half = nullptr, bool pool_only = false) {
    a.ktile = TF32 ? 32 : 64;
    if (!is_pow2(a.Mper) || !is_pow2(a.mpt) || a.mpt * a.bpt != kTileM || (a.Mper > kTileM && a.Mper % kTileM)) {
        set_error("run_tc_tap: Mper=%d mpt=%d bpt=%d is not a power-of-two tiling", a.Mper, a.mpt, a.bpt);
        return MG_ERR_INVALID;
    }
    a.mper_shift = 0; while ((1 << a.mper_shift) < a.Mper) ++a.mper_shift;
    a.mpt_shift = 0; while ((1 << a.mpt_shift) < a.mpt) ++a.mpt_shift;
    if (!am_halo || a.il < 1) a.il = 1;                  // interleaved halo tiles only come with a halo map
    a.il_shift = 0; while ((1 << a.il_shift) < a.il) ++a.il_shift;
    const long long rows = (long long)a.B * a.Mper;
    const int mtiles = (int)((rows + 127) / 128);
    // algorithmic traffic: the activation once, the output (and derivative tile), the mask tile
    ProbeScope probe(PROBE_TC_GEMM, 2.0 * (double)rows * a.N * a.ntaps * K,
                     (double)rows * (K * 2.0 + a.N * sizeof(TO) * (a.aux ? 2.0 : 1.0) +
                                     (a.mul_mode != MUL_NONE ? a.N * (double)sizeof(TMSK) : 0.0)), st);
    const int nslabs = a.N / BN;
    const Tuning& tn = tuning();
    if (!am_halo) build_tap_groups(a, 0);
    const size_t avail = (size_t)227 * 1024 - 1024 - kWsHeaderBytes;
    const size_t a_stage = (((size_t)(128 + a.halo * a.il) * 128) + 1023) / 1024 * 1024;
    // tiles can leave through a staging tile + TMA bulk stores: output rows uniformly strided (conv outputs [b][m][n], and
    // Linears: one row per sample), 16-byte aligned
    const long long row_stride = a.Mper == 1 ? a.o_bstride : a.o_mstride;
    const int per_tile = a.aux ? 2 : 1;                            // staging slots one tile needs (result + derivative tile)
    const size_t staging = (size_t)128 * BN * sizeof(TO) * per_tile;
    const bool tma_ok = tma_store_enabled() && !tn.no_tma_store && !a.accumulate &&
                        (a.Mper == 1 || a.o_bstride == (long long)a.Mper * a.o_mstride) && (row_stride * sizeof(TO)) % 16 == 0 &&
                        ((uintptr_t)((TO*)a.Out + a.o_off)) % 16 == 0 && (!a.aux || sizeof(TO) == 2);
    // CTA pairs (tc_tapgemm_ws2_kernel): bf16 operands, 128-wide slabs, an even number of tiles, the TMA-store epilogue, each
    // CTA keeps half of the slab's weights
    const size_t wbytes_pair = (size_t)a.ntaps * (K / a.ktile) * 64 * 128;
    // (measured per layer at the bench shapes, gpurun_out/r02o_*: never slower than one CTA per tile row, 5-35 % faster from
    // taps x K = 192 up -- critic conv.4 forward 616 -> 440 us, ED conv.1 forward 718 -> 490 us, ED conv.3 dgrad 2287 -> 1547 us)
    // More slabs than SM pairs (decoder.pre.2: 128 slabs, 64 row tiles): one pair per slab, in two waves of clusters -- half
    // the resident weights buy an 8-deep activation ring where the single CTA had 3 stages and sat on TMA latency.
    const bool pair = !TF32 && BN == 128 && bm_half && pair_enabled() && !tn.no_pair && (mtiles % 2 == 0) &&
                      (num_sms() / nslabs >= 2 || mtiles >= 8) && tma_ok && wbytes_pair + 3 * a_stage + staging <= avail;
    const size_t wbytes = pair ? wbytes_pair : (size_t)a.ntaps * (K / a.ktile) * BN * 128;
    int ctas_x = num_sms() / nslabs;
    if (pair) ctas_x = ctas_x < 2 ? 2 : (ctas_x & ~1);
    if (ctas_x < 1) ctas_x = 1;
    if (ctas_x > mtiles) ctas_x = mtiles;
    {   // Consecutive layers are 0.4-0.8 GB producer -> consumer hand-offs through a 126 MB L2: the consumer starts where
        // the producer stopped (its last tiles are still resident) if successive launches walk the rows in opposite order.
        static const bool snake = getenv("MELOGAN_NO_SNAKE") == nullptr;
        static unsigned launch_no = 0;
        a.reverse = snake ? (int)(launch_no++ & 1u) : 0;
        if (tn.reverse >= 0) a.reverse = tn.reverse;
    }
    { static const int dbg = getenv("MELOGAN_TC_DEBUG") ? atoi(getenv("MELOGAN_TC_DEBUG")) : 0; a.dbg = tn.dbg >= 0 ? tn.dbg : dbg; }
    // staggered walks when more than two slabs stream the same activation (an even step keeps a pair's tiles adjacent)
    a.rot_step = (nslabs > 2 && mtiles >= 8 && !tn.no_rot) ? 2 : 0;
    static const bool trace = getenv("MELOGAN_TRACE") != nullptr;
    const bool ws = ws_enabled() && !tn.no_ws && wbytes + 3 * a_stage <= avail && mtiles >= 4 * ctas_x &&
                    a.ntaps * (K / a.ktile) <= kWsMaxLoads;
    if (trace)
        fprintf(stderr, "[tc_tap] rows=%lld N=%d K=%d taps=%d groups=%d halo=%d BN=%d out%zu ws=%d stages=%d act=%d mul=%d aux=%d\n",
                rows, a.N, K, a.ntaps, a.ngroups, a.halo, BN, sizeof(TO), (int)ws, ws ? (int)((avail - wbytes) / a_stage) : 3, a.act,
                a.mul_mode, a.aux != nullptr);
    LaunchInfo& li = last_launch();
    const double flops0 = li.kind == 1 ? li.flops : 0.0, bytes0 = li.kind == 1 ? li.bytes : 0.0;   // sub-pixel phases add up
    li = LaunchInfo();
    li.kind = 1; li.rows = rows; li.N = a.N; li.K = K; li.taps = a.ntaps; li.groups = a.ngroups; li.halo = a.halo; li.BN = BN;
    li.out_bytes = (int)sizeof(TO); li.ws = ws; li.stages = 3; li.act = a.act; li.mul = a.mul_mode; li.aux = a.aux != nullptr;
    li.reverse = a.reverse; li.ctas_x = ws ? ctas_x : mtiles; li.slabs = nslabs; li.tf32 = TF32;
    li.flops = flops0 + 2.0 * (double)rows * a.N * a.ntaps * K;
    li.bytes = bytes0 + (double)rows * (K * (TF32 ? 4.0 : 2.0) + a.N * sizeof(TO) * (a.aux ? 2.0 : 1.0) +
                                        (a.mul_mode != MUL_NONE ? a.N * (double)sizeof(TMSK) : 0.0));
    if (ws) {
        // tiles leave through a staging tile + TMA bulk stores when the output rows are uniformly strided and the staging
        // tile fits next to >= 3 activation stages
        CUtensorMap om = am, xm = am;
        a.tma_store = 0; a.nsb = per_tile;
        if (tma_ok && wbytes + 3 * a_stage + staging <= avail) {
            int rc = make_out_map(&om, (TO*)a.Out + a.o_off, (int)sizeof(TO), a.N, rows, row_stride);
            if (rc != MG_OK) return rc;
            if (a.aux) {
                rc = make_out_map(&xm, (TO*)a.aux + a.o_off, (int)sizeof(TO), a.N, rows, row_stride);
                if (rc != MG_OK) return rc;
            }
            a.tma_store = 1;
        }
        // the mask tile (same layout as the output) by TMA too when it is bf16 and its buffer still leaves 3 stages
        CUtensorMap mm = am;
        const size_t maskbytes = (size_t)128 * BN * 2;
        a.tma_mask = 0;
        if (a.tma_store && a.mul_mode != MUL_NONE && sizeof(TMSK) == 2 && mask_tma_enabled() && !tn.no_tma_mask &&
            wbytes + 3 * a_stage + staging + maskbytes <= avail && ((uintptr_t)((const TMSK*)a.mul_src + a.o_off)) % 16 == 0) {
            const int rc = make_out_map(&mm, (const TMSK*)a.mul_src + a.o_off, 2, a.N, rows, row_stride);
            if (rc != MG_OK) return rc;
            a.tma_mask = 1;
        }
        // a second staging tile when it still leaves 4 activation stages: the drain of the bulk store (shared memory ->
        // L2) then overlaps the next tile's TMEM read and epilogue math instead of serialising with them
        // (GELU layers that also store the derivative tile are bound by their epilogue, not by the activation ring: two tiles'
        // worth of staging with a 2-deep ring beats one with 6 -- ED conv.2 forward 1450 -> 1284 us -- because with a single
        // staging set every tile waits for the previous tile's 64 KB bulk store to drain)
        const int ring2_min_stages = tn.ring2_stages > 0 ? tn.ring2_stages : ((a.aux && a.act == ACT_GELU) ? 2 : 4);
        if (a.tma_store && tn.staging_bufs != 1 &&
            wbytes + ring2_min_stages * a_stage + 2 * staging + (a.tma_mask ? maskbytes : 0) <= avail)
            a.nsb = 2 * per_tile;
        // a second mask tile (prefetch distance two tiles) when it still leaves 4 activation stages
        a.nmb = 1;
        if (a.tma_mask && tn.mask_bufs != 1 &&
            wbytes + 4 * a_stage + (size_t)128 * BN * sizeof(TO) * a.nsb + 2 * maskbytes <= avail)
            a.nmb = 2;
        // fused pooling (ws_pool_*): bf16 staging tile of a 128-wide slab, 1 or 2 samples per tile, 8 KB of partial sums
        const size_t poolbytes = 2 * 8 * 128 * sizeof(float);
        if (tn.no_fuse & 1) a.pool_out = nullptr;
        if (tn.no_fuse & 2) a.colsum_out = nullptr;
        if (tn.no_fuse & 4) a.stats_out = nullptr;
        const bool pool = a.pool_out && a.tma_store && BN == 128 && sizeof(TO) == 2 && a.bpt <= 2 && !a.accumulate &&
                          wbytes + 3 * a_stage + (size_t)128 * BN * sizeof(TO) * a.nsb + (a.tma_mask ? maskbytes * a.nmb : 0) + poolbytes <= avail;
        if (!pool) a.pool_out = nullptr;
        a.pool_atomic = a.Mper > 128;
        a.skip_out = pool && pool_only;
        // fused column sums (bias gradients): any bf16 staging tile; whole tiles only
        const bool colsum = a.colsum_out && !pool && a.tma_store && sizeof(TO) == 2 && !a.accumulate && rows % 128 == 0 &&
                            wbytes + 3 * a_stage + (size_t)128 * BN * sizeof(TO) * a.nsb + (a.tma_mask ? maskbytes * a.nmb : 0) + poolbytes <= avail;
        if (!colsum) a.colsum_out = nullptr;
        // fused BatchNorm statistics: float32 staging tile, whole tiles only
        const bool stats = a.stats_out && !pool && !colsum && a.tma_store && sizeof(TO) == 4 && !a.accumulate && rows % 128 == 0 &&
                           wbytes + 3 * a_stage + (size_t)128 * BN * sizeof(TO) * a.nsb + (a.tma_mask ? maskbytes * a.nmb : 0) + poolbytes <= avail;
        if (!stats) a.stats_out = nullptr;
        const size_t extra = (a.tma_store ? (size_t)128 * BN * sizeof(TO) * a.nsb : 0) + (a.tma_mask ? maskbytes * a.nmb : 0) +
                             (pool || colsum || stats ? poolbytes : 0);
        int nstages = (int)((avail - wbytes - extra) / a_stage);
        if (nstages > kWsMaxStages) nstages = kWsMaxStages;
        if (tn.max_stages > 0 && nstages > tn.max_stages) nstages = tn.max_stages < 2 ? 2 : tn.max_stages;
        li.stages = nstages; li.tma_store = a.tma_store; li.tma_mask = a.tma_mask; li.nsb = a.nsb; li.pair = pair;
        const size_t smem = 1024 + kWsHeaderBytes + wbytes + (size_t)nstages * a_stage + extra;
        const CUtensorMap& amap = am_halo ? *am_halo : am;
        if (pool) {
            if (a.pool_atomic) MG_CUDA_OK(cudaMemsetAsync(a.pool_out, 0, (size_t)a.B * a.N * sizeof(float), st));
            if (a.pool_done) *a.pool_done = 1;
        }
        if (colsum && a.colsum_done) *a.colsum_done = 1;
        if (stats && a.stats_done) *a.stats_done = 1;
        li.pool = pool ? 1 : (colsum ? 2 : (stats ? 3 : 0));
        if (pair) return launch_tc_tap_ws2<TO, TMSK>(amap, *bm_half, om, xm, mm, a, mtiles, nstages, ctas_x, smem, st);
        return (BN == 128) ? launch_tc_tap_ws<128, TO, TMSK, TF32>(amap, bm, om, xm, mm, a, mtiles, nstages, ctas_x, smem, st)
                           : launch_tc_tap_ws<64, TO, TMSK, TF32>(amap, bm, om, xm, mm, a, mtiles, nstages, ctas_x, smem, st);
    }
    a.il = 1; a.il_shift = 0;                            // one tile per CTA: plain boxes, accumulator lane = tile row
    a.pool_out = nullptr; a.colsum_out = nullptr; a.stats_out = nullptr;   // nothing fused here: the caller runs its own kernels
    return (BN == 128) ? launch_tc_tap<128, TO, TMSK, TF32>(am, bm, a, mtiles, st)
                       : launch_tc_tap<64, TO, TMSK, TF32>(am, bm, a, mtiles, st);
}

// Try to run a tap-GEMM described in SIMT terms on the tensor cores.  Returns 1 if launched, 0 if the shape
// does not qualify (caller falls back to the CUDA-core kernel), or a negative mg_status on error.
template <typename TA, typename TO, typename TMSK>
int try_tc_tapgemm(const TapGemmArgs& P, cudaStream_t st);

static inline int split_blocks(long long n4) {
    const long long b = (n4 + 255) / 256, cap = (long long)num_sms() * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// float32 tap-GEMM on the bf16 tensor cores (see split3 above): the activation is re-written once as six bf16 k-segments
// per row (12 bytes per element of scratch), the weights are packed as the matching six segments, and the bf16 kernels run
// with K' = 6 K.  Exact to O(2^-24) per product plus the fp32 accumulation order of the tensor core.
template <typename TO, typename TMSK>
int try_split_tapgemm(const TapGemmArgs& P, cudaStream_t st) {
    const long long rows = (long long)P.B * P.Mper;
    if (P.w_split || rows < 1024 || P.K % 64 || P.N % 64 || P.row_scale || P.ntaps < 1 || P.ntaps > kMaxTaps || !is_pow2(P.Mper) ||
        (P.Mper > 128 && P.Mper % 128))
        return 0;
    if (P.Mper > 1 && P.a_mstride != P.K && P.a_mstride != 2 * P.K) return 0;
    if (P.a_valid % P.K || P.a_bstride != P.a_valid || ((uintptr_t)P.A) % 16 || ((uintptr_t)P.Out) % 16) return 0;
    if (P.o_off % 8 || P.o_mstride % 8 || P.o_bstride % 8) return 0;
    for (int t = 0; t < P.ntaps; ++t)
        if (P.a_toff[t] % P.K) return 0;
    const long long nA = (long long)P.B * P.a_bstride;
    if (nA * 6 >= (1LL << 40) || 6LL * P.a_bstride >= (1LL << 31) || 6LL * P.ntaps * P.N * P.K >= (1LL << 31)) return 0;
    __nv_bfloat16* a6 = static_cast<__nv_bfloat16*>(split_scratch(0, (size_t)nA * 6 * sizeof(__nv_bfloat16), st));
    if (!a6) return 0;
    split3_kernel<kSplitA><<<split_blocks(nA / 4), 256, 0, st>>>(static_cast<const float4*>(P.A), a6, nA / 4, (long long)P.K);
    MG_LAUNCH_OK();
    TapGemmArgs Q = P;
    Q.A = a6; Q.K = 6 * P.K; Q.w_split = 1;
    Q.a_bstride = 6 * P.a_bstride; Q.a_mstride = 6 * P.a_mstride; Q.a_valid = 6 * P.a_valid;
    for (int t = 0; t < P.ntaps; ++t) Q.a_toff[t] = 6 * P.a_toff[t];
    // the fused reductions (pooling, column sums, BatchNorm statistics) are float32 atomics: the parity mode keeps the
    // callers' deterministic passes (their *_done flags stay 0)
    Q.pool_out = nullptr; Q.pool_done = nullptr; Q.pool_only = 0; Q.colsum_out = nullptr; Q.colsum_done = nullptr;
    Q.stats_out = nullptr; Q.stats_done = nullptr;
    const int rc = try_tc_tapgemm<__nv_bfloat16, TO, TMSK>(Q, st);
    if (rc == 1) {
        LaunchInfo& li = last_launch();
        li.fp32x6 = 1; li.K = P.K;
        li.flops /= 6.0;                      // the layer's float32 FLOPs, not the bf16 ones issued
    }
    return rc;
}

template <typename TA, typename TO, typename TMSK>
int try_tc_tapgemm(const TapGemmArgs& P, cudaStream_t st) {
    constexpr bool TF32 = std::is_same<TA, float>::value;
    constexpr int EB = TF32 ? 4 : 2, KT = TF32 ? 32 : 64;          // operand element bytes, elements per 128-byte k-block
    if (!enabled()) return 0;
    if (TF32 && !(tf32_enabled() && P.Mper == 1 && P.ntaps == 1 && P.B >= 128)) {   // TF32: the fp32 Linears of bf16 mode only
        if (fp32_tc_enabled() && !tf32_enabled()) return try_split_tapgemm<TO, TMSK>(P, st);
        return 0;
    }
    if (P.K % KT || P.N % 64 || P.row_scale || P.ntaps < 1 || P.ntaps > kMaxTaps) return 0;
    if (!is_pow2(P.Mper) || (P.Mper > 128 && P.Mper % 128)) return 0;
    int stride;
    if (P.Mper == 1) stride = 1;
    else if (P.a_mstride == P.K) stride = 1;
    else if (P.a_mstride == 2 * P.K) stride = 2;
    else return 0;
    if (P.a_valid % P.K || P.a_bstride != P.a_valid) return 0;
    const int LA = P.a_valid / P.K;                         // rows per sample of the A tensor
    if (stride == 2 && (LA % 2)) return 0;
    if (((uintptr_t)P.A) % 16 || ((uintptr_t)P.Out) % 16) return 0;
    if (P.o_off % 8 || P.o_mstride % 8 || P.o_bstride % 8) return 0;
    TcTapArgs a{};
    a.ntaps = P.ntaps; a.kblocks = P.K / KT; a.ktile = KT;
    for (int t = 0; t < P.ntaps; ++t) {
        if (P.a_toff[t] % P.K) return 0;
        const int r = P.a_toff[t] / P.K;                    // row shift in A rows
        if (stride == 1) { a.a_p[t] = 0; a.a_dm[t] = r; }
        else { const int p = ((r % 2) + 2) % 2; a.a_p[t] = p; a.a_dm[t] = (r - p) / 2; }
        a.b_row[t] = t * P.N;
    }
    a.mpt = P.Mper >= 128 ? 128 : P.Mper; a.bpt = 128 / a.mpt;
    a.Mper = P.Mper; a.B = P.B; a.N = P.N;
    a.Out = P.Out; a.o_bstride = P.o_bstride; a.o_mstride = P.o_mstride; a.o_off = P.o_off;
    a.bias = P.bias; a.col_scale = P.col_scale; a.act = P.act; a.mul_src = P.mul_src; a.mul_mode = P.mul_mode;
    a.aux = P.aux; a.alpha = P.alpha; a.accumulate = P.accumulate;
    a.kb_outer = P.w_split;
    a.n_perm_q = P.n_perm_q; a.n_perm_p = P.n_perm_p;
    a.pool_out = P.pool_out; a.pool_scale = P.pool_scale; a.pool_done = P.pool_done; a.skip_out = 0;
    a.colsum_out = P.colsum_out; a.colsum_tiles = (int)(P.colsum_rows / 128); a.colsum_done = P.colsum_done;
    a.stats_out = P.stats_out; a.stats_done = P.stats_done;

    // pack the weight taps [ntaps][N][K] (bf16, or fp32 for the TF32 path): into the packed-weight cache when the host has
    // promised that weights only change through mg_adam_step (melogan.trainer), else into the next scratch slot
    Scratch& sc = scratch();
    const size_t need = (size_t)P.ntaps * P.N * P.K;
    PackArgs pk{};
    for (int t = 0; t < P.ntaps; ++t) pk.w_toff[t] = P.w_toff[t];
    pk.W = P.W; pk.w_nstride = P.w_nstride; pk.w_kstride = P.w_kstride; pk.n_perm_q = P.n_perm_q; pk.n_perm_p = P.n_perm_p;
    pk.k_perm_q = P.k_perm_q; pk.k_perm_p = P.k_perm_p; pk.ntaps = P.ntaps; pk.N = P.N; pk.K = P.K; pk.f32 = TF32 ? 1 : 0;
    pk.split = P.w_split;
    bool packed = false;
    __nv_bfloat16* wp = static_cast<__nv_bfloat16*>(cache_lookup(pk, need * EB, &packed));
    if (!wp && P.w_split) {                 // six-segment weights do not fit the ring slots a context sized for bf16 weights
        wp = static_cast<__nv_bfloat16*>(split_scratch(1, need * EB, st));
        if (!wp) return 0;
    }
    if (!wp) {
        if (need * EB > sc.slot_elems * 2) return 0;
        wp = sc.slot[sc.next];
        sc.next = (sc.next + 1) & 3;
    }
    if (!packed) {
        pk.out = wp;
        long long blocks = ((long long)need + 255) / 256;
        if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
        pack_weight_kernel<<<(int)blocks, 256, 0, st>>>(pk);
        MG_LAUNCH_OK();
    }

    CUtensorMap am, bm;
    int BN = (P.N % 128 == 0) ? 128 : 64;
    if (BN == 128) {   // a 128-wide slab whose taps cannot stay resident next to 3 activation stages: take 64-wide slabs if
                       // THOSE can (the activation is then read once per slab, but no weight tile is ever re-fetched)
        const size_t avail = (size_t)227 * 1024 - 1024 - kWsHeaderBytes - 3 * 17408;
        const size_t w128 = (size_t)P.ntaps * (P.K / KT) * 128 * 128;
        const long long mt = ((long long)P.B * P.Mper + 127) / 128;
        static const bool narrow = getenv("MELOGAN_WS_NO_NARROW") == nullptr;
        // ... unless a CTA pair can take the 128-wide slab (half of it per CTA)
        const bool pair_fits = !TF32 && pair_enabled() && !tuning().no_pair && mt % 2 == 0 && !P.accumulate && tma_store_enabled() &&
                               !tuning().no_tma_store && w128 / 2 + (size_t)128 * 128 * sizeof(TO) * (P.aux ? 2 : 1) <= avail;
        if (narrow && w128 > avail && w128 / 2 <= avail && mt >= 4LL * (num_sms() / (P.N / 64)) && !pair_fits) BN = 64;
    }
    if (tuning().force_bn == 64 || (tuning().force_bn == 128 && P.N % 128 == 0)) BN = tuning().force_bn;
    int rc = make_act_map(&am, P.A, P.K, P.Mper == 1 ? 1 : LA, P.B, stride, a.mpt, a.bpt, EB);
    if (rc != MG_OK) return rc;
    rc = make_weight_map(&bm, wp, P.K, (long long)P.ntaps * P.N, BN, EB);
    if (rc != MG_OK) return rc;
    CUtensorMap amh;
    const CUtensorMap* halo_map = nullptr;
    if (reuse_enabled() && !tuning().no_reuse && P.ntaps > 1 && P.Mper > 1) {
        build_tap_groups(a, 8);
        if (a.ngroups < a.ntaps) {
            // one sample per tile: a box of 128 + halo rows; several samples per tile (Mper < 128): the same rows of all bpt
            // samples row-interleaved, so that one row shift addresses every sample of the tile (TcTapArgs::il)
            if (a.bpt == 1) rc = make_act_map(&amh, P.A, P.K, LA, P.B, stride, 128 + a.halo, 1, EB);
            else { rc = make_act_map_interleaved(&amh, P.A, P.K, LA, P.B, stride, a.mpt + a.halo, a.bpt, EB); a.il = a.bpt; }
            if (rc != MG_OK) return rc;
            halo_map = &amh;
        } else {
            build_tap_groups(a, 0);
        }
    }
    CUtensorMap bmh;
    const CUtensorMap* half_map = nullptr;
    if (BN == 128 && !TF32 && pair_enabled()) {          // boxes of 64 weight rows: each CTA of a pair keeps half a slab
        rc = make_weight_map(&bmh, wp, P.K, (long long)P.ntaps * P.N, 64, EB);
        if (rc != MG_OK) return rc;
        half_map = &bmh;
    }
    rc = run_tc_tap<TO, TMSK, TF32>(am, bm, a, BN, P.K, st, halo_map, half_map, P.pool_only != 0);
    return rc == MG_OK ? 1 : rc;
}

// fp32 -> bf16 copy (operands of the float32 Linears' weight gradients in bf16 mode, see try_tc_wgrad)
static __global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ y, long long n4) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const float4 v = __ldg(x + i);
        const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        y[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    }
}
// growable device scratch for those copies (allocated by the first eager call, never inside a stream capture)
__nv_bfloat16* wgrad_cast_scratch(size_t elems, cudaStream_t st);

template <typename TG, typename TA>
int try_tc_wgrad(const WgradArgs& P, cudaStream_t st);

// float32 weight gradient on the bf16 tensor cores: the reduction runs over the rows, so the six part pairs are stacked along
// the samples (G: h h m h l m, A: h m h l h m; six bf16 copies of each tensor in scratch) and tc_wgrad_kernel reduces over
// 6 x the rows into the same fp32 gradient.
inline thread_local int t_wgrad_wave_mult = 1;   // set around the inner call of try_split_wgrad (ONE variable for all
                                                 // translation units: the templates that read it are merged by the linker)
static inline int try_split_wgrad(const WgradArgs& P, cudaStream_t st) {
    const long long nrows = (long long)P.row_end - P.row_begin;
    if (P.row_begin != 0 || nrows < 1024 || nrows % P.Mper || P.K % 64 || P.N % 128 || P.g_off || P.ntaps < 1 || P.ntaps > kMaxTaps ||
        !is_pow2(P.Mper) || (P.Mper > 64 && P.Mper % 64))
        return 0;
    if (P.Mper > 1 && (P.g_mstride != P.N || P.g_bstride != (long long)P.Mper * P.N)) return 0;
    if (P.Mper == 1 && P.g_bstride != P.N) return 0;
    if (P.Mper > 1 && P.a_mstride != P.K && P.a_mstride != 2 * P.K) return 0;
    if (P.a_valid % P.K || P.a_bstride != P.a_valid || ((uintptr_t)P.A) % 16 || ((uintptr_t)P.G) % 16) return 0;
    const long long nb = nrows / P.Mper, nG = nrows * P.N, nA = nb * P.a_bstride;
    if (6 * nrows >= (1LL << 31)) return 0;
    __nv_bfloat16* g6 = static_cast<__nv_bfloat16*>(split_scratch(2, (size_t)nG * 6 * sizeof(__nv_bfloat16), st));
    __nv_bfloat16* a6 = static_cast<__nv_bfloat16*>(split_scratch(0, (size_t)nA * 6 * sizeof(__nv_bfloat16), st));
    if (!g6 || !a6) return 0;
    split3_kernel<kSplitA><<<split_blocks(nG / 4), 256, 0, st>>>(static_cast<const float4*>(P.G), g6, nG / 4, nG);
    MG_LAUNCH_OK();
    split3_kernel<kSplitW><<<split_blocks(nA / 4), 256, 0, st>>>(static_cast<const float4*>(P.A), a6, nA / 4, nA);
    MG_LAUNCH_OK();
    WgradArgs Q = P;
    Q.G = g6; Q.A = a6; Q.B = (int)(6 * nb); Q.row_end = (int)(6 * nrows);
    // six times the rows: six times the splits, so that one CTA's TMEM accumulation chain (whose truncation error grows with its
    // length) is as long as in bf16 mode and mostly covers ONE part pair; the partial sums meet in fp32 atomics
    t_wgrad_wave_mult = 6;
    const int rc = try_tc_wgrad<__nv_bfloat16, __nv_bfloat16>(Q, st);
    t_wgrad_wave_mult = 1;
    if (rc == 1) {
        LaunchInfo& li = last_launch();
        li.fp32x6 = 1; li.rows = nrows; li.flops /= 6.0;
    }
    return rc;
}

template <typename TG, typename TA>
int try_tc_wgrad(const WgradArgs& P, cudaStream_t st) {
    if (std::is_same<TG, float>::value && std::is_same<TA, float>::value) {
        if (enabled() && fp32_tc_enabled() && !tf32_enabled()) return try_split_wgrad(P, st);
        // bf16 mode: the weight gradients of the float32 Linears (critic fc.1, the generator's / encoder's MLPs) ran on the CUDA
        // cores (critic fc.1: 122 us for 3.2 GFLOP).  Their forward and dgrad already read the operands as TF32 (10 mantissa
        // bits); here both operands are rounded to bf16 copies (8 bits, like every conv's weight gradient) and the reduction over
        // the rows runs on the tensor cores (kind::tf32 has no MN-major form, which a reduction over rows needs).
        const long long nrows = (long long)P.row_end - P.row_begin;
        if (!enabled() || !tf32_enabled() || P.Mper != 1 || P.ntaps != 1 || P.row_begin != 0 || nrows < 1024 || P.K % 64 || P.N % 128 ||
            P.g_off || P.g_bstride != P.N || P.a_bstride != P.K || P.a_valid != P.K || P.a_toff[0] != 0 ||
            ((uintptr_t)P.A) % 16 || ((uintptr_t)P.G) % 16)
            return 0;
        const size_t ng = (size_t)nrows * P.N, na = (size_t)nrows * P.K;
        __nv_bfloat16* buf = wgrad_cast_scratch(ng + na, st);
        if (!buf) return 0;
        const int blocks_g = (int)std::min<long long>((ng / 4 + 255) / 256, (long long)num_sms() * 8);
        const int blocks_a = (int)std::min<long long>((na / 4 + 255) / 256, (long long)num_sms() * 8);
        f32_to_bf16_kernel<<<blocks_g, 256, 0, st>>>(reinterpret_cast<const float4*>(P.G), reinterpret_cast<uint2*>(buf), (long long)(ng / 4));
        MG_LAUNCH_OK();
        f32_to_bf16_kernel<<<blocks_a, 256, 0, st>>>(reinterpret_cast<const float4*>(P.A), reinterpret_cast<uint2*>(buf + ng), (long long)(na / 4));
        MG_LAUNCH_OK();
        WgradArgs Q = P;
        Q.G = buf; Q.A = buf + ng;
        return try_tc_wgrad<__nv_bfloat16, __nv_bfloat16>(Q, st);
    }
    if (!std::is_same<TG, __nv_bfloat16>::value || !std::is_same<TA, __nv_bfloat16>::value || !enabled()) return 0;
    if (P.K % 64 || P.N % 128 || P.ntaps < 1 || P.ntaps > kMaxTaps || P.g_off) return 0;
    if (!is_pow2(P.Mper) || (P.Mper > 64 && P.Mper % 64)) return 0;
    int stride;
    if (P.Mper == 1) stride = 1;
    else if (P.a_mstride == P.K) stride = 1;
    else if (P.a_mstride == 2 * P.K) stride = 2;
    else return 0;
    if (P.a_valid % P.K || P.a_bstride != P.a_valid) return 0;
    const int LA = P.a_valid / P.K;
    if (stride == 2 && (LA % 2)) return 0;
    if (P.Mper > 1 && (P.g_mstride != P.N || P.g_bstride != (long long)P.Mper * P.N)) return 0;
    if (P.Mper == 1 && P.g_bstride != P.N) return 0;
    if (((uintptr_t)P.A) % 16 || ((uintptr_t)P.G) % 16) return 0;
    if (P.row_begin % 64) return 0;
    TcWgradArgs a{};
    a.ntaps = P.ntaps; a.K = P.K; a.N = P.N;
    for (int t = 0; t < P.ntaps; ++t) {
        if (P.a_toff[t] % P.K) return 0;
        const int r = P.a_toff[t] / P.K;
        if (stride == 1) { a.a_p[t] = 0; a.a_dm[t] = r; }
        else { const int p = ((r % 2) + 2) % 2; a.a_p[t] = p; a.a_dm[t] = (r - p) / 2; }
        a.w_toff[t] = P.w_toff[t];
    }
    a.rpt = P.Mper >= 64 ? 64 : P.Mper; a.spt = 64 / a.rpt; a.Mper = P.Mper;
    a.row_begin = P.row_begin; a.row_end = P.row_end;
    a.dW = P.dW; a.w_nstride = P.w_nstride; a.w_kstride = P.w_kstride; a.n_perm_q = P.n_perm_q; a.n_perm_p = P.n_perm_p;
    a.alpha = P.alpha;
    const long long nrows = P.row_end - P.row_begin;
    if (nrows <= 0) return 1;
    const int BNK = (P.K % 128 == 0) ? 128 : 64;
    const int tiles = (P.N / 128) * P.ntaps * (P.K / BNK);
    // one full wave: as many CTAs as the GPU holds at once (2 per SM at BNK = 128, 3 at 64, by shared memory), never one
    // more -- a 445th CTA on 444 slots runs alone after everybody else and doubles the kernel
    const size_t smem_cta = (BNK == 128 ? sizeof(WgradSmem<128>) : sizeof(WgradSmem<64>)) + 1024;
    const long long cap = (long long)num_sms() * (long long)((227 * 1024) / smem_cta);
    static const int waves = getenv("MELOGAN_WGRAD_WAVES") ? atoi(getenv("MELOGAN_WGRAD_WAVES")) : 1;
    long long want = cap * (waves > 0 ? waves : 1) * t_wgrad_wave_mult / tiles;
    long long maxs = (nrows + 255) / 256;                    // at least 4 chunks per CTA
    long long splits = want < 1 ? 1 : (want > maxs ? maxs : want);
    if (splits > 65535) splits = 65535;
    long long rps = (nrows + splits - 1) / splits;
    rps = (rps + 63) / 64 * 64;
    splits = (nrows + rps - 1) / rps;
    a.rows_per_split = (int)rps;
    // samples covered by the maps: rows are (b, m) with Mper rows per sample
    const long long nsamples = (P.row_end + P.Mper - 1) / P.Mper;
    CUtensorMap gm, am;
    int rc = make_act_map(&gm, P.G, P.N, P.Mper, nsamples, 1, a.rpt, a.spt);
    if (rc != MG_OK) return rc;
    rc = make_act_map(&am, P.A, P.K, P.Mper == 1 ? 1 : LA, nsamples, stride, a.rpt, a.spt);
    if (rc != MG_OK) return rc;
    ProbeScope probe(PROBE_TC_WGRAD, 2.0 * (double)nrows * P.N * P.ntaps * P.K, (double)nrows * (P.N + P.K) * 2.0, st);
    {
        LaunchInfo& li = last_launch();
        li = LaunchInfo();
        li.kind = 2; li.rows = nrows; li.N = P.N; li.K = P.K; li.taps = P.ntaps; li.BN = BNK; li.splits = (int)splits;
        li.flops = 2.0 * (double)nrows * P.N * P.ntaps * P.K; li.bytes = (double)nrows * (P.N + P.K) * 2.0;
    }
    rc = (BNK == 128) ? launch_tc_wgrad<128>(gm, am, a, (int)splits, st) : launch_tc_wgrad<64>(gm, am, a, (int)splits, st);
    return rc == MG_OK ? 1 : rc;
}

}  // namespace tc
}  // namespace mg
