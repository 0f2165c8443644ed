// peer.cu -- the data-parallel gradient exchange (the path's one exchange step) as kernels over NVLink peer
// memory instead of an NCCL call: in-place sum over the ranks of one node (<= 8 GPUs, one process per GPU).
//
// Every rank owns an exchange region (cudaMalloc, shared through a CUDA IPC handle): two parities of max_floats floats, then
// a control block {flag, epoch}.  One all-reduce = three launches on the caller's stream:
//   stage    copy the local vector into parity (epoch + 1) & 1 of my region               (many CTAs, 16-byte accesses)
//   publish  epoch += 1; flag = epoch with a system-scope release                          (one thread)
//   reduce   every CTA waits until each peer's flag reached the epoch, then sums, element by element and in RANK order,
//            my local vector and the peers' staged copies read over NVLink (ld.volatile), in place
// From four ranks on the reduce step is split (reduce-scatter + all-gather): rank r reduces only slice r -- reading (N-1)/N of
// the vector from its peers instead of N-1 times the vector --, writes the sums into slice r of its own staged copy,
// publishes a second flag, and every rank then fetches the other slices from their owners: 2(N-1)/N vector reads per rank.
// Rank-order sums make the result bit-identical on all ranks (the parameters stay identical without a broadcast).  Two
// parities suffice: a rank stages epoch e + 2 only after every peer published e + 1, i.e. after they finished reducing e.
// No host synchronisation and no NCCL: the three launches are ordinary graph nodes, so a data-parallel training cycle is ONE
// CUDA graph again.  Reference: the all-reduce torch DDP would add to src/gan/train_gan.py:203,247 (loss.backward()).
#include <cstring>

#include "common.cuh"

using namespace mg;

namespace {
constexpr int kMaxRanks = 8;
struct Peers { float* base[kMaxRanks]; int world, rank; long long max_floats; };

__device__ __forceinline__ unsigned* ctrl_of(float* base, long long max_floats) {
    return reinterpret_cast<unsigned*>(base + 2 * max_floats);
}

__global__ void __launch_bounds__(256) peer_stage_kernel(const Peers P, const float* __restrict__ data, long long n) {
    float* mine = P.base[P.rank];
    const unsigned e = ctrl_of(mine, P.max_floats)[1] + 1u;
    float* dst = mine + (long long)(e & 1u) * P.max_floats;
    const long long n4 = n >> 2, stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride)
        reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(data) + i);
    for (long long i = (n4 << 2) + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) dst[i] = data[i];
}

// control block: c[0] = flag seen by the peers (2 * epoch - 1 after staging, 2 * epoch after the reduce-scatter), c[1] = epoch
__global__ void peer_publish_kernel(const Peers P, int phase) {
    unsigned* c = ctrl_of(P.base[P.rank], P.max_floats);
    unsigned e = c[1];
    if (phase == 0) { e += 1u; c[1] = e; }
    __threadfence_system();
    const unsigned f = 2u * e - (phase == 0 ? 1u : 0u);
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(c), "r"(f) : "memory");
}

__device__ __forceinline__ void peer_wait(const Peers& P, unsigned want) {     // threads 0..world-1 of a CTA, then a barrier
    if ((int)threadIdx.x < P.world && (int)threadIdx.x != P.rank) {
        const unsigned* pf = ctrl_of(P.base[threadIdx.x], P.max_floats);
        unsigned seen = 0;
        for (unsigned spins = 0;; ++spins) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(pf) : "memory");
            if ((int)(seen - want) >= 0) break;
            if (spins > (1u << 27)) asm volatile("trap;");             // a lost peer must not hang the GPU
            __nanosleep(100);
        }
    }
    __syncthreads();
}

// reduce-scatter: this rank sums slice `rank` (n4s float4 per slice) in rank order and leaves it in data AND in its staged copy
__global__ void __launch_bounds__(256) peer_rs_kernel(const Peers P, float* __restrict__ data, long long n4s, long long n4) {
    const unsigned e = ctrl_of(P.base[P.rank], P.max_floats)[1];
    peer_wait(P, 2u * e - 1u);
    const long long par = (long long)(e & 1u) * P.max_floats;
    const long long lo = (long long)P.rank * n4s, hi = lo + n4s < n4 ? lo + n4s : n4;
    float4* mine = reinterpret_cast<float4*>(P.base[P.rank] + par);
    for (long long i = lo + (long long)blockIdx.x * 256 + threadIdx.x; i < hi; i += (long long)gridDim.x * 256) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < P.world; ++r) {
            const float4 v = (r == P.rank) ? reinterpret_cast<const float4*>(data)[i]
                                           : __ldcv(reinterpret_cast<const float4*>(P.base[r] + par) + i);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        reinterpret_cast<float4*>(data)[i] = s;
        mine[i] = s;
    }
}
// all-gather: fetch every other slice from its owner's staged copy
__global__ void __launch_bounds__(256) peer_ag_kernel(const Peers P, float* __restrict__ data, long long n4s, long long n4) {
    const unsigned e = ctrl_of(P.base[P.rank], P.max_floats)[1];
    peer_wait(P, 2u * e);
    const long long par = (long long)(e & 1u) * P.max_floats;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const int r = (int)(i / n4s);
        if (r != P.rank) reinterpret_cast<float4*>(data)[i] = __ldcv(reinterpret_cast<const float4*>(P.base[r] + par) + i);
    }
}

__global__ void __launch_bounds__(256) peer_reduce_kernel(const Peers P, float* __restrict__ data, long long n) {
    const unsigned e = ctrl_of(P.base[P.rank], P.max_floats)[1];
    peer_wait(P, 2u * e - 1u);
    const long long par = (long long)(e & 1u) * P.max_floats;
    const long long n4 = n >> 2, stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < P.world; ++r) {
            const float4 v = (r == P.rank) ? reinterpret_cast<const float4*>(data)[i]
                                           : __ldcv(reinterpret_cast<const float4*>(P.base[r] + par) + i);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        reinterpret_cast<float4*>(data)[i] = s;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        float s = 0.f;
        for (int r = 0; r < P.world; ++r) s += (r == P.rank) ? data[i] : __ldcv(P.base[r] + par + i);
        data[i] = s;
    }
}
}  // namespace

struct mg_peer {
    Peers P{};
    void* opened[kMaxRanks] = {};
    bool connected = false;
};

extern "C" int mg_peer_create(int rank, int world, long long max_floats, mg_peer** out, unsigned char* handle_out) {
    MG_REQUIRE(out && handle_out && world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world && max_floats > 0,
               "peer_create: bad arguments (world <= %d)", kMaxRanks);
    max_floats = (max_floats + 3) / 4 * 4;
    mg_peer* p = new mg_peer();
    p->P.world = world; p->P.rank = rank; p->P.max_floats = max_floats;
    const size_t bytes = (size_t)(2 * max_floats) * 4 + 64;
    float* buf = nullptr;
    if (cudaMalloc(&buf, bytes) != cudaSuccess || cudaMemset(buf, 0, bytes) != cudaSuccess) {
        set_error("peer_create: %s", cudaGetErrorString(cudaGetLastError()));
        delete p;
        return MG_ERR_CUDA;
    }
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, buf) != cudaSuccess) {
        set_error("peer_create: cudaIpcGetMemHandle: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(buf);
        delete p;
        return MG_ERR_CUDA;
    }
    memcpy(handle_out, &h, 64);
    p->P.base[rank] = buf;
    *out = p;
    return MG_OK;
}

extern "C" int mg_peer_connect(mg_peer* p, const unsigned char* handles) {
    MG_REQUIRE(p && handles && !p->connected, "peer_connect: null or already connected");
    for (int r = 0; r < p->P.world; ++r) {
        if (r == p->P.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void* q = nullptr;
        MG_CUDA_OK(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
        p->opened[r] = q;
        p->P.base[r] = static_cast<float*>(q);
    }
    p->connected = true;
    return MG_OK;
}

extern "C" int mg_peer_allreduce_sum(mg_peer* p, float* data, long long n, void* stream) {
    MG_REQUIRE(p && data && n > 0, "peer_allreduce_sum: null or empty");
    MG_REQUIRE(p->connected || p->P.world == 1, "peer_allreduce_sum: call mg_peer_connect first");
    MG_REQUIRE(n <= p->P.max_floats, "peer_allreduce_sum: %lld floats exceed the exchange region (%lld)", n, p->P.max_floats);
    MG_REQUIRE(((uintptr_t)data) % 16 == 0, "peer_allreduce_sum: the vector must be 16-byte aligned");
    if (p->P.world == 1) return MG_OK;
    cudaStream_t st = as_stream(stream);
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = (long long)num_sms() * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    peer_stage_kernel<<<(unsigned)blocks, 256, 0, st>>>(p->P, data, n);
    MG_LAUNCH_OK();
    peer_publish_kernel<<<1, 1, 0, st>>>(p->P, 0);
    MG_LAUNCH_OK();
    if (p->P.world >= 4 && n % 4 == 0 && n >= 4096LL * p->P.world) {       // reduce-scatter + all-gather
        const long long n4 = n / 4, n4s = (n4 + p->P.world - 1) / p->P.world;
        peer_rs_kernel<<<(unsigned)blocks, 256, 0, st>>>(p->P, data, n4s, n4);
        MG_LAUNCH_OK();
        peer_publish_kernel<<<1, 1, 0, st>>>(p->P, 1);
        MG_LAUNCH_OK();
        peer_ag_kernel<<<(unsigned)blocks, 256, 0, st>>>(p->P, data, n4s, n4);
        MG_LAUNCH_OK();
    } else {                                                                // every rank reads every peer's whole copy
        peer_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(p->P, data, n);
        MG_LAUNCH_OK();
    }
    return MG_OK;
}

extern "C" void mg_peer_destroy(mg_peer* p) {
    if (!p) return;
    for (int r = 0; r < kMaxRanks; ++r)
        if (p->opened[r]) cudaIpcCloseMemHandle(p->opened[r]);
    if (p->P.base[p->P.rank]) cudaFree(p->P.base[p->P.rank]);
    delete p;
}
