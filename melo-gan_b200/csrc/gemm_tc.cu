// gemm_tc.cu -- host helpers of the tcgen05 kernels: TMA tensor-map encoding, packed-weight scratch ring.
#include <stdlib.h>

#include <vector>

#include "gemm_tc.cuh"

namespace mg {
namespace tc {

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
Scratch g_scratch;
}  // namespace

Scratch& scratch() { return g_scratch; }
// ---- packed-weight cache ----
namespace {
struct CacheEntry {
    PackArgs key;
    void* buf = nullptr;
    size_t bytes = 0;
    bool current = false;
};
std::vector<CacheEntry> g_cache;
thread_local bool g_cache_on = false;   // set per API call from the context (set_cache_mode)
bool same_key(const PackArgs& a, const PackArgs& b) {
    if (a.W != b.W || a.ntaps != b.ntaps || a.N != b.N || a.K != b.K || a.f32 != b.f32 || a.split != b.split || a.w_nstride != b.w_nstride ||
        a.w_kstride != b.w_kstride || a.n_perm_q != b.n_perm_q || a.n_perm_p != b.n_perm_p || a.k_perm_q != b.k_perm_q ||
        a.k_perm_p != b.k_perm_p)
        return false;
    for (int t = 0; t < a.ntaps; ++t)
        if (a.w_toff[t] != b.w_toff[t]) return false;
    return true;
}
}  // namespace

void* cache_lookup(const PackArgs& key, size_t bytes, bool* is_current) {
    *is_current = false;
    if (!g_cache_on) return nullptr;
    for (CacheEntry& e : g_cache)
        if (same_key(e.key, key) && e.bytes >= bytes) {
            *is_current = e.current;
            e.current = true;              // the caller packs now (stream-ordered before its consumer) if it was stale
            return e.buf;
        }
    CacheEntry e;
    e.key = key; e.bytes = bytes;
    if (cudaMalloc(&e.buf, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }   // e.g. under stream capture
    e.current = true;
    g_cache.push_back(e);
    return e.buf;
}

void weights_changed(const float* param, long long n) {
    for (CacheEntry& e : g_cache)
        if (!param || (e.key.W >= param && e.key.W < param + n)) e.current = false;
}

void set_cache_mode(bool on) { g_cache_on = on; }

static Tuning g_tuning;
Tuning& tuning() { return g_tuning; }
static thread_local LaunchInfo g_last_launch;
LaunchInfo& last_launch() { return g_last_launch; }

int ensure_scratch(size_t elems) {
    if (elems <= g_scratch.slot_elems) return MG_OK;
    // Grow by allocating NEW slots; the old ones stay alive until the process ends: CUDA graphs captured by an earlier, smaller
    // context (GanTrainer.capture_cycle) have the old slot addresses baked in and keep replaying on them.
    static std::vector<__nv_bfloat16*> retired;
    __nv_bfloat16* fresh[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < 4; ++i) {
        if (cudaMalloc(&fresh[i], elems * sizeof(__nv_bfloat16)) != cudaSuccess) {
            cudaGetLastError();
            for (int j = 0; j < i; ++j) cudaFree(fresh[j]);
            set_error("packed-weight scratch: cudaMalloc of %zu bytes failed", elems * sizeof(__nv_bfloat16));
            return MG_ERR_CUDA;
        }
    }
    for (int i = 0; i < 4; ++i) {
        if (g_scratch.slot[i]) retired.push_back(g_scratch.slot[i]);
        g_scratch.slot[i] = fresh[i];
    }
    g_scratch.slot_elems = elems;
    return MG_OK;
}

static int g_enabled = -1;
bool enabled() {
    if (g_enabled < 0) {
        const char* e = getenv("MELOGAN_DISABLE_TC");
        g_enabled = (e && e[0] == '1') ? 0 : 1;
    }
    return g_enabled == 1;
}
bool ws_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MELOGAN_DISABLE_WS");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}
__nv_bfloat16* wgrad_cast_scratch(size_t elems, cudaStream_t st) {
    // Grows by allocating a NEW buffer; the old ones stay alive until the process ends: CUDA graphs captured earlier have
    // their addresses baked in.  A growth request inside a stream capture fails (cudaMalloc is not capturable): the caller
    // then keeps the CUDA-core kernel for that launch -- callers run one eager step before they capture.
    static std::vector<__nv_bfloat16*> kept;
    static __nv_bfloat16* buf = nullptr;
    static size_t cap = 0;
    if (elems > cap) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (cs != cudaStreamCaptureStatusNone) return nullptr;      // (an allocation would invalidate the capture)
        __nv_bfloat16* nb = nullptr;
        if (cudaMalloc(&nb, elems * sizeof(__nv_bfloat16)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (buf) kept.push_back(buf);
        buf = nb;
        cap = elems;
    }
    return buf;
}
void* split_scratch(int kind, size_t bytes, cudaStream_t st) {
    static std::vector<void*> kept;
    static void* buf[3] = {nullptr, nullptr, nullptr};
    static size_t cap[3] = {0, 0, 0};
    if (kind < 0 || kind > 2) return nullptr;
    if (bytes > cap[kind]) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (cs != cudaStreamCaptureStatusNone) return nullptr;
        void* nb = nullptr;
        if (cudaMalloc(&nb, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (buf[kind]) kept.push_back(buf[kind]);
        buf[kind] = nb;
        cap[kind] = bytes;
    }
    return buf[kind];
}
bool fp32_tc_enabled() {
    static const bool env_on = getenv("MELOGAN_FP32_TC") != nullptr && getenv("MELOGAN_FP32_TC")[0] == '1';
    const int t = tuning().fp32_tc;
    return t >= 0 ? t != 0 : env_on;
}
bool pair_enabled() {
    static const bool on = getenv("MELOGAN_DISABLE_PAIR") == nullptr;
    return on;
}
bool mask_tma_enabled() {
    static const bool on = getenv("MELOGAN_DISABLE_TMA_MASK") == nullptr;
    return on;
}
bool tma_store_enabled() {
    static const bool on = getenv("MELOGAN_DISABLE_TMA_STORE") == nullptr;
    return on;
}
int make_out_map(CUtensorMap* map, const void* base, int elem_bytes, long long cols, long long rows, long long row_stride) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return MG_ERR_CUDA; }
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)row_stride * (cuuint64_t)elem_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), 128};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                    const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(out cols=%lld rows=%lld stride=%lld) failed: %d", cols, rows, row_stride, (int)r);
        return MG_ERR_CUDA;
    }
    return MG_OK;
}
bool reuse_enabled() {
    static const bool on = getenv("MELOGAN_DISABLE_TAP_REUSE") == nullptr;
    return on;
}
int set_enabled(int on) {
    const int prev = enabled() ? 1 : 0;
    g_enabled = on ? 1 : 0;
    return prev;
}

static bool g_tf32 = false;
bool tf32_enabled() {
    static const bool off = getenv("MELOGAN_DISABLE_TF32") != nullptr;
    return g_tf32 && !off;
}
void set_tf32(bool on) { g_tf32 = on; }

int make_act_map(CUtensorMap* map, const void* base, int C, int L, long long B, int stride, int box_rows,
                 int box_samples, int elem_bytes) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return MG_ERR_CUDA; }
    const cuuint64_t row_bytes = (cuuint64_t)C * (cuuint64_t)elem_bytes;
    cuuint64_t gdim[4], gstr[3];
    if (stride == 1) {
        gdim[0] = (cuuint64_t)C; gdim[1] = 1; gdim[2] = (cuuint64_t)L; gdim[3] = (cuuint64_t)B;
        gstr[0] = row_bytes; gstr[1] = row_bytes; gstr[2] = row_bytes * (cuuint64_t)L;
    } else {
        gdim[0] = (cuuint64_t)C; gdim[1] = 2; gdim[2] = (cuuint64_t)(L / 2); gdim[3] = (cuuint64_t)B;
        gstr[0] = row_bytes; gstr[1] = 2 * row_bytes; gstr[2] = row_bytes * (cuuint64_t)L;
    }
    const cuuint32_t box[4] = {(cuuint32_t)(128 / elem_bytes), 1, (cuuint32_t)box_rows, (cuuint32_t)box_samples};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(act C=%d L=%d B=%lld stride=%d box=%d,%d) failed: %d", C, L, B, stride, box_rows,
                  box_samples, (int)r);
        return MG_ERR_CUDA;
    }
    return MG_OK;
}

int make_act_map_interleaved(CUtensorMap* map, const void* base, int C, int L, long long B, int stride, int box_rows,
                             int box_samples, int elem_bytes) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return MG_ERR_CUDA; }
    const cuuint64_t row_bytes = (cuuint64_t)C * (cuuint64_t)elem_bytes;
    cuuint64_t gdim[4], gstr[3];
    gdim[0] = (cuuint64_t)C; gdim[1] = (cuuint64_t)B;
    gstr[0] = row_bytes * (cuuint64_t)L;                     // sample
    if (stride == 1) { gdim[2] = 1; gdim[3] = (cuuint64_t)L; gstr[1] = row_bytes; gstr[2] = row_bytes; }
    else { gdim[2] = 2; gdim[3] = (cuuint64_t)(L / 2); gstr[1] = row_bytes; gstr[2] = 2 * row_bytes; }
    const cuuint32_t box[4] = {(cuuint32_t)(128 / elem_bytes), (cuuint32_t)box_samples, 1, (cuuint32_t)box_rows};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(interleaved act C=%d L=%d B=%lld stride=%d box=%d,%d) failed: %d", C, L, B, stride,
                  box_rows, box_samples, (int)r);
        return MG_ERR_CUDA;
    }
    return MG_OK;
}

int make_view_map(CUtensorMap* map, const void* base, long long inner, long long planes, long long plane_stride,
                  long long rows, long long row_stride, long long samples, long long sample_stride, int box_rows,
                  int box_samples) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return MG_ERR_CUDA; }
    const cuuint64_t gdim[4] = {(cuuint64_t)inner, (cuuint64_t)planes, (cuuint64_t)rows, (cuuint64_t)samples};
    const cuuint64_t gstr[3] = {(cuuint64_t)plane_stride * 2, (cuuint64_t)row_stride * 2, (cuuint64_t)sample_stride * 2};
    const cuuint32_t box[4] = {64, 1, (cuuint32_t)box_rows, (cuuint32_t)box_samples};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(view inner=%lld planes=%lld rows=%lld samples=%lld) failed: %d", inner, planes,
                  rows, samples, (int)r);
        return MG_ERR_CUDA;
    }
    return MG_OK;
}

int make_weight_map(CUtensorMap* map, const void* base, int K, long long rows, int box_rows, int elem_bytes) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return MG_ERR_CUDA; }
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)K * (cuuint64_t)elem_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(weight K=%d rows=%lld box=%d) failed: %d", K, rows, box_rows, (int)r);
        return MG_ERR_CUDA;
    }
    return MG_OK;
}

}  // namespace tc
void tc_weights_changed(const float* param, long long n) { tc::weights_changed(param, n); }
}  // namespace mg

// A/B switch between the tcgen05 kernels and the CUDA-core kernels in bf16 mode (tests, profiling).
extern "C" int mg_tc_enable(int on) { return mg::tc::set_enabled(on); }
extern "C" int mg_weight_cache_invalidate(void) { mg::tc::weights_changed(nullptr, 0); return MG_OK; }
