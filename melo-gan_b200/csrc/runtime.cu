// runtime.cu -- error reporting, device queries, version info.
#include <stdarg.h>
#include <string.h>

#include <utility>
#include <vector>

#include "common.cuh"

namespace mg {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

static long long g_launches = 0;
void count_launch() { ++g_launches; }

// ---- probe ----
static int g_probe_family = PROBE_NONE;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_probe_events;
static double g_probe_flops = 0.0, g_probe_bytes = 0.0;
static size_t g_probe_used = 0;

ProbeScope::ProbeScope(int family, double flops, double bytes, cudaStream_t stream) : st(stream) {
    if (family != g_probe_family || g_probe_family == PROBE_NONE) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
    if (g_probe_used == g_probe_events.size()) {
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
        g_probe_events.push_back({a, b});
    }
    cudaEventRecord(g_probe_events[g_probe_used].first, stream);
    g_probe_flops += flops; g_probe_bytes += bytes;
    on = true;
}
ProbeScope::~ProbeScope() {
    if (!on) return;
    cudaEventRecord(g_probe_events[g_probe_used].second, st);
    ++g_probe_used;
}
}  // namespace mg

extern "C" long long mg_launch_count(void) { return mg::g_launches; }

extern "C" int mg_probe_begin(int family) {
    mg::g_probe_family = family;
    mg::g_probe_used = 0;
    mg::g_probe_flops = mg::g_probe_bytes = 0.0;
    return MG_OK;
}

// out[0] = launches, out[1] = total ms, out[2] = total algorithmic flops, out[3] = total algorithmic bytes
extern "C" int mg_probe_end(double* out) {
    MG_REQUIRE(out, "probe_end: null pointer");
    double ms = 0.0;
    for (size_t i = 0; i < mg::g_probe_used; ++i) {
        MG_CUDA_OK(cudaEventSynchronize(mg::g_probe_events[i].second));
        float t = 0.f;
        MG_CUDA_OK(cudaEventElapsedTime(&t, mg::g_probe_events[i].first, mg::g_probe_events[i].second));
        ms += t;
    }
    out[0] = (double)mg::g_probe_used; out[1] = ms; out[2] = mg::g_probe_flops; out[3] = mg::g_probe_bytes;
    mg::g_probe_family = mg::PROBE_NONE;
    return MG_OK;
}

extern "C" const char* mg_last_error(void) { return mg::g_err; }
extern "C" int mg_abi_version(void) { return 1; }
extern "C" const char* mg_build_info(void) { return "melogan_b200 sm_100a (nvcc " __DATE__ ")"; }

extern "C" uint32_t mg_scale_mask(const char* scale_name, int root_key) {
    // interval tables: reference src/gan/utils.py:14-26
    struct S { const char* name; int n; int iv[12]; };
    static const S scales[] = {
        {"major", 7, {0, 2, 4, 5, 7, 9, 11}},       {"minor", 7, {0, 2, 3, 5, 7, 8, 10}},
        {"chromatic", 12, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11}},
        {"dorian", 7, {0, 2, 3, 5, 7, 9, 10}},      {"phrygian", 7, {0, 1, 3, 5, 7, 8, 10}},
        {"lydian", 7, {0, 2, 4, 6, 7, 9, 11}},      {"mixolydian", 7, {0, 2, 4, 5, 7, 9, 10}},
        {"locrian", 7, {0, 1, 3, 5, 6, 8, 10}},     {"major_pentatonic", 5, {0, 2, 4, 7, 9}},
        {"minor_pentatonic", 5, {0, 3, 5, 7, 10}},  {"blues", 6, {0, 3, 5, 6, 7, 10}},
    };
    const S* s = &scales[2];
    if (scale_name)
        for (const S& c : scales)
            if (strcmp(c.name, scale_name) == 0) { s = &c; break; }
    uint32_t m = 0;
    for (int i = 0; i < s->n; ++i) m |= 1u << ((((s->iv[i] + root_key) % 12) + 12) % 12);
    return m;
}

extern "C" int mg_device_copy(void* dst, const void* src, long long nbytes, void* stream) {
    MG_REQUIRE(nbytes >= 0 && (nbytes == 0 || (dst && src)), "device_copy: bad arguments");
    if (nbytes == 0) return MG_OK;
    MG_CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)nbytes, cudaMemcpyDefault, mg::as_stream(stream)));
    return MG_OK;
}
