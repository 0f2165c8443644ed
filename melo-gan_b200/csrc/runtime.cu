// runtime.cu -- error reporting, device queries, version info.
#include <stdarg.h>
#include <string.h>

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
}  // namespace mg

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
