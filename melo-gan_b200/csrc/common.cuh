// common.cuh -- shared host/device helpers for the melogan_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/melogan_b200.h"

namespace mg {

void set_error(const char* fmt, ...);
int num_sms();
void count_launch();   // every kernel launch of this library goes through MG_LAUNCH_OK()
void tc_weights_changed(const float* param, long long n);   // gemm_tc.cu: invalidates packed-weight cache entries

// Optional per-kernel-family probe (bench.py roofline): CUDA events around each launch of one family.
enum ProbeFamily { PROBE_NONE = 0, PROBE_TAPGEMM = 1, PROBE_WGRAD = 2, PROBE_TC_GEMM = 3, PROBE_TC_WGRAD = 4,
                   PROBE_NOTES = 5, PROBE_ADAM = 6, PROBE_ELEM = 7 };
struct ProbeScope {
    bool on = false;
    cudaStream_t st;
    ProbeScope(int family, double flops, double bytes, cudaStream_t stream);
    ~ProbeScope();
};

#define MG_CUDA_OK(expr)                                                                    \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            mg::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return MG_ERR_CUDA;                                                             \
        }                                                                                   \
    } while (0)

#define MG_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            mg::set_error(__VA_ARGS__);       \
            return MG_ERR_INVALID;            \
        }                                     \
    } while (0)

#define MG_TRY(expr)                    \
    do {                                \
        int _rc = (expr);               \
        if (_rc != MG_OK) return _rc;   \
    } while (0)

#define MG_LAUNCH_OK()                      \
    do {                                    \
        MG_CUDA_OK(cudaGetLastError());     \
        mg::count_launch();                 \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace mg
