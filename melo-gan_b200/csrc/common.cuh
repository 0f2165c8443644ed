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

#define MG_LAUNCH_OK()  MG_CUDA_OK(cudaGetLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace mg
