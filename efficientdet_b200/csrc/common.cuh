// Shared helpers for the effdet_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "../../include/effdet_b200.h"

namespace effdet {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char *fmt, const char *a = "", long long b = 0, long long c = 0) {
    snprintf(g_err, sizeof(g_err), fmt, a, b, c);
    return code;
}

#define EFFDET_REQUIRE(cond, msg)                                                       \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            snprintf(::effdet::g_err, sizeof(::effdet::g_err), "%s: %s (%s)", __func__, msg, \
                     #cond);                                                            \
            return EFFDET_E_INVALID;                                                    \
        }                                                                               \
    } while (0)

#define EFFDET_CUDA(expr)                                                               \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            snprintf(::effdet::g_err, sizeof(::effdet::g_err), "%s: %s -> %s", __func__, #expr, \
                     cudaGetErrorString(_e));                                           \
            return EFFDET_E_CUDA;                                                       \
        }                                                                               \
    } while (0)

// call after every kernel launch: counts it and surfaces launch-configuration errors
#define EFFDET_LAUNCHED()                                                               \
    do {                                                                                \
        ::effdet::g_launches.fetch_add(1, std::memory_order_relaxed);                   \
        cudaError_t _e = cudaPeekAtLastError();                                         \
        if (_e != cudaSuccess) {                                                        \
            snprintf(::effdet::g_err, sizeof(::effdet::g_err), "%s: kernel launch -> %s", \
                     __func__, cudaGetErrorString(_e));                                 \
            (void)cudaGetLastError();                                                   \
            return EFFDET_E_CUDA;                                                       \
        }                                                                               \
    } while (0)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

constexpr int kNumSMs = 148;

// ---- activation storage types: float or bf16, math always in fp32
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) {
    return __bfloat162float(v);
}
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) {
    return __float2bfloat16_rn(v);
}

template <int ACT> __device__ __forceinline__ float activate(float x) {
    if (ACT == EFFDET_ACT_RELU) return fmaxf(x, 0.f);
    if (ACT == EFFDET_ACT_SWISH) return x / (1.f + __expf(-x));
    if (ACT == EFFDET_ACT_SIGMOID) return 1.f / (1.f + __expf(-x));
    return x;
}
__device__ __forceinline__ float activate_rt(float x, int act) {
    switch (act) {
        case EFFDET_ACT_RELU: return fmaxf(x, 0.f);
        case EFFDET_ACT_SWISH: return x / (1.f + __expf(-x));
        case EFFDET_ACT_SIGMOID: return 1.f / (1.f + __expf(-x));
        default: return x;
    }
}

}  // namespace effdet
