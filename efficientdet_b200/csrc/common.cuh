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

// ---- storage-type aware activations: bf16 tensors take the one-MUFU tanh forms (error ~2^-11 absolute,
//      below the bf16 rounding of the stored result); fp32 tensors keep the exact forms
__device__ __forceinline__ float tanh_fast(float x) {
    float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
template <typename T> __device__ __forceinline__ float activate_io(float x, int act) { return activate_rt(x, act); }
template <> __device__ __forceinline__ float activate_io<__nv_bfloat16>(float x, int act) {
    switch (act) {
        case EFFDET_ACT_RELU: return fmaxf(x, 0.f);
        case EFFDET_ACT_SWISH: { const float h = 0.5f * x; return fmaf(h, tanh_fast(h), h); }
        case EFFDET_ACT_SIGMOID: return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f);
        default: return x;
    }
}
// derivative of the activation applied to u, as a factor on the incoming gradient
template <typename T> __device__ __forceinline__ float act_grad_io(float u, int act) {
    if (act == EFFDET_ACT_RELU) return u > 0.f ? 1.f : 0.f;
    if (act == EFFDET_ACT_SWISH) {
        const float s = 1.f / (1.f + __expf(-u));
        return s * (1.f + u * (1.f - s));
    }
    return 1.f;
}
template <> __device__ __forceinline__ float act_grad_io<__nv_bfloat16>(float u, int act) {
    if (act == EFFDET_ACT_RELU) return u > 0.f ? 1.f : 0.f;
    if (act == EFFDET_ACT_SWISH) {
        const float s = fmaf(0.5f, tanh_fast(0.5f * u), 0.5f);
        return s * fmaf(u, 1.f - s, 1.f);
    }
    return 1.f;
}

}  // namespace effdet
