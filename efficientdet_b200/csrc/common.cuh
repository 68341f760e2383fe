// Shared helpers for the effdet_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <stdlib.h>
#include <utility>

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

// ---- programmatic dependent launch (PDL).  A kernel launched through launch_pdl() may be scheduled while
//      the previous kernel of the stream is still draining; it MUST execute EFFDET_PDL_SYNC() before its
//      first global-memory access (the wait returns once the previous grid has completed and its writes
//      are visible; the trigger then lets the next PDL kernel be scheduled behind this one).  Captured
//      CUDA graphs keep these edges, which removes most of the kernel-to-kernel launch gap of the chains
//      of small kernels in a step.  EFFDET_NO_PDL=1 falls back to plain stream order.
// Levels (EFFDET_PDL_LEVEL, default 2): 1 = tensor-core convolutions, 2 = + TMA depthwise family, SE, weight
// panels, weight gradients, 3 = + the element-wise / reduction kernels of the training step (measured:
// level 2 is +5 % on the D0 training step and +11 % on batch-1 inference; level 3 costs 5 % on D4
// training, where early-resident dependents take SM slots from the big memory-bound passes).
#ifndef EFFDET_PDL_TU_LEVEL
#define EFFDET_PDL_TU_LEVEL 3
#endif
inline int pdl_level() {
    static const int lvl = getenv("EFFDET_NO_PDL") ? 0 : (getenv("EFFDET_PDL_LEVEL") ? atoi(getenv("EFFDET_PDL_LEVEL")) : 2);
    return lvl;
}
#define EFFDET_PDL_SYNC() asm volatile("griddepcontrol.wait;\n\tgriddepcontrol.launch_dependents;" ::: "memory")
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args &&...args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool on = pdl_level() >= EFFDET_PDL_TU_LEVEL;
    cfg.attrs = attr; cfg.numAttrs = on ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

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

// ---- batched 16-byte loads for the streaming (HBM-bound) kernels.
// ptxas sinks every independent global load next to its first use, whatever the source / PTX order: an unrolled
// "load 4 rows, then reduce them" loop comes out as LDG, math, LDG, math ... with one or two loads in flight per
// thread, ~16 KB in flight per SM and ~3.5 TB/s (ncu: long-scoreboard stalls 11.6 per issue, DRAM 50 %).
// tie_loads() ORs every loaded word with (xor of ALL words of the batch) & zero, where `zero` is a kernel
// ARGUMENT that is always 0: the values are unchanged, but each of them now depends on all loads of the batch,
// so the assembler has to issue the whole batch before the first use (measured: 3.5 -> 4.7 TB/s on the
// BatchNorm backward passes).
__device__ __forceinline__ uint4 ld16(const void *p) { return *reinterpret_cast<const uint4 *>(p); }
template <int N> __device__ __forceinline__ void tie_loads(uint4 (&r)[N], uint32_t zero) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) t ^= r[i].x ^ r[i].y ^ r[i].z ^ r[i].w;
    t &= zero;
#pragma unroll
    for (int i = 0; i < N; ++i) { r[i].x |= t; r[i].y |= t; r[i].z |= t; r[i].w |= t; }
}
template <typename T, int CV> struct Unpack16;
template <> struct Unpack16<float, 4> {
    static __device__ __forceinline__ void run(const uint4 &t, float *v) {
        v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
    }
};
template <> struct Unpack16<__nv_bfloat16, 8> {
    static __device__ __forceinline__ void run(const uint4 &t, float *v) {
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
        v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
        v[4] = __uint_as_float(t.z << 16); v[5] = __uint_as_float(t.z & 0xffff0000u);
        v[6] = __uint_as_float(t.w << 16); v[7] = __uint_as_float(t.w & 0xffff0000u);
    }
};

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
