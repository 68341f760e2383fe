// TMA / mbarrier PTX wrappers and the host-side tensor-map encoder shared by the TMA-fed kernels
// (conv_tc.cu, dwconv_tma.cu).  sm_100a only.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace effdet {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t a, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(a), "r"(parity) : "memory");
    return done != 0;
}
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes or
// ~hint_ns elapse, so a waiting role does not burn issue slots (ncu: the polling loops of the
// producer / MMA / starved epilogue warps were ~18 % of all executed instructions of the
// HBM-bound layers, which are issue-limited in the epilogue).
__device__ __forceinline__ bool mbar_try_suspend(uint32_t a, uint32_t parity, uint32_t hint_ns) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(a), "r"(parity), "r"(hint_ns) : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    if (mbar_try(a, parity)) return;
    while (!mbar_try_suspend(a, parity, 2000u)) { }
}
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity, unsigned ns) {
    const uint32_t a = smem_u32(bar);
    if (mbar_try(a, parity)) return;
    while (!mbar_try_suspend(a, parity, 4000u)) { }
    (void)ns;
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0,
                                            int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
          "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0,
                                            int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
          "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, const void *src, int c0, int c1, int c2,
                                             int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
        ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// at most n of this thread's bulk stores may still be READING their shared-memory source
// same, source given as a 32-bit shared-window address
__device__ __forceinline__ void tma_store_4d_s(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
        ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// One lane of a converged warp.  Unlike `lane == 0` this tells the compiler that exactly one thread runs the
// guarded code, so the operands of TMA / tcgen05 instructions (uniform registers) are produced on the uniform
// datapath instead of a per-lane "waterfall" loop (R2UR.OR + BRA.U.ANY, ~22 instructions per TMA store).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// generic-proxy shared-memory writes -> visible to the async proxy (TMA)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------ host: cuTensorMapEncodeTiled via the runtime
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
// Stand-in used ONLY when no CUDA driver can be loaded and EFFDET_DRY_RUN is set (the CPU test that walks the
// launch list of every model size through the entry points' host side, tests/test_plan_hazards.py): it checks the
// documented argument rules of cuTensorMapEncodeTiled for the tiled, non-interleaved maps this library builds and
// writes nothing.  The launch that follows fails (no driver), so nothing ever consumes such a map.
static inline CUresult dry_run_encode_tiled(CUtensorMap *map, CUtensorMapDataType dt, cuuint32_t rank, void *addr,
                                            const cuuint64_t *dims, const cuuint64_t *strides, const cuuint32_t *box,
                                            const cuuint32_t *estr, CUtensorMapInterleave il, CUtensorMapSwizzle sw,
                                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
    const size_t es = dt == CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 ? 2 : dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 4 : 0;
    if (!map || !addr || es == 0 || rank < 1 || rank > 5 || il != CU_TENSOR_MAP_INTERLEAVE_NONE)
        return CUDA_ERROR_INVALID_VALUE;
    if (reinterpret_cast<uintptr_t>(addr) & 15) return CUDA_ERROR_INVALID_VALUE;
    for (cuuint32_t i = 0; i < rank; ++i) {
        if (dims[i] == 0 || dims[i] > (1ull << 32)) return CUDA_ERROR_INVALID_VALUE;
        if (box[i] == 0 || box[i] > 256) return CUDA_ERROR_INVALID_VALUE;
        if (estr[i] == 0 || estr[i] > 8) return CUDA_ERROR_INVALID_VALUE;
        if (i + 1 < rank && (strides[i] % 16 != 0 || strides[i] == 0 || strides[i] >= (1ull << 40)))
            return CUDA_ERROR_INVALID_VALUE;
    }
    const size_t inner = (size_t)box[0] * es;
    if (inner % 16 != 0) return CUDA_ERROR_INVALID_VALUE;
    const size_t span = sw == CU_TENSOR_MAP_SWIZZLE_128B ? 128 : sw == CU_TENSOR_MAP_SWIZZLE_64B ? 64 :
                        sw == CU_TENSOR_MAP_SWIZZLE_32B ? 32 : (size_t)-1;
    if (inner > span) return CUDA_ERROR_INVALID_VALUE;
    return CUDA_SUCCESS;
}

static inline EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
        else if (getenv("EFFDET_DRY_RUN"))
            fn = dry_run_encode_tiled;
    }
    return fn;
}


}  // namespace effdet
