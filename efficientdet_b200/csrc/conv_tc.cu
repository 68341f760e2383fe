// Tensor-core convolution for sm_100a: implicit GEMM on tcgen05.mma with the accumulator in
// TMEM, operands staged in shared memory by TMA (SWIZZLE_128B), warp-specialised
// producer / MMA-issuer / epilogue roles synchronised with mbarriers.
//
//   D[M = 128 pixels, N = Cout tile] += A[128 pixels, 64 channels] * B[Cout tile, 64 channels]^T
//
// A tile  = one TMA box of a 4-D NHWC tensor map (C, W, H, B): box (64, Wt, Ht, Bt), Wt*Ht*Bt = 128.
//           For a 3x3 convolution the box origin is shifted by the tap offset; TMA's out-of-bounds
//           zero fill IS the TF "SAME" zero padding, so there is no im2col buffer.  The K loop
//           runs over taps x 64-channel blocks.
// B tile  = TMA box (64, BLOCK_N, 1) of the bf16 weight panel (Kpad, Npad, taps | samples).
// Epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> scale/shift (folded BN or bias),
//           activation, ReLU-mask (data-gradient use), drop-connect scale + residual ->
//           bf16 / fp32 stores at per-group row / image strides (concat-free head outputs).
//
// Replaces the cuDNN calls behind efficientnet.py:228-237/289-304 (1x1 expand / project),
// model.py:71-90 (BiFPN 1x1 laterals), model.py:293-309/324-351 (3x3 head convs, all five pyramid
// levels in one launch) and their data gradients.
#include <cuda.h>
#include <stdlib.h>

#define EFFDET_PDL_TU_LEVEL 2
#include "common.cuh"
#include "tma.cuh"

namespace effdet {

// ------------------------------------------------------------------ tcgen05 wrappers (TMA / mbarrier ones: tma.cuh)
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"):
// rows are 128-byte lines (64 bf16), 8-row groups are 1024 bytes apart (SBO), LBO = 1 (unused).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

constexpr int kTcMaxGroups = 5;
constexpr int kTileM = 128, kTileK = 64;
constexpr int kATileBytes = kTileM * kTileK * 2;   // 16 KiB

struct TcGroup {
    void *y;
    const void *res;
    const void *mask;
    long long y_batch_stride;
    int ldc;
    int H, W;
    int Wt, Ht, Bt;           // tile extent (Wt*Ht*Bt == 128)
    int tiles_x, tiles_y, tiles_b;
    int tile_begin;
    int tma_store;            // output rows are 16-byte strided: store through out_map
    int lw, lh;               // log2(Wt), log2(Ht): tile extents are powers of two
};
struct alignas(64) TcParams {
    CUtensorMap a_map[kTcMaxGroups];
    CUtensorMap out_map[kTcMaxGroups];
    CUtensorMap b_map;
    TcGroup g[kTcMaxGroups];
    int n_groups, B, Cout, ksize, kblocks_per_tap, block_n, stages, tmem_cols;
    int split, kb_plane, plane_stride;   // split-bf16 fp32 mode: K blocks per plane segment, channels between planes
    int act, out_f32, b_per_sample, total_tiles, n_tiles, any_tma_store;
    // halo mode (3x3, stride 1, Cin <= 64, weights resident): one ring stage = three kx-shifted copies of
    // the activation patch with a halo row above / below; tap (ky, kx) is the 128-pixel range of copy kx
    // that starts ky rows in (see conv_wgrad_halo_kernel).  60 KiB instead of 9 x 16 KiB per tile.
    int halo, a_stage_bytes, box_stride, stage_bufs, box_bytes[kTcMaxGroups];
    int acc_bufs;        // TMEM accumulators per CTA: 2 (epilogue of tile i overlaps the MMAs of i+1) or 1
    int b_resident;      // all K blocks of the (single) weight tile stay in shared memory for the CTA's lifetime
    int stride, pad_t[kTcMaxGroups], pad_l[kTcMaxGroups];   // stride 2: TMA element strides sample every other pixel
    const float *scale, *shift, *keep;
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct TileCoord { int gi, x0, y0, b0, n0; };
// Each CTA walks a CONTIGUOUS range of tiles; after the one-time decode the coordinates advance
// by increments (no integer divisions in the per-tile path of the three roles).
// fp32 pair -> packed bf16x2, round-half-away (differs from RNE only on exact ties).  Three
// ALU-pipe instructions instead of one F2FP: F2FP shares the 16-lane/clk XU pipe with MUFU.TANH,
// which is the co-limiter of the swish epilogue (ncu: xu pipe 43 % busy, MUFU only 2/3 of it).
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
}

struct TileIter {
    int gi, tx, ty, tb, nt;
    __device__ __forceinline__ void init(const TcParams &p, int t) {
        const int m = t / p.n_tiles;
        nt = t - m * p.n_tiles;
        gi = 0;
#pragma unroll
        for (int i = 1; i < kTcMaxGroups; ++i)
            if (i < p.n_groups && m >= p.g[i].tile_begin) gi = i;
        const TcGroup &G = p.g[gi];
        int r = m - G.tile_begin;
        tx = r % G.tiles_x; r /= G.tiles_x;
        ty = r % G.tiles_y; tb = r / G.tiles_y;
    }
    __device__ __forceinline__ void next(const TcParams &p) {
        if (++nt < p.n_tiles) return;
        nt = 0;
        const TcGroup &G = p.g[gi];
        if (++tx < G.tiles_x) return;
        tx = 0;
        if (++ty < G.tiles_y) return;
        ty = 0;
        if (++tb < G.tiles_b) return;
        tb = 0; ++gi;
    }
    __device__ __forceinline__ TileCoord coord(const TcParams &p) const {
        const TcGroup &G = p.g[gi];
        TileCoord c;
        c.gi = gi; c.x0 = tx * G.Wt; c.y0 = ty * G.Ht; c.b0 = tb * G.Bt; c.n0 = nt * p.block_n;
        return c;
    }
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float tanh_approx(float x) {
    float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
// HALF_IN: the caller passes z/2 (the 1/2 is folded into the per-channel scale / shift), so that
// swish(z) = z * sigmoid(z) = (z/2) * (1 + tanh(z/2)) costs one MUFU and one FMA.  The special
// function unit is the co-limiter of the HBM-bound expand convolutions (one output element per
// 2 bytes written); bf16 outputs only -- tanh.approx is good to ~2^-11 absolute.
template <int ACT, bool HALF_IN> __device__ __forceinline__ float fast_act(float x) {
    if (ACT == EFFDET_ACT_RELU) return fmaxf(x, 0.f);
    if (ACT == EFFDET_ACT_SWISH) {
        if (HALF_IN) return fmaf(x, tanh_approx(x), x);
        return x * rcp_approx(1.f + ex2_approx(-1.4426950408889634f * x));
    }
    if (ACT == EFFDET_ACT_SIGMOID) return rcp_approx(1.f + ex2_approx(-1.4426950408889634f * x));
    return x;
}

// Packed fp32 FMA (FFMA2): two IEEE fma.rn results per issue slot -- bit-identical to two FFMAs.  The
// epilogue is issue-bound (r3a capture: ~390 warp-instructions per 32-column chunk), not FMA-pipe bound.
__device__ __forceinline__ void ffma2(float &d0, float &d1, float a0, float a1, float b0, float b1, float c0,
                                      float c1) {
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}

constexpr int kEpiWarps = 8;                 // two warps per TMEM lane quarter, alternating column chunks
constexpr int kTcThreads = 64 + 32 * kEpiWarps;

// Persistent: each CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  Barriers and TMEM
// are set up once; two TMEM accumulators let the epilogue of tile i overlap the MMAs of tile i+1.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..9 = epilogue.
template <int ACT, bool OUT_F32>
__global__ void __launch_bounds__(kTcThreads, 2)      // <= 96 registers: two CTAs (16 epilogue warps) per SM
conv_tc_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    // (pointer arithmetic on the __shared__ array, not through uintptr_t, keeps LDS/STS addressing)
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int b_tile_bytes = p.block_n * kTileK * 2;
    uint8_t *sA = smem;
    uint8_t *sB = smem + (size_t)p.stages * p.a_stage_bytes;
    // ring B tiles, or (b_resident) one slot per K block filled once
    const int num_k_all = p.ksize * p.ksize * p.kblocks_per_tap;
    float *sScale = reinterpret_cast<float *>(sB + (size_t)(p.b_resident ? num_k_all : p.stages) * b_tile_bytes);   // [2][256]
    float *sShift = sScale + 512;                                                          // [2][256] each: see the epilogue
    uint64_t *full = reinterpret_cast<uint64_t *>(sShift + 512);
    uint64_t *empty = full + p.stages;
    uint64_t *tmem_full = empty + p.stages;      // [2]
    uint64_t *tmem_empty = tmem_full + 2;        // [2]
    uint64_t *b_full = tmem_empty + 2;           // [1] resident weights landed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(b_full + 1);
    // epilogue staging (TMA-store path): per warp two [32 rows][32 columns] swizzled buffers
    constexpr int kStageBytes = 32 * 32 * (OUT_F32 ? 4 : 2);
    uint8_t *sStage = reinterpret_cast<uint8_t *>(tmem_slot + 4);
    sStage += (1024u - (smem_u32(sStage) & 1023u)) & 1023u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int taps = p.ksize * p.ksize;
    const int num_k = taps * p.kblocks_per_tap;
    const int t_begin = (int)(((long long)blockIdx.x * p.total_tiles) / gridDim.x);
    const int t_end = (int)(((long long)(blockIdx.x + 1) * p.total_tiles) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], kEpiWarps); }
        mbar_init(b_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // Programmatic dependent launch: everything above (barriers, TMEM, tensor-map fetch) overlaps the
    // tail of the previous kernel in the stream; nothing below may touch global memory before the
    // previous grid has completed.  The next PDL kernel may start ITS prologue as soon as our CTAs
    // have reached this point and an SM has room for it.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t acc_cols = (uint32_t)p.tmem_cols / (uint32_t)p.acc_bufs;
    const int two_acc = p.acc_bufs == 2;

    if (warp == 0) {
        if (elect_one()) {
            // ===== TMA producer
            int it = 0;
            if (p.b_resident && t_begin < t_end) {
                mbar_expect_tx(b_full, (uint32_t)(num_k * b_tile_bytes));
                for (int kb = 0; kb < num_k; ++kb)
                    tma_load_3d(sB + (size_t)kb * b_tile_bytes, &p.b_map, b_full, (kb % p.kblocks_per_tap) * kTileK, 0,
                                kb / p.kblocks_per_tap);
            }
            TileIter ti_;
            ti_.init(p, t_begin);
            for (int t = t_begin; t < t_end; ++t, ti_.next(p)) {
                const TileCoord c = ti_.coord(p);
                const CUtensorMap *amap = &p.a_map[c.gi];
                if (p.halo) {
                    const int s = it % p.stages, ph = (it / p.stages) & 1;
                    mbar_wait_backoff(&empty[s], ph ^ 1, 256);
                    mbar_expect_tx(&full[s], (uint32_t)(3 * p.box_bytes[c.gi]));
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
                        tma_load_4d(sA + (size_t)s * p.a_stage_bytes + (size_t)kx * p.box_stride, amap, &full[s], 0,
                                    c.x0 - 1 + kx, c.y0 - 1, c.b0);
                    ++it;
                    continue;
                }
                for (int kb = 0; kb < num_k; ++kb, ++it) {
                    const int s = it % p.stages, ph = (it / p.stages) & 1;
                    mbar_wait_backoff(&empty[s], ph ^ 1, 256);
                    const int tap = kb / p.kblocks_per_tap, kbt = kb - tap * p.kblocks_per_tap, kc = kbt * kTileK;
                    const int ky = tap / p.ksize, kx = tap - ky * p.ksize;
                    // split-bf16 fp32 mode: the K extent is [hi | lo | hi] against [Whi | Whi | Wlo]; the activation
                    // tensor holds the two planes side by side (plane_stride channels apart)
                    int ka = kc;
                    if (p.split) {
                        const int seg = kbt / p.kb_plane;
                        ka = (kbt - seg * p.kb_plane) * kTileK + (seg == 1 ? p.plane_stride : 0);
                    }
                    mbar_expect_tx(&full[s], (uint32_t)(kATileBytes + (p.b_resident ? 0 : b_tile_bytes)));
                    tma_load_4d(sA + (size_t)s * kATileBytes, amap, &full[s], ka, c.x0 * p.stride + kx - p.pad_l[c.gi],
                                c.y0 * p.stride + ky - p.pad_t[c.gi], c.b0);
                    if (!p.b_resident)
                        tma_load_3d(sB + (size_t)s * b_tile_bytes, &p.b_map, &full[s], kc, c.n0,
                                    p.b_per_sample ? c.b0 : tap);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ===== MMA issuer (single thread)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) |
                                   ((uint32_t)(kTileM >> 4) << 24);
            int it = 0, ti = 0;
            if (p.b_resident && t_begin < t_end) { mbar_wait(b_full, 0); tc_fence_after(); }
            TileIter mit;
            mit.init(p, t_begin);
            for (int t = t_begin; t < t_end; ++t, ++ti, mit.next(p)) {
                const int acc = two_acc ? (ti & 1) : 0, aph = (two_acc ? (ti >> 1) : ti) & 1;
                mbar_wait(&tmem_empty[acc], aph ^ 1);        // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_cols;
                if (p.halo) {
                    const int s = it % p.stages, ph = (it / p.stages) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + (size_t)s * p.a_stage_bytes);
                    const uint32_t row_bytes = (uint32_t)p.g[mit.gi].Wt * 128u;
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        const int ky = tap / 3, kx = tap - ky * 3;
                        const uint64_t da = make_kmajor_sw128_desc(a0 + (uint32_t)kx * (uint32_t)p.box_stride + (uint32_t)ky * row_bytes);
                        const uint64_t db = make_kmajor_sw128_desc(smem_u32(sB + (size_t)tap * b_tile_bytes));
#pragma unroll
                        for (int k = 0; k < kTileK / 16; ++k)
                            umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (tap | k) ? 1u : 0u);
                    }
                    umma_commit(&empty[s]);
                    ++it;
                    umma_commit(&tmem_full[acc]);
                    continue;
                }
                for (int kb = 0; kb < num_k; ++kb, ++it) {
                    const int s = it % p.stages, ph = (it / p.stages) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint64_t da = make_kmajor_sw128_desc(smem_u32(sA + (size_t)s * kATileBytes));
                    const uint64_t db = make_kmajor_sw128_desc(smem_u32(sB + (size_t)(p.b_resident ? kb : s) * b_tile_bytes));
#pragma unroll
                    for (int k = 0; k < kTileK / 16; ++k)      // UMMA_K = 16 bf16 = 32 bytes: +2 in the address field
                        umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
                    umma_commit(&empty[s]);                     // frees the smem slot when the MMAs retire
                }
                umma_commit(&tmem_full[acc]);                   // accumulator complete
            }
        }
    } else {
        // ===== epilogue: 8 warps; warp w owns TMEM lane quarter (w % 4) and every other 32-column chunk
        const int ew = warp - 2;
        const int q = warp & 3;
        const int half = ew >> 2;                           // which of the two warps of this quarter
        const int row = q * 32 + lane;                      // tile row == TMEM lane
        const int et = threadIdx.x - 64;                    // 0..255
        constexpr bool kHalfIn = (ACT == EFFDET_ACT_SWISH) && !OUT_F32;
        int ti = 0, cur_n0 = -1, sbuf = 0, ss_par = 256;
        uint8_t *my_stage = sStage + (size_t)ew * p.stage_bufs * kStageBytes;
        const uint32_t stage_s = smem_u32(my_stage);
        // bf16 staging row of this lane: 64 bytes, 16-byte chunk index ^= (row >> 1) & 3 (SWIZZLE_64B); the
        // XOR with j8 << 4 touches only bits 4-5, so it commutes with adding the row offset
        const uint32_t st_row = (uint32_t)lane * 64u + ((((uint32_t)lane >> 1) & 3u) << 4);
        TileIter tit;
        tit.init(p, t_begin);
        for (int t = t_begin; t < t_end; ++t, ++ti, tit.next(p)) {
            const TileCoord tc = tit.coord(p);
            const TcGroup &G = p.g[tc.gi];
            const int x0 = tc.x0, y0 = tc.y0, b0 = tc.b0, n0 = tc.n0;
            const int acc = two_acc ? (ti & 1) : 0, aph = (two_acc ? (ti >> 1) : ti) & 1;
            // Per-column scale / shift of this N tile -> shared memory (uniform across the CTA).  With several N
            // tiles (n varies fastest) EVERY tile restages them: the loads are issued here, ahead of the wait for
            // the accumulator, and land in the buffer the previous tile did not use, so one barrier per tile
            // (after the accumulator wait) orders both the new values and the reuse of the other buffer.  (r3b
            // capture of block4b_expand: the former barrier -> global load -> barrier chain of each tile held a
            // quarter of the epilogue warps' samples.)
            const bool restage = n0 != cur_n0;
            float st_scale = 1.f, st_shift = 0.f;
            if (restage) {
                const int n = n0 + et;
                const float pre = kHalfIn ? 0.5f : 1.f;
                st_scale = pre * ((p.scale && n < p.Cout) ? p.scale[n] : 1.f);
                st_shift = pre * ((p.shift && n < p.Cout) ? p.shift[n] : 0.f);
            }
            // per-row addressing is only needed for direct stores / residual / mask reads: the TMA
            // store clips rows outside the tensor by itself
            const bool tma_out = G.tma_store != 0;
            const bool has_rm = G.res != nullptr || G.mask != nullptr;
            bool row_ok = true;
            size_t base = 0;
            float kp = 1.f;
            if (!tma_out || G.res || G.mask) {
                const int xx = row & (G.Wt - 1), yy = (row >> G.lw) & (G.Ht - 1), bb = row >> (G.lw + G.lh);
                const int x = x0 + xx, y = y0 + yy, b = b0 + bb;
                row_ok = x < G.W && y < G.H && b < p.B;
                base = (size_t)b * G.y_batch_stride + ((size_t)y * G.W + x) * G.ldc;
                kp = (p.keep && row_ok) ? p.keep[b] : 1.f;
            }
            // origin of this warp's 32-row box (rows 32q .. 32q+31 of the tile) for the TMA store
            const int r0 = q * 32;
            const int sx = x0 + (r0 & (G.Wt - 1)), sy = y0 + ((r0 >> G.lw) & (G.Ht - 1)), sb = b0 + (r0 >> (G.lw + G.lh));
            if (restage) {
                ss_par ^= 256;
                sScale[ss_par + et] = st_scale;
                sShift[ss_par + et] = st_shift;
                cur_n0 = n0;
            }
            mbar_wait(&tmem_full[acc], aph);
            tc_fence_after();
            if (restage) asm volatile("bar.sync 1, 256;" ::: "memory");
            const float *cScale = sScale + ss_par, *cShift = sShift + ss_par;
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_cols + ((uint32_t)(q * 32) << 16);
            // 32-column chunks alternate between the two warps of a lane quarter; the starting warp
            // flips every tile so that an odd chunk count (N = 96: 3 chunks) balances over two tiles
            // columns of this tile that exist (warp-uniform); the accumulator goes back to the MMA warp as soon
            // as this warp's LAST chunk is in registers, so the next tile's MMAs (and, with one accumulator, the
            // whole wait -> MMA -> commit round trip) overlap that chunk's math, packing and store.  r3a capture
            // of block3a_expand (one accumulator, two CTAs per SM): 23 % of the epilogue warps' samples were
            // waits for tmem_full.
            const int c_end = min(p.block_n, p.Cout - n0);
            const int c_first = ((half ^ ti) & 1) * 32;
            bool released = false;
            for (int c0 = c_first; c0 < c_end; c0 += 64) {
                const int nbase = n0 + c0;
                const int ncols = min(32, c_end - c0);      // warp-uniform, >= 1
                uint32_t r[32];
                __syncwarp();                               // tcgen05.ld is warp-collective
                tmem_ld32(d_tmem + (uint32_t)c0, r);
                if (c0 + 64 >= c_end) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                    released = true;
                }
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 sc = *reinterpret_cast<const float4 *>(cScale + c0 + j);
                    const float4 sh = *reinterpret_cast<const float4 *>(cShift + c0 + j);
                    ffma2(v[j], v[j + 1], __uint_as_float(r[j]), __uint_as_float(r[j + 1]), sc.x, sc.y, sh.x, sh.y);
                    ffma2(v[j + 2], v[j + 3], __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]), sc.z, sc.w, sh.z, sh.w);
                    if (ACT == EFFDET_ACT_SWISH && kHalfIn) {
                        // swish(z) = h + h * tanh(h), h = z / 2 (the 1/2 is folded into scale / shift)
                        const float t0 = tanh_approx(v[j]), t1 = tanh_approx(v[j + 1]);
                        const float t2 = tanh_approx(v[j + 2]), t3 = tanh_approx(v[j + 3]);
                        ffma2(v[j], v[j + 1], v[j], v[j + 1], t0, t1, v[j], v[j + 1]);
                        ffma2(v[j + 2], v[j + 3], v[j + 2], v[j + 3], t2, t3, v[j + 2], v[j + 3]);
                    } else {
                        v[j] = fast_act<ACT, kHalfIn>(v[j]);
                        v[j + 1] = fast_act<ACT, kHalfIn>(v[j + 1]);
                        v[j + 2] = fast_act<ACT, kHalfIn>(v[j + 2]);
                        v[j + 3] = fast_act<ACT, kHalfIn>(v[j + 3]);
                    }
                }
                const int nvalid = row_ok ? ncols : 0;
                if (OUT_F32) {
                    float *Y = static_cast<float *>(G.y) + base + nbase;
                    const float *R = G.res ? static_cast<const float *>(G.res) + base + nbase : nullptr;
                    const float *MK = G.mask ? static_cast<const float *>(G.mask) + base + nbase : nullptr;
                    if (R || MK) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (j < nvalid) {
                                if (MK && !(MK[j] > 0.f)) v[j] = 0.f;
                                if (R) v[j] = v[j] * kp + R[j];
                            }
                        }
                    }
                    if (tma_out) {
                        // [32 rows][128 B] SWIZZLE_128B: 16-byte chunk index ^= row & 7 (bank-conflict free)
                        if (p.stage_bufs == 2) tma_store_wait_read<1>();   // the store that last read this buffer is done
                        else tma_store_wait_read<0>();
                        __syncwarp();
                        uint8_t *dst = my_stage + (size_t)sbuf * kStageBytes + lane * 128;
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4)
                            *reinterpret_cast<float4 *>(dst + ((j4 ^ (lane & 7)) << 4)) =
                                make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
                        fence_proxy_async();
                        __syncwarp();
                        if (elect_one())
                            tma_store_4d(&p.out_map[tc.gi], my_stage + (size_t)sbuf * kStageBytes, nbase, sx, sy, sb);
                        sbuf = (sbuf + 1) & (p.stage_bufs - 1);
                    } else {
                        // rows whose byte stride is not a multiple of 16 (9*C floats, C = 90: the class head's
                        // concatenated output) cannot go through TMA: transpose through the warp's staging
                        // buffer and let the 32 lanes write 32 consecutive floats of one row per instruction
                        // (one 128-byte segment instead of 32 scattered 4-byte stores)
                        (void)Y;
                        __syncwarp();
                        uint8_t *dst = my_stage + lane * 128;
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4)
                            *reinterpret_cast<float4 *>(dst + ((j4 ^ (lane & 7)) << 4)) =
                                make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
                        __syncwarp();
                        const unsigned ok_mask = __ballot_sync(0xffffffffu, row_ok);
                        const unsigned long long row_off = (unsigned long long)base + (unsigned long long)nbase;
                        float *Yb = static_cast<float *>(G.y);
#pragma unroll 4
                        for (int rr = 0; rr < 32; ++rr) {
                            const unsigned long long off = __shfl_sync(0xffffffffu, row_off, rr);
                            if (((ok_mask >> rr) & 1u) && lane < ncols) {
                                const float val = *reinterpret_cast<const float *>(
                                    my_stage + rr * 128 + ((((lane >> 2) ^ (rr & 7)) << 4) | ((lane & 3) << 2)));
                                Yb[off + lane] = val;
                            }
                        }
                    }
                } else {
                    if (has_rm && nvalid > 0) {       // (pointer arithmetic only on this path: expand convolutions skip it)
                        const __nv_bfloat16 *R = G.res ? static_cast<const __nv_bfloat16 *>(G.res) + base + nbase : nullptr;
                        const __nv_bfloat16 *MK = G.mask ? static_cast<const __nv_bfloat16 *>(G.mask) + base + nbase : nullptr;
                        const bool vec = nvalid == 32 && (!R || (reinterpret_cast<uintptr_t>(R) & 15) == 0) &&
                                         (!MK || (reinterpret_cast<uintptr_t>(MK) & 15) == 0);
                        if (vec) {
#pragma unroll
                            for (int j8 = 0; j8 < 32; j8 += 8) {
                                if (MK) {
                                    uint4 mv = *reinterpret_cast<const uint4 *>(MK + j8);
                                    const __nv_bfloat162 *mh = reinterpret_cast<const __nv_bfloat162 *>(&mv);
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        if (!(__low2float(mh[j]) > 0.f)) v[j8 + 2 * j] = 0.f;
                                        if (!(__high2float(mh[j]) > 0.f)) v[j8 + 2 * j + 1] = 0.f;
                                    }
                                }
                                if (R) {
                                    uint4 rv = *reinterpret_cast<const uint4 *>(R + j8);
                                    const __nv_bfloat162 *rh = reinterpret_cast<const __nv_bfloat162 *>(&rv);
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        v[j8 + 2 * j] = v[j8 + 2 * j] * kp + __low2float(rh[j]);
                                        v[j8 + 2 * j + 1] = v[j8 + 2 * j + 1] * kp + __high2float(rh[j]);
                                    }
                                }
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                if (j < nvalid) {
                                    if (MK && !(__bfloat162float(MK[j]) > 0.f)) v[j] = 0.f;
                                    if (R) v[j] = v[j] * kp + __bfloat162float(R[j]);
                                }
                            }
                        }
                    }
                    if (tma_out) {
                        // [32 rows][64 B] SWIZZLE_64B: 16-byte chunk index ^= (row >> 1) & 3 (bank-conflict free)
                        if (p.stage_bufs == 2) tma_store_wait_read<1>();
                        else tma_store_wait_read<0>();
                        __syncwarp();
                        // 32-bit shared-window addresses computed once per warp (stage_s, st_row): no generic ->
                        // shared conversion and no address IMAD chain per chunk
                        const uint32_t sbase = stage_s + (uint32_t)sbuf * (uint32_t)kStageBytes;
#pragma unroll
                        for (int j8 = 0; j8 < 4; ++j8) {
                            const uint32_t o0 = pack_bf16x2(v[8 * j8], v[8 * j8 + 1]);
                            const uint32_t o1 = pack_bf16x2(v[8 * j8 + 2], v[8 * j8 + 3]);
                            const uint32_t o2 = pack_bf16x2(v[8 * j8 + 4], v[8 * j8 + 5]);
                            const uint32_t o3 = pack_bf16x2(v[8 * j8 + 6], v[8 * j8 + 7]);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                                         ::"r"(sbase + (st_row ^ (uint32_t)(j8 << 4))), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (elect_one()) tma_store_4d_s(&p.out_map[tc.gi], sbase, nbase, sx, sy, sb);
                        sbuf = (sbuf + 1) & (p.stage_bufs - 1);
                    } else if (nvalid == 32 && (reinterpret_cast<uintptr_t>(static_cast<__nv_bfloat16 *>(G.y) + base + nbase) & 15) == 0) {
                        __nv_bfloat16 *Y = static_cast<__nv_bfloat16 *>(G.y) + base + nbase;
#pragma unroll
                        for (int j8 = 0; j8 < 4; ++j8) {
                            uint4 ov;
                            __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&ov);
#pragma unroll
                            for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(v[8 * j8 + 2 * j], v[8 * j8 + 2 * j + 1]);
                            *reinterpret_cast<uint4 *>(Y + 8 * j8) = ov;
                        }
                    } else {
                        __nv_bfloat16 *Y = static_cast<__nv_bfloat16 *>(G.y) + base + nbase;
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (j < nvalid) Y[j] = __float2bfloat16_rn(v[j]);
                    }
                }
            }
            // a warp without a chunk in this tile (narrow N) still owes its arrival
            if (!released) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // outstanding TMA stores land before exit
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ------------------------------------------------------------------ weight panels
// mode 0 (forward):       panel[t][n = co][k = ci] = w[t][ci][co]
// mode 1 (data gradient): panel[t][n = ci][k = co] = w[taps-1-t][ci][co]
// gate != NULL (forward 1x1 only): per-sample panels panel[b][co][ci] = w[ci][co] * gate[b][ci]
// (the squeeze-excite multiply of efficientnet.py:286 folded into the weights of project_conv)
__global__ void __launch_bounds__(256)
weight_panel_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ panel, int taps, int Cin,
                    int Cout, int Kpad, int Npad, int mode, const float *__restrict__ gate, int nb) {
    EFFDET_PDL_SYNC();
    const unsigned per = (unsigned)Npad * Kpad;
    const unsigned total = (unsigned)(gate ? nb : taps) * per;
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u) {
        const unsigned z = i / per, rem = i - z * per;
        const int n = (int)(rem / Kpad), k = (int)(rem - (unsigned)n * Kpad);
        float v = 0.f;
        if (mode == 0) {
            if (k < Cin && n < Cout) {
                if (gate) v = w[(size_t)k * Cout + n] * gate[(size_t)z * Cin + k];
                else v = w[((size_t)z * Cin + k) * Cout + n];
            }
        } else {
            if (k < Cout && n < Cin) v = w[((size_t)(taps - 1 - z) * Cin + n) * Cout + k];
        }
        panel[i] = __float2bfloat16_rn(v);
    }
}

// Per-sample gated panels of a 1x1 convolution: panel[b][n][k] = w[k][n] * gate[b][k].
// One block transposes a 32(k) x 32(n) tile of w through shared memory once (coalesced reads
// along n) and streams it out for every sample with k fastest (coalesced 64-byte rows).
__global__ void __launch_bounds__(256)
gated_panel_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ panel, int Cin, int Cout,
                   int Kpad, int Npad, const float *__restrict__ gate, int nb) {
    EFFDET_PDL_SYNC();
    __shared__ float tile[32][33];
    const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int k = k0 + r, n = n0 + tx;
        tile[r][tx] = (k < Cin && n < Cout) ? w[(size_t)k * Cout + n] : 0.f;
    }
    __syncthreads();
    const int k = k0 + tx;
    for (int b = blockIdx.z; b < nb; b += gridDim.z) {
        const float g = k < Cin ? gate[(size_t)b * Cin + k] : 0.f;
        __nv_bfloat16 *dst = panel + (size_t)b * Npad * Kpad;
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int n = n0 + r;
            if (n < Npad && k < Kpad) dst[(size_t)n * Kpad + k] = __float2bfloat16_rn(tile[tx][r] * g);
        }
    }
}

// ------------------------------------------------------------------ weight gradient on tcgen05
// dW[tap][ci][co] = sum_pixels X[pix + tap][ci] * dZ[pix][co]
// GEMM per tap:  D[M = 128 ci, N = co tile] += A[ci, K = 128 pixels] * B[K = 128 pixels, co]
// Both operands have the reduction dimension (pixels) as the slow one in memory, i.e. they are
// "MN-major" for the tensor core: the TMA boxes (64 channels x 128 pixels, SWIZZLE_128B) are
// exactly the canonical MN-major atoms stacked along K (SBO = 1024 B between 8-pixel groups,
// LBO = 16 KiB between 64-channel boxes).  Split-K over pixel-tile ranges; fp32 partials are
// reduced in a fixed order by wgrad_reduce_kernel (deterministic).
struct WgTcGroup {
    int H, W, Wt, Ht, Bt, tiles_x, tiles_y, tiles_b;
    int n_tiles, tile_begin;    // this group's pixel tiles are [tile_begin, tile_begin + n_tiles) of the flat list
    int lw, lh, box_bytes;      // col_taps: log2(Wt), log2(Ht), bytes of one 64-channel activation box (with halo rows)
};
struct alignas(64) WgTcParams {
    CUtensorMap x_map[kTcMaxGroups];
    CUtensorMap z_map[kTcMaxGroups];
    WgTcGroup g[kTcMaxGroups];
    int n_groups, B, Cin, Cout, ksize, m_tiles, block_n, stages, tmem_cols;
    int total_tiles, tiles_per_split;   // split-K over the FLAT list of pixel tiles of all groups: CTA x owns tiles
                                        // [x * tiles_per_split, ...) -- equal work per CTA, one balanced wave
    int pair_taps;       // Cin <= 64: the two 64-channel halves of the M = 128 tile hold two different taps
    int col_taps;        // 3x3, Cin > 64: one CTA = one kx and all three ky of a 128-channel block.  The activation
                         // patch is loaded ONCE with a halo row above / below (box (64 ch, Wt, Ht + 2, Bt), zero fill
                         // = SAME padding); tap ky is the dense pixel range starting ky rows in (Wt in {8, 16, 32}:
                         // every 16-pixel K step starts on a swizzle-atom boundary), three accumulators of block_n
                         // columns in TMEM.  Per 128 pixels 40 + 16 * ceil(bn / 64) KiB of shared-memory fill feed
                         // 3 x 128 x bn x 128 MACs: 162 FLOP/B at bn = 112 instead of 83 for the per-tap form.
    int a_box_stride;    // col_taps: bytes reserved per 64-channel activation box
    float *partial;
    float *bias_partial; // [split][Cout] or NULL: the spare half of the last pair multiplies ones -> sum of dz
};

__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(192, 1)
conv_wgrad_tc_kernel(const __grid_constant__ WgTcParams p) {
    EFFDET_PDL_SYNC();
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int n_boxes = (p.block_n + 63) / 64;
    const int a_bytes = p.col_taps ? 2 * p.a_box_stride : 2 * kATileBytes, b_bytes = n_boxes * kATileBytes;
    uint8_t *sA = smem;
    uint8_t *sB = smem + (size_t)p.stages * a_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(sB + (size_t)p.stages * b_bytes);
    uint64_t *empty = full + p.stages;
    uint64_t *tmem_full = empty + p.stages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int t_begin = (int)blockIdx.x * p.tiles_per_split;
    const int t_end = min(t_begin + p.tiles_per_split, p.total_tiles);
    auto group_of = [&](int T) {
        int gi = 0;
#pragma unroll
        for (int i = 1; i < kTcMaxGroups; ++i)
            if (i < p.n_groups && T >= p.g[i].tile_begin) gi = i;
        return gi;
    };
    const int taps_all = p.ksize * p.ksize;
    // pair_taps: rows 0..63 of the accumulator = tap 2*blockIdx.y, rows 64..127 = the next tap (both
    // over input channels 0..63), sharing one dz tile -- halves the operand traffic of the 64-channel
    // head convolutions
    // col_taps: blockIdx.y = kx * m_tiles + m_tile; `tap` is then the ky = 0 tap of that column (kx)
    const int tap = p.pair_taps ? 2 * (int)blockIdx.y : (int)blockIdx.y / p.m_tiles;
    const int tap2 = p.pair_taps ? min(tap + 1, taps_all - 1) : tap;
    const int ci0 = p.pair_taps ? 0 : ((int)blockIdx.y % p.m_tiles) * 128;
    const int n0 = blockIdx.z * p.block_n;
    const int ky = tap / p.ksize, kx = tap - ky * p.ksize, pad = p.ksize / 2;
    const int ky2 = tap2 / p.ksize, kx2 = tap2 - ky2 * p.ksize;
    // last pair of an odd tap count: its second half is free.  With a bias gradient requested it holds
    // bf16 ones (written once, never touched by TMA), so accumulator rows 64..127 = sum over pixels of dz.
    const bool ones_half = p.pair_taps && p.bias_partial && tap + 1 >= taps_all;
    const int num_k = t_end - t_begin;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    if (ones_half) {
        for (int s = 0; s < p.stages; ++s) {
            uint4 *dst = reinterpret_cast<uint4 *>(sA + (size_t)s * a_bytes + kATileBytes);
            for (int i = threadIdx.x; i < kATileBytes / 16; i += blockDim.x)
                dst[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
        }
        fence_proxy_async();              // generic writes -> visible to the tensor core's async proxy
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % p.stages, ph = (kb / p.stages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                const int gi = group_of(t_begin + kb);
                const WgTcGroup &G = p.g[gi];
                int t = t_begin + kb - G.tile_begin;
                const int tx = t % G.tiles_x; t /= G.tiles_x;
                const int ty = t % G.tiles_y; t /= G.tiles_y;
                const int x0 = tx * G.Wt, y0 = ty * G.Ht, b0 = t * G.Bt;
                uint8_t *a = sA + (size_t)s * a_bytes, *b = sB + (size_t)s * b_bytes;
                if (p.col_taps) {
                    // here `tap` = kx (blockIdx.y / m_tiles): column kx - 1, rows y0 - 1 .. y0 + Ht
                    mbar_expect_tx(&full[s], (uint32_t)(2 * G.box_bytes + b_bytes));
                    tma_load_4d(a, &p.x_map[gi], &full[s], ci0, x0 + tap - 1, y0 - 1, b0);
                    tma_load_4d(a + p.a_box_stride, &p.x_map[gi], &full[s], ci0 + 64, x0 + tap - 1, y0 - 1, b0);
                    for (int j = 0; j < n_boxes; ++j)
                        tma_load_4d(b + (size_t)j * kATileBytes, &p.z_map[gi], &full[s], n0 + 64 * j, x0, y0, b0);
                    continue;
                }
                mbar_expect_tx(&full[s], (uint32_t)(a_bytes + b_bytes - (ones_half ? kATileBytes : 0)));
                tma_load_4d(a, &p.x_map[gi], &full[s], ci0, x0 + kx - pad, y0 + ky - pad, b0);
                if (ones_half) { }
                else if (p.pair_taps) tma_load_4d(a + kATileBytes, &p.x_map[gi], &full[s], 0, x0 + kx2 - pad, y0 + ky2 - pad, b0);
                else tma_load_4d(a + kATileBytes, &p.x_map[gi], &full[s], ci0 + 64, x0 + kx - pad, y0 + ky - pad, b0);
                for (int j = 0; j < n_boxes; ++j)
                    tma_load_4d(b + (size_t)j * kATileBytes, &p.z_map[gi], &full[s], n0 + 64 * j, x0, y0, b0);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // a_major = b_major = MN (bits 15, 16)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % p.stages, ph = (kb / p.stages) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t a = smem_u32(sA + (size_t)s * a_bytes), b = smem_u32(sB + (size_t)s * b_bytes);
                if (p.col_taps) {
                    const WgTcGroup &G = p.g[group_of(t_begin + kb)];
#pragma unroll 1
                    for (int k = 0; k < 8; ++k) {
                        const int pix = 16 * k;
                        const int xx = pix & (G.Wt - 1), yy = (pix >> G.lw) & (G.Ht - 1), bb = pix >> (G.lw + G.lh);
                        const uint32_t row0 = (uint32_t)((bb * (G.Ht + 2) + yy) * G.Wt + xx) * 128u;      // ky = 0
                        const uint64_t db = make_mnmajor_sw128_desc(b + (uint32_t)k * 2048u, kATileBytes);
#pragma unroll
                        for (int kyy = 0; kyy < 3; ++kyy) {
                            const uint64_t da = make_mnmajor_sw128_desc(a + row0 + (uint32_t)(kyy * G.Wt) * 128u,
                                                                        (uint32_t)p.a_box_stride);
                            umma_bf16(tmem_base + (uint32_t)(kyy * p.block_n), da, db, idesc, (kb | k) ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty[s]);
                    continue;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {   // 128 pixels = 8 x UMMA_K(16); 16 pixel rows = 2048 bytes
                    const uint64_t da = make_mnmajor_sw128_desc(a + k * 2048, kATileBytes);
                    const uint64_t db = make_mnmajor_sw128_desc(b + k * 2048, kATileBytes);
                    umma_bf16(tmem_base, da, db, idesc, (kb | k) ? 1u : 0u);
                }
                umma_commit(&empty[s]);
            }
            umma_commit(tmem_full);
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int ci = p.pair_taps ? (row & 63) : ci0 + row;
        const int otap = p.pair_taps ? tap + (row >> 6) : tap;
        if (num_k > 0) {
            mbar_wait(tmem_full, 0);
            tc_fence_after();
        }
        if (p.col_taps) {
            // accumulator kyy (columns kyy * block_n ..) = tap (ky = kyy, kx = `tap`)
            for (int kyy = 0; kyy < 3; ++kyy) {
                float *o_tap = p.partial + ((size_t)blockIdx.x * taps_all + (kyy * 3 + tap)) * p.Cin * p.Cout;
                for (int c0 = 0; c0 < p.block_n; c0 += 16) {
                    uint32_t r[16];
                    __syncwarp();
                    if (num_k > 0) tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kyy * p.block_n + c0), r);
                    if (ci < p.Cin) {
                        float *o = o_tap + (size_t)ci * p.Cout + n0 + c0;
                        const int nv = min(16, p.Cout - (n0 + c0));
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (j < nv) o[j] = num_k > 0 ? __uint_as_float(r[j]) : 0.f;
                    }
                }
            }
        } else {
        float *out = p.partial + ((size_t)blockIdx.x * taps_all + min(otap, taps_all - 1)) * p.Cin * p.Cout;
        const bool row_ok = ci < p.Cin && otap < taps_all;
        for (int c0 = 0; c0 < p.block_n; c0 += 32) {
            uint32_t r[32];
            __syncwarp();
            if (num_k > 0) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            if (row_ok) {
                float *o = out + (size_t)ci * p.Cout + n0 + c0;
                const int nv = min(32, p.Cout - (n0 + c0));
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < nv) o[j] = num_k > 0 ? __uint_as_float(r[j]) : 0.f;
            }
            if (ones_half && row == 64) {
                float *o = p.bias_partial + (size_t)blockIdx.x * p.Cout + n0 + c0;
                const int nv = min(32, p.Cout - (n0 + c0));
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < nv) o[j] = num_k > 0 ? __uint_as_float(r[j]) : 0.f;
            }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}


// ------------------------------------------------------------------ 3x3 weight gradient, Cin <= 64: halo version
// The per-tap kernel above moves every activation tile through L2 -> shared memory once per tap
// (9 x 16 KiB per 128 pixels); measured, the 64-channel head convolutions sit at the L2 bandwidth
// (profiles/).  Here ONE CTA handles all nine taps of its pixel tiles.  Per tile it loads three
// copies of the activation patch with one row of halo above / below, shifted by kx = 0, 1, 2 columns
// (box (64 ch, Wt, Ht + 2, Bt), out-of-image = zero = SAME padding) and the dz tile once.  Inside copy
// kx, tap (ky, kx) is the dense pixel range starting ky rows in: with Wt in {8, 16, 32} every
// 16-pixel K step starts on an 8-pixel (1024-byte swizzle atom) boundary, so plain MN-major
// descriptors address it.  Two taps share one M = 128 instruction (LBO = distance between their
// two 64-channel operands); the ninth tap is paired with a block of ones, which yields the bias
// gradient.  Five accumulators of 64 columns live in TMEM.  ~76 KiB instead of 240 KiB per tile.
struct WgHaloGroup { int Wt, Ht, Bt, lw, lh, tiles_x, tiles_y, n_tiles, tile_begin, box_bytes; };
struct alignas(64) WgHaloParams {
    CUtensorMap x_map[kTcMaxGroups];
    CUtensorMap z_map[kTcMaxGroups];
    WgHaloGroup g[kTcMaxGroups];
    int n_groups, Cin, Cout, stages, box_stride;      // box_stride: bytes reserved per activation copy
    int ksteps, z_bytes;                              // 16-pixel K steps per tile (4 or 8), bytes of the dz tile
    int total_tiles, tiles_per_split;                 // flat split-K over all groups' tiles (see WgTcParams)
    float *partial, *bias_partial;
    int dbg;                                          // EFFDET_WG_DBG: 1 no MMAs, 2 one activation copy, 4 no partial stores (timing experiments)
};

__global__ void __launch_bounds__(192, 1)
conv_wgrad_halo_kernel(const __grid_constant__ WgHaloParams p) {
    EFFDET_PDL_SYNC();
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int stage_bytes = 3 * p.box_stride + p.z_bytes;            // 3 activation copies + dz tile
    uint8_t *sOnes = smem + (size_t)p.stages * stage_bytes;          // box_stride bytes of bf16 1.0
    uint64_t *full = reinterpret_cast<uint64_t *>(sOnes + p.box_stride);
    uint64_t *empty = full + p.stages;
    uint64_t *tmem_full = empty + p.stages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int t_begin = (int)blockIdx.x * p.tiles_per_split;
    const int t_end = min(t_begin + p.tiles_per_split, p.total_tiles);
    auto group_of = [&](int T) {
        int gi = 0;
#pragma unroll
        for (int i = 1; i < kTcMaxGroups; ++i)
            if (i < p.n_groups && T >= p.g[i].tile_begin) gi = i;
        return gi;
    };
    const int num_k = t_end - t_begin;
    const int n0 = blockIdx.y * 64;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512u);
    for (int i = threadIdx.x; i < p.box_stride / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(sOnes)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % p.stages, ph = (kb / p.stages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                const int gi = group_of(t_begin + kb);
                const WgHaloGroup &G = p.g[gi];
                int t = t_begin + kb - G.tile_begin;
                const int tx = t % G.tiles_x; t /= G.tiles_x;
                const int ty = t % G.tiles_y; t /= G.tiles_y;
                const int x0 = tx * G.Wt, y0 = ty * G.Ht, b0 = t * G.Bt;
                uint8_t *st = smem + (size_t)s * stage_bytes;
                const int ncopy = (p.dbg & 2) ? 1 : 3;
                mbar_expect_tx(&full[s], (uint32_t)(ncopy * G.box_bytes + p.z_bytes));
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
                    if (kx < ncopy)
                        tma_load_4d(st + (size_t)kx * p.box_stride, &p.x_map[gi], &full[s], 0, x0 - 1 + kx, y0 - 1, b0);
                tma_load_4d(st + 3 * (size_t)p.box_stride, &p.z_map[gi], &full[s], n0, x0, y0, b0);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // M = 128 (two taps x 64 channels), N = 64, both operands MN-major
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t ones_addr = smem_u32(sOnes);
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % p.stages, ph = (kb / p.stages) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const WgHaloGroup &G = p.g[group_of(t_begin + kb)];
                const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t zb = st + 3u * (uint32_t)p.box_stride;
                // Descriptors are built once per tile: the five tap-pair operands differ from K step to K step only
                // in their start address, i.e. by an addition to the descriptor's 14-bit address field (the sums stay
                // below 256 KiB, so nothing carries into the LBO field).  Measured (EFFDET_WG_DBG experiments, D0
                // trunk gradient): the 400 MMAs of a CTA cost 11.5 us = ~56 cycles each whether N is 32, 64 or 128 and
                // whether the issue loop has 9 or 30 instructions per MMA -- a per-instruction floor of the tensor
                // pipe at M = 128, K = 16; the other 21 us are loads (7.5), partial stores (5.7), the split-K
                // reduction and the prologue.
                const uint32_t wt128 = (uint32_t)G.Wt * 128u;
                uint64_t da0[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    // operand slots u = kx * 3 + ky are ordered by shared-memory address (copy kx, then
                    // ky rows in), so the second half of a pair always lies above the first (LBO > 0)
                    const int ua = 2 * j, ub = 2 * j + 1;
                    const uint32_t aa = st + (uint32_t)(ua / 3) * (uint32_t)p.box_stride + (uint32_t)(ua % 3) * wt128;
                    // second half: slot ub, or (j == 4) the block of ones at the same relative offset
                    const uint32_t ab = ub < 9 ? st + (uint32_t)(ub / 3) * (uint32_t)p.box_stride + (uint32_t)(ub % 3) * wt128
                                               : ones_addr;
                    da0[j] = make_mnmajor_sw128_desc(aa, ab - aa);
                }
                const uint64_t db0 = make_mnmajor_sw128_desc(zb, (uint32_t)p.z_bytes);
                const int ht2 = G.Ht + 2, wm = G.Wt - 1, hm = G.Ht - 1, lw = G.lw, lwh = G.lw + G.lh, wt = G.Wt;
                const bool first_tile = kb == 0;
                const bool no_mma = (p.dbg & 1) != 0;
#pragma unroll 1
                for (int k = 0; k < p.ksteps; ++k) {          // 16-pixel K steps of the tile
                    const int pix = 16 * k;
                    const int xx = pix & wm, yy = (pix >> lw) & hm, bb = pix >> lwh;
                    const uint64_t r4 = (uint64_t)(((uint32_t)((bb * ht2 + yy) * wt + xx) * 128u) >> 4);   // tap ky = 0
                    const uint64_t db = db0 + (uint64_t)(k * (2048 >> 4));
                    const uint32_t acc = (first_tile && k == 0) ? 0u : 1u;
                    if (!no_mma) {
#pragma unroll
                        for (int j = 0; j < 5; ++j) umma_bf16(tmem_base + (uint32_t)(j * 64), da0[j] + r4, db, idesc, acc);
                    }
                }
                umma_commit(&empty[s]);
            }
            umma_commit(tmem_full);
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        if (num_k > 0) {
            mbar_wait(tmem_full, 0);
            tc_fence_after();
        }
        const int ci = row & 63;
        for (int j = 0; j < 5; ++j) {
            const int u = 2 * j + (row >> 6);                       // operand slot kx * 3 + ky; 9 = ones
            const int tap = u < 9 ? (u % 3) * 3 + u / 3 : 9;        // HWIO tap index ky * 3 + kx
            float *out = p.partial + ((size_t)blockIdx.x * 9 + min(tap, 8)) * p.Cin * p.Cout;
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t r[32];
                __syncwarp();
                if (num_k > 0) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 64 + c0), r);
                const int nv = min(32, p.Cout - (n0 + c0));
                float *o = nullptr;
                if (tap < 9 && ci < p.Cin && !(p.dbg & 4)) o = out + (size_t)ci * p.Cout + n0 + c0;
                else if (tap == 9 && row == 64 && p.bias_partial) o = p.bias_partial + (size_t)blockIdx.x * p.Cout + n0 + c0;
                if (o) {
                    if (num_k == 0) {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) r[jj] = 0u;
                    }
                    if (nv == 32 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4)
                            reinterpret_cast<uint4 *>(o)[j4] = make_uint4(r[4 * j4], r[4 * j4 + 1], r[4 * j4 + 2], r[4 * j4 + 3]);
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj)
                            if (jj < nv) o[jj] = __uint_as_float(r[jj]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512u);
    }
}

// Split-K reduction.  A block of 256 threads owns 256 / S consecutive outputs; slice q of its S slices sums the split
// rows q, q + S, ... (coalesced reads, eight independent rows in flight per thread), then the S slice sums are
// combined in a fixed order: deterministic.  S = 8 for the many-row reductions of the D0 heads (137 rows x 147 KB:
// the thread-per-output form was a chain of 19 dependent L2 round trips, 11 us), S = 1 for few rows / many outputs.
__global__ void __launch_bounds__(256)
wgrad_tc_reduce_kernel(const float *__restrict__ partial, int nsplit, size_t n,
                       float *__restrict__ out, int accumulate,
                       const float *__restrict__ bias_partial, int n_bias,
                       float *__restrict__ dbias, int S) {
    EFFDET_PDL_SYNC();
    __shared__ float sred[256];
    const int per = 256 / S;
    const int e = threadIdx.x % per, slice = threadIdx.x / per;
    const size_t i = (size_t)blockIdx.x * per + e;
    const bool is_w = i < n, is_b = !is_w && i < n + (size_t)n_bias;
    const float *src = is_w ? partial + i : bias_partial + (i - n);
    const size_t stride = is_w ? n : (size_t)n_bias;
    float t = 0.f;
    if (is_w || is_b) {
        int s = slice;
        for (; s + 7 * S < nsplit; s += 8 * S) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + (size_t)(s + u * S) * stride);
#pragma unroll
            for (int u = 0; u < 8; ++u) t += v[u];
        }
        for (; s < nsplit; s += S) t += __ldcg(src + (size_t)s * stride);
    }
    if (S > 1) {
        sred[threadIdx.x] = t;
        __syncthreads();
        if (slice != 0) return;
        t = 0.f;
        for (int q = 0; q < S; ++q) t += sred[q * per + e];
    }
    if (is_w) out[i] = accumulate ? out[i] + t : t;
    else if (is_b) dbias[i - n] = t;
}
static int wg_reduce_slices(int nsplit) { return nsplit >= 64 ? 8 : nsplit >= 32 ? 4 : nsplit >= 16 ? 2 : 1; }

// picks (Wt, Ht, Bt) with Wt*Ht*Bt == 128 minimising the number of tiles
static void pick_tile(int W, int H, int B, int *Wt, int *Ht, int *Bt) {
    static const int cand[][3] = {{16, 8, 1}, {8, 16, 1}, {32, 4, 1}, {4, 32, 1}, {8, 8, 2}, {4, 8, 4}, {8, 4, 4},
                                  {4, 4, 8}, {2, 4, 16}, {4, 2, 16}, {2, 2, 32}, {1, 2, 64}, {2, 1, 64},
                                  {1, 1, 128}, {16, 4, 2}, {4, 16, 2}, {16, 2, 4}, {2, 16, 4}, {64, 2, 1}, {128, 1, 1}};
    long best = -1;
    for (auto &c : cand) {
        long n = (long)cdiv(W, c[0]) * cdiv(H, c[1]) * cdiv(B, c[2]);
        if (best < 0 || n < best) { best = n; *Wt = c[0]; *Ht = c[1]; *Bt = c[2]; }
    }
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

}  // namespace effdet

using namespace effdet;

// N tile of the tensor-core convolution for an (N = Cout) x (num_k = taps * ceil(K / 64)) problem.
// Up to 256 columns in one tile (the activations are then read once).  Wider N is split into equal
// tiles that are multiples of 32 (no 32-column epilogue chunk straddles two tiles): <= 256 columns
// for deep reductions (tensor-bound 3x3 head convolutions: the activation tile is re-read from L2 once
// per N tile, so wide tiles cut the shared-memory fill traffic), <= 128 columns for shallow ones
// (HBM-bound expand convolutions: two TMEM accumulators and two CTAs per SM hide the epilogue).
static int tc_block_n(int n, int num_k) {
    const int n16 = round_up(n, 16);
    if (n16 <= 256) return n16;
    const int cap = num_k >= 8 ? 256 : 128;
    const int tiles = (n16 + cap - 1) / cap;
    return round_up((n16 + tiles - 1) / tiles, 32);
}

extern "C" int effdet_conv_tc_block_n(int n) { return tc_block_n(n, 1); }

extern "C" size_t effdet_conv_weight_panel_elems(int taps_or_samples, int K, int N) {
    // large enough for either tiling rule (the caller does not say whether the first argument counts
    // taps or samples)
    const int a = round_up(N, tc_block_n(N, 1)), b = round_up(N, tc_block_n(N, 8));
    return (size_t)taps_or_samples * (a > b ? a : b) * round_up(K, kTileK);
}

extern "C" int effdet_conv_weight_panel(const float *w, void *panel, int taps, int Cin, int Cout, int mode,
                                        const float *gate, int B, void *stream) {
    EFFDET_REQUIRE(w && panel && taps > 0 && Cin > 0 && Cout > 0, "bad arguments");
    EFFDET_REQUIRE(mode == 0 || mode == 1, "mode 0 (forward) or 1 (data gradient)");
    EFFDET_REQUIRE(!gate || (mode == 0 && taps == 1 && B > 0), "gate only for forward 1x1");
    const int K = mode == 0 ? Cin : Cout, N = mode == 0 ? Cout : Cin;
    const int Kpad = round_up(K, kTileK);
    const int bn = tc_block_n(N, (gate ? 1 : taps) * (Kpad / kTileK));
    const int Npad = round_up(N, bn);
    const size_t total = (size_t)(gate ? B : taps) * Npad * Kpad;
    EFFDET_REQUIRE(total < 0xffffffffull, "panel too large");
    if (gate) {
        int zs = B < 8 ? B : 8;             // samples per block column: enough blocks to fill the machine
        dim3 grid(Kpad / 32, (Npad + 31) / 32, zs);
        EFFDET_CUDA(launch_pdl(gated_panel_kernel, dim3(grid), dim3(256), 0, as_stream(stream), w, static_cast<__nv_bfloat16 *>(panel), Cin, Cout,
                                                                Kpad, Npad, gate, B));
        EFFDET_LAUNCHED();
        return EFFDET_OK;
    }
    unsigned blocks = cdiv(total, 256);
    if (blocks > (unsigned)kNumSMs * 16) blocks = kNumSMs * 16;
    EFFDET_CUDA(launch_pdl(weight_panel_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), w, static_cast<__nv_bfloat16 *>(panel), taps, Cin,
                                                               Cout, Kpad, Npad, mode, gate, B));
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

// ------------------------------------------------------------------ fp32 accuracy mode on the tensor cores
// x = hi + lo + O(2^-18 |x|) with hi = bf16(x), lo = bf16(x - hi).  The convolution of an fp32 tensor by fp32
// weights becomes ONE bf16 GEMM over the K extent [hi | lo | hi] x [Whi | Whi | Wlo] with fp32 accumulation in
// TMEM (the lo * Wlo term, ~2^-18 of the product, is dropped): three times the MMAs of the bf16 speed mode instead
// of the SIMT fp32 kernel (78 of the 145 ms of a D2 / batch-64 forward were its head convolutions).
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ out, size_t n_vec, int C4) {
    // thread = 4 consecutive channels of one row: one 16-byte load, two 8-byte stores (hi plane, lo plane)
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * 256) {
        const size_t row = i / (size_t)C4;
        const int c4 = (int)(i - row * (size_t)C4);
        const float4 v = *reinterpret_cast<const float4 *>(x + i * 4);
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y),
                            h2 = __float2bfloat16_rn(v.z), h3 = __float2bfloat16_rn(v.w);
        __nv_bfloat162 hi[2] = {__halves2bfloat162(h0, h1), __halves2bfloat162(h2, h3)};
        __nv_bfloat162 lo[2] = {__floats2bfloat162_rn(v.x - __bfloat162float(h0), v.y - __bfloat162float(h1)),
                                __floats2bfloat162_rn(v.z - __bfloat162float(h2), v.w - __bfloat162float(h3))};
        __nv_bfloat16 *o = out + row * (size_t)(8 * C4) + (size_t)c4 * 4;
        *reinterpret_cast<uint2 *>(o) = *reinterpret_cast<uint2 *>(hi);
        *reinterpret_cast<uint2 *>(o + 4 * C4) = *reinterpret_cast<uint2 *>(lo);
    }
}

// panel[z][n][3 * Kp]: k in [0, Kp) = Whi, [Kp, 2 Kp) = Whi again, [2 Kp, 3 Kp) = Wlo of w[z][k][n] (* gate[z][k])
__global__ void __launch_bounds__(256)
weight_panel_split_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ panel, int taps, int Cin,
                          int Cout, int Kp, int Npad, const float *__restrict__ gate, int nb) {
    EFFDET_PDL_SYNC();
    const unsigned per = (unsigned)Npad * Kp;
    const unsigned total = (unsigned)(gate ? nb : taps) * per;
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u) {
        const unsigned z = i / per, rem = i - z * per;
        const int n = (int)(rem / Kp), k = (int)(rem - (unsigned)n * Kp);
        float v = 0.f;
        if (k < Cin && n < Cout) {
            if (gate) v = w[(size_t)k * Cout + n] * gate[(size_t)z * Cin + k];
            else v = w[((size_t)z * Cin + k) * Cout + n];
        }
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        __nv_bfloat16 *row = panel + ((size_t)z * Npad + n) * (size_t)(3 * Kp);
        row[k] = hi; row[Kp + k] = hi; row[2 * Kp + k] = lo;
    }
}

extern "C" int effdet_split_bf16(const float *x, void *out, size_t rows, int C, void *stream) {
    EFFDET_REQUIRE(x && out && C > 0 && C % 4 == 0, "C must be a multiple of 4");
    EFFDET_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "16-byte alignment");
    const size_t n_vec = rows * (size_t)(C / 4);
    if (n_vec == 0) return EFFDET_OK;
    unsigned blocks = cdiv(n_vec, 256);
    if (blocks > (unsigned)kNumSMs * 16) blocks = kNumSMs * 16;
    split_bf16_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, static_cast<__nv_bfloat16 *>(out), n_vec, C / 4);
    EFFDET_CUDA(cudaGetLastError());
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

extern "C" size_t effdet_conv_weight_panel_split_elems(int taps_or_samples, int Cin, int Cout) {
    const int K3 = 3 * round_up(Cin, kTileK);
    const int a = round_up(Cout, tc_block_n(Cout, 3)), b = round_up(Cout, tc_block_n(Cout, 8));
    return (size_t)taps_or_samples * (a > b ? a : b) * K3;
}

extern "C" int effdet_conv_weight_panel_split(const float *w, void *panel, int taps, int Cin, int Cout,
                                              const float *gate, int B, void *stream) {
    EFFDET_REQUIRE(w && panel && taps > 0 && Cin > 0 && Cout > 0, "bad arguments");
    EFFDET_REQUIRE(!gate || (taps == 1 && B > 0), "gate only for 1x1");
    const int Kp = round_up(Cin, kTileK);
    const int bn = tc_block_n(Cout, (gate ? 1 : taps) * 3 * (Kp / kTileK));
    const int Npad = round_up(Cout, bn);
    const size_t total = (size_t)(gate ? B : taps) * Npad * Kp;
    EFFDET_REQUIRE(total < 0xffffffffull, "panel too large");
    unsigned blocks = cdiv(total, 256);
    if (blocks > (unsigned)kNumSMs * 16) blocks = kNumSMs * 16;
    EFFDET_CUDA(launch_pdl(weight_panel_split_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), w,
                           static_cast<__nv_bfloat16 *>(panel), taps, Cin, Cout, Kp, Npad, gate, B));
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

// Shared memory of the halo form (two stages of three kx-shifted patches, nine resident weight tiles, barriers and
// scale / shift staging, ONE output staging buffer per epilogue warp) against the 227 KiB a CTA can have.  fp32
// outputs with 49..64 output channels (a class head of 6 or 7 classes on a 64-wide BiFPN) do not fit: 230 KiB --
// they take the ring form (found by examples/detect_host.c; the launch failed with "invalid argument").
static bool halo_fits(int bn, bool out_f32) {
    const size_t stage_out_min = (size_t)kEpiWarps * 32 * 32 * (out_f32 ? 4 : 2);
    const size_t smem = (size_t)2 * 3 * 20 * 1024 + (size_t)9 * bn * kTileK * 2 + 4096 + (2 * 2 + 5) * 8 + 16 + 1024 +
                        1024 + stage_out_min;
    return smem <= 227 * 1024;
}

// Called by effdet_conv2d when the descriptor qualifies.  Returns EFFDET_E_UNSUPPORTED (without
// touching the error string of a real failure) when the shape cannot take this path.
int effdet_conv2d_tc(const effdet_conv_desc *d, void *stream) {
    if (!d->weight_bf16 || d->in_dtype != EFFDET_BF16 || (d->stride != 1 && d->stride != 2)) return EFFDET_E_UNSUPPORTED;
    if (d->stride == 2 && (d->relu_mask[0] || d->weight_per_sample)) return EFFDET_E_UNSUPPORTED;
    if (d->Cin % 8 != 0 || d->kh != d->kw) return EFFDET_E_UNSUPPORTED;
    if (d->gate && !d->weight_per_sample) return EFFDET_E_UNSUPPORTED;
    EncodeTiledFn encode = get_encode();
    if (!encode) return fail(EFFDET_E_CUDA, "effdet_conv2d: cuTensorMapEncodeTiled unavailable%s", "");
    static TcParams p;      // large; filled per call (single-threaded use per the ABI contract)
    memset(&p, 0, sizeof(p));
    const int split = d->split_planes ? 1 : 0;       // fp32 accuracy mode: [hi | lo | hi] x [Whi | Whi | Wlo]
    if (split && (d->out_dtype != EFFDET_F32 || d->relu_mask[0])) return EFFDET_E_UNSUPPORTED;
    const int Kplane = round_up(d->Cin, kTileK);
    const int Kpad = (split ? 3 : 1) * Kplane;       // K extent of the GEMM = of the weight panel
    const int bn_panel = tc_block_n(d->Cout, d->kh * d->kw * (Kpad / kTileK));      // N tile the weight panel was padded for
    const int Npad = round_up(d->Cout, bn_panel);
    // Few M tiles (batch 1, coarse pyramid levels): a 16 x 16 x 1152 -> 192 project convolution is TWO tiles, i.e. two
    // SMs pulling 720 KB each through their TMA units while 146 idle (13 us; the batch-1 step is a chain of such
    // launches).  Narrower N tiles (still multiples of 32, dividing the panel's padding) spread the weight stream
    // over more SMs; the activation tile is re-read from L2 by each of them.  Not for the halo form (weights resident).
    int bn = bn_panel;
    {
        const bool halo_ok = d->kh == 3 && d->stride == 1 && Kpad == kTileK && !split && Npad == bn_panel &&
                             !d->weight_per_sample && 9 * bn_panel * kTileK * 2 <= 80 * 1024 &&
                             halo_fits(bn_panel, d->out_dtype == EFFDET_F32) && getenv("EFFDET_NO_CONV_HALO") == nullptr;
        static const int min_ctas = getenv("EFFDET_TC_MIN_CTAS") ? atoi(getenv("EFFDET_TC_MIN_CTAS")) : 64;
        long m_est = 0;
        for (int i = 0; i < d->n_groups; ++i) {
            const long Ho = (d->H[i] + d->stride - 1) / d->stride, Wo = (d->W[i] + d->stride - 1) / d->stride;
            m_est += ((long)d->B * Ho * Wo + kTileM - 1) / kTileM;
        }
        for (int c = (bn - 1) / 32 * 32; !halo_ok && c >= 32 && m_est * (Npad / bn) < min_ctas; c -= 32)
            if (Npad % c == 0) bn = c;      // multiples of 32: no 32-column epilogue chunk straddles two tiles
    }
    p.n_groups = d->n_groups; p.B = d->B; p.Cout = d->Cout; p.ksize = d->kh; p.stride = d->stride;
    p.kblocks_per_tap = Kpad / kTileK; p.block_n = bn;
    p.split = split; p.kb_plane = Kplane / kTileK; p.plane_stride = d->Cin;
    const int num_k = d->kh * d->kw * p.kblocks_per_tap;
    const int acc_pow2 = bn <= 32 ? 32 : bn <= 64 ? 64 : bn <= 128 ? 128 : 256;
    // two accumulators (epilogue of tile i overlaps the MMAs of tile i+1) unless the tile is wide
    // AND shallow (HBM-bound expand convolutions): then one accumulator and two CTAs per SM
    p.acc_bufs = (acc_pow2 == 256 && num_k <= 4) ? 1 : 2;
    p.tmem_cols = p.acc_bufs * acc_pow2;
    // weights resident in shared memory: single N tile, shared by all samples, <= 48 KiB
    p.b_resident = (Npad == bn && !d->weight_per_sample && (size_t)num_k * bn * kTileK * 2 <= 48 * 1024) ? 1 : 0;
    p.act = d->act; p.out_f32 = d->out_dtype == EFFDET_F32; p.b_per_sample = d->weight_per_sample ? 1 : 0;
    p.scale = d->scale; p.shift = d->shift; p.keep = d->keep;
    const int b_tile_bytes = bn * kTileK * 2;
    // halo mode: 3x3 stride 1, one K block per tap, one N tile, all nine weight tiles resident (<= 80 KiB)
    p.halo = (d->kh == 3 && d->stride == 1 && Kpad == kTileK && !split && Npad == bn && !d->weight_per_sample &&
              9 * b_tile_bytes <= 80 * 1024 && halo_fits(bn, p.out_f32) && getenv("EFFDET_NO_CONV_HALO") == nullptr) ? 1 : 0;
    p.box_stride = 20 * 1024;                         // (16, 8+2) or (8, 16+2) pixels x 128 bytes
    p.a_stage_bytes = p.halo ? 3 * p.box_stride : kATileBytes;
    p.stage_bufs = 2;
    if (p.halo) { p.b_resident = 1; p.acc_bufs = 2; p.tmem_cols = 2 * acc_pow2; }
    // <= 4 stages: with 64..128-wide N tiles two CTAs stay resident per SM, so one CTA's
    // epilogue overlaps the other's main loop
    // small tiles: 2 resident CTAs per SM (<= ~100 KiB each); wide tiles: 1
    int stage_out_bytes = kEpiWarps * 2 * 32 * 32 * (p.out_f32 ? 4 : 2);
    if (p.halo && 9 * b_tile_bytes + 2 * p.a_stage_bytes + stage_out_bytes + 7 * 1024 > 226 * 1024) {
        p.stage_bufs = 1;                             // one staging buffer per epilogue warp
        stage_out_bytes /= 2;
    }
    const int two_ctas = p.tmem_cols <= 256 && !p.halo;
    const int ring_bytes = p.halo ? p.a_stage_bytes : kATileBytes + (p.b_resident ? 0 : b_tile_bytes);
    const int fixed_bytes = stage_out_bytes + (p.b_resident ? num_k * b_tile_bytes : 0) + 7 * 1024;
    int stages = ((two_ctas ? 113 : 226) * 1024 - fixed_bytes) / ring_bytes;
    if (stages > (p.b_resident ? 6 : 4)) stages = p.b_resident ? 6 : 4;
    if (p.halo && stages > 2) stages = 2;
    // persistent kernel: the ring runs ahead ACROSS tiles, so its depth does not depend on the
    // number of K blocks of one tile (a 1x1 convolution with Cin <= 64 has a single K block)
    if (stages < 2) stages = 2;
    p.stages = stages;
    const size_t smem = (size_t)stages * ring_bytes + (p.b_resident ? (size_t)num_k * b_tile_bytes : 0) + 4096 +
                        (2 * stages + 5) * 8 + 16 + 1024 + 1024 + stage_out_bytes;
    if (smem > 227 * 1024)       // never for the shapes tc_block_n / halo_fits admit; a launch would fail with "invalid argument"
        return fail(EFFDET_E_INVALID, "effdet_conv2d: %s%lld bytes of shared memory exceed a CTA's 227 KiB", "", (long long)smem);
    p.any_tma_store = 0;
    int tiles = 0;
    for (int i = 0; i < d->n_groups; ++i) {
        TcGroup &g = p.g[i];
        g.y = d->y[i]; g.res = d->residual[i]; g.mask = d->relu_mask[i];
        // g.H / g.W are OUTPUT extents; stride 2 reads every other input pixel (TMA element strides)
        const int Hin = d->H[i], Win = d->W[i], sd = d->stride;
        g.H = (Hin + sd - 1) / sd; g.W = (Win + sd - 1) / sd;
        p.pad_t[i] = max((g.H - 1) * sd + d->kh - Hin, 0) / 2;      // TF SAME: before = total / 2
        p.pad_l[i] = max((g.W - 1) * sd + d->kw - Win, 0) / 2;
        g.ldc = d->ldc[i] ? d->ldc[i] : d->Cout;
        g.y_batch_stride = d->y_batch_stride[i] ? d->y_batch_stride[i] : (long long)g.H * g.W * g.ldc;
        if (p.halo) {
            // one image per tile; 16 x 8 or 8 x 16 pixels (the tap offsets stay swizzle-atom aligned)
            const long n_a = (long)cdiv(g.W, 16) * cdiv(g.H, 8), n_b = (long)cdiv(g.W, 8) * cdiv(g.H, 16);
            g.Bt = 1;
            if (n_a <= n_b) { g.Wt = 16; g.Ht = 8; } else { g.Wt = 8; g.Ht = 16; }
            p.box_bytes[i] = g.Wt * (g.Ht + 2) * 128;
        }
        else if (d->weight_per_sample) { g.Bt = 1; pick_tile(g.W, g.H, 1, &g.Wt, &g.Ht, &g.Bt); if (g.Bt != 1) { g.Wt = 16; g.Ht = 8; g.Bt = 1; } }
        else pick_tile(g.W, g.H, d->B, &g.Wt, &g.Ht, &g.Bt);
        g.tiles_x = cdiv(g.W, g.Wt); g.tiles_y = cdiv(g.H, g.Ht); g.tiles_b = cdiv(d->B, g.Bt);
        g.lw = 31 - __builtin_clz(g.Wt); g.lh = 31 - __builtin_clz(g.Ht);
        g.tile_begin = tiles;
        tiles += g.tiles_x * g.tiles_y * g.tiles_b;
        const long long ldx = d->ldx[i] ? d->ldx[i] : (split ? 2 : 1) * (long long)d->Cin;
        const long long xbs = d->x_batch_stride[i] ? d->x_batch_stride[i] : (long long)Hin * Win * ldx;
        if ((ldx * 2) % 16 || (xbs * 2) % 16 || (reinterpret_cast<uintptr_t>(d->x[i]) & 15)) return EFFDET_E_UNSUPPORTED;
        if (g.Wt * sd > 256 || g.Ht * sd > 256) return EFFDET_E_UNSUPPORTED;
        cuuint64_t dims[4] = {(cuuint64_t)((split ? 2 : 1) * d->Cin), (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)d->B};
        cuuint64_t strides[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)ldx * 2 * Win, (cuuint64_t)xbs * 2};
        cuuint32_t box[4] = {(cuuint32_t)kTileK, (cuuint32_t)(g.Wt * sd), (cuuint32_t)((g.Ht + (p.halo ? 2 : 0)) * sd),
                             (cuuint32_t)g.Bt};
        cuuint32_t es_in[4] = {1, (cuuint32_t)sd, (cuuint32_t)sd, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = encode(&p.a_map[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(d->x[i]), dims,
                            strides, box, es_in, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(EFFDET_E_CUDA, "effdet_conv2d: cuTensorMapEncodeTiled(A) failed %s(%lld)", "", (long long)r);
        // coalesced output: the epilogue stages 32-row x 32-column pieces in shared memory and TMA
        // stores them (rows / channels outside the tensor are clipped by the tensor map)
        const int es_out = p.out_f32 ? 4 : 2;
        g.tma_store = ((long long)g.ldc * es_out) % 16 == 0 && (g.y_batch_stride * es_out) % 16 == 0 &&
                      (reinterpret_cast<uintptr_t>(g.y) & 15) == 0;
        if (g.tma_store) {
            cuuint32_t obox[4];
            obox[0] = 32;
            if (g.Wt >= 32) { obox[1] = 32; obox[2] = 1; obox[3] = 1; }
            else if (g.Wt * g.Ht >= 32) { obox[1] = g.Wt; obox[2] = 32 / g.Wt; obox[3] = 1; }
            else { obox[1] = g.Wt; obox[2] = g.Ht; obox[3] = 32 / (g.Wt * g.Ht); }
            cuuint64_t odims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)d->B};
            cuuint64_t ostr[3] = {(cuuint64_t)g.ldc * es_out, (cuuint64_t)g.ldc * es_out * g.W,
                                  (cuuint64_t)g.y_batch_stride * es_out};
            CUresult ro = encode(&p.out_map[i], p.out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                                 4, g.y, odims, ostr, obox, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 p.out_f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (ro != CUDA_SUCCESS) g.tma_store = 0;
            else p.any_tma_store = 1;
        }
    }
    {
        const int nz = d->weight_per_sample ? d->B : d->kh * d->kw;
        cuuint64_t dims[3] = {(cuuint64_t)Kpad, (cuuint64_t)Npad, (cuuint64_t)nz};
        cuuint64_t strides[2] = {(cuuint64_t)Kpad * 2, (cuuint64_t)Kpad * 2 * Npad};
        cuuint32_t box[3] = {(cuuint32_t)kTileK, (cuuint32_t)bn, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = encode(&p.b_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(d->weight_bf16), dims,
                            strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(EFFDET_E_CUDA, "effdet_conv2d: cuTensorMapEncodeTiled(B) failed %s(%lld)", "", (long long)r);
    }
    p.n_tiles = Npad / bn;
    p.total_tiles = tiles * p.n_tiles;
    const int ctas_per_sm = (smem <= 113 * 1024 && p.tmem_cols <= 256 && !p.halo) ? 2 : 1;
    int grid = kNumSMs * ctas_per_sm;
    if (grid > p.total_tiles) grid = p.total_tiles;
    static bool attr_set = false;
#define TC_INST(A, F)                                                                                   \
    do {                                                                                                \
        if (!attr_set)                                                                                  \
            EFFDET_CUDA(cudaFuncSetAttribute(conv_tc_kernel<A, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             227 * 1024));                                              \
    } while (0)
    if (!attr_set) {
        TC_INST(EFFDET_ACT_NONE, false); TC_INST(EFFDET_ACT_RELU, false); TC_INST(EFFDET_ACT_SWISH, false);
        TC_INST(EFFDET_ACT_SIGMOID, false); TC_INST(EFFDET_ACT_NONE, true); TC_INST(EFFDET_ACT_RELU, true);
        TC_INST(EFFDET_ACT_SWISH, true); TC_INST(EFFDET_ACT_SIGMOID, true);
        attr_set = true;
    }
#undef TC_INST
    cudaStream_t st = as_stream(stream);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool pdl = pdl_level() >= 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
#define TC_LAUNCH(A)                                                                      \
    case A:                                                                               \
        if (p.out_f32) EFFDET_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<A, true>, p)); \
        else EFFDET_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<A, false>, p));          \
        break;
    switch (d->act) {
        TC_LAUNCH(EFFDET_ACT_NONE) TC_LAUNCH(EFFDET_ACT_RELU) TC_LAUNCH(EFFDET_ACT_SWISH) TC_LAUNCH(EFFDET_ACT_SIGMOID)
        default: return fail(EFFDET_E_INVALID, "effdet_conv2d: bad activation%s", "");
    }
#undef TC_LAUNCH
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

static bool wg_halo(const effdet_wgrad_desc *d) {
    return d->kh == 3 && d->kw == 3 && d->stride == 1 && d->Cin <= 64 && getenv("EFFDET_NO_WGRAD_HALO") == nullptr;
}
// tiles of the halo kernel: Wt in {8, 16, 32} (16-pixel K steps start on swizzle-atom boundaries)
// EFFDET_WG_HALO_PX=64: half-size tiles of the Cin <= 64 halo kernel (four 44 KiB ring stages instead of two of
// 76 KiB).  Measured 4-10 % SLOWER than 128-pixel tiles on the D0 heads (A/B in one job): the main loop is not
// starved by load latency; kept as a switch for the experiment.
static int wg_halo_px() {
    static const int px = getenv("EFFDET_WG_HALO_PX") ? atoi(getenv("EFFDET_WG_HALO_PX")) : 128;
    return px == 128 ? 128 : 64;
}
static void pick_tile_halo(int W, int H, int B, int *Wt, int *Ht, int *Bt, int px = 128) {
    static const int cand128[][3] = {{16, 8, 1}, {8, 16, 1}, {32, 4, 1}, {8, 8, 2}, {16, 4, 2}, {8, 4, 4}, {16, 2, 4},
                                     {8, 2, 8}, {16, 1, 8}};  // Wt * Ht >= 16: a K step never straddles images
    static const int cand64[][3] = {{16, 4, 1}, {8, 8, 1}, {32, 2, 1}, {8, 4, 2}, {16, 2, 2}, {8, 2, 4}, {16, 1, 4},
                                    {8, 8, 1}, {8, 8, 1}};
    const int (*cand)[3] = px == 64 ? cand64 : cand128;
    long best = -1;
    for (int q = 0; q < 9; ++q) {
        const int *c = cand[q];
        long n = (long)cdiv(W, c[0]) * cdiv(H, c[1]) * cdiv(B, c[2]);
        if (best < 0 || n < best) { best = n; *Wt = c[0]; *Ht = c[1]; *Bt = c[2]; }
    }
}

// 3x3, Cin > 64: column-of-taps form of conv_wgrad_tc_kernel (WgTcParams::col_taps)
static bool wg_col(const effdet_wgrad_desc *d) {
    return d->kh == 3 && d->kw == 3 && d->stride == 1 && d->Cin > 64 && getenv("EFFDET_NO_WGRAD_COL") == nullptr;
}
// N tile of the column-of-taps form: three accumulators must fit the 512 TMEM columns (<= 160 each, multiples of
// 16 = the UMMA N granularity at M = 128); equal tiles so that no CTA is mostly padding
static int wg_col_block_n(int Cout) {
    const int nt = (Cout + 159) / 160;
    return round_up((Cout + nt - 1) / nt, 16);
}

static int wg_tc_plan(const effdet_wgrad_desc *d, int *tiles_per_split_out, int *n_tiles_out, int *Wt, int *Ht,
                      int *Bt, int *block_n_out, int *m_tiles_out) {
    const int taps = d->kh * d->kw;
    const bool col = wg_col(d);
    const int bn = col ? wg_col_block_n(d->Cout) : (d->Cout <= 256 ? round_up(d->Cout, 64) : 256);
    const int m_tiles = (d->Cin + 127) / 128, n_tiles_n = (d->Cout + bn - 1) / bn;
    long total_tiles = 0;
    const bool halo = wg_halo(d);
    for (int i = 0; i < d->n_groups; ++i) {
        if (halo || col) pick_tile_halo(d->W[i], d->H[i], d->B, &Wt[i], &Ht[i], &Bt[i], halo ? wg_halo_px() : 128);
        else pick_tile(d->W[i], d->H[i], d->B, &Wt[i], &Ht[i], &Bt[i]);
        n_tiles_out[i] = (int)(cdiv(d->W[i], Wt[i]) * cdiv(d->H[i], Ht[i]) * cdiv(d->B, Bt[i]));
        total_tiles += n_tiles_out[i];
    }
    // ~2 CTAs per SM in flight (halo kernel: one CTA per SM, all taps in it; N tiles of 64);
    // at least 8 pixel tiles per CTA so the pipeline has work
    const long yz = halo ? (long)(d->Cout + 63) / 64
                    : col ? (long)3 * m_tiles * n_tiles_n
                          : (long)((d->Cin <= 64 && taps > 1) ? (taps + 1) / 2 : taps * m_tiles) * n_tiles_n;
    // ONE balanced wave: splits * yz <= resident CTAs (1 per SM for the halo / column forms, 2 otherwise); the
    // splits cut the FLAT tile list of all groups into equal ranges, so no CTA is a short remainder
    long want = ((long)kNumSMs * ((halo || col) ? 1 : 2)) / yz;
    if (want < 1) want = 1;
    long tps = (total_tiles + want - 1) / want;
    if (tps < (halo && wg_halo_px() == 64 ? 16 : 8)) tps = halo && wg_halo_px() == 64 ? 16 : 8;
    *tiles_per_split_out = (int)tps;
    *block_n_out = bn; *m_tiles_out = m_tiles;
    return (int)cdiv(total_tiles, tps);
}

extern "C" int effdet_conv_wgrad_tc_fuses_bias(const effdet_wgrad_desc *d) {
    return d && d->Cin <= 64 && d->kh == d->kw && ((d->kh * d->kw) & 1) && d->kh * d->kw > 1;
}

extern "C" int effdet_conv_wgrad_tc_splits(const effdet_wgrad_desc *d) {
    if (!d || d->n_groups < 1 || d->n_groups > kTcMaxGroups) return 0;
    int tps, nt[kTcMaxGroups], Wt[kTcMaxGroups], Ht[kTcMaxGroups], Bt[kTcMaxGroups], bn, mt;
    return wg_tc_plan(d, &tps, nt, Wt, Ht, Bt, &bn, &mt);
}

/* Tensor-core weight gradient: x (B,H,W,Cin) bf16 dense, dz (B,H,W,dz_ld) bf16 with the first Cout
 * channels meaningful (dz_ld >= Cout, multiple of 8), stride 1, k in {1,3}.  partial must hold
 * effdet_conv_wgrad_tc_splits(desc) * kh*kw*Cin*Cout floats. */
extern "C" int effdet_conv_wgrad_tc(const effdet_wgrad_desc *d, void *stream) {
    EFFDET_REQUIRE(d, "null descriptor");
    EFFDET_REQUIRE(d->n_groups >= 1 && d->n_groups <= kTcMaxGroups, "1..5 groups");
    EFFDET_REQUIRE(d->B > 0 && d->Cin > 0 && d->Cout > 0 && d->dweight && d->partial, "bad arguments");
    EFFDET_REQUIRE(d->stride == 1 && d->kh == d->kw && (d->kh == 1 || d->kh == 3), "stride 1, k in {1,3}");
    EFFDET_REQUIRE(d->x_dtype == EFFDET_BF16 && d->dz_dtype == EFFDET_BF16, "bf16 operands");
    EFFDET_REQUIRE(d->Cin % 8 == 0, "Cin must be a multiple of 8");
    EncodeTiledFn encode = get_encode();
    if (!encode) return fail(EFFDET_E_CUDA, "effdet_conv_wgrad_tc: cuTensorMapEncodeTiled unavailable%s", "");
    static WgTcParams p;
    memset(&p, 0, sizeof(p));
    int tps, nt[kTcMaxGroups], Wt[kTcMaxGroups], Ht[kTcMaxGroups], Bt[kTcMaxGroups], bn, mt;
    const int splits = wg_tc_plan(d, &tps, nt, Wt, Ht, Bt, &bn, &mt);
    EFFDET_REQUIRE(splits == d->n_splits, "n_splits must be effdet_conv_wgrad_tc_splits()");
    if (wg_halo(d)) {
        static WgHaloParams hp;
        memset(&hp, 0, sizeof(hp));
        hp.n_groups = d->n_groups; hp.Cin = d->Cin; hp.Cout = d->Cout; hp.partial = d->partial;
        hp.bias_partial = d->dbias ? d->partial + (size_t)splits * 9 * d->Cin * d->Cout : nullptr;
        int box_max = 0, zs = 0;
        for (int i = 0; i < d->n_groups; ++i) {
            WgHaloGroup &g = hp.g[i];
            g.Wt = Wt[i]; g.Ht = Ht[i]; g.Bt = Bt[i];
            g.lw = 31 - __builtin_clz(g.Wt); g.lh = 31 - __builtin_clz(g.Ht);
            g.tiles_x = cdiv(d->W[i], g.Wt); g.tiles_y = cdiv(d->H[i], g.Ht);
            g.n_tiles = nt[i]; g.tile_begin = zs;
            zs += nt[i];
            g.box_bytes = g.Wt * (g.Ht + 2) * g.Bt * 128;
            if (g.box_bytes > box_max) box_max = g.box_bytes;
            const long long ldz = d->dz_ld[i] ? d->dz_ld[i] : d->Cout;
            const long long zbs = d->dz_batch_stride[i] ? d->dz_batch_stride[i] : (long long)d->H[i] * d->W[i] * ldz;
            EFFDET_REQUIRE((ldz * 2) % 16 == 0 && (zbs * 2) % 16 == 0, "dz strides must be multiples of 16 bytes");
            EFFDET_REQUIRE(((reinterpret_cast<uintptr_t>(d->x[i]) | reinterpret_cast<uintptr_t>(d->dz[i])) & 15) == 0,
                           "operands must be 16-byte aligned");
            cuuint32_t es[4] = {1, 1, 1, 1};
            {
                cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W[i], (cuuint64_t)d->H[i], (cuuint64_t)d->B};
                cuuint64_t stx[3] = {(cuuint64_t)d->Cin * 2, (cuuint64_t)d->Cin * 2 * d->W[i],
                                     (cuuint64_t)d->Cin * 2 * d->W[i] * d->H[i]};
                cuuint32_t box[4] = {64, (cuuint32_t)g.Wt, (cuuint32_t)(g.Ht + 2), (cuuint32_t)g.Bt};
                CUresult r = encode(&hp.x_map[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(d->x[i]), dims,
                                    stx, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) return fail(EFFDET_E_CUDA, "effdet_conv_wgrad_tc: encode(x halo) failed %s(%lld)", "", (long long)r);
            }
            {
                cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->W[i], (cuuint64_t)d->H[i], (cuuint64_t)d->B};
                cuuint64_t stz[3] = {(cuuint64_t)ldz * 2, (cuuint64_t)ldz * 2 * d->W[i], (cuuint64_t)zbs * 2};
                cuuint32_t box[4] = {64, (cuuint32_t)g.Wt, (cuuint32_t)g.Ht, (cuuint32_t)g.Bt};
                CUresult r = encode(&hp.z_map[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(d->dz[i]), dims,
                                    stz, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) return fail(EFFDET_E_CUDA, "effdet_conv_wgrad_tc: encode(dz) failed %s(%lld)", "", (long long)r);
            }
        }
        hp.total_tiles = zs; hp.tiles_per_split = tps;
        { static const int dbg = getenv("EFFDET_WG_DBG") ? atoi(getenv("EFFDET_WG_DBG")) : 0; hp.dbg = dbg; }
        zs = splits;                                  // below: number of split rows
        hp.box_stride = round_up(box_max, 1024);
        hp.ksteps = wg_halo_px() / 16; hp.z_bytes = wg_halo_px() * 128;
        const int stage_bytes = 3 * hp.box_stride + hp.z_bytes;
        hp.stages = (int)((220 * 1024 - hp.box_stride - 2048) / stage_bytes);
        if (hp.stages > 4) hp.stages = 4;
        EFFDET_REQUIRE(hp.stages >= 1, "tile does not fit shared memory");
        const size_t hsmem = (size_t)hp.stages * stage_bytes + hp.box_stride + (2 * hp.stages + 1) * 8 + 16 + 1024;
        static bool hattr = false;
        if (!hattr) {
            EFFDET_CUDA(cudaFuncSetAttribute(conv_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            hattr = true;
        }
        cudaStream_t hst = as_stream(stream);
        dim3 hgrid(zs, (d->Cout + 63) / 64);
        EFFDET_CUDA(launch_pdl(conv_wgrad_halo_kernel, dim3(hgrid), dim3(192), hsmem, hst, hp));
        EFFDET_LAUNCHED();
        const size_t hn = (size_t)9 * d->Cin * d->Cout;
        const int hn_bias = d->dbias ? d->Cout : 0;
        EFFDET_CUDA(launch_pdl(wgrad_tc_reduce_kernel, dim3(cdiv(hn + hn_bias, 256 / wg_reduce_slices(zs))), dim3(256), 0, hst, d->partial, zs, hn, d->dweight, d->accumulate,
                                                                        hp.bias_partial, hn_bias, d->dbias, wg_reduce_slices(zs)));
        EFFDET_LAUNCHED();
        return EFFDET_OK;
    }
    p.n_groups = d->n_groups; p.B = d->B; p.Cin = d->Cin; p.Cout = d->Cout; p.ksize = d->kh;
    p.m_tiles = mt; p.block_n = bn; p.partial = d->partial;
    p.bias_partial = d->dbias ? d->partial + (size_t)splits * d->kh * d->kw * d->Cin * d->Cout : nullptr;
    p.pair_taps = (d->Cin <= 64 && d->kh * d->kw > 1) ? 1 : 0;
    EFFDET_REQUIRE(!d->dbias || effdet_conv_wgrad_tc_fuses_bias(d), "dbias only when effdet_conv_wgrad_tc_fuses_bias()");
    p.col_taps = wg_col(d) ? 1 : 0;
    if (p.col_taps) {
        int box_max = 0;
        for (int i = 0; i < d->n_groups; ++i) box_max = max(box_max, Wt[i] * (Ht[i] + 2) * Bt[i] * 128);
        p.a_box_stride = round_up(box_max, 1024);
    }
    p.tmem_cols = p.col_taps ? 512 : (bn <= 64 ? 64 : bn <= 128 ? 128 : 256);
    const int stage_bytes = (p.col_taps ? 2 * p.a_box_stride : 2 * kATileBytes) + ((bn + 63) / 64) * kATileBytes;
    int stages = ((p.col_taps ? 222 : 200) * 1024) / stage_bytes;
    if (stages > 4) stages = 4;
    if (stages < 1) stages = 1;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + (2 * stages + 1) * 8 + 16 + 1024;
    int z = 0;
    for (int i = 0; i < d->n_groups; ++i) {
        WgTcGroup &g = p.g[i];
        g.H = d->H[i]; g.W = d->W[i]; g.Wt = Wt[i]; g.Ht = Ht[i]; g.Bt = Bt[i];
        g.tiles_x = cdiv(g.W, g.Wt); g.tiles_y = cdiv(g.H, g.Ht); g.tiles_b = cdiv(d->B, g.Bt);
        g.n_tiles = nt[i]; g.tile_begin = z;
        g.lw = 31 - __builtin_clz(g.Wt); g.lh = 31 - __builtin_clz(g.Ht);
        g.box_bytes = g.Wt * (g.Ht + 2) * g.Bt * 128;
        z += nt[i];
        const long long ldz = d->dz_ld[i] ? d->dz_ld[i] : d->Cout;
        const long long zbs = d->dz_batch_stride[i] ? d->dz_batch_stride[i] : (long long)g.H * g.W * ldz;
        EFFDET_REQUIRE((ldz * 2) % 16 == 0 && (zbs * 2) % 16 == 0, "dz strides must be multiples of 16 bytes");
        EFFDET_REQUIRE(((reinterpret_cast<uintptr_t>(d->x[i]) | reinterpret_cast<uintptr_t>(d->dz[i])) & 15) == 0,
                       "operands must be 16-byte aligned");
        cuuint32_t box[4] = {64, (cuuint32_t)g.Wt, (cuuint32_t)g.Ht, (cuuint32_t)g.Bt};
        cuuint32_t xbox[4] = {64, (cuuint32_t)g.Wt, (cuuint32_t)(g.Ht + (p.col_taps ? 2 : 0)), (cuuint32_t)g.Bt};
        cuuint32_t es[4] = {1, 1, 1, 1};
        {
            cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)d->B};
            cuuint64_t st[3] = {(cuuint64_t)d->Cin * 2, (cuuint64_t)d->Cin * 2 * g.W, (cuuint64_t)d->Cin * 2 * g.W * g.H};
            CUresult r = encode(&p.x_map[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(d->x[i]), dims, st,
                                xbox, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(EFFDET_E_CUDA, "effdet_conv_wgrad_tc: encode(x) failed %s(%lld)", "", (long long)r);
        }
        {
            // only the first Cout channels are meaningful: the map's channel extent is Cout so that
            // everything beyond it is zero-filled by TMA
            cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)d->B};
            cuuint64_t st[3] = {(cuuint64_t)ldz * 2, (cuuint64_t)ldz * 2 * g.W, (cuuint64_t)zbs * 2};
            CUresult r = encode(&p.z_map[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(d->dz[i]), dims, st,
                                box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(EFFDET_E_CUDA, "effdet_conv_wgrad_tc: encode(dz) failed %s(%lld)", "", (long long)r);
        }
    }
    static bool attr_set = false;
    if (!attr_set) {
        EFFDET_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const int taps = d->kh * d->kw;
    p.total_tiles = z; p.tiles_per_split = tps;
    z = splits;
    dim3 grid(z, p.col_taps ? 3 * mt : (p.pair_taps ? (taps + 1) / 2 : taps * mt), (d->Cout + bn - 1) / bn);
    cudaStream_t st = as_stream(stream);
    EFFDET_CUDA(launch_pdl(conv_wgrad_tc_kernel, dim3(grid), dim3(192), smem, st, p));
    EFFDET_LAUNCHED();
    const size_t n = (size_t)taps * d->Cin * d->Cout;
    const int n_bias = d->dbias ? d->Cout : 0;
    EFFDET_CUDA(launch_pdl(wgrad_tc_reduce_kernel, dim3(cdiv(n + n_bias, 256 / wg_reduce_slices(z))), dim3(256), 0, st, d->partial, z, n, d->dweight, d->accumulate,
                                                                 p.bias_partial, n_bias, d->dbias, wg_reduce_slices(z)));
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}
