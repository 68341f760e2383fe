// Depthwise k x k convolution + folded BN + activation (+ squeeze-excite partial sums) for bf16
// NHWC tensors on sm_100a -- the HBM-bound half of every MBConv block (efficientnet.py:242-260)
// and the BiFPN DepthwiseConvBlock (model.py:48-68).  SURVEY section 8(a) row 4.
//
// Persistent blocks walk (x tile, y tile, channel block, image) tiles.  One elected thread feeds a
// two-deep shared-memory ring with TMA: the (8*S+K-S) x (16*S+K-S) x CB input patch of the next tile
// (out-of-image rows / columns / channels are zero-filled by the tensor map == TF "SAME"
// padding, asymmetric for stride 2) and its K*K x CB weight slice land while the current tile
// is being computed, so a block always has one tile of loads in flight and two resident blocks
// per SM keep ~60-170 KB per SM outstanding.
// A thread owns ONE channel pair (a bf16x2 word == one packed fp32x2 FFMA2 operand) and a 2 x 4
// register tile of outputs; its K*K weights live in registers for the whole tile, every input
// word is read from shared memory once and feeds up to 2*K FFMA2s, so the kernel needs ~12 bytes
// of shared-memory traffic per output element instead of ~45 for a vector-per-thread layout
// (which measured smem/LSU-bound: ncu short_scoreboard + mio_throttle, profiles/).  Lanes run
// along the channel pairs: a warp reads / writes whole 64..128-byte pixel rows.
// SE partial sums are reduced in a fixed order (registers -> shared rows -> one (image, tile,
// channel) partial) so the result is bit-reproducible run to run.
#define EFFDET_PDL_TU_LEVEL 2
#include "common.cuh"
#include "tma.cuh"

namespace effdet {

constexpr int kRtH = 2, kRtW = 4;          // outputs per thread (register tile)
#ifndef EFFDET_DW_RTH
#define EFFDET_DW_RTH 4
#endif

struct alignas(64) DwTmaParams {
    CUtensorMap x_map;                     // (C, W, H, B) bf16, box (CB, IW, IH, 1)
    CUtensorMap w_map;                     // (C, K*K) f32,      box (CB, K*K)
    CUtensorMap y_map;                     // (C, Wo, Ho, B) bf16, box (CB, TW, TH, 1): output tiles leave through TMA
    const float *scale, *shift;
    __nv_bfloat16 *y;
    float *se_sum;                         // [image][tile][C] sums of the outputs (SE squeeze), or NULL
    float *stats;                          // [image*tile][2][C] sum / sum of squares of the outputs (BN batch
                                           // statistics of the following BatchNormalization), or NULL
    int Ho, Wo, C, pad_t, pad_l, tiles_x, tiles_y, cblocks, total_tiles;
    int run;                               // tiles a block takes in a row before it strides by gridDim.x runs
};

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ float tanh_approx_f(float x) {
    float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}

// Strided walk over tiles t = first, first + stride, ... with coordinates (x, y, c, b) = t decomposed over the
// extents (nx, ny, nc): the stride is decomposed once and added with carries -- no integer division per tile
// (a division per coordinate is an I2F / MUFU.RCP / IMAD.HI chain of ~35 dependent instructions).
struct TileWalk {
    int t, x, y, c, b;
    int jx, jy, jc, jb, stride;
    __device__ __forceinline__ void init(int first, int stride_, int nx, int ny, int nc) {
        stride = stride_;
        int r = stride_;
        jx = r % nx; r /= nx; jy = r % ny; r /= ny; jc = r % nc; jb = r / nc;
        t = first; r = first;
        x = r % nx; r /= nx; y = r % ny; r /= ny; c = r % nc; b = r / nc;
    }
    __device__ __forceinline__ void next(int nx, int ny, int nc) {
        t += stride;
        x += jx; int carry = x >= nx; x -= carry ? nx : 0;
        y += jy + carry; carry = y >= ny; y -= carry ? ny : 0;
        c += jc + carry; carry = c >= nc; c -= carry ? nc : 0;
        b += jb + carry;
    }
};

// S = 1: 8 x 16 output tile, two register-tile passes per thread; S = 2: 8 x 8 tile, one pass
// register-tile height of the forward kernel: 4 rows for the stride-1 kernels (a 4 x 4 tile reads (3+K)^2 words
// for 16 outputs instead of (1+K)(3+K) for 8: 5x5 goes from 18 to 12 shared-memory load + unpack instructions per
// output pair, below its 12.5 FFMA2, so the FMA pipe -- 2 cycles per FFMA2 -- becomes the limiter)
template <int K, int S> struct FwdRt { static constexpr int H = S == 1 ? EFFDET_DW_RTH : 2; };

template <int K, int S, int CP, int RTH = kRtH> struct DwCfg {
    static constexpr int RT_H = RTH;
    static constexpr int TH = 8, TW = S == 1 ? 16 : 8;
    static constexpr int CB = CP * 2;
    static constexpr int IH = (TH - 1) * S + K, IW = (TW - 1) * S + K;
    static constexpr int NRT = (TH / RTH) * (TW / kRtW);               // register tiles per tile
    static constexpr int PASSES = NRT / 8;
    static constexpr int NT = CP * 8;
    static constexpr int IR = (RTH - 1) * S + K, NIN = (kRtW - 1) * S + K;
    static constexpr int IN_BYTES = IH * IW * CB * 2;
    static constexpr int IN_PAD = ((IN_BYTES + 127) / 128) * 128;       // TMA destinations are 128-byte aligned
    static constexpr int W_BYTES = K * K * CB * 4;
    static constexpr int STAGE_BYTES = ((IN_PAD + W_BYTES + 127) / 128) * 128;
    static constexpr int OUT_BYTES = TH * TW * CB * 2;                  // staged output tile [TH][TW][CB] bf16
    static constexpr int OUT_PAD = ((OUT_BYTES + 127) / 128) * 128;
    static constexpr size_t SMEM = 2 * (size_t)STAGE_BYTES + 2 * (size_t)OUT_PAD + 2 * 2 * 8 * CB * 4 + 64 + 128;
};

enum { DW_SUMS_NONE = 0, DW_SUMS_SE = 1, DW_SUMS_STATS = 2 };

// Epilogue (round 2): folded BN + activation in registers, bf16 pack, ONE 4-byte shared-memory store per output
// pair into a staged [TH][TW][CB] tile; after the tile barrier one thread hands the tile to the TMA engine
// (cp.async.bulk.tensor store: full 128-byte bursts, rows / columns / channels outside the tensor are clipped by
// the tensor map).  The round-1 epilogue stored straight to global memory: per output pair a 64-bit address
// IMAD chain, two bounds predicates and an STG -- 186 of the 330 issue slots of a 3x3 pass (profiles/
// r2_dwconv_sass.txt); the kernel was issue-bound at 53 % with DRAM at 30-37 %.  Two staging tiles: the store of
// tile i drains while tile i+1 is computed.  SUMS selects the per-tile reductions compiled in: none, the SE
// squeeze sums, or sum + sum of squares (BatchNorm batch statistics of the following layer).
template <int K, int S, int CP, int ACT, int SUMS>
__global__ void __launch_bounds__(DwCfg<K, S, CP, FwdRt<K, S>::H>::NT, (K == 3 ? 3 : 2))
dwconv_tma_kernel(const __grid_constant__ DwTmaParams p) {
    using Cfg = DwCfg<K, S, CP, FwdRt<K, S>::H>;
    constexpr int RTH = Cfg::RT_H;
    constexpr int CB = Cfg::CB, IW = Cfg::IW, NIN = Cfg::NIN, IR = Cfg::IR, TW = Cfg::TW, TH = Cfg::TH;
    extern __shared__ uint8_t dsm_raw[];
    // pointer arithmetic on the __shared__ array (not through uintptr_t) keeps LDS/STS addressing
    uint8_t *dsm = dsm_raw + ((128u - (smem_u32(dsm_raw) & 127u)) & 127u);
    uint8_t *sOut = dsm + 2 * Cfg::STAGE_BYTES;                                // [2][TH][TW][CB] bf16
    float *sRed = reinterpret_cast<float *>(sOut + 2 * Cfg::OUT_PAD);          // [2 parities][2 moments][8][CB]
    uint64_t *full = reinterpret_cast<uint64_t *>(sRed + 2 * 2 * 8 * CB);

    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    EFFDET_PDL_SYNC();

    // Tile order: x, y, channel block, image.  Block i takes runs of p.run consecutive tiles: run i, i + grid, ...
    // (block-cyclic), so the tiles in flight at any time are neighbours -- halo rows and the 128-byte lines two
    // channel blocks share are L2 hits (a contiguous range per block measured 2.3x the algorithmic DRAM reads on
    // 48-channel blocks) -- while a block still sees the same channel block for a whole run.  Tile coordinates
    // advance by add-with-carry of a pre-decomposed step: round 1 decomposed every tile index with three integer
    // divisions, twice (once for the TMA issue): 195 of the instructions of a 5x5 tile but 31 % of its stall
    // samples (profiles/r2_dwconv_ncu.txt: I2F / MUFU.RCP / IMAD.HI chains ahead of the barrier wait).  The K*K
    // weights and the folded-BN constants are re-read into registers only when the channel block changes.
    auto issue = [&](int tx, int ty, int cb, int b, int stage) {
        uint8_t *dst = dsm + (size_t)stage * Cfg::STAGE_BYTES;
        mbar_expect_tx(&full[stage], (uint32_t)(Cfg::IN_BYTES + Cfg::W_BYTES));
        tma_load_4d(dst, &p.x_map, &full[stage], cb * CB, tx * TW * S - p.pad_l, ty * TH * S - p.pad_t, b);
        tma_load_2d(dst + Cfg::IN_PAD, &p.w_map, &full[stage], cb * CB, 0);
    };

    const int pair = tid % CP, slot = tid / CP;        // slot 0..7: which register tile of a pass
    const int run = p.run;
    // step from the last tile of a run to the first tile of this block's next run, decomposed once
    int jx, jy, jc, jb;
    {
        int r = (int)(gridDim.x - 1) * run + 1;
        jx = r % p.tiles_x; r /= p.tiles_x;
        jy = r % p.tiles_y; r /= p.tiles_y;
        jc = r % p.cblocks; jb = r / p.cblocks;
    }
    int tx, ty, cb, b;                                 // current tile
    int nt = blockIdx.x * run, ntx, nty, ncb, nb, nk = 0;   // next tile to issue: index, coordinates, position in run
    {
        int r = nt;
        ntx = r % p.tiles_x; r /= p.tiles_x;
        nty = r % p.tiles_y; r /= p.tiles_y;
        ncb = r % p.cblocks; nb = r / p.cblocks;
    }
    auto advance = [&]() {
        const bool hop = ++nk == run;
        if (hop) nk = 0;
        nt += hop ? (int)(gridDim.x - 1) * run + 1 : 1;
        ntx += hop ? jx : 1;
        int carry = ntx >= p.tiles_x; ntx -= carry ? p.tiles_x : 0;
        nty += (hop ? jy : 0) + carry;
        carry = nty >= p.tiles_y; nty -= carry ? p.tiles_y : 0;
        ncb += (hop ? jc : 0) + carry;
        carry = ncb >= p.cblocks; ncb -= carry ? p.cblocks : 0;
        nb += (hop ? jb : 0) + carry;
    };
    if (tid == 0 && nt < p.total_tiles) issue(ntx, nty, ncb, nb, 0);
    int it = 0, cur_cb = -1;
    float2 wk[K * K];
    float2 sc = make_float2(0.f, 0.f), sh = make_float2(0.f, 0.f);
    while (nt < p.total_tiles) {
        const int stage = it & 1;
        tx = ntx; ty = nty; cb = ncb; b = nb;
        advance();
        if (tid == 0 && nt < p.total_tiles) issue(ntx, nty, ncb, nb, stage ^ 1);
        const uint8_t *sIn = dsm + (size_t)stage * Cfg::STAGE_BYTES;
        const float *sW = reinterpret_cast<const float *>(sIn + Cfg::IN_PAD);
        uint8_t *so = sOut + (size_t)(it & 1) * Cfg::OUT_PAD + (size_t)pair * 4;
        // rows / columns of this tile inside the output (edge tiles: the sums must skip the rest; the store is
        // clipped by the tensor map)
        const int vh = min(TH, p.Ho - ty * TH), vw = min(TW, p.Wo - tx * TW);
        const bool whole = vh == TH && vw == TW;
        const bool new_cb = cb != cur_cb;
        if (new_cb) {
            cur_cb = cb;
            const int c = cb * CB + pair * 2;
            // per-channel epilogue constants (swish: the 1/2 of z/2 * (1 + tanh(z/2)) is folded in)
            sc = make_float2(0.f, 0.f); sh = make_float2(0.f, 0.f);
            if (c < p.C) {
                const float pre = ACT == EFFDET_ACT_SWISH ? 0.5f : 1.f;
                sc = *reinterpret_cast<const float2 *>(p.scale + c);
                sh = *reinterpret_cast<const float2 *>(p.shift + c);
                sc.x *= pre; sc.y *= pre; sh.x *= pre; sh.y *= pre;
            }
        }
        mbar_wait(&full[stage], (it >> 1) & 1);
        if (new_cb) {
#pragma unroll
            for (int i = 0; i < K * K; ++i) wk[i] = *reinterpret_cast<const float2 *>(sW + i * CB + pair * 2);
        }

        float2 tot = make_float2(0.f, 0.f), tot2 = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int pass = 0; pass < Cfg::PASSES; ++pass) {
            const int rt = pass * 8 + slot;
            const int ry = (rt / (TW / kRtW)) * RTH, rx = (rt % (TW / kRtW)) * kRtW;   // output origin in the tile
            float2 acc[RTH][kRtW];
#pragma unroll
            for (int i = 0; i < RTH; ++i)
#pragma unroll
                for (int j = 0; j < kRtW; ++j) acc[i][j] = make_float2(0.f, 0.f);
            const uint8_t *base = sIn + ((size_t)((ry * S) * IW + rx * S) * CB + pair * 2) * 2;
#pragma unroll
            for (int rr = 0; rr < IR; ++rr) {
                float2 in[NIN];
#pragma unroll
                for (int j = 0; j < NIN; ++j) {
                    const uint32_t w2 = *reinterpret_cast<const uint32_t *>(base + (size_t)(rr * IW + j) * CB * 2);
                    in[j] = make_float2(__uint_as_float(w2 << 16), __uint_as_float(w2 & 0xffff0000u));
                }
#pragma unroll
                for (int orow = 0; orow < RTH; ++orow) {
                    const int ky = rr - orow * S;
                    if (ky < 0 || ky >= K) continue;
#pragma unroll
                    for (int kx = 0; kx < K; ++kx)
#pragma unroll
                        for (int oc = 0; oc < kRtW; ++oc)
                            acc[orow][oc] = __ffma2_rn(in[oc * S + kx], wk[ky * K + kx], acc[orow][oc]);
                }
            }
            uint8_t *srow = so + (size_t)(ry * TW + rx) * CB * 2;
#pragma unroll
            for (int orow = 0; orow < RTH; ++orow) {
#pragma unroll
                for (int oc = 0; oc < kRtW; ++oc) {
                    float2 z = __ffma2_rn(acc[orow][oc], sc, sh);
                    if (ACT == EFFDET_ACT_SWISH) {
                        z = __ffma2_rn(z, make_float2(tanh_approx_f(z.x), tanh_approx_f(z.y)), z);
                    } else if (ACT == EFFDET_ACT_RELU) {
                        z.x = fmaxf(z.x, 0.f); z.y = fmaxf(z.y, 0.f);
                    }
                    if (SUMS != DW_SUMS_NONE) {
                        if (whole || (ry + orow < vh && rx + oc < vw)) {
                            tot = __fadd2_rn(tot, z);
                            if (SUMS == DW_SUMS_STATS) tot2 = __ffma2_rn(z, z, tot2);
                        }
                    }
                    *reinterpret_cast<__nv_bfloat162 *>(srow + (size_t)(orow * TW + oc) * CB * 2) =
                        __floats2bfloat162_rn(z.x, z.y);
                }
            }
        }
        fence_proxy_async();                       // this thread's staged outputs -> visible to the TMA engine
        float *red = sRed + (size_t)(it & 1) * 2 * 8 * CB;
        if (SUMS != DW_SUMS_NONE) {
            *reinterpret_cast<float2 *>(red + slot * CB + pair * 2) = tot;
            if (SUMS == DW_SUMS_STATS) *reinterpret_cast<float2 *>(red + (8 + slot) * CB + pair * 2) = tot2;
        }
        if (tid == 0) tma_store_wait_read<0>();    // the previous tile's store has left the other staging tile
        __syncthreads();                           // tile staged; everyone is done reading this input stage
        if (tid == 0)
            tma_store_4d(&p.y_map, sOut + (size_t)(it & 1) * Cfg::OUT_PAD, cb * CB, tx * TW, ty * TH, b);
        if (SUMS != DW_SUMS_NONE) {
            const int moment = tid / CB, ch = tid - moment * CB;        // NT = 4 * CB threads
            if (moment < (SUMS == DW_SUMS_STATS ? 2 : 1) && cb * CB + ch < p.C) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) s += red[(moment * 8 + w) * CB + ch];
                const size_t row = (size_t)b * p.tiles_x * p.tiles_y + (size_t)ty * p.tiles_x + tx;
                if (SUMS == DW_SUMS_STATS) p.stats[(row * 2 + moment) * p.C + cb * CB + ch] = s;
                else p.se_sum[row * p.C + cb * CB + ch] = s;
            }
        }
        ++it;
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // outstanding stores land before exit
}

template <int K, int S, int CP>
static int launch_dw_tma(const void *x, const float *w, const float *scale, const float *shift, void *y,
                         float *se_sum, float *stats, int B, int H, int W, int C, int act, cudaStream_t st) {
    using Cfg = DwCfg<K, S, CP, FwdRt<K, S>::H>;
    EncodeTiledFn encode = get_encode();
    if (!encode) return fail(EFFDET_E_CUDA, "effdet_dwconv: cuTensorMapEncodeTiled unavailable%s", "");
    DwTmaParams p;
    memset(&p, 0, sizeof(p));
    const int Ho = (H + S - 1) / S, Wo = (W + S - 1) / S;
    p.Ho = Ho; p.Wo = Wo; p.C = C;
    p.pad_t = max((Ho - 1) * S + K - H, 0) / 2;
    p.pad_l = max((Wo - 1) * S + K - W, 0) / 2;
    p.tiles_x = (Wo + Cfg::TW - 1) / Cfg::TW; p.tiles_y = (Ho + Cfg::TH - 1) / Cfg::TH;
    p.cblocks = (C + Cfg::CB - 1) / Cfg::CB;
    p.total_tiles = p.tiles_x * p.tiles_y * p.cblocks * B;
    p.scale = scale; p.shift = shift; p.y = static_cast<__nv_bfloat16 *>(y); p.se_sum = se_sum; p.stats = stats;
    {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
        cuuint32_t box[4] = {(cuuint32_t)Cfg::CB, (cuuint32_t)Cfg::IW, (cuuint32_t)Cfg::IH, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = encode(&p.x_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(x), dims, strides, box,
                            es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(EFFDET_E_CUDA, "effdet_dwconv: cuTensorMapEncodeTiled(x) failed %s(%lld)", "", (long long)r);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)(K * K)};
        cuuint64_t strides[1] = {(cuuint64_t)C * 4};
        cuuint32_t box[2] = {(cuuint32_t)Cfg::CB, (cuuint32_t)(K * K)};
        cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&p.w_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(w), dims, strides, box,
                            es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(EFFDET_E_CUDA, "effdet_dwconv: cuTensorMapEncodeTiled(w) failed %s(%lld)", "", (long long)r);
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * Wo, (cuuint64_t)C * 2 * Wo * Ho};
        cuuint32_t box[4] = {(cuuint32_t)Cfg::CB, (cuuint32_t)Cfg::TW, (cuuint32_t)Cfg::TH, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = encode(&p.y_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, y, dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(EFFDET_E_CUDA, "effdet_dwconv: cuTensorMapEncodeTiled(y) failed %s(%lld)", "", (long long)r);
    }
#define DWT_LAUNCH(A, SU)                                                                                  \
    {                                                                                                      \
        auto kern = dwconv_tma_kernel<K, S, CP, A, SU>;                                                    \
        static int per_sm = 0;                                                                             \
        if (!per_sm) {                                                                                     \
            EFFDET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM)); \
            int nb = 0;                                                                                    \
            EFFDET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::NT, Cfg::SMEM));     \
            per_sm = nb < 1 ? 1 : nb;                                                                      \
        }                                                                                                  \
        int grid = kNumSMs * per_sm;                                                                       \
        if (grid > p.total_tiles) grid = p.total_tiles;                                                    \
        /* runs: at least ~4 per block (balance), at most 8 tiles long */                                  \
        p.run = p.total_tiles / (grid * 4);                                                                \
        if (const char *e = getenv("EFFDET_DW_RUN")) p.run = atoi(e);                                      \
        p.run = p.run < 1 ? 1 : (p.run > 8 ? 8 : p.run);                                                   \
        EFFDET_CUDA(launch_pdl(kern, dim3(grid), dim3(Cfg::NT), Cfg::SMEM, st, p));                        \
    }
    // compiled combinations: swish (+ SE squeeze sums) = MBConv forward; linear (+ BN batch statistics) = raw
    // convolution of the training step / data gradients; ReLU = stand-alone DepthwiseConvBlock
    if (stats && se_sum) return fail(EFFDET_E_INVALID, "effdet_dwconv: SE sums and BN statistics are exclusive%s", "");
    if (act == EFFDET_ACT_SWISH && se_sum) DWT_LAUNCH(EFFDET_ACT_SWISH, DW_SUMS_SE)
    else if (act == EFFDET_ACT_SWISH && !stats) DWT_LAUNCH(EFFDET_ACT_SWISH, DW_SUMS_NONE)
    else if (act == EFFDET_ACT_NONE && stats) DWT_LAUNCH(EFFDET_ACT_NONE, DW_SUMS_STATS)
    else if (act == EFFDET_ACT_NONE && !se_sum) DWT_LAUNCH(EFFDET_ACT_NONE, DW_SUMS_NONE)
    else if (act == EFFDET_ACT_RELU && !stats && !se_sum) DWT_LAUNCH(EFFDET_ACT_RELU, DW_SUMS_NONE)
    else return fail(EFFDET_E_UNSUPPORTED, "effdet_dwconv: unsupported activation / reduction combination%s", "");
#undef DWT_LAUNCH
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

// ------------------------------------------------------------------ depthwise weight gradient
// dW[ky][kx][c] = sum_{b,oy,ox} x[b, oy*S - pad + ky, ox*S - pad + kx, c] * dz[b,oy,ox,c]
// Same tiling and thread layout as the forward kernel: block (split, channel block) walks spatial
// tiles split, split + nsplit, ...; TMA double-buffers the x patch (with halo, zero padded) and the
// dz tile (zero filled outside the tensor, so edge tiles need no masks); a thread keeps the K*K
// partial sums of its channel pair in registers over ALL its tiles.  Partials are written per
// split in a fixed order and summed by sum_partials_warp_kernel (deterministic).
struct alignas(64) DwWgParams {
    CUtensorMap x_map;                     // (C, W, H, B) bf16, box (CB, IW, IH, 1)
    CUtensorMap z_map;                     // (C, Wo, Ho, B) bf16, box (CB, TW, TH, 1)
    float *partial;                        // [nsplit][K*K][C]
    int C, pad_t, pad_l, tiles_x, tiles_y, spatial_tiles;
};

template <int K, int S, int CP>
__global__ void __launch_bounds__(DwCfg<K, S, CP>::NT, 2)
dw_wgrad_tma_kernel(const __grid_constant__ DwWgParams p) {
    using Cfg = DwCfg<K, S, CP>;
    constexpr int CB = Cfg::CB, IW = Cfg::IW, NIN = Cfg::NIN, IR = Cfg::IR, TW = Cfg::TW, TH = Cfg::TH;
    constexpr int Z_BYTES = TH * TW * CB * 2;
    constexpr int STAGE = ((Cfg::IN_PAD + Z_BYTES + 127) / 128) * 128;
    extern __shared__ uint8_t dsm_raw[];
    uint8_t *dsm = dsm_raw + ((128u - (smem_u32(dsm_raw) & 127u)) & 127u);
    uint64_t *full = reinterpret_cast<uint64_t *>(dsm + 2 * STAGE);
    float *sRed = reinterpret_cast<float *>(dsm);          // reused after the main loop: [8][K*K][CB]

    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    EFFDET_PDL_SYNC();
    const int cb = blockIdx.y;
    auto issue = [&](int tx, int ty, int b, int stage) {
        uint8_t *dst = dsm + (size_t)stage * STAGE;
        mbar_expect_tx(&full[stage], (uint32_t)(Cfg::IN_BYTES + Z_BYTES));
        tma_load_4d(dst, &p.x_map, &full[stage], cb * CB, tx * TW * S - p.pad_l, ty * TH * S - p.pad_t, b);
        tma_load_4d(dst + Cfg::IN_PAD, &p.z_map, &full[stage], cb * CB, tx * TW, ty * TH, b);
    };
    const int pair = tid % CP, slot = tid / CP;
    float2 acc[K * K];
#pragma unroll
    for (int i = 0; i < K * K; ++i) acc[i] = make_float2(0.f, 0.f);

    TileWalk nxt;                                      // the tile to issue next (x, y, image; c unused)
    nxt.init(blockIdx.x, gridDim.x, p.tiles_x, p.tiles_y, 1 << 30);
    if (tid == 0 && nxt.t < p.spatial_tiles) issue(nxt.x, nxt.y, nxt.c, 0);
    int it = 0;
    for (; nxt.t < p.spatial_tiles; ++it) {
        const int stage = it & 1;
        nxt.next(p.tiles_x, p.tiles_y, 1 << 30);
        if (tid == 0 && nxt.t < p.spatial_tiles) issue(nxt.x, nxt.y, nxt.c, stage ^ 1);
        const uint8_t *sIn = dsm + (size_t)stage * STAGE;
        const uint8_t *sZ = sIn + Cfg::IN_PAD;
        mbar_wait(&full[stage], (it >> 1) & 1);
#pragma unroll 1
        for (int pass = 0; pass < Cfg::PASSES; ++pass) {
            const int rt = pass * 8 + slot;
            const int ry = (rt / (TW / kRtW)) * kRtH, rx = (rt % (TW / kRtW)) * kRtW;
            float2 dzv[kRtH][kRtW];
#pragma unroll
            for (int i = 0; i < kRtH; ++i)
#pragma unroll
                for (int j = 0; j < kRtW; ++j) {
                    const uint32_t w2 = *reinterpret_cast<const uint32_t *>(
                        sZ + ((size_t)((ry + i) * TW + rx + j) * CB + pair * 2) * 2);
                    dzv[i][j] = make_float2(__uint_as_float(w2 << 16), __uint_as_float(w2 & 0xffff0000u));
                }
            const uint8_t *base = sIn + ((size_t)((ry * S) * IW + rx * S) * CB + pair * 2) * 2;
#pragma unroll
            for (int rr = 0; rr < IR; ++rr) {
                float2 in[NIN];
#pragma unroll
                for (int j = 0; j < NIN; ++j) {
                    const uint32_t w2 = *reinterpret_cast<const uint32_t *>(base + (size_t)(rr * IW + j) * CB * 2);
                    in[j] = make_float2(__uint_as_float(w2 << 16), __uint_as_float(w2 & 0xffff0000u));
                }
#pragma unroll
                for (int orow = 0; orow < kRtH; ++orow) {
                    const int ky = rr - orow * S;
                    if (ky < 0 || ky >= K) continue;
#pragma unroll
                    for (int kx = 0; kx < K; ++kx)
#pragma unroll
                        for (int oc = 0; oc < kRtW; ++oc)
                            acc[ky * K + kx] = __ffma2_rn(in[oc * S + kx], dzv[orow][oc], acc[ky * K + kx]);
                }
            }
        }
        __syncthreads();            // everyone is done reading this stage before it is refilled
    }
    // block reduction over the 8 register-tile slots (fixed order), then one partial row per split
#pragma unroll
    for (int i = 0; i < K * K; ++i)
        *reinterpret_cast<float2 *>(sRed + ((size_t)slot * K * K + i) * CB + pair * 2) = acc[i];
    __syncthreads();
    for (int i = tid; i < K * K * CB; i += Cfg::NT) {
        const int c = cb * CB + i % CB;
        if (c < p.C) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += sRed[(size_t)w * K * K * CB + i];
            p.partial[((size_t)blockIdx.x * K * K + i / CB) * p.C + c] = s;
        }
    }
}

template <int K, int S, int CP>
static int launch_dw_wgrad_tma(const void *x, const void *dz, float *partial, int nsplit, int B, int H, int W,
                               int C, cudaStream_t st) {
    using Cfg = DwCfg<K, S, CP>;
    constexpr int Z_BYTES = Cfg::TH * Cfg::TW * Cfg::CB * 2;
    constexpr int STAGE = ((Cfg::IN_PAD + Z_BYTES + 127) / 128) * 128;
    constexpr size_t RED = (size_t)8 * K * K * Cfg::CB * 4;
    constexpr size_t SMEM = (2 * (size_t)STAGE > RED ? 2 * (size_t)STAGE : RED) + 64 + 128;
    EncodeTiledFn encode = get_encode();
    if (!encode) return fail(EFFDET_E_CUDA, "effdet_dw_backward: cuTensorMapEncodeTiled unavailable%s", "");
    DwWgParams p;
    memset(&p, 0, sizeof(p));
    const int Ho = (H + S - 1) / S, Wo = (W + S - 1) / S;
    p.C = C; p.partial = partial;
    p.pad_t = max((Ho - 1) * S + K - H, 0) / 2;
    p.pad_l = max((Wo - 1) * S + K - W, 0) / 2;
    p.tiles_x = (Wo + Cfg::TW - 1) / Cfg::TW; p.tiles_y = (Ho + Cfg::TH - 1) / Cfg::TH;
    p.spatial_tiles = p.tiles_x * p.tiles_y * B;
    cuuint32_t es[4] = {1, 1, 1, 1};
    {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
        cuuint32_t box[4] = {(cuuint32_t)Cfg::CB, (cuuint32_t)Cfg::IW, (cuuint32_t)Cfg::IH, 1};
        CUresult r = encode(&p.x_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(x), dims, strides, box,
                            es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(EFFDET_E_CUDA, "effdet_dw_backward: encode(x) failed %s(%lld)", "", (long long)r);
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * Wo, (cuuint64_t)C * 2 * Wo * Ho};
        cuuint32_t box[4] = {(cuuint32_t)Cfg::CB, (cuuint32_t)Cfg::TW, (cuuint32_t)Cfg::TH, 1};
        CUresult r = encode(&p.z_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(dz), dims, strides, box,
                            es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(EFFDET_E_CUDA, "effdet_dw_backward: encode(dz) failed %s(%lld)", "", (long long)r);
    }
    auto kern = dw_wgrad_tma_kernel<K, S, CP>;
    static bool attr = false;
    if (!attr) {
        EFFDET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        attr = true;
    }
    dim3 grid(nsplit, (C + Cfg::CB - 1) / Cfg::CB);
    EFFDET_CUDA(launch_pdl(kern, grid, dim3(Cfg::NT), SMEM, st, p));
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

// ------------------------------------------------------------------ fused BiFPN node
// out = ReLU(BN(DW3x3( fuse(resample(in0), in1 [, in2]) )))    (model.py:154-194 / :226-266, layers.py:26-31)
// resample = nearest x2 upsampling of the coarser level (MODE0 = 1) or 2x2 / stride-2 max pooling of the
// finer level (MODE0 = 2), applied while the tile is read.  Same persistent TMA pipeline as the
// depthwise kernel: per tile the halo patches of in1 / in2, the matching patch of in0 (half or double
// resolution) and the 9 x CB weights land in shared memory (zero fill outside the images: the fused
// tensor is zero there too, which is exactly the SAME padding of the depthwise conv); one cooperative
// pass writes the fused bf16 patch over in1's patch, then every thread runs its 2 x 4 register tile.
// HBM traffic: each input once, the output once (the fused tensor never leaves the SM).
struct alignas(64) NodeParams {
    CUtensorMap in0_map, in1_map, in2_map;  // (C, W*, H*, B) bf16
    CUtensorMap w_map;                      // (C, 9) f32
    const float *fw;                        // fusion weights (2 or 3) or NULL (plain Add)
    const float *scale, *shift;
    __nv_bfloat16 *y;
    float eps;
    int has_in2, H, W, C, tiles_x, tiles_y, cblocks, total_tiles;
};

template <int MODE0, int CP> struct NodeCfg {
    static constexpr int TH = 8, TW = MODE0 == 2 ? 8 : 16;
    static constexpr int CB = CP * 2;
    static constexpr int FH = TH + 2, FW = TW + 2;                       // fused patch (halo 1)
    static constexpr int SH = MODE0 == 1 ? TH / 2 + 2 : 2 * FH;          // in0 patch
    static constexpr int SW = MODE0 == 1 ? TW / 2 + 2 : 2 * FW;
    static constexpr int NRT = (TH / kRtH) * (TW / kRtW);
    static constexpr int PASSES = NRT / 8;
    static constexpr int NT = CP * 8;
    static constexpr int F_BYTES = FH * FW * CB * 2;
    static constexpr int F_PAD = ((F_BYTES + 127) / 128) * 128;
    static constexpr int S_BYTES = SH * SW * CB * 2;
    static constexpr int S_PAD = ((S_BYTES + 127) / 128) * 128;
    static constexpr int W_BYTES = 9 * CB * 4;
    static constexpr int STAGE_BYTES = ((2 * F_PAD + S_PAD + W_BYTES + 127) / 128) * 128;   // in1 | in2 | in0 | w
    static constexpr size_t SMEM = 2 * (size_t)STAGE_BYTES + 64 + 128;
};

template <int MODE0, int CP>
__global__ void __launch_bounds__(NodeCfg<MODE0, CP>::NT, 2)
bifpn_node_tma_kernel(const __grid_constant__ NodeParams p) {
    using Cfg = NodeCfg<MODE0, CP>;
    constexpr int CB = Cfg::CB, TH = Cfg::TH, TW = Cfg::TW, FH = Cfg::FH, FW = Cfg::FW, SW = Cfg::SW;
    extern __shared__ uint8_t dsm_raw[];
    uint8_t *dsm = dsm_raw + ((128u - (smem_u32(dsm_raw) & 127u)) & 127u);
    uint64_t *full = reinterpret_cast<uint64_t *>(dsm + 2 * Cfg::STAGE_BYTES);
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    EFFDET_PDL_SYNC();

    auto issue = [&](int tx, int ty, int cb, int b, int stage) {
        uint8_t *dst = dsm + (size_t)stage * Cfg::STAGE_BYTES;
        const int x0 = tx * TW - 1, y0 = ty * TH - 1;
        mbar_expect_tx(&full[stage], (uint32_t)(Cfg::F_BYTES * (p.has_in2 ? 2 : 1) + Cfg::S_BYTES + Cfg::W_BYTES));
        tma_load_4d(dst, &p.in1_map, &full[stage], cb * CB, x0, y0, b);
        if (p.has_in2) tma_load_4d(dst + Cfg::F_PAD, &p.in2_map, &full[stage], cb * CB, x0, y0, b);
        if (MODE0 == 1) tma_load_4d(dst + 2 * Cfg::F_PAD, &p.in0_map, &full[stage], cb * CB, tx * TW / 2 - 1, ty * TH / 2 - 1, b);
        else tma_load_4d(dst + 2 * Cfg::F_PAD, &p.in0_map, &full[stage], cb * CB, 2 * x0, 2 * y0, b);
        tma_load_2d(dst + 2 * Cfg::F_PAD + Cfg::S_PAD, &p.w_map, &full[stage], cb * CB, 0);
    };

    float w0 = 1.f, w1 = 1.f, w2 = 1.f, rinv = 1.f;
    const bool weighted = p.fw != nullptr;
    if (weighted) {
        w0 = fmaxf(p.fw[0], 0.f); w1 = fmaxf(p.fw[1], 0.f); w2 = p.has_in2 ? fmaxf(p.fw[2], 0.f) : 0.f;
        rinv = 1.f / (w0 + w1 + w2 + p.eps);
    }
    const int pair = tid % CP, slot = tid / CP;
    auto unpack = [](uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); };

    TileWalk nxt;
    nxt.init(blockIdx.x, gridDim.x, p.tiles_x, p.tiles_y, p.cblocks);
    if (tid == 0 && nxt.t < p.total_tiles) issue(nxt.x, nxt.y, nxt.c, nxt.b, 0);
    int it = 0;
    for (; nxt.t < p.total_tiles; ++it) {
        const int stage = it & 1;
        const int tx = nxt.x, ty = nxt.y, cb = nxt.c, b = nxt.b;
        nxt.next(p.tiles_x, p.tiles_y, p.cblocks);
        if (tid == 0 && nxt.t < p.total_tiles) issue(nxt.x, nxt.y, nxt.c, nxt.b, stage ^ 1);
        const int c = cb * CB + pair * 2;
        const bool c_ok = c < p.C;
        float2 sc = make_float2(0.f, 0.f), sh = make_float2(0.f, 0.f);
        if (c_ok) {
            sc = *reinterpret_cast<const float2 *>(p.scale + c);
            sh = *reinterpret_cast<const float2 *>(p.shift + c);
        }
        uint8_t *s1 = dsm + (size_t)stage * Cfg::STAGE_BYTES;
        const uint8_t *s2 = s1 + Cfg::F_PAD, *s0 = s1 + 2 * Cfg::F_PAD;
        const float *sW = reinterpret_cast<const float *>(s1 + 2 * Cfg::F_PAD + Cfg::S_PAD);
        mbar_wait(&full[stage], (it >> 1) & 1);

        // ---- fusion pass: fused patch (bf16) written over in1's patch
        for (int px = slot; px < FH * FW; px += 8) {
            const int hy = px / FW, hx = px - hy * FW;
            const size_t off = ((size_t)px * CB + pair * 2) * 2;
            float2 a;
            if (MODE0 == 1) {
                a = unpack(*reinterpret_cast<const uint32_t *>(
                    s0 + ((size_t)(((hy + 1) >> 1) * SW + ((hx + 1) >> 1)) * CB + pair * 2) * 2));
            } else {
                const uint8_t *q = s0 + ((size_t)((2 * hy) * SW + 2 * hx) * CB + pair * 2) * 2;
                const float2 q0 = unpack(*reinterpret_cast<const uint32_t *>(q));
                const float2 q1 = unpack(*reinterpret_cast<const uint32_t *>(q + CB * 2));
                const float2 q2 = unpack(*reinterpret_cast<const uint32_t *>(q + (size_t)SW * CB * 2));
                const float2 q3 = unpack(*reinterpret_cast<const uint32_t *>(q + (size_t)SW * CB * 2 + CB * 2));
                a = make_float2(fmaxf(fmaxf(q0.x, q1.x), fmaxf(q2.x, q3.x)), fmaxf(fmaxf(q0.y, q1.y), fmaxf(q2.y, q3.y)));
            }
            const float2 bb = unpack(*reinterpret_cast<const uint32_t *>(s1 + off));
            float2 f;
            if (weighted) { f.x = w0 * a.x + w1 * bb.x; f.y = w0 * a.y + w1 * bb.y; }
            else { f.x = a.x + bb.x; f.y = a.y + bb.y; }
            if (p.has_in2) {
                const float2 cc = unpack(*reinterpret_cast<const uint32_t *>(s2 + off));
                if (weighted) { f.x += w2 * cc.x; f.y += w2 * cc.y; }
                else { f.x += cc.x; f.y += cc.y; }
            }
            if (weighted) { f.x *= rinv; f.y *= rinv; }
            *reinterpret_cast<__nv_bfloat162 *>(s1 + off) = __floats2bfloat162_rn(f.x, f.y);
        }
        __syncthreads();

        // ---- depthwise 3x3 + BN + ReLU on the fused patch
        float2 wk[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) wk[i] = *reinterpret_cast<const float2 *>(sW + i * CB + pair * 2);
#pragma unroll 1
        for (int pass = 0; pass < Cfg::PASSES; ++pass) {
            const int rt = pass * 8 + slot;
            const int ry = (rt / (TW / kRtW)) * kRtH, rx = (rt % (TW / kRtW)) * kRtW;
            float2 acc[kRtH][kRtW];
#pragma unroll
            for (int i = 0; i < kRtH; ++i)
#pragma unroll
                for (int j = 0; j < kRtW; ++j) acc[i][j] = make_float2(0.f, 0.f);
            const uint8_t *base = s1 + ((size_t)(ry * FW + rx) * CB + pair * 2) * 2;
#pragma unroll
            for (int rr = 0; rr < kRtH + 2; ++rr) {
                float2 in[kRtW + 2];
#pragma unroll
                for (int j = 0; j < kRtW + 2; ++j)
                    in[j] = unpack(*reinterpret_cast<const uint32_t *>(base + (size_t)(rr * FW + j) * CB * 2));
#pragma unroll
                for (int orow = 0; orow < kRtH; ++orow) {
                    const int ky = rr - orow;
                    if (ky < 0 || ky >= 3) continue;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int oc = 0; oc < kRtW; ++oc)
                            acc[orow][oc] = __ffma2_rn(in[oc + kx], wk[ky * 3 + kx], acc[orow][oc]);
                }
            }
            if (c_ok) {
#pragma unroll
                for (int orow = 0; orow < kRtH; ++orow) {
                    const int oy = ty * TH + ry + orow;
                    if (oy >= p.H) continue;
                    __nv_bfloat16 *yrow = p.y + (((size_t)b * p.H + oy) * p.W) * p.C + c;
#pragma unroll
                    for (int oc = 0; oc < kRtW; ++oc) {
                        const int ox = tx * TW + rx + oc;
                        if (ox < p.W) {
                            const float2 z = __ffma2_rn(acc[orow][oc], sc, sh);
                            *reinterpret_cast<__nv_bfloat162 *>(yrow + (size_t)ox * p.C) =
                                __floats2bfloat162_rn(fmaxf(z.x, 0.f), fmaxf(z.y, 0.f));
                        }
                    }
                }
            }
        }
        __syncthreads();        // everyone is done with this stage before it is refilled
    }
}

template <int MODE0, int CP>
static int launch_node_tma(const void *in0, const void *in1, const void *in2, const float *fw, float eps,
                           const float *dw, const float *scale, const float *shift, void *out, int B, int H, int W,
                           int C, cudaStream_t st) {
    using Cfg = NodeCfg<MODE0, CP>;
    EncodeTiledFn encode = get_encode();
    if (!encode) return fail(EFFDET_E_CUDA, "effdet_bifpn_node: cuTensorMapEncodeTiled unavailable%s", "");
    NodeParams p;
    memset(&p, 0, sizeof(p));
    p.fw = fw; p.eps = eps; p.scale = scale; p.shift = shift; p.y = static_cast<__nv_bfloat16 *>(out);
    p.has_in2 = in2 != nullptr; p.H = H; p.W = W; p.C = C;
    p.tiles_x = (W + Cfg::TW - 1) / Cfg::TW; p.tiles_y = (H + Cfg::TH - 1) / Cfg::TH;
    p.cblocks = (C + Cfg::CB - 1) / Cfg::CB;
    p.total_tiles = p.tiles_x * p.tiles_y * p.cblocks * B;
    cuuint32_t es[4] = {1, 1, 1, 1};
    auto enc4 = [&](CUtensorMap *m, const void *ptr, int w, int h, int bw, int bh) -> CUresult {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * w, (cuuint64_t)C * 2 * w * h};
        cuuint32_t box[4] = {(cuuint32_t)Cfg::CB, (cuuint32_t)bw, (cuuint32_t)bh, 1};
        return encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(ptr), dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    const int H0 = MODE0 == 1 ? H / 2 : H * 2, W0 = MODE0 == 1 ? W / 2 : W * 2;
    CUresult r = enc4(&p.in1_map, in1, W, H, Cfg::FW, Cfg::FH);
    if (r == CUDA_SUCCESS && in2) r = enc4(&p.in2_map, in2, W, H, Cfg::FW, Cfg::FH);
    if (r == CUDA_SUCCESS) r = enc4(&p.in0_map, in0, W0, H0, Cfg::SW, Cfg::SH);
    if (r == CUDA_SUCCESS) {
        cuuint64_t dims[2] = {(cuuint64_t)C, 9};
        cuuint64_t strides[1] = {(cuuint64_t)C * 4};
        cuuint32_t box[2] = {(cuuint32_t)Cfg::CB, 9};
        cuuint32_t es2[2] = {1, 1};
        r = encode(&p.w_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(dw), dims, strides, box, es2,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return fail(EFFDET_E_CUDA, "effdet_bifpn_node: cuTensorMapEncodeTiled failed %s(%lld)", "", (long long)r);
    auto kern = bifpn_node_tma_kernel<MODE0, CP>;
    static int per_sm = 0;
    if (!per_sm) {
        EFFDET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        int nb = 0;
        EFFDET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::NT, Cfg::SMEM));
        per_sm = nb < 1 ? 1 : nb;
    }
    int grid = kNumSMs * per_sm;
    if (grid > p.total_tiles) grid = p.total_tiles;
    EFFDET_CUDA(launch_pdl(kern, dim3(grid), dim3(Cfg::NT), Cfg::SMEM, st, p));
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

// channel pairs per block: the largest of 32 / 24 / 16 (64 / 48 / 32 channels) that divides C,
// else the one wasting the fewest lanes
static int pick_cp(int C) {
    // 32 pairs (64 channels) = one warp per register-tile slot: shared-memory reads are conflict free.
    // 24 / 16 pairs put lanes of two slots into one warp; their pixel offsets are multiples of four
    // pixels = multiples of 32 words, so those reads are 2-way bank conflicted -- only worth it when
    // 64-channel blocks would idle more than ~1/8 of the lanes.
    if (const char *e = getenv("EFFDET_DW_CP")) { const int v = atoi(e); if (v == 32 || v == 24 || v == 16) return v; }
    const int cand[3] = {32, 24, 16};
    const long w64 = (long)((C + 63) / 64) * 64 - C;
    if (w64 * 8 <= C) return 32;
    for (int cp : cand) if (C % (2 * cp) == 0) return cp;
    int best = 32; long waste = -1;
    for (int cp : cand) {
        const long w = (long)((C + 2 * cp - 1) / (2 * cp)) * 2 * cp - C;
        if (waste < 0 || w < waste) { waste = w; best = cp; }
    }
    return best;
}

// SE partials: one per output tile (8 x 16 for stride 1, 8 x 8 for stride 2)
int dwconv_bf16_tma_se_blocks(int H, int W, int stride) {
    const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
    const int tw = stride == 1 ? 16 : 8;
    return ((Wo + tw - 1) / tw) * ((Ho + 7) / 8);
}

// bf16 entry used by effdet_dwconv (dwconv.cu)
int dwconv_bf16_tma(const void *x, const float *w, const float *scale, const float *shift, void *y, float *se_sum,
                    int B, int H, int W, int C, int k, int stride, int act, cudaStream_t st, float *stats) {
    const int cp = pick_cp(C);
#define DWT_CASE(K_, S_)                                                                                   \
    if (k == K_ && stride == S_) {                                                                         \
        if (cp == 32) return launch_dw_tma<K_, S_, 32>(x, w, scale, shift, y, se_sum, stats, B, H, W, C, act, st); \
        if (cp == 24) return launch_dw_tma<K_, S_, 24>(x, w, scale, shift, y, se_sum, stats, B, H, W, C, act, st); \
        return launch_dw_tma<K_, S_, 16>(x, w, scale, shift, y, se_sum, stats, B, H, W, C, act, st);              \
    }
    DWT_CASE(3, 1) DWT_CASE(5, 1) DWT_CASE(3, 2) DWT_CASE(5, 2)
#undef DWT_CASE
    return fail(EFFDET_E_INVALID, "effdet_dwconv: bad kernel / stride%s", "");
}

// number of spatial splits (partial rows) of the bf16 depthwise weight gradient
int dw_wgrad_bf16_splits(int B, int H, int W, int C, int stride) {
    const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
    const int tw = stride == 1 ? 16 : 8;
    const long tiles = (long)((Wo + tw - 1) / tw) * ((Ho + 7) / 8) * B;
    const int cp = pick_cp(C);
    const int cblocks = (C + 2 * cp - 1) / (2 * cp);
    long ns = (2L * kNumSMs + cblocks - 1) / cblocks;
    if (ns < 2) ns = 2;
    if (ns > tiles) ns = tiles < 2 ? 2 : tiles;
    return (int)ns;
}

int dw_wgrad_bf16_tma(const void *x, const void *dz, float *partial, int nsplit, int B, int H, int W, int C, int k,
                      int stride, cudaStream_t st) {
    const int cp = pick_cp(C);
#define DWW_CASE(K_, S_)                                                                                   \
    if (k == K_ && stride == S_) {                                                                         \
        if (cp == 32) return launch_dw_wgrad_tma<K_, S_, 32>(x, dz, partial, nsplit, B, H, W, C, st);       \
        if (cp == 24) return launch_dw_wgrad_tma<K_, S_, 24>(x, dz, partial, nsplit, B, H, W, C, st);       \
        return launch_dw_wgrad_tma<K_, S_, 16>(x, dz, partial, nsplit, B, H, W, C, st);                    \
    }
    DWW_CASE(3, 1) DWW_CASE(5, 1) DWW_CASE(3, 2) DWW_CASE(5, 2)
#undef DWW_CASE
    return fail(EFFDET_E_INVALID, "effdet_dw_backward: bad kernel / stride%s", "");
}

// bf16 entry used by effdet_bifpn_node (dwconv.cu): mode0 1 = upsample in0, 2 = max-pool in0
int bifpn_node_bf16_tma(const void *in0, int mode0, const void *in1, const void *in2, const float *fw, float eps,
                        const float *dw, const float *scale, const float *shift, void *out, int B, int H, int W,
                        int C, cudaStream_t st) {
    if (mode0 == 1) {
        const int cp = pick_cp(C);
        if (cp == 32) return launch_node_tma<1, 32>(in0, in1, in2, fw, eps, dw, scale, shift, out, B, H, W, C, st);
        if (cp == 24) return launch_node_tma<1, 24>(in0, in1, in2, fw, eps, dw, scale, shift, out, B, H, W, C, st);
        return launch_node_tma<1, 16>(in0, in1, in2, fw, eps, dw, scale, shift, out, B, H, W, C, st);
    }
    // max-pool mode: the in0 patch is 4x the tile, 32-channel blocks keep two stages + two blocks per SM
    return launch_node_tma<2, 16>(in0, in1, in2, fw, eps, dw, scale, shift, out, B, H, W, C, st);
}

}  // namespace effdet
