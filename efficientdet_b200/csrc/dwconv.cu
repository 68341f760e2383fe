// HBM-bound NHWC kernels of the backbone and BiFPN (fp32 math; fp32 or bf16 storage, 16-byte
// vector accesses along C):
//   dwconv_kernel      efficientnet.py:242-252  depthwise kxk + BN + swish (+ SE squeeze partial
//                      sums, :259-260) ; also model.py:48-68 when used stand-alone
//   se_gate_kernel     efficientnet.py:255-286  the two squeeze-excite FCs -> per-(b,c) gate
//   wbifpn_add_kernel  layers.py:26-31          fast normalised fusion / keras Add
//   bifpn_node_kernel  model.py:154-194, :226-266  upsample|maxpool-on-load + fusion +
//                      DepthwiseConv3x3 + BN + ReLU in one pass (shared-memory halo tile)
#define EFFDET_PDL_TU_LEVEL 2
#include "common.cuh"

namespace effdet {

template <typename T, int CV> struct Vec;
template <> struct Vec<float, 4> {
    static __device__ __forceinline__ void load(const float *p, float *v) {
        float4 t = *reinterpret_cast<const float4 *>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float *p, const float *v) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec<__nv_bfloat16, 8> {
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float *v) {
        uint4 t = *reinterpret_cast<const uint4 *>(p);
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float *v) {
        uint4 t;
        __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4 *>(p) = t;
    }
};
template <int CV> __device__ __forceinline__ void loadf(const float *p, float *v) {
#pragma unroll
    for (int i = 0; i < CV; i += 4) {
        float4 t = *reinterpret_cast<const float4 *>(p + i);
        v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
}

// ------------------------------------------------------------------ depthwise conv
// block = nvec x PY threads (nvec = C/CV channel vectors); a block covers `ppb` output pixels
// of one image.  SE partial sums are reduced in a fixed order (registers -> shared rows ->
// one (image, block, channel) partial in global memory; se_gate_kernel adds the blocks in
// order), so the forward pass is bit-reproducible run to run -- no float atomics.
template <typename T, int CV, int K>
__global__ void dwconv_kernel(const T *__restrict__ x, const float *__restrict__ w,
                              const float *__restrict__ scale, const float *__restrict__ shift,
                              T *__restrict__ y, float *__restrict__ se_sum, int H, int W, int Ho,
                              int Wo, int C, int stride, int pad_t, int pad_l, int ppb, int act) {
    extern __shared__ float s_sum[];
    const int nvec = C / CV;
    const int PY = blockDim.x / nvec;
    const int cv = threadIdx.x % nvec, py = threadIdx.x / nvec;
    const int b = blockIdx.y, c = cv * CV;
    const int p0 = blockIdx.x * ppb, p1 = min(p0 + ppb, Ho * Wo);
    float sc[CV], sh[CV], tot[CV];
    loadf<CV>(scale + c, sc);
    loadf<CV>(shift + c, sh);
#pragma unroll
    for (int i = 0; i < CV; ++i) tot[i] = 0.f;
    const T *xb = x + (size_t)b * H * W * C;
    T *yb = y + (size_t)b * Ho * Wo * C;
    if (py < PY) {
        for (int p = p0 + py; p < p1; p += PY) {
            const int oy = p / Wo, ox = p - oy * Wo;
            float acc[CV];
#pragma unroll
            for (int i = 0; i < CV; ++i) acc[i] = 0.f;
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
                const int iy = oy * stride - pad_t + ky;
                if (iy < 0 || iy >= H) continue;
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    const int ix = ox * stride - pad_l + kx;
                    if (ix < 0 || ix >= W) continue;
                    float v[CV], wk[CV];
                    Vec<T, CV>::load(xb + ((size_t)iy * W + ix) * C + c, v);
                    loadf<CV>(w + (size_t)(ky * K + kx) * C + c, wk);
#pragma unroll
                    for (int i = 0; i < CV; ++i) acc[i] = fmaf(v[i], wk[i], acc[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < CV; ++i) {
                acc[i] = activate_rt(acc[i] * sc[i] + sh[i], act);
                tot[i] += acc[i];
            }
            Vec<T, CV>::store(yb + (size_t)p * C + c, acc);
        }
    }
    if (se_sum) {
#pragma unroll
        for (int i = 0; i < CV; ++i) s_sum[(size_t)py * C + c + i] = tot[i];
        __syncthreads();
        float *dst = se_sum + ((size_t)b * gridDim.x + blockIdx.x) * C;
        for (int i = threadIdx.x; i < C; i += blockDim.x) {
            float t = 0.f;
            for (int r = 0; r < PY; ++r) t += s_sum[(size_t)r * C + i];
            dst[i] = t;
        }
    }
}

// ------------------------------------------------------------------ depthwise conv, tiled
// Shared-memory version: one block = 8 x 16 output pixels x CB channels.  The input tile
// (with its k-1 halo, zero-filled outside the image == TF SAME padding) is staged once with
// 16-byte cp.async copies, the k*k*CB weights sit next to it, and every thread produces a strip
// of 4 horizontally adjacent outputs for one 16-byte channel vector, so an input vector is read
// from shared memory once per kernel ROW and reused across the strip.  Lanes run along the
// channel vectors, so shared and global accesses of a warp are contiguous 128-byte lines.
constexpr int kDwTW = 16, kDwTH = 8, kDwStrip = 4;

__device__ __forceinline__ void cp_async16_zfill(void *dst, const void *src, bool valid) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

// PLANES (fp32 only): the output is written as the bf16 hi | lo split (B, Ho, Wo, 2*C) that the tensor-core
// convolution of the fp32 accuracy mode reads (effdet_split_bf16's layout) instead of the fp32 tensor: same bytes,
// and the separate split pass over the widest tensors of the network (6x expanded) disappears.
template <typename T, int CV, int K, int S, int NCV, int ACT, bool PLANES = false>
__global__ void __launch_bounds__(NCV * (kDwTW / kDwStrip) * kDwTH)
dwconv_tiled_kernel(const T *__restrict__ x, const float *__restrict__ w, const float *__restrict__ scale,
                    const float *__restrict__ shift, T *__restrict__ y, float *__restrict__ se_sum, int H,
                    int W, int Ho, int Wo, int C, int pad_t, int pad_l, int tiles_x) {
    constexpr int CB = NCV * CV;
    constexpr int IH = (kDwTH - 1) * S + K, IW = (kDwTW - 1) * S + K;
    constexpr int NT = NCV * (kDwTW / kDwStrip) * kDwTH;
    constexpr int NIN = (kDwStrip - 1) * S + K;            // input vectors per strip and kernel row
    extern __shared__ __align__(16) uint8_t dsm[];
    T *sIn = reinterpret_cast<T *>(dsm);                                   // [IH][IW][CB]
    float *sW = reinterpret_cast<float *>(dsm + (size_t)IH * IW * CB * sizeof(T));   // [K*K][CB]
    float *sRed = sW + K * K * CB;                                         // [strips][CB] (SE only)

    const int b = blockIdx.z, cb0 = blockIdx.y * CB;
    const int ty0 = (blockIdx.x / tiles_x) * kDwTH, tx0 = (blockIdx.x % tiles_x) * kDwTW;
    const int iy0 = ty0 * S - pad_t, ix0 = tx0 * S - pad_l;
    const T *xb = x + (size_t)b * H * W * C;
    // ---- stage input tile
    for (int i = threadIdx.x; i < IH * IW * NCV; i += NT) {
        const int v = i % NCV, pix = i / NCV;
        const int py = pix / IW, px = pix - py * IW;
        const int gy = iy0 + py, gx = ix0 + px, c = cb0 + v * CV;
        const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W && c < C;
        const T *src = ok ? xb + ((size_t)gy * W + gx) * C + c : xb;
        cp_async16_zfill(sIn + (size_t)pix * CB + v * CV, src, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int i = threadIdx.x; i < K * K * CB; i += NT) {
        const int c = cb0 + i % CB;
        sW[i] = c < C ? w[(size_t)(i / CB) * C + c] : 0.f;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int v = threadIdx.x % NCV, strip = threadIdx.x / NCV;
    const int sy = strip / (kDwTW / kDwStrip), sx = strip % (kDwTW / kDwStrip);
    const int c = cb0 + v * CV;
    float acc[kDwStrip][CV];
#pragma unroll
    for (int o = 0; o < kDwStrip; ++o)
#pragma unroll
        for (int k = 0; k < CV; ++k) acc[o][k] = 0.f;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
        float in[NIN][CV];
        const T *rowp = sIn + ((size_t)(sy * S + ky) * IW + sx * kDwStrip * S) * CB + v * CV;
#pragma unroll
        for (int j = 0; j < NIN; ++j) Vec<T, CV>::load(rowp + (size_t)j * CB, in[j]);
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
            float wk[CV];
            loadf<CV>(sW + (ky * K + kx) * CB + v * CV, wk);
#pragma unroll
            for (int o = 0; o < kDwStrip; ++o)
#pragma unroll
                for (int k = 0; k < CV; ++k) acc[o][k] = fmaf(in[o * S + kx][k], wk[k], acc[o][k]);
        }
    }
    float tot[CV];
#pragma unroll
    for (int k = 0; k < CV; ++k) tot[k] = 0.f;
    if (c < C) {
        float sc[CV], sh[CV];
        loadf<CV>(scale + c, sc);
        loadf<CV>(shift + c, sh);
        const int oy = ty0 + sy;
        T *yb = y + (size_t)b * Ho * Wo * C;
#pragma unroll
        for (int o = 0; o < kDwStrip; ++o) {
            const int ox = tx0 + sx * kDwStrip + o;
            if (oy < Ho && ox < Wo) {
#pragma unroll
                for (int k = 0; k < CV; ++k) {
                    acc[o][k] = activate<ACT>(acc[o][k] * sc[k] + sh[k]);
                    tot[k] += acc[o][k];
                }
                if (PLANES) {
                    static_assert(!PLANES || CV == 4, "split output is written for 4-channel fp32 threads");
                    __nv_bfloat16 *yp = reinterpret_cast<__nv_bfloat16 *>(y) +
                                        ((size_t)b * Ho * Wo + (size_t)oy * Wo + ox) * (size_t)(2 * C) + c;
                    __nv_bfloat16 h[4];
                    float l[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        h[k] = __float2bfloat16_rn(acc[o][k]);
                        l[k] = acc[o][k] - __bfloat162float(h[k]);
                    }
                    __nv_bfloat162 hi[2] = {__halves2bfloat162(h[0], h[1]), __halves2bfloat162(h[2], h[3])};
                    __nv_bfloat162 lo[2] = {__floats2bfloat162_rn(l[0], l[1]), __floats2bfloat162_rn(l[2], l[3])};
                    *reinterpret_cast<uint2 *>(yp) = *reinterpret_cast<uint2 *>(hi);
                    *reinterpret_cast<uint2 *>(yp + C) = *reinterpret_cast<uint2 *>(lo);
                } else {
                    Vec<T, CV>::store(yb + ((size_t)oy * Wo + ox) * C + c, acc[o]);
                }
            }
        }
    }
    if (se_sum) {
        // deterministic: strip partials -> shared rows -> fixed-order sum -> (image, tile, channel)
#pragma unroll
        for (int k = 0; k < CV; ++k) sRed[(size_t)strip * CB + v * CV + k] = tot[k];
        __syncthreads();
        constexpr int NS = (kDwTW / kDwStrip) * kDwTH;
        float *dst = se_sum + ((size_t)b * gridDim.x + blockIdx.x) * C;
        for (int i = threadIdx.x; i < CB; i += NT) {
            if (cb0 + i < C) {
                float t = 0.f;
                for (int r = 0; r < NS; ++r) t += sRed[(size_t)r * CB + i];
                dst[cb0 + i] = t;
            }
        }
    }
}

template <typename T, int CV, int K, int S, int NCV, bool PLANES = false>
static int launch_dw_tiled(const void *x, const float *w, const float *scale, const float *shift, void *y,
                           float *se_sum, int B, int H, int W, int C, int act, cudaStream_t st) {
    constexpr int CB = NCV * CV;
    constexpr int IH = (kDwTH - 1) * S + K, IW = (kDwTW - 1) * S + K;
    constexpr int NT = NCV * (kDwTW / kDwStrip) * kDwTH;
    const int Ho = (H + S - 1) / S, Wo = (W + S - 1) / S;
    const int pt = max((Ho - 1) * S + K - H, 0) / 2, pl = max((Wo - 1) * S + K - W, 0) / 2;
    const int tx = (Wo + kDwTW - 1) / kDwTW, ty = (Ho + kDwTH - 1) / kDwTH;
    const size_t smem = (size_t)IH * IW * CB * sizeof(T) + (size_t)K * K * CB * 4 +
                        (size_t)(kDwTW / kDwStrip) * kDwTH * CB * 4;
    dim3 grid(tx * ty, (C + CB - 1) / CB, B);
#define DW_LAUNCH(A)                                                                                       \
    {                                                                                                      \
        auto kern = dwconv_tiled_kernel<T, CV, K, S, NCV, A, PLANES>;                                      \
        if (smem > 48 * 1024) EFFDET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, NT, smem, st>>>(static_cast<const T *>(x), w, scale, shift, static_cast<T *>(y), se_sum, H, W, \
                                     Ho, Wo, C, pt, pl, tx);                                               \
    }
    if (act == EFFDET_ACT_SWISH) DW_LAUNCH(EFFDET_ACT_SWISH)
    else if (act == EFFDET_ACT_RELU) DW_LAUNCH(EFFDET_ACT_RELU)
    else if (act == EFFDET_ACT_NONE) DW_LAUNCH(EFFDET_ACT_NONE)
    else return fail(EFFDET_E_UNSUPPORTED, "effdet_dwconv: unsupported activation%s", "");
#undef DW_LAUNCH
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

// ------------------------------------------------------------------ squeeze-excite FCs
// One block per image.  The two FCs are tiny (C x R, R x C) but every step is a round trip to
// L2 / HBM, so the kernel is organised to keep many wide loads in flight per thread and few
// dependent rounds: 16-byte (8-byte when R % 4 != 0) weight loads, 4-8 independent accumulator
// chains, fixed summation orders (bit-reproducible).
constexpr int kSeThreads = 512;

template <int VW, int NTH = kSeThreads>
__device__ __forceinline__ void se_fc1(const float *__restrict__ mean, const float *__restrict__ w1, float *part,
                                       int C, int R, int tid) {
    const int JV = R / VW, G = NTH / JV;
    const int jv = tid % JV, cg = tid / JV;
    if (cg >= G) return;
    float s[4][VW];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < VW; ++k) s[u][k] = 0.f;
    int c = cg;
    for (; c + 3 * G < C; c += 4 * G) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float m = mean[c + u * G];
            const float *wp = w1 + (size_t)(c + u * G) * R + jv * VW;
            if (VW == 4) {
                const float4 q = *reinterpret_cast<const float4 *>(wp);
                s[u][0] = fmaf(m, q.x, s[u][0]); s[u][1 % VW] = fmaf(m, q.y, s[u][1 % VW]);
                s[u][2 % VW] = fmaf(m, q.z, s[u][2 % VW]); s[u][3 % VW] = fmaf(m, q.w, s[u][3 % VW]);
            } else if (VW == 2) {
                const float2 q = *reinterpret_cast<const float2 *>(wp);
                s[u][0] = fmaf(m, q.x, s[u][0]); s[u][1 % VW] = fmaf(m, q.y, s[u][1 % VW]);
            } else {
                s[u][0] = fmaf(m, wp[0], s[u][0]);
            }
        }
    }
    for (int u = 0; c < C; c += G, ++u) {
        const float m = mean[c];
#pragma unroll
        for (int k = 0; k < VW; ++k) s[u][k] = fmaf(m, w1[(size_t)c * R + jv * VW + k], s[u][k]);
    }
#pragma unroll
    for (int k = 0; k < VW; ++k) part[cg * R + jv * VW + k] = (s[0][k] + s[1][k]) + (s[2][k] + s[3][k]);
}

__global__ void __launch_bounds__(kSeThreads)
se_gate_kernel(const float *__restrict__ se_sum, int se_blocks, float inv_hw,
               const float *__restrict__ w1, const float *__restrict__ b1,
               const float *__restrict__ w2, const float *__restrict__ b2,
               float *__restrict__ gate, int C, int R) {
    EFFDET_PDL_SYNC();
    extern __shared__ float sm[];      // mean[C] | r[R] | part[4 * kSeThreads]
    float *mean = sm, *r = sm + C, *part = sm + C + R;
    const int b = blockIdx.x, tid = threadIdx.x;
    if (C >= kSeThreads || se_blocks < 8) {
        for (int c = tid; c < C; c += kSeThreads) {
            const float *src = se_sum + (size_t)b * se_blocks * C + c;
            float t = 0.f;
            for (int k = 0; k < se_blocks; ++k) t += src[(size_t)k * C];
            mean[c] = t * inv_hw;
        }
    } else {
        // few channels, many tile partials (early stages): KS threads per channel each add every
        // KS-th partial, then the KS slices are added in order -- still a fixed summation order
        const int KS = kSeThreads / C;
        const int c = tid % C, slice = tid / C;
        if (slice < KS) {
            const float *src = se_sum + (size_t)b * se_blocks * C + c;
            float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
            int k = slice;
            for (; k + 3 * KS < se_blocks; k += 4 * KS) {
                t0 += src[(size_t)k * C]; t1 += src[(size_t)(k + KS) * C];
                t2 += src[(size_t)(k + 2 * KS) * C]; t3 += src[(size_t)(k + 3 * KS) * C];
            }
            for (; k < se_blocks; k += KS) t0 += src[(size_t)k * C];
            part[slice * C + c] = (t0 + t1) + (t2 + t3);
        }
        __syncthreads();
        if (tid < C) {
            float t = 0.f;
            for (int s2 = 0; s2 < KS; ++s2) t += part[s2 * C + tid];
            mean[tid] = t * inv_hw;
        }
    }
    __syncthreads();
    // FC1 + swish: lanes along the R outputs (w1 is (C, R) row-major), thread groups split C
    const bool v4 = (R % 4 == 0) && ((reinterpret_cast<uintptr_t>(w1) & 15) == 0);
    const bool v2 = (R % 2 == 0) && ((reinterpret_cast<uintptr_t>(w1) & 7) == 0);
    int G;
    if (v4) { se_fc1<4>(mean, w1, part, C, R, tid); G = kSeThreads / (R / 4); }
    else if (v2) { se_fc1<2>(mean, w1, part, C, R, tid); G = kSeThreads / (R / 2); }
    else { se_fc1<1>(mean, w1, part, C, R, tid); G = kSeThreads / R; }
    __syncthreads();
    if (tid < R) {
        float s = 0.f;
        for (int g = 0; g < G; ++g) s += part[g * R + tid];
        r[tid] = activate<EFFDET_ACT_SWISH>(s + b1[tid]);
    }
    __syncthreads();
    // FC2 + sigmoid: a thread owns 4 consecutive channels (w2 is (R, C) row-major: 16-byte loads)
    for (int c = tid * 4; c < C; c += kSeThreads * 4) {
        float s[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < 4; ++k) s[u][k] = 0.f;
        int j = 0;
        for (; j + 3 < R; j += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 q = *reinterpret_cast<const float4 *>(w2 + (size_t)(j + u) * C + c);
                const float rj = r[j + u];
                s[u][0] = fmaf(rj, q.x, s[u][0]); s[u][1] = fmaf(rj, q.y, s[u][1]);
                s[u][2] = fmaf(rj, q.z, s[u][2]); s[u][3] = fmaf(rj, q.w, s[u][3]);
            }
        }
        for (; j < R; ++j) {
            const float4 q = *reinterpret_cast<const float4 *>(w2 + (size_t)j * C + c);
            const float rj = r[j];
            s[0][0] = fmaf(rj, q.x, s[0][0]); s[0][1] = fmaf(rj, q.y, s[0][1]);
            s[0][2] = fmaf(rj, q.z, s[0][2]); s[0][3] = fmaf(rj, q.w, s[0][3]);
        }
        const float4 bb = *reinterpret_cast<const float4 *>(b2 + c);
        float4 o;
        o.x = activate<EFFDET_ACT_SIGMOID>(bb.x + ((s[0][0] + s[1][0]) + (s[2][0] + s[3][0])));
        o.y = activate<EFFDET_ACT_SIGMOID>(bb.y + ((s[0][1] + s[1][1]) + (s[2][1] + s[3][1])));
        o.z = activate<EFFDET_ACT_SIGMOID>(bb.z + ((s[0][2] + s[1][2]) + (s[2][2] + s[3][2])));
        o.w = activate<EFFDET_ACT_SIGMOID>(bb.w + ((s[0][3] + s[1][3]) + (s[2][3] + s[3][3])));
        *reinterpret_cast<float4 *>(gate + (size_t)b * C + c) = o;
    }
}

// Thread-block-cluster version: the NC CTAs of a cluster share ONE image.  Each CTA owns a slice of the C
// channels: it reduces the squeeze partials of its slice, forms the partial FC1 products of that slice for all
// R hidden units, the cluster exchanges the NC partial vectors through distributed shared memory (summed in
// rank order: deterministic), and each CTA finishes FC2 + sigmoid for its own slice.  One image then streams
// its 2*C*R weights through NC SMs with NC-times shorter dependent-load chains -- the single-CTA kernel above
// is a chain of ~25 L2 round trips per image on one SM (11-35 us per block of the network, 18 % of the
// batch-1 inference step).
constexpr int kSeCT = 256;
template <int NC>
__global__ void __launch_bounds__(kSeCT)
se_gate_cluster_kernel(const float *__restrict__ se_sum, int se_blocks, float inv_hw,
                       const float *__restrict__ w1, const float *__restrict__ b1,
                       const float *__restrict__ w2, const float *__restrict__ b2,
                       float *__restrict__ gate, int C, int R, int Cs, float *__restrict__ mean_o,
                       float *__restrict__ s1_o, float *__restrict__ rr_o) {
    // mean_o / s1_o / rr_o (optional): squeeze mean (B,C), FC1 pre-activation and its swish (B,R) -- the
    // forward quantities the SE backward needs; gate == nullptr skips FC2 (backward recompute mode)
    EFFDET_PDL_SYNC();
    extern __shared__ __align__(16) float smc[];     // mean[Cs] | rpart[Rp] | r[Rp] | part[4 * kSeCT + Cs]
    const int Rp = (R + 3) & ~3;                     // keeps `part` 16-byte aligned
    float *mean = smc, *rpart = smc + Cs, *r = rpart + Rp, *part = r + Rp;
    unsigned rank;
    asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int b = blockIdx.x / NC, tid = threadIdx.x;
    const int c0 = (int)rank * Cs;
    int n = C - c0; n = n > Cs ? Cs : (n < 0 ? 0 : n);
    // ---- squeeze: mean of this CTA's channel slice (KS thread groups per channel, fixed order)
    if (n > 0) {
        const int KS = n >= kSeCT ? 1 : kSeCT / n;
        const float *src0 = se_sum + (size_t)b * se_blocks * C + c0;
        for (int idx = tid; idx < n * KS; idx += kSeCT) {
            const int c = idx % n, sl = idx / n;
            const float *src = src0 + c;
            float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
            int k = sl;
            // eight partial rows in flight (a wide slice, n > 256 channels, walks all se_blocks rows with one
            // thread per channel: 64 rows at four loads in flight were 16 dependent L2 round trips, most of the
            // 40 us this kernel took at C = 2688 in the r3 launch list)
            for (; k + 7 * KS < se_blocks; k += 8 * KS) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = src[(size_t)(k + u * KS) * C];
                t0 += v[0]; t1 += v[1]; t2 += v[2]; t3 += v[3];
                t0 += v[4]; t1 += v[5]; t2 += v[6]; t3 += v[7];
            }
            for (; k + 3 * KS < se_blocks; k += 4 * KS) {
                t0 += src[(size_t)k * C]; t1 += src[(size_t)(k + KS) * C];
                t2 += src[(size_t)(k + 2 * KS) * C]; t3 += src[(size_t)(k + 3 * KS) * C];
            }
            for (; k < se_blocks; k += KS) t0 += src[(size_t)k * C];
            part[sl * n + c] = (t0 + t1) + (t2 + t3);
        }
        __syncthreads();
        for (int c = tid; c < n; c += kSeCT) {
            float t = 0.f;
            for (int s2 = 0; s2 < KS; ++s2) t += part[s2 * n + c];
            mean[c] = t * inv_hw;
            if (mean_o) mean_o[(size_t)b * C + c0 + c] = t * inv_hw;
        }
    }
    __syncthreads();
    // ---- FC1, partial over the slice
    const float *w1s = w1 + (size_t)c0 * R;
    const bool v4 = (R % 4 == 0) && ((reinterpret_cast<uintptr_t>(w1s) & 15) == 0) && R / 4 <= kSeCT;
    const bool v2 = (R % 2 == 0) && ((reinterpret_cast<uintptr_t>(w1s) & 7) == 0) && R / 2 <= kSeCT;
    if (n > 0 && R <= kSeCT) {
        int G;
        if (v4) { se_fc1<4, kSeCT>(mean, w1s, part, n, R, tid); G = kSeCT / (R / 4); }
        else if (v2) { se_fc1<2, kSeCT>(mean, w1s, part, n, R, tid); G = kSeCT / (R / 2); }
        else { se_fc1<1, kSeCT>(mean, w1s, part, n, R, tid); G = kSeCT / R; }
        __syncthreads();
        if (tid < R) {
            float t = 0.f;
            for (int g = 0; g < G; ++g) t += part[g * R + tid];
            rpart[tid] = t;
        }
    } else {
        for (int j = tid; j < R; j += kSeCT) {          // R > 256 (not reached by EfficientNet) or empty slice
            float t = 0.f;
            for (int c = 0; c < n; ++c) t = fmaf(mean[c], w1s[(size_t)c * R + j], t);
            rpart[j] = t;
        }
    }
    // ---- exchange: every CTA adds the NC partial vectors in rank order (DSMEM reads)
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    for (int j = tid; j < R; j += kSeCT) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            uint32_t ra;
            asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra)
                : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(rpart + j))), "r"(k));
            float v;
            asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra));
            t += v;
        }
        const float pre = t + b1[j];
        r[j] = activate<EFFDET_ACT_SWISH>(pre);
        if (s1_o && rank == 0) { s1_o[(size_t)b * R + j] = pre; rr_o[(size_t)b * R + j] = r[j]; }
    }
    // nobody may leave (or reuse rpart) while a peer still reads it; also orders r[] for this CTA
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (n <= 0 || !gate) return;
    // ---- FC2 + sigmoid on the slice: thread = (channel quad, r-group); quads along lanes (16-byte loads)
    const int q = n / 4;
    int RG = kSeCT / q; if (RG < 1) RG = 1; if (RG > R) RG = R;
    const float *w2s = w2 + c0;
    for (int idx = tid; idx < q * RG; idx += kSeCT) {
        const int quad = idx % q, rg = idx / q;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
        int j = rg;
        for (; j + 3 * RG < R; j += 4 * RG) {
            const float4 q0 = *reinterpret_cast<const float4 *>(w2s + (size_t)j * C + quad * 4);
            const float4 q1 = *reinterpret_cast<const float4 *>(w2s + (size_t)(j + RG) * C + quad * 4);
            const float4 q2 = *reinterpret_cast<const float4 *>(w2s + (size_t)(j + 2 * RG) * C + quad * 4);
            const float4 q3 = *reinterpret_cast<const float4 *>(w2s + (size_t)(j + 3 * RG) * C + quad * 4);
            const float r0 = r[j], r1 = r[j + RG], r2 = r[j + 2 * RG], r3 = r[j + 3 * RG];
            a0.x = fmaf(r0, q0.x, a0.x); a0.y = fmaf(r0, q0.y, a0.y); a0.z = fmaf(r0, q0.z, a0.z); a0.w = fmaf(r0, q0.w, a0.w);
            a1.x = fmaf(r1, q1.x, a1.x); a1.y = fmaf(r1, q1.y, a1.y); a1.z = fmaf(r1, q1.z, a1.z); a1.w = fmaf(r1, q1.w, a1.w);
            a2.x = fmaf(r2, q2.x, a2.x); a2.y = fmaf(r2, q2.y, a2.y); a2.z = fmaf(r2, q2.z, a2.z); a2.w = fmaf(r2, q2.w, a2.w);
            a3.x = fmaf(r3, q3.x, a3.x); a3.y = fmaf(r3, q3.y, a3.y); a3.z = fmaf(r3, q3.z, a3.z); a3.w = fmaf(r3, q3.w, a3.w);
        }
        for (; j < R; j += RG) {
            const float4 q0 = *reinterpret_cast<const float4 *>(w2s + (size_t)j * C + quad * 4);
            const float r0 = r[j];
            a0.x = fmaf(r0, q0.x, a0.x); a0.y = fmaf(r0, q0.y, a0.y); a0.z = fmaf(r0, q0.z, a0.z); a0.w = fmaf(r0, q0.w, a0.w);
        }
        float4 o;
        o.x = (a0.x + a1.x) + (a2.x + a3.x); o.y = (a0.y + a1.y) + (a2.y + a3.y);
        o.z = (a0.z + a1.z) + (a2.z + a3.z); o.w = (a0.w + a1.w) + (a2.w + a3.w);
        *reinterpret_cast<float4 *>(part + (size_t)rg * n + quad * 4) = o;
    }
    __syncthreads();
    for (int quad = tid; quad < q; quad += kSeCT) {
        float4 t = *reinterpret_cast<const float4 *>(b2 + c0 + quad * 4);
        for (int rg = 0; rg < RG; ++rg) {
            const float4 v = *reinterpret_cast<const float4 *>(part + (size_t)rg * n + quad * 4);
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        float4 o;
        o.x = activate<EFFDET_ACT_SIGMOID>(t.x); o.y = activate<EFFDET_ACT_SIGMOID>(t.y);
        o.z = activate<EFFDET_ACT_SIGMOID>(t.z); o.w = activate<EFFDET_ACT_SIGMOID>(t.w);
        *reinterpret_cast<float4 *>(gate + (size_t)b * C + c0 + quad * 4) = o;
    }
}

template <int NC>
static cudaError_t launch_se_cluster(cudaStream_t st, const float *se_sum, int se_blocks, float inv_hw, const float *w1,
                                     const float *b1, const float *w2, const float *b2, float *gate, int B, int C,
                                     int R, float *mean_o = nullptr, float *s1_o = nullptr, float *rr_o = nullptr) {
    int Cs = (C + NC - 1) / NC;
    Cs = (Cs + 3) / 4 * 4;
    const size_t smem = (size_t)(2 * Cs + 2 * ((R + 3) & ~3) + 4 * kSeCT) * sizeof(float);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(B * NC); cfg.blockDim = dim3(kSeCT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    static const bool pdl = pdl_level() >= EFFDET_PDL_TU_LEVEL;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, se_gate_cluster_kernel<NC>, se_sum, se_blocks, inv_hw, w1, b1, w2, b2, gate, C, R, Cs,
                              mean_o, s1_o, rr_o);
}
// clusters of 8 / 4 / 2 CTAs per image (fewer as the batch alone fills the GPU)
cudaError_t se_cluster_launch(cudaStream_t st, const float *se_sum, int se_blocks, float inv_hw, const float *w1,
                              const float *b1, const float *w2, const float *b2, float *gate, int B, int C, int R,
                              float *mean_o, float *s1_o, float *rr_o) {
    if (B <= 18) return launch_se_cluster<8>(st, se_sum, se_blocks, inv_hw, w1, b1, w2, b2, gate, B, C, R, mean_o, s1_o, rr_o);
    if (B <= 74) return launch_se_cluster<4>(st, se_sum, se_blocks, inv_hw, w1, b1, w2, b2, gate, B, C, R, mean_o, s1_o, rr_o);
    return launch_se_cluster<2>(st, se_sum, se_blocks, inv_hw, w1, b1, w2, b2, gate, B, C, R, mean_o, s1_o, rr_o);
}
bool se_cluster_ok(int C, int R) {
    static const bool use_cluster = !(getenv("EFFDET_SE_CLUSTER") && atoi(getenv("EFFDET_SE_CLUSTER")) == 0);
    return use_cluster && C >= 32 && C % 4 == 0 && R <= kSeCT;
}

// ------------------------------------------------------------------ fusion (stand-alone layer)
struct FuseIn { const void *p[3]; };
template <typename T, int CV>
__global__ void __launch_bounds__(256)
wbifpn_add_kernel(FuseIn in, int n, const float *__restrict__ w, float eps, T *__restrict__ out,
                  size_t nvec) {
    float w0 = 1.f, w1 = 1.f, w2 = 1.f, inv = 1.f;
    if (w) {
        w0 = fmaxf(w[0], 0.f); w1 = fmaxf(w[1], 0.f); w2 = n > 2 ? fmaxf(w[2], 0.f) : 0.f;
        inv = w0 + w1 + w2 + eps;
    }
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        float a[CV], b[CV], c[CV], o[CV];
        Vec<T, CV>::load(static_cast<const T *>(in.p[0]) + i * CV, a);
        Vec<T, CV>::load(static_cast<const T *>(in.p[1]) + i * CV, b);
        if (n > 2) Vec<T, CV>::load(static_cast<const T *>(in.p[2]) + i * CV, c);
#pragma unroll
        for (int k = 0; k < CV; ++k) {
            if (w) {
                float s = w0 * a[k] + w1 * b[k];
                if (n > 2) s += w2 * c[k];
                o[k] = s / inv;
            } else {
                float s = a[k] + b[k];
                if (n > 2) s += c[k];
                o[k] = s;
            }
        }
        Vec<T, CV>::store(out + i * CV, o);
    }
}

// ------------------------------------------------------------------ fused BiFPN node
constexpr int kTile = 8;           // 8x8 output pixels per block
constexpr int kHalo = kTile + 2;
constexpr int kCB = 32;            // channels per block

template <typename T, int CV>
__global__ void __launch_bounds__(256)
bifpn_node_kernel(const T *__restrict__ in0, int mode0, const T *__restrict__ in1,
                  const T *__restrict__ in2, const float *__restrict__ fw, float eps,
                  const float *__restrict__ dw, const float *__restrict__ scale,
                  const float *__restrict__ shift, T *__restrict__ out, int H, int W, int C,
                  int tiles_x) {
    __shared__ __align__(16) float tile[kHalo * kHalo * kCB];
    constexpr int NV = kCB / CV;
    const int b = blockIdx.z, cb0 = blockIdx.y * kCB;
    const int ty0 = (blockIdx.x / tiles_x) * kTile, tx0 = (blockIdx.x % tiles_x) * kTile;
    float w0 = 1.f, w1 = 1.f, w2 = 1.f, inv = 1.f;
    const bool weighted = fw != nullptr;
    if (weighted) {
        w0 = fmaxf(fw[0], 0.f); w1 = fmaxf(fw[1], 0.f); w2 = in2 ? fmaxf(fw[2], 0.f) : 0.f;
        inv = w0 + w1 + w2 + eps;
    }
    const int H0 = mode0 == 1 ? H / 2 : (mode0 == 2 ? H * 2 : H);
    const int W0 = mode0 == 1 ? W / 2 : (mode0 == 2 ? W * 2 : W);
    const T *p0 = in0 + (size_t)b * H0 * W0 * C;
    const T *p1 = in1 + (size_t)b * H * W * C;
    const T *p2 = in2 ? in2 + (size_t)b * H * W * C : nullptr;

    for (int item = threadIdx.x; item < kHalo * kHalo * NV; item += 256) {
        const int v = item % NV, hp = item / NV;
        const int hy = hp / kHalo, hx = hp - hy * kHalo;
        const int yy = ty0 + hy - 1, xx = tx0 + hx - 1, c = cb0 + v * CV;
        float f[CV];
#pragma unroll
        for (int k = 0; k < CV; ++k) f[k] = 0.f;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W && c < C) {
            float a[CV], bb[CV];
            if (mode0 == 1) {
                Vec<T, CV>::load(p0 + ((size_t)(yy >> 1) * W0 + (xx >> 1)) * C + c, a);
            } else if (mode0 == 2) {
                const T *q = p0 + ((size_t)(2 * yy) * W0 + 2 * xx) * C + c;
                float t[CV];
                Vec<T, CV>::load(q, a);
                Vec<T, CV>::load(q + C, t);
#pragma unroll
                for (int k = 0; k < CV; ++k) a[k] = fmaxf(a[k], t[k]);
                Vec<T, CV>::load(q + (size_t)W0 * C, t);
#pragma unroll
                for (int k = 0; k < CV; ++k) a[k] = fmaxf(a[k], t[k]);
                Vec<T, CV>::load(q + (size_t)W0 * C + C, t);
#pragma unroll
                for (int k = 0; k < CV; ++k) a[k] = fmaxf(a[k], t[k]);
            } else {
                Vec<T, CV>::load(p0 + ((size_t)yy * W0 + xx) * C + c, a);
            }
            Vec<T, CV>::load(p1 + ((size_t)yy * W + xx) * C + c, bb);
            if (weighted) {
#pragma unroll
                for (int k = 0; k < CV; ++k) f[k] = w0 * a[k] + w1 * bb[k];
            } else {
#pragma unroll
                for (int k = 0; k < CV; ++k) f[k] = a[k] + bb[k];
            }
            if (p2) {
                float cc[CV];
                Vec<T, CV>::load(p2 + ((size_t)yy * W + xx) * C + c, cc);
#pragma unroll
                for (int k = 0; k < CV; ++k) f[k] += weighted ? w2 * cc[k] : cc[k];
            }
            if (weighted) {
#pragma unroll
                for (int k = 0; k < CV; ++k) f[k] = f[k] / inv;
            }
        }
        float *dst = tile + (size_t)hp * kCB + v * CV;
#pragma unroll
        for (int k = 0; k < CV; k += 4)
            *reinterpret_cast<float4 *>(dst + k) = make_float4(f[k], f[k + 1], f[k + 2], f[k + 3]);
    }
    __syncthreads();
    T *ob = out + (size_t)b * H * W * C;
    for (int item = threadIdx.x; item < kTile * kTile * NV; item += 256) {
        const int v = item % NV, pp = item / NV;
        const int py = pp / kTile, px = pp - py * kTile;
        const int yy = ty0 + py, xx = tx0 + px, c = cb0 + v * CV;
        if (yy >= H || xx >= W || c >= C) continue;
        float acc[CV];
#pragma unroll
        for (int k = 0; k < CV; ++k) acc[k] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float *src = tile + (size_t)((py + ky) * kHalo + px + kx) * kCB + v * CV;
                float wk[CV];
                loadf<CV>(dw + (size_t)(ky * 3 + kx) * C + c, wk);
#pragma unroll
                for (int k = 0; k < CV; ++k) acc[k] = fmaf(src[k], wk[k], acc[k]);
            }
        float sc[CV], sh[CV];
        loadf<CV>(scale + c, sc);
        loadf<CV>(shift + c, sh);
#pragma unroll
        for (int k = 0; k < CV; ++k) acc[k] = fmaxf(acc[k] * sc[k] + sh[k], 0.f);
        Vec<T, CV>::store(ob + ((size_t)yy * W + xx) * C + c, acc);
    }
}

static bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int dw_default_blocks(int B, int HW, int PY) {
    // enough blocks to fill the machine, each thread row handling up to 8 pixels
    int ppb = PY * 8;
    while (ppb > PY && (size_t)B * cdiv(HW, ppb) < (size_t)kNumSMs * 4) ppb >>= 1;
    return (int)cdiv(HW, ppb);
}

template <typename T, int CV>
static int launch_dw(const void *x, const float *w, const float *scale, const float *shift, void *y,
                     float *se_sum, int se_blocks, int B, int H, int W, int C, int k, int stride,
                     int act, cudaStream_t st) {
    const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
    const int pt = max((Ho - 1) * stride + k - H, 0) / 2, pl = max((Wo - 1) * stride + k - W, 0) / 2;
    const int nvec = C / CV;
    if (nvec > 1024) return fail(EFFDET_E_UNSUPPORTED, "effdet_dwconv: %sC=%lld too large", "", C);
    int PY = 256 / nvec; if (PY < 1) PY = 1;
    const int threads = nvec * PY;
    const int HW = Ho * Wo;
    int nblk = se_sum ? se_blocks : dw_default_blocks(B, HW, PY);
    if (nblk < 1 || nblk > HW) return fail(EFFDET_E_INVALID, "effdet_dwconv: bad se_blocks%s", "");
    const int ppb = (int)cdiv(HW, nblk);
    if ((int)cdiv(HW, ppb) != nblk)
        return fail(EFFDET_E_INVALID, "effdet_dwconv: se_blocks %smust come from effdet_dwconv_se_blocks", "");
    dim3 grid(nblk, B);
    const size_t sm = se_sum ? (size_t)PY * C * sizeof(float) : 0;
    if (k == 3)
        dwconv_kernel<T, CV, 3><<<grid, threads, sm, st>>>(
            static_cast<const T *>(x), w, scale, shift, static_cast<T *>(y), se_sum, H, W, Ho, Wo, C,
            stride, pt, pl, ppb, act);
    else
        dwconv_kernel<T, CV, 5><<<grid, threads, sm, st>>>(
            static_cast<const T *>(x), w, scale, shift, static_cast<T *>(y), se_sum, H, W, Ho, Wo, C,
            stride, pt, pl, ppb, act);
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

// dwconv_tma.cu: TMA-fed persistent kernel (bf16)
int dwconv_bf16_tma(const void *x, const float *w, const float *scale, const float *shift, void *y, float *se_sum,
                    int B, int H, int W, int C, int k, int stride, int act, cudaStream_t st, float *stats = nullptr);
int dwconv_bf16_tma_se_blocks(int H, int W, int stride);
int bifpn_node_bf16_tma(const void *in0, int mode0, const void *in1, const void *in2, const float *fw, float eps,
                        const float *dw, const float *scale, const float *shift, void *out, int B, int H, int W,
                        int C, cudaStream_t st);

}  // namespace effdet

using namespace effdet;

extern "C" int effdet_dwconv_se_blocks(int B, int H, int W, int C, int stride, int dtype) {
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || stride < 1) return 0;
    if (dtype == EFFDET_BF16) return dwconv_bf16_tma_se_blocks(H, W, stride);
    {   // tiled kernel: one partial per 8x16 output tile
        const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
        return ((Wo + kDwTW - 1) / kDwTW) * ((Ho + kDwTH - 1) / kDwTH);
    }
    const int CV = dtype == EFFDET_BF16 ? 8 : 4;
    const int nvec = C / CV;
    int PY = 256 / (nvec > 0 ? nvec : 1); if (PY < 1) PY = 1;
    const int HW = ((H + stride - 1) / stride) * ((W + stride - 1) / stride);
    int nblk = dw_default_blocks(B, HW, PY);
    const int ppb = (int)cdiv(HW, nblk);
    return (int)cdiv(HW, ppb);
}

extern "C" int effdet_dwconv(const void *x, const float *kernel, const float *scale,
                             const float *shift, void *y, float *se_sum, int se_blocks, int B,
                             int H, int W, int C, int k, int stride, int act, int dtype,
                             void *stream) {
    EFFDET_REQUIRE(x && kernel && scale && shift && y, "null pointer");
    EFFDET_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, "bad sizes");
    EFFDET_REQUIRE(C % 8 == 0, "C must be a multiple of 8");
    EFFDET_REQUIRE(k == 3 || k == 5, "kernel size 3 or 5");
    EFFDET_REQUIRE(stride == 1 || stride == 2, "stride 1 or 2");
    EFFDET_REQUIRE(al16(x) && al16(y) && al16(kernel) && al16(scale) && al16(shift), "16B alignment");
    if (se_sum)
        EFFDET_REQUIRE(se_blocks == effdet_dwconv_se_blocks(B, H, W, C, stride, dtype),
                       "se_blocks must come from effdet_dwconv_se_blocks");
    cudaStream_t st = as_stream(stream);
#define DW_CASE(T, CV, K, S, NCV) return launch_dw_tiled<T, CV, K, S, NCV>(x, kernel, scale, shift, y, se_sum, B, H, W, C, act, st)
    if (dtype == EFFDET_BF16) {
        return dwconv_bf16_tma(x, kernel, scale, shift, y, se_sum, B, H, W, C, k, stride, act, st);
    } else if (dtype == EFFDET_F32) {
        if (k == 3 && stride == 1) DW_CASE(float, 4, 3, 1, 8);
        if (k == 5 && stride == 1) DW_CASE(float, 4, 5, 1, 8);
        if (k == 3 && stride == 2) DW_CASE(float, 4, 3, 2, 4);
        if (k == 5 && stride == 2) DW_CASE(float, 4, 5, 2, 4);
    }
#undef DW_CASE
    return fail(EFFDET_E_INVALID, "effdet_dwconv: bad dtype%s", "");
}

/* fp32 depthwise convolution whose output is the bf16 hi | lo split (B, Ho, Wo, 2*C) -- see effdet_split_bf16 */
extern "C" int effdet_dwconv_split_out(const float *x, const float *kernel, const float *scale, const float *shift,
                                       void *y_planes, float *se_sum, int se_blocks, int B, int H, int W, int C,
                                       int k, int stride, int act, void *stream) {
    EFFDET_REQUIRE(x && kernel && scale && shift && y_planes, "null pointer");
    EFFDET_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad sizes");
    EFFDET_REQUIRE((k == 3 || k == 5) && (stride == 1 || stride == 2), "kernel 3 or 5, stride 1 or 2");
    EFFDET_REQUIRE(al16(x) && al16(y_planes) && al16(kernel) && al16(scale) && al16(shift), "16B alignment");
    if (se_sum)
        EFFDET_REQUIRE(se_blocks == effdet_dwconv_se_blocks(B, H, W, C, stride, EFFDET_F32),
                       "se_blocks must come from effdet_dwconv_se_blocks");
    cudaStream_t st = as_stream(stream);
#define DWP_CASE(K, S, NCV) return launch_dw_tiled<float, 4, K, S, NCV, true>(x, kernel, scale, shift, y_planes, se_sum, B, H, W, C, act, st)
    if (k == 3 && stride == 1) DWP_CASE(3, 1, 8);
    if (k == 5 && stride == 1) DWP_CASE(5, 1, 8);
    if (k == 3 && stride == 2) DWP_CASE(3, 2, 4);
    DWP_CASE(5, 2, 4);
#undef DWP_CASE
}

extern "C" int effdet_se_gate(const float *se_sum, int se_blocks, float inv_hw, const float *w1,
                              const float *b1, const float *w2, const float *b2, float *gate,
                              int B, int C, int R, void *stream) {
    EFFDET_REQUIRE(se_sum && w1 && b1 && w2 && b2 && gate, "null pointer");
    EFFDET_REQUIRE(B > 0 && C > 0 && R > 0 && se_blocks > 0, "bad sizes");
    EFFDET_REQUIRE(C % 4 == 0 && R <= kSeThreads, "C must be a multiple of 4, R <= 512");
    EFFDET_REQUIRE(((reinterpret_cast<uintptr_t>(w2) | reinterpret_cast<uintptr_t>(b2) |
                     reinterpret_cast<uintptr_t>(gate)) & 15) == 0, "w2 / b2 / gate must be 16B aligned");
    // clusters of 8 / 4 / 2 CTAs per image (fewer as the batch alone fills the GPU); EFFDET_SE_CLUSTER=0 keeps the
    // one-CTA-per-image kernel
    if (se_cluster_ok(C, R)) {
        cudaError_t ce = se_cluster_launch(as_stream(stream), se_sum, se_blocks, inv_hw, w1, b1, w2, b2, gate, B, C, R,
                                           nullptr, nullptr, nullptr);
        EFFDET_CUDA(ce);
        EFFDET_LAUNCHED();
        return EFFDET_OK;
    }
    const size_t sm = (size_t)(C + R + 4 * kSeThreads) * sizeof(float);
    EFFDET_REQUIRE(sm <= 48 * 1024, "C + R too large");
    EFFDET_CUDA(launch_pdl(se_gate_kernel, dim3(B), dim3(kSeThreads), sm, as_stream(stream), se_sum, se_blocks, inv_hw, w1, b1, w2, b2, gate,
                                                      C, R));
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

extern "C" int effdet_wbifpn_add(const void *const *inputs, int n, const float *w, float eps,
                                 void *out, size_t count, int dtype, void *stream) {
    EFFDET_REQUIRE(inputs && out, "null pointer");
    EFFDET_REQUIRE(n == 2 || n == 3, "2 or 3 inputs");
    if (count == 0) return EFFDET_OK;
    FuseIn in;
    for (int i = 0; i < 3; ++i) {
        in.p[i] = i < n ? inputs[i] : nullptr;
        EFFDET_REQUIRE(i >= n || (in.p[i] && al16(in.p[i])), "inputs must be non-null, 16B aligned");
    }
    EFFDET_REQUIRE(al16(out), "output must be 16B aligned");
    cudaStream_t st = as_stream(stream);
    if (dtype == EFFDET_F32) {
        EFFDET_REQUIRE(count % 4 == 0, "element count must be a multiple of 4");
        size_t nv = count / 4;
        unsigned blocks = cdiv(nv, 256); if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
        wbifpn_add_kernel<float, 4><<<blocks, 256, 0, st>>>(in, n, w, eps, static_cast<float *>(out), nv);
    } else if (dtype == EFFDET_BF16) {
        EFFDET_REQUIRE(count % 8 == 0, "element count must be a multiple of 8");
        size_t nv = count / 8;
        unsigned blocks = cdiv(nv, 256); if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
        wbifpn_add_kernel<__nv_bfloat16, 8><<<blocks, 256, 0, st>>>(
            in, n, w, eps, static_cast<__nv_bfloat16 *>(out), nv);
    } else {
        return fail(EFFDET_E_INVALID, "effdet_wbifpn_add: bad dtype%s", "");
    }
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

extern "C" int effdet_bifpn_node(const void *in0, int mode0, const void *in1, const void *in2,
                                 const float *w, float eps, const float *dw_kernel,
                                 const float *scale, const float *shift, void *out, int B, int H,
                                 int W, int C, int dtype, void *stream) {
    EFFDET_REQUIRE(in0 && in1 && dw_kernel && scale && shift && out, "null pointer");
    EFFDET_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad sizes (C % 8 == 0)");
    EFFDET_REQUIRE(mode0 >= 0 && mode0 <= 2, "mode0 in {0,1,2}");
    EFFDET_REQUIRE(mode0 != 1 || (H % 2 == 0 && W % 2 == 0), "upsample target must be even");
    EFFDET_REQUIRE(al16(in0) && al16(in1) && (!in2 || al16(in2)) && al16(out) && al16(dw_kernel) &&
                       al16(scale) && al16(shift), "16B alignment");
    const int tx = (W + kTile - 1) / kTile, ty = (H + kTile - 1) / kTile;
    dim3 grid(tx * ty, (C + kCB - 1) / kCB, B);
    cudaStream_t st = as_stream(stream);
    if (dtype == EFFDET_F32)
        bifpn_node_kernel<float, 4><<<grid, 256, 0, st>>>(
            static_cast<const float *>(in0), mode0, static_cast<const float *>(in1),
            static_cast<const float *>(in2), w, eps, dw_kernel, scale, shift,
            static_cast<float *>(out), H, W, C, tx);
    else if (dtype == EFFDET_BF16 && (mode0 == 1 || mode0 == 2))
        return bifpn_node_bf16_tma(in0, mode0, in1, in2, w, eps, dw_kernel, scale, shift, out, B, H, W, C, st);
    else if (dtype == EFFDET_BF16)
        bifpn_node_kernel<__nv_bfloat16, 8><<<grid, 256, 0, st>>>(
            static_cast<const __nv_bfloat16 *>(in0), mode0, static_cast<const __nv_bfloat16 *>(in1),
            static_cast<const __nv_bfloat16 *>(in2), w, eps, dw_kernel, scale, shift,
            static_cast<__nv_bfloat16 *>(out), H, W, C, tx);
    else
        return fail(EFFDET_E_INVALID, "effdet_bifpn_node: bad dtype%s", "");
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}
