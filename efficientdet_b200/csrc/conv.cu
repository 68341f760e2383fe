// Dense convolutions, SIMT path (fp32 accumulate; activations fp32 or bf16).
//   stem_conv_kernel     : efficientnet.py:413-423  Conv3x3 s2 (3 -> C0) + BN + swish
//   conv_igemm_kernel    : implicit-GEMM 1x1 / 3x3 (stride 1/2, TF SAME padding) with fused
//                          epilogue:  efficientnet.py:228-237 (expand), :289-304 (project + BN
//                          [+ drop-connect scale] + residual, SE gate folded into the A load,
//                          :255-286), model.py:71-90 (BiFPN ConvBlock), model.py:293-309 /
//                          :324-351 (head trunk + final convs, grouped over pyramid levels and
//                          written straight into the concatenated (B,N,4)/(B,N,C) outputs,
//                          model.py:393-398)
//   bn_fold_kernel       : BatchNormalization inference form -> per-channel scale/shift
// This is the accuracy-mode (fp32) path and the fallback for shapes the tcgen05 kernel
// (conv_tc.cu) does not take.
#include "common.cuh"

namespace effdet {

// ------------------------------------------------------------------ BN fold
__global__ void bn_fold_kernel(const float *__restrict__ gamma, const float *__restrict__ beta,
                               const float *__restrict__ mean, const float *__restrict__ var,
                               float eps, float *__restrict__ scale, float *__restrict__ shift,
                               int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = gamma[c] / sqrtf(var[c] + eps);
    scale[c] = s;
    shift[c] = beta[c] - mean[c] * s;
}

// ------------------------------------------------------------------ stem
// one thread = one output pixel x all C0 channels (C0 <= 64); weights broadcast from smem
// TI = float: normalised pixels; TI = uint8_t: raw letterboxed RGB bytes, normalised on the fly through
// lut[c][byte] = ((byte / 255) - mean_c) / std_c (train_tpu.py:135-140, generators/common.py:418-429), computed
// by the host in float32 exactly as the reference does -> bit-identical to feeding the normalised image
template <typename TI> struct StemPix;
template <> struct StemPix<float> {
    static __device__ __forceinline__ float get(const float *p, const float *, int) { return *p; }
};
template <> struct StemPix<uint8_t> {
    static __device__ __forceinline__ float get(const uint8_t *p, const float *lut, int c) { return lut[c * 256 + *p]; }
};

template <typename TI, typename TO, int C0>
__global__ void __launch_bounds__(128)
stem_conv_kernel(const TI *__restrict__ img, const float *__restrict__ lut, const float *__restrict__ w,
                 const float *__restrict__ scale, const float *__restrict__ shift,
                 TO *__restrict__ out, int B, int H, int W, int Ho, int Wo, int pad_t, int pad_l) {
    __shared__ float sw[27 * C0];
    __shared__ float ss[C0], sb[C0];
    __shared__ float slut[sizeof(TI) == 1 ? 768 : 1];
    for (int i = threadIdx.x; i < 27 * C0; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < C0; i += blockDim.x) { ss[i] = scale[i]; sb[i] = shift[i]; }
    if (sizeof(TI) == 1)
        for (int i = threadIdx.x; i < 768; i += blockDim.x) slut[i] = lut[i];
    __syncthreads();
    const size_t total = (size_t)B * Ho * Wo;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    const int ox = (int)(p % Wo), oy = (int)((p / Wo) % Ho), b = (int)(p / ((size_t)Wo * Ho));
    float acc[C0];
#pragma unroll
    for (int c = 0; c < C0; ++c) acc[c] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * 2 - pad_t + ky;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = ox * 2 - pad_l + kx;
            if (ix < 0 || ix >= W) continue;
            const TI *px = img + (((size_t)b * H + iy) * W + ix) * 3;
            const float v0 = StemPix<TI>::get(px, slut, 0), v1 = StemPix<TI>::get(px + 1, slut, 1),
                        v2 = StemPix<TI>::get(px + 2, slut, 2);
            const float *wk = sw + (ky * 3 + kx) * 3 * C0;
#pragma unroll
            for (int c = 0; c < C0; ++c)
                acc[c] += v0 * wk[c] + v1 * wk[C0 + c] + v2 * wk[2 * C0 + c];
        }
    }
    TO *o = out + p * C0;
#pragma unroll
    for (int c = 0; c < C0; ++c) o[c] = from_f<TO>(activate<EFFDET_ACT_SWISH>(acc[c] * ss[c] + sb[c]));
}


// bf16-output stem, tiled: one block = 4 x 32 output pixels of one image.  The fp32 RGB patch
// (9 x 65 pixels) is staged with coalesced row copies, the 27 x C0 weights sit next to it; a thread
// owns 8 output channels x a strip of 4 horizontally adjacent pixels (weights read once per tap
// and reused over the strip, packed FFMA2 math), and writes 16-byte bf16 vectors: the warp's
// stores are whole 64-byte pixel rows.  HBM-bound: 12 B in + 2*C0 B out per output pixel.
constexpr int kStemTW = 32, kStemTH = 4, kStemStrip = 4;
constexpr int kStemIW = 2 * kStemTW + 1, kStemIH = 2 * kStemTH + 1;

template <typename TI, int C0, int ACT, typename TO = __nv_bfloat16>
__global__ void __launch_bounds__((C0 / 8) * (kStemTW / kStemStrip) * kStemTH)
stem_conv_tiled_kernel(const TI *__restrict__ img, const float *__restrict__ lut, const float *__restrict__ w,
                       const float *__restrict__ scale, const float *__restrict__ shift,
                       TO *__restrict__ out, int H, int W, int Ho, int Wo, int pad_t, int pad_l,
                       int tiles_x, uint32_t zero) {
    constexpr int NO = C0 / 8;
    constexpr int NT = NO * (kStemTW / kStemStrip) * kStemTH;
    constexpr int ROW = kStemIW * 3;
    __shared__ __align__(16) float sw[27 * C0];
    __shared__ float sin_[kStemIH * ROW];
    __shared__ float slut[sizeof(TI) == 1 ? 768 : 1];
    if (sizeof(TI) == 1) {
        for (int i = threadIdx.x; i < 768; i += NT) slut[i] = lut[i];
        __syncthreads();
    }
    const int b = blockIdx.y;
    const int ty0 = (blockIdx.x / tiles_x) * kStemTH, tx0 = (blockIdx.x % tiles_x) * kStemTW;
    const int iy0 = ty0 * 2 - pad_t, ix0 = tx0 * 2 - pad_l;
    for (int i = threadIdx.x; i < 27 * C0 / 4; i += NT)
        reinterpret_cast<float4 *>(sw)[i] = reinterpret_cast<const float4 *>(w)[i];
    const TI *ib = img + (size_t)b * H * W * 3;
    // patch staging: four loads per thread in flight (tied together: ptxas otherwise sinks each load next to its
    // shared-memory store and the block waits one global round trip per element, 14 times)
    for (int i0 = threadIdx.x; i0 < kStemIH * ROW; i0 += 4 * NT) {
        uint32_t raw[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * NT;
            const int r = i / ROW, cidx = i - r * ROW;
            const int gy = iy0 + r, g3 = ix0 * 3 + cidx;
            // outside the image: zero in NORMALISED space (TF SAME padding acts on the network input)
            ok[u] = i < kStemIH * ROW && gy >= 0 && gy < H && g3 >= 0 && g3 < W * 3;
            raw[u] = 0u;
            if (ok[u]) {
                const TI *src = ib + (size_t)gy * W * 3 + g3;
                raw[u] = sizeof(TI) == 1 ? (uint32_t)*reinterpret_cast<const uint8_t *>(src)
                                         : __float_as_uint(*reinterpret_cast<const float *>(src));
            }
        }
        const uint32_t t = (raw[0] ^ raw[1] ^ raw[2] ^ raw[3]) & zero;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * NT;
            if (i < kStemIH * ROW) {
                const uint32_t w = raw[u] | t;
                const int cidx = i % ROW;
                sin_[i] = !ok[u] ? 0.f : (sizeof(TI) == 1 ? slut[(cidx % 3) * 256 + w] : __uint_as_float(w));
            }
        }
    }
    __syncthreads();
    const int oct = threadIdx.x % NO, strip = threadIdx.x / NO;
    const int sy = strip / (kStemTW / kStemStrip), sx = (strip % (kStemTW / kStemStrip)) * kStemStrip;
    float2 acc[kStemStrip][4];
#pragma unroll
    for (int o = 0; o < kStemStrip; ++o)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[o][k] = make_float2(0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const float *rowp = sin_ + (sy * 2 + ky) * ROW + sx * 2 * 3;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const float *wp = sw + ((ky * 3 + kx) * 3 + ci) * C0 + oct * 8;
                const float4 wa = *reinterpret_cast<const float4 *>(wp);
                const float4 wb = *reinterpret_cast<const float4 *>(wp + 4);
                const float2 wk[4] = {make_float2(wa.x, wa.y), make_float2(wa.z, wa.w), make_float2(wb.x, wb.y),
                                      make_float2(wb.z, wb.w)};
#pragma unroll
                for (int o = 0; o < kStemStrip; ++o) {
                    const float v = rowp[(o * 2 + kx) * 3 + ci];
                    const float2 vv = make_float2(v, v);
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[o][k] = __ffma2_rn(vv, wk[k], acc[o][k]);
                }
            }
    }
    const int oy = ty0 + sy;
    if (oy >= Ho) return;
    constexpr bool kF32Out = sizeof(TO) == 4;           // fp32 accuracy mode: exact swish, fp32 stores
    const float pre = (ACT == EFFDET_ACT_SWISH && !kF32Out) ? 0.5f : 1.f;
    float2 sc[4], sh[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        sc[k] = *reinterpret_cast<const float2 *>(scale + oct * 8 + 2 * k);
        sh[k] = *reinterpret_cast<const float2 *>(shift + oct * 8 + 2 * k);
        sc[k].x *= pre; sc[k].y *= pre; sh[k].x *= pre; sh[k].y *= pre;
    }
#pragma unroll
    for (int o = 0; o < kStemStrip; ++o) {
        const int ox = tx0 + sx + o;
        if (ox >= Wo) continue;
        if (kF32Out) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 z = __ffma2_rn(acc[o][k], sc[k], sh[k]);
                v[2 * k] = activate<ACT>(z.x); v[2 * k + 1] = activate<ACT>(z.y);
            }
            float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<float *>(out) +
                                                     (((size_t)b * Ho + oy) * Wo + ox) * C0 + oct * 8);
            dst[0] = make_float4(v[0], v[1], v[2], v[3]);
            dst[1] = make_float4(v[4], v[5], v[6], v[7]);
            continue;
        }
        uint4 ov;
        __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&ov);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float2 z = __ffma2_rn(acc[o][k], sc[k], sh[k]);
            if (ACT == EFFDET_ACT_SWISH) {
                float tx, ty;
                asm("tanh.approx.f32 %0, %1;" : "=f"(tx) : "f"(z.x));
                asm("tanh.approx.f32 %0, %1;" : "=f"(ty) : "f"(z.y));
                z.x = fmaf(z.x, tx, z.x); z.y = fmaf(z.y, ty, z.y);
            }
            oh[k] = __floats2bfloat162_rn(z.x, z.y);
        }
        *reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(out) + (((size_t)b * Ho + oy) * Wo + ox) * C0 + oct * 8) = ov;
    }
}

// ------------------------------------------------------------------ implicit GEMM
constexpr int kMaxGroups = 5;
struct ConvGroup {
    const void *x;        // (B,H,W,Cin)
    void *y;              // output base
    const void *res;      // optional residual, same indexing as y
    const void *mask;     // optional ReLU mask (zero the output where mask <= 0), indexed like y
    int H, W, Ho, Wo;
    long long x_batch_stride;   // elements between images in x
    int ldx;              // elements between consecutive input pixels in x
    int tile_begin;       // first M tile of this group
    long long y_batch_stride;   // elements between images in y
    int ldc;              // elements between consecutive output pixels in y
};
struct ConvParams {
    ConvGroup g[kMaxGroups];
    int n_groups;
    int B, Cin, Cout, kh, kw, stride;
    const float *w;       // (kh*kw*Cin, Cout) == Keras HWIO
    const float *scale;   // per-Cout multiplier (folded BN) or null
    const float *shift;   // per-Cout addend (folded BN / bias) or null
    const float *gate;    // (B,Cin) SE gate applied to the input, or null
    const float *keep;    // (B) drop-connect scale applied before the residual add, or null
    int act;
};

constexpr int BM = 64, BN = 64, BK = 16;

template <typename TI> __device__ __forceinline__ void load4(const TI *p, bool vec, int n, float *v);
template <> __device__ __forceinline__ void load4<float>(const float *p, bool vec, int n, float *v) {
    if (vec && n >= 4) {
        float4 t = *reinterpret_cast<const float4 *>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = i < n ? p[i] : 0.f;
    }
}
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16 *p, bool vec,
                                                                   int n, float *v) {
    if (vec && n >= 4) {
        uint2 t = *reinterpret_cast<const uint2 *>(p);
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162 *>(&t.x);
        __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162 *>(&t.y);
        v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = i < n ? __bfloat162float(p[i]) : 0.f;
    }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
conv_igemm_kernel(const ConvParams p) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN];
    const int tid = threadIdx.x;
    int gi = 0;
#pragma unroll
    for (int i = 1; i < kMaxGroups; ++i)
        if (i < p.n_groups && (int)blockIdx.x >= p.g[i].tile_begin) gi = i;
    const ConvGroup G = p.g[gi];
    const int HoWo = G.Ho * G.Wo;
    const int M = p.B * HoWo;
    const int m0 = ((int)blockIdx.x - G.tile_begin) * BM;
    const int n0 = blockIdx.y * BN;
    const int Cin = p.Cin, Cout = p.Cout;
    const int pad_t = max((G.Ho - 1) * p.stride + p.kh - G.H, 0) / 2;   // TF SAME: before = total/2
    const int pad_l = max((G.Wo - 1) * p.stride + p.kw - G.W, 0) / 2;

    // A-load role: one row, 4 consecutive k
    const int a_row = tid >> 2, a_k = (tid & 3) * 4;
    const int am = m0 + a_row;
    const bool a_ok = am < M;
    int ab = 0, aoy = 0, aox = 0;
    if (a_ok) { ab = am / HoWo; int r = am - ab * HoWo; aoy = r / G.Wo; aox = r - aoy * G.Wo; }
    const bool a_vec = (Cin & 3) == 0 && (G.ldx & 3) == 0 && (G.x_batch_stride & 3) == 0;
    // B-load role: one k row, 4 consecutive n
    const int b_k = tid >> 4, b_n = (tid & 15) * 4;
    const bool b_vec = (Cout & 3) == 0;
    // compute role
    const int ty = tid >> 4, tx = tid & 15;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const TI *X = static_cast<const TI *>(G.x);
    const int taps = p.kh * p.kw;
    for (int tap = 0; tap < taps; ++tap) {
        const int ky = tap / p.kw, kx = tap - ky * p.kw;
        const int iy = aoy * p.stride - pad_t + ky, ix = aox * p.stride - pad_l + kx;
        const bool pix_ok = a_ok && iy >= 0 && iy < G.H && ix >= 0 && ix < G.W;
        const TI *src = X + (size_t)ab * G.x_batch_stride +
                        ((size_t)(pix_ok ? iy : 0) * G.W + (pix_ok ? ix : 0)) * G.ldx;
        const float *gsrc = p.gate ? p.gate + (size_t)ab * Cin : nullptr;
        for (int c0 = 0; c0 < Cin; c0 += BK) {
            float av[4] = {0.f, 0.f, 0.f, 0.f};
            const int ck = c0 + a_k;
            if (pix_ok && ck < Cin) {
                load4<TI>(src + ck, a_vec, Cin - ck, av);
                if (gsrc) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) if (ck + i < Cin) av[i] *= gsrc[ck + i];
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) As[a_k + i][a_row] = av[i];
            float bv[4] = {0.f, 0.f, 0.f, 0.f};
            const int kk = c0 + b_k, nn = n0 + b_n;
            if (kk < Cin && nn < Cout)
                load4<float>(p.w + ((size_t)tap * Cin + kk) * Cout + nn, b_vec, Cout - nn, bv);
            *reinterpret_cast<float4 *>(&Bs[b_k][b_n]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                const float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
                const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
                const float ar[4] = {a.x, a.y, a.z, a.w}, br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

    // epilogue
    TO *Y = static_cast<TO *>(G.y);
    const TO *R = static_cast<const TO *>(G.res);
    const TO *MK = static_cast<const TO *>(G.mask);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
        const int b = m / HoWo, pix = m - b * HoWo;
        const size_t base = (size_t)b * G.y_batch_stride + (size_t)pix * G.ldc;
        const float kp = p.keep ? p.keep[b] : 1.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= Cout) continue;
            float v = acc[i][j];
            if (p.scale) v *= p.scale[n];
            if (p.shift) v += p.shift[n];
            v = activate_rt(v, p.act);
            if (MK && !(to_f<TO>(MK[base + n]) > 0.f)) v = 0.f;
            if (R) v = v * kp + to_f<TO>(R[base + n]);
            Y[base + n] = from_f<TO>(v);
        }
    }
}

}  // namespace effdet

using namespace effdet;

extern "C" int effdet_bn_fold(const float *gamma, const float *beta, const float *mean,
                              const float *var, float eps, float *scale, float *shift, int C,
                              void *stream) {
    EFFDET_REQUIRE(gamma && beta && mean && var && scale && shift && C > 0, "bad arguments");
    bn_fold_kernel<<<cdiv(C, 128), 128, 0, as_stream(stream)>>>(gamma, beta, mean, var, eps, scale,
                                                               shift, C);
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

template <typename TI, typename TO>
static int launch_stem(const TI *img, const float *lut, const float *w, const float *scale, const float *shift,
                       void *out, int B, int H, int W, int C0, cudaStream_t st) {
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    const int pad_t = ((Ho - 1) * 2 + 3 - H > 0 ? (Ho - 1) * 2 + 3 - H : 0) / 2;
    const int pad_l = ((Wo - 1) * 2 + 3 - W > 0 ? (Wo - 1) * 2 + 3 - W : 0) / 2;
    const size_t total = (size_t)B * Ho * Wo;
    dim3 grid(cdiv(total, 128));
#define STEM_CASE(C)                                                                             \
    case C:                                                                                      \
        stem_conv_kernel<TI, TO, C><<<grid, 128, 0, st>>>(img, lut, w, scale, shift, static_cast<TO *>(out), \
                                                      B, H, W, Ho, Wo, pad_t, pad_l);           \
        break;
    switch (C0) {
        STEM_CASE(32) STEM_CASE(40) STEM_CASE(48) STEM_CASE(56) STEM_CASE(64) STEM_CASE(8) STEM_CASE(16)
        default:
            return fail(EFFDET_E_UNSUPPORTED, "effdet_stem_conv: %sunsupported C0=%lld", "", C0);
    }
#undef STEM_CASE
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

template <typename TI>
static int launch_stem_bf16(const TI *img, const float *lut, const float *w, const float *scale, const float *shift,
                            void *out, int B, int H, int W, int C0, int act, cudaStream_t st) {
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    const int pad_t = ((Ho - 1) * 2 + 3 - H > 0 ? (Ho - 1) * 2 + 3 - H : 0) / 2;
    const int pad_l = ((Wo - 1) * 2 + 3 - W > 0 ? (Wo - 1) * 2 + 3 - W : 0) / 2;
    const int tx = (Wo + kStemTW - 1) / kStemTW, ty = (Ho + kStemTH - 1) / kStemTH;
    dim3 grid(tx * ty, B);
#define STEM_T(C)                                                                                          \
    case C:                                                                                                \
        if (act == EFFDET_ACT_SWISH)                                                                       \
            stem_conv_tiled_kernel<TI, C, EFFDET_ACT_SWISH><<<grid, (C / 8) * (kStemTW / kStemStrip) * kStemTH, 0, st>>>( \
                img, lut, w, scale, shift, static_cast<__nv_bfloat16 *>(out), H, W, Ho, Wo, pad_t, pad_l, tx, 0u);  \
        else                                                                                               \
            stem_conv_tiled_kernel<TI, C, EFFDET_ACT_NONE><<<grid, (C / 8) * (kStemTW / kStemStrip) * kStemTH, 0, st>>>( \
                img, lut, w, scale, shift, static_cast<__nv_bfloat16 *>(out), H, W, Ho, Wo, pad_t, pad_l, tx, 0u);  \
        break;
    switch (C0) {
        STEM_T(32) STEM_T(40) STEM_T(48) STEM_T(56) STEM_T(64)
        default:
            if (act != EFFDET_ACT_SWISH) return fail(EFFDET_E_UNSUPPORTED, "effdet_stem_conv_act: %sunsupported C0=%lld", "", C0);
            return launch_stem<TI, __nv_bfloat16>(img, lut, w, scale, shift, out, B, H, W, C0, st);
    }
#undef STEM_T
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

// fp32 output on the same tiled kernel (exact swish): the per-pixel fallback ran at 0.8 TB/s (2.0 ms of a D2 / batch-64
// fp32 forward)
template <typename TI>
static int launch_stem_f32_tiled(const TI *img, const float *lut, const float *w, const float *scale, const float *shift,
                                 void *out, int B, int H, int W, int C0, cudaStream_t st) {
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    const int pad_t = ((Ho - 1) * 2 + 3 - H > 0 ? (Ho - 1) * 2 + 3 - H : 0) / 2;
    const int pad_l = ((Wo - 1) * 2 + 3 - W > 0 ? (Wo - 1) * 2 + 3 - W : 0) / 2;
    const int tx = (Wo + kStemTW - 1) / kStemTW, ty = (Ho + kStemTH - 1) / kStemTH;
    dim3 grid(tx * ty, B);
#define STEM_F(C)                                                                                          \
    case C:                                                                                                \
        stem_conv_tiled_kernel<TI, C, EFFDET_ACT_SWISH, float><<<grid, (C / 8) * (kStemTW / kStemStrip) * kStemTH, 0, st>>>( \
            img, lut, w, scale, shift, static_cast<float *>(out), H, W, Ho, Wo, pad_t, pad_l, tx, 0u);     \
        break;
    switch (C0) {
        STEM_F(32) STEM_F(40) STEM_F(48) STEM_F(56) STEM_F(64)
        default: return launch_stem<TI, float>(img, lut, w, scale, shift, out, B, H, W, C0, st);
    }
#undef STEM_F
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

extern "C" int effdet_stem_conv(const float *images, const float *kernel, const float *scale,
                                const float *shift, void *out, int B, int H, int W, int C0,
                                int out_dtype, void *stream) {
    EFFDET_REQUIRE(images && kernel && scale && shift && out, "null pointer");
    EFFDET_REQUIRE(B > 0 && H > 0 && W > 0, "bad sizes");
    if (out_dtype == EFFDET_F32)
        return launch_stem_f32_tiled<float>(images, nullptr, kernel, scale, shift, out, B, H, W, C0, as_stream(stream));
    if (out_dtype == EFFDET_BF16)
        return launch_stem_bf16<float>(images, nullptr, kernel, scale, shift, out, B, H, W, C0, EFFDET_ACT_SWISH,
                                       as_stream(stream));
    return fail(EFFDET_E_INVALID, "effdet_stem_conv: bad dtype%s", "");
}

/* Stem fed with the raw letterboxed uint8 RGB image (what train_tpu.py:170-183 decodes from the TFRecord PNG and
 * what generators/common.py:406-417 builds before dividing by 255): normalize_image (train_tpu.py:135-140 ==
 * generators/common.py:418-429) is applied on the fly through lut (3 x 256 floats, lut[c][v] = ((v/255) - mean_c)
 * / std_c evaluated in float32 by the caller), so the float image (12 B/pixel) is never uploaded or stored.
 * act: EFFDET_ACT_SWISH (folded BN in scale/shift) or, bf16 only, EFFDET_ACT_NONE (raw z for training). */
extern "C" int effdet_stem_conv_u8(const unsigned char *images, const float *lut, const float *kernel,
                                   const float *scale, const float *shift, void *out, int B, int H, int W, int C0,
                                   int act, int out_dtype, void *stream) {
    EFFDET_REQUIRE(images && lut && kernel && scale && shift && out, "null pointer");
    EFFDET_REQUIRE(B > 0 && H > 0 && W > 0, "bad sizes");
    EFFDET_REQUIRE(act == EFFDET_ACT_SWISH || (act == EFFDET_ACT_NONE && out_dtype == EFFDET_BF16),
                   "act must be swish (or none with bf16 output)");
    if (out_dtype == EFFDET_F32)
        return launch_stem_f32_tiled<uint8_t>(images, lut, kernel, scale, shift, out, B, H, W, C0, as_stream(stream));
    if (out_dtype == EFFDET_BF16)
        return launch_stem_bf16<uint8_t>(images, lut, kernel, scale, shift, out, B, H, W, C0, act, as_stream(stream));
    return fail(EFFDET_E_INVALID, "effdet_stem_conv_u8: bad dtype%s", "");
}

/* out[p][c] = lut[c][in[p][c]]: normalize_image on the device for callers that need the float image itself
 * (weight gradient of a trainable stem). */
__global__ void __launch_bounds__(256)
normalize_u8_kernel(const unsigned char *__restrict__ in, const float *__restrict__ lut, float *__restrict__ out,
                    size_t n) {
    __shared__ float slut[768];
    for (int i = threadIdx.x; i < 768; i += 256) slut[i] = lut[i];
    __syncthreads();
    // 12 bytes = 4 pixels per thread: 3 x 4-byte loads, 3 x 16-byte stores
    const size_t ngrp = n / 12;
    for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < ngrp; g += (size_t)gridDim.x * 256) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(in) + g * 3;
        float4 *dst = reinterpret_cast<float4 *>(out) + g * 3;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const uint32_t w = src[q];
            float4 o;
            o.x = slut[((q * 4 + 0) % 3) * 256 + (w & 0xff)];
            o.y = slut[((q * 4 + 1) % 3) * 256 + ((w >> 8) & 0xff)];
            o.z = slut[((q * 4 + 2) % 3) * 256 + ((w >> 16) & 0xff)];
            o.w = slut[((q * 4 + 3) % 3) * 256 + (w >> 24)];
            dst[q] = o;
        }
    }
    if (blockIdx.x == 0)
        for (size_t i = ngrp * 12 + threadIdx.x; i < n; i += 256) out[i] = slut[(i % 3) * 256 + in[i]];
}
extern "C" int effdet_normalize_u8(const unsigned char *images, const float *lut, float *out, size_t n_values,
                                   void *stream) {
    EFFDET_REQUIRE(images && lut && out && n_values > 0 && n_values % 3 == 0, "bad arguments");
    EFFDET_REQUIRE((reinterpret_cast<uintptr_t>(images) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                   "unaligned buffers");
    size_t nb = cdiv(n_values / 12 + 1, 256);
    if (nb > (size_t)148 * 16) nb = (size_t)148 * 16;
    normalize_u8_kernel<<<(unsigned)nb, 256, 0, as_stream(stream)>>>(images, lut, out, n_values);
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

/* bf16 stem with a selectable epilogue: act = EFFDET_ACT_SWISH (inference, folded BN in scale/shift) or
 * EFFDET_ACT_NONE (training: raw convolution output z, scale = 1, shift = 0; BN runs on batch statistics). */
extern "C" int effdet_stem_conv_act(const float *images, const float *kernel, const float *scale,
                                    const float *shift, void *out, int B, int H, int W, int C0, int act,
                                    void *stream) {
    EFFDET_REQUIRE(images && kernel && scale && shift && out, "null pointer");
    EFFDET_REQUIRE(B > 0 && H > 0 && W > 0, "bad sizes");
    EFFDET_REQUIRE(act == EFFDET_ACT_SWISH || act == EFFDET_ACT_NONE, "act must be swish or none");
    return launch_stem_bf16<float>(images, nullptr, kernel, scale, shift, out, B, H, W, C0, act, as_stream(stream));
}

int effdet_conv2d_tc(const effdet_conv_desc *d, void *stream);     // conv_tc.cu

extern "C" int effdet_conv2d(const effdet_conv_desc *d, void *stream) {
    EFFDET_REQUIRE(d, "null descriptor");
    EFFDET_REQUIRE(d->n_groups >= 1 && d->n_groups <= kMaxGroups, "1..5 groups");
    EFFDET_REQUIRE(d->B > 0 && d->Cin > 0 && d->Cout > 0, "bad sizes");
    EFFDET_REQUIRE((d->kh == 1 && d->kw == 1) || (d->kh == 3 && d->kw == 3), "kernel 1x1 or 3x3");
    EFFDET_REQUIRE(d->stride == 1 || d->stride == 2, "stride 1 or 2");
    EFFDET_REQUIRE(d->weight || d->weight_bf16, "null weight");
    if (d->allow_tensor_core && d->weight_bf16) {
        const int rc = effdet_conv2d_tc(d, stream);
        if (rc != EFFDET_E_UNSUPPORTED) return rc;
    }
    EFFDET_REQUIRE(!d->split_planes, "split_planes descriptors run on the tensor-core path only (shape not supported)");
    EFFDET_REQUIRE(d->weight, "shape not supported by the tensor-core path and no fp32 weight given");
    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.n_groups = d->n_groups; p.B = d->B; p.Cin = d->Cin; p.Cout = d->Cout;
    p.kh = d->kh; p.kw = d->kw; p.stride = d->stride;
    p.w = d->weight; p.scale = d->scale; p.shift = d->shift; p.gate = d->gate; p.keep = d->keep;
    p.act = d->act;
    int tiles = 0;
    const size_t in_es = d->in_dtype == EFFDET_BF16 ? 2 : 4;
    for (int i = 0; i < d->n_groups; ++i) {
        ConvGroup &g = p.g[i];
        EFFDET_REQUIRE(d->x[i] && d->y[i] && d->H[i] > 0 && d->W[i] > 0, "bad group");
        g.x = d->x[i]; g.y = d->y[i]; g.res = d->residual[i]; g.mask = d->relu_mask[i];
        g.ldx = d->ldx[i] ? d->ldx[i] : d->Cin;
        g.x_batch_stride = d->x_batch_stride[i] ? d->x_batch_stride[i]
                                                : (long long)d->H[i] * d->W[i] * g.ldx;
        g.H = d->H[i]; g.W = d->W[i];
        g.Ho = (d->H[i] + d->stride - 1) / d->stride;
        g.Wo = (d->W[i] + d->stride - 1) / d->stride;
        g.tile_begin = tiles;
        g.ldc = d->ldc[i] ? d->ldc[i] : d->Cout;
        g.y_batch_stride = d->y_batch_stride[i] ? d->y_batch_stride[i]
                                                : (long long)g.Ho * g.Wo * g.ldc;
        EFFDET_REQUIRE((reinterpret_cast<uintptr_t>(g.x) % (4 * in_es)) == 0,
                       "input must be aligned to 4 elements");
        tiles += (int)cdiv((size_t)d->B * g.Ho * g.Wo, BM);
    }
    EFFDET_REQUIRE((reinterpret_cast<uintptr_t>(d->weight) & 15) == 0, "weights must be 16B aligned");
    dim3 grid(tiles, cdiv(d->Cout, BN));
    cudaStream_t st = as_stream(stream);
    if (d->in_dtype == EFFDET_F32 && d->out_dtype == EFFDET_F32)
        conv_igemm_kernel<float, float><<<grid, 256, 0, st>>>(p);
    else if (d->in_dtype == EFFDET_BF16 && d->out_dtype == EFFDET_BF16)
        conv_igemm_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(p);
    else if (d->in_dtype == EFFDET_BF16 && d->out_dtype == EFFDET_F32)
        conv_igemm_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>(p);
    else if (d->in_dtype == EFFDET_F32 && d->out_dtype == EFFDET_BF16)
        conv_igemm_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>(p);
    else
        return fail(EFFDET_E_UNSUPPORTED, "effdet_conv2d: unsupported dtype combination%s", "");
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}
