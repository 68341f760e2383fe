// Dense-convolution gradients (SIMT path; fp32 accumulate).
//   conv_wgrad_kernel         dW[tap][ci][co] = sum_pixels X[pix(+tap)][ci] * dZ[pix][co]; split-K over
//                             pixel ranges of up to 5 groups (pyramid levels sharing the weights,
//                             model.py:389-398), partials reduced in a fixed order
//   conv_dgrad_strided_kernel data gradient of a stride-2 3x3 SAME convolution (BiFPN layer-0
//                             P6/P7 laterals, model.py:105-110/208-211); stride-1 data gradients
//                             run through effdet_conv2d with effdet_conv_weight_transpose'd weights
// (backward of SURVEY section 8(a) rows 3, 7, 8 -- the cuDNN bwd-filter / bwd-data calls TF makes)
#include "common.cuh"

namespace effdet {

constexpr int kWgMaxGroups = 5;
struct WgGroup {
    const void *x;   // (B,H,W,Cin) dense
    const void *dz;  // gradient of the conv output, possibly strided
    int H, W, Ho, Wo;
    long long dz_batch_stride;
    int dz_ld;
    int split_begin;     // first z-slice of this group
    int rows_per_split;  // output pixels (over the whole batch) per slice
};
struct WgParams {
    WgGroup g[kWgMaxGroups];
    int n_groups, B, Cin, Cout, kh, kw, stride;
    float *partial;      // [n_splits][taps*Cin*Cout]
};

constexpr int WM = 64, WN = 64, WK = 16;

template <typename TX, typename TZ>
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(const WgParams p) {
    __shared__ float As[WK][WM + 4];   // X rows (pixels) x ci
    __shared__ float Bs[WK][WN];       // dZ rows (pixels) x co
    const int tid = threadIdx.x;
    const int ci_tiles = (p.Cin + WM - 1) / WM;
    const int tap = blockIdx.x / ci_tiles, ci0 = (blockIdx.x % ci_tiles) * WM;
    const int co0 = blockIdx.y * WN;
    int gi = 0;
#pragma unroll
    for (int i = 1; i < kWgMaxGroups; ++i)
        if (i < p.n_groups && (int)blockIdx.z >= p.g[i].split_begin) gi = i;
    const WgGroup G = p.g[gi];
    const int HoWo = G.Ho * G.Wo, M = p.B * HoWo;
    const int r0 = ((int)blockIdx.z - G.split_begin) * G.rows_per_split;
    const int r1 = min(r0 + G.rows_per_split, M);
    const int ky = tap / p.kw, kx = tap - ky * p.kw;
    const int pad_t = max((G.Ho - 1) * p.stride + p.kh - G.H, 0) / 2;
    const int pad_l = max((G.Wo - 1) * p.stride + p.kw - G.W, 0) / 2;
    const TX *X = static_cast<const TX *>(G.x);
    const TZ *DZ = static_cast<const TZ *>(G.dz);
    // load roles: 16 rows x 64 cols per tile, 4 elements per thread
    const int l_row = tid >> 4, l_col = (tid & 15) * 4;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int rb = r0; rb < r1; rb += WK) {
        const int m = rb + l_row;
        float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (m < r1) {
            const int b = m / HoWo, r = m - b * HoWo, oy = r / G.Wo, ox = r - oy * G.Wo;
            const int iy = oy * p.stride - pad_t + ky, ix = ox * p.stride - pad_l + kx;
            if (iy >= 0 && iy < G.H && ix >= 0 && ix < G.W) {
                const TX *src = X + (((size_t)b * G.H + iy) * G.W + ix) * p.Cin + ci0 + l_col;
#pragma unroll
                for (int i = 0; i < 4; ++i) if (ci0 + l_col + i < p.Cin) av[i] = to_f<TX>(src[i]);
            }
            const TZ *zs = DZ + (size_t)b * G.dz_batch_stride + (size_t)r * G.dz_ld + co0 + l_col;
#pragma unroll
            for (int i = 0; i < 4; ++i) if (co0 + l_col + i < p.Cout) bv[i] = to_f<TZ>(zs[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { As[l_row][l_col + i] = av[i]; Bs[l_row][l_col + i] = bv[i]; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < WK; ++k) {
            float ar[4], br[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { ar[i] = As[k][ty * 4 + i]; br[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *out = p.partial + (size_t)blockIdx.z * p.kh * p.kw * p.Cin * p.Cout;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ci = ci0 + ty * 4 + i;
        if (ci >= p.Cin) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co0 + tx * 4 + j;
            if (co < p.Cout) out[((size_t)tap * p.Cin + ci) * p.Cout + co] = acc[i][j];
        }
    }
}

__global__ void wgrad_reduce_kernel(const float *__restrict__ partial, int nsplit, size_t n,
                                    float *__restrict__ out, int accumulate) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float t = 0.f;
    for (int s = 0; s < nsplit; ++s) t += partial[(size_t)s * n + i];
    out[i] = accumulate ? out[i] + t : t;
}

// dx[b,iy,ix,ci] = sum_{tap,co : iy = oy*s - pad + ky} dz[b,oy,ox,co] * W[tap][ci][co]
template <typename T>
__global__ void __launch_bounds__(128)
conv_dgrad_strided_kernel(const T *__restrict__ dz, const float *__restrict__ w, T *__restrict__ dx,
                          int accumulate, int B, int H, int W, int Ho, int Wo, int Cin, int Cout,
                          int k, int stride, int pad_t, int pad_l) {
    const size_t total = (size_t)B * H * W * Cin;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ci = (int)(i % Cin);
    const size_t pix = i / Cin;
    const int ix = (int)(pix % W), iy = (int)((pix / W) % H), b = (int)(pix / ((size_t)W * H));
    float acc = 0.f;
    for (int ky = 0; ky < k; ++ky) {
        const int ny = iy + pad_t - ky;
        if (ny < 0 || ny % stride) continue;
        const int oy = ny / stride;
        if (oy >= Ho) continue;
        for (int kx = 0; kx < k; ++kx) {
            const int nx = ix + pad_l - kx;
            if (nx < 0 || nx % stride) continue;
            const int ox = nx / stride;
            if (ox >= Wo) continue;
            const T *g = dz + (((size_t)b * Ho + oy) * Wo + ox) * Cout;
            const float *ww = w + ((size_t)(ky * k + kx) * Cin + ci) * Cout;
            for (int co = 0; co < Cout; ++co) acc = fmaf(to_f<T>(g[co]), ww[co], acc);
        }
    }
    dx[i] = from_f<T>(accumulate ? to_f<T>(dx[i]) + acc : acc);
}

}  // namespace effdet

using namespace effdet;

extern "C" int effdet_conv_wgrad_splits(const effdet_wgrad_desc *d) {
    if (!d || d->n_groups < 1 || d->n_groups > kWgMaxGroups) return 0;
    // aim for ~4 waves of blocks in total
    const int taps = d->kh * d->kw;
    const int tiles = taps * (int)cdiv(d->Cin, WM) * (int)cdiv(d->Cout, WN);
    size_t rows_total = 0;
    for (int i = 0; i < d->n_groups; ++i)
        rows_total += (size_t)d->B * ((d->H[i] + d->stride - 1) / d->stride) *
                      ((d->W[i] + d->stride - 1) / d->stride);
    int want = (kNumSMs * 4 + tiles - 1) / tiles;
    if (want < 1) want = 1;
    size_t rps = (rows_total + want - 1) / want;
    if (rps < 256) rps = 256;
    rps = (rps + WK - 1) / WK * WK;
    int n = 0;
    for (int i = 0; i < d->n_groups; ++i) {
        size_t rows = (size_t)d->B * ((d->H[i] + d->stride - 1) / d->stride) *
                      ((d->W[i] + d->stride - 1) / d->stride);
        n += (int)cdiv(rows, rps);
    }
    return n;
}

extern "C" int effdet_conv_wgrad(const effdet_wgrad_desc *d, void *stream) {
    EFFDET_REQUIRE(d, "null descriptor");
    EFFDET_REQUIRE(d->n_groups >= 1 && d->n_groups <= kWgMaxGroups, "1..5 groups");
    EFFDET_REQUIRE(d->B > 0 && d->Cin > 0 && d->Cout > 0 && d->dweight && d->partial, "bad arguments");
    EFFDET_REQUIRE((d->kh == 1 && d->kw == 1) || (d->kh == 3 && d->kw == 3), "kernel 1x1 or 3x3");
    const int nsplit = effdet_conv_wgrad_splits(d);
    EFFDET_REQUIRE(nsplit > 0 && nsplit == d->n_splits, "n_splits must be effdet_conv_wgrad_splits()");
    WgParams p;
    memset(&p, 0, sizeof(p));
    p.n_groups = d->n_groups; p.B = d->B; p.Cin = d->Cin; p.Cout = d->Cout;
    p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.partial = d->partial;
    const int taps = d->kh * d->kw;
    const int tiles = taps * (int)cdiv(d->Cin, WM) * (int)cdiv(d->Cout, WN);
    size_t rows_total = 0;
    for (int i = 0; i < d->n_groups; ++i)
        rows_total += (size_t)d->B * ((d->H[i] + d->stride - 1) / d->stride) *
                      ((d->W[i] + d->stride - 1) / d->stride);
    int want = (kNumSMs * 4 + tiles - 1) / tiles; if (want < 1) want = 1;
    size_t rps = (rows_total + want - 1) / want; if (rps < 256) rps = 256;
    rps = (rps + WK - 1) / WK * WK;
    int z = 0;
    for (int i = 0; i < d->n_groups; ++i) {
        WgGroup &g = p.g[i];
        EFFDET_REQUIRE(d->x[i] && d->dz[i], "null group pointer");
        g.x = d->x[i]; g.dz = d->dz[i]; g.H = d->H[i]; g.W = d->W[i];
        g.Ho = (g.H + d->stride - 1) / d->stride; g.Wo = (g.W + d->stride - 1) / d->stride;
        g.dz_ld = d->dz_ld[i] ? d->dz_ld[i] : d->Cout;
        g.dz_batch_stride = d->dz_batch_stride[i] ? d->dz_batch_stride[i]
                                                  : (long long)g.Ho * g.Wo * g.dz_ld;
        g.split_begin = z; g.rows_per_split = (int)rps;
        z += (int)cdiv((size_t)d->B * g.Ho * g.Wo, rps);
    }
    dim3 grid(taps * (int)cdiv(d->Cin, WM), cdiv(d->Cout, WN), z);
    cudaStream_t st = as_stream(stream);
    if (d->x_dtype == EFFDET_F32 && d->dz_dtype == EFFDET_F32)
        conv_wgrad_kernel<float, float><<<grid, 256, 0, st>>>(p);
    else if (d->x_dtype == EFFDET_BF16 && d->dz_dtype == EFFDET_BF16)
        conv_wgrad_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(p);
    else if (d->x_dtype == EFFDET_BF16 && d->dz_dtype == EFFDET_F32)
        conv_wgrad_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>(p);
    else
        return fail(EFFDET_E_UNSUPPORTED, "effdet_conv_wgrad: unsupported dtype combination%s", "");
    EFFDET_LAUNCHED();
    const size_t n = (size_t)taps * d->Cin * d->Cout;
    wgrad_reduce_kernel<<<cdiv(n, 256), 256, 0, st>>>(d->partial, z, n, d->dweight, d->accumulate);
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

extern "C" int effdet_conv_dgrad_strided(const void *dz, const float *weight, void *dx, int accumulate,
                                         int B, int H, int W, int Cin, int Cout, int k, int stride,
                                         int dtype, void *stream) {
    EFFDET_REQUIRE(dz && weight && dx && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "bad arguments");
    EFFDET_REQUIRE((k == 1 || k == 3) && (stride == 1 || stride == 2), "k in {1,3}, stride in {1,2}");
    const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
    const int pt = max((Ho - 1) * stride + k - H, 0) / 2, pl = max((Wo - 1) * stride + k - W, 0) / 2;
    const size_t total = (size_t)B * H * W * Cin;
    cudaStream_t st = as_stream(stream);
    if (dtype == EFFDET_F32)
        conv_dgrad_strided_kernel<float><<<cdiv(total, 128), 128, 0, st>>>(
            (const float *)dz, weight, (float *)dx, accumulate, B, H, W, Ho, Wo, Cin, Cout, k, stride, pt, pl);
    else if (dtype == EFFDET_BF16)
        conv_dgrad_strided_kernel<__nv_bfloat16><<<cdiv(total, 128), 128, 0, st>>>(
            (const __nv_bfloat16 *)dz, weight, (__nv_bfloat16 *)dx, accumulate, B, H, W, Ho, Wo, Cin, Cout,
            k, stride, pt, pl);
    else
        return fail(EFFDET_E_INVALID, "effdet_conv_dgrad_strided: bad dtype%s", "");
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}
