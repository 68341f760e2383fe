// Training losses, forward + backward in one pass each (HBM-bound: one read of the
// predictions/targets, one write of the gradient).
//   focal_kernel      utils/tpu.py:84-155  tpu_focal(alpha, gamma) incl. keras
//                     binary_crossentropy (clip 1e-7 -> logit -> sigmoid-CE), ignore mask,
//                     normalised by max(1, #positive anchors); gradient is emitted w.r.t. the
//                     LOGITS of the class head (sigmoid backward folded in, model.py:351)
//   smooth_l1_kernel  utils/tpu.py:26-81   tpu_smooth_l1 (delta = 1)
// Loss sums use per-block partials reduced in a fixed order (deterministic).
#include "common.cuh"

namespace effdet {

__device__ __forceinline__ float block_sum_256(float v, float *sh) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x == 0)
        for (int i = 0; i < 8; ++i) t += sh[i];
    return t;      // valid in thread 0
}

// counts anchors with state == 1: per-block integer partials (exact, order-independent), then one
// small block finishes.  cnt[2*blk] = regression-state positives, cnt[2*blk+1] = class-state.
__global__ void __launch_bounds__(256)
count_pos_kernel(const float *__restrict__ reg_t, const float *__restrict__ labels_t,
                 const int8_t *__restrict__ state, size_t rows, int C, unsigned *__restrict__ cnt) {
    __shared__ unsigned sh[2][8];
    unsigned n_reg = 0, n_cls = 0;
    for (size_t r = (size_t)blockIdx.x * 256 + threadIdx.x; r < rows; r += (size_t)gridDim.x * 256) {
        n_reg += reg_t[r * 5 + 4] == 1.f;
        float s = labels_t ? labels_t[r * (size_t)(C + 1) + C] : (float)state[r];
        n_cls += s == 1.f;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        n_reg += __shfl_xor_sync(0xffffffffu, n_reg, d);
        n_cls += __shfl_xor_sync(0xffffffffu, n_cls, d);
    }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = n_reg; sh[1][threadIdx.x >> 5] = n_cls; }
    __syncthreads();
    if (threadIdx.x < 2) {
        unsigned t = 0;
        for (int i = 0; i < 8; ++i) t += sh[threadIdx.x][i];
        cnt[2 * blockIdx.x + threadIdx.x] = t;
    }
}
__global__ void __launch_bounds__(256)
count_pos_finalize_kernel(const unsigned *__restrict__ cnt, int nblk, float *__restrict__ out) {
    __shared__ unsigned sh[2][8];
    unsigned a = 0, b = 0;
    for (int i = threadIdx.x; i < nblk; i += 256) { a += cnt[2 * i]; b += cnt[2 * i + 1]; }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, d);
        b += __shfl_xor_sync(0xffffffffu, b, d);
    }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x < 2) {
        unsigned t = 0;
        for (int i = 0; i < 8; ++i) t += sh[threadIdx.x][i];
        out[2 + threadIdx.x] = (float)t;                        // [2] = #pos (regression), [3] = #pos (class)
        out[4 + threadIdx.x] = 1.f / fmaxf(1.f, (float)t);      // [4], [5] = 1/normaliser
    }
}

// optional second copy of the gradients: bf16, one dense channel-padded (B, cells, Cpad) buffer per
// pyramid level (channel = anchor*per + k) -- the layout the TMA-fed tensor-core data / weight
// gradient kernels consume (the concatenated fp32 tensors have rows of 36 / 9C elements, which
// TMA cannot stride over).  Padding channels are never written (the caller zeroes them once).
struct LevelGrads {
    __nv_bfloat16 *cls[5];
    __nv_bfloat16 *reg[5];
    int cells[5];
    int n_levels, cpad_cls, cpad_reg;
    unsigned N;
};
// (image, level, cell, anchor) of concatenated row r
struct LevelPos { int l; size_t cell_index; unsigned a; };
__device__ __forceinline__ LevelPos level_pos(const LevelGrads &L, unsigned b, unsigned n) {
    int l = 0;
    while (l + 1 < L.n_levels && n >= 9u * (unsigned)L.cells[l]) { n -= 9u * (unsigned)L.cells[l]; ++l; }
    const unsigned cell = n / 9u;
    LevelPos q;
    q.l = l; q.a = n - cell * 9u; q.cell_index = (size_t)b * L.cells[l] + cell;
    return q;
}

// log(1 - x) for 0 <= x <= 1/4: x' = -x, log1p(x') = x' + x'^2 P(x'), P = degree-7 minimax fit of
// (log1p(x') - x') / x'^2 on [-1/4, 0] (max relative error 7e-8 in fp32 arithmetic, i.e. rounding level --
// checked against float64 on 2e5 log-spaced points).  Nine FMA-pipe instructions instead of log1pf's ~30.
__device__ __forceinline__ float log1m_small(float x) {
    const float f = -x;
    float t = 0.3018442392349243f;
    t = fmaf(t, f, -0.0319586880505085f);
    t = fmaf(t, f, 0.16445574164390564f);
    t = fmaf(t, f, -0.1639823168516159f);
    t = fmaf(t, f, 0.2001783549785614f);
    t = fmaf(t, f, -0.2499942183494568f);
    t = fmaf(t, f, 0.33333340287208557f);
    t = fmaf(t, f, -0.5f);
    return fmaf(f * f, t, f);
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}

// focal loss of one element + its gradient w.r.t. the logit.  t in {0, 1}:
//   keras binary_crossentropy (clip -> logit -> sigmoid CE) == -log(pc) | -log(1 - pc) (the
//   three-transcendental form max(z,0) - z t + log1p(exp(-|z|)) of utils/tpu.py:135 is the same number);
//   (1-p_t)^gamma by sqrt for the reference's gamma = 1.5.  `lg` = log(pc) (foreground) or log(1 - pc)
//   (background), supplied by the caller: almost every element is a background anchor with a small score,
//   for which log(1 - pc) is the short polynomial above; the kernel takes the logf path only for warps that
//   hold a foreground element or a score above 1/4 (r2 capture: the kernel was issue-bound at 2.6 TB/s on
//   logf + log1pf + sqrtf evaluated for every element).
template <int GMODE>   // 0: generic gamma, 1: gamma == 1.5, 2: gamma == 2
__device__ __forceinline__ float focal_elem(float p, bool fg, float lg, float mask, float alpha, float gamma,
                                            float gscale, float &loss_acc) {
    const float af = fg ? alpha : 1.f - alpha;
    const float fw = fg ? 1.f - p : p;
    const float bce = -lg;
    float fwg, dfwg_abs;                       // fw^gamma, gamma * fw^(gamma-1)
    if (GMODE == 1) { const float sq = sqrt_approx(fw); fwg = fw * sq; dfwg_abs = 1.5f * sq; }
    else if (GMODE == 2) { fwg = fw * fw; dfwg_abs = 2.f * fw; }
    else { fwg = powf(fw, gamma); dfwg_abs = fw > 0.f ? gamma * powf(fw, gamma - 1.f) : 0.f; }
    loss_acc += af * fwg * bce * mask;
    const float dfwg = fg ? -dfwg_abs : dfwg_abs;
    const bool in_range = p > 1e-7f && p < 1.f - 1e-7f;
    // d/dlogit = dL/dp * p (1-p);  dbce/dp * p (1-p) = (p - t) inside the clip range, 0 outside
    const float t = fg ? 1.f : 0.f;
    const float g = af * (dfwg * bce * p * (1.f - p) + (in_range ? fwg * (p - t) : 0.f));
    return g * mask * gscale;
}

template <int VEC, int GMODE>
__global__ void __launch_bounds__(256)
focal_kernel(const float *__restrict__ p_, const float *__restrict__ labels_t,
             const int8_t *__restrict__ state, const int32_t *__restrict__ cls, unsigned rows, int C,
             float alpha, float gamma, float grad_scale, const float *__restrict__ norm,
             float *__restrict__ dlogit, float *__restrict__ partial, const LevelGrads L) {
    __shared__ float sh[8];
    const float gscale = norm[5] * grad_scale;
    const unsigned vpr = (unsigned)C / VEC;                // vectors per row
    const unsigned total = rows * vpr;
    float acc = 0.f;
    for (unsigned v = blockIdx.x * 256u + threadIdx.x; v < total; v += gridDim.x * 256u) {
        const unsigned r = v / vpr;
        const int c0 = (int)(v - r * vpr) * VEC;
        float p[VEC], t[VEC], st;
        const float *src = p_ + (size_t)r * C + c0;
        if (VEC == 4) { const float4 q = *reinterpret_cast<const float4 *>(src); p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }
        else if (VEC == 2) { const float2 q = *reinterpret_cast<const float2 *>(src); p[0] = q.x; p[1] = q.y; }
        else p[0] = src[0];
        if (labels_t) {
            const float *lt = labels_t + (size_t)r * (C + 1);
#pragma unroll
            for (int k = 0; k < VEC; ++k) t[k] = lt[c0 + k];
            st = lt[C];
        } else {
            const int cl = cls[r];
#pragma unroll
            for (int k = 0; k < VEC; ++k) t[k] = (cl == c0 + k) ? 1.f : 0.f;
            st = (float)state[r];
        }
        const float mask = st != -1.f ? 1.f : 0.f;
        float g[VEC], pc[VEC], lg[VEC];
        bool slow = false;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            pc[k] = fminf(fmaxf(p[k], 1e-7f), 1.f - 1e-7f);
            lg[k] = log1m_small(pc[k]);                         // right for background elements with pc <= 1/4
            slow |= (t[k] == 1.f) | (pc[k] > 0.25f);
        }
        if (__any_sync(__activemask(), slow)) {                 // warp-level branch: rare
#pragma unroll
            for (int k = 0; k < VEC; ++k)
                if ((t[k] == 1.f) | (pc[k] > 0.25f)) lg[k] = t[k] == 1.f ? logf(pc[k]) : logf(1.f - pc[k]);
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k)
            g[k] = focal_elem<GMODE>(p[k], t[k] == 1.f, lg[k], mask, alpha, gamma, gscale, acc);
        if (dlogit) {                                           // (optional fp32 copy: see effdet_detection_losses)
            float *dst = dlogit + (size_t)r * C + c0;
            if (VEC == 4) *reinterpret_cast<float4 *>(dst) = make_float4(g[0], g[1], g[2], g[3]);
            else if (VEC == 2) *reinterpret_cast<float2 *>(dst) = make_float2(g[0], g[1]);
            else dst[0] = g[0];
        }
        if (L.n_levels) {
            const unsigned b = r / L.N;
            const LevelPos q = level_pos(L, b, r - b * L.N);
            __nv_bfloat16 *o = L.cls[q.l] + q.cell_index * L.cpad_cls + q.a * C + c0;
            if (VEC == 4) {
                uint2 pk;
                *reinterpret_cast<__nv_bfloat162 *>(&pk.x) = __floats2bfloat162_rn(g[0], g[1]);
                *reinterpret_cast<__nv_bfloat162 *>(&pk.y) = __floats2bfloat162_rn(g[2], g[3]);
                *reinterpret_cast<uint2 *>(o) = pk;
            } else if (VEC == 2) {
                *reinterpret_cast<__nv_bfloat162 *>(o) = __floats2bfloat162_rn(g[0], g[1]);
            } else {
                o[0] = __float2bfloat16_rn(g[0]);
            }
        }
    }
    float t = block_sum_256(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void __launch_bounds__(256)
smooth_l1_kernel(const float *__restrict__ pred, const float *__restrict__ reg_t, unsigned rows,
                 float delta, float grad_scale, const float *__restrict__ norm,
                 float *__restrict__ dreg, float *__restrict__ partial, const LevelGrads L) {
    __shared__ float sh[8];
    const float inv_norm = norm[4];
    float acc = 0.f;
    for (unsigned r = blockIdx.x * 256u + threadIdx.x; r < rows; r += gridDim.x * 256u) {
        const float4 pr = *reinterpret_cast<const float4 *>(pred + (size_t)r * 4);
        const float *tg = reg_t + (size_t)r * 5;
        const bool fg = tg[4] == 1.f;
        const float pv[4] = {pr.x, pr.y, pr.z, pr.w};
        float g[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float d = pv[k] - tg[k], ad = fabsf(d);
            const float l = ad > delta ? ad - 0.5f : 0.5f * d * d;
            const float gd = ad > delta ? (d > 0.f ? 1.f : -1.f) : d;
            acc += fg ? l : 0.f;
            g[k] = fg ? gd * inv_norm * grad_scale : 0.f;
        }
        if (dreg) *reinterpret_cast<float4 *>(dreg + (size_t)r * 4) = make_float4(g[0], g[1], g[2], g[3]);
        if (L.n_levels) {
            const unsigned b = r / L.N;
            const LevelPos q = level_pos(L, b, r - b * L.N);
            uint2 pk;
            *reinterpret_cast<__nv_bfloat162 *>(&pk.x) = __floats2bfloat162_rn(g[0], g[1]);
            *reinterpret_cast<__nv_bfloat162 *>(&pk.y) = __floats2bfloat162_rn(g[2], g[3]);
            *reinterpret_cast<uint2 *>(L.reg[q.l] + q.cell_index * L.cpad_reg + q.a * 4) = pk;
        }
    }
    float t = block_sum_256(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// fixed-order parallel sums of the per-block partials (deterministic)
__global__ void __launch_bounds__(256)
finalize_losses_kernel(const float *__restrict__ pf, int nf, const float *__restrict__ ps,
                       int ns, float *__restrict__ out) {
    __shared__ float sh[8];
    float a = 0.f, b = 0.f;
    for (int i = threadIdx.x; i < nf; i += 256) a += pf[i];
    for (int i = threadIdx.x; i < ns; i += 256) b += ps[i];
    const float ta = block_sum_256(a, sh);
    __syncthreads();
    const float tb = block_sum_256(b, sh);
    if (threadIdx.x == 0) {
        out[0] = ta * out[5];     // focal
        out[1] = tb * out[4];     // smooth L1
    }
}

}  // namespace effdet

using namespace effdet;

static const int kLossBlocks = 148 * 8;

extern "C" size_t effdet_detection_losses_workspace_size(void) { return 4 * kLossBlocks * sizeof(float); }

extern "C" int effdet_detection_losses(const float *classification, const float *regression,
                                       const float *regression_t, const float *labels_t,
                                       const int8_t *state, const int32_t *cls, int B, size_t N,
                                       int C, float alpha, float gamma, float delta,
                                       float grad_scale, float *dcls_logits, float *dreg,
                                       float *out8, void *workspace, size_t workspace_bytes,
                                       void *const *dcls_levels_host, void *const *dreg_levels_host,
                                       const int *level_cells_host, int n_levels, int cpad_cls,
                                       int cpad_reg, void *stream) {
    EFFDET_REQUIRE(classification && regression && regression_t && out8 && workspace, "null pointer");
    // the concatenated fp32 gradients are optional when the per-level bf16 copies are requested (the tensor-core
    // gradient kernels read only those; D0 training: 150 MB of dead stores per step otherwise)
    EFFDET_REQUIRE((dcls_logits && dreg) || n_levels > 0, "dcls_logits / dreg may be NULL only with level outputs");
    EFFDET_REQUIRE(labels_t || (state && cls), "need dense labels or compact (state, cls) targets");
    EFFDET_REQUIRE(B > 0 && N > 0 && C > 0, "bad sizes");
    EFFDET_REQUIRE((reinterpret_cast<uintptr_t>(regression) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(dreg) & 15) == 0, "16B alignment");   // (NULL passes)
    if (workspace_bytes < effdet_detection_losses_workspace_size())
        return fail(EFFDET_E_CAPACITY, "effdet_detection_losses: workspace too small%s", "");
    cudaStream_t st = as_stream(stream);
    const size_t rows = (size_t)B * N;
    LevelGrads L;
    memset(&L, 0, sizeof(L));
    if (n_levels > 0) {
        EFFDET_REQUIRE(n_levels <= 5 && dcls_levels_host && dreg_levels_host && level_cells_host, "bad level outputs");
        EFFDET_REQUIRE(cpad_cls >= 9 * C && cpad_reg >= 36, "channel padding too small");
        size_t tot = 0;
        for (int l = 0; l < n_levels; ++l) {
            L.cls[l] = static_cast<__nv_bfloat16 *>(dcls_levels_host[l]);
            L.reg[l] = static_cast<__nv_bfloat16 *>(dreg_levels_host[l]);
            L.cells[l] = level_cells_host[l];
            tot += (size_t)9 * level_cells_host[l];
        }
        EFFDET_REQUIRE(tot == N, "level cells do not add up to N / 9");
        L.n_levels = n_levels; L.cpad_cls = cpad_cls; L.cpad_reg = cpad_reg; L.N = (unsigned)N;
    }
    float *pf = static_cast<float *>(workspace), *ps = pf + kLossBlocks;
    unsigned *cnt = reinterpret_cast<unsigned *>(ps + kLossBlocks);
    int nc = (int)cdiv(rows, 256 * 8); if (nc > kLossBlocks) nc = kLossBlocks; if (nc < 1) nc = 1;
    count_pos_kernel<<<nc, 256, 0, st>>>(regression_t, labels_t, state, rows, C, cnt);
    EFFDET_LAUNCHED();
    count_pos_finalize_kernel<<<1, 256, 0, st>>>(cnt, nc, out8);
    EFFDET_LAUNCHED();
    EFFDET_REQUIRE(rows * (size_t)C < 0xffffffffull, "B*N*C must fit 32 bits");
    const bool al = (reinterpret_cast<uintptr_t>(classification) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(dcls_logits) & 15) == 0 && (n_levels == 0 || cpad_cls % 4 == 0);
    const int vec = (C % 4 == 0 && al) ? 4 : (C % 2 == 0 && al) ? 2 : 1;
    const int gmode = gamma == 1.5f ? 1 : gamma == 2.f ? 2 : 0;
    int nf = (int)cdiv(rows * C / vec, 256 * 4); if (nf > kLossBlocks) nf = kLossBlocks; if (nf < 1) nf = 1;
#define FOCAL(V, G)                                                                                      \
    focal_kernel<V, G><<<nf, 256, 0, st>>>(classification, labels_t, state, cls, (unsigned)rows, C, alpha, gamma, \
                                           grad_scale, out8, dcls_logits, pf, L)
#define FOCAL_G(V)                                                                                       \
    do { if (gmode == 1) FOCAL(V, 1); else if (gmode == 2) FOCAL(V, 2); else FOCAL(V, 0); } while (0)
    if (vec == 4) FOCAL_G(4); else if (vec == 2) FOCAL_G(2); else FOCAL_G(1);
#undef FOCAL_G
#undef FOCAL
    EFFDET_LAUNCHED();
    int ns = (int)cdiv(rows, 256 * 4); if (ns > kLossBlocks) ns = kLossBlocks; if (ns < 1) ns = 1;
    smooth_l1_kernel<<<ns, 256, 0, st>>>(regression, regression_t, (unsigned)rows, delta, grad_scale, out8,
                                         dreg, ps, L);
    EFFDET_LAUNCHED();
    finalize_losses_kernel<<<1, 256, 0, st>>>(pf, nf, ps, ns, out8);
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}
