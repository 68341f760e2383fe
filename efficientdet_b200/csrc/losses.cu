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
__device__ __forceinline__ void level_store(const LevelGrads &L, bool is_cls, size_t r, int k, int per,
                                            float v) {
    const unsigned b = (unsigned)(r / L.N);
    unsigned n = (unsigned)(r - (size_t)b * L.N);
    int l = 0;
    while (l + 1 < L.n_levels && n >= 9u * (unsigned)L.cells[l]) { n -= 9u * (unsigned)L.cells[l]; ++l; }
    const unsigned cell = n / 9u, a = n - cell * 9u;
    const int cpad = is_cls ? L.cpad_cls : L.cpad_reg;
    __nv_bfloat16 *dst = is_cls ? L.cls[l] : L.reg[l];
    dst[((size_t)b * L.cells[l] + cell) * cpad + a * per + k] = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(256)
focal_kernel(const float *__restrict__ p_, const float *__restrict__ labels_t,
             const int8_t *__restrict__ state, const int32_t *__restrict__ cls, size_t rows, int C,
             float alpha, float gamma, float grad_scale, const float *__restrict__ norm,
             float *__restrict__ dlogit, float *__restrict__ partial, const LevelGrads L) {
    __shared__ float sh[8];
    const float inv_norm = norm[5];
    const size_t total = rows * (size_t)C;
    float acc = 0.f;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const size_t r = i / C;
        const int c = (int)(i - r * C);
        float t, st;
        if (labels_t) {
            t = labels_t[r * (size_t)(C + 1) + c];
            st = labels_t[r * (size_t)(C + 1) + C];
        } else {
            t = (cls[r] == c) ? 1.f : 0.f;
            st = (float)state[r];
        }
        const float p = p_[i];
        const bool fg = t == 1.f;
        const float af = fg ? alpha : 1.f - alpha;
        const float fw = fg ? 1.f - p : p;
        const float pc = fminf(fmaxf(p, 1e-7f), 1.f - 1e-7f);
        const float z = logf(pc / (1.f - pc));
        const float bce = fmaxf(z, 0.f) - z * t + log1pf(expf(-fabsf(z)));
        const float fwg = powf(fw, gamma);
        const float mask = st != -1.f ? 1.f : 0.f;
        acc += af * fwg * bce * mask;
        const float dbce = (p > 1e-7f && p < 1.f - 1e-7f) ? (pc - t) / (pc * (1.f - pc)) : 0.f;
        const float dfw = fg ? -1.f : 1.f;
        const float dfwg = fw > 0.f ? gamma * powf(fw, gamma - 1.f) * dfw : 0.f;
        const float dLdp = af * (dfwg * bce + fwg * dbce);
        const float gout = dLdp * p * (1.f - p) * mask * inv_norm * grad_scale;
        dlogit[i] = gout;
        if (L.n_levels) level_store(L, true, r, c, C, gout);
    }
    float t = block_sum_256(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void __launch_bounds__(256)
smooth_l1_kernel(const float *__restrict__ pred, const float *__restrict__ reg_t, size_t rows,
                 float delta, float grad_scale, const float *__restrict__ norm,
                 float *__restrict__ dreg, float *__restrict__ partial, const LevelGrads L) {
    __shared__ float sh[8];
    const float inv_norm = norm[4];
    float acc = 0.f;
    for (size_t r = (size_t)blockIdx.x * 256 + threadIdx.x; r < rows; r += (size_t)gridDim.x * 256) {
        const float4 pr = *reinterpret_cast<const float4 *>(pred + r * 4);
        const float *tg = reg_t + r * 5;
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
        *reinterpret_cast<float4 *>(dreg + r * 4) = make_float4(g[0], g[1], g[2], g[3]);
        if (L.n_levels) {
#pragma unroll
            for (int k = 0; k < 4; ++k) level_store(L, false, r, k, 4, g[k]);
        }
    }
    float t = block_sum_256(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void finalize_losses_kernel(const float *__restrict__ pf, int nf, const float *__restrict__ ps,
                                       int ns, float *__restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < nf; ++i) a += pf[i];
        for (int i = 0; i < ns; ++i) b += ps[i];
        out[0] = a * out[5];     // focal
        out[1] = b * out[4];     // smooth L1
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
    EFFDET_REQUIRE(classification && regression && regression_t && dcls_logits && dreg && out8 &&
                       workspace, "null pointer");
    EFFDET_REQUIRE(labels_t || (state && cls), "need dense labels or compact (state, cls) targets");
    EFFDET_REQUIRE(B > 0 && N > 0 && C > 0, "bad sizes");
    EFFDET_REQUIRE((reinterpret_cast<uintptr_t>(regression) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(dreg) & 15) == 0, "16B alignment");
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
    int nf = (int)cdiv(rows * C, 256 * 8); if (nf > kLossBlocks) nf = kLossBlocks; if (nf < 1) nf = 1;
    focal_kernel<<<nf, 256, 0, st>>>(classification, labels_t, state, cls, rows, C, alpha, gamma,
                                     grad_scale, out8, dcls_logits, pf, L);
    EFFDET_LAUNCHED();
    int ns = (int)cdiv(rows, 256 * 4); if (ns > kLossBlocks) ns = kLossBlocks; if (ns < 1) ns = 1;
    smooth_l1_kernel<<<ns, 256, 0, st>>>(regression, regression_t, rows, delta, grad_scale, out8,
                                         dreg, ps, L);
    EFFDET_LAUNCHED();
    finalize_losses_kernel<<<1, 32, 0, st>>>(pf, nf, ps, ns, out8);
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}
