// Training-mode kernels needed only when the EfficientNet backbone itself is trained
// (train_tpu.py without --freeze-backbone; BASELINE config 4): backward of
//   efficientnet.py:228-237  expand conv + BN + swish          -> bn_act_backward (swish')
//   efficientnet.py:242-252  depthwise k3/k5, stride 1/2       -> dw_wgrad_general, dw_dgrad_strided
//   efficientnet.py:255-286  squeeze-excite                    -> se_apply, se_bwd_reduce, se_fc_backward,
//                                                                 se_bwd_finish
//   efficientnet.py:413-423  stem conv                         -> stem_wgrad
// Same conventions as train.cu: NHWC, fp32 math, 16-byte vectors along C, deterministic
// fixed-order reductions.
#include "common.cuh"

namespace effdet {

template <typename T, int CV> struct VecB;
template <> struct VecB<float, 4> {
    static __device__ __forceinline__ void load(const float *p, float *v) {
        float4 t = *reinterpret_cast<const float4 *>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float *p, const float *v) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct VecB<__nv_bfloat16, 8> {
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
template <int CV> __device__ __forceinline__ void ldv(const float *p, float *v) {
#pragma unroll
    for (int i = 0; i < CV; i += 4) {
        float4 t = *reinterpret_cast<const float4 *>(p + i);
        v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
}

// ------------------------------------------------------------------ BN + activation backward
// u = z*a + b with a = gamma*invstd, b = beta - mean*a (training) or the folded inference scale/shift.
// pass 1: partial[blk][0][c] = sum dy*act'(u), partial[blk][1][c] = sum dy*act'(u)*(z - mean)
// (the finalize kernel multiplies the second sum by invstd).
// Both passes are pure streaming kernels, so what matters is bytes in flight per SM: a thread owns CV = 4
// channels (8-byte bf16 / 16-byte fp32 vectors) and keeps U = 4 rows of both tensors in flight; the
// per-channel coefficients then cost 20 registers instead of 48 and four blocks fit an SM (the 8-channel /
// 2-row version needed 100 registers: two blocks per SM, ~32 KB in flight, ~2.4 TB/s).
template <> struct VecB<__nv_bfloat16, 4> {
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float *v) {
        const uint2 t = *reinterpret_cast<const uint2 *>(p);
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
        v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float *v) {
        uint2 t;
        __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&t);
        h[0] = __floats2bfloat162_rn(v[0], v[1]);
        h[1] = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2 *>(p) = t;
    }
};
template <typename T, int ACT> __device__ __forceinline__ float act_grad_t(float u) {
    return act_grad_io<T>(u, ACT);
}
// Raw (still packed) 4-channel vectors.  ptxas sinks independent loads next to their uses, which leaves two
// loads in flight per thread however the loop is unrolled (seen in the SASS of every streaming kernel here:
// LDG, LDG, math, LDG, LDG, math ...; ncu: long-scoreboard stalls 11.6 per issue, DRAM at 50 %, 16 KB in
// flight per SM); all_loaded() ties every value of a batch to all of its loads.
template <typename T> struct Raw4;
template <> struct Raw4<__nv_bfloat16> {
    typedef uint2 type;
    static __device__ __forceinline__ type load(const __nv_bfloat16 *p) { return *reinterpret_cast<const uint2 *>(p); }
    static __device__ __forceinline__ void unpack(const type &t, float *v) {
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
        v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
    }
    // `zero` is a kernel argument that is always 0: every value is OR-ed with (xor of ALL loaded words) & zero,
    // a data dependence the assembler cannot remove, so the eight loads are in flight together (an empty asm
    // with in/out operands orders the PTX but ptxas re-sinks the loads)
    static __device__ __forceinline__ void all_loaded(type (&a)[4], type (&b)[4], uint32_t zero) {
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) t ^= a[i].x ^ a[i].y ^ b[i].x ^ b[i].y;
        t &= zero;
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i].x |= t; a[i].y |= t; b[i].x |= t; b[i].y |= t; }
    }
};
template <> struct Raw4<float> {
    typedef float4 type;
    static __device__ __forceinline__ type load(const float *p) { return *reinterpret_cast<const float4 *>(p); }
    static __device__ __forceinline__ void unpack(const type &t, float *v) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    static __device__ __forceinline__ void all_loaded(type (&)[4], type (&)[4], uint32_t) {}   // accuracy mode: as is
};
constexpr int kBnU = 4;       // rows per thread and loop iteration
// ptxas sinks each load of an unrolled batch next to its first use whatever the source order (SASS: LDG, LDG,
// math, LDG, LDG, math ...), so a thread never has more than two loads in flight and these passes ran at
// ~3.5 TB/s.  The bytes in flight are therefore supplied by the TMA engine instead: one thread per block asks
// for the block's next row ranges with cp.async.bulk.prefetch.L2 (no registers, no shared memory), kBnPD loop
// iterations ahead, and the register loads then hit L2.
constexpr int kBnPD = 3;
__device__ __forceinline__ void bulk_prefetch_l2(const void *p, size_t bytes) {
    if (bytes == 0) return;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"((uint32_t)bytes) : "memory");
}
// rows [ra, rb) of a dense (rows, C) matrix of T, clipped to r1; row size is a multiple of 16 bytes (C % 8 == 0)
template <typename T>
__device__ __forceinline__ void prefetch_rows(const T *base, size_t ra, size_t rb, size_t r1, int C) {
    if (rb > r1) rb = r1;
    if (ra < rb) bulk_prefetch_l2(base + ra * C, (rb - ra) * (size_t)C * sizeof(T));
}
template <typename T, int CV, int ACT>
__global__ void __launch_bounds__(1024, 1)      // <= 64 registers
bn_act_bwd_reduce_kernel(const T *__restrict__ z, const T *__restrict__ dy,
                                         const float *__restrict__ ua, const float *__restrict__ ub,
                                         const float *__restrict__ mean, size_t rows, int C,
                                         int rows_per_block, float *__restrict__ partial, uint32_t zero) {
    extern __shared__ float sred[];
    const int nvec = C / CV, PY = blockDim.x / nvec;
    const int cv = threadIdx.x % nvec, py = threadIdx.x / nvec, c = cv * CV;
    const size_t r0 = (size_t)blockIdx.x * rows_per_block;
    const size_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    float s1[CV], s2[CV], a[CV], b[CV], mu[CV];
    ldv<CV>(ua + c, a); ldv<CV>(ub + c, b); ldv<CV>(mean + c, mu);
#pragma unroll
    for (int k = 0; k < CV; ++k) { s1[k] = 0.f; s2[k] = 0.f; }
    size_t r = r0 + py;
    const size_t step = (size_t)kBnU * PY;
    if (threadIdx.x == 0) {
        prefetch_rows(z, r0, r0 + kBnPD * step, r1, C);
        prefetch_rows(dy, r0, r0 + kBnPD * step, r1, C);
    }
    size_t rq = r0 + kBnPD * step;             // first row not yet requested
    for (; r + (size_t)(kBnU - 1) * PY < r1; r += step, rq += step) {
        static_assert(CV == 4 && kBnU == 4, "raw-vector batch is written for 4 channels x 4 rows");
        if (threadIdx.x == 0) {
            prefetch_rows(z, rq, rq + step, r1, C);
            prefetch_rows(dy, rq, rq + step, r1, C);
        }
        typename Raw4<T>::type zr[kBnU], gr[kBnU];
#pragma unroll
        for (int u = 0; u < kBnU; ++u) {
            zr[u] = Raw4<T>::load(z + (r + (size_t)u * PY) * C + c);
            gr[u] = Raw4<T>::load(dy + (r + (size_t)u * PY) * C + c);
        }
        Raw4<T>::all_loaded(zr, gr, zero);
#pragma unroll
        for (int u = 0; u < kBnU; ++u) {
            float zz[CV], g[CV];
            Raw4<T>::unpack(zr[u], zz);
            Raw4<T>::unpack(gr[u], g);
#pragma unroll
            for (int k = 0; k < CV; ++k) {
                const float gm = g[k] * act_grad_t<T, ACT>(fmaf(zz[k], a[k], b[k]));
                s1[k] += gm;
                s2[k] = fmaf(gm, zz[k] - mu[k], s2[k]);
            }
        }
    }
    for (; r < r1; r += PY) {
        float zz[CV], g[CV];
        VecB<T, CV>::load(z + r * C + c, zz);
        VecB<T, CV>::load(dy + r * C + c, g);
#pragma unroll
        for (int k = 0; k < CV; ++k) {
            const float gm = g[k] * act_grad_t<T, ACT>(fmaf(zz[k], a[k], b[k]));
            s1[k] += gm;
            s2[k] = fmaf(gm, zz[k] - mu[k], s2[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < CV; ++k) {
        sred[((size_t)py * 2 + 0) * C + c + k] = s1[k];
        sred[((size_t)py * 2 + 1) * C + c + k] = s2[k];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        float t = 0.f;
        for (int q = 0; q < PY; ++q) t += sred[(size_t)q * 2 * C + i];
        partial[(size_t)blockIdx.x * 2 * C + i] = t;
    }
}
__global__ void bn_act_bwd_finalize_kernel(const float *__restrict__ partial, int nblk, double count,
                                           const float *__restrict__ gamma, const float *__restrict__ mean,
                                           const float *__restrict__ invstd, float *__restrict__ k123,
                                           float *__restrict__ dgamma, float *__restrict__ dbeta, int C) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= C) return;
    const int lane = threadIdx.x & 31;
    double s1 = 0.0, s2 = 0.0;
    int b = lane;
    for (; b + 7 * 32 < nblk; b += 8 * 32) {      // sixteen loads in flight, same summation order
        float v[8], q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float *row = partial + (size_t)(b + 32 * u) * 2 * C;
            v[u] = __ldcg(row + c);
            q[u] = __ldcg(row + C + c);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { s1 += (double)v[u]; s2 += (double)q[u]; }
    }
    for (; b < nblk; b += 32) {
        s1 += (double)__ldcg(partial + (size_t)b * 2 * C + c);
        s2 += (double)__ldcg(partial + (size_t)b * 2 * C + C + c);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, d);
        s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    }
    if (lane) return;
    const float g = gamma[c], is = invstd[c], mu = mean[c];
    s2 *= (double)is;                       // the reduce pass accumulated dy*act'*(z - mean)
    const float m1 = (float)(s1 / count), m2 = (float)(s2 / count);
    k123[c] = g * is;
    k123[C + c] = -g * is * is * m2;
    k123[2 * C + c] = -g * is * (m1 - mu * is * m2);
    if (dgamma) dgamma[c] = (float)s2;
    if (dbeta) dbeta[c] = (float)s1;
}
// pass 2: dz = k1*(dy*act'(u)) + k2*z + k3
// block = nvec channel vectors x PY rows; a thread keeps the seven per-channel coefficient vectors of
// ITS channels in registers and walks rows (the flat-index version re-read them from L1 for every
// 16-byte vector: 10 extra load instructions per 8 elements, LSU-bound at ~1.5 TB/s).
template <typename T, int CV, int ACT>
__global__ void __launch_bounds__(1024, 1)      // <= 64 registers: four 256-thread blocks per SM (blocks grow to C/4 threads)
bn_act_bwd_apply_kernel(const T *__restrict__ dy, const T *__restrict__ z, const float *__restrict__ ua,
                        const float *__restrict__ ub, const float *__restrict__ k123, T *__restrict__ dz,
                        size_t rows, int C, int rows_per_block, uint32_t zero) {
    const int nvec = C / CV, PY = blockDim.x / nvec;
    const int cv = threadIdx.x % nvec, py = threadIdx.x / nvec, c = cv * CV;
    if (py >= PY) return;
    float a[CV], b[CV], k1[CV], k2[CV], k3[CV];
    ldv<CV>(ua + c, a); ldv<CV>(ub + c, b);
    ldv<CV>(k123 + c, k1); ldv<CV>(k123 + C + c, k2); ldv<CV>(k123 + 2 * C + c, k3);
    const size_t r0 = (size_t)blockIdx.x * rows_per_block;
    const size_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    size_t r = r0 + py;
    const size_t step = (size_t)kBnU * PY;
    if (threadIdx.x == 0) {
        prefetch_rows(z, r0, r0 + kBnPD * step, r1, C);
        prefetch_rows(dy, r0, r0 + kBnPD * step, r1, C);
    }
    size_t rq = r0 + kBnPD * step;             // first row not yet requested
    for (; r + (size_t)(kBnU - 1) * PY < r1; r += step, rq += step) {
        static_assert(CV == 4 && kBnU == 4, "raw-vector batch is written for 4 channels x 4 rows");
        if (threadIdx.x == 0) {
            prefetch_rows(z, rq, rq + step, r1, C);
            prefetch_rows(dy, rq, rq + step, r1, C);
        }
        typename Raw4<T>::type zr[kBnU], gr[kBnU];
#pragma unroll
        for (int u = 0; u < kBnU; ++u) {
            gr[u] = Raw4<T>::load(dy + (r + (size_t)u * PY) * C + c);
            zr[u] = Raw4<T>::load(z + (r + (size_t)u * PY) * C + c);
        }
        Raw4<T>::all_loaded(zr, gr, zero);
#pragma unroll
        for (int u = 0; u < kBnU; ++u) {
            float g[CV], zz[CV];
            Raw4<T>::unpack(gr[u], g);
            Raw4<T>::unpack(zr[u], zz);
#pragma unroll
            for (int k = 0; k < CV; ++k)
                g[k] = k1[k] * (g[k] * act_grad_t<T, ACT>(fmaf(zz[k], a[k], b[k]))) + k2[k] * zz[k] + k3[k];
            VecB<T, CV>::store(dz + (r + (size_t)u * PY) * C + c, g);
        }
    }
    for (; r < r1; r += PY) {
        float g0[CV], z0[CV];
        VecB<T, CV>::load(dy + r * C + c, g0);
        VecB<T, CV>::load(z + r * C + c, z0);
#pragma unroll
        for (int k = 0; k < CV; ++k)
            g0[k] = k1[k] * (g0[k] * act_grad_t<T, ACT>(fmaf(z0[k], a[k], b[k]))) + k2[k] * z0[k] + k3[k];
        VecB<T, CV>::store(dz + r * C + c, g0);
    }
}

// ------------------------------------------------------------------ squeeze-excite
// yg = y * gate[b][c]
template <typename T, int CV>
__global__ void __launch_bounds__(256)
se_apply_kernel(const T *__restrict__ y, const float *__restrict__ gate, T *__restrict__ out, int HW, int C,
                size_t nvec_total) {
    const int nvec = C / CV;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvec_total; i += (size_t)gridDim.x * 256) {
        const int c = (int)(i % nvec) * CV;
        const size_t b = i / ((size_t)nvec * HW);
        float v[CV], g[CV];
        VecB<T, CV>::load(y + i * CV, v);
        ldv<CV>(gate + b * C + c, g);
#pragma unroll
        for (int k = 0; k < CV; ++k) v[k] *= g[k];
        VecB<T, CV>::store(out + i * CV, v);
    }
}
// partial[b][blk][c] = sum over the block's pixels of dyg * y     (d gate before the sigmoid);
// y == NULL: plain spatial sum of dyg (the squeeze of the forward pass)
template <typename T, int CV>
__global__ void se_bwd_reduce_kernel(const T *__restrict__ dyg, const T *__restrict__ y, int HW, int C,
                                     int rows_per_block, float *__restrict__ partial, uint32_t zero) {
    extern __shared__ float sred[];
    const int nvec = C / CV, PY = blockDim.x / nvec;
    const int cv = threadIdx.x % nvec, py = threadIdx.x / nvec, c = cv * CV;
    const int b = blockIdx.y;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, HW);
    const T *pa = dyg + (size_t)b * HW * C, *pb = y ? y + (size_t)b * HW * C : nullptr;
    float s[CV];
#pragma unroll
    for (int k = 0; k < CV; ++k) s[k] = 0.f;
    int r = r0 + py;
    if (pb) {
        for (; r + 3 * PY < r1; r += 4 * PY) {          // four rows of both tensors in flight
            uint4 raw[8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                raw[2 * u] = ld16(pa + (size_t)(r + u * PY) * C + c);
                raw[2 * u + 1] = ld16(pb + (size_t)(r + u * PY) * C + c);
            }
            tie_loads(raw, zero);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float a[CV], v[CV];
                Unpack16<T, CV>::run(raw[2 * u], a);
                Unpack16<T, CV>::run(raw[2 * u + 1], v);
#pragma unroll
                for (int k = 0; k < CV; ++k) s[k] = fmaf(a[k], v[k], s[k]);
            }
        }
    } else {
        for (; r + 3 * PY < r1; r += 4 * PY) {
            uint4 raw[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) raw[u] = ld16(pa + (size_t)(r + u * PY) * C + c);
            tie_loads(raw, zero);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float a[CV];
                Unpack16<T, CV>::run(raw[u], a);
#pragma unroll
                for (int k = 0; k < CV; ++k) s[k] += a[k];
            }
        }
    }
    for (; r < r1; r += PY) {
        float a[CV], v[CV];
        VecB<T, CV>::load(pa + (size_t)r * C + c, a);
        if (pb) {
            VecB<T, CV>::load(pb + (size_t)r * C + c, v);
#pragma unroll
            for (int k = 0; k < CV; ++k) s[k] = fmaf(a[k], v[k], s[k]);
        } else {
#pragma unroll
            for (int k = 0; k < CV; ++k) s[k] += a[k];
        }
    }
#pragma unroll
    for (int k = 0; k < CV; ++k) sred[(size_t)py * C + c + k] = s[k];
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        float t = 0.f;
        for (int r = 0; r < PY; ++r) t += sred[(size_t)r * C + i];
        partial[((size_t)b * gridDim.x + blockIdx.x) * C + i] = t;
    }
}
// SE FC backward in four short, wide phases (the old one-block-per-image kernel spent its time in
// serial chains of L2 round trips).  Scratch layout (floats), S = B*(C+R):
//   mean[B][C] | s1[B][R] | rr[B][R] | ds2[B][C] | ds1[B][R] | ds1p[B][nch][R]
// phase 1 (grid B): squeeze mean, FC1 pre-activation s1 and rr = swish(s1)        (forward recompute)
__global__ void __launch_bounds__(512)
se_bwd_phase1_kernel(const float *__restrict__ se_sum, int se_blocks, float inv_hw, const float *__restrict__ w1,
                     const float *__restrict__ b1, int C, int R, float *__restrict__ mean_o,
                     float *__restrict__ s1_o, float *__restrict__ rr_o) {
    extern __shared__ float sm[];       // mean[C] | part[512]
    float *mean = sm, *part = sm + C;
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int c = tid; c < C; c += 512) {
        const float *src = se_sum + (size_t)b * se_blocks * C + c;
        float t = 0.f;
        for (int k = 0; k < se_blocks; ++k) t += src[(size_t)k * C];
        mean[c] = t * inv_hw;
        mean_o[(size_t)b * C + c] = t * inv_hw;
    }
    __syncthreads();
    const int G = 512 / R;              // R <= 512
    const int j = tid % R, cg = tid / R;
    if (cg < G) {
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        int c = cg;
        for (; c + 3 * G < C; c += 4 * G) {
#pragma unroll
            for (int u = 0; u < 4; ++u) s4[u] = fmaf(mean[c + u * G], w1[(size_t)(c + u * G) * R + j], s4[u]);
        }
        for (int u = 0; c < C; c += G, ++u) s4[u] = fmaf(mean[c], w1[(size_t)c * R + j], s4[u]);
        part[cg * R + j] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    }
    __syncthreads();
    if (tid < R) {
        float s = b1[tid];
        for (int g = 0; g < G; ++g) s += part[g * R + tid];
        s1_o[(size_t)b * R + tid] = s;
        rr_o[(size_t)b * R + tid] = s / (1.f + __expf(-s));
    }
}
// phase 2 (grid nch x B, 256 threads = 256 channels): gate pre-activation, ds2 = dgate * g (1-g),
// partial sums over the chunk of ds1_pre[j] = sum_c ds2[c] * w2[j][c].
// The gate pre-activation keeps eight loads of w2 in flight per thread; the chunk's part of W2 ds2 is computed
// with warp w owning the rows j = w, w + 8, ... and its lanes striding over the 256 channels (eight independent
// coalesced loads, one shuffle reduction per row).  The first form -- every warp reducing every row j over its
// 32 channels, one dependent load and five shuffles per row and warp -- took 36 us per launch at C = 2688,
// R = 112 (r3 launch list), as long as a pass over the block's activation tensor.
__global__ void __launch_bounds__(256)
se_bwd_phase2_kernel(const float *__restrict__ rr_g, const float *__restrict__ dgate_partial, int dg_blocks,
                     const float *__restrict__ w2, const float *__restrict__ b2, int C, int R,
                     float *__restrict__ ds2_o, float *__restrict__ ds1p) {
    extern __shared__ float sm[];       // rr[R] (callers allocate 9 R floats)
    __shared__ float sd[256];           // ds2 of the block's channels
    float *rr = sm;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x * 256 + tid;
    for (int j = tid; j < R; j += 256) rr[j] = rr_g[(size_t)b * R + j];
    __syncthreads();
    float d = 0.f;
    if (c < C) {
        float s8[8] = {b2[c], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int j = 0;
        for (; j + 7 < R; j += 8) {
            float q[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) q[u] = w2[(size_t)(j + u) * C + c];
#pragma unroll
            for (int u = 0; u < 8; ++u) s8[u] = fmaf(rr[j + u], q[u], s8[u]);
        }
        for (; j < R; ++j) s8[0] = fmaf(rr[j], w2[(size_t)j * C + c], s8[0]);
        const float sg = ((s8[0] + s8[1]) + (s8[2] + s8[3])) + ((s8[4] + s8[5]) + (s8[6] + s8[7]));
        const float g = 1.f / (1.f + __expf(-sg));
        const float *src = dgate_partial + (size_t)b * dg_blocks * C + c;
        float dg = 0.f;
        for (int k = 0; k < dg_blocks; ++k) dg += src[(size_t)k * C];
        d = dg * g * (1.f - g);
        ds2_o[(size_t)b * C + c] = d;
    }
    sd[tid] = d;
    __syncthreads();
    const int cbase = blockIdx.x * 256;
    float *out = ds1p + ((size_t)b * gridDim.x + blockIdx.x) * R;
    for (int j = warp; j < R; j += 8) {
        const float *wr = w2 + (size_t)j * C + cbase;
        float q[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) q[i] = cbase + lane + 32 * i < C ? wr[lane + 32 * i] : 0.f;
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) v = fmaf(sd[lane + 32 * i], q[i], v);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) out[j] = v;
    }
}
// phase 3 (grid nch x B): ds1 = (sum of chunk partials) * swish'(s1); dmean[c] = sum_j w1[c][j] ds1[j]
__global__ void __launch_bounds__(256)
se_bwd_phase3_kernel(const float *__restrict__ ds1p, int nch, const float *__restrict__ s1_g,
                     const float *__restrict__ w1, int C, int R, float *__restrict__ ds1_o,
                     float *__restrict__ dmean) {
    extern __shared__ float sm[];       // ds1[R]
    float *ds1 = sm;
    const int b = blockIdx.y, tid = threadIdx.x;
    for (int j = tid; j < R; j += 256) {
        float t = 0.f;
        for (int k = 0; k < nch; ++k) t += ds1p[((size_t)b * nch + k) * R + j];
        const float u = s1_g[(size_t)b * R + j], sg = 1.f / (1.f + __expf(-u));
        const float v = t * sg * (1.f + u * (1.f - sg));
        ds1[j] = v;
        if (blockIdx.x == 0) ds1_o[(size_t)b * R + j] = v;
    }
    __syncthreads();
    const int c = blockIdx.x * 256 + tid;
    if (c < C) {
        const float *wr = w1 + (size_t)c * R;
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        int j = 0;
        if ((R & 3) == 0 && (reinterpret_cast<uintptr_t>(w1) & 15) == 0) {
            // a thread streams its own row of W1: four 16-byte loads in flight (r3 launch list: 15 us per launch
            // at C = 2688, R = 112 with four scalar loads in flight)
            for (; j + 15 < R; j += 16) {
                float4 q[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) q[u] = *reinterpret_cast<const float4 *>(wr + j + 4 * u);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    s4[u] = fmaf(q[u].x, ds1[j + 4 * u], s4[u]); s4[u] = fmaf(q[u].y, ds1[j + 4 * u + 1], s4[u]);
                    s4[u] = fmaf(q[u].z, ds1[j + 4 * u + 2], s4[u]); s4[u] = fmaf(q[u].w, ds1[j + 4 * u + 3], s4[u]);
                }
            }
        }
        for (; j + 3 < R; j += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) s4[u] = fmaf(wr[j + u], ds1[j + u], s4[u]);
        }
        for (; j < R; ++j) s4[0] = fmaf(wr[j], ds1[j], s4[0]);
        dmean[(size_t)b * C + c] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    }
}
// phase 4: weight / bias gradients summed over images in image order (deterministic)
//   dW1[c][j] = sum_b mean[b][c] ds1[b][j];  dW2[j][c] = sum_b rr[b][j] ds2[b][c];  db1 = sum_b ds1;  db2 = sum_b ds2
__global__ void __launch_bounds__(256)
se_bwd_phase4_kernel(const float *__restrict__ mean, const float *__restrict__ rr, const float *__restrict__ ds2,
                     const float *__restrict__ ds1, int B, int C, int R, float *__restrict__ dw1,
                     float *__restrict__ dw2, float *__restrict__ db1, float *__restrict__ db2) {
    const size_t cr = (size_t)C * R;
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < cr) {
        const int c = (int)(i / R), j = (int)(i % R);
        float t = 0.f;
        for (int b = 0; b < B; ++b) t = fmaf(mean[(size_t)b * C + c], ds1[(size_t)b * R + j], t);
        dw1[i] = t;
    } else if (i < 2 * cr) {
        const size_t k = i - cr;
        const int j = (int)(k / C), c = (int)(k % C);
        float t = 0.f;
        for (int b = 0; b < B; ++b) t = fmaf(rr[(size_t)b * R + j], ds2[(size_t)b * C + c], t);
        dw2[k] = t;
    } else if (i < 2 * cr + R) {
        const int j = (int)(i - 2 * cr);
        float t = 0.f;
        for (int b = 0; b < B; ++b) t += ds1[(size_t)b * R + j];
        db1[j] = t;
    } else if (i < 2 * cr + R + C) {
        const int c = (int)(i - 2 * cr - R);
        float t = 0.f;
        for (int b = 0; b < B; ++b) t += ds2[(size_t)b * C + c];
        db2[c] = t;
    }
}
// dy = dyg * gate + dmean * inv_hw
template <typename T, int CV>
__global__ void __launch_bounds__(256)
se_bwd_finish_kernel(const T *__restrict__ dyg, const float *__restrict__ gate, const float *__restrict__ dmean,
                     float inv_hw, T *__restrict__ dy, int HW, int C, size_t nvec_total, uint32_t zero) {
    const int nvec = C / CV;
    const size_t stride = (size_t)gridDim.x * 256;
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    for (; i + 3 * stride < nvec_total; i += 4 * stride) {      // four vectors in flight
        uint4 raw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) raw[u] = ld16(dyg + (i + u * stride) * CV);
        tie_loads(raw, zero);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const size_t iv = i + u * stride;
            const int c = (int)(iv % nvec) * CV;
            const size_t b = iv / ((size_t)nvec * HW);
            float v[CV], g[CV], m[CV];
            Unpack16<T, CV>::run(raw[u], v);
            ldv<CV>(gate + b * C + c, g);
            ldv<CV>(dmean + b * C + c, m);
#pragma unroll
            for (int k = 0; k < CV; ++k) v[k] = fmaf(v[k], g[k], m[k] * inv_hw);
            VecB<T, CV>::store(dy + iv * CV, v);
        }
    }
    for (; i < nvec_total; i += stride) {
        const int c = (int)(i % nvec) * CV;
        const size_t b = i / ((size_t)nvec * HW);
        float v[CV], g[CV], m[CV];
        VecB<T, CV>::load(dyg + i * CV, v);
        ldv<CV>(gate + b * C + c, g);
        ldv<CV>(dmean + b * C + c, m);
#pragma unroll
        for (int k = 0; k < CV; ++k) v[k] = fmaf(v[k], g[k], m[k] * inv_hw);
        VecB<T, CV>::store(dy + i * CV, v);
    }
}

// ------------------------------------------------------------------ depthwise backward (general)
// dW[tap][c] = sum_{b,oy,ox} x[b, oy*s - pad + ky, ox*s - pad + kx, c] * dz[b,oy,ox,c]
template <typename T, int CV, int K>
__global__ void dw_wgrad_general_kernel(const T *__restrict__ x, const T *__restrict__ dz, int B, int H, int W,
                                        int Ho, int Wo, int C, int stride, int pad_t, int pad_l,
                                        int pix_per_block, int cvb, float *__restrict__ partial) {
    // block (x, y): pixel range x, channel-vector chunk y of `cvb` vectors (wide layers -- D4..D6 reach 3456
    // channels -- would otherwise need > 1024 threads of ~100 registers)
    extern __shared__ float sred[];      // PY * K*K * cvb*CV
    const int PY = blockDim.x / cvb, CC = cvb * CV;
    const int cv = threadIdx.x % cvb, py = threadIdx.x / cvb;
    const int c = (blockIdx.y * cvb + cv) * CV;
    const bool active = c < C;
    const size_t total = (size_t)B * Ho * Wo;
    const size_t p0 = (size_t)blockIdx.x * pix_per_block;
    const size_t p1 = p0 + pix_per_block < total ? p0 + pix_per_block : total;
    float acc[K * K][CV];
#pragma unroll
    for (int t = 0; t < K * K; ++t)
#pragma unroll
        for (int k = 0; k < CV; ++k) acc[t][k] = 0.f;
    for (size_t p = p0 + py; active && p < p1; p += PY) {
        const int ox = (int)(p % Wo), oy = (int)((p / Wo) % Ho);
        const size_t b = p / ((size_t)Wo * Ho);
        float g[CV];
        VecB<T, CV>::load(dz + p * C + c, g);
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
            const int iy = oy * stride - pad_t + ky;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                const int ix = ox * stride - pad_l + kx;
                if (ix < 0 || ix >= W) continue;
                float v[CV];
                VecB<T, CV>::load(x + ((b * H + iy) * (size_t)W + ix) * C + c, v);
#pragma unroll
                for (int k = 0; k < CV; ++k) acc[ky * K + kx][k] = fmaf(v[k], g[k], acc[ky * K + kx][k]);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < K * K; ++t)
#pragma unroll
        for (int k = 0; k < CV; ++k) sred[((size_t)py * K * K + t) * CC + cv * CV + k] = acc[t][k];
    __syncthreads();
    for (int i = threadIdx.x; i < K * K * CC; i += blockDim.x) {
        const int tap = i / CC, cl = i - tap * CC;
        const int cg = blockIdx.y * CC + cl;
        if (cg >= C) continue;
        float t = 0.f;
        for (int r = 0; r < PY; ++r) t += sred[(size_t)r * K * K * CC + i];
        partial[((size_t)blockIdx.x * K * K + tap) * C + cg] = t;
    }
}
__global__ void sum_partials_warp_kernel(const float *__restrict__ partial, int nblk, int n,
                                         float *__restrict__ out) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    int b = lane;
    for (; b + 7 * 32 < nblk; b += 8 * 32) {      // eight loads in flight, same summation order
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(partial + (size_t)(b + 32 * u) * n + i);
#pragma unroll
        for (int u = 0; u < 8; ++u) s += (double)v[u];
    }
    for (; b < nblk; b += 32) s += (double)__ldcg(partial + (size_t)b * n + i);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) out[i] = (float)s;
}
// data gradient of a depthwise conv (any stride): gather form
template <typename T, int CV, int K>
__global__ void __launch_bounds__(256)
dw_dgrad_kernel(const T *__restrict__ dz, const float *__restrict__ w, T *__restrict__ dx, int B, int H, int W,
                int Ho, int Wo, int C, int stride, int pad_t, int pad_l) {
    const int nvec = C / CV;
    const size_t total = (size_t)B * H * W * nvec;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const int c = (int)(i % nvec) * CV;
        const size_t pix = i / nvec;
        const int ix = (int)(pix % W), iy = (int)((pix / W) % H);
        const size_t b = pix / ((size_t)W * H);
        float acc[CV];
#pragma unroll
        for (int k = 0; k < CV; ++k) acc[k] = 0.f;
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
            const int ny = iy + pad_t - ky;
            if (ny < 0 || ny % stride) continue;
            const int oy = ny / stride;
            if (oy >= Ho) continue;
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                const int nx = ix + pad_l - kx;
                if (nx < 0 || nx % stride) continue;
                const int ox = nx / stride;
                if (ox >= Wo) continue;
                float g[CV], wk[CV];
                VecB<T, CV>::load(dz + ((b * Ho + oy) * (size_t)Wo + ox) * C + c, g);
                ldv<CV>(w + (size_t)(ky * K + kx) * C + c, wk);
#pragma unroll
                for (int k = 0; k < CV; ++k) acc[k] = fmaf(g[k], wk[k], acc[k]);
            }
        }
        VecB<T, CV>::store(dx + pix * C + c, acc);
    }
}

// ------------------------------------------------------------------ stem weight gradient
// dW[ky][kx][ci][co] = sum_{b,oy,ox} img[b, 2oy - pad + ky, 2ox - pad + kx, ci] * dz[b,oy,ox,co]
// block = 27 x C0 threads-worth of outputs over a pixel range; partial[blk][27*C0]
template <typename T>
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const float *__restrict__ img, const T *__restrict__ dz, int B, int H, int W, int Ho, int Wo,
                  int C0, int pad_t, int pad_l, int pix_per_block, float *__restrict__ partial) {
    const int n_out = 27 * C0;
    const size_t total = (size_t)B * Ho * Wo;
    const size_t p0 = (size_t)blockIdx.x * pix_per_block;
    const size_t p1 = p0 + pix_per_block < total ? p0 + pix_per_block : total;
    for (int o = threadIdx.x; o < n_out; o += 256) {
        const int co = o % C0, ci = (o / C0) % 3, tap = o / (3 * C0);
        const int ky = tap / 3, kx = tap % 3;
        float acc = 0.f;
        for (size_t p = p0; p < p1; ++p) {
            const int ox = (int)(p % Wo), oy = (int)((p / Wo) % Ho);
            const size_t b = p / ((size_t)Wo * Ho);
            const int iy = oy * 2 - pad_t + ky, ix = ox * 2 - pad_l + kx;
            if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
            acc = fmaf(img[((b * H + iy) * (size_t)W + ix) * 3 + ci], to_f<T>(dz[p * C0 + co]), acc);
        }
        partial[(size_t)blockIdx.x * n_out + o] = acc;
    }
}

// flipped depthwise kernel (the data gradient of a stride-1 depthwise conv is the same conv with
// the taps reversed) followed by C ones and C zeros (identity scale / shift of the epilogue)
__global__ void dw_flip_kernel(const float *__restrict__ w, float *__restrict__ out, int taps, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < taps * C) out[i] = w[(size_t)(taps - 1 - i / C) * C + i % C];
    else if (i < (taps + 1) * C) out[i] = 1.f;
    else if (i < (taps + 2) * C) out[i] = 0.f;
}

// dwconv_tma.cu
int dwconv_bf16_tma(const void *x, const float *w, const float *scale, const float *shift, void *y, float *se_sum,
                    int B, int H, int W, int C, int k, int stride, int act, cudaStream_t st, float *stats = nullptr);
int dw_wgrad_bf16_splits(int B, int H, int W, int C, int stride);
int dw_wgrad_bf16_tma(const void *x, const void *dz, float *partial, int nsplit, int B, int H, int W, int C, int k,
                      int stride, cudaStream_t st);

// bf16 dz, tiled: block walks 4 x 32 output-pixel tiles (tile = blockIdx.x, += gridDim.x), staging
// the fp32 RGB patch (9 x 65 pixels, coalesced rows) and the dz tile in shared memory; a thread owns
// one output-channel PAIR and TPG of the 27 (tap, ci) rows and keeps their partial sums in
// registers over all its tiles (packed FFMA2), so the kernel is a streaming pass over images + dz.
constexpr int kSwTW = 32, kSwTH = 4, kSwIW = 2 * kSwTW + 1, kSwIH = 2 * kSwTH + 1;
template <int C0>
__global__ void __launch_bounds__(256)
stem_wgrad_tiled_kernel(const float *__restrict__ img, const __nv_bfloat16 *__restrict__ dz, int B, int H, int W,
                        int Ho, int Wo, int pad_t, int pad_l, int tiles_x, int tiles_y,
                        float *__restrict__ partial) {
    constexpr int CPn = C0 / 2, G = 256 / CPn, TPG = (27 + G - 1) / G, ROW = kSwIW * 3, NPIX = kSwTW * kSwTH;
    __shared__ float sin_[kSwIH * ROW];
    __shared__ __align__(16) uint32_t sdz[NPIX * CPn];
    const int cp = threadIdx.x % CPn, g = threadIdx.x / CPn;
    const bool active = g < G;
    int off[TPG];
#pragma unroll
    for (int j = 0; j < TPG; ++j) {
        const int tc = min(g * TPG + j, 26);
        const int tap = tc / 3, ci = tc - tap * 3;
        off[j] = (tap / 3) * ROW + (tap % 3) * 3 + ci;
    }
    float2 acc[TPG];
#pragma unroll
    for (int j = 0; j < TPG; ++j) acc[j] = make_float2(0.f, 0.f);
    const int total_tiles = tiles_x * tiles_y * B;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int r = t;
        const int tx = r % tiles_x; r /= tiles_x;
        const int ty = r % tiles_y; const int b = r / tiles_y;
        const int ty0 = ty * kSwTH, tx0 = tx * kSwTW;
        const int iy0 = ty0 * 2 - pad_t, ix0 = tx0 * 2 - pad_l;
        const float *ib = img + (size_t)b * H * W * 3;
        __syncthreads();                // previous tile fully consumed
        for (int i = threadIdx.x; i < kSwIH * ROW; i += 256) {
            const int rr = i / ROW, cidx = i - rr * ROW;
            const int gy = iy0 + rr, g3 = ix0 * 3 + cidx;
            sin_[i] = (gy >= 0 && gy < H && g3 >= 0 && g3 < W * 3) ? ib[(size_t)gy * W * 3 + g3] : 0.f;
        }
        for (int i = threadIdx.x; i < NPIX * CPn; i += 256) {
            const int pix = i / CPn, c2 = i - pix * CPn;
            const int oy = ty0 + pix / kSwTW, ox = tx0 + pix % kSwTW;
            sdz[i] = (oy < Ho && ox < Wo)
                ? reinterpret_cast<const uint32_t *>(dz + (((size_t)b * Ho + oy) * Wo + ox) * C0)[c2] : 0u;
        }
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int pix = 0; pix < NPIX; ++pix) {
                const uint32_t w2 = sdz[pix * CPn + cp];
                const float2 dzv = make_float2(__uint_as_float(w2 << 16), __uint_as_float(w2 & 0xffff0000u));
                const float *base = sin_ + (pix / kSwTW) * 2 * ROW + (pix % kSwTW) * 6;
#pragma unroll
                for (int j = 0; j < TPG; ++j) {
                    const float v = base[off[j]];
                    acc[j] = __ffma2_rn(make_float2(v, v), dzv, acc[j]);
                }
            }
        }
    }
    if (active) {
#pragma unroll
        for (int j = 0; j < TPG; ++j) {
            const int tc = g * TPG + j;
            if (tc < 27)
                *reinterpret_cast<float2 *>(partial + (size_t)blockIdx.x * 27 * C0 + (size_t)tc * C0 + cp * 2) = acc[j];
        }
    }
}

// kernels that batch four 16-byte loads per thread need ~56 registers: four 256-thread blocks per SM = one wave
static unsigned grid_for_n4(size_t n) {
    unsigned b = cdiv(n, 256 * 4);
    return b > (unsigned)kNumSMs * 4 ? kNumSMs * 4 : (b ? b : 1);
}
static unsigned grid_for_n(size_t n) {
    unsigned b = cdiv(n, 256);
    return b > (unsigned)kNumSMs * 8 ? kNumSMs * 8 : (b ? b : 1);
}

}  // namespace effdet

using namespace effdet;

#define DISPATCH_TB(dtype, EXPR_F32, EXPR_BF16)                                       \
    if ((dtype) == EFFDET_F32) { EXPR_F32; }                                          \
    else if ((dtype) == EFFDET_BF16) { EXPR_BF16; }                                   \
    else return fail(EFFDET_E_INVALID, "%s: bad dtype", __func__);

/* Backward of y = act(BN(z)) for act in {none, relu, swish}.  ua/ub: the affine map u = z*ua + ub the
 * forward applied (training: gamma*invstd, beta - mean*gamma*invstd; frozen BN: folded scale/shift).
 * frozen != 0: inference-mode BN (dz = ua * dy*act'(u); no dgamma/dbeta). */
extern "C" int effdet_bn_act_backward(const void *dy, const void *z, size_t rows, int C, const float *gamma,
                                      const float *save_mean, const float *save_invstd, const float *ua,
                                      const float *ub, int frozen, int act, float *dgamma, float *dbeta,
                                      void *dz, float *k123, float *partial, int nblk, int dtype, void *stream) {
    EFFDET_REQUIRE(dy && z && dz && k123 && ua && ub, "null pointer");
    EFFDET_REQUIRE(rows > 0 && C > 0 && C % 8 == 0 && nblk > 0, "bad sizes");
    cudaStream_t st = as_stream(stream);
    if (frozen) {
        EFFDET_CUDA(cudaMemsetAsync(k123, 0, 3 * (size_t)C * sizeof(float), st));
        EFFDET_CUDA(cudaMemcpyAsync(k123, ua, (size_t)C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else {
        EFFDET_REQUIRE(gamma && save_mean && save_invstd && partial, "null statistics");
        const int nvec = C / 4;
        EFFDET_REQUIRE(nvec <= 1024, "C too large");
        int PY = 256 / nvec; if (PY < 1) PY = 1;
        const int rpb = (int)cdiv(rows, nblk);
        EFFDET_REQUIRE((int)cdiv(rows, rpb) == nblk, "nblk must come from effdet_colreduce_blocks");
        const size_t sm = (size_t)PY * 2 * C * sizeof(float);
#define BN_RED(A)                                                                                              \
        DISPATCH_TB(dtype,                                                                                     \
            (bn_act_bwd_reduce_kernel<float, 4, A><<<nblk, nvec * PY, sm, st>>>(                                \
                (const float *)z, (const float *)dy, ua, ub, save_mean, rows, C, rpb, partial, 0u)),            \
            (bn_act_bwd_reduce_kernel<__nv_bfloat16, 4, A><<<nblk, nvec * PY, sm, st>>>(                        \
                (const __nv_bfloat16 *)z, (const __nv_bfloat16 *)dy, ua, ub, save_mean, rows, C, rpb, partial, 0u)))
        if (act == EFFDET_ACT_SWISH) { BN_RED(EFFDET_ACT_SWISH) }
        else if (act == EFFDET_ACT_RELU) { BN_RED(EFFDET_ACT_RELU) }
        else if (act == EFFDET_ACT_NONE) { BN_RED(EFFDET_ACT_NONE) }
        else return fail(EFFDET_E_UNSUPPORTED, "%s: unsupported activation", __func__);
#undef BN_RED
        EFFDET_LAUNCHED();
        bn_act_bwd_finalize_kernel<<<cdiv((size_t)C * 32, 256), 256, 0, st>>>(partial, nblk, (double)rows, gamma,
                                                                            save_mean, save_invstd, k123, dgamma,
                                                                            dbeta, C);
        EFFDET_LAUNCHED();
    }
    {
        const int nva = C / 4;
        EFFDET_REQUIRE(nva <= 1024, "C too large");
        int PYa = 256 / nva; if (PYa < 1) PYa = 1;
        // one balanced wave (<= 4 blocks per SM: 48 registers x 256 threads), each thread at least kBnU rows
        size_t rpb = cdiv(rows, (size_t)kNumSMs * 4);
        if (rpb < (size_t)PYa * kBnU) rpb = (size_t)PYa * kBnU;
        rpb = cdiv(rpb, PYa) * PYa;
        const unsigned nb = cdiv(rows, rpb);
#define BN_APP(A)                                                                                              \
        DISPATCH_TB(dtype,                                                                                     \
            (bn_act_bwd_apply_kernel<float, 4, A><<<nb, nva * PYa, 0, st>>>(                                    \
                (const float *)dy, (const float *)z, ua, ub, k123, (float *)dz, rows, C, (int)rpb, 0u)),        \
            (bn_act_bwd_apply_kernel<__nv_bfloat16, 4, A><<<nb, nva * PYa, 0, st>>>(                            \
                (const __nv_bfloat16 *)dy, (const __nv_bfloat16 *)z, ua, ub, k123, (__nv_bfloat16 *)dz, rows, C, \
                (int)rpb, 0u)))
        if (act == EFFDET_ACT_SWISH) { BN_APP(EFFDET_ACT_SWISH) }
        else if (act == EFFDET_ACT_RELU) { BN_APP(EFFDET_ACT_RELU) }
        else if (act == EFFDET_ACT_NONE) { BN_APP(EFFDET_ACT_NONE) }
        else return fail(EFFDET_E_UNSUPPORTED, "%s: unsupported activation", __func__);
#undef BN_APP
    }
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

extern "C" int effdet_se_apply(const void *y, const float *gate, void *out, int B, int HW, int C, int dtype,
                               void *stream) {
    EFFDET_REQUIRE(y && gate && out && B > 0 && HW > 0 && C > 0 && C % 8 == 0, "bad arguments");
    cudaStream_t st = as_stream(stream);
    const size_t n = (size_t)B * HW * C;
    DISPATCH_TB(dtype,
        (se_apply_kernel<float, 4><<<grid_for_n(n / 4), 256, 0, st>>>((const float *)y, gate, (float *)out, HW, C, n / 4)),
        (se_apply_kernel<__nv_bfloat16, 8><<<grid_for_n(n / 8), 256, 0, st>>>(
            (const __nv_bfloat16 *)y, gate, (__nv_bfloat16 *)out, HW, C, n / 8)))
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

/* Stochastic depth (FixedDropout with noise shape (None,1,1,1): efficientnet.py:147-188, applied to the
 * projected branch of every skip block when drop_rate > 0, efficientnet.py:300-304).  Keras dropout in the
 * training phase: keep = uniform[0,1) >= rate per (block, image); the kept branch is scaled by 1/(1-rate).
 * scales (nblocks, B) f32 holds keep/(1-rate).  The uniform numbers come from a counter-based generator
 * (splitmix64 of seed, step, block, image -- TensorFlow's own stream cannot be reproduced, only its
 * distribution); *step_counter is read and incremented on the device, so a captured graph draws a fresh
 * mask at every replay. */
__global__ void __launch_bounds__(256)
drop_connect_scales_kernel(const float *__restrict__ rates, int nblocks, int B, unsigned long long seed,
                           unsigned long long *__restrict__ step_counter, float *__restrict__ scales) {
    const unsigned long long step = *step_counter;
    __syncthreads();
    for (int i = threadIdx.x; i < nblocks * B; i += blockDim.x) {
        const int blk = i / B;
        unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (step * (unsigned long long)(nblocks * B) + i + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        const float u = (float)(z >> 40) * (1.0f / 16777216.0f);      // 24 random bits -> [0, 1)
        const float rate = rates[blk];
        scales[i] = u >= rate ? 1.0f / (1.0f - rate) : 0.0f;
    }
    if (threadIdx.x == 0) *step_counter = step + 1;
}
extern "C" int effdet_drop_connect_scales(const float *rates, int nblocks, int B, unsigned long long seed,
                                          unsigned long long *step_counter, float *scales, void *stream) {
    EFFDET_REQUIRE(rates && step_counter && scales && nblocks > 0 && B > 0, "bad arguments");
    drop_connect_scales_kernel<<<1, 256, 0, as_stream(stream)>>>(rates, nblocks, B, seed, step_counter, scales);
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

// out = y * scale[b] (+ res): forward of the dropped skip branch (res = block input) and, with res == NULL,
// its backward (d y = d out * scale[b])
template <typename T, int CV>
__global__ void __launch_bounds__(256)
sample_scale_add_kernel(const T *__restrict__ y, const float *__restrict__ scale, const T *__restrict__ res,
                        T *__restrict__ out, size_t vec_per_image, size_t nvec_total) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvec_total; i += (size_t)gridDim.x * 256) {
        const float s = scale[i / vec_per_image];
        float v[CV];
        VecB<T, CV>::load(y + i * CV, v);
        if (res) {
            float r[CV];
            VecB<T, CV>::load(res + i * CV, r);
#pragma unroll
            for (int k = 0; k < CV; ++k) v[k] = v[k] * s + r[k];
        } else {
#pragma unroll
            for (int k = 0; k < CV; ++k) v[k] *= s;
        }
        VecB<T, CV>::store(out + i * CV, v);
    }
}
extern "C" int effdet_sample_scale_add(const void *y, const float *scale, const void *res, void *out, int B,
                                       size_t per_image, int dtype, void *stream) {
    EFFDET_REQUIRE(y && scale && out && B > 0 && per_image > 0 && per_image % 8 == 0, "bad arguments");
    cudaStream_t st = as_stream(stream);
    const size_t n = (size_t)B * per_image;
    DISPATCH_TB(dtype,
        (sample_scale_add_kernel<float, 4><<<grid_for_n(n / 4), 256, 0, st>>>(
            (const float *)y, scale, (const float *)res, (float *)out, per_image / 4, n / 4)),
        (sample_scale_add_kernel<__nv_bfloat16, 8><<<grid_for_n(n / 8), 256, 0, st>>>(
            (const __nv_bfloat16 *)y, scale, (const __nv_bfloat16 *)res, (__nv_bfloat16 *)out, per_image / 8, n / 8)))
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

/* Per-(image, block, channel) spatial sums of y (B,HW,C): the SE squeeze when the depthwise
 * output is produced by a separate BN/activation pass (training mode).  partial (B, nblk, C) with
 * nblk = effdet_se_backward_blocks(). */
namespace effdet {      // dwconv.cu
cudaError_t se_cluster_launch(cudaStream_t st, const float *se_sum, int se_blocks, float inv_hw, const float *w1,
                              const float *b1, const float *w2, const float *b2, float *gate, int B, int C, int R,
                              float *mean_o, float *s1_o, float *rr_o);
bool se_cluster_ok(int C, int R);
}
extern "C" int effdet_se_backward_blocks(int HW, int C, int dtype);
extern "C" int effdet_spatial_sum(const void *y, float *partial, int nblk, int B, int HW, int C, int dtype,
                                  void *stream) {
    EFFDET_REQUIRE(y && partial && B > 0 && HW > 0 && C > 0 && C % 8 == 0, "bad arguments");
    EFFDET_REQUIRE(nblk == effdet_se_backward_blocks(HW, C, dtype), "nblk must come from effdet_se_backward_blocks");
    cudaStream_t st = as_stream(stream);
    const int CV = dtype == EFFDET_BF16 ? 8 : 4;
    const int nvec = C / CV;
    EFFDET_REQUIRE(nvec <= 1024, "C too large");
    int PY = 256 / nvec; if (PY < 1) PY = 1;
    const int rpb = (int)cdiv(HW, nblk);
    const size_t sm = (size_t)PY * C * sizeof(float);
    dim3 grid(nblk, B);
    DISPATCH_TB(dtype,
        (se_bwd_reduce_kernel<float, 4><<<grid, nvec * PY, sm, st>>>((const float *)y, nullptr, HW, C, rpb, partial, 0u)),
        (se_bwd_reduce_kernel<__nv_bfloat16, 8><<<grid, nvec * PY, sm, st>>>((const __nv_bfloat16 *)y, nullptr, HW, C,
                                                                              rpb, partial, 0u)))
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

extern "C" int effdet_se_backward_blocks(int HW, int C, int dtype) {
    const int CV = dtype == EFFDET_BF16 ? 8 : 4;
    int nvec = C / CV; if (nvec < 1) nvec = 1;
    int PY = 256 / nvec; if (PY < 1) PY = 1;
    int rpb = PY * 16;
    while (rpb > PY && cdiv(HW, rpb) < 8) rpb >>= 1;
    return (int)cdiv(HW, rpb);
}

/* Squeeze-excite backward (efficientnet.py:255-286).  dyg: gradient of the gated tensor y*gate.
 * Outputs: dy (B,HW,C) gradient of y (both the direct path and the path through the squeeze),
 * gradients of se_reduce / se_expand kernels and biases.  Scratch sizes (floats):
 * dg_partial B*effdet_se_backward_blocks()*C ; fc_scratch B*(2*C*R + R + C) ; dmean B*C. */
extern "C" int effdet_se_backward(const void *dyg, const void *y, const float *gate, const float *se_sum,
                                  int se_blocks, const float *w1, const float *b1, const float *w2,
                                  const float *b2, void *dy, float *dw1, float *db1, float *dw2, float *db2,
                                  float *dg_partial, int dg_blocks, float *fc_scratch, float *dmean, int B, int HW,
                                  int C, int R, int dtype, void *stream) {
    EFFDET_REQUIRE(dyg && y && gate && se_sum && w1 && b1 && w2 && b2 && dy && dw1 && db1 && dw2 && db2 &&
                       dg_partial && fc_scratch && dmean, "null pointer");
    EFFDET_REQUIRE(B > 0 && HW > 0 && C > 0 && C % 8 == 0 && R > 0 && se_blocks > 0, "bad sizes");
    EFFDET_REQUIRE(dg_blocks == effdet_se_backward_blocks(HW, C, dtype), "dg_blocks must come from effdet_se_backward_blocks");
    cudaStream_t st = as_stream(stream);
    const int CV = dtype == EFFDET_BF16 ? 8 : 4;
    const int nvec = C / CV;
    EFFDET_REQUIRE(nvec <= 1024, "C too large");
    int PY = 256 / nvec; if (PY < 1) PY = 1;
    const int rpb = (int)cdiv(HW, dg_blocks);
    const size_t sm = (size_t)PY * C * sizeof(float);
    dim3 grid(dg_blocks, B);
    DISPATCH_TB(dtype,
        (se_bwd_reduce_kernel<float, 4><<<grid, nvec * PY, sm, st>>>((const float *)dyg, (const float *)y, HW, C, rpb, dg_partial, 0u)),
        (se_bwd_reduce_kernel<__nv_bfloat16, 8><<<grid, nvec * PY, sm, st>>>(
            (const __nv_bfloat16 *)dyg, (const __nv_bfloat16 *)y, HW, C, rpb, dg_partial, 0u)))
    EFFDET_LAUNCHED();
    {
        EFFDET_REQUIRE(R <= 512, "R too large");
        const int nch = (C + 255) / 256;
        const size_t per = (size_t)2 * C * R + R + C;
        EFFDET_REQUIRE((size_t)B * (2 * C + 3 * R + (size_t)nch * R) <= (size_t)B * per, "fc_scratch too small");
        float *mean_g = fc_scratch, *s1_g = mean_g + (size_t)B * C, *rr_g = s1_g + (size_t)B * R;
        float *ds2_g = rr_g + (size_t)B * R, *ds1_g = ds2_g + (size_t)B * C, *ds1p = ds1_g + (size_t)B * R;
        const size_t sm1 = (size_t)(C + 512) * sizeof(float);
        EFFDET_REQUIRE(sm1 <= 48 * 1024, "C too large");
        if (se_cluster_ok(C, R)) {
            // forward recompute (squeeze mean, FC1) on thread-block clusters: the one-block-per-image kernel
            // was a 45-us chain of dependent L2 round trips per MBConv block at batch 8
            cudaError_t ce = se_cluster_launch(st, se_sum, se_blocks, 1.f / (float)HW, w1, b1, w2, b2, nullptr, B, C, R,
                                               mean_g, s1_g, rr_g);
            EFFDET_CUDA(ce);
        } else {
            se_bwd_phase1_kernel<<<B, 512, sm1, st>>>(se_sum, se_blocks, 1.f / (float)HW, w1, b1, C, R, mean_g, s1_g, rr_g);
        }
        EFFDET_LAUNCHED();
        dim3 g2(nch, B);
        se_bwd_phase2_kernel<<<g2, 256, (size_t)9 * R * sizeof(float), st>>>(rr_g, dg_partial, dg_blocks, w2, b2, C, R,
                                                                             ds2_g, ds1p);
        EFFDET_LAUNCHED();
        se_bwd_phase3_kernel<<<g2, 256, (size_t)R * sizeof(float), st>>>(ds1p, nch, s1_g, w1, C, R, ds1_g, dmean);
        EFFDET_LAUNCHED();
        se_bwd_phase4_kernel<<<cdiv(per, 256), 256, 0, st>>>(mean_g, rr_g, ds2_g, ds1_g, B, C, R, dw1, dw2, db1, db2);
        EFFDET_LAUNCHED();
    }
    const size_t n = (size_t)B * HW * C;
    DISPATCH_TB(dtype,
        (se_bwd_finish_kernel<float, 4><<<grid_for_n4(n / 4), 256, 0, st>>>(
            (const float *)dyg, gate, dmean, 1.f / (float)HW, (float *)dy, HW, C, n / 4, 0u)),
        (se_bwd_finish_kernel<__nv_bfloat16, 8><<<grid_for_n4(n / 8), 256, 0, st>>>(
            (const __nv_bfloat16 *)dyg, gate, dmean, 1.f / (float)HW, (__nv_bfloat16 *)dy, HW, C, n / 8, 0u)))
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

extern "C" int effdet_dw_backward_blocks(int B, int H, int W, int C, int k, int stride, int dtype) {
    if (dtype == EFFDET_BF16) return dw_wgrad_bf16_splits(B, H, W, C, stride);
    const int CV = dtype == EFFDET_BF16 ? 8 : 4;
    int nvec = C / CV; if (nvec < 1) nvec = 1;
    if (nvec > 128) nvec = 128;          // channel-vector chunk per block (see dw_wgrad_general_kernel)
    int PY = (k == 5 ? 64 : 128) / nvec; if (PY < 1) PY = 1;
    const size_t total = (size_t)B * ((H + stride - 1) / stride) * ((W + stride - 1) / stride);
    size_t ppb = (size_t)PY * 32;
    while (ppb > (size_t)PY && cdiv(total, ppb) < (unsigned)kNumSMs * 2) ppb >>= 1;
    return (int)cdiv(total, ppb);
}

/* Depthwise conv backward for k in {3,5}, stride in {1,2} (efficientnet.py:242-252):
 * dkernel (k,k,C) f32 and dx (B,H,W,C).  partial: k*k*C*effdet_dw_backward_blocks() floats. */
extern "C" int effdet_dw_backward(const void *x, const void *dz, const float *kernel, void *dx, float *dkernel,
                                  float *partial, int nblk, int B, int H, int W, int C, int k, int stride,
                                  int dtype, void *stream) {
    EFFDET_REQUIRE(x && dz && kernel && dkernel && partial, "null pointer");
    EFFDET_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad sizes");
    EFFDET_REQUIRE((k == 3 || k == 5) && (stride == 1 || stride == 2), "k in {3,5}, stride in {1,2}");
    EFFDET_REQUIRE(nblk == effdet_dw_backward_blocks(B, H, W, C, k, stride, dtype), "nblk must come from effdet_dw_backward_blocks");
    cudaStream_t st = as_stream(stream);
    const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
    const int pt = max((Ho - 1) * stride + k - H, 0) / 2, pl = max((Wo - 1) * stride + k - W, 0) / 2;
    if (dtype == EFFDET_BF16) {
        // TMA-tiled weight gradient (dwconv_tma.cu), partial rows summed in a fixed order
        int rc = dw_wgrad_bf16_tma(x, dz, partial, nblk, B, H, W, C, k, stride, st);
        if (rc) return rc;
        sum_partials_warp_kernel<<<cdiv((size_t)k * k * C * 32, 256), 256, 0, st>>>(partial, nblk, k * k * C, dkernel);
        EFFDET_LAUNCHED();
        if (dx && stride == 1) {
            // data gradient == the forward kernel on dz with reversed taps (scratch: partial is free now)
            dw_flip_kernel<<<cdiv((size_t)(k * k + 2) * C, 256), 256, 0, st>>>(kernel, partial, k * k, C);
            EFFDET_LAUNCHED();
            return dwconv_bf16_tma(dz, partial, partial + (size_t)k * k * C, partial + (size_t)(k * k + 1) * C, dx,
                                   nullptr, B, H, W, C, k, 1, EFFDET_ACT_NONE, st);
        }
        if (dx) {
            const size_t n = (size_t)B * H * W * C;
            if (k == 3) dw_dgrad_kernel<__nv_bfloat16, 8, 3><<<grid_for_n(n / 8), 256, 0, st>>>(
                    (const __nv_bfloat16 *)dz, kernel, (__nv_bfloat16 *)dx, B, H, W, Ho, Wo, C, stride, pt, pl);
            else dw_dgrad_kernel<__nv_bfloat16, 8, 5><<<grid_for_n(n / 8), 256, 0, st>>>(
                    (const __nv_bfloat16 *)dz, kernel, (__nv_bfloat16 *)dx, B, H, W, Ho, Wo, C, stride, pt, pl);
            EFFDET_LAUNCHED();
        }
        return EFFDET_OK;
    }
    const int CV = dtype == EFFDET_BF16 ? 8 : 4;
    const int nvec = C / CV;
    const int cvb = nvec > 128 ? 128 : nvec;                 // channel vectors per block
    const int cchunks = (nvec + cvb - 1) / cvb;
    int PY = (k == 5 ? 64 : 128) / cvb; if (PY < 1) PY = 1;
    const size_t total = (size_t)B * Ho * Wo;
    const int ppb = (int)cdiv(total, nblk);
    const size_t sm = (size_t)PY * k * k * cvb * CV * sizeof(float);
#define DWG(T, CVV, KK)                                                                                        \
    {                                                                                                          \
        auto kern = dw_wgrad_general_kernel<T, CVV, KK>;                                                       \
        if (sm > 48 * 1024) EFFDET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        kern<<<dim3(nblk, cchunks), cvb * PY, sm, st>>>((const T *)x, (const T *)dz, B, H, W, Ho, Wo, C, stride, pt, pl, ppb, cvb, partial); \
    }
    if (dtype == EFFDET_F32) { if (k == 3) DWG(float, 4, 3) else DWG(float, 4, 5) }
    else if (dtype == EFFDET_BF16) { if (k == 3) DWG(__nv_bfloat16, 8, 3) else DWG(__nv_bfloat16, 8, 5) }
    else return fail(EFFDET_E_INVALID, "effdet_dw_backward: bad dtype%s", "");
#undef DWG
    EFFDET_LAUNCHED();
    sum_partials_warp_kernel<<<cdiv((size_t)k * k * C * 32, 256), 256, 0, st>>>(partial, nblk, k * k * C, dkernel);
    EFFDET_LAUNCHED();
    if (dx) {
        const size_t n = (size_t)B * H * W * C;
#define DWD(T, CVV, KK) dw_dgrad_kernel<T, CVV, KK><<<grid_for_n(n / CVV), 256, 0, st>>>( \
            (const T *)dz, kernel, (T *)dx, B, H, W, Ho, Wo, C, stride, pt, pl)
        if (dtype == EFFDET_F32) { if (k == 3) DWD(float, 4, 3); else DWD(float, 4, 5); }
        else { if (k == 3) DWD(__nv_bfloat16, 8, 3); else DWD(__nv_bfloat16, 8, 5); }
#undef DWD
        EFFDET_LAUNCHED();
    }
    return EFFDET_OK;
}

extern "C" int effdet_stem_wgrad_blocks(int B, int H, int W) {
    const size_t total = (size_t)B * ((H + 1) / 2) * ((W + 1) / 2);
    size_t ppb = 2048;
    while (ppb > 64 && cdiv(total, ppb) < (unsigned)kNumSMs * 4) ppb >>= 1;
    return (int)cdiv(total, ppb);
}

/* Stem conv (3x3, stride 2, 3 -> C0) weight gradient; dz is the gradient of the conv output
 * (B, H/2, W/2, C0).  partial: 27*C0*effdet_stem_wgrad_blocks() floats. */
extern "C" int effdet_stem_wgrad(const float *images, const void *dz, float *dkernel, float *partial, int nblk,
                                 int B, int H, int W, int C0, int dtype, void *stream) {
    EFFDET_REQUIRE(images && dz && dkernel && partial && B > 0 && H > 0 && W > 0 && C0 > 0, "bad arguments");
    EFFDET_REQUIRE(nblk == effdet_stem_wgrad_blocks(B, H, W), "nblk must come from effdet_stem_wgrad_blocks");
    cudaStream_t st = as_stream(stream);
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    const int pt = max((Ho - 1) * 2 + 3 - H, 0) / 2, pl = max((Wo - 1) * 2 + 3 - W, 0) / 2;
    const size_t total = (size_t)B * Ho * Wo;
    const int ppb = (int)cdiv(total, nblk);
    if (dtype == EFFDET_BF16 && (C0 == 32 || C0 == 40 || C0 == 48 || C0 == 56 || C0 == 64) &&
        (reinterpret_cast<uintptr_t>(dz) & 3) == 0) {
        const int tx = (Wo + kSwTW - 1) / kSwTW, ty = (Ho + kSwTH - 1) / kSwTH;
        int grid = nblk < kNumSMs * 4 ? nblk : kNumSMs * 4;
#define SWG(C) case C: stem_wgrad_tiled_kernel<C><<<grid, 256, 0, st>>>(images, (const __nv_bfloat16 *)dz, B, H, W, Ho, Wo, pt, pl, tx, ty, partial); break;
        switch (C0) { SWG(32) SWG(40) SWG(48) SWG(56) SWG(64) }
#undef SWG
        EFFDET_LAUNCHED();
        sum_partials_warp_kernel<<<cdiv((size_t)27 * C0 * 32, 256), 256, 0, st>>>(partial, grid, 27 * C0, dkernel);
        EFFDET_LAUNCHED();
        return EFFDET_OK;
    }
    DISPATCH_TB(dtype,
        (stem_wgrad_kernel<float><<<nblk, 256, 0, st>>>(images, (const float *)dz, B, H, W, Ho, Wo, C0, pt, pl, ppb, partial)),
        (stem_wgrad_kernel<__nv_bfloat16><<<nblk, 256, 0, st>>>(images, (const __nv_bfloat16 *)dz, B, H, W, Ho, Wo, C0,
                                                                  pt, pl, ppb, partial)))
    EFFDET_LAUNCHED();
    sum_partials_warp_kernel<<<cdiv((size_t)27 * C0 * 32, 256), 256, 0, st>>>(partial, nblk, 27 * C0, dkernel);
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

// ------------------------------------------------------------------ fused squeeze-excite + BN/swish backward
// MBConv backward between the project data gradient and the depthwise backward (efficientnet.py:242-286):
//   y = swish(u), u = a z + b  (BatchNorm in training mode),  yg = y * gate[b][c],  gate = SE(mean_hw y)
//   given dyg:  dy = dyg * gate + dmean / HW,   dz = BN^T( dy * swish'(u) )
// The separate entry points (effdet_se_backward, then effdet_bn_act_backward) read dyg + y, read dyg and
// write dy, read dy + z, read dy + z and write dz: NINE passes over (B, HW, Cmid) tensors.  Every sum the
// two reductions need is linear in per-(image, channel) constants:
//   dgate            = sum dyg y
//   sum dy s'        = gate * sum dyg s'        + (dmean/HW) * sum s'
//   sum dy s' (z-mu) = gate * sum dyg s' (z-mu) + (dmean/HW) * sum s' (z-mu)
// so ONE reduction pass over (dyg, z) produces five per-(image, block, channel) partial sums (y and s' are
// recomputed from z), the tiny SE fully-connected backward and a per-(image, channel) combine follow, and ONE
// apply pass reads (dyg, z) and writes dz: five passes; dy is never materialised.
namespace effdet {

template <typename T> __device__ __forceinline__ void swish_and_grad(float u, float &y, float &g);
template <> __device__ __forceinline__ void swish_and_grad<float>(float u, float &y, float &g) {
    const float s = 1.f / (1.f + __expf(-u));
    y = u * s;
    g = s * (1.f + u * (1.f - s));
}
template <> __device__ __forceinline__ void swish_and_grad<__nv_bfloat16>(float u, float &y, float &g) {
    const float s = fmaf(0.5f, tanh_fast(0.5f * u), 0.5f);
    y = u * s;
    g = s * fmaf(u, 1.f - s, 1.f);
}

// dgp (B, nblk, C): sum dyg*y ; bnp (B, nblk, 4, C): sum dyg s', sum dyg s' z, sum s', sum s' z
// Blocks of <= 256 threads (85 registers: 20 accumulators + 8 constants + 16 words of loads in flight do not fit
// the 64 of a 1024-thread block): wide layers are cut into channel chunks along grid.z (nvb vectors per block).
template <typename T>
__global__ void __launch_bounds__(256, 3)
se_bn_bwd_reduce_kernel(const T *__restrict__ dyg, const T *__restrict__ z, const float *__restrict__ ua,
                        const float *__restrict__ ub, int HW, int C, int nvb,
                        int rows_per_block, float *__restrict__ dgp, float *__restrict__ bnp, uint32_t zero) {
    constexpr int CV = 4;
    extern __shared__ float sred[];
    const int PY = blockDim.x / nvb;
    const int cv = threadIdx.x % nvb, py = threadIdx.x / nvb, c = (blockIdx.z * nvb + cv) * CV;
    const int cl = cv * CV, CL = nvb * CV;            // channel index / count inside the block's chunk
    const bool live = py < PY && c < C;
    const int b = blockIdx.y;
    const size_t r0 = (size_t)blockIdx.x * rows_per_block;
    const size_t r1 = r0 + rows_per_block < (size_t)HW ? r0 + rows_per_block : (size_t)HW;
    const T *pg = dyg + (size_t)b * HW * C, *pz = z + (size_t)b * HW * C;
    // (the z-weighted sums are taken about zero here; the combine kernel shifts them by the batch mean: 64
    // registers leave no room for a fourth per-channel constant next to the 20 accumulators)
    float s[5][CV], a[CV], bb[CV];
#pragma unroll
    for (int k = 0; k < CV; ++k) { a[k] = 0.f; bb[k] = 0.f; }
    if (live) { ldv<CV>(ua + c, a); ldv<CV>(ub + c, bb); }
#pragma unroll
    for (int j = 0; j < 5; ++j)
#pragma unroll
        for (int k = 0; k < CV; ++k) s[j][k] = 0.f;
    auto accumulate = [&](const float *zz, const float *g) {
#pragma unroll
        for (int k = 0; k < CV; ++k) {
            float y, ag;
            swish_and_grad<T>(fmaf(zz[k], a[k], bb[k]), y, ag);
            const float ga = g[k] * ag;
            s[0][k] = fmaf(g[k], y, s[0][k]);
            s[1][k] += ga;
            s[2][k] = fmaf(ga, zz[k], s[2][k]);
            s[3][k] += ag;
            s[4][k] = fmaf(ag, zz[k], s[4][k]);
        }
    };
    size_t r = r0 + py;
    const size_t step = (size_t)kBnU * PY;
    const bool pf = threadIdx.x == 0 && blockIdx.z == 0;      // whole rows: one chunk asks for them
    if (live) {
        if (pf) {
            prefetch_rows(pz, r0, r0 + kBnPD * step, r1, C);
            prefetch_rows(pg, r0, r0 + kBnPD * step, r1, C);
        }
        size_t rq = r0 + kBnPD * step;             // first row not yet requested
        for (; r + (size_t)(kBnU - 1) * PY < r1; r += step, rq += step) {
            if (pf) {
                prefetch_rows(pz, rq, rq + step, r1, C);
                prefetch_rows(pg, rq, rq + step, r1, C);
            }
            typename Raw4<T>::type zr[kBnU], gr[kBnU];
#pragma unroll
            for (int u = 0; u < kBnU; ++u) {
                zr[u] = Raw4<T>::load(pz + (r + (size_t)u * PY) * C + c);
                gr[u] = Raw4<T>::load(pg + (r + (size_t)u * PY) * C + c);
            }
            Raw4<T>::all_loaded(zr, gr, zero);
#pragma unroll
            for (int u = 0; u < kBnU; ++u) {
                float zz[CV], g[CV];
                Raw4<T>::unpack(zr[u], zz);
                Raw4<T>::unpack(gr[u], g);
                accumulate(zz, g);
            }
        }
        for (; r < r1; r += PY) {
            float zz[CV], g[CV];
            VecB<T, CV>::load(pz + r * C + c, zz);
            VecB<T, CV>::load(pg + r * C + c, g);
            accumulate(zz, g);
        }
    }
    const size_t row = (size_t)b * gridDim.x + blockIdx.x;
    if (PY == 1) {                                  // one thread per channel vector: its sums are the block's
        if (live) {
#pragma unroll
            for (int k = 0; k < CV; ++k) {
                dgp[row * C + c + k] = s[0][k];
#pragma unroll
                for (int j = 0; j < 4; ++j) bnp[(row * 4 + j) * C + c + k] = s[j + 1][k];
            }
        }
        return;
    }
    if (py < PY) {
#pragma unroll
        for (int j = 0; j < 5; ++j)
#pragma unroll
            for (int k = 0; k < CV; ++k) sred[((size_t)py * 5 + j) * CL + cl + k] = s[j][k];
    }
    __syncthreads();
    const int c_base = blockIdx.z * CL;
    for (int i = threadIdx.x; i < 5 * CL; i += blockDim.x) {
        float t = 0.f;
        for (int q = 0; q < PY; ++q) t += sred[(size_t)q * 5 * CL + i];      // fixed order
        const int j = i / CL, cc = c_base + (i - j * CL);
        if (cc >= C) continue;
        if (j == 0) dgp[row * C + cc] = t;
        else bnp[(row * 4 + (j - 1)) * C + cc] = t;
    }
}

// Sums of the block partials per (image, channel): block = 32 channels x 16 slices of the block index; a slice adds
// its blocks k = slice, slice + 16, ... in order, the slices are added in order (deterministic).  One thread per
// channel walking all nblk rows (as the SE phase-2 kernel does for dgate) is a chain of nblk dependent L2 round
// trips: 60 us for the 107 row blocks of a 256 x 256 image.  tot_dg (B, C), tot_bn (B, 4, C).
__global__ void __launch_bounds__(512)
se_bn_partial_sum_kernel(const float *__restrict__ dgp, const float *__restrict__ bnp, int nblk, int C,
                         float *__restrict__ tot_dg, float *__restrict__ tot_bn) {
    __shared__ float red[16][5][32];
    const int cx = threadIdx.x & 31, ky = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx, b = blockIdx.y;
    float t[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (c < C) {
        const float *pd = dgp + (size_t)b * nblk * C + c;
        const float *pb = bnp + (size_t)b * nblk * 4 * C + c;
#pragma unroll 2
        for (int k = ky; k < nblk; k += 16) {
            const float v0 = __ldcg(pd + (size_t)k * C);
            const float v1 = __ldcg(pb + ((size_t)k * 4 + 0) * C), v2 = __ldcg(pb + ((size_t)k * 4 + 1) * C);
            const float v3 = __ldcg(pb + ((size_t)k * 4 + 2) * C), v4 = __ldcg(pb + ((size_t)k * 4 + 3) * C);
            t[0] += v0; t[1] += v1; t[2] += v2; t[3] += v3; t[4] += v4;
        }
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) red[ky][j][cx] = t[j];
    __syncthreads();
    if (ky < 5 && c < C) {
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) a += red[q][ky][cx];
        if (ky == 0) tot_dg[(size_t)b * C + c] = a;
        else tot_bn[((size_t)b * 4 + (ky - 1)) * C + c] = a;
    }
}

// per (image, channel): block partials summed in order, folded with gate and dmean/HW into the two sums of the
// BatchNorm backward: out (B, 2, C) = the `partial` matrix bn_act_bwd_finalize_kernel reduces (one row per image)
__global__ void __launch_bounds__(256)
se_bn_bwd_combine_kernel(const float *__restrict__ bnp, int nblk, const float *__restrict__ gate,
                         const float *__restrict__ dmean, const float *__restrict__ mean, float inv_hw, int C,
                         float *__restrict__ out) {
    const int c = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
    if (c >= C) return;
    double t[4] = {0.0, 0.0, 0.0, 0.0};
    const float *p = bnp + (size_t)b * nblk * 4 * C + c;
    for (int k = 0; k < nblk; ++k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] += (double)__ldcg(p + ((size_t)k * 4 + j) * C);
    }
    // sums about the batch mean (what bn_act_bwd_finalize_kernel expects), in double: sum w z - mu sum w
    const double mu = (double)mean[c];
    const double t1 = t[1] - mu * t[0], t3 = t[3] - mu * t[2];
    const double g = (double)gate[(size_t)b * C + c], m = (double)dmean[(size_t)b * C + c] * (double)inv_hw;
    out[((size_t)b * 2 + 0) * C + c] = (float)(g * t[0] + m * t[2]);
    out[((size_t)b * 2 + 1) * C + c] = (float)(g * t1 + m * t3);
}

// dz = k1 * ((dyg * gate + dmean/HW) * swish'(u)) + k2 * z + k3
template <typename T>
__global__ void __launch_bounds__(256, 3)
se_bn_bwd_apply_kernel(const T *__restrict__ dyg, const T *__restrict__ z, const float *__restrict__ ua,
                       const float *__restrict__ ub, const float *__restrict__ k123,
                       const float *__restrict__ gate, const float *__restrict__ dmean, float inv_hw,
                       T *__restrict__ dz, int HW, int C, int nvb, int rows_per_block, uint32_t zero) {
    constexpr int CV = 4;
    const int PY = blockDim.x / nvb;
    const int cv = threadIdx.x % nvb, py = threadIdx.x / nvb, c = (blockIdx.z * nvb + cv) * CV;
    if (py >= PY || c >= C) return;
    const int b = blockIdx.y;
    float a[CV], bb[CV], k1g[CV], k1m[CV], k2[CV], k3[CV];      // k1 * gate, k1 * dmean / HW
    ldv<CV>(ua + c, a); ldv<CV>(ub + c, bb);
    ldv<CV>(k123 + c, k2);
    ldv<CV>(gate + (size_t)b * C + c, k1g); ldv<CV>(dmean + (size_t)b * C + c, k1m);
#pragma unroll
    for (int k = 0; k < CV; ++k) { k1g[k] *= k2[k]; k1m[k] *= k2[k] * inv_hw; }
    ldv<CV>(k123 + C + c, k2); ldv<CV>(k123 + 2 * C + c, k3);
    const size_t r0 = (size_t)blockIdx.x * rows_per_block;
    const size_t r1 = r0 + rows_per_block < (size_t)HW ? r0 + rows_per_block : (size_t)HW;
    const T *pg = dyg + (size_t)b * HW * C, *pz = z + (size_t)b * HW * C;
    T *po = dz + (size_t)b * HW * C;
    auto apply = [&](float *gg, const float *zz) {
#pragma unroll
        for (int k = 0; k < CV; ++k) {
            float y, ag;
            swish_and_grad<T>(fmaf(zz[k], a[k], bb[k]), y, ag);
            gg[k] = fmaf(fmaf(gg[k], k1g[k], k1m[k]), ag, fmaf(k2[k], zz[k], k3[k]));
        }
    };
    size_t r = r0 + py;
    const size_t step = (size_t)kBnU * PY;
    const bool pf = threadIdx.x == 0 && blockIdx.z == 0;
    if (pf) {
        prefetch_rows(pz, r0, r0 + kBnPD * step, r1, C);
        prefetch_rows(pg, r0, r0 + kBnPD * step, r1, C);
    }
    size_t rq = r0 + kBnPD * step;
    for (; r + (size_t)(kBnU - 1) * PY < r1; r += step, rq += step) {
        if (pf) {
            prefetch_rows(pz, rq, rq + step, r1, C);
            prefetch_rows(pg, rq, rq + step, r1, C);
        }
        typename Raw4<T>::type zr[kBnU], gr[kBnU];
#pragma unroll
        for (int u = 0; u < kBnU; ++u) {
            gr[u] = Raw4<T>::load(pg + (r + (size_t)u * PY) * C + c);
            zr[u] = Raw4<T>::load(pz + (r + (size_t)u * PY) * C + c);
        }
        Raw4<T>::all_loaded(zr, gr, zero);
#pragma unroll
        for (int u = 0; u < kBnU; ++u) {
            float gg[CV], zz[CV];
            Raw4<T>::unpack(gr[u], gg);
            Raw4<T>::unpack(zr[u], zz);
            apply(gg, zz);
            VecB<T, CV>::store(po + (r + (size_t)u * PY) * C + c, gg);
        }
    }
    for (; r < r1; r += PY) {
        float gg[CV], zz[CV];
        VecB<T, CV>::load(pg + r * C + c, gg);
        VecB<T, CV>::load(pz + r * C + c, zz);
        apply(gg, zz);
        VecB<T, CV>::store(po + r * C + c, gg);
    }
}

}  // namespace effdet

/* Row blocks per image of the fused reduction (grid = blocks x B). */
// Block shape: nvb channel vectors x PY rows, <= 256 threads.  The channel range is cut into the number of
// chunks (grid.z) that fills the block best (C = 672: one chunk would be 168 threads, two chunks are 84 x 3 =
// 252), then rows_per_block is a multiple of kBnU * PY (no one-row-at-a-time tail except in an image's last
// block) sized for about two waves of three blocks per SM over (blocks, B, chunks).
static void se_bn_geometry(int C, int *chunks, int *nvb, int *PY) {
    int nvec = C / 4; if (nvec < 1) nvec = 1;
    int best_t = 0;
    for (int ch = 1; ch <= 8; ++ch) {
        const int v = (nvec + ch - 1) / ch;
        if (v > 256) continue;
        int py = 256 / v; if (py < 1) py = 1;
        if (v * py > best_t + 8) { best_t = v * py; *chunks = ch; *nvb = v; *PY = py; }
    }
}
static int se_bn_rows_per_block(int B, int HW, int chunks, int PY) {
    const int unit = kBnU * PY;
    long want = ((long)kNumSMs * 3 * 2 + (long)B * chunks - 1) / ((long)B * chunks);     // blocks per (image, chunk)
    if (want < 1) want = 1;
    long rpb = ((long)HW + want - 1) / want;
    rpb = (rpb + unit - 1) / unit * unit;
    if (rpb < 4L * unit) rpb = 4L * unit;
    return (int)rpb;
}
extern "C" int effdet_se_bn_backward_blocks(int B, int HW, int C, int dtype) {
    (void)dtype;
    int chunks = 1, nvb = 1, PY = 1;
    se_bn_geometry(C, &chunks, &nvb, &PY);
    return (int)cdiv((size_t)HW, (size_t)se_bn_rows_per_block(B, HW, chunks, PY));
}

/* Fused backward of the squeeze-excite gate and of the depthwise BatchNorm (training statistics) + swish of an
 * MBConv block (efficientnet.py:242-286): dyg (B,HW,C) gradient of y*gate, z (B,HW,C) the raw depthwise output
 * -> dz (B,HW,C) gradient of z, gradients of the SE kernels / biases and of gamma / beta.  Same results as
 * effdet_se_backward followed by effdet_bn_act_backward(act = swish) up to rounding (dy stays in fp32 here),
 * in five instead of nine passes over the tensors.  Scratch (floats): dg_partial B*(nblk+1)*C, bn_partial
 * B*(nblk+1)*4*C, bn_rows B*2*C, fc_scratch B*(2*C*R + R + C), dmean B*C, k123 3*C; nblk =
 * effdet_se_bn_backward_blocks(). */
extern "C" int effdet_se_bn_backward(const void *dyg, const void *z, const float *gate, const float *se_sum,
                                     int se_blocks, const float *w1, const float *b1, const float *w2,
                                     const float *b2, float *dw1, float *db1, float *dw2, float *db2,
                                     const float *gamma, const float *save_mean, const float *save_invstd,
                                     const float *ua, const float *ub, float *dgamma, float *dbeta, void *dz,
                                     float *k123, float *dg_partial, float *bn_partial, float *bn_rows, int nblk,
                                     float *fc_scratch, float *dmean, int B, int HW, int C, int R, int dtype,
                                     void *stream) {
    EFFDET_REQUIRE(dyg && z && gate && se_sum && w1 && b1 && w2 && b2 && dw1 && db1 && dw2 && db2 && gamma &&
                       save_mean && save_invstd && ua && ub && dz && k123 && dg_partial && bn_partial && bn_rows &&
                       fc_scratch && dmean, "null pointer");
    EFFDET_REQUIRE(B > 0 && HW > 0 && C > 0 && C % 8 == 0 && R > 0 && R <= 512 && se_blocks > 0, "bad sizes");
    EFFDET_REQUIRE(nblk == effdet_se_bn_backward_blocks(B, HW, C, dtype), "nblk must come from effdet_se_bn_backward_blocks");
    cudaStream_t st = as_stream(stream);
    int chunks = 1, nvb = 1, PY = 1;
    se_bn_geometry(C, &chunks, &nvb, &PY);
    const int rpb = se_bn_rows_per_block(B, HW, chunks, PY);
    const size_t sm = PY > 1 ? (size_t)PY * 5 * nvb * 4 * sizeof(float) : 0;
    EFFDET_REQUIRE(sm <= 48 * 1024, "reduction scratch too large");
    dim3 grid(nblk, B, chunks);
    DISPATCH_TB(dtype,
        (se_bn_bwd_reduce_kernel<float><<<grid, nvb * PY, sm, st>>>((const float *)dyg, (const float *)z, ua, ub,
                                                                    HW, C, nvb, rpb, dg_partial, bn_partial, 0u)),
        (se_bn_bwd_reduce_kernel<__nv_bfloat16><<<grid, nvb * PY, sm, st>>>(
            (const __nv_bfloat16 *)dyg, (const __nv_bfloat16 *)z, ua, ub, HW, C, nvb, rpb, dg_partial, bn_partial, 0u)))
    EFFDET_LAUNCHED();
    float *tot_dg = dg_partial + (size_t)B * nblk * C, *tot_bn = bn_partial + (size_t)B * nblk * 4 * C;
    se_bn_partial_sum_kernel<<<dim3(cdiv((size_t)C, 32), B), 512, 0, st>>>(dg_partial, bn_partial, nblk, C, tot_dg, tot_bn);
    EFFDET_LAUNCHED();
    {   // SE fully-connected backward: the launches of effdet_se_backward between its two tensor passes
        const int nch = (C + 255) / 256;
        const size_t per = (size_t)2 * C * R + R + C;
        EFFDET_REQUIRE((size_t)B * (2 * C + 3 * R + (size_t)nch * R) <= (size_t)B * per, "fc_scratch too small");
        float *mean_g = fc_scratch, *s1_g = mean_g + (size_t)B * C, *rr_g = s1_g + (size_t)B * R;
        float *ds2_g = rr_g + (size_t)B * R, *ds1_g = ds2_g + (size_t)B * C, *ds1p = ds1_g + (size_t)B * R;
        const size_t sm1 = (size_t)(C + 512) * sizeof(float);
        EFFDET_REQUIRE(sm1 <= 48 * 1024, "C too large");
        if (se_cluster_ok(C, R)) {
            EFFDET_CUDA(se_cluster_launch(st, se_sum, se_blocks, 1.f / (float)HW, w1, b1, w2, b2, nullptr, B, C, R,
                                          mean_g, s1_g, rr_g));
        } else {
            se_bwd_phase1_kernel<<<B, 512, sm1, st>>>(se_sum, se_blocks, 1.f / (float)HW, w1, b1, C, R, mean_g, s1_g, rr_g);
        }
        EFFDET_LAUNCHED();
        dim3 g2(nch, B);
        se_bwd_phase2_kernel<<<g2, 256, (size_t)9 * R * sizeof(float), st>>>(rr_g, tot_dg, 1, w2, b2, C, R, ds2_g, ds1p);
        EFFDET_LAUNCHED();
        se_bwd_phase3_kernel<<<g2, 256, (size_t)R * sizeof(float), st>>>(ds1p, nch, s1_g, w1, C, R, ds1_g, dmean);
        EFFDET_LAUNCHED();
        se_bwd_phase4_kernel<<<cdiv(per, 256), 256, 0, st>>>(mean_g, rr_g, ds2_g, ds1_g, B, C, R, dw1, dw2, db1, db2);
        EFFDET_LAUNCHED();
    }
    se_bn_bwd_combine_kernel<<<dim3(cdiv((size_t)C, 256), B), 256, 0, st>>>(tot_bn, 1, gate, dmean, save_mean,
                                                                           1.f / (float)HW, C, bn_rows);
    EFFDET_LAUNCHED();
    bn_act_bwd_finalize_kernel<<<cdiv((size_t)C * 32, 256), 256, 0, st>>>(bn_rows, B, (double)B * (double)HW, gamma,
                                                                        save_mean, save_invstd, k123, dgamma, dbeta, C);
    EFFDET_LAUNCHED();
    {
        const size_t rpa = (size_t)rpb;
        dim3 ga(cdiv((size_t)HW, rpa), B, chunks);
        DISPATCH_TB(dtype,
            (se_bn_bwd_apply_kernel<float><<<ga, nvb * PY, 0, st>>>((const float *)dyg, (const float *)z, ua, ub, k123,
                                                                    gate, dmean, 1.f / (float)HW, (float *)dz, HW, C,
                                                                    nvb, (int)rpa, 0u)),
            (se_bn_bwd_apply_kernel<__nv_bfloat16><<<ga, nvb * PY, 0, st>>>(
                (const __nv_bfloat16 *)dyg, (const __nv_bfloat16 *)z, ua, ub, k123, gate, dmean, 1.f / (float)HW,
                (__nv_bfloat16 *)dz, HW, C, nvb, (int)rpa, 0u)))
    }
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}
