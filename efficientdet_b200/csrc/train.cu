// Training-mode kernels of the BiFPN / heads (HBM-bound NHWC passes, fp32 math, fp32 or bf16
// storage).  Everything the Keras/TF autodiff + fused-BN + SGD ops did for the reference's
// model.fit (train_tpu.py:330-346), restated as explicit kernels:
//   resample_fuse_kernel    model.py:154-194/226-266 + layers.py:26-31  (forward, keeps f)
//   colreduce_kernel<MODE>  BN batch statistics / BN backward sums / bias gradients
//   bn_train_finalize       batch mean/var -> scale/shift, moving-average update (momentum .997)
//   scale_shift_act_kernel  y = act(z*scale + shift)
//   bn_bwd_finalize/apply   dz = k1*dy_masked + k2*z + k3 ; dgamma, dbeta
//   dw_wgrad_kernel         depthwise-kernel gradient
//   fuse_bwd_kernel         gradient routing through nearest-upsample (sum-pool) /
//                           2x2 max-pool (first-argmax) / same-resolution inputs
//   fuse_wgrad_kernel       fusion-weight gradients (layers.py:26-31)
//   sgd_momentum_kernel     train_tpu.py:268-269 keras SGD(lr, decay, momentum)
// All reductions use per-block partials summed in a fixed order: training is deterministic.
#include "common.cuh"

namespace effdet {

template <typename T, int CV> struct VecT;
template <> struct VecT<float, 4> {
    static __device__ __forceinline__ void load(const float *p, float *v) {
        float4 t = *reinterpret_cast<const float4 *>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float *p, const float *v) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct VecT<__nv_bfloat16, 8> {
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
template <int CV> __device__ __forceinline__ void ldf(const float *p, float *v) {
#pragma unroll
    for (int i = 0; i < CV; i += 4) {
        float4 t = *reinterpret_cast<const float4 *>(p + i);
        v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
}

struct FuseCoef { float c[3]; float inv; };
__device__ __forceinline__ FuseCoef fuse_coef(const float *fw, int n, float eps) {
    FuseCoef f;
    if (!fw) { f.c[0] = f.c[1] = f.c[2] = 1.f; f.inv = 1.f; return f; }
    float r0 = fmaxf(fw[0], 0.f), r1 = fmaxf(fw[1], 0.f), r2 = n > 2 ? fmaxf(fw[2], 0.f) : 0.f;
    f.c[0] = r0; f.c[1] = r1; f.c[2] = r2; f.inv = r0 + r1 + r2 + eps;
    return f;
}

// load the (resampled) value of input 0 at output pixel (y,x)
template <typename T, int CV>
__device__ __forceinline__ void load_resampled(const T *p0, int mode0, int y, int x, int W0, int C,
                                               int c, float *a) {
    if (mode0 == 1) {
        VecT<T, CV>::load(p0 + ((size_t)(y >> 1) * W0 + (x >> 1)) * C + c, a);
    } else if (mode0 == 2) {
        const T *q = p0 + ((size_t)(2 * y) * W0 + 2 * x) * C + c;
        float t[CV];
        VecT<T, CV>::load(q, a);
        VecT<T, CV>::load(q + C, t);
#pragma unroll
        for (int k = 0; k < CV; ++k) a[k] = fmaxf(a[k], t[k]);
        VecT<T, CV>::load(q + (size_t)W0 * C, t);
#pragma unroll
        for (int k = 0; k < CV; ++k) a[k] = fmaxf(a[k], t[k]);
        VecT<T, CV>::load(q + (size_t)W0 * C + C, t);
#pragma unroll
        for (int k = 0; k < CV; ++k) a[k] = fmaxf(a[k], t[k]);
    } else {
        VecT<T, CV>::load(p0 + ((size_t)y * W0 + x) * C + c, a);
    }
}

// ------------------------------------------------------------------ fusion forward (keeps f)
template <typename T, int CV>
__global__ void __launch_bounds__(256)
resample_fuse_kernel(const T *__restrict__ in0, int mode0, const T *__restrict__ in1,
                     const T *__restrict__ in2, const float *__restrict__ fw, float eps,
                     T *__restrict__ out, int B, int H, int W, int C) {
    EFFDET_PDL_SYNC();
    const int nvec = C / CV;
    const size_t total = (size_t)B * H * W * nvec;
    const FuseCoef fc = fuse_coef(fw, in2 ? 3 : 2, eps);
    const int H0 = mode0 == 1 ? H / 2 : (mode0 == 2 ? H * 2 : H);
    const int W0 = mode0 == 1 ? W / 2 : (mode0 == 2 ? W * 2 : W);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const int cv = (int)(i % nvec);
        size_t pix = i / nvec;
        const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((size_t)W * H));
        const int c = cv * CV;
        float a[CV], bb[CV], o[CV];
        load_resampled<T, CV>(in0 + (size_t)b * H0 * W0 * C, mode0, y, x, W0, C, c, a);
        VecT<T, CV>::load(in1 + pix * C + c, bb);
        if (fw) {
#pragma unroll
            for (int k = 0; k < CV; ++k) o[k] = fc.c[0] * a[k] + fc.c[1] * bb[k];
        } else {
#pragma unroll
            for (int k = 0; k < CV; ++k) o[k] = a[k] + bb[k];
        }
        if (in2) {
            float cc[CV];
            VecT<T, CV>::load(in2 + pix * C + c, cc);
#pragma unroll
            for (int k = 0; k < CV; ++k) o[k] += fw ? fc.c[2] * cc[k] : cc[k];
        }
        if (fw) {
#pragma unroll
            for (int k = 0; k < CV; ++k) o[k] = o[k] / fc.inv;
        }
        VecT<T, CV>::store(out + pix * C + c, o);
    }
}

// ------------------------------------------------------------------ column reductions
// x is (rows, C).  block = nvec x PY threads; blocks split the rows; partial[(blk)*2*C + s*C + c]
enum { RED_STATS = 0, RED_BNBWD = 1, RED_SUM = 2 };
template <typename T, int CV, int MODE>
__global__ void __launch_bounds__(1024, 1)      // <= 64 registers: four 256-thread blocks per SM (blocks grow to C/CV threads)
colreduce_kernel(const T *__restrict__ x, const T *__restrict__ y,
                                 const T *__restrict__ dy, const float *__restrict__ mean,
                                 const float *__restrict__ invstd, size_t rows, int C,
                                 int rows_per_block, float *__restrict__ partial, uint32_t zero) {
    EFFDET_PDL_SYNC();
    extern __shared__ float sred[];      // PY * 2 * C
    const int nvec = C / CV, PY = blockDim.x / nvec;
    const int cv = threadIdx.x % nvec, py = threadIdx.x / nvec, c = cv * CV;
    const size_t r0 = (size_t)blockIdx.x * rows_per_block;
    const size_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    float s1[CV], s2[CV], mu[CV], is[CV];
#pragma unroll
    for (int k = 0; k < CV; ++k) { s1[k] = 0.f; s2[k] = 0.f; mu[k] = 0.f; is[k] = 1.f; }
    if (MODE == RED_BNBWD) { ldf<CV>(mean + c, mu); ldf<CV>(invstd + c, is); }
    size_t r = r0 + py;
    if (MODE != RED_BNBWD) {
        // four rows (one 16-byte load each) in flight per thread
        for (; r + 3 * (size_t)PY < r1; r += 4 * (size_t)PY) {
            uint4 raw[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) raw[u] = ld16(x + (r + (size_t)u * PY) * C + c);
            tie_loads(raw, zero);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float v[CV];
                Unpack16<T, CV>::run(raw[u], v);
#pragma unroll
                for (int k = 0; k < CV; ++k) {
                    s1[k] += v[k];
                    if (MODE == RED_STATS) s2[k] = fmaf(v[k], v[k], s2[k]);
                }
            }
        }
    } else {
        // two rows of the three tensors (six 16-byte loads) in flight per thread
        for (; r + (size_t)PY < r1; r += 2 * (size_t)PY) {
            uint4 raw[6];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                raw[3 * u + 0] = ld16(x + (r + (size_t)u * PY) * C + c);
                raw[3 * u + 1] = ld16(dy + (r + (size_t)u * PY) * C + c);
                raw[3 * u + 2] = ld16(y + (r + (size_t)u * PY) * C + c);
            }
            tie_loads(raw, zero);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                float v[CV], g[CV], yy[CV];
                Unpack16<T, CV>::run(raw[3 * u + 0], v);
                Unpack16<T, CV>::run(raw[3 * u + 1], g);
                Unpack16<T, CV>::run(raw[3 * u + 2], yy);
#pragma unroll
                for (int k = 0; k < CV; ++k) {
                    const float gm = yy[k] > 0.f ? g[k] : 0.f;
                    s1[k] += gm;
                    s2[k] = fmaf(gm, (v[k] - mu[k]) * is[k], s2[k]);
                }
            }
        }
    }
    for (; r < r1; r += PY) {
        float v[CV];
        if (MODE == RED_STATS) {
            VecT<T, CV>::load(x + r * C + c, v);
#pragma unroll
            for (int k = 0; k < CV; ++k) { s1[k] += v[k]; s2[k] = fmaf(v[k], v[k], s2[k]); }
        } else if (MODE == RED_BNBWD) {
            float g[CV], yy[CV];
            VecT<T, CV>::load(x + r * C + c, v);
            VecT<T, CV>::load(dy + r * C + c, g);
            VecT<T, CV>::load(y + r * C + c, yy);
#pragma unroll
            for (int k = 0; k < CV; ++k) {
                const float gm = yy[k] > 0.f ? g[k] : 0.f;
                s1[k] += gm;
                s2[k] = fmaf(gm, (v[k] - mu[k]) * is[k], s2[k]);
            }
        } else {
            VecT<T, CV>::load(x + r * C + c, v);
#pragma unroll
            for (int k = 0; k < CV; ++k) s1[k] += v[k];
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
        for (int r = 0; r < PY; ++r) t += sred[(size_t)r * 2 * C + i];
        partial[(size_t)blockIdx.x * 2 * C + i] = t;
    }
}


// Sum over `nblk` partial rows for one channel with a whole warp: lane l adds blocks l, l+32, ...
// (fixed order), then a fixed-shape shuffle tree -> deterministic, and ~32x less serial latency
// than one thread walking all blocks.
__device__ __forceinline__ double warp_block_sum(const float *__restrict__ partial, int nblk,
                                                 size_t stride, size_t offset) {
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    int b = lane;
    // eight independent loads in flight per lane (one L2 round trip per 256 partial rows instead of
    // one per 32); the additions keep the sequential order, so the result is unchanged
    for (; b + 7 * 32 < nblk; b += 8 * 32) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(partial + (size_t)(b + 32 * u) * stride + offset);
#pragma unroll
        for (int u = 0; u < 8; ++u) s += (double)v[u];
    }
    for (; b < nblk; b += 32) s += (double)__ldcg(partial + (size_t)b * stride + offset);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    return s;
}
// two columns of the same partial matrix at once (sum and sum of squares / the two BN-backward sums):
// sixteen loads in flight per lane
__device__ __forceinline__ void warp_block_sum2(const float *__restrict__ partial, int nblk, size_t stride,
                                                size_t off1, size_t off2, double &o1, double &o2) {
    const int lane = threadIdx.x & 31;
    double s1 = 0.0, s2 = 0.0;
    int b = lane;
    for (; b + 7 * 32 < nblk; b += 8 * 32) {
        float v[8], q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float *row = partial + (size_t)(b + 32 * u) * stride;
            v[u] = __ldcg(row + off1);
            q[u] = __ldcg(row + off2);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { s1 += (double)v[u]; s2 += (double)q[u]; }
    }
    for (; b < nblk; b += 32) {
        const float *row = partial + (size_t)b * stride;
        s1 += (double)__ldcg(row + off1);
        s2 += (double)__ldcg(row + off2);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, d);
        s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    }
    o1 = s1; o2 = s2;
}

// BN training: batch statistics -> scale/shift (+ saved mean/invstd, moving averages)
__global__ void bn_train_finalize_kernel(const float *__restrict__ partial, int nblk, double count,
                                         const float *__restrict__ gamma,
                                         const float *__restrict__ beta, float eps, float momentum,
                                         float *__restrict__ moving_mean,
                                         float *__restrict__ moving_var, float *__restrict__ scale,
                                         float *__restrict__ shift, float *__restrict__ save_mean,
                                         float *__restrict__ save_invstd, int C) {
    EFFDET_PDL_SYNC();
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;     // one warp per channel
    if (c >= C) return;
    double s1, s2;
    warp_block_sum2(partial, nblk, 2 * (size_t)C, c, (size_t)C + c, s1, s2);
    if (threadIdx.x & 31) return;
    const double m = s1 / count;
    double var = s2 / count - m * m;
    if (var < 0.0) var = 0.0;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * is;
    scale[c] = sc;
    shift[c] = beta[c] - (float)m * sc;
    save_mean[c] = (float)m;
    save_invstd[c] = is;
    if (moving_mean) {
        const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
        moving_mean[c] = moving_mean[c] * momentum + (float)m * (1.f - momentum);
        moving_var[c] = moving_var[c] * momentum + (float)unb * (1.f - momentum);
    }
}

// sums `nblk` partial rows (layout [blk][2][C], first half) into out (bias / generic gradients)
// C = fold * Cout: the matrix was viewed as (rows/fold, fold*Cout) to get vectorisable rows
__global__ void colsum_finalize_kernel(const float *__restrict__ partial, int nblk, int C, int fold,
                                       float *__restrict__ out, int accumulate) {
    EFFDET_PDL_SYNC();
    const int Cout = C / fold;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;     // one warp per channel
    if (c >= Cout) return;
    double t = 0.0;
    for (int j = 0; j < fold; ++j) t += warp_block_sum(partial, nblk, 2 * (size_t)C, (size_t)j * Cout + c);
    if (threadIdx.x & 31) return;
    out[c] = accumulate ? out[c] + (float)t : (float)t;
}

// BN backward coefficients: dz = k1*dy_masked + k2*z + k3 ; dgamma, dbeta
__global__ void bn_bwd_finalize_kernel(const float *__restrict__ partial, int nblk, double count,
                                       const float *__restrict__ gamma,
                                       const float *__restrict__ mean,
                                       const float *__restrict__ invstd, float *__restrict__ k123,
                                       float *__restrict__ dgamma, float *__restrict__ dbeta, int C) {
    EFFDET_PDL_SYNC();
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;     // one warp per channel
    if (c >= C) return;
    double s1, s2;
    warp_block_sum2(partial, nblk, 2 * (size_t)C, c, (size_t)C + c, s1, s2);
    if (threadIdx.x & 31) return;
    const float g = gamma[c], is = invstd[c], mu = mean[c];
    const float m1 = (float)(s1 / count), m2 = (float)(s2 / count);
    k123[c] = g * is;
    k123[C + c] = -g * is * is * m2;
    k123[2 * C + c] = -g * is * (m1 - mu * is * m2);
    if (dgamma) dgamma[c] = (float)s2;
    if (dbeta) dbeta[c] = (float)s1;
}

// ------------------------------------------------------------------ BN + ReLU backward, one launch
// Thread-block clusters of 8 CTAs: a cluster owns CV channels (one 16-byte vector column), its CTAs split the rows.
// Pass 1 accumulates sum(dy_masked) and sum(dy_masked * xhat) of the CTA's rows, the eight partial pairs are
// exchanged through distributed shared memory and added in rank order (deterministic), every CTA derives the
// same k1/k2/k3, and pass 2 writes dz for its own rows (which it has just read: L2/L1-hot).  Replaces the
// reduce -> finalize -> apply chain (three launches, ~13-34 us on the small / medium BiFPN levels, which are
// launch- and latency-bound, not bandwidth-bound).
constexpr int kBnClu = 8, kBnCluThreads = 512;
template <typename T, int CV>
__global__ void __launch_bounds__(kBnCluThreads)
bn_relu_bwd_cluster_kernel(const T *__restrict__ dy, const T *__restrict__ y, const T *__restrict__ z, size_t rows,
                           int C, const float *__restrict__ gamma, const float *__restrict__ mean,
                           const float *__restrict__ invstd, float *__restrict__ dgamma,
                           float *__restrict__ dbeta, T *__restrict__ dz, float *__restrict__ k123,
                           uint32_t zero) {
    EFFDET_PDL_SYNC();
    __shared__ double wsum[kBnCluThreads / 32][2 * CV];
    __shared__ double mine[2 * CV];            // this CTA's partial sums (read by the peers)
    __shared__ float coef[3 * CV];
    unsigned rank;
    asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int c = (blockIdx.x / kBnClu) * CV, tid = threadIdx.x;
    const size_t rs = (rows + kBnClu - 1) / kBnClu;
    const size_t r0 = (size_t)rank * rs, r1 = r0 + rs < rows ? r0 + rs : rows;
    float s1[CV], s2[CV], mu[CV], is[CV];
    ldf<CV>(mean + c, mu); ldf<CV>(invstd + c, is);
#pragma unroll
    for (int k = 0; k < CV; ++k) { s1[k] = 0.f; s2[k] = 0.f; }
    size_t r = r0 + tid;
    for (; r + kBnCluThreads < r1; r += 2 * kBnCluThreads) {         // two rows of the three tensors in flight
        uint4 raw[6];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            raw[3 * u + 0] = ld16(z + (r + (size_t)u * kBnCluThreads) * C + c);
            raw[3 * u + 1] = ld16(dy + (r + (size_t)u * kBnCluThreads) * C + c);
            raw[3 * u + 2] = ld16(y + (r + (size_t)u * kBnCluThreads) * C + c);
        }
        tie_loads(raw, zero);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float v[CV], g[CV], yy[CV];
            Unpack16<T, CV>::run(raw[3 * u + 0], v);
            Unpack16<T, CV>::run(raw[3 * u + 1], g);
            Unpack16<T, CV>::run(raw[3 * u + 2], yy);
#pragma unroll
            for (int k = 0; k < CV; ++k) {
                const float gm = yy[k] > 0.f ? g[k] : 0.f;
                s1[k] += gm;
                s2[k] = fmaf(gm, (v[k] - mu[k]) * is[k], s2[k]);
            }
        }
    }
    for (; r < r1; r += kBnCluThreads) {
        float v[CV], g[CV], yy[CV];
        VecT<T, CV>::load(z + r * C + c, v);
        VecT<T, CV>::load(dy + r * C + c, g);
        VecT<T, CV>::load(y + r * C + c, yy);
#pragma unroll
        for (int k = 0; k < CV; ++k) {
            const float gm = yy[k] > 0.f ? g[k] : 0.f;
            s1[k] += gm;
            s2[k] = fmaf(gm, (v[k] - mu[k]) * is[k], s2[k]);
        }
    }
    // block reduction: fixed shuffle tree per warp, then the warps in order (double)
#pragma unroll
    for (int k = 0; k < CV; ++k) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], d);
            s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], d);
        }
    }
    if ((tid & 31) == 0) {
#pragma unroll
        for (int k = 0; k < CV; ++k) { wsum[tid >> 5][k] = (double)s1[k]; wsum[tid >> 5][CV + k] = (double)s2[k]; }
    }
    __syncthreads();
    if (tid < 2 * CV) {
        double t = 0.0;
        for (int w = 0; w < kBnCluThreads / 32; ++w) t += wsum[w][tid];
        mine[tid] = t;
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (tid < CV) {
        double t1 = 0.0, t2 = 0.0;
#pragma unroll
        for (int q = 0; q < kBnClu; ++q) {
            uint32_t a1, a2;
            asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a1)
                : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(&mine[tid]))), "r"(q));
            asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a2)
                : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(&mine[CV + tid]))), "r"(q));
            double v1, v2;
            asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v1) : "r"(a1));
            asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v2) : "r"(a2));
            t1 += v1; t2 += v2;
        }
        const double count = (double)rows;
        const float g = gamma[c + tid], isv = invstd[c + tid], muv = mean[c + tid];
        const float m1 = (float)(t1 / count), m2 = (float)(t2 / count);
        const float k1 = g * isv, k2 = -g * isv * isv * m2, k3 = -g * isv * (m1 - muv * isv * m2);
        coef[tid] = k1; coef[CV + tid] = k2; coef[2 * CV + tid] = k3;
        if (rank == 0) {
            k123[c + tid] = k1; k123[C + c + tid] = k2; k123[2 * C + c + tid] = k3;
            if (dgamma) dgamma[c + tid] = (float)t2;
            if (dbeta) dbeta[c + tid] = (float)t1;
        }
    }
    // peers may still be reading `mine`: nobody leaves / overwrites before everyone has read; also publishes coef
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    __syncthreads();
    float k1[CV], k2[CV], k3[CV];
#pragma unroll
    for (int k = 0; k < CV; ++k) { k1[k] = coef[k]; k2[k] = coef[CV + k]; k3[k] = coef[2 * CV + k]; }
    r = r0 + tid;
    for (; r + kBnCluThreads < r1; r += 2 * kBnCluThreads) {
        uint4 raw[6];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            raw[3 * u + 0] = ld16(z + (r + (size_t)u * kBnCluThreads) * C + c);
            raw[3 * u + 1] = ld16(dy + (r + (size_t)u * kBnCluThreads) * C + c);
            raw[3 * u + 2] = ld16(y + (r + (size_t)u * kBnCluThreads) * C + c);
        }
        tie_loads(raw, zero);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float v[CV], g[CV], yy[CV];
            Unpack16<T, CV>::run(raw[3 * u + 0], v);
            Unpack16<T, CV>::run(raw[3 * u + 1], g);
            Unpack16<T, CV>::run(raw[3 * u + 2], yy);
#pragma unroll
            for (int k = 0; k < CV; ++k) g[k] = k1[k] * (yy[k] > 0.f ? g[k] : 0.f) + k2[k] * v[k] + k3[k];
            VecT<T, CV>::store(dz + (r + (size_t)u * kBnCluThreads) * C + c, g);
        }
    }
    for (; r < r1; r += kBnCluThreads) {
        float v[CV], g[CV], yy[CV];
        VecT<T, CV>::load(z + r * C + c, v);
        VecT<T, CV>::load(dy + r * C + c, g);
        VecT<T, CV>::load(y + r * C + c, yy);
#pragma unroll
        for (int k = 0; k < CV; ++k) g[k] = k1[k] * (yy[k] > 0.f ? g[k] : 0.f) + k2[k] * v[k] + k3[k];
        VecT<T, CV>::store(dz + r * C + c, g);
    }
}
template <typename T, int CV>
static cudaError_t launch_bn_relu_bwd_cluster(cudaStream_t st, const void *dy, const void *y, const void *z,
                                              size_t rows, int C, const float *gamma, const float *mean,
                                              const float *invstd, float *dgamma, float *dbeta, void *dz,
                                              float *k123) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((C / CV) * kBnClu); cfg.blockDim = dim3(kBnCluThreads); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kBnClu; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    static const bool pdl = pdl_level() >= EFFDET_PDL_TU_LEVEL;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, bn_relu_bwd_cluster_kernel<T, CV>, (const T *)dy, (const T *)y, (const T *)z, rows, C,
                              gamma, mean, invstd, dgamma, dbeta, (T *)dz, k123, 0u);
}

template <typename T, int CV>
__global__ void __launch_bounds__(256)
scale_shift_act_kernel(const T *__restrict__ z, const float *__restrict__ scale,
                       const float *__restrict__ shift, T *__restrict__ y, size_t nvec_total,
                       int C, int act, uint32_t zero) {
    EFFDET_PDL_SYNC();
    const unsigned nvec = C / CV;
    const size_t stride = (size_t)gridDim.x * 256;
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    for (; i + 3 * stride < nvec_total; i += 4 * stride) {      // four vectors in flight
        uint4 raw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) raw[u] = ld16(z + (i + u * stride) * CV);
        tie_loads(raw, zero);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const size_t iv = i + u * stride;
            const int c = (int)((unsigned)iv % nvec) * CV;
            float v[CV], sc[CV], sh[CV];
            Unpack16<T, CV>::run(raw[u], v);
            ldf<CV>(scale + c, sc);
            ldf<CV>(shift + c, sh);
#pragma unroll
            for (int k = 0; k < CV; ++k) v[k] = activate_io<T>(fmaf(v[k], sc[k], sh[k]), act);
            VecT<T, CV>::store(y + iv * CV, v);
        }
    }
    for (; i < nvec_total; i += stride) {
        const int c = (int)((unsigned)i % nvec) * CV;        // nvec_total < 2^32 (checked by the launcher)
        float v[CV], sc[CV], sh[CV];
        VecT<T, CV>::load(z + i * CV, v);
        ldf<CV>(scale + c, sc);
        ldf<CV>(shift + c, sh);
#pragma unroll
        for (int k = 0; k < CV; ++k) v[k] = activate_io<T>(fmaf(v[k], sc[k], sh[k]), act);
        VecT<T, CV>::store(y + i * CV, v);
    }
}

// dz = k1*(dy * [y>0]) + k2*z + k3     (frozen BN: k2 = k3 = 0, k1 = scale)
template <typename T, int CV>
__global__ void __launch_bounds__(256, 4)
bn_bwd_apply_kernel(const T *__restrict__ dy, const T *__restrict__ y, const T *__restrict__ z,
                    const float *__restrict__ k123, T *__restrict__ dz, size_t nvec_total, int C, uint32_t zero) {
    EFFDET_PDL_SYNC();
    const unsigned nvec = C / CV;
    const size_t stride = (size_t)gridDim.x * 256;
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    for (; i + stride < nvec_total; i += 2 * stride) {      // two vectors of the three tensors in flight
        uint4 raw[6];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            raw[3 * u + 0] = ld16(dy + (i + u * stride) * CV);
            raw[3 * u + 1] = ld16(y + (i + u * stride) * CV);
            raw[3 * u + 2] = ld16(z + (i + u * stride) * CV);
        }
        tie_loads(raw, zero);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const size_t iv = i + u * stride;
            const int c = (int)((unsigned)iv % nvec) * CV;
            float g[CV], yy[CV], zz[CV], k1[CV], k2[CV], k3[CV];
            Unpack16<T, CV>::run(raw[3 * u + 0], g);
            Unpack16<T, CV>::run(raw[3 * u + 1], yy);
            Unpack16<T, CV>::run(raw[3 * u + 2], zz);
            ldf<CV>(k123 + c, k1);
            ldf<CV>(k123 + C + c, k2);
            ldf<CV>(k123 + 2 * C + c, k3);
#pragma unroll
            for (int k = 0; k < CV; ++k) g[k] = k1[k] * (yy[k] > 0.f ? g[k] : 0.f) + k2[k] * zz[k] + k3[k];
            VecT<T, CV>::store(dz + iv * CV, g);
        }
    }
    for (; i < nvec_total; i += stride) {
        const int c = (int)((unsigned)i % nvec) * CV;
        float g[CV], yy[CV], zz[CV], k1[CV], k2[CV], k3[CV];
        VecT<T, CV>::load(dy + i * CV, g);
        VecT<T, CV>::load(y + i * CV, yy);
        VecT<T, CV>::load(z + i * CV, zz);
        ldf<CV>(k123 + c, k1);
        ldf<CV>(k123 + C + c, k2);
        ldf<CV>(k123 + 2 * C + c, k3);
#pragma unroll
        for (int k = 0; k < CV; ++k) g[k] = k1[k] * (yy[k] > 0.f ? g[k] : 0.f) + k2[k] * zz[k] + k3[k];
        VecT<T, CV>::store(dz + i * CV, g);
    }
}

// ------------------------------------------------------------------ depthwise weight gradient
// dW[tap][c] = sum_{b,y,x} f[b, y+ky-1, x+kx-1, c] * dz[b,y,x,c]   (3x3, stride 1, SAME)
template <typename T, int CV>
__global__ void dw_wgrad_kernel(const T *__restrict__ f, const T *__restrict__ dz, int B, int H, int W,
                                int C, int pix_per_block, float *__restrict__ partial) {
    EFFDET_PDL_SYNC();
    extern __shared__ float sred[];      // PY * 9 * C
    const int nvec = C / CV, PY = blockDim.x / nvec;
    const int cv = threadIdx.x % nvec, py = threadIdx.x / nvec, c = cv * CV;
    const size_t total = (size_t)B * H * W;
    const size_t p0 = (size_t)blockIdx.x * pix_per_block;
    const size_t p1 = p0 + pix_per_block < total ? p0 + pix_per_block : total;
    float acc[9][CV];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int k = 0; k < CV; ++k) acc[t][k] = 0.f;
    for (size_t p = p0 + py; p < p1; p += PY) {
        const int x = (int)(p % W), y = (int)((p / W) % H);
        const size_t bbase = (p / ((size_t)W * H)) * (size_t)H * W;
        float g[CV];
        VecT<T, CV>::load(dz + p * C + c, g);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = y + ky - 1;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = x + kx - 1;
                if (ix < 0 || ix >= W) continue;
                float v[CV];
                VecT<T, CV>::load(f + (bbase + (size_t)iy * W + ix) * C + c, v);
#pragma unroll
                for (int k = 0; k < CV; ++k) acc[ky * 3 + kx][k] = fmaf(v[k], g[k], acc[ky * 3 + kx][k]);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int k = 0; k < CV; ++k) sred[((size_t)py * 9 + t) * C + c + k] = acc[t][k];
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) {
        float t = 0.f;
        for (int r = 0; r < PY; ++r) t += sred[(size_t)r * 9 * C + i];
        partial[(size_t)blockIdx.x * 9 * C + i] = t;
    }
}
__global__ void sum_partials_kernel(const float *__restrict__ partial, int nblk, int n,
                                    float *__restrict__ out, int accumulate) {
    EFFDET_PDL_SYNC();
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;     // one warp per element
    if (i >= n) return;
    const double t = warp_block_sum(partial, nblk, (size_t)n, (size_t)i);
    if (threadIdx.x & 31) return;
    out[i] = accumulate ? out[i] + (float)t : (float)t;
}

// ------------------------------------------------------------------ fusion backward
// which = 0: gradient of the resampled input (mode0 routing); which = 1/2: same-resolution inputs
template <typename T, int CV>
__global__ void __launch_bounds__(256)
fuse_bwd_kernel(const T *__restrict__ df, int which, int mode0, const T *__restrict__ in0,
                const float *__restrict__ fw, int n_in, float eps, T *__restrict__ dst,
                int accumulate, int B, int H, int W, int C) {
    EFFDET_PDL_SYNC();
    // H, W = resolution of df (the node's resolution)
    const FuseCoef fc = fuse_coef(fw, n_in, eps);
    const float coef = fw ? fc.c[which] / fc.inv : 1.f;
    const int nvec = C / CV;
    int Hd = H, Wd = W;
    if (which == 0 && mode0 == 1) { Hd = H / 2; Wd = W / 2; }
    if (which == 0 && mode0 == 2) { Hd = H * 2; Wd = W * 2; }
    const size_t total = (size_t)B * Hd * Wd * nvec;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const int cv = (int)(i % nvec), c = cv * CV;
        const size_t pix = i / nvec;
        const int x = (int)(pix % Wd), y = (int)((pix / Wd) % Hd), b = (int)(pix / ((size_t)Wd * Hd));
        const T *dfb = df + (size_t)b * H * W * C;
        float g[CV];
        if (which != 0 || mode0 == 0) {
            VecT<T, CV>::load(dfb + ((size_t)y * W + x) * C + c, g);
        } else if (mode0 == 1) {          // nearest upsample backward = 2x2 sum
            float t[CV];
            const T *q = dfb + ((size_t)(2 * y) * W + 2 * x) * C + c;
            VecT<T, CV>::load(q, g);
            VecT<T, CV>::load(q + C, t);
#pragma unroll
            for (int k = 0; k < CV; ++k) g[k] += t[k];
            VecT<T, CV>::load(q + (size_t)W * C, t);
#pragma unroll
            for (int k = 0; k < CV; ++k) g[k] += t[k];
            VecT<T, CV>::load(q + (size_t)W * C + C, t);
#pragma unroll
            for (int k = 0; k < CV; ++k) g[k] += t[k];
        } else {                          // max-pool backward: route to the first maximum
            const int wy = y >> 1, wx = x >> 1;
            if (wy >= H || wx >= W) {
#pragma unroll
                for (int k = 0; k < CV; ++k) g[k] = 0.f;
            } else {
                const T *q = in0 + (((size_t)b * Hd + 2 * wy) * Wd + 2 * wx) * C + c;
                float v[4][CV];
                VecT<T, CV>::load(q, v[0]);
                VecT<T, CV>::load(q + C, v[1]);
                VecT<T, CV>::load(q + (size_t)Wd * C, v[2]);
                VecT<T, CV>::load(q + (size_t)Wd * C + C, v[3]);
                float d[CV];
                VecT<T, CV>::load(dfb + ((size_t)wy * W + wx) * C + c, d);
                const int me = (y & 1) * 2 + (x & 1);
#pragma unroll
                for (int k = 0; k < CV; ++k) {
                    int arg = 0; float m = v[0][k];
#pragma unroll
                    for (int j = 1; j < 4; ++j) if (v[j][k] > m) { m = v[j][k]; arg = j; }
                    g[k] = arg == me ? d[k] : 0.f;
                }
            }
        }
        float o[CV];
        if (accumulate) {
            VecT<T, CV>::load(dst + pix * C + c, o);
#pragma unroll
            for (int k = 0; k < CV; ++k) o[k] += coef * g[k];
        } else {
#pragma unroll
            for (int k = 0; k < CV; ++k) o[k] = coef * g[k];
        }
        VecT<T, CV>::store(dst + pix * C + c, o);
    }
}

// partial[blk][4] = { sum df*(in0r - f), sum df*(in1 - f), sum df*(in2 - f), unused }
// (the difference is formed per element: subtracting two large sums would cancel badly)
template <typename T, int CV>
__global__ void __launch_bounds__(256)
fuse_wgrad_kernel(const T *__restrict__ df, const T *__restrict__ f, const T *__restrict__ in0,
                  int mode0, const T *__restrict__ in1, const T *__restrict__ in2, int B, int H,
                  int W, int C, float *__restrict__ partial) {
    EFFDET_PDL_SYNC();
    __shared__ float sh[4][8];
    const int nvec = C / CV;
    const size_t total = (size_t)B * H * W * nvec;
    const int H0 = mode0 == 1 ? H / 2 : (mode0 == 2 ? H * 2 : H);
    const int W0 = mode0 == 1 ? W / 2 : (mode0 == 2 ? W * 2 : W);
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const int cv = (int)(i % nvec), c = cv * CV;
        const size_t pix = i / nvec;
        const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((size_t)W * H));
        float g[CV], a[CV], v[CV], ff[CV];
        VecT<T, CV>::load(df + pix * C + c, g);
        VecT<T, CV>::load(f + pix * C + c, ff);
        load_resampled<T, CV>(in0 + (size_t)b * H0 * W0 * C, mode0, y, x, W0, C, c, a);
#pragma unroll
        for (int k = 0; k < CV; ++k) s[0] = fmaf(g[k], a[k] - ff[k], s[0]);
        VecT<T, CV>::load(in1 + pix * C + c, v);
#pragma unroll
        for (int k = 0; k < CV; ++k) s[1] = fmaf(g[k], v[k] - ff[k], s[1]);
        if (in2) {
            VecT<T, CV>::load(in2 + pix * C + c, v);
#pragma unroll
            for (int k = 0; k < CV; ++k) s[2] = fmaf(g[k], v[k] - ff[k], s[2]);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float v = s[j];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0) sh[j][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += sh[threadIdx.x][i];
        partial[(size_t)blockIdx.x * 4 + threadIdx.x] = t;
    }
}
__global__ void fuse_wgrad_finalize_kernel(const float *__restrict__ partial, int nblk,
                                           const float *__restrict__ fw, int n_in, float eps,
                                           float *__restrict__ dw) {
    EFFDET_PDL_SYNC();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s[4] = {0, 0, 0, 0};
    for (int b = 0; b < nblk; ++b)
        for (int j = 0; j < 4; ++j) s[j] += (double)partial[(size_t)b * 4 + j];
    float D = eps;
    for (int i = 0; i < n_in; ++i) D += fmaxf(fw[i], 0.f);
    for (int i = 0; i < n_in; ++i) dw[i] = fw[i] > 0.f ? (float)(s[i] / (double)D) : 0.f;
}

// ------------------------------------------------------------------ optimiser
__global__ void __launch_bounds__(256)
sgd_momentum_kernel(float *__restrict__ w, const float *__restrict__ g, float *__restrict__ v,
                    size_t n, float lr_t, float momentum, float grad_scale) {
    EFFDET_PDL_SYNC();
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const float vel = momentum * v[i] - lr_t * (g[i] * grad_scale);
        v[i] = vel;
        w[i] += vel;
    }
}

// the same update with the learning rate read from device memory: a CUDA graph holding this launch stays valid
// while the schedule changes lr_t every step (keras `decay`, cosine schedules)
__global__ void __launch_bounds__(256)
sgd_momentum_dev_lr_kernel(float *__restrict__ w, const float *__restrict__ g, float *__restrict__ v,
                           size_t n, const float *__restrict__ lr_dev, float momentum, float grad_scale) {
    EFFDET_PDL_SYNC();
    const float lr_t = *lr_dev;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const float vel = momentum * v[i] - lr_t * (g[i] * grad_scale);
        v[i] = vel;
        w[i] += vel;
    }
}

// flips a depthwise kernel (k,k,C) spatially: out[k*k-1-t][c] = in[t][c]
__global__ void flip_taps_kernel(const float *__restrict__ in, float *__restrict__ out, int taps, int n) {
    EFFDET_PDL_SYNC();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= taps * n) return;
    const int t = i / n, r = i - t * n;
    out[(size_t)(taps - 1 - t) * n + r] = in[i];
}
// dense conv weight (taps,Cin,Cout) -> (taps flipped, Cout, Cin)   (data-gradient kernel)
__global__ void conv_weight_transpose_kernel(const float *__restrict__ in, float *__restrict__ out,
                                             int taps, int Cin, int Cout) {
    EFFDET_PDL_SYNC();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)taps * Cin * Cout) return;
    const int co = (int)(i % Cout), ci = (int)((i / Cout) % Cin), t = (int)(i / ((size_t)Cout * Cin));
    out[((size_t)(taps - 1 - t) * Cout + co) * Cin + ci] = in[i];
}


// out (B,H,W,C) = dz (B,Ho,Wo,C) placed at (2*oy + ay, 2*ox + ax), zeros elsewhere.  The data gradient of
// a stride-2 convolution is then the stride-1 data gradient (tensor-core path) of this tensor.
template <typename T, int CV>
__global__ void __launch_bounds__(256)
zero_insert_kernel(const T *__restrict__ dz, T *__restrict__ out, int B, int Ho, int Wo, int C, int H, int W,
                   int ay, int ax) {
    EFFDET_PDL_SYNC();
    const int nvec = C / CV;
    const size_t total = (size_t)B * H * W * nvec;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const int c = (int)(i % nvec) * CV;
        const size_t pix = i / nvec;
        const int x = (int)(pix % W), y = (int)((pix / W) % H);
        const size_t b = pix / ((size_t)W * H);
        const int ny = y - ay, nx = x - ax;
        float v[CV];
#pragma unroll
        for (int k = 0; k < CV; ++k) v[k] = 0.f;
        if (ny >= 0 && nx >= 0 && !(ny & 1) && !(nx & 1) && (ny >> 1) < Ho && (nx >> 1) < Wo)
            VecT<T, CV>::load(dz + ((b * Ho + (ny >> 1)) * (size_t)Wo + (nx >> 1)) * C + c, v);
        VecT<T, CV>::store(out + pix * C + c, v);
    }
}

static bool a16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
// grid-stride kernels that need ~40 registers keep 6 x 256 threads per SM resident: 6 blocks per SM is one full
// wave (8 per SM would run as 1.33 waves)
static unsigned grid_for6(size_t n) {
    unsigned b = cdiv(n, 256);
    return b > (unsigned)kNumSMs * 6 ? kNumSMs * 6 : (b ? b : 1);
}
static unsigned grid_for4(size_t n) {            // batched-load kernels (<= 64 registers): four blocks per SM, one wave
    unsigned b = cdiv(n, 256 * 2);
    return b > (unsigned)kNumSMs * 4 ? kNumSMs * 4 : (b ? b : 1);
}
static unsigned grid_for(size_t n) {
    unsigned b = cdiv(n, 256);
    return b > (unsigned)kNumSMs * 8 ? kNumSMs * 8 : (b ? b : 1);
}

}  // namespace effdet

using namespace effdet;

#define DISPATCH_T(dtype, EXPR_F32, EXPR_BF16)                                        \
    if ((dtype) == EFFDET_F32) { EXPR_F32; }                                          \
    else if ((dtype) == EFFDET_BF16) { EXPR_BF16; }                                   \
    else return fail(EFFDET_E_INVALID, "%s: bad dtype", __func__);

extern "C" int effdet_resample_fuse(const void *in0, int mode0, const void *in1, const void *in2,
                                    const float *w, float eps, void *out, int B, int H, int W, int C,
                                    int dtype, void *stream) {
    EFFDET_REQUIRE(in0 && in1 && out && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad arguments");
    EFFDET_REQUIRE(mode0 >= 0 && mode0 <= 2 && (mode0 != 1 || (H % 2 == 0 && W % 2 == 0)), "bad mode");
    EFFDET_REQUIRE(a16(in0) && a16(in1) && (!in2 || a16(in2)) && a16(out), "16B alignment");
    cudaStream_t st = as_stream(stream);
    DISPATCH_T(dtype,
        ((void)launch_pdl(resample_fuse_kernel<float, 4>, dim3(grid_for((size_t)B * H * W * C / 4)), dim3(256), 0, st, 
            (const float *)in0, mode0, (const float *)in1, (const float *)in2, w, eps, (float *)out, B, H, W, C)),
        ((void)launch_pdl(resample_fuse_kernel<__nv_bfloat16, 8>, dim3(grid_for((size_t)B * H * W * C / 8)), dim3(256), 0, st, 
            (const __nv_bfloat16 *)in0, mode0, (const __nv_bfloat16 *)in1, (const __nv_bfloat16 *)in2, w,
            eps, (__nv_bfloat16 *)out, B, H, W, C)))
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

// number of row blocks the column reductions use for a (rows, C) matrix
extern "C" int effdet_colreduce_blocks(size_t rows, int C, int dtype) {
    if (rows == 0 || C <= 0) return 0;
    const int CV = dtype == EFFDET_BF16 ? 8 : 4;
    int nvec = C / CV; if (nvec < 1) nvec = 1;
    int PY = 256 / nvec; if (PY < 1) PY = 1;
    // ONE balanced wave: at most 4 blocks per SM (every reduction kernel fits 4 x 256 threads per SM), each
    // thread at least 4 rows.  More, smaller blocks ran as 1.3-2.6 waves whose last, partly filled wave cost
    // 10-30 % of the pass, and made the finalize kernels walk up to ~2300 partial rows.
    size_t rpb = cdiv(rows, (size_t)kNumSMs * 4);
    if (rpb < (size_t)PY * 4) rpb = (size_t)PY * 4;
    // a block's partial row is 2*C floats: at least 64 rows per block keeps the partial matrix (written here, read
    // by the finalize kernel) below 1/16 of the tensor (block7b of D4: 8192 rows x 2688 channels had 586 partial
    // rows = 12.6 MB for a 44 MB tensor, and a 14-us finalize)
    if (rpb < 64) rpb = 64;
    rpb = cdiv(rpb, PY) * PY;
    return (int)cdiv(rows, rpb);
}

template <typename T, int CV, int MODE>
static int launch_colreduce(const void *x, const void *y, const void *dy, const float *mean,
                            const float *invstd, size_t rows, int C, int nblk, float *partial,
                            cudaStream_t st) {
    const int nvec = C / CV;
    if (nvec > 1024) return fail(EFFDET_E_UNSUPPORTED, "colreduce: %sC=%lld too large", "", C);
    int PY = 256 / nvec; if (PY < 1) PY = 1;
    const int rpb = (int)cdiv(rows, nblk);
    if ((int)cdiv(rows, rpb) != nblk) return fail(EFFDET_E_INVALID, "colreduce: bad block count%s", "");
    const size_t sm = (size_t)PY * 2 * C * sizeof(float);
    if (sm > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(colreduce_kernel<T, CV, MODE>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return fail(EFFDET_E_CUDA, "colreduce: smem attribute: %s", cudaGetErrorString(e));
    }
    EFFDET_CUDA(launch_pdl(colreduce_kernel<T, CV, MODE>, dim3(nblk), dim3(nvec * PY), sm, st, 
        (const T *)x, (const T *)y, (const T *)dy, mean, invstd, rows, C, rpb, partial, 0u));
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

/* BN (training mode) forward statistics + finalize: z (rows,C) -> scale/shift/saved stats. */
extern "C" int effdet_bn_train_stats(const void *z, size_t rows, int C, const float *gamma,
                                     const float *beta, float eps, float momentum,
                                     float *moving_mean, float *moving_var, float *scale,
                                     float *shift, float *save_mean, float *save_invstd,
                                     float *partial, int nblk, int dtype, void *stream) {
    EFFDET_REQUIRE(z && gamma && beta && scale && shift && save_mean && save_invstd && partial, "null pointer");
    EFFDET_REQUIRE(rows > 0 && C > 0 && C % 8 == 0 && nblk > 0, "bad sizes");
    cudaStream_t st = as_stream(stream);
    int rc;
    DISPATCH_T(dtype,
        rc = (launch_colreduce<float, 4, RED_STATS>(z, nullptr, nullptr, nullptr, nullptr, rows, C, nblk, partial, st)),
        rc = (launch_colreduce<__nv_bfloat16, 8, RED_STATS>(z, nullptr, nullptr, nullptr, nullptr, rows, C, nblk, partial, st)))
    if (rc) return rc;
    EFFDET_CUDA(launch_pdl(bn_train_finalize_kernel, dim3(cdiv((size_t)C * 32, 256)), dim3(256), 0, st, partial, nblk, (double)rows, gamma, beta, eps,
                                                           momentum, moving_mean, moving_var, scale,
                                                           shift, save_mean, save_invstd, C));
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

/* Depthwise conv (raw output z, no activation) + the batch statistics of the BatchNormalization that follows
 * it, in one pass over the data: the depthwise kernel emits per-tile sums / sums of squares of its outputs
 * and only the finalize step of effdet_bn_train_stats runs afterwards (model.py:48-68 DepthwiseConvBlock in
 * training mode).  bf16 only.  partial: 2*C*nblk floats, nblk = B * effdet_dwconv_se_blocks(...). */
namespace effdet {
int dwconv_bf16_tma(const void *x, const float *w, const float *scale, const float *shift, void *y, float *se_sum,
                    int B, int H, int W, int C, int k, int stride, int act, cudaStream_t st, float *stats);
}
extern "C" int effdet_dwconv_bn_stats(const void *x, const float *kernel, const float *ones, const float *zeros,
                                      void *z, int B, int H, int W, int C, int k, int stride, const float *gamma,
                                      const float *beta, float eps, float momentum, float *moving_mean,
                                      float *moving_var, float *scale, float *shift, float *save_mean,
                                      float *save_invstd, float *partial, int nblk, int dtype, void *stream) {
    EFFDET_REQUIRE(x && kernel && ones && zeros && z && gamma && beta && scale && shift && save_mean && save_invstd &&
                       partial, "null pointer");
    EFFDET_REQUIRE(dtype == EFFDET_BF16, "bf16 only");
    EFFDET_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && (k == 3 || k == 5) && (stride == 1 || stride == 2),
                   "bad sizes");
    EFFDET_REQUIRE(nblk == B * effdet_dwconv_se_blocks(B, H, W, C, stride, dtype), "nblk must be B * effdet_dwconv_se_blocks()");
    cudaStream_t st = as_stream(stream);
    const int rc = dwconv_bf16_tma(x, kernel, ones, zeros, z, nullptr, B, H, W, C, k, stride, EFFDET_ACT_NONE, st, partial);
    if (rc) return rc;
    const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
    EFFDET_CUDA(launch_pdl(bn_train_finalize_kernel, dim3(cdiv((size_t)C * 32, 256)), dim3(256), 0, st, partial, nblk, (double)B * Ho * Wo, gamma, beta,
                                                                      eps, momentum, moving_mean, moving_var, scale,
                                                                      shift, save_mean, save_invstd, C));
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

extern "C" int effdet_scale_shift_act(const void *z, const float *scale, const float *shift, void *y,
                                      size_t rows, int C, int act, int dtype, void *stream) {
    EFFDET_REQUIRE(z && scale && shift && y && C > 0 && C % 8 == 0, "bad arguments");
    EFFDET_REQUIRE(rows * (size_t)C / 4 < 0xffffffffull, "tensor too large");
    if (rows == 0) return EFFDET_OK;
    cudaStream_t st = as_stream(stream);
    DISPATCH_T(dtype,
        ((void)launch_pdl(scale_shift_act_kernel<float, 4>, dim3(grid_for4(rows * C / 4)), dim3(256), 0, st, 
            (const float *)z, scale, shift, (float *)y, rows * C / 4, C, act, 0u)),
        ((void)launch_pdl(scale_shift_act_kernel<__nv_bfloat16, 8>, dim3(grid_for4(rows * C / 8)), dim3(256), 0, st, 
            (const __nv_bfloat16 *)z, scale, shift, (__nv_bfloat16 *)y, rows * C / 8, C, act, 0u)))
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

/* BN + ReLU backward (training-mode BN): dz from dy, y (ReLU output), z (BN input) and the saved
 * batch statistics; also dgamma/dbeta.  frozen != 0: BN ran in inference mode (k1 = scale). */
extern "C" int effdet_bn_relu_backward(const void *dy, const void *y, const void *z, size_t rows, int C,
                                       const float *gamma, const float *save_mean,
                                       const float *save_invstd, const float *frozen_scale,
                                       float *dgamma, float *dbeta, void *dz, float *k123,
                                       float *partial, int nblk, int dtype, void *stream) {
    EFFDET_REQUIRE(dy && y && z && dz && k123 && partial && gamma, "null pointer");
    EFFDET_REQUIRE(rows > 0 && C > 0 && C % 8 == 0 && nblk > 0, "bad sizes");
    cudaStream_t st = as_stream(stream);
    int rc;
    if (frozen_scale) {
        // inference-mode BN: dz = scale * dy_masked ; gamma/beta receive no gradient here
        EFFDET_CUDA(cudaMemsetAsync(k123, 0, 3 * (size_t)C * sizeof(float), st));
        EFFDET_CUDA(cudaMemcpyAsync(k123, frozen_scale, (size_t)C * sizeof(float),
                                    cudaMemcpyDeviceToDevice, st));
    } else {
        EFFDET_REQUIRE(save_mean && save_invstd, "null saved statistics");
        // small tensors (the coarse BiFPN levels, <= 512 KB): one cluster launch instead of reduce -> finalize ->
        // apply.  Measured on D0 training (batch 32, C = 64): 0.07 MB 8.2 -> 5.1 us, 0.26 MB 8.6 -> 6.2 us, 1 MB
        // 9.4 -> 9.2 us, 4 MB 13.7 -> 23 us, 17 MB 28 -> 78 us (64 CTAs cannot stream a large tensor), hence the
        // size limit.  EFFDET_BN_CLUSTER=0 keeps the three-kernel path.
        static const bool use_cluster = !(getenv("EFFDET_BN_CLUSTER") && atoi(getenv("EFFDET_BN_CLUSTER")) == 0);
        const int cvv = dtype == EFFDET_BF16 ? 8 : 4;
        const size_t tensor_bytes = rows * (size_t)C * (dtype == EFFDET_BF16 ? 2 : 4);
        if (use_cluster && (C / cvv) * kBnClu >= 32 && tensor_bytes <= (size_t)512 * 1024) {
            cudaError_t ce = dtype == EFFDET_BF16
                ? launch_bn_relu_bwd_cluster<__nv_bfloat16, 8>(st, dy, y, z, rows, C, gamma, save_mean, save_invstd,
                                                                dgamma, dbeta, dz, k123)
                : launch_bn_relu_bwd_cluster<float, 4>(st, dy, y, z, rows, C, gamma, save_mean, save_invstd, dgamma,
                                                       dbeta, dz, k123);
            EFFDET_CUDA(ce);
            EFFDET_LAUNCHED();
            return EFFDET_OK;
        }
        DISPATCH_T(dtype,
            rc = (launch_colreduce<float, 4, RED_BNBWD>(z, y, dy, save_mean, save_invstd, rows, C, nblk, partial, st)),
            rc = (launch_colreduce<__nv_bfloat16, 8, RED_BNBWD>(z, y, dy, save_mean, save_invstd, rows, C, nblk, partial, st)))
        if (rc) return rc;
        EFFDET_CUDA(launch_pdl(bn_bwd_finalize_kernel, dim3(cdiv((size_t)C * 32, 256)), dim3(256), 0, st, partial, nblk, (double)rows, gamma, save_mean,
                                                             save_invstd, k123, dgamma, dbeta, C));
        EFFDET_LAUNCHED();
    }
    DISPATCH_T(dtype,
        ((void)launch_pdl(bn_bwd_apply_kernel<float, 4>, dim3(grid_for4(rows * C / 4)), dim3(256), 0, st, 
            (const float *)dy, (const float *)y, (const float *)z, k123, (float *)dz, rows * C / 4, C, 0u)),
        ((void)launch_pdl(bn_bwd_apply_kernel<__nv_bfloat16, 8>, dim3(grid_for4(rows * C / 8)), dim3(256), 0, st, 
            (const __nv_bfloat16 *)dy, (const __nv_bfloat16 *)y, (const __nv_bfloat16 *)z, k123,
            (__nv_bfloat16 *)dz, rows * C / 8, C, 0u)))
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

/* out[c] += sum over rows [row0, row0 + nrows) of x[r][c] for a dense (*, C) matrix: scalar tail of
 * effdet_colsum for the few rows that do not fill a whole vector-aligned fold. */
template <typename T>
__global__ void colsum_tail_kernel(const T *__restrict__ x, size_t row0, int nrows, int C, float *__restrict__ out) {
    EFFDET_PDL_SYNC();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float t = 0.f;
    for (int r = 0; r < nrows; ++r) t += to_f<T>(x[(row0 + r) * (size_t)C + c]);
    out[c] += t;
}
extern "C" int effdet_colsum_tail(const void *x, size_t row0, int nrows, int C, float *out, int dtype, void *stream) {
    EFFDET_REQUIRE(x && out && C > 0 && nrows >= 0, "bad arguments");
    if (nrows == 0) return EFFDET_OK;
    cudaStream_t st = as_stream(stream);
    DISPATCH_T(dtype,
        ((void)launch_pdl(colsum_tail_kernel<float>, dim3(cdiv(C, 128)), dim3(128), 0, st, (const float *)x, row0, nrows, C, out)),
        ((void)launch_pdl(colsum_tail_kernel<__nv_bfloat16>, dim3(cdiv(C, 128)), dim3(128), 0, st, (const __nv_bfloat16 *)x, row0, nrows, C, out)))
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

/* column sum of a dense (rows*fold, C/fold) matrix -> out[C/fold] (bias gradients); the caller
 * passes it viewed as (rows, C) with C a multiple of the vector width. */
extern "C" int effdet_colsum(const void *x, size_t rows, int C, int fold, float *out, int accumulate,
                             float *partial, int nblk, int dtype, void *stream) {
    EFFDET_REQUIRE(x && out && partial && rows > 0 && C > 0 && nblk > 0 && fold >= 1, "bad arguments");
    EFFDET_REQUIRE(C % fold == 0, "fold must divide C");
    EFFDET_REQUIRE(C % (dtype == EFFDET_BF16 ? 8 : 4) == 0, "C must be a multiple of the vector width");
    cudaStream_t st = as_stream(stream);
    int rc;
    DISPATCH_T(dtype,
        rc = (launch_colreduce<float, 4, RED_SUM>(x, nullptr, nullptr, nullptr, nullptr, rows, C, nblk, partial, st)),
        rc = (launch_colreduce<__nv_bfloat16, 8, RED_SUM>(x, nullptr, nullptr, nullptr, nullptr, rows, C, nblk, partial, st)))
    if (rc) return rc;
    EFFDET_CUDA(launch_pdl(colsum_finalize_kernel, dim3(cdiv((size_t)(C / fold) * 32, 256)), dim3(256), 0, st, partial, nblk, C, fold, out, accumulate));
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

namespace effdet {      // dwconv_tma.cu
int dw_wgrad_bf16_splits(int B, int H, int W, int C, int stride);
int dw_wgrad_bf16_tma(const void *x, const void *dz, float *partial, int nsplit, int B, int H, int W, int C, int k,
                      int stride, cudaStream_t st);
}

extern "C" int effdet_dw_wgrad_blocks(int B, int H, int W, int C, int dtype) {
    if (dtype == EFFDET_BF16) return dw_wgrad_bf16_splits(B, H, W, C, 1);
    const int CV = dtype == EFFDET_BF16 ? 8 : 4;
    int nvec = C / CV; if (nvec < 1) nvec = 1;
    int PY = 128 / nvec; if (PY < 1) PY = 1;
    const size_t total = (size_t)B * H * W;
    size_t ppb = (size_t)PY * 32;
    while (ppb > (size_t)PY && cdiv(total, ppb) < (unsigned)kNumSMs * 2) ppb >>= 1;
    return (int)cdiv(total, ppb);
}

/* depthwise 3x3 (stride 1, SAME) weight gradient: dkernel (3,3,C) = sum f (*) dz. */
extern "C" int effdet_dw_wgrad(const void *f, const void *dz, int B, int H, int W, int C,
                               float *dkernel, float *partial, int nblk, int dtype, void *stream) {
    EFFDET_REQUIRE(f && dz && dkernel && partial && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && nblk > 0,
                   "bad arguments");
    cudaStream_t st = as_stream(stream);
    if (dtype == EFFDET_BF16) {
        // TMA-tiled kernel shared with the backbone's depthwise backward (dwconv_tma.cu)
        EFFDET_REQUIRE(nblk == effdet_dw_wgrad_blocks(B, H, W, C, dtype), "nblk must come from effdet_dw_wgrad_blocks");
        const int rc = dw_wgrad_bf16_tma(f, dz, partial, nblk, B, H, W, C, 3, 1, st);
        if (rc) return rc;
        EFFDET_CUDA(launch_pdl(sum_partials_kernel, dim3(cdiv((size_t)9 * C * 32, 256)), dim3(256), 0, st, partial, nblk, 9 * C, dkernel, 0));
        EFFDET_LAUNCHED();
        return EFFDET_OK;
    }
    const int CV = dtype == EFFDET_BF16 ? 8 : 4;
    const int nvec = C / CV;
    EFFDET_REQUIRE(nvec <= 1024, "C too large");
    int PY = 128 / nvec; if (PY < 1) PY = 1;
    const size_t total = (size_t)B * H * W;
    const int ppb = (int)cdiv(total, nblk);
    EFFDET_REQUIRE((int)cdiv(total, ppb) == nblk, "nblk must come from effdet_dw_wgrad_blocks");
    const size_t sm = (size_t)PY * 9 * C * sizeof(float);
    if (dtype == EFFDET_F32) {
        if (sm > 48 * 1024) EFFDET_CUDA(cudaFuncSetAttribute(dw_wgrad_kernel<float, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        EFFDET_CUDA(launch_pdl(dw_wgrad_kernel<float, 4>, dim3(nblk), dim3(nvec * PY), sm, st, (const float *)f, (const float *)dz, B, H, W, C, ppb, partial));
    } else if (dtype == EFFDET_BF16) {
        if (sm > 48 * 1024) EFFDET_CUDA(cudaFuncSetAttribute(dw_wgrad_kernel<__nv_bfloat16, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        EFFDET_CUDA(launch_pdl(dw_wgrad_kernel<__nv_bfloat16, 8>, dim3(nblk), dim3(nvec * PY), sm, st, (const __nv_bfloat16 *)f, (const __nv_bfloat16 *)dz, B, H, W, C, ppb, partial));
    } else {
        return fail(EFFDET_E_INVALID, "effdet_dw_wgrad: bad dtype%s", "");
    }
    EFFDET_LAUNCHED();
    EFFDET_CUDA(launch_pdl(sum_partials_kernel, dim3(cdiv((size_t)9 * C * 32, 256)), dim3(256), 0, st, partial, nblk, 9 * C, dkernel, 0));
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

/* gradient of one fusion input.  which = 0 routes through the resampling of in0 (mode0). */
extern "C" int effdet_fuse_backward_input(const void *df, int which, int mode0, const void *in0,
                                          const float *w, int n_inputs, float eps, void *dst,
                                          int accumulate, int B, int H, int W, int C, int dtype,
                                          void *stream) {
    EFFDET_REQUIRE(df && dst && which >= 0 && which < n_inputs && (n_inputs == 2 || n_inputs == 3), "bad arguments");
    EFFDET_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad sizes");
    EFFDET_REQUIRE(!(which == 0 && mode0 == 2) || in0, "max-pool routing needs the forward input");
    EFFDET_REQUIRE(a16(df) && a16(dst), "16B alignment");
    cudaStream_t st = as_stream(stream);
    size_t n = (size_t)B * H * W * C;
    if (which == 0 && mode0 == 1) n /= 4;
    if (which == 0 && mode0 == 2) n *= 4;
    DISPATCH_T(dtype,
        ((void)launch_pdl(fuse_bwd_kernel<float, 4>, dim3(grid_for(n / 4)), dim3(256), 0, st, 
            (const float *)df, which, mode0, (const float *)in0, w, n_inputs, eps, (float *)dst, accumulate, B, H, W, C)),
        ((void)launch_pdl(fuse_bwd_kernel<__nv_bfloat16, 8>, dim3(grid_for(n / 8)), dim3(256), 0, st, 
            (const __nv_bfloat16 *)df, which, mode0, (const __nv_bfloat16 *)in0, w, n_inputs, eps,
            (__nv_bfloat16 *)dst, accumulate, B, H, W, C)))
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

/* gradient of the fusion weights (layers.py:26-31): dw (n_inputs) f32; partial: 4*1184 floats. */
extern "C" int effdet_fuse_backward_weights(const void *df, const void *f, const void *in0, int mode0,
                                            const void *in1, const void *in2, const float *w,
                                            float eps, float *dw, float *partial, int B, int H, int W,
                                            int C, int dtype, void *stream) {
    EFFDET_REQUIRE(df && f && in0 && in1 && w && dw && partial, "null pointer");
    EFFDET_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad sizes");
    cudaStream_t st = as_stream(stream);
    const int n_in = in2 ? 3 : 2;
    const size_t n = (size_t)B * H * W * C;
    unsigned nblk;
    DISPATCH_T(dtype,
        (nblk = grid_for(n / 4), (void)launch_pdl(fuse_wgrad_kernel<float, 4>, dim3(nblk), dim3(256), 0, st, 
            (const float *)df, (const float *)f, (const float *)in0, mode0, (const float *)in1,
            (const float *)in2, B, H, W, C, partial)),
        (nblk = grid_for(n / 8), (void)launch_pdl(fuse_wgrad_kernel<__nv_bfloat16, 8>, dim3(nblk), dim3(256), 0, st, 
            (const __nv_bfloat16 *)df, (const __nv_bfloat16 *)f, (const __nv_bfloat16 *)in0, mode0,
            (const __nv_bfloat16 *)in1, (const __nv_bfloat16 *)in2, B, H, W, C, partial)))
    EFFDET_LAUNCHED();
    EFFDET_CUDA(launch_pdl(fuse_wgrad_finalize_kernel, dim3(1), dim3(32), 0, st, partial, (int)nblk, w, n_in, eps, dw));
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

/* keras SGD (train_tpu.py:268-269): v = momentum*v - lr_t*(g*grad_scale); w += v, with
 * lr_t = lr / (1 + decay*iterations) computed by the caller. */
/* Zero insertion for the data gradient of a stride-2 convolution: out (B,H,W,C) holds dz (B,Ho,Wo,C) at
 * rows 2*oy + ay / columns 2*ox + ax and zeros elsewhere (ay = (k-1)/2 - pad_top of the forward conv). */
extern "C" int effdet_zero_insert(const void *dz, void *out, int B, int Ho, int Wo, int C, int H, int W, int ay,
                                  int ax, int dtype, void *stream) {
    EFFDET_REQUIRE(dz && out && B > 0 && Ho > 0 && Wo > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad arguments");
    EFFDET_REQUIRE(a16(dz) && a16(out), "16B alignment");
    cudaStream_t st = as_stream(stream);
    const size_t n = (size_t)B * H * W * C;
    DISPATCH_T(dtype,
        ((void)launch_pdl(zero_insert_kernel<float, 4>, dim3(grid_for(n / 4)), dim3(256), 0, st, (const float *)dz, (float *)out, B, Ho, Wo, C, H, W, ay, ax)),
        ((void)launch_pdl(zero_insert_kernel<__nv_bfloat16, 8>, dim3(grid_for(n / 8)), dim3(256), 0, st, (const __nv_bfloat16 *)dz, (__nv_bfloat16 *)out, B, Ho, Wo, C, H, W, ay, ax)))
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

extern "C" int effdet_sgd_momentum_step_dev_lr(float *w, const float *g, float *v, size_t n, const float *lr_dev,
                                               float momentum, float grad_scale, void *stream) {
    if (n == 0) return EFFDET_OK;
    EFFDET_REQUIRE(w && g && v && lr_dev, "null pointer");
    EFFDET_CUDA(launch_pdl(sgd_momentum_dev_lr_kernel, dim3(grid_for(n)), dim3(256), 0, as_stream(stream), w, g, v, n, lr_dev, momentum, grad_scale));
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

extern "C" int effdet_sgd_momentum_step(float *w, const float *g, float *v, size_t n, float lr_t,
                                        float momentum, float grad_scale, void *stream) {
    if (n == 0) return EFFDET_OK;
    EFFDET_REQUIRE(w && g && v, "null pointer");
    EFFDET_CUDA(launch_pdl(sgd_momentum_kernel, dim3(grid_for(n)), dim3(256), 0, as_stream(stream), w, g, v, n, lr_t, momentum, grad_scale));
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

extern "C" int effdet_flip_taps(const float *in, float *out, int taps, int n, void *stream) {
    EFFDET_REQUIRE(in && out && taps > 0 && n > 0, "bad arguments");
    EFFDET_CUDA(launch_pdl(flip_taps_kernel, dim3(cdiv((size_t)taps * n, 256)), dim3(256), 0, as_stream(stream), in, out, taps, n));
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

extern "C" int effdet_conv_weight_transpose(const float *in, float *out, int taps, int Cin, int Cout,
                                            void *stream) {
    EFFDET_REQUIRE(in && out && taps > 0 && Cin > 0 && Cout > 0, "bad arguments");
    EFFDET_CUDA(launch_pdl(conv_weight_transpose_kernel, dim3(cdiv((size_t)taps * Cin * Cout, 256)), dim3(256), 0, as_stream(stream), 
        in, out, taps, Cin, Cout));
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}
