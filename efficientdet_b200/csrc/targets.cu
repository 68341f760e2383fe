// Geometry: anchor table (host, float64), IoU matrix and anchor-target assignment (device,
// float64).  Compiled with -fmad=false so the double arithmetic is bit-identical to the
// reference's numpy / Cython code (no FMA contraction).
//
// Replaces (reference file:line):
//   utils/anchors.py:372-403, :339-369, :296-336   -> effdet_anchors_for_shape_host
//   utils/compute_overlap.pyx:13-53                -> overlap_kernel
//   utils/anchors.py:130-239, :406-439             -> anchor_targets_kernel
#include <math.h>

#include "common.cuh"

namespace effdet {

__device__ __forceinline__ double iou_plus1(double b0, double b1, double b2, double b3,
                                            double q0, double q1, double q2, double q3) {
    // same op order as compute_overlap.pyx:31-52
    double box_area = (q2 - q0 + 1) * (q3 - q1 + 1);
    double iw = fmin(b2, q2) - fmax(b0, q0) + 1;
    if (iw > 0) {
        double ih = fmin(b3, q3) - fmax(b1, q1) + 1;
        if (ih > 0) {
            double ua = (b2 - b0 + 1) * (b3 - b1 + 1) + box_area - iw * ih;
            return iw * ih / ua;
        }
    }
    return 0.0;
}

__global__ void __launch_bounds__(256)
overlap_kernel(const double *__restrict__ boxes, size_t N, const double *__restrict__ query,
               size_t K, double *__restrict__ out) {
    size_t total = N * K, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        size_t n = i / K, k = i % K;
        const double *b = boxes + 4 * n, *q = query + 4 * k;
        out[i] = iou_plus1(b[0], b[1], b[2], b[3], q[0], q[1], q[2], q[3]);
    }
}

// One thread per (image, anchor).  Dense labels are pre-zeroed by a memset; this kernel
// scatters the one-hot entry and the state column, and writes the 5-float regression row.
__global__ void __launch_bounds__(256)
anchor_targets_kernel(const double *__restrict__ anchors, size_t N,
                      const double *__restrict__ gt_boxes, const int32_t *__restrict__ gt_labels,
                      const int32_t *__restrict__ gt_counts, int Kmax,
                      const double *__restrict__ image_hw, int C, double neg_ov, double pos_ov,
                      float *__restrict__ regression, float *__restrict__ labels,
                      int8_t *__restrict__ cstate, int32_t *__restrict__ ccls) {
    extern __shared__ double sgt[];      // Kmax * 4 boxes of this image
    const int b = blockIdx.y;
    const int K = gt_counts[b];
    for (int i = threadIdx.x; i < K * 4; i += blockDim.x) sgt[i] = gt_boxes[(size_t)b * Kmax * 4 + i];
    __syncthreads();
    const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const double a0 = anchors[4 * n], a1 = anchors[4 * n + 1], a2 = anchors[4 * n + 2],
                 a3 = anchors[4 * n + 3];
    float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f, state = 0.f;
    int cls = -1;
    if (K > 0) {
        double best = -1.0;
        int arg = 0;
        for (int k = 0; k < K; ++k) {          // np.argmax: first maximum wins
            double ov = iou_plus1(a0, a1, a2, a3, sgt[4 * k], sgt[4 * k + 1], sgt[4 * k + 2],
                                  sgt[4 * k + 3]);
            if (ov > best) { best = ov; arg = k; }
        }
        const bool pos = best >= pos_ov;
        const bool ign = (best > neg_ov) && !pos;
        state = pos ? 1.f : (ign ? -1.f : 0.f);
        if (pos) cls = gt_labels[(size_t)b * Kmax + arg];
        const double aw = a2 - a0, ah = a3 - a1;
        // bbox_transform: (gt - a) / wh, then (t - mean[0]) / std[0.2] in float64
        r0 = (float)(((sgt[4 * arg] - a0) / aw - 0.0) / 0.2);
        r1 = (float)(((sgt[4 * arg + 1] - a1) / ah - 0.0) / 0.2);
        r2 = (float)(((sgt[4 * arg + 2] - a2) / aw - 0.0) / 0.2);
        r3 = (float)(((sgt[4 * arg + 3] - a3) / ah - 0.0) / 0.2);
    }
    const double H = image_hw[2 * b], W = image_hw[2 * b + 1];
    if (H >= 0) {
        double cx = (a0 + a2) / 2, cy = (a1 + a3) / 2;
        if (cx >= W || cy >= H) state = -1.f;
    }
    const size_t row = (size_t)b * N + n;
    if (regression) {
        float *r = regression + row * 5;
        r[0] = r0; r[1] = r1; r[2] = r2; r[3] = r3; r[4] = state;
    }
    if (labels) {
        float *l = labels + row * (size_t)(C + 1);
        if (cls >= 0 && cls < C) l[cls] = 1.f;
        l[C] = state;
    }
    if (cstate) cstate[row] = (int8_t)state;
    if (ccls) ccls[row] = cls;
}

}  // namespace effdet

using namespace effdet;

extern "C" int effdet_anchors_for_shape_host(const int *level_hw, const int *sizes,
                                             const int *strides, int n_levels,
                                             const double *ratios, int n_ratios,
                                             const double *scales, int n_scales, double *out,
                                             size_t out_capacity_rows) {
    EFFDET_REQUIRE(level_hw && sizes && strides && ratios && scales && out, "null pointer");
    EFFDET_REQUIRE(n_levels >= 0 && n_ratios > 0 && n_scales > 0, "bad sizes");
    const int A = n_ratios * n_scales;
    size_t need = 0;
    for (int l = 0; l < n_levels; ++l) need += (size_t)level_hw[2 * l] * level_hw[2 * l + 1] * A;
    if (need > out_capacity_rows)
        return fail(EFFDET_E_CAPACITY, "effdet_anchors_for_shape_host: need %s%lld rows, have %lld",
                    "", (long long)need, (long long)out_capacity_rows);
    double *o = out;
    for (int l = 0; l < n_levels; ++l) {
        const int H = level_hw[2 * l], W = level_hw[2 * l + 1];
        double base[64][4];
        EFFDET_REQUIRE(A <= 64, "too many anchors per cell");
        for (int r = 0; r < n_ratios; ++r)
            for (int s = 0; s < n_scales; ++s) {
                // utils/anchors.py:391: base_size * scales is evaluated in float32
                volatile float side_f = (float)scales[s] * (float)sizes[l];
                double side = (double)side_f;
                double area = side * side;
                double w = sqrt(area / ratios[r]);
                double h = w * ratios[r];
                double *bb = base[r * n_scales + s];
                bb[0] = 0.0 - w * 0.5; bb[1] = 0.0 - h * 0.5;
                bb[2] = w - w * 0.5;   bb[3] = h - h * 0.5;
            }
        for (int y = 0; y < H; ++y) {
            double sy = (y + 0.5) * strides[l];
            for (int x = 0; x < W; ++x) {
                double sx = (x + 0.5) * strides[l];
                for (int a = 0; a < A; ++a) {
                    o[0] = base[a][0] + sx; o[1] = base[a][1] + sy;
                    o[2] = base[a][2] + sx; o[3] = base[a][3] + sy;
                    o += 4;
                }
            }
        }
    }
    return EFFDET_OK;
}

extern "C" int effdet_compute_overlap(const double *boxes, size_t N, const double *query, size_t K,
                                      double *overlaps, void *stream) {
    if (N * K == 0) return EFFDET_OK;
    EFFDET_REQUIRE(boxes && query && overlaps, "null pointer");
    unsigned blocks = cdiv(N * K, 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    overlap_kernel<<<blocks, 256, 0, as_stream(stream)>>>(boxes, N, query, K, overlaps);
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

extern "C" int effdet_anchor_targets(const double *anchors, size_t N, const double *gt_boxes,
                                     const int32_t *gt_labels, const int32_t *gt_counts, int B,
                                     int Kmax, const double *image_hw, int num_classes,
                                     double negative_overlap, double positive_overlap,
                                     float *regression, float *labels, int8_t *compact_state,
                                     int32_t *compact_cls, void *stream) {
    EFFDET_REQUIRE(B >= 1, "No data received to compute anchor targets for.");
    EFFDET_REQUIRE(anchors && gt_counts && image_hw, "null pointer");
    EFFDET_REQUIRE(Kmax >= 0 && num_classes >= 1, "bad sizes");
    EFFDET_REQUIRE(Kmax == 0 || (gt_boxes && gt_labels), "null gt pointer");
    EFFDET_REQUIRE((size_t)Kmax * 32 <= 48 * 1024, "too many boxes per image (max 1536)");
    if (N == 0) return EFFDET_OK;
    cudaStream_t st = as_stream(stream);
    if (labels)
        EFFDET_CUDA(cudaMemsetAsync(labels, 0, (size_t)B * N * (num_classes + 1) * sizeof(float), st));
    dim3 grid(cdiv(N, 256), B);
    anchor_targets_kernel<<<grid, 256, (size_t)(Kmax > 0 ? Kmax : 1) * 32, st>>>(
        anchors, N, gt_boxes, gt_labels, gt_counts, Kmax, image_hw, num_classes, negative_overlap,
        positive_overlap, regression, labels, compact_state, compact_cls);
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}
