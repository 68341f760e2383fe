/* Oracle (TEST INFRASTRUCTURE): plain-C restatement of
 *   /root/reference/utils/compute_overlap.pyx:13-53   (IoU, "+1" convention, float64)
 *   /root/reference/utils/anchors.py:210-239          (argmax / positive / ignore)
 *   /root/reference/utils/anchors.py:130-207,406-439  (dense targets, bbox_transform)
 * Single-threaded like the reference's Cython loop, so it doubles as the CPU
 * baseline for the target-assignment step.  Build: oracle/build_ref.py.
 */
#include <stddef.h>

static double dmin(double a, double b) { return a < b ? a : b; }
static double dmax(double a, double b) { return a > b ? a : b; }

/* overlaps is (N,K) row-major, pre-zeroed by the caller or not: we write every cell */
void oracle_compute_overlap(const double *boxes, size_t N, const double *query, size_t K,
                            double *overlaps) {
    for (size_t k = 0; k < K; ++k) {
        const double *q = query + 4 * k;
        double box_area = (q[2] - q[0] + 1) * (q[3] - q[1] + 1);
        for (size_t n = 0; n < N; ++n) {
            const double *b = boxes + 4 * n;
            double v = 0.0;
            double iw = dmin(b[2], q[2]) - dmax(b[0], q[0]) + 1;
            if (iw > 0) {
                double ih = dmin(b[3], q[3]) - dmax(b[1], q[1]) + 1;
                if (ih > 0) {
                    double ua = (b[2] - b[0] + 1) * (b[3] - b[1] + 1) + box_area - iw * ih;
                    v = iw * ih / ua;
                }
            }
            overlaps[n * K + k] = v;
        }
    }
}

/* One image.  anchors (N,4) f64; gt (K,4) f32-valued doubles; gt_labels (K) int.
 * regression (N,5) f32, labels (N,C+1) f32, both pre-zeroed by the caller.
 * img_h/img_w < 0 => the reference's `if image.shape:` branch is skipped. */
void oracle_anchor_targets_image(const double *anchors, size_t N, const double *gt,
                                 const int *gt_labels, size_t K, int num_classes,
                                 double neg_ov, double pos_ov, double img_h, double img_w,
                                 float *regression, float *labels) {
    size_t C1 = (size_t)num_classes + 1;
    for (size_t n = 0; n < N; ++n) {
        const double *a = anchors + 4 * n;
        float *r = regression + 5 * n;
        float *l = labels + C1 * n;
        if (K > 0) {
            double best = -1.0; size_t arg = 0;
            for (size_t k = 0; k < K; ++k) {       /* np.argmax: first maximum */
                double ov;
                oracle_compute_overlap(a, 1, gt + 4 * k, 1, &ov);
                if (ov > best) { best = ov; arg = k; }
            }
            int pos = best >= pos_ov;
            int ign = (best > neg_ov) && !pos;
            float state = pos ? 1.0f : (ign ? -1.0f : 0.0f);
            r[4] = state; l[num_classes] = state;
            if (pos) l[gt_labels[arg]] = 1.0f;
            const double *g = gt + 4 * arg;
            double aw = a[2] - a[0], ah = a[3] - a[1];
            r[0] = (float)(((g[0] - a[0]) / aw - 0) / 0.2);
            r[1] = (float)(((g[1] - a[1]) / ah - 0) / 0.2);
            r[2] = (float)(((g[2] - a[2]) / aw - 0) / 0.2);
            r[3] = (float)(((g[3] - a[3]) / ah - 0) / 0.2);
        }
        if (img_h >= 0) {
            double cx = (a[0] + a[2]) / 2, cy = (a[1] + a[3]) / 2;
            if (cx >= img_w || cy >= img_h) { r[4] = -1.0f; l[num_classes] = -1.0f; }
        }
    }
}
