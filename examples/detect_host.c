/* A host WITHOUT Python or torch driving the plan-level C ABI (include/effdet_b200.h): the C equivalent of
 *
 *     model, prediction_model = efficientdet(phi, num_classes=C, weighted_bifpn=..., anchors=anchors_for_shape(...))
 *     prediction_model.load_weights(path, by_name=True)
 *     boxes, scores, labels = prediction_model.predict_on_batch([images])        (reference: inference.py:32-59,
 *                                                                                  predict.py:74-105, model.py:356-452)
 *
 *     gcc -std=c99 -O2 -I include examples/detect_host.c -L efficientdet_b200 -leffdet_b200 \
 *         -Wl,-rpath,$PWD/efficientdet_b200 -lm -o detect_host
 *     ./detect_host [phi=0] [image_size=512] [batch=2] [classes=20]
 *
 * Weights: the plan's own manifest is walked (Keras names and shapes, what load_weights(by_name=True) matches) and
 * every tensor is filled deterministically the way the reference initialises it (kernels Glorot-uniform, BatchNorm
 * identity, class-head bias = -log((1 - 0.01) / 0.01), initializers.py:24) -- a real host reads them from the
 * reference's .h5 checkpoint instead.  Images: synthetic raw uint8 RGB (EFFDET_PLAN_U8_INPUT: normalize_image runs
 * inside the stem kernel).  Exit codes: 0 ok, 2 usage, 3 the library reported an error (e.g. no CUDA device: the
 * library has no CPU fallback and says so). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "effdet_b200.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;

static float uniform01(void) {           /* xorshift64*, 24 random bits -> [0, 1) */
    rng_state ^= rng_state >> 12;
    rng_state ^= rng_state << 25;
    rng_state ^= rng_state >> 27;
    return (float)((rng_state * 0x2545F4914F6CDD1Dull) >> 40) * (1.0f / 16777216.0f);
}

static int ends_with(const char *s, const char *suffix) {
    size_t n = strlen(s), m = strlen(suffix);
    return n >= m && strcmp(s + n - m, suffix) == 0;
}

static int check(int rc, const char *what) {
    if (rc != EFFDET_OK) {
        fprintf(stderr, "%s failed (%d): %s\n", what, rc, effdet_last_error());
        return 1;
    }
    return 0;
}

static void fill_weight(const char *name, int ndim, const int *dims, float *w, size_t n) {
    size_t i;
    if (ends_with(name, "/gamma") || ends_with(name, "/moving_variance")) {
        for (i = 0; i < n; ++i) w[i] = 1.0f;
    } else if (ends_with(name, "/beta") || ends_with(name, "/moving_mean")) {
        for (i = 0; i < n; ++i) w[i] = 0.0f;
    } else if (ends_with(name, "/bias")) {
        /* model.py:330-345: the classification convolution starts at PriorProbability(0.01) */
        float b = strstr(name, "class_head/pyramid_classification") ? -4.59512f : 0.0f;
        for (i = 0; i < n; ++i) w[i] = b;
    } else if (ndim == 4) {              /* conv HWIO / depthwise HWC1 */
        double fan_in = (double)dims[0] * dims[1] * (dims[3] == 1 ? 1 : dims[2]);
        double fan_out = (double)dims[0] * dims[1] * (dims[3] == 1 ? 1 : dims[3]);
        float a = (float)sqrt(6.0 / (fan_in + fan_out));
        for (i = 0; i < n; ++i) w[i] = (2.0f * uniform01() - 1.0f) * a;
    } else {                             /* wBiFPNAdd fusion weights (layers.py:17-23): ones; SE / dense vectors */
        for (i = 0; i < n; ++i) w[i] = 1.0f;
    }
}

int main(int argc, char **argv) {
    int phi = argc > 1 ? atoi(argv[1]) : 0;
    int size = argc > 2 ? atoi(argv[2]) : 512;
    int batch = argc > 3 ? atoi(argv[3]) : 2;
    int classes = argc > 4 ? atoi(argv[4]) : 20;
    const int max_det = 300;
    effdet_plan_t *plan = NULL;
    int nw, i, b, rc = 0;
    const char **names;
    void **ptrs;
    uint8_t *images;
    float *boxes, *scores;
    int32_t *labels;
    size_t npix;

    if (phi < 0 || phi > 6 || size <= 0 || batch <= 0 || classes <= 0) {
        fprintf(stderr, "usage: %s [phi 0..6] [image_size] [batch] [classes]\n", argv[0]);
        return 2;
    }
    if (check(effdet_plan_create(phi, size, batch, classes, /*weighted_bifpn=*/1, EFFDET_BF16, EFFDET_PLAN_U8_INPUT,
                                 &plan), "effdet_plan_create"))
        return 3;
    nw = effdet_plan_num_weights(plan);
    printf("EfficientDet-D%d %dx%d batch %d, %d classes: %d weights, %lu anchors, %d launches per forward\n", phi, size,
           size, batch, classes, nw, (unsigned long)effdet_plan_num_anchors(plan), effdet_plan_num_launches(plan));

    names = (const char **)calloc((size_t)nw, sizeof(*names));
    ptrs = (void **)calloc((size_t)nw, sizeof(*ptrs));
    for (i = 0; i < nw && !rc; ++i) {
        int nd = 0, dims[4] = {1, 1, 1, 1}, d;
        size_t n = 1;
        rc = check(effdet_plan_weight_info(plan, i, &names[i], &nd, dims), "effdet_plan_weight_info");
        for (d = 0; d < nd; ++d) n *= (size_t)dims[d];
        ptrs[i] = malloc(n * sizeof(float));
        fill_weight(names[i], nd, dims, (float *)ptrs[i], n);
    }
    if (!rc)
        rc = check(effdet_plan_bind_weights_host(plan, names, (const void *const *)ptrs, nw),
                   "effdet_plan_bind_weights_host");

    npix = (size_t)batch * size * size * 3;
    images = (uint8_t *)malloc(npix);
    for (npix = 0; npix < (size_t)batch * size * size * 3; ++npix) images[npix] = (uint8_t)(uniform01() * 256.0f);
    boxes = (float *)malloc((size_t)batch * max_det * 4 * sizeof(float));
    scores = (float *)malloc((size_t)batch * max_det * sizeof(float));
    labels = (int32_t *)malloc((size_t)batch * max_det * sizeof(int32_t));
    if (!rc)
        rc = check(effdet_detect_host(plan, images, /*anchors=*/NULL, /*score_threshold=*/0.005f,
                                      /*iou_threshold=*/0.5f, max_det, boxes, scores, labels), "effdet_detect_host");
    for (b = 0; b < batch && !rc; ++b) {
        int kept = 0, k;
        double sum = 0.0;
        for (k = 0; k < max_det; ++k)
            if (labels[b * max_det + k] >= 0) {
                ++kept;
                sum += scores[b * max_det + k];
            }
        printf("image %d: %d detections, score sum %.6f", b, kept, sum);
        if (kept) {
            const float *bx = boxes + (size_t)b * max_det * 4;
            printf(", best: class %d score %.4f box [%.1f %.1f %.1f %.1f]", (int)labels[b * max_det], scores[b * max_det],
                   bx[0], bx[1], bx[2], bx[3]);
        }
        printf("\n");
        for (k = 1; k < kept; ++k)       /* FilterDetections.py:97-99: sorted by score, padding (-1) at the end */
            if (scores[b * max_det + k] > scores[b * max_det + k - 1]) {
                fprintf(stderr, "detections of image %d are not sorted by score\n", b);
                rc = 1;
            }
    }
    effdet_plan_destroy(plan);
    for (i = 0; i < nw; ++i) free(ptrs[i]);
    free(ptrs); free((void *)names); free(images); free(boxes); free(scores); free(labels);
    return rc ? 3 : 0;
}
