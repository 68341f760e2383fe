/* A host WITHOUT Python driving the training half of the plan-level C ABI (include/effdet_b200.h, "compiled plans"):
 * the C equivalent of the reference's per-replica training loop
 *
 *     model.compile(optimizer=SGD(lr, decay, momentum), loss={'regression': smooth_l1, 'classification': focal})
 *     model.fit(dataset, ...)                                      (reference: train_tpu.py:249-346, train.py:333-388)
 *
 * The launch list of one optimizer step (anchor targets -> forward in training mode -> focal + smooth-L1 -> backward ->
 * SGD) is written once by efficientdet_b200.plan_export.export_train_plan(model, batch, "plan.efd", u8_input=True)
 * (raw uint8 RGB images, normalize_image on the device); this program loads it, fills the input regions with a
 * synthetic batch and steps it.  tests/test_c_host_example.py compiles and links it and checks its error paths on the
 * CPU; the same entry points run on the GPU through ctypes in tests/test_gpu_replay.py (bit-identical to
 * Trainer.train_on_batch).
 *
 *     gcc -std=c99 -O2 -I include -isystem /usr/local/cuda/include examples/train_host.c -L efficientdet_b200 -leffdet_b200 \
 *         -Wl,-rpath,$PWD/efficientdet_b200 -L /usr/local/cuda/lib64 -lcudart -o train_host
 *     ./train_host plan.efd [steps=10] [lr=0.01] [decay=4e-5]
 *
 * Exit codes: 0 ok, 2 usage, 3 the library or the CUDA runtime reported an error. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "effdet_b200.h"

static int check(int rc, const char *what) {
    if (rc != EFFDET_OK) {
        fprintf(stderr, "%s failed (%d): %s\n", what, rc, effdet_last_error());
        return 1;
    }
    return 0;
}

static int cuda_check(cudaError_t e, const char *what) {
    if (e != cudaSuccess) {
        fprintf(stderr, "%s failed: %s\n", what, cudaGetErrorString(e));
        return 1;
    }
    return 0;
}

int main(int argc, char **argv) {
    effdet_replay_t *plan = NULL;
    void *images = NULL, *gt_boxes = NULL, *gt_labels = NULL, *gt_counts = NULL, *losses = NULL;
    size_t n_images = 0, n_boxes = 0, n_labels = 0, n_counts = 0, n_losses = 0, i;
    long steps, it;
    double lr, decay;
    unsigned char *h_images;
    double *h_boxes;
    int32_t *h_labels, *h_counts;
    float h_losses[8];
    size_t batch, kmax;
    int rc = 0;

    if (argc < 2) {
        fprintf(stderr, "usage: %s plan.efd [steps] [lr] [decay]\n", argv[0]);
        return 2;
    }
    steps = argc > 2 ? atol(argv[2]) : 10;
    lr = argc > 3 ? atof(argv[3]) : 0.01;
    decay = argc > 4 ? atof(argv[4]) : 4e-5;
    if (check(effdet_replay_load(argv[1], 0, &plan), "effdet_replay_load")) return 3;
    printf("%s: %d launches per step\n", argv[1], effdet_replay_num_launches(plan));
    if (check(effdet_replay_region(plan, "images", &images, &n_images), "region images") ||
        check(effdet_replay_region(plan, "gt_boxes", &gt_boxes, &n_boxes), "region gt_boxes") ||
        check(effdet_replay_region(plan, "gt_labels", &gt_labels, &n_labels), "region gt_labels") ||
        check(effdet_replay_region(plan, "gt_counts", &gt_counts, &n_counts), "region gt_counts") ||
        check(effdet_replay_region(plan, "losses", &losses, &n_losses), "region losses")) {
        effdet_replay_destroy(plan);
        return 3;
    }
    batch = n_counts / sizeof(int32_t);
    kmax = n_labels / sizeof(int32_t) / batch;
    printf("batch %lu, %lu annotation slots per image, %lu image bytes per step\n", (unsigned long)batch,
           (unsigned long)kmax, (unsigned long)n_images);

    /* one synthetic batch: byte noise (the plan was exported with u8_input=True; a float plan takes the normalised
     * float32 image instead) and one box per image, (x1, y1, x2, y2) in pixels as float64, class 0 */
    h_images = (unsigned char *)malloc(n_images);
    h_boxes = (double *)calloc(n_boxes / sizeof(double), sizeof(double));
    h_labels = (int32_t *)calloc(n_labels / sizeof(int32_t), sizeof(int32_t));
    h_counts = (int32_t *)calloc(batch, sizeof(int32_t));
    for (i = 0; i < n_images; ++i) h_images[i] = (unsigned char)((i * 2654435761u) >> 24);
    for (i = 0; i < batch; ++i) {
        double *b = h_boxes + i * kmax * 4;
        b[0] = 16.0 + 8.0 * (double)i; b[1] = 24.0; b[2] = b[0] + 64.0; b[3] = 120.0;
        h_counts[i] = 1;
    }
    for (it = 0; it < steps && !rc; ++it) {
        rc = cuda_check(cudaMemcpyAsync(images, h_images, n_images, cudaMemcpyHostToDevice, 0), "copy images") ||
             cuda_check(cudaMemcpyAsync(gt_boxes, h_boxes, n_boxes, cudaMemcpyHostToDevice, 0), "copy gt_boxes") ||
             cuda_check(cudaMemcpyAsync(gt_labels, h_labels, n_labels, cudaMemcpyHostToDevice, 0), "copy gt_labels") ||
             cuda_check(cudaMemcpyAsync(gt_counts, h_counts, n_counts, cudaMemcpyHostToDevice, 0), "copy gt_counts");
        /* keras SGD: lr_t = lr / (1 + decay * iterations) */
        if (!rc) rc = check(effdet_replay_step(plan, lr / (1.0 + decay * (double)it), NULL), "effdet_replay_step");
        if (!rc)
            rc = cuda_check(cudaMemcpy(h_losses, losses, sizeof(h_losses) < n_losses ? sizeof(h_losses) : n_losses,
                                       cudaMemcpyDeviceToHost), "read losses");
        if (!rc) printf("step %ld: focal %.6f smooth-L1 %.6f\n", it, h_losses[0], h_losses[1]);
    }
    effdet_replay_destroy(plan);
    free(h_images); free(h_boxes); free(h_labels); free(h_counts);
    return rc ? 3 : 0;
}
