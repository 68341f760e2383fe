/* effdet_b200 -- C ABI of the B200-native EfficientDet hot path.
 *
 * Drop-in boundary for the path BASELINE.json names: the device work behind
 * Ely-S/EfficientDet's Python layer API.  The reference has no native/FFI layer
 * for this path (it is tf.keras Python; TensorFlow stock ops do the device work),
 * so each entry point cites the reference *Python* interface it replaces
 * (file:line under the reference tree) -- the host-side mirror in
 * the efficientdet_b200 Python package binds these through ctypes; INTEGRATION.md shows the
 * stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative EFFDET_E_* code;
 *     effdet_last_error() gives a thread-local message.
 *   - all pointers are DEVICE pointers unless the name ends in _host.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - activations are NHWC; convolution kernels are Keras HWIO; depthwise HWC(1).
 *   - dtype: EFFDET_F32 (accuracy mode) or EFFDET_BF16 (speed mode; fp32 accumulate).
 *   - functions only enqueue work on `stream`; they never synchronise, so they
 *     are CUDA-graph capturable (unless documented otherwise).
 */
#ifndef EFFDET_B200_H_
#define EFFDET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EFFDET_OK 0
#define EFFDET_E_INVALID (-1)     /* bad argument */
#define EFFDET_E_CUDA (-2)        /* CUDA runtime error (see effdet_last_error) */
#define EFFDET_E_CAPACITY (-3)    /* caller-provided workspace too small */
#define EFFDET_E_UNSUPPORTED (-4) /* shape / dtype combination not implemented */

#define EFFDET_F32 0
#define EFFDET_BF16 1

#define EFFDET_ACT_NONE 0
#define EFFDET_ACT_RELU 1
#define EFFDET_ACT_SWISH 2
#define EFFDET_ACT_SIGMOID 3

const char *effdet_last_error(void);
int effdet_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
long long effdet_launch_count(void);
/* cudaMemsetAsync(ptr, 0, bytes) on `stream` (zeroes accumulators inside captured graphs) */
int effdet_zero(void *ptr, size_t bytes, void *stream);

/* ---------------------------------------------------------------- geometry
 * utils/anchors.py:372-403 generate_anchors, :339-369 shift, :296-336 anchors_for_shape.
 * Host function: float64 table in level-major / cell-row-major / (ratio,scale) order.
 * level_hw_host: n_levels x {H, W}.  ratios/scales are the float32-rounded values
 * promoted to double (utils/anchors.py:46-52).  out_host: (sum H*W*nr*ns, 4) f64. */
int effdet_anchors_for_shape_host(const int *level_hw_host, const int *sizes_host,
                                  const int *strides_host, int n_levels,
                                  const double *ratios_host, int n_ratios,
                                  const double *scales_host, int n_scales,
                                  double *out_host, size_t out_capacity_rows);

/* utils/compute_overlap.pyx:13-53 -- IoU matrix (N,K) float64, "+1" convention. */
int effdet_compute_overlap(const double *boxes, size_t N, const double *query, size_t K,
                           double *overlaps, void *stream);

/* utils/anchors.py:130-207 anchor_targets_bbox (+ :210-239, :406-439).
 * anchors (N,4) f64; gt_boxes (B,Kmax,4) f64; gt_labels (B,Kmax) i32; gt_counts (B) i32;
 * image_hw (B,2) f64 {H,W}, H<0 => skip the centre-outside-image rule.
 * regression (B,N,5) f32 and labels (B,N,C+1) f32 are fully written (dense, reference
 * layout).  compact_* (optional, may be NULL): state (B,N) i8 {-1,0,1}, cls (B,N) i32
 * (class of the matched GT, -1 if not positive) -- the layout the fused loss consumes. */
int effdet_anchor_targets(const double *anchors, size_t N, const double *gt_boxes,
                          const int32_t *gt_labels, const int32_t *gt_counts, int B, int Kmax,
                          const double *image_hw, int num_classes, double negative_overlap,
                          double positive_overlap, float *regression, float *labels,
                          int8_t *compact_state, int32_t *compact_cls, void *stream);

/* ---------------------------------------------------------------- detection tail
 * RegressBoxes.py:126-164 apply_bbox_deltas: out = a + (d*std + mean) * [w,h,w,h],
 * each op individually rounded to float32 (no FMA contraction).
 * anchors: (1,N,4) broadcast when anchors_per_image == 0, else (B,N,4). */
int effdet_regress_boxes(const float *anchors, int anchors_per_image, const float *deltas,
                         const float mean_host[4], const float std_host[4], int B, size_t N,
                         float *out, void *stream);
/* ClipBoxes.py:9-24: x in [0, W-1], y in [0, H-1]. in == out allowed. */
int effdet_clip_boxes(const float *boxes, int B, size_t N, float height, float width, float *out,
                      void *stream);
/* the two fused (model.py:414-429): one read of anchors+deltas, one write. */
int effdet_regress_clip_boxes(const float *anchors, int anchors_per_image, const float *deltas,
                              const float mean_host[4], const float std_host[4], int B, size_t N,
                              float height, float width, float *out, void *stream);

/* FilterDetections.py:37-118 filter_detections mapped over the batch (:183-188).
 * boxes (B,N,4) f32, classification (B,N,C) f32.
 * Outputs: boxes (B,max_det,4) f32, scores (B,max_det) f32, labels (B,max_det) i32,
 * padded with -1; out_indices (optional, may be NULL): (B,max_det) i32 anchor index of
 * each detection.  nms == 0 reproduces `nms=False` (iou_threshold forced to 0 => no NMS,
 * FilterDetections.py:170-171, :11).  class_specific == 0 reproduces :86-92.
 * workspace: device scratch of effdet_filter_detections_workspace_size() bytes;
 * cand_capacity = max number of (anchor,class) pairs above the threshold over the whole
 * batch that the workspace can hold.  status (device int32[4]): [0] = 1 if the candidate
 * list overflowed cand_capacity (outputs invalid, re-run with [1] = required capacity),
 * [1] = total candidates, [2..3] reserved.  Bit-exact w.r.t. oracle/tail.py. */
size_t effdet_filter_detections_workspace_size(int B, size_t N, int C, size_t cand_capacity,
                                               int max_detections);
int effdet_filter_detections(const float *boxes, const float *classification, int B, size_t N,
                             int C, float score_threshold, float iou_threshold, int max_detections,
                             int class_specific, int nms, void *workspace, size_t workspace_bytes,
                             size_t cand_capacity, float *out_boxes, float *out_scores,
                             int32_t *out_labels, int32_t *out_indices, int32_t *status,
                             void *stream);

/* ---------------------------------------------------------------- network forward
 * BatchNormalization (inference form) folded to per-channel scale/shift:
 * scale = gamma/sqrt(var+eps), shift = beta - mean*scale  (efficientnet.py:233-236 eps 1e-3;
 * model.py:42-45 eps 1e-4). */
int effdet_bn_fold(const float *gamma, const float *beta, const float *moving_mean,
                   const float *moving_variance, float eps, float *scale, float *shift, int C,
                   void *stream);

/* efficientnet.py:413-423 stem: Conv3x3 stride 2 SAME (3 -> C0, kernel HWIO f32) + BN + swish.
 * images (B,H,W,3) f32 -> out (B,ceil(H/2),ceil(W/2),C0) of out_dtype. */
int effdet_stem_conv(const float *images, const float *kernel, const float *scale,
                     const float *shift, void *out, int B, int H, int W, int C0, int out_dtype,
                     void *stream);
/* Same stem, bf16 output, selectable epilogue: act = EFFDET_ACT_SWISH (as above) or EFFDET_ACT_NONE
 * (training mode: the raw convolution z with scale = 1 / shift = 0; BatchNormalization then runs on batch
 * statistics, efficientnet.py:413-423 with trainable BN). */
int effdet_stem_conv_act(const float *images, const float *kernel, const float *scale, const float *shift,
                         void *out, int B, int H, int W, int C0, int act, void *stream);
/* Same stem fed with the raw letterboxed uint8 RGB image -- what train_tpu.py:170-183 decodes from the
 * TFRecord PNG and what generators/common.py:406-417 builds before dividing by 255.  normalize_image
 * (train_tpu.py:135-140 == generators/common.py:418-429: ((v / 255) - mean_c) / std_c in float32) is applied on
 * the fly through lut (3 x 256 f32, lut[c*256 + v], evaluated by the caller in float32 exactly like the
 * reference), so the result is bit-identical to effdet_stem_conv on the normalised float image while the
 * 12 B/pixel float image is never uploaded or stored.  SAME padding is zero in normalised space.
 * act: EFFDET_ACT_SWISH, or EFFDET_ACT_NONE with bf16 output (training: raw z). */
int effdet_stem_conv_u8(const unsigned char *images, const float *lut, const float *kernel, const float *scale,
                        const float *shift, void *out, int B, int H, int W, int C0, int act, int out_dtype,
                        void *stream);
/* normalize_image alone: out[i] = lut[(i % 3)*256 + images[i]] over n_values = B*H*W*3 interleaved RGB bytes
 * (for callers that need the float image itself: the weight gradient of a trainable stem). */
int effdet_normalize_u8(const unsigned char *images, const float *lut, float *out, size_t n_values,
                        void *stream);

/* Letterbox resize of one raw uint8 RGB image (src_h, src_w, 3; rows src_row_stride bytes apart, 0 = dense) into
 * out (image_size, image_size, 3) uint8: utils/__init__.py:103-132 resize_image == generators/common.py:406-417
 * (cv2.resize bilinear to (resized_w, resized_h), pasted into the centre of a grey 128 square).  Bit-exact with
 * OpenCV's 8-bit INTER_LINEAR (fixed-point, 11-bit weights); the output feeds effdet_stem_conv_u8 directly.
 * effdet_letterbox_geometry is the host arithmetic of the same function: resized extent, paste offsets and the
 * scale the reference returns (0 when the image already is image_size x image_size, as the reference does). */
int effdet_letterbox_geometry(int src_h, int src_w, int image_size, int *resized_h, int *resized_w,
                              int *offset_h, int *offset_w, double *scale);
int effdet_letterbox_u8(const unsigned char *image, int src_h, int src_w, long long src_row_stride,
                        unsigned char *out, int image_size, void *stream);


/* Dense convolution as implicit GEMM (1x1 or 3x3, stride 1 or 2, TF SAME padding) with fused
 * epilogue  y = act(conv(x * gate) * scale + shift) [* keep[b] + residual].
 * Replaces: efficientnet.py:228-237 expand_conv+bn+swish, :289-304 project_conv+bn(+drop)+add
 * with the SE multiply (:286) folded into the input load; model.py:71-90 ConvBlock;
 * model.py:293-309 / :324-351 head convs.  Up to 5 "groups" (pyramid levels) share the
 * weights and run in ONE launch; each group may write with its own row stride (ldc) and
 * per-image stride so the head outputs land directly in the concatenated (B,N,4)/(B,N,C)
 * tensors (model.py:393-398) without a concat copy.  0 for ldc / y_batch_stride = dense. */
#define EFFDET_MAX_GROUPS 5
typedef struct effdet_conv_desc {
    int n_groups;
    const void *x[EFFDET_MAX_GROUPS];        /* (B,H,W,Cin) in_dtype */
    void *y[EFFDET_MAX_GROUPS];              /* out_dtype */
    const void *residual[EFFDET_MAX_GROUPS]; /* out_dtype, indexed like y; or NULL */
    int H[EFFDET_MAX_GROUPS], W[EFFDET_MAX_GROUPS];
    int ldc[EFFDET_MAX_GROUPS];
    long long y_batch_stride[EFFDET_MAX_GROUPS];
    int ldx[EFFDET_MAX_GROUPS];              /* input pixel stride (elements), 0 = Cin */
    long long x_batch_stride[EFFDET_MAX_GROUPS]; /* input image stride (elements), 0 = dense */
    const void *relu_mask[EFFDET_MAX_GROUPS];    /* out_dtype, indexed like y: output zeroed where
                                                    mask <= 0 (ReLU backward fused into a data-
                                                    gradient convolution); or NULL */
    int B, Cin, Cout, kh, kw, stride;
    const float *weight;  /* (kh,kw,Cin,Cout) f32, Keras HWIO */
    const float *scale;   /* (Cout) or NULL */
    const float *shift;   /* (Cout) folded-BN shift or conv bias, or NULL */
    const float *gate;    /* (B,Cin) f32 squeeze-excite gate, or NULL */
    const float *keep;    /* (B) f32 drop-connect scale (FixedDropout, efficientnet.py:300-303) or NULL */
    int act;              /* EFFDET_ACT_* */
    int in_dtype, out_dtype;
    const void *weight_bf16; /* optional bf16 panel from effdet_conv_weight_panel for the tcgen05 path */
    int allow_tensor_core;   /* 0 = force the SIMT fp32-accumulate kernel */
    int weight_per_sample;   /* weight_bf16 holds one panel per image (SE gate folded in) */
    int split_planes;        /* fp32 accuracy mode on the tensor cores: x is the (B,H,W,2*Cin) bf16 hi | lo split of an
                                fp32 tensor (effdet_split_bf16), weight_bf16 a panel from
                                effdet_conv_weight_panel_split, out_dtype EFFDET_F32.  The convolution is then the
                                three-term product hi*Whi + lo*Whi + hi*Wlo accumulated in fp32 (one GEMM with 3x the
                                K extent): ~2^-18 relative per product. */
} effdet_conv_desc;
int effdet_conv2d(const effdet_conv_desc *desc, void *stream);

/* bf16 weight panels for the tcgen05 path of effdet_conv2d (stride 1, bf16 activations):
 * mode 0: forward, panel[tap][Cout_pad][Cin_pad] = kernel[tap][ci][co];
 * mode 1: data gradient, panel[tap][Cin_pad][Cout_pad] = kernel[flipped tap][ci][co] (use with
 *         Cin/Cout swapped in the descriptor);
 * gate != NULL (mode 0, 1x1): one panel per image with the squeeze-excite gate (B,Cin) folded
 *         into the weights (efficientnet.py:286) -- set weight_per_sample in the descriptor.
 * effdet_conv_weight_panel_elems(taps or B, K, N) gives the bf16 element count to allocate. */
int effdet_conv_tc_block_n(int n);
size_t effdet_conv_weight_panel_elems(int taps_or_samples, int K, int N);
int effdet_conv_weight_panel(const float *kernel, void *panel, int taps, int Cin, int Cout, int mode,
                             const float *gate, int B, void *stream);
/* fp32 accuracy mode on the tensor cores (split_planes in the descriptor).  effdet_split_bf16: out (rows, 2*C) bf16
 * = [hi | lo] with hi = bf16(x), lo = bf16(x - hi) for x (rows, C) f32.  effdet_conv_weight_panel_split: forward
 * panel[tap or image][Cout_pad][3 * Cin_pad] = [Whi | Whi | Wlo] of the (gated) kernel; _elems gives its size. */
int effdet_split_bf16(const float *x, void *out, size_t rows, int C, void *stream);
/* effdet_dwconv for fp32 tensors with the OUTPUT written directly as that hi | lo split (B,Ho,Wo,2*C) bf16: the
 * depthwise convolution of an MBConv block feeds only the project convolution (efficientnet.py:242-296). */
int effdet_dwconv_split_out(const float *x, const float *kernel, const float *scale, const float *shift,
                            void *y_planes, float *se_sum, int se_blocks, int B, int H, int W, int C, int k,
                            int stride, int act, void *stream);
size_t effdet_conv_weight_panel_split_elems(int taps_or_samples, int Cin, int Cout);
int effdet_conv_weight_panel_split(const float *kernel, void *panel, int taps, int Cin, int Cout,
                                   const float *gate, int B, void *stream);

/* Depthwise kxk (k = 3 or 5, stride 1 or 2, SAME) + BN + activation; optionally writes
 * per-(image, block, channel) partial spatial SUMS of the activated output into se_sum
 * (B, se_blocks, C) f32 for squeeze-excite (deterministic: no atomics; every cell is written).
 * se_blocks must be effdet_dwconv_se_blocks(...) for the same arguments (ignored when se_sum
 * is NULL).  efficientnet.py:242-252 (+ :259-260 squeeze).
 * kernel (k,k,C) f32 = Keras depthwise_kernel (k,k,C,1).  C % 8 == 0. */
int effdet_dwconv_se_blocks(int B, int H, int W, int C, int stride, int dtype);
int effdet_dwconv(const void *x, const float *kernel, const float *scale, const float *shift,
                  void *y, float *se_sum, int se_blocks, int B, int H, int W, int C, int k,
                  int stride, int act, int dtype, void *stream);

/* Squeeze-excite FCs (efficientnet.py:255-286): mean = sum_blocks(se_sum)/(HW);
 * r = swish(W1^T mean + b1); gate = sigmoid(W2^T r + b2).  se_sum (B,se_blocks,C);
 * w1 (C,R), w2 (R,C) f32 (Keras 1x1 conv kernels). gate (B,C) f32.  Runs on thread-block clusters of 2-8 CTAs
 * per image (channel slices, partial FC1 sums exchanged through distributed shared memory in rank order:
 * deterministic); EFFDET_SE_CLUSTER=0 selects the one-CTA-per-image kernel. */
int effdet_se_gate(const float *se_sum, int se_blocks, float inv_hw, const float *w1,
                   const float *b1, const float *w2, const float *b2, float *gate, int B, int C,
                   int R, void *stream);

/* layers.py:11-39 wBiFPNAdd: out = sum_i relu(w_i) x_i / (sum_i relu(w_i) + eps); w == NULL
 * gives keras.layers.Add (plain sum).  n_inputs 2 or 3, `count` elements of dtype. */
int effdet_wbifpn_add(const void *const *inputs_host, int n_inputs, const float *w, float eps,
                      void *out, size_t count, int dtype, void *stream);

/* One fused BiFPN node (model.py:154-194 / :226-266): resample-on-load + (weighted) fusion +
 * DepthwiseConv3x3 SAME + BN + ReLU.  in0 is the resampled input: mode 1 = nearest x2 upsample
 * of a (B,H/2,W/2,C) tensor, mode 2 = 2x2/stride-2 max-pool of a (B,2H,2W,C) tensor, mode 0 =
 * same resolution.  in1 / in2 (in2 may be NULL) are (B,H,W,C).  Weight order = input order
 * (SURVEY Appendix D).  w NULL => unweighted Add. */
int effdet_bifpn_node(const void *in0, int mode0, const void *in1, const void *in2,
                      const float *w, float eps, const float *dw_kernel, const float *scale,
                      const float *shift, void *out, int B, int H, int W, int C, int dtype,
                      void *stream);

/* ---------------------------------------------------------------- training
 * utils/tpu.py:84-155 tpu_focal + :26-81 tpu_smooth_l1, forward and backward in one pass.
 * classification (B,N,C) f32 probabilities (sigmoid output of the class head, model.py:351),
 * regression (B,N,4) f32.  Targets in the reference layout: regression_t (B,N,5) f32 and either
 * dense labels_t (B,N,C+1) f32 or, when labels_t == NULL, the compact pair state (B,N) i8 /
 * cls (B,N) i32 written by effdet_anchor_targets.  Gradients: dcls_logits (B,N,C) f32 w.r.t.
 * the class head's pre-sigmoid outputs, dreg (B,N,4) f32; both already divided by
 * max(1,#positive) and multiplied by grad_scale.  out8 (device f32[8]): [0] focal loss,
 * [1] smooth-L1 loss, [2]/[3] #positive anchors, [4]/[5] 1/normaliser.
 * Optional (n_levels > 0): a second, bf16 copy of both gradients in per-pyramid-level dense
 * buffers (B, cells_l, cpad) with channel = anchor*per + k (what the TMA-fed tensor-core
 * gradient kernels read); level_cells_host[l] = H_l*W_l; padding channels are not written.
 * With n_levels > 0, dcls_logits and / or dreg may be NULL: the concatenated fp32 copies are then not
 * written (a caller whose backward reads only the level buffers saves their 4*N*(C+4) bytes per image). */
size_t effdet_detection_losses_workspace_size(void);
int effdet_detection_losses(const float *classification, const float *regression,
                            const float *regression_t, const float *labels_t, const int8_t *state,
                            const int32_t *cls, int B, size_t N, int C, float alpha, float gamma,
                            float delta, float grad_scale, float *dcls_logits, float *dreg,
                            float *out8, void *workspace, size_t workspace_bytes,
                            void *const *dcls_levels_host, void *const *dreg_levels_host,
                            const int *level_cells_host, int n_levels, int cpad_cls, int cpad_reg,
                            void *stream);

/* BiFPN fusion forward keeping the fused tensor (training): same semantics as the fusion stage
 * of effdet_bifpn_node (model.py:154-194/226-266, layers.py:26-31). */
int effdet_resample_fuse(const void *in0, int mode0, const void *in1, const void *in2,
                         const float *w, float eps, void *out, int B, int H, int W, int C, int dtype,
                         void *stream);

/* Row-block count used by the deterministic column reductions below for a (rows, C) matrix;
 * `partial` scratch must hold 2*C*blocks floats.  The count is one balanced wave on the 148 SMs (at most 4 blocks
 * per SM, at least 64 rows per block), so the partial matrix stays below 1/16 of the tensor. */
int effdet_colreduce_blocks(size_t rows, int C, int dtype);

/* BatchNormalization, training mode (model.py:59-62/81-84 with trainable=True; TF fused BN):
 * batch mean / biased variance of z (rows,C) -> scale, shift (y = z*scale + shift), saved
 * mean / invstd for the backward pass, and the moving-average update
 * moving = moving*momentum + batch*(1-momentum) (unbiased variance); moving_* may be NULL. */
int effdet_bn_train_stats(const void *z, size_t rows, int C, const float *gamma, const float *beta,
                          float eps, float momentum, float *moving_mean, float *moving_var,
                          float *scale, float *shift, float *save_mean, float *save_invstd,
                          float *partial, int nblk, int dtype, void *stream);
/* Depthwise conv with raw output z (no activation) + the batch statistics of the BatchNormalization that
 * follows it (model.py:48-68 DepthwiseConvBlock, trainable BN), one pass over the data: the depthwise kernel
 * emits per-tile sums / sums of squares and only the finalize step of effdet_bn_train_stats runs afterwards.
 * bf16 only; ones / zeros: C floats each; partial: 2*C*nblk floats, nblk = B * effdet_dwconv_se_blocks(). */
int effdet_dwconv_bn_stats(const void *x, const float *kernel, const float *ones, const float *zeros, void *z,
                           int B, int H, int W, int C, int k, int stride, const float *gamma, const float *beta,
                           float eps, float momentum, float *moving_mean, float *moving_var, float *scale,
                           float *shift, float *save_mean, float *save_invstd, float *partial, int nblk,
                           int dtype, void *stream);

/* y = act(z*scale + shift) over a (rows,C) matrix. */
int effdet_scale_shift_act(const void *z, const float *scale, const float *shift, void *y,
                           size_t rows, int C, int act, int dtype, void *stream);
/* Backward of y = relu(BN(z)): dz, dgamma, dbeta.  frozen_scale != NULL => BN ran in inference
 * mode (freeze_bn / trainable=False): dz = scale * dy*[y>0], no dgamma/dbeta.
 * k123: scratch 3*C floats; partial: 2*C*nblk floats. */
int effdet_bn_relu_backward(const void *dy, const void *y, const void *z, size_t rows, int C,
                            const float *gamma, const float *save_mean, const float *save_invstd,
                            const float *frozen_scale, float *dgamma, float *dbeta, void *dz,
                            float *k123, float *partial, int nblk, int dtype, void *stream);
/* Column sums of x viewed as (rows, C) -> out[C/fold] (bias gradients; `fold` consecutive logical
 * rows were packed into one physical row to keep rows vectorisable). */
int effdet_colsum(const void *x, size_t rows, int C, int fold, float *out, int accumulate,
                  float *partial, int nblk, int dtype, void *stream);
/* out[c] += sum over rows [row0, row0 + nrows) of the dense (*, C) matrix x: the scalar tail for the few rows
 * that do not fill a whole fold of effdet_colsum. */
int effdet_colsum_tail(const void *x, size_t row0, int nrows, int C, float *out, int dtype, void *stream);


/* Depthwise 3x3 stride-1 SAME weight gradient (model.py:48-55 DepthwiseConv2D backward-filter).
 * partial: 9*C*blocks floats with blocks = effdet_dw_wgrad_blocks(). */
int effdet_dw_wgrad_blocks(int B, int H, int W, int C, int dtype);
int effdet_dw_wgrad(const void *f, const void *dz, int B, int H, int W, int C, float *dkernel,
                    float *partial, int nblk, int dtype, void *stream);

/* Fusion backward (layers.py:26-31 + UpSampling2D / MaxPooling2D gradients).  df (B,H,W,C) is the
 * gradient of the fused tensor.  which = 0: gradient of in0 routed through its resampling
 * (mode 1: 2x2 sum into the (H/2,W/2) tensor; mode 2: to the first maximum of each 2x2 window of
 * the (2H,2W) tensor `in0`); which = 1/2: same-resolution inputs.  dst is overwritten or
 * accumulated into. */
int effdet_fuse_backward_input(const void *df, int which, int mode0, const void *in0, const float *w,
                               int n_inputs, float eps, void *dst, int accumulate, int B, int H,
                               int W, int C, int dtype, void *stream);
/* dw_i = [w_i > 0] * sum(df * (in_i - f)) / (sum relu(w) + eps); partial: 4*1184 floats. */
int effdet_fuse_backward_weights(const void *df, const void *f, const void *in0, int mode0,
                                 const void *in1, const void *in2, const float *w, float eps,
                                 float *dw, float *partial, int B, int H, int W, int C, int dtype,
                                 void *stream);

/* Dense-convolution weight gradient (cuDNN bwd-filter in the reference's TF graph):
 * dweight (kh,kw,Cin,Cout) f32 (+)= sum over groups/pixels x (*) dz.  dz may be strided (head
 * outputs inside the concatenated tensors).  partial: n_splits*kh*kw*Cin*Cout floats with
 * n_splits = effdet_conv_wgrad_splits(desc). */
typedef struct effdet_wgrad_desc {
    int n_groups;
    const void *x[EFFDET_MAX_GROUPS];   /* (B,H,W,Cin) x_dtype, dense */
    const void *dz[EFFDET_MAX_GROUPS];  /* (B,Ho,Wo,Cout) dz_dtype */
    int H[EFFDET_MAX_GROUPS], W[EFFDET_MAX_GROUPS];
    int dz_ld[EFFDET_MAX_GROUPS];
    long long dz_batch_stride[EFFDET_MAX_GROUPS];
    int B, Cin, Cout, kh, kw, stride;
    float *dweight;
    float *partial;
    int n_splits;
    int accumulate;
    int x_dtype, dz_dtype;
    /* effdet_conv_wgrad_tc only, optional: bias gradient dbias[Cout] = sum over pixels of dz, produced by
     * the same launch when effdet_conv_wgrad_tc_fuses_bias(desc) != 0 (3x3, Cin <= 64: the spare half of
     * the last tap pair's M tile multiplies a block of ones).  partial then needs n_splits * Cout more
     * floats.  Ignored (must be NULL) otherwise. */
    float *dbias;
} effdet_wgrad_desc;
int effdet_conv_wgrad_splits(const effdet_wgrad_desc *desc);
int effdet_conv_wgrad(const effdet_wgrad_desc *desc, void *stream);
/* Tensor-core (tcgen05, MN-major operands fed by TMA) version for bf16 operands, stride 1,
 * k in {1,3}: x dense (B,H,W,Cin), dz (B,H,W,dz_ld) bf16 with the first Cout channels used.
 * partial: effdet_conv_wgrad_tc_splits(desc) * kh*kw*Cin*Cout floats. */
int effdet_conv_wgrad_tc_splits(const effdet_wgrad_desc *desc);
int effdet_conv_wgrad_tc_fuses_bias(const effdet_wgrad_desc *desc);
int effdet_conv_wgrad_tc(const effdet_wgrad_desc *desc, void *stream);
/* Data gradient of a strided (stride 2) 1x1/3x3 SAME convolution; weight is the forward HWIO
 * kernel.  Stride-1 data gradients use effdet_conv2d on effdet_conv_weight_transpose'd weights. */
int effdet_conv_dgrad_strided(const void *dz, const float *weight, void *dx, int accumulate, int B,
                              int H, int W, int Cin, int Cout, int k, int stride, int dtype,
                              void *stream);
/* (taps,Cin,Cout) -> spatially flipped (taps,Cout,Cin): the kernel of the data-gradient conv. */
int effdet_conv_weight_transpose(const float *in, float *out, int taps, int Cin, int Cout,
                                 void *stream);
/* (taps, n) -> spatially flipped (depthwise data-gradient kernel). */
int effdet_flip_taps(const float *in, float *out, int taps, int n, void *stream);

/* Zero insertion for the data gradient of a stride-2 convolution (implicit in model.fit, train_tpu.py:330;
 * forward: model.py:205-211 BiFPN P6/P7 laterals): out (B,H,W,C) holds dz (B,Ho,Wo,C) at rows 2*oy + ay,
 * columns 2*ox + ax and zeros elsewhere, so that the stride-1 tensor-core data gradient applies. */
int effdet_zero_insert(const void *dz, void *out, int B, int Ho, int Wo, int C, int H, int W, int ay, int ax,
                       int dtype, void *stream);

/* keras SGD(lr, decay, momentum) of train_tpu.py:268-269 on a flat fp32 range:
 * v = momentum*v - lr_t*(g*grad_scale); w += v, lr_t = lr/(1+decay*iterations) from the caller. */
int effdet_sgd_momentum_step(float *w, const float *g, float *v, size_t n, float lr_t,
                             float momentum, float grad_scale, void *stream);
/* Same update, lr_t read from device memory (one float): the launch can sit in a CUDA graph while the schedule
 * changes the rate every step (compiled plans, effdet_replay_step). */
int effdet_sgd_momentum_step_dev_lr(float *w, const float *g, float *v, size_t n, const float *lr_dev,
                                    float momentum, float grad_scale, void *stream);

/* ---------------------------------------------------------------- training the backbone
 * (train_tpu.py without --freeze-backbone; BASELINE config 4).
 * Backward of y = act(BN(z)), act in {none, relu, swish} (efficientnet.py:228-252,289-296).
 * ua/ub: the affine map u = z*ua + ub applied in the forward pass (training-mode BN: the scale /
 * shift written by effdet_bn_train_stats; frozen BN: the folded inference scale/shift, frozen=1,
 * no dgamma/dbeta).  k123: 3*C floats scratch; partial: 2*C*nblk floats. */
int effdet_bn_act_backward(const void *dy, const void *z, size_t rows, int C, const float *gamma,
                           const float *save_mean, const float *save_invstd, const float *ua,
                           const float *ub, int frozen, int act, float *dgamma, float *dbeta, void *dz,
                           float *k123, float *partial, int nblk, int dtype, void *stream);
/* Per-(image, block, channel) spatial sums of y (B,HW,C) -> partial (B, nblk, C) with
 * nblk = effdet_se_backward_blocks(): the SE squeeze (efficientnet.py:259-260) in training mode. */
int effdet_spatial_sum(const void *y, float *partial, int nblk, int B, int HW, int C, int dtype,
                       void *stream);
/* out = y * gate[b][c]  (efficientnet.py:286 se_excite, materialised in training mode). */
int effdet_se_apply(const void *y, const float *gate, void *out, int B, int HW, int C, int dtype,
                    void *stream);
/* Stochastic depth = FixedDropout(drop_rate, noise_shape=(None,1,1,1)) on the projected branch of the skip
 * blocks in the training phase (efficientnet.py:147-188, :300-304; drop_rate per block :466-467).
 * effdet_drop_connect_scales draws scales[blk][b] = (uniform >= rates[blk]) / (1 - rates[blk]) for all
 * blocks of one step (Keras dropout semantics); the generator is counter based (seed, *step_counter, blk, b)
 * and increments *step_counter on the device, so every replay of a captured graph draws a new mask.
 * effdet_sample_scale_add: out = y * scale[b] (+ res), per image of `per_image` elements: the forward of the
 * dropped branch (res = block input) and, with res == NULL, its backward. */
int effdet_drop_connect_scales(const float *rates, int nblocks, int B, unsigned long long seed,
                               unsigned long long *step_counter, float *scales, void *stream);
int effdet_sample_scale_add(const void *y, const float *scale, const void *res, void *out, int B,
                            size_t per_image, int dtype, void *stream);
/* Squeeze-excite backward (efficientnet.py:255-286): dy (gradient of the SE input, both the direct
 * path and the path through GlobalAveragePooling), gradients of se_reduce / se_expand. */
int effdet_se_backward_blocks(int HW, int C, int dtype);
int effdet_se_backward(const void *dyg, const void *y, const float *gate, const float *se_sum,
                       int se_blocks, const float *w1, const float *b1, const float *w2, const float *b2,
                       void *dy, float *dw1, float *db1, float *dw2, float *db2, float *dg_partial,
                       int dg_blocks, float *fc_scratch, float *dmean, int B, int HW, int C, int R,
                       int dtype, void *stream);

/* Fused backward of the squeeze-excite gate and of the depthwise BatchNorm (training statistics) + swish of an
 * MBConv block (efficientnet.py:242-286): dyg (B,HW,C) gradient of y*gate, z (B,HW,C) the raw depthwise output
 * -> dz (B,HW,C) gradient of z, the gradients of the SE kernels / biases and of gamma / beta.  Same results as
 * effdet_se_backward followed by effdet_bn_act_backward(act = swish) up to rounding, in five instead of nine
 * passes over the tensors (y and swish' are recomputed from z; dy is never materialised).  ua / ub: the affine
 * map u = z*ua + ub of the forward pass.  Scratch (floats): dg_partial B*(nblk+1)*C, bn_partial B*(nblk+1)*4*C,
 * bn_rows B*2*C, fc_scratch B*(2*C*R + R + C), dmean B*C, k123 3*C; nblk = effdet_se_bn_backward_blocks(). */
int effdet_se_bn_backward_blocks(int B, int HW, int C, int dtype);
int effdet_se_bn_backward(const void *dyg, const void *z, const float *gate, const float *se_sum, int se_blocks,
                          const float *w1, const float *b1, const float *w2, const float *b2, float *dw1,
                          float *db1, float *dw2, float *db2, const float *gamma, const float *save_mean,
                          const float *save_invstd, const float *ua, const float *ub, float *dgamma,
                          float *dbeta, void *dz, float *k123, float *dg_partial, float *bn_partial,
                          float *bn_rows, int nblk, float *fc_scratch, float *dmean, int B, int HW, int C, int R,
                          int dtype, void *stream);
/* Depthwise conv backward, k in {3,5}, stride in {1,2}: dkernel (k,k,C) and (optional) dx. */
int effdet_dw_backward_blocks(int B, int H, int W, int C, int k, int stride, int dtype);
int effdet_dw_backward(const void *x, const void *dz, const float *kernel, void *dx, float *dkernel,
                       float *partial, int nblk, int B, int H, int W, int C, int k, int stride, int dtype,
                       void *stream);
/* Stem conv weight gradient (efficientnet.py:413-418). */
int effdet_stem_wgrad_blocks(int B, int H, int W);
int effdet_stem_wgrad(const float *images, const void *dz, float *dkernel, float *partial, int nblk, int B,
                      int H, int W, int C0, int dtype, void *stream);

/* ---------------------------------------------------------------- plan level (whole model)
 * What the reference's callers use: `model, prediction_model = efficientdet(phi, ...)` (model.py:356-452) and
 * `prediction_model.predict_on_batch([images (, anchors)])` (inference.py:57-59, predict.py:101-105).
 * A plan is the network lowered for a fixed (phi, image size, batch, classes, BiFPN kind, dtype) into a static
 * list of this library's launches over plan-owned buffers, captured in a CUDA graph.  Not thread-safe per plan;
 * one plan per (GPU, stream).  The caller owns every buffer it passes; the library owns the plan's workspace.
 *
 *   effdet_plan_create       phi 0..6 (model.py:367); image_size = image_sizes[phi] of model.py:29 or any multiple of
 *                            128; dtype EFFDET_F32 (accuracy mode) | EFFDET_BF16 (speed mode, tcgen05 convolutions);
 *                            flags: EFFDET_PLAN_U8_INPUT = images are the raw letterboxed uint8 RGB picture and
 *                            normalize_image runs inside the stem; EFFDET_PLAN_NO_GRAPH = launch eagerly.
 *   effdet_plan_weight_info  the weight manifest: Keras names "<layer>/<weight>" (nested heads "box_head/...",
 *                            "class_head/...") and Keras shapes (conv HWIO, depthwise HWC1, BN vectors), in the
 *                            creation order of the reference graph -- what `load_weights(by_name=True)` matches
 *                            (train.py:329-332).
 *   effdet_plan_bind_weights float32 DEVICE pointers by name (a trailing ":0" is accepted; unknown names are
 *                            ignored like by_name=True).  The values are COPIED into the plan (asynchronously on
 *                            `stream`); re-bind after the weights changed.  _host: the same from host memory.
 *   effdet_forward           images (B,S,S,3) f32 (uint8 with EFFDET_PLAN_U8_INPUT), device -> regression (B,N,4) f32,
 *                            classification (B,N,C) f32 (model.py:393-407).  Output pointers may be NULL: the
 *                            results then stay in the plan's own buffers (effdet_plan_buffers).  Every weight must
 *                            have been bound.  Enqueues only (graph launch).
 *   effdet_detect            forward + RegressBoxes (mean 0, std 0.2) + ClipBoxes + FilterDetections (model.py:
 *                            414-450): boxes (B,max_det,4) f32, scores (B,max_det) f32, labels (B,max_det) i32, padded
 *                            with -1.  anchors (1,N,4) f32 device, or NULL = utils/anchors.py anchors_for_shape of the
 *                            plan's image size (the `anchors=` baked form, model.py:414-419).  Synchronises the stream
 *                            once (candidate-capacity word).  _host: host images / anchors / outputs, copies inside.
 * The training step (losses, backward, optimizer) is lowered on the host side by efficientdet_b200/train.py over the
 * per-layer entry points above. */
typedef struct effdet_plan effdet_plan_t;
#define EFFDET_PLAN_U8_INPUT 1u
#define EFFDET_PLAN_NO_GRAPH 2u
int effdet_plan_create(int phi, int image_size, int batch, int num_classes, int weighted_bifpn, int dtype,
                       unsigned flags, effdet_plan_t **out);
int effdet_plan_destroy(effdet_plan_t *plan);
int effdet_plan_num_weights(const effdet_plan_t *plan);
size_t effdet_plan_num_anchors(const effdet_plan_t *plan);
int effdet_plan_num_launches(const effdet_plan_t *plan);
/* Test facility for machines WITHOUT a CUDA driver: with the environment variable EFFDET_DRY_RUN set,
 * effdet_plan_create builds the lowering over host memory and effdet_plan_dry_run issues every launch of the plan
 * (BatchNorm folds, weight panels, the forward launch list) through its entry point; each must pass that entry
 * point's validation and launch configuration and fail only at its first CUDA call.  Returns the first other error.
 * With a driver present plans are never built this way and the call returns EFFDET_E_INVALID. */
int effdet_plan_dry_run(effdet_plan_t *plan, int *launches_checked);
int effdet_plan_weight_info(const effdet_plan_t *plan, int index, const char **name, int *ndim, int dims[4]);
int effdet_plan_bind_weights(effdet_plan_t *plan, const char *const *names, const void *const *device_ptrs, int n,
                             void *stream);
int effdet_plan_bind_weights_host(effdet_plan_t *plan, const char *const *names, const void *const *host_ptrs, int n);
int effdet_forward(effdet_plan_t *plan, const void *images, float *regression_out, float *classification_out,
                   void *stream);
int effdet_plan_buffers(effdet_plan_t *plan, void **images, float **regression, float **classification);
int effdet_detect(effdet_plan_t *plan, const void *images, const float *anchors, float score_threshold,
                  float iou_threshold, int max_detections, float *boxes_out, float *scores_out,
                  int32_t *labels_out, void *stream);
int effdet_detect_host(effdet_plan_t *plan, const void *images_host, const float *anchors_host,
                       float score_threshold, float iou_threshold, int max_detections, float *boxes_host,
                       float *scores_host, int32_t *labels_host);

/* ---------------------------------------------------------------- plan level, training: compiled plans
 * The training step the reference runs per replica (train_tpu.py:249-346: anchor targets -> forward in training
 * mode -> focal + smooth-L1 -> backward -> SGD) as ONE call from any host language.  The launch list is lowered by
 * efficientdet_b200/train.py and written by efficientdet_b200.plan_export.export_train_plan(model, batch, path)
 * ("compile once"); this library loads it, owns the device regions and a CUDA graph of the launches, and replays
 * them ("run anywhere") -- same launches, same arguments, bit-identical results to Trainer.train_on_batch.
 * Named regions (effdet_replay_region): "images" (B,S,S,3) f32 (u8 for a uint8 plan), "gt_boxes" (B,kmax,4) f64
 * x1,y1,x2,y2, "gt_labels" (B,kmax) i32, "gt_counts" (B,) i32, "image_hw" (B,2) f64, "anchors" (N,4) f64,
 * "losses" (8,) f32 ([0] focal, [1] smooth-L1), "weights" (the flat fp32 parameter buffer; the .json sidecar of
 * the plan file lists offset and shape of every Keras weight), "lr" (one f32, written by effdet_replay_step).  The caller copies a batch into the input regions,
 * calls effdet_replay_step with the step's learning rate (keras: lr / (1 + decay * iterations)) and reads
 * "losses".  flags: 1 = no CUDA graph (launch eagerly). */
typedef struct effdet_replay effdet_replay_t;
int effdet_replay_load(const char *path, int flags, effdet_replay_t **out);
int effdet_replay_num_launches(const effdet_replay_t *plan);
int effdet_replay_region(effdet_replay_t *plan, const char *name, void **device_ptr, size_t *bytes);
int effdet_replay_step(effdet_replay_t *plan, double learning_rate, void *stream);
int effdet_replay_destroy(effdet_replay_t *plan);

#ifdef __cplusplus
}
#endif
#endif /* EFFDET_B200_H_ */
