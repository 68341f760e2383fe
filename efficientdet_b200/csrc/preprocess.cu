// Letterbox resize on the device: utils/__init__.py:103-132 resize_image (== generators/common.py:406-417 and the
// preprocess step of inference.py / predict.py), i.e. cv2.resize(image, (rw, rh)) -- bilinear, uint8 -- pasted into
// the centre of a grey (128) square.  SURVEY section 8(f) rank 4, second half (the normalisation is in the stem).
//
// Bit-exact with OpenCV's INTER_LINEAR for 8-bit images (the reference pins opencv-python 3.4.2.17; the arithmetic
// is unchanged through 4.x): coordinates fx = (float)((dx + 0.5) * scale - 0.5) with scale = 1 / (dst / src) in
// double, 11-bit fixed-point weights saturate_cast<short>(w * 2048) (round half to even), horizontal pass in int32,
// vertical pass ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.  Horizontally OpenCV clamps the
// coordinate (fx = 0 at the borders); vertically it keeps the weights and clips the two row indices.
// This translation unit is compiled with -fmad=false (build.py): the coordinate arithmetic must not be contracted.
#include "common.cuh"

namespace effdet {

struct LbCoef { int s0, s1; int a0, a1; };

// one axis: destination index d -> source indices and the two 11-bit weights
__device__ __forceinline__ LbCoef lb_coef(int d, int ssize, double scale, bool clamp_coord) {
    float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    LbCoef c;
    if (clamp_coord) {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
        c.s0 = s; c.s1 = min(s + 1, ssize - 1);
    } else {
        c.s0 = min(max(s, 0), ssize - 1); c.s1 = min(max(s + 1, 0), ssize - 1);
    }
    c.a0 = __float2int_rn((1.f - f) * 2048.f);
    c.a1 = __float2int_rn(f * 2048.f);
    return c;
}

__global__ void __launch_bounds__(256)
letterbox_u8_kernel(const uint8_t *__restrict__ src, int sh, int sw, long long src_row_stride,
                    uint8_t *__restrict__ dst, int S, int rh, int rw, int off_h, int off_w,
                    double scale_x, double scale_y) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= S) return;
    uint8_t *o = dst + ((size_t)y * S + x) * 3;
    const int dy = y - off_h, dx = x - off_w;
    if (dy < 0 || dy >= rh || dx < 0 || dx >= rw) { o[0] = 128; o[1] = 128; o[2] = 128; return; }
    if (rh == sh && rw == sw) {                      // cv2.resize to the same size copies
        const uint8_t *p = src + (size_t)dy * src_row_stride + (size_t)dx * 3;
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
        return;
    }
    const LbCoef cx = lb_coef(dx, sw, scale_x, true), cy = lb_coef(dy, sh, scale_y, false);
    const uint8_t *r0 = src + (size_t)cy.s0 * src_row_stride, *r1 = src + (size_t)cy.s1 * src_row_stride;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int h0 = (int)r0[cx.s0 * 3 + c] * cx.a0 + (int)r0[cx.s1 * 3 + c] * cx.a1;
        const int h1 = (int)r1[cx.s0 * 3 + c] * cx.a0 + (int)r1[cx.s1 * 3 + c] * cx.a1;
        const int v = (((cy.a0 * (h0 >> 4)) >> 16) + ((cy.a1 * (h1 >> 4)) >> 16) + 2) >> 2;
        o[c] = (uint8_t)min(max(v, 0), 255);
    }
}

}  // namespace effdet

using namespace effdet;

extern "C" int effdet_letterbox_geometry(int src_h, int src_w, int image_size, int *resized_h, int *resized_w,
                                         int *offset_h, int *offset_w, double *scale) {
    EFFDET_REQUIRE(src_h > 0 && src_w > 0 && image_size > 0, "bad arguments");
    int rh, rw;
    double sc;
    if (src_h == src_w && src_h == image_size) { rh = rw = image_size; sc = 0.0; }   // utils/__init__.py:108-109
    else if (src_h > src_w) { sc = (double)image_size / src_h; rh = image_size; rw = (int)(src_w * sc); }
    else { sc = (double)image_size / src_w; rh = (int)(src_h * sc); rw = image_size; }
    if (resized_h) *resized_h = rh;
    if (resized_w) *resized_w = rw;
    if (offset_h) *offset_h = (rh == image_size && rw == image_size && sc == 0.0) ? 0 : (image_size - rh) / 2;
    if (offset_w) *offset_w = (rh == image_size && rw == image_size && sc == 0.0) ? 0 : (image_size - rw) / 2;
    if (scale) *scale = sc;
    return EFFDET_OK;
}

extern "C" int effdet_letterbox_u8(const unsigned char *image, int src_h, int src_w, long long src_row_stride,
                                   unsigned char *out, int image_size, void *stream) {
    EFFDET_REQUIRE(image && out && src_h > 0 && src_w > 0 && image_size > 0, "bad arguments");
    int rh, rw, oh, ow;
    double sc;
    effdet_letterbox_geometry(src_h, src_w, image_size, &rh, &rw, &oh, &ow, &sc);
    EFFDET_REQUIRE(rh > 0 && rw > 0, "image too thin for this image_size");
    if (src_row_stride == 0) src_row_stride = (long long)src_w * 3;
    // OpenCV: inv_scale = dsize / ssize (double), scale = 1 / inv_scale
    const double scale_x = 1.0 / ((double)rw / (double)src_w), scale_y = 1.0 / ((double)rh / (double)src_h);
    dim3 grid(cdiv((size_t)image_size, 256), image_size);
    letterbox_u8_kernel<<<grid, 256, 0, as_stream(stream)>>>(image, src_h, src_w, src_row_stride, out, image_size, rh, rw,
                                                             oh, ow, scale_x, scale_y);
    EFFDET_CUDA(cudaGetLastError());
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}
