// Micro-benchmark: issue rate of packed FFMA2 vs scalar FFMA on sm_100a (per SM, per clock).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float *out, int iters, float a, float b) {
    float2 acc[8];
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    float2 x = make_float2(a, b), w = make_float2(b, a);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) acc[i] = __ffma2_rn(acc[i], x, w);
                else { acc[i].x = fmaf(acc[i].x, x.x, w.x); acc[i].y = fmaf(acc[i].y, x.y, w.y); }
            }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
int main() {
    float *d; cudaMalloc(&d, 148 * 1024 * 4);
    const int iters = 2000;
    for (int mode = 0; mode < 2; ++mode)
        for (int nt : {128, 256, 512, 1024}) {
            if (mode == 0) k<0><<<148, nt>>>(d, iters, 1.0001f, 0.5f); else k<1><<<148, nt>>>(d, iters, 1.0001f, 0.5f);
            cudaDeviceSynchronize();
            float clk; cudaMemcpy(&clk, d, 4, cudaMemcpyDeviceToHost);
            double fma = (double)iters * 64 * 2 * nt;      // scalar FMAs per block (= per SM)
            printf("%s threads/SM %4d: %.1f clk, %.1f FMA/clk/SM (%.2f warp-inst/clk/SMSP)\n", mode ? "FFMA " : "FFMA2", nt,
                   clk, fma / clk, fma / clk / 4 / (mode ? 32 : 64));
        }
    return 0;
}
