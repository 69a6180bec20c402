// Microbenchmark: scalar FADD/FFMA vs packed FADD2/FFMA2 throughput on sm_100a (dev tool).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float2* out, float2 c, int iters) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x += c.x; a[i].y += c.y; }
            if (MODE == 1) { a[i] = __fadd2_rn(a[i], c); }
            if (MODE == 2) { a[i].x = fmaf(a[i].x, c.x, c.y); a[i].y = fmaf(a[i].y, c.x, c.y); }
            if (MODE == 3) { a[i] = __ffma2_rn(a[i], c, c); }
            if (MODE == 4) { a[i].x *= c.x; a[i].y *= c.y; }
            if (MODE == 5) { a[i] = __fmul2_rn(a[i], c); }
            if (MODE == 6) { a[i] = __fadd2_rn(a[i], c); a[i].x = fmaf(a[i].x, c.x, c.y); }   // mix packed + scalar
        }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int blocks, int threads) {
    float2* out; cudaMalloc(&out, sizeof(float2) * blocks * threads);
    const int iters = 4096;
    k<MODE><<<blocks, threads>>>(out, make_float2(1.0001f, 0.5f), iters);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<blocks, threads>>>(out, make_float2(1.0001f, 0.5f), iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double lane_ops = (double)blocks * threads * iters * 16;   // 8 float2 = 16 scalar ops per iteration
    if (MODE == 6) lane_ops = (double)blocks * threads * iters * 24;
    printf("%-28s blocks=%d threads=%d  %.3f ms  %.1f Glane-op/s  (%.1f lane-ops/clk/SM @1.965GHz,148SM)\n", name, blocks, threads, ms,
           lane_ops / ms / 1e6, lane_ops / (ms * 1e-3) / 1.965e9 / 148);
    cudaFree(out);
}
int main() {
    for (int threads : {256, 512, 1024}) {
        run<0>("FADD  scalar", 148 * 2, threads);
        run<1>("FADD2 packed", 148 * 2, threads);
        run<2>("FFMA  scalar", 148 * 2, threads);
        run<3>("FFMA2 packed", 148 * 2, threads);
        run<4>("FMUL  scalar", 148 * 2, threads);
        run<5>("FMUL2 packed", 148 * 2, threads);
        run<6>("FADD2+FFMA mix", 148 * 2, threads);
    }
    return 0;
}
