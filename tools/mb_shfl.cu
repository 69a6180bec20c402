// Microbenchmark (dev tool): do warp shuffles contend with shared-memory loads/stores for the same SM data path?
// MODE 0: LDS.64+STS.64 only, 1: SHFL only, 2: both interleaved, 3: FFMA2 only, 4: SHFL + FFMA2, 5: LDS/STS + FFMA2
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float2* out, int iters) {
    __shared__ float2 sm[256 * 17];
    const int t = threadIdx.x;
    float2 v[8], f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[j] = make_float2(t + j, t - j); f[j] = make_float2(t * 0.01f, j); }
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2 || MODE == 5) {
#pragma unroll
            for (int j = 0; j < 8; ++j) sm[((t >> 4) * 16 + j) * 17 + (t & 15)] = v[j];
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = sm[t * 17 + j];
            __syncwarp();
        }
        if (MODE == 1 || MODE == 2 || MODE == 4) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[j].x = __shfl_xor_sync(0xffffffffu, v[j].x, 1 + (j & 7));
                v[j].y = __shfl_xor_sync(0xffffffffu, v[j].y, 1 + (j & 7));
            }
        }
        if (MODE >= 3) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = __ffma2_rn(f[j], make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f));
        }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s.x += v[j].x + f[j].x; s.y += v[j].y + f[j].y; }
    out[blockIdx.x * 256 + t] = s;
}
template <int MODE> void run(const char* name) {
    float2* out; cudaMalloc(&out, 8 * 296 * 256);
    const int iters = 4000;
    k<MODE><<<296, 256>>>(out, iters);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<296, 256>>>(out, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double clk = ms * 1e-3 * 1.965e9 / iters;     // SM cycles per iteration (16 warps resident)
    printf("%-36s %.3f ms  %.1f SM-clk per iteration  (per warp-iteration: %.1f)\n", name, ms, clk, clk / 16);
    cudaFree(out);
}
int main() {
    run<0>("8 STS.64 + 8 LDS.64");
    run<1>("16 SHFL.32");
    run<2>("8 STS.64 + 8 LDS.64 + 16 SHFL.32");
    run<3>("32 FFMA2");
    run<4>("16 SHFL.32 + 32 FFMA2");
    run<5>("8 STS.64 + 8 LDS.64 + 32 FFMA2");
    return 0;
}
