// Microbenchmark of the stft_kernel input access pattern (dev tool): each 256-thread CTA walks a 1 MB
// region frame by frame (4096 sf window, hop 2048), 16 x 8-byte loads per thread per frame.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ld_stream(const float2* p) {
    float2 r; asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p)); return r; }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// MODE 0: 16 loads (full frame) no_allocate; 1: + L2 prefetch of next new half; 2: default-cached loads;
// 3: only the new half (8 loads), old half kept in registers; 4: new half only + prefetch 2 frames ahead
template <int MODE>
__global__ void __launch_bounds__(256, 2) k(const float2* in, float2* out, int units, int frames_per_unit, int do_store) {
    const int t = threadIdx.x;
    float2 acc = make_float2(0, 0);
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const float2* base = in + (size_t)u * frames_per_unit * 2048;
        float2 old[8];
        for (int j = 0; j < 8; ++j) old[j] = make_float2(0, 0);
        for (int f = 0; f < frames_per_unit - 1; ++f) {
            const float2* src = base + (size_t)f * 2048 + t;
            float2 v[16];
            if (MODE == 3 || MODE == 4) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { v[j] = old[j]; v[j + 8] = ld_stream(src + 256 * (j + 8)); }
                if (MODE == 4 && (t & 15) == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) prefetch_l2(src + 256 * (j + 8) + 2 * 2048);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = (MODE == 2) ? __ldg(src + 256 * j) : ld_stream(src + 256 * j);
                if (MODE == 1 && (t & 15) == 0) {
#pragma unroll
                    for (int j = 16; j < 24; ++j) prefetch_l2(src + 256 * j);
                }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) { acc.x += v[j].x * 1.0001f; acc.y += v[j].y * 0.9999f; }
#pragma unroll
            for (int j = 0; j < 8; ++j) old[j] = v[j + 8];
            if (do_store) {
                float2* dst = out + (size_t)u * frames_per_unit * 2048 + (size_t)f * 2048 + t;
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[256 * j] = make_float2(v[j].x + acc.x, v[j].y);
            }
        }
    }
    if (acc.x == 123.456f) out[t] = acc;
}
template <int MODE> void run(const char* name, const float2* in, float2* out, int units, int fpu, int grid, int do_store) {
    k<MODE><<<grid, 256>>>(in, out, units, fpu, do_store);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<grid, 256>>>(in, out, units, fpu, do_store);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double sf = (double)units * (fpu - 1) * 2048;
    printf("%-44s grid=%4d store=%d  %.3f ms  %.1f Gsf/s  new-bytes %.0f GB/s\n", name, grid, do_store, ms, sf / ms / 1e6, sf * 8 * (1 + do_store) / ms / 1e6);
}
int main() {
    const int units = 3520, fpu = 60;                       // ~ 32 tracks worth
    size_t n = (size_t)units * fpu * 2048;
    float2 *in, *out; cudaMalloc(&in, n * 8); cudaMalloc(&out, n * 8); cudaMemset(in, 0, n * 8);
    for (int store = 0; store < 2; ++store)
        for (int grid : {148, 296}) {
            run<0>("16 loads no_allocate", in, out, units, fpu, grid, store);
            run<1>("16 loads + L2 prefetch next", in, out, units, fpu, grid, store);
            run<2>("16 loads default cache (__ldg)", in, out, units, fpu, grid, store);
            run<3>("8 new loads, old half in regs", in, out, units, fpu, grid, store);
            run<4>("8 new loads + L2 prefetch 2 ahead", in, out, units, fpu, grid, store);
        }
    return 0;
}
