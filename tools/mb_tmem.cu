// Microbenchmark (dev tool): tcgen05.ld / tcgen05.st .32x32b.x16 throughput when tensor memory is used as
// thread-private parking space (no MMA), alone and next to FFMA2 work.  2 CTAs x 256 threads per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void tmem_st16(uint32_t a, const float (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(a), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]),
                   "f"(r[8]), "f"(r[9]), "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t a, float (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
                   "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
                 : "r"(a) : "memory");
}
// MODE 0: loads only; 1: stores only; 2: loads + 64 FFMA2 per load (independent); 3: FFMA2 only
template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float* out, int iters) {
    __shared__ uint32_t slot;
    const int t = threadIdx.x, warp = t >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
    float r[16], acc[16];
    float2 f[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) { r[j] = t * 0.001f + j; acc[j] = 0.f; }
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = make_float2(t * 0.01f, j);
    for (int c = 0; c < 8; ++c) tmem_st16(base + 16 * c, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (MODE == 0 || MODE == 2) {
                tmem_ld16(base + 16 * c, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] += r[j];
            }
            if (MODE == 1) {
                tmem_st16(base + 16 * c, acc);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                acc[c] += 1.f;
            }
            if (MODE == 2 || MODE == 3) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = __ffma2_rn(f[j], make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += acc[j];
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[j].x + f[j].y;
    out[blockIdx.x * 256 + t] = s;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(slot) : "memory");
}
template <int MODE> void run(const char* name) {
    float* out; cudaMalloc(&out, 4 * 296 * 256);
    const int iters = 2000;
    k<MODE><<<296, 256>>>(out, iters);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<296, 256>>>(out, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    const double ops = (double)iters * 8;                 // x16 transfers per thread
    const double bytes_sm = ops * 64 * 512;               // per SM: 512 threads x 64 B per transfer
    const double clk = ms * 1e-3 * 1.965e9;
    printf("%-28s %.3f ms  %.1f B/clk/SM  (%.1f clk per warp-level x16 transfer per SM)  ffma2 lane-ops/clk/SM %.1f  [%s]\n", name, ms,
           bytes_sm / clk, clk / (ops * 16), (MODE >= 2) ? ops * 64 * 2 * 512 / clk : 0.0, cudaGetErrorString(e));
    cudaFree(out);
}
int main() {
    run<0>("LDTM.x16 only");
    run<1>("STTM.x16 only");
    run<3>("FFMA2 only");
    run<2>("LDTM.x16 + 64 FFMA2 each");
    return 0;
}
