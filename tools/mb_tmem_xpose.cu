// Microbenchmark (dev tool): a 16 x 16 transpose of float2 among the 16 lanes that share lane bit 0 of a warp, done through
// TENSOR MEMORY instead of shared memory.  tcgen05.st/ld .32x32b gives (lane i, column c) <-> (thread i, register c);
// .16x256b gives (lane r, column c) <-> (thread 4*(r%8) + (c%8)/2, register 4*(c/8) + 2*(r/8) + c%2)  [CuTe copy traits of
// SM100_TMEM_LOAD_16dp256b]: a store in one shape followed by a load in the other swaps two thread bits with two register
// bits, so two round trips make the 4-bit exchange a radix-16 FFT stage boundary needs -- with no shared-memory wavefronts.
//   part 1: dump of the data movement (checked on the host against the predicted permutation, both directions);
//   part 2: cycles per exchange (2 CTAs x 256 threads per SM), alone and under FFMA2 work, next to the shared-memory version.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define TM_ST32(a, r, o) \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
                 ::"r"(a), "f"(r[o+0]), "f"(r[o+1]), "f"(r[o+2]), "f"(r[o+3]), "f"(r[o+4]), "f"(r[o+5]), "f"(r[o+6]), "f"(r[o+7]), \
                   "f"(r[o+8]), "f"(r[o+9]), "f"(r[o+10]), "f"(r[o+11]), "f"(r[o+12]), "f"(r[o+13]), "f"(r[o+14]), "f"(r[o+15]) : "memory")
#define TM_LD32(a, r, o) \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                 : "=f"(r[o+0]), "=f"(r[o+1]), "=f"(r[o+2]), "=f"(r[o+3]), "=f"(r[o+4]), "=f"(r[o+5]), "=f"(r[o+6]), "=f"(r[o+7]), \
                   "=f"(r[o+8]), "=f"(r[o+9]), "=f"(r[o+10]), "=f"(r[o+11]), "=f"(r[o+12]), "=f"(r[o+13]), "=f"(r[o+14]), "=f"(r[o+15]) \
                 : "r"(a) : "memory")
#define TM_ST256(a, r, o) \
    asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
                 ::"r"(a), "f"(r[o+0]), "f"(r[o+1]), "f"(r[o+2]), "f"(r[o+3]), "f"(r[o+4]), "f"(r[o+5]), "f"(r[o+6]), "f"(r[o+7]), \
                   "f"(r[o+8]), "f"(r[o+9]), "f"(r[o+10]), "f"(r[o+11]), "f"(r[o+12]), "f"(r[o+13]), "f"(r[o+14]), "f"(r[o+15]) : "memory")
#define TM_LD256(a, r, o) \
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                 : "=f"(r[o+0]), "=f"(r[o+1]), "=f"(r[o+2]), "=f"(r[o+3]), "=f"(r[o+4]), "=f"(r[o+5]), "=f"(r[o+6]), "=f"(r[o+7]), \
                   "=f"(r[o+8]), "=f"(r[o+9]), "=f"(r[o+10]), "=f"(r[o+11]), "=f"(r[o+12]), "=f"(r[o+13]), "=f"(r[o+14]), "=f"(r[o+15]) \
                 : "r"(a) : "memory")
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// one round trip "forward": thread-major store, 16-lane-row load (two instructions: lanes 0-15 and 16-31 of the warp's quarter)
__device__ __forceinline__ void pass_fwd(uint32_t base, float (&r)[32]) {
    TM_ST32(base, r, 0);
    TM_ST32(base + 16, r, 16);
    wait_st();
    TM_LD256(base, r, 0);
    TM_LD256(base + (16u << 16), r, 16);
    wait_ld();
}
// the exact inverse data movement
__device__ __forceinline__ void pass_inv(uint32_t base, float (&r)[32]) {
    TM_ST256(base, r, 0);
    TM_ST256(base + (16u << 16), r, 16);
    wait_st();
    TM_LD32(base, r, 0);
    TM_LD32(base + 16, r, 16);
    wait_ld();
}

__device__ __forceinline__ uint32_t tmem_setup(uint32_t* slot, int warp) {
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"((uint32_t)__cvta_generic_to_shared(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
}
__device__ __forceinline__ void tmem_release(uint32_t* slot, int warp) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(*slot) : "memory");
}

// part 1: out[stage][thread][reg] = the tag (source thread * 32 + source register) found there after
// stage 0: one forward pass, 1: second forward pass (regs re-ordered in between as the FFT exchange does), 2: first inverse pass,
// 3: second inverse pass (must be the identity again)
__global__ void __launch_bounds__(256, 2) dump_kernel(float* out) {
    __shared__ uint32_t slot;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t base = tmem_setup(&slot, warp);
    float r[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) r[c] = (float)(lane * 32 + c);
    pass_fwd(base, r);
    if (warp == 5) for (int c = 0; c < 32; ++c) out[(0 * 32 + lane) * 32 + c] = r[c];
    // received register 16*I + 4*n + 2*h + e = element (m = 2*I + h, n), component e; second pass wants float2 column q' = 4*m + n
    float s[32];
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int e = 0; e < 2; ++e) s[2 * (4 * (2 * I + h) + n) + e] = r[16 * I + 4 * n + 2 * h + e];
    pass_fwd(base, s);
    if (warp == 5) for (int c = 0; c < 32; ++c) out[(1 * 32 + lane) * 32 + c] = s[c];
    pass_inv(base, s);
    if (warp == 5) for (int c = 0; c < 32; ++c) out[(2 * 32 + lane) * 32 + c] = s[c];
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int e = 0; e < 2; ++e) r[16 * I + 4 * n + 2 * h + e] = s[2 * (4 * (2 * I + h) + n) + e];
    pass_inv(base, r);
    if (warp == 5) for (int c = 0; c < 32; ++c) out[(3 * 32 + lane) * 32 + c] = r[c];
    tmem_release(&slot, warp);
}

// part 2.  MODE 0: tensor-memory exchange (forward two passes + inverse two passes per iteration = 2 exchanges);
// 1: shared-memory exchange of the current kernel (16 STS.64 + 8 LDS.128, then 8 STS.128 + 16 LDS.64 = 2 exchanges);
// 2 / 3: the same with NF FFMA2 per exchange in between; 4: FFMA2 only
template <int MODE, int NF>
__global__ void __launch_bounds__(256, 2) time_kernel(float* out, int iters) {
    __shared__ uint32_t slot;
    extern __shared__ __align__(16) float2 ex[];          // 256 rows x 18 float2 (the kernel's padded E2 buffer)
    const int t = threadIdx.x, warp = t >> 5;
    const uint32_t base = tmem_setup(&slot, warp);
    float r[32];
    float2 f[8];
#pragma unroll
    for (int c = 0; c < 32; ++c) r[c] = (float)(t * 32 + c) * 1e-6f;
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = make_float2(t * 0.01f, j);
    auto work = [&]() {
#pragma unroll
        for (int q = 0; q < NF / 8; ++q)
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __ffma2_rn(f[j], make_float2(1.0001f, 0.9999f), make_float2(r[2 * j], r[2 * j + 1]));
#pragma unroll
        for (int j = 0; j < 8; ++j) { r[2 * j] += f[j].x * 1e-9f; r[17 + 2 * (j & 6)] += f[j].y * 1e-9f; }
    };
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
            pass_fwd(base, r);
            pass_fwd(base, r);
            if (MODE == 2) work();
            pass_inv(base, r);
            pass_inv(base, r);
            if (MODE == 2) work();
        } else if (MODE == 1 || MODE == 3) {
            float2* rowb = ex + ((t >> 4) * 16) * 18 + (t & 15);
#pragma unroll
            for (int j = 0; j < 16; ++j) rowb[j * 18] = make_float2(r[2 * j], r[2 * j + 1]);
            __syncwarp();
            float4* rowc = reinterpret_cast<float4*>(ex + t * 18);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float4 x = rowc[j]; r[4 * j] = x.x; r[4 * j + 1] = x.y; r[4 * j + 2] = x.z; r[4 * j + 3] = x.w; }
            if (MODE == 3) work();
#pragma unroll
            for (int j = 0; j < 8; ++j) rowc[j] = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 16; ++j) { const float2 x = rowb[j * 18]; r[2 * j] = x.x; r[2 * j + 1] = x.y; }
            if (MODE == 3) work();
            __syncwarp();
        } else {
            work();
            work();
        }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 32; ++c) s += r[c];
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[j].x + f[j].y;
    out[blockIdx.x * 256 + t] = s;
    tmem_release(&slot, warp);
}

template <int MODE, int NF> void run(const char* name, int grid = 296, int threads = 256) {
    float* out;
    cudaMalloc(&out, 4 * 296 * 256);
    const int iters = 2000, smem = 256 * 18 * 8;
    cudaFuncSetAttribute(time_kernel<MODE, NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    time_kernel<MODE, NF><<<grid, threads, smem>>>(out, iters);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    time_kernel<MODE, NF><<<grid, threads, smem>>>(out, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const cudaError_t e = cudaGetLastError();
    const double clk = ms * 1e-3 * 1.965e9;
    // per SM: 16 warps, each does 2 exchanges per iteration
    if (threads == 32)          // one warp per SM: pure latency of the dependent chain
        printf("%-52s %.3f ms   %.1f clk per exchange, one warp alone on the SM (latency)  [%s]\n", name, ms, clk / (iters * 2.0), cudaGetErrorString(e));
    else
    printf("%-52s %.3f ms   %.1f SM-clk per warp-level exchange   (FFMA2 lane-ops/clk/SM %.1f)  [%s]\n", name, ms, clk / (iters * 2.0 * 16.0),
           (NF > 0) ? (double)iters * 2 * NF * 64 * 16 / clk : 0.0, cudaGetErrorString(e));
    cudaFree(out);
}

int main() {
    float* d;
    cudaMalloc(&d, 4 * 32 * 32 * 4);
    dump_kernel<<<2, 256>>>(d);
    std::vector<float> h(4 * 32 * 32);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    printf("dump: %s\n", cudaGetErrorString(cudaGetLastError()));
    // predicted: after one forward pass thread T register 16*I + 4*n + 2*h + e holds (source lane T/4 + 8*(2*I + h), source register 2*(4*n + T%4) + e)
    int bad0 = 0, bad1 = 0, bad2 = 0, bad3 = 0;
    for (int T = 0; T < 32; ++T)
        for (int I = 0; I < 2; ++I) for (int n = 0; n < 4; ++n) for (int hh = 0; hh < 2; ++hh) for (int e = 0; e < 2; ++e) {
            const int reg = 16 * I + 4 * n + 2 * hh + e;
            const int want = (T / 4 + 8 * (2 * I + hh)) * 32 + 2 * (4 * n + T % 4) + e;
            if ((int)h[(0 * 32 + T) * 32 + reg] != want) { if (bad0 < 4) printf("  pass1 T=%d reg=%d got %d want %d\n", T, reg, (int)h[(0 * 32 + T) * 32 + reg], want); ++bad0; }
            // after the second pass: thread T2 element (m', n') = register 2*(4*m' + n')... received as 16*I + 4*n' + 2*h + e with m' = 2*I + h:
            // source lane = T2/16 + 2*m' + 8*n', source float2 index q = 4*(T2%4) + (T2/4)%4
            const int mp = 2 * I + hh;
            const int want2 = (T / 16 + 2 * mp + 8 * n) * 32 + 2 * (4 * (T % 4) + (T / 4) % 4) + e;
            if ((int)h[(1 * 32 + T) * 32 + reg] != want2) { if (bad1 < 4) printf("  pass2 T=%d reg=%d got %d want %d\n", T, reg, (int)h[(1 * 32 + T) * 32 + reg], want2); ++bad1; }
        }
    for (int T = 0; T < 32; ++T) for (int c = 0; c < 32; ++c) {
        // after the first inverse pass the registers are those after forward pass 1, in the re-ordered (column) naming
        const int q = c / 2, e = c % 2, m = q / 4, n = q % 4;
        const int want = (T / 4 + 8 * m) * 32 + 2 * (4 * n + T % 4) + e;
        if ((int)h[(2 * 32 + T) * 32 + c] != want) { if (bad2 < 4) printf("  inv1 T=%d reg=%d got %d want %d\n", T, c, (int)h[(2 * 32 + T) * 32 + c], want); ++bad2; }
        if ((int)h[(3 * 32 + T) * 32 + c] != T * 32 + c) { if (bad3 < 4) printf("  inv2 T=%d reg=%d got %d want %d\n", T, c, (int)h[(3 * 32 + T) * 32 + c], T * 32 + c); ++bad3; }
    }
    printf("layout check: forward pass 1 mismatches %d, pass 2 %d, inverse pass 1 %d, inverse pass 2 (identity) %d\n", bad0, bad1, bad2, bad3);
    if (bad0) { printf("raw dump of pass 1, threads 0..3:\n"); for (int T = 0; T < 4; ++T) { for (int c = 0; c < 32; ++c) printf("%d ", (int)h[T * 32 + c]); printf("\n"); } }
    run<0, 0>("tensor-memory exchange (2 round trips)");
    run<1, 0>("shared-memory exchange (16 STS.64 + 8 LDS.128)");
    run<4, 96>("96 FFMA2 only");
    run<2, 96>("tensor-memory exchange + 96 FFMA2");
    run<3, 96>("shared-memory exchange + 96 FFMA2");
    run<0, 0>("tensor-memory exchange, latency", 148, 32);
    run<3, 8>("shared-memory exchange (+8 FFMA2), latency", 148, 32);
    run<2, 8>("tensor-memory exchange (+8 FFMA2), latency", 148, 32);
    run<4, 8>("8 FFMA2 only, latency", 148, 32);
    return 0;
}
