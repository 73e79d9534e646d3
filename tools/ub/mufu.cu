// Microbenchmarks: SFU ex2 throughput (fp32 and f16x2), FFMA throughput, tcgen05.ld throughput is not covered here.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ex2(float* out, int iters) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = a[i] - 1.0f;
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.f) out[0] = s;
}
__global__ void k_ex2h(float* out, int iters) {
    unsigned a[8];
    for (int i = 0; i < 8; ++i) a[i] = 0xB000B000u + threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] ^= 0x80008000u;
    }
    unsigned s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345u) out[0] = s;
}
__global__ void k_ffma(float* out, int iters) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = 0.001f * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], 1.0001f, 0.5f);
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.f) out[0] = s;
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float* d; cudaMalloc(&d, 4);
    const int iters = 4096, blocks = 148 * 8, threads = 256;
    double n = (double)blocks * threads * iters * 8;
    float ms = timeit([&] { k_ex2<<<blocks, threads>>>(d, iters); });
    printf("ex2.f32   : %.2f Tops/s (%.2f /clk/SM @1.9GHz)\n", n / ms / 1e9, n / ms / 1e9 * 1e12 / 148 / 1.9e9 / 1e12 * 1e0);
    ms = timeit([&] { k_ex2h<<<blocks, threads>>>(d, iters); });
    printf("ex2.f16x2 : %.2f T instr/s = %.2f T exps/s\n", n / ms / 1e9, 2 * n / ms / 1e9);
    ms = timeit([&] { k_ffma<<<blocks, threads>>>(d, iters); });
    printf("ffma      : %.2f T instr/s\n", n / ms / 1e9);
    return 0;
}
