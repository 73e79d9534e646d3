// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM, alone and next to MUFU.EX2 / FMA-pipe work.
// Decides what bounds the head_dim-16 attention kernel (attention_tc.cuh): S is 4 B per score in TMEM and costs one exponential.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ub/tmem_ld tools/ub/tmem_ld.cu && tools/ub/tmem_ld
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MODE 0: loads only; 1: loads + one ex2 per loaded word (the attention ratio); 2: ex2 only (same count); 3: loads + 4 FFMA per word
template <int MODE>
__global__ void __launch_bounds__(256) k_ldtm(unsigned long long* cyc, float* sink, int iters, int cols) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    float acc = 0.f;
    float e[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) e[i] = -0.001f * (threadIdx.x + i);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t v[32];
        if (MODE != 2) {
            tmem_ld32(tmem + lane_off + (uint32_t)((it * 32) % cols), v);
            tmem_ld_wait();
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float x = (MODE == 1) ? __uint_as_float(v[i] & 0x3fffffffu) * -1e-30f + e[i] : e[i];
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(x));
                e[i] = x - 1.0f;
            }
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float x = __uint_as_float(v[i] & 0x3fffffffu);
                x = fmaf(x, 1.0001f, e[i]); x = fmaf(x, 0.5f, 0.25f); x = fmaf(x, x, 0.125f); e[i] = fmaf(x, 1e-30f, -0.001f);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += __uint_as_float(v[i] & 1u);
        }
    }
    const long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) acc += e[i];
    if (acc == 12345.f) sink[0] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols) : "memory");
    }
}

template <int MODE>
static void run(const char* name, int warps, int ctas_per_sm, int cols) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * ctas_per_sm, iters = 4096;
    unsigned long long* d_cyc;
    float* d_sink;
    cudaMalloc(&d_cyc, blocks * 8);
    cudaMalloc(&d_sink, 4);
    k_ldtm<MODE><<<blocks, warps * 32>>>(d_cyc, d_sink, 64, cols);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_ldtm<MODE><<<blocks, warps * 32>>>(d_cyc, d_sink, iters, cols);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[148 * 8], mx = 0;
    cudaMemcpy(h, d_cyc, blocks * 8, cudaMemcpyDeviceToHost);
    for (int i = 0; i < blocks; ++i) mx = h[i] > mx ? h[i] : mx;
    const double words_per_sm = (double)ctas_per_sm * warps * 32 * 32.0 * iters;
    printf("%-28s warps/CTA=%d CTAs/SM=%d: %8llu clk  -> %6.1f B/clk/SM (ld)  %5.2f words(exp)/clk/SM   [%.3f ms, %s]\n", name, warps,
           ctas_per_sm, mx, words_per_sm * 4.0 / mx, words_per_sm / mx, ms, cudaGetErrorString(err));
    cudaFree(d_cyc);
    cudaFree(d_sink);
}

int main() {
    for (int w : {4, 8}) {
        for (int c : {1, 2, 3}) {
            const int cols = c == 3 ? 128 : 256;
            run<0>("tcgen05.ld x32 only", w, c, cols);
        }
    }
    for (int c : {1, 2, 3}) run<2>("ex2 only", 4, c, 128);
    for (int c : {1, 2, 3}) run<1>("ld + 1 ex2/word", 4, c, 128);
    run<1>("ld + 1 ex2/word", 8, 2, 128);
    for (int c : {1, 2, 3}) run<3>("ld + 4 FFMA/word", 4, c, 128);
    return 0;
}
