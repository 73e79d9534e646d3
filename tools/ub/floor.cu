// Where does the ~6 us floor of a tcgen05 launch go?  A chain of dependent launches (plain stream, PDL attribute like the
// real kernels, captured in a CUDA graph as well) of kernels that do progressively more of the conv_tc skeleton:
//   L0 empty                       L1 + mbarrier init, TMEM alloc/dealloc, __syncthreads
//   L2 + one TMA load (A 16 KB + B 8 KB) and wait        L3 + 4 UMMAs, commit, wait
//   L4 + tcgen05.ld, st.shared staging, fence, TMA store, wait_group.read      L5 = L4 with the tensor maps prefetched
//   L6 = L3 + tcgen05.ld and direct 16-byte st.global from registers (no staging, no TMA store)
// Build: nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -o tools_ub/floor tools_ub/floor.cu -lcuda
#include "../diffusionmodelscustom_b200/csrc/conv.cuh"
namespace b2d { thread_local Status g_status; int g_pdl_enabled = 1; }
using namespace b2d;

template <int LEVEL>
__global__ void __launch_bounds__(192, 1) k_floor(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                  const __grid_constant__ CUtensorMap tmO, f16* gout) {
    pdl_launch_dependents();
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + CONV_A_BYTES;
    uint8_t* sO = sB + 8192;
    uint64_t* full = reinterpret_cast<uint64_t*>(sO + CONV_A_BYTES);
    uint64_t* acc = full + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(full + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (LEVEL == 5 && threadIdx.x == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmO); }
    if (LEVEL >= 1) {
        if (warp == 1) {
            if (lane == 0) { mbar_init(full, 1); mbar_init(acc, 1); fence_mbar_init(); }
            __syncwarp();
            tmem_alloc(slot, 64);
            tmem_relinquish();
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    pdl_wait();
    if (LEVEL >= 2) {
        const uint32_t tmem = *slot;
        if (warp == 0 && lane == 0) {
            mbar_arrive_expect_tx(full, CONV_A_BYTES + 8192);
            tma_load_2d(sA, &tmA, full, 0, blockIdx.x * 128);
            tma_load_2d(sB, &tmB, full, 0, 0);
        }
        if (warp == 1 && lane == 0) {
            mbar_wait(full, 0);
            if (LEVEL >= 3) {
                tc_fence_after();
                const uint64_t da = umma_desc_sw128(smem_u32(sA)), db = umma_desc_sw128(smem_u32(sB));
                for (int k = 0; k < 4; ++k) umma_f16(tmem, da + k * 2, db + k * 2, umma_idesc_f16(128, 64), k != 0);
                umma_commit(acc);
            }
        }
        if (LEVEL >= 3 && warp >= 2) {
            mbar_wait(acc, 0);
            tc_fence_after();
            if (LEVEL == 6) {
                const int q = warp & 3, row = q * 32 + lane;
                uint4* op = reinterpret_cast<uint4*>(gout + ((size_t)blockIdx.x * 128 + row) * 64);
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + h * 32, v);
                    tmem_ld_wait();
                    for (int j = 0; j < 4; ++j) {
                        uint4 o;
                        o.x = pack_h2(__uint_as_float(v[j * 8]), __uint_as_float(v[j * 8 + 1]));
                        o.y = pack_h2(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3]));
                        o.z = pack_h2(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5]));
                        o.w = pack_h2(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7]));
                        op[h * 4 + j] = o;
                    }
                }
            } else if (LEVEL >= 4) {
                const int q = warp & 3, row = q * 32 + lane;
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + h * 32, v);
                    tmem_ld_wait();
                    for (int j = 0; j < 4; ++j) {
                        uint4 o;
                        o.x = pack_h2(__uint_as_float(v[j * 8]), __uint_as_float(v[j * 8 + 1]));
                        o.y = pack_h2(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3]));
                        o.z = pack_h2(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5]));
                        o.w = pack_h2(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7]));
                        *reinterpret_cast<uint4*>(sO + row * 128 + (((h * 4 + j) ^ (row & 7)) << 4)) = o;
                    }
                }
                fence_proxy_async();
                named_bar_sync(1, 128);
                if (threadIdx.x == 64) {
                    tma_store_2d(&tmO, sO, 0, blockIdx.x * 128);
                    tma_store_commit();
                    tma_store_wait_read();
                }
            }
        }
    }
    if (LEVEL >= 1) {
        tc_fence_before();
        __syncthreads();
        if (warp == 1) { tc_fence_after(); tmem_dealloc(*slot, 64); }
    }
}

template <int LEVEL>
static void run(const char* name, CUtensorMap a, CUtensorMap b, CUtensorMap o, int ctas, f16* gout) {
    const int smem = 1024 + CONV_A_BYTES * 2 + 8192 + 64;
    cudaFuncSetAttribute(k_floor<LEVEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int N = 200;
    for (int pdl = 1; pdl >= 0; --pdl) {
        g_pdl_enabled = pdl;
        for (int i = 0; i < 20; ++i) launch_k(k_floor<LEVEL>, dim3(ctas), dim3(192), smem, st, a, b, o, gout);
        cudaEventRecord(e0, st);
        for (int i = 0; i < N; ++i) launch_k(k_floor<LEVEL>, dim3(ctas), dim3(192), smem, st, a, b, o, gout);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaGraph_t g; cudaGraphExec_t ge;
        cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        for (int i = 0; i < N; ++i) launch_k(k_floor<LEVEL>, dim3(ctas), dim3(192), smem, st, a, b, o, gout);
        cudaStreamEndCapture(st, &g);
        cudaGraphInstantiate(&ge, g, 0);
        cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
        cudaEventRecord(e0, st); cudaGraphLaunch(ge, st); cudaEventRecord(e1, st); cudaEventSynchronize(e1);
        float msg; cudaEventElapsedTime(&msg, e0, e1);
        printf("%-46s ctas %3d pdl %d: stream %.2f us/launch, graph %.2f us/launch  (%s)\n", name, ctas, pdl, ms * 1e3 / N, msg * 1e3 / N,
               cudaGetErrorString(cudaGetLastError()));
        cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
    }
}

int main() {
    const int M = 128 * 148;
    f16 *A, *B, *O;
    cudaMalloc(&A, (size_t)M * 64 * 2); cudaMalloc(&B, 64 * 64 * 2); cudaMalloc(&O, (size_t)M * 64 * 2);
    cudaMemset(A, 0, (size_t)M * 64 * 2); cudaMemset(B, 0, 64 * 64 * 2);
    CUtensorMap ta, tb, to;
    uint64_t ad[2] = {64, (uint64_t)M}; uint64_t as[1] = {128}; uint32_t ab[2] = {64, 128};
    uint64_t bd[2] = {64, 64}; uint32_t bb[2] = {64, 64};
    if (make_tmap_f16(&ta, A, 2, ad, as, ab) || make_tmap_f16(&tb, B, 2, bd, as, bb) || make_tmap_f16(&to, O, 2, ad, as, ab)) {
        printf("tensor map failed: %s\n", g_status.msg.c_str());
        return 1;
    }
    for (int ctas : {128}) {
        run<0>("L0 empty", ta, tb, to, ctas, O);
        run<1>("L1 + barriers, TMEM alloc/dealloc", ta, tb, to, ctas, O);
        run<2>("L2 + TMA load A,B + wait", ta, tb, to, ctas, O);
        run<3>("L3 + 4 UMMA + commit + wait", ta, tb, to, ctas, O);
        run<4>("L4 + tcgen05.ld, staging, TMA store", ta, tb, to, ctas, O);
        run<5>("L5 = L4 + tensor maps prefetched", ta, tb, to, ctas, O);
        run<6>("L6 = L3 + tcgen05.ld, direct st.global", ta, tb, to, ctas, O);
    }
    return 0;
}
