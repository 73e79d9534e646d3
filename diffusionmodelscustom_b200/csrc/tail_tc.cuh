// Decoder.final_layer on tcgen05 (modules_DANRA_conditional.py:503-509): InstanceNorm(ConvT out) -> Conv3x3(64 -> 1) + bias, fp32
// NCHW result (eps_hat).  tail_mma_kernel (elementwise.cuh) treats it as a GEMM with M = pixel, K = 9 taps x 64, N = 1 padded to
// 8: seven eighths of the legacy HMMA work is padding and the tile is re-read from shared memory nine times — measured 2.1 TB/s
// at B=256, 128x128 (ncu: legacy tensor pipe 28 %, DRAM 44 %), neither roofline.
//
// Here the taps are the N dimension:  T[p][tap] = sum_c x[p][c] * w'[tap][c]  (M = 128 consecutive pixels, K = 64, N = 9 -> 16:
// four UMMA 128x16x16 per 16 KB of activations, i.e. per TMA load), and the 3x3 stencil becomes a shifted sum of T over the
// pixel's neighbours:  out[y][x] = bias + sum_{taps inside the image} ( T[(y+r-1, x+s-1)][3r+s] - k[tap] ),
// where w' = w * rstd (InstanceNorm folded per sample, fp16 like every weight) and k[tap] = sum_c w*rstd*mean is the mean term,
// dropped together with T for taps outside the image because the zero padding applies to the NORMALISED tensor.  Every
// activation byte crosses shared memory once: the kernel is an HBM stream.
//   CTA = one band of `rows` image rows of one sample (+ one halo tile above and below);  tile = 128 consecutive pixels
//   warp 0: TMA producer (3-stage ring)   warp 1: MMA issuer (two 16-column TMEM accumulators)
//   warps 2-5: T tile -> shared-memory ring (fp32, tap-major, 4 slots) -> stencil sum of the previous tile -> coalesced fp32 stores
#pragma once
#include "common.cuh"
#include "conv.cuh"

namespace b2d {

constexpr int TT_STAGES = 3;
constexpr int TT_SLOTS = 4;
constexpr int TT_TPITCH = 128 + 4;                                  // floats per tap row of a T slot
constexpr int TT_SLOT_FLOATS = 9 * TT_TPITCH;
constexpr int TT_THREADS = 192;
constexpr int TT_SMEM = 1024 + TT_STAGES * CONV_A_BYTES + 2048 + TT_SLOTS * TT_SLOT_FLOATS * 4 + 256;

__global__ void __launch_bounds__(TT_THREADS, 3)
    tail_tc_kernel(const __grid_constant__ CUtensorMap tmA,           // [B*H*W pixels][64] fp16, box 64 x 128, 128B swizzle
                   const float* __restrict__ stats,                    // [B][64] x {mean, rstd}
                   const float* __restrict__ w,                        // K-major [tap * 64 + c] (c_out == 1)
                   const float* __restrict__ bias, float* __restrict__ out,   // [B,1,H,W]
                   int H, int W, int band_tiles) {
    pdl_launch_dependents();
    extern __shared__ uint8_t tt_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tt_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                                              // TT_STAGES x 16 KB
    uint8_t* sW = sA + TT_STAGES * CONV_A_BYTES;                     // 16 rows x 128 B (rows 9..15 zero), 128B swizzle
    float* sT = reinterpret_cast<float*>(sW + 2048);                 // TT_SLOTS x [9][TT_TPITCH]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sT + TT_SLOTS * TT_SLOT_FLOATS);
    uint64_t* a_full = bars;                 // [TT_STAGES]
    uint64_t* a_empty = bars + TT_STAGES;    // [TT_STAGES]
    uint64_t* t_full = bars + 2 * TT_STAGES; // [2]
    uint64_t* t_empty = t_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
    __shared__ float s_k[9];
    __shared__ float s_st[128];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int tiles_per_img = (H * W) >> 7;
    const int t0 = blockIdx.x * band_tiles;                          // first output tile of this band
    const int nt = min(band_tiles, tiles_per_img - t0);
    const int lo = t0 > 0 ? t0 - 1 : t0;                             // loaded tiles: [lo, hi)
    const int hi = min(t0 + nt + 1, tiles_per_img);
    const int nload = hi - lo;

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmA);
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < TT_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 32);
        tmem_relinquish();
    }
    pdl_wait();
    if (threadIdx.x < 128) s_st[threadIdx.x] = stats[(size_t)b * 128 + threadIdx.x];
    __syncthreads();
    // folded weights (rows = taps, K-major, 128B swizzle: 16-byte chunk j of row n sits at chunk j ^ (n & 7)); rows 9..15 zero
    for (int i = threadIdx.x; i < 16 * 8; i += TT_THREADS) {
        const int n = i >> 3, j = i & 7;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (n < 9) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __ldg(w + n * 64 + j * 8 + e) * s_st[2 * (j * 8 + e) + 1];
            v.x = pack_h2(f[0], f[1]); v.y = pack_h2(f[2], f[3]); v.z = pack_h2(f[4], f[5]); v.w = pack_h2(f[6], f[7]);
        }
        *reinterpret_cast<uint4*>(sW + n * 128 + ((j ^ (n & 7)) << 4)) = v;
    }
    for (int tap = warp; tap < 9; tap += TT_THREADS / 32) {           // mean term per tap (fp32 weights, as tail_mma_kernel)
        const float* wr = w + tap * 64;
        float m = __ldg(wr + lane) * s_st[2 * lane + 1] * s_st[2 * lane] +
                  __ldg(wr + lane + 32) * s_st[2 * (lane + 32) + 1] * s_st[2 * (lane + 32)];
        m = warp_sum(m);
        if (lane == 0) s_k[tap] = m;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        const int pix0 = b * H * W;
        for (int i = 0; i < nload; ++i) {
            const int st = i % TT_STAGES;
            mbar_wait(&a_empty[st], ((i / TT_STAGES) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&a_full[st], CONV_A_BYTES);
                tma_load_2d(sA + st * CONV_A_BYTES, &tmA, &a_full[st], 0, pix0 + (lo + i) * 128);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: T tile = A tile x W'^T =====================
        constexpr uint32_t idesc = umma_idesc_f16(128, 16);
        const uint64_t db = umma_desc_sw128(smem_u32(sW));
        for (int i = 0; i < nload; ++i) {
            const int st = i % TT_STAGES, acc = i & 1;
            mbar_wait(&t_empty[acc], ((i >> 1) & 1) ^ 1);
            mbar_wait(&a_full[st], (i / TT_STAGES) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t da = umma_desc_sw128(smem_u32(sA + st * CONV_A_BYTES));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(tmem + acc * 16, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, k != 0);
                umma_commit(&a_empty[st]);
                umma_commit(&t_full[acc]);
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue: thread = pixel of the tile =====================
        const int q4 = warp & 3;                                     // TMEM lane quarter of this warp
        const int px = q4 * 32 + lane;
        const float bias0 = __ldg(bias);
        float kk[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) kk[t] = s_k[t];
        const int wmask = W - 1, wshift = 31 - __clz(W);             // W is a power of two <= 128
        for (int i = 0; i < nload; ++i) {
            const int acc = i & 1;
            mbar_wait(&t_full[acc], (i >> 1) & 1);
            tc_fence_after();
            uint32_t v[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(acc * 16))
                : "memory");
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[acc]);
            float* slot = sT + (i % TT_SLOTS) * TT_SLOT_FLOATS;
#pragma unroll
            for (int t = 0; t < 9; ++t) slot[t * TT_TPITCH + px] = __uint_as_float(v[t]) - kk[t];
            named_bar_sync(1, 128);                                  // tile i of T is complete in shared memory
            // stencil sum for the tile loaded one step earlier (its lower neighbour has just arrived), and for the last tile of
            // the image when it is the last one loaded
            for (int o = i - 1; o <= i; ++o) {
                const int tile = lo + o;                             // output tile (image-local index)
                if (o < 0 || tile < t0 || tile >= t0 + nt) continue;
                if (o == i && tile + 1 < tiles_per_img) continue;    // its lower neighbour is still to come
                const int q = tile * 128 + px;                       // pixel index inside the image
                const int y = q >> wshift, x = q & wmask;
                float acc_o = bias0;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int yy = y + r - 1;
                    if (yy < 0 || yy >= H) continue;
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        const int xx = x + s - 1;
                        if (xx < 0 || xx >= W) continue;
                        const int qs = (yy << wshift) + xx;          // source pixel
                        const int so = (qs >> 7) - lo;               // load index of its tile (within [o-1, o+1])
                        acc_o += sT[(so % TT_SLOTS) * TT_SLOT_FLOATS + (r * 3 + s) * TT_TPITCH + (qs & 127)];
                    }
                }
                out[(size_t)b * H * W + q] = acc_o;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 32);
    }
}

struct TailTcPlan {
    CUtensorMap tmA;
    int band_tiles = 0;
    dim3 grid;
};

inline bool tail_tc_supported(int H, int W, int c_out) {
    return c_out == 1 && is_pow2(W) && W <= 128 && W >= 4 && (H * W) % 128 == 0;
}

inline int tail_tc_plan_build(TailTcPlan& pl, const f16* x, int B, int H, int W) {
    uint64_t dims[2] = {64, (uint64_t)B * H * W};
    uint64_t str[1] = {128};
    uint32_t box[2] = {64, 128};
    B2D_TRY(make_tmap_f16(&pl.tmA, x, 2, dims, str, box));
    const int tiles = H * W / 128;
    pl.band_tiles = tiles < 16 ? tiles : 16;
    pl.grid = dim3((tiles + pl.band_tiles - 1) / pl.band_tiles, B, 1);
    return 0;
}

inline int tail_tc_init_attrs() {
    B2D_CUDA(cudaFuncSetAttribute(tail_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TT_SMEM));
    B2D_CUDA(cudaFuncSetAttribute(tail_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return 0;
}

inline int tail_tc_launch(const TailTcPlan& pl, const float* stats, const float* w, const float* bias, float* out, int H, int W,
                          cudaStream_t st) {
    B2D_CUDA(launch_k(tail_tc_kernel, pl.grid, dim3(TT_THREADS), TT_SMEM, st, pl.tmA, stats, w, bias, out, H, W, pl.band_tiles));
    return 0;
}

}  // namespace b2d
