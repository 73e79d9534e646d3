// tcgen05 flash attention for head_dim 32 (C = 128 with 4 heads: UNet_downscale's sa1 / sa4 at L = 1024 / 256, Family R's 128-channel
// level at 128x128) — the occupancy design of attention_tc3.cuh (one 128-query tile per CTA, optimistic lazy maximum, 32 scores
// in registers at a time, part of the exponentials on the FMA pipe, four CTAs per SM) with the operand shapes of head_dim 32:
//   * Q [128 x 32] and K [64 x 32] tiles are K-major rows of 64 bytes (64-byte swizzle), S = Q K^T takes two K = 16 MMAs;
//   * V [64 keys x 32] is the MN-major B operand of O += P V on its own (N = 32): with a ones tile next to it ([V | ones], N = 48)
//     S 64 + P 32 + O 48 columns would not fit the 128 TMEM columns that four co-resident CTAs leave each other, so the softmax
//     denominator is summed in registers instead — from the SAME fp16-rounded P the MMA consumes (packed HADD2 partial sums
//     per 32-key chunk, accumulated in fp32), and rescaled together with O.
// Replaces flash_attn_kernel<32> (mma.sync) wherever L % 128 == 0.
#pragma once
#include "attention_tc3.cuh"

namespace b2d {

constexpr int AT5_D = 32;
constexpr int AT5_Q_BYTES = ATC_BLK * AT5_D * 2;      // 8 KB
constexpr int AT5_KV_BYTES = ATC_BN * AT5_D * 2;      // 4 KB
constexpr int AT5_STAGES = 4;
constexpr int AT5_P_COL = 64, AT5_O_COL = 96;
constexpr int AT5_SMEM = 1024 + AT5_Q_BYTES + 2 * AT5_STAGES * AT5_KV_BYTES + 256;

__device__ __forceinline__ float h2_sum(uint32_t a, uint32_t b) {      // (a.lo + a.hi) + (b.lo + b.hi), halves -> fp32
    const float2 x = unpack_h2(a), y = unpack_h2(b);
    return (x.x + x.y) + (y.x + y.y);
}

template <int POLY>
__global__ void __launch_bounds__(AT3_THREADS, AT3_CTAS_PER_SM)
    attn_tc5_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmkv, f16* __restrict__ o, int L,
                    int C, float scale_log2e) {
    pdl_launch_dependents();
    extern __shared__ uint8_t at5_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at5_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + AT5_Q_BYTES;
    uint8_t* sV = sK + AT5_STAGES * AT5_KV_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + AT5_STAGES * AT5_KV_BYTES);
    constexpr int KV_FULL = 0, KV_EMPTY = 4, S_FULL = 8, S_EMPTY = 9, P_FULL = 10, P_EMPTY = 11, Q_FULL = 12, O_FULL = 13, NBARS = 14;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
    const uint32_t bar0 = smem_u32(bars);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.y, b = blockIdx.z;
    const int q0 = blockIdx.x * ATC_BLK;
    const int nb = L / ATC_BN;
    const int row_base = b * L;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm); tma_prefetch_desc(&tmkv); }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < AT5_STAGES; ++i) { mbar_init(&bars[KV_FULL + i], 1); mbar_init(&bars[KV_EMPTY + i], 1); }
            mbar_init(&bars[S_FULL], 1); mbar_init(&bars[S_EMPTY], 4);
            mbar_init(&bars[P_FULL], 4); mbar_init(&bars[P_EMPTY], 1);
            mbar_init(&bars[Q_FULL], 1); mbar_init(&bars[O_FULL], 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, AT3_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(&bars[Q_FULL], AT5_Q_BYTES);
            tma_load_2d(sQ, &tm, &bars[Q_FULL], head * AT5_D, row_base + q0);
        }
        __syncwarp();
        for (int t = 0; t < nb; ++t) {
            const int st = t & (AT5_STAGES - 1);
            mbar_wait_a(bar0 + 8 * (KV_EMPTY + st), ((t >> 2) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&bars[KV_FULL + st], 2 * AT5_KV_BYTES);
                tma_load_2d(sK + st * AT5_KV_BYTES, &tmkv, &bars[KV_FULL + st], C + head * AT5_D, row_base + t * ATC_BN);
                tma_load_2d(sV + st * AT5_KV_BYTES, &tmkv, &bars[KV_FULL + st], 2 * C + head * AT5_D, row_base + t * ATC_BN);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc_s = umma_idesc_f16_ex(128, ATC_BN, 0);
        constexpr uint32_t idesc_o = umma_idesc_f16_ex(128, AT5_D, 1);           // B MN-major, N = 32
        // K-major operands with 64-byte rows, 64-byte swizzle: 8-row atoms 512 B apart (layout code 4)
        const uint64_t dq = umma_desc(smem_u32(sQ), 0, 512, 4);
        const uint64_t dk0 = umma_desc(smem_u32(sK), 0, 512, 4);
        // V [key][32] as MN-major B: one 32-wide MN atom (64 B), 8-key K atoms 512 B apart; 16 keys per MMA = 1024 B
        const uint64_t dv0 = umma_desc(smem_u32(sV), AT5_KV_BYTES, 512, 4);
        mbar_wait_a(bar0 + 8 * Q_FULL, 0);
        auto issue_s = [&](int j) {
            const int ks = j & (AT5_STAGES - 1);
            mbar_wait_a(bar0 + 8 * (KV_FULL + ks), (j >> 2) & 1);
            mbar_wait_a(bar0 + 8 * S_EMPTY, (j & 1) ^ 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dk = dk0 + (uint64_t)(ks * (AT5_KV_BYTES / 16));
                umma_f16(tmem, dq, dk, idesc_s, 0);
                umma_f16(tmem, dq + 2, dk + 2, idesc_s, 1);                       // second 16 of the head dimension: +32 B
                umma_commit_a(bar0 + 8 * S_FULL);
            }
            __syncwarp();
        };
        auto issue_pv = [&](int j) {
            const int vs = j & (AT5_STAGES - 1);
            mbar_wait_a(bar0 + 8 * P_FULL, j & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dv = dv0 + (uint64_t)(vs * (AT5_KV_BYTES / 16));
#pragma unroll
                for (int kk = 0; kk < ATC_BN / 16; ++kk)
                    umma_f16_ts(tmem + AT5_O_COL, tmem + AT5_P_COL + kk * 8, dv + (uint64_t)(kk * (1024 / 16)), idesc_o, (j | kk) != 0);
                umma_commit_a(bar0 + 8 * P_EMPTY);
                umma_commit_a(bar0 + 8 * (KV_EMPTY + vs));
                if (j == nb - 1) umma_commit_a(bar0 + 8 * O_FULL);
            }
            __syncwarp();
        };
        for (int j = 0; j < nb; ++j) {
            issue_s(j);
            if (j > 0) issue_pv(j - 1);
        }
        issue_pv(nb - 1);
    } else {
        const int q = warp & 3;
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
        const int row = q * 32 + lane;
        float m_ref = 0.f, m_pending = 0.f, l_sum = 0.f;
        bool pending = false;
        auto rescale_o = [&](float fac) {        // O (32 columns) *= fac in TMEM; the register denominator with it
            uint32_t ov[32];
            tmem_ld32(tl + AT5_O_COL, ov);
            tmem_ld_wait();
            uint32_t o0[16], o1[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                o0[i] = __float_as_uint(__uint_as_float(ov[i]) * fac);
                o1[i] = __float_as_uint(__uint_as_float(ov[16 + i]) * fac);
            }
            tmem_st16(tl + AT5_O_COL, o0);
            tmem_st16(tl + AT5_O_COL + 16, o1);
            tmem_st_wait();
            l_sum *= fac;
        };
        for (int j = 0; j < nb; ++j) {
            mbar_wait_a(bar0 + 8 * S_FULL, j & 1);
            tc_fence_after();
            if (j == 0) {
                float mx = -INFINITY;
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    uint32_t v[32];
                    tmem_ld32(tl + ch * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
                }
                m_ref = mx * scale_log2e;
            }
            if (__any_sync(0xffffffffu, pending)) {
                mbar_wait_a(bar0 + 8 * P_EMPTY, (j - 1) & 1);
                tc_fence_after();
                rescale_o(pending ? ex2_approx(m_ref - m_pending) : 1.0f);
                if (pending) m_ref = m_pending;
                pending = false;
            }
            float bmx, blk_sum;
            bool waited_p = false;
            for (int pass = 0; pass < 2; ++pass) {
                const float neg_m = -m_ref, neg_m15 = 15.0f - m_ref;
                bmx = -INFINITY;
                blk_sum = 0.f;
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    uint32_t v[32];
                    tmem_ld32(tl + ch * 32, v);
                    tmem_ld_wait();
                    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
                    for (int i = 0; i < 32; i += 2)
                        mx[(i >> 1) & 1] = fmaxf(mx[(i >> 1) & 1], fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
                    bmx = fmaxf(bmx, fmaxf(mx[0], mx[1]));
                    uint32_t pk[16];
#pragma unroll
                    for (int idx = 0; idx < 16 + AT_PIPE; ++idx) {
                        if (idx < 16) {
                            const float s0 = __uint_as_float(v[2 * idx]), s1 = __uint_as_float(v[2 * idx + 1]);
                            if ((idx & 7) < POLY) {
                                pk[idx] = ex2_pair_poly(fmaf(s0, scale_log2e, neg_m15), fmaf(s1, scale_log2e, neg_m15));
                            } else {
                                v[2 * idx] = __float_as_uint(ex2_ordered(fmaf(s0, scale_log2e, neg_m)));
                                v[2 * idx + 1] = __float_as_uint(ex2_ordered(fmaf(s1, scale_log2e, neg_m)));
                            }
                        }
                        if (idx >= AT_PIPE) {
                            const int i = idx - AT_PIPE;
                            if ((i & 7) >= POLY) pk[i] = pack_h2_ordered(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                        }
                    }
                    // denominator of this chunk from the rounded P: four packed-half partial sums of 4 pairs each (<= 8 x 2^15
                    // would overflow fp16, so the packed sums stay short), then fp32
                    uint32_t a0 = pk[0], a1 = pk[4], a2 = pk[8], a3 = pk[12];
#pragma unroll
                    for (int i = 1; i < 4; ++i) {
                        asm("add.rn.f16x2 %0, %0, %1;" : "+r"(a0) : "r"(pk[i]));
                        asm("add.rn.f16x2 %0, %0, %1;" : "+r"(a1) : "r"(pk[4 + i]));
                        asm("add.rn.f16x2 %0, %0, %1;" : "+r"(a2) : "r"(pk[8 + i]));
                        asm("add.rn.f16x2 %0, %0, %1;" : "+r"(a3) : "r"(pk[12 + i]));
                    }
                    blk_sum += h2_sum(a0, a1) + h2_sum(a2, a3);
                    if (!waited_p) {
                        mbar_wait_a(bar0 + 8 * P_EMPTY, (j & 1) ^ 1);
                        tc_fence_after();
                        waited_p = true;
                    }
                    tmem_st16(tl + AT5_P_COL + ch * 16, pk);
                }
                const float bm = bmx * scale_log2e;
                const bool overflow = bm > m_ref + 13.0f;            // packed partial sums of four P must stay finite: 4 x 2^13 < 65504
                if (!__any_sync(0xffffffffu, overflow)) {
                    if (bm > m_ref + 8.0f) {
                        pending = true;
                        m_pending = bm;
                    }
                    break;
                }
                tmem_st_wait();
                rescale_o(overflow ? ex2_approx(m_ref - bm) : 1.0f);
                if (overflow) m_ref = bm;
            }
            l_sum += blk_sum;
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (elect_one()) {
                mbar_arrive_a(bar0 + 8 * S_EMPTY);
                mbar_arrive_a(bar0 + 8 * P_FULL);
            }
            __syncwarp();
        }
        // ---- epilogue: O / l -> fp16 (64 bytes per row)
        mbar_wait_a(bar0 + 8 * O_FULL, 0);
        tc_fence_after();
        uint32_t ov[32];
        tmem_ld32(tl + AT5_O_COL, ov);
        tmem_ld_wait();
        const float inv = 1.0f / l_sum;
        f16* op = o + ((size_t)(row_base + q0 + row)) * C + head * AT5_D;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4 ov4;
            ov4.x = pack_h2(__uint_as_float(ov[8 * j + 0]) * inv, __uint_as_float(ov[8 * j + 1]) * inv);
            ov4.y = pack_h2(__uint_as_float(ov[8 * j + 2]) * inv, __uint_as_float(ov[8 * j + 3]) * inv);
            ov4.z = pack_h2(__uint_as_float(ov[8 * j + 4]) * inv, __uint_as_float(ov[8 * j + 5]) * inv);
            ov4.w = pack_h2(__uint_as_float(ov[8 * j + 6]) * inv, __uint_as_float(ov[8 * j + 7]) * inv);
            reinterpret_cast<uint4*>(op)[j] = ov4;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, AT3_TMEM_COLS);
    }
}

inline bool attn_tc5_supported(int L, int C, int heads) {
    static const bool off = getenv("B2D_NO_TC_ATTN32") != nullptr;
    return !off && heads > 0 && C % heads == 0 && C / heads == AT5_D && L % ATC_BLK == 0;
}

template <int POLY>
inline int attn_tc5_attr() {
    B2D_CUDA(cudaFuncSetAttribute(attn_tc5_kernel<POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT5_SMEM));
    B2D_CUDA(cudaFuncSetAttribute(attn_tc5_kernel<POLY>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return 0;
}
inline int attn_tc5_init_attrs() {
    B2D_TRY(attn_tc5_attr<0>());
    B2D_TRY(attn_tc5_attr<2>());
    B2D_TRY(attn_tc5_attr<4>());
    return 0;
}

// Tensor maps over the qkv buffer [B*L][3C] fp16: boxes of 32 channels x 128 (queries) / 64 (keys) tokens, 64-byte swizzle.
inline int attn_tc5_make_map(AttnTcMaps* m, const f16* qkv, int B, int L, int C) {
    uint64_t dims[2] = {(uint64_t)3 * C, (uint64_t)B * L};
    uint64_t str[1] = {(uint64_t)3 * C * 2};
    uint32_t boxq[2] = {AT5_D, ATC_BLK}, boxk[2] = {AT5_D, ATC_BN};
    B2D_TRY(make_tmap_f16(&m->q, qkv, 2, dims, str, boxq, CU_TENSOR_MAP_SWIZZLE_64B));
    return make_tmap_f16(&m->kv, qkv, 2, dims, str, boxk, CU_TENSOR_MAP_SWIZZLE_64B);
}

inline int attn_tc5_launch(const AttnTcMaps& m, f16* o, int B, int L, int C, int heads, cudaStream_t st) {
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)AT5_D);
    const dim3 grid(L / ATC_BLK, heads, B), block(AT3_THREADS);
    const int poly = attn_tc3_poly();
    if (poly <= 0) B2D_CUDA(launch_k(attn_tc5_kernel<0>, grid, block, AT5_SMEM, st, m.q, m.kv, o, L, C, scale_log2e));
    else if (poly <= 2) B2D_CUDA(launch_k(attn_tc5_kernel<2>, grid, block, AT5_SMEM, st, m.q, m.kv, o, L, C, scale_log2e));
    else B2D_CUDA(launch_k(attn_tc5_kernel<4>, grid, block, AT5_SMEM, st, m.q, m.kv, o, L, C, scale_log2e));
    return 0;
}

}  // namespace b2d
