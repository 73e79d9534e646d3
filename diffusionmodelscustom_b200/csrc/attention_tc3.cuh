// tcgen05 flash attention for head_dim 16, occupancy variant: the structure of attn_tc_kernel (one 128-query tile per CTA,
// S / P / O in 128 TMEM columns) with the softmax reorganised so that a CTA needs few enough registers for FOUR co-resident
// CTAs per SM (4 x 128 TMEM columns = all of TMEM, 16 softmax warps = 4 per scheduler).
//
// Why occupancy: a softmax warp's path through one 64-key block is ~1800 clk, of which only 512 are its own SFU time
// (64 MUFU.EX2 x 8 clk); the rest are latencies of a chain of special operations (mbarrier try_wait, tcgen05.ld/st and their
// waits, fences, the hand-off through the single-thread MMA issuer) that no amount of scheduling inside ONE in-order warp
// hides (measured with the clock64 trace of attention_tc2.cuh: two warps per scheduler reach 57 % of the SFU rate, three in
// attn_tc_kernel 70 %).  Independent CTAs de-phase naturally, so four warps per scheduler keep the SFU busy.
//
// What makes it fit in <= 80 registers: the 64 scores of a row are never all in registers.  The row maximum of a block is
// NOT needed before its exponentials: they are taken against the reference maximum carried over from earlier blocks
// ("optimistic" lazy maximum) while the block maximum is accumulated on the side, 32 columns at a time:
//   * growth <= 2^8 over the reference (the normal case): nothing to do, P <= 2^8 sits well inside fp16;
//   * growth in (2^8, 2^15]: P is still finite, the reference moves for the FOLLOWING blocks and O is rescaled once P.V of this
//     block has retired;
//   * growth > 2^15 (would overflow fp16; only while the running maximum is still being found, i.e. the first blocks of peaky
//     rows): O is rescaled, the reference moves and the block is exponentiated again from S, which is still in TMEM.
// Block 0 has no reference yet and reads its maximum in a pre-pass (S_0 is read twice).
// POLY of every 8 fp16 pairs are exponentiated on the FMA pipe in packed half precision (ex2_pair_poly, attention_tc2.cuh).
//   warp 0: TMA producer   warp 1: MMA issuer (converged loop, elect.sync)   warps 2-5: softmax (thread = query row)
#pragma once
#include "attention_tc2.cuh"

namespace b2d {

constexpr int AT3_THREADS = 192;
constexpr int AT3_CTAS_PER_SM = 4;
constexpr int AT3_TMEM_COLS = 128;
constexpr int AT3_P_COL = ATC_BN, AT3_O_COL = ATC_BN + ATC_BN / 2;
constexpr int AT3_SMEM = 1024 + ATC_TILE_BYTES + ATC_KV_BYTES * 3 * ATC_STAGES + 256;

template <int POLY>
__global__ void __launch_bounds__(AT3_THREADS, AT3_CTAS_PER_SM)
    attn_tc3_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmkv, f16* __restrict__ o, int L,
                    int C, float scale_log2e) {
    pdl_launch_dependents();
    extern __shared__ uint8_t at3_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at3_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + ATC_TILE_BYTES;
    uint8_t* sV = sK + ATC_STAGES * ATC_KV_BYTES;                    // [V tile 2 KB | ones tile 2 KB] per stage
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATC_STAGES * 2 * ATC_KV_BYTES);
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
            for (int i = 0; i < ATC_STAGES; ++i) { mbar_init(&bars[KV_FULL + i], 1); mbar_init(&bars[KV_EMPTY + i], 1); }
            mbar_init(&bars[S_FULL], 1); mbar_init(&bars[S_EMPTY], 4);     // one elected arrival per softmax warp
            mbar_init(&bars[P_FULL], 4); mbar_init(&bars[P_EMPTY], 1);
            mbar_init(&bars[Q_FULL], 1); mbar_init(&bars[O_FULL], 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, AT3_TMEM_COLS);
        tmem_relinquish();
    }
    if (warp >= 2) {   // constant ones tiles (generic-proxy writes -> visible to the async proxy after the fence)
        for (int i = threadIdx.x - 64; i < ATC_STAGES * ATC_BN * 2; i += 128) {
            const int st = i / (ATC_BN * 2), r = i % (ATC_BN * 2);
            *reinterpret_cast<uint4*>(sV + st * 2 * ATC_KV_BYTES + ATC_KV_BYTES + r * 16) = make_uint4(0x00003C00u, 0u, 0u, 0u);
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        // ===================== TMA producer (converged loop, one elected thread issues) =====================
        if (elect_one()) {
            mbar_arrive_expect_tx(&bars[Q_FULL], ATC_TILE_BYTES);
            tma_load_2d(sQ, &tm, &bars[Q_FULL], head * ATC_D, row_base + q0);
        }
        __syncwarp();
        for (int t = 0; t < nb; ++t) {
            const int st = t & (ATC_STAGES - 1);
            mbar_wait_a(bar0 + 8 * (KV_EMPTY + st), ((t >> 2) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&bars[KV_FULL + st], 2 * ATC_KV_BYTES);
                tma_load_2d(sK + st * ATC_KV_BYTES, &tmkv, &bars[KV_FULL + st], C + head * ATC_D, row_base + t * ATC_BN);
                tma_load_2d(sV + st * 2 * ATC_KV_BYTES, &tmkv, &bars[KV_FULL + st], 2 * C + head * ATC_D, row_base + t * ATC_BN);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: S_j, then P.V of block j-1 =====================
        constexpr uint32_t idesc_s = umma_idesc_f16_ex(128, ATC_BN, 0);
        constexpr uint32_t idesc_o = umma_idesc_f16_ex(128, 2 * ATC_D, 1);
        const uint64_t dq = umma_desc(smem_u32(sQ), 0, 256, 6);
        const uint64_t dk0 = umma_desc(smem_u32(sK), 0, 256, 6);
        const uint64_t dv0 = umma_desc(smem_u32(sV), ATC_KV_BYTES, 256, 6);
        mbar_wait_a(bar0 + 8 * Q_FULL, 0);
        auto issue_s = [&](int j) {
            const int ks = j & (ATC_STAGES - 1);
            mbar_wait_a(bar0 + 8 * (KV_FULL + ks), (j >> 2) & 1);
            mbar_wait_a(bar0 + 8 * S_EMPTY, (j & 1) ^ 1);                // softmax is done with S_{j-1}
            tc_fence_after();
            if (elect_one()) {
                umma_f16(tmem, dq, dk0 + (uint64_t)(ks * (ATC_KV_BYTES / 16)), idesc_s, 0);
                umma_commit_a(bar0 + 8 * S_FULL);
            }
            __syncwarp();
        };
        auto issue_pv = [&](int j) {
            const int vs = j & (ATC_STAGES - 1);
            mbar_wait_a(bar0 + 8 * P_FULL, j & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dv = dv0 + (uint64_t)(vs * (2 * ATC_KV_BYTES / 16));
#pragma unroll
                for (int kk = 0; kk < ATC_BN / 16; ++kk)
                    umma_f16_ts(tmem + AT3_O_COL, tmem + AT3_P_COL + kk * 8, dv + (uint64_t)(kk * (512 / 16)), idesc_o, (j | kk) != 0);
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
        // ===================== softmax warps =====================
        const int q = warp & 3;
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
        const int row = q * 32 + lane;
        float m_ref = 0.f;                       // reference maximum of this row (log2 units), set by block 0's pre-pass
        float m_pending = 0.f;                   // a moved reference waiting for P.V of the previous block before O is rescaled
        bool pending = false;
        auto rescale_o = [&](float fac) {        // O (16 columns) and the denominator column *= fac, in TMEM
            uint32_t ov[32];
            tmem_ld32(tl + AT3_O_COL, ov);
            tmem_ld_wait();
            uint32_t o0[16], o1[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                o0[i] = __float_as_uint(__uint_as_float(ov[i]) * fac);
                o1[i] = __float_as_uint(__uint_as_float(ov[16 + i]) * fac);
            }
            tmem_st16(tl + AT3_O_COL, o0);
            tmem_st16(tl + AT3_O_COL + 16, o1);
            tmem_st_wait();
        };
        for (int j = 0; j < nb; ++j) {
            mbar_wait_a(bar0 + 8 * S_FULL, j & 1);
            tc_fence_after();
            if (j == 0) {                        // no reference yet: maximum of S_0 first
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
            if (__any_sync(0xffffffffu, pending)) {      // reference moved during block j-1: O *= 2^(old - new) once P.V_{j-1} is in
                mbar_wait_a(bar0 + 8 * P_EMPTY, (j - 1) & 1);
                tc_fence_after();
                rescale_o(pending ? ex2_approx(m_ref - m_pending) : 1.0f);
                if (pending) m_ref = m_pending;
                pending = false;
            }
            float bmx;
            bool waited_p = false;
            for (int pass = 0; pass < 2; ++pass) {       // pass 1 only after an fp16-overflowing growth of the maximum (rare)
                const float neg_m = -m_ref, neg_m15 = 15.0f - m_ref;
                bmx = -INFINITY;
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
                    if (!waited_p) {                     // P.V of block j-1 has finished reading the P buffer
                        mbar_wait_a(bar0 + 8 * P_EMPTY, (j & 1) ^ 1);
                        tc_fence_after();
                        waited_p = true;
                    }
                    tmem_st16(tl + AT3_P_COL + ch * 16, pk);
                }
                const float bm = bmx * scale_log2e;
                const bool overflow = bm > m_ref + 15.0f;            // some P of this row left the fp16 range
                if (!__any_sync(0xffffffffu, overflow)) {
                    if (bm > m_ref + 8.0f) {                         // finite, but move the reference for the blocks to come
                        pending = true;
                        m_pending = bm;
                    }
                    break;
                }
                // redo: P.V_{j-1} has retired (waited above), so O can be rescaled now; then exponentiate S_j again
                tmem_st_wait();
                rescale_o(overflow ? ex2_approx(m_ref - bm) : 1.0f);
                if (overflow) m_ref = bm;
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (elect_one()) {
                mbar_arrive_a(bar0 + 8 * S_EMPTY);                   // S_j is no longer needed
                mbar_arrive_a(bar0 + 8 * P_FULL);
            }
            __syncwarp();
        }
        // ---- epilogue: O / l -> fp16
        mbar_wait_a(bar0 + 8 * O_FULL, 0);
        tc_fence_after();
        uint32_t ov[32];
        tmem_ld32(tl + AT3_O_COL, ov);
        tmem_ld_wait();
        const float inv = 1.0f / __uint_as_float(ov[16]);
        f16* op = o + ((size_t)(row_base + q0 + row)) * C + head * ATC_D;
        uint4 o0, o1;
        o0.x = pack_h2(__uint_as_float(ov[0]) * inv, __uint_as_float(ov[1]) * inv);
        o0.y = pack_h2(__uint_as_float(ov[2]) * inv, __uint_as_float(ov[3]) * inv);
        o0.z = pack_h2(__uint_as_float(ov[4]) * inv, __uint_as_float(ov[5]) * inv);
        o0.w = pack_h2(__uint_as_float(ov[6]) * inv, __uint_as_float(ov[7]) * inv);
        o1.x = pack_h2(__uint_as_float(ov[8]) * inv, __uint_as_float(ov[9]) * inv);
        o1.y = pack_h2(__uint_as_float(ov[10]) * inv, __uint_as_float(ov[11]) * inv);
        o1.z = pack_h2(__uint_as_float(ov[12]) * inv, __uint_as_float(ov[13]) * inv);
        o1.w = pack_h2(__uint_as_float(ov[14]) * inv, __uint_as_float(ov[15]) * inv);
        reinterpret_cast<uint4*>(op)[0] = o0;
        reinterpret_cast<uint4*>(op)[1] = o1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, AT3_TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ 32-key blocks, double-buffered
// Same occupancy design with 32-key blocks: S (2 x 32 columns) and P (2 x 16) are double-buffered inside the same 128 TMEM
// columns, so neither hand-off through the MMA issuer (S_{j+1} after S_j is consumed, storing P_j after P.V_{j-1}) is on a
// softmax warp's path any more; the price is one barrier round per 32 keys instead of 64.
// TMEM: S0 [0,32) | S1 [32,64) | P0 [64,80) | P1 [80,96) | O [96,112) | denominators [112,128)
constexpr int AT4_BN = 32;
constexpr int AT4_KV_BYTES = AT4_BN * ATC_D * 2;     // 1 KB
constexpr int AT4_STAGES = 4;
constexpr int AT4_P_COL = 64, AT4_O_COL = 96;
constexpr int AT4_SMEM = 1024 + ATC_TILE_BYTES + AT4_KV_BYTES * 3 * AT4_STAGES + 256;

template <int POLY>
__global__ void __launch_bounds__(AT3_THREADS, AT3_CTAS_PER_SM)
    attn_tc4_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmkv, f16* __restrict__ o, int L,
                    int C, float scale_log2e) {
    pdl_launch_dependents();
    extern __shared__ uint8_t at4_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at4_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + ATC_TILE_BYTES;
    uint8_t* sV = sK + AT4_STAGES * AT4_KV_BYTES;                    // [V tile 1 KB | ones tile 1 KB] per stage
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + AT4_STAGES * 2 * AT4_KV_BYTES);
    constexpr int KV_FULL = 0, KV_EMPTY = 4, S_FULL = 8, S_EMPTY = 10, P_FULL = 12, P_EMPTY = 14, Q_FULL = 16, O_FULL = 17, NBARS = 18;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
    const uint32_t bar0 = smem_u32(bars);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.y, b = blockIdx.z;
    const int q0 = blockIdx.x * ATC_BLK;
    const int nb = L / AT4_BN;
    const int row_base = b * L;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm); tma_prefetch_desc(&tmkv); }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < AT4_STAGES; ++i) { mbar_init(&bars[KV_FULL + i], 1); mbar_init(&bars[KV_EMPTY + i], 1); }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&bars[S_FULL + i], 1); mbar_init(&bars[S_EMPTY + i], 4);
                mbar_init(&bars[P_FULL + i], 4); mbar_init(&bars[P_EMPTY + i], 1);
            }
            mbar_init(&bars[Q_FULL], 1); mbar_init(&bars[O_FULL], 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, AT3_TMEM_COLS);
        tmem_relinquish();
    }
    if (warp >= 2) {   // constant ones tiles
        for (int i = threadIdx.x - 64; i < AT4_STAGES * AT4_BN * 2; i += 128) {
            const int st = i / (AT4_BN * 2), r = i % (AT4_BN * 2);
            *reinterpret_cast<uint4*>(sV + st * 2 * AT4_KV_BYTES + AT4_KV_BYTES + r * 16) = make_uint4(0x00003C00u, 0u, 0u, 0u);
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(&bars[Q_FULL], ATC_TILE_BYTES);
            tma_load_2d(sQ, &tm, &bars[Q_FULL], head * ATC_D, row_base + q0);
        }
        __syncwarp();
        for (int t = 0; t < nb; ++t) {
            const int st = t & (AT4_STAGES - 1);
            mbar_wait_a(bar0 + 8 * (KV_EMPTY + st), ((t >> 2) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&bars[KV_FULL + st], 2 * AT4_KV_BYTES);
                tma_load_2d(sK + st * AT4_KV_BYTES, &tmkv, &bars[KV_FULL + st], C + head * ATC_D, row_base + t * AT4_BN);
                tma_load_2d(sV + st * 2 * AT4_KV_BYTES, &tmkv, &bars[KV_FULL + st], 2 * C + head * ATC_D, row_base + t * AT4_BN);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc_s = umma_idesc_f16_ex(128, AT4_BN, 0);
        constexpr uint32_t idesc_o = umma_idesc_f16_ex(128, 2 * ATC_D, 1);
        const uint64_t dq = umma_desc(smem_u32(sQ), 0, 256, 6);
        const uint64_t dk0 = umma_desc(smem_u32(sK), 0, 256, 6);
        const uint64_t dv0 = umma_desc(smem_u32(sV), AT4_KV_BYTES, 256, 6);
        mbar_wait_a(bar0 + 8 * Q_FULL, 0);
        auto issue_s = [&](int j) {
            const int ks = j & (AT4_STAGES - 1), sb = j & 1;
            mbar_wait_a(bar0 + 8 * (KV_FULL + ks), (j >> 2) & 1);
            mbar_wait_a(bar0 + 8 * (S_EMPTY + sb), ((j >> 1) & 1) ^ 1);        // softmax is done with S_{j-2}
            tc_fence_after();
            if (elect_one()) {
                umma_f16(tmem + sb * AT4_BN, dq, dk0 + (uint64_t)(ks * (AT4_KV_BYTES / 16)), idesc_s, 0);
                umma_commit_a(bar0 + 8 * (S_FULL + sb));
            }
            __syncwarp();
        };
        auto issue_pv = [&](int j) {
            const int vs = j & (AT4_STAGES - 1), pb = j & 1;
            mbar_wait_a(bar0 + 8 * (P_FULL + pb), (j >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dv = dv0 + (uint64_t)(vs * (2 * AT4_KV_BYTES / 16));
#pragma unroll
                for (int kk = 0; kk < AT4_BN / 16; ++kk)
                    umma_f16_ts(tmem + AT4_O_COL, tmem + AT4_P_COL + pb * 16 + kk * 8, dv + (uint64_t)(kk * (512 / 16)), idesc_o,
                                (j | kk) != 0);
                umma_commit_a(bar0 + 8 * (P_EMPTY + pb));
                umma_commit_a(bar0 + 8 * (KV_EMPTY + vs));
                if (j == nb - 1) umma_commit_a(bar0 + 8 * O_FULL);
            }
            __syncwarp();
        };
        issue_s(0);
        issue_s(1);                                       // nb >= 4
        for (int j = 0; j < nb; ++j) {
            if (j + 2 < nb) issue_s(j + 2);
            issue_pv(j);
        }
    } else {
        const int q = warp & 3;
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
        const int row = q * 32 + lane;
        float m_ref = 0.f;
        auto rescale_o = [&](float fac) {
            uint32_t ov[32];
            tmem_ld32(tl + AT4_O_COL, ov);
            tmem_ld_wait();
            uint32_t o0[16], o1[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                o0[i] = __float_as_uint(__uint_as_float(ov[i]) * fac);
                o1[i] = __float_as_uint(__uint_as_float(ov[16 + i]) * fac);
            }
            tmem_st16(tl + AT4_O_COL, o0);
            tmem_st16(tl + AT4_O_COL + 16, o1);
            tmem_st_wait();
        };
        for (int j = 0; j < nb; ++j) {
            const int sb = j & 1;
            mbar_wait_a(bar0 + 8 * (S_FULL + sb), (j >> 1) & 1);
            tc_fence_after();
            uint32_t v[32];
            tmem_ld32(tl + sb * AT4_BN, v);
            tmem_ld_wait_regs(v);
            float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < 32; i += 2) mx[(i >> 1) & 1] = fmaxf(mx[(i >> 1) & 1], fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
            const float bm = fmaxf(mx[0], mx[1]) * scale_log2e;
            // one chunk IS the block, so its maximum is known before the exponentials: the plain lazy reference of attn_tc_kernel
            // (move only on > 2^8 growth, O rescaled in TMEM by the owning warp once P.V of the previous block has retired)
            const bool move = (j == 0) || bm > m_ref + 8.0f;
            const bool need_fix = move && j > 0;
            if (__any_sync(0xffffffffu, need_fix)) {
                mbar_wait_a(bar0 + 8 * (P_EMPTY + ((j - 1) & 1)), ((j - 1) >> 1) & 1);
                tc_fence_after();
                rescale_o(need_fix ? ex2_approx(m_ref - bm) : 1.0f);
            }
            if (move) m_ref = bm;
            const float neg_m = -m_ref, neg_m15 = 15.0f - m_ref;
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
            mbar_wait_a(bar0 + 8 * (P_EMPTY + sb), ((j >> 1) & 1) ^ 1);        // P.V of block j-2 has finished reading this P buffer
            tc_fence_after();
            tmem_st16(tl + AT4_P_COL + sb * 16, pk);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (elect_one()) {
                mbar_arrive_a(bar0 + 8 * (S_EMPTY + sb));
                mbar_arrive_a(bar0 + 8 * (P_FULL + sb));
            }
            __syncwarp();
        }
        mbar_wait_a(bar0 + 8 * O_FULL, 0);
        tc_fence_after();
        uint32_t ov[32];
        tmem_ld32(tl + AT4_O_COL, ov);
        tmem_ld_wait();
        const float inv = 1.0f / __uint_as_float(ov[16]);
        f16* op = o + ((size_t)(row_base + q0 + row)) * C + head * ATC_D;
        uint4 o0, o1;
        o0.x = pack_h2(__uint_as_float(ov[0]) * inv, __uint_as_float(ov[1]) * inv);
        o0.y = pack_h2(__uint_as_float(ov[2]) * inv, __uint_as_float(ov[3]) * inv);
        o0.z = pack_h2(__uint_as_float(ov[4]) * inv, __uint_as_float(ov[5]) * inv);
        o0.w = pack_h2(__uint_as_float(ov[6]) * inv, __uint_as_float(ov[7]) * inv);
        o1.x = pack_h2(__uint_as_float(ov[8]) * inv, __uint_as_float(ov[9]) * inv);
        o1.y = pack_h2(__uint_as_float(ov[10]) * inv, __uint_as_float(ov[11]) * inv);
        o1.z = pack_h2(__uint_as_float(ov[12]) * inv, __uint_as_float(ov[13]) * inv);
        o1.w = pack_h2(__uint_as_float(ov[14]) * inv, __uint_as_float(ov[15]) * inv);
        reinterpret_cast<uint4*>(op)[0] = o0;
        reinterpret_cast<uint4*>(op)[1] = o1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, AT3_TMEM_COLS);
    }
}

// fraction of fp16 pairs (of every 8) exponentiated on the FMA pipe; B2D_ATTN_POLY overrides (0..4)
inline int attn_tc3_poly() {
    static const int v = [] {
        const char* e = getenv("B2D_ATTN_POLY");
        int p = e ? atoi(e) : 4;
        return p < 0 ? 0 : (p > 8 ? 8 : p);
    }();
    return v;
}

template <int POLY>
inline int attn_tc3_attr() {
    B2D_CUDA(cudaFuncSetAttribute(attn_tc3_kernel<POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT3_SMEM));
    B2D_CUDA(cudaFuncSetAttribute(attn_tc3_kernel<POLY>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return 0;
}
template <int POLY>
inline int attn_tc4_attr() {
    B2D_CUDA(cudaFuncSetAttribute(attn_tc4_kernel<POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT4_SMEM));
    B2D_CUDA(cudaFuncSetAttribute(attn_tc4_kernel<POLY>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return 0;
}
inline int attn_tc3_init_attrs() {
    B2D_TRY(attn_tc4_attr<0>());
    B2D_TRY(attn_tc4_attr<1>());
    B2D_TRY(attn_tc4_attr<2>());
    B2D_TRY(attn_tc4_attr<3>());
    B2D_TRY(attn_tc4_attr<4>());
    B2D_TRY(attn_tc3_attr<0>());
    B2D_TRY(attn_tc3_attr<1>());
    B2D_TRY(attn_tc3_attr<2>());
    B2D_TRY(attn_tc3_attr<3>());
    B2D_TRY(attn_tc3_attr<4>());
    B2D_TRY(attn_tc3_attr<5>());
    B2D_TRY(attn_tc3_attr<6>());
    B2D_TRY(attn_tc3_attr<8>());
    return 0;
}

inline int attn_tc3_launch(const AttnTcMaps& m, f16* o, int B, int L, int C, int heads, cudaStream_t st) {
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)ATC_D);
    const dim3 grid(L / ATC_BLK, heads, B), block(AT3_THREADS);
    switch (attn_tc3_poly()) {
        case 0: B2D_CUDA(launch_k(attn_tc3_kernel<0>, grid, block, AT3_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        case 1: B2D_CUDA(launch_k(attn_tc3_kernel<1>, grid, block, AT3_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        case 2: B2D_CUDA(launch_k(attn_tc3_kernel<2>, grid, block, AT3_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        case 3: B2D_CUDA(launch_k(attn_tc3_kernel<3>, grid, block, AT3_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        case 4: B2D_CUDA(launch_k(attn_tc3_kernel<4>, grid, block, AT3_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        case 5: B2D_CUDA(launch_k(attn_tc3_kernel<5>, grid, block, AT3_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        case 6: case 7: B2D_CUDA(launch_k(attn_tc3_kernel<6>, grid, block, AT3_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        default: B2D_CUDA(launch_k(attn_tc3_kernel<8>, grid, block, AT3_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
    }
    return 0;
}

// 32-key-block variant: its own K/V tensor map (box of 32 keys)
inline int attn_tc4_make_map(AttnTcMaps* m, const f16* qkv, int B, int L, int C) {
    uint64_t dims[2] = {(uint64_t)3 * C, (uint64_t)B * L};
    uint64_t str[1] = {(uint64_t)3 * C * 2};
    uint32_t boxq[2] = {ATC_D, ATC_BLK}, boxk[2] = {ATC_D, AT4_BN};
    B2D_TRY(make_tmap_f16(&m->q, qkv, 2, dims, str, boxq, CU_TENSOR_MAP_SWIZZLE_32B));
    return make_tmap_f16(&m->kv, qkv, 2, dims, str, boxk, CU_TENSOR_MAP_SWIZZLE_32B);
}
inline int attn_tc4_launch(const AttnTcMaps& m, f16* o, int B, int L, int C, int heads, cudaStream_t st) {
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)ATC_D);
    const dim3 grid(L / ATC_BLK, heads, B), block(AT3_THREADS);
    switch (attn_tc3_poly() > 4 ? 4 : attn_tc3_poly()) {
        case 0: B2D_CUDA(launch_k(attn_tc4_kernel<0>, grid, block, AT4_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        case 1: B2D_CUDA(launch_k(attn_tc4_kernel<1>, grid, block, AT4_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        case 2: B2D_CUDA(launch_k(attn_tc4_kernel<2>, grid, block, AT4_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        case 3: B2D_CUDA(launch_k(attn_tc4_kernel<3>, grid, block, AT4_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        default: B2D_CUDA(launch_k(attn_tc4_kernel<4>, grid, block, AT4_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
    }
    return 0;
}

}  // namespace b2d
