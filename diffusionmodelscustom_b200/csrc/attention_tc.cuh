// tcgen05 flash attention for head_dim 16 (C = 64, 4 heads: the L = 1024 / 4096 layers that dominate the step).
//
// One CTA = 128 queries of one (sample, head).  S = Q K^T and O += P V run on the 5th-gen tensor cores with accumulators in
// TMEM; each softmax thread owns one full query row (thread <-> TMEM lane), so row max / exp need no shuffles, and P goes
// back to TMEM with tcgen05.st and is consumed as the A operand of the P.V MMA (never touches shared memory).
// S is read from TMEM exactly once per block (measured: the TMEM->register path, ~64 B/clk/SM, is what bounds this
// kernel together with the SFU): online softmax with a LAZY reference maximum — the row's reference m only moves when the
// block maximum exceeds it by more than 8 (in log2 units), so P <= 2^8 stays well inside fp16 and the O accumulator in
// TMEM is rescaled (tcgen05.ld / scale / tcgen05.st by the warp that owns those lanes) only on those rare blocks.
// The softmax denominator comes from the same MMA as O: the V tile and a constant "ones" tile form one MN-major B operand
// with N = 32, so the denominator is accumulated in fp32 from exactly the fp16-rounded P the numerator uses.
//   warp 0: TMA producer (Q once; K and V tiles through 4-stage rings, 32B-swizzled 128 x 16 tiles)
//   warp 1: MMA issuer + TMEM owner       warps 2-5: softmax / correction / epilogue (128 threads)
// TMEM (256 columns per CTA, two CTAs per SM so that one CTA's load/max/store phases overlap the other's exponentials):
//   S [0,128) | P [128,192) | O [192,208) | denominators [208,224)
#pragma once
#include "common.cuh"
#include "conv.cuh"

namespace b2d {

constexpr int ATC_BLK = 128;   // queries per CTA (UMMA M)
constexpr int ATC_BN = 64;     // keys per block (UMMA N of S, K extent of P.V)
constexpr int ATC_CTAS_PER_SM = 3;
constexpr int ATC_POLY = 0;    // of every 8 fp16 pairs of P, this many are exponentiated on the FMA pipe (rest on the SFU)

// 2^x on the FMA/ALU pipes (the SFU does 16 ex2/clk/SM and is what bounds this kernel): round-to-nearest split
// x = n + f, f in [-0.5, 0.5] via the 1.5*2^23 magic constant, degree-3 minimax polynomial for 2^f (max rel. error 7.7e-5,
// below the 4.9e-4 resolution of the fp16 P it feeds), exponent patched in with an integer add.
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -125.0f);
    const float xr = x + 12582912.0f;
    const float f = x - (xr - 12582912.0f);
    float p = fmaf(0.055088773f, f, 0.24260406f);
    p = fmaf(p, f, 0.69327623f);
    p = fmaf(p, f, 0.99992895f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}
constexpr int ATC_D = 16;
constexpr int ATC_STAGES = 4;
constexpr int ATC_TILE_BYTES = ATC_BLK * ATC_D * 2;   // Q tile, 4 KB
constexpr int ATC_KV_BYTES = ATC_BN * ATC_D * 2;      // K / V / ones tile
constexpr int ATC_TMEM_COLS = (ATC_BN + ATC_BN / 2 + 32) <= 128 ? 128 : 256;
constexpr int ATC_P_COL = ATC_BN, ATC_O_COL = ATC_BN + ATC_BN / 2;
constexpr int ATC_SMEM = 1024 + ATC_TILE_BYTES + ATC_KV_BYTES * 3 * ATC_STAGES + 256 + 512;

constexpr int ATC_THREADS = 192;
__global__ void __launch_bounds__(ATC_THREADS, ATC_CTAS_PER_SM) attn_tc_kernel(const __grid_constant__ CUtensorMap tm,
                                                                                const __grid_constant__ CUtensorMap tmkv, f16* __restrict__ o, int L, int C,
                                                      float scale_log2e) {
    pdl_launch_dependents();
    extern __shared__ uint8_t atc_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(atc_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + ATC_TILE_BYTES;
    // V stages are 8 KB: [V tile 4 KB][ones tile 4 KB].  The pair is ONE MN-major B operand with N = 32: two 16-wide
    // MN-atoms LBO = 4096 B apart, so a single MMA per 16 keys yields O (16 columns) and the softmax denominator
    // (ones rows are [1,0..0 | 1,0..0], invariant under the 32-byte swizzle).
    uint8_t* sV = sK + ATC_STAGES * ATC_KV_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATC_STAGES * 2 * ATC_KV_BYTES);
    uint64_t* k_full = bars;             // [4]
    uint64_t* k_empty = bars + 4;        // [4]
    uint64_t* v_full = bars + 8;         // [4]
    uint64_t* v_empty = bars + 12;       // [4]
    uint64_t* s_full = bars + 16;        // [2]
    uint64_t* s_empty = bars + 18;       // [2]
    uint64_t* p_full = bars + 20;        // [2]
    uint64_t* p_empty = bars + 22;       // [2]
    uint64_t* q_full = bars + 24;
    uint64_t* o_full = bars + 25;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);
    float* s_max = reinterpret_cast<float*>(bars + 28);   // [4][128] pass-1 partial maxima

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.y, b = blockIdx.z;
    const int q0 = blockIdx.x * ATC_BLK;
    const int nb = L / ATC_BN;
    const int row_base = b * L;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm); tma_prefetch_desc(&tmkv); }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < ATC_STAGES; ++i) {
                mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
                mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4);    // one elected arrival per softmax warp
                mbar_init(&p_full[i], 4); mbar_init(&p_empty[i], 1);
            }
            mbar_init(q_full, 1);
            mbar_init(o_full, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, ATC_TMEM_COLS);
        tmem_relinquish();
    }
    if (warp >= 2) {   // constant ones tiles (generic-proxy writes -> visible to the async proxy after the fence)
        for (int i = threadIdx.x - 64; i < ATC_STAGES * ATC_BN * 2; i += 128) {
            const int st = i / (ATC_BN * 2), r = i % (ATC_BN * 2);        // 16-byte chunk r of stage st's ones tile
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
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, ATC_TILE_BYTES);
            tma_load_2d(sQ, &tm, q_full, head * ATC_D, row_base + q0);
            for (int t = 0; t < nb; ++t) {
                const int st = t % ATC_STAGES;
                mbar_wait(&k_empty[st], ((t / ATC_STAGES) & 1) ^ 1);
                mbar_arrive_expect_tx(&k_full[st], ATC_KV_BYTES);
                tma_load_2d(sK + st * ATC_KV_BYTES, &tmkv, &k_full[st], C + head * ATC_D, row_base + t * ATC_BN);
                mbar_wait(&v_empty[st], ((t / ATC_STAGES) & 1) ^ 1);
                mbar_arrive_expect_tx(&v_full[st], ATC_KV_BYTES);
                tma_load_2d(sV + st * 2 * ATC_KV_BYTES, &tmkv, &v_full[st], 2 * C + head * ATC_D, row_base + t * ATC_BN);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = umma_idesc_f16_ex(128, ATC_BN, 0);    // S: A,B K-major, N = keys per block
            constexpr uint32_t idesc_o = umma_idesc_f16_ex(128, 2 * ATC_D, 1); // [O | denominators]: A from TMEM, B MN-major, N = 32
            const uint64_t dq = umma_desc(smem_u32(sQ), 0, 256, 6);            // K-major, 32B swizzle: 8-row atoms 256 B apart
            mbar_wait(q_full, 0);
            tc_fence_after();
            auto issue_s = [&](int g) {
                const int ks = g % ATC_STAGES;
                mbar_wait(&k_full[ks], (g / ATC_STAGES) & 1);
                mbar_wait(&s_empty[0], (g & 1) ^ 1);              // softmax has pulled S_{g-1} into registers
                tc_fence_after();
                umma_f16(tmem, dq, umma_desc(smem_u32(sK + ks * ATC_KV_BYTES), 0, 256, 6), idesc_s, 0);
                umma_commit(&s_full[0]);
                umma_commit(&k_empty[ks]);
            };
            auto issue_pv = [&](int j) {
                const int vs = j % ATC_STAGES;
                mbar_wait(&p_full[0], j & 1);
                mbar_wait(&v_full[vs], (j / ATC_STAGES) & 1);
                tc_fence_after();
                const uint32_t vaddr = smem_u32(sV + vs * 2 * ATC_KV_BYTES);
#pragma unroll
                for (int kk = 0; kk < ATC_BN / 16; ++kk) {
                    const uint32_t pa = tmem + ATC_P_COL + kk * 8;           // 16 fp16 of K = 8 packed columns
                    // V tile [key][16] is MN-major for B: K-atoms (8 keys x 32 B) 256 B apart; 16 keys per MMA = 512 B
                    umma_f16_ts(tmem + ATC_O_COL, pa, umma_desc(vaddr + kk * 512, ATC_KV_BYTES, 256, 6), idesc_o, (j | kk) != 0);
                }
                umma_commit(&p_empty[0]);
                umma_commit(&v_empty[vs]);
            };
            for (int j = 0; j < nb; ++j) {                            // S_j first, then P.V of the previous block
                issue_s(j);
                if (j > 0) issue_pv(j - 1);
            }
            issue_pv(nb - 1);
            umma_commit(o_full);
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int row = q * 32 + lane;
        float m_ref = -INFINITY;                                 // reference maximum of this row, in scaled (log2) units
        for (int j = 0; j < nb; ++j) {
            mbar_wait(&s_full[0], j & 1);
            tc_fence_after();
            constexpr int NCH = ATC_BN / 32;
            uint32_t v[NCH][32];
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) tmem_ld32(tmem + lane_off + ch * 32, v[ch]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[0]);            // scores are in registers: the S buffer can be refilled
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
                for (int i = 0; i < 32; ++i) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(v[ch][i]));
            const float bm = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * scale_log2e;
            // lazy reference update: move m_ref only if this block exceeds it by more than 2^8
            const bool move = bm > m_ref + 8.0f;
            const float m_new = move ? bm : m_ref;
            const bool need_fix = move && (j > 0);              // O already holds contributions relative to the old m_ref
            if (__any_sync(0xffffffffu, need_fix)) {
                // P.V of block j-1 must have retired before O is touched; P.V of block j waits for our p_full arrival
                mbar_wait(&p_empty[0], (j - 1) & 1);
                tc_fence_after();
                uint32_t ov[32];
                tmem_ld32(tmem + lane_off + ATC_O_COL, ov);
                tmem_ld_wait();
                const float fac = need_fix ? ex2_approx(m_ref - m_new) : 1.0f;
                uint32_t o0[16], o1[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    o0[i] = __float_as_uint(__uint_as_float(ov[i]) * fac);
                    o1[i] = __float_as_uint(__uint_as_float(ov[16 + i]) * fac);
                }
                tmem_st16(tmem + lane_off + ATC_O_COL, o0);
                tmem_st16(tmem + lane_off + ATC_O_COL + 16, o1);
                tmem_st_wait();
            }
            m_ref = m_new;
            // exponentials first (the long phase), the P buffer is only needed when they are done
            uint32_t pk[NCH][16];
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float x0 = fmaf(__uint_as_float(v[ch][2 * i]), scale_log2e, -m_ref);
                    const float x1 = fmaf(__uint_as_float(v[ch][2 * i + 1]), scale_log2e, -m_ref);
                    pk[ch][i] = ((i & 7) < ATC_POLY) ? pack_h2_nosat(ex2_poly(x0), ex2_poly(x1))
                                                     : pack_h2_nosat(ex2_approx(x0), ex2_approx(x1));
                }
            }
            mbar_wait(&p_empty[0], (j & 1) ^ 1);                // P.V of block j-1 has finished reading the P buffer
            tc_fence_after();
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) tmem_st16(tmem + lane_off + ATC_P_COL + ch * 16, pk[ch]);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[0]);
        }
        // ---- epilogue: O / l -> fp16
        mbar_wait(o_full, 0);
        tc_fence_after();
        uint32_t ov[32];
        tmem_ld32(tmem + lane_off + ATC_O_COL, ov);
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
        tmem_dealloc(tmem, ATC_TMEM_COLS);
    }
}

inline bool attn_tc_supported(int L, int C, int heads) { return C / heads == ATC_D && C % heads == 0 && L % ATC_BLK == 0; }

inline int attn_tc_init_attrs() {
    B2D_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM));
    return 0;
}

// Tensor map over the qkv buffer [B*L][3C] fp16: box = 16 channels x 128 tokens, 32-byte swizzle.
struct AttnTcMaps {
    CUtensorMap q, kv;   // same tensor, boxes of 128 (queries) / ATC_BN (keys) tokens
};
inline int attn_tc_make_map(AttnTcMaps* m, const f16* qkv, int B, int L, int C) {
    uint64_t dims[2] = {(uint64_t)3 * C, (uint64_t)B * L};
    uint64_t str[1] = {(uint64_t)3 * C * 2};
    uint32_t boxq[2] = {ATC_D, ATC_BLK}, boxk[2] = {ATC_D, ATC_BN};
    B2D_TRY(make_tmap_f16(&m->q, qkv, 2, dims, str, boxq, CU_TENSOR_MAP_SWIZZLE_32B));
    return make_tmap_f16(&m->kv, qkv, 2, dims, str, boxk, CU_TENSOR_MAP_SWIZZLE_32B);
}

inline int attn_tc_launch(const AttnTcMaps& m, f16* o, int B, int L, int C, int heads, cudaStream_t st) {
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)ATC_D);
    B2D_CUDA(launch_k(attn_tc_kernel, dim3(L / ATC_BLK, heads, B), dim3(ATC_THREADS), ATC_SMEM, st, m.q, m.kv, o, L, C,
                      scale_log2e));
    return 0;
}

}  // namespace b2d
