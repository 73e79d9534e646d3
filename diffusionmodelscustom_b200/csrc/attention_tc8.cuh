// tcgen05 flash attention for head_dim 16: the issuer-on-named-barriers structure of attention_tc6.cuh (MODE 1) with the softmax
// argument computed BY THE TENSOR CORE.  In attn_tc3 / attn_tc6 every score costs one FFMA (s * scale*log2e - m_ref) before its
// exponential: 64 of the ~380 warp instructions per 64-key block, in a kernel whose issue slots are ~70 % busy.  Here
//   * Q is multiplied by scale*log2e once per CTA, in place in shared memory (fp16, elementwise: layout-agnostic);
//   * every S half gets a second K = 16 MMA step  E_q . E_k^T  accumulated onto  Q_s K^T :  row r of E_q holds
//     [-m_hi, -m_lo, 0.. | -m_hi, -m_lo, 0..] (the row's reference maximum split into two fp16, written twice so that the row is
//     invariant under the 32-byte swizzle), every row of E_k holds [.5, .5, 0.. | .5, .5, 0..]  =>  S'' = S*scale*log2e - m_ref
//     arrives in TMEM ready for MUFU.EX2: an SFU pair costs 2 MUFU + 1 pack (was 5 instructions), a polynomial pair 11 (was 12).
// The reference only moves on > 2^8 growth (lazy maximum at half granularity as in attention_tc6.cuh), so E_q is rewritten
// rarely: by the row's own thread, after the S MMAs that may still be reading it have completed, with a proxy fence before the
// thread's next named-barrier arrival (which is what releases the next S MMA).  A half whose MMA was issued BEFORE a move carries
// the old reference; the thread remembers the value baked into each in-flight half and subtracts the difference in the (rare)
// slow path.  m_ref is kept equal to hi + lo exactly, so what the tensor core subtracts is what the bookkeeping assumes.
// The same MMA step adds a PER-COLUMN constant (E_q column 2 is 1, E_k column 2 holds it): P is stored as p * 2^7 (the factor
// cancels in O / l; 2^8 * 2^7 still fits fp16 and the flush point moves from 2^-15 down to 2^-22), so SFU columns receive x + 7 and
// polynomial columns x + 22, which is what lets the converter's own relu clamp stand in for the lower clamp of the exponent
// range: a polynomial pair costs 10 instructions.
// Tried on top of this and measured slower (kept out): P.V issued per 32-key half with its own barrier (460 vs 456 us at B=32,
// L=4096: the P buffer is not what couples the warps), polynomial and SFU pairs alternating one by one (463 us).
#pragma once
#include "attention_tc6.cuh"

namespace b2d {

constexpr int AT8_THREADS = 160;
constexpr int AT8_EQ_BYTES = ATC_BLK * ATC_D * 2;             // 4 KB
constexpr int AT8_EK_BYTES = (ATC_BN / 2) * ATC_D * 2;        // 1 KB
constexpr int AT8_SMEM = 1024 + ATC_TILE_BYTES + ATC_KV_BYTES * 3 * ATC_STAGES + AT8_EQ_BYTES + AT8_EK_BYTES + 256;

// per-column constants added by the tensor core (log2 units): P is stored as p * 2^7 (the factor cancels in O / l; 2^8 * 2^7 still
// fits fp16), SFU columns receive x + 7, polynomial columns x + 22 (relu clamp = flush below 2^-22).
// Accuracy trade-off, measured with tools/attn_stress.py (200 random cases, rel-L2 of the attention output against fp64):
//   * the polynomial path rounds x + OFF_POLY to fp16: in [16, 32) the ulp is 2^-6, i.e. +-0.54 % (0.31 % rms) on every P of a
//     polynomial column — worst case 2.7e-3 on near-uniform attention (an average of many keys: nothing averages the noise against
//     the signal).  OFF_POLY = 15 / OFF_SFU = 0 halves that for x <= 1 (1.4e-3) but flushes at 2^-15, and a peaked row over a broad
//     floor of 4096 keys then loses up to a few per cent of its mass in the polynomial columns (3.3e-3 in the stress set, the same
//     case where attn_tc3 — which flushed at 2^-15 too — shows 3.5e-3); a clamp by HMNMX2 instead of relu (offset 0: ulp 2^-11 near
//     x = 0) costs one instruction per pair = ~3.5 % of the kernel.
//   * at model level none of this is visible: eps_hat rel-L2 against the reference goldens is 1.02e-3 mean / 1.84e-3 max with this
//     kernel, with attn_tc3 and with the fp32-softmax mma.sync kernel alike (tools/eps_error.py) — the attention output is a small
//     residual contribution next to the fp16 storage of every activation.
constexpr float AT8_OFF_SFU = 7.0f, AT8_OFF_POLY = 22.0f;

// 2^(x' - 15) for a pair of fp32 arguments x' in (-inf, 30] on the FMA pipe in packed half precision; x' <= 0 flushes to +0 exactly
__device__ __forceinline__ uint32_t ex2_pair_poly_off(float xa, float xb) {
    uint32_t h, xr, nf, f, p, r;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(xb), "f"(xa));             // low half <- xa; negative -> +0
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(xr) : "r"(h), "r"(0x66006600u));              // + 1536: integer part lands in the mantissa
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(nf) : "r"(xr), "r"(0x66006600u));             // n' as a half
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(f) : "r"(h), "r"(nf));                        // f in [-0.5, 0.5]
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(0x2B0D2B0Du), "r"(f), "r"(0x33C333C3u));   // 0.05509 f + 0.24260
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(p), "r"(f), "r"(0x398C398Cu));              // .. f + 0.69328
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(p), "r"(f), "r"(0x3C003C00u));              // .. f + 1
    // 2^(n' - 15) as half bits in ONE integer multiply-add: xr << 10 puts n' (mantissa bits 0..4, n' <= 31) into each exponent
    // field; what the low half spills into the high half's mantissa is the constant 0x6600 >> 6, subtracted again.  n' = 0 -> +0.0
    const uint32_t e = xr * 1024u - 0x01980000u;
    asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(p), "r"(e));
    return r;
}

template <int POLY>
__global__ void __launch_bounds__(AT8_THREADS, AT6_CTAS_PER_SM)
    attn_tc8_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmkv, f16* __restrict__ o, int L,
                    int C, float scale_log2e) {
    pdl_launch_dependents();
    extern __shared__ uint8_t at8_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at8_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + ATC_TILE_BYTES;
    uint8_t* sV = sK + ATC_STAGES * ATC_KV_BYTES;                    // [V tile 2 KB | ones tile 2 KB] per stage
    uint8_t* sEq = sV + ATC_STAGES * 2 * ATC_KV_BYTES;               // [128 rows][16] fp16: -m_ref of each row
    uint8_t* sEk = sEq + AT8_EQ_BYTES;                               // [32 rows][16] fp16: constant
    uint64_t* bars = reinterpret_cast<uint64_t*>(sEk + AT8_EK_BYTES);
    constexpr int KV_FULL = 0, S_FULL = 4 /* two halves */, P_EMPTY = 6, Q_FULL = 7, NBARS = 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
    const uint32_t bar0 = smem_u32(bars);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.y, b = blockIdx.z;
    const int q0 = blockIdx.x * ATC_BLK;
    const int nb = L / ATC_BN;
    const int row_base = b * L;
    constexpr int IW = 4, NT = AT8_THREADS;

    if (warp == IW) {
        if (lane == 0) {
            tma_prefetch_desc(&tm);
            tma_prefetch_desc(&tmkv);
            for (int i = 0; i < NBARS; ++i) mbar_init(&bars[i], 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, AT3_TMEM_COLS);
        tmem_relinquish();
    }
    // constant tiles (generic-proxy writes -> visible to the async proxy after the fence): ones next to V, E_k, E_q = 0
    for (int i = threadIdx.x; i < ATC_STAGES * ATC_BN * 2; i += NT) {
        const int st = i / (ATC_BN * 2), r = i % (ATC_BN * 2);
        *reinterpret_cast<uint4*>(sV + st * 2 * ATC_KV_BYTES + ATC_KV_BYTES + r * 16) = make_uint4(0x00003C00u, 0u, 0u, 0u);
    }
    // E_q rows [0, 0, 1, 0.. | same]; E_k row c [.5, .5, off_c / 2, 0.. | same] (every term enters twice: once per 16-byte half)
    for (int i = threadIdx.x; i < AT8_EQ_BYTES / 16; i += NT) *reinterpret_cast<uint4*>(sEq + i * 16) = make_uint4(0u, 0x00003C00u, 0u, 0u);
    for (int i = threadIdx.x; i < AT8_EK_BYTES / 16; i += NT) {
        const int pair = (i >> 1) >> 1;                                  // S column = E_k row = i >> 1; two columns per fp16 pair
        const uint32_t half_off = ((pair & 7) < POLY) ? 0x4980u /* OFF_POLY / 2 = 11 */ : 0x4300u /* OFF_SFU / 2 = 3.5 */;
        *reinterpret_cast<uint4*>(sEk + i * 16) = make_uint4(0x38003800u, half_off, 0u, 0u);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();

    if (warp == IW) {
        // ===================== issuer warp: sleeps in the named barriers between hand-offs =====================
        constexpr uint32_t idesc_s = umma_idesc_f16_ex(128, ATC_BN / 2, 0);
        constexpr uint32_t idesc_o = umma_idesc_f16_ex(128, 2 * ATC_D, 1);
        const uint64_t dq = umma_desc(smem_u32(sQ), 0, 256, 6);
        const uint64_t dk0 = umma_desc(smem_u32(sK), 0, 256, 6);
        const uint64_t dv0 = umma_desc(smem_u32(sV), ATC_KV_BYTES, 256, 6);
        const uint64_t deq = umma_desc(smem_u32(sEq), 0, 256, 6);
        const uint64_t dek = umma_desc(smem_u32(sEk), 0, 256, 6);
        constexpr uint64_t K_HALF = (ATC_BN / 2) * ATC_D * 2 / 16;       // second 32 keys of a K tile, in descriptor units
        auto load_kv = [&](int t) {                                      // one elected thread
            const int st = t & (ATC_STAGES - 1);
            mbar_arrive_expect_tx(&bars[KV_FULL + st], 2 * ATC_KV_BYTES);
            tma_load_2d(sK + st * ATC_KV_BYTES, &tmkv, &bars[KV_FULL + st], C + head * ATC_D, row_base + t * ATC_BN);
            tma_load_2d(sV + st * 2 * ATC_KV_BYTES, &tmkv, &bars[KV_FULL + st], 2 * C + head * ATC_D, row_base + t * ATC_BN);
        };
        auto issue_s_half = [&](int t, int ch) {                         // one elected thread
            const int ks = t & (ATC_STAGES - 1);
            mbar_wait_a(bar0 + 8 * (KV_FULL + ks), (t >> 2) & 1);
            tc_fence_after();
            const uint32_t d = tmem + ch * (ATC_BN / 2);
            umma_f16(d, dq, dk0 + (uint64_t)(ks * (ATC_KV_BYTES / 16)) + (uint64_t)ch * K_HALF, idesc_s, 0);
            umma_f16(d, deq, dek, idesc_s, 1);                           // - m_ref of every row
            umma_commit_a(bar0 + 8 * (S_FULL + ch));
        };
        if (elect_one()) {
            mbar_arrive_expect_tx(&bars[Q_FULL], ATC_TILE_BYTES);
            tma_load_2d(sQ, &tm, &bars[Q_FULL], head * ATC_D, row_base + q0);
            const int pre = nb < ATC_STAGES ? nb : ATC_STAGES;
            for (int t = 0; t < pre; ++t) load_kv(t);
        }
        __syncwarp();
        named_bar_sync(4, NT);                                   // Q has been rescaled in place by the softmax warps
        if (elect_one()) {
            issue_s_half(0, 0);
            issue_s_half(0, 1);
        }
        __syncwarp();
        for (int j = 0; j < nb; ++j) {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                named_bar_sync(1 + ch, NT);                      // all four softmax warps hold this half of S_j in registers
                tc_fence_after();
                if (j + 1 < nb) {
                    if (elect_one()) issue_s_half(j + 1, ch);
                    __syncwarp();
                }
            }
            named_bar_sync(3, NT);                               // P_j stored (and P.V_{j-1} seen retired by every softmax warp)
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dv = dv0 + (uint64_t)((j & (ATC_STAGES - 1)) * (2 * ATC_KV_BYTES / 16));
#pragma unroll
                for (int kk = 0; kk < ATC_BN / 16; ++kk)
                    umma_f16_ts(tmem + AT3_O_COL, tmem + AT3_P_COL + kk * 8, dv + (uint64_t)(kk * (512 / 16)), idesc_o, (j | kk) != 0);
                umma_commit_a(bar0 + 8 * P_EMPTY);
                if (j >= 1 && j + 3 < nb) load_kv(j + 3);        // the stage P.V_{j-1} has retired from
            }
            __syncwarp();
        }
    } else {
        // ===================== softmax warps (thread = query row) =====================
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        const int row = warp * 32 + lane;
        {   // Q *= scale * log2(e), in place (any 32 bytes of the tile per thread)
            mbar_wait_a(bar0 + 8 * Q_FULL, 0);
            const f162 sc = __float2half2_rn(scale_log2e);
            f162* qp = reinterpret_cast<f162*>(sQ + threadIdx.x * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) qp[i] = __hmul2(qp[i], sc);
            fence_proxy_async();
            named_bar_arrive(4, NT);
        }
        uint32_t* eq_row = reinterpret_cast<uint32_t*>(sEq + row * 32);
        float m_ref = 0.f;                       // reference maximum of this row (log2 units), always == hi + lo of its E_q row
        float m_eq = 0.f;                        // what the E_q row holds (negated) right now
        float mb[2] = {0.f, 0.f};                // reference baked into the S half that is loaded next
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
        auto rescale_p_first_half = [&](float fac) {     // the 32 fp16 P values of this block's first half (16 packed columns)
            tmem_st_wait();
            uint32_t pv[32];
            tmem_ld32(tl + AT3_P_COL, pv);               // columns 16..31 are read and dropped
            tmem_ld_wait();
            uint32_t p0[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float2 f = unpack_h2(pv[i]);
                p0[i] = pack_h2_nosat(f.x * fac, f.y * fac);
            }
            tmem_st16(tl + AT3_P_COL, p0);
        };

        for (int j = 0; j < nb; ++j) {
            bool waited_p = false;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                mbar_wait_a(bar0 + 8 * (S_FULL + ch), j & 1);
                tc_fence_after();
                uint32_t v[32];
                tmem_ld32(tl + ch * 32, v);
                tmem_ld_wait();
                tc_fence_before();
                named_bar_arrive(1 + ch, NT);                    // releases this half: S_{j+1} is issued with E_q as it is NOW
                const float d0 = m_ref - mb[ch];                 // this half was issued d0 ago (0 unless the reference moved since)
                mb[ch] = m_eq;
                float mx[2] = {-INFINITY, -INFINITY};            // [0]: SFU columns (x + 7), [1]: polynomial columns (x + 22)
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const int cls = (((i >> 1) & 7) < POLY) ? 1 : 0;
                    mx[cls] = fmaxf(mx[cls], fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
                }
                const bool first = (j == 0) && (ch == 0);
                // fast-path test without forming the maximum: with d0 == 0 the class maxima are relative to m_ref already
                const bool slow = d0 != 0.0f || mx[0] > 8.0f + AT8_OFF_SFU || mx[1] > 8.0f + AT8_OFF_POLY || first;
                if (__any_sync(0xffffffffu, slow)) {                         // rare after the first blocks
                    const float cm = fmaxf(mx[0] - AT8_OFF_SFU, mx[1] - AT8_OFF_POLY) - d0;   // maximum of the half relative to m_ref
                    const bool move = cm > 8.0f || first;
                    float m_new = m_ref;
                    if (move) {                                  // new reference, exactly representable as hi + lo
                        const float t = m_ref + cm;
                        const f16 hi = __float2half_rn(t);
                        const f16 lo = __float2half_rn(t - __half2float(hi));
                        m_new = __half2float(hi) + __half2float(lo);
                    }
                    if (__any_sync(0xffffffffu, move)) {
                        if (!waited_p) {
                            mbar_wait_a(bar0 + 8 * P_EMPTY, (j & 1) ^ 1);
                            tc_fence_after();
                            waited_p = true;
                        }
                        const float fac = move ? ex2_approx(m_ref - m_new) : 1.0f;
                        if (j > 0) rescale_o(fac);               // block 0: O is not initialised yet (P.V_0 overwrites)
                        if (ch == 1) rescale_p_first_half(fac);
                        // E_q may only change while no S MMA that reads it is in flight: the halves released so far have been issued
                        // (this warp has arrived for them), wait until they have completed
                        if (j + 1 < nb) mbar_wait_a(bar0 + 8 * (S_FULL + 0), (j + 1) & 1);
                        if (ch == 0) mbar_wait_a(bar0 + 8 * (S_FULL + 1), j & 1);
                        else if (j + 1 < nb) mbar_wait_a(bar0 + 8 * (S_FULL + 1), (j + 1) & 1);
                        if (move) {
                            const f16 hi = __float2half_rn(m_new);
                            const f16 lo = __float2half_rn(m_new - __half2float(hi));
                            const uint32_t w = (uint32_t)__half_as_ushort(__hneg(hi)) | ((uint32_t)__half_as_ushort(__hneg(lo)) << 16);
                            eq_row[0] = w;                       // both 16-byte halves of the row: invariant under the 32 B swizzle
                            eq_row[4] = w;
                            m_eq = m_new;
                        }
                        fence_proxy_async();
                    }
                    const float shift = m_new - (m_ref - d0);    // what this half still has to lose: d0 + (m_new - m_ref)
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) - shift);
                    m_ref = m_new;
                }
                uint32_t pk[16];
#pragma unroll
                for (int idx = 0; idx < 16 + AT_PIPE; ++idx) {
                    if (idx < 16) {
                        const float s0 = __uint_as_float(v[2 * idx]), s1 = __uint_as_float(v[2 * idx + 1]);
                        if ((idx & 7) < POLY) {
                            pk[idx] = ex2_pair_poly_off(s0, s1);
                        } else {
                            v[2 * idx] = __float_as_uint(ex2_ordered(s0));
                            v[2 * idx + 1] = __float_as_uint(ex2_ordered(s1));
                        }
                    }
                    if (idx >= AT_PIPE) {
                        const int i = idx - AT_PIPE;
                        if ((i & 7) >= POLY) pk[i] = pack_h2_ordered(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                    }
                }
                if (!waited_p) {                                 // P.V of block j-1 has finished reading the P buffer
                    mbar_wait_a(bar0 + 8 * P_EMPTY, (j & 1) ^ 1);
                    tc_fence_after();
                    waited_p = true;
                }
                tmem_st16(tl + AT3_P_COL + ch * 16, pk);
            }
            tmem_st_wait();
            tc_fence_before();
            named_bar_arrive(3, NT);
        }
        // ---- epilogue: O / l -> fp16
        mbar_wait_a(bar0 + 8 * P_EMPTY, (nb - 1) & 1);
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
    if (warp == IW) {
        tc_fence_after();
        tmem_dealloc(tmem, AT3_TMEM_COLS);
    }
}

// fp16 pairs (of every 8) exponentiated on the FMA pipe; B2D_ATTN_POLY overrides
inline int attn_tc8_poly() {
    static const int v = [] {
        const char* e = getenv("B2D_ATTN_POLY");
        const int p = e ? atoi(e) : 4;
        return p < 0 ? 0 : (p > 5 ? 5 : p);
    }();
    return v;
}

template <int POLY>
inline int attn_tc8_attr() {
    B2D_CUDA(cudaFuncSetAttribute(attn_tc8_kernel<POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT8_SMEM));
    B2D_CUDA(cudaFuncSetAttribute(attn_tc8_kernel<POLY>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return 0;
}
inline int attn_tc8_init_attrs() {
    B2D_TRY(attn_tc8_attr<0>());
    B2D_TRY(attn_tc8_attr<2>());
    B2D_TRY(attn_tc8_attr<3>());
    B2D_TRY(attn_tc8_attr<4>());
    B2D_TRY(attn_tc8_attr<5>());
    return 0;
}

inline int attn_tc8_launch(const AttnTcMaps& m, f16* o, int B, int L, int C, int heads, cudaStream_t st) {
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)ATC_D);
    const dim3 grid(L / ATC_BLK, heads, B), block(AT8_THREADS);
    switch (attn_tc8_poly()) {
        case 0: case 1: B2D_CUDA(launch_k(attn_tc8_kernel<0>, grid, block, AT8_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        case 2: B2D_CUDA(launch_k(attn_tc8_kernel<2>, grid, block, AT8_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        case 3: B2D_CUDA(launch_k(attn_tc8_kernel<3>, grid, block, AT8_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        case 4: B2D_CUDA(launch_k(attn_tc8_kernel<4>, grid, block, AT8_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        default: B2D_CUDA(launch_k(attn_tc8_kernel<5>, grid, block, AT8_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
    }
    return 0;
}

}  // namespace b2d
