// tcgen05 flash attention for head_dim 16, SELF-ISSUING variant: the four-CTAs-per-SM design of attention_tc3.cuh without the
// TMA-producer and MMA-issuer warps.  A CTA is just its four softmax warps (128 threads, thread = query row).
//
// Why: in attn_tc3_kernel the two auxiliary warps of each CTA spin on mbarriers for the whole kernel — eight spinning warps per SM
// next to sixteen working ones.  ncu (profiles/r2_ncu_attn_tc3_full_*): 119 of the 493 warp instructions executed per 64-key block
// are try_wait / branch / yield, with the issue slots 74 % busy; and every hand-off goes softmax -> mbarrier -> issuer wake-up ->
// MMA -> mbarrier -> softmax wake-up, which sits on each CTA's serial chain once per block.  Here the hand-offs are counted in
// shared memory (atom.inc, wraps at 4): the LAST of the four warps to finish a phase issues the dependent tensor-core / TMA work
// itself, from one elected lane, and carries on.  Nothing polls except a warp that genuinely has to wait.
//
// S is produced in two 32-key halves with their own barriers, and a half is released as soon as all four warps hold it in
// registers: S_{j+1} (first half) is issued while block j is still being exponentiated, so the S round trip through the tensor
// core is off the critical path (attn_tc4_kernel tried that with full 32-key blocks and paid a P.V hand-off per 32 keys; here P
// and P.V stay 64 keys wide).  The early release rules out re-reading S, so the reference maximum is the plain lazy one at HALF
// granularity: maximum of the 32 scores in registers first, exponentials second; a move of the reference (> 2^8 growth, rare
// after the first blocks) rescales O in TMEM and — when it happens in the second half — the already stored first half of P.
//
// phase hand-offs of block j (counter -> what its last arriver issues):
//   CNT_A: S first half in registers   -> S_{j+1} first half  (K_{j+1} has landed long ago)
//   CNT_B: S second half in registers  -> S_{j+1} second half; TMA of K/V block j+3 into the stage P.V_{j-1} has retired from
//   CNT_P: P_j stored                  -> P.V_j (4 MMAs, [V|ones] B operand), commit -> P_EMPTY
#pragma once
#include "attention_tc3.cuh"

namespace b2d {

constexpr int AT6_THREADS = 128;
constexpr int AT6_CTAS_PER_SM = 4;
constexpr int AT6_SMEM = 1024 + ATC_TILE_BYTES + ATC_KV_BYTES * 3 * ATC_STAGES + 256;

// shared-memory phase counter: returns the number of earlier arrivals of this phase (0..3) and wraps back to 0 after the fourth
__device__ __forceinline__ uint32_t at6_arrive(uint32_t addr) {
    uint32_t old;
    asm volatile("atom.acq_rel.cta.shared::cta.inc.u32 %0, [%1], 3;" : "=r"(old) : "r"(addr) : "memory");
    return old;
}

// MODE 0: self-issuing (128 threads).  MODE 1: a fifth warp issues all tensor-core / TMA work and BLOCKS on hardware named barriers
// (bar.sync against the softmax warps' bar.arrive) instead of polling mbarriers: a waiting issuer takes no issue slots at all.
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int POLY, int MODE>
__global__ void __launch_bounds__(MODE ? AT6_THREADS + 32 : AT6_THREADS, AT6_CTAS_PER_SM)
    attn_tc6_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmkv, f16* __restrict__ o, int L,
                    int C, float scale_log2e) {
    pdl_launch_dependents();
    extern __shared__ uint8_t at6_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at6_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + ATC_TILE_BYTES;
    uint8_t* sV = sK + ATC_STAGES * ATC_KV_BYTES;                    // [V tile 2 KB | ones tile 2 KB] per stage
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATC_STAGES * 2 * ATC_KV_BYTES);
    constexpr int KV_FULL = 0, S_FULL = 4 /* two halves */, P_EMPTY = 6, Q_FULL = 7, NBARS = 8;
    uint32_t* cnt = reinterpret_cast<uint32_t*>(bars + NBARS);       // CNT_A, CNT_B, CNT_P
    uint32_t* tmem_slot = cnt + 4;
    const uint32_t bar0 = smem_u32(bars), cnt0 = smem_u32(cnt);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.y, b = blockIdx.z;
    const int q0 = blockIdx.x * ATC_BLK;
    const int nb = L / ATC_BN;
    const int row_base = b * L;

    constexpr int IW = MODE ? 4 : 0;           // the warp that sets up barriers / TMEM and runs the prologue
    constexpr int NT = MODE ? AT6_THREADS + 32 : AT6_THREADS;
    if (warp == IW) {
        if (lane == 0) {
            tma_prefetch_desc(&tm);
            tma_prefetch_desc(&tmkv);
            for (int i = 0; i < NBARS; ++i) mbar_init(&bars[i], 1);
            cnt[0] = cnt[1] = cnt[2] = 0;
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, AT3_TMEM_COLS);
        tmem_relinquish();
    }
    // constant ones tiles (generic-proxy writes -> visible to the async proxy after the fence)
    for (int i = threadIdx.x; i < ATC_STAGES * ATC_BN * 2; i += NT) {
        const int st = i / (ATC_BN * 2), r = i % (ATC_BN * 2);
        *reinterpret_cast<uint4*>(sV + st * 2 * ATC_KV_BYTES + ATC_KV_BYTES + r * 16) = make_uint4(0x00003C00u, 0u, 0u, 0u);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();

    constexpr uint32_t idesc_s = umma_idesc_f16_ex(128, ATC_BN / 2, 0);
    constexpr uint32_t idesc_o = umma_idesc_f16_ex(128, 2 * ATC_D, 1);
    const uint64_t dq = umma_desc(smem_u32(sQ), 0, 256, 6);
    const uint64_t dk0 = umma_desc(smem_u32(sK), 0, 256, 6);
    const uint64_t dv0 = umma_desc(smem_u32(sV), ATC_KV_BYTES, 256, 6);
    constexpr uint64_t K_HALF = (ATC_BN / 2) * ATC_D * 2 / 16;       // second 32 keys of a K tile, in descriptor units

    auto load_kv = [&](int t) {                                      // one elected thread
        const int st = t & (ATC_STAGES - 1);
        mbar_arrive_expect_tx(&bars[KV_FULL + st], 2 * ATC_KV_BYTES);
        tma_load_2d(sK + st * ATC_KV_BYTES, &tmkv, &bars[KV_FULL + st], C + head * ATC_D, row_base + t * ATC_BN);
        tma_load_2d(sV + st * 2 * ATC_KV_BYTES, &tmkv, &bars[KV_FULL + st], 2 * C + head * ATC_D, row_base + t * ATC_BN);
    };
    auto issue_s_half = [&](int t, int ch) {                         // one elected thread; K_t's barrier is (re-)observed here
        const int ks = t & (ATC_STAGES - 1);
        mbar_wait_a(bar0 + 8 * (KV_FULL + ks), (t >> 2) & 1);
        tc_fence_after();
        umma_f16(tmem + ch * (ATC_BN / 2), dq, dk0 + (uint64_t)(ks * (ATC_KV_BYTES / 16)) + (uint64_t)ch * K_HALF, idesc_s, 0);
        umma_commit_a(bar0 + 8 * (S_FULL + ch));
    };

    if (warp == IW) {
        if (elect_one()) {
            mbar_arrive_expect_tx(&bars[Q_FULL], ATC_TILE_BYTES);
            tma_load_2d(sQ, &tm, &bars[Q_FULL], head * ATC_D, row_base + q0);
            const int pre = nb < ATC_STAGES ? nb : ATC_STAGES;
            for (int t = 0; t < pre; ++t) load_kv(t);
            mbar_wait_a(bar0 + 8 * Q_FULL, 0);
            issue_s_half(0, 0);
            issue_s_half(0, 1);
        }
        __syncwarp();
    }
    if (MODE && warp == IW) {
        // ===================== issuer warp: sleeps in the named barriers between hand-offs =====================
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
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
    const int row = warp * 32 + lane;
    float m_ref = -INFINITY;                 // reference maximum of this row (log2 units)
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
            if (MODE) {
                named_bar_arrive(1 + ch, NT);
            } else if (elect_one()) {
                if (at6_arrive(cnt0 + 4 * ch) == 3) {            // all four warps hold this half of S_j in registers
                    if (j + 1 < nb) issue_s_half(j + 1, ch);
                    // P.V_{j-1} has retired (this thread waited for it before its first P store of block j): its K/V stage is free
                    if (ch == 1 && j >= 1 && j + 3 < nb) load_kv(j + 3);
                }
            }
            __syncwarp();
            float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < 32; i += 2)
                mx[(i >> 1) & 1] = fmaxf(mx[(i >> 1) & 1], fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
            const float cm = fmaxf(mx[0], mx[1]) * scale_log2e;
            const bool move = cm > m_ref + 8.0f;
            if (__any_sync(0xffffffffu, move)) {                 // rare after the first blocks
                if (!waited_p) {
                    mbar_wait_a(bar0 + 8 * P_EMPTY, (j & 1) ^ 1);
                    tc_fence_after();
                    waited_p = true;
                }
                const float fac = move ? ex2_approx(m_ref - cm) : 1.0f;   // first block: 2^(-inf) = 0
                if (j > 0) rescale_o(fac);                               // block 0: O is not initialised yet (P.V_0 overwrites)
                if (ch == 1) rescale_p_first_half(fac);
                if (move) m_ref = cm;
            }
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
            if (!waited_p) {                                     // P.V of block j-1 has finished reading the P buffer
                mbar_wait_a(bar0 + 8 * P_EMPTY, (j & 1) ^ 1);
                tc_fence_after();
                waited_p = true;
            }
            tmem_st16(tl + AT3_P_COL + ch * 16, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        if (MODE) {
            named_bar_arrive(3, NT);
        } else if (elect_one()) {
            if (at6_arrive(cnt0 + 8) == 3) {                     // P_j complete: O += P_j [V_j | ones]
                tc_fence_after();
                const uint64_t dv = dv0 + (uint64_t)((j & (ATC_STAGES - 1)) * (2 * ATC_KV_BYTES / 16));
#pragma unroll
                for (int kk = 0; kk < ATC_BN / 16; ++kk)
                    umma_f16_ts(tmem + AT3_O_COL, tmem + AT3_P_COL + kk * 8, dv + (uint64_t)(kk * (512 / 16)), idesc_o, (j | kk) != 0);
                umma_commit_a(bar0 + 8 * P_EMPTY);
            }
        }
        __syncwarp();
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

template <int POLY>
inline int attn_tc6_attr() {
    B2D_CUDA(cudaFuncSetAttribute(attn_tc6_kernel<POLY, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT6_SMEM));
    B2D_CUDA(cudaFuncSetAttribute(attn_tc6_kernel<POLY, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    B2D_CUDA(cudaFuncSetAttribute(attn_tc6_kernel<POLY, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT6_SMEM));
    B2D_CUDA(cudaFuncSetAttribute(attn_tc6_kernel<POLY, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return 0;
}
inline int attn_tc6_init_attrs() {
    B2D_TRY(attn_tc6_attr<0>());
    B2D_TRY(attn_tc6_attr<2>());
    B2D_TRY(attn_tc6_attr<3>());
    B2D_TRY(attn_tc6_attr<4>());
    B2D_TRY(attn_tc6_attr<5>());
    B2D_TRY(attn_tc6_attr<6>());
    return 0;
}

template <int MODE>
inline int attn_tc6_launch(const AttnTcMaps& m, f16* o, int B, int L, int C, int heads, cudaStream_t st) {
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)ATC_D);
    const dim3 grid(L / ATC_BLK, heads, B), block(MODE ? AT6_THREADS + 32 : AT6_THREADS);
    switch (attn_tc3_poly()) {
        case 0: case 1: B2D_CUDA(launch_k(attn_tc6_kernel<0, MODE>, grid, block, AT6_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        case 2: B2D_CUDA(launch_k(attn_tc6_kernel<2, MODE>, grid, block, AT6_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        case 3: B2D_CUDA(launch_k(attn_tc6_kernel<3, MODE>, grid, block, AT6_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        case 4: B2D_CUDA(launch_k(attn_tc6_kernel<4, MODE>, grid, block, AT6_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        case 5: B2D_CUDA(launch_k(attn_tc6_kernel<5, MODE>, grid, block, AT6_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
        default: B2D_CUDA(launch_k(attn_tc6_kernel<6, MODE>, grid, block, AT6_SMEM, st, m.q, m.kv, o, L, C, scale_log2e)); break;
    }
    return 0;
}

}  // namespace b2d
