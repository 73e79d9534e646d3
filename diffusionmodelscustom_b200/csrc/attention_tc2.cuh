// Persistent tcgen05 flash attention for head_dim 16 (C = 64, 4 heads) — the L = 1024 / 4096 layers that are 59-76 % of
// the 128x128 step.  Replaces attn_tc_kernel (attention_tc.cuh, kept for L % 256 != 0 and as the A/B baseline).
//
// What bounds this layer: one exponential per 64 MMA-FLOP.  The SFU issues 16 ex2/clk/SM (measured 4.63 Texp/s over the chip,
// tools/ub/mufu.cu) while TMEM reads are NOT the wall (tools/ub/tmem_ld.cu: tcgen05.ld sustains >= 160 B/clk/SM, S needs 4 B per
// score).  attn_tc_kernel reached 69 % of the SFU rate: its softmax warps alternate load / max / exp / store phases and only
// the occupancy of three CTAs overlaps them.  This kernel removes the phases instead:
//   * ONE persistent CTA per SM walks a static list of work items (sample, head, 256 consecutive queries = two 128-row
//     tiles that share every K/V block); the TMA producer and the MMA issuers run ahead across item boundaries, so
//     pipeline fill/drain is paid once per CTA, not once per tile.
//   * eight softmax warps (thread = query row, as before); each keeps TWO register copies of S: while block n is being
//     exponentiated the tcgen05.ld of block n+1 is already in flight, and the MMA issuer of the tile produces S two blocks
//     ahead (two S buffers in TMEM), so a warp's instruction stream is MUFU/FMA work back to back.
//   * a compile-time fraction of the exponentials (POLY of every 8 fp16 pairs) is evaluated on the FMA pipe in PACKED HALF
//     precision instead of the SFU: x' = s*scale - m + 15 is clamped at 0 by the converter itself (cvt.rn.relu.f16x2), split
//     with the 1536 magic constant into n' + f, 2^f is a degree-3 HFMA2 polynomial and 2^(n'-15) is built by shifting n'
//     into the exponent field (n' = 0 gives +0.0: far-below-maximum scores flush to zero exactly like the SFU path): 10
//     instructions per PAIR against 2 MUFU slots of 8 clk each.
// Online softmax with the lazy reference maximum of attn_tc_kernel (O is rescaled in TMEM only when a block exceeds the row's
// reference by more than 2^8); the denominator comes out of the P.V MMA through the [V | ones] MN-major operand.
//   warp 0: TMA producer   warp 1 / warp 10: MMA issuer of tile 0 / tile 1 (warp 1 owns TMEM)   warps 2-9: softmax
// Nothing on a softmax warp's path waits for an event of the SAME block: S is double-buffered in TMEM and in registers (its
// buffer is released as soon as the prefetch has landed), P is double-buffered in TMEM (storing P_n only needs P.V_{n-2}), so the
// single-thread MMA issuers (~100 clk per tcgen05.mma issued, five per block and tile) have a whole block of slack.
// TMEM, per tile (256 columns each): S0 [0,64) | S1 [64,128) | P0 [128,160) | P1 [160,192) | O [192,208) | denominators [208,224)
#pragma once
#include "attention_tc.cuh"

namespace b2d {

constexpr int AT2_THREADS = 352;
constexpr int AT2_QTILES = 2;
constexpr int AT2_STAGES = 4;
constexpr int AT2_TILE_COLS = 256;
constexpr int AT2_S_COL = 0, AT2_P_COL = 128, AT2_O_COL = 192;   // S0 S1 | P0 P1 | O, denominators
constexpr int AT2_SMEM = 1024 + 2 * AT2_QTILES * ATC_TILE_BYTES + ATC_KV_BYTES * 3 * AT2_STAGES + 512;

// 2^x' for a pair of fp32 arguments, on the FMA pipe in packed half precision; x' >= 0 after the converter's clamp, x' <= 24.
__device__ __forceinline__ uint32_t ex2_pair_poly(float xa, float xb) {
    uint32_t h, xr, nf, f, p, r;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(xb), "f"(xa));             // low half <- xa; negative -> +0
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(xr) : "r"(h), "r"(0x66006600u));              // + 1536: integer part lands in the mantissa
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(nf) : "r"(xr), "r"(0x66006600u));             // n' as a half
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(f) : "r"(h), "r"(nf));                        // f in [-0.5, 0.5]
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(0x2B0D2B0Du), "r"(f), "r"(0x33C333C3u));   // 0.05509 f + 0.24260
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(p), "r"(f), "r"(0x398C398Cu));              // .. f + 0.69328
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(p), "r"(f), "r"(0x3C003C00u));              // .. f + 1
    const uint32_t e = (xr << 10) & 0x7C007C00u;                                         // 2^(n' - 15) as half bits; n' = 0 -> +0.0
    asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(p), "r"(e));
    return r;
}

constexpr int AT_PIPE = 6;   // pairs between an SFU request and the pack that consumes it
__device__ __forceinline__ float ex2_ordered(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_h2_ordered(float a, float b) {   // inputs in [0, 2^8]: no saturation needed
    uint32_t r;
    asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

// tcgen05.wait::ld that also names the destination registers, so that no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
                   "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                   "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}

// mbarrier / commit helpers on raw shared-memory addresses (the issuer loops keep them in registers: no address arithmetic
// or generic->shared conversion per use)
__device__ __forceinline__ void mbar_wait_a(uint32_t addr, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP_A:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"   // %2: suspend-time hint — the thread sleeps in the
        "@P1 bra DONE_A;\n\t"                                               // barrier unit instead of re-issuing try_wait + branch
        "bra WAIT_LOOP_A;\n\t"                                              // (a fifth of all issued instructions without it)
        "DONE_A:\n\t"
        "}" ::"r"(addr),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void umma_commit_a(uint32_t addr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}

// Bring-up aid (-DAT2_TRACE): clock64 stamps of one softmax warp and one MMA issuer of CTA 0, per key block
constexpr int AT2_TRACE_BLOCKS = 256;
__device__ long long g_at2_trace[AT2_TRACE_BLOCKS][12];
#ifdef AT2_TRACE
#define AT2_STAMP(cond, n, slot) do { if ((cond) && (n) < AT2_TRACE_BLOCKS) g_at2_trace[(n)][(slot)] = clock64(); } while (0)
#else
#define AT2_STAMP(cond, n, slot) do { } while (0)
#endif

template <int POLY>
__global__ void __launch_bounds__(AT2_THREADS, 1)
    attn_tc2_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmkv, f16* __restrict__ o, int L,
                    int C, int heads, int items, float scale_log2e) {
    pdl_launch_dependents();
    extern __shared__ uint8_t at2_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at2_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                               // [2 stages][2 tiles] x 4 KB
    uint8_t* sK = sQ + 2 * AT2_QTILES * ATC_TILE_BYTES;               // [4] x 2 KB
    uint8_t* sV = sK + AT2_STAGES * ATC_KV_BYTES;                     // [4] x (V 2 KB | ones 2 KB)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + AT2_STAGES * 2 * ATC_KV_BYTES);
    // barrier indices (8 bytes each)
    constexpr int KV_FULL = 0, KV_EMPTY = 4, Q_FULL = 8, Q_EMPTY = 10, S_FULL = 12, S_EMPTY = 16, P_FULL = 20, P_EMPTY = 24, O_FULL = 28,
                  O_EMPTY = 30, NBARS = 32;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
    const uint32_t bar0 = smem_u32(bars);
#define AT2_BAR(idx) (bar0 + 8u * (uint32_t)(idx))

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nb = L / ATC_BN;                                        // key blocks per item
    const int items_per_bh = L / (AT2_QTILES * ATC_BLK);
    const int my_items = blockIdx.x < items ? (items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int nblocks = my_items * nb;                                // the CTA's stream of key blocks over all its items

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm); tma_prefetch_desc(&tmkv); }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < AT2_STAGES; ++i) {
                mbar_init(&bars[KV_FULL + i], 1);
                mbar_init(&bars[KV_EMPTY + i], 2);       // the P.V MMAs of both tiles
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&bars[Q_FULL + i], 1); mbar_init(&bars[Q_EMPTY + i], 2);
                mbar_init(&bars[O_FULL + i], 1); mbar_init(&bars[O_EMPTY + i], 4);
            }
            for (int i = 0; i < 4; ++i) {     // [tile][buffer]
                mbar_init(&bars[S_FULL + i], 1); mbar_init(&bars[S_EMPTY + i], 4);
                mbar_init(&bars[P_FULL + i], 4); mbar_init(&bars[P_EMPTY + i], 1);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    if (warp >= 2 && warp < 10) {   // constant ones tiles (generic-proxy writes -> visible to the async proxy after the fence)
        for (int i = threadIdx.x - 64; i < AT2_STAGES * ATC_BN * 2; i += 256) {
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
        // ===================== TMA producer: Q per item, one K+V stage per key block =====================
        if (lane == 0) {
            int n = 0;
            for (int it = 0; it < my_items; ++it) {
                const int item = blockIdx.x + it * gridDim.x;
                const int bh = item / items_per_bh, qi = item - bh * items_per_bh;
                const int b = bh / heads, head = bh - b * heads;
                const int row_base = b * L, q0 = qi * (AT2_QTILES * ATC_BLK);
                const int qs = it & 1;
                mbar_wait_a(AT2_BAR(Q_EMPTY + qs), ((it >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&bars[Q_FULL + qs], AT2_QTILES * ATC_TILE_BYTES);
                for (int g = 0; g < AT2_QTILES; ++g)
                    tma_load_2d(sQ + (qs * AT2_QTILES + g) * ATC_TILE_BYTES, &tm, &bars[Q_FULL + qs], head * ATC_D, row_base + q0 + g * ATC_BLK);
                for (int t = 0; t < nb; ++t, ++n) {
                    const int st = n & (AT2_STAGES - 1);
                    mbar_wait_a(AT2_BAR(KV_EMPTY + st), ((n >> 2) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bars[KV_FULL + st], 2 * ATC_KV_BYTES);
                    tma_load_2d(sK + st * ATC_KV_BYTES, &tmkv, &bars[KV_FULL + st], C + head * ATC_D, row_base + t * ATC_BN);
                    tma_load_2d(sV + st * 2 * ATC_KV_BYTES, &tmkv, &bars[KV_FULL + st], 2 * C + head * ATC_D, row_base + t * ATC_BN);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1 || warp == 10) {
        // ===================== MMA issuer of tile g: S two blocks ahead of P.V =====================
        // The loop is kept as lean as possible (running counters instead of divisions, descriptors and barrier addresses in
        // registers): a single thread executes it at one instruction per ~7 clk next to the softmax warps of its scheduler, and
        // with the first version (~300 instructions per iteration) the issuer, not the SFU, paced the kernel (2100 clk per block).
        // All 32 lanes run the loop CONVERGED (waits, counters and descriptor arithmetic are warp-uniform, so they live on the
        // uniform datapath); only the tcgen05.mma / tcgen05.commit instructions themselves are issued by lane 0.  With the whole
        // loop inside `if (lane == 0)` every operand of those instructions had to be moved register -> uniform register first.
        if (nblocks > 0) {
            const int g = warp == 1 ? 0 : 1;
            const uint32_t tbase = tmem + g * AT2_TILE_COLS;
            constexpr uint32_t idesc_s = umma_idesc_f16_ex(128, ATC_BN, 0);    // S: A,B K-major, N = keys per block
            constexpr uint32_t idesc_o = umma_idesc_f16_ex(128, 2 * ATC_D, 1); // [O | denominators]: A from TMEM, B MN-major, N = 32
            const uint64_t dq0 = umma_desc(smem_u32(sQ + g * ATC_TILE_BYTES), 0, 256, 6);     // + (qs * 2 tiles * 4096) >> 4
            const uint64_t dk0 = umma_desc(smem_u32(sK), 0, 256, 6);                          // + (ks * 2048) >> 4
            const uint64_t dv0 = umma_desc(smem_u32(sV), ATC_KV_BYTES, 256, 6);               // + (vs * 4096 + kk * 512) >> 4
            const uint32_t b_sfull = AT2_BAR(S_FULL + g * 2), b_sempty = AT2_BAR(S_EMPTY + g * 2);
            const uint32_t b_pfull = AT2_BAR(P_FULL + g * 2), b_pempty = AT2_BAR(P_EMPTY + g * 2);
            const uint32_t b_ofull = AT2_BAR(O_FULL + g), b_oempty = AT2_BAR(O_EMPTY + g);
            int ns = 0, ts = 0, its = 0;                 // S cursor: global block, block within item, item
            int np = 0, tp = 0, itp = 0;                 // P.V cursor
            auto issue_s = [&]() {
                const int qs = its & 1, ks = ns & 3, sb = ns & 1;
                if (ts == 0) mbar_wait_a(AT2_BAR(Q_FULL + qs), (its >> 1) & 1);
                mbar_wait_a(AT2_BAR(KV_FULL + ks), (ns >> 2) & 1);
                mbar_wait_a(b_sempty + 8 * sb, ((ns >> 1) & 1) ^ 1);            // the softmax warps hold S_{n-2} in registers
                tc_fence_after();
                const bool last = ts == nb - 1;
                if (elect_one()) {
                    umma_f16(tbase + AT2_S_COL + sb * ATC_BN, dq0 + (uint64_t)(qs * (AT2_QTILES * ATC_TILE_BYTES / 16)),
                             dk0 + (uint64_t)(ks * (ATC_KV_BYTES / 16)), idesc_s, 0);
                    umma_commit_a(b_sfull + 8 * sb);
                    if (last) umma_commit_a(AT2_BAR(Q_EMPTY + qs));
                }
                __syncwarp();
                ++ns;
                if (last) {
                    ts = 0;
                    ++its;
                } else {
                    ++ts;
                }
            };
            auto issue_pv = [&]() {
                const int vs = np & 3, pb = np & 1;                              // its K+V stage was waited for by S_n already
                if (tp == 0) mbar_wait_a(b_oempty, (itp & 1) ^ 1);              // the previous item's O has been read out
                mbar_wait_a(b_pfull + 8 * pb, (np >> 1) & 1);
                AT2_STAMP(blockIdx.x == 0 && g == 0 && lane == 0, np, 10);
                tc_fence_after();
                const bool last = tp == nb - 1;
                if (elect_one()) {
                    const uint64_t dv = dv0 + (uint64_t)(vs * (2 * ATC_KV_BYTES / 16));
#pragma unroll
                    for (int kk = 0; kk < ATC_BN / 16; ++kk)
                        umma_f16_ts(tbase + AT2_O_COL, tbase + AT2_P_COL + pb * 32 + kk * 8, dv + (uint64_t)(kk * (512 / 16)), idesc_o,
                                    (tp | kk) != 0);
                    umma_commit_a(b_pempty + 8 * pb);
                    umma_commit_a(AT2_BAR(KV_EMPTY + vs));
                    if (last) umma_commit_a(b_ofull);
                }
                __syncwarp();
                ++np;
                if (last) {
                    tp = 0;
                    ++itp;
                } else {
                    ++tp;
                }
            };
            issue_s();
            if (nblocks > 1) issue_s();
            for (int n = 0; n < nblocks; ++n) {
                AT2_STAMP(blockIdx.x == 0 && g == 0 && lane == 0, n, 8);
                if (n + 2 < nblocks) issue_s();
                AT2_STAMP(blockIdx.x == 0 && g == 0 && lane == 0, n, 9);
                issue_pv();
                AT2_STAMP(blockIdx.x == 0 && g == 0 && lane == 0, n, 11);
            }
        }
        __syncwarp();
    } else {
        // ===================== softmax warps: tile g, TMEM lane quarter q =====================
        const int g = (warp - 2) >> 2;
        const int q = warp & 3;
        const uint32_t tbase = tmem + g * AT2_TILE_COLS + ((uint32_t)(q * 32) << 16);
        const int row = q * 32 + lane;
        uint32_t va[2][32], vb[2][32];
        const uint32_t b_sfull = AT2_BAR(S_FULL + g * 2), b_sempty = AT2_BAR(S_EMPTY + g * 2);
        const uint32_t b_pfull = AT2_BAR(P_FULL + g * 2), b_pempty = AT2_BAR(P_EMPTY + g * 2);
        if (nblocks > 0) {
            mbar_wait_a(b_sfull, 0);
            tc_fence_after();
            tmem_ld32(tbase + AT2_S_COL, va[0]);
            tmem_ld32(tbase + AT2_S_COL + 32, va[1]);
            tmem_ld_wait_regs(va[0]);
            tmem_ld_wait_regs(va[1]);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(b_sempty);
        }
        float m_ref = -INFINITY;
        int n = 0;
        // one key block: v holds S_n; vn receives S_{n+1} while the exponentials run
        auto block = [&](uint32_t (&v)[2][32], uint32_t (&vn)[2][32], int t) {
            const bool has_next = n + 1 < nblocks;
            const bool tr = blockIdx.x == 0 && warp == 2 && lane == 0;
            AT2_STAMP(tr, n, 0);
            if (has_next) {
                const int sb = (n + 1) & 1;
                mbar_wait_a(b_sfull + 8 * sb, ((n + 1) >> 1) & 1);
                tc_fence_after();
                tmem_ld32(tbase + AT2_S_COL + sb * ATC_BN, vn[0]);
                tmem_ld32(tbase + AT2_S_COL + sb * ATC_BN + 32, vn[1]);
            }
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                for (int i = 0; i < 32; i += 2)
                    mx[(i >> 1) & 3] = fmaxf(mx[(i >> 1) & 3], fmaxf(__uint_as_float(v[ch][i]), __uint_as_float(v[ch][i + 1])));
            if (has_next) {                                          // the prefetch has landed under the max: release its S buffer
                tmem_ld_wait_regs(vn[0]);
                tmem_ld_wait_regs(vn[1]);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_a(b_sempty + 8 * ((n + 1) & 1));
            }
            AT2_STAMP(tr, n, 1);
            const float bm = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * scale_log2e;
            if (t == 0) m_ref = -INFINITY;
            const bool move = bm > m_ref + 8.0f;                     // lazy reference: moves only on > 2^8 growth
            const float m_new = move ? bm : m_ref;
            const bool need_fix = move && (t > 0);                   // O already holds contributions relative to the old reference
            if (__any_sync(0xffffffffu, need_fix)) {
                mbar_wait_a(b_pempty + 8 * ((n - 1) & 1), ((n - 1) >> 1) & 1);   // P.V of block n-1 has retired; P.V of block n waits for our p_full
                tc_fence_after();
                uint32_t ov[32];
                tmem_ld32(tbase + AT2_O_COL, ov);
                tmem_ld_wait();
                const float fac = need_fix ? ex2_approx(m_ref - m_new) : 1.0f;
                uint32_t o0[16], o1[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    o0[i] = __float_as_uint(__uint_as_float(ov[i]) * fac);
                    o1[i] = __float_as_uint(__uint_as_float(ov[16 + i]) * fac);
                }
                tmem_st16(tbase + AT2_O_COL, o0);
                tmem_st16(tbase + AT2_O_COL + 16, o1);
                tmem_st_wait();
            }
            m_ref = m_new;
            const float neg_m = -m_ref, neg_m15 = 15.0f - m_ref;
            // The SFU results are consumed AT_PIPE pairs after they were requested: MUFU.EX2 has a latency of tens of cycles, and a
            // pack that directly follows its two exponentials (what the compiler schedules to keep live ranges short) stalls the
            // warp once per pair — measured 2000 clk per 64-key block for a warp, 4x the SFU time.  Volatile asm pins the order
            // MUFU(i) ... pack(i - AT_PIPE); the polynomial pairs and the FFMAs are ordinary instructions that fill the gaps.
            AT2_STAMP(tr, n, 2);
            uint32_t pk[2][16];
#pragma unroll
            for (int idx = 0; idx < 32 + AT_PIPE; ++idx) {
                if (idx < 32) {
                    const int ch = idx >> 4, i = idx & 15;
                    const float s0 = __uint_as_float(v[ch][2 * i]), s1 = __uint_as_float(v[ch][2 * i + 1]);
                    if ((i & 7) < POLY) {
                        pk[ch][i] = ex2_pair_poly(fmaf(s0, scale_log2e, neg_m15), fmaf(s1, scale_log2e, neg_m15));
                    } else {
                        v[ch][2 * i] = __float_as_uint(ex2_ordered(fmaf(s0, scale_log2e, neg_m)));
                        v[ch][2 * i + 1] = __float_as_uint(ex2_ordered(fmaf(s1, scale_log2e, neg_m)));
                    }
                }
                if (idx >= AT_PIPE) {
                    const int ch = (idx - AT_PIPE) >> 4, i = (idx - AT_PIPE) & 15;
                    if ((i & 7) >= POLY) pk[ch][i] = pack_h2_ordered(__uint_as_float(v[ch][2 * i]), __uint_as_float(v[ch][2 * i + 1]));
                }
            }
            AT2_STAMP(tr, n, 3);
            mbar_wait_a(b_pempty + 8 * (n & 1), ((n >> 1) & 1) ^ 1); // P.V of block n-2 has finished reading this P buffer
            AT2_STAMP(tr, n, 4);
            tc_fence_after();
            tmem_st16(tbase + AT2_P_COL + (n & 1) * 32, pk[0]);
            tmem_st16(tbase + AT2_P_COL + (n & 1) * 32 + 16, pk[1]);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(b_pfull + 8 * (n & 1));
            AT2_STAMP(tr, n, 5);
            AT2_STAMP(tr, n, 6);
            ++n;
        };
        for (int it = 0; it < my_items; ++it) {
            for (int t = 0; t < nb; t += 2) {                        // nb is even (L % 256 == 0)
                block(va, vb, t);
                block(vb, va, t + 1);
            }
            // ---- item epilogue: O / l -> fp16
            const int item = blockIdx.x + it * gridDim.x;
            const int bh = item / items_per_bh, qi = item - bh * items_per_bh;
            const int b = bh / heads, head = bh - b * heads;
            mbar_wait_a(AT2_BAR(O_FULL + g), it & 1);
            tc_fence_after();
            uint32_t ov[32];
            tmem_ld32(tbase + AT2_O_COL, ov);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(AT2_BAR(O_EMPTY + g));
            const float inv = 1.0f / __uint_as_float(ov[16]);
            f16* op = o + ((size_t)(b * L + qi * (AT2_QTILES * ATC_BLK) + g * ATC_BLK + row)) * C + head * ATC_D;
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
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

inline bool attn_tc2_supported(int L, int C, int heads) {
    return C / heads == ATC_D && C % heads == 0 && L % (AT2_QTILES * ATC_BLK) == 0;
}

// fraction of fp16 pairs (of every 8) exponentiated on the FMA pipe; B2D_ATTN_POLY overrides (0..4)
inline int attn_tc2_poly() {
    static const int v = [] {
        const char* e = getenv("B2D_ATTN_POLY");
        int p = e ? atoi(e) : 3;
        return p < 0 ? 0 : (p > 4 ? 4 : p);
    }();
    return v;
}

inline int attn_tc2_init_attrs() {
    B2D_CUDA(cudaFuncSetAttribute(attn_tc2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT2_SMEM));
    B2D_CUDA(cudaFuncSetAttribute(attn_tc2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT2_SMEM));
    B2D_CUDA(cudaFuncSetAttribute(attn_tc2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT2_SMEM));
    B2D_CUDA(cudaFuncSetAttribute(attn_tc2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT2_SMEM));
    B2D_CUDA(cudaFuncSetAttribute(attn_tc2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT2_SMEM));
    return 0;
}

inline int attn_tc2_launch(const AttnTcMaps& m, f16* o, int B, int L, int C, int heads, int num_sms, cudaStream_t st) {
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)ATC_D);
    const int items = B * heads * (L / (AT2_QTILES * ATC_BLK));
    const int grid = items < num_sms ? items : num_sms;
    switch (attn_tc2_poly()) {
        case 0: B2D_CUDA(launch_k(attn_tc2_kernel<0>, dim3(grid), dim3(AT2_THREADS), AT2_SMEM, st, m.q, m.kv, o, L, C, heads, items, scale_log2e)); break;
        case 1: B2D_CUDA(launch_k(attn_tc2_kernel<1>, dim3(grid), dim3(AT2_THREADS), AT2_SMEM, st, m.q, m.kv, o, L, C, heads, items, scale_log2e)); break;
        case 2: B2D_CUDA(launch_k(attn_tc2_kernel<2>, dim3(grid), dim3(AT2_THREADS), AT2_SMEM, st, m.q, m.kv, o, L, C, heads, items, scale_log2e)); break;
        case 3: B2D_CUDA(launch_k(attn_tc2_kernel<3>, dim3(grid), dim3(AT2_THREADS), AT2_SMEM, st, m.q, m.kv, o, L, C, heads, items, scale_log2e)); break;
        default: B2D_CUDA(launch_k(attn_tc2_kernel<4>, dim3(grid), dim3(AT2_THREADS), AT2_SMEM, st, m.q, m.kv, o, L, C, heads, items, scale_log2e)); break;
    }
    return 0;
}

}  // namespace b2d
