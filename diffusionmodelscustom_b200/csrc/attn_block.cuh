// Fused low-resolution self-attention: LayerNorm + QKV projection + softmax(QK^T)V for the levels with L = H*W <= 64 tokens
// (ImageSelfAttention at 8x8, 4x4, 2x2, 1x1: modules_DANRA_conditional.py:91-110; unet_ms.py:6-27), one launch instead of
// LayerNorm / QKV GEMM / attention / out-projection.
//
// CTA = (tile of 128 consecutive tokens = 128/L whole samples, one head).  All MMAs are tcgen05 with accumulators in TMEM:
//   1. QKV_h = x W'_h^T over K = C in 64-wide k-blocks (TMA ring; A = raw activations, B = the head's D rows of the
//      gamma-folded Wq, Wk, Wv as one 3-D TMA box) -> TMEM columns [0,3D).  The epilogue warps read the A stages as they pass
//      and derive each row's LayerNorm mean / rstd, which is applied algebraically at read-out: rstd*acc - rstd*mu*c1 + b'.
//   2. read-out (thread = token row): Q (pre-scaled by log2(e)/sqrt(D)) and K are written to shared memory as 128B-swizzled
//      K-major operands, V is written TRANSPOSED ([D][128 keys]) so that it is a K-major B operand for P.V.
//   3. S = Q K^T for the whole tile (128 x 128); each row only uses the L columns of its own sample (block-diagonal mask),
//      softmax in registers, P (fp16, zero outside the sample) goes back to TMEM and is the A operand of O = P V.
//   4. O / rowsum -> fp16.  Without the fused out-projection it is stored to global [rows][C] at the head's columns.  With it
//      (the CTAs of a tile's heads form a (1, heads, 1) cluster) O_h becomes the A operand of out_h = O_h Wo[:, hD:(h+1)D]^T
//      (128 x C, fp32 in TMEM); the per-head partials meet in an L2 workspace, and after a cluster barrier CTA r finalises
//      128/heads rows: sum over heads in fixed order + bias + residual (+ ReLU) -> fp16.  A second variant (fuse_out == 2) exchanges
//      the fp16 head outputs instead: every CTA scatters O_h through DSMEM into all peers' gathered [128][C] operand and then
//      computes its own D output channels over K = C (no fp32 partials, no workspace).
// warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: statistics / read-out / softmax (two warps per TMEM
// lane quarter, splitting the columns).
#pragma once
#include "common.cuh"
#include "conv.cuh"

namespace b2d {

constexpr int AB_THREADS = 320;
template <int D>
__host__ __device__ constexpr int ab_stages() { return D == 128 ? 3 : 4; }
template <int D>
__host__ __device__ constexpr int ab_stage_bytes() { return CONV_A_BYTES + 3 * D * 128; }
template <int D>
__host__ __device__ constexpr int ab_dpad() { return D < 64 ? 64 : D; }
template <int D>
__host__ __device__ constexpr int ab_smem_bytes() { return 1024 + ab_stages<D>() * ab_stage_bytes<D>() + 3 * D * 8 + 128 * 16 + 128 * 4 + 512; }

struct AttnBlockParams {
    const float* c1;     // [3C] column sums of the gamma-folded weights
    const float* bias;   // [3C] folded bias
    f16* out;            // [M][C] attention output (before the out-projection)
    int M, C, L;
    float scale_log2e;
    // fused out-projection (fuse_out != 0)
    int fuse_out;
    const float* out_bias;   // [C]
    const f16* residual;     // [M][C] (the block input)
    f16* out_final;          // [M][C]
    float* ws;               // [tiles][heads][128][C] fp32 partials
    int final_act;           // 0 none, 1 ReLU
};

template <int D>
__global__ void __launch_bounds__(AB_THREADS, 1)
    attn_block_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmW, const AttnBlockParams p) {
    pdl_launch_dependents();
    constexpr int STAGES = ab_stages<D>();
    constexpr int SB = ab_stage_bytes<D>();
    constexpr int DPAD = ab_dpad<D>();
    constexpr int S_COL = 384, P_COL = 0, O_COL = 64;
    extern __shared__ uint8_t ab_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ab_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    // operands of the attention phase alias the (drained) ring
    uint8_t* sQ = ring;                                         // DPAD/64 atoms of [128 rows][128 B]
    uint8_t* sK = sQ + (DPAD / 64) * CONV_A_BYTES;
    uint8_t* sVt = sK + (DPAD / 64) * CONV_A_BYTES;             // 2 atoms (keys 0-63, 64-127) of [D rows][128 B]
    static_assert(2 * (DPAD / 64) * CONV_A_BYTES + 2 * D * 128 <= STAGES * SB, "attention operands must fit the ring");
    float* s_c1 = reinterpret_cast<float*>(ring + STAGES * SB);  // [3D]
    float* s_bias = s_c1 + 3 * D;                                // [3D]
    float2* s_stat = reinterpret_cast<float2*>(s_bias + 3 * D);  // [2][128]
    float* s_l = reinterpret_cast<float*>(s_stat + 256);         // [128] softmax denominators
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_l + 128);
    uint64_t* full = bars;                // [STAGES]
    uint64_t* empty = bars + STAGES;      // [STAGES]
    uint64_t* accum_full = bars + 2 * STAGES;
    uint64_t* qk_ready = accum_full + 1;
    uint64_t* s_full = accum_full + 2;
    uint64_t* p_ready = accum_full + 3;
    uint64_t* o_full = accum_full + 4;
    uint64_t* w_full = accum_full + 5;    // [2] out-projection weight halves
    uint64_t* o_ready = accum_full + 7;
    uint64_t* out_full = accum_full + 8;
    uint64_t* w2_full = accum_full + 9;   // [4] gather mode: out_proj weight k-blocks
    uint64_t* w2_empty = accum_full + 13; // [4]
    uint64_t* out2_full = accum_full + 17;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 18);
    // out-projection weights: Wo[:, head*D .. +DPAD) as DPAD/64 K-atoms of [NW rows][128 B], in halves of NW <= 256 output channels
    constexpr int OPER_BYTES = 2 * (DPAD / 64) * CONV_A_BYTES + 2 * D * 128;
    const int NW = p.C > 256 ? 256 : p.C;
    const int nhalf = p.C > 256 ? 2 : 1;
    const int hbytes = (DPAD / 64) * NW * 128;
    uint8_t* sW0 = ring + OPER_BYTES;
    const bool w1_early = STAGES * SB - OPER_BYTES >= 2 * hbytes;
    uint8_t* sW1 = w1_early ? sW0 + hbytes : sK;       // late variant: over K / V^T once P.V has retired
    // gather mode (fuse_out == 2): every head's O (fp16) is scattered through DSMEM into each CTA's G = [C/64 atoms][128][128 B]
    // (over the dead attention operands); this CTA then computes out[:, head*D .. +D) = G Wo[head*D .. +D, :]^T over K = C with
    // the weight k-blocks ([D rows][128 B]) streamed through NS2 stages placed behind G.
    const int gbytes = 128 * p.C * 2;
    uint8_t* sG = ring;
    uint8_t* sW2 = ring + (gbytes > OPER_BYTES ? gbytes : OPER_BYTES);
    const int NS2 = (p.C >> 6) < 4 ? (p.C >> 6) : 4;
    float* s_ob = s_c1;                                // out_proj bias slice [D] (s_c1 is dead after the QKV read-out)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.y;
    const int m0 = blockIdx.x * 128;
    const int nkb = p.C >> 6;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (p.fuse_out) tma_prefetch_desc(&tmW);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < STAGES; ++i) {
                mbar_init(&full[i], 1);
                mbar_init(&empty[i], 1 + 8);     // MMA commit + the eight statistics warps
            }
            mbar_init(accum_full, 1);
            mbar_init(qk_ready, 8);
            mbar_init(s_full, 1);
            mbar_init(p_ready, 4);
            mbar_init(o_full, 1);
            mbar_init(&w_full[0], 1);
            mbar_init(&w_full[1], 1);
            mbar_init(o_ready, 8);
            mbar_init(out_full, 1);
            for (int i = 0; i < 4; ++i) {
                mbar_init(&w2_full[i], 1);
                mbar_init(&w2_empty[i], 1);
            }
            mbar_init(out2_full, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();

    uint32_t okeep[2][16];   // gather mode: this thread's share of O_h (fp16 pairs), alive across the cluster barrier
#pragma unroll
    for (int i = 0; i < 16; ++i) { okeep[0][i] = 0u; okeep[1][i] = 0u; }
    // single-thread roles are guarded by elect.sync (not `lane == 0`): ptxas then keeps the operands of the TMA / tcgen05
    // instructions on the uniform datapath instead of wrapping each one in an ELECT / R2UR.BROADCAST waterfall loop
    if (warp == 0) {
        if (elect_one()) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % STAGES;
                mbar_wait(&empty[st], ((kb / STAGES) & 1) ^ 1);
                mbar_arrive_expect_tx(&full[st], SB);
                tma_load_2d(ring + st * SB, &tmA, &full[st], kb * 64, m0);
                tma_load_3d(ring + st * SB + CONV_A_BYTES, &tmB, &full[st], kb * 64, head * D, 0);
            }
            if (p.fuse_out == 2) {
                mbar_wait(accum_full, 0);                      // the ring is drained: the region behind G is free
                for (int kb = 0; kb < NS2; ++kb) {
                    mbar_arrive_expect_tx(&w2_full[kb], (uint32_t)(D * 128));
                    tma_load_2d(sW2 + kb * (D * 128), &tmW, &w2_full[kb], kb * 64, head * D);
                }
            }
            if (p.fuse_out == 1) {
                mbar_wait(accum_full, 0);                      // the ring is drained: its tail is free for Wo
                for (int hh = 0; hh < nhalf; ++hh) {
                    if (hh == 1 && !w1_early) mbar_wait(o_full, 0);
                    uint8_t* dst = hh == 0 ? sW0 : sW1;
                    mbar_arrive_expect_tx(&w_full[hh], (uint32_t)hbytes);
                    for (int a = 0; a < DPAD / 64; ++a) tma_load_2d(dst + a * NW * 128, &tmW, &w_full[hh], head * D + a * 64, hh * NW);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            // ---- 1. QKV projection
            constexpr int N1 = (3 * D <= 256) ? 3 * D : 256;
            constexpr uint32_t idesc1 = umma_idesc_f16(128, N1);
            constexpr uint32_t idesc2 = umma_idesc_f16(128, 3 * D - N1 > 0 ? 3 * D - N1 : 16);
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % STAGES;
                mbar_wait(&full[st], (kb / STAGES) & 1);
                tc_fence_after();
                const uint64_t da = umma_desc_sw128(smem_u32(ring + st * SB));
                const uint64_t db = umma_desc_sw128(smem_u32(ring + st * SB + CONV_A_BYTES));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma_f16(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc1, (kb | k) != 0);
                    if (3 * D > N1) {
                        const uint64_t db2 = umma_desc_sw128(smem_u32(ring + st * SB + CONV_A_BYTES + N1 * 128));
                        umma_f16(tmem + N1, da + (uint64_t)(k * 2), db2 + (uint64_t)(k * 2), idesc2, (kb | k) != 0);
                    }
                }
                umma_commit(&empty[st]);
            }
            umma_commit(accum_full);
            // ---- 3. S = Q K^T
            mbar_wait(qk_ready, 0);
            tc_fence_after();
            constexpr uint32_t idesc_s = umma_idesc_f16(128, 128);
#pragma unroll
            for (int kk = 0; kk < DPAD / 16; ++kk) {
                const uint64_t dq = umma_desc_sw128(smem_u32(sQ + (kk >> 2) * CONV_A_BYTES)) + (uint64_t)((kk & 3) * 2);
                const uint64_t dk = umma_desc_sw128(smem_u32(sK + (kk >> 2) * CONV_A_BYTES)) + (uint64_t)((kk & 3) * 2);
                umma_f16(tmem + S_COL, dq, dk, idesc_s, kk != 0);
            }
            umma_commit(s_full);
            // ---- O = P V
            mbar_wait(p_ready, 0);
            tc_fence_after();
            constexpr uint32_t idesc_o = umma_idesc_f16(128, D);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const uint64_t dv = umma_desc_sw128(smem_u32(sVt + (kk >> 2) * (D * 128))) + (uint64_t)((kk & 3) * 2);
                umma_f16_ts(tmem + O_COL, tmem + P_COL + kk * 8, dv, idesc_o, kk != 0);
            }
            umma_commit(o_full);
            // ---- out_h = O_h Wo_h^T
            if (p.fuse_out == 1) {
                mbar_wait(o_ready, 0);
                tc_fence_after();
                const uint32_t idesc_w = umma_idesc_f16(128, NW);
                for (int hh = 0; hh < nhalf; ++hh) {
                    mbar_wait(&w_full[hh], 0);
                    tc_fence_after();
                    const uint8_t* wb = hh == 0 ? sW0 : sW1;
#pragma unroll
                    for (int kk = 0; kk < DPAD / 16; ++kk) {
                        const uint64_t da = umma_desc_sw128(smem_u32(sQ + (kk >> 2) * CONV_A_BYTES)) + (uint64_t)((kk & 3) * 2);
                        const uint64_t dw = umma_desc_sw128(smem_u32(wb + (kk >> 2) * NW * 128)) + (uint64_t)((kk & 3) * 2);
                        umma_f16(tmem + hh * 256, da, dw, idesc_w, kk != 0);
                    }
                }
                umma_commit(out_full);
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int hf = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const int sw = row & 7;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        // LN column sums / folded bias of this head's 3D columns (visible to all epilogue warps after the statistics barrier)
        for (int i = threadIdx.x - 64; i < 3 * D; i += 256) {
            const int mat = i / D, d = i - mat * D;
            s_c1[i] = __ldg(p.c1 + mat * p.C + head * D + d);
            s_bias[i] = __ldg(p.bias + mat * p.C + head * D + d);
        }
        // ---- LayerNorm statistics from the A stages (each half-warp-group sums 4 of the 8 chunks of a 64-channel block)
        float s1 = 0.f, s2 = 0.f;
        for (int kb = 0; kb < nkb; ++kb) {
            const int st = kb % STAGES;
            mbar_wait(&full[st], (kb / STAGES) & 1);
            const uint4* arow = reinterpret_cast<const uint4*>(ring + st * SB + row * 128);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const uint4 a4 = arow[(hf * 4 + jj) ^ sw];
                float2 tt;
                tt = unpack_h2(a4.x); s1 += tt.x + tt.y; s2 = fmaf(tt.x, tt.x, s2); s2 = fmaf(tt.y, tt.y, s2);
                tt = unpack_h2(a4.y); s1 += tt.x + tt.y; s2 = fmaf(tt.x, tt.x, s2); s2 = fmaf(tt.y, tt.y, s2);
                tt = unpack_h2(a4.z); s1 += tt.x + tt.y; s2 = fmaf(tt.x, tt.x, s2); s2 = fmaf(tt.y, tt.y, s2);
                tt = unpack_h2(a4.w); s1 += tt.x + tt.y; s2 = fmaf(tt.x, tt.x, s2); s2 = fmaf(tt.y, tt.y, s2);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
        s_stat[hf * 128 + row] = make_float2(s1, s2);
        named_bar_sync(1, 256);
        float ln_a, ln_b;
        {
            const float2 v0 = s_stat[row], v1 = s_stat[128 + row];
            const float invk = 1.0f / (float)p.C;
            const float mu = (v0.x + v1.x) * invk;
            ln_a = rsqrtf(fmaxf((v0.y + v1.y) * invk - mu * mu, 0.f) + 1e-5f);
            ln_b = -ln_a * mu;
        }
        // ---- 2. read-out of Q, K, V (32-column chunks, alternating between the two warps of a lane quarter)
        mbar_wait(accum_full, 0);
        tc_fence_after();
        constexpr int CPM = D / 32;                                // chunks per matrix
#pragma unroll 1
        for (int c = hf; c < 3 * CPM; c += 2) {
            const int mat = c / CPM, col0 = (c - mat * CPM) * 32;
            uint32_t v[32];
            tmem_ld32(tmem + lane_off + (uint32_t)(mat * D + col0), v);
            tmem_ld_wait();
            float f[32];
            const float post = (mat == 0) ? p.scale_log2e : 1.0f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int ci = mat * D + col0 + j;
                f[j] = fmaf(ln_a, __uint_as_float(v[j]), fmaf(ln_b, s_c1[ci], s_bias[ci])) * post;
            }
            if (mat < 2) {
                uint8_t* base = (mat == 0 ? sQ : sK) + (col0 >> 6) * CONV_A_BYTES + row * 128;
                const int ch0 = (col0 & 63) >> 3;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 o;
                    o.x = pack_h2(f[j * 8 + 0], f[j * 8 + 1]);
                    o.y = pack_h2(f[j * 8 + 2], f[j * 8 + 3]);
                    o.z = pack_h2(f[j * 8 + 4], f[j * 8 + 5]);
                    o.w = pack_h2(f[j * 8 + 6], f[j * 8 + 7]);
                    *reinterpret_cast<uint4*>(base + (((ch0 + j) ^ sw) << 4)) = o;
                }
            } else {
                // V^T: element (d, key = row) of atom row>>6: [d][128 B], 16-byte chunk ((row & 63) >> 3) ^ (d & 7)
                uint8_t* base = sVt + (row >> 6) * (D * 128) + (row & 7) * 2;
                const int kc = (row & 63) >> 3;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int d = col0 + j;
                    *reinterpret_cast<f16*>(base + d * 128 + ((kc ^ (d & 7)) << 4)) = __float2half_rn(fminf(fmaxf(f[j], -65504.f), 65504.f));
                }
            }
        }
        if (D < 64 && hf == 1) {   // zero the K padding (columns D..63) of this row in Q and K
#pragma unroll
            for (int j = D / 8; j < 8; ++j) {
                *reinterpret_cast<uint4*>(sQ + row * 128 + ((j ^ sw) << 4)) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(sK + row * 128 + ((j ^ sw) << 4)) = make_uint4(0, 0, 0, 0);
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(qk_ready);
        // ---- 3. block-diagonal softmax (one thread per row; the hf == 0 warps)
        if (hf == 0) {
            mbar_wait(s_full, 0);
            tc_fence_after();
            const int L = p.L;
            const int span = L > 32 ? 64 : 32;                       // columns this warp loads
            const int col_base = L > 32 ? (row / L) * L : q * 32;    // warp-uniform
            uint32_t v[2][32];
            tmem_ld32(tmem + lane_off + (uint32_t)(S_COL + col_base), v[0]);
            if (span == 64) tmem_ld32(tmem + lane_off + (uint32_t)(S_COL + col_base + 32), v[1]);
            tmem_ld_wait();
            const int own0 = (row / L) * L - col_base;               // first valid column (relative), L valid columns
            float mx = -INFINITY;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int cj = hh * 32 + j;
                    const bool ok = cj < span && cj >= own0 && cj < own0 + L;
                    if (ok) mx = fmaxf(mx, __uint_as_float(v[hh][j]));
                }
            float lsum = 0.f;
            uint32_t pk[2][16];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int c0 = hh * 32 + 2 * j, c1 = c0 + 1;
                    const bool ok0 = c0 < span && c0 >= own0 && c0 < own0 + L;
                    const bool ok1 = c1 < span && c1 >= own0 && c1 < own0 + L;
                    const float p0 = ok0 ? ex2_approx(__uint_as_float(v[hh][2 * j]) - mx) : 0.f;
                    const float p1 = ok1 ? ex2_approx(__uint_as_float(v[hh][2 * j + 1]) - mx) : 0.f;
                    lsum += p0 + p1;
                    pk[hh][j] = pack_h2_nosat(p0, p1);
                }
            s_l[row] = lsum;
            // P: 128 keys = 64 packed columns = 4 chunks of 16; zero outside this row's sample
            uint32_t zero[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) zero[j] = 0u;
            const int cb = col_base >> 5;                             // first loaded 32-key chunk
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const uint32_t addr = tmem + lane_off + (uint32_t)(P_COL + ch * 16);
                if (ch == cb) tmem_st16(addr, pk[0]);
                else if (span == 64 && ch == cb + 1) tmem_st16(addr, pk[1]);
                else tmem_st16(addr, zero);
            }
            tmem_st_wait();
            __threadfence_block();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready);
        }
        // ---- 4. O / l -> fp16 -> global
        mbar_wait(o_full, 0);
        tc_fence_after();
        constexpr int OCH = D / 32;                                 // 32-column chunks of O
        const float inv = 1.0f / s_l[row];
        const int grow = m0 + row;
#pragma unroll 1
        for (int c = hf; c < OCH; c += 2) {
            uint32_t v[32];
            tmem_ld32(tmem + lane_off + (uint32_t)(O_COL + c * 32), v);
            tmem_ld_wait();
            if (p.fuse_out == 2) {
                // keep this chunk (fp16) in registers: it is scattered to every head's CTA after the cluster barrier
                const int kc = c >> 1;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t pk = pack_h2(__uint_as_float(v[2 * j]) * inv, __uint_as_float(v[2 * j + 1]) * inv);
                    if (kc == 0) okeep[0][j] = pk; else okeep[1][j] = pk;
                }
            } else if (p.fuse_out) {
                // O_h -> A operand (K-major, 128B swizzle) over the Q tile; padding columns keep Q's zeros
                uint8_t* base = sQ + ((c * 32) >> 6) * CONV_A_BYTES + row * 128;
                const int ch0 = ((c * 32) & 63) >> 3;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 o;
                    o.x = pack_h2(__uint_as_float(v[j * 8 + 0]) * inv, __uint_as_float(v[j * 8 + 1]) * inv);
                    o.y = pack_h2(__uint_as_float(v[j * 8 + 2]) * inv, __uint_as_float(v[j * 8 + 3]) * inv);
                    o.z = pack_h2(__uint_as_float(v[j * 8 + 4]) * inv, __uint_as_float(v[j * 8 + 5]) * inv);
                    o.w = pack_h2(__uint_as_float(v[j * 8 + 6]) * inv, __uint_as_float(v[j * 8 + 7]) * inv);
                    *reinterpret_cast<uint4*>(base + (((ch0 + j) ^ sw) << 4)) = o;
                }
            } else if (grow < p.M) {
                uint4* op = reinterpret_cast<uint4*>(p.out + (size_t)grow * p.C + head * D + c * 32);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 o;
                    o.x = pack_h2(__uint_as_float(v[j * 8 + 0]) * inv, __uint_as_float(v[j * 8 + 1]) * inv);
                    o.y = pack_h2(__uint_as_float(v[j * 8 + 2]) * inv, __uint_as_float(v[j * 8 + 3]) * inv);
                    o.z = pack_h2(__uint_as_float(v[j * 8 + 4]) * inv, __uint_as_float(v[j * 8 + 5]) * inv);
                    o.w = pack_h2(__uint_as_float(v[j * 8 + 6]) * inv, __uint_as_float(v[j * 8 + 7]) * inv);
                    op[j] = o;
                }
            }
        }
        if (p.fuse_out == 2) {
            for (int i = threadIdx.x - 64; i < D; i += 256) s_ob[i] = __ldg(p.out_bias + head * D + i);
            tc_fence_before();
        }
        if (p.fuse_out == 1) {
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_ready);
            // ---- 5. this head's partial of the out-projection -> L2 workspace
            mbar_wait(out_full, 0);
            tc_fence_after();
            float* wrow = p.ws + (((size_t)blockIdx.x * gridDim.y + head) * 128 + row) * p.C;
            const int nch = p.C >> 5;
#pragma unroll 1
            for (int c = hf; c < nch; c += 2) {
                uint32_t v[32];
                tmem_ld32(tmem + lane_off + (uint32_t)(c * 32), v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    __stcg(reinterpret_cast<float4*>(wrow + c * 32 + j),
                           make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                       __uint_as_float(v[j + 3])));
            }
        }
    }
    if (p.fuse_out == 2) {
        const int heads = gridDim.y;
        // (1) every CTA of the cluster is done with its attention operands -> their memory may be overwritten by the peers
        cluster_arrive_release();
        cluster_wait_acquire();
        if (warp >= 2) {
            const int q = warp & 3, hf = (warp - 2) >> 2;
            const int row = q * 32 + lane, sw = row & 7;
            constexpr int OCH = D / 32;
#pragma unroll
            for (int kc = 0; kc < 2; ++kc) {
                const int c = hf + 2 * kc;
                if (c < OCH) {
                    const int col = head * D + c * 32;                       // first column of this chunk in the gathered [128][C]
                    const uint32_t local = smem_u32(sG + (col >> 6) * CONV_A_BYTES + row * 128);
                    const int ch0 = (col & 63) >> 3;
                    for (int peer = 0; peer < heads; ++peer) {
                        uint32_t remote;
                        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(peer));
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t a = remote + (uint32_t)(((ch0 + j) ^ sw) << 4);
                            asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(okeep[kc][4 * j]),
                                         "r"(okeep[kc][4 * j + 1]), "r"(okeep[kc][4 * j + 2]), "r"(okeep[kc][4 * j + 3])
                                         : "memory");
                        }
                    }
                }
            }
        }
        fence_proxy_async();            // writer side: generic-proxy (DSMEM) stores before the peers' tensor-core reads
        // (2) all scatters have landed
        cluster_arrive_release();
        cluster_wait_acquire();
        fence_proxy_async();
        tc_fence_after();
        if (warp == 0) {
            if (lane == 0) {
                for (int kb = NS2; kb < nkb; ++kb) {
                    const int st = kb % NS2;
                    mbar_wait(&w2_empty[st], ((kb / NS2) & 1) ^ 1);
                    mbar_arrive_expect_tx(&w2_full[st], (uint32_t)(D * 128));
                    tma_load_2d(sW2 + st * (D * 128), &tmW, &w2_full[st], kb * 64, head * D);
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            if (lane == 0) {
                constexpr uint32_t idesc_w = umma_idesc_f16(128, D);
                for (int kb = 0; kb < nkb; ++kb) {
                    const int st = kb % NS2;
                    mbar_wait(&w2_full[st], (kb / NS2) & 1);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(smem_u32(sG + kb * CONV_A_BYTES));
                    const uint64_t dw = umma_desc_sw128(smem_u32(sW2 + st * (D * 128)));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16(tmem, da + (uint64_t)(k * 2), dw + (uint64_t)(k * 2), idesc_w, (kb | k) != 0);
                    umma_commit(&w2_empty[st]);
                }
                umma_commit(out2_full);
            }
            __syncwarp();
        } else {
            const int q = warp & 3, hf = (warp - 2) >> 2;
            const int row = q * 32 + lane;
            const uint32_t lane_off = (uint32_t)(q * 32) << 16;
            const int grow = m0 + row;
            const int growc = grow < p.M ? grow : p.M - 1;
            constexpr int OCH = D / 32;
            // residual rows are fetched before waiting for the MMA
            uint4 res[2][4];
#pragma unroll
            for (int kc = 0; kc < 2; ++kc) {
                const int c = hf + 2 * kc;
                if (c < OCH) {
                    const uint4* rp = reinterpret_cast<const uint4*>(p.residual + (size_t)growc * p.C + head * D + c * 32);
#pragma unroll
                    for (int j = 0; j < 4; ++j) res[kc][j] = __ldg(rp + j);
                }
            }
            mbar_wait(out2_full, 0);
            tc_fence_after();
#pragma unroll
            for (int kc = 0; kc < 2; ++kc) {
                const int c = hf + 2 * kc;
                if (c < OCH) {
                    uint32_t v[32];
                    tmem_ld32(tmem + lane_off + (uint32_t)(c * 32), v);
                    tmem_ld_wait();
                    if (grow < p.M) {
                        uint4* op = reinterpret_cast<uint4*>(p.out_final + (size_t)grow * p.C + head * D + c * 32);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float f[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j * 8 + e]) + s_ob[c * 32 + j * 8 + e];
                            const uint4 r4 = res[kc][j];
                            float2 t;
                            t = unpack_h2(r4.x); f[0] += t.x; f[1] += t.y;
                            t = unpack_h2(r4.y); f[2] += t.x; f[3] += t.y;
                            t = unpack_h2(r4.z); f[4] += t.x; f[5] += t.y;
                            t = unpack_h2(r4.w); f[6] += t.x; f[7] += t.y;
                            if (p.final_act == 1) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
                            }
                            uint4 o;
                            o.x = pack_h2(f[0], f[1]); o.y = pack_h2(f[2], f[3]);
                            o.z = pack_h2(f[4], f[5]); o.w = pack_h2(f[6], f[7]);
                            op[j] = o;
                        }
                    }
                }
            }
        }
    }
    if (p.fuse_out == 1) {
        // every head of the tile has written its partial; CTA `head` finalises 128/heads rows in fixed head order
        // barrier.cluster arrive.release / wait.acquire order the partial-tile writes (st.global.cg) before the peers' ld.global.cg
        cluster_arrive_release();
        cluster_wait_acquire();
        if (warp >= 2) {
            const int et = threadIdx.x - 64;                        // 0..255
            const int heads = gridDim.y;
            const int rows_per = 128 / heads;
            const int c8n = p.C >> 3;
            const float* wtile = p.ws + (size_t)blockIdx.x * heads * 128 * p.C;
            for (int it = et; it < rows_per * c8n; it += 256) {
                const int row = head * rows_per + it / c8n;
                const int c8 = (it % c8n) * 8;
                const int grow = m0 + row;
                if (grow >= p.M) continue;
                float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                float4 pa[8], pb[8];                                  // all heads' partials in flight at once
#pragma unroll
                for (int hh = 0; hh < 8; ++hh) {
                    if (hh < heads) {
                        const float4* src = reinterpret_cast<const float4*>(wtile + ((size_t)hh * 128 + row) * p.C + c8);
                        pa[hh] = __ldcg(src);
                        pb[hh] = __ldcg(src + 1);
                    }
                }
#pragma unroll
                for (int hh = 0; hh < 8; ++hh) {
                    if (hh < heads) {
                        f[0] += pa[hh].x; f[1] += pa[hh].y; f[2] += pa[hh].z; f[3] += pa[hh].w;
                        f[4] += pb[hh].x; f[5] += pb[hh].y; f[6] += pb[hh].z; f[7] += pb[hh].w;
                    }
                }
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.out_bias + c8));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.out_bias + c8 + 4));
                f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                const uint4 r4 = *reinterpret_cast<const uint4*>(p.residual + (size_t)grow * p.C + c8);
                float2 t;
                t = unpack_h2(r4.x); f[0] += t.x; f[1] += t.y;
                t = unpack_h2(r4.y); f[2] += t.x; f[3] += t.y;
                t = unpack_h2(r4.z); f[4] += t.x; f[5] += t.y;
                t = unpack_h2(r4.w); f[6] += t.x; f[7] += t.y;
                if (p.final_act == 1) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
                }
                uint4 o;
                o.x = pack_h2(f[0], f[1]); o.y = pack_h2(f[2], f[3]);
                o.z = pack_h2(f[4], f[5]); o.w = pack_h2(f[6], f[7]);
                *reinterpret_cast<uint4*>(p.out_final + (size_t)grow * p.C + c8) = o;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

struct AttnBlockPlan {
    CUtensorMap tmA, tmB, tmW;
    AttnBlockParams p;
    int D = 0, heads = 0;
};

inline bool attn_block_supported(int L, int C, int heads) {
    if (heads <= 0 || C % heads != 0 || C % 64 != 0) return false;
    const int D = C / heads;
    if (D != 32 && D != 64 && D != 128) return false;
    return L >= 1 && L <= 64 && (128 % L) == 0;
}

// x: [M][C] fp16 activations (raw, un-normalised); w: gamma-folded in_proj weight [3C][C] fp16.
inline int attn_block_plan_build(AttnBlockPlan& pl, const f16* x, const f16* w, const float* c1, const float* bias, f16* out, int M,
                                 int C, int L, int heads) {
    pl.D = C / heads;
    pl.heads = heads;
    pl.p.c1 = c1; pl.p.bias = bias; pl.p.out = out; pl.p.M = M; pl.p.C = C; pl.p.L = L;
    pl.p.scale_log2e = 1.4426950408889634f / sqrtf((float)pl.D);
    uint64_t ad[2] = {(uint64_t)C, (uint64_t)M};
    uint64_t as[1] = {(uint64_t)C * 2};
    uint32_t ab[2] = {64, 128};
    B2D_TRY(make_tmap_f16(&pl.tmA, x, 2, ad, as, ab));
    uint64_t bd[3] = {(uint64_t)C, (uint64_t)C, 3};
    uint64_t bs[2] = {(uint64_t)C * 2, (uint64_t)C * C * 2};
    uint32_t bb[3] = {64, (uint32_t)pl.D, 3};
    B2D_TRY(make_tmap_f16(&pl.tmB, w, 3, bd, bs, bb));
    pl.tmW = pl.tmA;
    pl.p.fuse_out = 0;
    return 0;
}

// Adds the fused out-projection: wo = out_proj.weight [C][C] fp16 (K-major), ws = fp32 workspace of attn_block_ws_floats().
inline size_t attn_block_ws_floats(int M, int C, int heads) { return (size_t)((M + 127) / 128) * heads * 128 * C; }
inline bool attn_block_out_supported(int C, int heads) { return heads <= 8 && (128 % heads) == 0 && C <= 512 && (C <= 256 || C == 512); }
// mode 1: per-head fp32 partials through the L2 workspace `ws`; mode 2: fp16 head outputs gathered through DSMEM (ws unused).
inline bool attn_block_gather_supported(int C, int heads) {
    if (heads > 8 || C % heads != 0) return false;
    const int D = C / heads;
    const int ring = (D == 128 ? 3 : 4) * (CONV_A_BYTES + 3 * D * 128);
    const int dpad = D < 64 ? 64 : D;
    const int oper = 2 * (dpad / 64) * CONV_A_BYTES + 2 * D * 128;
    const int g = 128 * C * 2;
    const int ns = (C / 64) < 4 ? (C / 64) : 4;
    return (g > oper ? g : oper) + ns * D * 128 <= ring;
}
inline int attn_block_plan_fuse_out(AttnBlockPlan& pl, const f16* wo, const float* out_bias, const f16* residual, f16* out_final,
                                    float* ws, int final_act, int mode = 1) {
    const int C = pl.p.C;
    uint64_t wd[2] = {(uint64_t)C, (uint64_t)C};
    uint64_t wsb[1] = {(uint64_t)C * 2};
    uint32_t wb[2] = {64, (uint32_t)(mode == 2 ? pl.D : (C > 256 ? 256 : C))};
    B2D_TRY(make_tmap_f16(&pl.tmW, wo, 2, wd, wsb, wb));
    pl.p.fuse_out = mode;
    pl.p.out_bias = out_bias; pl.p.residual = residual; pl.p.out_final = out_final; pl.p.ws = ws; pl.p.final_act = final_act;
    return 0;
}

inline int attn_block_init_attrs() {
    B2D_CUDA(cudaFuncSetAttribute(attn_block_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ab_smem_bytes<32>()));
    B2D_CUDA(cudaFuncSetAttribute(attn_block_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ab_smem_bytes<64>()));
    B2D_CUDA(cudaFuncSetAttribute(attn_block_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, ab_smem_bytes<128>()));
    return 0;
}

template <int D>
inline int attn_block_launch_t(const AttnBlockPlan& pl, cudaStream_t st) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((pl.p.M + 127) / 128, pl.heads);
    cfg.blockDim = dim3(AB_THREADS);
    cfg.dynamicSmemBytes = ab_smem_bytes<D>();
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = pl.p.fuse_out ? pl.heads : 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = g_pdl_enabled;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    B2D_CUDA(cudaLaunchKernelEx(&cfg, attn_block_kernel<D>, pl.tmA, pl.tmB, pl.tmW, pl.p));
    return 0;
}

inline int attn_block_launch(const AttnBlockPlan& pl, cudaStream_t st) {
    if (pl.D == 32) return attn_block_launch_t<32>(pl, st);
    if (pl.D == 64) return attn_block_launch_t<64>(pl, st);
    return attn_block_launch_t<128>(pl, st);
}

}  // namespace b2d
