// Persistent streaming GEMM for the 1x1 projections with shallow K (K = Cin <= 128: QKV / out / FF projections of the
// C = 64 and C = 128 attention blocks, ConvTranspose GEMMs of the last decoder levels).  These have M = B*H*W up to 10^5-10^6
// rows and are bound by activation traffic, not by the tensor pipe, so the kernel is organised around keeping loads and
// stores in flight:
//   * one CTA per SM, persistent over M tiles (stride gridDim.x); blockIdx.y picks the N block (<= 256 columns) whose
//     weights are TMA-loaded ONCE and stay resident in shared memory;
//   * A tiles (128 rows x K) stream through a 4-6 stage TMA ring;
//   * two TMEM accumulator stages: the MMA of tile i+1 is issued while the epilogue warps drain tile i
//     (tcgen05.ld -> bias/residual/activation -> fp16 stores);
//   * optional fused LayerNorm (ConvParams::ln_c1): the epilogue thread of row m reads that row of the A tile from shared
//     memory (conflict-free in swizzle order), derives mean/rstd, and applies them algebraically to the accumulator, so the
//     separate LayerNorm launch and its normalised copy of the activations disappear.
// Same operand layouts, descriptors and epilogue semantics as conv_tc_kernel (conv.cuh).
#pragma once
#include "common.cuh"
#include "conv.cuh"

namespace b2d {

constexpr int GS_EPI_WARPS = 16;
constexpr int GS_EPI_THREADS = GS_EPI_WARPS * 32;
constexpr int GS_THREADS = 64 + GS_EPI_THREADS;   // warp 0: TMA producer, warp 1: MMA issuer, warps 2..17: epilogue

template <int KB>
__host__ __device__ constexpr int gs_stages() { return KB == 1 ? 4 : 2; }
template <int KB>
__host__ __device__ constexpr int gs_out_bufs() { return KB == 1 ? 4 : 2; }   // output staging tiles (TMA stores in flight)
template <int KB>
__host__ __device__ constexpr int gs_smem_bytes() {
    return gs_stages<KB>() * KB * CONV_A_BYTES + KB * 256 * 128 + (gs_out_bufs<KB>() + 2) * CONV_A_BYTES /*store staging + residual ring*/ +
           1024 + 256;
}

template <int KB>
__global__ void __launch_bounds__(GS_THREADS, 1)
    gemm_stream_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, const ConvParams p,
                       int M, int NB) {
    pdl_launch_dependents();
    constexpr int STAGES = gs_stages<KB>();
    extern __shared__ uint8_t gs_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gs_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                                        // STAGES x KB x 16 KB
    uint8_t* sB = smem + STAGES * KB * CONV_A_BYTES;           // KB x NB x 128 B (resident weights)
    constexpr int OB = gs_out_bufs<KB>();
    constexpr int OBP = OB / 2;                                // staging tiles per epilogue pair
    uint8_t* sO = sB + KB * 256 * 128;                         // OB x 16 KB output staging (128 rows x 64 ch, 128B swizzle)
    uint8_t* sR = sO + OB * CONV_A_BYTES;                      // 2 x 16 KB residual blocks, TMA-prefetched by the producer
    uint64_t* bars = reinterpret_cast<uint64_t*>(sR + 2 * CONV_A_BYTES);
    uint64_t* a_full = bars;                 // [STAGES]
    uint64_t* a_empty = bars + STAGES;       // [STAGES]
    uint64_t* t_full = bars + 2 * STAGES;    // [2] accumulator ready
    uint64_t* t_empty = t_full + 2;          // [2] accumulator drained (epilogue warps)
    uint64_t* b_full = t_empty + 2;
    uint64_t* r_full = b_full + 1;           // [2]
    uint64_t* r_empty = r_full + 2;          // [2] (epilogue warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nblk = blockIdx.y;
    const int num_tiles = (M + 127) / 128;
    __shared__ __align__(16) float2 s_stat[2 * 512];          // per-group partial (sum, sum of squares) of the LN statistics, double-buffered
    __shared__ __align__(16) float s_bias[256], s_c1[256];   // bias / LN column sums of this CTA's N block (persistent => loaded once)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < STAGES; ++i) {
                mbar_init(&a_full[i], 1);
                mbar_init(&a_empty[i], p.ln_c1 ? 1 + GS_EPI_WARPS : 1);   // + the epilogue warps when they read A for the LN statistics
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&t_full[i], 1);
                mbar_init(&t_empty[i], GS_EPI_WARPS);
            }
            mbar_init(b_full, 1);
            for (int i = 0; i < 2; ++i) {
                mbar_init(&r_full[i], 1);
                mbar_init(&r_empty[i], GS_EPI_WARPS / 2);
            }
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

    // Producer and MMA loops run converged on all 32 lanes, one hardware-elected lane issues (conv.cuh explains why: no
    // ELECT / R2UR waterfall around every TMA and tcgen05 instruction).
    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(b_full, (uint32_t)(KB * NB * 128));
            for (int kb = 0; kb < KB; ++kb) tma_load_2d(sB + kb * NB * 128, &tmB, b_full, kb * 64, nblk * NB);
        }
        __syncwarp();
        int it = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
            const int st = it % STAGES;
            mbar_wait(&a_empty[st], ((it / STAGES) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&a_full[st], KB * CONV_A_BYTES);
                for (int kb = 0; kb < KB; ++kb)
                    tma_load_2d(sA + (st * KB + kb) * CONV_A_BYTES, &tmA, &a_full[st], kb * 64, t * 128);
            }
            __syncwarp();
            if (p.residual != nullptr) {
                for (int jb = 0; jb < NB / 64; ++jb) {
                    const int blk = it * (NB / 64) + jb, rb = blk & 1;
                    mbar_wait(&r_empty[rb], ((blk >> 1) & 1) ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&r_full[rb], CONV_A_BYTES);
                        tma_load_2d(sR + rb * CONV_A_BYTES, &tmR, &r_full[rb], nblk * NB + jb * 64, t * 128);
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        {
            const uint32_t idesc = umma_idesc_f16(128, NB);
            mbar_wait(b_full, 0);
            int it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
                const int st = it % STAGES, acc = it & 1;
                mbar_wait(&t_empty[acc], ((it >> 1) & 1) ^ 1);
                mbar_wait(&a_full[st], (it / STAGES) & 1);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb) {
                        const uint64_t da = umma_desc_sw128(smem_u32(sA + (st * KB + kb) * CONV_A_BYTES));
                        const uint64_t db = umma_desc_sw128(smem_u32(sB + kb * NB * 128));
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16(tmem + acc * 256, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    }
                    umma_commit(&a_empty[st]);
                    umma_commit(&t_full[acc]);
                }
                __syncwarp();
            }
        }
        __syncwarp();
    } else {
        // 16 epilogue warps = 4 groups of 128 threads (one TMEM lane quarter per warp, warp % 4).  Groups 2p and 2p+1 form
        // pair p: the pair owns every second 64-column output block (running block counter & 1), its two groups take the
        // low / high 32 columns of it.  One warp per scheduler cannot hide the tcgen05.ld -> FMA -> pack -> st.shared
        // dependency chain (measured 5 clk per issued instruction); four per scheduler can.
        const int q = warp & 3;
        const int grp = (warp - 2) >> 2, pair = grp >> 1, half = grp & 1;
        const int row = q * 32 + lane;
        const int HW = p.Ho * p.Wo;
        const int sw = row & 7;
        // The two warps that share a TMEM lane quarter inside a pair (low / high 32 columns of the pair's 64-column blocks) form a
        // sub-group with its own 32-row staging tiles, its own 64-thread named barrier and its own TMA stores: no barrier in
        // the block loop is wider than two warps.
        const int sg = pair * 4 + q;                                     // 0..7
        const bool elected = half == 0 && lane == 0;                    // issues this sub-group's TMA stores
        const int bar_id = 1 + sg;                                       // named barriers 1..8 (9..12: statistics per quarter, 13: init)
        // bias / LN column sums of this CTA's N block: fetched by the epilogue warps only, so the TMA and MMA warps do not wait
        // for this L2 round trip at the prologue's __syncthreads
        for (int i = threadIdx.x - 64; i < NB; i += GS_EPI_THREADS) {
            const int col = nblk * NB + i;
            const int cb = p.convt ? col % p.CoutT : col;
            s_bias[i] = p.bias ? __ldg(p.bias + cb) : 0.f;
            s_c1[i] = p.ln_c1 ? __ldg(p.ln_c1 + cb) : 0.f;
        }
        named_bar_sync(13, GS_EPI_THREADS);
        int it = 0, blk = 0;                                   // blk: running 64-column block counter
        float satm = 0.f;                                      // running |value| maximum of this thread's conversions (sat_flush per tile)
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            const int m0 = t * 128;
            const int m = m0 + row;
            const int mr = m < M ? m : M - 1;                 // rows past the end are clipped by the TMA store
            const int n = mr / HW;
            float ln_a = 1.f, ln_b = 0.f;                     // out = ln_a * acc + (ln_b * c1 + bias)
            if (p.ln_c1 != nullptr) {
                const int st = it % STAGES;
                mbar_wait(&a_full[st], (it / STAGES) & 1);
                float s1 = 0.f, s2 = 0.f;
                // each group sums a quarter of the row (swizzle order: conflict-free, and a sum does not care)
                constexpr int PER = KB * 2;
#pragma unroll
                for (int jj = 0; jj < PER; ++jj) {
                    const int j = grp * PER + jj;               // 16-byte chunk index within the K = KB*64 row
                    const uint4* arow = reinterpret_cast<const uint4*>(sA + (st * KB + (j >> 3)) * CONV_A_BYTES + row * 128);
                    const uint4 a4 = arow[(j & 7) ^ sw];
                    float2 tt;
                    tt = unpack_h2(a4.x); s1 += tt.x + tt.y; s2 = fmaf(tt.x, tt.x, s2); s2 = fmaf(tt.y, tt.y, s2);
                    tt = unpack_h2(a4.y); s1 += tt.x + tt.y; s2 = fmaf(tt.x, tt.x, s2); s2 = fmaf(tt.y, tt.y, s2);
                    tt = unpack_h2(a4.z); s1 += tt.x + tt.y; s2 = fmaf(tt.x, tt.x, s2); s2 = fmaf(tt.y, tt.y, s2);
                    tt = unpack_h2(a4.w); s1 += tt.x + tt.y; s2 = fmaf(tt.x, tt.x, s2); s2 = fmaf(tt.y, tt.y, s2);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_empty[st]);
                float2* sst = s_stat + (it & 1) * 512;
                sst[grp * 128 + row] = make_float2(s1, s2);
                named_bar_sync(9 + q, 128);                    // the four warps that own this lane quarter
                s1 = 0.f; s2 = 0.f;
#pragma unroll
                for (int g = 0; g < 4; ++g) {                   // fixed order: every group derives bit-identical statistics
                    const float2 v2 = sst[g * 128 + row];
                    s1 += v2.x; s2 += v2.y;
                }
                const float invk = 1.0f / (float)(KB * 64);
                const float mu = s1 * invk;
                ln_a = rsqrtf(fmaxf(s2 * invk - mu * mu, 0.f) + 1e-5f);
                ln_b = -ln_a * mu;
            }
            mbar_wait(&t_full[acc], (it >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int jb = 0; jb < NB / 64; ++jb, ++blk) {
                if ((blk & 1) != pair) continue;
                uint8_t* stg = sO + (sg * OBP + ((blk >> 1) % OBP)) * (CONV_A_BYTES / 4);   // [32 rows][128 B], 128B swizzle
                uint4* srow = reinterpret_cast<uint4*>(stg + lane * 128);
                if (elected) tma_store_wait_read_n<OBP - 1>();   // the store that used this buffer OBP blocks ago has read it
                named_bar_sync(bar_id, 64);
                const int col = nblk * NB + jb * 64;           // GEMM column of this block
                int cbase = col, ab = 0;
                if (p.convt) {
                    ab = col / p.CoutT;
                    cbase = col - ab * p.CoutT;
                }
                {
                    uint32_t v[32];
                    tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + jb * 64 + half * 32), v);
                    tmem_ld_wait();
                    const int c0 = cbase + half * 32;
                    float f[32];
                    const int lc = jb * 64 + half * 32;             // column within this CTA's N block
                    if (p.ln_c1) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 c4 = *reinterpret_cast<const float4*>(&s_c1[lc + j]);
                            const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[lc + j]);
                            f[j] = fmaf(ln_a, __uint_as_float(v[j]), fmaf(ln_b, c4.x, b4.x));
                            f[j + 1] = fmaf(ln_a, __uint_as_float(v[j + 1]), fmaf(ln_b, c4.y, b4.y));
                            f[j + 2] = fmaf(ln_a, __uint_as_float(v[j + 2]), fmaf(ln_b, c4.z, b4.z));
                            f[j + 3] = fmaf(ln_a, __uint_as_float(v[j + 3]), fmaf(ln_b, c4.w, b4.w));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[lc + j]);
                            f[j] = __uint_as_float(v[j]) + b4.x;
                            f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
                            f[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
                            f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
                        }
                    }
                    if (p.residual) {   // non-convT only (checked on the host): residual block TMA-prefetched into sR[pair]
                        mbar_wait(&r_full[pair], (blk >> 1) & 1);
                        const uint4* rrow = reinterpret_cast<const uint4*>(sR + pair * CONV_A_BYTES + row * 128);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 r4 = rrow[(half * 4 + j) ^ sw];
                            float2 tt;
                            tt = unpack_h2(r4.x); f[j * 8 + 0] += tt.x; f[j * 8 + 1] += tt.y;
                            tt = unpack_h2(r4.y); f[j * 8 + 2] += tt.x; f[j * 8 + 3] += tt.y;
                            tt = unpack_h2(r4.z); f[j * 8 + 4] += tt.x; f[j * 8 + 5] += tt.y;
                            tt = unpack_h2(r4.w); f[j * 8 + 6] += tt.x; f[j * 8 + 7] += tt.y;
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&r_empty[pair]);
                    }
                    if (p.act) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], p.act);
                    }
                    if (p.post_add) {
                        const float* pa = p.post_add + (size_t)n * p.post_stride + c0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(pa + j));
                            f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 o;
                        o.x = pack_h2_acc(f[j * 8 + 0], f[j * 8 + 1], satm);
                        o.y = pack_h2_acc(f[j * 8 + 2], f[j * 8 + 3], satm);
                        o.z = pack_h2_acc(f[j * 8 + 4], f[j * 8 + 5], satm);
                        o.w = pack_h2_acc(f[j * 8 + 6], f[j * 8 + 7], satm);
                        srow[(half * 4 + j) ^ sw] = o;
                    }
                }
                fence_proxy_async();
                named_bar_sync(bar_id, 64);
                if (elected) {
                    const int ms = m0 + q * 32;                 // first GEMM row of this sub-group's 32-row slice
                    if (p.convt) {
                        const int n0 = ms / HW, rem = ms - n0 * HW;
                        tma_store_5d(&tmO, stg, (ab & 1) * p.CoutT + cbase, rem % p.Wo, ab >> 1, rem / p.Wo, n0);
                    } else {
                        tma_store_2d(&tmO, stg, col, ms);
                    }
                    tma_store_commit();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[acc]);
            sat_flush(satm);
        }
        if (elected) tma_store_wait_read();   // smem must outlive the bulk reads; the writes complete with the grid
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

struct GemmStreamPlan {
    ConvParams p;
    CUtensorMap tmA, tmB, tmO, tmR;
    int KB = 1, NB = 64, M = 0;
    dim3 grid;
};

// Eligible: 1x1, stride 1 (incl. the ConvTranspose GEMM), K in {64, 128}, N a multiple of one of {256,192,128,64}, big M.
constexpr int GS_MIN_ROWS = 8192;
inline bool gemm_stream_supported(const ConvParams& p, int min_rows = GS_MIN_ROWS) {
    if (p.R != 1 || p.S != 1 || p.stride != 1 || p.pad != 0) return false;
    if (p.Cin != 64 && p.Cin != 128) return false;
    if (p.convt && p.residual) return false;
    if (p.convt && !(is_pow2(p.Ho) && is_pow2(p.Wo) && p.Wo <= 128)) return false;
    const long long M = (long long)p.B * p.Ho * p.Wo;
    return M >= min_rows && p.Cout % 64 == 0;
}

inline int gemm_stream_plan_build(GemmStreamPlan& pl, int num_sms) {
    ConvParams& p = pl.p;
    pl.KB = p.Cin / 64;
    pl.M = p.B * p.Ho * p.Wo;
    int NB = 64;
    const int cands[4] = {256, 192, 128, 64};
    for (int c : cands)
        if (p.Cout % c == 0 && (!p.convt || p.CoutT % 32 == 0)) { NB = c; break; }
    pl.NB = NB;
    const int nblocks = p.Cout / NB;
    const int tiles = (pl.M + 127) / 128;
    int gx = num_sms / nblocks;
    if (gx < 1) gx = 1;
    if (gx > tiles) gx = tiles;
    pl.grid = dim3(gx, nblocks, 1);
    p.splits = 1;
    uint64_t ad[2] = {(uint64_t)p.Cin, (uint64_t)pl.M};
    uint64_t as[1] = {(uint64_t)p.Cin * 2};
    uint32_t ab[2] = {64, 128};
    B2D_TRY(make_tmap_f16(&pl.tmA, p.in, 2, ad, as, ab));
    uint64_t bd[2] = {(uint64_t)p.Cin, (uint64_t)p.Cout};
    uint64_t bs[1] = {(uint64_t)p.Cin * 2};
    uint32_t bb[2] = {64, (uint32_t)NB};
    B2D_TRY(make_tmap_f16(&pl.tmB, p.w, 2, bd, bs, bb));
    if (p.convt) {   // stored tensor [B][2Ho][2Wo][Co] viewed as [n][h][a][w][(b,c)]; 128 consecutive GEMM rows = a (TW,TH,TN) box
        const int TW = p.Wo < 128 ? p.Wo : 128;
        const int TH = (128 / TW) < p.Ho ? (128 / TW) : p.Ho;
        const int TN = 128 / (TW * TH);
        const uint64_t Co = p.CoutT, Wo = p.Wo, Ho = p.Ho;
        uint64_t dims[5] = {2 * Co, Wo, 2, Ho, (uint64_t)p.B};
        uint64_t str[4] = {2 * Co * 2, 2 * Wo * Co * 2, 4 * Wo * Co * 2, 4 * Ho * Wo * Co * 2};
        (void)TN;
        const int tw = TW < 32 ? TW : 32;                       // 32 consecutive GEMM rows = a (tw, th, tn) box
        const int th = (32 / tw) < TH ? (32 / tw) : TH;
        const int tn = 32 / (tw * th);
        uint32_t box[5] = {64, (uint32_t)tw, 1, (uint32_t)th, (uint32_t)tn};
        B2D_TRY(make_tmap_f16(&pl.tmO, p.out, 5, dims, str, box));
    } else {
        uint64_t od[2] = {(uint64_t)p.Cout, (uint64_t)pl.M};
        uint64_t os[1] = {(uint64_t)p.Cout * 2};
        uint32_t ob[2] = {64, 128}, ob32[2] = {64, 32};
        B2D_TRY(make_tmap_f16(&pl.tmO, p.out, 2, od, os, ob32));      // stores: 32-row slices (one per epilogue sub-group)
        if (p.residual) B2D_TRY(make_tmap_f16(&pl.tmR, p.residual, 2, od, os, ob));
    }
    if (!p.residual) pl.tmR = pl.tmO;
    return 0;
}

inline int gemm_stream_init_attrs() {
    B2D_CUDA(cudaFuncSetAttribute(gemm_stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, gs_smem_bytes<1>()));
    B2D_CUDA(cudaFuncSetAttribute(gemm_stream_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, gs_smem_bytes<2>()));
    return 0;
}

inline int gemm_stream_launch(const GemmStreamPlan& pl, cudaStream_t st) {
    if (pl.KB == 1)
        B2D_CUDA(launch_k(gemm_stream_kernel<1>, pl.grid, dim3(GS_THREADS), gs_smem_bytes<1>(), st, pl.tmA, pl.tmB, pl.tmO, pl.tmR, pl.p, pl.M, pl.NB));
    else
        B2D_CUDA(launch_k(gemm_stream_kernel<2>, pl.grid, dim3(GS_THREADS), gs_smem_bytes<2>(), st, pl.tmA, pl.tmB, pl.tmO, pl.tmR, pl.p, pl.M, pl.NB));
    return 0;
}

}  // namespace b2d
