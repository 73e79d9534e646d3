// Persistent implicit-GEMM convolution for the large layers (enough 128 x BN tiles to cover the SMs at least twice, no split-K).
//
// conv_tc_kernel (conv.cuh) computes ONE tile per CTA: barrier init, TMEM allocation, pipeline fill, K loop, epilogue and the
// store drain are paid per tile and nothing overlaps the epilogue — measured on the 64 -> 64 channel 3x3 layers of the 128x128
// network at batch 256 (2048 tiles of nine k-blocks): 65 us for 0.6 us of tensor time and 10 us of HBM time per SM.
// Here one CTA per SM walks a static list of tiles with
//   * the TMA ring running across tile boundaries (the producer never drains),
//   * TWO TMEM accumulator stages: the MMAs of tile i+1 run while eight epilogue warps drain tile i,
//   * dedicated staging / residual tiles (the one-tile kernel borrows drained ring stages for them),
//   * converged producer / issuer loops with elect.sync issue (no R2UR waterfall).
// Same operand layouts, tensor maps, epilogue semantics (bias, residual, activation, time projection, GroupNorm partial
// sums, ConvTranspose pixel-shuffle store) and the same ConvPlan as conv_tc_kernel; the stored tensor is bit-identical to it
// (the GroupNorm partial sums are added in a different, equally fixed order).
//   warp 0: TMA producer   warp 1: MMA issuer + TMEM owner   warps 2-9: epilogue (warp % 4 = TMEM lane quarter; group (warp-2)/4
//   takes the low / high 32 columns of a 64-channel block for BN = 64, the first / second 64-channel block for BN = 128)
#pragma once
#include "conv.cuh"

namespace b2d {

constexpr int CONVP_EPI_THREADS = 256;                 // per epilogue group
// EG = 2: TWO epilogue groups of eight warps, group e drains the tiles with (tile counter & 1) == e, i.e. it owns accumulator stage
// e, with its own staging / residual tiles, bias slice and named barrier.  For BN = 64 the epilogue of a tile (tcgen05.ld ->
// bias / residual / activation / vector -> pack -> staging -> TMA store, ~600 dependent instructions per thread with two warps per
// scheduler) takes ~5000 clk against ~2100-2500 clk of shared-memory fill per tile: ncu showed the 64 -> 64 3x3 layers at 27 % tensor
// pipe with the tile rate set by the epilogue.  Two groups drain two tiles at a time.
template <int EG>
__host__ __device__ constexpr int convp_threads() { return 64 + EG * CONVP_EPI_THREADS; }

// SLAB mode (3x3, stride 1, tiles of 16 x 8 pixels inside one image): the A operand of the three taps of one filter COLUMN is one
// (8+2)-row x 16-pixel slab, loaded once; tap r is the same shared-memory tile 16 rows (= 2 KB, a multiple of the 1 KB swizzle
// atom) further down, addressed through the UMMA descriptor.  One ring stage = slab + the three weight tiles of that column.
// The deep-K layers are bound by the L2 -> shared-memory fill rate (measured 62 B/clk/SM: 430 clk per 24 KB k-block at BN = 64,
// 515 clk per 32 KB at BN = 128, against 128 / 256 clk of tensor time); the slab cuts the A bytes per filter column from 48 KB to
// 20 KB.
constexpr int CONVP_SLAB_TW = 16, CONVP_SLAB_TH = 8;
constexpr int CONVP_SLAB_A_BYTES = (CONVP_SLAB_TH + 2) * CONVP_SLAB_TW * 128;     // 20 KB
template <int BN, bool SLAB>
__host__ __device__ constexpr int convp_stage_bytes() {
    return SLAB ? CONVP_SLAB_A_BYTES + 3 * BN * 128 : conv_stage_bytes<BN>();
}
// RES = false: a build without the residual tiles (layers without a residual operand): their shared memory goes to the ring
template <int BN, int STAGES, bool SLAB, int EG, bool RES>
__host__ __device__ constexpr int convp_smem_bytes() {
    return STAGES * convp_stage_bytes<BN, SLAB>() + EG * (RES ? 2 : 1) * (BN / 64) * CONV_A_BYTES /* staging + residual */ + 1024 /*align*/ +
           512 /*barriers*/;
}

template <int BN, int STAGES, bool SLAB, int EG, bool RES>
__global__ void __launch_bounds__(convp_threads<EG>(), 1)
    conv_tcp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, const ConvParams p,
                    const int mtiles, const int ntiles) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw_p[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_p) + 1023) & ~uintptr_t(1023));
    constexpr int A_BYTES = SLAB ? CONVP_SLAB_A_BYTES : CONV_A_BYTES;     // per stage
    constexpr int B_BYTES = SLAB ? 3 * BN * 128 : BN * 128;
    uint8_t* sA = smem;                                             // STAGES x A_BYTES
    uint8_t* sB = sA + STAGES * A_BYTES;                            // STAGES x B_BYTES
    uint8_t* sO = sB + STAGES * B_BYTES;                            // EG x BN/64 staging tiles (128 rows x 128 B, 128B swizzle)
    uint8_t* sR = sO + EG * (BN / 64) * CONV_A_BYTES;               // EG x BN/64 residual tiles (RES builds only)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sR + (RES ? EG * (BN / 64) * CONV_A_BYTES : 0));
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* acc_full = bars + 2 * STAGES;        // [2]
    uint64_t* acc_empty = bars + 2 * STAGES + 2;   // [2]  eight arrivals (epilogue warps)
    uint64_t* res_full = bars + 2 * STAGES + 4;    // [EG]
    uint64_t* res_empty = bars + 2 * STAGES + 6;   // [EG] eight arrivals
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 8);
    __shared__ __align__(16) float s_bias[EG][BN];
    __shared__ float s_red[EG][8][2];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = mtiles * ntiles;
    constexpr int TCOLS = 2 * BN;                                   // two accumulator stages

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < STAGES; ++i) {
                mbar_init(&full[i], 1);
                mbar_init(&empty[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&acc_full[i], 1);
                mbar_init(&acc_empty[i], 8);
            }
            for (int i = 0; i < EG; ++i) {
                mbar_init(&res_full[i], 1);
                mbar_init(&res_empty[i], 8);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, TCOLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    const int cblocks = p.Cin >> 6;
    const int num_kb = p.R * p.S * cblocks;
    const int per_img = p.tiles_w * p.tiles_h;

    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
            const int mt = t / ntiles, nblk = t - mt * ntiles;
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tb = mt / per_img;
            const int w0 = tw * p.TW, h0 = th * p.TH, n0 = tb * p.TN;
            const int nbn = nblk * BN;
            if (SLAB) {
                // one stage per (channel block, filter column): the slab and the three weight tiles of that column
                int cb = 0, sc = 0;
                for (int kk = 0; kk < 3 * cblocks; ++kk) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&full[stage], (uint32_t)(A_BYTES + B_BYTES));
                        tma_load_4d(sA + stage * A_BYTES, &tmA, &full[stage], cb * 64, w0 + sc - 1, h0 - 1, n0);
#pragma unroll
                        for (int r = 0; r < 3; ++r)
                            tma_load_2d(sB + stage * B_BYTES + r * (BN * 128), &tmB, &full[stage], (r * 3 + sc) * p.Cin + cb * 64, nbn);
                    }
                    __syncwarp();
                    if (++sc == 3) {
                        sc = 0;
                        ++cb;
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            } else {
            // (cb, s, r) advance as counters: the k-block loop must issue faster than the tensor core drains a stage (128 clk for
            // BN = 64), and two runtime integer divisions per k-block alone cost more than that on a single thread
            int cb = 0, s = 0, r = 0, tapc = 0;            // tapc = (r * S + s) * Cin: weight column of the tap
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full[stage], conv_stage_bytes<BN>());
                    void* a_dst = sA + stage * CONV_A_BYTES;
                    void* b_dst = sB + stage * (BN * 128);
                    if (p.stride == 1) {
                        tma_load_4d(a_dst, &tmA, &full[stage], cb * 64, w0 + s - p.pad, h0 + r - p.pad, n0);
                    } else {
                        const int hr = r - p.pad, wr = s - p.pad;
                        const int ph = hr & 1, pw = wr & 1;
                        const int dh = (hr - ph) >> 1, dw = (wr - pw) >> 1;
                        tma_load_5d(a_dst, &tmA, &full[stage], pw * p.Cin + cb * 64, w0 + dw, ph, h0 + dh, n0);
                    }
                    tma_load_2d(b_dst, &tmB, &full[stage], tapc + cb * 64, nbn);
                }
                __syncwarp();
                if (++cb == cblocks) {
                    cb = 0;
                    tapc += p.Cin;
                    if (++s == p.S) {
                        s = 0;
                        ++r;
                    }
                }
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            }
            if (RES && p.residual != nullptr) {
                const int e = (EG == 2) ? (it & 1) : 0;        // the group that will drain this tile
                const int ei = (EG == 2) ? (it >> 1) : it;     // ... and how many tiles it has drained before
                mbar_wait(&res_empty[e], (ei & 1) ^ 1);        // that group has read its previous tile's residual
                if (elect_one()) {
                    mbar_arrive_expect_tx(&res_full[e], (BN / 64) * CONV_A_BYTES);
                    for (int jb = 0; jb < BN / 64; ++jb) {
                        void* dst = sR + (e * (BN / 64) + jb) * CONV_A_BYTES;
                        if (p.convt) {
                            const int ab = (nblk * BN) / p.CoutT;
                            const int cb0 = (nblk * BN) - ab * p.CoutT + jb * 64;
                            tma_load_5d(dst, &tmR, &res_full[e], (ab & 1) * p.CoutT + cb0, w0, ab >> 1, h0, n0);
                        } else {
                            tma_load_4d(dst, &tmR, &res_full[e], nblk * BN + jb * 64, w0, h0, n0);
                        }
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = umma_idesc_f16(128, BN);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator stage
            tc_fence_after();
            const int nstage = SLAB ? 3 * cblocks : num_kb;
            for (int kb = 0; kb < nstage; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da = umma_desc_sw128(smem_u32(sA + stage * A_BYTES));
                    const uint64_t db = umma_desc_sw128(smem_u32(sB + stage * B_BYTES));
                    if (SLAB) {
#pragma unroll
                        for (int r = 0; r < 3; ++r)              // tap r: the slab 16 rows further down, its own weight tile
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_f16(tmem_base + acc * BN, da + (uint64_t)(r * (CONVP_SLAB_TW * 128 / 16) + k * 2),
                                         db + (uint64_t)(r * (BN * 128 / 16) + k * 2), idesc, (kb | r | k) != 0);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16(tmem_base + acc * BN, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty[stage]);
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (elect_one()) umma_commit(&acc_full[acc]);
            __syncwarp();
        }
    } else {
        // ===================== epilogue warps =====================
        const int q = warp & 3;
        const int eg = (warp - 2) >> 3;                         // epilogue group (0 when EG == 1)
        const int g = ((warp - 2) >> 2) & 1;                    // 0 / 1 within the group
        const int row = q * 32 + lane;
        const int sw = row & 7;
        const int et = threadIdx.x - 64 - eg * CONVP_EPI_THREADS;   // 0..255 within the group
        const int bar_id = 1 + eg;
        uint8_t* const sOg = sO + eg * (BN / 64) * CONV_A_BYTES;
        uint8_t* const sRg = sR + eg * (BN / 64) * CONV_A_BYTES;
        const int lw = row % p.TW;
        const int lh = (row / p.TW) % p.TH;
        const int ln = row / (p.TW * p.TH);
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        constexpr int NCH_ = BN / 64;
        int bias_nblk = -1;
        int it = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
            if (EG == 2 && (it & 1) != eg) continue;            // the other group's tile
            const int acc = it & 1;
            const int gi = (EG == 2) ? (it >> 1) : it;          // tiles this group has drained before
            const int mt = t / ntiles, nblk = t - mt * ntiles;
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tb = mt / per_img;
            const int w0 = tw * p.TW, h0 = th * p.TH, n0 = tb * p.TN;
            const int n = n0 + ln;
            const bool valid = n < p.B;
            const int nr = valid ? n : p.B - 1;
            int cbase, ab = 0;
            if (p.convt) {
                ab = (nblk * BN) / p.CoutT;
                cbase = (nblk * BN) - ab * p.CoutT;
            } else {
                cbase = nblk * BN;
            }
            // the bias slice only changes with the N block (never, for the many single-N-block layers): no global load per tile
            float* bias_s = s_bias[eg];
            if (nblk != bias_nblk) {
                named_bar_sync(bar_id, CONVP_EPI_THREADS);      // every reader of the previous slice is done
                if (et < BN) bias_s[et] = p.bias ? __ldg(p.bias + cbase + et) : 0.f;
                bias_nblk = nblk;
            }
            // per-sample channel vector (time projection): requested before the accumulator wait, consumed after it
            float4 pav[NCH_][8];
            if (p.post_add) {
#pragma unroll
                for (int c = 0; c < NCH_; ++c) {
                    const int ch = (BN == 64) ? g : g * 2 + c;
                    const float* pa = p.post_add + (size_t)nr * p.post_stride + cbase + ch * 32;
#pragma unroll
                    for (int j = 0; j < 8; ++j) pav[c][j] = __ldg(reinterpret_cast<const float4*>(pa) + j);
                }
            }
            // the stores of the previous tile must have read the staging tiles before they are overwritten
            if (et < BN / 64) tma_store_wait_read();            // bulk groups are per thread: the threads that issued the stores wait
            named_bar_sync(bar_id, CONVP_EPI_THREADS);

            mbar_wait(&acc_full[acc], (it >> 1) & 1);
            tc_fence_after();
            if (RES && p.residual != nullptr) mbar_wait(&res_full[eg], gi & 1);
            float gs1 = 0.f, gs2 = 0.f, satm = 0.f;
            constexpr int NCH = BN / 64;                        // 32-column chunks per warp: 1 (BN=64) or 2 (BN=128)
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int jb = (BN == 64) ? 0 : g;              // 64-channel block
                const int half = (BN == 64) ? g : c;            // 32-column half of it
                const int ch = jb * 2 + half;
                uint32_t v[32];
                tmem_ld32(tmem_base + lane_off + (uint32_t)(acc * BN + ch * 32), v);
                tmem_ld_wait();
                uint4* srow = reinterpret_cast<uint4*>(sOg + jb * CONV_A_BYTES + row * 128);
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(&bias_s[ch * 32 + j]);
                    f[j] = __uint_as_float(v[j]) + b4.x;
                    f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
                    f[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
                    f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
                }
                if (RES && p.residual) {
                    const uint4* rrow = reinterpret_cast<const uint4*>(sRg + jb * CONV_A_BYTES + row * 128);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 r4 = rrow[(half * 4 + j) ^ sw];
                        float2 tt;
                        tt = unpack_h2(r4.x); f[j * 8 + 0] += tt.x; f[j * 8 + 1] += tt.y;
                        tt = unpack_h2(r4.y); f[j * 8 + 2] += tt.x; f[j * 8 + 3] += tt.y;
                        tt = unpack_h2(r4.z); f[j * 8 + 4] += tt.x; f[j * 8 + 5] += tt.y;
                        tt = unpack_h2(r4.w); f[j * 8 + 6] += tt.x; f[j * 8 + 7] += tt.y;
                    }
                }
                if (p.act) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], p.act);
                }
                if (p.post_add) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b4 = pav[c][j];
                        f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w;
                    }
                }
                if (p.gn_partial != nullptr && valid) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        gs1 += f[j];
                        gs2 = fmaf(f[j], f[j], gs2);
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
            sat_flush(satm);
            // accumulator stage and residual tile are consumed: hand them back before the store
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&acc_empty[acc]);
                if (RES && p.residual != nullptr) mbar_arrive(&res_empty[eg]);
            }
            if (p.gn_partial != nullptr) {
                gs1 = warp_sum(gs1);
                gs2 = warp_sum(gs2);
                if (lane == 0) {
                    s_red[eg][(warp - 2) & 7][0] = gs1;
                    s_red[eg][(warp - 2) & 7][1] = gs2;
                }
            }
            fence_proxy_async();                                // generic-proxy smem writes -> visible to the TMA engine
            named_bar_sync(bar_id, CONVP_EPI_THREADS);
            if (et < BN / 64) {                                 // one thread per 64-channel block issues its store
                const int jb = et;
                if (p.convt) tma_store_5d(&tmO, sOg + jb * CONV_A_BYTES, (ab & 1) * p.CoutT + cbase + jb * 64, w0, ab >> 1, h0, n0);
                else tma_store_4d(&tmO, sOg + jb * CONV_A_BYTES, cbase + jb * 64, w0, h0, n0);
                tma_store_commit();
            }
            if (p.gn_partial != nullptr) {                      // fixed-order combination: deterministic
                if (p.gn_sub == 4) {
                    if (g == 0 && lane == 0) {
                        const size_t o = (((size_t)nblk * mtiles + mt) * 4 + q) * 2;
                        // same association as the one-tile kernel: a quarter's 64 (or 128) columns are summed per row first
                        p.gn_partial[o] = s_red[eg][q][0] + s_red[eg][4 + q][0];
                        p.gn_partial[o + 1] = s_red[eg][q][1] + s_red[eg][4 + q][1];
                    }
                } else if (et == 64) {
                    const size_t o = ((size_t)nblk * mtiles + mt) * 2;
                    float a = 0.f, b = 0.f;
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        a += s_red[eg][w][0];
                        b += s_red[eg][w][1];
                    }
                    p.gn_partial[o] = a;
                    p.gn_partial[o + 1] = b;
                }
                // s_red is rewritten only after the next tile's first barrier, which every reader passes after reading
            }
        }
        if (et < BN / 64) tma_store_wait_read();                // smem must outlive the bulk reads
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TCOLS);
    }
}

template <int BN, int STAGES, bool SLAB, int EG, bool RES>
inline int conv_tcp_set_attr() {
    static_assert(convp_smem_bytes<BN, STAGES, SLAB, EG, RES>() <= 227 * 1024, "shared memory budget");
    B2D_CUDA(cudaFuncSetAttribute(conv_tcp_kernel<BN, STAGES, SLAB, EG, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  convp_smem_bytes<BN, STAGES, SLAB, EG, RES>()));
    B2D_CUDA(cudaFuncSetAttribute(conv_tcp_kernel<BN, STAGES, SLAB, EG, RES>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
    return 0;
}
inline bool conv_tcp_one_group() {
    static const bool v = getenv("B2D_CONV_ONE_EPI_GROUP") != nullptr;      // A/B: one epilogue group for BN = 64 too
    return v;
}
inline int conv_tcp_init_attrs() {
    B2D_TRY((conv_tcp_set_attr<64, 8, false, 1, true>()));
    B2D_TRY((conv_tcp_set_attr<64, 6, false, 2, true>()));
    B2D_TRY((conv_tcp_set_attr<64, 7, false, 2, false>()));
    B2D_TRY((conv_tcp_set_attr<128, 5, false, 1, true>()));
    B2D_TRY((conv_tcp_set_attr<64, 4, true, 1, true>()));
    B2D_TRY((conv_tcp_set_attr<64, 3, true, 2, true>()));
    B2D_TRY((conv_tcp_set_attr<64, 4, true, 2, false>()));
    B2D_TRY((conv_tcp_set_attr<128, 2, true, 1, true>()));
    return 0;
}

// Eligible: an un-split plan with enough tiles to give every SM at least two (B2D_NO_CONV_PERSIST disables; a slab plan can only
// run here)
inline bool conv_tcp_eligible(const ConvPlan& pl, int num_sms, int min_tiles_per_sm = 2) {
    if (!pl.tc_ready || pl.p.splits != 1) return false;
    if (pl.slab) return true;
    if (conv_persist_disabled()) return false;
    const int tiles = (int)(pl.grid.x * pl.grid.y);
    return tiles >= min_tiles_per_sm * num_sms;
}

template <int BN, int STAGES, bool SLAB, int EG, bool RES>
inline int conv_tcp_launch_t(const ConvPlan& pl, int num_sms, cudaStream_t st) {
    const int mtiles = (int)pl.grid.x, ntiles = (int)pl.grid.y;
    const int tiles = mtiles * ntiles;
    const int grid = tiles < num_sms ? tiles : num_sms;
    B2D_CUDA(launch_k(conv_tcp_kernel<BN, STAGES, SLAB, EG, RES>, dim3(grid), dim3(convp_threads<EG>()),
                      (size_t)convp_smem_bytes<BN, STAGES, SLAB, EG, RES>(), st, pl.tmA, pl.tmB, pl.tmO, pl.tmR, pl.p, mtiles, ntiles));
    return 0;
}
inline int conv_launch_tcp(const ConvPlan& pl, int num_sms, cudaStream_t st) {
    B2D_CHECK(pl.tc_ready && pl.p.splits == 1, "persistent conv needs an un-split plan");
    // BN = 64: two epilogue groups — the tile rate of these layers is set by the epilogue (64 -> 64 3x3 + ReLU: 51 -> 33 us at
    // B=256) — unless the K loop is so deep that the ring depth matters more.  Layers without a residual operand use the builds
    // without residual tiles, which keep the ring depth of the one-group kernel next to the second group's staging tile.
    const int kblocks = pl.p.R * pl.p.S * (pl.p.Cin >> 6);
    const bool res = pl.p.residual != nullptr;
    const bool two = pl.bn == 64 && !conv_tcp_one_group() && kblocks <= 36;
    if (pl.slab) {
        if (pl.bn != 64) return conv_tcp_launch_t<128, 2, true, 1, true>(pl, num_sms, st);
        if (!two) return conv_tcp_launch_t<64, 4, true, 1, true>(pl, num_sms, st);
        return res ? conv_tcp_launch_t<64, 3, true, 2, true>(pl, num_sms, st) : conv_tcp_launch_t<64, 4, true, 2, false>(pl, num_sms, st);
    }
    // deepest rings that fit next to the staging / residual tiles (224 KB): the deep-K layers are paced by the bytes a single
    // SM keeps in flight
    if (pl.bn != 64) return conv_tcp_launch_t<128, 5, false, 1, true>(pl, num_sms, st);
    if (!two) return conv_tcp_launch_t<64, 8, false, 1, true>(pl, num_sms, st);
    return res ? conv_tcp_launch_t<64, 6, false, 2, true>(pl, num_sms, st) : conv_tcp_launch_t<64, 7, false, 2, false>(pl, num_sms, st);
}

}  // namespace b2d
