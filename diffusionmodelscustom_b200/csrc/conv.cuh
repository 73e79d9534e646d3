// Implicit-GEMM convolution / projection kernels.
//
//   conv_tc_kernel   : tcgen05.mma (UMMA 128 x BN x 16, f16 -> fp32 in TMEM), operands staged by TMA into
//                      128B-swizzled K-major shared-memory tiles, warp-specialised (TMA / MMA / 4 epilogue warps),
//                      mbarrier pipeline.  One 128-pixel x BN-channel output tile per CTA.
//   conv_simt_kernel : one-thread-per-output CUDA-core restatement with the same epilogue; bring-up / unit-test
//                      cross-check only (selected explicitly through b2d_op_conv2d(impl=1) or the debug flag).
//
// Activations are NHWC f16.  Weights are packed [Cout][R*S*Cin] (tap-major, Cin innermost, K-major for UMMA).
// GEMM view: M = B*Ho*Wo output pixels, N = Cout, K = R*S*Cin; K-blocks of 64 = one 128-byte swizzle row.
//   * stride 1: the A tile of tap (r,s) is a shifted 4-D TMA box (64ch, TW, TH, TN) of the input; the zero padding is
//     TMA out-of-bounds fill (signed start coordinates).
//   * stride 2: the input is viewed as [B][Hi/2][2][Wi/2][2*Cin] (row/column parity split); tap (r,s) has a fixed parity
//     and the box is again dense -> 5-D TMA.
//   * ConvTranspose2d(k=2,s=2) is the GEMM [B*h*w, Cin] x [Cin, 4*Cout] with a pixel-shuffle store
//     (SURVEY.md App. A); packed weight rows are (a*2+b)*Cout + co.
// Epilogue (per output element): v = acc + bias[c]; v += residual; v = act(v); v += post_add[b][c]; store f16.
#pragma once
#include "common.cuh"

namespace b2d {

struct ConvParams {
    int B, Hi, Wi, Cin;
    int Ho, Wo, Cout;  // GEMM pixel grid and GEMM N (convT: Ho=Hi, Wo=Wi, Cout = 4*CoutT)
    int R, S, stride, pad;
    int convt;   // 1 = ConvTranspose k2 s2 store
    int CoutT;   // channels of the stored tensor (== Cout unless convt)
    // M tiling: 128 rows = TN images x TH rows x TW cols (all powers of two)
    int TW, TH, TN, tiles_w, tiles_h;
    // epilogue
    const float* bias;
    const f16* residual;
    const float* post_add;
    int post_stride;
    int act;  // 0 none, 1 relu, 2 gelu(erf)
    f16* out;
    // LayerNorm folded into the GEMM (gemm_stream only): A holds the UN-normalised rows, the weights are pre-scaled by gamma,
    // out = rstd_m * (acc - mean_m * ln_c1[n]) + bias[n]  with ln_c1 = column sums of the scaled weights and
    // bias = W beta + b.  Row statistics are computed by the epilogue threads from the A tile in shared memory.
    const float* ln_c1;
    // GroupNorm(1,C) statistics of the OUTPUT fused into the epilogue (conv_tc, unsplit, one sample per M tile): every CTA
    // writes {sum, sum of squares} of its tile to gn_partial[(nblk * mtiles + mtile) * 2]; the consumer adds them in order.
    float* gn_partial;
    int gn_sub;          // 1: one {sum, sumsq} per 128-row tile; 4: one per 32-row quarter (tiles that span several samples)
    // split-K: gridDim.z = splits CTAs of one cluster share an output tile; fp32 partials go through `ws`
    int splits;
    float* ws;  // [tiles][splits][128][BN] fp32
    // simt only
    const f16* in;
    const f16* w;
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == 1) return fmaxf(v, 0.0f);
    if (act == 2) return gelu_erf(v);
    return v;
}

// Epilogue of 8 consecutive channels of one output pixel: bias -> +residual -> act -> +post_add -> fp16 store (16 B).
__device__ __forceinline__ void conv_epilogue8(const ConvParams& p, float (&f)[8], int n, size_t off, int c0) {
    if (p.bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c0));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + 4));
        f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
        f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
    }
    if (p.residual) {
        const uint4 r4 = __ldg(reinterpret_cast<const uint4*>(p.residual + off));
        float2 t;
        t = unpack_h2(r4.x); f[0] += t.x; f[1] += t.y;
        t = unpack_h2(r4.y); f[2] += t.x; f[3] += t.y;
        t = unpack_h2(r4.z); f[4] += t.x; f[5] += t.y;
        t = unpack_h2(r4.w); f[6] += t.x; f[7] += t.y;
    }
    if (p.act) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = apply_act(f[j], p.act);
    }
    if (p.post_add) {
        const float* pa = p.post_add + (size_t)n * p.post_stride + c0;
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(pa));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(pa + 4));
        f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
        f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
    }
    uint4 o;
    o.x = pack_h2(f[0], f[1]); o.y = pack_h2(f[2], f[3]);
    o.z = pack_h2(f[4], f[5]); o.w = pack_h2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p.out + off) = o;
}

__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// 16-byte load from the shared memory of CTA `rank` of this cluster (distributed shared memory)
__device__ __forceinline__ float4 ld_dsmem_f4(const void* local_ptr, int rank) {
    uint32_t a = smem_u32(local_ptr), ra;
    float4 v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra));
    return v;
}

// ------------------------------------------------------------------------------------------------ tcgen05 path
constexpr int CONV_TC_THREADS = 192;  // warp0 TMA, warp1 MMA(+TMEM alloc), warps2-5 epilogue
constexpr int CONV_A_BYTES = 128 * 64 * 2;

template <int BN>
__host__ __device__ constexpr int conv_stage_bytes() {
    return CONV_A_BYTES + BN * 64 * 2;
}
template <int BN, int STAGES>
__host__ __device__ constexpr int conv_smem_bytes() {
    return STAGES * conv_stage_bytes<BN>() + 1024 /*align slack*/ + 256 /*barriers*/;
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(CONV_TC_THREADS)
    conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, const ConvParams p) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                              // STAGES x 16 KB
    uint8_t* sB = smem + STAGES * CONV_A_BYTES;      // STAGES x BN*128 B
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * conv_stage_bytes<BN>());
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* accum_full = bars + 2 * STAGES;
    uint64_t* res_full = bars + 2 * STAGES + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);
    __shared__ __align__(16) float s_bias[BN];   // bias of this N tile, staged while the main loop runs

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // tile coordinates
    const int mt = blockIdx.x;
    const int tw = mt % p.tiles_w;
    const int th = (mt / p.tiles_w) % p.tiles_h;
    const int tb = mt / (p.tiles_w * p.tiles_h);
    const int w0 = tw * p.TW, h0 = th * p.TH, n0 = tb * p.TN;
    const int nblk = blockIdx.y;

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
            mbar_init(accum_full, 1);
            mbar_init(res_full, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched only below

    const int cblocks = p.Cin >> 6;
    const int total_kb = p.R * p.S * cblocks;
    const int split = blockIdx.z;                              // == rank in the (1,1,splits) cluster
    const int kb_begin = (int)(((long long)total_kb * split) / p.splits);
    const int kb_end = (int)(((long long)total_kb * (split + 1)) / p.splits);
    const int num_kb = kb_end - kb_begin;

    if (warp == 0) {
        // ===================== TMA producer =====================
        // The loop runs CONVERGED on all 32 lanes (waits and index arithmetic are warp-uniform and stay on the uniform
        // datapath); one hardware-elected lane issues the TMA instructions.  Inside an `if (lane == 0)` region ptxas wraps
        // every TMA / tcgen05 instruction in an ELECT + R2UR.BROADCAST waterfall loop (~100 clk per instruction, measured on
        // the attention issuer: 543 -> 185 clk for four MMAs and two commits).
        {
            int stage = 0;
            uint32_t phase = 0;
            // (cb, s, r) advance as counters (one division pair at the start for split-K): two runtime integer divisions per
            // k-block made this single-thread loop slower (~540 clk per k-block, measured through the persistent variant) than the
            // tensor core drains a stage (128 clk at BN = 64) — it paced every deep-K launch
            int tap0 = kb_begin / cblocks;
            int cb = kb_begin - tap0 * cblocks, r = tap0 / p.S, s = tap0 - r * p.S, tapc = tap0 * p.Cin;
            const int nbn = nblk * BN;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
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
            // residual tile(s): extra ring slots that only the epilogue consumes; they land while the last MMAs run
            if (p.residual != nullptr && p.splits == 1) {
                if (elect_one()) mbar_arrive_expect_tx(res_full, (BN / 64) * CONV_A_BYTES);
                __syncwarp();
                for (int jb = 0; jb < BN / 64; ++jb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (elect_one()) {
                        void* dst = sA + stage * CONV_A_BYTES;
                        if (p.convt) {
                            const int ab = (nblk * BN) / p.CoutT;
                            const int cb0 = (nblk * BN) - ab * p.CoutT + jb * 64;
                            tma_load_5d(dst, &tmR, res_full, (ab & 1) * p.CoutT + cb0, w0, ab >> 1, h0, n0);
                        } else {
                            tma_load_4d(dst, &tmR, res_full, nblk * BN + jb * 64, w0, h0, n0);
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (converged loop, elected lane issues) =====================
        {
            constexpr uint32_t idesc = umma_idesc_f16(128, BN);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da = umma_desc_sw128(smem_u32(sA + stage * CONV_A_BYTES));
                    const uint64_t db = umma_desc_sw128(smem_u32(sB + stage * (BN * 128)));
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // advance 16 f16 = 32 B along K inside the swizzle atom: +2 in the (addr >> 4) field
                        umma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty[stage]);  // smem slot reusable once these MMAs retire
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (elect_one()) umma_commit(accum_full);
        }
        __syncwarp();
    } else {
        // ===================== epilogue warps: TMEM -> regs -> global =====================
        const int q = warp & 3;             // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;      // tile row = output pixel
        const int lw = row % p.TW;
        const int lh = (row / p.TW) % p.TH;
        const int ln = row / (p.TW * p.TH);
        const int n = n0 + ln, h = h0 + lh, w = w0 + lw;
        const bool valid = n < p.B;

        int cbase;      // first channel of this N tile in the stored tensor
        size_t pix;     // pixel index in the stored tensor
        if (p.convt) {
            const int ab = (nblk * BN) / p.CoutT;
            cbase = (nblk * BN) - ab * p.CoutT;
            pix = ((size_t)n * (2 * p.Ho) + (2 * h + (ab >> 1))) * (size_t)(2 * p.Wo) + (2 * w + (ab & 1));
        } else {
            cbase = nblk * BN;
            pix = ((size_t)n * p.Ho + h) * (size_t)p.Wo + w;
        }
        const size_t obase = pix * (size_t)p.CoutT + cbase;
        (void)obase;
        {   // stage the bias slice of this tile in shared memory before the accumulator is ready
            const int et = threadIdx.x - 64;
            if (et < BN) s_bias[et] = p.bias ? __ldg(p.bias + cbase + et) : 0.f;
            named_bar_sync(1, 128);
        }

        mbar_wait(accum_full, 0);
        tc_fence_after();
        if (p.splits > 1) {
            if (p.ws == nullptr) {
                // raw fp32 partial tile -> this CTA's own (drained) pipeline memory, read by the cluster peers through DSMEM;
                // 16-byte chunks XOR-swizzled by the row so that a warp's 32 rows hit all banks
                float4* prow = reinterpret_cast<float4*>(sA) + (size_t)row * (BN / 4);
#pragma unroll 1
                for (int ch = 0; ch < BN / 32; ++ch) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        prow[(ch * 8 + j) ^ (row & 7)] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                     __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                }
            } else {
            // raw fp32 partial tile -> workspace (this thread's row: BN contiguous floats)
            const size_t tile_id = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
            float* wrow = p.ws + ((tile_id * p.splits + split) * 128 + row) * BN;
#pragma unroll 1
            for (int ch = 0; ch < BN / 32; ++ch) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    __stcg(reinterpret_cast<float4*>(wrow + ch * 32 + j),
                           make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                       __uint_as_float(v[j + 3])));
            }
            }
        } else {
            // Coalesced epilogue: each thread (= output pixel) builds its 128-byte row of a 64-channel block in a
            // 128B-swizzled staging tile (a free pipeline stage; the residual block, if any, was TMA-loaded into the same
            // tile and is read by the same thread first), then one thread issues a TMA store of the whole block.
            const int nr = n < p.B ? n : p.B - 1;                 // out-of-range rows are clipped by the store; keep reads in range
            float gs1 = 0.f, gs2 = 0.f;                           // GroupNorm partial sums of this thread's row
            float satm = 0.f;                                     // |value| maximum of this thread's conversions (sat_flush below)
            if (p.residual != nullptr) mbar_wait(res_full, 0);
#pragma unroll 1
            for (int jb = 0; jb < BN / 64; ++jb) {
                uint8_t* stg = sA + ((num_kb + jb) % STAGES) * CONV_A_BYTES;
                uint4* srow = reinterpret_cast<uint4*>(stg + row * 128);
                const int sw = row & 7;
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    const int ch = jb * 2 + half;
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), v);
                    tmem_ld_wait();
                    const int c0 = cbase + ch * 32;
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[ch * 32 + j]);
                            f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                        }
                    }
                    if (p.residual) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 r4 = srow[(half * 4 + j) ^ sw];
                            float2 t;
                            t = unpack_h2(r4.x); f[j * 8 + 0] += t.x; f[j * 8 + 1] += t.y;
                            t = unpack_h2(r4.y); f[j * 8 + 2] += t.x; f[j * 8 + 3] += t.y;
                            t = unpack_h2(r4.z); f[j * 8 + 4] += t.x; f[j * 8 + 5] += t.y;
                            t = unpack_h2(r4.w); f[j * 8 + 6] += t.x; f[j * 8 + 7] += t.y;
                        }
                    }
                    if (p.act) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], p.act);
                    }
                    if (p.post_add) {
                        const float* pa = p.post_add + (size_t)nr * p.post_stride + c0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(pa + j));
                            f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
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
                fence_proxy_async();                               // generic-proxy smem writes -> visible to the TMA engine
                named_bar_sync(1, 128);                            // the four epilogue warps
                if (threadIdx.x == 64) {
                    if (p.convt) {
                        const int ab = (nblk * BN) / p.CoutT;
                        tma_store_5d(&tmO, stg, (ab & 1) * p.CoutT + cbase + jb * 64, w0, ab >> 1, h0, n0);
                    } else {
                        tma_store_4d(&tmO, stg, cbase + jb * 64, w0, h0, n0);
                    }
                    tma_store_commit();
                }
            }
            sat_flush(satm);
            if (p.gn_partial != nullptr) {                         // block-reduce the tile's {sum, sumsq} in fixed order
                gs1 = warp_sum(gs1);
                gs2 = warp_sum(gs2);
                if (lane == 0) {
                    s_bias[2 * (warp - 2)] = gs1;                  // s_bias is dead by now: reuse as scratch
                    s_bias[2 * (warp - 2) + 1] = gs2;
                }
                named_bar_sync(1, 128);
                if (p.gn_sub == 4) {                               // per 32-row quarter (row order: quarter q = warp & 3)
                    if (lane == 0) {
                        const size_t t = (((size_t)nblk * gridDim.x + blockIdx.x) * 4 + q) * 2;
                        p.gn_partial[t] = gs1;
                        p.gn_partial[t + 1] = gs2;
                    }
                } else if (threadIdx.x == 64) {
                    const size_t t = ((size_t)nblk * gridDim.x + blockIdx.x) * 2;
                    p.gn_partial[t] = (s_bias[0] + s_bias[2]) + (s_bias[4] + s_bias[6]);
                    p.gn_partial[t + 1] = (s_bias[1] + s_bias[3]) + (s_bias[5] + s_bias[7]);
                }
            }
            if (threadIdx.x == 64) tma_store_wait_read();          // smem must outlive the bulk reads; the writes complete with the grid
        }
    }
    if (p.splits > 1) {
        // all CTAs of the cluster have written their partial tiles; each finalises 128/splits rows in fixed split order
        // barrier.cluster arrive.release / wait.acquire order the partial-tile writes (st.global.cg) before the peers' ld.global.cg
        cluster_arrive_release();
        cluster_wait_acquire();
        if (warp >= 2) {
            const int et = threadIdx.x - 64;                       // 0..127
            const int rows_per = 128 / p.splits;
            const int items = rows_per * (BN / 8);
            const size_t tile_id = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
            const float* wtile = p.ws + tile_id * p.splits * 128 * BN;
            for (int it = et; it < items; it += 128) {
                const int row = split * rows_per + it / (BN / 8);
                const int c8 = (it % (BN / 8)) * 8;
                float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                float4 pa[8], pb[8];      // every split's partial in flight at once (one L2 round trip, not `splits`), summed in split order
                if (p.ws == nullptr) {
                    const float4* prow = reinterpret_cast<const float4*>(sA) + (size_t)row * (BN / 4);
                    const int jc = c8 >> 2;
#pragma unroll
                    for (int sp = 0; sp < 8; ++sp) {
                        if (sp < p.splits) {
                            pa[sp] = ld_dsmem_f4(prow + (jc ^ (row & 7)), sp);
                            pb[sp] = ld_dsmem_f4(prow + ((jc + 1) ^ (row & 7)), sp);
                        }
                    }
                } else {
#pragma unroll
                for (int sp = 0; sp < 8; ++sp) {
                    if (sp < p.splits) {
                        const float4* src = reinterpret_cast<const float4*>(wtile + ((size_t)sp * 128 + row) * BN + c8);
                        pa[sp] = __ldcg(src);
                        pb[sp] = __ldcg(src + 1);
                    }
                }
                }
#pragma unroll
                for (int sp = 0; sp < 8; ++sp) {
                    if (sp < p.splits) {
                        f[0] += pa[sp].x; f[1] += pa[sp].y; f[2] += pa[sp].z; f[3] += pa[sp].w;
                        f[4] += pb[sp].x; f[5] += pb[sp].y; f[6] += pb[sp].z; f[7] += pb[sp].w;
                    }
                }
                const int lw = row % p.TW;
                const int lh = (row / p.TW) % p.TH;
                const int ln = row / (p.TW * p.TH);
                const int n = n0 + ln, h = h0 + lh, w = w0 + lw;
                if (n >= p.B) continue;
                int cbase;
                size_t pix;
                if (p.convt) {
                    const int ab = (nblk * BN) / p.CoutT;
                    cbase = (nblk * BN) - ab * p.CoutT;
                    pix = ((size_t)n * (2 * p.Ho) + (2 * h + (ab >> 1))) * (size_t)(2 * p.Wo) + (2 * w + (ab & 1));
                } else {
                    cbase = nblk * BN;
                    pix = ((size_t)n * p.Ho + h) * (size_t)p.Wo + w;
                }
                conv_epilogue8(p, f, n, pix * (size_t)p.CoutT + cbase + c8, cbase + c8);
            }
        }
        if (p.ws == nullptr) {   // no CTA may exit (and release its shared memory) while a peer can still read its partial tile
            cluster_arrive_release();
            cluster_wait_acquire();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
    }
}

// ------------------------------------------------------------------------------------------------ SIMT cross-check
__global__ void conv_simt_kernel(const ConvParams p) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t total = (size_t)p.B * p.Ho * p.Wo * p.Cout;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int co = (int)(idx % p.Cout);
        size_t m = idx / p.Cout;
        const int w = (int)(m % p.Wo);
        m /= p.Wo;
        const int h = (int)(m % p.Ho);
        const int n = (int)(m / p.Ho);
        const int K = p.R * p.S * p.Cin;
        const f16* wrow = p.w + (size_t)co * K;
        float acc = 0.0f;
        for (int r = 0; r < p.R; ++r) {
            const int hi = h * p.stride + r - p.pad;
            if (hi < 0 || hi >= p.Hi) continue;
            for (int s = 0; s < p.S; ++s) {
                const int wi = w * p.stride + s - p.pad;
                if (wi < 0 || wi >= p.Wi) continue;
                const f16* ip = p.in + (((size_t)n * p.Hi + hi) * p.Wi + wi) * p.Cin;
                const f16* wp = wrow + (r * p.S + s) * p.Cin;
                for (int c = 0; c < p.Cin; c += 2) {
                    const float2 a = __half22float2(*reinterpret_cast<const f162*>(ip + c));
                    const float2 b = __half22float2(*reinterpret_cast<const f162*>(wp + c));
                    acc = fmaf(a.x, b.x, acc);
                    acc = fmaf(a.y, b.y, acc);
                }
            }
        }
        int c;
        size_t pix;
        if (p.convt) {
            const int ab = co / p.CoutT;
            c = co - ab * p.CoutT;
            pix = ((size_t)n * (2 * p.Ho) + (2 * h + (ab >> 1))) * (size_t)(2 * p.Wo) + (2 * w + (ab & 1));
        } else {
            c = co;
            pix = ((size_t)n * p.Ho + h) * (size_t)p.Wo + w;
        }
        const size_t o = pix * p.CoutT + c;
        float v = acc;
        if (p.bias) v += p.bias[c];
        if (p.residual) v += __half2float(p.residual[o]);
        v = apply_act(v, p.act);
        if (p.post_add) v += p.post_add[(size_t)n * p.post_stride + c];
        p.out[o] = __float2half_rn(sat_h(v));
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

inline int make_tmap_f16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
                          const uint32_t* box, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    PFN_encodeTiled fn = get_encode_fn();
    B2D_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gd[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i < rank - 1) gs[i] = strides_b[i];
    }
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B2D_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return 0;
}

// A fully resolved convolution launch (tensor maps are encoded once; buffers never move).
struct ConvPlan {
    ConvParams p;
    CUtensorMap tmA, tmB, tmO, tmR;
    int bn = 0;
    int stages = 4;
    dim3 grid;
    bool tc_ready = false;
    bool slab = false;     // 16 x 8 pixel tiles with (8+2)-row A slabs: only the persistent kernel (conv_persist.cuh) runs such a plan
    size_t ws_floats = 0;  // split-K workspace the caller must provide in p.ws before launching
};
inline bool conv_persist_disabled() {
    static const bool off = getenv("B2D_NO_CONV_PERSIST") != nullptr;
    return off;
}

inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// Fills geometry-derived fields of p (tiling) and encodes the tensor maps.
inline int conv_plan_build(ConvPlan& pl, int num_sms, bool allow_slab = true) {
    ConvParams& p = pl.p;
    B2D_CHECK(p.Cin % 64 == 0, "tcgen05 conv needs Cin % 64 == 0");
    B2D_CHECK(p.Cout % 64 == 0, "tcgen05 conv needs Cout % 64 == 0");
    B2D_CHECK(is_pow2(p.Ho) && is_pow2(p.Wo) && p.Wo <= 128, "output extent must be a power of two (<=128 wide)");
    B2D_CHECK(p.stride == 1 || p.stride == 2, "stride must be 1 or 2");
    if (p.stride == 2) B2D_CHECK(p.Hi % 2 == 0 && p.Wi % 2 == 0, "stride-2 conv needs even input extent");
    p.TW = p.Wo < 128 ? p.Wo : 128;
    p.TH = (128 / p.TW) < p.Ho ? (128 / p.TW) : p.Ho;
    p.TN = 128 / (p.TW * p.TH);
    p.tiles_w = p.Wo / p.TW;
    p.tiles_h = p.Ho / p.TH;
    // slab tiling for the large 3x3 / stride-1 layers (see conv_persist.cuh): 16 x 8 pixel tiles inside one image
    static const bool no_slab = getenv("B2D_NO_CONV_SLAB") != nullptr;
    pl.slab = false;
    if (allow_slab && !no_slab && !conv_persist_disabled() && !p.convt && p.R == 3 && p.S == 3 && p.stride == 1 && p.pad == 1 &&
        p.Wo % 16 == 0 && p.Ho % 8 == 0) {
        const int mt = (p.Wo / 16) * (p.Ho / 8) * p.B;
        const int nt = (p.Cout % 128 == 0 && mt * (p.Cout / 128) >= num_sms) ? p.Cout / 128 : p.Cout / 64;
        if (mt * nt >= 2 * num_sms) {
            pl.slab = true;
            p.TW = 16; p.TH = 8; p.TN = 1;
            p.tiles_w = p.Wo / 16;
            p.tiles_h = p.Ho / 8;
        }
    }
    const int mtiles = p.tiles_w * p.tiles_h * ((p.B + p.TN - 1) / p.TN);
    // N tile: 128 only when that still fills the machine and does not straddle a convT sub-pixel block
    int bn = 64;
    if (p.Cout % 128 == 0 && (!p.convt || p.CoutT % 128 == 0) && mtiles * (p.Cout / 128) >= num_sms) bn = 128;
    pl.bn = bn;
    // split-K over a (1,1,S) cluster when the tile grid cannot fill the machine and K is deep
    const int tiles = mtiles * (p.Cout / bn);
    const int total_kb = p.R * p.S * (p.Cin / 64);
    int splits = 1;
    while (splits < 8 && tiles * splits * 2 <= num_sms && total_kb / (splits * 2) >= 4) splits *= 2;
    p.splits = splits;
    static const bool splitk_l2 = getenv("B2D_SPLITK_L2") != nullptr;
    pl.ws_floats = (splits > 1 && splitk_l2) ? (size_t)tiles * splits * 128 * bn : 0;
    // deep pipelines for deep-K problems that leave SMs to spare anyway (the TMA round trip paces them)
    const int kb_cta = total_kb / splits;
    const bool deep = tiles * splits <= num_sms && (splits == 1 ? kb_cta >= 6 : kb_cta >= 12);   // measured: 8-stage CTAs in
    pl.stages = (bn == 64) ? (deep ? 8 : 4) : (deep ? 6 : 3);                                     // 8-CTA clusters schedule worse
    static const bool shallow_only = getenv("B2D_CONV_SHALLOW") != nullptr;                       // A/B: always the small rings
    if (shallow_only) pl.stages = (bn == 64) ? 4 : 3;
    pl.grid = dim3(mtiles, p.Cout / bn, splits);
    const uint64_t C = p.Cin, W = p.Wi, H = p.Hi, B = p.B;
    if (p.stride == 1) {
        uint64_t dims[4] = {C, W, H, B};
        uint64_t str[3] = {C * 2, W * C * 2, H * W * C * 2};
        uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)(pl.slab ? p.TH + 2 : p.TH), (uint32_t)p.TN};
        B2D_TRY(make_tmap_f16(&pl.tmA, p.in, 4, dims, str, box));
    } else {
        uint64_t dims[5] = {2 * C, W / 2, 2, H / 2, B};
        uint64_t str[4] = {2 * C * 2, W * C * 2, 2 * W * C * 2, H * W * C * 2};
        uint32_t box[5] = {64, (uint32_t)p.TW, 1, (uint32_t)p.TH, (uint32_t)p.TN};
        B2D_TRY(make_tmap_f16(&pl.tmA, p.in, 5, dims, str, box));
    }
    const uint64_t K = (uint64_t)p.R * p.S * p.Cin;
    uint64_t wd[2] = {K, (uint64_t)p.Cout};
    uint64_t ws[1] = {K * 2};
    uint32_t wb[2] = {64, (uint32_t)bn};
    B2D_TRY(make_tmap_f16(&pl.tmB, p.w, 2, wd, ws, wb));
    // output (and residual) tiles for the TMA-store epilogue: the 128-pixel tile is a dense box of the stored tensor
    auto out_map = [&](CUtensorMap* tm, const void* base) -> int {
        const uint64_t Co = p.CoutT, Wo = p.Wo, Ho = p.Ho;
        if (p.convt) {   // stored tensor [B][2Ho][2Wo][Co] viewed as [n][h][a][w][(b,c)]
            uint64_t dims[5] = {2 * Co, Wo, 2, Ho, B};
            uint64_t str[4] = {2 * Co * 2, 2 * Wo * Co * 2, 4 * Wo * Co * 2, 4 * Ho * Wo * Co * 2};
            uint32_t box[5] = {64, (uint32_t)p.TW, 1, (uint32_t)p.TH, (uint32_t)p.TN};
            return make_tmap_f16(tm, base, 5, dims, str, box);
        }
        uint64_t dims[4] = {Co, Wo, Ho, B};
        uint64_t str[3] = {Co * 2, Wo * Co * 2, Ho * Wo * Co * 2};
        uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN};
        return make_tmap_f16(tm, base, 4, dims, str, box);
    };
    B2D_TRY(out_map(&pl.tmO, p.out));
    if (p.residual) B2D_TRY(out_map(&pl.tmR, p.residual));
    else pl.tmR = pl.tmO;
    pl.tc_ready = true;
    return 0;
}

template <int BN, int STAGES>
inline int conv_tc_set_attr() {
    B2D_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  conv_smem_bytes<BN, STAGES>()));
    return 0;
}
inline int conv_tc_init_attrs() {
    B2D_TRY((conv_tc_set_attr<64, 4>()));
    B2D_TRY((conv_tc_set_attr<64, 8>()));
    B2D_TRY((conv_tc_set_attr<128, 3>()));
    B2D_TRY((conv_tc_set_attr<128, 6>()));
    return 0;
}

template <int BN, int STAGES>
inline int conv_tc_launch_t(const ConvPlan& pl, cudaStream_t st) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = pl.grid;
    cfg.blockDim = dim3(CONV_TC_THREADS);
    cfg.dynamicSmemBytes = conv_smem_bytes<BN, STAGES>();
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = pl.p.splits;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = g_pdl_enabled;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    B2D_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, STAGES>, pl.tmA, pl.tmB, pl.tmO, pl.tmR, pl.p));
    return 0;
}

inline int conv_launch_tc(const ConvPlan& pl, cudaStream_t st) {
    B2D_CHECK(pl.tc_ready, "conv plan not built");
    // split-K partial tiles: through DSMEM (p.ws == nullptr) or through an L2 workspace (p.ws set: B2D_SPLITK_L2 A/B switch)
    if (pl.bn == 64) return pl.stages == 8 ? conv_tc_launch_t<64, 8>(pl, st) : conv_tc_launch_t<64, 4>(pl, st);
    return pl.stages == 6 ? conv_tc_launch_t<128, 6>(pl, st) : conv_tc_launch_t<128, 3>(pl, st);
}

inline int conv_launch_simt(const ConvParams& p, cudaStream_t st) {
    const size_t total = (size_t)p.B * p.Ho * p.Wo * p.Cout;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    B2D_CUDA(launch_k(conv_simt_kernel, dim3(blocks), dim3(256), 0, st, p));
    B2D_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b2d
