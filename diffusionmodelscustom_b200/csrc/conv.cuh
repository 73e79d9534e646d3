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
    // simt only
    const f16* in;
    const f16* w;
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == 1) return fmaxf(v, 0.0f);
    if (act == 2) return gelu_erf(v);
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
    conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                              // STAGES x 16 KB
    uint8_t* sB = smem + STAGES * CONV_A_BYTES;      // STAGES x BN*128 B
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * conv_stage_bytes<BN>());
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* accum_full = bars + 2 * STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

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
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < STAGES; ++i) {
                mbar_init(&full[i], 1);
                mbar_init(&empty[i], 1);
            }
            mbar_init(accum_full, 1);
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

    const int cblocks = p.Cin >> 6;
    const int num_kb = p.R * p.S * cblocks;

    if (warp == 0) {
        // ===================== TMA producer (one lane) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tap = 0; tap < p.R * p.S; ++tap) {
                const int r = tap / p.S, s = tap - r * p.S;
                for (int cb = 0; cb < cblocks; ++cb) {
                    mbar_wait(&empty[stage], phase ^ 1);
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
                    tma_load_2d(b_dst, &tmB, &full[stage], tap * p.Cin + cb * 64, nblk * BN);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one lane) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(128, BN);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint64_t da = umma_desc_sw128(smem_u32(sA + stage * CONV_A_BYTES));
                const uint64_t db = umma_desc_sw128(smem_u32(sB + stage * (BN * 128)));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // advance 16 f16 = 32 B along K inside the swizzle atom: +2 in the (addr >> 4) field
                    umma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                }
                umma_commit(&empty[stage]);  // smem slot reusable once these MMAs retire
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit(accum_full);
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

        mbar_wait(accum_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), v);
            tmem_ld_wait();
            if (valid) {
                const int c0 = cbase + ch * 32;
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                if (p.bias) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
                        f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                    }
                }
                if (p.residual) {
                    const uint4* rp = reinterpret_cast<const uint4*>(p.residual + obase + ch * 32);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 r4 = __ldg(rp + j);
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
                    const float* pa = p.post_add + (size_t)n * p.post_stride + c0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(pa + j));
                        f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                    }
                }
                uint4* op = reinterpret_cast<uint4*>(p.out + obase + ch * 32);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 o;
                    o.x = pack_h2(f[j * 8 + 0], f[j * 8 + 1]);
                    o.y = pack_h2(f[j * 8 + 2], f[j * 8 + 3]);
                    o.z = pack_h2(f[j * 8 + 4], f[j * 8 + 5]);
                    o.w = pack_h2(f[j * 8 + 6], f[j * 8 + 7]);
                    op[j] = o;
                }
            }
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
                          const uint32_t* box) {
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
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B2D_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return 0;
}

// A fully resolved convolution launch (tensor maps are encoded once; buffers never move).
struct ConvPlan {
    ConvParams p;
    CUtensorMap tmA, tmB;
    int bn = 0;
    dim3 grid;
    bool tc_ready = false;
};

inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// Fills geometry-derived fields of p (tiling) and encodes the tensor maps.
inline int conv_plan_build(ConvPlan& pl, int num_sms) {
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
    const int mtiles = p.tiles_w * p.tiles_h * ((p.B + p.TN - 1) / p.TN);
    // N tile: 128 only when that still fills the machine and does not straddle a convT sub-pixel block
    int bn = 64;
    if (p.Cout % 128 == 0 && (!p.convt || p.CoutT % 128 == 0) && mtiles * (p.Cout / 128) >= num_sms) bn = 128;
    pl.bn = bn;
    pl.grid = dim3(mtiles, p.Cout / bn, 1);
    const uint64_t C = p.Cin, W = p.Wi, H = p.Hi, B = p.B;
    if (p.stride == 1) {
        uint64_t dims[4] = {C, W, H, B};
        uint64_t str[3] = {C * 2, W * C * 2, H * W * C * 2};
        uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN};
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
    pl.tc_ready = true;
    return 0;
}

constexpr int CONV_STAGES_64 = 4;
constexpr int CONV_STAGES_128 = 3;

inline int conv_tc_init_attrs() {
    B2D_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64, CONV_STAGES_64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  conv_smem_bytes<64, CONV_STAGES_64>()));
    B2D_CUDA(cudaFuncSetAttribute(conv_tc_kernel<128, CONV_STAGES_128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  conv_smem_bytes<128, CONV_STAGES_128>()));
    return 0;
}

inline int conv_launch_tc(const ConvPlan& pl, cudaStream_t st) {
    B2D_CHECK(pl.tc_ready, "conv plan not built");
    if (pl.bn == 64)
        conv_tc_kernel<64, CONV_STAGES_64>
            <<<pl.grid, CONV_TC_THREADS, conv_smem_bytes<64, CONV_STAGES_64>(), st>>>(pl.tmA, pl.tmB, pl.p);
    else
        conv_tc_kernel<128, CONV_STAGES_128>
            <<<pl.grid, CONV_TC_THREADS, conv_smem_bytes<128, CONV_STAGES_128>(), st>>>(pl.tmA, pl.tmB, pl.p);
    B2D_CUDA(cudaGetLastError());
    return 0;
}

inline int conv_launch_simt(const ConvParams& p, cudaStream_t st) {
    const size_t total = (size_t)p.B * p.Ho * p.Wo * p.Cout;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    conv_simt_kernel<<<blocks, 256, 0, st>>>(p);
    B2D_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b2d
