// Self-attention pieces for ImageSelfAttention (modules_DANRA_conditional.py:91-110) / SelfAttention (unet_ms.py:21-27).
//
// NHWC activations are already the (B, L=H*W, C) token layout, so no permutes exist on this path.
//   layernorm_rows_kernel : LayerNorm over C per token, fp32 statistics, one warp per token.
//   flash_attn_kernel     : softmax(q k^T / sqrt(d)) v per (sample, head) with streaming (online) softmax —
//                           the L x L score matrix the reference materialises (need_weights=True) never exists.
//                           mma.sync m16n8k16 f16 (these layers are ex2-bound at head_dim 16),
//                           K/V tiles double-buffered in shared memory with cp.async.
// The QKV and output projections run through the tcgen05 GEMM path (conv.cuh, 1x1 case) with the residual add
// (+ ReLU in the decoder) in its epilogue.
#pragma once
#include "common.cuh"

namespace b2d {

// ------------------------------------------------------------------------------------------------ LayerNorm
// x,y: [rows][C] f16.  C in {64,...,512}, C % 64 == 0.  One warp per row, each lane owns C/32 contiguous pairs.
template <int C>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const f16* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, f16* __restrict__ y,
                                                             int rows) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int PER = C / 32;  // elements per lane (2..16), contiguous
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const f16* xr = x + (size_t)warp * C + lane * PER;
    float v[PER];
#pragma unroll
    for (int i = 0; i < PER; i += 2) {
        const float2 t = __half22float2(*reinterpret_cast<const f162*>(xr + i));
        v[i] = t.x;
        v[i + 1] = t.y;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) s += v[i];
    const float mean = warp_sum(s) * (1.0f / C);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const float d = v[i] - mean;
        ss += d * d;
    }
    const float rstd = rsqrtf(warp_sum(ss) * (1.0f / C) + 1e-5f);
    f16* yr = y + (size_t)warp * C + lane * PER;
#pragma unroll
    for (int i = 0; i < PER; i += 2) {
        const int c = lane * PER + i;
        const float a = (v[i] - mean) * rstd * gamma[c] + beta[c];
        const float b = (v[i + 1] - mean) * rstd * gamma[c + 1] + beta[c + 1];
        *reinterpret_cast<f162*>(yr + i) = __floats2half2_rn(a, b);
    }
}

inline int layernorm_launch(const f16* x, const float* g, const float* b, f16* y, int rows, int C, cudaStream_t st) {
    const int blocks = (rows + 7) / 8;
    switch (C) {
        case 64: B2D_CUDA(launch_k(layernorm_rows_kernel<64>, dim3(blocks), dim3(256), 0, st, x, g, b, y, rows)); break;
        case 128: B2D_CUDA(launch_k(layernorm_rows_kernel<128>, dim3(blocks), dim3(256), 0, st, x, g, b, y, rows)); break;
        case 256: B2D_CUDA(launch_k(layernorm_rows_kernel<256>, dim3(blocks), dim3(256), 0, st, x, g, b, y, rows)); break;
        case 512: B2D_CUDA(launch_k(layernorm_rows_kernel<512>, dim3(blocks), dim3(256), 0, st, x, g, b, y, rows)); break;
        default: return fail(-1, "layernorm: unsupported channel count " + std::to_string(C));
    }
    B2D_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ flash attention
// qkv: [B*L][3C] f16 (q | k | v, head j owns columns [j*D,(j+1)*D) of each third); o: [B*L][C] f16.
// grid = (ceil(L/64), heads, B); 128 threads; each warp owns 16 query rows; KV tiles of 64 keys.
constexpr int FA_BQ = 64;
constexpr int FA_BK = 64;

template <int D>
__host__ __device__ constexpr int fa_smem_bytes() {
    return 2 /*stages*/ * 2 /*K,V*/ * FA_BK * (D + 8) * 2;
}

template <int D>
__global__ void __launch_bounds__(128) flash_attn_kernel(const f16* __restrict__ qkv, f16* __restrict__ o, int L, int C,
                                                         float scale_log2e) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int LDS = D + 8;  // padded row (elements): conflict-free 32-bit fragment loads and ldmatrix
    extern __shared__ __align__(16) uint8_t fa_smem[];
    f16* sK = reinterpret_cast<f16*>(fa_smem);          // [2][FA_BK][LDS]
    f16* sV = sK + 2 * FA_BK * LDS;                      // [2][FA_BK][LDS]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int head = blockIdx.y, b = blockIdx.z;
    const int q0 = blockIdx.x * FA_BQ + warp * 16;
    const size_t row_stride = (size_t)3 * C;
    const f16* base = qkv + (size_t)b * L * row_stride + (size_t)head * D;

    // ---- Q fragments (A operand, 16 x D), rows q0+g and q0+g+8
    uint32_t qf[D / 16][4];
    {
        const int r0 = q0 + g, r1 = q0 + g + 8;
        const f16* p0 = base + (size_t)r0 * row_stride;
        const f16* p1 = base + (size_t)r1 * row_stride;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
            const int c = kk * 16 + 2 * t;
            qf[kk][0] = r0 < L ? *reinterpret_cast<const uint32_t*>(p0 + c) : 0u;
            qf[kk][1] = r1 < L ? *reinterpret_cast<const uint32_t*>(p1 + c) : 0u;
            qf[kk][2] = r0 < L ? *reinterpret_cast<const uint32_t*>(p0 + c + 8) : 0u;
            qf[kk][3] = r1 < L ? *reinterpret_cast<const uint32_t*>(p1 + c + 8) : 0u;
        }
    }

    // O accumulators; n-tile D/8 is the "ones" column of V (pad block of the smem rows): its column 0 accumulates the
    // softmax denominator in fp32 on the tensor pipe, with the same online rescaling as O.
    float oacc[D / 8 + 1][4];
#pragma unroll
    for (int i = 0; i <= D / 8; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY;

    const int ntiles = (L + FA_BK - 1) / FA_BK;
    constexpr int CPR = D / 8;  // 16-byte chunks per row

    // pad block of every V row (never touched by cp.async): [1, 0, 0, 0, 0, 0, 0, 0]
    for (int i = threadIdx.x; i < 2 * FA_BK; i += 128) {
        uint4 one = make_uint4(0x00003C00u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(sV + (size_t)i * LDS + D) = one;
    }

    auto load_tile = [&](int tile, int buf) {
        const int k0 = tile * FA_BK;
        f16* dK = sK + buf * FA_BK * LDS;
        f16* dV = sV + buf * FA_BK * LDS;
        for (int i = threadIdx.x; i < FA_BK * CPR; i += 128) {
            const int r = i / CPR, c = (i - r * CPR) * 8;
            const bool ok = (k0 + r) < L;
            const f16* src = base + (size_t)(ok ? (k0 + r) : 0) * row_stride + c;
            cp_async16(dK + r * LDS + c, src + C, ok);
            cp_async16(dV + r * LDS + c, src + 2 * C, ok);
        }
    };

    load_tile(0, 0);
    cp_async_commit();

    for (int tile = 0; tile < ntiles; ++tile) {
        const int buf = tile & 1;
        if (tile + 1 < ntiles) load_tile(tile + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const f16* tK = sK + buf * FA_BK * LDS;
        const f16* tV = sV + buf * FA_BK * LDS;

        // ---- S = Q K^T (16 x 64 per warp), fp32
        float s[FA_BK / 8][4];
#pragma unroll
        for (int j = 0; j < FA_BK / 8; ++j) {
            s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
            const f16* kr = tK + (j * 8 + g) * LDS + 2 * t;
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk) {
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr + kk * 16);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kr + kk * 16 + 8);
                mma_f16_16816(s[j], qf[kk], b0, b1);
            }
        }
        // ---- mask keys beyond L (last tile only)
        const int k0 = tile * FA_BK;
        if (k0 + FA_BK > L) {
#pragma unroll
            for (int j = 0; j < FA_BK / 8; ++j) {
                const int kc = k0 + j * 8 + 2 * t;
                if (kc >= L) s[j][0] = s[j][2] = -INFINITY;
                if (kc + 1 >= L) s[j][1] = s[j][3] = -INFINITY;
            }
        }
        // ---- online softmax (rows g and g+8; a row is spread over the 4 lanes of a quad)
        float mx0 = m0, mx1 = m1;
#pragma unroll
        for (int j = 0; j < FA_BK / 8; ++j) {
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float corr0 = ex2_approx((m0 - mx0) * scale_log2e);
        const float corr1 = ex2_approx((m1 - mx1) * scale_log2e);
        m0 = mx0;
        m1 = mx1;
        const float ms0 = mx0 * scale_log2e, ms1 = mx1 * scale_log2e;
        // P = 2^(s*scale - m*scale) straight into the fp16x2 A fragments (2 exponentials per SFU op)
        uint32_t pf[FA_BK / 16][4];
#pragma unroll
        for (int j = 0; j < FA_BK / 8; ++j) {
            pf[j >> 1][(j & 1) * 2 + 0] = ex2_h2(fmaf(s[j][0], scale_log2e, -ms0), fmaf(s[j][1], scale_log2e, -ms0));
            pf[j >> 1][(j & 1) * 2 + 1] = ex2_h2(fmaf(s[j][2], scale_log2e, -ms1), fmaf(s[j][3], scale_log2e, -ms1));
        }
#pragma unroll
        for (int i = 0; i <= D / 8; ++i) {
            oacc[i][0] *= corr0;
            oacc[i][1] *= corr0;
            oacc[i][2] *= corr1;
            oacc[i][3] *= corr1;
        }
        // ---- O += P V : A = P (from registers), B[k][n] = V[key k][dim n] via ldmatrix.trans (+ the ones column)
#pragma unroll
        for (int kk = 0; kk < FA_BK / 16; ++kk) {
            const uint32_t vrow = smem_u32(tV + (kk * 16 + (lane & 15)) * LDS);
#pragma unroll
            for (int i = 0; i <= D / 8; ++i) {
                uint32_t b0, b1;
                ldmatrix_x2_trans(b0, b1, vrow + i * 16);
                mma_f16_16816(oacc[i], pf[kk], b0, b1);
            }
        }
        __syncthreads();  // everyone done with this buffer before it is refilled
    }
    cp_async_wait<0>();

    // ---- finalise: the denominator sits in column 0 of the ones tile (lane t == 0 of each quad)
    const float l0 = __shfl_sync(0xffffffffu, oacc[D / 8][0], lane & ~3);
    const float l1 = __shfl_sync(0xffffffffu, oacc[D / 8][2], lane & ~3);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    const int r0 = q0 + g, r1 = q0 + g + 8;
    f16* ob = o + (size_t)b * L * C + (size_t)head * D + 2 * t;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
        if (r0 < L)
            *reinterpret_cast<uint32_t*>(ob + (size_t)r0 * C + i * 8) = pack_h2(oacc[i][0] * inv0, oacc[i][1] * inv0);
        if (r1 < L)
            *reinterpret_cast<uint32_t*>(ob + (size_t)r1 * C + i * 8) = pack_h2(oacc[i][2] * inv1, oacc[i][3] * inv1);
    }
}

// Head dims below the m16n8k16 K extent (n_heads = 8 or H/2 in some reference scripts, e.g. test/unet_test.py:20-159):
// one thread per query, K/V tiles of 128 keys in shared memory, online softmax in blocks of 8 keys.
template <int D>
__global__ void __launch_bounds__(128) attn_small_d_kernel(const f16* __restrict__ qkv, f16* __restrict__ o, int L, int C,
                                                           float scale_log2e) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sK[128][D];
    __shared__ float sV[128][D];
    const int head = blockIdx.y, b = blockIdx.z;
    const int qi = blockIdx.x * 128 + threadIdx.x;
    const size_t rs = (size_t)3 * C;
    const f16* base = qkv + (size_t)b * L * rs + (size_t)head * D;
    float q[D], acc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        q[d] = qi < L ? __half2float(base[(size_t)qi * rs + d]) * scale_log2e : 0.f;
        acc[d] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    for (int k0 = 0; k0 < L; k0 += 128) {
        __syncthreads();
        {
            const int kr = k0 + threadIdx.x;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                sK[threadIdx.x][d] = kr < L ? __half2float(base[(size_t)kr * rs + C + d]) : 0.f;
                sV[threadIdx.x][d] = kr < L ? __half2float(base[(size_t)kr * rs + 2 * C + d]) : 0.f;
            }
        }
        __syncthreads();
        const int kn = min(128, L - k0);
        for (int kb = 0; kb < kn; kb += 8) {
            float s[8];
            float mx = m;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float a = 0.f;
#pragma unroll
                for (int d = 0; d < D; ++d) a = fmaf(q[d], sK[kb + j][d], a);
                s[j] = (kb + j) < kn ? a : -INFINITY;
                mx = fmaxf(mx, s[j]);
            }
            const float corr = ex2_approx(m - mx);
            m = mx;
            l *= corr;
#pragma unroll
            for (int d = 0; d < D; ++d) acc[d] *= corr;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float p = ex2_approx(s[j] - mx);
                l += p;
#pragma unroll
                for (int d = 0; d < D; ++d) acc[d] = fmaf(p, sV[(kb + j) & 127][d], acc[d]);
            }
        }
    }
    if (qi < L) {
        const float inv = 1.0f / l;
        f16* op = o + ((size_t)b * L + qi) * C + (size_t)head * D;
#pragma unroll
        for (int d = 0; d < D; ++d) op[d] = __float2half_rn(sat_h(acc[d] * inv));
    }
}

// ------------------------------------------------------------------------------------------------ wide heads, few tokens
// head_dim 256 / 512 only occurs with n_heads = 1 (the launcher default of the clean application, test/launch.py:62) at the
// low-resolution levels (C = 256: L = 16 / 64 tokens, C = 512: L = 4 / 16): a few MFLOP per sample.  One CTA per (sample, head):
// Q, K, V of the head in shared memory, the L x L scores in fp32, softmax per row, O = P V.  Requires L <= 64, L * D <= 16384.
constexpr int AW_MAX_L = 64;
constexpr int AW_MAX_LD = 16384;
__global__ void __launch_bounds__(256) attn_wide_kernel(const f16* __restrict__ qkv, f16* __restrict__ o, int L, int C, int D,
                                                        float scale) {
    pdl_launch_dependents();
    extern __shared__ uint8_t aw_raw[];
    f16* sq = reinterpret_cast<f16*>(aw_raw);                 // [L][D]
    f16* sk = sq + (size_t)L * D;
    f16* sv = sk + (size_t)L * D;
    float* sp = reinterpret_cast<float*>(sv + (size_t)L * D);  // [L][L]
    const int head = blockIdx.x, b = blockIdx.y;
    pdl_wait();
    const f16* base = qkv + (size_t)b * L * 3 * C + (size_t)head * D;
    const int d8 = D >> 3;
    for (int i = threadIdx.x; i < L * d8; i += blockDim.x) {
        const int r = i / d8, c8 = (i - r * d8) * 8;
        const f16* row = base + (size_t)r * 3 * C + c8;
        *reinterpret_cast<uint4*>(sq + (size_t)r * D + c8) = __ldg(reinterpret_cast<const uint4*>(row));
        *reinterpret_cast<uint4*>(sk + (size_t)r * D + c8) = __ldg(reinterpret_cast<const uint4*>(row + C));
        *reinterpret_cast<uint4*>(sv + (size_t)r * D + c8) = __ldg(reinterpret_cast<const uint4*>(row + 2 * C));
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int ij = warp; ij < L * L; ij += nw) {               // one warp per score: lanes stride the head dimension
        const int i = ij / L, j = ij - i * L;
        float acc = 0.f;
        for (int c = lane * 2; c < D; c += 64) {
            const float2 a = __half22float2(*reinterpret_cast<const f162*>(sq + (size_t)i * D + c));
            const float2 k2 = __half22float2(*reinterpret_cast<const f162*>(sk + (size_t)j * D + c));
            acc = fmaf(a.x, k2.x, acc);
            acc = fmaf(a.y, k2.y, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) sp[ij] = acc * scale;
    }
    __syncthreads();
    for (int i = warp; i < L; i += nw) {                       // softmax of row i
        float m = -INFINITY;
        for (int j = lane; j < L; j += 32) m = fmaxf(m, sp[i * L + j]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
        float sum = 0.f;
        for (int j = lane; j < L; j += 32) {
            const float e = __expf(sp[i * L + j] - m);
            sp[i * L + j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int j = lane; j < L; j += 32) sp[i * L + j] *= inv;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < L * (D >> 1); idx += blockDim.x) {   // O[i][c..c+1] = sum_j P[i][j] V[j][c..c+1]
        const int i = idx / (D >> 1), c = (idx - i * (D >> 1)) * 2;
        float a0 = 0.f, a1 = 0.f;
        for (int j = 0; j < L; ++j) {
            const float pj = sp[i * L + j];
            const float2 v2 = __half22float2(*reinterpret_cast<const f162*>(sv + (size_t)j * D + c));
            a0 = fmaf(pj, v2.x, a0);
            a1 = fmaf(pj, v2.y, a1);
        }
        *reinterpret_cast<uint32_t*>(o + ((size_t)b * L + i) * C + (size_t)head * D + c) = pack_h2(a0, a1);
    }
}
inline bool attn_wide_supported(int L, int D) { return (D == 256 || D == 512) && L <= AW_MAX_L && L * D <= AW_MAX_LD; }

inline int flash_attn_init_attrs() {
    B2D_CUDA(cudaFuncSetAttribute(attn_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  3 * AW_MAX_LD * 2 + AW_MAX_L * AW_MAX_L * 4));
    B2D_CUDA(cudaFuncSetAttribute(flash_attn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  fa_smem_bytes<128>()));
    B2D_CUDA(cudaFuncSetAttribute(flash_attn_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  fa_smem_bytes<64>()));
    return 0;
}

inline int flash_attn_launch(const f16* qkv, f16* o, int B, int L, int C, int heads, cudaStream_t st) {
    B2D_CHECK(C % heads == 0, "attention: C must be divisible by n_heads");
    const int D = C / heads;
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)D);
    dim3 grid((L + FA_BQ - 1) / FA_BQ, heads, B);
    switch (D) {
        case 2: B2D_CUDA(launch_k(attn_small_d_kernel<2>, dim3(dim3((L + 127) / 128, heads, B)), dim3(128), 0, st, qkv, o, L, C, scale_log2e)); break;
        case 4: B2D_CUDA(launch_k(attn_small_d_kernel<4>, dim3(dim3((L + 127) / 128, heads, B)), dim3(128), 0, st, qkv, o, L, C, scale_log2e)); break;
        case 8: B2D_CUDA(launch_k(attn_small_d_kernel<8>, dim3(dim3((L + 127) / 128, heads, B)), dim3(128), 0, st, qkv, o, L, C, scale_log2e)); break;
        case 16: B2D_CUDA(launch_k(flash_attn_kernel<16>, dim3(grid), dim3(128), fa_smem_bytes<16>(), st, qkv, o, L, C, scale_log2e)); break;
        case 32: B2D_CUDA(launch_k(flash_attn_kernel<32>, dim3(grid), dim3(128), fa_smem_bytes<32>(), st, qkv, o, L, C, scale_log2e)); break;
        case 64: B2D_CUDA(launch_k(flash_attn_kernel<64>, dim3(grid), dim3(128), fa_smem_bytes<64>(), st, qkv, o, L, C, scale_log2e)); break;
        case 128: B2D_CUDA(launch_k(flash_attn_kernel<128>, dim3(grid), dim3(128), fa_smem_bytes<128>(), st, qkv, o, L, C, scale_log2e)); break;
        default:
            if (!attn_wide_supported(L, D))
                return fail(-1, "attention: unsupported head_dim " + std::to_string(D) + " at " + std::to_string(L) + " tokens");
            B2D_CUDA(launch_k(attn_wide_kernel, dim3(heads, B), dim3(256), (size_t)3 * L * D * 2 + (size_t)L * L * 4, st, qkv, o, L, C, D,
                              1.0f / sqrtf((float)D)));
            break;
    }
    B2D_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b2d
