// One-pass normalisation kernels: statistics AND application from a single read of the tensor.
// A thread-block cluster (1, 2, 4 or 8 CTAs) holds the whole normalisation slab in shared memory (up to 8 x 64 KB):
//   MODE 0 InstanceNorm2d : slab = one sample x 64 channels x all pixels; per-channel mean / rstd
//                           (modules_DANRA_conditional.py:409,417 — biased variance, eps 1e-5, no affine)
//   MODE 1 GroupNorm(1,C) : slab = one whole sample (C*H*W elements); scalar mean / rstd, per-channel affine, optional
//                           residual add + GELU, optional time-projection add (unet_ms.py:39-47)
// Each CTA loads its rows with cp.async, reduces them, publishes partial sums in its own shared memory; after a cluster
// barrier every CTA adds the partials of all ranks in rank order through distributed shared memory (deterministic), then
// normalises its rows straight out of shared memory.  Replaces plane_stats+instnorm_apply / sample_stats+groupnorm_apply
// (two launches, two reads) wherever the slab fits.
#pragma once
#include "common.cuh"

namespace b2d {

constexpr int NF_MAX_ROWS = 512;                  // rows of 128 B per CTA (64 KB => up to 3 CTAs per SM overlap their phases)
constexpr int NF_SMEM = NF_MAX_ROWS * 128 + 1024;

__device__ __forceinline__ float ld_dsmem_f32(const float* local_ptr, int rank) {
    uint32_t a = smem_u32(local_ptr), ra;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra));
    return v;
}

struct NormParams {
    const f16* x;
    f16* y;
    const f16* add;          // skip (IN) / residual (GN), same shape as x, or null
    const float* vec;        // [B][vec_stride] per-sample per-channel vector added at the end, or null
    int vec_stride;
    const float* gamma;      // GN affine [C] (MODE 1)
    const float* beta;
    int act;                 // MODE 1: 0 none, 2 GELU(erf) (applied after the optional residual add)
    int C;                   // channels of the tensor
    int rows;                // rows of 128 B in the slab (MODE 0: pixels; MODE 1: HW*C/64)
    int rows_per_cta;        // rows handled by one CTA of the cluster
    long long slab_stride;   // elements between consecutive slabs along blockIdx.z (sample)
};

template <int MODE>
__global__ void __launch_bounds__(256, 3) norm_fused_kernel(const NormParams p) {
    pdl_launch_dependents();
    extern __shared__ __align__(128) uint8_t nf_smem[];
    uint8_t* tile = nf_smem;                                   // [rows_per_cta][128 B]
    __shared__ float s_part[2][64];                            // this CTA's partial sums (read by the whole cluster)
    __shared__ float s_red[8][2][64];
    __shared__ float s_mean[64], s_rstd[64];
    const int csize = gridDim.x;                               // cluster = all CTAs along x
    const int rank = blockIdx.x;
    const int grp = blockIdx.y;                                // MODE 0: 64-channel group
    const int b = blockIdx.z;
    const int row_pitch = (MODE == 0) ? p.C : 64;              // elements between consecutive rows in global memory
    const size_t base = (size_t)b * p.slab_stride + (MODE == 0 ? (size_t)grp * 64 : 0);
    const int r0 = rank * p.rows_per_cta;
    const int nrows = min(p.rows_per_cta, p.rows - r0);
    pdl_wait();
    // ---- load rows into shared memory
    for (int i = threadIdx.x; i < nrows * 8; i += 256) {
        const int r = i >> 3, c = i & 7;
        cp_async16(tile + (size_t)r * 128 + c * 16, p.x + base + (size_t)(r0 + r) * row_pitch + c * 8, true);
    }
    cp_async_commit();
    // operands of the apply phase that do not depend on the statistics are fetched now, under the reduction's latency:
    // the per-sample vector (MODE 0) and, when a thread owns at most 4 items, its share of the skip tensor
    __shared__ float s_vec[64];
    if (MODE == 0 && threadIdx.x < 64) s_vec[threadIdx.x] = p.vec ? p.vec[(size_t)b * p.vec_stride + grp * 64 + threadIdx.x] : 0.f;
    const bool pre_add = p.add != nullptr && nrows * 8 <= 4 * 256;
    uint4 addv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        addv[k] = make_uint4(0, 0, 0, 0);
        const int i = threadIdx.x + k * 256;
        if (pre_add && i < nrows * 8)
            addv[k] = *reinterpret_cast<const uint4*>(p.add + base + (size_t)(r0 + (i >> 3)) * row_pitch + (i & 7) * 8);
    }
    cp_async_wait<0>();
    __syncthreads();
    // ---- partial statistics (per channel of the 64-wide row; MODE 1 folds the 64 columns afterwards)
    {
        const int cp = threadIdx.x & 31, pl = threadIdx.x >> 5;
        float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
        for (int r = pl; r < nrows; r += 8) {
            const float2 v = __half22float2(*reinterpret_cast<const f162*>(tile + (size_t)r * 128 + cp * 4));
            a0 += v.x; a1 += v.y;
            q0 = fmaf(v.x, v.x, q0); q1 = fmaf(v.y, v.y, q1);
        }
        s_red[pl][0][2 * cp] = a0; s_red[pl][0][2 * cp + 1] = a1;
        s_red[pl][1][2 * cp] = q0; s_red[pl][1][2 * cp + 1] = q1;
    }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int k = threadIdx.x >> 6, c = threadIdx.x & 63;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += s_red[i][k][c];
        s_part[k][c] = s;
    }
    // ---- cluster-wide reduction in rank order through distributed shared memory
    cluster_arrive_release();
    cluster_wait_acquire();
    if (threadIdx.x < 128) {
        const int k = threadIdx.x >> 6, c = threadIdx.x & 63;
        float pv[8];                                              // all ranks' partials in flight at once, summed in rank order
#pragma unroll
        for (int r = 0; r < 8; ++r) pv[r] = (r < csize) ? ld_dsmem_f32(&s_part[k][c], r) : 0.f;
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += pv[r];
        s_red[0][k][c] = s;
    }
    __syncthreads();
    if (MODE == 0) {
        if (threadIdx.x < 64) {
            const float inv = 1.0f / (float)p.rows;
            const float mean = s_red[0][0][threadIdx.x] * inv;
            const float var = fmaxf(s_red[0][1][threadIdx.x] * inv - mean * mean, 0.f);
            s_mean[threadIdx.x] = mean;
            s_rstd[threadIdx.x] = rsqrtf(var + 1e-5f);
        }
    } else {
        if (threadIdx.x == 0) {
            double ts = 0.0, tq = 0.0;
            for (int c = 0; c < 64; ++c) { ts += (double)s_red[0][0][c]; tq += (double)s_red[0][1][c]; }
            const double cnt = (double)p.rows * 64.0;
            const double mean = ts / cnt;
            const double var = fmax(tq / cnt - mean * mean, 0.0);
            s_mean[0] = (float)mean;
            s_rstd[0] = (float)(1.0 / sqrt(var + 1e-5));
        }
    }
    __syncthreads();
    // no CTA may exit (and release its shared memory) while a peer can still read its partials
    cluster_arrive_release();
    // ---- apply from shared memory, four items per thread at a time: the skip / residual vectors of all four are requested
    // before any arithmetic (one dependent global load per iteration kept a single 16-byte load in flight per thread:
    // ncu long-scoreboard stall 7.0, 2.9 TB/s)
    const int nitems = nrows * 8;
    float satm = 0.f;
    for (int i0 = threadIdx.x, k0 = 0; i0 < nitems; i0 += 4 * 256, k0 += 4) {
        uint4 av[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 256;
            av[u] = make_uint4(0, 0, 0, 0);
            if (p.add && i < nitems) {
                if (pre_add) av[u] = addv[u];                    // nitems <= 4 * 256: k0 == 0
                else av[u] = *reinterpret_cast<const uint4*>(p.add + base + (size_t)(r0 + (i >> 3)) * row_pitch + (i & 7) * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 256;
            if (i >= nitems) break;
            const int r = i >> 3, c8 = (i & 7) * 8;
            const uint4 xv = *reinterpret_cast<const uint4*>(tile + (size_t)r * 128 + c8 * 2);
            const size_t e = base + (size_t)(r0 + r) * row_pitch + c8;
            float f[8];
            float2 t;
            t = unpack_h2(xv.x); f[0] = t.x; f[1] = t.y;
            t = unpack_h2(xv.y); f[2] = t.x; f[3] = t.y;
            t = unpack_h2(xv.z); f[4] = t.x; f[5] = t.y;
            t = unpack_h2(xv.w); f[6] = t.x; f[7] = t.y;
            int ch;                                                 // channel of f[0] in the tensor
            if (MODE == 0) {
                ch = grp * 64 + c8;
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = (f[j] - s_mean[c8 + j]) * s_rstd[c8 + j];
            } else {
                ch = (int)(((size_t)(r0 + r) * 64 + c8) % p.C);
                const float mean = s_mean[0], rstd = s_rstd[0];
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + ch)), g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + ch + 4));
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + ch)), b1 = __ldg(reinterpret_cast<const float4*>(p.beta + ch + 4));
                const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = (f[j] - mean) * rstd * gg[j] + bb[j];
            }
            if (p.add) {
                t = unpack_h2(av[u].x); f[0] += t.x; f[1] += t.y;
                t = unpack_h2(av[u].y); f[2] += t.x; f[3] += t.y;
                t = unpack_h2(av[u].z); f[4] += t.x; f[5] += t.y;
                t = unpack_h2(av[u].w); f[6] += t.x; f[7] += t.y;
            }
            if (MODE == 1 && p.act == 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = gelu_erf(f[j]);
            }
            if (p.vec) {
                if (MODE == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] += s_vec[c8 + j];
                } else {
                    const float* vp = p.vec + (size_t)b * p.vec_stride + ch;
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] += vp[j];
                }
            }
            uint4 o;
            o.x = pack_h2_acc(f[0], f[1], satm); o.y = pack_h2_acc(f[2], f[3], satm);
            o.z = pack_h2_acc(f[4], f[5], satm); o.w = pack_h2_acc(f[6], f[7], satm);
            *reinterpret_cast<uint4*>(p.y + e) = o;
        }
    }
    sat_flush(satm);
    cluster_wait_acquire();
}

// cluster size (1,2,4,8) and rows per CTA for a slab of `rows` 128-byte rows; 0 if it does not fit 8 x 128 KB
inline int norm_fused_cluster(int rows) {
    for (int cs = 1; cs <= 8; cs *= 2)
        if ((rows + cs - 1) / cs <= NF_MAX_ROWS) return cs;
    return 0;
}

template <int MODE>
inline int norm_fused_launch(const NormParams& p, int B, int groups, cudaStream_t st) {
    int cs = norm_fused_cluster(p.rows);
    B2D_CHECK(cs > 0, "normalisation slab does not fit a cluster");
    // small batches: spread each slab over a wider cluster until the grid covers the SMs about twice (the kernel is a
    // load -> reduce -> cluster barrier -> store chain per CTA, so its duration follows rows per CTA, not total bytes)
    while (cs < 8 && cs * groups * B < 2 * 148 && p.rows / (2 * cs) >= 32) cs *= 2;
    NormParams q = p;
    q.rows_per_cta = (p.rows + cs - 1) / cs;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs, groups, B);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = (size_t)q.rows_per_cta * 128;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = g_pdl_enabled;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    B2D_CUDA(cudaLaunchKernelEx(&cfg, norm_fused_kernel<MODE>, q));
    return 0;
}

inline int norm_fused_init_attrs() {
    B2D_CUDA(cudaFuncSetAttribute(norm_fused_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, NF_MAX_ROWS * 128));
    B2D_CUDA(cudaFuncSetAttribute(norm_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, NF_MAX_ROWS * 128));
    return 0;
}

}  // namespace b2d
