// Forward-process, loss and evaluation-statistics kernels around the sampler — fused, fp32, HBM-bound:
//   noise_image_kernel      q(x_t | x_0) of DiffusionUtils.noiseImage (diffusion_DANRA_conditional.py:85-103): one pass that writes
//                           x_t AND the noise it used (given, or Philox keyed by the global sample index)
//   weighted_mse_*          SDFWeightedMSELoss.forward (training_DANRA_conditional.py:33-56) / nn.MSELoss as ONE reduction pass
//                           (sigmoid weights evaluated in flight), deterministic two-level sum
//   eval_daily_kernel       per-sample nan-aware MAE / RMSE over the spatial dimensions (evaluation_DANRA_conditional.py:121-122)
//   eval_pixel_kernel       per-pixel MAE / RMSE / bias over the samples (the "pixel-wise" statistics of the same script)
//   histogram_kernel        fixed-range histogram (numpy semantics: NaN and out-of-range dropped, right edge closed) with
//                           shared-memory privatised bins
#pragma once
#include "elementwise.cuh"

namespace b2d {

__global__ void __launch_bounds__(256) noise_image_kernel(const float* __restrict__ x0, const long long* __restrict__ t,
                                                          const float* __restrict__ alpha_hat, const float* __restrict__ noise_in,
                                                          float* __restrict__ x_t, float* __restrict__ noise_out, size_t n,
                                                          size_t per_sample, unsigned long long seed, unsigned long long sample_offset,
                                                          float noise_scale) {
    const size_t n4 = n >> 2;
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < n4; v += (size_t)gridDim.x * blockDim.x) {
        const size_t e = v * 4;
        const size_t b = e / per_sample;
        const float ah = alpha_hat[t[b]];
        const float ca = __fsqrt_rn(ah), cn = __fsqrt_rn(__fsub_rn(1.0f, ah));
        const float4 xv = reinterpret_cast<const float4*>(x0)[v];
        float4 zv;
        if (noise_in) {
            zv = reinterpret_cast<const float4*>(noise_in)[v];
        } else {
            const unsigned long long sample = sample_offset + b;
            const unsigned long long within = (e % per_sample) >> 2;
            uint32_t c[4] = {(uint32_t)within, 0x7FFFFFFEu, (uint32_t)sample, (uint32_t)(sample >> 32)};
            philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
            const float2 g0 = box_muller(c[0], c[1]), g1 = box_muller(c[2], c[3]);
            zv = make_float4(g0.x, g0.y, g1.x, g1.y);
        }
        zv = make_float4(__fmul_rn(zv.x, noise_scale), __fmul_rn(zv.y, noise_scale), __fmul_rn(zv.z, noise_scale), __fmul_rn(zv.w, noise_scale));
        float4 o;   // (sqrt(ahat) * x) + (sqrt(1 - ahat) * noise), reference op order, no FMA contraction
        o.x = __fadd_rn(__fmul_rn(ca, xv.x), __fmul_rn(cn, zv.x));
        o.y = __fadd_rn(__fmul_rn(ca, xv.y), __fmul_rn(cn, zv.y));
        o.z = __fadd_rn(__fmul_rn(ca, xv.z), __fmul_rn(cn, zv.z));
        o.w = __fadd_rn(__fmul_rn(ca, xv.w), __fmul_rn(cn, zv.w));
        reinterpret_cast<float4*>(x_t)[v] = o;
        reinterpret_cast<float4*>(noise_out)[v] = zv;
    }
}

// block-level sum of doubles in fixed order (warp shuffles, then warp 0)
__device__ __forceinline__ double block_sum_d(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = l < (int)(blockDim.x >> 5) ? sh[l] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;
}

constexpr int WMSE_BLOCKS = 592;
__global__ void __launch_bounds__(256) weighted_mse_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                   const float* __restrict__ sdf, float w_span, float w_min,
                                                                   double* __restrict__ partial, size_t n) {
    __shared__ double sh[8];
    double acc = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float d = a[i] - b[i];
        float w = 1.0f;
        if (sdf) w = (1.0f / (1.0f + expf(-sdf[i]))) * w_span + w_min;   // sigmoid(sdf) * (max_land - min_sea) + min_sea
        acc += (double)(w * (d * d));
    }
    const double s = block_sum_d(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) weighted_mse_final_kernel(const double* __restrict__ partial, int nblocks, double inv_n,
                                                                 float* __restrict__ out) {
    __shared__ double sh[8];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) acc += partial[i];
    const double s = block_sum_d(acc, sh);
    if (threadIdx.x == 0) out[0] = (float)(s * inv_n);
}

// one block per sample: nan-aware mean |g - e| and sqrt(mean (g - e)^2) over the sample's hw pixels
__global__ void __launch_bounds__(256) eval_daily_kernel(const float* __restrict__ gen, const float* __restrict__ ev,
                                                         float* __restrict__ mae, float* __restrict__ rmse, size_t hw) {
    __shared__ double sh[8];
    const size_t base = (size_t)blockIdx.x * hw;
    double sa = 0.0, ss = 0.0, cnt = 0.0;
    for (size_t i = threadIdx.x; i < hw; i += blockDim.x) {
        const float d = gen[base + i] - ev[base + i];
        if (d == d) {
            sa += fabs((double)d);
            ss += (double)d * (double)d;
            cnt += 1.0;
        }
    }
    sa = block_sum_d(sa, sh);
    ss = block_sum_d(ss, sh);
    cnt = block_sum_d(cnt, sh);
    if (threadIdx.x == 0) {
        mae[blockIdx.x] = cnt > 0 ? (float)(sa / cnt) : __int_as_float(0x7fc00000);
        rmse[blockIdx.x] = cnt > 0 ? (float)sqrt(ss / cnt) : __int_as_float(0x7fc00000);
    }
}

// one thread per pixel, loop over the samples (coalesced across pixels): nan-aware MAE, RMSE and bias (mean of g - e)
__global__ void __launch_bounds__(256) eval_pixel_kernel(const float* __restrict__ gen, const float* __restrict__ ev,
                                                         float* __restrict__ mae, float* __restrict__ rmse, float* __restrict__ bias,
                                                         int n_samples, size_t hw) {
    const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (p >= hw) return;
    double sa = 0.0, ss = 0.0, sb = 0.0;
    int cnt = 0;
    for (int s = 0; s < n_samples; ++s) {
        const float d = gen[(size_t)s * hw + p] - ev[(size_t)s * hw + p];
        if (d == d) {
            sa += fabs((double)d);
            ss += (double)d * (double)d;
            sb += (double)d;
            ++cnt;
        }
    }
    const float nanv = __int_as_float(0x7fc00000);
    mae[p] = cnt ? (float)(sa / cnt) : nanv;
    rmse[p] = cnt ? (float)sqrt(ss / cnt) : nanv;
    bias[p] = cnt ? (float)(sb / cnt) : nanv;
}

constexpr int HIST_MAX_BINS = 4096;
__global__ void __launch_bounds__(256) histogram_kernel(const float* __restrict__ x, size_t n, float lo, float hi, int bins,
                                                        unsigned long long* __restrict__ counts) {
    extern __shared__ unsigned int sh_bins[];
    for (int i = threadIdx.x; i < bins; i += blockDim.x) sh_bins[i] = 0;
    __syncthreads();
    // bin index in double, like numpy ((x - first_edge) * bins / (last_edge - first_edge) on float64): with thousands of bins an
    // fp32 product misplaces every value within ~1e-4 of an edge
    const double scale = (double)bins / ((double)hi - (double)lo);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        if (v >= lo && v <= hi) {                       // NaN fails both comparisons
            int b = (int)(((double)v - (double)lo) * scale);
            if (b >= bins) b = bins - 1;                // right edge belongs to the last bin
            atomicAdd(&sh_bins[b], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += blockDim.x)
        if (sh_bins[i]) atomicAdd(&counts[i], (unsigned long long)sh_bins[i]);
}

}  // namespace b2d
