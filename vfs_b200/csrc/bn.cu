// Train-mode BatchNorm pieces around the tcgen05 conv: per-channel statistics of an fp32 tensor, finalisation
// (mean / invstd / running-stat update, exactly F.batch_norm(training=True)'s bookkeeping) and the normalise +
// residual + ReLU + split pass.
#include "host_common.h"
#include "ptx.cuh"

namespace vfs {

// x fp32 [M, C] (C multiple of 4): block = 256 threads covering C/4 float4 columns x row groups; the row groups of a
// block are folded in shared memory so that one fp64 atomic per (block, channel, moment) reaches global memory.
__global__ void __launch_bounds__(256) channel_stats_kernel(const float* __restrict__ x, double* __restrict__ stats,
                                                            long long M, int C) {
  __shared__ float red[256][9];
  const int cols_here = min(1024, C - static_cast<int>(blockIdx.y) * 1024);  // this block's column range
  const int c4 = cols_here / 4;
  const int lcol = threadIdx.x % c4;
  const int col = lcol + blockIdx.y * 256;              // float4 column (global)
  const int rgrp = threadIdx.x / c4;                    // row group inside the block
  const int rows_per_block = blockDim.x / c4;
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (rgrp < rows_per_block) {
    for (long long r = static_cast<long long>(blockIdx.x) * rows_per_block + rgrp; r < M;
         r += static_cast<long long>(gridDim.x) * rows_per_block) {
      const float4 v = *reinterpret_cast<const float4*>(x + r * C + col * 4);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      q[0] = fmaf(v.x, v.x, q[0]); q[1] = fmaf(v.y, v.y, q[1]);
      q[2] = fmaf(v.z, v.z, q[2]); q[3] = fmaf(v.w, v.w, q[3]);
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    red[threadIdx.x][e] = s[e];
    red[threadIdx.x][4 + e] = q[e];
  }
  __syncthreads();
  if (rgrp == 0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      double ts = 0.0, tq = 0.0;
      for (int r = 0; r < rows_per_block; ++r) {
        ts += static_cast<double>(red[r * c4 + lcol][e]);
        tq += static_cast<double>(red[r * c4 + lcol][4 + e]);
      }
      atomicAdd(stats + col * 4 + e, ts);
      atomicAdd(stats + C + col * 4 + e, tq);
    }
  }
}

__global__ void bn_finalize_kernel(double* __restrict__ stats, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float momentum, float eps,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ save_mean, float* __restrict__ save_invstd, int C) {
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = stats[c] / count;
  double var = stats[C + c] / count - mean * mean;  // biased variance, fp64 so the subtraction is benign
  if (var < 0.0) var = 0.0;
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
  const float sc = g * invstd;
  scale[c] = sc;
  shift[c] = b - static_cast<float>(mean) * sc;
  if (save_mean) save_mean[c] = static_cast<float>(mean);
  if (save_invstd) save_invstd[c] = invstd;
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
    running_var[c] = (1.0f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}

// y = z*scale + shift (+res) (relu) -> split; one thread per 8 channels
__global__ void bn_apply_kernel(const float* __restrict__ z, const h16* __restrict__ z_hi,
                                const h16* __restrict__ z_lo, const float* __restrict__ scale,
                                const float* __restrict__ shift, const h16* __restrict__ res_hi,
                                const h16* __restrict__ res_lo, h16* __restrict__ out_hi,
                                h16* __restrict__ out_lo, long long total8, int C, int relu) {
  pdl_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long o = i * 8;
    const int c = static_cast<int>(o % C);
    float4 a, b;
    if (z != nullptr) {
      a = *reinterpret_cast<const float4*>(z + o);
      b = *reinterpret_cast<const float4*>(z + o + 4);
    } else {   // raw conv output kept as a split tensor (train-mode forward through the TMA epilogue)
      const uint4 zh = *reinterpret_cast<const uint4*>(z_hi + o), zl = *reinterpret_cast<const uint4*>(z_lo + o);
      a = make_float4(lo16_to_float(zh.x) + lo16_to_float(zl.x), hi16_to_float(zh.x) + hi16_to_float(zl.x),
                      lo16_to_float(zh.y) + lo16_to_float(zl.y), hi16_to_float(zh.y) + hi16_to_float(zl.y));
      b = make_float4(lo16_to_float(zh.z) + lo16_to_float(zl.z), hi16_to_float(zh.z) + hi16_to_float(zl.z),
                      lo16_to_float(zh.w) + lo16_to_float(zl.w), hi16_to_float(zh.w) + hi16_to_float(zl.w));
    }
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c)), s1 = __ldg(reinterpret_cast<const float4*>(scale + c + 4));
    const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift + c)), h1 = __ldg(reinterpret_cast<const float4*>(shift + c + 4));
    float y[8] = {fmaf(a.x, s0.x, h0.x), fmaf(a.y, s0.y, h0.y), fmaf(a.z, s0.z, h0.z), fmaf(a.w, s0.w, h0.w),
                  fmaf(b.x, s1.x, h1.x), fmaf(b.y, s1.y, h1.y), fmaf(b.z, s1.z, h1.z), fmaf(b.w, s1.w, h1.w)};
    if (res_hi) {
      const uint4 rh = *reinterpret_cast<const uint4*>(res_hi + o), rl = *reinterpret_cast<const uint4*>(res_lo + o);
      y[0] += lo16_to_float(rh.x) + lo16_to_float(rl.x);
      y[1] += hi16_to_float(rh.x) + hi16_to_float(rl.x);
      y[2] += lo16_to_float(rh.y) + lo16_to_float(rl.y);
      y[3] += hi16_to_float(rh.y) + hi16_to_float(rl.y);
      y[4] += lo16_to_float(rh.z) + lo16_to_float(rl.z);
      y[5] += hi16_to_float(rh.z) + hi16_to_float(rl.z);
      y[6] += lo16_to_float(rh.w) + lo16_to_float(rl.w);
      y[7] += hi16_to_float(rh.w) + hi16_to_float(rl.w);
    }
    if (relu) {
#pragma unroll
      for (int e = 0; e < 8; ++e) y[e] = fmaxf(y[e], 0.0f);
    }
    h16 hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split16(y[e], hi[e], lo[e]);
    uint4 oh, ol;
    oh.x = pack16x2(hi[0], hi[1]); oh.y = pack16x2(hi[2], hi[3]);
    oh.z = pack16x2(hi[4], hi[5]); oh.w = pack16x2(hi[6], hi[7]);
    ol.x = pack16x2(lo[0], lo[1]); ol.y = pack16x2(lo[2], lo[3]);
    ol.z = pack16x2(lo[4], lo[5]); ol.w = pack16x2(lo[6], lo[7]);
    *reinterpret_cast<uint4*>(out_hi + o) = oh;
    *reinterpret_cast<uint4*>(out_lo + o) = ol;
  }
}

// y = y*scale[c] + shift[c] (+ReLU) in place on fp32 [M, C] (BatchNorm1d apply of the two-phase / SyncBN path)
__global__ void affine_act_kernel(float* __restrict__ y, const float* __restrict__ scale,
                                  const float* __restrict__ shift, long long total, int C, int relu) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    float v = fmaf(y[i], scale[c], shift[c]);
    if (relu) v = fmaxf(v, 0.0f);
    y[i] = v;
  }
}

int affine_act_f32(float* y, const float* scale, const float* shift, long long M, int C, int relu, cudaStream_t s) {
  VFS_REQUIRE(y && scale && shift, VFS_EINVAL, "affine_act: null argument");
  const long long total = M * C;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  affine_act_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(y, scale, shift, total, C, relu);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int channel_stats_f32(const float* x, double* stats, long long M, int C, cudaStream_t s) {
  VFS_REQUIRE(x && stats, VFS_EINVAL, "channel_stats: null argument");
  VFS_REQUIRE(M > 0 && C > 0 && C % 4 == 0 && (C <= 1024 || C % 1024 == 0), VFS_ESHAPE,
              "channel_stats: C=%d unsupported (multiple of 4, and of 1024 above 1024)", C);
  const int c4 = (C < 1024 ? C : 1024) / 4;
  const int rows_per_block = 256 / c4;
  long long blocks = (M + rows_per_block - 1) / rows_per_block;
  if (blocks > 148 * 4) blocks = 148 * 4;
  channel_stats_kernel<<<dim3(static_cast<int>(blocks), (C + 1023) / 1024), 256, 0, s>>>(x, stats, M, C);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int bn_finalize(double* stats, double count, const float* gamma, const float* beta, float* running_mean,
                float* running_var, float momentum, float eps, float* scale, float* shift, float* save_mean,
                float* save_invstd, int C, cudaStream_t s) {
  VFS_REQUIRE(stats && scale && shift, VFS_EINVAL, "bn_finalize: null argument");
  VFS_REQUIRE(count > 1.0, VFS_ESHAPE, "bn_finalize: Expected more than 1 value per channel when training");
  VFS_CUDA_OK(launch_pdl(bn_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, s, stats, count, gamma, beta,
                         running_mean, running_var, momentum, eps, scale, shift, save_mean, save_invstd, C));
  return VFS_OK;
}

int bn_apply(const float* z, const void* z_split, const float* scale, const float* shift, const void* residual_split,
             void* out_split, long long M, int C, int relu, cudaStream_t s) {
  VFS_REQUIRE((z || z_split) && scale && shift && out_split, VFS_EINVAL, "bn_apply: null argument");
  VFS_REQUIRE(M > 0 && C > 0 && C % 8 == 0, VFS_ESHAPE, "bn_apply: C=%d must be a multiple of 8", C);
  const long long total8 = M * C / 8;
  const h16* rh = reinterpret_cast<const h16*>(residual_split);
  h16* oh = reinterpret_cast<h16*>(out_split);
  long long blocks = (total8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const h16* zh = z ? nullptr : reinterpret_cast<const h16*>(z_split);
  VFS_CUDA_OK(launch_pdl(bn_apply_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, s, z, zh,
                         zh ? zh + M * C : nullptr, scale, shift, rh, rh ? rh + M * C : nullptr, oh, oh + M * C, total8,
                         C, relu));
  return VFS_OK;
}

VFS_DEFINE_OVERFLOW_ACCESSOR(overflow_bn)

}  // namespace vfs
