// Backward-pass kernels around the tensor-core dgrad/wgrad convolutions: BatchNorm(+ReLU) backward on split NHWC
// activations, stem (max-pool / BN / 7x7 weight gradient), SimSiam head (Linear, BatchNorm1d, avg-pool, cosine loss)
// and the fused SGD-momentum update.  All exact fp32 (fp64 for cross-pixel reductions), fixed formulas of
// torch.nn.functional.batch_norm / max_pool2d / linear backward.
#include <math.h>
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"

namespace vfs {

__device__ __forceinline__ void unpack8(const uint4 h, const uint4 l, float (&v)[8]) {
  v[0] = lo16_to_float(h.x) + lo16_to_float(l.x);
  v[1] = hi16_to_float(h.x) + hi16_to_float(l.x);
  v[2] = lo16_to_float(h.y) + lo16_to_float(l.y);
  v[3] = hi16_to_float(h.y) + hi16_to_float(l.y);
  v[4] = lo16_to_float(h.z) + lo16_to_float(l.z);
  v[5] = hi16_to_float(h.z) + hi16_to_float(l.z);
  v[6] = lo16_to_float(h.w) + lo16_to_float(l.w);
  v[7] = hi16_to_float(h.w) + hi16_to_float(l.w);
}
__device__ __forceinline__ void pack8(const float (&v)[8], uint4& h, uint4& l) {
  h16 hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) split16(v[e], hi[e], lo[e]);
  h.x = pack16x2(hi[0], hi[1]); h.y = pack16x2(hi[2], hi[3]);
  h.z = pack16x2(hi[4], hi[5]); h.w = pack16x2(hi[6], hi[7]);
  l.x = pack16x2(lo[0], lo[1]); l.y = pack16x2(lo[2], lo[3]);
  l.z = pack16x2(lo[4], lo[5]); l.w = pack16x2(lo[6], lo[7]);
}

// ------------------------------------------------------------------------------------------------
// BatchNorm (+ReLU, +residual fan-out) backward on [M, C] activations.
//   g  = dY * 1[y > 0]            (relu; y = forward output of the layer)
//   reduce: sums[c] = sum g, sums[C + c] = sum g * xhat,  xhat = (z - mean) * invstd
//   apply : dz = gamma * invstd * (g - sums[c]/M - xhat * sums[C+c]/M)     -> split;  optionally also g -> split
// ------------------------------------------------------------------------------------------------
struct BnBwdArgs {
  const h16* dy_hi; const h16* dy_lo;  // split dY (or null when dy_f32 is used)
  const float* dy_f32;
  const h16* y_hi; const h16* y_lo;    // forward output for the ReLU mask (null: no ReLU)
  const float* y_f32;                  // same as fp32 (SimSiam head tensors)
  const float* z;                                          // raw conv output fp32 [M, C] (or null when z_hi is set)
  const h16* z_hi; const h16* z_lo;                        // raw conv output as a split tensor
  const float* mean; const float* invstd; const float* gamma;
  double* sums;                                            // [2C] (reduce: accumulated; apply: read)
  double count;
  h16* dz_hi; h16* dz_lo;              // split dz (or null)
  float* dz_f32;                                           // fp32 dz (stem)
  h16* g_hi; h16* g_lo;                // optional: masked gradient for the residual branch
  float* dgamma; float* dbeta;         // optional (apply): parameter gradients = param_scale * sums, written by block row 0
  int accumulate; float param_scale;
  int eval_mode;                       // BN ran on running statistics: dz = gamma * invstd * g (no batch-statistic terms)
  long long M; int C;
};

__device__ __forceinline__ void load_z8(const BnBwdArgs& a, long long o, float (&zz)[8]) {
  if (a.z) {
    const float4 z0 = *reinterpret_cast<const float4*>(a.z + o), z1 = *reinterpret_cast<const float4*>(a.z + o + 4);
    zz[0] = z0.x; zz[1] = z0.y; zz[2] = z0.z; zz[3] = z0.w; zz[4] = z1.x; zz[5] = z1.y; zz[6] = z1.z; zz[7] = z1.w;
  } else {
    unpack8(*reinterpret_cast<const uint4*>(a.z_hi + o), *reinterpret_cast<const uint4*>(a.z_lo + o), zz);
  }
}

__device__ __forceinline__ void bn_bwd_load_g(const BnBwdArgs& a, long long o, float (&g)[8]) {
  if (a.dy_f32) {
    const float4 p = *reinterpret_cast<const float4*>(a.dy_f32 + o), q = *reinterpret_cast<const float4*>(a.dy_f32 + o + 4);
    g[0] = p.x; g[1] = p.y; g[2] = p.z; g[3] = p.w; g[4] = q.x; g[5] = q.y; g[6] = q.z; g[7] = q.w;
  } else {
    unpack8(*reinterpret_cast<const uint4*>(a.dy_hi + o), *reinterpret_cast<const uint4*>(a.dy_lo + o), g);
  }
  if (a.y_hi) {
    float y[8];
    unpack8(*reinterpret_cast<const uint4*>(a.y_hi + o), *reinterpret_cast<const uint4*>(a.y_lo + o), y);
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = (y[e] > 0.0f) ? g[e] : 0.0f;
  } else if (a.y_f32) {
    const float4 p = *reinterpret_cast<const float4*>(a.y_f32 + o), q = *reinterpret_cast<const float4*>(a.y_f32 + o + 4);
    const float y[8] = {p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = (y[e] > 0.0f) ? g[e] : 0.0f;
  }
}

// ReLU mask from the forward output: y = hi + lo with hi = half(y) >= 0, so y > 0 <=> hi > 0 except for 0 < y < 2^-25
// (hi rounds to zero; such a pixel's gradient is dropped -- immeasurable, and it saves the lo plane's 2 bytes/element).
__device__ __forceinline__ void bn_bwd_load_g2(const BnBwdArgs& a, long long o, float (&g)[8]) {
  if (a.dy_f32) {
    const float4 p = *reinterpret_cast<const float4*>(a.dy_f32 + o), q = *reinterpret_cast<const float4*>(a.dy_f32 + o + 4);
    g[0] = p.x; g[1] = p.y; g[2] = p.z; g[3] = p.w; g[4] = q.x; g[5] = q.y; g[6] = q.z; g[7] = q.w;
  } else {
    unpack8(*reinterpret_cast<const uint4*>(a.dy_hi + o), *reinterpret_cast<const uint4*>(a.dy_lo + o), g);
  }
  if (a.y_hi) {
    const uint4 h = *reinterpret_cast<const uint4*>(a.y_hi + o);
    const float y[8] = {lo16_to_float(h.x), hi16_to_float(h.x), lo16_to_float(h.y), hi16_to_float(h.y),
                        lo16_to_float(h.z), hi16_to_float(h.z), lo16_to_float(h.w), hi16_to_float(h.w)};
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = (y[e] > 0.0f) ? g[e] : 0.0f;
  } else if (a.y_f32) {
    const float4 p = *reinterpret_cast<const float4*>(a.y_f32 + o), q = *reinterpret_cast<const float4*>(a.y_f32 + o + 4);
    const float y[8] = {p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = (y[e] > 0.0f) ? g[e] : 0.0f;
  }
}

// Channel-slab layout shared by the reduce and apply kernels: a block of kBnThreads threads covers a slab of
// `slab` <= 512 channels (blockIdx.y) as slab/8 eight-channel pieces x row groups; every thread keeps ITS eight channels
// for the whole kernel (row stride = whole rows), so the per-channel constants live in registers and every warp access
// is a run of whole 128-byte lines.  Two rows per loop iteration keep ~12 16-byte loads in flight per thread.
constexpr int kBnThreads = 512;

// Per-channel sums of g and g * xhat over all M pixels: row groups of a block are folded in shared memory, one fp64
// atomic per (block, channel, sum) reaches global memory.  (History: 16 fp64 atomics per THREAD = 310 us per call;
// 256-thread blocks without unrolling and a serial 8-thread fold = 2.4 TB/s, profiles/r02_bn_bench_v1.log.)
__global__ void __launch_bounds__(kBnThreads) bn_bwd_reduce_kernel(const BnBwdArgs a, int slab) {
  __shared__ float red[kBnThreads][17];
  pdl_wait();
  const int pieces = slab / 8;
  const int piece = threadIdx.x % pieces, rgrp = threadIdx.x / pieces;
  const int rows_per_block = kBnThreads / pieces;
  const bool active = rgrp < rows_per_block;
  const int c = blockIdx.y * slab + piece * 8;
  float sg[8], sx[8], mu[8], is[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sg[e] = sx[e] = 0.0f;
    mu[e] = a.mean[c + e];
    is[e] = a.invstd[c + e];
  }
  if (active) {
    const long long stride = static_cast<long long>(gridDim.x) * rows_per_block;
    long long r = static_cast<long long>(blockIdx.x) * rows_per_block + rgrp;
    for (; r + stride < a.M; r += 2 * stride) {
      const long long o0 = r * a.C + c, o1 = (r + stride) * a.C + c;
      float g0[8], g1[8];
      bn_bwd_load_g2(a, o0, g0);
      bn_bwd_load_g2(a, o1, g1);
      float z0[8], z1[8];
      load_z8(a, o0, z0);
      load_z8(a, o1, z1);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        sg[e] += g0[e];
        sx[e] = fmaf(g0[e], (z0[e] - mu[e]) * is[e], sx[e]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        sg[e] += g1[e];
        sx[e] = fmaf(g1[e], (z1[e] - mu[e]) * is[e], sx[e]);
      }
    }
    if (r < a.M) {
      const long long o0 = r * a.C + c;
      float g0[8], z0[8];
      bn_bwd_load_g2(a, o0, g0);
      load_z8(a, o0, z0);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        sg[e] += g0[e];
        sx[e] = fmaf(g0[e], (z0[e] - mu[e]) * is[e], sx[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    red[threadIdx.x][e] = sg[e];
    red[threadIdx.x][8 + e] = sx[e];
  }
  __syncthreads();
  // value v of the block's 2 * slab outputs: which = v / slab (0: sum g, 1: sum g xhat), channel cc = v % slab
  for (int v = threadIdx.x; v < 2 * slab; v += kBnThreads) {
    const int which = v / slab, cc = v - which * slab;
    const int pc = cc >> 3, e = (cc & 7) + 8 * which;
    double t = 0.0;
    for (int rg = 0; rg < rows_per_block; ++rg) t += static_cast<double>(red[rg * pieces + pc][e]);
    atomicAdd(a.sums + which * a.C + blockIdx.y * slab + cc, t);
  }
}

__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(const BnBwdArgs a, int slab) {
  pdl_wait();
  const int pieces = slab / 8;
  const int piece = threadIdx.x % pieces, rgrp = threadIdx.x / pieces;
  const int rows_per_block = kBnThreads / pieces;
  if (rgrp >= rows_per_block) return;
  const int c = blockIdx.y * slab + piece * 8;
  const float inv_count = static_cast<float>(1.0 / a.count);
  float mu[8], is[8], ga[8], sgm[8], sxm[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    mu[e] = a.mean[c + e];
    is[e] = a.invstd[c + e];
    ga[e] = (a.gamma ? a.gamma[c + e] : 1.0f) * is[e];
    sgm[e] = a.eval_mode ? 0.0f : static_cast<float>(a.sums[c + e]) * inv_count;
    sxm[e] = a.eval_mode ? 0.0f : static_cast<float>(a.sums[a.C + c + e]) * inv_count;
  }
  if (blockIdx.x == 0 && rgrp == 0 && (a.dgamma || a.dbeta) && a.sums != nullptr) {
    // dgamma = param_scale * sum g xhat, dbeta = param_scale * sum g (one writer per channel; formerly a third launch)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float dg = static_cast<float>(a.sums[a.C + c + e]) * a.param_scale;
      const float db = static_cast<float>(a.sums[c + e]) * a.param_scale;
      if (a.dgamma) a.dgamma[c + e] = a.accumulate ? a.dgamma[c + e] + dg : dg;
      if (a.dbeta) a.dbeta[c + e] = a.accumulate ? a.dbeta[c + e] + db : db;
    }
  }
  const long long stride = static_cast<long long>(gridDim.x) * rows_per_block;
  auto emit = [&](long long o, const float (&g)[8], const float (&zz)[8]) {
    float dz[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float xhat = (zz[e] - mu[e]) * is[e];
      dz[e] = ga[e] * (g[e] - sgm[e] - xhat * sxm[e]);
    }
    if (a.dz_hi) {
      uint4 h, l;
      pack8(dz, h, l);
      *reinterpret_cast<uint4*>(a.dz_hi + o) = h;
      *reinterpret_cast<uint4*>(a.dz_lo + o) = l;
    }
    if (a.dz_f32) {
      *reinterpret_cast<float4*>(a.dz_f32 + o) = make_float4(dz[0], dz[1], dz[2], dz[3]);
      *reinterpret_cast<float4*>(a.dz_f32 + o + 4) = make_float4(dz[4], dz[5], dz[6], dz[7]);
    }
    if (a.g_hi) {
      uint4 h, l;
      pack8(g, h, l);
      *reinterpret_cast<uint4*>(a.g_hi + o) = h;
      *reinterpret_cast<uint4*>(a.g_lo + o) = l;
    }
  };
  long long r = static_cast<long long>(blockIdx.x) * rows_per_block + rgrp;
  for (; r + stride < a.M; r += 2 * stride) {
    const long long o0 = r * a.C + c, o1 = (r + stride) * a.C + c;
    float g0[8], g1[8], z0[8], z1[8];
    bn_bwd_load_g2(a, o0, g0);
    bn_bwd_load_g2(a, o1, g1);
    load_z8(a, o0, z0);
    load_z8(a, o1, z1);
    emit(o0, g0, z0);
    emit(o1, g1, z1);
  }
  if (r < a.M) {
    const long long o0 = r * a.C + c;
    float g0[8], z0[8];
    bn_bwd_load_g2(a, o0, g0);
    load_z8(a, o0, z0);
    emit(o0, g0, z0);
  }
}

// sums fp64 [2C] -> dgamma/dbeta fp32 (optionally accumulated)
__global__ void bn_bwd_param_kernel(const double* __restrict__ sums, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int C, int accumulate, float out_scale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float dg = static_cast<float>(sums[C + c]) * out_scale, db = static_cast<float>(sums[c]) * out_scale;
  if (dgamma) dgamma[c] = accumulate ? dgamma[c] + dg : dg;
  if (dbeta) dbeta[c] = accumulate ? dbeta[c] + db : db;
}

static int bn_bwd_fill(BnBwdArgs& a, const void* dy_split, const float* dy_f32, const void* y_split, const float* y_f32,
                       const float* z, const void* z_split,
                       const float* mean, const float* invstd, const float* gamma, double* sums, double count,
                       void* dz_split, float* dz_f32, void* g_split, long long M, int C) {
  memset(&a, 0, sizeof(a));
  const long long plane = M * C;
  if (dy_split) {
    a.dy_hi = reinterpret_cast<const h16*>(dy_split);
    a.dy_lo = a.dy_hi + plane;
  }
  a.dy_f32 = dy_f32;
  if (y_split) {
    a.y_hi = reinterpret_cast<const h16*>(y_split);
    a.y_lo = a.y_hi + plane;
  }
  a.y_f32 = y_f32;
  a.z = z; a.mean = mean; a.invstd = invstd; a.gamma = gamma; a.sums = sums; a.count = count;
  if (!z && z_split) {
    a.z_hi = reinterpret_cast<const h16*>(z_split);
    a.z_lo = a.z_hi + plane;
  }
  if (dz_split) {
    a.dz_hi = reinterpret_cast<h16*>(dz_split);
    a.dz_lo = a.dz_hi + plane;
  }
  a.dz_f32 = dz_f32;
  if (g_split) {
    a.g_hi = reinterpret_cast<h16*>(g_split);
    a.g_lo = a.g_hi + plane;
  }
  a.M = M; a.C = C;
  return VFS_OK;
}

// bulk-copy pipelined forms (csrc/bn_stream.cu) for the residual-stage layout: split dy / z / y, split outputs
bool bn_stream_eligible(const void* dy_split, const float* dy_f32, const float* y_f32, const float* z, const void* z_split,
                        float* dz_f32, long long M, int C);
int bn_bwd_reduce_stream(const void* dy_split, const void* y_split, const void* z_split, const float* mean,
                         const float* invstd, double* sums, long long M, int C, cudaStream_t s);
int bn_bwd_apply_stream(const void* dy_split, const void* y_split, const void* z_split, const float* mean,
                        const float* invstd, const float* gamma, const double* sums, double count, void* dz_split,
                        void* g_split, float* dgamma, float* dbeta, int accumulate, float param_scale, int eval_mode,
                        long long M, int C, cudaStream_t s);

// channel slab of a block: the whole C when C <= 512 (C/8 must divide kBnThreads), else 512-channel slabs
static bool bn_slab(int C, int* slab) {
  if (C <= 0 || C % 8 != 0) return false;
  if (C <= 512) {
    if (kBnThreads % (C / 8) != 0) return false;
    *slab = C;
    return true;
  }
  if (C % 512 != 0) return false;
  *slab = 512;
  return true;
}
// grid.x blocks walk the rows (each thread at least ~4 rows, at most 2 blocks per SM in total), grid.y the channel slabs
static dim3 bn_grid(long long M, int C, int slab) {
  const int rows_per_block = kBnThreads / (slab / 8);
  const int slabs = C / slab;
  long long bx = (M + 4LL * rows_per_block - 1) / (4LL * rows_per_block);
  const long long cap = (2LL * 148 + slabs - 1) / slabs;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3(static_cast<unsigned>(bx), static_cast<unsigned>(slabs));
}

int bn_bwd_reduce(const void* dy_split, const float* dy_f32, const void* y_split, const float* y_f32, const float* z,
                  const void* z_split, const float* mean, const float* invstd, double* sums, long long M, int C,
                  cudaStream_t s) {
  VFS_REQUIRE((dy_split || dy_f32) && (z || z_split) && mean && invstd && sums, VFS_EINVAL,
              "bn_bwd_reduce: null argument");
  if (M > 0 && bn_stream_eligible(dy_split, dy_f32, y_f32, z, z_split, nullptr, M, C))
    return bn_bwd_reduce_stream(dy_split, y_split, z_split, mean, invstd, sums, M, C, s);
  int slab = 0;
  VFS_REQUIRE(M > 0 && bn_slab(C, &slab), VFS_ESHAPE, "bn_bwd_reduce: C=%d unsupported", C);
  BnBwdArgs a;
  bn_bwd_fill(a, dy_split, dy_f32, y_split, y_f32, z, z_split, mean, invstd, nullptr, sums, 1.0, nullptr, nullptr,
              nullptr, M, C);
  const dim3 grid = bn_grid(M, C, slab);
  VFS_CUDA_OK(launch_pdl(bn_bwd_reduce_kernel, dim3(grid), dim3(kBnThreads), 0, s, a, slab));
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int bn_bwd_apply(const void* dy_split, const float* dy_f32, const void* y_split, const float* y_f32, const float* z,
                 const void* z_split, const float* mean, const float* invstd, const float* gamma, const double* sums,
                 double count, void* dz_split,
                 float* dz_f32, void* g_split, float* dgamma, float* dbeta, int accumulate, float param_scale,
                 long long M, int C, cudaStream_t s) {
  // count <= 0 selects the eval-mode form (running statistics): sums may then be NULL when no parameter gradient is
  // wanted
  const bool eval_mode = !(count > 0);
  if (eval_mode) count = 1.0;
  VFS_REQUIRE((dy_split || dy_f32) && (z || z_split) && mean && invstd && (sums || eval_mode) && (dz_split || dz_f32),
              VFS_EINVAL, "bn_bwd_apply: null argument");
  if (M > 0 && dz_split && bn_stream_eligible(dy_split, dy_f32, y_f32, z, z_split, dz_f32, M, C))
    return bn_bwd_apply_stream(dy_split, y_split, z_split, mean, invstd, gamma, sums, count, dz_split, g_split, dgamma,
                               dbeta, accumulate, param_scale, eval_mode ? 1 : 0, M, C, s);
  int slab = 0;
  VFS_REQUIRE(M > 0 && count > 0 && bn_slab(C, &slab), VFS_ESHAPE, "bn_bwd_apply: C=%d unsupported", C);
  BnBwdArgs a;
  bn_bwd_fill(a, dy_split, dy_f32, y_split, y_f32, z, z_split, mean, invstd, gamma, const_cast<double*>(sums), count,
              dz_split, dz_f32, g_split, M, C);
  a.dgamma = dgamma;
  a.dbeta = dbeta;
  a.accumulate = accumulate;
  a.param_scale = param_scale;
  a.eval_mode = eval_mode ? 1 : 0;
  const dim3 grid = bn_grid(M, C, slab);
  VFS_CUDA_OK(launch_pdl(bn_bwd_apply_kernel, dim3(grid), dim3(kBnThreads), 0, s, a, slab));
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

// masked gradient only (no BN): g = dY * 1[y > 0]  -> split   (blocks whose last op is add+ReLU with eval-mode BN)
__global__ void relu_bwd_split_kernel(const h16* dy_hi, const h16* dy_lo, const h16* y_hi,
                                      const h16* y_lo, h16* g_hi, h16* g_lo,
                                      long long total8) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long o = i * 8;
    float g[8], y[8];
    unpack8(*reinterpret_cast<const uint4*>(dy_hi + o), *reinterpret_cast<const uint4*>(dy_lo + o), g);
    unpack8(*reinterpret_cast<const uint4*>(y_hi + o), *reinterpret_cast<const uint4*>(y_lo + o), y);
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = (y[e] > 0.0f) ? g[e] : 0.0f;
    uint4 h, l;
    pack8(g, h, l);
    *reinterpret_cast<uint4*>(g_hi + o) = h;
    *reinterpret_cast<uint4*>(g_lo + o) = l;
  }
}

// ------------------------------------------------------------------------------------------------
// Stem backward: max-pool(3,2,1) + ReLU backward, then 7x7 weight gradient.
// ------------------------------------------------------------------------------------------------
// g[n,y,x,c] = 1[a > 0] * sum over pooling windows whose (first) arg-max is (y,x) of dPool, with a = relu(z*sc+sh)
__global__ void stem_pool_relu_bwd_kernel(const h16* __restrict__ dp_hi, const h16* __restrict__ dp_lo,
                                          const float* __restrict__ z, const float* __restrict__ scale,
                                          const float* __restrict__ shift, float* __restrict__ g, int N, int Hc, int Wc,
                                          int Hp, int Wp) {
  const size_t total = static_cast<size_t>(N) * Hc * Wc * 8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int grp = static_cast<int>(i & 7);
    size_t t = i >> 3;
    const int x = static_cast<int>(t % Wc);
    t /= Wc;
    const int y = static_cast<int>(t % Hc);
    const int n = static_cast<int>(t / Hc);
    float sc[8], sh[8], acc[8], a0[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sc[e] = scale[grp * 8 + e];
      sh[e] = shift[grp * 8 + e];
      acc[e] = 0.0f;
    }
    auto act = [&](int yy, int xx, float (&out)[8]) {
      const float4* src = reinterpret_cast<const float4*>(z + ((static_cast<size_t>(n) * Hc + yy) * Wc + xx) * 64 + grp * 8);
      const float4 p = src[0], q = src[1];
      const float v[8] = {p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) out[e] = fmaxf(fmaf(v[e], sc[e], sh[e]), 0.0f);
    };
    act(y, x, a0);
    // pooling windows (py, px) containing this pixel: 2*py-1 <= y <= 2*py+1  <=>  y/2 <= py <= (y+1)/2
    const int py_lo = y / 2, py_hi = (y + 1) / 2;
    const int px_lo = x / 2, px_hi = (x + 1) / 2;
    for (int py = py_lo; py <= py_hi; ++py) {
      if (py >= Hp) continue;
      for (int px = px_lo; px <= px_hi; ++px) {
        if (px >= Wp) continue;
        // is (y, x) the first maximum of window (py, px) in row-major scan order?
        bool first[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) first[e] = true;
        for (int dy = -1; dy <= 1; ++dy) {
          const int yy = 2 * py + dy;
          if (yy < 0 || yy >= Hc) continue;
          for (int dx = -1; dx <= 1; ++dx) {
            const int xx = 2 * px + dx;
            if (xx < 0 || xx >= Wc || (yy == y && xx == x)) continue;
            float o[8];
            act(yy, xx, o);
            const bool before = (yy < y) || (yy == y && xx < x);
#pragma unroll
            for (int e = 0; e < 8; ++e) first[e] = first[e] && (before ? (o[e] < a0[e]) : (o[e] <= a0[e]));
          }
        }
        float d[8];
        const size_t po = ((static_cast<size_t>(n) * Hp + py) * Wp + px) * 64 + grp * 8;
        unpack8(*reinterpret_cast<const uint4*>(dp_hi + po), *reinterpret_cast<const uint4*>(dp_lo + po), d);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += first[e] ? d[e] : 0.0f;
      }
    }
    float out[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) out[e] = (a0[e] > 0.0f) ? acc[e] : 0.0f;
    float* dst = g + ((static_cast<size_t>(n) * Hc + y) * Wc + x) * 64 + grp * 8;
    *reinterpret_cast<float4*>(dst) = make_float4(out[0], out[1], out[2], out[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(out[4], out[5], out[6], out[7]);
  }
}

// dW[co][c][r][s] += sum_{n,oy,ox} dz[n,oy,ox,co] * x[n,c,2oy+r-3,2ox+s-3].
// Block = one segment of <= 32 conv pixels of one output row per iteration (grid-stride); the 7 input rows the segment
// touches and its dz vectors are staged in shared memory.  Thread = 4 output channels x 10 filter taps (k = kg + 16 j),
// i.e. 40 accumulators: per pixel one 16-byte load of dz and 10 broadcast loads of the input patch feed 40 FMAs -- the
// first version held 1 x 37 accumulators and issued two shared-memory loads per FMA (1.1 ms per call at 64 x 224^2,
// profiles/r02_train_profile_cfg4_v1.log).  Partial sums are combined with fp32 atomics.
constexpr int kStemSegMax = 32;
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                                         float* __restrict__ dw, int N, int H, int W, int Hc, int Wc,
                                                         int seg_w, float out_scale) {
  __shared__ float patch[3 * 7 * 72];              // [c][r][pc], pc < 2*seg_w + 5 <= 69 (row pitch 72)
  __shared__ __align__(16) float dzs[kStemSegMax][64];
  const int cg = threadIdx.x & 15, kg = threadIdx.x >> 4;   // co = 4 cg .. 4 cg + 3 ; k = kg + 16 j
  int poff[10];                                             // patch offset of tap j (without the 2*px term)
  bool kok[10];
#pragma unroll
  for (int j = 0; j < 10; ++j) {
    const int k = kg + 16 * j;
    kok[j] = k < 147;
    const int kk = kok[j] ? k : 0;
    const int c = kk / 49, r = (kk / 7) % 7, sx = kk % 7;
    poff[j] = (c * 7 + r) * 72 + sx;
  }
  float acc[10][4];
#pragma unroll
  for (int j = 0; j < 10; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.0f;
  const int segs_per_row = (Wc + seg_w - 1) / seg_w;
  const int pw = 2 * seg_w + 5;
  const long long total_segs = static_cast<long long>(N) * Hc * segs_per_row;
  for (long long sidx = blockIdx.x; sidx < total_segs; sidx += gridDim.x) {
    const int seg = static_cast<int>(sidx % segs_per_row);
    const long long t = sidx / segs_per_row;
    const int oy = static_cast<int>(t % Hc);
    const int n = static_cast<int>(t / Hc);
    const int ox0 = seg * seg_w;
    __syncthreads();
    for (int i = threadIdx.x; i < 21 * pw; i += 256) {
      const int pc = i % pw, cr = i / pw;            // cr = c * 7 + r
      const int c = cr / 7, r = cr - 7 * c;
      const int iy = 2 * oy + r - 3, ix = 2 * ox0 + pc - 3;
      patch[cr * 72 + pc] = (iy >= 0 && iy < H && ix >= 0 && ix < W)
                                ? __ldg(x + (static_cast<size_t>(n) * 3 + c) * H * W + static_cast<size_t>(iy) * W + ix)
                                : 0.0f;
    }
    const float4* dz4 = reinterpret_cast<const float4*>(dz + ((static_cast<size_t>(n) * Hc + oy) * Wc + ox0) * 64);
    for (int i = threadIdx.x; i < seg_w * 16; i += 256) {
      const int px = i >> 4;
      reinterpret_cast<float4*>(&dzs[0][0])[i] = (ox0 + px < Wc) ? __ldg(dz4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
#pragma unroll 2
    for (int px = 0; px < seg_w; ++px) {
      const float4 d = *reinterpret_cast<const float4*>(&dzs[px][4 * cg]);
#pragma unroll
      for (int j = 0; j < 10; ++j) {
        const float pv = patch[poff[j] + 2 * px];
        acc[j][0] = fmaf(d.x, pv, acc[j][0]);
        acc[j][1] = fmaf(d.y, pv, acc[j][1]);
        acc[j][2] = fmaf(d.z, pv, acc[j][2]);
        acc[j][3] = fmaf(d.w, pv, acc[j][3]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 10; ++j) {
    if (!kok[j]) continue;
    const int k = kg + 16 * j;
#pragma unroll
    for (int e = 0; e < 4; ++e) atomicAdd(dw + (4 * cg + e) * 147 + k, acc[j][e] * out_scale);
  }
}

// Max-pool(3,2,1) + ReLU backward as a SCATTER over the pooled grid: thread = (pool position, 8 channels) finds the
// first maximum of its 3x3 window of a = relu(z*scale + shift) in row-major scan order (torch's arg-max rule) and adds
// its gradient there when a > 0 (the ReLU mask).  g must be zero-filled by the caller.  The gather form above visits up
// to 36 activations per conv pixel; this one reads 9 per pool position (a quarter as many positions).
__global__ void stem_pool_relu_bwd_scatter_kernel(const h16* __restrict__ dp_hi, const h16* __restrict__ dp_lo,
                                                  const float* __restrict__ z, const float* __restrict__ scale,
                                                  const float* __restrict__ shift, float* __restrict__ g, int N, int Hc,
                                                  int Wc, int Hp, int Wp) {
  const size_t total = static_cast<size_t>(N) * Hp * Wp * 8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int grp = static_cast<int>(i & 7);
    size_t t = i >> 3;
    const int px = static_cast<int>(t % Wp);
    t /= Wp;
    const int py = static_cast<int>(t % Hp);
    const int n = static_cast<int>(t / Hp);
    float sc[8], sh[8], best[8];
    int arg[8];
    *reinterpret_cast<float4*>(sc) = __ldg(reinterpret_cast<const float4*>(scale + grp * 8));
    *reinterpret_cast<float4*>(sc + 4) = __ldg(reinterpret_cast<const float4*>(scale + grp * 8 + 4));
    *reinterpret_cast<float4*>(sh) = __ldg(reinterpret_cast<const float4*>(shift + grp * 8));
    *reinterpret_cast<float4*>(sh + 4) = __ldg(reinterpret_cast<const float4*>(shift + grp * 8 + 4));
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      best[e] = -1.0f;     // activations are >= 0: the first visited position always wins over the sentinel
      arg[e] = 0;
    }
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = 2 * py + dy;
      if (yy < 0 || yy >= Hc) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = 2 * px + dx;
        if (xx < 0 || xx >= Wc) continue;
        const float4* src =
            reinterpret_cast<const float4*>(z + ((static_cast<size_t>(n) * Hc + yy) * Wc + xx) * 64 + grp * 8);
        const float4 p = src[0], q = src[1];
        const float v[8] = {p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w};
        const int code = (dy + 1) * 3 + (dx + 1);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float a = fmaxf(fmaf(v[e], sc[e], sh[e]), 0.0f);
          if (a > best[e]) {
            best[e] = a;
            arg[e] = code;
          }
        }
      }
    }
    float d[8];
    const size_t po = ((static_cast<size_t>(n) * Hp + py) * Wp + px) * 64 + grp * 8;
    unpack8(*reinterpret_cast<const uint4*>(dp_hi + po), *reinterpret_cast<const uint4*>(dp_lo + po), d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (best[e] > 0.0f && d[e] != 0.0f) {
        const int yy = 2 * py + arg[e] / 3 - 1, xx = 2 * px + arg[e] % 3 - 1;
        atomicAdd(g + ((static_cast<size_t>(n) * Hc + yy) * Wc + xx) * 64 + grp * 8 + e, d[e]);
      }
    }
  }
}

int stem_pool_relu_bwd(const void* dpool_split, const float* z, const float* scale, const float* shift, float* g,
                       int N, int H, int W, cudaStream_t s) {
  VFS_REQUIRE(dpool_split && z && scale && shift && g, VFS_EINVAL, "stem_pool_relu_bwd: null argument");
  const int Hc = (H + 6 - 7) / 2 + 1, Wc = (W + 6 - 7) / 2 + 1;
  const int Hp = (Hc + 2 - 3) / 2 + 1, Wp = (Wc + 2 - 3) / 2 + 1;
  const h16* hi = reinterpret_cast<const h16*>(dpool_split);
  static int gather = -1;   // VFS_STEM_POOL_BWD_GATHER=1 selects the deterministic gather form (debug / comparison)
  if (gather < 0) {
    const char* e = getenv("VFS_STEM_POOL_BWD_GATHER");
    gather = (e && atoi(e) != 0) ? 1 : 0;
  }
  if (gather) {
    const size_t total = static_cast<size_t>(N) * Hc * Wc * 8;
    const int blocks = static_cast<int>((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    stem_pool_relu_bwd_kernel<<<blocks, 256, 0, s>>>(hi, hi + static_cast<size_t>(N) * Hp * Wp * 64, z, scale, shift, g,
                                                     N, Hc, Wc, Hp, Wp);
  } else {
    VFS_CUDA_OK(cudaMemsetAsync(g, 0, static_cast<size_t>(N) * Hc * Wc * 64 * sizeof(float), s));
    const size_t total = static_cast<size_t>(N) * Hp * Wp * 8;
    const int blocks = static_cast<int>((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    stem_pool_relu_bwd_scatter_kernel<<<blocks, 256, 0, s>>>(hi, hi + static_cast<size_t>(N) * Hp * Wp * 64, z, scale,
                                                             shift, g, N, Hc, Wc, Hp, Wp);
  }
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int stem_wgrad_tc(const float* x, const float* dz, float* dw, float out_scale, int N, int H, int W, cudaStream_t s);

int stem_wgrad(const float* x, const float* dz, float* dw, int accumulate, float out_scale, int N, int H, int W,
               cudaStream_t s) {
  VFS_REQUIRE(x && dz && dw, VFS_EINVAL, "stem_wgrad: null argument");
  const int Hc = (H + 6 - 7) / 2 + 1, Wc = (W + 6 - 7) / 2 + 1;
  if (!accumulate) VFS_CUDA_OK(cudaMemsetAsync(dw, 0, 64 * 147 * sizeof(float), s));
  static int simt = -1;
  if (simt < 0) {
    const char* e = getenv("VFS_STEM_WGRAD_SIMT");
    simt = (e && e[0] == '1') ? 1 : 0;
  }
  if (!simt) return stem_wgrad_tc(x, dz, dw, out_scale, N, H, W, s);   // csrc/stem.cu: tcgen05 version
  const int segs = (Wc + kStemSegMax - 1) / kStemSegMax;
  const int seg_w = (Wc + segs - 1) / segs;          // e.g. Wc = 112 -> 4 segments of 28 pixels
  stem_wgrad_kernel<<<148 * 3, 256, 0, s>>>(x, dz, dw, N, H, W, Hc, Wc, seg_w, out_scale);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

// ------------------------------------------------------------------------------------------------
// SimSiam head backward
// ------------------------------------------------------------------------------------------------
// dX[m, k] = sum_n dY[m, n] W[n, k]; block = 256 consecutive k, rows tiled by 16, the n range split over blockIdx.z
// in slices of 128 (the weight matrix is the whole traffic: 8 blocks cannot pull it through; dX is zeroed by the caller
// and the slices are combined with fp32 atomics).
__global__ void __launch_bounds__(256) linear_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ W,
                                                              float* __restrict__ dx, int M, int N, int K) {
  __shared__ float dys[16][128];
  const int k = blockIdx.x * 256 + threadIdx.x;
  const int m0 = blockIdx.y * 16;
  float acc[16];
#pragma unroll
  for (int m = 0; m < 16; ++m) acc[m] = 0.0f;
  {
    const int n0 = blockIdx.z * 128;
    __syncthreads();
    for (int i = threadIdx.x; i < 16 * 128; i += 256) {
      const int m = i >> 7, nn = i & 127;
      dys[m][nn] = (m0 + m < M && n0 + nn < N) ? dy[static_cast<size_t>(m0 + m) * N + n0 + nn] : 0.0f;
    }
    __syncthreads();
    if (k < K) {
      const int nmax = min(128, N - n0);
      for (int nn = 0; nn < nmax; ++nn) {
        const float w = __ldg(W + static_cast<size_t>(n0 + nn) * K + k);
#pragma unroll
        for (int m = 0; m < 16; ++m) acc[m] = fmaf(dys[m][nn], w, acc[m]);
      }
    }
  }
  if (k < K) {
#pragma unroll
    for (int m = 0; m < 16; ++m)
      if (m0 + m < M) atomicAdd(dx + static_cast<size_t>(m0 + m) * K + k, acc[m]);
  }
}

// dW[n, k] (+)= sum_m dY[m, n] X[m, k];  block = 8 rows n x 256 k; db[n] (+)= sum_m dY[m, n]
__global__ void __launch_bounds__(256) linear_bwd_weight_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                float* __restrict__ dW, float* __restrict__ db, int M,
                                                                int N, int K, int accumulate) {
  __shared__ float dys[128][8];
  const int k = blockIdx.x * 256 + threadIdx.x;
  const int n0 = blockIdx.y * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
  float bsum = 0.0f;
  for (int m0 = 0; m0 < M; m0 += 128) {
    __syncthreads();
    for (int i = threadIdx.x; i < 128 * 8; i += 256) {
      const int m = i >> 3, j = i & 7;
      dys[m][j] = (m0 + m < M && n0 + j < N) ? dy[static_cast<size_t>(m0 + m) * N + n0 + j] : 0.0f;
    }
    __syncthreads();
    const int mmax = min(128, M - m0);
    if (k < K) {
      for (int m = 0; m < mmax; ++m) {
        const float xv = __ldg(x + static_cast<size_t>(m0 + m) * K + k);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(dys[m][j], xv, acc[j]);
      }
    }
    if (blockIdx.x == 0 && threadIdx.x < 8)
      for (int m = 0; m < mmax; ++m) bsum += dys[m][threadIdx.x];
  }
  if (k < K) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (n0 + j < N) {
        float* d = dW + static_cast<size_t>(n0 + j) * K + k;
        *d = accumulate ? *d + acc[j] : acc[j];
      }
  }
  if (db && blockIdx.x == 0 && threadIdx.x < 8 && n0 + threadIdx.x < N)
    db[n0 + threadIdx.x] = accumulate ? db[n0 + threadIdx.x] + bsum : bsum;
}

// BatchNorm1d (+ReLU) backward, one thread per feature.  pre = linear output before BN, out = layer output.
__global__ void bn1d_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ pre,
                                const float* __restrict__ out, float* __restrict__ dpre, int M, int N,
                                const float* __restrict__ gamma, const float* __restrict__ mean,
                                const float* __restrict__ invstd, int training, int relu, float* __restrict__ dgamma,
                                float* __restrict__ dbeta, int accumulate) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float mu = mean[n], is = invstd[n], g = gamma ? gamma[n] : 1.0f;
  float sg = 0.0f, sx = 0.0f;
  for (int m = 0; m < M; ++m) {
    const size_t o = static_cast<size_t>(m) * N + n;
    float d = dy[o];
    if (relu && !(out[o] > 0.0f)) d = 0.0f;
    sg += d;
    sx = fmaf(d, (pre[o] - mu) * is, sx);
  }
  const float invM = 1.0f / static_cast<float>(M);
  for (int m = 0; m < M; ++m) {
    const size_t o = static_cast<size_t>(m) * N + n;
    float d = dy[o];
    if (relu && !(out[o] > 0.0f)) d = 0.0f;
    const float xhat = (pre[o] - mu) * is;
    dpre[o] = training ? g * is * (d - sg * invM - xhat * sx * invM) : g * is * d;
  }
  if (dgamma) dgamma[n] = accumulate ? dgamma[n] + sx : sx;
  if (dbeta) dbeta[n] = accumulate ? dbeta[n] + sg : sg;
}

__global__ void relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ out, float* __restrict__ dx,
                                size_t n) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n) dx[i] = (out[i] > 0.0f) ? dy[i] : 0.0f;
}

// global average pool backward straight into the backbone's gradient format: dY [B, C] -> split NHWC [B, HW, C]
__global__ void avgpool_bwd_split_kernel(const float* __restrict__ dy, h16* __restrict__ hi,
                                         h16* __restrict__ lo, int B, int HW, int C) {
  const size_t total8 = static_cast<size_t>(B) * HW * C / 8;
  const float inv = 1.0f / static_cast<float>(HW);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t o = i * 8;
    const int c = static_cast<int>(o % C);
    const int b = static_cast<int>(o / (static_cast<size_t>(HW) * C));
    const float4 p = *reinterpret_cast<const float4*>(dy + static_cast<size_t>(b) * C + c);
    const float4 q = *reinterpret_cast<const float4*>(dy + static_cast<size_t>(b) * C + c + 4);
    const float v[8] = {p.x * inv, p.y * inv, p.z * inv, p.w * inv, q.x * inv, q.y * inv, q.z * inv, q.w * inv};
    uint4 h, l;
    pack8(v, h, l);
    *reinterpret_cast<uint4*>(hi + o) = h;
    *reinterpret_cast<uint4*>(lo + o) = l;
  }
}

// global average pool backward in the reference layout: dY [B, C] -> dX NCHW [B, C, HW]
__global__ void avgpool_bwd_nchw_kernel(const float* __restrict__ dy, float* __restrict__ dx, size_t total, int HW) {
  const float inv = 1.0f / static_cast<float>(HW);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dx[i] = dy[i / HW] * inv;
}

int avgpool_backward_nchw(const float* dy, float* dx, int B, int C, int HW, cudaStream_t s) {
  VFS_REQUIRE(dy && dx, VFS_EINVAL, "avgpool_backward_nchw: null argument");
  const size_t total = static_cast<size_t>(B) * C * HW;
  const int blocks = static_cast<int>((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  avgpool_bwd_nchw_kernel<<<blocks, 256, 0, s>>>(dy, dx, total, HW);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

// cosine loss backward w.r.t. p (z is detached in SimSiam):  L = 2 - 2 cos  ->  dp = g * (-2) (zhat - phat cos)/|p|
__global__ void __launch_bounds__(128) cosine_loss_bwd_kernel(const float* __restrict__ p, const float* __restrict__ z,
                                                              const float* __restrict__ gout, float* __restrict__ dp,
                                                              int D, int with_norm, int negative) {
  __shared__ float red[3][4];
  __shared__ float bc[3];
  const int b = blockIdx.x;
  const float* pp = p + static_cast<size_t>(b) * D;
  const float* zz = z + static_cast<size_t>(b) * D;
  float spp = 0.0f, szz = 0.0f, spz = 0.0f;
  for (int i = threadIdx.x; i < D; i += 128) {
    const float a = pp[i], c = zz[i];
    spp = fmaf(a, a, spp);
    szz = fmaf(c, c, szz);
    spz = fmaf(a, c, spz);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    spp += __shfl_xor_sync(0xffffffffu, spp, o);
    szz += __shfl_xor_sync(0xffffffffu, szz, o);
    spz += __shfl_xor_sync(0xffffffffu, spz, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = spp;
    red[1][threadIdx.x >> 5] = szz;
    red[2][threadIdx.x >> 5] = spz;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    bc[0] = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    bc[1] = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    bc[2] = red[2][0] + red[2][1] + red[2][2] + red[2][3];
  }
  __syncthreads();
  const float coef = (negative ? -1.0f : -2.0f) * gout[b];
  if (!with_norm) {
    for (int i = threadIdx.x; i < D; i += 128) dp[static_cast<size_t>(b) * D + i] = coef * zz[i];
    return;
  }
  const float np = fmaxf(sqrtf(bc[0]), 1e-12f), nz = fmaxf(sqrtf(bc[1]), 1e-12f);
  const float cosv = bc[2] / (np * nz);
  for (int i = threadIdx.x; i < D; i += 128)
    dp[static_cast<size_t>(b) * D + i] = coef * (zz[i] / nz - pp[i] / np * cosv) / np;
}

// ------------------------------------------------------------------------------------------------
// SGD with momentum and weight decay (torch.optim.SGD semantics, dampening 0, no nesterov):
//   g' = g + wd * p ;  buf = first ? g' : momentum * buf + g' ;  p -= lr * buf
// ------------------------------------------------------------------------------------------------
__global__ void sgd_momentum_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                                    size_t n, float lr, float momentum, float wd, int first, float grad_scale) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float gg = fmaf(wd, p[i], g[i] * grad_scale);
    const float b = first ? gg : fmaf(momentum, buf[i], gg);
    buf[i] = b;
    p[i] = p[i] - lr * b;
  }
}

// Same update with the hyper-parameters read from device memory ({lr, momentum, weight_decay, grad_scale}): a CUDA
// graph that captured the launch follows an LR schedule by copying four floats, no re-capture.  Momentum buffers start
// zeroed (momentum * 0 + g' = g' is torch's first-step rule).  float4-vectorised (n % 4 == 0 and 16-byte alignment
// checked by the host); the tail runs scalar.
__global__ void sgd_momentum_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                                        size_t n, const float* __restrict__ hyper, int vec) {
  const float lr = hyper[0], momentum = hyper[1], wd = hyper[2], grad_scale = hyper[3];
  if (vec) {
    const size_t n4 = n / 4;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
      float4 pv = reinterpret_cast<float4*>(p)[i];
      const float4 gv = reinterpret_cast<const float4*>(g)[i];
      float4 bv = reinterpret_cast<float4*>(buf)[i];
      bv.x = fmaf(momentum, bv.x, fmaf(wd, pv.x, gv.x * grad_scale));
      bv.y = fmaf(momentum, bv.y, fmaf(wd, pv.y, gv.y * grad_scale));
      bv.z = fmaf(momentum, bv.z, fmaf(wd, pv.z, gv.z * grad_scale));
      bv.w = fmaf(momentum, bv.w, fmaf(wd, pv.w, gv.w * grad_scale));
      pv.x -= lr * bv.x; pv.y -= lr * bv.y; pv.z -= lr * bv.z; pv.w -= lr * bv.w;
      reinterpret_cast<float4*>(buf)[i] = bv;
      reinterpret_cast<float4*>(p)[i] = pv;
    }
    return;
  }
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float gg = fmaf(wd, p[i], g[i] * grad_scale);
    const float b = fmaf(momentum, buf[i], gg);
    buf[i] = b;
    p[i] = p[i] - lr * b;
  }
}

int sgd_momentum_step_dev(float* p, const float* g, float* buf, size_t n, const float* hyper, cudaStream_t s) {
  VFS_REQUIRE(p && g && buf && hyper, VFS_EINVAL, "sgd_momentum_step_dev: null argument");
  if (n == 0) return VFS_OK;
  const bool vec = n % 4 == 0 && (reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                                  reinterpret_cast<uintptr_t>(buf)) % 16 == 0;
  const size_t work = vec ? n / 4 : n;
  size_t blocks = (work + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  sgd_momentum_dev_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(p, g, buf, n, hyper, vec ? 1 : 0);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int relu_bwd_split(const void* dy_split, const void* y_split, void* g_split, long long elems, cudaStream_t s) {
  VFS_REQUIRE(dy_split && y_split && g_split && elems % 8 == 0, VFS_EINVAL, "relu_bwd_split: bad argument");
  const h16* dh = reinterpret_cast<const h16*>(dy_split);
  const h16* yh = reinterpret_cast<const h16*>(y_split);
  h16* gh = reinterpret_cast<h16*>(g_split);
  long long blocks = (elems / 8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  relu_bwd_split_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(dh, dh + elems, yh, yh + elems, gh, gh + elems, elems / 8);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

bool linear_mma_eligible(int M, int N, int K);
bool linear_mma_enabled();
int linear_bwd_data_mma(const float* dy, const float* W, float* dx, int M, int N, int K, cudaStream_t s);
int linear_bwd_weight_mma(const float* dy, const float* x, float* dW, float* db, int M, int N, int K, int accumulate,
                          cudaStream_t s);

int linear_backward(const float* dy, const float* x, const float* W, float* dx, float* dW, float* db, int M, int N,
                    int K, int accumulate, cudaStream_t s) {
  VFS_REQUIRE(dy && x && W, VFS_EINVAL, "linear_backward: null argument");
  VFS_REQUIRE(M > 0 && N > 0 && K > 0, VFS_ESHAPE, "linear_backward: bad shape");
  const bool mma = linear_mma_enabled() && linear_mma_eligible(M, N, K);
  if (dx && mma) {
    const int rc = linear_bwd_data_mma(dy, W, dx, M, N, K, s);
    if (rc != VFS_OK) return rc;
  } else if (dx) {
    VFS_CUDA_OK(cudaMemsetAsync(dx, 0, static_cast<size_t>(M) * K * sizeof(float), s));
    linear_bwd_data_kernel<<<dim3((K + 255) / 256, (M + 15) / 16, (N + 127) / 128), 256, 0, s>>>(dy, W, dx, M, N, K);
    VFS_CUDA_OK(cudaGetLastError());
  }
  if (dW && mma && M <= 32) return linear_bwd_weight_mma(dy, x, dW, db, M, N, K, accumulate, s);
  if (dW) {
    linear_bwd_weight_kernel<<<dim3((K + 255) / 256, (N + 7) / 8), 256, 0, s>>>(dy, x, dW, db, M, N, K, accumulate);
    VFS_CUDA_OK(cudaGetLastError());
  }
  return VFS_OK;
}

int bn1d_backward(const float* dy, const float* pre, const float* out, float* dpre, int M, int N, const float* gamma,
                  const float* mean, const float* invstd, int training, int relu, float* dgamma, float* dbeta,
                  int accumulate, cudaStream_t s) {
  VFS_REQUIRE(dy && pre && out && dpre && mean && invstd, VFS_EINVAL, "bn1d_backward: null argument");
  bn1d_bwd_kernel<<<(N + 127) / 128, 128, 0, s>>>(dy, pre, out, dpre, M, N, gamma, mean, invstd, training, relu, dgamma,
                                                  dbeta, accumulate);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int relu_backward(const float* dy, const float* out, float* dx, size_t n, cudaStream_t s) {
  VFS_REQUIRE(dy && out && dx, VFS_EINVAL, "relu_backward: null argument");
  relu_bwd_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, s>>>(dy, out, dx, n);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int avgpool_backward_split(const float* dy, void* out_split, int B, int HW, int C, cudaStream_t s) {
  VFS_REQUIRE(dy && out_split && C % 8 == 0, VFS_EINVAL, "avgpool_backward: bad argument");
  h16* hi = reinterpret_cast<h16*>(out_split);
  const size_t total8 = static_cast<size_t>(B) * HW * C / 8;
  const int blocks = static_cast<int>((total8 + 255) / 256 < 148 * 16 ? (total8 + 255) / 256 : 148 * 16);
  avgpool_bwd_split_kernel<<<blocks, 256, 0, s>>>(dy, hi, hi + static_cast<size_t>(B) * HW * C, B, HW, C);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int cosine_loss_backward(const float* p, const float* z, const float* gout, float* dp, int B, int D, int with_norm,
                         int negative, cudaStream_t s) {
  VFS_REQUIRE(p && z && gout && dp, VFS_EINVAL, "cosine_loss_backward: null argument");
  cosine_loss_bwd_kernel<<<B, 128, 0, s>>>(p, z, gout, dp, D, with_norm, negative);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int sgd_momentum_step(float* p, const float* g, float* buf, size_t n, float lr, float momentum, float wd, int first,
                      float grad_scale, cudaStream_t s) {
  VFS_REQUIRE(p && g && buf, VFS_EINVAL, "sgd_momentum_step: null argument");
  if (n == 0) return VFS_OK;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  sgd_momentum_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(p, g, buf, n, lr, momentum, wd, first, grad_scale);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

VFS_DEFINE_OVERFLOW_ACCESSOR(overflow_train)

}  // namespace vfs
