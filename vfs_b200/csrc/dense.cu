// Dense HW x HW affinity post-processing (the exported helpers compute_affinity / propagate of
// mmaction/models/common/affinity_utils.py:6-50; not used by VanillaTracker, kept for API completeness).
// The affinity GEMM itself runs on the tcgen05 conv kernel (src pixels as the "image", dst pixels as the 1x1 filter
// bank); this file holds the masked softmax along either axis and the top-k-thresholded propagation.
#include <math.h>

#include "host_common.h"
#include "ptx.cuh"

namespace vfs {

struct DenseMask {
  int mode;  // 0 none, 1 circle, 2 square
  int ry, rx, W;
};

__device__ __forceinline__ bool dense_mask_ok(const DenseMask& m, int i, int j) {
  if (m.mode == 0) return true;
  const int dy = i / m.W - j / m.W, dx = i % m.W - j % m.W;
  return (m.mode == 1) ? (dy * dy + dx * dx < m.ry * m.ry) : (abs(dy) <= m.ry && abs(dx) <= m.rx);
}

// A [B][R][ld] (R rows, Cc valid columns, leading dimension ld) -> out [B][R][Cc] contiguous.
// softmax_dim: 0 none (copy + mask), 1 over rows (per column), 2 over columns (per row).
// Block (32, 8): 32 consecutive lines of the non-reduced axis, 8 groups striding the reduced axis.
__global__ void masked_softmax_kernel(const float* __restrict__ A, float* __restrict__ out, int R, int Cc, int ld,
                                      int softmax_dim, DenseMask mask, int nan_to_zero) {
  __shared__ float red[8][33];
  const int b = blockIdx.y;
  const float* src = A + static_cast<size_t>(b) * R * ld;
  float* dst = out + static_cast<size_t>(b) * R * Cc;
  const int line = blockIdx.x * 32 + threadIdx.x;  // column (dim 1) or row (dim 2) index
  const int nlines = (softmax_dim == 2) ? R : Cc;
  const int len = (softmax_dim == 2) ? Cc : R;
  auto at = [&](int l, int k) -> size_t {  // element k of line l
    return (softmax_dim == 2) ? static_cast<size_t>(l) * ld + k : static_cast<size_t>(k) * ld + l;
  };
  auto at_out = [&](int l, int k) -> size_t {
    return (softmax_dim == 2) ? static_cast<size_t>(l) * Cc + k : static_cast<size_t>(k) * Cc + l;
  };
  auto masked = [&](int l, int k) -> bool {  // mask[src i, dst j]
    return (softmax_dim == 2) ? dense_mask_ok(mask, l, k) : dense_mask_ok(mask, k, l);
  };
  const bool live = line < nlines;
  if (softmax_dim == 0) {
    if (live)
      for (int k = threadIdx.y; k < len; k += 8) {
        const float v = src[at(line, k)];
        dst[at_out(line, k)] = masked(line, k) ? v : -INFINITY;
      }
    return;
  }
  float m = -INFINITY;
  if (live)
    for (int k = threadIdx.y; k < len; k += 8)
      if (masked(line, k)) m = fmaxf(m, src[at(line, k)]);
  red[threadIdx.y][threadIdx.x] = m;
  __syncthreads();
  m = red[0][threadIdx.x];
#pragma unroll
  for (int g = 1; g < 8; ++g) m = fmaxf(m, red[g][threadIdx.x]);
  __syncthreads();
  float s = 0.0f;
  if (live)
    for (int k = threadIdx.y; k < len; k += 8)
      if (masked(line, k)) s += expf(src[at(line, k)] - m);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  s = 0.0f;
#pragma unroll
  for (int g = 0; g < 8; ++g) s += red[g][threadIdx.x];
  if (live)
    for (int k = threadIdx.y; k < len; k += 8) {
      float v = masked(line, k) ? expf(src[at(line, k)] - m) / s : 0.0f;  // exp(-inf - m) = 0 for masked entries
      if (m == -INFINITY) v = nan_to_zero ? 0.0f : NAN;                  // fully masked line: softmax gives NaN
      dst[at_out(line, k)] = v;
    }
}

// new_img[b, c, j] = sum_i img[b, c, i] * A'[b, i, j];  A' = A, or clamp(A - kth_j, 0) / max(sum, 1e-12) with kth_j the
// topk-th largest value of column j.  Block (32, 8) = 32 columns x 8 row groups.
template <int KMAX, int CVMAX>
__global__ void propagate_dense_kernel(const float* __restrict__ img, const float* __restrict__ A,
                                       float* __restrict__ out, int Cv, int HW, int topk) {
  __shared__ float sh[8][KMAX][33];
  __shared__ float red[8][CVMAX + 1][33];
  const int b = blockIdx.y;
  const int j = blockIdx.x * 32 + threadIdx.x;
  const bool live = j < HW;
  const float* Ab = A + static_cast<size_t>(b) * HW * HW;
  const float* ib = img + static_cast<size_t>(b) * Cv * HW;
  float kth = -INFINITY;
  if (topk > 0) {
    float tv[KMAX];
#pragma unroll
    for (int i = 0; i < KMAX; ++i) tv[i] = -INFINITY;
    if (live)
      for (int i = threadIdx.y; i < HW; i += 8) {
        const float x = Ab[static_cast<size_t>(i) * HW + j];
        if (x > tv[KMAX - 1]) {
#pragma unroll
          for (int t = KMAX - 1; t >= 0; --t) {
            if (t > 0 && x > tv[t - 1]) tv[t] = tv[t - 1];
            else if (x > tv[t]) tv[t] = x;
          }
        }
      }
#pragma unroll
    for (int t = 0; t < KMAX; ++t) sh[threadIdx.y][t][threadIdx.x] = tv[t];
    __syncthreads();
    // merge the 8 sorted lists of this column (every thread of the column redundantly; 8*KMAX values)
    float mv[KMAX];
#pragma unroll
    for (int t = 0; t < KMAX; ++t) mv[t] = -INFINITY;
    for (int g = 0; g < 8; ++g)
#pragma unroll
      for (int t = 0; t < KMAX; ++t) {
        const float x = sh[g][t][threadIdx.x];
        if (x > mv[KMAX - 1]) {
#pragma unroll
          for (int u = KMAX - 1; u >= 0; --u) {
            if (u > 0 && x > mv[u - 1]) mv[u] = mv[u - 1];
            else if (x > mv[u]) mv[u] = x;
          }
        }
      }
    kth = -INFINITY;
#pragma unroll
    for (int t = 0; t < KMAX; ++t)
      if (t == topk - 1) kth = mv[t];
  }
  float acc[CVMAX + 1];
#pragma unroll
  for (int c = 0; c <= CVMAX; ++c) acc[c] = 0.0f;
  if (live)
    for (int i = threadIdx.y; i < HW; i += 8) {
      float a = Ab[static_cast<size_t>(i) * HW + j];
      if (topk > 0) a = fmaxf(a - kth, 0.0f);
      acc[CVMAX] += a;
#pragma unroll
      for (int c = 0; c < CVMAX; ++c)
        if (c < Cv) acc[c] = fmaf(ib[static_cast<size_t>(c) * HW + i], a, acc[c]);
    }
#pragma unroll
  for (int c = 0; c <= CVMAX; ++c) red[threadIdx.y][c][threadIdx.x] = acc[c];
  __syncthreads();
  if (threadIdx.y == 0 && live) {
    float tot[CVMAX + 1];
#pragma unroll
    for (int c = 0; c <= CVMAX; ++c) {
      tot[c] = 0.0f;
      for (int g = 0; g < 8; ++g) tot[c] += red[g][c][threadIdx.x];
    }
    const float inv = (topk > 0) ? 1.0f / fmaxf(tot[CVMAX], 1e-12f) : 1.0f;
#pragma unroll
    for (int c = 0; c < CVMAX; ++c)
      if (c < Cv) out[(static_cast<size_t>(b) * Cv + c) * HW + j] = tot[c] * inv;
  }
}

int masked_softmax(const float* A, float* out, int B, int R, int Cc, int ld, int softmax_dim, int mask_mode, int ry,
                   int rx, int W, int nan_to_zero, cudaStream_t s) {
  VFS_REQUIRE(A && out, VFS_EINVAL, "masked_softmax: null argument");
  VFS_REQUIRE(B > 0 && R > 0 && Cc > 0 && ld >= Cc && softmax_dim >= 0 && softmax_dim <= 2, VFS_ESHAPE,
              "masked_softmax: bad shape");
  VFS_REQUIRE(mask_mode == 0 || (R == Cc && W > 0), VFS_ESHAPE, "masked_softmax: the analytic mask needs a square matrix");
  DenseMask m{mask_mode, ry, rx, W};
  const int nlines = (softmax_dim == 2) ? R : Cc;
  masked_softmax_kernel<<<dim3((nlines + 31) / 32, B), dim3(32, 8), 0, s>>>(A, out, R, Cc, ld, softmax_dim, m,
                                                                            nan_to_zero);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int propagate_dense(const float* img, const float* A, float* out, int B, int Cv, int HW, int topk, cudaStream_t s) {
  VFS_REQUIRE(img && A && out, VFS_EINVAL, "propagate_dense: null argument");
  VFS_REQUIRE(B > 0 && HW > 0 && Cv >= 1 && Cv <= 16, VFS_ESHAPE, "propagate_dense: Cv=%d outside [1,16]", Cv);
  VFS_REQUIRE(topk >= 0 && topk <= 16 && topk <= HW, VFS_ESHAPE, "propagate_dense: topk=%d outside [0,16]", topk);
  propagate_dense_kernel<16, 16><<<dim3((HW + 31) / 32, B), dim3(32, 8), 0, s>>>(img, A, out, Cv, HW, topk);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}


// ------------------------------------------------------------------------------------------------------------------
// General form of masked_attention_efficient (common/local_attention.py:287-342) for the cases the fused window kernel
// (affinity.cu) does not take: an ARBITRARY boolean mask tensor and / or topk = None (softmax over every key).  The
// affinity A [rows = T * HWk][ld >= HWq] of one batch item (already divided by the temperature) comes from the tcgen05
// conv kernel (key pixels as the image, query pixels as the 1x1 filter bank); this kernel does, per query column,
//   mask -> top-k (or all keys) -> softmax | clamp(min=0)^2 -> weighted sum of the value vectors.
// mask: uint8 [HWk][HWq] (key-major like the reference asserts), applied to key frames t >= non_mask_len; NULL = none.
// Block (32, 8): 32 consecutive queries x 8 row groups.
// ------------------------------------------------------------------------------------------------------------------
template <int KMAX>
__global__ void generic_attention_topk_kernel(const float* __restrict__ A, int rows, int ld, int HWk, int HWq,
                                              const unsigned char* __restrict__ mask, int non_mask_len,
                                              const float* __restrict__ values, int Cv, int topk, int mode,
                                              float* __restrict__ out) {
  __shared__ float sv[8][KMAX][33];
  __shared__ int si[8][KMAX][33];
  const int q = blockIdx.x * 32 + threadIdx.x;
  const bool live = q < HWq;
  float tv[KMAX];
  int ti[KMAX];
#pragma unroll
  for (int i = 0; i < KMAX; ++i) {
    tv[i] = -INFINITY;
    ti[i] = -1;
  }
  auto insert = [&](float x, int idx) {
#pragma unroll
    for (int i = KMAX - 1; i >= 0; --i) {
      if (i > 0 && x > tv[i - 1]) {
        tv[i] = tv[i - 1];
        ti[i] = ti[i - 1];
      } else if (x > tv[i]) {
        tv[i] = x;
        ti[i] = idx;
      }
    }
  };
  // contiguous row ranges per group: ties resolve towards the lower key index like a serial scan
  const int per = (rows + 7) / 8;
  const int r_end = min(rows, (static_cast<int>(threadIdx.y) + 1) * per);
  if (live)
    for (int r = threadIdx.y * per; r < r_end; ++r) {
      if (mask != nullptr && r / HWk >= non_mask_len && mask[static_cast<size_t>(r % HWk) * HWq + q] == 0) continue;
      const float x = A[static_cast<size_t>(r) * ld + q];
      if (x > tv[KMAX - 1]) insert(x, r);
    }
#pragma unroll
  for (int i = 0; i < KMAX; ++i) {
    sv[threadIdx.y][i][threadIdx.x] = tv[i];
    si[threadIdx.y][i][threadIdx.x] = ti[i];
  }
  __syncthreads();
  if (threadIdx.y != 0 || !live) return;
  for (int g = 1; g < 8; ++g)
#pragma unroll 1
    for (int i = 0; i < KMAX; ++i) {
      const float x = sv[g][i][threadIdx.x];
      if (!(x > tv[KMAX - 1])) break;
      insert(x, si[g][i][threadIdx.x]);
    }
  float w[KMAX];
  float wsum = 0.0f;
#pragma unroll
  for (int i = 0; i < KMAX; ++i) {
    if (i < topk && ti[i] >= 0) {
      w[i] = (mode == 0) ? expf(tv[i] - tv[0]) : fmaxf(tv[i], 0.0f) * fmaxf(tv[i], 0.0f);
      wsum += w[i];
    } else {
      w[i] = 0.0f;
    }
  }
  const float inv = (mode == 0) ? __fdiv_rn(1.0f, wsum) : 1.0f;
  for (int c = 0; c < Cv; ++c) {
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < KMAX; ++i)
      if (i < topk && ti[i] >= 0) acc = fmaf(values[static_cast<size_t>(c) * rows + ti[i]], w[i] * inv, acc);
    out[static_cast<size_t>(c) * HWq + q] = acc;
  }
}

// topk = None: softmax (or clamp^2) over ALL unmasked keys.  Two passes over the column (max, then weights), value
// channels in groups of 8 accumulators.
__global__ void generic_attention_dense_kernel(const float* __restrict__ A, int rows, int ld, int HWk, int HWq,
                                               const unsigned char* __restrict__ mask, int non_mask_len,
                                               const float* __restrict__ values, int Cv, int mode,
                                               float* __restrict__ out) {
  __shared__ float red[8][9][33];
  const int q = blockIdx.x * 32 + threadIdx.x;
  const bool live = q < HWq;
  auto valid = [&](int r) {
    return mask == nullptr || r / HWk < non_mask_len || mask[static_cast<size_t>(r % HWk) * HWq + q] != 0;
  };
  float m = -INFINITY;
  if (live && mode == 0)
    for (int r = threadIdx.y; r < rows; r += 8)
      if (valid(r)) m = fmaxf(m, A[static_cast<size_t>(r) * ld + q]);
  red[threadIdx.y][0][threadIdx.x] = m;
  __syncthreads();
#pragma unroll
  for (int g = 0; g < 8; ++g) m = fmaxf(m, red[g][0][threadIdx.x]);
  __syncthreads();
  for (int c0 = 0; c0 < Cv; c0 += 8) {
    float acc[8], wsum = 0.0f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
    if (live)
      for (int r = threadIdx.y; r < rows; r += 8) {
        if (!valid(r)) continue;
        const float x = A[static_cast<size_t>(r) * ld + q];
        const float w = (mode == 0) ? expf(x - m) : fmaxf(x, 0.0f) * fmaxf(x, 0.0f);
        wsum += w;
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (c0 + e < Cv) acc[e] = fmaf(values[static_cast<size_t>(c0 + e) * rows + r], w, acc[e]);
      }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[threadIdx.y][e][threadIdx.x] = acc[e];
    red[threadIdx.y][8][threadIdx.x] = wsum;
    __syncthreads();
    if (threadIdx.y == 0 && live) {
      float tot[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) {
        tot[e] = 0.0f;
        for (int g = 0; g < 8; ++g) tot[e] += red[g][e][threadIdx.x];
      }
      const float inv = (mode == 0) ? __fdiv_rn(1.0f, tot[8]) : 1.0f;
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (c0 + e < Cv) out[static_cast<size_t>(c0 + e) * HWq + q] = tot[e] * inv;
    }
    __syncthreads();
  }
}

int generic_attention(const float* A, int rows, int ld, int HWk, int HWq, const unsigned char* mask, int non_mask_len,
                      const float* values, int Cv, int topk, int mode, float* out, cudaStream_t s) {
  VFS_REQUIRE(A && values && out, VFS_EINVAL, "generic_attention: null argument");
  VFS_REQUIRE(rows > 0 && HWk > 0 && rows % HWk == 0 && HWq > 0 && ld >= HWq && Cv >= 1, VFS_ESHAPE,
              "generic_attention: bad shape (rows %d, HWk %d, HWq %d, ld %d)", rows, HWk, HWq, ld);
  VFS_REQUIRE(topk >= 0 && topk <= 16, VFS_ESHAPE, "generic_attention: topk=%d outside [1,16] (0 = all keys)", topk);
  VFS_REQUIRE(mode == 0 || mode == 1, VFS_EINVAL, "generic_attention: bad mode");
  VFS_REQUIRE(non_mask_len >= 0 && non_mask_len <= rows / HWk, VFS_EINVAL, "generic_attention: bad non_mask_len");
  const dim3 grid((HWq + 31) / 32), block(32, 8);
  if (topk > 0)
    generic_attention_topk_kernel<16><<<grid, block, 0, s>>>(A, rows, ld, HWk, HWq, mask, non_mask_len, values, Cv,
                                                             topk, mode, out);
  else
    generic_attention_dense_kernel<<<grid, block, 0, s>>>(A, rows, ld, HWk, HWq, mask, non_mask_len, values, Cv, mode,
                                                          out);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

}  // namespace vfs
