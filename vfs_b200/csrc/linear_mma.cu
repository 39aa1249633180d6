// SimSiam head Linear layers (projector / predictor, reference sim_siam_head.py:56-117: 2048 x 2048 weights, M = clips
// per GPU = 8..32 rows) on warp-level tensor-core MMAs.
//
// These layers are weight-streaming problems: 16.8 MB of fp32 weights against 0.27 GFLOP.  The SIMT kernels they
// replace were bound by shared-memory operand reads and by their own staging barriers (36-41 us per layer; 10 layers
// forward + 10 backward per cfg-4 step).  Here every warp streams its slab of W straight from global memory into MMA
// fragments with 16-byte loads and multiplies with `mma.sync.m16n8k8` in 3xTF32 (x = hi + lo in tf32, the three
// significant products, fp32 accumulate: fp32-grade results, error ~2^-21 relative per product) -- tcgen05 would need
// the weights repacked into split-fp16 tiles every step, which costs more than the whole layer.
//
// The fragment layouts of m16n8k8 (g = lane / 4, t = lane % 4):
//   A[16 x 8]  a0 (g, t)   a1 (g + 8, t)   a2 (g, t + 4)   a3 (g + 8, t + 4)
//   B[ 8 x 8]  b0 (k = t, n = g)           b1 (k = t + 4, n = g)
//   C[16 x 8]  c0 (g, 2t)  c1 (g, 2t + 1)  c2 (g + 8, 2t)  c3 (g + 8, 2t + 1)
// Both the reduction index and the output column index may be permuted freely as long as A and B (resp. B and C) use
// the same permutation; the kernels choose permutations that turn every fragment load into one float4.
#include "host_common.h"

namespace vfs {
namespace {

__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
  const float r = v - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct Frag4 {  // hi / lo tf32 planes of one float4
  uint32_t hi[4], lo[4];
};
__device__ __forceinline__ Frag4 split4(const float4& v) {
  Frag4 f;
  split_tf32(v.x, f.hi[0], f.lo[0]);
  split_tf32(v.y, f.hi[1], f.lo[1]);
  split_tf32(v.z, f.hi[2], f.lo[2]);
  split_tf32(v.w, f.hi[3], f.lo[3]);
  return f;
}
// d += A B with A = (a, b rows g / g + 8; reduction slots i0, i1), B = (w slots j0, j1): lo x hi, hi x lo, hi x hi
__device__ __forceinline__ void mma3(float (&d)[4], const Frag4& a, const Frag4& b, int i0, int i1, const Frag4& w,
                                     const Frag4& w2, int j0, int j1) {
  mma_tf32(d, a.lo[i0], b.lo[i0], a.lo[i1], b.lo[i1], w.hi[j0], w2.hi[j1]);
  mma_tf32(d, a.hi[i0], b.hi[i0], a.hi[i1], b.hi[i1], w.lo[j0], w2.lo[j1]);
  mma_tf32(d, a.hi[i0], b.hi[i0], a.hi[i1], b.hi[i1], w.hi[j0], w2.hi[j1]);
}

__device__ __forceinline__ float4 ldg4(const float* p, bool ok) {
  return ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---------------------------------------------------------------------------------------------------------------
// y[m, n] = sum_k x[m, k] W[n, k] + bias[n].   Block = 8 warps = 2 column tiles of 8 features x 4 quarters of K;
// grid = (N / 16, M / (16 MT)).  Inside a 16-wide block of k, lane (g, t) loads W[n0 + g][kb + 4t .. 4t + 3] and the x
// rows g, g + 8 at the same k: MMA step s in {0, 1} takes physical k = kb + 4t + 2s for slot t and kb + 4t + 2s + 1
// for slot t + 4.  The quarters are summed through shared memory in a fixed order (deterministic).
// ---------------------------------------------------------------------------------------------------------------
template <int MT>
__global__ void __launch_bounds__(256) linear_fwd_mma_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                             const float* __restrict__ bias, float* __restrict__ y,
                                                             int M, int N, int K) {
  __shared__ float red[3][MT * 16][17];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int ntile = warp & 1, kq = warp >> 1;
  const int n0 = blockIdx.x * 16 + ntile * 8;
  const int m_base = blockIdx.y * (16 * MT);
  const int kq_len = K / 4;  // K % 64 == 0
  const int k_begin = kq * kq_len, k_end = k_begin + kq_len;
  const bool n_ok = n0 + g < N;
  const float* wp = W + static_cast<size_t>(n_ok ? n0 + g : 0) * K + 4 * t;
  const float* xa[MT];
  const float* xb[MT];
  bool a_ok[MT], b_ok[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int ra = m_base + mt * 16 + g, rb = ra + 8;
    a_ok[mt] = ra < M;
    b_ok[mt] = rb < M;
    xa[mt] = x + static_cast<size_t>(a_ok[mt] ? ra : 0) * K + 4 * t;
    xb[mt] = x + static_cast<size_t>(b_ok[mt] ? rb : 0) * K + 4 * t;
  }
  // The tensor-core adder truncates: a long accumulation chain drifts (measured 3e-6 relative at K = 2048, several
  // times the error of an fp32 FMA chain).  Every 64 values of k the MMA accumulator is folded into an fp32 sum with a
  // round-to-nearest add and restarted.
  float acc[MT][4], part[MT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[mt][i] = 0.0f;

  for (int k64 = k_begin; k64 < k_end; k64 += 64) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) part[mt][i] = 0.0f;
    const int k64_end = min(k64 + 64, k_end);
#pragma unroll 4
    for (int kb = k64; kb < k64_end; kb += 16) {
      const Frag4 w = split4(ldg4(wp + kb, n_ok));
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const Frag4 a = split4(ldg4(xa[mt] + kb, a_ok[mt]));
        const Frag4 b = split4(ldg4(xb[mt] + kb, b_ok[mt]));
        mma3(part[mt], a, b, 0, 1, w, w, 0, 1);
        mma3(part[mt], a, b, 2, 3, w, w, 2, 3);
      }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][i] += part[mt][i];
  }

  if (kq > 0) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      red[kq - 1][mt * 16 + g][ntile * 8 + 2 * t] = acc[mt][0];
      red[kq - 1][mt * 16 + g][ntile * 8 + 2 * t + 1] = acc[mt][1];
      red[kq - 1][mt * 16 + g + 8][ntile * 8 + 2 * t] = acc[mt][2];
      red[kq - 1][mt * 16 + g + 8][ntile * 8 + 2 * t + 1] = acc[mt][3];
    }
  }
  __syncthreads();
  if (kq == 0) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = mt * 16 + g + (i >> 1) * 8, col = ntile * 8 + 2 * t + (i & 1);
        const int m = m_base + row, n = blockIdx.x * 16 + col;
        if (m < M && n < N) {
          float v = acc[mt][i];
          v += red[0][row][col];
          v += red[1][row][col];
          v += red[2][row][col];
          y[static_cast<size_t>(m) * N + n] = v + (bias ? bias[n] : 0.0f);
        }
      }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// dX[m, k] (+)= sum_n dY[m, n] W[n, k].   Block = 32 consecutive k (four column tiles: lane (g, t) loads
// W[row][kc + 4g .. 4g + 3], so tile j holds physical column kc + 4c + j at logical column c) x one slice of n
// (blockIdx.z), the slice's 16-row blocks dealt round-robin to the 8 warps.  Inside a 16-row block lane (g, t) loads
// dY[g | g + 8][nb + 4t .. 4t + 3] and W rows nb + 4t .. 4t + 3: step s takes row nb + 4t + 2s for slot t and
// nb + 4t + 2s + 1 for slot t + 4.  Warps are summed through shared memory in a fixed order; slices of n are combined
// with fp32 atomics into the zeroed dX (like the kernel this replaces).
// ---------------------------------------------------------------------------------------------------------------
template <int MT>
__global__ void __launch_bounds__(256) linear_bwd_data_mma_kernel(const float* __restrict__ dy,
                                                                  const float* __restrict__ W, float* __restrict__ dx,
                                                                  int M, int N, int K, int n_slice) {
  __shared__ float red[8][MT * 16][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int kc = blockIdx.x * 32;
  const int m_base = blockIdx.y * (16 * MT);
  const int n_begin = blockIdx.z * n_slice, n_end = min(N, n_begin + n_slice);
  const float* da[MT];
  const float* db[MT];
  bool a_ok[MT], b_ok[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int ra = m_base + mt * 16 + g, rb = ra + 8;
    a_ok[mt] = ra < M;
    b_ok[mt] = rb < M;
    da[mt] = dy + static_cast<size_t>(a_ok[mt] ? ra : 0) * N + 4 * t;
    db[mt] = dy + static_cast<size_t>(b_ok[mt] ? rb : 0) * N + 4 * t;
  }
  float acc[MT][4][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][j][i] = 0.0f;

  const float* wcol = W + kc + 4 * g;
  // (per 16 rows of n a fresh MMA accumulator, folded into the fp32 sum with a round-to-nearest add: see the forward
  // kernel)
#pragma unroll 2
  for (int nb = n_begin + 16 * warp; nb < n_end; nb += 16 * 8) {  // N % 16 == 0
    Frag4 w[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) w[r] = split4(ldg4(wcol + static_cast<size_t>(nb + 4 * t + r) * K, true));
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const Frag4 a = split4(ldg4(da[mt] + nb, a_ok[mt]));
      const Frag4 b = split4(ldg4(db[mt] + nb, b_ok[mt]));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float part[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        mma3(part, a, b, 0, 1, w[0], w[1], j, j);
        mma3(part, a, b, 2, 3, w[2], w[3], j, j);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][j][i] += part[i];
      }
    }
  }

  // tile j, c0 / c1: logical columns 2t, 2t + 1 -> physical kc + 8t + j, kc + 8t + 4 + j
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[warp][mt * 16 + g][8 * t + j] = acc[mt][j][0];
      red[warp][mt * 16 + g][8 * t + 4 + j] = acc[mt][j][1];
      red[warp][mt * 16 + g + 8][8 * t + j] = acc[mt][j][2];
      red[warp][mt * 16 + g + 8][8 * t + 4 + j] = acc[mt][j][3];
    }
  __syncthreads();
  for (int i = threadIdx.x; i < MT * 16 * 32; i += 256) {
    const int row = i >> 5, col = i & 31;
    const int m = m_base + row;
    if (m >= M) continue;
    float v = 0.0f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) v += red[w8][row][col];
    float* d = dx + static_cast<size_t>(m) * K + kc + col;
    if (gridDim.z > 1) atomicAdd(d, v);
    else *d = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// dW[n, k] (+)= sum_m dY[m, n] X[m, k],  db[n] (+)= sum_m dY[m, n]   (M <= 32: the reduction is 4 MMA steps).
// Block = 32 consecutive k (four column tiles, permuted like in the data kernel) x a range of 16-row tiles of n dealt
// round-robin to the 8 warps; the X fragments of the block's columns (all M rows) stay in registers and are reused for
// every row tile, so the kernel streams dW once (read-modify-write when accumulating) and nothing else.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) linear_bwd_weight_mma_kernel(const float* __restrict__ dy,
                                                                    const float* __restrict__ x, float* __restrict__ dW,
                                                                    float* __restrict__ db, int M, int N, int K,
                                                                    int tiles_per_block, int accumulate) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int kc = blockIdx.x * 32;
  Frag4 xf[4][2];  // step (8 rows of m each) x {row t, row t + 4}
#pragma unroll
  for (int st = 0; st < 4; ++st) {
    const int m0 = st * 8 + t, m1 = m0 + 4;
    xf[st][0] = split4(ldg4(x + static_cast<size_t>(m0 < M ? m0 : 0) * K + kc + 4 * g, m0 < M));
    xf[st][1] = split4(ldg4(x + static_cast<size_t>(m1 < M ? m1 : 0) * K + kc + 4 * g, m1 < M));
  }
  const int tile_begin = blockIdx.y * tiles_per_block;
  const int tile_end = min(N / 16, tile_begin + tiles_per_block);
  for (int nt = tile_begin + warp; nt < tile_end; nt += 8) {
    const int n0 = nt * 16;
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;
    float bs0 = 0.0f, bs1 = 0.0f;
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      const int m0 = st * 8 + t, m1 = m0 + 4;
      const float v0 = m0 < M ? __ldg(dy + static_cast<size_t>(m0) * N + n0 + g) : 0.0f;       // A(g, t)
      const float v1 = m0 < M ? __ldg(dy + static_cast<size_t>(m0) * N + n0 + g + 8) : 0.0f;   // A(g + 8, t)
      const float v2 = m1 < M ? __ldg(dy + static_cast<size_t>(m1) * N + n0 + g) : 0.0f;       // A(g, t + 4)
      const float v3 = m1 < M ? __ldg(dy + static_cast<size_t>(m1) * N + n0 + g + 8) : 0.0f;   // A(g + 8, t + 4)
      bs0 += v0 + v2;
      bs1 += v1 + v3;
      uint32_t h0, l0, h1, l1, h2, l2, h3, l3;
      split_tf32(v0, h0, l0);
      split_tf32(v1, h1, l1);
      split_tf32(v2, h2, l2);
      split_tf32(v3, h3, l3);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        mma_tf32(acc[j], l0, l1, l2, l3, xf[st][0].hi[j], xf[st][1].hi[j]);
        mma_tf32(acc[j], h0, h1, h2, h3, xf[st][0].lo[j], xf[st][1].lo[j]);
        mma_tf32(acc[j], h0, h1, h2, h3, xf[st][0].hi[j], xf[st][1].hi[j]);
      }
    }
    // row n0 + g: columns kc + 8t + {0..3} (c0 of tiles 0..3) and kc + 8t + 4 + {0..3} (c1); row n0 + g + 8: c2 / c3
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float* row = dW + static_cast<size_t>(n0 + g + 8 * half) * K + kc + 8 * t;
      float4 lo4 = make_float4(acc[0][2 * half], acc[1][2 * half], acc[2][2 * half], acc[3][2 * half]);
      float4 hi4 = make_float4(acc[0][2 * half + 1], acc[1][2 * half + 1], acc[2][2 * half + 1], acc[3][2 * half + 1]);
      if (accumulate) {
        const float4 o0 = *reinterpret_cast<const float4*>(row);
        const float4 o1 = *reinterpret_cast<const float4*>(row + 4);
        lo4.x += o0.x; lo4.y += o0.y; lo4.z += o0.z; lo4.w += o0.w;
        hi4.x += o1.x; hi4.y += o1.y; hi4.z += o1.z; hi4.w += o1.w;
      }
      *reinterpret_cast<float4*>(row) = lo4;
      *reinterpret_cast<float4*>(row + 4) = hi4;
    }
    if (db && blockIdx.x == 0) {
      bs0 += __shfl_xor_sync(0xffffffffu, bs0, 1);
      bs0 += __shfl_xor_sync(0xffffffffu, bs0, 2);
      bs1 += __shfl_xor_sync(0xffffffffu, bs1, 1);
      bs1 += __shfl_xor_sync(0xffffffffu, bs1, 2);
      if (t == 0) {
        db[n0 + g] = accumulate ? db[n0 + g] + bs0 : bs0;
        db[n0 + g + 8] = accumulate ? db[n0 + g + 8] + bs1 : bs1;
      }
    }
  }
}

}  // namespace

bool linear_mma_eligible(int M, int N, int K) { return M > 0 && N % 16 == 0 && K % 64 == 0; }

int linear_forward_mma(const float* x, const float* W, const float* bias, float* y, int M, int N, int K,
                       cudaStream_t s) {
  if (M <= 16) linear_fwd_mma_kernel<1><<<dim3(N / 16, 1), 256, 0, s>>>(x, W, bias, y, M, N, K);
  else linear_fwd_mma_kernel<2><<<dim3(N / 16, (M + 31) / 32), 256, 0, s>>>(x, W, bias, y, M, N, K);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

// dx must be zeroed by the caller when the returned slice count is > 1 (this function zeroes it itself)
int linear_bwd_data_mma(const float* dy, const float* W, float* dx, int M, int N, int K, cudaStream_t s) {
  // slices of n so that the grid has >= 2 blocks per SM; every slice a multiple of 128 rows (8 warps x 16)
  const int col_blocks = K / 32;
  const int m_blocks = M <= 16 ? 1 : (M + 31) / 32;
  int slices = (2 * 148 + col_blocks * m_blocks - 1) / (col_blocks * m_blocks);
  int n_slice = ((N + slices - 1) / slices + 127) / 128 * 128;
  slices = (N + n_slice - 1) / n_slice;
  if (slices > 1) VFS_CUDA_OK(cudaMemsetAsync(dx, 0, static_cast<size_t>(M) * K * sizeof(float), s));
  const dim3 grid(col_blocks, m_blocks, slices);
  if (M <= 16) linear_bwd_data_mma_kernel<1><<<grid, 256, 0, s>>>(dy, W, dx, M, N, K, n_slice);
  else linear_bwd_data_mma_kernel<2><<<grid, 256, 0, s>>>(dy, W, dx, M, N, K, n_slice);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int linear_bwd_weight_mma(const float* dy, const float* x, float* dW, float* db, int M, int N, int K, int accumulate,
                          cudaStream_t s) {
  const int tiles = N / 16;
  int splits = (2 * 148 + K / 32 - 1) / (K / 32);                 // >= 2 blocks per SM
  int per = ((tiles + splits - 1) / splits + 7) / 8 * 8;          // whole rounds of the 8 warps
  splits = (tiles + per - 1) / per;
  linear_bwd_weight_mma_kernel<<<dim3(K / 32, splits), 256, 0, s>>>(dy, x, dW, db, M, N, K, per, accumulate);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

}  // namespace vfs

