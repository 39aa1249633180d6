// SimSiam head kernels: global average pool, Linear (skinny M), BatchNorm1d(+ReLU), cosine-similarity loss.
// Reference: mmaction/models/heads/sim_siam_head.py:143-174, losses/sim_loss.py:42-63.
//
// M (batch) is 8..128 rows, so each Linear is bound by streaming its fp32 weight matrix (up to 16.8 MB) once from
// HBM; tensor cores would not help.  Exact fp32 FMAs with a fixed summation order.
#include <math.h>

#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"

namespace vfs {

// ---- global average pool: NCHW fp32 [B,C,HW] -> [B,C]; one warp per (b,c)
__global__ void avgpool_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int BC, int HW) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= BC) return;
  const float* src = in + static_cast<size_t>(w) * HW;
  float s = 0.0f;
  for (int i = lane; i < HW; i += 32) s += src[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[w] = s / static_cast<float>(HW);
}

// ---- y[m, n] = sum_k x[m,k] W[n,k] + bias[n];  block = 8 warps = 8 output features, rows tiled by MT
template <int MT>
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                     const float* __restrict__ bias, float* __restrict__ y, int M,
                                                     int N, int K) {
  constexpr int KC = 8192 / MT;  // floats of K staged per iteration (32 KB of shared memory)
  __shared__ __align__(16) float xs[MT * KC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  const int m0 = blockIdx.y * MT;
  float acc[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m] = 0.0f;
  for (int k0 = 0; k0 < K; k0 += KC) {
    for (int i = threadIdx.x; i < MT * KC; i += 256) {
      const int m = i / KC, kk = i - m * KC;
      xs[i] = (m0 + m < M && k0 + kk < K) ? x[static_cast<size_t>(m0 + m) * K + k0 + kk] : 0.0f;
    }
    __syncthreads();
    if (n < N) {
      for (int kk = lane * 4; kk < KC; kk += 128) {
        if (k0 + kk < K) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(W + static_cast<size_t>(n) * K + k0 + kk));
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            const float4 x4 = *reinterpret_cast<const float4*>(xs + m * KC + kk);
            acc[m] = fmaf(w4.x, x4.x, acc[m]);
            acc[m] = fmaf(w4.y, x4.y, acc[m]);
            acc[m] = fmaf(w4.z, x4.z, acc[m]);
            acc[m] = fmaf(w4.w, x4.w, acc[m]);
          }
        }
      }
    }
    __syncthreads();
  }
  if (n >= N) return;
  const float b = bias ? bias[n] : 0.0f;
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    float v = acc[m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == (m & 31) && m0 + m < M) y[static_cast<size_t>(m0 + m) * N + n] = v + b;
  }
}

// ---- BatchNorm1d (+ReLU) in place over y[M,N]; one thread per feature.
// training: batch statistics (biased variance for the normalisation, unbiased into running_var, momentum update
// of the running stats exactly like torch.nn.functional.batch_norm); eval: running statistics.
__global__ void bn1d_act_kernel(float* __restrict__ y, int M, int N, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float* __restrict__ running_mean,
                                float* __restrict__ running_var, float eps, float momentum, int training, int relu,
                                float* __restrict__ save_mean, float* __restrict__ save_var) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float mean, var;
  if (training) {
    float s = 0.0f;
    for (int m = 0; m < M; ++m) s += y[static_cast<size_t>(m) * N + n];
    mean = s / static_cast<float>(M);
    float q = 0.0f;
    for (int m = 0; m < M; ++m) {
      const float d = y[static_cast<size_t>(m) * N + n] - mean;
      q = fmaf(d, d, q);
    }
    var = q / static_cast<float>(M);
    if (running_mean) {
      const float unbiased = (M > 1) ? q / static_cast<float>(M - 1) : var;
      running_mean[n] = (1.0f - momentum) * running_mean[n] + momentum * mean;
      running_var[n] = (1.0f - momentum) * running_var[n] + momentum * unbiased;
    }
  } else {
    mean = running_mean[n];
    var = running_var[n];
  }
  const float invstd = 1.0f / sqrtf(var + eps);
  if (save_mean) save_mean[n] = mean;
  if (save_var) save_var[n] = invstd;  // inverse standard deviation (what the backward pass needs)
  const float g = gamma ? gamma[n] : 1.0f, b = beta ? beta[n] : 0.0f;
  for (int m = 0; m < M; ++m) {
    float v = (y[static_cast<size_t>(m) * N + n] - mean) * invstd * g + b;
    if (relu) v = fmaxf(v, 0.0f);
    y[static_cast<size_t>(m) * N + n] = v;
  }
}

__global__ void relu_kernel(float* __restrict__ y, size_t n) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n) y[i] = fmaxf(y[i], 0.0f);
}

// ---- cosine similarity loss, one block per sample: 2 - 2*cos(p, z)  (or -cos)
__global__ void __launch_bounds__(128) cosine_loss_kernel(const float* __restrict__ p, const float* __restrict__ z,
                                                          float* __restrict__ loss, int D, int with_norm,
                                                          int negative) {
  __shared__ float red[3][4];
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
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = spp;
    red[1][warp] = szz;
    red[2][warp] = spz;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    spp = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    szz = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    spz = red[2][0] + red[2][1] + red[2][2] + red[2][3];
    float prod = spz;
    if (with_norm) prod = spz / (fmaxf(sqrtf(spp), 1e-12f) * fmaxf(sqrtf(szz), 1e-12f));  // F.normalize eps
    loss[b] = negative ? -prod : 2.0f - 2.0f * prod;
  }
}

// ------------------------------------------------------------------------------------------------
int global_avg_pool_nchw(const float* in, float* out, int B, int C, int HW, cudaStream_t s) {
  VFS_REQUIRE(in && out, VFS_EINVAL, "global_avg_pool: null argument");
  VFS_REQUIRE(B > 0 && C > 0 && HW > 0, VFS_ESHAPE, "global_avg_pool: empty tensor");
  const long long threads = static_cast<long long>(B) * C * 32;
  avgpool_nchw_kernel<<<static_cast<int>((threads + 255) / 256), 256, 0, s>>>(in, out, B * C, HW);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

// csrc/linear_mma.cu: warp-MMA (3xTF32) kernels for N % 16 == 0, K % 64 == 0; VFS_LINEAR_MMA=0 keeps the SIMT kernels
bool linear_mma_eligible(int M, int N, int K);
int linear_forward_mma(const float* x, const float* W, const float* bias, float* y, int M, int N, int K, cudaStream_t s);
bool linear_mma_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("VFS_LINEAR_MMA");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

int linear_forward(const float* x, const float* W, const float* bias, float* y, int M, int N, int K,
                   cudaStream_t s) {
  VFS_REQUIRE(x && W && y, VFS_EINVAL, "linear: null argument");
  VFS_REQUIRE(M > 0 && N > 0 && K > 0 && K % 4 == 0, VFS_ESHAPE, "linear: bad shape M=%d N=%d K=%d (K %% 4 == 0)", M,
              N, K);
  if (linear_mma_enabled() && linear_mma_eligible(M, N, K)) return linear_forward_mma(x, W, bias, y, M, N, K, s);
  const int nb = (N + 7) / 8;
  if (M <= 8) linear_kernel<8><<<dim3(nb, 1), 256, 0, s>>>(x, W, bias, y, M, N, K);
  else if (M <= 16) linear_kernel<16><<<dim3(nb, 1), 256, 0, s>>>(x, W, bias, y, M, N, K);
  else linear_kernel<32><<<dim3(nb, (M + 31) / 32), 256, 0, s>>>(x, W, bias, y, M, N, K);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int bn1d_act(float* y, int M, int N, const float* gamma, const float* beta, float* running_mean, float* running_var,
             float eps, float momentum, int training, int relu, float* save_mean, float* save_invstd, cudaStream_t s) {
  VFS_REQUIRE(y, VFS_EINVAL, "bn1d_act: null argument");
  VFS_REQUIRE(training || (running_mean && running_var), VFS_EINVAL, "bn1d_act: eval mode needs running stats");
  VFS_REQUIRE(!training || M > 1, VFS_ESHAPE,
              "bn1d_act: Expected more than 1 value per channel when training, got M=%d", M);
  bn1d_act_kernel<<<(N + 127) / 128, 128, 0, s>>>(y, M, N, gamma, beta, running_mean, running_var, eps, momentum,
                                                  training, relu, save_mean, save_invstd);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int relu_inplace(float* y, size_t n, cudaStream_t s) {
  VFS_REQUIRE(y, VFS_EINVAL, "relu: null argument");
  relu_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, s>>>(y, n);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int cosine_sim_loss(const float* p, const float* z, float* loss, int B, int D, int with_norm, int negative,
                    cudaStream_t s) {
  VFS_REQUIRE(p && z && loss, VFS_EINVAL, "cosine_sim_loss: null argument");
  VFS_REQUIRE(B > 0 && D > 0, VFS_ESHAPE, "cosine_sim_loss: empty input");
  cosine_loss_kernel<<<B, 128, 0, s>>>(p, z, loss, D, with_norm, negative);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

}  // namespace vfs
