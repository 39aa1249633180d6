// SiamFC linear-probe training (TrackerSiamFC.train_step, projects/siamfc-pytorch/siamfc/siamfc_tracker_base.py:364-386):
// the backbone is frozen (default_config_base.py:40-49: frozen_stages=4, norm_eval), so the step is
//   responses = xcorr(z_convs(f_z), x_convs(f_x)) * out_scale          (siamfc/heads.py:46-58)
//   loss      = FocalLoss / BalancedLoss(responses, labels)             (siamfc/losses.py:27-64)
//   backward  : d responses -> d(adapter outputs) through the correlation -> weight / bias gradients of the two 1x1 convs
//   optimizer : Adam (default) or SGD
// The 1x1 convolutions and their weight gradients run on the tcgen05 kernels (conv_tc.cu / wgrad_tc.cu); this file holds
// the loss (value + gradient in one launch), the two correlation gradients and the Adam update.
#include <math.h>

#include "host_common.h"

namespace vfs {

namespace {

__device__ __forceinline__ float log_sigmoid_ref(float x) {      // losses.py:8-14
  return fminf(x, 0.0f) - logf(1.0f + expf(-fabsf(x)));
}
__device__ __forceinline__ float log_minus_sigmoid_ref(float x) {  // losses.py:17-23
  return fminf(-x, 0.0f) - logf(1.0f + expf(-fabsf(x)));
}

__device__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.0f;
  for (int k = 0; k < static_cast<int>(blockDim.x >> 5); ++k) s += red[k];
  return s;
}

// mode 0: FocalLoss(gamma): loss_i = -(t w+ log sig(x) + (1-t) w- log(1-sig(x))), w+ = (1-p)^g, w- = p^g,
//         loss = mean(loss_i / mean(t w+ + (1-t) w-))            (the normaliser is differentiated too, like autograd)
// mode 1: BalancedLoss: weights 1/#pos, neg_weight/#neg normalised to sum 1, sum of weighted BCE-with-logits
// One block (n is a few thousand response values).  grad = d loss / d x.
__global__ void __launch_bounds__(1024) siamfc_loss_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                           float* __restrict__ loss, float* __restrict__ grad, int n,
                                                           int mode, float gamma, float neg_weight) {
  __shared__ float red[32];
  if (mode == 0) {
    float sl = 0.0f, sw = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float xi = x[i], ti = t[i];
      const float p = 1.0f / (1.0f + expf(-xi));
      const float wp = powf(1.0f - p, gamma), wn = powf(p, gamma);
      sl += -(ti * wp * log_sigmoid_ref(xi) + (1.0f - ti) * wn * log_minus_sigmoid_ref(xi));
      sw += ti * wp + (1.0f - ti) * wn;
    }
    const float S = block_sum(sl, red);
    const float Wt = block_sum(sw, red);
    const float A = Wt / n;                       // avg_weight.mean()
    if (threadIdx.x == 0) loss[0] = S / (A * n);
    if (grad == nullptr) return;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float xi = x[i], ti = t[i];
      const float p = 1.0f / (1.0f + expf(-xi));
      const float q = 1.0f - p;
      const float wp = powf(q, gamma), wn = powf(p, gamma);
      const float ls = log_sigmoid_ref(xi), lm = log_minus_sigmoid_ref(xi);
      // d w+/dx = -g (1-p)^(g-1) p (1-p) = -g w+ p ; d w-/dx = g p^(g-1) p (1-p) = g w- (1-p)
      const float dwp = -gamma * wp * p, dwn = gamma * wn * q;
      // d log sig / dx = 1-p ; d log(1-sig) / dx = -p
      const float dl = -(ti * (dwp * ls + wp * q) + (1.0f - ti) * (dwn * lm - wn * p));
      const float dw = ti * dwp + (1.0f - ti) * dwn;
      // L = S / (A n), A = W / n  =>  L = S / W ; dL/dx_i = dl_i / W - S dw_i / W^2
      grad[i] = dl / Wt - S * dw / (Wt * Wt);
    }
  } else {
    float np_ = 0.0f, nn_ = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      np_ += (t[i] == 1.0f) ? 1.0f : 0.0f;
      nn_ += (t[i] == 0.0f) ? 1.0f : 0.0f;
    }
    const float P = block_sum(np_, red), Nn = block_sum(nn_, red);
    // weight[pos] = 1/P, weight[neg] = neg_weight/Nn, then / sum(weight) = 1 + neg_weight (when both classes exist)
    const float wsum = (P > 0.0f ? 1.0f : 0.0f) + (Nn > 0.0f ? neg_weight : 0.0f);
    float sl = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float xi = x[i], ti = t[i];
      float w = 0.0f;
      if (ti == 1.0f) w = 1.0f / P;
      else if (ti == 0.0f) w = neg_weight / Nn;
      w /= wsum;
      // binary_cross_entropy_with_logits: max(x,0) - x t + log(1 + exp(-|x|))
      sl += w * (fmaxf(xi, 0.0f) - xi * ti + logf(1.0f + expf(-fabsf(xi))));
      if (grad != nullptr) grad[i] = w * (1.0f / (1.0f + expf(-xi)) - ti);
    }
    const float S = block_sum(sl, red);
    if (threadIdx.x == 0) loss[0] = S;
  }
}

// dZ[i, dy, dx, c] = scale * sum_{oy,ox} dR[i, oy, ox] * X[i, oy+dy, ox+dx, c]      (pairs: exemplar i <-> search i)
// block = one (dx, dy, i); thread = 4 channels
__global__ void __launch_bounds__(128) xcorr_bwd_z_kernel(const float* __restrict__ dr, const float* __restrict__ x,
                                                          float* __restrict__ dz, int C, int hz, int wz, int h, int w,
                                                          int ho, int wo, float scale) {
  extern __shared__ float s_dr[];
  const int dx = blockIdx.x, dy = blockIdx.y, i = blockIdx.z;
  for (int k = threadIdx.x; k < ho * wo; k += blockDim.x) s_dr[k] = dr[static_cast<size_t>(i) * ho * wo + k];
  __syncthreads();
  const float* xx = x + static_cast<size_t>(i) * h * w * C;
  for (int c4 = threadIdx.x; c4 < C / 4; c4 += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int oy = 0; oy < ho; ++oy) {
      const float4* row = reinterpret_cast<const float4*>(xx + (static_cast<size_t>(oy + dy) * w + dx) * C) + c4;
      for (int ox = 0; ox < wo; ++ox) {
        const float g = s_dr[oy * wo + ox];
        const float4 v = __ldg(row + static_cast<size_t>(ox) * (C / 4));
        acc.x = fmaf(g, v.x, acc.x); acc.y = fmaf(g, v.y, acc.y);
        acc.z = fmaf(g, v.z, acc.z); acc.w = fmaf(g, v.w, acc.w);
      }
    }
    acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
    reinterpret_cast<float4*>(dz + ((static_cast<size_t>(i) * hz + dy) * wz + dx) * C)[c4] = acc;
  }
}

// dX[i, py, px, c] = scale * sum_{dy,dx : 0 <= py-dy < ho, 0 <= px-dx < wo} dR[i, py-dy, px-dx] * Z[i, dy, dx, c]
__global__ void __launch_bounds__(128) xcorr_bwd_x_kernel(const float* __restrict__ dr, const float* __restrict__ z,
                                                          float* __restrict__ dxo, int C, int hz, int wz, int h, int w,
                                                          int ho, int wo, float scale) {
  extern __shared__ float s_dr[];
  const int px = blockIdx.x, py = blockIdx.y, i = blockIdx.z;
  for (int k = threadIdx.x; k < ho * wo; k += blockDim.x) s_dr[k] = dr[static_cast<size_t>(i) * ho * wo + k];
  __syncthreads();
  const float* zz = z + static_cast<size_t>(i) * hz * wz * C;
  const int dy0 = max(0, py - (ho - 1)), dy1 = min(hz - 1, py);
  const int dx0 = max(0, px - (wo - 1)), dx1 = min(wz - 1, px);
  for (int c4 = threadIdx.x; c4 < C / 4; c4 += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int dy = dy0; dy <= dy1; ++dy) {
      for (int dx = dx0; dx <= dx1; ++dx) {
        const float g = s_dr[(py - dy) * wo + (px - dx)];
        const float4 v = __ldg(reinterpret_cast<const float4*>(zz + (static_cast<size_t>(dy) * wz + dx) * C) + c4);
        acc.x = fmaf(g, v.x, acc.x); acc.y = fmaf(g, v.y, acc.y);
        acc.z = fmaf(g, v.z, acc.z); acc.w = fmaf(g, v.w, acc.w);
      }
    }
    acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
    reinterpret_cast<float4*>(dxo + ((static_cast<size_t>(i) * h + py) * w + px) * C)[c4] = acc;
  }
}

// torch.optim.Adam (no amsgrad): g' = g + wd p ; m = b1 m + (1-b1) g' ; v = b2 v + (1-b2) g'^2 ;
// p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps, float wd, float bc1,
                            float bc2_sqrt) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float gg = fmaf(wd, p[i], g[i]);
    const float mi = b1 * m[i] + (1.0f - b1) * gg;
    const float vi = b2 * v[i] + (1.0f - b2) * gg * gg;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - (lr / bc1) * (mi / denom);
  }
}

}  // namespace

int siamfc_loss(const float* responses, const float* labels, float* loss, float* grad, int n, int mode, float gamma,
                float neg_weight, cudaStream_t s) {
  VFS_REQUIRE(responses && labels && loss, VFS_EINVAL, "siamfc_loss: null argument");
  VFS_REQUIRE(n > 0 && (mode == 0 || mode == 1), VFS_EINVAL, "siamfc_loss: bad argument");
  siamfc_loss_kernel<<<1, 1024, 0, s>>>(responses, labels, loss, grad, n, mode, gamma, neg_weight);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int xcorr_backward_nhwc(const float* dr, const float* z, const float* x, float* dz, float* dx, int n, int C, int hz, int wz,
                        int h, int w, float out_scale, cudaStream_t s) {
  VFS_REQUIRE(dr && z && x && (dz || dx), VFS_EINVAL, "xcorr_backward: null argument");
  VFS_REQUIRE(n > 0 && C > 0 && C % 4 == 0 && hz > 0 && wz > 0 && h >= hz && w >= wz, VFS_ESHAPE,
              "xcorr_backward: bad shape");
  const int ho = h - hz + 1, wo = w - wz + 1;
  const size_t smem = static_cast<size_t>(ho) * wo * sizeof(float);
  VFS_REQUIRE(smem <= 48 * 1024, VFS_ESHAPE, "xcorr_backward: response map %dx%d too large", ho, wo);
  if (dz) {
    xcorr_bwd_z_kernel<<<dim3(wz, hz, n), 128, smem, s>>>(dr, x, dz, C, hz, wz, h, w, ho, wo, out_scale);
    VFS_CUDA_OK(cudaGetLastError());
  }
  if (dx) {
    xcorr_bwd_x_kernel<<<dim3(w, h, n), 128, smem, s>>>(dr, z, dx, C, hz, wz, h, w, ho, wo, out_scale);
    VFS_CUDA_OK(cudaGetLastError());
  }
  return VFS_OK;
}

int adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
              float weight_decay, int step, cudaStream_t s) {
  VFS_REQUIRE(p && g && m && v && step >= 1, VFS_EINVAL, "adam_step: bad argument");
  if (n == 0) return VFS_OK;
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay,
                                                      static_cast<float>(bc1), static_cast<float>(sqrt(bc2)));
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

}  // namespace vfs
