// ResNet stem: conv 7x7/s2/p3 (3 -> 64) + BN(eval) + ReLU, then maxpool 3x3/s2/p1 -> split NHWC.
// Reference: ResNet._make_stem_layer / forward, mmaction/models/backbones/resnet.py:422-435, 565-566.
//
// K = 3*7*7 = 147 is too ragged for a 64-wide TMA/UMMA K-chunk and the layer is 3% of ResNet-50's
// FLOPs, so it runs as exact fp32 FMAs: one thread per conv-output pixel holding all 64 output channels
// in registers, input patch (parity-split columns -> conflict-free stride-2 reads) and the whole filter
// bank in shared memory (broadcast float4 reads).
#include "host_common.h"
#include "ptx.cuh"

namespace vfs {

constexpr int kStemTileH = 8, kStemTileW = 32;               // conv-output pixels per block (256 threads)
constexpr int kPatchH = kStemTileH * 2 + 5;                  // 21 input rows
constexpr int kPatchWHalf = 36;                              // (32*2+5 = 69 cols) split by parity -> 35, padded
constexpr int kStemK = 147;
constexpr int kStemSmemFloats = 3 * kPatchH * 2 * kPatchWHalf + kStemK * 64;

__global__ void __launch_bounds__(256) stem_conv_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                        const float* __restrict__ scale,
                                                        const float* __restrict__ shift, float* __restrict__ out,
                                                        int H, int W, int Hc, int Wc) {
  extern __shared__ float smem[];
  float* patch = smem;                                    // [3][21][2][36]
  float* wsm = smem + 3 * kPatchH * 2 * kPatchWHalf;      // [147][64]
  const int n = blockIdx.z;
  const int oy0 = blockIdx.y * kStemTileH, ox0 = blockIdx.x * kStemTileW;
  const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
  const float* src = in + static_cast<size_t>(n) * 3 * H * W;

  for (int i = threadIdx.x; i < kStemK * 64; i += 256) {
    const int k = i >> 6, co = i & 63;
    wsm[i] = w[co * kStemK + k];  // OIHW flatten: k = (c*7 + r)*7 + s
  }
  for (int i = threadIdx.x; i < 3 * kPatchH * 70; i += 256) {
    const int pc = i % 70;
    const int t = i / 70;
    const int pr = t % kPatchH;
    const int c = t / kPatchH;
    const int iy = iy0 + pr, ix = ix0 + pc;
    float v = 0.0f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = src[(static_cast<size_t>(c) * H + iy) * W + ix];
    patch[((c * kPatchH + pr) * 2 + (pc & 1)) * kPatchWHalf + (pc >> 1)] = v;
  }
  __syncthreads();

  const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
  float acc[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = 0.0f;

  for (int c = 0; c < 3; ++c) {
    for (int r = 0; r < 7; ++r) {
      const float* prow = patch + ((c * kPatchH + 2 * ty + r) * 2) * kPatchWHalf;
#pragma unroll
      for (int s = 0; s < 7; ++s) {
        const float x = prow[(s & 1) * kPatchWHalf + tx + (s >> 1)];
        const float4* wk = reinterpret_cast<const float4*>(wsm + ((c * 7 + r) * 7 + s) * 64);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 wv = wk[j];
          acc[4 * j + 0] = fmaf(x, wv.x, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(x, wv.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(x, wv.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(x, wv.w, acc[4 * j + 3]);
        }
      }
    }
  }

  const int oy = oy0 + ty, ox = ox0 + tx;
  if (oy < Hc && ox < Wc) {
    float4* dst = reinterpret_cast<float4*>(out + ((static_cast<size_t>(n) * Hc + oy) * Wc + ox) * 64);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float4 y;
      if (scale != nullptr) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + j);
        const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + j);
        y.x = fmaxf(fmaf(acc[4 * j + 0], sc.x, sh.x), 0.0f);
        y.y = fmaxf(fmaf(acc[4 * j + 1], sc.y, sh.y), 0.0f);
        y.z = fmaxf(fmaf(acc[4 * j + 2], sc.z, sh.z), 0.0f);
        y.w = fmaxf(fmaf(acc[4 * j + 3], sc.w, sh.w), 0.0f);
      } else {  // raw convolution output (train-mode BN: statistics are taken before normalisation)
        y = make_float4(acc[4 * j + 0], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
      }
      dst[j] = y;
    }
  }
}

// maxpool 3x3/s2/p1 over fp32 NHWC (C = 64) -> split NHWC.  One thread per (pixel, 8 channels).
__global__ void stem_pool_kernel(const float* __restrict__ in, const float* __restrict__ scale,
                                 const float* __restrict__ shift, h16* __restrict__ out_hi,
                                 h16* __restrict__ out_lo, int N, int Hc, int Wc, int Hp, int Wp) {
  const size_t total = static_cast<size_t>(N) * Hp * Wp * 8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i & 7);
    size_t t = i >> 3;
    const int px = static_cast<int>(t % Wp);
    t /= Wp;
    const int py = static_cast<int>(t % Hp);
    const int n = static_cast<int>(t / Hp);
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
    for (int dy = -1; dy <= 1; ++dy) {
      const int y = py * 2 + dy;
      if (y < 0 || y >= Hc) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        const int x = px * 2 + dx;
        if (x < 0 || x >= Wc) continue;
        const float4* src =
            reinterpret_cast<const float4*>(in + ((static_cast<size_t>(n) * Hc + y) * Wc + x) * 64 + g * 8);
        float4 a = src[0], b = src[1];
        if (scale != nullptr) {  // BN(batch statistics) + ReLU applied on the fly to the raw conv output
          const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + g * 8));
          const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale + g * 8 + 4));
          const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift + g * 8));
          const float4 h1 = __ldg(reinterpret_cast<const float4*>(shift + g * 8 + 4));
          a = make_float4(fmaxf(fmaf(a.x, s0.x, h0.x), 0.f), fmaxf(fmaf(a.y, s0.y, h0.y), 0.f),
                          fmaxf(fmaf(a.z, s0.z, h0.z), 0.f), fmaxf(fmaf(a.w, s0.w, h0.w), 0.f));
          b = make_float4(fmaxf(fmaf(b.x, s1.x, h1.x), 0.f), fmaxf(fmaf(b.y, s1.y, h1.y), 0.f),
                          fmaxf(fmaf(b.z, s1.z, h1.z), 0.f), fmaxf(fmaf(b.w, s1.w, h1.w), 0.f));
        }
        m[0] = fmaxf(m[0], a.x); m[1] = fmaxf(m[1], a.y); m[2] = fmaxf(m[2], a.z); m[3] = fmaxf(m[3], a.w);
        m[4] = fmaxf(m[4], b.x); m[5] = fmaxf(m[5], b.y); m[6] = fmaxf(m[6], b.z); m[7] = fmaxf(m[7], b.w);
      }
    }
    h16 hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split16(m[e], hi[e], lo[e]);
    uint4 oh, ol;
    oh.x = pack16x2(hi[0], hi[1]); oh.y = pack16x2(hi[2], hi[3]);
    oh.z = pack16x2(hi[4], hi[5]); oh.w = pack16x2(hi[6], hi[7]);
    ol.x = pack16x2(lo[0], lo[1]); ol.y = pack16x2(lo[2], lo[3]);
    ol.z = pack16x2(lo[4], lo[5]); ol.w = pack16x2(lo[6], lo[7]);
    const size_t o = ((static_cast<size_t>(n) * Hp + py) * Wp + px) * 64 + g * 8;
    *reinterpret_cast<uint4*>(out_hi + o) = oh;
    *reinterpret_cast<uint4*>(out_lo + o) = ol;
  }
}

static inline void stem_dims(int H, int W, int* Hc, int* Wc, int* Hp, int* Wp) {
  *Hc = (H + 6 - 7) / 2 + 1;
  *Wc = (W + 6 - 7) / 2 + 1;
  *Hp = (*Hc + 2 - 3) / 2 + 1;
  *Wp = (*Wc + 2 - 3) / 2 + 1;
}

size_t stem_workspace_bytes(int N, int H, int W) {
  int Hc, Wc, Hp, Wp;
  stem_dims(H, W, &Hc, &Wc, &Hp, &Wp);
  return static_cast<size_t>(N) * Hc * Wc * 64 * sizeof(float);
}

static int stem_conv_launch(const float* in, const float* weight, const float* scale, const float* shift,
                            float* conv_out, int N, int H, int W, cudaStream_t s) {
  int Hc, Wc, Hp, Wp;
  stem_dims(H, W, &Hc, &Wc, &Hp, &Wp);
  static bool configured = false;
  const int smem_bytes = kStemSmemFloats * static_cast<int>(sizeof(float));
  if (!configured) {
    VFS_CUDA_OK(cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  dim3 grid((Wc + kStemTileW - 1) / kStemTileW, (Hc + kStemTileH - 1) / kStemTileH, N);
  stem_conv_kernel<<<grid, 256, smem_bytes, s>>>(in, weight, scale, shift, conv_out, H, W, Hc, Wc);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

static int stem_pool_launch(const float* conv_out, const float* scale, const float* shift, void* out_split, int N,
                            int H, int W, cudaStream_t s) {
  int Hc, Wc, Hp, Wp;
  stem_dims(H, W, &Hc, &Wc, &Hp, &Wp);
  h16* hi = reinterpret_cast<h16*>(out_split);
  h16* lo = hi + static_cast<size_t>(N) * Hp * Wp * 64;
  const size_t total = static_cast<size_t>(N) * Hp * Wp * 8;
  const int blocks = static_cast<int>((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  stem_pool_kernel<<<blocks, 256, 0, s>>>(conv_out, scale, shift, hi, lo, N, Hc, Wc, Hp, Wp);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int stem_forward(const float* in, const float* weight, const float* scale, const float* shift, void* out_split,
                 void* workspace, int N, int H, int W, cudaStream_t s) {
  VFS_REQUIRE(in && weight && scale && shift && out_split && workspace, VFS_EINVAL, "stem_forward: null argument");
  VFS_REQUIRE(N > 0 && H >= 7 && W >= 7, VFS_ESHAPE, "stem_forward: input %dx%dx%d too small", N, H, W);
  float* conv_out = reinterpret_cast<float*>(workspace);
  int rc = stem_conv_launch(in, weight, scale, shift, conv_out, N, H, W, s);
  if (rc != VFS_OK) return rc;
  return stem_pool_launch(conv_out, nullptr, nullptr, out_split, N, H, W, s);
}

// train-mode pieces: raw conv output (statistics are computed on it), then BN+ReLU fused into the max-pool
int stem_conv_raw(const float* in, const float* weight, void* conv_out, int N, int H, int W, cudaStream_t s) {
  VFS_REQUIRE(in && weight && conv_out, VFS_EINVAL, "stem_conv_raw: null argument");
  VFS_REQUIRE(N > 0 && H >= 7 && W >= 7, VFS_ESHAPE, "stem_conv_raw: input %dx%dx%d too small", N, H, W);
  return stem_conv_launch(in, weight, nullptr, nullptr, reinterpret_cast<float*>(conv_out), N, H, W, s);
}

int stem_bn_relu_pool(const void* conv_out, const float* scale, const float* shift, void* out_split, int N, int H,
                      int W, cudaStream_t s) {
  VFS_REQUIRE(conv_out && scale && shift && out_split, VFS_EINVAL, "stem_bn_relu_pool: null argument");
  return stem_pool_launch(reinterpret_cast<const float*>(conv_out), scale, shift, out_split, N, H, W, s);
}

VFS_DEFINE_OVERFLOW_ACCESSOR(overflow_stem)

}  // namespace vfs
