// Tracker post-processing (reference VanillaTracker.forward_test, trackers/vanilla_tracker.py:162-181):
// bilinear upsample of the propagated label logits to image resolution (F.interpolate, align_corners=False),
// per-channel min-max normalisation (only where max > 0) and arg-max over channels -> uint8 label map.
// Two passes over the *output* grid, both recomputing the 4-tap bilinear sample instead of materialising the
// [Cv,H,W] upsampled tensor.
#include <limits.h>
#include <math.h>

#include "host_common.h"
#include "ptx.cuh"

namespace vfs {

__device__ __forceinline__ int float_to_ordered(float f) {
  const int i = __float_as_int(f);
  return (i >= 0) ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float((i >= 0) ? i : i ^ 0x7fffffff); }

struct Bilinear {
  int y0, y1, x0, x1;
  float ly, lx;
};
// torch area_pixel_compute_source_index(scale, dst, align_corners=false, cubic=false)
__device__ __forceinline__ Bilinear bilinear_setup(int oy, int ox, int h, int w, float sy, float sx) {
  Bilinear b;
  float fy = sy * (static_cast<float>(oy) + 0.5f) - 0.5f;
  float fx = sx * (static_cast<float>(ox) + 0.5f) - 0.5f;
  fy = fy < 0.0f ? 0.0f : fy;
  fx = fx < 0.0f ? 0.0f : fx;
  b.y0 = static_cast<int>(fy);
  b.x0 = static_cast<int>(fx);
  b.y1 = b.y0 + ((b.y0 < h - 1) ? 1 : 0);
  b.x1 = b.x0 + ((b.x0 < w - 1) ? 1 : 0);
  b.ly = fy - static_cast<float>(b.y0);
  b.lx = fx - static_cast<float>(b.x0);
  return b;
}
__device__ __forceinline__ float bilinear_sample(const float* __restrict__ p, int w, const Bilinear& b) {
  const float hy = 1.0f - b.ly, hx = 1.0f - b.lx;
  return hy * (hx * p[b.y0 * w + b.x0] + b.lx * p[b.y0 * w + b.x1]) +
         b.ly * (hx * p[b.y1 * w + b.x0] + b.lx * p[b.y1 * w + b.x1]);
}

// workspace layout: per problem [min[Cv], max[Cv]] as order-preserving ints
__global__ void minmax_init_kernel(int* mm, int Cv) {
  const int c = threadIdx.x;
  mm += static_cast<size_t>(blockIdx.x) * 2 * Cv;
  if (c < Cv) {
    mm[c] = INT_MAX;        // min
    mm[Cv + c] = INT_MIN;   // max
  }
}

__global__ void upsample_minmax_kernel(const float* __restrict__ logit, int Cv, int h, int w, int H, int W, float sy,
                                       float sx, int* __restrict__ mm) {
  const int c = blockIdx.y;
  const float* p = logit + (static_cast<size_t>(blockIdx.z) * Cv + c) * h * w;
  mm += static_cast<size_t>(blockIdx.z) * 2 * Cv;
  float lo = INFINITY, hi = -INFINITY;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    const Bilinear b = bilinear_setup(i / W, i % W, h, w, sy, sx);
    const float v = bilinear_sample(p, w, b);
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mm + c, float_to_ordered(lo));
    atomicMax(mm + Cv + c, float_to_ordered(hi));
  }
}

__global__ void normalize_argmax_kernel(const float* __restrict__ logit, int Cv, int h, int w, int H, int W, float sy,
                                        float sx, const int* __restrict__ mm, unsigned char* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  logit += static_cast<size_t>(blockIdx.y) * Cv * h * w;
  mm += static_cast<size_t>(blockIdx.y) * 2 * Cv;
  out += static_cast<size_t>(blockIdx.y) * H * W;
  const Bilinear b = bilinear_setup(i / W, i % W, h, w, sy, sx);
  float best = -INFINITY;
  int arg = 0;
  for (int c = 0; c < Cv; ++c) {
    float v = bilinear_sample(logit + static_cast<size_t>(c) * h * w, w, b);
    const float mn = ordered_to_float(mm[c]), mx = ordered_to_float(mm[Cv + c]);
    if (mx > 0.0f) v = (v - mn) / (mx - mn + 1e-12f);
    if (v > best) {  // first maximum wins, like torch.argmax on ties
      best = v;
      arg = c;
    }
  }
  out[i] = static_cast<unsigned char>(arg);
}

size_t seg_postprocess_workspace_bytes(int Cv) { return 2 * static_cast<size_t>(Cv) * sizeof(int); }

// P independent label maps (logit [P][Cv][h*w] -> out [P][H][W]) in three launches; workspace P * 2 * Cv ints.
int seg_postprocess(const float* logit, unsigned char* out, void* workspace, int P, int Cv, int h, int w, int H,
                    int W, cudaStream_t s) {
  VFS_REQUIRE(logit && out && workspace, VFS_EINVAL, "seg_postprocess: null argument");
  VFS_REQUIRE(P >= 1 && P <= 65535 && Cv >= 1 && Cv <= 256 && h > 0 && w > 0 && H > 0 && W > 0, VFS_ESHAPE,
              "seg_postprocess: bad shape");
  int* mm = reinterpret_cast<int*>(workspace);
  const float sy = static_cast<float>(h) / static_cast<float>(H), sx = static_cast<float>(w) / static_cast<float>(W);
  minmax_init_kernel<<<P, 256, 0, s>>>(mm, Cv);
  VFS_CUDA_OK(cudaGetLastError());
  int blocks = (H * W + 255) / 256;
  const int cap = (148 * 4 + P - 1) / P;
  if (blocks > cap) blocks = cap;
  upsample_minmax_kernel<<<dim3(blocks, Cv, P), 256, 0, s>>>(logit, Cv, h, w, H, W, sy, sx, mm);
  VFS_CUDA_OK(cudaGetLastError());
  normalize_argmax_kernel<<<dim3((H * W + 255) / 256, P), 256, 0, s>>>(logit, Cv, h, w, H, W, sy, sx, mm, out);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

}  // namespace vfs
