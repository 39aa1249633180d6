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



// ------------------------------------------------------------------------------------------------------------------
// SiamFC response post-processing (TrackerSiamFC.update, projects/siamfc-pytorch/siamfc/siamfc_tracker_base.py:263-291):
//   cv2.resize(INTER_CUBIC) of each scale's response map R x R -> U x U, scale penalty on the non-centre scales, the
//   scale with the largest peak, then on that map: subtract the minimum, divide by the sum, blend with the Hann window,
//   arg-max.  The reference moves the maps to the host and runs cv2/numpy; here the maps never leave the device and
//   only {scale id, peak row, peak column} are read back.
// cv2 bicubic: a = -0.75, source coordinate (d + 0.5) * R / U - 0.5, taps floor-1 .. floor+2 with replicated borders,
// horizontal pass then vertical pass in fp32 (imgproc/resize.cpp, HResizeCubic / VResizeCubic).
// ------------------------------------------------------------------------------------------------------------------
namespace {

struct Cubic {
  int i0;       // index of the second tap (floor of the source coordinate)
  float w[4];   // weights of taps i0-1 .. i0+2
};
__device__ __forceinline__ Cubic cubic_setup(int d, float scale) {
  Cubic c;
  float f = (static_cast<float>(d) + 0.5f) * scale - 0.5f;
  const int s = static_cast<int>(floorf(f));
  f -= static_cast<float>(s);
  const float A = -0.75f;
  c.i0 = s;
  c.w[0] = ((A * (f + 1.0f) - 5.0f * A) * (f + 1.0f) + 8.0f * A) * (f + 1.0f) - 4.0f * A;
  c.w[1] = ((A + 2.0f) * f - (A + 3.0f)) * f * f + 1.0f;
  c.w[2] = ((A + 2.0f) * (1.0f - f) - (A + 3.0f)) * (1.0f - f) * (1.0f - f) + 1.0f;
  c.w[3] = 1.0f - c.w[0] - c.w[1] - c.w[2];
  return c;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// stats[s] = {ordered max, ordered min} (ints), sums[s] = sum of the upsampled, penalised map (fp64)
__global__ void siamfc_stats_init_kernel(int* stats, double* sums, int S) {
  const int s = threadIdx.x;
  if (s < S) {
    stats[2 * s] = INT_MIN;
    stats[2 * s + 1] = INT_MAX;
    sums[s] = 0.0;
  }
}

// grid (U, S), block = 32 * ceil(U / 32): one upsampled row of one scale
__global__ void siamfc_upsample_kernel(const float* __restrict__ resp, int R, int U, int centre, float penalty,
                                       float* __restrict__ up, int* __restrict__ stats, double* __restrict__ sums) {
  const int y = blockIdx.x, s = blockIdx.y, x = threadIdx.x;
  const float scale = static_cast<float>(R) / static_cast<float>(U);
  const float* src = resp + static_cast<size_t>(s) * R * R;
  float v = 0.0f;
  float lo = INFINITY, hi = -INFINITY;
  if (x < U) {
    const Cubic cy = cubic_setup(y, scale), cx = cubic_setup(x, scale);
    float h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float* row = src + clampi(cy.i0 - 1 + j, 0, R - 1) * R;
      h[j] = row[clampi(cx.i0 - 1, 0, R - 1)] * cx.w[0] + row[clampi(cx.i0, 0, R - 1)] * cx.w[1] +
             row[clampi(cx.i0 + 1, 0, R - 1)] * cx.w[2] + row[clampi(cx.i0 + 2, 0, R - 1)] * cx.w[3];
    }
    v = h[0] * cy.w[0] + h[1] * cy.w[1] + h[2] * cy.w[2] + h[3] * cy.w[3];
    if (s != centre) v *= penalty;
    up[(static_cast<size_t>(s) * U + y) * U + x] = v;
    lo = hi = v;
  }
  double sum = static_cast<double>(v);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(stats + 2 * s, float_to_ordered(hi));
    atomicMin(stats + 2 * s + 1, float_to_ordered(lo));
    atomicAdd(sums + s, sum);
  }
}

// one block: pick the scale, normalise its map, blend with the window, first arg-max -> out = {scale, row, col}
__global__ void siamfc_peak_kernel(const float* __restrict__ up, const int* __restrict__ stats,
                                   const double* __restrict__ sums, const double* __restrict__ hann, int S, int U,
                                   float window_influence, int* __restrict__ out) {
  __shared__ double best_v[32];
  __shared__ int best_i[32];
  int sid = 0;
  float peak = ordered_to_float(stats[0]);
  for (int s = 1; s < S; ++s) {
    const float m = ordered_to_float(stats[2 * s]);
    if (m > peak) {  // np.argmax: first maximum
      peak = m;
      sid = s;
    }
  }
  const float mn = ordered_to_float(stats[2 * sid + 1]);
  const int n = U * U;
  // numpy: response -= min (fp32); response /= response.sum() + 1e-16 (fp32 array / fp64 scalar -> fp32)
  const float denom = static_cast<float>((sums[sid] - static_cast<double>(n) * static_cast<double>(mn)) + 1e-16);
  const float keep = static_cast<float>(1.0 - static_cast<double>(window_influence));
  const double wi = static_cast<double>(window_influence);
  const float* map = up + static_cast<size_t>(sid) * n;
  double bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float r = __fdiv_rn(map[i] - mn, denom);
    const double v = static_cast<double>(keep * r) + wi * hann[i];  // fp32 product promoted by the fp64 window
    if (v > bv) {
      bv = v;
      bi = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) {
      bv = ov;
      bi = oi;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    best_v[threadIdx.x >> 5] = bv;
    best_i[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w)
      if (best_v[w] > bv || (best_v[w] == bv && best_i[w] < bi)) {
        bv = best_v[w];
        bi = best_i[w];
      }
    out[0] = sid;
    out[1] = bi / U;
    out[2] = bi % U;
  }
}

}  // namespace

size_t siamfc_peak_workspace_bytes(int S, int U) {
  return static_cast<size_t>(S) * U * U * sizeof(float) + static_cast<size_t>(S) * (2 * sizeof(int) + sizeof(double)) +
         64;
}

int siamfc_response_peak(const float* responses, int S, int R, int U, const double* hann, float scale_penalty,
                         float window_influence, void* workspace, int* out3, cudaStream_t s) {
  VFS_REQUIRE(responses && hann && workspace && out3, VFS_EINVAL, "siamfc_response_peak: null argument");
  VFS_REQUIRE(S >= 1 && S <= 32 && R >= 2 && U >= R && U <= 1024, VFS_ESHAPE,
              "siamfc_response_peak: S=%d R=%d U=%d unsupported", S, R, U);
  char* w = reinterpret_cast<char*>(workspace);
  double* sums = reinterpret_cast<double*>(w);                       // 8-byte aligned first
  int* stats = reinterpret_cast<int*>(w + ((S * sizeof(double) + 15) / 16) * 16);
  float* up = reinterpret_cast<float*>(w + ((S * sizeof(double) + 15) / 16) * 16 + ((2 * S * sizeof(int) + 15) / 16) * 16);
  siamfc_stats_init_kernel<<<1, 32, 0, s>>>(stats, sums, S);
  VFS_CUDA_OK(cudaGetLastError());
  siamfc_upsample_kernel<<<dim3(U, S), ((U + 31) / 32) * 32, 0, s>>>(responses, R, U, S / 2, scale_penalty, up, stats,
                                                                     sums);
  VFS_CUDA_OK(cudaGetLastError());
  siamfc_peak_kernel<<<1, 1024, 0, s>>>(up, stats, sums, hann, S, U, window_influence, out3);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

}  // namespace vfs
