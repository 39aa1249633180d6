// Layout boundary kernels (NCHW fp32 <-> split NHWC), weight packing, and the fp32 SIMT conv used as
// an on-device test instrument.
#include "host_common.h"
#include "ptx.cuh"

namespace vfs {

// ------------------------------------------------------------------------------------------------
// NCHW fp32 -> split NHWC (tile transpose through shared memory)
// ------------------------------------------------------------------------------------------------
__global__ void nchw_to_split_kernel(const float* __restrict__ in, h16* __restrict__ out_hi,
                                     h16* __restrict__ out_lo, int C, int HW, float scale) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* src = in + static_cast<size_t>(n) * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? src[static_cast<size_t>(c) * HW + p] * scale : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < HW && c < C) {
      h16 hi, lo;
      split16(tile[threadIdx.x][i], hi, lo);
      const size_t o = (static_cast<size_t>(n) * HW + p) * C + c;
      out_hi[o] = hi;
      out_lo[o] = lo;
    }
  }
}

__global__ void split_to_nchw_kernel(const h16* __restrict__ in_hi, const h16* __restrict__ in_lo,
                                     float* __restrict__ out, int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    float v = 0.0f;
    if (p < HW && c < C) {
      const size_t o = (static_cast<size_t>(n) * HW + p) * C + c;
      v = h16_to_float(in_hi[o]) + h16_to_float(in_lo[o]);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  float* dst = out + static_cast<size_t>(n) * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) dst[static_cast<size_t>(c) * HW + p] = tile[threadIdx.x][i];
  }
}

int nchw_f32_to_split(const float* in, void* out_split, int N, int C, int H, int W, float scale, cudaStream_t s) {
  VFS_REQUIRE(in && out_split, VFS_EINVAL, "nchw_f32_to_split: null argument");
  VFS_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, VFS_ESHAPE, "nchw_f32_to_split: empty tensor");
  const int HW = H * W;
  h16* hi = reinterpret_cast<h16*>(out_split);
  h16* lo = hi + static_cast<size_t>(N) * HW * C;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
  nchw_to_split_kernel<<<grid, block, 0, s>>>(in, hi, lo, C, HW, scale);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int split_to_nchw_f32(const void* in_split, float* out, int N, int C, int H, int W, cudaStream_t s) {
  VFS_REQUIRE(in_split && out, VFS_EINVAL, "split_to_nchw_f32: null argument");
  VFS_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, VFS_ESHAPE, "split_to_nchw_f32: empty tensor");
  const int HW = H * W;
  const h16* hi = reinterpret_cast<const h16*>(in_split);
  const h16* lo = hi + static_cast<size_t>(N) * HW * C;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
  split_to_nchw_kernel<<<grid, block, 0, s>>>(hi, lo, out, C, HW);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

// ------------------------------------------------------------------------------------------------
// OIHW fp32 -> split [2][Cout][(r*k+s)*Cin + ci]
// ------------------------------------------------------------------------------------------------
// ``wscale`` (a power of two) multiplies the weights before the split: kaiming-initialised weights are ~0.02, where
// the lo plane (2^-11 of the value) is an fp16 subnormal and the pair keeps ~17 instead of 22 significant bits; scaled
// by 256 both planes are normal numbers for |w| >= 5e-4.  The caller folds 1 / wscale into the epilogue's scale vector.
__global__ void pack_weight_kernel(const float* __restrict__ w, h16* __restrict__ hi,
                                   h16* __restrict__ lo, int Cout, int Cin, int k, float wscale) {
  const size_t total = static_cast<size_t>(Cout) * Cin * k * k;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // destination index i = (co * k*k + tap) * Cin + ci
    const int ci = static_cast<int>(i % Cin);
    const size_t t = i / Cin;
    const int tap = static_cast<int>(t % (k * k));
    const int co = static_cast<int>(t / (k * k));
    const float v = w[(static_cast<size_t>(co) * Cin + ci) * k * k + tap] * wscale;
    h16 h, l;
    split16(v, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// OIHW fp32 -> split [2][Cin][(r'*k+s')*Cout + co] with the kernel flipped (operand of the data-gradient conv)
__global__ void pack_weight_dgrad_kernel(const float* __restrict__ w, h16* __restrict__ hi,
                                         h16* __restrict__ lo, int Cout, int Cin, int k, float wscale) {
  const size_t total = static_cast<size_t>(Cout) * Cin * k * k;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // destination index i = (ci * k*k + tap') * Cout + co
    const int co = static_cast<int>(i % Cout);
    const size_t t = i / Cout;
    const int tap = static_cast<int>(t % (k * k));
    const int ci = static_cast<int>(t / (k * k));
    const int src_tap = k * k - 1 - tap;  // (k-1-r', k-1-s')
    const float v = w[(static_cast<size_t>(co) * Cin + ci) * k * k + src_tap] * wscale;
    h16 h, l;
    split16(v, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// Multi-tensor form: every conv of the network (forward and / or data-gradient layout) in ONE launch.  A training step
// re-packs all weights after each optimiser update; per-tensor launches were 208 of its ~1200 kernels
// (profiles/r02_train_profile_cfg4_v1.log).  Block b serves the item whose [first_block, next first_block) contains it.
constexpr int kPackPerBlock = 2048;
__global__ void __launch_bounds__(256) pack_weights_multi_kernel(const VfsPackItem* __restrict__ items, int n) {
  int lo_i = 0, hi_i = n - 1;
  while (lo_i < hi_i) {   // last item with first_block <= blockIdx.x
    const int mid = (lo_i + hi_i + 1) >> 1;
    if (items[mid].first_block <= static_cast<int>(blockIdx.x)) lo_i = mid;
    else hi_i = mid - 1;
  }
  const VfsPackItem it = items[lo_i];
  const float wscale = exp2f(static_cast<float>(it.scale_log2));
  const int Cout = it.Cout, Cin = it.Cin, kk = it.ksize * it.ksize;
  const size_t total = static_cast<size_t>(Cout) * Cin * kk;
  h16* hi = reinterpret_cast<h16*>(it.dst_split);
  h16* lo = hi + total;
  const size_t begin = static_cast<size_t>(blockIdx.x - it.first_block) * kPackPerBlock;
  size_t end = begin + kPackPerBlock;
  if (end > total) end = total;
  for (size_t i = begin + threadIdx.x; i < end; i += 256) {
    float v;
    if (it.mode == 0) {   // forward operand: i = (co * k*k + tap) * Cin + ci
      const int ci = static_cast<int>(i % Cin);
      const size_t t = i / Cin;
      const int tap = static_cast<int>(t % kk);
      const int co = static_cast<int>(t / kk);
      v = it.w[(static_cast<size_t>(co) * Cin + ci) * kk + tap] * wscale;
    } else {              // data-gradient operand: i = (ci * k*k + tap') * Cout + co, kernel flipped
      const int co = static_cast<int>(i % Cout);
      const size_t t = i / Cout;
      const int tap = static_cast<int>(t % kk);
      const int ci = static_cast<int>(t / kk);
      v = it.w[(static_cast<size_t>(co) * Cin + ci) * kk + (kk - 1 - tap)] * wscale;
    }
    h16 h, l;
    split16(v, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

int pack_conv_weights_multi(const VfsPackItem* items_dev, int n, int total_blocks, cudaStream_t s) {
  VFS_REQUIRE(items_dev && n > 0 && total_blocks > 0, VFS_EINVAL, "pack_conv_weights_multi: bad argument");
  pack_weights_multi_kernel<<<total_blocks, 256, 0, s>>>(items_dev, n);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int pack_conv_weight_dgrad(const float* w, void* wt_split, int Cout, int Cin, int k, float wscale, cudaStream_t s) {
  VFS_REQUIRE(w && wt_split, VFS_EINVAL, "pack_conv_weight_dgrad: null argument");
  VFS_REQUIRE(Cout > 0 && Cin > 0 && k > 0, VFS_ESHAPE, "pack_conv_weight_dgrad: bad shape");
  const size_t total = static_cast<size_t>(Cout) * Cin * k * k;
  h16* hi = reinterpret_cast<h16*>(wt_split);
  const int blocks = static_cast<int>((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  pack_weight_dgrad_kernel<<<blocks, 256, 0, s>>>(w, hi, hi + total, Cout, Cin, k, wscale);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int pack_conv_weight(const float* w, void* w_split, int Cout, int Cin, int k, float wscale, cudaStream_t s) {
  VFS_REQUIRE(w && w_split, VFS_EINVAL, "pack_conv_weight: null argument");
  VFS_REQUIRE(Cout > 0 && Cin > 0 && k > 0, VFS_ESHAPE, "pack_conv_weight: bad shape");
  const size_t total = static_cast<size_t>(Cout) * Cin * k * k;
  h16* hi = reinterpret_cast<h16*>(w_split);
  const int blocks = static_cast<int>((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  pack_weight_kernel<<<blocks, 256, 0, s>>>(w, hi, hi + total, Cout, Cin, k, wscale);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

// ------------------------------------------------------------------------------------------------
// fp32 SIMT conv (test instrument): one thread per output element, fixed summation order.
// ------------------------------------------------------------------------------------------------
__global__ void conv_simt_kernel(VfsConvDesc d, const h16* __restrict__ in_hi,
                                 const h16* __restrict__ in_lo, const h16* __restrict__ w_hi,
                                 const h16* __restrict__ w_lo, const float* __restrict__ scale,
                                 const float* __restrict__ shift, const h16* __restrict__ res_hi,
                                 const h16* __restrict__ res_lo, float* __restrict__ out, int Ho, int Wo,
                                 int pad, int dil) {
  const size_t total = static_cast<size_t>(d.N) * Ho * Wo * d.Cout;
  const int k = d.ksize;
  const size_t Ktot = static_cast<size_t>(k) * k * d.Cin;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i % d.Cout);
    size_t t = i / d.Cout;
    const int ox = static_cast<int>(t % Wo);
    t /= Wo;
    const int oy = static_cast<int>(t % Ho);
    const int n = static_cast<int>(t / Ho);
    float acc = 0.0f;
    for (int r = 0; r < k; ++r) {
      const int iy = oy * d.stride + r * dil - pad;
      if (iy < 0 || iy >= d.H) continue;
      for (int c = 0; c < k; ++c) {
        const int ix = ox * d.stride + c * dil - pad;
        if (ix < 0 || ix >= d.W) continue;
        const size_t ibase = ((static_cast<size_t>(n) * d.H + iy) * d.W + ix) * d.Cin;
        const size_t wbase = co * Ktot + static_cast<size_t>(r * k + c) * d.Cin;
        for (int ci = 0; ci < d.Cin; ++ci) {
          const float x = h16_to_float(in_hi[ibase + ci]) + h16_to_float(in_lo[ibase + ci]);
          const float ww = h16_to_float(w_hi[wbase + ci]) + h16_to_float(w_lo[wbase + ci]);
          acc = fmaf(x, ww, acc);
        }
      }
    }
    float y = fmaf(acc, scale[co], shift[co]);
    if (res_hi) y += h16_to_float(res_hi[i]) + h16_to_float(res_lo[i]);
    if (d.relu) y = fmaxf(y, 0.0f);
    out[i] = y;
  }
}

int conv_bn_act_simt(const VfsConvDesc* d, const void* in_split, const void* w_split, const float* scale,
                     const float* shift, const void* residual_split, float* out_f32, cudaStream_t stream) {
  VFS_REQUIRE(d && in_split && w_split && scale && shift && out_f32, VFS_EINVAL, "conv_simt: null argument");
  const int k = d->ksize, s = d->stride, dil = (k == 1) ? 1 : d->dilation;
  const int pad = (k == 1) ? 0 : dil;
  const int Ho = (d->H + 2 * pad - dil * (k - 1) - 1) / s + 1;
  const int Wo = (d->W + 2 * pad - dil * (k - 1) - 1) / s + 1;
  const size_t in_plane = static_cast<size_t>(d->N) * d->H * d->W * d->Cin;
  const size_t out_plane = static_cast<size_t>(d->N) * Ho * Wo * d->Cout;
  const size_t w_plane = static_cast<size_t>(d->Cout) * k * k * d->Cin;
  const h16* in_hi = reinterpret_cast<const h16*>(in_split);
  const h16* w_hi = reinterpret_cast<const h16*>(w_split);
  const h16* r_hi = reinterpret_cast<const h16*>(residual_split);
  const int blocks = static_cast<int>((out_plane + 255) / 256 < 65535 ? (out_plane + 255) / 256 : 65535);
  conv_simt_kernel<<<blocks, 256, 0, stream>>>(*d, in_hi, in_hi + in_plane, w_hi, w_hi + w_plane, scale, shift, r_hi,
                                               r_hi ? r_hi + out_plane : nullptr, out_f32, Ho, Wo, pad, dil);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Data feed: the reference's CPU pipeline steps ``Normalize`` (mmaction/datasets/pipelines/
// augmentations.py:711-757 -> mmcv.imnormalize_: optional channel swap, (x - mean) * (1/std) in fp32) and
// ``FormatShape('NCTHW')`` (formating.py:248-258: [M,H,W,C] -> [M/T, C, T, H, W]) on the device, so that frames cross
// PCIe as uint8 HWC (a quarter of the fp32 NCTHW bytes).
//   in  uint8 [clips][T][H][W][3]      out fp32 [clips][3][T][H][W]
// One thread = 4 consecutive pixels: three 4-byte loads (12 bytes), three float4 stores (one per channel plane).
// ------------------------------------------------------------------------------------------------------------------
__global__ void frames_u8_to_ncthw_kernel(const unsigned char* __restrict__ in, float* __restrict__ out, long long clips,
                                          int T, long long HW, float m0, float m1, float m2, double s0, double s1, double s2,
                                          int swap_rb) {
  const long long quads = HW / 4;  // HW % 4 == 0 is required by the caller
  const long long total = clips * T * quads;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long qd = i % quads;
    const long long ft = i / quads;  // clip * T + t
    const long long t = ft % T, clip = ft / T;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(in + (ft * HW + qd * 4) * 3);
    const uint32_t w0 = src[0], w1 = src[1], w2 = src[2];
    // bytes: p0 = (b0 b1 b2) p1 = (b3 b4 b5) p2 = (b6 b7 b8) p3 = (b9 b10 b11)
    float c0[4], c1[4], c2[4];
    c0[0] = static_cast<float>(w0 & 0xff);         c1[0] = static_cast<float>((w0 >> 8) & 0xff);
    c2[0] = static_cast<float>((w0 >> 16) & 0xff); c0[1] = static_cast<float>(w0 >> 24);
    c1[1] = static_cast<float>(w1 & 0xff);         c2[1] = static_cast<float>((w1 >> 8) & 0xff);
    c0[2] = static_cast<float>((w1 >> 16) & 0xff); c1[2] = static_cast<float>(w1 >> 24);
    c2[2] = static_cast<float>(w2 & 0xff);         c0[3] = static_cast<float>((w2 >> 8) & 0xff);
    c1[3] = static_cast<float>((w2 >> 16) & 0xff); c2[3] = static_cast<float>(w2 >> 24);
    const float* first = swap_rb ? c2 : c0;   // channel written to output plane 0
    const float* third = swap_rb ? c0 : c2;
    const long long plane = T * HW;
    float* dst = out + (clip * 3 * T + t) * HW + qd * 4;
    // cv2.subtract on a CV_32F image rounds to fp32; cv2.multiply by the float64 1/std scalar multiplies in fp64 and
    // rounds once to fp32 (checked against cv2 bit for bit by the parity test)
    auto norm = [](float x, float m, double sinv) {
      return __double2float_rn(static_cast<double>(__fsub_rn(x, m)) * sinv);
    };
    *reinterpret_cast<float4*>(dst) =
        make_float4(norm(first[0], m0, s0), norm(first[1], m0, s0), norm(first[2], m0, s0), norm(first[3], m0, s0));
    *reinterpret_cast<float4*>(dst + plane) =
        make_float4(norm(c1[0], m1, s1), norm(c1[1], m1, s1), norm(c1[2], m1, s1), norm(c1[3], m1, s1));
    *reinterpret_cast<float4*>(dst + 2 * plane) =
        make_float4(norm(third[0], m2, s2), norm(third[1], m2, s2), norm(third[2], m2, s2), norm(third[3], m2, s2));
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Training data feed on the device: RandomResizedCrop (the crop box itself is drawn on the host) -> Resize(keep_ratio
// False, cv2 bilinear) -> Flip(horizontal) -> Normalize -> FormatShape('NCTHW') of the reference's train_pipeline
// (configs/*:48-92; augmentations.py:171-330, 487-597, 600-711, 711-757) as ONE kernel per batch.
// The resize is cv2.resize(INTER_LINEAR) on uint8 bit for bit (resize.cpp: fx = float((dx+0.5)*scale-0.5); coefficients
// saturate_cast<short>(x * 2048); horizontal pass in int32; vertical pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2 >> 2;
// left/right border columns collapse to one source pixel, rows clamp).  One thread = one output pixel, 3 channels.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cv_linear_coef(int d, double scale, int src_n, bool collapse, int& ofs, int& c0, int& c1) {
  float f = static_cast<float>((d + 0.5) * scale - 0.5);
  int s0 = static_cast<int>(floorf(f));
  f -= static_cast<float>(s0);
  if (collapse) {   // horizontal axis: outside [0, n-1) the sample is one source pixel
    if (s0 < 0) { f = 0.0f; s0 = 0; }
    if (s0 >= src_n - 1) { f = 0.0f; s0 = src_n - 1; }
  }
  ofs = s0;
  c0 = __float2int_rn((1.0f - f) * 2048.0f);
  c1 = __float2int_rn(f * 2048.0f);
}

__global__ void augment_u8_to_ncthw_kernel(const VfsAugItem* __restrict__ items, float* __restrict__ out, int T, int dst_h,
                                           int dst_w, float m0, float m1, float m2, double s0, double s1, double s2,
                                           int swap_rb) {
  const int f = blockIdx.y;                       // frame = clip * T + t
  const VfsAugItem it = items[f];
  const int clip = f / T, t = f - clip * T;
  const long long HW = static_cast<long long>(dst_h) * dst_w;
  const double scale_x = static_cast<double>(it.crop_w) / dst_w, scale_y = static_cast<double>(it.crop_h) / dst_h;
  const unsigned char* src = it.src + (static_cast<long long>(it.crop_y0) * it.W + it.crop_x0) * 3;
  const long long row_pitch = static_cast<long long>(it.W) * 3;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < HW;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int oy = static_cast<int>(i / dst_w), ox = static_cast<int>(i - static_cast<long long>(oy) * dst_w);
    const int dx = it.flip ? dst_w - 1 - ox : ox;     // Flip runs after Resize: mirrored column of the resized image
    int xo, a0, a1, yo, b0, b1;
    cv_linear_coef(dx, scale_x, it.crop_w, true, xo, a0, a1);
    cv_linear_coef(oy, scale_y, it.crop_h, false, yo, b0, b1);
    const int x1 = min(xo + 1, it.crop_w - 1);
    const int y0 = min(max(yo, 0), it.crop_h - 1), y1 = min(max(yo + 1, 0), it.crop_h - 1);
    const unsigned char* r0 = src + y0 * row_pitch;
    const unsigned char* r1 = src + y1 * row_pitch;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int h0 = r0[xo * 3 + c] * a0 + r0[x1 * 3 + c] * a1;
      const int h1 = r1[xo * 3 + c] * a0 + r1[x1 * 3 + c] * a1;
      int px = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
      px = min(max(px, 0), 255);
      v[c] = static_cast<float>(px);
    }
    auto norm = [](float x, float m, double sinv) {
      return __double2float_rn(static_cast<double>(__fsub_rn(x, m)) * sinv);
    };
    float* dst = out + (static_cast<long long>(clip) * 3 * T + t) * HW + i;
    const long long plane = static_cast<long long>(T) * HW;
    dst[0] = norm(swap_rb ? v[2] : v[0], m0, s0);
    dst[plane] = norm(v[1], m1, s1);
    dst[2 * plane] = norm(swap_rb ? v[0] : v[2], m2, s2);
  }
}

int augment_u8_to_ncthw_f32(const VfsAugItem* items_dev, float* out, long long clips, int T, int dst_h, int dst_w,
                            const float* mean3, const double* stdinv3, int swap_rb, cudaStream_t s) {
  VFS_REQUIRE(items_dev && out && mean3 && stdinv3, VFS_EINVAL, "augment_u8_to_ncthw_f32: null argument");
  VFS_REQUIRE(clips > 0 && T > 0 && dst_h > 0 && dst_w > 0 && clips * T <= 65535, VFS_ESHAPE,
              "augment_u8_to_ncthw_f32: bad shape (at most 65535 frames per call)");
  const long long HW = static_cast<long long>(dst_h) * dst_w;
  int bx = static_cast<int>((HW + 255) / 256);
  if (bx > 64) bx = 64;
  augment_u8_to_ncthw_kernel<<<dim3(bx, static_cast<unsigned>(clips * T)), 256, 0, s>>>(
      items_dev, out, T, dst_h, dst_w, mean3[0], mean3[1], mean3[2], stdinv3[0], stdinv3[1], stdinv3[2], swap_rb);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int frames_u8_to_ncthw_f32(const unsigned char* in, float* out, long long clips, int T, int H, int W, const float* mean3,
                           const double* stdinv3, int swap_rb, cudaStream_t s) {
  VFS_REQUIRE(in && out && mean3 && stdinv3, VFS_EINVAL, "frames_u8_to_ncthw_f32: null argument");
  VFS_REQUIRE(clips > 0 && T > 0 && H > 0 && W > 0, VFS_ESHAPE, "frames_u8_to_ncthw_f32: empty input");
  const long long HW = static_cast<long long>(H) * W;
  VFS_REQUIRE(HW % 4 == 0, VFS_ESHAPE, "frames_u8_to_ncthw_f32: H*W = %lld must be a multiple of 4", HW);
  const long long total = clips * T * (HW / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  frames_u8_to_ncthw_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(in, out, clips, T, HW, mean3[0], mean3[1], mean3[2],
                                                                    stdinv3[0], stdinv3[1], stdinv3[2], swap_rb);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

VFS_DEFINE_OVERFLOW_ACCESSOR(overflow_layout)

}  // namespace vfs
