// SiamFC cross-correlation head: out[i, 0, oy, ox] = out_scale * sum_{dy,dx,c} x[i, oy+dy, ox+dx, c] * z[i % nz, dy, dx, c]
// Reference: SiamFC._fast_xcorr, projects/siamfc-pytorch/siamfc/heads.py:16-23 (grouped F.conv2d).
// 75.8 MFLOP on 2.6 MB of operands per pair -> bound by reading x/z (L2-resident), not by FLOPs; fp32 FMAs over
// NHWC operands so the channel run is the coalesced axis.
#include "host_common.h"
#include "ptx.cuh"

namespace vfs {

__global__ void nchw_to_nhwc_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* src = in + static_cast<size_t>(n) * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? src[static_cast<size_t>(c) * HW + p] : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < HW && c < C) out[(static_cast<size_t>(n) * HW + p) * C + c] = tile[threadIdx.x][i];
  }
}

// one block per output element; threads stride over (dy, dx, c/4)
__global__ void __launch_bounds__(256) xcorr_nhwc_kernel(const float* __restrict__ z, const float* __restrict__ x,
                                                         float* __restrict__ out, int nz, int C, int hz, int wz,
                                                         int h, int w, int ho, int wo, float out_scale) {
  __shared__ float red[8];
  const int ox = blockIdx.x, oy = blockIdx.y, i = blockIdx.z;
  const float* zz = z + static_cast<size_t>(i % nz) * hz * wz * C;
  const float* xx = x + static_cast<size_t>(i) * h * w * C;
  const int c4 = C / 4;
  const int total = hz * wz * c4;
  float acc = 0.0f;
  for (int t = threadIdx.x; t < total; t += 256) {
    const int cc = t % c4;
    const int r = t / c4;
    const int dx = r % wz, dy = r / wz;
    const float4 a = __ldg(reinterpret_cast<const float4*>(zz + (static_cast<size_t>(dy) * wz + dx) * C) + cc);
    const float4 b = __ldg(reinterpret_cast<const float4*>(xx + (static_cast<size_t>(oy + dy) * w + ox + dx) * C) + cc);
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    acc = fmaf(a.w, b.w, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k];
    out[(static_cast<size_t>(i) * ho + oy) * wo + ox] = s * out_scale;
  }
}

int nchw_to_nhwc_f32(const float* in, float* out, int N, int C, int H, int W, cudaStream_t s) {
  VFS_REQUIRE(in && out, VFS_EINVAL, "nchw_to_nhwc_f32: null argument");
  VFS_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, VFS_ESHAPE, "nchw_to_nhwc_f32: empty tensor");
  dim3 grid((H * W + 31) / 32, (C + 31) / 32, N), block(32, 8);
  nchw_to_nhwc_f32_kernel<<<grid, block, 0, s>>>(in, out, C, H * W);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int xcorr_nhwc(const float* z, const float* x, float* out, int nz, int nx, int C, int hz, int wz, int h, int w,
               float out_scale, cudaStream_t s) {
  VFS_REQUIRE(z && x && out, VFS_EINVAL, "xcorr: null argument");
  VFS_REQUIRE(nz > 0 && nx > 0 && nx % nz == 0, VFS_ESHAPE, "xcorr: nx=%d must be a positive multiple of nz=%d", nx, nz);
  VFS_REQUIRE(C > 0 && C % 4 == 0, VFS_ESHAPE, "xcorr: C=%d must be a multiple of 4", C);
  VFS_REQUIRE(hz > 0 && wz > 0 && h >= hz && w >= wz, VFS_ESHAPE, "xcorr: exemplar larger than search region");
  const int ho = h - hz + 1, wo = w - wz + 1;
  xcorr_nhwc_kernel<<<dim3(wo, ho, nx), 256, 0, s>>>(z, x, out, nz, C, hz, wz, h, w, ho, wo, out_scale);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

}  // namespace vfs
