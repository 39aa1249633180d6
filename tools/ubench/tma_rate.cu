// Micro-benchmark: how fast can one SM's TMA unit fill shared memory with the boxes conv_tc_kernel uses?
// One CTA per SM; thread 0 issues the loads of a stage into a ring of S slots, thread 32 consumes (waits for the
// bytes, frees the slot at once).  Prints SM cycles per stage for several box decompositions of the same 64 KB stage
// (activation tile 128 pixels x 64 ch x 2 planes + weight tile 128 rows x 64 x 2 planes).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I vfs_b200/csrc \
//        tools/ubench/tma_rate.cu -o tools/ubench/tma_rate
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace vfs;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled encode_fn() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  return reinterpret_cast<PFN_encodeTiled>(fn);
}

static CUtensorMap make_map(void* base, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box,
                            CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B) {
  CUtensorMap m;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides[i - 1];
  }
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, base, gdim, gstr, bdim, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d\n", (int)r);
    exit(1);
  }
  return m;
}

struct Params {
  CUtensorMap a5;   // {C, W, H, N, 2} box {64, 16, 8, 1, 2}
  CUtensorMap a4;   // {C, W, H, N} one plane, box {64, 16, 8, 1}
  CUtensorMap a2;   // {C, pixels} one plane, box {64, 128}
  CUtensorMap a3f;  // {C, pixels, 2} box {64, 128, 2}
  CUtensorMap b3;   // {K, Cout, 2} box {64, 128, 2}
  CUtensorMap b2;   // {K, Cout} one plane box {64, 128}
  CUtensorMap a5h;  // box {32, 16, 8, 1, 2}, 64B swizzle (half-K stage)
  CUtensorMap b3h;  // box {32, 128, 2}, 64B swizzle
  int stage_bytes;
  long long* cycles;
  int iters, mode, stages;
  long long plane_a_elems, plane_b_elems;
};

// mode 0: a5 + b3 (what conv_tc_kernel issues for a 3x3 layer)      2 instructions, 64 KB
// mode 1: a4 x2 + b2 x2                                             4 instructions, 64 KB
// mode 2: a5 only (32 KB)         mode 3: b3 only (32 KB)
// mode 4: a3f + b3 (1x1 conv, flat pixel axis)                      mode 5: a2 x2 + b2 x2
__global__ void __launch_bounds__(96, 1) tma_rate_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + p.stages * p.stage_bytes;
  auto full = [&](int s) { return bar + 8u * s; };
  auto empty = [&](int s) { return bar + 8u * (8 + s); };
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    fence_mbar_init();
  }
  __syncthreads();
  const int tile = blockIdx.x;  // 148 CTAs over 16 images x (2 x 4) tiles of 16x8
  const int n0 = tile % 16, h0 = ((tile / 16) % 4) * 8, w0 = ((tile / 64) % 2) * 16;
  const int pix0 = (tile % 128) * 128;
  const int bytes = (p.mode == 2 || p.mode == 3 || p.mode == 6 || p.mode == 7) ? 32768 : 65536;
  if (threadIdx.x == 0) {
    int stage = 0;
    uint32_t phase = 0;
    const long long t0 = clock64();
    long long t_first = 0;
    for (int it = 0; it < p.iters; ++it) {
      mbar_wait(empty(stage), phase ^ 1u, 1);
      const uint32_t sa = base + stage * p.stage_bytes, sb = sa + p.stage_bytes / 2;
      const int tap = it % 9, dh = tap / 3 - 1, dw = tap % 3 - 1, c0 = ((it / 9) % 4) * 64;
      const int kb = (it % 36) * 64, nb = (it / 36 % 2) * 128;
      mbar_arrive_expect_tx(full(stage), bytes);
      switch (p.mode) {
        case 0:
          tma_load_5d(sa, &p.a5, full(stage), c0, w0 + dw, h0 + dh, n0, 0);
          tma_load_3d(sb, &p.b3, full(stage), kb, nb, 0);
          break;
        case 1:
          tma_load_4d(sa, &p.a4, full(stage), c0, w0 + dw, h0 + dh, n0);
          tma_load_4d(sa + 16384, &p.a4, full(stage), c0, w0 + dw, h0 + dh, n0);
          tma_load_2d(sb, &p.b2, full(stage), kb, nb);
          tma_load_2d(sb + 16384, &p.b2, full(stage), kb, nb);
          break;
        case 2:
          tma_load_5d(sa, &p.a5, full(stage), c0, w0 + dw, h0 + dh, n0, 0);
          break;
        case 3:
          tma_load_3d(sb, &p.b3, full(stage), kb, nb, 0);
          break;
        case 4:
          tma_load_3d(sa, &p.a3f, full(stage), c0, pix0, 0);
          tma_load_3d(sb, &p.b3, full(stage), kb, nb, 0);
          break;
        case 6:  // half-K stage: 32 channels of both planes, A + B = 32 KB, issued by this one thread
          tma_load_5d(sa, &p.a5h, full(stage), c0 / 2, w0 + dw, h0 + dh, n0, 0);
          tma_load_3d(sb, &p.b3h, full(stage), kb / 2, nb, 0);
          break;
        case 7:  // half-K stage, A only from this thread (B comes from thread 64)
          tma_load_5d(sa, &p.a5h, full(stage), c0 / 2, w0 + dw, h0 + dh, n0, 0);
          break;
        default:
          tma_load_2d(sa, &p.a2, full(stage), c0, pix0);
          tma_load_2d(sa + 16384, &p.a2, full(stage), c0, pix0);
          tma_load_2d(sb, &p.b2, full(stage), kb, nb);
          tma_load_2d(sb + 16384, &p.b2, full(stage), kb, nb);
          break;
      }
      if (it == p.stages - 1) t_first = clock64() - t0;
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
    p.cycles[blockIdx.x * 2 + 1] = t_first;
  } else if (threadIdx.x == 64 && p.mode == 7) {
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < p.iters; ++it) {
      mbar_wait(empty(stage), phase ^ 1u, 3);
      const uint32_t sb = base + stage * p.stage_bytes + p.stage_bytes / 2;
      const int kb = (it % 36) * 64, nb = (it / 36 % 2) * 128;
      tma_load_3d(sb, &p.b3h, full(stage), kb / 2, nb, 0);
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (threadIdx.x == 32) {
    int stage = 0;
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
      mbar_wait(full(stage), phase, 2);
      mbar_arrive(empty(stage));
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
    p.cycles[blockIdx.x * 2] = clock64() - t0;
  }
}

int main() {
  const int N = 16, H = 32, W = 32, C = 256, Cout = 256, K = 9 * 256;
  const size_t a_plane = static_cast<size_t>(N) * H * W * C, b_plane = static_cast<size_t>(Cout) * K;
  __half *a, *b;
  cudaMalloc(&a, 2 * a_plane * 2);
  cudaMalloc(&b, 2 * b_plane * 2);
  cudaMemset(a, 0, 2 * a_plane * 2);
  cudaMemset(b, 0, 2 * b_plane * 2);
  Params p;
  {
    uint64_t d[5] = {C, W, H, N, 2}, s[4] = {C * 2ull, W * C * 2ull, H * W * C * 2ull, a_plane * 2};
    uint32_t bx[5] = {64, 16, 8, 1, 2};
    p.a5 = make_map(a, 5, d, s, bx);
    p.a4 = make_map(a, 4, d, s, bx);
  }
  {
    uint64_t d[3] = {C, (uint64_t)N * H * W, 2}, s[2] = {C * 2ull, a_plane * 2};
    uint32_t bx[3] = {64, 128, 2};
    p.a3f = make_map(a, 3, d, s, bx);
    p.a2 = make_map(a, 2, d, s, bx);
  }
  {
    uint64_t d[3] = {K, Cout, 2}, s[2] = {K * 2ull, b_plane * 2};
    uint32_t bx[3] = {64, 128, 2};
    p.b3 = make_map(b, 3, d, s, bx);
    p.b2 = make_map(b, 2, d, s, bx);
  }
  {
    uint64_t d[5] = {C, W, H, N, 2}, s5[4] = {C * 2ull, W * C * 2ull, H * W * C * 2ull, a_plane * 2};
    uint32_t bx[5] = {32, 16, 8, 1, 2};
    p.a5h = make_map(a, 5, d, s5, bx, CU_TENSOR_MAP_SWIZZLE_64B);
    uint64_t d3[3] = {K, Cout, 2}, s3[2] = {K * 2ull, b_plane * 2};
    uint32_t bx3[3] = {32, 128, 2};
    p.b3h = make_map(b, 3, d3, s3, bx3, CU_TENSOR_MAP_SWIZZLE_64B);
  }
  cudaMalloc(&p.cycles, 148 * 2 * sizeof(long long));
  p.iters = 720;
  const char* names[] = {"5D act(2 planes) + 3D wgt(2 planes)  [conv_tc 3x3]", "4D act x2 + 2D wgt x2 (per plane)",
                         "5D act only (32 KB)", "3D wgt only (32 KB)", "3D flat act + 3D wgt [conv_tc 1x1]",
                         "2D flat act x2 + 2D wgt x2", "half-K: 5D act + 3D wgt, 64B swizzle (32 KB stage)",
                         "half-K, act and wgt issued by two threads"};
  for (int pass = 0; pass < 3; ++pass)
    for (int mode = (pass == 0 ? 0 : 6); mode < (pass == 0 ? 6 : 8); ++mode) {
      const int stages = (pass == 0) ? 3 : (pass == 1 ? 6 : 3);
      p.mode = mode;
      p.stages = stages;
      p.stage_bytes = (mode >= 6) ? 32768 : 65536;
      const int smem = stages * p.stage_bytes + 1024 + 256;
      cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      for (int rep = 0; rep < 2; ++rep) tma_rate_kernel<<<148, 96, smem>>>(p);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[296];
      cudaMemcpy(h, p.cycles, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0, fi = 0;
      for (int i = 0; i < 148; ++i) {
        mx = h[2 * i] > mx ? h[2 * i] : mx;
        fi = h[2 * i + 1] > fi ? h[2 * i + 1] : fi;
      }
      const double per = double(mx) / p.iters;
      const int bytes = (mode == 2 || mode == 3 || mode >= 6) ? 32768 : 65536;
      printf("stages=%d  %-52s %7.1f cyc/stage  %6.1f B/cyc/SM   first %d issues took %lld cyc   %s\n", stages,
             names[mode], per, bytes / per, stages, fi, cudaGetErrorString(e));
    }
  return 0;
}
