// Micro-benchmark: tcgen05.mma.kind::f16 issue rate from shared-memory operands (SS mode), M = 128, N in {64,128,256},
// one CTA per SM, one issuing thread.  Answers: how many SM cycles does one K=16 MMA cost per N, with and without
// concurrent TMA-like traffic into shared memory (other warps storing to unrelated smem while the MMAs run)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I vfs_b200/csrc tools/ubench/mma_rate.cu -o mma_rate
#include <cstdio>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace vfs;

template <int N>
__global__ void __launch_bounds__(160, 1) mma_rate_kernel(long long* cycles, int iters, int distinct_stages,
                                                          int store_traffic) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 200 * 1024;
  const uint32_t tptr = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // zero the operand area (values irrelevant for timing, but keep them finite)
  for (uint32_t o = threadIdx.x * 16; o < 200 * 1024; o += blockDim.x * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + o), "r"(0));
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr));
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16_f32(128, N);
      unsigned long long g0, g1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t st = base + (it % distinct_stages) * (64 * 1024);
        const uint32_t a = st, b = st + 32 * 1024;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = umma_desc_sw128_kmajor(a + k * 32);
          const uint64_t da2 = umma_desc_sw128_kmajor(a + 16 * 1024 + k * 32);
          const uint64_t db = umma_desc_sw128_kmajor(b + k * 32);
          const uint64_t db2 = umma_desc_sw128_kmajor(b + N * 128 + k * 32);
          umma_f16(tmem_base, da2, db, idesc, 1u);
          umma_f16(tmem_base, da, db2, idesc, 1u);
          umma_f16(tmem_base, da, db, idesc, 1u);
        }
      }
      umma_commit(bar);
      mbar_wait(bar, 0, 1);
      const long long t1 = clock64();
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
      cycles[blockIdx.x] = t1 - t0;
      cycles[148 + blockIdx.x] = static_cast<long long>(g1 - g0);  // wall ns of the same interval
      stop = 1;
    }
  } else if (store_traffic) {
    // warps 1..4: stream 16-byte stores over a 64 KB window that the MMAs do not read
    const uint32_t win = base + 3 * 64 * 1024 - 64 * 1024 + 0;  // stage 2 region (unused when distinct_stages <= 2)
    uint32_t o = (threadIdx.x - 32) * 16;
    while (!stop) {
#pragma unroll 8
      for (int r = 0; r < 8; ++r) {
        asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(win + o), "r"(r));
        o = (o + 128 * 16) & (64 * 1024 - 1);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int N>
void run(int iters, int stages, int traffic) {
  long long* d;
  cudaMalloc(&d, 2 * 148 * sizeof(long long));
  const int smem = 200 * 1024 + 1024 + 64;
  cudaFuncSetAttribute(mma_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) mma_rate_kernel<N><<<148, 160, smem>>>(d, iters, stages, traffic);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2 * 148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1LL << 60;
  for (int i = 0; i < 148; ++i) {
    mx = h[i] > mx ? h[i] : mx;
    mn = h[i] < mn ? h[i] : mn;
  }
  const double per = static_cast<double>(mx) / (iters * 12.0);
  printf("N=%3d stages=%d store_traffic=%d iters=%d: %.1f cyc per MMA (K=16)  -> %.0f FLOP/cyc/SM  [min CTA %.1f]  "
         "SM clock during the run %.0f MHz (%.2f ms)  -> %.0f TFLOP/s chip  %s\n", N,
         stages, traffic, iters, per, 2.0 * 128 * N * 16 / per, static_cast<double>(mn) / (iters * 12.0),
         1e3 * static_cast<double>(h[0]) / static_cast<double>(h[148]), h[148] * 1e-6,
         148.0 * 2.0 * 128 * N * 16 * iters * 12.0 / static_cast<double>(h[148]) * 1e-3, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int traffic = 0; traffic < 2; ++traffic) {
    run<64>(2000, 2, traffic);
    run<128>(2000, 2, traffic);
    run<256>(2000, 2, traffic);
  }
  run<128>(2000, 1, 0);
  // sustained: does the chip hold its clock under all-SM tensor load?
  run<256>(200, 2, 0);
  run<256>(20000, 2, 0);
  run<256>(200000, 2, 0);
  return 0;
}
