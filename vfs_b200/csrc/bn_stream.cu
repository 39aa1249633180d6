// BatchNorm backward (reduce / apply) as bulk-copy pipelined streaming kernels.
//
// Both passes are pure HBM streams (10 bytes in per element for the reduce; 10 in + 4..8 out for the apply).  The
// register-resident versions in train.cu top out at 3.5-4.4 TB/s: a thread has its loads in flight only while it waits
// for them, and at ~100 registers per thread only 16 warps fit an SM.  Here ONE producer thread per CTA keeps a 4-stage
// ring of shared memory full with cp.async.bulk (TMA's linear form, 8 KB per stream and stage -> up to 160 KB in flight
// per SM, independent of how many threads compute), 8 consumer warps read the staged tiles with conflict-free 16-byte
// loads, and hand the slot back through an mbarrier.  Tiles are runs of 4096 consecutive elements of the [M][C] tensors;
// C divides 2048, so a thread's eight channels are the same in every tile and its per-channel constants stay in
// registers.
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"

namespace vfs {

namespace {

constexpr int kTile = 4096;                    // elements per tile (8 KB per fp16 stream)
constexpr int kStages = 4;
constexpr int kMaxStreams = 5;                 // dy hi / lo, z hi / lo, y hi
constexpr int kStreamBytes = kTile * 2;
constexpr int kStageBytes = kMaxStreams * kStreamBytes;
constexpr int kConsumers = 256;
constexpr int kThreads = kConsumers + 32;      // + the producer warp
constexpr int kSmemBytes = kStages * kStageBytes + 256 + 1024;

struct StreamArgs {
  const h16* in[kMaxStreams];                  // dy_hi, dy_lo, z_hi, z_lo, y_hi (y_hi may be null: no ReLU)
  int nin;
  const float* mean; const float* invstd; const float* gamma;
  double* sums; double count;
  h16* dz_hi; h16* dz_lo; h16* g_hi; h16* g_lo;
  float* dgamma; float* dbeta; int accumulate; float param_scale; int eval_mode;
  long long total;                             // M * C
  int C;
};

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint4 lds16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void unpack8s(const uint4 h, const uint4 l, float (&v)[8]) {
  v[0] = lo16_to_float(h.x) + lo16_to_float(l.x); v[1] = hi16_to_float(h.x) + hi16_to_float(l.x);
  v[2] = lo16_to_float(h.y) + lo16_to_float(l.y); v[3] = hi16_to_float(h.y) + hi16_to_float(l.y);
  v[4] = lo16_to_float(h.z) + lo16_to_float(l.z); v[5] = hi16_to_float(h.z) + hi16_to_float(l.z);
  v[6] = lo16_to_float(h.w) + lo16_to_float(l.w); v[7] = hi16_to_float(h.w) + hi16_to_float(l.w);
}
__device__ __forceinline__ void pack8s(const float (&v)[8], uint4& h, uint4& l) {
  split16x2(v[0], v[1], h.x, l.x);
  split16x2(v[2], v[3], h.y, l.y);
  split16x2(v[4], v[5], h.z, l.z);
  split16x2(v[6], v[7], h.w, l.w);
}

template <bool APPLY>
__global__ void __launch_bounds__(kThreads, 1) bn_bwd_stream_kernel(const StreamArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + kStages * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), kConsumers / 32);
    }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();   // launched with launch_pdl: the barrier set-up above overlaps the tail of the previous kernel
  const long long num_tiles = (a.total + kTile - 1) / kTile;

  if (warp == kConsumers / 32) {
    // ======================= producer =======================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(empty_bar(stage), phase ^ 1u, 900 + stage);
        const long long e0 = tile * kTile;
        const long long left = a.total - e0;
        const uint32_t bytes = static_cast<uint32_t>((left < kTile ? left : kTile) * 2);
        mbar_arrive_expect_tx(full_bar(stage), bytes * a.nin);
        const uint32_t dst = base + stage * kStageBytes;
        for (int k = 0; k < a.nin; ++k) bulk_load(dst + k * kStreamBytes, a.in[k] + e0, bytes, full_bar(stage));
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    // ======================= consumers =======================
    const int ct = threadIdx.x;                 // 0..255; pieces at tile elements ct*8 and 2048 + ct*8
    const int c = (ct * 8) % a.C;               // (2048 % C == 0: both pieces see the same channels)
    float mu[8], is[8], ga[8], sgm[8], sxm[8], sg[8], sx[8];
    const float inv_count = static_cast<float>(1.0 / a.count);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      mu[e] = a.mean[c + e];
      is[e] = a.invstd[c + e];
      sg[e] = sx[e] = 0.0f;
      if (APPLY) {
        ga[e] = (a.gamma ? a.gamma[c + e] : 1.0f) * is[e];
        sgm[e] = a.eval_mode ? 0.0f : static_cast<float>(a.sums[c + e]) * inv_count;
        sxm[e] = a.eval_mode ? 0.0f : static_cast<float>(a.sums[a.C + c + e]) * inv_count;
      }
    }
    if (APPLY && blockIdx.x == 0 && ct < a.C / 8 && (a.dgamma || a.dbeta) && a.sums != nullptr) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float dg = static_cast<float>(a.sums[a.C + c + e]) * a.param_scale;
        const float db = static_cast<float>(a.sums[c + e]) * a.param_scale;
        if (a.dgamma) a.dgamma[c + e] = a.accumulate ? a.dgamma[c + e] + dg : dg;
        if (a.dbeta) a.dbeta[c + e] = a.accumulate ? a.dbeta[c + e] + db : db;
      }
    }
    int stage = 0;
    uint32_t phase = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(full_bar(stage), phase, 950 + stage);
      const uint32_t src = base + stage * kStageBytes;
      const long long e0 = tile * kTile;
#pragma unroll
      for (int piece = 0; piece < 2; ++piece) {
        const int off = piece * 2048 + ct * 8;
        if (e0 + off < a.total) {
          const uint32_t so = static_cast<uint32_t>(off) * 2u;
          float g[8], zz[8];
          unpack8s(lds16(src + so), lds16(src + kStreamBytes + so), g);
          unpack8s(lds16(src + 2 * kStreamBytes + so), lds16(src + 3 * kStreamBytes + so), zz);
          if (a.nin == 5) {
            const uint4 yh = lds16(src + 4 * kStreamBytes + so);
            const float y[8] = {lo16_to_float(yh.x), hi16_to_float(yh.x), lo16_to_float(yh.y), hi16_to_float(yh.y),
                                lo16_to_float(yh.z), hi16_to_float(yh.z), lo16_to_float(yh.w), hi16_to_float(yh.w)};
#pragma unroll
            for (int e = 0; e < 8; ++e) g[e] = (y[e] > 0.0f) ? g[e] : 0.0f;
          }
          if (APPLY) {
            float dz[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float xhat = (zz[e] - mu[e]) * is[e];
              dz[e] = ga[e] * (g[e] - sgm[e] - xhat * sxm[e]);
            }
            uint4 h, l;
            pack8s(dz, h, l);
            *reinterpret_cast<uint4*>(a.dz_hi + e0 + off) = h;
            *reinterpret_cast<uint4*>(a.dz_lo + e0 + off) = l;
            if (a.g_hi) {
              pack8s(g, h, l);
              *reinterpret_cast<uint4*>(a.g_hi + e0 + off) = h;
              *reinterpret_cast<uint4*>(a.g_lo + e0 + off) = l;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              sg[e] += g[e];
              sx[e] = fmaf(g[e], (zz[e] - mu[e]) * is[e], sx[e]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(stage));
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1u;
      }
    }
    if (!APPLY) {
      // fold the 256 consumers (threads ct, ct + C/8, ... share channels) in the (now idle) stage memory
      asm volatile("bar.sync 1, 256;" ::: "memory");
      float* red = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)));
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        red[ct * 17 + e] = sg[e];
        red[ct * 17 + 8 + e] = sx[e];
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int pieces = a.C / 8;               // <= 256
      const int groups = kConsumers / pieces;
      for (int v = ct; v < 2 * a.C; v += kConsumers) {
        const int which = v / a.C, cc = v - which * a.C;
        const int pc = cc >> 3, e = (cc & 7) + 8 * which;
        double t = 0.0;
        for (int r = 0; r < groups; ++r) t += static_cast<double>(red[(r * pieces + pc) * 17 + e]);
        atomicAdd(a.sums + v, t);
      }
    }
  }
}

bool stream_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("VFS_BN_STREAM");
    on = (e && atoi(e) == 0) ? 0 : 1;
  }
  return on != 0;
}

template <bool APPLY>
int launch_stream(const StreamArgs& a, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    VFS_CUDA_OK(cudaFuncSetAttribute(bn_bwd_stream_kernel<APPLY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kSmemBytes));
    configured = true;
  }
  const long long num_tiles = (a.total + kTile - 1) / kTile;
  long long grid = device_sm_count();
  if (grid > num_tiles) grid = num_tiles;
  VFS_CUDA_OK(launch_pdl(bn_bwd_stream_kernel<APPLY>, dim3(static_cast<unsigned>(grid)), dim3(kThreads), kSmemBytes, s, a));
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

}  // namespace

// true when the streaming form can take the call (split dy and z, optional split y, split outputs, C | 2048)
bool bn_stream_eligible(const void* dy_split, const float* dy_f32, const float* y_f32, const float* z, const void* z_split,
                        float* dz_f32, long long M, int C) {
  return stream_enabled() && dy_split && !dy_f32 && !y_f32 && !z && z_split && !dz_f32 && C >= 64 && C <= 2048 &&
         2048 % C == 0 && M * C >= 4 * kTile;
}

int bn_bwd_reduce_stream(const void* dy_split, const void* y_split, const void* z_split, const float* mean,
                         const float* invstd, double* sums, long long M, int C, cudaStream_t s) {
  StreamArgs a;
  memset(&a, 0, sizeof(a));
  const long long plane = M * C;
  a.in[0] = reinterpret_cast<const h16*>(dy_split);
  a.in[1] = a.in[0] + plane;
  a.in[2] = reinterpret_cast<const h16*>(z_split);
  a.in[3] = a.in[2] + plane;
  a.nin = 4;
  if (y_split) {
    a.in[4] = reinterpret_cast<const h16*>(y_split);
    a.nin = 5;
  }
  a.mean = mean; a.invstd = invstd; a.sums = sums; a.count = 1.0; a.total = plane; a.C = C;
  return launch_stream<false>(a, s);
}

int bn_bwd_apply_stream(const void* dy_split, const void* y_split, const void* z_split, const float* mean,
                        const float* invstd, const float* gamma, const double* sums, double count, void* dz_split,
                        void* g_split, float* dgamma, float* dbeta, int accumulate, float param_scale, int eval_mode,
                        long long M, int C, cudaStream_t s) {
  StreamArgs a;
  memset(&a, 0, sizeof(a));
  const long long plane = M * C;
  a.in[0] = reinterpret_cast<const h16*>(dy_split);
  a.in[1] = a.in[0] + plane;
  a.in[2] = reinterpret_cast<const h16*>(z_split);
  a.in[3] = a.in[2] + plane;
  a.nin = 4;
  if (y_split) {
    a.in[4] = reinterpret_cast<const h16*>(y_split);
    a.nin = 5;
  }
  a.mean = mean; a.invstd = invstd; a.gamma = gamma; a.sums = const_cast<double*>(sums); a.count = count;
  a.dz_hi = reinterpret_cast<h16*>(dz_split);
  a.dz_lo = a.dz_hi + plane;
  if (g_split) {
    a.g_hi = reinterpret_cast<h16*>(g_split);
    a.g_lo = a.g_hi + plane;
  }
  a.dgamma = dgamma; a.dbeta = dbeta; a.accumulate = accumulate; a.param_scale = param_scale; a.eval_mode = eval_mode;
  a.total = plane; a.C = C;
  return launch_stream<true>(a, s);
}

}  // namespace vfs
