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

// ------------------------------------------------------------------------------------------------
// Tensor-core stem: the same 7x7/s2 convolution as an implicit GEMM on tcgen05.
//   M = 128 conv-output pixels (4 rows x 32 columns), N = 64 output channels, K = 147 padded to 192 (3 chunks of 64).
// The A operand (im2col tile) cannot come from TMA (Cin = 3, stride 2), so the CTA builds it: the fp32 input patch
// is staged in shared memory, every thread gathers 8 consecutive K values of one pixel row, splits them into hi/lo
// halves and stores the two 16-byte chunks at the 128B-swizzled position the UMMA descriptor expects.  The split
// weights [2][64][192] are brought once per CTA by TMA.  36 MMAs (128x64x16) per tile replace 128 x 9408 FMAs.
// ------------------------------------------------------------------------------------------------
constexpr int kTcRows = 4, kTcCols = 32;                 // pixel tile
constexpr int kTcPatchH = kTcRows * 2 + 5;               // 13 input rows
constexpr int kTcPatchW = 72;                            // 32*2+5 = 69 -> padded
constexpr int kTcKPad = 192;
constexpr int kTcABytes = 3 * 2 * 128 * 128;             // 3 K-chunks x (hi, lo) x 128 rows x 128 B = 96 KB
constexpr int kTcBBytes = 3 * 2 * 64 * 128;              // 48 KB
constexpr int kTcPatchBytes = 3 * kTcPatchH * kTcPatchW * 4;
constexpr int kTcSmemBytes = kTcABytes + kTcBBytes + kTcPatchBytes + 64 + 1024;

struct alignas(64) StemTcParams {
  CUtensorMap tmap_w;  // {192, 64, 2} split weights
  const float* in;     // NCHW fp32
  const float* scale;  // nullable: raw output
  const float* shift;
  float* out;          // fp32 NHWC [N, Hc, Wc, 64]
  int N, H, W, Hc, Wc, tiles_x, tiles_y, num_tiles;
};

constexpr int kTcThreads = 512;                          // 4 thread groups of 128 (one per pixel row of the tile)
constexpr int kTcPatchElems = 3 * kTcPatchH * kTcPatchW;
constexpr int kTcPatchPerThread = (kTcPatchElems + kTcThreads - 1) / kTcThreads;

template <int G>
__device__ __forceinline__ void stem_build_row(const float* __restrict__ patch, uint32_t a_base, int m, int py, int px) {
  // chunks j = 4*i + G (16-byte chunks of 8 consecutive k); everything about k is a compile-time constant
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int j = 4 * i + G;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = j * 8 + e;
      if (k < 147) {
        const int c = k / 49, r = (k / 7) % 7, s = k % 7;
        v[e] = patch[(c * kTcPatchH + 2 * py + r) * kTcPatchW + 2 * px + s];
      } else {
        v[e] = 0.0f;
      }
    }
    uint4 h, l;
    split16x2(v[0], v[1], h.x, l.x);
    split16x2(v[2], v[3], h.y, l.y);
    split16x2(v[4], v[5], h.z, l.z);
    split16x2(v[6], v[7], h.w, l.w);
    const int kc = j >> 3, c16 = j & 7;
    const uint32_t addr = a_base + kc * (2 * 128 * 128) + m * 128 + ((c16 ^ (m & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(h.x), "r"(h.y), "r"(h.z), "r"(h.w) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + 128 * 128), "r"(l.x), "r"(l.y), "r"(l.z), "r"(l.w)
                 : "memory");
  }
}

__global__ void __launch_bounds__(kTcThreads, 1) stem_tc_kernel(const __grid_constant__ StemTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;                        // [3][hi|lo][128 rows][128 B]
  const uint32_t b_base = smem_base + kTcABytes;            // [3][hi|lo][64 rows][128 B]
  const uint32_t patch_addr = b_base + kTcBBytes;
  float* patch = reinterpret_cast<float*>(smem_raw + (patch_addr - smem_u32(smem_raw)));
  const uint32_t bar_w = patch_addr + kTcPatchBytes;        // weights landed
  const uint32_t bar_mma = bar_w + 8;                       // MMAs of the current tile retired
  const uint32_t tmem_ptr_addr = bar_w + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmap_w);
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_addr, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_w, kTcBBytes);
    for (int kc = 0; kc < 3; ++kc) tma_load_3d(b_base + kc * (2 * 64 * 128), &p.tmap_w, bar_w, kc * 64, 0, 0);
  }
  mbar_wait(bar_w, 0, 500);

  const int m = threadIdx.x & 127, g = threadIdx.x >> 7;  // pixel row of the tile, chunk group (j mod 4)
  const int py = m >> 5, px = m & 31;
  const int q = warp & 3, part = warp >> 2;               // epilogue: TMEM lane quarter, 16-column quarter
  constexpr uint32_t idesc = umma_idesc_f16_f32(128, 64);
  uint32_t mma_phase = 0;

  // input patch of a tile, global -> registers (the loads of tile i+1 are in flight while tile i is multiplied and
  // written out; they reach shared memory once the im2col build of tile i no longer reads the patch buffer)
  float pre[kTcPatchPerThread];
  auto load_patch = [&](int tile) {
    const int tx = tile % p.tiles_x;
    const int t2 = tile / p.tiles_x;
    const int ty = t2 % p.tiles_y;
    const int n = t2 / p.tiles_y;
    const int iy0 = ty * kTcRows * 2 - 3, ix0 = tx * kTcCols * 2 - 3;
    const float* src = p.in + static_cast<size_t>(n) * 3 * p.H * p.W;
#pragma unroll
    for (int u = 0; u < kTcPatchPerThread; ++u) {
      const int i = threadIdx.x + u * kTcThreads;
      const int pc = i % kTcPatchW;
      const int t = i / kTcPatchW;
      const int pr = t % kTcPatchH;
      const int c = t / kTcPatchH;
      const int iy = iy0 + pr, ix = ix0 + pc;
      pre[u] = (i < kTcPatchElems && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
                   ? __ldg(src + (static_cast<size_t>(c) * p.H + iy) * p.W + ix)
                   : 0.0f;
    }
  };
  auto store_patch = [&]() {
#pragma unroll
    for (int u = 0; u < kTcPatchPerThread; ++u) {
      const int i = threadIdx.x + u * kTcThreads;
      if (i < kTcPatchElems) patch[i] = pre[u];
    }
  };
  if (static_cast<int>(blockIdx.x) < p.num_tiles) {
    load_patch(blockIdx.x);
    store_patch();
  }
  __syncthreads();

  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const int tx = tile % p.tiles_x;
    const int t2 = tile / p.tiles_x;
    const int ty = t2 % p.tiles_y;
    const int n = t2 / p.tiles_y;
    const int oy0 = ty * kTcRows, ox0 = tx * kTcCols;
    if (g == 0) stem_build_row<0>(patch, a_base, m, py, px);
    else if (g == 1) stem_build_row<1>(patch, a_base, m, py, px);
    else if (g == 2) stem_build_row<2>(patch, a_base, m, py, px);
    else stem_build_row<3>(patch, a_base, m, py, px);
    fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    const int next_tile = tile + gridDim.x;
    if (next_tile < p.num_tiles) load_patch(next_tile);   // in flight during the MMAs and the epilogue
    if (threadIdx.x == 0) {
      tc_fence_after();
#pragma unroll
      for (int kc = 0; kc < 3; ++kc) {
        const uint32_t a_hi = a_base + kc * (2 * 128 * 128), a_lo = a_hi + 128 * 128;
        const uint32_t b_hi = b_base + kc * (2 * 64 * 128), b_lo = b_hi + 64 * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t koff = k * 32;
          umma_f16(tmem_base, umma_desc_sw128_kmajor(a_lo + koff), umma_desc_sw128_kmajor(b_hi + koff), idesc,
                   (kc | k) != 0 ? 1u : 0u);
          umma_f16(tmem_base, umma_desc_sw128_kmajor(a_hi + koff), umma_desc_sw128_kmajor(b_lo + koff), idesc, 1u);
          umma_f16(tmem_base, umma_desc_sw128_kmajor(a_hi + koff), umma_desc_sw128_kmajor(b_hi + koff), idesc, 1u);
        }
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, mma_phase, 600);
    mma_phase ^= 1u;
    tc_fence_after();
    // epilogue: warp (q, part) reads rows 32q..32q+31, columns 16*part..+15
    {
      uint32_t acc[16];
      tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + part * 16, acc);
      tmem_ld_wait();
      const int r = q * 32 + lane;
      const int oy = oy0 + (r >> 5), ox = ox0 + (r & 31);
      if (oy < p.Hc && ox < p.Wc) {
        float4* dst = reinterpret_cast<float4*>(p.out + ((static_cast<size_t>(n) * p.Hc + oy) * p.Wc + ox) * 64 + part * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 y = make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                                 __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
          if (p.scale != nullptr) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + part * 16) + j);
            const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + part * 16) + j);
            y.x = fmaxf(fmaf(y.x, sc.x, sh.x), 0.0f);
            y.y = fmaxf(fmaf(y.y, sc.y, sh.y), 0.0f);
            y.z = fmaxf(fmaf(y.z, sc.z, sh.z), 0.0f);
            y.w = fmaxf(fmaf(y.w, sc.w, sh.w), 0.0f);
          }
          dst[j] = y;
        }
      }
    }
    if (next_tile < p.num_tiles) store_patch();  // the build of this tile finished reading the patch buffer long ago
    tc_fence_before();
    __syncthreads();  // TMEM drained, A free and the next patch in place
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient of the stem convolution on tcgen05:
//     dW[co][k] (+)= out_scale * sum_pixels dZ[pixel][co] * im2col[pixel][k]        (k = (c*7 + r)*7 + s)
// Per 128-pixel tile the CTA builds the same swizzled im2col tile as the forward kernel ([3 chunks][hi|lo][128 pixel
// rows][128 B]) and a split copy of the fp32 dZ tile ([hi|lo][128 pixel rows][64 co]).  With the pixel axis as the
// reduction both are MN-major operands (rows = reduction index, 128 B = 64 consecutive M / N elements per row), the
// layout csrc/wgrad_tc.cu feeds from TMA boxes: A = dZ^T (M = 128: the 64 channels plus a 64-row block that points
// at a zeroed region), B = im2col (N = 192: three 64-column blocks one chunk apart).  The fp32 accumulator
// [128 x 192] lives in TMEM for the whole life of the CTA; one epilogue adds it into dW with fp32 atomics.
// The SIMT kernel this replaces (train.cu stem_wgrad_kernel) took 0.49 ms per call at 32 x 224^2.
// ------------------------------------------------------------------------------------------------
static inline void stem_dims(int H, int W, int* Hc, int* Wc, int* Hp, int* Wp);
constexpr int kSwDzBytes = 2 * 128 * 128;                // hi, lo planes of [128 pixels][64 co]
constexpr int kSwZeroBytes = 128 * 128;
constexpr int kSwSmemBytes = kTcABytes + kSwDzBytes + kSwZeroBytes + kTcPatchBytes + 64 + 1024;

struct StemWgradParams {
  const float* in;   // NCHW fp32
  const float* dz;   // fp32 NHWC [N, Hc, Wc, 64]
  float* dw;         // fp32 [64][147]
  float out_scale;
  int N, H, W, Hc, Wc, tiles_x, tiles_y, num_tiles;
};

__global__ void __launch_bounds__(kTcThreads, 1) stem_wgrad_tc_kernel(const StemWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;                        // im2col [3][hi|lo][128 rows][128 B]
  const uint32_t dz_base = a_base + kTcABytes;              // dZ [hi|lo][128 rows][128 B]
  const uint32_t zero_base = dz_base + kSwDzBytes;          // 16 KB of zeros (upper 64 rows of the M = 128 operand)
  const uint32_t patch_addr = zero_base + kSwZeroBytes;
  float* patch = reinterpret_cast<float*>(smem_raw + (patch_addr - smem_u32(smem_raw)));
  const uint32_t bar_mma = patch_addr + kTcPatchBytes;
  const uint32_t tmem_ptr_addr = bar_mma + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_addr, 256);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < kSwZeroBytes / 16; i += kTcThreads)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(zero_base + i * 16), "r"(0u) : "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  const int m = threadIdx.x & 127, g = threadIdx.x >> 7;  // pixel row of the tile, chunk group
  const int py = m >> 5, px = m & 31;
  constexpr uint32_t idesc = umma_idesc_f16_f32_mn(128, 192);
  uint32_t mma_phase = 0;

  float pre[kTcPatchPerThread];
  float4 dpre[4];  // 16 channels (16 g .. 16 g + 15) of this thread's pixel
  auto load_tile = [&](int tile) {
    const int tx = tile % p.tiles_x;
    const int t2 = tile / p.tiles_x;
    const int ty = t2 % p.tiles_y;
    const int n = t2 / p.tiles_y;
    const int iy0 = ty * kTcRows * 2 - 3, ix0 = tx * kTcCols * 2 - 3;
    const float* src = p.in + static_cast<size_t>(n) * 3 * p.H * p.W;
#pragma unroll
    for (int u = 0; u < kTcPatchPerThread; ++u) {
      const int i = threadIdx.x + u * kTcThreads;
      const int pc = i % kTcPatchW;
      const int t = i / kTcPatchW;
      const int pr = t % kTcPatchH;
      const int c = t / kTcPatchH;
      const int iy = iy0 + pr, ix = ix0 + pc;
      pre[u] = (i < kTcPatchElems && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
                   ? __ldg(src + (static_cast<size_t>(c) * p.H + iy) * p.W + ix)
                   : 0.0f;
    }
    const int oy = ty * kTcRows + py, ox = tx * kTcCols + px;
    const bool ok = oy < p.Hc && ox < p.Wc;
    const float4* d4 =
        reinterpret_cast<const float4*>(p.dz + ((static_cast<size_t>(n) * p.Hc + (ok ? oy : 0)) * p.Wc + (ok ? ox : 0)) * 64 + 16 * g);
#pragma unroll
    for (int j = 0; j < 4; ++j) dpre[j] = ok ? __ldg(d4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int u = 0; u < kTcPatchPerThread; ++u) {
      const int i = threadIdx.x + u * kTcThreads;
      if (i < kTcPatchElems) patch[i] = pre[u];
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {  // two 16-byte chunks of 8 channels: c16 = 2 g + j
      const float4 a = dpre[2 * j], b = dpre[2 * j + 1];
      uint4 h, l;
      split16x2(a.x, a.y, h.x, l.x);
      split16x2(a.z, a.w, h.y, l.y);
      split16x2(b.x, b.y, h.z, l.z);
      split16x2(b.z, b.w, h.w, l.w);
      const int c16 = 2 * g + j;
      const uint32_t addr = dz_base + m * 128 + ((c16 ^ (m & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(h.x), "r"(h.y), "r"(h.z), "r"(h.w) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + 128 * 128), "r"(l.x), "r"(l.y), "r"(l.z), "r"(l.w)
                   : "memory");
    }
  };

  bool first = true;
  if (static_cast<int>(blockIdx.x) < p.num_tiles) load_tile(blockIdx.x);
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    store_tile();       // patch + dZ of this tile (the previous tile's MMAs have retired: see the wait below)
    __syncthreads();    // patch complete
    if (g == 0) stem_build_row<0>(patch, a_base, m, py, px);
    else if (g == 1) stem_build_row<1>(patch, a_base, m, py, px);
    else if (g == 2) stem_build_row<2>(patch, a_base, m, py, px);
    else stem_build_row<3>(patch, a_base, m, py, px);
    fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    const int next_tile = tile + gridDim.x;
    if (next_tile < p.num_tiles) load_tile(next_tile);   // global loads in flight during the MMAs
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t dz_hi = dz_base, dz_lo = dz_base + 128 * 128;
      const uint32_t b_hi = a_base, b_lo = a_base + 128 * 128;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t koff = k * 16 * 128;  // 16 pixel rows of 128 B
        const uint64_t da_hi = umma_desc_sw128_mnmajor(dz_hi + koff, zero_base - dz_hi);
        const uint64_t da_lo = umma_desc_sw128_mnmajor(dz_lo + koff, zero_base - dz_lo);
        const uint64_t db_hi = umma_desc_sw128_mnmajor(b_hi + koff, 2 * 128 * 128);
        const uint64_t db_lo = umma_desc_sw128_mnmajor(b_lo + koff, 2 * 128 * 128);
        umma_f16(tmem_base, da_lo, db_hi, idesc, (first && k == 0) ? 0u : 1u);
        umma_f16(tmem_base, da_hi, db_lo, idesc, 1u);
        umma_f16(tmem_base, da_hi, db_hi, idesc, 1u);
      }
      umma_commit(bar_mma);
    }
    first = false;
    mbar_wait(bar_mma, mma_phase, 700);   // operands free again (and, after the last tile, the accumulator final)
    mma_phase ^= 1u;
    tc_fence_after();
  }

  // epilogue: accumulator rows 0..63 = output channels (TMEM lanes 0..63: warps with warp % 4 in {0, 1}), 192 columns
  // split over the four warps that share a lane quarter
  if (!first && (warp & 3) < 2) {
    const int co = (warp & 3) * 32 + lane;
    const int part = warp >> 2;
#pragma unroll 1
    for (int c0 = part * 48; c0 < part * 48 + 48; c0 += 16) {
      uint32_t acc[16];
      tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + c0, acc);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c0 + j < 147) atomicAdd(p.dw + co * 147 + c0 + j, __uint_as_float(acc[j]) * p.out_scale);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

int stem_wgrad_tc(const float* x, const float* dz, float* dw, float out_scale, int N, int H, int W, cudaStream_t s) {
  StemWgradParams p;
  int Hc, Wc, Hp, Wp;
  stem_dims(H, W, &Hc, &Wc, &Hp, &Wp);
  p.in = x; p.dz = dz; p.dw = dw; p.out_scale = out_scale;
  p.N = N; p.H = H; p.W = W; p.Hc = Hc; p.Wc = Wc;
  p.tiles_x = (Wc + kTcCols - 1) / kTcCols;
  p.tiles_y = (Hc + kTcRows - 1) / kTcRows;
  p.num_tiles = p.tiles_x * p.tiles_y * N;
  static bool configured = false;
  if (!configured) {
    VFS_CUDA_OK(cudaFuncSetAttribute(stem_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSwSmemBytes));
    configured = true;
  }
  const int sms = device_sm_count();
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  stem_wgrad_tc_kernel<<<grid, kTcThreads, kSwSmemBytes, s>>>(p);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

// OIHW fp32 [64,3,7,7] -> split [2][64][192] (K = (c*7+r)*7+s zero-padded to 192)
__global__ void stem_pack_weight_kernel(const float* __restrict__ w, h16* __restrict__ hi, h16* __restrict__ lo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * kTcKPad) return;
  const int k = i % kTcKPad, co = i / kTcKPad;
  const float v = (k < 147) ? w[co * 147 + k] : 0.0f;
  h16 h, l;
  split16(v, h, l);
  hi[i] = h;
  lo[i] = l;
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

size_t stem_packed_weight_bytes() { return 2 * 64 * kTcKPad * sizeof(h16); }

int stem_pack_weight(const float* w, void* w_split, cudaStream_t s) {
  VFS_REQUIRE(w && w_split, VFS_EINVAL, "stem_pack_weight: null argument");
  h16* hi = reinterpret_cast<h16*>(w_split);
  stem_pack_weight_kernel<<<(64 * kTcKPad + 255) / 256, 256, 0, s>>>(w, hi, hi + 64 * kTcKPad);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

// weight_split: output of stem_pack_weight (tensor-core path)
static int stem_conv_launch(const float* in, const void* weight_split, const float* scale, const float* shift,
                            float* conv_out, int N, int H, int W, cudaStream_t s) {
  int Hc, Wc, Hp, Wp;
  stem_dims(H, W, &Hc, &Wc, &Hp, &Wp);
  StemTcParams p;
  memset(&p, 0, sizeof(p));
  const uint64_t dims[3] = {kTcKPad, 64, 2};
  const uint64_t strides[2] = {kTcKPad * 2, kTcKPad * 64 * 2};
  const uint32_t box[3] = {64, 64, 2};
  int rc = make_tmap_16b_sw128(&p.tmap_w, weight_split, 3, dims, strides, box);
  if (rc != VFS_OK) return rc;
  p.in = in; p.scale = scale; p.shift = shift; p.out = conv_out;
  p.N = N; p.H = H; p.W = W; p.Hc = Hc; p.Wc = Wc;
  p.tiles_x = (Wc + kTcCols - 1) / kTcCols;
  p.tiles_y = (Hc + kTcRows - 1) / kTcRows;
  p.num_tiles = p.tiles_x * p.tiles_y * N;
  static bool configured = false;
  if (!configured) {
    VFS_CUDA_OK(cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
    configured = true;
  }
  const int sms = device_sm_count();
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  stem_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, s>>>(p);
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

int stem_forward(const float* in, const void* weight, const float* scale, const float* shift, void* out_split,
                 void* workspace, int N, int H, int W, cudaStream_t s) {
  VFS_REQUIRE(in && weight && scale && shift && out_split && workspace, VFS_EINVAL, "stem_forward: null argument");
  VFS_REQUIRE(N > 0 && H >= 7 && W >= 7, VFS_ESHAPE, "stem_forward: input %dx%dx%d too small", N, H, W);
  float* conv_out = reinterpret_cast<float*>(workspace);
  int rc = stem_conv_launch(in, weight, scale, shift, conv_out, N, H, W, s);
  if (rc != VFS_OK) return rc;
  return stem_pool_launch(conv_out, nullptr, nullptr, out_split, N, H, W, s);
}

// train-mode pieces: raw conv output (statistics are computed on it), then BN+ReLU fused into the max-pool
int stem_conv_raw(const float* in, const void* weight, void* conv_out, int N, int H, int W, cudaStream_t s) {
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
