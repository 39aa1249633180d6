// tcgen05 weight-gradient convolution for sm_100a (training backward):
//
//   dW[co, tap, ci] = sum_{n, ho, wo} dZ[n, ho, wo, co] * X[n, ho*s + r*d - p, wo*s + q*d - p, ci]
//
// GEMM view: M = Cout (128 per tile), N = Cin (128 or 64 per tile), K = output pixels walked in chunks of 64.
// Both operands are pixel-major in memory (NHWC: channels contiguous), i.e. *MN-major* for this GEMM, so the very
// same TMA boxes the forward kernel uses ({64 channels, pixel tile}, 128B swizzle) are consumed by the tensor core
// through MN-major shared-memory descriptors -- no transposition pass.  Operands are split-fp16 (3 MMAs per
// product).  Work is split over (Cout block, Cin block, filter tap, pixel range); every unit reduces its fp32
// partial into dW with vector atomics (red.global.add.v4.f32).  Same warp roles as conv_tc.cu.
#include "host_common.h"
#include "ptx.cuh"

namespace vfs {

constexpr int kWgThreads = 192;
constexpr int kWgStages = 3;
constexpr int kWgPixels = 64;                          // K chunk: 64 pixels = 64 rows of 128 B
constexpr int kWgBoxBytes = 2 * kWgPixels * 128;       // one TMA box: hi + lo plane of [64 px][64 ch]  (16 KB)
constexpr int kWgStageBytes = 4 * kWgBoxBytes;         // 2 Cout blocks (A) + up to 2 Cin blocks (B)     (64 KB)
constexpr int kWgSmemBytes = kWgStages * kWgStageBytes + 256 + 1024;
constexpr int kWgMaxTaps = 9;
constexpr int kWgMaxViews = 4;

struct alignas(64) WgradParams {
  CUtensorMap tmap_dz;               // {Cout, Wo, Ho, N, 2}
  CUtensorMap tmap_x[kWgMaxViews];   // stride-parity views of X {Cin, Wv, Hv, N, 2}
  int tiles_w, tiles_h, num_chunks;  // pixel tiles of 64: chunk = (tn_i*tiles_h + th_i)*tiles_w + tw_i
  int tw, th, tn;
  int co_blocks, ci_blocks, bn;  // bn = 128 or 64 (Cin columns per tile)
  int num_taps, splits, chunks_per_split, num_units;
  int tap_view[kWgMaxTaps], tap_dh[kWgMaxTaps], tap_dw[kWgMaxTaps];
  int Cin, Cout, taps_total;
  float* dw;  // fp32 [Cout][taps][Cin], accumulated
  float out_scale;  // applied to the partial tiles before the atomics (1 when an unpack pass scales instead)
};

struct WgUnit {
  int co_blk, ci_blk, tap, c_begin, c_end;
};

__device__ __forceinline__ WgUnit wg_decode(const WgradParams& p, int unit) {
  WgUnit u;
  const int s = unit % p.splits;
  int r = unit / p.splits;
  u.tap = r % p.num_taps;
  r /= p.num_taps;
  u.ci_blk = r % p.ci_blocks;
  u.co_blk = r / p.ci_blocks;
  u.c_begin = min(p.num_chunks, s * p.chunks_per_split);
  u.c_end = min(p.num_chunks, u.c_begin + p.chunks_per_split);
  return u;
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kWgStages * kWgStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kWgStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kWgStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kWgStages + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * kWgStages + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 256;  // 2 accumulator stages x 128 columns

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_dz);
    for (int v = 0; v < kWgMaxViews; ++v) tma_prefetch_desc(&p.tmap_x[v]);
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_addr, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
  pdl_wait();   // launched with launch_pdl: everything above overlaps the tail of the previous kernel

  const int nb = p.bn / 64;                                        // Cin boxes per stage (1 or 2)
  const uint32_t stage_tx = (2 + nb) * kWgBoxBytes;                // bytes landing per stage

  if (warp == 0) {
    // ======================= TMA producer =======================
    int stage = 0;
    uint32_t phase = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      const WgUnit u = wg_decode(p, unit);
      const int view = p.tap_view[u.tap], dh = p.tap_dh[u.tap], dw = p.tap_dw[u.tap];
      for (int c = u.c_begin; c < u.c_end; ++c) {
        const int tw_i = c % p.tiles_w;
        const int t2 = c / p.tiles_w;
        const int w0 = tw_i * p.tw, h0 = (t2 % p.tiles_h) * p.th, n0 = (t2 / p.tiles_h) * p.tn;
        mbar_wait(empty_bar(stage), phase ^ 1u, 100 + stage);
        if (lane == 0) {
          const uint32_t sa = smem_base + stage * kWgStageBytes;
          mbar_arrive_expect_tx(full_bar(stage), stage_tx);
          tma_load_5d(sa, &p.tmap_dz, full_bar(stage), u.co_blk * 128, w0, h0, n0, 0);
          tma_load_5d(sa + kWgBoxBytes, &p.tmap_dz, full_bar(stage), u.co_blk * 128 + 64, w0, h0, n0, 0);
          for (int j = 0; j < nb; ++j)
            tma_load_5d(sa + (2 + j) * kWgBoxBytes, &p.tmap_x[view], full_bar(stage), u.ci_blk * p.bn + j * 64,
                        w0 + dw, h0 + dh, n0, 0);
        }
        __syncwarp();
        if (++stage == kWgStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    const uint32_t idesc = umma_idesc_f16_f32_mn(128, p.bn);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      const WgUnit u = wg_decode(p, unit);
      if (u.c_begin >= u.c_end) continue;  // empty unit: nothing to accumulate, nothing to reduce
      mbar_wait(tempty_bar(as), aphase ^ 1u, 200 + as);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * 128;
      for (int c = u.c_begin; c < u.c_end; ++c) {
        mbar_wait(full_bar(stage), phase, 300 + stage);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_base + stage * kWgStageBytes;
          // box layout: [hi plane 8 KB][lo plane 8 KB]; the two 64-channel blocks of an operand are one box apart
          const uint32_t a_hi = sa, a_lo = sa + kWgPixels * 128;
          const uint32_t b_hi = sa + 2 * kWgBoxBytes, b_lo = b_hi + kWgPixels * 128;
#pragma unroll
          for (int k = 0; k < kWgPixels / 16; ++k) {
            const uint32_t koff = k * 16 * 128;  // 16 pixel rows of 128 B
            const uint64_t da_hi = umma_desc_sw128_mnmajor(a_hi + koff, kWgBoxBytes);
            const uint64_t da_lo = umma_desc_sw128_mnmajor(a_lo + koff, kWgBoxBytes);
            const uint64_t db_hi = umma_desc_sw128_mnmajor(b_hi + koff, kWgBoxBytes);
            const uint64_t db_lo = umma_desc_sw128_mnmajor(b_lo + koff, kWgBoxBytes);
            umma_f16(d_tmem, da_lo, db_hi, idesc, (c > u.c_begin || k > 0) ? 1u : 0u);
            umma_f16(d_tmem, da_hi, db_lo, idesc, 1u);
            umma_f16(d_tmem, da_hi, db_hi, idesc, 1u);
          }
          umma_commit(empty_bar(stage));
          if (c == u.c_end - 1) umma_commit(tfull_bar(as));
        }
        __syncwarp();
        if (++stage == kWgStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1u;
      }
    }
  } else {
    // ======================= epilogue: reduce the fp32 partial tile into dW =======================
    const int q = warp & 3;
    const int row = q * 32 + lane;  // Cout row inside the tile
    int as = 0;
    uint32_t aphase = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      const WgUnit u = wg_decode(p, unit);
      if (u.c_begin >= u.c_end) continue;
      mbar_wait(tfull_bar(as), aphase, 400 + as);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 128;
      const int co = u.co_blk * 128 + row;
      const bool co_ok = co < p.Cout;  // Cout = 64: the upper half of the tile is TMA zero fill
      float* dst = p.dw + (static_cast<size_t>(co) * p.taps_total + u.tap) * p.Cin + u.ci_blk * p.bn;
#pragma unroll 1
      for (int c0 = 0; c0 < p.bn; c0 += 32) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(t_row + c0, acc);
        tmem_ld_wait();
        if (co_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4(dst + c0 + j, __uint_as_float(acc[j]) * p.out_scale, __uint_as_float(acc[j + 1]) * p.out_scale,
                       __uint_as_float(acc[j + 2]) * p.out_scale, __uint_as_float(acc[j + 3]) * p.out_scale);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (++as == 2) {
        as = 0;
        aphase ^= 1u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// [Cout][taps][Cin] fp32 -> OIHW fp32 (optionally accumulating into an existing gradient)
__global__ void wgrad_unpack_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout, int Cin, int kk,
                                    int accumulate, float out_scale) {
  pdl_wait();
  const size_t total = static_cast<size_t>(Cout) * Cin * kk;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // i indexes dst: ((co*Cin + ci)*kk + tap)
    const int tap = static_cast<int>(i % kk);
    const size_t t = i / kk;
    const int ci = static_cast<int>(t % Cin);
    const int co = static_cast<int>(t / Cin);
    const float v = src[(static_cast<size_t>(co) * kk + tap) * Cin + ci] * out_scale;
    dst[i] = accumulate ? dst[i] + v : v;
  }
}

namespace {
inline int floordiv_i(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

void choose_tile64(int Wo, int Ho, int N, int* tw, int* th, int* tn) {
  long best = -1;
  for (int a = 1; a <= 64; a <<= 1)
    for (int b = 1; a * b <= 64; b <<= 1) {
      const int c = 64 / (a * b);
      const long tiles = static_cast<long>((Wo + a - 1) / a) * ((Ho + b - 1) / b) * ((N + c - 1) / c);
      if (best < 0 || tiles < best) {
        best = tiles;
        *tw = a; *th = b; *tn = c;
      }
    }
}
}  // namespace

size_t wgrad_workspace_bytes(int Cout, int Cin, int ksize) {
  return static_cast<size_t>(Cout) * Cin * ksize * ksize * sizeof(float);
}

// d = forward descriptor.  x_split [N,H,W,Cin], dz_split [N,Ho,Wo,Cout], workspace fp32 [Cout][k*k][Cin],
// dw_oihw fp32 [Cout][Cin][k][k] = out_scale * dW (overwritten, or accumulated into when accumulate != 0).
int conv_wgrad_tc(const VfsConvDesc* d, const void* x_split, const void* dz_split, void* workspace,
                  float* dw_oihw, int accumulate, float out_scale, cudaStream_t stream) {
  VFS_REQUIRE(d && x_split && dz_split && workspace && dw_oihw, VFS_EINVAL, "conv_wgrad: null argument");
  VFS_REQUIRE(d->ksize == 1 || d->ksize == 3, VFS_ESHAPE, "conv_wgrad: ksize %d unsupported", d->ksize);
  VFS_REQUIRE(d->stride == 1 || d->stride == 2, VFS_ESHAPE, "conv_wgrad: stride %d unsupported", d->stride);
  VFS_REQUIRE(d->Cin % 64 == 0 && d->Cout % 64 == 0, VFS_ESHAPE,
              "conv_wgrad: Cin and Cout must be multiples of 64 (got %d, %d)", d->Cin, d->Cout);
  const int N = d->N, H = d->H, W = d->W, Cin = d->Cin, Cout = d->Cout;
  const int k = d->ksize, s = d->stride, dil = (k == 1) ? 1 : d->dilation;
  const int pad = (k == 1) ? 0 : dil;
  const int Ho = (H + 2 * pad - dil * (k - 1) - 1) / s + 1;
  const int Wo = (W + 2 * pad - dil * (k - 1) - 1) / s + 1;
  const size_t x_plane = static_cast<size_t>(N) * H * W * Cin;
  const size_t z_plane = static_cast<size_t>(N) * Ho * Wo * Cout;

  WgradParams p;
  memset(&p, 0, sizeof(p));
  const bool flat = (k == 1 && s == 1);
  int pW = Wo, pH = Ho, pN = N;
  if (flat) {
    p.tw = 64; p.th = 1; p.tn = 1;
    pW = N * H * W; pH = 1; pN = 1;
  } else {
    choose_tile64(Wo, Ho, N, &p.tw, &p.th, &p.tn);
  }
  p.tiles_w = (pW + p.tw - 1) / p.tw;
  p.tiles_h = (pH + p.th - 1) / p.th;
  p.num_chunks = p.tiles_w * p.tiles_h * ((pN + p.tn - 1) / p.tn);
  p.bn = (Cin % 128 == 0) ? 128 : 64;
  p.co_blocks = (Cout + 127) / 128;
  p.Cout = Cout;
  p.ci_blocks = Cin / p.bn;
  p.num_taps = k * k;
  p.taps_total = k * k;
  p.Cin = Cin;
  p.dw = reinterpret_cast<float*>(workspace);
  const int base_units = p.co_blocks * p.ci_blocks * p.num_taps;
  const int sms = device_sm_count();
  int splits = (2 * sms + base_units - 1) / base_units;
  if (splits > p.num_chunks) splits = p.num_chunks;
  if (splits < 1) splits = 1;
  p.chunks_per_split = (p.num_chunks + splits - 1) / splits;
  p.splits = (p.num_chunks + p.chunks_per_split - 1) / p.chunks_per_split;
  p.num_units = base_units * p.splits;

  const uint32_t box[5] = {64u, static_cast<uint32_t>(p.tw), static_cast<uint32_t>(p.th), static_cast<uint32_t>(p.tn), 2u};
  {
    int rc;
    if (flat) {
      const uint64_t dims[5] = {static_cast<uint64_t>(Cout), static_cast<uint64_t>(N) * Ho * Wo, 1, 1, 2};
      const uint64_t strides[4] = {static_cast<uint64_t>(Cout) * 2, z_plane * 2, z_plane * 2, z_plane * 2};
      rc = make_tmap_16b_sw128(&p.tmap_dz, dz_split, 5, dims, strides, box);
    } else {
      const uint64_t dims[5] = {static_cast<uint64_t>(Cout), static_cast<uint64_t>(Wo), static_cast<uint64_t>(Ho),
                                static_cast<uint64_t>(N), 2};
      const uint64_t strides[4] = {static_cast<uint64_t>(Cout) * 2, static_cast<uint64_t>(Wo) * Cout * 2,
                                   static_cast<uint64_t>(Ho) * Wo * Cout * 2, z_plane * 2};
      rc = make_tmap_16b_sw128(&p.tmap_dz, dz_split, 5, dims, strides, box);
    }
    if (rc != VFS_OK) return rc;
  }
  bool view_used[kWgMaxViews] = {false, false, false, false};
  for (int r = 0; r < k; ++r)
    for (int q = 0; q < k; ++q) {
      const int oh = r * dil - pad, ow = q * dil - pad;
      const int ph = ((oh % s) + s) % s, pw = ((ow % s) + s) % s;
      const int t = r * k + q;
      p.tap_view[t] = ph * 2 + pw;
      p.tap_dh[t] = floordiv_i(oh - ph, s);
      p.tap_dw[t] = floordiv_i(ow - pw, s);
      view_used[ph * 2 + pw] = true;
    }
  const char* x_base = reinterpret_cast<const char*>(x_split);
  int first_valid = -1;
  bool built[kWgMaxViews] = {false, false, false, false};
  for (int v = 0; v < kWgMaxViews; ++v) {
    const int ph = v / 2, pw = v % 2;
    int rc = VFS_OK;
    if (flat) {
      const uint64_t dims[5] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(N) * H * W, 1, 1, 2};
      const uint64_t strides[4] = {static_cast<uint64_t>(Cin) * 2, x_plane * 2, x_plane * 2, x_plane * 2};
      rc = make_tmap_16b_sw128(&p.tmap_x[v], x_base, 5, dims, strides, box);
      built[v] = true;
    } else if (view_used[v] && ph < H && pw < W) {
      const int Hv = (H - ph + s - 1) / s, Wv = (W - pw + s - 1) / s;
      const uint64_t dims[5] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(Wv), static_cast<uint64_t>(Hv),
                                static_cast<uint64_t>(N), 2};
      const uint64_t strides[4] = {static_cast<uint64_t>(s) * Cin * 2, static_cast<uint64_t>(s) * W * Cin * 2,
                                   static_cast<uint64_t>(H) * W * Cin * 2, x_plane * 2};
      rc = make_tmap_16b_sw128(&p.tmap_x[v], x_base + (static_cast<size_t>(ph) * W + pw) * Cin * 2, 5, dims, strides,
                                box);
      built[v] = true;
    }
    if (rc != VFS_OK) return rc;
    if (built[v] && first_valid < 0) first_valid = v;
  }
  VFS_REQUIRE(first_valid >= 0, VFS_EINVAL, "conv_wgrad: no usable input view");
  for (int v = 0; v < kWgMaxViews; ++v)
    if (!built[v]) p.tmap_x[v] = p.tmap_x[first_valid];

  // 1x1 convolutions: [Cout][1][Cin] IS the OIHW layout -- when the caller accumulates (the flat gradient buffer of
  // the training step, zeroed once per step) the partial tiles go straight into dW: no workspace memset, no unpack pass
  const bool direct = (k == 1) && accumulate != 0;
  p.out_scale = direct ? out_scale : 1.0f;
  if (direct) p.dw = dw_oihw;
  else VFS_CUDA_OK(cudaMemsetAsync(workspace, 0, wgrad_workspace_bytes(Cout, Cin, k), stream));
  static bool configured = false;
  if (!configured) {
    VFS_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes));
    configured = true;
  }
  const int grid = p.num_units < sms ? p.num_units : sms;
  VFS_CUDA_OK(launch_pdl(wgrad_tc_kernel, dim3(grid), dim3(kWgThreads), kWgSmemBytes, stream, p));
  if (direct) return VFS_OK;
  const size_t total = static_cast<size_t>(Cout) * Cin * k * k;
  const int blocks = static_cast<int>((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
  VFS_CUDA_OK(launch_pdl(wgrad_unpack_kernel, dim3(blocks), dim3(256), 0, stream, static_cast<const float*>(p.dw), dw_oihw,
                         Cout, Cin, k * k, accumulate, out_scale));
  return VFS_OK;
}

}  // namespace vfs
