// tcgen05 implicit-GEMM convolution with fused BN(eval)/residual/ReLU epilogue for sm_100a.
//
//   out[pix, co] = act( scale[co] * sum_{tap, c} in[pix + off(tap), c] * w[co, tap, c] + shift[co] (+ res[pix, co]) )
//
// GEMM view: M = output pixels (tiles of 128 = tn x th x tw pixels), N = Cout (tiles of BN),
// K = taps * Cin walked in chunks of 64 channels of one filter tap.
//
// Operands are "split fp16" (x = hi + lo, both IEEE half): each K-chunk issues three tcgen05.mma products
// hi*hi + hi*lo + lo*hi into the same fp32 TMEM accumulator, which restores ~2^-22 relative operand
// accuracy (the reference runs fp32; parity bar is 1e-3 after ~50 stacked layers).
//
// Warp roles (352 threads, persistent over output tiles):
//   warp 0     TMA producer: per K-chunk one 5-D box load of the activation tile (both planes; the filter
//              tap is a coordinate offset, image borders are TMA out-of-bounds zero fill) and one 3-D box
//              load of the weight tile, into a STAGES-deep 128B-swizzled shared-memory ring.
//   warp 1     TMEM allocator + MMA issuer (one elected lane), accumulators double-buffered in TMEM.
//   warps 2-9  epilogue math: tcgen05.ld -> scale/shift (+residual) (ReLU) -> split -> swizzled staging buffer
//              (or, legacy epilogue for fp32 output / BN statistics: fp32 transpose tile -> global stores).
//   warp 10    epilogue TMA: tensor stores of the staged 64-channel chunks, residual chunks loaded ahead.
// PAIR form: clusters of two CTAs, tcgen05.mma.cta_group::2 (see the kernel's comment).
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"

namespace vfs {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // fp16 elements = 128 B = one swizzle row
constexpr int kTileABytes = kBlockM * kBlockK * 2;  // one plane of the activation tile (16 KB)
constexpr int kChunkBytes = kBlockM * 64 * 2 * 2;  // one 64-column output chunk, hi + lo planes (32 KB)
// warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue math (two per TMEM lane quarter), warp 10 epilogue TMA
// (tensor stores + residual loads; idle in the legacy epilogue)
constexpr int kNumThreads = 352;
constexpr int kMaxTaps = 9;
constexpr int kMaxViews = 4;

struct alignas(64) ConvKernelParams {
  CUtensorMap tmap_a[kMaxViews];  // activation views (one per stride-parity), dims {C, Wv, Hv, N, 2}
  CUtensorMap tmap_b;             // weights, dims {K, Cout, 2}
  CUtensorMap tmap_out;           // TMA epilogue: output view, dims {Cout, Wo, Ho, N, 2}, box {64, tw, th, tn, 2}
  CUtensorMap tmap_res;           // TMA epilogue: residual through the same view
  int num_m_tiles, num_n_tiles;
  int num_sched_tiles;  // tiles walked by the persistent loop: m x n tiles, or (m-tile pairs) x n tiles for CTA pairs
  int tiles_w, tiles_h;  // m_tile = (tn_i * tiles_h + th_i) * tiles_w + tw_i
  int tw, th, tn;        // tile extent in pixels, tw*th*tn == 128
  int Wo, Ho, N;         // output extent
  int Cout;
  int kchunks_per_tap;  // Cin / 64
  int num_taps;
  int tap_view[kMaxTaps];
  int tap_dh[kMaxTaps];
  int tap_dw[kMaxTaps];
  int tap_koff[kMaxTaps];  // K coordinate (elements) of this tap's first channel chunk in the weight operand
  // output / residual addressing: pixel (n, h, w) of this launch lives at
  //   ((n*out_H + h*out_sy + out_oy)*out_W + w*out_sx + out_ox) * Cout   (strided views are used by stride-2 dgrad)
  int out_H, out_W, out_sy, out_sx, out_oy, out_ox;
  const float* scale;
  const float* shift;
  const h16* res_hi;  // nullable
  const h16* res_lo;
  h16* out_hi;  // nullable
  h16* out_lo;
  float* out_f32;  // nullable, NHWC fp32
  int relu;
  double* stat_sum;    // nullable: per-channel sum / sum of squares of the epilogue output over all valid pixels
  double* stat_sqsum;  // (train-mode BatchNorm statistics; accumulated with fp64 atomics)
};

// Optional in-kernel timeline (debug instrument, vfs_debug_conv_trace): lane 0 of the TMA warp, the MMA warp and the
// first epilogue warp append (code, clock64) pairs to a per-CTA, per-role region of a caller-provided buffer.
__device__ long long* g_conv_trace = nullptr;
__device__ int g_conv_trace_cap = 0;

struct TraceCursor {
  long long* base;
  int cap, n;
  __device__ __forceinline__ void init(int role) {
    long long* t = g_conv_trace;
    cap = g_conv_trace_cap;
    n = 0;
    base = (t != nullptr && (threadIdx.x & 31) == 0) ? t + (static_cast<size_t>(blockIdx.x) * 3 + role) * cap * 2
                                                     : nullptr;
  }
  __device__ __forceinline__ void mark(int code) {
    if (base != nullptr && n < cap) {
      base[2 * n] = code;
      base[2 * n + 1] = clock64();
      ++n;
    }
  }
};

template <int BN, int STAGES, int NBUF, bool PAIR = false, bool STATS = false>
struct ConvSmem {
  static constexpr int kRowsB = PAIR ? BN / 2 : BN;                       // weight rows this CTA stages
  static constexpr int kTileBBytes = kRowsB * kBlockK * 2;                // one plane of the weight tile
  static constexpr int kStageBytes = 2 * kTileABytes + 2 * kTileBBytes;  // hi+lo of A and B
  // epilogue staging: legacy (NBUF == 0) one 128 x 64 fp32 transpose tile; TMA epilogue NBUF output chunks
  static constexpr int kStagingBytes = (NBUF == 0) ? kBlockM * 64 * 4 : NBUF * kChunkBytes;
  // 8 B x (2*STAGES + 4 pipeline + tmem ptr + 3 x 4 epilogue) <= 200, then (legacy epilogue) 2 x BN fp32 per-tile
  // partial sums of the BatchNorm statistics
  static constexpr int kBarrierBytes = 256 + ((NBUF == 0 || STATS) ? 2 * BN * 4 : 0);
  static constexpr int kTotal = STAGES * kStageBytes + kStagingBytes + kBarrierBytes + 1024;  // +1024 align slack
};

// NBUF == 0: legacy epilogue (fp32 transpose tile, direct global stores; fp32 output and BN statistics supported).
// NBUF  > 0: TMA epilogue (split output only): results staged as hi/lo planes in the 128B-swizzled box layout and
//            written with one cp.async.bulk.tensor store per 64-column chunk; RES adds a residual that is brought in by
//            TMA into the same staging buffer NBUF-1 chunks ahead.
// PAIR: the kernel runs as clusters of two CTAs on one TPC and issues tcgen05.mma.cta_group::2 (M = 256): CTA rank r
//       owns output pixel tile 2*pair + r (its 128 accumulator rows live in its own TMEM) and stages only HALF of the
//       weight tile (rows r*BN/2 ...), the tensor cores of the two SMs exchange the halves.  Per SM and MMA this
//       halves the shared-memory reads of B and the TMA fill traffic of B -- the 1-CTA kernel is bound by exactly that
//       (profiles/r01_conv_trace_v7.log: 1093 cycles per K-chunk against an MMA floor of 768).  Only the leader CTA's
//       warp 1 issues MMAs; its commits are multicast to the barriers of both CTAs; the peer's TMA loads signal the
//       leader's full barriers; the peer's epilogue releases accumulator stages on the leader's barrier.
// Sum over the warp's 32 lanes (accumulator rows) of 8 per-lane values (columns): recursive halving over lane bits
// 4, 3, 2 (7 shuffles) + a 2-step butterfly; every lane returns the total of column lane >> 2.
__device__ __forceinline__ float warp_colsum8(float (&v)[8], int lane) {
  {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? v[i] : v[i + 4];
      const float keep = up ? v[i + 4] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? v[i] : v[i + 2];
      const float keep = up ? v[i + 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = (lane & 4) != 0;
    const float send = up ? v[0] : v[1];
    const float keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];
}

// STATS (TMA epilogue only): per-channel sum / sum of squares of the epilogue output over the valid pixels of every
// tile (train-mode BatchNorm statistics) -- column sums by warp shuffles, combined across the eight math warps in shared
// memory, one fp64 atomic per channel and tile.  The raw conv output leaves as a split tensor through the TMA store.
template <int BN, int STAGES, int NBUF, bool RES, bool PAIR = false, bool STATS = false>
__global__ void __launch_bounds__(kNumThreads, 1) conv_tc_kernel(const __grid_constant__ ConvKernelParams p) {
  using S = ConvSmem<BN, STAGES, NBUF, PAIR, STATS>;
  static_assert(!STATS || (NBUF > 0 && !RES), "statistics ride on the residual-free TMA epilogue");
  static_assert(!RES || NBUF >= 2, "the residual prefetch needs at least two staging buffers");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = smem_base + STAGES * S::kStageBytes;
  const uint32_t bar_base = stg_base + S::kStagingBytes;
  // barrier layout (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then tmem ptr (4 B)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 4);
  auto rfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 5 + b); };   // residual chunk landed
  auto staged_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 9 + b); };  // output chunk written to smem
  auto free_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 13 + b); };   // store finished reading the buffer

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 2 * BN;  // two accumulator stages (power of two: 128, 256 or 512)
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  // persistent schedule: this CTA (pair) walks tiles tile_first, tile_first + tile_step, ...
  const int tile_first = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  auto m_tile_of = [&](int tile) {
    const int m = tile / p.num_n_tiles;
    return PAIR ? 2 * m + static_cast<int>(cta_rank) : m;
  };

  if (warp == 0 && lane == 0) {
    for (int v = 0; v < kMaxViews; ++v) tma_prefetch_desc(&p.tmap_a[v]);
    tma_prefetch_desc(&p.tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), PAIR ? 16 : 8);  // epilogue warps of both CTAs release the leader's accumulator
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(rfull_bar(b), 1);
      mbar_init(staged_bar(b), 8);
      mbar_init(free_bar(b), 1);
    }
    if (NBUF > 0) {
      tma_prefetch_desc(&p.tmap_out);
      if (RES) tma_prefetch_desc(&p.tmap_res);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) {
      tmem_alloc_2sm(tmem_ptr_addr, kTmemCols);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_ptr_addr, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // the peer's barriers must be initialised before anything remote touches them
  else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
  const int num_tiles = p.num_sched_tiles;
  const int num_kchunks = p.num_taps * p.kchunks_per_tap;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) may overlap
  // the tail of the previous kernel in the stream; from the wait on, its results are read and its inputs overwritten.
  // NOTHING that reads global memory may move above the wait -- not even the weight tiles: callers do produce weights
  // with the immediately preceding launch (pack -> conv in training, features -> conv in the dense affinity paths),
  // and a kernel that never triggers early only guarantees visibility of its stores at griddepcontrol.wait.  (Round 1
  // tried issuing the first weight tiles before the wait: -1 % on the bench, one flaky parity failure.)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ======================= TMA producer =======================
    int stage = 0;
    uint32_t phase = 0;
    TraceCursor tr;
    tr.init(0);
    tr.mark(0);
    for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
      const int m_tile = m_tile_of(tile);
      const int n_tile = tile % p.num_n_tiles;
      const int tw_i = m_tile % p.tiles_w;
      const int t2 = m_tile / p.tiles_w;
      const int th_i = t2 % p.tiles_h;
      const int tn_i = t2 / p.tiles_h;  // >= tiles_n for the phantom tile of an odd pair: every load is zero fill
      const int w0 = tw_i * p.tw, h0 = th_i * p.th, n0 = tn_i * p.tn;
      for (int kc = 0; kc < num_kchunks; ++kc) {
        mbar_wait(empty_bar(stage), phase ^ 1u, 100 + stage);
        tr.mark(1);
        if (lane == 0) {
          const int tap = kc / p.kchunks_per_tap;
          const int c0 = (kc - tap * p.kchunks_per_tap) * kBlockK;
          const uint32_t sa = smem_base + stage * S::kStageBytes;
          const uint32_t sb = sa + 2 * kTileABytes;
          if (PAIR) {
            // both CTAs' bytes are counted on the leader's barrier (the MMA of the pair waits there)
            if (cta_rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * S::kStageBytes);
            tma_load_5d_2sm(sa, &p.tmap_a[p.tap_view[tap]], full_bar(stage), c0, w0 + p.tap_dw[tap],
                            h0 + p.tap_dh[tap], n0, 0);
            tma_load_3d_2sm(sb, &p.tmap_b, full_bar(stage), p.tap_koff[tap] + c0,
                            n_tile * BN + static_cast<int>(cta_rank) * S::kRowsB, 0);
          } else {
            mbar_arrive_expect_tx(full_bar(stage), S::kStageBytes);
            tma_load_5d(sa, &p.tmap_a[p.tap_view[tap]], full_bar(stage), c0, w0 + p.tap_dw[tap], h0 + p.tap_dh[tap],
                        n0, 0);
            tma_load_3d(sb, &p.tmap_b, full_bar(stage), p.tap_koff[tap] + c0, n_tile * BN, 0);
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    constexpr uint32_t idesc = umma_idesc_f16_f32(PAIR ? 2 * kBlockM : kBlockM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    TraceCursor tr;
    tr.init(1);
    tr.mark(0);
    // PAIR: the peer CTA's warp 1 only takes part in the TMEM allocation
    for (int tile = tile_first; tile < num_tiles && cta_rank == 0; tile += tile_step) {
      mbar_wait(tempty_bar(as), aphase ^ 1u, 200 + as);
      tr.mark(8);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kc = 0; kc < num_kchunks; ++kc) {
        mbar_wait(full_bar(stage), phase, 300 + stage);
        tr.mark(2);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_base + stage * S::kStageBytes;
          const uint32_t a_hi = sa, a_lo = sa + kTileABytes;
          const uint32_t b_hi = sa + 2 * kTileABytes, b_lo = b_hi + S::kTileBBytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint32_t koff = k * 32;  // 16 halves = 32 B inside the 128 B swizzle row
            const uint64_t da_hi = umma_desc_sw128_kmajor(a_hi + koff);
            const uint64_t da_lo = umma_desc_sw128_kmajor(a_lo + koff);
            const uint64_t db_hi = umma_desc_sw128_kmajor(b_hi + koff);
            const uint64_t db_lo = umma_desc_sw128_kmajor(b_lo + koff);
            if (PAIR) {
              umma_f16_2sm(d_tmem, da_lo, db_hi, idesc, (kc | k) != 0 ? 1u : 0u);
              umma_f16_2sm(d_tmem, da_hi, db_lo, idesc, 1u);
              umma_f16_2sm(d_tmem, da_hi, db_hi, idesc, 1u);
            } else {
              umma_f16(d_tmem, da_lo, db_hi, idesc, (kc | k) != 0 ? 1u : 0u);  // small terms first
              umma_f16(d_tmem, da_hi, db_lo, idesc, 1u);
              umma_f16(d_tmem, da_hi, db_hi, idesc, 1u);
            }
          }
          if (PAIR) {
            umma_commit_2sm(empty_bar(stage));  // frees the slot in BOTH CTAs when these MMAs retire
            if (kc == num_kchunks - 1) umma_commit_2sm(tfull_bar(as));
          } else {
            umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs retire
            if (kc == num_kchunks - 1) umma_commit(tfull_bar(as));
          }
        }
        if (kc == num_kchunks - 1) tr.mark(3);
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1u;
      }
    }
  } else if (warp == 10) {
    // ======================= epilogue TMA warp =======================
    // Per 64-column chunk g (staging buffer g % NBUF): wait until the eight math warps have staged it, issue the
    // tensor store, then recycle the buffer once the store has finished reading it -- for RES by loading the residual
    // of chunk g + NBUF into it, otherwise by signalling it free.
    if constexpr (NBUF > 0) {
      constexpr int kCPT = BN / 64;  // chunks per tile
      auto chunk_coords = [&](int g, int& cc, int& w0, int& h0, int& n0) -> bool {
        const int tile = tile_first + (g / kCPT) * tile_step;
        if (tile >= num_tiles) return false;
        const int m_tile = m_tile_of(tile);
        const int n_tile = tile % p.num_n_tiles;
        const int tw_i = m_tile % p.tiles_w;
        const int t2 = m_tile / p.tiles_w;
        cc = n_tile * BN + (g % kCPT) * 64;
        w0 = tw_i * p.tw;
        h0 = (t2 % p.tiles_h) * p.th;
        n0 = (t2 / p.tiles_h) * p.tn;
        return true;
      };
      auto issue_res = [&](int g) {
        int cc, w0, h0, n0;
        if (chunk_coords(g, cc, w0, h0, n0)) {
          const int b = g % NBUF;
          mbar_arrive_expect_tx(rfull_bar(b), kChunkBytes);
          tma_load_5d(stg_base + b * kChunkBytes, &p.tmap_res, rfull_bar(b), cc, w0, h0, n0, 0);
        }
      };
      if (lane == 0) {
        if (RES) {
          for (int g0 = 0; g0 < NBUF; ++g0) issue_res(g0);
        }
        int cc, w0, h0, n0;
        for (int g = 0; chunk_coords(g, cc, w0, h0, n0); ++g) {
          const int b = g % NBUF;
          mbar_wait(staged_bar(b), (g / NBUF) & 1, 600 + b);
          tma_store_5d(&p.tmap_out, stg_base + b * kChunkBytes, cc, w0, h0, n0, 0);
          bulk_commit_group();
          if (RES) {
            // Wait until THIS store has drained its buffer (a few hundred ns; this warp has nothing else to do until
            // the next chunk is staged) and refill the same buffer with the residual of chunk g + NBUF: the residual
            // ring then runs NBUF chunks ahead of the math warps instead of one (the L2 -> smem latency of a residual
            // chunk is longer than one chunk's epilogue math).
            bulk_wait_group_read<0>();
            issue_res(g + NBUF);
          } else {
            bulk_wait_group_read<NBUF - 1>();  // store g - (NBUF - 1) has released its buffer
            if (g >= NBUF - 1) mbar_arrive(free_bar((g - (NBUF - 1)) % NBUF));
          }
        }
        bulk_wait_group_all();
      }
    }
  } else if constexpr (NBUF > 0) {
    // ======================= TMA epilogue, math warps 2..9 =======================
    // thread = accumulator row (pixel) x 32 of the chunk's 64 columns: TMEM -> registers -> scale/shift (+residual read
    // from the staging buffer) -> ReLU -> split -> hi/lo planes of the staging buffer in the swizzled box layout.
    // No per-pixel address arithmetic, no bounds checks (TMA clips), no block-wide barrier: each thread only touches
    // its own 8 x 16 bytes of the buffer, hand-over to / from the TMA warp goes through mbarriers.
    const int ew = warp - 2;
    const int q = warp & 3;   // TMEM lane quarter this warp may access (hardware rule: warp id % 4)
    const int ch = ew >> 2;   // which 32-column half of the 64-column chunk
    const int row = q * 32 + lane;
    constexpr int kCPT = BN / 64;  // chunks per tile
    uint32_t off[4];               // byte offsets of this thread's four 16-byte pieces inside one plane
#pragma unroll
    for (int j = 0; j < 4; ++j) off[j] = row * 128 + (((ch * 4 + j) ^ (row & 7)) << 4);
    const float lo_clamp = p.relu ? 0.0f : -INFINITY;
    float amax = 0.0f;
    TraceCursor tr;
    tr.init(2);
    if (ew != 0) tr.base = nullptr;
    tr.mark(0);
    int as = 0;
    uint32_t aphase = 0;
    int g = 0;
    const int et = threadIdx.x - 64;  // 0..255 over the math warps
    float* s_stat = reinterpret_cast<float*>(smem_raw + (bar_base + 256 - smem_u32(smem_raw)));
    if constexpr (STATS) {
      for (int i = et; i < 2 * BN; i += 256) s_stat[i] = 0.0f;
      asm volatile("bar.sync 2, 256;" ::: "memory");
    }
    for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
      const int n_tile = tile % p.num_n_tiles;
      bool row_valid = true;
      if constexpr (STATS) {   // rows beyond the image (ragged tiles, the phantom tile of an odd pair) do not count
        const int m_tile = m_tile_of(tile);
        const int tw_i = m_tile % p.tiles_w;
        const int t2 = m_tile / p.tiles_w;
        const int r2 = row / p.tw;
        const int w = tw_i * p.tw + row % p.tw;
        const int hh = (t2 % p.tiles_h) * p.th + r2 % p.th;
        const int n = (t2 / p.tiles_h) * p.tn + r2 / p.th;
        row_valid = (w < p.Wo) && (hh < p.Ho) && (n < p.N);
      }
      // pull this tile's scale / shift lines into L1 while the accumulator is still being produced
      if (lane < BN / 16) {
        const float* base = (lane < BN / 32) ? p.scale : p.shift;
        const int line = (lane < BN / 32) ? lane : lane - BN / 32;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(base + n_tile * BN + line * 32));
      }
      tr.mark(4);
      mbar_wait(tfull_bar(as), aphase, 400 + as);
      tr.mark(5);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = 0; c < kCPT; ++c, ++g) {
        const int b = g % NBUF;
        const uint32_t buf = stg_base + b * kChunkBytes;
        uint32_t acc[32];
        tmem_ld_32x32b_x32(t_row + c * 64 + ch * 32, acc);
        tmem_ld_wait();
        if (c == kCPT - 1) {  // last TMEM read of this tile: hand the accumulator stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_cluster(tempty_bar(as), 0);  // the pair's MMA issuer lives in the leader CTA
            else mbar_arrive(tempty_bar(as));
          }
        }
        tr.mark(9);
        if (RES) mbar_wait(rfull_bar(b), (g / NBUF) & 1, 500 + b);
        tr.mark(10);
        uint4 oh[4], ol[4];
        uint4 rh[4], rl[4];
        if (RES) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(rh[j].x), "=r"(rh[j].y), "=r"(rh[j].z), "=r"(rh[j].w) : "r"(buf + off[j]));
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(rl[j].x), "=r"(rl[j].y), "=r"(rl[j].z), "=r"(rl[j].w)
                         : "r"(buf + kChunkBytes / 2 + off[j]));
          }
        }
        const float* sc_ptr = p.scale + n_tile * BN + c * 64 + ch * 32;
        const float* sh_ptr = p.shift + n_tile * BN + c * 64 + ch * 32;
        float my_sum = 0.0f, my_sq = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 sc0 = __ldg(reinterpret_cast<const float4*>(sc_ptr + 8 * j));
          const float4 sc1 = __ldg(reinterpret_cast<const float4*>(sc_ptr + 8 * j + 4));
          const float4 sh0 = __ldg(reinterpret_cast<const float4*>(sh_ptr + 8 * j));
          const float4 sh1 = __ldg(reinterpret_cast<const float4*>(sh_ptr + 8 * j + 4));
          float2 y[4];
          y[0] = fma2(make_float2(__uint_as_float(acc[8 * j]), __uint_as_float(acc[8 * j + 1])),
                      make_float2(sc0.x, sc0.y), make_float2(sh0.x, sh0.y));
          y[1] = fma2(make_float2(__uint_as_float(acc[8 * j + 2]), __uint_as_float(acc[8 * j + 3])),
                      make_float2(sc0.z, sc0.w), make_float2(sh0.z, sh0.w));
          y[2] = fma2(make_float2(__uint_as_float(acc[8 * j + 4]), __uint_as_float(acc[8 * j + 5])),
                      make_float2(sc1.x, sc1.y), make_float2(sh1.x, sh1.y));
          y[3] = fma2(make_float2(__uint_as_float(acc[8 * j + 6]), __uint_as_float(acc[8 * j + 7])),
                      make_float2(sc1.z, sc1.w), make_float2(sh1.z, sh1.w));
          if (RES) {
            y[0] = add2(add2(y[0], h2_to_float2(rl[j].x)), h2_to_float2(rh[j].x));
            y[1] = add2(add2(y[1], h2_to_float2(rl[j].y)), h2_to_float2(rh[j].y));
            y[2] = add2(add2(y[2], h2_to_float2(rl[j].z)), h2_to_float2(rh[j].z));
            y[3] = add2(add2(y[3], h2_to_float2(rl[j].w)), h2_to_float2(rh[j].w));
          }
          uint32_t h2[4], l2[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            y[e].x = fmaxf(y[e].x, lo_clamp);
            y[e].y = fmaxf(y[e].y, lo_clamp);
            amax = fmaxf(amax, fmaxf(fabsf(y[e].x), fabsf(y[e].y)));
            const __half2 h = __floats2half2_rn(y[e].x, y[e].y);
            const float2 d = sub2(y[e], __half22float2(h));
            const __half2 l = __floats2half2_rn(d.x, d.y);
            h2[e] = *reinterpret_cast<const uint32_t*>(&h);
            l2[e] = *reinterpret_cast<const uint32_t*>(&l);
          }
          oh[j] = make_uint4(h2[0], h2[1], h2[2], h2[3]);
          ol[j] = make_uint4(l2[0], l2[1], l2[2], l2[3]);
          if constexpr (STATS) {
            float v[8], q[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              v[2 * e] = row_valid ? y[e].x : 0.0f;
              v[2 * e + 1] = row_valid ? y[e].y : 0.0f;
              q[2 * e] = v[2 * e] * v[2 * e];
              q[2 * e + 1] = v[2 * e + 1] * v[2 * e + 1];
            }
            const float cs = warp_colsum8(v, lane), cq = warp_colsum8(q, lane);
            if ((lane & 3) == j) {   // lane keeps column 8 (lane & 3) + (lane >> 2) of this warp's 32 columns
              my_sum = cs;
              my_sq = cq;
            }
          }
        }
        if constexpr (STATS) {
          const int col = c * 64 + ch * 32 + 8 * (lane & 3) + (lane >> 2);
          atomicAdd(s_stat + col, my_sum);
          atomicAdd(s_stat + BN + col, my_sq);
        }
        tr.mark(11);
        // RES: the landed residual implies the buffer was free; otherwise wait for the store NBUF chunks ago
        if (!RES) mbar_wait(free_bar(b), ((g / NBUF) & 1) ^ 1u, 700 + b);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + off[j]), "r"(oh[j].x), "r"(oh[j].y),
                       "r"(oh[j].z), "r"(oh[j].w) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + kChunkBytes / 2 + off[j]), "r"(ol[j].x),
                       "r"(ol[j].y), "r"(ol[j].z), "r"(ol[j].w) : "memory");
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(staged_bar(b));
        tr.mark(7);
      }
      if constexpr (STATS) {
        asm volatile("bar.sync 2, 256;" ::: "memory");  // every math warp's partial sums of this tile are in
        if (et < BN) {
          atomicAdd(p.stat_sum + n_tile * BN + et, static_cast<double>(s_stat[et]));
          atomicAdd(p.stat_sqsum + n_tile * BN + et, static_cast<double>(s_stat[BN + et]));
          s_stat[et] = 0.0f;
          s_stat[BN + et] = 0.0f;
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");  // zeroed before the next tile accumulates
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1u;
      }
    }
    if (!(amax <= 65504.0f)) atomicAdd(&g_split_overflow, 1u);
  } else if (warp < 10) {
    // ======================= legacy epilogue (warps 2..9) =======================
    // Phase A (thread = accumulator row, warp = 32 rows x 32 of the chunk's 64 columns): TMEM -> fp32 staging tile
    // in shared memory (16-byte chunks XOR-swizzled so both phases are bank-conflict free).
    // Phase B (8 threads per pixel row, 8 channels each): staging -> scale/shift (+residual) -> ReLU -> split ->
    // global, every warp access covering whole 128-byte lines.  Residual pieces are prefetched before phase A.
    const int ew = warp - 2;   // 0..7
    const int q = warp & 3;    // TMEM lane quarter this warp may access (hardware rule: warp id % 4)
    const int ch = ew >> 2;    // which 32-column half of the 64-column chunk this warp stages
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;         // 0..255 over the epilogue threads
    const int piece = et & 7, rr = et >> 3;  // phase B: 8-channel piece of the chunk, first of 4 rows (stride 32)
    // per-tile partial sums of the BN statistics: [sum[BN], sum of squares[BN]] fp32 in shared memory, flushed with
    // ONE fp64 atomic per channel and tile (the first version sent 2 x 8 fp64 atomics per warp and chunk to global)
    float* s_stat = reinterpret_cast<float*>(smem_raw + (bar_base + 256 - smem_u32(smem_raw)));
    if (p.stat_sum != nullptr) {
      for (int i = et; i < 2 * BN; i += 256) s_stat[i] = 0.0f;
      asm volatile("bar.sync 2, 256;" ::: "memory");
    }
    auto phys_chunk = [](int r, int c) { return (c & 8) | ((c ^ (c >> 3) ^ r) & 7); };
    int as = 0;
    uint32_t aphase = 0;
    TraceCursor tr;
    tr.init(2);
    if (ew != 0) tr.base = nullptr;
    tr.mark(0);
    static_assert(!PAIR || NBUF > 0, "CTA pairs use the TMA epilogue");
    for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
      const int m_tile = tile / p.num_n_tiles;
      const int n_tile = tile - m_tile * p.num_n_tiles;
      const int tw_i = m_tile % p.tiles_w;
      const int t2 = m_tile / p.tiles_w;
      const int th_i = t2 % p.tiles_h;
      const int tn_i = t2 / p.tiles_h;
      const int w0 = tw_i * p.tw, h0 = th_i * p.th, n0 = tn_i * p.tn;
      // element offsets of this thread's 4 output pixels (index math once per tile, not per chunk)
      size_t obase[4];
      bool ok[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = rr + 32 * i;
        const int dw = r % p.tw;
        const int r2 = r / p.tw;
        const int w = w0 + dw, hh = h0 + r2 % p.th, n = n0 + r2 / p.th;
        ok[i] = (w < p.Wo) && (hh < p.Ho) && (n < p.N);
        obase[i] = ((static_cast<size_t>(n) * p.out_H + hh * p.out_sy + p.out_oy) * p.out_W + w * p.out_sx + p.out_ox) *
                       p.Cout + n_tile * BN + piece * 8;
      }

      tr.mark(4);
      mbar_wait(tfull_bar(as), aphase, 400 + as);
      tr.mark(5);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 64) {
        const int cg = n_tile * BN + c0 + piece * 8;
        uint4 pre_h[4], pre_l[4];
        if (p.res_hi != nullptr) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            pre_h[i] = make_uint4(0, 0, 0, 0);
            pre_l[i] = make_uint4(0, 0, 0, 0);
            if (ok[i]) {
              pre_h[i] = *reinterpret_cast<const uint4*>(p.res_hi + obase[i] + c0);
              pre_l[i] = *reinterpret_cast<const uint4*>(p.res_lo + obase[i] + c0);
            }
          }
        }
        const float4 sc0 = __ldg(reinterpret_cast<const float4*>(p.scale + cg));
        const float4 sc1 = __ldg(reinterpret_cast<const float4*>(p.scale + cg + 4));
        const float4 sh0 = __ldg(reinterpret_cast<const float4*>(p.shift + cg));
        const float4 sh1 = __ldg(reinterpret_cast<const float4*>(p.shift + cg + 4));
        asm volatile("bar.sync 1, 256;" ::: "memory");  // staging buffer free
        {
          uint32_t acc[32];
          tmem_ld_32x32b_x32(t_row + c0 + ch * 32, acc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t addr = stg_base + row * 256 + phys_chunk(row, ch * 8 + j) * 16;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(acc[4 * j]), "r"(acc[4 * j + 1]),
                         "r"(acc[4 * j + 2]), "r"(acc[4 * j + 3])
                         : "memory");
          }
        }
        if (c0 + 64 >= BN) {  // last TMEM read of this tile: hand the accumulator stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(as));
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // staging buffer full
        tr.mark(6);
        float st_s[8], st_q[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) st_s[e] = st_q[e] = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (ok[i]) {
            const int r = rr + 32 * i;
            const size_t o = obase[i] + c0;
            float y[8];
            const uint32_t a0 = stg_base + r * 256 + phys_chunk(r, 2 * piece) * 16;
            const uint32_t a1 = stg_base + r * 256 + phys_chunk(r, 2 * piece + 1) * 16;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(y[0]), "=f"(y[1]), "=f"(y[2]), "=f"(y[3]) : "r"(a0));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(y[4]), "=f"(y[5]), "=f"(y[6]), "=f"(y[7]) : "r"(a1));
            y[0] = fmaf(y[0], sc0.x, sh0.x); y[1] = fmaf(y[1], sc0.y, sh0.y);
            y[2] = fmaf(y[2], sc0.z, sh0.z); y[3] = fmaf(y[3], sc0.w, sh0.w);
            y[4] = fmaf(y[4], sc1.x, sh1.x); y[5] = fmaf(y[5], sc1.y, sh1.y);
            y[6] = fmaf(y[6], sc1.z, sh1.z); y[7] = fmaf(y[7], sc1.w, sh1.w);
            if (p.res_hi != nullptr) {
              const uint4 rh = pre_h[i], rl = pre_l[i];
              const float2 h0f = h2_to_float2(rh.x), l0f = h2_to_float2(rl.x);
              const float2 h1f = h2_to_float2(rh.y), l1f = h2_to_float2(rl.y);
              const float2 h2f = h2_to_float2(rh.z), l2f = h2_to_float2(rl.z);
              const float2 h3f = h2_to_float2(rh.w), l3f = h2_to_float2(rl.w);
              y[0] += h0f.x + l0f.x; y[1] += h0f.y + l0f.y;
              y[2] += h1f.x + l1f.x; y[3] += h1f.y + l1f.y;
              y[4] += h2f.x + l2f.x; y[5] += h2f.y + l2f.y;
              y[6] += h3f.x + l3f.x; y[7] += h3f.y + l3f.y;
            }
            if (p.relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) y[e] = fmaxf(y[e], 0.0f);
            }
            if (p.stat_sum != nullptr) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                st_s[e] += y[e];
                st_q[e] = fmaf(y[e], y[e], st_q[e]);
              }
            }
            if (p.out_f32 != nullptr) {
              float4* of = reinterpret_cast<float4*>(p.out_f32 + o);
              of[0] = make_float4(y[0], y[1], y[2], y[3]);
              of[1] = make_float4(y[4], y[5], y[6], y[7]);
            }
            if (p.out_hi != nullptr) {
              uint4 oh, ol;
              split16x2(y[0], y[1], oh.x, ol.x);
              split16x2(y[2], y[3], oh.y, ol.y);
              split16x2(y[4], y[5], oh.z, ol.z);
              split16x2(y[6], y[7], oh.w, ol.w);
              note_overflow8(y);
              *reinterpret_cast<uint4*>(p.out_hi + o) = oh;
              *reinterpret_cast<uint4*>(p.out_lo + o) = ol;
            }
          }
        }
        tr.mark(7);
        if (p.stat_sum != nullptr) {
          // lanes {l, l+8, l+16, l+24} hold the same 8 channels for different rows: fold them, then one fp64
          // atomic per channel and warp
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            st_s[e] += __shfl_xor_sync(0xffffffffu, st_s[e], 8);
            st_q[e] += __shfl_xor_sync(0xffffffffu, st_q[e], 8);
            st_s[e] += __shfl_xor_sync(0xffffffffu, st_s[e], 16);
            st_q[e] += __shfl_xor_sync(0xffffffffu, st_q[e], 16);
          }
          if (lane < 8) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              atomicAdd(s_stat + c0 + piece * 8 + e, st_s[e]);
              atomicAdd(s_stat + BN + c0 + piece * 8 + e, st_q[e]);
            }
          }
        }
      }
      if (p.stat_sum != nullptr) {
        asm volatile("bar.sync 2, 256;" ::: "memory");  // every warp's partial sums of this tile are in
        if (et < BN) {
          atomicAdd(p.stat_sum + n_tile * BN + et, static_cast<double>(s_stat[et]));
          atomicAdd(p.stat_sqsum + n_tile * BN + et, static_cast<double>(s_stat[BN + et]));
          s_stat[et] = 0.0f;
          s_stat[BN + et] = 0.0f;
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");  // zeroed before the next tile accumulates
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1u;
      }
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();  // both CTAs are done with TMEM and with each other's barriers
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2sm(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
namespace {

inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// choose (tw, th, tn), tw*th*tn = 128, powers of two, minimising the number of tiles
void choose_tile(int Wo, int Ho, int N, int* tw, int* th, int* tn) {
  long best = -1;
  for (int a = 1; a <= 128; a <<= 1) {      // tw
    for (int b = 1; a * b <= 128; b <<= 1) {  // th
      const int c = 128 / (a * b);            // tn
      const long tiles = static_cast<long>((Wo + a - 1) / a) * ((Ho + b - 1) / b) * ((N + c - 1) / c);
      // prefer fewer tiles; tie-break towards wider rows (longer contiguous runs)
      if (best < 0 || tiles < best) {
        best = tiles;
        *tw = a;
        *th = b;
        *tn = c;
      }
    }
  }
}

// CTA-pair policy: mode 0 = never, 1 = whenever the shape allows, 2 = when the launch has >= min_tiles pair tiles.
// Initialised from VFS_CONV_PAIR / VFS_CONV_PAIR_MIN_TILES, changed at run time by vfs_conv_set_pair_policy (tests,
// tuning).
int g_pair_mode = -1, g_pair_min = 48;
void conv_pair_policy(int* mode, int* min_tiles) {
  if (g_pair_mode < 0) {
    const char* e = getenv("VFS_CONV_PAIR");
    g_pair_mode = e ? atoi(e) : 2;
    const char* m = getenv("VFS_CONV_PAIR_MIN_TILES");
    if (m) g_pair_min = atoi(m);
  }
  *mode = g_pair_mode;
  *min_tiles = g_pair_min;
}

// Persistent launch: one CTA per SM (or one CTA pair per TPC) walking the tile list; p.num_m_tiles / num_n_tiles must
// be final.
template <int BN, int STAGES, int NBUF, bool RES, bool PAIR = false, bool STATS = false>
int launch(ConvKernelParams p, cudaStream_t stream) {
  using S = ConvSmem<BN, STAGES, NBUF, PAIR, STATS>;
  static_assert(S::kTotal <= 232448, "shared memory budget (227 KB) exceeded");
  static_assert(2 * BN <= 512, "two accumulator stages must fit the 512 TMEM columns");
  static bool configured = false;
  if (!configured) {
    VFS_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES, NBUF, RES, PAIR, STATS>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured = true;
  }
  const int sms = device_sm_count();
  int grid;
  if (PAIR) {
    p.num_sched_tiles = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
    grid = 2 * p.num_sched_tiles < (sms & ~1) ? 2 * p.num_sched_tiles : (sms & ~1);
  } else {
    p.num_sched_tiles = p.num_m_tiles * p.num_n_tiles;
    grid = p.num_sched_tiles < sms ? p.num_sched_tiles : sms;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  static int pdl = -1;  // VFS_CONV_PDL=0 disables programmatic dependent launch (experiments with interleaved streams)
  if (pdl < 0) {
    const char* e = getenv("VFS_CONV_PDL");
    pdl = (e && atoi(e) == 0) ? 0 : 1;
  }
  attr[0].val.programmaticStreamSerializationAllowed = pdl;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = PAIR ? 2 : 1;
  VFS_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, STAGES, NBUF, RES, PAIR, STATS>, p));
  return VFS_OK;
}

}  // namespace

struct TapSpec {
  int view, dh, dw, koff;
};

// One launch of the implicit-GEMM kernel: A = pixels of a split NHWC tensor (optionally through stride-parity
// views), B = [2][Nout][Ktot] split weights, output pixels written through a (possibly strided) view.
struct ConvSpec {
  const void* a_split;
  size_t a_plane;
  int N, H, W, C;
  int view_stride;
  bool flat;
  int num_taps;
  TapSpec taps[kMaxTaps];
  const void* b_split;
  int Nout;
  uint64_t Ktot;
  int Ho, Wo;
  int out_H, out_W, out_sy, out_sx, out_oy, out_ox;
  const float* scale;
  const float* shift;
  const void* res_split;
  void* out_split;
  size_t out_plane;
  float* out_f32;
  double* stats;
  int relu;
};

static int run_spec(const ConvSpec& c, cudaStream_t stream) {
  ConvKernelParams p;
  memset(&p, 0, sizeof(p));
  p.Cout = c.Nout;
  p.kchunks_per_tap = c.C / 64;
  p.num_taps = c.num_taps;
  p.scale = c.scale;
  p.shift = c.shift;
  p.relu = c.relu;
  p.out_f32 = c.out_f32;
  if (c.stats) {
    p.stat_sum = c.stats;
    p.stat_sqsum = c.stats + c.Nout;
  }
  if (c.out_split) {
    p.out_hi = reinterpret_cast<h16*>(c.out_split);
    p.out_lo = p.out_hi + c.out_plane;
  }
  if (c.res_split) {
    p.res_hi = reinterpret_cast<const h16*>(c.res_split);
    p.res_lo = p.res_hi + c.out_plane;
  }
  const int s = c.view_stride;
  if (c.flat) {
    p.tw = 128; p.th = 1; p.tn = 1;
    p.Wo = c.N * c.H * c.W; p.Ho = 1; p.N = 1;
    p.out_H = 1; p.out_W = p.Wo; p.out_sy = 1; p.out_sx = 1; p.out_oy = 0; p.out_ox = 0;
  } else {
    choose_tile(c.Wo, c.Ho, c.N, &p.tw, &p.th, &p.tn);
    p.Wo = c.Wo; p.Ho = c.Ho; p.N = c.N;
    p.out_H = c.out_H; p.out_W = c.out_W; p.out_sy = c.out_sy; p.out_sx = c.out_sx;
    p.out_oy = c.out_oy; p.out_ox = c.out_ox;
  }
  p.tiles_w = (p.Wo + p.tw - 1) / p.tw;
  p.tiles_h = (p.Ho + p.th - 1) / p.th;
  const int tiles_n = (p.N + p.tn - 1) / p.tn;
  p.num_m_tiles = p.tiles_w * p.tiles_h * tiles_n;

  const uint32_t box_a[5] = {64u, static_cast<uint32_t>(p.tw), static_cast<uint32_t>(p.th),
                             static_cast<uint32_t>(p.tn), 2u};
  const char* in_base = reinterpret_cast<const char*>(c.a_split);
  bool view_used[kMaxViews] = {false, false, false, false};
  for (int t = 0; t < c.num_taps; ++t) {
    p.tap_view[t] = c.taps[t].view;
    p.tap_dh[t] = c.taps[t].dh;
    p.tap_dw[t] = c.taps[t].dw;
    p.tap_koff[t] = c.taps[t].koff;
    view_used[c.taps[t].view] = true;
  }
  int first_valid = -1;
  for (int v = 0; v < kMaxViews; ++v) {
    const int ph = v / 2, pw = v % 2;
    int rc = VFS_OK;
    if (c.flat) {
      const uint64_t dims[5] = {static_cast<uint64_t>(c.C), static_cast<uint64_t>(c.N) * c.H * c.W, 1, 1, 2};
      const uint64_t strides[4] = {static_cast<uint64_t>(c.C) * 2, c.a_plane * 2, c.a_plane * 2, c.a_plane * 2};
      rc = make_tmap_16b_sw128(&p.tmap_a[v], in_base, 5, dims, strides, box_a);
      if (first_valid < 0) first_valid = v;
    } else if (view_used[v] && ph < c.H && pw < c.W && (s > 1 || v == 0)) {
      const int Hv = (c.H - ph + s - 1) / s, Wv = (c.W - pw + s - 1) / s;
      const uint64_t dims[5] = {static_cast<uint64_t>(c.C), static_cast<uint64_t>(Wv), static_cast<uint64_t>(Hv),
                                static_cast<uint64_t>(c.N), 2};
      const uint64_t strides[4] = {static_cast<uint64_t>(s) * c.C * 2, static_cast<uint64_t>(s) * c.W * c.C * 2,
                                   static_cast<uint64_t>(c.H) * c.W * c.C * 2, c.a_plane * 2};
      rc = make_tmap_16b_sw128(&p.tmap_a[v], in_base + (static_cast<size_t>(ph) * c.W + pw) * c.C * 2, 5, dims,
                                strides, box_a);
      if (first_valid < 0) first_valid = v;
    }
    if (rc != VFS_OK) return rc;
  }
  VFS_REQUIRE(first_valid >= 0, VFS_EINVAL, "conv: no usable input view");
  for (int v = 0; v < kMaxViews; ++v) {
    const bool built = c.flat || (view_used[v] && (v / 2) < c.H && (v % 2) < c.W && (s > 1 || v == 0));
    if (!built) p.tmap_a[v] = p.tmap_a[first_valid];  // never used by a tap; keeps the prefetch valid
  }

  int BN = (c.Nout % 128 == 0) ? 128 : 64;
  {
    static int forced = -1;  // VFS_CONV_BN=64|128|256 overrides the tile width (experiments)
    if (forced < 0) {
      const char* e = getenv("VFS_CONV_BN");
      forced = e ? atoi(e) : 0;
    }
    if (forced > 0 && c.Nout % forced == 0) BN = forced;
  }
  p.num_n_tiles = c.Nout / BN;
  {
    const uint64_t dims[3] = {c.Ktot, static_cast<uint64_t>(c.Nout), 2};
    const uint64_t strides[2] = {c.Ktot * 2, c.Ktot * c.Nout * 2};
    const uint32_t box_b[3] = {64u, static_cast<uint32_t>(BN), 2u};
    int rc = make_tmap_16b_sw128(&p.tmap_b, c.b_split, 3, dims, strides, box_b);
    if (rc != VFS_OK) return rc;
  }
  auto set_weight_map = [&](int box_rows) {
    const uint64_t dims[3] = {c.Ktot, static_cast<uint64_t>(c.Nout), 2};
    const uint64_t strides[2] = {c.Ktot * 2, c.Ktot * c.Nout * 2};
    const uint32_t box_b[3] = {64u, static_cast<uint32_t>(box_rows), 2u};
    return make_tmap_16b_sw128(&p.tmap_b, c.b_split, 3, dims, strides, box_b);
  };
  // fp32 output / BN statistics go through the legacy epilogue; the split-only contract uses the TMA epilogue
  static int force_legacy = -1;
  if (force_legacy < 0) {
    const char* e = getenv("VFS_CONV_LEGACY_EPILOGUE");
    force_legacy = (e && atoi(e) != 0) ? 1 : 0;
  }
  // (statistics with a split output and no residual ride on the TMA epilogue: the train-mode forward)
  const bool tma_stats = c.stats != nullptr && c.out_split != nullptr && c.out_f32 == nullptr &&
                         c.res_split == nullptr && !force_legacy;
  const bool legacy = force_legacy || c.out_f32 != nullptr || (c.stats != nullptr && !tma_stats) ||
                      c.out_split == nullptr;
  if (legacy) {
    if (BN == 256) {
      BN = 128;
      p.num_n_tiles = c.Nout / 128;
      int rc = set_weight_map(128);
      if (rc != VFS_OK) return rc;
    }
    if (BN == 128) return launch<128, 3, 0, false>(p, stream);
    return launch<64, 4, 0, false>(p, stream);
  }
  {
    // output (and residual) through the same possibly strided pixel view the legacy epilogue addresses by hand
    const uint32_t box_o[5] = {64u, static_cast<uint32_t>(p.tw), static_cast<uint32_t>(p.th),
                               static_cast<uint32_t>(p.tn), 2u};
    const uint64_t Co = static_cast<uint64_t>(c.Nout);
    const uint64_t dims[5] = {Co, static_cast<uint64_t>(p.Wo), static_cast<uint64_t>(p.Ho),
                              static_cast<uint64_t>(p.N), 2};
    const uint64_t strides[4] = {static_cast<uint64_t>(p.out_sx) * Co * 2,
                                 static_cast<uint64_t>(p.out_sy) * p.out_W * Co * 2,
                                 static_cast<uint64_t>(p.out_H) * p.out_W * Co * 2, c.out_plane * 2};
    const size_t view_off = (static_cast<size_t>(p.out_oy) * p.out_W + p.out_ox) * Co * 2;
    int rc = make_tmap_16b_sw128(&p.tmap_out, reinterpret_cast<char*>(c.out_split) + view_off, 5, dims, strides, box_o);
    if (rc != VFS_OK) return rc;
    if (c.res_split) {
      rc = make_tmap_16b_sw128(&p.tmap_res, reinterpret_cast<const char*>(c.res_split) + view_off, 5, dims, strides,
                               box_o);
      if (rc != VFS_OK) return rc;
    }
  }
  // CTA pairs (cta_group::2, 256-pixel x BNp tiles).  Policy 2 (default) uses them where they measured faster on B200
  // (profiles/r01_layers_pair_v8.log, r01_layers_pair_v9.log): 256-wide tiles with enough pair tiles to occupy most
  // TPCs and at least 8 K-chunks, or at least 4 when a residual streams through the epilogue (layer3 expand).
  // Shorter-K expand layers and small launches (per-video calls, SiamFC crops) keep the finer 1-CTA tiles.  Policy 1 forces pairs wherever
  // the shape allows (tests), 0 disables them.
  {
    int pair_mode, pair_min;
    conv_pair_policy(&pair_mode, &pair_min);
    const int BNp = (c.Nout % 256 == 0) ? 256 : ((c.Nout % 128 == 0) ? 128 : 0);
    if (pair_mode != 0 && BNp != 0 && p.num_m_tiles >= 2) {
      const int pair_tiles = ((p.num_m_tiles + 1) / 2) * (c.Nout / BNp);
      const int kchunks = p.num_taps * p.kchunks_per_tap;
      const bool profitable = BNp == 256 && pair_tiles >= pair_min && kchunks >= (c.res_split ? 4 : 8);
      if (pair_mode == 1 || profitable) {
        p.num_n_tiles = c.Nout / BNp;
        int rc = set_weight_map(BNp / 2);
        if (rc != VFS_OK) return rc;
        if (c.res_split) {
          if (BNp == 256) return launch<256, 2, 3, true, true>(p, stream);
          return launch<128, 2, 3, true, true>(p, stream);
        }
        if (tma_stats) {   // (two stages: the third would not leave room for the per-tile sums)
          if (BNp == 256) return launch<256, 2, 1, false, true, true>(p, stream);
          return launch<128, 4, 1, false, true, true>(p, stream);
        }
        if (BNp == 256) return launch<256, 3, 1, false, true>(p, stream);
        return launch<128, 4, 1, false, true>(p, stream);
      }
    }
  }
  if (c.res_split) {
    if (BN == 256) {
      BN = 128;
      p.num_n_tiles = c.Nout / 128;
      int rc = set_weight_map(128);
      if (rc != VFS_OK) return rc;
    }
    if (BN == 128) return launch<128, 2, 3, true>(p, stream);
    return launch<64, 2, 3, true>(p, stream);
  }
  if (tma_stats) {
    if (BN == 256) {   // (no room for the per-tile sums next to two 96 KB stages)
      BN = 128;
      p.num_n_tiles = c.Nout / 128;
      int rc = set_weight_map(128);
      if (rc != VFS_OK) return rc;
    }
    if (BN == 128) return launch<128, 3, 1, false, false, true>(p, stream);
    return launch<64, 4, 1, false, false, true>(p, stream);
  }
  if (BN == 256) return launch<256, 2, 1, false>(p, stream);
  if (BN == 128) return launch<128, 3, 1, false>(p, stream);
  return launch<64, 4, 1, false>(p, stream);
}

static int check_desc(const VfsConvDesc* d, const char* who) {
  VFS_REQUIRE(d->ksize == 1 || d->ksize == 3, VFS_ESHAPE, "%s: ksize %d unsupported", who, d->ksize);
  VFS_REQUIRE(d->stride == 1 || d->stride == 2, VFS_ESHAPE, "%s: stride %d unsupported", who, d->stride);
  VFS_REQUIRE(d->dilation >= 1, VFS_ESHAPE, "%s: dilation %d", who, d->dilation);
  VFS_REQUIRE(d->Cin % 64 == 0 && d->Cout % 64 == 0, VFS_ESHAPE, "%s: Cin/Cout must be multiples of 64", who);
  VFS_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0, VFS_ESHAPE, "%s: empty input", who);
  return VFS_OK;
}

int conv_bn_act_tc(const VfsConvDesc* d, const void* in_split, const void* w_split, const float* scale,
                   const float* shift, const void* residual_split, void* out_split, float* out_f32,
                   double* stats, cudaStream_t stream) {
  VFS_REQUIRE(d && in_split && w_split && scale && shift, VFS_EINVAL, "conv_bn_act: null argument");
  VFS_REQUIRE(out_split || out_f32, VFS_EINVAL, "conv_bn_act: no output buffer");
  int rc = check_desc(d, "conv_bn_act");
  if (rc != VFS_OK) return rc;
  const int N = d->N, H = d->H, W = d->W, Cin = d->Cin, Cout = d->Cout;
  const int k = d->ksize, s = d->stride, dil = (k == 1) ? 1 : d->dilation;
  const int pad = (k == 1) ? 0 : dil;
  const int Ho = (H + 2 * pad - dil * (k - 1) - 1) / s + 1;
  const int Wo = (W + 2 * pad - dil * (k - 1) - 1) / s + 1;

  ConvSpec c;
  memset(&c, 0, sizeof(c));
  c.a_split = in_split;
  c.a_plane = static_cast<size_t>(N) * H * W * Cin;
  c.N = N; c.H = H; c.W = W; c.C = Cin;
  c.view_stride = s;
  c.flat = (k == 1 && s == 1);
  c.num_taps = k * k;
  // activation views, one per stride parity (ph, pw): view[y', x'] = in[s*y' + ph, s*x' + pw]
  for (int r = 0; r < k; ++r) {
    for (int q = 0; q < k; ++q) {
      const int oh = r * dil - pad, ow = q * dil - pad;
      const int ph = ((oh % s) + s) % s, pw = ((ow % s) + s) % s;
      TapSpec& t = c.taps[r * k + q];
      t.view = ph * 2 + pw;
      t.dh = floordiv(oh - ph, s);
      t.dw = floordiv(ow - pw, s);
      t.koff = (r * k + q) * Cin;
    }
  }
  c.b_split = w_split;
  c.Nout = Cout;
  c.Ktot = static_cast<uint64_t>(k) * k * Cin;
  c.Ho = Ho; c.Wo = Wo;
  c.out_H = Ho; c.out_W = Wo; c.out_sy = 1; c.out_sx = 1;
  c.scale = scale; c.shift = shift;
  c.res_split = residual_split;
  c.out_split = out_split;
  c.out_plane = static_cast<size_t>(N) * Ho * Wo * Cout;
  c.out_f32 = out_f32;
  c.stats = stats;
  c.relu = d->relu;
  return run_spec(c, stream);
}

// ------------------------------------------------------------------------------------------------
// Data gradient of the same convolution (training backward): dX = conv_transpose(dZ, W).
//   d          forward descriptor (N, H, W = forward INPUT extent; Cin, Cout forward channels)
//   dz_split   split NHWC [N, Ho, Wo, Cout]
//   wt_split   split [2][Cin][k*k*Cout], K index = (r'*k + s')*Cout + co with the kernel flipped:
//              wt[ci][r'][s'][co] = w[co][ci][k-1-r'][k-1-s']        (vfs_pack_conv_weight_dgrad)
//   add_split  NULL or split NHWC [N,H,W,Cin] added to the result (gradient arriving through the other branch)
//   dx_split   split NHWC [N, H, W, Cin]
// Stride 1 is one launch of the forward kernel with the roles of Cin/Cout swapped.  Stride 2 is the sum over the
// four output-parity classes (dX[2a+pa, 2b+pb] only receives the taps with matching parity), each one launch that
// writes its strided view of dX; positions no tap reaches get the plain `add` term (or zero).
// ------------------------------------------------------------------------------------------------
int conv_set_pair_policy(int mode, int min_pair_tiles) {
  VFS_REQUIRE(mode >= 0 && mode <= 2 && min_pair_tiles >= 1, VFS_EINVAL, "conv_set_pair_policy: mode %d min %d", mode,
              min_pair_tiles);
  g_pair_mode = mode;
  g_pair_min = min_pair_tiles;
  return VFS_OK;
}

int conv_set_trace(long long* buffer, int events_per_role) {
  VFS_CUDA_OK(cudaMemcpyToSymbol(g_conv_trace, &buffer, sizeof(buffer)));
  VFS_CUDA_OK(cudaMemcpyToSymbol(g_conv_trace_cap, &events_per_role, sizeof(events_per_role)));
  return VFS_OK;
}

int conv_dgrad_tc(const VfsConvDesc* d, const void* dz_split, const void* wt_split, const float* ones,
                  const float* zeros, const void* add_split, void* dx_split, cudaStream_t stream) {
  VFS_REQUIRE(d && dz_split && wt_split && ones && zeros && dx_split, VFS_EINVAL, "conv_dgrad: null argument");
  int rc = check_desc(d, "conv_dgrad");
  if (rc != VFS_OK) return rc;
  const int N = d->N, H = d->H, W = d->W, Cin = d->Cin, Cout = d->Cout;
  const int k = d->ksize, s = d->stride, dil = (k == 1) ? 1 : d->dilation;
  const int pad = (k == 1) ? 0 : dil;
  const int Ho = (H + 2 * pad - dil * (k - 1) - 1) / s + 1;
  const int Wo = (W + 2 * pad - dil * (k - 1) - 1) / s + 1;
  VFS_REQUIRE(s == 1 || dil == 1, VFS_ESHAPE, "conv_dgrad: stride 2 with dilation > 1 unsupported");

  ConvSpec c;
  memset(&c, 0, sizeof(c));
  c.a_split = dz_split;
  c.a_plane = static_cast<size_t>(N) * Ho * Wo * Cout;
  c.N = N; c.H = Ho; c.W = Wo; c.C = Cout;
  c.view_stride = 1;
  c.b_split = wt_split;
  c.Nout = Cin;
  c.Ktot = static_cast<uint64_t>(k) * k * Cout;
  c.scale = ones; c.shift = zeros;
  c.res_split = add_split;
  c.out_split = dx_split;
  c.out_plane = static_cast<size_t>(N) * H * W * Cin;
  c.out_H = H; c.out_W = W;
  if (s == 1) {
    // dX[h] = sum_r' dZ[h + (r'-c)*dil] * wt[r'] : a same-size correlation with the flipped kernel
    c.flat = (k == 1);
    c.num_taps = k * k;
    for (int r = 0; r < k; ++r)
      for (int q = 0; q < k; ++q) {
        TapSpec& t = c.taps[r * k + q];
        t.view = 0;
        t.dh = r * dil - pad;
        t.dw = q * dil - pad;
        t.koff = (r * k + q) * Cout;
      }
    c.Ho = H; c.Wo = W;
    c.out_sy = 1; c.out_sx = 1; c.out_oy = 0; c.out_ox = 0;
    return run_spec(c, stream);
  }
  if (k == 1) {
    // only the (even, even) positions receive a tap; everything else is the `add` term (or zero)
    const size_t bytes = 2 * c.out_plane * sizeof(h16);
    if (add_split) VFS_CUDA_OK(cudaMemcpyAsync(dx_split, add_split, bytes, cudaMemcpyDeviceToDevice, stream));
    else VFS_CUDA_OK(cudaMemsetAsync(dx_split, 0, bytes, stream));
  }
  // stride 2: forward z[ho] += x[2*ho + r - pad] * w[r]  =>  dX[i] = sum_{r : (i + pad - r) even} dZ[(i+pad-r)/2] * w[r]
  // in flipped-kernel coordinates r' = k-1-r the weight chunk of forward tap r is (k-1-r).
  for (int pa = 0; pa < 2; ++pa) {
    for (int pb = 0; pb < 2; ++pb) {
      const int Hc = (H - pa + 1) / 2, Wc = (W - pb + 1) / 2;  // number of positions 2a+pa < H
      if (Hc <= 0 || Wc <= 0) continue;
      int nt = 0;
      int rs[3], dhs[3], qs[3], dws[3];
      int nr = 0, nq = 0;
      for (int r = 0; r < k; ++r)
        if (((pa + pad - r) % 2 + 2) % 2 == 0) { rs[nr] = r; dhs[nr] = floordiv(pa + pad - r, 2); ++nr; }
      for (int q = 0; q < k; ++q)
        if (((pb + pad - q) % 2 + 2) % 2 == 0) { qs[nq] = q; dws[nq] = floordiv(pb + pad - q, 2); ++nq; }
      for (int i = 0; i < nr; ++i)
        for (int j = 0; j < nq; ++j) {
          TapSpec& t = c.taps[nt++];
          t.view = 0;
          t.dh = dhs[i];
          t.dw = dws[j];
          t.koff = ((k - 1 - rs[i]) * k + (k - 1 - qs[j])) * Cout;
        }
      c.flat = false;
      c.Ho = Hc; c.Wo = Wc;
      c.out_sy = 2; c.out_sx = 2; c.out_oy = pa; c.out_ox = pb;
      if (nt == 0) continue;  // no tap reaches this parity class (1x1/s2): pre-filled below
      c.num_taps = nt;
      rc = run_spec(c, stream);
      if (rc != VFS_OK) return rc;
    }
  }
  return VFS_OK;
}

VFS_DEFINE_OVERFLOW_ACCESSOR(overflow_conv)

}  // namespace vfs
