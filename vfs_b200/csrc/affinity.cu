// Fused restricted-attention label propagation for sm_100a (DAVIS inference hot loop).
//
// Replaces the reference's per-32-query chunk loop  einsum -> /T -> masked_fill(-inf) -> topk -> index_select ->
// softmax -> einsum  (mmaction/models/common/local_attention.py:287-342) by two kernels:
//
//  A  attn_scores_topk_kernel : S = Q K^T on tcgen05 (split-fp16 operands, 3 products per K-chunk, fp32 TMEM
//     accumulators), for one tile of 128 queries (8 rows x 16 cols of the feature map) against 128-key tiles
//     (8 x 16) that intersect the query tile's neighbour window -- key tiles outside the window are never loaded.
//     Each epilogue thread owns one query row of the accumulator and keeps a sorted top-k (value, key index) list
//     in registers; the radius mask is an integer predicate on (dy, dx), the HW x T*HW affinity never exists in
//     memory.  Work is split into units (query tile, key frame, window slice) so that all 148 SMs are busy even
//     for a single frame pair; every unit writes its partial top-k.
//  B  attn_merge_propagate_kernel : per query, merge the partial lists, scale by 1/temperature, softmax over
//     the k survivors (or clamp^2, mode 'cosine'), gather the k value vectors and write the propagated labels.
//
// plus the feature preparation (L2 normalisation over channels fused with the NCHW->split-NHWC transpose,
// local_attention.py:277-279).
#include <math.h>

#include "host_common.h"
#include "ptx.cuh"

namespace vfs {

constexpr int kQTileW = 16, kQTileH = 8;  // 128 queries / keys per tile
// Two forms of the scores kernel.  Narrow: one 128-key tile per step (N = 128 MMAs), 3 stages of 64 KB.  WIDE: two key
// tiles per step (N = 256 MMAs: the query tile is staged once per two key tiles and read from shared memory once per
// 256 key columns), 2 stages of 96 KB, 2 x 256 TMEM columns.
template <bool WIDE>
struct AttnCfg {
  static constexpr int kStages = WIDE ? 2 : 3;
  static constexpr int kKeyBytes = (WIDE ? 2 : 1) * 2 * 16384;  // K hi + K lo planes of the step's key tile(s)
  static constexpr int kStageBytes = 2 * 16384 + kKeyBytes;     // Q hi, Q lo, then K hi, K lo
  static constexpr int kNC = WIDE ? 256 : 128;                  // key columns per accumulator stage
};
constexpr int kAttnEpiHalves = 2;                // epilogue warps per TMEM lane quarter (each owns 64 key columns)
constexpr int kAttnScratchBytes = kAttnEpiHalves * 32 * 128 * 4;  // one 32-column score slab per row and half
constexpr int kAttnSmemBytes = 3 * 65536 + 256 + kAttnScratchBytes + 1024;  // both forms: 192 KB of stages
constexpr int kAttnThreads = 64 + 128 * kAttnEpiHalves;  // TMA warp, MMA warp, 4 x kAttnEpiHalves epilogue warps
constexpr int kMaxKeyFrames = 32;
constexpr int kMaxProblems = 32;   // query frames per launch
constexpr int kMaxKeySlots = 256;  // problems x key frames per launch

struct alignas(64) AttnParams {
  CUtensorMap tmap_q;  // {C, W, H, q frames, 2}
  CUtensorMap tmap_k;  // {C, W, H, k frames, 2}
  CUtensorMap tmap_k1; // same tensor, box of ONE plane (WIDE: the two key tiles interleave per plane in smem)
  int H, W, kchunks;
  int T;  // number of key frame slots per problem
  int num_problems;
  short q_ids[kMaxProblems];      // bank frame of each problem's query
  short frame_ids[kMaxKeySlots];  // [problem][slot] -> key bank frame
  int mask_mode;  // 0 none, 1 circle (dy^2+dx^2 < ry^2), 2 square (|dy| <= ry, |dx| <= rx)
  int ry, rx;
  int non_mask_len;
  int q_tiles_x, q_tiles_y, splits;
  int num_units;
  // key tile = kth x ktw pixels (kth * ktw <= 128 accumulator columns; the WIDE form picks the shape that covers the
  // neighbour window of a query tile with the fewest tiles, e.g. 5 x 25 for radius 18: 18 tiles instead of 24 of 8 x 16)
  int kth, ktw, kvalid;
  int stage_tx_bytes;
  float* part_val;  // [problems*T*splits*kAttnEpiHalves][KMAX][HW]
  int* part_idx;
};

struct KeyWindow {
  int wy0, wx0, ny, nx;  // origin and number of key tiles
};

__device__ __forceinline__ KeyWindow key_window(const AttnParams& p, int qy0, int qx0, int t) {
  KeyWindow w;
  int wy1, wx1;
  if (p.mask_mode == 0 || t < p.non_mask_len) {
    w.wy0 = 0; w.wx0 = 0; wy1 = p.H - 1; wx1 = p.W - 1;
  } else {
    const int ey = (p.mask_mode == 1) ? p.ry - 1 : p.ry;
    const int ex = (p.mask_mode == 1) ? p.ry - 1 : p.rx;
    w.wy0 = max(0, qy0 - ey);
    w.wx0 = max(0, qx0 - ex);
    wy1 = min(p.H - 1, min(qy0 + kQTileH - 1, p.H - 1) + ey);
    wx1 = min(p.W - 1, min(qx0 + kQTileW - 1, p.W - 1) + ex);
  }
  const int wh = wy1 - w.wy0 + 1, ww = wx1 - w.wx0 + 1;
  w.ny = wh > 0 ? (wh + p.kth - 1) / p.kth : 0;
  w.nx = ww > 0 ? (ww + p.ktw - 1) / p.ktw : 0;
  if (w.ny == 0 || w.nx == 0) {
    w.ny = 0;
    w.nx = 1;  // keeps j / nx well defined for the (empty) tile loop
  }
  return w;
}

struct UnitInfo {
  int b, qy0, qx0, t, j_begin, j_end;  // j_*: steps (WIDE: pairs of key tiles)
  int n;                               // key tiles in the window
  KeyWindow w;
};

template <bool WIDE>
__device__ __forceinline__ UnitInfo decode_unit(const AttnParams& p, int unit) {
  UnitInfo u;
  // unit order: window slice fastest, then query tile, then key frame, then problem.  The CTAs of a wave then work
  // on ONE key frame (26 MB of split features at 480p: L2-resident) and every frame is read from DRAM about once;
  // with the key frame varying faster than the query tile all T frames were live at once (550 MB at T = 21) and
  // the window overlap of neighbouring query tiles was re-read from DRAM: 2.63 GB per launch against 0.58 GB of
  // compulsory bytes (profiles/r02_affinity_480p_ncu.csv).
  const int s = unit % p.splits;
  const int r = unit / p.splits;
  const int q_tiles = p.q_tiles_x * p.q_tiles_y;
  const int qt = r % q_tiles;
  const int r3 = r / q_tiles;
  u.t = r3 % p.T;
  u.b = r3 / p.T;
  u.qx0 = (qt % p.q_tiles_x) * kQTileW;
  u.qy0 = (qt / p.q_tiles_x) * kQTileH;
  u.w = key_window(p, u.qy0, u.qx0, u.t);
  u.n = u.w.ny * u.w.nx;
  const int steps = WIDE ? (u.n + 1) / 2 : u.n;
  const int per = (steps + p.splits - 1) / p.splits;
  u.j_begin = min(steps, s * per);
  u.j_end = min(steps, u.j_begin + per);
  return u;
}
// origin of key tile j of the unit's window; a tile past the end (odd window in the WIDE form) lies outside the image:
// its TMA load is pure zero fill and the epilogue skips its columns
__device__ __forceinline__ void key_tile_origin(const AttnParams& p, const UnitInfo& u, int j, int& ky0, int& kx0) {
  if (j < u.n) {
    ky0 = u.w.wy0 + (j / u.w.nx) * p.kth;
    kx0 = u.w.wx0 + (j % u.w.nx) * p.ktw;
  } else {
    ky0 = p.H;
    kx0 = 0;
  }
}

template <int KMAX>
__device__ __forceinline__ void topk_insert(float (&v)[KMAX], int (&id)[KMAX], float x, int idx) {
#pragma unroll
  for (int i = KMAX - 1; i >= 0; --i) {
    if (i > 0 && x > v[i - 1]) {
      v[i] = v[i - 1];
      id[i] = id[i - 1];
    } else if (x > v[i]) {
      v[i] = x;
      id[i] = idx;
    }
  }
}

// KTH x KTW: key tile shape (compile time, so that the epilogue's column -> (row, column) map unrolls into constants;
// a run-time shape with a lookup table was measured 12 % slower end to end: the top-k epilogue then paces the loop)
template <int KMAX, bool WIDE, int KTH, int KTW>
__global__ void __launch_bounds__(kAttnThreads, 1) attn_scores_topk_kernel(const __grid_constant__ AttnParams p) {
  static_assert(KTH * KTW <= 128 && KTW <= 32, "key tile: at most 128 pixels, rows of at most 32");
  using A = AttnCfg<WIDE>;
  constexpr int kAttnStages = A::kStages;
  constexpr int kAttnStageBytes = A::kStageBytes;
  constexpr int kNC = A::kNC;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kAttnStages * kAttnStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kAttnStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kAttnStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kAttnStages + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * kAttnStages + 4);
  const uint32_t scratch_base = bar_base + 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 2 * kNC;  // 2 accumulator stages

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_q);
    tma_prefetch_desc(&p.tmap_k);
    if (WIDE) tma_prefetch_desc(&p.tmap_k1);
    for (int s = 0; s < kAttnStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4 * kAttnEpiHalves);
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

  if (warp == 0) {
    // ======================= TMA producer =======================
    int stage = 0;
    uint32_t phase = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      const UnitInfo u = decode_unit<WIDE>(p, unit);
      const int frame = p.frame_ids[u.b * p.T + u.t];
      const int qframe = p.q_ids[u.b];
      for (int j = u.j_begin; j < u.j_end; ++j) {
        int ky0, kx0, ky1 = 0, kx1 = 0;
        key_tile_origin(p, u, WIDE ? 2 * j : j, ky0, kx0);
        if (WIDE) key_tile_origin(p, u, 2 * j + 1, ky1, kx1);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(empty_bar(stage), phase ^ 1u, 100 + stage);
          if (lane == 0) {
            const uint32_t sq = smem_base + stage * kAttnStageBytes;
            mbar_arrive_expect_tx(full_bar(stage), static_cast<uint32_t>(p.stage_tx_bytes));
            tma_load_5d(sq, &p.tmap_q, full_bar(stage), kc * 64, u.qx0, u.qy0, qframe, 0);
            if (WIDE) {
              // K hi = [tile 2j rows 0..127 | tile 2j+1 rows 128..255], then K lo likewise: one plane per load
              const uint32_t sk = sq + 32768;
              tma_load_5d(sk, &p.tmap_k1, full_bar(stage), kc * 64, kx0, ky0, frame, 0);
              tma_load_5d(sk + 16384, &p.tmap_k1, full_bar(stage), kc * 64, kx1, ky1, frame, 0);
              tma_load_5d(sk + 32768, &p.tmap_k1, full_bar(stage), kc * 64, kx0, ky0, frame, 1);
              tma_load_5d(sk + 49152, &p.tmap_k1, full_bar(stage), kc * 64, kx1, ky1, frame, 1);
            } else {
              tma_load_5d(sq + 32768, &p.tmap_k, full_bar(stage), kc * 64, kx0, ky0, frame, 0);
            }
          }
          __syncwarp();
          if (++stage == kAttnStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    constexpr uint32_t idesc = umma_idesc_f16_f32(128, kNC);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      const UnitInfo u = decode_unit<WIDE>(p, unit);
      for (int j = u.j_begin; j < u.j_end; ++j) {
        mbar_wait(tempty_bar(as), aphase ^ 1u, 200 + as);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kNC;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(full_bar(stage), phase, 300 + stage);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t q_hi = smem_base + stage * kAttnStageBytes, q_lo = q_hi + 16384;
            const uint32_t k_hi = q_hi + 32768, k_lo = k_hi + A::kKeyBytes / 2;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t koff = k * 32;
              const uint64_t dq_hi = umma_desc_sw128_kmajor(q_hi + koff), dq_lo = umma_desc_sw128_kmajor(q_lo + koff);
              const uint64_t dk_hi = umma_desc_sw128_kmajor(k_hi + koff), dk_lo = umma_desc_sw128_kmajor(k_lo + koff);
              umma_f16(d_tmem, dq_lo, dk_hi, idesc, (kc | k) != 0 ? 1u : 0u);
              umma_f16(d_tmem, dq_hi, dk_lo, idesc, 1u);
              umma_f16(d_tmem, dq_hi, dk_hi, idesc, 1u);
            }
            umma_commit(empty_bar(stage));
            if (kc == p.kchunks - 1) umma_commit(tfull_bar(as));
          }
          __syncwarp();
          if (++stage == kAttnStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (++as == 2) {
          as = 0;
          aphase ^= 1u;
        }
      }
    }
  } else {
    // ======================= epilogue: per-query running top-k =======================
    // Two warps per TMEM lane quarter: warp half h scans key columns [64h, 64h+64) of every key tile and keeps its own
    // sorted list; the lists are merged by kernel B like the window slices.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int dqx = row % kQTileW, dqy = row / kQTileW;
    const int HW = p.H * p.W;
    int as = 0;
    uint32_t aphase = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      const UnitInfo u = decode_unit<WIDE>(p, unit);
      const int qy = u.qy0 + dqy, qx = u.qx0 + dqx;
      const bool q_valid = (qy < p.H) && (qx < p.W);
      const bool masked = (p.mask_mode != 0) && (u.t >= p.non_mask_len);
      const int r2 = p.ry * p.ry;
      float tv[KMAX];
      int ti[KMAX];
#pragma unroll
      for (int i = 0; i < KMAX; ++i) {
        tv[i] = -INFINITY;
        ti[i] = 0;
      }
      for (int j = u.j_begin; j < u.j_end; ++j) {
        // narrow: this warp half scans columns [64h, 64h+64) of the step's key tile; WIDE: half h owns key tile 2j+h
        // (accumulator columns [128h, 128h+128))
        const int jt = WIDE ? 2 * j + half : j;
        int ky0, kx0;
        key_tile_origin(p, u, jt, ky0, kx0);
        mbar_wait(tfull_bar(as), aphase, 400 + as);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kNC + (WIDE ? 128 * half : 0);
        // per key tile: squared horizontal distances / in-image flags of its KTW columns and the per-row limits
        // (circle: dx^2 < r^2 - dy^2; square: |dx| <= rx and |dy| <= ry; unmasked: always), shared by all slabs
        int dx2[KTW], lim[KTH];
        uint32_t xok = 0, rowok = 0;
#pragma unroll
        for (int c = 0; c < KTW; ++c) {
          const int dx = kx0 + c - qx;
          dx2[c] = (p.mask_mode == 2) ? abs(dx) : dx * dx;
          if (kx0 + c < p.W) xok |= (1u << c);
        }
#pragma unroll
        for (int r = 0; r < KTH; ++r) {
          const int ky = ky0 + r;
          const int dy = ky - qy;
          bool ok = ky < p.H;
          int l = 0x7fffffff;
          if (masked) {
            if (p.mask_mode == 1) l = r2 - dy * dy;
            else {
              l = p.rx + 1;
              ok = ok && (abs(dy) <= p.ry);
            }
          }
          lim[r] = l;
          if (ok) rowok |= (1u << r);
        }
        constexpr int kValid = KTH * KTW;
        const int c_begin = WIDE ? 0 : 64 * half;
        int c_end = (jt < u.n) ? (WIDE ? 128 : 64 * half + 64) : c_begin;   // phantom tile: nothing to scan
        if (c_end > kValid) c_end = kValid;                                  // columns past the tile's pixels: stale rows
#pragma unroll
        for (int slab_i = 0; slab_i < 4; ++slab_i) {
          const int c0 = slab_i * 32;                                        // compile-time after unrolling
          if (c0 < c_begin || c0 >= c_end) continue;
          uint32_t acc[32];
          tmem_ld_32x32b_x32(t_row + c0, acc);
          tmem_ld_wait();
          // Two passes: (1) a bit mask of the columns that are inside the image / radius and beat the list's current
          // k-th value -- a conservative filter, the k-th value only grows; (2) only those columns go through the sorted
          // insertion, in ascending column order (identical result to testing every column, but the warp executes the
          // ~40-instruction insertion max-over-lanes(#candidates) times instead of 32 times).  The dynamic column index
          // of pass 2 reads the slab back from shared memory.
          const float thr = tv[KMAX - 1];
          uint32_t cand = 0;
#pragma unroll
          for (int cc = 0; cc < 32; ++cc) {
            const int r = c0 + cc;                                           // constants: r, r / KTW, r % KTW
            if (r < kValid) {
              const bool ok = ((rowok >> (r / KTW)) & 1u) && ((xok >> (r % KTW)) & 1u) && (dx2[r % KTW] < lim[r / KTW]);
              if (ok && __uint_as_float(acc[cc]) > thr) cand |= (1u << cc);
            }
          }
          if (__any_sync(0xffffffffu, cand != 0)) {
            const uint32_t slab = scratch_base + static_cast<uint32_t>(half) * (32 * 128 * 4) +
                                  static_cast<uint32_t>(row) * 4u;
#pragma unroll
            for (int jj = 0; jj < 32; ++jj)
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(slab + jj * 512), "r"(acc[jj]) : "memory");
            while (cand != 0) {
              const int jj = __ffs(cand) - 1;
              cand &= cand - 1;
              float sc;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(sc) : "r"(slab + jj * 512));
              if (sc > tv[KMAX - 1]) {
                const int r = c0 + jj;
                const int ky = ky0 + r / KTW, kx = kx0 + r % KTW;
                topk_insert<KMAX>(tv, ti, sc, u.t * HW + ky * p.W + kx);
              }
            }
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
      if (q_valid) {
        const int slot = ((u.b * p.T + u.t) * p.splits + (unit % p.splits)) * kAttnEpiHalves + half;
        const size_t base = static_cast<size_t>(slot) * KMAX * HW + static_cast<size_t>(qy) * p.W + qx;
#pragma unroll
        for (int i = 0; i < KMAX; ++i) {
          p.part_val[base + static_cast<size_t>(i) * HW] = tv[i];
          p.part_idx[base + static_cast<size_t>(i) * HW] = ti[i];
        }
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

// ------------------------------------------------------------------------------------------------
// Kernel B: merge partial lists, softmax over the k survivors, gather values.
// ------------------------------------------------------------------------------------------------
struct MergeParams {
  const float* part_val;
  const int* part_idx;
  int slots, HW, topk, mode;  // slots per problem; mode 0 softmax, 1 cosine
  float temperature;
  // element (problem b, slot t, channel c, pos) at
  //   values[b*v_batch_stride + val_ids[b*T+t]*v_frame_stride + c*v_chan_stride + pos]
  const float* values;
  long long v_batch_stride, v_frame_stride, v_chan_stride;
  int Cv, T;
  short val_ids[kMaxKeySlots];
  float* out;      // [problems][Cv][HW]
  float* out_val;  // optional [problems][topk][HW] (affinity / temperature of the selected keys)
  int* out_idx;    // optional [problems][topk][HW] (flat key index slot*HW + pos)
};

// Block = 32 queries x 8 sub-lanes (warp index = sub-lane, lane = query: every global access of a warp covers 32
// consecutive queries).  Sub-lane w reduces a contiguous range of partial lists to a local top-k, warp 0 merges the
// eight sorted lists (in slot order, so ties resolve exactly like a serial pass over the slots), computes the
// softmax weights, and all eight warps split the value channels of the propagation.
constexpr int kMergeSub = 8;

template <int KMAX>
__global__ void __launch_bounds__(32 * kMergeSub) attn_merge_propagate_kernel(const MergeParams p) {
  __shared__ float sv[kMergeSub][KMAX][32];
  __shared__ int si[kMergeSub][KMAX][32];
  __shared__ float sw[KMAX][32];
  const int lane = threadIdx.x & 31, sub = threadIdx.x >> 5;
  const int qi = blockIdx.x * 32 + lane;
  const int b = blockIdx.y;
  const bool valid = qi < p.HW;
  float tv[KMAX];
  int ti[KMAX];
#pragma unroll
  for (int i = 0; i < KMAX; ++i) {
    tv[i] = -INFINITY;
    ti[i] = 0;
  }
  const int per = (p.slots + kMergeSub - 1) / kMergeSub;
  const int s_end = min(p.slots, (sub + 1) * per);
  if (valid) {
    for (int s = sub * per; s < s_end; ++s) {
      const size_t base = (static_cast<size_t>(b) * p.slots + s) * KMAX * p.HW + qi;
#pragma unroll
      for (int i = 0; i < KMAX; ++i) {
        const float v = p.part_val[base + static_cast<size_t>(i) * p.HW];
        if (v > tv[KMAX - 1]) topk_insert<KMAX>(tv, ti, v, p.part_idx[base + static_cast<size_t>(i) * p.HW]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < KMAX; ++i) {
    sv[sub][i][lane] = tv[i];
    si[sub][i][lane] = ti[i];
  }
  __syncthreads();
  if (sub == 0) {
    for (int w = 1; w < kMergeSub; ++w) {
#pragma unroll 1
      for (int i = 0; i < KMAX; ++i) {
        const float v = sv[w][i][lane];
        if (!(v > tv[KMAX - 1])) break;  // the list is sorted: nothing further can enter
        topk_insert<KMAX>(tv, ti, v, si[w][i][lane]);
      }
    }
    float wgt[KMAX];
    float wsum = 0.0f;
    const float vmax = __fdiv_rn(tv[0], p.temperature);
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
      const float a = __fdiv_rn(tv[i], p.temperature);  // reference divides the whole affinity by temperature
      if (i < p.topk) {
        if (valid && p.out_val) p.out_val[(static_cast<size_t>(b) * p.topk + i) * p.HW + qi] = a;
        if (valid && p.out_idx) p.out_idx[(static_cast<size_t>(b) * p.topk + i) * p.HW + qi] = ti[i];
        if (p.mode == 0) {
          wgt[i] = expf(a - vmax);
        } else {
          const float c = fmaxf(a, 0.0f);
          wgt[i] = c * c;
        }
        wsum += wgt[i];
      } else {
        wgt[i] = 0.0f;
      }
    }
    const float inv = (p.mode == 0) ? __fdiv_rn(1.0f, wsum) : 1.0f;
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
      sw[i][lane] = (p.mode == 0) ? wgt[i] * inv : wgt[i];
      si[0][i][lane] = ti[i];
    }
  }
  __syncthreads();
  if (!valid) return;
  for (int c = sub; c < p.Cv; c += kMergeSub) {
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
      if (i < p.topk) {
        const int idx = si[0][i][lane];
        const int slot = idx / p.HW, pos = idx - slot * p.HW;
        const float val = p.values[b * p.v_batch_stride + p.val_ids[b * p.T + slot] * p.v_frame_stride +
                                   c * p.v_chan_stride + pos];
        acc = fmaf(val, sw[i][lane], acc);
      }
    }
    p.out[(static_cast<size_t>(b) * p.Cv + c) * p.HW + qi] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// Feature preparation: optional L2 normalisation over channels (F.normalize, eps 1e-12) fused with the
// conversion to split NHWC.
// ------------------------------------------------------------------------------------------------
__global__ void pixel_inv_norm_nchw_kernel(const float* __restrict__ in, float* __restrict__ inv, int C, int HW) {
  const int n = blockIdx.y;
  const int pidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (pidx >= HW) return;
  const float* src = in + static_cast<size_t>(n) * C * HW + pidx;
  float s = 0.0f;
  for (int c = 0; c < C; ++c) {
    const float x = src[static_cast<size_t>(c) * HW];
    s = fmaf(x, x, s);
  }
  inv[static_cast<size_t>(n) * HW + pidx] = 1.0f / fmaxf(sqrtf(s), 1e-12f);
}

__global__ void nchw_to_split_scaled_kernel(const float* __restrict__ in, const float* __restrict__ pix_scale,
                                            h16* __restrict__ out_hi, h16* __restrict__ out_lo,
                                            int C, int HW, int Cs) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* src = in + static_cast<size_t>(n) * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, pp = p0 + threadIdx.x;
    float v = 0.0f;
    if (c < C && pp < HW) {
      v = src[static_cast<size_t>(c) * HW + pp];
      if (pix_scale) v *= pix_scale[static_cast<size_t>(n) * HW + pp];
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int pp = p0 + i, c = c0 + threadIdx.x;
    if (pp < HW && c < C) {
      h16 hi, lo;
      split16(tile[threadIdx.x][i], hi, lo);
      const size_t o = (static_cast<size_t>(n) * HW + pp) * Cs + c;  // Cs >= C: channel-padded destination rows
      out_hi[o] = hi;
      out_lo[o] = lo;
    }
  }
}

// split NHWC -> L2-normalised split NHWC; one warp per pixel, C multiple of 64.
__global__ void normalize_split_kernel(const h16* __restrict__ in_hi, const h16* __restrict__ in_lo,
                                       h16* __restrict__ out_hi, h16* __restrict__ out_lo,
                                       size_t num_pixels, int C) {
  const size_t pix = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (pix >= num_pixels) return;
  const size_t base = pix * C;
  float s = 0.0f;
  for (int c = lane * 2; c < C; c += 64) {
    const uint32_t h = *reinterpret_cast<const uint32_t*>(in_hi + base + c);
    const uint32_t l = *reinterpret_cast<const uint32_t*>(in_lo + base + c);
    const float a = lo16_to_float(h) + lo16_to_float(l);
    const float b = hi16_to_float(h) + hi16_to_float(l);
    s = fmaf(a, a, s);
    s = fmaf(b, b, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
  for (int c = lane * 2; c < C; c += 64) {
    const uint32_t h = *reinterpret_cast<const uint32_t*>(in_hi + base + c);
    const uint32_t l = *reinterpret_cast<const uint32_t*>(in_lo + base + c);
    const float a = (lo16_to_float(h) + lo16_to_float(l)) * inv;
    const float b = (hi16_to_float(h) + hi16_to_float(l)) * inv;
    h16 ah, al, bh, bl;
    split16(a, ah, al);
    split16(b, bh, bl);
    *reinterpret_cast<uint32_t*>(out_hi + base + c) = pack16x2(ah, bh);
    *reinterpret_cast<uint32_t*>(out_lo + base + c) = pack16x2(al, bl);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int features_to_split(const float* in_nchw, void* out_split, void* inv_norm_ws, int N, int C, int H, int W,
                      int normalize, int c_stride, long long plane_stride, cudaStream_t s) {
  VFS_REQUIRE(in_nchw && out_split, VFS_EINVAL, "features_to_split: null argument");
  VFS_REQUIRE(!normalize || inv_norm_ws, VFS_EINVAL, "features_to_split: normalisation needs a workspace");
  VFS_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, VFS_ESHAPE, "features_to_split: empty tensor");
  const int HW = H * W;
  float* inv = reinterpret_cast<float*>(inv_norm_ws);
  if (normalize) {
    dim3 grid((HW + 127) / 128, N);
    pixel_inv_norm_nchw_kernel<<<grid, 128, 0, s>>>(in_nchw, inv, C, HW);
    VFS_CUDA_OK(cudaGetLastError());
  }
  if (c_stride <= 0) c_stride = C;
  if (plane_stride <= 0) plane_stride = static_cast<long long>(N) * HW * c_stride;
  VFS_REQUIRE(c_stride >= C && plane_stride >= static_cast<long long>(N) * HW * c_stride, VFS_EINVAL,
              "features_to_split: destination strides smaller than the tensor");
  h16* hi = reinterpret_cast<h16*>(out_split);
  h16* lo = hi + plane_stride;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
  nchw_to_split_scaled_kernel<<<grid, block, 0, s>>>(in_nchw, normalize ? inv : nullptr, hi, lo, C, HW, c_stride);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int normalize_split(const void* in_split, void* out_split, long long num_pixels, int C, long long in_plane_stride,
                    long long out_plane_stride, cudaStream_t s) {
  VFS_REQUIRE(in_split && out_split, VFS_EINVAL, "normalize_split: null argument");
  VFS_REQUIRE(C % 64 == 0 && num_pixels > 0, VFS_ESHAPE, "normalize_split: C must be a multiple of 64");
  const h16* ih = reinterpret_cast<const h16*>(in_split);
  h16* oh = reinterpret_cast<h16*>(out_split);
  const long long threads = num_pixels * 32;
  const int blocks = static_cast<int>((threads + 255) / 256);
  normalize_split_kernel<<<blocks, 256, 0, s>>>(ih, ih + in_plane_stride, oh, oh + out_plane_stride,
                                                static_cast<size_t>(num_pixels), C);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

static int attn_kmax(int topk) { return topk <= 10 ? 10 : 16; }

// 1 (default) = WIDE form (two key tiles per step), 0 = one key tile per step.  VFS_ATTN_WIDE sets the default,
// vfs_attention_set_wide changes it at run time (tests, tuning).  Results are identical.
static int g_attn_wide = -1;
static bool attn_use_wide() {
  if (g_attn_wide < 0) {
    const char* e = getenv("VFS_ATTN_WIDE");
    g_attn_wide = e ? (atoi(e) != 0) : 1;
  }
  return g_attn_wide != 0;
}
int attention_set_wide(int mode) {
  VFS_REQUIRE(mode == 0 || mode == 1, VFS_EINVAL, "attention_set_wide: mode %d", mode);
  g_attn_wide = mode;
  return VFS_OK;
}

static int attn_splits(const VfsAttnDesc* d, int B) {
  const int q_tiles = ((d->W + kQTileW - 1) / kQTileW) * ((d->H + kQTileH - 1) / kQTileH);
  const int target = 2 * device_sm_count();
  const int base = q_tiles * d->T * (B > 0 ? B : 1);
  int splits = (target + base - 1) / base;
  if (splits < 1) splits = 1;
  if (splits > 8) splits = 8;
  return splits;
}

size_t attention_workspace_bytes(const VfsAttnDesc* d, int B) {
  if (!d || d->T <= 0 || B <= 0) return 0;
  const size_t slots = static_cast<size_t>(B) * d->T * attn_splits(d, B) * kAttnEpiHalves;
  return slots * attn_kmax(d->topk) * d->H * d->W * (sizeof(float) + sizeof(int));
}

int masked_attention_batched(const VfsAttnDesc* d, int B, const void* q_bank_split, long long q_plane_stride,
                             int q_bank_frames, const int* q_ids, const void* k_bank_split, long long k_plane_stride,
                             int k_bank_frames, const int* key_ids, const float* values, const int* val_ids,
                             long long v_batch_stride, long long v_frame_stride, long long v_chan_stride, float* out,
                             float* out_topk_val, int* out_topk_idx, void* workspace, size_t workspace_bytes,
                             cudaStream_t stream) {
  VFS_REQUIRE(d && q_bank_split && q_ids && k_bank_split && key_ids && values && val_ids && out && workspace,
              VFS_EINVAL, "masked_attention: null argument");
  VFS_REQUIRE(d->H > 0 && d->W > 0 && d->C > 0 && d->C % 64 == 0, VFS_ESHAPE,
              "masked_attention: C=%d must be a positive multiple of 64", d->C);
  VFS_REQUIRE(d->T >= 1 && d->T <= kMaxKeyFrames, VFS_ESHAPE, "masked_attention: T=%d outside [1,%d]", d->T,
              kMaxKeyFrames);
  VFS_REQUIRE(B >= 1 && B <= kMaxProblems && B * d->T <= kMaxKeySlots, VFS_ESHAPE,
              "masked_attention: %d problems x %d key frames exceeds the per-launch limit (%d, %d)", B, d->T,
              kMaxProblems, kMaxKeySlots);
  VFS_REQUIRE(d->topk >= 1 && d->topk <= 16, VFS_ESHAPE, "masked_attention: topk=%d outside [1,16]", d->topk);
  VFS_REQUIRE(d->temperature > 0.0f, VFS_EINVAL, "masked_attention: temperature must be positive");
  VFS_REQUIRE(d->mask_mode >= 0 && d->mask_mode <= 2, VFS_EINVAL, "masked_attention: bad mask_mode");
  VFS_REQUIRE(d->non_mask_len >= 0 && d->non_mask_len < d->T, VFS_EINVAL, "masked_attention: bad non_mask_len");
  VFS_REQUIRE(d->Cv >= 1, VFS_ESHAPE, "masked_attention: Cv must be >= 1");
  VFS_REQUIRE(q_bank_frames < 32768 && k_bank_frames < 32768, VFS_ESHAPE, "masked_attention: bank too large");
  VFS_REQUIRE(workspace_bytes >= attention_workspace_bytes(d, B), VFS_EINVAL, "masked_attention: workspace too small");
  for (int i = 0; i < B; ++i)
    VFS_REQUIRE(q_ids[i] >= 0 && q_ids[i] < q_bank_frames, VFS_EINVAL, "masked_attention: query frame id %d out of range",
                q_ids[i]);
  for (int i = 0; i < B * d->T; ++i) {
    VFS_REQUIRE(key_ids[i] >= 0 && key_ids[i] < k_bank_frames, VFS_EINVAL,
                "masked_attention: key frame id %d out of range", key_ids[i]);
    VFS_REQUIRE(val_ids[i] >= 0 && val_ids[i] < 32768, VFS_EINVAL, "masked_attention: value frame id out of range");
  }

  const int HW = d->H * d->W;
  const int KMAX = attn_kmax(d->topk);
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.H = d->H; p.W = d->W; p.kchunks = d->C / 64; p.T = d->T; p.num_problems = B;
  for (int i = 0; i < B; ++i) p.q_ids[i] = static_cast<short>(q_ids[i]);
  for (int i = 0; i < B * d->T; ++i) p.frame_ids[i] = static_cast<short>(key_ids[i]);
  p.mask_mode = d->mask_mode; p.ry = d->radius_y; p.rx = d->radius_x; p.non_mask_len = d->non_mask_len;
  p.q_tiles_x = (d->W + kQTileW - 1) / kQTileW;
  p.q_tiles_y = (d->H + kQTileH - 1) / kQTileH;
  p.splits = attn_splits(d, B);
  p.num_units = B * p.q_tiles_x * p.q_tiles_y * d->T * p.splits;
  const size_t slots = static_cast<size_t>(B) * d->T * p.splits * kAttnEpiHalves;
  p.part_val = reinterpret_cast<float*>(workspace);
  p.part_idx = reinterpret_cast<int*>(p.part_val + slots * KMAX * HW);
  const uint32_t box[5] = {64, kQTileW, kQTileH, 1, 2};
  {
    const uint64_t dims[5] = {static_cast<uint64_t>(d->C), static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H),
                              static_cast<uint64_t>(q_bank_frames), 2};
    const uint64_t strides[4] = {static_cast<uint64_t>(d->C) * 2, static_cast<uint64_t>(d->W) * d->C * 2,
                                 static_cast<uint64_t>(HW) * d->C * 2, static_cast<uint64_t>(q_plane_stride) * 2};
    int rc = make_tmap_16b_sw128(&p.tmap_q, q_bank_split, 5, dims, strides, box);
    if (rc != VFS_OK) return rc;
  }
  {
    const uint64_t dims[5] = {static_cast<uint64_t>(d->C), static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H),
                              static_cast<uint64_t>(k_bank_frames), 2};
    const uint64_t strides[4] = {static_cast<uint64_t>(d->C) * 2, static_cast<uint64_t>(d->W) * d->C * 2,
                                 static_cast<uint64_t>(HW) * d->C * 2, static_cast<uint64_t>(k_plane_stride) * 2};
    int rc = make_tmap_16b_sw128(&p.tmap_k, k_bank_split, 5, dims, strides, box);
    if (rc != VFS_OK) return rc;
  }
  const bool wide = attn_use_wide();
  p.kth = kQTileH;
  p.ktw = kQTileW;
  if (wide) {
    // key tile shape covering an interior query tile's window with the fewest tiles (then the widest rows)
    int wh = d->H, ww = d->W;
    if (d->mask_mode == 1) {
      wh = kQTileH + 2 * (d->radius_y - 1);
      ww = kQTileW + 2 * (d->radius_y - 1);
    } else if (d->mask_mode == 2) {
      wh = kQTileH + 2 * d->radius_y;
      ww = kQTileW + 2 * d->radius_x;
    }
    if (wh > d->H) wh = d->H;
    if (ww > d->W) ww = d->W;
    if (wh < 1) wh = 1;
    if (ww < 1) ww = 1;
    static int fixed = -1;   // VFS_ATTN_KEYTILE=0 keeps the 8 x 16 key tiles (comparison runs)
    if (fixed < 0) {
      const char* e = getenv("VFS_ATTN_KEYTILE");
      fixed = (e && atoi(e) == 0) ? 1 : 0;
    }
    if (!fixed) {
      // instantiated shapes: 8 x 16 (general), 5 x 25 (radius 18: 42 x 50 window -> 18 tiles instead of 24),
      // 6 x 19 (radius 12: 30 x 38 window -> 10 instead of 12), 4 x 32
      const int shapes[4][2] = {{8, 16}, {5, 25}, {6, 19}, {4, 32}};
      int best_tiles = 1 << 30;
      for (int k = 0; k < 4; ++k) {
        const int th = shapes[k][0], tw = shapes[k][1];
        const int tiles = ((wh + th - 1) / th) * ((ww + tw - 1) / tw);
        if (tiles < best_tiles) {
          best_tiles = tiles;
          p.kth = th;
          p.ktw = tw;
        }
      }
    }
  }
  p.kvalid = p.kth * p.ktw;
  p.stage_tx_bytes = 2 * 16384 + (wide ? 4 * p.kvalid * 128 : 2 * 16384);
  if (wide) {
    const uint32_t box1[5] = {64, static_cast<uint32_t>(p.ktw), static_cast<uint32_t>(p.kth), 1, 1};
    const uint64_t dims[5] = {static_cast<uint64_t>(d->C), static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H),
                              static_cast<uint64_t>(k_bank_frames), 2};
    const uint64_t strides[4] = {static_cast<uint64_t>(d->C) * 2, static_cast<uint64_t>(d->W) * d->C * 2,
                                 static_cast<uint64_t>(HW) * d->C * 2, static_cast<uint64_t>(k_plane_stride) * 2};
    int rc = make_tmap_16b_sw128(&p.tmap_k1, k_bank_split, 5, dims, strides, box1);
    if (rc != VFS_OK) return rc;
  }
  const int sms = device_sm_count();
  const int grid = p.num_units < sms ? p.num_units : sms;
  {
    const int shape_id = (p.kth == 5) ? 1 : (p.kth == 6) ? 2 : (p.kth == 4) ? 3 : 0;
    int rc = VFS_OK;
#define VFS_ATTN_LAUNCH(K, WD, TH, TW)                                                                          \
  do {                                                                                                            \
    static bool cfgd = false;                                                                                     \
    if (!cfgd) {                                                                                                  \
      rc = check_cuda(cudaFuncSetAttribute(attn_scores_topk_kernel<K, WD, TH, TW>,                               \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes),          \
                      "cudaFuncSetAttribute(attn_scores_topk_kernel)");                                           \
      cfgd = (rc == VFS_OK);                                                                                      \
    }                                                                                                             \
    if (rc == VFS_OK) attn_scores_topk_kernel<K, WD, TH, TW><<<grid, kAttnThreads, kAttnSmemBytes, stream>>>(p); \
  } while (0)
    if (!wide) {
      if (KMAX == 10) VFS_ATTN_LAUNCH(10, false, 8, 16);
      else VFS_ATTN_LAUNCH(16, false, 8, 16);
    } else if (KMAX == 10) {
      switch (shape_id) {
        case 1: VFS_ATTN_LAUNCH(10, true, 5, 25); break;
        case 2: VFS_ATTN_LAUNCH(10, true, 6, 19); break;
        case 3: VFS_ATTN_LAUNCH(10, true, 4, 32); break;
        default: VFS_ATTN_LAUNCH(10, true, 8, 16); break;
      }
    } else {
      switch (shape_id) {
        case 1: VFS_ATTN_LAUNCH(16, true, 5, 25); break;
        case 2: VFS_ATTN_LAUNCH(16, true, 6, 19); break;
        case 3: VFS_ATTN_LAUNCH(16, true, 4, 32); break;
        default: VFS_ATTN_LAUNCH(16, true, 8, 16); break;
      }
    }
#undef VFS_ATTN_LAUNCH
    if (rc != VFS_OK) return rc;
  }
  VFS_CUDA_OK(cudaGetLastError());

  MergeParams m;
  memset(&m, 0, sizeof(m));
  m.part_val = p.part_val; m.part_idx = p.part_idx;
  m.slots = d->T * p.splits * kAttnEpiHalves; m.HW = HW; m.topk = d->topk; m.mode = d->mode; m.temperature = d->temperature;
  m.values = values; m.v_batch_stride = v_batch_stride; m.v_frame_stride = v_frame_stride;
  m.v_chan_stride = v_chan_stride; m.Cv = d->Cv; m.T = d->T;
  for (int i = 0; i < B * d->T; ++i) m.val_ids[i] = static_cast<short>(val_ids[i]);
  m.out = out; m.out_val = out_topk_val; m.out_idx = out_topk_idx;
  const dim3 mgrid((HW + 31) / 32, B);
  if (KMAX == 10) attn_merge_propagate_kernel<10><<<mgrid, 32 * kMergeSub, 0, stream>>>(m);
  else attn_merge_propagate_kernel<16><<<mgrid, 32 * kMergeSub, 0, stream>>>(m);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int masked_attention(const VfsAttnDesc* d, const void* q_split, long long q_plane_stride, const void* k_bank_split,
                     long long k_plane_stride, int k_bank_frames, const int* key_frame_ids, const float* values,
                     long long v_frame_stride, long long v_chan_stride, float* out, float* out_topk_val,
                     int* out_topk_idx, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const int q_id = 0;
  return masked_attention_batched(d, 1, q_split, q_plane_stride, 1, &q_id, k_bank_split, k_plane_stride,
                                  k_bank_frames, key_frame_ids, values, key_frame_ids, 0, v_frame_stride,
                                  v_chan_stride, out, out_topk_val, out_topk_idx, workspace, workspace_bytes, stream);
}

VFS_DEFINE_OVERFLOW_ACCESSOR(overflow_affinity)

}  // namespace vfs
