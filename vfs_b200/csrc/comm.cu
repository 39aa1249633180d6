// Peer-memory communicator for the data-parallel training step: one process per GPU of ONE NVLink / NVSwitch box.
//
// The reference trains with MMDistributedDataParallel (mmaction/apis/train.py:58-66: bucketed NCCL all-reduce of the
// gradients) and torch.nn.SyncBatchNorm (configs/*:9,15: one NCCL exchange of the batch statistics per BN layer,
// forward and backward).  At B200 step times those ~230 latency-bound collectives per step, each issued from the host,
// cost more than the arithmetic between them.  Here every rank owns a *symmetric segment* (cudaMalloc + CUDA IPC,
// mapped into every peer), and the exchanges are ordinary kernels on the caller's stream that store into / load from
// peer memory over NVLink:
//
//   comm_allreduce_small   [<= 4096 values, fp64 or fp32]  SyncBN statistics / logged scalars.  Every rank PUSHES its
//                          values into a slot of every peer's segment as epoch-tagged 8-byte words (NCCL's LL
//                          protocol: no fence, no separate flag) and polls the cells of its own segment; the sum is
//                          formed in rank order from local memory, so all ranks hold bit-identical results.
//   comm_barrier           flags only.
//   comm_allreduce_f32     the gradient all-reduce over a range of the segment's data region: barrier, then rank r
//                          reduces the r-th chunk reading the W peer copies (fixed order), scales it and writes the
//                          result into all W copies, barrier.  Two-shot: every byte crosses NVLink once in, once out.
//
// Nothing here involves the host after setup, so the whole multi-rank step is capturable in a CUDA graph.
// Every spin-wait has a wall-clock watchdog (globaltimer): on expiry it sets a sticky error word in the local segment
// (read by comm_error) and all later waits fall through -- a protocol bug or a dead peer must not hang the box.
#include <stdlib.h>

#include "host_common.h"

namespace vfs {

namespace {

constexpr int kMaxWorld = 8;
constexpr int kSlots = 32;                       // small-exchange slots, used round-robin by the host
constexpr int kSmallBytes = 64 * 1024;           // per (slot, parity, rank): 4096 cells of 16 bytes
constexpr unsigned long long kDefaultTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

// control block at the start of every segment
struct Control {
  unsigned int small_flag[kSlots][kMaxWorld];    // written by peer r: epoch of its push into this slot
  unsigned int small_epoch[kSlots];              // local: last completed epoch of the slot
  unsigned int barrier_flag[kMaxWorld];          // written by peer r
  unsigned int barrier_epoch;                    // local
  unsigned int error;                            // sticky watchdog flag (local)
  unsigned int pad[7];
};

constexpr size_t kControlBytes = (sizeof(Control) + 1023) / 1024 * 1024;
inline size_t small_region_bytes(int world) { return static_cast<size_t>(kSlots) * 2 * world * kSmallBytes; }

struct CommDev {   // passed by value to the kernels
  int rank, world;
  unsigned long long timeout_ns;
  char* seg[kMaxWorld];                          // base of every rank's segment as mapped in THIS process
};

__device__ __forceinline__ Control* ctl(const CommDev& c, int r) { return reinterpret_cast<Control*>(c.seg[r]); }
__device__ __forceinline__ char* small_buf(const CommDev& c, int r, int slot, int parity, int src) {
  return c.seg[r] + kControlBytes +
         ((static_cast<size_t>(slot) * 2 + parity) * c.world + src) * static_cast<size_t>(kSmallBytes);
}

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// wait until *flag has reached `epoch` (wrap-safe), with the watchdog
__device__ __forceinline__ void wait_flag(const CommDev& c, const unsigned int* flag, unsigned int epoch) {
  Control* me = ctl(c, c.rank);
  if (static_cast<int>(ld_acquire_sys(flag) - epoch) >= 0) return;
  if (*reinterpret_cast<volatile unsigned int*>(&me->error)) return;
  const unsigned long long t0 = global_ns();
  while (static_cast<int>(ld_acquire_sys(flag) - epoch) < 0) {
    if (global_ns() - t0 > c.timeout_ns) {
      *reinterpret_cast<volatile unsigned int*>(&me->error) = 1u;
      __threadfence_system();
      return;
    }
    __nanosleep(64);
  }
}

// LL ("low latency") exchange, the protocol NCCL uses for small messages: every 4-byte half of a value travels in an
// 8-byte word {payload, epoch} -- 8-byte stores are single transactions over NVLink, so the receiver needs no separate
// flag and the sender no system-scope fence: it polls each word until its epoch tag matches.  One one-way NVLink trip
// instead of store -> fence (round trip) -> flag store -> flag poll.  A slot holds [parity][rank][n] 16-byte cells
// {lo32, epoch, hi32, epoch} (fp32 values use the first word only).
__device__ __forceinline__ void st_ll(uint4* p, unsigned int a, unsigned int b, unsigned int epoch) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(epoch), "r"(b), "r"(epoch)
               : "memory");
}
__device__ __forceinline__ uint4 ld_ll(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

template <typename T>
__global__ void __launch_bounds__(1024) allreduce_small_kernel(const CommDev c, T* __restrict__ data, int n, int slot) {
  Control* me = ctl(c, c.rank);
  const unsigned int epoch = me->small_epoch[slot] + 1u;   // every thread reads it before thread 0 advances it
  const int parity = static_cast<int>(epoch & 1u);
  // push my values into every rank's copy of the slot (own copy included), tagged with the epoch
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    unsigned int lo, hi = 0u;
    if (sizeof(T) == 8) {
      const unsigned long long bits = __double_as_longlong(static_cast<double>(data[i]));
      lo = static_cast<unsigned int>(bits);
      hi = static_cast<unsigned int>(bits >> 32);
    } else {
      lo = __float_as_uint(static_cast<float>(data[i]));
    }
    for (int r = 0; r < c.world; ++r)
      st_ll(reinterpret_cast<uint4*>(small_buf(c, r, slot, parity, c.rank)) + i, lo, hi, epoch);
  }
  // collect: poll every rank's cell until both tags carry this epoch, sum in rank order (bit-identical on all ranks)
  const unsigned long long t0 = global_ns();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    T sum = static_cast<T>(0);
    for (int r = 0; r < c.world; ++r) {
      const uint4* cell = reinterpret_cast<const uint4*>(small_buf(c, c.rank, slot, parity, r)) + i;
      uint4 v = ld_ll(cell);
      while (v.y != epoch || v.w != epoch) {
        if (*reinterpret_cast<volatile unsigned int*>(&me->error)) break;
        if (global_ns() - t0 > c.timeout_ns) {
          *reinterpret_cast<volatile unsigned int*>(&me->error) = 1u;
          break;
        }
        v = ld_ll(cell);
      }
      if (sizeof(T) == 8)
        sum += static_cast<T>(__longlong_as_double(static_cast<long long>((static_cast<unsigned long long>(v.z) << 32) | v.x)));
      else
        sum += static_cast<T>(__uint_as_float(v.x));
    }
    data[i] = sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) me->small_epoch[slot] = epoch;
}

__global__ void barrier_kernel(const CommDev c) {
  Control* me = ctl(c, c.rank);
  const unsigned int epoch = me->barrier_epoch + 1u;
  __threadfence_system();   // everything this device wrote before (earlier kernels included) precedes the flag
  __syncthreads();
  if (threadIdx.x < c.world) {
    st_release_sys(&ctl(c, threadIdx.x)->barrier_flag[c.rank], epoch);
    wait_flag(c, &me->barrier_flag[threadIdx.x], epoch);
  }
  __syncthreads();
  if (threadIdx.x == 0) me->barrier_epoch = epoch;
}

__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

// rank r owns float4 range [r*chunk4, min((r+1)*chunk4, n4)) of the buffer at byte offset `off` of every segment
template <int W>
__global__ void __launch_bounds__(512) allreduce_chunk_kernel(const CommDev c, size_t off, size_t n4, size_t chunk4,
                                                              float scale) {
  const size_t begin = static_cast<size_t>(c.rank) * chunk4;
  size_t end = begin + chunk4;
  if (end > n4) end = n4;
  float* base[W];
#pragma unroll
  for (int r = 0; r < W; ++r) base[r] = reinterpret_cast<float*>(c.seg[r] + off);
  for (size_t i = begin + blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < end;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 v[W];
#pragma unroll
    for (int r = 0; r < W; ++r) v[r] = ld_peer_f4(base[r] + 4 * i);
    float4 s = v[0];
#pragma unroll
    for (int r = 1; r < W; ++r) {
      s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w;
    }
    s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
#pragma unroll
    for (int r = 0; r < W; ++r) *reinterpret_cast<float4*>(base[r] + 4 * i) = s;
  }
}

}  // namespace

struct Comm {
  CommDev dev;
  void* local = nullptr;          // this rank's segment (cudaMalloc)
  size_t seg_bytes = 0, data_bytes = 0, data_off = 0;
  bool connected = false;
  int next_slot = 0;
};

size_t comm_handle_bytes() { return sizeof(cudaIpcMemHandle_t); }

int comm_create(int rank, int world, size_t data_bytes, Comm** out, void* handle_out) {
  VFS_REQUIRE(out && handle_out, VFS_EINVAL, "comm_create: null argument");
  VFS_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, VFS_EINVAL,
              "comm_create: rank %d / world %d unsupported (1..%d ranks of one box)", rank, world, kMaxWorld);
  Comm* c = new Comm();
  memset(&c->dev, 0, sizeof(c->dev));
  c->dev.rank = rank;
  c->dev.world = world;
  c->dev.timeout_ns = kDefaultTimeoutNs;
  if (const char* env = getenv("VFS_COMM_TIMEOUT_MS")) {
    const long long ms = atoll(env);
    if (ms > 0) c->dev.timeout_ns = static_cast<unsigned long long>(ms) * 1000000ull;
  }
  c->data_off = (kControlBytes + small_region_bytes(world) + 4095) / 4096 * 4096;
  c->data_bytes = (data_bytes + 4095) / 4096 * 4096;
  c->seg_bytes = c->data_off + c->data_bytes;
  int rc = check_cuda(cudaMalloc(&c->local, c->seg_bytes), "cudaMalloc(symmetric segment)");
  if (rc != VFS_OK) { delete c; return rc; }
  rc = check_cuda(cudaMemset(c->local, 0, c->seg_bytes), "cudaMemset(symmetric segment)");
  if (rc == VFS_OK)
    rc = check_cuda(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle_out), c->local),
                    "cudaIpcGetMemHandle");
  if (rc != VFS_OK) { cudaFree(c->local); delete c; return rc; }
  c->dev.seg[rank] = reinterpret_cast<char*>(c->local);
  if (world == 1) c->connected = true;
  *out = c;
  return VFS_OK;
}

int comm_connect(Comm* c, const void* all_handles) {
  VFS_REQUIRE(c && all_handles, VFS_EINVAL, "comm_connect: null argument");
  const cudaIpcMemHandle_t* h = reinterpret_cast<const cudaIpcMemHandle_t*>(all_handles);
  for (int r = 0; r < c->dev.world; ++r) {
    if (r == c->dev.rank) continue;
    void* p = nullptr;
    VFS_CUDA_OK(cudaIpcOpenMemHandle(&p, h[r], cudaIpcMemLazyEnablePeerAccess));
    c->dev.seg[r] = reinterpret_cast<char*>(p);
  }
  c->connected = true;
  return VFS_OK;
}

int comm_destroy(Comm* c) {
  if (!c) return VFS_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < c->dev.world; ++r)
    if (r != c->dev.rank && c->dev.seg[r]) cudaIpcCloseMemHandle(c->dev.seg[r]);
  if (c->local) cudaFree(c->local);
  delete c;
  return VFS_OK;
}

void* comm_data_ptr(Comm* c) { return c ? reinterpret_cast<char*>(c->local) + c->data_off : nullptr; }
size_t comm_data_bytes(Comm* c) { return c ? c->data_bytes : 0; }

int comm_error(Comm* c) {
  if (!c) return 0;
  unsigned int e = 0;
  cudaMemcpy(&e, &reinterpret_cast<Control*>(c->local)->error, sizeof(e), cudaMemcpyDeviceToHost);
  return static_cast<int>(e);
}

int comm_allreduce_small(Comm* c, void* data, int n, int is_f64, cudaStream_t s) {
  VFS_REQUIRE(c && data, VFS_EINVAL, "comm_allreduce_small: null argument");
  VFS_REQUIRE(c->connected, VFS_EINVAL, "comm_allreduce_small: communicator is not connected");
  VFS_REQUIRE(n > 0 && static_cast<size_t>(n) * 16 <= static_cast<size_t>(kSmallBytes), VFS_ESHAPE,
              "comm_allreduce_small: %d values exceed the %d cells of a slot", n, kSmallBytes / 16);
  const int slot = c->next_slot;
  c->next_slot = (c->next_slot + 1) % kSlots;
  int threads = (n + 31) / 32 * 32;
  if (threads > 1024) threads = 1024;
  if (threads < 32) threads = 32;
  if (is_f64)
    allreduce_small_kernel<double><<<1, threads, 0, s>>>(c->dev, reinterpret_cast<double*>(data), n, slot);
  else
    allreduce_small_kernel<float><<<1, threads, 0, s>>>(c->dev, reinterpret_cast<float*>(data), n, slot);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int comm_barrier(Comm* c, cudaStream_t s) {
  VFS_REQUIRE(c && c->connected, VFS_EINVAL, "comm_barrier: communicator is not connected");
  barrier_kernel<<<1, 32, 0, s>>>(c->dev);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

int comm_allreduce_f32(Comm* c, size_t offset_bytes, size_t n, float scale, cudaStream_t s) {
  VFS_REQUIRE(c && c->connected, VFS_EINVAL, "comm_allreduce_f32: communicator is not connected");
  VFS_REQUIRE(offset_bytes % 16 == 0 && n % 4 == 0 && offset_bytes + n * 4 <= c->data_bytes, VFS_ESHAPE,
              "comm_allreduce_f32: range [%zu, +%zu floats) must be 16-byte aligned, a multiple of 4 floats and inside "
              "the %zu-byte data region", offset_bytes, n, c->data_bytes);
  if (n == 0) return VFS_OK;
  const int W = c->dev.world;
  const size_t n4 = n / 4;
  const size_t chunk4 = (n4 + W - 1) / W;
  const size_t off = c->data_off + offset_bytes;
  VFS_CUDA_OK(cudaGetLastError());
  barrier_kernel<<<1, 32, 0, s>>>(c->dev);
  size_t blocks = (chunk4 + 511) / 512;
  const size_t cap = static_cast<size_t>(device_sm_count()) * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const int g = static_cast<int>(blocks);
  switch (W) {
    case 1: allreduce_chunk_kernel<1><<<g, 512, 0, s>>>(c->dev, off, n4, chunk4, scale); break;
    case 2: allreduce_chunk_kernel<2><<<g, 512, 0, s>>>(c->dev, off, n4, chunk4, scale); break;
    case 3: allreduce_chunk_kernel<3><<<g, 512, 0, s>>>(c->dev, off, n4, chunk4, scale); break;
    case 4: allreduce_chunk_kernel<4><<<g, 512, 0, s>>>(c->dev, off, n4, chunk4, scale); break;
    case 5: allreduce_chunk_kernel<5><<<g, 512, 0, s>>>(c->dev, off, n4, chunk4, scale); break;
    case 6: allreduce_chunk_kernel<6><<<g, 512, 0, s>>>(c->dev, off, n4, chunk4, scale); break;
    case 7: allreduce_chunk_kernel<7><<<g, 512, 0, s>>>(c->dev, off, n4, chunk4, scale); break;
    default: allreduce_chunk_kernel<8><<<g, 512, 0, s>>>(c->dev, off, n4, chunk4, scale); break;
  }
  VFS_CUDA_OK(cudaGetLastError());
  barrier_kernel<<<1, 32, 0, s>>>(c->dev);
  VFS_CUDA_OK(cudaGetLastError());
  return VFS_OK;
}

}  // namespace vfs
