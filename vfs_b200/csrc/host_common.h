// Host-side helpers shared by all translation units of libvfs_b200.so: error reporting for the
// C ABI, and TMA tensor-map construction through the driver entry point (no link-time libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/vfs_b200.h"

namespace vfs {

void set_last_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);  // returns VFS_OK or VFS_ECUDA (and records the message)

#define VFS_CUDA_OK(expr)                                 \
  do {                                                    \
    int _rc = ::vfs::check_cuda((expr), #expr);           \
    if (_rc != VFS_OK) return _rc;                        \
  } while (0)

#define VFS_REQUIRE(cond, code, ...)                      \
  do {                                                    \
    if (!(cond)) {                                        \
      ::vfs::set_last_error(__VA_ARGS__);                 \
      return (code);                                      \
    }                                                     \
  } while (0)

// Encodes a bf16 tiled tensor map with 128-byte swizzle.  dims/box innermost-first; strides_bytes has
// rank-1 entries (stride of dim 1.. rank-1).  Returns VFS_OK or an error code.
int make_tmap_16b_sw128(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box);

int device_sm_count();

// Programmatic dependent launch for kernels that start with pdl_wait() (ptx.cuh): the grid may be scheduled while its
// predecessor in the stream drains and waits on the device for the predecessor's completion, which hides the ~1 us
// launch latency between dependent kernels (also as programmatic edges inside a captured CUDA graph -- the training
// step is ~900 dependent launches).  VFS_PDL=0 turns the attribute off.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace vfs
