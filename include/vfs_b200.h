/*
 * libvfs_b200 -- C ABI of the B200-native (sm_100a) hot path of xvjiarui/VFS.
 *
 * The reference has no FFI of its own (it is pure Python on top of torch/mmcv); every entry point
 * below replaces a *library call site* of the reference and cites it.  Conventions:
 *   - plain device pointers + sizes; no torch types; no allocation, no synchronisation, no host
 *     callbacks inside; everything is enqueued on the caller's stream.
 *   - return value: VFS_OK (0) or a negative error code; vfs_last_error_string() describes the last
 *     failure of the calling thread.  Nothing throws across the boundary.
 *   - "split" tensors are the library's activation format: fp32 values stored as two bf16 planes
 *     (hi, lo with x ~= hi + lo, relative error <= 2^-16), layout [2][N][H][W][C] (plane-major NHWC).
 *     They feed the tcgen05 tensor cores three products at a time (hi*hi + hi*lo + lo*hi), which
 *     keeps fp32-level accuracy (the reference computes in fp32; parity bar 1e-3).
 */
#ifndef VFS_B200_H_
#define VFS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VFS_OK 0
#define VFS_EINVAL (-1) /* bad argument (null pointer, unsupported value) */
#define VFS_ESHAPE (-2) /* shape not supported by the kernel family */
#define VFS_EARCH (-3)  /* device is not sm_100 */
#define VFS_ECUDA (-4)  /* CUDA runtime / driver error, see vfs_last_error_string() */

typedef struct CUstream_st* vfs_stream_t; /* == cudaStream_t */

const char* vfs_last_error_string(void);
int vfs_abi_version(void);
/* VFS_OK iff the current device is compute capability 10.x. */
int vfs_check_device(void);

/* ------------------------------------------------------------------------------------------------
 * Layout boundary.  Reference tensors are NCHW fp32 contiguous.
 * ---------------------------------------------------------------------------------------------- */
/* NCHW fp32 -> split NHWC.  C must be a multiple of 8. */
int vfs_nchw_f32_to_split(const float* in, void* out_split, int N, int C, int H, int W, vfs_stream_t s);
/* split NHWC -> NCHW fp32 (hi + lo). */
int vfs_split_to_nchw_f32(const void* in_split, float* out, int N, int C, int H, int W, vfs_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * ResNet stem: conv 7x7/s2/p3 (3->64) + BN(eval, folded scale/shift) + ReLU + maxpool 3x3/s2/p1.
 * Replaces ResNet.forward's `self.conv1(x); self.maxpool(x)`
 * (mmaction/models/backbones/resnet.py:565-566, _make_stem_layer :422-435).
 *   in        NCHW fp32 [N,3,H,W]           weight [64,3,7,7] fp32
 *   scale/shift [64] fp32 (gamma/sqrt(var+eps), beta - mean*scale)
 *   out_split split NHWC [N, Hp, Wp, 64], Hc = (H+6-7)/2+1, Hp = (Hc+2-3)/2+1
 *   workspace fp32 [N*Hc*Wc*64] (vfs_stem_workspace_bytes)
 * ---------------------------------------------------------------------------------------------- */
size_t vfs_stem_workspace_bytes(int N, int H, int W);
int vfs_stem_forward(const float* in, const float* weight, const float* scale, const float* shift, void* out_split,
                     void* workspace, int N, int H, int W, vfs_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * conv(k x k, stride, dilation, pad = dilation*(k/2), bias=False) -> y*scale[c]+shift[c] (+residual) (ReLU)
 * as one tcgen05 implicit-GEMM kernel.  Replaces one mmcv ConvModule (conv -> BN(eval) -> ReLU) plus,
 * for the last conv of a block, `out += identity; relu(out)`:
 *   Bottleneck.forward resnet.py:200-232, BasicBlock.forward :83-113, downsample :267-277.
 *   in_split   split NHWC [N,H,W,Cin]           Cin, Cout multiples of 64
 *   w_split    split [2][Cout][k*k*Cin], K index = (r*k + s)*Cin + c   (vfs_pack_conv_weight)
 *   residual_split  NULL or split NHWC [N,Ho,Wo,Cout]
 *   out_split  NULL or split NHWC [N,Ho,Wo,Cout];  out_f32_nhwc  NULL or fp32 [N,Ho,Wo,Cout]
 * ---------------------------------------------------------------------------------------------- */
typedef struct VfsConvDesc {
  int32_t N, H, W, Cin, Cout;
  int32_t ksize;    /* 1 or 3 */
  int32_t stride;   /* 1 or 2 */
  int32_t dilation; /* >= 1 (3x3 only) */
  int32_t relu;     /* 0/1 */
} VfsConvDesc;

int vfs_conv_bn_act(const VfsConvDesc* d, const void* in_split, const void* w_split, const float* scale,
                    const float* shift, const void* residual_split, void* out_split, float* out_f32_nhwc,
                    vfs_stream_t s);

/* OIHW fp32 [Cout,Cin,k,k] -> split [2][Cout][k*k*Cin] (device to device). */
int vfs_pack_conv_weight(const float* w_oihw, void* w_split, int Cout, int Cin, int ksize, vfs_stream_t s);

/* Test instrument: the same contract as vfs_conv_bn_act computed with plain fp32 FMAs (one thread per
 * output element, fixed summation order).  Exists so the tensor-core path can be checked on the device
 * at sizes the CPU oracle cannot reach; never called by the product path. */
int vfs_debug_conv_bn_act_simt(const VfsConvDesc* d, const void* in_split, const void* w_split, const float* scale,
                               const float* shift, const void* residual_split, float* out_f32_nhwc,
                               vfs_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* VFS_B200_H_ */
