/*
 * libvfs_b200 -- C ABI of the B200-native (sm_100a) hot path of xvjiarui/VFS.
 *
 * The reference has no FFI of its own (it is pure Python on top of torch/mmcv); every entry point
 * below replaces a *library call site* of the reference and cites it.  Conventions:
 *   - plain device pointers + sizes; no torch types; no allocation, no synchronisation, no host
 *     callbacks inside; everything is enqueued on the caller's stream.
 *   - return value: VFS_OK (0) or a negative error code; vfs_last_error_string() describes the last
 *     failure of the calling thread.  Nothing throws across the boundary.
 *   - "split" tensors are the library's activation format: fp32 values stored as two IEEE-half planes
 *     (hi, lo with x ~= hi + lo: 22 significant bits, relative error <= 2^-22 for |x| >~ 0.1, absolute error
 *     <= 3e-8 below), layout [2][N][H][W][C] (plane-major NHWC).  They feed the tcgen05 tensor cores three
 *     products at a time (lo*hi + hi*lo + hi*hi, fp32 accumulation), which keeps fp32-class accuracy (the
 *     reference computes in fp32; parity bar 1e-3, exact top-k indices).  Values must stay inside the fp16 range
 *     (|x| <= 65504); violations are counted, see vfs_overflow_count().
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
/* Values that left the fp16 range while being split since the last reset (synchronises the device). */
unsigned int vfs_overflow_count(int reset);

/* ------------------------------------------------------------------------------------------------
 * Layout boundary.  Reference tensors are NCHW fp32 contiguous.
 * ---------------------------------------------------------------------------------------------- */
/* NCHW fp32 -> split NHWC.  C must be a multiple of 8. */
int vfs_nchw_f32_to_split(const float* in, void* out_split, int N, int C, int H, int W, vfs_stream_t s);
/* same, of scale * in (used to enter the backward pass with power-of-two scaled gradients). */
int vfs_nchw_f32_to_split_scaled(const float* in, void* out_split, int N, int C, int H, int W, float scale,
                                 vfs_stream_t s);
/* split NHWC -> NCHW fp32 (hi + lo). */
int vfs_split_to_nchw_f32(const void* in_split, float* out, int N, int C, int H, int W, vfs_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * ResNet stem: conv 7x7/s2/p3 (3->64) + BN(eval, folded scale/shift) + ReLU + maxpool 3x3/s2/p1.
 * Replaces ResNet.forward's `self.conv1(x); self.maxpool(x)`
 * (mmaction/models/backbones/resnet.py:565-566, _make_stem_layer :422-435).
 *   in        NCHW fp32 [N,3,H,W]           weight: split [2][64][192] packed by vfs_stem_pack_weight from the
 *             OIHW fp32 [64,3,7,7] filter bank (K = 147 zero-padded to 192; the conv runs on tcgen05 with the im2col
 *             tile built in shared memory)
 *   scale/shift [64] fp32 (gamma/sqrt(var+eps), beta - mean*scale)
 *   out_split split NHWC [N, Hp, Wp, 64], Hc = (H+6-7)/2+1, Hp = (Hc+2-3)/2+1
 *   workspace fp32 [N*Hc*Wc*64] (vfs_stem_workspace_bytes)
 * ---------------------------------------------------------------------------------------------- */
size_t vfs_stem_workspace_bytes(int N, int H, int W);
size_t vfs_stem_packed_weight_bytes(void);
int vfs_stem_pack_weight(const float* w_oihw, void* w_split, vfs_stream_t s);
int vfs_stem_forward(const float* in, const void* weight, const float* scale, const float* shift, void* out_split,
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

/* ------------------------------------------------------------------------------------------------
 * Train-mode BatchNorm (batch statistics), reference: mmcv ConvModule with norm_cfg SyncBN/BN in training mode
 * (torch.nn.SyncBatchNorm / BatchNorm2d forward: batch_norm_stats, gather_stats, batch_norm_elemt).
 *   1. vfs_conv_stats     conv as above (scale/shift normally 1/0) writing the raw fp32 NHWC output AND accumulating
 *                         per-channel [sum | sum of squares] into stats (fp64 [2*Cout], zero-initialised by the caller;
 *                         across GPUs the caller all-reduces stats -- that is the SyncBN exchange)
 *   2. vfs_bn_finalize    stats,count -> scale = gamma*invstd, shift = beta - mean*scale, save_mean/save_invstd,
 *                         running_mean/var momentum update (unbiased variance), like F.batch_norm(training=True)
 *   3. vfs_bn_apply       y = z*scale + shift (+residual) (ReLU) -> split NHWC
 * vfs_channel_stats_f32 accumulates the same statistics for an fp32 [M,C] tensor (stem output).
 * ---------------------------------------------------------------------------------------------- */
/* stem in train mode: raw 7x7/s2 conv output (fp32 NHWC [N,Hc,Wc,64], vfs_stem_workspace_bytes) for the batch
 * statistics, then BN(scale/shift from vfs_bn_finalize)+ReLU fused into the 3x3/s2 max-pool -> split NHWC */
int vfs_stem_conv_raw(const float* in, const void* weight, void* conv_out_f32_nhwc, int N, int H, int W,
                      vfs_stream_t s);
int vfs_stem_bn_relu_pool(const void* conv_out_f32_nhwc, const float* scale, const float* shift, void* out_split,
                          int N, int H, int W, vfs_stream_t s);
int vfs_conv_stats(const VfsConvDesc* d, const void* in_split, const void* w_split, const float* scale,
                   const float* shift, float* out_f32_nhwc, double* stats, vfs_stream_t s);
/* same with the raw conv output kept as a SPLIT tensor (z_split, split NHWC [N,Ho,Wo,Cout]) written by the TMA
 * epilogue, the statistics accumulated by its math warps (column sums by warp shuffles): the fast train-mode forward.
 * ones / zeros: fp32 [Cout] constant vectors. */
int vfs_conv_stats_split(const VfsConvDesc* d, const void* in_split, const void* w_split, const float* ones,
                         const float* zeros, void* z_split, double* stats, vfs_stream_t s);
int vfs_channel_stats_f32(const float* x, double* stats, long long M, int C, vfs_stream_t s);
int vfs_bn_finalize(double* stats, double count, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float momentum, float eps, float* scale, float* shift, float* save_mean,
                    float* save_invstd, int C, vfs_stream_t s);
/* y = y*scale[c] + shift[c] (+ReLU) in place on fp32 [M,C]: the apply step of the two-phase (SyncBN) BatchNorm1d */
int vfs_affine_act_f32(float* y, const float* scale, const float* shift, long long M, int C, int relu,
                       vfs_stream_t s);
/* z (fp32 [M,C]) or, when z is NULL, z_split (split [2][M][C]) is the raw conv output */
int vfs_bn_apply(const float* z, const void* z_split, const float* scale, const float* shift,
                 const void* residual_split, void* out_split, long long M, int C, int relu, vfs_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * Backward of the convolution w.r.t. its input (training): dX = conv_transpose(dZ, W) (+ add), the same tcgen05
 * implicit-GEMM kernel with the roles of Cin/Cout swapped and a flipped kernel; stride-2 layers are evaluated per
 * output-parity class.  Replaces torch autograd's cudnn_convolution_backward_input for the ConvModules of
 * resnet.py.  `d` is the FORWARD descriptor (N,H,W = forward input extent).
 *   dz_split  split NHWC [N,Ho,Wo,Cout]     wt_split  split [2][Cin][k*k*Cout] (vfs_pack_conv_weight_dgrad)
 *   ones/zeros fp32 [Cin] constant vectors  add_split NULL or split NHWC [N,H,W,Cin] (gradient of another branch)
 *   dx_split  split NHWC [N,H,W,Cin]
 * ---------------------------------------------------------------------------------------------- */
int vfs_conv_dgrad(const VfsConvDesc* d, const void* dz_split, const void* wt_split, const float* ones,
                   const float* zeros, const void* add_split, void* dx_split, vfs_stream_t s);
int vfs_pack_conv_weight_dgrad(const float* w_oihw, void* wt_split, int Cout, int Cin, int ksize, vfs_stream_t s);

/* Backward of the convolution w.r.t. its weight: dW[co,ci,r,q] = sum_pixels dZ[pix,co] * X[pix + off(r,q), ci] as a
 * tcgen05 GEMM over the pixel axis with MN-major split-fp16 operands (csrc/wgrad_tc.cu).  Replaces torch autograd's
 * cudnn_convolution_backward_weight.  `d` is the FORWARD descriptor.
 *   x_split split NHWC [N,H,W,Cin]; dz_split split NHWC [N,Ho,Wo,Cout]; workspace vfs_conv_wgrad_workspace_bytes;
 *   dw_oihw fp32 [Cout,Cin,k,k]: overwritten, or accumulated into when accumulate != 0 (second SimSiam view). */
size_t vfs_conv_wgrad_workspace_bytes(int Cout, int Cin, int ksize);
int vfs_conv_wgrad(const VfsConvDesc* d, const void* x_split, const void* dz_split, void* workspace, float* dw_oihw,
                   int accumulate, float out_scale, vfs_stream_t s);

/* BatchNorm(+ReLU) backward around the conv gradients (torch batch_norm_backward_reduce / _elemt):
 *   g = dY * 1[y > 0];  sums = [sum g | sum g*xhat] (fp64, caller zero-initialises; all-reduced across ranks for
 *   SyncBN);  dz = gamma*invstd*(g - sums0/count - xhat*sums1/count);  dgamma = param_scale*sums1, dbeta = param_scale*sums0.
 * dY comes as a split tensor or as fp32 (stem); dz goes out split (residual stages) and/or fp32 (stem); g_split
 * (optional) is the masked gradient that also flows into the block's identity branch. */
int vfs_bn_bwd_reduce(const void* dy_split, const float* dy_f32, const void* y_split, const float* y_f32,
                      const float* z, const void* z_split, const float* mean, const float* invstd, double* sums,
                      long long M, int C, vfs_stream_t s);
int vfs_bn_bwd_apply(const void* dy_split, const float* dy_f32, const void* y_split, const float* y_f32,
                     const float* z, const void* z_split, const float* mean, const float* invstd, const float* gamma,
                     const double* sums, double count, void* dz_split, float* dz_f32, void* g_split, float* dgamma,
                     float* dbeta, int accumulate, float param_scale, long long M, int C, vfs_stream_t s);
int vfs_relu_bwd_split(const void* dy_split, const void* y_split, void* g_split, long long elems, vfs_stream_t s);
/* stem backward: max-pool(3,2,1)+ReLU backward onto the raw conv output grid (g fp32 [N,Hc,Wc,64]), and the 7x7
 * weight gradient dw[64,3,7,7] (+)= out_scale * sum dz * x */
int vfs_stem_pool_relu_bwd(const void* dpool_split, const float* z, const float* scale, const float* shift, float* g,
                           int N, int H, int W, vfs_stream_t s);
int vfs_stem_wgrad(const float* x, const float* dz, float* dw, int accumulate, float out_scale, int N, int H, int W,
                   vfs_stream_t s);
/* SimSiam head / loss backward and the optimiser update */
int vfs_linear_backward(const float* dy, const float* x, const float* W, float* dx, float* dW, float* db, int M,
                        int N, int K, int accumulate, vfs_stream_t s);
int vfs_bn1d_backward(const float* dy, const float* pre, const float* out, float* dpre, int M, int N,
                      const float* gamma, const float* mean, const float* invstd, int training, int relu,
                      float* dgamma, float* dbeta, int accumulate, vfs_stream_t s);
int vfs_relu_backward(const float* dy, const float* out, float* dx, size_t n, vfs_stream_t s);
int vfs_avgpool_backward(const float* dy, float* dx_nchw, int B, int C, int HW, vfs_stream_t s);
int vfs_cosine_loss_backward(const float* p, const float* z, const float* gout, float* dp, int B, int D,
                             int with_norm, int negative, vfs_stream_t s);
/* torch.optim.SGD update (momentum, weight decay, dampening 0): g' = grad_scale*g + wd*p; buf = first ? g' :
 * momentum*buf + g'; p -= lr*buf.  Replaces the optimizer step of mmcv's OptimizerHook (configs/*:134). */
int vfs_sgd_momentum_step(float* p, const float* g, float* buf, size_t n, float lr, float momentum, float wd,
                          int first, float grad_scale, vfs_stream_t s);

/* The same update with {lr, momentum, weight_decay, grad_scale} read from DEVICE memory (hyper fp32[4]): a captured
 * CUDA graph follows the per-iteration LR schedule of the configs (lr_config CosineAnnealing by_epoch=False,
 * configs/*:136) by copying four floats.  Momentum buffers must start zeroed (== torch's first-step rule). */
int vfs_sgd_momentum_step_dev(float* p, const float* g, float* buf, size_t n, const float* hyper, vfs_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * Peer-memory communicator: the collectives of the data-parallel training step as kernels over NVLink peer memory
 * (one process per GPU of one box).  Replaces, for this path, the NCCL calls torch issues for
 *   - torch.nn.SyncBatchNorm (norm_cfg=dict(type='SyncBN'), configs/*:9,15): per-layer statistics exchange, fwd + bwd
 *   - MMDistributedDataParallel's gradient all-reduce (mmaction/apis/train.py:58-66)
 *   - BaseTracker._parse_losses' all-reduce of the logged scalars (mmaction/models/trackers/base.py:103-108)
 * Every rank allocates a symmetric segment (control words + small-exchange slots + a caller-visible data region) and
 * exports a CUDA IPC handle; the caller moves the handles between the processes (torch.distributed here) and
 * connects.  After that nothing involves the host: every call below only enqueues kernels on `s`, so the whole
 * multi-rank step can live in one CUDA graph.  All ranks must issue the same sequence of calls.
 *   vfs_comm_allreduce_small_*  in-place sum over ranks of up to 4096 fp64 / 8192 fp32 values (device pointer
 *                               anywhere); summed in rank order -> bit-identical on all ranks
 *   vfs_comm_allreduce_f32      in-place sum (times `scale`) of n floats at byte `offset` of the DATA region of
 *                               every rank (16-byte aligned, n % 4 == 0): barrier, two-shot reduce/broadcast, barrier
 *   vfs_comm_error              1 if a spin-wait's watchdog expired since creation (VFS_COMM_TIMEOUT_MS, default
 *                               20000; synchronous read)
 * ---------------------------------------------------------------------------------------------- */
typedef struct VfsComm VfsComm;
size_t vfs_comm_handle_bytes(void);
int vfs_comm_create(int rank, int world, size_t data_bytes, VfsComm** out, void* handle_out);
int vfs_comm_connect(VfsComm* c, const void* all_handles /* [world][vfs_comm_handle_bytes()] in rank order */);
int vfs_comm_destroy(VfsComm* c);
void* vfs_comm_data_ptr(VfsComm* c);
size_t vfs_comm_data_bytes(VfsComm* c);
int vfs_comm_error(VfsComm* c);
int vfs_comm_allreduce_small_f64(VfsComm* c, double* data, int n, vfs_stream_t s);
int vfs_comm_allreduce_small_f32(VfsComm* c, float* data, int n, vfs_stream_t s);
int vfs_comm_barrier(VfsComm* c, vfs_stream_t s);
int vfs_comm_allreduce_f32(VfsComm* c, size_t offset_bytes, size_t n, float scale, vfs_stream_t s);

/* The packers with a power-of-two weight scale: w * wscale is split.  Weights of ~0.02 (kaiming initialisation) have
 * their lo plane in the fp16 subnormal range and keep ~17 significant bits; scaled by 256 both planes are normal for
 * |w| >= 5e-4 and the pair keeps 22.  The caller multiplies the conv epilogue's scale vector (vfs_conv_bn_act `scale`,
 * the `ones` of vfs_conv_stats_split / vfs_conv_dgrad) by 1 / wscale -- exact, a power of two. */
int vfs_pack_conv_weight_scaled(const float* w_oihw, void* w_split, int Cout, int Cin, int ksize, float wscale,
                                vfs_stream_t s);
int vfs_pack_conv_weight_dgrad_scaled(const float* w_oihw, void* wt_split, int Cout, int Cin, int ksize, float wscale,
                                      vfs_stream_t s);
/* OIHW fp32 [Cout,Cin,k,k] -> split [2][Cout][k*k*Cin] (device to device). */
int vfs_pack_conv_weight(const float* w_oihw, void* w_split, int Cout, int Cin, int ksize, vfs_stream_t s);

/* All conv weights of a network in one launch (training re-packs them after every optimiser step).  items: DEVICE
 * array of n entries sorted by first_block; entry i owns blocks [first_block_i, first_block_{i+1}) of a launch of
 * total_blocks blocks, vfs_pack_blocks(Cout, Cin, ksize) blocks each.  mode 0 = vfs_pack_conv_weight layout,
 * 1 = vfs_pack_conv_weight_dgrad layout. */
typedef struct VfsPackItem {
  const float* w;   /* OIHW fp32 [Cout,Cin,k,k] */
  void* dst_split;  /* split [2][...] */
  int32_t Cout, Cin, ksize, mode;
  int32_t first_block;
  int32_t scale_log2; /* the weights are multiplied by 2^scale_log2 before the split (see vfs_pack_conv_weight_scaled) */
} VfsPackItem;
int vfs_pack_blocks(int Cout, int Cin, int ksize);
int vfs_pack_conv_weights_multi(const VfsPackItem* items_dev, int n, int total_blocks, vfs_stream_t s);

/* Test instrument: the same contract as vfs_conv_bn_act computed with plain fp32 FMAs (one thread per
 * output element, fixed summation order).  Exists so the tensor-core path can be checked on the device
 * at sizes the CPU oracle cannot reach; never called by the product path. */
int vfs_debug_conv_bn_act_simt(const VfsConvDesc* d, const void* in_split, const void* w_split, const float* scale,
                               const float* shift, const void* residual_split, float* out_f32_nhwc,
                               vfs_stream_t s);

/* Performance instrument: while `buffer` (device memory, 148 CTAs x 3 roles x events_per_role x 2 int64) is set,
 * vfs_conv_bn_act / vfs_conv_stats / vfs_conv_dgrad launches record an in-kernel timeline of (event code, SM
 * clock) pairs per CTA for the TMA warp (role 0), the MMA warp (1) and the first epilogue warp (2); see
 * tools/conv_trace.py.  Pass NULL to switch it off (the default; the product path never enables it). */
int vfs_debug_conv_trace(long long* buffer, int events_per_role);
/* Tile policy of vfs_conv_bn_act / vfs_conv_dgrad.  The kernel has a CTA-pair form (thread-block clusters of two,
 * tcgen05.mma.cta_group::2, 256-pixel x 256/128-channel tiles, each CTA staging half of the weight tile) that is
 * faster on large launches and a 1-CTA form with finer tiles for small ones.
 *   mode 0 = never pair, 1 = pair whenever Cout % 128 == 0, 2 (default) = pair when the launch has at least
 *   min_pair_tiles pair tiles (default 48 of the 74 TPCs).  Results are identical in every mode. */
int vfs_conv_set_pair_policy(int mode, int min_pair_tiles);

/* ------------------------------------------------------------------------------------------------
 * Restricted-attention label propagation (DAVIS inference).  Replaces masked_attention_efficient,
 * mmaction/models/common/local_attention.py:237-348 (F.normalize :277-279, einsum :289-291, masked_fill
 * :292-313, topk :316, index_select :320-326, softmax + einsum :327-334), with the HW x HW bool mask of
 * spatial_neighbor (common/affinity_utils.py:119-156) replaced by its analytic predicate.
 * ---------------------------------------------------------------------------------------------- */
typedef struct VfsAttnDesc {
  int32_t H, W;         /* feature-map size (queries and keys) */
  int32_t C;            /* attention channels, multiple of 64 */
  int32_t Cv;           /* value channels (objects + background) */
  int32_t T;            /* number of key frames, 1..32 */
  int32_t topk;         /* 1..16 */
  int32_t mask_mode;    /* 0 none, 1 circle: dy^2+dx^2 < radius_y^2, 2 square: |dy|<=radius_y && |dx|<=radius_x */
  int32_t radius_y, radius_x;
  int32_t non_mask_len; /* leading key frames exempt from the mask */
  int32_t mode;         /* 0 softmax, 1 cosine (clamp(min=0)^2) */
  float temperature;    /* > 0 */
} VfsAttnDesc;

/* NCHW fp32 features [N,C,H,W] -> split NHWC, optionally L2-normalised over C (F.normalize, eps 1e-12).
 * inv_norm_ws: fp32 [N*H*W] scratch (only when normalize != 0). */
int vfs_features_to_split(const float* in_nchw, void* out_split, void* inv_norm_ws, int N, int C, int H, int W,
                          int normalize, vfs_stream_t s);
/* same with a channel-padded / row-padded destination (pixel rows c_stride >= C elements apart, lo plane at
 * +plane_stride elements; the caller zero-fills the padding) -- used to feed odd channel counts to the GEMM. */
int vfs_features_to_split_ex(const float* in_nchw, void* out_split, void* inv_norm_ws, int N, int C, int H, int W,
                             int normalize, int c_stride, long long plane_stride, vfs_stream_t s);
/* Label post-processing of the tracker (trackers/vanilla_tracker.py:162-181): bilinear upsample (align_corners=False)
 * of logit fp32 [Cv][h][w] to [H][W], per-channel min-max normalisation where max > 0, arg-max over channels ->
 * out_labels uint8 [H][W].  workspace: vfs_seg_postprocess_workspace_bytes(Cv). */
size_t vfs_seg_postprocess_workspace_bytes(int Cv);
int vfs_seg_postprocess(const float* logit, unsigned char* out_labels, void* workspace, int Cv, int h, int w, int H,
                        int W, vfs_stream_t s);
/* num_maps independent maps in the same three launches (one frame of num_maps videos): logit [num_maps][Cv][h][w] ->
 * out_labels [num_maps][H][W]; workspace num_maps * vfs_seg_postprocess_workspace_bytes(Cv). */
int vfs_seg_postprocess_batched(const float* logit, unsigned char* out_labels, void* workspace, int num_maps, int Cv,
                                int h, int w, int H, int W, vfs_stream_t s);
/* General form of masked_attention_efficient (local_attention.py:287-342) for an arbitrary boolean mask and / or
 * topk = None -- the cases the fused window kernel (vfs_masked_attention*) does not take.  affinity fp32
 * [rows = T*HWk][ld >= HWq] of ONE batch item, already divided by the temperature (vfs_conv_bn_act: key pixels as the
 * image, query pixels as 1x1 filters, scale = 1/temperature); mask uint8 [HWk][HWq] or NULL, applied to key frames
 * t >= non_mask_len; values fp32 [Cv][rows]; topk in [1,16] or 0 = all keys; mode 0 softmax, 1 clamp(min=0)^2.
 * out fp32 [Cv][HWq]. */
int vfs_generic_attention(const float* affinity, int rows, int ld, int HWk, int HWq, const unsigned char* mask,
                          int non_mask_len, const float* values, int Cv, int topk, int mode, float* out,
                          vfs_stream_t s);
/* SiamFC response post-processing (TrackerSiamFC.update, projects/siamfc-pytorch/siamfc/siamfc_tracker_base.py:263-291):
 * responses fp32 [num_scales][R][R] -> cv2-compatible bicubic upsample to [U][U], scale_penalty on the non-centre
 * scales, scale with the largest peak, (map - min) / sum blended with hann_window (fp64 [U][U], already normalised)
 * by window_influence, first arg-max.  out_scale_row_col int32[3] = {scale id, peak row, peak column} (device memory).
 * workspace: vfs_siamfc_peak_workspace_bytes(num_scales, U). */
size_t vfs_siamfc_peak_workspace_bytes(int num_scales, int upscaled_size);
int vfs_siamfc_response_peak(const float* responses, int num_scales, int response_size, int upscaled_size,
                             const double* hann_window, float scale_penalty, float window_influence, void* workspace,
                             int32_t* out_scale_row_col, vfs_stream_t s);
/* Dense helpers of common/affinity_utils.py:6-50 (compute_affinity / propagate; exported by the reference, unused by
 * its trackers).  The HW x HW GEMM itself is vfs_conv_bn_act with the dst pixels as 1x1 filters; these finish it:
 *   vfs_masked_softmax  A [B][R][ld] -> out [B][R][Cc]: mask (analytic window, mask[i,j]) to -inf, softmax over
 *                       dim 1 (rows, per column) / 2 (columns, per row) / 0 (none); fully masked lines -> 0 (or NaN)
 *   vfs_propagate_dense out[b,c,j] = sum_i img[b,c,i] A'[b,i,j], A' = A or clamp(A - kth_j, 0)/max(sum, 1e-12) */
int vfs_masked_softmax(const float* A, float* out, int B, int R, int Cc, int ld, int softmax_dim, int mask_mode,
                       int radius_y, int radius_x, int W, int nan_to_zero, vfs_stream_t s);
int vfs_propagate_dense(const float* img, const float* A, float* out, int B, int Cv, int HW, int topk, vfs_stream_t s);
/* split NHWC -> L2-normalised split NHWC (plane strides in elements; in == out allowed). */
int vfs_normalize_split(const void* in_split, void* out_split, long long num_pixels, int C,
                        long long in_plane_stride, long long out_plane_stride, vfs_stream_t s);

/* Form of the scores kernel: 1 (default) = two 128-key tiles per step (N = 256 MMAs, the query tile is staged once per
 * two key tiles), 0 = one key tile per step.  Identical results. */
int vfs_attention_set_wide(int mode);
size_t vfs_attention_workspace_bytes(const VfsAttnDesc* d, int num_problems);
/*   q_split        hi plane of the query frame [H][W][C]; lo plane at +q_plane_stride elements
 *   k_bank_split   hi plane of a bank of frames [k_bank_frames][H][W][C]; lo plane at +k_plane_stride elements
 *   key_frame_ids  HOST array [T]: bank frame used by key slot t (a frame may repeat, cf. vanilla_tracker.py:133-149)
 *   values         fp32; element (slot t, channel c, position p) at values[key_frame_ids[t]*v_frame_stride +
 *                  c*v_chan_stride + p]
 *   out            fp32 [Cv][H*W]
 *   out_topk_val / out_topk_idx   NULL or [topk][H*W]: selected affinities (already / temperature) and flat key
 *                  indices t*H*W + p, sorted descending -- for index-parity tests */
int vfs_masked_attention(const VfsAttnDesc* d, const void* q_split, long long q_plane_stride,
                         const void* k_bank_split, long long k_plane_stride, int k_bank_frames,
                         const int32_t* key_frame_ids, const float* values, long long v_frame_stride,
                         long long v_chan_stride, float* out, float* out_topk_val, int32_t* out_topk_idx,
                         void* workspace, size_t workspace_bytes, vfs_stream_t s);

/* Several independent (query frame, key set) problems of the same shape in ONE launch (e.g. the frame pairs of a
 * batch of clips, or the N batch items of masked_attention_efficient).  num_problems <= 32, num_problems*T <= 256.
 *   q_frame_ids [P] (host): query frame of problem p inside q_bank_split [q_bank_frames][H][W][C]
 *   key_frame_ids [P*T] (host): key bank frame of (p, slot t);  value_frame_ids [P*T]: frame index used to address
 *   values[p*v_batch_stride + value_frame_ids[p*T+t]*v_frame_stride + c*v_chan_stride + pos]
 *   out fp32 [P][Cv][H*W]; out_topk_* NULL or [P][topk][H*W] */
int vfs_masked_attention_batched(const VfsAttnDesc* d, int num_problems, const void* q_bank_split,
                                 long long q_plane_stride, int q_bank_frames, const int32_t* q_frame_ids,
                                 const void* k_bank_split, long long k_plane_stride, int k_bank_frames,
                                 const int32_t* key_frame_ids, const float* values, const int32_t* value_frame_ids,
                                 long long v_batch_stride, long long v_frame_stride, long long v_chan_stride,
                                 float* out, float* out_topk_val, int32_t* out_topk_idx, void* workspace,
                                 size_t workspace_bytes, vfs_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * SimSiam head + loss.  Replaces SimSiamHead.forward (heads/sim_siam_head.py:143-163: AdaptiveAvgPool2d,
 * nn.Linear, BatchNorm1d/SyncBatchNorm, ReLU) and CosineSimLoss._forward (losses/sim_loss.py:42-63).
 * ---------------------------------------------------------------------------------------------- */
int vfs_global_avg_pool(const float* in_nchw, float* out, int B, int C, int HW, vfs_stream_t s);
/* y[M,N] = x[M,K] W[N,K]^T + bias */
int vfs_linear(const float* x, const float* W, const float* bias, float* y, int M, int N, int K, vfs_stream_t s);
/* in-place BatchNorm1d over y[M,N] (+ReLU); training != 0 uses batch statistics and updates the running ones */
int vfs_bn1d_act(float* y, int M, int N, const float* gamma, const float* beta, float* running_mean,
                 float* running_var, float eps, float momentum, int training, int relu, float* save_mean,
                 float* save_invstd, vfs_stream_t s);
int vfs_relu(float* y, size_t n, vfs_stream_t s);
/* loss[b] = 2 - 2*cos(p[b], z[b])  (negative != 0: -cos; with_norm == 0: raw dot product) */
int vfs_cosine_sim_loss(const float* p, const float* z, float* loss, int B, int D, int with_norm, int negative,
                        vfs_stream_t s);

/* Data feed: ``Normalize`` (mmaction/datasets/pipelines/augmentations.py:711-757 -> mmcv.imnormalize_)
 * + ``FormatShape('NCTHW')`` (formating.py:248-258) on the device.  frames uint8 [clips][T][H][W][3] (device memory,
 * H*W % 4 == 0) -> out fp32 [clips][3][T][H][W] = fp32(fp64(fp32(x - mean)) * stdinv) per channel, which is what
 * cv2.subtract / cv2.multiply produce for mmcv's float64 mean and 1/std vectors; swap_rb != 0 swaps channels 0 and 2
 * first (the pipeline's ``to_bgr``).  mean3 (three floats) and stdinv3 (three doubles = 1 / double(std)) are HOST
 * pointers. */
int vfs_frames_u8_to_ncthw_f32(const unsigned char* frames, float* out, long long clips, int T, int H, int W,
                               const float* mean3, const double* stdinv3, int swap_rb, vfs_stream_t s);
/* Training data feed: RandomResizedCrop's crop (box drawn on the host, augmentations.py:214-262) ->
 * Resize(scale, keep_ratio=False) = cv2.resize INTER_LINEAR on uint8, bit for bit (:487-597 -> mmcv.imresize) ->
 * Flip horizontal (:600-711) -> Normalize -> FormatShape('NCTHW'), one kernel per batch.  items: DEVICE array of
 * clips*T entries (frame f = clip*T + t); frames may differ in size.  out fp32 [clips][3][T][dst_h][dst_w]. */
typedef struct VfsAugItem {
  const unsigned char* src; /* uint8 HWC frame [H][W][3] in device memory */
  int32_t H, W;
  int32_t crop_x0, crop_y0, crop_w, crop_h; /* crop box inside the frame */
  int32_t flip;                             /* mirror horizontally after the resize */
  int32_t reserved;
} VfsAugItem;
int vfs_augment_u8_to_ncthw_f32(const VfsAugItem* items_dev, float* out, long long clips, int T, int dst_h, int dst_w,
                                const float* mean3, const double* stdinv3, int swap_rb, vfs_stream_t s);
/* ------------------------------------------------------------------------------------------------
 * SiamFC cross-correlation.  Replaces SiamFC._fast_xcorr (projects/siamfc-pytorch/siamfc/heads.py:16-23).
 *   z fp32 NHWC [nz,hz,wz,C], x fp32 NHWC [nx,h,w,C] -> out fp32 [nx,1,h-hz+1,w-wz+1], x[i] pairs with z[i % nz]
 * ---------------------------------------------------------------------------------------------- */
int vfs_nchw_to_nhwc_f32(const float* in, float* out, int N, int C, int H, int W, vfs_stream_t s);
int vfs_xcorr_nhwc(const float* z, const float* x, float* out, int nz, int nx, int C, int hz, int wz, int h, int w,
                   float out_scale, vfs_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * SiamFC linear-probe training (TrackerSiamFC.train_step, projects/siamfc-pytorch/siamfc/siamfc_tracker_base.py:364-386;
 * the backbone is frozen, default_config_base.py:40-49, so only the head trains).
 *   vfs_siamfc_loss          FocalLoss (mode 0, losses.py:43-64) / BalancedLoss (mode 1, :27-40) of responses vs labels
 *                            (n values each): loss[0] and, when grad != NULL, d loss / d responses, one launch
 *   vfs_xcorr_backward_nhwc  gradients of vfs_xcorr_nhwc for n (exemplar, search) pairs: dz [n,hz,wz,C], dx [n,h,w,C]
 *                            from dr [n,1,h-hz+1,w-wz+1] (either output may be NULL)
 *   vfs_adam_step            torch.optim.Adam update (no amsgrad), `step` = 1-based step count
 * The 1x1 adapter convolutions and their weight gradients use vfs_conv_bn_act / vfs_conv_wgrad. */
int vfs_siamfc_loss(const float* responses, const float* labels, float* loss, float* grad, int n, int mode,
                    float gamma, float neg_weight, vfs_stream_t s);
int vfs_xcorr_backward_nhwc(const float* dr, const float* z, const float* x, float* dz, float* dx, int n, int C, int hz,
                            int wz, int h, int w, float out_scale, vfs_stream_t s);
int vfs_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
                  float weight_decay, int step, vfs_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* VFS_B200_H_ */
