// C-ABI surface of libvfs_b200.so (see include/vfs_b200.h) + host helpers.
#include <stdarg.h>

#include <stdlib.h>

#include "host_common.h"

namespace vfs {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return VFS_OK;
  set_last_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return VFS_ECUDA;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("VFS_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) {
      set_last_error("cuTensorMapEncodeTiled entry point unavailable (err %d)", static_cast<int>(e));
      return nullptr;
    }
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

int make_tmap_16b_sw128(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return VFS_ECUDA;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                  gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u]",
                   static_cast<int>(r), rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
                   (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0),
                   (unsigned long long)(rank > 4 ? gdim[4] : 0), bdim[0], rank > 1 ? bdim[1] : 0,
                   rank > 2 ? bdim[2] : 0, rank > 3 ? bdim[3] : 0, rank > 4 ? bdim[4] : 0);
    return VFS_ECUDA;
  }
  return VFS_OK;
}

// implemented in the kernel translation units
int conv_bn_act_tc(const VfsConvDesc* d, const void* in_split, const void* w_split, const float* scale,
                   const float* shift, const void* residual_split, void* out_split, float* out_f32,
                   double* stats, cudaStream_t stream);
int conv_set_trace(long long* buffer, int events_per_role);
int conv_set_pair_policy(int mode, int min_pair_tiles);
int conv_dgrad_tc(const VfsConvDesc* d, const void* dz_split, const void* wt_split, const float* ones,
                  const float* zeros, const void* add_split, void* dx_split, cudaStream_t stream);
int pack_conv_weight_dgrad(const float* w, void* wt_split, int Cout, int Cin, int k, float wscale, cudaStream_t s);
size_t wgrad_workspace_bytes(int Cout, int Cin, int ksize);
int conv_wgrad_tc(const VfsConvDesc* d, const void* x_split, const void* dz_split, void* workspace, float* dw_oihw,
                  int accumulate, float out_scale, cudaStream_t stream);
int affine_act_f32(float* y, const float* scale, const float* shift, long long M, int C, int relu, cudaStream_t s);
int bn_bwd_reduce(const void* dy_split, const float* dy_f32, const void* y_split, const float* y_f32, const float* z,
                  const void* z_split, const float* mean, const float* invstd, double* sums, long long M, int C,
                  cudaStream_t s);
int bn_bwd_apply(const void* dy_split, const float* dy_f32, const void* y_split, const float* y_f32, const float* z,
                 const void* z_split, const float* mean, const float* invstd, const float* gamma, const double* sums,
                 double count, void* dz_split,
                 float* dz_f32, void* g_split, float* dgamma, float* dbeta, int accumulate, float param_scale,
                 long long M, int C, cudaStream_t s);
int relu_bwd_split(const void* dy_split, const void* y_split, void* g_split, long long elems, cudaStream_t s);
int stem_pool_relu_bwd(const void* dpool_split, const float* z, const float* scale, const float* shift, float* g,
                       int N, int H, int W, cudaStream_t s);
int stem_wgrad(const float* x, const float* dz, float* dw, int accumulate, float out_scale, int N, int H, int W,
               cudaStream_t s);
int linear_backward(const float* dy, const float* x, const float* W, float* dx, float* dW, float* db, int M, int N,
                    int K, int accumulate, cudaStream_t s);
int bn1d_backward(const float* dy, const float* pre, const float* out, float* dpre, int M, int N, const float* gamma,
                  const float* mean, const float* invstd, int training, int relu, float* dgamma, float* dbeta,
                  int accumulate, cudaStream_t s);
int relu_backward(const float* dy, const float* out, float* dx, size_t n, cudaStream_t s);
int avgpool_backward_nchw(const float* dy, float* dx, int B, int C, int HW, cudaStream_t s);
int cosine_loss_backward(const float* p, const float* z, const float* gout, float* dp, int B, int D, int with_norm,
                         int negative, cudaStream_t s);
int sgd_momentum_step(float* p, const float* g, float* buf, size_t n, float lr, float momentum, float wd, int first,
                      float grad_scale, cudaStream_t s);
int sgd_momentum_step_dev(float* p, const float* g, float* buf, size_t n, const float* hyper, cudaStream_t s);
struct Comm;
size_t comm_handle_bytes();
int comm_create(int rank, int world, size_t data_bytes, Comm** out, void* handle_out);
int comm_connect(Comm* c, const void* all_handles);
int comm_destroy(Comm* c);
void* comm_data_ptr(Comm* c);
size_t comm_data_bytes(Comm* c);
int comm_error(Comm* c);
int comm_allreduce_small(Comm* c, void* data, int n, int is_f64, cudaStream_t s);
int comm_barrier(Comm* c, cudaStream_t s);
int comm_allreduce_f32(Comm* c, size_t offset_bytes, size_t n, float scale, cudaStream_t s);
int channel_stats_f32(const float* x, double* stats, long long M, int C, cudaStream_t s);
int bn_finalize(double* stats, double count, const float* gamma, const float* beta, float* running_mean,
                float* running_var, float momentum, float eps, float* scale, float* shift, float* save_mean,
                float* save_invstd, int C, cudaStream_t s);
int bn_apply(const float* z, const void* z_split, const float* scale, const float* shift, const void* residual_split,
             void* out_split, long long M, int C, int relu, cudaStream_t s);
int conv_bn_act_simt(const VfsConvDesc* d, const void* in_split, const void* w_split, const float* scale,
                     const float* shift, const void* residual_split, float* out_f32, cudaStream_t stream);
int nchw_f32_to_split(const float* in, void* out_split, int N, int C, int H, int W, float scale, cudaStream_t s);
int split_to_nchw_f32(const void* in_split, float* out, int N, int C, int H, int W, cudaStream_t s);
int pack_conv_weight(const float* w, void* w_split, int Cout, int Cin, int k, float wscale, cudaStream_t s);
int pack_conv_weights_multi(const VfsPackItem* items_dev, int n, int total_blocks, cudaStream_t s);
size_t stem_workspace_bytes(int N, int H, int W);
int stem_forward(const float* in, const void* weight, const float* scale, const float* shift, void* out_split,
                 void* workspace, int N, int H, int W, cudaStream_t s);
size_t stem_packed_weight_bytes();
int stem_pack_weight(const float* w, void* w_split, cudaStream_t s);

int stem_conv_raw(const float* in, const void* weight, void* conv_out, int N, int H, int W, cudaStream_t s);
int stem_bn_relu_pool(const void* conv_out, const float* scale, const float* shift, void* out_split, int N, int H,
                      int W, cudaStream_t s);
int features_to_split(const float* in_nchw, void* out_split, void* inv_norm_ws, int N, int C, int H, int W,
                      int normalize, int c_stride, long long plane_stride, cudaStream_t s);
size_t seg_postprocess_workspace_bytes(int Cv);
size_t siamfc_peak_workspace_bytes(int S, int U);
int siamfc_response_peak(const float* responses, int S, int R, int U, const double* hann, float scale_penalty,
                         float window_influence, void* workspace, int* out3, cudaStream_t s);
int seg_postprocess(const float* logit, unsigned char* out, void* workspace, int P, int Cv, int h, int w, int H,
                    int W, cudaStream_t s);
int masked_softmax(const float* A, float* out, int B, int R, int Cc, int ld, int softmax_dim, int mask_mode, int ry,
                   int rx, int W, int nan_to_zero, cudaStream_t s);
int propagate_dense(const float* img, const float* A, float* out, int B, int Cv, int HW, int topk, cudaStream_t s);
int generic_attention(const float* A, int rows, int ld, int HWk, int HWq, const unsigned char* mask, int non_mask_len,
                      const float* values, int Cv, int topk, int mode, float* out, cudaStream_t s);
int normalize_split(const void* in_split, void* out_split, long long num_pixels, int C, long long in_plane_stride,
                    long long out_plane_stride, cudaStream_t s);
size_t attention_workspace_bytes(const VfsAttnDesc* d, int B);
int attention_set_wide(int mode);
int masked_attention_batched(const VfsAttnDesc* d, int B, const void* q_bank_split, long long q_plane_stride,
                             int q_bank_frames, const int* q_ids, const void* k_bank_split, long long k_plane_stride,
                             int k_bank_frames, const int* key_ids, const float* values, const int* val_ids,
                             long long v_batch_stride, long long v_frame_stride, long long v_chan_stride, float* out,
                             float* out_topk_val, int* out_topk_idx, void* workspace, size_t workspace_bytes,
                             cudaStream_t stream);
int masked_attention(const VfsAttnDesc* d, const void* q_split, long long q_plane_stride, const void* k_bank_split,
                     long long k_plane_stride, int k_bank_frames, const int* key_frame_ids, const float* values,
                     long long v_frame_stride, long long v_chan_stride, float* out, float* out_topk_val,
                     int* out_topk_idx, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int global_avg_pool_nchw(const float* in, float* out, int B, int C, int HW, cudaStream_t s);
int linear_forward(const float* x, const float* W, const float* bias, float* y, int M, int N, int K, cudaStream_t s);
int bn1d_act(float* y, int M, int N, const float* gamma, const float* beta, float* running_mean, float* running_var,
             float eps, float momentum, int training, int relu, float* save_mean, float* save_invstd, cudaStream_t s);
int relu_inplace(float* y, size_t n, cudaStream_t s);
int cosine_sim_loss(const float* p, const float* z, float* loss, int B, int D, int with_norm, int negative,
                    cudaStream_t s);
int nchw_to_nhwc_f32(const float* in, float* out, int N, int C, int H, int W, cudaStream_t s);
int frames_u8_to_ncthw_f32(const unsigned char* in, float* out, long long clips, int T, int H, int W, const float* mean3,
                           const double* stdinv3, int swap_rb, cudaStream_t s);
int augment_u8_to_ncthw_f32(const VfsAugItem* items_dev, float* out, long long clips, int T, int dst_h, int dst_w,
                            const float* mean3, const double* stdinv3, int swap_rb, cudaStream_t s);
int siamfc_loss(const float* responses, const float* labels, float* loss, float* grad, int n, int mode, float gamma,
                float neg_weight, cudaStream_t s);
int xcorr_backward_nhwc(const float* dr, const float* z, const float* x, float* dz, float* dx, int n, int C, int hz, int wz,
                        int h, int w, float out_scale, cudaStream_t s);
int adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
              float weight_decay, int step, cudaStream_t s);
int xcorr_nhwc(const float* z, const float* x, float* out, int nz, int nx, int C, int hz, int wz, int h, int w,
               float out_scale, cudaStream_t s);

unsigned int overflow_conv(int), overflow_layout(int), overflow_stem(int), overflow_affinity(int), overflow_bn(int),
    overflow_train(int);

}  // namespace vfs

extern "C" {

const char* vfs_last_error_string(void) { return vfs::g_last_error; }
int vfs_abi_version(void) { return 1; }

/* Number of values that exceeded the fp16 range (|x| > 65504) when a tensor was split since the last reset.
 * Synchronises the device. */
unsigned int vfs_overflow_count(int reset) {
  return vfs::overflow_conv(reset) + vfs::overflow_layout(reset) + vfs::overflow_stem(reset) +
         vfs::overflow_affinity(reset) + vfs::overflow_bn(reset) + vfs::overflow_train(reset);
}

int vfs_check_device(void) {
  int dev = 0;
  VFS_CUDA_OK(cudaGetDevice(&dev));
  int major = 0;
  VFS_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  VFS_REQUIRE(major == 10, VFS_EARCH, "device compute capability major %d, need 10 (sm_100a)", major);
  return VFS_OK;
}

int vfs_nchw_f32_to_split(const float* in, void* out_split, int N, int C, int H, int W, vfs_stream_t s) {
  return vfs::nchw_f32_to_split(in, out_split, N, C, H, W, 1.0f, s);
}
int vfs_nchw_f32_to_split_scaled(const float* in, void* out_split, int N, int C, int H, int W, float scale,
                                 vfs_stream_t s) {
  return vfs::nchw_f32_to_split(in, out_split, N, C, H, W, scale, s);
}
int vfs_split_to_nchw_f32(const void* in_split, float* out, int N, int C, int H, int W, vfs_stream_t s) {
  return vfs::split_to_nchw_f32(in_split, out, N, C, H, W, s);
}
size_t vfs_stem_workspace_bytes(int N, int H, int W) { return vfs::stem_workspace_bytes(N, H, W); }
size_t vfs_stem_packed_weight_bytes(void) { return vfs::stem_packed_weight_bytes(); }
int vfs_stem_pack_weight(const float* w_oihw, void* w_split, vfs_stream_t s) { return vfs::stem_pack_weight(w_oihw, w_split, s); }
int vfs_stem_forward(const float* in, const void* weight, const float* scale, const float* shift, void* out_split,
                     void* workspace, int N, int H, int W, vfs_stream_t s) {
  return vfs::stem_forward(in, weight, scale, shift, out_split, workspace, N, H, W, s);
}
int vfs_stem_conv_raw(const float* in, const void* weight, void* conv_out_f32_nhwc, int N, int H, int W,
                      vfs_stream_t s) {
  return vfs::stem_conv_raw(in, weight, conv_out_f32_nhwc, N, H, W, s);
}
int vfs_stem_bn_relu_pool(const void* conv_out_f32_nhwc, const float* scale, const float* shift, void* out_split,
                          int N, int H, int W, vfs_stream_t s) {
  return vfs::stem_bn_relu_pool(conv_out_f32_nhwc, scale, shift, out_split, N, H, W, s);
}
int vfs_conv_bn_act(const VfsConvDesc* d, const void* in_split, const void* w_split, const float* scale,
                    const float* shift, const void* residual_split, void* out_split, float* out_f32_nhwc,
                    vfs_stream_t s) {
  return vfs::conv_bn_act_tc(d, in_split, w_split, scale, shift, residual_split, out_split, out_f32_nhwc, nullptr, s);
}
int vfs_conv_stats(const VfsConvDesc* d, const void* in_split, const void* w_split, const float* scale,
                   const float* shift, float* out_f32_nhwc, double* stats, vfs_stream_t s) {
  return vfs::conv_bn_act_tc(d, in_split, w_split, scale, shift, nullptr, nullptr, out_f32_nhwc, stats, s);
}
int vfs_conv_dgrad(const VfsConvDesc* d, const void* dz_split, const void* wt_split, const float* ones,
                   const float* zeros, const void* add_split, void* dx_split, vfs_stream_t s) {
  return vfs::conv_dgrad_tc(d, dz_split, wt_split, ones, zeros, add_split, dx_split, s);
}
int vfs_pack_conv_weight_dgrad(const float* w_oihw, void* wt_split, int Cout, int Cin, int ksize, vfs_stream_t s) {
  return vfs::pack_conv_weight_dgrad(w_oihw, wt_split, Cout, Cin, ksize, 1.0f, s);
}
int vfs_pack_conv_weight_dgrad_scaled(const float* w_oihw, void* wt_split, int Cout, int Cin, int ksize, float wscale,
                                      vfs_stream_t s) {
  return vfs::pack_conv_weight_dgrad(w_oihw, wt_split, Cout, Cin, ksize, wscale, s);
}
size_t vfs_conv_wgrad_workspace_bytes(int Cout, int Cin, int ksize) {
  return vfs::wgrad_workspace_bytes(Cout, Cin, ksize);
}
int vfs_conv_wgrad(const VfsConvDesc* d, const void* x_split, const void* dz_split, void* workspace, float* dw_oihw,
                   int accumulate, float out_scale, vfs_stream_t s) {
  return vfs::conv_wgrad_tc(d, x_split, dz_split, workspace, dw_oihw, accumulate, out_scale, s);
}
int vfs_affine_act_f32(float* y, const float* scale, const float* shift, long long M, int C, int relu,
                       vfs_stream_t s) {
  return vfs::affine_act_f32(y, scale, shift, M, C, relu, s);
}
int vfs_bn_bwd_reduce(const void* dy_split, const float* dy_f32, const void* y_split, const float* y_f32,
                      const float* z, const void* z_split, const float* mean, const float* invstd, double* sums,
                      long long M, int C, vfs_stream_t s) {
  return vfs::bn_bwd_reduce(dy_split, dy_f32, y_split, y_f32, z, z_split, mean, invstd, sums, M, C, s);
}
int vfs_bn_bwd_apply(const void* dy_split, const float* dy_f32, const void* y_split, const float* y_f32,
                     const float* z, const void* z_split, const float* mean, const float* invstd, const float* gamma,
                     const double* sums, double count, void* dz_split, float* dz_f32, void* g_split, float* dgamma,
                     float* dbeta, int accumulate, float param_scale, long long M, int C, vfs_stream_t s) {
  return vfs::bn_bwd_apply(dy_split, dy_f32, y_split, y_f32, z, z_split, mean, invstd, gamma, sums, count, dz_split, dz_f32,
                           g_split, dgamma, dbeta, accumulate, param_scale, M, C, s);
}
int vfs_relu_bwd_split(const void* dy_split, const void* y_split, void* g_split, long long elems, vfs_stream_t s) {
  return vfs::relu_bwd_split(dy_split, y_split, g_split, elems, s);
}
int vfs_stem_pool_relu_bwd(const void* dpool_split, const float* z, const float* scale, const float* shift, float* g,
                           int N, int H, int W, vfs_stream_t s) {
  return vfs::stem_pool_relu_bwd(dpool_split, z, scale, shift, g, N, H, W, s);
}
int vfs_stem_wgrad(const float* x, const float* dz, float* dw, int accumulate, float out_scale, int N, int H, int W,
                   vfs_stream_t s) {
  return vfs::stem_wgrad(x, dz, dw, accumulate, out_scale, N, H, W, s);
}
int vfs_linear_backward(const float* dy, const float* x, const float* W, float* dx, float* dW, float* db, int M,
                        int N, int K, int accumulate, vfs_stream_t s) {
  return vfs::linear_backward(dy, x, W, dx, dW, db, M, N, K, accumulate, s);
}
int vfs_bn1d_backward(const float* dy, const float* pre, const float* out, float* dpre, int M, int N,
                      const float* gamma, const float* mean, const float* invstd, int training, int relu,
                      float* dgamma, float* dbeta, int accumulate, vfs_stream_t s) {
  return vfs::bn1d_backward(dy, pre, out, dpre, M, N, gamma, mean, invstd, training, relu, dgamma, dbeta, accumulate, s);
}
int vfs_relu_backward(const float* dy, const float* out, float* dx, size_t n, vfs_stream_t s) {
  return vfs::relu_backward(dy, out, dx, n, s);
}
int vfs_avgpool_backward(const float* dy, float* dx_nchw, int B, int C, int HW, vfs_stream_t s) {
  return vfs::avgpool_backward_nchw(dy, dx_nchw, B, C, HW, s);
}
int vfs_cosine_loss_backward(const float* p, const float* z, const float* gout, float* dp, int B, int D,
                             int with_norm, int negative, vfs_stream_t s) {
  return vfs::cosine_loss_backward(p, z, gout, dp, B, D, with_norm, negative, s);
}
int vfs_sgd_momentum_step(float* p, const float* g, float* buf, size_t n, float lr, float momentum, float wd,
                          int first, float grad_scale, vfs_stream_t s) {
  return vfs::sgd_momentum_step(p, g, buf, n, lr, momentum, wd, first, grad_scale, s);
}
int vfs_sgd_momentum_step_dev(float* p, const float* g, float* buf, size_t n, const float* hyper, vfs_stream_t s) {
  return vfs::sgd_momentum_step_dev(p, g, buf, n, hyper, s);
}
size_t vfs_comm_handle_bytes(void) { return vfs::comm_handle_bytes(); }
int vfs_comm_create(int rank, int world, size_t data_bytes, VfsComm** out, void* handle_out) {
  return vfs::comm_create(rank, world, data_bytes, reinterpret_cast<vfs::Comm**>(out), handle_out);
}
int vfs_comm_connect(VfsComm* c, const void* all_handles) {
  return vfs::comm_connect(reinterpret_cast<vfs::Comm*>(c), all_handles);
}
int vfs_comm_destroy(VfsComm* c) { return vfs::comm_destroy(reinterpret_cast<vfs::Comm*>(c)); }
void* vfs_comm_data_ptr(VfsComm* c) { return vfs::comm_data_ptr(reinterpret_cast<vfs::Comm*>(c)); }
size_t vfs_comm_data_bytes(VfsComm* c) { return vfs::comm_data_bytes(reinterpret_cast<vfs::Comm*>(c)); }
int vfs_comm_error(VfsComm* c) { return vfs::comm_error(reinterpret_cast<vfs::Comm*>(c)); }
int vfs_comm_allreduce_small_f64(VfsComm* c, double* data, int n, vfs_stream_t s) {
  return vfs::comm_allreduce_small(reinterpret_cast<vfs::Comm*>(c), data, n, 1, s);
}
int vfs_comm_allreduce_small_f32(VfsComm* c, float* data, int n, vfs_stream_t s) {
  return vfs::comm_allreduce_small(reinterpret_cast<vfs::Comm*>(c), data, n, 0, s);
}
int vfs_comm_barrier(VfsComm* c, vfs_stream_t s) { return vfs::comm_barrier(reinterpret_cast<vfs::Comm*>(c), s); }
int vfs_comm_allreduce_f32(VfsComm* c, size_t offset_bytes, size_t n, float scale, vfs_stream_t s) {
  return vfs::comm_allreduce_f32(reinterpret_cast<vfs::Comm*>(c), offset_bytes, n, scale, s);
}
int vfs_channel_stats_f32(const float* x, double* stats, long long M, int C, vfs_stream_t s) {
  return vfs::channel_stats_f32(x, stats, M, C, s);
}
int vfs_bn_finalize(double* stats, double count, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float momentum, float eps, float* scale, float* shift, float* save_mean,
                    float* save_invstd, int C, vfs_stream_t s) {
  return vfs::bn_finalize(stats, count, gamma, beta, running_mean, running_var, momentum, eps, scale, shift,
                          save_mean, save_invstd, C, s);
}
int vfs_bn_apply(const float* z, const void* z_split, const float* scale, const float* shift,
                 const void* residual_split, void* out_split, long long M, int C, int relu, vfs_stream_t s) {
  return vfs::bn_apply(z, z_split, scale, shift, residual_split, out_split, M, C, relu, s);
}
int vfs_conv_stats_split(const VfsConvDesc* d, const void* in_split, const void* w_split, const float* ones,
                         const float* zeros, void* z_split, double* stats, vfs_stream_t s) {
  return vfs::conv_bn_act_tc(d, in_split, w_split, ones, zeros, nullptr, z_split, nullptr, stats, s);
}
int vfs_pack_blocks(int Cout, int Cin, int ksize) {
  const long long total = static_cast<long long>(Cout) * Cin * ksize * ksize;
  return static_cast<int>((total + 2047) / 2048);
}
int vfs_pack_conv_weights_multi(const VfsPackItem* items_dev, int n, int total_blocks, vfs_stream_t s) {
  return vfs::pack_conv_weights_multi(items_dev, n, total_blocks, s);
}
int vfs_pack_conv_weight(const float* w_oihw, void* w_split, int Cout, int Cin, int ksize, vfs_stream_t s) {
  return vfs::pack_conv_weight(w_oihw, w_split, Cout, Cin, ksize, 1.0f, s);
}
int vfs_pack_conv_weight_scaled(const float* w_oihw, void* w_split, int Cout, int Cin, int ksize, float wscale,
                                vfs_stream_t s) {
  return vfs::pack_conv_weight(w_oihw, w_split, Cout, Cin, ksize, wscale, s);
}
int vfs_debug_conv_bn_act_simt(const VfsConvDesc* d, const void* in_split, const void* w_split, const float* scale,
                               const float* shift, const void* residual_split, float* out_f32_nhwc,
                               vfs_stream_t s) {
  return vfs::conv_bn_act_simt(d, in_split, w_split, scale, shift, residual_split, out_f32_nhwc, s);
}

int vfs_debug_conv_trace(long long* buffer, int events_per_role) {
  return vfs::conv_set_trace(buffer, events_per_role);
}
int vfs_conv_set_pair_policy(int mode, int min_pair_tiles) { return vfs::conv_set_pair_policy(mode, min_pair_tiles); }

int vfs_features_to_split(const float* in_nchw, void* out_split, void* inv_norm_ws, int N, int C, int H, int W,
                          int normalize, vfs_stream_t s) {
  return vfs::features_to_split(in_nchw, out_split, inv_norm_ws, N, C, H, W, normalize, 0, 0, s);
}
int vfs_features_to_split_ex(const float* in_nchw, void* out_split, void* inv_norm_ws, int N, int C, int H, int W,
                             int normalize, int c_stride, long long plane_stride, vfs_stream_t s) {
  return vfs::features_to_split(in_nchw, out_split, inv_norm_ws, N, C, H, W, normalize, c_stride, plane_stride, s);
}
size_t vfs_seg_postprocess_workspace_bytes(int Cv) { return vfs::seg_postprocess_workspace_bytes(Cv); }
int vfs_seg_postprocess(const float* logit, unsigned char* out_labels, void* workspace, int Cv, int h, int w, int H,
                        int W, vfs_stream_t s) {
  return vfs::seg_postprocess(logit, out_labels, workspace, 1, Cv, h, w, H, W, s);
}
size_t vfs_siamfc_peak_workspace_bytes(int num_scales, int upscaled_size) {
  return vfs::siamfc_peak_workspace_bytes(num_scales, upscaled_size);
}
int vfs_siamfc_response_peak(const float* responses, int num_scales, int response_size, int upscaled_size,
                             const double* hann_window, float scale_penalty, float window_influence, void* workspace,
                             int32_t* out_scale_row_col, vfs_stream_t s) {
  return vfs::siamfc_response_peak(responses, num_scales, response_size, upscaled_size, hann_window, scale_penalty,
                                   window_influence, workspace, out_scale_row_col, s);
}
int vfs_seg_postprocess_batched(const float* logit, unsigned char* out_labels, void* workspace, int num_maps, int Cv,
                                int h, int w, int H, int W, vfs_stream_t s) {
  return vfs::seg_postprocess(logit, out_labels, workspace, num_maps, Cv, h, w, H, W, s);
}
int vfs_masked_softmax(const float* A, float* out, int B, int R, int Cc, int ld, int softmax_dim, int mask_mode,
                       int radius_y, int radius_x, int W, int nan_to_zero, vfs_stream_t s) {
  return vfs::masked_softmax(A, out, B, R, Cc, ld, softmax_dim, mask_mode, radius_y, radius_x, W, nan_to_zero, s);
}
int vfs_propagate_dense(const float* img, const float* A, float* out, int B, int Cv, int HW, int topk, vfs_stream_t s) {
  return vfs::propagate_dense(img, A, out, B, Cv, HW, topk, s);
}
int vfs_generic_attention(const float* affinity, int rows, int ld, int HWk, int HWq, const unsigned char* mask,
                          int non_mask_len, const float* values, int Cv, int topk, int mode, float* out,
                          vfs_stream_t s) {
  return vfs::generic_attention(affinity, rows, ld, HWk, HWq, mask, non_mask_len, values, Cv, topk, mode, out, s);
}
int vfs_normalize_split(const void* in_split, void* out_split, long long num_pixels, int C,
                        long long in_plane_stride, long long out_plane_stride, vfs_stream_t s) {
  return vfs::normalize_split(in_split, out_split, num_pixels, C, in_plane_stride, out_plane_stride, s);
}
int vfs_attention_set_wide(int mode) { return vfs::attention_set_wide(mode); }
size_t vfs_attention_workspace_bytes(const VfsAttnDesc* d, int num_problems) {
  return vfs::attention_workspace_bytes(d, num_problems);
}
int vfs_masked_attention_batched(const VfsAttnDesc* d, int num_problems, const void* q_bank_split,
                                 long long q_plane_stride, int q_bank_frames, const int32_t* q_frame_ids,
                                 const void* k_bank_split, long long k_plane_stride, int k_bank_frames,
                                 const int32_t* key_frame_ids, const float* values, const int32_t* value_frame_ids,
                                 long long v_batch_stride, long long v_frame_stride, long long v_chan_stride,
                                 float* out, float* out_topk_val, int32_t* out_topk_idx, void* workspace,
                                 size_t workspace_bytes, vfs_stream_t s) {
  return vfs::masked_attention_batched(d, num_problems, q_bank_split, q_plane_stride, q_bank_frames, q_frame_ids,
                                       k_bank_split, k_plane_stride, k_bank_frames, key_frame_ids, values,
                                       value_frame_ids, v_batch_stride, v_frame_stride, v_chan_stride, out,
                                       out_topk_val, out_topk_idx, workspace, workspace_bytes, s);
}
int vfs_masked_attention(const VfsAttnDesc* d, const void* q_split, long long q_plane_stride,
                         const void* k_bank_split, long long k_plane_stride, int k_bank_frames,
                         const int32_t* key_frame_ids, const float* values, long long v_frame_stride,
                         long long v_chan_stride, float* out, float* out_topk_val, int32_t* out_topk_idx,
                         void* workspace, size_t workspace_bytes, vfs_stream_t s) {
  return vfs::masked_attention(d, q_split, q_plane_stride, k_bank_split, k_plane_stride, k_bank_frames,
                               key_frame_ids, values, v_frame_stride, v_chan_stride, out, out_topk_val, out_topk_idx,
                               workspace, workspace_bytes, s);
}
int vfs_global_avg_pool(const float* in_nchw, float* out, int B, int C, int HW, vfs_stream_t s) {
  return vfs::global_avg_pool_nchw(in_nchw, out, B, C, HW, s);
}
int vfs_linear(const float* x, const float* W, const float* bias, float* y, int M, int N, int K, vfs_stream_t s) {
  return vfs::linear_forward(x, W, bias, y, M, N, K, s);
}
int vfs_bn1d_act(float* y, int M, int N, const float* gamma, const float* beta, float* running_mean,
                 float* running_var, float eps, float momentum, int training, int relu, float* save_mean,
                 float* save_invstd, vfs_stream_t s) {
  return vfs::bn1d_act(y, M, N, gamma, beta, running_mean, running_var, eps, momentum, training, relu, save_mean,
                       save_invstd, s);
}
int vfs_relu(float* y, size_t n, vfs_stream_t s) { return vfs::relu_inplace(y, n, s); }
int vfs_cosine_sim_loss(const float* p, const float* z, float* loss, int B, int D, int with_norm, int negative,
                        vfs_stream_t s) {
  return vfs::cosine_sim_loss(p, z, loss, B, D, with_norm, negative, s);
}
int vfs_frames_u8_to_ncthw_f32(const unsigned char* frames, float* out, long long clips, int T, int H, int W,
                               const float* mean3, const double* stdinv3, int swap_rb, vfs_stream_t s) {
  return vfs::frames_u8_to_ncthw_f32(frames, out, clips, T, H, W, mean3, stdinv3, swap_rb, s);
}
int vfs_nchw_to_nhwc_f32(const float* in, float* out, int N, int C, int H, int W, vfs_stream_t s) {
  return vfs::nchw_to_nhwc_f32(in, out, N, C, H, W, s);
}
int vfs_xcorr_nhwc(const float* z, const float* x, float* out, int nz, int nx, int C, int hz, int wz, int h, int w,
                   float out_scale, vfs_stream_t s) {
  return vfs::xcorr_nhwc(z, x, out, nz, nx, C, hz, wz, h, w, out_scale, s);
}
int vfs_augment_u8_to_ncthw_f32(const VfsAugItem* items_dev, float* out, long long clips, int T, int dst_h, int dst_w,
                                const float* mean3, const double* stdinv3, int swap_rb, vfs_stream_t s) {
  return vfs::augment_u8_to_ncthw_f32(items_dev, out, clips, T, dst_h, dst_w, mean3, stdinv3, swap_rb, s);
}
int vfs_siamfc_loss(const float* responses, const float* labels, float* loss, float* grad, int n, int mode,
                    float gamma, float neg_weight, vfs_stream_t s) {
  return vfs::siamfc_loss(responses, labels, loss, grad, n, mode, gamma, neg_weight, s);
}
int vfs_xcorr_backward_nhwc(const float* dr, const float* z, const float* x, float* dz, float* dx, int n, int C, int hz,
                            int wz, int h, int w, float out_scale, vfs_stream_t s) {
  return vfs::xcorr_backward_nhwc(dr, z, x, dz, dx, n, C, hz, wz, h, w, out_scale, s);
}
int vfs_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
                  float weight_decay, int step, vfs_stream_t s) {
  return vfs::adam_step(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, s);
}

}  // extern "C"
