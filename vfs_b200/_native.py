"""ctypes binding of libvfs_b200.so (the C ABI declared in include/vfs_b200.h).

There is deliberately NO fallback: if the library is missing or a call fails, a RuntimeError is raised.
torch is only used by callers for device memory and streams; the ABI itself is torch-free.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_lib', 'libvfs_b200.so')

VFS_OK = 0


class VfsConvDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32)
                for n in ('N', 'H', 'W', 'Cin', 'Cout', 'ksize', 'stride', 'dilation', 'relu')]


class VfsPackItem(ctypes.Structure):
    _fields_ = [('w', ctypes.c_void_p), ('dst_split', ctypes.c_void_p)] + \
        [(n, ctypes.c_int32) for n in ('Cout', 'Cin', 'ksize', 'mode', 'first_block', 'scale_log2')]


class VfsAugItem(ctypes.Structure):
    _fields_ = [('src', ctypes.c_void_p)] + \
        [(n, ctypes.c_int32) for n in ('H', 'W', 'crop_x0', 'crop_y0', 'crop_w', 'crop_h', 'flip', 'reserved')]


class VfsAttnDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ('H', 'W', 'C', 'Cv', 'T', 'topk', 'mask_mode', 'radius_y', 'radius_x',
                                              'non_mask_len', 'mode')] + [('temperature', ctypes.c_float)]


_vp, _i, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
_sz, _ll = ctypes.c_size_t, ctypes.c_longlong
_attn_p = ctypes.POINTER(VfsAttnDesc)

# name -> (restype, argtypes); every symbol include/vfs_b200.h declares
PROTOTYPES = {
    'vfs_last_error_string': (ctypes.c_char_p, []),
    'vfs_abi_version': (_i, []),
    'vfs_check_device': (_i, []),
    'vfs_overflow_count': (ctypes.c_uint, [_i]),
    'vfs_nchw_f32_to_split': (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    'vfs_nchw_f32_to_split_scaled': (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp]),
    'vfs_split_to_nchw_f32': (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    'vfs_stem_workspace_bytes': (_sz, [_i, _i, _i]),
    'vfs_stem_packed_weight_bytes': (_sz, []),
    'vfs_stem_pack_weight': (_i, [_vp, _vp, _vp]),
    'vfs_stem_forward': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    'vfs_conv_bn_act': (_i, [ctypes.POINTER(VfsConvDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'vfs_pack_conv_weight': (_i, [_vp, _vp, _i, _i, _i, _vp]),
    'vfs_pack_conv_weight_scaled': (_i, [_vp, _vp, _i, _i, _i, _f, _vp]),
    'vfs_pack_conv_weight_dgrad_scaled': (_i, [_vp, _vp, _i, _i, _i, _f, _vp]),
    'vfs_pack_blocks': (_i, [_i, _i, _i]),
    'vfs_pack_conv_weights_multi': (_i, [_vp, _i, _i, _vp]),
    'vfs_debug_conv_trace': (_i, [_vp, _i]),
    'vfs_conv_set_pair_policy': (_i, [_i, _i]),
    'vfs_attention_set_wide': (_i, [_i]),
    'vfs_debug_conv_bn_act_simt': (_i, [ctypes.POINTER(VfsConvDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'vfs_stem_conv_raw': (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    'vfs_stem_bn_relu_pool': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    'vfs_conv_stats': (_i, [ctypes.POINTER(VfsConvDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'vfs_conv_dgrad': (_i, [ctypes.POINTER(VfsConvDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'vfs_pack_conv_weight_dgrad': (_i, [_vp, _vp, _i, _i, _i, _vp]),
    'vfs_conv_wgrad_workspace_bytes': (_sz, [_i, _i, _i]),
    'vfs_conv_wgrad': (_i, [ctypes.POINTER(VfsConvDesc), _vp, _vp, _vp, _vp, _i, _f, _vp]),
    'vfs_affine_act_f32': (_i, [_vp, _vp, _vp, _ll, _i, _i, _vp]),
    'vfs_bn_bwd_reduce': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _vp]),
    'vfs_bn_bwd_apply': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_double, _vp, _vp, _vp, _vp,
                              _vp, _i, _f, _ll, _i, _vp]),
    'vfs_conv_stats_split': (_i, [ctypes.POINTER(VfsConvDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'vfs_relu_bwd_split': (_i, [_vp, _vp, _vp, _ll, _vp]),
    'vfs_stem_pool_relu_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    'vfs_stem_wgrad': (_i, [_vp, _vp, _vp, _i, _f, _i, _i, _i, _vp]),
    'vfs_linear_backward': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'vfs_bn1d_backward': (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _i, _vp]),
    'vfs_relu_backward': (_i, [_vp, _vp, _vp, _sz, _vp]),
    'vfs_avgpool_backward': (_i, [_vp, _vp, _i, _i, _i, _vp]),
    'vfs_cosine_loss_backward': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'vfs_sgd_momentum_step': (_i, [_vp, _vp, _vp, _sz, _f, _f, _f, _i, _f, _vp]),
    'vfs_sgd_momentum_step_dev': (_i, [_vp, _vp, _vp, _sz, _vp, _vp]),
    'vfs_comm_handle_bytes': (_sz, []),
    'vfs_comm_create': (_i, [_i, _i, _sz, ctypes.POINTER(ctypes.c_void_p), _vp]),
    'vfs_comm_connect': (_i, [_vp, _vp]),
    'vfs_comm_destroy': (_i, [_vp]),
    'vfs_comm_data_ptr': (_vp, [_vp]),
    'vfs_comm_data_bytes': (_sz, [_vp]),
    'vfs_comm_error': (_i, [_vp]),
    'vfs_comm_allreduce_small_f64': (_i, [_vp, _vp, _i, _vp]),
    'vfs_comm_allreduce_small_f32': (_i, [_vp, _vp, _i, _vp]),
    'vfs_comm_barrier': (_i, [_vp, _vp]),
    'vfs_comm_allreduce_f32': (_i, [_vp, _sz, _sz, _f, _vp]),
    'vfs_channel_stats_f32': (_i, [_vp, _vp, _ll, _i, _vp]),
    'vfs_bn_finalize': (_i, [_vp, ctypes.c_double, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _i, _vp]),
    'vfs_bn_apply': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _vp]),
    'vfs_features_to_split': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'vfs_features_to_split_ex': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _ll, _vp]),
    'vfs_seg_postprocess_workspace_bytes': (_sz, [_i]),
    'vfs_seg_postprocess': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'vfs_generic_attention': (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _i, _i, _i, _vp, _vp]),
    'vfs_siamfc_peak_workspace_bytes': (_sz, [_i, _i]),
    'vfs_siamfc_response_peak': (_i, [_vp, _i, _i, _i, _vp, _f, _f, _vp, _vp, _vp]),
    'vfs_seg_postprocess_batched': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    'vfs_masked_softmax': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    'vfs_propagate_dense': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'vfs_normalize_split': (_i, [_vp, _vp, _ll, _i, _ll, _ll, _vp]),
    'vfs_attention_workspace_bytes': (_sz, [_attn_p, _i]),
    'vfs_masked_attention_batched': (_i, [_attn_p, _i, _vp, _ll, _i, ctypes.POINTER(ctypes.c_int32), _vp, _ll, _i,
                                          ctypes.POINTER(ctypes.c_int32), _vp, ctypes.POINTER(ctypes.c_int32), _ll,
                                          _ll, _ll, _vp, _vp, _vp, _vp, _sz, _vp]),
    'vfs_masked_attention': (_i, [_attn_p, _vp, _ll, _vp, _ll, _i, ctypes.POINTER(ctypes.c_int32), _vp, _ll, _ll,
                                  _vp, _vp, _vp, _vp, _sz, _vp]),
    'vfs_global_avg_pool': (_i, [_vp, _vp, _i, _i, _i, _vp]),
    'vfs_linear': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    'vfs_bn1d_act': (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _f, _f, _i, _i, _vp, _vp, _vp]),
    'vfs_relu': (_i, [_vp, _sz, _vp]),
    'vfs_cosine_sim_loss': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'vfs_nchw_to_nhwc_f32': (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    'vfs_frames_u8_to_ncthw_f32': (_i, [_vp, _vp, _ll, _i, _i, _i, ctypes.POINTER(ctypes.c_float),
                                        ctypes.POINTER(ctypes.c_double), _i, _vp]),
    'vfs_augment_u8_to_ncthw_f32': (_i, [_vp, _vp, _ll, _i, _i, _i, ctypes.POINTER(ctypes.c_float),
                                         ctypes.POINTER(ctypes.c_double), _i, _vp]),
    'vfs_siamfc_loss': (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _f, _vp]),
    'vfs_xcorr_backward_nhwc': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp]),
    'vfs_adam_step': (_i, [_vp, _vp, _vp, _vp, _sz, _f, _f, _f, _f, _f, _i, _vp]),
    'vfs_xcorr_nhwc': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp]),
}

_lib = None


def lib():
    """Load (once) and return the ctypes handle; raises if the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} not found: build it with `python -m vfs_b200.build` (no CPU fallback exists)')
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what=''):
    if rc != VFS_OK:
        msg = lib().vfs_last_error_string().decode(errors='replace')
        raise RuntimeError(f'libvfs_b200 {what} failed with code {rc}: {msg}')


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


_raw_stream = None


def current_stream():
    """Raw cudaStream_t of torch's current stream on the current device.  ``torch.cuda.current_stream()`` builds a
    Python Stream object (~20 us per call, measured in tools/e2e_profile.py) -- per kernel launch that was almost half
    of the eager host time; the private raw-stream getter is a plain C call."""
    global _raw_stream
    import torch
    if _raw_stream is None:
        _raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', False)
    if _raw_stream:
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
