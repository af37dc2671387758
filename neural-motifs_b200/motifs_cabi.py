"""ctypes binding of libmotifs_b200.so (the C ABI declared in include/motifs_b200.h).

The product path has NO fallback: if the shared library is missing, was built for another
architecture, or a call fails, an exception is raised.  Nothing here imports `oracle/`.
"""
import ctypes
import os
from ctypes import c_int, c_float, c_double, c_void_p, c_longlong, c_size_t, c_char_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmotifs_b200.so")

P = c_void_p  # every device / host pointer crosses the boundary as a raw address

# name -> (restype, argtypes); mirrors include/motifs_b200.h one to one.
SIGNATURES = {
    "ROIAlignForwardLaucher": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P, P]),
    "ROIAlignBackwardLaucher": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P]),
    "ApplyNMSGPU": (c_int, [P, P, c_int, c_float, c_int]),
    "highway_lstm_forward_ongpu": (None, [c_int] * 5 + [P] * 10 + [c_int, P, P]),
    "highway_lstm_backward_ongpu": (None, [c_int] * 5 + [P] * 16 + [c_int, c_int, P, P]),
    "mb200_last_error": (c_char_p, []),
    "mb200_abi_version": (c_int, []),
    "mb200_compiled_arch": (c_int, []),
    "mb200_device_ok": (c_int, []),
    "mb200_roi_align_forward_nhwc": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P, P]),
    "mb200_roi_align_forward_nhwc_to_nchw": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P, P]),
    "mb200_highway_lstm_layer_forward": (c_int, [c_int] * 4 + [P] * 9),
    "mb200_highway_lstm_layer_backward": (c_int, [c_int] * 4 + [P] * 11),
    "mb200_nms_mask_words": (c_longlong, [P, c_int]),
    "mb200_nms_segmented": (c_int, [P, P, P, c_int, c_int, c_float, c_int, P, P, P, P]),
    "mb200_bbox_overlaps_f32": (c_int, [P, c_int, P, c_int, P, P]),
    "mb200_bbox_overlaps_f64": (c_int, [P, c_int, P, c_int, c_int, P, P]),
    "mb200_union_rois": (c_int, [P, P, c_int, P, P, P]),
    "mb200_draw_union_boxes": (c_int, [P, c_int, c_int, c_float, P, P]),
    "mb200_bbox_preds": (c_int, [P, P, c_longlong, c_int, P, P, P, P]),
    "mb200_highway_lstm_scratch_floats": (c_size_t, [c_int, c_int, c_int]),
    "mb200_highway_lstm_forward": (c_int, [c_int] * 5 + [P] * 9 + [P]),
    "mb200_highway_lstm_backward": (c_int, [c_int] * 5 + [P] * 14 + [c_int, P, P]),
    "mb200_gemm_workspace_floats": (c_longlong, [c_int, c_int, c_int]),
    "mb200_gemm_bf16x3": (c_int, [P, P, P, P, c_int, c_int, c_int, P, c_int, P, c_longlong, P, P, c_longlong, P, P]),
    "mb200_conv3x3_bf16x3": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, P, c_int, P, P, P, P]),
    "mb200_split_bf16": (c_int, [P, c_longlong, c_int, c_longlong, c_int, P, P, P]),
    "mb200_split_transpose_bf16": (c_int, [P, c_int, c_int, c_longlong, c_int, P, P, P]),
    "mb200_conv_weight_split": (c_int, [P, c_int, c_int, c_int, P, P, P]),
    "mb200_im2col3_split": (c_int, [P, c_int, c_int, c_int, P, P, P]),
    "mb200_stem_weight_split": (c_int, [P, c_int, P, P, P]),
    "mb200_conv3x3_stem_split": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "mb200_maxpool2_nhwc_split": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P, P]),
    "mb200_maxpool3s2_forward": (c_int, [P, c_longlong, c_int, c_int, P, P, P]),
    "mb200_maxpool3s2_backward": (c_int, [P, P, c_longlong, c_int, c_int, P, P]),
    "mb200_im2col7s2_split": (c_int, [P, c_int, c_int, c_int, c_longlong, P, P, P]),
    "mb200_im2col3_nhwc_split": (c_int, [P, c_int, c_int, c_int, c_int, c_int, c_longlong, P, P, P]),
    "mb200_bn_stats": (c_int, [P, c_longlong, c_int, c_float, c_float, P, P, P, P, P, P]),
    "mb200_bn_pool3s2_nhwc": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, P, P, P]),
    "mb200_unpool3s2_nhwc": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P]),
    "mb200_bn_nhwc_to_nchw": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, P, P]),
    "mb200_nchw_to_nhwc": (c_int, [P, c_int, c_int, c_int, P, P]),
    "mb200_bn_relu_backward": (c_int, [P, P, P, P, P, c_longlong, c_int, P, P, P, P]),
    "mb200_bn_relu_backward_split": (c_int, [P, P, P, P, P, P, c_longlong, c_longlong, c_int, c_int, c_int, P, P, P, P, P, P, P]),
    "mb200_col2im3_nhwc": (c_int, [P, c_int, c_int, c_int, c_int, P, P]),
    "mb200_gemm_mn_workspace_floats": (c_longlong, [c_int, c_int, c_int]),
    "mb200_gemm_bf16x3_mn": (c_int, [P, P, c_longlong, P, P, c_longlong, c_int, c_int, c_int, P, c_longlong, P, P]),
    "mb200_sumsq_accum": (c_int, [P, c_longlong, P, P]),
    "mb200_sgd_momentum_clip": (c_int, [P, P, P, c_longlong, c_float, c_float, c_float, P, c_float, c_int, c_int, P]),
    "mb200_sgd_momentum_clip_scaled": (c_int, [P, P, P, c_longlong, c_float, c_float, c_float, P, c_float, c_float, c_int, c_int, P]),
    "mb200_anchor_targets": (c_int, [P, c_int, P, c_int, c_double, c_double, P, P, P, P, P]),
    "mb200_gemm_set_pair_mode": (c_int, [c_int]),
    "mb200_set_sm_budget": (c_int, [c_int]),
    "mb200_zero_async": (c_int, [P, c_longlong, P]),
    "mb200_conv_set_halo_mode": (c_int, [c_int]),
    "mb200_decoder_commit": (c_int, [P, P, c_int, c_int, c_float, P, P]),
    "mb200_sgd_momentum_clip_split": (c_int, [P, P, P, P, P, c_longlong, c_float, c_float, c_float, P, c_float, c_float, c_int, c_int, P]),
    "mb200_optim_set_background": (c_int, [c_int]),
    "mb200_dp_reduce_shard_sumsq": (c_int, [P, P, c_longlong, P, P]),
    "mb200_dp_bcast_slot": (c_int, [P, P, c_int, P]),
    "mb200_dp_reduce_staged_sumsq": (c_int, [P, P, c_longlong, c_int, c_longlong, P, P]),
    "mb200_sgd_momentum_clip_mc": (c_int, [P, P, P, P, P, P, c_longlong, c_float, c_float, c_float, P, c_float, c_float, c_int, c_int, P]),
    "mb200_highway_lstm_tc_supported": (c_int, [c_int, c_int]),
    "mb200_highway_lstm_layer_forward_tc": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P, P, P, P, P]),
    "mb200_sgemm": (c_int, [c_int, c_int, c_int, c_int, c_int, c_float, P, c_int, P, c_int, c_float, P, c_int, P]),
}

_lib = None


class MotifsB200Error(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle; raises MotifsB200Error if it cannot."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MotifsB200Error(
            "libmotifs_b200.so not found at %s — run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or library fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise MotifsB200Error("libmotifs_b200.so does not export %s" % name) from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    mode = os.environ.get("MOTIFS_GEMM_PAIR_MODE")      # A/B runs: 0 = 1-CTA tcgen05 kernels only, 1 = per shape (default), 2 = force pairs
    if mode is not None:
        lib.mb200_gemm_set_pair_mode(int(mode))
    mode = os.environ.get("MOTIFS_CONV_HALO_MODE")      # 0 = tap-by-tap conv kernels, 1 = halo kernel (default)
    if mode is not None:
        lib.mb200_conv_set_halo_mode(int(mode))
    return lib


def last_error():
    return load().mb200_last_error().decode()


LAUNCHER_CALLS = 0   # successful C-ABI launcher calls (each launches >= 1 kernel of this library)


def check(rc, what):
    """The launchers return 1 on success (reference convention); anything else raises."""
    global LAUNCHER_CALLS
    LAUNCHER_CALLS += 1
    if rc != 1:
        raise MotifsB200Error("%s failed (code %d): %s" % (what, rc, last_error()))


def ptr(t):
    """Raw address of a torch tensor (or None)."""
    if t is None:
        return None
    return c_void_p(t.data_ptr())


def cur_stream():
    """Raw cudaStream_t of torch's current stream on the current device (fast path: no Stream object)."""
    import torch
    return c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise MotifsB200Error("motifs_b200 operators run on CUDA tensors only (got a %s tensor); "
                                  "there is no CPU path" % t.device)
