"""ctypes binding of libu2b200.so (the C ABI declared in include/u2b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the
product path raises. Build with `python -m u2seg_b200.build` (nvcc, sm_100a).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libu2b200.so")

_lib = None

c_void_p, c_int, c_int64, c_size_t, c_float = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_float)

# name -> (restype, argtypes); must list every symbol include/u2b200.h declares
# (tests/test_abi.py parses the header and checks this table and the .so against it).
SIGNATURES = {
    "u2b_last_error": (ctypes.c_char_p, []),
    "u2b_version": (c_int, []),
    "u2b_sm_count": (c_int, []),
    "u2b_kmeans_kpad": (c_int64, [c_int64]),
    "u2b_kmeans_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "u2b_kmeans_xnorm_max": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "u2b_kmeans_prepare": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "u2b_kmeans_assign": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "u2b_kmeans_set_cluster": (c_int, [c_int]),
    "u2b_kmeans_set_mstep": (c_int, [c_int]),
    "u2b_knn_candidates_per_row": (c_int, []),
    "u2b_knn_npad": (c_int64, [c_int64]),
    "u2b_knn_set_cluster": (c_int, [c_int]),
    "u2b_knn_candidates": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                   c_void_p]),
    "u2b_knn_refine": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_float,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "u2b_kmeans_accumulate": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                      c_size_t, c_void_p]),
    "u2b_kmeans_finalize": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "u2b_assign_levels": (c_int, [c_void_p, c_int64, c_int, c_int, c_float, c_int, c_void_p, c_void_p]),
    "u2b_roi_align_fwd": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                  c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "u2b_roi_align_bwd": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                  c_void_p, c_int64, c_int, c_void_p, c_float, c_void_p]),
    "u2b_roi_align_chw_supported": (c_int, [c_int64, c_int]),
    "u2b_roi_align_fwd_chw": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                      c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "u2b_roi_align_bwd_chw": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                      c_void_p, c_int64, c_int, c_void_p, c_float, c_void_p]),
    "u2b_paste_masks": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "u2b_crop_resize_masks": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p,
                                      c_void_p, c_void_p]),
    "u2b_gn_supported": (c_int, [c_int, c_int]),
    "u2b_gn_finalize": (c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p]),
    "u2b_gn_bwd_coeff": (c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                 c_void_p]),
    "u2b_stem_conv_supported": (c_int, [c_int] * 6),
    "u2b_stem_conv_fwd": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "u2b_stem_conv_wgrad_num_partials": (c_int, [c_int64, c_int, c_int]),
    "u2b_stem_conv_wgrad": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "u2b_conv2d_wgrad_supported": (c_int, [c_int] * 6),
    "u2b_conv2d_wgrad_ksplit": (c_int, [c_int] * 9),
    "u2b_conv2d_nhwc_wgrad": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_int, c_void_p, c_void_p]),
    "u2b_upsample_bilinear_supported": (c_int, [c_int, c_int]),
    "u2b_upsample_bilinear": (c_int, [c_int, c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p]),
    "u2b_rpn_losses_num_partials": (c_int64, [c_int64]),
    "u2b_box_losses_num_partials": (c_int64, [c_int64]),
    "u2b_rpn_losses": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "u2b_box_losses": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p,
                               c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "u2b_rpn_decode_selected": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p,
                                        c_float, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "u2b_cascade_relabel": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_float, c_float,
                                    c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "u2b_sgd_step_segments": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_float, c_int, c_int64, c_void_p]),
    "u2b_debug_nms_profile": (c_int, [c_int, c_void_p]),
    "u2b_upsample_ce_num_partials": (c_int64, [c_int64, c_int, c_int]),
    "u2b_upsample_ce_supported": (c_int, [c_int, c_int]),
    "u2b_upsample_ce": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int64, c_void_p,
                                c_void_p, c_void_p]),
    "u2b_iou_match": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_void_p]),
    "u2b_conv2d_supported": (c_int, [c_int] * 6),
    "u2b_conv2d_set_cluster": (c_int, [c_int]),
    "u2b_conv2d_nhwc_fwd": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                    c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "u2b_conv2_supported": (c_int, [c_int] * 6),
    "u2b_conv2_stats_rows": (c_int64, [c_int] * 7),
    "u2b_conv2_set_tile_n": (c_int, [c_int]),
    "u2b_conv2_nhwc_fwd": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                   c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "u2b_conv2_dgrad_supported": (c_int, [c_int] * 6),
    "u2b_conv2_nhwc_dgrad": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p]),
    "u2b_deconv2x2_supported": (c_int, [c_int, c_int]),
    "u2b_deconv2x2_nhwc_fwd": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int,
                                       c_void_p, c_void_p]),
    "u2b_conv_wgrad2_supported": (c_int, [c_int] * 6),
    "u2b_conv_wgrad2_workspace_floats": (c_int64, [c_int] * 9),
    "u2b_conv_wgrad2": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_void_p, c_int, c_void_p, c_void_p]),
    "u2b_mask_loss_supported": (c_int, [c_int]),
    "u2b_mask_loss_num_partials": (c_int, []),
    "u2b_mask_loss_fwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int,
                                  c_void_p, c_void_p, c_void_p]),
    "u2b_mask_loss_bwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    "u2b_maxpool3x3s2_fwd": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "u2b_maxpool3x3s2_bwd": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "u2b_sum2x2_nhwc": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "u2b_bn_apply_resup": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "u2b_preprocess_u8_nhwc": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_float),
                                       ctypes.POINTER(c_float), c_int, c_void_p, c_void_p]),
    "u2b_bn_supported": (c_int, [c_int]),
    "u2b_bn_num_strips": (c_int, [c_int64, c_int]),
    "u2b_bn_stats": (c_int, [c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "u2b_bn_sum_partials": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "u2b_bn_finalize": (c_int, [c_void_p, c_int, ctypes.c_double, c_void_p, c_void_p, c_float, c_float, c_void_p,
                                c_void_p, c_void_p, c_int, c_void_p]),
    "u2b_bn_apply": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int, c_void_p]),
    "u2b_bn_bwd_reduce": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "u2b_bn_bwd_reduce_relu_x": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "u2b_bn_bwd_apply_relu_x": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "u2b_bn_bwd_coeff": (c_int, [c_void_p, c_int, ctypes.c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                 c_void_p]),
    "u2b_bn_xchg_buffer_bytes": (c_size_t, [c_int, c_int]),
    "u2b_bn_xchg_finalize": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, ctypes.c_double, c_void_p,
                                     c_void_p, c_float, c_float, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "u2b_bn_xchg_bwd_coeff": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, ctypes.c_double, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "u2b_roi_align_set_impl": (c_int, [c_int]),
    "u2b_set_sm_budget": (c_int, [c_int]),
    "u2b_set_pdl": (c_int, [c_int]),
    "u2b_conv2_set_staging": (c_int, [c_int]),
    "u2b_bn_xchg2_max_ctas": (c_int, []),
    "u2b_bn_xchg2_finalize": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, ctypes.c_double, c_void_p,
                                      c_void_p, c_float, c_float, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "u2b_bn_xchg2_bwd_coeff": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, ctypes.c_double, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "u2b_bn_bwd_apply": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                 c_void_p]),
    "u2b_nms_workspace_bytes": (c_size_t, [c_int64]),
    "u2b_batched_nms": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_int64, c_void_p, c_void_p,
                                c_void_p, c_size_t, c_void_p]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libu2b200.so not found at %s — build it with `python -m u2seg_b200.build` "
                "(no CPU or PyTorch fallback exists for the B200 path)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if os.environ.get("U2B_CONV2_STAGING", "0") == "1":   # csrc/conv2.cu conv2_pick_pipeline (A/B timing; default off)
            L.u2b_conv2_set_staging(1)
        if os.environ.get("U2B_PDL", "1") == "0":      # programmatic dependent launch between libu2b200 kernels (common.cuh)
            L.u2b_set_pdl(0)
        _lib = L
    return _lib


class U2BError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = lib().u2b_last_error()
        raise U2BError("%s failed (code %d): %s" % (what or "libu2b200 call", rc,
                                                    msg.decode() if msg else "?"))


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL). Tensors must be contiguous."""
    if t is None:
        return ctypes.c_void_p(0)
    assert t.is_contiguous(), "libu2b200 expects contiguous tensors"
    return ctypes.c_void_p(t.data_ptr())


# number of kernels launched through the library since import (bench.py's `gpu_launches`)
launch_count = 0


def count_launches(n):
    global launch_count
    launch_count += n
