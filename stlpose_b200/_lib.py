"""ctypes binding of libstlpose_b200.so (the C ABI declared in include/stlpose_b200.h).

There is no fallback: if the shared library is missing the import of any compute entry point raises,
and every entry point itself fails without a CUDA device.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# STLPOSE_LIB: another build of the same library (measurement builds, e.g. with -DSTL_CONV_COUNTERS)
LIB_PATH = os.environ.get("STLPOSE_LIB") or os.path.join(_HERE, "libstlpose_b200.so")
STL_MAX_UP = 3

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)
vp = ctypes.c_void_p


class StlError(RuntimeError):
    pass


class ConvDesc(ctypes.Structure):
    _fields_ = [
        ("in_", vp), ("N", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int), ("Cin", ctypes.c_int),
        ("out", vp), ("Cout", ctypes.c_int), ("Cout_pad", ctypes.c_int),
        ("ksize", ctypes.c_int), ("stride", ctypes.c_int),
        ("w_packed", vp), ("bias_packed", vp), ("residual", vp),
        ("n_up", ctypes.c_int), ("up_src", vp * STL_MAX_UP), ("up_shift", ctypes.c_int * STL_MAX_UP),
        ("relu", ctypes.c_int), ("out_nchw", ctypes.c_int), ("impl", ctypes.c_int),
        ("force_mb", ctypes.c_int), ("max_ctas", ctypes.c_int), ("dbg_counters", vp),
        ("in2", vp), ("Cin2", ctypes.c_int), ("pdl", ctypes.c_int),
    ]


class HrnetCfg(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int), ("num_joints", ctypes.c_int), ("stage_modules", ctypes.c_int * 3),
                ("blocks", ctypes.c_int), ("image_h", ctypes.c_int), ("image_w", ctypes.c_int)]


class ConvInfo(ctypes.Structure):
    _fields_ = [("conv_key", ctypes.c_char * 96), ("bn_key", ctypes.c_char * 96),
                ("cout", ctypes.c_int), ("cin", ctypes.c_int), ("ksize", ctypes.c_int), ("stride", ctypes.c_int)]


class OpInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("kind", "layer", "cin", "cout", "ksize", "stride", "out_h", "out_w")] + \
               [("flops_per_image", ctypes.c_double), ("bytes_per_image", ctypes.c_double)] + \
               [(n, ctypes.c_int) for n in ("grid", "smem", "mb", "nt", "ck", "a_stages", "b_stages", "a_shift",
                                            "tiles", "subs")]


# name -> (restype, argtypes); mirrors include/stlpose_b200.h one to one
PROTOTYPES = {
    "stl_abi_version": (ctypes.c_int, []),
    "stl_last_error": (ctypes.c_char_p, []),
    "stl_flip_avg": (ctypes.c_int, [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_int_p,
                                    ctypes.c_int, vp]),
    "stl_flip_back": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_int_p,
                                     ctypes.c_int, vp]),
    "stl_decode": (ctypes.c_int, [vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_int_p,
                                  ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp]),
    "stl_mse_workspace_bytes": (ctypes.c_size_t, []),
    "stl_oks_nms": (ctypes.c_int, [vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.c_float,
                                   ctypes.c_double, ctypes.c_float, ctypes.c_int, vp, vp, vp]),
    "stl_generate_target": (ctypes.c_int, [vp, vp, vp] + [ctypes.c_int] * 7 + [vp, vp, vp]),
    "stl_upsampled_argmax": (ctypes.c_int, [vp] + [ctypes.c_int] * 6 + [vp, vp, vp]),
    "stl_warp_affine_crops": (ctypes.c_int, [vp] + [ctypes.c_int] * 2 + [vp] + [ctypes.c_int] * 3 + [vp, vp, vp, vp, vp]),
    "stl_sgd_step_batched": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                            ctypes.c_float, ctypes.c_int, vp]),
    "stl_stem_im2col": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]),
    "stl_warp_affine_crops_f32": (ctypes.c_int, [vp] + [ctypes.c_int] * 2 + [vp] + [ctypes.c_int] * 3 + [vp, vp]),
    "stl_pck_accuracy": (ctypes.c_int, [vp, vp] + [ctypes.c_int] * 4 + [ctypes.c_float, vp, vp, vp, vp]),
    "stl_scale_inplace": (ctypes.c_int, [vp, vp, ctypes.c_longlong, vp]),
    "stl_mse_loss_fwd_bwd": (ctypes.c_int, [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]),
    "stl_padded_bytes": (ctypes.c_size_t, [ctypes.c_int] * 4),
    "stl_nchw_to_padded": (ctypes.c_int, [vp, vp] + [ctypes.c_int] * 5 + [vp]),
    "stl_padded_to_nchw": (ctypes.c_int, [vp, vp] + [ctypes.c_int] * 5 + [vp]),
    "stl_pack_conv_weights": (ctypes.c_int, [vp] * 6 + [ctypes.c_float] + [ctypes.c_int] * 5 + [vp, vp, vp]),
    "stl_basic_block": (ctypes.c_int, [vp] * 6 + [ctypes.c_int] * 4 + [vp]),
    "stl_bottleneck_link": (ctypes.c_int, [vp] * 8 + [ctypes.c_int] * 4 + [vp]),
    "stl_bottleneck_link2": (ctypes.c_int, [vp] * 8 + [ctypes.c_int] * 4 + [vp]),
    "stl_pack_conv_weights_dgrad": (ctypes.c_int, [vp] + [ctypes.c_int] * 5 + [vp, vp, vp]),
    "stl_conv2d": (ctypes.c_int, [ctypes.POINTER(ConvDesc), vp]),
    "stl_conv2d_stats_floats": (ctypes.c_size_t, [ctypes.c_int]),
    "stl_conv2d_stats": (ctypes.c_int, [ctypes.POINTER(ConvDesc), vp, c_int_p, vp]),
    "stl_conv2d_bn": (ctypes.c_int, [ctypes.POINTER(ConvDesc), vp, vp, ctypes.c_float, ctypes.c_float, vp, vp, vp, vp,
                                     c_int_p, vp]),
    "stl_bn_apply": (ctypes.c_int, [vp] * 6 + [ctypes.c_int] * 5 + [vp, vp]),
    "stl_bn_train_forward_fused": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, ctypes.c_int,
                                                  ctypes.c_float, ctypes.c_float] + [ctypes.c_int] * 4 + [vp] * 6),
    "stl_plan_create": (vp, [ctypes.POINTER(HrnetCfg)]),
    "stl_plan_destroy": (None, [vp]),
    "stl_plan_num_convs": (ctypes.c_int, [vp]),
    "stl_plan_conv_info": (ctypes.c_int, [vp, ctypes.c_int, ctypes.POINTER(ConvInfo)]),
    "stl_plan_weight_bytes": (ctypes.c_size_t, [vp]),
    "stl_plan_pack_conv": (ctypes.c_int, [vp, ctypes.c_int] + [vp] * 6 + [ctypes.c_float, vp, vp]),
    "stl_plan_workspace_bytes": (ctypes.c_size_t, [vp, ctypes.c_int]),
    "stl_plan_forward": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, ctypes.c_size_t, vp]),
    "stl_plan_launches_per_forward": (ctypes.c_int, [vp]),
    "stl_plan_kernel_launches": (ctypes.c_int, [vp]),
    "stl_plan_forward_timed": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, ctypes.c_size_t, vp,
                                              c_float_p]),
    "stl_plan_op_info": (ctypes.c_int, [vp, ctypes.c_int, ctypes.POINTER(OpInfo)]),
    "stl_pack_conv_weights_batched": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, vp]),
    "stl_bn_train_forward": (ctypes.c_int, [vp, vp, vp, vp, ctypes.c_int, ctypes.c_float, ctypes.c_float] +
                             [ctypes.c_int] * 4 + [vp] * 7),
    "stl_bn_train_backward": (ctypes.c_int, [vp] * 6 + [ctypes.c_int] * 5 + [vp] * 4),
    "stl_bn_train_forward_ticket": (ctypes.c_int, [vp, vp, vp, vp, ctypes.c_int, ctypes.c_float, ctypes.c_float] +
                                    [ctypes.c_int] * 4 + [vp] * 8),
    "stl_bn_train_backward_ticket": (ctypes.c_int, [vp] * 6 + [ctypes.c_int] * 5 + [vp] * 6),
    "stl_bn_train_forward_coop": (ctypes.c_int, [vp, vp, vp, vp, ctypes.c_int, ctypes.c_float, ctypes.c_float] +
                                  [ctypes.c_int] * 4 + [vp] * 8 + [vp]),
    "stl_bn_train_backward_coop": (ctypes.c_int, [vp] * 7 + [ctypes.c_int] * 5 + [vp] * 6 + [vp]),
    "stl_bn_train_backward_ticket_z": (ctypes.c_int, [vp] * 6 + [ctypes.c_int] * 4 + [vp] * 5),
    "stl_sum_relu_forward": (ctypes.c_int, [ctypes.POINTER(vp), ctypes.c_int, ctypes.POINTER(vp), c_int_p, ctypes.c_int,
                                            vp] + [ctypes.c_int] * 4 + [vp]),
    "stl_relu_mask": (ctypes.c_int, [vp, vp, vp, ctypes.c_longlong, vp]),
    "stl_upsample_backward": (ctypes.c_int, [vp, vp] + [ctypes.c_int] * 5 + [vp]),
    "stl_conv_dgrad": (ctypes.c_int, [vp, vp, vp] + [ctypes.c_int] * 7 + [vp]),
    "stl_conv_wgrad_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 8),
    "stl_conv_wgrad": (ctypes.c_int, [vp, vp, vp] + [ctypes.c_int] * 8 + [vp, ctypes.c_size_t, vp]),
    "stl_bn_workspace_floats": (ctypes.c_size_t, [ctypes.c_int]),
    "stl_zero_stuff": (ctypes.c_int, [vp, vp] + [ctypes.c_int] * 4 + [vp]),
    "stl_conv_wgrad_naive_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 8),
    "stl_conv_wgrad_naive": (ctypes.c_int, [vp, vp, vp] + [ctypes.c_int] * 8 + [vp, ctypes.c_size_t, vp]),
}

_lib = None


def lib():
    """Load the shared library (once). Raises StlError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise StlError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           f"or `make -C stlpose_b200/csrc`. There is no CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.stl_abi_version() != 1:
            raise StlError("libstlpose_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(status):
    if status != 0:
        raise StlError(lib().stl_last_error().decode() or "stlpose_b200 call failed")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
