"""ctypes binding of the C ABI (include/dadetect_b200.h -> libdadetect_b200.so).

The library is built in-tree by ``__graft_entry__.build()`` (or ``make -C da-detect_b200/csrc``).
There is deliberately no fallback: if the shared object is missing or a call fails, an
exception is raised — the product path never routes through a CPU or library implementation.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdadetect_b200.so")

_P, _I, _F, _Q, _Z = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong, ctypes.c_size_t

# name -> (restype, argtypes)   p = pointer, i = int, f = float, q = long long, d = double
_SIGS = {
    "dd_last_error": (ctypes.c_char_p, ""),
    "dd_abi_version": (_I, ""),
    "dd_launch_count": (_Q, ""),
    "dd_set_sm_budget": (_I, "i"),
    "dd_tcgen05_built": (_I, ""),
    "dd_roi_align_forward": (_I, "pppiiiiifiiiip"),
    "dd_roi_align_backward": (_I, "pppiiiiifiiiip"),
    "dd_roi_align_forward_nchw": (_I, "pppiiiiifiiip"),
    "dd_roi_align_backward_nchw": (_I, "pppiiiiifiiip"),
    "dd_nms_workspace_bytes": (_Z, "i"),
    "dd_nms": (_I, "ppifpppp"),
    "dd_nms_sorted": (_I, "pifipppp"),
    "dd_nms_batched_workspace_bytes": (_Z, "ii"),
    "dd_nms_sorted_batched": (_I, "ppiifipippp"),
    "dd_anchor_grid": (_I, "piiiiiiippp"),
    "dd_rpn_topk_workspace_bytes": (_Z, "ii"),
    "dd_rpn_topk_decode": (_I, "pppiiiiiiifpppppp"),
    "dd_match": (_I, "pippiffipppp"),
    "dd_box_encode": (_I, "pipppiffffipp"),
    "dd_rpn_anchor_labels": (_I, "ppipp"),
    "dd_roi_labels": (_I, "ppipipp"),
    "dd_roi_gather_sampled": (_I, "ppppppppppiiiffffppppppp"),
    "dd_rpn_sampled_losses": (_I, "pppiiipppppppfpppp"),
    "dd_box_decode": (_I, "ppiiffffpp"),
    "dd_conv2d_forward_workspace_bytes": (_Z, "iiiii"),
    "dd_conv2d_forward": (_I, "ppppppiiiiiiiiiiipp"),
    "dd_conv2d_forward_prepared": (_I, "ppppppiiiiiiiiiiipp"),
    "dd_conv2d_forward_prepare_batch": (_I, "ippppppip"),
    "dd_stem_workspace_bytes": (_Z, "iiii"),
    "dd_stem_conv7x7s2_forward": (_I, "pppppiiiiiipp"),
    "dd_conv2d_dgrad_workspace_bytes": (_Z, "iiii"),
    "dd_conv2d_dgrad": (_I, "ppppppiiiiiiiiiipip"),
    "dd_conv2d_dgrad_prepare_batch": (_I, "ipppppppip"),
    "dd_conv2d_wgrad_workspace_bytes": (_Z, "iiiiiiiii"),
    "dd_conv2d_wgrad": (_I, "ppppiiiiiiiiiiipp"),
    "dd_bias_grad": (_I, "ppiiip"),
    "dd_nchw_to_nhwc": (_I, "ppiiiip"),
    "dd_nhwc_to_nchw": (_I, "ppiiiip"),
    "dd_maxpool3x3s2": (_I, "ppiiiip"),
    "dd_avgpool_forward": (_I, "ppiiip"),
    "dd_avgpool_backward": (_I, "ppiiip"),
    "dd_avgpool_relu_backward": (_I, "pppiiip"),
    "dd_relu_backward": (_I, "pppqp"),
    "dd_grl_backward": (_I, "pfpqip"),
    "dd_grl_backward_dev": (_I, "pppqip"),
    "dd_adv_grl_weight": (_I, "pffffpp"),
    "dd_dropout_apply": (_I, "pppqp"),
    "dd_bce_logits_mean": (_I, "pppqqppp"),
    "dd_softmax_ce_mean": (_I, "pppiippp"),
    "dd_smooth_l1_sum": (_I, "ppqffppp"),
    "dd_box_reg_loss": (_I, "ppppiippp"),
    "dd_consistency_loss": (_I, "pqpiipppppp"),
    "dd_proposals_gather": (_I, "ppppppppiiiipppp"),
    "dd_balanced_sample": (_I, "pppiiiippp"),
    "dd_sgd_momentum_dev": (_I, "pppqpffffp"),
    "dd_triplet_margin_loss": (_I, "pppqiqfpppppp"),
    "dd_adaptive_margin_update": (_I, "ppdddpp"),
    "dd_sgd_momentum": (_I, "pppqffffip"),
    "dd_upsample2x_forward": (_I, "ppiiiip"),
    "dd_upsample2x_backward": (_I, "ppiiiip"),
    "dd_subsample2_forward": (_I, "ppiiiip"),
    "dd_subsample2_backward": (_I, "ppiiiip"),
    "dd_fpn_level_map": (_I, "piiififpp"),
    "dd_roi_align_level_forward": (_I, "pppipiiiiifiiip"),
    "dd_roi_align_level_backward": (_I, "pppipiiiiifiiip"),
    "dd_resample_ksize": (_I, "ii"),
    "dd_resample_coeffs": (_I, "iipp"),
    "dd_preprocess_image": (_I, "piiiqppippiiiiipppiip"),
}
_CT = {"p": _P, "i": _I, "f": _F, "q": _Q, "d": ctypes.c_double}

EXPORTED_SYMBOLS = tuple(sorted(_SIGS))


class NativeError(RuntimeError):
    """A C-ABI call returned non-zero (cudaError or argument error)."""


_lib = None


def load():
    """Load libdadetect_b200.so and attach argtypes.  Raises ImportError when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "dadetect_b200: native library {} is missing — run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C da-detect_b200/csrc`).  There is no CPU fallback.".format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)            # AttributeError here = header/library drift
        fn.restype = res
        fn.argtypes = [_CT[c] for c in args]
    _lib = lib
    return lib


def call(name, *args):
    """Invoke an int-returning entry point; raise NativeError with dd_last_error() on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise NativeError("{} failed (code {}): {}".format(name, rc, lib.dd_last_error().decode()))


def launch_count():
    return int(load().dd_launch_count())
