"""ctypes binding of libv2f_b200.so (the C ABI declared in include/v2f.h).

There is no CPU or eager fallback: if the shared library is missing, or a tensor is not a
contiguous CUDA tensor of the expected dtype, the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libv2f_b200.so")
ABI_VERSION = 4

_lib = None

c_int, c_ll, c_float, c_vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p

_ERRORS = {-1: "bad argument", -2: "pointer not 16-byte aligned", -3: "kernel launch failed",
           -4: "unsupported size"}


class DecodeParams(ctypes.Structure):
    """Mirror of ``struct v2f_decode_params`` (include/v2f.h), field for field."""
    _INTS = ["N", "B", "W", "E", "H", "Li", "Lt", "T", "variant", "mod_mask"]
    _PTRS = ["Himg", "Vimg", "Htr", "Ptr", "Mst", "HMst", "h0", "x0", "y",
             "Wcat", "bcat", "w_att", "beta_att", "b_tl", "We_mm", "W_me", "b_me", "W_ihc", "w_x",
             "b_ih", "w_fc", "b_fc",
             "yhat", "h_all", "S_all", "alpha_img", "alpha_tr", "alpha_mm", "C", "HC", "U", "CTX",
             "GI", "RZN", "xin",
             "dY", "dh", "DScat", "DGI", "DCTX", "dU", "DHC", "DC", "DE_img", "DE_tr", "DYH", "dxn",
             "dw_acc", "dMst_acc", "dHMst_acc", "dHimg", "dVimg", "dHtr", "dPtr", "dMst", "dHMst",
             "dWcat", "dbcat", "dw_att", "db_tl", "dWe_mm", "dW_me", "db_me", "dW_ihc", "dw_x",
             "db_ih", "dw_fc", "db_fc", "WcatT", "W_ihcT", "W_meT", "We_mmT", "ws"]
    _fields_ = ([(n, c_int) for n in _INTS] + [("tf_mask", ctypes.c_uint), ("precision", c_int)] +
                [(n, c_vp) for n in _PTRS] + [("ws_floats", c_ll), ("attn_ws", c_vp), ("tf_mask_dev", c_vp), ("persist_ws", c_vp),
                 ("team_ws", c_vp), ("team_ws_floats", c_ll)])


class AfDesc(ctypes.Structure):
    """Mirror of ``struct v2f_af_desc`` (include/v2f.h)."""
    _fields_ = [("p", c_vp), ("row", c_vp), ("col", c_vp), ("sq", c_vp), ("rms", c_vp), ("numel", c_ll),
                ("nmat", c_int), ("R", c_int), ("C", c_int), ("kind", c_int), ("inner", c_int), ("pad_", c_int),
                ("sO", c_ll), ("sI", c_ll), ("sR", c_ll), ("sC", c_ll)]


class AfPlan(ctypes.Structure):
    """Mirror of ``struct v2f_adafactor_plan`` (include/v2f.h)."""
    _fields_ = [("descs", c_vp), ("grads", c_vp), ("acc", c_vp), ("vec_units", c_vp), ("small_units", c_vp * 3),
                ("row_units", c_vp), ("col_units", c_vp), ("n_desc", c_int), ("n_vec", c_int), ("n_small", c_int * 3),
                ("n_rows", c_int), ("n_cols", c_int), ("eps1", c_float), ("eps2", c_float),
                ("clip_threshold", c_float), ("scale_parameter", c_int)]


def _declare(lib):
    lib.v2f_version.restype = c_int
    lib.v2f_launch_count.restype = c_ll
    lib.v2f_gemm_f32.argtypes = [c_int, c_int, c_int, c_int, c_int, c_vp, c_int, c_ll, c_vp, c_int, c_ll,
                                 c_vp, c_int, c_ll, c_int, c_vp, c_float, c_int, c_vp]
    lib.v2f_colsum_f32.argtypes = [c_int, c_int, c_vp, c_int, c_vp, c_float, c_vp]
    lib.v2f_mul_f32.argtypes = [c_ll, c_vp, c_vp, c_vp, c_vp]
    lib.v2f_decode_fwd.argtypes = [ctypes.POINTER(DecodeParams), c_vp]
    lib.v2f_decode_bwd.argtypes = [ctypes.POINTER(DecodeParams), c_vp]
    lib.v2f_gru_seq_fwd.argtypes = [c_int] * 4 + [c_vp] * 11 + [c_int, c_vp]
    lib.v2f_gru_seq_bwd.argtypes = [c_int] * 4 + [c_vp] * 19 + [c_vp, c_vp, c_ll, c_int, c_vp]
    sd = [c_int] * 5 + [c_vp, c_int, c_ll] * 3
    lib.v2f_sdpa_fwd.argtypes = sd + [c_vp, c_int, c_ll, c_vp, c_vp, c_vp, c_float, c_vp]
    lib.v2f_sdpa_bwd.argtypes = sd + [c_vp, c_int, c_ll, c_vp, c_vp] + [c_vp, c_int, c_ll] * 3 + [c_float, c_vp]
    lib.v2f_embed_fwd.argtypes = [c_int, c_int, c_vp, c_vp, c_vp, ctypes.POINTER(c_vp), c_vp, c_vp, c_vp, c_vp]
    lib.v2f_embed_bwd.argtypes = [c_int, c_int, c_vp, c_vp, c_vp, c_vp, ctypes.POINTER(c_int), c_vp, c_vp,
                                  ctypes.POINTER(c_vp), c_vp]
    lib.v2f_gemm_tc.argtypes = [c_int, c_int, c_int, c_int, c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, c_vp, c_float,
                                c_int, c_int, c_vp]
    lib.v2f_gemm_tc_batched.argtypes = [c_int, c_int, c_int, c_int, c_vp, c_ll, c_ll, c_vp, c_ll, c_ll, c_vp, c_ll,
                                        c_ll, c_int, c_vp, c_float, c_int, c_int, c_vp]
    lib.v2f_cast_bf16.argtypes = [c_ll, c_vp, c_vp, c_vp]
    lib.v2f_transpose.argtypes = [c_int, c_int, c_vp, c_ll, c_int, c_vp, c_ll, c_int, c_vp]
    lib.v2f_prof_enable.argtypes = [c_int]
    lib.v2f_gru_persistent_enable.argtypes = [c_int]
    lib.v2f_gru_persistent_enable.restype = c_int
    lib.v2f_image_normalize_u8.argtypes = [c_ll, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp]
    lib.v2f_image_normalize_u8.restype = c_int
    lib.v2f_adafactor_step.argtypes = [ctypes.POINTER(AfPlan), ctypes.c_double, ctypes.c_double, c_vp]
    lib.v2f_adafactor_step.restype = c_int
    lib.v2f_decode_persistent_enable.argtypes = [c_int]
    lib.v2f_decode_persistent_enable.restype = c_int
    lib.v2f_decode_persist_stamps_enable.argtypes = [c_int]
    lib.v2f_decode_persist_stamps_enable.restype = c_int
    lib.v2f_decode_persist_ws_floats.argtypes = [c_int, c_int, c_int, c_int]
    lib.v2f_decode_persist_ws_floats.restype = c_ll
    lib.v2f_decode_persist_stamps_offset.argtypes = [c_int, c_int, c_int]
    lib.v2f_decode_persist_stamps_offset.restype = c_ll
    c_double = ctypes.c_double
    lib.v2f_bn1d_stats.argtypes = [c_int, c_int, c_vp, c_vp, c_vp]
    lib.v2f_bn1d_apply.argtypes = [c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_double, c_vp, c_vp, c_float, c_float, c_vp,
                                   c_vp, c_vp, c_vp]
    lib.v2f_bn1d_bwd_stats.argtypes = [c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]
    lib.v2f_bn1d_bwd_apply.argtypes = [c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_double, c_vp, c_vp]
    lib.v2f_decode_team_ws_floats.argtypes = [c_int] * 5
    lib.v2f_decode_team_ws_floats.restype = c_ll
    lib.v2f_decode_team_stamps_offset.argtypes = [c_int] * 5
    lib.v2f_decode_team_stamps_offset.restype = c_ll
    lib.v2f_decode_team_enable.argtypes = [c_int]
    lib.v2f_decode_team_bwd_enable.argtypes = [c_int]
    lib.v2f_decode_team_bwd_ws_floats.argtypes = [c_int, c_int]
    lib.v2f_decode_team_bwd_ws_floats.restype = c_ll
    lib.v2f_decode_team_stamps_enable.argtypes = [c_int]
    lib.v2f_prof_read.argtypes = [c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_ll)]
    lib.v2f_prof_read_bytes.argtypes = [c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_ll),
                                        ctypes.POINTER(c_ll)]
    lib.v2f_prof_read_bytes.restype = c_int
    # ---- GTM-family row operators (csrc/gtm_ops.cu)
    lib.v2f_add_ln_fwd.argtypes = [c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_float, c_vp, c_vp, c_vp, c_vp]
    lib.v2f_add_ln_bwd_blocks.argtypes = [c_int]
    lib.v2f_add_ln_bwd.argtypes = [c_int, c_int] + [c_vp] * 10
    lib.v2f_bn1d_fwd.argtypes = [c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_float, c_float, c_vp, c_vp,
                                 c_vp, c_vp]
    lib.v2f_bn1d_bwd.argtypes = [c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]
    lib.v2f_dropout.argtypes = [c_ll, c_vp, c_vp, c_float, c_vp, c_vp]
    lib.v2f_gate_fwd.argtypes = [c_ll, c_vp, c_vp, c_int, c_vp, c_vp]
    lib.v2f_gate_bwd.argtypes = [c_ll, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp]
    lib.v2f_add_f32.argtypes = [c_ll, c_vp, c_vp, c_vp, c_vp]
    lib.v2f_relu_bwd.argtypes = [c_ll, c_vp, c_vp, c_vp, c_vp]
    lib.v2f_relu_fwd.argtypes = [c_ll, c_vp, c_vp, c_vp]
    lib.v2f_add_bcast.argtypes = [c_ll, c_ll, c_vp, c_vp, c_vp, c_vp]
    lib.v2f_copy2d.argtypes = [c_int, c_int, c_vp, c_ll, c_vp, c_ll, c_vp]
    lib.v2f_repeat_rows.argtypes = [c_int, c_int, c_ll, c_vp, c_vp, c_vp]
    lib.v2f_fold_rows.argtypes = [c_int, c_int, c_ll, c_vp, c_vp, c_vp]
    lib.v2f_gather4_fwd.argtypes = [c_int, c_int, ctypes.POINTER(c_vp), c_vp, c_vp, c_vp, c_vp]
    lib.v2f_gather4_bwd.argtypes = [c_int, c_int, c_vp, c_vp, c_vp, ctypes.POINTER(c_int), ctypes.POINTER(c_vp), c_vp]
    lib.v2f_feat4_fwd.argtypes = [c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]
    lib.v2f_feat4_bwd.argtypes = [c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]
    lib.v2f_meanpool_fwd.argtypes = [c_int, c_int, c_int, c_vp, c_int, c_int, c_vp, c_vp]
    lib.v2f_meanpool_bwd.argtypes = [c_int, c_int, c_int, c_vp, c_int, c_int, c_vp, c_vp]
    # ---- fused BatchNorm2d (+add) (+ReLU) for the bf16 channels_last trunk (csrc/bn_act.cu)
    lib.v2f_bn2d_blocks.argtypes = [c_ll, c_int]
    lib.v2f_bn2d_act_fwd.argtypes = [c_ll, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_float, c_float, c_int,
                                     c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]
    lib.v2f_bn2d_act_bwd.argtypes = [c_ll, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp,
                                     c_vp, c_vp, c_vp, c_vp, c_vp]
    lib.v2f_bn2d_relu_maxpool_fwd.argtypes = [c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_float,
                                              c_float, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]
    lib.v2f_bn2d_relu_maxpool_fwd_parts.argtypes = [c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_float,
                                                    c_float, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp]
    lib.v2f_bn2d_pdl_enable.argtypes = [c_int]
    lib.v2f_bn2d_pdl_enable.restype = c_int
    lib.v2f_stem_conv_blocks.argtypes = [c_int, c_int, c_int, c_int, c_int]
    lib.v2f_stem_conv_fwd.argtypes = [c_int, c_int, c_int, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp]
    for name in ("v2f_bn2d_blocks", "v2f_bn2d_act_fwd", "v2f_bn2d_act_bwd", "v2f_bn2d_relu_maxpool_fwd",
                 "v2f_bn2d_relu_maxpool_fwd_parts", "v2f_stem_conv_blocks", "v2f_stem_conv_fwd"):
        getattr(lib, name).restype = c_int
    for name in ("v2f_add_ln_fwd", "v2f_add_ln_bwd_blocks", "v2f_add_ln_bwd", "v2f_bn1d_fwd", "v2f_bn1d_bwd",
                 "v2f_gate_fwd", "v2f_gate_bwd", "v2f_add_f32", "v2f_relu_bwd", "v2f_relu_fwd", "v2f_add_bcast", "v2f_copy2d",
                 "v2f_repeat_rows", "v2f_fold_rows", "v2f_gather4_fwd", "v2f_gather4_bwd", "v2f_feat4_fwd",
                 "v2f_feat4_bwd", "v2f_meanpool_fwd", "v2f_meanpool_bwd"):
        getattr(lib, name).restype = c_int
    for name in ("v2f_gemm_tc", "v2f_gemm_tc_batched", "v2f_cast_bf16", "v2f_transpose", "v2f_prof_enable", "v2f_prof_read", "v2f_gemm_f32", "v2f_colsum_f32", "v2f_mul_f32", "v2f_decode_fwd", "v2f_decode_bwd",
                 "v2f_gru_seq_fwd", "v2f_gru_seq_bwd", "v2f_sdpa_fwd", "v2f_sdpa_bwd", "v2f_embed_fwd",
                 "v2f_embed_bwd"):
        getattr(lib, name).restype = c_int


def lib():
    """The loaded library.  Raises (loudly) if it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                "(nvcc, sm_100a).  This package has no CPU / eager fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        _declare(handle)
        if handle.v2f_version() != ABI_VERSION:
            raise RuntimeError("libv2f_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {_ERRORS.get(rc, rc)}")


def ptr(t, dtype=torch.float32, allow_none=False):
    """Device pointer of a contiguous CUDA tensor (or NULL)."""
    if t is None:
        if allow_none:
            return None
        raise ValueError("tensor required")
    if not t.is_cuda:
        raise RuntimeError("visuelle2-multimodal-fusion_b200 runs on CUDA only (got a CPU tensor)")
    if t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(lib().v2f_launch_count())


(K_ATTN_FWD, K_ATTN_BWD, K_TILEGRAD, K_BN_STATS, K_BN_APPLY, K_BN_BWD_REDUCE, K_BN_BWD_ELEMT, K_DECODE_PERSIST_FWD,
 K_DECODE_PERSIST_BWD, K_STEM_CONV, K_GEMM_TC) = range(11)


def prof_enable(on):
    check(lib().v2f_prof_enable(1 if on else 0), "v2f_prof_enable")


def prof_read(kernel_id):
    """(total_ms, launches) of one kernel id since the previous read."""
    ms, n = ctypes.c_double(0.0), c_ll(0)
    check(lib().v2f_prof_read(kernel_id, ctypes.byref(ms), ctypes.byref(n)), "v2f_prof_read")
    return ms.value, n.value


def prof_read_bytes(kernel_id):
    """(total_ms, launches, algorithmic_bytes) of one kernel id since the previous read."""
    ms, n, b = ctypes.c_double(0.0), c_ll(0), c_ll(0)
    check(lib().v2f_prof_read_bytes(kernel_id, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(b)),
          "v2f_prof_read_bytes")
    return ms.value, n.value, b.value
