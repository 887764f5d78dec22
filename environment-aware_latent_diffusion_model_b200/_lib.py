"""ctypes binding of libealdm_b200.so (the C ABI declared in include/ealdm_b200.h).

The library is the product: there is no Python/PyTorch fallback for any operator.  If the shared
object is missing or the device is not sm_100, loading fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libealdm_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

F32, BF16 = 0, 1
ACT_NONE, ACT_SILU, ACT_GEGLU, ACT_RELU, ACT_SOFTMAX4 = 0, 1, 2, 3, 4
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05 = 0, 1, 2
TC_OPT_CTA2, TC_OPT_WIDE, TC_OPT_RELAXED_WAIT, TC_OPT_BN, TC_OPT_GELU_ERF, TC_OPT_STREAMK = 0, 1, 2, 3, 4, 5

EXPORTS = [
    "ealdm_abi_version", "ealdm_last_error", "ealdm_device_check", "ealdm_launch_count", "ealdm_tc_set_option", "ealdm_set_pdl",
    "ealdm_conv", "ealdm_conv_ln_parts", "ealdm_ff_geglu_fused", "ealdm_group_norm", "ealdm_group_norm_workspace_bytes", "ealdm_layer_norm", "ealdm_attention",
    "ealdm_timestep_embedding", "ealdm_nchw_to_nhwc", "ealdm_nhwc_to_nchw",
    "ealdm_upsample_nearest2x", "ealdm_copy2d", "ealdm_softmax_rows", "ealdm_ddim_step",
    "ealdm_q_sample", "ealdm_cfg_mse",
    # backward pass
    "ealdm_conv_wgrad", "ealdm_conv_wgrad_workspace_bytes", "ealdm_group_norm_bwd",
    "ealdm_group_norm_bwd_workspace_bytes", "ealdm_layer_norm_bwd", "ealdm_layer_norm_bwd_workspace_bytes",
    "ealdm_attention_bwd", "ealdm_attention_bwd_workspace_bytes", "ealdm_geglu", "ealdm_geglu_bwd",
    "ealdm_silu", "ealdm_silu_bwd", "ealdm_colsum", "ealdm_colsum_workspace_bytes", "ealdm_zero_insert2x",
    "ealdm_sumpool2x2", "ealdm_cfg_mse_bwd", "ealdm_gn_partial", "ealdm_adamw_ema_step", "ealdm_plms_eps", "ealdm_vq_nearest", "ealdm_ddpm_step",
    # conditioner
    "ealdm_fourier_style", "ealdm_lstm_cell", "ealdm_adain", "ealdm_batch_norm_relu",
    "ealdm_adain_bwd", "ealdm_batch_norm_relu_bwd", "ealdm_lstm_cell_bwd", "ealdm_relu_bwd",
]
WGRAD_PACKED, WGRAD_OIHW = 0, 1


class ConvSrc(C.Structure):
    _fields_ = [("x", C.c_void_p), ("n", C.c_int64), ("h", C.c_int64), ("w", C.c_int64),
                ("c", C.c_int64), ("ld", C.c_int64), ("ksize", C.c_int32), ("stride", C.c_int32),
                ("pad", C.c_int32), ("upsample", C.c_int32)]


class ConvArgs(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("impl", C.c_int32), ("n_src", C.c_int32), ("act", C.c_int32),
                ("src", ConvSrc * 2), ("weight", C.c_void_p), ("n_out", C.c_int64),
                ("k_total", C.c_int64), ("h_out", C.c_int64), ("w_out", C.c_int64),
                ("bias", C.c_void_p), ("rowvec", C.c_void_p), ("ld_rowvec", C.c_int64),
                ("residual", C.c_void_p), ("ld_res", C.c_int64), ("out", C.c_void_p),
                ("ld_out", C.c_int64), ("out_f32", C.c_int32), ("res_f32", C.c_int32),
                ("out2", C.c_void_p), ("ld_out2", C.c_int64), ("gn_partial", C.c_void_p), ("gn_ld", C.c_int64),
                ("weight_adjoint", C.c_int32), ("upsample_phases", C.c_int32), ("ld_weight", C.c_int64),
                ("ln_partial_out", C.c_void_p), ("ln_partial_in", C.c_void_p), ("ln_parts_in", C.c_int64),
                ("ln_c1", C.c_void_p), ("ln_channels", C.c_int64), ("ln_eps", C.c_float), ("wi_tokens", C.c_int32),
                ("wi_heads", C.c_int32), ("gn_unit", C.c_int32), ("wi_ld", C.c_int64), ("wi_head_stride", C.c_int64),
                ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("gn_gamma", C.c_void_p), ("gn_beta", C.c_void_p),
                ("gn_eps", C.c_float), ("gn_groups", C.c_int32), ("gn_silu", C.c_int32), ("gn_only", C.c_int32)]


class GroupNormArgs(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("act", C.c_int32), ("x", C.c_void_p), ("n", C.c_int64),
                ("hw", C.c_int64), ("c", C.c_int64), ("ld_x", C.c_int64), ("groups", C.c_int32),
                ("eps", C.c_float), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("y", C.c_void_p),
                ("ld_y", C.c_int64), ("workspace", C.c_void_p), ("x_f32", C.c_int32), ("partial_unit", C.c_int32),
                ("stats_out", C.c_void_p), ("partial", C.c_void_p), ("partial_ld", C.c_int64)]


class FfFusedArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("rows", C.c_int64), ("ld_x", C.c_int64), ("c", C.c_int64), ("hidden", C.c_int64),
                ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
                ("residual", C.c_void_p), ("ld_res", C.c_int64), ("out", C.c_void_p), ("ld_out", C.c_int64),
                ("out_f32", C.c_int32), ("reserved", C.c_int32)]


class LayerNormArgs(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("x_f32", C.c_int32), ("x", C.c_void_p),
                ("rows", C.c_int64), ("c", C.c_int64), ("ld_x", C.c_int64), ("eps", C.c_float),
                ("reserved2", C.c_int32), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("y", C.c_void_p), ("ld_y", C.c_int64)]


class AttentionArgs(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("impl", C.c_int32), ("q", C.c_void_p), ("k", C.c_void_p),
                ("v", C.c_void_p), ("ld_q", C.c_int64), ("ld_kv", C.c_int64),
                ("head_stride_q", C.c_int64), ("head_stride_kv", C.c_int64), ("batch", C.c_int64),
                ("heads", C.c_int64), ("n_q", C.c_int64), ("n_kv", C.c_int64),
                ("head_dim", C.c_int64), ("scale", C.c_float), ("reserved", C.c_int32),
                ("out", C.c_void_p), ("ld_out", C.c_int64), ("lse", C.c_void_p)]


class DdimStepArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("e_uncond", C.c_void_p), ("e_cond", C.c_void_p),
                ("noise", C.c_void_p), ("x_prev", C.c_void_p), ("pred_x0", C.c_void_p),
                ("e_out", C.c_void_p), ("numel", C.c_int64), ("cfg_scale", C.c_float),
                ("sqrt_one_minus_at", C.c_float), ("sqrt_at", C.c_float),
                ("sqrt_a_prev", C.c_float), ("dir_coef", C.c_float), ("sigma_t", C.c_float),
                ("temperature", C.c_float), ("reserved", C.c_int32)]


class ConvWgradArgs(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("impl", C.c_int32), ("src", ConvSrc), ("dy", C.c_void_p),
                ("ld_dy", C.c_int64), ("n_out", C.c_int64), ("h_out", C.c_int64), ("w_out", C.c_int64),
                ("dw", C.c_void_p), ("ld_dw", C.c_int64), ("layout", C.c_int32), ("accumulate", C.c_int32),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64)]


class GroupNormBwdArgs(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("act", C.c_int32), ("x", C.c_void_p), ("n", C.c_int64),
                ("hw", C.c_int64), ("c", C.c_int64), ("ld_x", C.c_int64), ("groups", C.c_int32),
                ("x_f32", C.c_int32), ("stats", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("dy", C.c_void_p), ("ld_dy", C.c_int64), ("add", C.c_void_p), ("ld_add", C.c_int64),
                ("add2", C.c_void_p), ("ld_add2", C.c_int64), ("dx", C.c_void_p), ("ld_dx", C.c_int64),
                ("dx_f32", C.c_int32), ("reserved", C.c_int32), ("dx2", C.c_void_p), ("ld_dx2", C.c_int64),
                ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("workspace", C.c_void_p)]


class LayerNormBwdArgs(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("x_f32", C.c_int32), ("x", C.c_void_p), ("rows", C.c_int64),
                ("c", C.c_int64), ("ld_x", C.c_int64), ("eps", C.c_float), ("dx_f32", C.c_int32),
                ("gamma", C.c_void_p), ("dy", C.c_void_p), ("ld_dy", C.c_int64), ("add", C.c_void_p),
                ("ld_add", C.c_int64), ("dx", C.c_void_p), ("ld_dx", C.c_int64), ("dx2", C.c_void_p),
                ("ld_dx2", C.c_int64), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("workspace", C.c_void_p)]


class AttentionBwdArgs(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("impl", C.c_int32), ("q", C.c_void_p), ("k", C.c_void_p),
                ("v", C.c_void_p), ("out", C.c_void_p), ("dout", C.c_void_p), ("ld_q", C.c_int64),
                ("ld_kv", C.c_int64), ("ld_out", C.c_int64), ("ld_dout", C.c_int64),
                ("head_stride_q", C.c_int64), ("head_stride_kv", C.c_int64), ("batch", C.c_int64),
                ("heads", C.c_int64), ("n_q", C.c_int64), ("n_kv", C.c_int64), ("head_dim", C.c_int64),
                ("scale", C.c_float), ("reserved", C.c_int32), ("dq", C.c_void_p), ("ld_dq", C.c_int64),
                ("head_stride_dq", C.c_int64), ("dk", C.c_void_p), ("dv", C.c_void_p), ("ld_dkv", C.c_int64),
                ("head_stride_dkv", C.c_int64), ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
                ("lse", C.c_void_p)]


class AdamWArgs(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("ema", C.c_void_p), ("param_bf16", C.c_void_p), ("numel", C.c_int64), ("step", C.c_int64),
                ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("weight_decay", C.c_float), ("grad_scale", C.c_float), ("ema_decay", C.c_float),
                ("reserved", C.c_int32)]


class EaldmError(RuntimeError):
    pass


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libealdm_b200.so (in-tree) with nvcc via make."""
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    r = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise EaldmError("building libealdm_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


_lib = None


def _declare(lib):
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.ealdm_abi_version.restype = C.c_int
    lib.ealdm_abi_version.argtypes = []
    lib.ealdm_last_error.restype = C.c_char_p
    lib.ealdm_last_error.argtypes = []
    lib.ealdm_device_check.restype = C.c_int
    lib.ealdm_device_check.argtypes = []
    lib.ealdm_launch_count.restype = C.c_int64
    lib.ealdm_launch_count.argtypes = []
    lib.ealdm_tc_set_option.restype = C.c_int
    lib.ealdm_tc_set_option.argtypes = [C.c_int, C.c_int]
    lib.ealdm_set_pdl.restype = C.c_int
    lib.ealdm_set_pdl.argtypes = [C.c_int]
    lib.ealdm_conv_ln_parts.restype = C.c_int64
    lib.ealdm_conv_ln_parts.argtypes = [C.POINTER(ConvArgs)]
    lib.ealdm_group_norm_workspace_bytes.restype = C.c_int64
    lib.ealdm_group_norm_workspace_bytes.argtypes = [i64, i64, i64]
    for name, argt in [
        ("ealdm_conv_wgrad_workspace_bytes", [C.POINTER(ConvWgradArgs)]),
        ("ealdm_group_norm_bwd_workspace_bytes", [i64, i64, i64]),
        ("ealdm_layer_norm_bwd_workspace_bytes", [i64, i64]),
        ("ealdm_attention_bwd_workspace_bytes", [C.POINTER(AttentionBwdArgs)]),
        ("ealdm_colsum_workspace_bytes", [i64, i64, i64]),
    ]:
        fn = getattr(lib, name)
        fn.restype = C.c_int64
        fn.argtypes = argt
    for name, argt in [
        ("ealdm_conv", [C.POINTER(ConvArgs), vp]),
        ("ealdm_group_norm", [C.POINTER(GroupNormArgs), vp]),
        ("ealdm_layer_norm", [C.POINTER(LayerNormArgs), vp]),
        ("ealdm_ff_geglu_fused", [C.POINTER(FfFusedArgs), vp]),
        ("ealdm_attention", [C.POINTER(AttentionArgs), vp]),
        ("ealdm_timestep_embedding", [vp, i64, i32, vp, i32, vp, vp]),
        ("ealdm_nchw_to_nhwc", [vp, i64, i64, i64, i64, i32, vp, i64, vp]),
        ("ealdm_nhwc_to_nchw", [vp, i64, i32, i64, i64, i64, i64, vp, vp]),
        ("ealdm_upsample_nearest2x", [vp, i64, i32, i64, i64, i64, i64, vp, i64, vp]),
        ("ealdm_copy2d", [vp, i64, i32, vp, i64, i32, i64, i64, vp]),
        ("ealdm_softmax_rows", [vp, i64, i32, i64, i64, f32, vp]),
        ("ealdm_ddim_step", [C.POINTER(DdimStepArgs), vp]),
        ("ealdm_q_sample", [vp, vp, vp, vp, vp, i64, i64, vp, vp]),
        ("ealdm_cfg_mse", [vp, vp, vp, f32, i64, i64, vp, vp]),
        ("ealdm_conv_wgrad", [C.POINTER(ConvWgradArgs), vp]),
        ("ealdm_group_norm_bwd", [C.POINTER(GroupNormBwdArgs), vp]),
        ("ealdm_layer_norm_bwd", [C.POINTER(LayerNormBwdArgs), vp]),
        ("ealdm_attention_bwd", [C.POINTER(AttentionBwdArgs), vp]),
        ("ealdm_geglu", [vp, i64, i32, i64, i64, vp, i64, vp]),
        ("ealdm_geglu_bwd", [vp, i64, vp, i64, i32, i64, i64, vp, i64, vp]),
        ("ealdm_silu", [vp, i64, i64, i64, i32, vp, i64, vp]),
        ("ealdm_silu_bwd", [vp, i64, vp, i64, i32, i64, i64, vp, i64, vp]),
        ("ealdm_colsum", [vp, i64, i32, i64, i64, i64, vp, i64, i32, vp, vp]),
        ("ealdm_zero_insert2x", [vp, i64, i32, i64, i64, i64, i64, vp, i64, vp]),
        ("ealdm_sumpool2x2", [vp, i64, i32, i64, i64, i64, i64, vp, i64, vp, i64, vp, i64, vp]),
        ("ealdm_cfg_mse_bwd", [vp, vp, vp, vp, f32, i64, i64, vp, vp, vp]),
        ("ealdm_gn_partial", [vp, i64, i32, i64, i64, i64, vp, i64, vp]),
        ("ealdm_adamw_ema_step", [C.POINTER(AdamWArgs), vp]),
        ("ealdm_plms_eps", [vp, vp, f32, vp, vp, vp, i32, vp, vp, i64, vp]),
        ("ealdm_vq_nearest", [vp, i64, i64, i64, vp, i64, vp, vp, vp]),
        ("ealdm_ddpm_step", [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, f32, i64, i64, vp, vp, vp]),
        ("ealdm_fourier_style", [vp, i64, vp, i32, i32, f32, vp, i64, f32, vp, vp, vp]),
        ("ealdm_lstm_cell", [vp, i64, i64, i64, vp, vp, vp, vp, vp, i64, vp, i64, vp, vp]),
        ("ealdm_adain", [vp, i64, i64, i64, i64, vp, i64, f32, vp, i64, vp]),
        ("ealdm_batch_norm_relu", [vp, i64, i64, i64, vp, vp, vp, vp, i32, f32, i32, vp, i64, vp, vp]),
        ("ealdm_adain_bwd", [vp, i64, i64, i64, i64, vp, i64, f32, vp, i64, vp]),
        ("ealdm_batch_norm_relu_bwd", [vp, i64, i64, i64, vp, vp, vp, i32, f32, i32, vp, i64, vp, i64, vp, i64, vp, vp, vp]),
        ("ealdm_lstm_cell_bwd", [vp, i64, i64, i64, vp, vp, vp, vp, vp, i64, vp, i64, vp, vp, vp, vp]),
        ("ealdm_relu_bwd", [vp, i64, vp, i64, vp, i64, i64, i64, vp, i64, vp]),
    ]:
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = argt


def load():
    """Load the shared library (never builds implicitly on a GPU box; see __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EaldmError(
                f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the CUDA kernels)")
        lib = C.CDLL(LIB_PATH)
        _declare(lib)
        if lib.ealdm_abi_version() != 2:
            raise EaldmError("libealdm_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        msg = load().ealdm_last_error().decode("utf-8", "replace")
        raise EaldmError(f"libealdm_b200 error {rc}: {msg}")
