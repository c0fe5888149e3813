"""ctypes binding of libkmunet.so (the C ABI declared in include/kmunet.h).

There is no CPU path and no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkmunet.so")

KMU_PREC_FP32 = 0
KMU_PREC_BF16 = 1

_f32p = C.c_void_p  # device pointers travel as integers


class KanDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "Cin", "H", "W", "Cout", "ksize", "stride", "padding", "grid_size",
                                         "spline_order", "precision", "has_scaler", "grid_uniform")] + \
               [("grid_t0", C.c_float), ("grid_h", C.c_float)]


class KanFwdArgs(C.Structure):
    _fields_ = [("d", KanDesc), ("x", _f32p), ("base_weight", _f32p), ("spline_weight", _f32p), ("spline_scaler", _f32p),
                ("grid", _f32p), ("y", _f32p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class KanBwdArgs(C.Structure):
    _fields_ = [("d", KanDesc), ("x", _f32p), ("dy", _f32p), ("base_weight", _f32p), ("spline_weight", _f32p),
                ("spline_scaler", _f32p), ("grid", _f32p), ("dx", _f32p), ("d_base_weight", _f32p),
                ("d_spline_weight", _f32p), ("d_spline_scaler", _f32p), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_size_t)]


class HsmDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "C", "L", "H", "N", "precision")]


class HsmFwdArgs(C.Structure):
    _fields_ = [("d", HsmDesc)] + [(n, _f32p) for n in ("x", "w_bcdt", "w_dw", "w_hz", "w_out", "A", "D", "y", "h", "P",
                                                        "stats", "hs", "hz")] + \
               [("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class HsmBwdArgs(C.Structure):
    _fields_ = [("d", HsmDesc)] + [(n, _f32p) for n in ("x", "dy", "dh", "w_bcdt", "w_dw", "w_hz", "w_out", "A", "D", "P", "stats",
                                                        "hs", "hz", "h", "dx", "d_w_bcdt", "d_w_dw", "d_w_hz", "d_w_out",
                                                        "d_A", "d_D")] + \
               [("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class DysDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "C", "H", "W", "scale", "groups")]


class DysFwdArgs(C.Structure):
    _fields_ = [("d", DysDesc)] + [(n, _f32p) for n in ("x", "w_offset", "b_offset", "init_pos", "offset", "out")]


class DysBwdArgs(C.Structure):
    _fields_ = [("d", DysDesc)] + [(n, _f32p) for n in ("x", "w_offset", "offset", "dout", "dx", "d_w_offset", "d_b_offset")] + \
               [("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class DeformDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "C", "H", "W", "Cout")]


class AdamWArgs(C.Structure):          # kmu_adamw_args
    _fields_ = [("entries", C.c_void_p), ("chunks", C.c_void_p), ("n_entries", C.c_int32), ("n_chunks", C.c_int32),
                ("step", _f32p), ("lr_dev", _f32p)] + \
               [(n, C.c_double) for n in ("lr", "beta1", "beta2", "eps", "weight_decay", "grad_scale")]


class DagemDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "C", "H", "W", "training")] + [("momentum", C.c_float), ("eps", C.c_float)]


class DagemBn(C.Structure):
    _fields_ = [(n, _f32p) for n in ("weight", "bias", "running_mean", "running_var")]


class DagemFwdArgs(C.Structure):
    _fields_ = [("d", DagemDesc)] + [(n, _f32p) for n in ("x", "deformed", "ea_w", "ea_b", "vu_w", "vu_b", "eu_w", "eu_b",
                                                          "er_w", "er_b", "wf")] + \
               [("bn", DagemBn * 5), ("out", _f32p), ("saved", _f32p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class DagemBwdArgs(C.Structure):
    _fields_ = [("d", DagemDesc)] + [(n, _f32p) for n in ("x", "deformed", "dout", "saved", "ea_w", "vu_w", "eu_w", "er_w", "wf",
                                                          "dx", "d_deformed", "d_ea_w", "d_ea_b", "d_vu_w", "d_vu_b", "d_eu_w",
                                                          "d_eu_b", "d_er_w", "d_er_b", "d_wf")] + \
               [("d_bn_weight", _f32p * 5), ("d_bn_bias", _f32p * 5), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class BnMixDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "C", "HW", "training", "relu", "mix")] + [("momentum", C.c_float), ("eps", C.c_float)]


class BnMixFwdArgs(C.Structure):
    _fields_ = [("d", BnMixDesc)] + [(n, _f32p) for n in ("x", "weight", "bias", "running_mean", "running_var", "res", "alpha", "y",
                                                          "stat")] + [("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class BnMixBwdArgs(C.Structure):
    _fields_ = [("d", BnMixDesc)] + [(n, _f32p) for n in ("x", "dy", "weight", "bias", "stat", "res", "alpha", "dx", "d_weight",
                                                          "d_bias", "d_res", "d_alpha")] + \
               [("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class DwDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "C", "H", "W")]


class PwDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "Cin", "Cout", "HW")]


class TnDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "C", "HW")] + [("eps_gn", C.c_float), ("eps_ln", C.c_float)]


class TnFwdArgs(C.Structure):
    _fields_ = [("d", TnDesc)] + [(n, _f32p) for n in ("x", "gh", "bh", "gw", "bw", "gc", "bc", "y", "gstat")] + \
               [("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class TnBwdArgs(C.Structure):
    _fields_ = [("d", TnDesc)] + [(n, _f32p) for n in ("x", "dy", "gstat", "gh", "gw", "gc", "dx", "d_gh", "d_bh", "d_gw", "d_bw",
                                                       "d_gc", "d_bc")] + [("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class ScDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "Cin", "Cout", "H", "W", "kh", "kw")]


# every symbol include/kmunet.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "kmu_version": (C.c_int, []),
    "kmu_last_error": (C.c_char_p, []),
    "kmu_device_supported": (C.c_int, []),
    "kmu_launch_count": (C.c_uint64, []),
    "kmu_kanconv2d_fwd_workspace_bytes": (C.c_size_t, [C.POINTER(KanDesc)]),
    "kmu_kanconv2d_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(KanDesc)]),
    "kmu_kanconv2d_fwd": (C.c_int, [C.POINTER(KanFwdArgs), C.c_void_p]),
    "kmu_kanconv2d_bwd": (C.c_int, [C.POINTER(KanBwdArgs), C.c_void_p]),
    "kmu_kanconv2d_path": (C.c_int, [C.POINTER(KanDesc)]),
    "kmu_debug_flags": (None, [C.c_int]),
    "kmu_hsmssd_fwd_workspace_bytes": (C.c_size_t, [C.POINTER(HsmDesc)]),
    "kmu_hsmssd_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(HsmDesc)]),
    "kmu_hsmssd_fwd": (C.c_int, [C.POINTER(HsmFwdArgs), C.c_void_p]),
    "kmu_hsmssd_bwd": (C.c_int, [C.POINTER(HsmBwdArgs), C.c_void_p]),
    "kmu_layernorm1d_fwd": (C.c_int, [_f32p, _f32p, _f32p, _f32p, _f32p, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                      C.c_void_p]),
    "kmu_layernorm1d_bwd_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "kmu_layernorm1d_bwd": (C.c_int, [_f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                      C.c_void_p, C.c_size_t, C.c_void_p]),
    "kmu_dysample_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(DysDesc)]),
    "kmu_dysample_fwd": (C.c_int, [C.POINTER(DysFwdArgs), C.c_void_p]),
    "kmu_dysample_bwd": (C.c_int, [C.POINTER(DysBwdArgs), C.c_void_p]),
    "kmu_dysample_sample_fwd": (C.c_int, [C.POINTER(DysDesc), _f32p, _f32p, _f32p, C.c_void_p]),
    "kmu_dysample_sample_bwd": (C.c_int, [C.POINTER(DysDesc), _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p]),
    "kmu_set_deterministic": (None, [C.c_int]),
    "kmu_get_deterministic": (C.c_int, []),
    "kmu_deformconv3x3_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(DeformDesc)]),
    "kmu_deformconv3x3_fwd": (C.c_int, [C.POINTER(DeformDesc), _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p]),
    "kmu_deformconv3x3_bwd": (C.c_int, [C.POINTER(DeformDesc), _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p,
                                        C.c_size_t, C.c_void_p]),
    "kmu_adamw_chunk_elems": (C.c_int32, []),
    "kmu_adamw_step": (C.c_int, [C.POINTER(AdamWArgs), C.c_void_p]),
    "kmu_bnmix_workspace_bytes": (C.c_size_t, [C.POINTER(BnMixDesc)]),
    "kmu_bnmix_fwd": (C.c_int, [C.POINTER(BnMixFwdArgs), C.c_void_p]),
    "kmu_bnmix_bwd": (C.c_int, [C.POINTER(BnMixBwdArgs), C.c_void_p]),
    "kmu_dwconv3x3_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(DwDesc)]),
    "kmu_dwconv3x3_fwd": (C.c_int, [C.POINTER(DwDesc), _f32p, _f32p, _f32p, _f32p, C.c_void_p]),
    "kmu_dwconv3x3_bwd": (C.c_int, [C.POINTER(DwDesc), _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kmu_dwconv3x3_bwd_add": (C.c_int, [C.POINTER(DwDesc), _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p, C.c_size_t,
                                        C.c_void_p]),
    "kmu_dwconv3x3_scaled_fwd": (C.c_int, [C.POINTER(DwDesc), _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p]),
    "kmu_dwconv3x3_scaled_bwd": (C.c_int, [C.POINTER(DwDesc), _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p,
                                           C.c_size_t, C.c_void_p]),
    "kmu_pwconv_wgrad_supported": (C.c_int, [C.POINTER(PwDesc)]),
    "kmu_pwconv_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(PwDesc)]),
    "kmu_pwconv_fwd": (C.c_int, [C.POINTER(PwDesc), _f32p, _f32p, _f32p, _f32p, C.c_void_p]),
    "kmu_pwconv_bwd": (C.c_int, [C.POINTER(PwDesc), _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kmu_triplenorm_workspace_bytes": (C.c_size_t, [C.POINTER(TnDesc)]),
    "kmu_triplenorm_fwd": (C.c_int, [C.POINTER(TnFwdArgs), C.c_void_p]),
    "kmu_triplenorm_bwd": (C.c_int, [C.POINTER(TnBwdArgs), C.c_void_p]),
    "kmu_qkv_gate_fwd": (C.c_int, [_f32p, _f32p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "kmu_qkv_gate_bwd": (C.c_int, [_f32p, _f32p, _f32p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "kmu_lerpmix_bwd_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int64]),
    "kmu_lerpmix_fwd": (C.c_int, [_f32p, _f32p, _f32p, _f32p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p]),
    "kmu_lerpmix_bwd": (C.c_int, [_f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p,
                                  C.c_size_t, C.c_void_p]),
    "kmu_hybridloss_workspace_bytes": (C.c_size_t, []),
    "kmu_hybridloss_stats": (C.c_int, [_f32p, _f32p, C.c_int64, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kmu_hybridloss_stack": (C.c_int, [_f32p, _f32p, _f32p, _f32p, C.c_int64, C.c_void_p]),
    "kmu_hybridloss_ssim": (C.c_int, [_f32p, _f32p, _f32p, _f32p, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                      C.c_size_t, C.c_void_p]),
    "kmu_hybridloss_bwd": (C.c_int, [_f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int64, C.c_float, C.c_void_p]),
    "kmu_groupnorm_fwd_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int64, C.c_int32]),
    "kmu_groupnorm_fwd": (C.c_int, [_f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_float,
                                    C.c_void_p, C.c_size_t, C.c_void_p]),
    "kmu_resize_bilinear_ac_fwd": (C.c_int, [_f32p, _f32p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "kmu_combine3_bwd_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int64]),
    "kmu_combine3_fwd": (C.c_int, [_f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int32, C.c_int64, C.c_void_p]),
    "kmu_combine3_bwd": (C.c_int, [_f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int32, C.c_int64, C.c_void_p,
                                   C.c_size_t, C.c_void_p]),
    "kmu_smallconv_supported": (C.c_int, [C.POINTER(ScDesc)]),
    "kmu_smallconv_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(ScDesc)]),
    "kmu_smallconv_fwd": (C.c_int, [C.POINTER(ScDesc), _f32p, _f32p, _f32p, _f32p, C.c_void_p]),
    "kmu_smallconv_bwd": (C.c_int, [C.POINTER(ScDesc), _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kmu_iwp_bwd_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "kmu_iwp_fwd": (C.c_int, [_f32p, _f32p, _f32p, _f32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "kmu_iwp_bwd": (C.c_int, [_f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                              C.c_size_t, C.c_void_p]),
    "kmu_pwconv_tc_supported": (C.c_int, [C.POINTER(PwDesc)]),
    "kmu_pwconv_tc_wgrad_supported": (C.c_int, [C.POINTER(PwDesc)]),
    "kmu_pwconv_tc_workspace_bytes": (C.c_size_t, [C.POINTER(PwDesc)]),
    "kmu_pwconv_tc_fwd": (C.c_int, [C.POINTER(PwDesc), _f32p, _f32p, _f32p, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kmu_pwconv_tc_bwd": (C.c_int, [C.POINTER(PwDesc), _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kmu_pwconv_tma_fwd_supported": (C.c_int, [C.POINTER(PwDesc)]),
    "kmu_pwconv_tma_fwd_workspace_bytes": (C.c_size_t, [C.POINTER(PwDesc)]),
    "kmu_pwconv_tma_fwd": (C.c_int, [C.POINTER(PwDesc), _f32p, _f32p, _f32p, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kmu_pwconv_fused_bwd_supported": (C.c_int, [C.POINTER(PwDesc)]),
    "kmu_pwconv_fused_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(PwDesc)]),
    "kmu_pwconv_fused_bwd": (C.c_int, [C.POINTER(PwDesc), _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kmu_dagem_saved_bytes": (C.c_size_t, [C.POINTER(DagemDesc)]),
    "kmu_dagem_fwd_workspace_bytes": (C.c_size_t, [C.POINTER(DagemDesc)]),
    "kmu_dagem_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(DagemDesc)]),
    "kmu_dagem_fwd": (C.c_int, [C.POINTER(DagemFwdArgs), C.c_void_p]),
    "kmu_dagem_bwd": (C.c_int, [C.POINTER(DagemBwdArgs), C.c_void_p]),
}

_lib = None


def lib():
    """Load libkmunet.so once; raise loudly if it has not been built (python -m km_unet_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m km_unet_b200.build` (or __graft_entry__.build()). "
                "km_unet_b200 has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError here = header and library out of sync
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error():
    return lib().kmu_last_error().decode("utf-8", "replace")


def check(status, what):
    if status != 0:
        raise RuntimeError(f"{what} failed (kmu_status {status}): {last_error()}")


def ptr(t):
    """Device pointer of a contiguous fp32 CUDA tensor (or NULL for None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("km_unet_b200 kernels take CUDA tensors only (there is no CPU fallback)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"km_unet_b200 kernels take float32 tensors, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError("km_unet_b200 kernels take contiguous tensors")
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(lib().kmu_launch_count())
