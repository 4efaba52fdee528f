"""ctypes binding of libanoddpm_b200.so (the C ABI declared in include/anoddpm_b200.h).

The library is built in-tree by ``python -m anoddpm_b200.build`` (nvcc, sm_100a).  There is NO
fallback: if the shared object is missing or a call fails, this module raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ANODDPM_B200_LIB") or os.path.join(_HERE, "libanoddpm_b200.so")   # override: A/B runs of two builds

c_int, c_i64, c_f32, c_f64, c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double, ctypes.c_void_p


class ConvDesc(ctypes.Structure):
    """Mirror of ``anoddpm_conv_desc`` (include/anoddpm_b200.h)."""
    _fields_ = [("out", c_vp), ("a", c_vp), ("r1", c_vp), ("r2", c_vp), ("w", c_vp), ("bias", c_vp), ("emb", c_vp),
                ("res", c_vp), ("B", c_int), ("H", c_int), ("W", c_int), ("Ca", c_int), ("Cr1", c_int), ("Cr2", c_int),
                ("Cout", c_int), ("ksize", c_int), ("emb_stride", c_int), ("res_mode", c_int), ("impl", c_int), ("operand_f16", c_int), ("gn_stats", c_vp),
                ("out_f16", c_int), ("a2", c_vp), ("Ca2", c_int), ("gn_table", c_vp), ("out16", c_vp), ("workspace", c_vp),
                ("workspace_bytes", ctypes.c_longlong)]


# name -> (restype, argtypes); every symbol of include/anoddpm_b200.h
SIGNATURES = {
    "anoddpm_abi_version": (c_int, []),
    "anoddpm_last_error": (ctypes.c_char_p, []),
    "anoddpm_build_info": (ctypes.c_char_p, []),
    "anoddpm_fp64_fma_probe": (c_int, [c_vp, c_int, c_int, c_vp]),
    "anoddpm_simplex_init": (c_int, [c_i64, c_vp]),
    "anoddpm_simplex_init_batch": (c_int, [c_vp, c_int, c_vp]),
    "anoddpm_simplex_noise": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp, c_int, c_f64, c_f64, c_vp]),
    "anoddpm_simplex_noise_steps": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp, c_int, c_f64, c_f64, c_vp]),
    "anoddpm_simplex_octaves": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_int, c_f64, c_f64, c_vp]),
    "anoddpm_simplex_grid3": (c_int, [c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_int, c_vp, c_vp]),
    "anoddpm_simplex_grid2": (c_int, [c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_vp]),
    "anoddpm_q_sample": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_vp]),
    "anoddpm_p_sample": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_vp]),
    "anoddpm_p_sample_simplex": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_f64,
                                         c_f64, c_vp]),
    "anoddpm_detection_counts": (c_int, [c_vp, c_vp, c_vp, c_vp, c_f32, c_int, c_i64, c_vp]),
    "anoddpm_vlb_terms": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_vp]),
    "anoddpm_conv_wgrad_supported": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "anoddpm_conv_wgrad": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_f32, c_vp]),
    "anoddpm_gn_silu_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                                    c_vp, c_int, c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_f32, c_int, c_int,
                                    c_f32, c_vp]),
    "anoddpm_cast_colsum": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_f32, c_f32, c_vp]),
    "anoddpm_colsum_f16": (c_int, [c_vp, c_vp, c_i64, c_int, c_f32, c_vp]),
    "anoddpm_gemm_batched": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_int,
                                     c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_f32, c_int, c_vp]),
    "anoddpm_softmax_rows": (c_int, [c_vp, c_i64, c_int, c_vp]),
    "anoddpm_softmax_bwd_rows": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp]),
    "anoddpm_silu_bwd": (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    "anoddpm_stem_wgrad": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_f32, c_vp]),
    "anoddpm_head_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_f32, c_vp]),
    "anoddpm_pack_conv_w": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_i64, c_int, c_int, c_vp]),
    "anoddpm_sumsq": (c_int, [c_vp, c_vp, c_i64, c_vp]),
    "anoddpm_adamw_ema_step": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_f32, c_f32, c_f32, c_int,
                                       c_f32, c_f32, c_vp]),
    "anoddpm_timestep_embedding": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp]),
    "anoddpm_gather_rows": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp]),
    "anoddpm_linear": (c_int, [c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp]),
    "anoddpm_gn_stats": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_vp]),
    "anoddpm_gn_apply": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_f32,
                                 c_int, c_int, c_int, c_vp]),
    "anoddpm_gn_apply_f16in": (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_f32,
                               c_int, c_int, c_vp]),
    "anoddpm_gn_table": (c_int, [c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_f32, c_vp]),
    "anoddpm_conv": (c_int, [ctypes.POINTER(ConvDesc), c_vp]),
    "anoddpm_conv_workspace_bytes": (ctypes.c_longlong, [ctypes.POINTER(ConvDesc)]),
    "anoddpm_conv_tc_supported": (c_int, [ctypes.POINTER(ConvDesc)]),
    "anoddpm_conv_tc3_supported": (c_int, [ctypes.POINTER(ConvDesc)]),
    "anoddpm_conv_stem": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "anoddpm_conv_stem_fused": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "anoddpm_slice_resize_norm": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_i64, c_int, c_int, c_int, c_vp]),
    "anoddpm_conv_head": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "anoddpm_conv_head_gn": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_f32,
                             c_vp]),
    "anoddpm_conv_head_sep_supported": (c_int, [c_int, c_int]),
    "anoddpm_conv_head_sep_workspace_bytes": (ctypes.c_longlong, [c_int, c_int, c_int]),
    "anoddpm_conv_head_sep": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp]),
    "anoddpm_attention": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "anoddpm_attention_tc_supported": (c_int, [c_int, c_int, c_int]),
    "anoddpm_attention_tc_workspace_bytes": (ctypes.c_longlong, [c_int, c_int, c_int]),
    "anoddpm_attention_tc": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def lib():
    """The loaded shared library (loads on first use; raises NativeLibraryError if it is not built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} is missing: build it with `python -m anoddpm_b200.build` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU/PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the header and the library ever disagree
            fn.restype = res
            fn.argtypes = args
        if L.anoddpm_abi_version() != 1:
            raise NativeLibraryError("libanoddpm_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError(f"anoddpm_b200 native call failed ({rc}): {lib().anoddpm_last_error().decode()}")


def call(name, *args):
    check(getattr(lib(), name)(*args))


def ptr(t):
    """Device (or pinned-host) address of a tensor as c_void_p; None -> NULL."""
    if t is None:
        return None
    return c_vp(t.data_ptr())


def stream():
    return c_vp(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"{what} must be a CUDA tensor: anoddpm_b200 runs on sm_100a only (no CPU fallback)")
