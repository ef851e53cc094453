"""ctypes binding of libviewneti_sm100a.so (C-ABI declared in include/viewneti.h).

Host code stays Python: PyTorch owns device memory and streams, tensors cross the boundary as raw
device pointers + sizes.  There is NO CPU fallback: if the library is missing or a call fails,
`VNError` is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

import torch

import os as _os

# VN_LIB_SUFFIX selects an experiment build of the same ABI (see build.py); the default is the product library
_LIB_PATH = Path(__file__).resolve().parent / "lib" / f"libviewneti_sm100a{_os.environ.get('VN_LIB_SUFFIX', '')}.so"
_lib: Optional[C.CDLL] = None


class VNError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("nb", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64),
        ("B", C.c_void_p), ("ldb", C.c_int64),
        ("D", C.c_void_p), ("ldd", C.c_int64),
        ("bias", C.c_void_p),
        ("rowbias", C.c_void_p), ("ld_rowbias", C.c_int64), ("rows_per_batch", C.c_int32),
        ("R", C.c_void_p), ("ldr", C.c_int64),
        ("out_fp32", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("force_bn", C.c_int32), ("force_split", C.c_int32),
        ("r_fp32", C.c_int32),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("nb", C.c_int32), ("heads", C.c_int32), ("nq", C.c_int32), ("nk", C.c_int32),
        ("scale", C.c_float),
        ("q", C.c_void_p), ("ldq", C.c_int64), ("bsq", C.c_int64),
        ("k", C.c_void_p), ("ldk", C.c_int64), ("bsk", C.c_int64),
        ("v", C.c_void_p), ("ldv", C.c_int64), ("bsv", C.c_int64),
        ("o", C.c_void_p), ("ldo", C.c_int64), ("bso", C.c_int64),
        ("lse", C.c_void_p),
        ("d_o", C.c_void_p), ("lddo", C.c_int64), ("bsdo", C.c_int64),
        ("delta", C.c_void_p),
        ("dq", C.c_void_p), ("lddq", C.c_int64), ("bsdq", C.c_int64),
        ("dk", C.c_void_p), ("lddk", C.c_int64), ("bsdk", C.c_int64),
        ("dv", C.c_void_p), ("lddv", C.c_int64), ("bsdv", C.c_int64),
        ("dkv_acc", C.c_void_p),
        ("causal", C.c_int32),
        ("ws", C.c_void_p), ("ws_bytes", C.c_int64),
        ("defer_dkv_finish", C.c_int32),
    ]


_P, _I, _L, _F, _Z = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol include/viewneti.h declares
SIGNATURES = {
    "vn_version": (C.c_int, []),
    "vn_last_error": (C.c_char_p, []),
    "vn_launch_count": (C.c_int64, []),
    "vn_launch_count_reset": (None, []),
    "vn_set_pdl": (None, [C.c_int]),
    "vn_gemm_workspace_bytes": (C.c_size_t, [_I, _I]),
    "vn_gemm": (C.c_int, [C.POINTER(GemmDesc), _P]),
    "vn_set_debug_buffer": (None, [_P]),
    "vn_groupnorm_stats": (C.c_int, [_P, _L, _I, _I, _I, _I, _P, _P]),
    "vn_groupnorm_apply": (C.c_int, [_P, _L, _P, _P, _P, _F, _I, _P, _L, _I, _I, _I, _I, _P]),
    "vn_groupnorm_bwd_stats": (C.c_int, [_P, _L, _P, _L, _P, _P, _P, _F, _I, _P, _I, _I, _I, _I, _P]),
    "vn_groupnorm_bwd_apply": (C.c_int, [_P, _L, _P, _L, _P, _P, _P, _P, _F, _I, _P, _L, _P, _L, _P, _L,
                                         _I, _I, _I, _I, _P]),
    "vn_groupnorm_fwd": (C.c_int, [_P, _L, _P, _P, _F, _I, _P, _L, _I, _I, _I, _I, _P, _P, _P]),
    "vn_groupnorm_bwd": (C.c_int, [_P, _L, _P, _L, _P, _P, _P, _F, _I, _P, _L, _P, _L, _P, _L, _I, _I, _I, _I, _P, _P,
                                   _P]),
    "vn_groupnorm_partial_floats": (C.c_size_t, [_I]),
    "vn_set_groupnorm_fused": (None, [C.c_int]),
    "vn_layernorm_fwd": (C.c_int, [_P, _L, _P, _P, _F, _P, _L, _P, _I, _I, _P]),
    "vn_layernorm_bwd": (C.c_int, [_P, _L, _P, _L, _P, _P, _P, _L, _P, _L, _I, _I, _P]),
    "vn_layernorm_fwd_f32": (C.c_int, [_P, _L, _P, _P, _F, _P, _L, _P, _I, _I, _P]),
    "vn_layernorm_bwd_f32": (C.c_int, [_P, _L, _P, _L, _P, _P, _P, _L, _P, _L, _P, _L, _I, _I, _P]),
    "vn_nhwc_to_nchw_thin": (C.c_int, [_P, _L, _P, _I, _I, _L, _P]),
    "vn_attention_fwd_workspace_bytes": (C.c_size_t, [_I, _I, _I, _I]),
    "vn_attention_bwd_workspace_bytes": (C.c_size_t, [_I, _I, _I, _I, _I]),
    "vn_geglu_fwd": (C.c_int, [_P, _L, _P, _L, _I, _I, _P]),
    "vn_geglu_bwd": (C.c_int, [_P, _L, _P, _L, _P, _L, _I, _I, _P]),
    "vn_gelu_fwd": (C.c_int, [_P, _L, _P, _L, _I, _I, _P]),
    "vn_gelu_bwd": (C.c_int, [_P, _L, _P, _L, _P, _L, _I, _I, _P]),
    "vn_seq_attention_fwd": (C.c_int, [C.POINTER(AttnDesc), _I, _P]),
    "vn_seq_attention_bwd": (C.c_int, [C.POINTER(AttnDesc), _I, _P]),
    "vn_attention_fwd": (C.c_int, [C.POINTER(AttnDesc), _P]),
    "vn_attention_bwd": (C.c_int, [C.POINTER(AttnDesc), _P]),
    "vn_attention_dkv_finish": (C.c_int, [C.POINTER(AttnDesc), _P]),
    "vn_upsample2x_fwd": (C.c_int, [_P, _L, _P, _L, _I, _I, _I, _I, _P]),
    "vn_upsample2x_bwd": (C.c_int, [_P, _L, _P, _L, _I, _I, _I, _I, _P]),
    "vn_im2col_s2": (C.c_int, [_P, _L, _P, _I, _I, _I, _I, _P]),
    "vn_im2col_s2_pad0": (C.c_int, [_P, _L, _P, _I, _I, _I, _I, _P]),
    "vn_im2col_thin": (C.c_int, [_P, _P, _L, _I, _I, _I, _I, _P]),
    "vn_softmax_rows": (C.c_int, [_P, _L, _P, _L, _I, _I, _F, _P]),
    "vn_col2im_s2": (C.c_int, [_P, _P, _L, _P, _L, _I, _I, _I, _I, _P]),
    "vn_conv_in_fwd": (C.c_int, [_P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _P]),
    "vn_conv_out_fwd": (C.c_int, [_P, _L, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "vn_conv_out_bwd": (C.c_int, [_P, _P, _P, _L, _I, _I, _I, _I, _I, _P]),
    "vn_timestep_sinusoid": (C.c_int, [_P, _P, _I, _I, _P]),
    "vn_gemv": (C.c_int, [_P, _L, _P, _P, _P, _L, _I, _I, _I, _I, _P]),
    "vn_cast_f32_bf16": (C.c_int, [_P, _P, _L, _P]),
    "vn_cast_bf16_f32": (C.c_int, [_P, _P, _L, _P]),
    "vn_copy2d": (C.c_int, [_P, _L, _P, _L, _P, _L, _L, _I, _P]),
    "vn_mse_loss": (C.c_int, [_P, _P, _L, _F, _P, _P, _P]),
    "vn_cfg_dpmpp_step": (C.c_int, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _F, _P]),
    "vn_cfg_ddim_step": (C.c_int, [_P, _P, _P, _L, _F, _F, _F, _I, _P]),
    "vn_memset_zero": (C.c_int, [_P, _Z, _P]),
    "vn_memset": (C.c_int, [_P, _I, _Z, _P]),
    "vn_mapper_param_count": (C.c_int, [_I]),
    "vn_mapper_saved_floats": (C.c_int, [_I]),
    "vn_mapper_fwd": (C.c_int, [_P, _P, _P, _F, _P, _P, _P, _I, _I, _I, _P]),
    "vn_mapper_bwd": (C.c_int, [_P, _P, _P, _P, _F, _P, _P, _I, _I, _P]),
    "vn_adamw_step": (C.c_int, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _F, _P]),
}


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """dlopen the library and bind every declared symbol (works on a CPU-only box: symbols only)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise VNError(f"{_LIB_PATH} is missing - run `python -m view_neti_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise VNError(f"{_LIB_PATH.name} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if lib.vn_version() != 1:
        raise VNError(f"ABI version mismatch: library {lib.vn_version()}, binding 1")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().vn_last_error().decode("utf-8", "replace")
        raise VNError(f"{what or 'viewneti call'} failed (rc={rc}): {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise VNError("viewneti kernels need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream
