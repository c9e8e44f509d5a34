"""ctypes binding of libspoofsv_b200.so (the C ABI in include/spoofsv_b200.h).

There is no fallback: if the library is missing it is built with nvcc, and if
that is impossible the import of any compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libspoofsv_b200.so"

PREC_FP32 = 0          # FP32-accurate: tensor cores (3xTF32 split) where built, CUDA cores elsewhere
PREC_BF16 = 1
PREC_FP32_FFMA = 2     # FP32-accurate on the CUDA cores only

_c_f32p = C.c_void_p
_c_i64p = C.c_void_p

# name -> (restype, argtypes).  Must list every symbol include/spoofsv_b200.h declares
# (tests/test_abi.py cross-checks this table against the header).
SIGNATURES = {
    "ssv_version": (C.c_int, []),
    "ssv_last_error": (C.c_char_p, []),
    "ssv_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "ssv_launch_count": (C.c_long, [C.c_int]),
    "ssv_highway_conv_fwd": (C.c_int, [_c_f32p] * 7 + [C.c_int] * 6 + [_c_f32p, C.c_int, C.c_void_p]),
    "ssv_text2mel_create": (C.c_int, [C.POINTER(C.c_char_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
                            + [C.c_int] * 6 + [C.POINTER(C.c_void_p)]),
    "ssv_text2mel_destroy": (C.c_int, [C.c_void_p]),
    "ssv_text_encoder_fwd": (C.c_int, [C.c_void_p, _c_i64p, C.c_int, C.c_int, _c_f32p, _c_f32p, C.c_int, C.c_void_p]),
    "ssv_text2mel_train_fwd": (C.c_int, [C.c_void_p, _c_f32p, _c_i64p, _c_f32p, C.c_int, C.c_int, C.c_int, _c_f32p, _c_f32p,
                                         C.c_int, C.c_void_p]),
    "ssv_decoder_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "ssv_decoder_destroy": (C.c_int, [C.c_void_p]),
    "ssv_decoder_begin": (C.c_int, [C.c_void_p, _c_f32p, _c_f32p, _c_f32p, C.c_int, C.c_int, _c_f32p, _c_f32p,
                                    _c_i64p, C.c_int, C.c_void_p]),
    "ssv_decoder_step": (C.c_int, [C.c_void_p, _c_f32p, C.c_long, C.c_long, _c_i64p, C.c_void_p]),
    "ssv_decoder_run": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "ssv_decoder_frames": (C.c_int, [C.c_void_p]),
    "ssv_decoder_set_plan": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "ssv_text2mel_check": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ssv_decoder_check": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ssv_ssrn_create": (C.c_int, [C.POINTER(C.c_char_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
                        + [C.c_int] * 4 + [C.POINTER(C.c_void_p)]),
    "ssv_ssrn_destroy": (C.c_int, [C.c_void_p]),
    "ssv_ssrn_fwd": (C.c_int, [C.c_void_p, _c_f32p, C.c_long, C.c_long, C.c_long, C.c_int, C.c_int, _c_f32p,
                               C.c_int, C.c_void_p]),
    "ssv_synthesize_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _c_i64p, _c_f32p, C.c_int, C.c_int, C.c_int,
                                      _c_f32p, _c_f32p, _c_f32p, _c_i64p, C.c_int, C.c_int, C.c_void_p]),
    "ssv_synthesize_host_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _c_i64p, _c_f32p, C.c_int, C.c_int, C.c_int,
                                             _c_f32p, _c_f32p, _c_f32p, _c_i64p, C.c_int, C.c_int, C.c_void_p,
                                             C.POINTER(C.c_int)]),
    "ssv_synthesize_host_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "ssv_decoder_set_lin_output": (C.c_int, [C.c_void_p, C.c_int]),
    "ssv_deemphasis": (C.c_int, [_c_f32p, _c_f32p, C.c_int, C.c_long, C.c_float, C.c_void_p]),
    "ssv_highway_conv_bwd": (C.c_int, [_c_f32p] * 8 + [C.c_int] * 6 + [_c_f32p] * 8 + [C.c_int, C.c_int, C.c_void_p]),
    "ssv_highway_conv_fwd_save": (C.c_int, [_c_f32p] * 7 + [C.c_int] * 6 + [_c_f32p] * 2 + [C.c_int, C.c_int, C.c_void_p]),
    "ssv_conv_ln_fwd_save": (C.c_int, [_c_f32p] * 6 + [C.c_int] * 5 + [_c_f32p] * 2 + [C.c_int, C.c_void_p]),
    "ssv_conv_ln_bwd": (C.c_int, [_c_f32p] * 5 + [C.c_int] * 5 + [_c_f32p] * 6 + [C.c_int, C.c_void_p]),
    "ssv_attention_train_fwd": (C.c_int, [_c_f32p] * 2 + [C.c_int] * 3 + [_c_f32p] * 2 + [C.c_void_p]),
    "ssv_attention_train_bwd": (C.c_int, [_c_f32p] * 5 + [C.c_int] * 3 + [_c_f32p] * 2 + [C.c_void_p]),
    "ssv_text_embedding_fwd": (C.c_int, [_c_i64p, _c_f32p, _c_f32p] + [C.c_int] * 4 + [_c_f32p, C.c_void_p]),
    "ssv_text_embedding_bwd": (C.c_int, [_c_i64p, _c_f32p] + [C.c_int] * 4 + [_c_f32p, _c_f32p, C.c_void_p]),
    "ssv_linear_small_fwd": (C.c_int, [_c_f32p] * 3 + [C.c_int] * 3 + [_c_f32p, C.c_void_p]),
    "ssv_linear_small_bwd": (C.c_int, [_c_f32p] * 2 + [C.c_int] * 3 + [_c_f32p] * 2 + [C.c_void_p]),
    "ssv_griffin_lim_workspace": (C.c_longlong, [C.c_int, C.c_int]),
    "ssv_griffin_lim": (C.c_int, [_c_f32p, _c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                  _c_f32p, _c_f32p, C.c_longlong, C.c_void_p]),
    "ssv_spec_features": (C.c_int, [_c_f32p, C.c_int, C.c_int, _c_f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                    C.c_int, _c_f32p, _c_f32p, _c_f32p, C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


class SsvError(RuntimeError):
    pass


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first if needed) the CUDA library.  Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    import os
    override = os.environ.get("SSV_B200_LIB")          # development aid: an A/B build from `build.py --variant=...`
    if override:
        lib = C.CDLL(override)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib
    if build_if_missing:
        from . import build as _build
        try:
            _build.build()
        except Exception as e:  # a prebuilt, up-to-date .so is still fine
            if not LIB_PATH.exists():
                raise SsvError(f"libspoofsv_b200.so is missing and could not be built: {e}") from e
    if not LIB_PATH.exists():
        raise SsvError(f"{LIB_PATH} not found; run `python -m spoofsv_b200.build`")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().ssv_last_error().decode("utf-8", "replace")
        if status == 1:
            raise ValueError(f"spoofsv_b200: {msg}")
        raise SsvError(f"spoofsv_b200 (status {status}): {msg}")


def current_stream_ptr() -> int:
    import torch
    return int(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, what: str) -> None:
    if not t.is_cuda:
        raise SsvError(
            f"{what}: tensor is on {t.device}; spoofsv_b200 has no CPU path (CUDA sm_100a kernels only)")


def pack_params(state: dict):
    """state_dict-like {name: cuda fp32 tensor} -> ctypes arrays (names, ptrs, numels) + keepalive."""
    import torch
    names, ptrs, numels, keep = [], [], [], []
    for k, v in state.items():
        t = v.detach()
        require_cuda(t, f"parameter {k}")
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(torch.float32).contiguous()
        keep.append(t)
        names.append(k.encode())
        ptrs.append(t.data_ptr())
        numels.append(t.numel())
    n = len(names)
    return ((C.c_char_p * n)(*names), (C.c_void_p * n)(*ptrs), (C.c_int64 * n)(*numels), n, keep)
