"""ctypes binding of libdrs_b200.so (the C ABI declared in include/drs_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing or a call fails the caller gets an
exception carrying drs_last_error().
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdrs_b200.so")
CSRC = os.path.join(_HERE, "csrc")

DRS_OK = 0
DRS_E_INVALID, DRS_E_MISSING, DRS_E_CUDA, DRS_E_PIPELINE, DRS_E_STATE = -1, -2, -3, -4, -5
MODEL_SUPERRES, MODEL_SAR_TO_NDVI, MODEL_GENERATION = 0, 1, 2
CONV_3x3, CONV_3x3_S2, CONV_1x1, CONV_2x2_S2, CONV_T3x3_S2 = 0, 1, 2, 3, 4


class DrsError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"drs_b200 error {code}: {message}")
        self.code = code


class DrsModelDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("x_channels", C.c_int), ("cond_channels", C.c_int), ("out_channels", C.c_int),
                ("num_classes", C.c_int)]


class DrsTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.POINTER(C.c_float)), ("numel", C.c_int64)]


# name -> (restype, argtypes); must list every symbol include/drs_b200.h declares
SIGNATURES = {
    "drs_last_error": (C.c_char_p, []),
    "drs_version": (C.c_int, []),
    "drs_model_create": (C.c_int, [C.POINTER(DrsModelDesc), C.POINTER(DrsTensor), C.c_int, C.c_int,
                                   C.POINTER(C.c_void_p)]),
    "drs_model_destroy": (None, [C.c_void_p]),
    "drs_plan_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "drs_plan_destroy": (None, [C.c_void_p]),
    "drs_plan_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "drs_plan_check": (C.c_int, [C.c_void_p, C.c_void_p]),
    "drs_cond_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "drs_time_embed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "drs_unet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "drs_sampler_prepare": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                      C.c_void_p]),
    "drs_sampler_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "drs_sampler_step": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "drs_sampler_launches_per_step": (C.c_int, [C.c_void_p]),
    "drs_plan_launch_count": (C.c_int, [C.c_void_p]),
    "drs_plan_launch_info": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_double),
                                       C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "drs_plan_profile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "drs_plan_time_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "drs_sampler_time_hbm_kernels": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]),
    "drs_ddpm_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_size_t,
                                  C.c_void_p]),
    "drs_noise_images": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_size_t,
                                   C.c_void_p]),
    "drs_blend": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                            C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "drs_debug_conv2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "drs_debug_l2_flush": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "drs_debug_timeline": (C.c_int, [C.c_void_p, C.c_int]),
    "drs_debug_spans": (C.c_int, [C.c_void_p, C.c_int]),
    "drs_debug_fetch": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]),
}

# include/drs_b200_diag.h, part 1: the separate diagnostics library (micro-benchmarks; never on the product path)
DIAG_LIB_PATH = os.path.join(_HERE, "libdrs_b200_diag.so")
DIAG_SIGNATURES = {
    "drs_diag_last_error": (C.c_char_p, []),
    "drs_debug_mma_rate": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "drs_debug_mma_rate2": (C.c_int, [C.c_int] * 7 + [C.c_void_p]),
}

_lib: Optional[C.CDLL] = None
_diag: Optional[C.CDLL] = None


def build(verbose: bool = False) -> str:
    """Compiles the CUDA sources for sm_100a with nvcc (cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libdrs_b200.so failed:\n" + out.stdout[-4000:] + out.stderr[-4000:])
    if verbose:
        print(out.stdout[-2000:])
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the sampling path)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def diag_lib() -> C.CDLL:
    """libdrs_b200_diag.so (scripts/diag_mma_rate*.py only)."""
    global _diag
    if _diag is None:
        if not os.path.exists(DIAG_LIB_PATH):
            raise ImportError(f"{DIAG_LIB_PATH} is missing: build it with `make -C {CSRC}`")
        handle = C.CDLL(DIAG_LIB_PATH)
        for name, (res, args) in DIAG_SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _diag = handle
    return _diag


def check_diag(code: int) -> None:
    if code != 0:
        msg = diag_lib().drs_diag_last_error()
        raise DrsError(code, msg.decode("utf-8", "replace") if msg else "unknown error")


def check(code: int) -> None:
    if code != DRS_OK:
        msg = lib().drs_last_error()
        raise DrsError(code, msg.decode("utf-8", "replace") if msg else "unknown error")


def ptr(t) -> Optional[int]:
    """Device / host pointer of a contiguous torch tensor (None passes NULL)."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream
