# -*- coding: utf-8 -*-
"""ctypes binding of libcellvit_b200.so (the C-ABI declared in include/cellvit_b200.h).

There is no CPU fallback: if the shared library is missing, ``lib()`` raises. PyTorch tensors are only
used as device-memory handles (``data_ptr()``) and for the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcellvit_b200.so")
_LIB = None

# enums of csrc/tc_gemm.h
EPI_F16, EPI_RES_F32, EPI_CONVT, EPI_HEAD = 0, 1, 2, 3
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
ROW_IDENTITY, ROW_SEQ, ROW_WINDOW, ROW_TO_WINDOW = 0, 1, 2, 3


class TcEpilogue(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("act", C.c_int),
        ("scale", C.c_void_p), ("shift", C.c_void_p), ("out", C.c_void_p), ("ldc", C.c_longlong),
        ("res", C.c_void_p), ("ldres", C.c_longlong),
        ("res_mod", C.c_int), ("res_off", C.c_int),
        ("row_map", C.c_int), ("row_seq", C.c_int), ("row_pad", C.c_int), ("row_off", C.c_int),
        ("win_size", C.c_int), ("win_grid", C.c_int), ("tok_h", C.c_int), ("tok_w", C.c_int),
        ("ct_cout", C.c_int), ("ct_hin", C.c_int), ("ct_win", C.c_int),
        ("head_w", C.c_void_p), ("head_b", C.c_void_p), ("head_nc", C.c_int), ("head_hw", C.c_int),
        ("head_out", C.c_void_p), ("head_argmax", C.c_void_p), ("head_argmax_nc", C.c_int), ("sched_counter", C.c_void_p),
    ]


class CvbError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise CvbError(f"{LIB_PATH} is missing: run `python -m cellvit_b200.build` (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        L.cvb_last_error.restype = C.c_char_p
        if L.cvb_tc_epilogue_bytes() != C.sizeof(TcEpilogue):
            raise CvbError("TcEpilogue layout mismatch between csrc/tc_gemm.h and _lib.py")
        _LIB = L
    return _LIB


def check(rc: int, what: str = ""):
    if rc != 0:
        raise CvbError(f"{what} failed with status {rc}: {lib().cvb_last_error().decode()}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
