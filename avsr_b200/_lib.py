"""ctypes binding of libavsr_b200.so (the C ABI declared in include/avsr_b200.h).

The library is built in-tree (``avsr_b200/libavsr_b200.so``) by ``__graft_entry__.build()`` /
``make -C avsr_b200/csrc``.  There is no CPU or PyTorch fallback: if the library is missing or a call
fails, a RuntimeError is raised (reference convention: print + re-raise, script/evaluation.py:290-294).
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libavsr_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "avsr_b200.h")

_lib: Optional[C.CDLL] = None


class Epilogue(C.Structure):
    _fields_ = [
        ("bias", C.c_void_p), ("bias_mode", C.c_int), ("act", C.c_int), ("prelu", C.c_void_p),
        ("residual", C.c_void_p), ("res_dtype", C.c_int), ("ldr", C.c_longlong),
        ("out_bf16", C.c_void_p), ("ld_bf16", C.c_longlong), ("out_f32", C.c_void_p), ("ld_f32", C.c_longlong),
        ("row_mask", C.c_void_p), ("act_after_residual", C.c_int),
    ]


class BeamState(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("B", "beam", "S", "V", "lmax", "tmax", "blank", "eos", "cap", "no_end_detect")] + [
        (n, C.c_void_p) for n in (
            "utt_T", "step", "n_run", "row_active", "last_tok", "score", "dec_sc", "ctc_sc", "s_prev", "rprev_idx",
            "anc", "hist_tok", "hist_prev", "run2j", "n_ended", "end_step", "end_j", "end_score", "end_dec",
            "end_ctc", "end_len", "best_len", "best_all", "done", "overflow")
    ] + [("d_end", C.c_double), ("utt_maxlen", C.c_void_p)]


ACT_NONE, ACT_GELU, ACT_RELU, ACT_PRELU = 0, 1, 2, 3


def declared_symbols() -> list:
    """Every function name declared in include/avsr_b200.h."""
    src = open(HEADER_PATH).read()
    return sorted(set(re.findall(r"\b(avsr_[a-z0-9_]+)\s*\(", src)))


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the avsr_b200 hot path)")
        lib = C.CDLL(LIB_PATH)
        lib.avsr_last_error.restype = C.c_char_p
        for name in declared_symbols():
            fn = getattr(lib, name)           # AttributeError if the .so does not export a declared symbol
            if name != "avsr_last_error":
                fn.restype = C.c_int
        _lib = lib
    return _lib


launch_count = 0      # kernels enqueued through check() since the caller last reset it (bench.py's gpu_launches)


def check(rc: int, what: str) -> None:
    global launch_count
    launch_count += 1
    if rc != 0:
        msg = load().avsr_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t: torch.Tensor, dtype=None, name: str = "tensor") -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device (avsr_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def ll(v: int) -> C.c_longlong:
    return C.c_longlong(int(v))


def make_epilogue(bias=None, bias_mode=1, act=ACT_NONE, prelu=None, residual=None, ldr=0, out_bf16=None, ld_bf16=0,
                  out_f32=None, ld_f32=0, row_mask=None, act_after_residual=False) -> Epilogue:
    ep = Epilogue()
    ep.bias = bias.data_ptr() if bias is not None else None
    ep.bias_mode = bias_mode
    ep.act = act
    ep.prelu = prelu.data_ptr() if prelu is not None else None
    if residual is not None:
        ep.residual = residual.data_ptr()
        ep.res_dtype = 1 if residual.dtype == torch.bfloat16 else 0
        ep.ldr = ldr
    if out_bf16 is not None:
        ep.out_bf16 = out_bf16.data_ptr()
        ep.ld_bf16 = ld_bf16
    if out_f32 is not None:
        ep.out_f32 = out_f32.data_ptr()
        ep.ld_f32 = ld_f32
    ep.row_mask = row_mask.data_ptr() if row_mask is not None else None
    ep.act_after_residual = 1 if act_after_residual else 0
    return ep


# ---------------------------------------------------------------------------------------------- thin op wrappers
def gemm_bf16(a: torch.Tensor, b: torch.Tensor, M: int, N: int, K: int, ep: Epilogue, lda=None, ldb=None, bn_hint=0):
    """C[M,N] = epilogue(A[M,K] @ B[N,K]^T) on tcgen05 tensor cores."""
    lib = load()
    check(lib.avsr_gemm_bf16_tc(ptr(a), ll(lda if lda is not None else a.stride(0)), ptr(b),
                                ll(ldb if ldb is not None else b.stride(0)), M, N, K, C.byref(ep), bn_hint, stream()),
          "avsr_gemm_bf16_tc")


def conv2d_bf16(x: torch.Tensor, w: torch.Tensor, nf: int, H: int, W: int, Cin: int, Cout: int, ks: int, stride: int, ep: Epilogue):
    """ks x ks / stride / pad ks//2 convolution of NHWC bf16 `x` [nf,H,W,Cin] with w [Cout, ks*ks*Cin] as an implicit GEMM."""
    lib = load()
    check(lib.avsr_conv2d_bf16_tc(ptr(x), ptr(w), ll(nf), H, W, Cin, Cout, ks, stride, C.byref(ep), stream()), "avsr_conv2d_bf16_tc")


def sgemm(a: torch.Tensor, w: torch.Tensor, M: int, N: int, K: int, ep: Epilogue, lda=None, ldw=None):
    lib = load()
    check(lib.avsr_sgemm(ptr(a), ll(lda if lda is not None else a.stride(0)), ptr(w),
                         ll(ldw if ldw is not None else w.stride(0)), M, N, K, C.byref(ep), stream()), "avsr_sgemm")


def layernorm(x: torch.Tensor, gamma, beta, eps: float, out_bf16=None, out_f32=None):
    rows, n = x.shape
    lib = load()
    check(lib.avsr_layernorm(ptr(x), ll(x.stride(0)), ll(rows), n, ptr(gamma), ptr(beta), C.c_float(eps),
                             ptr(out_bf16), ll(out_bf16.stride(0) if out_bf16 is not None else 0),
                             ptr(out_f32), ll(out_f32.stride(0) if out_f32 is not None else 0), stream()), "avsr_layernorm")
