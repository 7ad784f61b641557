"""Tensor-level wrappers over the C ABI (one Python function per ``mh_*`` entry point).

PyTorch is used for device memory and streams only; every function here launches exactly
the hand-written sm_100a kernels in ``csrc/`` on ``torch.cuda.current_stream()``.
"""
import ctypes
from ctypes import byref, c_float, c_int, c_longlong, c_uint32, c_uint64, c_void_p

import torch

from . import lib as L
from .lib import EPI_ADD, EPI_BF16, EPI_DGELU, EPI_F32, EPI_GELU, EPI_RES  # noqa: F401

bf16 = torch.bfloat16


def _p(t):
    return None if t is None else c_void_p(t.data_ptr())


def _s():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk2d(t, name, dtype=None):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D tensor with unit inner stride, got {tuple(t.shape)} / {t.stride()}")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{name}: expected {dtype}, got {t.dtype}")


def gemm(a, b, out, *, a_mn=False, b_mn=False, epilogue=EPI_BF16, bias=None, aux_in=None, aux_out=None,
         mask=None, p_drop=0.0, seed=0, site=0, block_n=0, split_k=0):
    """out[M,N] (op)= A[M,K] @ B[N,K]^T.  ``a``: [M,K] (or [K,M] when a_mn), ``b``: [N,K] (or
    [K,N] when b_mn); bf16, unit inner stride.  ``out`` bf16 (or fp32 for EPI_F32, accumulated)."""
    _chk2d(a, "a", bf16)
    _chk2d(b, "b", bf16)
    _chk2d(out, "out")
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if K != Kb or out.shape[0] != M or out.shape[1] != N:
        raise ValueError(f"gemm shape mismatch: A {tuple(a.shape)} B {tuple(b.shape)} out {tuple(out.shape)}")
    want = torch.float32 if epilogue == EPI_F32 else bf16
    if out.dtype != want:
        raise ValueError(f"gemm: out dtype {out.dtype}, expected {want}")
    args = L.GemmArgs()
    args.M, args.N, args.K = M, N, K
    args.A, args.lda, args.a_mn = a.data_ptr(), a.stride(0), int(a_mn)
    args.B, args.ldb, args.b_mn = b.data_ptr(), b.stride(0), int(b_mn)
    args.D, args.ldd = out.data_ptr(), out.stride(0)
    args.epilogue = epilogue
    if bias is not None:
        if bias.dtype != torch.float32 or bias.numel() != N:
            raise ValueError("gemm: bias must be fp32 [N]")
        args.bias = bias.data_ptr()
    ld_aux = 0
    for t, nm in ((aux_in, "aux_in"), (aux_out, "aux_out")):
        if t is not None:
            _chk2d(t, nm, bf16)
            if tuple(t.shape) != (M, N):
                raise ValueError(f"gemm: {nm} shape {tuple(t.shape)} != {(M, N)}")
            if ld_aux and ld_aux != t.stride(0):
                raise ValueError("gemm: aux_in / aux_out leading dims differ")
            ld_aux = t.stride(0)
    args.aux_in = None if aux_in is None else aux_in.data_ptr()
    args.aux_out = None if aux_out is None else aux_out.data_ptr()
    args.ld_aux = ld_aux
    if mask is not None:
        if mask.dtype not in (torch.uint8, torch.bool) or tuple(mask.shape) != (M, N) or mask.stride(0) != out.stride(0):
            raise ValueError("gemm: mask must be u8/bool [M,N] with the same leading dim as out")
        args.mask = mask.data_ptr()
    args.p_drop, args.seed, args.site = float(p_drop), int(seed), int(site)
    args.block_n, args.split_k = block_n, split_k
    L.check(L.lib().mh_gemm(byref(args), _s()), "mh_gemm")
    return out
