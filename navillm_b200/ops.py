"""Python-side launchers for the sm_100a kernels (thin: argument checking + C-ABI call).

Every function here writes into caller-provided or freshly ``torch.empty``-allocated CUDA tensors and
launches on the current stream.  Nothing falls back to torch math.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, f32, i32, i64, ptr, stream_ptr, u32

bf16 = torch.bfloat16


def _rowmajor(t: torch.Tensor, name: str) -> None:
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D tensor with unit inner stride, got shape {tuple(t.shape)} strides {t.stride()}")
    if not t.is_cuda:
        raise _lib.NvError(f"{name}: navillm_b200 kernels need CUDA tensors (no CPU fallback)")


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False, out: torch.Tensor | None = None,
         addend: torch.Tensor | None = None, out_f32: bool = False, block_n: int = 0) -> torch.Tensor:
    """C[M,N] = A·B (+ addend) on tcgen05 (csrc/gemm_bf16.cu).

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True);  b: [N,K] (b_mn=False) or [K,N] (b_mn=True); bf16.
    """
    _rowmajor(a, "a"); _rowmajor(b, "b")
    assert a.dtype == bf16 and b.dtype == bf16
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if K != Kb:
        raise ValueError(f"gemm: contraction mismatch {K} vs {Kb}")
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32 if out_f32 else bf16, device=a.device)
    _rowmajor(out, "out")
    assert out.shape == (M, N) and out.dtype == (torch.float32 if out_f32 else bf16)
    flags = 0
    if addend is not None:
        _rowmajor(addend, "addend")
        assert addend.shape == (M, N) and addend.dtype == bf16
        flags |= _lib.GEMM_ADD
    if out_f32:
        flags |= _lib.GEMM_OUT_F32
    lib = _lib.load()
    check(lib.nv_gemm_bf16(ptr(a), i64(a.stride(0)), i32(a_mn), ptr(b), i64(b.stride(0)), i32(b_mn), ptr(out),
                           i64(out.stride(0)), ptr(addend), i64(addend.stride(0) if addend is not None else 0),
                           i32(M), i32(N), i32(K), u32(flags), i32(block_n), stream_ptr()), "nv_gemm_bf16")
    return out
