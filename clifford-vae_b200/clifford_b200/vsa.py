"""HRR / VSA ops with the reference's function signatures (reference utils/vsa.py:9-96), backed by
the fused FFT-domain kernels.  CUDA tensors only: the reference's experiment harnesses that pin the
item memory to the CPU (utils/vsa.py:266-267) must move it to the GPU -- there is no CPU fallback.
"""
from __future__ import annotations

import math

import torch

from . import ops


def hrr_init(n: int, d: int, device="cuda", dtype=torch.float32) -> torch.Tensor:
    """n vectors ~ N(0, 1/d) (utils/vsa.py:9-12)."""
    return ops.hrr_init(n, d, device).to(dtype)


def unitary_init(n: int, d: int, device="cuda", dtype=torch.float32, eps=1e-3) -> torch.Tensor:
    """n vectors with unit Fourier magnitude (utils/vsa.py:15-36): one batched kernel instead of a
    Python loop of n tiny iffts."""
    return ops.unitary_init(n, d, device, eps).to(dtype)


def normalize_vectors(x: torch.Tensor) -> torch.Tensor:
    return ops.Normalize.apply(x)


def bind(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """circular convolution (utils/vsa.py:43-46)."""
    return ops.Bind.apply(a, b, ops.BIND_MUL)


def invert(a: torch.Tensor) -> torch.Tensor:
    """[a0, a1, ..., a_{n-1}] -> [a0, a_{n-1}, ..., a1] (utils/vsa.py:49-53)."""
    return ops.Invert.apply(a)


def unbind(ab: torch.Tensor, b: torch.Tensor, method: str = "inv") -> torch.Tensor:
    """utils/vsa.py:56-72.  'inv'/'*' multiplies by conj(FFT b) (== bind with invert(b)) in the same
    fused kernel; 'dagger'/'deconv' divides by FFT(b) + 1e-12."""
    if method == "inv" or method == "*":
        return ops.Bind.apply(ab, b, ops.BIND_MUL_CONJ)
    elif method == "†" or method == "deconv":
        return ops.Bind.apply(ab, b, ops.BIND_DIV)
    else:
        raise ValueError(f"unsupported unbind method: {method}")


def bundle(vectors: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    """sum over dim 0, optionally / sqrt(k) (utils/vsa.py:75-79)."""
    scale = 1.0 / math.sqrt(vectors.shape[0]) if normalize else 1.0
    return ops.Bundle.apply(vectors, scale)


def permute_vector(v: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    return ops.Permute.apply(v, perm.to(v.device), False)


def unpermute_vector(v: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    return ops.Permute.apply(v, perm.to(v.device), True)


def similarity(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """cosine similarity over the last dim (utils/vsa.py:93-96)."""
    if a.device != b.device:
        b = b.to(a.device)
    return ops.Cosine.apply(a, b)
