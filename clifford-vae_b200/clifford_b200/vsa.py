"""HRR / VSA ops with the reference's function signatures (reference utils/vsa.py:9-96), backed by
the fused FFT-domain kernels.

Device policy.  Every op runs on a CUDA (sm_100a) device; there is no CPU implementation.  The reference's
harnesses hand these functions CPU tensors in places (utils/vsa.py:266-267,278 pins the item memory and the
random roles to the CPU; utils/wandb_utils.py:165 builds its baselines with ``device="cpu"``), so a CPU argument is
*staged*: copied to the CUDA device of the other operand (or the current CUDA device), computed there, and the result
is returned on the device the reference would have returned it on.  Without a CUDA device the call raises
CliffordB200Error -- never a host computation.  The autograd functions in ``ops`` stay strict (CUDA tensors only).
"""
from __future__ import annotations

import math

import torch

from . import ops
from ._lib import CliffordB200Error


def _compute_device(*tensors) -> torch.device:
    for t in tensors:
        if torch.is_tensor(t) and t.device.type == "cuda":
            return t.device
    if not torch.cuda.is_available():
        raise CliffordB200Error(
            "clifford_b200 ops run only on a CUDA (sm_100a) device and no CUDA device is available. "
            "There is deliberately no CPU fallback.")
    return torch.device("cuda", torch.cuda.current_device())


def _stage(*tensors):
    """-> (tensors on one CUDA device, device the result goes back to | None when nothing was staged)."""
    if all(t.device.type == "cuda" for t in tensors):
        return tensors, None
    dev = _compute_device(*tensors)
    return tuple(t.to(dev) for t in tensors), tensors[0].device


def _home(x: torch.Tensor, home) -> torch.Tensor:
    return x if home is None else x.to(home)


def _init_device(device):
    """(compute device, device to return on | None) for the generators' ``device`` argument."""
    dev = torch.device(device)
    if dev.type == "cuda":
        return dev, None
    return _compute_device(), dev


def hrr_init(n: int, d: int, device="cpu", dtype=torch.float32) -> torch.Tensor:
    """n vectors ~ N(0, 1/d) (utils/vsa.py:9-12)."""
    dev, home = _init_device(device)
    return _home(ops.hrr_init(n, d, dev).to(dtype), home)


def unitary_init(n: int, d: int, device="cpu", dtype=torch.float32, eps=1e-3) -> torch.Tensor:
    """n vectors with unit Fourier magnitude (utils/vsa.py:15-36): one batched kernel instead of a
    Python loop of n tiny iffts."""
    dev, home = _init_device(device)
    return _home(ops.unitary_init(n, d, dev, eps).to(dtype), home)


def normalize_vectors(x: torch.Tensor) -> torch.Tensor:
    (x,), home = _stage(x)
    return _home(ops.Normalize.apply(x), home)


def bind(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """circular convolution (utils/vsa.py:43-46)."""
    (a, b), home = _stage(a, b)
    return _home(ops.Bind.apply(a, b, ops.BIND_MUL), home)


def invert(a: torch.Tensor) -> torch.Tensor:
    """[a0, a1, ..., a_{n-1}] -> [a0, a_{n-1}, ..., a1] (utils/vsa.py:49-53)."""
    (a,), home = _stage(a)
    return _home(ops.Invert.apply(a), home)


def unbind(ab: torch.Tensor, b: torch.Tensor, method: str = "inv") -> torch.Tensor:
    """utils/vsa.py:56-72.  'inv'/'*' multiplies by conj(FFT b) (== bind with invert(b)) in the same
    fused kernel; 'dagger'/'deconv' divides by FFT(b) + 1e-12."""
    if method == "inv" or method == "*":
        mode = ops.BIND_MUL_CONJ
    elif method == "†" or method == "deconv":
        mode = ops.BIND_DIV
    else:
        raise ValueError(f"unsupported unbind method: {method}")
    (ab, b), home = _stage(ab, b)
    return _home(ops.Bind.apply(ab, b, mode), home)


def bundle(vectors: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    """sum over dim 0, optionally / sqrt(k) (utils/vsa.py:75-79)."""
    scale = 1.0 / math.sqrt(vectors.shape[0]) if normalize else 1.0
    (vectors,), home = _stage(vectors)
    return _home(ops.Bundle.apply(vectors, scale), home)


def permute_vector(v: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    (v,), home = _stage(v)
    return _home(ops.Permute.apply(v, perm.to(v.device), False), home)


def unpermute_vector(v: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    (v,), home = _stage(v)
    return _home(ops.Permute.apply(v, perm.to(v.device), True), home)


def similarity(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """cosine similarity over the last dim (utils/vsa.py:93-96); the result lives on a's device."""
    home = None if a.device.type == "cuda" else a.device
    if a.device.type != "cuda":
        a = a.to(_compute_device(b))
    if a.device != b.device:
        b = b.to(a.device)
    return _home(ops.Cosine.apply(a, b), home)
