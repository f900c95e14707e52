"""Batched device versions of the reference's VSA experiment loops.

The reference runs these as Python triple loops with one tiny FFT launch and one `.item()` host sync
per query (scripts/binding_depth_heatmap.py:16-39, scripts/rolefiller_heatmap.py:17-44).  Here all
trials of one (dimension, depth | k) cell go through the fused bind / unbind kernels as one batch and
the cleanup is one GEMM + argmax; only the per-cell mean leaves the device.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np
import torch

from . import vsa


def clifford_init(n: int, d: int, device="cuda", dtype=torch.float32) -> torch.Tensor:
    """n vectors of length 2d on the Clifford torus (unit-magnitude spectrum, DC = Nyquist = 1): the
    uniform prior's sampler.  (scripts/bundle_heatmap.py:16-29 builds the same family but leaves a
    1-radian phase on the DC / Nyquist bins.)"""
    from .distributions import CliffordTorusUniform
    return CliffordTorusUniform(d, device=device, dtype=dtype).rsample((n,))


def binding_depth_cell(vecs: torch.Tensor) -> torch.Tensor:
    """vecs (T, m+1, d): per trial bind vecs[:,0] with partners 1..m in order, unbind them in reverse
    order, return cos(recovered, target) (T,).  (run_depth_sweep's inner trial loop, batched over T.)"""
    target = vecs[:, 0].contiguous()
    bound = target
    m = vecs.shape[1] - 1
    for k in range(1, m + 1):
        bound = vsa.bind(bound, vecs[:, k].contiguous())
    for k in range(m, 0, -1):
        bound = vsa.unbind(bound, vecs[:, k].contiguous())
    return vsa.similarity(bound, target)


def binding_depth_cell_fused(vecs: torch.Tensor) -> torch.Tensor:
    """Same result as binding_depth_cell, in ONE kernel: the whole bind/unbind chain of a trial is evaluated in the
    frequency domain (X0 * prod_j |Y_j|^2) and the cosine by Parseval -- every vector is read once and nothing but
    the (T,) similarities is written.  Power-of-two d in [32, 16384]; no autograd (an evaluation workload)."""
    from . import _lib, ops
    T, mp1, d = vecs.shape
    v = vecs.detach().float().contiguous()
    _lib.ensure_device(v.device)
    ops._CUR_DEV[0] = v.device
    ops._EMPTY[0] = T == 0
    out = torch.empty(T, device=v.device, dtype=torch.float32)
    ops._launch("cvb_vsa_depth_chain_cosine", v.data_ptr(), out.data_ptr(), T, mp1, d)
    return out


def run_depth_sweep(init_fn: Callable, dims: Sequence[int], max_depth: int = 40, n_trials: int = 10, device="cuda",
                    fused_chain: bool = True):
    """Same signature/return as the reference's run_depth_sweep minus the label: (sim_matrix, depths)."""
    depths = list(range(1, max_depth + 1))
    sim = np.full((len(dims), len(depths)), np.nan)
    for i, d in enumerate(dims):
        cells = []
        for m in depths:
            vecs = vsa.normalize_vectors(init_fn(n_trials * (m + 1), d, device=device))
            vecs = vecs.view(n_trials, m + 1, vecs.shape[-1])
            dd = vecs.shape[-1]
            fused = fused_chain and dd >= 32 and dd <= 16384 and (dd & (dd - 1)) == 0
            cells.append((binding_depth_cell_fused(vecs) if fused else binding_depth_cell(vecs)).mean())
        sim[i] = torch.stack(cells).cpu().numpy()            # one host sync per dimension
    return sim, depths


def rolefiller_cell(items: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """items (M, d) normalised item memory, idx (T, 2k) distinct item indices per trial: roles are
    idx[:, :k], fillers idx[:, k:].  Returns the per-trial accuracy (T,) of recovering every filler from
    the bundled role-filler pairs by unbinding with its role and cleaning up against the item memory."""
    T, k2 = idx.shape
    k = k2 // 2
    roles = items[idx[:, :k].T.contiguous()]                  # (k, T, d)
    fillers = items[idx[:, k:].T.contiguous()]
    pairs = vsa.bind(roles, fillers)                          # (k, T, d): k*T rows in one launch
    bundled = vsa.bundle(pairs, normalize=True)               # (T, d)
    recovered = vsa.unbind(bundled.unsqueeze(0), roles)       # (k, T, d)
    # cleanup: cosine against the whole item memory is a (kT x d) x (d x M) GEMM (cuBLAS) + argmax
    rec = vsa.normalize_vectors(recovered.reshape(k * T, -1))
    best = (rec @ items.T).argmax(dim=1).view(k, T)
    return (best == idx[:, k:].T).float().mean(dim=0)


def run_rolefiller_sweep(init_fn: Callable, dims: Sequence[int], k_range: Sequence[int], n_items: int = 1000,
                         n_trials: int = 10, device="cuda", generator: Optional[torch.Generator] = None):
    acc = np.full((len(dims), len(k_range)), np.nan)
    for i, d in enumerate(dims):
        items = vsa.normalize_vectors(init_fn(n_items, d, device=device))
        cells, cols = [], []
        for j, k in enumerate(k_range):
            if 2 * k > n_items:
                continue
            idx = torch.stack([torch.randperm(n_items, generator=generator)[:2 * k] for _ in range(n_trials)]).to(device)
            cells.append(rolefiller_cell(items, idx).mean())
            cols.append(j)
        if cells:
            acc[i, cols] = torch.stack(cells).cpu().numpy()
    return acc
