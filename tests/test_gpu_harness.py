"""GPU: batched VSA experiment harness vs the reference's loops restated with the
CPU oracle on the same vectors, plus the expected curve shapes at BASELINE config 5's dimension."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _depth_cell_oracle(vecs):
    from oracle import latent_oracle as O
    out = []
    for v in vecs:                      # scripts/binding_depth_heatmap.py:25-35
        target, partners = v[0:1], v[1:]
        bound = target.clone()
        for k in range(partners.shape[0]):
            bound = O.bind(bound, partners[k:k + 1])
        for k in range(partners.shape[0] - 1, -1, -1):
            bound = O.unbind(bound, partners[k:k + 1])
        out.append(O.similarity(bound, target).mean())
    return torch.stack(out)


@pytest.mark.parametrize("d,m", [(64, 3), (256, 7), (1024, 5), (144, 4)])
def test_depth_cell_matches_oracle_loops(d, m):
    from clifford_b200 import harness
    from oracle import latent_oracle as O
    torch.manual_seed(d + m)
    T = 6
    vecs = O.normalize_vectors(torch.randn(T, m + 1, d) / d ** 0.5)
    got = harness.binding_depth_cell(vecs.to(DEV)).cpu()
    ref = _depth_cell_oracle(vecs)
    assert float((got - ref).abs().max()) < 5e-5
    if d & (d - 1) == 0:
        fused = harness.binding_depth_cell_fused(vecs.to(DEV)).cpu()     # frequency-domain chain, one kernel
        assert float((fused - ref).abs().max()) < 5e-5


def test_rolefiller_cell_matches_oracle_loops():
    from clifford_b200 import harness
    from oracle import latent_oracle as O
    torch.manual_seed(0)
    M, d, k, T = 200, 512, 5, 4
    items = O.normalize_vectors(torch.randn(M, d) / d ** 0.5)
    idx = torch.stack([torch.randperm(M)[:2 * k] for _ in range(T)])
    got = harness.rolefiller_cell(items.to(DEV), idx.to(DEV)).cpu()
    ref = []
    for t in range(T):                  # scripts/rolefiller_heatmap.py:31-43
        roles, fillers = items[idx[t, :k]], items[idx[t, k:]]
        bundled = O.bundle(O.bind(roles, fillers), normalize=True)
        correct = 0
        for ii in range(k):
            rec = O.unbind(bundled.unsqueeze(0), roles[ii].unsqueeze(0)).squeeze()
            correct += int(torch.argmax(O.similarity(rec, items)) == idx[t, k + ii])
        ref.append(correct / k)
    assert torch.allclose(got, torch.tensor(ref))


def test_c5_curve_shapes_d8192():
    """BASELINE config 5: d = 8192, depth sweep.  Unitary / Clifford keys unbind exactly at every depth;
    HRR similarity decays with depth."""
    from clifford_b200 import harness
    from utils import vsa
    torch.manual_seed(4)
    sim_u, depths = harness.run_depth_sweep(vsa.unitary_init, [8192], max_depth=8, n_trials=16, device=DEV)
    assert depths == list(range(1, 9)) and sim_u.shape == (1, 8)
    assert sim_u.min() > 0.999
    sim_c, _ = harness.run_depth_sweep(harness.clifford_init, [4096], max_depth=8, n_trials=16, device=DEV)
    assert sim_c.min() > 0.999
    sim_h, _ = harness.run_depth_sweep(vsa.hrr_init, [8192], max_depth=8, n_trials=16, device=DEV)
    assert sim_h[0, 0] > sim_h[0, 3] > sim_h[0, 7] and sim_h[0, 0] < 0.95
    acc = harness.run_rolefiller_sweep(vsa.hrr_init, [1024], [2, 8, 600], n_items=1000, n_trials=8, device=DEV)
    assert acc[0, 0] > 0.95 and np.isnan(acc[0, 2]) and acc[0, 1] <= acc[0, 0] + 1e-6
