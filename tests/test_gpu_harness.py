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


def test_bundle_capacity_cell_matches_oracle_loops():
    from clifford_b200 import harness
    from oracle import latent_oracle as O
    torch.manual_seed(1)
    M, d, k, T = 120, 256, 6, 5
    items = O.normalize_vectors(torch.randn(M, d) / d ** 0.5)
    idx = torch.stack([torch.randperm(M)[:2 * k] for _ in range(T)])
    got = harness.bundle_capacity_cell(items.to(DEV), idx.to(DEV)).cpu()
    ref = []
    for t in range(T):                  # utils/vsa.py:138-160
        X, Xp = items[idx[t, :k]], items[idx[t, k:]]
        C1, C2 = O.bundle(X, normalize=True), O.bundle(Xp, normalize=True)
        ref.append(float((O.similarity(X, C1.unsqueeze(0).expand(k, -1)) >
                          O.similarity(X, C2.unsqueeze(0).expand(k, -1))).float().mean()))
    assert torch.allclose(got, torch.tensor(ref))
    res = harness.run_bundle_capacity(d=512, n_items=200, k_range=[2, 40], n_trials=8, device=DEV)
    assert res["k"] == [2, 40] and res["accuracy"][0] == 1.0 and 0.5 < res["accuracy"][1] <= 1.0


@pytest.mark.parametrize("method,braid,random_roles", [("inv", False, True), ("†", False, True), ("inv", True, True),
                                                         ("inv", False, False)])
def test_binding_pairs_cell_matches_oracle_loops(method, braid, random_roles):
    from clifford_b200 import harness
    from oracle import latent_oracle as O
    torch.manual_seed(2)
    M, d, k, T = 150, 512, 4, 3
    items = O.normalize_vectors(torch.randn(M, d) / d ** 0.5)
    if random_roles:
        fidx = torch.stack([torch.randperm(M)[:k] for _ in range(T)])
        a, r = torch.rand(k * T, d // 2 - 1), torch.rand(k * T, d // 2 - 1)
        roles = O.normalize_vectors(O.unitary_init_from_uniform(a, r, d)).view(k, T, d)
    else:
        idx = torch.stack([torch.randperm(M)[:2 * k] for _ in range(T)])
        fidx, roles = idx[:, k:], items[idx[:, :k].T]
    perms = torch.stack([torch.randperm(d) for _ in range(k * T)]).view(k, T, d) if braid else None
    got = harness.binding_pairs_cell(items.to(DEV), fidx.to(DEV), roles.to(DEV), method,
                                     None if perms is None else perms.to(DEV)).cpu()
    ref = []
    for t in range(T):                  # utils/vsa.py:273-322
        pairs = O.bind(roles[:, t], items[fidx[t]])
        if braid:
            pairs = torch.stack([O.permute_vector(pairs[i], perms[i, t]) for i in range(k)])
        bundled = O.bundle(pairs, normalize=True)
        correct = 0
        for i in range(k):
            src = O.unpermute_vector(bundled, perms[i, t]) if braid else bundled
            rec = O.unbind(src.unsqueeze(0), roles[i, t].unsqueeze(0), method=method).squeeze()
            correct += int(torch.argmax(O.similarity(rec, items)) == fidx[t, i])
        ref.append(correct / k)
    assert torch.allclose(got, torch.tensor(ref))


def test_run_binding_unbinding_pairs_contract():
    from clifford_b200 import harness
    torch.manual_seed(3)
    res = harness.run_binding_unbinding_pairs(d=1024, n_items=300, k_range=[2, 6], n_trials=6, device=DEV)
    assert res["k"] == [2, 6] and res["accuracy"][0] > 0.9 and len(res["std"]) == 2
    res_b = harness.run_binding_unbinding_pairs(d=1024, n_items=300, k_range=[2], n_trials=6, device=DEV,
                                                use_braiding=True, bind_with_random=False)
    assert res_b["accuracy"][0] > 0.9
    with pytest.raises(ValueError):
        harness.run_binding_unbinding_pairs(d=64, n_items=20, k_range=[2], n_trials=1, device=DEV, unbind_method="x")


@pytest.mark.parametrize("d,method", [(256, "inv"), (200, "inv"), (256, "†")])
def test_self_binding_curves_match_oracle_loops(d, method):
    from clifford_b200 import harness
    from oracle import latent_oracle as O
    torch.manual_seed(5)
    N, T, D = 40, 4, 3
    all_z = O.normalize_vectors(torch.randn(N, d) / d ** 0.5)
    tidx = torch.randint(0, N, (T,))
    pidx = torch.stack([torch.tensor([j for j in torch.randperm(N).tolist() if j != int(tidx[t])][:D]) for t in range(T)])
    s_self, s_rand = harness.self_binding_curves(all_z.to(DEV), tidx.to(DEV), pidx.to(DEV), D, method)
    for t in range(T):                  # utils/wandb_utils.py:96-124
        target, partners = all_z[tidx[t]:tidx[t] + 1], all_z[pidx[t]]
        for m in range(1, D + 1):
            b1 = target.clone()
            b2 = target.clone()
            for i in range(m):
                b1 = O.bind(b1, target)
                b2 = O.bind(b2, partners[i:i + 1])
            for i in range(m - 1, -1, -1):
                b1 = O.unbind(b1, target, method=method)
                b2 = O.unbind(b2, partners[i:i + 1], method=method)
            tol = 5e-5 if method == "inv" else 2e-3      # deconvolution amplifies round-off at small |F_b|
            assert abs(float(s_self[m - 1, t]) - float(O.similarity(b1, target).mean())) < tol
            assert abs(float(s_rand[m - 1, t]) - float(O.similarity(b2, target).mean())) < tol


def test_leading_dim_broadcast_is_not_materialised():
    """(k, T, d) op (1, T, d): the (T, d) operand is indexed modulo T inside the kernel, forward and backward."""
    from utils import vsa
    torch.manual_seed(6)
    k, T, d = 3, 5, 128
    a = torch.randn(k, T, d, device=DEV, requires_grad=True)
    b = torch.randn(1, T, d, device=DEV, requires_grad=True)
    out = vsa.bind(a, b)
    ref = torch.fft.irfft(torch.fft.rfft(a.detach().double()) * torch.fft.rfft(b.detach().double()), n=d)
    assert rel_err(out.detach().cpu(), ref.cpu()) < 1e-5
    w = torch.randn_like(out)
    ga, gb = torch.autograd.grad((out * w).sum(), [a, b])
    a64, b64 = a.detach().double().requires_grad_(), b.detach().double().requires_grad_()
    (torch.fft.irfft(torch.fft.rfft(a64) * torch.fft.rfft(b64), n=d) * w.double()).sum().backward()
    assert gb.shape == b.shape and rel_err(ga.cpu(), a64.grad.cpu()) < 1e-5 and rel_err(gb.cpu(), b64.grad.cpu()) < 1e-5
    s = vsa.similarity(a, b[0])
    assert rel_err(s.detach().cpu(), torch.nn.functional.cosine_similarity(a.detach().double(), b.detach().double(), dim=-1).cpu()) < 1e-5


@pytest.mark.parametrize("d", [16, 100, 512])
def test_angles_to_clifford_vector_and_interpolation(d):
    """utils/wandb_utils.py:506-521 and mnist/mnist_clifpws.py:121-135 restated with torch.fft in fp64."""
    import math
    from clifford_b200 import harness
    torch.manual_seed(d)
    n = 2 * d

    def ref(angles, ortho):
        th = torch.zeros(*angles.shape[:-1], n, dtype=torch.float64)
        th[..., 1:d] = angles[..., 1:]
        th[..., -d + 1:] = -torch.flip(angles[..., 1:], (-1,))
        s = torch.exp(1j * th)
        return torch.fft.ifft(s, dim=-1, norm="ortho" if ortho else None).real

    ang = (torch.rand(3, 7, d, dtype=torch.float64) * 2 - 1) * math.pi
    got = harness.angles_to_clifford_vector(ang.float().to(DEV))
    assert got.shape == (3, 7, n) and rel_err(got.cpu(), ref(ang.float().double(), False)) < 1e-5
    got_o = harness.angles_to_clifford_vector(ang.float().to(DEV), ortho=True)
    assert rel_err(got_o.cpu(), ref(ang.float().double(), True)) < 1e-5
    a1, a2 = ang[0, 0].float(), ang[0, 1].float()
    steps = 9
    zi = harness.clifford_interpolate(a1.to(DEV), a2.to(DEV), steps)
    delta = (a2 - a1 + math.pi) % (2 * math.pi) - math.pi
    ia = a1 + torch.linspace(0, 1, steps).view(-1, 1) * delta
    assert rel_err(zi.cpu(), ref(ia.double(), True)) < 2e-5
    assert float((zi.norm(dim=-1) - n ** 0.5).abs().max()) < 1e-3       # ortho scaling: ||z|| = sqrt(n)
