"""GPU: the reference-facing boundary -- harness entry points under the reference's names (utils/vsa.py:99-630)
against the oracle loops, CPU-argument staging, and launch-state hygiene (an op over an empty tensor or on another
thread between a forward and its backward must not disturb the backward)."""
import threading

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_reference_named_entry_points_run_and_keep_the_result_contract():
    from utils import vsa
    torch.manual_seed(0)
    r = vsa.test_bundle_capacity(d=512, n_items=200, k_range=[2, 40], n_trials=8, device=DEV)
    assert set(r) == {"k", "accuracy", "std"} and r["k"] == [2, 40] and r["accuracy"][0] == 1.0
    assert 0.5 < r["accuracy"][1] <= 1.0
    # the reference's default device is "cpu" and drivers pass CPU item memories: both run on the GPU here
    mem = torch.randn(300, 1024) / 32.0
    r = vsa.test_binding_unbinding_pairs(d=1024, n_items=300, k_range=[2, 6], n_trials=6, item_memory=mem)
    assert r["k"] == [2, 6] and r["accuracy"][0] > 0.9 and len(r["std"]) == 2
    r = vsa.test_binding_unbinding_pairs(d=1024, n_items=300, k_range=[2], n_trials=4, device=DEV, item_memory=mem,
                                         use_braiding=True, bind_with_random=False, unbind_method="†", plot=True)
    assert r["accuracy"][0] >= 0.5          # deconvolution by an HRR role is noisy (small |F_role| bins); 8 queries
    with pytest.raises(ValueError):
        vsa.test_binding_unbinding_pairs(d=64, n_items=20, k_range=[2], n_trials=1, device=DEV, unbind_method="x")
    r = vsa.test_per_class_bundle_capacity_k_items(d=256, n_items=100, n_classes=5, items_per_class=2, device=DEV)
    assert r["avg_similarity_matrix"].shape == (10, 10) and r["n_bundles"] == 10
    assert np.allclose(np.diag(r["avg_similarity_matrix"]), 1.0, atol=1e-5)
    assert (r["std_similarity_matrix"] == 0).all()
    # a class without enough items -> the reference's early-out dict
    lab = torch.zeros(10, dtype=torch.long)
    r = vsa.test_per_class_bundle_capacity_k_items(d=64, n_items=10, n_classes=1, items_per_class=20,
                                                   item_memory=torch.randn(10, 64), labels=lab)
    assert r == {"avg_similarity_matrix": None}


@pytest.mark.parametrize("braid,per_class", [(False, False), (True, False), (True, True)])
def test_per_class_similarity_matches_oracle_loops(braid, per_class):
    from oracle import latent_oracle as O
    from utils import vsa
    torch.manual_seed(7)
    n_items, d, n_classes, ipc = 60, 128, 4, 3
    mem = torch.randn(n_items, d)
    labels = torch.randint(0, n_classes, (n_items,))
    perms = (torch.stack([torch.randperm(d) for _ in range(n_classes)]) if per_class else
             torch.stack([torch.randperm(d) for _ in range(n_items)])) if braid else None
    got = vsa.test_per_class_bundle_capacity_k_items(
        d=d, n_items=n_items, n_classes=n_classes, items_per_class=ipc, n_trials=2, device=DEV, item_memory=mem,
        labels=labels, use_braiding=braid, per_class_braid=per_class,
        _perms=None if perms is None else perms.to(DEV))
    # utils/vsa.py:431-513 restated with the oracle ops
    items = O.normalize_vectors(mem)
    if braid:
        items = torch.stack([O.permute_vector(items[i], perms[int(labels[i])] if per_class else perms[i])
                             for i in range(n_items)])
    sel = []
    for c in torch.unique(labels).tolist()[:n_classes]:
        idx = torch.where(labels == c)[0]
        assert len(idx) >= ipc
        sel += idx[:ipc].tolist()
    B = items[sel]
    ref = np.stack([O.similarity(B[i].unsqueeze(0), B).numpy() for i in range(len(sel))])
    assert got["n_bundles"] == n_classes * ipc and got["items_per_class"] == ipc
    assert np.abs(got["avg_similarity_matrix"] - ref).max() < 1e-5


def test_cpu_arguments_are_staged_to_the_gpu_and_returned_home():
    """utils/vsa.py:266-267,278 and wandb_utils.py:165 hand the ops CPU tensors: computed on the GPU, returned on CPU."""
    from clifford_b200 import _lib
    from oracle import latent_oracle as O
    from utils import vsa
    torch.manual_seed(1)
    a, b = torch.randn(5, 256) / 16, torch.randn(5, 256) / 16
    n0 = _lib.launch_count()
    ab = vsa.bind(a, b)
    assert ab.device.type == "cpu" and _lib.launch_count() > n0
    assert rel_err(ab, O.bind(a, b)) < 1e-5
    assert rel_err(vsa.unbind(ab, b.to(DEV), "†"), O.unbind(ab, b, "†")) < 2e-3
    assert vsa.unbind(ab, b.to(DEV)).device.type == "cpu"
    s = vsa.similarity(a[0], b.to(DEV))
    assert s.device.type == "cpu" and rel_err(s, O.similarity(a[0], b)) < 1e-5
    s = vsa.similarity(a.to(DEV), b)
    assert s.device.type == "cuda"
    h = vsa.hrr_init(7, 64)
    u = vsa.unitary_init(7, 64)
    assert h.device.type == "cpu" and u.device.type == "cpu"
    assert float((torch.fft.rfft(u).abs() - 1).abs().max()) < 1e-4
    assert vsa.normalize_vectors(a).device.type == "cpu" and vsa.bundle(a).device.type == "cpu"
    assert vsa.invert(a).device.type == "cpu"
    perm = torch.randperm(256)
    assert torch.equal(vsa.unpermute_vector(vsa.permute_vector(a, perm), perm), a)
    ag = a.clone().requires_grad_()
    vsa.bind(ag, b).sum().backward()                       # autograd flows through the staging copies
    assert ag.grad is not None and ag.grad.device.type == "cpu"


def test_backward_is_not_disturbed_by_an_intervening_empty_op(golden_clifford):
    """Round-1 defect: launch state lived in module globals set by forward; an op over an empty tensor between a
    forward and its backward made the backward skip its launch and return uninitialised memory."""
    from dists.clifford import CliffordPowerSphericalDistribution
    from oracle import latent_oracle as O
    from utils import vsa
    torch.manual_seed(2)
    B, d = 6, 64
    loc = torch.randn(B, d)
    kap = torch.rand(B, 1) * 4 + 0.2
    tp = torch.distributions.Beta(0.5 + kap + 1e-7, torch.tensor(0.5)).sample((d,)).squeeze(-1).T.contiguous()
    g = torch.randn(B, d)
    w = torch.randn(B, 2 * d)
    loc_r, kap_r = loc.clone().requires_grad_(), kap.clone().requires_grad_()
    (O.clifford_ps_rsample(loc_r, kap_r, tp, g) * w).sum().backward()

    loc_g, kap_g = loc.to(DEV).requires_grad_(), kap.to(DEV).requires_grad_()
    z = CliffordPowerSphericalDistribution(loc_g, kap_g).rsample(_base_draws=(tp.to(DEV), g.to(DEV)))
    a = torch.randn(4, 128, device=DEV, requires_grad=True)
    b = torch.randn(4, 128, device=DEV)
    ab = vsa.bind(a, b)
    # ops over empty tensors between the forwards and the backwards
    e = torch.empty(0, 128, device=DEV)
    assert vsa.bind(e, e).shape == (0, 128)
    assert vsa.similarity(e, e).shape == (0,)
    assert vsa.normalize_vectors(e).shape == (0, 128)
    h = vsa.hrr_init(3, 32, device=DEV)                    # and a generator right after them
    assert float(h.abs().max()) > 0 and torch.isfinite(h).all()
    (z * w.to(DEV)).sum().backward()
    assert rel_err(loc_g.grad.cpu(), loc_r.grad) < 2e-5
    assert rel_err(kap_g.grad.cpu(), kap_r.grad) < 3e-4
    ab.sum().backward()
    ref = torch.fft.irfft(torch.fft.rfft(torch.ones(4, 128, dtype=torch.float64)) *
                          torch.fft.rfft(b.double().cpu()).conj(), n=128)
    assert rel_err(a.grad.cpu(), ref) < 1e-5


def test_ops_from_two_threads_do_not_share_launch_state():
    """Autograd runs backward on its own thread; user threads may interleave ops over empty and non-empty tensors."""
    from utils import vsa
    torch.manual_seed(3)
    d = 256
    a = torch.randn(64, d, device=DEV)
    b = torch.randn(64, d, device=DEV)
    ref = torch.fft.irfft(torch.fft.rfft(a.double()) * torch.fft.rfft(b.double()), n=d).cpu()
    stop = threading.Event()
    errs = []

    def empties():
        e = torch.empty(0, d, device=DEV)
        while not stop.is_set():
            try:
                vsa.bind(e, e)
                vsa.similarity(e, e)
            except Exception as ex:                        # pragma: no cover
                errs.append(ex)
                return

    t = threading.Thread(target=empties)
    t.start()
    try:
        for _ in range(200):
            ar = a.clone().requires_grad_()
            out = vsa.bind(ar, b)
            out.sum().backward()
            assert rel_err(out.detach().cpu(), ref) < 1e-5
            assert float(ar.grad.abs().max()) > 0 and torch.isfinite(ar.grad).all()
    finally:
        stop.set()
        t.join()
    assert not errs


def test_samplers_inside_a_cuda_graph_draw_fresh_streams_on_every_replay():
    """(seed, offset) are host-chosen kernel arguments, so a captured graph used to replay the same
    draws.  With the device-resident launch counter (cvb_set_rng_device_counter, _lib.enable_graph_rng) every replay of
    a captured rsample + KL + bind step yields new samples of the right law."""
    from clifford_b200 import _lib
    from dists.clifford import CliffordPowerSphericalDistribution, CliffordTorusUniform, PowerSpherical
    from utils import vsa
    torch.manual_seed(0)
    B, d = 256, 512
    loc = torch.zeros(B, d, device=DEV)
    kap = torch.full((B, 1), 2.0, device=DEV)
    roles = vsa.unitary_init(B, 2 * d, device=DEV)
    ploc = torch.nn.functional.normalize(torch.randn(B, 33, device=DEV), dim=-1)
    pk = torch.full((B,), 5.0, device=DEV)
    _lib.enable_graph_rng(DEV)
    prior = CliffordTorusUniform(d, device=DEV)

    def step():
        q = CliffordPowerSphericalDistribution(loc, kap, validate_args=False)
        z = q.rsample()
        kl = torch.distributions.kl.kl_divergence(q, prior)
        return z, kl, vsa.bind(z, roles), PowerSpherical(ploc, pk).rsample()

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step()                                          # warm-up outside capture (allocator, cvb_init)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        z, kl, bound, zs = step()
    outs = []
    for _ in range(3):
        g.replay()
        torch.cuda.synchronize()
        outs.append((z.clone(), bound.clone(), zs.clone()))
        assert float((z.norm(dim=-1) - 1).abs().max()) < 2e-5
        assert float((torch.fft.rfft(z.double(), dim=-1).abs() - 1).abs().max()) < 2e-5
        rec = vsa.unbind(bound, roles)
        assert float((rec - z).abs().max()) < 1e-4
        assert float((zs.norm(dim=-1) - 1).abs().max()) < 1e-5
    for i in range(3):
        for j in range(i + 1, 3):
            assert float((outs[i][0] - outs[j][0]).abs().max()) > 1e-2       # different draws on every replay
            assert float((outs[i][2] - outs[j][2]).abs().max()) > 1e-2
    # the law is unchanged: KS of the pooled phases of the three replays
    from test_gpu_sampler_ks import _phase_cdf
    th = torch.cat([torch.angle(torch.fft.rfft(o[0].double(), dim=-1)[:, 1:d]).reshape(-1) for o in outs]).cpu().numpy()
    x = np.sort(th)
    n = x.size
    F = _phase_cdf(x, 2.0)
    i = np.arange(1, n + 1)
    assert max(np.max(i / n - F), np.max(F - (i - 1) / n)) < 1.95 / np.sqrt(n) + 2e-4
    # eager sampling afterwards still works and differs from the graph's draws
    z2 = CliffordPowerSphericalDistribution(loc, kap, validate_args=False).rsample()
    assert float((z2 - outs[-1][0]).abs().max()) > 1e-2


def test_captured_training_step_replays_the_right_draws_in_backward():
    """Inside a captured graph the sphere samplers' backward regenerates the forward's tangent normals from the counter-
    based generator, so it must read the launch-counter value its OWN forward read -- also when other sampling launches
    (which bump the counter) sit between the two.  Check through ||z|| = 1: with consistent draws d(z . z)/d loc = 0
    (J^T z = 0); with any other draws it is O(1).  Two samplers forward, then both backward, captured once, replayed."""
    from clifford_b200 import _lib
    from dists.clifford import PowerSpherical
    from hyperspherical_vae.distributions import VonMisesFisher
    torch.manual_seed(1)
    B, D = 64, 129
    loc1 = torch.nn.functional.normalize(torch.randn(B, D, device=DEV), dim=-1).requires_grad_()
    loc2 = torch.nn.functional.normalize(torch.randn(B, D, device=DEV), dim=-1).requires_grad_()
    k1 = torch.full((B,), 4.0, device=DEV, requires_grad=True)
    k2 = torch.full((B, 1), 6.0, device=DEV, requires_grad=True)
    w = torch.randn(B, D, device=DEV)
    _lib.enable_graph_rng(DEV)

    def step():
        z1 = PowerSpherical(loc1, k1).rsample()
        z2 = VonMisesFisher(loc2, k2, validate_args=False).rsample()       # (argument validation syncs: not capturable)
        z3 = PowerSpherical(loc1, k1).rsample()                    # one more bump before any backward runs
        g_unit = torch.autograd.grad((z1 * z1.detach()).sum() + (z2 * z2.detach()).sum() + (z3 * z3.detach()).sum(),
                                     [loc1, loc2], retain_graph=True)
        g_w = torch.autograd.grad((z1 * w).sum() + (z2 * w).sum(), [loc1, loc2])
        return z1, z2, z3, g_unit, g_w

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        z1, z2, z3, g_unit, g_w = step()
    prev = None
    for _ in range(3):
        g.replay()
        torch.cuda.synchronize()
        for z in (z1, z2, z3):
            assert float((z.norm(dim=-1) - 1).abs().max()) < 2e-5
        assert float(g_unit[0].abs().max()) < 5e-5 and float(g_unit[1].abs().max()) < 5e-5     # the draws were replayed
        assert float(g_w[0].abs().max()) > 1e-2 and float(g_w[1].abs().max()) > 1e-2            # and the test is not vacuous
        assert float((z1 - z3).abs().max()) > 1e-2                                               # distinct launches differ
        if prev is not None:
            assert float((z1 - prev).abs().max()) > 1e-2                                         # fresh draws per replay
        prev = z1.clone()
    # eager mode afterwards: same identity (host offsets identify the launch there)
    z = PowerSpherical(loc1, k1).rsample()
    (gl,) = torch.autograd.grad((z * z.detach()).sum(), [loc1])
    assert float(gl.abs().max()) < 5e-5
