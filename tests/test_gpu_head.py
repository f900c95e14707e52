"""GPU parity of the folded concentration head: the `from_head` constructors take the RAW output of
the models' `fc_scale` / `fc_concentration` layer and evaluate kappa = clamp(softplus(raw) + floor, max=kmax)
(reference mnist/mlp_vae.py:69-71, cnn/models.py:96,99) inside the sampling kernels and their backward.

Checked against the unfused composition -- the same torch ops the reference models run, feeding the plain constructors
(which are themselves pinned to the reference's golden vectors) -- on the same injected base draws: sample, KL, and the
gradients with respect to loc and to the RAW tensor, including rows where the clamp is active (gradient exactly 0) and
rows past softplus's linear threshold (raw > 20).  Tolerance: 1e-5 max-norm relative on values, 2e-5 on gradients
(the two paths differ only in the rounding of softplus: torch's CUDA kernel vs log1pf(expf(x)) in ours)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"

# raw head outputs covering: deep negative (kappa ~ floor), around 0, the clamp boundary, clamped, past threshold 20
RAW = [-9.0, -2.5, -0.3, 0.0, 0.7, 3.1, 8.9, 9.5, 9.99, 10.4, 12.0, 19.5, 20.5, 33.0]


def _raw(B, gen):
    base = torch.tensor(RAW, dtype=torch.float32)
    extra = torch.randn(max(B - len(RAW), 0), generator=gen) * 3.0
    return torch.cat([base, extra])[:B].reshape(B, 1).to(DEV)


def _head(raw, floor, kmax):
    return torch.clamp(F.softplus(raw) + floor, max=kmax)


@pytest.mark.parametrize("d,S", [(16, 1), (64, 1), (512, 1), (2048, 1), (20, 1), (256, 3)])
@pytest.mark.parametrize("floor", [0.03, 0.5])
def test_clifford_head_matches_unfused(d, S, floor):
    from dists.clifford import CliffordPowerSphericalDistribution, CliffordTorusUniform
    gen = torch.Generator().manual_seed(100 + d)
    B = 24
    loc0 = (torch.rand(B, d, generator=gen) * 6.283 - 3.14).to(DEV)
    raw0 = _raw(B, gen)
    kap0 = _head(raw0, floor, 10.0)
    rows = S * B
    tp = torch.distributions.Beta(0.5 + kap0.cpu().expand(B, d) + 1e-7, torch.tensor(0.5)).sample((S,)).reshape(rows, d)
    tp = tp.clamp(1e-6, 1 - 1e-6).to(DEV)
    g = torch.randn(rows, d, generator=gen).to(DEV)
    gz = torch.randn(S, B, 2 * d, generator=gen).to(DEV) if S > 1 else torch.randn(B, 2 * d, generator=gen).to(DEV)
    gkl = torch.randn(B, generator=gen).to(DEV)
    sshape = torch.Size([S]) if S > 1 else torch.Size()
    prior = CliffordTorusUniform(d, device=DEV)

    def run(fused):
        loc = loc0.clone().requires_grad_()
        raw = raw0.clone().requires_grad_()
        if fused:
            q = CliffordPowerSphericalDistribution.from_head(loc, raw, floor=floor, max=10.0)
        else:
            q = CliffordPowerSphericalDistribution(loc, _head(raw, floor, 10.0))
        z = q.rsample(sshape, _base_draws=(tp, g))
        kl = torch.distributions.kl.kl_divergence(q, prior)
        dl, dr = torch.autograd.grad((z * gz).sum() + (kl * gkl).sum(), [loc, raw])
        return z.detach(), kl.detach(), dl, dr, q

    z1, kl1, dl1, dr1, q1 = run(True)
    z0, kl0, dl0, dr0, _ = run(False)
    assert rel_err(z1.cpu(), z0.cpu()) < 1e-5
    assert rel_err(kl1.cpu(), kl0.cpu()) < 1e-5
    assert rel_err(dl1.cpu(), dl0.cpu()) < 2e-5
    assert rel_err(dr1.cpu(), dr0.cpu()) < 2e-5
    # clamp active (softplus(raw) + floor > 10): no gradient reaches the raw value, exactly
    clamped = (F.softplus(raw0) + floor > 10.0).reshape(-1)
    assert clamped.any() and float(dr1.reshape(-1)[clamped].abs().max()) == 0.0
    # the lazily materialised attribute is the head's value
    assert rel_err(q1.concentration.detach().cpu(), kap0.expand(B, d).cpu()) < 1e-6


def test_clifford_head_device_rng_backward_consistent():
    """Device-RNG forward + backward from the raw tensor: same draws (same Philox seed/offset) through the fused and the
    unfused constructors give the same sample and the same gradients."""
    from dists.clifford import CliffordPowerSphericalDistribution
    from clifford_b200 import _lib
    gen = torch.Generator().manual_seed(7)
    B, d = 64, 1024
    loc0 = (torch.rand(B, d, generator=gen) * 6.283 - 3.14).to(DEV)
    raw0 = _raw(B, gen)
    gz = torch.randn(B, 2 * d, generator=gen).to(DEV)
    outs = []
    for fused in (True, False):
        torch.manual_seed(1234)
        loc = loc0.clone().requires_grad_()
        raw = raw0.clone().requires_grad_()
        q = (CliffordPowerSphericalDistribution.from_head(loc, raw, floor=0.03, max=10.0) if fused
             else CliffordPowerSphericalDistribution(loc, _head(raw, 0.03, 10.0)))
        z = q.rsample()
        dl, dr = torch.autograd.grad((z * gz).sum() + q.entropy().sum(), [loc, raw])
        outs.append((z.detach(), dl, dr))
    (z1, dl1, dr1), (z0, dl0, dr0) = outs
    assert rel_err(z1.cpu(), z0.cpu()) < 1e-5
    assert rel_err(dl1.cpu(), dl0.cpu()) < 2e-5
    assert rel_err(dr1.cpu(), dr0.cpu()) < 5e-5


@pytest.mark.parametrize("D", [3, 40, 513])
def test_powerspherical_head_matches_unfused(D):
    from dists.clifford import PowerSpherical, HypersphericalUniform
    gen = torch.Generator().manual_seed(D)
    B = 20
    loc0 = F.normalize(torch.randn(B, D, generator=gen), dim=-1).to(DEV)
    raw0 = _raw(B, gen)
    kap0 = _head(raw0, 0.8, 10.0).squeeze(-1)
    a = (D - 1) / 2
    tp = torch.distributions.Beta(a + kap0.cpu() + 1e-7, torch.tensor(float(a))).sample().to(DEV)
    g = torch.randn(B, D - 1, generator=gen).to(DEV)
    gz = torch.randn(B, D, generator=gen).to(DEV)
    gkl = torch.randn(B, generator=gen).to(DEV)
    prior = HypersphericalUniform(D, device=DEV)

    def run(fused):
        loc = loc0.clone().requires_grad_()
        raw = raw0.clone().requires_grad_()
        q = (PowerSpherical.from_head(loc, raw, floor=0.8, max=10.0) if fused
             else PowerSpherical(loc, _head(raw, 0.8, 10.0).squeeze(-1)))
        z = q.rsample(_base_draws=(tp, g))
        kl = torch.distributions.kl.kl_divergence(q, prior)
        dl, dr = torch.autograd.grad((z * gz).sum() + (kl * gkl).sum(), [loc, raw])
        return z.detach(), kl.detach(), dl, dr, q

    z1, kl1, dl1, dr1, q1 = run(True)
    z0, kl0, dl0, dr0, _ = run(False)
    assert rel_err(z1.cpu(), z0.cpu()) < 1e-5
    assert float((kl1 - kl0).abs().max()) < 1e-5 * max(1.0, float(kl0.abs().max()))
    assert rel_err(dl1.cpu(), dl0.cpu()) < 2e-5
    assert rel_err(dr1.cpu(), dr0.cpu()) < 2e-5
    clamped = (F.softplus(raw0) + 0.8 > 10.0).reshape(-1)
    assert clamped.any() and float(dr1.reshape(-1)[clamped].abs().max()) == 0.0
    assert rel_err(q1.scale.detach().cpu(), kap0.cpu()) < 1e-6
    # a method off the fused path (log_prob) works on the materialised concentration
    lp1 = q1.log_prob(z1)
    lp0 = PowerSpherical(loc0, kap0).log_prob(z1)
    assert rel_err(lp1.detach().cpu(), lp0.cpu()) < 1e-5


@pytest.mark.parametrize("D", [3, 41, 513])
def test_vmf_head_matches_unfused(D):
    """Device RNG (same seed -> same Wood draws, as kappa is identical in both paths): sample, KL and gradients."""
    from hyperspherical_vae.distributions import VonMisesFisher
    from hyperspherical_vae.distributions.hyperspherical_uniform import HypersphericalUniform
    from clifford_b200 import _lib
    gen = torch.Generator().manual_seed(D)
    B = 20
    loc0 = F.normalize(torch.randn(B, D, generator=gen), dim=-1).to(DEV)
    raw0 = _raw(B, gen)
    gz = torch.randn(B, D, generator=gen).to(DEV)
    gkl = torch.randn(B, generator=gen).to(DEV)
    prior = HypersphericalUniform(D - 1, device=DEV)
    outs = []
    for fused in (True, False):
        torch.manual_seed(99)
        loc = loc0.clone().requires_grad_()
        raw = raw0.clone().requires_grad_()
        q = (VonMisesFisher.from_head(loc, raw, floor=0.8, max=10.0) if fused
             else VonMisesFisher(loc, _head(raw, 0.8, 10.0)))
        z = q.rsample()
        kl = torch.distributions.kl.kl_divergence(q, prior)
        dl, dr = torch.autograd.grad((z * gz).sum() + (kl * gkl).sum(), [loc, raw])
        outs.append((z.detach(), kl.detach(), dl, dr))
    (z1, kl1, dl1, dr1), (z0, kl0, dl0, dr0) = outs
    # softplus rounding (1 ulp of kappa) can flip a Wood acceptance only with negligible probability; compare directly
    assert rel_err(z1.cpu(), z0.cpu()) < 2e-5
    assert float((kl1 - kl0).abs().max()) < 1e-5 * max(1.0, float(kl0.abs().max()))
    assert rel_err(dl1.cpu(), dl0.cpu()) < 5e-5
    assert rel_err(dr1.cpu(), dr0.cpu()) < 5e-5


def test_head_step_is_one_launch():
    """rsample + KL from the raw head output: exactly one kernel launch of ours (the unfused composition adds three
    elementwise torch kernels for softplus / add / clamp before it)."""
    from dists.clifford import CliffordPowerSphericalDistribution, CliffordTorusUniform, PowerSpherical, HypersphericalUniform
    from clifford_b200 import _lib
    loc = torch.randn(128, 512, device=DEV, requires_grad=True)
    raw = torch.randn(128, 1, device=DEV, requires_grad=True)
    n0 = _lib.launch_count()
    q = CliffordPowerSphericalDistribution.from_head(loc, raw)
    z = q.rsample()
    kl = torch.distributions.kl.kl_divergence(q, CliffordTorusUniform(512, device=DEV))
    assert _lib.launch_count() - n0 == 1
    assert kl.requires_grad and z.requires_grad
    loc2 = F.normalize(torch.randn(64, 513, device=DEV), dim=-1).requires_grad_()
    raw2 = torch.randn(64, 1, device=DEV, requires_grad=True)
    n0 = _lib.launch_count()
    q2 = PowerSpherical.from_head(loc2, raw2)
    z2 = q2.rsample()
    kl2 = torch.distributions.kl.kl_divergence(q2, HypersphericalUniform(513, device=DEV))
    assert _lib.launch_count() - n0 == 1
    (z2.sum() + kl2.sum() + z.sum() + kl.sum()).backward()
    assert raw.grad is not None and raw2.grad is not None and torch.isfinite(raw.grad).all() and torch.isfinite(raw2.grad).all()
