"""CPU: host logic of the `from_head` constructors (the concentration head folded into the samplers).  The
kernels need a GPU (tests/test_gpu_head.py); what runs here is the lazily materialised concentration -- the reference
models' own formula (mnist/mlp_vae.py:69-71, cnn/models.py:96,99) -- shapes, KL dispatch and the refusal of CPU tensors."""
import pytest
import torch
import torch.nn.functional as F


def test_lazy_concentration_is_the_reference_head_formula():
    from dists.clifford import CliffordPowerSphericalDistribution, CliffordTorusDistribution, PowerSpherical
    from hyperspherical_vae.distributions import VonMisesFisher
    torch.manual_seed(0)
    B, d = 6, 8
    loc = torch.randn(B, d)
    raw = torch.tensor([[-9.0], [0.0], [3.0], [9.5], [12.0], [25.0]], requires_grad=True)
    q = CliffordPowerSphericalDistribution.from_head(loc, raw, floor=0.03, max=10.0)
    assert isinstance(q, CliffordTorusDistribution) and q.batch_shape == (B,) and q.event_shape == (2 * d,) and q.orig_dim == d
    assert "concentration" not in q.__dict__                       # nothing evaluated yet
    want = torch.clamp(F.softplus(raw) + 0.03, max=10.0)
    assert torch.equal(q.concentration, want.expand(B, d))
    (g,) = torch.autograd.grad(q.concentration.sum(), [raw])        # differentiable, clamp-active rows get zero
    assert float(g[4]) == 0.0 and float(g[5]) == 0.0 and float(g[1]) == pytest.approx(0.5 * d)
    mu = F.normalize(torch.randn(B, d + 1), dim=-1)
    ps = PowerSpherical.from_head(mu, raw, floor=0.8, max=10.0)
    assert ps.batch_shape == (B,) and ps.event_shape == (d + 1,)
    assert torch.equal(ps.scale, torch.clamp(F.softplus(raw) + 0.8, max=10.0).squeeze(-1))
    vmf = VonMisesFisher.from_head(mu, raw, floor=0.8, max=10.0)
    assert torch.equal(vmf.scale, torch.clamp(F.softplus(raw) + 0.8, max=10.0)) and vmf.batch_shape == mu.shape
    with pytest.raises(ValueError):
        CliffordPowerSphericalDistribution.from_head(loc, torch.zeros(B, d))       # one raw value per row


def test_head_built_distributions_refuse_cpu_sampling_and_dispatch_kl():
    from clifford_b200._lib import CliffordB200Error
    from dists.clifford import CliffordPowerSphericalDistribution, CliffordTorusUniform
    from torch.distributions.kl import _dispatch_kl
    q = CliffordPowerSphericalDistribution.from_head(torch.zeros(2, 16), torch.zeros(2, 1))
    assert _dispatch_kl(type(q), CliffordTorusUniform) is not NotImplemented
    with pytest.raises(CliffordB200Error):
        q.rsample()                                                  # no CPU fallback
