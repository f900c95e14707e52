"""GPU parity: D-dim PowerSpherical and von Mises-Fisher (ctypes -> C ABI) vs golden vectors from the
reference, plus KS tests of the on-device samplers against reference samples."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("name", ["b6_D5", "b4_D513", "b8_D3", "b5_D512", "b3_D40_s2"])
def test_powerspherical_matches_reference(golden_ps, name):
    from dists.clifford import PowerSpherical, HypersphericalUniform
    c = golden_ps[name]
    loc = T(c["loc"]).requires_grad_()
    kap = T(c["kappa"]).requires_grad_()
    D = loc.shape[-1]
    q = PowerSpherical(loc, kap)
    p = HypersphericalUniform(D, device=DEV)
    sshape = c["z"].shape[:-2]
    z = q.rsample(torch.Size(sshape), _base_draws=(T(c["tprime"]), T(c["g"])))
    assert z.shape == c["z"].shape
    assert rel_err(z.detach().cpu(), c["z"]) < 1e-5
    dloc, dkap = torch.autograd.grad((z * T(c["grad_z"])).sum(), [loc, kap])
    assert rel_err(dloc.cpu(), c["dloc"]) < 2e-5
    assert rel_err(dkap.cpu(), c["dkappa"]) < 2e-4     # saddle-point Beta gradient, cancellation-prone
    ent = q.entropy()
    kl = torch.distributions.kl.kl_divergence(q, p)
    assert rel_err(ent.detach().cpu(), c["entropy"]) < 1e-5
    assert np.max(np.abs(kl.detach().cpu().numpy() - c["kl"])) < 1e-5 * max(1.0, abs(float(c["prior_entropy"].reshape(-1)[0])))
    (dk,) = torch.autograd.grad((kl * T(c["grad_kl"])).sum(), [kap])
    assert rel_err(dk.cpu(), c["dkappa_kl"]) < 2e-4
    lp = q.log_prob(T(c["value"]))
    assert rel_err(lp.detach().cpu(), c["log_prob"]) < 1e-5
    dl, dk2 = torch.autograd.grad((lp * T(c["grad_lp"])).sum(), [loc, kap])
    assert rel_err(dl.cpu(), c["dloc_lp"]) < 2e-5
    assert rel_err(dk2.cpu(), c["dkappa_lp"]) < 1e-3   # the reference's fp32 lgamma/digamma differences
    assert rel_err(p.entropy().cpu(), c["prior_entropy"]) < 1e-6
    assert rel_err(p.log_prob(T(c["value"])).cpu(), c["prior_log_prob"]) < 1e-6


def test_sphere_uniform(golden_ps):
    from clifford_b200 import ops
    c = golden_ps["uniform_D33"]
    z = ops.sphere_uniform_rsample(9, 33, DEV, 1e-7, gnoise=T(c["g"]))
    assert rel_err(z.cpu(), c["z"]) < 1e-6
    from dists.clifford import HypersphericalUniform
    s = HypersphericalUniform(65, device=DEV).rsample(torch.Size([4096]))
    assert float((s.norm(dim=-1) - 1).abs().max()) < 1e-5
    assert float(s.mean(0).abs().max()) < 0.02


@pytest.mark.parametrize("name", ["b6_D5", "b4_D513", "b8_D3", "b16_D41", "b5_D512"])
def test_vmf_matches_reference(golden_vmf, name):
    from hyperspherical_vae.distributions import VonMisesFisher
    from hyperspherical_vae.distributions.hyperspherical_uniform import HypersphericalUniform as VMFUniform
    c = golden_vmf[name]
    loc = T(c["loc"]).requires_grad_()
    kap = T(c["kappa"]).requires_grad_()
    m = loc.shape[-1]
    q = VonMisesFisher(loc, kap)
    p = VMFUniform(m - 1, device=DEV)
    if m == 3:
        draws = (None, T(c["u"]).reshape(1, -1), T(c["g"]))
    else:
        R = c["e_rounds"].shape[0]
        draws = (T(c["e_rounds"]).reshape(R, -1), T(c["u_rounds"]).reshape(R, -1), T(c["g"]))
    z = q.rsample(_base_draws=draws)
    assert rel_err(z.detach().cpu(), c["z"]) < 1e-5
    dloc, dkap = torch.autograd.grad((z * T(c["grad_z"])).sum(), [loc, kap])
    assert rel_err(dloc.cpu(), c["dloc"]) < 2e-5
    assert rel_err(dkap.cpu(), c["dkappa"]) < 1e-4
    ent = q.entropy()
    assert ent.shape == c["entropy"].shape
    assert np.max(np.abs(ent.detach().cpu().numpy() - c["entropy"])) < 1e-5 * max(1.0, np.max(np.abs(c["entropy"])))
    kl = torch.distributions.kl.kl_divergence(q, p)
    assert np.max(np.abs(kl.detach().cpu().numpy() - c["kl"])) < 1e-5 * max(1.0, np.max(np.abs(c["entropy"])))
    (dk,) = torch.autograd.grad((kl * T(c["grad_kl"])).sum(), [kap])
    assert rel_err(dk.cpu(), c["dkappa_kl"]) < 1e-4
    lp = q.log_prob(T(c["value"]))
    assert np.max(np.abs(lp.detach().cpu().numpy() - c["log_prob"])) < 1e-5 * max(1.0, np.max(np.abs(c["log_prob"])))
    dl, dk2 = torch.autograd.grad((lp * T(c["grad_lp"])).sum(), [loc, kap])
    assert rel_err(dl.cpu(), c["dloc_lp"]) < 2e-5
    assert rel_err(dk2.cpu(), c["dkappa_lp"]) < 1e-4
    assert rel_err(p.entropy().cpu(), c["prior_entropy"]) < 1e-6


def test_ks_device_samplers_vs_reference(golden_ks):
    from scipy.stats import ks_2samp
    from dists.clifford import PowerSpherical
    from hyperspherical_vae.distributions import VonMisesFisher
    torch.manual_seed(3)
    n = 8192
    for D, kap in ((513, 5.0), (16, 2.0), (3, 4.0)):
        loc = torch.zeros(n, D, device=DEV)
        loc[:, 0] = 1
        z = PowerSpherical(loc, torch.full((n,), kap, device=DEV)).rsample().cpu().numpy()
        assert np.max(np.abs(np.linalg.norm(z, axis=-1) - 1)) < 5e-5
        for col, key in ((0, "ps_t"), (1, "ps_z1")):
            pval = ks_2samp(z[:, col], golden_ks[f"{key}_D{D}_k{kap}"]).pvalue
            assert pval > 1e-3, ("ps", D, col, pval)
        z = VonMisesFisher(loc, torch.full((n, 1), kap, device=DEV)).rsample().cpu().numpy()
        assert np.max(np.abs(np.linalg.norm(z, axis=-1) - 1)) < 5e-5
        for col, key in ((0, "vmf_w"), (1, "vmf_z1")):
            pval = ks_2samp(z[:, col], golden_ks[f"{key}_D{D}_k{kap}"]).pvalue
            assert pval > 1e-3, ("vmf", D, col, pval)


@pytest.mark.parametrize("B,D", [(64, 513), (40, 5), (33, 128), (35, 129), (70, 300), (20, 1024), (9, 1025), (6, 2051)])
@pytest.mark.parametrize("family", ["ps", "vmf"])
def test_rng_mode_backward_consistency(family, B, D):
    """RNG mode: the backward replays the tangent normals from Philox; check it against autograd of
    the oracle fed the realised sample's own decomposition.  D <= 1024 runs the register-resident kernels (every
    K = ceil(D / 128) boundary), larger D the two-pass kernels; B > 32 exercises the 32-row scalar batches."""
    from oracle import latent_oracle as O
    from dists.clifford import PowerSpherical
    from hyperspherical_vae.distributions import VonMisesFisher
    torch.manual_seed(9)
    loc = torch.nn.functional.normalize(torch.randn(B, D, device=DEV), dim=-1).requires_grad_()
    if family == "ps":
        kap = (torch.rand(B, device=DEV) * 9 + 0.8).requires_grad_()
        z = PowerSpherical(loc, kap).rsample()
    else:
        kap = (torch.rand(B, 1, device=DEV) * 9 + 0.8).requires_grad_()
        z = VonMisesFisher(loc, kap).rsample()
    gz = torch.randn_like(z)
    dloc, dkap = torch.autograd.grad((z * gz).sum(), [loc, kap])
    # undo the Householder reflection on the CPU to recover y = [t, sqrt(1-t^2) v]
    eps = 1e-7 if family == "ps" else 1e-5
    lo = loc.detach().cpu().double()
    e1 = torch.zeros_like(lo); e1[:, 0] = 1
    u = e1 - lo
    u = u / (u.norm(dim=-1, keepdim=True) + eps)
    y = z.detach().cpu().double()
    y = y - 2 * (y * u).sum(-1, keepdim=True) * u
    t = y[:, 0]
    g = y[:, 1:].float()              # any positive multiple of v gives the same sample
    loc_c = loc.detach().cpu().requires_grad_()
    kap_c = kap.detach().cpu().requires_grad_()
    if family == "ps":
        tprime = ((t + 1) / 2).float()
        zo = O.powerspherical_rsample(loc_c, kap_c, tprime, g)
    else:
        # dw/dkappa needs the accepted proposal e: invert w(e, b) with the oracle's constants
        b, a, dd = O.vmf_wood_constants(kap_c.detach().double(), D)
        w = t.reshape(-1, 1)
        e = (1 - w) / ((1 + b) - w * (1 - b))
        wo = O.vmf_sample_w(kap_c, D, [e], [torch.full_like(e, 1e-300)])
        zo = O.vmf_rsample(loc_c, kap_c, wo, torch.cat([torch.zeros(B, 1), g], -1))
    assert rel_err(zo.detach(), z.detach().cpu()) < 1e-4
    dlo, dko = torch.autograd.grad((zo * gz.cpu()).sum(), [loc_c, kap_c])
    assert rel_err(dloc.cpu(), dlo) < 1e-4
    assert rel_err(dkap.cpu(), dko) < 2e-3


@pytest.mark.parametrize("family", ["powerspherical", "vmf"])
def test_fused_row_scalars_one_launch_and_separate_backward_passes(family, golden_ps, golden_vmf):
    """rsample produces entropy (and the vMF log-normaliser) in the SAME launch; they are separate autograd nodes, so the
    sample and the KL can be differentiated in two backward passes (like the reference's independent graphs) or in one."""
    from clifford_b200 import _lib
    from dists.clifford import PowerSpherical, HypersphericalUniform
    from hyperspherical_vae.distributions import VonMisesFisher
    from hyperspherical_vae.distributions.hyperspherical_uniform import HypersphericalUniform as VMFUniform
    torch.manual_seed(0)
    B, D = 64, 33
    loc = torch.nn.functional.normalize(torch.randn(B, D, device=DEV), dim=-1).requires_grad_()
    if family == "powerspherical":
        kap = (torch.rand(B, device=DEV) * 8 + 0.5).requires_grad_()
        make = lambda: (PowerSpherical(loc, kap), HypersphericalUniform(D, device=DEV))
    else:
        kap = (torch.rand(B, 1, device=DEV) * 8 + 0.5).requires_grad_()
        make = lambda: (VonMisesFisher(loc, kap), VMFUniform(D - 1, device=DEV))
    q, p = make()
    n0 = _lib.launch_count()
    z = q.rsample()
    kl = torch.distributions.kl.kl_divergence(q, p)
    ent = q.entropy()
    assert _lib.launch_count() - n0 == 1                        # sampler + entropy + KL: one kernel
    w = torch.randn_like(z)
    (g1,) = torch.autograd.grad((z * w).sum(), [kap], retain_graph=False)
    (g2,) = torch.autograd.grad(kl.sum(), [kap])               # second, independent backward pass
    # against the unfused evaluation of the same quantities
    q2, p2 = make()
    kl2 = torch.distributions.kl.kl_divergence(q2, p2)         # no rsample before: the stand-alone entropy kernel
    (g2_ref,) = torch.autograd.grad(kl2.sum(), [kap])
    assert rel_err(kl.detach().cpu(), kl2.detach().cpu()) < 1e-6 and rel_err(g2.cpu(), g2_ref.cpu()) < 1e-6
    assert rel_err(ent.detach().cpu(), q2.entropy().detach().cpu()) < 1e-6
    assert torch.isfinite(g1).all() and float(g1.abs().max()) > 0
    # one joint backward equals the sum
    q3, p3 = make()
    torch.manual_seed(1)
    z3 = q3.rsample()
    loss = (z3 * w).sum() + torch.distributions.kl.kl_divergence(q3, p3).sum()
    (g3,) = torch.autograd.grad(loss, [kap])
    torch.manual_seed(1)
    q4, _ = make()
    (g4,) = torch.autograd.grad((q4.rsample() * w).sum(), [kap])
    assert rel_err(g3.cpu(), (g4 + g2_ref).cpu()) < 1e-5


@pytest.mark.parametrize("D", [41, 130, 202, 256, 513, 1024, 2050])
def test_vmf_row_scalars_order_and_concentration_sweep(D):
    """Entropy / log-normaliser and their kappa-derivatives over four decades of kappa at small and large orders against
    the float64 restatement of the reference (von_mises_fisher.py:183-212, ops/ive.py:29-34, SciPy's ive).  Orders
    v = D/2 - 1 >= 100 run through the four-term Debye expansion for every kappa; where ive(v, kappa) sits far below the
    reference's 1e-20 regulariser (small kappa at large D) the kernel skips the Bessel evaluation altogether; large kappa at
    the same D does not underflow and crosses that switch inside one launch."""
    from clifford_b200 import ops
    from oracle import latent_oracle as O
    kap = torch.logspace(-2, 3.5, 97, dtype=torch.float32).reshape(-1, 1)
    k_ref = kap.clone().to(torch.float64).requires_grad_()
    ent_ref = O.vmf_entropy(k_ref, D)
    ln_ref = O.vmf_log_normalization(k_ref, D)
    (dent_ref,) = torch.autograd.grad(ent_ref.sum(), k_ref, retain_graph=True)
    (dln_ref,) = torch.autograd.grad(ln_ref.sum(), k_ref)
    k_dev = kap.to(DEV).requires_grad_()
    ent, ln = ops.VMFEntropyLogNorm.apply(k_dev, D)
    (dent,) = torch.autograd.grad(ent.sum(), k_dev, retain_graph=True)
    (dln,) = torch.autograd.grad(ln.sum(), k_dev)
    scale = max(1.0, float(ln_ref.detach().abs().max()))
    assert float((ln.detach().cpu().double() - ln_ref.detach()).abs().max()) < 1e-5 * scale
    assert float((ent.detach().cpu().double() - ent_ref.detach()).abs().max()) < 1e-5 * scale
    # element-wise: derivative error relative to max(1, |value|) per element
    for got, ref in ((dln, dln_ref), (dent, dent_ref)):
        err = (got.cpu().double().reshape(-1) - ref.reshape(-1)).abs() / ref.reshape(-1).abs().clamp_min(1.0)
        assert float(err.max()) < 2e-5, float(err.max())
