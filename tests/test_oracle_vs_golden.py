"""Pin the CPU oracle (oracle/latent_oracle.py) against outputs of the real reference.

tests/golden/*.npz were produced by oracle/gen_golden.py importing /root/reference; the
oracle is fed the recorded base draws and must reproduce the reference's outputs (fp32
round-off only) -- forward values AND autograd gradients.
"""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import latent_oracle as O

T = torch.from_numpy
CLIFF = ["b4_d16_rowk", "b3_d8_fullk", "b2_d512_rowk", "b5_d5_rowk", "b3_d64_rowk_s2", "b6_d2048_rowk",
         "b4_d20_rowk"]
TOL = 2e-6   # same arithmetic, same draws: only re-association level differences are allowed
GTOL = 2e-5


@pytest.mark.parametrize("name", CLIFF)
def test_clifford_forward_and_grads(golden_clifford, name):
    c = golden_clifford[name]
    loc = T(c["loc"]).requires_grad_()
    kap = T(c["kappa"]).requires_grad_()
    z = O.clifford_ps_rsample(loc, kap, T(c["tprime"]), T(c["g"]))
    assert rel_err(z.detach(), c["z"]) < TOL
    dloc, dkap = torch.autograd.grad((z * T(c["grad_z"])).sum(), [loc, kap])
    assert rel_err(dloc, c["dloc"]) < GTOL
    # dkappa is a d-term fp32 sum with cancellation (reference's own round-off ~1e-5 of the summands)
    assert rel_err(dkap, c["dkappa"]) < 3e-4
    kbd = kap.expand_as(loc)
    ent = O.clifford_ps_entropy(kbd)
    kl = O.clifford_ps_kl(kbd)
    assert rel_err(ent.detach(), c["entropy"]) < TOL
    assert rel_err(kl.detach(), c["kl"]) < TOL
    (dk,) = torch.autograd.grad((kl * T(c["grad_kl"])).sum(), [kap])
    assert rel_err(dk, c["dkappa_kl"]) < GTOL
    lp = O.clifford_ps_log_prob(T(c["z"]), loc, kbd)
    assert rel_err(lp.detach(), c["log_prob_z"]) < 1e-5
    lpv = O.clifford_ps_log_prob(T(c["value"]), loc, kbd)
    assert rel_err(lpv.detach(), c["log_prob_value"]) < TOL
    dl, dk2 = torch.autograd.grad((lpv * T(c["grad_lp"])).sum(), [loc, kap])
    assert rel_err(dl, c["dloc_lp"]) < GTOL
    assert rel_err(dk2, c["dkappa_lp"]) < GTOL
    d = loc.shape[-1]
    assert rel_err(O.clifford_uniform_log_prob(T(c["z"]), d), c["prior_log_prob"]) < 1e-7


@pytest.mark.parametrize("name", CLIFF)
def test_clifford_closed_form_backward(golden_clifford, name):
    """The closed forms the CUDA backward kernels implement == reference autograd."""
    c = golden_clifford[name]
    loc, kap = T(c["loc"]), T(c["kappa"])
    dtheta, dk_el = O.clifford_ps_rsample_backward(loc, kap, T(c["tprime"]), T(c["g"]), T(c["grad_z"]))
    while dtheta.dim() > loc.dim():
        dtheta, dk_el = dtheta.sum(0), dk_el.sum(0)
    dk = dk_el.sum(-1, keepdim=True) if kap.shape[-1] == 1 else dk_el
    assert rel_err(dtheta, c["dloc"]) < GTOL
    assert rel_err(dk, c["dkappa"]) < 3e-4
    kbd = kap.expand_as(loc)
    dh = O.clifford_ps_entropy_backward(kbd)
    dkl = -dh * T(c["grad_kl"]).unsqueeze(-1)
    dkl = dkl.sum(-1, keepdim=True) if kap.shape[-1] == 1 else dkl
    assert rel_err(dkl, c["dkappa_kl"]) < GTOL


def test_clifford_identities(golden_clifford):
    """|rfft z| = 1, ||z|| = 1, sum z = 1."""
    for name in CLIFF:
        z = T(golden_clifford[name]["z"])
        F = torch.fft.rfft(z, dim=-1)
        assert torch.allclose(F.abs(), torch.ones_like(F.abs()), atol=2e-5)
        assert torch.allclose(z.norm(dim=-1), torch.ones(z.shape[:-1]), atol=1e-5)
        assert torch.allclose(z.sum(-1), torch.ones(z.shape[:-1]), atol=1e-4)


@pytest.mark.parametrize("name", ["uni_s7_d16", "uni_s3_d512", "uni_s4_d5"])
def test_clifford_uniform(golden_clifford, name):
    c = golden_clifford[name]
    assert rel_err(O.clifford_uniform_rsample(T(c["u"])), c["z"]) < TOL
    assert abs(O.clifford_uniform_entropy(c["u"].shape[-1]) - float(c["entropy"])) < 1e-9


@pytest.mark.parametrize("name", ["b6_D5", "b4_D513", "b8_D3", "b5_D512", "b3_D40_s2"])
def test_powerspherical(golden_ps, name):
    c = golden_ps[name]
    loc = T(c["loc"]).requires_grad_()
    kap = T(c["kappa"]).requires_grad_()
    D = loc.shape[-1]
    z = O.powerspherical_rsample(loc, kap, T(c["tprime"]), T(c["g"]))
    assert rel_err(z.detach(), c["z"]) < TOL
    dloc, dkap = torch.autograd.grad((z * T(c["grad_z"])).sum(), [loc, kap])
    assert rel_err(dloc, c["dloc"]) < GTOL
    assert rel_err(dkap, c["dkappa"]) < 2e-4  # saddle-point dirichlet_grad, cancellation-prone
    assert rel_err(O.powerspherical_entropy(kap, D).detach(), c["entropy"]) < TOL
    kl = O.powerspherical_kl(kap, D)
    assert np.max(np.abs(kl.detach().numpy() - c["kl"])) < 2e-4  # difference of ~1e3-sized fp32 terms
    lp = O.powerspherical_log_prob(T(c["value"]), loc, kap)
    assert rel_err(lp.detach(), c["log_prob"]) < TOL
    dl, dk = torch.autograd.grad((lp * T(c["grad_lp"])).sum(), [loc, kap])
    assert rel_err(dl, c["dloc_lp"]) < GTOL
    assert rel_err(dk, c["dkappa_lp"]) < 1e-3
    assert abs(O.sphere_uniform_entropy(D) - float(np.asarray(c["prior_entropy"]).reshape(-1)[0])) < 1e-3 * max(1, abs(float(np.asarray(c["prior_entropy"]).reshape(-1)[0])))


def test_sphere_uniform(golden_ps):
    c = golden_ps["uniform_D33"]
    assert rel_err(O.sphere_uniform_rsample(T(c["g"])), c["z"]) < TOL


@pytest.mark.parametrize("name", ["b6_D5", "b4_D513", "b8_D3", "b16_D41", "b5_D512"])
def test_vmf(golden_vmf, name):
    c = golden_vmf[name]
    loc = T(c["loc"]).requires_grad_()
    kap = T(c["kappa"]).requires_grad_()
    m = loc.shape[-1]
    if m == 3:
        w = O.vmf_sample_w3(kap, T(c["u"]))
    else:
        w = O.vmf_sample_w(kap, m, list(T(c["e_rounds"])), list(T(c["u_rounds"])))
    z = O.vmf_rsample(loc, kap, w, T(c["g"]))
    assert rel_err(z.detach(), c["z"]) < TOL
    dloc, dkap = torch.autograd.grad((z * T(c["grad_z"])).sum(), [loc, kap])
    assert rel_err(dloc, c["dloc"]) < GTOL
    assert rel_err(dkap, c["dkappa"]) < 1e-4
    ent = O.vmf_entropy(kap, m)
    assert np.max(np.abs(ent.detach().numpy() - c["entropy"])) < 1e-3 * max(1.0, np.max(np.abs(c["entropy"])) * 1e-3)
    kl = O.vmf_kl(kap, m)
    assert np.max(np.abs(kl.detach().numpy() - c["kl"])) < 2e-3
    (dk,) = torch.autograd.grad((kl * T(c["grad_kl"])).sum(), [kap])
    assert rel_err(dk, c["dkappa_kl"]) < 1e-4
    lp = O.vmf_log_prob(T(c["value"]), loc, kap)
    assert np.max(np.abs(lp.detach().numpy() - c["log_prob"])) < 1e-3 * max(1.0, np.max(np.abs(c["log_prob"])) * 1e-3)
    dl, dk2 = torch.autograd.grad((lp * T(c["grad_lp"])).sum(), [loc, kap])
    assert rel_err(dl, c["dloc_lp"]) < GTOL
    assert rel_err(dk2, c["dkappa_lp"]) < 1e-4
    assert abs(O.vmf_uniform_entropy(m - 1) - float(np.asarray(c["prior_entropy"]).reshape(-1)[0])) < 1e-3 * max(1, abs(float(np.asarray(c["prior_entropy"]).reshape(-1)[0])))


VSA = ["k5_d64", "k3_d1024", "k4_d37", "k2_d513", "k2_d4096", "k1_d16384", "k3_d144"]


@pytest.mark.parametrize("name", VSA)
def test_vsa(golden_vsa, name):
    c = golden_vsa[name]
    a, cc, b = T(c["a"]), T(c["c"]), T(c["b_unitary"])
    d = a.shape[-1]
    assert rel_err(O.hrr_init_from_normal(T(c["g_a"])), c["a"]) < 1e-7
    assert rel_err(O.unitary_init_from_uniform(T(c["ua"]), T(c["ur"]), d), c["b_unitary"]) < TOL
    assert rel_err(O.bind(a, cc), c["bind_ac"]) < TOL
    go = T(c["grad_out"])
    assert rel_err(O.bind(go, O.invert(cc)), c["da"]) < GTOL      # grad_a bind(a,c) = bind(g, invert(c))
    assert rel_err(O.bind(go, O.invert(a)), c["dc"]) < GTOL
    assert rel_err(O.unbind(T(c["bind_ac"]), cc, "inv"), c["unbind_inv"]) < TOL
    assert rel_err(O.unbind(T(c["bind_ac"]), cc, "deconv"), c["unbind_deconv"]) < 1e-4
    assert rel_err(O.bind(a, b), c["bind_a_unitary"]) < TOL
    assert rel_err(O.unbind(T(c["bind_a_unitary"]), b, "*"), c["unbind_unitary"]) < TOL
    assert rel_err(c["unbind_unitary"], c["a"]) < 1e-4             # unitary unbind recovers a
    assert rel_err(O.invert(a), c["invert_a"]) == 0
    assert rel_err(O.bundle(a, True), c["bundle_norm"]) < TOL
    assert rel_err(O.bundle(a, False), c["bundle_raw"]) < TOL
    assert rel_err(O.normalize_vectors(a), c["normalize_a"]) < TOL
    assert rel_err(O.similarity(a, cc), c["sim_ac"]) < 1e-5
    assert rel_err(O.similarity(a[0], cc), c["sim_bcast"]) < 1e-5
    perm = T(c["perm"])
    assert rel_err(O.permute_vector(a, perm), c["permute_a"]) == 0
    assert rel_err(O.unpermute_vector(a, perm), c["unpermute_a"]) == 0
    with pytest.raises(ValueError):
        O.unbind(a, cc, "nope")


def test_dirichlet_grad_restatement(golden_special):
    """numpy restatement of ATen dirichlet_grad_one vs torch._dirichlet_grad (fp32 CPU)."""
    s = golden_special
    x, al, be, ref = s["dg_x"], s["dg_alpha"], s["dg_beta"], s["dg"]
    got = np.array([O.dirichlet_grad_one_np(float(xx), float(a), float(a) + float(b)) for xx, a, b in zip(x, al, be)])
    err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-6)
    assert np.max(err) < 5e-4, (np.max(err), np.argmax(err))


def test_log_ive_restatement(golden_special):
    s = golden_special
    for i, v in enumerate(s["ive_v"]):
        got = O.log_ive_np(float(v), s["ive_z"])
        ref = s["ive"][i]
        ok = ref > 1e-300
        if not ok.any():
            continue  # scipy underflows for every z at this order
        assert np.max(np.abs(got[ok] - np.log(ref[ok]))) < 1e-9 * np.maximum(1, np.abs(np.log(ref[ok]))).max() + 1e-9, v
