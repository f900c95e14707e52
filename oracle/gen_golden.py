"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE.  Run once in the build container (the GPU box has no /root/reference):

    python oracle/gen_golden.py

The reference draws its random variates internally, so this script wraps the torch RNG
entry points it uses (Beta.rsample / Beta.sample / Uniform.sample / Normal.sample /
torch.randn / torch.rand) to RECORD the draws; the recorded draws are stored next to the
reference outputs so the oracle and the CUDA kernels can be fed the identical variates.
Nothing from the reference is copied; only its inputs/outputs are stored.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
import zlib

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _import_reference():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "vmf"))
    for m in ("matplotlib", "matplotlib.pyplot"):  # utils/vsa.py:4 imports it; not installed here
        sys.modules.setdefault(m, types.ModuleType(m))
    import dists.clifford as rc
    from hyperspherical_vae.distributions import VonMisesFisher
    from hyperspherical_vae.distributions.hyperspherical_uniform import HypersphericalUniform as VMFUniform
    spec = importlib.util.spec_from_file_location("ref_vsa", os.path.join(REF, "utils", "vsa.py"))
    rv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rv)
    return rc, VonMisesFisher, VMFUniform, rv


class Recorder:
    """Context manager that records every base draw the reference makes."""

    def __init__(self):
        self.log = []

    def __enter__(self):
        D = torch.distributions
        self._saved = (D.Beta.rsample, D.Beta.sample, D.Uniform.sample, D.Normal.sample, torch.randn, torch.rand)
        rec = self.log

        def wrap(fn, tag, is_method=True):
            def inner(*a, **k):
                o = fn(*a, **k)
                rec.append((tag, o.detach().clone()))
                return o
            return inner

        D.Beta.rsample = wrap(self._saved[0], "beta_rsample")
        D.Beta.sample = wrap(self._saved[1], "beta_sample")
        D.Uniform.sample = wrap(self._saved[2], "uniform_sample")
        D.Normal.sample = wrap(self._saved[3], "normal_sample")
        torch.randn = wrap(self._saved[4], "randn")
        torch.rand = wrap(self._saved[5], "rand")
        return self

    def __exit__(self, *exc):
        D = torch.distributions
        D.Beta.rsample, D.Beta.sample, D.Uniform.sample, D.Normal.sample, torch.randn, torch.rand = self._saved

    def get(self, tag):
        return [t for g, t in self.log if g == tag]


def np_(t):
    return t.detach().cpu().numpy()


def gen_clifford(rc):
    cases = {}
    kl_div = torch.distributions.kl.kl_divergence
    specs = [  # name, B, d, kappa_mode, sample_shape
        ("b4_d16_rowk", 4, 16, "row", ()),
        ("b3_d8_fullk", 3, 8, "full", ()),
        ("b2_d512_rowk", 2, 512, "row", ()),
        ("b5_d5_rowk", 5, 5, "row", ()),          # n = 10, not a power of two
        ("b3_d64_rowk_s2", 3, 64, "row", (2,)),    # IWAE-style sample_shape
        ("b6_d2048_rowk", 6, 2048, "row", ()),
        ("b4_d20_rowk", 4, 20, "row", ()),         # n = 40
    ]
    for name, B, d, kmode, sshape in specs:
        torch.manual_seed(zlib.crc32(name.encode()))
        loc = (torch.randn(B, d) * 2.0).requires_grad_()
        if kmode == "row":
            kap = (torch.rand(B, 1) * 9.9 + 0.03).requires_grad_()
        else:
            kap = (torch.rand(B, d) * 9.9 + 0.03).requires_grad_()
        q = rc.CliffordPowerSphericalDistribution(loc, kap)
        p = rc.CliffordTorusUniform(d)
        with Recorder() as r:
            z = q.rsample(torch.Size(sshape))
        (tprime,) = r.get("beta_rsample")
        (g,) = r.get("randn")
        gz = torch.randn_like(z)
        dloc, dkap = torch.autograd.grad((z * gz).sum(), [loc, kap], retain_graph=True)
        ent = q.entropy()
        kl = kl_div(q, p)
        gkl = torch.randn_like(kl)
        (dkap_kl,) = torch.autograd.grad((kl * gkl).sum(), [kap])
        lp_z = q.log_prob(z.detach())
        value = torch.randn(*z.shape) / np.sqrt(2 * d)  # generic vector, not on the torus
        glp = torch.randn_like(lp_z)
        lp_v = q.log_prob(value)
        dloc_lp, dkap_lp = torch.autograd.grad((lp_v * glp).sum(), [loc, kap])
        cases[name] = dict(
            loc=np_(loc), kappa=np_(kap), tprime=np_(tprime), g=np_(g.squeeze(-1)), z=np_(z),
            grad_z=np_(gz), dloc=np_(dloc), dkappa=np_(dkap), entropy=np_(ent), kl=np_(kl),
            grad_kl=np_(gkl), dkappa_kl=np_(dkap_kl), log_prob_z=np_(lp_z), value=np_(value),
            log_prob_value=np_(lp_v), grad_lp=np_(glp), dloc_lp=np_(dloc_lp), dkappa_lp=np_(dkap_lp),
            prior_log_prob=np_(p.log_prob(z.detach())),
        )
    # uniform prior sampler
    for name, S, d in [("uni_s7_d16", 7, 16), ("uni_s3_d512", 3, 512), ("uni_s4_d5", 4, 5)]:
        torch.manual_seed(zlib.crc32(name.encode()))
        p = rc.CliffordTorusUniform(d)
        with Recorder() as r:
            z = p.rsample(torch.Size([S]))
        (u,) = r.get("rand")
        cases[name] = dict(u=np_(u), z=np_(z), entropy=np.float64(p.entropy()))
    flat = {f"{c}/{k}": v for c, d_ in cases.items() for k, v in d_.items()}
    np.savez_compressed(os.path.join(OUT, "clifford.npz"), **flat)
    print("clifford.npz", len(flat), "arrays")


def gen_powerspherical(rc):
    cases = {}
    kl_div = torch.distributions.kl.kl_divergence
    for name, B, D, sshape in [("b6_D5", 6, 5, ()), ("b4_D513", 4, 513, ()), ("b8_D3", 8, 3, ()),
                               ("b5_D512", 5, 512, ()), ("b3_D40_s2", 3, 40, (2,))]:
        torch.manual_seed(zlib.crc32(name.encode()))
        loc_raw = torch.randn(B, D, requires_grad=True)
        loc = torch.nn.functional.normalize(loc_raw, dim=-1)
        kap = (torch.rand(B) * 9.2 + 0.8).requires_grad_()
        q = rc.PowerSpherical(loc, kap)
        p = rc.HypersphericalUniform(D)
        with Recorder() as r:
            z = q.rsample(torch.Size(sshape))
        (tprime,) = r.get("beta_rsample")
        (g,) = r.get("randn")
        gz = torch.randn_like(z)
        dloc, dkap = torch.autograd.grad((z * gz).sum(), [loc, kap], retain_graph=True)
        ent = q.entropy()
        kl = kl_div(q, p)
        gkl = torch.randn_like(kl)
        (dkap_kl,) = torch.autograd.grad((kl * gkl).sum(), [kap])
        value = torch.nn.functional.normalize(torch.randn(*z.shape), dim=-1)
        lp = q.log_prob(value)
        glp = torch.randn_like(lp)
        dloc_lp, dkap_lp = torch.autograd.grad((lp * glp).sum(), [loc, kap])
        cases[name] = dict(
            loc=np_(loc), kappa=np_(kap), tprime=np_(tprime), g=np_(g), z=np_(z), grad_z=np_(gz),
            dloc=np_(dloc), dkappa=np_(dkap), entropy=np_(ent), kl=np_(kl), grad_kl=np_(gkl),
            dkappa_kl=np_(dkap_kl), value=np_(value), log_prob=np_(lp), grad_lp=np_(glp),
            dloc_lp=np_(dloc_lp), dkappa_lp=np_(dkap_lp), prior_entropy=np_(p.entropy()),
            prior_log_prob=np_(p.log_prob(value)),
        )
    torch.manual_seed(5)
    p = rc.HypersphericalUniform(33)
    with Recorder() as r:
        s = p.rsample(torch.Size([9]))
    cases["uniform_D33"] = dict(g=np_(r.get("randn")[0]), z=np_(s))
    flat = {f"{c}/{k}": v for c, d_ in cases.items() for k, v in d_.items()}
    np.savez_compressed(os.path.join(OUT, "powerspherical.npz"), **flat)
    print("powerspherical.npz", len(flat), "arrays")


def gen_vmf(VMF, VMFUniform):
    cases = {}
    kl_div = torch.distributions.kl.kl_divergence
    for name, B, D in [("b6_D5", 6, 5), ("b4_D513", 4, 513), ("b8_D3", 8, 3), ("b16_D41", 16, 41),
                       ("b5_D512", 5, 512)]:
        torch.manual_seed(zlib.crc32(name.encode()))
        loc_raw = torch.randn(B, D)
        loc = torch.nn.functional.normalize(loc_raw, dim=-1).requires_grad_()
        kap = (torch.rand(B, 1) * 9.2 + 0.8)
        if name == "b16_D41":
            kap = kap * 3.0  # exercise kappa > 10 (b_app blend)
        kap.requires_grad_()
        q = VMF(loc, kap)
        p = VMFUniform(D - 1)
        p.device = torch.device("cpu")
        with Recorder() as r:
            z = q.rsample()
        e_rounds = r.get("beta_sample")
        uni = r.get("uniform_sample")
        (g,) = r.get("normal_sample")
        gz = torch.randn_like(z)
        dloc, dkap = torch.autograd.grad((z * gz).sum(), [loc, kap], retain_graph=True, allow_unused=True)
        ent = q.entropy()
        kl = kl_div(q, p)
        gkl = torch.randn_like(kl)
        (dkap_kl,) = torch.autograd.grad((kl * gkl).sum(), [kap])
        value = torch.nn.functional.normalize(torch.randn(B, D), dim=-1)
        lp = q.log_prob(value)
        glp = torch.randn_like(lp)
        dloc_lp, dkap_lp = torch.autograd.grad((lp * glp).sum(), [loc, kap])
        c = dict(
            loc=np_(loc), kappa=np_(kap), g=np_(g), z=np_(z), grad_z=np_(gz), dloc=np_(dloc),
            dkappa=np_(dkap) if dkap is not None else np.zeros((B, 1), np.float32),
            entropy=np_(ent), kl=np_(kl), grad_kl=np_(gkl), dkappa_kl=np_(dkap_kl), value=np_(value),
            log_prob=np_(lp), grad_lp=np_(glp), dloc_lp=np_(dloc_lp), dkappa_lp=np_(dkap_lp),
            prior_entropy=np_(p.entropy()), n_rounds=np.int64(len(e_rounds)),
        )
        if D == 3:
            c["u"] = np_(uni[0])
        else:
            c["e_rounds"] = np.stack([np_(e) for e in e_rounds])
            c["u_rounds"] = np.stack([np_(u) for u in uni])
        cases[name] = c
    flat = {f"{c}/{k}": v for c, d_ in cases.items() for k, v in d_.items()}
    np.savez_compressed(os.path.join(OUT, "vmf.npz"), **flat)
    print("vmf.npz", len(flat), "arrays")


def gen_vsa(rv):
    cases = {}
    for name, k, d in [("k5_d64", 5, 64), ("k3_d1024", 3, 1024), ("k4_d37", 4, 37), ("k2_d513", 2, 513),
                       ("k2_d4096", 2, 4096), ("k1_d16384", 1, 16384), ("k3_d144", 3, 144)]:
        torch.manual_seed(zlib.crc32(name.encode()))
        with Recorder() as r:
            a = rv.hrr_init(k, d)
        g_a = r.get("randn")[0]
        with Recorder() as r:
            b = rv.unitary_init(k, d)
        rands = r.get("rand")
        ua = torch.stack(rands[0::2]) if rands else torch.zeros(k, 0)
        ur = torch.stack(rands[1::2]) if rands else torch.zeros(k, 0)
        c = rv.hrr_init(k, d)
        a.requires_grad_(); c.requires_grad_()
        ab = rv.bind(a, c)
        gab = torch.randn_like(ab)
        da, dc = torch.autograd.grad((ab * gab).sum(), [a, c])
        ub_inv = rv.unbind(ab.detach(), c.detach(), "inv")
        ub_dec = rv.unbind(ab.detach(), c.detach(), "†")
        abu = rv.bind(a.detach(), b)
        ub_unitary = rv.unbind(abu, b, "inv")
        perm = torch.randperm(d)
        cases[name] = dict(
            g_a=np_(g_a), a=np_(a), ua=np_(ua), ur=np_(ur), b_unitary=np_(b), c=np_(c), bind_ac=np_(ab),
            grad_out=np_(gab), da=np_(da), dc=np_(dc), unbind_inv=np_(ub_inv), unbind_deconv=np_(ub_dec),
            bind_a_unitary=np_(abu), unbind_unitary=np_(ub_unitary), invert_a=np_(rv.invert(a.detach())),
            bundle_norm=np_(rv.bundle(a.detach(), True)), bundle_raw=np_(rv.bundle(a.detach(), False)),
            normalize_a=np_(rv.normalize_vectors(a.detach())), sim_ac=np_(rv.similarity(a.detach(), c.detach())),
            sim_bcast=np_(rv.similarity(a.detach()[0], c.detach())), perm=np_(perm),
            permute_a=np_(rv.permute_vector(a.detach(), perm)),
            unpermute_a=np_(rv.unpermute_vector(a.detach(), perm)),
        )
    flat = {f"{c}/{k}": v for c, d_ in cases.items() for k, v in d_.items()}
    np.savez_compressed(os.path.join(OUT, "vsa.npz"), **flat)
    print("vsa.npz", len(flat), "arrays")


def gen_special():
    """Third-party arithmetic pins: torch._dirichlet_grad, lgamma/digamma/trigamma, scipy ive."""
    import scipy.special as sp
    rng = np.random.default_rng(0)
    # (x, alpha, beta) grids covering all four branches of dirichlet_grad_one
    xs, als, bes = [], [], []
    for beta in (0.5, 2.0, 7.5, 19.5, 256.0):
        for alpha in (0.53, 0.9, 1.5, 3.3, 6.5, 10.5, 20.0, 260.0, 266.0):
            x = np.concatenate([rng.beta(alpha, beta, 24), [1e-4, 0.02, 0.5, 0.98, 0.9999]])
            x = np.clip(x, 1e-6, 1 - 1e-6)
            xs.append(x); als.append(np.full_like(x, alpha)); bes.append(np.full_like(x, beta))
    x = np.concatenate(xs).astype(np.float32); al = np.concatenate(als).astype(np.float32)
    be = np.concatenate(bes).astype(np.float32)
    dg = torch._dirichlet_grad(torch.from_numpy(x), torch.from_numpy(al), torch.from_numpy(al + be)).numpy()
    arg = np.concatenate([np.linspace(0.05, 12, 200), np.linspace(12, 600, 100)]).astype(np.float32)
    ta = torch.from_numpy(arg)
    vs = np.array([0.5, 1.0, 1.5, 4.0, 19.5, 255.0, 255.5, 1023.0])
    zs = np.concatenate([np.linspace(0.01, 12, 60), np.linspace(12, 400, 40)])
    ive = np.stack([sp.ive(v, zs) for v in vs])
    np.savez_compressed(
        os.path.join(OUT, "special.npz"), dg_x=x, dg_alpha=al, dg_beta=be, dg=dg, arg=arg,
        lgamma=torch.lgamma(ta).numpy(), digamma=torch.digamma(ta).numpy(),
        trigamma=torch.polygamma(1, ta).numpy(), ive_v=vs, ive_z=zs, ive=ive,
    )
    print("special.npz")


def gen_ks(rc, VMF):
    """Sorted reference samples for two-sample KS tests of the on-device RNG."""
    out = {}
    torch.manual_seed(1234)
    n = 8192
    for kap in (0.1, 1.0, 10.0):
        loc = torch.zeros(n, 2)  # d=2: circle 1 carries the phase
        q = rc.CliffordPowerSphericalDistribution(loc, torch.full((n, 1), kap))
        z = q.rsample()  # (n,4): z = irfft([1, e^{i th}, 1]) -> th from fft bin 1
        th = torch.angle(torch.fft.fft(z, dim=-1)[:, 1])
        out[f"clifford_phi_k{kap}"] = np.sort(np_(th))
    for D, kap in ((513, 5.0), (16, 2.0), (3, 4.0)):
        loc = torch.zeros(n, D); loc[:, 0] = 1
        z = rc.PowerSpherical(loc, torch.full((n,), kap)).rsample()
        out[f"ps_t_D{D}_k{kap}"] = np.sort(np_(z[:, 0]))
        out[f"ps_z1_D{D}_k{kap}"] = np.sort(np_(z[:, 1]))
        z = VMF(loc, torch.full((n, 1), kap)).rsample()
        out[f"vmf_w_D{D}_k{kap}"] = np.sort(np_(z[:, 0]))
        out[f"vmf_z1_D{D}_k{kap}"] = np.sort(np_(z[:, 1]))
    np.savez_compressed(os.path.join(OUT, "ks_samples.npz"), **out)
    print("ks_samples.npz", list(out))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    rc, VMF, VMFUniform, rv = _import_reference()
    gen_clifford(rc)
    gen_powerspherical(rc)
    gen_vmf(VMF, VMFUniform)
    gen_vsa(rv)
    gen_special()
    gen_ks(rc, VMF)
    gen_vae_step()
    gen_conv_vae_step()
    gen_ks_extra(rc)


def gen_vae_step():
    """One training-loss evaluation of the reference's MLP VAE (mnist/mlp_vae.py:19-143) per latent family, with
    seeded weights (torch.manual_seed -> deterministic nn.Linear / xavier init under the same torch version),
    seeded synthetic binarised inputs and RECORDED latent draws: loss terms and parameter gradients."""
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "vmf"))
    from mnist.mlp_vae import MLPVAE, vae_loss
    out = {}
    for dist_name, z_dim in (("clifford", 16), ("powerspherical", 9), ("vmf", 9)):
        torch.manual_seed(20240 + z_dim)
        model = MLPVAE(h_dim=128, z_dim=z_dim, distribution=dist_name)
        x = (torch.rand(8, 1, 28, 28) > torch.rand(8, 1, 28, 28)).float()
        with Recorder() as r:
            res = vae_loss(model, x, beta=1.0, return_dict=True)
        res["total"].backward()
        c = {"x": np_(x), "recon": np_(res["recon"]), "kl": np_(res["kl"]), "total": np_(res["total"]),
             "entropy": np_(res["entropy"])}
        if dist_name == "vmf":
            c["e_rounds"] = np.stack([np_(e) for e in r.get("beta_sample")])
            c["u_rounds"] = np.stack([np_(u) for u in r.get("uniform_sample")])
            c["g"] = np_(r.get("normal_sample")[0])
        else:
            c["tprime"] = np_(r.get("beta_rsample")[0])
            g = r.get("randn")[0]
            c["g"] = np_(g.squeeze(-1) if dist_name == "clifford" else g)
        for name, prm in model.named_parameters():
            c["grad_norm/" + name] = np.float64(prm.grad.norm().item())
        c["grad/fc_scale.weight"] = np_(model.fc_scale.weight.grad)
        c["grad/fc_mean.bias"] = np_(model.fc_mean.bias.grad)
        c["grad/decoder.0.bias"] = np_(model.decoder[0].bias.grad)
        c["param_checksum"] = np.float64(sum(p.double().sum().item() for p in model.parameters()))
        for k, v in c.items():
            out[f"{dist_name}/{k}"] = v
    np.savez_compressed(os.path.join(OUT, "vae_step.npz"), **out)
    print("vae_step.npz", len(out), "arrays")


def gen_ks_extra(rc):
    """Reference-class phase samples at the concentrations the round-2 sampler tests add (0.001, 0.13, 100)."""
    out = {}
    torch.manual_seed(4321)
    n = 8192
    for kap in (0.001, 0.13, 100.0):
        loc = torch.zeros(n, 2)
        q = rc.CliffordPowerSphericalDistribution(loc, torch.full((n, 1), kap))
        z = q.rsample()
        th = torch.angle(torch.fft.fft(z, dim=-1)[:, 1])
        out[f"clifford_phi_k{kap}"] = np.sort(np_(th))
    np.savez_compressed(os.path.join(OUT, "ks_samples_extra.npz"), **out)
    print("ks_samples_extra.npz", len(out), "arrays")


def gen_conv_vae_step():
    """One training-loss evaluation of the reference's conv VAE (cnn/models.py:134-315; the C3 model at a small latent)
    with seeded weights, seeded synthetic inputs and RECORDED latent draws: loss terms and parameter gradients."""
    sys.path.insert(0, REF)
    from cnn.models import VAE
    out = {}
    for dist_name, latent_dim in (("clifford", 64), ("powerspherical", 33)):
        torch.manual_seed(777 + latent_dim)
        model = VAE(latent_dim=latent_dim, in_channels=3, distribution=dist_name, device="cpu", recon_loss_type="l1")
        x = torch.rand(4, 3, 32, 32) * 2 - 1
        with Recorder() as r:
            x_recon, q_z, p_z, mu = model(x)
            losses = model.compute_loss(x, x_recon, q_z, p_z, 1.0)
        losses["total_loss"].backward()
        c = {"x": np_(x), "recon": np_(losses["recon_loss"]), "kl": np_(losses["kld_loss"]),
             "total": np_(losses["total_loss"]), "entropy": np_(losses["entropy"]), "mu": np_(mu),
             "x_recon": np_(x_recon)}
        c["tprime"] = np_(r.get("beta_rsample")[0])
        g = r.get("randn")[0]
        c["g"] = np_(g.squeeze(-1) if dist_name == "clifford" else g)
        for name, prm in model.named_parameters():
            c["grad_norm/" + name] = np.float64(prm.grad.norm().item())
        c["grad/encoder.fc_concentration.weight"] = np_(model.encoder.fc_concentration.weight.grad)
        c["grad/encoder.fc_mu.bias"] = np_(model.encoder.fc_mu.bias.grad)
        c["grad/decoder.fc.bias"] = np_(model.decoder.fc.bias.grad)
        c["param_checksum"] = np.float64(sum(p.double().sum().item() for p in model.parameters()))
        for k, v in c.items():
            out[f"{dist_name}/{k}"] = v
    np.savez_compressed(os.path.join(OUT, "conv_vae_step.npz"), **out)
    print("conv_vae_step.npz", len(out), "arrays")


if __name__ == "__main__":
    if os.environ.get("GEN_ONLY") == "vae_step":      # regenerate just this fixture
        os.makedirs(OUT, exist_ok=True)
        _import_reference()
        gen_vae_step()
    elif os.environ.get("GEN_ONLY") == "ks_extra":
        os.makedirs(OUT, exist_ok=True)
        gen_ks_extra(_import_reference()[0])
    elif os.environ.get("GEN_ONLY") == "conv_vae_step":
        os.makedirs(OUT, exist_ok=True)
        _import_reference()
        gen_conv_vae_step()
    else:
        main()
