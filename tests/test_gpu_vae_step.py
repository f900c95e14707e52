"""GPU, model level: one loss evaluation + backward of the reference's MNIST MLP VAE
(mnist/mlp_vae.py:19-143) reproduced with the drop-in distributions and the reference's recorded
latent draws.  The fixture (tests/golden/vae_step.npz, oracle/gen_golden.py:gen_vae_step) was produced by
the real reference on CPU with seeded weights; the module below restates its layer layout so that the
same seed yields the same weights (checked through a parameter checksum)."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


class _MirrorMLPVAE(nn.Module):
    """784-256-128 encoder, mean / scale heads, (2)z-128-256-784 decoder, xavier weights, zero biases --
    registration order as in the reference so that seeded construction is bit-identical."""

    def __init__(self, z_dim, distribution):
        super().__init__()
        self.z_dim, self.distribution = z_dim, distribution
        self.encoder = nn.Sequential(nn.Linear(784, 256), nn.ReLU(), nn.Linear(256, 128), nn.ReLU())
        self.fc_mean = nn.Linear(128, z_dim)
        self.fc_scale = nn.Linear(128, 1)
        dec_in = 2 * z_dim if distribution == "clifford" else z_dim
        self.decoder = nn.Sequential(nn.Linear(dec_in, 128), nn.ReLU(), nn.Linear(128, 256), nn.ReLU(), nn.Linear(256, 784))

        def init(m):
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)

        self.apply(init)


def _case(dist_name):
    z = np.load(os.path.join(GOLDEN, "vae_step.npz"))
    return {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(dist_name + "/")}


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("dist_name,z_dim", [("clifford", 16), ("powerspherical", 9), ("vmf", 9)])
def test_mlp_vae_loss_and_gradients_match_reference(dist_name, z_dim):
    from dists.clifford import (CliffordPowerSphericalDistribution, CliffordTorusUniform, HypersphericalUniform,
                                PowerSpherical)
    c = _case(dist_name)
    torch.manual_seed(20240 + z_dim)
    model = _MirrorMLPVAE(z_dim, dist_name)
    assert abs(sum(p.double().sum().item() for p in model.parameters()) - float(c["param_checksum"])) < 1e-9
    model = model.to(DEV)
    x = T(c["x"])
    h = model.encoder(x.view(-1, 784))
    if dist_name == "clifford":
        loc = model.fc_mean(h)
        scale = torch.clamp(F.softplus(model.fc_scale(h)) + 0.03, max=10.0)
        q = CliffordPowerSphericalDistribution(loc, scale)
        p = CliffordTorusUniform(z_dim, device=DEV, validate_args=False)
        z = q.rsample(_base_draws=(T(c["tprime"]), T(c["g"])))
    else:
        loc = F.normalize(model.fc_mean(h), p=2, dim=-1)
        scale = torch.clamp(F.softplus(model.fc_scale(h)) + 0.8, max=10.0)
        if dist_name == "powerspherical":
            q = PowerSpherical(loc, scale.squeeze(-1))
            p = HypersphericalUniform(z_dim, device=DEV, validate_args=False)
            z = q.rsample(_base_draws=(T(c["tprime"]), T(c["g"])))
        else:
            from hyperspherical_vae.distributions import VonMisesFisher
            from hyperspherical_vae.distributions.hyperspherical_uniform import HypersphericalUniform as VU
            q = VonMisesFisher(loc, scale)
            p = VU(z_dim - 1, device=DEV, validate_args=False)
            R = c["e_rounds"].shape[0]
            z = q.rsample(_base_draws=(T(c["e_rounds"]).reshape(R, -1), T(c["u_rounds"]).reshape(R, -1), T(c["g"])))
    recon = F.binary_cross_entropy_with_logits(model.decoder(z), x.view(-1, 784), reduction="sum") / x.size(0)
    kl = torch.distributions.kl.kl_divergence(q, p).mean()
    ent = q.entropy().mean()
    total = recon + kl
    total.backward()
    assert abs(float(recon) - float(c["recon"])) < 2e-5 * abs(float(c["recon"]))
    assert abs(float(kl) - float(c["kl"])) < 2e-5 * max(1.0, abs(float(c["kl"])), abs(float(c["entropy"])))
    assert abs(float(ent) - float(c["entropy"])) < 2e-5 * max(1.0, abs(float(c["entropy"])))
    assert abs(float(total) - float(c["total"])) < 2e-5 * abs(float(c["total"]))
    for name, prm in model.named_parameters():
        ref = float(c["grad_norm/" + name])
        assert abs(float(prm.grad.norm()) - ref) < 2e-4 * max(ref, 1e-3), (name, float(prm.grad.norm()), ref)
    assert rel_err(model.fc_scale.weight.grad.cpu(), c["grad/fc_scale.weight"]) < 5e-4
    assert rel_err(model.fc_mean.bias.grad.cpu(), c["grad/fc_mean.bias"]) < 1e-4
    assert rel_err(model.decoder[0].bias.grad.cpu(), c["grad/decoder.0.bias"]) < 1e-4
