"""GPU: distribution of the on-device Clifford phase sampler at the sizes that take the inverse-CDF table path
(d >= 512, one concentration per row <= 32; csrc/icdf_table.cuh) and the exact rejection path (larger concentrations):
one-sample KS of >= 2^20 device draws against the analytic law of the reference's draw -- phi = +-acos(2 t' - 1),
t' ~ Beta(1/2 + k, 1/2) (dists/clifford.py:124-134, :295-301), with its sqrt(eps) phase clamp (:44-48) -- and two-sample
KS against samples drawn from the reference class itself (tests/golden/ks_samples*.npz)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _phase_cdf(phi, k):
    """CDF of the phase on (-pi, pi): 1/2 +- G(|phi| / 2) / 2 with G(psi) = 1 - I_{cos^2 psi}(k + 1/2, 1/2)."""
    from scipy.special import betainc
    g = 1.0 - betainc(k + 0.5, 0.5, np.cos(np.abs(phi) / 2) ** 2)
    return 0.5 + 0.5 * np.sign(phi) * g


def _device_phases(k, rows, d):
    from dists.clifford import CliffordPowerSphericalDistribution
    loc = torch.zeros(rows, d, device=DEV)
    q = CliffordPowerSphericalDistribution(loc, torch.full((rows, 1), k, device=DEV), validate_args=False)
    z = q.rsample()
    return torch.angle(torch.fft.rfft(z.double(), dim=-1)[:, 1:d]).cpu().numpy()      # (rows, d-1) phases, loc = 0


@pytest.mark.parametrize("k", [1e-3, 0.13, 1.0, 10.0, 31.0, 100.0])
def test_one_sample_ks_against_the_analytic_law(k):
    torch.manual_seed(int(k * 1000) + 5)
    d, rows = 512, 2112                                   # 2112 * 511 = 1,079,232 >= 2^20 draws
    th = _device_phases(k, rows, d).reshape(-1)
    n = th.size
    assert n >= 1 << 20
    x = np.sort(th)
    F = _phase_cdf(x, k)
    i = np.arange(1, n + 1)
    D = max(np.max(i / n - F), np.max(F - (i - 1) / n))
    # the reference never returns |phi| < sqrt(eps) (clamped to +-sqrt(eps)): that mass sits on two atoms
    atom = _phase_cdf(np.array([3.16227766e-4]), k)[0] - 0.5
    crit = 1.95 / np.sqrt(n)                              # alpha = 0.001
    assert D < crit + atom, (k, D, crit, atom)
    assert np.abs(th).min() >= 3.0e-4 and np.abs(th).max() <= np.pi
    # signs are fair and independent of the magnitude
    assert abs(np.mean(th > 0) - 0.5) < 4 / np.sqrt(n)
    assert abs(np.corrcoef(np.abs(th[::2][: n // 2 - 1]), np.abs(th[1::2][: n // 2 - 1]))[0, 1]) < 5e-3


@pytest.mark.parametrize("k,fname", [(0.001, "ks_samples_extra.npz"), (0.13, "ks_samples_extra.npz"),
                                     (100.0, "ks_samples_extra.npz"), (0.1, "ks_samples.npz"), (1.0, "ks_samples.npz"),
                                     (10.0, "ks_samples.npz")])
def test_two_sample_ks_against_reference_class_samples(k, fname):
    from scipy.stats import ks_2samp
    ref = np.load(os.path.join(GOLDEN, fname))[f"clifford_phi_k{k}"]
    torch.manual_seed(11)
    th = _device_phases(k, 64, 512)
    for col in (0, 255, 510):                             # single circles across rows ...
        stat, pval = ks_2samp(th[:, col], ref)
        assert pval > 1e-3, (k, col, stat, pval)
    stat, pval = ks_2samp(th.reshape(-1)[:8192], ref)     # ... and across circles of the same rows
    assert pval > 1e-3, (k, stat, pval)


def test_rows_with_different_concentrations_use_their_own_table():
    """Consecutive rows switch between table rows (k <= 32) and exact rows (k > 32) inside one launch."""
    from dists.clifford import CliffordPowerSphericalDistribution
    torch.manual_seed(3)
    ks = torch.tensor([0.05, 40.0, 2.0, 0.5, 64.0, 9.0, 31.9, 32.1], device=DEV).repeat(96)[:, None]
    rows, d = ks.shape[0], 1024
    z = CliffordPowerSphericalDistribution(torch.zeros(rows, d, device=DEV), ks, validate_args=False).rsample()
    th = torch.angle(torch.fft.rfft(z.double(), dim=-1)[:, 1:d]).cpu().numpy()
    for j, k in enumerate(ks[:8, 0].tolist()):
        x = np.sort(th[j::8].reshape(-1))
        n = x.size
        F = _phase_cdf(x, k)
        i = np.arange(1, n + 1)
        D = max(np.max(i / n - F), np.max(F - (i - 1) / n))
        atom = _phase_cdf(np.array([3.16227766e-4]), k)[0] - 0.5
        assert D < 1.95 / np.sqrt(n) + atom, (k, D)


@pytest.mark.parametrize("d", [16, 512, 20])
def test_von_mises_torus_distribution_sampling(d):
    """CliffordTorusDistribution.rsample (dists/clifford.py:261-275): every circle's phase is loc + VonMises(0, kappa);
    one-sample KS against scipy's von Mises CDF, unit-modulus spectrum, per-element and row-scalar concentrations."""
    from scipy.stats import vonmises
    from dists.clifford import CliffordTorusDistribution, CliffordTorusUniform
    torch.manual_seed(d)
    rows = 4096 if d <= 20 else 512
    for k in (0.05, 2.0, 30.0):
        loc = torch.full((rows, d), 0.3, device=DEV)
        q = CliffordTorusDistribution(loc, torch.full((rows, 1), k, device=DEV))
        z = q.rsample()
        assert z.shape == (rows, 2 * d) and q.batch_shape == (rows,) and q.event_shape == (2 * d,)
        F = torch.fft.rfft(z.double(), dim=-1)
        assert float((F.abs() - 1).abs().max()) < 2e-5
        th = (torch.angle(F[:, 1:d]) - 0.3).cpu().numpy().reshape(-1)
        th = (th + np.pi) % (2 * np.pi) - np.pi
        x = np.sort(th)
        n = x.size
        Fx = vonmises.cdf(x, k)
        i = np.arange(1, n + 1)
        D = max(np.max(i / n - Fx), np.max(Fx - (i - 1) / n))
        assert D < 1.95 / np.sqrt(n), (k, D)
    # per-element concentrations, sample_shape, KL registration with the uniform prior
    kap = torch.rand(rows, d, device=DEV) * 5 + 0.1
    q = CliffordTorusDistribution(torch.zeros(rows, d, device=DEV), kap)
    z = q.rsample(torch.Size([2]))
    assert z.shape == (2, rows, 2 * d)
    assert float((torch.fft.rfft(z.double(), dim=-1).abs() - 1).abs().max()) < 2e-5
    kl = torch.distributions.kl.kl_divergence(q, CliffordTorusUniform(d, device=DEV))
    assert kl.shape == (rows,) and bool((kl > 0).all())


@pytest.mark.parametrize("d", [16, 512])
def test_per_element_concentrations_straddling_the_envelope_switch(d):
    """with per-element concentrations on both sides of 1/pi (Gaussian vs uniform envelope) two circles that
    share one Philox result must still get independent proposals: marginals by KS, pairs by correlation."""
    from dists.clifford import CliffordPowerSphericalDistribution
    torch.manual_seed(5)
    rows = 1 << 15 if d == 16 else 2048
    kap = torch.where(torch.arange(d, device=DEV) % 2 == 0, 0.2, 0.5).expand(rows, d).contiguous()
    if d == 512:      # pairs are (k, k + 32) there: alternate in blocks of 32
        kap = torch.where((torch.arange(d, device=DEV) // 32) % 2 == 0, 0.2, 0.5).expand(rows, d).contiguous()
    q = CliffordPowerSphericalDistribution(torch.zeros(rows, d, device=DEV), kap, validate_args=False)
    th = torch.angle(torch.fft.rfft(q.rsample().double(), dim=-1)[:, :d]).cpu().numpy()
    kv = kap[0].cpu().numpy()
    for k in (0.2, 0.5):
        x = np.sort(th[:, (kv == k) & (np.arange(d) > 0)].reshape(-1))
        n = x.size
        F = _phase_cdf(x, float(np.float32(k)))
        i = np.arange(1, n + 1)
        D = max(np.max(i / n - F), np.max(F - (i - 1) / n))
        atom = _phase_cdf(np.array([3.16227766e-4]), k)[0] - 0.5
        assert D < 1.95 / np.sqrt(n) + atom, (k, D)
    step = 1 if d == 16 else 32
    # per column pair (k, k + step) -- the two circles that share a Philox result -- after removing the column means
    # (the two columns have different concentrations, hence different mean |phi|)
    a, b = np.abs(th[:, 2:d - step:1]), np.abs(th[:, 2 + step:d:1])
    a = (a - a.mean(0)) / a.std(0)
    b = (b - b.mean(0)) / b.std(0)
    c = (a * b).mean(0)                                   # one correlation coefficient per pair, sd 1 / sqrt(rows)
    assert abs(c.mean()) < 5 / np.sqrt(a.size) and np.abs(c).max() < 6 / np.sqrt(rows), (c.mean(), np.abs(c).max())
