"""CPU: the half-angle inverse-CDF table the device sampler interpolates (csrc/icdf_table.cuh, built on the host in
double precision by cvb_ps_halfangle_icdf_table -- no GPU needed) against SciPy's inverse regularised incomplete beta:
node values, and the CDF error of the full pipeline (Lagrange in log1p(kappa) + cubic Hermite in s) that the CUDA
kernel evaluates, restated here in numpy."""
import ctypes

import numpy as np
import pytest

sp = pytest.importorskip("scipy.special")


def _table():
    from clifford_b200 import _lib
    lib = _lib.load()
    nk, nn, km = ctypes.c_int(), ctypes.c_int(), ctypes.c_float()
    assert lib.cvb_ps_halfangle_icdf_table(None, 0, ctypes.addressof(nk), ctypes.addressof(nn), ctypes.addressof(km)) == 0
    buf = np.zeros((nk.value, nn.value, 2), dtype=np.float32)
    assert lib.cvb_ps_halfangle_icdf_table(buf.ctypes.data, buf.size, None, None, None) == 0
    return buf, float(km.value)


def _G(psi, k):
    return 1.0 - sp.betainc(k + 0.5, 0.5, np.cos(psi) ** 2)


def test_nodes_match_scipy_inverse_incomplete_beta():
    tab, kmax = _table()
    nk, nn, _ = tab.shape
    assert (nk, nn, kmax) == (64, 257, 32.0)
    s = np.linspace(0, 1, nn)
    for ki in (0, 1, 7, 20, 40, 63):
        k = np.expm1(np.log1p(kmax) * ki / (nk - 1))
        v = s ** (2 * k + 1)
        want = np.arccos(np.sqrt(sp.betaincinv(k + 0.5, 0.5, v)))
        want[0], want[-1] = np.pi / 2, 0.0
        assert np.abs(tab[ki, :, 0] - want).max() < 3e-7, ki
        assert np.all(np.diff(tab[ki, :, 0]) < 0) and np.all(tab[ki, :, 1] < 0)
        # slopes: dH/ds = -p s^(p-1) / g(H), g = cos^{2k} / Z the density of |psi|
        lnZ = np.log(0.5 * np.sqrt(np.pi)) + sp.gammaln(k + 0.5) - sp.gammaln(k + 1.0)
        p = 2 * k + 1
        slope = -p * s[1:-1] ** (p - 1) / np.exp(2 * k * np.log(np.cos(want[1:-1])) - lnZ) / (nn - 1)
        assert np.abs(tab[ki, 1:-1, 1] - slope).max() < 2e-6 * np.abs(slope).max()


@pytest.mark.parametrize("k", [1e-3, 0.03, 0.13, 1.0, 3.7, 10.0, 31.5])
def test_interpolated_sampler_reproduces_the_cdf(k):
    """numpy restatement of icdf_build_row + icdf_sample: max |CDF(sample(v)) - (1 - v)| over a fine set of v."""
    tab, kmax = _table()
    nk, nn, _ = tab.shape
    M = nn - 1
    x = np.log1p(k) * (nk - 1) / np.log1p(kmax)
    i = int(np.clip(np.floor(x), 1, nk - 3))
    u = x - i
    w = [-u * (u - 1) * (u - 2) / 6, (u + 1) * (u - 1) * (u - 2) / 2, -(u + 1) * u * (u - 2) / 2, (u + 1) * u * (u - 1) / 6]
    node = sum(w[a] * tab[i - 1 + a].astype(np.float64) for a in range(4))
    H, D = node[:, 0], node[:, 1]
    v = np.concatenate([np.linspace(2.0 ** -24, 1, 200001), 2.0 ** -np.linspace(0, 24, 4001)])
    s = v ** (1 / (2 * k + 1))
    j = np.minimum((s * M).astype(int), M - 1)
    tau = s * M - j
    dh = H[j + 1] - H[j]
    psi = H[j] + tau * (D[j] + tau * ((3 * dh - 2 * D[j] - D[j + 1]) + tau * (-2 * dh + D[j] + D[j + 1])))
    err = np.abs((1 - _G(psi, k)) - v)
    assert err.max() < (2e-6 if k <= 10 else 8e-5), err.max()


def _lagrange(u):
    w = np.array([-u * (u - 1) * (u - 2) / 6, (u + 1) * (u - 1) * (u - 2) / 2, -(u + 1) * u * (u - 2) / 2, (u + 1) * u * (u - 1) / 6])
    a, b, c, e = u + 1, u, u - 1, u - 2
    dw = np.array([-(c * e + b * e + b * c) / 6, (c * e + a * e + a * c) / 2, -(b * e + a * e + a * b) / 2, (b * c + a * c + a * b) / 6])
    return w, dw


def _hermite(node, s, M):
    H, S = node[:, 0], node[:, 1]
    x = s * M
    j = np.minimum(x.astype(int), M - 1)
    tau = x - j
    d0 = H[j + 1] - H[j]
    c1, c2, c3 = S[j], 3 * d0 - 2 * S[j] - S[j + 1], -2 * d0 + S[j] + S[j + 1]
    return H[j] + tau * (c1 + tau * (c2 + tau * c3)), (c1 + tau * (2 * c2 + 3 * tau * c3)) * M


@pytest.mark.parametrize("k", [0.03, 0.13, 0.5, 1.0, 3.7, 10.0, 20.0, 31.0])
def test_table_derivative_is_the_implicit_reparameterisation_gradient(k):
    """The backward of table-sampled rows (csrc/icdf_table.cuh: icdf_build_row<true>, icdf_phi_and_dkappa, restated here
    in numpy) differentiates the table map: d psi / d kappa at fixed uniform draw v = [Lagrange weights differentiated] +
    (d psi / d s)(d s / d kappa), s = v^(1/p).  Against the analytic implicit gradient -- central differences of SciPy's
    inverse regularised incomplete beta in double -- it holds 1e-4 relative (measured 2.5e-5), which is tighter than ATen's
    piecewise approximation of the same quantity that the reference's backward uses."""
    tab, kmax = _table()
    tab = tab.astype(np.float64)
    nk, nn, _ = tab.shape
    M = nn - 1
    qs = (nk - 1) / np.log1p(kmax)
    x = np.log1p(k) * qs
    i = int(np.clip(np.floor(x), 1, nk - 3))
    w, dw = _lagrange(x - i)
    node = sum(w[a] * tab[i - 1 + a] for a in range(4))
    dnode = sum(dw[a] * tab[i - 1 + a] for a in range(4)) * qs / (1 + k)

    def exact(s, kk):
        return np.arccos(np.sqrt(sp.betaincinv(kk + 0.5, 0.5, s ** (2 * kk + 1))))

    s = np.linspace(0.02, 0.98, 193)
    _, dpsi_ds = _hermite(node, s, M)
    dfix, _ = _hermite(dnode, s, M)
    h = 1e-5 * max(k, 1e-2)
    dfix_exact = (exact(s, k + h) - exact(s, k - h)) / (2 * h)
    assert np.abs(dfix - dfix_exact).max() < 5e-5 * np.abs(dfix_exact).max()
    p = 2 * k + 1
    total = dfix + dpsi_ds * (-2 * s * np.log(s) / p)
    v = s ** p
    total_exact = (exact(v ** (1 / (2 * (k + h) + 1)), k + h) - exact(v ** (1 / (2 * (k - h) + 1)), k - h)) / (2 * h)
    assert np.abs(total - total_exact).max() < 1e-4 * np.abs(total_exact).max()
