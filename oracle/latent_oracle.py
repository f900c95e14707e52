"""CPU oracle for the latent hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The shipped path
(``clifford-vae_b200/``) never imports it and has no CPU fallback.

This is a flat, functional restatement (torch CPU tensors, fp32 unless noted) of the
reference's latent-space algorithms, with the base random variates INJECTED so a CUDA
kernel can be fed the same draws:

  * Clifford-torus power-spherical distribution   /root/reference/dists/clifford.py:281-327
  * uniform torus prior                           /root/reference/dists/clifford.py:215-242
  * D-dim PowerSpherical + HypersphericalUniform  /root/reference/dists/clifford.py:85-212,335-337
  * von Mises-Fisher (Wood rejection, fp64)       /root/reference/vmf/hyperspherical_vae/distributions/von_mises_fisher.py:50-217
  * HRR / VSA ops                                 /root/reference/utils/vsa.py:9-96

Parity pin: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, generated in the build
container by ``oracle/gen_golden.py`` (imports /root/reference, records its RNG draws)
and committed under ``tests/golden/``; ``tests/test_oracle_vs_golden.py`` checks every
function here against them.

Third-party arithmetic the reference leans on (not under /root/reference):
torch 2.11.0 (``torch.fft`` / pocketfft, ``torch._sample_dirichlet``,
``torch._dirichlet_grad`` = ATen/native/Distributions.h ``dirichlet_grad_one``,
``lgamma``/``digamma``) and scipy 1.18.1 (``scipy.special.ive``).  ``dirichlet_grad_one``
and ``log_ive`` are restated below in numpy (published algorithms) and pinned against
those libraries in the tests.
"""
from __future__ import annotations

import math

import numpy as np
import torch

EPS = 1e-7  # dists/clifford.py:17-18 (_get_eps)
LOG2 = math.log(2.0)
LOGPI = math.log(math.pi)
LOG2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------
# implicit-reparameterisation gradient of a Beta draw (torch/distributions/dirichlet.py:16-19)
# --------------------------------------------------------------------------------------
class _BetaDraw(torch.autograd.Function):
    """Attach d t'/d alpha to an injected Beta(alpha, beta) draw t'.

    Beta.rsample is Dirichlet([alpha, beta]).rsample()[..., 0]; its backward is
    grad_alpha = dirichlet_grad_one(t', alpha, alpha+beta) * (1 - t') * grad_t'
    (torch/distributions/dirichlet.py:16-19 with x = (t', 1-t'), grad = (g, 0)).
    beta is a constant on every path of the reference.
    """

    @staticmethod
    def forward(ctx, tprime, alpha, beta):
        ctx.save_for_backward(tprime, alpha, beta)
        return tprime.clone()

    @staticmethod
    def backward(ctx, grad):
        tprime, alpha, beta = ctx.saved_tensors
        tp, al, be = torch.broadcast_tensors(tprime, alpha, beta)
        d = torch._dirichlet_grad(tp.contiguous(), al.contiguous(), (al + be).contiguous())
        ga = d * (1 - tp) * grad
        # reduce broadcast dims back to alpha's shape
        while ga.dim() > alpha.dim():
            ga = ga.sum(0)
        for i, s in enumerate(alpha.shape):
            if s == 1 and ga.shape[i] != 1:
                ga = ga.sum(i, keepdim=True)
        return None, ga, None


def beta_draw(tprime, alpha, beta):
    beta = torch.as_tensor(beta, dtype=alpha.dtype).expand_as(alpha)
    return _BetaDraw.apply(tprime, alpha, beta)


# --------------------------------------------------------------------------------------
# per-circle power-spherical pieces (dim = 2)
# --------------------------------------------------------------------------------------
def ps_log_normalizer(kappa, dim):
    """clifford.py:187-196. alpha = (dim-1)/2 + kappa + eps, beta = (dim-1)/2."""
    s = kappa + EPS
    alpha = (dim - 1) / 2 + s
    beta = (dim - 1) / 2
    return -((alpha + beta) * LOG2 + torch.lgamma(alpha) - torch.lgamma(alpha + beta) + beta * LOGPI)


def ps_entropy(kappa, dim):
    """clifford.py:204-212."""
    s = kappa + EPS
    alpha = (dim - 1) / 2 + s
    beta = (dim - 1) / 2
    return -(ps_log_normalizer(kappa, dim) + s * (LOG2 + torch.digamma(alpha) - torch.digamma(alpha + beta)))


def hermitian_phases_to_vector(theta):
    """clifford.py:301-308: theta (..., d) -> real vector (..., 2d).

    S[0]=1, S[k]=exp(i theta_k) for k=1..d-1, S[d]=1, S[n-k]=conj S[k]; z = Re ifft(S).
    theta[..., 0] is ignored.
    """
    d = theta.shape[-1]
    n = 2 * d
    ts = torch.zeros(theta.shape[:-1] + (n,), dtype=theta.dtype)
    ts[..., 1:d] = theta[..., 1:]
    ts[..., n - d + 1:] = -torch.flip(theta[..., 1:], (-1,))
    return torch.fft.ifft(torch.exp(1j * ts), dim=-1).real


def clifford_ps_theta(loc, kappa, tprime, g):
    """clifford.py:295-300 with the PowerSpherical(dim=2, loc=e1) chain flattened.

    tprime ~ Beta(1/2 + kappa + eps, 1/2) (clifford.py:124-134), g ~ N(0,1) gives the
    sign s = g / (|g| + eps) (clifford.py:100-107 with dim-1 = 1); t = 2 t' - 1;
    y = (t, s * sqrt(max(1 - t^2, eps))) (clifford.py:44-48); the Householder step is the
    identity because loc = e1 (clifford.py:72-76); theta = loc + atan2(y1, y0).
    """
    alpha = 0.5 + (kappa + EPS)
    tp = beta_draw(tprime, alpha.expand_as(tprime) if alpha.dim() == tprime.dim() else alpha, 0.5)
    t = 2.0 * tp - 1.0
    s = g / (g.abs() + EPS)
    y1 = s * torch.sqrt(torch.clamp(1 - t * t, min=EPS))
    return loc + torch.atan2(y1, t)


def clifford_ps_rsample(loc, kappa, tprime, g):
    """z (..., 2d) of CliffordPowerSphericalDistribution.rsample (clifford.py:295-308)."""
    return hermitian_phases_to_vector(clifford_ps_theta(loc, kappa, tprime, g))


def clifford_ps_entropy(kappa_bd):
    """clifford.py:318-322: sum over circles 1..d-1 of the dim-2 PS entropy. kappa_bd (..., d)."""
    return ps_entropy(kappa_bd, 2)[..., 1:].sum(-1)


def clifford_uniform_entropy(d):
    """clifford.py:241-242."""
    return (d - 1) * LOG2PI


def clifford_ps_kl(kappa_bd):
    """clifford.py:325-327."""
    return -clifford_ps_entropy(kappa_bd) + clifford_uniform_entropy(kappa_bd.shape[-1])


def clifford_ps_log_prob(value, loc, kappa_bd):
    """clifford.py:310-316 + :198-202. Sum includes circle 0 (unlike entropy)."""
    d = loc.shape[-1]
    freq = torch.fft.fft(value, dim=-1)[..., :d]
    ang = torch.angle(freq)
    dot = torch.cos(loc) * torch.cos(ang) + torch.sin(loc) * torch.sin(ang)
    dot = torch.clamp(dot, min=-1.0 + EPS, max=1.0 - EPS)
    return (ps_log_normalizer(kappa_bd, 2) + kappa_bd * torch.log1p(dot)).sum(-1)


def clifford_uniform_rsample(u):
    """clifford.py:228-236: u ~ U[0,1) (..., d) -> (..., 2d)."""
    return hermitian_phases_to_vector(u * 2 * math.pi)


def clifford_uniform_log_prob(value, d):
    """clifford.py:238-239."""
    return -torch.ones_like(value[..., 0]) * clifford_uniform_entropy(d)


# closed forms the CUDA backward kernels implement; checked against
# autograd of the functions above in tests/test_oracle_vs_golden.py
def clifford_ps_rsample_backward(loc, kappa, tprime, g, grad_z):
    """Returns (dL/dloc (..., d), dL/dkappa per element (..., d)) for z = clifford_ps_rsample."""
    d = loc.shape[-1]
    n = 2 * d
    with torch.no_grad():
        alpha = (0.5 + (kappa + EPS)).expand_as(tprime).contiguous()
        t = 2.0 * tprime - 1.0
        s = g / (g.abs() + EPS)
        om = 1 - t * t
        y1 = s * torch.sqrt(torch.clamp(om, min=EPS))
        theta = loc + torch.atan2(y1, t)
        G = torch.fft.rfft(grad_z, dim=-1)[..., :d]
        dtheta = -(2.0 / n) * (torch.exp(1j * theta) * G.conj()).imag
        dtheta[..., 0] = 0
        # d atan2(y1, t)/dt with y1 = s sqrt(max(om, eps)); r2 = t^2 + y1^2
        r2 = t * t + y1 * y1
        dy1_dt = torch.where(om > EPS, -s * t / torch.sqrt(torch.clamp(om, min=EPS)), torch.zeros_like(t))
        dphi_dt = (t * dy1_dt - y1) / r2
        dtp_dalpha = torch._dirichlet_grad(tprime.contiguous(), alpha, alpha + 0.5) * (1 - tprime)
        dkappa = dtheta * dphi_dt * 2.0 * dtp_dalpha
    return dtheta, dkappa


def clifford_ps_entropy_backward(kappa_bd):
    """dH/dkappa per element for k>=1: -(kappa+eps) (trigamma(alpha) - trigamma(alpha+1/2))."""
    s = kappa_bd + EPS
    alpha = 0.5 + s
    dh = -s * (torch.polygamma(1, alpha) - torch.polygamma(1, alpha + 0.5))
    dh = dh.clone()
    dh[..., 0] = 0
    return dh


# --------------------------------------------------------------------------------------
# D-dimensional PowerSpherical (clifford.py:162-212) and the uniform sphere prior (:85-121)
# --------------------------------------------------------------------------------------
def householder_e1_to_loc(y, loc, eps):
    """clifford.py:72-76 (eps=1e-7) / von_mises_fisher.py:177-181 (eps=1e-5)."""
    e1 = torch.zeros_like(loc)
    e1[..., 0] = 1
    u = e1 - loc
    u = u / (u.norm(dim=-1, keepdim=True) + eps)
    return y - 2 * (y * u).sum(-1, keepdim=True) * u


def powerspherical_rsample(loc, kappa, tprime, g):
    """loc (..., D) unit rows, kappa (...,), tprime (...,) ~ Beta((D-1)/2+kappa+eps, (D-1)/2),
    g (..., D-1) ~ N(0,1).  clifford.py:152-159, :100-107, :44-48, :72-76."""
    D = loc.shape[-1]
    alpha = (D - 1) / 2 + (kappa + EPS)
    tp = beta_draw(tprime, alpha.expand_as(tprime), (D - 1) / 2)
    t = (2.0 * tp - 1.0).unsqueeze(-1)
    v = g / (g.norm(dim=-1, keepdim=True) + EPS)
    y = torch.cat((t, v * torch.sqrt(torch.clamp(1 - t * t, min=EPS))), -1)
    return householder_e1_to_loc(y, loc, EPS)


def powerspherical_log_prob(value, loc, kappa):
    """clifford.py:198-202."""
    D = loc.shape[-1]
    dot = (loc * value).sum(-1)
    dot = torch.clamp(dot, min=-1.0 + EPS, max=1.0 - EPS)
    return ps_log_normalizer(kappa, D) + kappa * torch.log1p(dot)


def powerspherical_entropy(kappa, D):
    return ps_entropy(kappa, D)


def sphere_uniform_entropy(D):
    """clifford.py:109-121 (dim = D)."""
    return -(math.lgamma(D / 2) - (LOG2 + (D / 2) * LOGPI))


def powerspherical_kl(kappa, D):
    """clifford.py:335-337."""
    return -ps_entropy(kappa, D) + sphere_uniform_entropy(D)


def sphere_uniform_rsample(g):
    """clifford.py:100-107."""
    return g / (g.norm(dim=-1, keepdim=True) + EPS)


# --------------------------------------------------------------------------------------
# von Mises-Fisher (vendored s-vae-pytorch)
# --------------------------------------------------------------------------------------
def vmf_wood_constants(kappa64, m):
    """von_mises_fisher.py:90-114 (all fp64). Returns b, a, d."""
    c = torch.sqrt(4 * kappa64 ** 2 + (m - 1) ** 2)
    b_true = (-2 * kappa64 + c) / (m - 1)
    b_app = (m - 1) / (4 * kappa64)
    s = torch.clamp(kappa64 - 10, 0.0, 1.0)
    b = b_app * s + b_true * (1 - s)
    a = (m - 1 + 2 * kappa64 + c) / 4
    d = (4 * a * b) / (1 + b) - (m - 1) * math.log(m - 1)
    return b, a, d


def vmf_sample_w(kappa, m, e_rounds, u_rounds):
    """Wood rejection loop with injected proposals (von_mises_fisher.py:126-175, k=1).

    kappa (B,1) fp32; e_rounds/u_rounds: lists of (B,1) fp64 tensors, one per round
    (e ~ Beta((m-1)/2,(m-1)/2), u ~ U(1e-20, 1-1e-20)).  A row takes the first round in
    which it is accepted.  Gradient reaches kappa only through b (as in the reference).
    Returns w (B,1) in kappa's dtype.
    """
    k64 = kappa.to(torch.float64)
    b, a, d = vmf_wood_constants(k64, m)
    w = torch.zeros_like(b)
    done = torch.zeros_like(b, dtype=torch.bool)
    for e, u in zip(e_rounds, u_rounds):
        w_ = (1 - (1 + b) * e) / (1 - (1 - b) * e)
        t = (2 * a * b) / (1 - (1 - b) * e)
        acc = ((m - 1.0) * t.log() - t + d) > torch.log(u)
        take = acc & ~done
        w = torch.where(take, w_, w)
        done = done | acc
    assert bool(done.all()), "not enough recorded rejection rounds"
    return w.to(kappa.dtype)


def vmf_sample_w3(kappa, u):
    """von_mises_fisher.py:73-88 (m == 3 closed form); u (B,1) ~ U(0,1)."""
    k64 = kappa.to(torch.float64)
    u64 = u.to(torch.float64)
    w = 1 + torch.stack([torch.log(u64), torch.log(1 - u64) - 2 * k64], dim=0).logsumexp(0) / k64
    return w.to(kappa.dtype)


def vmf_rsample(loc, kappa, w, g):
    """von_mises_fisher.py:50-71: w (B,1) from vmf_sample_w, g (B,D) ~ N(0,1) whose column 0
    is discarded (:59-65)."""
    v = g[..., 1:]
    v = v / v.norm(dim=-1, keepdim=True)
    w_ = torch.sqrt(torch.clamp(1 - w ** 2, 1e-10))
    x = torch.cat((w, w_ * v), -1)
    return householder_e1_to_loc(x, loc, 1e-5).to(loc.dtype)


def log_ive_np(v, z):
    """log of the exponentially scaled modified Bessel function I_v(z) e^{-z}, fp64.

    Restated published algorithms (the reference calls scipy.special.ive on the host,
    ops/ive.py:9-26): ascending series (A&S 9.6.10) when it converges quickly, otherwise the
    uniform (Debye) asymptotic expansion in v (A&S 9.7.7) for v >= 12, otherwise Hankel's
    large-argument expansion (A&S 9.7.1).  Pinned against scipy in the tests.
    """
    z = np.asarray(z, dtype=np.float64)
    out = np.empty_like(z)
    it = np.nditer([z, out], op_flags=[["readonly"], ["writeonly"]])
    for zz, oo in it:
        x = float(zz)
        if x * x <= 80.0 * (v + 1.0) or (v < 12.0 and x <= 30.0):
            # series: sum_k (x/2)^{2k+v} / (k! Gamma(k+v+1))
            q = 0.25 * x * x
            term = 1.0
            ssum = 1.0
            k = 1
            while True:
                term *= q / (k * (k + v))
                ssum += term
                if term < 1e-17 * ssum:
                    break
                k += 1
            oo[...] = v * math.log(0.5 * x) - math.lgamma(v + 1.0) + math.log(ssum) - x
        elif v >= 12.0:
            t2 = x / v
            r = math.sqrt(1.0 + t2 * t2)
            p = 1.0 / r
            eta = r + math.log(t2 / (1.0 + r))
            p2 = p * p
            u1 = p * (3.0 - 5.0 * p2) / 24.0
            u2 = p2 * (81.0 - 462.0 * p2 + 385.0 * p2 * p2) / 1152.0
            u3 = p * p2 * (30375.0 - 369603.0 * p2 + 765765.0 * p2 * p2 - 425425.0 * p2 * p2 * p2) / 414720.0
            u4 = p2 * p2 * (4465125.0 - 94121676.0 * p2 + 349922430.0 * p2 * p2 - 446185740.0 * p2 ** 3
                            + 185910725.0 * p2 ** 4) / 39813120.0
            ser = 1.0 + u1 / v + u2 / v ** 2 + u3 / v ** 3 + u4 / v ** 4
            oo[...] = v * eta - 0.5 * math.log(2.0 * math.pi * v) - 0.5 * math.log(r) + math.log(ser) - x
        else:
            mu = 4.0 * v * v
            term = 1.0
            ssum = 1.0
            for k in range(1, 40):
                nt = -term * (mu - (2 * k - 1) ** 2) / (k * 8.0 * x)
                if abs(nt) > abs(term):
                    break
                term = nt
                ssum += term
                if abs(term) < 1e-17 * abs(ssum):
                    break
            oo[...] = -0.5 * math.log(2.0 * math.pi * x) + math.log(ssum)
    return out


def ive_fraction_approx2(v, z, eps=1e-20):
    """ops/ive.py:63-79 (fp64 tensors in, fp64 out)."""
    def delta(a):
        lamb = v + (a - 1.0) / 2.0
        return (v - 0.5) + lamb / (2 * torch.sqrt((lamb ** 2 + z ** 2).clamp(eps)))
    d0, d2 = delta(0.0), delta(2.0)
    b0 = z / (d0 + torch.sqrt(d0 ** 2 + z ** 2).clamp(eps))
    b2 = z / (d2 + torch.sqrt(d2 ** 2 + z ** 2).clamp(eps))
    return (b0 + b2) / 2.0


class _LogIve(torch.autograd.Function):
    """log(ive(v, z) + 1e-20) with the reference's derivative (ops/ive.py:29-34):
    d ive/dz = ive(v-1, z) - ive(v, z) (v+z)/z."""

    @staticmethod
    def forward(ctx, v, z):
        ctx.v = v
        ctx.save_for_backward(z)
        val = np.exp(log_ive_np(v, z.detach().numpy()))
        return torch.log(torch.from_numpy(val) + 1e-20)

    @staticmethod
    def backward(ctx, grad):
        (z,) = ctx.saved_tensors
        v = ctx.v
        zn = z.detach().numpy()
        iv = np.exp(log_ive_np(v, zn))
        ivm1 = np.exp(log_ive_np(v - 1, zn))
        dive = torch.from_numpy(ivm1 - iv * (v + zn) / zn)
        return None, grad * dive / (torch.from_numpy(iv) + 1e-20)


def vmf_log_normalization(kappa, m):
    """von_mises_fisher.py:200-212; kappa (B,1) -> (B,) in kappa's dtype."""
    k64 = kappa.to(torch.float64)
    log_ive = _LogIve.apply(m / 2 - 1, k64)
    out = -((m / 2 - 1) * torch.log(k64) - (m / 2) * LOG2PI - (k64 + log_ive))
    return out.view(*out.shape[:-1]).to(kappa.dtype)


def vmf_entropy(kappa, m):
    """von_mises_fisher.py:183-191."""
    k64 = kappa.to(torch.float64)
    out = -k64 * ive_fraction_approx2(torch.tensor(m / 2, dtype=torch.float64), k64)
    # fp64 + the fp32-rounded log-normaliser, then cast (von_mises_fisher.py:189-191)
    return (out.view(*out.shape[:-1]) + vmf_log_normalization(kappa, m)).to(kappa.dtype)


def vmf_log_prob(x, loc, kappa):
    """von_mises_fisher.py:193-198."""
    m = loc.shape[-1]
    out = kappa * (loc * x).sum(-1, keepdim=True)
    return out.view(*out.shape[:-1]) - vmf_log_normalization(kappa, m)


def vmf_uniform_entropy(dim):
    """hyperspherical_uniform.py:47-54 with dim = m-1 (log surface area of S^dim)."""
    return LOG2 + ((dim + 1) / 2) * LOGPI - math.lgamma((dim + 1) / 2)


def vmf_kl(kappa, m):
    """von_mises_fisher.py:215-217 with the prior HypersphericalUniform(m-1)."""
    return -vmf_entropy(kappa, m) + vmf_uniform_entropy(m - 1)


# --------------------------------------------------------------------------------------
# VSA / HRR ops (utils/vsa.py:9-96)
# --------------------------------------------------------------------------------------
def hrr_init_from_normal(g):
    """vsa.py:9-12 with the N(0,1) draw injected: g (n,d)."""
    return g / math.sqrt(g.shape[-1])


def unitary_init_from_uniform(a, r, d, eps=1e-3):
    """vsa.py:15-36 with both uniform draws injected: a, r (n, (d-1)//2) ~ U[0,1)."""
    n = a.shape[0]
    out = torch.zeros(n, d, dtype=torch.float32)
    lo, hi = 1, (d + 1) // 2
    for i in range(n):
        sign = torch.sign(r[i] - 0.5)
        phi = sign * math.pi * (eps + a[i] * (1 - 2 * eps))
        fv = torch.zeros(d, dtype=torch.complex64)
        fv[0] = 1.0
        fv[lo:hi] = torch.cos(phi) + 1j * torch.sin(phi)
        fv[d // 2 + 1:] = torch.flip(torch.conj(fv[lo:hi]), dims=(0,))
        if d % 2 == 0:
            fv[d // 2] = 1.0
        out[i] = torch.fft.ifft(fv).real
    return out


def normalize_vectors(x):
    """vsa.py:39-40: x / max(||x||, 1e-12)."""
    return x / x.norm(dim=-1, keepdim=True).clamp_min(1e-12)


def bind(a, b):
    """vsa.py:43-46: circular convolution."""
    return torch.fft.ifft(torch.fft.fft(a, dim=-1) * torch.fft.fft(b, dim=-1), dim=-1).real


def invert(a):
    """vsa.py:49-53: index reversal modulo d."""
    return torch.cat([a[..., :1], torch.flip(a[..., 1:], dims=[-1])], dim=-1)


def unbind(ab, b, method="inv"):
    """vsa.py:56-72."""
    if method in ("inv", "*"):
        return bind(ab, invert(b))
    if method in ("†", "deconv"):
        fa = torch.fft.fft(ab, dim=-1) / (torch.fft.fft(b, dim=-1) + 1e-12)
        return torch.fft.ifft(fa, dim=-1).real
    raise ValueError(f"unsupported unbind method: {method}")


def bundle(vectors, normalize=True):
    """vsa.py:75-79."""
    s = vectors.sum(0)
    return s / math.sqrt(vectors.shape[0]) if normalize else s


def permute_vector(v, perm):
    return v[..., perm]


def unpermute_vector(v, perm):
    return v[..., torch.argsort(perm)]


def similarity(a, b):
    """vsa.py:93-96: cosine with each norm clamped at 1e-8 (F.cosine_similarity)."""
    a, b = torch.broadcast_tensors(a, b)
    na = a.norm(dim=-1).clamp_min(1e-8)
    nb = b.norm(dim=-1).clamp_min(1e-8)
    return (a * b).sum(-1) / (na * nb)


# --------------------------------------------------------------------------------------
# numpy restatement of ATen's dirichlet_grad_one (torch/include/ATen/native/Distributions.h:374-511)
# -- the arithmetic spec for the CUDA device function; pinned against torch._dirichlet_grad.
# --------------------------------------------------------------------------------------
_DG_C = np.array([
    [[[1.003668233, -0.01061107488, -0.0657888334, 0.01201642863],
      [0.6336835991, -0.3557432599, 0.05486251648, -0.001465281033],
      [-0.03276231906, 0.004474107445, 0.002429354597, -0.0001557569013]],
     [[0.221950385, -0.3187676331, 0.01799915743, 0.01074823814],
      [-0.2951249643, 0.06219954479, 0.01535556598, 0.001550077057],
      [0.02155310298, 0.004170831599, 0.001292462449, 6.976601077e-05]],
     [[-0.05980841433, 0.008441916499, 0.01085618172, 0.002319392565],
      [0.02911413504, 0.01400243777, -0.002721828457, 0.000751041181],
      [0.005900514878, -0.001936558688, -9.495446725e-06, 5.385558597e-05]]],
    [[[1, -0.02924021934, -0.04438342661, 0.007285809825],
      [0.6357567472, -0.3473456711, 0.05454656494, -0.002407477521],
      [-0.03301322327, 0.004845219414, 0.00231480583, -0.0002307248149]],
     [[0.5925320577, -0.1757678135, 0.01505928619, 0.000564515273],
      [0.1014815858, -0.06589186703, 0.01272886114, -0.0007316646956],
      [-0.007258481865, 0.001096195486, 0.0003934994223, -4.12701925e-05]],
     [[0.06469649321, -0.0236701437, 0.002902096474, -5.896963079e-05],
      [0.001925008108, -0.002869809258, 0.0008000589141, -6.063713228e-05],
      [-0.0003477407336, 6.959756487e-05, 1.097287507e-05, -1.650964693e-06]]],
])


def _digamma64(x):
    from scipy.special import digamma
    return float(digamma(x))


def dirichlet_grad_one_np(x, alpha, total):
    """Scalar fp64 restatement: -(d/dalpha cdf(x; alpha, beta)) / pdf / (1-x), ATen's piecewise form."""
    x = float(x); alpha = float(alpha); total = float(total)
    beta = total - alpha
    boundary = total * x * (1 - x)
    if x <= 0.5 and boundary < 2.5:                      # Taylor series near x = 0
        factor = _digamma64(alpha) - _digamma64(alpha + beta) - math.log(x)
        numer = 1.0
        series = numer / alpha * (factor + 1 / alpha)
        for i in range(1, 11):
            numer *= (i - beta) * x / i
            den = alpha + i
            series += numer / den * (factor + 1 / den)
        r = x * (1 - x) ** (-beta) * series
        return 0.0 if math.isnan(r) else r
    if x >= 0.5 and boundary < 0.75:                     # Taylor series near x = 1 (roles swapped)
        xx, aa, bb = 1 - x, beta, alpha
        factor = _digamma64(aa + bb) - _digamma64(bb)
        numer, betas, dbetas, series = 1.0, 1.0, 0.0, factor / aa
        for i in range(1, 9):
            numer *= -xx / i
            dbetas = dbetas * (bb - i) + betas
            betas = betas * (bb - i)
            series += numer / (aa + i) * (dbetas + factor * betas)
        r = -((1 - xx) ** (1 - bb)) * series
        r = 0.0 if math.isnan(r) else r
        return -r
    if alpha > 6 and beta > 6:                           # Rice saddle point
        mean = alpha / total
        std = math.sqrt(alpha * beta / (total + 1)) / total
        if mean - 0.1 * std <= x <= mean + 0.1 * std:
            b2 = beta * beta
            poly = 47 * x * b2 * b2 + alpha * (
                (43 + 20 * (16 + 27 * beta) * x) * b2 * beta + alpha * (
                    3 * (59 + 180 * beta - 90 * x) * b2 + alpha * (
                        (453 + 1620 * beta * (1 - x) - 455 * x) * beta + alpha * (
                            8 * (1 - x) * (135 * beta - 11)))))
            pn = (1 + 12 * alpha) * (1 + 12 * beta) / (total * total)
            pd = 12960 * alpha ** 3 * beta * beta * (1 + 12 * total)
            return pn / (1 - x) * poly / pd
        prefactor = -x / math.sqrt(2 * alpha * beta / total)
        stirling = ((1 + 1 / (12 * alpha) + 1 / (288 * alpha * alpha))
                    * (1 + 1 / (12 * beta) + 1 / (288 * beta * beta))
                    / (1 + 1 / (12 * total) + 1 / (288 * total * total)))
        t1n = 2 * alpha * alpha * (x - 1) + alpha * beta * (x - 1) - x * beta * beta
        axbx = alpha * (x - 1) + beta * x
        t1d = math.sqrt(2 * alpha / beta) * total ** 1.5 * axbx * axbx
        term1 = t1n / t1d
        term2 = 0.5 * math.log(alpha / (total * x))
        term3 = math.sqrt(8 * alpha * beta / total) / (beta * x + alpha * (x - 1))
        t4b = beta * math.log(beta / (total * (1 - x))) + alpha * math.log(alpha / (total * x))
        term4 = t4b ** -1.5
        return stirling * prefactor * (term1 + term2 * (term3 + (term4 if x < mean else -term4)))
    u = math.log(x)                                      # rational correction
    a = math.log(alpha) - u
    b = math.log(total) - a
    pu = (1.0, u, u * u)
    pa = (1.0, a, a * a)
    p = q = 0.0
    for i in range(3):
        for j in range(3):
            ua = pu[i] * pa[j]
            c0, c1 = _DG_C[0][i][j], _DG_C[1][i][j]
            p += ua * (c0[0] + b * (c0[1] + b * (c0[2] + b * c0[3])))
            q += ua * (c1[0] + b * (c1[1] + b * (c1[2] + b * c1[3])))
    approx = x * (_digamma64(total) - _digamma64(alpha)) / beta
    return p / q * approx
