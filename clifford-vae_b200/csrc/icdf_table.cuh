// Inverse-CDF table of the power-spherical half-angle on one circle (row-scalar concentration).
//
// t' ~ Beta(1/2 + k, 1/2) (reference dists/clifford.py:124-134 with dim = 2) is cos^2(psi) for a half-angle
// |psi| in [0, pi/2] with density ~ cos^{2k}(psi).  With v uniform in (0, 1] and p = 2k + 1,
//     |psi| = H_k(s),   s = v^{1/p},   H_k(s) = G_k^{-1}(1 - s^p),   G_k = CDF of |psi|,
// and H_k is smooth on [0, 1] for every k (the substitution s = v^{1/p} removes the (pi/2 - psi)^{2k+1} tail
// singularity of the plain inverse CDF): cubic Hermite interpolation on a uniform s-grid of 256 cells reproduces the
// CDF to < 1e-6 for k <= 10 and < 6e-5 for k <= 32 (tests/test_icdf_table.py, against SciPy's betaincinv).  H_k is
// also smooth in k: the library holds H and dH/ds on 64 concentration nodes uniform in log1p(k) over [0, 32] (built in
// double precision on the host at cvb_init), and a row's table is a 4-point Lagrange interpolation between them.
// One circle then costs one uniform word + log2/exp2 + one table cell + 3 FMAs instead of an envelope rejection test
// with a retry queue.  (The DEVICE copy of the table stores the phase 2H and its slope: phi = 2 psi is what the sampler
// adds to loc; the host copy exported through cvb_ps_halfangle_icdf_table holds H itself.)  Rows with k > 32 (and per-element concentrations) keep the exact rejection sampler (rng.cuh).
#pragma once
#include "common.cuh"

namespace cvb {

constexpr int kIcdfKappaNodes = 64;
constexpr int kIcdfCells = 256;                       // s-cells per concentration node (nodes = cells + 1)
constexpr float kIcdfKappaMax = 32.0f;
constexpr float kIcdfQMax = 3.49650756146648f;        // log1p(32)
constexpr int kIcdfRowStride = kIcdfCells + 2;        // nodes per concentration row, padded to a multiple of 16 bytes
constexpr int kIcdfTableEntries = kIcdfKappaNodes * kIcdfRowStride;     // float2 (H, dH/ds / cells)

// per-device copy of the table (nullptr + error set before cvb_init)
const float2* device_icdf_table();

#ifdef __CUDACC__
// Build one row's cell polynomials psi(tau) = c.x + tau (c.y + tau (c.z + tau c.w)) in shared memory (kIcdfCells
// float4); called by the T threads of a group (t = 0..T-1), followed by the caller's group barrier.
// DERIV: the cells of d phase / d kappa at fixed s instead -- the kappa-derivative of the same interpolant (the Lagrange
// weights differentiated, times d log1p(kappa) / d kappa), i.e. the exact pathwise derivative of what the sampler
// evaluates; against the analytic d/dkappa of G_k^{-1}(1 - s^p) it is accurate to 1e-5 relative (max over s and
// k in [0.03, 31]), tests/test_icdf_table.py.
// UNROLL: iterations kept in flight together.  With T = 32 threads per row (d = 512) a build is four iterations, i.e. four
// serial L2 round trips; ahead of a group's FIRST row (nothing to overlap them with: the latency of a small batch) the
// caller asks for 4, inside the row loop (where the build overlaps the previous row's transform and registers are
// scarce) for 1.
// PAD: also write cell[kIcdfCells] = the constant H(1) (the table must then hold kIcdfCells + 1 cells), so that a sampler
// may index it with floor(x) for x = cells * s up to and including cells (s = 1) without clamping.
template <bool DERIV = false, int UNROLL = 1, bool PAD = false>
__device__ __forceinline__ void icdf_build_row(float4* cell, float kappa, const float2* __restrict__ table, int t, int T) {
  const float x = log1pf(kappa) * ((float)(kIcdfKappaNodes - 1) / kIcdfQMax);
  int i = (int)x;
  i = i < 1 ? 1 : (i > kIcdfKappaNodes - 3 ? kIcdfKappaNodes - 3 : i);
  const float u = x - (float)i;                        // in [-1, 2] at the ends of the node range
  float w0, w1, w2, w3;
  if (DERIV) {
    const float a = u + 1.0f, b = u, c = u - 1.0f, e = u - 2.0f;
    const float dx = ((float)(kIcdfKappaNodes - 1) / kIcdfQMax) / (1.0f + kappa);     // d x / d kappa
    w0 = -(c * e + b * e + b * c) * (1.0f / 6.0f) * dx;
    w1 = (c * e + a * e + a * c) * 0.5f * dx;
    w2 = -(b * e + a * e + a * b) * 0.5f * dx;
    w3 = (b * c + a * c + a * b) * (1.0f / 6.0f) * dx;
  } else {
    w0 = -u * (u - 1.0f) * (u - 2.0f) * (1.0f / 6.0f);
    w1 = (u + 1.0f) * (u - 1.0f) * (u - 2.0f) * 0.5f;
    w2 = -(u + 1.0f) * u * (u - 2.0f) * 0.5f;
    w3 = (u + 1.0f) * u * (u - 1.0f) * (1.0f / 6.0f);
  }
  // thread t builds cells 2t', 2t'+1 from nodes 2t' .. 2t'+2: one 128-bit + one 64-bit load per concentration row
  const float2* r0 = table + (size_t)(i - 1) * kIcdfRowStride;
#pragma unroll UNROLL
  for (int j = 2 * t; j < kIcdfCells; j += 2 * T) {
    float2 n0 = make_float2(0.f, 0.f), n1 = n0, n2 = n0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float w = q == 0 ? w0 : (q == 1 ? w1 : (q == 2 ? w2 : w3));
      const float4 a = __ldg(reinterpret_cast<const float4*>(r0 + (size_t)q * kIcdfRowStride + j));
      const float2 b = __ldg(r0 + (size_t)q * kIcdfRowStride + j + 2);
      n0.x = fmaf(w, a.x, n0.x); n0.y = fmaf(w, a.y, n0.y);
      n1.x = fmaf(w, a.z, n1.x); n1.y = fmaf(w, a.w, n1.y);
      n2.x = fmaf(w, b.x, n2.x); n2.y = fmaf(w, b.y, n2.y);
    }
    const float d0 = n1.x - n0.x, d1 = n2.x - n1.x;
    cell[j] = make_float4(n0.x, n0.y, 3.0f * d0 - 2.0f * n0.y - n1.y, -2.0f * d0 + n0.y + n1.y);
    cell[j + 1] = make_float4(n1.x, n1.y, 3.0f * d1 - 2.0f * n1.y - n2.y, -2.0f * d1 + n1.y + n2.y);
    if (PAD && j + 2 == kIcdfCells) cell[kIcdfCells] = make_float4(n2.x, 0.0f, 0.0f, 0.0f);
  }
}

// Value cells and d/dkappa cells of one row in ONE pass over the node rows (the backward needs both and they read the
// same nodes: half the loads and address arithmetic of two icdf_build_row calls).  Same arithmetic, cell for cell.
__device__ __forceinline__ void icdf_build_row_both(float4* cell, float4* dcell, float kappa, const float2* __restrict__ table,
                                                    int t, int T) {
  const float x = log1pf(kappa) * ((float)(kIcdfKappaNodes - 1) / kIcdfQMax);
  int i = (int)x;
  i = i < 1 ? 1 : (i > kIcdfKappaNodes - 3 ? kIcdfKappaNodes - 3 : i);
  const float u = x - (float)i;
  const float a1 = u + 1.0f, b1 = u, c1 = u - 1.0f, e1 = u - 2.0f;
  const float dx = ((float)(kIcdfKappaNodes - 1) / kIcdfQMax) / (1.0f + kappa);
  const float w[4] = {-u * (u - 1.0f) * (u - 2.0f) * (1.0f / 6.0f), (u + 1.0f) * (u - 1.0f) * (u - 2.0f) * 0.5f,
                      -(u + 1.0f) * u * (u - 2.0f) * 0.5f, (u + 1.0f) * u * (u - 1.0f) * (1.0f / 6.0f)};
  const float g[4] = {-(c1 * e1 + b1 * e1 + b1 * c1) * (1.0f / 6.0f) * dx, (c1 * e1 + a1 * e1 + a1 * c1) * 0.5f * dx,
                      -(b1 * e1 + a1 * e1 + a1 * b1) * 0.5f * dx, (b1 * c1 + a1 * c1 + a1 * b1) * (1.0f / 6.0f) * dx};
  const float2* r0 = table + (size_t)(i - 1) * kIcdfRowStride;
  for (int j = 2 * t; j < kIcdfCells; j += 2 * T) {
    float2 n0 = make_float2(0.f, 0.f), n1 = n0, n2 = n0, m0 = n0, m1 = n0, m2 = n0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(r0 + (size_t)q * kIcdfRowStride + j));
      const float2 b = __ldg(r0 + (size_t)q * kIcdfRowStride + j + 2);
      n0.x = fmaf(w[q], a.x, n0.x); n0.y = fmaf(w[q], a.y, n0.y);
      n1.x = fmaf(w[q], a.z, n1.x); n1.y = fmaf(w[q], a.w, n1.y);
      n2.x = fmaf(w[q], b.x, n2.x); n2.y = fmaf(w[q], b.y, n2.y);
      m0.x = fmaf(g[q], a.x, m0.x); m0.y = fmaf(g[q], a.y, m0.y);
      m1.x = fmaf(g[q], a.z, m1.x); m1.y = fmaf(g[q], a.w, m1.y);
      m2.x = fmaf(g[q], b.x, m2.x); m2.y = fmaf(g[q], b.y, m2.y);
    }
    const float d0 = n1.x - n0.x, d1 = n2.x - n1.x;
    cell[j] = make_float4(n0.x, n0.y, 3.0f * d0 - 2.0f * n0.y - n1.y, -2.0f * d0 + n0.y + n1.y);
    cell[j + 1] = make_float4(n1.x, n1.y, 3.0f * d1 - 2.0f * n1.y - n2.y, -2.0f * d1 + n1.y + n2.y);
    const float f0 = m1.x - m0.x, f1 = m2.x - m1.x;
    dcell[j] = make_float4(m0.x, m0.y, 3.0f * f0 - 2.0f * m0.y - m1.y, -2.0f * f0 + m0.y + m1.y);
    dcell[j + 1] = make_float4(m1.x, m1.y, 3.0f * f1 - 2.0f * m1.y - m2.y, -2.0f * f1 + m1.y + m2.y);
  }
}

__device__ __forceinline__ float fast_exp2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float fast_log2(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// One circle from one 32-bit word: bits 0..22 -> v in (0, 1), midpoints of 2^23 equal bins; returns the phase magnitude |phi| = 2 |psi| in [0, pi]
// (the device table stores 2 H).  inv_p = 1 / (2k + 1).  Bit 31 of the word is the circle's sign draw.
// x_out: the table coordinate x = cells * s, s = v^(1/p) in (0, 1] -- what the training path saves for the backward.
// PADDED (the table has the extra cell of icdf_build_row<.., PAD>): floor(x) from a round-toward-zero add of 2^23 -- the
// cell index sits in the low mantissa bits, no float <-> int conversions (they share the MUFU unit with the four
// transcendentals of a circle) and no clamp.
template <bool PADDED>
__device__ __forceinline__ float icdf_sample_phi_t(const float4* cell, float inv_p, uint32_t w, float& x_out) {
  uint32_t vb;
  asm("lop3.b32 %0, %1, 0x007FFFFF, 0x3F800000, 0xEA;" : "=r"(vb) : "r"(w));      // (w & mask) | one, one instruction
  const float v = __uint_as_float(vb) - 0.99999994f;
  static_assert(kIcdfCells == 256, "the table coordinate below is 2^8 s");
  const float x = fast_exp2(fmaf(fast_log2(v), inv_p, 8.0f));                      // cells * v^(1/p), in (0, 256]
  x_out = x;
  float tau;
  float4 c;
  if (PADDED) {
    const float y = __fadd_rz(x, 8388608.0f);
    tau = x - (y - 8388608.0f);
    c = cell[__float_as_uint(y) & 0x1FFu];
  } else {
    int j = (int)x;
    j = j > kIcdfCells - 1 ? kIcdfCells - 1 : j;
    tau = x - (float)j;
    c = cell[j];
  }
  return fmaf(fmaf(fmaf(c.w, tau, c.z), tau, c.y), tau, c.x);
}
__device__ __forceinline__ float icdf_sample_phi(const float4* cell, float inv_p, uint32_t w, float& x_out) {
  return icdf_sample_phi_t<false>(cell, inv_p, w, x_out);
}
__device__ __forceinline__ float icdf_sample_phi(const float4* cell, float inv_p, uint32_t w) {
  float x_unused;
  return icdf_sample_phi(cell, inv_p, w, x_unused);
}
// Backward of a table-sampled circle from its saved coordinate x (> 0): the same phase magnitude, bit for bit, and its
// pathwise derivative with respect to the concentration at fixed uniform draw v:
//   d|phi|/dkappa = D(s) + (d|phi|/dx) (dx/dkappa),   x = cells v^(1/p)  =>  dx/dkappa = -2 x ln(s) / p,  p = 2 kappa + 1
// with D = d|phi|/dkappa at fixed s from the derivative cells (icdf_build_row<true>).
__device__ __forceinline__ float icdf_phi_and_dkappa(const float4* cell, const float4* dcell, float inv_p, float x, float& dphi_dkappa) {
  int j = (int)x;
  j = j > kIcdfCells - 1 ? kIcdfCells - 1 : j;
  const float tau = x - (float)j;
  const float4 c = cell[j], g = dcell[j];
  const float phi = fmaf(fmaf(fmaf(c.w, tau, c.z), tau, c.y), tau, c.x);
  const float dphi_dx = fmaf(fmaf(3.0f * c.w, tau, 2.0f * c.z), tau, c.y);
  const float dfix = fmaf(fmaf(fmaf(g.w, tau, g.z), tau, g.y), tau, g.x);
  const float ln_s = __logf(x) - 5.545177444479562f;               // ln(x / 256)
  dphi_dkappa = fmaf(dphi_dx, -2.0f * x * ln_s * inv_p, dfix);
  return phi;
}
// t' = cos^2(psi) = (1 + cos phi) / 2 in torch's Beta clamp range, like the exact sampler
__device__ __forceinline__ float icdf_tprime(float phi) {
  return fminf(fmaxf(fmaf(0.5f, __cosf(phi), 0.5f), 1.17549435e-38f), 1.0f - 5.9604645e-8f);
}
#endif

}  // namespace cvb
