// Row kernels for the D-dimensional samplers: PowerSpherical (reference dists/clifford.py:85-212),
// von Mises-Fisher (reference vmf/hyperspherical_vae/distributions/von_mises_fisher.py:50-217) and
// the uniform sphere prior.  Both samplers share one structure:
//     y = [t, sqrt(max(1 - t^2, clamp_eps)) * g / (||g|| + norm_eps)],   z = y - 2 (y.u) u,
//     u = (e1 - loc) / (||e1 - loc|| + house_eps)
// and differ only in how the scalar t is drawn (Beta marginal vs Wood rejection) -- so one warp per
// row makes one pass for the three reductions and one pass for the Householder reflection
// (8D + 8 algorithmic bytes per row).
#pragma once
#include "common.cuh"
#include "rng.cuh"
#include "special.cuh"

namespace cvb {

enum SphereFamily : int { kFamilyPS = 0, kFamilyVMF = 1, kFamilyUniform = 2 };

struct SphereParams {
  const float* loc;          // (loc_rows, D)
  const float* kappa;        // (loc_rows)
  long long loc_rows;
  const float* tprime;       // PS injected (rows)
  const double* e_rounds;    // vMF injected proposals (n_rounds, rows)
  const double* u_rounds;    // vMF injected uniforms  (n_rounds, rows); D == 3: row 0 is the closed-form uniform
  int n_rounds;
  const float* gnoise;       // injected normals; tangent component i (1..D-1) at gnoise[row*g_pitch + g_off + i - 1]
  int g_pitch, g_off;
  float* z;                  // fwd: (rows, D) out
  float* save;               // (rows, 2): PS (t', 0) ; vMF (w, dw/dkappa).  fwd: out, bwd: in
  const float* grad_z;       // bwd: (rows, D)
  float* dloc;               // bwd: (rows, D) out
  float* dkappa;             // bwd: (rows) out
  long long rows;
  int D;
  float norm_eps, clamp_eps, house_eps;
  // fused row scalars of the first sample of every parameter row (row < loc_rows), all optional:
  float* entropy;            // (loc_rows)
  float* kl;                 // (loc_rows): prior_entropy - entropy
  float* dentropy;           // (loc_rows): d entropy / d kappa
  float* log_norm;           // (loc_rows)  vMF only
  float* dlog_norm;          // (loc_rows)  vMF only
  double prior_entropy;
  KappaHead head;            // on: `kappa` holds the raw head output; every kappa-derivative written is then d / d raw
  PhiloxKey key;
};

// generic float Gamma(alpha) draw (Marsaglia-Tsang with the alpha < 1 boost), independent stream per elem
__device__ __forceinline__ float gamma_draw_float(float alpha, const PhiloxKey& key, uint64_t elem) {
  const GammaMT g(alpha);
  uint32_t attempt = 0;
  float boost = 1.0f, x = 0.f;
  bool first = true;
  for (;;) {
    const uint4 r = philox_draw(key, elem, attempt++);
    const float2 nn = box_muller(r.x, r.y);
    if (first && g.inv_alpha > 0.f) boost = __powf(u01_open0(r.w), g.inv_alpha);
    first = false;
    if (gamma_mt_attempt(g, nn.x, u01_open0(r.z), x)) break;
  }
  return x * boost;
}

// Four N(0,1) draws for element group q of `row` (elements 4q .. 4q+3): ONE Philox call + two Box-Muller
// transforms per four elements.  `first` is the tangent index of element 0 of the row (1 for the samplers,
// whose element 0 is the scalar coordinate; 1 for the uniform sphere too, which has no such slot but uses
// the same injected layout gnoise[row * g_pitch + g_off + tangent_index - 1]).
__device__ __forceinline__ void normals4(const SphereParams& p, long long row, int q, int elem_to_tangent, float (&g)[4]) {
  if (p.gnoise) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ti = 4 * q + j + elem_to_tangent;          // tangent index (>= 1 is a real draw)
      g[j] = (ti >= 1 && ti - 1 < p.g_pitch - p.g_off) ? p.gnoise[row * p.g_pitch + p.g_off + ti - 1] : 0.0f;
    }
    return;
  }
  PhiloxKey k = p.key;
  k.stream = 7;
  const uint4 r = philox_draw(k, (uint64_t)(row * (long long)((p.D + 3) / 4) + q), 0);
  const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
  g[0] = a.x; g[1] = a.y; g[2] = b.x; g[3] = b.y;
}

struct WoodDraw { float w, dw_dkappa; };

// Wood (1994) rejection sampler for the vMF mixture coordinate w, fp64 like the reference (:90-175);
// m == 3 uses the closed form (:73-88).  Executed redundantly by every lane of the row's warp.
__device__ __forceinline__ WoodDraw vmf_draw_w(const SphereParams& p, long long row, float kappa_f) {
  const double kap = (double)kappa_f;
  const int m = p.D;
  WoodDraw o;
  if (m == 3) {
    double u;
    if (p.u_rounds) u = p.u_rounds[row];
    else { PhiloxKey k = p.key; k.stream = 5; const uint4 r = philox_draw(k, (uint64_t)row, 0); u = (double)u01_open1(r.x); u = fmin(fmax(u, 1e-12), 1.0 - 1e-12); }
    const double a = log(u), b = log(1.0 - u) - 2.0 * kap;
    const double mx = fmax(a, b);
    const double L = mx + log(exp(a - mx) + exp(b - mx));
    const double sb = exp(b - L);
    o.w = (float)(1.0 + L / kap);
    o.dw_dkappa = (float)(-2.0 * sb / kap - L / (kap * kap));
    return o;
  }
  const double m1 = (double)(m - 1);
  const double c = sqrt(4.0 * kap * kap + m1 * m1);
  const double b_true = (-2.0 * kap + c) / m1;
  const double b_app = m1 / (4.0 * kap);
  const double s = fmin(fmax(kap - 10.0, 0.0), 1.0);
  const double b = b_app * s + b_true * (1.0 - s);
  const double a = (m1 + 2.0 * kap + c) / 4.0;
  const double dd = (4.0 * a * b) / (1.0 + b) - m1 * log(m1);
  const double ds = (kap > 10.0 && kap < 11.0) ? 1.0 : 0.0;
  const double db = (-m1 / (4.0 * kap * kap)) * s + b_app * ds + ((-2.0 + 4.0 * kap / c) / m1) * (1.0 - s) - b_true * ds;
  double e = 0.5;
  bool ok = false;
  if (p.e_rounds) {
    for (int r = 0; r < p.n_rounds && !ok; ++r) {
      e = p.e_rounds[(long long)r * p.rows + row];
      const double u = p.u_rounds[(long long)r * p.rows + row];
      const double t = (2.0 * a * b) / (1.0 - (1.0 - b) * e);
      ok = (m1 * log(t) - t + dd) > log(u);
    }
  } else {
    PhiloxKey k = p.key;
    k.stream = 5;
    const float h = 0.5f * (float)(m - 1);
    for (uint32_t round = 0; !ok; ++round) {
      const float x = gamma_draw_float(h, k, (uint64_t)(row * 64 + 2 * round));
      const float y = gamma_draw_float(h, k, (uint64_t)(row * 64 + 2 * round + 1));
      e = (double)x / ((double)x + (double)y);
      PhiloxKey ku = p.key;
      ku.stream = 6;
      const uint4 r = philox_draw(ku, (uint64_t)row, round);
      const double u = fmin(fmax(u01_double(r.x, r.y), 1e-20), 1.0 - 1e-20);
      const double t = (2.0 * a * b) / (1.0 - (1.0 - b) * e);
      ok = (m1 * log(t) - t + dd) > log(u);
      if (round > 1000) ok = true;   // cannot happen for finite kappa; guards a hang
    }
  }
  const double den = 1.0 - (1.0 - b) * e;
  o.w = (float)((1.0 - (1.0 + b) * e) / den);
  o.dw_dkappa = (float)(-2.0 * e * (1.0 - e) / (den * den) * db);
  return o;
}

// One warp per row.  FAMILY kFamilyUniform: z = g / (||g|| + norm_eps) with D tangent components.
// vMF entropy (:183-191) / log-normaliser (:200-212) and kappa-derivatives, one thread per row, fp64.
// Reproduces the reference's log(ive + 1e-20) (ive underflows to 0 for large orders) and its Bessel-
// ratio bound ive_fraction_approx2 (ops/ive.py:63-79).
__device__ __forceinline__ void bessel_ratio_bound(double v, double z, double a, double& B, double& dB) {
  const double lam = v + (a - 1.0) / 2.0;
  const double r = sqrt(fmax(lam * lam + z * z, 1e-20));
  const double delta = (v - 0.5) + lam / (2.0 * r);
  const double ddelta = -lam * z / (2.0 * r * r * r);
  const double S = fmax(sqrt(delta * delta + z * z), 1e-20);
  const double dS = (delta * ddelta + z) / S;
  const double den = delta + S;
  B = z / den;
  dB = (den - z * (ddelta + dS)) / (den * den);
}
struct VmfRowScalars { float entropy, log_norm, dentropy, dlog_norm; };
// The four independent fp64 chains behind the row scalars: ive(v, k), ive(v - 1, k) and the two Bessel-ratio bounds.
// One lane can run them back to back (vmf_row_scalars), or four lanes one each (sphere_row_scalars with few rows per
// warp: the serial chain -- two Bessel series with their lgamma -- was 6 us of the 14 us C2 vMF step).
// ive(v, k) as the row scalars consume it.  The reference evaluates log(ive + 1e-20) and d ive / (ive + 1e-20) in float64
// (von_mises_fisher.py:200-212, ops/ive.py:29-34): once ive < e^-110 = 1.7e-48 both are bit-identical to their values at
// ive = 0, and at the reference's 512 / 513-dimensional latents with kappa <= 10 ive is ~e^-750.  From the series,
// ive(v, k) <= (k/2)^v e^{k^2 / (4 (v + 1)) - k} / Gamma(v + 1) and Gamma(v + 1) >= (v / e)^v, so
//   log ive <= v (log(k / (2 v)) + 1) + k^2 / (4 (v + 1)) - k:
// one log and one division decide it, and the Bessel evaluation drops out of the serial chain ahead of the row.
__device__ __forceinline__ double vmf_ive_or_zero(double v, double k) {
  if (v >= 1.0 && v * (log(k / (2.0 * v)) + 1.0) + k * k / (4.0 * (v + 1.0)) - k < -110.0) return 0.0;
  return exp(log_ive(v, k));
}
__device__ __forceinline__ double vmf_piece_ive(double k, int D) { return vmf_ive_or_zero(0.5 * (double)D - 1.0, k); }
__device__ __forceinline__ double vmf_piece_ive_m1(double k, int D) {
  const double v = 0.5 * (double)D - 1.0;
  // ive(v - 1, k); orders below zero only occur for m = 2 (I_{-1} = I_1) and m = 3 (closed form)
  if (v >= 1.0) return vmf_ive_or_zero(v - 1.0, k);
  if (v == 0.0) return exp(log_ive(1.0, k));
  return sqrt(2.0 / (3.14159265358979323846 * k)) * 0.5 * (1.0 + exp(-2.0 * k));
}
__device__ __forceinline__ VmfRowScalars vmf_row_scalars_combine(double k, int D, double ive, double im1, double B0, double dB0,
                                                                  double B2, double dB2);
__device__ __forceinline__ VmfRowScalars vmf_row_scalars(double k, int D) {
  const double m2 = 0.5 * (double)D;
  double B0, dB0, B2, dB2;
  bessel_ratio_bound(m2, k, 0.0, B0, dB0);
  bessel_ratio_bound(m2, k, 2.0, B2, dB2);
  return vmf_row_scalars_combine(k, D, vmf_piece_ive(k, D), vmf_piece_ive_m1(k, D), B0, dB0, B2, dB2);
}
__device__ __forceinline__ VmfRowScalars vmf_row_scalars_combine(double k, int D, double ive, double im1, double B0, double dB0,
                                                                  double B2, double dB2) {
  const double m2 = 0.5 * (double)D, v = m2 - 1.0;
  const double lval = log(ive + 1e-20);
  const double ln = -(v * log(k) - m2 * 1.83787706640934548356 - (k + lval));
  // d ive/dk = ive(v-1) - ive(v) (v + k)/k   (ops/ive.py:29-34)
  const double dive = im1 - ive * (v + k) / k;
  const double dlval = dive / (ive + 1e-20);
  const double dln = -(v / k - (1.0 + dlval));
  const float ln_f = (float)ln;
  const double frac = 0.5 * (B0 + B2), dfrac = 0.5 * (dB0 + dB2);
  VmfRowScalars r;
  r.log_norm = ln_f;
  r.dlog_norm = (float)dln;
  r.entropy = (float)(-k * frac + (double)ln_f);
  r.dentropy = (float)(-frac - k * dfrac + dln);
  return r;
}

// Fused row scalars (entropy / KL / their kappa-derivatives) of parameter row `prow`, written by ONE lane.
template <int FAMILY>
__device__ __forceinline__ void sphere_row_scalars(const SphereParams& p, long long prow) {
  const float raw = __ldg(p.kappa + prow);
  const double kap = (double)head_kappa(p.head, raw);
  const float chain = head_dkappa(p.head, raw);
  if (FAMILY == kFamilyVMF) {
    const VmfRowScalars r = vmf_row_scalars(kap, p.D);
    if (p.entropy) p.entropy[prow] = r.entropy;
    if (p.kl) p.kl[prow] = (float)(p.prior_entropy - (double)r.entropy);
    if (p.dentropy) p.dentropy[prow] = r.dentropy * chain;
    if (p.log_norm) p.log_norm[prow] = r.log_norm;
    if (p.dlog_norm) p.dlog_norm[prow] = r.dlog_norm * chain;
  } else {
    const PsConsts c = ps_consts(kap, 0.5 * (double)(p.D - 1));
    if (p.entropy) p.entropy[prow] = (float)c.entropy;
    if (p.kl) p.kl[prow] = (float)(p.prior_entropy - c.entropy);
    if (p.dentropy) p.dentropy[prow] = (float)c.dentropy * chain;
  }
}

template <int FAMILY>
__global__ void __launch_bounds__(256)
sphere_rsample_kernel(const SphereParams p) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int D = p.D;
  for (long long row = warp; row < p.rows; row += nwarps) {
    float* zr = p.z + row * D;
    if (FAMILY != kFamilyUniform && (p.entropy || p.kl || p.dentropy || p.log_norm) && row < p.loc_rows && lane == 31)
      sphere_row_scalars<FAMILY>(p, row);
    if (FAMILY == kFamilyUniform) {
      float ss = 0.f;
      for (int q = lane; 4 * q < D; q += 32) {
        float g[4];
        normals4(p, row, q, 1, g);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = 4 * q + j;
          if (i < D) { zr[i] = g[j]; ss += g[j] * g[j]; }
        }
      }
      ss = warp_sum(ss);
      const float inv = 1.0f / (sqrtf(ss) + p.norm_eps);
      for (int q = lane; 4 * q < D; q += 32) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = 4 * q + j;
          if (i < D) zr[i] *= inv;        // same lane that stashed it
        }
      }
      continue;
    }
    const long long prow = row % p.loc_rows;
    const float* lr = p.loc + prow * D;
    const float kap = head_kappa(p.head, __ldg(p.kappa + prow));
    // scalar coordinate t
    float t, save0, save1 = 0.f;
    if (FAMILY == kFamilyPS) {
      float tp;
      if (p.tprime) {
        tp = p.tprime[row];
      } else {
        const float half = 0.5f * (float)(D - 1);
        PhiloxKey k = p.key;
        k.stream = 4;
        // lanes 0 / 1 draw the two gammas side by side
        const float al = (lane & 1) ? half : half + (kap + 1e-7f);
        const float gm = gamma_draw_float(al, k, (uint64_t)(row * 2 + (lane & 1)));
        const float x = __shfl_sync(0xffffffffu, gm, 0), y = __shfl_sync(0xffffffffu, gm, 1);
        tp = fminf(fmaxf(x / (x + y), 1.17549435e-38f), 1.0f - 5.9604645e-8f);
      }
      t = 2.0f * tp - 1.0f;
      save0 = tp;
    } else {
      const WoodDraw wd = vmf_draw_w(p, row, kap);
      t = wd.w;
      save0 = wd.w;
      save1 = wd.dw_dkappa;
    }
    // pass 1: stash g in the output row; ||g||^2, ||u||^2, sum g u
    float sgg = 0.f, suu = 0.f, sgu = 0.f;
    for (int q = lane; 4 * q < D; q += 32) {
      float g[4];
      normals4(p, row, q, 0, g);           // element i has tangent index i (element 0 = scalar coordinate)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = 4 * q + j;
        if (i < D) {
          const float u = (i == 0 ? 1.0f : 0.0f) - lr[i];
          suu += u * u;
          if (i > 0) {
            zr[i] = g[j];
            sgg += g[j] * g[j];
            sgu += g[j] * u;
          }
        }
      }
    }
    sgg = warp_sum(sgg); suu = warp_sum(suu); sgu = warp_sum(sgu);
    const float sq = sqrtf(fmaxf(1.0f - t * t, p.clamp_eps));
    const float cg = sq / (sqrtf(sgg) + p.norm_eps);           // y_i = cg * g_i, i >= 1
    const float un = sqrtf(suu);
    const float iu = 1.0f / (un + p.house_eps);
    const float u0 = 1.0f - lr[0];
    const float ydotu = (t * u0 + cg * sgu) * iu;               // y . u_hat
    // pass 2: z = y - 2 (y.u_hat) u_hat
    for (int q = lane; 4 * q < D; q += 32) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = 4 * q + j;
        if (i < D) {
          const float u = ((i == 0 ? 1.0f : 0.0f) - lr[i]) * iu;
          const float y = (i == 0) ? t : cg * zr[i];      // stashed by this same lane in pass 1
          zr[i] = y - 2.0f * ydotu * u;
        }
      }
    }
    if (p.save && lane == 0) { p.save[2 * row] = save0; p.save[2 * row + 1] = save1; }
  }
  if (!p.gnoise) rng_launch_done(p.key);      // device RNG: this launch consumed a counter value
}

// Backward: grad_z -> dloc (rows, D), dkappa (rows).
// Register-resident variant for D <= 128 K (K = 1..8; the reference's latents are D = 512 / 513): one warp per row, lane
// `lane` owns the element groups q = lane + 32 kq (4 consecutive elements each, the unit one Philox call serves).  The
// loc row is loaded ONCE with all loads in flight together (the two-pass kernel above re-reads it and stashes the normals
// in the output row: `long_scoreboard` 52 % of its stalls), the normals stay in registers, z is written once, and the
// compile-time trip counts remove most of the per-element index arithmetic (IMAD/IADD3/LEA were 35 % of its instructions).
__device__ __forceinline__ double shfl_double(double v, int src) {
  return __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(v), src), __shfl_sync(0xffffffffu, __double2loint(v), src));
}
// The same row scalars with the independent fp64 chains of one row spread over a QUAD of lanes: lane 16 + 4 j + part
// serves row j (j < 4) of the warp's batch.  Two steps, so that the (divergent) chains can interleave with the scalar
// draws the first lanes run in between: `pieces` (no convergence point) and `combine` (shuffles inside the quad; all
// 32 lanes must call it).  Used when a warp has at most 4 rows to serve.
struct RowScalarPieces { double r0, r1, r2; float raw; bool mine; long long prow; };
template <int FAMILY>
__device__ __forceinline__ RowScalarPieces sphere_row_scalar_pieces(const SphereParams& p, long long base, long long nwarps, int lane) {
  RowScalarPieces o;
  o.r0 = o.r1 = o.r2 = 0.0; o.raw = 1.0f; o.mine = false; o.prow = 0;
  if (lane < 16) return o;
  const int jr = (lane - 16) >> 2, part = lane & 3;
  o.prow = base + (long long)jr * nwarps;
  o.mine = o.prow < p.rows && o.prow < p.loc_rows;
  if (!o.mine) return o;
  o.raw = __ldg(p.kappa + o.prow);
  const double kap = (double)head_kappa(p.head, o.raw);
  if (FAMILY == kFamilyVMF) {
    // lanes of a quad must run the SAME instruction stream to run side by side (divergent paths of one warp are
    // serialised): the two Bessel series differ only in their order, the two ratio bounds only in `a`
    const double v = 0.5 * (double)p.D - 1.0;
    if (part < 2) {
      if (part == 1 && v < 1.0 && v != 0.0) o.r0 = vmf_piece_ive_m1(kap, p.D);                  // m = 3: closed form
      else o.r0 = vmf_ive_or_zero(part == 0 ? v : (v >= 1.0 ? v - 1.0 : 1.0), kap);
    } else {
      bessel_ratio_bound(0.5 * (double)p.D, kap, part == 2 ? 0.0 : 2.0, o.r0, o.r1);
    }
  } else if (part < 2) {
    // lgamma / digamma / trigamma of a (lane 0 of the quad) and of a + b (lane 1)
    const double half = 0.5 * (double)(p.D - 1);
    gamma_family(part == 0 ? half + (kap + 1e-7) : half + (kap + 1e-7) + half, o.r0, o.r1, o.r2);
  }
  return o;
}
template <int FAMILY>
__device__ __forceinline__ void sphere_row_scalar_combine(const SphereParams& p, const RowScalarPieces& o, int lane) {
  const int qb = lane & ~3;
  const double a0 = shfl_double(o.r0, qb + 1), a1 = shfl_double(o.r1, qb + 1), a2 = shfl_double(o.r2, qb + 1);
  const double b0 = shfl_double(o.r0, qb + 2), b1 = shfl_double(o.r1, qb + 2);
  const double c0 = shfl_double(o.r0, qb + 3), c1 = shfl_double(o.r1, qb + 3);
  if (!o.mine || (lane & 3) != 0) return;
  const double kap = (double)head_kappa(p.head, o.raw);
  const float chain = head_dkappa(p.head, o.raw);
  if (FAMILY == kFamilyVMF) {
    const VmfRowScalars r = vmf_row_scalars_combine(kap, p.D, o.r0, a0, b0, b1, c0, c1);
    if (p.entropy) p.entropy[o.prow] = r.entropy;
    if (p.kl) p.kl[o.prow] = (float)(p.prior_entropy - (double)r.entropy);
    if (p.dentropy) p.dentropy[o.prow] = r.dentropy * chain;
    if (p.log_norm) p.log_norm[o.prow] = r.log_norm;
    if (p.dlog_norm) p.dlog_norm[o.prow] = r.dlog_norm * chain;
  } else {
    const PsConsts c = ps_consts_from(kap, 0.5 * (double)(p.D - 1), o.r0, o.r1, o.r2, a0, a1, a2);
    if (p.entropy) p.entropy[o.prow] = (float)c.entropy;
    if (p.kl) p.kl[o.prow] = (float)(p.prior_entropy - c.entropy);
    if (p.dentropy) p.dentropy[o.prow] = (float)c.dentropy * chain;
  }
}

template <int FAMILY, int K>
__global__ void __launch_bounds__(256)
sphere_rsample_reg_kernel(const SphereParams p) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int D = p.D;
  // A warp owns rows warp, warp + nwarps, ...; it takes them 32 at a time so that the vMF mixture coordinates (Wood's
  // fp64 rejection sampler) of a batch are drawn by 32 lanes in parallel, one row each, instead of redundantly by every
  // lane for one row (that was 37 % of the vMF kernel).  With fewer than 32 rows per warp the idle lanes simply skip.
  for (long long base = warp; base < p.rows; base += 32 * nwarps) {
  float w_l = 0.f, dw_l = 0.f;      // lane j: scalar draw of the batch's j-th row (vMF: w, dw/dkappa; PS: t')
  const bool want_scalars = FAMILY != kFamilyUniform && (p.entropy || p.kl || p.dentropy || p.log_norm);
  // at most 4 rows in this batch (small batches: C2 has one row per warp): each row's independent fp64 chains run on
  // four lanes side by side (lanes 16 .. 31), interleaved with the draws on lanes 0 .. 3; combined after the draws
  const bool quad_scalars = want_scalars && (base + 4 * nwarps >= p.rows);
  RowScalarPieces pieces{};
  if (quad_scalars) pieces = sphere_row_scalar_pieces<FAMILY>(p, base, nwarps, lane);
  if (want_scalars) {
    if (!quad_scalars) {
      // fused entropy / KL of the batch's rows: lane (j + 16) % 32 takes row j, so that the fp64 chains interleave with
      // the draws of the first lanes
      const long long re = base + (long long)((lane + 16) & 31) * nwarps;
      if (re < p.rows && re < p.loc_rows) sphere_row_scalars<FAMILY>(p, re);
    }
  }
  {
    const long long rj = base + (long long)lane * nwarps;
    if (rj < p.rows) {
      const float kapj = head_kappa(p.head, __ldg(p.kappa + rj % p.loc_rows));
      if (FAMILY == kFamilyVMF) {
        const WoodDraw wd = vmf_draw_w(p, rj, kapj);
        w_l = wd.w;
        dw_l = wd.dw_dkappa;
      } else if (p.tprime) {
        w_l = p.tprime[rj];
      } else {
        // t' = X / (X + Y), X ~ Gamma((D-1)/2 + kappa + eps), Y ~ Gamma((D-1)/2): same streams as the two-pass kernel
        const float half = 0.5f * (float)(D - 1);
        PhiloxKey k = p.key;
        k.stream = 4;
        const float x = gamma_draw_float(half + (kapj + 1e-7f), k, (uint64_t)(rj * 2));
        const float y = gamma_draw_float(half, k, (uint64_t)(rj * 2 + 1));
        w_l = fminf(fmaxf(x / (x + y), 1.17549435e-38f), 1.0f - 5.9604645e-8f);
      }
    }
  }
  if (quad_scalars) sphere_row_scalar_combine<FAMILY>(p, pieces, lane);
  for (int jr = 0; jr < 32; ++jr) {
    const long long row = base + (long long)jr * nwarps;
    if (row >= p.rows) break;
    const long long prow = row % p.loc_rows;
    const float* lr = p.loc + prow * D;
    float u[K][4], g[K][4] = {};
#pragma unroll
    for (int kq = 0; kq < K; ++kq) {
      const int i0 = 4 * (lane + 32 * kq);
#pragma unroll
      for (int j = 0; j < 4; ++j) u[kq][j] = (i0 + j < D) ? __ldg(lr + i0 + j) : 0.0f;
    }
    // scalar coordinate t (drawn above for the whole batch)
    float t, save0, save1 = 0.f;
    if (FAMILY == kFamilyPS) {
      const float tp = __shfl_sync(0xffffffffu, w_l, jr);
      t = 2.0f * tp - 1.0f;
      save0 = tp;
    } else {
      t = __shfl_sync(0xffffffffu, w_l, jr);
      save0 = t;
      save1 = __shfl_sync(0xffffffffu, dw_l, jr);
    }
    // tangent normals and the three reductions; u = e1 - loc; slots past D (and the scalar slot 0) hold zeros
    float sgg = 0.f, suu = 0.f, sgu = 0.f;
#pragma unroll
    for (int kq = 0; kq < K; ++kq) {
      const int q = lane + 32 * kq, i0 = 4 * q;
      if (i0 < D) normals4(p, row, q, 0, g[kq]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j;
        const bool in = i < D;
        u[kq][j] = in ? ((i == 0 ? 1.0f : 0.0f) - u[kq][j]) : 0.0f;
        g[kq][j] = (in && i > 0) ? g[kq][j] : 0.0f;
        suu = fmaf(u[kq][j], u[kq][j], suu);
        sgg = fmaf(g[kq][j], g[kq][j], sgg);
        sgu = fmaf(g[kq][j], u[kq][j], sgu);
      }
    }
    sgg = warp_sum(sgg); suu = warp_sum(suu); sgu = warp_sum(sgu);
    const float sq = sqrtf(fmaxf(1.0f - t * t, p.clamp_eps));
    const float cg = sq / (sqrtf(sgg) + p.norm_eps);           // y_i = cg * g_i, i >= 1
    const float un = sqrtf(suu);
    const float iu = 1.0f / (un + p.house_eps);
    const float u0 = __shfl_sync(0xffffffffu, u[0][0], 0);     // 1 - loc_0
    const float ydotu = (t * u0 + cg * sgu) * iu;               // y . u_hat
    const float c2 = 2.0f * ydotu * iu;
    float* zr = p.z + row * D;
#pragma unroll
    for (int kq = 0; kq < K; ++kq) {
      const int i0 = 4 * (lane + 32 * kq);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j;
        if (i < D) {
          const float y = (i == 0) ? t : cg * g[kq][j];
          zr[i] = fmaf(-c2, u[kq][j], y);                      // z = y - 2 (y.u_hat) u_hat
        }
      }
    }
    if (p.save && lane == 0) { p.save[2 * row] = save0; p.save[2 * row + 1] = save1; }
  }
  }
  if (!p.gnoise) rng_launch_done(p.key);      // device RNG: this launch consumed a counter value
}

template <int FAMILY>
__global__ void __launch_bounds__(256)
sphere_rsample_bwd_kernel(const SphereParams p) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int D = p.D;
  for (long long row = warp; row < p.rows; row += nwarps) {
    const long long prow = row % p.loc_rows;
    const float* lr = p.loc + prow * D;
    const float* gz = p.grad_z + row * D;
    float* dl = p.dloc + row * D;
    const float kap_raw = __ldg(p.kappa + prow);
    const float kap = head_kappa(p.head, kap_raw);
    float t, tp = 0.f, dt_dk_direct = 0.f;
    if (FAMILY == kFamilyPS) {
      tp = p.tprime ? p.tprime[row] : p.save[2 * row];
      t = 2.0f * tp - 1.0f;
    } else {
      t = p.save[2 * row];
      dt_dk_direct = p.save[2 * row + 1];
    }
    // pass 1: five reductions; stash g in the dloc row
    float sgg = 0.f, suu = 0.f, sgu = 0.f, sug = 0.f, szg = 0.f;
    for (int q = lane; 4 * q < D; q += 32) {
      float g[4];
      normals4(p, row, q, 0, g);           // replays the forward's draws from the counter-based generator
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = 4 * q + j;
        if (i < D) {
          const float u = (i == 0 ? 1.0f : 0.0f) - lr[i];
          const float gzi = gz[i];
          suu += u * u;
          sug += u * gzi;
          if (i > 0) {
            dl[i] = g[j];
            sgg += g[j] * g[j];
            sgu += g[j] * u;
            szg += gzi * g[j];
          }
        }
      }
    }
    sgg = warp_sum(sgg); suu = warp_sum(suu); sgu = warp_sum(sgu); sug = warp_sum(sug); szg = warp_sum(szg);
    const float om = 1.0f - t * t;
    const float sq = sqrtf(fmaxf(om, p.clamp_eps));
    const float ng = sqrtf(sgg) + p.norm_eps;
    const float cg = sq / ng;
    const float un = sqrtf(suu);
    const float ne = un + p.house_eps;
    const float u0 = 1.0f - lr[0];
    const float a = sug / ne;                                   // grad_z . u_hat
    const float b = (t * u0 + cg * sgu) / ne;                   // y . u_hat
    // grad wrt y = H grad_z; scalar chain through t
    const float gy0 = gz[0] - 2.0f * a * u0 / ne;
    const float s_gv = (szg - 2.0f * a * sgu / ne) / ng;        // grad_y[1:] . v
    const float dsq = (om > p.clamp_eps) ? (-t / sq) : 0.0f;
    const float gt = gy0 + dsq * s_gv;
    // dloc_i = -gu_i,  gu_i = -2 (a y_i + b gz_i)/ne + u_i 4ab / (un ne)
    const float k2 = (un > 0.f) ? 4.0f * a * b / (un * ne) : 0.0f;
    for (int q = lane; 4 * q < D; q += 32) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = 4 * q + j;
        if (i < D) {
          const float u = (i == 0 ? 1.0f : 0.0f) - lr[i];
          const float y = (i == 0) ? t : cg * dl[i];
          dl[i] = 2.0f * (a * y + b * gz[i]) / ne - u * k2;
        }
      }
    }
    if (lane == 0) {
      float dk;
      if (FAMILY == kFamilyPS) {
        const float half = 0.5f * (float)(D - 1);
        const BetaGradConsts bc(half + (kap + 1e-7f), half);
        dk = 2.0f * gt * dirichlet_grad_one(tp, bc) * (1.0f - tp);
      } else {
        dk = gt * dt_dk_direct;
      }
      p.dkappa[row] = dk * head_dkappa(p.head, kap_raw);
    }
  }
}

// Register-resident backward (D <= 128 K), same structure as sphere_rsample_reg_kernel: loc and grad_z are read once with
// all loads in flight, the replayed normals stay in registers, dloc is written once, and the per-row scalar tail (the
// implicit Beta gradient for PowerSpherical, fp32 + the fp64 saddle branch) is evaluated for 32 rows in parallel lanes
// instead of on lane 0 of every row.
template <int FAMILY, int K>
__global__ void __launch_bounds__(256)
sphere_rsample_bwd_reg_kernel(const SphereParams p) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int D = p.D;
  for (long long base = warp; base < p.rows; base += 32 * nwarps) {
    // lane j: saved scalars of the batch's j-th row
    const long long rj = base + (long long)lane * nwarps;
    float s0_l = 0.f, s1_l = 0.f, gt_l = 0.f;
    if (rj < p.rows) {
      if (FAMILY == kFamilyPS) {
        s0_l = p.tprime ? p.tprime[rj] : p.save[2 * rj];
      } else {
        s0_l = p.save[2 * rj];
        s1_l = p.save[2 * rj + 1];
      }
    }
    for (int jr = 0; jr < 32; ++jr) {
      const long long row = base + (long long)jr * nwarps;
      if (row >= p.rows) break;
      const long long prow = row % p.loc_rows;
      const float* lr = p.loc + prow * D;
      const float* gzr = p.grad_z + row * D;
      float u[K][4], gz[K][4], g[K][4] = {};
#pragma unroll
      for (int kq = 0; kq < K; ++kq) {
        const int i0 = 4 * (lane + 32 * kq);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool in = i0 + j < D;
          u[kq][j] = in ? __ldg(lr + i0 + j) : 0.0f;
          gz[kq][j] = in ? ldg_stream1(gzr + i0 + j) : 0.0f;
        }
      }
      const float s0 = __shfl_sync(0xffffffffu, s0_l, jr);
      const float t = (FAMILY == kFamilyPS) ? 2.0f * s0 - 1.0f : s0;
      float sgg = 0.f, suu = 0.f, sgu = 0.f, sug = 0.f, szg = 0.f;
#pragma unroll
      for (int kq = 0; kq < K; ++kq) {
        const int q = lane + 32 * kq, i0 = 4 * q;
        if (i0 < D) normals4(p, row, q, 0, g[kq]);       // replays the forward's draws from the counter-based generator
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = i0 + j;
          const bool in = i < D;
          u[kq][j] = in ? ((i == 0 ? 1.0f : 0.0f) - u[kq][j]) : 0.0f;
          g[kq][j] = (in && i > 0) ? g[kq][j] : 0.0f;
          suu = fmaf(u[kq][j], u[kq][j], suu);
          sug = fmaf(u[kq][j], gz[kq][j], sug);
          sgg = fmaf(g[kq][j], g[kq][j], sgg);
          sgu = fmaf(g[kq][j], u[kq][j], sgu);
          szg = fmaf(gz[kq][j], g[kq][j], szg);
        }
      }
      sgg = warp_sum(sgg); suu = warp_sum(suu); sgu = warp_sum(sgu); sug = warp_sum(sug); szg = warp_sum(szg);
      const float om = 1.0f - t * t;
      const float sq = sqrtf(fmaxf(om, p.clamp_eps));
      const float ng = sqrtf(sgg) + p.norm_eps;
      const float cg = sq / ng;
      const float un = sqrtf(suu);
      const float ne = un + p.house_eps;
      const float u0 = __shfl_sync(0xffffffffu, u[0][0], 0);      // 1 - loc_0
      const float gz0 = __shfl_sync(0xffffffffu, gz[0][0], 0);
      const float a = sug / ne;                                   // grad_z . u_hat
      const float b = (t * u0 + cg * sgu) / ne;                   // y . u_hat
      const float gy0 = gz0 - 2.0f * a * u0 / ne;
      const float s_gv = (szg - 2.0f * a * sgu / ne) / ng;        // grad_y[1:] . v
      const float dsq = (om > p.clamp_eps) ? (-t / sq) : 0.0f;
      const float gt = gy0 + dsq * s_gv;
      if (lane == jr) gt_l = gt;
      // dloc_i = 2 (a y_i + b gz_i)/ne - u_i 4ab / (un ne)
      const float k2 = (un > 0.f) ? 4.0f * a * b / (un * ne) : 0.0f;
      const float ca = 2.0f * a / ne, cb = 2.0f * b / ne;
      float* dl = p.dloc + row * D;
#pragma unroll
      for (int kq = 0; kq < K; ++kq) {
        const int i0 = 4 * (lane + 32 * kq);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = i0 + j;
          if (i < D) {
            const float y = (i == 0) ? t : cg * g[kq][j];
            dl[i] = fmaf(ca, y, fmaf(cb, gz[kq][j], -u[kq][j] * k2));
          }
        }
      }
    }
    // scalar tails of the batch, one row per lane
    if (rj < p.rows) {
      float dk;
      const float kap_raw = __ldg(p.kappa + rj % p.loc_rows);
      if (FAMILY == kFamilyPS) {
        const float half = 0.5f * (float)(D - 1);
        const BetaGradConsts bc(half + (head_kappa(p.head, kap_raw) + 1e-7f), half);
        dk = 2.0f * gt_l * dirichlet_grad_one(s0_l, bc) * (1.0f - s0_l);
      } else {
        dk = gt_l * s1_l;
      }
      p.dkappa[rj] = dk * head_dkappa(p.head, kap_raw);
    }
  }
}

// PowerSpherical.log_prob (clifford.py:198-202): lp = logC(kappa) + kappa log1p(clamp(loc . x)).
// Optional outputs for the backward: coef (rows) = kappa / (1 + dot) inside the clamp else 0
// (d lp / d loc = coef * value, d lp / d value = coef * loc) and dlp_dkappa (rows).
struct SphereLogProbParams {
  const float* value; const float* loc; const float* kappa; long long loc_rows;
  float* log_prob; float* coef; float* dlp_dkappa; long long rows; int D;
};
static __global__ void __launch_bounds__(256) powerspherical_log_prob_kernel(const SphereLogProbParams p) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = warp; row < p.rows; row += nwarps) {
    const long long prow = row % p.loc_rows;
    const float* lr = p.loc + prow * p.D;
    const float* vr = p.value + row * p.D;
    float dot = 0.f;
    for (int i = lane; i < p.D; i += 32) dot += lr[i] * vr[i];
    dot = warp_sum(dot);
    if (lane == 0) {
      const float kap = p.kappa[prow];
      const PsConsts c = ps_consts((double)kap, 0.5 * (double)(p.D - 1));
      const float dc = fminf(fmaxf(dot, -1.0f + 1e-7f), 1.0f - 1e-7f);
      const float l1p = log1pf(dc);
      p.log_prob[row] = (float)c.log_norm + kap * l1p;
      if (p.coef) {
        const bool inside = (dot >= -1.0f + 1e-7f) && (dot <= 1.0f - 1e-7f);
        p.coef[row] = inside ? kap / (1.0f + dc) : 0.0f;
        p.dlp_dkappa[row] = (float)c.dlog_norm + l1p;
      }
    }
  }
}

// log C(kappa) and its derivative for a batch of concentrations (PowerSpherical.log_normalizer)
static __global__ void ps_log_normalizer_kernel(const float* kappa, long long rows, double half_dm1, float* log_norm,
                                                float* dlog_norm) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
    const PsConsts c = ps_consts((double)kappa[i], half_dm1);
    log_norm[i] = (float)c.log_norm;
    if (dlog_norm) dlog_norm[i] = (float)c.dlog_norm;
  }
}

static __global__ void vmf_entropy_kernel(const float* kappa, long long rows, int D, float* entropy, float* log_norm,
                                          float* dentropy, float* dlog_norm) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
    const VmfRowScalars r = vmf_row_scalars((double)kappa[i], D);
    if (log_norm) log_norm[i] = r.log_norm;
    if (dlog_norm) dlog_norm[i] = r.dlog_norm;
    if (entropy) entropy[i] = r.entropy;
    if (dentropy) dentropy[i] = r.dentropy;
  }
}

// VonMisesFisher.log_prob (von_mises_fisher.py:193-212): lp = kappa <loc, x> - log_norm(kappa), log_norm = the
// reference's `_log_normalization` (passed in per parameter row, from the sampler launch or vmf_entropy_kernel).
// Optional `dot` (rows) = <loc, x> for the backward.  One warp per row.
struct VmfLogProbParams {
  const float* value; const float* loc; const float* kappa; const float* log_norm; long long loc_rows;
  float* log_prob; float* dot; long long rows; int D;
};
static __global__ void __launch_bounds__(256) vmf_log_prob_kernel(const VmfLogProbParams p) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = warp; row < p.rows; row += nwarps) {
    const long long prow = row % p.loc_rows;
    const float* lr = p.loc + prow * p.D;
    const float* vr = p.value + row * p.D;
    float dot = 0.f;
    for (int i = lane; i < p.D; i += 32) dot = fmaf(__ldg(lr + i), ldg_stream1(vr + i), dot);
    dot = warp_sum(dot);
    if (lane == 0) {
      p.log_prob[row] = p.kappa[prow] * dot - p.log_norm[prow];
      if (p.dot) p.dot[row] = dot;
    }
  }
}

// Backward helper of the row log-densities lp = f(<loc, x>): with w (rows) = upstream * d lp / d <loc, x>,
//   dvalue[r, :] = w[r] loc[r % loc_rows, :]   and   dloc[q, :] = sum_{s} w[s loc_rows + q] value[s loc_rows + q, :]
// (the sum over the sample dimension is done here, fixed order: deterministic).  Either output may be null.
struct RowScaleParams {
  const float* w; const float* value; const float* loc; long long loc_rows; long long rows; int D;
  float* dvalue; float* dloc;
};
static __global__ void __launch_bounds__(256) row_scale_pair_kernel(const RowScaleParams p) {
  const long long total = p.loc_rows * (long long)p.D;
  const long long S = p.rows / p.loc_rows;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long q = i / p.D;
    const float l = p.dvalue ? __ldg(p.loc + i) : 0.f;
    float acc = 0.f;
    for (long long s = 0; s < S; ++s) {
      const long long r = s * p.loc_rows + q;
      const float w = __ldg(p.w + r);
      if (p.dvalue) p.dvalue[s * total + i] = w * l;
      if (p.dloc) acc = fmaf(w, ldg_stream1(p.value + s * total + i), acc);
    }
    if (p.dloc) p.dloc[i] = acc;
  }
}

}  // namespace cvb
