// Counter-based on-device RNG: Philox4x32-10 (Salmon et al., SC'11) plus the draws the latent
// samplers need.  Every draw is addressed by (seed, call offset, element index, attempt), so
// kernels are reproducible, need no RNG state in memory, and ranks stay disjoint by seed.
#pragma once
#include "common.cuh"

namespace cvb {

struct PhiloxKey {
  uint32_t k0, k1;     // seed
  uint32_t offset;     // per-call offset (host increments it for every launch that draws)
  uint32_t stream;     // which variate family inside a call (0 = gamma proposals, 1 = normals ...)
  // Optional device-resident launch counter added to `offset` (cvb_set_rng_device_counter): under CUDA-graph capture
  // the host-chosen offset is frozen into the graph, so the caller keeps a counter in device memory and bumps it with a
  // captured kernel after every sampling launch -- each replay then draws a fresh stream.  nullptr: host offset only.
  const unsigned long long* dev_counter;
  // Self-bumping mode (cvb_set_rng_device_counter_autobump): a zero-initialised arrival word; the last CTA of every
  // sampling launch to retire adds 1 to *dev_counter and re-arms the word (rng_launch_done), so a captured graph needs
  // no separate `counter += 1` kernel between sampling launches.  nullptr: the caller bumps the counter.
  unsigned int* arrive;
  // Round keys k + r W of the ten rounds, filled by make_key on the host: the kernels' XORs read them straight from the
  // constant bank (kernel parameters) instead of bumping both key words with two integer adds in every round -- 18 of
  // the ~66 instructions of one Philox call.
  uint32_t rk0[10], rk1[10];
};

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    // one 32x32->64 multiply per product (IMAD.WIDE.U32) instead of separate high / low halves
    const uint64_t p0 = (uint64_t)M0 * c.x, p1 = (uint64_t)M1 * c.z;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), hi1 = (uint32_t)(p1 >> 32);
    const uint32_t lo0 = (uint32_t)p0, lo1 = (uint32_t)p1;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += W0;
    k1 += W1;
  }
  return c;
}

// the same rounds with the precomputed round keys of a PhiloxKey (bit-identical to philox4x32_10(c, key.k0, key.k1))
__device__ __forceinline__ uint4 philox4x32_10_rk(uint4 c, const PhiloxKey& key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c.x, p1 = (uint64_t)M1 * c.z;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), hi1 = (uint32_t)(p1 >> 32);
    const uint32_t lo0 = (uint32_t)p0, lo1 = (uint32_t)p1;
    c = make_uint4(hi1 ^ c.y ^ key.rk0[r], lo1, hi0 ^ c.w ^ key.rk1[r], lo0);
  }
  return c;
}
inline void philox_fill_round_keys(PhiloxKey& key) {
  uint32_t k0 = key.k0, k1 = key.k1;
  for (int r = 0; r < 10; ++r) { key.rk0[r] = k0; key.rk1[r] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}

// the call offset of this launch: the host-chosen offset plus the device launch counter when one is registered
__device__ __forceinline__ uint32_t philox_call_offset(const PhiloxKey& key) {
  uint32_t off = key.offset;
  if (key.dev_counter) off += (uint32_t)__ldg(key.dev_counter) * 0x9E3779B9u;   // uniform branch on a kernel parameter
  return off;
}
// hot loops: stream and call offset resolved once by the caller (philox_call_offset), the key read in place (round keys
// from the constant bank)
__device__ __forceinline__ uint4 philox_draw_at(const PhiloxKey& key, uint32_t stream, uint32_t call_offset, uint64_t elem) {
  return philox4x32_10_rk(make_uint4((uint32_t)elem, (uint32_t)(elem >> 32), stream << 24, call_offset), key);
}
// counter layout: (element lo, element hi, attempt | stream << 24, call offset)
__device__ __forceinline__ uint4 philox_draw(const PhiloxKey& key, uint64_t elem, uint32_t attempt) {
  return philox4x32_10_rk(make_uint4((uint32_t)elem, (uint32_t)(elem >> 32), attempt | (key.stream << 24), philox_call_offset(key)), key);
}

// Last statement of every kernel that draws from the device generator; reached by ALL threads of the CTA.  Every CTA has
// read the counter for the last time before it arrives, the last arrival bumps it; the next launch on the stream (also
// one that was admitted early by programmatic dependent launch: it reads nothing before griddepcontrol.wait) sees the new
// value.  One arrival word per device: launches in this mode must be serialised on one stream.
__device__ __forceinline__ void rng_launch_done(const PhiloxKey& key) {
  if (!key.arrive) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(key.arrive, 1u) == gridDim.x - 1) {
      *key.arrive = 0u;
      atomicAdd(const_cast<unsigned long long*>(key.dev_counter), 1ULL);
    }
  }
}

// uniform in (0, 1]: never 0, so logs are finite
__device__ __forceinline__ float u01_open0(uint32_t x) { return fmaf((float)(x >> 8), 0x1p-24f, 0x1p-24f); }
// uniform in [0, 1)
__device__ __forceinline__ float u01_open1(uint32_t x) { return (float)(x >> 8) * 0x1p-24f; }
__device__ __forceinline__ double u01_double(uint32_t hi, uint32_t lo) {
  // 53-bit uniform in (0,1)
  const uint64_t m = ((uint64_t)hi << 21) ^ (lo >> 11);
  return ((double)(m & ((1ull << 53) - 1)) + 0.5) * 0x1p-53;
}

// Box-Muller: two independent N(0,1) from two 32-bit words
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u = u01_open0(a);
  const float r = fast_sqrt(-2.0f * __logf(u));
  float s, c;
  __sincosf(6.283185307179586f * u01_open1(b), &s, &c);
  return make_float2(r * c, r * s);
}
__device__ __forceinline__ double2 box_muller_double(uint4 r) {
  const double u = u01_double(r.x, r.y), v = u01_double(r.z, r.w);
  const double rad = sqrt(-2.0 * log(u));
  double s, c;
  sincospi(2.0 * v, &s, &c);
  return make_double2(rad * c, rad * s);
}

// Marsaglia-Tsang (2000) constants for Gamma(alpha), alpha > 0: boosts alpha < 1.
struct GammaMT {
  float d, c, inv_alpha;   // inv_alpha > 0 only when the alpha+1 boost is active
  __device__ __forceinline__ explicit GammaMT(float alpha) {
    inv_alpha = 0.f;
    if (alpha < 1.0f) { inv_alpha = 1.0f / alpha; alpha += 1.0f; }
    d = alpha - 0.333333333333f;
    c = rsqrtf(9.0f * d);
  }
};

// One Marsaglia-Tsang attempt from proposal normal x and uniform u (0,1]; returns accepted flag.
__device__ __forceinline__ bool gamma_mt_attempt(const GammaMT& g, float x, float u, float& out) {
  // branch-free: in a warp some lane nearly always needs the log test, so every lane evaluates it
  const float y = fmaf(g.c, x, 1.0f);
  const float v = y * y * y;
  const float xx = x * x;
  const float rhs = fmaf(0.5f, xx, g.d * (1.0f - v + __logf(fmaxf(v, 1e-30f))));
  out = g.d * v;
  return (y > 0.f) & ((u < fmaf(-0.0331f * xx, xx, 1.0f)) | (__logf(u) < rhs));
}

// ---- exact sampler for the circle's power-spherical phase ------------------------------------------
// t' ~ Beta(1/2 + k, 1/2) is cos^2(psi) for a half-angle psi in (-pi/2, pi/2) with density ~ cos^{2k}(psi)
// (phi = 2 psi is the phase whose density is ~ (1 + cos phi)^k, reference dists/clifford.py:124-134 with
// dim = 2).  Rejection from a cheap envelope, two proposals per Philox call, no Gamma draws:
//   k >= 1/pi : psi ~ N(0, 1/(2k)), accept if |psi| < pi/2 and u < cos^{2k}(psi) e^{k psi^2}  (cos x <= e^{-x^2/2})
//   k <  1/pi : psi ~ U(-pi/2, pi/2), accept if u < cos^{2k}(psi)
// Acceptance >= 0.72 per proposal, so >= 92 % of the elements finish in the first call; the rest are
// queued by the caller.  The sign of psi is the circle's (independent, fair) sign draw.
struct HalfAngle {
  float k2;        // 2 k
  float kl2e;      // k * log2(e)
  float sigma;     // sqrt(1 / (2k)); 0 selects the uniform proposal
  __device__ __forceinline__ explicit HalfAngle(float k) {
    k2 = 2.0f * k;
    kl2e = k * 1.4426950408889634f;
    sigma = (k >= 0.3183098861837907f) ? rsqrtf(k2) : 0.0f;
  }
};
__device__ __forceinline__ bool half_angle_accept(const HalfAngle& h, float psi, uint32_t uword) {
  const float c = __cosf(psi);
  const float lhs = __log2f(u01_open0(uword));
  const float rhs = fmaf(h.k2, __log2f(fmaxf(c, 1e-30f)), (h.sigma > 0.f) ? h.kl2e * psi * psi : 0.0f);
  return (fabsf(psi) < 1.5707962f) & (lhs < rhs);
}
// proposals from one Philox result: returns true and (t', sign) if either was accepted
__device__ __forceinline__ bool half_angle_try(const HalfAngle& h, uint4 r, float& tp, float& sign) {
  float p0, p1;
  uint32_t u0, u1;
  if (h.sigma > 0.f) {
    const float2 nn = box_muller(r.x, r.y);
    p0 = nn.x * h.sigma; p1 = nn.y * h.sigma;
    u0 = r.z; u1 = r.w;
  } else {
    p0 = (u01_open1(r.x) - 0.5f) * 3.14159265358979f; p1 = (u01_open1(r.z) - 0.5f) * 3.14159265358979f;
    u0 = r.y; u1 = r.w;
  }
  const bool a0 = half_angle_accept(h, p0, u0), a1 = half_angle_accept(h, p1, u1);
  const float psi = a0 ? p0 : p1;
  const float c = __cosf(psi);
  tp = fminf(fmaxf(c * c, 1.17549435e-38f), 1.0f - 5.9604645e-8f);   // torch's Beta clamp range
  sign = (psi < 0.f) ? -1.0f : 1.0f;
  return a0 | a1;
}
// One Philox result serves TWO elements with one proposal each (a rejected one is queued):
// element j in {0, 1} of the pair gets (t', sign) and returns its acceptance in bit j.
// `r2`: a SECOND, independent Philox result, used only in the mixed case (one bin in the normal regime, the other in
// the uniform one -- per-element concentrations straddling 1/pi): there the normal proposal depends on both r.x and r.y
// through the Box-Muller pair, so the uniform bin draws its proposal and accept word from r2 instead of re-using r.y
// (which made the two circles' proposals dependent).  Callers pass r2 = r when both bins share a regime.
__device__ __forceinline__ uint32_t half_angle_pair(const HalfAngle& h0, const HalfAngle& h1, uint4 r, uint4 r2,
                                                    float (&tp)[2], float (&sign)[2]) {
  float p0, p1;
  uint32_t u0 = r.z, u1 = r.w;
  const float2 nn = box_muller(r.x, r.y);
  const bool n0 = h0.sigma > 0.f, n1 = h1.sigma > 0.f;
  if (n0 == n1) {
    p0 = n0 ? nn.x * h0.sigma : (u01_open1(r.x) - 0.5f) * 3.14159265358979f;
    p1 = n1 ? nn.y * h1.sigma : (u01_open1(r.y) - 0.5f) * 3.14159265358979f;
  } else if (n0) {
    p0 = nn.x * h0.sigma;
    p1 = (u01_open1(r2.x) - 0.5f) * 3.14159265358979f; u1 = r2.y;
  } else {
    p1 = nn.y * h1.sigma;
    p0 = (u01_open1(r2.x) - 0.5f) * 3.14159265358979f; u0 = r2.y;
  }
  const bool a0 = half_angle_accept(h0, p0, u0), a1 = half_angle_accept(h1, p1, u1);
  const float c0 = __cosf(p0), c1 = __cosf(p1);
  tp[0] = fminf(fmaxf(c0 * c0, 1.17549435e-38f), 1.0f - 5.9604645e-8f);
  tp[1] = fminf(fmaxf(c1 * c1, 1.17549435e-38f), 1.0f - 5.9604645e-8f);
  sign[0] = (p0 < 0.f) ? -1.0f : 1.0f;
  sign[1] = (p1 < 0.f) ? -1.0f : 1.0f;
  return (a0 ? 1u : 0u) | (a1 ? 2u : 0u);
}
__device__ __forceinline__ bool circle_beta_first(const HalfAngle& h, const PhiloxKey& key, uint64_t elem, float& tp,
                                                  float& sign) {
  return half_angle_try(h, philox_draw(key, elem, 0), tp, sign);
}
__device__ __forceinline__ float circle_beta_retry(const HalfAngle& h, const PhiloxKey& key, uint64_t elem, float& sign) {
  float tp = 0.5f;
  bool ok = false;
  for (uint32_t attempt = 1; !ok && attempt < 256; ++attempt) ok = half_angle_try(h, philox_draw(key, elem, attempt), tp, sign);
  return tp;
}

// Gamma(alpha) in double for alpha >= 1 (Marsaglia-Tsang), used by the per-row Beta draws of the
// D-dimensional samplers (alpha ~ D/2, one draw per row, so cost is irrelevant).
__device__ inline double gamma_mt_double(double alpha, const PhiloxKey& key, uint64_t elem, uint32_t& attempt) {
  double boost = 1.0;
  if (alpha < 1.0) {
    uint4 r = philox_draw(key, elem, attempt++);
    boost = pow(u01_double(r.x, r.y), 1.0 / alpha);
    alpha += 1.0;
  }
  const double d = alpha - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (;;) {
    const uint4 r1 = philox_draw(key, elem, attempt++);
    const uint4 r2 = philox_draw(key, elem, attempt++);
    const double2 nn = box_muller_double(r1);
    const double x = nn.x, y = 1.0 + c * x;
    if (y <= 0.0) continue;
    const double v = y * y * y, u = u01_double(r2.x, r2.y);
    if (log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return boost * d * v;
  }
}
__device__ inline double beta_draw_double(double a, double b, const PhiloxKey& key, uint64_t elem, uint32_t& attempt) {
  const double x = gamma_mt_double(a, key, elem, attempt);
  const double y = gamma_mt_double(b, key, elem, attempt);
  return x / (x + y);
}

}  // namespace cvb
