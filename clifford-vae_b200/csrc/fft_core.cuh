// Batched power-of-two FFT engine: register-resident Stockham passes with shared-memory
// exchanges, plus the half-length real-FFT (R2C / C2R) untangling steps.
//
// One FFT of N = 2^LOG2N complex points is owned by T = N/E consecutive threads ("a group");
// thread t of the group holds the E points {t + e*T}.  Every pass applies radix-R butterflies
// (R <= E, E/R butterflies per thread) entirely in registers; between passes the group
// transposes through its padded shared-memory exchange buffer.  The ownership pattern
// {t + e*T} is the same at the input of every pass and at the final output, so callers load
// and store global memory with consecutive threads touching consecutive float2 (coalesced).
//
// All functions that touch the exchange buffer call group_sync(): every thread of a group must call
// them the same number of times (callers keep loop trip counts uniform).
#pragma once
#include "common.cuh"

namespace cvb {

// LOGE = log2(points per thread).  The default plan keeps 16 points per thread for every N >= 256; the bind kernels use
// 32 points per thread at N = 8192 (radix 32 x 32 x 8 on 256 threads): with 16 points that size needs 512 threads, which
// caps the kernel at 128 registers per thread (it spilled) and costs a fourth pass.  (The Clifford kernels at N = 8192
// were measured 11-24 % SLOWER on the wide plan -- their sampling loops want the extra warps -- so it is opt-in; a
// two-pass 32 x 32 plan for bind at N = 1024 was also measured: 43.9 % vs 47.1 % of the HBM roofline, not adopted.)
constexpr int default_loge(int log2n) { return (log2n >= 8) ? 4 : (log2n >= 6 ? 3 : 2); }

template <int LOG2N_, int LOGE_, int MINTHREADS_ = 128>
struct FftPlanT {
  static_assert(LOG2N_ >= 4 && LOG2N_ <= 13, "fast path covers N = 16 .. 8192 complex points");
  static constexpr int LOG2N = LOG2N_;
  static constexpr int N = 1 << LOG2N;
  static constexpr int LOGE = LOGE_;
  static constexpr int E = 1 << LOGE;              // points per thread
  static constexpr int T = N / E;                  // threads per FFT
  static constexpr int LOGT = LOG2N - LOGE;
  static constexpr int XCH = pad16(N) + 2;         // float2 slots in one exchange buffer (index N usable)
  static constexpr int THREADS = (T >= MINTHREADS_) ? T : MINTHREADS_;   // CTA size
  static constexpr int GROUPS = THREADS / T;       // FFTs processed side by side in one CTA
};
template <int LOG2N>
using FftPlan = FftPlanT<LOG2N, default_loge(LOG2N)>;
template <int LOG2N>
using WideFftPlan = FftPlanT<LOG2N, (LOG2N >= 13) ? 5 : default_loge(LOG2N)>;

// Barrier among the T threads that own one FFT.  A group of <= 32 threads is (part of) one warp:
// __syncwarp; a group that is a proper part of the CTA uses its own named barrier, so the groups of a
// CTA run independently; a group that is the whole CTA uses __syncthreads.
template <class Pl>
__device__ __forceinline__ void group_sync_p() {
  if constexpr (Pl::T <= 32) {
    __syncwarp();
  } else if constexpr (Pl::T < Pl::THREADS) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)(threadIdx.x / Pl::T)), "n"(Pl::T) : "memory");
  } else {
    __syncthreads();
  }
}
template <int LOG2N>
__device__ __forceinline__ void group_sync() { group_sync_p<FftPlan<LOG2N>>(); }

// ---- small in-register DFTs (natural order in, natural order out) ---------------------------
template <bool INV>
__device__ __forceinline__ void dft2(cplx& a, cplx& b) {
  cplx t = a;
  a = cadd(t, b);
  b = csub(t, b);
}

template <bool INV>
__device__ __forceinline__ void dft4(cplx& a0, cplx& a1, cplx& a2, cplx& a3) {
  cplx s0 = cadd(a0, a2), d0 = csub(a0, a2);
  cplx s1 = cadd(a1, a3), d1 = csub(a1, a3);
  cplx r = INV ? cmul_i(d1) : cmul_mi(d1);   // (-i)^1 * d1 forward, (+i) * d1 inverse
  a0 = cadd(s0, s1);
  a2 = csub(s0, s1);
  a1 = cadd(d0, r);
  a3 = csub(d0, r);
}

// z * exp(-+ i*phi) with (c, s) = (cos phi, sin phi): forward uses exp(-i phi)
template <bool INV>
__device__ __forceinline__ cplx rot(cplx z, float c, float s) {
  // z * c + swap(z) * (-+s, +-s): two packed instructions
  return INV ? __ffma2_rn(z, make_float2(c, c), __fmul2_rn(make_float2(z.y, z.x), make_float2(-s, s)))
             : __ffma2_rn(z, make_float2(c, c), __fmul2_rn(make_float2(z.y, z.x), make_float2(s, -s)));
}

template <int R, bool INV>
struct Dft;

template <bool INV>
struct Dft<2, INV> {
  static __device__ __forceinline__ void run(cplx (&u)[2]) { dft2<INV>(u[0], u[1]); }
};
template <bool INV>
struct Dft<4, INV> {
  static __device__ __forceinline__ void run(cplx (&u)[4]) { dft4<INV>(u[0], u[1], u[2], u[3]); }
};
template <bool INV>
struct Dft<8, INV> {
  static __device__ __forceinline__ void run(cplx (&u)[8]) {
    constexpr float h = 0.70710678118654752440f;
    dft4<INV>(u[0], u[2], u[4], u[6]);     // even samples -> E[k] in u[0],u[2],u[4],u[6]
    dft4<INV>(u[1], u[3], u[5], u[7]);     // odd samples  -> O[k] in u[1],u[3],u[5],u[7]
    cplx o0 = u[1];
    cplx o1 = rot<INV>(u[3], h, h);
    cplx o2 = INV ? cmul_i(u[5]) : cmul_mi(u[5]);
    cplx o3 = rot<INV>(u[7], -h, h);
    cplx e0 = u[0], e1 = u[2], e2 = u[4], e3 = u[6];
    u[0] = cadd(e0, o0); u[4] = csub(e0, o0);
    u[1] = cadd(e1, o1); u[5] = csub(e1, o1);
    u[2] = cadd(e2, o2); u[6] = csub(e2, o2);
    u[3] = cadd(e3, o3); u[7] = csub(e3, o3);
  }
};
template <bool INV>
struct Dft<16, INV> {
  static __device__ __forceinline__ void run(cplx (&u)[16]) {
    constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
    constexpr float h = 0.70710678118654752440f;
    // stage 1: for each n1, DFT-4 over n2 of x[n1 + 4 n2]  -> Y[n1][k2] left at slot n1 + 4 k2
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) dft4<INV>(u[n1], u[n1 + 4], u[n1 + 8], u[n1 + 12]);
    // stage 2: Y[n1][k2] *= W16^(n1 k2)
    u[5] = rot<INV>(u[5], c1, s1);                       // m = 1
    u[6] = rot<INV>(u[6], h, h);                         // m = 2
    u[7] = rot<INV>(u[7], s1, c1);                       // m = 3
    u[9] = rot<INV>(u[9], h, h);                         // m = 2
    u[10] = INV ? cmul_i(u[10]) : cmul_mi(u[10]);        // m = 4
    u[11] = rot<INV>(u[11], -h, h);                      // m = 6
    u[13] = rot<INV>(u[13], s1, c1);                     // m = 3
    u[14] = rot<INV>(u[14], -h, h);                      // m = 6
    u[15] = rot<INV>(u[15], -c1, -s1);                   // m = 9
    // stage 3: for each k2, DFT-4 over n1 -> X[4 k1 + k2] left at slot k1 + 4 k2
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) dft4<INV>(u[4 * k2], u[4 * k2 + 1], u[4 * k2 + 2], u[4 * k2 + 3]);
    // 4x4 register transpose (renaming only after unrolling)
    cplx t;
#define CVB_SWAP(a, b) t = u[a]; u[a] = u[b]; u[b] = t;
    CVB_SWAP(1, 4) CVB_SWAP(2, 8) CVB_SWAP(3, 12) CVB_SWAP(6, 9) CVB_SWAP(7, 13) CVB_SWAP(11, 14)
#undef CVB_SWAP
  }
};

template <bool INV>
struct Dft<32, INV> {
  // decimation in time over two 16-point transforms: X[k] = E[k] + W32^k O[k], X[k + 16] = E[k] - W32^k O[k]
  static __device__ __forceinline__ void run(cplx (&u)[32]) {
    constexpr float c[16] = {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                             0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f,
                             0.0f, -0.19509032201612826785f, -0.38268343236508977173f, -0.55557023301960222474f,
                             -0.70710678118654752440f, -0.83146961230254523708f, -0.92387953251128675613f, -0.98078528040323044913f};
    constexpr float s[16] = {0.0f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                             0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f, 0.98078528040323044913f,
                             1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                             0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f};
    cplx ev[16], od[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { ev[i] = u[2 * i]; od[i] = u[2 * i + 1]; }
    Dft<16, INV>::run(ev);
    Dft<16, INV>::run(od);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      cplx o;
      if (k == 0) o = od[0];
      else if (k == 8) o = INV ? cmul_i(od[8]) : cmul_mi(od[8]);
      else o = rot<INV>(od[k], c[k], s[k]);
      u[k] = cadd(ev[k], o);
      u[k + 16] = csub(ev[k], o);
    }
  }
};


// 64-point DFT as 8 x 8 (decimation n = n1 + 8 n2, k = 8 k1 + k2): used by the 64-points-per-thread plans (one warp per
// 2048-point transform, ONE shared-memory exchange per transform instead of two).
template <bool INV>
struct Dft<64, INV> {
  static __device__ __forceinline__ void run(cplx (&u)[64]) {
    // cos / sin of 2 pi m / 64, m = 0 .. 15 (first quadrant; the rest by symmetry)
    constexpr float C[17] = {1.0f, 0.99518472667219688624f, 0.98078528040323044913f, 0.95694033573220886494f,
                             0.92387953251128675613f, 0.88192126434835502971f, 0.83146961230254523708f, 0.77301045336273696081f,
                             0.70710678118654752440f, 0.63439328416364549822f, 0.55557023301960222474f, 0.47139673682599764856f,
                             0.38268343236508977173f, 0.29028467725446236764f, 0.19509032201612826785f, 0.09801714032956060199f, 0.0f};
    // stage 1: for each n1, DFT-8 over n2 of x[n1 + 8 n2] -> Y[n1][k2] left at slot n1 + 8 k2
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) {
      cplx w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = u[n1 + 8 * j];
      Dft<8, INV>::run(w);
#pragma unroll
      for (int j = 0; j < 8; ++j) u[n1 + 8 * j] = w[j];
    }
    // stage 2: Y[n1][k2] *= W64^(n1 k2)
#pragma unroll
    for (int n1 = 1; n1 < 8; ++n1) {
#pragma unroll
      for (int k2 = 1; k2 < 8; ++k2) {
        const int m = n1 * k2;              // 1 .. 49
        const int q = m >> 4, r = m & 15;   // quadrant, offset: angle = (16 q + r) 2 pi / 64
        // cos / sin of the first-quadrant offset, then rotate by q quarter turns
        const float c0 = C[r], s0 = C[16 - r];
        const float c = (q == 0) ? c0 : (q == 1) ? -s0 : (q == 2) ? -c0 : s0;
        const float sn = (q == 0) ? s0 : (q == 1) ? c0 : (q == 2) ? -s0 : -c0;
        cplx& z = u[n1 + 8 * k2];
        if (r == 0) {
          z = (q == 1) ? (INV ? cmul_i(z) : cmul_mi(z)) : (q == 2) ? make_float2(-z.x, -z.y) : (INV ? cmul_mi(z) : cmul_i(z));
        } else {
          z = rot<INV>(z, c, sn);
        }
      }
    }
    // stage 3: for each k2, DFT-8 over n1 -> X[8 k1 + k2] left at slot k1 + 8 k2
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {
      cplx w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = u[8 * k2 + j];
      Dft<8, INV>::run(w);
#pragma unroll
      for (int j = 0; j < 8; ++j) u[8 * k2 + j] = w[j];
    }
    // 8 x 8 register transpose (renaming only after unrolling): slot k1 + 8 k2 holds X[8 k1 + k2]
#pragma unroll
    for (int a = 0; a < 8; ++a) {
#pragma unroll
      for (int b = a + 1; b < 8; ++b) {
        const cplx tmp = u[a + 8 * b];
        u[a + 8 * b] = u[b + 8 * a];
        u[b + 8 * a] = tmp;
      }
    }
  }
};

// u[r] *= w1^r, r = 1..R-1 (powers built by squaring / one extra multiply: depth <= 2 log2 R)
template <int R>
__device__ __forceinline__ void apply_twiddle_powers(cplx (&u)[R], cplx w1) {
  if constexpr (R > 16) {
    // streaming form for the wide butterflies: w^r = w^(r-1) w, re-anchored by squaring at every power of two
    // (w^2, w^4, ... are the only powers kept live) so the chain's round-off stays at a few ulp
    cplx pw = w1, cur = w1;              // pw: last power-of-two power; cur: w^(r-1)
#pragma unroll
    for (int i = 1; i < R; ++i) {
      if (i > 1) {
        if ((i & (i - 1)) == 0) { pw = cmul(pw, pw); cur = pw; }
        else cur = cmul(cur, w1);
      }
      u[i] = cmul(u[i], cur);
    }
    return;
  }
  cplx w[R];
  w[1] = w1;
#pragma unroll
  for (int i = 2; i < R; ++i) {
    if (i & 1) {
      w[i] = cmul(w[i - 1], w1);
    } else {
      w[i] = cmul(w[i >> 1], w[i >> 1]);
    }
  }
#pragma unroll
  for (int i = 1; i < R; ++i) u[i] = cmul(u[i], w[i]);
}

// This thread's points {t + e T} of a padded exchange buffer.  pad16(t + e T) = pad16(t) + e (T + T/16) whenever T is a
// multiple of 16 (adding a multiple of 16 never carries out of the low four bits), so one base address and compile-time
// offsets replace a shift / add / scale per point.
template <class Pl>
__device__ __forceinline__ void load_own_points(cplx (&v)[Pl::E], const cplx* xch, int t) {
  if constexpr (Pl::T % 16 == 0) {
    constexpr int S = Pl::T + Pl::T / 16;
    const cplx* xr = xch + pad16(t);
#pragma unroll
    for (int e = 0; e < Pl::E; ++e) v[e] = xr[e * S];
  } else {
#pragma unroll
    for (int e = 0; e < Pl::E; ++e) v[e] = xch[pad16(t + e * Pl::T)];
  }
}

// BASEPTR: read the exchanged points through load_own_points (the Clifford kernels: fewer address instructions in their
// issue-bound loops); the bind kernels keep the indexed form (measured: -0.8 % at d = 4096 with the base-pointer form).
template <class Pl, int P, bool INV, bool BASEPTR = false>
__device__ __forceinline__ void fft_pass_p(cplx (&v)[Pl::E], cplx* xch, const int t, const cplx* __restrict__ tw) {
  constexpr int LOG2N = Pl::LOG2N;
  constexpr int LOGNS = P * Pl::LOGE;
  constexpr int LOGR = (LOG2N - LOGNS) < Pl::LOGE ? (LOG2N - LOGNS) : Pl::LOGE;
  constexpr int R = 1 << LOGR, Q = Pl::E / R, NS = 1 << LOGNS;
  constexpr bool LAST = (LOGNS + LOGR == LOG2N);
  if (!LAST) group_sync_p<Pl>();                // earlier readers of xch are done
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    cplx u[R];
#pragma unroll
    for (int r = 0; r < R; ++r) u[r] = v[q + r * Q];
    const int j = t + q * Pl::T;
    const int k = j & (NS - 1);
    if (P > 0) {
      cplx w1 = __ldg(&tw[twiddle_offset(LOG2N) + (k << (LOG2N + 1 - LOGNS - LOGR))]);   // exp(-2 pi i k / (NS R))
      if (INV) w1.y = -w1.y;
      apply_twiddle_powers<R>(u, w1);
    }
    Dft<R, INV>::run(u);
    if (LAST) {
#pragma unroll
      for (int r = 0; r < R; ++r) v[q + r * Q] = u[r];
    } else {
      const int j0 = ((j - k) << LOGR) + k;
#pragma unroll
      for (int r = 0; r < R; ++r) xch[pad16(j0 + r * NS)] = u[r];
    }
  }
  if constexpr (!LAST) {
    group_sync_p<Pl>();
    if constexpr (BASEPTR) {
      load_own_points<Pl>(v, xch, t);
    } else {
#pragma unroll
      for (int e = 0; e < Pl::E; ++e) v[e] = xch[pad16(t + e * Pl::T)];
    }
    fft_pass_p<Pl, P + 1, INV, BASEPTR>(v, xch, t, tw);
  }
}

// In: v[e] = x[t + e T].  Out: v[e] = X[t + e T], X[k] = sum_j x[j] exp(-+ 2 pi i j k / N) (unnormalised).
template <class Pl, bool INV>
__device__ __forceinline__ void fft_run_p(cplx (&v)[Pl::E], cplx* xch, int t, const cplx* __restrict__ tw) {
  fft_pass_p<Pl, 0, INV>(v, xch, t, tw);
}
template <int LOG2N, bool INV>
__device__ __forceinline__ void fft_run(cplx (&v)[FftPlan<LOG2N>::E], cplx* xch, int t,
                                        const cplx* __restrict__ tw) {
  fft_pass_p<FftPlan<LOG2N>, 0, INV, true>(v, xch, t, tw);
}

// R2C untangle.  In: v = Z = FFT_N(z), z[m] = x[2m] + i x[2m+1] of a real row of length n = 2N.
// Out: v[e] = X[k], k = t + e T, the first N bins of the length-n real FFT; returns X[N] (real,
// meaningful on the thread with t == 0 only).
template <int LOG2N>
__device__ __forceinline__ float r2c_untangle(cplx (&v)[FftPlan<LOG2N>::E], cplx* xch, int t,
                                              const cplx* __restrict__ tw) {
  using Pl = FftPlan<LOG2N>;
  constexpr int N = Pl::N;
  if constexpr (Pl::T % 16 == 0) {
    // base addresses with compile-time offsets (see load_own_points): own slots from pad16(t), partners N - k downwards from
    // pad16(N - t); the one wrap-around (k = 0 pairs with itself) is a select in the first iteration of thread 0
    constexpr int S = Pl::T + Pl::T / 16;
    cplx* xk = xch + pad16(t);
    const cplx* xm = xch + pad16(N - t);
    const cplx* twp = tw + twiddle_offset(LOG2N) + t;
    group_sync<LOG2N>();
#pragma unroll
    for (int e = 0; e < Pl::E; ++e) xk[e * S] = v[e];
    group_sync<LOG2N>();
    const float nyq0 = v[0].x - v[0].y;      // for t == 0: Re Z0 - Im Z0
#pragma unroll
    for (int e = 0; e < Pl::E; ++e) {
      const cplx z = v[e];
      const cplx* pp = (e == 0) ? ((t == 0) ? xch : xm) : xm - e * S;
      const cplx zp = cconj(*pp);
      const cplx w = __ldg(twp + e * Pl::T);                              // exp(-2 pi i k / n)
      const cplx s = cadd(z, zp), d = csub(z, zp);
      const cplx wd = cmul_mi(cmul(w, d));                                 // -i w (z - zp)
      v[e] = cadd_scaled(s, wd, 0.5f);
    }
    return nyq0;
  }
  group_sync<LOG2N>();
#pragma unroll
  for (int e = 0; e < Pl::E; ++e) xch[pad16(t + e * Pl::T)] = v[e];
  group_sync<LOG2N>();
  const float nyq = v[0].x - v[0].y;      // for t == 0: Re Z0 - Im Z0
#pragma unroll
  for (int e = 0; e < Pl::E; ++e) {
    const int k = t + e * Pl::T;
    const cplx z = v[e];
    const cplx zp = cconj(xch[pad16((N - k) & (N - 1))]);
    const cplx w = __ldg(&tw[twiddle_offset(LOG2N) + k]);   // exp(-2 pi i k / n)
    const cplx s = cadd(z, zp), d = csub(z, zp);
    const cplx wd = cmul_mi(cmul(w, d));                                 // -i w (z - zp)
    v[e] = cadd_scaled(s, wd, 0.5f);
  }
  return nyq;
}

// C2R pre-processing.  In: v[e] = X[k] (k = t + e T, bins 0..N-1 of a Hermitian half spectrum),
// x_nyq = X[N] (real; only the value passed by t == 0 is used).  Out: v = Z such that the
// unnormalised inverse FFT_N(Z)[m] = (x[2m], x[2m+1]) with x = irfft(X, n = 2N) (1/n included).
template <int LOG2N>
__device__ __forceinline__ void c2r_pretangle(cplx (&v)[FftPlan<LOG2N>::E], float x_nyq, cplx* xch, int t,
                                              const cplx* __restrict__ tw) {
  using Pl = FftPlan<LOG2N>;
  constexpr int N = Pl::N;
  constexpr float scale = 1.0f / (2.0f * N);
  group_sync<LOG2N>();
#pragma unroll
  for (int e = 0; e < Pl::E; ++e) xch[pad16(t + e * Pl::T)] = v[e];
  if (t == 0) xch[pad16(N)] = make_float2(x_nyq, 0.0f);
  group_sync<LOG2N>();
#pragma unroll
  for (int e = 0; e < Pl::E; ++e) {
    const int k = t + e * Pl::T;
    const cplx x = v[e];
    const cplx xp = cconj(xch[pad16(N - k)]);
    const cplx w = cconj(__ldg(&tw[twiddle_offset(LOG2N) + k]));   // exp(+2 pi i k / n)
    const cplx s = cadd(x, xp), d = csub(x, xp);
    const cplx wd = cmul_i(cmul(w, d));                                       // +i w (x - xp)
    v[e] = cadd_scaled(s, wd, scale);
  }
}

// C2R pre-processing when the half spectrum X[0..N] already sits in the exchange buffer (written by
// the caller, followed by a __syncthreads()).  Out: v = Z as in c2r_pretangle.
// OWN: on entry v[e] already holds this thread's own bins X[t + e T] (the caller kept what it wrote to the buffer in
// registers): only the partner bins are read.
template <int LOG2N, bool OWN = false>
__device__ __forceinline__ void c2r_pretangle_load(cplx (&v)[FftPlan<LOG2N>::E], const cplx* xch, int t,
                                                   const cplx* __restrict__ tw) {
  using Pl = FftPlan<LOG2N>;
  constexpr int N = Pl::N;
  constexpr float scale = 1.0f / (2.0f * N);
  if constexpr (Pl::T % 16 == 0) {
    // bins k = t + e T and N - k from two base addresses with compile-time offsets (see load_own_points), the twiddles
    // from one base pointer
    constexpr int S = Pl::T + Pl::T / 16;
    const cplx* xk = xch + pad16(t);
    const cplx* xm = xch + pad16(N - t);
    const cplx* twp = tw + twiddle_offset(LOG2N) + t;
#pragma unroll
    for (int e = 0; e < Pl::E; ++e) {
      const cplx x = OWN ? v[e] : xk[e * S];
      const cplx xp = cconj(xm[-e * S]);
      const cplx w = cconj(__ldg(twp + e * Pl::T));               // exp(+2 pi i k / n)
      const cplx s = cadd(x, xp), d = csub(x, xp);
      const cplx wd = cmul_i(cmul(w, d));
      v[e] = cadd_scaled(s, wd, scale);
    }
    return;
  }
#pragma unroll
  for (int e = 0; e < Pl::E; ++e) {
    const int k = t + e * Pl::T;
    const cplx x = OWN ? v[e] : xch[pad16(k)];
    const cplx xp = cconj(xch[pad16(N - k)]);
    const cplx w = cconj(__ldg(&tw[twiddle_offset(LOG2N) + k]));   // exp(+2 pi i k / n)
    const cplx s = cadd(x, xp), d = csub(x, xp);
    const cplx wd = cmul_i(cmul(w, d));
    v[e] = cadd_scaled(s, wd, scale);
  }
}

// sum over the T threads of a group; scratch: >= 32 floats per group; uses __syncthreads when T > 32
template <int LOG2N>
__device__ __forceinline__ float group_sum(float val, float* scratch, int t) {
  constexpr int T = FftPlan<LOG2N>::T;
  constexpr int W = T < 32 ? T : 32;
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
  if constexpr (T > 32) {
    group_sync<LOG2N>();
    if ((t & 31) == 0) scratch[t >> 5] = val;
    group_sync<LOG2N>();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < T / 32; ++i) s += scratch[i];
    val = s;
  }
  return val;
}

}  // namespace cvb
