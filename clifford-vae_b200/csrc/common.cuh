// Shared device/host helpers for the clifford_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>

#ifndef __CUDA_ARCH__
#define CVB_HOST_PASS 1
#endif

namespace cvb {

// ---- status codes returned through the C ABI (include/clifford_b200.h) -------------------
enum Status : int {
  kOk = 0,
  kBadArgument = 1,      // null pointer, non-positive size, misaligned row
  kUnsupported = 2,      // shape outside what the kernels implement
  kCudaError = 3,        // a CUDA runtime call failed (see cvb_last_error_string)
};

typedef float2 cplx;

// Complex arithmetic on (re, im) pairs with sm_100's packed fp32 instructions (FADD2 / FMUL2 / FFMA2).  A complex
// add is ONE instruction and a complex multiply TWO: ptxas folds the lane swap, the per-lane sign and the scalar
// broadcast of the expressions below into operand modifiers (R.F32x2.LO_HI, .NP, R.F32), so no MOV is emitted.  The
// fp32 pipes retire the same number of lane operations as the scalar forms (measured, tools/microbench/
// packed_fp32_throughput.cu: 124 lane-ops / SM / clk either way) -- what halves is the issue-slot count.
// Rounding is identical to the scalar fmaf forms these replace.
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  // (a.x b.x - a.y b.y, a.x b.y + a.y b.x) = b * a.x + swap(b) * (-a.y, a.y)
  return __ffma2_rn(b, make_float2(a.x, a.x), __fmul2_rn(make_float2(b.y, b.x), make_float2(-a.y, a.y)));
}
// a * conj(b) = a * b.x + swap(a) * (b.y, -b.y)
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {
  return __ffma2_rn(a, make_float2(b.x, b.x), __fmul2_rn(make_float2(a.y, a.x), make_float2(b.y, -b.y)));
}
__device__ __forceinline__ cplx cconj(cplx a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ cplx cscale(cplx a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
// s * (a + b), s * a + b
__device__ __forceinline__ cplx cadd_scaled(cplx a, cplx b, float s) { return __fmul2_rn(__fadd2_rn(a, b), make_float2(s, s)); }
__device__ __forceinline__ cplx caxpy(float s, cplx a, cplx b) { return __ffma2_rn(a, make_float2(s, s), b); }
// multiply by +i / -i (a lane swap with one sign: folded into the consumer's operand modifiers)
__device__ __forceinline__ cplx cmul_i(cplx a) { return make_float2(-a.y, a.x); }
__device__ __forceinline__ cplx cmul_mi(cplx a) { return make_float2(a.y, -a.x); }
__device__ __forceinline__ float cabs2(cplx a) { return fmaf(a.x, a.x, a.y * a.y); }

// Shared-memory exchange buffers hold complex values as float2 with one pad slot every 16
// entries: stride-R (R = 2..16) writes of the first Stockham pass and the contiguous reads of
// the next pass are then both conflict-free for 64-bit accesses (DESIGN.md "smem layout").
__host__ __device__ __forceinline__ constexpr int pad16(int i) { return i + (i >> 4); }

// Twiddle tables, one compact table per transform size so that consecutive lanes read consecutive
// entries: for a complex length N = 2^L (real length n = 2N) the table exp(-2*pi*i*m / n), m in [0, N),
// starts at entry 2^L of one concatenated per-device array (L = 4 .. 13, 16384 entries, 128 KB).
// Filled once per device in double precision on the host (api_core.cu: cvb_init()).
constexpr int kTwiddleMaxLog2N = 13;
constexpr int kTwiddleEntries = 1 << (kTwiddleMaxLog2N + 1);
__host__ __device__ __forceinline__ constexpr int twiddle_offset(int log2n) { return 1 << log2n; }

// streaming (read-once / write-once) global accesses: keep them out of L1
__device__ __forceinline__ float2 ldg_stream2(const float2* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream2(float2* p, float2 v) {
  asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream1(float* p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// Programmatic dependent launch: wait for the preceding kernel on the stream (no-op when launched without the attribute),
// then let the NEXT kernel start becoming resident as this one's CTAs retire.
__device__ __forceinline__ void pdl_wait_and_release() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// single-instruction MUFU.SQRT (max relative error 2^-23); sqrtf() expands to a refinement sequence
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Concentration head folded into the samplers (reference mnist/mlp_vae.py:69-71, cnn/models.py:96,99):
//     kappa = min(softplus(raw) + floor, kmax)
// evaluated by the kernel from the raw output of the `fc_scale` / `fc_concentration` layer, with the chain factor
// d kappa / d raw applied to every kappa-gradient the kernel writes.  on == 0: the input already is kappa.
// softplus as torch (beta 1, threshold 20: softplus(x) = x for x > 20); clamp(max=) passes the gradient where x <= max.
struct KappaHead {
  int on;
  float floor, kmax;
};
__device__ __forceinline__ float head_kappa(const KappaHead& h, float raw) {
  if (!h.on) return raw;
  const float sp = raw > 20.0f ? raw : log1pf(expf(raw));
  return fminf(sp + h.floor, h.kmax);
}
__device__ __forceinline__ float head_dkappa(const KappaHead& h, float raw) {
  if (!h.on) return 1.0f;
  const float e = expf(raw);
  const float sp = raw > 20.0f ? raw : log1pf(e);
  if (!(sp + h.floor <= h.kmax)) return 0.0f;
  return raw > 20.0f ? 1.0f : e / (e + 1.0f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace cvb
