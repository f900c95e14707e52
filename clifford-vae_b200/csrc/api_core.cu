// Library state: error string, per-device twiddle table, launch counter, occupancy cache, and the
// RNG test hook.  C ABI declared in include/clifford_b200.h.
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>
#include "launch.cuh"
#include "icdf_table.cuh"
#include "../../include/clifford_b200.h"

namespace cvb {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launch_count{0};
static std::mutex g_mu;
static const cplx* g_tw[64] = {nullptr};
static int g_sms[64] = {0};
static int* g_sched[64] = {nullptr};
static std::atomic<unsigned> g_sched_seq{0};
constexpr int kSchedSlots = 64;
// The dynamic-smem opt-in (cudaFuncSetAttribute) is per (device, kernel) and must only ever be RAISED (kernels with a
// run-time shared-memory size -- the direct-DFT and short-row kernels -- are launched with different sizes); the
// occupancy answer additionally depends on the launch's threads and shared memory.
static std::unordered_map<unsigned long long, size_t> g_smem_optin;   // (device, kernel) -> largest opt-in applied
static std::unordered_map<unsigned long long, int> g_occ;             // (device, kernel, threads, smem) -> resident CTAs per SM
static unsigned long long func_key(const void* kernel) {
  int dev = 0;
  cudaGetDevice(&dev);
  return (unsigned long long)reinterpret_cast<uintptr_t>(kernel) * 64ull + (unsigned long long)(dev & 63);
}
static unsigned long long occ_key(const void* kernel, int threads, size_t smem) {
  unsigned long long h = func_key(kernel);
  h ^= ((unsigned long long)smem + 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
  h ^= ((unsigned long long)threads + 0x94D049BB133111EBull) * 0xD6E8FEB86659FD93ull;
  return h;
}
static const float2* g_icdf[64] = {nullptr};
static const unsigned long long* g_rng_ctr[64] = {nullptr};
static unsigned int* g_rng_arrive[64] = {nullptr};      // per-device arrival word (allocated by cvb_init, self re-arming)
static bool g_rng_autobump[64] = {false};
static std::vector<float2> g_icdf_host;

// ---- host construction of the half-angle inverse-CDF table (icdf_table.cuh), double precision ----------------
// regularised incomplete beta I_x(a, b) by the modified Lentz continued fraction; xc = 1 - x passed separately so
// that both ends keep their relative accuracy
static double beta_cf(double a, double b, double x) {
  const double tiny = 1e-300, eps = 1e-16;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < tiny) d = tiny;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 2000; ++m) {
    const int m2 = 2 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d; h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < eps) break;
  }
  return h;
}
// (I_x(a, b), 1 - I_x(a, b)) with x + xc = 1
static void beta_inc(double a, double b, double x, double xc, double* lower, double* upper) {
  if (x <= 0.0) { *lower = 0.0; *upper = 1.0; return; }
  if (xc <= 0.0) { *lower = 1.0; *upper = 0.0; return; }
  const double lbt = lgamma(a + b) - lgamma(a) - lgamma(b) + a * log(x) + b * log(xc);
  if (x < (a + 1.0) / (a + b + 2.0)) {
    *lower = exp(lbt) * beta_cf(a, b, x) / a;
    *upper = 1.0 - *lower;
  } else {
    *upper = exp(lbt) * beta_cf(b, a, xc) / b;
    *lower = 1.0 - *upper;
  }
}

static void build_icdf_host(std::vector<float2>& out) {
  const double half_pi = 1.57079632679489661923;
  out.assign(kIcdfTableEntries, make_float2(0.f, 0.f));
  for (int ki = 0; ki < kIcdfKappaNodes; ++ki) {
    const double kap = expm1((double)kIcdfQMax * ki / (kIcdfKappaNodes - 1));
    const double a = kap + 0.5, p = 2.0 * kap + 1.0;
    const double Z = exp(log(0.5 * sqrt(3.14159265358979323846)) + lgamma(kap + 0.5) - lgamma(kap + 1.0));   // int_0^{pi/2} cos^{2k}
    float2* row = out.data() + (size_t)ki * kIcdfRowStride;
    double prev = half_pi;
    for (int j = 0; j <= kIcdfCells; ++j) {
      const double s = (double)j / kIcdfCells;
      double psi, dpsi;
      if (j == 0) {
        psi = half_pi; dpsi = -pow(p * Z, 1.0 / p);
      } else if (j == kIcdfCells) {
        psi = 0.0; dpsi = -p * Z;
      } else {
        // solve tail(psi) = P(|psi'| > psi) = I_{cos^2 psi}(k + 1/2, 1/2) = s^p: safeguarded Newton on log tail
        const double lv = p * log(s);
        double lo = 0.0, hi = prev;                     // H is decreasing in s
        psi = 0.5 * (lo + hi);
        for (int it = 0; it < 200; ++it) {
          const double cs = cos(psi), sn = sin(psi);
          double low, up;
          beta_inc(a, 0.5, cs * cs, sn * sn, &low, &up);
          const double f = log(low) - lv;               // decreasing in psi
          if (f > 0.0) lo = psi; else hi = psi;
          if (hi - lo < 1e-15 * half_pi) break;
          // d log(tail) / d psi = -g / tail, g = cos^{2k} / Z
          const double g = exp(2.0 * kap * log(fmax(cs, 1e-300))) / Z;
          double nxt = psi + f * low / g;
          if (!(nxt > lo && nxt < hi)) nxt = 0.5 * (lo + hi);
          if (fabs(nxt - psi) < 1e-15) { psi = nxt; break; }
          psi = nxt;
        }
        const double g = exp(2.0 * kap * log(cos(psi))) / Z;
        dpsi = -p * exp((p - 1.0) * log(s)) / g;
        prev = psi;
      }
      row[j] = make_float2((float)psi, (float)(dpsi / kIcdfCells));
    }
    row[kIcdfCells + 1] = row[kIcdfCells];          // padding entry (never interpolated)
  }
}

const float2* device_icdf_table() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || !g_icdf[dev]) {
    set_last_error("cvb_init() has not been called on the current device");
    return nullptr;
  }
  return g_icdf[dev];
}

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const cplx* device_twiddles() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || !g_tw[dev]) {
    set_last_error("cvb_init() has not been called on the current device");
    return nullptr;
  }
  return g_tw[dev];
}

const unsigned long long* rng_device_counter() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  return g_rng_ctr[dev];
}

unsigned int* rng_arrive_word() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || !g_rng_autobump[dev]) return nullptr;
  return g_rng_arrive[dev];
}

int* next_sched_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || !g_sched[dev]) return nullptr;
  return g_sched[dev] + 2 * (g_sched_seq.fetch_add(1, std::memory_order_relaxed) % kSchedSlots);
}

int sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && g_sms[dev] > 0) return g_sms[dev];
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (dev >= 0 && dev < 64) g_sms[dev] = n;
  return n;
}

int cached_ctas_per_sm(const void* kernel, int threads, size_t smem, bool* found) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_occ.find(occ_key(kernel, threads, smem));
  *found = it != g_occ.end();
  return *found ? it->second : 0;
}
void store_ctas_per_sm(const void* kernel, int threads, size_t smem, int per_sm) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_occ[occ_key(kernel, threads, smem)] = per_sm;
}
int ensure_smem_optin(const void* kernel, size_t smem) {
  if (smem <= 48 * 1024) return kOk;
  std::lock_guard<std::mutex> lk(g_mu);
  size_t& cur = g_smem_optin[func_key(kernel)];
  if (smem > cur) {
    CVB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cur = smem;
  }
  return kOk;
}

__global__ void philox_fill_kernel(uint4* out, long long n, PhiloxKey key) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = philox4x32_10_rk(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 0u, key.offset), key);   // the kernels' form (round keys from the host)
}

}  // namespace cvb

using namespace cvb;

extern "C" {

int cvb_version(void) { return 100; }
const char* cvb_last_error_string(void) { return g_err; }
long long cvb_launch_count(void) { return g_launch_count.load(); }

int cvb_init(void) {
  int dev = 0;
  CVB_CUDA(cudaGetDevice(&dev));
  CVB_REQUIRE(dev >= 0 && dev < 64, kUnsupported, "device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_tw[dev]) return kOk;
  cudaDeviceProp prop;
  CVB_CUDA(cudaGetDeviceProperties(&prop, dev));
  CVB_REQUIRE(prop.major == 10, kUnsupported,
              "clifford_b200 is built for sm_100a only; device %d is sm_%d%d", dev, prop.major, prop.minor);
  std::vector<float2> host(kTwiddleEntries, make_float2(1.f, 0.f));
  for (int L = 4; L <= kTwiddleMaxLog2N; ++L) {
    const int N = 1 << L;
    for (int m = 0; m < N; ++m) {
      const double ang = -2.0 * 3.14159265358979323846264338327950288 * (double)m / (double)(2 * N);
      host[twiddle_offset(L) + m] = make_float2((float)cos(ang), (float)sin(ang));
    }
  }
  cplx* dptr = nullptr;
  CVB_CUDA(cudaMalloc(&dptr, sizeof(cplx) * kTwiddleEntries));
  CVB_CUDA(cudaMemcpy(dptr, host.data(), sizeof(cplx) * kTwiddleEntries, cudaMemcpyHostToDevice));
  int* sched = nullptr;
  CVB_CUDA(cudaMalloc(&sched, sizeof(int) * 2 * kSchedSlots));
  CVB_CUDA(cudaMemset(sched, 0, sizeof(int) * 2 * kSchedSlots));
  if (g_icdf_host.empty()) build_icdf_host(g_icdf_host);
  float2* icdf = nullptr;
  CVB_CUDA(cudaMalloc(&icdf, sizeof(float2) * kIcdfTableEntries));
  {
    std::vector<float2> dev_copy(g_icdf_host);      // the device samples the phase phi = 2 psi directly
    for (auto& e : dev_copy) { e.x *= 2.0f; e.y *= 2.0f; }
    CVB_CUDA(cudaMemcpy(icdf, dev_copy.data(), sizeof(float2) * kIcdfTableEntries, cudaMemcpyHostToDevice));
  }
  g_icdf[dev] = icdf;
  g_sched[dev] = sched;
  unsigned int* arrive = nullptr;
  CVB_CUDA(cudaMalloc(&arrive, sizeof(unsigned int)));
  CVB_CUDA(cudaMemset(arrive, 0, sizeof(unsigned int)));
  g_rng_arrive[dev] = arrive;
  g_tw[dev] = dptr;
  g_sms[dev] = prop.multiProcessorCount;
  return kOk;
}

// Register (or clear, with NULL) a device-resident 64-bit launch counter for the CURRENT device: every sampling kernel
// launched afterwards adds it to its Philox offset.  The library only reads it; the caller increments it on the stream
// (e.g. a captured `counter += 1`) after each sampling launch, which is what makes the samplers replayable inside a
// CUDA graph with fresh draws per replay.
int cvb_set_rng_device_counter(const unsigned long long* counter) {
  int dev = 0;
  CVB_CUDA(cudaGetDevice(&dev));
  CVB_REQUIRE(dev >= 0 && dev < 64, kUnsupported, "device ordinal %d out of range", dev);
  g_rng_ctr[dev] = counter;
  g_rng_autobump[dev] = false;
  return kOk;
}

// The same registration in self-bumping mode: every launch that draws from the device generator adds 1 to *counter when
// its last CTA retires (rng_launch_done), so a captured graph needs no `counter += 1` kernel of its own.  Launches in this
// mode must be serialised on one stream (one arrival word per device).  Backward launches that REPLAY draws (the sphere
// samplers regenerate their tangent normals) must see the forward's counter value: snapshot it before the forward and
// register the snapshot with cvb_set_rng_device_counter around the backward launch (clifford_b200/ops.py does).
int cvb_set_rng_device_counter_autobump(unsigned long long* counter) {
  int dev = 0;
  CVB_CUDA(cudaGetDevice(&dev));
  CVB_REQUIRE(dev >= 0 && dev < 64, kUnsupported, "device ordinal %d out of range", dev);
  CVB_REQUIRE(counter == nullptr || g_rng_arrive[dev] != nullptr, kBadArgument, "cvb_set_rng_device_counter_autobump: call cvb_init() first");
  g_rng_ctr[dev] = counter;
  g_rng_autobump[dev] = counter != nullptr;
  return kOk;
}

// Host-only: the half-angle inverse-CDF table the device samplers interpolate (icdf_table.cuh); no GPU needed.
int cvb_ps_halfangle_icdf_table(float* out, long long capacity_floats, int* n_kappa, int* n_nodes, float* kappa_max) {
  if (n_kappa) *n_kappa = kIcdfKappaNodes;
  if (n_nodes) *n_nodes = kIcdfCells + 1;
  if (kappa_max) *kappa_max = kIcdfKappaMax;
  if (!out) return kOk;
  const long long need = 2LL * kIcdfKappaNodes * (kIcdfCells + 1);
  CVB_REQUIRE(capacity_floats >= need, kBadArgument, "cvb_ps_halfangle_icdf_table: need room for %lld floats", need);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_icdf_host.empty()) build_icdf_host(g_icdf_host);
  }
  for (int ki = 0; ki < kIcdfKappaNodes; ++ki)       // drop the row padding
    memcpy(out + 2LL * ki * (kIcdfCells + 1), g_icdf_host.data() + (size_t)ki * kIcdfRowStride, sizeof(float2) * (kIcdfCells + 1));
  return kOk;
}

int cvb_philox_fill(unsigned int* out, long long n_vec4, unsigned long long seed, unsigned long long offset,
                    void* stream) {
  CVB_REQUIRE(out && n_vec4 > 0, kBadArgument, "cvb_philox_fill: bad arguments");
  PhiloxKey key = make_key(seed, 0, 0);
  key.offset = (uint32_t)offset;
  key.dev_counter = nullptr;          // the known-answer test hook is addressed by (seed, offset) alone
  key.arrive = nullptr;
  philox_fill_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint4*>(out), n_vec4, key);
  return check_launch("philox_fill_kernel");
}

}  // extern "C"
