// Library state: error string, per-device twiddle table, launch counter, occupancy cache, and the
// RNG test hook.  C ABI declared in include/clifford_b200.h.
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>
#include "launch.cuh"
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
static std::unordered_map<const void*, int> g_occ;

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

int cached_ctas_per_sm(const void* kernel, int, size_t, bool* found) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_occ.find(kernel);
  *found = it != g_occ.end();
  return *found ? it->second : 0;
}
void store_ctas_per_sm(const void* kernel, int per_sm) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_occ[kernel] = per_sm;
}

__global__ void philox_fill_kernel(uint4* out, long long n, PhiloxKey key) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 0u, key.offset), key.k0, key.k1);
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
  g_sched[dev] = sched;
  g_tw[dev] = dptr;
  g_sms[dev] = prop.multiProcessorCount;
  return kOk;
}

int cvb_philox_fill(unsigned int* out, long long n_vec4, unsigned long long seed, unsigned long long offset,
                    void* stream) {
  CVB_REQUIRE(out && n_vec4 > 0, kBadArgument, "cvb_philox_fill: bad arguments");
  PhiloxKey key = make_key(seed, 0, 0);
  key.offset = (uint32_t)offset;
  philox_fill_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint4*>(out), n_vec4, key);
  return check_launch("philox_fill_kernel");
}

}  // extern "C"
