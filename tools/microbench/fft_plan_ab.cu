// On-chip cost of one N = 2048 / 4096 complex FFT under different plans of csrc/fft_core.cuh (points per thread E,
// threads per transform T = N / E): the production 16-point plan (3 passes, 2 shared-memory exchanges) against the
// 32- and 64-point plans (64 points: 2 passes, ONE exchange, one warp per 2048-point transform).  Data stays resident:
// each group loads a row once and chains `reps` forward transforms.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I ../../clifford-vae_b200/csrc fft_plan_ab.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cmath>
#include <vector>
#include "fft_core.cuh"
using namespace cvb;

template <class Pl>
__global__ void __launch_bounds__(Pl::THREADS) fft_chain(const float2* __restrict__ in, float2* __restrict__ out, int rows, int reps,
                                                         const cplx* __restrict__ tw) {
  constexpr int N = Pl::N, T = Pl::T, E = Pl::E, G = Pl::GROUPS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int group = threadIdx.x / T, t = threadIdx.x % T;
  cplx* xch = reinterpret_cast<cplx*>(smem_raw) + (size_t)group * Pl::XCH;
  for (long long row = (long long)blockIdx.x * G + group; row < rows; row += (long long)gridDim.x * G) {
    cplx v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = in[row * N + t + e * T];
    for (int r = 0; r < reps; ++r) {
      fft_run_p<Pl, false>(v, xch, t, tw);
      if (r + 1 < reps) {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = cscale(v[e], 1.0f / 32.0f);      // keep magnitudes bounded over the chain
      }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) out[row * N + t + e * T] = v[e];
  }
}

template <class Pl>
float run(const char* name, const float2* in, float2* out, const cplx* tw, int reps, int* ctas_per_sm_out) {
  const size_t smem = sizeof(cplx) * Pl::XCH * Pl::GROUPS;
  auto kern = fft_chain<Pl>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Pl::THREADS, smem);
  const int grid = 148 * per_sm, rows = grid * Pl::GROUPS * 4;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  kern<<<grid, Pl::THREADS, smem>>>(in, out, rows, reps, tw);
  cudaEventRecord(a);
  kern<<<grid, Pl::THREADS, smem>>>(in, out, rows, reps, tw);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
  const double cyc = ms * 1e-3 * 1.965e9 * 148 / ((double)rows * reps);
  printf("%-34s E=%2d T=%3d threads/CTA=%3d regs=%3d smem/CTA=%6zu CTAs/SM=%2d (rows in flight/SM %2d): %8.1f us  -> %6.0f cycles per FFT per SM\n",
         name, Pl::E, Pl::T, Pl::THREADS, fa.numRegs, smem, per_sm, per_sm * Pl::GROUPS, ms * 1e3, cyc);
  *ctas_per_sm_out = per_sm;
  return ms;
}

template <class PlA, class PlB>
void check(const float2* d_in, float2* d_o1, float2* d_o2, const cplx* tw, int rows) {
  constexpr int N = PlA::N;
  fft_chain<PlA><<<rows / PlA::GROUPS, PlA::THREADS, sizeof(cplx) * PlA::XCH * PlA::GROUPS>>>(d_in, d_o1, rows, 1, tw);
  fft_chain<PlB><<<rows / PlB::GROUPS, PlB::THREADS, sizeof(cplx) * PlB::XCH * PlB::GROUPS>>>(d_in, d_o2, rows, 1, tw);
  cudaDeviceSynchronize();
  std::vector<float2> a((size_t)rows * N), b((size_t)rows * N), x((size_t)rows * N);
  cudaMemcpy(a.data(), d_o1, a.size() * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(b.data(), d_o2, b.size() * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(x.data(), d_in, x.size() * 8, cudaMemcpyDeviceToHost);
  double mx = 0, dab = 0, ea = 0, eb = 0;
  // fp64 DFT of row 0 only (N^2)
  for (int k = 0; k < N; ++k) {
    double yr = 0, yi = 0;
    for (int j = 0; j < N; ++j) {
      const double th = -2.0 * M_PI * (double)((long long)j * k % N) / N;
      yr += x[j].x * cos(th) - x[j].y * sin(th); yi += x[j].x * sin(th) + x[j].y * cos(th);
    }
    mx = fmax(mx, fmax(fabs(yr), fabs(yi)));
    ea = fmax(ea, fmax(fabs(a[k].x - yr), fabs(a[k].y - yi)));
    eb = fmax(eb, fmax(fabs(b[k].x - yr), fabs(b[k].y - yi)));
  }
  for (size_t i = 0; i < a.size(); ++i) dab = fmax(dab, fmax(fabs(a[i].x - b[i].x), fabs(a[i].y - b[i].y)));
  printf("N=%d: max-norm relative error vs fp64 DFT (row 0): 16-point plan %.2e, wide plan %.2e; plans differ by %.2e (all rows)\n",
         N, ea / mx, eb / mx, dab / mx);
}

int main() {
  std::vector<float2> host(kTwiddleEntries, make_float2(1.f, 0.f));
  for (int L = 4; L <= kTwiddleMaxLog2N; ++L) {
    const int N = 1 << L;
    for (int m = 0; m < N; ++m) {
      const double ang = -2.0 * M_PI * (double)m / (double)(2 * N);
      host[twiddle_offset(L) + m] = make_float2((float)cos(ang), (float)sin(ang));
    }
  }
  cplx* tw; cudaMalloc(&tw, sizeof(cplx) * kTwiddleEntries);
  cudaMemcpy(tw, host.data(), sizeof(cplx) * kTwiddleEntries, cudaMemcpyHostToDevice);
  const size_t n = (size_t)148 * 16 * 4 * 8 * 4096;
  std::vector<float2> h(1 << 22);
  srand(3);
  for (auto& v : h) v = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
  float2 *d_in, *d_o1, *d_o2;
  cudaMalloc(&d_in, n * 8); cudaMalloc(&d_o1, n * 8); cudaMalloc(&d_o2, n * 8);
  for (size_t off = 0; off < n; off += h.size()) cudaMemcpy(d_in + off, h.data(), std::min(h.size(), n - off) * 8, cudaMemcpyHostToDevice);
  check<FftPlan<11>, FftPlanT<11, 6, 32>>(d_in, d_o1, d_o2, tw, 128);
  check<FftPlan<12>, FftPlanT<12, 6, 64>>(d_in, d_o1, d_o2, tw, 128);
  const int reps = 32;
  int c;
  printf("N = 2048 (bind d = 4096, Clifford d = 2048)\n");
  run<FftPlan<11>>("16 pt: 16.16.8, 2 exchanges", d_in, d_o1, tw, reps, &c);
  run<FftPlanT<11, 5, 64>>("32 pt: 32.32.2, 2 exchanges", d_in, d_o1, tw, reps, &c);
  run<FftPlanT<11, 5, 128>>("32 pt, 2 transforms per CTA", d_in, d_o1, tw, reps, &c);
  run<FftPlanT<11, 6, 32>>("64 pt: 64.32, 1 exchange", d_in, d_o1, tw, reps, &c);
  run<FftPlanT<11, 6, 64>>("64 pt, 2 transforms per CTA", d_in, d_o1, tw, reps, &c);
  run<FftPlanT<11, 6, 128>>("64 pt, 4 transforms per CTA", d_in, d_o1, tw, reps, &c);
  printf("N = 4096 (bind d = 8192)\n");
  run<FftPlan<12>>("16 pt: 16.16.16, 2 exchanges", d_in, d_o1, tw, reps, &c);
  run<FftPlanT<12, 5, 128>>("32 pt: 32.32.4, 2 exchanges", d_in, d_o1, tw, reps, &c);
  run<FftPlanT<12, 6, 64>>("64 pt: 64.64, 1 exchange", d_in, d_o1, tw, reps, &c);
  run<FftPlanT<12, 6, 128>>("64 pt, 2 transforms per CTA", d_in, d_o1, tw, reps, &c);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
