#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
// throughput microbenchmark: scalar FFMA vs packed FFMA2 / FADD2 / FMUL2 (sm_100a)
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float c0) {
  float2 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  float2 b = make_float2(c0, c0 * 0.5f), c = make_float2(1e-3f, -1e-3f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
        if (MODE == 1) a[i] = __ffma2_rn(a[i], b, c);
        if (MODE == 2) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }
        if (MODE == 3) a[i] = __fadd2_rn(a[i], c);
        if (MODE == 4) { a[i].x = a[i].x * b.x; a[i].y = a[i].y * b.y; }
        if (MODE == 5) a[i] = __fmul2_rn(a[i], b);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, float* out, int iters) {
  int grid = 148 * 8;
  k<MODE><<<grid, 256>>>(out, 10, 1.0001f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(out, iters, 1.0001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double lane_ops = (double)grid * 256 * iters * 4 * 8 * 2;   // fp32 lane operations (each fma/add/mul on one float)
  printf("%-8s %.3f ms  %.2f T lane-ops/s  (%.1f per SM per clk at 1.965 GHz)\n", name, ms, lane_ops / ms / 1e9,
         lane_ops / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  int iters = 20000;
  run<0>("FFMA", out, iters); run<1>("FFMA2", out, iters); run<2>("FADD", out, iters); run<3>("FADD2", out, iters);
  run<4>("FMUL", out, iters); run<5>("FMUL2", out, iters);
  cudaError_t e = cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(e));
  return 0;
}
