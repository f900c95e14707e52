// DFT-as-GEMM, measured: ONE radix-16 pass of the N = 2048 complex FFT (128 butterflies of 16
// points = one latent row at d = 2048) done two ways on the same register / shared-memory data flow:
//   (A) CUDA cores : the production in-register radix-16 butterfly (packed fp32, csrc/fft_core.cuh Dft<16>) followed by
//                    the Stockham exchange through shared memory (16 STS.64 + barrier + 16 LDS.64 per thread);
//   (B) tcgen05    : the same 128 x (32 x 32 real) products as one M=128, N=32, K=32 GEMM on the 5th-gen tensor cores
//                    with 3xTF32 error compensation (A = A_hi + A_lo, B = B_hi + B_lo; A_hi B_hi + A_lo B_hi + A_hi B_lo),
//                    A written by the threads in the canonical K-major 128B-swizzled layout (this IS the inter-pass
//                    exchange), B (the DFT matrix) resident in shared memory, D in TMEM, read back with tcgen05.ld.
// Both variants chain R passes on resident data (no HBM traffic inside the timed loop), so the number reported is the
// on-chip cost of a pass: cycles per 128-butterfly tile per SM at the occupancy each variant reaches.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../clifford-vae_b200/csrc dft16_tcgen05.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "fft_core.cuh"

using namespace cvb;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------------------------------------- (A) CUDA cores
__global__ void __launch_bounds__(128) dft16_cuda_core(const float2* __restrict__ in, float2* __restrict__ out, int tiles, int passes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* xch = reinterpret_cast<cplx*>(smem_raw);
  const int t = threadIdx.x;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    cplx u[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) u[e] = in[(size_t)tile * 2048 + t + e * 128];
    for (int ps = 0; ps < passes; ++ps) {
      Dft<16, false>::run(u);
      if (ps + 1 < passes) {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) xch[pad16(16 * t + r)] = u[r];      // Stockham write of pass 0 (stride-16 scatter)
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 16; ++e) u[e] = xch[pad16(t + e * 128)];      // next pass's ownership pattern
      }
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) out[(size_t)tile * 2048 + 16 * t + e] = u[e];
  }
}

// ---------------------------------------------------------------------------------------------- (B) tcgen05
// shared memory: A_hi (16 KB) | A_lo (16 KB) | B_hi (4 KB) | B_lo (4 KB) | mbarrier | tmem base     (1024-byte aligned tiles)
constexpr int kAOff = 0, kALoOff = 16384, kBOff = 32768, kBLoOff = 36864, kBarOff = 40960, kSmemB = 40960 + 64;

// Operand row (= TMEM lane = thread) that holds butterfly b: rows 8r .. 8r+7 hold butterflies r, 16+r, ..., 112+r, so the
// Stockham scatter of one output index r (a warp writes butterflies 16 j + r) lands in ONE 8-row swizzle atom instead
// of eight rows with identical bank alignment (that version, 8-way conflicted, measured 2560 cycles per tile-pass).
__device__ __forceinline__ int row_of(int b) { return ((b & 15) << 3) | (b >> 4); }
__device__ __forceinline__ int butterfly_of(int row) { return ((row & 7) << 4) | (row >> 3); }

__device__ __forceinline__ uint32_t sw128(int row, int chunk) {        // byte offset of 16-byte chunk `chunk` of row `row`
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  // K-major, SWIZZLE_128B: start address >> 4 | LBO (unused for swizzled K-major) | SBO = 1024 B between 8-row groups |
  // version 1 (sm_100) at bit 46 | layout type 2 (SWIZZLE_128B) at bit 61
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                       // leading byte offset (16 B, ignored)
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset
  d |= (uint64_t)1 << 46;                       // descriptor version
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
constexpr uint32_t kIdesc = (1u << 4)      // D format F32
                          | (2u << 7)      // A format TF32
                          | (2u << 10)     // B format TF32
                          | (0u << 15) | (0u << 16)   // A, B K-major
                          | ((32u >> 3) << 17)        // N = 32
                          | ((128u >> 4) << 24);      // M = 128

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128) dft16_tcgen05(const float2* __restrict__ in, float2* __restrict__ out, int tiles, int passes,
                                                     const float* __restrict__ bmat /* [2][32][32]: hi, lo; row n, col k */) {
  extern __shared__ __align__(16) unsigned char smem_unaligned[];
  // the 128B-swizzled operand tiles must start on a 1024-byte boundary of the shared window
  unsigned char* smem = smem_unaligned + ((1024u - (smem_addr(smem_unaligned) & 1023u)) & 1023u);
  const int t = threadIdx.x, warp = t >> 5;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kBarOff + 16);
  // B tiles: row n = 0..31 (output 2f + c'), 32 floats (k = 2e + c) -> 8 chunks of 16 B
  for (int i = t; i < 2 * 32 * 8; i += 128) {
    const int which = i / 256, n = (i / 8) % 32, c = i % 8;
    const float4 v = reinterpret_cast<const float4*>(bmat)[(which * 32 + n) * 8 + c];
    *reinterpret_cast<float4*>(smem + (which ? kBLoOff : kBOff) + sw128(n, c)) = v;
  }
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_addr(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;
  const uint64_t dA = make_desc(smem_addr(smem + kAOff)), dAl = make_desc(smem_addr(smem + kALoOff));
  const uint64_t dB = make_desc(smem_addr(smem + kBOff)), dBl = make_desc(smem_addr(smem + kBLoOff));
  uint32_t parity = 0;

  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    float u[32];                       // this thread's butterfly: u[2e] = Re x_e, u[2e+1] = Im x_e
    const int bfly = butterfly_of(t);  // thread t = operand row t = TMEM lane t works on butterfly bfly in every pass
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const float2 v = in[(size_t)tile * 2048 + bfly + e * 128];
      u[2 * e] = v.x; u[2 * e + 1] = v.y;
    }
    for (int ps = 0; ps < passes; ++ps) {
      // Pass 0: the thread writes its own row.  Later passes: the Stockham scatter IS the operand write -- output r of
      // butterfly b becomes point 16 b + r = element (16 b + r) >> 7 of butterfly (16 b + r) & 127 (one 8-byte store per
      // output, hi and lo, instead of the 16-byte row chunks of pass 0).
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float x = u[4 * c + j];
          hi[j] = x;                                                     // the tensor core reads the top 19 bits
          lo[j] = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);  // exact remainder
        }
        if (ps == 0) {
          *reinterpret_cast<float4*>(smem + kAOff + sw128(t, c)) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<float4*>(smem + kALoOff + sw128(t, c)) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {                                  // complex outputs r = 2c, 2c + 1
            const int p = 16 * bfly + 2 * c + h, row = row_of(p & 127), e = p >> 7;
            const uint32_t off = sw128(row, e >> 1) + (e & 1) * 8;
            *reinterpret_cast<float2*>(smem + kAOff + off) = make_float2(hi[2 * h], hi[2 * h + 1]);
            *reinterpret_cast<float2*>(smem + kALoOff + off) = make_float2(lo[2 * h], lo[2 * h + 1]);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (t == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_tf32(tmem_d, dA + 2 * k, dB + 2 * k, k > 0);       // +32 bytes per K = 8 step
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_tf32(tmem_d, dAl + 2 * k, dB + 2 * k, 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_tf32(tmem_d, dA + 2 * k, dBl + 2 * k, 1);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
      }
      // wait for the MMAs
      asm volatile(
          "{\n.reg .pred P1;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_addr(bar)),
          "r"(parity)
          : "memory");
      parity ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[32];
      const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
            "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
            "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) u[i] = __uint_as_float(r[i]);
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) out[(size_t)tile * 2048 + 16 * bfly + e] = make_float2(u[2 * e], u[2 * e + 1]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem_d) : "memory");
}

// ---------------------------------------------------------------------------------------------- host
static float tf32_trunc(float x) { uint32_t b; memcpy(&b, &x, 4); b &= 0xFFFFE000u; memcpy(&x, &b, 4); return x; }

int main(int argc, char** argv) {
  const int tiles_check = 296, passes_bench = 24;
  const int sms = 148, tiles_max = 148 * 6 * 4;
  // DFT matrix, split hi / lo
  std::vector<float> bm(2 * 32 * 32);
  for (int f = 0; f < 16; ++f)
    for (int e = 0; e < 16; ++e) {
      const double th = 2.0 * M_PI * e * f / 16.0;
      const double v[2][2] = {{cos(th), sin(th)}, {-sin(th), cos(th)}};     // [c'][c]
      for (int cp = 0; cp < 2; ++cp)
        for (int c = 0; c < 2; ++c) {
          const float full = (float)v[cp][c], hi = tf32_trunc(full), lo = (float)(v[cp][c] - (double)hi);
          bm[(0 * 32 + 2 * f + cp) * 32 + 2 * e + c] = hi;
          bm[(1 * 32 + 2 * f + cp) * 32 + 2 * e + c] = lo;
        }
    }
  float* d_b; cudaMalloc(&d_b, bm.size() * 4); cudaMemcpy(d_b, bm.data(), bm.size() * 4, cudaMemcpyHostToDevice);
  const size_t n = (size_t)tiles_check * 2048, n_all = (size_t)tiles_max * 2048;
  std::vector<float2> h(n_all);
  srand(1);
  for (auto& v : h) v = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
  float2 *d_in, *d_o1, *d_o2;
  cudaMalloc(&d_in, n_all * 8); cudaMalloc(&d_o1, n_all * 8); cudaMalloc(&d_o2, n_all * 8);
  cudaMemcpy(d_in, h.data(), n_all * 8, cudaMemcpyHostToDevice);
  const size_t smemA = sizeof(cplx) * (pad16(2048) + 2);
  cudaFuncSetAttribute(dft16_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemB + 1024);
  // ---- correctness of ONE pass against a double-precision DFT-16
  dft16_cuda_core<<<tiles_check, 128, smemA>>>(d_in, d_o1, tiles_check, 1);
  dft16_tcgen05<<<tiles_check, 128, kSmemB + 1024>>>(d_in, d_o2, tiles_check, 1, d_b);
  cudaError_t err = cudaDeviceSynchronize();
  printf("one pass: %s\n", cudaGetErrorString(err));
  if (err != cudaSuccess) return 1;
  std::vector<float2> o1(n), o2(n);
  cudaMemcpy(o1.data(), d_o1, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(o2.data(), d_o2, n * 8, cudaMemcpyDeviceToHost);
  double e1 = 0, e2 = 0, mx = 0;
  for (int tile = 0; tile < tiles_check; ++tile)
    for (int t = 0; t < 128; ++t)
      for (int f = 0; f < 16; ++f) {
        double yr = 0, yi = 0;
        for (int e = 0; e < 16; ++e) {
          const float2 x = h[(size_t)tile * 2048 + t + e * 128];
          const double th = -2.0 * M_PI * e * f / 16.0;
          yr += x.x * cos(th) - x.y * sin(th); yi += x.x * sin(th) + x.y * cos(th);
        }
        const size_t idx = (size_t)tile * 2048 + 16 * t + f;
        mx = fmax(mx, fmax(fabs(yr), fabs(yi)));
        e1 = fmax(e1, fmax(fabs(o1[idx].x - yr), fabs(o1[idx].y - yi)));
        e2 = fmax(e2, fmax(fabs(o2[idx].x - yr), fabs(o2[idx].y - yi)));
      }
  printf("max-norm relative error of one radix-16 pass vs fp64:  CUDA cores %.3e   tcgen05 3xTF32 %.3e\n", e1 / mx, e2 / mx);
  // ---- chained passes agree between the two variants (same Stockham chain)
  dft16_cuda_core<<<tiles_check, 128, smemA>>>(d_in, d_o1, tiles_check, 3);
  dft16_tcgen05<<<tiles_check, 128, kSmemB + 1024>>>(d_in, d_o2, tiles_check, 3, d_b);
  cudaDeviceSynchronize();
  cudaMemcpy(o1.data(), d_o1, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(o2.data(), d_o2, n * 8, cudaMemcpyDeviceToHost);
  double dmax = 0, m3 = 0;
  for (size_t i = 0; i < n; ++i) { m3 = fmax(m3, fabs(o1[i].x)); dmax = fmax(dmax, fmax(fabs(o1[i].x - o2[i].x), fabs(o1[i].y - o2[i].y))); }
  printf("three chained passes, tcgen05 vs CUDA cores: max-norm relative difference %.3e\n", dmax / m3);
  // ---- on-chip cost of a pass: resident data, `passes_bench` chained passes per tile
  for (int per_sm = 1; per_sm <= 6; ++per_sm) {
    const int grid = sms * per_sm, tiles = grid * 4;
    float ms[2];
    for (int variant = 0; variant < 2; ++variant) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      for (int rep = 0; rep < 3; ++rep) {
        if (rep == 1) cudaEventRecord(a);
        if (variant == 0) dft16_cuda_core<<<grid, 128, smemA>>>(d_in, d_o1, tiles, passes_bench);
        else dft16_tcgen05<<<grid, 128, kSmemB + 1024>>>(d_in, d_o2, tiles, passes_bench, d_b);
      }
      cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms[variant], a, b);
      ms[variant] *= 0.5f;
    }
    const double tile_passes = (double)tiles * passes_bench;
    printf("CTAs/SM %d: CUDA cores %8.1f us (%5.0f cyc per tile-pass per SM @1.965 GHz)   tcgen05 %8.1f us (%5.0f cyc per tile-pass per SM)\n",
           per_sm, ms[0] * 1e3, ms[0] * 1e-3 * 1.965e9 * sms / tile_passes, ms[1] * 1e3, ms[1] * 1e-3 * 1.965e9 * sms / tile_passes);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
