// fp64_lat.cu -- dependent-issue latency and per-warp throughput of the operations on the SSOR
// sweep's chain (tuning aid; nvcc -arch=sm_100a -o fp64_lat fp64_lat.cu && ./fp64_lat).
// One warp, clock64 around N dependent (or 8-way independent) operations.
#include <cstdio>
#include <cuda_runtime.h>

#define N 4096

template <int OP>
__global__ void lat(double* out, long long* cyc, double x0, double y0)
{
  __shared__ double sm[64];
  __shared__ unsigned long long mb[64];
  sm[threadIdx.x & 63] = x0;
  mb[threadIdx.x & 63] = 0;
  __syncthreads();
  double x = x0 + threadIdx.x, y = y0;
  double a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
  if (OP == 8) { a1 = y; a2 = y * 0.125; a3 = y * 0.5; a4 = y * 0.125; a5 = y * 0.25; a6 = y * 0.0625; a7 = y * 4.0; }   // contraction
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = __dadd_rn(x, y);
    if (OP == 1) x = __dmul_rn(x, y);
    if (OP == 2) x = __fma_rn(x, y, y);
    if (OP == 3) x = __ddiv_rn(x, y);
    if (OP == 4) x = __dsqrt_rn(x);
    if (OP == 5) x = __shfl_up_sync(0xffffffffu, x, 1);
    if (OP == 6) { x = sm[(__double2loint(x) & 31)]; }                      // dependent LDS.64
    if (OP == 7) {                                                           // 8 independent DFMA chains
      x = __fma_rn(x, y, y); a1 = __fma_rn(a1, y, y); a2 = __fma_rn(a2, y, y); a3 = __fma_rn(a3, y, y);
      a4 = __fma_rn(a4, y, y); a5 = __fma_rn(a5, y, y); a6 = __fma_rn(a6, y, y); a7 = __fma_rn(a7, y, y);
    }
    if (OP == 8) {                                                           // the forward SSOR chain
      const double zh = __shfl_up_sync(0xffffffffu, x, 1);
      double s = __dadd_rn(a1, __dmul_rn(a2, zh));
      s = __dadd_rn(s, a3);
      s = __dadd_rn(s, __dmul_rn(a4, x));
      s = __dadd_rn(s, a5);
      x = __dadd_rn(a6, __ddiv_rn(__dmul_rn(y, s), a7));
    }
    if (OP == 9) { asm volatile("bar.sync 1, 32;" ::: "memory"); }
    if (OP == 10) {                                                          // 8 independent DADD
      x = __dadd_rn(x, y); a1 = __dadd_rn(a1, y); a2 = __dadd_rn(a2, y); a3 = __dadd_rn(a3, y);
      a4 = __dadd_rn(a4, y); a5 = __dadd_rn(a5, y); a6 = __dadd_rn(a6, y); a7 = __dadd_rn(a7, y);
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
static void run(const char* name, int per_iter, double x0, double y0)
{
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * sizeof(double));
  cudaMallocManaged(&cyc, sizeof(long long));
  for (int r = 0; r < 3; ++r) { lat<OP><<<1, 32>>>(out, cyc, x0, y0); cudaDeviceSynchronize(); }
  printf("%-28s %8.2f cycles per op (%d ops per iteration, %.1f cycles per iteration)\n", name,
         (double)*cyc / N / per_iter, per_iter, (double)*cyc / N);
  cudaFree(out); cudaFree(cyc);
}

int main()
{
  run<0>("DADD dependent", 1, 1.0, 1e-9);
  run<1>("DMUL dependent", 1, 1.0, 1.0000001);
  run<2>("DFMA dependent", 1, 0.5, 0.999);
  run<3>("ddiv_rn dependent", 1, 1.0, 1.0000001);
  run<4>("dsqrt_rn dependent", 1, 2.0, 0.0);
  run<5>("SHFL.64 dependent", 1, 1.0, 0.0);
  run<6>("LDS.64 dependent", 1, 3.0, 0.0);
  run<7>("DFMA 8 independent", 8, 0.5, 0.999);
  run<10>("DADD 8 independent", 8, 1.0, 1e-9);
  run<8>("SSOR forward chain (1 cell)", 1, 0.1, 1.4);
  run<9>("bar.sync (1 warp)", 1, 0.0, 0.0);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
