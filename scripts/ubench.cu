// scripts/ubench.cu -- FP64 pipe micro-benchmarks on B200 (developer aid; numbers quoted in DESIGN.md / profiles/).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench scripts/ubench.cu && /tmp/ubench
// Measures, with clock64() around an unrolled loop executed by ONE warp on one SM:
//   dependent-issue latency of DFMA / DADD / DMUL, issue cost with k independent chains, MUFU.RSQ64H, SHFL (64-bit),
//   I2F.F64.S64; and the full-chip DFMA rate with all-register operands.
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void k_lat_dfma(double* out, long long* cyc, int iters, double a, double b) {
  double x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) x[c] = threadIdx.x * 1e-3 + c;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int c = 0; c < CHAINS; c++) x[c] = fma(x[c], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += x[c];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_lat_dadd(double* out, long long* cyc, int iters, double b) {
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) x = x + b;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_lat_rsq(double* out, long long* cyc, int iters) {
  double x = 1.0 + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x));
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_lat_shfl(double* out, long long* cyc, int iters) {
  double x = 1.0 + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int CHAINS>
__global__ void k_i2f(double* out, long long* cyc, int iters) {
  long long v[CHAINS];
  double acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) { v[c] = threadIdx.x + c; acc[c] = 0; }
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int c = 0; c < CHAINS; c++) { double d = __ll2double_rn(v[c]); v[c] += __double_as_longlong(d) & 3; acc[c] = d; }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += acc[c] + (double)v[c];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int CHAINS>
__global__ void __launch_bounds__(256) k_full(double* out, int iters, double a, double b) {
  double x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) x[c] = threadIdx.x * 1e-3 + c;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int c = 0; c < CHAINS; c++) x[c] = fma(x[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  double* d; long long* c; long long hc;
  cudaMalloc(&d, 148 * 16 * 256 * 8); cudaMalloc(&c, 8);
  const int it = 4096;
#define RUN1(name, launch, nops)                                                          \
  launch; launch; cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);                          \
  printf("%-34s %8.2f cycles/op (1 warp)\n", name, (double)hc / (double)(nops));
  RUN1("DFMA dependent chain", (k_lat_dfma<1><<<1, 32>>>(d, c, it, 0.999, 1e-9)), it * 8)
  RUN1("DFMA 2 chains (per DFMA)", (k_lat_dfma<2><<<1, 32>>>(d, c, it, 0.999, 1e-9)), it * 16)
  RUN1("DFMA 4 chains (per DFMA)", (k_lat_dfma<4><<<1, 32>>>(d, c, it, 0.999, 1e-9)), it * 32)
  RUN1("DFMA 8 chains (per DFMA)", (k_lat_dfma<8><<<1, 32>>>(d, c, it, 0.999, 1e-9)), it * 64)
  RUN1("DADD dependent chain", (k_lat_dadd<<<1, 32>>>(d, c, it, 1e-9)), it * 8)
  RUN1("MUFU.RSQ64H dependent", (k_lat_rsq<<<1, 32>>>(d, c, it)), it * 8)
  RUN1("SHFL.64 + DADD dependent", (k_lat_shfl<<<1, 32>>>(d, c, it)), it * 8)
  RUN1("I2F.F64.S64 1 chain (+LOP,IADD)", (k_i2f<1><<<1, 32>>>(d, c, it)), it * 8)
  RUN1("I2F.F64.S64 4 chains (per I2F)", (k_i2f<4><<<1, 32>>>(d, c, it)), it * 32)
  // 4 warps on one SM, one per SMSP? (128 threads): per-SMSP issue cost with 8 chains
  k_lat_dfma<8><<<1, 128>>>(d, c, it, 0.999, 1e-9); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
  printf("%-34s %8.2f cycles/op per warp (4 warps/SM)\n", "DFMA 8 chains", (double)hc / (it * 64.0));
  k_lat_dfma<8><<<1, 256>>>(d, c, it, 0.999, 1e-9); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
  printf("%-34s %8.2f cycles/op per warp (8 warps/SM = 2 per SMSP)\n", "DFMA 8 chains", (double)hc / (it * 64.0));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int occ = 4; occ <= 8; occ += 4) {
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      k_full<8><<<148 * occ, 256>>>(d, 1 << 13, 0.999999, 1e-9);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 2) printf("full chip DFMA, %d CTAs/SM x 256 thr x 8 chains: %.2f TFLOP/s\n", occ, 2.0 * 8 * 4 * (1 << 13) * 148.0 * occ * 256 / (ms * 1e-3) / 1e12);
    }
  }
  return 0;
}
