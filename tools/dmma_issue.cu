// Microbenchmark: issue interval and dependent latency of mma.sync.m8n8k4.f64 (DMMA.8x8x4) on one SM
// as a function of the number of warps per SM sub-partition.  Development aid for the lift kernels.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void k(double *out, long long *cyc, int iters) {
  double acc[CHAINS][2];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc[i][0] = acc[i][1] = threadIdx.x * 1e-9;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) dmma(acc[i][0], acc[i][1], a, b);
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CHAINS>
void run(int warps, double *out, long long *cyc) {
  const int iters = 2000;
  k<CHAINS><<<148, warps * 32>>>(out, cyc, iters);
  k<CHAINS><<<148, warps * 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / (iters * CHAINS);
  printf("warps/SM %2d (per SMSP %4.1f)  chains %d  cycles per DMMA per warp %6.1f  -> per SMSP issue interval %5.1f\n",
         warps, warps / 4.0, CHAINS, per, per / (warps / 4.0 < 1 ? 1 : warps / 4.0));
}

int main() {
  double *out;
  long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&cyc, 8);
  for (int w : {1, 4, 8, 16, 32}) {
    run<1>(w, out, cyc);
    run<2>(w, out, cyc);
    run<4>(w, out, cyc);
    run<8>(w, out, cyc);
  }
  return 0;
}
