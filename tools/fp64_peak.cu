// FP64 pipe micro-benchmark for the roofline denominator of the per-permutation kernel
// (MEASURED_PEAKS.json has HBM and bf16 only).  Prints one JSON object.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void dfma_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void dmma_kernel(double *out, int iters, double a, double b) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) dmma(c[j][0], c[j][1], a, b);
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void mixed_kernel(double *out, int iters, double a, double b) {
  double c[4][2];
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  for (int i = 0; i < 4; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) dmma(c[j][0], c[j][1], a, b);
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
  }
  double s = x0 + x1 + x2 + x3;
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
static double time_ms(K launch, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8, threads = 256, iters = 20000;
  double *out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  double t1 = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
  double f1 = 2.0 * 8 * iters * (double)blocks * threads / (t1 * 1e-3) / 1e12;
  double t2 = time_ms([&] { dmma_kernel<<<blocks, threads>>>(out, iters / 4, 1.0000001, 1e-9); }, 5);
  double f2 = 2.0 * 8 * 8 * 4 * 8 * (iters / 4) * (double)blocks * (threads / 32) / (t2 * 1e-3) / 1e12;
  double t3 = time_ms([&] { mixed_kernel<<<blocks, threads>>>(out, iters / 4, 1.0000001, 1e-9); }, 5);
  double f3 = (2.0 * 256 * 4 * (threads / 32) + 2.0 * 4 * threads) * (iters / 4) * (double)blocks / (t3 * 1e-3) / 1e12;
  cudaError_t err = cudaDeviceSynchronize();
  printf("{\"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, \"mixed_tflops\": %.2f, \"dfma_ms\": %.3f, \"dmma_ms\": %.3f, \"mixed_ms\": %.3f, \"cuda\": \"%s\"}\n",
         sms, f1, f2, f3, t1, t2, t3, cudaGetErrorString(err));
  return err != cudaSuccess;
}
