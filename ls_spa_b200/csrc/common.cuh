// Shared helpers for the LS-SPA sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "lsspa.h"

#define LSSPA_CUDA_TRY(expr)                                   \
  do {                                                         \
    cudaError_t err__ = (expr);                                \
    if (err__ != cudaSuccess) return -(1000 + (int)err__);     \
  } while (0)

#define LSSPA_LAUNCH_CHECK() LSSPA_CUDA_TRY(cudaGetLastError())

// Cycle probes of the lift kernels (tools/lifts_cycles.py): compiled in only with
// -DLSSPA_LIFTS_TIMING (LSSPA_EXTRA_NVCC_FLAGS); the shipped library carries none.
#ifdef LSSPA_LIFTS_TIMING
#define LSSPA_CLOCK() clock64()
#else
#define LSSPA_CLOCK() 0LL
#endif

namespace lsspa {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
// A column whose sub-pivot part has squared norm below this is treated as already zero (no
// reflector): such values are round-off residue, and squaring them again would underflow and
// turn 1 / (beta * u1) into inf.
constexpr double kTinySig = 1e-280;

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Leading dimension for column-major fp64 tiles in shared memory: the smallest
// ld >= n with ld % 16 == 4, so that 4 consecutive rows x 8 columns (the access
// shape of one warp: lane&3 -> row, lane>>2 -> column) hit 32 distinct banks.
__host__ __device__ inline int padded_ld(int n) { return n + ((4 - (n % 16)) + 16) % 16; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// sum over the 4 lanes that share lane>>2 (lane&3 = row group)
__device__ __forceinline__ double quad_sum(double v) {
  v += __shfl_xor_sync(kFull, v, 1);
  v += __shfl_xor_sync(kFull, v, 2);
  return v;
}

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

struct DeviceInfo {
  int sm_count;
  int smem_optin;
  int cc_major;
};
// cached per process (device of the first call); 0 on failure
const DeviceInfo &device_info();

// DMMA lift kernel (lifts_mma.cu), used by lsspa_lifts for 9 <= p <= 128
bool lifts_mma_supported(int p);
int lifts_mma_launch(int p, const double *R_tr_cm, const double *c_tr, const double *R_te_cm, const double *c_te,
                     double y_norm_sq, const int32_t *perms, int64_t count, int antithetical, double *lifts_out,
                     cudaStream_t st);

// wide problems (lifts_big.cu): p > 152
bool lifts_big_supported(int p);
int lifts_big_cond(int p, const double *R_tr_cm, const double *D, const double *Gh, double *Xinv, double *colstat,
                   double *info, cudaStream_t st);

}  // namespace lsspa
