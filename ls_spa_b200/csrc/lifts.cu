// Per-permutation core of LS-SPA on sm_100a.
//
// Replaces square_shapley (reference ls_spa/ls_spa.py:256-287) and the antithetic
// pair average (:205-208).  One CTA owns one sample (a permutation, or a
// permutation and its reverse) at a time and keeps two column-major
// p x (p+1) fp64 tiles resident (shared memory when they fit, an L2-resident
// global slice otherwise):
//
//   A = [ R_tr[:, perm] | c_tr ]   Householder-triangularised in place; after step k
//                                  row k holds R[k, k:] and c[k] (the reference forms
//                                  Q explicitly, :275-278 -- only R and Q^T c matter)
//   X = [ R_te[:, perm] | c_te ]   eliminated against the rows of R as they become
//                                  final: X[:, j] -= (X[:, k] / R[k,k]) * R[k, j].
//                                  The multipliers are the columns of W = X R^-1 and
//                                  the last column is the test residual of the prefix
//                                  model, c_te - X[:, :k+1] theta_{k+1}  (:279-283).
//
// cost_k = |residual_k|^2, lift at position k = (cost_k - cost_{k+1}) / |y_test|^2
// scattered to feature perm[k] (:284-285).
//
// Warp work unit ("task"): 8 columns x 4 row groups (lane>>2 -> column, lane&3 -> row
// residue mod 4); with ld % 16 == 4 every fp64 access of a task is bank-conflict free
// and the Householder dot products need two shuffle steps only.

#include "common.cuh"

#include <stdlib.h>

namespace lsspa {

struct LiftParams {
  int p;
  int ld;
  const double *Rtr;  // column-major p x p
  const double *ctr;
  const double *Rte;  // column-major p x p
  const double *cte;
  double inv_ynsq;
  const int32_t *perms;
  int64_t count;
  int anti;
  double *out;
  double *gws;        // global workspace (used when the tiles do not fit in smem)
  int64_t gws_stride; // doubles per CTA
};

__device__ __forceinline__ void hh_vector(double *A, int ld, int p, int k, int lane, double *s_tau) {
  // Householder reflector for A[k:p, k] (LAPACK dlarfg convention: v[0] = 1 implied)
  double *col = A + (size_t)k * ld;
  double sig = 0.0;
  for (int i = k + 1 + lane; i < p; i += kWarp) {
    double x = col[i];
    sig = fma(x, x, sig);
  }
  sig = warp_sum(sig);
  double x0 = col[k];
  double tau = 0.0, scale = 0.0, beta = x0;
  if (sig > kTinySig) {
    double nrm = sqrt(fma(x0, x0, sig));
    beta = (x0 >= 0.0) ? -nrm : nrm;
    tau = (beta - x0) / beta;
    scale = 1.0 / (x0 - beta);
  }
  for (int i = k + 1 + lane; i < p; i += kWarp) col[i] *= scale;
  if (lane == 0) {
    col[k] = beta;
    *s_tau = tau;
  }
}

// apply H_k = I - tau v v^T to columns k+1..p of A  (v = [1; A[k+1:p, k]])
__device__ __forceinline__ void hh_update(double *A, int ld, int p, int k, double tau, int warp,
                                          int nwarps, int lane) {
  const int rg = lane & 3, cg = lane >> 2;
  const double *v = A + (size_t)k * ld;
  const int ntasks = (p - k + 7) >> 3;
  for (int t = warp; t < ntasks; t += nwarps) {
    const int j = k + 1 + 8 * t + cg;
    const bool valid = j <= p;
    double *cj = A + (size_t)(valid ? j : k + 1) * ld;
    double w = 0.0, akj = 0.0;
    if (valid) {
      akj = cj[k];
      for (int i = k + 1 + rg; i < p; i += 4) w = fma(v[i], cj[i], w);
    }
    w = quad_sum(w);
    __syncwarp();
    if (valid) {
      w = tau * (akj + w);
      if (rg == 0) cj[k] = akj - w;
      for (int i = k + 1 + rg; i < p; i += 4) cj[i] = fma(-w, v[i], cj[i]);
    }
  }
}

// X[:, j] -= (X[:, k] / R[k,k]) * R[k, j]   for j = k+1..p
__device__ __forceinline__ void x_update(const double *A, double *X, int ld, int p, int k, int warp,
                                         int nwarps, int lane) {
  const int rg = lane & 3, cg = lane >> 2;
  const double rinv = 1.0 / A[(size_t)k * ld + k];
  const double *xk = X + (size_t)k * ld;
  const int ntasks = (p - k + 7) >> 3;
  for (int t = warp; t < ntasks; t += nwarps) {
    const int j = k + 1 + 8 * t + cg;
    if (j <= p) {
      const double akj = A[(size_t)j * ld + k] * rinv;
      double *xj = X + (size_t)j * ld;
      for (int i = rg; i < p; i += 4) xj[i] = fma(-xk[i], akj, xj[i]);
    }
  }
}

__device__ __forceinline__ void resid_cost(const double *X, int ld, int p, int lane, double *dst) {
  const double *r = X + (size_t)p * ld;
  double s = 0.0;
  for (int i = lane; i < p; i += kWarp) s = fma(r[i], r[i], s);
  s = warp_sum(s);
  if (lane == 0) *dst = s;
}

template <bool kSmemTiles>
__global__ void __launch_bounds__(512) lifts_kernel(LiftParams a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int p = a.p, ld = a.ld;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;

  double *sm = reinterpret_cast<double *>(smem_raw);
  double *A, *X, *small;
  const size_t tile = (size_t)(p + 1) * ld;
  if (kSmemTiles) {
    A = sm;
    X = sm + tile;
    small = sm + 2 * tile;
  } else {
    A = a.gws + (size_t)blockIdx.x * a.gws_stride;
    X = A + tile;
    small = sm;
  }
  double *cost = small;            // p+1
  double *acc = cost + (p + 1);    // p
  double *s_tau = acc + p;         // 1 (+1 pad)
  int *perm_s = reinterpret_cast<int *>(s_tau + 2);  // p

  const int halves = a.anti ? 2 : 1;
  const double weight = a.anti ? 0.5 : 1.0;

  for (int64_t s = blockIdx.x; s < a.count; s += gridDim.x) {
    for (int h = 0; h < halves; ++h) {
      __syncthreads();  // previous user of perm_s / tiles is done
      for (int k = tid; k < p; k += nt)
        perm_s[k] = a.perms[s * p + (h == 0 ? k : p - 1 - k)];
      __syncthreads();
      // gather the permuted columns; both reduced factors are upper triangular, so
      // column `col` is non-zero in rows 0..col only
      for (int e = tid; e < p * p; e += nt) {
        const int k = e / p, i = e - k * p;
        const int col = perm_s[k];
        const bool nz = i <= col;
        A[(size_t)k * ld + i] = nz ? a.Rtr[(size_t)col * p + i] : 0.0;
        X[(size_t)k * ld + i] = nz ? a.Rte[(size_t)col * p + i] : 0.0;
      }
      for (int i = tid; i < p; i += nt) {
        A[(size_t)p * ld + i] = a.ctr[i];
        X[(size_t)p * ld + i] = a.cte[i];
      }
      __syncthreads();
      if (warp == nwarps - 1) resid_cost(X, ld, p, lane, &cost[0]);

      for (int k = 0; k < p; ++k) {
        // phase A: reflector of column k  ||  elimination of X with pivot row k-1
        if (warp == 0) hh_vector(A, ld, p, k, lane, s_tau);
        if (k > 0) x_update(A, X, ld, p, k - 1, warp, nwarps, lane);
        __syncthreads();
        // phase B: trailing update with H_k  ||  cost of prefix k
        if (k > 0 && warp == nwarps - 1) resid_cost(X, ld, p, lane, &cost[k]);
        hh_update(A, ld, p, k, *s_tau, warp, nwarps, lane);
        __syncthreads();
      }
      x_update(A, X, ld, p, p - 1, warp, nwarps, lane);
      __syncthreads();
      if (warp == 0) resid_cost(X, ld, p, lane, &cost[p]);
      __syncthreads();
      for (int k = tid; k < p; k += nt) {
        const double lift = (cost[k] - cost[k + 1]) * a.inv_ynsq;
        const int f = perm_s[k];
        acc[f] = (h == 0 ? 0.0 : acc[f]) + weight * lift;
      }
    }
    __syncthreads();
    for (int f = tid; f < p; f += nt) a.out[s * p + f] = acc[f];
  }
}

static size_t small_smem_bytes(int p) {
  // cost[p+1] + acc[p] + tau[2] doubles, perm_s[p] ints
  return (size_t)(2 * p + 3) * sizeof(double) + (size_t)p * sizeof(int) + 16;
}

static size_t tile_smem_bytes(int p) {
  return 2 * (size_t)(p + 1) * padded_ld(p) * sizeof(double);
}

static bool tiles_fit_smem(int p) {
  const DeviceInfo &d = device_info();
  int limit = d.smem_optin > 0 ? d.smem_optin : 227 * 1024;
  return tile_smem_bytes(p) + small_smem_bytes(p) <= (size_t)limit;
}

static int lifts_grid(int p, int64_t count) {
  const DeviceInfo &d = device_info();
  int sms = d.sm_count > 0 ? d.sm_count : 148;
  int64_t per_sm = 1;
  if (tiles_fit_smem(p)) {
    size_t per_cta = tile_smem_bytes(p) + small_smem_bytes(p) + 1024;
    per_sm = (int64_t)((227 * 1024) / per_cta);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 16) per_sm = 16;
  }
  int64_t g = per_sm * sms;
  if (g > count) g = count;
  if (g < 1) g = 1;
  return (int)g;
}

static int lifts_threads(int p) {
  if (p <= 12) return 64;
  if (p <= 32) return 128;
  if (p <= 64) return 256;
  return 512;
}

}  // namespace lsspa

using namespace lsspa;

// LSSPA_LIFTS_IMPL=v1 forces the scalar kernel of this file (debugging / A-B timing)
static bool use_mma(int p) {
  static const bool forced_v1 = [] {
    const char *e = getenv("LSSPA_LIFTS_IMPL");
    return e && e[0] == 'v' && e[1] == '1';
  }();
  return !forced_v1 && lifts_mma_supported(p);
}

extern "C" size_t lsspa_lifts_workspace_bytes(int p, int64_t count) {
  if (p < 1 || count < 1) return 0;
  if (use_mma(p) || tiles_fit_smem(p)) return 0;
  return (size_t)lifts_grid(p, count) * 2 * (size_t)(p + 1) * padded_ld(p) * sizeof(double);
}

extern "C" int lsspa_lifts(int p, const double *R_tr_cm, const double *c_tr, const double *R_te_cm,
                           const double *c_te, double y_norm_sq, const int32_t *perms, int64_t count,
                           int antithetical, double *lifts_out, void *workspace,
                           size_t workspace_bytes, void *stream) {
  if (p < 1 || p > 32767 || !R_tr_cm || !c_tr || !R_te_cm || !c_te || !perms || !lifts_out)
    return LSSPA_E_BADARG;
  if (count == 0) return LSSPA_OK;
  if (count < 0) return LSSPA_E_BADARG;
  if (use_mma(p))
    return lifts_mma_launch(p, R_tr_cm, c_tr, R_te_cm, c_te, y_norm_sq, perms, count, antithetical, lifts_out,
                            as_stream(stream));
  LiftParams a;
  a.p = p;
  a.ld = padded_ld(p);
  a.Rtr = R_tr_cm;
  a.ctr = c_tr;
  a.Rte = R_te_cm;
  a.cte = c_te;
  a.inv_ynsq = 1.0 / y_norm_sq;
  a.perms = perms;
  a.count = count;
  a.anti = antithetical ? 1 : 0;
  a.out = lifts_out;
  a.gws = nullptr;
  a.gws_stride = 0;
  const int grid = lifts_grid(p, count);
  const int nt = lifts_threads(p);
  cudaStream_t st = as_stream(stream);
  if (tiles_fit_smem(p)) {
    size_t smem = tile_smem_bytes(p) + small_smem_bytes(p);
    LSSPA_CUDA_TRY(cudaFuncSetAttribute(lifts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    lifts_kernel<true><<<grid, nt, smem, st>>>(a);
  } else {
    size_t need = lsspa_lifts_workspace_bytes(p, count);
    if (!workspace || workspace_bytes < need) return LSSPA_E_WORKSPACE;
    a.gws = reinterpret_cast<double *>(workspace);
    a.gws_stride = 2 * (int64_t)(p + 1) * a.ld;
    size_t smem = small_smem_bytes(p);
    LSSPA_CUDA_TRY(cudaFuncSetAttribute(lifts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    lifts_kernel<false><<<grid, nt, smem, st>>>(a);
  }
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}
