// Tall-skinny reduction of [X | y] to a (p+1) x (p+1) triangular factor (TSQR).
//
// Replaces reduce_data (reference ls_spa/ls_spa.py:290-318): the reference calls LAPACK
// on the whole (N+p) x p block and materialises Q (:314-317); only R and Q^T y are used
// afterwards, and Q^T y is the last column of the R factor of [X | y].  Here every CTA
// streams a contiguous slab of rows through shared memory in blocks of kRB rows and
// folds each block into a running triangle T with Householder reflectors that exploit
// the [triangle; dense block] structure (one reflector = 1 + kRB entries).  The per-CTA
// triangles are then merged group-wise by the same kernel body (lsspa_tsqr_merge); the
// ridge rows sqrt(reg) I (:310) join the stack as one more triangle.
//
// Layout: T row-major q x q (q = p+1), block B row-major kRB x q; thread j owns column
// j, so T[k][j] / B[i][j] accesses are contiguous across the warp (no bank conflicts,
// coalesced when T lives in global memory for large p) and the reflector dot products
// need no cross-thread reduction.

#include "common.cuh"

#include <stdlib.h>

namespace lsspa {

constexpr int kRBMax = 32;   // rows per block (one warp computes the reflector)
constexpr int kSlotExtra = 8;

struct TsqrGeom {
  int q;          // p + 1
  int rb;         // rows per block
  int threads;
  bool t_in_smem;
  size_t smem;
};

static TsqrGeom tsqr_geom(int p) {
  TsqrGeom g;
  g.q = p + 1;
  const DeviceInfo &d = device_info();
  const size_t limit = (size_t)(d.smem_optin > 0 ? d.smem_optin : 227 * 1024);
  const size_t half = (limit - 2048) / 2;  // aim for two CTAs per SM
  const size_t tbytes = (size_t)g.q * g.q * sizeof(double);
  const size_t bbytes32 = (size_t)kRBMax * g.q * sizeof(double);
  const size_t misc = (size_t)(kRBMax + 8) * sizeof(double);
  if (tbytes + bbytes32 + misc <= half) {
    g.t_in_smem = true;
    g.rb = kRBMax;
  } else if (tbytes + bbytes32 + misc <= limit) {
    g.t_in_smem = true;
    g.rb = kRBMax;
  } else {
    g.t_in_smem = false;
    int rb = (int)((limit - misc - 1024) / ((size_t)g.q * sizeof(double)));
    if (rb > kRBMax) rb = kRBMax;
    if (rb < 1) rb = 1;
    g.rb = rb;
  }
  g.smem = (g.t_in_smem ? tbytes : 0) + (size_t)g.rb * g.q * sizeof(double) + misc;
  int nt = ((g.q + 31) / 32) * 32;
  if (nt > 1024) nt = 1024;
  if (nt < 64) nt = 64;
  g.threads = nt;
  return g;
}

// Fold the rb x q block B (rows with leading zeros in columns < kstart) into T.
__device__ __forceinline__ void absorb_block(double *T, double *B, double *v, double *s_tau, int q,
                                             int rb, int kstart) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int k = kstart; k < q; ++k) {
    if (tid < kWarp) {
      const double x = (tid < rb) ? B[(size_t)tid * q + k] : 0.0;
      const double sig = warp_sum(x * x);
      const double x0 = T[(size_t)k * q + k];
      double tau = 0.0, scale = 0.0, beta = x0;
      if (sig > kTinySig) {
        const double nrm = sqrt(fma(x0, x0, sig));
        beta = (x0 >= 0.0) ? -nrm : nrm;
        tau = (beta - x0) / beta;
        scale = 1.0 / (x0 - beta);
      }
      if (tid < rb) v[tid] = x * scale;
      if (tid == 0) {
        T[(size_t)k * q + k] = beta;
        *s_tau = tau;
      }
    }
    __syncthreads();
    const double tau = *s_tau;
    if (tau != 0.0) {
      for (int j = k + 1 + tid; j < q; j += nt) {
        const double tkj = T[(size_t)k * q + j];
        double w = tkj;
        for (int i = 0; i < rb; ++i) w = fma(v[i], B[(size_t)i * q + j], w);
        w *= tau;
        T[(size_t)k * q + j] = tkj - w;
        for (int i = 0; i < rb; ++i) B[(size_t)i * q + j] = fma(-w, v[i], B[(size_t)i * q + j]);
      }
    }
    __syncthreads();
  }
}

struct RowsParams {
  const double *X;
  int64_t ldx;
  const double *y;
  int64_t nrows;
  int p;
  double divisor;
  double *parts;
  int64_t slot;
  int nparts;
  int rb;
  int t_in_smem;
};

__global__ void tsqr_rows_kernel(RowsParams a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int q = a.p + 1, rb = a.rb;
  const int tid = threadIdx.x, nt = blockDim.x;
  double *sm = reinterpret_cast<double *>(smem_raw);
  double *slot = a.parts + (size_t)blockIdx.x * a.slot;
  double *T = a.t_in_smem ? sm : slot;
  double *B = a.t_in_smem ? sm + (size_t)q * q : sm;
  double *v = B + (size_t)rb * q;
  double *s_tau = v + kRBMax;
  double *s_acc = s_tau + 1;

  for (int e = tid; e < q * q; e += nt) T[e] = 0.0;
  if (tid == 0) *s_acc = 0.0;

  // contiguous slab of rows per CTA, a multiple of rb
  int64_t per = ceil_div(ceil_div(a.nrows, (int64_t)a.nparts), (int64_t)rb) * rb;
  const int64_t r_begin = (int64_t)blockIdx.x * per;
  const int64_t r_end = (r_begin + per < a.nrows) ? r_begin + per : a.nrows;
  double ysq = 0.0;  // thread i < rb accumulates its rows
  __syncthreads();
  for (int64_t r0 = r_begin; r0 < r_end; r0 += rb) {
    const int rows = (int)((r_end - r0 < rb) ? r_end - r0 : rb);
    for (int e = tid; e < rb * q; e += nt) {
      const int i = e / q, j = e - i * q;
      double val = 0.0;
      if (i < rows) {
        val = (j < a.p) ? a.X[(r0 + i) * a.ldx + j] : a.y[r0 + i];
        val /= a.divisor;
      }
      B[e] = val;
    }
    __syncthreads();
    if (tid < rows) {
      const double yy = B[(size_t)tid * q + a.p];
      ysq = fma(yy, yy, ysq);
    }
    absorb_block(T, B, v, s_tau, q, rows, 0);
  }
  // block-level sum of ysq (threads 0..rb-1 hold partials; rb <= 32 -> warp 0)
  if (tid < kWarp) {
    const double s = warp_sum(tid < rb ? ysq : 0.0);
    if (tid == 0) *s_acc = s;
  }
  __syncthreads();
  if (a.t_in_smem)
    for (int e = tid; e < q * q; e += nt) slot[e] = T[e];
  if (tid < kSlotExtra) slot[(size_t)q * q + tid] = (tid == 0) ? *s_acc : 0.0;
}

struct MergeParams {
  const double *parts;
  int count;
  int group;
  int p;
  double *out;
  int64_t slot;
  int rb;
  int t_in_smem;
};

__global__ void tsqr_merge_kernel(MergeParams a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int q = a.p + 1, rb = a.rb;
  const int tid = threadIdx.x, nt = blockDim.x;
  double *sm = reinterpret_cast<double *>(smem_raw);
  double *slot = a.out + (size_t)blockIdx.x * a.slot;
  double *T = a.t_in_smem ? sm : slot;
  double *B = a.t_in_smem ? sm + (size_t)q * q : sm;
  double *v = B + (size_t)rb * q;
  double *s_tau = v + kRBMax;

  const int first = blockIdx.x * a.group;
  const int last = (first + a.group < a.count) ? first + a.group : a.count;
  const double *src0 = a.parts + (size_t)first * a.slot;
  for (int e = tid; e < q * q; e += nt) {
    const int i = e / q, j = e - i * q;
    T[e] = (j >= i) ? src0[e] : 0.0;
  }
  double extra = (tid == 0) ? src0[(size_t)q * q] : 0.0;
  __syncthreads();
  for (int t = first + 1; t < last; ++t) {
    const double *src = a.parts + (size_t)t * a.slot;
    if (tid == 0) extra += src[(size_t)q * q];
    for (int r0 = 0; r0 < q; r0 += rb) {
      const int rows = (q - r0 < rb) ? q - r0 : rb;
      for (int e = tid; e < rb * q; e += nt) {
        const int i = e / q, j = e - i * q;
        B[e] = (i < rows && j >= r0 + i) ? src[(size_t)(r0 + i) * q + j] : 0.0;
      }
      __syncthreads();
      absorb_block(T, B, v, s_tau, q, rows, r0);
    }
  }
  if (a.t_in_smem)
    for (int e = tid; e < q * q; e += nt) slot[e] = T[e];
  if (tid < kSlotExtra) slot[(size_t)q * q + tid] = (tid == 0) ? extra : 0.0;
}


// ---------------------------------------------------------------------------------------------
// Register-resident variant (q <= 512; 128 registers per thread cap the CTA at 512 threads): thread (j, h) keeps 32 rows of column j of the current
// block in registers (h = which half of a 64-row block when two threads share a column), so the
// block never touches shared memory.  Per reflector k the owner threads of column k compute
// u = x - beta e1 and tt = -1/(beta u1) themselves (no scaling pass) and publish u through a
// double-buffered 64-entry scratch: ONE barrier per reflector instead of two, 16-byte broadcast
// loads only.  T stays row-major (shared memory when it fits, else the output slot in L2).
template <int HALVES>
__device__ __forceinline__ void absorb_regs(double *T, double (&b)[32], double *us, int q, int j, int h,
                                            bool active, int kstart, int kend) {
  // us: 2 x (64 + 8) doubles: [u(0..63), u1, tt] per buffer
  for (int k = kstart; k < kend; ++k) {
    double *ub = us + (k & 1) * 72;
    if (active && j == k) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, s5 = 0.0, s6 = 0.0, s7 = 0.0;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        s0 = fma(b[i], b[i], s0);
        s1 = fma(b[i + 1], b[i + 1], s1);
        s2 = fma(b[i + 2], b[i + 2], s2);
        s3 = fma(b[i + 3], b[i + 3], s3);
        s4 = fma(b[i + 4], b[i + 4], s4);
        s5 = fma(b[i + 5], b[i + 5], s5);
        s6 = fma(b[i + 6], b[i + 6], s6);
        s7 = fma(b[i + 7], b[i + 7], s7);
      }
      double sig = ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
      if (HALVES == 2) sig += __shfl_xor_sync(__activemask(), sig, 1);
      const double x0 = T[(size_t)k * q + k];
      double u1 = 0.0, tt = 0.0, beta = x0;
      if (sig > kTinySig) {
        // dependent fp64 operations cost ~20 cycles each: one rsqrt, one reciprocal
        const double ss = fma(x0, x0, sig);
        const double nrm = ss * rsqrt(ss);
        tt = 1.0 / fma(fabs(x0), nrm, ss);      // = 1 / (|x| |u1|) = -1 / (beta u1)
        beta = (x0 >= 0.0) ? -nrm : nrm;
        u1 = x0 - beta;
      }
      double2 *dst = reinterpret_cast<double2 *>(ub + 32 * h);
#pragma unroll
      for (int i = 0; i < 16; ++i) dst[i] = make_double2(b[2 * i], b[2 * i + 1]);
      if (h == 0) {
        ub[64] = u1;
        ub[65] = tt;
        T[(size_t)k * q + k] = beta;
      }
    }
    __syncthreads();
    const double tt = ub[65];
    if (active && j > k && tt != 0.0) {
      const double2 *src = reinterpret_cast<const double2 *>(ub + 32 * h);
      double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0, d4 = 0.0, d5 = 0.0, d6 = 0.0, d7 = 0.0;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const double2 a = src[i], c = src[i + 1], e = src[i + 2], g = src[i + 3];
        d0 = fma(a.x, b[2 * i], d0);
        d1 = fma(a.y, b[2 * i + 1], d1);
        d2 = fma(c.x, b[2 * i + 2], d2);
        d3 = fma(c.y, b[2 * i + 3], d3);
        d4 = fma(e.x, b[2 * i + 4], d4);
        d5 = fma(e.y, b[2 * i + 5], d5);
        d6 = fma(g.x, b[2 * i + 6], d6);
        d7 = fma(g.y, b[2 * i + 7], d7);
      }
      double d = ((d0 + d1) + (d2 + d3)) + ((d4 + d5) + (d6 + d7));
      if (HALVES == 2) d += __shfl_xor_sync(__activemask(), d, 1);
      const double u1 = ub[64];
      const double tkj = T[(size_t)k * q + j];
      const double w = tt * fma(u1, tkj, d);
      if (h == 0) T[(size_t)k * q + j] = fma(-w, u1, tkj);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const double2 a = src[i];
        b[2 * i] = fma(-w, a.x, b[2 * i]);
        b[2 * i + 1] = fma(-w, a.y, b[2 * i + 1]);
      }
    }
  }
  __syncthreads();  // the scratch buffers may be reused by the next block
}

template <int HALVES>
__global__ void __launch_bounds__(512) tsqr_rows_regs_kernel(RowsParams a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int q = a.p + 1;
  const int tid = threadIdx.x, nt = blockDim.x;
  double *sm = reinterpret_cast<double *>(smem_raw);
  double *slot = a.parts + (size_t)blockIdx.x * a.slot;
  double *us = sm;                 // 144 doubles
  double *s_acc = sm + 144;        // 32 doubles (per-warp partial sums of y^2)
  double *T = a.t_in_smem ? sm + 176 : slot;
  const int j = tid / HALVES, h = tid % HALVES;
  const bool active = j < q;
  constexpr int RB = 32 * HALVES;

  for (int e = tid; e < q * q; e += nt) T[e] = 0.0;
  const int64_t per = ceil_div(ceil_div(a.nrows, (int64_t)a.nparts), (int64_t)RB) * RB;
  const int64_t r_begin = (int64_t)blockIdx.x * per;
  const int64_t r_end = (r_begin + per < a.nrows) ? r_begin + per : a.nrows;
  double ysq = 0.0;
  __syncthreads();
  double b[32];
  for (int64_t r0 = r_begin; r0 < r_end; r0 += RB) {
    const int64_t rbase = r0 + 32 * h;
    if (active) {
      if (j < a.p) {
        const double *src = a.X + rbase * a.ldx + j;
#pragma unroll
        for (int i = 0; i < 32; ++i) b[i] = (rbase + i < r_end) ? src[(int64_t)i * a.ldx] / a.divisor : 0.0;
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          b[i] = (rbase + i < r_end) ? a.y[rbase + i] / a.divisor : 0.0;
          ysq = fma(b[i], b[i], ysq);
        }
      }
    }
    absorb_regs<HALVES>(T, b, us, q, j, h, active, 0, q);
  }
  // sum of squares of the scaled targets (held by the threads of column p)
  if (tid < 32) s_acc[tid] = 0.0;
  __syncthreads();
  if (active && j == a.p) atomicAdd(&s_acc[0], ysq);
  __syncthreads();
  if (a.t_in_smem)
    for (int e = tid; e < q * q; e += nt) slot[e] = T[e];
  if (tid < kSlotExtra) slot[(size_t)q * q + tid] = (tid == 0) ? s_acc[0] : 0.0;
}

template <int HALVES>
__global__ void __launch_bounds__(512) tsqr_merge_regs_kernel(MergeParams a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int q = a.p + 1;
  const int tid = threadIdx.x, nt = blockDim.x;
  double *sm = reinterpret_cast<double *>(smem_raw);
  double *slot = a.out + (size_t)blockIdx.x * a.slot;
  double *us = sm;
  double *T = a.t_in_smem ? sm + 176 : slot;
  const int j = tid / HALVES, h = tid % HALVES;
  const bool active = j < q;
  constexpr int RB = 32 * HALVES;

  const int first = blockIdx.x * a.group;
  const int last = (first + a.group < a.count) ? first + a.group : a.count;
  const double *src0 = a.parts + (size_t)first * a.slot;
  for (int e = tid; e < q * q; e += nt) {
    const int i = e / q, jj = e - i * q;
    T[e] = (jj >= i) ? src0[e] : 0.0;
  }
  double extra = (tid == 0) ? src0[(size_t)q * q] : 0.0;
  __syncthreads();
  double b[32];
  for (int t = first + 1; t < last; ++t) {
    const double *src = a.parts + (size_t)t * a.slot;
    if (tid == 0) extra += src[(size_t)q * q];
    for (int r0 = 0; r0 < q; r0 += RB) {
      const int rbase = r0 + 32 * h;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int r = rbase + i;
        b[i] = (active && r < q && j >= r) ? src[(size_t)r * q + j] : 0.0;
      }
      // rows r0.. of a triangle are zero left of column r0; nothing below row q
      absorb_regs<HALVES>(T, b, us, q, j, h, active, r0, q);
    }
  }
  if (a.t_in_smem)
    for (int e = tid; e < q * q; e += nt) slot[e] = T[e];
  if (tid < kSlotExtra) slot[(size_t)q * q + tid] = (tid == 0) ? extra : 0.0;
}

struct RegsGeom {
  int halves;   // 0 = not applicable
  int threads;
  bool t_in_smem;
  size_t smem;
};

static RegsGeom regs_geom(int p) {
  RegsGeom g{0, 0, false, 0};
  const int q = p + 1;
  if (q > 512) return g;
  g.halves = (q <= 256) ? 2 : 1;
  if (const char *e = getenv("LSSPA_TSQR_HALVES")) g.halves = (e[0] == '1') ? 1 : g.halves;
  g.threads = ((q * g.halves + 31) / 32) * 32;
  const DeviceInfo &d = device_info();
  const size_t limit = (size_t)(d.smem_optin > 0 ? d.smem_optin : 227 * 1024);
  const size_t tbytes = (size_t)q * q * sizeof(double);
  g.t_in_smem = tbytes + 176 * sizeof(double) <= limit;
  g.smem = 176 * sizeof(double) + (g.t_in_smem ? tbytes : 0);
  return g;
}

}  // namespace lsspa

using namespace lsspa;

extern "C" int64_t lsspa_tsqr_slot_doubles(int p) {
  if (p < 1) return 0;
  return (int64_t)(p + 1) * (p + 1) + kSlotExtra;
}

extern "C" int lsspa_tsqr_num_parts(int p, int64_t nrows) {
  if (p < 1 || nrows < 1) return 0;
  const TsqrGeom g0 = tsqr_geom(p);
  const RegsGeom rg = regs_geom(p);
  TsqrGeom g = g0;
  if (rg.halves) {
    g.rb = 32 * rg.halves;
    g.smem = rg.smem;
    g.t_in_smem = rg.t_in_smem;
  }
  const DeviceInfo &d = device_info();
  const int sms = d.sm_count > 0 ? d.sm_count : 148;
  int64_t cap = g.t_in_smem ? (int64_t)sms * ((g.smem * 2 + 2048 <= 227 * 1024) ? 2 : 1) : sms;
  if (const char *e = getenv("LSSPA_TSQR_CTAS_PER_SM")) cap = (int64_t)sms * atoi(e);
  int64_t want = ceil_div(nrows, (int64_t)g.rb * 4);  // at least ~4 blocks per CTA
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

extern "C" int lsspa_tsqr_rows(const double *X, int64_t ldx, const double *y, int64_t nrows, int p,
                               double divisor, double *parts, int nparts, void *stream) {
  if (!X || !y || !parts || p < 1 || nrows < 1 || ldx < p || nparts < 1 || divisor == 0.0)
    return LSSPA_E_BADARG;
  const TsqrGeom g = tsqr_geom(p);
  RowsParams a;
  a.X = X;
  a.ldx = ldx;
  a.y = y;
  a.nrows = nrows;
  a.p = p;
  a.divisor = divisor;
  a.parts = parts;
  a.slot = lsspa_tsqr_slot_doubles(p);
  a.nparts = nparts;
  a.rb = g.rb;
  a.t_in_smem = g.t_in_smem ? 1 : 0;
  const RegsGeom rg = regs_geom(p);
  if (rg.halves) {
    a.t_in_smem = rg.t_in_smem ? 1 : 0;
    if (rg.halves == 2) {
      LSSPA_CUDA_TRY(cudaFuncSetAttribute(tsqr_rows_regs_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)rg.smem));
      tsqr_rows_regs_kernel<2><<<nparts, rg.threads, rg.smem, as_stream(stream)>>>(a);
    } else {
      LSSPA_CUDA_TRY(cudaFuncSetAttribute(tsqr_rows_regs_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)rg.smem));
      tsqr_rows_regs_kernel<1><<<nparts, rg.threads, rg.smem, as_stream(stream)>>>(a);
    }
    LSSPA_LAUNCH_CHECK();
    return LSSPA_OK;
  }
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(tsqr_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)g.smem));
  tsqr_rows_kernel<<<nparts, g.threads, g.smem, as_stream(stream)>>>(a);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

extern "C" int lsspa_tsqr_merge(const double *parts, int count, int group, int p, double *out,
                                void *stream) {
  if (!parts || !out || p < 1 || count < 1 || group < 1) return LSSPA_E_BADARG;
  const TsqrGeom g = tsqr_geom(p);
  MergeParams a;
  a.parts = parts;
  a.count = count;
  a.group = group;
  a.p = p;
  a.out = out;
  a.slot = lsspa_tsqr_slot_doubles(p);
  a.rb = g.rb;
  a.t_in_smem = g.t_in_smem ? 1 : 0;
  const int nout = (int)ceil_div(count, group);
  const RegsGeom rg = regs_geom(p);
  if (rg.halves) {
    a.t_in_smem = rg.t_in_smem ? 1 : 0;
    if (rg.halves == 2) {
      LSSPA_CUDA_TRY(cudaFuncSetAttribute(tsqr_merge_regs_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)rg.smem));
      tsqr_merge_regs_kernel<2><<<nout, rg.threads, rg.smem, as_stream(stream)>>>(a);
    } else {
      LSSPA_CUDA_TRY(cudaFuncSetAttribute(tsqr_merge_regs_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)rg.smem));
      tsqr_merge_regs_kernel<1><<<nout, rg.threads, rg.smem, as_stream(stream)>>>(a);
    }
    LSSPA_LAUNCH_CHECK();
    return LSSPA_OK;
  }
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(tsqr_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)g.smem));
  tsqr_merge_kernel<<<nout, g.threads, g.smem, as_stream(stream)>>>(a);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}
