// CholeskyQR2 reduction of [X | y] to its (p+1) x (p+1) triangular factor, p <= 119.
//
// Second implementation of reduce_data (reference ls_spa/ls_spa.py:290-318) next to the Householder
// TSQR of reduce.cu.  The Householder kernel is bound by the latency of one reflector per column
// per 64-row block (~2000 cycles each); here almost all work is dense 8x8x4 fp64 tensor products:
//
//   pass 1   G1 = Z^T Z           (Z = [X | y], streamed once)      R1 = chol(G1)
//   pass 2   G2 = Q1^T Q1, Q1 = Z R1^-1 formed on the fly per 32-row chunk      R2 = chol(G2)
//   result   R = R2 R1
//
// One Cholesky-QR pass loses cond(Z)^2 eps; repeating it on Q1 = Z R1^-1 (whose condition number is
// ~1) restores Householder-level accuracy as long as cond(Z) <~ 1e7 (Yamamoto et al. 2015).  The
// caller checks the condition estimate / pivot status written by lsspa_chol_factor and falls back to
// the Householder TSQR otherwise (rank-deficient inputs such as the reference's own "hard" test data).
//
// Fragments: chunk rows live in shared memory row-major with row stride ldr (ldr % 16 == 4), so
// the access "rows 4s+q, columns 8t+c" (lane = 4c + q) is bank-conflict free; the same fragment
// f_t is the A operand of tile row t and the B operand of tile column t of the Gram product.

#include "common.cuh"

namespace lsspa {

constexpr int kGR = 32;          // rows per chunk
constexpr int kGramThreads = 256;

__device__ __forceinline__ void dmma_g(double &d0, double &d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

static int gram_nt(int p) { return (p + 1 + 7) / 8; }
static int gram_ldr(int nt) {
  int n = 8 * nt;
  return n + ((4 - (n % 16)) + 16) % 16;   // smallest >= n with % 16 == 4
}

struct GramParams {
  const double *X;
  int64_t ldx;
  const double *y;
  int64_t nrows;
  int p;
  const double *Rinv;  // pass 2: [8nt][ldr] row-major upper triangular, else nullptr
  double *parts;       // [nparts][(8nt)^2] row-major partial Gram matrices (upper tiles valid)
  int nparts;
  int nt;
  int ldr;
};

// accumulate the Gram of the kGR x 8nt chunk C (row-major, stride ldr) into the warp's tiles:
// warp g owns tile rows g and i2 = nt-1-g, i.e. the nt+1 tiles (g, g..nt-1) and (i2, i2..nt-1)
// (every group has the same count: balanced).  acc[idx]: idx < nt-g -> tile (g, g+idx), else
// tile (i2, i2 + idx - (nt-g)).
template <int MAXNT>
__device__ __forceinline__ void gram_chunk(const double *C, int ldr, int nt, int g, int lane,
                                           double (&acc)[MAXNT + 1][2]) {
  const int c = lane >> 2, q = lane & 3;
  const int i2 = nt - 1 - g, nA = nt - g;
  const bool two = i2 > g;
#pragma unroll 2
  for (int ks = 0; ks < kGR / 4; ++ks) {
    const double *row = C + (size_t)(4 * ks + q) * ldr + c;
    const double fa = row[8 * g], fb = row[8 * i2];
#pragma unroll
    for (int idx = 0; idx <= MAXNT; ++idx) {
      if (idx < nA) {
        dmma_g(acc[idx][0], acc[idx][1], fa, row[8 * (g + idx)]);
      } else if (two && idx <= nt) {
        dmma_g(acc[idx][0], acc[idx][1], fb, row[8 * (i2 + idx - nA)]);
      }
    }
  }
}

template <int MAXNT, int MINB>
__global__ void __launch_bounds__(kGramThreads, MINB) gram_rows_kernel(GramParams a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int p = a.p, nt = a.nt, ldr = a.ldr, nc = 8 * nt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = lane >> 2, q = lane & 3;
  const bool pass2 = a.Rinv != nullptr;
  double *sm = reinterpret_cast<double *>(smem_raw);
  double *Zc0 = sm;                               // kGR x ldr
  double *Zc1 = Zc0 + (size_t)kGR * ldr;
  double *Qc = Zc1 + (size_t)kGR * ldr;           // pass 2 only
  double *Ri = Qc + (size_t)kGR * ldr;            // pass 2 only: nc x ldr
  if (pass2)
    for (int e = tid; e < nc * ldr; e += kGramThreads) Ri[e] = a.Rinv[e];

  double acc[MAXNT + 1][2];
#pragma unroll
  for (int t = 0; t <= MAXNT; ++t) acc[t][0] = acc[t][1] = 0.0;
  const int ngroups = (nt + 1) / 2;
  const bool has_group = warp < ngroups;

  const int64_t per = ceil_div(ceil_div(a.nrows, (int64_t)a.nparts), (int64_t)kGR) * kGR;
  const int64_t r_begin = (int64_t)blockIdx.x * per;
  const int64_t r_end = (r_begin + per < a.nrows) ? r_begin + per : a.nrows;

  // staging: thread (rr = tid / 128, col = tid % 128) moves rows 2u + rr, u = 0..15, of column col
  constexpr int NPT = kGR / 2;
  const int rr = tid >> 7, col = tid & 127;
  const bool live = col < nc;
  double stage[NPT];
  auto fetch = [&](int64_t r0) {
#pragma unroll
    for (int u = 0; u < NPT; ++u) {
      const int64_t r = r0 + 2 * u + rr;
      double v = 0.0;
      if (live && r < r_end) {
        if (col < p) v = a.X[r * a.ldx + col];
        else if (col == p) v = a.y[r];
      }
      stage[u] = v;
    }
  };
  auto commit = [&](double *Z) {
    if (live) {
#pragma unroll
      for (int u = 0; u < NPT; ++u) Z[(size_t)(2 * u + rr) * ldr + col] = stage[u];
    }
  };
  if (r_begin < r_end) {
    fetch(r_begin);
    commit(Zc0);
  }
  __syncthreads();
  int buf = 0;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += kGR) {
    double *Z = buf ? Zc1 : Zc0;
    double *Zn = buf ? Zc0 : Zc1;
    const bool more = r0 + kGR < r_end;
    if (more) fetch(r0 + kGR);
    const double *G_in = Z;
    if (pass2) {
      // Q = Z Rinv for this chunk: warp w owns output column tiles j = w, w+8, all 4 row tiles
      for (int j = warp; j < nt; j += kGramThreads / 32) {
        double o[kGR / 8][2];
#pragma unroll
        for (int rt = 0; rt < kGR / 8; ++rt) o[rt][0] = o[rt][1] = 0.0;
        for (int kt = 0; kt <= j; ++kt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double b = Ri[(size_t)(8 * kt + 4 * e + q) * ldr + 8 * j + c];
#pragma unroll
            for (int rt = 0; rt < kGR / 8; ++rt) {
              const double av = Z[(size_t)(8 * rt + c) * ldr + 8 * kt + 4 * e + q];
              dmma_g(o[rt][0], o[rt][1], av, b);
            }
          }
        }
#pragma unroll
        for (int rt = 0; rt < kGR / 8; ++rt)
          *reinterpret_cast<double2 *>(Qc + (size_t)(8 * rt + c) * ldr + 8 * j + 2 * q) = make_double2(o[rt][0], o[rt][1]);
      }
      __syncthreads();
      G_in = Qc;
    }
    if (has_group) gram_chunk<MAXNT>(G_in, ldr, nt, warp, lane, acc);
    if (more) commit(Zn);
    __syncthreads();
    buf ^= 1;
  }
  // write this CTA's partial Gram (upper tiles): C layout lane (m = c, n = 2q+e)
  double *out = a.parts + (size_t)blockIdx.x * nc * nc;
  for (int e = tid; e < nc * nc; e += kGramThreads) out[e] = 0.0;
  __syncthreads();
  if (has_group) {
    const int g = warp, i2 = nt - 1 - g, nA = nt - g;
#pragma unroll
    for (int idx = 0; idx <= MAXNT; ++idx) {
      if (idx < nA)
        *reinterpret_cast<double2 *>(out + (size_t)(8 * g + c) * nc + 8 * (g + idx) + 2 * q) = make_double2(acc[idx][0], acc[idx][1]);
      else if (i2 > g && idx <= nt)
        *reinterpret_cast<double2 *>(out + (size_t)(8 * i2 + c) * nc + 8 * (i2 + idx - nA) + 2 * q) = make_double2(acc[idx][0], acc[idx][1]);
    }
  }
}

// G[e] = scale * sum_cta parts[cta][e], fixed order (deterministic)
__global__ void gram_sum_kernel(const double *parts, int count, int n2, double scale, double *G) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n2) return;
  double s = 0.0;
  for (int k = 0; k < count; ++k) s += parts[(size_t)k * n2 + e];
  G[e] = s * scale;
}

// Cholesky G = R^T R (upper, G row-major nc x nc, leading q x q block used), then R^-1.
// info[0] = 0 ok / 1 non-positive or tiny pivot (relative to the column's own norm),
// info[1] = a bound >= cond_2(R') of the column-equilibrated factor R' = R D^-1 (the smaller of
// |R'|_F |R'^-1|_F and sqrt(max row sum |G'|) sqrt(|R'^-1|_1 |R'^-1|_inf)),
// D = diag(sqrt(G_jj)): the accuracy of a Cholesky factor is governed by the conditioning of the
// equilibrated matrix (van der Sluis / Demmel), so units of the columns must not count.
// R is written row-major q x q (slot layout of reduce.cu); Rinv row-major [nc][ldr], zero padded.
__global__ void __launch_bounds__(1024) chol_factor_kernel(const double *G, int q, int nc, int ldr, double *R_out,
                                                           double *Rinv_out, double *info) {
  extern __shared__ double sm[];
  double *A = sm;                        // nc x nc working copy (upper part), row-major
  double *Vi = A + (size_t)nc * nc;      // nc x nc inverse, row-major
  __shared__ double rk[128];             // scaled pivot row of the current column
  __shared__ double s_fail, red[64];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int tx = tid & 31, ty = tid >> 5, lane = tx, w = ty;
  for (int e = tid; e < nc * nc; e += nt) {
    const int i = e / nc, j = e - i * nc;
    A[e] = (i < q && j < q && j >= i) ? G[e] : 0.0;
    Vi[e] = 0.0;
  }
  if (tid == 0) s_fail = 0.0;
  __syncthreads();
  double dmax = 0.0;
  for (int i = 0; i < q; ++i) dmax = fmax(dmax, A[(size_t)i * nc + i]);
  // right-looking Cholesky, two barriers per column: (a) the pivot row is scaled into rk,
  // (b) the trailing block is updated by a 32 x 32 thread grid (no index divisions)
  for (int k = 0; k < q; ++k) {
    const double d = A[(size_t)k * nc + k];
    const double r = sqrt(d > 0.0 ? d : 1.0), ri = 1.0 / r;
    if (tid == 0 && (!(d > 1e-14 * G[(size_t)k * nc + k]) || !(dmax > 0.0))) s_fail = 1.0;
    const int m = q - k - 1;
    for (int j = tid; j < m; j += nt) rk[j] = A[(size_t)k * nc + k + 1 + j] * ri;
    __syncthreads();
    if (tid == 0) A[(size_t)k * nc + k] = r;
    for (int j = tid; j < m; j += nt) A[(size_t)k * nc + k + 1 + j] = rk[j];
    for (int i = ty; i < m; i += 32) {
      const double ai = rk[i];
      double *row = A + (size_t)(k + 1 + i) * nc + k + 1;
      for (int j = i + tx; j < m; j += 32) row[j] = fma(-ai, rk[j], row[j]);   // upper part: j >= i
    }
    __syncthreads();
  }
  // inverse: warp w solves R x = e_j for its columns j by back substitution, the row dot products
  // across the lanes (column j of R^-1 lives in column j of Vi)
  for (int j = w; j < q; j += 32) {
    for (int i = j; i >= 0; --i) {
      const double *row = A + (size_t)i * nc;
      double sacc = 0.0;
      for (int k = i + 1 + lane; k <= j; k += 32) sacc = fma(-row[k], Vi[(size_t)k * nc + j], sacc);
      sacc = warp_sum(sacc);
      if (lane == 0) Vi[(size_t)i * nc + j] = (sacc + ((i == j) ? 1.0 : 0.0)) / row[i];
      __syncwarp();
    }
  }
  __syncthreads();
  double fr = 0.0, fi = 0.0;
  for (int e = tid; e < nc * nc; e += nt) {
    const int i = e / nc, j = e - i * nc;
    if (i < q && j < q) {
      const double di = sqrt(fmax(G[(size_t)i * nc + i], 0.0)), dj = sqrt(fmax(G[(size_t)j * nc + j], 0.0));
      const double rr = (dj > 0.0) ? A[e] / dj : 0.0;
      const double v = Vi[e] * di;
      fr = fma(rr, rr, fr);
      fi = fma(v, v, fi);
    }
  }
  fr = warp_sum(fr);
  fi = warp_sum(fi);
  if (lane == 0) {
    red[w] = fr;
    red[32 + w] = fi;
  }
  __syncthreads();
  // second bound: sqrt(Gershgorin bound on the largest eigenvalue of the equilibrated Gram matrix)
  // * sqrt(|R'^-1|_1 |R'^-1|_inf) with R'^-1 = D R^-1; thread t takes row / column t
  __shared__ double b_g[128], b_c[128], b_r[128];
  for (int t = tid; t < q; t += nt) {
    const double dt = sqrt(fmax(G[(size_t)t * nc + t], 0.0));
    double g = 0.0, cs = 0.0, rs = 0.0;
    for (int k = 0; k < q; ++k) {
      const double dk = sqrt(fmax(G[(size_t)k * nc + k], 0.0));
      const double gij = (k >= t) ? G[(size_t)t * nc + k] : G[(size_t)k * nc + t];   // upper part of G
      g += (dt > 0.0 && dk > 0.0) ? fabs(gij) / (dt * dk) : 0.0;
      cs += fabs(Vi[(size_t)k * nc + t]) * dk;     // column t of D R^-1
      rs += fabs(Vi[(size_t)t * nc + k]) * dt;     // row t
    }
    b_g[t] = g;
    b_c[t] = cs;
    b_r[t] = rs;
  }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, b = 0.0, gm = 0.0, cm = 0.0, rm = 0.0;
    for (int k = 0; k < nt / 32; ++k) {
      a += red[k];
      b += red[32 + k];
    }
    for (int t = 0; t < q; ++t) {
      gm = fmax(gm, b_g[t]);
      cm = fmax(cm, b_c[t]);
      rm = fmax(rm, b_r[t]);
    }
    double frob = sqrt(a) * sqrt(b), sharp = sqrt(gm) * sqrt(cm * rm);
    if (!(frob == frob)) frob = INFINITY;
    if (!(sharp == sharp)) sharp = INFINITY;
    info[0] = s_fail;
    info[1] = fmin(frob, sharp);
  }
  for (int e = tid; e < q * q; e += nt) {
    const int i = e / q, j = e - i * q;
    R_out[e] = A[(size_t)i * nc + j];
  }
  for (int e = tid; e < nc * ldr; e += nt) {
    const int i = e / ldr, j = e - i * ldr;
    Rinv_out[e] = (j < nc) ? Vi[(size_t)i * nc + j] : 0.0;
  }
}

// out (slot layout: q x q row-major, then [q*q] = sum of squares of the y column) = R2 R1
__global__ void tri_product_kernel(const double *R2, const double *R1, int q, const double *G1, int nc,
                                   double *out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < q * q) {
    const int i = e / q, j = e - i * q;
    double s = 0.0;
    for (int k = i; k <= j; ++k) s = fma(R2[(size_t)i * q + k], R1[(size_t)k * q + j], s);
    out[e] = (j >= i) ? s : 0.0;
  }
  if (e < 8) out[(size_t)q * q + e] = (e == 0) ? G1[(size_t)(q - 1) * nc + (q - 1)] : 0.0;
}

template <int MAXNT>
static int launch_gram(const GramParams &a, size_t smem, cudaStream_t st) {
  if (a.Rinv == nullptr) {   // pass 1: small shared-memory footprint, two CTAs per SM
    LSSPA_CUDA_TRY(cudaFuncSetAttribute(gram_rows_kernel<MAXNT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gram_rows_kernel<MAXNT, 2><<<a.nparts, kGramThreads, smem, st>>>(a);
    LSSPA_LAUNCH_CHECK();
    return LSSPA_OK;
  }
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(gram_rows_kernel<MAXNT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gram_rows_kernel<MAXNT, 1><<<a.nparts, kGramThreads, smem, st>>>(a);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

}  // namespace lsspa

using namespace lsspa;

// (p + 1 <= 120: the Cholesky kernel keeps the factor and its inverse, 2 x (8 nt)^2 doubles, in shared memory)
extern "C" int lsspa_gram_supported(int p) { return (p >= 1 && p + 1 <= 120) ? 1 : 0; }

extern "C" int64_t lsspa_gram_slot_doubles(int p) {
  if (!lsspa_gram_supported(p)) return 0;
  const int nc = 8 * gram_nt(p);
  return (int64_t)nc * nc;
}

extern "C" int64_t lsspa_gram_rinv_doubles(int p) {
  if (!lsspa_gram_supported(p)) return 0;
  const int nt = gram_nt(p);
  return (int64_t)8 * nt * gram_ldr(nt);
}

extern "C" int lsspa_gram_num_parts(int p, int64_t nrows, int pass2) {
  if (!lsspa_gram_supported(p) || nrows < 1) return 0;
  const DeviceInfo &d = device_info();
  const int sms = d.sm_count > 0 ? d.sm_count : 148;
  int64_t cap = (int64_t)sms * (pass2 ? 1 : 2);  // pass 2 keeps R1^-1 in shared memory: one CTA per SM
  int64_t want = ceil_div(nrows, (int64_t)kGR * 4);
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

extern "C" int lsspa_gram_rows(const double *X, int64_t ldx, const double *y, int64_t nrows, int p,
                               const double *Rinv_or_null, double *parts, int nparts, void *stream) {
  if (!X || !y || !parts || !lsspa_gram_supported(p) || nrows < 1 || ldx < p || nparts < 1) return LSSPA_E_BADARG;
  GramParams a;
  a.X = X;
  a.ldx = ldx;
  a.y = y;
  a.nrows = nrows;
  a.p = p;
  a.Rinv = Rinv_or_null;
  a.parts = parts;
  a.nparts = nparts;
  a.nt = gram_nt(p);
  a.ldr = gram_ldr(a.nt);
  size_t smem = (size_t)2 * kGR * a.ldr * sizeof(double);
  if (Rinv_or_null) smem += ((size_t)kGR * a.ldr + (size_t)8 * a.nt * a.ldr) * sizeof(double);
  cudaStream_t st = as_stream(stream);
  if (a.nt <= 4) return launch_gram<4>(a, smem, st);
  if (a.nt <= 8) return launch_gram<8>(a, smem, st);
  if (a.nt <= 13) return launch_gram<13>(a, smem, st);
  return launch_gram<16>(a, smem, st);
}

extern "C" int lsspa_gram_finish(const double *parts, int count, int p, double scale, double *G_out,
                                 void *stream) {
  if (!parts || !G_out || !lsspa_gram_supported(p) || count < 1) return LSSPA_E_BADARG;
  const int n2 = (int)lsspa_gram_slot_doubles(p);
  gram_sum_kernel<<<(n2 + 255) / 256, 256, 0, as_stream(stream)>>>(parts, count, n2, scale, G_out);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

// G (layout of lsspa_gram_finish) += reg on the first p diagonal entries: the sqrt(reg) I rows of the train
// block (reference ls_spa/ls_spa.py:310) contribute exactly reg I to the Gram matrix
__global__ void gram_ridge_kernel(double *G, int nc, int p, double reg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p) G[(size_t)i * nc + i] += reg;
}

extern "C" int lsspa_gram_add_ridge(double *G, int p, double reg, void *stream) {
  if (!G || !lsspa_gram_supported(p)) return LSSPA_E_BADARG;
  gram_ridge_kernel<<<(p + 127) / 128, 128, 0, as_stream(stream)>>>(G, 8 * gram_nt(p), p, reg);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

extern "C" int lsspa_chol_factor(const double *G, int p, double *R_out, double *Rinv_out, double *info,
                                 void *stream) {
  if (!G || !R_out || !Rinv_out || !info || !lsspa_gram_supported(p)) return LSSPA_E_BADARG;
  const int nt = gram_nt(p), nc = 8 * nt, ldr = gram_ldr(nt);
  const size_t smem = (size_t)2 * nc * nc * sizeof(double);
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(chol_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  chol_factor_kernel<<<1, 1024, smem, as_stream(stream)>>>(G, p + 1, nc, ldr, R_out, Rinv_out, info);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

extern "C" int lsspa_tri_product(const double *R2, const double *R1, int p, const double *G1, double *out_slot,
                                 void *stream) {
  if (!R2 || !R1 || !G1 || !out_slot || !lsspa_gram_supported(p)) return LSSPA_E_BADARG;
  const int q = p + 1, nc = 8 * gram_nt(p);
  tri_product_kernel<<<(q * q + 255) / 256, 256, 0, as_stream(stream)>>>(R2, R1, q, G1, nc, out_slot);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}
