// Per-permutation core of LS-SPA for WELL-CONDITIONED reduced problems: the triangular factor of
// R_tr[:, perm] (reference ls_spa/ls_spa.py:268, np.linalg.qr) is obtained as the Cholesky factor
// of the permuted Gram matrix instead of by Householder reflections.
//
//   Gh = [R_tr | c_tr]^T [R_tr | c_tr]           (p+1) x (p+1), once per reduced problem
//   Gh[pi^, pi^] = [R c]^T [R c]                  with pi^ = (perm, p):  R = chol, c = R^-T b
//
// R agrees with the QR factor up to row signs, which the lifts do not see.  The factorisation
// is a left-looking blocked Cholesky on 8x8 tiles: every product is a DMMA, the only serial
// part is the 8x8 diagonal block (8 pivots instead of the 8 full-height reflectors a Householder
// panel needs), and the inverse of the diagonal block -- which the elimination phase needs
// anyway -- falls out of the same pivots.  Forward error ~ eps * cond(R_tr)^2, so the host only
// takes this route when lsspa_lifts_gram reports a small condition estimate; otherwise the
// Householder kernel (lifts_mma.cu) runs.  Phase 2 (elimination of X = R_te[:, perm], reference
// :279-283) is the same as in lifts_mma.cu.
//
// Fragment conventions as in lifts_mma.cu (lane = 4c + q): A[m=c][k=q], B[k=q][n=c],
// C[m=c][n=2q+e]; "tile access" = lane touches M[i0+2q..+1][j0+c] of the column-major matrix.

#include "common.cuh"

#include <stdlib.h>

namespace lsspa {
namespace {

struct CholParams {
  int p;
  int ld;
  int rt;   // row tiles      ceil(p / 8)
  int pt;   // column tiles   ceil((p + 1) / 8)
  const double *Gh;   // (p+1) x (p+1) symmetric, leading dimension p+1
  const double *Rte;  // column-major p x p
  const double *cte;
  double inv_ynsq;
  const int32_t *perms;
  int64_t count;
  int anti;
  double *out;
  double *fact;        // split route: per-evaluation factors (R upper tiles, c, inverses of the diagonal blocks)
  long long *dbg;  // optional cycle counters of block 0 (development aid), else nullptr
};

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 ld_tile(const double *M, int ld, int i0, int j0, int c, int q) {
  return *reinterpret_cast<const double2 *>(M + (size_t)(j0 + c) * ld + i0 + 2 * q);
}
__device__ __forceinline__ void st_tile(double *M, int ld, int i0, int j0, int c, int q, double2 v) {
  *reinterpret_cast<double2 *>(M + (size_t)(j0 + c) * ld + i0 + 2 * q) = v;
}

// 1/d and 1/sqrt(d) from the 20-bit MUFU seeds and one cubic correction step each (relative
// error ~ seed^3 ~ 1e-18 before the final rounding): 3 dependent fp64 operations after the seed
// instead of the ~7 of the IEEE-exact library sequences.  These sit on the pivot chain.
__device__ __forceinline__ double rcp_fast(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  const double e = fma(-d, r, 1.0);
  return fma(r, fma(e, e, e), r);
}
__device__ __forceinline__ double rsqrt_fast(double d) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  const double e = fma(-d * r, r, 1.0);
  return fma(r * e, fma(0.375, e, 0.5), r);
}

// Pivots outside this range (or NaN) are not factored: the host never sends such a problem here,
// the guard only keeps the arithmetic finite.
constexpr double kPivMin = 1e-200, kPivMax = 1e200;

// In-register factorisation of the diagonal tile.  (t0, t1) = T[m=c][n=2q+e], symmetric, only the
// part m >= n is used.  nf pivots.  On return (t0, t1) = Lo = U^T (lower triangular Cholesky
// factor; rows m >= nf carry the finished entries of the non-pivot rows, i.e. the c row), and
// (y0, y1) = U^-1 in the layout Dbuf wants: lane (c, q) holds Uinv[2q+e][c].
__device__ __forceinline__ void diag_factor(double &t0, double &t1, double &y0, double &y1, int nf, int lane) {
  const int c = lane >> 2, q = lane & 3;
  y0 = (c == 2 * q) ? 1.0 : 0.0;
  y1 = (c == 2 * q + 1) ? 1.0 : 0.0;
  double ds0 = 0.0, ds1 = 0.0, dsc = 0.0;  // pivots of columns 2q, 2q+1 and c (their rsqrt is needed at the end)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (j < nf) {
      const double sel = (j & 1) ? t1 : t0;         // this lane's entry of column (2q' + (j&1)); column j on lanes q == j/2
      const int jq = j >> 1;
      const double d = __shfl_sync(kFull, sel, 4 * j + jq);
      const double colm = __shfl_sync(kFull, sel, (lane & ~3) | jq);      // T[c][j]
      const double cn0 = __shfl_sync(kFull, sel, (2 * q) * 4 + jq);       // T[2q][j]
      const double cn1 = __shfl_sync(kFull, sel, (2 * q + 1) * 4 + jq);   // T[2q+1][j]
      const double yj0 = __shfl_sync(kFull, y0, 4 * j + q);               // Y[j][2q]
      const double yj1 = __shfl_sync(kFull, y1, 4 * j + q);               // Y[j][2q+1]
      const bool ok = (d > kPivMin) && (d < kPivMax);
      const double ri = ok ? rcp_fast(d) : 0.0;
      if (2 * q == j) ds0 = d;
      if (2 * q + 1 == j) ds1 = d;
      if (c == j) dsc = d;
      if (c > j) {
        const double f = colm * ri;
        if (2 * q > j) t0 = fma(-f, cn0, t0);
        if (2 * q + 1 > j) t1 = fma(-f, cn1, t1);
        y0 = fma(-f, yj0, y0);
        y1 = fma(-f, yj1, y1);
      }
    }
  }
  const double rs0 = (ds0 > kPivMin && ds0 < kPivMax) ? rsqrt_fast(ds0) : 0.0;
  const double rs1 = (ds1 > kPivMin && ds1 < kPivMax) ? rsqrt_fast(ds1) : 0.0;
  const double rsc = (dsc > kPivMin && dsc < kPivMax) ? rsqrt_fast(dsc) : 0.0;
  // Lo[m][n] = T_n[m][n] * rsqrt(d_n) for m >= n, n < nf
  t0 = (c >= 2 * q && 2 * q < nf) ? t0 * rs0 : 0.0;
  t1 = (c >= 2 * q + 1 && 2 * q + 1 < nf) ? t1 * rs1 : 0.0;
  // Uinv[u][jj] = Y[jj][u] * rsqrt(d_jj); columns jj >= nf are zero
  y0 = (c < nf) ? y0 * rsc : 0.0;
  y1 = (c < nf) ? y1 * rsc : 0.0;
}

// Rows 8 it .. 8 it + 7 of X = R_te[:, perm] in C / A-operand layout (lane (c, q): row 8 it + c,
// columns 8 L + 2q + e) and the test target of the row.  Every load is issued from a valid (clamped)
// address and masked afterwards, so all of them are in flight together.
template <int RT>
__device__ __forceinline__ void load_x(double (&xr)[RT][2], double &r_in, const CholParams &a, const int *perm_s,
                                       int it, int p, int c, int q) {
  const int row = 8 * it + c;
  const int rowc = row < p ? row : p - 1;
#pragma unroll
  for (int L = 0; L < RT; ++L) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int l = 8 * L + 2 * q + e;
      const int col = perm_s[l < p ? l : p - 1];
      const double v = __ldg(a.Rte + (size_t)col * p + rowc);
      xr[L][e] = (l < p && row <= col) ? v : 0.0;
    }
  }
  const double t = __ldg(a.cte + rowc);
  r_in = (row < p) ? t : 0.0;
}

// Elimination step J of one 8-row tile of X (reference :279-283): M_J = X_J R_JJ^-1 are 8 columns
// of W = X R^-1; the running test residual of each row after each of these columns gives
// cost_{k+1} (summed over the rows into wc); X_L -= M_J R_JL for the column tiles to the right.
// Needs row block J of R, Dinv_J and c_J only -- i.e. it can run as soon as step J of the
// factorisation is complete.
template <int RT, int LD, int J>
__device__ __forceinline__ void elim_step(double (&xr)[RT][2], double &r_in, double *wc, const double *A,
                                          const double *Dbuf, const double *cvec, int p, int c, int q) {
  const double2 dv = ld_tile(Dbuf + J * 64, 8, 0, 0, c, q);
  double m0 = 0.0, m1 = 0.0;  // C layout (row, column 2q+e of the tile)
  dmma(m0, m1, xr[J][0], dv.x);
  dmma(m0, m1, xr[J][1], dv.y);
  // residual of row c after each of the 8 columns of the tile as one more product:
  // R[m][n] = r_in[m] - sum_{k <= n} M[m][k] c_k, i.e. M times the upper-triangular matrix whose
  // row k holds c_k (B fragment: k = 2q + e, n = c); plain fp64 arithmetic would queue behind the
  // DMMAs of the other warps operation by operation
  const double2 cv = *reinterpret_cast<const double2 *>(cvec + 8 * J + 2 * q);
  double ra = r_in, rb = r_in;
  dmma(ra, rb, -m0, (2 * q <= c) ? cv.x : 0.0);
  dmma(ra, rb, -m1, (2 * q + 1 <= c) ? cv.y : 0.0);
  // sum of the squares over the 8 rows (lane bits 2..4) of both values with three shuffles: the
  // first round hands each value to one half of the lanes; even rows end up with the total of the
  // first column, odd rows with that of the second
  const double d0 = ra * ra, d1 = rb * rb;
  const bool odd = (c & 1) != 0;
  double tot = (odd ? d1 : d0) + __shfl_xor_sync(kFull, odd ? d0 : d1, 4);
  tot += __shfl_xor_sync(kFull, tot, 8);
  tot += __shfl_xor_sync(kFull, tot, 16);
  if (c < 2) {
    const int k0 = 8 * J + 2 * q + c;
    if (k0 < p) wc[k0] += tot;          // wc[k] collects cost_{k+1}
  }
  r_in = __shfl_sync(kFull, rb, 3, 4);   // residual after the last column of the tile
  m0 = -m0;
  m1 = -m1;
#pragma unroll
  for (int L = J + 1; L < RT; ++L) {
    const double2 rt = ld_tile(A, LD, 8 * J, 8 * L, c, q);
    dmma(xr[L][0], xr[L][1], m0, rt.x);
    dmma(xr[L][0], xr[L][1], m1, rt.y);
  }
}

// run-time step index -> the statically indexed instance (the tile registers need constant indices)
template <int RT, int LD, int J = 0>
__device__ __forceinline__ void elim_dispatch(int s, double (&xr)[RT][2], double &r_in, double *wc, const double *A,
                                              const double *Dbuf, const double *cvec, int p, int c, int q) {
  if constexpr (J < RT) {
    if (s == J) elim_step<RT, LD, J>(xr, r_in, wc, A, Dbuf, cvec, p, c, q);
    else elim_dispatch<RT, LD, J + 1>(s, xr, r_in, wc, A, Dbuf, cvec, p, c, q);
  }
}

template <int RT, int LD, int J = 0>
__device__ __forceinline__ void elim_all(double (&xr)[RT][2], double &r_in, double *wc, const double *A,
                                         const double *Dbuf, const double *cvec, int p, int c, int q) {
  if constexpr (J < RT) {
    elim_step<RT, LD, J>(xr, r_in, wc, A, Dbuf, cvec, p, c, q);
    elim_all<RT, LD, J + 1>(xr, r_in, wc, A, Dbuf, cvec, p, c, q);
  }
}

// sum_{k < kend} R(k, L)^T R(k, S) in C layout (the transpose of what tile (S, L) loses): four
// independent accumulation chains, a dependent DMMA costs more than an issue slot
template <int LD>
__device__ __forceinline__ double2 acc_tile(const double *A, int kend, int L, int S, int c, int q) {
  double p0 = 0.0, p1 = 0.0, r0 = 0.0, r1 = 0.0, e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0;
  int k = 0;
  for (; k + 1 < kend; k += 2) {
    const double2 xa = ld_tile(A, LD, 8 * k, 8 * L, c, q);
    const double2 xb = ld_tile(A, LD, 8 * k, 8 * S, c, q);
    const double2 ya = ld_tile(A, LD, 8 * k + 8, 8 * L, c, q);
    const double2 yb = ld_tile(A, LD, 8 * k + 8, 8 * S, c, q);
    dmma(p0, p1, xa.x, xb.x);
    dmma(r0, r1, ya.x, yb.x);
    dmma(e0, e1, xa.y, xb.y);
    dmma(f0, f1, ya.y, yb.y);
  }
  if (k < kend) {
    const double2 xa = ld_tile(A, LD, 8 * k, 8 * L, c, q);
    const double2 xb = ld_tile(A, LD, 8 * k, 8 * S, c, q);
    dmma(p0, p1, xa.x, xb.x);
    dmma(e0, e1, xa.y, xb.y);
  }
  return make_double2((p0 + r0) + (e0 + f0), (p1 + r1) + (e1 + f1));
}

// leading dimension of the tile array (column-major, ld % 16 == 8: conflict-free tile accesses)
__host__ __device__ constexpr int chol_ld(int rt) { return ((8 * rt) % 16 == 8) ? 8 * rt : 8 * rt + 8; }

// The tile geometry (RT row tiles, PT = RT or RT + 1 column tiles) is a template parameter: every
// tile address is then base + immediate and no tile loop carries a run-time guard.
// MODE 0: the whole evaluation (factor + eliminate, fused).  MODE 1: factor only -- needs only the
// train side; R (upper tiles, with c) and the inverses of its diagonal blocks go to a.fact.
// MODE 2: eliminate only -- loads them back and runs phase 2 against the test factor.  The split
// lets the factorisations of a host-resident job run while the test rows are still crossing PCIe.
__host__ __device__ constexpr int64_t chol_fact_doubles(int rt, int pt) {
  int64_t t = 0;
  for (int L = 0; L < pt; ++L) t += (L < rt ? L : rt - 1) + 1;
  return 64 * t + 64 * rt;
}

template <int RT, int PT, int MINB, int MODE>
__global__ void __launch_bounds__(256, MINB) lifts_chol_kernel(CholParams a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int p = a.p;
  constexpr int ld = chol_ld(RT);
  constexpr int NR = 8 * RT, NC = 8 * PT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = lane >> 2, q = lane & 3;

  double *A = reinterpret_cast<double *>(smem_raw);  // upper tiles of the permuted Gram matrix -> R, c
  double *Dbuf = A + (size_t)NC * ld;                 // RT x 64: inverses of the diagonal blocks of R
  double *wcost = Dbuf + (size_t)RT * 64;             // 8 x NR per-warp cost partials
  double *cost = wcost + (size_t)8 * NR;              // p + 2
  double *acc = cost + (p + 2);                       // p
  int *perm_s = reinterpret_cast<int *>(acc + p + (p & 1));  // p + 1

  const int halves = a.anti ? 2 : 1;
  const double weight = a.anti ? 0.5 : 1.0;
  const int ldg = p + 1;

  for (int64_t sidx = blockIdx.x; sidx < a.count; sidx += gridDim.x) {
    for (int h = 0; h < halves; ++h) {
      __syncthreads();
      for (int k = tid; k <= p; k += 256)
        perm_s[k] = (k == p) ? p : a.perms[sidx * p + (h == 0 ? k : p - 1 - k)];
      __syncthreads();
      const long long t_a = LSSPA_CLOCK();
      constexpr int64_t FD = chol_fact_doubles(RT, PT);
      const int64_t eval = sidx * halves + h;
      if constexpr (MODE == 2) {
        // ---- phase 0 (split route): R, c and the diagonal inverses of this evaluation from global
        // (tile column L holds 8 columns of 8 (min(L, RT-1) + 1) rows; every bound below is a
        // compile-time constant after unrolling, and all loads are issued before the first use)
        const double2 *src = reinterpret_cast<const double2 *>(a.fact + eval * FD);
#pragma unroll
        for (int L = 0; L < PT; ++L) {
          const int nrow2 = 4 * ((L < RT ? L : RT - 1) + 1);   // double2 per column
          int off2 = 0;
#pragma unroll
          for (int k = 0; k < L; ++k) off2 += 32 * ((k < RT ? k : RT - 1) + 1);
#pragma unroll
          for (int e0 = 0; e0 < 8 * nrow2; e0 += 256) {
            const int e = e0 + tid;
            if (e < 8 * nrow2) {
              const int cc = e / nrow2, r2 = e - cc * nrow2;
              *reinterpret_cast<double2 *>(A + (size_t)(8 * L + cc) * ld + 2 * r2) = __ldg(src + off2 + e);
            }
          }
        }
#pragma unroll
        for (int e0 = 0; e0 < 32 * RT; e0 += 256) {
          const int e = e0 + tid;
          if (e < 32 * RT) reinterpret_cast<double2 *>(Dbuf)[e] = __ldg(src + (FD - 64 * RT) / 2 + e);
        }
      } else {
      // ---- phase 0: gather the upper tiles of Gh[pi^, pi^] (rows < p, columns <= p), zero padding.
        // Warp w takes the columns l = w + 8 g (column tile g); lane -> rows 2 lane, 2 lane + 1 (+ 64).
        // All loads of a batch of column tiles are issued from valid addresses before the first store
        // and masked afterwards, so they are in flight together.
        {
          constexpr int NRR = (NR + 63) / 64;   // lane -> rows 2 lane + 64 r, r < NRR
          int pr[NRR][2];
          bool pv[NRR][2];
#pragma unroll
          for (int r = 0; r < NRR; ++r) {
            const int i0 = 2 * lane + 64 * r;
            pv[r][0] = i0 < p;
            pv[r][1] = i0 + 1 < p;
            pr[r][0] = pv[r][0] ? perm_s[i0] : 0;
            pr[r][1] = pv[r][1] ? perm_s[i0 + 1] : 0;
          }
          constexpr int GB = (PT + 1) / 2;
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            double vv[GB][NRR][2];
#pragma unroll
            for (int gg = 0; gg < GB; ++gg) {
              const int g = b * GB + gg;
              if (g < PT) {
                const int l = warp + 8 * g;
                const double *src = a.Gh + (size_t)perm_s[(l <= p) ? l : p] * ldg;
#pragma unroll
                for (int r = 0; r < NRR; ++r) {
                  if (8 * g + 8 > 64 * r) {  // compile-time: the column tile reaches below row 64 r
                    vv[gg][r][0] = __ldg(src + pr[r][0]);
                    vv[gg][r][1] = __ldg(src + pr[r][1]);
                  }
                }
              }
            }
#pragma unroll
            for (int gg = 0; gg < GB; ++gg) {
              const int g = b * GB + gg;
              if (g < PT) {
                const int l = warp + 8 * g;
                const bool colv = l <= p;
#pragma unroll
                for (int r = 0; r < NRR; ++r) {
                  const int i0 = 2 * lane + 64 * r;
                  if (64 * r < 8 * g + 8 && i0 < 8 * g + 8 && i0 < NR) {
                    double2 v;
                    v.x = (colv && pv[r][0]) ? vv[gg][r][0] : 0.0;
                    v.y = (colv && pv[r][1]) ? vv[gg][r][1] : 0.0;
                    *reinterpret_cast<double2 *>(A + (size_t)l * ld + i0) = v;
                  }
                }
              }
            }
          }
        }
      }
      if constexpr (MODE != 1) {
        if (warp == 7) {
          double s0 = 0.0;
          for (int i = lane; i < p; i += 32) s0 = fma(a.cte[i], a.cte[i], s0);
          s0 = warp_sum(s0);
          if (lane == 0) cost[0] = s0;
        }
        for (int e = tid; e < 8 * NR; e += 256) wcost[e] = 0.0;
      }
      __syncthreads();

      const long long t_b = LSSPA_CLOCK();
      long long t_acc = 0, t_diag = 0, t_wait = 0;
      // ---- phase 1: left-looking blocked Cholesky, software-pipelined around the serial part.
      // The 8x8 diagonal blocks form a chain (factor s needs row block s - 1), so warp 0 does
      // nothing else: D(s) = last update of the diagonal tile + factorisation + inverse.  Meanwhile
      // warps 1..7 pre-accumulate row block s + 1 over the finished rows k < s (P), after barrier 1
      // scale their tiles of row block s with Dinv_s (S), and after barrier 2 add the missing k = s
      // term (F).  Tile (r, L), L > r, belongs to warp 1 + (L - r - 1) % 7; the diagonal tile of row
      // r is pre-accumulated in place by warp 1 + r % 7.
      constexpr int NSL = (PT - 1 + 6) / 7;
      // The last NF warps also carry an 8-row tile of X in registers and eliminate it against row
      // block s right after that block is complete (phase 2 fused into the shadow of the diagonal
      // chain); NF is chosen so that the remaining row tiles are exactly one round of 8 warps.
      double *wc = wcost + (size_t)warp * NR;
      const double *cvec = A + (size_t)p * ld;
      double xr[RT][2];
      double r_in = 0.0;
      constexpr int NF = (MODE != 0 || RT - 8 < 0) ? 0 : (RT - 8 > 6 ? 6 : RT - 8);  // fused tiles: the rest is one round of 8
      // fused warps: 7, 6, 5, 3, 2, 1 in that order -- never warp 4, which issues on the same SM
      // sub-partition as the diagonal chain of warp 0 (measured: 3 % faster than warps 3..7)
      const int frank = (warp == 0 || warp == 4) ? 99 : (warp > 4 ? 7 - warp : 6 - warp);
      const bool fused = frank < NF;
      if (fused) load_x<RT>(xr, r_in, a, perm_s, frank, p, c, q);
      if constexpr (MODE != 2) {
        double2 tv[NSL];
#pragma unroll
        for (int sl = 0; sl < NSL; ++sl) {
          const int L = warp + 7 * sl;  // row block 0
          tv[sl] = make_double2(0.0, 0.0);
          if (warp != 0 && L < PT) tv[sl] = ld_tile(A, ld, 0, 8 * L, c, q);
        }
        for (int s = 0; s < RT; ++s) {
          const int nf = (p - 8 * s < 8) ? p - 8 * s : 8;
          const long long u0 = LSSPA_CLOCK();
          double2 tvn[NSL];
          if (warp == 0) {
            double2 t = ld_tile(A, ld, 8 * s, 8 * s, c, q);
            if (s > 0) {
              const double2 xa = ld_tile(A, ld, 8 * (s - 1), 8 * s, c, q);
              double p0 = 0.0, p1 = 0.0, e0 = 0.0, e1 = 0.0;
              dmma(p0, p1, xa.x, xa.x);
              dmma(e0, e1, xa.y, xa.y);
              t.x -= p0 + e0;
              t.y -= p1 + e1;
            }
            double y0, y1;
            diag_factor(t.x, t.y, y0, y1, nf, lane);
            st_tile(A, ld, 8 * s, 8 * s, c, q, t);
            *reinterpret_cast<double2 *>(Dbuf + s * 64 + c * 8 + 2 * q) = make_double2(y0, y1);
          } else if (s + 1 < RT) {
#pragma unroll
            for (int sl = 0; sl < NSL; ++sl) {
              const int L = s + 1 + warp + 7 * sl;
              tvn[sl] = make_double2(0.0, 0.0);
              if (L < PT) {
                const double2 g = ld_tile(A, ld, 8 * (s + 1), 8 * L, c, q);
                const double2 z = acc_tile<ld>(A, s, L, s + 1, c, q);
                tvn[sl] = make_double2(g.x - z.x, g.y - z.y);
              }
            }
            if (warp == 1 + (s + 1) % 7 && s > 0) {
              const double2 g = ld_tile(A, ld, 8 * (s + 1), 8 * (s + 1), c, q);
              const double2 z = acc_tile<ld>(A, s, s + 1, s + 1, c, q);
              st_tile(A, ld, 8 * (s + 1), 8 * (s + 1), c, q, make_double2(g.x - z.x, g.y - z.y));
            }
          }
          const long long u1 = LSSPA_CLOCK();
          t_acc += u1 - u0;
          __syncthreads();
          const long long u2 = LSSPA_CLOCK();
          t_wait += u2 - u1;
          if (warp != 0) {
            const double2 dv = ld_tile(Dbuf + s * 64, 8, 0, 0, c, q);
#pragma unroll
            for (int sl = 0; sl < NSL; ++sl) {
              const int L = s + warp + 7 * sl;
              if (L < PT) {
                double r0 = 0.0, r1 = 0.0;
                dmma(r0, r1, tv[sl].x, dv.x);
                dmma(r0, r1, tv[sl].y, dv.y);
                st_tile(A, ld, 8 * s, 8 * L, c, q, make_double2(r0, r1));
              }
            }
          }
          __syncthreads();
          if (warp != 0 && s + 1 < RT) {
#pragma unroll
            for (int sl = 0; sl < NSL; ++sl) {
              const int L = s + 1 + warp + 7 * sl;
              tv[sl] = tvn[sl];
              if (L < PT) {
                const double2 xa = ld_tile(A, ld, 8 * s, 8 * L, c, q);
                const double2 xb = ld_tile(A, ld, 8 * s, 8 * (s + 1), c, q);
                double p0 = 0.0, p1 = 0.0, e0 = 0.0, e1 = 0.0;
                dmma(p0, p1, xa.x, xb.x);
                dmma(e0, e1, xa.y, xb.y);
                tv[sl].x -= p0 + e0;
                tv[sl].y -= p1 + e1;
              }
            }
          }
          if (fused) elim_dispatch<RT, ld>(s, xr, r_in, wc, A, Dbuf, cvec, p, c, q);
          t_diag += LSSPA_CLOCK() - u2;
        }
      }

      const long long t_c = LSSPA_CLOCK();
      if constexpr (MODE == 1) {
        // ---- split route: R (upper tiles, with c) and the diagonal inverses go to global memory
        double2 *dst = reinterpret_cast<double2 *>(a.fact + eval * FD);
#pragma unroll
        for (int L = 0; L < PT; ++L) {
          const int nrow2 = 4 * ((L < RT ? L : RT - 1) + 1);   // double2 per column
          int off2 = 0;
#pragma unroll
          for (int k = 0; k < L; ++k) off2 += 32 * ((k < RT ? k : RT - 1) + 1);
#pragma unroll
          for (int e0 = 0; e0 < 8 * nrow2; e0 += 256) {
            const int e = e0 + tid;
            if (e < 8 * nrow2) {
              const int cc = e / nrow2, r2 = e - cc * nrow2;
              dst[off2 + e] = *reinterpret_cast<const double2 *>(A + (size_t)(8 * L + cc) * ld + 2 * r2);
            }
          }
        }
#pragma unroll
        for (int e0 = 0; e0 < 32 * RT; e0 += 256) {
          const int e = e0 + tid;
          if (e < 32 * RT) dst[(FD - 64 * RT) / 2 + e] = reinterpret_cast<const double2 *>(Dbuf)[e];
        }
        (void)t_a; (void)t_b; (void)t_c; (void)t_acc; (void)t_diag; (void)t_wait;
        (void)wc; (void)cvec; (void)xr; (void)r_in; (void)weight;
      } else {
      // ---- phase 2: the row tiles of X that were not carried through the factorisation
      for (int it = NF + warp; it < RT; it += 8) {
        load_x<RT>(xr, r_in, a, perm_s, it, p, c, q);
        elim_all<RT, ld>(xr, r_in, wc, A, Dbuf, cvec, p, c, q);
      }
      const long long t_d = LSSPA_CLOCK();
      __syncthreads();
#ifdef LSSPA_LIFTS_TIMING
      if (a.dbg != nullptr && blockIdx.x == 0 && lane == 0 && sidx == blockIdx.x && h == 0) {
        long long *d = a.dbg + warp * 8;
        d[0] = t_b - t_a;   // gather
        d[1] = t_diag;      // diagonal blocks (warp 0)
        d[2] = t_acc;       // accumulation of the row block
        d[3] = t_wait;      // waiting for the diagonal block
        d[4] = t_c - t_b;   // whole phase 1
        d[5] = t_d - t_c;   // phase 2 (this warp)
      }
#else
      (void)t_a; (void)t_b; (void)t_c; (void)t_d; (void)t_acc; (void)t_diag; (void)t_wait;
#endif
      for (int k = tid; k < p; k += 256) {
        double sacc = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sacc += wcost[(size_t)w * NR + k];
        cost[k + 1] = sacc;
      }
      __syncthreads();
      for (int k = tid; k < p; k += 256) {
        const double lift = (cost[k] - cost[k + 1]) * a.inv_ynsq;
        const int f = perm_s[k];
        acc[f] = (h == 0 ? 0.0 : acc[f]) + weight * lift;
      }
      }
    }
    __syncthreads();
    if constexpr (MODE != 1)
      for (int f = tid; f < p; f += 256) a.out[sidx * p + f] = acc[f];
  }
}


// =====================================================================================================
// Packed variant of the fused kernel (MODE 0): four warps per CTA and only the UPPER tiles of the
// permuted Gram matrix in shared memory, so that four independent factorisations are resident per SM
// at p = 100 (one serial diagonal chain per SM sub-partition instead of two per SM) -- 55 KB per CTA
// instead of 102 KB.  Column tile L (8 columns) is stored column-major with its own leading dimension
// ld(L) = 8 (L + 1) rounded up to 8 mod 16 (conflict-free tile accesses) at offset off(L).  The inverse
// of diagonal block s overwrites the diagonal tile (s, s), which nothing reads once it is factored --
// except in the last row block, whose diagonal tile may hold entries of c~ (column p): that one keeps U
// and its inverse goes to a 64-double side buffer.  The chain warp rotates with blockIdx so that the
// chains of co-resident CTAs issue on different sub-partitions.
__host__ __device__ constexpr int pk_rows(int L, int RT) { return 8 * ((L < RT ? L : RT - 1) + 1); }
__host__ __device__ constexpr int pk_ld(int L, int RT) { return (pk_rows(L, RT) % 16 == 8) ? pk_rows(L, RT) : pk_rows(L, RT) + 8; }
__host__ __device__ constexpr int pk_off(int L, int RT) {
  int o = 0;
  for (int l = 0; l < L; ++l) o += 8 * pk_ld(l, RT);
  return o;
}
// run-time forms (L <= RT): ld = 8 (L' + 1) + 8 (L' & 1) with L' = min(L, RT - 1); off = 64 (L (L + 1) / 2 + L / 2)
template <int RT>
__device__ __forceinline__ int pk_ld_rt(int L) {
  const int l = L < RT ? L : RT - 1;
  return 8 * (l + 1) + 8 * (l & 1);
}
template <int RT>
__device__ __forceinline__ int pk_off_rt(int L) {
  // columns before tile L: tiles 0..L-1, all with index < RT (L <= RT)
  return 32 * L * (L + 1) + 64 * (L >> 1);
}
// tile (k, L) of the packed array: rows 8k.., column tile L
template <int RT>
__device__ __forceinline__ double2 ldp(const double *A, int k, int L, int c, int q) {
  return *reinterpret_cast<const double2 *>(A + pk_off_rt<RT>(L) + c * pk_ld_rt<RT>(L) + 8 * k + 2 * q);
}
template <int RT>
__device__ __forceinline__ void stp(double *A, int k, int L, int c, int q, double2 v) {
  *reinterpret_cast<double2 *>(A + pk_off_rt<RT>(L) + c * pk_ld_rt<RT>(L) + 8 * k + 2 * q) = v;
}
// inverse of diagonal block J (see above): tile (J, J) for J < RT - 1, the side buffer for the last block
template <int RT>
__device__ __forceinline__ double2 ld_dinv(const double *A, const double *Dlast, int J, int c, int q) {
  if (J < RT - 1) return ldp<RT>(A, J, J, c, q);
  return *reinterpret_cast<const double2 *>(Dlast + c * 8 + 2 * q);
}

// Elimination step J of one 8-row tile of X on the packed array (see elim_step)
template <int RT, int J>
__device__ __forceinline__ void elim_step4(double (&xr)[RT][2], double &r_in, double *wc, const double *A,
                                           const double *Dlast, const double *cvec, int p, int c, int q) {
  const double2 dv = ld_dinv<RT>(A, Dlast, J, c, q);
  double m0 = 0.0, m1 = 0.0;
  dmma(m0, m1, xr[J][0], dv.x);
  dmma(m0, m1, xr[J][1], dv.y);
  const double2 cv = *reinterpret_cast<const double2 *>(cvec + 8 * J + 2 * q);
  double ra = r_in, rb = r_in;
  dmma(ra, rb, -m0, (2 * q <= c) ? cv.x : 0.0);
  dmma(ra, rb, -m1, (2 * q + 1 <= c) ? cv.y : 0.0);
  const double d0 = ra * ra, d1 = rb * rb;
  const bool odd = (c & 1) != 0;
  double tot = (odd ? d1 : d0) + __shfl_xor_sync(kFull, odd ? d0 : d1, 4);
  tot += __shfl_xor_sync(kFull, tot, 8);
  tot += __shfl_xor_sync(kFull, tot, 16);
  if (c < 2) {
    const int k0 = 8 * J + 2 * q + c;
    if (k0 < p) wc[k0] += tot;
  }
  r_in = __shfl_sync(kFull, rb, 3, 4);
  m0 = -m0;
  m1 = -m1;
#pragma unroll
  for (int L = J + 1; L < RT; ++L) {
    const double2 rt = *reinterpret_cast<const double2 *>(A + pk_off(L, RT) + c * pk_ld(L, RT) + 8 * J + 2 * q);
    dmma(xr[L][0], xr[L][1], m0, rt.x);
    dmma(xr[L][0], xr[L][1], m1, rt.y);
  }
}
template <int RT, int J = 0>
__device__ __forceinline__ void elim_all4(double (&xr)[RT][2], double &r_in, double *wc, const double *A,
                                          const double *Dlast, const double *cvec, int p, int c, int q) {
  if constexpr (J < RT) {
    elim_step4<RT, J>(xr, r_in, wc, A, Dlast, cvec, p, c, q);
    elim_all4<RT, J + 1>(xr, r_in, wc, A, Dlast, cvec, p, c, q);
  }
}

// sum_{k < kend} R(k, L)^T R(k, S), packed array, four independent chains
template <int RT>
__device__ __forceinline__ double2 acc_tile4(const double *A, int kend, int L, int S, int c, int q) {
  const double *pl = A + pk_off_rt<RT>(L) + c * pk_ld_rt<RT>(L) + 2 * q;
  const double *ps = A + pk_off_rt<RT>(S) + c * pk_ld_rt<RT>(S) + 2 * q;
  double p0 = 0.0, p1 = 0.0, r0 = 0.0, r1 = 0.0, e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0;
  int k = 0;
  for (; k + 1 < kend; k += 2) {
    const double2 xa = *reinterpret_cast<const double2 *>(pl + 8 * k);
    const double2 xb = *reinterpret_cast<const double2 *>(ps + 8 * k);
    const double2 ya = *reinterpret_cast<const double2 *>(pl + 8 * k + 8);
    const double2 yb = *reinterpret_cast<const double2 *>(ps + 8 * k + 8);
    dmma(p0, p1, xa.x, xb.x);
    dmma(r0, r1, ya.x, yb.x);
    dmma(e0, e1, xa.y, xb.y);
    dmma(f0, f1, ya.y, yb.y);
  }
  if (k < kend) {
    const double2 xa = *reinterpret_cast<const double2 *>(pl + 8 * k);
    const double2 xb = *reinterpret_cast<const double2 *>(ps + 8 * k);
    dmma(p0, p1, xa.x, xb.x);
    dmma(e0, e1, xa.y, xb.y);
  }
  return make_double2((p0 + r0) + (e0 + f0), (p1 + r1) + (e1 + f1));
}

constexpr int kC4Threads = 128;

// MODE 0: the whole evaluation.  MODE 1: factor only (split route): phases 0 and 1, then R's upper tiles (with
// c~) and the inverses of the diagonal blocks go to a.fact in the record layout lifts_elim4_kernel reads.
template <int RT, int PT, int MINB, int MODE>
__global__ void __launch_bounds__(kC4Threads, MINB) lifts_chol4_kernel(CholParams a, int sms) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int p = a.p;
  constexpr int NR = 8 * RT;
  constexpr int TOT = pk_off(PT, RT);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = lane >> 2, q = lane & 3;
  // role 0 = diagonal chain, roles 1..3 = workers; rotated so that co-resident CTAs (blockIdx differing by
  // multiples of the SM count on a full grid) run their chains on different SM sub-partitions
  const int role = (warp - (int)((blockIdx.x / (unsigned)sms) & 3u)) & 3;

  double *A = reinterpret_cast<double *>(smem_raw);   // packed upper tiles
  double *Dlast = A + TOT;                            // 64: inverse of the last diagonal block
  double *wcost = Dlast + 64;                         // 4 x NR per-warp cost partials
  double *cost = wcost + (size_t)4 * NR;              // p + 2
  double *acc = cost + (p + 2);                       // p
  int *perm_s = reinterpret_cast<int *>(acc + p + (p & 1));  // p + 1

  const int halves = a.anti ? 2 : 1;
  const double weight = a.anti ? 0.5 : 1.0;
  const int ldg = p + 1;
  constexpr int LP = PT - 1;                          // column tile holding column p (c~)
  const double *cvec = A + pk_off(LP, RT) + (p - 8 * LP) * pk_ld(LP, RT);

  for (int64_t sidx = blockIdx.x; sidx < a.count; sidx += gridDim.x) {
    for (int h = 0; h < halves; ++h) {
      __syncthreads();
      for (int k = tid; k <= p; k += kC4Threads)
        perm_s[k] = (k == p) ? p : a.perms[sidx * p + (h == 0 ? k : p - 1 - k)];
      __syncthreads();
      const long long t_a = LSSPA_CLOCK();
      // ---- phase 0: gather the upper tiles of Gh[pi^, pi^] with 8-byte cp.async (zero-fill where masked):
      // warp w takes the columns l = w + 4 g, lane the rows 2 lane, 2 lane + 1 (+ 64); every copy of the
      // evaluation is in flight before the first one is awaited
      {
        constexpr int NRR = (NR + 63) / 64;
        int pr[NRR][2];
        bool pv[NRR][2];
#pragma unroll
        for (int r = 0; r < NRR; ++r) {
          const int i0 = 2 * lane + 64 * r;
          pv[r][0] = i0 < p;
          pv[r][1] = i0 + 1 < p;
          pr[r][0] = pv[r][0] ? perm_s[i0] : 0;
          pr[r][1] = pv[r][1] ? perm_s[i0 + 1] : 0;
        }
#pragma unroll 2
        for (int g = 0; g < 2 * PT; ++g) {
          const int l = warp + 4 * g;
          const int L = l >> 3, cc = l & 7;
          const bool colv = l <= p;
          const double *src = a.Gh + (size_t)perm_s[colv ? l : p] * ldg;
          const int rows = 8 * ((L < RT ? L : RT - 1) + 1);
          double *dst = A + pk_off_rt<RT>(L) + cc * pk_ld_rt<RT>(L);
#pragma unroll
          for (int r = 0; r < NRR; ++r) {
            const int i0 = 2 * lane + 64 * r;
            if (i0 < rows) {
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const unsigned dsta = (unsigned)__cvta_generic_to_shared(dst + i0 + e);
                const int nbytes = (colv && pv[r][e]) ? 8 : 0;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dsta), "l"(src + pr[r][e]), "r"(nbytes)
                             : "memory");
              }
            }
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      if constexpr (MODE == 0) {
        if (warp == 3) {
          double s0 = 0.0;
          for (int i = lane; i < p; i += 32) s0 = fma(a.cte[i], a.cte[i], s0);
          s0 = warp_sum(s0);
          if (lane == 0) cost[0] = s0;
        }
        for (int e = tid; e < 4 * NR; e += kC4Threads) wcost[e] = 0.0;
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      const long long t_b = LSSPA_CLOCK();
      long long t_work = 0, t_w1 = 0, t_w2 = 0, t_sf = 0;

      // ---- phase 1: left-looking blocked Cholesky around the serial diagonal chain (see lifts_chol_kernel).
      // Tile (r, L), L > r, belongs to worker (L - r - 1) % 3; the diagonal tile of row r is pre-accumulated in
      // place by worker r % 3.
      constexpr int NSL = (PT - 1 + 2) / 3;
      double *wc = wcost + (size_t)warp * NR;
      {
        double2 tv[NSL];
#pragma unroll
        for (int sl = 0; sl < NSL; ++sl) {
          const int L = role + 3 * sl;  // row block 0
          tv[sl] = make_double2(0.0, 0.0);
          if (role != 0 && L < PT) tv[sl] = ldp<RT>(A, 0, L, c, q);
        }
        for (int s = 0; s < RT; ++s) {
          const int nf = (p - 8 * s < 8) ? p - 8 * s : 8;
          double2 tvn[NSL];
          const long long u0 = LSSPA_CLOCK();
          if (role == 0) {
            double2 t = ldp<RT>(A, s, s, c, q);
            if (s > 0) {
              const double2 xa = ldp<RT>(A, s - 1, s, c, q);
              double p0 = 0.0, p1 = 0.0, e0 = 0.0, e1 = 0.0;
              dmma(p0, p1, xa.x, xa.x);
              dmma(e0, e1, xa.y, xa.y);
              t.x -= p0 + e0;
              t.y -= p1 + e1;
            }
            double y0, y1;
            diag_factor(t.x, t.y, y0, y1, nf, lane);
            if (s < RT - 1) {
              stp<RT>(A, s, s, c, q, make_double2(y0, y1));       // inverse over the (dead) diagonal tile
            } else {
              stp<RT>(A, s, s, c, q, t);                          // last block: U stays (it may hold c~)
              *reinterpret_cast<double2 *>(Dlast + c * 8 + 2 * q) = make_double2(y0, y1);
            }
          } else if (s + 1 < RT) {
#pragma unroll
            for (int sl = 0; sl < NSL; ++sl) {
              const int L = s + 1 + role + 3 * sl;
              tvn[sl] = make_double2(0.0, 0.0);
              if (L < PT) {
                const double2 g = ldp<RT>(A, s + 1, L, c, q);
                const double2 z = acc_tile4<RT>(A, s, L, s + 1, c, q);
                tvn[sl] = make_double2(g.x - z.x, g.y - z.y);
              }
            }
            if (role == 1 + (s + 1) % 3 && s > 0) {
              const double2 g = ldp<RT>(A, s + 1, s + 1, c, q);
              const double2 z = acc_tile4<RT>(A, s, s + 1, s + 1, c, q);
              stp<RT>(A, s + 1, s + 1, c, q, make_double2(g.x - z.x, g.y - z.y));
            }
          }
          const long long u1 = LSSPA_CLOCK();
          __syncthreads();
          const long long u2 = LSSPA_CLOCK();
          t_work += u1 - u0;
          t_w1 += u2 - u1;
          if (role != 0) {
            const double2 dv = ld_dinv<RT>(A, Dlast, s, c, q);
#pragma unroll
            for (int sl = 0; sl < NSL; ++sl) {
              const int L = s + role + 3 * sl;
              if (L < PT) {
                double r0 = 0.0, r1 = 0.0;
                dmma(r0, r1, tv[sl].x, dv.x);
                dmma(r0, r1, tv[sl].y, dv.y);
                stp<RT>(A, s, L, c, q, make_double2(r0, r1));
              }
            }
          }
          const long long u3 = LSSPA_CLOCK();
          __syncthreads();
          const long long u4 = LSSPA_CLOCK();
          t_sf += u3 - u2;
          t_w2 += u4 - u3;
          if (role != 0 && s + 1 < RT) {
#pragma unroll
            for (int sl = 0; sl < NSL; ++sl) {
              const int L = s + 1 + role + 3 * sl;
              tv[sl] = tvn[sl];
              if (L < PT) {
                const double2 xa = ldp<RT>(A, s, L, c, q);
                const double2 xb = ldp<RT>(A, s, s + 1, c, q);
                double p0 = 0.0, p1 = 0.0, e0 = 0.0, e1 = 0.0;
                dmma(p0, p1, xa.x, xb.x);
                dmma(e0, e1, xa.y, xb.y);
                tv[sl].x -= p0 + e0;
                tv[sl].y -= p1 + e1;
              }
            }
          }
          t_sf += LSSPA_CLOCK() - u4;
        }
      }
      const long long t_c = LSSPA_CLOCK();
      if constexpr (MODE == 1) {
        // ---- split route: the record of this evaluation (tile column L = 8 columns of rows(L) doubles,
        // contiguous; then the RT inverses of the diagonal blocks, 64 doubles each in (c, 2q + e) order).
        // The diagonal tiles of the record hold what the packed array holds there (the inverse; U in the
        // last block) -- the elimination kernel overwrites them from the tail anyway.
        constexpr int64_t FD = chol_fact_doubles(RT, PT);
        double2 *dst = reinterpret_cast<double2 *>(a.fact + (sidx * halves + h) * FD);
#pragma unroll
        for (int L = 0; L < PT; ++L) {
          constexpr int dummy = 0;
          (void)dummy;
          const int rows = pk_rows(L, RT);
          int off = 0;
#pragma unroll
          for (int k = 0; k < L; ++k) off += 8 * pk_rows(k, RT);
          const int n2 = 8 * rows / 2;
          for (int e = tid; e < n2; e += kC4Threads) {
            const int cc = e / (rows / 2), r2 = e - cc * (rows / 2);
            dst[off / 2 + e] = *reinterpret_cast<const double2 *>(A + pk_off(L, RT) + cc * pk_ld(L, RT) + 2 * r2);
          }
        }
        for (int e = tid; e < 32 * RT; e += kC4Threads) {
          const int sblk = e >> 5, l = e & 31, cc = l >> 2, qq = l & 3;
          const double *srcd = (sblk < RT - 1) ? A + pk_off_rt<RT>(sblk) + cc * pk_ld_rt<RT>(sblk) + 8 * sblk + 2 * qq
                                               : Dlast + cc * 8 + 2 * qq;
          dst[(FD - 64 * RT) / 2 + e] = *reinterpret_cast<const double2 *>(srcd);
        }
        (void)t_a; (void)t_b; (void)t_c; (void)t_work; (void)t_w1; (void)t_w2; (void)t_sf;
        (void)wc; (void)cvec; (void)weight; (void)cost; (void)acc;
        continue;
      }
      // ---- phase 2: elimination of X = R_te[:, perm], row tiles warp, warp + 4, ...
      {
        double xr[RT][2];
        double r_in = 0.0;
        for (int it = warp; it < RT; it += 4) {
          load_x<RT>(xr, r_in, a, perm_s, it, p, c, q);
          elim_all4<RT>(xr, r_in, wc, A, Dlast, cvec, p, c, q);
        }
      }
      const long long t_d = LSSPA_CLOCK();
      __syncthreads();
#ifdef LSSPA_LIFTS_TIMING
      if (a.dbg != nullptr && blockIdx.x == 0 && lane == 0 && sidx == blockIdx.x && h == 0) {
        long long *d = a.dbg + warp * 8;
        d[0] = t_b - t_a;   // gather
        d[1] = t_work;      // chain: diagonal blocks / workers: pre-accumulation
        d[2] = t_w1;        // wait at barrier 1
        d[3] = t_sf;        // scale + finish phases
        d[4] = t_w2;        // wait at barrier 2
        d[5] = t_c - t_b;   // whole phase 1
        d[6] = t_d - t_c;   // phase 2 (this warp)
        d[7] = LSSPA_CLOCK() - t_d;   // wait for the slowest warp of phase 2
      }
#else
      (void)t_a; (void)t_b; (void)t_c; (void)t_d; (void)t_work; (void)t_w1; (void)t_w2; (void)t_sf;
#endif
      for (int k = tid; k < p; k += kC4Threads) {
        double sacc = 0.0;
#pragma unroll
        for (int w = 0; w < 4; ++w) sacc += wcost[(size_t)w * NR + k];
        cost[k + 1] = sacc;
      }
      __syncthreads();
      for (int k = tid; k < p; k += kC4Threads) {
        const double lift = (cost[k] - cost[k + 1]) * a.inv_ynsq;
        const int f = perm_s[k];
        acc[f] = (h == 0 ? 0.0 : acc[f]) + weight * lift;
      }
    }
    __syncthreads();
    if constexpr (MODE == 0)
      for (int f = tid; f < p; f += kC4Threads) a.out[sidx * p + f] = acc[f];
  }
}

// Elimination half of the split route (MODE 2 of lifts_chol_kernel) on the packed layout: four warps,
// four CTAs per SM.  The factors of one evaluation (R's upper tiles with c~, then the inverses of the
// diagonal blocks, as lsspa_lifts_chol_factor wrote them) are copied into the packed tile array with
// 16-byte cp.async, the inverses over the diagonal tiles (last block: side buffer, see above), then
// phase 2 and the cost / lift / scatter epilogue of lifts_chol4_kernel.
template <int RT, int PT, int MINB>
__global__ void __launch_bounds__(kC4Threads, MINB) lifts_elim4_kernel(CholParams a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int p = a.p;
  constexpr int NR = 8 * RT;
  constexpr int TOT = pk_off(PT, RT);
  constexpr int64_t FD = chol_fact_doubles(RT, PT);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = lane >> 2, q = lane & 3;
  double *A = reinterpret_cast<double *>(smem_raw);
  double *Dlast = A + TOT;
  double *wcost = Dlast + 64;
  double *cost = wcost + (size_t)4 * NR;
  double *acc = cost + (p + 2);
  int *perm_s = reinterpret_cast<int *>(acc + p + (p & 1));
  const int halves = a.anti ? 2 : 1;
  const double weight = a.anti ? 0.5 : 1.0;
  constexpr int LP = PT - 1;
  const double *cvec = A + pk_off(LP, RT) + (p - 8 * LP) * pk_ld(LP, RT);

  for (int64_t sidx = blockIdx.x; sidx < a.count; sidx += gridDim.x) {
    for (int h = 0; h < halves; ++h) {
      __syncthreads();
      for (int k = tid; k <= p; k += kC4Threads)
        perm_s[k] = (k == p) ? p : a.perms[sidx * p + (h == 0 ? k : p - 1 - k)];
      const double *src = a.fact + (sidx * halves + h) * FD;
      // tile column L: 8 columns of rows(L) doubles, contiguous in the record; 16 bytes per copy
#pragma unroll
      for (int L = 0; L < PT; ++L) {
        constexpr int dummy = 0;
        (void)dummy;
        const int rows = pk_rows(L, RT);
        int off = 0;
#pragma unroll
        for (int k = 0; k < L; ++k) off += 8 * pk_rows(k, RT);
        const int n2 = 8 * rows / 2;   // double2 in this tile column
        for (int e = tid; e < n2; e += kC4Threads) {
          const int cc = e / (rows / 2), r2 = e - cc * (rows / 2);
          const unsigned dsta = (unsigned)__cvta_generic_to_shared(A + pk_off(L, RT) + cc * pk_ld(L, RT) + 2 * r2);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dsta), "l"(src + off + 2 * e) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (warp == 3) {
        double s0 = 0.0;
        for (int i = lane; i < p; i += 32) s0 = fma(a.cte[i], a.cte[i], s0);
        s0 = warp_sum(s0);
        if (lane == 0) cost[0] = s0;
      }
      for (int e = tid; e < 4 * NR; e += kC4Threads) wcost[e] = 0.0;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      // inverses of the diagonal blocks (8 x 8, leading dimension 8) over the diagonal tiles
      {
        const double *dsrc = src + (FD - 64 * RT);
        for (int e = tid; e < 32 * RT; e += kC4Threads) {
          const int sblk = e >> 5, l = e & 31;          // block, lane-like index: 2 l = cc * 8 + 2 qq
          const double2 v = __ldg(reinterpret_cast<const double2 *>(dsrc) + e);
          const int cc = l >> 2, qq = l & 3;
          if (sblk < RT - 1)
            *reinterpret_cast<double2 *>(A + pk_off_rt<RT>(sblk) + cc * pk_ld_rt<RT>(sblk) + 8 * sblk + 2 * qq) = v;
          else
            *reinterpret_cast<double2 *>(Dlast + cc * 8 + 2 * qq) = v;
        }
      }
      __syncthreads();
      double *wc = wcost + (size_t)warp * NR;
      {
        double xr[RT][2];
        double r_in = 0.0;
        for (int it = warp; it < RT; it += 4) {
          load_x<RT>(xr, r_in, a, perm_s, it, p, c, q);
          elim_all4<RT>(xr, r_in, wc, A, Dlast, cvec, p, c, q);
        }
      }
      __syncthreads();
      for (int k = tid; k < p; k += kC4Threads) {
        double sacc = 0.0;
#pragma unroll
        for (int w = 0; w < 4; ++w) sacc += wcost[(size_t)w * NR + k];
        cost[k + 1] = sacc;
      }
      __syncthreads();
      for (int k = tid; k < p; k += kC4Threads) {
        const double lift = (cost[k] - cost[k + 1]) * a.inv_ynsq;
        const int f = perm_s[k];
        acc[f] = (h == 0 ? 0.0 : acc[f]) + weight * lift;
      }
    }
    __syncthreads();
    for (int f = tid; f < p; f += kC4Threads) a.out[sidx * p + f] = acc[f];
  }
}

template <int RT>
size_t chol4_smem_bytes(int p, int pt) {
  const int tot = (pt == RT) ? pk_off(RT, RT) : pk_off(RT + 1, RT);
  const size_t d = (size_t)tot + 64 + (size_t)4 * 8 * RT + (size_t)(p + 2) + (size_t)(p + 1);
  return d * sizeof(double) + (size_t)(p + 1) * sizeof(int) + 16;
}

template <int RT, int PT, int MINB, int MODE>
int launch_chol4(const CholParams &a, int grid, size_t smem, int sms, cudaStream_t st) {
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(lifts_chol4_kernel<RT, PT, MINB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(lifts_chol4_kernel<RT, PT, MINB, MODE>,
                                      cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  lifts_chol4_kernel<RT, PT, MINB, MODE><<<grid, kC4Threads, smem, st>>>(a, sms);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

// p = 49 .. 128 (RT = 7 .. 16): CTAs per SM from the packed footprint
template <int RT, int PT, int MINB>
int launch_elim4(const CholParams &a, int grid, size_t smem, cudaStream_t st) {
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(lifts_elim4_kernel<RT, PT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(lifts_elim4_kernel<RT, PT, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
  lifts_elim4_kernel<RT, PT, MINB><<<grid, kC4Threads, smem, st>>>(a);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

template <int RT>
int run_chol4(const CholParams &a, int64_t count, int mode, cudaStream_t st) {
  const DeviceInfo &d = device_info();
  const int sms = d.sm_count > 0 ? d.sm_count : 148;
  const size_t smem = chol4_smem_bytes<RT>(a.p, a.pt);
  int per_sm = (int)((size_t)(228 * 1024) / (smem + 1024));
  if (per_sm > 4) per_sm = 4;
  static const int cap = [] { const char *e = getenv("LSSPA_C4_PER_SM"); return e ? atoi(e) : 4; }();   // experiments
  if (per_sm > cap) per_sm = cap;
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)per_sm * sms;
  if (grid > count) grid = count;
  constexpr int MINB = RT <= 13 ? 4 : (RT <= 15 ? 3 : 2);   // resident CTAs the register budget is cut for
  if (mode == 2)
    return (a.pt == RT) ? launch_elim4<RT, RT, MINB>(a, (int)grid, smem, st) : launch_elim4<RT, RT + 1, MINB>(a, (int)grid, smem, st);
  if (mode == 1)
    return (a.pt == RT) ? launch_chol4<RT, RT, MINB, 1>(a, (int)grid, smem, sms, st)
                        : launch_chol4<RT, RT + 1, MINB, 1>(a, (int)grid, smem, sms, st);
  return (a.pt == RT) ? launch_chol4<RT, RT, MINB, 0>(a, (int)grid, smem, sms, st)
                      : launch_chol4<RT, RT + 1, MINB, 0>(a, (int)grid, smem, sms, st);
}

size_t chol_smem_bytes(int p) {
  const int rt = (p + 7) / 8, pt = (p + 8) / 8;
  const int ld = chol_ld(rt);
  const size_t d = (size_t)8 * pt * ld + (size_t)rt * 64 + (size_t)64 * rt + (size_t)(p + 2) + (size_t)(p + 1);
  return d * sizeof(double) + (size_t)(p + 1) * sizeof(int) + 32;
}

template <int RT, int PT, int MINB, int MODE>
int launch_chol(const CholParams &a, int grid, size_t smem, cudaStream_t st) {
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(lifts_chol_kernel<RT, PT, MINB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(lifts_chol_kernel<RT, PT, MINB, MODE>,
                                      cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  lifts_chol_kernel<RT, PT, MINB, MODE><<<grid, 256, smem, st>>>(a);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

template <int RT>
int launch_chol_rt(const CholParams &a, int grid, size_t smem, int mode, cudaStream_t st) {
  constexpr int MINB = RT <= 6 ? 3 : (RT <= 13 ? 2 : 1);   // resident CTAs per SM the register budget is cut for
  if constexpr (RT >= 7 && RT <= 16) {   // the split route is instantiated where the DMMA estimator also works
    if (mode == 1)
      return (a.pt == RT) ? launch_chol<RT, RT, MINB, 1>(a, grid, smem, st) : launch_chol<RT, RT + 1, MINB, 1>(a, grid, smem, st);
    if (mode == 2)
      return (a.pt == RT) ? launch_chol<RT, RT, MINB, 2>(a, grid, smem, st) : launch_chol<RT, RT + 1, MINB, 2>(a, grid, smem, st);
  }
  if (mode != 0) return LSSPA_E_UNSUPPORTED;
  return (a.pt == RT) ? launch_chol<RT, RT, MINB, 0>(a, grid, smem, st) : launch_chol<RT, RT + 1, MINB, 0>(a, grid, smem, st);
}

// Column norms of the upper-triangular R (column-major, ld p): D[j] = |R[:, j]|, 1 where the column is zero
__global__ void lift_scale_kernel(int p, const double *__restrict__ R, double *__restrict__ D) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p) return;
  double s0 = 0.0;
  for (int k = 0; k <= j; ++k) s0 = fma(R[(size_t)j * p + k], R[(size_t)j * p + k], s0);
  D[j] = s0 > 0.0 ? sqrt(s0) : 1.0;
}

// Gh = [R D^-1 | c]^T [R D^-1 | c] for upper-triangular R (column-major, ld p): block (i), thread (j).
// The columns are equilibrated (unit norm): the lifts are invariant under a scaling of the features
// when the test factor is scaled alike, and the scaling removes the part of the conditioning
// that is only due to units (within sqrt(p) of the best diagonal scaling, van der Sluis).
__global__ void lift_gram_kernel(int p, const double *__restrict__ R, const double *__restrict__ cvec,
                                 const double *__restrict__ D, double *__restrict__ Gh) {
  const int i = blockIdx.x;  // 0..p
  const double *ci = (i < p) ? R + (size_t)i * p : cvec;
  const int ni = (i < p) ? i + 1 : p;
  const double di = (i < p) ? D[i] : 1.0;
  for (int j = threadIdx.x; j <= p; j += blockDim.x) {
    const double *cj = (j < p) ? R + (size_t)j * p : cvec;
    const int nj = (j < p) ? j + 1 : p;
    const double dj = (j < p) ? D[j] : 1.0;
    const int n = ni < nj ? ni : nj;
    double s0 = 0.0, s1 = 0.0;
    int k = 0;
    for (; k + 1 < n; k += 2) {
      s0 = fma(ci[k], cj[k], s0);
      s1 = fma(ci[k + 1], cj[k + 1], s1);
    }
    if (k < n) s0 = fma(ci[k], cj[k], s0);
    Gh[(size_t)i * (p + 1) + j] = (s0 + s1) / (di * dj);
  }
}

// Condition bound of the equilibrated factor R' = R D^-1 (its Gram matrix Gh has a unit diagonal):
//   info[0] = min(info[2], info[3]) >= cond_2(R'),
//   info[2] = |R'|_F |R'^-1|_F,
//   info[3] = sqrt(max_i sum_j |Gh_ij|) * sqrt(|R'^-1|_1 |R'^-1|_inf)   (Gershgorin for the largest
//             eigenvalue of Gh, |A|_2^2 <= |A|_1 |A|_inf for the inverse; p <= 128 only) -- about
//             three times sharper on correlated data, so fewer benign problems are sent to the
//             Householder kernel;
//   info[1] = min |R'_kk| / max |R'_kk|.
// One CTA of 32 warps; a warp solves R' x = e_j by back substitution for its columns j (x in shared
// memory, the row dot products across the lanes).  inf when R is singular.
__global__ void __launch_bounds__(1024) lift_cond_kernel(int p, const double *__restrict__ R,
                                                         const double *__restrict__ D, const double *__restrict__ Gh,
                                                         double *__restrict__ info, int tight) {
  extern __shared__ double xs[];  // 32 x p solution vectors | R' (row-major p x p) | tight: 32 x p row sums of |R'^-1|
  __shared__ double red[4][32];
  __shared__ double rmin[32], rmax[32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  double *x = xs + (size_t)w * p;
  double *Rs = xs + (size_t)32 * p;   // Rs[i * p + k] = R'[i][k] = R[i][k] / D[k]
  double *rsum = Rs + (size_t)p * p + (size_t)w * p;
  for (int e = threadIdx.x; e < p * p; e += blockDim.x) {
    const int k = e / p, i = e - k * p;   // R is column-major: element e = R[i][k]
    Rs[(size_t)i * p + k] = (i <= k) ? R[e] / D[k] : 0.0;
  }
  if (tight)
    for (int e = threadIdx.x; e < 32 * p; e += blockDim.x) Rs[(size_t)p * p + e] = 0.0;
  __syncthreads();
  double fr = 0.0, fi = 0.0, dmin = 1e300, dmax = 0.0, colmax = 0.0, gersh = 0.0;
  for (int j = w; j < p; j += 32) {
    double colsum = 0.0;
    for (int i = j; i >= 0; --i) {
      const double *row = Rs + (size_t)i * p;
      double sacc = 0.0;
      for (int k = i + 1 + l; k <= j; k += 32) sacc = fma(-row[k], x[k], sacc);
      sacc = warp_sum(sacc);
      const double v = (sacc + ((i == j) ? 1.0 : 0.0)) / row[i];
      __syncwarp();
      if (l == 0) {
        x[i] = v;
        fi = fma(v, v, fi);
        colsum += fabs(v);
        if (tight) rsum[i] += fabs(v);
      }
      __syncwarp();
    }
    colmax = fmax(colmax, colsum);
    for (int i = l; i <= j; i += 32) {
      const double r = Rs[(size_t)i * p + j];
      fr = fma(r, r, fr);
    }
    if (l == 0) {
      const double d = fabs(Rs[(size_t)j * p + j]);
      dmin = fmin(dmin, d);
      dmax = fmax(dmax, d);
    }
    // Gershgorin row sum of the equilibrated Gram matrix (row j)
    double g = 0.0;
    for (int k = l; k < p; k += 32) g += fabs(Gh[(size_t)j * (p + 1) + k]);
    gersh = fmax(gersh, warp_sum(g));
  }
  fr = warp_sum(fr);
  fi = warp_sum(fi);
  for (int o = 16; o > 0; o >>= 1) {
    dmin = fmin(dmin, __shfl_xor_sync(kFull, dmin, o));
    dmax = fmax(dmax, __shfl_xor_sync(kFull, dmax, o));
    colmax = fmax(colmax, __shfl_xor_sync(kFull, colmax, o));
  }
  if (l == 0) {
    red[0][w] = fr;
    red[1][w] = fi;
    red[2][w] = colmax;
    red[3][w] = gersh;
    rmin[w] = dmin;
    rmax[w] = dmax;
  }
  __syncthreads();
  // |R'^-1|_inf: row i sums over all columns = over the 32 per-warp partial sums (fixed order)
  double rowmax = 0.0;
  if (tight) {
    for (int i = threadIdx.x; i < p; i += blockDim.x) {
      double t = 0.0;
      for (int k = 0; k < 32; ++k) t += Rs[(size_t)p * p + (size_t)k * p + i];
      rowmax = fmax(rowmax, t);
    }
    for (int o = 16; o > 0; o >>= 1) rowmax = fmax(rowmax, __shfl_xor_sync(kFull, rowmax, o));
  }
  __syncthreads();
  if (l == 0) x[0] = rowmax;   // the solution vectors are done: x[0] of every warp carries its row maximum
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0, mn = 1e300, mx = 0.0, cm = 0.0, gs = 0.0, rm = 0.0;
    for (int k = 0; k < 32; ++k) {
      a += red[0][k];
      b += red[1][k];
      cm = fmax(cm, red[2][k]);
      gs = fmax(gs, red[3][k]);
      rm = fmax(rm, xs[(size_t)k * p]);
      mn = fmin(mn, rmin[k]);
      mx = fmax(mx, rmax[k]);
    }
    double frob = sqrt(a) * sqrt(b);
    if (!(frob == frob)) frob = INFINITY;
    double sharp = tight ? sqrt(gs) * sqrt(cm * rm) : INFINITY;
    if (!(sharp == sharp)) sharp = INFINITY;
    info[0] = fmin(frob, sharp);
    info[1] = (mx > 0.0) ? mn / mx : 0.0;
    info[2] = frob;
    info[3] = sharp;
  }
}

}  // namespace

extern long long *g_lifts_dbg;  // lifts_mma.cu

bool lifts_chol_supported(int p) { return p >= 17 && p <= 152; }   // 152: the tile array of 19 row tiles still fits one SM

}  // namespace lsspa

using namespace lsspa;

extern "C" int lsspa_lifts_chol_supported(int p) { return lifts_chol_supported(p) ? 1 : 0; }

extern "C" int64_t lsspa_lifts_gram_doubles(int p) {
  if (p < 1) return 0;
  // Gh, info[8], D[p], scratch of the estimate (wide problems: the inverse of the equilibrated factor + 2p column sums)
  return (int64_t)(p + 1) * (p + 1) + 8 + p + (int64_t)p * p + 2 * (int64_t)p;
}

extern "C" int lsspa_lifts_gram(int p, const double *R_tr_cm, const double *c_tr, double *gram_out, void *stream) {
  if (p < 1 || !R_tr_cm || !c_tr || !gram_out) return LSSPA_E_BADARG;
  if (!lifts_chol_supported(p) && !lifts_big_supported(p)) return LSSPA_E_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  double *info = gram_out + (size_t)(p + 1) * (p + 1);
  double *D = info + 8;
  lift_scale_kernel<<<(p + 127) / 128, 128, 0, st>>>(p, R_tr_cm, D);
  LSSPA_LAUNCH_CHECK();
  lift_gram_kernel<<<p + 1, 128, 0, st>>>(p, R_tr_cm, c_tr, D, gram_out);
  LSSPA_LAUNCH_CHECK();
  if (lifts_big_supported(p)) {
    double *Xinv = D + p;
    return lifts_big_cond(p, R_tr_cm, D, gram_out, Xinv, Xinv + (size_t)p * p, info, st);
  }
  const int tight = p <= 128 ? 1 : 0;   // the sharper bound needs 32 p more doubles of shared memory
  const size_t cond_smem = ((size_t)32 * p * (1 + tight) + (size_t)p * p) * sizeof(double);
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(lift_cond_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cond_smem));
  lift_cond_kernel<<<1, 1024, cond_smem, st>>>(p, R_tr_cm, D, gram_out, info, tight);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

// mode 0: fused; 1: factor only (fact written); 2: eliminate only (fact read)
static int chol_run(int mode, int p, const double *gram, const double *R_te_cm, const double *c_te, double y_norm_sq,
                    const int32_t *perms, int64_t count, int antithetical, double *lifts_out, double *fact,
                    void *stream) {
  if (!lifts_chol_supported(p)) return LSSPA_E_UNSUPPORTED;
  if (count > 0 && !perms) return LSSPA_E_BADARG;
  if (mode != 2 && !gram) return LSSPA_E_BADARG;
  if (mode != 1 && (!R_te_cm || !c_te || !lifts_out || !(y_norm_sq > 0.0))) return LSSPA_E_BADARG;
  if (mode != 0 && !fact) return LSSPA_E_BADARG;
  if (count == 0) return LSSPA_OK;
  if (count < 0) return LSSPA_E_BADARG;
  CholParams a;
  a.p = p;
  a.rt = (p + 7) / 8;
  a.pt = (p + 8) / 8;
  a.ld = chol_ld(a.rt);
  a.Gh = gram;
  a.Rte = R_te_cm;
  a.cte = c_te;
  a.inv_ynsq = (mode != 1) ? 1.0 / y_norm_sq : 0.0;
  a.perms = perms;
  a.count = count;
  a.anti = antithetical ? 1 : 0;
  a.out = lifts_out;
  a.fact = fact;
  a.dbg = (mode == 0) ? g_lifts_dbg : nullptr;
  const size_t smem = chol_smem_bytes(p);
  const DeviceInfo &d = device_info();
  const int sms = d.sm_count > 0 ? d.sm_count : 148;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  const int cap = a.rt <= 6 ? 3 : 2;   // = MINB of the instantiation
  if (per_sm > cap) per_sm = cap;
  int64_t grid = (int64_t)per_sm * sms;
  if (grid > count) grid = count;
  cudaStream_t st = as_stream(stream);
  static const bool packed = [] {
    const char *e = getenv("LSSPA_CHOL_PACKED");     // 0: the eight-warp kernel everywhere (A/B timing)
    return !(e && e[0] == '0');
  }();
  if (packed) {
    switch (a.rt) {
      case 7: return run_chol4<7>(a, count, mode, st);
      case 8: return run_chol4<8>(a, count, mode, st);
      case 9: return run_chol4<9>(a, count, mode, st);
      case 10: return run_chol4<10>(a, count, mode, st);
      case 11: return run_chol4<11>(a, count, mode, st);
      case 12: return run_chol4<12>(a, count, mode, st);
      case 13: return run_chol4<13>(a, count, mode, st);
      case 14: return run_chol4<14>(a, count, mode, st);
      case 15: return run_chol4<15>(a, count, mode, st);
      case 16: return run_chol4<16>(a, count, mode, st);
      default: break;
    }
  }
  switch (a.rt) {
    case 3: return launch_chol_rt<3>(a, (int)grid, smem, mode, st);
    case 4: return launch_chol_rt<4>(a, (int)grid, smem, mode, st);
    case 5: return launch_chol_rt<5>(a, (int)grid, smem, mode, st);
    case 6: return launch_chol_rt<6>(a, (int)grid, smem, mode, st);
    case 7: return launch_chol_rt<7>(a, (int)grid, smem, mode, st);
    case 8: return launch_chol_rt<8>(a, (int)grid, smem, mode, st);
    case 9: return launch_chol_rt<9>(a, (int)grid, smem, mode, st);
    case 10: return launch_chol_rt<10>(a, (int)grid, smem, mode, st);
    case 11: return launch_chol_rt<11>(a, (int)grid, smem, mode, st);
    case 12: return launch_chol_rt<12>(a, (int)grid, smem, mode, st);
    case 13: return launch_chol_rt<13>(a, (int)grid, smem, mode, st);
    case 14: return launch_chol_rt<14>(a, (int)grid, smem, mode, st);
    case 15: return launch_chol_rt<15>(a, (int)grid, smem, mode, st);
    case 16: return launch_chol_rt<16>(a, (int)grid, smem, mode, st);
    case 17: return launch_chol_rt<17>(a, (int)grid, smem, mode, st);
    case 18: return launch_chol_rt<18>(a, (int)grid, smem, mode, st);
    case 19: return launch_chol_rt<19>(a, (int)grid, smem, mode, st);
  }
  return LSSPA_E_UNSUPPORTED;
}

extern "C" int lsspa_lifts_chol(int p, const double *gram, const double *R_te_cm, const double *c_te,
                                double y_norm_sq, const int32_t *perms, int64_t count, int antithetical,
                                double *lifts_out, void *stream) {
  return chol_run(0, p, gram, R_te_cm, c_te, y_norm_sq, perms, count, antithetical, lifts_out, nullptr, stream);
}

extern "C" int64_t lsspa_lifts_chol_factor_doubles(int p) {
  if (p < 49 || p > 128) return 0;   // the split route
  return chol_fact_doubles((p + 7) / 8, (p + 8) / 8);
}

extern "C" int lsspa_lifts_chol_factor(int p, const double *gram, const int32_t *perms, int64_t count,
                                       int antithetical, double *factors_out, void *stream) {
  if (lsspa_lifts_chol_factor_doubles(p) == 0) return LSSPA_E_UNSUPPORTED;
  return chol_run(1, p, gram, nullptr, nullptr, 0.0, perms, count, antithetical, nullptr, factors_out, stream);
}

extern "C" int lsspa_lifts_chol_eliminate(int p, const double *factors, const double *R_te_cm, const double *c_te,
                                          double y_norm_sq, const int32_t *perms, int64_t count, int antithetical,
                                          double *lifts_out, void *stream) {
  if (lsspa_lifts_chol_factor_doubles(p) == 0) return LSSPA_E_UNSUPPORTED;
  return chol_run(2, p, nullptr, R_te_cm, c_te, y_norm_sq, perms, count, antithetical, lifts_out,
                  const_cast<double *>(factors), stream);
}
