// Per-permutation core of LS-SPA, tensor-pipe version (mma.sync m8n8k4 f64 = DMMA on sm_100a;
// tcgen05 has no fp64, so this is the tensor path the FP64 pipe offers).
//
// Replaces square_shapley (reference ls_spa/ls_spa.py:256-287) and the antithetic pair average
// (:205-208) for 49 <= p <= 128, any conditioning (the Householder route; well-conditioned problems
// take lifts_chol.cu).  Same mathematics as lifts.cu, reorganised so that almost all
// flops are 8x8x4 matrix products and the number of block-wide barriers drops from 2p to 2p/8:
//
//  phase 1  blocked Householder QR of A = [R_tr[:, perm] | c_tr] (shared memory, column-major,
//           ld % 16 == 8).  Per panel of 8 columns: four warps factor the panel cooperatively in
//           registers (reflectors V, compact-WY factor T from G = V^T V) one panel ahead, while
//           the other four take whole trailing column tiles:  W^T = (A2^T V) T  and
//           A2^T -= W^T V^T, all DMMA.
//  phase 2  elimination of X = R_te[:, perm] against the rows of R, one warp per 8-row tile of
//           X held entirely in registers (rows of X are independent):  M_J = X_J R_JJ^-1,
//           X_L -= M_J R_JL.  The multipliers M are the columns of W = X R^-1 (reference :279),
//           so the prefix residuals / costs (:282-283) come from a running residual per row.
//
// Fragment conventions (lane = 4c + q):  A[m=c][k=q], B[k=q][n=c], C[m=c][n=2q+e].
// "tile access": lane touches the 16 bytes M[i0+2q .. i0+2q+1][j0+c] of a column-major matrix;
// with ld % 16 == 8 the 32 lanes hit 512 distinct bytes in 4 conflict-free wavefronts.  Such a
// fragment is at once the B operand of the tile, the A operand of its transpose and the C
// operand of its transpose (k index <-> row 2q+e in both k-steps), and a C result is reused as
// the A operand of the next product without leaving the registers.
// The lane-level Python model of this file is tests/lifts_v2_model.py.

#include "common.cuh"

namespace lsspa {

struct LiftParams2 {
  int p;
  int ld;       // leading dimension of A and Vb (>= 8*RT, ld % 16 == 8)
  int rt;       // row tiles      ceil(p / 8)
  int pt;       // column tiles   ceil((p + 1) / 8)
  const double *Rtr;  // column-major p x p
  const double *ctr;
  const double *Rte;  // column-major p x p
  const double *cte;
  double inv_ynsq;
  const int32_t *perms;
  int64_t count;
  int anti;
  double *out;
  long long *dbg;  // optional: cycle counters of block 0 (development aid), else nullptr
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

constexpr int kVS = 10;  // row stride (doubles) of the row-major V buffers: rows 2q hit banks 8q

// V fragment for the tile access pattern (rows 2q+e of the tile, column c): B operand of V /
// A operand of V^T.  Two 8-byte loads, conflict-free with kVS = 10.
__device__ __forceinline__ double2 ld_vfrag(const double *V, int r0, int c, int q) {
  const double *ptr = V + (size_t)(r0 + 2 * q) * kVS + c;
  return make_double2(ptr[0], ptr[kVS]);
}

// ---------------------------------------------------------------- trailing column tile j
// A2[:, tile j] -= V T^T (V^T A2[:, tile j]) for the reflectors of panel s
__device__ __forceinline__ void trailing_tile(double *A, const double *V, const double *Tb, int ld, int RT,
                                              int s, int j, int lane) {
  const int c = lane >> 2, q = lane & 3;
  double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
  int t = s;
  for (; t + 1 < RT; t += 2) {
    const double2 x = ld_tile(A, ld, 8 * t, 8 * j, c, q);
    const double2 v = ld_vfrag(V, 8 * t, c, q);
    const double2 x2 = ld_tile(A, ld, 8 * t + 8, 8 * j, c, q);
    const double2 v2 = ld_vfrag(V, 8 * t + 8, c, q);
    dmma(a0, a1, x.x, v.x);
    dmma(b0, b1, x2.x, v2.x);
    dmma(a0, a1, x.y, v.y);
    dmma(b0, b1, x2.y, v2.y);
  }
  if (t < RT) {
    const double2 x = ld_tile(A, ld, 8 * t, 8 * j, c, q);
    const double2 v = ld_vfrag(V, 8 * t, c, q);
    dmma(a0, a1, x.x, v.x);
    dmma(a0, a1, x.y, v.y);
  }
  a0 += b0;  // C layout of Wraw^T: lane (m = column of the tile, n = reflector 2q+e)
  a1 += b1;
  const double2 tt = ld_tile(Tb, 8, 0, 0, c, q);
  double w0 = 0.0, w1 = 0.0;  // W'^T = Wraw^T T
  dmma(w0, w1, a0, tt.x);
  dmma(w0, w1, a1, tt.y);
  w0 = -w0;
  w1 = -w1;
  for (t = s; t < RT; ++t) {
    double2 cf = ld_tile(A, ld, 8 * t, 8 * j, c, q);
    // B[k <-> reflector 2q+e][n = row c] = V[row][2q+e]: 16 contiguous bytes of a V row
    const double2 vt = *reinterpret_cast<const double2 *>(V + (size_t)(8 * t + c) * kVS + 2 * q);
    dmma(cf.x, cf.y, w0, vt.x);
    dmma(cf.x, cf.y, w1, vt.y);
    st_tile(A, ld, 8 * t, 8 * j, c, q, cf);
  }
}

// ---------------------------------------------------------------- cooperative panel (4 warps)
// The reflector loop of a panel is a serial chain (norm -> sqrt/rcp -> dot -> update, ~1000 cycles
// per reflector for one warp holding the whole strip).  Four "panel warps" share a strip by rows
// (warp w holds tiles w, w+4, w+8, ... relative to the top tile): the per-reflector data work
// shrinks 4x and the two reductions (column norm, dot products) are completed through a small
// shared-memory exchange and a 128-thread named barrier.  The other four warps of the CTA apply
// the previous panel to the remaining column tiles meanwhile.
constexpr int kPW = 4;  // panel warps (warps 0..3)

__device__ __forceinline__ void bar_panel() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct PanelXch {      // exchange scratch in shared memory (doubles)
  double *pn;          // [2][kPW]      partial squared norms (double-buffered by reflector parity)
  double *x0;          // [2]           pivot value
  double *pw;          // [kPW][8]      partial dot products
  double *gp;          // [kPW][64]     partial G / partial W^T (C layout)
  double *vs;          // [kPW][8*KL]   per-warp u scratch column
};

// A2[:, tile j] -= U T^T (U^T A2[:, tile j]) for the reflectors of panel s, rows split over the
// panel warps (cooperative version of trailing_tile)
template <int KL>
__device__ __forceinline__ void trailing_coop(double *A, const double *V, const double *Tb, const PanelXch &x,
                                              int ld, int RT, int s, int j, int lane, int w) {
  const int c = lane >> 2, q = lane & 3;
  const int nt = RT - s;
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int kk = 0; kk < KL; ++kk) {
    const int k = w + kPW * kk;
    if (k < nt) {
      const double2 xx = ld_tile(A, ld, 8 * (s + k), 8 * j, c, q);
      const double2 v = ld_vfrag(V, 8 * (s + k), c, q);
      dmma(a0, a1, xx.x, v.x);
      dmma(a0, a1, xx.y, v.y);
    }
  }
  *reinterpret_cast<double2 *>(x.gp + w * 64 + c * 8 + 2 * q) = make_double2(a0, a1);
  bar_panel();
  a0 = 0.0;
  a1 = 0.0;
#pragma unroll
  for (int ww = 0; ww < kPW; ++ww) {
    const double2 t = *reinterpret_cast<const double2 *>(x.gp + ww * 64 + c * 8 + 2 * q);
    a0 += t.x;
    a1 += t.y;
  }
  const double2 tt = ld_tile(Tb, 8, 0, 0, c, q);
  double w0 = 0.0, w1 = 0.0;  // W'^T = Wraw^T T
  dmma(w0, w1, a0, tt.x);
  dmma(w0, w1, a1, tt.y);
  w0 = -w0;
  w1 = -w1;
#pragma unroll
  for (int kk = 0; kk < KL; ++kk) {
    const int k = w + kPW * kk;
    if (k < nt) {
      double2 cf = ld_tile(A, ld, 8 * (s + k), 8 * j, c, q);
      const double2 vt = *reinterpret_cast<const double2 *>(V + (size_t)(8 * (s + k) + c) * kVS + 2 * q);
      dmma(cf.x, cf.y, w0, vt.x);
      dmma(cf.x, cf.y, w1, vt.y);
      st_tile(A, ld, 8 * (s + k), 8 * j, c, q, cf);
    }
  }
  bar_panel();  // the tile is complete (and gp free) before anyone reloads it with another ownership
}

// Factor panel s (column tile s, row tiles s..RT-1) with the four panel warps.  Same outputs as the
// single-warp version: R (top tile) to A, U row-major to V, T (column-major 8x8) to Tb.
template <int KL>
__device__ __forceinline__ void panel_coop(double *A, double *V, double *Tb, const PanelXch &x, int ld, int p,
                                           int RT, int s, int lane, int w) {
  const int c = lane >> 2, q = lane & 3;
  const int j0 = 8 * s, r0 = 8 * s;
  const int nf = (p - j0 < 8) ? p - j0 : 8;
  const int nt = RT - s;
  const int l0 = 2 * q, l1 = 2 * q + 1;
  const bool top = (w == 0);  // this warp's local tile 0 is the top tile of the strip
  double vr[KL][2];
  double na = 0.0, nb = 0.0;
  const double *Acol = A + (size_t)(j0 + c) * ld + r0 + 2 * q;
#pragma unroll
  for (int kk = 0; kk < KL; ++kk) {
    const int k = w + kPW * kk;
    double2 v = make_double2(0.0, 0.0);
    if (k < nt) v = *reinterpret_cast<const double2 *>(Acol + 8 * k);
    vr[kk][0] = v.x;
    vr[kk][1] = v.y;
    if (kk == 0 && top) {
      if (l0 > c) na = v.x * v.x;
      if (l1 > c) nb = v.y * v.y;
    } else {
      na = fma(v.x, v.x, na);
      nb = fma(v.y, v.y, nb);
    }
  }
  double tau_r[8];
  double dg = 0.0, du = 0.0;
  double2 *vs2 = reinterpret_cast<double2 *>(x.vs + w * (8 * KL)) + q;
  // ONE exchange per reflector.  u = x - beta e1 is unnormalised, so everything a warp needs for its
  // partial dot products (the raw sub-pivot entries of column cc) is known before beta is: each
  // warp publishes {partial |x|^2, partial u^T a_c without the pivot row}, warp 0 adds the pivot
  // row a_piv,c, and after the barrier every warp derives beta, u1, tt and w_c = tt (sum + u1 a_piv,c).
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) {
    tau_r[cc] = 0.0;
    if (cc < nf) {
      double *ex = x.pw + (cc & 1) * 48;   // [kPW][8] partial dots, [kPW] partial norms, [8] pivot row
      __syncwarp();                        // every lane is done reading the previous reflector's column
      if (c == cc) {
        // publish the raw column (top tile: zeros on and above the pivot row)
        if (top)
          vs2[0] = make_double2((l0 <= cc) ? 0.0 : vr[0][0], (l1 <= cc) ? 0.0 : vr[0][1]);
        else
          vs2[0] = make_double2(vr[0][0], vr[0][1]);
#pragma unroll
        for (int kk = 1; kk < KL; ++kk) vs2[4 * kk] = make_double2(vr[kk][0], vr[kk][1]);
      }
      __syncwarp();
      double wa = 0.0, wb = 0.0;
#pragma unroll
      for (int kk = 0; kk < KL; ++kk) {
        const double2 u = vs2[4 * kk];
        wa = fma(u.x, vr[kk][0], wa);
        wb = fma(u.y, vr[kk][1], wb);
      }
      // both reductions over the quad at once: lanes of column cc carry the norm, all carry the dot
      double wpart = wa + wb, npart = na + nb;
      wpart += __shfl_xor_sync(kFull, wpart, 1);
      npart += __shfl_xor_sync(kFull, npart, 1);
      wpart += __shfl_xor_sync(kFull, wpart, 2);
      npart += __shfl_xor_sync(kFull, npart, 2);
      if (q == 0) ex[w * 8 + c] = wpart;
      if (lane == 4 * cc) ex[32 + w] = npart;
      if (top && q == (cc >> 1)) ex[36 + c] = (cc & 1) ? vr[0][1] : vr[0][0];   // pivot-row entries a_piv,c
      bar_panel();
      const double sig = (ex[32] + ex[33]) + (ex[34] + ex[35]);
      const double x0 = ex[36 + cc];
      if (sig > kTinySig) {  // uniform over the four warps
        // beta = -sign(x0) |x|, u1 = x0 - beta (|u1| = |x0| + |x|), tt = 1 / (|x| |u1|)
        // (dependent fp64 operations cost ~20 cycles each here, so the chain is kept short:
        //  one rsqrt and one reciprocal, 1/(|x| |u1|) = 1/(|x|^2 + |x0| |x|))
        const double s2 = fma(x0, x0, sig);
        const double nrm = s2 * rsqrt(s2);
        const double tt = 1.0 / fma(fabs(x0), nrm, s2);
        const double beta = (x0 >= 0.0) ? -nrm : nrm;
        const double u1 = x0 - beta;
        tau_r[cc] = tt;
        if (c == cc && top) {
          dg = beta;
          du = u1;
        }
        const double apiv = ex[36 + c];
        const double wt = tt * (((ex[c] + ex[8 + c]) + (ex[16 + c] + ex[24 + c])) + u1 * apiv);
        if (c > cc) {
          na = 0.0;
          nb = 0.0;
#pragma unroll
          for (int kk = 0; kk < KL; ++kk) {
            const double2 u = vs2[4 * kk];
            vr[kk][0] = fma(-wt, u.x, vr[kk][0]);
            vr[kk][1] = fma(-wt, u.y, vr[kk][1]);
            if (kk == 0 && top) {
              // the pivot row is not part of the published column: a_piv,c -= wt * u1
              if (l0 == cc) vr[0][0] = fma(-wt, u1, vr[0][0]);
              if (l1 == cc) vr[0][1] = fma(-wt, u1, vr[0][1]);
              if (l0 > c) na = vr[0][0] * vr[0][0];
              if (l1 > c) nb = vr[0][1] * vr[0][1];
            } else {
              na = fma(vr[kk][0], vr[kk][0], na);
              nb = fma(vr[kk][1], vr[kk][1], nb);
            }
          }
        }
      } else if (c == cc && top) {
        dg = x0;  // column already zero below the pivot: H = I
        du = 0.0;
      }
    }
  }
  // U (row-major) for the trailing updates and R / untouched columns back to A
  {
    const bool fact = c < nf;
    // a column carries a reflector iff its tau is non-zero (known to every warp)
    double mytau = 0.0;
#pragma unroll
    for (int cc = 0; cc < 8; ++cc)
      if (c == cc) mytau = tau_r[cc];
    const bool refl = fact && (mytau != 0.0);
    double *Vc = V + (size_t)(r0 + l0) * kVS + c;
    double *Aw = A + (size_t)(j0 + c) * ld + r0 + 2 * q;
#pragma unroll
    for (int kk = 0; kk < KL; ++kk) {
      const int k = w + kPW * kk;
      if (k < nt) {
        if (kk == 0 && top) {
          Vc[0] = refl ? ((l0 < c) ? 0.0 : ((l0 == c) ? du : vr[0][0])) : 0.0;
          Vc[kVS] = refl ? ((l1 < c) ? 0.0 : ((l1 == c) ? du : vr[0][1])) : 0.0;
          double2 v = make_double2(vr[0][0], vr[0][1]);
          if (fact) {
            v.x = (l0 < c) ? vr[0][0] : ((l0 == c) ? dg : 0.0);
            v.y = (l1 < c) ? vr[0][1] : ((l1 == c) ? dg : 0.0);
          }
          *reinterpret_cast<double2 *>(Aw) = v;
        } else {
          Vc[(size_t)8 * k * kVS] = refl ? vr[kk][0] : 0.0;
          Vc[(size_t)8 * k * kVS + kVS] = refl ? vr[kk][1] : 0.0;
          *reinterpret_cast<double2 *>(Aw + 8 * k) =
              fact ? make_double2(0.0, 0.0) : make_double2(vr[kk][0], vr[kk][1]);
        }
      }
    }
  }
  __syncwarp();
  // partial G = U^T U over this warp's tiles
  double g0 = 0.0, g1 = 0.0;
  {
    const double *Vf = V + (size_t)(r0 + 2 * q) * kVS + c;
#pragma unroll
    for (int kk = 0; kk < KL; ++kk) {
      const int k = w + kPW * kk;
      if (k < nt) {
        const double vx = Vf[(size_t)8 * k * kVS];
        const double vy = Vf[(size_t)8 * k * kVS + kVS];
        dmma(g0, g1, vx, vx);
        dmma(g0, g1, vy, vy);
      }
    }
  }
  *reinterpret_cast<double2 *>(x.gp + w * 64 + c * 8 + 2 * q) = make_double2(g0, g1);  // [m][n] row-major
  bar_panel();
  // T (dlarft, forward / columnwise) by warp 0: the 32 lanes first add the four partial G's, then
  // lane u < 8 runs the recurrence of row u with two accumulators (dependent fp64 ops are ~20
  // cycles each and this sits on the critical path of the panel chain)
  if (w == 0) {
    {
      const int o = c * 8 + 2 * q;
      const double2 p0 = *reinterpret_cast<const double2 *>(x.gp + o);
      const double2 p1 = *reinterpret_cast<const double2 *>(x.gp + 64 + o);
      const double2 p2 = *reinterpret_cast<const double2 *>(x.gp + 128 + o);
      const double2 p3 = *reinterpret_cast<const double2 *>(x.gp + 192 + o);
      __syncwarp();
      *reinterpret_cast<double2 *>(x.gp + o) = make_double2((p0.x + p1.x) + (p2.x + p3.x), (p0.y + p1.y) + (p2.y + p3.y));
      __syncwarp();
    }
    if (lane < 8) {
      const int u = lane;
      double tr[8];
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        double val = 0.0;
        if (cc == u) {
          val = tau_r[cc];
        } else if (cc > u) {
          double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k >= u && k < cc) {
              if (k & 1) acc1 = fma(tr[k], x.gp[k * 8 + cc], acc1);
              else acc0 = fma(tr[k], x.gp[k * 8 + cc], acc0);
            }
          val = -tau_r[cc] * (acc0 + acc1);
        }
        tr[cc] = val;
        Tb[cc * 8 + u] = val;
      }
    }
  }
}

// ---------------------------------------------------------------- kernel
template <int MAXT, int MINB>
__global__ void __launch_bounds__(256, MINB) lifts_mma_kernel(LiftParams2 a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int p = a.p, ld = a.ld, RT = a.rt, PT = a.pt;
  const int NR = 8 * RT, NC = 8 * PT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = lane >> 2, q = lane & 3;

  double *A = reinterpret_cast<double *>(smem_raw);
  double *V0 = A + (size_t)NC * ld;        // NR x kVS, double-buffered   (phase 2: Dbuf, RT x 64)
  double *V1 = V0 + (size_t)NR * kVS;      //                             (phase 2: cost partials, 8 x NR)
  double *Tb = V1 + (size_t)NR * kVS;      // 2 x 64
  double *Gs = Tb + 128;                   // 64
  double *vs = Gs + 64;                    // 128: per-warp u scratch columns of the panel warps
  double *xch = vs + 128;                  // 16 (unused) + 96 (2 x {32 dots, 4 norms, 8 pivot row, pad}) + 256 (gp)
  double *cost = xch + 368;                // p + 1
  PanelXch px;
  px.pn = xch;
  px.x0 = xch + 8;
  px.pw = xch + 16;
  px.gp = xch + 112;
  px.vs = vs;
  constexpr int KL = (MAXT + kPW - 1) / kPW;
  double *acc = cost + (p + 2);            // p
  int *perm_s = reinterpret_cast<int *>(acc + p + (p & 1));
  double *Dbuf = V0;
  double *wcost = V1;

  const int halves = a.anti ? 2 : 1;
  const double weight = a.anti ? 0.5 : 1.0;

  for (int64_t sidx = blockIdx.x; sidx < a.count; sidx += gridDim.x) {
    for (int h = 0; h < halves; ++h) {
      __syncthreads();
      for (int k = tid; k < p; k += 256) perm_s[k] = a.perms[sidx * p + (h == 0 ? k : p - 1 - k)];
      __syncthreads();
      long long t_a = LSSPA_CLOCK();
      // ---- phase 0: gather A = [R_tr[:, perm] | c_tr], zero padding
      {
        const int half = NR / 2, tot = NC * half;
#pragma unroll 4
        for (int e = tid; e < tot; e += 256) {
          const int k = e / half, i = 2 * (e - k * half);
          double2 v = make_double2(0.0, 0.0);
          if (k < p) {
            const int col = perm_s[k];
            const double *src = a.Rtr + (size_t)col * p;
            if (i <= col) v.x = __ldg(src + i);
            if (i + 1 <= col) v.y = __ldg(src + i + 1);
          } else if (k == p) {
            if (i < p) v.x = __ldg(a.ctr + i);
            if (i + 1 < p) v.y = __ldg(a.ctr + i + 1);
          }
          *reinterpret_cast<double2 *>(A + (size_t)k * ld + i) = v;
        }
      }
      if (warp == 7) {
        double s0 = 0.0;
        for (int i = lane; i < p; i += 32) s0 = fma(a.cte[i], a.cte[i], s0);
        s0 = warp_sum(s0);
        if (lane == 0) cost[0] = s0;
      }
      __syncthreads();

      // ---- phase 1: blocked Householder with one panel of look-ahead.  In step s warp 0 first
      // brings column tile s+1 up to date and factors it (into the other V/T buffer) while
      // warps 1..7 apply panel s to the remaining tiles: one barrier per panel.
      long long t_b = LSSPA_CLOCK(), t_pan = 0, t_trc = 0, t_wait = 0;
      if (warp < kPW) panel_coop<KL>(A, V0, Tb, px, ld, p, RT, 0, lane, warp);
      t_pan += LSSPA_CLOCK() - t_b;
      __syncthreads();
      for (int s = 0; s < RT; ++s) {
        const double *Vc = (s & 1) ? V1 : V0;
        double *Vn = (s & 1) ? V0 : V1;
        const double *Tc = Tb + 64 * (s & 1);
        double *Tn = Tb + 64 * ((s + 1) & 1);
        long long t0 = LSSPA_CLOCK();
        if (warp < kPW) {
          if (s + 1 < PT) trailing_coop<KL>(A, Vc, Tc, px, ld, RT, s, s + 1, lane, warp);
          long long t1 = LSSPA_CLOCK();
          t_trc += t1 - t0;
          if (s + 1 < RT) panel_coop<KL>(A, Vn, Tn, px, ld, p, RT, s + 1, lane, warp);
          t_pan += LSSPA_CLOCK() - t1;
        } else {
          for (int j = s + 2 + (warp - kPW); j < PT; j += 8 - kPW) trailing_tile(A, Vc, Tc, ld, RT, s, j, lane);
          t_trc += LSSPA_CLOCK() - t0;
        }
        long long t2 = LSSPA_CLOCK();
        __syncthreads();
        t_wait += LSSPA_CLOCK() - t2;
      }
      long long t_c = LSSPA_CLOCK();

      // ---- phase 1.5: inverses of the 8x8 diagonal blocks of R (column-major 8x8 each), zero the
      //      per-warp cost partials (both overlay the V buffers)
      if (tid < RT * 8) {
        const int J = tid >> 3, jj = tid & 7;
        double x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = 0.0;
        if (8 * J + jj < p) {
          const double *D = A + (size_t)(8 * J) * ld + 8 * J;  // D[u][k] = D[k * ld + u]
#pragma unroll
          for (int u = 7; u >= 0; --u) {
            if (u == jj) {
              x[u] = 1.0 / D[(size_t)u * ld + u];
            } else if (u < jj) {
              double sacc = 0.0;
#pragma unroll
              for (int k = 0; k < 8; ++k)
                if (k > u && k <= jj) sacc = fma(D[(size_t)k * ld + u], x[k], sacc);
              x[u] = -sacc / D[(size_t)u * ld + u];
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) Dbuf[J * 64 + jj * 8 + u] = x[u];
      }
      for (int e = tid; e < 8 * NR; e += 256) wcost[e] = 0.0;
      __syncthreads();

      // ---- phase 2: X = R_te[:, perm] eliminated against R, one warp per 8-row tile, registers only
      double *wc = wcost + (size_t)warp * NR;
      const double *cvec = A + (size_t)p * ld;
      for (int it = warp; it < RT; it += 8) {
        const int row = 8 * it + c;  // C layout: lane (m = row, n = columns 2q+e)
        double xr[MAXT][2];
#pragma unroll
        for (int L = 0; L < MAXT; ++L) {
          xr[L][0] = 0.0;
          xr[L][1] = 0.0;
          if (L < RT) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int l = 8 * L + 2 * q + e;
              if (l < p) {
                const int col = perm_s[l];
                if (row <= col) xr[L][e] = a.Rte[(size_t)col * p + row];
              }
            }
          }
        }
        double r_in = (row < p) ? a.cte[row] : 0.0;
#pragma unroll
        for (int J = 0; J < MAXT; ++J) {
          if (J < RT) {
            const double2 dv = ld_tile(Dbuf + J * 64, 8, 0, 0, c, q);
            double m0 = 0.0, m1 = 0.0;  // M_J = X_J R_JJ^-1, C layout (row, column 2q+e of the panel)
            dmma(m0, m1, xr[J][0], dv.x);
            dmma(m0, m1, xr[J][1], dv.y);
            // running test residual of this row after each of the 8 columns of the panel
            // residual of row c after each of the 8 columns of the panel as one more product with the
            // upper-triangular matrix whose row k holds c_k (B fragment: k = 2q + e, n = c)
            const double2 cv = *reinterpret_cast<const double2 *>(cvec + 8 * J + 2 * q);
            double ra = r_in, rb = r_in;
            dmma(ra, rb, -m0, (2 * q <= c) ? cv.x : 0.0);
            dmma(ra, rb, -m1, (2 * q + 1 <= c) ? cv.y : 0.0);
            // row sums of both squares with three shuffles: the first round hands each value to one
            // half of the lanes; even rows end up with the first column's total, odd rows with the second's
            const double d0 = ra * ra, d1 = rb * rb;
            const bool odd = (c & 1) != 0;
            double tot = (odd ? d1 : d0) + __shfl_xor_sync(kFull, odd ? d0 : d1, 4);
            tot += __shfl_xor_sync(kFull, tot, 8);
            tot += __shfl_xor_sync(kFull, tot, 16);
            if (c < 2) {
              const int k0 = 8 * J + 2 * q + c;
              if (k0 < p) wc[k0] += tot;          // wc[k] collects cost_{k+1}
            }
            r_in = __shfl_sync(kFull, rb, 3, 4);   // residual after the last column of the panel
            m0 = -m0;
            m1 = -m1;
#pragma unroll
            for (int L = 0; L < MAXT; ++L) {
              if (L > J && L < RT) {
                const double2 rt = ld_tile(A, ld, 8 * J, 8 * L, c, q);
                dmma(xr[L][0], xr[L][1], m0, rt.x);
                dmma(xr[L][0], xr[L][1], m1, rt.y);
              }
            }
          }
        }
      }
      long long t_d = LSSPA_CLOCK();
      __syncthreads();
#ifdef LSSPA_LIFTS_TIMING
      if (a.dbg != nullptr && blockIdx.x == 0 && lane == 0 && sidx == blockIdx.x && h == 0) {
        long long *d = a.dbg + warp * 8;
        d[0] = t_b - t_a;      // gather
        d[1] = t_pan;          // panel (warps 0-3)
        d[2] = t_trc;          // trailing work
        d[3] = t_wait;         // waiting at the step barrier
        d[4] = t_c - t_b;      // whole phase 1
        d[5] = t_d - t_c;      // phase 1.5 + phase 2 (this warp)
      }
#else
      (void)t_a; (void)t_b; (void)t_c; (void)t_d; (void)t_pan; (void)t_trc; (void)t_wait;
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
    __syncthreads();
    for (int f = tid; f < p; f += 256) a.out[sidx * p + f] = acc[f];
  }
}

static int mma_ld(int rt) {
  int n = 8 * rt;
  return (n % 16 == 8) ? n : n + 8;
}

static size_t mma_smem_bytes(int p) {
  const int rt = (p + 7) / 8, pt = (p + 8) / 8;
  const int ld = mma_ld(rt);
  size_t d = (size_t)8 * pt * ld + 2 * (size_t)8 * rt * 10 + 192 + 128 + 368 + (size_t)(p + 2) + (size_t)(p + 1);
  return d * sizeof(double) + (size_t)p * sizeof(int) + 32;
}

template <int MAXT, int MINB>
static int launch_mma(const LiftParams2 &a, int grid, size_t smem, cudaStream_t st) {
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(lifts_mma_kernel<MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(lifts_mma_kernel<MAXT, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
  lifts_mma_kernel<MAXT, MINB><<<grid, 256, smem, st>>>(a);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

// below ~48 features the scalar kernel (many small CTAs per SM) is faster (profiles/r01_quick_bench_v2.log)
long long *g_lifts_dbg = nullptr;  // set through lsspa_debug_set_lifts_counters (development aid)

bool lifts_mma_supported(int p) { return p >= 49 && p <= 128; }

int lifts_mma_launch(int p, const double *R_tr_cm, const double *c_tr, const double *R_te_cm, const double *c_te,
                     double y_norm_sq, const int32_t *perms, int64_t count, int antithetical, double *lifts_out,
                     cudaStream_t st) {
  LiftParams2 a;
  a.p = p;
  a.rt = (p + 7) / 8;
  a.pt = (p + 8) / 8;
  a.ld = mma_ld(a.rt);
  a.Rtr = R_tr_cm;
  a.ctr = c_tr;
  a.Rte = R_te_cm;
  a.cte = c_te;
  a.inv_ynsq = 1.0 / y_norm_sq;
  a.perms = perms;
  a.count = count;
  a.anti = antithetical ? 1 : 0;
  a.out = lifts_out;
  a.dbg = g_lifts_dbg;
  const size_t smem = mma_smem_bytes(p);
  const DeviceInfo &d = device_info();
  const int sms = d.sm_count > 0 ? d.sm_count : 148;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int64_t grid = (int64_t)per_sm * sms;
  if (grid > count) grid = count;
  if (a.rt <= 4) return launch_mma<4, 3>(a, (int)grid, smem, st);
  if (a.rt <= 8) return launch_mma<8, 2>(a, (int)grid, smem, st);
  if (a.rt <= 13) return launch_mma<13, 2>(a, (int)grid, smem, st);
  return launch_mma<16, 1>(a, (int)grid, smem, st);
}

}  // namespace lsspa

#ifdef LSSPA_LIFTS_TIMING
// development aid (tools/lifts_cycles.py), only in builds with -DLSSPA_LIFTS_TIMING and therefore not
// part of include/lsspa.h: 64 device long longs receiving block 0's cycle counters
extern "C" __attribute__((visibility("default"))) void lsspa_debug_set_lifts_counters(long long *dev_ptr) {
  lsspa::g_lifts_dbg = dev_ptr;
}
#endif
