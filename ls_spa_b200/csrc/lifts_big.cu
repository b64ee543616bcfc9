// Per-permutation core of LS-SPA for WIDE problems (p > 152: the tile array of one permutation no
// longer fits one SM's shared memory; BASELINE.json config 5 is p = 1000).
//
// Same mathematics as lifts_chol.cu (reference ls_spa/ls_spa.py:256-287 through the Cholesky factor
// of the permuted Gram matrix), organised as a BATCHED blocked factorisation over an L2/HBM-resident
// tile workspace, many permutations in flight:
//
//   M  = [ Gh[pi^, pi^] | X^T ]      q x (q + p),  q = p + 1,  pi^ = (perm, p),  X = R_te[:, perm]
//   left-looking block row j (64 rows):   C_jc = M_jc - sum_{k<j} R_kj^T R_kc      for every column tile c
//                                         R_jj = chol(C_jj),   R_jc = R_jj^-T C_jc
//   the G part yields R and c~ = R[:p, p] (= Q^T c_tr of the reference, :278); the X part yields
//   V = R^-T X^T = W^T, the transposed multipliers of the reference's triangular solve (:279-283);
//   cost_{k+1} = sum_i (c_te[i] - sum_{j<=k} c~_j V[j][i])^2, lift_k = (cost_k - cost_{k+1}) / |y_te|^2.
//
// Workspace layout ("fragment-major" tiles): a 64 x 64 tile is 8 x 8 micro-tiles of 8 x 8 doubles,
// micro-tile (kk, nn) at (kk * 8 + nn) * 64, column-major inside (element (r, c) at c * 8 + r).  Lane
// l of a warp reads the 16 bytes at 2 l: that double2 is at once the B fragment (k = 2q + e, n = c) of
// the micro-tile and the A fragment of its transpose for the two m8n8k4 DMMAs covering k = 0..7
// (c = l >> 2, q = l & 3), and a C fragment is stored with the same addressing -- every shared and
// global access of the hot loop is a contiguous 512-byte warp access.  A 16-row slab of a tile
// (two micro-rows) is 8 KB contiguous: the unit the TMA moves (cp.async.bulk + mbarrier pipeline).
//
// Kernels per batch of evaluations: big_gather (permuted Gram and test columns -> tiles), then per
// block row big_diag (diagonal tile: update, 64 x 64 Cholesky and inverse in shared memory) and
// big_panel (all other tiles of the block row: CTA = 4 column tiles x one evaluation, 8 DMMA warps
// + 1 TMA producer warp, 4-stage ring), finally big_cost (residual recurrence, lifts, scatter).
// Bound: FP64 pipe (DMMA); operands stream from L2 at ~13 flop/B per CTA.

#include "common.cuh"

namespace lsspa {
namespace {

constexpr int kNB = 64;                  // tile edge
constexpr int kTileD = kNB * kNB;        // doubles per tile (32 KB)
constexpr int kSlabRows = 16;            // k rows per pipeline stage
constexpr int kSlabD = kSlabRows * kNB;  // doubles per slab (8 KB)
constexpr int kSlabsPerTile = kNB / kSlabRows;
constexpr int kGroup = 4;                // column tiles per panel CTA
constexpr int kStages = 4;
constexpr int kStageD = (1 + kGroup) * kSlabD;   // B slab + kGroup A slabs (40 KB)
constexpr int kPanelThreads = 288;       // 8 consumer warps + 1 producer warp

struct BigParams {
  int p, q;
  int T;        // tile rows / columns of the Gram part: ceil(q / 64)
  int TX;       // tile columns of the X^T part: ceil(p / 64)
  int W;        // T + TX
  int halves;   // 2 with antithetic pairs
  int j;        // block row of this launch
  int nevals;   // evaluations in this batch
  int64_t s0;   // first sample of the batch
  size_t evalD; // doubles per evaluation: (T * W + T) tiles
  double *ws;
  const double *Gh;    // (q x q), leading dimension q
  const double *Rte;   // column-major p x p (columns scaled like the train columns)
  const double *cte;
  const int32_t *perms;
  double inv_ynsq;
  double *out;
  int *status;
};

__device__ __forceinline__ double *tile_ptr(const BigParams &a, int ev, int r, int c) {
  return a.ws + (size_t)ev * a.evalD + ((size_t)r * a.W + c) * kTileD;
}
__device__ __forceinline__ double *dinv_ptr(const BigParams &a, int ev, int r) {
  return a.ws + (size_t)ev * a.evalD + ((size_t)a.T * a.W + r) * kTileD;
}

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

// ---- mbarrier / TMA bulk copy (PTX ISA 8.x, sm_90+)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- gather
// CTA (tile row r, evaluation ev): the 64 rows of M in fragment-major tiles.  Eight source rows at a
// time are staged in shared memory with coalesced reads (a row of Gh, then a column of R_te), and every
// warp writes whole micro-tiles (512 contiguous bytes).
__global__ void __launch_bounds__(256) big_gather_kernel(BigParams a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int p = a.p, q = a.q, T = a.T, TX = a.TX;
  const int r = blockIdx.x, ev = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = lane >> 2, qq = lane & 3;
  const int rowlen = q + (q & 1);
  double *rows = reinterpret_cast<double *>(smem_raw);          // 8 x rowlen
  int *perm_s = reinterpret_cast<int *>(rows + (size_t)8 * rowlen);   // 64 T entries: pi^ or -1 (padding)
  const int64_t sample = a.s0 + ev / a.halves;
  const bool rev = (ev % a.halves) == 1;
  for (int i = tid; i < kNB * T; i += 256)
    perm_s[i] = (i < p) ? a.perms[sample * p + (rev ? p - 1 - i : i)] : (i == p ? p : -1);
  __syncthreads();
  for (int mr = 0; mr < 8; ++mr) {
    const int i0 = kNB * r + 8 * mr;
    // ---- Gram part: rows Gh[pi^_i, :]
    {
      const int src = perm_s[i0 + warp];
      double *dst = rows + (size_t)warp * rowlen;
      if (src >= 0) {
        const double *g = a.Gh + (size_t)src * q;
        for (int k = lane; k < q; k += 32) dst[k] = __ldg(g + k);
      }
    }
    __syncthreads();
    {
      const int s0r = perm_s[i0 + 2 * qq], s1r = perm_s[i0 + 2 * qq + 1];
      const double *row0 = rows + (size_t)(2 * qq) * rowlen, *row1 = row0 + rowlen;
      const int nmicro = (T - r) * 8;
      for (int m = warp; m < nmicro; m += 8) {
        const int ct = r + m / 8, nn = m % 8;
        const int col = kNB * ct + 8 * nn + c;
        const int pc = perm_s[col];
        double2 v;
        v.x = (pc >= 0 && s0r >= 0) ? row0[pc] : ((i0 + 2 * qq == col) ? 1.0 : 0.0);
        v.y = (pc >= 0 && s1r >= 0) ? row1[pc] : ((i0 + 2 * qq + 1 == col) ? 1.0 : 0.0);
        *reinterpret_cast<double2 *>(tile_ptr(a, ev, r, ct) + (size_t)(mr * 8 + nn) * 64 + 2 * lane) = v;
      }
    }
    __syncthreads();
    // ---- X^T part: row i = column pi_i of R_te (rows >= p, incl. the target row, are zero)
    {
      const int i = i0 + warp;
      double *dst = rows + (size_t)warp * rowlen;
      if (i < p) {
        const double *g = a.Rte + (size_t)perm_s[i] * p;
        for (int k = lane; k < p; k += 32) dst[k] = __ldg(g + k);
      }
    }
    __syncthreads();
    {
      const bool v0 = i0 + 2 * qq < p, v1 = i0 + 2 * qq + 1 < p;
      const double *row0 = rows + (size_t)(2 * qq) * rowlen, *row1 = row0 + rowlen;
      for (int m = warp; m < TX * 8; m += 8) {
        const int ct = m / 8, nn = m % 8;
        const int col = kNB * ct + 8 * nn + c;
        double2 v;
        v.x = (v0 && col < p) ? row0[col] : 0.0;
        v.y = (v1 && col < p) ? row1[col] : 0.0;
        *reinterpret_cast<double2 *>(tile_ptr(a, ev, r, T + ct) + (size_t)(mr * 8 + nn) * 64 + 2 * lane) = v;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- diagonal tile of block row j
// CTA = evaluation.  C_jj = M_jj - sum_{k<j} R_kj^T R_kj (DMMA, operands straight from L2), then the
// 64 x 64 Cholesky factor U (U^T U = C_jj) and U^-1 in shared memory.  U -> tile (j, j), U^-1 -> the
// evaluation's Dinv tile j.  A non-positive pivot zeroes its row (only the padding / target rows may
// legitimately do that); one in a feature row raises the status flag.
__global__ void __launch_bounds__(256) big_diag_kernel(BigParams a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double (*S)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(smem_raw);
  double (*V)[kNB + 1] = S + kNB;
  double *urow = reinterpret_cast<double *>(V + kNB);
  const int ev = blockIdx.x, j = a.j;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = lane >> 2, qq = lane & 3;
  const int mi = warp;
  double2 acc[8];
  {
    const double *t = tile_ptr(a, ev, j, j);
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) acc[ni] = *reinterpret_cast<const double2 *>(t + (size_t)(ni * 8 + mi) * 64 + 2 * lane);
  }
  for (int k = 0; k < j; ++k) {
    const double *t = tile_ptr(a, ev, k, j);
#pragma unroll 4
    for (int kk = 0; kk < 8; ++kk) {
      const double2 av = *reinterpret_cast<const double2 *>(t + (size_t)(kk * 8 + mi) * 64 + 2 * lane);
      double2 bv[8];
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) bv[ni] = *reinterpret_cast<const double2 *>(t + (size_t)(kk * 8 + ni) * 64 + 2 * lane);
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) {
        dmma(acc[ni].x, acc[ni].y, -av.x, bv[ni].x);
        dmma(acc[ni].x, acc[ni].y, -av.y, bv[ni].y);
      }
    }
  }
  // accumulator (m = 8 mi + c, n = 8 ni + 2 qq + e) = C_jj[n][m] = C_jj[m][n]
#pragma unroll
  for (int ni = 0; ni < 8; ++ni) {
    S[8 * mi + c][8 * ni + 2 * qq] = acc[ni].x;
    S[8 * mi + c][8 * ni + 2 * qq + 1] = acc[ni].y;
  }
  for (int e = tid; e < kNB * kNB; e += 256) V[e / kNB][e % kNB] = 0.0;
  __syncthreads();
  // right-looking Cholesky on the upper triangle with the tile in REGISTERS: thread (ty = warp, tx = lane)
  // owns the entries (i, j) = (ty + 8 ii, tx + 32 jj).  Row k belongs to warp k % 8, which gets the pivot
  // by one shuffle, scales its row and publishes it through the double-buffered urow: one barrier per
  // column.  (The column loop is split into blocks of eight so that every register index is static.)
  {
    double ar[8][2];
#pragma unroll
    for (int ii = 0; ii < 8; ++ii)
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) ar[ii][jj] = S[warp + 8 * ii][lane + 32 * jj];
    double *ub = urow;            // 2 x kNB
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const int k = 8 * kb + kk;
        double *uc = ub + (k & 1) * kNB;
        if (warp == kk) {
          const double d = __shfl_sync(kFull, ar[kb][k >> 5], k & 31);
          const bool ok = d > 1e-280;
          const double ri = ok ? rsqrt(d) : 0.0;
          if (!ok && lane == 0 && kNB * j + k < a.p) atomicExch(a.status, 1);
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int n = lane + 32 * jj;
            if (n >= k) {
              const double v = (n == k) ? (ok ? d * ri : 0.0) : ar[kb][jj] * ri;
              ar[kb][jj] = v;
              uc[n] = v;
            }
          }
        }
        __syncthreads();
#pragma unroll
        for (int ii = kb; ii < 8; ++ii) {
          const int m = warp + 8 * ii;
          if (m > k) {
            const double um = uc[m];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int n = lane + 32 * jj;
              if (n >= m) ar[ii][jj] = fma(-um, uc[n], ar[ii][jj]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int ii = 0; ii < 8; ++ii)
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) S[warp + 8 * ii][lane + 32 * jj] = ar[ii][jj];
  }
  __syncthreads();
  // inverse of the upper-triangular U: column n by back substitution, four lanes per column (they split
  // the dot product of each step and combine with two shuffles; a column only ever touches its own
  // entries of V, so warp-level synchronisation suffices).  Zero pivots give zero rows / columns.
  {
    const int n = tid >> 2, part = tid & 3;
    for (int i = kNB - 1; i >= 0; --i) {          // uniform trip count: columns n < i idle through the step
      double s = 0.0;
      if (i <= n)
        for (int k = i + 1 + part; k <= n; k += 4) s = fma(-S[i][k], V[k][n], s);
      s += __shfl_xor_sync(kFull, s, 1);
      s += __shfl_xor_sync(kFull, s, 2);
      if (i <= n && part == 0) {
        const double d = S[i][i];
        V[i][n] = (d != 0.0) ? (s + ((i == n) ? 1.0 : 0.0)) / d : 0.0;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // store both as fragment-major tiles (rows = k, columns = n; lower parts zero)
  double *tu = tile_ptr(a, ev, j, j), *td = dinv_ptr(a, ev, j);
  for (int e = tid; e < kNB * kNB / 2; e += 256) {
    const int micro = e / 32, l = e % 32;
    const int kk = micro / 8, nn = micro % 8;
    const int cc = l >> 2, q2 = l & 3;
    const int row = 8 * kk + 2 * q2, col = 8 * nn + cc;
    double2 u, v;
    u.x = (row <= col) ? S[row][col] : 0.0;
    u.y = (row + 1 <= col) ? S[row + 1][col] : 0.0;
    v.x = (row <= col) ? V[row][col] : 0.0;
    v.y = (row + 1 <= col) ? V[row + 1][col] : 0.0;
    *reinterpret_cast<double2 *>(tu + (size_t)micro * 64 + 2 * l) = u;
    *reinterpret_cast<double2 *>(td + (size_t)micro * 64 + 2 * l) = v;
  }
}

// ---------------------------------------------------------------- the other tiles of block row j
// CTA (group g of kGroup column tiles, evaluation).  Consumer warp w: column tile w / 2 of the group,
// micro-rows 4 (w & 1) .. +3 of the TRANSPOSED tile Ct = C_jc^T (m = column within tile c, n = row
// within block row j), all eight micro-columns: 32 accumulator tiles.  The producer warp streams, per
// 16 rows of k, the slab of R_kj (B operand, shared by the group) and the slabs of R_kc (A operands).
// Epilogue: R_jc^T = Ct U^-1 (C -> A register reuse), stored straight into tile (j, c).
__global__ void __launch_bounds__(kPanelThreads, 1) big_panel_kernel(BigParams a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *stage = reinterpret_cast<double *>(smem_raw);            // kStages x kStageD
  double *dinv_s = stage + (size_t)kStages * kStageD;              // one tile
  uint64_t *full = reinterpret_cast<uint64_t *>(dinv_s + kTileD);  // kStages
  uint64_t *empty = full + kStages;                                 // kStages
  uint64_t *dbar = empty + kStages;
  const int ev = blockIdx.y, j = a.j;
  const int c0 = j + 1 + kGroup * blockIdx.x;
  const int ng = (a.W - c0 < kGroup) ? a.W - c0 : kGroup;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2 * ng);
    }
    mbar_init(dbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nsteps = kSlabsPerTile * j;
  if (warp == 8) {
    if (lane == 0) {
      mbar_expect_tx(dbar, kTileD * sizeof(double));
      bulk_g2s(dinv_s, dinv_ptr(a, ev, j), kTileD * sizeof(double), dbar);
      for (int step = 0; step < nsteps; ++step) {
        const int s = step % kStages;
        if (step >= kStages) mbar_wait(&empty[s], ((step / kStages) - 1) & 1);
        const int k = step / kSlabsPerTile, sl = step % kSlabsPerTile;
        double *dst = stage + (size_t)s * kStageD;
        mbar_expect_tx(&full[s], (uint32_t)((1 + ng) * kSlabD * sizeof(double)));
        bulk_g2s(dst, tile_ptr(a, ev, k, j) + (size_t)sl * kSlabD, kSlabD * sizeof(double), &full[s]);
        for (int t = 0; t < ng; ++t)
          bulk_g2s(dst + (size_t)(1 + t) * kSlabD, tile_ptr(a, ev, k, c0 + t) + (size_t)sl * kSlabD,
                   kSlabD * sizeof(double), &full[s]);
      }
    }
    return;
  }
  const int ct = warp >> 1, mh = warp & 1;
  if (ct >= ng) return;
  double *tile = tile_ptr(a, ev, j, c0 + ct);
  double2 acc[4][8];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 8; ++ni)
      acc[mi][ni] = *reinterpret_cast<const double2 *>(tile + (size_t)(ni * 8 + 4 * mh + mi) * 64 + 2 * lane);
  for (int step = 0; step < nsteps; ++step) {
    const int s = step % kStages;
    mbar_wait(&full[s], (step / kStages) & 1);
    const double *Bs = stage + (size_t)s * kStageD;
    const double *As = Bs + (size_t)(1 + ct) * kSlabD;
#pragma unroll
    for (int kk = 0; kk < kSlabRows / 8; ++kk) {
      double2 av[4], bv[8];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        av[mi] = *reinterpret_cast<const double2 *>(As + (size_t)(kk * 8 + 4 * mh + mi) * 64 + 2 * lane);
        av[mi].x = -av[mi].x;
        av[mi].y = -av[mi].y;
      }
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) bv[ni] = *reinterpret_cast<const double2 *>(Bs + (size_t)(kk * 8 + ni) * 64 + 2 * lane);
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
          dmma(acc[mi][ni].x, acc[mi][ni].y, av[mi].x, bv[ni].x);
          dmma(acc[mi][ni].x, acc[mi][ni].y, av[mi].y, bv[ni].y);
        }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
  // epilogue: Out(mi, ni) = sum_{ki <= ni} Ct(mi, ki) Dinv(ki, ni), descending ni (Ct(mi, ni) is dead afterwards)
  mbar_wait(dbar, 0);
#pragma unroll
  for (int ni = 7; ni >= 0; --ni) {
    double2 dv[8];
#pragma unroll
    for (int ki = 0; ki < 8; ++ki)
      if (ki <= ni) dv[ki] = *reinterpret_cast<const double2 *>(dinv_s + (size_t)(ki * 8 + ni) * 64 + 2 * lane);
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      double o0 = 0.0, o1 = 0.0, r0 = 0.0, r1 = 0.0;   // two chains
#pragma unroll
      for (int ki = 0; ki < 8; ++ki)
        if (ki <= ni) {
          dmma(o0, o1, acc[mi][ki].x, dv[ki].x);
          dmma(r0, r1, acc[mi][ki].y, dv[ki].y);
        }
      *reinterpret_cast<double2 *>(tile + (size_t)(ni * 8 + 4 * mh + mi) * 64 + 2 * lane) = make_double2(o0 + r0, o1 + r1);
    }
  }
}

// ---------------------------------------------------------------- residual recurrence, lifts, scatter
// CTA = sample (both halves of an antithetic pair).  Thread <-> columns i of V (test-factor rows); the
// recurrence over the feature positions j is sequential per column, the squared norms per position
// are reduced across the CTA in chunks of 32 positions (fixed order: deterministic).
constexpr int kCostThreads = 1024;
constexpr int kCostMaxU = 2;   // columns per thread: p <= 2048

__global__ void __launch_bounds__(kCostThreads) big_cost_kernel(BigParams a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int p = a.p, T = a.T;
  double *cvec = reinterpret_cast<double *>(smem_raw);   // p
  double *cost = cvec + p;                               // p + 1
  double *accf = cost + p + 1;                           // p
  double *part = accf + p;                               // 32 x 32
  int *perm_s = reinterpret_cast<int *>(part + 32 * 32); // p
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sl = blockIdx.x;
  const int64_t sample = a.s0 + sl;
  const double weight = a.halves == 2 ? 0.5 : 1.0;
  const int pt = p / kNB, pc = p % kNB;   // tile column / column within it of the target column of R
  for (int h = 0; h < a.halves; ++h) {
    const int ev = sl * a.halves + h;
    __syncthreads();
    for (int k = tid; k < p; k += kCostThreads) {
      perm_s[k] = a.perms[sample * p + (h == 0 ? k : p - 1 - k)];
      const int jr = k % kNB;
      cvec[k] = tile_ptr(a, ev, k / kNB, pt)[(size_t)((jr / 8) * 8 + pc / 8) * 64 + (pc % 8) * 8 + (jr % 8)];
    }
    double res[kCostMaxU];
    const double *vbase[kCostMaxU];
    double s0 = 0.0;
#pragma unroll
    for (int u = 0; u < kCostMaxU; ++u) {
      const int i = tid + kCostThreads * u;
      res[u] = (i < p) ? a.cte[i] : 0.0;
      s0 = fma(res[u], res[u], s0);
      const int ic = (i < p) ? i : 0;
      // V[jrow][i]: tile (jrow / 64, T + i / 64), micro ((jrow % 64) / 8, (i % 64) / 8), offset (i % 8) * 8 + jrow % 8
      vbase[u] = tile_ptr(a, ev, 0, T + ic / kNB) + (size_t)((ic % kNB) / 8) * 64 + (ic % 8) * 8;
    }
    s0 = warp_sum(s0);
    if (lane == 0) part[warp * 32] = s0;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < 32; ++w) t += part[w * 32];
      cost[0] = t;
    }
    __syncthreads();
    const size_t row_tile_stride = (size_t)a.W * kTileD;
    for (int j0 = 0; j0 < p; j0 += 32) {
#pragma unroll 4
      for (int jj = 0; jj < 32; ++jj) {
        const int jrow = j0 + jj;
        double s = 0.0;
        if (jrow < p) {
          const double cj = cvec[jrow];
          const size_t off = (size_t)(jrow / kNB) * row_tile_stride + (size_t)((jrow % kNB) / 8) * 512 + (jrow % 8);
#pragma unroll
          for (int u = 0; u < kCostMaxU; ++u) {
            if (tid + kCostThreads * u < p) {
              res[u] = fma(-cj, vbase[u][off], res[u]);
              s = fma(res[u], res[u], s);
            }
          }
        }
        s = warp_sum(s);
        if (lane == 0) part[warp * 32 + jj] = s;
      }
      __syncthreads();
      if (tid < 32 && j0 + tid < p) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 32; ++w) t += part[w * 32 + tid];
        cost[j0 + tid + 1] = t;
      }
      __syncthreads();
    }
    for (int k = tid; k < p; k += kCostThreads) {
      const double lift = (cost[k] - cost[k + 1]) * a.inv_ynsq;
      const int f = perm_s[k];
      accf[f] = (h == 0 ? 0.0 : accf[f]) + weight * lift;
    }
  }
  __syncthreads();
  for (int f = tid; f < p; f += kCostThreads) a.out[sample * p + f] = accf[f];
}

// ---------------------------------------------------------------- condition bound for wide problems
// Column j of R'^-1 (R' = R D^-1, upper triangular) by column-oriented back substitution: CTA = column,
// x in shared memory, column i of R' is contiguous in the column-major R.  Writes the column to
// Xinv (column-major, ld p) and its sum of squares / of absolute values to colstat[2][p].
__global__ void __launch_bounds__(256) big_inv_col_kernel(int p, const double *__restrict__ R, const double *__restrict__ D,
                                                          double *__restrict__ Xinv, double *__restrict__ colstat) {
  extern __shared__ double x[];
  __shared__ double red[2][8];
  const int j = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < p; i += 256) x[i] = (i == j) ? 1.0 : 0.0;
  __syncthreads();
  for (int i = j; i >= 0; --i) {
    const double *col = R + (size_t)i * p;      // R[k][i], k <= i
    const double di = D[i];
    const double rii = col[i] / di;
    const double xi = x[i] / rii;               // inf / nan on a singular factor: propagates to the bound
    __syncthreads();
    if (tid == 0) x[i] = xi;
    for (int k = tid; k < i; k += 256) x[k] = fma(-(col[k] / di), xi, x[k]);
    __syncthreads();
  }
  double s2 = 0.0, s1 = 0.0;
  for (int i = tid; i < p; i += 256) {
    const double v = (i <= j) ? x[i] : 0.0;
    Xinv[(size_t)j * p + i] = v;
    s2 = fma(v, v, s2);
    s1 += fabs(v);
  }
  s2 = warp_sum(s2);
  s1 = warp_sum(s1);
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = s2;
    red[1][tid >> 5] = s1;
  }
  __syncthreads();
  if (tid == 0) {
    double a2 = 0.0, a1 = 0.0;
    for (int w = 0; w < 8; ++w) {
      a2 += red[0][w];
      a1 += red[1][w];
    }
    colstat[j] = a2;
    colstat[p + j] = a1;
  }
}

// Largest singular value of the upper-triangular A (column-major, ld p; column j divided by scale[j] when
// scale != nullptr) by power iteration on A^T A, one CTA of 1024 threads.  x, xs, y: p doubles each.
__device__ double power_sigma(int p, const double *__restrict__ A, const double *__restrict__ scale, double *x,
                              double *xs, double *y, double *red, int iters) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto block_norm = [&](const double *v) {
    double s = 0.0;
    for (int i = tid; i < p; i += 1024) s = fma(v[i], v[i], s);
    s = warp_sum(s);
    __syncthreads();
    if (lane == 0) red[warp] = s;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t += red[w];
    return sqrt(t);
  };
  for (int i = tid; i < p; i += 1024) x[i] = 1.0 + 0.5 * sin(1.0 * i);
  __syncthreads();
  double sigma = 0.0;
  for (int it = 0; it < iters; ++it) {
    const double nx = block_norm(x);
    for (int j = tid; j < p; j += 1024) xs[j] = x[j] / nx / (scale ? scale[j] : 1.0);
    __syncthreads();
    for (int i = tid; i < p; i += 1024) {          // y = A x: row i, columns j >= i (coalesced over i)
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int j = i;
#pragma unroll 2
      for (; j + 3 < p; j += 4) {
        s0 = fma(A[(size_t)j * p + i], xs[j], s0);
        s1 = fma(A[(size_t)(j + 1) * p + i], xs[j + 1], s1);
        s2 = fma(A[(size_t)(j + 2) * p + i], xs[j + 2], s2);
        s3 = fma(A[(size_t)(j + 3) * p + i], xs[j + 3], s3);
      }
      for (; j < p; ++j) s0 = fma(A[(size_t)j * p + i], xs[j], s0);
      y[i] = (s0 + s1) + (s2 + s3);
    }
    __syncthreads();
    sigma = block_norm(y);
    for (int j = warp; j < p; j += 32) {           // x = A^T y: column j, rows i <= j (contiguous)
      const double *col = A + (size_t)j * p;
      double s0 = 0.0;
      for (int i = lane; i <= j; i += 32) s0 = fma(col[i], y[i], s0);
      s0 = warp_sum(s0);
      if (lane == 0) x[j] = s0 / (scale ? scale[j] : 1.0);
    }
    __syncthreads();
  }
  return sigma;
}

// info[2] = |R'|_F |R'^-1|_F and info[3] = sqrt(max_i sum_j |Gh_ij|) sqrt(|R'^-1|_1 |R'^-1|_inf) are rigorous
// bounds >= cond_2(R') but over-state it by sqrt(p)-like factors at these widths (24x at p = 1000 on the
// benchmark data), so info[4] = sigma_max(R') sigma_max(R'^-1) by 24 power iterations each (converges from
// below; within 4 % here) is used with a safety factor: info[0] = min(info[2], info[3], 1.25 info[4]).
// info[1] = min|R'_kk| / max|R'_kk|.
__global__ void __launch_bounds__(1024) big_cond_kernel(int p, const double *__restrict__ R, const double *__restrict__ D,
                                                        const double *__restrict__ Gh, const double *__restrict__ Xinv,
                                                        const double *__restrict__ colstat, double *__restrict__ info) {
  extern __shared__ double pw[];   // 3 p
  __shared__ double red[6][32];
  const double s_r = power_sigma(p, R, D, pw, pw + p, pw + 2 * p, &red[0][0], 24);
  const double s_i = power_sigma(p, Xinv, nullptr, pw, pw + p, pw + 2 * p, &red[0][0], 24);
  __syncthreads();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double fr = 0.0, fi = 0.0, cm = 0.0, rm = 0.0, gs = 0.0, dmin = 1e300, dmax = 0.0;
  for (int i = tid; i < p; i += 1024) {
    double rs = 0.0, g = 0.0;
    for (int jc = i; jc < p; ++jc) rs += fabs(Xinv[(size_t)jc * p + i]);   // row i of the inverse
    for (int jc = 0; jc < p; ++jc) g += fabs(Gh[(size_t)i * (p + 1) + jc]);
    rm = fmax(rm, rs);
    gs = fmax(gs, g);
    fr += Gh[(size_t)i * (p + 1) + i];          // |R'|_F^2 = trace of the equilibrated Gram matrix
    fi += colstat[i];
    cm = fmax(cm, colstat[p + i]);
    const double d = fabs(R[(size_t)i * p + i] / D[i]);
    dmin = fmin(dmin, d);
    dmax = fmax(dmax, d);
  }
  fr = warp_sum(fr);
  fi = warp_sum(fi);
  for (int o = 16; o > 0; o >>= 1) {
    cm = fmax(cm, __shfl_xor_sync(kFull, cm, o));
    rm = fmax(rm, __shfl_xor_sync(kFull, rm, o));
    gs = fmax(gs, __shfl_xor_sync(kFull, gs, o));
    dmin = fmin(dmin, __shfl_xor_sync(kFull, dmin, o));
    dmax = fmax(dmax, __shfl_xor_sync(kFull, dmax, o));
  }
  if (lane == 0) {
    red[0][warp] = fr;
    red[1][warp] = fi;
    red[2][warp] = cm;
    red[3][warp] = rm;
    red[4][warp] = gs;
    red[5][warp] = dmin;
  }
  __shared__ double rmax[32];
  if (lane == 0) rmax[warp] = dmax;
  __syncthreads();
  if (tid == 0) {
    double A = 0.0, B = 0.0, C = 0.0, Rm = 0.0, G = 0.0, mn = 1e300, mx = 0.0;
    for (int w = 0; w < 32; ++w) {
      A += red[0][w];
      B += red[1][w];
      C = fmax(C, red[2][w]);
      Rm = fmax(Rm, red[3][w]);
      G = fmax(G, red[4][w]);
      mn = fmin(mn, red[5][w]);
      mx = fmax(mx, rmax[w]);
    }
    double frob = sqrt(A) * sqrt(B), sharp = sqrt(G) * sqrt(C * Rm);
    if (!(frob == frob)) frob = INFINITY;
    if (!(sharp == sharp)) sharp = INFINITY;
    double power = s_r * s_i;
    if (!(power == power)) power = INFINITY;
    info[0] = fmin(fmin(frob, sharp), 1.25 * power);
    info[1] = (mx > 0.0) ? mn / mx : 0.0;
    info[2] = frob;
    info[3] = sharp;
    info[4] = power;
  }
}

// ---------------------------------------------------------------- dense Gram matrix <-> tiles (reduction of wide problems)
// tiles (r <= c) of the symmetric matrix scale * G + reg * diag(1..1, 0) (G: q x q row-major, upper part
// valid; reg on the first p diagonal entries: the sqrt(reg) I rows of reference :310), identity padding.
__global__ void __launch_bounds__(256) dense_to_tiles_kernel(BigParams a, const double *G, double scale, double reg) {
  const int r = blockIdx.x, ct = blockIdx.y;
  if (ct < r) return;
  const int q = a.q, p = a.p;
  double *t = tile_ptr(a, 0, r, ct);
  for (int e = threadIdx.x; e < kTileD / 2; e += 256) {
    const int micro = e / 32, l = e % 32;
    const int row0 = kNB * r + 8 * (micro / 8) + 2 * (l & 3), col = kNB * ct + 8 * (micro % 8) + (l >> 2);
    double v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = row0 + h;
      if (row < q && col < q) {
        const double gv = (row <= col) ? G[(size_t)row * q + col] : G[(size_t)col * q + row];
        v[h] = scale * gv + ((row == col && row < p) ? reg : 0.0);
      } else {
        v[h] = (row == col) ? 1.0 : 0.0;
      }
    }
    *reinterpret_cast<double2 *>(t + (size_t)micro * 64 + 2 * l) = make_double2(v[0], v[1]);
  }
}

// slot layout of reduce.cu: q x q row-major upper-triangular factor, then [q*q] = sum of squares of the y column
__global__ void tiles_to_slot_kernel(BigParams a, const double *G, double scale, double *slot) {
  const int q = a.q;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < q * q) {
    const int i = e / q, j = e - i * q;
    double v = 0.0;
    if (i <= j) {
      const double *t = tile_ptr(a, 0, i / kNB, j / kNB);
      const int ir = i % kNB, jc = j % kNB;
      v = t[(size_t)((ir / 8) * 8 + jc / 8) * 64 + (jc % 8) * 8 + (ir % 8)];
    }
    slot[e] = v;
  }
  if (e < 8) slot[(size_t)q * q + e] = (e == 0) ? scale * G[(size_t)(q - 1) * q + (q - 1)] : 0.0;
}

}  // namespace

bool lifts_big_supported(int p) { return p > 152 && p <= kCostThreads * kCostMaxU - 1; }

int lifts_big_cond(int p, const double *R_tr_cm, const double *D, const double *Gh, double *Xinv, double *colstat,
                   double *info, cudaStream_t st) {
  big_inv_col_kernel<<<p, 256, (size_t)p * sizeof(double), st>>>(p, R_tr_cm, D, Xinv, colstat);
  LSSPA_LAUNCH_CHECK();
  big_cond_kernel<<<1, 1024, (size_t)3 * p * sizeof(double), st>>>(p, R_tr_cm, D, Gh, Xinv, colstat, info);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

static void big_geometry(int p, int antithetical, BigParams &a) {
  a.p = p;
  a.q = p + 1;
  a.T = (a.q + kNB - 1) / kNB;
  a.TX = (p + kNB - 1) / kNB;
  a.W = a.T + a.TX;
  a.halves = antithetical ? 2 : 1;
  a.evalD = ((size_t)a.T * a.W + a.T) * kTileD;
}

}  // namespace lsspa

using namespace lsspa;

extern "C" size_t lsspa_gram_big_factor_workspace_bytes(int p) {
  if (p < 1 || p > 2047) return 0;
  const int T = (p + 1 + kNB - 1) / kNB;
  return ((size_t)T * T + T) * kTileD * sizeof(double);
}

// slot = Cholesky factor (slot layout) of scale * G_acc + reg * diag(I_p, 0); status_flag raised on a bad feature pivot
extern "C" int lsspa_gram_big_factor(const double *G_acc, int p, double scale, double reg, double *slot_out,
                                     void *workspace, size_t workspace_bytes, int *status_flag, void *stream) {
  if (!G_acc || !slot_out || !status_flag || p < 1 || p > 2047) return LSSPA_E_BADARG;
  if (!workspace || workspace_bytes < lsspa_gram_big_factor_workspace_bytes(p)) return LSSPA_E_WORKSPACE;
  BigParams a;
  a.p = p;
  a.q = p + 1;
  a.T = (a.q + kNB - 1) / kNB;
  a.TX = 0;
  a.W = a.T;
  a.halves = 1;
  a.evalD = ((size_t)a.T * a.W + a.T) * kTileD;
  a.ws = reinterpret_cast<double *>(workspace);
  a.Gh = nullptr;
  a.Rte = nullptr;
  a.cte = nullptr;
  a.perms = nullptr;
  a.inv_ynsq = 0.0;
  a.out = nullptr;
  a.status = status_flag;
  a.s0 = 0;
  a.nevals = 1;
  cudaStream_t st = as_stream(stream);
  const size_t panel_smem = ((size_t)kStages * kStageD + kTileD) * sizeof(double) + (2 * kStages + 1) * sizeof(uint64_t);
  const size_t diag_smem = ((size_t)2 * kNB * (kNB + 1) + 2 * kNB) * sizeof(double);
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(big_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)diag_smem));
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(big_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem));
  dense_to_tiles_kernel<<<dim3((unsigned)a.T, (unsigned)a.T), 256, 0, st>>>(a, G_acc, scale, reg);
  LSSPA_LAUNCH_CHECK();
  for (int j = 0; j < a.T; ++j) {
    a.j = j;
    big_diag_kernel<<<1, 256, diag_smem, st>>>(a);
    LSSPA_LAUNCH_CHECK();
    const int ncols = a.W - j - 1;
    if (ncols > 0) {
      big_panel_kernel<<<dim3((unsigned)((ncols + kGroup - 1) / kGroup), 1), kPanelThreads, panel_smem, st>>>(a);
      LSSPA_LAUNCH_CHECK();
    }
  }
  tiles_to_slot_kernel<<<(unsigned)ceil_div((int64_t)a.q * a.q, 256), 256, 0, st>>>(a, G_acc, scale, slot_out);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

extern "C" int lsspa_lifts_big_supported(int p) { return lifts_big_supported(p) ? 1 : 0; }

extern "C" size_t lsspa_lifts_big_workspace_bytes(int p, int64_t count, int antithetical, size_t budget_bytes) {
  if (!lifts_big_supported(p) || count < 1) return 0;
  BigParams a;
  big_geometry(p, antithetical, a);
  const size_t per_sample = a.evalD * sizeof(double) * a.halves;
  size_t samples = budget_bytes / per_sample;
  if (samples < 1) samples = 1;
  if ((int64_t)samples > count) samples = (size_t)count;
  return samples * per_sample;
}

extern "C" int lsspa_lifts_big(int p, const double *gram, const double *R_te_cm, const double *c_te, double y_norm_sq,
                               const int32_t *perms, int64_t count, int antithetical, double *lifts_out,
                               void *workspace, size_t workspace_bytes, int *status_flag, void *stream) {
  if (!lifts_big_supported(p)) return LSSPA_E_UNSUPPORTED;
  if (!gram || !R_te_cm || !c_te || !lifts_out || !status_flag || !(y_norm_sq > 0.0) || count < 0) return LSSPA_E_BADARG;
  if (count == 0) return LSSPA_OK;
  if (!perms) return LSSPA_E_BADARG;
  BigParams a;
  big_geometry(p, antithetical, a);
  const size_t per_sample = a.evalD * sizeof(double) * a.halves;
  if (!workspace || workspace_bytes < per_sample) return LSSPA_E_WORKSPACE;
  int64_t batch = (int64_t)(workspace_bytes / per_sample);
  if (batch > count) batch = count;
  if (batch > 32767 / a.halves) batch = 32767 / a.halves;       // gridDim.y
  a.ws = reinterpret_cast<double *>(workspace);
  a.Gh = gram;
  a.Rte = R_te_cm;
  a.cte = c_te;
  a.perms = perms;
  a.inv_ynsq = 1.0 / y_norm_sq;
  a.out = lifts_out;
  a.status = status_flag;
  cudaStream_t st = as_stream(stream);
  const int rowlen = a.q + (a.q & 1);
  const size_t gather_smem = (size_t)8 * rowlen * sizeof(double) + (size_t)kNB * a.T * sizeof(int);
  const size_t panel_smem = ((size_t)kStages * kStageD + kTileD) * sizeof(double) + (2 * kStages + 1) * sizeof(uint64_t);
  const size_t cost_smem = ((size_t)3 * p + 1 + 32 * 32) * sizeof(double) + (size_t)p * sizeof(int);
  const size_t diag_smem = ((size_t)2 * kNB * (kNB + 1) + 2 * kNB) * sizeof(double);
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(big_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)diag_smem));
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(big_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gather_smem));
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(big_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem));
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(big_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cost_smem));
  for (int64_t s0 = 0; s0 < count; s0 += batch) {
    const int ns = (int)((count - s0 < batch) ? count - s0 : batch);
    a.s0 = s0;
    a.nevals = ns * a.halves;
    big_gather_kernel<<<dim3((unsigned)a.T, (unsigned)a.nevals), 256, gather_smem, st>>>(a);
    LSSPA_LAUNCH_CHECK();
    for (int j = 0; j < a.T; ++j) {
      a.j = j;
      big_diag_kernel<<<(unsigned)a.nevals, 256, diag_smem, st>>>(a);
      LSSPA_LAUNCH_CHECK();
      const int ncols = a.W - j - 1;
      if (ncols > 0) {
        big_panel_kernel<<<dim3((unsigned)((ncols + kGroup - 1) / kGroup), (unsigned)a.nevals), kPanelThreads, panel_smem, st>>>(a);
        LSSPA_LAUNCH_CHECK();
      }
    }
    big_cost_kernel<<<(unsigned)ns, kCostThreads, cost_smem, st>>>(a);
    LSSPA_LAUNCH_CHECK();
  }
  return LSSPA_OK;
}
