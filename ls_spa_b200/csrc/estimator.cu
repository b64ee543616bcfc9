// Online estimator of LS-SPA on the device.
//
// Replaces the bookkeeping of the reference's sample loop (ls_spa/ls_spa.py:186-236):
//   merge_sample_mean / merge_sample_cov (:103-119)  -> Chan merge of batch moments
//   error_estimates (:321-341)                        -> factor-free Gaussian draws
//   stop test (:229), error_history (:225)            -> device flag + history array
//   attribution_history (:217-219)                    -> prefix means
//   theta / r_squared epilogue (:240-243)             -> triangular solve + residual
//
// Work is organised per SUPER-BATCH (a run of consecutive batches): one launch folds all its
// batches into the state in order and emits the squared error draws of the batches this rank
// owns, one launch takes all their quantiles in parallel.  The host (engine.py) scans the
// per-batch errors for the first one below the tolerance and, if the stop falls inside the
// super-batch, restores the snapshot and replays the merges up to that batch -- the merges are
// deterministic, so this is exactly the reference's `break`.
//
// error_estimates draws 1024 vectors from N(0, unbiased_cov / n).  The lift covariance is
// exactly singular (every lift vector sums to the full-model R^2), so instead of
// factorising it we use z_s = sum_k g_ks (l_k - mean) / sqrt(n (n-1)), g iid N(0,1), which
// has exactly that covariance.  S_s = sum_k g_ks (l_k - mean) and G_s = sum_k g_ks are
// kept as running sums and re-centred whenever the mean moves (same algebra as the
// covariance merge), so nothing cancels catastrophically.

#include "common.cuh"

namespace lsspa {

constexpr int kDraws = LSSPA_ERR_DRAWS;  // 1024
// State (doubles): mean[2][p] (ping-pong), G[2][kDraws] (ping-pong), cov[p][p] (biased),
// S[p][kDraws] feature-major.  Row f of cov and column f of S are owned by CTA f of the absorb
// kernel and updated in place; mean and G are read by every CTA, hence double-buffered.
__host__ __device__ inline size_t state_doubles(int p) {
  return 2 * (size_t)p + 2 * (size_t)kDraws + (size_t)p * p + (size_t)p * kDraws;
}
struct StateView {
  double *mean[2];
  double *G[2];
  double *cov;
  double *S;
};
__host__ __device__ inline StateView view_state(double *base, int p) {
  StateView v;
  double *c = base;
  v.mean[0] = c; c += p;
  v.mean[1] = c; c += p;
  v.G[0] = c; c += kDraws;
  v.G[1] = c; c += kDraws;
  v.cov = c; c += (size_t)p * p;
  v.S = c;
  return v;
}

// partial moment block of one (batch, rank)
constexpr int kPartHdr = 8;
__host__ __device__ inline size_t partial_doubles(int p) {
  return kPartHdr + (size_t)p + (size_t)p * p + kDraws + (size_t)p * kDraws;
}
struct PartView {
  const double *hdr;   // [0] = n
  const double *mean;  // p
  const double *m2;    // p x p   sum (l-mean)(l-mean)^T
  const double *G;     // kDraws
  const double *S;     // [p][kDraws]  sum g (l-mean)
};
__host__ __device__ inline PartView view_part(const double *base, int p) {
  PartView v;
  v.hdr = base;
  v.mean = base + kPartHdr;
  v.m2 = v.mean + p;
  v.G = v.m2 + (size_t)p * p;
  v.S = v.G + kDraws;
  return v;
}

// ---------------------------------------------------------------- Gaussian stream
// counter-based: (seed, global sample index, pair index) -> two N(0,1) floats
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ void gauss_pair(uint64_t seed, uint64_t sample, uint32_t pair, float &g0,
                                           float &g1) {
  uint64_t z = mix64(seed + 0x9E3779B97F4A7C15ULL * (sample * (uint64_t)(kDraws / 2) + pair + 1));
  z = mix64(z + 0xD1B54A32D192ED03ULL);
  const float u1 = (float)((uint32_t)(z >> 40) + 1u) * 5.9604644775390625e-8f;  // (0,1]
  const float u2 = (float)((uint32_t)z & 0xFFFFFFu) * 5.9604644775390625e-8f;   // [0,1)
  // fast intrinsics (MUFU): ~1e-6 relative, far below the Monte-Carlo noise of 1024 draws, and a
  // third of the instructions of the IEEE-accurate routines (the generator feeds DMMA tiles)
  const float r = __fsqrt_rn(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  g0 = r * c;
  g1 = r * s;
}

// ---------------------------------------------------------------- per-batch partial moments
// batch_desc[b] = {first row in `lifts`, row count, global index of the first sample}
__global__ void part_mean_kernel(int p, const double *lifts, const int64_t *desc, double *partials,
                                 size_t pstride) {
  const int b = blockIdx.x;
  const int64_t r0 = desc[3 * b], cnt = desc[3 * b + 1];
  double *base = partials + (size_t)b * pstride;
  if (threadIdx.x < kPartHdr) base[threadIdx.x] = (threadIdx.x == 0) ? (double)cnt : 0.0;
  for (int f = threadIdx.x; f < p; f += blockDim.x) {
    double s = 0.0;
    for (int64_t k = 0; k < cnt; ++k) s += lifts[(r0 + k) * p + f];
    base[kPartHdr + f] = cnt > 0 ? s / (double)cnt : 0.0;
  }
}

__global__ void part_m2_kernel(int p, const double *lifts, const int64_t *desc, double *partials,
                               size_t pstride) {
  const int b = blockIdx.z;
  const int64_t r0 = desc[3 * b], cnt = desc[3 * b + 1];
  double *base = partials + (size_t)b * pstride;
  const double *mean = base + kPartHdr;
  const int fa = blockIdx.y * 16 + threadIdx.y, fb = blockIdx.x * 16 + threadIdx.x;
  if (fa >= p || fb >= p) return;
  const double ma = mean[fa], mb = mean[fb];
  double s = 0.0;
  for (int64_t k = 0; k < cnt; ++k) {
    const double *row = lifts + (r0 + k) * p;
    s = fma(row[fa] - ma, row[fb] - mb, s);
  }
  base[kPartHdr + p + (size_t)fa * p + fb] = s;
}

constexpr int kDrawTile = 16;  // draws per CTA
constexpr int kRowChunk = 32;  // rows whose Gaussians are staged at once

__global__ void __launch_bounds__(1024) part_s_kernel(int p, const double *lifts, const int64_t *desc, uint64_t seed,
                              double *partials, size_t pstride) {
  __shared__ float g[kRowChunk][kDrawTile];
  const int b = blockIdx.y;
  const int tile = blockIdx.x;  // draws [tile*16, tile*16+16)
  const int64_t r0 = desc[3 * b], cnt = desc[3 * b + 1], gidx0 = desc[3 * b + 2];
  double *base = partials + (size_t)b * pstride;
  const double *mean = base + kPartHdr;
  double *Gout = base + kPartHdr + p + (size_t)p * p;
  double *Sout = Gout + kDraws;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int nf = (p + nt - 1) / nt;  // features per thread (1 unless p > blockDim)
  double gsum = 0.0;                 // thread d < 16: sum of draw d over rows
  for (int fi = 0; fi < nf; ++fi) {
    const int f = tid + fi * nt;
    double acc[kDrawTile];
#pragma unroll
    for (int d = 0; d < kDrawTile; ++d) acc[d] = 0.0;
    const double mf = (f < p) ? mean[f] : 0.0;
    for (int64_t k0 = 0; k0 < cnt; k0 += kRowChunk) {
      const int rows = (int)((cnt - k0 < kRowChunk) ? cnt - k0 : kRowChunk);
      __syncthreads();
      for (int e = tid; e < kRowChunk * (kDrawTile / 2); e += nt) {
        const int r = e / (kDrawTile / 2), pr = e % (kDrawTile / 2);
        float a0 = 0.f, a1 = 0.f;
        if (r < rows)
          gauss_pair(seed, (uint64_t)(gidx0 + k0 + r), (uint32_t)(tile * (kDrawTile / 2) + pr), a0, a1);
        g[r][2 * pr] = a0;
        g[r][2 * pr + 1] = a1;
      }
      __syncthreads();
      if (fi == 0 && tid < kDrawTile)
        for (int r = 0; r < rows; ++r) gsum += (double)g[r][tid];
      if (f < p) {
        for (int r = 0; r < rows; ++r) {
          const double x = lifts[(r0 + k0 + r) * p + f] - mf;
#pragma unroll
          for (int d = 0; d < kDrawTile; ++d) acc[d] = fma((double)g[r][d], x, acc[d]);
        }
      }
    }
    if (f < p) {
#pragma unroll
      for (int d = 0; d < kDrawTile; ++d) Sout[(size_t)f * kDraws + tile * kDrawTile + d] = acc[d];
    }
  }
  if (tid < kDrawTile) Gout[tile * kDrawTile + tid] = gsum;
}

// ---------------------------------------------------------------- per-batch partial moments on DMMA (p <= 128)
// Both second-order partials are products over the rows of the batch:
//   M2 = Xc^T Xc   (p x p)      and      S = Xc^T g   (p x 1024 draws),     Xc = lifts - mean.
// A CTA stages up to 128 centered rows in shared memory (row-major, stride % 16 == 8, so that the
// 8 bytes a lane needs, Xc[4 ks + q][8 t + c], are conflict-free) -- that value is at once the A
// fragment of feature tile t (A[m=c][k=q]) and the B fragment of the same tile (B[k=q][n=c]).
constexpr int kMmaRows = 128;

__device__ __forceinline__ void dmma_e(double &d0, double &d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

// row pitch of the staged rows: the A / B fragments read element (row 4 ks + q, column 8 t + c), so within
// a half-warp (c = 0..3, q = 0..3) the bank pairs are distinct iff q * pitch covers 0, 4, 8, 12 (mod 16):
// pitch = 4 (mod 8).  (A pitch of 8 mod 16 makes q = 0 / 2 and 1 / 3 collide: two-way conflicts on the one
// shared-memory load per DMMA, which then saturates the LSU before the FP64 pipe.)
__host__ __device__ inline int mma_ldx(int ft) { return 8 * ft + 4; }

// rows [k0, k0 + rows) of the batch, centered, zero padded to kMmaRows x ldx.  A warp takes whole
// rows (lane -> features lane, lane + 32, ...): coalesced, no index arithmetic.  The loads of four rows
// (from clamped, always valid addresses) are issued into registers BEFORE the first shared-memory store:
// written as load - subtract - store per element, every load waited for its predecessor's store (one
// L2 round trip per element, 64 in a row per warp -- a third of part_s_mma_kernel's run time).
__device__ __forceinline__ void stage_rows(double *Xs, int ldx, const double *__restrict__ lifts,
                                           const double *__restrict__ mean, int p, int64_t row0, int rows, int tid,
                                           int nt) {
  const int lane = tid & 31, w = tid >> 5, nw = nt >> 5;
  double mu[4];
  int fc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int f = lane + 32 * j;
    fc[j] = f < p ? f : p - 1;
    mu[j] = mean[fc[j]];
  }
  constexpr int RB = 4;
  for (int rb = w; rb < kMmaRows; rb += RB * nw) {
    double v[RB][4];
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      const int r = rb + u * nw;
      const double *src = lifts + (row0 + (r < rows ? r : rows - 1)) * p;
#pragma unroll
      for (int j = 0; j < 4; ++j) v[u][j] = __ldg(src + fc[j]);
    }
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      const int r = rb + u * nw;
      if (r < kMmaRows) {
        double *dst = Xs + (size_t)r * ldx;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int f = lane + 32 * j;
          if (f < ldx) dst[f] = (r < rows && f < p) ? v[u][j] - mu[j] : 0.0;
        }
      }
    }
  }
}

// M2: warp g owns tile rows g and FT-1-g of the upper triangle (FT + 1 tiles, balanced)
template <int MAXFT>
__global__ void __launch_bounds__(256, 2) part_m2_mma_kernel(int p, const double *lifts, const int64_t *desc,
                                                          double *partials, size_t pstride) {
  extern __shared__ __align__(16) double Xs[];
  const int b = blockIdx.x;
  const int64_t r0 = desc[3 * b], cnt = desc[3 * b + 1];
  double *base = partials + (size_t)b * pstride;
  const double *mean = base + kPartHdr;
  double *m2 = base + kPartHdr + p;
  const int FT = (p + 7) / 8, ldx = mma_ldx(FT);
  const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
  const int c = lane >> 2, q = lane & 3;
  const int rowA = g, rowB = FT - 1 - g;           // tile rows of this warp (rowB skipped if <= rowA)
  const bool hasA = rowA < FT && rowA <= rowB, hasB = rowB > rowA && rowB < FT;
  double acc[MAXFT + 1][2];
#pragma unroll
  for (int i = 0; i <= MAXFT; ++i) acc[i][0] = acc[i][1] = 0.0;
  for (int64_t k0 = 0; k0 < cnt; k0 += kMmaRows) {
    const int rows = (int)((cnt - k0 < kMmaRows) ? cnt - k0 : kMmaRows);
    __syncthreads();
    stage_rows(Xs, ldx, lifts, mean, p, r0 + k0, rows, tid, blockDim.x);
    __syncthreads();
    const int ksteps = (rows + 3) / 4;
    for (int ks = 0; ks < ksteps; ++ks) {
      const double *xrow = Xs + (size_t)(4 * ks + q) * ldx + c;
      const double xa = hasA ? xrow[8 * rowA] : 0.0, xb = hasB ? xrow[8 * rowB] : 0.0;
      // accumulator i: i < FT - rowA -> tile (rowA, rowA + i); else tile (rowB, rowB + i - (FT - rowA))
      const int nA = FT - rowA;
#pragma unroll
      for (int i = 0; i <= MAXFT; ++i) {
        if (hasA && i < nA) {
          dmma_e(acc[i][0], acc[i][1], xa, xrow[8 * (rowA + i)]);
        } else if (hasB && i >= nA && i - nA < FT - rowB) {
          dmma_e(acc[i][0], acc[i][1], xb, xrow[8 * (rowB + i - nA)]);
        }
      }
    }
  }
  // C layout: lane (c, q) holds M2[8 t1 + c][8 t2 + 2q + e]; both triangles are written
#pragma unroll
  for (int i = 0; i <= MAXFT; ++i) {
    const int nA = FT - rowA;
    int t1 = -1, t2 = -1;
    if (hasA && i < nA) {
      t1 = rowA;
      t2 = rowA + i;
    } else if (hasB && i >= nA && i - nA < FT - rowB) {
      t1 = rowB;
      t2 = rowB + i - nA;
    }
    if (t1 >= 0) {
      const int fa = 8 * t1 + c;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int fb = 8 * t2 + 2 * q + e;
        if (fa < p && fb < p) {
          m2[(size_t)fa * p + fb] = acc[i][e];
          m2[(size_t)fb * p + fa] = acc[i][e];
        }
      }
    }
  }
}

// S and G: CTA = (block of kSDraws draws, batch).  A warp task = (all feature tiles, one tile of 8 draws).
// The Gaussians are generated in the lanes that need them, one generator call per lane for TWO k-steps:
// lane (c, q) draws the pair (draws 8 T + (c & ~1), + 1) of row 4 (ks + (c & 1)) + q, and one shuffle with
// lane ^ 4 gives every lane column c of both k-steps (even c: own g0 now, the partner's g0 next; odd c: the
// partner's g1 now, own g1 next) -- 2 x FT DMMAs per generator call.
constexpr int kSDraws = 128;

template <int MAXFT>
__global__ void __launch_bounds__(256, 2) part_s_mma_kernel(int p, const double *lifts, const int64_t *desc,
                                                         uint64_t seed, double *partials, size_t pstride) {
  extern __shared__ __align__(16) double Xs[];
  const int b = blockIdx.y, db = blockIdx.x;
  const int64_t r0 = desc[3 * b], cnt = desc[3 * b + 1], gidx0 = desc[3 * b + 2];
  double *base = partials + (size_t)b * pstride;
  const double *mean = base + kPartHdr;
  double *Gout = base + kPartHdr + p + (size_t)p * p;
  double *Sout = Gout + kDraws;
  const int FT = (p + 7) / 8, ldx = mma_ldx(FT);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int c = lane >> 2, q = lane & 3;
  const bool even = (c & 1) == 0;
  constexpr int kTiles = kSDraws / 8;               // draw tiles per CTA
  const int64_t nchunks = (cnt + kMmaRows - 1) / kMmaRows;
  for (int task = w; task < kTiles; task += 8) {
    const int T = db * kTiles + task;
    double acc[MAXFT][2];
#pragma unroll
    for (int i = 0; i < MAXFT; ++i) acc[i][0] = acc[i][1] = 0.0;
    double gs = 0.0;
    const uint32_t pr = (uint32_t)((T * 8 + (c & ~1)) >> 1);
    for (int64_t ch = 0; ch < nchunks; ++ch) {
      const int64_t k0 = ch * kMmaRows;
      const int rows = (int)((cnt - k0 < kMmaRows) ? cnt - k0 : kMmaRows);
      if (nchunks > 1 || task < 8) {   // one staging serves every task unless the batch needs several chunks
        __syncthreads();
        stage_rows(Xs, ldx, lifts, mean, p, r0 + k0, rows, tid, blockDim.x);
        __syncthreads();
      }
      const int ksteps = (rows + 3) / 4;
      for (int ks = 0; ks < ksteps; ks += 2) {
        const int rg = 4 * (ks + (even ? 0 : 1)) + q;
        float g0 = 0.f, g1 = 0.f;
        if (rg < rows) gauss_pair(seed, (uint64_t)(gidx0 + k0 + rg), pr, g0, g1);
        const float recv = __shfl_xor_sync(kFull, even ? g1 : g0, 4);
        const double b0 = (double)(even ? g0 : recv);     // k-step ks
        const double b1 = (double)(even ? recv : g1);     // k-step ks + 1 (rows beyond the batch: zeros)
        gs += b0 + b1;
        const double *x0 = Xs + (size_t)(4 * ks + q) * ldx + c;
#pragma unroll
        for (int i = 0; i < MAXFT; ++i)
          if (i < FT) dmma_e(acc[i][0], acc[i][1], x0[8 * i], b0);
        if (ks + 1 < ksteps) {
          const double *x1 = x0 + (size_t)4 * ldx;
#pragma unroll
          for (int i = 0; i < MAXFT; ++i)
            if (i < FT) dmma_e(acc[i][0], acc[i][1], x1[8 * i], b1);
        }
      }
    }
    // C layout: lane (c, q) holds S[feature 8 i + c][draw 8 T + 2q + e]
#pragma unroll
    for (int i = 0; i < MAXFT; ++i) {
      const int fa = 8 * i + c;
      if (i < FT && fa < p)
        *reinterpret_cast<double2 *>(Sout + (size_t)fa * kDraws + 8 * T + 2 * q) = make_double2(acc[i][0], acc[i][1]);
    }
    gs = quad_sum(gs);
    if (q == 0) Gout[8 * T + c] = gs;
  }
}

// ---------------------------------------------------------------- batch step
// Order statistics of 1024 non-negative doubles held by one warp (32 per lane), without sorting:
// the IEEE bit pattern of a non-negative double is monotone, so the k-th largest value is the
// largest key K with count(key >= K) >= k, built bit by bit from the top; one warp-wide integer
// reduction per bit, no barriers.
constexpr int kPerLane = kDraws / 32;

__device__ __forceinline__ unsigned long long kth_largest_key(const unsigned long long (&key)[kPerLane], int k) {
  unsigned long long K = 0;
  for (int bit = 62; bit >= 0; --bit) {
    const unsigned long long trial = K | (1ULL << bit);
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) cnt += (key[i] >= trial) ? 1 : 0;
    cnt = __reduce_add_sync(kFull, cnt);
    if (cnt >= k) K = trial;
  }
  return K;
}

// np.quantile(sqrt(v), 0.95) of the 1024 non-negative values a warp holds (32 per lane; valid on lane 0).
// A NaN value (an estimate after a single sample: 0 / sqrt(n (n-1)), the reference returns nan there too)
// makes the whole quantile NaN, so that it can never pass the `< tolerance` stop test.
__device__ __forceinline__ double quantile95_sqrt(const double (&v)[kPerLane], int lane) {
  bool nan_row = false;
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) nan_row |= !(v[i] == v[i]);
  nan_row = __any_sync(kFull, nan_row);
  unsigned long long key[kPerLane];
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) key[i] = (unsigned long long)__double_as_longlong(v[i] > 0.0 ? v[i] : 0.0);
  // linear interpolation (reference :340-341) between the ascending ranks lo and lo + 1, i.e. the
  // (kDraws - lo)-th and (kDraws - lo - 1)-th largest
  const double pos = (double)(kDraws - 1) * 0.95;
  const int lo = (int)floor(pos);
  const unsigned long long klo = kth_largest_key(key, kDraws - lo);
  // the next rank up is klo again when ties leave fewer than kDraws - lo - 1 keys above it,
  // otherwise the smallest key above klo
  int n_gt = 0;
  unsigned long long above = ~0ULL;
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) {
    if (key[i] > klo) {
      ++n_gt;
      above = key[i] < above ? key[i] : above;
    }
  }
  n_gt = __reduce_add_sync(kFull, n_gt);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(kFull, above, o);
    above = other < above ? other : above;
  }
  const unsigned long long khi = (n_gt >= kDraws - lo - 1) ? above : klo;
  const double t = pos - (double)lo;
  const double a = sqrt(__longlong_as_double((long long)klo)), c = sqrt(__longlong_as_double((long long)khi));
  return nan_row ? __longlong_as_double(0x7ff8000000000000LL) : c - (c - a) * (1.0 - t);
}

// Fold `nb` consecutive batches (partial blocks partials[slot_map[b]]) into the state, in order.
// grid = p CTAs of 1024 threads; CTA f owns covariance row f and draw-sum column f.  For the
// batches b in [own0, own1) the squared error draws z_sf^2 (covariance unbiased_cov / n after
// batch b) are written to zsq[b - own0][f][s].
__global__ void __launch_bounds__(1024) est_absorb_kernel(double *state, int p, int cur, double n_before,
                                                           const double *partials, size_t pstride,
                                                           const int *slot_map, int nb, int own0, int own1,
                                                           double *zsq, int with_draws) {
  extern __shared__ double smem[];
  double *mrun = smem;                           // (nb+1) x p running means
  double *nrun = mrun + (size_t)(nb + 1) * p;    // nb+1 running counts
  double *n2s = nrun + (nb + 1);                 // nb batch counts
  double *fa = n2s + nb;                         // per-batch merge weights (no divisions in the serial loops):
  double *fb = fa + nb;                          //   fa = n1/nn, fb = n2/nn, fc = 1/nn, fd = fa*fb,
  double *fc = fb + nb;                          //   fz = 1/sqrt(nn (nn-1))
  double *fd = fc + nb;
  double *fz = fd + nb;
  int *slots = reinterpret_cast<int *>(fz + nb + (nb & 1));  // nb block indices
  StateView st = view_state(state, p);
  const int nxt = cur ^ 1;
  const int tid = threadIdx.x;
  const int f = blockIdx.x;
  // The folds below are serial in b, but nothing they load depends on the running values: block
  // indices and counts go to shared memory first, and the loops are unrolled so that the global
  // loads of several batches are in flight together (they used to be one latency per batch).
  for (int b = tid; b < nb; b += blockDim.x) {
    const int sl = slot_map[b];
    slots[b] = sl;
    n2s[b] = partials[(size_t)sl * pstride];
  }
  for (int j = tid; j < p; j += blockDim.x) mrun[j] = st.mean[cur][j];
  __syncthreads();
  if (tid == 0) {
    double n = n_before;
    nrun[0] = n;
    for (int b = 0; b < nb; ++b) {
      n += n2s[b];
      nrun[b + 1] = n;
    }
  }
  __syncthreads();
  for (int b = tid; b < nb; b += blockDim.x) {
    const double n1 = nrun[b], n2 = n2s[b], nn = nrun[b + 1];
    fa[b] = n1 / nn;
    fb[b] = n2 / nn;
    fc[b] = 1.0 / nn;
    fd[b] = (n1 / nn) * (n2 / nn);
    fz[b] = 1.0 / sqrt(nn * (nn - 1.0));
  }
  __syncthreads();
  for (int j = tid; j < p; j += blockDim.x) {
    double m = mrun[j];
#pragma unroll 8
    for (int b = 0; b < nb; ++b) {
      const double pm = partials[(size_t)slots[b] * pstride + kPartHdr + j];
      if (n2s[b] > 0.0) m = fa[b] * m + fb[b] * pm;
      mrun[(size_t)(b + 1) * p + j] = m;
    }
  }
  __syncthreads();
  // ---- covariance row f (biased): Chan merge batch by batch (reference merge_sample_cov)
  for (int j = tid; j < p; j += blockDim.x) {
    double c = st.cov[(size_t)f * p + j];
#pragma unroll 8
    for (int b = 0; b < nb; ++b) {
      const double *blk = partials + (size_t)slots[b] * pstride + kPartHdr;
      const double pmf = blk[f], pmj = blk[j], pm2 = blk[p + (size_t)f * p + j];
      if (n2s[b] > 0.0) {
        const double df = mrun[(size_t)b * p + f] - pmf;
        const double dj = mrun[(size_t)b * p + j] - pmj;
        c = fa[b] * c + pm2 * fc[b] + fd[b] * df * dj;
      }
    }
    st.cov[(size_t)f * p + j] = c;
  }
  if (tid == 0) st.mean[nxt][f] = mrun[(size_t)nb * p + f];
  // ---- draw sums: S column f and G, re-centred on the moving mean
  if (with_draws) {
    double s = st.S[(size_t)f * kDraws + tid];
    double gr = st.G[cur][tid];
#pragma unroll 4
    for (int b = 0; b < nb; ++b) {
      const double *blk = partials + (size_t)slots[b] * pstride + kPartHdr;
      const double pmf = blk[f];
      const double g2 = blk[p + (size_t)p * p + tid];
      const double ps = blk[p + (size_t)p * p + kDraws + (size_t)f * kDraws + tid];
      if (n2s[b] > 0.0) {
        const double m_old = mrun[(size_t)b * p + f], m_new = mrun[(size_t)(b + 1) * p + f];
        s += (m_old - m_new) * gr + ps + (pmf - m_new) * g2;
        gr += g2;
      }
      if (zsq != nullptr && b >= own0 && b < own1) {
        const double z = s * fz[b];       // n = 1: 0 * inf = NaN, as the reference's 0 / 0
        zsq[((size_t)(b - own0) * (p + 1) + f) * kDraws + tid] = z * z;
      }
    }
    st.S[(size_t)f * kDraws + tid] = s;
    if (f == 0) st.G[nxt][tid] = gr;
  }
}

// 0.95 quantiles of the error draws of `nown` batches: grid (p + 1, nown), block 1024.
// CTA (f < p, b): |z_sf| over the draws -> feat_out[b][f];  CTA (p, b): |z_s|_2 -> overall_out[b].
// Sorting z^2 and taking the square root of the two order statistics before interpolating is the
// same as numpy.quantile(|z|, 0.95).
// zsq[b][p][s] = sum_f zsq[b][f][s] (the squared L2 norm of draw s, for the overall error): CTA =
// (batch, 128 draws), 8 feature groups x 128 draws, the groups are added in a fixed order.
__global__ void __launch_bounds__(1024) est_rowsum_kernel(int p, double *zsq) {
  __shared__ double part[8][128];
  const int b = blockIdx.y, s = blockIdx.x * 128 + (threadIdx.x & 127), grp = threadIdx.x >> 7;
  double *zb = zsq + (size_t)b * (p + 1) * kDraws;
  double a0 = 0.0, a1 = 0.0;
  int ff = grp;
  for (; ff + 8 < p; ff += 16) {
    a0 += zb[(size_t)ff * kDraws + s];
    a1 += zb[(size_t)(ff + 8) * kDraws + s];
  }
  if (ff < p) a0 += zb[(size_t)ff * kDraws + s];
  part[grp][threadIdx.x & 127] = a0 + a1;
  __syncthreads();
  if (grp == 0) {
    double t = 0.0;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += part[g][threadIdx.x];
    zb[(size_t)p * kDraws + s] = t;
  }
}

// One warp per (batch, feature) row of zsq, and one more per batch for the overall error (the sum
// over the features); rows = nown * (p + 1), 8 per CTA.
__global__ void __launch_bounds__(256) est_quantile_kernel(int p, int nown, const double *zsq, double *overall_out,
                                                           double *feat_out) {
  const int lane = threadIdx.x & 31;
  const int64_t rowid = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (rowid >= (int64_t)nown * (p + 1)) return;
  const int b = (int)(rowid / (p + 1)), f = (int)(rowid - (int64_t)b * (p + 1));
  const double *zb = zsq + ((size_t)b * (p + 1) + f) * kDraws;   // row p = sum over the features (est_rowsum_kernel)
  double v[kPerLane];
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) v[i] = zb[32 * i + lane];
  const double q = quantile95_sqrt(v, lane);
  if (lane == 0) {
    if (f < p) feat_out[(size_t)b * p + f] = q;
    else overall_out[b] = q;
  }
}

// ---------------------------------------------------------------- draw sums + error norms in one pass
// The draw-sum half of est_absorb_kernel for callers that only need what the sample loop consumes: the
// OVERALL error after every owned batch and the per-feature errors after ONE batch (the last one, or the
// batch the job stops at).  CTA = (group of kErrFG features, 256 draws); a thread folds its draw for the
// group's features (independent chains) and adds their squared z to the group's share of |z_s|^2, so the
// 100 MB of squared draws per super-batch that est_absorb_kernel writes (and est_rowsum / est_quantile
// read back) shrink to [owned batch][group][draw] partial norms.  Same recurrences, same order in b.
constexpr int kErrFG = 4;

__global__ void __launch_bounds__(256) est_draws_kernel(double *__restrict__ state, int p, int cur, double n_before,
                                                         const double *__restrict__ partials, size_t pstride,
                                                         const int *__restrict__ slot_map, int nb, int own0, int own1,
                                                         int feat_batch, double *__restrict__ pnorm,
                                                         double *__restrict__ zlast) {
  extern __shared__ double smem[];
  double *nrun = smem;                       // nb + 1 running counts
  double *n2s = nrun + (nb + 1);             // nb batch counts
  double *fz = n2s + nb;                     // 1 / sqrt(nn (nn - 1))
  double *fa = fz + nb;                      // n1 / nn
  double *fb = fa + nb;                      // n2 / nn
  double *pm = fb + nb;                      // nb x FG batch means of the group's features
  double *d1 = pm + (size_t)nb * kErrFG;     // nb x FG  m_old - m_new
  double *d2 = d1 + (size_t)nb * kErrFG;     // nb x FG  batch mean - m_new
  int *slots = reinterpret_cast<int *>(d2 + (size_t)nb * kErrFG);
  StateView st = view_state(state, p);
  const int nxt = cur ^ 1;
  const int tid = threadIdx.x;
  const int grp = blockIdx.x, ngrp = gridDim.x;
  const int f0 = grp * kErrFG;
  const int s = blockIdx.y * 256 + tid;
  for (int b = tid; b < nb; b += 256) {
    const int sl = slot_map[b];
    slots[b] = sl;
    n2s[b] = partials[(size_t)sl * pstride];
  }
  __syncthreads();
  for (int e = tid; e < nb * kErrFG; e += 256) {
    const int b = e / kErrFG, k = e - b * kErrFG;
    pm[e] = (f0 + k < p) ? partials[(size_t)slots[b] * pstride + kPartHdr + f0 + k] : 0.0;
  }
  if (tid == 0) {
    double n = n_before;
    nrun[0] = n;
    for (int b = 0; b < nb; ++b) {
      n += n2s[b];
      nrun[b + 1] = n;
    }
  }
  __syncthreads();
  for (int b = tid; b < nb; b += 256) {
    const double nn = nrun[b + 1];
    fz[b] = 1.0 / sqrt(nn * (nn - 1.0));
    fa[b] = nrun[b] / nn;
    fb[b] = n2s[b] / nn;
  }
  __syncthreads();
  if (tid < kErrFG) {
    // running mean of one feature: the recurrence of est_absorb_kernel (m = n1/nn m + n2/nn pm)
    const int k = tid;
    double m = (f0 + k < p) ? st.mean[cur][f0 + k] : 0.0;
    for (int b = 0; b < nb; ++b) {
      double mn = m;
      if (n2s[b] > 0.0) mn = fa[b] * m + fb[b] * pm[b * kErrFG + k];
      d1[b * kErrFG + k] = m - mn;
      d2[b * kErrFG + k] = pm[b * kErrFG + k] - mn;
      m = mn;
    }
  }
  __syncthreads();
  double sk[kErrFG];
#pragma unroll
  for (int k = 0; k < kErrFG; ++k) sk[k] = (f0 + k < p) ? st.S[(size_t)(f0 + k) * kDraws + s] : 0.0;
  double gr = st.G[cur][s];
  const size_t offG = kPartHdr + (size_t)p + (size_t)p * p;
  constexpr int UB = 8;       // batches whose loads are in flight together
  for (int b0 = 0; b0 < nb; b0 += UB) {
    double g2[UB], ps[UB][kErrFG];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int b = (b0 + u < nb) ? b0 + u : nb - 1;
      const double *blk = partials + (size_t)slots[b] * pstride + offG;
      g2[u] = blk[s];
#pragma unroll
      for (int k = 0; k < kErrFG; ++k)
        ps[u][k] = blk[kDraws + (size_t)((f0 + k < p) ? f0 + k : p - 1) * kDraws + s];
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int b = b0 + u;
      if (b < nb) {
        if (n2s[b] > 0.0) {
#pragma unroll
          for (int k = 0; k < kErrFG; ++k) sk[k] += d1[b * kErrFG + k] * gr + ps[u][k] + d2[b * kErrFG + k] * g2[u];
          gr += g2[u];
        }
        if (b >= own0 && b < own1) {
          const double w = fz[b];       // n = 1: 0 * inf = NaN, as the reference's 0 / 0
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k < kErrFG; ++k) {
            const double z = sk[k] * w;
            if (f0 + k < p) {
              acc += z * z;
              if (b == feat_batch) zlast[(size_t)(f0 + k) * kDraws + s] = z * z;
            }
          }
          pnorm[((size_t)(b - own0) * ngrp + grp) * kDraws + s] = acc;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kErrFG; ++k)
    if (f0 + k < p) st.S[(size_t)(f0 + k) * kDraws + s] = sk[k];
  if (grp == 0) st.G[nxt][s] = gr;
}

// |z_s|^2 of every owned batch: the group partial norms of est_draws_kernel added in a fixed order
__global__ void est_normsum_kernel(int nown, int ngrp, const double *__restrict__ pnorm, double *__restrict__ norm2) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)nown * kDraws) return;
  const size_t b = e / kDraws, sd = e - b * kDraws;
  const double *pb = pnorm + b * ngrp * kDraws + sd;
  double acc = 0.0;
#pragma unroll 8
  for (int g = 0; g < ngrp; ++g) acc += pb[(size_t)g * kDraws];
  norm2[e] = acc;
}

// Quantiles for est_draws_kernel: warp per row; rows [0, nown) = overall error of an owned batch (from
// est_normsum_kernel), rows [nown, nown + p) = features of the one batch whose squared draws were kept
// (only when feat_out != nullptr).
__global__ void __launch_bounds__(256) est_quantile_fused_kernel(int p, int nown, const double *norm2,
                                                                  const double *zlast, double *overall_out,
                                                                  double *feat_out) {
  const int lane = threadIdx.x & 31;
  const int rowid = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int rows = nown + (feat_out != nullptr ? p : 0);
  if (rowid >= rows) return;
  const double *zb = rowid < nown ? norm2 + (size_t)rowid * kDraws : zlast + (size_t)(rowid - nown) * kDraws;
  double v[kPerLane];
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) v[i] = zb[32 * i + lane];
  const double q = quantile95_sqrt(v, lane);
  if (lane == 0) {
    if (rowid < nown) overall_out[rowid] = q;
    else feat_out[rowid - nown] = q;
  }
}

// ---------------------------------------------------------------- run total of a rank (multi-GPU)
// One partial block equal to the Chan merge of nb consecutive partial blocks, computed as PARALLEL sums
// (the merge is associative): n = sum n_b, mean = sum n_b m_b / n, and with d_b = m_b - mean
//   M2 = sum_b M2_b + n_b d_b d_b^T,   G = sum_b G_b,   S = sum_b S_b + d_b G_b^T.
// A rank ships this block instead of folding its batches twice; the sequential fold (est_absorb_kernel)
// and this sum agree to rounding.
// The kernels of that sum (lsspa_estimator_block_total, and the moment half of lsspa_estimator_absorb_errors):
// the batches are ALSO spread over threads -- thread (element, chunk c) adds the batches c, c + 8, ... (their
// loads are independent and in flight together), the eight chunk sums are added in a fixed order.
// run_mean: CTA = 128 features x 8 chunks; run_total: CTA = 32 elements x 8 chunks.  slot_map: block b is
// partials[slot_map[b]] (nullptr = consecutive blocks).
__global__ void __launch_bounds__(1024) run_mean_kernel(int p, const double *__restrict__ partials, size_t pstride, int nb,
                                                         const int *__restrict__ slot_map, double *__restrict__ out) {
  __shared__ double part[8][128];
  __shared__ double s_n;
  const int fl = threadIdx.x & 127, ch = threadIdx.x >> 7;
  const int f = blockIdx.x * 128 + fl;
  double cnt = 0.0, acc = 0.0;
  for (int b = ch; b < nb; b += 8) {
    const double *blk = partials + (size_t)(slot_map ? slot_map[b] : b) * pstride;
    const double n_b = blk[0], m = blk[kPartHdr + (f < p ? f : 0)];
    cnt += n_b;
    if (n_b > 0.0) acc = fma(n_b, m, acc);
  }
  part[ch][fl] = acc;
  __syncthreads();
  // total count: every chunk row saw all its batches; add the eight chunk counts through shared memory
  __shared__ double cpart[8];
  if (fl == 0) cpart[ch] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    double n = 0.0;
    for (int c = 0; c < 8; ++c) n += cpart[c];
    s_n = n;
  }
  __syncthreads();
  const double n = s_n;
  if (blockIdx.x == 0 && threadIdx.x < kPartHdr) out[threadIdx.x] = (threadIdx.x == 0) ? n : 0.0;
  if (ch == 0 && f < p) {
    double t = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) t += part[c][fl];
    out[kPartHdr + f] = n > 0.0 ? t / n : 0.0;
  }
}

__global__ void __launch_bounds__(256) run_total_kernel(int p, const double *__restrict__ partials, size_t pstride,
                                                         int nb, const int *__restrict__ slot_map, int with_draws,
                                                         double *__restrict__ out) {
  __shared__ double part[8][32];
  const int el = threadIdx.x & 31, ch = threadIdx.x >> 5;
  const size_t nm2 = (size_t)p * p, ng = kDraws, ns = (size_t)p * kDraws;
  const size_t total = nm2 + (with_draws ? ng + ns : 0);
  const size_t e = (size_t)blockIdx.x * 32 + el;
  const size_t ec = e < total ? e : total - 1;
  const double *mean = out + kPartHdr;
  double acc = 0.0;
  if (ec < nm2) {                 // M2 = sum_b M2_b + n_b d_b d_b^T
    const int f = (int)(ec / p), j = (int)(ec - (size_t)f * p);
    const double mf = mean[f], mj = mean[j];
#pragma unroll 4
    for (int b = ch; b < nb; b += 8) {
      const double *blk = partials + (size_t)(slot_map ? slot_map[b] : b) * pstride;
      const double n_b = blk[0], m2 = blk[kPartHdr + p + ec], bf = blk[kPartHdr + f], bj = blk[kPartHdr + j];
      if (n_b > 0.0) acc += m2 + n_b * (bf - mf) * (bj - mj);
    }
  } else if (ec < nm2 + ng) {     // G = sum_b G_b
#pragma unroll 4
    for (int b = ch; b < nb; b += 8) {
      const double *blk = partials + (size_t)(slot_map ? slot_map[b] : b) * pstride;
      const double n_b = blk[0], g = blk[kPartHdr + p + ec];
      if (n_b > 0.0) acc += g;
    }
  } else {                        // S = sum_b S_b + d_b G_b^T
    const size_t r = ec - nm2 - ng;
    const int f = (int)(r / kDraws), sd = (int)(r - (size_t)f * kDraws);
    const double mf = mean[f];
#pragma unroll 4
    for (int b = ch; b < nb; b += 8) {
      const double *blk = partials + (size_t)(slot_map ? slot_map[b] : b) * pstride;
      const double n_b = blk[0], sv = blk[kPartHdr + p + ec], bf = blk[kPartHdr + f], g = blk[kPartHdr + p + nm2 + sd];
      if (n_b > 0.0) acc += sv + (bf - mf) * g;
    }
  }
  part[ch][el] = acc;
  __syncthreads();
  if (ch == 0 && e < total) {
    double t = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) t += part[c][el];
    out[kPartHdr + p + e] = t;
  }
}

// ---------------------------------------------------------------- standalone error_estimates
// L L^T = cov for a positive semi-definite cov (one CTA, right-looking, L row-major in `L`): a pivot
// that is not above 1e-13 of the largest diagonal entry is round-off of a singular direction -- its
// column is set to zero, exactly what a factorisation of a PSD matrix of that rank leaves there.
__global__ void __launch_bounds__(1024) psd_factor_kernel(int p, const double *cov, double *L) {
  __shared__ double s_piv;
  const int tid = threadIdx.x, nt = blockDim.x;
  double dmax = 0.0;
  for (int i = 0; i < p; ++i) dmax = fmax(dmax, cov[(size_t)i * p + i]);
  for (int e = tid; e < p * p; e += nt) {
    const int i = e / p, j = e - i * p;
    L[e] = (j <= i) ? cov[e] : 0.0;
  }
  __syncthreads();
  for (int k = 0; k < p; ++k) {
    if (tid == 0) {
      const double d = L[(size_t)k * p + k];
      s_piv = (d > 1e-13 * dmax && d > 0.0) ? 1.0 / sqrt(d) : 0.0;
    }
    __syncthreads();
    const double ri = s_piv;
    for (int i = k + tid; i < p; i += nt) L[(size_t)i * p + k] *= ri;     // column k (incl. the diagonal: d / sqrt(d))
    __syncthreads();
    const int m = p - k - 1;
    for (int e = tid; e < m * m; e += nt) {
      const int i = k + 1 + e / m, j = k + 1 + e % m;
      if (j <= i) L[(size_t)i * p + j] = fma(-L[(size_t)i * p + k], L[(size_t)j * p + k], L[(size_t)i * p + j]);
    }
    __syncthreads();
  }
}

// zsq[f][s] = (sum_k L[f][k] g[s][k])^2: block f, thread s
__global__ void __launch_bounds__(kDraws) err_draws_kernel(int p, const double *L, uint64_t seed, double *zsq) {
  const int f = blockIdx.x, s = threadIdx.x;
  double z = 0.0;
  for (int k = 0; k <= f; ++k) {
    float g0, g1;
    gauss_pair(seed, (uint64_t)k, (uint32_t)s >> 1, g0, g1);   // column k, draws 2 (s/2) and 2 (s/2) + 1
    z = fma(L[(size_t)f * p + k], (double)((s & 1) ? g1 : g0), z);
  }
  zsq[(size_t)f * kDraws + s] = z * z;
}

// ---------------------------------------------------------------- history / merge / epilogue
__global__ void prefix_means_kernel(int p, const double *lifts, int64_t rows, double *carry_sum,
                                    double carry_count, double *hist) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= p) return;
  double s = carry_sum[f];
  for (int64_t k = 0; k < rows; ++k) {
    s += lifts[k * p + f];
    hist[k * p + f] = s / (carry_count + (double)(k + 1));
  }
  carry_sum[f] = s;
}

__global__ void merge_moments_kernel(int p, double *mean, double *cov, double n1, const double *mean2,
                                     const double *cov2, double n2) {
  // cov first (needs the old mean), elementwise over the p x p grid; the mean is updated
  // by a second launch so that no thread reads a half-updated vector
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)p * p) return;
  const int i = (int)(e / p), j = (int)(e % p);
  const double nn = n1 + n2;
  const double di = mean[i] - mean2[i], dj = mean[j] - mean2[j];
  const double c2 = cov2 ? cov2[e] : 0.0;
  cov[e] = (n1 / nn) * cov[e] + (n2 / nn) * c2 + (n1 / nn) * (n2 / nn) * di * dj;
}
__global__ void merge_mean_kernel(int p, double *mean, double n1, const double *mean2, double n2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p) return;
  const double nn = n1 + n2;
  mean[i] = (n1 / nn) * mean[i] + (n2 / nn) * mean2[i];
}

// theta = minimum-norm least-squares solution of R_tr theta = c_tr, as np.linalg.lstsq(...,
// rcond=None) returns it (reference ls_spa/ls_spa.py:240): one-sided (Hestenes) Jacobi SVD of
// R_tr, singular values below eps * p * sigma_max dropped.  The reference's own test data
// contain an exactly rank-deficient train block (100 centred rows, 100 features), so a plain
// triangular solve is not enough.  r2 = (|c_te|^2 - |c_te - R_te theta|^2) / ynsq  (:241-243).
// One CTA; A (copy of R_tr, column-major) and V live in the global workspace (L2-resident).
__global__ void __launch_bounds__(512) theta_r2_kernel(int p, const double *Rtr, const double *ctr,
                                                        const double *Rte, const double *cte, double ynsq,
                                                        double *out, double *ws, int in_smem) {
  extern __shared__ double sm[];
  double *coef = sm, *theta = sm + p, *red = sm + 2 * p;  // red[64]
  __shared__ int s_rot;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  // Fast path: when the diagonal of R_tr shows no sign of rank deficiency (min |R_kk| > 1e-7 max),
  // lstsq's singular-value cut-off (eps * p * sigma_max) is not active and the minimum-norm
  // solution is R^-1 c: plain back substitution.  Otherwise the Jacobi SVD below.
  __shared__ int s_fast;
  {
    double dmin = 1e300, dmax = 0.0;
    for (int i = tid; i < p; i += nt) {
      const double d = fabs(Rtr[(size_t)i * p + i]);
      dmin = fmin(dmin, d);
      dmax = fmax(dmax, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      dmin = fmin(dmin, __shfl_xor_sync(kFull, dmin, o));
      dmax = fmax(dmax, __shfl_xor_sync(kFull, dmax, o));
    }
    if (lane == 0) {
      red[warp] = dmin;
      red[32 + warp] = dmax;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < nwarps; ++w) {
        dmin = fmin(dmin, red[w]);
        dmax = fmax(dmax, red[32 + w]);
      }
      s_fast = (dmin > 1e-7 * dmax) ? 1 : 0;
    }
    __syncthreads();
  }
  if (s_fast) {
    for (int i = tid; i < p; i += nt) coef[i] = ctr[i];
    __syncthreads();
    for (int j = p - 1; j >= 0; --j) {
      if (tid == 0) theta[j] = coef[j] / Rtr[(size_t)j * p + j];
      __syncthreads();
      const double tj = theta[j];
      for (int i = tid; i < j; i += nt) coef[i] = fma(-Rtr[(size_t)j * p + i], tj, coef[i]);
      __syncthreads();
    }
  } else {
    // Jacobi work matrices: shared memory when 2 p^2 doubles fit (p <= ~118), else the workspace
    double *A = in_smem ? red + 64 : ws;
    double *V = A + (size_t)p * p;
    for (int e = tid; e < p * p; e += nt) {
      const int j = e / p, i = e - j * p;
      A[e] = (i <= j) ? Rtr[e] : 0.0;
      V[e] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    const int n = p + (p & 1);
    // stop when every pair is orthogonal to ~sqrt(p) eps (the dgesvj criterion); a tighter bound
    // sits below the round-off of the dot products and the sweeps never end
    const double jtol = 4.0 * 2.220446049250313e-16 * sqrt((double)p);
    for (int sweep = 0; sweep < 40; ++sweep) {
      if (tid == 0) s_rot = 0;
      __syncthreads();
      for (int s = 0; s < n - 1; ++s) {
        for (int k = warp; k < n / 2; k += nwarps) {
          int i = (k == 0) ? n - 1 : (s + k) % (n - 1);
          int j = (k == 0) ? s : (s - k + (n - 1)) % (n - 1);
          if (i > j) { const int t = i; i = j; j = t; }
          if (j >= p) continue;  // dummy column of an odd-sized tournament
          double *ai = A + (size_t)i * p, *aj = A + (size_t)j * p;
          double al = 0.0, be = 0.0, ga = 0.0;
          for (int r = lane; r < p; r += kWarp) {
            const double x = ai[r], y = aj[r];
            al = fma(x, x, al);
            be = fma(y, y, be);
            ga = fma(x, y, ga);
          }
          al = warp_sum(al);
          be = warp_sum(be);
          ga = warp_sum(ga);
          if (fabs(ga) > jtol * sqrt(al * be) && ga != 0.0) {
            const double zeta = (be - al) / (2.0 * ga);
            const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
            double *vi = V + (size_t)i * p, *vj = V + (size_t)j * p;
            for (int r = lane; r < p; r += kWarp) {
              const double x = ai[r], y = aj[r];
              ai[r] = c * x - sn * y;
              aj[r] = sn * x + c * y;
              const double u = vi[r], w = vj[r];
              vi[r] = c * u - sn * w;
              vj[r] = sn * u + c * w;
            }
            if (lane == 0) s_rot = 1;
          }
        }
        __syncthreads();
      }
      const int any = s_rot;
      __syncthreads();
      if (!any) break;
    }
    // coef_i = (a_i . c) / sigma_i^2 for the retained singular directions
    double smax = 0.0;
    for (int i = warp; i < p; i += nwarps) {
      const double *ai = A + (size_t)i * p;
      double nn = 0.0, dc = 0.0;
      for (int r = lane; r < p; r += kWarp) {
        nn = fma(ai[r], ai[r], nn);
        dc = fma(ai[r], ctr[r], dc);
      }
      nn = warp_sum(nn);
      dc = warp_sum(dc);
      if (lane == 0) {
        coef[i] = dc;
        theta[i] = nn;  // sigma_i^2, reused as scratch
      }
      smax = fmax(smax, nn);
    }
    if (lane == 0) red[warp] = smax;
    __syncthreads();
    smax = 0.0;
    for (int w = 0; w < nwarps; ++w) smax = fmax(smax, red[w]);
    const double rcond = 2.220446049250313e-16 * (double)p;
    const double cut = smax * rcond * rcond;  // compare squared singular values
    __syncthreads();
    for (int i = tid; i < p; i += nt) coef[i] = (theta[i] > cut && theta[i] > 0.0) ? coef[i] / theta[i] : 0.0;
    __syncthreads();
    for (int r = tid; r < p; r += nt) {
      double acc = 0.0;
      for (int i = 0; i < p; ++i) acc = fma(V[(size_t)i * p + r], coef[i], acc);
      theta[r] = acc;
    }
    __syncthreads();
  }
  double s_c = 0.0, s_r = 0.0;
  for (int i = tid; i < p; i += nt) {
    double pred = 0.0;
    for (int j = i; j < p; ++j) pred = fma(Rte[(size_t)j * p + i], theta[j], pred);
    const double c = cte[i], r = c - pred;
    s_c = fma(c, c, s_c);
    s_r = fma(r, r, s_r);
    out[i] = theta[i];
  }
  s_c = warp_sum(s_c);
  s_r = warp_sum(s_r);
  __syncthreads();
  if (lane == 0) {
    red[warp] = s_c;
    red[32 + warp] = s_r;
  }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, r = 0.0;
    for (int w = 0; w < nwarps; ++w) {
      a += red[w];
      r += red[32 + w];
    }
    // same association as the reference: (|c_te|^2 - |resid|^2) / |y_test|^2
    out[p] = (a - r) / ynsq;
  }
}

}  // namespace lsspa

using namespace lsspa;

extern "C" size_t lsspa_estimator_state_bytes(int p) {
  if (p < 1) return 0;
  return state_doubles(p) * sizeof(double);
}

extern "C" int64_t lsspa_estimator_partial_doubles(int p) {
  return p < 1 ? 0 : (int64_t)partial_doubles(p);
}

template <int MAXFT>
static int launch_part_mma(int p, const double *lifts, const int64_t *desc, int nbatch, uint64_t seed, int estimate_errors,
                           double *partials, size_t pstride, size_t smem, cudaStream_t st) {
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(part_m2_mma_kernel<MAXFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  part_m2_mma_kernel<MAXFT><<<nbatch, 256, smem, st>>>(p, lifts, desc, partials, pstride);
  LSSPA_LAUNCH_CHECK();
  if (estimate_errors) {
    LSSPA_CUDA_TRY(cudaFuncSetAttribute(part_s_mma_kernel<MAXFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    part_s_mma_kernel<MAXFT><<<dim3(kDraws / kSDraws, (unsigned)nbatch), 256, smem, st>>>(p, lifts, desc, seed, partials, pstride);
    LSSPA_LAUNCH_CHECK();
  }
  return LSSPA_OK;
}

extern "C" int lsspa_estimator_partials(int p, const double *lifts, const int64_t *batch_desc, int nbatch,
                                        uint64_t seed, int estimate_errors, double *partials,
                                        void *stream) {
  if (p < 1 || !lifts || !batch_desc || !partials || nbatch < 0) return LSSPA_E_BADARG;
  if (nbatch == 0) return LSSPA_OK;
  cudaStream_t st = as_stream(stream);
  const size_t pstride = partial_doubles(p);
  int nt = ((p + 31) / 32) * 32;
  if (nt > 1024) nt = 1024;
  if (nt < 32) nt = 32;
  part_mean_kernel<<<nbatch, nt, 0, st>>>(p, lifts, batch_desc, partials, pstride);
  LSSPA_LAUNCH_CHECK();
  if (p >= 17 && p <= 128) {
    // tensor-pipe versions (the scalar kernels below remain for every other width)
    const int FT = (p + 7) / 8;
    const size_t smem = (size_t)kMmaRows * mma_ldx(FT) * sizeof(double);
    int rc;
    if (FT <= 8) rc = launch_part_mma<8>(p, lifts, batch_desc, nbatch, seed, estimate_errors, partials, pstride, smem, st);
    else if (FT <= 13) rc = launch_part_mma<13>(p, lifts, batch_desc, nbatch, seed, estimate_errors, partials, pstride, smem, st);
    else rc = launch_part_mma<16>(p, lifts, batch_desc, nbatch, seed, estimate_errors, partials, pstride, smem, st);
    return rc;
  }
  dim3 g2((unsigned)ceil_div(p, 16), (unsigned)ceil_div(p, 16), (unsigned)nbatch);
  part_m2_kernel<<<g2, dim3(16, 16), 0, st>>>(p, lifts, batch_desc, partials, pstride);
  LSSPA_LAUNCH_CHECK();
  if (estimate_errors) {
    dim3 g3(kDraws / kDrawTile, (unsigned)nbatch);
    part_s_kernel<<<g3, nt < 128 ? 128 : nt, 0, st>>>(p, lifts, batch_desc, seed, partials, pstride);
    LSSPA_LAUNCH_CHECK();
  }
  return LSSPA_OK;
}

extern "C" int lsspa_estimator_max_batches(int p) {
  // running means, counts and block indices of all batches of one absorb call live in shared
  // memory: (nb + 1) * (p + 1) + 7 nb + 16 doubles
  const DeviceInfo &d = device_info();
  const size_t limit = (size_t)(d.smem_optin > 0 ? d.smem_optin : 227 * 1024) - 1024;
  long nb = (long)(limit / ((size_t)(p + 8) * sizeof(double))) - 3;
  if (nb > 4096) nb = 4096;
  return nb < 1 ? 0 : (int)nb;
}

extern "C" int lsspa_estimator_absorb(void *state, int p, int cur, double n_before, const double *partials,
                                      const int32_t *slot_map, int nb, int own0, int own1, double *zsq,
                                      int with_draws, void *stream) {
  if (!state || !partials || !slot_map || p < 1 || nb < 0 || (cur != 0 && cur != 1)) return LSSPA_E_BADARG;
  if (nb == 0) return LSSPA_OK;
  if (nb > lsspa_estimator_max_batches(p)) return LSSPA_E_UNSUPPORTED;
  const size_t smem = ((size_t)(nb + 1) * p + (nb + 1) + 7 * (size_t)nb + 16) * sizeof(double);
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(est_absorb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  est_absorb_kernel<<<p, kDraws, smem, as_stream(stream)>>>(reinterpret_cast<double *>(state), p, cur, n_before,
                                                            partials, partial_doubles(p), slot_map, nb, own0,
                                                            own1, zsq, with_draws);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

extern "C" int lsspa_estimator_quantiles(int p, double *zsq, int nown, double *overall_out,
                                         double *feat_out, void *stream) {
  if (!zsq || !overall_out || !feat_out || p < 1 || nown < 0) return LSSPA_E_BADARG;
  if (nown == 0) return LSSPA_OK;
  est_rowsum_kernel<<<dim3(kDraws / 128, (unsigned)nown), 1024, 0, as_stream(stream)>>>(p, zsq);
  LSSPA_LAUNCH_CHECK();
  const int64_t rows = (int64_t)nown * (p + 1);
  est_quantile_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, as_stream(stream)>>>(p, nown, zsq, overall_out, feat_out);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

extern "C" int64_t lsspa_estimator_errors_workspace_doubles(int p, int nown) {
  if (p < 1 || nown < 0) return 0;
  const int ngrp = (p + kErrFG - 1) / kErrFG;
  // group partial norms, |z|^2 per owned batch, squared draws of one batch, the run total block, one slot index
  return ((int64_t)nown * ngrp + nown + p) * kDraws + (int64_t)partial_doubles(p) + 2;
}

extern "C" int lsspa_estimator_absorb_errors(void *state, int p, int cur, double n_before, const double *partials,
                                             const int32_t *slot_map, int nb, int own0, int own1, int feat_batch,
                                             double *overall_out, double *feat_out, double *workspace,
                                             void *stream) {
  if (!state || !partials || !slot_map || p < 1 || nb < 0 || (cur != 0 && cur != 1)) return LSSPA_E_BADARG;
  if (own0 < 0 || own1 > nb || own1 < own0 || !workspace) return LSSPA_E_BADARG;
  const int nown = own1 - own0;
  if (nown > 0 && !overall_out) return LSSPA_E_BADARG;
  if (feat_batch >= 0 && (feat_batch < own0 || feat_batch >= own1 || !feat_out)) return LSSPA_E_BADARG;
  if (nb == 0) return LSSPA_OK;
  if (nb > lsspa_estimator_max_batches(p)) return LSSPA_E_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  const size_t pstride = partial_doubles(p);
  const int ngrp = (p + kErrFG - 1) / kErrFG;
  double *pnorm = workspace;
  double *norm2 = pnorm + (size_t)nown * ngrp * kDraws;
  double *zlast = norm2 + (size_t)nown * kDraws;
  double *total = zlast + (size_t)p * kDraws;
  int *slot0 = reinterpret_cast<int *>(total + pstride);
  // draw sums, error norms (reads the live mean; writes S, the other G copy and the norms)
  {
    const size_t smem = ((size_t)(nb + 1) + 4 * (size_t)nb + 3 * (size_t)nb * kErrFG) * sizeof(double) +
                        (size_t)nb * sizeof(int) + 16;
    LSSPA_CUDA_TRY(cudaFuncSetAttribute(est_draws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    est_draws_kernel<<<dim3((unsigned)ngrp, kDraws / 256), 256, smem, st>>>(
        reinterpret_cast<double *>(state), p, cur, n_before, partials, pstride, slot_map, nb, own0, own1, feat_batch,
        pnorm, zlast);
    LSSPA_LAUNCH_CHECK();
  }
  // mean and covariance: the moments of the whole run as PARALLEL sums (merge_sample_mean / merge_sample_cov
  // are associative, reference test/test_ls_spa.py:20-44), then one Chan merge into the state -- instead of a
  // fold whose nb dependent steps each wait for a load
  {
    run_mean_kernel<<<(p + 127) / 128, 1024, 0, st>>>(p, partials, pstride, nb, slot_map, total);
    LSSPA_LAUNCH_CHECK();
    run_total_kernel<<<(unsigned)ceil_div((int64_t)p * p, 32), 256, 0, st>>>(p, partials, pstride, nb, slot_map, 0, total);
    LSSPA_LAUNCH_CHECK();
    LSSPA_CUDA_TRY(cudaMemsetAsync(slot0, 0, 2 * sizeof(int), st));
    const size_t smem = ((size_t)2 * p + 2 + 7 + 16) * sizeof(double);
    LSSPA_CUDA_TRY(cudaFuncSetAttribute(est_absorb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    est_absorb_kernel<<<p, 128, smem, st>>>(reinterpret_cast<double *>(state), p, cur, n_before, total, pstride, slot0,
                                            1, 0, 0, nullptr, 0);
    LSSPA_LAUNCH_CHECK();
  }
  if (nown > 0) {
    est_normsum_kernel<<<(unsigned)ceil_div((int64_t)nown * kDraws, 256), 256, 0, st>>>(nown, ngrp, pnorm, norm2);
    LSSPA_LAUNCH_CHECK();
    const int rows = nown + (feat_batch >= 0 ? p : 0);
    est_quantile_fused_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, st>>>(p, nown, norm2, zlast, overall_out,
                                                                        feat_batch >= 0 ? feat_out : nullptr);
    LSSPA_LAUNCH_CHECK();
  }
  return LSSPA_OK;
}

extern "C" int lsspa_estimator_block_total(int p, const double *partials, int nb, int with_draws, double *out_block,
                                           void *stream) {
  if (p < 1 || !partials || !out_block || nb < 0) return LSSPA_E_BADARG;
  cudaStream_t st = as_stream(stream);
  const size_t pstride = partial_doubles(p);
  if (nb == 0 || !with_draws) LSSPA_CUDA_TRY(cudaMemsetAsync(out_block, 0, pstride * sizeof(double), st));
  if (nb == 0) return LSSPA_OK;
  run_mean_kernel<<<(p + 127) / 128, 1024, 0, st>>>(p, partials, pstride, nb, nullptr, out_block);
  LSSPA_LAUNCH_CHECK();
  const int64_t total = (int64_t)p * p + (with_draws ? kDraws + (int64_t)p * kDraws : 0);
  run_total_kernel<<<(unsigned)ceil_div(total, 32), 256, 0, st>>>(p, partials, pstride, nb, nullptr, with_draws, out_block);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

extern "C" int lsspa_error_draws(int p, const double *cov, uint64_t seed, double *zsq, double *workspace,
                                 void *stream) {
  if (p < 1 || !cov || !zsq || !workspace) return LSSPA_E_BADARG;
  cudaStream_t st = as_stream(stream);
  psd_factor_kernel<<<1, 1024, 0, st>>>(p, cov, workspace);
  LSSPA_LAUNCH_CHECK();
  err_draws_kernel<<<p, kDraws, 0, st>>>(p, workspace, seed, zsq);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

extern "C" int lsspa_prefix_means(int p, const double *lifts, int64_t rows, double *carry_sum,
                                  double carry_count, double *hist_out, void *stream) {
  if (p < 1 || !lifts || !carry_sum || !hist_out || rows < 0) return LSSPA_E_BADARG;
  if (rows == 0) return LSSPA_OK;
  prefix_means_kernel<<<(unsigned)ceil_div(p, 128), 128, 0, as_stream(stream)>>>(p, lifts, rows, carry_sum,
                                                                                carry_count, hist_out);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

extern "C" int lsspa_merge_moments(int p, double *mean, double *cov, double old_n, const double *new_mean,
                                   const double *new_cov_or_null, double new_n, void *stream) {
  if (p < 1 || !mean || !new_mean || old_n + new_n <= 0.0) return LSSPA_E_BADARG;
  cudaStream_t st = as_stream(stream);
  if (cov) {
    merge_moments_kernel<<<(unsigned)ceil_div((int64_t)p * p, 256), 256, 0, st>>>(p, mean, cov, old_n, new_mean,
                                                                                 new_cov_or_null, new_n);
    LSSPA_LAUNCH_CHECK();
  }
  merge_mean_kernel<<<(unsigned)ceil_div(p, 256), 256, 0, st>>>(p, mean, old_n, new_mean, new_n);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

extern "C" size_t lsspa_theta_r2_workspace_bytes(int p) {
  return p < 1 ? 0 : 2 * (size_t)p * p * sizeof(double);
}

extern "C" int lsspa_theta_r2(int p, const double *R_tr_cm, const double *c_tr, const double *R_te_cm,
                              const double *c_te, double y_norm_sq, double *out, void *workspace,
                              size_t workspace_bytes, void *stream) {
  if (p < 1 || !R_tr_cm || !c_tr || !R_te_cm || !c_te || !out) return LSSPA_E_BADARG;
  if (!workspace || workspace_bytes < lsspa_theta_r2_workspace_bytes(p)) return LSSPA_E_WORKSPACE;
  size_t smem = (size_t)(2 * p + 64) * sizeof(double);
  const size_t big = smem + 2 * (size_t)p * p * sizeof(double);
  const DeviceInfo &d = device_info();
  const int in_smem = big <= (size_t)(d.smem_optin > 0 ? d.smem_optin : 227 * 1024) ? 1 : 0;
  if (in_smem) smem = big;
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(theta_r2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  theta_r2_kernel<<<1, 512, smem, as_stream(stream)>>>(p, R_tr_cm, c_tr, R_te_cm, c_te, y_norm_sq, out,
                                                        reinterpret_cast<double *>(workspace), in_smem);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}
