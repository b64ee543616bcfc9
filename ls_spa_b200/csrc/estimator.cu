// Online estimator of LS-SPA on the device.
//
// Replaces the bookkeeping of the reference's sample loop (ls_spa/ls_spa.py:186-236):
//   merge_sample_mean / merge_sample_cov (:103-119)  -> Chan merge of batch moments
//   error_estimates (:321-341)                        -> factor-free Gaussian draws
//   stop test (:229), error_history (:225)            -> device flag + history array
//   attribution_history (:217-219)                    -> prefix means
//   theta / r_squared epilogue (:240-243)             -> triangular solve + residual
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
constexpr int kHdr = 16;
// header slots (doubles)
enum { H_N = 0, H_STOP, H_NHIST, H_OVERALL, H_TOL, H_CUR, H_EST, H_MAXH };

struct StateView {
  double *hdr;
  double *mean[2];
  double *cov[2];
  double *G[2];
  double *S[2];  // feature-major [p][kDraws]
  double *feat_err;
  double *feat_err_tmp;
  double *err_hist;
  double *zsq;      // [p][kDraws] scratch of squared draws (overall-error reduction)
  double *ticket;   // 8 bytes used as an unsigned counter
};

__host__ __device__ inline size_t state_doubles(int p, int max_batches) {
  return kHdr + 2 * (size_t)p + 2 * (size_t)p * p + 2 * (size_t)kDraws + 2 * (size_t)p * kDraws +
         2 * (size_t)p + (size_t)max_batches + 8 + (size_t)p * kDraws + 8;
}

__host__ __device__ inline StateView view_state(double *base, int p, int max_batches) {
  StateView v;
  double *c = base;
  v.hdr = c; c += kHdr;
  v.mean[0] = c; c += p;
  v.mean[1] = c; c += p;
  v.cov[0] = c; c += (size_t)p * p;
  v.cov[1] = c; c += (size_t)p * p;
  v.G[0] = c; c += kDraws;
  v.G[1] = c; c += kDraws;
  v.S[0] = c; c += (size_t)p * kDraws;
  v.S[1] = c; c += (size_t)p * kDraws;
  v.feat_err = c; c += p;
  v.feat_err_tmp = c; c += p;
  v.err_hist = c; c += max_batches + 8;
  v.zsq = c; c += (size_t)p * kDraws;
  v.ticket = c;
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
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
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

// ---------------------------------------------------------------- batch step
__device__ __forceinline__ void bitonic_sort_1024(double *buf, int tid) {
  // blockDim.x == 1024, ascending
  for (int k = 2; k <= kDraws; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      const int ixj = tid ^ j;
      if (ixj > tid) {
        const double a = buf[tid], b = buf[ixj];
        const bool up = (tid & k) == 0;
        if ((a > b) == up) {
          buf[tid] = b;
          buf[ixj] = a;
        }
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ double quantile95_sorted(const double *buf) {
  // numpy.quantile(..., 0.95), default 'linear' method on n = 1024 sorted values
  const double pos = (double)(kDraws - 1) * 0.95;
  const int lo = (int)floor(pos);
  const double t = pos - (double)lo;
  const double a = buf[lo], b = buf[lo + 1];
  return b - (b - a) * (1.0 - t);
}

// One launch per batch: grid = p CTAs of 1024 threads.  CTA f merges row f of the covariance,
// re-centres column f of the draw sums, takes the 0.95 quantile of |z_sf| and leaves z_sf^2 in a
// scratch column.  The CTA that finishes last (atomic ticket) adds the scratch columns in feature
// order (deterministic: every rank of a multi-GPU job must take the same stop decision), takes the
// quantile of the norms and commits the batch: count, buffer flip, error history, stop flag.
__global__ void __launch_bounds__(1024) est_step_kernel(double *state, int p, int max_batches,
                                                         const double *partials, size_t rank_stride,
                                                         int nranks, double *zsq, unsigned int *ticket) {
  extern __shared__ double smem[];
  __shared__ bool s_last;
  double *sortbuf = smem;            // kDraws
  double *mrun = smem + kDraws;      // (nranks+1) x p running means
  double *nrun = mrun + (size_t)(nranks + 1) * p;  // nranks+1 running counts
  StateView st = view_state(state, p, max_batches);
  if (st.hdr[H_STOP] != 0.0) return;
  const int cur = (int)st.hdr[H_CUR], nxt = cur ^ 1;
  const bool est = st.hdr[H_EST] != 0.0;
  const int tid = threadIdx.x;
  const int f = blockIdx.x;

  // running means / counts after merging 0..r partials (every CTA recomputes them)
  if (tid == 0) {
    double n = st.hdr[H_N];
    nrun[0] = n;
    for (int r = 0; r < nranks; ++r) {
      n += partials[(size_t)r * rank_stride];
      nrun[r + 1] = n;
    }
  }
  for (int j = tid; j < p; j += blockDim.x) mrun[j] = st.mean[cur][j];
  __syncthreads();
  for (int r = 0; r < nranks; ++r) {
    const PartView pv = view_part(partials + (size_t)r * rank_stride, p);
    const double n1 = nrun[r], n2 = pv.hdr[0], nn = nrun[r + 1];
    for (int j = tid; j < p; j += blockDim.x) {
      const double m1 = mrun[(size_t)r * p + j];
      mrun[(size_t)(r + 1) * p + j] = (nn > 0.0) ? (n1 / nn) * m1 + (n2 / nn) * pv.mean[j] : m1;
    }
  }
  __syncthreads();
  const double ntot = nrun[nranks];
  const double scale = 1.0 / sqrt(ntot * (ntot - 1.0));

  // ---- covariance row f (biased), Chan merge rank by rank
  for (int j = tid; j < p; j += blockDim.x) {
    double c = st.cov[cur][(size_t)f * p + j];
    for (int r = 0; r < nranks; ++r) {
      const PartView pv = view_part(partials + (size_t)r * rank_stride, p);
      const double n1 = nrun[r], n2 = pv.hdr[0], nn = nrun[r + 1];
      if (n2 > 0.0) {
        const double df = mrun[(size_t)r * p + f] - pv.mean[f];
        const double dj = mrun[(size_t)r * p + j] - pv.mean[j];
        c = (n1 / nn) * c + pv.m2[(size_t)f * p + j] / nn + (n1 / nn) * (n2 / nn) * df * dj;
      }
    }
    st.cov[nxt][(size_t)f * p + j] = c;
  }
  if (tid == 0) st.mean[nxt][f] = mrun[(size_t)nranks * p + f];
  if (est) {
    // ---- S column f and G, re-centred on the moving mean
    double s = st.S[cur][(size_t)f * kDraws + tid];
    double gr = st.G[cur][tid];
    for (int r = 0; r < nranks; ++r) {
      const PartView pv = view_part(partials + (size_t)r * rank_stride, p);
      if (pv.hdr[0] > 0.0) {
        const double m_old = mrun[(size_t)r * p + f], m_new = mrun[(size_t)(r + 1) * p + f];
        const double g2 = pv.G[tid];
        s += (m_old - m_new) * gr + pv.S[(size_t)f * kDraws + tid] + (pv.mean[f] - m_new) * g2;
        gr += g2;
      }
    }
    st.S[nxt][(size_t)f * kDraws + tid] = s;
    if (f == 0) st.G[nxt][tid] = gr;
    const double z = s * scale;
    zsq[(size_t)f * kDraws + tid] = z * z;
    sortbuf[tid] = fabs(z);
    bitonic_sort_1024(sortbuf, tid);
    if (tid == 0) st.feat_err_tmp[f] = quantile95_sorted(sortbuf);
  }
  // ---- last CTA: overall error + commit
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double overall = 0.0;
  if (est) {
    double ss = 0.0;
    for (int ff = 0; ff < p; ++ff) ss += __ldcg(zsq + (size_t)ff * kDraws + tid);
    sortbuf[tid] = sqrt(ss);
    bitonic_sort_1024(sortbuf, tid);
    overall = quantile95_sorted(sortbuf);
    for (int j = tid; j < p; j += blockDim.x) st.feat_err[j] = __ldcg(st.feat_err_tmp + j);
  }
  __syncthreads();
  if (tid == 0) {
    *ticket = 0;
    st.hdr[H_N] = ntot;
    st.hdr[H_CUR] = (double)nxt;
    if (est) {
      st.hdr[H_OVERALL] = overall;
      const int nh = (int)st.hdr[H_NHIST];
      if (nh < (int)st.hdr[H_MAXH]) st.err_hist[nh] = overall;
      st.hdr[H_NHIST] = (double)(nh + 1);
      if (overall < st.hdr[H_TOL]) st.hdr[H_STOP] = 1.0;
    }
  }
}

__global__ void est_init_kernel(double *state, size_t total, int max_batches, double tol, int est) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  double v = 0.0;
  if (i == H_TOL) v = tol;
  if (i == H_EST) v = est ? 1.0 : 0.0;
  if (i == H_MAXH) v = (double)max_batches;
  state[i] = v;
}

__global__ void est_read_kernel(const double *state, int p, int max_batches, double *summary4,
                                double *mean, double *feat_err, double *err_hist, double *cov) {
  StateView st = view_state(const_cast<double *>(state), p, max_batches);
  const int cur = (int)st.hdr[H_CUR];
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    summary4[0] = st.hdr[H_N];
    summary4[1] = st.hdr[H_STOP];
    summary4[2] = st.hdr[H_NHIST];
    summary4[3] = st.hdr[H_OVERALL];
  }
  if (i < (size_t)p) {
    mean[i] = st.mean[cur][i];
    feat_err[i] = st.feat_err[i];
  }
  if (i < (size_t)max_batches) err_hist[i] = st.err_hist[i];
  if (cov && i < (size_t)p * p) cov[i] = st.cov[cur][i];
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
        if (fabs(ga) > 1e-15 * sqrt(al * be) && ga != 0.0) {
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

extern "C" size_t lsspa_estimator_state_bytes(int p, int max_batches) {
  if (p < 1 || max_batches < 0) return 0;
  return state_doubles(p, max_batches) * sizeof(double);
}

extern "C" int64_t lsspa_estimator_partial_doubles(int p) {
  return p < 1 ? 0 : (int64_t)partial_doubles(p);
}

extern "C" int lsspa_estimator_init(void *state, int p, int max_batches, double tolerance, int estimate_errors,
                                    void *stream) {
  if (!state || p < 1 || max_batches < 0) return LSSPA_E_BADARG;
  const size_t total = state_doubles(p, max_batches);
  est_init_kernel<<<(unsigned)ceil_div((int64_t)total, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<double *>(state), total, max_batches, tolerance, estimate_errors);
  LSSPA_LAUNCH_CHECK();
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

extern "C" int lsspa_estimator_update(void *state, int p, int max_batches, const double *partials,
                                      int nbatch, int nranks, int estimate_errors, void *stream) {
  if (!state || !partials || p < 1 || nbatch < 0 || nranks < 1) return LSSPA_E_BADARG;
  cudaStream_t st = as_stream(stream);
  const size_t pstride = partial_doubles(p);
  // gathered layout: partials[rank][batch][block]
  const size_t rank_stride = (size_t)nbatch * pstride;
  const size_t smem = ((size_t)kDraws + (size_t)(nranks + 1) * p + (nranks + 1) + 8) * sizeof(double);
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(est_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  (void)estimate_errors;
  StateView sv = view_state(reinterpret_cast<double *>(state), p, max_batches);
  for (int b = 0; b < nbatch; ++b) {
    est_step_kernel<<<p, kDraws, smem, st>>>(reinterpret_cast<double *>(state), p, max_batches,
                                             partials + (size_t)b * pstride, rank_stride, nranks, sv.zsq,
                                             reinterpret_cast<unsigned int *>(sv.ticket));
    LSSPA_LAUNCH_CHECK();
  }
  return LSSPA_OK;
}

extern "C" int lsspa_estimator_read(const void *state, int p, int max_batches, double *summary4, double *mean,
                                    double *feat_err, double *err_hist, double *cov_or_null, void *stream) {
  if (!state || !summary4 || !mean || !feat_err || p < 1) return LSSPA_E_BADARG;
  if (max_batches > 0 && !err_hist) return LSSPA_E_BADARG;
  size_t n = (size_t)p;
  if ((size_t)max_batches > n) n = (size_t)max_batches;
  if (cov_or_null && (size_t)p * p > n) n = (size_t)p * p;
  est_read_kernel<<<(unsigned)ceil_div((int64_t)n, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const double *>(state), p, max_batches, summary4, mean, feat_err, err_hist, cov_or_null);
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
