// Device permutation sources, bit-exact with the host generators the reference uses.
//
//   exact          itertools.permutations(range(p))            reference ls_spa/ls_spa.py:171
//   pcg64          np.random.default_rng(seed).permutation(p)  reference ls_spa/ls_spa.py:168,175
//   sobol_argsort  np.argsort(Sobol(p).random(n), axis=1)      reference experiments/ground_truth_medium.py:70-71
//   permutohedron  permutohedron_samples(MultivariateNormalQMC(...))  reference experiments/ground_truth_medium.py:56-67
//
// The arithmetic restated here is third-party (numpy 2.3.5 PCG64 / Generator.shuffle,
// scipy 1.18.1 qmc.Sobol and MultivariateNormalQMC); the reference only calls it.

#include "common.cuh"

namespace lsspa {

typedef unsigned __int128 u128;

// ---------------------------------------------------------------- exact (lexicographic)
__global__ void perms_exact_kernel(int p, uint64_t first, int64_t count, int32_t *out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  uint64_t r = first + (uint64_t)idx;
  const int m = p < 20 ? p : 20;  // 20! < 2^64 <= 21!: only the last 20 positions can move
  const int base = p - m;
  unsigned char pool[20];
  unsigned char dig[20];
  for (int i = 0; i < m; ++i) pool[i] = (unsigned char)i;
  for (int i = 1; i <= m; ++i) {
    dig[m - i] = (unsigned char)(r % (uint64_t)i);
    r /= (uint64_t)i;
  }
  int32_t *row = out + idx * p;
  for (int i = 0; i < base; ++i) row[i] = i;
  for (int pos = 0; pos < m; ++pos) {
    const int d = dig[pos];
    row[base + pos] = base + pool[d];
    for (int i = d; i + 1 < m - pos; ++i) pool[i] = pool[i + 1];
  }
}

// ---------------------------------------------------------------- PCG64 (numpy Generator)
// 128-bit LCG, multiplier below, output XSL-RR 128/64 taken AFTER the step.
__device__ __forceinline__ u128 pcg_mult() {
  return ((u128)0x2360ED051FC65DA4ULL << 64) | (u128)0x4385DF649FCCF645ULL;
}
__device__ __forceinline__ uint64_t pcg_output(u128 s) {
  const uint64_t hi = (uint64_t)(s >> 64), lo = (uint64_t)s;
  const unsigned rot = (unsigned)(hi >> 58);
  const uint64_t x = hi ^ lo;
  return (x >> rot) | (x << ((64u - rot) & 63u));
}
// state after `delta` steps (Brown's O(log n) LCG jump, as numpy's pcg_advance_lcg_128)
__device__ u128 pcg_advance(u128 state, u128 inc, uint64_t delta) {
  u128 acc_mult = 1, acc_plus = 0, cur_mult = pcg_mult(), cur_plus = inc;
  while (delta > 0) {
    if (delta & 1) {
      acc_mult *= cur_mult;
      acc_plus = acc_plus * cur_mult + cur_plus;
    }
    cur_plus = (cur_mult + 1) * cur_plus;
    cur_mult *= cur_mult;
    delta >>= 1;
  }
  return acc_mult * state + acc_plus;
}

constexpr int kRawPerThread = 128;  // 64-bit outputs generated sequentially by one thread

// raw[0] (optional) = buffered uinteger; then lo32, hi32 of outputs 1, 2, ... of the stream
__global__ void pcg64_raw_kernel(const uint64_t *gen_state, int64_t nout, uint32_t *raw) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t o0 = t * kRawPerThread;
  if (o0 >= nout) return;
  const u128 s0 = ((u128)gen_state[0] << 64) | gen_state[1];
  const u128 inc = ((u128)gen_state[2] << 64) | gen_state[3];
  const int has = gen_state[4] != 0;
  if (t == 0 && has) raw[0] = (uint32_t)gen_state[5];
  uint32_t *dst = raw + has;
  u128 s = pcg_advance(s0, inc, (uint64_t)o0);
  const u128 mult = pcg_mult();
  const int64_t o1 = (o0 + kRawPerThread < nout) ? o0 + kRawPerThread : nout;
  for (int64_t o = o0; o < o1; ++o) {
    s = s * mult + inc;
    const uint64_t v = pcg_output(s);
    dst[2 * o] = (uint32_t)v;
    dst[2 * o + 1] = (uint32_t)(v >> 32);
  }
}

// One warp walks the raw draws and resolves numpy's masked rejection (random_interval):
// for i = p-1 .. 1 redraw while (u32 & mask(i)) > i.  Lane l speculates how many of the
// lanes before it accept; iterating the ballot fixes at least one more lane per round.
__global__ void pcg64_scan_kernel(int p, int64_t count, const uint32_t *raw, int64_t ndraws,
                                  int32_t *accepted, uint64_t *gen_state, int *status_flag) {
  const int lane = threadIdx.x;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int steps = p - 1;
  const int64_t total = count * (int64_t)steps;
  int64_t t = 0, pos = 0, consumed = 0;
  int tmod = 0;
  // the walk is serial, so the loads are not: keep the next kAhead windows of draws in registers
  constexpr int kAhead = 8;
  uint32_t win[kAhead];
#pragma unroll
  for (int a = 0; a < kAhead; ++a) win[a] = (32 * a + lane < ndraws) ? raw[32 * a + lane] : 0u;
  while (t < total && pos < ndraws) {
    const bool in_range = pos + lane < ndraws;
    const uint32_t d = win[0];
#pragma unroll
    for (int a = 0; a + 1 < kAhead; ++a) win[a] = win[a + 1];
    {
      const int64_t nxt = pos + 32 * kAhead + lane;
      win[kAhead - 1] = (nxt < ndraws) ? raw[nxt] : 0u;
    }
    unsigned accmask = kFull;
    uint32_t val = 0;
    int64_t tl = 0;
    bool acc = false;
    for (int it = 0; it < 33; ++it) {
      const int prior = __popc(accmask & lt_mask);
      tl = t + prior;
      int m = tmod + prior;  // < steps + 32
      if (steps >= 32) {
        if (m >= steps) m -= steps;
      } else {
        m %= steps;
      }
      const uint32_t i = (uint32_t)(steps - m);  // Fisher-Yates index p-1 .. 1
      const uint32_t msk = 0xffffffffu >> __clz(i);
      val = d & msk;
      acc = in_range && (tl < total) && (val <= i);
      const unsigned nm = __ballot_sync(kFull, acc);
      if (nm == accmask) break;
      accmask = nm;
    }
    if (acc) accepted[tl] = (int32_t)val;
    const int nacc = __popc(accmask);
    if (t + nacc >= total) {
      const unsigned last = __ballot_sync(kFull, acc && (tl + 1 == total));
      consumed = pos + (31 - __clz(last)) + 1;
    } else {
      const int64_t rem = ndraws - pos;
      consumed = pos + (rem < 32 ? rem : 32);
    }
    t += nacc;
    tmod = (int)((tmod + nacc) % steps);
    pos += 32;
  }
  if (lane == 0) {
    if (t < total) *status_flag = 1;  // raw budget exhausted (caller sized it too small)
    const u128 s0 = ((u128)gen_state[0] << 64) | gen_state[1];
    const u128 inc = ((u128)gen_state[2] << 64) | gen_state[3];
    const int has = gen_state[4] != 0;
    int64_t from_outputs = consumed - ((has && consumed > 0) ? 1 : 0);
    if (consumed > 0) {
      const uint64_t nout = (uint64_t)((from_outputs + 1) / 2);
      const u128 s = pcg_advance(s0, inc, nout);
      gen_state[0] = (uint64_t)(s >> 64);
      gen_state[1] = (uint64_t)s;
      if (from_outputs & 1) {
        gen_state[4] = 1;
        gen_state[5] = pcg_output(s) >> 32;
      } else {
        gen_state[4] = 0;
        gen_state[5] = 0;
      }
    }
  }
}

// thread per permutation: a = arange(p); for i = p-1..1: swap(a[i], a[j_i])
__global__ void pcg64_shuffle_kernel(int p, int64_t count, const int32_t *accepted, int32_t *out) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= count) return;
  int32_t *a = out + n * p;
  const int32_t *j = accepted + n * (int64_t)(p - 1);
  for (int i = 0; i < p; ++i) a[i] = i;
  for (int i = p - 1, s = 0; i >= 1; --i, ++s) {
    const int jj = j[s];
    const int32_t tmp = a[i];
    a[i] = a[jj];
    a[jj] = tmp;
  }
}

static int64_t pcg64_raw_budget(int p, int64_t count) {
  // expected draws per permutation = sum_i (mask(i)+1)/(i+1); 15 % + 64k head-room
  double e = 0.0;
  for (int i = 1; i < p; ++i) {
    uint32_t m = 0xffffffffu >> __builtin_clz((unsigned)i);
    e += ((double)m + 1.0) / ((double)i + 1.0);
  }
  double tot = e * (double)count * 1.15 + 65536.0;
  int64_t n = (int64_t)tot;
  return (n + 1) & ~(int64_t)1;  // even
}

// ---------------------------------------------------------------- scrambled Sobol' points
// scipy: point k = shift ^ XOR_{b in bits(gray(k))} sv[:, b], value = int * 2^-bits
__device__ __forceinline__ uint32_t sobol_coord(const uint32_t *sv_row, uint32_t shift, int bits,
                                                uint64_t gray) {
  uint32_t x = shift;
  for (int b = 0; b < bits && gray; ++b, gray >>= 1)
    if (gray & 1) x ^= sv_row[b];
  return x;
}

// argsort by counting: rank_j = #{l : (key_l, l) < (key_j, j)}  (stable; numpy's argsort
// on distinct keys gives the same permutation, tie order is unpinned in the reference)
template <typename Key>
__device__ __forceinline__ void rank_scatter(const Key *keys, int p, int32_t *row) {
  for (int j = threadIdx.x; j < p; j += blockDim.x) {
    const Key kj = keys[j];
    int r = 0;
    for (int l = 0; l < p; ++l) {
      const Key kl = keys[l];
      r += (kl < kj) || (kl == kj && l < j);
    }
    row[r] = j;
  }
}

__global__ void sobol_argsort_kernel(int p, const uint32_t *sv, const uint32_t *shift, int bits,
                                     uint64_t first, int64_t count, int32_t *out) {
  extern __shared__ uint32_t keys_u[];
  for (int64_t n = blockIdx.x; n < count; n += gridDim.x) {
    const uint64_t k = first + (uint64_t)n;
    const uint64_t gray = k ^ (k >> 1);
    __syncthreads();
    for (int j = threadIdx.x; j < p; j += blockDim.x)
      keys_u[j] = sobol_coord(sv + (size_t)j * bits, shift[j], bits, gray);
    __syncthreads();
    rank_scatter<uint32_t>(keys_u, p, out + n * p);
  }
}

// MultivariateNormalQMC(inv_transform=False): Box-Muller on consecutive Sobol' coordinate
// pairs, then projection on U (rows r = 0..p-2: 1/n_r in columns 0..r, -(r+1)/n_r in
// column r+1, n_r = sqrt((r+1)(r+2))), i.e. proj_j = sum_{r>=j} z_r/n_r - j z_{j-1}/n_{j-1}.
// The per-row normalisation of z (reference :59) is a positive scalar and cannot change
// the argsort, so it is not applied.
__global__ void permutohedron_kernel(int p, const uint32_t *sv, const uint32_t *shift, int bits,
                                     uint64_t first, int64_t count, int32_t *out) {
  extern __shared__ double zbuf[];  // w[p] (scaled normals), proj[p]
  double *w = zbuf;
  double *proj = zbuf + p + 1;
  const int dim = p - 1;
  const int npairs = (dim + 1) / 2;
  const double scale = ldexp(1.0, -bits);
  for (int64_t n = blockIdx.x; n < count; n += gridDim.x) {
    const uint64_t k = first + (uint64_t)n;
    const uint64_t gray = k ^ (k >> 1);
    __syncthreads();
    for (int t = threadIdx.x; t < npairs; t += blockDim.x) {
      const double u0 = scale * (double)sobol_coord(sv + (size_t)(2 * t) * bits, shift[2 * t], bits, gray);
      const double u1 =
          scale * (double)sobol_coord(sv + (size_t)(2 * t + 1) * bits, shift[2 * t + 1], bits, gray);
      const double rad = sqrt(-2.0 * log(u0));
      const double th = (2.0 * 3.141592653589793) * u1;
      const int r0 = 2 * t, r1 = 2 * t + 1;
      w[r0] = rad * cos(th) / sqrt((double)(r0 + 1) * (double)(r0 + 2));
      if (r1 < dim) w[r1] = rad * sin(th) / sqrt((double)(r1 + 1) * (double)(r1 + 2));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double suffix = 0.0;
      proj[p - 1] = 0.0;
      for (int j = dim - 1; j >= 0; --j) {
        suffix += w[j];
        proj[j] = suffix;
      }
    }
    __syncthreads();
    for (int j = threadIdx.x + 1; j < p; j += blockDim.x) proj[j] -= (double)j * w[j - 1];
    __syncthreads();
    rank_scatter<double>(proj, p, out + n * p);
  }
}

static int block_for(int p) {
  int b = ((p + 31) / 32) * 32;
  if (b > 1024) b = 1024;
  return b;
}

static int grid_for(int64_t count) {
  const DeviceInfo &d = device_info();
  int64_t cap = (int64_t)(d.sm_count > 0 ? d.sm_count : 148) * 32;
  return (int)(count < cap ? count : cap);
}

}  // namespace lsspa

using namespace lsspa;

extern "C" int lsspa_perms_exact(int p, uint64_t first_rank, int64_t count, int32_t *perms_out,
                                 void *stream) {
  if (p < 1 || count < 0 || !perms_out) return LSSPA_E_BADARG;
  if (count == 0) return LSSPA_OK;
  const int nt = 128;
  perms_exact_kernel<<<(unsigned)ceil_div(count, nt), nt, 0, as_stream(stream)>>>(p, first_rank, count,
                                                                                 perms_out);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

extern "C" size_t lsspa_perms_pcg64_workspace_bytes(int p, int64_t count) {
  if (p < 1 || count < 1) return 0;
  const int64_t nraw = pcg64_raw_budget(p, count) + 2;
  const int64_t nacc = count * (int64_t)(p > 1 ? p - 1 : 1);
  return (size_t)(nraw + nacc) * sizeof(uint32_t) + 256;
}

extern "C" int lsspa_perms_pcg64(int p, uint64_t *gen_state, int64_t count, int32_t *perms_out,
                                 void *workspace, size_t workspace_bytes, int *status_flag,
                                 void *stream) {
  if (p < 1 || count < 0 || !gen_state || !perms_out || !status_flag) return LSSPA_E_BADARG;
  if (count == 0) return LSSPA_OK;
  cudaStream_t st = as_stream(stream);
  const size_t need = lsspa_perms_pcg64_workspace_bytes(p, count);
  if (!workspace || workspace_bytes < need) return LSSPA_E_WORKSPACE;
  const int64_t budget = pcg64_raw_budget(p, count);  // even number of 32-bit draws
  uint32_t *raw = reinterpret_cast<uint32_t *>(workspace);
  int32_t *accepted = reinterpret_cast<int32_t *>(raw + budget + 2);
  const int nt = 128;
  if (p > 1) {
    const int64_t nout = budget / 2;
    const int64_t nthreads = ceil_div(nout, kRawPerThread);
    pcg64_raw_kernel<<<(unsigned)ceil_div(nthreads, nt), nt, 0, st>>>(gen_state, nout, raw);
    LSSPA_LAUNCH_CHECK();
    // `budget` draws come from outputs; a buffered uinteger (if any) adds one more in front
    pcg64_scan_kernel<<<1, 32, 0, st>>>(p, count, raw, budget, accepted, gen_state, status_flag);
    LSSPA_LAUNCH_CHECK();
  }
  pcg64_shuffle_kernel<<<(unsigned)ceil_div(count, nt), nt, 0, st>>>(p, count, accepted, perms_out);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

extern "C" int lsspa_perms_sobol_argsort(int p, const uint32_t *sv, const uint32_t *shift, int bits,
                                         uint64_t first_index, int64_t count, int32_t *perms_out,
                                         void *stream) {
  if (p < 1 || count < 0 || !sv || !shift || !perms_out || bits < 1 || bits > 32) return LSSPA_E_BADARG;
  if (count == 0) return LSSPA_OK;
  sobol_argsort_kernel<<<grid_for(count), block_for(p), (size_t)p * sizeof(uint32_t), as_stream(stream)>>>(
      p, sv, shift, bits, first_index, count, perms_out);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

extern "C" int lsspa_perms_permutohedron(int p, const uint32_t *sv, const uint32_t *shift, int bits,
                                         uint64_t first_index, int64_t count, int32_t *perms_out,
                                         void *stream) {
  if (p < 2 || count < 0 || !sv || !shift || !perms_out || bits < 1 || bits > 32) return LSSPA_E_BADARG;
  if (count == 0) return LSSPA_OK;
  permutohedron_kernel<<<grid_for(count), block_for(p), (size_t)(2 * p + 2) * sizeof(double),
                         as_stream(stream)>>>(p, sv, shift, bits, first_index, count, perms_out);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}
