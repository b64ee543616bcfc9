// Device permutation sources, bit-exact with the host generators the reference uses.
//
//   exact          itertools.permutations(range(p))            reference ls_spa/ls_spa.py:171
//   pcg64          np.random.default_rng(seed).permutation(p)  reference ls_spa/ls_spa.py:168,175
//   sobol_argsort  np.argsort(Sobol(p).random(n), axis=1)      reference experiments/ground_truth_medium.py:70-71
//   permutohedron  permutohedron_samples(MultivariateNormalQMC(...))  reference experiments/ground_truth_medium.py:56-67
//
// The arithmetic restated here is third-party (numpy 2.3.5 PCG64 / Generator.shuffle,
// scipy 1.18.1 qmc.Sobol and MultivariateNormalQMC); the reference only calls it.

#include <stdlib.h>

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

// ---- parallel resolution of the masked rejection -------------------------------------------
// Which draws are accepted depends on the Fisher-Yates step m = (acceptances so far) mod (p - 1):
// a finite-state machine over the draw stream.  It is evaluated in parallel the standard way:
//   1. every chunk of kChunk draws is simulated from EVERY start state (block = chunk, thread =
//      state): acc_tab[chunk][m0] = acceptances inside the chunk;
//   2. the same per group of kGroup chunks (composition of the chunk maps), then one short serial
//      walk over the groups gives the global acceptance count at every group start;
//   3. every chunk is replayed from its now known start state (one thread per chunk) and writes
//      the accepted values; the thread that produces the last needed acceptance records how many
//      draws the stream consumed.
constexpr int kChunk = 256;
constexpr int kGroup = 64;

__device__ __forceinline__ bool fy_accept(uint32_t d, int steps, int m, uint32_t &val) {
  const uint32_t i = (uint32_t)(steps - m);  // Fisher-Yates index p-1 .. 1
  val = d & (0xffffffffu >> __clz(i));
  return val <= i;
}

__global__ void __launch_bounds__(128) pcg64_fsm_kernel(int steps, const uint32_t *raw, int64_t ndraws,
                                                        uint16_t *acc_tab) {
  __shared__ uint32_t d[kChunk];
  const int64_t c = blockIdx.x, n0 = c * kChunk;
  const int nvalid = (int)((ndraws - n0 < kChunk) ? ndraws - n0 : kChunk);
  for (int i = threadIdx.x; i < nvalid; i += blockDim.x) d[i] = raw[n0 + i];
  __syncthreads();
  for (int m0 = threadIdx.x; m0 < steps; m0 += blockDim.x) {
    int m = m0, a = 0;
    for (int n = 0; n < nvalid; ++n) {
      uint32_t v;
      if (fy_accept(d[n], steps, m, v)) {
        ++a;
        m = (m + 1 == steps) ? 0 : m + 1;
      }
    }
    acc_tab[c * steps + m0] = (uint16_t)a;
  }
}

__global__ void __launch_bounds__(128) pcg64_group_kernel(int steps, int64_t nchunks, const uint16_t *acc_tab,
                                                          uint32_t *grp_tab) {
  const int64_t g = blockIdx.x, c0 = g * kGroup;
  const int64_t c1 = (c0 + kGroup < nchunks) ? c0 + kGroup : nchunks;
  for (int m0 = threadIdx.x; m0 < steps; m0 += blockDim.x) {
    int m = m0;
    uint32_t t = 0;
    for (int64_t c = c0; c < c1; ++c) {
      const int a = acc_tab[c * steps + m];
      t += (uint32_t)a;
      m = (m + a) % steps;
    }
    grp_tab[g * steps + m0] = t;
  }
}

// Tg[g] = acceptances before group g (Tg[ngroups] = all of them); consumed = -1 (not yet known)
__global__ void pcg64_prefix_kernel(int steps, int64_t ngroups, const uint32_t *grp_tab, int64_t *Tg,
                                    int64_t *consumed) {
  int64_t t = 0;
  for (int64_t g = 0; g < ngroups; ++g) {
    Tg[g] = t;
    t += grp_tab[g * steps + (int)(t % steps)];
  }
  Tg[ngroups] = t;
  *consumed = -1;
}

__global__ void __launch_bounds__(kGroup) pcg64_emit_kernel(int steps, int64_t total, const uint32_t *raw,
                                                            int64_t ndraws, int64_t nchunks, const uint16_t *acc_tab,
                                                            const int64_t *Tg, int32_t *accepted, int64_t *consumed) {
  __shared__ int64_t tc[kGroup];
  const int64_t g = blockIdx.x, c0 = g * kGroup;
  if (threadIdx.x == 0) {
    int64_t t = Tg[g];
    for (int k = 0; k < kGroup; ++k) {
      tc[k] = t;
      if (c0 + k < nchunks) t += acc_tab[(c0 + k) * steps + (int)(t % steps)];
    }
  }
  __syncthreads();
  const int64_t c = c0 + threadIdx.x;
  if (c >= nchunks) return;
  int64_t t = tc[threadIdx.x];
  if (t >= total) return;
  int m = (int)(t % steps);
  const int64_t n0 = c * kChunk;
  const int nvalid = (int)((ndraws - n0 < kChunk) ? ndraws - n0 : kChunk);
  for (int n = 0; n < nvalid; ++n) {
    uint32_t v;
    if (fy_accept(raw[n0 + n], steps, m, v)) {
      accepted[t] = (int32_t)v;
      ++t;
      m = (m + 1 == steps) ? 0 : m + 1;
      if (t == total) {
        *consumed = n0 + n + 1;
        break;
      }
    }
  }
}

// advance the generator by the draws the stream consumed (the buffered half word included)
__global__ void pcg64_finish_kernel(int64_t ndraws, const int64_t *consumed_in, uint64_t *gen_state,
                                    int *status_flag) {
  int64_t consumed = *consumed_in;
  if (consumed < 0) {  // raw budget exhausted (caller sized it too small)
    *status_flag = 1;
    consumed = ndraws;
  }
  const u128 s0 = ((u128)gen_state[0] << 64) | gen_state[1];
  const u128 inc = ((u128)gen_state[2] << 64) | gen_state[3];
  const int has = gen_state[4] != 0;
  const int64_t from_outputs = consumed - ((has && consumed > 0) ? 1 : 0);
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

// workspace: raw draws | accepted values | chunk tables | group tables | group prefix | consumed
struct PcgLayout {
  int64_t budget, nchunks, ngroups, steps;
  size_t off_acc, off_tab, off_grp, off_tg, off_cons, bytes;
};
static PcgLayout pcg64_layout(int p, int64_t count) {
  PcgLayout L;
  L.steps = p > 1 ? p - 1 : 1;
  L.budget = pcg64_raw_budget(p, count);
  L.nchunks = ceil_div(L.budget, kChunk);
  L.ngroups = ceil_div(L.nchunks, kGroup);
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t o = up((size_t)(L.budget + 2) * sizeof(uint32_t));
  L.off_acc = o;
  o = up(o + (size_t)count * L.steps * sizeof(int32_t));
  L.off_tab = o;
  o = up(o + (size_t)L.nchunks * L.steps * sizeof(uint16_t));
  L.off_grp = o;
  o = up(o + (size_t)L.ngroups * L.steps * sizeof(uint32_t));
  L.off_tg = o;
  o = up(o + (size_t)(L.ngroups + 1) * sizeof(int64_t));
  L.off_cons = o;
  L.bytes = o + 256;
  return L;
}

extern "C" size_t lsspa_perms_pcg64_workspace_bytes(int p, int64_t count) {
  if (p < 1 || count < 1) return 0;
  return pcg64_layout(p, count).bytes;
}

extern "C" int lsspa_perms_pcg64(int p, uint64_t *gen_state, int64_t count, int32_t *perms_out,
                                 void *workspace, size_t workspace_bytes, int *status_flag,
                                 void *stream) {
  if (p < 1 || count < 0 || !gen_state || !perms_out || !status_flag) return LSSPA_E_BADARG;
  if (count == 0) return LSSPA_OK;
  cudaStream_t st = as_stream(stream);
  const PcgLayout L = pcg64_layout(p, count);
  if (!workspace || workspace_bytes < L.bytes) return LSSPA_E_WORKSPACE;
  char *ws = reinterpret_cast<char *>(workspace);
  const int64_t budget = L.budget;  // even number of 32-bit draws
  uint32_t *raw = reinterpret_cast<uint32_t *>(ws);
  int32_t *accepted = reinterpret_cast<int32_t *>(ws + L.off_acc);
  const int nt = 128;
  if (p > 1) {
    const int64_t nout = budget / 2;
    const int64_t nthreads = ceil_div(nout, kRawPerThread);
    pcg64_raw_kernel<<<(unsigned)ceil_div(nthreads, nt), nt, 0, st>>>(gen_state, nout, raw);
    LSSPA_LAUNCH_CHECK();
    // `budget` draws come from outputs; a buffered uinteger (if any) adds one more in front
    static const bool serial = [] {
      const char *e = getenv("LSSPA_PCG_SCAN");
      return e && e[0] == 's';
    }();
    if (serial) {  // the single-warp walk (A/B timing, debugging)
      pcg64_scan_kernel<<<1, 32, 0, st>>>(p, count, raw, budget, accepted, gen_state, status_flag);
      LSSPA_LAUNCH_CHECK();
    } else {
      const int steps = (int)L.steps;
      uint16_t *tab = reinterpret_cast<uint16_t *>(ws + L.off_tab);
      uint32_t *grp = reinterpret_cast<uint32_t *>(ws + L.off_grp);
      int64_t *Tg = reinterpret_cast<int64_t *>(ws + L.off_tg);
      int64_t *cons = reinterpret_cast<int64_t *>(ws + L.off_cons);
      pcg64_fsm_kernel<<<(unsigned)L.nchunks, 128, 0, st>>>(steps, raw, budget, tab);
      LSSPA_LAUNCH_CHECK();
      pcg64_group_kernel<<<(unsigned)L.ngroups, 128, 0, st>>>(steps, L.nchunks, tab, grp);
      LSSPA_LAUNCH_CHECK();
      pcg64_prefix_kernel<<<1, 1, 0, st>>>(steps, L.ngroups, grp, Tg, cons);
      LSSPA_LAUNCH_CHECK();
      pcg64_emit_kernel<<<(unsigned)L.ngroups, kGroup, 0, st>>>(steps, count * (int64_t)steps, raw, budget, L.nchunks,
                                                                tab, Tg, accepted, cons);
      LSSPA_LAUNCH_CHECK();
      pcg64_finish_kernel<<<1, 1, 0, st>>>(budget, cons, gen_state, status_flag);
      LSSPA_LAUNCH_CHECK();
    }
  }
  pcg64_shuffle_kernel<<<(unsigned)ceil_div(count, nt), nt, 0, st>>>(p, count, accepted, perms_out);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

namespace lsspa {
// One warp per row: every entry in [0, p) and every value hit exactly once (bitmap in shared memory).
// bad_flag (device int, never cleared here) receives 1 + the first offending row seen by some warp.
__global__ void __launch_bounds__(256) perms_validate_kernel(int p, const int32_t *perms, int64_t count, int *bad_flag) {
  extern __shared__ unsigned int bm_all[];
  const int words = (p + 31) / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned int *bm = bm_all + (size_t)warp * words;
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < count; row += (int64_t)gridDim.x * 8) {
    for (int w = lane; w < words; w += 32) bm[w] = 0u;
    __syncwarp();
    bool bad = false;
    for (int k = lane; k < p; k += 32) {
      const int v = perms[row * p + k];
      if (v < 0 || v >= p) bad = true;
      else if (atomicOr(&bm[v >> 5], 1u << (v & 31)) & (1u << (v & 31))) bad = true;   // seen before
    }
    __syncwarp();
    if (__any_sync(kFull, bad) && lane == 0) atomicCAS(bad_flag, 0, (int)(row < 0x7ffffffe ? row + 1 : 0x7fffffff));
    __syncwarp();
  }
}
}  // namespace lsspa

extern "C" int lsspa_perms_validate(int p, const int32_t *perms, int64_t count, int *bad_flag, void *stream) {
  if (p < 1 || count < 0 || !bad_flag || (count > 0 && !perms)) return LSSPA_E_BADARG;
  if (count == 0) return LSSPA_OK;
  const size_t smem = (size_t)8 * ((p + 31) / 32) * sizeof(unsigned int);
  if (smem > 48 * 1024) return LSSPA_E_UNSUPPORTED;
  int64_t grid = lsspa::ceil_div(count, 8);
  if (grid > 148 * 8) grid = 148 * 8;
  lsspa::perms_validate_kernel<<<(unsigned)grid, 256, smem, lsspa::as_stream(stream)>>>(p, perms, count, bad_flag);
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
