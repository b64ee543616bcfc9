// CholeskyQR2 reduction of [X | y] to its (p+1) x (p+1) triangular factor, p <= 111.
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

#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)
#include <stdio.h>
#include <stdlib.h>

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
  extern __shared__ __align__(128) unsigned char smem_raw[];
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

// ---------------------------------------------------------------- TMA variant of pass 1
// Same Gram accumulation, but the 32-row chunks arrive through the TMA: a 2-D tensor map over X
// (box = 32 rows x (8 nt + 4) columns: the columns beyond p are OUT OF BOUNDS of the tensor and come back
// as zeros, and the row pitch 8 nt + 4 keeps the fragment loads conflict-free) and a 1-D map over y
// (32 entries), both completing on the stage's mbarrier.  Warp 7 is the producer (one lane issues
// cp.async.bulk.tensor for a ring of kTmaStages chunks), warps 0..6 own the tile-row pairs as before and
// release a stage with one mbarrier arrive each: no __syncthreads in the row loop, loads run up to
// three chunks ahead of the DMMAs (warp 8 is the producer, warps 0..7 consume).  The y column is merged into the fragments of tile column nt - 1
// in registers; the 32 targets of a chunk are one coalesced load per warp, handed out by shuffle.  Needs a
// 16-byte aligned X with an even leading dimension (TMA global strides); otherwise gram_rows_kernel runs.
constexpr int kTmaStages = 4;   // ring depth (3 when four stages of two CTAs no longer fit an SM)

__device__ __forceinline__ unsigned tma_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tma_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_mbar_expect(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tma_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tma_smem(bar)) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint64_t *bar, unsigned parity) {
  unsigned ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(tma_smem(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   tma_smem(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(tma_smem(bar))
               : "memory");
}
// The upper tiles (row-major order: (0,0) .. (0,NT-1), (1,1) ..) are dealt to the eight consumer warps in
// contiguous runs of T/8 tiles, T = NT (NT + 1) / 2: every warp -- and with it every SM sub-partition --
// gets the same number of DMMAs (+-1).  NT and the warp index are template parameters, so the run is a
// compile-time list: accumulators, fragment addresses (base + immediate) and the few slots that touch
// the y column are all static, and fragments shared by several slots of a run are loaded once.
__host__ __device__ constexpr int gt_total(int nt) { return nt * (nt + 1) / 2; }
__host__ __device__ constexpr int gt_first(int nt, int w) { return w * gt_total(nt) / 8; }
__host__ __device__ constexpr int gt_row(int nt, int n) {
  int r = 0;
  while (n >= nt - r) {
    n -= nt - r;
    ++r;
  }
  return r;
}
__host__ __device__ constexpr int gt_col(int nt, int n) {
  int r = 0;
  while (n >= nt - r) {
    n -= nt - r;
    ++r;
  }
  return r + n;
}
constexpr int kGtMaxSlots = 15;   // NT <= 15: ceil(120 / 8)

template <int NT, int W, int S>
__device__ __forceinline__ void gt_step(double (&acc)[kGtMaxSlots][2], const double *row, double yv) {
  constexpr int kCount = gt_first(NT, W + 1) - gt_first(NT, W);
  if constexpr (S < kCount) {
    constexpr int n = gt_first(NT, W) + S;
    constexpr int tr = gt_row(NT, n), tc = gt_col(NT, n);
    double fa = row[8 * tr], fb = row[8 * tc];
    if constexpr (tr == NT - 1) fa += yv;
    if constexpr (tc == NT - 1) fb += yv;
    dmma_g(acc[S][0], acc[S][1], fa, fb);
    gt_step<NT, W, S + 1>(acc, row, yv);
  }
}
template <int NT, int W, int S>
__device__ __forceinline__ void gt_store(const double (&acc)[kGtMaxSlots][2], double *out, int nc, int c, int q) {
  constexpr int kCount = gt_first(NT, W + 1) - gt_first(NT, W);
  if constexpr (S < kCount) {
    constexpr int n = gt_first(NT, W) + S;
    constexpr int tr = gt_row(NT, n), tc = gt_col(NT, n);
    *reinterpret_cast<double2 *>(out + (size_t)(8 * tr + c) * nc + 8 * tc + 2 * q) = make_double2(acc[S][0], acc[S][1]);
    gt_store<NT, W, S + 1>(acc, out, nc, c, q);
  }
}

struct GramTmaCtx {
  const unsigned char *smem;
  size_t stage_bytes;
  uint64_t *full, *empty;
  int nst, nchunks, pitch, p;
  int64_t r_begin, r_end;
  const double *y;
  double *out;
};

template <int NT, int W>
__device__ __forceinline__ void gt_consume(const GramTmaCtx &g, int lane) {
  const int c = lane >> 2, q = lane & 3;
  constexpr int nc = 8 * NT;
  double acc[kGtMaxSlots][2];
#pragma unroll
  for (int t = 0; t < kGtMaxSlots; ++t) acc[t][0] = acc[t][1] = 0.0;
  const bool ycol = (nc - 8 + c) == g.p;          // this lane's column of the last tile column is the y column
  // the 32 targets of a chunk: one coalesced load per warp, fetched a chunk ahead, handed out by shuffle
  auto load_y = [&](int i) {
    const int64_t r = g.r_begin + (int64_t)i * kGR + lane;
    return (i < g.nchunks && r < g.r_end) ? __ldg(g.y + r) : 0.0;
  };
  double ynext = load_y(0);
  for (int i = 0; i < g.nchunks; ++i) {
    const int s = i % g.nst;
    const double ycur = ynext;
    ynext = load_y(i + 1);
    tma_mbar_wait(&g.full[s], (i / g.nst) & 1);
    const double *Z = reinterpret_cast<const double *>(g.smem + (size_t)s * g.stage_bytes);
#pragma unroll 4
    for (int ks = 0; ks < kGR / 4; ++ks) {
      const double *row = Z + (size_t)(4 * ks + q) * g.pitch + c;
      const double ysh = __shfl_sync(kFull, ycur, 4 * ks + q);
      gt_step<NT, W, 0>(acc, row, ycol ? ysh : 0.0);
    }
    __syncwarp();
    if (lane == 0) tma_mbar_arrive(&g.empty[s]);
  }
  gt_store<NT, W, 0>(acc, g.out, nc, c, q);
}

constexpr int kGramTmaThreads = 288;   // eight consumer warps + the producer warp

template <int NT>
__global__ void __launch_bounds__(kGramTmaThreads, 2) gram_tma_kernel(const __grid_constant__ CUtensorMap mapX, GramParams a,
                                                                      int nst) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int nc = 8 * NT, pitch = nc + 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t stage_bytes = ((size_t)kGR * pitch * sizeof(double) + 127) / 128 * 128;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + nst * stage_bytes);
  uint64_t *empty = full + kTmaStages;
  if (tid == 0) {
    for (int s = 0; s < nst; ++s) {
      tma_mbar_init(&full[s], 1);
      tma_mbar_init(&empty[s], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t per = ceil_div(ceil_div(a.nrows, (int64_t)a.nparts), (int64_t)kGR) * kGR;
  const int64_t r_begin = (int64_t)blockIdx.x * per;
  const int64_t r_end = (r_begin + per < a.nrows) ? r_begin + per : a.nrows;
  const int nchunks = r_begin < r_end ? (int)ceil_div(r_end - r_begin, (int64_t)kGR) : 0;
  const unsigned tx = (unsigned)((size_t)kGR * pitch * sizeof(double));
  if (warp == 8) {
    if (lane == 0) {
      for (int i = 0; i < nchunks; ++i) {
        const int s = i % nst;
        if (i >= nst) tma_mbar_wait(&empty[s], ((i / nst) - 1) & 1);
        tma_mbar_expect(&full[s], tx);
        tma_load_2d(smem_raw + (size_t)s * stage_bytes, &mapX, 0, (int)(r_begin + (int64_t)i * kGR), &full[s]);
      }
    }
    return;
  }
  GramTmaCtx g;
  g.smem = smem_raw;
  g.stage_bytes = stage_bytes;
  g.full = full;
  g.empty = empty;
  g.nst = nst;
  g.nchunks = nchunks;
  g.pitch = pitch;
  g.p = a.p;
  g.r_begin = r_begin;
  g.r_end = r_end;
  g.y = a.y;
  g.out = a.parts + (size_t)blockIdx.x * nc * nc;
  switch (warp) {
    case 0: gt_consume<NT, 0>(g, lane); break;
    case 1: gt_consume<NT, 1>(g, lane); break;
    case 2: gt_consume<NT, 2>(g, lane); break;
    case 3: gt_consume<NT, 3>(g, lane); break;
    case 4: gt_consume<NT, 4>(g, lane); break;
    case 5: gt_consume<NT, 5>(g, lane); break;
    case 6: gt_consume<NT, 6>(g, lane); break;
    default: gt_consume<NT, 7>(g, lane); break;
  }
}

typedef CUresult (*TmaEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmaEncodeFn tma_encoder() {
  static TmaEncodeFn fn = [] {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return reinterpret_cast<TmaEncodeFn>(ptr);
  }();
  return fn;
}

// the upper tiles are written by the consumer warps only: the sum kernel must not read the lower ones
template <int NT>
static int launch_gram_tma_nt(const CUtensorMap &mx, const GramParams &a, cudaStream_t st) {
  constexpr int pitch = 8 * NT + 4;
  const size_t stage_bytes = ((size_t)kGR * pitch * sizeof(double) + 127) / 128 * 128;
  const int nst = (2 * (kTmaStages * stage_bytes + 1200) <= (size_t)227 * 1024) ? kTmaStages : 3;
  const size_t smem = nst * stage_bytes + 2 * kTmaStages * sizeof(uint64_t);
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(gram_tma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gram_tma_kernel<NT><<<a.nparts, kGramTmaThreads, smem, st>>>(mx, a, nst);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

static int launch_gram_tma(const GramParams &a, cudaStream_t st, bool &used) {
  used = false;
  static const bool off = [] { const char *e = getenv("LSSPA_GRAM_TMA"); return e && e[0] == '0'; }();
  TmaEncodeFn enc = tma_encoder();
  if (off || !enc || a.Rinv != nullptr || a.nt > 15) return LSSPA_OK;
  if ((reinterpret_cast<uintptr_t>(a.X) & 15) || (a.ldx & 1) || a.nrows >= (int64_t)1 << 31) return LSSPA_OK;
  const int nc = 8 * a.nt, pitch = nc + 4;
  CUtensorMap mx;
  const cuuint64_t gdim[2] = {(cuuint64_t)a.p, (cuuint64_t)a.nrows};
  const cuuint64_t gstr[1] = {(cuuint64_t)a.ldx * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)pitch, (cuuint32_t)kGR};
  const cuuint32_t estr[2] = {1, 1};
  static const bool verbose = [] { const char *e = getenv("LSSPA_GRAM_TMA"); return e && e[0] == '2'; }();
  const CUresult er = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(a.X), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (er != CUDA_SUCCESS) {
    if (verbose) fprintf(stderr, "lsspa: cuTensorMapEncodeTiled failed (%d): p=%d rows=%lld ldx=%lld box=%dx%d\n", (int)er, a.p,
                         (long long)a.nrows, (long long)a.ldx, pitch, kGR);
    return LSSPA_OK;
  }
  int rc = LSSPA_E_UNSUPPORTED;
  switch (a.nt) {
    case 1: rc = launch_gram_tma_nt<1>(mx, a, st); break;
    case 2: rc = launch_gram_tma_nt<2>(mx, a, st); break;
    case 3: rc = launch_gram_tma_nt<3>(mx, a, st); break;
    case 4: rc = launch_gram_tma_nt<4>(mx, a, st); break;
    case 5: rc = launch_gram_tma_nt<5>(mx, a, st); break;
    case 6: rc = launch_gram_tma_nt<6>(mx, a, st); break;
    case 7: rc = launch_gram_tma_nt<7>(mx, a, st); break;
    case 8: rc = launch_gram_tma_nt<8>(mx, a, st); break;
    case 9: rc = launch_gram_tma_nt<9>(mx, a, st); break;
    case 10: rc = launch_gram_tma_nt<10>(mx, a, st); break;
    case 11: rc = launch_gram_tma_nt<11>(mx, a, st); break;
    case 12: rc = launch_gram_tma_nt<12>(mx, a, st); break;
    case 13: rc = launch_gram_tma_nt<13>(mx, a, st); break;
    case 14: rc = launch_gram_tma_nt<14>(mx, a, st); break;
    case 15: rc = launch_gram_tma_nt<15>(mx, a, st); break;
    default: break;
  }
  used = rc == LSSPA_OK;
  return rc;
}

// G[e] = scale * sum_cta parts[cta][e], fixed order (deterministic)
__global__ void gram_sum_kernel(const double *parts, int count, int n2, int nc, double scale, double *G) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n2) return;
  const int i = e / nc, j = e - i * nc;
  double s = 0.0;
  if ((j >> 3) >= (i >> 3))      // upper tiles only (the TMA kernel leaves the lower ones unwritten)
    for (int k = 0; k < count; ++k) s += parts[(size_t)k * n2 + e];
  G[e] = s * scale;
}

// 1/sqrt(d) from the 20-bit MUFU seed and two correction steps (cubic, then quadratic): the library
// rsqrt() carries a slow-path call on the pivot chain of chol_factor_kernel
__device__ __forceinline__ double rsqrt_seed3(double d) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d * r, r, 1.0);
  r = fma(r * e, fma(0.375, e, 0.5), r);
  e = fma(-d * r, r, 1.0);
  return fma(0.5 * r, e, r);
}

// Cholesky G = R^T R (upper, G row-major nc x nc, leading q x q block used), then R^-1.
// One CTA, every instruction executed a handful of times: the run time is the INSTRUCTION FETCH of the code
// touched (cold: the row passes evict it from L2 between launches), so the cold loops are kept rolled.
// info[0] = 0 ok / 1 non-positive or tiny pivot (relative to the column's own norm),
// info[1] = a bound >= cond_2(R') of the column-equilibrated factor R' = R D^-1 (the smaller of
// |R'|_F |R'^-1|_F and sqrt(max row sum |G'|) sqrt(|R'^-1|_1 |R'^-1|_inf)),
// D = diag(sqrt(G_jj)): the accuracy of a Cholesky factor is governed by the conditioning of the
// equilibrated matrix (van der Sluis / Demmel), so units of the columns must not count.
// R is written row-major q x q (slot layout of reduce.cu); Rinv row-major [nc][ldr], zero padded.
__global__ void __launch_bounds__(1024) chol_factor_kernel(const double *G, int q, int nc, int ldr, double *R_out,
                                                           double *Rinv_out, double *info, double *gram_out) {
  extern __shared__ double sm[];
  double *A = sm;                        // nc x nc working copy (upper part), row-major
  double *Vi = A + (size_t)nc * nc;      // nc x nc inverse, row-major
  __shared__ double s_fail, red[64];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int tx = tid & 31, ty = tid >> 5, lane = tx, w = ty;
#pragma unroll 1
  for (int e = tid; e < nc * nc; e += nt) {
    const int i = e / nc, j = e - i * nc;
    A[e] = (i < q && j < q && j >= i) ? G[e] : 0.0;
    Vi[e] = 0.0;
  }
  if (tid == 0) s_fail = 0.0;
  __syncthreads();
  // everything that needs the Gram matrix itself is taken from the shared-memory copy BEFORE it is
  // overwritten by the factor: column scales d_k = sqrt(G_kk), the Gershgorin row sums of the equilibrated
  // matrix (whole matrix / leading block without the target column) and the record for the lift route
  __shared__ double dsq[128], rdsq[128], rdiag[128], gdiag[128];
  __shared__ double b_g[2][128], b_c[2][128], b_r[2][128];
  const int pp = q - 1;   // leading block = the features without the target column
#pragma unroll 1
  for (int k = tid; k < q; k += nt) {
    gdiag[k] = A[(size_t)k * nc + k];
    const double d = sqrt(fmax(A[(size_t)k * nc + k], 0.0));
    dsq[k] = d;
    rdsq[k] = d > 0.0 ? 1.0 / d : 0.0;
  }
  __syncthreads();
#pragma unroll 1
  for (int t = w; t < q; t += 32) {
    double g = 0.0, gp = 0.0;
#pragma unroll 1
    for (int k = lane; k < q; k += 32) {
      const double gij = (k >= t) ? A[(size_t)t * nc + k] : A[(size_t)k * nc + t];   // upper part
      const double ge = fabs(gij) * (rdsq[t] * rdsq[k]);
      g += ge;
      if (k < pp) gp += ge;
    }
    g = warp_sum(g);
    gp = warp_sum(gp);
    if (lane == 0) {
      b_g[0][t] = g;
      b_g[1][t] = gp;
    }
  }
  if (gram_out != nullptr) {
    // Gh = [R D^-1 | c]^T [R D^-1 | c] = the Gram matrix with its feature rows / columns scaled to a unit
    // diagonal, and D -- what lsspa_lifts_gram derives from the factor, here straight from G (= R^T R)
#pragma unroll 1
    for (int e = tid; e < q * q; e += nt) {
      const int i = e / q, j = e - i * q;
      const double gij = (j >= i) ? A[(size_t)i * nc + j] : A[(size_t)j * nc + i];
      const double di = (i < pp) ? (dsq[i] > 0.0 ? dsq[i] : 1.0) : 1.0;
      const double dj = (j < pp) ? (dsq[j] > 0.0 ? dsq[j] : 1.0) : 1.0;
      gram_out[e] = gij / (di * dj);
    }
#pragma unroll 1
    for (int k = tid; k < pp; k += nt) gram_out[(size_t)q * q + 8 + k] = dsq[k] > 0.0 ? dsq[k] : 1.0;
  }
  __syncthreads();
  const long long c0 = clock64();
  double dmax = 0.0;
#pragma unroll 1
  for (int i = 0; i < q; ++i) dmax = fmax(dmax, A[(size_t)i * nc + i]);
  // Right-looking Cholesky with the matrix in REGISTERS: thread (ty, tx) owns the entries (i, j) with
  // i = ty + 32 ii, j = tx + 32 jj.  Row k belongs to warp k % 32, which gets the pivot by one shuffle,
  // scales its row and publishes it through the double-buffered rk: one barrier per column.
  double ar[4][4];
#pragma unroll
  for (int ii = 0; ii < 4; ++ii)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int i = ty + 32 * ii, j = tx + 32 * jj;
      ar[ii][jj] = (i < nc && j < nc) ? A[(size_t)i * nc + j] : 0.0;
    }
  __shared__ double rkb[2][128];
  // (the column loop is split into 32-column blocks so that every register index is a compile-time constant)
#pragma unroll
  for (int kb = 0; kb < 4; ++kb) {
    for (int kk = 0; kk < 32; ++kk) {
      const int k = 32 * kb + kk;
      if (k >= q) break;
      double *rkc = rkb[k & 1];
      if (ty == kk) {
        const double d = __shfl_sync(kFull, ar[kb][kb], kk);
        const bool good = d > 1e-14 * gdiag[k] && dmax > 0.0;
        if (!good && tx == 0) s_fail = 1.0;
        const double dd = d > 0.0 ? d : 1.0;
        const double ri = rsqrt_seed3(dd);
#pragma unroll
        for (int jj = kb; jj < 4; ++jj) {
          const int j = tx + 32 * jj;
          if (j >= k && j < nc) {
            const double v = (j == k) ? dd * ri : ar[kb][jj] * ri;
            ar[kb][jj] = v;
            rkc[j] = v;
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int ii = kb; ii < 4; ++ii) {
        const int i = ty + 32 * ii;
        if (i > k && i < q) {
          const double ai = rkc[i];
#pragma unroll
          for (int jj = ii; jj < 4; ++jj) {
            const int j = tx + 32 * jj;
            if (j >= i && j < q) ar[ii][jj] = fma(-ai, rkc[j], ar[ii][jj]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int ii = 0; ii < 4; ++ii)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int i = ty + 32 * ii, j = tx + 32 * jj;
      if (i < nc && j < nc) A[(size_t)i * nc + j] = (j >= i && i < q && j < q) ? ar[ii][jj] : 0.0;
    }
  __syncthreads();
  const long long c1 = clock64();
  // inverse: column n of R^-1 by back substitution, four lanes per column (they split the dot product of
  // each step and combine with two shuffles; a column only touches its own entries of Vi).  The reciprocals
  // of the diagonal are taken once, in parallel: no division on the chain of 101 dependent steps.
#pragma unroll 1
  for (int k = tid; k < q; k += nt) rdiag[k] = 1.0 / A[(size_t)k * nc + k];
  __syncthreads();
  {
    const int n = tid >> 2, part = tid & 3;
#pragma unroll 1
    for (int i = q - 1; i >= 0; --i) {
      const double *row = A + (size_t)i * nc;
      double sacc = 0.0;
      if (n < q && i <= n) {
#pragma unroll 4
        for (int k = i + 1 + part; k <= n; k += 4) sacc = fma(-row[k], Vi[(size_t)k * nc + n], sacc);
      }
      sacc += __shfl_xor_sync(kFull, sacc, 1);
      sacc += __shfl_xor_sync(kFull, sacc, 2);
      if (n < q && i <= n && part == 0) Vi[(size_t)i * nc + n] = (sacc + ((i == n) ? 1.0 : 0.0)) * rdiag[i];
      __syncwarp();
    }
  }
  __syncthreads();
  const long long c2 = clock64();
  // Frobenius bounds of the whole factor (fr, fi) and of its leading pp x pp block (frp, fip): the
  // inverse of the leading block of a triangular matrix is the leading block of its inverse
  double fr = 0.0, fi = 0.0, frp = 0.0, fip = 0.0;
  // (R and R^-1 are upper triangular: warp per row, lanes over the columns j >= i; reciprocal scales from rdsq)
#pragma unroll 1
  for (int i = w; i < q; i += 32) {
    const double di = dsq[i];
#pragma unroll 1
    for (int j = i + lane; j < q; j += 32) {
      const double rr = A[(size_t)i * nc + j] * rdsq[j];
      const double v = Vi[(size_t)i * nc + j] * di;
      fr = fma(rr, rr, fr);
      fi = fma(v, v, fi);
      if (j < pp) {          // i <= j < pp
        frp = fma(rr, rr, frp);
        fip = fma(v, v, fip);
      }
    }
  }
  fr = warp_sum(fr);
  fi = warp_sum(fi);
  frp = warp_sum(frp);
  fip = warp_sum(fip);
  __shared__ double redp[64];
  if (lane == 0) {
    red[w] = fr;
    red[32 + w] = fi;
    redp[w] = frp;
    redp[32 + w] = fip;
  }
  __syncthreads();
  // second bound: sqrt(Gershgorin bound on the largest eigenvalue of the equilibrated Gram matrix)
  // * sqrt(|R'^-1|_1 |R'^-1|_inf) with R'^-1 = D R^-1; thread t takes row / column t.  [1]: leading block.
#pragma unroll 1
  for (int t = w; t < q; t += 32) {          // warp per row / column t, lanes over k
    const double dt = dsq[t];
    double cs = 0.0, rs = 0.0, csp = 0.0, rsp = 0.0;
#pragma unroll 1
    for (int k = lane; k < q; k += 32) {
      const double ce = fabs(Vi[(size_t)k * nc + t]) * dsq[k];     // column t of D R^-1
      const double re = fabs(Vi[(size_t)t * nc + k]) * dt;         // row t
      cs += ce;
      rs += re;
      if (k < pp) {
        csp += ce;
        rsp += re;
      }
    }
    cs = warp_sum(cs);
    rs = warp_sum(rs);
    csp = warp_sum(csp);
    rsp = warp_sum(rsp);
    if (lane == 0) {
      b_c[0][t] = cs;
      b_r[0][t] = rs;
      b_c[1][t] = csp;
      b_r[1][t] = rsp;
    }
  }
  __syncthreads();
  // the final combination is done by warp 0 with its lanes over the rows / warps (a single thread walking
  // these ~600 shared-memory values and ~100 FP64 divisions was half of this kernel's time)
  double bound0 = 0.0, bound1 = 0.0, dlo = 1e300, dhi = 0.0;
  if (w == 0) {
#pragma unroll 1
    for (int blk = 0; blk < 2; ++blk) {
      const int n = blk == 0 ? q : pp;
      const double *rd = blk == 0 ? red : redp;
      double a = (lane < nt / 32) ? rd[lane] : 0.0, b = (lane < nt / 32) ? rd[32 + lane] : 0.0;
      a = warp_sum(a);
      b = warp_sum(b);
      double gm = 0.0, cm = 0.0, rm = 0.0;
#pragma unroll 1
      for (int t = lane; t < n; t += 32) {
        gm = fmax(gm, b_g[blk][t]);
        cm = fmax(cm, b_c[blk][t]);
        rm = fmax(rm, b_r[blk][t]);
      }
      gm = warp_max(gm);
      cm = warp_max(cm);
      rm = warp_max(rm);
      double frob = sqrt(a) * sqrt(b), sharp = sqrt(gm) * sqrt(cm * rm);
      if (!(frob == frob)) frob = INFINITY;
      if (!(sharp == sharp)) sharp = INFINITY;
      if (blk == 0) bound0 = fmin(frob, sharp);
      else bound1 = fmin(frob, sharp);
    }
    double dneg = -1e300;
#pragma unroll 1
    for (int k = lane; k < pp; k += 32) {
      const double d = (dsq[k] > 0.0) ? fabs(A[(size_t)k * nc + k]) / dsq[k] : 0.0;
      dneg = fmax(dneg, -d);
      dhi = fmax(dhi, d);
    }
    dlo = -warp_max(dneg);
    dhi = warp_max(dhi);
  }
  if (tid == 0) {
    info[0] = s_fail;
    info[1] = bound0;
    if (gram_out != nullptr) {
      // the record lsspa_lifts_gram leaves behind its Gram matrix: [0] condition bound of the equilibrated
      // train factor (leading block), [1] min / max of its diagonal
      double *ginfo = gram_out + (size_t)q * q;
      ginfo[0] = (s_fail != 0.0) ? INFINITY : bound1;
      ginfo[1] = (dhi > 0.0) ? dlo / dhi : 0.0;
      ginfo[2] = bound1;
      ginfo[3] = bound1;
      // phase cycles of this (single-CTA) kernel as thread 0 sees them, for tools/prof_gram.py: factorisation,
      // inverse, bounds.  (BAR.SYNC defers blocking: a clock read after a barrier is taken when THIS warp
      // arrives, so the wait for the slowest warp of a phase is booked to the next phase.)
      ginfo[4] = (double)(c1 - c0);
      ginfo[5] = (double)(c2 - c1);
      ginfo[6] = (double)(clock64() - c2);
    }
  }
#pragma unroll 1
  for (int e = tid; e < q * q; e += nt) {
    const int i = e / q, j = e - i * q;
    R_out[e] = A[(size_t)i * nc + j];
  }
#pragma unroll 1
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

// (p + 1 <= 112: the Cholesky kernel keeps the factor and its inverse, 2 x (8 nt)^2 doubles, plus ~11 KB of
// static tables in shared memory -- 2 x 120^2 doubles alone are 225 KB and do not leave room for them)
extern "C" int lsspa_gram_supported(int p) { return (p >= 1 && p + 1 <= 112) ? 1 : 0; }

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
  bool used = false;
  const int rc = launch_gram_tma(a, st, used);
  if (rc != LSSPA_OK || used) return rc;
  if (a.nt <= 4) return launch_gram<4>(a, smem, st);
  if (a.nt <= 8) return launch_gram<8>(a, smem, st);
  if (a.nt <= 13) return launch_gram<13>(a, smem, st);
  return launch_gram<16>(a, smem, st);
}

extern "C" int lsspa_gram_finish(const double *parts, int count, int p, double scale, double *G_out,
                                 void *stream) {
  if (!parts || !G_out || !lsspa_gram_supported(p) || count < 1) return LSSPA_E_BADARG;
  const int n2 = (int)lsspa_gram_slot_doubles(p);
  gram_sum_kernel<<<(n2 + 255) / 256, 256, 0, as_stream(stream)>>>(parts, count, n2, 8 * gram_nt(p), scale, G_out);
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

extern "C" int lsspa_chol_factor_gram(const double *G, int p, double *R_out, double *Rinv_out, double *info,
                                      double *gram_out_or_null, void *stream) {
  if (!G || !R_out || !Rinv_out || !info || !lsspa_gram_supported(p)) return LSSPA_E_BADARG;
  const int nt = gram_nt(p), nc = 8 * nt, ldr = gram_ldr(nt);
  const size_t smem = (size_t)2 * nc * nc * sizeof(double);
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(chol_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  chol_factor_kernel<<<1, 1024, smem, as_stream(stream)>>>(G, p + 1, nc, ldr, R_out, Rinv_out, info, gram_out_or_null);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

extern "C" int lsspa_chol_factor(const double *G, int p, double *R_out, double *Rinv_out, double *info,
                                 void *stream) {
  return lsspa_chol_factor_gram(G, p, R_out, Rinv_out, info, nullptr, stream);
}

extern "C" int lsspa_tri_product(const double *R2, const double *R1, int p, const double *G1, double *out_slot,
                                 void *stream) {
  if (!R2 || !R1 || !G1 || !out_slot || !lsspa_gram_supported(p)) return LSSPA_E_BADARG;
  const int q = p + 1, nc = 8 * gram_nt(p);
  tri_product_kernel<<<(q * q + 255) / 256, 256, 0, as_stream(stream)>>>(R2, R1, q, G1, nc, out_slot);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}
