// Gram reduction of [X | y] for WIDE problems (p + 1 > 112: the upper triangle no longer fits the
// registers of one CTA as in gram.cu).  Replaces np.linalg.qr of the tall blocks (reference
// ls_spa/ls_spa.py:314-315) by G = Z^T Z (fp64 tensor tiles) followed by a blocked Cholesky factorisation
// of G (lifts_big.cu, the batched tile kernels with a batch of one).
//
//   gram_big_kernel   CTA = (pair of 128-column blocks a <= b, row split s): streams its rows in chunks
//                     of 32 through double-buffered shared memory (row-major, stride % 16 == 4:
//                     conflict-free fragments) and keeps the 128 x 128 block of G in registers
//                     (8 warps x 32 DMMA tiles).  Partial blocks -> parts[s][pair].
//   gram_big_acc      G_acc += sum_s parts[s]  (dense q x q row-major, blocks a <= b), fixed order.
//
// Intensity: a CTA reads 2 x 128 columns per row for 128 x 128 x 2 flop -> 16 flop/B, above the FP64
// ridge; algorithmic bytes 8 N q are re-read nb/2 times from L2/HBM (nb = q / 128 column blocks).

#include "common.cuh"

namespace lsspa {
namespace {

constexpr int kBB = 128;        // column block
constexpr int kBR = 32;         // rows per chunk
constexpr int kLD = kBB + 4;    // shared-memory row stride (doubles), % 16 == 4

__device__ __forceinline__ void dmma_b(double &d0, double &d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

struct GramBigParams {
  const double *X;
  int64_t ldx;
  const double *y;
  int64_t nrows;
  int p, q, nb, nsplit;
  double *parts;   // [nsplit][npairs][128 * 128]
};

__device__ __forceinline__ void pair_from_index(int idx, int nb, int &a, int &b) {
  a = 0;
  while (idx >= nb - a) {
    idx -= nb - a;
    ++a;
  }
  b = a + idx;
}

__global__ void __launch_bounds__(256, 1) gram_big_kernel(GramBigParams g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sm = reinterpret_cast<double *>(smem_raw);   // [2 buffers][2 blocks][kBR][kLD]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = lane >> 2, q4 = lane & 3;
  int ba, bb;
  pair_from_index(blockIdx.x, g.nb, ba, bb);
  const bool same = ba == bb;
  const int split = blockIdx.y;
  const int64_t per = ceil_div(ceil_div(g.nrows, (int64_t)g.nsplit), (int64_t)kBR) * kBR;
  const int64_t r_begin = (int64_t)split * per;
  const int64_t r_end = (r_begin + per < g.nrows) ? r_begin + per : g.nrows;

  // loader: thread (rg = tid / 128, col = tid % 128) moves rows 2u + rg, u = 0..15, of both blocks
  const int rg = tid >> 7, col = tid & 127;
  const int colA = kBB * ba + col, colB = kBB * bb + col;
  double stA[kBR / 2], stB[kBR / 2];
  auto fetch = [&](int64_t r0) {
#pragma unroll
    for (int u = 0; u < kBR / 2; ++u) {
      const int64_t r = r0 + 2 * u + rg;
      double va = 0.0, vb = 0.0;
      if (r < r_end) {
        if (colA < g.p) va = g.X[r * g.ldx + colA];
        else if (colA == g.p) va = g.y[r];
        if (!same) {
          if (colB < g.p) vb = g.X[r * g.ldx + colB];
          else if (colB == g.p) vb = g.y[r];
        }
      }
      stA[u] = va;
      stB[u] = vb;
    }
  };
  auto commit = [&](double *buf) {
#pragma unroll
    for (int u = 0; u < kBR / 2; ++u) {
      buf[(size_t)(2 * u + rg) * kLD + col] = stA[u];
      if (!same) buf[(size_t)kBR * kLD + (size_t)(2 * u + rg) * kLD + col] = stB[u];
    }
  };

  double acc[4][8][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  const int wm = warp >> 1, wn = warp & 1;   // micro-rows 4 wm .. +3, micro-columns 8 wn .. +7

  if (r_begin < r_end) {
    fetch(r_begin);
    commit(sm);
  }
  __syncthreads();
  int buf = 0;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += kBR) {
    const double *As = sm + (size_t)buf * 2 * kBR * kLD;
    const double *Bs = same ? As : As + (size_t)kBR * kLD;
    const bool more = r0 + kBR < r_end;
    if (more) fetch(r0 + kBR);
#pragma unroll 2
    for (int ks = 0; ks < kBR / 4; ++ks) {
      const double *ra = As + (size_t)(4 * ks + q4) * kLD + c;
      const double *rb = Bs + (size_t)(4 * ks + q4) * kLD + c;
      double fa[4], fb[8];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) fa[mi] = ra[8 * (4 * wm + mi)];
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) fb[ni] = rb[8 * (8 * wn + ni)];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) dmma_b(acc[mi][ni][0], acc[mi][ni][1], fa[mi], fb[ni]);
    }
    if (more) commit(sm + (size_t)(buf ^ 1) * 2 * kBR * kLD);
    __syncthreads();
    buf ^= 1;
  }
  const int npairs = g.nb * (g.nb + 1) / 2;
  double *out = g.parts + ((size_t)split * npairs + blockIdx.x) * kBB * kBB;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 8; ++ni)
      *reinterpret_cast<double2 *>(out + (size_t)(8 * (4 * wm + mi) + c) * kBB + 8 * (8 * wn + ni) + 2 * q4) =
          make_double2(acc[mi][ni][0], acc[mi][ni][1]);
}

// G_acc[i][j] += sum_s parts[s][pair(i / 128, j / 128)][i % 128][j % 128]   for block pairs a <= b
__global__ void gram_big_acc_kernel(const double *parts, int nsplit, int q, int nb, double *G) {
  const int npairs = nb * (nb + 1) / 2;
  const int pair = blockIdx.y;
  int a, b;
  pair_from_index(pair, nb, a, b);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;   // element of the block
  if (e >= kBB * kBB) return;
  const int i = kBB * a + e / kBB, j = kBB * b + e % kBB;
  if (i >= q || j >= q) return;
  double s = 0.0;
  for (int k = 0; k < nsplit; ++k) s += parts[((size_t)k * npairs + pair) * kBB * kBB + e];
  G[(size_t)i * q + j] += s;
}

}  // namespace
}  // namespace lsspa

using namespace lsspa;

extern "C" int lsspa_gram_big_supported(int p) { return (p >= 1 && p <= 2047) ? 1 : 0; }

static int gram_big_nb(int p) { return (p + 1 + kBB - 1) / kBB; }

extern "C" int lsspa_gram_big_num_splits(int p, int64_t nrows) {
  if (!lsspa_gram_big_supported(p) || nrows < 1) return 0;
  const DeviceInfo &d = device_info();
  const int sms = d.sm_count > 0 ? d.sm_count : 148;
  const int nb = gram_big_nb(p), npairs = nb * (nb + 1) / 2;
  int64_t s = sms / npairs;
  if (s < 1) s = 1;
  const int64_t most = ceil_div(nrows, (int64_t)4 * kBR);
  if (s > most) s = most;
  return (int)s;
}

extern "C" int64_t lsspa_gram_big_part_doubles(int p) {
  if (!lsspa_gram_big_supported(p)) return 0;
  const int nb = gram_big_nb(p);
  return (int64_t)(nb * (nb + 1) / 2) * kBB * kBB;
}

extern "C" int lsspa_gram_big_rows(const double *X, int64_t ldx, const double *y, int64_t nrows, int p, double *parts,
                                   int nsplit, void *stream) {
  if (!X || !y || !parts || !lsspa_gram_big_supported(p) || nrows < 1 || ldx < p || nsplit < 1) return LSSPA_E_BADARG;
  GramBigParams g;
  g.X = X;
  g.ldx = ldx;
  g.y = y;
  g.nrows = nrows;
  g.p = p;
  g.q = p + 1;
  g.nb = gram_big_nb(p);
  g.nsplit = nsplit;
  g.parts = parts;
  const size_t smem = (size_t)2 * 2 * kBR * kLD * sizeof(double);
  LSSPA_CUDA_TRY(cudaFuncSetAttribute(gram_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gram_big_kernel<<<dim3((unsigned)(g.nb * (g.nb + 1) / 2), (unsigned)nsplit), 256, smem, as_stream(stream)>>>(g);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}

extern "C" int lsspa_gram_big_accumulate(const double *parts, int nsplit, int p, double *G_acc, void *stream) {
  if (!parts || !G_acc || !lsspa_gram_big_supported(p) || nsplit < 1) return LSSPA_E_BADARG;
  const int nb = gram_big_nb(p);
  gram_big_acc_kernel<<<dim3(kBB * kBB / 256, (unsigned)(nb * (nb + 1) / 2)), 256, 0, as_stream(stream)>>>(
      parts, nsplit, p + 1, nb, G_acc);
  LSSPA_LAUNCH_CHECK();
  return LSSPA_OK;
}
