// SELL-32 device code and Krylov bookkeeping shared by the single-GPU drivers (krylov.cu) and the
// peer-to-peer multi-GPU CG (dcg.cu).
//
// Solver-internal matrix layout (sliced ELLPACK, slice height 32 = one warp, no row sorting):
// slice t holds rows [32t, 32t+32); its width W_t is the longest row rounded up to an even number;
// entry k of row (32t + lane) lives at  slice_ptr[t] + (k/2)*64 + lane*2 + (k%2),  i.e. every lane
// reads ITS row with 128-bit loads while the warp as a whole reads 512 contiguous bytes per
// instruction (4 L1 wavefronts per 64 nonzeros — the CSR-chunk kernel of spmv.cuh needs ~1 wavefront per
// nonzero, which caps it at ~35 % of HBM bandwidth; see profiles/). FEM rows of neighbouring nodes
// have equal length, so padding is < 1 % on the Hexa1 cube. Padding entries are (col = own row, 0.0).
//
// Node-block column indices: FEM rows come in groups of dpn (the DOFs of one node) that share their column
// BLOCKS: entry k of a row has column dpn*adj[k/dpn] + k%dpn. Storing one int32 per (node, block) instead
// of one per entry cuts the index stream from 4 B to 4/dpn^2 B per nonzero (8.5 instead of 12 B/nnz for
// dpn = 3). Layout, in slice order so the load is one wavefront: bcols[bslice_ptr[t] + kb*NPS + m] = block
// column kb of the m-th node touched by slice t (first node nf = 32t/dpn, NPS = 12 nodes for dpn 3, 16 for
// dpn 2). Padding repeats the node's last block (values there are 0.0). Matrices with unreferenced nodes
// (rows of length 1) keep scalar columns.
#pragma once
#include <stdlib.h>

#include "spmv.cuh"

namespace tfem {
namespace {

constexpr int kVecThreads = 256;
constexpr int kSellWarps = 8;

// device-resident scalars of a Krylov solve
enum Sc {
  SC_RHO = 0, SC_RHO_PREV, SC_PQ, SC_RR, SC_TOL, SC_BNRM, SC_DONE, SC_ITERS, SC_ALPHA, SC_BETA,
  // MINRES recurrences (scipy/sparse/linalg/_isolve/minres.py)
  SC_M_BETA1, SC_M_OLDB, SC_M_BETA, SC_M_DBAR, SC_M_EPSLN, SC_M_PHIBAR, SC_M_CS, SC_M_SN, SC_M_TNORM2,
  SC_M_GMAX, SC_M_GMIN, SC_M_ALFA, SC_M_YNORM2, SC_M_PHI, SC_M_DENOM, SC_M_OLDEPS, SC_M_DELTA,
  SC_M_RNORM, SC_M_ISTOP, SC_COUNT = 32
};

struct Sell {
  int64_t n, n_slices;
  const int64_t* slice_ptr;  // [n_slices+1], element offsets (multiples of 64)
  const int32_t* cols;
  const double* vals;
  int64_t dot_lo = 0, dot_hi = INT64_MAX;  // rows that enter the fused x.y dot (owned rows of a rank)
  // distributed multigrid (amg.cu, k_amg_spmv): only rows [dot_lo, dot_hi) are computed and WRITTEN (a peer may be
  // storing into the halo entries of the output at the same time); slices outside [slice_lo, slice_hi) are skipped
  int64_t slice_lo = 0, slice_hi = -1;     // -1: all slices
  // rows longer than TFEM_SELL_LONG_ROW are empty in the slices and computed from the CSR arrays (k_sell_long)
  int n_long = 0;
  const int32_t* long_rows = nullptr;
  const int64_t* csr_indptr = nullptr;
  const int32_t* csr_cols = nullptr;
  const double* csr_vals = nullptr;
  // optional node-block column indices: one int per (node, block) instead of one per entry
  const int64_t* bslice_ptr = nullptr;
  const int32_t* bcols = nullptr;
  int dpn = 0;
};

inline Sell make_sell(const tfem_sell_t* a) {
  Sell A;
  A.n = a->n_rows;
  A.n_slices = (a->n_rows + 31) / 32;
  A.slice_ptr = a->slice_ptr;
  A.cols = a->cols;
  A.vals = a->vals;
  A.bslice_ptr = a->bslice_ptr;
  A.bcols = a->bcols;
  A.dpn = a->bcols ? a->dpn : 0;
  A.n_long = a->n_long;
  A.long_rows = a->long_rows;
  A.csr_indptr = a->csr_indptr;
  A.csr_cols = a->csr_cols;
  A.csr_vals = a->csr_vals;
  return A;
}

inline int check_sell(const tfem_sell_t* a) {
  TFEM_REQUIRE(a && a->slice_ptr && a->vals && a->n_rows > 0, "SELL matrix: null pointer or empty");
  TFEM_REQUIRE(a->cols || (a->bcols && a->bslice_ptr), "SELL matrix: neither scalar nor block columns given");
  TFEM_REQUIRE(aligned16(a->vals) && (!a->cols || aligned16(a->cols)), "SELL arrays must be 16-byte aligned");
  TFEM_REQUIRE(!a->bcols || a->dpn == 2 || a->dpn == 3, "block columns need 2 or 3 DOFs per node");
  TFEM_REQUIRE(a->n_long >= 0 && (a->n_long == 0 || (a->long_rows && a->csr_indptr && a->csr_cols && a->csr_vals)),
               "SELL matrix: long rows need the row list and the CSR arrays");
  return TFEM_OK;
}

template <int DPN> struct Nps { static constexpr int v = (DPN == 3) ? 12 : 16; };

// (A x)[32 t + lane] for the calling warp's lane, scalar column indices. Sequential accumulation per row
// (the order of scipy's CSR matvec).
template <bool COHERENT>
__device__ __forceinline__ double ldx(const double* p) {
  if constexpr (COHERENT) return __ldcg(p);
  else return __ldg(p);
}

template <bool COHERENT = false>
__device__ __forceinline__ double sell_slice_row(const Sell& A, int64_t t, const double* __restrict__ x,
                                                 int lane) {
  const int2* c2 = reinterpret_cast<const int2*>(A.cols);
  const double2* v2 = reinterpret_cast<const double2*>(A.vals);
  const int64_t b2 = (A.slice_ptr[t] >> 1) + lane;
  const int w2 = (int)((A.slice_ptr[t + 1] - A.slice_ptr[t]) >> 6);  // 128-bit steps
  double acc = 0.0;
  int s = 0;
  for (; s + 4 <= w2; s += 4) {
    int2 c[4];
    double2 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t at = b2 + (int64_t)(s + u) * 32;
      asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(c[u].x), "=r"(c[u].y) : "l"(c2 + at));
      v[u] = ldg_stream_double2(v2 + at);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      acc = fma(v[u].x, ldx<COHERENT>(x + c[u].x), acc);
      acc = fma(v[u].y, ldx<COHERENT>(x + c[u].y), acc);
    }
  }
  for (; s < w2; ++s) {
    const int64_t at = b2 + (int64_t)s * 32;
    int2 c;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(c.x), "=r"(c.y) : "l"(c2 + at));
    const double2 v = ldg_stream_double2(v2 + at);
    acc = fma(v.x, ldx<COHERENT>(x + c.x), acc);
    acc = fma(v.y, ldx<COHERENT>(x + c.y), acc);
  }
  return acc;
}

// The same with node-block column indices.
template <int DPN, bool COHERENT = false>
__device__ __forceinline__ double bsell_slice_row(const Sell& A, int64_t t, const double* __restrict__ x,
                                                  int lane) {
  constexpr int NPS = Nps<DPN>::v;
  const double2* v2 = reinterpret_cast<const double2*>(A.vals);
  const int64_t b2 = (A.slice_ptr[t] >> 1) + lane;
  const int w2 = (int)((A.slice_ptr[t + 1] - A.slice_ptr[t]) >> 6);  // 128-bit steps
  const int64_t row = t * 32 + lane;
  const int m = (int)(row / DPN - (t * 32) / DPN);  // my node within the slice
  const int32_t* bc = A.bcols + A.bslice_ptr[t] + m;
  double acc = 0.0;
  if (DPN == 2) {
    int s = 0;
    for (; s + 4 <= w2; s += 4) {
      double2 v[4];
      int c[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = ldg_stream_double2(v2 + b2 + (int64_t)(s + u) * 32);
        c[u] = 2 * __ldg(bc + (s + u) * NPS);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc = fma(v[u].x, ldx<COHERENT>(x + c[u]), acc);
        acc = fma(v[u].y, ldx<COHERENT>(x + c[u] + 1), acc);
      }
    }
    for (; s < w2; ++s) {
      const double2 v = ldg_stream_double2(v2 + b2 + (int64_t)s * 32);
      const int c = 2 * __ldg(bc + s * NPS);
      acc = fma(v.x, ldx<COHERENT>(x + c), acc);
      acc = fma(v.y, ldx<COHERENT>(x + c + 1), acc);
    }
  } else {
    // 6 entries = 3 x 128-bit value loads = 2 column blocks per step
    int s = 0, kb = 0;
    for (; s + 6 <= w2; s += 6, kb += 4) {
      double2 v[6];
      int c[4];
#pragma unroll
      for (int u = 0; u < 6; ++u) v[u] = ldg_stream_double2(v2 + b2 + (int64_t)(s + u) * 32);
#pragma unroll
      for (int u = 0; u < 4; ++u) c[u] = 3 * __ldg(bc + (kb + u) * NPS);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const double* x0 = x + c[2 * h];
        const double* x1 = x + c[2 * h + 1];
        acc = fma(v[3 * h].x, ldx<COHERENT>(x0), acc);
        acc = fma(v[3 * h].y, ldx<COHERENT>(x0 + 1), acc);
        acc = fma(v[3 * h + 1].x, ldx<COHERENT>(x0 + 2), acc);
        acc = fma(v[3 * h + 1].y, ldx<COHERENT>(x1), acc);
        acc = fma(v[3 * h + 2].x, ldx<COHERENT>(x1 + 1), acc);
        acc = fma(v[3 * h + 2].y, ldx<COHERENT>(x1 + 2), acc);
      }
    }
    for (; s + 3 <= w2; s += 3, kb += 2) {
      const double2 va = ldg_stream_double2(v2 + b2 + (int64_t)s * 32);
      const double2 vb = ldg_stream_double2(v2 + b2 + (int64_t)(s + 1) * 32);
      const double2 vc = ldg_stream_double2(v2 + b2 + (int64_t)(s + 2) * 32);
      const double* x0 = x + 3 * __ldg(bc + kb * NPS);
      const double* x1 = x + 3 * __ldg(bc + (kb + 1) * NPS);
      acc = fma(va.x, ldx<COHERENT>(x0), acc);
      acc = fma(va.y, ldx<COHERENT>(x0 + 1), acc);
      acc = fma(vb.x, ldx<COHERENT>(x0 + 2), acc);
      acc = fma(vb.y, ldx<COHERENT>(x1), acc);
      acc = fma(vc.x, ldx<COHERENT>(x1 + 1), acc);
      acc = fma(vc.y, ldx<COHERENT>(x1 + 2), acc);
    }
    // tail: fewer than 6 entries left; entry k uses block k/3, component k%3
    const int nblk = (2 * w2 + 2) / 3;  // blocks stored for this slice (ceil(W/3))
    for (; s < w2; ++s) {
      const double2 v = ldg_stream_double2(v2 + b2 + (int64_t)s * 32);
      const int k0 = 2 * s, k1 = 2 * s + 1;
      const int ka = k0 / 3 < nblk ? k0 / 3 : nblk - 1, kc = k1 / 3 < nblk ? k1 / 3 : nblk - 1;
      acc = fma(v.x, ldx<COHERENT>(x + 3 * __ldg(bc + ka * NPS) + k0 % 3), acc);
      acc = fma(v.y, ldx<COHERENT>(x + 3 * __ldg(bc + kc * NPS) + k1 % 3), acc);
    }
  }
  return acc;
}

// DPN = 0: scalar columns
template <int DPN, bool COHERENT = false>
__device__ __forceinline__ double slice_row(const Sell& A, int64_t t, const double* __restrict__ x, int lane) {
  if constexpr (DPN == 0) return sell_slice_row<COHERENT>(A, t, x, lane);
  else return bsell_slice_row<DPN, COHERENT>(A, t, x, lane);
}

// Persistent grids are sized from the kernel's REAL occupancy (registers may allow fewer resident CTAs
// than the 2048-thread limit; an oversized grid would run a second, nearly empty wave).
template <typename K>
inline int resident_ctas(K kernel, int threads) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  const int sms = num_sms() < kSMs ? num_sms() : kSMs;
  if (const char* cap = getenv("TFEM_CTAS_PER_SM")) {  // tuning knob for experiments (tools/prof_driver.py)
    const int c = atoi(cap);
    if (c >= 1 && c < per_sm) per_sm = c;
  }
  return sms * (per_sm > 8 ? 8 : per_sm);
}

// grids are cached per (instantiation, device): one process may drive several devices
template <typename K>
int cached_resident_ctas(K kernel, int threads) {
  static int g[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!g[dev]) g[dev] = resident_ctas(kernel, threads);
  return g[dev];
}

inline int vec_grid_cap() { return (num_sms() < kSMs ? num_sms() : kSMs) * 8; }

inline int vec_grid(int64_t n) {
  const int64_t want = (n + kVecThreads * 4 - 1) / (kVecThreads * 4);
  const int64_t cap = vec_grid_cap();
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

// y = A x, one warp per slice, persistent grid with the slices strided over it so that all SMs walk the same
// band of the matrix (the gathered x window stays in L2). DPN = 0: scalar columns, 2/3: node-block columns.
// DOT fuses x.y over the rows [dot_lo, dot_hi) (fixed-order reduction, last CTA writes *out_scalar).
template <int DPN, bool DOT>
__global__ void __launch_bounds__(kSellWarps * 32, 8)
    k_sell_spmv(Sell A, const double* __restrict__ x, double* __restrict__ y, const double* sc,
                double* partials, unsigned int* ticket, double* out_scalar) {
  __shared__ double s_red[kSellWarps];
  if (DOT && sc[SC_DONE] != 0.0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double dot = 0.0;
  const int64_t t_hi = A.slice_hi < 0 ? A.n_slices : A.slice_hi;
  for (int64_t t = A.slice_lo + (int64_t)blockIdx.x * kSellWarps + warp; t < t_hi;
       t += (int64_t)gridDim.x * kSellWarps) {
    const int64_t row = t * 32 + lane;
    // x[row] for the fused dot is requested BEFORE the row is streamed: asked for afterwards it costs every
    // warp one exposed memory latency per slice (64 us of a 1.2 ms launch at config B)
    double xr = 0.0;
    if (DOT && row < A.n && row >= A.dot_lo && row < A.dot_hi) xr = __ldg(x + row);
    const double acc = slice_row<DPN>(A, t, x, lane);
    if (row < A.n) {
      y[row] = acc;
      if (DOT) dot = fma(acc, xr, dot);
    }
  }
  if (DOT) {
    const double b = block_sum<kSellWarps * 32>(dot, s_red);
    double mine[1] = {b}, tot[1];
    if (publish_and_reduce<1>(mine, partials, ticket, tot) && threadIdx.x == 0) *out_scalar = tot[0];
  }
}

// Y[:, 0:nb] = A X[:, 0:nb] for row-major blocks of vectors (the eigensolver's K X / M X, reference sparse.py:798-1011 hands
// blocks to LOBPCG): the matrix is streamed once per MB vectors instead of once per vector, and each gathered row of X
// is nb contiguous doubles. Per row and column the entries are added in the order of k_sell_spmv: the result equals nb
// single products bit for bit.
// MODE 0: Y = A X;  1: Y = B - A X;  2: Y = X + omega dinv (B - A X)  (the residual and the damped-Jacobi sweep of the V cycle
// on a block of vectors, same expressions as k_amg_spmv). B and Y share the layout of X.
template <int DPN, int MB, int MODE = 0>
__global__ void __launch_bounds__(kSellWarps * 32)
    k_sell_spmm(Sell A, const double* __restrict__ X, int64_t ldx, double* __restrict__ Y, int64_t ldy, int nb,
                const double* __restrict__ B = nullptr, const double* __restrict__ dinv = nullptr, double omega = 0.0) {
  constexpr int NPS = Nps<DPN == 0 ? 3 : DPN>::v;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double2* v2 = reinterpret_cast<const double2*>(A.vals);
  for (int64_t t = (int64_t)blockIdx.x * kSellWarps + warp; t < A.n_slices; t += (int64_t)gridDim.x * kSellWarps) {
    const int64_t row = t * 32 + lane;
    const int64_t b2 = (A.slice_ptr[t] >> 1) + lane;
    const int w2 = (int)((A.slice_ptr[t + 1] - A.slice_ptr[t]) >> 6);  // 128-bit steps = entry pairs
    const int32_t* bc = nullptr;
    if constexpr (DPN != 0) bc = A.bcols + A.bslice_ptr[t] + (int)(row / DPN - (t * 32) / DPN);
    // the epilogue operands are requested before the row is streamed (one exposed latency per slice otherwise)
    double br[MB], xr[MB], dr = 0.0;
    if (MODE != 0 && row < A.n) {
      if (MODE == 2) dr = omega * __ldg(dinv + row);
#pragma unroll
      for (int j = 0; j < MB; ++j)
        if (j < nb) {
          br[j] = __ldg(B + row * ldx + j);
          if (MODE == 2) xr[j] = __ldg(X + row * ldx + j);
        }
    }
    double acc[MB];
#pragma unroll
    for (int j = 0; j < MB; ++j) acc[j] = 0.0;
#pragma unroll 2
    for (int s = 0; s < w2; ++s) {
      const double2 v = ldg_stream_double2(v2 + b2 + (int64_t)s * 32);
      int64_t c0, c1;
      if constexpr (DPN == 0) {
        const int2 c = __ldg(reinterpret_cast<const int2*>(A.cols) + b2 + (int64_t)s * 32);
        c0 = c.x;
        c1 = c.y;
      } else {
        const int k0 = 2 * s, k1 = 2 * s + 1;
        c0 = (int64_t)DPN * __ldg(bc + (k0 / DPN) * NPS) + k0 % DPN;
        c1 = (int64_t)DPN * __ldg(bc + (k1 / DPN) * NPS) + k1 % DPN;
      }
      const double* x0 = X + c0 * ldx;
      const double* x1 = X + c1 * ldx;
#pragma unroll
      for (int j = 0; j < MB; ++j)
        if (j < nb) acc[j] = fma(v.x, __ldg(x0 + j), acc[j]);
#pragma unroll
      for (int j = 0; j < MB; ++j)
        if (j < nb) acc[j] = fma(v.y, __ldg(x1 + j), acc[j]);
    }
    if (row < A.n) {
#pragma unroll
      for (int j = 0; j < MB; ++j)
        if (j < nb) {
          double out = acc[j];
          if (MODE == 1) out = br[j] - acc[j];
          if (MODE == 2) out = fma(dr, br[j] - acc[j], xr[j]);
          Y[row * ldy + j] = out;
        }
    }
  }
}

// 4 vectors per pass over the matrix, measured at config B (tools/time_spmm.py, profiles/r2_spmm.txt): 2.25 ms per pass
// against 4 x 1.15 ms of single products. 8 per pass is slower per vector (5.6 ms: the gathers of 64-byte rows of X, not
// the matrix stream, bound the kernel), and so is any leading dimension wider than the pass (the gathered rows waste their
// sectors: m = 16 unpacked took 19.5 ms, more than 16 single products) — the caller packs 4 columns per call (ldx = 4).
constexpr int kSpmmBlock = 4;

template <int DPN>
int launch_spmm_t(const Sell& A, int64_t m, const double* X, int64_t ldx, double* Y, int64_t ldy, cudaStream_t st) {
  const int g = cached_resident_ctas(k_sell_spmm<DPN, kSpmmBlock>, kSellWarps * 32);
  const int64_t want = A.n_slices > 0 ? (A.n_slices + kSellWarps - 1) / kSellWarps : 1;
  for (int64_t j0 = 0; j0 < m; j0 += kSpmmBlock) {
    const int nb = (int)(m - j0 < kSpmmBlock ? m - j0 : kSpmmBlock);
    k_sell_spmm<DPN, kSpmmBlock><<<(int)(want < g ? want : g), kSellWarps * 32, 0, st>>>(A, X + j0, ldx, Y + j0, ldy, nb);
  }
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

// The long rows (empty in the slices): one CTA per row straight from the CSR arrays. Every thread sums its strided
// entries in order, the partial sums are combined in a fixed tree: deterministic. DOT: the row's term x_r y_r is added
// to *out_scalar (which k_sell_spmv has written) by the CTA that finishes last, in row order.
constexpr int kLongThreads = 256;
template <bool DOT>
__global__ void __launch_bounds__(kLongThreads)
    k_sell_long(Sell A, const double* __restrict__ x, double* __restrict__ y, const double* sc, double* partials,
                unsigned int* ticket, double* out_scalar) {
  __shared__ double s_red[kLongThreads / 32];
  if (DOT && sc[SC_DONE] != 0.0) return;
  const int64_t r = A.long_rows[blockIdx.x];
  const int64_t b = A.csr_indptr[r], e = A.csr_indptr[r + 1];
  double acc = 0.0;
  for (int64_t k = b + threadIdx.x; k < e; k += kLongThreads) acc = fma(A.csr_vals[k], __ldg(x + A.csr_cols[k]), acc);
  const double row_sum = block_sum<kLongThreads>(acc, s_red);   // valid in thread 0
  double term = 0.0;
  if (threadIdx.x == 0) {
    y[r] = row_sum;
    if (DOT && r >= A.dot_lo && r < A.dot_hi) term = row_sum * x[r];
  }
  if (DOT) {
    double mine[1] = {term}, tot[1];
    if (publish_and_reduce<1>(mine, partials, ticket, tot) && threadIdx.x == 0) *out_scalar += tot[0];
  }
}

template <int DPN, bool DOT>
int launch_sell_t(const Sell& A, const double* x, double* y, const double* sc, double* partials,
                  unsigned int* ticket, double* out_scalar, cudaStream_t st) {
  const int g = cached_resident_ctas(k_sell_spmv<DPN, DOT>, kSellWarps * 32);  // per instantiation and device
  const int64_t n_sl = (A.slice_hi < 0 ? A.n_slices : A.slice_hi) - A.slice_lo;
  const int64_t want = n_sl > 0 ? (n_sl + kSellWarps - 1) / kSellWarps : 1;
  k_sell_spmv<DPN, DOT><<<(int)(want < g ? want : g), kSellWarps * 32, 0, st>>>(A, x, y, sc, partials, ticket,
                                                                                out_scalar);
  if (A.n_long > 0) k_sell_long<DOT><<<A.n_long, kLongThreads, 0, st>>>(A, x, y, sc, partials, ticket, out_scalar);
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

template <bool DOT>
int launch_sell(const Sell& A, const double* x, double* y, const double* sc, double* partials,
                unsigned int* ticket, double* out_scalar, cudaStream_t st) {
  if (A.dpn == 3) return launch_sell_t<3, DOT>(A, x, y, sc, partials, ticket, out_scalar, st);
  if (A.dpn == 2) return launch_sell_t<2, DOT>(A, x, y, sc, partials, ticket, out_scalar, st);
  return launch_sell_t<0, DOT>(A, x, y, sc, partials, ticket, out_scalar, st);
}

// ---- Krylov work buffer: 6 vectors + device scalars + reduction partials + ticket
struct Work {
  double *r, *p, *q;                      // CG
  double *r1, *r2, *y, *v, *w1, *w2;      // MINRES (aliases r/p/q for the first three)
  double* sc;
  double* partials;
  unsigned int* ticket;
};

constexpr int64_t kMaxPartials = 148 * 8 * 4;  // >= any grid used here, x up to 3 values per kernel

inline int64_t pad32(int64_t n) { return (n + 31) & ~(int64_t)31; }

inline Work carve(double* work, int64_t n) {
  const int64_t np = pad32(n);
  Work w;
  w.r = work;
  w.p = work + np;
  w.q = work + 2 * np;
  w.r1 = w.r;
  w.r2 = w.p;
  w.y = w.q;
  w.v = work + 3 * np;
  w.w1 = work + 4 * np;
  w.w2 = work + 5 * np;
  w.sc = work + 6 * np;
  w.partials = w.sc + SC_COUNT;
  w.ticket = reinterpret_cast<unsigned int*>(w.partials + kMaxPartials);
  return w;
}

}  // namespace
}  // namespace tfem
