// K4/K5/K6/K7 — CSR SpMV, Jacobi preconditioner and the fused Jacobi-PCG / MINRES drivers.
//
// Replaces the CuPy path of the reference (src/torchfem/sparse.py:350-421): COO->CSR conversion,
// `cupy_diags(1/A.diagonal())`, and cupy_cg / cupy_minres, which launch cuSPARSE SpMV + ~6 separate
// vector kernels per iteration and synchronise with the host every iteration to test convergence.
//
// Here one CG iteration is three launches (SpMV fused with p.q; x/r update fused with the Jacobi apply
// and both dot products; direction update), all scalars live on the device, reductions are
// fixed-order ("last block" pattern, no FP atomics) and the host only polls a flag every
// `check_every` iterations. Kernels of already-converged iterations exit at their first instruction.
#include <math.h>
#include <string.h>

#include <cooperative_groups.h>
#include <cub/device/device_scan.cuh>

#include "ebe.cuh"

namespace tfem {

__global__ void k_spmv_plan(int64_t n_rows, int64_t n_chunks, const int64_t* __restrict__ indptr,
                            int32_t* __restrict__ chunk_rows) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c > n_chunks) return;
  if (c == n_chunks) {
    chunk_rows[c] = (int32_t)n_rows;
    return;
  }
  const int64_t target = c * (int64_t)TFEM_SPMV_CHUNK;  // first row with indptr[row] >= target
  int64_t lo = 0, hi = n_rows;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (indptr[mid] < target) lo = mid + 1; else hi = mid;
  }
  chunk_rows[c] = (int32_t)lo;
}

namespace {

struct Csr {
  int64_t n, nnz, n_chunks;
  const int64_t* indptr;
  const int32_t* cols;
  const double* vals;
  const int32_t* chunk_rows;
};

// ------------------------------------------------------------------------------------------ SpMV
template <int G, bool DOT>
__global__ void __launch_bounds__(kSpmvWarps * 32)
    k_spmv(Csr A, const double* __restrict__ x, double* __restrict__ y, const double* sc,
           double* partials, unsigned int* ticket, double* out_scalar) {
  __shared__ __align__(16) double s_prod[kSpmvWarps][kSpmvCap];
  __shared__ double s_red[kSpmvWarps];
  if (DOT && sc[SC_DONE] != 0.0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double dot = 0.0;
  for (int64_t c = (int64_t)blockIdx.x * kSpmvWarps + warp; c < A.n_chunks;
       c += (int64_t)gridDim.x * kSpmvWarps)
    dot += spmv_chunk<G>(c, A.indptr, A.cols, A.vals, A.chunk_rows, x, y, DOT ? x : nullptr,
                         s_prod[warp], lane);
  if (DOT) {
    const double b = block_sum<kSpmvWarps * 32>(dot, s_red);
    double mine[1] = {b}, tot[1];
    if (publish_and_reduce<1>(mine, partials, ticket, tot) && threadIdx.x == 0) *out_scalar = tot[0];
  }
}

template <bool DOT>
int launch_spmv(const Csr& A, const double* x, double* y, const double* sc, double* partials,
                unsigned int* ticket, double* out_scalar, int grid, cudaStream_t st) {
  const double avg = A.n > 0 ? (double)A.nnz / (double)A.n : 0.0;
  if (avg >= 48.0)
    k_spmv<16, DOT><<<grid, kSpmvWarps * 32, 0, st>>>(A, x, y, sc, partials, ticket, out_scalar);
  else if (avg >= 12.0)
    k_spmv<8, DOT><<<grid, kSpmvWarps * 32, 0, st>>>(A, x, y, sc, partials, ticket, out_scalar);
  else
    k_spmv<4, DOT><<<grid, kSpmvWarps * 32, 0, st>>>(A, x, y, sc, partials, ticket, out_scalar);
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

inline int spmv_grid(int64_t n_chunks) {
  const int64_t want = (n_chunks + kSpmvWarps - 1) / kSpmvWarps;
  const int64_t cap = (int64_t)(num_sms() < kSMs ? num_sms() : kSMs) * 6;  // 6 CTAs x 32 KB smem per SM
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

// ------------------------------------------------------------------------------------------ SELL-32
// (layout and per-slice device code: sell.cuh)
__global__ void k_sell_widths(int64_t n, int64_t n_slices, const int64_t* __restrict__ indptr,
                              int64_t* __restrict__ slice_elems, int64_t long_cap) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (t >= n_slices) return;
  const int64_t r = t * 32 + lane;
  int64_t len64 = (r < n) ? indptr[r + 1] - indptr[r] : 0;
  if (len64 > long_cap) len64 = 0;   // long rows stay out of the slices (side path k_sell_long)
  int len = (int)len64;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
  if (lane == 0) slice_elems[t] = (int64_t)((len + 1) & ~1) * 32;
}

// one warp per slice; output-coalesced transposition CSR -> SELL (cols and/or vals)
__global__ void k_sell_fill(int64_t n, int64_t n_cols, int64_t n_slices, const int64_t* __restrict__ indptr,
                            const int32_t* __restrict__ cols, const double* __restrict__ vals,
                            const int64_t* __restrict__ slice_ptr, int32_t* __restrict__ s_cols,
                            double* __restrict__ s_vals, int64_t long_cap) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (t >= n_slices) return;
  const int64_t base = slice_ptr[t];
  const int total = (int)(slice_ptr[t + 1] - base);  // 32 * W
  const int64_t my_row = t * 32 + lane;
  const int64_t my_beg = (my_row < n) ? indptr[my_row] : 0;
  int64_t my_len64 = (my_row < n) ? indptr[my_row + 1] - my_beg : 0;
  if (my_len64 > long_cap) my_len64 = 0;   // long rows stay out of the slices (side path k_sell_long)
  const int my_len = (int)my_len64;
  for (int o = lane; o < total; o += 32) {
    const int s2 = o >> 6, rem = o & 63;
    const int rl = rem >> 1, k = s2 * 2 + (rem & 1);
    const int64_t beg = __shfl_sync(0xffffffffu, my_beg, rl);
    const int len = __shfl_sync(0xffffffffu, my_len, rl);
    const bool real = k < len;
    if (s_cols) {
      int64_t own = t * 32 + rl;
      if (own >= n) own = n - 1;
      if (own >= n_cols) own = n_cols - 1;  // rectangular operators (AMG prolongation / restriction)
      s_cols[base + o] = real ? cols[beg + k] : (int32_t)own;
    }
    if (s_vals) s_vals[base + o] = real ? vals[beg + k] : 0.0;
  }
}

__global__ void k_bsell_widths(int64_t n_slices, int dpn, int nps, const int64_t* __restrict__ slice_ptr,
                               int64_t* __restrict__ out) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_slices) return;
  const int64_t W = (slice_ptr[t + 1] - slice_ptr[t]) >> 5;
  out[t] = ((W + dpn - 1) / dpn) * nps;
}

__global__ void k_bsell_fill(int64_t n_slices, int64_t n_nod, int dpn, int nps,
                             const int64_t* __restrict__ node_ptr, const int32_t* __restrict__ adj,
                             const int64_t* __restrict__ bslice_ptr, int32_t* __restrict__ bcols) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (t >= n_slices) return;
  const int64_t base = bslice_ptr[t];
  const int total = (int)(bslice_ptr[t + 1] - base);
  const int64_t nf = (t * 32) / dpn;
  for (int o = lane; o < total; o += 32) {
    const int kb = o / nps, m = o - kb * nps;
    int64_t node = nf + m;
    if (node >= n_nod) node = n_nod - 1;
    const int64_t nb = node_ptr[node];
    const int cnt = (int)(node_ptr[node + 1] - nb);
    // an empty row (halo rows of a distributed operator) pads with block column 0: always inside the operand, also for
    // rectangular operators (its own node index may lie beyond a shorter coarse vector); the values there are 0.0
    bcols[base + o] = cnt > 0 ? adj[nb + (kb < cnt ? kb : cnt - 1)] : 0;
  }
}

// The linear operator of a Krylov solve: the assembled matrix in SELL-32 form, or the matrix-free element
// operator (ebe.cuh).
struct Op {
  bool ebe = false;
  Sell sell;
  Ebe el;
  int64_t n = 0;
};

template <bool DOT>
int apply_op(const Op& op, const double* x, double* y, const double* sc, double* partials, unsigned int* ticket,
             double* out_scalar, cudaStream_t st) {
  if (op.ebe) return launch_ebe<DOT>(op.el, x, y, sc, partials, ticket, out_scalar, st);
  return launch_sell<DOT>(op.sell, x, y, sc, partials, ticket, out_scalar, st);
}

// ------------------------------------------------------------------------------------------ CG
// Scalar recurrences. On one GPU the last CTA of the producing kernel runs them in place; across ranks
// the producing kernel only writes its local sums to `red`, the host all-reduces `red` (NCCL) and
// k_cg_scalars runs the same code on the reduced values, identically on every rank.
__device__ __forceinline__ void cg_scalars_init(double* sc, const double* tot, double rtol, double atol) {
  const double bnrm = sqrt(tot[2]);
  const double tol = fmax(atol, rtol * bnrm);
  sc[SC_RR] = tot[0];
  sc[SC_RHO] = tot[1];
  sc[SC_RHO_PREV] = tot[1];
  sc[SC_BNRM] = bnrm;
  sc[SC_TOL] = tol;
  sc[SC_ITERS] = 0.0;
  sc[SC_DONE] = (bnrm == 0.0 || sqrt(tot[0]) < tol) ? 1.0 : 0.0;
}

__device__ __forceinline__ void cg_scalars_update(double* sc, const double* tot) {
  const double rho_prev = sc[SC_RHO];
  sc[SC_ALPHA] = rho_prev / sc[SC_PQ];
  sc[SC_RHO_PREV] = rho_prev;
  sc[SC_RHO] = tot[1];
  sc[SC_RR] = tot[0];
  sc[SC_BETA] = tot[1] / rho_prev;
  sc[SC_ITERS] += 1.0;
  if (!isfinite(tot[0])) sc[SC_DONE] = 2.0;  // breakdown (scipy would iterate on NaNs to maxiter)
  else if (sqrt(tot[0]) < sc[SC_TOL]) sc[SC_DONE] = 1.0;
}

// which: 0 after init (red = rr, rho, bb), 1 after SpMV (red = p.q), 2 after update (red = rr, rho)
__global__ void k_cg_scalars(int which, double* sc, const double* red, double rtol, double atol) {
  if (which == 0) cg_scalars_init(sc, red, rtol, atol);
  else if (sc[SC_DONE] != 0.0) return;
  else if (which == 1) sc[SC_PQ] = red[0];
  else cg_scalars_update(sc, red);
}

// r = b - q (q = A x0) or r = b ; p = z = dinv*r ; rr = r.r ; rho = r.z ; bb = b.b
__global__ void __launch_bounds__(kVecThreads)
    k_cg_init(int64_t n, const double* __restrict__ b, const double* __restrict__ q_or_null,
              const double* __restrict__ dinv, double* __restrict__ r, double* __restrict__ p,
              double* sc, double rtol, double atol, double* partials, unsigned int* ticket,
              double* red) {
  __shared__ double s_red[kVecThreads / 32];
  double rr = 0.0, rho = 0.0, bb = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * kVecThreads) {
    const double bi = b[i];
    const double ri = q_or_null ? bi - q_or_null[i] : bi;
    const double zi = dinv[i] * ri;
    r[i] = ri;
    p[i] = zi;
    rr += ri * ri;
    rho += ri * zi;
    bb += bi * bi;
  }
  double mine[3], tot[3];
  mine[0] = block_sum<kVecThreads>(rr, s_red);
  mine[1] = block_sum<kVecThreads>(rho, s_red);
  mine[2] = block_sum<kVecThreads>(bb, s_red);
  if (publish_and_reduce<3>(mine, partials, ticket, tot) && threadIdx.x == 0) {
    if (red) { red[0] = tot[0]; red[1] = tot[1]; red[2] = tot[2]; }
    else cg_scalars_init(sc, tot, rtol, atol);
  }
}

// alpha = rho / p.q ; x += alpha p ; r -= alpha q ; rr = r.r ; rho' = r.(dinv r) ; convergence test
__global__ void __launch_bounds__(kVecThreads)
    k_cg_update(int64_t n, const double* __restrict__ p, const double* __restrict__ q,
                const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r,
                double* sc, double* partials, unsigned int* ticket, double* red) {
  __shared__ double s_red[kVecThreads / 32];
  if (sc[SC_DONE] != 0.0) return;
  const double pq = sc[SC_PQ];
  const double alpha = sc[SC_RHO] / pq;
  double rr = 0.0, rho = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * kVecThreads) {
    const double pi = p[i], qi = q[i];
    x[i] = fma(alpha, pi, x[i]);
    const double ri = fma(-alpha, qi, r[i]);
    r[i] = ri;
    rr = fma(ri, ri, rr);
    rho = fma(ri * dinv[i], ri, rho);
  }
  double mine[2], tot[2];
  mine[0] = block_sum<kVecThreads>(rr, s_red);
  mine[1] = block_sum<kVecThreads>(rho, s_red);
  if (publish_and_reduce<2>(mine, partials, ticket, tot) && threadIdx.x == 0) {
    if (red) { red[0] = tot[0]; red[1] = tot[1]; }
    else cg_scalars_update(sc, tot);
  }
}

// p = dinv*r + beta p
__global__ void __launch_bounds__(kVecThreads)
    k_cg_direction(int64_t n, const double* __restrict__ r, const double* __restrict__ dinv,
                   double* __restrict__ p, const double* sc) {
  if (sc[SC_DONE] != 0.0) return;
  const double beta = sc[SC_BETA];
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * kVecThreads)
    p[i] = fma(beta, p[i], dinv[i] * r[i]);
}

// ---- small systems: the whole CG loop in ONE cooperative kernel.
// Below ~1 M rows an iteration of the three-kernel driver costs ~38 us whatever the size: three dependent launches,
// each ending in a "last CTA" reduction. Here the grid stays resident and the phases of an iteration are separated
// by grid-wide barriers (~2-3 us each); every CTA sums the per-CTA partials itself in the same fixed order, so all
// CTAs see bit-identical scalars and take the same decisions without a broadcast. A thread owns the same rows in
// every phase (row = 32 t + lane of its slices), so x, r, q and its entries of p are only ever read by their writer;
// the one cross-thread read, the gather of p inside the SpMV, goes to L2 (ld.global.cg: an L1 line could predate
// another CTA's store of the same iteration).
__device__ __forceinline__ double grid_total(const double* partials, int nb, double* s_bcast) {
  __syncthreads();  // s_bcast of the previous reduction has been consumed
  if (threadIdx.x < 32) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nb; i += 32) s += __ldcg(partials + i);
    s = warp_sum(s);
    if (threadIdx.x == 0) *s_bcast = s;
  }
  __syncthreads();
  return *s_bcast;
}

template <int DPN>
__global__ void __launch_bounds__(kSellWarps * 32, 4)
    k_cg_coop(Sell A, const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r, double* p,
              double* __restrict__ q, double* sc, double* partials, int batch) {
  namespace cgr = cooperative_groups;
  cgr::grid_group grid = cgr::this_grid();
  __shared__ double s_red[kSellWarps];
  __shared__ double s_bcast;
  const int nb = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t t0 = (int64_t)blockIdx.x * kSellWarps + warp, stride = (int64_t)nb * kSellWarps;
  if (sc[SC_DONE] != 0.0) return;  // read-only so far: uniform over the grid
  double rho = sc[SC_RHO], iters = sc[SC_ITERS], rr_tot = sc[SC_RR], done = 0.0;
  const double tol = sc[SC_TOL];
  for (int it = 0; it < batch; ++it) {
    // q = A p on my slices, partial p.q
    double dot = 0.0;
    for (int64_t t = t0; t < A.n_slices; t += stride) {
      const int64_t row = t * 32 + lane;
      const double acc = slice_row<DPN, true>(A, t, p, lane);
      if (row < A.n) {
        q[row] = acc;
        dot = fma(acc, p[row], dot);
      }
    }
    double bs = block_sum<kSellWarps * 32>(dot, s_red);
    if (threadIdx.x == 0) partials[blockIdx.x] = bs;
    grid.sync();
    const double alpha = rho / grid_total(partials, nb, &s_bcast);
    // x += alpha p ; r -= alpha q ; partial r.r and r.(D^-1 r)
    double rr = 0.0, rz = 0.0;
    for (int64_t t = t0; t < A.n_slices; t += stride) {
      const int64_t row = t * 32 + lane;
      if (row < A.n) {
        x[row] = fma(alpha, p[row], x[row]);
        const double ri = fma(-alpha, q[row], r[row]);
        r[row] = ri;
        rr = fma(ri, ri, rr);
        rz = fma(ri * dinv[row], ri, rz);
      }
    }
    bs = block_sum<kSellWarps * 32>(rr, s_red);
    if (threadIdx.x == 0) partials[nb + blockIdx.x] = bs;
    bs = block_sum<kSellWarps * 32>(rz, s_red);
    if (threadIdx.x == 0) partials[2 * nb + blockIdx.x] = bs;
    grid.sync();
    rr_tot = grid_total(partials + nb, nb, &s_bcast);
    const double rz_tot = grid_total(partials + 2 * nb, nb, &s_bcast);
    const double beta = rz_tot / rho;
    rho = rz_tot;
    iters += 1.0;
    if (!isfinite(rr_tot)) done = 2.0;  // breakdown (scipy would iterate on NaNs to maxiter)
    else if (sqrt(rr_tot) < tol) done = 1.0;
    if (done != 0.0) break;  // identical in every CTA
    // p = D^-1 r + beta p
    for (int64_t t = t0; t < A.n_slices; t += stride) {
      const int64_t row = t * 32 + lane;
      if (row < A.n) p[row] = fma(beta, p[row], dinv[row] * r[row]);
    }
    grid.sync();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sc[SC_RHO] = rho;
    sc[SC_RR] = rr_tot;
    sc[SC_ITERS] = iters;
    sc[SC_DONE] = done;
  }
}

constexpr int64_t kCoopMaxRows = 1000000;

// Launches batches of a cooperative kernel until convergence / maxiter. Returns TFEM_ERR_INVALID (and touches nothing)
// if the device cannot launch cooperatively, so that the caller falls back to the kernel-per-phase driver.
template <typename K>
int coop_grid(K kernel, const Sell& A, int* grid_out) {
  int dev = 0, attr = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&attr, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !attr)
    return TFEM_ERR_INVALID;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kSellWarps * 32, 0) != cudaSuccess || per_sm < 1)
    return TFEM_ERR_INVALID;
  const int64_t want = (A.n_slices + kSellWarps - 1) / kSellWarps;
  const int64_t cap = (int64_t)num_sms() * per_sm;
  const int64_t grid = want < cap ? want : cap;
  if (3 * grid > kMaxPartials) return TFEM_ERR_INVALID;
  *grid_out = (int)grid;
  return TFEM_OK;
}

template <int DPN>
int cg_coop_batches(const Sell& A, const double* dinv, double* x, const Work& w, int64_t maxiter, double* sc_host,
                    double* launches, double* spmvs, cudaStream_t st) {
  int g = 0;
  if (coop_grid(k_cg_coop<DPN>, A, &g) != TFEM_OK) return TFEM_ERR_INVALID;
  int64_t issued = 0;
  while (true) {
    TFEM_CUDA(cudaMemcpyAsync(sc_host, w.sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, st));
    TFEM_CUDA(cudaStreamSynchronize(st));
    if (sc_host[SC_DONE] != 0.0 || issued >= maxiter) break;
    int batch = (int)(maxiter - issued < 256 ? maxiter - issued : 256);
    Sell a = A;
    const double* dv = dinv;
    double *xx = x, *rr = w.r, *pp = w.p, *qq = w.q, *scp = w.sc, *part = w.partials;
    void* args[] = {&a, &dv, &xx, &rr, &pp, &qq, &scp, &part, &batch};
    TFEM_CUDA(cudaLaunchCooperativeKernel((void*)k_cg_coop<DPN>, dim3(g), dim3(kSellWarps * 32), args, 0, st));
    issued += batch;
    *launches += 1;
  }
  *spmvs += sc_host[SC_ITERS];
  return TFEM_OK;
}

__global__ void k_copy_or_zero(int64_t n, const double* __restrict__ src, double* __restrict__ dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src ? src[i] : 0.0;
}

// ------------------------------------------------------------------------------------------ MINRES
// Paige-Saunders MINRES with a diagonal preconditioner, restating scipy `minres` (the algorithm
// behind cupy_minres, sparse.py:411). Vector roles: r1, r2, y, v, w, w1(w_old), w2(w_cur).
//
// init: r1 = b - A x0 (or b); y = dinv*r1; beta1^2 = r1.y ; r2 = r1
__global__ void __launch_bounds__(kVecThreads)
    k_mr_init(int64_t n, const double* __restrict__ b, const double* __restrict__ q_or_null,
              const double* __restrict__ dinv, double* __restrict__ r1, double* __restrict__ r2,
              double* __restrict__ y, double* __restrict__ w, double* __restrict__ w2, double* sc,
              double* partials, unsigned int* ticket) {
  __shared__ double s_red[kVecThreads / 32];
  double ry = 0.0, bb = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * kVecThreads) {
    const double bi = b[i];
    const double ri = q_or_null ? bi - q_or_null[i] : bi;
    const double yi = dinv[i] * ri;
    r1[i] = ri;
    r2[i] = ri;
    y[i] = yi;
    w[i] = 0.0;
    w2[i] = 0.0;
    ry += ri * yi;
    bb += bi * bi;
  }
  double mine[2], tot[2];
  mine[0] = block_sum<kVecThreads>(ry, s_red);
  mine[1] = block_sum<kVecThreads>(bb, s_red);
  if (publish_and_reduce<2>(mine, partials, ticket, tot) && threadIdx.x == 0) {
    const double beta1 = sqrt(fmax(tot[0], 0.0));
    sc[SC_BNRM] = sqrt(tot[1]);
    sc[SC_M_BETA1] = beta1;
    sc[SC_M_OLDB] = 0.0;
    sc[SC_M_BETA] = beta1;
    sc[SC_M_DBAR] = 0.0;
    sc[SC_M_EPSLN] = 0.0;
    sc[SC_M_PHIBAR] = beta1;
    sc[SC_M_CS] = -1.0;
    sc[SC_M_SN] = 0.0;
    sc[SC_M_TNORM2] = 0.0;
    sc[SC_M_GMAX] = 0.0;
    sc[SC_M_GMIN] = 1.7976931348623157e308;
    sc[SC_M_YNORM2] = 0.0;
    sc[SC_M_RNORM] = beta1;
    sc[SC_M_ISTOP] = 0.0;
    sc[SC_ITERS] = 0.0;
    // beta1 == 0 -> x0 is the solution; ||b|| == 0 -> x = b (scipy returns early in both cases)
    double done = 0.0;
    if (tot[0] < 0.0) done = 2.0;
    else if (beta1 == 0.0 || tot[1] == 0.0) done = 1.0;
    sc[SC_DONE] = done;
  }
}

// v = y / beta  (v is what the SpMV consumes)
__global__ void __launch_bounds__(kVecThreads)
    k_mr_v(int64_t n, const double* __restrict__ y, double* __restrict__ v, const double* sc) {
  if (sc[SC_DONE] != 0.0) return;
  const double s = 1.0 / sc[SC_M_BETA];
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * kVecThreads)
    v[i] = s * y[i];
}

// scalar recurrences of one MINRES iteration (scipy minres.py main loop), on the scalar block `sc` (global memory in
// the kernel-per-phase driver, a per-CTA shared-memory copy in the cooperative kernel); ry = r2 . (D^-1 r2)
__device__ __forceinline__ void mr_scalars_lanczos(double* sc, double ry) {
    const double eps = 2.220446049250313e-16;
    const double itn = sc[SC_ITERS] + 1.0;
    const double alfa = sc[SC_M_ALFA];
    const double oldb = sc[SC_M_BETA];
    if (ry < 0.0 || !isfinite(ry)) {
      sc[SC_DONE] = 2.0;
      return;
    }
    const double beta = sqrt(ry);
    const double beta1 = sc[SC_M_BETA1];
    double tnorm2 = sc[SC_M_TNORM2] + alfa * alfa + oldb * oldb + beta * beta;
    double istop = 0.0;
    if (itn == 1.0 && beta / beta1 <= 10.0 * eps) istop = -1.0;
    const double cs0 = sc[SC_M_CS], sn0 = sc[SC_M_SN], dbar0 = sc[SC_M_DBAR];
    const double oldeps = sc[SC_M_EPSLN];
    const double delta = cs0 * dbar0 + sn0 * alfa;
    const double gbar = sn0 * dbar0 - cs0 * alfa;
    const double epsln = sn0 * beta;
    const double dbar = -cs0 * beta;
    const double root = hypot(gbar, dbar);
    double gamma = hypot(gbar, beta);
    gamma = fmax(gamma, eps);
    const double cs = gbar / gamma, sn = beta / gamma;
    const double phibar0 = sc[SC_M_PHIBAR];
    const double phi = cs * phibar0;
    const double phibar = sn * phibar0;
    sc[SC_M_OLDB] = oldb;
    sc[SC_M_BETA] = beta;
    sc[SC_M_TNORM2] = tnorm2;
    sc[SC_M_DBAR] = dbar;
    sc[SC_M_EPSLN] = epsln;
    sc[SC_M_CS] = cs;
    sc[SC_M_SN] = sn;
    sc[SC_M_PHIBAR] = phibar;
    sc[SC_M_PHI] = phi;
    sc[SC_M_DENOM] = 1.0 / gamma;
    sc[SC_M_OLDEPS] = oldeps;
    sc[SC_M_DELTA] = delta;
    const double gmax = fmax(sc[SC_M_GMAX], gamma), gmin = fmin(sc[SC_M_GMIN], gamma);
    sc[SC_M_GMAX] = gmax;
    sc[SC_M_GMIN] = gmin;
    sc[SC_M_RNORM] = phibar;
    sc[SC_ITERS] = itn;
    sc[SC_M_ISTOP] = istop;
    // the ||x||-dependent stopping tests are finished in k_mr_xupdate (needs the new x)
    sc[SC_RHO] = root;  // reused slot: root for test2
}

// the ||x||-dependent stopping tests of scipy minres; xx = x . x
__device__ __forceinline__ void mr_scalars_xupdate(double* sc, double xx, double rtol, double maxiter) {
    const double eps = 2.220446049250313e-16;
    const double Anorm = sqrt(sc[SC_M_TNORM2]);
    const double ynorm = sqrt(xx);
    const double epsx = Anorm * ynorm * eps;
    const double rnorm = sc[SC_M_RNORM];
    const double root = sc[SC_RHO];
    const double test1 = (ynorm == 0.0 || Anorm == 0.0) ? INFINITY : rnorm / (Anorm * ynorm);
    const double test2 = (Anorm == 0.0) ? INFINITY : root / Anorm;
    const double Acond = sc[SC_M_GMAX] / sc[SC_M_GMIN];
    double istop = sc[SC_M_ISTOP];
    if (istop == 0.0) {
      const double t1 = 1.0 + test1, t2 = 1.0 + test2;
      if (t2 <= 1.0) istop = 2.0;
      if (t1 <= 1.0) istop = 1.0;
      if (sc[SC_ITERS] >= maxiter) istop = 6.0;
      if (Acond >= 0.1 / eps) istop = 4.0;
      if (epsx >= sc[SC_M_BETA1]) istop = 3.0;
      if (test2 <= rtol) istop = 2.0;
      if (test1 <= rtol) istop = 1.0;
    }
    sc[SC_M_ISTOP] = istop;
    sc[SC_RR] = rnorm * rnorm;
    if (istop != 0.0) sc[SC_DONE] = (istop == 6.0) ? 3.0 : 1.0;
}

// after y = A v and alfa_raw = v.(A v) (from the SpMV kernel, stored in SC_PQ):
//   y -= (beta/oldb) r1 (itn>=2) ; alfa = v.y ; y -= (alfa/beta) r2 ; r1 = r2 ; r2 = y ; y = dinv*r2 ;
//   beta_new^2 = r2.y
// v.r1 = 0 in exact arithmetic but scipy computes alfa AFTER the r1 correction, so alfa is reduced here
// in a first pass (k_mr_alfa) and applied in the second (k_mr_lanczos).
__global__ void __launch_bounds__(kVecThreads)
    k_mr_alfa(int64_t n, const double* __restrict__ v, double* __restrict__ y,
              const double* __restrict__ r1, double* sc, double* partials, unsigned int* ticket) {
  __shared__ double s_red[kVecThreads / 32];
  if (sc[SC_DONE] != 0.0) return;
  const bool second = sc[SC_ITERS] >= 1.0;
  const double f = second ? sc[SC_M_BETA] / sc[SC_M_OLDB] : 0.0;
  double a = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * kVecThreads) {
    double yi = y[i];
    if (second) {
      yi = yi - f * r1[i];
      y[i] = yi;
    }
    a += v[i] * yi;
  }
  double mine[1], tot[1];
  mine[0] = block_sum<kVecThreads>(a, s_red);
  if (publish_and_reduce<1>(mine, partials, ticket, tot) && threadIdx.x == 0) sc[SC_M_ALFA] = tot[0];
}

__global__ void __launch_bounds__(kVecThreads)
    k_mr_lanczos(int64_t n, const double* __restrict__ dinv, double* __restrict__ y,
                 double* __restrict__ r1, double* __restrict__ r2, double* sc, double rtol,
                 double* partials, unsigned int* ticket) {
  __shared__ double s_red[kVecThreads / 32];
  if (sc[SC_DONE] != 0.0) return;
  const double f = sc[SC_M_ALFA] / sc[SC_M_BETA];
  double ry = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * kVecThreads) {
    const double r2i = r2[i];
    const double yi = y[i] - f * r2i;
    r1[i] = r2i;
    r2[i] = yi;
    const double zi = dinv[i] * yi;
    y[i] = zi;
    ry += yi * zi;
  }
  double mine[1], tot[1];
  mine[0] = block_sum<kVecThreads>(ry, s_red);
  if (publish_and_reduce<1>(mine, partials, ticket, tot) && threadIdx.x == 0) mr_scalars_lanczos(sc, tot[0]);
}

// w = (v - oldeps*w1 - delta*w2)/gamma ; x += phi*w ; ynorm^2 = x.x ; then the stopping tests
__global__ void __launch_bounds__(kVecThreads)
    k_mr_xupdate(int64_t n, const double* __restrict__ v, double* __restrict__ w1,
                 double* __restrict__ w2, double* __restrict__ x, double* sc, double rtol,
                 double maxiter, double* partials, unsigned int* ticket) {
  __shared__ double s_red[kVecThreads / 32];
  if (sc[SC_DONE] != 0.0) return;
  const double oldeps = sc[SC_M_OLDEPS], delta = sc[SC_M_DELTA], denom = sc[SC_M_DENOM];
  const double phi = sc[SC_M_PHI];
  double xx = 0.0;
  // storage rotation: w1 <- w2(old), w2 <- w(new): w1[i] holds w_{k-2}, w2[i] holds w_{k-1}
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * kVecThreads) {
    const double wk2 = w1[i], wk1 = w2[i];
    const double wn = (v[i] - oldeps * wk2 - delta * wk1) * denom;
    w1[i] = wk1;
    w2[i] = wn;
    const double xi = fma(phi, wn, x[i]);
    x[i] = xi;
    xx = fma(xi, xi, xx);
  }
  double mine[1], tot[1];
  mine[0] = block_sum<kVecThreads>(xx, s_red);
  if (publish_and_reduce<1>(mine, partials, ticket, tot) && threadIdx.x == 0) mr_scalars_xupdate(sc, tot[0], rtol, maxiter);
}

// MINRES the same way: one cooperative kernel, three grid barriers per iteration (after the SpMV + alfa pass, after
// the Lanczos pass, after the x update). Every CTA carries its own copy of the scalar recurrences in shared memory and
// advances it from the same reduction totals, so no scalar ever has to be broadcast; CTA 0 writes the block back at
// the end. v ping-pongs between two buffers: the x update of iteration k also writes v_{k+1} = y / beta_{k+1}, so the
// barrier that ends the iteration is also the one the next SpMV's gather needs.
template <int DPN>
__global__ void __launch_bounds__(kSellWarps * 32, 4)
    k_mr_coop(Sell A, const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r1,
              double* __restrict__ r2, double* __restrict__ y, double* va, double* vb, double* __restrict__ w1,
              double* __restrict__ w2, double* sc, double* partials, double rtol, double maxiter, int batch) {
  namespace cgr = cooperative_groups;
  cgr::grid_group grid = cgr::this_grid();
  __shared__ double s_red[kSellWarps];
  __shared__ double s_bcast;
  __shared__ double s_sc[SC_COUNT];
  const int nb = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t t0 = (int64_t)blockIdx.x * kSellWarps + warp, stride = (int64_t)nb * kSellWarps;
  if (threadIdx.x < SC_COUNT) s_sc[threadIdx.x] = sc[threadIdx.x];
  __syncthreads();
  if (s_sc[SC_DONE] != 0.0) return;  // uniform over the grid
  double* v_cur = va;
  double* v_nxt = vb;
  {  // v = y / beta for the first iteration of this launch
    const double s0 = 1.0 / s_sc[SC_M_BETA];
    for (int64_t t = t0; t < A.n_slices; t += stride) {
      const int64_t row = t * 32 + lane;
      if (row < A.n) v_cur[row] = s0 * y[row];
    }
    grid.sync();
  }
  for (int it = 0; it < batch; ++it) {
    // y = A v ; y -= (beta/oldb) r1 (from the second iteration on) ; partial v.y
    const bool second = s_sc[SC_ITERS] >= 1.0;
    const double f = second ? s_sc[SC_M_BETA] / s_sc[SC_M_OLDB] : 0.0;
    double a = 0.0;
    for (int64_t t = t0; t < A.n_slices; t += stride) {
      const int64_t row = t * 32 + lane;
      const double acc = slice_row<DPN, true>(A, t, v_cur, lane);
      if (row < A.n) {
        const double yi = second ? acc - f * r1[row] : acc;
        y[row] = yi;
        a += v_cur[row] * yi;
      }
    }
    double bs = block_sum<kSellWarps * 32>(a, s_red);
    if (threadIdx.x == 0) partials[blockIdx.x] = bs;
    grid.sync();
    const double alfa = grid_total(partials, nb, &s_bcast);
    // y -= (alfa/beta) r2 ; r1 = r2 ; r2 = y ; y = D^-1 r2 ; partial r2.y
    const double f2 = alfa / s_sc[SC_M_BETA];
    double ry = 0.0;
    for (int64_t t = t0; t < A.n_slices; t += stride) {
      const int64_t row = t * 32 + lane;
      if (row < A.n) {
        const double r2i = r2[row];
        const double yi = y[row] - f2 * r2i;
        r1[row] = r2i;
        r2[row] = yi;
        const double zi = dinv[row] * yi;
        y[row] = zi;
        ry += yi * zi;
      }
    }
    bs = block_sum<kSellWarps * 32>(ry, s_red);
    if (threadIdx.x == 0) partials[nb + blockIdx.x] = bs;
    grid.sync();
    const double ry_tot = grid_total(partials + nb, nb, &s_bcast);
    if (threadIdx.x == 0) {
      s_sc[SC_M_ALFA] = alfa;
      mr_scalars_lanczos(s_sc, ry_tot);
    }
    __syncthreads();
    if (s_sc[SC_DONE] != 0.0) break;  // breakdown; identical in every CTA
    // w = (v - oldeps w1 - delta w2) / gamma ; x += phi w ; partial x.x ; next v = y / beta
    const double oldeps = s_sc[SC_M_OLDEPS], delta = s_sc[SC_M_DELTA], denom = s_sc[SC_M_DENOM];
    const double phi = s_sc[SC_M_PHI], sv = 1.0 / s_sc[SC_M_BETA];
    double xx = 0.0;
    for (int64_t t = t0; t < A.n_slices; t += stride) {
      const int64_t row = t * 32 + lane;
      if (row < A.n) {
        const double wk2 = w1[row], wk1 = w2[row];
        const double wn = (v_cur[row] - oldeps * wk2 - delta * wk1) * denom;
        w1[row] = wk1;
        w2[row] = wn;
        const double xi = fma(phi, wn, x[row]);
        x[row] = xi;
        xx = fma(xi, xi, xx);
        v_nxt[row] = sv * y[row];
      }
    }
    bs = block_sum<kSellWarps * 32>(xx, s_red);
    if (threadIdx.x == 0) partials[2 * nb + blockIdx.x] = bs;
    grid.sync();
    const double xx_tot = grid_total(partials + 2 * nb, nb, &s_bcast);
    if (threadIdx.x == 0) mr_scalars_xupdate(s_sc, xx_tot, rtol, maxiter);
    __syncthreads();
    if (s_sc[SC_DONE] != 0.0) break;
    double* tmp = v_cur;
    v_cur = v_nxt;
    v_nxt = tmp;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < SC_COUNT) sc[threadIdx.x] = s_sc[threadIdx.x];
}

template <int DPN>
int mr_coop_batches(const Sell& A, const double* dinv, double* x, const Work& w, double rtol, int64_t maxiter,
                    double* sc_host, double* launches, double* spmvs, cudaStream_t st) {
  int g = 0;
  if (coop_grid(k_mr_coop<DPN>, A, &g) != TFEM_OK) return TFEM_ERR_INVALID;
  double* vb = nullptr;  // second v buffer (ping-pong)
  TFEM_CUDA(malloc_async(&vb, pad32(A.n) * sizeof(double), st));
  int64_t issued = 0;
  int rc = TFEM_OK;
  while (true) {
    if ((rc = check_cuda(cudaMemcpyAsync(sc_host, w.sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, st), "memcpy")))
      break;
    if ((rc = check_cuda(cudaStreamSynchronize(st), "sync"))) break;
    if (sc_host[SC_DONE] != 0.0 || issued >= maxiter) break;
    int batch = (int)(maxiter - issued < 256 ? maxiter - issued : 256);
    Sell a = A;
    const double* dv = dinv;
    double *xx = x, *r1 = w.r1, *r2 = w.r2, *yy = w.y, *va = w.v, *w1 = w.w1, *w2 = w.w2, *scp = w.sc, *part = w.partials;
    double rt = rtol, mi = (double)maxiter;
    void* args[] = {&a, &dv, &xx, &r1, &r2, &yy, &va, &vb, &w1, &w2, &scp, &part, &rt, &mi, &batch};
    if ((rc = check_cuda(cudaLaunchCooperativeKernel((void*)k_mr_coop<DPN>, dim3(g), dim3(kSellWarps * 32), args, 0, st),
                         "cooperative launch")))
      break;
    issued += batch;
    *launches += 1;
  }
  cudaFreeAsync(vb, st);
  if (rc != TFEM_OK) return rc;
  *spmvs += sc_host[SC_ITERS];
  return TFEM_OK;
}

// ------------------------------------------------------------------------------------------ misc
__global__ void k_diag_positions(int64_t n, const int64_t* __restrict__ indptr,
                                 const int32_t* __restrict__ cols, int64_t* __restrict__ pos) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  int64_t lo = indptr[r], hi = indptr[r + 1];
  int64_t found = -1;
  while (lo < hi) {  // columns are sorted within a row
    int64_t mid = (lo + hi) >> 1;
    int32_t c = cols[mid];
    if (c == (int32_t)r) {
      found = mid;
      break;
    }
    if (c < (int32_t)r) lo = mid + 1; else hi = mid;
  }
  pos[r] = found;
}

__global__ void k_jacobi(int64_t n, const double* __restrict__ vals, const int64_t* __restrict__ pos,
                         double* __restrict__ dinv) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int64_t p = pos[r];
  dinv[r] = 1.0 / (p >= 0 ? vals[p] : 0.0);
}

__global__ void k_adjoint_grad(int64_t n, const int64_t* __restrict__ indptr,
                               const int32_t* __restrict__ cols, const double* __restrict__ lam,
                               const double* __restrict__ x, double* __restrict__ g) {
  // one warp per row
  const int lane = threadIdx.x & 31;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= n) return;
  const double ml = -lam[r];
  for (int64_t k = indptr[r] + lane; k < indptr[r + 1]; k += 32) g[k] = ml * x[cols[k]];
}

// transpose helpers (integer counting sort; order inside a column fixed by a per-column sort)
__global__ void k_tr_count(int64_t nnz, const int32_t* __restrict__ cols, int64_t* __restrict__ cnt) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k < nnz) atomicAdd((unsigned long long*)&cnt[cols[k] + 1], 1ull);
}
__global__ void k_tr_scan_serial(int64_t n_cols, int64_t* cnt) {  // tiny matrices only; see host code
  for (int64_t i = 0; i < n_cols; ++i) cnt[i + 1] += cnt[i];
}
__global__ void k_tr_fill(int64_t n_rows, const int64_t* __restrict__ indptr,
                          const int32_t* __restrict__ cols, const double* __restrict__ vals,
                          const int64_t* __restrict__ t_indptr, int64_t* __restrict__ cursor,
                          int32_t* __restrict__ t_cols, double* __restrict__ t_vals) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  for (int64_t k = indptr[r]; k < indptr[r + 1]; ++k) {
    const int32_t c = cols[k];
    const int64_t dst = t_indptr[c] + (int64_t)atomicAdd((unsigned long long*)&cursor[c], 1ull);
    t_cols[dst] = (int32_t)r;
    t_vals[dst] = vals[k];
  }
}
__global__ void k_tr_sort(int64_t n_cols, const int64_t* __restrict__ t_indptr,
                          int32_t* __restrict__ t_cols, double* __restrict__ t_vals) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  const int64_t b = t_indptr[c], e = t_indptr[c + 1];
  for (int64_t i = b + 1; i < e; ++i) {
    const int32_t kc = t_cols[i];
    const double kv = t_vals[i];
    int64_t j = i - 1;
    while (j >= b && t_cols[j] > kc) {
      t_cols[j + 1] = t_cols[j];
      t_vals[j + 1] = t_vals[j];
      --j;
    }
    t_cols[j + 1] = kc;
    t_vals[j + 1] = kv;
  }
}

}  // namespace
}  // namespace tfem

using namespace tfem;

extern "C" int64_t tfem_spmv_num_chunks(int64_t nnz) {
  return nnz <= 0 ? 1 : (nnz + TFEM_SPMV_CHUNK - 1) / TFEM_SPMV_CHUNK;
}

extern "C" int tfem_spmv_plan(int64_t n_rows, int64_t nnz, const int64_t* indptr, int32_t* chunk_rows,
                              void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(indptr && chunk_rows && n_rows > 0, "spmv_plan: bad arguments");
  TFEM_REQUIRE(n_rows < (int64_t)INT32_MAX, "spmv_plan: n_rows must be < 2^31");
  const int64_t nc = tfem_spmv_num_chunks(nnz);
  k_spmv_plan<<<grid_for(nc + 1, 256), 256, 0, st>>>(n_rows, nc, indptr, chunk_rows);
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

extern "C" int tfem_spmv(int64_t n_rows, int64_t nnz, const int64_t* indptr, const int32_t* cols,
                         const double* vals, const int32_t* chunk_rows, const double* x, double* y,
                         void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(indptr && cols && vals && chunk_rows && x && y, "spmv: null pointer");
  TFEM_REQUIRE(aligned16(cols) && aligned16(vals), "spmv: indices/values must be 16-byte aligned");
  Csr A{n_rows, nnz, tfem_spmv_num_chunks(nnz), indptr, cols, vals, chunk_rows};
  return launch_spmv<false>(A, x, y, nullptr, nullptr, nullptr, nullptr, spmv_grid(A.n_chunks), st);
}

extern "C" int tfem_csr_diag_positions(int64_t n_rows, const int64_t* indptr, const int32_t* cols,
                                       int64_t* pos, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(indptr && cols && pos && n_rows > 0, "diag_positions: bad arguments");
  k_diag_positions<<<grid_for(n_rows, 256), 256, 0, st>>>(n_rows, indptr, cols, pos);
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

extern "C" int tfem_jacobi_setup(int64_t n_rows, const double* vals, const int64_t* pos, double* dinv,
                                 void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(vals && pos && dinv && n_rows > 0, "jacobi_setup: bad arguments");
  k_jacobi<<<grid_for(n_rows, 256), 256, 0, st>>>(n_rows, vals, pos, dinv);
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

extern "C" int tfem_adjoint_matrix_grad(int64_t n_rows, const int64_t* indptr, const int32_t* cols,
                                        const double* lam, const double* x, double* g, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(indptr && cols && lam && x && g && n_rows > 0, "adjoint_matrix_grad: bad arguments");
  k_adjoint_grad<<<grid_for(n_rows * 32, 256), 256, 0, st>>>(n_rows, indptr, cols, lam, x, g);
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

extern "C" int tfem_csr_transpose(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* indptr,
                                  const int32_t* cols, const double* vals, int64_t* t_indptr,
                                  int32_t* t_cols, double* t_vals, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(indptr && cols && vals && t_indptr && t_cols && t_vals, "transpose: null pointer");
  TFEM_CUDA(cudaMemsetAsync(t_indptr, 0, (n_cols + 1) * sizeof(int64_t), st));
  if (nnz == 0) return TFEM_OK;
  k_tr_count<<<grid_for(nnz, 256), 256, 0, st>>>(nnz, cols, t_indptr);
  TFEM_LAUNCH_CHECK();
  // general (non-symmetric) matrices reach this path only through differentiable_sparse_solve on
  // user matrices; FEM tangents are symmetric and never transposed. A serial scan is adequate there.
  k_tr_scan_serial<<<1, 1, 0, st>>>(n_cols, t_indptr);
  TFEM_LAUNCH_CHECK();
  int64_t* cursor = nullptr;
  TFEM_CUDA(malloc_async(&cursor, n_cols * sizeof(int64_t), st));
  TFEM_CUDA(cudaMemsetAsync(cursor, 0, n_cols * sizeof(int64_t), st));
  k_tr_fill<<<grid_for(n_rows, 128), 128, 0, st>>>(n_rows, indptr, cols, vals, t_indptr, cursor, t_cols, t_vals);
  TFEM_LAUNCH_CHECK();
  k_tr_sort<<<grid_for(n_cols, 128), 128, 0, st>>>(n_cols, t_indptr, t_cols, t_vals);
  TFEM_LAUNCH_CHECK();
  TFEM_CUDA(cudaFreeAsync(cursor, st));
  return TFEM_OK;
}

extern "C" int tfem_sell_slice_ptr(int64_t n_rows, const int64_t* indptr, int64_t* slice_ptr, void* stream_) {
  return tfem_sell_slice_ptr_capped(n_rows, indptr, INT64_MAX, slice_ptr, stream_);
}

extern "C" int tfem_sell_slice_ptr_capped(int64_t n_rows, const int64_t* indptr, int64_t long_cap,
                                          int64_t* slice_ptr, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(indptr && slice_ptr && n_rows > 0, "sell_slice_ptr: bad arguments");
  const int64_t ns = (n_rows + 31) / 32;
  TFEM_CUDA(cudaMemsetAsync(slice_ptr, 0, sizeof(int64_t), st));
  k_sell_widths<<<grid_for(ns * 32, 256), 256, 0, st>>>(n_rows, ns, indptr, slice_ptr + 1, long_cap);
  TFEM_LAUNCH_CHECK();
  size_t bytes = 0;
  TFEM_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, slice_ptr + 1, slice_ptr + 1, (int)ns, st));
  void* tmp = nullptr;
  TFEM_CUDA(malloc_async(&tmp, bytes ? bytes : 16, st));
  TFEM_CUDA(cub::DeviceScan::InclusiveSum(tmp, bytes, slice_ptr + 1, slice_ptr + 1, (int)ns, st));
  TFEM_CUDA(cudaFreeAsync(tmp, st));
  return TFEM_OK;
}

extern "C" int tfem_sell_fill_rect(int64_t n_rows, int64_t n_cols, const int64_t* indptr, const int32_t* cols,
                                   const double* vals, const int64_t* slice_ptr, int32_t* sell_cols,
                                   double* sell_vals, void* stream_) {
  return tfem_sell_fill_capped(n_rows, n_cols, indptr, cols, vals, INT64_MAX, slice_ptr, sell_cols, sell_vals, stream_);
}

extern "C" int tfem_sell_fill_capped(int64_t n_rows, int64_t n_cols, const int64_t* indptr, const int32_t* cols,
                                     const double* vals, int64_t long_cap, const int64_t* slice_ptr,
                                     int32_t* sell_cols, double* sell_vals, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(indptr && slice_ptr && n_rows > 0 && n_cols > 0, "sell_fill: bad arguments");
  TFEM_REQUIRE((!sell_cols || cols) && (!sell_vals || vals), "sell_fill: output without input");
  const int64_t ns = (n_rows + 31) / 32;
  k_sell_fill<<<grid_for(ns * 32, 256), 256, 0, st>>>(n_rows, n_cols, ns, indptr, cols, vals, slice_ptr, sell_cols,
                                                      sell_vals, long_cap);
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

extern "C" int tfem_sell_fill(int64_t n_rows, const int64_t* indptr, const int32_t* cols,
                              const double* vals, const int64_t* slice_ptr, int32_t* sell_cols,
                              double* sell_vals, void* stream_) {
  return tfem_sell_fill_rect(n_rows, n_rows, indptr, cols, vals, slice_ptr, sell_cols, sell_vals, stream_);
}

extern "C" int tfem_bsell_slice_ptr(int64_t n_rows, int dpn, const int64_t* slice_ptr,
                                    int64_t* bslice_ptr, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(slice_ptr && bslice_ptr && n_rows > 0 && (dpn == 2 || dpn == 3), "bsell_slice_ptr: bad arguments");
  const int64_t ns = (n_rows + 31) / 32;
  const int nps = dpn == 3 ? 12 : 16;
  TFEM_CUDA(cudaMemsetAsync(bslice_ptr, 0, sizeof(int64_t), st));
  k_bsell_widths<<<grid_for(ns, 256), 256, 0, st>>>(ns, dpn, nps, slice_ptr, bslice_ptr + 1);
  TFEM_LAUNCH_CHECK();
  size_t bytes = 0;
  TFEM_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, bslice_ptr + 1, bslice_ptr + 1, (int)ns, st));
  void* tmp = nullptr;
  TFEM_CUDA(malloc_async(&tmp, bytes ? bytes : 16, st));
  TFEM_CUDA(cub::DeviceScan::InclusiveSum(tmp, bytes, bslice_ptr + 1, bslice_ptr + 1, (int)ns, st));
  TFEM_CUDA(cudaFreeAsync(tmp, st));
  return TFEM_OK;
}

extern "C" int tfem_bsell_fill(int64_t n_rows, int dpn, int64_t n_nod, const int64_t* node_ptr,
                               const int32_t* adj, const int64_t* bslice_ptr, int32_t* bcols,
                               void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(node_ptr && adj && bslice_ptr && bcols && n_rows == n_nod * dpn && (dpn == 2 || dpn == 3),
               "bsell_fill: bad arguments");
  const int64_t ns = (n_rows + 31) / 32;
  k_bsell_fill<<<grid_for(ns * 32, 256), 256, 0, st>>>(ns, n_nod, dpn, dpn == 3 ? 12 : 16, node_ptr, adj,
                                                       bslice_ptr, bcols);
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

extern "C" int tfem_sell_spmv(const tfem_sell_t* a, const double* x, double* y, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  int rc = check_sell(a);
  if (rc != TFEM_OK) return rc;
  TFEM_REQUIRE(x && y, "sell_spmv: null pointer");
  Sell A = make_sell(a);
  return launch_sell<false>(A, x, y, nullptr, nullptr, nullptr, nullptr, st);
}

extern "C" int tfem_sell_spmm(const tfem_sell_t* a, int64_t m, const double* X, int64_t ldx, double* Y, int64_t ldy,
                              void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  int rc = check_sell(a);
  if (rc != TFEM_OK) return rc;
  TFEM_REQUIRE(X && Y && X != Y, "sell_spmm: null or aliased blocks");
  TFEM_REQUIRE(m >= 1 && ldx >= m && ldy >= m, "sell_spmm: bad block shape");
  Sell A = make_sell(a);
  TFEM_REQUIRE(A.n_long == 0, "sell_spmm: matrices with long rows take one product per vector (tfem_sell_spmv)");
  if (A.dpn == 3) return launch_spmm_t<3>(A, m, X, ldx, Y, ldy, st);
  if (A.dpn == 2) return launch_spmm_t<2>(A, m, X, ldx, Y, ldy, st);
  return launch_spmm_t<0>(A, m, X, ldx, Y, ldy, st);
}

extern "C" int64_t tfem_krylov_work_doubles(int64_t n_rows) {
  return 6 * pad32(n_rows) + SC_COUNT + kMaxPartials + 32;
}

// A batch of `check_every` Krylov iterations captured once into a CUDA graph and replayed: below a few hundred
// thousand unknowns the three (CG) / five (MINRES) kernels of an iteration take less time than their launches, and a
// graph launch removes the per-kernel launch gap. Stream capture is not possible on the legacy default stream (which is
// what torch hands over by default), so such solves run on a private non-blocking stream fenced with an event.
struct BatchGraph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  ~BatchGraph() {
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
  }
};

constexpr int64_t kGraphMaxRows = 3000000;  // larger systems are bandwidth-bound: launch gaps do not matter there

static cudaStream_t private_stream() {
  static thread_local cudaStream_t ps = nullptr;
  if (!ps && cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking) != cudaSuccess) ps = nullptr;
  return ps;
}

static int krylov_solve_impl(int method, const Op& A, const double* dinv, const double* b, const double* x0,
                             double rtol, double atol, int64_t maxiter, int check_every, double* x,
                             double* work, double* info, cudaStream_t st_caller) {
  TFEM_REQUIRE(dinv && b && x && work && info, "krylov_solve: null pointer");
  TFEM_REQUIRE(method == TFEM_METHOD_CG || method == TFEM_METHOD_MINRES, "krylov_solve: unknown method");
  const int64_t n = A.n;
  if (maxiter <= 0) maxiter = (method == TFEM_METHOD_CG ? 10 : 5) * n;
  if (check_every <= 0) check_every = 32;

  // work stream: the caller's, unless batches are to be replayed as graphs and the caller's stream cannot capture
  static const bool graphs_off = getenv("TFEM_CG_GRAPH") && atoi(getenv("TFEM_CG_GRAPH")) == 0;
  bool use_graph = !graphs_off && n <= kGraphMaxRows;
  cudaStream_t st = st_caller;
  if (use_graph && (st_caller == nullptr || st_caller == cudaStreamLegacy)) {
    cudaStream_t ps = private_stream();
    cudaEvent_t ev = nullptr;
    if (ps && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) {
      TFEM_CUDA(cudaEventRecord(ev, st_caller));
      TFEM_CUDA(cudaStreamWaitEvent(ps, ev, 0));
      TFEM_CUDA(cudaEventDestroy(ev));
      st = ps;  // the solve ends with a host synchronisation of this stream, so the caller sees finished results
    } else {
      use_graph = false;
    }
  }
  BatchGraph bg;

  Work w = carve(work, n);
  const int vg = vec_grid(n);
  TFEM_CUDA(cudaMemsetAsync(w.sc, 0, (SC_COUNT + kMaxPartials + 32) * sizeof(double), st));
  double launches = 0, spmvs = 0;

  // x = x0 (or 0); initial residual
  k_copy_or_zero<<<vg, kVecThreads, 0, st>>>(n, x0, x);
  TFEM_LAUNCH_CHECK();
  const double* q0 = nullptr;
  if (x0) {
    int rc = apply_op<false>(A, x, w.q, nullptr, nullptr, nullptr, nullptr, st);
    if (rc != TFEM_OK) return rc;
    q0 = w.q;
    spmvs += 1;
    launches += 1;
  }
  const int per_it = method == TFEM_METHOD_CG ? 3 : 5;
  auto iteration = [&]() -> int {
    if (method == TFEM_METHOD_CG) {
      int rc = apply_op<true>(A, w.p, w.q, w.sc, w.partials, w.ticket, w.sc + SC_PQ, st);
      if (rc != TFEM_OK) return rc;
      k_cg_update<<<vg, kVecThreads, 0, st>>>(n, w.p, w.q, dinv, x, w.r, w.sc, w.partials, w.ticket, nullptr);
      k_cg_direction<<<vg, kVecThreads, 0, st>>>(n, w.r, dinv, w.p, w.sc);
    } else {
      // MINRES vectors: r1=w.r, r2=w.p, y=w.q, v, w1, w2
      k_mr_v<<<vg, kVecThreads, 0, st>>>(n, w.y, w.v, w.sc);
      int rc = apply_op<true>(A, w.v, w.y, w.sc, w.partials, w.ticket, w.sc + SC_PQ, st);
      if (rc != TFEM_OK) return rc;
      k_mr_alfa<<<vg, kVecThreads, 0, st>>>(n, w.v, w.y, w.r1, w.sc, w.partials, w.ticket);
      k_mr_lanczos<<<vg, kVecThreads, 0, st>>>(n, dinv, w.y, w.r1, w.r2, w.sc, rtol, w.partials, w.ticket);
      k_mr_xupdate<<<vg, kVecThreads, 0, st>>>(n, w.v, w.w1, w.w2, x, w.sc, rtol, (double)maxiter, w.partials, w.ticket);
    }
    return TFEM_OK;
  };
  double sc_host[SC_COUNT];
  int64_t issued = 0;
  if (method == TFEM_METHOD_CG)
    k_cg_init<<<vg, kVecThreads, 0, st>>>(n, b, q0, dinv, w.r, w.p, w.sc, rtol, atol, w.partials, w.ticket, nullptr);
  else  // the first SpMV consumes q0's storage (y), which k_mr_init has already folded into r1
    k_mr_init<<<vg, kVecThreads, 0, st>>>(n, b, q0, dinv, w.r1, w.r2, w.y, w.w1, w.w2, w.sc, w.partials, w.ticket);
  TFEM_LAUNCH_CHECK();
  launches += 2;
  // small systems on the assembled matrix: the whole loop in one cooperative kernel (see k_cg_coop / k_mr_coop)
  static const bool coop_off = getenv("TFEM_CG_COOP") && atoi(getenv("TFEM_CG_COOP")) == 0;
  bool coop_done = false;
  if (!A.ebe && !coop_off && n <= kCoopMaxRows && A.sell.n_long == 0) {  // (the long-row side path is a second kernel)
    int rc = TFEM_ERR_INVALID;
    if (method == TFEM_METHOD_CG) {
      if (A.sell.dpn == 3) rc = cg_coop_batches<3>(A.sell, dinv, x, w, maxiter, sc_host, &launches, &spmvs, st);
      else if (A.sell.dpn == 2) rc = cg_coop_batches<2>(A.sell, dinv, x, w, maxiter, sc_host, &launches, &spmvs, st);
      else rc = cg_coop_batches<0>(A.sell, dinv, x, w, maxiter, sc_host, &launches, &spmvs, st);
    } else {
      if (A.sell.dpn == 3) rc = mr_coop_batches<3>(A.sell, dinv, x, w, rtol, maxiter, sc_host, &launches, &spmvs, st);
      else if (A.sell.dpn == 2) rc = mr_coop_batches<2>(A.sell, dinv, x, w, rtol, maxiter, sc_host, &launches, &spmvs, st);
      else rc = mr_coop_batches<0>(A.sell, dinv, x, w, rtol, maxiter, sc_host, &launches, &spmvs, st);
    }
    if (rc == TFEM_OK) coop_done = true;
    else if (rc != TFEM_ERR_INVALID) return rc;
  }
  while (!coop_done) {
    TFEM_CUDA(cudaMemcpyAsync(sc_host, w.sc, sizeof(sc_host), cudaMemcpyDeviceToHost, st));
    TFEM_CUDA(cudaStreamSynchronize(st));
    if (sc_host[SC_DONE] != 0.0 || issued >= maxiter) break;
    const int64_t batch = maxiter - issued < check_every ? maxiter - issued : check_every;
    // the first batch is launched kernel by kernel (it also runs the one-time occupancy queries); full batches after
    // it replay the captured graph
    if (use_graph && issued > 0 && batch == check_every) {
      if (!bg.exec) {
        TFEM_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        int rc = TFEM_OK;
        for (int64_t it = 0; it < batch && rc == TFEM_OK; ++it) rc = iteration();
        const cudaError_t ce = cudaStreamEndCapture(st, &bg.graph);
        if (rc != TFEM_OK) return rc;
        TFEM_CUDA(ce);
        TFEM_CUDA(cudaGraphInstantiate(&bg.exec, bg.graph, 0));
      }
      TFEM_CUDA(cudaGraphLaunch(bg.exec, st));
    } else {
      for (int64_t it = 0; it < batch; ++it) {
        int rc = iteration();
        if (rc != TFEM_OK) return rc;
      }
      TFEM_LAUNCH_CHECK();
    }
    issued += batch;
    launches += (double)per_it * batch;
    spmvs += batch;
  }
  info[0] = sc_host[SC_ITERS];
  info[1] = sqrt(sc_host[SC_RR]);
  info[2] = sc_host[SC_BNRM];
  info[3] = sc_host[SC_DONE] == 1.0 ? 1.0 : 0.0;
  info[4] = spmvs;
  info[5] = launches;
  info[6] = sc_host[SC_DONE];
  info[7] = method == TFEM_METHOD_MINRES ? sc_host[SC_M_ISTOP] : 0.0;
  if (sc_host[SC_DONE] == 2.0) {
    set_last_error("breakdown", "non-finite residual or non-positive curvature (matrix not SPD?)");
    return TFEM_ERR_BREAKDOWN;
  }
  if (sc_host[SC_DONE] != 1.0) {
    set_last_error("not converged", "iteration limit reached");
    return TFEM_ERR_NOT_CONVERGED;
  }
  return TFEM_OK;
}

extern "C" int tfem_krylov_solve(int method, const tfem_sell_t* a, const double* dinv, const double* b,
                                 const double* x0, double rtol, double atol, int64_t maxiter,
                                 int check_every, double* x, double* work, double* info, void* stream_) {
  int rc0 = check_sell(a);
  if (rc0 != TFEM_OK) return rc0;
  Op op;
  op.sell = make_sell(a);
  op.n = a->n_rows;
  return krylov_solve_impl(method, op, dinv, b, x0, rtol, atol, maxiter, check_every, x, work, info,
                           (cudaStream_t)stream_);
}

extern "C" int tfem_krylov_solve_ebe(int method, const tfem_ebe_t* a, const double* dinv, const double* b,
                                     const double* x0, double rtol, double atol, int64_t maxiter,
                                     int check_every, double* x, double* work, double* info, void* stream_) {
  int rc0 = check_ebe(a);
  if (rc0 != TFEM_OK) return rc0;
  Op op;
  op.ebe = true;
  op.el = make_ebe(a);
  op.n = a->n_nod * a->dpn;
  return krylov_solve_impl(method, op, dinv, b, x0, rtol, atol, maxiter, check_every, x, work, info,
                           (cudaStream_t)stream_);
}

extern "C" int tfem_ebe_spmv(const tfem_ebe_t* a, const double* x, double* y, void* stream_) {
  int rc0 = check_ebe(a);
  if (rc0 != TFEM_OK) return rc0;
  TFEM_REQUIRE(x && y, "ebe_spmv: null pointer");
  return launch_ebe<false>(make_ebe(a), x, y, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream_);
}

extern "C" int tfem_ebe_diag(const tfem_ebe_t* a, double* diag, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  int rc0 = check_ebe(a);
  if (rc0 != TFEM_OK) return rc0;
  TFEM_REQUIRE(diag, "ebe_diag: null pointer");
  const Ebe A = make_ebe(a);
  const unsigned grid = grid_for(A.n_nod, 256);
  if (A.dpn == 3) k_ebe_diag<3><<<grid, 256, 0, st>>>(A, diag);
  else if (A.dpn == 2) k_ebe_diag<2><<<grid, 256, 0, st>>>(A, diag);
  else k_ebe_diag<1><<<grid, 256, 0, st>>>(A, diag);
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

// ---------------------------------------------------------------------------------------------------
// Multi-GPU CG, one call per stage (the host interleaves NCCL collectives between stages; see
// torch-fem_b200/distributed.py). Vectors have n_local entries (owned rows [row_lo, row_lo+n_owned) plus
// halo rows); vector kernels touch the owned range only, the SpMV runs over all local rows and its fused
// dot covers the owned rows. `red_dev` (double[4]) carries local sums out and reduced sums back in.
extern "C" int tfem_cg_stage(int stage, const tfem_sell_t* a, int64_t row_lo, int64_t n_owned,
                             const double* dinv, const double* b, double* x, double* work, double* red,
                             double rtol, double atol, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(a && a->n_rows > 0, "cg_stage: null matrix");
  const int64_t n_local = a->n_rows;
  TFEM_REQUIRE(work && red && n_owned > 0 && row_lo >= 0 && row_lo + n_owned <= n_local,
               "cg_stage: bad arguments");
  Work w = carve(work, n_local);
  const int vg = vec_grid(n_owned);
  const int64_t o = row_lo;
  switch (stage) {
    case 0:  // reset state, x = 0, r = b, p = dinv r ; red <- (rr, rho, bb)
      TFEM_REQUIRE(dinv && b && x, "cg_stage: null pointer");
      TFEM_CUDA(cudaMemsetAsync(w.sc, 0, (SC_COUNT + kMaxPartials + 32) * sizeof(double), st));
      TFEM_CUDA(cudaMemsetAsync(w.p, 0, n_local * sizeof(double), st));
      k_copy_or_zero<<<vg, kVecThreads, 0, st>>>(n_owned, nullptr, x + o);
      k_cg_init<<<vg, kVecThreads, 0, st>>>(n_owned, b + o, nullptr, dinv + o, w.r + o, w.p + o, w.sc, rtol,
                                           atol, w.partials, w.ticket, red);
      break;
    case 1:
    case 3:
    case 5:  // scalar recurrences on the all-reduced sums
      k_cg_scalars<<<1, 1, 0, st>>>(stage / 2, w.sc, red, rtol, atol);
      break;
    case 2: {  // q = A p (all local rows), red <- p.q over owned rows
      int rcs = check_sell(a);
      if (rcs != TFEM_OK) return rcs;
      Sell A = make_sell(a);
      A.dot_lo = row_lo;
      A.dot_hi = row_lo + n_owned;
      int rc = launch_sell<true>(A, w.p, w.q, w.sc, w.partials, w.ticket, red, st);
      if (rc != TFEM_OK) return rc;
      break;
    }
    case 4:  // x += alpha p ; r -= alpha q ; red <- (rr, rho)
      TFEM_REQUIRE(dinv && x, "cg_stage: null pointer");
      k_cg_update<<<vg, kVecThreads, 0, st>>>(n_owned, w.p + o, w.q + o, dinv + o, x + o, w.r + o, w.sc,
                                             w.partials, w.ticket, red);
      break;
    case 6:  // p = dinv r + beta p
      TFEM_REQUIRE(dinv, "cg_stage: null pointer");
      k_cg_direction<<<vg, kVecThreads, 0, st>>>(n_owned, w.r + o, dinv + o, w.p + o, w.sc);
      break;
    default:
      set_last_error("invalid argument", "cg_stage: unknown stage");
      return TFEM_ERR_INVALID;
  }
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

// Offsets (in doubles) inside the Krylov work buffer: which = 0 r, 1 p, 2 q, 3 scalars.
extern "C" int64_t tfem_krylov_work_offset(int64_t n_rows, int which) {
  const int64_t np = pad32(n_rows);
  return which <= 2 ? which * np : 6 * np;
}

// Copies {iterations, ||r||, ||b||, done flag} of the solve living in `work` to the host (synchronises).
extern "C" int tfem_krylov_state(int64_t n_rows, const double* work, double* info_host, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(work && info_host, "krylov_state: null pointer");
  double sc_host[SC_COUNT];
  TFEM_CUDA(cudaMemcpyAsync(sc_host, work + 6 * pad32(n_rows), sizeof(sc_host), cudaMemcpyDeviceToHost, st));
  TFEM_CUDA(cudaStreamSynchronize(st));
  info_host[0] = sc_host[SC_ITERS];
  info_host[1] = sqrt(sc_host[SC_RR]);
  info_host[2] = sc_host[SC_BNRM];
  info_host[3] = sc_host[SC_DONE];
  return TFEM_OK;
}
