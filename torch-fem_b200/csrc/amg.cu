// K11-K16 — aggregation algebraic multigrid: setup kernels, V cycle and AMG-preconditioned CG.
//
// Replaces the two third-party AMG back ends of the reference: pyamg `smoothed_aggregation_solver(A, B,
// smooth="jacobi")` + scipy cg on the CPU (src/torchfem/sparse.py:493-512) and the AmgX aggregation-AMG solver on the
// GPU (src/torchfem/amgx.py:71-98, sparse.py:422-442; V cycle, one pre/post sweep, dense solve on the coarsest
// level). The algorithm is smoothed aggregation (Vanek/Mandel/Brezina) on the NODE graph with d x d blocks:
//   K11 row info        dinv, "isolated" DOFs (Dirichlet rows), zero-diagonal repair on coarse levels
//   K12 aggregation     maximal independent set with fixed pseudo-random keys (Luby rounds), members join the
//                       adjacent root with the largest key                       — integer only, deterministic
//   K13 prolongator     P = (I - w D^-1 A) T, T = masked piecewise-constant translations; pattern + values per row
//   K14 transpose       R = P^T (counting sort on columns, rows sorted by rank; values gathered through `src`)
//   K15 SpGEMM          C = X Y on block-CSR operands: symbolic (shared-memory hash set per row, rank sort) and
//                       numeric (shared-memory accumulators, contributions added in the order of X's row — fixed
//                       summation order, no FP atomics); used for A P and R (A P)
//   K16 V cycle / PCG   SELL-32 SpMV (sell.cuh) with fused epilogues: residual, damped-Jacobi update, prolongation
//                       add, and the r.z / p.q dot products of CG
//
// Every operator of the hierarchy (A_l, P_l, R_l) is a block-CSR matrix over nodes (bptr int64, bcol int32 sorted)
// whose values are stored in the SCALAR CSR order of the assembled matrix: block row I with m blocks owns
// d*d*m values at d*d*bptr[I]; entry (a, s, c) = row DOF a, s-th block, column DOF c sits at (a*m + s)*d + c.
// That is exactly the layout tfem_assemble writes, so level 0 is the assembled matrix itself and every level can be
// handed to the SELL-32 converter unchanged.
#include <math.h>

#include <cub/device/device_scan.cuh>

#include <string.h>

#include "sell.cuh"
#include "peer.cuh"

namespace tfem {
namespace {

constexpr int kRowCap = 768;   // blocks per block row the prolongator kernel can hold in shared memory
constexpr int kAggCap = 256;   // distinct aggregates one row of P may touch (+1)
constexpr int kStageMax = 160;  // most blocks of a block row the prolongator kernel stages in shared memory (longer rows: global loads)

__device__ __forceinline__ uint32_t hash32(uint32_t h) {  // MurmurHash3 finaliser, as oracle/amg_oracle.py
  h += 0x9E3779B9u;
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
__device__ __forceinline__ uint64_t mis_key(int64_t i) {
  return ((uint64_t)hash32((uint32_t)i) << 32) | (uint64_t)(i + 1);
}

// ------------------------------------------------------------------------------------------ K11 row info
template <int D>
__global__ void k_row_info(int64_t nb, const int64_t* __restrict__ bptr, const int32_t* __restrict__ bcol,
                           double* vals, int fix_zero_diag, double* __restrict__ dinv,
                           uint8_t* __restrict__ iso) {
  const int lane = threadIdx.x & 31;
  const int64_t I = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (I >= nb) return;
  const int64_t b0 = bptr[I];
  const int m = (int)(bptr[I + 1] - b0);
  for (int a = 0; a < D; ++a) {
    double* row = vals + D * D * b0 + (int64_t)a * m * D;
    double diag = 0.0;
    int dpos = -1;
    bool off = false;
    for (int e = lane; e < m * D; e += 32) {
      const int s = e / D, c = e - s * D;
      const double v = row[e];
      if (bcol[b0 + s] == (int32_t)I && c == a) {
        diag = v;
        dpos = e;
      } else if (v != 0.0) {
        off = true;
      }
    }
    off = __any_sync(0xffffffffu, off);
    const double dsum = warp_sum(diag);  // exactly one lane holds it
    if (fix_zero_diag && dsum == 0.0 && dpos >= 0) row[dpos] = 1.0;
    if (lane == 0) {
      dinv[I * D + a] = 1.0 / ((fix_zero_diag && dsum == 0.0) ? 1.0 : dsum);
      iso[I * D + a] = off ? 0 : 1;
    }
  }
}

// ------------------------------------------------------------------------------------------ power iteration
__global__ void k_pw_init(int64_t n, double* __restrict__ x) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = 1.0 + (double)(hash32((uint32_t)i) & 1023u) / 1024.0;
}

// y <- dinv*y ; sums y.y and x.x
__global__ void __launch_bounds__(kVecThreads)
    k_pw_step(int64_t n, const double* __restrict__ dinv, const double* __restrict__ x, double* __restrict__ y,
              double* sc, double* partials, unsigned int* ticket) {
  __shared__ double s_red[kVecThreads / 32];
  double yy = 0.0, xx = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kVecThreads) {
    const double yi = dinv[i] * y[i], xi = x[i];
    y[i] = yi;
    yy = fma(yi, yi, yy);
    xx = fma(xi, xi, xx);
  }
  double mine[2], tot[2];
  mine[0] = block_sum<kVecThreads>(yy, s_red);
  mine[1] = block_sum<kVecThreads>(xx, s_red);
  if (publish_and_reduce<2>(mine, partials, ticket, tot) && threadIdx.x == 0) {
    sc[0] = tot[0];
    sc[1] = tot[1];
  }
}

__global__ void k_pw_scale(int64_t n, const double* __restrict__ y, double* __restrict__ x, const double* sc) {
  const double s = 1.0 / sqrt(sc[0]);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = s * y[i];
}

// ------------------------------------------------------------------------------------------ K12 aggregation
__global__ void k_mis_select(int64_t nb, const int64_t* __restrict__ bptr, const int32_t* __restrict__ bcol,
                             const int8_t* __restrict__ state, uint8_t* __restrict__ newroot) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nb) return;
  uint8_t nr = 0;
  if (state[i] == 0) {
    const uint64_t me = mis_key(i);
    nr = 1;
    for (int64_t k = bptr[i]; k < bptr[i + 1]; ++k) {
      const int32_t j = bcol[k];
      if (j != i && state[j] == 0 && mis_key(j) > me) {
        nr = 0;
        break;
      }
    }
  }
  newroot[i] = nr;
}

__global__ void k_mis_apply(int64_t nb, const int64_t* __restrict__ bptr, const int32_t* __restrict__ bcol,
                            int8_t* __restrict__ state, const uint8_t* __restrict__ newroot, int* n_undecided) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nb || state[i] != 0) return;
  if (newroot[i]) {
    state[i] = 1;
    return;
  }
  for (int64_t k = bptr[i]; k < bptr[i + 1]; ++k)
    if (newroot[bcol[k]]) {
      state[i] = 2;
      return;
    }
  atomicAdd(n_undecided, 1);
}

// nodes a rank does not own (halo nodes of a partitioned mesh) take no part in the aggregation: state 3
__global__ void k_mis_exclude(int64_t nb, const uint8_t* __restrict__ exclude, int8_t* __restrict__ state) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < nb && exclude[i]) state[i] = 3;
}

__global__ void k_mis_flags(int64_t nb, const int8_t* __restrict__ state, int32_t* __restrict__ flag) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < nb) flag[i] = state[i] == 1 ? 1 : 0;
}

__global__ void k_mis_assign(int64_t nb, const int64_t* __restrict__ bptr, const int32_t* __restrict__ bcol,
                             const int8_t* __restrict__ state, const int32_t* __restrict__ root_index,
                             int32_t* __restrict__ agg) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nb) return;
  uint64_t best = state[i] == 1 ? mis_key(i) : 0ull;
  for (int64_t k = bptr[i]; k < bptr[i + 1]; ++k) {
    const int32_t j = bcol[k];
    if (state[j] == 1) {
      const uint64_t kj = mis_key(j);
      if (kj > best) best = kj;
    }
  }
  const int64_t root = (int64_t)(best & 0xFFFFFFFFull) - 1;
  agg[i] = root >= 0 ? root_index[root] : -1;
}

// distance-2 variant (graphs of low degree — Tetra1: a radius-1 aggregate holds only ~3 nodes there, the coarse
// operators fill in and the complexity explodes; radius-2 aggregates hold ~20): a node becomes a root if its key is
// the largest among the undecided nodes within distance 2, everything within distance 2 of a root is covered.
__global__ void k_mis2_t1(int64_t nb, const int64_t* __restrict__ bptr, const int32_t* __restrict__ bcol,
                          const int8_t* __restrict__ state, unsigned long long* __restrict__ t1) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nb) return;
  unsigned long long best = state[i] == 0 ? mis_key(i) : 0ull;
  for (int64_t k = bptr[i]; k < bptr[i + 1]; ++k) {
    const int32_t j = bcol[k];
    if (state[j] == 0) {
      const unsigned long long kj = mis_key(j);
      if (kj > best) best = kj;
    }
  }
  t1[i] = best;
}
__global__ void k_mis2_select(int64_t nb, const int64_t* __restrict__ bptr, const int32_t* __restrict__ bcol,
                              const int8_t* __restrict__ state, const unsigned long long* __restrict__ t1,
                              uint8_t* __restrict__ newroot) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nb) return;
  uint8_t nr = 0;
  if (state[i] == 0) {
    const unsigned long long me = mis_key(i);
    nr = t1[i] == me ? 1 : 0;
    for (int64_t k = bptr[i]; nr && k < bptr[i + 1]; ++k)
      if (t1[bcol[k]] > me) nr = 0;
  }
  newroot[i] = nr;
}
__global__ void k_mis2_near(int64_t nb, const int64_t* __restrict__ bptr, const int32_t* __restrict__ bcol,
                            const uint8_t* __restrict__ newroot, uint8_t* __restrict__ near1) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nb) return;
  uint8_t c = newroot[i];
  for (int64_t k = bptr[i]; !c && k < bptr[i + 1]; ++k) c = newroot[bcol[k]];
  near1[i] = c;
}
__global__ void k_mis2_apply(int64_t nb, const int64_t* __restrict__ bptr, const int32_t* __restrict__ bcol,
                             int8_t* __restrict__ state, const uint8_t* __restrict__ newroot,
                             const uint8_t* __restrict__ near1, int* n_undecided) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nb || state[i] != 0) return;
  if (newroot[i]) {
    state[i] = 1;
    return;
  }
  uint8_t c = near1[i];
  for (int64_t k = bptr[i]; !c && k < bptr[i + 1]; ++k) c = near1[bcol[k]];
  if (c) state[i] = 2;
  else atomicAdd(n_undecided, 1);
}
// nodes two steps away from every root join the aggregate of the neighbour (one step from a root) with the largest key
__global__ void k_mis2_assign_far(int64_t nb, const int64_t* __restrict__ bptr, const int32_t* __restrict__ bcol,
                                  const int32_t* __restrict__ agg1, int32_t* __restrict__ agg) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nb) return;
  int32_t a = agg1[i];
  if (a < 0) {
    unsigned long long best = 0ull;
    for (int64_t k = bptr[i]; k < bptr[i + 1]; ++k) {
      const int32_t j = bcol[k];
      if (agg1[j] >= 0) {
        const unsigned long long kj = mis_key(j);
        if (kj > best) {
          best = kj;
          a = agg1[j];
        }
      }
    }
  }
  agg[i] = a;
}

// ------------------------------------------------------------------------------------------ K13 prolongator
// One warp per fine node: the distinct aggregates of its neighbours, sorted (rank by counting in shared memory),
// then  P[i,J] = delta(J, agg i) diag(1-iso_i) - w Dinv_i sum_{j in adj(i), agg j = J} A_ij diag(1-iso_j)
// with the sum taken in adjacency order.
template <int D, bool FILL>
__global__ void __launch_bounds__(128)
    k_prolongator(int64_t nb, const int64_t* __restrict__ bptr, const int32_t* __restrict__ bcol,
                  const double* __restrict__ vals, const int32_t* __restrict__ agg,
                  const double* __restrict__ dinv, const uint8_t* __restrict__ iso, double omega,
                  int64_t* __restrict__ pcount, const int64_t* __restrict__ pptr, int32_t* __restrict__ pcol,
                  double* __restrict__ pvals, int* err, int stage_cap, int64_t row0) {
  __shared__ int s_key[4][kRowCap];
  __shared__ int s_ukey[4][FILL ? kAggCap : 1];
  __shared__ short s_rank[4][kRowCap];
  __shared__ unsigned char s_first[4][kRowCap];
  __shared__ int s_start[4][FILL ? kAggCap : 1];   // FILL: m < kAggCap distinct aggregates per row (checked below)
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rows [row0, row0 + nb) of the operator; the outputs (pcount / pptr rows) are numbered from 0
  const int64_t Io = blockIdx.x * 4ll + w;
  if (Io >= nb) return;
  const int64_t I = row0 + Io;
  const int64_t b0 = bptr[I];
  const int L = (int)(bptr[I + 1] - b0);
  if (L > kRowCap) {
    if (lane == 0) *err = 1;
    return;
  }
  int* key = s_key[w];
  for (int e = lane; e < L; e += 32) key[e] = agg[bcol[b0 + e]];
  __syncwarp();
  int firsts = 0;
  for (int e = lane; e < L; e += 32) {
    const int k = key[e];
    bool first = true;
    for (int e2 = 0; e2 < e; ++e2)
      if (key[e2] == k) {
        first = false;
        break;
      }
    s_first[w][e] = first ? 1 : 0;
    firsts += first ? 1 : 0;
  }
  const int m = warp_sum_int(firsts);
  if (!FILL) {
    if (lane == 0) pcount[Io] = m;
    return;
  }
  if (m > kAggCap - 1) {
    if (lane == 0) *err = 2;
    return;
  }
  __syncwarp();
  for (int e = lane; e < L; e += 32) {
    const int k = key[e];
    int r = 0;
    for (int e2 = 0; e2 < L; ++e2) r += (s_first[w][e2] && key[e2] < k) ? 1 : 0;
    s_rank[w][e] = (short)r;
    if (s_first[w][e]) s_ukey[w][r] = k;
  }
  __syncwarp();
  const int64_t p0 = pptr[Io];
  for (int r = lane; r < m; r += 32) pcol[p0 + r] = s_ukey[w][r];
  // entries grouped by aggregate, adjacency order kept inside a group: lane r collects the entries of rank r
  // (s_key is free now: it becomes the permutation, s_first/s_rank stay), group r = perm[start[r] .. start[r+1])
  int* perm = s_key[w];
  int* start = s_start[w];
  for (int r = lane; r < m; r += 32) {
    int cnt = 0;
    for (int e = 0; e < L; ++e) cnt += s_rank[w][e] == r ? 1 : 0;
    start[r + 1] = cnt;
  }
  if (lane == 0) start[0] = 0;
  __syncwarp();
  if (lane == 0)
    for (int r = 0; r < m; ++r) start[r + 1] += start[r];
  __syncwarp();
  for (int r = lane; r < m; r += 32) {
    int at = start[r];
    for (int e = 0; e < L; ++e)
      if (s_rank[w][e] == r) perm[at++] = e;
  }
  __syncwarp();
  const int mine = agg[I];
  const double* arow = vals + D * D * b0;
  double* prow = pvals + D * D * p0;
  // the block row of A is staged in shared memory with coalesced loads (read straight from global memory the
  // per-output walk below issues scattered 8-byte loads: 44 % of the kernel's stall samples), and the isolated-DOF
  // flags of the neighbours are packed into one byte per entry (s_first is free after the ranking)
  extern __shared__ double s_stage[];
  double* sa = s_stage + (size_t)w * stage_cap * D * D;
  const bool staged = L <= stage_cap;
  if (staged)
    for (int o = lane; o < L * D * D; o += 32) sa[o] = arow[o];
  for (int e = lane; e < L; e += 32) {
    const int64_t j = bcol[b0 + e];
    unsigned char flags = 0;
#pragma unroll
    for (int c = 0; c < D; ++c) flags |= (unsigned char)((iso[j * D + c] ? 1 : 0) << c);
    s_first[w][e] = flags;
  }
  __syncwarp();
  for (int o = lane; o < m * D * D; o += 32) {
    const int r = o / (D * D), a = (o / D) % D, c = o % D;
    double sum = 0.0;
    for (int q = start[r]; q < start[r + 1]; ++q) {
      const int e = perm[q];
      if (!((s_first[w][e] >> c) & 1)) sum += staged ? sa[(a * L + e) * D + c] : arow[((int64_t)a * L + e) * D + c];
    }
    double v = -omega * dinv[I * D + a] * sum;
    if (s_ukey[w][r] == mine && a == c && !iso[I * D + a]) v += 1.0;
    prow[((int64_t)a * m + r) * D + c] = v;
  }
}

// ------------------------------------------------------------------------------------------ K14 transpose
__global__ void k_bt_count(int64_t nblk, const int32_t* __restrict__ col, int64_t* __restrict__ tptr) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k < nblk) atomicAdd((unsigned long long*)&tptr[col[k] + 1], 1ull);
}
__global__ void k_bt_fill(int64_t nrows, const int64_t* __restrict__ ptr, const int32_t* __restrict__ col,
                          const int64_t* __restrict__ tptr, int32_t* __restrict__ cursor,
                          int32_t* __restrict__ tcol_tmp, int32_t* __restrict__ tsrc_tmp) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  for (int64_t k = ptr[i]; k < ptr[i + 1]; ++k) {
    const int32_t c = col[k];
    const int64_t dst = tptr[c] + atomicAdd(&cursor[c], 1);
    tcol_tmp[dst] = (int32_t)i;
    tsrc_tmp[dst] = (int32_t)k;
  }
}
// one warp per transposed row: rank sort by (unique) column
__global__ void k_bt_sort(int64_t ntrows, const int64_t* __restrict__ tptr, const int32_t* __restrict__ tcol_tmp,
                          const int32_t* __restrict__ tsrc_tmp, int32_t* __restrict__ tcol,
                          int32_t* __restrict__ tsrc) {
  const int lane = threadIdx.x & 31;
  const int64_t J = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (J >= ntrows) return;
  const int64_t b = tptr[J];
  const int L = (int)(tptr[J + 1] - b);
  for (int e = lane; e < L; e += 32) {
    const int32_t k = tcol_tmp[b + e];
    int r = 0;
    for (int e2 = 0; e2 < L; ++e2) r += tcol_tmp[b + e2] < k ? 1 : 0;
    tcol[b + r] = k;
    tsrc[b + r] = tsrc_tmp[b + e];
  }
}
template <int D>
__global__ void k_bt_vals(int64_t ntrows, const int64_t* __restrict__ ptr, const double* __restrict__ vals,
                          const int64_t* __restrict__ tptr, const int32_t* __restrict__ tcol,
                          const int32_t* __restrict__ tsrc, double* __restrict__ tvals) {
  const int lane = threadIdx.x & 31;
  const int64_t J = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (J >= ntrows) return;
  const int64_t b = tptr[J];
  const int L = (int)(tptr[J + 1] - b);
  for (int o = lane; o < L * D * D; o += 32) {
    const int t = o / (D * D), c = (o / D) % D, a = o % D;  // transposed entry (c, a) <- source entry (a, c)
    const int64_t i = tcol[b + t];
    const int64_t p0 = ptr[i];
    const int mi = (int)(ptr[i + 1] - p0);
    const int s = (int)(tsrc[b + t] - p0);
    tvals[D * D * b + ((int64_t)c * L + t) * D + a] = vals[D * D * p0 + ((int64_t)a * mi + s) * D + c];
  }
}

// ------------------------------------------------------------------------------------------ K15 SpGEMM
// symbolic: one warp per row of C = X Y; the distinct columns are collected in a shared-memory hash set (integer CAS;
// the resulting SET does not depend on the insertion order), compacted and rank-sorted.
template <int HC, bool FILL>
__global__ void k_spgemm_sym(int64_t nx, const int64_t* __restrict__ xptr, const int32_t* __restrict__ xcol,
                             const int64_t* __restrict__ yptr, const int32_t* __restrict__ ycol,
                             int64_t* __restrict__ ccount, const int64_t* __restrict__ cptr,
                             int32_t* __restrict__ ccol, int* err) {
  extern __shared__ int s_dyn[];
  const int W = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* table = s_dyn + (size_t)w * HC;
  int* list = s_dyn + (size_t)W * HC + (size_t)w * HC;
  int* counter = s_dyn + (size_t)2 * W * HC + w;
  __shared__ int s_prefix[8][32];
  __shared__ long long s_y0s[8][32];
  int* prefix = s_prefix[w];
  long long* y0s = s_y0s[w];
  for (int64_t i = blockIdx.x * (int64_t)W + w; i < nx; i += (int64_t)gridDim.x * W) {
    for (int e = lane; e < HC; e += 32) table[e] = -1;
    if (lane == 0) *counter = 0;
    __syncwarp();
    bool overflow = false;
    const int64_t x0 = xptr[i];
    const int LX = (int)(xptr[i + 1] - x0);
    for (int kc = 0; kc < LX; kc += 32) {
      // 32 entries of X's row at a time: their Y rows are flattened over the lanes (insertion order is irrelevant
      // for a set), prefix[k] = entries of the first k+1 Y rows
      int64_t my_y0 = 0;
      int my_ly = 0;
      if (kc + lane < LX) {
        const int64_t j = xcol[x0 + kc + lane];
        my_y0 = yptr[j];
        my_ly = (int)(yptr[j + 1] - my_y0);
      }
      const int incl = warp_scan_incl(my_ly, lane);
      prefix[lane] = incl;
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      y0s[lane] = my_y0;
      __syncwarp();
      for (int q = lane; q < total; q += 32) {
        int lo = 0, hi = 31;  // first k with prefix[k] > q
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (prefix[mid] > q) hi = mid; else lo = mid + 1;
        }
        const int e = q - (lo ? prefix[lo - 1] : 0);
        const int key = ycol[y0s[lo] + e];
        unsigned h = ((unsigned)key * 2654435761u) & (HC - 1);
        int probes = 0;
        while (true) {
          const int old = atomicCAS(&table[h], -1, key);
          if (old == -1) {
            atomicAdd(counter, 1);
            break;
          }
          if (old == key) break;
          h = (h + 1) & (HC - 1);
          if (++probes >= HC) {
            overflow = true;
            break;
          }
        }
      }
      __syncwarp();
    }
    __syncwarp();
    const int m = *counter;
    if (__any_sync(0xffffffffu, overflow) || m > (HC / 4) * 3) {
      if (lane == 0) *err = 1;
      if (!FILL && lane == 0) ccount[i] = 0;
      __syncwarp();
      continue;
    }
    if (!FILL) {
      if (lane == 0) ccount[i] = m;
      __syncwarp();
      continue;
    }
    int off = 0;
    for (int base = 0; base < HC; base += 32) {
      const int v = table[base + lane];
      const unsigned mask = __ballot_sync(0xffffffffu, v != -1);
      if (v != -1) list[off + __popc(mask & ((1u << lane) - 1u))] = v;
      off += __popc(mask);
    }
    __syncwarp();
    const int64_t c0 = cptr[i];
    for (int e = lane; e < m; e += 32) {
      const int k = list[e];
      int r = 0;
      for (int e2 = 0; e2 < m; ++e2) r += list[e2] < k ? 1 : 0;
      ccol[c0 + r] = k;
    }
    __syncwarp();
  }
}

// numeric: one group of G threads (a warp, or the whole CTA for long rows) per row of C; accumulators for the row
// live in shared memory. X's row is walked sequentially, the threads spread over (entry of Y's row, row DOF): every
// accumulator receives its contributions in the order of X's row, one per step — a fixed summation order.
template <int G>
__device__ __forceinline__ void group_sync() {
  if constexpr (G == 32) __syncwarp();
  else __syncthreads();
}

// position of block column `key` in the sorted column list cc[0..m) of the row: binary search (CTA variant, long
// rows) or — warp variant — a lookup in a small open-addressing table built once per row (the search was 36 % of
// the stall samples and 24 % of the instructions of A*P at config B)
__device__ __forceinline__ int find_sorted(const int* cc, int m, int key) {
  int lo = 0, hi = m - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cc[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <int D, int G>
__global__ void __launch_bounds__(256, 4)
    k_spgemm_num(int64_t nx, const int64_t* __restrict__ xptr, const int32_t* __restrict__ xcol,
                 const double* __restrict__ xvals, const int64_t* __restrict__ yptr,
                 const int32_t* __restrict__ ycol, const double* __restrict__ yvals,
                 const int64_t* __restrict__ cptr, const int32_t* __restrict__ ccol,
                 double* __restrict__ cvals, int MC, int HT) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ long long s_y0[256];  // Y-row offsets / lengths of a chunk of X's row (one entry per thread), so that
  __shared__ int s_ly[256];        // the walk over X's row costs one memory round trip per step instead of three
  constexpr int DD = D * D;
  constexpr bool HASH = G == 32;
  const int W = blockDim.x / G, w = threadIdx.x / G, lane = threadIdx.x % G;
  double* acc = reinterpret_cast<double*>(s_raw) + (size_t)w * MC * DD;
  int* cc = reinterpret_cast<int*>(reinterpret_cast<double*>(s_raw) + (size_t)W * MC * DD) + (size_t)w * MC;
  // warp variant: hash table of HT (power of two >= 2 MC) entries per warp: key in the high, position in the low half
  long long* table = reinterpret_cast<long long*>(reinterpret_cast<int*>(reinterpret_cast<double*>(s_raw) +
                                                  (size_t)W * MC * DD) + (size_t)W * MC + ((W * MC) & 1)) + (size_t)w * HT;
  // G == 256: one row per CTA, the loop bound is uniform over the CTA (barriers inside)
  for (int64_t i = blockIdx.x * (int64_t)W + w; i < nx; i += (int64_t)gridDim.x * W) {
    const int64_t c0 = cptr[i];
    const int m = (int)(cptr[i + 1] - c0);
    for (int e = lane; e < m * DD; e += G) acc[e] = 0.0;
    if constexpr (HASH) {
      for (int e = lane; e < HT; e += G) table[e] = -1;
      __syncwarp();
      for (int e = lane; e < m; e += G) {
        const int key = ccol[c0 + e];
        unsigned h = ((unsigned)key * 2654435761u) & (unsigned)(HT - 1);
        const long long entry = ((long long)key << 32) | (unsigned)e;
        while (atomicCAS(reinterpret_cast<unsigned long long*>(&table[h]), ~0ull, (unsigned long long)entry) != ~0ull)
          h = (h + 1) & (unsigned)(HT - 1);
      }
    } else {
      for (int e = lane; e < m; e += G) cc[e] = ccol[c0 + e];
    }
    const int64_t x0 = xptr[i];
    const int LX = (int)(xptr[i + 1] - x0);
    for (int kc = 0; kc < LX; kc += G) {
      if (kc + lane < LX) {
        const int64_t j = xcol[x0 + kc + lane];
        const int64_t y0 = yptr[j];
        s_y0[w * G + lane] = y0;
        s_ly[w * G + lane] = (int)(yptr[j + 1] - y0);
      }
      group_sync<G>();
      const int kend = LX - kc < G ? LX - kc : G;
      for (int kk = 0; kk < kend; ++kk) {
        const int kx = kc + kk;
        const int64_t y0 = s_y0[w * G + kk];
        const int LY = s_ly[w * G + kk];
        const double* xrow = xvals + DD * x0 + (int64_t)kx * D;   // + a*LX*D + b
        // task = (entry t of Y's row, row DOFs): a whole block per thread while the row fits the group in one pass
        // that way, one row DOF per thread otherwise
        const bool whole = LY * D > G && LY <= G;
        const int ntask = whole ? LY : LY * D;
        for (int o = lane; o < ntask; o += G) {
          const int t = whole ? o : o / D;
          const int a0 = whole ? 0 : o - t * D, a1 = whole ? D : a0 + 1;
          const int key = ycol[y0 + t];
          const double* yb0 = yvals + DD * y0 + (int64_t)t * D;     // + b*LY*D + c
          double yb[D][D];
#pragma unroll
          for (int b = 0; b < D; ++b)
#pragma unroll
            for (int c = 0; c < D; ++c) yb[b][c] = yb0[(int64_t)b * LY * D + c];
          int pos;
          if constexpr (HASH) {
            unsigned h = ((unsigned)key * 2654435761u) & (unsigned)(HT - 1);
            long long entry = table[h];
            while ((int)(entry >> 32) != key) {
              h = (h + 1) & (unsigned)(HT - 1);
              entry = table[h];
            }
            pos = (int)(entry & 0xffffffffll);
          } else {
            pos = find_sorted(cc, m, key);
          }
          double* dst = acc + pos * DD;
          for (int a = a0; a < a1; ++a) {
            double xa[D];
#pragma unroll
            for (int b = 0; b < D; ++b) xa[b] = xrow[(int64_t)a * LX * D + b];
#pragma unroll
            for (int c = 0; c < D; ++c) {
              double sum = 0.0;
#pragma unroll
              for (int b = 0; b < D; ++b) sum = fma(xa[b], yb[b][c], sum);
              dst[a * D + c] += sum;
            }
          }
        }
        group_sync<G>();
      }
    }
    for (int o = lane; o < m * DD; o += G) {
      const int pos = o / DD, a = (o / D) % D, c = o % D;
      cvals[DD * c0 + ((int64_t)a * m + pos) * D + c] = acc[pos * DD + a * D + c];
    }
    group_sync<G>();
  }
}

// ------------------------------------------------------------------------------------------ K16 V cycle
enum { M_AX = 0, M_RES = 1, M_JAC = 2, M_ADD = 3 };
#ifndef TFEM_AMG_EPILOGUE_PREFETCH
#define TFEM_AMG_EPILOGUE_PREFETCH 1   // 5.48 -> 5.40 ms per AMG-PCG iteration at config B; 0: operands loaded before the row
#endif

// y = A x (M_AX) | b - A x (M_RES) | x + w dinv (b - A x) (M_JAC, y != x) | y + A x (M_ADD), one warp per SELL
// slice, persistent grid. DOT: x.(A x) for M_AX, b.y for M_JAC (fixed-order reduction, last CTA writes *out_scalar).
template <int DPN, int MODE, bool DOT, int MINB = 8>
__global__ void __launch_bounds__(kSellWarps * 32, MINB)
    k_amg_spmv(Sell A, const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ b,
               const double* __restrict__ dinv, double omega, double* partials, unsigned int* ticket,
               double* out_scalar) {
  __shared__ double s_red[kSellWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double dot = 0.0;
  const int64_t t_hi = A.slice_hi < 0 ? A.n_slices : A.slice_hi;
  for (int64_t t = A.slice_lo + (int64_t)blockIdx.x * kSellWarps + warp; t < t_hi; t += (int64_t)gridDim.x * kSellWarps) {
    const int64_t row = t * 32 + lane;
    const bool live = row < A.n && row >= A.dot_lo && row < A.dot_hi;
    double xr = 0.0, br = 0.0, dr = 0.0;
#if TFEM_AMG_EPILOGUE_PREFETCH
    // the epilogue operands are only PREFETCHED before the row is streamed and loaded after it: held in registers
    // through the streaming loop they push the 32-register kernels into spills
    if (live) {
      if (MODE == M_JAC || (MODE == M_AX && DOT)) asm volatile("prefetch.global.L1 [%0];" ::"l"(x + row));
      if (MODE == M_RES || MODE == M_JAC) asm volatile("prefetch.global.L1 [%0];" ::"l"(b + row));
      if (MODE == M_JAC) asm volatile("prefetch.global.L1 [%0];" ::"l"(dinv + row));
      if (MODE == M_ADD) asm volatile("prefetch.global.L1 [%0];" ::"l"(y + row));
    }
    const double acc = slice_row<DPN>(A, t, x, lane);
    if (live) {
      if (MODE == M_JAC || (MODE == M_AX && DOT)) xr = __ldg(x + row);
      if (MODE == M_RES || MODE == M_JAC) br = __ldg(b + row);
      if (MODE == M_JAC) dr = __ldg(dinv + row);
      if (MODE == M_ADD) br = y[row];
    }
#else
    if (live) {
      if (MODE == M_JAC || (MODE == M_AX && DOT)) xr = __ldg(x + row);
      if (MODE == M_RES || MODE == M_JAC) br = __ldg(b + row);
      if (MODE == M_JAC) dr = __ldg(dinv + row);
      if (MODE == M_ADD) br = y[row];
    }
    const double acc = slice_row<DPN>(A, t, x, lane);
#endif
    if (live) {
      double v;
      if (MODE == M_AX) v = acc;
      else if (MODE == M_RES) v = br - acc;
      else if (MODE == M_JAC) v = fma(omega * dr, br - acc, xr);
      else v = br + acc;
      y[row] = v;
      if (DOT) dot = fma(MODE == M_AX ? xr : br, v, dot);
    }
  }
  if (DOT) {
    const double s = block_sum<kSellWarps * 32>(dot, s_red);
    double mine[1] = {s}, tot[1];
    if (publish_and_reduce<1>(mine, partials, ticket, tot) && threadIdx.x == 0) *out_scalar = tot[0];
  }
}

template <int DPN, int MODE, bool DOT>
int launch_amg_spmv_t(const Sell& A, const double* x, double* y, const double* b, const double* dinv, double omega,
                      double* partials, unsigned int* ticket, double* out_scalar, cudaStream_t st) {
  // 8 CTAs per SM (32 registers; the epilogue-heavy variants spill ~80 bytes outside the streaming loop) beat 6 CTAs
  // with 40 registers: 5.92 -> 5.55 ms per AMG-PCG iteration at config B. TFEM_AMG_OCC6=1 selects the latter.
  static const bool occ6 = getenv("TFEM_AMG_OCC6") && atoi(getenv("TFEM_AMG_OCC6")) != 0;
  const int64_t n_sl = (A.slice_hi < 0 ? A.n_slices : A.slice_hi) - A.slice_lo;
  const int64_t want = n_sl > 0 ? (n_sl + kSellWarps - 1) / kSellWarps : 1;
  if (occ6) {
    const int g = cached_resident_ctas(k_amg_spmv<DPN, MODE, DOT, 6>, kSellWarps * 32);
    k_amg_spmv<DPN, MODE, DOT, 6><<<(int)(want < g ? want : g), kSellWarps * 32, 0, st>>>(A, x, y, b, dinv, omega,
                                                                                         partials, ticket, out_scalar);
  } else {
    const int g = cached_resident_ctas(k_amg_spmv<DPN, MODE, DOT>, kSellWarps * 32);  // per instantiation and device
    k_amg_spmv<DPN, MODE, DOT><<<(int)(want < g ? want : g), kSellWarps * 32, 0, st>>>(A, x, y, b, dinv, omega,
                                                                                      partials, ticket, out_scalar);
  }
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

template <int MODE, bool DOT>
int launch_amg_spmv(const Sell& A, const double* x, double* y, const double* b, const double* dinv, double omega,
                    double* partials, unsigned int* ticket, double* out_scalar, cudaStream_t st) {
  if (A.dpn == 3) return launch_amg_spmv_t<3, MODE, DOT>(A, x, y, b, dinv, omega, partials, ticket, out_scalar, st);
  if (A.dpn == 2) return launch_amg_spmv_t<2, MODE, DOT>(A, x, y, b, dinv, omega, partials, ticket, out_scalar, st);
  return launch_amg_spmv_t<0, MODE, DOT>(A, x, y, b, dinv, omega, partials, ticket, out_scalar, st);
}

// The same operations on a block-CSR operator, TPR threads per block row (d scalar rows at once): for the coarse
// levels, whose rows are few and long (hundreds to thousands of blocks) — one row per lane (SELL) would leave the
// GPU idle there. Lanes stride over the blocks of the row, partial sums are combined in a fixed tree.
struct Bcsr {
  int64_t nbr = 0, nblk = 0;
  const int64_t* bptr = nullptr;
  const int32_t* bcol = nullptr;
  const double* vals = nullptr;
  int d = 0;
  int64_t row_lo = 0, row_hi = -1;  // block rows computed and written (-1: all); see Sell::slice_lo
};

template <int D, int MODE, int TPR>
__global__ void __launch_bounds__(256)
    k_bcsr_spmv(Bcsr A, const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ b,
                const double* __restrict__ dinv, double omega) {
  constexpr int RPC = 256 / TPR;
  __shared__ double s_part[D][8];
  const int g = threadIdx.x / TPR, l = threadIdx.x % TPR;
  // the loop bound is uniform over the CTA (shuffles / barriers inside); groups past the end idle with m = 0
  const int64_t I_hi = A.row_hi < 0 ? A.nbr : A.row_hi;
  for (int64_t base = A.row_lo + blockIdx.x * (int64_t)RPC; base < I_hi; base += (int64_t)gridDim.x * RPC) {
    const int64_t I = base + g;
    const bool active = I < I_hi;
    const int64_t b0 = active ? A.bptr[I] : 0;
    const int m = active ? (int)(A.bptr[I + 1] - b0) : 0;
    const double* v = A.vals + D * D * b0;
    double acc[D];
#pragma unroll
    for (int a = 0; a < D; ++a) acc[a] = 0.0;
    for (int s = l; s < m; s += TPR) {
      const int64_t col = A.bcol[b0 + s];
      double xs[D];
#pragma unroll
      for (int c = 0; c < D; ++c) xs[c] = x[col * D + c];
#pragma unroll
      for (int a = 0; a < D; ++a)
#pragma unroll
        for (int c = 0; c < D; ++c) acc[a] = fma(v[((int64_t)a * m + s) * D + c], xs[c], acc[a]);
    }
    constexpr int SW = TPR < 32 ? TPR : 32;
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int o = SW / 2; o > 0; o >>= 1) acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], o);
    if constexpr (TPR == 256) {
      __syncthreads();  // s_part of the previous row has been consumed
      if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int a = 0; a < D; ++a) s_part[a][threadIdx.x >> 5] = acc[a];
      __syncthreads();
      if (l < D) {
        double t = 0.0;
        for (int q = 0; q < 8; ++q) t += s_part[l][q];
        acc[0] = t;  // thread l < D holds row DOF l in acc[0]
      }
    }
    if (active && l < D) {
      double sum = acc[0];
      if constexpr (TPR != 256) {
#pragma unroll
        for (int a = 1; a < D; ++a)
          if (l == a) sum = acc[a];
      }
      const int64_t row = I * D + l;
      double out;
      if (MODE == M_AX) out = sum;
      else if (MODE == M_RES) out = b[row] - sum;
      else if (MODE == M_JAC) out = fma(omega * dinv[row], b[row] - sum, x[row]);
      else out = y[row] + sum;
      y[row] = out;
    }
  }
}

template <int D, int MODE>
int launch_bcsr_d(const Bcsr& A, const double* x, double* y, const double* b, const double* dinv, double omega,
                  cudaStream_t st) {
  const double avg = (double)A.nblk / (double)(A.nbr > 0 ? A.nbr : 1);
  const int64_t cap = (int64_t)num_sms() * 8;
  int64_t rows = (A.row_hi < 0 ? A.nbr : A.row_hi) - A.row_lo;
  if (rows < 1) rows = 1;
  if (avg > 1024.0 || rows < 2048) {
    k_bcsr_spmv<D, MODE, 256><<<(unsigned)(rows < cap ? rows : cap), 256, 0, st>>>(A, x, y, b, dinv, omega);
  } else if (avg > 16.0) {
    const int64_t want = (rows + 7) / 8;
    k_bcsr_spmv<D, MODE, 32><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(A, x, y, b, dinv, omega);
  } else {
    const int64_t want = (rows + 31) / 32;
    k_bcsr_spmv<D, MODE, 8><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(A, x, y, b, dinv, omega);
  }
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

template <int MODE>
int launch_bcsr(const Bcsr& A, const double* x, double* y, const double* b, const double* dinv, double omega,
                cudaStream_t st) {
  if (A.d == 3) return launch_bcsr_d<3, MODE>(A, x, y, b, dinv, omega, st);
  if (A.d == 2) return launch_bcsr_d<2, MODE>(A, x, y, b, dinv, omega, st);
  return launch_bcsr_d<1, MODE>(A, x, y, b, dinv, omega, st);
}

// an operator of the hierarchy in whichever layout the host chose for it
struct Oper {
  bool bcsr = false;
  Sell sell;
  Bcsr blk;
  int64_t n = 0;  // scalar rows
};

template <int MODE>
int apply_oper(const Oper& A, const double* x, double* y, const double* b, const double* dinv, double omega,
               cudaStream_t st) {
  if (A.bcsr) return launch_bcsr<MODE>(A.blk, x, y, b, dinv, omega, st);
  return launch_amg_spmv<MODE, false>(A.sell, x, y, b, dinv, omega, nullptr, nullptr, nullptr, st);
}

// x = w dinv b  (one damped-Jacobi sweep from the zero vector)
__global__ void __launch_bounds__(kVecThreads)
    k_jacobi_first(int64_t n, double omega, const double* __restrict__ dinv, const double* __restrict__ b,
                   double* __restrict__ x) {
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kVecThreads)
    x[i] = omega * dinv[i] * b[i];
}

// coarsest level: x = Ainv b, dense row-major, one warp per row
__global__ void k_dense_mv(int n, const double* __restrict__ Minv, const double* __restrict__ b,
                           double* __restrict__ x) {
  const int lane = threadIdx.x & 31;
  const int r = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (r >= n) return;
  double s = 0.0;
  for (int k = lane; k < n; k += 32) s = fma(Minv[(int64_t)r * n + k], b[k], s);
  s = warp_sum(s);
  if (lane == 0) x[r] = s;
}

// ---- AMG-PCG scalars (indices into the device scalar block)
enum { P_RHO = 0, P_RHO_NEW, P_PQ, P_RR, P_TOL, P_BNRM, P_DONE, P_ITERS, P_COUNT = 16 };

__device__ __forceinline__ void pcg_scalars_init(double* sc, double rr, double bb, double rtol, double atol) {
  const double bnrm = sqrt(bb);
  const double tol = fmax(atol, rtol * bnrm);
  sc[P_RR] = rr;
  sc[P_BNRM] = bnrm;
  sc[P_TOL] = tol;
  sc[P_ITERS] = 0.0;
  sc[P_DONE] = (bnrm == 0.0 || sqrt(rr) < tol) ? 1.0 : 0.0;
}

__device__ __forceinline__ void pcg_scalars_update(double* sc, double rr) {
  sc[P_RR] = rr;
  sc[P_ITERS] += 1.0;
  // breakdown = a non-finite residual only, like scipy's cg and the Jacobi kernels: Newton tangents of a state that
  // is not yet in equilibrium can be slightly indefinite (p.q < 0 in some iteration) and CG still gets through
  if (!isfinite(rr)) sc[P_DONE] = 2.0;
  else if (sqrt(rr) < sc[P_TOL]) sc[P_DONE] = 1.0;
}

// r = b - q (q = A x0) or r = b ; rr, bb ; tolerance and the convergence test of the initial iterate
__global__ void __launch_bounds__(kVecThreads)
    k_pcg_init(int64_t n, const double* __restrict__ b, const double* __restrict__ q_or_null, double* __restrict__ r,
               double* sc, double rtol, double atol, double* partials, unsigned int* ticket, double* red = nullptr) {
  __shared__ double s_red[kVecThreads / 32];
  double rr = 0.0, bb = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kVecThreads) {
    const double bi = b[i];
    const double ri = q_or_null ? bi - q_or_null[i] : bi;
    r[i] = ri;
    rr = fma(ri, ri, rr);
    bb = fma(bi, bi, bb);
  }
  double mine[2], tot[2];
  mine[0] = block_sum<kVecThreads>(rr, s_red);
  mine[1] = block_sum<kVecThreads>(bb, s_red);
  if (publish_and_reduce<2>(mine, partials, ticket, tot) && threadIdx.x == 0) {
    if (red) {  // distributed: local sums only, k_allreduce applies pcg_scalars_init to the global ones
      red[0] = tot[0];
      red[1] = tot[1];
    } else {
      pcg_scalars_init(sc, tot[0], tot[1], rtol, atol);
    }
  }
}

// alpha = rho / p.q ; x += alpha p ; r -= alpha q ; rr = r.r ; convergence test (scipy cg: ||r|| < tol)
__global__ void __launch_bounds__(kVecThreads)
    k_pcg_update(int64_t n, const double* __restrict__ p, const double* __restrict__ q, double* __restrict__ x,
                 double* __restrict__ r, double* sc, double* partials, unsigned int* ticket, double* red = nullptr) {
  __shared__ double s_red[kVecThreads / 32];
  const double pq = sc[P_PQ];
  const double alpha = sc[P_RHO] / pq;
  double rr = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kVecThreads) {
    x[i] = fma(alpha, p[i], x[i]);
    const double ri = fma(-alpha, q[i], r[i]);
    r[i] = ri;
    rr = fma(ri, ri, rr);
  }
  double mine[1], tot[1];
  mine[0] = block_sum<kVecThreads>(rr, s_red);
  if (publish_and_reduce<1>(mine, partials, ticket, tot) && threadIdx.x == 0) {
    if (red) red[0] = tot[0];
    else pcg_scalars_update(sc, tot[0]);
  }
}

// p = z + (rho_new / rho) p   (first: p = z)
__global__ void __launch_bounds__(kVecThreads)
    k_pcg_direction(int64_t n, const double* __restrict__ z, double* __restrict__ p, const double* sc, int first) {
  const double beta = first ? 0.0 : sc[P_RHO_NEW] / sc[P_RHO];
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kVecThreads)
    p[i] = first ? z[i] : fma(beta, p[i], z[i]);
}
__global__ void k_pcg_roll(double* sc) { sc[P_RHO] = sc[P_RHO_NEW]; }

// *out = a.b (single-level "hierarchy": the dense solve has no smoother to fuse the dot into)
__global__ void __launch_bounds__(kVecThreads)
    k_dot(int64_t n, const double* __restrict__ a, const double* __restrict__ b, double* partials,
          unsigned int* ticket, double* out) {
  __shared__ double s_red[kVecThreads / 32];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kVecThreads)
    s = fma(a[i], b[i], s);
  double mine[1], tot[1];
  mine[0] = block_sum<kVecThreads>(s, s_red);
  if (publish_and_reduce<1>(mine, partials, ticket, tot) && threadIdx.x == 0) *out = tot[0];
}

struct Level {
  Oper A, P, R;
  int64_t n;
  const double* dinv;
  double omega;
  double *x, *b, *t;
};

int make_oper(const tfem_amg_operator_t* o, Oper* out) {
  if (o->bcsr.nb_rows > 0) {
    TFEM_REQUIRE(o->bcsr.bptr && o->bcsr.bcol && o->bcsr.vals && o->bcsr.d >= 1 && o->bcsr.d <= 3,
                 "amg operator: bad block-CSR descriptor");
    out->bcsr = true;
    out->blk.nbr = o->bcsr.nb_rows;
    out->blk.nblk = o->bcsr.n_blocks;
    out->blk.bptr = o->bcsr.bptr;
    out->blk.bcol = o->bcsr.bcol;
    out->blk.vals = o->bcsr.vals;
    out->blk.d = o->bcsr.d;
    out->n = o->bcsr.nb_rows * o->bcsr.d;
    return TFEM_OK;
  }
  int rc = check_sell(&o->sell);
  if (rc != TFEM_OK) return rc;
  TFEM_REQUIRE(o->sell.n_long == 0, "amg: matrices with long rows (reference-point couplings) need method cg / minres");
  out->bcsr = false;
  out->sell = make_sell(&o->sell);
  out->n = o->sell.n_rows;
  return TFEM_OK;
}

int make_levels(const tfem_amg_level_t* lv, int n_levels, Level* out, bool need_sell0 = true) {
  int rc;
  for (int l = 0; l < n_levels; ++l) {
    if ((rc = make_oper(&lv[l].A, &out[l].A)) != TFEM_OK) return rc;
    out[l].n = out[l].A.n;
    out[l].dinv = lv[l].dinv;
    out[l].omega = lv[l].omega;
    out[l].x = lv[l].x;
    out[l].b = lv[l].b;
    out[l].t = lv[l].t;
  }
  TFEM_REQUIRE(!need_sell0 || !out[0].A.bcsr, "amg: the finest level must be given in SELL-32 form");
  for (int l = 0; l + 1 < n_levels; ++l) {
    if ((rc = make_oper(&lv[l].P, &out[l].P)) != TFEM_OK || (rc = make_oper(&lv[l].R, &out[l].R)) != TFEM_OK) return rc;
    TFEM_REQUIRE(lv[l].dinv && lv[l].x && lv[l].t, "amg level: null work vector");
    TFEM_REQUIRE(out[l].P.n == out[l].n && out[l].R.n == out[l + 1].n, "amg level: P / R shapes");
    TFEM_REQUIRE(lv[l + 1].b && lv[l + 1].x, "amg level: null work vector");
  }
  return TFEM_OK;
}

// z = M r. Level l > 0 reads its right-hand side from L[l].b; the result of a non-coarsest level lands in L[l].t,
// of the coarsest in L[l].x. At level 0 the input is `r`, the output `z`; dot != nullptr fuses r.z into the last
// kernel.
int vcycle(const Level* L, int n_levels, const double* coarse_inv, const double* r, double* z, double* partials,
           unsigned int* ticket, double* dot_out, int64_t* launches, cudaStream_t st) {
  int rc;
  for (int l = 0; l < n_levels - 1; ++l) {  // downward leg
    const Level& v = L[l];
    const double* b = l == 0 ? r : v.b;
    k_jacobi_first<<<vec_grid(v.n), kVecThreads, 0, st>>>(v.n, v.omega, v.dinv, b, v.x);
    if ((rc = apply_oper<M_RES>(v.A, v.x, v.t, b, nullptr, 0.0, st))) return rc;
    if ((rc = apply_oper<M_AX>(v.R, v.t, L[l + 1].b, nullptr, nullptr, 0.0, st))) return rc;
    *launches += 3;
  }
  {
    const Level& c = L[n_levels - 1];
    const double* b = n_levels == 1 ? r : c.b;
    double* x = n_levels == 1 ? z : c.x;
    k_dense_mv<<<grid_for(c.n * 32, 256), 256, 0, st>>>((int)c.n, coarse_inv, b, x);
    *launches += 1;
    if (n_levels == 1 && dot_out) {
      k_dot<<<vec_grid(c.n), kVecThreads, 0, st>>>(c.n, r, z, partials, ticket, dot_out);
      *launches += 1;
    }
  }
  for (int l = n_levels - 2; l >= 0; --l) {  // upward leg
    const Level& v = L[l];
    const double* b = l == 0 ? r : v.b;
    const double* xc = (l + 1 == n_levels - 1) ? L[l + 1].x : L[l + 1].t;
    double* out = l == 0 ? z : v.t;
    if ((rc = apply_oper<M_ADD>(v.P, xc, v.x, nullptr, nullptr, 0.0, st))) return rc;
    if (l == 0 && dot_out)
      rc = launch_amg_spmv<M_JAC, true>(v.A.sell, v.x, out, b, v.dinv, v.omega, partials, ticket, dot_out, st);
    else
      rc = apply_oper<M_JAC>(v.A, v.x, out, b, v.dinv, v.omega, st);
    if (rc) return rc;
    *launches += 2;
  }
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

// ---- V cycle on a block of nb <= 4 vectors stored row-major [n, nb] (the eigensolver's preconditioner step): the two
// products with the finest operator — more than half of a cycle — read the matrix once for the block
// (k_sell_spmm<., 4, MODE>); restriction, the coarse levels and prolongation run per vector on the levels' work vectors.
// Every expression and summation order is the one of vcycle(): each column equals the single-vector cycle bit for bit.
__global__ void __launch_bounds__(kVecThreads)
    k_jacobi_first_block(int64_t n, int nb, double omega, const double* __restrict__ dinv, const double* __restrict__ B,
                         double* __restrict__ X) {
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n * nb; i += (int64_t)gridDim.x * kVecThreads)
    X[i] = omega * dinv[i / nb] * B[i];
}

__global__ void __launch_bounds__(kVecThreads)
    k_block_column(int64_t n, int nb, int j, double* __restrict__ block, double* __restrict__ vec, int to_block) {
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kVecThreads) {
    if (to_block) block[i * nb + j] = vec[i];
    else vec[i] = block[i * nb + j];
  }
}

template <int DPN, int MODE>
void launch_spmm_mode(const Sell& A, const double* X, double* Y, int nb, const double* B, const double* dinv, double omega,
                      cudaStream_t st) {
  const int g = cached_resident_ctas(k_sell_spmm<DPN, kSpmmBlock, MODE>, kSellWarps * 32);
  const int64_t want = A.n_slices > 0 ? (A.n_slices + kSellWarps - 1) / kSellWarps : 1;
  k_sell_spmm<DPN, kSpmmBlock, MODE><<<(int)(want < g ? want : g), kSellWarps * 32, 0, st>>>(A, X, nb, Y, nb, nb, B, dinv,
                                                                                            omega);
}

template <int MODE>
void spmm_mode(const Sell& A, const double* X, double* Y, int nb, const double* B, const double* dinv, double omega,
               cudaStream_t st) {
  if (A.dpn == 3) launch_spmm_mode<3, MODE>(A, X, Y, nb, B, dinv, omega, st);
  else if (A.dpn == 2) launch_spmm_mode<2, MODE>(A, X, Y, nb, B, dinv, omega, st);
  else launch_spmm_mode<0, MODE>(A, X, Y, nb, B, dinv, omega, st);
}

int vcycle_block(const Level* L, int n_levels, const double* coarse_inv, int nb, const double* Rb, double* Zb,
                 double* work, cudaStream_t st) {
  const Level& v = L[0];
  int64_t launches = 0;
  int rc;
  const bool batched = n_levels >= 2 && !v.A.bcsr && v.A.sell.n_long == 0 && v.A.sell.slice_lo == 0 &&
                       v.A.sell.slice_hi < 0;
  if (!batched) {  // a single level (dense) or a block-CSR finest operator: cycle by cycle through the work vectors
    for (int j = 0; j < nb; ++j) {
      double* rj = work;
      double* zj = work + v.n;
      k_block_column<<<vec_grid(v.n), kVecThreads, 0, st>>>(v.n, nb, j, const_cast<double*>(Rb), rj, 0);
      if ((rc = vcycle(L, n_levels, coarse_inv, rj, zj, nullptr, nullptr, nullptr, &launches, st))) return rc;
      k_block_column<<<vec_grid(v.n), kVecThreads, 0, st>>>(v.n, nb, j, Zb, zj, 1);
    }
    TFEM_LAUNCH_CHECK();
    return TFEM_OK;
  }
  double* X0 = work;                    // [n, nb]
  double* T0 = work + v.n * (int64_t)nb;  // [n, nb]
  k_jacobi_first_block<<<vec_grid(v.n * nb), kVecThreads, 0, st>>>(v.n, nb, v.omega, v.dinv, Rb, X0);
  spmm_mode<M_RES>(v.A.sell, X0, T0, nb, Rb, nullptr, 0.0, st);
  for (int j = 0; j < nb; ++j) {
    k_block_column<<<vec_grid(v.n), kVecThreads, 0, st>>>(v.n, nb, j, T0, v.t, 0);
    if ((rc = apply_oper<M_AX>(v.R, v.t, L[1].b, nullptr, nullptr, 0.0, st))) return rc;
    // levels 1.. as a cycle of their own: right-hand side L[1].b, result in L[1].t
    if ((rc = vcycle(L + 1, n_levels - 1, coarse_inv, L[1].b, L[1].t, nullptr, nullptr, nullptr, &launches, st)))
      return rc;
    k_block_column<<<vec_grid(v.n), kVecThreads, 0, st>>>(v.n, nb, j, X0, v.x, 0);
    if ((rc = apply_oper<M_ADD>(v.P, L[1].t, v.x, nullptr, nullptr, 0.0, st))) return rc;
    k_block_column<<<vec_grid(v.n), kVecThreads, 0, st>>>(v.n, nb, j, X0, v.x, 1);
  }
  spmm_mode<M_JAC>(v.A.sell, X0, Zb, nb, Rb, v.dinv, v.omega, st);
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

int scan_in_place(int64_t* ptr, int64_t n, cudaStream_t st) {  // ptr[0] = 0, ptr[1..n] counts -> inclusive sums
  TFEM_CUDA(cudaMemsetAsync(ptr, 0, sizeof(int64_t), st));
  size_t bytes = 0;
  TFEM_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, ptr + 1, ptr + 1, (int)n, st));
  void* tmp = nullptr;
  TFEM_CUDA(malloc_async(&tmp, bytes ? bytes : 16, st));
  TFEM_CUDA(cub::DeviceScan::InclusiveSum(tmp, bytes, ptr + 1, ptr + 1, (int)n, st));
  TFEM_CUDA(cudaFreeAsync(tmp, st));
  return TFEM_OK;
}

int read_flag(int* dev_flag, int* host, cudaStream_t st) {
  TFEM_CUDA(cudaMemcpyAsync(host, dev_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  TFEM_CUDA(cudaStreamSynchronize(st));
  return TFEM_OK;
}

}  // namespace
}  // namespace tfem

using namespace tfem;

extern "C" int tfem_amg_row_info(int d, int64_t nb, const int64_t* bptr, const int32_t* bcol, double* vals,
                                 int fix_zero_diag, double* dinv, uint8_t* iso, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(bptr && bcol && vals && dinv && iso && nb > 0, "amg_row_info: bad arguments");
  TFEM_REQUIRE(d >= 1 && d <= 3, "amg: 1, 2 or 3 DOFs per node");
  const unsigned grid = grid_for(nb * 32, 256);
  if (d == 3) k_row_info<3><<<grid, 256, 0, st>>>(nb, bptr, bcol, vals, fix_zero_diag, dinv, iso);
  else if (d == 2) k_row_info<2><<<grid, 256, 0, st>>>(nb, bptr, bcol, vals, fix_zero_diag, dinv, iso);
  else k_row_info<1><<<grid, 256, 0, st>>>(nb, bptr, bcol, vals, fix_zero_diag, dinv, iso);
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

extern "C" int64_t tfem_amg_work_doubles(int64_t n_rows) { return 4 * pad32(n_rows) + P_COUNT + kMaxPartials + 32; }

extern "C" int tfem_amg_rho(const tfem_amg_operator_t* a, const double* dinv, int iterations, double* work,
                            double* rho_host, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(a && dinv && work && rho_host && iterations > 0, "amg_rho: bad arguments");
  Oper A;
  int rc = make_oper(a, &A);
  if (rc != TFEM_OK) return rc;
  const int64_t n = A.n, np = pad32(n);
  double *x = work, *y = work + np, *sc = work + 4 * np, *partials = sc + P_COUNT;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(partials + kMaxPartials);
  TFEM_CUDA(cudaMemsetAsync(sc, 0, (P_COUNT + kMaxPartials + 32) * sizeof(double), st));
  const int vg = vec_grid(n);
  k_pw_init<<<vg, kVecThreads, 0, st>>>(n, x);
  for (int it = 0; it < iterations; ++it) {
    if ((rc = apply_oper<M_AX>(A, x, y, nullptr, nullptr, 0.0, st))) return rc;
    k_pw_step<<<vg, kVecThreads, 0, st>>>(n, dinv, x, y, sc, partials, ticket);
    if (it + 1 < iterations) k_pw_scale<<<vg, kVecThreads, 0, st>>>(n, y, x, sc);
  }
  TFEM_LAUNCH_CHECK();
  double h[2];
  TFEM_CUDA(cudaMemcpyAsync(h, sc, sizeof(h), cudaMemcpyDeviceToHost, st));
  TFEM_CUDA(cudaStreamSynchronize(st));
  *rho_host = sqrt(h[0]) / sqrt(h[1]);
  return TFEM_OK;
}

extern "C" int tfem_amg_aggregate(int64_t nb, const int64_t* bptr, const int32_t* bcol, int distance,
                                  int8_t* state_work, uint8_t* flag_work, int32_t* index_work, int32_t* agg,
                                  int64_t* n_agg_host, int32_t* rounds_host, void* stream_) {
  return tfem_amg_aggregate_masked(nb, bptr, bcol, distance, nullptr, state_work, flag_work, index_work, agg,
                                   n_agg_host, rounds_host, stream_);
}

extern "C" int tfem_amg_aggregate_masked(int64_t nb, const int64_t* bptr, const int32_t* bcol, int distance,
                                         const uint8_t* exclude, int8_t* state_work, uint8_t* flag_work,
                                         int32_t* index_work, int32_t* agg, int64_t* n_agg_host,
                                         int32_t* rounds_host, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(bptr && bcol && state_work && flag_work && index_work && agg && n_agg_host && nb > 0,
               "amg_aggregate: bad arguments");
  TFEM_REQUIRE(distance == 1 || distance == 2, "amg_aggregate: distance must be 1 or 2");
  TFEM_REQUIRE(nb < (int64_t)INT32_MAX, "amg_aggregate: too many nodes");
  int* counter = nullptr;
  unsigned long long* t1 = nullptr;
  uint8_t* near1 = nullptr;
  int32_t* agg1 = nullptr;
  TFEM_CUDA(malloc_async(&counter, sizeof(int), st));
  if (distance == 2) {
    TFEM_CUDA(malloc_async(&t1, nb * sizeof(unsigned long long), st));
    TFEM_CUDA(malloc_async(&near1, nb, st));
    TFEM_CUDA(malloc_async(&agg1, nb * sizeof(int32_t), st));
  }
  auto release = [&]() {
    cudaFreeAsync(counter, st);
    if (t1) cudaFreeAsync(t1, st);
    if (near1) cudaFreeAsync(near1, st);
    if (agg1) cudaFreeAsync(agg1, st);
  };
  TFEM_CUDA(cudaMemsetAsync(state_work, 0, nb, st));
  const unsigned grid = grid_for(nb, 256);
  if (exclude) k_mis_exclude<<<grid, 256, 0, st>>>(nb, exclude, state_work);
  int rounds = 0, undecided = 1;
  while (undecided > 0) {
    if (rounds >= 200) {  // Luby rounds finish in O(log n) with overwhelming probability
      release();
      set_last_error("capacity", "amg_aggregate: independent set did not finish in 200 rounds");
      return TFEM_ERR_CAPACITY;
    }
    TFEM_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
    if (distance == 1) {
      k_mis_select<<<grid, 256, 0, st>>>(nb, bptr, bcol, state_work, flag_work);
      k_mis_apply<<<grid, 256, 0, st>>>(nb, bptr, bcol, state_work, flag_work, counter);
    } else {
      k_mis2_t1<<<grid, 256, 0, st>>>(nb, bptr, bcol, state_work, t1);
      k_mis2_select<<<grid, 256, 0, st>>>(nb, bptr, bcol, state_work, t1, flag_work);
      k_mis2_near<<<grid, 256, 0, st>>>(nb, bptr, bcol, flag_work, near1);
      k_mis2_apply<<<grid, 256, 0, st>>>(nb, bptr, bcol, state_work, flag_work, near1, counter);
    }
    TFEM_LAUNCH_CHECK();
    int rc = read_flag(counter, &undecided, st);
    if (rc != TFEM_OK) {
      release();
      return rc;
    }
    ++rounds;
  }
  k_mis_flags<<<grid, 256, 0, st>>>(nb, state_work, index_work);
  TFEM_LAUNCH_CHECK();
  int32_t last_flag = 0, last_idx = 0;
  TFEM_CUDA(cudaMemcpyAsync(&last_flag, index_work + nb - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  size_t bytes = 0;
  TFEM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, index_work, index_work, (int)nb, st));
  void* tmp = nullptr;
  TFEM_CUDA(malloc_async(&tmp, bytes ? bytes : 16, st));
  TFEM_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, index_work, index_work, (int)nb, st));
  TFEM_CUDA(cudaFreeAsync(tmp, st));
  TFEM_CUDA(cudaMemcpyAsync(&last_idx, index_work + nb - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (distance == 1) {
    k_mis_assign<<<grid, 256, 0, st>>>(nb, bptr, bcol, state_work, index_work, agg);
  } else {
    k_mis_assign<<<grid, 256, 0, st>>>(nb, bptr, bcol, state_work, index_work, agg1);
    k_mis2_assign_far<<<grid, 256, 0, st>>>(nb, bptr, bcol, agg1, agg);
  }
  TFEM_LAUNCH_CHECK();
  TFEM_CUDA(cudaStreamSynchronize(st));
  release();
  *n_agg_host = (int64_t)last_idx + last_flag;
  if (rounds_host) *rounds_host = rounds;
  return TFEM_OK;
}

template <bool FILL>
static int prolongator_launch(int d, int64_t nb, const int64_t* bptr, const int32_t* bcol, const double* vals,
                              const int32_t* agg, const double* dinv, const uint8_t* iso, double omega,
                              int64_t* pcount, const int64_t* pptr, int32_t* pcol, double* pvals, int max_row,
                              cudaStream_t st, int64_t row0 = 0) {
  int* err = nullptr;
  TFEM_CUDA(malloc_async(&err, sizeof(int), st));
  TFEM_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
  const unsigned grid = grid_for(nb, 4);
  // staging capacity = the longest block row, up to kStageMax (static ~30 KB + staging may exceed 48 KB: opt in)
  const int stage_cap = FILL ? (max_row < 8 ? 8 : (max_row > kStageMax ? kStageMax : max_row)) : 0;
  const size_t stage = FILL ? (size_t)4 * stage_cap * d * d * sizeof(double) : 0;
  if (FILL) {
    {  // (set on every call: the attribute is per device and costs nothing)
      if (d == 3) TFEM_CUDA(cudaFuncSetAttribute(k_prolongator<3, FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      else if (d == 2) TFEM_CUDA(cudaFuncSetAttribute(k_prolongator<2, FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      else TFEM_CUDA(cudaFuncSetAttribute(k_prolongator<1, FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    }
  }
  if (d == 3) k_prolongator<3, FILL><<<grid, 128, stage, st>>>(nb, bptr, bcol, vals, agg, dinv, iso, omega, pcount, pptr, pcol, pvals, err, stage_cap, row0);
  else if (d == 2) k_prolongator<2, FILL><<<grid, 128, stage, st>>>(nb, bptr, bcol, vals, agg, dinv, iso, omega, pcount, pptr, pcol, pvals, err, stage_cap, row0);
  else k_prolongator<1, FILL><<<grid, 128, stage, st>>>(nb, bptr, bcol, vals, agg, dinv, iso, omega, pcount, pptr, pcol, pvals, err, stage_cap, row0);
  TFEM_LAUNCH_CHECK();
  int h = 0;
  int rc = read_flag(err, &h, st);
  cudaFreeAsync(err, st);
  if (rc != TFEM_OK) return rc;
  if (h) {
    set_last_error("capacity", h == 1 ? "amg_prolongator: a node has more than 768 neighbours"
                                      : "amg_prolongator: a node touches more than 255 aggregates");
    return TFEM_ERR_CAPACITY;
  }
  return TFEM_OK;
}

extern "C" int tfem_amg_prolongator_count(int d, int64_t nb, const int64_t* bptr, const int32_t* bcol,
                                          const int32_t* agg, int64_t* pptr, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(bptr && bcol && agg && pptr && nb > 0 && d >= 1 && d <= 3, "amg_prolongator_count: bad arguments");
  int rc = prolongator_launch<false>(d, nb, bptr, bcol, nullptr, agg, nullptr, nullptr, 0.0, pptr + 1, nullptr,
                                     nullptr, nullptr, 0, st);
  if (rc != TFEM_OK) return rc;
  return scan_in_place(pptr, nb, st);
}

extern "C" int tfem_amg_prolongator_fill(int d, int64_t nb, const int64_t* bptr, const int32_t* bcol,
                                         const double* vals, const int32_t* agg, const double* dinv,
                                         const uint8_t* iso, double omega, const int64_t* pptr, int32_t* pcol,
                                         double* pvals, int max_row, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(bptr && bcol && vals && agg && dinv && iso && pptr && pcol && pvals && nb > 0 && d >= 1 && d <= 3,
               "amg_prolongator_fill: bad arguments");
  return prolongator_launch<true>(d, nb, bptr, bcol, vals, agg, dinv, iso, omega, nullptr, pptr, pcol, pvals, max_row, st);
}

// The same for the rows [row0, row0 + n_rows) of an operator whose other rows belong to other ranks (distributed
// hierarchy): `agg`, `dinv`, `iso` are indexed like the operator's rows / columns, pptr / pcol / pvals from 0.
extern "C" int tfem_amg_prolongator_count_rows(int d, int64_t row0, int64_t n_rows, const int64_t* bptr,
                                               const int32_t* bcol, const int32_t* agg, int64_t* pptr, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(bptr && bcol && agg && pptr && n_rows > 0 && row0 >= 0 && d >= 1 && d <= 3,
               "amg_prolongator_count_rows: bad arguments");
  int rc = prolongator_launch<false>(d, n_rows, bptr, bcol, nullptr, agg, nullptr, nullptr, 0.0, pptr + 1, nullptr,
                                     nullptr, nullptr, 0, st, row0);
  if (rc != TFEM_OK) return rc;
  return scan_in_place(pptr, n_rows, st);
}

extern "C" int tfem_amg_prolongator_fill_rows(int d, int64_t row0, int64_t n_rows, const int64_t* bptr,
                                              const int32_t* bcol, const double* vals, const int32_t* agg,
                                              const double* dinv, const uint8_t* iso, double omega,
                                              const int64_t* pptr, int32_t* pcol, double* pvals, int max_row,
                                              void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(bptr && bcol && vals && agg && dinv && iso && pptr && pcol && pvals && n_rows > 0 && row0 >= 0 &&
                   d >= 1 && d <= 3, "amg_prolongator_fill_rows: bad arguments");
  return prolongator_launch<true>(d, n_rows, bptr, bcol, vals, agg, dinv, iso, omega, nullptr, pptr, pcol, pvals,
                                  max_row, st, row0);
}

extern "C" int tfem_amg_transpose_structure(int64_t n_rows, int64_t n_cols, const int64_t* ptr, const int32_t* col,
                                            int64_t nblk, int64_t* tptr, int32_t* tcol, int32_t* tsrc,
                                            void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(ptr && col && tptr && tcol && tsrc && n_rows > 0 && n_cols > 0, "amg_transpose_structure: bad arguments");
  TFEM_REQUIRE(nblk < (int64_t)INT32_MAX, "amg_transpose_structure: too many blocks");
  TFEM_CUDA(cudaMemsetAsync(tptr, 0, (n_cols + 1) * sizeof(int64_t), st));
  if (nblk == 0) return TFEM_OK;
  k_bt_count<<<grid_for(nblk, 256), 256, 0, st>>>(nblk, col, tptr);
  TFEM_LAUNCH_CHECK();
  int rc = scan_in_place(tptr, n_cols, st);
  if (rc != TFEM_OK) return rc;
  int32_t *cursor = nullptr, *tmp = nullptr;
  TFEM_CUDA(malloc_async(&cursor, n_cols * sizeof(int32_t), st));
  TFEM_CUDA(malloc_async(&tmp, 2 * nblk * sizeof(int32_t), st));
  TFEM_CUDA(cudaMemsetAsync(cursor, 0, n_cols * sizeof(int32_t), st));
  k_bt_fill<<<grid_for(n_rows, 128), 128, 0, st>>>(n_rows, ptr, col, tptr, cursor, tmp, tmp + nblk);
  k_bt_sort<<<grid_for(n_cols * 32, 256), 256, 0, st>>>(n_cols, tptr, tmp, tmp + nblk, tcol, tsrc);
  TFEM_LAUNCH_CHECK();
  TFEM_CUDA(cudaFreeAsync(cursor, st));
  TFEM_CUDA(cudaFreeAsync(tmp, st));
  return TFEM_OK;
}

extern "C" int tfem_amg_transpose_values(int d, int64_t n_cols, const int64_t* ptr, const double* vals,
                                         const int64_t* tptr, const int32_t* tcol, const int32_t* tsrc,
                                         double* tvals, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(ptr && vals && tptr && tcol && tsrc && tvals && n_cols > 0 && d >= 1 && d <= 3,
               "amg_transpose_values: bad arguments");
  const unsigned grid = grid_for(n_cols * 32, 256);
  if (d == 3) k_bt_vals<3><<<grid, 256, 0, st>>>(n_cols, ptr, vals, tptr, tcol, tsrc, tvals);
  else if (d == 2) k_bt_vals<2><<<grid, 256, 0, st>>>(n_cols, ptr, vals, tptr, tcol, tsrc, tvals);
  else k_bt_vals<1><<<grid, 256, 0, st>>>(n_cols, ptr, vals, tptr, tcol, tsrc, tvals);
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

template <bool FILL>
static int spgemm_sym_launch(int64_t nx, const int64_t* xptr, const int32_t* xcol, const int64_t* yptr,
                             const int32_t* ycol, int64_t* ccount, const int64_t* cptr, int32_t* ccol,
                             cudaStream_t st) {
  int* err = nullptr;
  TFEM_CUDA(malloc_async(&err, sizeof(int), st));
  for (int attempt = 0; attempt < 2; ++attempt) {
    TFEM_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
    if (attempt == 0) {
      constexpr int HC = 1024, W = 4;
      const size_t smem = (2 * W * HC + W) * sizeof(int);
      const int64_t want = (nx + W - 1) / W;
      const int64_t cap = (int64_t)num_sms() * 6;
      k_spgemm_sym<HC, FILL><<<(unsigned)(want < cap ? want : cap), W * 32, smem, st>>>(nx, xptr, xcol, yptr, ycol,
                                                                                        ccount, cptr, ccol, err);
    } else {
      constexpr int HC = 16384, W = 1;
      const size_t smem = (2 * W * HC + W) * sizeof(int);
      TFEM_CUDA(cudaFuncSetAttribute(k_spgemm_sym<HC, FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int64_t cap = (int64_t)num_sms();
      k_spgemm_sym<HC, FILL><<<(unsigned)(nx < cap ? nx : cap), W * 32, smem, st>>>(nx, xptr, xcol, yptr, ycol, ccount,
                                                                                    cptr, ccol, err);
    }
    TFEM_LAUNCH_CHECK();
    int h = 0;
    int rc = read_flag(err, &h, st);
    if (rc != TFEM_OK) return rc;
    if (!h) {
      cudaFreeAsync(err, st);
      return TFEM_OK;
    }
  }
  cudaFreeAsync(err, st);
  set_last_error("capacity", "amg_spgemm: a product row has more than 12288 distinct block columns");
  return TFEM_ERR_CAPACITY;
}

extern "C" int tfem_amg_spgemm_count(int64_t nx, const int64_t* xptr, const int32_t* xcol, const int64_t* yptr,
                                     const int32_t* ycol, int64_t* cptr, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(xptr && xcol && yptr && ycol && cptr && nx > 0, "amg_spgemm_count: bad arguments");
  int rc = spgemm_sym_launch<false>(nx, xptr, xcol, yptr, ycol, cptr + 1, nullptr, nullptr, st);
  if (rc != TFEM_OK) return rc;
  return scan_in_place(cptr, nx, st);
}

extern "C" int tfem_amg_spgemm_fill(int64_t nx, const int64_t* xptr, const int32_t* xcol, const int64_t* yptr,
                                    const int32_t* ycol, const int64_t* cptr, int32_t* ccol, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(xptr && xcol && yptr && ycol && cptr && ccol && nx > 0, "amg_spgemm_fill: bad arguments");
  return spgemm_sym_launch<true>(nx, xptr, xcol, yptr, ycol, nullptr, cptr, ccol, st);
}

template <int D, int G>
static int spgemm_num_launch(int64_t nx, const int64_t* xptr, const int32_t* xcol, const double* xvals,
                             const int64_t* yptr, const int32_t* ycol, const double* yvals, const int64_t* cptr,
                             const int32_t* ccol, double* cvals, int max_row, cudaStream_t st) {
  int ht = 0;  // hash-table entries per warp (warp variant only): power of two >= 2 max_row
  if (G == 32) {
    ht = 16;
    while (ht < 2 * max_row) ht <<= 1;
  }
  const size_t per_group = (size_t)max_row * (D * D * sizeof(double) + sizeof(int)) + (size_t)ht * sizeof(long long);
  if (per_group > 200 * 1024) {
    set_last_error("capacity", "amg_spgemm_numeric: a product row does not fit in shared memory");
    return TFEM_ERR_CAPACITY;
  }
  int W = 256 / G;
  if (G == 32) {
    // ~4 CTAs per SM. 44 KB, not 48: the kernel's static shared memory counts against the 48 KB a launch may use
    // without opting in (max_row = 78 gave 47,888 B dynamic + static > 48 KB -> "invalid argument" at launch)
    const int fit = (int)((44 * 1024) / (per_group ? per_group : 1));
    W = fit < 1 ? 1 : (fit > 8 ? 8 : fit);
  }
  const size_t smem = W * per_group + 32;
  if (smem > 40 * 1024)
    TFEM_CUDA(cudaFuncSetAttribute(k_spgemm_num<D, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(208 * 1024)));
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_spgemm_num<D, G>, W * G, smem) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  const int64_t want = (nx + W - 1) / W;
  const int64_t cap = (int64_t)num_sms() * per_sm;
  k_spgemm_num<D, G><<<(unsigned)(want < cap ? want : cap), W * G, smem, st>>>(nx, xptr, xcol, xvals, yptr, ycol,
                                                                               yvals, cptr, ccol, cvals, max_row, ht);
  const cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) {
    char msg[256];
    snprintf(msg, sizeof(msg), "k_spgemm_num<%d,%d> launch: %s (nx %lld, max_row %d, W %d, smem %zu, grid %lld, per_sm %d)",
             D, G, cudaGetErrorString(le), (long long)nx, max_row, W, smem, (long long)(want < cap ? want : cap), per_sm);
    set_last_error("cuda", msg);
    return TFEM_ERR_CUDA;
  }
  return TFEM_OK;
}

extern "C" int tfem_amg_spgemm_numeric(int d, int64_t nx, const int64_t* xptr, const int32_t* xcol,
                                       const double* xvals, const int64_t* yptr, const int32_t* ycol,
                                       const double* yvals, const int64_t* cptr, const int32_t* ccol, double* cvals,
                                       int max_row, int threads_per_row, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(xptr && xcol && xvals && yptr && ycol && yvals && cptr && ccol && cvals && nx > 0 && max_row > 0,
               "amg_spgemm_numeric: bad arguments");
  TFEM_REQUIRE(d >= 1 && d <= 3, "amg: 1, 2 or 3 DOFs per node");
  // a warp per row cannot hold rows whose accumulators exceed ~40 KB: those go to the CTA-per-row variant too
  const bool cta = threads_per_row > 32 || (size_t)max_row * (d * d * 8 + 4 + 32) > 24 * 1024;
#define TFEM_SPGEMM_CASE(D)                                                                                         \
  if (d == D)                                                                                                       \
    return cta ? spgemm_num_launch<D, 256>(nx, xptr, xcol, xvals, yptr, ycol, yvals, cptr, ccol, cvals, max_row, st) \
               : spgemm_num_launch<D, 32>(nx, xptr, xcol, xvals, yptr, ycol, yvals, cptr, ccol, cvals, max_row, st);
  TFEM_SPGEMM_CASE(3)
  TFEM_SPGEMM_CASE(2)
  TFEM_SPGEMM_CASE(1)
#undef TFEM_SPGEMM_CASE
  return TFEM_ERR_INVALID;
}

extern "C" int tfem_amg_vcycle(const tfem_amg_level_t* levels, int n_levels, const double* coarse_inv,
                               const double* r, double* z, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(levels && n_levels >= 1 && n_levels <= TFEM_AMG_MAX_LEVELS && coarse_inv && r && z && r != z,
               "amg_vcycle: bad arguments");
  Level L[TFEM_AMG_MAX_LEVELS];
  int rc = make_levels(levels, n_levels, L);
  if (rc != TFEM_OK) return rc;
  int64_t launches = 0;
  return vcycle(L, n_levels, coarse_inv, r, z, nullptr, nullptr, nullptr, &launches, st);
}

extern "C" int tfem_amg_vcycle_block(const tfem_amg_level_t* levels, int n_levels, const double* coarse_inv, int nb,
                                     const double* r_block, double* z_block, double* work, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(levels && n_levels >= 1 && n_levels <= TFEM_AMG_MAX_LEVELS && coarse_inv && r_block && z_block && work &&
                   r_block != z_block,
               "amg_vcycle_block: bad arguments");
  TFEM_REQUIRE(nb >= 1 && nb <= kSpmmBlock, "amg_vcycle_block: 1 to 4 vectors per call");
  Level L[TFEM_AMG_MAX_LEVELS];
  int rc = make_levels(levels, n_levels, L);
  if (rc != TFEM_OK) return rc;
  return vcycle_block(L, n_levels, coarse_inv, nb, r_block, z_block, work, st);
}

extern "C" int tfem_amg_pcg_solve(const tfem_amg_level_t* levels, int n_levels, const double* coarse_inv,
                                  const double* b, const double* x0, double rtol, double atol, int64_t maxiter,
                                  double* x, double* work, double* info, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(levels && n_levels >= 1 && n_levels <= TFEM_AMG_MAX_LEVELS && coarse_inv && b && x && work && info,
               "amg_pcg_solve: bad arguments");
  Level L[TFEM_AMG_MAX_LEVELS];
  int rc = make_levels(levels, n_levels, L);
  if (rc != TFEM_OK) return rc;
  const Sell& A = L[0].A.sell;
  const int64_t n = A.n, np = pad32(n);
  if (maxiter <= 0) maxiter = 10 * n;
  double *r = work, *p = work + np, *q = work + 2 * np, *z = work + 3 * np, *sc = work + 4 * np;
  double* partials = sc + P_COUNT;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(partials + kMaxPartials);
  TFEM_CUDA(cudaMemsetAsync(sc, 0, (P_COUNT + kMaxPartials + 32) * sizeof(double), st));
  const int vg = vec_grid(n);
  int64_t launches = 0, spmvs = 0;

  if (x0) TFEM_CUDA(cudaMemcpyAsync(x, x0, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  else TFEM_CUDA(cudaMemsetAsync(x, 0, n * sizeof(double), st));
  const double* q0 = nullptr;
  if (x0) {
    if ((rc = launch_amg_spmv<M_AX, false>(A, x, q, nullptr, nullptr, 0.0, nullptr, nullptr, nullptr, st))) return rc;
    q0 = q;
    ++spmvs;
    ++launches;
  }
  k_pcg_init<<<vg, kVecThreads, 0, st>>>(n, b, q0, r, sc, rtol, atol, partials, ticket);
  TFEM_LAUNCH_CHECK();
  ++launches;
  double h[P_COUNT];
  TFEM_CUDA(cudaMemcpyAsync(h, sc, sizeof(h), cudaMemcpyDeviceToHost, st));
  TFEM_CUDA(cudaStreamSynchronize(st));
  int64_t it = 0;
  while (h[P_DONE] == 0.0 && it < maxiter) {
    if ((rc = vcycle(L, n_levels, coarse_inv, r, z, partials, ticket, sc + P_RHO_NEW, &launches, st))) return rc;
    k_pcg_direction<<<vg, kVecThreads, 0, st>>>(n, z, p, sc, it == 0 ? 1 : 0);
    k_pcg_roll<<<1, 1, 0, st>>>(sc);
    if ((rc = launch_amg_spmv<M_AX, true>(A, p, q, nullptr, nullptr, 0.0, partials, ticket, sc + P_PQ, st))) return rc;
    k_pcg_update<<<vg, kVecThreads, 0, st>>>(n, p, q, x, r, sc, partials, ticket);
    TFEM_LAUNCH_CHECK();
    launches += 4;
    spmvs += 3;
    ++it;
    TFEM_CUDA(cudaMemcpyAsync(h, sc, sizeof(h), cudaMemcpyDeviceToHost, st));
    TFEM_CUDA(cudaStreamSynchronize(st));
  }
  info[0] = h[P_ITERS];
  info[1] = sqrt(h[P_RR]);
  info[2] = h[P_BNRM];
  info[3] = h[P_DONE] == 1.0 ? 1.0 : 0.0;
  info[4] = (double)spmvs;
  info[5] = (double)launches;
  info[6] = h[P_DONE];
  info[7] = 0.0;
  if (h[P_DONE] == 2.0) {
    set_last_error("breakdown", "non-finite residual or non-positive curvature (matrix or preconditioner not SPD?)");
    return TFEM_ERR_BREAKDOWN;
  }
  if (h[P_DONE] != 1.0) {
    set_last_error("not converged", "iteration limit reached");
    return TFEM_ERR_NOT_CONVERGED;
  }
  return TFEM_OK;
}

// ---------------------------------------------------------------------------------------------------
// Distributed AMG-PCG: the cycle above with its exchange steps over peer memory (peer.cuh).
// ---------------------------------------------------------------------------------------------------
namespace tfem {
namespace {

enum { AR_INIT = 0, AR_RHO, AR_PQ, AR_RR };

// One CTA: publish this rank's local sums (LL protocol), wait for all ranks, sum in rank order, apply the scalar
// recurrence the single-GPU kernels apply in their last CTA. Bit-identical on every rank.
template <int K>
__global__ void k_allreduce(Peers P, int set, unsigned long long epoch, const double* local, double* sc, int op,
                            double rtol, double atol) {
  __shared__ double s_out[4];
  __shared__ int s_ok;
  if (sc[P_DONE] == 4.0) return;
  if (threadIdx.x < 32) {
    double v[K];
#pragma unroll
    for (int j = 0; j < K; ++j) v[j] = local[j];
    ll_publish<K>(P, set, epoch, v);
  }
  double tot[K];
  if (!ll_wait_sum<K>(P, set, epoch, tot, s_out, &s_ok)) {
    if (threadIdx.x == 0) sc[P_DONE] = 4.0;
    return;
  }
  if (threadIdx.x == 0) {
    if (op == AR_INIT) pcg_scalars_init(sc, tot[0], K > 1 ? tot[K - 1] : 0.0, rtol, atol);
    else if (op == AR_RHO) sc[P_RHO_NEW] = tot[0];
    else if (op == AR_PQ) sc[P_PQ] = tot[0];
    else pcg_scalars_update(sc, tot[0]);
  }
}

// Halo exchange of a heap vector in ONE kernel: the CTAs store my boundary entries into the neighbours' copies (the last
// one releases the channel flag there), then thread 0 of CTA 0 waits until every neighbour has delivered — so the next
// kernel on the stream may read the halo. No deadlock: every rank's stores precede its wait and depend on nothing remote.
__global__ void __launch_bounds__(kVecThreads)
    k_halo_exchange(Peers P, Halo H, int64_t vec_off, int channel, unsigned long long epoch, unsigned int* halo_ticket,
                    double* sc) {
  if (sc[P_DONE] == 4.0) return;
  if (H.n_send > 0 && H.send_total > 0) {
    const double* v = heap(P, P.rank) + vec_off;
    halo_send_and_release(P, H, vec_off, channel, epoch, (int)gridDim.x, halo_ticket, [&](int64_t i) { return v[i]; });
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && H.n_recv > 0 && !halo_wait_thread(P, H, channel, epoch)) sc[P_DONE] = 4.0;
}

__global__ void __launch_bounds__(kVecThreads)
    k_zero(int64_t n, double* __restrict__ x) {
  for (int64_t i = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kVecThreads) x[i] = 0.0;
}

struct DLevel {
  Level L;
  int64_t lo, hi, c_lo, c_hi;
  Halo H;
  int64_t x_off, t_off;  // heap offsets of L.x / L.t
};

struct DCtx {
  Comm* c;
  Peers P;
  double* heap_base;
  double* sc;
  unsigned int* halo_ticket;
  cudaStream_t st;
  int64_t launches;
};

void restrict_rows(Oper& A, int64_t lo, int64_t hi) {  // scalar rows [lo, hi) are computed and written
  if (A.bcsr) {
    const int d = A.blk.d;
    A.blk.row_lo = lo / d;
    A.blk.row_hi = (hi + d - 1) / d;
  } else {
    A.sell.dot_lo = lo;
    A.sell.dot_hi = hi;
    A.sell.slice_lo = lo / 32;
    A.sell.slice_hi = (hi + 31) / 32;
  }
}

int halo_exchange(DCtx& X, const Halo& H, int64_t vec_off, int channel) {
  if (X.c->world == 1) return TFEM_OK;
  const unsigned long long ep = ++X.c->chan_epoch[channel];
  if ((H.n_send > 0 && H.send_total > 0) || H.n_recv > 0) {
    int g = H.send_total > 0 ? halo_ctas(H.send_total, kVecThreads) : 1;
    if (g > 64) g = 64;
    k_halo_exchange<<<g, kVecThreads, 0, X.st>>>(X.P, H, vec_off, channel, ep, X.halo_ticket, X.sc);
    ++X.launches;
  }
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

template <int K>
int allreduce(DCtx& X, const double* local, int op, double rtol, double atol) {
  const unsigned long long ep = ++X.c->epoch;
  k_allreduce<K><<<1, 32, 0, X.st>>>(X.P, (int)(ep & 1ull), ep, local, X.sc, op, rtol, atol);
  ++X.launches;
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

constexpr int CH_CG_P = 0, CH_GATHER = 15;
inline int ch_x(int l) { return 1 + 3 * l; }
inline int ch_t(int l) { return 2 + 3 * l; }
inline int ch_xc(int l) { return 3 + 3 * l; }   // level l's result vector as the coarse input of level l-1

// z = M r with the distributed levels D[0..nd) and the replicated tail T[0..nt). r, z: level-0 local vectors.
int dvcycle(DCtx& X, DLevel* D, int nd, const Level* T, int nt, const double* tail_inv, const Halo& gatherH,
            int64_t tail_b_off, double* tail_b, double* tail_x, const double* r, double* z, double* partials,
            unsigned int* ticket, double* dot_out) {
  int rc;
  cudaStream_t st = X.st;
  for (int l = 0; l < nd; ++l) {  // downward leg
    DLevel& d = D[l];
    const Level& v = d.L;
    const int64_t n = d.hi - d.lo;
    const double* b = l == 0 ? r : v.b;
    k_jacobi_first<<<vec_grid(n), kVecThreads, 0, st>>>(n, v.omega, v.dinv + d.lo, b + d.lo, v.x + d.lo);
    if ((rc = halo_exchange(X, d.H, d.x_off, ch_x(l)))) return rc;
    if ((rc = apply_oper<M_RES>(v.A, v.x, v.t, b, nullptr, 0.0, st))) return rc;
    if ((rc = halo_exchange(X, d.H, d.t_off, ch_t(l)))) return rc;
    double* nb = l + 1 < nd ? D[l + 1].L.b : tail_b;
    if ((rc = apply_oper<M_AX>(v.R, v.t, nb, nullptr, nullptr, 0.0, st))) return rc;
    X.launches += 3;
  }
  // the tail: every rank stores its segment of the right-hand side into every rank's copy, then solves redundantly
  if ((rc = halo_exchange(X, gatherH, tail_b_off, CH_GATHER))) return rc;
  if ((rc = vcycle(T, nt, tail_inv, tail_b, tail_x, nullptr, nullptr, nullptr, &X.launches, st))) return rc;
  for (int l = nd - 1; l >= 0; --l) {  // upward leg
    DLevel& d = D[l];
    const Level& v = d.L;
    const double* b = l == 0 ? r : v.b;
    const double* xc = tail_x;
    if (l + 1 < nd) {
      xc = D[l + 1].L.t;
      if ((rc = halo_exchange(X, D[l + 1].H, D[l + 1].t_off, ch_xc(l + 1)))) return rc;
    }
    if ((rc = apply_oper<M_ADD>(v.P, xc, v.x, nullptr, nullptr, 0.0, st))) return rc;
    if ((rc = halo_exchange(X, d.H, d.x_off, ch_x(l)))) return rc;
    double* out = l == 0 ? z : v.t;
    if (l == 0 && dot_out)
      rc = launch_amg_spmv<M_JAC, true>(v.A.sell, v.x, out, b, v.dinv, v.omega, partials, ticket, dot_out, st);
    else
      rc = apply_oper<M_JAC>(v.A, v.x, out, b, v.dinv, v.omega, st);
    if (rc) return rc;
    X.launches += 2;
  }
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

}  // namespace
}  // namespace tfem

extern "C" int tfem_damg_pcg_solve(void* comm, const tfem_damg_level_t* levels, int n_levels,
                                   const tfem_amg_level_t* tail_levels, int n_tail, const double* tail_inv,
                                   int64_t tail_n, double* tail_b_heap, double* tail_x, const double* b, double* x,
                                   double* p_heap, double* work, double rtol, double atol, int64_t maxiter,
                                   double timeout_s, double* info, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  Comm* c = static_cast<Comm*>(comm);
  TFEM_REQUIRE(c && c->connected, "damg_pcg_solve: communicator missing or not connected");
  TFEM_REQUIRE(!c->broken, "damg_pcg_solve: communicator is out of step after a failed solve; create a new one");
  TFEM_REQUIRE(levels && n_levels >= 1 && n_levels <= 4 && tail_levels && n_tail >= 1 && n_tail <= TFEM_AMG_MAX_LEVELS &&
                   tail_inv && tail_b_heap && tail_x && b && x && p_heap && work && info && tail_n > 0,
               "damg_pcg_solve: bad arguments");
  TFEM_REQUIRE(c->world - 1 <= kMaxNbr, "damg_pcg_solve: the coarse gather supports at most 9 ranks");
  if (!(timeout_s > 0.0)) timeout_s = 20.0;
  double* heap_base = reinterpret_cast<double*>(c->base[c->rank] + HEADER_BYTES);
  auto in_heap = [&](const double* p, int64_t n) {
    return p >= heap_base && p + n <= heap_base + c->heap_doubles;
  };

  DLevel D[4];
  Level T[TFEM_AMG_MAX_LEVELS];
  int rc;
  for (int l = 0; l < n_levels; ++l) {
    const tfem_damg_level_t& dl = levels[l];
    Level& v = D[l].L;
    if ((rc = make_oper(&dl.lv.A, &v.A)) || (rc = make_oper(&dl.lv.P, &v.P)) || (rc = make_oper(&dl.lv.R, &v.R))) return rc;
    v.n = v.A.n;
    v.dinv = dl.lv.dinv;
    v.omega = dl.lv.omega;
    v.x = dl.lv.x;
    v.b = dl.lv.b;
    v.t = dl.lv.t;
    TFEM_REQUIRE(v.dinv && v.x && v.t && (l == 0 || v.b), "damg level: null vector");
    TFEM_REQUIRE(in_heap(v.x, v.n) && in_heap(v.t, v.n), "damg level: x and t must live in the communicator's heap");
    TFEM_REQUIRE(dl.own_lo >= 0 && dl.own_lo < dl.own_hi && dl.own_hi <= v.n, "damg level: bad owned range");
    TFEM_REQUIRE(v.P.n == v.n, "damg level: P must have the level's local rows");
    D[l].lo = dl.own_lo;
    D[l].hi = dl.own_hi;
    D[l].c_lo = dl.c_own_lo;
    D[l].c_hi = dl.c_own_hi;
    D[l].x_off = v.x - heap_base;
    D[l].t_off = v.t - heap_base;
    if ((rc = fill_halo(D[l].H, c, dl.n_sends, dl.sends, dl.n_recv, dl.recv_peers))) return rc;
    restrict_rows(v.A, dl.own_lo, dl.own_hi);
    restrict_rows(v.P, dl.own_lo, dl.own_hi);
    restrict_rows(v.R, dl.c_own_lo, dl.c_own_hi);
    if (l > 0) TFEM_REQUIRE(D[l - 1].L.R.n == v.n && D[l - 1].c_lo == dl.own_lo && D[l - 1].c_hi == dl.own_hi,
                            "damg level: R of the finer level does not match this level");
  }
  TFEM_REQUIRE(!D[0].L.A.bcsr, "damg: the finest level must be given in SELL-32 form");
  if ((rc = make_levels(tail_levels, n_tail, T, false))) return rc;
  TFEM_REQUIRE(T[0].n == tail_n && D[n_levels - 1].L.R.n == tail_n, "damg: tail size mismatch");
  TFEM_REQUIRE(in_heap(tail_b_heap, tail_n), "damg: the gathered right-hand side must live in the heap");
  TFEM_REQUIRE(in_heap(p_heap, D[0].L.n), "damg: p must live in the heap");

  // the gather of the tail's right-hand side as a halo plan: my segment goes to every other rank, same offsets
  Halo G;
  memset(&G, 0, sizeof(G));
  const int64_t seg_lo = D[n_levels - 1].c_lo, seg_n = D[n_levels - 1].c_hi - D[n_levels - 1].c_lo;
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    const int s = G.n_send++;
    G.send_peer[s] = r;
    G.send_count[s] = seg_n;
    G.src0[s] = seg_lo;
    G.dst0[s] = seg_lo;
    G.send_total += seg_n;
    G.recv_peer[G.n_recv++] = r;
  }

  const Sell& A = D[0].L.A.sell;
  const int64_t n = D[0].L.n, np = pad32(n), lo = D[0].lo, no = D[0].hi - D[0].lo;
  if (maxiter <= 0) maxiter = 10 * n * c->world;
  double *r = work, *q = work + 2 * np, *z = work + 3 * np, *sc = work + 4 * np, *p = p_heap;
  double* partials = sc + P_COUNT;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(partials + kMaxPartials);
  double* loc = sc + 8;   // P_COUNT = 16 scalars: slots 8..11 carry the local sums handed to k_allreduce
  TFEM_CUDA(cudaMemsetAsync(sc, 0, (P_COUNT + kMaxPartials + 32) * sizeof(double), st));
  DCtx X;
  X.c = c;
  X.P = make_peers(c, timeout_s);
  X.heap_base = heap_base;
  X.sc = sc;
  X.halo_ticket = ticket + 1;
  X.st = st;
  X.launches = 0;
  const int vg = vec_grid(no);
  int64_t spmvs = 0;

  k_zero<<<vg, kVecThreads, 0, st>>>(no, x + lo);
  k_pcg_init<<<vg, kVecThreads, 0, st>>>(no, b + lo, nullptr, r + lo, sc, rtol, atol, partials, ticket, loc);
  TFEM_LAUNCH_CHECK();
  if ((rc = allreduce<2>(X, loc, AR_INIT, rtol, atol))) return rc;
  X.launches += 2;
  double h[P_COUNT];
  TFEM_CUDA(cudaMemcpyAsync(h, sc, sizeof(h), cudaMemcpyDeviceToHost, st));
  TFEM_CUDA(cudaStreamSynchronize(st));
  int64_t it = 0;
  const int64_t tail_b_off = tail_b_heap - heap_base, p_off = p_heap - heap_base;
  while (h[P_DONE] == 0.0 && it < maxiter) {
    if ((rc = dvcycle(X, D, n_levels, T, n_tail, tail_inv, G, tail_b_off, tail_b_heap, tail_x, r, z, partials, ticket,
                      loc))) { c->broken = true; return rc; }
    if ((rc = allreduce<1>(X, loc, AR_RHO, rtol, atol))) return rc;
    k_pcg_direction<<<vg, kVecThreads, 0, st>>>(no, z + lo, p + lo, sc, it == 0 ? 1 : 0);
    k_pcg_roll<<<1, 1, 0, st>>>(sc);
    if ((rc = halo_exchange(X, D[0].H, p_off, CH_CG_P))) return rc;
    if ((rc = launch_amg_spmv<M_AX, true>(A, p, q, nullptr, nullptr, 0.0, partials, ticket, loc, st))) return rc;
    if ((rc = allreduce<1>(X, loc, AR_PQ, rtol, atol))) return rc;
    k_pcg_update<<<vg, kVecThreads, 0, st>>>(no, p + lo, q + lo, x + lo, r + lo, sc, partials, ticket, loc);
    TFEM_LAUNCH_CHECK();
    if ((rc = allreduce<1>(X, loc, AR_RR, rtol, atol))) return rc;
    X.launches += 4;
    spmvs += 3;
    ++it;
    TFEM_CUDA(cudaMemcpyAsync(h, sc, sizeof(h), cudaMemcpyDeviceToHost, st));
    TFEM_CUDA(cudaStreamSynchronize(st));
  }
  info[0] = h[P_ITERS];
  info[1] = sqrt(h[P_RR]);
  info[2] = h[P_BNRM];
  info[3] = h[P_DONE] == 1.0 ? 1.0 : 0.0;
  info[4] = (double)spmvs;
  info[5] = (double)X.launches;
  info[6] = h[P_DONE];
  info[7] = 0.0;
  if (h[P_DONE] == 4.0) {
    c->broken = true;
    set_last_error("communication", "a peer did not deliver its halo / reduction within the timeout");
    return TFEM_ERR_COMM;
  }
  if (h[P_DONE] == 2.0) {
    set_last_error("breakdown", "non-finite residual or non-positive curvature (matrix or preconditioner not SPD?)");
    return TFEM_ERR_BREAKDOWN;
  }
  if (h[P_DONE] != 1.0) {
    set_last_error("not converged", "iteration limit reached");
    return TFEM_ERR_NOT_CONVERGED;
  }
  return TFEM_OK;
}
