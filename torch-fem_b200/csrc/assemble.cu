// K2/K3 — deterministic assembly of element matrices into CSR values, fused with Dirichlet masking.
//
// Replaces FEM.assemble_matrix (src/torchfem/base.py:398-426): `val.index_add_(0, k_map, k.ravel())`
// (atomicAdd scatter on CUDA: non-deterministic FP order) + two [nnz] int64 gathers for the row/column
// masks. Here every CSR entry is OWNED by one thread, which sums its element contributions in the fixed
// order of the precomputed `src` permutation (ascending element, then local slot — the order in which
// the reference's CPU index_add_ meets them), so the result is bitwise reproducible.
//
// One warp per node: its dpn rows are contiguous in `vals`, so the warp writes them coalesced; the
// contributions of one element to these rows are dpn consecutive rows of k_e (contiguous dpn*nd
// doubles), so the reads of neighbouring lanes fall into the same lines. Each k_e entry is read once:
// HBM traffic = 8 B * n_elem*nd^2 (read) + 4 B * n_elem*nn^2 (src) + 8 B * nnz (write).
#include "common.cuh"

namespace tfem {
namespace {

constexpr int kWarps = 8;
#ifndef TFEM_K2_INFLIGHT
#define TFEM_K2_INFLIGHT 2
#endif
#ifndef TFEM_K2_SOLVER_MINB
#define TFEM_K2_SOLVER_MINB 4  // resident CTAs the solver-order instantiation must allow (64 registers)
#endif
constexpr int kInFlight = TFEM_K2_INFLIGHT;  // contributions whose loads are issued before the first is added

// NN = nodes per element as a compile-time constant (0: run-time value) so the slot decoding needs no
// integer division. One lane per node BLOCK (neighbour p of the node): it decodes every contribution once and
// accumulates the whole dpn x dpn block (the previous version decoded it once per scalar entry and was bound by
// instruction issue: 71 % of the issue slots at config B); each entry still receives its contributions in slot
// order, two contributions in flight.
template <int DPN, int NN>
__global__ void __launch_bounds__(kWarps * 32)
    k_assemble(int64_t n_nod, int nn_rt, const int64_t* __restrict__ node_ptr,
               const int32_t* __restrict__ adj, const int64_t* __restrict__ indptr,
               const int64_t* __restrict__ src_ptr, const int32_t* __restrict__ src,
               const double* __restrict__ k, const uint8_t* __restrict__ is_con,
               const double* __restrict__ ubc, double* __restrict__ vals, double* __restrict__ lift) {
  const int nn = NN > 0 ? NN : nn_rt;
  const int lane = threadIdx.x & 31;
  const int64_t node = blockIdx.x * (int64_t)kWarps + (threadIdx.x >> 5);
  if (node >= n_nod) return;
  const int64_t nb = node_ptr[node];
  const int cnt = (int)(node_ptr[node + 1] - nb);
  const int64_t row0 = node * DPN;
  if (cnt == 0) {  // unreferenced node: lone diagonal, value 0 unless constrained (base.py:419)
    if (lane < DPN) {
      vals[indptr[row0 + lane]] = (is_con && is_con[row0 + lane]) ? 1.0 : 0.0;
      if (lift) lift[row0 + lane] = 0.0;
    }
    return;
  }
  const int nd = nn * DPN;
  const int64_t nd2 = (int64_t)nd * nd;
  bool row_con[DPN];
  int64_t rp[DPN];
  double lsum[DPN];
#pragma unroll
  for (int i = 0; i < DPN; ++i) {
    row_con[i] = is_con && is_con[row0 + i];
    rp[i] = indptr[row0 + i];
    lsum[i] = 0.0;
  }
  for (int p = lane; p < cnt; p += 32) {
    const int64_t sb = src_ptr[nb + p], se = src_ptr[nb + p + 1];
    double acc[DPN][DPN];
#pragma unroll
    for (int i = 0; i < DPN; ++i)
#pragma unroll
      for (int j = 0; j < DPN; ++j) acc[i][j] = 0.0;
    for (int64_t s = sb; s < se; s += kInFlight) {
      const double* kb[kInFlight];
#pragma unroll
      for (int u = 0; u < kInFlight; ++u) {
        const int cc = (s + u < se) ? src[s + u] : -1;  // e*nn*nn + a*nn + b
        if (cc < 0) {
          kb[u] = nullptr;
        } else {
          const int ea = cc / nn, bq = cc - ea * nn;
          const int e = ea / nn, aq = ea - e * nn;
          kb[u] = k + e * nd2 + (int64_t)(aq * DPN) * nd + bq * DPN;
        }
      }
      double v[kInFlight][DPN][DPN];
#pragma unroll
      for (int u = 0; u < kInFlight; ++u)
#pragma unroll
        for (int i = 0; i < DPN; ++i)
#pragma unroll
          for (int j = 0; j < DPN; ++j) v[u][i][j] = kb[u] ? kb[u][i * nd + j] : 0.0;
#pragma unroll
      for (int u = 0; u < kInFlight; ++u)  // + 0.0 for the missing ones: exact; slot order kept
#pragma unroll
        for (int i = 0; i < DPN; ++i)
#pragma unroll
          for (int j = 0; j < DPN; ++j) acc[i][j] += v[u][i][j];
    }
    const int64_t col0 = (int64_t)adj[nb + p] * DPN;
#pragma unroll
    for (int j = 0; j < DPN; ++j) {
      const int64_t col = col0 + j;
      const bool col_con = is_con && is_con[col];
      const double uc = (lift && col_con) ? ubc[col] : 0.0;
#pragma unroll
      for (int i = 0; i < DPN; ++i) {
        double a = acc[i][j];
        // Dirichlet lifting: what the prescribed values contribute to the free rows, K[row, con] u[con]
        if (lift && col_con && !row_con[i]) lsum[i] = fma(a, uc, lsum[i]);
        if (row_con[i] || col_con) a = (col == row0 + i) ? 1.0 : 0.0;
        vals[rp[i] + (int64_t)p * DPN + j] = a;
      }
    }
  }
  if (lift) {
#pragma unroll
    for (int i = 0; i < DPN; ++i) {
      const double t = warp_sum(lsum[i]);
      if (lane == 0) lift[row0 + i] = t;
    }
  }
}

// The same, also (or only) writing the values in the solver's SELL-32 order and 1/diagonal. A separate kernel: with the
// extra outputs compiled into k_assemble the CSR-only path went from 6.5 to 9.3 ms at config B.
template <int DPN, int NN>
__global__ void __launch_bounds__(kWarps * 32, TFEM_K2_SOLVER_MINB)
    k_assemble_solver(int64_t n_nod, int nn_rt, const int64_t* __restrict__ node_ptr,
               const int32_t* __restrict__ adj, const int64_t* __restrict__ indptr,
               const int64_t* __restrict__ src_ptr, const int32_t* __restrict__ src,
               const double* __restrict__ k, const uint8_t* __restrict__ is_con,
               const double* __restrict__ ubc, double* __restrict__ vals, double* __restrict__ lift,
               const int64_t* __restrict__ slice_ptr, double* __restrict__ sell_vals, double* __restrict__ dinv) {
  const int nn = NN > 0 ? NN : nn_rt;
  const int lane = threadIdx.x & 31;
  const int64_t node = blockIdx.x * (int64_t)kWarps + (threadIdx.x >> 5);
  if (node >= n_nod) return;
  const int64_t nb = node_ptr[node];
  const int cnt = (int)(node_ptr[node + 1] - nb);
  const int64_t row0 = node * DPN;
  // SELL-32 position of entry k of row r: slice_ptr[r / 32] + (k / 2) * 64 + (r % 32) * 2 + (k % 2)   (sell.cuh)
  auto sell_base = [&](int64_t r) { return slice_ptr[r >> 5] + (r & 31) * 2; };
  auto sell_pad = [&](int len) {  // zero entries [len, width) of my rows
#pragma unroll
    for (int i = 0; i < DPN; ++i) {
      const int64_t r = row0 + i, s0 = slice_ptr[r >> 5];
      const int width = (int)((slice_ptr[(r >> 5) + 1] - s0) >> 5);
      for (int k = len + lane; k < width; k += 32) sell_vals[s0 + (r & 31) * 2 + (k >> 1) * 64 + (k & 1)] = 0.0;
    }
  };
  if (sell_vals && node == n_nod - 1) {  // the rows past the end of the last slice
    const int64_t n_rows = n_nod * DPN, r_end = (n_rows + 31) & ~(int64_t)31;
    for (int64_t r = n_rows; r < r_end; ++r) {
      const int64_t s0 = slice_ptr[r >> 5];
      const int width = (int)((slice_ptr[(r >> 5) + 1] - s0) >> 5);
      for (int k = lane; k < width; k += 32) sell_vals[s0 + (r & 31) * 2 + (k >> 1) * 64 + (k & 1)] = 0.0;
    }
  }
  if (cnt == 0) {  // unreferenced node: lone diagonal, value 0 unless constrained (base.py:419)
    if (lane < DPN) {
      const double d = (is_con && is_con[row0 + lane]) ? 1.0 : 0.0;
      if (vals) vals[indptr[row0 + lane]] = d;
      if (lift) lift[row0 + lane] = 0.0;
      if (dinv) dinv[row0 + lane] = 1.0 / d;
      if (sell_vals) sell_vals[sell_base(row0 + lane)] = d;
    }
    if (sell_vals) sell_pad(1);
    return;
  }
  const int nd = nn * DPN;
  const int64_t nd2 = (int64_t)nd * nd;
  bool row_con[DPN];
  int64_t rp[DPN];
  double lsum[DPN];
#pragma unroll
  for (int i = 0; i < DPN; ++i) {
    row_con[i] = is_con && is_con[row0 + i];
    rp[i] = vals ? indptr[row0 + i] : 0;
    lsum[i] = 0.0;
  }
  // The SELL-32 values leave through shared memory: a lane owns the 3 consecutive entries of a node block, but in the
  // slice two consecutive entries of ONE row are adjacent (16 B) and the node's rows follow each other, so the warp
  // regroups a round of 32 blocks into (entry pair, row) items and writes 48-byte runs instead of scattered 8-byte words
  // (8.8 -> 7.x ms at config B; written straight from the block registers the kernel is slower than the CSR order plus
  // the copy it is meant to replace).
  __shared__ __align__(16) double s_stage[kWarps][DPN][32 * DPN];
  __shared__ int64_t s_sbase[kWarps][DPN];
  const int warp = threadIdx.x >> 5;
  if (sell_vals && lane < DPN) s_sbase[warp][lane] = sell_base(row0 + lane);
  for (int p0 = 0; p0 < cnt; p0 += 32) {
    const int p = p0 + lane;
    const bool mine = p < cnt;
    double acc[DPN][DPN];
#pragma unroll
    for (int i = 0; i < DPN; ++i)
#pragma unroll
      for (int j = 0; j < DPN; ++j) acc[i][j] = 0.0;
    if (mine) {
      const int64_t sb = src_ptr[nb + p], se = src_ptr[nb + p + 1];
      for (int64_t s = sb; s < se; s += kInFlight) {
        const double* kb[kInFlight];
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) {
          const int cc = (s + u < se) ? src[s + u] : -1;  // e*nn*nn + a*nn + b
          if (cc < 0) {
            kb[u] = nullptr;
          } else {
            const int ea = cc / nn, bq = cc - ea * nn;
            const int e = ea / nn, aq = ea - e * nn;
            kb[u] = k + e * nd2 + (int64_t)(aq * DPN) * nd + bq * DPN;
          }
        }
        double v[kInFlight][DPN][DPN];
#pragma unroll
        for (int u = 0; u < kInFlight; ++u)
#pragma unroll
          for (int i = 0; i < DPN; ++i)
#pragma unroll
            for (int j = 0; j < DPN; ++j) v[u][i][j] = kb[u] ? kb[u][i * nd + j] : 0.0;
#pragma unroll
        for (int u = 0; u < kInFlight; ++u)  // + 0.0 for the missing ones: exact; slot order kept
#pragma unroll
          for (int i = 0; i < DPN; ++i)
#pragma unroll
            for (int j = 0; j < DPN; ++j) acc[i][j] += v[u][i][j];
      }
      const int64_t col0 = (int64_t)adj[nb + p] * DPN;
#pragma unroll
      for (int j = 0; j < DPN; ++j) {
        const int64_t col = col0 + j;
        const bool col_con = is_con && is_con[col];
        const double uc = (lift && col_con) ? ubc[col] : 0.0;
#pragma unroll
        for (int i = 0; i < DPN; ++i) {
          double a = acc[i][j];
          // Dirichlet lifting: what the prescribed values contribute to the free rows, K[row, con] u[con]
          if (lift && col_con && !row_con[i]) lsum[i] = fma(a, uc, lsum[i]);
          if (row_con[i] || col_con) a = (col == row0 + i) ? 1.0 : 0.0;
          if (vals) vals[rp[i] + (int64_t)p * DPN + j] = a;
          acc[i][j] = a;
        }
      }
      if (dinv && col0 == row0) {  // the lane of the diagonal block: 1 / diagonal after the masking
#pragma unroll
        for (int i = 0; i < DPN; ++i) dinv[row0 + i] = 1.0 / acc[i][i];
      }
    }
    if (sell_vals) {
#pragma unroll
      for (int i = 0; i < DPN; ++i)
#pragma unroll
        for (int j = 0; j < DPN; ++j) s_stage[warp][i][lane * DPN + j] = acc[i][j];  // 0.0 beyond the last block
      __syncwarp();
      const int ne = (cnt - p0 < 32 ? cnt - p0 : 32) * DPN;  // entries of this round, per row
      const int n_items = ((ne + 1) >> 1) * DPN;             // (entry pair, row); an odd tail pairs with a 0.0
      const int64_t pair0 = ((int64_t)p0 * DPN) >> 1;         // p0 * DPN is even (p0 = 0, 32, ...)
      for (int item = lane; item < n_items; item += 32) {
        const int sp = item / DPN, i = item - sp * DPN;
        const double2 v2 = *reinterpret_cast<const double2*>(&s_stage[warp][i][2 * sp]);
        *reinterpret_cast<double2*>(sell_vals + s_sbase[warp][i] + (pair0 + sp) * 64) = v2;
      }
      __syncwarp();
    }
  }
  if (sell_vals) sell_pad((cnt * DPN + 1) & ~1);
  if (lift) {
#pragma unroll
    for (int i = 0; i < DPN; ++i) {
      const double t = warp_sum(lsum[i]);
      if (lane == 0) lift[row0 + i] = t;
    }
  }
}

template <int DPN>
int launch_assemble(int64_t n_nod, int nn, const int64_t* node_ptr, const int32_t* adj, const int64_t* indptr,
                    const int64_t* src_ptr, const int32_t* src, const double* k, const uint8_t* is_con,
                    const double* ubc, double* vals, double* lift, const int64_t* slice_ptr, double* sell_vals,
                    double* dinv, cudaStream_t st) {
  const unsigned grid = (unsigned)((n_nod + kWarps - 1) / kWarps);
#define TFEM_ASM(NNC)                                                                                             \
  if (sell_vals || dinv)                                                                                          \
    k_assemble_solver<DPN, NNC><<<grid, kWarps * 32, 0, st>>>(n_nod, nn, node_ptr, adj, indptr, src_ptr, src, k,  \
                                                              is_con, ubc, vals, lift, slice_ptr, sell_vals, dinv); \
  else                                                                                                            \
    k_assemble<DPN, NNC><<<grid, kWarps * 32, 0, st>>>(n_nod, nn, node_ptr, adj, indptr, src_ptr, src, k, is_con, \
                                                       ubc, vals, lift)
  switch (nn) {
    case 3: TFEM_ASM(3); break;
    case 4: TFEM_ASM(4); break;
    case 6: TFEM_ASM(6); break;
    case 8: TFEM_ASM(8); break;
    case 10: TFEM_ASM(10); break;
    case 20: TFEM_ASM(20); break;
    default: TFEM_ASM(0); break;
  }
#undef TFEM_ASM
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

}  // namespace
}  // namespace tfem

using namespace tfem;

extern "C" int tfem_assemble_solve(int64_t n_nod, int nn, int dpn, const int64_t* node_ptr, const int32_t* adj,
                                   const int64_t* indptr, const int64_t* src_ptr, const int32_t* src,
                                   const double* k, const uint8_t* is_con, const double* ubc, double* vals,
                                   double* lift, const int64_t* slice_ptr, double* sell_vals, double* dinv,
                                   void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(node_ptr && adj && indptr && src_ptr && src && k, "assemble: null pointer");
  TFEM_REQUIRE(vals || sell_vals, "assemble: no output (CSR values, SELL values or both)");
  TFEM_REQUIRE(!sell_vals || slice_ptr, "assemble: SELL values need the slice offsets of tfem_sell_slice_ptr");
  TFEM_REQUIRE(n_nod > 0 && nn > 0, "assemble: bad sizes");
  TFEM_REQUIRE(!lift || (is_con && ubc), "assemble: the Dirichlet lifting needs is_con and the prescribed values");
  switch (dpn) {
    case 1:
      return launch_assemble<1>(n_nod, nn, node_ptr, adj, indptr, src_ptr, src, k, is_con, ubc, vals, lift, slice_ptr,
                                sell_vals, dinv, st);
    case 2:
      return launch_assemble<2>(n_nod, nn, node_ptr, adj, indptr, src_ptr, src, k, is_con, ubc, vals, lift, slice_ptr,
                                sell_vals, dinv, st);
    case 3:
      return launch_assemble<3>(n_nod, nn, node_ptr, adj, indptr, src_ptr, src, k, is_con, ubc, vals, lift, slice_ptr,
                                sell_vals, dinv, st);
  }
  set_last_error("invalid argument", "assemble: dofs per node must be 1, 2 or 3");
  return TFEM_ERR_INVALID;
}

extern "C" int tfem_assemble_bc(int64_t n_nod, int nn, int dpn, const int64_t* node_ptr, const int32_t* adj,
                                const int64_t* indptr, const int64_t* src_ptr, const int32_t* src,
                                const double* k, const uint8_t* is_con, const double* ubc, double* vals,
                                double* lift, void* stream_) {
  TFEM_REQUIRE(vals, "assemble: null pointer");
  return tfem_assemble_solve(n_nod, nn, dpn, node_ptr, adj, indptr, src_ptr, src, k, is_con, ubc, vals, lift, nullptr,
                             nullptr, nullptr, stream_);
}

extern "C" int tfem_assemble(int64_t n_nod, int nn, int dpn, const int64_t* node_ptr, const int32_t* adj,
                             const int64_t* indptr, const int64_t* src_ptr, const int32_t* src,
                             const double* k, const uint8_t* is_con, double* vals, void* stream_) {
  return tfem_assemble_bc(n_nod, nn, dpn, node_ptr, adj, indptr, src_ptr, src, k, is_con, nullptr, vals, nullptr,
                          stream_);
}
