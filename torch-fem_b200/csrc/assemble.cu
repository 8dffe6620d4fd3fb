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

template <int DPN>
__global__ void __launch_bounds__(kWarps * 32)
    k_assemble(int64_t n_nod, int nn, const int64_t* __restrict__ node_ptr,
               const int32_t* __restrict__ adj, const int64_t* __restrict__ indptr,
               const int64_t* __restrict__ src_ptr, const int32_t* __restrict__ src,
               const double* __restrict__ k, const uint8_t* __restrict__ is_con,
               double* __restrict__ vals) {
  const int lane = threadIdx.x & 31;
  const int64_t node = blockIdx.x * (int64_t)kWarps + (threadIdx.x >> 5);
  if (node >= n_nod) return;
  const int64_t nb = node_ptr[node];
  const int cnt = (int)(node_ptr[node + 1] - nb);
  const int64_t row0 = node * DPN;
  if (cnt == 0) {  // unreferenced node: lone diagonal, value 0 unless constrained (base.py:419)
    if (lane < DPN) vals[indptr[row0 + lane]] = (is_con && is_con[row0 + lane]) ? 1.0 : 0.0;
    return;
  }
  const int nd = nn * DPN;
  const int64_t nd2 = (int64_t)nd * nd;
  const int rowlen = cnt * DPN;
#pragma unroll
  for (int i = 0; i < DPN; ++i) {
    const int64_t row = row0 + i;
    const int64_t rp = indptr[row];
    const bool row_con = is_con && is_con[row];
    for (int t = lane; t < rowlen; t += 32) {
      const int p = t / DPN, j = t - p * DPN;
      const int64_t sb = src_ptr[nb + p], se = src_ptr[nb + p + 1];
      double acc = 0.0;
      for (int64_t s = sb; s < se; ++s) {
        const int c = src[s];  // e*nn*nn + a*nn + b
        const int b = c % nn;
        const int ea = c / nn;  // e*nn + a
        const int a = ea % nn;
        const int64_t e = ea / nn;
        acc += k[e * nd2 + (int64_t)(a * DPN + i) * nd + b * DPN + j];
      }
      const int64_t col = (int64_t)adj[nb + p] * DPN + j;
      if (row_con || (is_con && is_con[col])) acc = (col == row) ? 1.0 : 0.0;
      vals[rp + t] = acc;
    }
  }
}

}  // namespace
}  // namespace tfem

using namespace tfem;

extern "C" int tfem_assemble(int64_t n_nod, int nn, int dpn, const int64_t* node_ptr, const int32_t* adj,
                             const int64_t* indptr, const int64_t* src_ptr, const int32_t* src,
                             const double* k, const uint8_t* is_con, double* vals, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  TFEM_REQUIRE(node_ptr && adj && indptr && src_ptr && src && k && vals, "assemble: null pointer");
  TFEM_REQUIRE(n_nod > 0 && nn > 0, "assemble: bad sizes");
  const unsigned grid = (unsigned)((n_nod + kWarps - 1) / kWarps);
  switch (dpn) {
    case 1:
      k_assemble<1><<<grid, kWarps * 32, 0, st>>>(n_nod, nn, node_ptr, adj, indptr, src_ptr, src, k, is_con, vals);
      break;
    case 2:
      k_assemble<2><<<grid, kWarps * 32, 0, st>>>(n_nod, nn, node_ptr, adj, indptr, src_ptr, src, k, is_con, vals);
      break;
    case 3:
      k_assemble<3><<<grid, kWarps * 32, 0, st>>>(n_nod, nn, node_ptr, adj, indptr, src_ptr, src, k, is_con, vals);
      break;
    default:
      set_last_error("invalid argument", "assemble: dofs per node must be 1, 2 or 3");
      return TFEM_ERR_INVALID;
  }
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}
