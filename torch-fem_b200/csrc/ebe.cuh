// K8 — matrix-free (element-by-element) operator on stored element matrices:
//     y = sum_e P_e^T k_e P_e x      with the Dirichlet masking of the assembly applied on the fly
// (rows / columns of constrained DOFs dropped, unit diagonal there — what base.py:414-419 does to the
// assembled matrix). The reference has no matrix-free path; BASELINE.json's north_star lists it as the optional
// operator of the Krylov solve: no sparsity pattern values, no assembly, no format conversion — useful when the
// tangent changes at every Newton iteration and only a few Krylov iterations are spent per matrix.
//
// Deterministic gather (no scatter, no atomics): one warp per node walks the node's incident (element, local
// node) pairs in ascending slot order — the incidence lists the pattern build produces anyway — lane c owns
// column c of the element matrix: the warp reads dpn rows of k_e (contiguous nd doubles each) per pair and the
// element's x values once. Bytes per product: all of k (8 nd^2 per element; 15.6 GB at config B, against 7.0 GB
// for the block-SELL matrix) — so per iteration it is ~2.2x slower than the assembled SpMV and pays off only
// below ~10 iterations per matrix.
#pragma once
#include "sell.cuh"

namespace tfem {
namespace {

struct Ebe {
  int64_t n_nod;
  int nn, dpn;
  const int32_t* inc_ptr;    // [n_nod+1]
  const int32_t* inc_list;   // slots e*nn + a, ascending per node
  const int64_t* elements;   // [n_elem*nn]
  const double* k;           // [n_elem, nd, nd]
  const uint8_t* is_con;     // [n_dofs] or nullptr
};

constexpr int kEbeWarps = 8;

template <int DPN, bool DOT>
__global__ void __launch_bounds__(kEbeWarps * 32)
    k_ebe_spmv(Ebe A, const double* __restrict__ x, double* __restrict__ y, const double* sc, double* partials,
               unsigned int* ticket, double* out_scalar) {
  __shared__ double s_red[kEbeWarps];
  if (DOT && sc[SC_DONE] != 0.0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nn = A.nn, nd = nn * DPN;
  const int64_t nd2 = (int64_t)nd * nd;
  double dot = 0.0;
  for (int64_t node = (int64_t)blockIdx.x * kEbeWarps + warp; node < A.n_nod; node += (int64_t)gridDim.x * kEbeWarps) {
    const int b = A.inc_ptr[node], e_ = A.inc_ptr[node + 1];
    double acc[DPN];
#pragma unroll
    for (int i = 0; i < DPN; ++i) acc[i] = 0.0;
    for (int s = b; s < e_; ++s) {
      const int slot = A.inc_list[s];
      const int e = slot / nn, a = slot - e * nn;
      const double* ke = A.k + e * nd2 + (int64_t)(a * DPN) * nd;
      for (int c = lane; c < nd; c += 32) {
        const int bn = c / DPN;
        const int64_t col = A.elements[(int64_t)e * nn + bn] * DPN + (c - bn * DPN);
        const double xc = (A.is_con && A.is_con[col]) ? 0.0 : __ldg(x + col);
#pragma unroll
        for (int i = 0; i < DPN; ++i) acc[i] = fma(ldg_stream_double(ke + (int64_t)i * nd + c), xc, acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < DPN; ++i) acc[i] = warp_sum(acc[i]);
    if (lane < DPN) {
      const int64_t row = node * DPN + lane;
      double v = acc[0];
#pragma unroll
      for (int i = 1; i < DPN; ++i) v = (lane == i) ? acc[i] : v;
      const double xr = __ldg(x + row);
      if (A.is_con && A.is_con[row]) v = xr;          // unit diagonal on constrained rows
      y[row] = v;
      if (DOT) dot = fma(v, xr, dot);
    }
  }
  if (DOT) {
    const double bsum = block_sum<kEbeWarps * 32>(dot, s_red);
    double mine[1] = {bsum}, tot[1];
    if (publish_and_reduce<1>(mine, partials, ticket, tot) && threadIdx.x == 0) *out_scalar = tot[0];
  }
}

// diag[row] = sum over incident (e, a) of k_e[(a,i),(a,i)]  (1 on constrained rows)
template <int DPN>
__global__ void k_ebe_diag(Ebe A, double* __restrict__ diag) {
  const int64_t node = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (node >= A.n_nod) return;
  const int nn = A.nn, nd = nn * DPN;
  const int64_t nd2 = (int64_t)nd * nd;
  double d[DPN];
#pragma unroll
  for (int i = 0; i < DPN; ++i) d[i] = 0.0;
  for (int s = A.inc_ptr[node]; s < A.inc_ptr[node + 1]; ++s) {
    const int slot = A.inc_list[s];
    const int e = slot / nn, a = slot - e * nn;
#pragma unroll
    for (int i = 0; i < DPN; ++i) d[i] += A.k[e * nd2 + (int64_t)(a * DPN + i) * nd + a * DPN + i];
  }
#pragma unroll
  for (int i = 0; i < DPN; ++i) {
    const int64_t row = node * DPN + i;
    diag[row] = (A.is_con && A.is_con[row]) ? 1.0 : d[i];
  }
}

template <int DPN, bool DOT>
int launch_ebe_t(const Ebe& A, const double* x, double* y, const double* sc, double* partials,
                 unsigned int* ticket, double* out_scalar, cudaStream_t st) {
  const int g = cached_resident_ctas(k_ebe_spmv<DPN, DOT>, kEbeWarps * 32);  // per instantiation and device
  const int64_t want = (A.n_nod + kEbeWarps - 1) / kEbeWarps;
  k_ebe_spmv<DPN, DOT><<<(int)(want < g ? want : g), kEbeWarps * 32, 0, st>>>(A, x, y, sc, partials, ticket, out_scalar);
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

template <bool DOT>
int launch_ebe(const Ebe& A, const double* x, double* y, const double* sc, double* partials,
               unsigned int* ticket, double* out_scalar, cudaStream_t st) {
  if (A.dpn == 3) return launch_ebe_t<3, DOT>(A, x, y, sc, partials, ticket, out_scalar, st);
  if (A.dpn == 2) return launch_ebe_t<2, DOT>(A, x, y, sc, partials, ticket, out_scalar, st);
  return launch_ebe_t<1, DOT>(A, x, y, sc, partials, ticket, out_scalar, st);
}

inline int check_ebe(const tfem_ebe_t* a) {
  TFEM_REQUIRE(a && a->inc_ptr && a->inc_list && a->elements && a->k && a->n_nod > 0 && a->nn > 0,
               "element operator: null pointer or empty");
  TFEM_REQUIRE(a->dpn >= 1 && a->dpn <= 3, "element operator: dofs per node must be 1, 2 or 3");
  return TFEM_OK;
}

inline Ebe make_ebe(const tfem_ebe_t* a) {
  Ebe A;
  A.n_nod = a->n_nod;
  A.nn = a->nn;
  A.dpn = a->dpn;
  A.inc_ptr = a->inc_ptr;
  A.inc_list = a->inc_list;
  A.elements = a->elements;
  A.k = a->k;
  A.is_con = a->is_con;
  return A;
}

}  // namespace
}  // namespace tfem
