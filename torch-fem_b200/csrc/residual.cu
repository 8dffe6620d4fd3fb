// K9/K10 — the two geometry contractions of the residual evaluation, for all Gauss points in one launch:
//   tfem_elem_grad   H[q,e,i,J] = s_q,e * sum_n u_e[e,n,i] B_q[e,J,n]          (reference base.py:1052 `du @ B[i]^T`,
//                                                                               heat: base.py:1241-1243)
//   tfem_elem_force  f_e[e,n,i] = sum_q s_q,e * sum_J B_q[e,J,n] P[q,e,i,J]     (base.py:1082-1083 + compute_f,
//                                                                               solid.py:56-58 / planar.py:90-92)
// with B_q = J_q^-1 b_q recomputed from the node coordinates (closed-form inverse) instead of read from a cached
// [n_int, n_elem, dim, nn] tensor, and s = 1 (unweighted) or w_q detJ_q (weighted). Each is the transpose of the
// other, so the pair also serves as each other's autograd backward (torch-fem_b200/residual.py): the material
// update between them stays in torch (the adjoint differentiates through it, reference sparse.py:689-705).
// The reference runs ~6 batched-matmul / einsum launches per Gauss point for these two lines; on B200 those
// tiny-matrix bmm kernels took 750 ms of a 1.8 s `Solid.solve` at config B.
//
// Mapping: one thread per element, loop over Gauss points; node coordinates (and u_e) of the CTA's 64 elements
// are staged in shared memory with an odd per-element stride (conflict-free); H / P rows of neighbouring lanes
// are contiguous in memory. Bytes: 72 B per (element, Gauss point) + ~600 B per element; FLOPs negligible.
#include "common.cuh"

namespace tfem {
namespace {

template <int DIM, int NN, int NINT>
struct RTables {
  double bref[NINT * DIM * NN];
  double w[NINT];
};

template <int DIM>
__device__ __forceinline__ double inv_det_r(const double (&J)[DIM][DIM], double (&inv)[DIM][DIM]);

template <>
__device__ __forceinline__ double inv_det_r<2>(const double (&J)[2][2], double (&inv)[2][2]) {
  const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
  const double id = 1.0 / det;
  inv[0][0] = J[1][1] * id;
  inv[0][1] = -J[0][1] * id;
  inv[1][0] = -J[1][0] * id;
  inv[1][1] = J[0][0] * id;
  return det;
}

template <>
__device__ __forceinline__ double inv_det_r<3>(const double (&J)[3][3], double (&inv)[3][3]) {
  const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
  const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
  const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
  const double id = 1.0 / det;
  inv[0][0] = c00 * id;
  inv[1][0] = c01 * id;
  inv[2][0] = c02 * id;
  inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
  inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
  inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
  inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
  inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
  inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
  return det;
}

constexpr int kEPC = 64;  // elements (= threads) per CTA

// J_q = b_q X_e (reference base.py:306-309), its inverse and determinant
template <int DIM, int NN, int NINT>
__device__ __forceinline__ double jacobian(const RTables<DIM, NN, NINT>& tab, int q, const double* X,
                                           double (&inv)[DIM][DIM]) {
  double J[DIM][DIM];
#pragma unroll
  for (int i = 0; i < DIM; ++i)
#pragma unroll
    for (int j = 0; j < DIM; ++j) J[i][j] = 0.0;
  for (int n = 0; n < NN; ++n) {
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
      const double b = tab.bref[(q * DIM + i) * NN + n];
#pragma unroll
      for (int j = 0; j < DIM; ++j) J[i][j] = fma(b, X[n * DIM + j], J[i][j]);
    }
  }
  return inv_det_r<DIM>(J, inv);
}

template <int DIM, int NN>
__device__ __forceinline__ void stage_coords(double* sX, int strideX, const double* __restrict__ nodes,
                                             const int64_t* __restrict__ elements, int64_t e0, int64_t n_elem) {
  for (int t = threadIdx.x; t < kEPC * NN; t += blockDim.x) {
    const int le = t / NN, n = t - le * NN;
    const int64_t e = e0 + le;
    if (e < n_elem) {
      const int64_t node = elements[e * NN + n];
#pragma unroll
      for (int j = 0; j < DIM; ++j) sX[le * strideX + n * DIM + j] = nodes[node * DIM + j];
    }
  }
}

template <int DIM, int NN, int NINT, int DPN>
__global__ void __launch_bounds__(kEPC)
    k_elem_grad(const __grid_constant__ RTables<DIM, NN, NINT> tab, const double* __restrict__ nodes,
                const int64_t* __restrict__ elements, int64_t n_elem, const double* __restrict__ u_e,
                const double* __restrict__ scale, int weighted, double* __restrict__ H,
                int32_t* __restrict__ neg_jac) {
  constexpr int SX = (NN * DIM) | 1, SU = (NN * DPN) | 1;
  extern __shared__ double smem[];
  double* sX = smem;
  double* sU = smem + kEPC * SX;
  const int64_t e0 = (int64_t)blockIdx.x * kEPC;
  stage_coords<DIM, NN>(sX, SX, nodes, elements, e0, n_elem);
  for (int t = threadIdx.x; t < kEPC * NN * DPN; t += blockDim.x) {
    const int le = t / (NN * DPN), r = t - le * (NN * DPN);
    if (e0 + le < n_elem) sU[le * SU + r] = u_e[(e0 + le) * (NN * DPN) + r];
  }
  __syncthreads();
  const int64_t e = e0 + threadIdx.x;
  if (e >= n_elem) return;
  const double* X = sX + threadIdx.x * SX;
  const double* U = sU + threadIdx.x * SU;
  const double se = (weighted && scale) ? scale[e] : 1.0;
#pragma unroll 1
  for (int q = 0; q < NINT; ++q) {
    double inv[DIM][DIM];
    const double det = jacobian<DIM, NN, NINT>(tab, q, X, inv);
    if (!(det > 0.0)) atomicOr(neg_jac, 1);
    // G[i][j] = sum_n u[n,i] b_q[j,n]  (gradient in reference coordinates), H[i][J] = sum_j inv[J][j] G[i][j]
    double G[DPN][DIM];
#pragma unroll
    for (int i = 0; i < DPN; ++i)
#pragma unroll
      for (int j = 0; j < DIM; ++j) G[i][j] = 0.0;
    for (int n = 0; n < NN; ++n) {
#pragma unroll
      for (int j = 0; j < DIM; ++j) {
        const double b = tab.bref[(q * DIM + j) * NN + n];
#pragma unroll
        for (int i = 0; i < DPN; ++i) G[i][j] = fma(U[n * DPN + i], b, G[i][j]);
      }
    }
    const double s = weighted ? tab.w[q] * det * se : 1.0;
    double* out = H + ((int64_t)q * n_elem + e) * (DPN * DIM);
#pragma unroll
    for (int i = 0; i < DPN; ++i)
#pragma unroll
      for (int Jx = 0; Jx < DIM; ++Jx) {
        double h = 0.0;
#pragma unroll
        for (int j = 0; j < DIM; ++j) h = fma(inv[Jx][j], G[i][j], h);
        out[i * DIM + Jx] = s * h;
      }
  }
}

template <int DIM, int NN, int NINT, int DPN>
__global__ void __launch_bounds__(kEPC)
    k_elem_force(const __grid_constant__ RTables<DIM, NN, NINT> tab, const double* __restrict__ nodes,
                 const int64_t* __restrict__ elements, int64_t n_elem, const double* __restrict__ P,
                 const double* __restrict__ scale, int weighted, double* __restrict__ f_e,
                 int32_t* __restrict__ neg_jac) {
  constexpr int SX = (NN * DIM) | 1;
  extern __shared__ double smem[];
  double* sX = smem;
  const int64_t e0 = (int64_t)blockIdx.x * kEPC;
  stage_coords<DIM, NN>(sX, SX, nodes, elements, e0, n_elem);
  __syncthreads();
  const int64_t e = e0 + threadIdx.x;
  if (e >= n_elem) return;
  const double* X = sX + threadIdx.x * SX;
  const double se = (weighted && scale) ? scale[e] : 1.0;
  double acc[NN][DPN];
#pragma unroll
  for (int n = 0; n < NN; ++n)
#pragma unroll
    for (int i = 0; i < DPN; ++i) acc[n][i] = 0.0;
#pragma unroll 1
  for (int q = 0; q < NINT; ++q) {
    double inv[DIM][DIM];
    const double det = jacobian<DIM, NN, NINT>(tab, q, X, inv);
    if (!(det > 0.0)) atomicOr(neg_jac, 1);
    const double s = weighted ? tab.w[q] * det * se : 1.0;
    const double* Pq = P + ((int64_t)q * n_elem + e) * (DPN * DIM);
    // T[i][j] = s * sum_J P[i][J] inv[J][j] ; f[n][i] += sum_j b_q[j,n] T[i][j]
    double T[DPN][DIM];
#pragma unroll
    for (int i = 0; i < DPN; ++i) {
      double p[DIM];
#pragma unroll
      for (int Jx = 0; Jx < DIM; ++Jx) p[Jx] = Pq[i * DIM + Jx];
#pragma unroll
      for (int j = 0; j < DIM; ++j) {
        double t = 0.0;
#pragma unroll
        for (int Jx = 0; Jx < DIM; ++Jx) t = fma(p[Jx], inv[Jx][j], t);
        T[i][j] = s * t;
      }
    }
#pragma unroll
    for (int n = 0; n < NN; ++n) {
#pragma unroll
      for (int j = 0; j < DIM; ++j) {
        const double b = tab.bref[(q * DIM + j) * NN + n];
#pragma unroll
        for (int i = 0; i < DPN; ++i) acc[n][i] = fma(b, T[i][j], acc[n][i]);
      }
    }
  }
  double* out = f_e + e * (NN * DPN);
#pragma unroll
  for (int n = 0; n < NN; ++n)
#pragma unroll
    for (int i = 0; i < DPN; ++i) out[n * DPN + i] = acc[n][i];
}

template <int DIM, int NN, int NINT, int DPN>
int launch_pair(bool grad, const double* bref, const double* w, const double* nodes, const int64_t* elements,
                int64_t n_elem, const double* in, const double* scale, int weighted, double* out, int32_t* neg_jac,
                cudaStream_t st) {
  RTables<DIM, NN, NINT> tab;
  for (int i = 0; i < NINT * DIM * NN; ++i) tab.bref[i] = bref[i];
  for (int i = 0; i < NINT; ++i) tab.w[i] = w[i];
  const unsigned grid = (unsigned)((n_elem + kEPC - 1) / kEPC);
  constexpr int SX = (NN * DIM) | 1, SU = (NN * DPN) | 1;
  if (grad) {
    const size_t bytes = (size_t)kEPC * (SX + SU) * sizeof(double);
    if (bytes > 48 * 1024)
      TFEM_CUDA(cudaFuncSetAttribute(k_elem_grad<DIM, NN, NINT, DPN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    k_elem_grad<DIM, NN, NINT, DPN><<<grid, kEPC, bytes, st>>>(tab, nodes, elements, n_elem, in, scale, weighted, out, neg_jac);
  } else {
    const size_t bytes = (size_t)kEPC * SX * sizeof(double);
    if (bytes > 48 * 1024)
      TFEM_CUDA(cudaFuncSetAttribute(k_elem_force<DIM, NN, NINT, DPN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    k_elem_force<DIM, NN, NINT, DPN><<<grid, kEPC, bytes, st>>>(tab, nodes, elements, n_elem, in, scale, weighted, out, neg_jac);
  }
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

int dispatch_pair(bool grad, int dim, int nn, int n_int, int dpn, const double* bref, const double* w,
                  const double* nodes, const int64_t* elements, int64_t n_elem, const double* in,
                  const double* scale, int weighted, double* out, int32_t* neg_jac, cudaStream_t st) {
#define TFEM_CASE(D, N, Q)                                                                                    \
  if (dim == D && nn == N && n_int == Q) {                                                                    \
    if (dpn == D)                                                                                             \
      return launch_pair<D, N, Q, D>(grad, bref, w, nodes, elements, n_elem, in, scale, weighted, out, neg_jac, st); \
    if (dpn == 1)                                                                                             \
      return launch_pair<D, N, Q, 1>(grad, bref, w, nodes, elements, n_elem, in, scale, weighted, out, neg_jac, st); \
  }
  TFEM_CASE(3, 8, 8)
  TFEM_CASE(3, 20, 8)
  TFEM_CASE(3, 4, 1)
  TFEM_CASE(3, 10, 4)
  TFEM_CASE(2, 4, 4)
  TFEM_CASE(2, 8, 4)
  TFEM_CASE(2, 3, 1)
  TFEM_CASE(2, 6, 3)
#undef TFEM_CASE
  set_last_error("invalid argument", "unsupported (dim, nodes per element, integration points, dofs per node)");
  return TFEM_ERR_INVALID;
}

// ------------------------------------------------------------------------------------------ K17 tangent contraction
// out[q,e,i] = sum_k C[e,i,k] E[q,e,k]   (TRANSPOSE: C[e,k,i]) with M = d*d flattened index pairs — the elastic stress
// update sigma = C : eps at every Gauss point (reference elasticity.py:119-127 `einsum("...ijkl,...kl->...ij")`) and,
// transposed, its backward with respect to the strain. One thread per (element, row) keeps its row of C in registers
// and walks the Gauss points: C is read once, E / out once each (bytes: 8 M^2 + 16 Q M per element).
template <int M, bool TRANSPOSE>
__global__ void __launch_bounds__(256)
    k_ddot(int64_t n_q, int64_t n_elem, const double* __restrict__ C, const double* __restrict__ E,
           double* __restrict__ out) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t e = t / M;
  const int i = (int)(t - e * M);
  if (e >= n_elem) return;
  double c[M];
#pragma unroll
  for (int k = 0; k < M; ++k) c[k] = TRANSPOSE ? C[(e * M + k) * M + i] : C[(e * M + i) * M + k];
  for (int64_t q = 0; q < n_q; ++q) {
    const double* h = E + (q * n_elem + e) * M;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) s = fma(c[k], h[k], s);
    out[(q * n_elem + e) * M + i] = s;
  }
}

// gC[e,i,k] = sum_q G[q,e,i] E[q,e,k]  — the backward of k_ddot with respect to the tangent
template <int M>
__global__ void __launch_bounds__(256)
    k_ddot_outer(int64_t n_q, int64_t n_elem, const double* __restrict__ G, const double* __restrict__ E,
                 double* __restrict__ gC) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t e = t / M;
  const int i = (int)(t - e * M);
  if (e >= n_elem) return;
  double acc[M];
#pragma unroll
  for (int k = 0; k < M; ++k) acc[k] = 0.0;
  for (int64_t q = 0; q < n_q; ++q) {
    const double g = G[(q * n_elem + e) * M + i];
    const double* h = E + (q * n_elem + e) * M;
#pragma unroll
    for (int k = 0; k < M; ++k) acc[k] = fma(g, h[k], acc[k]);
  }
#pragma unroll
  for (int k = 0; k < M; ++k) gC[(e * M + i) * M + k] = acc[k];
}

// a9 — deterministic `assemble_rhs` (src/torchfem/base.py:428-445: F.index_add_(0, idx.ravel(), f.ravel()), atomics on
// CUDA): a gather over the node -> (element, local node) incidence lists of the pattern build. One thread per global
// DOF sums its contributions in ascending slot order — the order the reference's CPU index_add_ meets them — so the
// result is bitwise reproducible. Neighbouring threads read neighbouring doubles of the same element rows.
__global__ void k_assemble_rhs(int64_t n_dofs, int dpn, const int32_t* __restrict__ inc_ptr,
                               const int32_t* __restrict__ inc_list, const double* __restrict__ f,
                               double* __restrict__ F) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_dofs) return;
  const int64_t node = r / dpn;
  const int i = (int)(r - node * dpn);
  double s = 0.0;
  for (int k = inc_ptr[node]; k < inc_ptr[node + 1]; ++k) s += f[(int64_t)inc_list[k] * dpn + i];
  F[r] = s;
}

}  // namespace
}  // namespace tfem

using namespace tfem;

extern "C" int tfem_assemble_rhs(int64_t n_nod, int dpn, const int32_t* inc_ptr, const int32_t* inc_list,
                                 const double* f_e, double* F, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(inc_ptr && inc_list && f_e && F && dpn >= 1, "assemble_rhs: bad arguments");
  if (n_nod <= 0) return TFEM_OK;
  const int64_t n = n_nod * dpn;
  k_assemble_rhs<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, dpn, inc_ptr, inc_list, f_e, F);
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

extern "C" int tfem_elem_grad(int dim, int nn, int n_int, int dpn, const double* bref_host, const double* w_host,
                              const double* nodes, const int64_t* elements, int64_t n_elem, const double* u_e,
                              const double* scale, int weighted, double* H, int32_t* neg_jac, void* stream_) {
  TFEM_REQUIRE(bref_host && w_host && nodes && elements && u_e && H && neg_jac, "elem_grad: null pointer");
  if (n_elem <= 0) return TFEM_OK;
  return dispatch_pair(true, dim, nn, n_int, dpn, bref_host, w_host, nodes, elements, n_elem, u_e, scale, weighted, H,
                       neg_jac, (cudaStream_t)stream_);
}

extern "C" int tfem_elem_force(int dim, int nn, int n_int, int dpn, const double* bref_host, const double* w_host,
                               const double* nodes, const int64_t* elements, int64_t n_elem, const double* P,
                               const double* scale, int weighted, double* f_e, int32_t* neg_jac, void* stream_) {
  TFEM_REQUIRE(bref_host && w_host && nodes && elements && P && f_e && neg_jac, "elem_force: null pointer");
  if (n_elem <= 0) return TFEM_OK;
  return dispatch_pair(false, dim, nn, n_int, dpn, bref_host, w_host, nodes, elements, n_elem, P, scale, weighted, f_e,
                       neg_jac, (cudaStream_t)stream_);
}

extern "C" int tfem_ddot(int m, int64_t n_q, int64_t n_elem, const double* C, const double* E, int transpose,
                         double* out, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(C && E && out && n_q >= 0, "ddot: null pointer");
  TFEM_REQUIRE(m == 1 || m == 4 || m == 9, "ddot: m must be 1, 4 or 9 (d*d)");
  if (n_elem <= 0 || n_q == 0) return TFEM_OK;
  const unsigned grid = (unsigned)((n_elem * m + 255) / 256);
#define TFEM_DDOT(M)                                                                  \
  if (m == M) {                                                                       \
    if (transpose) k_ddot<M, true><<<grid, 256, 0, st>>>(n_q, n_elem, C, E, out);     \
    else k_ddot<M, false><<<grid, 256, 0, st>>>(n_q, n_elem, C, E, out);              \
  }
  TFEM_DDOT(9)
  TFEM_DDOT(4)
  TFEM_DDOT(1)
#undef TFEM_DDOT
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

extern "C" int tfem_ddot_outer(int m, int64_t n_q, int64_t n_elem, const double* G, const double* E, double* gC,
                               void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(G && E && gC && n_q >= 0, "ddot_outer: null pointer");
  TFEM_REQUIRE(m == 1 || m == 4 || m == 9, "ddot_outer: m must be 1, 4 or 9 (d*d)");
  if (n_elem <= 0) return TFEM_OK;
  const unsigned grid = (unsigned)((n_elem * m + 255) / 256);
  if (m == 9) k_ddot_outer<9><<<grid, 256, 0, st>>>(n_q, n_elem, G, E, gC);
  else if (m == 4) k_ddot_outer<4><<<grid, 256, 0, st>>>(n_q, n_elem, G, E, gC);
  else k_ddot_outer<1><<<grid, 256, 0, st>>>(n_q, n_elem, G, E, gC);
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}
