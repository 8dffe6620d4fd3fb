// K1 — element matrices k_e = sum_q w_q detJ_q s_e B_q^T C B_q in FP64.
//
// Replaces, in one launch, what the reference does with ~10 torch ops per Gauss point
// (src/torchfem/base.py:293-314 eval_shape_functions: einsum + batched LU det/inv; base.py:1086-1090 /
// :1272-1278: 3-operand einsum + compute_k + accumulate) and its [n_int,n_elem,d,nn] temporaries.
//
// Mapping: one thread per (element, RPT row nodes a, DOF i, chunk of CB = 8 column nodes); it owns RPT x 8 dpn entries
// of RPT rows of k_e and contracts in two stages per Gauss point — U[r][k,L] = sum_J (w detJ B[J,a_r]) C[i,J,k,L]
// once, then k[(a_r,i),(b,k)] += sum_L U[r][k,L] B[L,b] for its columns — 19 kFMA per Hexa1 element instead of the
// 52 kFMA of the direct triple product. Register tiling RPT = 2 rows that share the DOF i (mechanics, even node
// counts) halves the shared-memory operand traffic: the column gradients B[:,b] and the tangent row C[i,:,:,:] are
// loaded once for both rows (3.5 FMA per LDS instead of 1.8; round 1's kernel was bound by the LDS pipe at 56 % of
// its wavefront rate with the FP64 pipe at 30 %), and an elastic tangent (one tensor for all Gauss points) is held in
// registers over the whole Gauss loop. Per element the node coordinates, the physical gradients B_q = J_q^-1 b_q of all Gauss points,
// w_q detJ_q and the material tangent are staged in shared memory (the reference-space table b_q and the
// weights come in as a __grid_constant__ kernel parameter, so the call is re-entrant across streams). The
// finished element matrix is staged in shared memory and written with fully coalesced stores: HBM traffic is
// the 8*(nn*dpn)^2 B/element output plus ~1 kB/element of inputs.
#include "common.cuh"

namespace tfem {
namespace {

template <int DIM, int NN, int NINT>
struct Tables {
  double bref[NINT * DIM * NN];  // [q][i][N] = d N_N / d xi_i at Gauss point q
  double w[NINT];
};

template <int DIM>
__device__ __forceinline__ double inv_det(const double (&J)[DIM][DIM], double (&inv)[DIM][DIM]);

template <>
__device__ __forceinline__ double inv_det<2>(const double (&J)[2][2], double (&inv)[2][2]) {
  const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
  const double id = 1.0 / det;
  inv[0][0] = J[1][1] * id;
  inv[0][1] = -J[0][1] * id;
  inv[1][0] = -J[1][0] * id;
  inv[1][1] = J[0][0] * id;
  return det;
}

template <>
__device__ __forceinline__ double inv_det<3>(const double (&J)[3][3], double (&inv)[3][3]) {
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

constexpr int kCB = 8;  // column nodes per thread
#ifndef TFEM_K1_MINB
#define TFEM_K1_MINB 1     // minimum resident CTAs per SM the register allocation must allow (tuning: tools/time_k1.py)
#endif
#ifndef TFEM_K1_RPT
#define TFEM_K1_RPT 1      // row nodes per thread (mechanics, even node counts). 2 halves the LDS operand traffic but
                           // runs SLOWER on B200 (17.2-17.7 vs 15.1 ms at config B: 220 registers -> 8 warps per SM;
                           // profiles/r2_e_k1_variants.txt, r2_f_k1_variants.txt)
#endif

template <int KIND, int NN>
struct K1Rows { static constexpr int v = (KIND == TFEM_KIND_MECH && NN % TFEM_K1_RPT == 0) ? TFEM_K1_RPT : 1; };

template <int KIND, int DIM, int NN, int NINT, int EPC, bool PERGP>
__global__ void __launch_bounds__(((EPC * (NN / K1Rows<KIND, NN>::v) * ((KIND == TFEM_KIND_MECH) ? DIM : 1) * ((NN + kCB - 1) / kCB) + 31) / 32) * 32,
                                  TFEM_K1_MINB)
    k_integrate(const __grid_constant__ Tables<DIM, NN, NINT> tab, const double* __restrict__ nodes,
                const int64_t* __restrict__ elements, int64_t n_elem,
                const double* __restrict__ tangent, const double* __restrict__ scale,
                double* __restrict__ k_out, int32_t* __restrict__ neg_jac) {
  constexpr int DPN = (KIND == TFEM_KIND_MECH) ? DIM : 1;
  constexpr int ND = NN * DPN;
  constexpr int TS = (KIND == TFEM_KIND_MECH) ? DIM * DIM * DIM * DIM : DIM * DIM;
  constexpr int NQC = PERGP ? NINT : 1;
  constexpr int NCH = (NN + kCB - 1) / kCB;
  constexpr int RPT = K1Rows<KIND, NN>::v;
  constexpr int TPE = (ND / RPT) * NCH;  // threads per element
  // per-element shared layout (doubles)
  constexpr int OFF_X = 0;
  constexpr int OFF_B = OFF_X + NN * DIM;
  constexpr int OFF_WD = OFF_B + NINT * DIM * NN;
  constexpr int OFF_C = OFF_WD + NINT;
  constexpr int OFF_K = OFF_C + NQC * TS;
  constexpr int OFF_J = OFF_K + ND * ND;  // Jacobians and their inverses at the Gauss points
  constexpr int PER_ELEM = OFF_J + 2 * NINT * DIM * DIM;

  extern __shared__ double smem[];
  const int tid = threadIdx.x;
  const int el = tid / TPE;
  const int pr = tid - el * TPE;
  const int64_t e0 = (int64_t)blockIdx.x * EPC;
  const int64_t e = e0 + el;
  const bool active = (el < EPC) && (e < n_elem);
  double* S = smem + (size_t)(el < EPC ? el : 0) * PER_ELEM;

  if (active) {
    for (int t = pr; t < NN * DIM; t += TPE) {
      const int a = t / DIM, c = t - a * DIM;
      S[OFF_X + t] = nodes[elements[e * NN + a] * DIM + c];
    }
    for (int t = pr; t < NQC * TS; t += TPE) {
      const int q = t / TS, c = t - q * TS;
      S[OFF_C + t] = tangent[((int64_t)q * n_elem + e) * TS + c];
    }
  }
  __syncthreads();

  // geometry of the element, all TPE threads of the element at work (round 1: one thread per Gauss point computed
  // J, J^-1 and all of B_q — a serial chain of ~300 FMAs on a third of the threads, with divergent constant-bank reads
  // of the reference table). The table is staged in shared memory once per CTA (Tq).
  double* Tq = smem + (size_t)EPC * PER_ELEM;
  for (int t = tid; t < NINT * DIM * NN; t += blockDim.x) Tq[t] = tab.bref[t];
  __syncthreads();
  // J[q][i][j] = sum_n bref[q][i][n] X[n][j]
  for (int t = pr; active && t < NINT * DIM * DIM; t += TPE) {
    const int q = t / (DIM * DIM), ij = t - q * DIM * DIM, i = ij / DIM, j = ij - i * DIM;
    double sum = 0.0;
#pragma unroll
    for (int n = 0; n < NN; ++n) sum = fma(Tq[(q * DIM + i) * NN + n], S[OFF_X + n * DIM + j], sum);
    S[OFF_J + t] = sum;
  }
  __syncthreads();
  for (int q = pr; active && q < NINT; q += TPE) {
    double J[DIM][DIM], inv[DIM][DIM];
#pragma unroll
    for (int i = 0; i < DIM; ++i)
#pragma unroll
      for (int j = 0; j < DIM; ++j) J[i][j] = S[OFF_J + (q * DIM + i) * DIM + j];
    const double det = inv_det<DIM>(J, inv);
    if (!(det > 0.0)) atomicOr(neg_jac, 1);
    S[OFF_WD + q] = tab.w[q] * det * (scale ? scale[e] : 1.0);
#pragma unroll
    for (int i = 0; i < DIM; ++i)
#pragma unroll
      for (int j = 0; j < DIM; ++j) S[OFF_J + NINT * DIM * DIM + (q * DIM + i) * DIM + j] = inv[i][j];
  }
  __syncthreads();
  // B[q][i][n] = sum_j J^-1[q][i][j] bref[q][j][n]
  for (int t = pr; active && t < NINT * DIM * NN; t += TPE) {
    const int qi = t / NN, n = t - qi * NN, q = qi / DIM;
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < DIM; ++j) sum = fma(S[OFF_J + NINT * DIM * DIM + qi * DIM + j], Tq[(q * DIM + j) * NN + n], sum);
    S[OFF_B + t] = sum;
  }
  __syncthreads();

  if (active) {
    // my RPT matrix rows (a_r, i), a_r = ap * RPT + r, and my first column node
    const int rg = pr / NCH, b0 = (pr - rg * NCH) * kCB;
    const int ap = rg / DPN, i = rg - ap * DPN;
    double acc[RPT][kCB][DPN];
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
      for (int c = 0; c < kCB; ++c)
#pragma unroll
        for (int k = 0; k < DPN; ++k) acc[r][c][k] = 0.0;

    // mechanics: row i of the tangent, C[i, :, :, :] (DIM^3 doubles); heat: kappa. One tensor for all Gauss points
    // (elastic): in registers for the whole loop.
    constexpr int CS = DIM * DPN * DIM;
    double Creg[PERGP ? 1 : CS];
    if (!PERGP) {
      const double* C0 = S + OFF_C + ((KIND == TFEM_KIND_MECH) ? i * DIM * DIM * DIM : 0);
#pragma unroll
      for (int t = 0; t < CS; ++t) Creg[t] = C0[t];
    }

#pragma unroll 1
    for (int q = 0; q < NINT; ++q) {
      const double* Cq = S + OFF_C + (PERGP ? q * TS : 0) + ((KIND == TFEM_KIND_MECH) ? i * DIM * DIM * DIM : 0);
      const double wd = S[OFF_WD + q];
      // stage 1: U[r][k][L] = sum_J (w detJ B[J,a_r]) C[i,J,k,L]   (heat: U[0][L] = sum_J (w detJ B[J,a]) kappa[J,L])
      double U[RPT][DPN][DIM];
      double bp[RPT][DIM];
#pragma unroll
      for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int j = 0; j < DIM; ++j) bp[r][j] = wd * S[OFF_B + (q * DIM + j) * NN + ap * RPT + r];
#pragma unroll
      for (int k = 0; k < DPN; ++k)
#pragma unroll
        for (int L = 0; L < DIM; ++L) {
          double cj[DIM];
#pragma unroll
          for (int Jx = 0; Jx < DIM; ++Jx) cj[Jx] = PERGP ? Cq[(Jx * DPN + k) * DIM + L] : Creg[(Jx * DPN + k) * DIM + L];
#pragma unroll
          for (int r = 0; r < RPT; ++r) {
            double u = 0.0;
#pragma unroll
            for (int Jx = 0; Jx < DIM; ++Jx) u = fma(bp[r][Jx], cj[Jx], u);
            U[r][k][L] = u;
          }
        }
      // stage 2: my columns; the column gradient is loaded once for the RPT rows
#pragma unroll
      for (int c = 0; c < kCB; ++c) {
        if (b0 + c < NN) {
          double br[DIM];
#pragma unroll
          for (int L = 0; L < DIM; ++L) br[L] = S[OFF_B + (q * DIM + L) * NN + b0 + c];
#pragma unroll
          for (int r = 0; r < RPT; ++r)
#pragma unroll
            for (int k = 0; k < DPN; ++k) {
              double t = acc[r][c][k];
#pragma unroll
              for (int L = 0; L < DIM; ++L) t = fma(U[r][k][L], br[L], t);
              acc[r][c][k] = t;
            }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int row = (ap * RPT + r) * DPN + i;
#pragma unroll
      for (int c = 0; c < kCB; ++c)
        if (b0 + c < NN) {
#pragma unroll
          for (int k = 0; k < DPN; ++k) S[OFF_K + row * ND + (b0 + c) * DPN + k] = acc[r][c][k];
        }
    }
  }
  __syncthreads();

  // coalesced write of the CTA's element matrices (contiguous in k_out)
  const int n_here = (int)((n_elem - e0) < EPC ? (n_elem - e0) : EPC);
  double* dst = k_out + e0 * (int64_t)(ND * ND);
  for (int t = tid; t < n_here * ND * ND; t += blockDim.x) {
    const int le = t / (ND * ND);
    dst[t] = smem[(size_t)le * PER_ELEM + OFF_K + (t - le * ND * ND)];
  }
}

template <int KIND, int DIM, int NN, int NINT>
int launch(const double* bref, const double* w, const double* nodes, const int64_t* elements,
           int64_t n_elem, const double* tangent, int per_gp, const double* scale, double* k_out,
           int32_t* neg_jac, cudaStream_t st) {
  constexpr int TPE = (NN / K1Rows<KIND, NN>::v) * ((KIND == TFEM_KIND_MECH) ? DIM : 1) * ((NN + kCB - 1) / kCB);
  constexpr int EPC = (128 / TPE) > 0 ? (128 / TPE) : 1;
  constexpr int THREADS = ((EPC * TPE + 31) / 32) * 32;
  constexpr int DPN = (KIND == TFEM_KIND_MECH) ? DIM : 1;
  constexpr int ND = NN * DPN;
  constexpr int TS = (KIND == TFEM_KIND_MECH) ? DIM * DIM * DIM * DIM : DIM * DIM;
  Tables<DIM, NN, NINT> tab;
  for (int i = 0; i < NINT * DIM * NN; ++i) tab.bref[i] = bref[i];
  for (int i = 0; i < NINT; ++i) tab.w[i] = w[i];
  const unsigned grid = (unsigned)((n_elem + EPC - 1) / EPC);
  if (per_gp) {
    const size_t per_elem = NN * DIM + NINT * DIM * NN + NINT + NINT * TS + ND * ND + 2 * NINT * DIM * DIM;
    const size_t bytes = (per_elem * EPC + NINT * DIM * NN) * sizeof(double);
    auto kern = k_integrate<KIND, DIM, NN, NINT, EPC, true>;
    if (bytes > 48 * 1024)
      TFEM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    kern<<<grid, THREADS, bytes, st>>>(tab, nodes, elements, n_elem, tangent, scale, k_out, neg_jac);
  } else {
    const size_t per_elem = NN * DIM + NINT * DIM * NN + NINT + TS + ND * ND + 2 * NINT * DIM * DIM;
    const size_t bytes = (per_elem * EPC + NINT * DIM * NN) * sizeof(double);
    auto kern = k_integrate<KIND, DIM, NN, NINT, EPC, false>;
    if (bytes > 48 * 1024)
      TFEM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    kern<<<grid, THREADS, bytes, st>>>(tab, nodes, elements, n_elem, tangent, scale, k_out, neg_jac);
  }
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

template <int KIND>
int dispatch(int dim, int nn, int n_int, const double* bref, const double* w, const double* nodes,
             const int64_t* elements, int64_t n_elem, const double* tangent, int per_gp,
             const double* scale, double* k_out, int32_t* neg_jac, cudaStream_t st) {
#define TFEM_CASE(D, N, Q)                                                                           \
  if (dim == D && nn == N && n_int == Q)                                                             \
    return launch<KIND, D, N, Q>(bref, w, nodes, elements, n_elem, tangent, per_gp, scale, k_out,    \
                                 neg_jac, st);
  TFEM_CASE(3, 8, 8)   // Hexa1  (elements.py:936-1075)
  TFEM_CASE(3, 20, 8)  // Hexa2  (elements.py:1078-1406, 2x2x2 reduced integration)
  TFEM_CASE(3, 4, 1)   // Tetra1 (elements.py:727-790)
  TFEM_CASE(3, 10, 4)  // Tetra2 (elements.py:793-933)
  TFEM_CASE(2, 4, 4)   // Quad1  (elements.py:491-567)
  TFEM_CASE(2, 8, 4)   // Quad2  (elements.py:599-695)
  TFEM_CASE(2, 3, 1)   // Tria1  (elements.py:309-360)
  TFEM_CASE(2, 6, 3)   // Tria2  (elements.py:385-466)
#undef TFEM_CASE
  set_last_error("invalid argument", "unsupported (dim, nodes per element, integration points)");
  return TFEM_ERR_INVALID;
}

}  // namespace
}  // namespace tfem

using namespace tfem;

extern "C" int tfem_integrate_k(int kind, int dim, int nn, int n_int, const double* bref_host,
                                const double* w_host, const double* nodes, const int64_t* elements,
                                int64_t n_elem, const double* tangent, int tangent_per_gp,
                                const double* scale, double* k_out, int32_t* neg_jac, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(bref_host && w_host && nodes && elements && tangent && k_out && neg_jac,
               "integrate_k: null pointer");
  TFEM_REQUIRE(n_elem >= 0, "integrate_k: negative n_elem");
  if (n_elem == 0) return TFEM_OK;
  if (kind == TFEM_KIND_MECH)
    return dispatch<TFEM_KIND_MECH>(dim, nn, n_int, bref_host, w_host, nodes, elements, n_elem, tangent,
                                    tangent_per_gp, scale, k_out, neg_jac, st);
  if (kind == TFEM_KIND_HEAT)
    return dispatch<TFEM_KIND_HEAT>(dim, nn, n_int, bref_host, w_host, nodes, elements, n_elem, tangent,
                                    tangent_per_gp, scale, k_out, neg_jac, st);
  set_last_error("invalid argument", "kind must be TFEM_KIND_MECH or TFEM_KIND_HEAT");
  return TFEM_ERR_INVALID;
}
