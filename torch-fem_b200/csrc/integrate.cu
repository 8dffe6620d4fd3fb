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
//
// Two kernels: k_integrate (above; tangents that differ per Gauss point: hyperelasticity, plasticity) and
// k_integrate_elastic (one tensor per element: linear elasticity, heat, topology optimisation — the benchmark's case),
// which sums over the Gauss points FIRST (a 3 x 3 geometric block per node pair) and contracts the tangent once:
// 10.6 kFMA per Hexa1 element, geometry in registers, k written from registers in whole sectors, global loads one
// element ahead. 15.1 -> 5.1 ms at BASELINE configs[1] (profiles/r2_k1_elastic.txt). TFEM_K1_ELASTIC=0 forces the
// general kernel.
#include <stdlib.h>

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

// ---- elastic tangent (ONE tensor for all Gauss points of an element: the benchmark's case, topology optimisation,
// every linear material). The sum over the Gauss points is taken FIRST:
//     M_ab[J][L] = sum_q w_q detJ_q B_q[J,a] B_q[L,b]            (a 3 x 3 "geometric" block per node pair)
//     k[(a,i),(b,k)] = sum_{J,L} C[i,J,k,L] M_ab[J][L]           (mechanics)      k[a,b] = sum kappa[J,L] M_ab[J][L] (heat)
// which halves the arithmetic (10.6 instead of 19 kFMA per Hexa1 element) and the shared-memory operand traffic:
// a thread owns one row node a and TWO column nodes (18 entries of M in registers, accumulated over the Gauss points
// from 10 LDS per 18 FMA), then contracts its two blocks with the tangent read by BROADCAST loads (every thread of an
// element reads the same address). One warp per Hexa1 element (32 = 8 x 4 tasks).
#ifndef TFEM_K1E_MINB
#define TFEM_K1E_MINB 4  // resident CTAs of 128 threads the register allocation must allow (<= 128 registers)
#endif

template <int NN>
struct ElasticTasks {
  static constexpr int v = NN * ((NN + 1) / 2);
  static constexpr int tpe = v >= 32 ? 32 : (v >= 16 ? 16 : 8);
};

template <int KIND, int DIM, int NN, int NINT>
struct ElasticLayout {  // doubles of shared memory per element; every region starts on a 16-byte boundary
  static constexpr int TS = (KIND == TFEM_KIND_MECH) ? DIM * DIM * DIM * DIM : DIM * DIM;
  static constexpr int even(int v) { return (v + 1) & ~1; }
  static constexpr int off_x = 0;
  static constexpr int off_b = off_x + even(NN * DIM);
  static constexpr int bq_stride = DIM * NN + 4;  // Gauss-point stride of B: 28 doubles (Hexa1) halves the conflicts of
                                                  // the geometry lanes' stores, which write 8 Gauss points at once
  static constexpr int off_wd = off_b + NINT * bq_stride;
  static constexpr int off_c = off_wd + even(NINT);
  static constexpr int per_elem = off_c + even(TS);
  static constexpr int tq_stride = DIM * NN + 2;  // one row per Gauss point: bref[q][:][:], then w[q]; the stride of
                                                  // 26 doubles (Hexa1) puts the 8 rows of a warp on different banks
  static constexpr int tq_size = NINT * tq_stride;
};

__device__ __forceinline__ void st_global_v4(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// Three phases per element, all operands of the arithmetic in registers, the element matrix written straight from
// registers in whole 32-byte sectors (no staging):
//   load      coordinates and tangent to shared memory
//   geometry  lane (q, part): J_q (redundantly in the lanes of q), its inverse, and B_q[:, n] for the lane's nodes
//   product   lane (a, b-pair): M over the Gauss points, contraction with C by 16-byte broadcast loads, stores
template <int KIND, int DIM, int NN, int NINT, int EPC>
__global__ void __launch_bounds__(((EPC * ElasticTasks<NN>::tpe + 31) / 32) * 32, TFEM_K1E_MINB)
    k_integrate_elastic(const __grid_constant__ Tables<DIM, NN, NINT> tab, const double* __restrict__ nodes,
                        const int64_t* __restrict__ elements, int64_t n_elem, const double* __restrict__ tangent,
                        const double* __restrict__ scale, double* __restrict__ k_out, int32_t* __restrict__ neg_jac) {
  using Lay = ElasticLayout<KIND, DIM, NN, NINT>;
  constexpr int DPN = (KIND == TFEM_KIND_MECH) ? DIM : 1;
  constexpr int ND = NN * DPN;
  constexpr int TS = Lay::TS;
  constexpr int TPE = ElasticTasks<NN>::tpe;
  constexpr bool WARP = (TPE == 32);  // one warp per element: every phase boundary is a warp barrier
  constexpr int NBP = (NN + 1) / 2;   // column-node pairs
  constexpr int NPART = (TPE / NINT < 1) ? 1 : ((TPE / NINT > NN) ? NN : TPE / NINT);  // lanes sharing a Gauss point
  constexpr int NPN = (NN + NPART - 1) / NPART;                                         // nodes per lane
  constexpr int OFF_X = Lay::off_x, OFF_B = Lay::off_b, OFF_WD = Lay::off_wd, OFF_C = Lay::off_c;
  constexpr int PER_ELEM = Lay::per_elem;
  constexpr int TQS = Lay::tq_stride, BQS = Lay::bq_stride;

  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x;
  const int el = tid / TPE;
  const int pr = tid - el * TPE;
  double* S = smem + (size_t)(el < EPC ? el : 0) * PER_ELEM;
  double* Tq = smem + (size_t)EPC * PER_ELEM;
  auto sync_elem = [] {
    if constexpr (WARP) __syncwarp(); else __syncthreads();
  };
  // the reference-gradient table once per (persistent) CTA
  for (int t = tid; t < NINT * DIM * NN; t += blockDim.x) {
    const int q = t / (DIM * NN);
    Tq[q * TQS + (t - q * DIM * NN)] = tab.bref[t];
  }
  for (int q = tid; q < NINT; q += blockDim.x) Tq[q * TQS + DIM * NN] = tab.w[q];
  __syncthreads();

  // The global loads of an element are issued one element ahead (node indices and tangent while the current
  // element's geometry runs, coordinates while its product runs) and land in shared memory at the top of the next
  // turn: a warp works through its elements serially, so without this every element pays two dependent DRAM round trips.
  constexpr int NXP = (NN * DIM + TPE - 1) / TPE;
  constexpr int NCP = (TS + TPE - 1) / TPE;
  int64_t pre_idx[NXP];
  double pre_x[NXP], pre_c[NCP], pre_scale = 1.0;
  auto fetch_idx_c = [&](int64_t ee) {
#pragma unroll
    for (int m = 0; m < NXP; ++m) {
      const int t = pr + m * TPE;
      if (NN * DIM % TPE == 0 || t < NN * DIM) pre_idx[m] = elements[ee * NN + t / DIM];  // used in fetch_x only
    }
#pragma unroll
    for (int m = 0; m < NCP; ++m) {
      const int t = pr + m * TPE;
      if (TS % TPE == 0 || t < TS) pre_c[m] = tangent[ee * TS + t];
    }
    if (scale) pre_scale = scale[ee];
  };
  auto fetch_x = [&] {
#pragma unroll
    for (int m = 0; m < NXP; ++m) {
      const int t = pr + m * TPE;
      if (NN * DIM % TPE == 0 || t < NN * DIM) pre_x[m] = nodes[pre_idx[m] * DIM + (t % DIM)];
    }
  };

  const int64_t n_blk = (n_elem + EPC - 1) / EPC;
  if ((int64_t)blockIdx.x < n_blk && el < EPC && (int64_t)blockIdx.x * EPC + el < n_elem) {
    fetch_idx_c((int64_t)blockIdx.x * EPC + el);
    fetch_x();
  }
  for (int64_t blk = blockIdx.x; blk < n_blk; blk += gridDim.x) {
    const int64_t e = blk * EPC + el;
    const bool active = (el < EPC) && (e < n_elem);
    const int64_t e_next = e + (int64_t)gridDim.x * EPC;
    const bool next_active = (el < EPC) && (blk + gridDim.x < n_blk) && (e_next < n_elem);
    const double my_scale = pre_scale;

    if (active) {
#pragma unroll
      for (int m = 0; m < NXP; ++m) {
        const int t = pr + m * TPE;
        if (NN * DIM % TPE == 0 || t < NN * DIM) S[OFF_X + t] = pre_x[m];
      }
#pragma unroll
      for (int m = 0; m < NCP; ++m) {
        const int t = pr + m * TPE;
        if (TS % TPE == 0 || t < TS) S[OFF_C + t] = pre_c[m];
      }
    }
    sync_elem();
    if (next_active) fetch_idx_c(e_next);

#pragma unroll 1
    for (int g = pr; active && g < NINT * NPART; g += TPE) {
      const int q = g / NPART, part = g - q * NPART;
      const double* T = Tq + q * TQS;
      // J[i][j] = sum_n bref[q][i][n] X[n][j], n ascending (the order of the general kernel)
      double J[DIM][DIM], inv[DIM][DIM];
#pragma unroll
      for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = 0; j < DIM; ++j) J[i][j] = 0.0;
      {
        double xr[NN * DIM];
#pragma unroll
        for (int m = 0; m < NN * DIM / 2; ++m) {
          const double2 v = reinterpret_cast<const double2*>(S + OFF_X)[m];
          xr[2 * m] = v.x;
          xr[2 * m + 1] = v.y;
        }
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
          double tr[NN];
#pragma unroll
          for (int m = 0; m < NN / 2; ++m) {
            const double2 v = reinterpret_cast<const double2*>(T + i * NN)[m];  // (i * NN) even for even NN only
            tr[2 * m] = v.x;
            tr[2 * m + 1] = v.y;
          }
          if constexpr (NN % 2 == 1) tr[NN - 1] = T[i * NN + NN - 1];
#pragma unroll
          for (int n = 0; n < NN; ++n)
#pragma unroll
            for (int j = 0; j < DIM; ++j) J[i][j] = fma(tr[n], xr[n * DIM + j], J[i][j]);
        }
      }
      const double det = inv_det<DIM>(J, inv);
      if (part == 0) {
        if (!(det > 0.0)) atomicOr(neg_jac, 1);
        S[OFF_WD + q] = T[DIM * NN] * det * my_scale;
      }
      // B[q][i][n] = sum_j J^-1[i][j] bref[q][j][n] for my nodes
#pragma unroll
      for (int m = 0; m < NPN; ++m) {
        const int n = part + m * NPART;
        if (NN % NPART == 0 || n < NN) {
          double tn[DIM];
#pragma unroll
          for (int j = 0; j < DIM; ++j) tn[j] = T[j * NN + n];
#pragma unroll
          for (int i = 0; i < DIM; ++i) {
            double sum = 0.0;
#pragma unroll
            for (int j = 0; j < DIM; ++j) sum = fma(inv[i][j], tn[j], sum);
            S[OFF_B + q * BQS + i * NN + n] = sum;
          }
        }
      }
    }
    sync_elem();
    if (next_active) fetch_x();

#pragma unroll 1
    for (int task = pr; active && task < NN * NBP; task += TPE) {
      asm volatile("" ::: "memory");  // keeps the 81 tangent entries out of registers across the tasks of a lane
      const int a = task / NBP, bp = task - a * NBP, b0 = bp * 2;
      const bool two = (NN % 2 == 0) || (b0 + 1 < NN);
      // M[J][c*DIM + L], c = 0, 1: the two column nodes
      double M[DIM][2 * DIM];
#pragma unroll
      for (int Jx = 0; Jx < DIM; ++Jx)
#pragma unroll
        for (int c = 0; c < 2 * DIM; ++c) M[Jx][c] = 0.0;
#pragma unroll 2
      for (int q = 0; q < NINT; ++q) {
        const double wd = S[OFF_WD + q];
        double ba[DIM], bb[2 * DIM];
#pragma unroll
        for (int Jx = 0; Jx < DIM; ++Jx) ba[Jx] = wd * S[OFF_B + q * BQS + Jx * NN + a];
#pragma unroll
        for (int L = 0; L < DIM; ++L) {
          if constexpr (NN % 2 == 0) {
            const double2 v = *reinterpret_cast<const double2*>(S + OFF_B + q * BQS + L * NN + b0);
            bb[L] = v.x;
            bb[DIM + L] = v.y;
          } else {
            bb[L] = S[OFF_B + q * BQS + L * NN + b0];
            bb[DIM + L] = two ? S[OFF_B + q * BQS + L * NN + b0 + 1] : 0.0;
          }
        }
#pragma unroll
        for (int Jx = 0; Jx < DIM; ++Jx)
#pragma unroll
          for (int c = 0; c < 2 * DIM; ++c) M[Jx][c] = fma(ba[Jx], bb[c], M[Jx][c]);
      }
      // contraction with the tangent: C is walked linearly in 16-byte broadcast loads; for each (i, k) the terms are
      // added in the order (J, L)
      double o0[DPN][DPN], o1[DPN][DPN];
#pragma unroll
      for (int i = 0; i < DPN; ++i)
#pragma unroll
        for (int k = 0; k < DPN; ++k) o0[i][k] = o1[i][k] = 0.0;
#pragma unroll
      for (int m = 0; m < (TS + 1) / 2; ++m) {
        double cv[2];
        if (2 * m + 1 < TS) {  // (two 8-byte loads instead: 5.10 ms either way)
          const double2 v = reinterpret_cast<const double2*>(S + OFF_C)[m];
          cv[0] = v.x;
          cv[1] = v.y;
        } else {
          cv[0] = S[OFF_C + 2 * m];
          cv[1] = 0.0;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int idx = 2 * m + h;
          if (idx < TS) {
            if constexpr (KIND == TFEM_KIND_MECH) {
              const int i = idx / (DIM * DIM * DIM), Jx = (idx / (DIM * DIM)) % DIM, k = (idx / DIM) % DIM, L = idx % DIM;
              o0[i][k] = fma(cv[h], M[Jx][L], o0[i][k]);
              o1[i][k] = fma(cv[h], M[Jx][DIM + L], o1[i][k]);
            } else {
              const int Jx = idx / DIM, L = idx % DIM;
              o0[0][0] = fma(cv[h], M[Jx][L], o0[0][0]);
              o1[0][0] = fma(cv[h], M[Jx][DIM + L], o1[0][0]);
            }
          }
        }
      }
      // the thread's 2*DPN consecutive entries of each of its DPN rows
      double* row = k_out + e * (int64_t)(ND * ND) + (int64_t)(a * DPN) * ND + b0 * DPN;
      if constexpr (NN % 2 == 1) {
#pragma unroll
        for (int i = 0; i < DPN; ++i)
#pragma unroll
          for (int k = 0; k < DPN; ++k) {
            row[i * ND + k] = o0[i][k];
            if (two) row[i * ND + DPN + k] = o1[i][k];
          }
      } else if constexpr (DPN == 1) {
        *reinterpret_cast<double2*>(row) = make_double2(o0[0][0], o1[0][0]);
      } else if constexpr (DPN == 2) {
#pragma unroll
        for (int i = 0; i < 2; ++i) st_global_v4(row + i * ND, o0[i][0], o0[i][1], o1[i][0], o1[i][1]);
      } else if constexpr ((ND * 8) % 32 == 0 && NBP % 2 == 0) {
        // 48 bytes per row and thread: an even pair starts a sector, an odd pair starts in the middle of one; the two
        // 16-byte halves of the shared sector go out in the same instruction
        const bool odd = bp & 1;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          double* r = row + i * ND;
          st_global_v4(odd ? r + 2 : r, odd ? o0[i][2] : o0[i][0], odd ? o1[i][0] : o0[i][1],
                       odd ? o1[i][1] : o0[i][2], odd ? o1[i][2] : o1[i][0]);
          *reinterpret_cast<double2*>(odd ? r : r + 4) =
              make_double2(odd ? o0[i][0] : o1[i][1], odd ? o0[i][1] : o1[i][2]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          double2* r = reinterpret_cast<double2*>(row + i * ND);
          r[0] = make_double2(o0[i][0], o0[i][1]);
          r[1] = make_double2(o0[i][2], o1[i][0]);
          r[2] = make_double2(o1[i][1], o1[i][2]);
        }
      }
    }
    sync_elem();
  }
}

inline bool use_elastic_kernel() {  // TFEM_K1_ELASTIC=0: the general kernel also for one-tensor tangents (A/B timing)
  static const bool off = getenv("TFEM_K1_ELASTIC") && atoi(getenv("TFEM_K1_ELASTIC")) == 0;
  return !off;
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
  } else if (use_elastic_kernel()) {
    using Lay = ElasticLayout<KIND, DIM, NN, NINT>;
    constexpr int TPE2 = ElasticTasks<NN>::tpe;
    constexpr int EPC2 = 128 / TPE2;
    constexpr int THREADS2 = ((EPC2 * TPE2 + 31) / 32) * 32;
    const size_t bytes = ((size_t)Lay::per_elem * EPC2 + Lay::tq_size) * sizeof(double);
    auto kern = k_integrate_elastic<KIND, DIM, NN, NINT, EPC2>;
    if (bytes > 48 * 1024)
      TFEM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    int resident = 0;  // persistent CTAs: the reference-gradient table is staged once per CTA
    TFEM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, THREADS2, bytes));
    const int64_t n_blk = (n_elem + EPC2 - 1) / EPC2;
    const int64_t cap = (int64_t)num_sms() * (resident > 0 ? resident : 1);
    kern<<<(unsigned)(n_blk < cap ? n_blk : cap), THREADS2, bytes, st>>>(tab, nodes, elements, n_elem, tangent, scale,
                                                                        k_out, neg_jac);
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
