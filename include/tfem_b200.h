/*
 * tfem_b200.h — C ABI of libtfem_b200.so: the B200 (sm_100a) kernels behind torch-fem's implicit-solve
 * hot path (element stiffness integration -> sparse assembly -> Jacobi-preconditioned Krylov solve).
 *
 * Conventions (modelled on the reference's only native boundary, the AmgX ctypes binding,
 * /root/reference/src/torchfem/amgx.py:148-208):
 *   - every entry point returns `int rc` (0 = TFEM_OK); `tfem_get_error_string(rc, buf, len)` describes it
 *     (amgx.py:195-201 `_check` / `AMGX_get_error_string`);
 *   - plain pointers and sizes only; "dev" pointers are device memory owned by the caller (torch tensors,
 *     `tensor.data_ptr()`), borrowed for the duration of the call; "host" pointers are host memory;
 *   - every call is stream-ordered on the `stream` argument (a `cudaStream_t` passed as void*; the Python
 *     side passes `torch.cuda.current_stream().cuda_stream`); no hidden synchronisation unless stated;
 *   - no global state besides a once-initialised attribute cache; float64 values, int32/int64 indices.
 *
 * Each entry point cites the reference code it replaces (paths relative to /root/reference).
 */
#ifndef TFEM_B200_H
#define TFEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFEM_OK 0
#define TFEM_ERR_INVALID 1      /* bad argument (null pointer, unsupported element, misaligned buffer) */
#define TFEM_ERR_CUDA 2         /* a CUDA runtime call or kernel launch failed; see error string */
#define TFEM_ERR_CAPACITY 3     /* a documented static limit was exceeded (node valence, index width) */
#define TFEM_ERR_NOT_CONVERGED 4/* Krylov solver hit maxiter (reference: RuntimeError "CG failed ...") */
#define TFEM_ERR_BREAKDOWN 5    /* Krylov breakdown (non-finite or non-positive curvature) */
#define TFEM_ERR_NCCL 6
#define TFEM_ERR_COMM 7          /* a peer did not deliver its halo / reduction within the timeout */

#define TFEM_KIND_MECH 0        /* vector field, tangent [.., d,d,d,d]  (base.py:1086-1090) */
#define TFEM_KIND_HEAT 1        /* scalar field, tangent [.., d,d]      (base.py:1272-1278) */

#define TFEM_METHOD_CG 0        /* sparse.py:414-421 (cupy_cg + Jacobi) */
#define TFEM_METHOD_MINRES 1    /* sparse.py:406-413 (cupy_minres + Jacobi) */

/* Library version (major*10000 + minor*100 + patch). */
int tfem_version(void);

/* Human-readable text for `rc`; for TFEM_ERR_CUDA/NCCL includes the last runtime error of this thread.
 * Mirrors AMGX_get_error_string (amgx.py:195-201). */
int tfem_get_error_string(int rc, char* buf, int len);

/* ---------------------------------------------------------------------------------------------------
 * K0 — sparsity pattern.  Replaces the packed-key sort/unique/searchsorted of FEM.__init__
 * (src/torchfem/base.py:78-118) with a node-graph build: node->element incidence, per-node sorted
 * unique neighbour list, expanded to dpn x dpn blocks. The result is bit-identical to the reference's
 * `glob_idx` (as CSR), `k_map` and `diag_map`, including the lone diagonal entry of nodes that no
 * element references (base.py:89-91).
 *
 * Phase 1 sizes the pattern, phase 2 fills caller-allocated arrays.
 * ------------------------------------------------------------------------------------------------- */

/* elements_dev: int64 [n_elem*nn] (the reference's connectivity dtype).
 * inc_ptr_dev : int32 [n_nod+1]   out: CSR offsets of the node->element-slot incidence
 * inc_list_dev: int32 [n_elem*nn] out: for every node the slots (e*nn + a) that reference it, ascending
 * blk_cnt_dev : int32 [n_nod]     out: number of distinct neighbour nodes (0 for an unreferenced node)
 * totals_dev  : int64 [4]         out: {nnzb (node blocks), nnz (scalar entries for dpn), max blk_cnt,
 *                                       max incident slots}
 * Fails with TFEM_ERR_CAPACITY if a node has more than 2048/nn incident elements or n_elem*nn*nn >= 2^31. */
int tfem_pattern_phase1(int64_t n_nod, int64_t n_elem, int nn, int dpn, const int64_t* elements_dev,
                        int32_t* inc_ptr_dev, int32_t* inc_list_dev, int32_t* blk_cnt_dev,
                        int64_t* totals_dev, void* stream);

/* node_ptr_dev: int64 [n_nod+1] out: block-CSR offsets (exclusive scan of blk_cnt)
 * adj_dev     : int32 [nnzb]    out: neighbour node ids, ascending per node
 * indptr_dev  : int64 [n_dofs+1] out: scalar CSR row offsets (== bincount/cumsum of glob_idx[0], sparse.py:389-390)
 * indices_dev : int32 [nnz]     out: scalar CSR column indices (== glob_idx[1])
 * diag_map_dev: int32 [n_dofs]  out: position of (i,i) (base.py:106-108)
 * src_ptr_dev : int64 [nnzb+1]  out: offsets into src of the element contributions of every node block
 * src_dev     : int32 [n_elem*nn*nn] out: contributions e*nn*nn + a*nn + b, ascending per block — the
 *               element-slot -> CSR permutation of the deterministic assembly (replaces k_map in K2) */
int tfem_pattern_phase2(int64_t n_nod, int64_t n_elem, int nn, int dpn, const int64_t* elements_dev,
                        const int32_t* inc_ptr_dev, const int32_t* inc_list_dev,
                        const int32_t* blk_cnt_dev, int64_t* node_ptr_dev, int32_t* adj_dev,
                        int64_t* indptr_dev, int32_t* indices_dev, int32_t* diag_map_dev,
                        int64_t* src_ptr_dev, int32_t* src_dev, void* stream);

/* Reference-compatible `k_map` (int32 [n_elem*(nn*dpn)^2], base.py:94-104): CSR position of every element
 * slot in k.ravel() order. Only needed for API compatibility; the assembly kernel uses `src`. */
int tfem_pattern_k_map(int64_t n_nod, int64_t n_elem, int nn, int dpn, const int64_t* elements_dev,
                       const int64_t* node_ptr_dev, const int32_t* adj_dev, const int64_t* indptr_dev,
                       int32_t* k_map_dev, void* stream);

/* Reference-compatible `glob_idx` rows (int64 [nnz]) from indptr (base.py:110-118). */
int tfem_pattern_coo_rows(int64_t n_dofs, const int64_t* indptr_dev, int64_t* rows_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * K1 — element matrices.  Replaces eval_shape_functions (base.py:293-314) + the Gauss-point loop's
 * stiffness branch (base.py:1086-1090 mechanics, :1272-1278 heat) + compute_k (solid.py:52-54,
 * planar.py:86-88).
 *   k_e[(p,i),(r,k)] = sum_q w_q detJ_q s_e sum_{J,L} B_q[J,p] C[i,J,k,L] B_q[L,r]     (MECH)
 *   k_e[p,r]         = sum_q w_q detJ_q s_e sum_{i,j} kappa[i,j] B_q[i,p] B_q[j,r]     (HEAT)
 * bref_host  : double [n_int*dim*nn]  reference-space gradients etype.B(ipoints) (host memory)
 * w_host     : double [n_int]         etype.iweights (host memory)
 * nodes_dev  : double [n_nod*dim];  elements_dev: int64 [n_elem*nn]
 * tangent_dev: double [(n_int if tangent_per_gp else 1), n_elem, dim^4 | dim^2]
 * scale_dev  : double [n_elem] (planar thickness) or NULL
 * k_dev      : double [n_elem*(nn*dpn)^2] out, row-major, local DOF order node-major / dof-minor
 * neg_jac_dev: int32 [1] in/out: set to 1 if any detJ <= 0 (caller raises the reference's
 *              ValueError("Negative Jacobian. Check element numbering."), base.py:311-312)
 * Supported (dim,nn,n_int): (3,8,8) Hexa1, (3,20,8) Hexa2, (3,4,1) Tetra1, (3,10,4) Tetra2,
 *                           (2,4,4) Quad1, (2,8,4) Quad2, (2,3,1) Tria1, (2,6,3) Tria2. */
int tfem_integrate_k(int kind, int dim, int nn, int n_int, const double* bref_host,
                     const double* w_host, const double* nodes_dev, const int64_t* elements_dev,
                     int64_t n_elem, const double* tangent_dev, int tangent_per_gp,
                     const double* scale_dev, double* k_dev, int32_t* neg_jac_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * K9/K10 — geometry contractions of the RESIDUAL evaluation, all Gauss points in one launch. The material
 * update between them stays with the caller (torch; the adjoint differentiates through it, sparse.py:689-705).
 *   tfem_elem_grad :  H[q,e,i,J] = s * sum_n u_e[e,n,i] B_q[e,J,n]        replaces `du @ B[i]^T` inside the Gauss
 *                     loop (base.py:1052; heat base.py:1241-1243) and eval_shape_functions (base.py:293-314)
 *   tfem_elem_force:  f_e[e,n,i] = sum_q s * sum_J B_q[e,J,n] P[q,e,i,J]  replaces `f += w * compute_f(detJ, B, P)`
 *                     (base.py:1082-1083; solid.py:56-58, planar.py:90-92)
 * s = 1 if !weighted, else w_q * detJ_q * scale[e] (scale_dev may be NULL). Each call is the transpose of the
 * other with the same `weighted` on the other side, so they are also each other's backward.
 * u_e_dev: double [n_elem, nn, dpn]; H_dev / P_dev: double [n_int, n_elem, dpn, dim]; f_e_dev: [n_elem, nn, dpn];
 * dpn = dim (mechanics) or 1 (heat). neg_jac_dev as in tfem_integrate_k. Element types as tfem_integrate_k. */
int tfem_elem_grad(int dim, int nn, int n_int, int dpn, const double* bref_host, const double* w_host,
                   const double* nodes_dev, const int64_t* elements_dev, int64_t n_elem, const double* u_e_dev,
                   const double* scale_dev, int weighted, double* H_dev, int32_t* neg_jac_dev, void* stream);
int tfem_elem_force(int dim, int nn, int n_int, int dpn, const double* bref_host, const double* w_host,
                    const double* nodes_dev, const int64_t* elements_dev, int64_t n_elem, const double* P_dev,
                    const double* scale_dev, int weighted, double* f_e_dev, int32_t* neg_jac_dev, void* stream);

/* K17 — tangent contraction of the elastic stress update, all Gauss points in one launch:
 *   out[q,e,i] = sum_k C[e,i,k] E[q,e,k]   (transpose != 0: C[e,k,i]),   i, k = flattened index pairs, m = d*d
 * replaces `einsum("...ijkl,...kl->...ij", C, de)` inside Material.step (materials/elasticity.py:119-127), which the
 * reference evaluates per Gauss point with a temporary of the size of C per point. C_dev: double [n_elem, m, m];
 * E_dev / out_dev: double [n_q, n_elem, m]. tfem_ddot_outer is its backward with respect to C:
 *   gC[e,i,k] = sum_q G[q,e,i] E[q,e,k]. */
int tfem_ddot(int m, int64_t n_q, int64_t n_elem, const double* C_dev, const double* E_dev, int transpose,
              double* out_dev, void* stream);
int tfem_ddot_outer(int m, int64_t n_q, int64_t n_elem, const double* G_dev, const double* E_dev, double* gC_dev,
                    void* stream);

/* ---------------------------------------------------------------------------------------------------
 * K2/K3 — deterministic assembly.  Replaces FEM.assemble_matrix (base.py:398-426): index_add_ scatter
 * (atomics on CUDA) becomes a gather over the precomputed `src` permutation with a fixed summation
 * order (element order), fused with the Dirichlet masking (rows/cols of constrained DOFs zeroed, unit
 * diagonal, base.py:414-419).
 * is_con_dev: uint8 [n_dofs] (1 = constrained) or NULL for no constraints (assembly.py:19,509 EMPTY)
 * vals_dev  : double [nnz] out, in the CSR order of tfem_pattern_phase2 (== reference COO order) */
int tfem_assemble(int64_t n_nod, int nn, int dpn, const int64_t* node_ptr_dev, const int32_t* adj_dev,
                  const int64_t* indptr_dev, const int64_t* src_ptr_dev, const int32_t* src_dev,
                  const double* k_dev, const uint8_t* is_con_dev, double* vals_dev, void* stream);

/* The same, fused with the Dirichlet lifting the first Newton step needs (base.py:708-741: the residual of the
 * prescribed increment du_bc is K_unconstrained du_bc on the free rows): lift_dev[row] = sum over constrained
 * columns c of K[row, c] * ubc_dev[c] for free rows, 0 for constrained rows — taken from the entries the
 * masking is about to zero, so no unconstrained copy of K is ever assembled. Fixed summation order.
 * ubc_dev: double [n_dofs] (entries at constrained DOFs are read); lift_dev: double [n_dofs] out or NULL. */
int tfem_assemble_bc(int64_t n_nod, int nn, int dpn, const int64_t* node_ptr_dev, const int32_t* adj_dev,
                     const int64_t* indptr_dev, const int64_t* src_ptr_dev, const int32_t* src_dev,
                     const double* k_dev, const uint8_t* is_con_dev, const double* ubc_dev, double* vals_dev,
                     double* lift_dev, void* stream);

/* The same, writing what the Krylov solve consumes in the same pass (the CSR -> SELL copy of tfem_sell_fill and
 * tfem_jacobi_setup become part of the assembly: 13 GB less traffic per step at BASELINE configs[1]):
 * slice_ptr_dev : int64 [ceil(n_dofs/32)+1], tfem_sell_slice_ptr of the same pattern (no long rows)
 * sell_vals_dev : double [slice_ptr[last]] out or NULL — the values in SELL-32 order (`tfem_sell_t.vals`), padding
 *                 entries and the rows past n_dofs of the last slice written as 0.0; bitwise equal to
 *                 tfem_sell_fill(vals_dev)
 * dinv_dev      : double [n_dofs] out or NULL — 1 / diagonal after the masking (== tfem_jacobi_setup)
 * vals_dev      : CSR values out, or NULL when only the solver's copy is wanted (one of the two must be given) */
int tfem_assemble_solve(int64_t n_nod, int nn, int dpn, const int64_t* node_ptr_dev, const int32_t* adj_dev,
                        const int64_t* indptr_dev, const int64_t* src_ptr_dev, const int32_t* src_dev,
                        const double* k_dev, const uint8_t* is_con_dev, const double* ubc_dev, double* vals_dev,
                        double* lift_dev, const int64_t* slice_ptr_dev, double* sell_vals_dev, double* dinv_dev,
                        void* stream);

/* ---------------------------------------------------------------------------------------------------
 * K5 — CSR SpMV y = A x (fp64 values, int32 columns, int64 row offsets).  Replaces cusparseSpMV inside
 * cupy_cg / cupy_minres (sparse.py:411,419). Algorithmic bytes: 12*nnz + 20*n_rows.
 *
 * The plan splits the nonzero stream into chunks of TFEM_SPMV_CHUNK entries; chunk c owns the rows whose
 * first entry lies in [c*CHUNK, (c+1)*CHUNK).
 * chunk_rows_dev: int32 [n_chunks+1] out, n_chunks = tfem_spmv_num_chunks(nnz). */
#define TFEM_SPMV_CHUNK 512
int64_t tfem_spmv_num_chunks(int64_t nnz);
int tfem_spmv_plan(int64_t n_rows, int64_t nnz, const int64_t* indptr_dev, int32_t* chunk_rows_dev,
                   void* stream);
int tfem_spmv(int64_t n_rows, int64_t nnz, const int64_t* indptr_dev, const int32_t* indices_dev,
              const double* vals_dev, const int32_t* chunk_rows_dev, const double* x_dev,
              double* y_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * SELL-32 — solver-internal matrix layout (sliced ELLPACK, slice height 32, no row permutation), built
 * from the CSR arrays above. It replaces the CuPy CSR object the reference hands to cuSPARSE
 * (sparse.py:391-393). Slice t holds rows [32t, 32t+32), width W_t = longest row rounded up to even;
 * entry k of row 32t+l is at slice_ptr[t] + (k/2)*64 + 2*l + (k%2): each lane streams its own row with
 * 128-bit loads and the warp reads 512 contiguous bytes per instruction. Padding = (own row, 0.0).
 * slice_ptr_dev: int64 [ceil(n/32)+1] out; slice_ptr[last] = padded entry count (read it back to size
 * sell_cols_dev int32 / sell_vals_dev double). tfem_sell_fill converts cols and/or vals (NULL = skip),
 * so a values-only refresh (new Newton iteration, same pattern) re-converts 8 B/nnz only. */
int tfem_sell_slice_ptr(int64_t n_rows, const int64_t* indptr_dev, int64_t* slice_ptr_dev, void* stream);
int tfem_sell_fill(int64_t n_rows, const int64_t* indptr_dev, const int32_t* indices_dev,
                   const double* vals_dev, const int64_t* slice_ptr_dev, int32_t* sell_cols_dev,
                   double* sell_vals_dev, void* stream);

/* The same for a rectangular operator with n_cols columns (AMG prolongation / restriction): the column index of a
 * padding entry is clamped to n_cols-1 so that it never reads past the end of x. */
int tfem_sell_fill_rect(int64_t n_rows, int64_t n_cols, const int64_t* indptr_dev, const int32_t* indices_dev,
                        const double* vals_dev, const int64_t* slice_ptr_dev, int32_t* sell_cols_dev,
                        double* sell_vals_dev, void* stream);

/* The same with the long-row rule of tfem_sell_t: rows with more than long_cap entries are left EMPTY in the slices
 * (pass TFEM_SELL_LONG_ROW; the uncapped functions above keep every row — the AMG operators use those). */
int tfem_sell_slice_ptr_capped(int64_t n_rows, const int64_t* indptr_dev, int64_t long_cap, int64_t* slice_ptr_dev,
                               void* stream);
int tfem_sell_fill_capped(int64_t n_rows, int64_t n_cols, const int64_t* indptr_dev, const int32_t* cols_dev,
                          const double* vals_dev, int64_t long_cap, const int64_t* slice_ptr_dev,
                          int32_t* sell_cols_dev, double* sell_vals_dev, void* stream);

/* Node-block column indices (optional, dpn = 2 or 3): FEM rows come in groups of dpn that share their column
 * blocks (column of entry k = dpn*adj[k/dpn] + k%dpn), so one int32 per (node, block) replaces one per entry
 * and the index stream drops from 4 to 4/dpn^2 bytes per nonzero (8.5 instead of 12 B/nnz for dpn = 3).
 * bcols[bslice_ptr[t] + kb*NPS + m] = block column kb of the m-th node touched by slice t (NPS = 12 for dpn 3,
 * 16 for dpn 2). Not valid for patterns with unreferenced nodes (rows of length 1): use scalar columns then.
 * bslice_ptr_dev: int64 [ceil(n/32)+1] out (last entry = length of bcols). */
int tfem_bsell_slice_ptr(int64_t n_rows, int dpn, const int64_t* slice_ptr_dev, int64_t* bslice_ptr_dev,
                         void* stream);
int tfem_bsell_fill(int64_t n_rows, int dpn, int64_t n_nod, const int64_t* node_ptr_dev,
                    const int32_t* adj_dev, const int64_t* bslice_ptr_dev, int32_t* bcols_dev, void* stream);

/* A SELL-32 matrix as the solver kernels take it (all device pointers). Either `cols` (scalar columns) or
 * `bcols` + `bslice_ptr` + `dpn` (node-block columns) must be given; if both are, block columns are used.
 *
 * Long rows: a row with more than TFEM_SELL_LONG_ROW entries (the rows of a reference point coupled to a whole face,
 * reference assembly.py:295-335: 11,169 entries against 81) would pad its whole slice to its own length and be walked
 * by ONE lane. tfem_sell_slice_ptr / tfem_sell_fill therefore leave such rows EMPTY in the slices, and the SpMV
 * computes them on a side path straight from the CSR arrays — one CTA per long row, products summed in a fixed
 * order (deterministic). n_long > 0 requires long_rows (ascending row numbers) and the three CSR arrays. Supported
 * by tfem_sell_spmv, tfem_krylov_solve and tfem_cg_stage; the AMG and the multi-GPU kernels reject such matrices. */
#define TFEM_SELL_LONG_ROW 1024
typedef struct tfem_sell {
  int64_t n_rows;
  const int64_t* slice_ptr;   /* [ceil(n/32)+1] */
  const int32_t* cols;        /* [padded nnz] or NULL */
  const double* vals;         /* [padded nnz] */
  const int64_t* bslice_ptr;  /* [ceil(n/32)+1] or NULL */
  const int32_t* bcols;       /* or NULL */
  int32_t dpn;                /* DOFs per node of the block structure (2 or 3), 0 if none */
  int32_t n_long;             /* rows longer than TFEM_SELL_LONG_ROW (0: none) */
  const int32_t* long_rows;   /* [n_long] */
  const int64_t* csr_indptr;  /* the CSR arrays the matrix was converted from (read for the long rows only) */
  const int32_t* csr_cols;
  const double* csr_vals;
} tfem_sell_t;

int tfem_sell_spmv(const tfem_sell_t* A, const double* x_dev, double* y_dev, void* stream);

/* Y = A X for a block of m vectors stored row-major (X[row * ldx + j], Y[row * ldy + j], j < m; ldx, ldy >= m): what the
 * eigensolver applies to its block (reference sparse.py:798-1011 passes K, M and the preconditioner to LOBPCG as block
 * operators). The matrix is streamed once per 4 vectors; keep m <= 4 and ldx = m per call (pack the columns): wider
 * leading dimensions waste the sectors of the gathered rows and are slower than single products. Column j of Y equals tfem_sell_spmv on column j bit for bit.
 * Matrices with long rows (n_long > 0) are refused with TFEM_ERR_INVALID: use tfem_sell_spmv per vector. */
int tfem_sell_spmm(const tfem_sell_t* A, int64_t m, const double* X_dev, int64_t ldx, double* Y_dev, int64_t ldy,
                   void* stream);

/* y = A^T x for a general (non-symmetric) CSR matrix — the adjoint of `Solve` with non-symmetric A
 * (sparse.py:203; tests/test_sparse.py:94-158). Deterministic: builds on a transposed copy made by the
 * caller with tfem_csr_transpose. */
int tfem_csr_transpose(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* indptr_dev,
                       const int32_t* indices_dev, const double* vals_dev, int64_t* t_indptr_dev,
                       int32_t* t_indices_dev, double* t_vals_dev, void* stream);

/* K4 — Jacobi preconditioner: dinv[i] = 1 / vals[diag_pos[i]]  (sparse.py:408-409 `diags(1/A.diagonal())`).
 * diag_pos_dev: int64 [n] position of the diagonal entry of every row, or -1 (then dinv = inf like 1/0). */
int tfem_csr_diag_positions(int64_t n_rows, const int64_t* indptr_dev, const int32_t* indices_dev,
                            int64_t* diag_pos_dev, void* stream);
int tfem_jacobi_setup(int64_t n_rows, const double* vals_dev, const int64_t* diag_pos_dev,
                      double* dinv_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * K5+K6 — Jacobi-preconditioned Krylov solve on one GPU.  Replaces cupy_cg / cupy_minres + the
 * per-iteration host synchronisation of the reference GPU path (sparse.py:406-421).
 * CG follows scipy/cupy `cg`: stop when ||r||_2 < max(atol, rtol*||b||_2), tested before every
 * iteration; maxiter <= 0 means 10*n. MINRES follows Paige-Saunders as in scipy `minres` (test1/test2),
 * maxiter <= 0 means 5*n.
 * Per iteration: SpMV fused with the p.q dot; one fused axpy/axpy/precondition/dot kernel; one direction
 * update. All reductions are fixed-order (deterministic); convergence is tested on the device and the
 * host polls a flag every `check_every` iterations (<=0: default 32).
 * The matrix is passed in the solver-internal SELL-32 layout (tfem_sell_* below).
 * x0_dev may be NULL (zero initial guess). work_dev: double [tfem_krylov_work_doubles(n)] scratch.
 * info_host: double [8] out (host): {iterations, final ||r||_2, ||b||_2, converged(1/0), spmv count,
 *            kernel launches, reserved, reserved}.
 * Returns TFEM_ERR_NOT_CONVERGED at maxiter (x still holds the last iterate). */
int64_t tfem_krylov_work_doubles(int64_t n_rows);
int tfem_krylov_solve(int method, const tfem_sell_t* A, const double* dinv_dev, const double* b_dev,
                      const double* x0_dev, double rtol, double atol, int64_t maxiter, int check_every,
                      double* x_dev, double* work_dev, double* info_host, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * K8 — matrix-free (element-by-element) operator on stored element matrices (optional operator of the Krylov
 * solve; the reference has none):  y = sum_e P_e^T k_e P_e x  with the Dirichlet masking of tfem_assemble applied
 * on the fly (constrained rows / columns dropped, unit diagonal). Deterministic gather over the node -> element
 * incidence lists of tfem_pattern_phase1; no pattern values, no assembly, no format conversion. Reads all of k per
 * product (8 nd^2 B per element), so it pays off when few Krylov iterations are spent per tangent. */
typedef struct tfem_ebe {
  int64_t n_nod;
  int32_t nn, dpn;
  const int32_t* inc_ptr;    /* [n_nod+1]      (tfem_pattern_phase1) */
  const int32_t* inc_list;   /* [n_elem*nn]    slots e*nn + a, ascending per node */
  const int64_t* elements;   /* [n_elem*nn] */
  const double* k;           /* [n_elem, nn*dpn, nn*dpn] */
  const uint8_t* is_con;     /* [n_nod*dpn] or NULL */
} tfem_ebe_t;
int tfem_ebe_spmv(const tfem_ebe_t* A, const double* x_dev, double* y_dev, void* stream);
/* diag[row] = sum of the diagonal entries of the incident element matrices (1 on constrained rows). */
int tfem_ebe_diag(const tfem_ebe_t* A, double* diag_dev, void* stream);
/* tfem_krylov_solve with the element operator in place of the assembled matrix (same arguments otherwise). */
int tfem_krylov_solve_ebe(int method, const tfem_ebe_t* A, const double* dinv_dev, const double* b_dev,
                          const double* x0_dev, double rtol, double atol, int64_t maxiter, int check_every,
                          double* x_dev, double* work_dev, double* info_host, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Multi-GPU Jacobi-PCG, one call per stage. The reference has no multi-GPU path; the partitioning
 * (contiguous node blocks, ghost elements integrated redundantly) is described in DESIGN.md §multi-GPU.
 * The host (torch-fem_b200/distributed.py) interleaves the NCCL collectives:
 *   stage 0  init: x = 0, r = b, p = dinv r              red <- local (r.r, r.z, b.b)     -> all-reduce(3)
 *   stage 1  scalars after init (tolerance, convergence of the zero iterate)
 *   -- per iteration: halo exchange of p, then
 *   stage 2  q = A p over all local rows                  red <- local p.q (owned rows)    -> all-reduce(1)
 *   stage 3  scalars (p.q)
 *   stage 4  x += alpha p ; r -= alpha q                  red <- local (r.r, r.z)          -> all-reduce(2)
 *   stage 5  scalars (alpha, beta, iteration count, convergence flag)
 *   stage 6  p = dinv r + beta p
 * Vectors have n_local = A_local->n_rows entries: owned rows [row_lo, row_lo + n_owned) plus halo rows; b, dinv, x are
 * indexed like the local rows. work_dev as in tfem_krylov_solve (sized for n_local); red_dev: double[4]. */
int tfem_cg_stage(int stage, const tfem_sell_t* A_local, int64_t row_lo, int64_t n_owned,
                  const double* dinv_dev, const double* b_dev, double* x_dev, double* work_dev,
                  double* red_dev, double rtol, double atol, void* stream);
/* Offset (in doubles) of a vector inside work_dev: which = 0 r, 1 p, 2 q, 3 device scalars. */
int64_t tfem_krylov_work_offset(int64_t n_rows, int which);
/* info_host[4] <- {iterations, ||r||_2, ||b||_2, done flag (0 running, 1 converged, 2 breakdown)}; synchronises. */
int tfem_krylov_state(int64_t n_rows, const double* work_dev, double* info_host, void* stream);


/* ---------------------------------------------------------------------------------------------------
 * Multi-GPU Jacobi-PCG with the halo exchange and the dot-product all-reduces fused into the compute
 * kernels over peer memory (CUDA IPC mappings of the other ranks' buffers; NVLink / NVSwitch stores).
 * One process per GPU. No NCCL call and no host work inside the iteration: the direction-update kernel
 * stores the entries a neighbour needs straight into that neighbour's copy of p, reductions are `world`
 * stores of <= 3 doubles plus a flag, summed in rank order by every consumer (bit-identical on all ranks).
 * Replaces the same reference loop as tfem_krylov_solve (cupy_cg, sparse.py:414-421).
 *
 * Setup: every rank calls tfem_comm_create (allocates its buffer: an 8 KB header of flags + a heap of 2*vec_doubles
 * doubles — the SAME vec_doubles on every rank, >= the longest local vector — and returns a 64-byte IPC handle), the
 * host exchanges the handles (e.g. torch.distributed.all_gather), every rank calls tfem_comm_connect with all `world`
 * handles in rank order, then a host barrier. world <= 16. The heap is symmetric: the same offset means the same
 * vector on every rank, so a kernel stores halo entries straight into a peer's copy (tfem_dcg_solve keeps its two
 * copies of p there; tfem_damg_* carve it into the vectors of the multigrid levels).
 * ------------------------------------------------------------------------------------------------- */
#define TFEM_IPC_HANDLE_BYTES 64
#define TFEM_MAX_NEIGHBOURS 8
#define TFEM_TRACE_SLOTS 16
int tfem_comm_create(int rank, int world, int64_t vec_doubles, void** comm_out, void* ipc_handle_out);
int tfem_comm_connect(void* comm, const void* all_handles /* world * TFEM_IPC_HANDLE_BYTES, host */);
int tfem_comm_destroy(void* comm);
/* This rank's heap (device pointer) and its length in doubles. */
int tfem_comm_heap(void* comm, void** heap_dev_out, int64_t* heap_doubles_out);
/* In-kernel profile of the cross-GPU waits (nsys is not available and ncu serialises kernels): with trace_dev != NULL
 * (device uint64 [n_iterations * TFEM_TRACE_SLOTS], zeroed by the caller) the kernels of iterations
 * [first_iteration, first_iteration + n_iterations) of the following solves stamp %globaltimer (ns) at fixed points:
 * slot 0 SpMV begin, 1 longest halo wait inside the SpMV (duration), 2 SpMV last CTA, 3 update begin, 4 update
 * reduction arrived, 5 update last CTA, 6 direction begin, 7 direction reduction arrived, 8 halo flags released,
 * 9 direction last CTA. time_spmv != 0: tfem_dcg_solve also times the SpMV launches of its second batch with CUDA
 * events and returns their mean (ms) in info_host[7]. trace_dev == NULL switches the trace off. */
int tfem_comm_set_trace(void* comm, void* trace_dev, int64_t first_iteration, int n_iterations, int time_spmv);

/* One entry per neighbour this rank sends halo values to: `count` entries of the local vector, taken at
 * src_idx[k] (device int32, local numbering) or src_start + k if src_idx is NULL, stored at dst_idx[k] (device
 * int32, the PEER's local numbering) or dst_start + k of the peer's vector. */
typedef struct tfem_halo_send {
  int32_t peer;
  int64_t count;
  const int32_t* src_idx;
  const int32_t* dst_idx;
  int64_t src_start, dst_start;
} tfem_halo_send_t;

/* A_local: rows in local numbering [low halo | owned | high halo]; only owned rows [row_lo, row_lo+n_owned) are
 * used. Rows in [interior_lo, interior_hi) must not reference halo columns (they are processed before the halo
 * has arrived); pass interior_lo == interior_hi if unknown. recv_peers_host: ranks that send halo values to this
 * rank. dinv, b, x: local-length device vectors (owned entries meaningful). work_dev as tfem_krylov_solve.
 * maxiter must be the same on every rank. timeout_s bounds every wait on a peer (<= 0: 20 s).
 * info_host as tfem_krylov_solve. Zero initial guess. All ranks must call this collectively. */
int tfem_dcg_solve(void* comm, const tfem_sell_t* A_local, int64_t row_lo, int64_t n_owned,
                   int64_t interior_lo, int64_t interior_hi, int n_sends, const tfem_halo_send_t* sends_host,
                   int n_recv, const int32_t* recv_peers_host, const double* dinv_dev, const double* b_dev,
                   double* x_dev, double* work_dev, double rtol, double atol, int64_t maxiter, int check_every,
                   double timeout_s, double* info_host, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * K11-K16 — aggregation algebraic multigrid (smoothed aggregation on the node graph) and AMG-preconditioned CG.
 * Replaces the reference's third-party AMG back ends: pyamg smoothed_aggregation_solver(A, B, smooth="jacobi") + scipy
 * cg/minres on the CPU (src/torchfem/sparse.py:493-512) and the AmgX aggregation-AMG solver on the GPU
 * (src/torchfem/amgx.py:71-98 config: V cycle, 1 pre / 1 post sweep, dense coarse solve; sparse.py:422-442).
 *
 * Operators are block CSR over nodes with d x d blocks (d = DOFs per node, 1..3): bptr int64 [nb+1], bcol int32 sorted
 * per row, values in the scalar-CSR order of the assembled matrix: block row I with m blocks owns d*d*m doubles at
 * d*d*bptr[I], entry (row DOF a, s-th block, column DOF c) at (a*m + s)*d + c. Level 0 is (node_ptr, adj, vals) of
 * tfem_pattern_phase2 / tfem_assemble unchanged. The host (torch-fem_b200/amg.py) owns all buffers and drives the
 * setup level by level:
 *   row_info -> rho -> aggregate -> prolongator_count/fill -> transpose_structure/values -> spgemm (A P, then R (A P)).
 * Everything is deterministic (integer atomics only; floating-point sums in a fixed order).
 * ------------------------------------------------------------------------------------------------- */

/* dinv[i] = 1/A_ii; iso[i] = 1 if every off-diagonal entry of row i is zero (Dirichlet rows after the masking of
 * tfem_assemble, base.py:414-419). fix_zero_diag != 0 (coarse levels): a zero diagonal (aggregate made of isolated
 * DOFs only) is replaced by 1 in vals_dev. */
int tfem_amg_row_info(int d, int64_t nb, const int64_t* bptr_dev, const int32_t* bcol_dev, double* vals_dev,
                      int fix_zero_diag, double* dinv_dev, uint8_t* iso_dev, void* stream);

/* Aggregation: maximal independent set of the node graph in Luby rounds with fixed pseudo-random keys; aggregates are
 * numbered in root order. distance = 1: roots pairwise non-adjacent, every node joins the adjacent root with the
 * largest key (radius-1 aggregates, ~12 nodes on Hexa1). distance = 2: roots more than two steps apart, nodes one step
 * from a root join it, the others join the aggregate of their neighbour with the largest key (radius-2 aggregates: for
 * graphs of low degree such as Tetra1, where radius 1 gives ~3 nodes per aggregate and the coarse operators fill in).
 * agg_dev: int32 [nb] out. state_work int8 [nb], flag_work uint8 [nb], index_work int32 [nb]: scratch. Synchronises
 * (one flag per round). */
int tfem_amg_aggregate(int64_t nb, const int64_t* bptr_dev, const int32_t* bcol_dev, int distance,
                       int8_t* state_work_dev, uint8_t* flag_work_dev, int32_t* index_work_dev, int32_t* agg_dev,
                       int64_t* n_agg_host, int32_t* rounds_host, void* stream);
/* The same on the subgraph of the nodes with exclude_dev[i] == 0 (uint8 [nb], NULL = none): excluded nodes (the halo
 * nodes of a partitioned mesh) are never roots and join nothing; their agg entries are undefined. */
int tfem_amg_aggregate_masked(int64_t nb, const int64_t* bptr_dev, const int32_t* bcol_dev, int distance,
                              const uint8_t* exclude_dev, int8_t* state_work_dev, uint8_t* flag_work_dev,
                              int32_t* index_work_dev, int32_t* agg_dev, int64_t* n_agg_host, int32_t* rounds_host,
                              void* stream);

/* Smoothed prolongator P = (I - omega D^-1 A) T, T[i, agg(i)] = diag(1 - iso_i). count: pptr_dev int64 [nb+1] out
 * (offsets); fill: pcol_dev int32 [pptr[nb]] (sorted per row), pvals_dev double [d*d*pptr[nb]]; max_row = longest
 * block row of A (rows up to 160 blocks are staged in shared memory; 0 = do not stage).
 * TFEM_ERR_CAPACITY if a node has more than 768 neighbours or touches more than 255 aggregates. */
int tfem_amg_prolongator_count(int d, int64_t nb, const int64_t* bptr_dev, const int32_t* bcol_dev,
                               const int32_t* agg_dev, int64_t* pptr_dev, void* stream);
int tfem_amg_prolongator_fill(int d, int64_t nb, const int64_t* bptr_dev, const int32_t* bcol_dev,
                              const double* vals_dev, const int32_t* agg_dev, const double* dinv_dev,
                              const uint8_t* iso_dev, double omega, const int64_t* pptr_dev, int32_t* pcol_dev,
                              double* pvals_dev, int max_row, void* stream);

/* The same for the block rows [row0, row0 + n_rows) only (the rows a rank owns): agg / dinv / iso are indexed like the
 * operator's rows and columns (halo columns included; agg values are arbitrary ints, e.g. GLOBAL aggregate numbers),
 * pptr [n_rows+1] / pcol / pvals are numbered from 0. */
int tfem_amg_prolongator_count_rows(int d, int64_t row0, int64_t n_rows, const int64_t* bptr_dev,
                                    const int32_t* bcol_dev, const int32_t* agg_dev, int64_t* pptr_dev, void* stream);
int tfem_amg_prolongator_fill_rows(int d, int64_t row0, int64_t n_rows, const int64_t* bptr_dev,
                                   const int32_t* bcol_dev, const double* vals_dev, const int32_t* agg_dev,
                                   const double* dinv_dev, const uint8_t* iso_dev, double omega,
                                   const int64_t* pptr_dev, int32_t* pcol_dev, double* pvals_dev, int max_row,
                                   void* stream);

/* Block transpose. structure: tptr int64 [n_cols+1], tcol int32 [nblk] sorted per row, tsrc int32 [nblk] = index of
 * the source block; values: tvals = transposed blocks gathered through tsrc (a values-only refresh repeats only this). */
int tfem_amg_transpose_structure(int64_t n_rows, int64_t n_cols, const int64_t* ptr_dev, const int32_t* col_dev,
                                 int64_t nblk, int64_t* tptr_dev, int32_t* tcol_dev, int32_t* tsrc_dev, void* stream);
int tfem_amg_transpose_values(int d, int64_t n_cols, const int64_t* ptr_dev, const double* vals_dev,
                              const int64_t* tptr_dev, const int32_t* tcol_dev, const int32_t* tsrc_dev,
                              double* tvals_dev, void* stream);

/* Block SpGEMM C = X Y. count: cptr int64 [nx+1] out; fill: ccol int32 sorted per row; numeric: cvals with every
 * entry summed in the order of X's row (fixed order). max_row = longest row of C (blocks); threads_per_row = 32 (a warp
 * per row of C) or 256 (a CTA per row: few, long rows).
 * TFEM_ERR_CAPACITY if a row of C has more than 12288 distinct block columns. */
int tfem_amg_spgemm_count(int64_t nx, const int64_t* xptr_dev, const int32_t* xcol_dev, const int64_t* yptr_dev,
                          const int32_t* ycol_dev, int64_t* cptr_dev, void* stream);
int tfem_amg_spgemm_fill(int64_t nx, const int64_t* xptr_dev, const int32_t* xcol_dev, const int64_t* yptr_dev,
                         const int32_t* ycol_dev, const int64_t* cptr_dev, int32_t* ccol_dev, void* stream);
int tfem_amg_spgemm_numeric(int d, int64_t nx, const int64_t* xptr_dev, const int32_t* xcol_dev,
                            const double* xvals_dev, const int64_t* yptr_dev, const int32_t* ycol_dev,
                            const double* yvals_dev, const int64_t* cptr_dev, const int32_t* ccol_dev,
                            double* cvals_dev, int max_row, int threads_per_row, void* stream);

/* A block-CSR operator as the coarse-level SpMV takes it (nb_rows == 0: not given). */
typedef struct tfem_bcsr {
  int64_t nb_rows;
  int64_t n_blocks;
  const int64_t* bptr;
  const int32_t* bcol;
  const double* vals;
  int32_t d;
} tfem_bcsr_t;

/* An operator of the hierarchy: SELL-32 (one row per lane; right for many short rows — the fine levels) or, if
 * bcsr.nb_rows > 0, block CSR with 8 / 32 / 256 threads per block row (few, long rows — the coarse levels and their
 * restrictions). The finest-level operator must be SELL-32. */
typedef struct tfem_amg_operator {
  tfem_sell_t sell;
  tfem_bcsr_t bcsr;
} tfem_amg_operator_t;

/* One level of the hierarchy as the cycle takes it. P / R are unused on the coarsest level. x, b, t: work vectors of
 * the level's length (b unused on level 0). */
#define TFEM_AMG_MAX_LEVELS 16
typedef struct tfem_amg_level {
  tfem_amg_operator_t A; /* level operator */
  tfem_amg_operator_t P; /* prolongation from the next coarser level: rows = this level */
  tfem_amg_operator_t R; /* restriction to the next coarser level:   rows = next level */
  const double* dinv;    /* 1 / diag(A) */
  double omega;          /* damped-Jacobi weight 4 / (3 rho(D^-1 A)) */
  double* x;
  double* b;
  double* t;
} tfem_amg_level_t;

/* Spectral radius of D^-1 A by `iterations` steps of the power method from a fixed start vector (synchronises).
 * work_dev: double [tfem_amg_work_doubles(n)]. */
int64_t tfem_amg_work_doubles(int64_t n_rows);
int tfem_amg_rho(const tfem_amg_operator_t* A, const double* dinv_dev, int iterations, double* work_dev,
                 double* rho_host, void* stream);

/* z = M r: V(1,1) cycle with damped Jacobi, dense inverse (coarse_inv_dev, row-major [n_c, n_c]) on the coarsest
 * level. Symmetric positive definite for SPD A, hence a valid CG preconditioner. r and z must not alias. */
int tfem_amg_vcycle(const tfem_amg_level_t* levels_host, int n_levels, const double* coarse_inv_dev,
                    const double* r_dev, double* z_dev, void* stream);

/* One V cycle on a block of nb <= 4 vectors stored row-major (r_block[row * nb + j]): the preconditioner step of the
 * eigensolver (reference sparse.py:798-1011 hands the preconditioner to LOBPCG as a block operator). The residual and
 * the post-smoothing sweep on the finest level read the matrix once for the block (tfem_sell_spmm with the cycle's
 * epilogues); restriction, coarse levels and prolongation run per vector. Column j of z_block equals tfem_amg_vcycle on
 * column j bit for bit. work_dev: 2 * n * nb doubles. */
int tfem_amg_vcycle_block(const tfem_amg_level_t* levels, int n_levels, const double* coarse_inv_dev, int nb,
                          const double* r_block_dev, double* z_block_dev, double* work_dev, void* stream);

/* CG preconditioned with the V cycle; stopping rule, maxiter and info_host as tfem_krylov_solve (CG). The r.z and p.q
 * dot products are fused into the last smoother / the SpMV. The host polls the convergence flag every iteration (one
 * iteration is a whole cycle). work_dev: double [tfem_amg_work_doubles(n)]. */
int tfem_amg_pcg_solve(const tfem_amg_level_t* levels_host, int n_levels, const double* coarse_inv_dev,
                       const double* b_dev, const double* x0_dev, double rtol, double atol, int64_t maxiter,
                       double* x_dev, double* work_dev, double* info_host, void* stream);

/* a9 — deterministic right-hand-side assembly, F[dpn*n + i] = sum over the slots (e, a) with elements[e,a] == n of
 * f_e[e, a*dpn + i], summed in ascending slot order (bitwise reproducible; replaces base.py:428-445
 * `F.index_add_(0, idx.ravel(), f.ravel())`, which uses floating-point atomics on CUDA). inc_ptr / inc_list: the
 * incidence lists of tfem_pattern_phase1. f_e: double [n_elem, nn*dpn]; F: double [n_nod*dpn] out. */
int tfem_assemble_rhs(int64_t n_nod, int dpn, const int32_t* inc_ptr_dev, const int32_t* inc_list_dev,
                      const double* f_e_dev, double* F_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Distributed AMG-PCG (SURVEY §8(f)-3 x §8(e)): the hierarchy of tfem_amg_* partitioned by rows over the ranks of a
 * tfem_comm. Aggregates are local to a rank, the prolongator smoothing and the Galerkin products are the global ones
 * (the host exchanges halo rows of P and A P at setup, torch-fem_b200/damg.py), levels below a size limit are gathered
 * and continued redundantly on every rank (`tail`, an ordinary tfem_amg_level_t hierarchy over the GLOBAL numbering of
 * its finest level). All exchange steps of the cycle run over peer memory inside this call: halo entries are stored
 * straight into the neighbours' copies of the vectors (which therefore live in the communicator's symmetric heap),
 * dot products are LL-protocol all-reduces; no NCCL call, no host work besides the convergence poll.
 *
 * Level l (distributed): A = square operator over this rank's LOCAL numbering of the level ([halo | owned | halo];
 * only the owned rows [own_lo, own_hi) are computed and written), P: rows = local numbering of level l, columns =
 * index space of level l+1 (local numbering of the next distributed level, or the global numbering of the tail);
 * R = rows of P^T for the owned coarse rows [c_own_lo, c_own_hi) of that index space, columns = local numbering of
 * level l (halo columns included: the residual's halo is exchanged before the restriction). lv.x and lv.t MUST be
 * vectors of the communicator's heap at the same offset on every rank; sends / recv_peers: halo plan of the level's
 * vectors in scalar indices (as tfem_dcg_solve).
 * ------------------------------------------------------------------------------------------------- */
typedef struct tfem_damg_level {
  tfem_amg_level_t lv;
  int64_t own_lo, own_hi;
  int64_t c_own_lo, c_own_hi;
  int32_t n_sends;
  const tfem_halo_send_t* sends;
  int32_t n_recv;
  const int32_t* recv_peers;
} tfem_damg_level_t;

/* CG on the distributed level-0 operator preconditioned with the distributed V(1,1) cycle. b, x: local-length device
 * vectors of level 0 (owned entries meaningful), zero initial guess. p_heap: heap vector of level-0 length (search
 * direction, exchanged every iteration). tail_b_heap: heap vector [tail_n] — the gathered right-hand side of the tail
 * (every rank stores its owned segment [c_own_lo, c_own_hi) of the last distributed level into every rank's copy);
 * tail_x: device vector [tail_n]. tail_levels / n_tail / tail_coarse_inv as tfem_amg_vcycle (the tail's finest level
 * may be block CSR). work_dev: double [tfem_amg_work_doubles(n_local of level 0)]. Collective; maxiter must be equal on
 * all ranks. info_host as tfem_krylov_solve. */
int tfem_damg_pcg_solve(void* comm, const tfem_damg_level_t* levels_host, int n_levels,
                        const tfem_amg_level_t* tail_levels_host, int n_tail, const double* tail_coarse_inv_dev,
                        int64_t tail_n, double* tail_b_heap, double* tail_x_dev, const double* b_dev, double* x_dev,
                        double* p_heap, double* work_dev, double rtol, double atol, int64_t maxiter, double timeout_s,
                        double* info_host, void* stream);

/* K7 — adjoint matrix gradient on the pattern: g[p] = -lambda[row(p)] * x[col(p)]
 * (sparse.py:212-216 `val = -gradb[row] * x[col]`). */
int tfem_adjoint_matrix_grad(int64_t n_rows, const int64_t* indptr_dev, const int32_t* indices_dev,
                             const double* lambda_dev, const double* x_dev, double* g_dev,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TFEM_B200_H */
