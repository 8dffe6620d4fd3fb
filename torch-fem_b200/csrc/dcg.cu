// Multi-GPU Jacobi-PCG with the two exchange steps of the path fused INTO the compute kernels, over peer
// memory (CUDA IPC mappings, NVLink 5 / NVSwitch stores) — no NCCL call and no host involvement inside
// the iteration. Peer-memory primitives: peer.cuh.
//
// The reference is single-GPU; its Krylov loop is cupy_cg (src/torchfem/sparse.py:414-421).
// Partitioning (torch-fem_b200/distributed.py): rank r owns a contiguous block of rows
// [row_lo, row_lo + n_owned) of its LOCAL numbering [low halo | owned | high halo]; owned rows are complete.
//
// Every rank's communication buffer holds two copies of the search direction p in its symmetric heap.
// Per iteration `it` (epochs are monotone counters, never reset):
//   k_sell_spmv      q = A p over the interior slices (rows that read no halo entry): the single-GPU kernel itself
//   k_dcg_spmv_halo  the slices that read halo entries: waits (once per warp) for the neighbours' halo flags of this
//                    iteration, fused p.q; the last CTA stores interior + halo sum into slot [my rank] of LL set 0
//                    on every rank (data + epoch in the same 8-byte words: one NVLink store, no fence, no flag).
//   k_dcg_update     one warp per CTA polls set 0 until all `world` slots carry this epoch, sums them in rank
//                    order (bit-identical on every rank), alpha = rho / p.q; x += alpha p; r -= alpha q; r.r and
//                    r.(D^-1 r); the last CTA publishes both sums to set 1 of every rank.
//   k_dcg_direction  waits for set 1, beta = rho' / rho, convergence test; p' = D^-1 r + beta p into the OTHER
//                    p buffer. The FIRST CTAs of the grid store the entries a neighbour needs straight into that
//                    neighbour's p' halo, fence once per CTA, and the last of them releases the halo flag — while
//                    the rest of the grid is still streaming, so the exchange overlaps the kernel.
// Hazards: the p buffers ping-pong and each reduction set is consumed before the barrier-like reduction of
// the other set completes, so no slot is overwritten while a peer may still read it.
// Every spin is bounded by `timeout_s`: on expiry the solve ends with TFEM_ERR_COMM on every rank instead of
// hanging the GPU.
#include <math.h>
#include <string.h>

#include "sell.cuh"
#include "peer.cuh"

namespace tfem {
namespace {

// trace slots of one CG iteration (ns of %globaltimer; *_WAIT_MAX are durations, maximum over the waiting threads)
enum Trace {
  TR_SPMV_BEGIN = 0, TR_SPMV_HALO_WAIT_MAX, TR_SPMV_LAST_CTA, TR_UPD_BEGIN, TR_UPD_WAITED, TR_UPD_LAST_CTA,
  TR_DIR_BEGIN, TR_DIR_WAITED, TR_DIR_HALO_RELEASED, TR_DIR_LAST_CTA
};
static_assert(TR_DIR_LAST_CTA < kTraceSlots, "trace slots");

constexpr int CH_P = 0;  // halo channel of the search direction

__device__ __forceinline__ double* p_buf(const Peers& P, int r, int which, int64_t vec) {
  return heap(P, r) + (int64_t)which * vec;
}

// true in every thread of the CTA that takes the last ticket
__device__ __forceinline__ bool last_cta(unsigned int* ticket) {
  __shared__ bool s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
    if (s_last) *ticket = 0u;
  }
  __syncthreads();
  return s_last;
}

// x = 0 ; r = b ; p0 = D^-1 b (owned rows [o, o+n)) ; halo of p0 to the neighbours ; (r.r, r.z, b.b) -> set 1
__global__ void __launch_bounds__(kVecThreads)
    k_dcg_init(int64_t o, int64_t n, int64_t vec, const double* __restrict__ b, const double* __restrict__ dinv,
               double* __restrict__ r, double* __restrict__ x, double* partials, unsigned int* ticket,
               Peers P, Halo H, int n_halo_ctas, unsigned long long ep_halo, unsigned long long ep_red) {
  __shared__ double s_red[kVecThreads / 32];
  double* p = p_buf(P, P.rank, 0, vec);
  if ((int)blockIdx.x < n_halo_ctas)
    halo_send_and_release(P, H, 0, CH_P, ep_halo, n_halo_ctas, ticket + 1, [&](int64_t i) { return dinv[i] * b[i]; });
  double rr = 0.0, rho = 0.0;
  for (int64_t k = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; k < n; k += (int64_t)gridDim.x * kVecThreads) {
    const int64_t i = o + k;
    const double bi = b[i];
    const double zi = dinv[i] * bi;
    x[i] = 0.0;
    r[i] = bi;
    p[i] = zi;
    rr = fma(bi, bi, rr);
    rho = fma(bi, zi, rho);
  }
  double mine[2], tot[2];
  mine[0] = block_sum<kVecThreads>(rr, s_red);
  mine[1] = block_sum<kVecThreads>(rho, s_red);
  if (publish_and_reduce<2>(mine, partials, ticket, tot)) {
    const double out[3] = {tot[0], tot[1], tot[0]};
    ll_publish<3>(P, 1, ep_red, out);
  }
}

// scalar state after the initial reduction (same on every rank)
__global__ void k_dcg_scalars_init(double* sc, Peers P, unsigned long long ep_red, double rtol, double atol) {
  __shared__ double s_out[4];
  __shared__ int s_ok;
  double tot[3];
  if (!ll_wait_sum<3>(P, 1, ep_red, tot, s_out, &s_ok)) {
    if (threadIdx.x == 0) sc[SC_DONE] = 4.0;
    return;
  }
  if (threadIdx.x == 0) {
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
}

// q = A p over the owned rows, in two kernels. The INTERIOR slices [ti_lo, ti_hi) — rows that reference no halo
// entry — go through the very kernel of the single-GPU driver (k_sell_spmv<DPN, true> on a slice range, sell.cuh; its
// fused p.q lands in sc[SC_PQ]): 98 % of the rows at config B, and no wait of any kind. The slices that read halo
// entries follow in k_dcg_spmv_halo: by then the neighbours' halo flags of this iteration are long set (they were
// released during the neighbours' direction kernels, a whole SpMV ago), so the exchange is hidden behind the interior
// rows. Keeping the halo logic out of the streaming kernel matters: fused into one kernel it cost 80 bytes of spills
// inside the slice loop and 3.5 % of the SpMV (1184 vs 1143 us at config B, profiles/r2_*_wait_trace*).
// p.q = interior sum + halo sum (fixed order) -> LL set 0 of every rank.
template <int DPN>
__global__ void __launch_bounds__(kSellWarps * 32, 8)
    k_dcg_spmv_halo(Sell A, int64_t ts_lo, int64_t ts_hi, int64_t ti_lo, int64_t ti_hi, int which, int64_t vec,
                    double* __restrict__ q, double* sc, double* partials, unsigned int* ticket, Peers P, Halo H,
                    long long it, unsigned long long ep_halo, unsigned long long ep_red) {
  __shared__ double s_red[kSellWarps];
  if (sc[SC_DONE] != 0.0) return;
  const double* __restrict__ p = p_buf(P, P.rank, which, vec);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * kSellWarps;
  const int64_t n_low = ti_lo - ts_lo, n_halo = n_low + (ts_hi - ti_hi);
  double dot = 0.0;
  int64_t j = (int64_t)blockIdx.x * kSellWarps + warp;
  if (j < n_halo) {
    int ok = 1;
    if (lane == 0) {
      const unsigned long long t0 = global_ns();
      ok = halo_wait_thread(P, H, CH_P, ep_halo) ? 1 : 0;
      if (!ok) sc[SC_DONE] = 4.0;
      trace_max(P, it, TR_SPMV_HALO_WAIT_MAX, global_ns() - t0);
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    for (; ok && j < n_halo; j += stride) {
      const int64_t t = j < n_low ? ts_lo + j : ti_hi + (j - n_low);
      // halo entries were written by a peer while the interior kernel was running: read them from L2 (ld.global.cg),
      // never through an L1 sector an interior row could have pulled in earlier
      const int64_t row = t * 32 + lane;
      const bool mine = row >= A.dot_lo && row < A.dot_hi;
      const double pr = mine ? __ldcg(p + row) : 0.0;
      const double acc = slice_row<DPN, true>(A, t, p, lane);
      if (mine) {
        q[row] = acc;
        dot = fma(acc, pr, dot);
      }
    }
  }
  const double bsum = block_sum<kSellWarps * 32>(dot, s_red);
  double mine[1] = {bsum}, tot[1];
  if (publish_and_reduce<1>(mine, partials, ticket, tot)) {
    double out[1];
    out[0] = sc[SC_PQ] + tot[0];   // interior + halo, the same order on every run
    ll_publish<1>(P, 0, ep_red, out);
    if (threadIdx.x == 0) trace_mark(P, it, TR_SPMV_LAST_CTA);
  }
}

// alpha = rho / p.q ; x += alpha p ; r -= alpha q ; (r.r, r.D^-1 r) -> LL set 1
__global__ void __launch_bounds__(kVecThreads, 8)
    k_dcg_update(int64_t o, int64_t n, int which, int64_t vec, const double* __restrict__ q,
                 const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r, double* sc,
                 double* partials, unsigned int* ticket, Peers P, long long it, unsigned long long ep_wait,
                 unsigned long long ep_red) {
  __shared__ double s_red[kVecThreads / 32];
  __shared__ double s_out[4];
  __shared__ int s_ok;
  if (sc[SC_DONE] != 0.0) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) trace_mark(P, it, TR_UPD_BEGIN);
  double pq[1];
  if (!ll_wait_sum<1>(P, 0, ep_wait, pq, s_out, &s_ok)) {
    if (threadIdx.x == 0) sc[SC_DONE] = 4.0;
    return;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) trace_mark(P, it, TR_UPD_WAITED);
  const double* __restrict__ p = p_buf(P, P.rank, which, vec);
  const double alpha = sc[SC_RHO] / pq[0];
  double rr = 0.0, rho = 0.0;
  for (int64_t k = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; k < n; k += (int64_t)gridDim.x * kVecThreads) {
    const int64_t i = o + k;
    x[i] = fma(alpha, p[i], x[i]);
    const double ri = fma(-alpha, q[i], r[i]);
    r[i] = ri;
    rr = fma(ri, ri, rr);
    rho = fma(ri * dinv[i], ri, rho);
  }
  double mine[2], tot[2];
  mine[0] = block_sum<kVecThreads>(rr, s_red);
  mine[1] = block_sum<kVecThreads>(rho, s_red);
  if (publish_and_reduce<2>(mine, partials, ticket, tot)) {
    ll_publish<2>(P, 1, ep_red, tot);
    if (threadIdx.x == 0) {
      sc[SC_PQ] = pq[0];
      sc[SC_ALPHA] = alpha;
      trace_mark(P, it, TR_UPD_LAST_CTA);
    }
  }
}

// beta = rho' / rho ; p' = D^-1 r + beta p (other buffer) ; the FIRST CTAs store the entries the neighbours need
// straight into their p' halos and release the halo flag while the rest of the grid is still updating p'
__global__ void __launch_bounds__(kVecThreads, 8)
    k_dcg_direction(int64_t o, int64_t n, int which, int64_t vec, const double* __restrict__ r,
                    const double* __restrict__ dinv, double* sc, unsigned int* ticket, Peers P, Halo H,
                    int n_halo_ctas, long long it, unsigned long long ep_wait, unsigned long long ep_halo) {
  __shared__ double s_out[4];
  __shared__ int s_ok;
  if (sc[SC_DONE] != 0.0) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) trace_mark(P, it, TR_DIR_BEGIN);
  double t2[2];
  if (!ll_wait_sum<2>(P, 1, ep_wait, t2, s_out, &s_ok)) {
    if (threadIdx.x == 0) sc[SC_DONE] = 4.0;
    return;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) trace_mark(P, it, TR_DIR_WAITED);
  const double rho_prev = sc[SC_RHO];
  const double beta = t2[1] / rho_prev;
  const double* __restrict__ p = p_buf(P, P.rank, which, vec);
  double* __restrict__ pn = p_buf(P, P.rank, which ^ 1, vec);
  if ((int)blockIdx.x < n_halo_ctas) {
    // dedicated halo CTAs (no share of the streaming loop, so they do not lengthen the kernel): the exchange
    // overlaps the update of p'. The neighbours need these entries only at the END of their next SpMV.
    const bool released = halo_send_and_release(P, H, (int64_t)(which ^ 1) * vec, CH_P, ep_halo, n_halo_ctas,
                                                ticket + 1, [&](int64_t i) { return fma(beta, p[i], dinv[i] * r[i]); });
    if (released && threadIdx.x == 0) trace_mark(P, it, TR_DIR_HALO_RELEASED);
  } else {
    const int64_t nb = (int64_t)gridDim.x - n_halo_ctas, b0 = (int64_t)blockIdx.x - n_halo_ctas;
    for (int64_t k = b0 * kVecThreads + threadIdx.x; k < n; k += nb * kVecThreads) {
      const int64_t i = o + k;
      pn[i] = fma(beta, p[i], dinv[i] * r[i]);
    }
  }
  if (last_cta(ticket) && threadIdx.x == 0) {
    sc[SC_RHO_PREV] = rho_prev;
    sc[SC_RHO] = t2[1];
    sc[SC_RR] = t2[0];
    sc[SC_BETA] = beta;
    sc[SC_ITERS] += 1.0;
    if (!isfinite(t2[0])) sc[SC_DONE] = 2.0;
    else if (sqrt(t2[0]) < sc[SC_TOL]) sc[SC_DONE] = 1.0;
    trace_mark(P, it, TR_DIR_LAST_CTA);
  }
}

// barrier at the end of a solve: nobody starts the next solve (whose first kernel overwrites reduction slots
// and p halos on its peers) before every rank has finished the last kernel of this one
__global__ void k_comm_barrier(double* sc, Peers P, unsigned long long ep) {
  if (sc && sc[SC_DONE] == 4.0) return;
  if (threadIdx.x < P.world) {
    __threadfence_system();
    st_release_sys(barrier_flag(P, threadIdx.x) + P.rank, ep);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long t0 = global_ns();
    const unsigned long long* fl = barrier_flag(P, P.rank);
    bool ok = true;
    for (int r = 0; r < P.world && ok; ++r) ok = spin_until(fl + r, ep, t0, P.timeout_ns);
    if (ok) fence_acq_rel_sys();
    if (!ok && sc) sc[SC_DONE] = 4.0;
  }
}

__global__ void k_dcg_trace_begin(Peers P, long long it, const double* sc) {
  if (sc[SC_DONE] == 0.0) trace_mark(P, it, TR_SPMV_BEGIN);
}

template <int DPN>
int launch_dcg_spmv(const Sell& A0, int64_t ts_lo, int64_t ts_hi, int64_t ti_lo, int64_t ti_hi, int which,
                    int64_t vec, const double* p, double* q, double* sc, double* partials, unsigned int* ticket,
                    const Peers& P, const Halo& H, long long it, unsigned long long ep_halo,
                    unsigned long long ep_red, cudaStream_t st) {
  if (P.trace && it >= P.trace_it0 && it < P.trace_it0 + P.trace_n) k_dcg_trace_begin<<<1, 1, 0, st>>>(P, it, sc);
  Sell A = A0;
  A.slice_lo = ti_lo;
  A.slice_hi = ti_hi;
  if (ti_hi > ti_lo) {
    int rc = launch_sell_t<DPN, true>(A, p, q, sc, partials, ticket, sc + SC_PQ, st);
    if (rc != TFEM_OK) return rc;
  } else {
    TFEM_CUDA(cudaMemsetAsync(sc + SC_PQ, 0, sizeof(double), st));
  }
  const int g = cached_resident_ctas(k_dcg_spmv_halo<DPN>, kSellWarps * 32);
  const int64_t n_halo = (ti_lo - ts_lo) + (ts_hi - ti_hi);
  const int64_t want = (n_halo + kSellWarps - 1) / kSellWarps;
  k_dcg_spmv_halo<DPN><<<(int)(want < g ? (want > 0 ? want : 1) : g), kSellWarps * 32, 0, st>>>(
      A0, ts_lo, ts_hi, ti_lo, ti_hi, which, vec, q, sc, partials, ticket, P, H, it, ep_halo, ep_red);
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

}  // namespace
}  // namespace tfem

using namespace tfem;

extern "C" int tfem_comm_create(int rank, int world, int64_t vec_doubles, void** comm_out,
                                void* ipc_handle_out) {
  TFEM_REQUIRE(comm_out && ipc_handle_out, "comm_create: null pointer");
  TFEM_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "comm_create: bad rank/world (<= 16 ranks)");
  TFEM_REQUIRE(vec_doubles > 0, "comm_create: empty vector");
  static_assert(sizeof(cudaIpcMemHandle_t) == TFEM_IPC_HANDLE_BYTES, "IPC handle size");
  Comm* c = new Comm();
  c->rank = rank;
  c->world = world;
  c->heap_doubles = 2 * pad32(vec_doubles);
  const size_t bytes = (size_t)HEADER_BYTES + (size_t)c->heap_doubles * sizeof(double);
  void* mem = nullptr;
  int rc = check_cuda(cudaMalloc(&mem, bytes), "cudaMalloc(comm buffer)");
  if (rc != TFEM_OK) { delete c; return rc; }
  rc = check_cuda(cudaMemset(mem, 0, bytes), "cudaMemset(comm buffer)");
  if (rc == TFEM_OK) rc = check_cuda(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
  memset(ipc_handle_out, 0, TFEM_IPC_HANDLE_BYTES);
  if (rc == TFEM_OK && world > 1) {
    cudaIpcMemHandle_t h;
    rc = check_cuda(cudaIpcGetMemHandle(&h, mem), "cudaIpcGetMemHandle");
    if (rc == TFEM_OK) memcpy(ipc_handle_out, &h, sizeof(h));
  }
  if (rc != TFEM_OK) { cudaFree(mem); delete c; return rc; }
  c->base[rank] = static_cast<char*>(mem);
  c->connected = (world == 1);
  *comm_out = c;
  return TFEM_OK;
}

extern "C" int tfem_comm_connect(void* comm, const void* all_handles) {
  Comm* c = static_cast<Comm*>(comm);
  TFEM_REQUIRE(c && all_handles, "comm_connect: null pointer");
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank || c->opened[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(all_handles) + (size_t)r * TFEM_IPC_HANDLE_BYTES, sizeof(h));
    void* ptr = nullptr;
    TFEM_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    c->base[r] = static_cast<char*>(ptr);
    c->opened[r] = true;
  }
  c->connected = true;
  return TFEM_OK;
}

extern "C" int tfem_comm_destroy(void* comm) {
  Comm* c = static_cast<Comm*>(comm);
  if (!c) return TFEM_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r)
    if (c->opened[r]) cudaIpcCloseMemHandle(c->base[r]);
  if (c->base[c->rank]) cudaFree(c->base[c->rank]);
  delete c;
  return TFEM_OK;
}

extern "C" int tfem_comm_heap(void* comm, void** heap_dev_out, int64_t* heap_doubles_out) {
  Comm* c = static_cast<Comm*>(comm);
  TFEM_REQUIRE(c && heap_dev_out && heap_doubles_out, "comm_heap: null pointer");
  *heap_dev_out = c->base[c->rank] + HEADER_BYTES;
  *heap_doubles_out = c->heap_doubles;
  return TFEM_OK;
}

extern "C" int tfem_comm_set_trace(void* comm, void* trace_dev, int64_t first_iteration, int n_iterations,
                                   int time_spmv) {
  Comm* c = static_cast<Comm*>(comm);
  TFEM_REQUIRE(c, "comm_set_trace: null communicator");
  TFEM_REQUIRE(!trace_dev || n_iterations > 0, "comm_set_trace: empty trace window");
  c->trace = static_cast<unsigned long long*>(trace_dev);
  c->trace_it0 = first_iteration;
  c->trace_n = trace_dev ? n_iterations : 0;
  c->time_spmv = time_spmv;
  return TFEM_OK;
}

extern "C" int tfem_dcg_solve(void* comm, const tfem_sell_t* a, int64_t row_lo, int64_t n_owned,
                              int64_t interior_lo, int64_t interior_hi, int n_sends,
                              const tfem_halo_send_t* sends, int n_recv, const int32_t* recv_peers,
                              const double* dinv, const double* b, double* x, double* work, double rtol,
                              double atol, int64_t maxiter, int check_every, double timeout_s,
                              double* info, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  Comm* c = static_cast<Comm*>(comm);
  TFEM_REQUIRE(c && c->connected, "dcg_solve: communicator missing or not connected");
  TFEM_REQUIRE(!c->broken, "dcg_solve: communicator is out of step after a failed solve; create a new one");
  int rc0 = check_sell(a);
  if (rc0 != TFEM_OK) return rc0;
  TFEM_REQUIRE(a->n_long == 0, "dcg_solve: matrices with long rows are not supported by the multi-GPU kernels");
  TFEM_REQUIRE(dinv && b && x && work && info, "dcg_solve: null pointer");
  const int64_t n_local = a->n_rows;
  const int64_t vec = c->heap_doubles / 2;
  TFEM_REQUIRE(n_owned > 0 && row_lo >= 0 && row_lo + n_owned <= n_local, "dcg_solve: bad row range");
  TFEM_REQUIRE(n_local <= vec, "dcg_solve: communicator vectors are too short for this matrix");
  if (maxiter <= 0) maxiter = 10 * n_local * c->world;
  if (check_every <= 0) check_every = 32;
  if (!(timeout_s > 0.0)) timeout_s = 20.0;

  const Peers P = make_peers(c, timeout_s);
  Halo H;
  rc0 = fill_halo(H, c, n_sends, sends, n_recv, recv_peers);
  if (rc0 != TFEM_OK) return rc0;

  Sell A = make_sell(a);
  const int64_t row_hi = row_lo + n_owned;
  A.dot_lo = row_lo;
  A.dot_hi = row_hi;
  if (interior_lo < row_lo) interior_lo = row_lo;
  if (interior_hi > row_hi) interior_hi = row_hi;
  const int64_t ts_lo = row_lo / 32, ts_hi = (row_hi + 31) / 32;
  int64_t ti_lo = (interior_lo + 31) / 32, ti_hi = interior_hi / 32;
  if (ti_lo < ts_lo) ti_lo = ts_lo;
  if (ti_hi > ts_hi) ti_hi = ts_hi;
  if (ti_hi < ti_lo) ti_hi = ti_lo;

  Work w = carve(work, n_local);
  const int vg = vec_grid(n_owned);
  const double* pbuf[2] = {reinterpret_cast<const double*>(c->base[c->rank] + HEADER_BYTES),
                           reinterpret_cast<const double*>(c->base[c->rank] + HEADER_BYTES) + vec};
  // init kernel: the first CTAs also send p0's halo; direction kernel: `hd` DEDICATED halo CTAs out of the grid
  int hc = halo_ctas(H.send_total, kVecThreads);
  if (hc > vg) hc = vg;
  // <= 32 entries per thread, four in flight (peer.cuh): at 433 k entries (two neighbours at config B) 32 CTAs with 53
  // entries per thread, one at a time, took longer than the whole streaming loop (direction kernel 68 vs 54 us,
  // profiles/r2_n8_*wait_trace*)
  int hd = H.send_total > 0 ? (int)((H.send_total + 8191) / 8192) : 0;
  if (hd > 64) hd = 64;
  if (H.send_total > 0 && hd < 1) hd = 1;
  const int vgd = vg + (vg + hd <= vec_grid_cap() ? hd : 0);             // stay within one resident wave
  if (vgd == vg && hd >= vg) hd = vg > 1 ? vg - 1 : 0;
  const unsigned long long E0 = c->epoch;
  TFEM_CUDA(cudaMemsetAsync(w.sc, 0, (SC_COUNT + kMaxPartials + 32) * sizeof(double), st));
  k_dcg_init<<<vg, kVecThreads, 0, st>>>(row_lo, n_owned, vec, b, dinv, w.r, x, w.partials, w.ticket, P, H, hc,
                                        E0 + 1, E0 + 1);
  TFEM_LAUNCH_CHECK();
  k_dcg_scalars_init<<<1, 32, 0, st>>>(w.sc, P, E0 + 1, rtol, atol);
  TFEM_LAUNCH_CHECK();
  double launches = 2, sc_host[SC_COUNT];
  int64_t issued = 0;
  int rc = TFEM_OK;
  // optional CUDA-event timing of the SpMV launches of the second batch (the kernel the roofline is quoted on)
  const int kEv = 32;
  cudaEvent_t ev[2 * kEv] = {};
  int n_ev = 0;
  double spmv_ms = 0.0;
  while (true) {
    TFEM_CUDA(cudaMemcpyAsync(sc_host, w.sc, sizeof(sc_host), cudaMemcpyDeviceToHost, st));
    TFEM_CUDA(cudaStreamSynchronize(st));
    if (n_ev > 0 && spmv_ms == 0.0) {
      float acc = 0.f, ms = 0.f;
      for (int e = 0; e < n_ev; ++e)
        if (cudaEventElapsedTime(&ms, ev[2 * e], ev[2 * e + 1]) == cudaSuccess) acc += ms;
      spmv_ms = acc / n_ev;
    }
    if (sc_host[SC_DONE] != 0.0 || issued >= maxiter) break;
    const int64_t batch = maxiter - issued < check_every ? maxiter - issued : check_every;
    const bool timed = c->time_spmv && issued > 0 && n_ev == 0;
    for (int64_t k = 0; k < batch; ++k) {
      const long long it = (long long)(issued + k);
      const int which = (int)(it & 1);
      const unsigned long long e1 = E0 + 1 + (unsigned long long)it, e2 = e1 + 1;
      const bool tm = timed && k < kEv;
      if (tm) {
        cudaEventCreate(&ev[2 * n_ev]);
        cudaEventCreate(&ev[2 * n_ev + 1]);
        cudaEventRecord(ev[2 * n_ev], st);
      }
      if (A.dpn == 3)
        rc = launch_dcg_spmv<3>(A, ts_lo, ts_hi, ti_lo, ti_hi, which, vec, pbuf[which], w.q, w.sc, w.partials, w.ticket, P, H, it, e1, e1, st);
      else if (A.dpn == 2)
        rc = launch_dcg_spmv<2>(A, ts_lo, ts_hi, ti_lo, ti_hi, which, vec, pbuf[which], w.q, w.sc, w.partials, w.ticket, P, H, it, e1, e1, st);
      else
        rc = launch_dcg_spmv<0>(A, ts_lo, ts_hi, ti_lo, ti_hi, which, vec, pbuf[which], w.q, w.sc, w.partials, w.ticket, P, H, it, e1, e1, st);
      if (rc != TFEM_OK) { c->broken = true; return rc; }
      if (tm) cudaEventRecord(ev[2 * n_ev++ + 1], st);
      k_dcg_update<<<vg, kVecThreads, 0, st>>>(row_lo, n_owned, which, vec, w.q, dinv, x, w.r, w.sc, w.partials,
                                              w.ticket, P, it, e1, e2);
      k_dcg_direction<<<vgd, kVecThreads, 0, st>>>(row_lo, n_owned, which, vec, w.r, dinv, w.sc, w.ticket, P, H, hd,
                                                  it, e2, e2);
    }
    rc = check_cuda(cudaGetLastError(), "dcg launch");
    if (rc != TFEM_OK) { c->broken = true; return rc; }
    issued += batch;
    launches += 4.0 * batch;
  }
  for (int e = 0; e < 2 * n_ev; ++e) cudaEventDestroy(ev[e]);
  c->epoch = E0 + (unsigned long long)issued + 2;
  if (sc_host[SC_DONE] != 4.0) {
    const double done = sc_host[SC_DONE];
    k_comm_barrier<<<1, 32, 0, st>>>(w.sc, P, ++c->barrier_epoch);
    TFEM_LAUNCH_CHECK();
    TFEM_CUDA(cudaMemcpyAsync(sc_host, w.sc, sizeof(sc_host), cudaMemcpyDeviceToHost, st));
    TFEM_CUDA(cudaStreamSynchronize(st));
    if (sc_host[SC_DONE] != 4.0) sc_host[SC_DONE] = done;
    launches += 1;
  }
  info[0] = sc_host[SC_ITERS];
  info[1] = sqrt(sc_host[SC_RR]);
  info[2] = sc_host[SC_BNRM];
  info[3] = sc_host[SC_DONE] == 1.0 ? 1.0 : 0.0;
  info[4] = (double)issued;
  info[5] = launches;
  info[6] = sc_host[SC_DONE];
  info[7] = spmv_ms;
  if (sc_host[SC_DONE] == 4.0) {
    c->broken = true;
    set_last_error("communication", "a peer did not deliver its halo / reduction within the timeout");
    return TFEM_ERR_COMM;
  }
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
