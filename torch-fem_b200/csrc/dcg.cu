// Multi-GPU Jacobi-PCG with the two exchange steps of the path fused INTO the compute kernels, over peer
// memory (CUDA IPC mappings, NVLink 5 / NVSwitch stores) — no NCCL call and no host involvement inside
// the iteration.
//
// The reference is single-GPU; its Krylov loop is cupy_cg (src/torchfem/sparse.py:414-421).
// Partitioning (torch-fem_b200/distributed.py): rank r owns a contiguous block of rows
// [row_lo, row_lo + n_owned) of its LOCAL numbering [low halo | owned | high halo]; owned rows are complete.
//
// Every rank allocates one communication buffer (header + two copies of the search direction p) and maps
// the buffers of all peers. Per iteration `it` (epochs are monotone counters, never reset):
//   k_dcg_spmv       q = A p over owned rows. Interior slices first; then the CTA waits until every
//                    neighbour's halo flag has reached the epoch of this iteration and does the slices
//                    that read halo entries. Fused p.q; the last CTA STORES the local sum into slot
//                    [my rank] of reduction set A in every rank's buffer, then releases a flag there.
//   k_dcg_update     waits for all `world` flags of set A, sums the slots in rank order (bit-identical on
//                    every rank), alpha = rho / p.q; x += alpha p; r -= alpha q; r.r and r.(D^-1 r);
//                    the last CTA publishes both sums to set B of every rank.
//   k_dcg_direction  waits for set B, beta = rho' / rho, convergence test; p' = D^-1 r + beta p into the OTHER
//                    p buffer and, for the entries a neighbour needs, straight into that neighbour's p' halo
//                    (peer stores); the last CTA releases the halo flags of the next epoch on the neighbours.
// All-reduce = world stores of <= 3 doubles + a flag per rank (latency of one NVLink store, ~2 us) instead
// of an NCCL kernel (~15-25 us at 8 GPUs); halo exchange = the direction kernel's own stores.
// Hazards: the p buffers ping-pong and each reduction set is consumed before the barrier-like reduction of
// the other set completes, so no slot is overwritten while a peer may still read it.
// Every spin is bounded by `timeout_s`: on expiry the solve ends with TFEM_ERR_COMM on every rank instead of
// hanging the GPU.
#include <math.h>
#include <string.h>

#include "sell.cuh"

namespace tfem {
namespace {

constexpr int kMaxRanks = 16;
constexpr int kMaxNbr = TFEM_MAX_NEIGHBOURS;
// header layout of a communication buffer (bytes)
constexpr int64_t OFF_HALO_FLAG = 0;     // uint64 [kMaxRanks]  epoch of the last halo delivered by rank s
constexpr int64_t OFF_RED_FLAG = 256;    // uint64 [2][kMaxRanks]
constexpr int64_t OFF_RED_VAL = 1024;    // double [2][kMaxRanks][4]
constexpr int64_t HEADER_BYTES = 4096;

struct Comm {
  int rank = 0, world = 1;
  int64_t vec = 0;  // doubles per p buffer (padded)
  char* base[kMaxRanks] = {};
  bool opened[kMaxRanks] = {};
  unsigned long long epoch = 0;
  bool broken = false;
  bool connected = false;
};

struct Peers {
  int rank, world;
  int64_t vec;
  char* base[kMaxRanks];
  unsigned long long timeout_ns;
};

struct Halo {
  int n_send;
  int send_peer[kMaxNbr];
  int64_t send_count[kMaxNbr];
  const int32_t* send_src[kMaxNbr];  // local indices, or nullptr: contiguous from src0
  const int32_t* send_dst[kMaxNbr];  // indices in the peer's local numbering, or nullptr: contiguous from dst0
  int64_t src0[kMaxNbr], dst0[kMaxNbr];
  int n_recv;
  int recv_peer[kMaxNbr];
};

__device__ __forceinline__ double* p_buf(const Peers& P, int r, int which) {
  return reinterpret_cast<double*>(P.base[r] + HEADER_BYTES) + (int64_t)which * P.vec;
}
__device__ __forceinline__ unsigned long long* halo_flag(const Peers& P, int r) {
  return reinterpret_cast<unsigned long long*>(P.base[r] + OFF_HALO_FLAG);
}
__device__ __forceinline__ unsigned long long* red_flag(const Peers& P, int r, int set) {
  return reinterpret_cast<unsigned long long*>(P.base[r] + OFF_RED_FLAG) + set * kMaxRanks;
}
__device__ __forceinline__ double* red_val(const Peers& P, int r, int set) {
  return reinterpret_cast<double*>(P.base[r] + OFF_RED_VAL) + set * kMaxRanks * 4;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// one thread: spin until *flag >= epoch; false on timeout
__device__ __forceinline__ bool spin_until(const unsigned long long* flag, unsigned long long epoch,
                                           unsigned long long t0, unsigned long long timeout_ns) {
  while (ld_acquire_sys(flag) < epoch) {
    if (global_ns() - t0 > timeout_ns) return false;
    __nanosleep(40);
  }
  return true;
}

// Whole CTA: wait until every rank has published reduction `set` for `epoch`, sum the slots in rank order.
// Returns false (in all threads) on timeout; sc[SC_DONE] is then 4.
template <int K>
__device__ __forceinline__ bool reduce_wait(const Peers& P, int set, unsigned long long epoch, double (&out)[K],
                                            double* sc) {
  __shared__ double s_out[4];
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    const unsigned long long t0 = global_ns();
    const unsigned long long* fl = red_flag(P, P.rank, set);
    bool ok = true;
    for (int r = 0; r < P.world && ok; ++r) ok = spin_until(fl + r, epoch, t0, P.timeout_ns);
    if (ok) {
      const double* v = red_val(P, P.rank, set);
#pragma unroll
      for (int j = 0; j < K; ++j) {
        double s = 0.0;
        for (int r = 0; r < P.world; ++r) s += __ldcv(v + r * 4 + j);
        s_out[j] = s;
      }
    } else {
      sc[SC_DONE] = 4.0;
    }
    s_ok = ok ? 1 : 0;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < K; ++j) out[j] = s_out[j];
  const bool ok = s_ok != 0;
  __syncthreads();
  return ok;
}

// warp 0 of the last CTA: lane r stores this rank's sums into rank r's slot and releases the flag there
template <int K>
__device__ __forceinline__ void reduce_publish(const Peers& P, int set, unsigned long long epoch,
                                               const double (&tot)[K]) {
  const int lane = threadIdx.x & 31;
  if (lane < P.world) {
    double* v = red_val(P, lane, set) + P.rank * 4;
#pragma unroll
    for (int j = 0; j < K; ++j) v[j] = tot[j];
    __threadfence_system();
    st_release_sys(red_flag(P, lane, set) + P.rank, epoch);
  }
}

// true in every thread of the CTA that takes the last ticket; all peer stores of the grid are then visible
// system-wide once the caller has executed __threadfence_system()
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
  if (s_last) __threadfence_system();
  return s_last;
}

__device__ __forceinline__ void release_halo_flags(const Peers& P, const Halo& H, unsigned long long epoch) {
  const int lane = threadIdx.x;
  if (lane < H.n_send) st_release_sys(halo_flag(P, H.send_peer[lane]) + P.rank, epoch);
}

// value(i) for every entry a neighbour needs -> the neighbour's p buffer `which`
template <typename F>
__device__ __forceinline__ void halo_send(const Peers& P, const Halo& H, int which, F value) {
  bool stored = false;
  for (int s = 0; s < H.n_send; ++s) {
    double* dst = p_buf(P, H.send_peer[s], which);
    const int32_t* si = H.send_src[s];
    const int32_t* di = H.send_dst[s];
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < H.send_count[s];
         k += (int64_t)gridDim.x * blockDim.x) {
      const int64_t i = si ? (int64_t)si[k] : H.src0[s] + k;
      const int64_t d = di ? (int64_t)di[k] : H.dst0[s] + k;
      dst[d] = value(i);
      stored = true;
    }
  }
  if (stored) __threadfence_system();  // my peer stores are visible system-wide before I take a ticket
}

// x = 0 ; r = b ; p0 = D^-1 b (owned rows [o, o+n)) ; halo of p0 to the neighbours ; (r.r, r.z, b.b) -> set B
__global__ void __launch_bounds__(kVecThreads)
    k_dcg_init(int64_t o, int64_t n, const double* __restrict__ b, const double* __restrict__ dinv,
               double* __restrict__ r, double* __restrict__ x, double* partials, unsigned int* ticket,
               Peers P, Halo H, unsigned long long ep_halo, unsigned long long ep_red) {
  __shared__ double s_red[kVecThreads / 32];
  double* p = p_buf(P, P.rank, 0);
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
  halo_send(P, H, 0, [&](int64_t i) { return dinv[i] * b[i]; });
  double mine[2], tot[2];
  mine[0] = block_sum<kVecThreads>(rr, s_red);
  mine[1] = block_sum<kVecThreads>(rho, s_red);
  if (publish_and_reduce<2>(mine, partials, ticket, tot)) {
    __threadfence_system();
    const double out[3] = {tot[0], tot[1], tot[0]};
    reduce_publish<3>(P, 1, ep_red, out);
    release_halo_flags(P, H, ep_halo);
  }
}

// scalar state after the initial reduction (same on every rank)
__global__ void k_dcg_scalars_init(double* sc, Peers P, unsigned long long ep_red, double rtol, double atol) {
  double tot[3];
  if (!reduce_wait<3>(P, 1, ep_red, tot, sc)) return;
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

// q = A p over the owned slices [ts_lo, ts_hi). One strided index space: the interior slices [ti_lo, ti_hi)
// first, then the slices that read halo entries. A warp that reaches the second part waits (once) until every
// neighbour's halo flag has reached ep_halo; by then the flags are normally long set, so the exchange is
// hidden behind the interior rows. p.q over owned rows -> set A.
template <int DPN>
__global__ void __launch_bounds__(kSellWarps * 32, 8)
    k_dcg_spmv(Sell A, int64_t ts_lo, int64_t ts_hi, int64_t ti_lo, int64_t ti_hi, int which,
               double* __restrict__ q, double* sc, double* partials, unsigned int* ticket, Peers P, Halo H,
               unsigned long long ep_halo, unsigned long long ep_red) {
  __shared__ double s_red[kSellWarps];
  if (sc[SC_DONE] != 0.0) return;
  const double* __restrict__ p = p_buf(P, P.rank, which);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * kSellWarps;
  const int64_t n_int = ti_hi - ti_lo, n_low = ti_lo - ts_lo, n_all = ts_hi - ts_lo;
  double dot = 0.0;
  int64_t j = (int64_t)blockIdx.x * kSellWarps + warp;
  for (; j < n_int; j += stride) {
    const int64_t t = ti_lo + j;
    const int64_t row = t * 32 + lane;
    const bool mine = row >= A.dot_lo && row < A.dot_hi;
    const double pr = mine ? __ldg(p + row) : 0.0;   // requested before the row is streamed (see krylov.cu)
    const double acc = slice_row<DPN>(A, t, p, lane);
    if (mine) {
      q[row] = acc;
      dot = fma(acc, pr, dot);
    }
  }
  if (j < n_all) {
    int ok = 1;
    if (lane == 0) {
      const unsigned long long t0 = global_ns();
      const unsigned long long* fl = halo_flag(P, P.rank);
      for (int s = 0; s < H.n_recv && ok; ++s) ok = spin_until(fl + H.recv_peer[s], ep_halo, t0, P.timeout_ns) ? 1 : 0;
      if (!ok) sc[SC_DONE] = 4.0;
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    for (; ok && j < n_all; j += stride) {
      const int64_t jb = j - n_int;
      const int64_t t = jb < n_low ? ts_lo + jb : ti_hi + (jb - n_low);
      // halo entries were written by a peer while this kernel may already have been running: read them from
      // L2 (ld.global.cg), never through an L1 sector an interior row could have pulled in earlier
      const int64_t row = t * 32 + lane;
      const bool mine = row >= A.dot_lo && row < A.dot_hi;
      const double pr = mine ? __ldg(p + row) : 0.0;
      const double acc = slice_row<DPN, true>(A, t, p, lane);
      if (mine) {
        q[row] = acc;
        dot = fma(acc, pr, dot);
      }
    }
  }
  const double bsum = block_sum<kSellWarps * 32>(dot, s_red);
  double mine[1] = {bsum}, tot[1];
  if (publish_and_reduce<1>(mine, partials, ticket, tot)) reduce_publish<1>(P, 0, ep_red, tot);
}

// alpha = rho / p.q ; x += alpha p ; r -= alpha q ; (r.r, r.D^-1 r) -> set B
__global__ void __launch_bounds__(kVecThreads)
    k_dcg_update(int64_t o, int64_t n, int which, const double* __restrict__ q, const double* __restrict__ dinv,
                 double* __restrict__ x, double* __restrict__ r, double* sc, double* partials,
                 unsigned int* ticket, Peers P, unsigned long long ep_wait, unsigned long long ep_red) {
  __shared__ double s_red[kVecThreads / 32];
  if (sc[SC_DONE] != 0.0) return;
  double pq[1];
  if (!reduce_wait<1>(P, 0, ep_wait, pq, sc)) return;
  const double* __restrict__ p = p_buf(P, P.rank, which);
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
    reduce_publish<2>(P, 1, ep_red, tot);
    if (threadIdx.x == 0) {
      sc[SC_PQ] = pq[0];
      sc[SC_ALPHA] = alpha;
    }
  }
}

// beta = rho' / rho ; p' = D^-1 r + beta p (other buffer) + peer stores of the halo entries ; scalar state
__global__ void __launch_bounds__(kVecThreads)
    k_dcg_direction(int64_t o, int64_t n, int which, const double* __restrict__ r,
                    const double* __restrict__ dinv, double* sc, unsigned int* ticket, Peers P, Halo H,
                    unsigned long long ep_wait, unsigned long long ep_halo) {
  if (sc[SC_DONE] != 0.0) return;
  double t2[2];
  if (!reduce_wait<2>(P, 1, ep_wait, t2, sc)) return;
  const double rho_prev = sc[SC_RHO];
  const double beta = t2[1] / rho_prev;
  const double* __restrict__ p = p_buf(P, P.rank, which);
  double* __restrict__ pn = p_buf(P, P.rank, which ^ 1);
  for (int64_t k = blockIdx.x * (int64_t)kVecThreads + threadIdx.x; k < n; k += (int64_t)gridDim.x * kVecThreads) {
    const int64_t i = o + k;
    pn[i] = fma(beta, p[i], dinv[i] * r[i]);
  }
  halo_send(P, H, which ^ 1, [&](int64_t i) { return fma(beta, p[i], dinv[i] * r[i]); });
  if (last_cta(ticket)) {
    if (threadIdx.x < 32) release_halo_flags(P, H, ep_halo);
    if (threadIdx.x == 0) {
      sc[SC_RHO_PREV] = rho_prev;
      sc[SC_RHO] = t2[1];
      sc[SC_RR] = t2[0];
      sc[SC_BETA] = beta;
      sc[SC_ITERS] += 1.0;
      if (!isfinite(t2[0])) sc[SC_DONE] = 2.0;
      else if (sqrt(t2[0]) < sc[SC_TOL]) sc[SC_DONE] = 1.0;
    }
  }
}

// barrier at the end of a solve: nobody starts the next solve (whose first kernel overwrites reduction slots
// and p halos on its peers) before every rank has finished the last kernel of this one
__global__ void k_dcg_finish(double* sc, Peers P, unsigned long long ep) {
  if (sc[SC_DONE] == 4.0) return;
  if (threadIdx.x < P.world) st_release_sys(red_flag(P, threadIdx.x, 0) + P.rank, ep);
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long t0 = global_ns();
    const unsigned long long* fl = red_flag(P, P.rank, 0);
    bool ok = true;
    for (int r = 0; r < P.world && ok; ++r) ok = spin_until(fl + r, ep, t0, P.timeout_ns);
    if (!ok) sc[SC_DONE] = 4.0;
  }
}

template <int DPN>
int launch_dcg_spmv(const Sell& A, int64_t ts_lo, int64_t ts_hi, int64_t ti_lo, int64_t ti_hi, int which,
                    double* q, double* sc, double* partials, unsigned int* ticket, const Peers& P,
                    const Halo& H, unsigned long long ep_halo, unsigned long long ep_red, cudaStream_t st) {
  static int g = 0;
  if (!g) g = resident_ctas(k_dcg_spmv<DPN>, kSellWarps * 32);
  const int64_t want = (ts_hi - ts_lo + kSellWarps - 1) / kSellWarps;
  k_dcg_spmv<DPN><<<(int)(want < g ? (want > 0 ? want : 1) : g), kSellWarps * 32, 0, st>>>(
      A, ts_lo, ts_hi, ti_lo, ti_hi, which, q, sc, partials, ticket, P, H, ep_halo, ep_red);
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
  c->vec = pad32(vec_doubles);
  const size_t bytes = (size_t)HEADER_BYTES + 2 * (size_t)c->vec * sizeof(double);
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

extern "C" int tfem_dcg_solve(void* comm, const tfem_sell_t* a, int64_t row_lo, int64_t n_owned,
                              int64_t interior_lo, int64_t interior_hi, int n_sends,
                              const tfem_halo_send_t* sends, int n_recv, const int32_t* recv_peers,
                              const double* dinv, const double* b, double* x, double* work, double rtol,
                              double atol, int64_t maxiter, int check_every, double timeout_s,
                              double* info, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  Comm* c = static_cast<Comm*>(comm);
  TFEM_REQUIRE(c && c->connected, "dcg_solve: communicator missing or not connected");
  TFEM_REQUIRE(!c->broken, "dcg_solve: communicator is out of step after a failed solve; create a new one");
  int rc0 = check_sell(a);
  if (rc0 != TFEM_OK) return rc0;
  TFEM_REQUIRE(dinv && b && x && work && info, "dcg_solve: null pointer");
  const int64_t n_local = a->n_rows;
  TFEM_REQUIRE(n_owned > 0 && row_lo >= 0 && row_lo + n_owned <= n_local, "dcg_solve: bad row range");
  TFEM_REQUIRE(n_local <= c->vec, "dcg_solve: communicator vectors are too short for this matrix");
  TFEM_REQUIRE(n_sends >= 0 && n_sends <= kMaxNbr && n_recv >= 0 && n_recv <= kMaxNbr && (n_sends == 0 || sends) &&
                   (n_recv == 0 || recv_peers), "dcg_solve: bad halo plan");
  if (maxiter <= 0) maxiter = 10 * n_local * c->world;
  if (check_every <= 0) check_every = 32;
  if (!(timeout_s > 0.0)) timeout_s = 20.0;

  Peers P;
  P.rank = c->rank;
  P.world = c->world;
  P.vec = c->vec;
  for (int r = 0; r < kMaxRanks; ++r) P.base[r] = c->base[r];
  P.timeout_ns = (unsigned long long)(timeout_s * 1e9);
  Halo H;
  memset(&H, 0, sizeof(H));
  H.n_send = n_sends;
  for (int s = 0; s < n_sends; ++s) {
    TFEM_REQUIRE(sends[s].peer >= 0 && sends[s].peer < c->world && sends[s].peer != c->rank, "dcg_solve: bad peer");
    H.send_peer[s] = sends[s].peer;
    H.send_count[s] = sends[s].count;
    H.send_src[s] = sends[s].src_idx;
    H.send_dst[s] = sends[s].dst_idx;
    H.src0[s] = sends[s].src_start;
    H.dst0[s] = sends[s].dst_start;
  }
  H.n_recv = n_recv;
  for (int s = 0; s < n_recv; ++s) {
    TFEM_REQUIRE(recv_peers[s] >= 0 && recv_peers[s] < c->world, "dcg_solve: bad peer");
    H.recv_peer[s] = recv_peers[s];
  }

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
  const unsigned long long E0 = c->epoch;
  TFEM_CUDA(cudaMemsetAsync(w.sc, 0, (SC_COUNT + kMaxPartials + 32) * sizeof(double), st));
  k_dcg_init<<<vg, kVecThreads, 0, st>>>(row_lo, n_owned, b, dinv, w.r, x, w.partials, w.ticket, P, H, E0 + 1, E0 + 1);
  TFEM_LAUNCH_CHECK();
  k_dcg_scalars_init<<<1, 32, 0, st>>>(w.sc, P, E0 + 1, rtol, atol);
  TFEM_LAUNCH_CHECK();
  double launches = 2, sc_host[SC_COUNT];
  int64_t issued = 0;
  int rc = TFEM_OK;
  while (true) {
    TFEM_CUDA(cudaMemcpyAsync(sc_host, w.sc, sizeof(sc_host), cudaMemcpyDeviceToHost, st));
    TFEM_CUDA(cudaStreamSynchronize(st));
    if (sc_host[SC_DONE] != 0.0 || issued >= maxiter) break;
    const int64_t batch = maxiter - issued < check_every ? maxiter - issued : check_every;
    for (int64_t k = 0; k < batch; ++k) {
      const unsigned long long it = (unsigned long long)(issued + k);
      const int which = (int)(it & 1);
      if (A.dpn == 3)
        rc = launch_dcg_spmv<3>(A, ts_lo, ts_hi, ti_lo, ti_hi, which, w.q, w.sc, w.partials, w.ticket, P, H, E0 + 1 + it, E0 + 1 + it, st);
      else if (A.dpn == 2)
        rc = launch_dcg_spmv<2>(A, ts_lo, ts_hi, ti_lo, ti_hi, which, w.q, w.sc, w.partials, w.ticket, P, H, E0 + 1 + it, E0 + 1 + it, st);
      else
        rc = launch_dcg_spmv<0>(A, ts_lo, ts_hi, ti_lo, ti_hi, which, w.q, w.sc, w.partials, w.ticket, P, H, E0 + 1 + it, E0 + 1 + it, st);
      if (rc != TFEM_OK) { c->broken = true; return rc; }
      k_dcg_update<<<vg, kVecThreads, 0, st>>>(row_lo, n_owned, which, w.q, dinv, x, w.r, w.sc, w.partials, w.ticket, P,
                                              E0 + 1 + it, E0 + 2 + it);
      k_dcg_direction<<<vg, kVecThreads, 0, st>>>(row_lo, n_owned, which, w.r, dinv, w.sc, w.ticket, P, H, E0 + 2 + it,
                                                 E0 + 2 + it);
    }
    rc = check_cuda(cudaGetLastError(), "dcg launch");
    if (rc != TFEM_OK) { c->broken = true; return rc; }
    issued += batch;
    launches += 3.0 * batch;
  }
  c->epoch = E0 + (unsigned long long)issued + 2;
  if (sc_host[SC_DONE] != 4.0) {
    const double done = sc_host[SC_DONE];
    k_dcg_finish<<<1, 32, 0, st>>>(w.sc, P, c->epoch);
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
  info[7] = 0.0;
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
