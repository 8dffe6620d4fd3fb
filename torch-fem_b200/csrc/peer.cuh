// Peer-memory primitives shared by the multi-GPU solvers (dcg.cu: fused Jacobi-PCG; damg.cu: distributed AMG-PCG).
//
// Every rank owns one communication buffer (cudaMalloc + CUDA IPC) and maps the buffers of all peers; the same
// offsets mean the same thing on every rank (a symmetric heap). Layout:
//   [header 8 KB | heap: `heap_doubles` doubles carved by the host into vectors at equal offsets on all ranks]
// Header:
//   halo flags   uint64 [kMaxChannels][kMaxRanks]   epoch of the last halo rank s delivered on that channel
//   LL slots     uint64 [2 sets][kMaxRanks][8]      all-reduce payload, NCCL-LL style: every 8-byte word carries
//                                                   32 bits of data and the 32-bit epoch, so data and "flag" arrive
//                                                   in ONE single-copy-atomic NVLink store: no fence, no second trip
// All waits are bounded by `timeout_ns` (then the solve ends with TFEM_ERR_COMM instead of hanging the GPU).
//
// Optional in-kernel trace (`tfem_comm_set_trace`): kernels stamp %globaltimer at fixed points of an iteration into
// trace[(it - it0) * kTraceSlots + slot]; this is the profile of the waits nsys would give (nsys is not in the
// image, and ncu serialises kernels, which hides exactly the cross-GPU waits).
#pragma once
#include "common.cuh"

namespace tfem {
namespace {

constexpr int kMaxRanks = 16;
constexpr int kMaxNbr = TFEM_MAX_NEIGHBOURS;
constexpr int kMaxChannels = 16;
constexpr int kLLWords = 8;                       // 4 doubles per (set, rank)
constexpr int64_t OFF_HALO_FLAG = 0;              // uint64 [kMaxChannels][kMaxRanks] = 2 KB
constexpr int64_t OFF_LL = 2048;                  // uint64 [2][kMaxRanks][kLLWords]  = 2 KB
constexpr int64_t OFF_BARRIER = 4096;             // uint64 [kMaxRanks]
constexpr int64_t HEADER_BYTES = 8192;
constexpr int kTraceSlots = TFEM_TRACE_SLOTS;

struct Peers {
  int rank, world;
  char* base[kMaxRanks];
  unsigned long long timeout_ns;
  unsigned long long* trace;  // device, or nullptr
  long long trace_it0;
  int trace_n;
};

struct Halo {
  int n_send;
  int send_peer[kMaxNbr];
  int64_t send_count[kMaxNbr];
  const int32_t* send_src[kMaxNbr];  // local indices, or nullptr: contiguous from src0
  const int32_t* send_dst[kMaxNbr];  // indices in the peer's local numbering, or nullptr: contiguous from dst0
  int64_t src0[kMaxNbr], dst0[kMaxNbr];
  int64_t send_total;
  int n_recv;
  int recv_peer[kMaxNbr];
};

__device__ __forceinline__ double* heap(const Peers& P, int r) {
  return reinterpret_cast<double*>(P.base[r] + HEADER_BYTES);
}
__device__ __forceinline__ unsigned long long* halo_flag(const Peers& P, int r, int channel) {
  return reinterpret_cast<unsigned long long*>(P.base[r] + OFF_HALO_FLAG) + channel * kMaxRanks;
}
__device__ __forceinline__ unsigned long long* ll_slot(const Peers& P, int r, int set) {
  return reinterpret_cast<unsigned long long*>(P.base[r] + OFF_LL) + set * kMaxRanks * kLLWords;
}
__device__ __forceinline__ unsigned long long* barrier_flag(const Peers& P, int r) {
  return reinterpret_cast<unsigned long long*>(P.base[r] + OFF_BARRIER);
}

__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void trace_mark(const Peers& P, long long it, int slot) {
  if (P.trace && it >= P.trace_it0 && it < P.trace_it0 + P.trace_n)
    P.trace[(it - P.trace_it0) * kTraceSlots + slot] = global_ns();
}
__device__ __forceinline__ void trace_max(const Peers& P, long long it, int slot, unsigned long long v) {
  if (P.trace && it >= P.trace_it0 && it < P.trace_it0 + P.trace_n)
    atomicMax(P.trace + (it - P.trace_it0) * kTraceSlots + slot, v);
}

// one thread: spin (relaxed loads: no fence per poll) until *flag >= epoch, then ONE acquire fence; false on timeout
__device__ __forceinline__ bool spin_until(const unsigned long long* flag, unsigned long long epoch,
                                           unsigned long long t0, unsigned long long timeout_ns) {
  int polls = 0;
  while (ld_relaxed_sys(flag) < epoch) {
    if ((++polls & 63) == 0 && global_ns() - t0 > timeout_ns) return false;
    __nanosleep(20);
  }
  return true;
}

// ---- all-reduce (sum) of K <= 4 doubles over the ranks, LL protocol.
// publish: lanes 0..world-1 of one warp store this rank's K values into slot [my rank] of rank `lane`.
template <int K>
__device__ __forceinline__ void ll_publish(const Peers& P, int set, unsigned long long epoch, const double (&v)[K]) {
  const int lane = threadIdx.x & 31;
  if (lane < P.world) {
    unsigned long long* dst = ll_slot(P, lane, set) + P.rank * kLLWords;
    const unsigned long long tag = (epoch & 0xffffffffull) << 32;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const unsigned long long bits = (unsigned long long)__double_as_longlong(v[j]);
      st_relaxed_sys(dst + 2 * j, tag | (bits & 0xffffffffull));
      st_relaxed_sys(dst + 2 * j + 1, tag | (bits >> 32));
    }
  }
}

// wait: ONE warp of the CTA polls (lane r = rank r's slot), sums in rank order (bit-identical on every rank) and
// broadcasts through shared memory. Returns false in all threads on timeout.
template <int K>
__device__ __forceinline__ bool ll_wait_sum(const Peers& P, int set, unsigned long long epoch, double (&out)[K],
                                            double* s_out /* shared double[4] */, int* s_ok /* shared int */) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const unsigned long long want = epoch & 0xffffffffull;
    const unsigned long long* src = ll_slot(P, P.rank, set) + lane * kLLWords;
    double mine[K];
    bool ok = true;
    if (lane < P.world) {
      const unsigned long long t0 = global_ns();
#pragma unroll
      for (int j = 0; j < K; ++j) {
        unsigned long long lo, hi;
        int polls = 0;
        while (true) {
          lo = ld_relaxed_sys(src + 2 * j);
          hi = ld_relaxed_sys(src + 2 * j + 1);
          if ((lo >> 32) == want && (hi >> 32) == want) break;
          if ((++polls & 63) == 0 && global_ns() - t0 > P.timeout_ns) { ok = false; break; }
          __nanosleep(20);
        }
        mine[j] = __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
      }
    } else {
#pragma unroll
      for (int j = 0; j < K; ++j) mine[j] = 0.0;
    }
    ok = __all_sync(0xffffffffu, ok);
#pragma unroll
    for (int j = 0; j < K; ++j) {
      double s = 0.0;
      for (int r = 0; r < P.world; ++r) s += __shfl_sync(0xffffffffu, mine[j], r);  // rank order
      if (lane == 0) s_out[j] = s;
    }
    if (lane == 0) *s_ok = ok ? 1 : 0;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < K; ++j) out[j] = s_out[j];
  const bool ok = *s_ok != 0;
  __syncthreads();
  return ok;
}

// ---- halo exchange by peer stores.
// The CTAs [0, n_halo_ctas) of a kernel store value(i) for every entry a neighbour needs into that neighbour's
// vector at heap offset `vec_off`; each of them then fences ONCE (thread 0, after the CTA barrier) and takes a
// ticket; the last one releases the channel's flag on every neighbour. Doing this in the FIRST CTAs of the grid
// lets the exchange overlap the rest of the kernel.
__host__ __device__ inline int halo_ctas(int64_t send_total, int threads) {
  const int64_t per_cta = (int64_t)threads * 4;
  const int64_t c = (send_total + per_cta - 1) / per_cta;
  return (int)(c < 1 ? 1 : c);
}

template <typename F>
__device__ __forceinline__ bool halo_send_and_release(const Peers& P, const Halo& H, int64_t vec_off, int channel,
                                                      unsigned long long epoch, int n_halo_ctas,
                                                      unsigned int* halo_ticket, F value) {
  __shared__ bool s_last_halo;
  for (int s = 0; s < H.n_send; ++s) {
    double* dst = heap(P, H.send_peer[s]) + vec_off;
    const int32_t* si = H.send_src[s];
    const int32_t* di = H.send_dst[s];
    const int64_t stride = (int64_t)n_halo_ctas * blockDim.x, cnt = H.send_count[s];
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    // four entries in flight per thread: the loop is a chain of dependent memory latencies otherwise
    for (; k + 3 * stride < cnt; k += 4 * stride) {
      int64_t i[4], d[4];
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t ku = k + u * stride;
        i[u] = si ? (int64_t)si[ku] : H.src0[s] + ku;
        d[u] = di ? (int64_t)di[ku] : H.dst0[s] + ku;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = value(i[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) dst[d[u]] = v[u];
    }
    for (; k < cnt; k += stride) {
      const int64_t i = si ? (int64_t)si[k] : H.src0[s] + k;
      const int64_t d = di ? (int64_t)di[k] : H.dst0[s] + k;
      dst[d] = value(i);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();  // cumulative: the CTA's peer stores (ordered before by the barrier) become visible first
    const unsigned t = atomicAdd(halo_ticket, 1u);
    s_last_halo = (t == (unsigned)n_halo_ctas - 1);
    if (s_last_halo) *halo_ticket = 0u;
  }
  __syncthreads();
  if (s_last_halo && threadIdx.x < H.n_send) {
    __threadfence_system();
    st_release_sys(halo_flag(P, H.send_peer[threadIdx.x], channel) + P.rank, epoch);
  }
  return s_last_halo;  // true in every thread of the CTA that released the flags
}

// one thread per neighbour waits for the channel's flags; returns false on timeout (valid in the calling thread)
__device__ __forceinline__ bool halo_wait_thread(const Peers& P, const Halo& H, int channel,
                                                 unsigned long long epoch) {
  const unsigned long long t0 = global_ns();
  const unsigned long long* fl = halo_flag(P, P.rank, channel);
  bool ok = true;
  for (int s = 0; s < H.n_recv && ok; ++s) ok = spin_until(fl + H.recv_peer[s], epoch, t0, P.timeout_ns);
  if (ok) fence_acq_rel_sys();
  return ok;
}

struct Comm {
  int rank = 0, world = 1;
  int64_t heap_doubles = 0;
  char* base[kMaxRanks] = {};
  bool opened[kMaxRanks] = {};
  unsigned long long epoch = 0;          // dcg: reduction / halo epochs
  unsigned long long chan_epoch[kMaxChannels] = {};
  unsigned long long barrier_epoch = 0;
  bool broken = false;
  bool connected = false;
  unsigned long long* trace = nullptr;   // device buffer owned by the caller
  long long trace_it0 = 0;
  int trace_n = 0;
  int time_spmv = 0;
};

inline Peers make_peers(const Comm* c, double timeout_s) {
  Peers P;
  P.rank = c->rank;
  P.world = c->world;
  for (int r = 0; r < kMaxRanks; ++r) P.base[r] = c->base[r];
  P.timeout_ns = (unsigned long long)(timeout_s * 1e9);
  P.trace = c->trace;
  P.trace_it0 = c->trace_it0;
  P.trace_n = c->trace_n;
  return P;
}

inline int fill_halo(Halo& H, const Comm* c, int n_sends, const tfem_halo_send_t* sends, int n_recv,
                     const int32_t* recv_peers) {
  memset(&H, 0, sizeof(H));
  TFEM_REQUIRE(n_sends >= 0 && n_sends <= kMaxNbr && n_recv >= 0 && n_recv <= kMaxNbr && (n_sends == 0 || sends) &&
                   (n_recv == 0 || recv_peers), "bad halo plan");
  H.n_send = n_sends;
  for (int s = 0; s < n_sends; ++s) {
    TFEM_REQUIRE(sends[s].peer >= 0 && sends[s].peer < c->world && sends[s].peer != c->rank, "halo plan: bad peer");
    H.send_peer[s] = sends[s].peer;
    H.send_count[s] = sends[s].count;
    H.send_src[s] = sends[s].src_idx;
    H.send_dst[s] = sends[s].dst_idx;
    H.src0[s] = sends[s].src_start;
    H.dst0[s] = sends[s].dst_start;
    H.send_total += sends[s].count;
  }
  H.n_recv = n_recv;
  for (int s = 0; s < n_recv; ++s) {
    TFEM_REQUIRE(recv_peers[s] >= 0 && recv_peers[s] < c->world, "halo plan: bad peer");
    H.recv_peer[s] = recv_peers[s];
  }
  return TFEM_OK;
}

}  // namespace
}  // namespace tfem
