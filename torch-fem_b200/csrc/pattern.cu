// K0 — sparsity pattern from the node graph (replaces src/torchfem/base.py:78-118).
//
// The reference sorts n_elem*(nn*dpn)^2 packed 64-bit keys. Here the pattern is built on NODES:
//   1. node -> element-slot incidence (integer counting sort; ascending slot ids per node),
//   2. per node (one warp): the sorted distinct set of all nodes of its incident elements,
//   3. expansion to dpn x dpn scalar blocks (rows sorted by column, diagonal always present; a node that
//      no element references keeps the reference's lone diagonal entry per DOF, base.py:89-91),
//   4. the element-slot -> block permutation `src` used by the deterministic assembly.
// Only integer atomics are used and every list they fill is sorted afterwards, so the output is
// deterministic and bit-identical to the reference's glob_idx / k_map / diag_map.
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>

#include "common.cuh"

namespace tfem {
namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kCandCap = 2048;  // candidates (incident elements * nn) one warp can hold in shared memory

__global__ void k_count_incidence(int64_t n_slots, int64_t n_nod, const int64_t* __restrict__ elements,
                                  int32_t* __restrict__ inc_ptr, int32_t* __restrict__ bad) {
  int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  int64_t node = elements[s];
  if (node < 0 || node >= n_nod) {
    *bad = 1;
    return;
  }
  atomicAdd(&inc_ptr[node + 1], 1);
}

__global__ void k_fill_incidence(int64_t n_slots, int64_t n_nod, const int64_t* __restrict__ elements,
                                 const int32_t* __restrict__ inc_ptr, int32_t* __restrict__ cursor,
                                 int32_t* __restrict__ inc_list) {
  int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  int64_t node = elements[s];
  if (node < 0 || node >= n_nod) return;
  int pos = atomicAdd(&cursor[node], 1);
  inc_list[inc_ptr[node] + pos] = (int32_t)s;
}

// ascending slot ids per node (the atomic cursor above fills them in arbitrary order)
__global__ void k_sort_incidence(int64_t n_nod, const int32_t* __restrict__ inc_ptr,
                                 int32_t* __restrict__ inc_list) {
  int64_t node = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (node >= n_nod) return;
  int b = inc_ptr[node], e = inc_ptr[node + 1];
  for (int i = b + 1; i < e; ++i) {
    int32_t v = inc_list[i];
    int j = i - 1;
    while (j >= b && inc_list[j] > v) {
      inc_list[j + 1] = inc_list[j];
      --j;
    }
    inc_list[j + 1] = v;
  }
}

// Loads the candidate neighbour nodes of `node` (all nodes of all incident elements, slot-major) into
// shared memory and flags the first occurrence of every distinct value. Returns the candidate count.
__device__ __forceinline__ int load_candidates(int64_t node, int nn, const int64_t* __restrict__ elements,
                                               const int32_t* __restrict__ inc_ptr,
                                               const int32_t* __restrict__ inc_list, int32_t* cand,
                                               uint8_t* first, int lane) {
  const int b = inc_ptr[node];
  const int n_inc = inc_ptr[node + 1] - b;
  const int L = n_inc * nn;
  for (int j = lane; j < L; j += 32) {
    int s = inc_list[b + j / nn];
    int e = s / nn;
    cand[j] = (int32_t)elements[(int64_t)e * nn + (j % nn)];
  }
  __syncwarp();
  for (int j = lane; j < L; j += 32) {
    int v = cand[j];
    bool f = true;
    for (int t = 0; t < j; ++t)
      if (cand[t] == v) {
        f = false;
        break;
      }
    first[j] = f ? 1 : 0;
  }
  __syncwarp();
  return L;
}

__global__ void __launch_bounds__(kWarpsPerCta * 32)
    k_node_count(int64_t n_nod, int nn, int dpn, const int64_t* __restrict__ elements,
                 const int32_t* __restrict__ inc_ptr, const int32_t* __restrict__ inc_list,
                 int32_t* __restrict__ blk_cnt, unsigned long long* __restrict__ totals,
                 int32_t* __restrict__ bad) {
  __shared__ int32_t s_cand[kWarpsPerCta][kCandCap];
  __shared__ uint8_t s_first[kWarpsPerCta][kCandCap];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t node = blockIdx.x * (int64_t)kWarpsPerCta + warp;
  if (node >= n_nod) return;
  const int n_inc = inc_ptr[node + 1] - inc_ptr[node];
  if (n_inc * nn > kCandCap) {
    if (lane == 0) *bad = 2;
    return;
  }
  const int L = load_candidates(node, nn, elements, inc_ptr, inc_list, s_cand[warp], s_first[warp], lane);
  int c = 0;
  for (int j = lane; j < L; j += 32) c += s_first[warp][j];
  c = warp_sum_int(c);
  if (lane == 0) {
    blk_cnt[node] = c;
    atomicAdd(&totals[0], (unsigned long long)c);
    atomicAdd(&totals[1], (unsigned long long)(c > 0 ? (long long)c * dpn * dpn : dpn));
    atomicMax(&totals[2], (unsigned long long)c);
    atomicMax(&totals[3], (unsigned long long)n_inc);
  }
}

struct BlkToI64 {
  __host__ __device__ int64_t operator()(int32_t c) const { return (int64_t)c; }
};
struct NodeNnz {
  int dpn;
  __host__ __device__ int64_t operator()(int32_t c) const {
    return c > 0 ? (int64_t)c * dpn * dpn : (int64_t)dpn;
  }
};

// One warp per node: neighbour list, scalar rows, diagonal positions and the assembly permutation.
__global__ void __launch_bounds__(kWarpsPerCta * 32)
    k_node_fill(int64_t n_nod, int nn, int dpn, const int64_t* __restrict__ elements,
                const int32_t* __restrict__ inc_ptr, const int32_t* __restrict__ inc_list,
                const int64_t* __restrict__ node_ptr, const int64_t* __restrict__ node_base,
                int32_t* __restrict__ adj, int64_t* __restrict__ indptr, int32_t* __restrict__ indices,
                int32_t* __restrict__ diag_map, int64_t* __restrict__ src_ptr, int32_t* __restrict__ src) {
  __shared__ int32_t s_cand[kWarpsPerCta][kCandCap];
  __shared__ uint8_t s_first[kWarpsPerCta][kCandCap];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t node = blockIdx.x * (int64_t)kWarpsPerCta + warp;
  if (node >= n_nod) return;
  int32_t* cand = s_cand[warp];
  uint8_t* first = s_first[warp];
  const int n_inc = inc_ptr[node + 1] - inc_ptr[node];
  if (n_inc * nn > kCandCap) return;  // reported by phase 1
  const int L = load_candidates(node, nn, elements, inc_ptr, inc_list, cand, first, lane);
  const int64_t nb = node_ptr[node];
  const int cnt = (int)(node_ptr[node + 1] - nb);
  const int64_t base = node_base[node];
  const int64_t row0 = node * dpn;

  if (cnt == 0) {  // unreferenced node: lone diagonal per DOF (base.py:89-91)
    if (lane < dpn) {
      indptr[row0 + lane] = base + lane;
      indices[base + lane] = (int32_t)(row0 + lane);
      diag_map[row0 + lane] = (int32_t)(base + lane);
    }
    if (node == n_nod - 1 && lane == 0) indptr[n_nod * dpn] = base + dpn;
    return;
  }

  // sorted distinct neighbours: rank of a first occurrence = number of smaller first occurrences
  for (int j = lane; j < L; j += 32) {
    if (!first[j]) continue;
    int v = cand[j], rank = 0;
    for (int t = 0; t < L; ++t) rank += (first[t] && cand[t] < v) ? 1 : 0;
    adj[nb + rank] = v;
  }
  __syncwarp();

  // scalar CSR rows of this node: dpn rows, each cnt*dpn entries, column = dpn*neighbour + j
  const int rowlen = cnt * dpn;
  for (int i = 0; i < dpn; ++i) {
    const int64_t rp = base + (int64_t)i * rowlen;
    if (lane == 0) indptr[row0 + i] = rp;
    for (int t = lane; t < rowlen; t += 32) {
      int nbr = adj[nb + t / dpn];
      int col = nbr * dpn + (t % dpn);
      indices[rp + t] = col;
      if (col == (int32_t)(row0 + i)) diag_map[row0 + i] = (int32_t)(rp + t);
    }
  }
  if (node == n_nod - 1 && lane == 0) indptr[n_nod * dpn] = base + (int64_t)dpn * rowlen;

  // assembly permutation: contributions (slot s = e*nn + a, local column node b) of every block, in
  // candidate order (ascending e, then a, then b) == the reference's k.ravel() slot order
  const int64_t sbase = (int64_t)inc_ptr[node] * nn;
  int carry = 0;
  for (int p0 = 0; p0 < cnt; p0 += 32) {
    const int p = p0 + lane;
    int v = (p < cnt) ? adj[nb + p] : -1;
    int c = 0;
    if (p < cnt)
      for (int t = 0; t < L; ++t) c += (cand[t] == v) ? 1 : 0;
    int incl = warp_scan_incl(c, lane);
    int off = carry + incl - c;
    if (p < cnt) {
      src_ptr[nb + p] = sbase + off;
      int w = 0;
      for (int t = 0; t < L; ++t)
        if (cand[t] == v) {
          int s = inc_list[inc_ptr[node] + t / nn];
          src[sbase + off + w] = s * nn + (t % nn);
          ++w;
        }
    }
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
}

__global__ void k_set_src_end(const int64_t* node_ptr, int64_t n_nod, int64_t* src_ptr, int64_t total) {
  src_ptr[node_ptr[n_nod]] = total;
}

// reference-compatible k_map: one thread per (element, local row node a, local col node b)
__global__ void k_kmap(int64_t n_pairs, int nn, int dpn, const int64_t* __restrict__ elements,
                       const int64_t* __restrict__ node_ptr, const int32_t* __restrict__ adj,
                       const int64_t* __restrict__ indptr, int32_t* __restrict__ k_map) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_pairs) return;
  const int b = (int)(t % nn);
  const int a = (int)((t / nn) % nn);
  const int64_t e = t / ((int64_t)nn * nn);
  const int64_t na = elements[e * nn + a];
  const int32_t nbn = (int32_t)elements[e * nn + b];
  int64_t lo = node_ptr[na], hi = node_ptr[na + 1] - 1;
  const int64_t beg = lo;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (adj[mid] < nbn) lo = mid + 1; else hi = mid;
  }
  const int pos = (int)(lo - beg);
  const int nd = nn * dpn;
  for (int i = 0; i < dpn; ++i) {
    const int64_t rp = indptr[na * dpn + i] + (int64_t)pos * dpn;
    for (int j = 0; j < dpn; ++j)
      k_map[e * nd * nd + (int64_t)(a * dpn + i) * nd + (b * dpn + j)] = (int32_t)(rp + j);
  }
}

__global__ void k_coo_rows(int64_t n_rows, int64_t nnz, const int64_t* __restrict__ indptr,
                           int64_t* __restrict__ rows) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  int64_t lo = 0, hi = n_rows - 1;  // last row with indptr[row] <= p
  while (lo < hi) {
    int64_t mid = (lo + hi + 1) >> 1;
    if (indptr[mid] <= p) lo = mid; else hi = mid - 1;
  }
  rows[p] = lo;
}

inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

}  // namespace
}  // namespace tfem

using namespace tfem;

extern "C" int tfem_pattern_phase1(int64_t n_nod, int64_t n_elem, int nn, int dpn,
                                   const int64_t* elements, int32_t* inc_ptr, int32_t* inc_list,
                                   int32_t* blk_cnt, int64_t* totals, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(n_nod > 0 && n_elem >= 0 && nn > 0 && dpn > 0, "pattern: bad sizes");
  TFEM_REQUIRE(elements && inc_ptr && inc_list && blk_cnt && totals, "pattern: null pointer");
  const int64_t n_slots = n_elem * nn;
  if (n_slots * nn >= (int64_t)INT32_MAX || n_nod * dpn >= (int64_t)INT32_MAX) {
    set_last_error("capacity", "n_elem*nn*nn and n_dofs must be < 2^31 on one device (partition the mesh)");
    return TFEM_ERR_CAPACITY;
  }
  int32_t* bad = nullptr;
  TFEM_CUDA(malloc_async(&bad, sizeof(int32_t), st));
  TFEM_CUDA(cudaMemsetAsync(bad, 0, sizeof(int32_t), st));
  TFEM_CUDA(cudaMemsetAsync(inc_ptr, 0, (n_nod + 1) * sizeof(int32_t), st));
  TFEM_CUDA(cudaMemsetAsync(blk_cnt, 0, n_nod * sizeof(int32_t), st));
  TFEM_CUDA(cudaMemsetAsync(totals, 0, 4 * sizeof(int64_t), st));
  if (n_slots > 0) {
    k_count_incidence<<<grid_for(n_slots, 256), 256, 0, st>>>(n_slots, n_nod, elements, inc_ptr, bad);
    TFEM_LAUNCH_CHECK();
  }
  size_t tmp_bytes = 0;
  TFEM_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, inc_ptr, inc_ptr, (int)(n_nod + 1), st));
  void* tmp = nullptr;
  TFEM_CUDA(malloc_async(&tmp, tmp_bytes ? tmp_bytes : 16, st));
  TFEM_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, inc_ptr, inc_ptr, (int)(n_nod + 1), st));
  if (n_slots > 0) {
    k_fill_incidence<<<grid_for(n_slots, 256), 256, 0, st>>>(n_slots, n_nod, elements, inc_ptr, blk_cnt,
                                                             inc_list);
    TFEM_LAUNCH_CHECK();
  }
  k_sort_incidence<<<grid_for(n_nod, 128), 128, 0, st>>>(n_nod, inc_ptr, inc_list);
  TFEM_LAUNCH_CHECK();
  TFEM_CUDA(cudaMemsetAsync(blk_cnt, 0, n_nod * sizeof(int32_t), st));
  k_node_count<<<grid_for(n_nod, kWarpsPerCta), kWarpsPerCta * 32, 0, st>>>(
      n_nod, nn, dpn, elements, inc_ptr, inc_list, blk_cnt, (unsigned long long*)totals, bad);
  TFEM_LAUNCH_CHECK();
  int32_t bad_h = 0;
  TFEM_CUDA(cudaMemcpyAsync(&bad_h, bad, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  TFEM_CUDA(cudaStreamSynchronize(st));
  TFEM_CUDA(cudaFreeAsync(tmp, st));
  TFEM_CUDA(cudaFreeAsync(bad, st));
  if (bad_h == 1) {
    set_last_error("invalid argument", "element connectivity references a node outside [0, n_nod)");
    return TFEM_ERR_INVALID;
  }
  if (bad_h == 2) {
    set_last_error("capacity", "a node has more than 2048/nn incident elements");
    return TFEM_ERR_CAPACITY;
  }
  return TFEM_OK;
}

extern "C" int tfem_pattern_phase2(int64_t n_nod, int64_t n_elem, int nn, int dpn,
                                   const int64_t* elements, const int32_t* inc_ptr,
                                   const int32_t* inc_list, const int32_t* blk_cnt, int64_t* node_ptr,
                                   int32_t* adj, int64_t* indptr, int32_t* indices, int32_t* diag_map,
                                   int64_t* src_ptr, int32_t* src, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(elements && inc_ptr && inc_list && blk_cnt && node_ptr && adj && indptr && indices &&
                   diag_map && src_ptr && src, "pattern: null pointer");
  // exclusive scans over nodes: block offsets and scalar-entry offsets (n_nod+1 outputs each)
  int64_t* node_base = nullptr;
  TFEM_CUDA(malloc_async(&node_base, (n_nod + 1) * sizeof(int64_t), st));
  auto it_blk = thrust::make_transform_iterator(blk_cnt, BlkToI64());
  auto it_nnz = thrust::make_transform_iterator(blk_cnt, NodeNnz{dpn});
  size_t b1 = 0, b2 = 0;
  TFEM_CUDA(cub::DeviceScan::InclusiveSum(nullptr, b1, it_blk, node_ptr + 1, (int)n_nod, st));
  TFEM_CUDA(cub::DeviceScan::InclusiveSum(nullptr, b2, it_nnz, node_base + 1, (int)n_nod, st));
  size_t tmp_bytes = b1 > b2 ? b1 : b2;
  void* tmp = nullptr;
  TFEM_CUDA(malloc_async(&tmp, tmp_bytes ? tmp_bytes : 16, st));
  TFEM_CUDA(cudaMemsetAsync(node_ptr, 0, sizeof(int64_t), st));
  TFEM_CUDA(cudaMemsetAsync(node_base, 0, sizeof(int64_t), st));
  TFEM_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, it_blk, node_ptr + 1, (int)n_nod, st));
  TFEM_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, it_nnz, node_base + 1, (int)n_nod, st));
  k_node_fill<<<grid_for(n_nod, kWarpsPerCta), kWarpsPerCta * 32, 0, st>>>(
      n_nod, nn, dpn, elements, inc_ptr, inc_list, node_ptr, node_base, adj, indptr, indices, diag_map,
      src_ptr, src);
  TFEM_LAUNCH_CHECK();
  // src_ptr[nnzb] = total number of contributions; nnzb is node_ptr[n_nod] (device) -> tiny kernel
  // reads it there instead of a host round trip
  k_set_src_end<<<1, 1, 0, st>>>(node_ptr, n_nod, src_ptr, n_elem * (int64_t)nn * nn);
  TFEM_LAUNCH_CHECK();
  TFEM_CUDA(cudaFreeAsync(tmp, st));
  TFEM_CUDA(cudaFreeAsync(node_base, st));
  return TFEM_OK;
}

extern "C" int tfem_pattern_k_map(int64_t n_nod, int64_t n_elem, int nn, int dpn,
                                  const int64_t* elements, const int64_t* node_ptr, const int32_t* adj,
                                  const int64_t* indptr, int32_t* k_map, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  (void)n_nod;
  TFEM_REQUIRE(elements && node_ptr && adj && indptr && k_map, "k_map: null pointer");
  const int64_t n_pairs = n_elem * nn * nn;
  if (n_pairs == 0) return TFEM_OK;
  k_kmap<<<grid_for(n_pairs, 256), 256, 0, st>>>(n_pairs, nn, dpn, elements, node_ptr, adj, indptr, k_map);
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}

extern "C" int tfem_pattern_coo_rows(int64_t n_dofs, const int64_t* indptr, int64_t* rows,
                                     void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)cudaGetLastError();  // a stale non-sticky error of another library (NCCL pointer queries) is not ours
  TFEM_REQUIRE(indptr && rows && n_dofs > 0, "coo_rows: bad arguments");
  int64_t nnz = 0;
  TFEM_CUDA(cudaMemcpyAsync(&nnz, indptr + n_dofs, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  TFEM_CUDA(cudaStreamSynchronize(st));
  if (nnz == 0) return TFEM_OK;
  k_coo_rows<<<grid_for(nnz, 256), 256, 0, st>>>(n_dofs, nnz, indptr, rows);
  TFEM_LAUNCH_CHECK();
  return TFEM_OK;
}
