// K5 — CSR SpMV device code shared by tfem_spmv and the Krylov drivers.
//
// Layout: fp64 values, int32 column indices, int64 row offsets (algorithmic bytes 12*nnz + 20*n_rows).
// FEM rows are short (24..81 entries for Hexa1, <=243 for Hexa2), so a row-per-warp kernel would waste
// lanes and load values with unaligned 8-byte accesses. Instead the nonzero STREAM is cut into chunks of
// TFEM_SPMV_CHUNK entries; one warp owns a chunk = the rows that start inside it:
//   phase A  streams cols/vals of the chunk with 16-byte-aligned 128-bit loads (L1 no-allocate, they are
//            read once), gathers x through L1/L2 and parks the products in shared memory;
//   phase B  reduces each row from shared memory with G lanes per row (fixed order -> deterministic).
// The grid is persistent (a multiple of the SM count); CTA b takes chunks b, b+grid, ... so that all SMs
// walk the same band of the matrix and the gathered x window stays L2-resident.
#pragma once
#include "common.cuh"

namespace tfem {

constexpr int kSpmvWarps = 4;
constexpr int kSpmvCap = 1024;  // products one warp can park (chunk + longest row + alignment slack)

__global__ void k_spmv_plan(int64_t n_rows, int64_t n_chunks, const int64_t* __restrict__ indptr,
                            int32_t* __restrict__ chunk_rows);

// One warp processes chunk `c`. Returns this lane's contribution to dot(w, y) over the chunk's rows
// (0 if w == nullptr).
template <int G>
__device__ __forceinline__ double spmv_chunk(int64_t c, const int64_t* __restrict__ indptr,
                                             const int32_t* __restrict__ cols,
                                             const double* __restrict__ vals,
                                             const int32_t* __restrict__ chunk_rows,
                                             const double* __restrict__ x, double* __restrict__ y,
                                             const double* __restrict__ w, double* prod, int lane) {
  const int r0 = chunk_rows[c], r1 = chunk_rows[c + 1];
  if (r1 <= r0) return 0.0;
  const int64_t p0 = indptr[r0], p1 = indptr[r1];
  const int64_t pa = p0 & ~(int64_t)3;
  double dot = 0.0;

  if (p1 - pa <= kSpmvCap) {
    // ---- phase A: stream the chunk, park products
    const int4* c4 = reinterpret_cast<const int4*>(cols);
    const double2* v2 = reinterpret_cast<const double2*>(vals);
#pragma unroll 1
    for (int64_t b0 = pa; b0 < p1; b0 += 512) {
      int4 ci[4];
      double2 va[4], vb[4];
      bool full[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t base = b0 + u * 128 + lane * 4;
        full[u] = base + 4 <= p1;
        if (full[u]) {
          ci[u] = ldg_stream_int4(c4 + (base >> 2));
          va[u] = ldg_stream_double2(v2 + (base >> 1));
          vb[u] = ldg_stream_double2(v2 + (base >> 1) + 1);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t base = b0 + u * 128 + lane * 4;
        double* dst = prod + (base - pa);
        if (full[u]) {
          double2 o0, o1;
          o0.x = va[u].x * __ldg(x + ci[u].x);
          o0.y = va[u].y * __ldg(x + ci[u].y);
          o1.x = vb[u].x * __ldg(x + ci[u].z);
          o1.y = vb[u].y * __ldg(x + ci[u].w);
          reinterpret_cast<double2*>(dst)[0] = o0;
          reinterpret_cast<double2*>(dst)[1] = o1;
        } else {
          for (int k = 0; k < 4; ++k)
            if (base + k < p1) dst[k] = ldg_stream_double(vals + base + k) * __ldg(x + ldg_stream_int(cols + base + k));
        }
      }
    }
    __syncwarp();
    // ---- phase B: G lanes per row
    constexpr int R = 32 / G;
    const int g = lane % G, sub = lane / G;
    const int nrows = r1 - r0;
    for (int rb = 0; rb < nrows; rb += R) {
      const int rr = rb + sub;
      double acc = 0.0;
      if (rr < nrows) {
        const int s = (int)(indptr[r0 + rr] - pa), e = (int)(indptr[r0 + rr + 1] - pa);
        for (int k = s + g; k < e; k += G) acc += prod[k];
      }
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (rr < nrows && g == 0) {
        y[r0 + rr] = acc;
        if (w) dot += acc * w[r0 + rr];
      }
    }
    __syncwarp();
  } else {
    // ---- long rows (do not fit the parking area): one row at a time, whole warp, fixed order
    for (int r = r0; r < r1; ++r) {
      const int64_t s = indptr[r], e = indptr[r + 1];
      double acc = 0.0;
      for (int64_t k = s + lane; k < e; k += 32) acc += vals[k] * __ldg(x + cols[k]);
      acc = warp_sum(acc);
      if (lane == 0) {
        y[r] = acc;
        if (w) dot += acc * w[r];
      }
    }
  }
  return dot;
}

// Fixed-order block reduction of one double per thread; result valid in thread 0.
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* red /*[THREADS/32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < THREADS / 32; ++i) s += red[i];
  return s;
}

// "last block done" finalisation: every CTA publishes its partial(s), the CTA that takes the last ticket
// sums all partials in index order (deterministic for a fixed grid) — returns true in that CTA's
// warp 0 after filling out[0..NV).
template <int NV>
__device__ __forceinline__ bool publish_and_reduce(const double (&mine)[NV], double* partials,
                                                   unsigned int* ticket, double (&out)[NV]) {
  __shared__ bool s_last;
  const unsigned nb = gridDim.x;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int v = 0; v < NV; ++v) partials[(size_t)v * nb + blockIdx.x] = mine[v];
    __threadfence();
    const unsigned t = atomicAdd(ticket, 1u);
    s_last = (t == nb - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  if (threadIdx.x >= 32) return false;
  const int lane = threadIdx.x;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    double s = 0.0;
    for (unsigned i = lane; i < nb; i += 32) s += __ldcg(partials + (size_t)v * nb + i);
    out[v] = warp_sum(s);
  }
  if (lane == 0) *ticket = 0u;
  return true;
}

}  // namespace tfem
