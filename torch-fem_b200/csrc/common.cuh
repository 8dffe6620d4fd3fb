// Shared helpers of libtfem_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "tfem_b200.h"

namespace tfem {

// last runtime error text of this host thread (reported by tfem_get_error_string)
void set_last_error(const char* what, const char* detail);

inline int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return TFEM_OK;
  set_last_error(what, cudaGetErrorString(e));
  return TFEM_ERR_CUDA;
}

#define TFEM_STR2(x) #x
#define TFEM_STR(x) TFEM_STR2(x)
#define TFEM_CUDA(call)                                                             \
  do {                                                                              \
    int _rc = ::tfem::check_cuda((call), #call " at " __FILE__ ":" TFEM_STR(__LINE__)); \
    if (_rc != TFEM_OK) return _rc;                                                 \
  } while (0)

#define TFEM_LAUNCH_CHECK(name) TFEM_CUDA((cudaGetLastError()))

#define TFEM_REQUIRE(cond, msg)                          \
  do {                                                   \
    if (!(cond)) {                                       \
      ::tfem::set_last_error("invalid argument", msg);   \
      return TFEM_ERR_INVALID;                           \
    }                                                    \
  } while (0)

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs

inline int num_sms() {  // cached per device (a process may drive several)
  static int sms[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return kSMs;
  dev &= 63;
  if (sms[dev] == 0 &&
      (cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms[dev] <= 0))
    sms[dev] = kSMs;
  return sms[dev];
}

// Stream-ordered scratch allocations come from the device's default memory pool. Its default release threshold is 0:
// every stream synchronisation hands unused pool memory back to the driver and the next allocation maps it again
// (measured: stalls of 10-300 ms inside setup calls that allocate a 4-byte flag). Keep the pool's memory instead —
// once per device. This changes a process-wide setting of the CUDA default pool (INTEGRATION.md says so);
// TFEM_KEEP_POOL=0 leaves the pool alone.
inline void keep_pool_memory() {
  static bool done[64] = {};
  int dev = 0;
  cudaMemPool_t pool;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  if (done[dev & 63]) return;
  done[dev & 63] = true;
  const char* off = getenv("TFEM_KEEP_POOL");
  if (off && off[0] == '0') return;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) return;
  uint64_t threshold = UINT64_MAX;
  cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
}

template <typename P>
inline cudaError_t malloc_async(P** ptr, size_t bytes, cudaStream_t st) {
  keep_pool_memory();
  return cudaMallocAsync(reinterpret_cast<void**>(ptr), bytes, st);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// inclusive warp scan
__device__ __forceinline__ int warp_scan_incl(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// streaming (read-once) loads: bypass L1 allocation so the cache is kept for the gathered vector
__device__ __forceinline__ int4 ldg_stream_int4(const int4* p) {
  int4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ double2 ldg_stream_double2(const double2* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ double ldg_stream_double(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ldg_stream_int(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

}  // namespace tfem
