// Shared helpers for the concepthash_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/concepthash_b200.h"

struct ch_ws {
  int device;
  int sm_count;
  int max_smem_optin;
  int l2_bytes;
  int clock_khz;
  // host->device staging (ch_pack_sign / ch_pack_labels with CH_MEM_HOST)
  void* stage[2];
  size_t stage_bytes;
  cudaStream_t copy_stream;
  cudaEvent_t ev_copied[2];
  cudaEvent_t ev_consumed[2];
  // pageable host sources: pinned bounce buffers filled by a few host threads (the driver's own pageable path
  // copies single-threaded, ~9 GB/s)
  void* bounce[2];
  int host_threads;
  int pack_threads;   // host threads of the pageable fp32 sign/bit-pack (host_pack.cpp)
  int64_t launches;
};

// host_pack.cpp: sign/bit-pack of pageable fp32 rows on the host's cores -> flag bits (1: a zero, 2: NaN)
uint32_t ch_host_pack_f32(const float* src, int64_t n, int ncols, int64_t rs, int words, uint32_t* out, int nthreads);
void ch_host_parallel_copy(void* dst, const void* src, size_t bytes, int nthreads);   // host_pack.cpp

// thread-local last error ---------------------------------------------------------------------------
void ch_set_error(const char* fmt, ...);

#define CH_FAIL(...)            \
  do {                          \
    ch_set_error(__VA_ARGS__);  \
    return 1;                   \
  } while (0)

#define CH_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ch_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

#define CH_LAUNCH_CHECK(ws)                                                                \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      ch_set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return 1;                                                                            \
    }                                                                                      \
    __atomic_fetch_add(&(ws)->launches, 1, __ATOMIC_RELAXED); /* (a loader thread may launch too) */                                                                     \
  } while (0)

struct ChDeviceGuard {
  int prev;
  bool ok;
  explicit ChDeviceGuard(int dev) : prev(-1), ok(true) {
    if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
    if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~ChDeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

static inline int64_t ch_round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---- device helpers -------------------------------------------------------------------------------
#ifdef __CUDACC__

// device-side view of ch_hist_args (+ derived geometry), shared by the POPC and the tensor-core kernels
struct HistDev {
  const uint32_t* q_bits; const uint32_t* q_nz;
  const uint32_t* g_bits; const uint32_t* g_nz;
  const uint32_t* q_lab;  const uint32_t* g_lab;
  uint32_t* slab_all; uint32_t* slab_rel;
  const uint32_t* thresh;
  const uint32_t* rec_off; const uint32_t* rec_cap; uint32_t* rec_cnt; uint4* recs;
  uint32_t* err_flag;
  long long nq, nq_pad, ndb, row_base;
  int nbit, nbins, lw, emit_mode;
  int nqtiles, rows_per_stripe, tile_rows, flush_tiles;
};



__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// mbarrier + 1-D bulk async copy (TMA engine, SASS: UBLKCP) --------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared::cta bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

#endif  // __CUDACC__
