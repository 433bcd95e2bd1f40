// K1: sign + bit-pack of real-valued codes, and the label packer (one-/multi-hot -> ids + bitmasks).
// HBM-bound: algorithmic bytes per code row = elem_size * nbit read + nbit / 8 written.
//
// Work unit = one 32-bit output word (row r, word w).  A warp owns 32 consecutive units: for unit u
// every lane loads element (r, 32 w + lane) -- a fully coalesced 128-byte request when the column
// stride is 1 -- and one __ballot_sync turns the 32 predicates into the packed word; lane u keeps it.
// After 32 units the warp stores its 32 words as one coalesced 128-byte line.  Loads are unrolled 8x
// so every warp keeps 1 KB of HBM requests in flight.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <thread>
#include <vector>

#include "common.cuh"

int ch_ws_ensure_stage(ch_ws* ws);   // api.cu
int ch_ws_ensure_bounce(ch_ws* ws);  // api.cu

namespace {

template <typename T>
struct Elem;
template <>
struct Elem<float> {
  typedef float cmp_t;
  static __device__ __forceinline__ float load(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ float round16(float x) { return x; }
};
template <>
struct Elem<double> {
  typedef double cmp_t;
  static __device__ __forceinline__ double load(const double* p) { return __ldg(p); }
  static __device__ __forceinline__ float round16(float x) { return x; }
};
template <>
struct Elem<__half> {
  typedef float cmp_t;
  static __device__ __forceinline__ float load(const __half* p) {
    return __half2float(__ushort_as_half(__ldg(reinterpret_cast<const unsigned short*>(p))));
  }
  static __device__ __forceinline__ float round16(float x) { return __half2float(__float2half_rn(x)); }
};
template <>
struct Elem<__nv_bfloat16> {
  typedef float cmp_t;
  static __device__ __forceinline__ float load(const __nv_bfloat16* p) {
    return __uint_as_float(static_cast<uint32_t>(__ldg(reinterpret_cast<const unsigned short*>(p))) << 16);
  }
  static __device__ __forceinline__ float round16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
};
template <>
struct Elem<int64_t> {
  typedef double cmp_t;
  static __device__ __forceinline__ double load(const int64_t* p) { return static_cast<double>(__ldg(p)); }
  static __device__ __forceinline__ float round16(float x) { return x; }
};
template <>
struct Elem<int32_t> {
  typedef double cmp_t;
  static __device__ __forceinline__ double load(const int32_t* p) { return static_cast<double>(__ldg(p)); }
  static __device__ __forceinline__ float round16(float x) { return x; }
};
template <>
struct Elem<int16_t> {
  typedef float cmp_t;
  static __device__ __forceinline__ float load(const int16_t* p) { return static_cast<float>(__ldg(p)); }
  static __device__ __forceinline__ float round16(float x) { return x; }
};
template <>
struct Elem<int8_t> {
  typedef float cmp_t;
  static __device__ __forceinline__ float load(const int8_t* p) {
    return static_cast<float>(static_cast<int8_t>(__ldg(reinterpret_cast<const signed char*>(p))));
  }
  static __device__ __forceinline__ float round16(float x) { return x; }
};
template <>
struct Elem<uint8_t> {
  typedef float cmp_t;
  static __device__ __forceinline__ float load(const uint8_t* p) { return static_cast<float>(__ldg(p)); }
  static __device__ __forceinline__ float round16(float x) { return x; }
};

// rows [row0, row_end) of the output are produced; rows >= n are zero pad rows.
// element (r, k) lives at src[(r - row0) * rs + k * cs].
template <typename T>
__global__ void __launch_bounds__(256) pack_bits_kernel(const T* __restrict__ src, int64_t row0, int64_t row_end,
                                                        int64_t n, int ncols, int64_t rs, int64_t cs,
                                                        double thr, const double* __restrict__ sub, int words,
                                                        uint32_t* __restrict__ out_pos,
                                                        uint32_t* __restrict__ out_nz, uint32_t* __restrict__ flags) {
  typedef typename Elem<T>::cmp_t C;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int64_t unit_begin = row0 * words;
  const int64_t unit_end = row_end * words;
  const C cthr = static_cast<C>(thr);
  const bool has_thr = thr != 0.0;
  uint32_t fl = 0;
  for (int64_t u0 = unit_begin + warp * 32; u0 < unit_end; u0 += nwarps * 32) {
    uint32_t my_pos = 0, my_nz = 0;
    // (row, word) of the first unit: one division per 32 units, then incremental
    int64_t r = u0 / words;
    int w = static_cast<int>(u0 - r * words);
    const T* rowp = src + (r - row0) * rs;
#pragma unroll 8
    for (int u = 0; u < 32; ++u) {
      const int col = w * 32 + lane;
      bool pos = false, nzb = false;
      if (r < n && col < ncols) {   // units past unit_end have r >= rows_pad > n
        C x = Elem<T>::load(rowp + static_cast<int64_t>(col) * cs);
        if (sub != nullptr) {   // zero_mean_eval: `codes - mean` as torch computes it, i.e. rounded to the codes' dtype
          x -= static_cast<C>(__ldg(sub + col));
          if (sizeof(T) == 2) x = static_cast<C>(Elem<T>::round16(static_cast<float>(x)));
        }
        if (x != x) fl |= 2u;
        if (has_thr && (x < C(0) ? -x : x) < cthr) x = C(0);
        pos = x > C(0);
        nzb = x != C(0);
        if (!nzb) fl |= 1u;
      }
      const uint32_t bp = __ballot_sync(0xffffffffu, pos);
      const uint32_t bn = __ballot_sync(0xffffffffu, nzb);
      if (lane == u) {
        my_pos = bp;
        my_nz = bn;
      }
      if (++w == words) {
        w = 0;
        ++r;
        rowp += rs;
      }
    }
    const int64_t mine = u0 + lane;
    if (mine < unit_end) {
      out_pos[mine] = my_pos;
      if (out_nz != nullptr) out_nz[mine] = my_nz;
    }
  }
  fl = __reduce_or_sync(0xffffffffu, fl);
  if (lane == 0 && fl != 0 && flags != nullptr) atomicOr(flags, fl);
}

// Fast path of K1: contiguous (n, nbit) codes with nbit in {32, 64, 128, 256}, no threshold, no non-zero plane.
// The tensor is then a flat stream of 128-byte units (32 elements -> one output word); a warp turns 32
// consecutive units into one coalesced 128-byte store with 32 fully unrolled, independent, coalesced loads in
// flight per thread (4 KB per warp).  Zeros / NaNs are only DETECTED here (flags); if a zero shows up the
// caller re-packs with the general kernel to obtain the non-zero plane.
template <typename T>
__global__ void __launch_bounds__(256) pack_sign_flat_kernel(const T* __restrict__ src, int64_t n, int nbit, int words,
                                                             int64_t nblocks_full, int64_t nblocks_all,
                                                             uint32_t* __restrict__ out_pos,
                                                             uint32_t* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  uint32_t fl = 0;
  for (int64_t blk = warp; blk < nblocks_all; blk += nwarps) {
    uint32_t mine = 0;
    if (blk < nblocks_full) {
      // 32 units entirely inside the tensor: no bounds checks, 32 independent coalesced loads in flight
      const T* p = src + blk * 1024 + lane;
      float x[32];
#pragma unroll
      for (int u = 0; u < 32; ++u) x[u] = static_cast<float>(Elem<T>::load(p + u * 32));
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        const uint32_t b = __ballot_sync(0xffffffffu, x[u] > 0.0f);
        if (!(fabsf(x[u]) > 0.0f)) fl |= (x[u] != x[u]) ? 2u : 1u;   // zero or NaN: rare
        if (lane == u) mine = b;
      }
    } else {
      // the last rows of the tensor and the zero pad rows
      for (int u = 0; u < 32; ++u) {
        const int64_t unit = blk * 32 + u;
        const int64_t r = unit / words;
        const int col = static_cast<int>(unit - r * words) * 32 + lane;
        bool pos = false;
        if (r < n && col < nbit) {
          const float x = static_cast<float>(Elem<T>::load(src + r * nbit + col));
          pos = x > 0.0f;
          if (!(fabsf(x) > 0.0f)) fl |= (x != x) ? 2u : 1u;
        }
        const uint32_t b = __ballot_sync(0xffffffffu, pos);
        if (lane == u) mine = b;
      }
    }
    out_pos[blk * 32 + lane] = mine;
  }
  fl = __reduce_or_sync(0xffffffffu, fl);
  if (lane == 0 && fl != 0) atomicOr(flags, fl);
}

// per-row summary of a label bitmask: id of the first positive (or nolabel), statistics
__global__ void label_rows_kernel(const uint32_t* __restrict__ masks, int64_t n, int64_t rows_pad, int words,
                                  uint32_t nolabel, uint32_t* __restrict__ ids, uint32_t* __restrict__ info) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  uint32_t cnt = 0, first = nolabel;
  if (r < n) {
    for (int w = 0; w < words; ++w) {
      const uint32_t m = masks[r * words + w];
      if (m != 0 && cnt == 0) first = static_cast<uint32_t>(w * 32 + __ffs(m) - 1);
      cnt += __popc(m);
    }
  }
  if (r < rows_pad) ids[r] = first;
  // warp-aggregate the statistics
  uint32_t mx = __reduce_max_sync(0xffffffffu, cnt);
  uint32_t mid = __reduce_max_sync(0xffffffffu, (r < n && cnt > 0) ? first + 1 : 0u);
  uint32_t nz = __popc(__ballot_sync(0xffffffffu, r < n && cnt == 0));
  // ... then over the block (launched with 256 threads): three atomics per block, not per warp
  __shared__ uint32_t s_mx[8], s_mid[8], s_nz[8];
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    s_mx[warp] = mx;
    s_mid[warp] = mid;
    s_nz[warp] = nz;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) {
      mx = max(mx, s_mx[w]);
      mid = max(mid, s_mid[w]);
      nz += s_nz[w];
    }
    if (mx) atomicMax(info + 0, mx);
    if (mid) atomicMax(info + 1, mid);
    if (nz) atomicAdd(info + 2, nz);
  }
}

// 1-D integer class ids -> u32 ids (negative -> nolabel).  kLabelRowsPerThread rows per thread, the label
// statistics reduced over the block first: three atomics per 1024 rows (one per WARP and 32 rows made the 12.5 M ids
// of a cfg5 shard a 0.57 ms queue on two L2 addresses; the data themselves take 0.03 ms)
constexpr int kLabelRowsPerThread = 4;
template <typename T>
__global__ void __launch_bounds__(256) label_ids_kernel(const T* __restrict__ src, int64_t row0, int64_t row_end,
                                                        int64_t n, int64_t rs, uint32_t nolabel,
                                                        uint32_t* __restrict__ ids, uint32_t* __restrict__ info) {
  __shared__ uint32_t s_mid[8], s_nz[8];
  const int64_t base = row0 + static_cast<int64_t>(blockIdx.x) * (256 * kLabelRowsPerThread) + threadIdx.x;
  uint32_t mid = 0u, nz = 0u;
#pragma unroll
  for (int j = 0; j < kLabelRowsPerThread; ++j) {
    const int64_t r = base + static_cast<int64_t>(j) * 256;
    uint32_t id = nolabel;
    if (r < n) {
      const double v = static_cast<double>(Elem<T>::load(src + (r - row0) * rs));
      if (v >= 0.0 && v < 4294967000.0) {
        id = static_cast<uint32_t>(v);
        mid = max(mid, id + 1u);
      } else {
        ++nz;
      }
    }
    if (r < row_end) ids[r] = id;
  }
  mid = __reduce_max_sync(0xffffffffu, mid);
  nz = __reduce_add_sync(0xffffffffu, nz);
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    s_mid[warp] = mid;
    s_nz[warp] = nz;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      mid = max(mid, s_mid[w]);
      nz += s_nz[w];
    }
    if (mid) {                       // (some row of the block holds a valid id)
      atomicMax(info + 0, 1u);
      atomicMax(info + 1, mid);
    }
    if (nz) atomicAdd(info + 2, nz);
  }
}

// zero_mean_eval (experiments/train_helper.py:223-226): column sums of the gallery codes, deterministic --
// every warp sums the rows r = warp, warp + nwarps, ... of 32 columns at a time into its own partial row,
// a second kernel adds the partials in warp order.
template <typename T>
__global__ void __launch_bounds__(256) column_partials_kernel(const T* __restrict__ src, int64_t n, int ncols,
                                                              int64_t rs, int64_t cs, double* __restrict__ part) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int c0 = 0; c0 < ncols; c0 += 32) {
    const int col = c0 + lane;
    double acc = 0.0;
    if (col < ncols)
      for (int64_t r = warp; r < n; r += nwarps)
        acc += static_cast<double>(Elem<T>::load(src + r * rs + static_cast<int64_t>(col) * cs));
    if (col < ncols) part[warp * ncols + col] = acc;
  }
}
__global__ void column_reduce_kernel(const double* __restrict__ part, int64_t nwarps, int ncols,
                                     double* __restrict__ sums) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  double acc = 0.0;
  for (int64_t w = 0; w < nwarps; ++w) acc += part[w * ncols + col];
  sums[col] = acc;
}

__global__ void or_flags_kernel(uint32_t* __restrict__ flags, uint32_t bits) { atomicOr(flags, bits); }

size_t elem_size(int dtype) {
  switch (dtype) {
    case CH_F32: return 4;
    case CH_F16: return 2;
    case CH_BF16: return 2;
    case CH_F64: return 8;
    case CH_I64: return 8;
    case CH_I32: return 4;
    case CH_U8: return 1;
    case CH_I16: return 2;
    case CH_I8: return 1;
    default: return 0;
  }
}

int launch_pack(ch_ws* ws, const void* src, int dtype, int64_t row0, int64_t row_end, int64_t n, int ncols,
                int64_t rs, int64_t cs, double thr, int words, uint32_t* out_pos, uint32_t* out_nz,
                uint32_t* flags, cudaStream_t st, const double* sub = nullptr) {
  const int64_t units = (row_end - row0) * words;
  if (units <= 0) return 0;
  const int64_t warps = (units + 31) / 32;
  int64_t blocks = (warps + 7) / 8;
  const int64_t cap = static_cast<int64_t>(ws->sm_count) * 16;
  if (blocks > cap) blocks = cap;
  const dim3 grid(static_cast<unsigned>(blocks)), block(256);
#define CH_PACK_CASE(ENUM, TYPE)                                                                            \
  case ENUM:                                                                                                \
    pack_bits_kernel<TYPE><<<grid, block, 0, st>>>(static_cast<const TYPE*>(src), row0, row_end, n, ncols,  \
                                                   rs, cs, thr, sub, words, out_pos, out_nz, flags);        \
    break;
  switch (dtype) {
    CH_PACK_CASE(CH_F32, float)
    CH_PACK_CASE(CH_F16, __half)
    CH_PACK_CASE(CH_BF16, __nv_bfloat16)
    CH_PACK_CASE(CH_F64, double)
    CH_PACK_CASE(CH_I64, int64_t)
    CH_PACK_CASE(CH_I32, int32_t)
    CH_PACK_CASE(CH_U8, uint8_t)
    CH_PACK_CASE(CH_I16, int16_t)
    CH_PACK_CASE(CH_I8, int8_t)
    default:
      CH_FAIL("unsupported dtype %d", dtype);
  }
#undef CH_PACK_CASE
  CH_LAUNCH_CHECK(ws);
  return 0;
}

// Runs `launch_pack` over a host buffer in pipelined chunks: H2D of chunk c+1 (copy stream) overlaps
// the pack kernel of chunk c (compute stream).  The host buffer must be row-contiguous (cs == 1).
int pack_from_host(ch_ws* ws, const void* src, int dtype, int64_t n, int ncols, int64_t rs, int64_t cs,
                   double thr, int words, int64_t rows_pad, uint32_t* out_pos, uint32_t* out_nz,
                   uint32_t* flags, cudaStream_t st, const double* sub = nullptr) {
  const size_t es = elem_size(dtype);
  if (cs != 1 && ncols > 1) CH_FAIL("host buffers must have unit column stride");
  if (ch_ws_ensure_stage(ws)) return 1;
  const size_t row_bytes = static_cast<size_t>(rs) * es;          // host pitch
  const size_t dense_bytes = static_cast<size_t>(ncols) * es;     // what a row really holds
  if (row_bytes == 0 || rs < ncols) CH_FAIL("bad row stride");
  // rows are staged DENSELY (a strided 2-D DMA drops the gaps of column slices / row-sampled views)
  int64_t chunk_rows = static_cast<int64_t>(ws->stage_bytes / dense_bytes) / 64 * 64;
  if (chunk_rows < 64) CH_FAIL("row of %zu bytes does not fit the staging buffer", dense_bytes);
  if (n == 0)
    return launch_pack(ws, nullptr, dtype, 0, rows_pad, 0, ncols, ncols, 1, thr, words, out_pos, out_nz, flags, st, sub);
  // pageable source?  (cudaMemcpyAsync from pageable memory is staged by the driver on one thread, ~9 GB/s: copy
  // into pinned bounce buffers with a few host threads instead and DMA from there -- the memcpy of chunk c + 1
  // overlaps the DMA of chunk c)
  bool pageable = true;
  {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, src) == cudaSuccess)
      pageable = attr.type == cudaMemoryTypeUnregistered;
    else
      cudaGetLastError();   // older drivers report unregistered host pointers as an error
  }
  if (pageable && ch_ws_ensure_bounce(ws)) return 1;
  // fp32 codes in host memory, plain sign test: for pageable memory the host's cores have to touch every byte anyway
  // (nothing can DMA from it), so they do the sign test in that one pass and only the BITS travel (host_pack.cpp).
  // Pinned memory takes the same route: the cores read it at 150-200 GB/s, the DMA engine at ~43 GB/s (NABirds from
  // pinned tensors: 1.95 ms by DMA + kernel, 1.19 ms from pageable tensors by this path).
  if (dtype == CH_F32 && thr == 0.0 && sub == nullptr && out_nz == nullptr &&
      ws->pack_threads > 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0) {
    if (ch_ws_ensure_bounce(ws)) return 1;
    const size_t out_row = static_cast<size_t>(words) * 4;
    const int64_t rows_per_buf = static_cast<int64_t>(ws->stage_bytes / out_row);
    int c = 0;
    uint32_t fl = 0;
    for (int64_t r0 = 0; r0 < n; r0 += rows_per_buf, ++c) {
      const int b = c & 1;
      const int64_t r1 = (r0 + rows_per_buf < n) ? r0 + rows_per_buf : n;
      CH_CUDA(cudaEventSynchronize(ws->ev_copied[b]));          // the DMA that last read this bounce buffer is done
      uint32_t* dst = static_cast<uint32_t*>(ws->bounce[b]);
      fl |= ch_host_pack_f32(static_cast<const float*>(src) + static_cast<size_t>(r0) * rs, r1 - r0, ncols, rs, words,
                             dst, ws->pack_threads);
      CH_CUDA(cudaMemcpyAsync(out_pos + static_cast<size_t>(r0) * words, dst, static_cast<size_t>(r1 - r0) * out_row,
                              cudaMemcpyHostToDevice, st));
      CH_CUDA(cudaEventRecord(ws->ev_copied[b], st));
    }
    CH_CUDA(cudaMemsetAsync(out_pos + static_cast<size_t>(n) * words, 0, static_cast<size_t>(rows_pad - n) * out_row, st));
    if (fl != 0u && flags != nullptr) {
      or_flags_kernel<<<1, 1, 0, st>>>(flags, fl);
      CH_LAUNCH_CHECK(ws);
    }
    return 0;
  }
  int c = 0;
  for (int64_t r0 = 0; r0 < n; r0 += chunk_rows, ++c) {
    const int b = c & 1;
    const int64_t r1 = (r0 + chunk_rows < n) ? r0 + chunk_rows : n;
    const bool last = r1 == n;
    // the staging buffer may still be read by the pack kernel of chunk c - 2 (or of an earlier call)
    CH_CUDA(cudaStreamWaitEvent(ws->copy_stream, ws->ev_consumed[b], 0));
    const char* src_rows = static_cast<const char*>(src) + static_cast<size_t>(r0) * row_bytes;
    const size_t rows_c = static_cast<size_t>(r1 - r0);
    if (pageable) {
      // the bounce buffer is free once the DMA of chunk c - 2 has been issued AND completed
      CH_CUDA(cudaEventSynchronize(ws->ev_copied[b]));
      char* dst = static_cast<char*>(ws->bounce[b]);
      int nt = ws->host_threads;
      if (rows_c * dense_bytes < (static_cast<size_t>(4) << 20)) nt = 1;
      auto work = [=](int t) {
        const size_t ra = rows_c * t / nt, rb = rows_c * (t + 1) / nt;
        if (rs == ncols) {
          memcpy(dst + ra * dense_bytes, src_rows + ra * dense_bytes, (rb - ra) * dense_bytes);
        } else {
          for (size_t r = ra; r < rb; ++r) memcpy(dst + r * dense_bytes, src_rows + r * row_bytes, dense_bytes);
        }
      };
      if (nt == 1) {
        work(0);
      } else {
        std::vector<std::thread> pool;
        pool.reserve(nt - 1);
        for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
        work(0);
        for (auto& th : pool) th.join();
      }
      CH_CUDA(cudaMemcpyAsync(ws->stage[b], dst, rows_c * dense_bytes, cudaMemcpyHostToDevice, ws->copy_stream));
    } else if (rs == ncols) {
      CH_CUDA(cudaMemcpyAsync(ws->stage[b], src_rows, rows_c * dense_bytes, cudaMemcpyHostToDevice, ws->copy_stream));
    } else {
      CH_CUDA(cudaMemcpy2DAsync(ws->stage[b], dense_bytes, src_rows, row_bytes, dense_bytes, rows_c,
                                cudaMemcpyHostToDevice, ws->copy_stream));
    }
    CH_CUDA(cudaEventRecord(ws->ev_copied[b], ws->copy_stream));
    CH_CUDA(cudaStreamWaitEvent(st, ws->ev_copied[b], 0));
    if (launch_pack(ws, ws->stage[b], dtype, r0, last ? rows_pad : r1, n, ncols, ncols, 1, thr, words, out_pos, out_nz,
                    flags, st, sub))
      return 1;
    CH_CUDA(cudaEventRecord(ws->ev_consumed[b], st));
  }
  // the caller may free / reuse the host buffer after return: a pageable source has been copied out by the host
  // threads already (its last DMA may still be in flight -- the next call's host copy overlaps it)
  if (!pageable) CH_CUDA(cudaStreamSynchronize(ws->copy_stream));
  return 0;
}

}  // namespace

extern "C" int ch_host_pack_sign(const float* codes_host, int64_t n, int nbit, int64_t row_stride,
                                 uint32_t* out_bits_host, uint32_t* flags_host, int threads) {
  const int words = ch_code_words(nbit);
  if (words == 0) CH_FAIL("nbit=%d unsupported (1..%d)", nbit, CH_MAX_NBIT);
  if (n < 0 || (n > 0 && (codes_host == nullptr || out_bits_host == nullptr)) || row_stride < nbit)
    CH_FAIL("bad arguments to ch_host_pack_sign");
  const uint32_t fl = ch_host_pack_f32(codes_host, n, nbit, row_stride, words, out_bits_host, threads < 1 ? 1 : threads);
  if (flags_host != nullptr) *flags_host |= fl;
  return 0;
}

extern "C" int64_t ch_padded_rows(int64_t n) { return ch_round_up(n < 0 ? 0 : n, 64) + 64; }

extern "C" int ch_code_words(int nbit) {
  if (nbit <= 0 || nbit > CH_MAX_NBIT) return 0;
  if (nbit <= 32) return 1;
  if (nbit <= 64) return 2;
  if (nbit <= 128) return 4;
  return 8;
}

int ch_ws_scratch(ch_ws* ws, size_t bytes, void** out);  // api.cu

extern "C" int ch_column_sums(ch_ws* ws, const void* codes_dev, int dtype, int64_t n, int ncols, int64_t row_stride,
                              int64_t col_stride, double* sums_dev, void* stream) {
  if (ws == nullptr || sums_dev == nullptr || (codes_dev == nullptr && n > 0)) CH_FAIL("null argument to ch_column_sums");
  if (ncols <= 0 || n < 0) CH_FAIL("bad shape (%lld, %d)", (long long)n, ncols);
  if (!(dtype == CH_F32 || dtype == CH_F16 || dtype == CH_BF16 || dtype == CH_F64))
    CH_FAIL("codes must be a floating dtype (got %d)", dtype);
  ChDeviceGuard g(ws->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t blocks = (n + 7) / 8;
  const int64_t cap = static_cast<int64_t>(ws->sm_count) * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const int64_t nwarps = blocks * 8;
  void* scratch = nullptr;
  if (ch_ws_scratch(ws, static_cast<size_t>(nwarps) * ncols * sizeof(double), &scratch)) return 1;
  double* part = static_cast<double*>(scratch);
  const dim3 grid(static_cast<unsigned>(blocks)), block(256);
  switch (dtype) {
    case CH_F32:
      column_partials_kernel<float><<<grid, block, 0, st>>>(static_cast<const float*>(codes_dev), n, ncols,
                                                            row_stride, col_stride, part);
      break;
    case CH_F16:
      column_partials_kernel<__half><<<grid, block, 0, st>>>(static_cast<const __half*>(codes_dev), n, ncols,
                                                             row_stride, col_stride, part);
      break;
    case CH_BF16:
      column_partials_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(static_cast<const __nv_bfloat16*>(codes_dev), n,
                                                                    ncols, row_stride, col_stride, part);
      break;
    default:
      column_partials_kernel<double><<<grid, block, 0, st>>>(static_cast<const double*>(codes_dev), n, ncols,
                                                             row_stride, col_stride, part);
      break;
  }
  CH_LAUNCH_CHECK(ws);
  column_reduce_kernel<<<(ncols + 127) / 128, 128, 0, st>>>(part, nwarps, ncols, sums_dev);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_pack_sign(ch_ws* ws, const void* codes, int mem, int dtype, int64_t n, int nbit,
                            int64_t row_stride, int64_t col_stride, double threshold, const double* col_sub_dev,
                            uint32_t* out_bits_dev, uint32_t* out_nz_dev, uint32_t* flags_dev, void* stream) {
  if (ws == nullptr) CH_FAIL("null workspace");
  const int words = ch_code_words(nbit);
  if (words == 0) CH_FAIL("nbit=%d unsupported (1..%d)", nbit, CH_MAX_NBIT);
  if (n < 0 || out_bits_dev == nullptr) CH_FAIL("bad arguments to ch_pack_sign");
  if (!(dtype == CH_F32 || dtype == CH_F16 || dtype == CH_BF16 || dtype == CH_F64))
    CH_FAIL("codes must be a floating dtype (got %d)", dtype);
  ChDeviceGuard g(ws->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t rows_pad = ch_padded_rows(n);
  if (mem == CH_MEM_HOST)
    return pack_from_host(ws, codes, dtype, n, nbit, row_stride, col_stride, threshold, words, rows_pad,
                          out_bits_dev, out_nz_dev, flags_dev, st, col_sub_dev);
  // flat fast path (see pack_sign_flat_kernel): one launch covers the tensor, its tail and the pad rows
  const bool flat = out_nz_dev == nullptr && threshold == 0.0 && col_sub_dev == nullptr &&
                    (col_stride == 1 || nbit == 1) &&
                    (row_stride == nbit || n <= 1) && nbit == words * 32 && flags_dev != nullptr &&
                    (dtype == CH_F32 || dtype == CH_F16 || dtype == CH_BF16) &&
                    (reinterpret_cast<uintptr_t>(codes) & 3) == 0;
  if (flat) {
    const int64_t nblocks_full = n * words / 32;             // 32-unit blocks without any bounds check
    const int64_t nblocks_all = rows_pad * words / 32;       // rows_pad is a multiple of 64
    int64_t blocks = (nblocks_all + 7) / 8;
    const int64_t cap = static_cast<int64_t>(ws->sm_count) * 8;
    if (blocks > cap) blocks = cap;
    const dim3 grid(static_cast<unsigned>(blocks)), block(256);
    if (dtype == CH_F32)
      pack_sign_flat_kernel<float><<<grid, block, 0, st>>>(static_cast<const float*>(codes), n, nbit, words,
                                                           nblocks_full, nblocks_all, out_bits_dev, flags_dev);
    else if (dtype == CH_F16)
      pack_sign_flat_kernel<__half><<<grid, block, 0, st>>>(static_cast<const __half*>(codes), n, nbit, words,
                                                            nblocks_full, nblocks_all, out_bits_dev, flags_dev);
    else
      pack_sign_flat_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(static_cast<const __nv_bfloat16*>(codes), n, nbit,
                                                                   words, nblocks_full, nblocks_all, out_bits_dev,
                                                                   flags_dev);
    CH_LAUNCH_CHECK(ws);
    return 0;
  }
  return launch_pack(ws, codes, dtype, 0, rows_pad, n, nbit, row_stride, col_stride, threshold, words, out_bits_dev,
                     out_nz_dev, flags_dev, st, col_sub_dev);
}

extern "C" int ch_pack_labels(ch_ws* ws, const void* labels, int mem, int dtype, int64_t n, int C,
                              int64_t row_stride, int64_t col_stride, uint32_t nolabel, uint32_t* out_ids_dev,
                              uint32_t* out_masks_dev, uint32_t* info_dev, void* stream) {
  if (ws == nullptr) CH_FAIL("null workspace");
  if (n < 0 || out_ids_dev == nullptr || info_dev == nullptr) CH_FAIL("bad arguments to ch_pack_labels");
  if (elem_size(dtype) == 0) CH_FAIL("unsupported label dtype %d", dtype);
  ChDeviceGuard g(ws->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t rows_pad = ch_padded_rows(n);
  if (C > 0) {
    if (out_masks_dev == nullptr) CH_FAIL("2-D labels need a mask buffer");
    const int words = (C + 31) / 32;
    int rc;
    if (mem == CH_MEM_HOST)
      rc = pack_from_host(ws, labels, dtype, n, C, row_stride, col_stride, 0.0, words, rows_pad, out_masks_dev,
                          nullptr, nullptr, st);
    else
      rc = launch_pack(ws, labels, dtype, 0, rows_pad, n, C, row_stride, col_stride, 0.0, words, out_masks_dev,
                       nullptr, nullptr, st);
    if (rc) return rc;
    const unsigned blocks = static_cast<unsigned>((rows_pad + 255) / 256);
    label_rows_kernel<<<blocks, 256, 0, st>>>(out_masks_dev, n, rows_pad, words, nolabel, out_ids_dev, info_dev);
    CH_LAUNCH_CHECK(ws);
    return 0;
  }
  // 1-D ids (host arrays go through the staging buffer in stream-ordered chunks)
  const size_t es = elem_size(dtype);
  const int64_t chunk_rows = mem == CH_MEM_HOST
                                 ? static_cast<int64_t>(ws->stage_bytes / (static_cast<size_t>(row_stride) * es)) / 64 * 64
                                 : rows_pad;
  if (mem == CH_MEM_HOST) {
    if (chunk_rows < 64) CH_FAIL("label row stride too large for the staging buffer");
    if (ch_ws_ensure_stage(ws)) return 1;
  }
  bool user_buffer_in_flight = false;   // a DMA may still be reading the CALLER's memory when the loop ends
  for (int64_t r0 = 0; r0 == 0 || r0 < n; r0 += chunk_rows) {
    const int64_t r1 = (r0 + chunk_rows < n) ? r0 + chunk_rows : n;
    const bool last = r1 >= n;
    const void* src = labels;
    if (mem == CH_MEM_HOST) {
      const size_t bytes = r1 > r0 ? (static_cast<size_t>(r1 - r0 - 1) * row_stride + 1) * es : 0;
      const char* hsrc = static_cast<const char*>(labels) + static_cast<size_t>(r0) * row_stride * es;
      if (bytes >= (static_cast<size_t>(1) << 20) && ws->pack_threads > 1) {
        // large pageable label arrays: the pool copies them into the pinned bounce buffer, the DMA runs from there
        cudaPointerAttributes attr;
        bool pageable = true;
        if (cudaPointerGetAttributes(&attr, hsrc) == cudaSuccess) pageable = attr.type == cudaMemoryTypeUnregistered;
        else cudaGetLastError();
        if (pageable) {
          if (ch_ws_ensure_bounce(ws)) return 1;
          CH_CUDA(cudaEventSynchronize(ws->ev_copied[0]));   // the DMA that last read this bounce buffer is done
          ch_host_parallel_copy(ws->bounce[0], hsrc, bytes, ws->pack_threads);
          hsrc = static_cast<const char*>(ws->bounce[0]);
        }
      }
      // (the staging buffer may still be read by a kernel of an earlier call on another stream: pack_from_host's protocol)
      CH_CUDA(cudaStreamWaitEvent(st, ws->ev_consumed[0], 0));
      if (bytes) CH_CUDA(cudaMemcpyAsync(ws->stage[0], hsrc, bytes, cudaMemcpyHostToDevice, st));
      if (hsrc == static_cast<const char*>(ws->bounce[0])) CH_CUDA(cudaEventRecord(ws->ev_copied[0], st));
      else if (bytes) user_buffer_in_flight = true;
      src = ws->stage[0];
    } else {
      src = static_cast<const char*>(labels);
    }
    const int64_t row_end = last ? rows_pad : r1;
    const unsigned blocks = static_cast<unsigned>((row_end - r0 + 256 * kLabelRowsPerThread - 1) / (256 * kLabelRowsPerThread));
#define CH_ID_CASE(ENUM, TYPE)                                                                              \
  case ENUM:                                                                                                \
    label_ids_kernel<TYPE><<<blocks, 256, 0, st>>>(static_cast<const TYPE*>(src), r0, row_end, n, row_stride, \
                                                   nolabel, out_ids_dev, info_dev);                         \
    break;
    switch (dtype) {
      CH_ID_CASE(CH_I64, int64_t)
      CH_ID_CASE(CH_I32, int32_t)
      CH_ID_CASE(CH_I16, int16_t)
      CH_ID_CASE(CH_I8, int8_t)
      CH_ID_CASE(CH_U8, uint8_t)
      CH_ID_CASE(CH_F32, float)
      CH_ID_CASE(CH_F64, double)
      default:
        CH_FAIL("unsupported dtype %d for 1-D labels", dtype);
    }
    CH_LAUNCH_CHECK(ws);
    if (mem == CH_MEM_HOST) CH_CUDA(cudaEventRecord(ws->ev_consumed[0], st));
    if (last) break;
  }
#undef CH_ID_CASE
  // The caller's host buffer must be released on return.  A pageable array that went through the bounce buffer is
  // (the pool copied it; the bounce buffer is guarded by its event, the staging buffer is reused in stream order):
  // no wait -- the 8 MB of int64 ids of a 1M-row gallery cost 1.5 ms of host time here, a sixth of the whole
  // evaluation.  Anything still being read by a DMA is waited for.
  if (user_buffer_in_flight) CH_CUDA(cudaStreamSynchronize(st));
  return 0;
}
