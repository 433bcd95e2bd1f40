// Development probe: tcgen05.mma kind::i8 with the A operand in TENSOR MEMORY (written there with tcgen05.st,
// lane = row, four int8 per 32-bit column) and B in shared memory (no-swizzle K-major core matrices).
//   part 1: correctness of the assumed A layout against a CPU product (and a layout-revealing pattern)
//   part 2: back-to-back issue rate, and the latency of a shared-memory CAS / a 20 KB bulk copy beside it
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tmem_a_probe tmem_a_probe.cu && ./tmem_a_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra.uni WAIT_DONE;\nbra.uni WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

constexpr int KB = 160;
constexpr int kACol = 384;     // A: columns 384 .. 384 + KB / 4

// MODE 0: one product, D written out.  MODE 1: rate (iters batches of KB / 32 MMAs over 3 accumulators) with one
// bulk-copy thread and SPIN spinning threads beside it.
template <int MODE, int SPIN>
__global__ void __launch_bounds__(128 + 32 * (1 + SPIN)) probe(const int8_t* __restrict__ A, const int8_t* __restrict__ Bt,
                                                             int* __restrict__ D, int iters, long long* out,
                                                             const unsigned char* gsrc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar_mma, bar_cp;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop_s;
  __shared__ unsigned int spin_word;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * KB / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = reinterpret_cast<const uint32_t*>(Bt)[i];
  if (tid == 0) {
    stop_s = 0;
    spin_word = 1u;
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_cp, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  if (warp < 4) {
    // thread = row of A = TMEM lane: word j holds bytes k = 4 j .. 4 j + 3
    const uint32_t* row = reinterpret_cast<const uint32_t*>(A + (size_t)tid * KB);
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + kACol;
#pragma unroll
    for (int j = 0; j < KB / 4; j += 4) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr + j), "r"(row[j]),
                   "r"(row[j + 1]), "r"(row[j + 2]), "r"(row[j + 3])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | (8u << 24);
  const uint32_t sB = smem_u32(smem);
  if (tid == 0) {
    const long long t0 = clock64();
    const int n = MODE == 0 ? 1 : iters;
    for (int it = 0; it < n; ++it) {
      const uint32_t d = tmem_base + (uint32_t)(it % 3) * 128;
#pragma unroll
      for (int k = 0; k < KB / 32; ++k)
        mma_ts(d, tmem_base + kACol + k * 8, make_desc(sB + k * 256, 128, 8 * KB), idesc, k > 0 ? 1u : 0u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    mbar_wait(&bar_mma, 0);
    const long long t1 = clock64();
    if (MODE == 1) out[blockIdx.x] = t1 - t0;
    stop_s = 1;
  } else if (MODE == 1 && warp == 4) {
    if ((tid & 31) == 0) {
      unsigned char* dst = smem + 128 * KB;
      uint32_t ph = 0;
      long long n = 0;
      while (stop_s == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_cp)), "r"(20480u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(gsrc + (size_t)blockIdx.x * 20480 + (n & 7) * 20480 * 148), "r"(20480u), "r"(smem_u32(&bar_cp)) : "memory");
        mbar_wait(&bar_cp, ph);
        ph ^= 1u;
        ++n;
      }
      out[148 + blockIdx.x] = n;
    }
  } else if (MODE == 1 && SPIN > 0 && warp > 4) {
    if ((tid & 31) == 0) {
      long long n = 0;
      while (stop_s == 0) {
        if (atomicCAS(&spin_word, 0u, 1u) == 0u) break;
        ++n;
      }
      if (warp == 5) out[296 + blockIdx.x] = n;
    }
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (MODE == 0 && warp < 4) {
    const int row = tid;
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) D[row * 128 + c0 + j] = (int)r[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

static size_t tiled(int r, int k) { return (size_t)(r / 8) * (8 * KB) + (size_t)(k / 16) * 128 + (r % 8) * 16 + (k % 16); }

int main() {
  std::vector<int8_t> a(128 * KB), b(128 * KB), bt(128 * KB);
  int8_t *dA, *dB;
  int* dD;
  long long* dOut;
  unsigned char* gsrc;
  CK(cudaMalloc(&dA, a.size())); CK(cudaMalloc(&dB, bt.size())); CK(cudaMalloc(&dD, 128 * 128 * 4));
  CK(cudaMalloc(&dOut, 3 * 148 * 8)); CK(cudaMalloc(&gsrc, (size_t)20480 * 148 * 8));
  CK(cudaMemset(gsrc, 1, (size_t)20480 * 148 * 8));
  const int smem = 128 * KB + 20480;
  CK(cudaFuncSetAttribute(probe<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(probe<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(probe<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  std::vector<int> d(128 * 128);
  for (int test = 0; test < 2; ++test) {
    srand(7);
    for (int r = 0; r < 128; ++r)
      for (int k = 0; k < KB; ++k) {
        a[r * KB + k] = test == 0 ? (int8_t)(rand() % 7 - 3) : (int8_t)(k < 32 ? k + 1 : 0);
        b[r * KB + k] = test == 0 ? (int8_t)(rand() % 7 - 3) : (int8_t)(k == r ? 1 : 0);
      }
    for (int r = 0; r < 128; ++r) for (int k = 0; k < KB; ++k) bt[tiled(r, k)] = b[r * KB + k];
    CK(cudaMemcpy(dA, a.data(), a.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, bt.data(), bt.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0x7f, 128 * 128 * 4));
    probe<0, 0><<<1, 160, smem>>>(dA, dB, dD, 1, dOut, gsrc);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(d.data(), dD, d.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int i = 0; i < 128; ++i)
      for (int j = 0; j < 128; ++j) {
        int ref = 0;
        for (int k = 0; k < KB; ++k) ref += (int)a[i * KB + k] * (int)b[j * KB + k];
        if (ref != d[i * 128 + j]) ++bad;
      }
    printf("test %d (%s): %ld mismatches of %d\n", test, test == 0 ? "random" : "layout pattern", bad, 128 * 128);
    if (test == 1) {
      printf("  row 0, D[0][0..39] (expect 1..32, then 0):");
      for (int j = 0; j < 40; ++j) printf(" %d", d[j]);
      printf("\n  row 37, D[37][0..39]:");
      for (int j = 0; j < 40; ++j) printf(" %d", d[37 * 128 + j]);
      printf("\n");
    }
  }
  for (int spin = 0; spin < 2; ++spin) {
    const int iters = 2000;
    CK(cudaMemset(dOut, 0, 3 * 148 * 8));
    for (int rep = 0; rep < 2; ++rep) {
      if (spin == 0) probe<1, 0><<<148, 160, smem>>>(dA, dB, dD, iters, dOut, gsrc);
      else probe<1, 3><<<148, 256, smem>>>(dA, dB, dD, iters, dOut, gsrc);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
    }
    std::vector<long long> h(3 * 148);
    CK(cudaMemcpy(h.data(), dOut, h.size() * 8, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("A in TMEM, N=128, 3 accumulators, %d spinners: %.1f clk/MMA; bulk copy %.0f clk per 20 KB tile", spin ? 3 : 0,
           (double)mx / (iters * (KB / 32)), (double)mx / (h[148] ? h[148] : 1));
    if (spin) printf("; %.0f clk per CAS", (double)mx / (h[296] ? h[296] : 1));
    printf("\n");
  }
  return 0;
}
