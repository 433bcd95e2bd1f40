// Development probe: back-to-back tcgen05.mma kind::i8 throughput (clocks per instruction) for M=128 and
// N=128 / N=256 with operands resident in shared memory (no-swizzle K-major core-matrix layout, K = 160 bytes),
// one CTA per SM on every SM.  Answers: is the M=N=128 instruction limited by its operand fetch?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o mma_rate_probe mma_rate_probe.cu && ./mma_rate_probe
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

constexpr int KB = 160;

// N = MMA width; NACC accumulators used round-robin; KBLK k-blocks (32 bytes) per batch.
template <int N, int NACC, int KBLK, int LDW, int TMA, int SPIN>
__global__ void __launch_bounds__(128 + 32 * LDW + 32 + 32 * (1 + SPIN)) probe(int iters, long long* out, const unsigned char* gsrc, int commit_every) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ __align__(8) uint64_t bar_batch[4];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop_s;
  __shared__ __align__(8) uint64_t bar_cp;
  __shared__ unsigned int spin_word;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + 256) * KB / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01ff01ffu;
  if (tid == 0) {
    stop_s = 0;
    spin_word = 1u;
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_cp, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_batch[i], 1);
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
  if (tid == 0) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t sbo = 8 * KB, lbo = 128;
    const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + 128 * KB);
    const long long t0 = clock64();
    if (NACC == 0) {                       // idle baseline: no MMA at all, just let the other threads run
      while (clock64() - t0 < 1000000) {
      }
    }
    for (int it = 0; it < (NACC == 0 ? 0 : iters); ++it) {
      const uint32_t acc_col = (uint32_t)(it % (NACC == 0 ? 1 : NACC)) * N;
#pragma unroll
      for (int k = 0; k < KBLK; ++k) {
        const uint64_t da = make_desc(sA + k * 256, lbo, sbo);
        const uint64_t db = make_desc(sB + k * 256, lbo, sbo);
        const uint32_t acc = k > 0 ? 1u : 0u;
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n}\n" ::"r"(tmem_base + acc_col),
            "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
            : "memory");
      }
      // a commit after every batch (nobody waits on these barriers): does the commit itself cost tensor time?
      if (commit_every) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_batch[it & 3])) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    mbar_wait(&bar_mma, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
    stop_s = 1;
  } else if (TMA && warp == 4 + LDW) {
    // one thread streams 20 KB tiles from global (L2-resident) memory into a spare shared-memory region, back to
    // back, like the gallery ring of the select kernel
    if ((tid & 31) == 0) {
      unsigned char* dst = smem + (128 + 256) * KB;
      uint32_t ph = 0;
      long long n = 0;
      while (stop_s == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_cp)), "r"(20480u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(gsrc + (size_t)blockIdx.x * 20480 + (n & 7) * 20480 * 148), "r"(20480u), "r"(smem_u32(&bar_cp)) : "memory");
        mbar_wait(&bar_cp, ph);
        ph ^= 1u;
        ++n;
      }
      out[296 + blockIdx.x] = n;
    }
  } else if (SPIN > 0 && warp > 4 + LDW && warp <= 4 + LDW + SPIN) {
    // spinning threads: compare-and-swap on a shared word that never becomes free (the MMA lock / barrier polls)
    if ((tid & 31) == 0) {
      long long n = 0;
      while (stop_s == 0) {
        if (atomicCAS(&spin_word, 0u, 1u) == 0u) break;
        ++n;
      }
      if (warp == 5 + LDW) out[444 + blockIdx.x] = n;
    }
  } else if (warp >= 4 && warp < 4 + LDW) {
    // LDW warps keep reading (packed, 64 columns per instruction) the LAST 128 columns of TMEM, which the MMAs
    // of the NACC < 512 / N configurations never touch: does the read-out slow the MMA down?
    uint32_t sink = 0;
    long long n = 0;
    while (stop_s == 0) {
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 384 + (n & 1) * 64;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) sink ^= r[j];
      ++n;
    }
    if (sink == 0x12345u) out[0] = n;
    if ((tid & 31) == 0 && warp == 4) out[148 + blockIdx.x] = n;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

template <int N, int NACC, int KBLK, int LDW = 0, int TMA = 0, int SPIN = 0>
void run(const char* name, int commit_every = 0) {
  long long* d;
  const int ctas = 148, iters = 2000;
  CK(cudaMalloc(&d, 4 * ctas * 8));
  CK(cudaMemset(d, 0, 4 * ctas * 8));
  unsigned char* gsrc;
  CK(cudaMalloc(&gsrc, (size_t)20480 * 148 * 8));
  CK(cudaMemset(gsrc, 1, (size_t)20480 * 148 * 8));
  const int smem = 384 * KB + 20480;
  CK(cudaFuncSetAttribute(probe<N, NACC, KBLK, LDW, TMA, SPIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int rep = 0; rep < 2; ++rep) {
    probe<N, NACC, KBLK, LDW, TMA, SPIN><<<ctas, 128 + 32 * LDW + 32 + 32 * (1 + SPIN), smem>>>(iters, d, gsrc, commit_every);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> h(4 * ctas);
  CK(cudaMemcpy(h.data(), d, 4 * ctas * 8, cudaMemcpyDeviceToHost));
  long long mx = 0, mn = 1LL << 60;
  for (int i = 0; i < ctas; ++i) {
    const long long v = h[i]; mx = v > mx ? v : mx; mn = v < mn ? v : mn; }
  const double mmas = (double)iters * KBLK;
  printf("%-34s clk/MMA min %.1f max %.1f   MAC/clk/SM %.0f (of 8192)\n", name, mn / mmas, mx / mmas,
         128.0 * N * 32 * mmas / mx);
  if (LDW > 0) printf("    with %d warps reading TMEM: %.1f clk per packed x32 load and warp\n", LDW, (double)mx / h[ctas]);
  if (TMA) printf("    with a bulk-copy stream: %.0f clk per 20 KB tile (%.1f B/clk/SM)\n", (double)mx / h[2 * ctas], 20480.0 * h[2 * ctas] / mx);
  if (SPIN) printf("    with %d spinning threads: %.0f clk per CAS\n", SPIN, (double)mx / h[3 * ctas]);
  cudaFree(d);
  cudaFree(gsrc);
}

int main() {
  run<128, 4, 5>("N=128, 4 acc, 5 k-blocks");
  run<128, 4, 5>("N=128, 4 acc, 5 k-blocks, commit per batch", 1);
  run<128, 4, 1>("N=128, 4 acc, 1 k-block, commit per batch", 1);
  run<256, 2, 5>("N=256, 2 acc, 5 k-blocks, commit per batch", 1);
  run<64, 8, 5>("N=64, 8 acc, 5 k-blocks");
  run<128, 3, 5, 16, 0, 0>("N=128 + LDTM x16");
  run<128, 3, 5, 16, 0, 0>("N=128 + LDTM x16, commit per batch", 1);
  run<128, 3, 3, 16, 0, 0>("N=128, 3 k-blocks + LDTM x16, commit per batch", 1);
  run<128, 3, 5, 4, 0, 0>("N=128 + LDTM x4, commit per batch", 1);
  run<128, 4, 5, 0, 1, 0>("N=128 + bulk copies, commit per batch", 1);
  run<128, 4, 5, 0, 0, 1>("N=128 + 1 CAS spinner, commit per batch", 1);
  run<128, 3, 5, 16, 1, 1>("N=128 + LDTM x16 + copies + spinner");
  run<128, 3, 5, 16, 1, 1>("N=128 + LDTM x16 + copies + spinner, commit per batch", 1);
  run<128, 0, 5, 0, 1, 1>("IDLE tensor pipe: copies + 1 spinner");
  run<128, 0, 5, 16, 0, 0>("IDLE tensor pipe: LDTM x16");
  return 0;
}
