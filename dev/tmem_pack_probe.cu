// Development probe: does tcgen05.ld ... .pack::16b on s32 accumulators return the LOW 16 bits of two adjacent
// columns in one register?  (one tcgen05.mma kind::i8 tile, M=128, N=256, as in tc_probe.cu)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tmem_pack_probe tmem_pack_probe.cu && ./tmem_pack_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra.uni WAIT_DONE;\nbra.uni WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__host__ __device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;          // version = 1 (Blackwell)
  return d;                        // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (0)
}

template <int KB>
__global__ void __launch_bounds__(128) probe(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  int8_t* sA = (int8_t*)smem;                 // 128 x KB
  int8_t* sB = (int8_t*)(smem + 128 * KB);    // 256 x KB
  __shared__ __align__(8) uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  if (tid == 0) {
    mbar_expect_tx(&bar_load, 128 * KB + 256 * KB);
    bulk_g2s(sA, A, 128 * KB, &bar_load);
    bulk_g2s(sB, B, 256 * KB, &bar_load);
    mbar_wait(&bar_load, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (32u << 17) | (8u << 24);  // S32 acc, s8 x s8, N=256, M=128
    const uint32_t sbo = 8 * KB, lbo = 128;
#pragma unroll
    for (int k = 0; k < KB / 32; ++k) {
      const uint64_t da = make_desc(smem_u32(sA) + k * 256, lbo, sbo);
      const uint64_t db = make_desc(smem_u32(sB) + k * 256, lbo, sbo);
      const uint32_t acc = k > 0 ? 1u : 0u;
      asm volatile(
          "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n}\n" ::"r"(tmem_base),
          "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
  }
  mbar_wait(&bar_mma, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: warp w reads TMEM lanes 32w..32w+31; thread = one row of D; 64 columns per load, packed 2 x 16 bit
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < 256; c0 += 64) {
    uint32_t r[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) D[row * 128 + c0 / 2 + j] = (int)r[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

static size_t tiled(int r, int k, int KB) { return (size_t)(r / 8) * (8 * KB) + (size_t)(k / 16) * 128 + (r % 8) * 16 + (k % 16); }

template <int KB>
int run() {
  std::vector<int8_t> a(128 * KB), b(256 * KB), at(128 * KB), bt(256 * KB);
  srand(1 + KB);
  for (auto& v : a) v = (rand() & 1) ? 1 : -1;
  for (auto& v : b) v = (rand() & 1) ? 1 : -1;
  for (int r = 0; r < 128; ++r) for (int k = 0; k < KB; ++k) at[tiled(r, k, KB)] = a[r * KB + k];
  for (int r = 0; r < 256; ++r) for (int k = 0; k < KB; ++k) bt[tiled(r, k, KB)] = b[r * KB + k];
  int8_t *dA, *dB; int* dD;
  CK(cudaMalloc(&dA, at.size())); CK(cudaMalloc(&dB, bt.size())); CK(cudaMalloc(&dD, 128 * 256 * 4));
  CK(cudaMemcpy(dA, at.data(), at.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, bt.data(), bt.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0x7f, 128 * 256 * 4));
  const int smem = 384 * KB;
  CK(cudaFuncSetAttribute(probe<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe<KB><<<1, 128, smem>>>(dA, dB, dD);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<int> d(128 * 128);
  CK(cudaMemcpy(d.data(), dD, d.size() * 4, cudaMemcpyDeviceToHost));
  long bad = 0;
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 128; ++j) {
      int ref[2];
      for (int h = 0; h < 2; ++h) {
        ref[h] = 0;
        for (int k = 0; k < KB; ++k) ref[h] += (int)a[i * KB + k] * (int)b[(2 * j + h) * KB + k];
      }
      const unsigned want = ((unsigned)ref[0] & 0xffffu) | ((unsigned)ref[1] << 16);
      if (want != (unsigned)d[i * 128 + j]) { if (bad < 8) printf("KB=%d mismatch P[%d][%d] = %08x, want %08x (cols %d, %d)\n", KB, i, j, d[i * 128 + j], want, ref[0], ref[1]); ++bad; }
    }
  printf("KB=%d: %ld mismatches of %d\n", KB, bad, 128 * 128);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return bad != 0;
}

int main() {
  int rc = 0;
  rc |= run<32>();
  rc |= run<64>();
  rc |= run<128>();
  printf(rc ? "PROBE FAILED\n" : "PROBE OK\n");
  return rc;
}
