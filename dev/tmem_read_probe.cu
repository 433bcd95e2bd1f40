// Development probe: TMEM -> register read bandwidth (tcgen05.ld 32x32b.x32, plain and .pack::16b) with 1..4 warps
// per scheduler reading their lane quadrants in a loop.  Prints TMEM bytes per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tmem_read_probe tmem_read_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

#define OUT32 "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
#define REGS "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}"

template <bool PACK>
__global__ void __launch_bounds__(512) probe(int iters, unsigned long long* out, uint32_t* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t r[32], acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint32_t taddr = base + (uint32_t)(((warp >> 2) * 128 + (i & 1) * 64) & 511);
    if (PACK)
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 " REGS ", [%32];\n" : OUT32 : "r"(taddr) : "memory");
    else
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " REGS ", [%32];\n" : OUT32 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) acc ^= r[j];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 0x12345u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512u) : "memory");
}

int main() {
  unsigned long long* out; uint32_t* sink;
  CK(cudaMalloc(&out, 8 * 148)); CK(cudaMalloc(&sink, 4));
  const int iters = 4000;
  CK(cudaFuncSetAttribute(probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CK(cudaFuncSetAttribute(probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  for (int pack = 0; pack < 2; ++pack)
    for (int warps = 4; warps <= 16; warps += 4) {
      for (int rep = 0; rep < 2; ++rep) {
        if (pack) probe<true><<<148, warps * 32, 100 * 1024>>>(iters, out, sink); else probe<false><<<148, warps * 32, 100 * 1024>>>(iters, out, sink);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
      }
      unsigned long long clk = 0; CK(cudaMemcpy(&clk, out, 8, cudaMemcpyDeviceToHost));
      // TMEM cells read per load: 32 lanes x (pack ? 64 : 32) columns x 4 bytes
      const double bytes = (double)warps * iters * 32.0 * (pack ? 64 : 32) * 4.0;
      printf("%s warps/SM %2d (%d per scheduler): %.1f clk per load per warp, %.1f TMEM bytes/clk/SM, %.1f columns x lanes / clk / SM\n",
             pack ? "x32.pack::16b" : "x32          ", warps, warps / 4, (double)clk / iters, bytes / clk, bytes / 4 / clk);
    }
  return 0;
}
