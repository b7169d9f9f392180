// Microbenchmark: cycles per tcgen05.mma (SS mode, kind::f16, cta_group::1) as a function of M and N.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "../../consistencytta_b200/csrc/ctta_ptx.cuh"
using namespace ctta;

// acc_mode: 1 = every MMA accumulates into the same TMEM tile; 0 = switch tile every 4 MMAs; 2 = round-robin over 4 tiles
// a_shift: byte offset (multiple of 128) added to the A start address: a row-shifted view of the swizzled tile
__global__ void __launch_bounds__(128, 1) k(int m, int n, int iters, int same_acc, int a_shift, long long* cyc, long long* ns) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tslot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc(m, n, 0);
    long long t0 = clock64();
    unsigned long long g0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g0));
    for (int i = 0; i < iters; ++i) {
      const uint32_t a = base + (i & 3) * 32 + a_shift;    // 4 k-slices of a 128x64 tile
      const uint32_t b = base + 16384 + (i & 3) * 32;
      const uint32_t d = same_acc == 1 ? tm : (same_acc == 0 ? tm + ((i >> 2) & 1) * 256 : tm + (i & 3) * 128);
      umma_f16(d, umma_desc_sw128(a), umma_desc_sw128(b), idesc, 1u);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    unsigned long long g1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
    if (blockIdx.x == 0) { *cyc = t1 - t0; *ns = (long long)(g1 - g0); }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long *cyc, *ns;
  cudaMallocManaged(&cyc, 8); cudaMallocManaged(&ns, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 20000;
  for (int grid : {1, 148}) for (int m : {128}) for (int n : {32, 64, 128, 256}) for (int same : {1, 0, 2}) for (int shift : {0, 384}) {
    if (same == 2 && n > 128) continue;
    k<<<grid, 128, 64 * 1024>>>(m, n, iters, same, shift, cyc, ns);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    printf("grid %3d M %3d N %3d acc_mode %d a_shift %3d: %7.1f cycles/UMMA  %7.1f ns/UMMA  (%.2f GHz)  %.0f TFLOP/s chip-wide\n", grid, m, n, same, shift,
           (double)*cyc / iters, (double)*ns / iters, (double)*cyc / (double)*ns,
           2.0 * m * n * 16 * iters * grid / ((double)*ns * 1e-9) / 1e12);
  }
  return 0;
}
