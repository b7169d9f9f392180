// MUFU.EX2 issue rate: f32 vs packed f16x2 (results per clock per SM).  nvcc -arch=sm_100a -O3 -o ex2_rate ex2_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>

template <int MODE>
__global__ void k(float* out, int iters, long long* cycles) {
  float a[8];
  uint32_t h[8];
  for (int i = 0; i < 8; ++i) {
    a[i] = -0.001f * (threadIdx.x + i);
    h[i] = 0xB000B000u + threadIdx.x + i;   // two small negative halves
  }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      else asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
    }
  }
  long long t1 = clock64();
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  for (int mode = 0; mode < 2; ++mode) {
    for (int threads = 256; threads <= 1024; threads *= 2) {
      if (mode == 0) k<0><<<148, threads>>>(out, iters, cyc); else k<1><<<148, threads>>>(out, iters, cyc);
      cudaDeviceSynchronize();
      long long h;
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const double insts = double(iters) * 8 * threads;   // thread-instructions per SM
      printf("%s threads/SM=%4d: %.2f thread-instr/clk/SM = %.2f results/clk/SM\n", mode ? "ex2.f16x2" : "ex2.f32  ", threads,
             insts / h, insts * (mode ? 2 : 1) / h);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
