// MUFU.EX2 latency / throughput with few warps per scheduler (the attention softmax has 2 per SMSP).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_lat mufu_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, bool CONSUME>
__global__ void k(float* out, long long* clk, int iters) {
  float v[ILP], s[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = -0.001f * (threadIdx.x + i), s[i] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (CONSUME) asm volatile("add.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(v[i]));  // consumer right behind (ILP apart)
    }
  }
  long long t1 = clock64();
  float a = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) a += v[i] + s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int ILP, bool CONSUME>
void run(int threads) {
  int sms = 148, iters = 4096;
  float* out; long long* clk;
  cudaMalloc(&out, sms * threads * sizeof(float));
  cudaMalloc(&clk, sms * sizeof(long long));
  k<ILP, CONSUME><<<sms, threads>>>(out, clk, iters);
  k<ILP, CONSUME><<<sms, threads>>>(out, clk, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
  printf("warps/SMSP=%d ILP=%2d consume=%d: %6.1f clk per MUFU per warp, %5.2f MUFU lanes/clk/SM\n", threads / 128, ILP,
         (int)CONSUME, avg / (double(iters) * ILP), double(threads) * iters * ILP / avg);
  cudaFree(out); cudaFree(clk);
}

int main() {
  run<1, false>(128); run<2, false>(128); run<4, false>(128); run<8, false>(128);
  run<1, false>(256); run<2, false>(256); run<4, false>(256); run<8, false>(256);
  run<2, true>(256); run<4, true>(256); run<8, true>(256); run<16, true>(256);
  return 0;
}
