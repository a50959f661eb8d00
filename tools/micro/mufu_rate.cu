// Per-SM issue rate of the special-function / min-max instructions the attention softmax leans on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu && ./mufu_rate
// One 1024-thread block per SM, 8 independent chains per thread, clock64() around the loop.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024, 1) rate_kernel(float* out, long long* clk, int iters) {
  float v[8];
  unsigned u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = -0.001f * (threadIdx.x + i), u[i] = 0x3c003c00u + threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
      if (MODE == 3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(v[(i + 1) & 7]), "f"(v[(i + 2) & 7]));
      if (MODE == 4) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(v[(i + 1) & 7]), "f"(v[(i + 2) & 7]));
      if (MODE == 5) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(v[i]), "f"(v[(i + 1) & 7])); v[i] = __uint_as_float(u[i]); }
      if (MODE == 7) asm volatile("add.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(v[(i + 1) & 7]));
      if (MODE == 8) asm volatile("mul.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(v[(i + 1) & 7]));
      if (MODE == 9) asm volatile("shl.b32 %0, %0, 3; add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
      if (MODE == 10) asm volatile("selp.f32 %0, %0, %1, p;" : "+f"(v[i]) : "f"(v[(i + 1) & 7]));
      if (MODE == 6) asm volatile("max.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(v[(i + 1) & 7]));
    }
  }
  long long t1 = clock64();
  __syncthreads();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_instr) {
  int sms = 148, iters = 4096;
  float* out; long long* clk;
  cudaMalloc(&out, sms * 1024 * sizeof(float));
  cudaMalloc(&clk, sms * sizeof(long long));
  rate_kernel<MODE><<<sms, 1024>>>(out, clk, iters);
  rate_kernel<MODE><<<sms, 1024>>>(out, clk, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms; ++i) avg += h[i];
  avg /= sms;
  double thread_instr = 1024.0 * iters * 8;
  printf("%-28s %7.1f thread-instr/clk/SM  (%6.1f results/clk/SM)  err=%s\n", name, thread_instr / avg,
         thread_instr * per_instr / avg, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(clk);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.bf16x2", 2);
  run<2>("ex2.approx.f16x2", 2);
  run<3>("max.f32 (3-input, FMNMX3)", 1);
  run<6>("max.f32 (2-input)", 1);
  run<4>("fma.rn.f32", 1);
  run<5>("cvt.rn.bf16x2.f32 (F2FP)", 2);
  run<7>("add.f32", 1);
  run<8>("mul.f32", 1);
  run<9>("shl+add (int)", 1);
  return 0;
}
