// TMEM -> register (tcgen05.ld) and register -> TMEM (tcgen05.st) bandwidth per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../open_provence_b200/csrc -o tmem_bw tmem_bw.cu
// CTAS_PER_SM blocks of 4 or 8 warps per SM, each warp loops over its 32 lanes x COLS columns.
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "common.cuh"
#include "tmem_ldst.cuh"
using namespace opv;

template <int MODE>  // 0: ld x64, 1: ld 2x64 (one wait), 2: st x64, 3: ld x32 (32 columns, float)
__global__ void tmem_kernel(float* out, long long* clk, int iters, int cols_alloc) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, cols_alloc); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t r[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) r[i] = threadIdx.x + i;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) { uint32_t (&q)[64] = reinterpret_cast<uint32_t (&)[64]>(r); tmem_ld_32x32b_x64(base + (it & 1) * 64, q); acc += q[0] ^ q[63]; }
    if (MODE == 1) { tmem_ld_32x32b_2x64(base, r); acc += r[0] ^ r[127]; }
    if (MODE == 2) { uint32_t (&q)[64] = reinterpret_cast<uint32_t (&)[64]>(r); q[0] += it; tmem_st_32x32b_x64(base + (it & 1) * 64, q); }
    if (MODE == 3) { float (&q)[32] = reinterpret_cast<float (&)[32]>(r); tmem_ld_32x32(base + (it & 3) * 32, q); acc += r[0] ^ r[31]; }
  }
  long long t1 = clock64();
  tc_fence_before(); __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + r[5];
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  if (warp == 0) tmem_dealloc(slot, cols_alloc);
}

template <int MODE>
void run(const char* name, int bytes_per_thread_iter, int threads, int ctas_per_sm) {
  int sms = 148, iters = 2000, blocks = sms * ctas_per_sm;
  float* out; long long* clk;
  cudaMalloc(&out, blocks * threads * sizeof(float));
  cudaMalloc(&clk, blocks * sizeof(long long));
  for (int rep = 0; rep < 2; ++rep) tmem_kernel<MODE><<<blocks, threads>>>(out, clk, iters, 256);
  cudaError_t e = cudaDeviceSynchronize();
  long long* h = new long long[blocks];
  cudaMemcpy(h, clk, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
  double bytes_per_sm = double(bytes_per_thread_iter) * threads * iters * ctas_per_sm;
  printf("%-22s threads=%3d ctas/SM=%d: %7.1f B/clk/SM  (%.0f clk per iteration)  %s\n", name, threads, ctas_per_sm,
         bytes_per_sm / avg, avg / iters, cudaGetErrorString(e));
  cudaFree(out); cudaFree(clk); delete[] h;
}

int main() {
  for (int c = 1; c <= 2; ++c) {
    run<0>("ld 32x32b.x64", 256, 128, c);
    run<1>("ld 32x32b 2x64", 512, 128, c);
    run<3>("ld 32x32b.x32", 128, 128, c);
    run<2>("st 32x32b.x64", 256, 128, c);
  }
  run<0>("ld 32x32b.x64", 256, 256, 1);
  run<1>("ld 32x32b 2x64", 512, 256, 1);
  return 0;
}
