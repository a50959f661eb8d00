// The attention softmax inner loop in isolation (no MMA / TMEM / barriers): per "block" each thread does
// 64 FMNMX3 + 128 x (FFMA, MUFU.EX2, FADD) + 64 F2FP over 128 register-resident scores.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_loop softmax_loop.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ unsigned pack(float lo, float hi) { unsigned r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

template <int VARIANT>  // 0 = full, 1 = no MUFU (FMUL instead), 2 = no max, 3 = no pack
__global__ void __launch_bounds__(256, 1) k(const float* in, unsigned* out, long long* clk, int iters) {
  float sr[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) sr[i] = in[i * 256 + threadIdx.x];
  float m = 0.f, l = 0.f;
  unsigned acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float mx0 = -1e30f, mx1 = -1e30f, mx2 = -1e30f, mx3 = -1e30f;
    if (VARIANT != 2) {
#pragma unroll
      for (int c = 0; c < 128; c += 8) {
        mx0 = fmax3(mx0, sr[c], sr[c + 1]); mx1 = fmax3(mx1, sr[c + 2], sr[c + 3]);
        mx2 = fmax3(mx2, sr[c + 4], sr[c + 5]); mx3 = fmax3(mx3, sr[c + 6], sr[c + 7]);
      }
    }
    m = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * 0.18f + l * 1e-30f;
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
    for (int c = 0; c < 128; c += 4) {
      const float xa = fmaf(sr[c], 0.18f, -m), xb = fmaf(sr[c + 1], 0.18f, -m), xc = fmaf(sr[c + 2], 0.18f, -m), xd = fmaf(sr[c + 3], 0.18f, -m);
      float a, b, e, f;
      if (VARIANT == 1) { a = xa * 0.5f, b = xb * 0.5f, e = xc * 0.5f, f = xd * 0.5f; }
      else { a = ex2(xa), b = ex2(xb), e = ex2(xc), f = ex2(xd); }
      s0 += a, s1 += b, s2 += e, s3 += f;
      if (VARIANT != 3) acc ^= pack(a, b) + pack(e, f);
    }
    l = l * 0.5f + (s0 + s1) + (s2 + s3);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(l);
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// variant with the kernel's register budget: 128 scores refreshed per block, the 64 packed probabilities kept until
// the end of the block (then written to shared memory), at most 216 registers per thread
__global__ void __maxnreg__(216) k216(const float* in, unsigned* out, long long* clk, int iters) {
  __shared__ uint4 sp[256 * 4];
  float sr[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) sr[i] = in[i * 256 + threadIdx.x];
  float m = 0.f, l = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float mx0 = -1e30f, mx1 = -1e30f, mx2 = -1e30f, mx3 = -1e30f;
#pragma unroll
    for (int c = 0; c < 128; c += 8) {
      mx0 = fmax3(mx0, sr[c], sr[c + 1]); mx1 = fmax3(mx1, sr[c + 2], sr[c + 3]);
      mx2 = fmax3(mx2, sr[c + 4], sr[c + 5]); mx3 = fmax3(mx3, sr[c + 6], sr[c + 7]);
    }
    m = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * 0.18f + l * 1e-30f;
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    unsigned pr[64];
#pragma unroll
    for (int c = 0; c < 64; c += 2) {
      const float a = ex2(fmaf(sr[2 * c], 0.18f, -m)), b = ex2(fmaf(sr[2 * c + 1], 0.18f, -m));
      const float e = ex2(fmaf(sr[2 * c + 2], 0.18f, -m)), f = ex2(fmaf(sr[2 * c + 3], 0.18f, -m));
      s0 += a, s1 += b, s2 += e, s3 += f;
      pr[c] = pack(a, b); pr[c + 1] = pack(e, f);
    }
    l = l * 0.5f + (s0 + s1) + (s2 + s3);
    // all 64 probabilities leave together, like the single tcgen05.st of the kernel
    asm volatile("" :: "r"(pr[0]), "r"(pr[63]));
#pragma unroll
    for (int g = 0; g < 16; ++g) sp[(threadIdx.x & 255) * 4 + (g & 3)] = make_uint4(pr[4 * g], pr[4 * g + 1], pr[4 * g + 2], pr[4 * g + 3]);
    // next block's scores depend on this block (cannot be hoisted)
#pragma unroll
    for (int i = 0; i < 128; i += 16) sr[i] += l * 1e-30f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(l) + sp[threadIdx.x].x;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

void run216(int threads) {
  int sms = 148, iters = 2000;
  float* in; unsigned* out; long long* clk;
  cudaMalloc(&in, 32768 * 4); cudaMemset(in, 0, 32768 * 4);
  cudaMalloc(&out, sms * 256 * 4); cudaMalloc(&clk, sms * 8);
  k216<<<sms, threads>>>(in, out, clk, iters); k216<<<sms, threads>>>(in, out, clk, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
  printf("%-28s threads=%d %7.0f clk per 128-score block  %s\n", "216 regs, P kept per block", threads, avg / iters, cudaGetErrorString(cudaGetLastError()));
}

template <int V>
void run(const char* name, int threads = 256) {
  int sms = 148, iters = 2000;
  float* in; unsigned* out; long long* clk;
  cudaMalloc(&in, 32768 * 4); cudaMemset(in, 0, 32768 * 4);
  cudaMalloc(&out, sms * 256 * 4); cudaMalloc(&clk, sms * 8);
  k<V><<<sms, threads>>>(in, out, clk, iters); k<V><<<sms, threads>>>(in, out, clk, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
  printf("%-28s threads=%d %7.0f clk per 128-score block  %s\n", name, threads, avg / iters, cudaGetErrorString(cudaGetLastError()));
}
int main() { run216(128); run216(256); run<0>("full", 128); run<1>("no MUFU (FMUL)", 128); run<0>("full"); run<1>("no MUFU (FMUL)"); run<2>("no max"); run<3>("no pack"); return 0; }
