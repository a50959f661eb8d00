// fp32 FFMA GEMM for the parity mode (OPV_DTYPE_F32): C (+)= A[M,K] . W[N,K]^T with fp32 operands and
// fp32 accumulation, so the engine can be compared with the fp32 reference forward at 1e-5.
// 64x64 tile, BK = 16, 256 threads, 4x4 outputs per thread.  Not the throughput path.
#pragma once

#include "common.cuh"

namespace opv {

template <bool ACCUMULATE>
__global__ void __launch_bounds__(256)
gemm_f32_simt_kernel(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ C, const int M,
                     const int N, const int K, const int64_t ldc) {
  __shared__ float As[16][64 + 4];
  __shared__ float Ws[16][64 + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = static_cast<int64_t>(blockIdx.y) * 64;
  const int n0 = blockIdx.x * 64;
  const int lr = tid >> 2;        // tile row loaded by this thread
  const int lk = (tid & 3) * 4;   // first of its 4 k values

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += 16) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + lr < M) a = *reinterpret_cast<const float4*>(A + (m0 + lr) * K + k0 + lk);
    const float4 w = *reinterpret_cast<const float4*>(W + static_cast<int64_t>(n0 + lr) * K + k0 + lk);
    As[lk + 0][lr] = a.x, As[lk + 1][lr] = a.y, As[lk + 2][lr] = a.z, As[lk + 3][lr] = a.w;
    Ws[lk + 0][lr] = w.x, Ws[lk + 1][lr] = w.y, Ws[lk + 2][lr] = w.z, Ws[lk + 3][lr] = w.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float av[4], wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i], wv[i] = Ws[k][tx * 4 + i];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = m0 + ty * 4 + i;
    if (row < M) {
      float4* dst = reinterpret_cast<float4*>(C + row * ldc + n0 + tx * 4);
      float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      if (ACCUMULATE) {
        const float4 o = *dst;
        v.x += o.x, v.y += o.y, v.z += o.z, v.w += o.w;
      }
      *dst = v;
    }
  }
}

}  // namespace opv
