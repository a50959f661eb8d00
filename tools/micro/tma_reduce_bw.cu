// Micro-benchmark: throughput of cp.reduce.async.bulk.tensor (.add.f32) vs plain TMA store for a
// [M, 512] fp32 tensor written in [128 x 32] tiles (the residual-stream update of the Wo GEMMs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_reduce_bw tma_reduce_bw.cu && ./tma_reduce_bw
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>  // 0 = store, 1 = reduce add
__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap tm, int M, int N) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* buf = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < 128 * 32 * 2; i += 128) buf[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int tiles_m = M / 128, tiles_n = N / 32;
  if (threadIdx.x == 0) {
    int it = 0;
    for (int t = blockIdx.x; t < tiles_m * tiles_n; t += gridDim.x, ++it) {
      const int m = t / tiles_n, n = t % tiles_n;
      const uint32_t src = smem_u32(buf + (it & 1) * 128 * 32);
      if (MODE == 0)
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tm), "r"(src), "r"(n * 32), "r"(m * 128) : "memory");
      else
        asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tm), "r"(src), "r"(n * 32), "r"(m * 128) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main() {
  const int M = 131072, N = 512;
  float* d;
  CK(cudaMalloc(&d, (size_t)M * N * 4));
  CK(cudaMemset(d, 0, (size_t)M * N * 4));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)N * 4};
  cuuint32_t box[2] = {32, 128}, el[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int mode = 0; mode < 2; ++mode) {
    for (int grid : {148, 296}) {
      float best = 1e9;
      for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        if (mode == 0) k<0><<<grid, 128, 32768>>>(tm, M, N); else k<1><<<grid, 128, 32768>>>(tm, M, N);
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
      }
      printf("%s grid=%d: %.3f ms, %.1f GB/s payload\n", mode ? "reduce.add.f32" : "store", grid, best, (double)M * N * 4 / best / 1e6);
    }
  }
  std::vector<float> h(8);
  CK(cudaMemcpy(h.data(), d, 32, cudaMemcpyDeviceToHost));
  printf("d[0]=%f (expect 10 reduce passes + stores of 1.0)\n", h[0]);
  return 0;
}
