// Reference-precision varlen self-attention over packed sequences (RoPE already applied to q,k), fp32 parity mode:
//   out[i] = softmax_j(q_i.k_j / 8  over allowed j) . v_j,   allowed = same sequence and
//   (global layer: all j) | (local layer: |i - j| <= half_window)     (HF:175-194, masking_utils.py:121-131)
// qkv is [T, 3H] with q | k | v column thirds, 64 columns per head (HF:280-282); out is [T, H].
// The bf16 product path is attention_tcgen05*.cuh; this kernel only backs dtype = fp32 (the 1e-5 cross-check).
#pragma once

#include <math_constants.h>

#include "common.cuh"


namespace opv {

// --------------------------------------------------------------------------------------------------
// Reference-precision kernel: one warp per (query, head); each lane walks keys lane, lane+32, ...
// with its own online-softmax state, merged across the warp at the end.
// --------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
attention_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, const int32_t* __restrict__ cu_seqlens,
                      const int H, const int half_window) {
  __shared__ float sq[4][64];
  const int seq = blockIdx.z, head = blockIdx.y;
  const int begin = cu_seqlens[seq];
  const int n = cu_seqlens[seq + 1] - begin;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int qi = blockIdx.x * 4 + warp;
  if (qi >= n) return;  // warp-uniform
  const int64_t ld = 3 * static_cast<int64_t>(H);
  const T* qptr = qkv + (static_cast<int64_t>(begin) + qi) * ld + head * 64;
  const T* kbase = qkv + static_cast<int64_t>(begin) * ld + H + head * 64;
  const T* vbase = kbase + H;
  sq[warp][lane] = OperandCast<T>::to_float(qptr[lane]) * 0.125f;
  sq[warp][lane + 32] = OperandCast<T>::to_float(qptr[lane + 32]) * 0.125f;
  __syncwarp();

  const int lo = half_window < 0 ? 0 : max(0, qi - half_window);
  const int hi = half_window < 0 ? n : min(n, qi + half_window + 1);
  float m = -CUDART_INF_F, l = 0.f;
  float o[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) o[d] = 0.f;
  for (int j = lo + lane; j < hi; j += 32) {
    const T* kr = kbase + static_cast<int64_t>(j) * ld;
    const T* vr = vbase + static_cast<int64_t>(j) * ld;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 64; ++d) s = fmaf(sq[warp][d], OperandCast<T>::to_float(kr[d]), s);
    const float mn = fmaxf(m, s);
    const float corr = expf(m - mn);  // m = -inf on the first key -> 0
    const float p = expf(s - mn);
    l = l * corr + p;
#pragma unroll
    for (int d = 0; d < 64; ++d) o[d] = o[d] * corr + p * OperandCast<T>::to_float(vr[d]);
    m = mn;
  }
  const float mw = warp_max(m);  // finite: every query sees at least itself
  const float sc = (m == -CUDART_INF_F) ? 0.f : expf(m - mw);
  const float lw = warp_sum(l * sc);
  const float inv = 1.0f / lw;
  T* optr = out + (static_cast<int64_t>(begin) + qi) * H + head * 64;
#pragma unroll
  for (int d = 0; d < 64; ++d) {
    const float v = warp_sum(o[d] * sc);
    if (lane == (d & 31)) optr[d] = OperandCast<T>::from_float(v * inv);
  }
}

}  // namespace opv
