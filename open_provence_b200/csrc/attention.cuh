// Varlen self-attention over packed sequences (RoPE already applied to q,k):
//   out[i] = softmax_j(q_i.k_j / 8  over allowed j) . v_j,   allowed = same sequence and
//   (global layer: all j) | (local layer: |i - j| <= half_window)     (HF:175-194, masking_utils.py:121-131)
// qkv is [T, 3H] with q | k | v column thirds, 64 columns per head (HF:280-282); out is [T, H].
//
//  * attention_mma_kernel   bf16 operands, flash-style online softmax in registers (warp-shuffle row
//                           reductions), QK^T and PV on tensor cores, K/V tiles double-buffered with
//                           cp.async; local layers only visit the key tiles that intersect the band.
//  * attention_simt_kernel  fp32 (or bf16) reference-precision kernel for the parity mode: one warp per
//                           query, lanes over keys, expf.
#pragma once

#include <math_constants.h>

#include "common.cuh"

namespace opv {

constexpr int kAttBlockM = 64;
constexpr int kAttBlockN = 64;
constexpr int kAttThreads = 128;

// [64 rows][64 bf16] tile, 128 B rows; 16 B chunk c of row r lives at chunk c ^ (r & 7)
__device__ __forceinline__ uint32_t att_off(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }

__global__ void __launch_bounds__(kAttThreads)
attention_mma_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                     const int32_t* __restrict__ cu_seqlens, const int H, const int half_window) {
  __shared__ __align__(128) uint8_t sQ[kAttBlockM * 128];
  __shared__ __align__(128) uint8_t sK[2][kAttBlockN * 128];
  __shared__ __align__(128) uint8_t sV[2][kAttBlockN * 128];

  const int seq = blockIdx.z, head = blockIdx.y;
  const int begin = cu_seqlens[seq];
  const int n = cu_seqlens[seq + 1] - begin;
  const int q0 = blockIdx.x * kAttBlockM;
  if (q0 >= n) return;
  const bool global = half_window < 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t ld = 3 * static_cast<int64_t>(H);
  const __nv_bfloat16* qbase = qkv + static_cast<int64_t>(begin) * ld + head * 64;
  const __nv_bfloat16* kbase = qbase + H;
  const __nv_bfloat16* vbase = qbase + 2 * H;

  int kt_lo = 0, kt_hi = (n + kAttBlockN - 1) / kAttBlockN;
  if (!global) {
    const int lo = max(0, q0 - half_window);
    const int hi = min(n, q0 + kAttBlockM + half_window);  // exclusive
    kt_lo = lo / kAttBlockN;
    kt_hi = (hi + kAttBlockN - 1) / kAttBlockN;
  }

  auto load_kv = [&](int kt, int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * kAttThreads;
      const int r = idx >> 3, c = idx & 7;
      const int key = kt * kAttBlockN + r;
      const bool ok = key < n;
      const int64_t roff = static_cast<int64_t>(ok ? key : n - 1) * ld + c * 8;
      cp_async_16(sK[buf] + att_off(r, c), kbase + roff, ok);
      cp_async_16(sV[buf] + att_off(r, c), vbase + roff, ok);
    }
  };

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * kAttThreads;
    const int r = idx >> 3, c = idx & 7;
    const bool ok = q0 + r < n;
    cp_async_16(sQ + att_off(r, c), qbase + static_cast<int64_t>(ok ? q0 + r : n - 1) * ld + c * 8, ok);
  }
  load_kv(kt_lo, 0);
  cp_async_commit();

  const float scale_log2 = 0.125f * 1.44269504088896340736f;  // head_dim^-0.5 * log2(e), d = 64
  const int r0 = q0 + warp * 16 + (lane >> 2);                // this thread's two query rows
  const int r1 = r0 + 8;
  float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, l0 = 0.f, l1 = 0.f;
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  uint32_t qf[4][4];

  for (int kt = kt_lo; kt < kt_hi; ++kt) {
    const int buf = (kt - kt_lo) & 1;
    if (kt + 1 < kt_hi) {
      load_kv(kt + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    if (kt == kt_lo) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int chunk = kk * 2 + (lane >> 4);
        ldmatrix_x4(qf[kk], smem_u32(sQ + att_off(row, chunk)));
      }
    }

    // S = Q K^T (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t kb[4];
        const int row = np * 16 + (lane >> 4) * 8 + (lane & 7);
        const int chunk = kk * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(kb, smem_u32(sK[buf] + att_off(row, chunk)));
        mma_bf16_16816(s[2 * np], qf[kk], kb[0], kb[1]);
        mma_bf16_16816(s[2 * np + 1], qf[kk], kb[2], kb[3]);
      }
    }

    // mask + online softmax (fp32, base-2)
    float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kt * kAttBlockN + j * 8 + (lane & 3) * 2 + (e & 1);
        const int qrow = (e < 2) ? r0 : r1;
        const bool ok = key < n && (global || abs(qrow - key) <= half_window);
        s[j][e] = ok ? s[j][e] * scale_log2 : -CUDART_INF_F;
      }
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float base0 = (mn0 == -CUDART_INF_F) ? 0.f : mn0;  // fully masked so far: keep exp2 finite
    const float base1 = (mn1 == -CUDART_INF_F) ? 0.f : mn1;
    const float corr0 = exp2f(m0 - base0), corr1 = exp2f(m1 - base1);
    m0 = mn0, m1 = mn1;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = exp2f(s[j][0] - base0), s[j][1] = exp2f(s[j][1] - base0);
      s[j][2] = exp2f(s[j][2] - base1), s[j][3] = exp2f(s[j][3] - base1);
      sum0 += s[j][0] + s[j][1];
      sum1 += s[j][2] + s[j][3];
      o[j][0] *= corr0, o[j][1] *= corr0, o[j][2] *= corr1, o[j][3] *= corr1;
    }
    l0 = l0 * corr0 + sum0;  // per-thread partial row sums; reduced over the 4-lane group at the end
    l1 = l1 * corr1 + sum1;

    // O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t vb[4];
        const int row = kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
        const int chunk = dp * 2 + (lane >> 4);
        ldmatrix_x4_trans(vb, smem_u32(sV[buf] + att_off(row, chunk)));
        mma_bf16_16816(o[2 * dp], pa, vb[0], vb[1]);
        mma_bf16_16816(o[2 * dp + 1], pa, vb[2], vb[3]);
      }
    }
    __syncthreads();  // all warps done with buf before the next iteration's prefetch overwrites it
  }

  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
  __nv_bfloat16* obase = out + static_cast<int64_t>(begin) * H + head * 64 + (lane & 3) * 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (r0 < n)
      *reinterpret_cast<uint32_t*>(obase + static_cast<int64_t>(r0) * H + j * 8) =
          pack_bf16x2(o[j][0] * inv0, o[j][1] * inv0);
    if (r1 < n)
      *reinterpret_cast<uint32_t*>(obase + static_cast<int64_t>(r1) * H + j * 8) =
          pack_bf16x2(o[j][2] * inv1, o[j][3] * inv1);
  }
}

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
