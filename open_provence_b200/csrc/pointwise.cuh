// HBM-bound row kernels: embedding gather + LayerNorm, LayerNorm, final LayerNorm fused with the
// token-mask (pruning) head, the CLS rerank head, unfused RoPE / GeGLU (fp32 parity mode and the
// unfused bf16 debug path), positions, and the score-conversion / per-sentence prune kernels.
// One warp owns one token row; rows are read once with 16 B loads and written once.
#pragma once

#include "common.cuh"

namespace opv {

constexpr int kRowWarps = 8;  // warps (rows) per CTA for the row kernels

// ---------------------------------------------------------------------------------------------
// row load / store helpers (VEC float4 per lane, H = VEC * 128)
// ---------------------------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void load_row_f32(const float* __restrict__ src, int lane, float4 (&v)[VEC]) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) v[i] = *reinterpret_cast<const float4*>(src + (i * 32 + lane) * 4);
}
template <int VEC>
__device__ __forceinline__ void load_row_bf16(const __nv_bfloat16* __restrict__ src, int lane, float4 (&v)[VEC]) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const uint2 u = *reinterpret_cast<const uint2*>(src + (i * 32 + lane) * 4);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    v[i] = make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
  }
}
__device__ __forceinline__ void store4(float* dst, float4 v) { *reinterpret_cast<float4*>(dst) = v; }
__device__ __forceinline__ void store4(__nv_bfloat16* dst, float4 v) {
  uint2 u;
  u.x = pack_bf16x2(v.x, v.y);
  u.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(dst) = u;
}

// LayerNorm statistics over one row held by a warp (nn.LayerNorm: biased variance, HF:60,321,323,432).
template <int VEC>
__device__ __forceinline__ void row_norm_stats(const float4 (&v)[VEC], float eps, float& mean, float& rstd) {
  constexpr float inv_h = 1.0f / static_cast<float>(VEC * 128);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  mean = warp_sum(s) * inv_h;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  rstd = 1.0f / sqrtf(warp_sum(q) * inv_h + eps);
}
template <int VEC>
__device__ __forceinline__ void row_normalize(float4 (&v)[VEC], const float* __restrict__ w, int lane, float mean,
                                              float rstd) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(w + (i * 32 + lane) * 4));
    v[i].x = (v[i].x - mean) * rstd * g.x;
    v[i].y = (v[i].y - mean) * rstd * g.y;
    v[i].z = (v[i].z - mean) * rstd * g.z;
    v[i].w = (v[i].w - mean) * rstd * g.w;
  }
}

// ---------------------------------------------------------------------------------------------
// K1: h = LN(E[ids]) (fp32 residual stream), x = cast(h) (layer 0 has no attn_norm, HF:318-319)
// ---------------------------------------------------------------------------------------------
template <typename EmbT, typename OutT, int VEC>
__global__ void __launch_bounds__(kRowWarps * 32)
embed_ln_kernel(const int32_t* __restrict__ ids, const EmbT* __restrict__ emb, const float* __restrict__ w,
                float* __restrict__ h, OutT* __restrict__ x, const int64_t M, const int vocab, const float eps) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel in the stream
  constexpr int H = VEC * 128;
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
  if (row >= M) return;
  int id = ids[row];
  id = min(max(id, 0), vocab - 1);  // ids are validated on the host; never read out of bounds
  float4 v[VEC];
  if constexpr (sizeof(EmbT) == 2) {
    load_row_bf16<VEC>(reinterpret_cast<const __nv_bfloat16*>(emb) + static_cast<int64_t>(id) * H, lane, v);
  } else {
    load_row_f32<VEC>(reinterpret_cast<const float*>(emb) + static_cast<int64_t>(id) * H, lane, v);
  }
  float mean, rstd;
  row_norm_stats<VEC>(v, eps, mean, rstd);
  row_normalize<VEC>(v, w, lane, mean, rstd);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    store4(h + row * H + (i * 32 + lane) * 4, v[i]);
    store4(x + row * H + (i * 32 + lane) * 4, v[i]);
  }
}

// K2/K8: x = LN(h) * w
template <typename OutT, int VEC>
__global__ void __launch_bounds__(kRowWarps * 32)
layernorm_kernel(const float* __restrict__ h, const float* __restrict__ w, OutT* __restrict__ x, const int64_t M,
                 const float eps) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel in the stream
  constexpr int H = VEC * 128;
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
  if (row >= M) return;
  float4 v[VEC];
  load_row_f32<VEC>(h + row * H, lane, v);
  float mean, rstd;
  row_norm_stats<VEC>(v, eps, mean, rstd);
  row_normalize<VEC>(v, w, lane, mean, rstd);
#pragma unroll
  for (int i = 0; i < VEC; ++i) store4(x + row * H + (i * 32 + lane) * 4, v[i]);
}

// K11 + K12: prune_logits[t] = LN(h[t]; final_norm) . Wp^T + bp  (HF:488 + standalone:446-447).
// The final hidden state is never written to HBM.
template <int VEC>
__global__ void __launch_bounds__(kRowWarps * 32)
final_ln_prune_kernel(const float* __restrict__ h, const float* __restrict__ w, const float* __restrict__ wp,
                      const float* __restrict__ bp, float* __restrict__ logits, const int64_t M, const float eps) {
  pdl_launch_dependents();
  pdl_wait();  // inputs come from the previous kernel in the stream
  constexpr int H = VEC * 128;
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
  if (row >= M) return;
  float4 v[VEC];
  load_row_f32<VEC>(h + row * H, lane, v);
  float mean, rstd;
  row_norm_stats<VEC>(v, eps, mean, rstd);
  row_normalize<VEC>(v, w, lane, mean, rstd);
  float d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(wp + (i * 32 + lane) * 4));
    const float4 b = __ldg(reinterpret_cast<const float4*>(wp + H + (i * 32 + lane) * 4));
    d0 += v[i].x * a.x + v[i].y * a.y + v[i].z * a.z + v[i].w * a.w;
    d1 += v[i].x * b.x + v[i].y * b.y + v[i].z * b.z + v[i].w * b.w;
  }
  d0 = warp_sum(d0);
  d1 = warp_sum(d1);
  if (lane == 0) {
    logits[row * 2 + 0] = d0 + bp[0];
    logits[row * 2 + 1] = d1 + bp[1];
  }
}

// K13: rank_logits[s] = classifier(LN(gelu(dense(pooled[s])); head.norm)) (HF:621-634,493-502) with
// pooled[s] = LN(h[cls_s]; final_norm) for "cls" pooling (mean pooling: further down).  Three small launches: the first version ran everything in one CTA per sequence and spent
// 337 us per step re-reading the [H, H] dense matrix once per sequence with one dependent load at a time.
//   stage 1 (one warp per sequence)      cls[s]  = LN(h[first token of s]; final_norm)
//   stage 2 (8 features x 8 sequences)   y[s][j] = gelu(cls[s] . dense[j])   dense row held in registers,
//                                                                            the 8 cls rows shared through smem
//   stage 3 (one warp per sequence)      rank[s] = LN(y[s]; head.norm) . classifier^T + bias
constexpr int kRankSeqTile = 8;  // 8 x H x 4 B <= 32 KB of shared memory for H <= 1024
constexpr int kRankFeatTile = 8;  // = warps per CTA of stage 2

__global__ void __launch_bounds__(32)
rank_head_cls_ln_kernel(const float* __restrict__ h, const int32_t* __restrict__ cu_seqlens,
                        const float* __restrict__ final_norm, float* __restrict__ cls, const int H, const float eps,
                        const int64_t M) {
  const int s = blockIdx.x, lane = threadIdx.x;
  const int64_t first = cu_seqlens[s];
  const float* row = h + (first < M ? first : M - 1) * H;  // an empty trailing sequence must not read past the buffer
  const float inv_h = 1.0f / static_cast<float>(H);
  float part = 0.f;
  for (int i = lane; i < H; i += 32) part += row[i];
  const float mean = warp_sum(part) * inv_h;
  part = 0.f;
  for (int i = lane; i < H; i += 32) {
    const float d = row[i] - mean;
    part += d * d;
  }
  const float rstd = 1.0f / sqrtf(warp_sum(part) * inv_h + eps);
  for (int i = lane; i < H; i += 32) cls[static_cast<int64_t>(s) * H + i] = (row[i] - mean) * rstd * final_norm[i];
}

// classifier_pooling = "mean" (HF:623-630): pooled[s] = (1 / n_s) * sum over the sequence's tokens of
// LN(h[t]; final_norm), in place of the CLS row.  Two launches, summation order fixed by the sequence alone (so the
// result does not depend on what else is in the batch):
//   partial  grid (chunks, sequences): the 8 warps of a CTA walk a chunk of kPoolChunkRows rows (warp w takes rows
//            w, w + 8, ...), each lane accumulating its 4 * VEC columns; the warps' sums are added in warp order
//   finish   one CTA per sequence adds the chunk sums in chunk order and divides by n_s
// A chunk's slot in `partial` is begin / kPoolChunkRows + s + chunk: unique per (s, chunk) because
// floor(n / C) + 1 >= ceil(n / C), and below T / C + n_seqs + 1.
constexpr int kPoolChunkRows = 256;

template <int VEC>  // H = VEC * 128
__global__ void __launch_bounds__(kRowWarps * 32)
rank_head_mean_partial_kernel(const float* __restrict__ h, const int32_t* __restrict__ cu_seqlens,
                              const float* __restrict__ final_norm, float* __restrict__ partial, const float eps) {
  constexpr int H = VEC * 128;
  __shared__ float4 sm_acc[kRowWarps][VEC * 32];
  const int s = blockIdx.y, chunk = blockIdx.x;
  const int begin = cu_seqlens[s], n = cu_seqlens[s + 1] - begin;
  const int r0 = chunk * kPoolChunkRows;
  if (r0 >= n) return;
  const int r1 = min(n, r0 + kPoolChunkRows);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = r0 + warp; r < r1; r += kRowWarps) {
    float4 v[VEC];
    load_row_f32<VEC>(h + static_cast<int64_t>(begin + r) * H, lane, v);
    float mean, rstd;
    row_norm_stats<VEC>(v, eps, mean, rstd);
    row_normalize<VEC>(v, final_norm, lane, mean, rstd);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      acc[i].x += v[i].x, acc[i].y += v[i].y, acc[i].z += v[i].z, acc[i].w += v[i].w;
    }
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) sm_acc[warp][i * 32 + lane] = acc[i];
  __syncthreads();
  float* out = partial + (static_cast<int64_t>(begin / kPoolChunkRows) + s + chunk) * H;
  for (int i = threadIdx.x; i < VEC * 32; i += blockDim.x) {
    float4 t = sm_acc[0][i];
#pragma unroll
    for (int w = 1; w < kRowWarps; ++w) {
      const float4 u = sm_acc[w][i];
      t.x += u.x, t.y += u.y, t.z += u.z, t.w += u.w;
    }
    *reinterpret_cast<float4*>(out + i * 4) = t;
  }
}

__global__ void __launch_bounds__(128)
rank_head_mean_finish_kernel(const float* __restrict__ partial, const int32_t* __restrict__ cu_seqlens,
                             float* __restrict__ pooled, const int H) {
  const int s = blockIdx.x;
  const int begin = cu_seqlens[s], n = cu_seqlens[s + 1] - begin;
  const int chunks = (n + kPoolChunkRows - 1) / kPoolChunkRows;
  const float* src = partial + (static_cast<int64_t>(begin / kPoolChunkRows) + s) * H;
  const float inv_n = 1.0f / static_cast<float>(n > 0 ? n : 1);
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    float t = 0.f;
    for (int c = 0; c < chunks; ++c) t += src[static_cast<int64_t>(c) * H + i];
    pooled[static_cast<int64_t>(s) * H + i] = t * inv_n;
  }
}

template <int VEC>  // H = VEC * 128
__global__ void __launch_bounds__(kRankFeatTile * 32)
rank_head_dense_kernel(const float* __restrict__ cls, const float* __restrict__ dense, float* __restrict__ y,
                       const int n_seqs) {
  constexpr int H = VEC * 128;
  extern __shared__ float sm_cls[];  // [kRankSeqTile][H]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s0 = blockIdx.y * kRankSeqTile;
  const int ns = min(kRankSeqTile, n_seqs - s0);
  for (int i = threadIdx.x; i < ns * (H / 4); i += blockDim.x)
    reinterpret_cast<float4*>(sm_cls)[i] = reinterpret_cast<const float4*>(cls + static_cast<int64_t>(s0) * H)[i];
  const int j = blockIdx.x * kRankFeatTile + warp;  // dense has no bias (classifier_bias = False)
  float4 w[VEC];
  load_row_f32<VEC>(dense + static_cast<int64_t>(j) * H, lane, w);
  __syncthreads();
  for (int s = 0; s < ns; ++s) {
    const float* c = sm_cls + s * H;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float4 v = *reinterpret_cast<const float4*>(c + (i * 32 + lane) * 4);
      acc += (v.x * w[i].x + v.y * w[i].y) + (v.z * w[i].z + v.w * w[i].w);
    }
    acc = warp_sum(acc);
    if (lane == 0) y[static_cast<int64_t>(s0 + s) * H + j] = gelu_erf(acc);
  }
}

__global__ void __launch_bounds__(32)
rank_head_out_kernel(const float* __restrict__ y, const float* __restrict__ head_norm, const float* __restrict__ cls_w,
                     const float* __restrict__ cls_b, float* __restrict__ rank_logits, const int H,
                     const int num_labels, const float eps) {
  const int s = blockIdx.x, lane = threadIdx.x;
  const float* row = y + static_cast<int64_t>(s) * H;
  const float inv_h = 1.0f / static_cast<float>(H);
  float part = 0.f;
  for (int i = lane; i < H; i += 32) part += row[i];
  const float mean = warp_sum(part) * inv_h;
  part = 0.f;
  for (int i = lane; i < H; i += 32) {
    const float d = row[i] - mean;
    part += d * d;
  }
  const float rstd = 1.0f / sqrtf(warp_sum(part) * inv_h + eps);
  for (int l = 0; l < num_labels; ++l) {
    const float* wr = cls_w + static_cast<int64_t>(l) * H;
    float acc = 0.f;
    for (int i = lane; i < H; i += 32) acc = fmaf((row[i] - mean) * rstd * head_norm[i], wr[i], acc);
    acc = warp_sum(acc);
    if (lane == 0) rank_logits[static_cast<int64_t>(s) * num_labels + l] = acc + cls_b[l];
  }
}

// ---------------------------------------------------------------------------------------------
// positions: pos[t] = t - cu_seqlens[seq(t)]  (RoPE restarts at 0 for every packed sequence)
// ---------------------------------------------------------------------------------------------
__global__ void positions_kernel(const int32_t* __restrict__ cu_seqlens, int32_t* __restrict__ pos) {
  const int s = blockIdx.x;
  const int begin = cu_seqlens[s], end = cu_seqlens[s + 1];
  for (int i = begin + threadIdx.x; i < end; i += blockDim.x) pos[i] = i - begin;
}

// The forward's version: cu_seqlens is a DEVICE array the host cannot look at without a synchronisation, so it is made
// memory-safe here instead of trusted.  cu_clean[s] = clamp(cu_seqlens[s], 0, T) is what every later kernel reads: all
// row indices derived from it stay inside [0, T) (a non-increasing pair is an empty sequence to them), positions stay
// inside the RoPE table.  status[s] != 0 marks a boundary pair that was not 0 <= begin <= end <= T with
// end - begin <= max_seqlen, first begin == 0, last end == T; opv_forward_status() counts them on request.
__global__ void positions_checked_kernel(const int32_t* __restrict__ cu_seqlens, int32_t* __restrict__ cu_clean,
                                         int32_t* __restrict__ status, int32_t* __restrict__ pos, const int64_t T,
                                         const int max_seqlen, const int max_pos) {
  const int s = blockIdx.x, last = gridDim.x - 1;
  const int64_t b0 = cu_seqlens[s], e0 = cu_seqlens[s + 1];
  const int begin = static_cast<int>(b0 < 0 ? 0 : (b0 > T ? T : b0));
  const int end = static_cast<int>(e0 < 0 ? 0 : (e0 > T ? T : e0));
  if (threadIdx.x == 0) {
    cu_clean[s] = begin;
    if (s == last) cu_clean[s + 1] = end;
    status[s] = (b0 != begin || e0 != end || e0 < b0 || e0 - b0 > max_seqlen || (s == 0 && b0 != 0) ||
                 (s == last && e0 != T))
                    ? 1
                    : 0;
  }
  for (int i = begin + threadIdx.x; i < end; i += blockDim.x) pos[i] = min(i - begin, max_pos - 1);
}

// ---------------------------------------------------------------------------------------------
// unfused RoPE (in place on the q,k thirds of qkv [M, 3H]) and GeGLU -- fp32 parity mode
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void rope_inplace_kernel(T* __restrict__ qkv, const int32_t* __restrict__ pos,
                                    const float* __restrict__ cos_t, const float* __restrict__ sin_t, const int64_t M,
                                    const int H, const int max_pos = 0) {
  const int heads2 = 2 * (H / 64);  // q heads then k heads: contiguous first 2H columns
  const int64_t total = M * heads2 * 32;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(idx & 31);
    const int64_t rh = idx >> 5;
    const int hd = static_cast<int>(rh % heads2);
    const int64_t row = rh / heads2;
    int p = pos[row];
    if (max_pos > 0) p = min(max(p, 0), max_pos - 1);  // rows no sequence covers hold no position
    const float c = cos_t[static_cast<int64_t>(p) * 32 + i], s = sin_t[static_cast<int64_t>(p) * 32 + i];
    T* base = qkv + row * (3 * static_cast<int64_t>(H)) + hd * 64;
    const float a = OperandCast<T>::to_float(base[i]), b = OperandCast<T>::to_float(base[i + 32]);
    base[i] = OperandCast<T>::from_float(a * c - b * s);
    base[i + 32] = OperandCast<T>::from_float(b * c + a * s);
  }
}

template <typename T>
__global__ void geglu_kernel(const T* __restrict__ u, T* __restrict__ act, const int64_t M, const int I) {
  const int64_t total = M * I;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = idx / I;
    const int j = static_cast<int>(idx - row * I);
    const float a = OperandCast<T>::to_float(u[row * 2 * I + j]);
    const float g = OperandCast<T>::to_float(u[row * 2 * I + I + j]);
    act[idx] = OperandCast<T>::from_float(gelu_erf(a) * g);
  }
}

// ---------------------------------------------------------------------------------------------
// K14 + K15: score conversion and the per-sentence threshold / prune
// ---------------------------------------------------------------------------------------------
// keep-prob of a token = softmax(l)[1] evaluated like the reference's fp32 softmax (standalone:2918)
__device__ __forceinline__ float keep_prob(float l0, float l1) {
  const float m = fmaxf(l0, l1);
  const float e0 = expf(l0 - m), e1 = expf(l1 - m);
  return e1 / (e0 + e1);
}

// token-level keep probabilities (encoder.py:429-430): p[t] = softmax(prune_logits[t])[1]
__global__ void token_keep_prob_kernel(const float* __restrict__ logits, const int64_t n_tokens, float* __restrict__ prob) {
  const float2* l2 = reinterpret_cast<const float2*>(logits);
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < n_tokens;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float2 l = __ldg(l2 + t);
    prob[t] = keep_prob(l.x, l.y);
  }
}

// one warp per fragment: coalesced read of the fragment's [start, end) logits, warp-shuffle sum
__global__ void __launch_bounds__(kRowWarps * 32)
fragment_mean_kernel(const float* __restrict__ logits, const int64_t n_tokens, const int32_t* __restrict__ ranges,
                     const int n_frags, float* __restrict__ frag_mean) {
  const int lane = threadIdx.x & 31;
  const int f = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (f >= n_frags) return;
  int64_t start = ranges[2 * f], end = ranges[2 * f + 1];
  start = max(static_cast<int64_t>(0), min(start, n_tokens));
  end = max(start, min(end, n_tokens));
  float s = 0.f;
  const float2* l2 = reinterpret_cast<const float2*>(logits);
  for (int64_t t = start + lane; t < end; t += 32) {
    const float2 l = __ldg(l2 + t);
    s += keep_prob(l.x, l.y);
  }
  s = warp_sum(s);
  if (lane == 0) frag_mean[f] = (end <= start) ? 1.0f : s / static_cast<float>(end - start);  // standalone:3081
}

__global__ void rank_score_kernel(const float* __restrict__ rank_logits, const int n_seqs, const int num_labels,
                                  float* __restrict__ score) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_seqs) score[s] = 1.0f / (1.0f + expf(-rank_logits[static_cast<int64_t>(s) * num_labels]));
}

// one thread per sentence: unweighted fp64 mean of its fragment means (standalone:3116-3134)
__global__ void sentence_prune_kernel(const float* __restrict__ frag_mean, const int32_t* __restrict__ sent_offsets,
                                      const int32_t* __restrict__ sent_frag_index, const int n_sents,
                                      const double threshold, const double guard, double* __restrict__ sent_prob,
                                      uint8_t* __restrict__ keep, uint8_t* __restrict__ near) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_sents) return;
  const int b = sent_offsets[s], e = sent_offsets[s + 1];
  double acc = 0.0;
  for (int k = b; k < e; ++k) acc += static_cast<double>(frag_mean[sent_frag_index[k]]);
  double p = (e > b) ? acc / static_cast<double>(e - b) : 0.0;
  p = fmax(0.0, fmin(p, 1.0));
  sent_prob[s] = p;
  keep[s] = p > threshold ? 1 : 0;
  near[s] = fabs(p - threshold) <= guard ? 1 : 0;
}

// x (fp32) -> three bf16 planes with x = hi + mid + lo up to 2^-24 |x| (OPV_DTYPE_F32_TC: fp32 GEMMs as six bf16 passes
// on the tcgen05 kernel).  Plane stride = n elements.
__global__ void split3_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ planes, const int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = x[i];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(hi);
    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
    const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
    planes[i] = hi;
    planes[n + i] = mid;
    planes[2 * n + i] = lo;
  }
}

}  // namespace opv
