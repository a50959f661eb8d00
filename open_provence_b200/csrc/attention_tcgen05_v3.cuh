// Varlen self-attention on tcgen05, second softmax organisation: TWO threads per query row.
//
// Same tile loop, barriers, TMA producer and MMA issuer as attention_tcgen05_kernel (attention_tcgen05.cuh);
// what changes is who evaluates the softmax.  There one thread owned a whole row (128 scores per key block)
// and each SM scheduler hosted 2 softmax warps; the clock64 timeline showed a warp needing ~2850 cycles per
// block for work that takes 2111 in isolation (tools/micro/softmax_loop.cu), with MUFU.EX2 -- the bounding
// unit at 16/clk/SM -- only ~70 % busy, because two in-order warps per scheduler cannot cover each other's
// TMEM round trips, barrier hand-offs and row-max phase.  Here warps w and w + 4 share TMEM lane quarter w:
// thread (w, lane) handles keys [64 * (w / 4), +64) of row 32 * (w % 4) + lane.  Every scheduler hosts 4 softmax
// warps with half-length dependency chains, and a thread needs ~70 registers (32 scores + 16 packed
// probabilities at a time: the scores are read from TMEM twice, once for the row max and once for the
// exponentials), so no setmaxnreg juggling is needed.
//   - row max: each thread reduces its 64 scores, the two halves meet through shared memory + a 64-thread barrier
//   - row sum: kept per half, added in the epilogue
//   - lazy O rescale and the final O / l store: each half owns 32 of the 64 output columns
#pragma once

#include <math_constants.h>

#include "attention_tcgen05.cuh"

namespace opv {

constexpr int kFa3Threads = 384;  // warps 0..7 softmax, 8 TMA producer, 9 MMA issuer, 10 TMEM allocation, 11 idle

struct Fa3SmemLayout {
  static constexpr int kQ = 0;
  static constexpr int kK = kQ + kFaTileBytes;
  static constexpr int kV = kK + kFaKvStages * kFaTileBytes;
  static constexpr int kExchange = kV + kFaKvStages * kFaTileBytes;  // float [2 parities][2 halves][128 rows]
  static constexpr int kBars = kExchange + 2 * 2 * 128 * 4;
  static constexpr int kTotal = kBars + 256 + 1024;
};

__global__ void __launch_bounds__(kFa3Threads, 2)
attention_tcgen05_v3_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out,
                            const int32_t* __restrict__ cu_seqlens, const int H, const int half_window,
                            const int n_seqs, const int tiles_per_seq, const int pdl_late) {
  using L = Fa3SmemLayout;
  const int heads = H / 64;
  const int total_tiles = n_seqs * heads * tiles_per_seq;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024 B alignment (128 B swizzle) established in the SHARED address space: the shared-window address of a __shared__
  // symbol is a compile-time constant.  Rounding the GENERIC pointer made ptxas rebuild the window base (S2UR
  // SR_SWINHI / SR_CgaCtaId + uniform arithmetic) in front of every barrier operation (profiles/r2_attention_notes.md).
  const uint32_t raw0 = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw0 + 1023u) & ~1023u) - raw0);
  uint8_t* sQ = smem + L::kQ;
  uint8_t* sK = smem + L::kK;
  uint8_t* sV = smem + L::kV;
  float* xchg = reinterpret_cast<float*>(smem + L::kExchange);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + kFaKvStages;
  uint64_t* v_full = k_empty + kFaKvStages;
  uint64_t* v_empty = v_full + kFaKvStages;
  uint64_t* s_full = v_empty + kFaKvStages;  // S(i) complete in TMEM             (tcgen05.commit)
  uint64_t* s_empty = s_full + 1;            // S(i) no longer needed             (8 warp arrivals)
  uint64_t* p_full = s_empty + 1;            // P(i) written (+ O rescaled)       (8 warp arrivals)
  uint64_t* pv_done = p_full + 1;            // O += P(i).V(i) complete           (tcgen05.commit)
  uint64_t* q_empty = pv_done + 1;           // last S of a tile complete         (tcgen05.commit)
  uint64_t* o_empty = q_empty + 1;           // O of a tile read by the epilogue  (8 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const bool global = half_window < 0;

  struct Tile {
    int begin, n, q0, head, key_base, nb;
  };
  auto decode = [&](const int t, Tile& tile) -> bool {
    const int qt = t % tiles_per_seq;
    const int sh = t / tiles_per_seq;
    const int seq = sh / heads;
    tile.head = sh - seq * heads;
    tile.begin = cu_seqlens[seq];
    tile.n = cu_seqlens[seq + 1] - tile.begin;
    tile.q0 = qt * kFaBlockM;
    if (tile.q0 >= tile.n) return false;
    tile.key_base = global ? 0 : tile.q0 - half_window;
    const int key_end = global ? tile.n : min(tile.n, tile.q0 + kFaBlockM + half_window);
    tile.nb = (key_end - tile.key_base + kFaBlockN - 1) / kFaBlockN;
    return true;
  };

  if (warp == 8 && lane == 0) tma_prefetch_desc(&tm_qkv);
  if (warp == 9 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kFaKvStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 8);
    mbar_init(p_full, 8);
    mbar_init(pv_done, 1);
    mbar_init(q_empty, 1);
    mbar_init(o_empty, 8);
    fence_mbar_init();
  }
  if (warp == 10) {
    tmem_alloc(tmem_slot, kFaTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (!pdl_late) pdl_launch_dependents();
  pdl_wait();  // the prologue above overlapped the previous kernel; qkv is visible from here on
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform (single UTCHMMA per MMA)

  if (warp == 8) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      uint32_t tiles_done = 0, kc = 0, vc = 0;
      Tile tl;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        if (!decode(t, tl)) continue;
        if (tiles_done > 0) mbar_wait_spin(q_empty, (tiles_done - 1) & 1);
        mbar_expect_tx(q_full, kFaTileBytes);
        tma_load_2d(sQ, &tm_qkv, q_full, tl.head * 64, tl.begin + tl.q0);
        for (int i = 0; i <= tl.nb; ++i) {  // consumption order of the MMA warp: K0, K1, V0, K2, V1, ...
          if (i < tl.nb) {
            const uint32_t st = kc % kFaKvStages;
            mbar_wait_spin(&k_empty[st], ((kc / kFaKvStages) & 1) ^ 1);
            mbar_expect_tx(&k_full[st], kFaTileBytes);
            tma_load_2d(sK + st * kFaTileBytes, &tm_qkv, &k_full[st], H + tl.head * 64,
                        tl.begin + tl.key_base + i * kFaBlockN);
            ++kc;
          }
          if (i >= 1) {
            const uint32_t st = vc % kFaKvStages;
            mbar_wait_spin(&v_empty[st], ((vc / kFaKvStages) & 1) ^ 1);
            mbar_expect_tx(&v_full[st], kFaTileBytes);
            tma_load_2d(sV + st * kFaTileBytes, &tm_qkv, &v_full[st], 2 * H + tl.head * 64,
                        tl.begin + tl.key_base + (i - 1) * kFaBlockN);
            ++vc;
          }
        }
        ++tiles_done;
      }
    }
  } else if (warp == 9) {
    // ------------------------------ MMA issuer --------------------------------
    constexpr uint32_t idesc_s = umma_idesc_bf16_f32(kFaBlockM, kFaBlockN);  // Q.K^T: both K-major
    constexpr uint32_t idesc_o = umma_idesc_bf16_f32_bmn(kFaBlockM, 64);     // P.V: V is MN-major
    const uint32_t t_s = tmem_base, t_p = tmem_base + 128, t_o = tmem_base + 192;
    const uint32_t q_addr = smem_u32(sQ);
    uint32_t tiles_done = 0, sc = 0, pc = 0;
    auto issue_s = [&](const bool last_of_tile) {  // S = Q . K^T for the next key block
      const uint32_t st = sc % kFaKvStages;
      mbar_wait_spin(&k_full[st], (sc / kFaKvStages) & 1);
      if (sc > 0) mbar_wait_spin(s_empty, (sc - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t k_addr = smem_u32(sK + st * kFaTileBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(t_s, umma_desc_k_sw128(q_addr + k * 32), umma_desc_k_sw128(k_addr + k * 32), idesc_s,
                       k != 0 ? 1u : 0u);
        umma_commit(&k_empty[st]);
        umma_commit(s_full);
        if (last_of_tile) umma_commit(q_empty);
      }
      __syncwarp();
      ++sc;
    };
    Tile tl;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      if (!decode(t, tl)) continue;
      mbar_wait_spin(q_full, tiles_done & 1);
      issue_s(tl.nb == 1);
      for (int i = 0; i < tl.nb; ++i) {
        if (i + 1 < tl.nb) issue_s(i + 2 == tl.nb);
        const uint32_t st = pc % kFaKvStages;
        mbar_wait_spin(&v_full[st], (pc / kFaKvStages) & 1);
        mbar_wait_spin(p_full, pc & 1);
        if (i == 0 && tiles_done > 0) mbar_wait_spin(o_empty, (tiles_done - 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t v_addr = smem_u32(sV + st * kFaTileBytes);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16_ts(t_o, t_p + k * 8, umma_desc_mn_sw128(v_addr + k * 2048), idesc_o, (i | k) != 0 ? 1u : 0u);
          umma_commit(&v_empty[st]);
          umma_commit(pv_done);
        }
        __syncwarp();
        ++pc;
      }
      ++tiles_done;
    }
  } else if (warp < 8) {
    // ------------------------------ softmax warps: two threads per query row ----
    const int quarter = warp & 3, half = warp >> 2;
    const int r_tile = quarter * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_s = lane_base + 64 * half;        // this thread's 64 score columns
    const uint32_t t_p = lane_base + 128 + 32 * half;  // its 32 packed-probability columns
    const uint32_t t_o = lane_base + 192 + 32 * half;  // the 32 output columns it rescales / stores
    const float scale_log2 = 0.125f * 1.44269504088896340736f;
    const int pair_bar = 1 + quarter;  // named barrier of warps {quarter, quarter + 4}
    // row max / row sum exchange between the two halves of a row; slots alternate with the block parity so that a
    // thread that runs ahead cannot overwrite a value its partner has not read yet
    float* const my_slots = xchg + half * 128 + r_tile;
    const float* const other_slots = xchg + (half ^ 1) * 128 + r_tile;
    uint32_t bc = 0;
    Tile tl;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      if (!decode(t, tl)) continue;
      const int n = tl.n, q0 = tl.q0;
      const int row = q0 + r_tile;
      float m_run = -CUDART_INF_F, l_run = 0.f;
      // keys this row may attend to: [k_lo, k_lo + k_span]; keys every one of the warp's 32 rows sees: [all_lo, all_hi];
      // keys at least one of them sees: [any_lo, any_hi]
      const int row_first = q0 + quarter * 32, row_last = row_first + 31;
      const int k_lo = global ? 0 : max(row - half_window, 0);
      const int k_hi = global ? n - 1 : min(row + half_window, n - 1);
      const uint32_t k_span = static_cast<uint32_t>(k_hi - k_lo);
      const int all_lo = global ? 0 : max(row_last - half_window, 0);
      const int all_hi = global ? n - 1 : min(row_first + half_window, n - 1);
      const int any_lo = global ? 0 : max(row_first - half_window, 0);
      const int any_hi = global ? n - 1 : min(row_last + half_window, n - 1);

      for (int i = 0; i < tl.nb; ++i, ++bc) {
        const int key0 = tl.key_base + i * kFaBlockN + 64 * half;  // first key of this thread's 64
        int kind[2];  // per 32-key chunk, warp-uniform: 0 = no row sees it, 1 = every row sees all of it, 2 = mixed
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c_lo = key0 + 32 * q, c_hi = c_lo + 31;
          kind[q] = (c_hi < any_lo || c_lo > any_hi) ? 0 : ((c_lo >= all_lo && c_hi <= all_hi) ? 1 : 2);
        }
        mbar_wait_spin(s_full, bc & 1);
        tc_fence_after();

        // ---- pass 1: row max over this thread's 64 scores
        float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (kind[q] == 0) continue;
          uint32_t sr[32];
          tmem_ld_32x32_raw(t_s + 32 * q, sr);
          if (kind[q] == 2) {
            const int d = key0 + 32 * q - k_lo;
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (static_cast<uint32_t>(d + c) > k_span) sr[c] = 0xff800000u;  // -inf
          }
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            mx0 = fmax3(mx0, __uint_as_float(sr[c + 0]), __uint_as_float(sr[c + 1]));
            mx1 = fmax3(mx1, __uint_as_float(sr[c + 2]), __uint_as_float(sr[c + 3]));
          }
        }
        my_slots[(bc & 1) * 256] = fmaxf(mx0, mx1);
        named_bar_sync(pair_bar, 64);
        const float mx = fmaxf(fmaxf(mx0, mx1), other_slots[(bc & 1) * 256]);
        const float m_cand = fmaxf(m_run, mx * scale_log2);
        const bool upd = (m_cand - m_run) > kFaRescaleThreshold;  // false when both are -inf (NaN); same in both halves
        const float m_new = upd ? m_cand : m_run;
        const float corr = upd ? ex2_approx(m_run - m_new) : 1.0f;
        const float m_use = (m_new == -CUDART_INF_F) ? 0.f : m_new;
        m_run = m_new;

        if (i > 0) {
          mbar_wait_spin(pv_done, (bc - 1) & 1);  // O holds blocks < i and the P buffer is free again
          tc_fence_after();
          if (__any_sync(0xffffffffu, upd)) {
            uint32_t orr[32];
            tmem_ld_32x32_raw(t_o, orr);
#pragma unroll
            for (int c = 0; c < 32; ++c) orr[c] = __float_as_uint(__uint_as_float(orr[c]) * corr);
            tmem_st_32x32b_x32(t_o, orr);
          }
        }

        // ---- pass 2: exponentials, 32 keys at a time; the packed probabilities leave in 16-column stores
        // packed fp32 pairs (FFMA2 / FADD2, sm_100): half an issue slot per score for scale-and-shift and row sum
        float2 acc01 = make_float2(0.f, 0.f), acc23 = make_float2(0.f, 0.f);
        const float2 sc2 = make_float2(scale_log2, scale_log2), nm2 = make_float2(-m_use, -m_use);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint32_t pq[16];
          if (kind[q] == 0) {
#pragma unroll
            for (int c = 0; c < 16; ++c) pq[c] = 0u;
          } else {
            uint32_t sr[32];
            tmem_ld_32x32_raw(t_s + 32 * q, sr);
            if (kind[q] == 2) {
              const int d = key0 + 32 * q - k_lo;
#pragma unroll
              for (int c = 0; c < 32; ++c)
                if (static_cast<uint32_t>(d + c) > k_span) sr[c] = 0xff800000u;
            }
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
              const float2 x01 = __ffma2_rn(make_float2(__uint_as_float(sr[2 * c]), __uint_as_float(sr[2 * c + 1])), sc2, nm2);
              const float2 x23 = __ffma2_rn(make_float2(__uint_as_float(sr[2 * c + 2]), __uint_as_float(sr[2 * c + 3])), sc2, nm2);
              const float2 p01 = make_float2(ex2_approx(x01.x), ex2_approx(x01.y));
              const float2 p23 = make_float2(ex2_approx(x23.x), ex2_approx(x23.y));
              acc01 = __fadd2_rn(acc01, p01);
              acc23 = __fadd2_rn(acc23, p23);
              pq[c] = pack_bf16x2(p01.x, p01.y);
              pq[c + 1] = pack_bf16x2(p23.x, p23.y);
            }
          }
          tmem_st_32x32b_x16_nowait(t_p + 16 * q, pq);
        }
        // every score of this block has been read twice by now: the MMA warp may overwrite S with S(i+1)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty);
        const float2 acc = __fadd2_rn(acc01, acc23);
        l_run = l_run * corr + (acc.x + acc.y);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }

      // epilogue: O / l -> bf16 -> out[begin + row, head*64 + 32*half : +32]
      my_slots[(bc & 1) * 256] = l_run;  // parity of the NEXT block: last used two blocks ago
      mbar_wait_spin(pv_done, (bc - 1) & 1);
      tc_fence_after();
      uint32_t orr[32];
      tmem_ld_32x32_raw(t_o, orr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);  // the next tile's first P.V may overwrite O
      named_bar_sync(pair_bar, 64);
      const float l_total = l_run + other_slots[(bc & 1) * 256];
      named_bar_sync(pair_bar, 64);  // both halves have read the sums before the next tile's max exchange
      if (row < n) {
        const float inv = 1.0f / l_total;
        uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<int64_t>(tl.begin) + row) * H + tl.head * 64 + 32 * half);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(orr[8 * g + 0]) * inv, __uint_as_float(orr[8 * g + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(orr[8 * g + 2]) * inv, __uint_as_float(orr[8 * g + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(orr[8 * g + 4]) * inv, __uint_as_float(orr[8 * g + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(orr[8 * g + 6]) * inv, __uint_as_float(orr[8 * g + 7]) * inv);
          dst[g] = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) tmem_dealloc(tmem_base, kFaTmemCols);
}

}  // namespace opv
