// Varlen self-attention on the 5th-gen tensor cores (bf16 operands, fp32 softmax), head_dim = 64.
//   out[i] = softmax_j(q_i.k_j / 8 over allowed j) . v_j     (HF:175-194, masking_utils.py:121-131)
// qkv is the packed [T, 3H] output of the Wqkv GEMM (RoPE already applied): q | k | v column thirds,
// 64 columns per head (HF:280-282); out is [T, H].
//
// Work unit = one (sequence, head, 128-query tile).  The kernel is PERSISTENT: 2 CTAs per SM walk the tile
// list with a grid stride, so TMEM allocation, barrier setup and the first TMA round trip are paid once per
// CTA and the producer / MMA warps run ahead into the next tile while the softmax warps finish the current
// one (a local-attention tile is only two key blocks long: launched one CTA per tile, more than half of its
// ~11000 cycles were prologue and epilogue).  Two CTAs per SM so that one CTA's softmax overlaps the other's
// MMAs.  Warp roles (256 threads):
//   warps 0..3  softmax      ONE THREAD PER QUERY ROW (TMEM lane = row): tcgen05.ld the 128 scores of
//                            the row, thread-local max / exp2 / sum (no shuffles), write P as packed
//                            bf16 back to TMEM (tcgen05.st), rescale O in TMEM only when the running max
//                            grew by more than 2^8 (lazy rescale), final 1/l normalisation + store
//   warp 4      TMA producer Q once, then K/V 128x64 tiles (128 B swizzle) through 2-stage rings
//   warp 5      MMA issuer   S = Q.K^T (SS, 128x128x64) into TMEM cols [0,128);
//                            O += P.V  (A = P from TMEM cols [128,192), B = V MN-major from smem) into
//                            cols [192,256).  S(i+1) is issued as soon as the softmax warps have READ
//                            S(i), so it runs under softmax(i).
//   warps 6..7  TMEM alloc / idle (they donate their registers: setmaxnreg)
// Local (sliding window) layers visit only the key blocks that intersect the band |i-j| <= half_window.
#pragma once

#include <math_constants.h>

#include "common.cuh"
#include "tmem_ldst.cuh"

namespace opv {

constexpr int kFaBlockM = 128;
constexpr int kFaBlockN = 128;
constexpr int kFaThreads = 256;
constexpr int kFaKvStages = 2;
constexpr int kFaTileBytes = 128 * 64 * 2;  // one 128-row x 64-col bf16 tile
constexpr int kFaTmemCols = 256;            // S [0,128) | P [128,192) | O [192,256)
constexpr float kFaRescaleThreshold = 8.0f; // log2 domain: P stays below 2^8

template <bool P_IN_TMEM>
struct FaSmemLayout {
  static constexpr int kQ = 0;
  static constexpr int kK = kQ + kFaTileBytes;
  static constexpr int kV = kK + kFaKvStages * kFaTileBytes;
  static constexpr int kP = kV + kFaKvStages * kFaTileBytes;  // smem-P variant only: 2 x [128][64] bf16
  static constexpr int kBars = kP + (P_IN_TMEM ? 0 : 2 * kFaTileBytes);
  static constexpr int kTotal = kBars + 256 + 1024;  // + barriers + slack for the 1024 B alignment
};

template <bool P_IN_TMEM>
__global__ void __launch_bounds__(kFaThreads, 2)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out,
                         const int32_t* __restrict__ cu_seqlens, const int H, const int half_window,
                         const int n_seqs, const int tiles_per_seq, long long* __restrict__ trace,
                         const int pdl_late) {
  // trace (tools/attn_check.py only; nullptr in the product): clock64() stamps of the first tile of CTA 1, 16 slots
  // per key block: softmax warp 0 [0 s_full seen, 1 scores read, 4 pv_done seen, 5 P published], MMA warp
  // [6 S(i) issued, 7 PV(i) issued], softmax warp w < 3 [8+2w scores read, 9+2w P published]; block 0 slots
  // 14 / 15 = kernel entry / exit.
  using L = FaSmemLayout<P_IN_TMEM>;
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
  uint8_t* sP = smem + L::kP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;                 // [kFaKvStages]
  uint64_t* k_empty = k_full + kFaKvStages;    // [kFaKvStages]
  uint64_t* v_full = k_empty + kFaKvStages;    // [kFaKvStages]
  uint64_t* v_empty = v_full + kFaKvStages;    // [kFaKvStages]
  uint64_t* s_full = v_empty + kFaKvStages;    // S(i) complete in TMEM          (tcgen05.commit)
  uint64_t* s_empty = s_full + 1;              // S(i) read into registers        (4 warp arrivals)
  uint64_t* p_full = s_empty + 1;              // P(i) written (+ O rescaled)     (4 warp arrivals)
  uint64_t* pv_done = p_full + 1;              // O += P(i).V(i) complete         (tcgen05.commit)
  uint64_t* q_empty = pv_done + 1;             // last S of a tile complete: Q may be overwritten (tcgen05.commit)
  uint64_t* o_empty = q_empty + 1;             // O of a tile read by the epilogue (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const bool global = half_window < 0;
  const bool trace_cta = trace != nullptr && blockIdx.x == 1 && lane == 0;
  bool tracing = trace_cta;  // cleared after the CTA's first tile
#define OPV_FA_STAMP(blk, slot) do { if (tracing) trace[(blk) * 16 + (slot)] = clock64(); } while (0)
  if (warp == 0) OPV_FA_STAMP(0, 14);  // kernel entry

  // Tile t -> (sequence, head, query tile); consecutive t are consecutive query tiles of one (sequence, head), so
  // the CTAs running at the same time share K / V in L2.  Key blocks of 128: global layers walk [0, n); local
  // layers walk [q0 - w, q0 + 128 + w) -- the band of this query tile -- starting at an UNALIGNED key (TMA
  // zero-fills rows before the tensor, rows of the previous sequence are masked), so a tile needs 2 blocks instead
  // of the 3 that 128-aligned blocks would touch.  Every role decodes the same tiles and skips the empty ones.
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
    tile.nb = (key_end - tile.key_base + kFaBlockN - 1) / kFaBlockN;  // >= 1
    return true;
  };

  if (warp == 4 && lane == 0) tma_prefetch_desc(&tm_qkv);
  if (warp == 5 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kFaKvStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    mbar_init(q_empty, 1);
    mbar_init(o_empty, 4);
    fence_mbar_init();
  }
  if (warp == 6) {
    tmem_alloc(tmem_slot, kFaTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (!pdl_late) pdl_launch_dependents();
  pdl_wait();  // the prologue above overlapped the previous kernel; qkv is visible from here on
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform for ptxas: a per-thread TMEM address makes every tcgen05.mma an ELECT / R2UR.BROADCAST waterfall loop

  if (warp >= 4) {
    setmaxnreg_dec<40>();
    if (warp == 4) {
      // ------------------------------ TMA producer ------------------------------
      if (lane == 0) {
        uint32_t tiles_done = 0, kc = 0, vc = 0;  // running counts -> ring stage and barrier parity
        Tile tl;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
          if (!decode(t, tl)) continue;
          if (tiles_done > 0) mbar_wait_spin(q_empty, (tiles_done - 1) & 1);  // previous tile's S MMAs have read Q
          mbar_expect_tx(q_full, kFaTileBytes);
          tma_load_2d(sQ, &tm_qkv, q_full, tl.head * 64, tl.begin + tl.q0);
          // consumption order of the MMA warp: K0, K1, V0, K2, V1, ...
          for (int i = 0; i <= tl.nb; ++i) {
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
    } else if (warp == 5) {
      // ------------------------------ MMA issuer --------------------------------
      constexpr uint32_t idesc_s = umma_idesc_bf16_f32(kFaBlockM, kFaBlockN);  // Q.K^T: both K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16_f32_bmn(kFaBlockM, 64);     // P.V: V is MN-major
      const uint32_t t_s = tmem_base, t_p = tmem_base + 128, t_o = tmem_base + 192;
      const uint32_t q_addr = smem_u32(sQ);
      uint32_t tiles_done = 0, sc = 0, pc = 0;  // running counts of S / PV issues -> ring stage and barrier parity
      auto issue_s = [&](const int i, const bool last_of_tile) {  // S(i) = Q . K(i)^T
        const uint32_t st = sc % kFaKvStages;
        mbar_wait_spin(&k_full[st], (sc / kFaKvStages) & 1);
        if (sc > 0) mbar_wait_spin(s_empty, (sc - 1) & 1);  // the softmax warps have read the previous S
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
          OPV_FA_STAMP(i, 6);
        }
        __syncwarp();
        ++sc;
      };
      Tile tl;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        if (!decode(t, tl)) continue;
        mbar_wait_spin(q_full, tiles_done & 1);
        issue_s(0, tl.nb == 1);
        for (int i = 0; i < tl.nb; ++i) {
          if (i + 1 < tl.nb) issue_s(i + 1, i + 2 == tl.nb);
          const uint32_t st = pc % kFaKvStages;
          mbar_wait_spin(&v_full[st], (pc / kFaKvStages) & 1);
          mbar_wait_spin(p_full, pc & 1);
          if (i == 0 && tiles_done > 0) mbar_wait_spin(o_empty, (tiles_done - 1) & 1);  // epilogue has read the previous O
          tc_fence_after();
          if (elect_one()) {
            const uint32_t v_addr = smem_u32(sV + st * kFaTileBytes);
#pragma unroll
            for (int k = 0; k < 8; ++k) {  // 16 keys per MMA: two 8-key groups of 1024 B
              const uint64_t b_desc = umma_desc_mn_sw128(v_addr + k * 2048);
              const uint32_t acc = (i | k) != 0 ? 1u : 0u;
              if constexpr (P_IN_TMEM) {
                umma_bf16_ts(t_o, t_p + k * 8, b_desc, idesc_o, acc);
              } else {
                const uint32_t p_addr = smem_u32(sP) + (k >> 2) * kFaTileBytes + (k & 3) * 32;
                umma_bf16_ss(t_o, umma_desc_k_sw128(p_addr), b_desc, idesc_o, acc);
              }
            }
            umma_commit(&v_empty[st]);
            umma_commit(pv_done);
            OPV_FA_STAMP(i, 7);
          }
          __syncwarp();
          ++pc;
        }
        ++tiles_done;
        tracing = false;
      }
    }
  } else {
    // ------------------------------ softmax warps (one thread per query row) ----
    setmaxnreg_inc<216>();
    const int r_tile = warp * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t t_p = t_s + 128, t_o = t_s + 192;
    const float scale_log2 = 0.125f * 1.44269504088896340736f;  // head_dim^-0.5 * log2(e)
    uint32_t tiles_done = 0, bc = 0;  // running count of key blocks -> barrier parity
    Tile tl;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    if (!decode(t, tl)) continue;
    const int begin = tl.begin, n = tl.n, q0 = tl.q0, head = tl.head, key_base = tl.key_base, nb = tl.nb;
    const int row = q0 + r_tile;          // row inside the sequence
    float m_run = -CUDART_INF_F, l_run = 0.f;
    // keys this row may attend to: [k_lo, k_lo + k_span]; for the warp's 32 rows: keys every row sees
    // [all_lo, all_hi] and keys some row sees [any_lo, any_hi]
    const int row_first = q0 + warp * 32, row_last = row_first + 31;
    const int k_lo = global ? 0 : max(row - half_window, 0);
    const int k_hi = global ? n - 1 : min(row + half_window, n - 1);
    const uint32_t k_span = static_cast<uint32_t>(k_hi - k_lo);
    const int all_lo = global ? 0 : max(row_last - half_window, 0);
    const int all_hi = global ? n - 1 : min(row_first + half_window, n - 1);
    const int any_lo = global ? 0 : max(row_first - half_window, 0);
    const int any_hi = global ? n - 1 : min(row_last + half_window, n - 1);

    for (int i = 0; i < nb; ++i, ++bc) {
      mbar_wait_spin(s_full, bc & 1);
      tc_fence_after();
      if (warp == 0) OPV_FA_STAMP(i, 0);
      uint32_t sr[128];
      tmem_ld_32x32b_2x64(t_s, sr);
      if (warp == 0) OPV_FA_STAMP(i, 1);
      if (warp < 3) OPV_FA_STAMP(i, 8 + 2 * warp);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);  // the MMA warp may overwrite S with S(i+1)

      // Columns are handled in four 32-key chunks; each chunk is classified for the WHOLE warp (rows
      // row_first..row_first+31): skipped when no row may see it, unmasked when every row sees all of it.
      const int key0 = key_base + i * kFaBlockN;
      int kind[4];  // 0 = skip, 1 = full, 2 = mixed
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const int c_lo = key0 + 32 * ch, c_hi = c_lo + 31;
        kind[ch] = (c_hi < any_lo || c_lo > any_hi) ? 0 : ((c_lo >= all_lo && c_hi <= all_hi) ? 1 : 2);
      }
      // Every block of a global layer except a sequence's last one is fully visible to every row.  That case
      // runs as ONE straight-line basic block (no per-chunk branches) so that ptxas can keep MUFU.EX2 results in
      // flight across the whole row: the softmax loop alone needs ~1200 cycles per block for a single warp
      // (tools/micro/softmax_loop.cu), the branchy general path below measured ~2500 inside this kernel.
      const bool all_full = (kind[0] & kind[1] & kind[2] & kind[3]) == 1 && (kind[0] | kind[1] | kind[2] | kind[3]) == 1;
      uint32_t pr[64];
      float corr;
      bool upd;
      if (all_full) {
        float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F, mx2 = -CUDART_INF_F, mx3 = -CUDART_INF_F;
#pragma unroll
        for (int c = 0; c < 128; c += 8) {
          mx0 = fmax3(mx0, __uint_as_float(sr[c + 0]), __uint_as_float(sr[c + 1]));
          mx1 = fmax3(mx1, __uint_as_float(sr[c + 2]), __uint_as_float(sr[c + 3]));
          mx2 = fmax3(mx2, __uint_as_float(sr[c + 4]), __uint_as_float(sr[c + 5]));
          mx3 = fmax3(mx3, __uint_as_float(sr[c + 6]), __uint_as_float(sr[c + 7]));
        }
        const float m_cand = fmaxf(m_run, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2);  // finite
        upd = (m_cand - m_run) > kFaRescaleThreshold;
        const float m_new = upd ? m_cand : m_run;
        corr = upd ? ex2_approx(m_run - m_new) : 1.0f;
        m_run = m_new;
        // packed fp32 pairs (FFMA2 / FADD2, sm_100): half an issue slot per score for the scale-and-shift and for the
        // row sum; the loop is co-limited by MUFU.EX2 and by issue slots (profiles/r2_attention_notes.md)
        float2 acc01 = make_float2(0.f, 0.f), acc23 = make_float2(0.f, 0.f);
        const float2 sc2 = make_float2(scale_log2, scale_log2), nm2 = make_float2(-m_new, -m_new);
#pragma unroll
        for (int c = 0; c < 64; c += 2) {
          const float2 x01 = __ffma2_rn(make_float2(__uint_as_float(sr[2 * c]), __uint_as_float(sr[2 * c + 1])), sc2, nm2);
          const float2 x23 = __ffma2_rn(make_float2(__uint_as_float(sr[2 * c + 2]), __uint_as_float(sr[2 * c + 3])), sc2, nm2);
          const float2 p01 = make_float2(ex2_approx(x01.x), ex2_approx(x01.y));
          const float2 p23 = make_float2(ex2_approx(x23.x), ex2_approx(x23.y));
          acc01 = __fadd2_rn(acc01, p01);
          acc23 = __fadd2_rn(acc23, p23);
          pr[c] = pack_bf16x2(p01.x, p01.y);
          pr[c + 1] = pack_bf16x2(p23.x, p23.y);
        }
        const float2 acc = __fadd2_rn(acc01, acc23);
        l_run = l_run * corr + (acc.x + acc.y);
      } else {
        float mx = -CUDART_INF_F;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          if (kind[ch] == 0) continue;
          if (kind[ch] == 2) {
            const int c_lo = key0 + 32 * ch;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const bool ok = static_cast<uint32_t>(c_lo + c - k_lo) <= k_span;
              if (!ok) sr[32 * ch + c] = 0xff800000u;  // -inf
            }
          }
          float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F;
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            mx0 = fmax3(mx0, __uint_as_float(sr[32 * ch + c + 0]), __uint_as_float(sr[32 * ch + c + 1]));
            mx1 = fmax3(mx1, __uint_as_float(sr[32 * ch + c + 2]), __uint_as_float(sr[32 * ch + c + 3]));
          }
          mx = fmax3(mx, mx0, mx1);
        }
        const float m_cand = fmaxf(m_run, mx * scale_log2);
        upd = (m_cand - m_run) > kFaRescaleThreshold;  // false when both are -inf (NaN)
        const float m_new = upd ? m_cand : m_run;
        corr = upd ? ex2_approx(m_run - m_new) : 1.0f;
        const float m_use = (m_new == -CUDART_INF_F) ? 0.f : m_new;
        m_run = m_new;
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          if (kind[ch] == 0) {
#pragma unroll
            for (int c = 0; c < 16; ++c) pr[16 * ch + c] = 0u;
          } else {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const float a = ex2_approx(fmaf(__uint_as_float(sr[32 * ch + 2 * c]), scale_log2, -m_use));
              const float b = ex2_approx(fmaf(__uint_as_float(sr[32 * ch + 2 * c + 1]), scale_log2, -m_use));
              sum0 += a, sum1 += b;
              pr[16 * ch + c] = pack_bf16x2(a, b);
            }
          }
        }
        l_run = l_run * corr + (sum0 + sum1);
      }

      if (i > 0) {
        mbar_wait_spin(pv_done, (bc - 1) & 1);  // O holds blocks < i and the P buffer is free again
        tc_fence_after();
        if (warp == 0) OPV_FA_STAMP(i, 4);
        if (__any_sync(0xffffffffu, upd)) {
          uint32_t orr[64];
          tmem_ld_32x32b_x64(t_o, orr);
#pragma unroll
          for (int c = 0; c < 64; ++c) orr[c] = __float_as_uint(__uint_as_float(orr[c]) * corr);
          tmem_st_32x32b_x64(t_o, orr);
        }
      }
      if constexpr (P_IN_TMEM) {
        tmem_st_32x32b_x64(t_p, pr);
        tc_fence_before();
      } else {
        // K-major, 128 B swizzle: 16 B chunk g of row r lives at chunk (g ^ (r & 7)) of its 128 B row
#pragma unroll
        for (int g = 0; g < 16; ++g) {
          uint8_t* dst = sP + (g >> 3) * kFaTileBytes + r_tile * 128 + (((g & 7) ^ (r_tile & 7)) << 4);
          *reinterpret_cast<uint4*>(dst) = make_uint4(pr[4 * g], pr[4 * g + 1], pr[4 * g + 2], pr[4 * g + 3]);
        }
        fence_proxy_async_smem();
        tc_fence_before();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (warp == 0) OPV_FA_STAMP(i, 5);
      if (warp < 3) OPV_FA_STAMP(i, 9 + 2 * warp);
    }

    // epilogue: O / l -> bf16 -> out[begin + row, head*64 : head*64+64]
    mbar_wait_spin(pv_done, (bc - 1) & 1);
    tc_fence_after();
    uint32_t orr[64];
    tmem_ld_32x32b_x64(t_o, orr);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(o_empty);  // the next tile's first P.V may overwrite O
    if (row < n) {
      const float inv = 1.0f / l_run;
      uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<int64_t>(begin) + row) * H + head * 64);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(orr[8 * g + 0]) * inv, __uint_as_float(orr[8 * g + 1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(orr[8 * g + 2]) * inv, __uint_as_float(orr[8 * g + 3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(orr[8 * g + 4]) * inv, __uint_as_float(orr[8 * g + 5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(orr[8 * g + 6]) * inv, __uint_as_float(orr[8 * g + 7]) * inv);
        dst[g] = u;
      }
    }
    ++tiles_done;
    tracing = false;
    }  // tile loop
  }

  tc_fence_before();
  __syncthreads();
  tracing = trace_cta;
  if (warp == 0) OPV_FA_STAMP(0, 15);  // all roles done (before the TMEM release)
#undef OPV_FA_STAMP
  if (warp == 6) tmem_dealloc(tmem_base, kFaTmemCols);
}

}  // namespace opv
