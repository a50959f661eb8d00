// Varlen self-attention on tcgen05, third organisation ("pp"): ONE CTA PER SM, TWO 128-row QUERY TILES IN FLIGHT.
//
// Why (profiles/r1v_ncu_full_gemm_attention.md): the 2-CTAs-per-SM kernels reach 0.36 of the tensor peak on
// the global layers -- MUFU.EX2 (16 / clk / SM) is the bounding unit, but with two in-order softmax warps per
// scheduler it is only ~55-70 % busy, and every CTA streams its own copy of K and V from L2 (32 KB per key block per
// 128 query rows).  Here one CTA owns the whole SM (all 512 TMEM columns, 640 threads):
//   * work unit = (sequence, head, 256 query rows) = Q tiles A and B; each K / V block is loaded ONCE for both
//     (half the L2 -> smem traffic per query row) through 3-stage rings;
//   * 16 softmax warps: tile j = warp / 8, column half = (warp / 4) % 2, TMEM lane quarter = warp % 4.  Thread
//     (j, half, quarter, lane) owns keys [64 * half, +64) of row 32 * quarter + lane of tile j: FOUR softmax warps per
//     scheduler with 64-long dependency chains, the scores read from TMEM once (64 registers);
//   * TMEM: S_A [0,128) | S_B [128,256) | P_A [256,320) | P_B [320,384) | O_A [384,448) | O_B [448,512).  P has its own
//     columns, so S_j(i+1) is issued as soon as the softmax warps have READ S_j(i) and runs under softmax_j(i);
//     P_j(i).V(i) follows p_full_j(i).  The single MMA warp interleaves the two tiles:
//     S_A(i+1), S_B(i+1), PV_A(i), PV_B(i), ... -- the order in which their operands become ready when tile A runs
//     half a block ahead of tile B;
//   * the two halves of a row agree on the running max through shared memory + a 64-thread named barrier (lazy
//     rescale, threshold 2^8, as in the other kernels), keep separate row sums (added in the epilogue) and each
//     rescales / normalises / stores 32 of the 64 output columns.
// Key-block ranges are per tile, so sliding-window layers work too (tile A walks blocks [0,2), tile B [1,3) of the
// 3-block band of a 256-row super tile); masking is the same warp-uniform chunk classification as attention_tcgen05.
#pragma once

#include <math_constants.h>

#include "attention_tcgen05.cuh"

namespace opv {

constexpr int kPpThreads = 640;   // warps 0..15 softmax, 16 TMA producer, 17 MMA issuer, 18 TMEM allocation, 19 idle
constexpr int kPpKvStages = 3;
constexpr int kPpTmemCols = 512;
constexpr int kPpSuperM = 2 * kFaBlockM;  // query rows per work unit

struct PpSmemLayout {
  static constexpr int kQ = 0;                                        // 2 tiles x [128][64] bf16
  static constexpr int kK = kQ + 2 * kFaTileBytes;
  static constexpr int kV = kK + kPpKvStages * kFaTileBytes;
  static constexpr int kExchange = kV + kPpKvStages * kFaTileBytes;   // float [2 tiles][2 parities][2 halves][128 rows]
  static constexpr int kBars = kExchange + 2 * 2 * 2 * 128 * 4;
  static constexpr int kTotal = kBars + 512 + 1024;                   // + barriers + slack for the 1024 B alignment
};

// DBG (tools/attn_check.py only; 0 in the product): 1 = FMUL instead of MUFU.EX2, 2 = no tcgen05.ld of the scores,
// 3 = the MMA warp issues no MMA (commits only), 4 = no row-max exchange between the halves.  Results are wrong by
// construction; the variants exist to time what is left when one resource is taken out.
template <int DBG>
__global__ void __launch_bounds__(kPpThreads, 1)
attention_tcgen05_pp_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out,
                            const int32_t* __restrict__ cu_seqlens, const int H, const int half_window,
                            const int n_seqs, const int supers_per_seq, const int pdl_late,
                            long long* __restrict__ trace) {
  // trace (tools/attn_check.py only; nullptr in the product): clock64() stamps of CTA 0's first super tile, 32 slots per
  // key block: softmax thread (tile j, half 0, quarter 0, lane 0) at 8 j + [0 s_full seen, 1 scores loaded, 2 row max
  // agreed, 3 exponentials done, 4 pv_done seen, 5 P published]; MMA warp 16 + [0 S_A issued, 1 S_B issued, 2 PV_A issued,
  // 3 PV_B issued]; TMA warp 24 + [0 K(i) requested, 1 V(i) requested]
  using L = PpSmemLayout;
  bool tracing = trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0;
#define OPV_PP_STAMP(blk, slot) do { if (tracing) trace[(blk) * 32 + (slot)] = clock64(); } while (0)
#define OPV_PP_STAMP_F(blk, slot, val) do { if (tracing) { long long t__; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t__), "+f"(val)); trace[(blk) * 32 + (slot)] = t__; } } while (0)
  constexpr int S = kPpKvStages;
  const int heads = H / 64;
  const int total_tiles = n_seqs * heads * supers_per_seq;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + L::kQ;
  uint8_t* sK = smem + L::kK;
  uint8_t* sV = smem + L::kV;
  float* xchg = reinterpret_cast<float*>(smem + L::kExchange);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
  uint64_t* q_full = bars;             // [2]  Q_j landed                                    (TMA)
  uint64_t* q_empty = q_full + 2;      // [2]  last S_j of a super tile complete             (tcgen05.commit)
  uint64_t* k_full = q_empty + 2;      // [S]
  uint64_t* k_empty = k_full + S;      // [S]  every S MMA reading the stage complete        (tcgen05.commit)
  uint64_t* v_full = k_empty + S;      // [S]
  uint64_t* v_empty = v_full + S;      // [S]
  uint64_t* s_full = v_empty + S;      // [2]  S_j(i) complete in TMEM                       (tcgen05.commit)
  uint64_t* s_empty = s_full + 2;      // [2]  S_j(i) read into registers                    (8 warp arrivals)
  uint64_t* p_full = s_empty + 2;      // [2]  P_j(i) written (+ O_j rescaled)               (8 warp arrivals)
  uint64_t* pv_done = p_full + 2;      // [2]  O_j += P_j(i).V(i) complete                   (tcgen05.commit)
  uint64_t* o_empty = pv_done + 2;     // [2]  O_j of a super tile read by the epilogue      (8 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const bool global = half_window < 0;
  auto ex2 = [](const float x) { return DBG == 1 ? x * 0.5f : ex2_approx(x); };

  // Super tile t -> (sequence, head, 256-row query range); consecutive t are consecutive query ranges of one
  // (sequence, head), so the CTAs running at the same time share K / V in L2.  Key blocks of 128 start at key_base
  // (0 for global layers; q0 - half_window, UNALIGNED, for sliding-window layers: TMA zero-fills rows before the
  // tensor, rows of the previous sequence are masked).  Tile j walks key blocks [lo[j], hi[j]) of the nb loaded ones.
  struct Super {  // scalars only (indexable arrays would live in local memory)
    int begin, n, q0, head, key_base, nb, lo0, hi0, lo1, hi1;
    __device__ __forceinline__ int lo(int j) const { return j ? lo1 : lo0; }
    __device__ __forceinline__ int hi(int j) const { return j ? hi1 : hi0; }
  };
  auto decode = [&](const int t, Super& st) -> bool {
    const int qt = t % supers_per_seq;
    const int sh = t / supers_per_seq;
    const int seq = sh / heads;
    st.head = sh - seq * heads;
    st.begin = cu_seqlens[seq];
    st.n = cu_seqlens[seq + 1] - st.begin;
    st.q0 = qt * kPpSuperM;
    if (st.q0 >= st.n) return false;
    const bool two = st.n - st.q0 > kFaBlockM;
    if (global) {
      st.key_base = 0;
      st.nb = (st.n + kFaBlockN - 1) / kFaBlockN;
      st.lo0 = 0, st.hi0 = st.nb;
      st.lo1 = 0, st.hi1 = two ? st.nb : 0;
    } else {
      st.key_base = st.q0 - half_window;
      const int key_end = min(st.n, st.q0 + (two ? 2 : 1) * kFaBlockM + half_window);
      st.nb = (key_end - st.key_base + kFaBlockN - 1) / kFaBlockN;
      // tile j sees keys [q0 + 128 j - w, q0 + 128 j + 127 + w] = relative [128 j, 128 j + 127 + 2 w]
      st.lo0 = 0, st.hi0 = min(st.nb, (kFaBlockM - 1 + 2 * half_window) / kFaBlockN + 1);
      st.lo1 = two ? 1 : 0, st.hi1 = two ? min(st.nb, (2 * kFaBlockM - 1 + 2 * half_window) / kFaBlockN + 1) : 0;
    }
    return true;
  };

  if (warp == 16 && lane == 0) tma_prefetch_desc(&tm_qkv);
  if (warp == 17 && lane == 0) {
    for (int j = 0; j < 2; ++j) {
      mbar_init(&q_full[j], 1);
      mbar_init(&q_empty[j], 1);
      mbar_init(&s_full[j], 1);
      mbar_init(&s_empty[j], 8);
      mbar_init(&p_full[j], 8);
      mbar_init(&pv_done[j], 1);
      mbar_init(&o_empty[j], 8);
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 18) {
    tmem_alloc(tmem_slot, kPpTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (!pdl_late) pdl_launch_dependents();
  pdl_wait();  // the prologue above overlapped the previous kernel; qkv is visible from here on
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform (single UTCHMMA per MMA)

  // Register pool of the CTA = 96 (ptxas cap for 640 threads) x 640 = 61440: 16 softmax warps x 32 x 112 + 4 x 32 x 32.
  // setmaxnreg.inc only draws from what the CTA's own warps released, never from the SM's unallocated registers.
  if (warp >= 16) {
    setmaxnreg_dec<32>();
    if (warp == 16) {
      // ------------------------------ TMA producer ------------------------------
      if (lane == 0) {
        uint32_t td0 = 0, td1 = 0, kc = 0, vc = 0;  // running counts -> ring stage and barrier parity
        Super st;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
          if (!decode(t, st)) continue;
          auto load_q = [&](const int j, uint32_t& td) {
            if (st.hi(j) <= st.lo(j)) return;
            if (td > 0) mbar_wait(&q_empty[j], (td - 1) & 1);  // the previous super tile's S_j MMAs have read Q_j
            mbar_expect_tx(&q_full[j], kFaTileBytes);
            tma_load_2d(sQ + j * kFaTileBytes, &tm_qkv, &q_full[j], st.head * 64, st.begin + st.q0 + j * kFaBlockM);
            ++td;
          };
          load_q(0, td0);
          load_q(1, td1);
          // consumption order of the MMA warp: K0, K1, V0, K2, V1, ...
          for (int i = 0; i <= st.nb; ++i) {
            if (i < st.nb) {
              const uint32_t sg = kc % S;
              mbar_wait(&k_empty[sg], ((kc / S) & 1) ^ 1);
              mbar_expect_tx(&k_full[sg], kFaTileBytes);
              tma_load_2d(sK + sg * kFaTileBytes, &tm_qkv, &k_full[sg], H + st.head * 64,
                          st.begin + st.key_base + i * kFaBlockN);
              OPV_PP_STAMP(i, 24);
              ++kc;
            }
            if (i >= 1) {
              const uint32_t sg = vc % S;
              mbar_wait(&v_empty[sg], ((vc / S) & 1) ^ 1);
              mbar_expect_tx(&v_full[sg], kFaTileBytes);
              tma_load_2d(sV + sg * kFaTileBytes, &tm_qkv, &v_full[sg], 2 * H + st.head * 64,
                          st.begin + st.key_base + (i - 1) * kFaBlockN);
              OPV_PP_STAMP(i - 1, 25);
              ++vc;
            }
          }
          tracing = false;
        }
      }
    } else if (warp == 17) {
      // ------------------------------ MMA issuer --------------------------------
      constexpr uint32_t idesc_s = umma_idesc_bf16_f32(kFaBlockM, kFaBlockN);  // Q.K^T: both K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16_f32_bmn(kFaBlockM, 64);     // P.V: V is MN-major
      const uint32_t q_addr = smem_u32(sQ);
      uint32_t td0 = 0, td1 = 0, sc0 = 0, sc1 = 0, pc0 = 0, pc1 = 0, kc = 0, vc = 0;
      bool tracing_mma = true;
      Super st;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        if (!decode(t, st)) continue;
        for (int i = 0; i <= st.nb; ++i) {
          if (i < st.nb) {  // S_j(i) = Q_j . K(i)^T for the tiles whose range holds block i
            const uint32_t sg = kc % S;
            mbar_wait(&k_full[sg], (kc / S) & 1);
            const uint32_t k_addr = smem_u32(sK + sg * kFaTileBytes);
            auto issue_s = [&](const int j, const uint32_t td, uint32_t& sc) {
              if (i < st.lo(j) || i >= st.hi(j)) return;
              if (i == st.lo(j)) mbar_wait(&q_full[j], td & 1);
              if (sc > 0) mbar_wait(&s_empty[j], (sc - 1) & 1);  // the softmax warps have read the previous S_j
              tc_fence_after();
              if (elect_one()) {
                const uint32_t t_s = tmem_base + j * 128;
#pragma unroll
                for (int k = 0; k < (DBG == 3 ? 0 : 4); ++k)
                  umma_bf16_ss(t_s, umma_desc_k_sw128(q_addr + j * kFaTileBytes + k * 32),
                               umma_desc_k_sw128(k_addr + k * 32), idesc_s, k != 0 ? 1u : 0u);
                umma_commit(&s_full[j]);
                if (i == st.hi(j) - 1) umma_commit(&q_empty[j]);
                if (trace != nullptr && blockIdx.x == 0 && tracing_mma) trace[i * 32 + 16 + j] = clock64();
              }
              __syncwarp();
              ++sc;
            };
            issue_s(0, td0, sc0);
            issue_s(1, td1, sc1);
            if (elect_one()) umma_commit(&k_empty[sg]);
            __syncwarp();
            ++kc;
          }
          if (i >= 1) {  // O_j += P_j(b) . V(b)
            const int b = i - 1;
            const uint32_t sg = vc % S;
            mbar_wait(&v_full[sg], (vc / S) & 1);
            const uint32_t v_addr = smem_u32(sV + sg * kFaTileBytes);
            auto issue_pv = [&](const int j, const uint32_t td, uint32_t& pc) {
              if (b < st.lo(j) || b >= st.hi(j)) return;
              mbar_wait(&p_full[j], pc & 1);
              if (b == st.lo(j) && td > 0) mbar_wait(&o_empty[j], (td - 1) & 1);  // epilogue has read the previous O_j
              tc_fence_after();
              if (elect_one()) {
                const uint32_t t_p = tmem_base + 256 + j * 64, t_o = tmem_base + 384 + j * 64;
                const uint32_t first = b == st.lo(j) ? 0u : 1u;
#pragma unroll
                for (int k = 0; k < (DBG == 3 ? 0 : 8); ++k)  // 16 keys per MMA: two 8-key groups of 1024 B
                  umma_bf16_ts(t_o, t_p + k * 8, umma_desc_mn_sw128(v_addr + k * 2048), idesc_o, (k != 0) ? 1u : first);
                umma_commit(&pv_done[j]);
                if (trace != nullptr && blockIdx.x == 0 && tracing_mma) trace[b * 32 + 18 + j] = clock64();
              }
              __syncwarp();
              ++pc;
            };
            issue_pv(0, td0, pc0);
            issue_pv(1, td1, pc1);
            if (elect_one()) umma_commit(&v_empty[sg]);
            __syncwarp();
            ++vc;
          }
        }
        if (st.hi0 > st.lo0) ++td0;
        if (st.hi1 > st.lo1) ++td1;
        tracing_mma = false;
      }
    }
  } else {
    // ------------------------------ softmax warps: two threads per query row, two tiles ----
    setmaxnreg_inc<112>();
    const int j = warp >> 3, half = (warp >> 2) & 1, quarter = warp & 3;
    const int r_tile = quarter * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_s = lane_base + j * 128 + 64 * half;        // this thread's 64 score columns
    const uint32_t t_p = lane_base + 256 + j * 64 + 32 * half;   // its 32 packed-probability columns
    const uint32_t t_o = lane_base + 384 + j * 64 + 32 * half;   // the 32 output columns it rescales / stores
    const float scale_log2 = 0.125f * 1.44269504088896340736f;   // head_dim^-0.5 * log2(e)
    const int pair_bar = 1 + j * 4 + quarter;                    // named barrier of the two warps sharing these rows
    // row max / row sum exchange between the two halves of a row; slots alternate with the block parity so that a
    // thread that runs ahead cannot overwrite a value its partner has not read yet
    float* const my_slots = xchg + j * 512 + half * 128 + r_tile;
    const float* const other_slots = xchg + j * 512 + (half ^ 1) * 128 + r_tile;
    uint64_t* const my_s_full = &s_full[j];
    uint64_t* const my_s_empty = &s_empty[j];
    uint64_t* const my_p_full = &p_full[j];
    uint64_t* const my_pv_done = &pv_done[j];
    uint64_t* const my_o_empty = &o_empty[j];
    uint32_t bc = 0;  // running count of this tile's key blocks -> barrier parity
    tracing = tracing && half == 0 && quarter == 0;
    const int tslot = 8 * j;
    Super st;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      if (!decode(t, st)) continue;
      const int lo = st.lo(j), hi = st.hi(j);
      if (hi <= lo) continue;  // second tile of a super tile that ends inside the first one
      const int n = st.n, q0 = st.q0 + j * kFaBlockM;
      const int row = q0 + r_tile;  // row inside the sequence
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

      for (int i = lo; i < hi; ++i, ++bc) {
        const int key0 = st.key_base + i * kFaBlockN + 64 * half;  // first key of this thread's 64
        int kind[2];  // per 32-key chunk, warp-uniform: 0 = no row sees it, 1 = every row sees all of it, 2 = mixed
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c_lo = key0 + 32 * q, c_hi = c_lo + 31;
          kind[q] = (c_hi < any_lo || c_lo > any_hi) ? 0 : ((c_lo >= all_lo && c_hi <= all_hi) ? 1 : 2);
        }
        mbar_wait(my_s_full, bc & 1);
        tc_fence_after();
        OPV_PP_STAMP(i, tslot + 0);
        uint32_t sr[64];
        if constexpr (DBG == 2) {
#pragma unroll
          for (int c = 0; c < 64; ++c) sr[c] = __float_as_uint(static_cast<float>((lane * 7 + c * 3 + i) & 31) * 0.01f);
        } else {
          tmem_ld_32x32b_x64(t_s, sr);
        }
        OPV_PP_STAMP(i, tslot + 1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(my_s_empty);  // the MMA warp may overwrite S_j with S_j(i+1)

        uint32_t pr[32];
        float corr, sum;
        bool upd;
        if (kind[0] == 1 && kind[1] == 1) {
          // Every block of a global layer except a sequence's last one: ONE straight-line basic block.
          float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F, mx2 = -CUDART_INF_F, mx3 = -CUDART_INF_F;
#pragma unroll
          for (int c = 0; c < 64; c += 8) {
            mx0 = fmax3(mx0, __uint_as_float(sr[c + 0]), __uint_as_float(sr[c + 1]));
            mx1 = fmax3(mx1, __uint_as_float(sr[c + 2]), __uint_as_float(sr[c + 3]));
            mx2 = fmax3(mx2, __uint_as_float(sr[c + 4]), __uint_as_float(sr[c + 5]));
            mx3 = fmax3(mx3, __uint_as_float(sr[c + 6]), __uint_as_float(sr[c + 7]));
          }
          const float mine = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
          float mx = mine;
          if constexpr (DBG != 4) {
            my_slots[(bc & 1) * 256] = mine;
            named_bar_sync(pair_bar, 64);
            mx = fmaxf(mine, other_slots[(bc & 1) * 256]);
          }
          OPV_PP_STAMP_F(i, tslot + 2, mx);
          const float m_cand = fmaxf(m_run, mx * scale_log2);  // finite
          upd = (m_cand - m_run) > kFaRescaleThreshold;
          const float m_new = upd ? m_cand : m_run;
          corr = upd ? ex2_approx(m_run - m_new) : 1.0f;
          m_run = m_new;
          float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            const float a = ex2(fmaf(__uint_as_float(sr[2 * c]), scale_log2, -m_new));
            const float b = ex2(fmaf(__uint_as_float(sr[2 * c + 1]), scale_log2, -m_new));
            const float e = ex2(fmaf(__uint_as_float(sr[2 * c + 2]), scale_log2, -m_new));
            const float f = ex2(fmaf(__uint_as_float(sr[2 * c + 3]), scale_log2, -m_new));
            sum0 += a, sum1 += b, sum2 += e, sum3 += f;
            pr[c] = pack_bf16x2(a, b);
            pr[c + 1] = pack_bf16x2(e, f);
          }
          sum = (sum0 + sum1) + (sum2 + sum3);
        } else {
          float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (kind[q] == 0) continue;
            if (kind[q] == 2) {
              const int d = key0 + 32 * q - k_lo;
#pragma unroll
              for (int c = 0; c < 32; ++c)
                if (static_cast<uint32_t>(d + c) > k_span) sr[32 * q + c] = 0xff800000u;  // -inf
            }
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              mx0 = fmax3(mx0, __uint_as_float(sr[32 * q + c + 0]), __uint_as_float(sr[32 * q + c + 1]));
              mx1 = fmax3(mx1, __uint_as_float(sr[32 * q + c + 2]), __uint_as_float(sr[32 * q + c + 3]));
            }
          }
          const float mine = fmaxf(mx0, mx1);
          my_slots[(bc & 1) * 256] = mine;
          named_bar_sync(pair_bar, 64);
          const float mx = fmaxf(mine, other_slots[(bc & 1) * 256]);
          const float m_cand = fmaxf(m_run, mx * scale_log2);
          upd = (m_cand - m_run) > kFaRescaleThreshold;  // false when both are -inf (NaN); same in both halves
          const float m_new = upd ? m_cand : m_run;
          corr = upd ? ex2_approx(m_run - m_new) : 1.0f;
          const float m_use = (m_new == -CUDART_INF_F) ? 0.f : m_new;
          m_run = m_new;
          float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (kind[q] == 0) {
#pragma unroll
              for (int c = 0; c < 16; ++c) pr[16 * q + c] = 0u;
            } else {
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const float a = ex2(fmaf(__uint_as_float(sr[32 * q + 2 * c]), scale_log2, -m_use));
                const float b = ex2(fmaf(__uint_as_float(sr[32 * q + 2 * c + 1]), scale_log2, -m_use));
                sum0 += a, sum1 += b;
                pr[16 * q + c] = pack_bf16x2(a, b);
              }
            }
          }
          sum = sum0 + sum1;
        }
        OPV_PP_STAMP_F(i, tslot + 3, sum);
        l_run = l_run * corr + sum;

        if (i > lo) {
          mbar_wait(my_pv_done, (bc - 1) & 1);  // O_j holds blocks < i and the P_j buffer is free again
          tc_fence_after();
          OPV_PP_STAMP(i, tslot + 4);
          if (__any_sync(0xffffffffu, upd)) {
            uint32_t orr[32];
            tmem_ld_32x32_raw(t_o, orr);
#pragma unroll
            for (int c = 0; c < 32; ++c) orr[c] = __float_as_uint(__uint_as_float(orr[c]) * corr);
            tmem_st_32x32b_x32(t_o, orr);
          }
        }
        tmem_st_32x32b_x32(t_p, pr);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(my_p_full);
        OPV_PP_STAMP(i, tslot + 5);
      }
      tracing = false;

      // epilogue: O / l -> bf16 -> out[begin + row, head*64 + 32*half : +32]
      my_slots[(bc & 1) * 256] = l_run;  // parity of the NEXT block: last used two blocks ago
      mbar_wait(my_pv_done, (bc - 1) & 1);
      tc_fence_after();
      uint32_t orr[32];
      tmem_ld_32x32_raw(t_o, orr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(my_o_empty);  // the next super tile's first P_j.V may overwrite O_j
      named_bar_sync(pair_bar, 64);
      const float l_total = l_run + other_slots[(bc & 1) * 256];
      named_bar_sync(pair_bar, 64);  // both halves have read the sums before the next tile's max exchange
      if (row < n) {
        const float inv = 1.0f / l_total;
        uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<int64_t>(st.begin) + row) * H + st.head * 64 + 32 * half);
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
  if (warp == 18) tmem_dealloc(tmem_base, kPpTmemCols);
#undef OPV_PP_STAMP
#undef OPV_PP_STAMP_F
}

}  // namespace opv
