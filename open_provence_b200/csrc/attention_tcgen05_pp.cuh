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

constexpr int kPpThreads = 640;   // warps 0..15 softmax, 16 TMA producer, 17 / 18 MMA issuers of tile A / B, 19 TMEM allocation
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

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024 B alignment (128 B swizzle) is established in the SHARED address space: the shared-window address of a
  // __shared__ symbol is a compile-time constant, so everything derived from it is free to rematerialise.  Rounding the
  // GENERIC pointer instead made ptxas rebuild the window base (S2UR SR_SWINHI / SR_CgaCtaId + uniform arithmetic) in
  // front of every barrier operation of the softmax loop.
  const uint32_t raw0 = smem_u32(smem_raw);
  const uint32_t sm0 = (raw0 + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (sm0 - raw0);
  // Every shared-memory address below is the 32-bit shared-window address of the aligned base + a constant, computed
  // once: the barrier / TMA / descriptor helpers take such addresses (see common.cuh, *_a).
  const uint32_t a_q = sm0 + L::kQ, a_k = sm0 + L::kK, a_v = sm0 + L::kV, a_xchg = sm0 + L::kExchange;
  const uint32_t q_full = sm0 + L::kBars;     // [2]  Q_j landed                                    (TMA)
  const uint32_t q_empty = q_full + 8 * 2;    // [2]  last S_j of a super tile complete             (tcgen05.commit)
  const uint32_t k_full = q_empty + 8 * 2;    // [S]
  const uint32_t k_empty = k_full + 8 * S;    // [S]  every S MMA reading the stage complete        (2 x tcgen05.commit)
  const uint32_t v_full = k_empty + 8 * S;    // [S]
  const uint32_t v_empty = v_full + 8 * S;    // [S]
  const uint32_t s_full = v_empty + 8 * S;    // [2]  S_j(i) complete in TMEM                       (tcgen05.commit)
  const uint32_t s_empty = s_full + 8 * 2;    // [2]  S_j(i) read into registers                    (8 warp arrivals)
  const uint32_t p_full = s_empty + 8 * 2;    // [2]  P_j(i) written (+ O_j rescaled)               (8 warp arrivals)
  const uint32_t pv_done = p_full + 8 * 2;    // [2]  O_j += P_j(i).V(i) complete                   (tcgen05.commit)
  const uint32_t o_empty = pv_done + 8 * 2;   // [2]  O_j of a super tile read by the epilogue      (8 warp arrivals)
  const uint32_t x_a = o_empty + 8 * 2;       // [4]  per lane quarter: tile B's exponentials issued -> tile A may start
  const uint32_t x_b = x_a + 8 * 4;           // [4]  per lane quarter: tile A's exponentials issued -> tile B may start
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kBars + 8 * (2 * 7 + 4 * S + 8));

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
      mbar_init_a(q_full + 8 * j, 1);
      mbar_init_a(q_empty + 8 * j, 1);
      mbar_init_a(s_full + 8 * j, 1);
      mbar_init_a(s_empty + 8 * j, 8);
      mbar_init_a(p_full + 8 * j, 8);
      mbar_init_a(pv_done + 8 * j, 1);
      mbar_init_a(o_empty + 8 * j, 8);
    }
    for (int q = 0; q < 4; ++q) {
      mbar_init_a(x_a + 8 * q, 2);
      mbar_init_a(x_b + 8 * q, 2);
    }
    for (int s = 0; s < S; ++s) {
      mbar_init_a(k_full + 8 * s, 1);
      mbar_init_a(k_empty + 8 * s, 2);  // one tcgen05.commit per issuer
      mbar_init_a(v_full + 8 * s, 1);
      mbar_init_a(v_empty + 8 * s, 2);
    }
    fence_mbar_init();
  }
  if (warp == 19) {
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
            if (td > 0) mbar_wait_a(q_empty + 8 * j, (td - 1) & 1);  // the previous super tile's S_j MMAs have read Q_j
            mbar_expect_tx_a(q_full + 8 * j, kFaTileBytes);
            tma_load_2d_a(a_q + j * kFaTileBytes, &tm_qkv, q_full + 8 * j, st.head * 64, st.begin + st.q0 + j * kFaBlockM);
            ++td;
          };
          load_q(0, td0);
          load_q(1, td1);
          // consumption order of the MMA warp: K0, K1, V0, K2, V1, ...
          for (int i = 0; i <= st.nb; ++i) {
            if (i < st.nb) {
              const uint32_t sg = kc % S;
              mbar_wait_a(k_empty + 8 * sg, ((kc / S) & 1) ^ 1);
              mbar_expect_tx_a(k_full + 8 * sg, kFaTileBytes);
              tma_load_2d_a(a_k + sg * kFaTileBytes, &tm_qkv, k_full + 8 * sg, H + st.head * 64,
                          st.begin + st.key_base + i * kFaBlockN);
              OPV_PP_STAMP(i, 24);
              ++kc;
            }
            if (i >= 1) {
              const uint32_t sg = vc % S;
              mbar_wait_a(v_empty + 8 * sg, ((vc / S) & 1) ^ 1);
              mbar_expect_tx_a(v_full + 8 * sg, kFaTileBytes);
              tma_load_2d_a(a_v + sg * kFaTileBytes, &tm_qkv, v_full + 8 * sg, 2 * H + st.head * 64,
                          st.begin + st.key_base + (i - 1) * kFaBlockN);
              OPV_PP_STAMP(i - 1, 25);
              ++vc;
            }
          }
          tracing = false;
        }
      }
    } else if ((warp == 17 || warp == 18) && elect_one()) {
      // ------------------------------ MMA issuers: warp 17 for tile A, warp 18 for tile B (ONE thread each) -------
      // Everything here is on the critical path of its tile (S_j(i+1) follows s_empty_j(i), P_j(i).V follows
      // p_full_j(i)).  Every tcgen05.mma / commit / try_wait travels through the scheduler's MIO queue, which the
      // softmax warps of the same scheduler keep full of MUFU.EX2 (r2 traces: ~130 cycles per dependent trip, a single
      // issuer needed 3400 cycles for the 24 MMAs + 7 commits + 7 waits of one key block of both tiles).  Hence one
      // issuer per tile, on two different schedulers, each a single thread with lean waits and nothing in local memory.
      const int j = warp - 17;
      constexpr uint32_t idesc_s = umma_idesc_bf16_f32(kFaBlockM, kFaBlockN);  // Q.K^T: both K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16_f32_bmn(kFaBlockM, 64);     // P.V: V is MN-major
      const uint32_t q_addr = a_q + j * kFaTileBytes, k_base = a_k, v_base = a_v;
      const uint32_t t_s = tmem_base + j * 128, t_p = tmem_base + 256 + j * 64, t_o = tmem_base + 384 + j * 64;
      const uint32_t my_q_full = q_full + 8 * j;
      const uint32_t my_q_empty = q_empty + 8 * j;
      const uint32_t my_s_full = s_full + 8 * j;
      const uint32_t my_s_empty = s_empty + 8 * j;
      const uint32_t my_p_full = p_full + 8 * j;
      const uint32_t my_pv_done = pv_done + 8 * j;
      const uint32_t my_o_empty = o_empty + 8 * j;
      uint32_t td = 0, sc = 0, pc = 0, kc = 0, vc = 0;
      bool tracing_mma = trace != nullptr && blockIdx.x == 0;
      Super st;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        if (!decode(t, st)) continue;
        const int lo = st.lo(j), hi = st.hi(j);
        for (int i = 0; i <= st.nb; ++i) {
          if (i < st.nb) {  // S_j(i) = Q_j . K(i)^T when this tile's range holds block i
            const uint32_t sg = kc % S;
            // also when this tile skips the block: the wait keeps this issuer from arriving on k_empty[sg] for the
            // NEXT use of the stage before the other issuer has arrived for this one (2 arrivals complete a phase)
            mbar_wait_a(k_full + 8 * sg, (kc / S) & 1);
            if (i >= lo && i < hi) {
              const uint32_t k_addr = k_base + sg * kFaTileBytes;
              if (i == lo) mbar_wait_a(my_q_full, td & 1);
              if (sc > 0) mbar_wait_a(my_s_empty, (sc - 1) & 1);  // the softmax warps have read the previous S_j
              tc_fence_after();
#pragma unroll
              for (int k = 0; k < (DBG == 3 ? 0 : 4); ++k)
                umma_bf16_ss(t_s, umma_desc_k_sw128(q_addr + k * 32), umma_desc_k_sw128(k_addr + k * 32), idesc_s,
                             k != 0 ? 1u : 0u);
              umma_commit_a(my_s_full);
              if (i == hi - 1) umma_commit_a(my_q_empty);
              if (tracing_mma) trace[i * 32 + 16 + j] = clock64();
              ++sc;
            }
            umma_commit_a(k_empty + 8 * sg);  // both issuers arrive for every block, whether their tile used it or not
            ++kc;
          }
          if (i >= 1) {  // O_j += P_j(b) . V(b)
            const int b = i - 1;
            const uint32_t sg = vc % S;
            mbar_wait_a(v_full + 8 * sg, (vc / S) & 1);
            if (b >= lo && b < hi) {
              const uint32_t v_addr = v_base + sg * kFaTileBytes;
              mbar_wait_a(my_p_full, pc & 1);
              if (b == lo && td > 0) mbar_wait_a(my_o_empty, (td - 1) & 1);  // epilogue has read the previous O_j
              tc_fence_after();
              const uint32_t first = b == lo ? 0u : 1u;
#pragma unroll
              for (int k = 0; k < (DBG == 3 ? 0 : 8); ++k)  // 16 keys per MMA: two 8-key groups of 1024 B
                umma_bf16_ts(t_o, t_p + k * 8, umma_desc_mn_sw128(v_addr + k * 2048), idesc_o, (k != 0) ? 1u : first);
              umma_commit_a(my_pv_done);
              if (tracing_mma) trace[b * 32 + 18 + j] = clock64();
              ++pc;
            }
            umma_commit_a(v_empty + 8 * sg);
            ++vc;
          }
        }
        if (hi > lo) ++td;
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
    const uint32_t t_other = lane_base + j * 128 + 64 * (half ^ 1);  // the other half's (row max only)
    const uint32_t t_p = lane_base + 256 + j * 64 + 32 * half;   // its 32 packed-probability columns
    const uint32_t t_o = lane_base + 384 + j * 64 + 32 * half;   // the 32 output columns it rescales / stores
    const float scale_log2 = 0.125f * 1.44269504088896340736f;   // head_dim^-0.5 * log2(e)
    const int pair_bar = 1 + j * 4 + quarter;                    // named barrier of the two warps sharing these rows
    // row max / row sum exchange between the two halves of a row; slots alternate with the block parity so that a
    // thread that runs ahead cannot overwrite a value its partner has not read yet
    const uint32_t my_slots = a_xchg + 4 * (j * 512 + half * 128 + r_tile);        // + 1024 B for the odd parity
    const uint32_t other_slots = a_xchg + 4 * (j * 512 + (half ^ 1) * 128 + r_tile);
    const uint32_t tok_wait = j == 0 ? x_a + 8 * quarter : x_b + 8 * quarter;   // exponential-phase token (see below)
    const uint32_t tok_pass = j == 0 ? x_b + 8 * quarter : x_a + 8 * quarter;
    uint32_t tok = 0;  // token phases consumed by this tile
    const uint32_t my_s_full = s_full + 8 * j;
    const uint32_t my_s_empty = s_empty + 8 * j;
    const uint32_t my_p_full = p_full + 8 * j;
    const uint32_t my_pv_done = pv_done + 8 * j;
    const uint32_t my_o_empty = o_empty + 8 * j;
    uint32_t bc = 0;  // running count of this tile's key blocks -> barrier parity
    tracing = tracing && half == 0 && quarter == 0;
    const int tslot = 8 * j;
    Super st;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      if (!decode(t, st)) continue;
      const int lo = st.lo(j), hi = st.hi(j);
      if (hi <= lo) continue;  // second tile of a super tile that ends inside the first one
      const bool use_token = DBG == 0 && global && st.hi1 > st.lo1;  // both tiles walk the same key blocks
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
        const int blk_key0 = st.key_base + i * kFaBlockN;
        // the block's four 32-key chunks, warp-uniform and the same in both halves of a row pair:
        // 0 = no row of this warp sees it, 1 = every row sees all of it, 2 = mixed
        auto kind_of = [&](const int q) {  // arithmetic only: an indexable array would live in local memory
          const int c_lo = blk_key0 + 32 * q, c_hi = c_lo + 31;
          return (c_hi < any_lo || c_lo > any_hi) ? 0 : ((c_lo >= all_lo && c_hi <= all_hi) ? 1 : 2);
        };
        const int own = 2 * half, oth = 2 * (half ^ 1);
        int k_own0 = 1, k_own1 = 1, k_oth0 = 1, k_oth1 = 1;
        // global layers: every block but a sequence's last one is fully visible -- one compare instead of four chunk
        // classifications (this loop is instruction-issue bound: ~480 instructions per 64 scores before this was trimmed)
        bool all_full = global && blk_key0 + kFaBlockN <= n;
        if (!all_full) {
          k_own0 = kind_of(own), k_own1 = kind_of(own + 1), k_oth0 = kind_of(oth), k_oth1 = kind_of(oth + 1);
          all_full = (k_own0 & k_own1 & k_oth0 & k_oth1) == 1;
        }
        const int key0 = blk_key0 + 64 * half;  // first key of this thread's 64
        // -inf for the keys of a 32-column chunk starting at key `first` that this row does not see
        auto mask32 = [&](uint32_t* v, const int first) {
          const int d = first - k_lo;
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (static_cast<uint32_t>(d + c) > k_span) v[c] = 0xff800000u;
        };
        auto max32 = [&](const uint32_t* v, float& a, float& b) {
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            a = fmax3(a, __uint_as_float(v[c + 0]), __uint_as_float(v[c + 1]));
            b = fmax3(b, __uint_as_float(v[c + 2]), __uint_as_float(v[c + 3]));
          }
        };

        mbar_wait_a(my_s_full, bc & 1);
        tc_fence_after();
        OPV_PP_STAMP(i, tslot + 0);
        // ---- scores: this thread's 64 columns stay in registers; the OTHER half's 64 are only read for the row max,
        // 32 at a time.  Both threads of a row thus compute the same full-row max on their own: no exchange through
        // shared memory and no pair barrier on the per-block path (three dependent trips through the MIO queue, each
        // ~130 cycles while the other tile's MUFU.EX2 stream keeps that queue full -- r2 trace).
        uint32_t sr[64], ot[32];
        if constexpr (DBG == 2) {
#pragma unroll
          for (int c = 0; c < 64; ++c) sr[c] = __float_as_uint(static_cast<float>((lane * 7 + c * 3 + i) & 31) * 0.01f);
#pragma unroll
          for (int c = 0; c < 32; ++c) ot[c] = sr[c];
        } else {
          tmem_ld_32x32b_x64_nowait(t_s, sr);
          tmem_ld_32x32b_x32_nowait(t_other, ot);
          tmem_wait_ld_fence64(sr);
          tmem_ld_fence32(ot);
        }
        OPV_PP_STAMP(i, tslot + 1);
        float mo0 = -CUDART_INF_F, mo1 = -CUDART_INF_F;
        if (all_full) {
          max32(ot, mo0, mo1);
        } else if (k_oth0 != 0) {
          if (k_oth0 == 2) mask32(ot, blk_key0 + 32 * oth);
          max32(ot, mo0, mo1);
        }
        if constexpr (DBG != 2) tmem_ld_32x32b_x32_nowait(t_other + 32, ot);  // in flight under the own-half max
        float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F;
        if (all_full) {
          max32(sr, mx0, mx1);
          max32(sr + 32, mx0, mx1);
        } else {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int kq = q ? k_own1 : k_own0;
            if (kq == 0) continue;
            if (kq == 2) mask32(sr + 32 * q, key0 + 32 * q);
            max32(sr + 32 * q, mx0, mx1);
          }
        }
        if constexpr (DBG != 2) tmem_wait_ld_fence32(ot);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(my_s_empty);  // the MMA warp may overwrite S_j with S_j(i+1)
        if (all_full) {
          max32(ot, mo0, mo1);
        } else if (k_oth1 != 0) {
          if (k_oth1 == 2) mask32(ot, blk_key0 + 32 * (oth + 1));
          max32(ot, mo0, mo1);
        }
        float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mo0, mo1));
        OPV_PP_STAMP_F(i, tslot + 2, mx);
        const float m_cand = fmaxf(m_run, mx * scale_log2);
        const bool upd = (m_cand - m_run) > kFaRescaleThreshold;  // false when both are -inf (NaN); same in both halves
        const float m_new = upd ? m_cand : m_run;
        const float corr = upd ? ex2_approx(m_run - m_new) : 1.0f;
        const float m_use = (m_new == -CUDART_INF_F) ? 0.f : m_new;
        m_run = m_new;

        // ---- O_j holds blocks < i and the P_j buffer is free again; lazy rescale of this thread's 32 output columns.
        // Done BEFORE the exponentials: this wait and the TMEM round trip then sit in the shadow of the other tile's
        // exponential phase instead of between this tile's last MUFU and its p_full.
        if (i > lo) {
          mbar_wait_a(my_pv_done, (bc - 1) & 1);
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

        // ---- exponential-phase token.  MUFU.EX2 (16 / clk / SM) bounds this kernel, and the four softmax warps of a
        // scheduler otherwise drift into phase: all of them issue exponentials at once (each at a quarter of the
        // rate), then all of them sit in TMEM round trips and barrier hand-offs with the unit idle (r2 trace: 1800
        // cycles of exponentials + 1250 of everything else per block, XU 62 % busy).  With both tiles walking the same
        // key blocks, tile B starts its exponentials of block i when tile A has issued (most of) its own, and tile A
        // those of block i+1 when B is through with block i: one tile's loads, row max, hand-offs and stores run under
        // the other tile's exponentials.  Per lane quarter (= per scheduler), 2 warp arrivals per phase.
        if (use_token && (j == 1 || i > lo)) {
          mbar_wait_a(tok_wait, tok & 1);
          ++tok;
        }
        // Packed fp32 pairs (FFMA2 / FADD2, new on sm_100): the scale-and-shift and the row-sum accumulation cost half
        // an issue slot per score instead of one each.  This loop is co-limited by MUFU.EX2 and by issue slots (r2
        // profile: 57 % issue utilisation, 62 % XU), so instructions per score are the lever.
        float2 acc01 = make_float2(0.f, 0.f), acc23 = make_float2(0.f, 0.f);
        const float2 sc2 = make_float2(scale_log2, scale_log2), nm2 = make_float2(-m_use, -m_use);
        auto exp8 = [&](uint32_t* pq, const int q, const int c0) {  // 16 scores -> 8 packed probability pairs
#pragma unroll
          for (int c = c0; c < c0 + 8; c += 2) {
            const float2 x01 = __ffma2_rn(make_float2(__uint_as_float(sr[32 * q + 2 * c]), __uint_as_float(sr[32 * q + 2 * c + 1])), sc2, nm2);
            const float2 x23 = __ffma2_rn(make_float2(__uint_as_float(sr[32 * q + 2 * c + 2]), __uint_as_float(sr[32 * q + 2 * c + 3])), sc2, nm2);
            const float2 p01 = make_float2(ex2(x01.x), ex2(x01.y));
            const float2 p23 = make_float2(ex2(x23.x), ex2(x23.y));
            acc01 = __fadd2_rn(acc01, p01);
            acc23 = __fadd2_rn(acc23, p23);
            pq[c] = pack_bf16x2(p01.x, p01.y);
            pq[c + 1] = pack_bf16x2(p23.x, p23.y);
          }
        };
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint32_t pq[16];
          const bool skip = !all_full && (q ? k_own1 : k_own0) == 0;  // no row of this warp sees these 32 keys
          if (skip) {
#pragma unroll
            for (int c = 0; c < 16; ++c) pq[c] = 0u;
          } else {
            exp8(pq, q, 0);
          }
          if (q == 1 && use_token && (j == 0 || i + 1 < hi)) {
            // three quarters of this block's exponentials are issued: hand the token over now, so that the other
            // tile's wake-up overlaps the tail (tile B's last block of a super tile has no successor to release)
            __syncwarp();
            if (lane == 0) mbar_arrive_a(tok_pass);
          }
          if (!skip) exp8(pq, q, 8);
          tmem_st_32x32b_x16_nowait(t_p + 16 * q, pq);  // drains under the remaining exponentials
        }
        const float2 acc = __fadd2_rn(acc01, acc23);
        float sum = acc.x + acc.y;
        OPV_PP_STAMP_F(i, tslot + 3, sum);
        l_run = l_run * corr + sum;
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(my_p_full);
        OPV_PP_STAMP(i, tslot + 5);
      }
      tracing = false;

      // epilogue: O / l -> bf16 -> out[begin + row, head*64 + 32*half : +32]
      st_shared_f32(my_slots + (bc & 1) * 1024, l_run);  // parity of the NEXT block: last used two blocks ago
      mbar_wait_a(my_pv_done, (bc - 1) & 1);
      tc_fence_after();
      uint32_t orr[32];
      tmem_ld_32x32_raw(t_o, orr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(my_o_empty);  // the next super tile's first P_j.V may overwrite O_j
      named_bar_sync(pair_bar, 64);
      const float l_total = l_run + ld_shared_f32(other_slots + (bc & 1) * 1024);
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
  if (warp == 19) tmem_dealloc(tmem_base, kPpTmemCols);
#undef OPV_PP_STAMP
#undef OPV_PP_STAMP_F
}

}  // namespace opv
