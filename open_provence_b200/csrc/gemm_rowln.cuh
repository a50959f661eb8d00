// Residual GEMM with the FOLLOWING LayerNorm computed from on-chip data (N = hidden size = 256 or 512):
//   R(fp32) = R + A[M,K] . W[N,K]^T        (HF:340-341: attn.Wo / mlp.Wo residual adds)
//   X(bf16) = LayerNorm(R) * w             (HF:318-341: mlp_norm after Wo, the next layer's attn_norm after mlp.Wo)
//
// Why a kernel of its own: the RESIDUAL epilogue of gemm_tcgen05.cuh adds in L2 (TMA reduce-add), so
// the SM never sees the new rows, and re-reading them "from L2" goes to DRAM at this streaming rate.  Here a CTA pair
// owns FULL rows: its two 256-column accumulators (N = 512; one for N = 256) are the whole row block.  The epilogue
//   pass 1   loads the old residual chunk by TMA (prefetched into L2 six chunks ahead), adds the accumulator,
//            stores the new residual by TMA, keeps it in TMEM in place of the accumulator, sums the row;
//   pass 1b  re-reads TMEM for the centred sum of squares (two-pass variance, like layernorm_kernel);
//   pass 2   re-reads TMEM, normalises, stores X by TMA;
// and only then hands the accumulators back.  For N = 512 that gives up the overlap of one row block's epilogue with the
// next one's MMAs -- affordable where the GEMM is HBM-bound (K = H: attn.Wo), not for mlp.Wo (K = I), which stays on
// the reduce-add kernel + standalone LayerNorm.  For N = 256 the other accumulator keeps the MMAs running.
// HBM bytes per token: A + 4N (old) + 4N (new) + 2N (X) against A + 8N + (4N + 2N) for GEMM + LayerNorm launches.
//
// Same producer / MMA-issuer structure as gemm_bf16_tcgen05_pair_kernel (cta_group::2, 256 x 256 tiles, group_rows
// tile order); 4 operand stages, 8 epilogue warps in 2 column groups with three 16 KB buffers each (load -> modify in
// place -> store, two loads ahead).  Measured, attn.Wo per step: 3 stages + 4 buffers 3.40-3.45 ms, 4 + 3: 3.27-3.35,
// 5 + 2 (one load ahead): 4.04.
#pragma once

#include "gemm_tcgen05.cuh"

namespace opv {

constexpr int kRowLnStages = 4;
constexpr int kRowLnGroups = 2;                  // epilogue groups of 4 warps; group g takes the 32-column chunks c = g (mod groups). Measured with 4 groups (16 warps, 2 buffers each, one load ahead): attn.Wo 3.45 -> 3.92 ms per step -- the depth of the load pipeline matters, not the number of warps
constexpr int kRowLnBufs = 3;                    // 16 KB buffers per epilogue group
constexpr int kRowLnLoadAhead = kRowLnBufs - 1;  // old-residual loads in flight per group
constexpr int kRowLnThreads = 64 + 128 * kRowLnGroups;
constexpr int kRowLnChunksPerTile = 8 / kRowLnGroups;  // "L" slots per 256-column tile and group
constexpr int kRowLnXPerTile = 4 / kRowLnGroups;       // "X" slots per tile and group

struct RowLnSmemLayout {
  static constexpr int kStageA = kGemmBlockM * kGemmBlockK * 2;
  static constexpr int kStageB = 128 * kGemmBlockK * 2;
  static constexpr int kTileBytes = kRowLnStages * (kStageA + kStageB);  // 128 KB
  static constexpr int kBufBytes = kRowLnGroups * kRowLnBufs * kGemmChunkBytes;  //  96 KB
  static constexpr int kStatBytes = kRowLnGroups * kGemmBlockM * 4;      // [group][row] fp32 partial sums
  static constexpr int kBarrierBytes = 192;
  static constexpr int kTotal = kTileBytes + kBufBytes + kStatBytes + kBarrierBytes + 1024;  // + alignment slack
  static constexpr int kTmemCols = 512;
};
static_assert(RowLnSmemLayout::kTotal <= 232448, "row-LN GEMM: shared memory budget");

__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1)
               : "memory");
}

// One 128 B row of a staging chunk (16 B pieces XOR-swizzled like TMA's SWIZZLE_128B), read side.
__device__ __forceinline__ void staging_read_row(const uint8_t* buf, int r, uint32_t (&w)[32]) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const uint4 v = *reinterpret_cast<const uint4*>(buf + r * 128 + ((g ^ (r & 7)) << 4));
    w[4 * g] = v.x, w[4 * g + 1] = v.y, w[4 * g + 2] = v.z, w[4 * g + 3] = v.w;
  }
}

// tm_r: fp32 [M, N] residual stream, 32-column x 128-row box (loads AND stores); tm_x: bf16 [M, N], 64-column box.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kRowLnThreads, 1)
gemm_rowln_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                       const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_x,
                       const float* __restrict__ ln_w, const float ln_eps, const int pdl_late, const int M, const int N,
                       const int K) {
  using L = RowLnSmemLayout;
  constexpr int BLOCK_N = 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw0 = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw0 + 1023u) & ~1023u) - raw0);  // 1024-aligned in the shared address space
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kRowLnStages * L::kStageA;
  uint8_t* bufs = smem + L::kTileBytes;                                      // [group][kRowLnBufs] x 16 KB
  float* stats = reinterpret_cast<float*>(smem + L::kTileBytes + L::kBufBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kTileBytes + L::kBufBytes + L::kStatBytes);
  uint64_t* empty_bar = full_bar + kRowLnStages;
  uint64_t* tmem_full = empty_bar + kRowLnStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* ld_full = tmem_empty + 2;  // [group][kRowLnBufs]: old-residual chunk landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ld_full + kRowLnGroups * kRowLnBufs);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int cta_rank = static_cast<int>(cluster_ctarank());
  const int num_pairs_m = (M + 2 * kGemmBlockM - 1) / (2 * kGemmBlockM);
  const int tiles_per_row = N / BLOCK_N;  // 1 or 2
  const int num_kb = K / kGemmBlockK;
  const int first_pair = static_cast<int>(cluster_id_x());
  const int pair_step = static_cast<int>(cluster_nctaid_x());

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_r);
    tma_prefetch_desc(&tm_x);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kRowLnStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 2 * 4 * kRowLnGroups);  // epilogue warps of both CTAs
    }
    for (int s = 0; s < kRowLnGroups * kRowLnBufs; ++s) mbar_init(&ld_full[s], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, L::kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (!pdl_late) pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int m_pair = first_pair; m_pair < num_pairs_m; m_pair += pair_step) {
      const int row0 = (m_pair * 2 + cta_rank) * kGemmBlockM;
      for (int n_blk = 0; n_blk < tiles_per_row; ++n_blk) {
        const int wrow0 = n_blk * BLOCK_N + cta_rank * 128;
        // (Measured and not kept: an L2 prefetch of the NEXT row block's A rows from here, to shorten the first tile after
        // the accumulators come back -- ncu: 31 % of the epilogue warps' time is the wait for that tile.  attn.Wo
        // 3.45 -> 3.70 ms per step, xsmall mlp.Wo 1.14 -> 1.43: the prefetched lines are evicted before their load.)
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * (L::kStageA + L::kStageB));
            tma_load_2d_pair(smem_a + stage * L::kStageA, &tm_a, &full_bar[stage], kb * kGemmBlockK, row0);
            tma_load_2d_pair(smem_b + stage * L::kStageB, &tm_b, &full_bar[stage], kb * kGemmBlockK, wrow0);
          }
          __syncwarp();
          if (++stage == kRowLnStages) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) --------------
    if (cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_f32(2 * kGemmBlockM, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int m_pair = first_pair; m_pair < num_pairs_m; m_pair += pair_step) {
        for (int n_blk = 0; n_blk < tiles_per_row; ++n_blk) {
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_base = smem_u32(smem_a + stage * L::kStageA);
              const uint32_t b_base = smem_u32(smem_b + stage * L::kStageB);
#pragma unroll
              for (int k = 0; k < kGemmBlockK / kUmmaK; ++k)
                umma_bf16_ss_pair(d_tmem, umma_desc_k_sw128(a_base + k * 32), umma_desc_k_sw128(b_base + k * 32), idesc,
                                  (kb | k) != 0 ? 1u : 0u);
              umma_commit_pair(&empty_bar[stage], 0b11);
              if (kb == num_kb - 1) umma_commit_pair(&tmem_full[acc], 0b11);
            }
            __syncwarp();
            if (++stage == kRowLnStages) stage = 0, phase ^= 1;
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------ epilogue warps (both CTAs) ----------------
    // kRowLnGroups groups of 4 warps (one warp per TMEM lane quarter) split the columns.
    // Per row block and group: n_load "L" slots (one 32-column fp32 chunk each: load old, add, store new) followed
    // by n_x "X" slots (one 64-column bf16 chunk of X each).  Slot u of the group's running count lives in buffer
    // u % kRowLnBufs.  Before the group barrier of slot u the leader waits until the store of slot u - 1 has read its
    // buffer (cp.async.bulk.wait_group.read 0; issued a whole slot earlier), so past the barrier every buffer but
    // slot u's is free: the leader loads slot u + kRowLnLoadAhead into one, the threads may write the next X slot.
    const int quarter = warp & 3;
    const int group = (warp - 2) >> 2;
    const int r_tile = quarter * 32 + lane;
    const bool leader = ((warp - 2) & 3) == 0 && lane == 0;
    const int bar_id = 1 + group;
    constexpr int kAllBar = 1 + kRowLnGroups, kAllThreads = 128 * kRowLnGroups;
    constexpr int CPT = kRowLnChunksPerTile, XPT = kRowLnXPerTile, G = kRowLnGroups;
    uint8_t* gbufs = bufs + group * kRowLnBufs * kGemmChunkBytes;
    uint64_t* gld = ld_full + group * kRowLnBufs;
    const int n_load = CPT * tiles_per_row, n_x = XPT * tiles_per_row, n_slots = n_load + n_x;
    const float inv_n = 1.0f / static_cast<float>(N);
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);

    // leader only: issue the old-residual load of slot `u` (row block u / n_slots of this CTA) if it is an L slot
    // `prefetch_only`: pull the chunk into L2 (a few us ahead -- a line prefetched a whole row block ahead is evicted
    // again before its load at this streaming rate)
    auto issue_load = [&](int u, bool prefetch_only) {
      const int k = u / n_slots, j = u - k * n_slots;
      const int m_pair = first_pair + k * pair_step;
      if (j >= n_load || m_pair >= num_pairs_m) return;
      const int col0 = (j / CPT) * BLOCK_N + (group + G * (j % CPT)) * 32;
      const int row0 = (m_pair * 2 + cta_rank) * kGemmBlockM;
      if (prefetch_only) {
        tma_prefetch_l2_2d(&tm_r, col0, row0);
        return;
      }
      const int b = u % kRowLnBufs;
      mbar_expect_tx(&gld[b], kGemmChunkBytes);
      tma_load_2d(gbufs + b * kGemmChunkBytes, &tm_r, &gld[b], col0, row0);
    };
    // n_slots divides 6 * tiles_per_row * ... : with 2 groups n_slots is 6 or 12, with 4 groups 3 or 6, so slot u + 6
    // is an L slot exactly when slot u's position in its row block is one: every L slot is prefetched once
    constexpr int kPrefetchAhead = 6;
    if (leader) {
      for (int i = kRowLnLoadAhead; i < kPrefetchAhead; ++i) issue_load(i, true);
      for (int i = 0; i < kRowLnLoadAhead; ++i) issue_load(i, false);
    }
    int u = 0;            // running slot count of this group
    uint32_t ld_par = 0;  // bit b = parity of the next load completion of buffer b
    int acc_first = 0;    // accumulator buffer of this row block's first tile
    uint32_t acc_phase = 0;
    for (int m_pair = first_pair; m_pair < num_pairs_m; m_pair += pair_step) {
      const int row0 = (m_pair * 2 + cta_rank) * kGemmBlockM;
      // ---- pass 1: new residual = old + accumulator; keep it in TMEM; row sum
      float s1a = 0.f, s1b = 0.f;
      for (int j = 0; j < n_load; ++j, ++u) {
        const int t = j / CPT, c = group + G * (j % CPT);
        const int acc = (acc_first + t) & 1;
        if (j % CPT == 0) {
          mbar_wait(&tmem_full[acc], acc_phase);  // accumulator of tile t complete
          tc_fence_after();
        }
        const int b = u % kRowLnBufs;
        uint8_t* buf = gbufs + b * kGemmChunkBytes;
        mbar_wait(&gld[b], (ld_par >> b) & 1u);
        ld_par ^= 1u << b;
        uint32_t a[32], o[32];
        const uint32_t taddr = lane_base + acc * BLOCK_N + c * 32;
        tmem_ld_32x32_raw(taddr, a);
        staging_read_row(buf, r_tile, o);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float v0 = __uint_as_float(o[i]) + __uint_as_float(a[i]);
          const float v1 = __uint_as_float(o[i + 1]) + __uint_as_float(a[i + 1]);
          s1a += v0;
          s1b += v1;
          o[i] = __float_as_uint(v0);
          o[i + 1] = __float_as_uint(v1);
        }
        staging_write_row(buf, r_tile, o);
        tmem_st_32x32b_x32(taddr, o);
        fence_proxy_async_smem();
        if (leader) tma_store_wait_read<0>();  // slot u - 1's store has read its buffer (see above)
        named_bar_sync(bar_id, 128);
        if (leader) {
          tma_store_2d(&tm_r, buf, t * BLOCK_N + c * 32, row0);
          tma_store_commit();
          issue_load(u + kRowLnLoadAhead, false);
          issue_load(u + kPrefetchAhead, true);
        }
      }
      // ---- mean, then pass 1b: centred sum of squares over this thread's chunks.  One statistics array: a barrier
      // after every write AND after every read keeps the next write behind all readers.
      stats[group * kGemmBlockM + r_tile] = s1a + s1b;
      named_bar_sync(kAllBar, kAllThreads);
      float tot = 0.f;
#pragma unroll
      for (int g = 0; g < G; ++g) tot += stats[g * kGemmBlockM + r_tile];
      const float mean = tot * inv_n;
      named_bar_sync(kAllBar, kAllThreads);
      float s2a = 0.f, s2b = 0.f;
      for (int j = 0; j < n_load; ++j) {
        const int acc = (acc_first + j / CPT) & 1;
        uint32_t a[32];
        tmem_ld_32x32_raw(lane_base + acc * BLOCK_N + (group + G * (j % CPT)) * 32, a);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float d0 = __uint_as_float(a[i]) - mean, d1 = __uint_as_float(a[i + 1]) - mean;
          s2a += d0 * d0;
          s2b += d1 * d1;
        }
      }
      stats[group * kGemmBlockM + r_tile] = s2a + s2b;
      named_bar_sync(kAllBar, kAllThreads);
      tot = 0.f;
#pragma unroll
      for (int g = 0; g < G; ++g) tot += stats[g * kGemmBlockM + r_tile];
      const float rstd = 1.0f / sqrtf(tot * inv_n + ln_eps);
      named_bar_sync(kAllBar, kAllThreads);
      // ---- pass 2: X = (R - mean) * rstd * w as bf16, 64 columns per slot
      for (int j = 0; j < n_x; ++j, ++u) {
        const int xc = group + G * j;  // 64-column chunk of the row
        const int acc = (acc_first + (xc >> 2)) & 1;
        const uint32_t taddr = lane_base + acc * BLOCK_N + (xc & 3) * 64;
        uint32_t w[32];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32];
          tmem_ld_32x32_raw(taddr + half * 32, v);
          const float4* g4 = reinterpret_cast<const float4*>(ln_w + xc * 64 + half * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 ga = __ldg(g4 + i);
            w[16 * half + 2 * i] = pack_bf16x2((__uint_as_float(v[4 * i]) - mean) * rstd * ga.x,
                                               (__uint_as_float(v[4 * i + 1]) - mean) * rstd * ga.y);
            w[16 * half + 2 * i + 1] = pack_bf16x2((__uint_as_float(v[4 * i + 2]) - mean) * rstd * ga.z,
                                                   (__uint_as_float(v[4 * i + 3]) - mean) * rstd * ga.w);
          }
        }
        uint8_t* buf = gbufs + (u % kRowLnBufs) * kGemmChunkBytes;  // free: see the slot comment above
        staging_write_row(buf, r_tile, w);
        fence_proxy_async_smem();
        if (leader) tma_store_wait_read<0>();
        named_bar_sync(bar_id, 128);
        if (leader) {
          tma_store_2d(&tm_x, buf, xc * 64, row0);
          tma_store_commit();
          issue_load(u + kRowLnLoadAhead, false);
          issue_load(u + kPrefetchAhead, true);
        }
        // ---- an accumulator goes back to the MMA warp as soon as this warp has read its last column of it
        if (j % XPT == XPT - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_pair_leader(&tmem_empty[(acc_first + j / XPT) & 1]);
        }
      }
      if (tiles_per_row == 2) {
        acc_phase ^= 1;  // both buffers used once per row block
      } else {
        acc_first ^= 1;
        if (acc_first == 0) acc_phase ^= 1;
      }
    }
    if (leader) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair(tmem_base, L::kTmemCols);
}

}  // namespace opv
