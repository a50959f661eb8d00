// bf16 GEMM for the projections / FFN on the 5th-gen tensor cores:
//   C = epilogue(A[M,K] . W[N,K]^T),  A = activations (K-major), W = nn.Linear weight (K-major)
//
// Persistent, warp-specialised CTA (one per SM):
//   warp 0        TMA producer   cp.async.bulk.tensor 2D tiles (128B swizzle) -> 4/6-stage smem ring
//   warp 1        MMA issuer     one lane issues tcgen05.mma (128 x BLOCK_N x 16), accumulators in TMEM,
//                                double-buffered (2 x BLOCK_N fp32 columns) so the epilogue of tile i
//                                overlaps the mainloop of tile i+1
//   warps 2..5    epilogue group 0   tcgen05.ld TMEM -> registers -> fused op -> swizzled smem staging
//   warps 6..9    epilogue group 1   tile -> ONE TMA store (or TMA reduce-add) per 128-row x 128-byte chunk
// All global traffic goes through TMA: the r1a profile showed the per-thread row stores of the first
// version (32 different 128 B lines per warp instruction) bound by L1TEX, not by HBM or the tensor pipe.
// Fused epilogues (SURVEY.md section 7 hard part 2: the H=512 GEMMs are HBM-bound unless the
// elementwise work rides in the epilogue):
//   STORE     C bf16
//   ROPE      rotate-half RoPE on the q,k column thirds with per-row positions (HF:205-228)
//   RESIDUAL  R(fp32) += acc, performed in L2 by cp.reduce.async.bulk.tensor .add.f32  (HF:340-341)
//   GEGLU     C[:, j] = gelu_erf(acc[:, in_j]) * acc[:, gate_j]  (HF:90-91), W rows interleaved per 128;
//             gelu through the branch-free exp2 form of the Gaussian CDF (common.cuh: gelu_fast)
//   RESIDUAL_LN  (CTA-pair kernel) RESIDUAL, then the NEXT LayerNorm (HF:318-341: mlp_norm after Wo, the following
//             layer's attn_norm after Wo2) of the finished rows while they are still in L2: x = LN(R) * w as bf16
#pragma once

#include "common.cuh"
#include "pointwise.cuh"

namespace opv {

constexpr int kGemmBlockM = 128;
constexpr int kGemmBlockK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmChunkBytes = 128 * 128;  // staging chunk: 128 rows x 128 B (64 bf16 or 32 fp32 columns)

enum : int { kEpiStore = 0, kEpiRope = 1, kEpiResidual = 2, kEpiGeglu = 3, kEpiResidualLn = 4 };

// Two epilogue groups (8 warps) for every epilogue: the r1b profile showed the K = 512 GEMMs waiting on a
// single 4-warp epilogue (tensor pipe 64 % active on Wqkv+RoPE); the groups take alternate 128 B chunks.
__host__ __device__ constexpr int gemm_epi_groups(int /*epi*/) { return 2; }
constexpr int kLnWarps = 4;  // RESIDUAL_LN: warps 10..13 run the LayerNorm of finished row blocks beside the epilogue
__host__ __device__ constexpr int gemm_threads(int epi) {
  return 64 + 128 * gemm_epi_groups(epi) + (epi == 4 /* kEpiResidualLn */ ? 32 * kLnWarps : 0);
}

struct GemmEpilogueArgs {
  const int32_t* pos;  // ROPE: [M] position of each row inside its sequence
  const float* cos;    // ROPE: [max_pos, 32]
  const float* sin;    // ROPE: [max_pos, 32]
  int32_t rope_cols;   // ROPE: output columns < rope_cols (= 2H: q and k) are rotated
  int32_t rope_rows;   // ROPE: rows of the cos / sin tables; positions are clamped into [0, rope_rows) (0 = trusted)
  // CTA-pair kernel, ROPE: 1 = a cluster takes ALL column tiles of a 256-row block before moving to its next row
  // block, so the rows' cos|sin table lines are staged once per row block instead of once per rotated tile
  // (set by the host when there are at least as many row blocks as clusters)
  int32_t group_rows;
  int32_t pdl_late;  // 1 = do not release the dependent kernel early (large forwards, see engine.cu g_pdl_late)
  // RESIDUAL_LN: after the last column tile of a row block, x[row] = LN(R[row]) * ln_w for the block's rows
  const float* ln_w;    // [N] LayerNorm weight (no bias: ModernBERT norm_bias = false)
  const float* ln_r;    // [M, N] fp32, the tensor behind tm_c (the residual stream)
  __nv_bfloat16* ln_x;  // [M, N] bf16
  float ln_eps;
};

// ROPE epilogue: the cos|sin rows (2 x 128 B) of the tile's 128 tokens are staged in shared memory by
// coalesced loads (one pipeline stage is given up for the 32 KB); see rope_stage_rows().
constexpr int kRopeRowBytes = 256;
__host__ __device__ constexpr int gemm_rope_bytes(int epi) { return epi == kEpiRope ? kGemmBlockM * kRopeRowBytes : 0; }

template <int BLOCK_N, int EPI>
struct GemmSmemLayout {
  static constexpr int kStageA = kGemmBlockM * kGemmBlockK * 2;
  static constexpr int kStageB = BLOCK_N * kGemmBlockK * 2;
  static constexpr int kStages = ((BLOCK_N == 256) ? 4 : 6) - (EPI == kEpiRope ? 1 : 0);
  static constexpr int kTileBytes = kStages * (kStageA + kStageB);
  static constexpr int kStagingBytes = 2 * kGemmChunkBytes;  // one buffer per epilogue group
  static constexpr int kRopeBytes = gemm_rope_bytes(EPI);
  static constexpr int kBarrierBytes = 256;
  static constexpr int kTotal = kTileBytes + kStagingBytes + kRopeBytes + kBarrierBytes + 1024;  // + alignment slack
  static constexpr int kTmemCols = 2 * BLOCK_N;                                                   // 512 or 256
};

// One 128 B row of the staging chunk, 16 B pieces XOR-swizzled exactly like TMA's SWIZZLE_128B.
__device__ __forceinline__ void staging_write_row(uint8_t* buf, int r, const uint32_t (&w)[32]) {
#pragma unroll
  for (int g = 0; g < 8; ++g)
    *reinterpret_cast<uint4*>(buf + r * 128 + ((g ^ (r & 7)) << 4)) =
        make_uint4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
}


// RESIDUAL_LN hand-off from the epilogue-group leaders to the LayerNorm warps: `ready[s]` (2 arrivals, one per group
// leader) says every reduce-add into this CTA's row block number `blocks` is complete in L2; `done[s]` (kLnWarps
// arrivals) gives the slot back.  The leaders do not stall for the completion: they signal two chunks into the NEXT
// tile (cp.async.bulk.wait_group 2 then covers exactly the finished row block) or at the end of the kernel.
struct LnHandoff {
  uint64_t* ready;
  uint64_t* done;
  int blocks = 0;        // row blocks this CTA has finished issuing
  bool pending = false;  // the last finished row block has not been signalled yet
  __device__ __forceinline__ void signal() {  // group leader only
    const int s = blocks & 1;
    mbar_wait(&done[s], ((blocks >> 1) & 1) ^ 1);
    mbar_arrive(&ready[s]);
  }
};

// Epilogue-group state: which staging buffer is next and how many buffers the group owns.
template <int NBUF>
struct StagingRing {
  uint8_t* base;
  int next = 0;
  int bar_id;
  bool leader;  // the one thread that issues the TMA stores of this group

  // returns the buffer to fill; on return the TMA store that last read it has finished reading
  __device__ __forceinline__ uint8_t* acquire() {
    if (leader) tma_store_wait_read<NBUF - 1>();
    named_bar_sync(bar_id, 128);
    return base + next * kGemmChunkBytes;
  }
  // all 128 threads call after writing their row; the leader then issues `op(buf)`
  template <typename Issue>
  __device__ __forceinline__ void release(uint8_t* buf, Issue&& issue) {
    fence_proxy_async_smem();
    named_bar_sync(bar_id, 128);
    if (leader) {
      issue(buf);
      tma_store_commit();
    }
    next = (next + 1 == NBUF) ? 0 : next + 1;
  }
};

// ROPE: all 256 epilogue threads copy the cos|sin table rows of the tile's 128 tokens into shared memory.
// 16 lanes read one token's 256 B (cos row, then sin row) with 16 B loads, so a warp instruction touches 4
// lines instead of the 32 that "one lane = one token" reads cost (the r1c profile: 8 warps x 16 such loads per
// head made L1 the limiter of the Wqkv GEMM).  16 B pieces are XOR-swizzled on the low 3 row bits so both
// this store and the per-row reads of the epilogue are bank-conflict free.
__device__ __forceinline__ void rope_stage_rows(const GemmEpilogueArgs& ep, uint8_t* rope_cs, int epi_tid, int row0,
                                                int M) {
  const int piece = epi_tid & 15;  // 0..7 = cos, 8..15 = sin
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 16 + (epi_tid >> 4);
    const int64_t row = static_cast<int64_t>(row0) + r;
    int p = row < M ? __ldg(ep.pos + row) : 0;
    if (ep.rope_rows > 0) p = min(max(p, 0), ep.rope_rows - 1);  // rows no sequence covers hold no position
    const float* src = (piece < 8 ? ep.cos : ep.sin) + static_cast<int64_t>(p) * 32 + (piece & 7) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src));
    *reinterpret_cast<float4*>(rope_cs + r * kRopeRowBytes + ((piece ^ (r & 7)) << 4)) = v;
  }
}

template <int BLOCK_N, int EPI, int NBUF>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmEpilogueArgs& ep, const CUtensorMap* tm_c,
                                                   StagingRing<NBUF>& ring, const uint8_t* rope_cs, uint32_t taddr,
                                                   int r_tile, int64_t row, int M, int m_blk, int n_blk, int group,
                                                   LnHandoff* ln = nullptr) {
  const int row0 = m_blk * kGemmBlockM;
  if constexpr (EPI == kEpiStore) {
#pragma unroll 1
    for (int c = group; c < BLOCK_N / 64; c += gemm_epi_groups(EPI)) {
      float lo[32], hi[32];
      tmem_ld_32x32(taddr + c * 64, lo);
      tmem_ld_32x32(taddr + c * 64 + 32, hi);
      uint32_t w[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(lo[2 * i], lo[2 * i + 1]), w[16 + i] = pack_bf16x2(hi[2 * i], hi[2 * i + 1]);
      uint8_t* buf = ring.acquire();
      staging_write_row(buf, r_tile, w);
      const int col0 = n_blk * BLOCK_N + c * 64;
      ring.release(buf, [&](uint8_t* b) { tma_store_2d(tm_c, b, col0, row0); });
    }
  } else if constexpr (EPI == kEpiRope) {
    const bool rotate = n_blk * BLOCK_N < ep.rope_cols;  // tile-uniform: rope_cols (2H) is a multiple of BLOCK_N
    const uint8_t* cs_row = rope_cs + r_tile * kRopeRowBytes;
#pragma unroll 1
    for (int hd = group; hd < BLOCK_N / 64; hd += gemm_epi_groups(EPI)) {  // one 64-wide head per chunk
      float lo[32], hi[32];
      tmem_ld_32x32(taddr + hd * 64, lo);
      tmem_ld_32x32(taddr + hd * 64 + 32, hi);
      if (rotate) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 c = *reinterpret_cast<const float4*>(cs_row + ((i ^ (r_tile & 7)) << 4));
          const float4 sn = *reinterpret_cast<const float4*>(cs_row + (((8 + i) ^ (r_tile & 7)) << 4));
          const float cs_[4] = {c.x, c.y, c.z, c.w}, sn_[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float a = lo[4 * i + j], b = hi[4 * i + j];
            lo[4 * i + j] = a * cs_[j] - b * sn_[j];  // t*cos + rotate_half(t)*sin, first half
            hi[4 * i + j] = b * cs_[j] + a * sn_[j];  // second half
          }
        }
      }
      uint32_t w[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(lo[2 * i], lo[2 * i + 1]), w[16 + i] = pack_bf16x2(hi[2 * i], hi[2 * i + 1]);
      uint8_t* buf = ring.acquire();
      staging_write_row(buf, r_tile, w);
      const int col0 = n_blk * BLOCK_N + hd * 64;
      ring.release(buf, [&](uint8_t* b) { tma_store_2d(tm_c, b, col0, row0); });
    }
  } else if constexpr (EPI == kEpiResidual || EPI == kEpiResidualLn) {
#pragma unroll 1
    for (int c = group; c < BLOCK_N / 32; c += gemm_epi_groups(EPI)) {  // 32 fp32 columns = 128 B per row
      uint32_t w[32];
      tmem_ld_32x32_raw(taddr + c * 32, w);
      uint8_t* buf = ring.acquire();
      staging_write_row(buf, r_tile, w);
      const int col0 = n_blk * BLOCK_N + c * 32;
      ring.release(buf, [&](uint8_t* b) { tma_reduce_add_2d(tm_c, b, col0, row0); });
      if constexpr (EPI == kEpiResidualLn) {
        if (c == group + gemm_epi_groups(EPI) && ln->pending) {  // second chunk of this tile issued
          if (ring.leader) {
            tma_store_wait_complete<2>();  // all but this tile's two bulk groups: the previous row block is final in L2
            ln->signal();
          }
          ln->pending = false;
          ++ln->blocks;
        }
      }
    }
  } else {  // kEpiGeglu: tile columns [0,128) = "input", [128,256) = "gate" of the same 128 features
    static_assert(EPI != kEpiGeglu || BLOCK_N == 256, "GeGLU epilogue needs BLOCK_N = 256");
    // group g produces output columns [64g, 64g + 64) of this tile's 128
    uint32_t w[32];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float a[32], g[32];
      tmem_ld_32x32(taddr + group * 64 + half * 32, a);
      tmem_ld_32x32(taddr + 128 + group * 64 + half * 32, g);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        w[16 * half + i] = pack_bf16x2(gelu_fast(a[2 * i]) * g[2 * i], gelu_fast(a[2 * i + 1]) * g[2 * i + 1]);
    }
    uint8_t* buf = ring.acquire();
    staging_write_row(buf, r_tile, w);
    const int col0 = n_blk * 128 + group * 64;
    ring.release(buf, [&](uint8_t* b) { tma_store_2d(tm_c, b, col0, row0); });
  }
}

// RESIDUAL_LN: LayerNorm of the 128 rows [row0, row0 + 128) of R by the kLnWarps LayerNorm warps, one warp per row
// and four (N <= 512) or two rows in flight per warp (ld.global.cg: never a stale L1 line).  Same statistics code as
// layernorm_kernel (pointwise.cuh), hence bit-identical x.
// MEASURED (B200, base-130M, 131 072 tokens): correct, but NOT a win, so `ln_fuse` is off by
// default.  The idea was that the rows are re-read from L2 right after this CTA's own reduce-adds completed them.
// They are not there any more: with all 148 CTAs streaming, ~113 MB pass through the 126 MB L2 per round of row
// blocks (~16 us), so the re-reads go to DRAM (Wo: 404 -> 653-669 MB read per launch, also with an evict_last
// policy on the reduce-adds and with the LayerNorm signalled at once), and giving a CTA pair whole row blocks
// separates the two column tiles' reads of the same A rows in time (Wo2: 807 -> 1396-1468 MB).  Wo 2.52 -> 4.5,
// Wo2 4.25 -> 6.2 ms per step against 2.45 ms of standalone LayerNorm launches saved.
template <int VEC>
__device__ __forceinline__ void epilogue_ln_rows(const GemmEpilogueArgs& ep, int row0, int M, int lw, int lane) {
  constexpr int N = VEC * 128;
  constexpr int R = VEC <= 4 ? 4 : 2;  // rows in flight per warp (register budget: R * VEC float4)  // rows in flight per warp (register budget: R * VEC float4)
  float4 g[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) g[i] = __ldg(reinterpret_cast<const float4*>(ep.ln_w + (i * 32 + lane) * 4));
#pragma unroll 1
  for (int rb = lw; rb < kGemmBlockM; rb += kLnWarps * R) {
    float4 v[R][VEC];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int64_t row = min(static_cast<int64_t>(row0) + rb + kLnWarps * j, static_cast<int64_t>(M) - 1);  // clamped: loads are unconditional
#pragma unroll
      for (int i = 0; i < VEC; ++i) v[j][i] = __ldcg(reinterpret_cast<const float4*>(ep.ln_r + row * N + (i * 32 + lane) * 4));
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int64_t row = static_cast<int64_t>(row0) + rb + kLnWarps * j;
      if (row < M) {
        float mean, rstd;
        row_norm_stats<VEC>(v[j], ep.ln_eps, mean, rstd);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          float4 o;
          o.x = (v[j][i].x - mean) * rstd * g[i].x;
          o.y = (v[j][i].y - mean) * rstd * g[i].y;
          o.z = (v[j][i].z - mean) * rstd * g[i].z;
          o.w = (v[j][i].w - mean) * rstd * g[i].w;
          store4(ep.ln_x + row * N + (i * 32 + lane) * 4, o);
        }
      }
    }
  }
}


// tm_c: STORE/ROPE/GEGLU bf16 [M, ldc] with a 64-column x 128-row box; RESIDUAL fp32 [M, N] with a
// 32-column x 128-row box (both SWIZZLE_128B).  TMA clips rows >= M.
template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(gemm_threads(EPI), 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                         const __grid_constant__ CUtensorMap tm_c, const GemmEpilogueArgs ep, const int M,
                         const int N, const int K) {
  using L = GemmSmemLayout<BLOCK_N, EPI>;
  constexpr int kGroups = gemm_epi_groups(EPI);
  constexpr int kBufsPerGroup = 2 / kGroups;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + L::kStages * L::kStageA;
  uint8_t* staging = smem + L::kTileBytes;
  uint8_t* rope_cs = staging + L::kStagingBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kTileBytes + L::kStagingBytes + L::kRopeBytes);
  uint64_t* empty_bar = full_bar + L::kStages;
  uint64_t* tmem_full = empty_bar + L::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int num_m = (M + kGemmBlockM - 1) / kGemmBlockM;
  const int num_n = N / BLOCK_N;
  const int num_tiles = num_m * num_n;
  const int num_kb = K / kGemmBlockK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_c);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < L::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4 * kGroups);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, L::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (!ep.pdl_late) pdl_launch_dependents();
  pdl_wait();  // the prologue above overlapped the previous kernel; its outputs are visible from here on
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform for ptxas: a per-thread TMEM address makes every tcgen05.mma an ELECT / R2UR.BROADCAST waterfall loop

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], L::kStageA + L::kStageB);
          tma_load_2d(smem_a + stage * L::kStageA, &tm_a, &full_bar[stage], kb * kGemmBlockK, m_blk * kGemmBlockM);
          tma_load_2d(smem_b + stage * L::kStageB, &tm_b, &full_bar[stage], kb * kGemmBlockK, n_blk * BLOCK_N);
        }
        __syncwarp();
        if (++stage == L::kStages) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kGemmBlockM, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);  // TMA bytes have landed
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_base = smem_u32(smem_a + stage * L::kStageA);
          const uint32_t b_base = smem_u32(smem_b + stage * L::kStageB);
#pragma unroll
          for (int k = 0; k < kGemmBlockK / kUmmaK; ++k) {
            // advancing 16 bf16 (32 B) along K inside the 128 B swizzle row = +32 B on the start address
            umma_bf16_ss(d_tmem, umma_desc_k_sw128(a_base + k * 32), umma_desc_k_sw128(b_base + k * 32), idesc,
                         (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                      // smem slot free once these MMAs retire
          if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == L::kStages) stage = 0, phase ^= 1;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ------------------------------ epilogue warps ----------------------------
    const int quarter = warp & 3;         // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int group = (warp - 2) >> 2;    // 0 (warps 2..5) or 1 (warps 6..9)
    const int r_tile = quarter * 32 + lane;
    StagingRing<kBufsPerGroup> ring;
    ring.base = staging + group * kBufsPerGroup * kGemmChunkBytes;
    ring.bar_id = 1 + group;
    ring.leader = ((warp - 2) & 3) == 0 && lane == 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      if constexpr (EPI == kEpiRope) {
        // stage this tile's cos|sin rows while its mainloop is still running (TMEM is double-buffered)
        if (n_blk * BLOCK_N < ep.rope_cols) {
          named_bar_sync(3, 256);  // every epilogue thread is done reading the previous tile's rows
          rope_stage_rows(ep, rope_cs, threadIdx.x - 64, m_blk * kGemmBlockM, M);
          named_bar_sync(3, 256);
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int64_t row = static_cast<int64_t>(m_blk) * kGemmBlockM + r_tile;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
      gemm_epilogue_tile<BLOCK_N, EPI>(ep, &tm_c, ring, rope_cs, taddr, r_tile, row, M, m_blk, n_blk, group);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (ring.leader) tma_store_wait_all();  // smem must outlive the bulk stores that read it
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, L::kTmemCols);
}

// --------------------------------------------------------------------------------------------------
// CTA-pair version (cta_group::2), BLOCK_N = 256: a 2-CTA cluster on one TPC computes a 256 x 256 tile.
// Each CTA stages its own 128 rows of A and HALF of the W tile (128 of the 256 rows) per k-block, so the
// L2 -> SM operand traffic per FLOP drops by a third against the single-CTA kernel (which the r1c
// measurements showed capped near 1 PFLOP/s by L2 bandwidth: 384 KB of operands per 33.5 MFLOP tile).
// The even CTA's warp 1 issues tcgen05.mma.cta_group::2 (M = 256) for the pair; TMA bytes of both CTAs
// are counted on the even CTA's full barrier; tcgen05.commit multicasts the smem-slot / accumulator
// hand-offs to both CTAs; each CTA's epilogue drains its own 128 TMEM lanes exactly as above.
// --------------------------------------------------------------------------------------------------
// u-th tile of the cluster `first` (of `step` clusters).  Default: tile index first + u * step over the
// row-major [row block][column tile] list.  group_rows: all column tiles of row block first, then of first + step, ...
__device__ __forceinline__ bool pair_tile(int u, int first, int step, int num_tiles, int num_pairs_m, int num_n,
                                          int group_rows, int& m_pair, int& n_blk) {
  if (group_rows) {
    m_pair = first + (u / num_n) * step;
    n_blk = u % num_n;
    return m_pair < num_pairs_m;
  }
  const int tile = first + u * step;
  m_pair = tile / num_n;
  n_blk = tile % num_n;
  return tile < num_tiles;
}

template <int EPI>
struct GemmPairSmemLayout {
  static constexpr int kStageA = kGemmBlockM * kGemmBlockK * 2;  // 16 KB: this CTA's 128 rows of A
  static constexpr int kStageB = 128 * kGemmBlockK * 2;          // 16 KB: this CTA's half of the 256 W rows
  static constexpr int kStages = EPI == kEpiRope ? 5 : 6;
  static constexpr int kTileBytes = kStages * (kStageA + kStageB);
  static constexpr int kStagingBytes = 2 * kGemmChunkBytes;
  static constexpr int kRopeBytes = gemm_rope_bytes(EPI);
  static constexpr int kBarrierBytes = 256;
  static constexpr int kTotal = kTileBytes + kStagingBytes + kRopeBytes + kBarrierBytes + 1024;
  static constexpr int kTmemCols = 512;  // two 256-column accumulators per CTA
};

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(gemm_threads(EPI), 1)
gemm_bf16_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                              const __grid_constant__ CUtensorMap tm_c, const GemmEpilogueArgs ep, const int M,
                              const int N, const int K) {
  using L = GemmPairSmemLayout<EPI>;
  constexpr int BLOCK_N = 256;
  constexpr int kGroups = gemm_epi_groups(EPI);
  constexpr int kBufsPerGroup = 2 / kGroups;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + L::kStages * L::kStageA;
  uint8_t* staging = smem + L::kTileBytes;
  uint8_t* rope_cs = staging + L::kStagingBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kTileBytes + L::kStagingBytes + L::kRopeBytes);
  uint64_t* empty_bar = full_bar + L::kStages;
  uint64_t* tmem_full = empty_bar + L::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* ln_ready = tmem_empty + 4;  // RESIDUAL_LN: [2] row block final in L2 (one arrival per epilogue-group leader)
  uint64_t* ln_done = ln_ready + 2;     //              [2] row block normalised (one arrival per LayerNorm warp)

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int cta_rank = static_cast<int>(cluster_ctarank());  // 0 = leader (issues the MMAs)
  const int num_pairs_m = (M + 2 * kGemmBlockM - 1) / (2 * kGemmBlockM);
  const int num_n = N / BLOCK_N;
  const int num_tiles = num_pairs_m * num_n;  // 256 x 256 tiles
  const int num_kb = K / kGemmBlockK;
  const int first_tile = static_cast<int>(cluster_id_x());
  const int tile_step = static_cast<int>(cluster_nctaid_x());

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_c);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < L::kStages; ++s) {
      mbar_init(&full_bar[s], 1);   // leader's producer arrives (expect_tx covers both CTAs' bytes)
      mbar_init(&empty_bar[s], 1);  // multicast tcgen05.commit from the leader
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);                  // multicast tcgen05.commit from the leader
      mbar_init(&tmem_empty[s], 2 * 4 * kGroups);   // epilogue warps of BOTH CTAs (used on the leader only)
      mbar_init(&ln_ready[s], kGroups);
      mbar_init(&ln_done[s], kLnWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, L::kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // barrier inits of both CTAs visible before any remote arrive / multicast commit
  tc_fence_after();
  if (!ep.pdl_late) pdl_launch_dependents();
  pdl_wait();  // the prologue above overlapped the previous kernel; its outputs are visible from here on
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform for ptxas: a per-thread TMEM address makes every tcgen05.mma an ELECT / R2UR.BROADCAST waterfall loop

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int u = 0;; ++u) {
      int m_pair, n_blk;
      if (!pair_tile(u, first_tile, tile_step, num_tiles, num_pairs_m, num_n, ep.group_rows, m_pair, n_blk)) break;
      const int row0 = (m_pair * 2 + cta_rank) * kGemmBlockM;
      const int wrow0 = n_blk * BLOCK_N + cta_rank * 128;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * (L::kStageA + L::kStageB));
          tma_load_2d_pair(smem_a + stage * L::kStageA, &tm_a, &full_bar[stage], kb * kGemmBlockK, row0);
          tma_load_2d_pair(smem_b + stage * L::kStageB, &tm_b, &full_bar[stage], kb * kGemmBlockK, wrow0);
        }
        __syncwarp();
        if (++stage == L::kStages) stage = 0, phase ^= 1;
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
      for (int u = 0;; ++u) {
        int m_pair, n_blk;
        if (!pair_tile(u, first_tile, tile_step, num_tiles, num_pairs_m, num_n, ep.group_rows, m_pair, n_blk)) break;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);  // both CTAs' epilogues drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);  // both CTAs' TMA bytes have landed
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_base = smem_u32(smem_a + stage * L::kStageA);
            const uint32_t b_base = smem_u32(smem_b + stage * L::kStageB);
#pragma unroll
            for (int k = 0; k < kGemmBlockK / kUmmaK; ++k)
              umma_bf16_ss_pair(d_tmem, umma_desc_k_sw128(a_base + k * 32), umma_desc_k_sw128(b_base + k * 32), idesc,
                                (kb | k) != 0 ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage], 0b11);                      // slot free in both CTAs
            if (kb == num_kb - 1) umma_commit_pair(&tmem_full[acc], 0b11);  // accumulators complete in both
          }
          __syncwarp();
          if (++stage == L::kStages) stage = 0, phase ^= 1;
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp < 2 + 4 * kGroups) {
    // ------------------------------ epilogue warps (both CTAs) ----------------
    const int quarter = warp & 3;
    const int group = (warp - 2) >> 2;
    const int r_tile = quarter * 32 + lane;
    StagingRing<kBufsPerGroup> ring;
    ring.base = staging + group * kBufsPerGroup * kGemmChunkBytes;
    ring.bar_id = 1 + group;
    ring.leader = ((warp - 2) & 3) == 0 && lane == 0;
    LnHandoff ln;
    ln.ready = ln_ready, ln.done = ln_done;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = 0;; ++u) {
      int m_pair, n_blk;
      if (!pair_tile(u, first_tile, tile_step, num_tiles, num_pairs_m, num_n, ep.group_rows, m_pair, n_blk)) break;
      const int m_blk = m_pair * 2 + cta_rank;
      if constexpr (EPI == kEpiRope) {
        if (n_blk * BLOCK_N < ep.rope_cols && (!ep.group_rows || n_blk == 0)) {
          named_bar_sync(3, 256);
          rope_stage_rows(ep, rope_cs, threadIdx.x - 64, m_blk * kGemmBlockM, M);
          named_bar_sync(3, 256);
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int64_t row = static_cast<int64_t>(m_blk) * kGemmBlockM + r_tile;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
      gemm_epilogue_tile<BLOCK_N, EPI>(ep, &tm_c, ring, rope_cs, taddr, r_tile, row, M, m_blk, n_blk, group, &ln);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_pair_leader(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
      // group_rows order: the last column tile ends this CTA's 128-row block; its LayerNorm is handed over two chunks
      // into the next tile (gemm_epilogue_tile) or below
      if constexpr (EPI == kEpiResidualLn) ln.pending = n_blk == num_n - 1;
    }
    if (ring.leader) {
      tma_store_wait_all();
      if constexpr (EPI == kEpiResidualLn) {
        if (ln.pending) ln.signal();
      }
    }
  } else {
    // ------------------------------ LayerNorm warps (RESIDUAL_LN, both CTAs) --
    if constexpr (EPI == kEpiResidualLn) {
      const int lw = warp - (2 + 4 * kGroups);
      int blocks = 0;
      for (int u = 0;; ++u) {
        int m_pair, n_blk;
        if (!pair_tile(u, first_tile, tile_step, num_tiles, num_pairs_m, num_n, ep.group_rows, m_pair, n_blk)) break;
        if (n_blk != num_n - 1) continue;
        const int row0 = (m_pair * 2 + cta_rank) * kGemmBlockM;
        const int s = blocks & 1;
        mbar_wait(&ln_ready[s], (blocks >> 1) & 1);
        switch (N >> 7) {
          case 2: epilogue_ln_rows<2>(ep, row0, M, lw, lane); break;
          case 4: epilogue_ln_rows<4>(ep, row0, M, lw, lane); break;
          case 6: epilogue_ln_rows<6>(ep, row0, M, lw, lane); break;
          default: epilogue_ln_rows<8>(ep, row0, M, lw, lane); break;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ln_done[s]);
        ++blocks;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer can still touch its smem / TMEM
  if (warp == 2) tmem_dealloc_pair(tmem_base, L::kTmemCols);
}

}  // namespace opv
