// bf16 GEMM for the projections / FFN on the 5th-gen tensor cores:
//   C = epilogue(A[M,K] . W[N,K]^T),  A = activations (K-major), W = nn.Linear weight (K-major)
//
// Persistent, warp-specialised CTA (one per SM):
//   warp 0      TMA producer   cp.async.bulk.tensor 2D tiles (128B swizzle) -> 4/6-stage smem ring
//   warp 1      MMA issuer     one lane issues tcgen05.mma (128 x BLOCK_N x 16), accumulators in TMEM,
//                              double-buffered (2 x BLOCK_N fp32 columns) so the epilogue of tile i
//                              overlaps the mainloop of tile i+1
//   warps 2..5  epilogue       tcgen05.ld TMEM -> registers -> fused op -> global
// Fused epilogues (SURVEY.md section 7 hard part 2: the H=512 GEMMs are HBM-bound unless the
// elementwise work rides in the epilogue):
//   STORE     C bf16
//   ROPE      rotate-half RoPE on the q,k column thirds with per-row positions (HF:205-228)
//   RESIDUAL  R(fp32) += acc                                    (HF:340-341)
//   GEGLU     C[:, j] = gelu_erf(acc[:, in_j]) * acc[:, gate_j]  (HF:90-91), W rows interleaved per 128
#pragma once

#include "common.cuh"

namespace opv {

constexpr int kGemmBlockM = 128;
constexpr int kGemmBlockK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kGemmThreads = 192;
constexpr int kUmmaK = 16;

enum : int { kEpiStore = 0, kEpiRope = 1, kEpiResidual = 2, kEpiGeglu = 3 };

struct GemmEpilogueArgs {
  void* c;             // STORE/ROPE/GEGLU: bf16 [M, ldc]; RESIDUAL: fp32 [M, ldc] (read-modify-write)
  int64_t ldc;         // row pitch of c in elements
  const int32_t* pos;  // ROPE: [M] position of each row inside its sequence
  const float* cos;    // ROPE: [max_pos, 32]
  const float* sin;    // ROPE: [max_pos, 32]
  int32_t rope_cols;   // ROPE: output columns < rope_cols (= 2H: q and k) are rotated
};

template <int BLOCK_N>
struct GemmSmemLayout {
  static constexpr int kStageA = kGemmBlockM * kGemmBlockK * 2;
  static constexpr int kStageB = BLOCK_N * kGemmBlockK * 2;
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : 6;
  static constexpr int kTileBytes = kStages * (kStageA + kStageB);
  static constexpr int kBarrierBytes = 256;
  static constexpr int kTotal = kTileBytes + kBarrierBytes + 1024;  // + slack for the 1024 B alignment
  static constexpr int kTmemCols = 2 * BLOCK_N;                     // 512 or 256: power of two >= 32
};

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&v)[32]) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
    d4[i] = u;
  }
}

template <int BLOCK_N, int EPI>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmEpilogueArgs& ep, uint32_t taddr, int64_t row, int M,
                                                   int n0, int n_blk) {
  const bool row_ok = row < M;
  if constexpr (EPI == kEpiStore) {
    __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(ep.c) + row * ep.ldc + n0;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      float v[32];
      tmem_ld_32x32(taddr + c * 32, v);
      if (row_ok) store_bf16x32(crow + c * 32, v);
    }
  } else if constexpr (EPI == kEpiRope) {
    __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(ep.c) + row * ep.ldc + n0;
    const bool rotate = n0 < ep.rope_cols;  // tile-uniform: rope_cols (2H) is a multiple of BLOCK_N
    float cs[32], sn[32];
    if (rotate) {
      const int p = row_ok ? ep.pos[row] : 0;
      const float4* c4 = reinterpret_cast<const float4*>(ep.cos + static_cast<int64_t>(p) * 32);
      const float4* s4 = reinterpret_cast<const float4*>(ep.sin + static_cast<int64_t>(p) * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 a = __ldg(c4 + i), b = __ldg(s4 + i);
        cs[4 * i + 0] = a.x, cs[4 * i + 1] = a.y, cs[4 * i + 2] = a.z, cs[4 * i + 3] = a.w;
        sn[4 * i + 0] = b.x, sn[4 * i + 1] = b.y, sn[4 * i + 2] = b.z, sn[4 * i + 3] = b.w;
      }
    }
#pragma unroll 1
    for (int hd = 0; hd < BLOCK_N / 64; ++hd) {  // one 64-wide head per iteration
      float lo[32], hi[32];
      tmem_ld_32x32(taddr + hd * 64, lo);
      tmem_ld_32x32(taddr + hd * 64 + 32, hi);
      if (rotate) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float a = lo[i], b = hi[i];
          lo[i] = a * cs[i] - b * sn[i];  // t*cos + rotate_half(t)*sin, first half
          hi[i] = b * cs[i] + a * sn[i];  // second half
        }
      }
      if (row_ok) {
        store_bf16x32(crow + hd * 64, lo);
        store_bf16x32(crow + hd * 64 + 32, hi);
      }
    }
  } else if constexpr (EPI == kEpiResidual) {
    float* rrow = reinterpret_cast<float*>(ep.c) + row * ep.ldc + n0;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      float v[32];
      tmem_ld_32x32(taddr + c * 32, v);
      if (row_ok) {
        float4* r4 = reinterpret_cast<float4*>(rrow + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 r = r4[i];
          r.x += v[4 * i + 0], r.y += v[4 * i + 1], r.z += v[4 * i + 2], r.w += v[4 * i + 3];
          r4[i] = r;
        }
      }
    }
  } else {  // kEpiGeglu: tile columns [0,128) = "input", [128,256) = "gate" of the same 128 features
    static_assert(EPI != kEpiGeglu || BLOCK_N == 256, "GeGLU epilogue needs BLOCK_N = 256");
    __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(ep.c) + row * ep.ldc + n_blk * 128;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float a[32], g[32];
      tmem_ld_32x32(taddr + c * 32, a);
      tmem_ld_32x32(taddr + 128 + c * 32, g);
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] = gelu_erf(a[i]) * g[i];
      if (row_ok) store_bf16x32(crow + c * 32, a);
    }
  }
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                         const GemmEpilogueArgs ep, const int M, const int N, const int K) {
  using L = GemmSmemLayout<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + L::kStages * L::kStageA;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kTileBytes);
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
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < L::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);  // one arrive per epilogue warp
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
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
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
        if (lane == 0) {
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
    const int quarter = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int64_t row = static_cast<int64_t>(m_blk) * kGemmBlockM + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
      gemm_epilogue_tile<BLOCK_N, EPI>(ep, taddr, row, M, n_blk * BLOCK_N, n_blk);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, L::kTmemCols);
}

}  // namespace opv
