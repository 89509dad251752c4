// dalm_b200 — bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands fed by TMA).
//
//     D[M,N] = epilogue( alpha * A[M,K] . B[N,K]^T )            A, B bf16 row-major ("TN": both K-contiguous)
//
// This single kernel is every dense contraction of the training step: the linear layers of the encoder / decoder
// forward (B = W[out,in]), their dgrad (B = W^T[in,out], kept as a resident transposed copy - weights are frozen in
// PEFT mode and HBM is 180 GB), the LoRA adapters (folded into the K dimension: A = [x | x A_l^T], B = [W | s B_l])
// and the lm_head. It replaces the cuBLAS calls made by HF modeling code under
// dalm/models/rag_e2e_base_model.py:93,105 (reference) and the autograd backward of those.
//
// Structure (one persistent CTA per SM, 384 threads):
//   warp 0      : TMA producer   - cp.async.bulk.tensor 128B-swizzled tiles into a STAGES-deep smem ring
//   warp 1      : MMA issuer     - one elected lane issues tcgen05.mma (128 x BN x 16) into a double-buffered TMEM
//                                  accumulator, tcgen05.commit releases smem slots / publishes the accumulator
//   warp 2      : TMEM allocator
//   warps 4..11 : epilogue       - tcgen05.ld 32x32b -> registers -> alpha/bias/GELU/dropout/residual -> swizzled smem
//                                  -> TMA store (two groups of 4 warps interleave 128-byte-wide store blocks),
//                                  overlapped with the next tile's MMAs through the second accumulator stage
#include "common.cuh"
#include "ptx.cuh"
#include <cudaTypedefs.h>
#include <mutex>
#include <unordered_map>

namespace dalm {
using namespace ptx;

struct GemmEpilogue {
  void* out;            // [M, ldo]
  long long ldo;
  int out_f32;          // 0: bf16, 1: fp32
  const float* bias;    // [N] fp32 or nullptr
  const void* resid;    // [M, ldr] added after activation, or nullptr (may alias out)
  long long ldr;
  int resid_f32;
  int act;              // 0: none, 1: GELU(erf), 2: GELU backward - out = bf16(alpha*acc + bias) * gelu'(resid) (resid = the bf16
                        //    pre-activation; multiplied, not added): the dgrad GEMM of the FFN output projection emits d(pre) directly
  float alpha;
  int M, N, K;
  DropCfg drop;         // dropout on act(alpha*acc+bias) BEFORE the residual add (BertSelfOutput / BertOutput); p = 0 => off
  int group_m;          // tile rasterisation: bands of group_m m-tiles, n-tiles walked serpentine inside a band (0 = m-fastest)
  int fuse;             // 0: none; 1: SwiGLU forward - tile columns [0,128) = gate, [128,256) = up of the same 128 features (weight
                        // rows interleaved); besides `out` (gate|up, the backward's input) the tile's silu(gate)*up goes to tmap_out2
                        // 2: rotary position embedding (head_dim 128, HF rotate_half) on output columns < rope_cols
                        // 4: SwiGLU backward in the down-projection's dgrad: the accumulator is d(act) [M,F]; with gate / up read from
                        //    `resid` (= the interleaved gate|up buffer [M,2F], which is also `out`) the tile leaves as
                        //    d gate = d act * up * s(g)(1 + g(1 - s(g))) and d up = d act * silu(g), written in place over gate / up
                        // 3: GELU forward with both tensors kept - `out` = pre-activation (bf16, what the backward needs),
                        //    tmap_out2 = gelu(pre) (bf16, the next GEMM's operand): one launch instead of GEMM + a 2-pass kernel
  const float* rope_cos; const float* rope_sin;   // fuse == 2: fp32 [rope_L, 64]; the position of output row m is m % rope_L
  int rope_L, rope_cols;
  int l2_hints;         // TMA L2 eviction priorities: bit 0 = A loads evict_last (the panel the resident CTAs share across waves),
                        // bit 1 = B loads evict_first (streamed once per band), bit 2 = output stores evict_first
};

// Tile rasterisation. Persistent CTA i works on tiles i, i + grid, ...: the tiles resident at one moment are ~148 consecutive
// indices, and (L2 serving the CTAs that share a panel) DRAM sees each A panel / B panel of that footprint once per wave.
// m-fastest order makes the footprint num_m x 4 tiles: all of A is re-read by every wave (profiles/r01_gemm_v3_ncu_full_raw.csv:
// 652 MB for the 343 MB down-projection). Bands of group_m m-tiles make it ~square (group_m x 148/group_m), which minimises
// rows-of-A + rows-of-B per wave, and the serpentine n order lets consecutive waves of a band re-use its A panel from L2.
__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int gm, int& m_blk, int& n_blk) {
  if (gm <= 0) { m_blk = tile % num_m; n_blk = tile / num_m; return; }
  const int band_tiles = gm * num_n;
  const int band = tile / band_tiles;
  const int r = tile - band * band_tiles;
  const int m0 = band * gm;
  const int h = min(gm, num_m - m0);                            // the last band may be shorter
  const int n = r / h;
  m_blk = m0 + (r - n * h);
  n_blk = (band & 1) ? (num_n - 1 - n) : n;
}

// Epilogue math for one thread = one output row, 32 consecutive columns [col0, col0+32): alpha, bias, GELU, residual.
template <bool PRE = false>
__device__ __forceinline__ void epilogue_math(const GemmEpilogue& ep, const uint32_t* v, float* f, int row, int col0, bool row_ok,
                                              const float4* rpre = nullptr) {
  const int N = ep.N;
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) * ep.alpha;
  if (ep.bias) {
#pragma unroll
    for (int i = 0; i < 32; ++i) if (col0 + i < N) f[i] += __ldg(ep.bias + col0 + i);
  }
  if (ep.act == 1) {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = gelu_erf(f[i]);
  }
  if (ep.drop.p > 0.f) {
    const unsigned long long dstream = drop_stream(ep.drop);
    const unsigned long long base = ((unsigned long long)row * (unsigned long long)N + (unsigned long long)col0) >> 3;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float sc[8];
      drop_scale8(ep.drop, dstream, base + g, sc);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[g * 8 + j] *= sc[j];
    }
  }
  if constexpr (PRE) {                                          // fp32 residual already in registers (prefetched a block ahead)
#pragma unroll
    for (int g = 0; g < 8; ++g) { f[g * 4 + 0] += rpre[g].x; f[g * 4 + 1] += rpre[g].y; f[g * 4 + 2] += rpre[g].z; f[g * 4 + 3] += rpre[g].w; }
  } else if (ep.resid && row_ok) {
    if (ep.resid_f32) {
      const float* r = reinterpret_cast<const float*>(ep.resid) + (size_t)row * ep.ldr + col0;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        if (col0 + g * 4 < N) {
          const float4 t = *reinterpret_cast<const float4*>(r + g * 4);
          f[g * 4 + 0] += t.x; f[g * 4 + 1] += t.y; f[g * 4 + 2] += t.z; f[g * 4 + 3] += t.w;
        }
      }
    } else {
      const __nv_bfloat16* r = reinterpret_cast<const __nv_bfloat16*>(ep.resid) + (size_t)row * ep.ldr + col0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (col0 + g * 8 < N) {
          const bf16x8 t = *reinterpret_cast<const bf16x8*>(r + g * 8);
          float tf[8];
          unpack8(t, tf);
          if (ep.act == 2) {     // same roundings as the un-fused pair (bf16 dgrad output, then gelu_bwd_kernel): bit-identical results
#pragma unroll
            for (int i = 0; i < 8; ++i) f[g * 8 + i] = __bfloat162float(__float2bfloat16_rn(f[g * 8 + i])) * gelu_erf_grad(tf[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[g * 8 + i] += tf[i];
          }
        }
      }
    }
  }
}

// Drains one 128 x BN fp32 accumulator (this thread: TMEM lane = row_in_tile) to HBM through a 128B-swizzled staging tile
// and TMA stores: the 128 epilogue threads write their rows into shared memory (16-byte pieces XOR-swizzled by row % 8:
// conflict-free per quarter warp, and exactly the layout CU_TENSOR_MAP_SWIZZLE_128B expects), one thread issues
// cp.async.bulk.tensor stores of [128 rows x 128 bytes] boxes. HBM sees full 128-byte lines; ragged M / N edges are
// clipped by the TMA unit. Two groups of four epilogue warps (8 warps: TMEM lane quarter = warp % 4) take alternate
// store blocks, each with its own staging tile, so one group's TMA store overlaps the other's TMEM reads and math.
// (Direct per-thread row stores were measured at 1170 TFLOP/s vs 1540 TFLOP/s for the bare mainloop: profiles/.)
constexpr int kStageTileBytes = 128 * 128;
constexpr int kEpiGroups = 2;                                   // two groups of 4 epilogue warps split a tile's store blocks
constexpr int kGemmThreads = 128 + kEpiGroups * 128;            // warps 0-3: TMA / MMA / TMEM alloc / spare ; warps 4-11: epilogue
template <int BN, int SPECIAL = 0>        // SPECIAL 3: the GELU two-output epilogue (fuse == 3) is compiled into its own kernel
__device__ __forceinline__ void epilogue_drain_tile(const GemmEpilogue& ep, const CUtensorMap* tmap_out, const CUtensorMap* tmap_out2,
                                                    unsigned char* staging, int grp, uint32_t t_row, int row_in_tile, int tile_row0,
                                                    int tile_col0) {
  const int N = ep.N;
  const int row = tile_row0 + row_in_tile;
  const bool row_ok = row < ep.M;
  const int sb_cols = ep.out_f32 ? 32 : 64;                    // 128 bytes of output per row per store block
  const bool issuer = (threadIdx.x == 128 + grp * 128);        // first thread of this epilogue group
  unsigned char* tile = staging + grp * kStageTileBytes;       // one staging tile per group
  unsigned char* st = tile + row_in_tile * 128;
  const int sw = row_in_tile & 7;
  // fp32 residual (o_proj / down-projection: x_out = x + y): this thread's 128 bytes of the NEXT store block are fetched
  // while the current block is processed, so the HBM latency of the residual no longer sits in the block's serial chain
  // (tmem ld -> residual ld -> math -> st.shared -> TMA store), which made this epilogue the slowest GEMM shape in round 1
  const bool rpf = ep.out_f32 && ep.resid != nullptr && ep.resid_f32;
  float4 rnext[8];
  auto fetch_resid = [&](int cc) {
    const int cg0 = tile_col0 + cc;
    const float* r = reinterpret_cast<const float*>(ep.resid) + (size_t)row * ep.ldr + cg0;
#pragma unroll
    for (int g = 0; g < 8; ++g)
      rnext[g] = (row_ok && cc < BN && cg0 + g * 4 < N) ? *reinterpret_cast<const float4*>(r + g * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  if (rpf) fetch_resid(grp * sb_cols);
#pragma unroll 1
  for (int c = grp * sb_cols; c < BN; c += kEpiGroups * sb_cols) {   // the two groups interleave store blocks
    const int col0 = tile_col0 + c;
    if (col0 >= N) break;                                       // uniform across the group's 4 warps
    if (issuer) bulk_wait_read<0>();                            // the previous store has finished reading the staging tile
    named_bar_sync(1 + grp, 128);
    if (ep.out_f32) {
      uint32_t v[32]; float f[32];
      tmem_ld_32x32(t_row + (uint32_t)c, v);
      float4 rcur[8];
      if (rpf) {
#pragma unroll
        for (int g = 0; g < 8; ++g) rcur[g] = rnext[g];
        fetch_resid(c + kEpiGroups * sb_cols);
      }
      tmem_ld_wait();
      if (rpf) epilogue_math<true>(ep, v, f, row, col0, row_ok, rcur);
      else     epilogue_math<false>(ep, v, f, row, col0, row_ok);
#pragma unroll
      for (int g = 0; g < 8; ++g)
        *reinterpret_cast<float4*>(st + ((g ^ sw) << 4)) = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
    } else if (ep.fuse == 2 && col0 < ep.rope_cols) {
      // RoPE in the QKV epilogue: this 64-column block is one rotate_half HALF of a head (x1 = columns 0..63, x2 = 64..127 of the
      // head; tiles are 256 columns = two whole heads); its partner half sits 64 TMEM columns away in the same accumulator.
      //   x1' = x1 cos - x2 sin ,  x2' = x2 cos + x1 sin        (cos / sin of the row's position, element j = column % 64)
      // Replaces a separate in-place pass over the q|k columns of every layer's QKV output (32 x 34 us per cfg-3 step).
      const bool second = ((col0 >> 6) & 1) != 0;               // this block holds x2
      const int cp = second ? c - 64 : c + 64;
      const int pos = row_ok ? (row % ep.rope_L) : 0;
      const float* cs = ep.rope_cos + (size_t)pos * 64;
      const float* sn = ep.rope_sin + (size_t)pos * 64;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t vo[32], vq[32]; float f[32];
        tmem_ld_32x32(t_row + (uint32_t)(c + h * 32), vo);
        tmem_ld_32x32(t_row + (uint32_t)(cp + h * 32), vq);
        float cc[32], ss[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          *reinterpret_cast<float4*>(cc + i) = __ldg(reinterpret_cast<const float4*>(cs + h * 32 + i));
          *reinterpret_cast<float4*>(ss + i) = __ldg(reinterpret_cast<const float4*>(sn + h * 32 + i));
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float own = __uint_as_float(vo[i]), oth = __uint_as_float(vq[i]);
          f[i] = second ? fmaf(own, cc[i], oth * ss[i]) : fmaf(own, cc[i], -oth * ss[i]);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<bf16x8*>(st + (((h * 4 + g) ^ sw) << 4)) = pack8(f + g * 8);
      }
    } else if (SPECIAL == 4) {
      // SwiGLU backward (see GemmEpilogue::fuse): this block = 64 features f0.. of d(act); their gate columns sit at
      // 256*(f0/128) + f0%128 of the interleaved buffer, the up columns 128 further. Same arithmetic and roundings as
      // swiglu_bwd_kernel on a bf16 d(act), so the fused launch is bit-identical to the dgrad GEMM + that kernel.
      const int gcol = ((col0 >> 7) << 8) + (col0 & 127);
      const __nv_bfloat16* gp = reinterpret_cast<const __nv_bfloat16*>(ep.resid) + (size_t)row * ep.ldr + gcol;
      bf16x8 pu[8];                                            // d gate goes straight to the staging tile, d up waits in registers
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        bf16x8 gq[4], uq[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          gq[g] = row_ok ? *reinterpret_cast<const bf16x8*>(gp + h * 32 + g * 8) : bf16x8{};
          uq[g] = row_ok ? *reinterpret_cast<const bf16x8*>(gp + 128 + h * 32 + g * 8) : bf16x8{};
        }
        uint32_t v[32];
        tmem_ld_32x32(t_row + (uint32_t)(c + h * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float gg[8], uu[8], dg[8], du[8];
          unpack8(gq[g], gg);
          unpack8(uq[g], uu);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float d = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[g * 8 + k]) * ep.alpha));
            const float sg = 1.f / (1.f + __expf(-gg[k]));
            const float silu = gg[k] * sg;
            dg[k] = d * uu[k] * sg * (1.f + gg[k] * (1.f - sg));
            du[k] = d * silu;
          }
          *reinterpret_cast<bf16x8*>(st + (((h * 4 + g) ^ sw) << 4)) = pack8(dg);
          pu[h * 4 + g] = pack8(du);
        }
      }
      fence_proxy_async();
      named_bar_sync(1 + grp, 128);
      if (issuer) {
        tma_store_2d(tmap_out, tile, gcol, tile_row0);
        bulk_commit();
        bulk_wait_read<0>();
      }
      named_bar_sync(1 + grp, 128);
#pragma unroll
      for (int g = 0; g < 8; ++g) *reinterpret_cast<bf16x8*>(st + ((g ^ sw) << 4)) = pu[g];
      fence_proxy_async();
      named_bar_sync(1 + grp, 128);
      if (issuer) {
        tma_store_2d(tmap_out, tile, gcol + 128, tile_row0);
        bulk_commit();
      }
      continue;
    } else if (SPECIAL == 3) {
      // GELU forward, both tensors: the 64-column block goes out twice through the same staging tile - first the bf16
      // pre-activation, then gelu() of those ROUNDED values (exactly what gelu_fwd_kernel would read back from HBM).
      bf16x8 pk[8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (col0 + h * 32 < N) {
          uint32_t v[32]; float f[32];
          tmem_ld_32x32(t_row + (uint32_t)(c + h * 32), v);
          tmem_ld_wait();
          epilogue_math(ep, v, f, row, col0 + h * 32, row_ok);
#pragma unroll
          for (int g = 0; g < 4; ++g) pk[h * 4 + g] = pack8(f + g * 8);
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) pk[h * 4 + g] = bf16x8{};
        }
      }
#pragma unroll
      for (int g = 0; g < 8; ++g) *reinterpret_cast<bf16x8*>(st + ((g ^ sw) << 4)) = pk[g];
      fence_proxy_async();
      named_bar_sync(1 + grp, 128);
      if (issuer) {
        if (ep.l2_hints & 4) tma_store_2d_hint(tmap_out, tile, col0, tile_row0, l2_policy_evict_first());
        else                 tma_store_2d(tmap_out, tile, col0, tile_row0);
        bulk_commit();
      }
#pragma unroll
      for (int g = 0; g < 8; ++g) {                             // overlaps the TMA unit reading the staging tile
        float x[8];
        unpack8(pk[g], x);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = gelu_erf(x[i]);
        pk[g] = pack8(x);
      }
      if (issuer) bulk_wait_read<0>();
      named_bar_sync(1 + grp, 128);
#pragma unroll
      for (int g = 0; g < 8; ++g) *reinterpret_cast<bf16x8*>(st + ((g ^ sw) << 4)) = pk[g];
      fence_proxy_async();
      named_bar_sync(1 + grp, 128);
      if (issuer) {
        tma_store_2d(tmap_out2, tile, col0, tile_row0);
        bulk_commit();
      }
      continue;
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (col0 + h * 32 < N) {
          uint32_t v[32]; float f[32];
          tmem_ld_32x32(t_row + (uint32_t)(c + h * 32), v);
          tmem_ld_wait();
          epilogue_math(ep, v, f, row, col0 + h * 32, row_ok);
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<bf16x8*>(st + (((h * 4 + g) ^ sw) << 4)) = pack8(f + g * 8);
        }
      }
    }
    fence_proxy_async();                                        // generic-proxy smem writes -> visible to the TMA unit
    named_bar_sync(1 + grp, 128);
    if (issuer) {
      if (ep.l2_hints & 4) tma_store_2d_hint(tmap_out, tile, col0, tile_row0, l2_policy_evict_first());
      else                 tma_store_2d(tmap_out, tile, col0, tile_row0);
      bulk_commit();
    }
  }
  if constexpr (BN == 256) {
    if (ep.fuse == 1) {
      // SwiGLU forward: act[:, 128 n_blk + 64 j + ...] = silu(gate) * up from the fp32 accumulators (gate columns 64j.., up columns
      // 128 + 64j..): two more 128-byte-wide store blocks per tile, one per epilogue group. Replaces a separate pass that re-read
      // the 203 MB gate|up buffer (32 x 51 us per cfg-3 step).
      const int j = grp;
      const int acol0 = (tile_col0 >> 1) + j * 64;              // column of the [M, N/2] activation matrix
      if (issuer) bulk_wait_read<0>();
      named_bar_sync(1 + grp, 128);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t vg[32], vu[32]; float f[32];
        tmem_ld_32x32(t_row + (uint32_t)(j * 64 + h * 32), vg);
        tmem_ld_32x32(t_row + (uint32_t)(128 + j * 64 + h * 32), vu);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float g = __uint_as_float(vg[i]);
          f[i] = g / (1.f + __expf(-g)) * __uint_as_float(vu[i]);
        }
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) *reinterpret_cast<bf16x8*>(st + (((h * 4 + g4) ^ sw) << 4)) = pack8(f + g4 * 8);
      }
      fence_proxy_async();
      named_bar_sync(1 + grp, 128);
      if (issuer) {
        if (ep.l2_hints & 4) tma_store_2d_hint(tmap_out2, tile, acol0, tile_row0, l2_policy_evict_first());
        else                 tma_store_2d(tmap_out2, tile, acol0, tile_row0);
        bulk_commit();
      }
    }
  }
}

template <int BN> struct GemmCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int ACC_STAGES = 2;
  static constexpr int TMEM_COLS = ACC_STAGES * BN;              // 128 / 256 / 512: powers of two
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * kStageTileBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

// LAYOUT selects how the two bf16 operands lie in HBM (the tensor core reads either major directly; nothing is transposed):
//   0  "TN"  A[M,K], B[N,K]  both K-contiguous                 forward y = x W^T, and PEFT dgrad against resident W^T copies
//   1  "NN"  A[M,K], B[K,N]  B is MN-major (N-contiguous)      dgrad dx = dy W straight from W[out,in] (full fine-tuning:
//                                                              weights change every step, so no transposed copy is kept)
//   2  "wgrad" A[K,M], B[K,N] both MN-major                    dW[out,in] = dy^T x : contraction over the token rows
// MN-major stage tiles are stored [64 k-rows][64 elements = 128 B] per 64-wide chunk (8 KB, chunks LBO = 8 KB apart,
// 8-row groups SBO = 1 KB apart), each chunk one TMA box of the row-major source.
template <int BN, int LAYOUT, int SPECIAL = 0>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_out2, const GemmEpilogue ep) {
  using Cfg = GemmCfg<BN>;
  constexpr int BM = Cfg::BM, BK = Cfg::BK, STAGES = Cfg::STAGES;
  extern __shared__ unsigned char smem_raw[];
  // 128B swizzle atoms need 1024-byte aligned tile bases
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* staging = smem + STAGES * Cfg::STAGE_BYTES;             // 2 x [128 rows x 128 B], 1024-aligned
  uint64_t* full_bar   = reinterpret_cast<uint64_t*>(staging + 2 * kStageTileBytes);
  uint64_t* empty_bar  = full_bar + STAGES;
  uint64_t* tfull_bar  = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + Cfg::ACC_STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + Cfg::ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = ep.M, N = ep.N, K = ep.K;
  const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_out);
    if (ep.fuse == 1 || ep.fuse == 3) prefetch_tmap(&tmap_out2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < Cfg::ACC_STAGES; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4 * kEpiGroups); }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const bool hint_a = (ep.l2_hints & 1) != 0, hint_b = (ep.l2_hints & 2) != 0;
      const uint64_t pol_a = hint_a ? l2_policy_evict_last() : 0ull, pol_b = hint_b ? l2_policy_evict_first() : 0ull;
      auto load = [](void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, bool hinted, uint64_t pol) {
        if (hinted) tma_load_2d_hint(dst, tm, bar, c0, c1, pol);
        else        tma_load_2d(dst, tm, bar, c0, c1);
      };
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(tile, num_m, num_n, ep.group_m, m_blk, n_blk);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sa = smem + stage * Cfg::STAGE_BYTES;
          unsigned char* sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if constexpr (LAYOUT == 2) {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) load(sa + c * 8192, &tmap_a, &full_bar[stage], m_blk * BM + c * 64, kb * BK, hint_a, pol_a);
          } else {
            load(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM, hint_a, pol_a);
          }
          if constexpr (LAYOUT >= 1) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) load(sb + c * 8192, &tmap_b, &full_bar[stage], n_blk * BN + c * 64, kb * BK, hint_b, pol_b);
          } else {
            load(sb, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN, hint_b, pol_b);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN) | (LAYOUT == 2 ? (1u << 15) : 0u) | (LAYOUT >= 1 ? (1u << 16) : 0u);
      // k-step (16 elements of K) in 16-byte units of the descriptor start address: K-major +32 B inside the swizzle atom,
      // MN-major +16 rows * 128 B
      constexpr uint64_t a_step = LAYOUT == 2 ? 128 : 2, b_step = LAYOUT >= 1 ? 128 : 2;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);            // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);                  // TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t adesc = LAYOUT == 2 ? make_sw128_mnmajor_desc(sa, 8192, 1024) : make_sw128_kmajor_desc(sa);
          const uint64_t bdesc = LAYOUT >= 1 ? make_sw128_mnmajor_desc(sb, 8192, 1024) : make_sw128_kmajor_desc(sb);
          const int krem = K - kb * BK;
          const int ksteps = krem >= BK ? (BK / 16) : ((krem + 15) / 16);   // skip all-zero (OOB) k-slices
          for (int k = 0; k < ksteps; ++k) {
            umma_f16(d_tmem, adesc + a_step * (uint64_t)k, bdesc + b_step * (uint64_t)k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                       // smem slot free once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);                           // accumulator complete -> epilogue
        if (++acc == Cfg::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // =============================== epilogue ===============================
    const int q = warp & 3;                                     // TMEM lane quarter == warp % 4
    const int grp = (warp - 4) >> 2;                            // epilogue group 0 / 1
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(tile, num_m, num_n, ep.group_m, m_blk, n_blk);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      epilogue_drain_tile<BN, SPECIAL>(ep, &tmap_out, &tmap_out2, staging, grp, t_row, q * 32 + lane, m_blk * BM, n_blk * BN);
      // all TMEM reads of this warp are complete (wait::ld): hand the accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == Cfg::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    if ((threadIdx.x & 127) == 0) bulk_wait<0>();               // staging tiles must outlive their TMA reads
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ============================================================================================================
// CTA-pair variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x BN tile. Each CTA stages its own
// 128 rows of A and HALF (BN/2 rows) of B per k-block, so shared-memory fill + operand-read traffic per SM drops from
// 96 KB to 64 KB per 128x256x64 of MMA work - the single-CTA kernel above is shared-memory-bandwidth bound at ~70 % of
// the tensor peak. The leader CTA's elected thread issues tcgen05.mma.cta_group::2 (M = 256); completion is multicast
// to both CTAs' barriers; both CTAs run TMA producers and epilogues for their own half.
// ============================================================================================================
template <int BN, int ST = 0> struct Gemm2Cfg {
  static constexpr int BM = 128 /*per CTA*/, BK = 64, BNH = BN / 2;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BNH * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = ST > 0 ? ST : ((BN == 256) ? 6 : 8);
  static constexpr int ACC_STAGES = 2;
  static constexpr int TMEM_COLS = ACC_STAGES * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * kStageTileBytes + 1024 + 256;
};

template <int BN, int ST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ CUtensorMap tmap_out, const GemmEpilogue ep) {
  using Cfg = Gemm2Cfg<BN, ST>;
  constexpr int BM = Cfg::BM, BK = Cfg::BK, STAGES = Cfg::STAGES, BNH = Cfg::BNH;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* staging = smem + STAGES * Cfg::STAGE_BYTES;             // 2 x [128 rows x 128 B], 1024-aligned
  uint64_t* full_bar   = reinterpret_cast<uint64_t*>(staging + 2 * kStageTileBytes);
  uint64_t* empty_bar  = full_bar + STAGES;
  uint64_t* tfull_bar  = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + Cfg::ACC_STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + Cfg::ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int M = ep.M, N = ep.N, K = ep.K;
  const int num_m = (M + 2 * BM - 1) / (2 * BM), num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + BK - 1) / BK;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b); prefetch_tmap(&tmap_out); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    // tfull: one multicast commit per tile; tempty (used on the leader): 4 epilogue warps of EACH CTA arrive
    for (int a = 0; a < Cfg::ACC_STAGES; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8 * kEpiGroups); }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_holder, Cfg::TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // =============================== TMA producer (both CTAs) ===============================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m_blk = tile % num_m, n_blk = tile / num_m;
        const int row_a = m_blk * 2 * BM + (int)rank * BM;
        const int row_b = n_blk * BN + (int)rank * BNH;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);             // own copy, released by the multicast commit
          unsigned char* sa = smem + stage * Cfg::STAGE_BYTES;
          unsigned char* sb = sa + Cfg::A_BYTES;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);   // both CTAs' bytes
          tma_load_2d_2sm(sa, &tmap_a, &full_bar[stage], kb * BK, row_a);
          tma_load_2d_2sm(sb, &tmap_b, &full_bar[stage], kb * BK, row_b);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) ===============================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);            // both CTAs' epilogues drained this stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t adesc = make_sw128_kmajor_desc(sa);
          const uint64_t bdesc = make_sw128_kmajor_desc(sb);
          const int krem = K - kb * BK;
          const int ksteps = krem >= BK ? (BK / 16) : ((krem + 15) / 16);
          for (int k = 0; k < ksteps; ++k)
            umma_f16_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage], 0x3);              // frees the slot in BOTH CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&tfull_bar[acc], 0x3);                  // accumulator halves ready in both CTAs
        if (++acc == Cfg::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // =============================== epilogue (both CTAs, own 128 rows) ===============================
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m_blk = tile % num_m, n_blk = tile / num_m;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      epilogue_drain_tile<BN>(ep, &tmap_out, &tmap_out, staging, grp, t_row, q * 32 + lane, m_blk * 2 * BM + (int)rank * BM, n_blk * BN);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tempty_bar[acc], 0);   // leader's barrier
      if (++acc == Cfg::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    if ((threadIdx.x & 127) == 0) bulk_wait<0>();
  }

  tc_fence_before();
  cluster_sync_all();                                           // peer may still be reading our smem / signalling us
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side: tensor-map construction (driver entry point fetched at run time, no libcuda link dependency) + cache
// ------------------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

struct TmapKey {
  const void* ptr; long long rows, cols, ld; int box_rows; int f32;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && f32 == o.f32;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h = h * 1000003u ^ (size_t)k.rows; h = h * 1000003u ^ (size_t)k.cols;
    h = h * 1000003u ^ (size_t)k.ld;   h = h * 1000003u ^ (size_t)(k.box_rows * 2 + k.f32);
    return h;
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;
static std::mutex g_tmap_mu;

// row-major [rows, cols] (bf16, or fp32 when f32 != 0) with row stride ld (elements);
// box = {128 bytes of columns, box_rows}, 128B swizzle, OOB loads -> 0, OOB stores clipped
int get_tmap(const void* ptr, long long rows, long long cols, long long ld, int box_rows, CUtensorMap* out, int f32) {
  TmapKey key{ptr, rows, cols, ld, box_rows, f32};
  {
    std::lock_guard<std::mutex> g(g_tmap_mu);
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) { *out = it->second; return 0; }
  }
  auto fn = get_encode_fn();
  DALM_REQUIRE(fn != nullptr, "gemm: cuTensorMapEncodeTiled driver entry point unavailable");
  DALM_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "gemm: operand base %p is not 16-byte aligned", ptr);
  DALM_REQUIRE((ld % (f32 ? 4 : 8)) == 0, "gemm: row stride %lld is not a multiple of 16 bytes", ld);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (f32 ? 4 : 2)};
  cuuint32_t box[2] = {f32 ? 32u : 64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUtensorMap m;
  CUresult r = fn(&m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DALM_REQUIRE(r == CUDA_SUCCESS, "gemm: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box_rows=%d",
               (int)r, rows, cols, ld, box_rows);
  {
    std::lock_guard<std::mutex> g(g_tmap_mu);
    if (g_tmaps.size() > 8192) g_tmaps.clear();
    g_tmaps[key] = m;
  }
  *out = m;
  return 0;
}

template <int BN, int LAYOUT = 0, int SPECIAL = 0>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const GemmEpilogue& ep,
                       int max_ctas, cudaStream_t stream, const CUtensorMap* to2 = nullptr) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    DALM_CUDA(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, LAYOUT, SPECIAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int num_tiles = ((ep.M + 127) / 128) * ((ep.N + BN - 1) / BN);
  int grid = num_tiles < kNumSMs ? num_tiles : kNumSMs;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  gemm_bf16_tn_kernel<BN, LAYOUT, SPECIAL><<<grid, kGemmThreads, Cfg::SMEM_BYTES, stream>>>(ta, tb, to, to2 ? *to2 : to, ep);
  count_launch();
  return check_launch("gemm_bf16_tn_kernel");
}

template <int BN, int ST = 0>
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const GemmEpilogue& ep,
                        int max_ctas, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN, ST>;
  static bool attr_set = false;
  if (!attr_set) {
    DALM_CUDA(cudaFuncSetAttribute(gemm2_bf16_tn_kernel<BN, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int num_tiles = ((ep.M + 255) / 256) * ((ep.N + BN - 1) / BN);
  int clusters = num_tiles < kNumSMs / 2 ? num_tiles : kNumSMs / 2;
  if (max_ctas > 0 && clusters > max_ctas / 2) clusters = max_ctas / 2 > 0 ? max_ctas / 2 : 1;
  gemm2_bf16_tn_kernel<BN, ST><<<2 * clusters, kGemmThreads, Cfg::SMEM_BYTES, stream>>>(ta, tb, to, ep);
  count_launch();
  return check_launch("gemm2_bf16_tn_kernel");
}

}  // namespace dalm

using namespace dalm;

extern "C" int dalm_b200_gemm_bf16(int layout, const void* A, long long lda, const void* B, long long ldb, void* out,
                                   long long ldo, int out_f32, int M, int N, int K, float alpha, const float* bias,
                                   int act, const void* resid, long long ldr, int resid_f32, int block_n,
                                   int max_ctas, float drop_p, unsigned long long drop_seed,
                                   unsigned long long drop_stream_id, const void* drop_offset, void* stream);

// tile rasterisation override: -1 = m-fastest order everywhere, 0 = automatic (default: pick_group_m below), -2 = the round-2a rule
// (bands for every multi-wave problem), > 0 = that many m-tiles per band. Initial value: env DALM_B200_GEMM_RASTER.
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}
static int g_group_m_override = env_int("DALM_B200_GEMM_RASTER", 0);
extern "C" void dalm_b200_gemm_set_raster(int group_m) { g_group_m_override = group_m; }
// TMA L2 eviction hints (GemmEpilogue::l2_hints bit mask): -1 = automatic (default), 0..7 = that mask on every launch.
// Initial value: env DALM_B200_GEMM_L2_HINTS.
static int g_l2_hints = env_int("DALM_B200_GEMM_L2_HINTS", -1);
extern "C" void dalm_b200_gemm_set_l2_hints(int mask) { g_l2_hints = mask < 0 ? -1 : (mask & 7); }
// Automatic choice, from profiles/r02b_gemm_l2_probe.txt (ncu DRAM bytes + CUDA-event times per raster x hint mask, cfg-3 shapes):
// in the m-fastest regime (A [M,K] <= 40 MB, bf16 output) "A evict_last, B and stores evict_first" brings DRAM traffic from
// 1.11-1.19x to 0.96-0.99x of the algorithmic bytes at unchanged time (QKV 289 -> 242 MB, gate|up 622 -> 514, lm_head 663 -> 592);
// with banded rasters (A too big for L2, or fp32 output + residual streams) every mask cost 2-5 % time and did not cut traffic.
static int pick_l2_hints(int M, int N, int K, int tile_n, int group_m, bool stream_out) {
  if (g_l2_hints >= 0) return g_l2_hints;
  const long long tiles = (long long)((M + 127) / 128) * ((N + tile_n - 1) / tile_n);
  return (group_m == 0 && tiles > kNumSMs && 2.0 * M * K <= 40e6 && !stream_out) ? 7 : 0;
}

// tile-shape heuristic: estimated time = waves of 148 CTAs x tile width x an efficiency penalty for narrow tiles (a
// 128 x BN tile re-reads its A operand from shared memory for every BN columns: profiles/r01_gemm_probe_tiles.jsonl).
// The 128x256 tile wins whenever the problem has more than about one wave of work - including N = 1024 at 3 204
// encoder rows, where 104 tiles in ONE wave take 32 us against 54 us for 208 half-width tiles in two waves
// (profiles/r01_gemm_shapes_in_step.txt had those shapes at 280-410 TFLOP/s). The CTA-pair kernel (block_n 2128/2256) is
// correct and tested but not faster on these shapes, so never auto-picked.
static int pick_block_n(int M, int N) {
  const long long m1 = (M + 127) / 128;
  double best = 1e30;
  int bn = 64;
  const int cand[3] = {256, 128, 64};
  const double penalty[3] = {1.0, 1.55, 2.7};      // measured at 26700x1024x4096: 203 / 313 / 549 us
  for (int i = 0; i < 3; ++i) {
    if (cand[i] > 64 && N < cand[i]) continue;                // do not pad N by more than one tile
    const long long tiles = m1 * ((N + cand[i] - 1) / cand[i]);
    const double cost = (double)((tiles + kNumSMs - 1) / kNumSMs) * cand[i] * penalty[i];
    if (cost < best) { best = cost; bn = cand[i]; }
  }
  return bn;
}

// Band height of the tile rasterisation (0 = m-fastest). One wave = the kNumSMs tiles resident at a time.
//  * m-fastest: a wave spans every m-tile of ~kNumSMs/num_m n-tiles, so all of A is touched by every wave. When A [M,K] is small
//    enough to stay in L2 across waves (measured: 38 MB does, next to a bf16 output stream) DRAM sees A once and every B panel
//    once - the algorithmic minimum (profiles/r01_gemm_v3: gate|up 447 MB vs 421, QKV 279 vs 252).
//  * bands of g m-tiles walked serpentine in n: only the band's rows of A have to stay resident, B is streamed once per band:
//    traffic ~ A + B * nbands. Wins when A does not fit (down-projection, K = 11008: 652 -> 559 MB) or when fp32 output + residual
//    streams (150 MB at cfg-3) push A out of L2 anyway (o_proj: 276 -> 240 MB).
// (profiles/r02_gemm_v4_ncu_full_raw.csv showed the cost of banding everything: B of gate|up / QKV read twice, 1.3x algorithmic.)
static int pick_group_m(int M, int N, int K, int tile_n, bool stream_out) {
  const int num_m = (M + 127) / 128, num_n = (N + tile_n - 1) / tile_n;
  int group_m = 0;
  if (g_group_m_override > 0) return g_group_m_override;
  if (g_group_m_override == -1 || (long long)num_m * num_n <= kNumSMs) return 0;
  const double a_bytes = 2.0 * M * K;
  if (g_group_m_override == 0 && a_bytes <= 40e6 && !stream_out) return 0;
  const double ideal = sqrt((double)kNumSMs * tile_n / 128.0);   // footprint of a wave ~square: min rows-of-A + rows-of-B
  int nbands = (int)(num_m / ideal + 0.5);
  if (nbands < 1) nbands = 1;
  group_m = (num_m + nbands - 1) / nbands;                        // bands equalised over num_m
  if (group_m >= num_m) group_m = 0;                              // one band == m-fastest
  return group_m;
}

// D[M,N] = act(alpha * A[M,K] B[N,K]^T + bias) + resid
//   A: bf16 [M,K] row stride lda;  B: bf16 [N,K] row stride ldb;  out: bf16|fp32 [M,N] row stride ldo
//   block_n: 0 = auto, or one of 64/128/256.   max_ctas: 0 = all SMs (used by tests to force multi-tile-per-CTA paths)
extern "C" int dalm_b200_gemm_bf16_tn(const void* A, long long lda, const void* B, long long ldb, void* out,
                                      long long ldo, int out_f32, int M, int N, int K, float alpha, const float* bias,
                                      int act, const void* resid, long long ldr, int resid_f32, int block_n,
                                      int max_ctas, float drop_p, unsigned long long drop_seed,
                                      unsigned long long drop_stream_id, const void* drop_offset, void* stream) {
  return dalm_b200_gemm_bf16(0, A, lda, B, ldb, out, ldo, out_f32, M, N, K, alpha, bias, act, resid, ldr, resid_f32, block_n,
                             max_ctas, drop_p, drop_seed, drop_stream_id, drop_offset, stream);
}

// layout 0: A[M,K] B[N,K] (TN)   1: A[M,K] B[K,N] (NN, dgrad from W[out,in])   2: A[K,M] B[K,N] (wgrad, contraction over rows)
extern "C" int dalm_b200_gemm_bf16(int layout, const void* A, long long lda, const void* B, long long ldb, void* out,
                                   long long ldo, int out_f32, int M, int N, int K, float alpha, const float* bias,
                                   int act, const void* resid, long long ldr, int resid_f32, int block_n,
                                   int max_ctas, float drop_p, unsigned long long drop_seed,
                                   unsigned long long drop_stream_id, const void* drop_offset, void* stream) {
  DALM_REQUIRE(layout >= 0 && layout <= 2, "gemm: layout must be 0 (TN), 1 (NN) or 2 (wgrad)");
  DALM_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  DALM_REQUIRE((N % 8) == 0, "gemm: N=%d must be a multiple of 8", N);
  DALM_REQUIRE((K % 8) == 0 || layout == 2, "gemm: K=%d must be a multiple of 8", K);
  DALM_REQUIRE((M % 8) == 0 || layout != 2, "gemm: M=%d must be a multiple of 8 for the wgrad layout", M);
  DALM_REQUIRE(lda >= (layout == 2 ? M : K) && ldb >= (layout >= 1 ? N : K) && ldo >= N, "gemm: leading dimensions too small");
  DALM_REQUIRE((ldo % (out_f32 ? 4 : 8)) == 0, "gemm: ldo=%lld breaks 16-byte row alignment", ldo);
  DALM_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "gemm: out is not 16-byte aligned");
  if (resid) {
    DALM_REQUIRE((ldr % (resid_f32 ? 4 : 8)) == 0, "gemm: ldr=%lld breaks 16-byte row alignment", ldr);
    DALM_REQUIRE((reinterpret_cast<uintptr_t>(resid) & 15) == 0, "gemm: resid is not 16-byte aligned");
  }
  DALM_REQUIRE(act == 0 || act == 1 || act == 2, "gemm: act must be 0 (none), 1 (gelu) or 2 (gelu backward: multiply by gelu'(resid))");
  DALM_REQUIRE(act != 2 || (resid != nullptr && !resid_f32 && !out_f32 && drop_p == 0.f),
               "gemm: act 2 (gelu backward) needs a bf16 `resid` (the pre-activation), a bf16 output and no dropout");
  int bn = block_n;
  if (bn == 0) bn = pick_block_n(M, N);
  DALM_REQUIRE(bn == 64 || bn == 128 || bn == 256 || bn == 2128 || bn == 2256 || bn == 3256 || bn == 4256,
               "gemm: block_n must be 0, 64/128/256 (single CTA) or 2128/2256 (CTA pair)");
  const bool pair = bn > 1000;
  DALM_REQUIRE(!(pair && layout != 0), "gemm: the CTA-pair kernel only takes the TN layout");
  const int tile_n = pair ? bn % 1000 : bn;
  CUtensorMap ta, tb, to;
  if (layout == 2) { if (int e = get_tmap(A, K, M, lda, 64, &ta)) return e; }
  else             { if (int e = get_tmap(A, M, K, lda, 128, &ta)) return e; }
  if (layout >= 1) { if (int e = get_tmap(B, K, N, ldb, 64, &tb)) return e; }
  else             { if (int e = get_tmap(B, N, K, ldb, pair ? tile_n / 2 : tile_n, &tb)) return e; }
  if (int e = get_tmap(out, M, N, ldo, 128, &to, out_f32)) return e;
  DALM_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "gemm: dropout p must be in [0,1)");
  const int group_m = pick_group_m(M, N, K, tile_n, out_f32 != 0 || resid != nullptr);
  GemmEpilogue ep{out, ldo, out_f32, bias, resid, ldr, resid_f32, act, alpha, M, N, K,
                  make_drop(drop_p, drop_seed, drop_stream_id, drop_offset), group_m, 0, nullptr, nullptr, 0, 0,
                  pick_l2_hints(M, N, K, tile_n, group_m, out_f32 != 0 || resid != nullptr)};
  cudaStream_t st = (cudaStream_t)stream;
  if (bn == 2256) return launch_gemm2<256>(ta, tb, to, ep, max_ctas, st);
  if (bn == 3256) return launch_gemm2<256, 3>(ta, tb, to, ep, max_ctas, st);     // tuning probes (fewer stages)
  if (bn == 4256) return launch_gemm2<256, 4>(ta, tb, to, ep, max_ctas, st);
  if (bn == 2128) return launch_gemm2<128>(ta, tb, to, ep, max_ctas, st);
  if (layout == 1) {
    if (bn == 256) return launch_gemm<256, 1>(ta, tb, to, ep, max_ctas, st);
    if (bn == 128) return launch_gemm<128, 1>(ta, tb, to, ep, max_ctas, st);
    return launch_gemm<64, 1>(ta, tb, to, ep, max_ctas, st);
  }
  if (layout == 2) {
    if (bn == 256) return launch_gemm<256, 2>(ta, tb, to, ep, max_ctas, st);
    if (bn == 128) return launch_gemm<128, 2>(ta, tb, to, ep, max_ctas, st);
    return launch_gemm<64, 2>(ta, tb, to, ep, max_ctas, st);
  }
  if (bn == 256) return launch_gemm<256>(ta, tb, to, ep, max_ctas, st);
  if (bn == 128) return launch_gemm<128>(ta, tb, to, ep, max_ctas, st);
  return launch_gemm<64>(ta, tb, to, ep, max_ctas, st);
}

// gate|up projection of LlamaMLP with SiLU(gate) * up fused into the epilogue (see include/dalm_b200.h)
extern "C" int dalm_b200_gemm_bf16_swiglu(const void* A, long long lda, const void* B, long long ldb, void* gu, long long ldgu,
                                          void* act, long long ldact, int M, int N, int K, void* stream) {
  DALM_REQUIRE(M > 0 && K > 0 && N >= 256 && (N % 256) == 0, "gemm_swiglu: N=%d must be a positive multiple of 256 (128-feature gate / up blocks)", N);
  DALM_REQUIRE((K % 8) == 0 && lda >= K && ldb >= K && ldgu >= N && ldact >= N / 2, "gemm_swiglu: bad K / leading dimensions");
  DALM_REQUIRE((ldgu % 8) == 0 && (ldact % 8) == 0 && ((uintptr_t)gu & 15) == 0 && ((uintptr_t)act & 15) == 0, "gemm_swiglu: output alignment");
  CUtensorMap ta, tb, to, to2;
  if (int e = get_tmap(A, M, K, lda, 128, &ta)) return e;
  if (int e = get_tmap(B, N, K, ldb, 256, &tb)) return e;
  if (int e = get_tmap(gu, M, N, ldgu, 128, &to, 0)) return e;
  if (int e = get_tmap(act, M, N / 2, ldact, 128, &to2, 0)) return e;
  const int group_m = pick_group_m(M, N, K, 256, false);
  GemmEpilogue ep{gu, ldgu, 0, nullptr, nullptr, 0, 0, 0, 1.f, M, N, K, make_drop(0.f, 0, 0, nullptr), group_m, 1, nullptr, nullptr, 0, 0,
                  pick_l2_hints(M, N, K, 256, group_m, false)};
  return launch_gemm<256>(ta, tb, to, ep, 0, (cudaStream_t)stream, &to2);
}

// down-projection dgrad of LlamaMLP with the SwiGLU backward in its epilogue: d(act) = dY WdT^T never reaches HBM; the interleaved
// gate|up buffer of the forward (gemm_bf16_swiglu) is overwritten in place with [d gate | d up] (what swiglu_bwd produced).
extern "C" int dalm_b200_gemm_bf16_swiglu_bwd(const void* dY, long long lddy, const void* WdT, long long ldw, void* gu, long long ldgu,
                                              int M, int F, int K, void* stream) {
  DALM_REQUIRE(M > 0 && K > 0 && F >= 256 && (F % 128) == 0, "gemm_swiglu_bwd: F=%d must be >= 256 and a multiple of 128 (interleave block)", F);
  DALM_REQUIRE((K % 8) == 0 && lddy >= K && ldw >= K && ldgu >= 2LL * F && (ldgu % 8) == 0 && ((uintptr_t)gu & 15) == 0,
               "gemm_swiglu_bwd: bad K / leading dimensions / alignment");
  CUtensorMap ta, tb, to;
  if (int e = get_tmap(dY, M, K, lddy, 128, &ta)) return e;
  if (int e = get_tmap(WdT, F, K, ldw, 256, &tb)) return e;
  if (int e = get_tmap(gu, M, 2LL * F, ldgu, 128, &to, 0)) return e;
  const int group_m = pick_group_m(M, F, K, 256, false);
  GemmEpilogue ep{gu, ldgu, 0, nullptr, gu, ldgu, 0, 0, 1.f, M, F, K, make_drop(0.f, 0, 0, nullptr), group_m, 4, nullptr, nullptr, 0, 0, 0};   // in-place output: no hints
  return launch_gemm<256, 0, 4>(ta, tb, to, ep, 0, (cudaStream_t)stream);
}

// intermediate projection of a GELU MLP (BertIntermediate, Falcon dense_h_to_4h) with the activation fused into the epilogue and
// BOTH tensors written: pre = A B^T + bias (bf16; gelu_bwd needs it) and act = gelu(pre) (bf16; the output projection's operand).
extern "C" int dalm_b200_gemm_bf16_gelu(const void* A, long long lda, const void* B, long long ldb, void* pre, long long ldpre,
                                        void* act, long long ldact, int M, int N, int K, const float* bias, void* stream) {
  DALM_REQUIRE(M > 0 && N > 0 && K > 0 && (N % 8) == 0 && (K % 8) == 0, "gemm_gelu: bad shape M=%d N=%d K=%d", M, N, K);
  DALM_REQUIRE(lda >= K && ldb >= K && ldpre >= N && ldact >= N, "gemm_gelu: leading dimensions too small");
  DALM_REQUIRE((ldpre % 8) == 0 && (ldact % 8) == 0 && ((uintptr_t)pre & 15) == 0 && ((uintptr_t)act & 15) == 0, "gemm_gelu: output alignment");
  const int bn = pick_block_n(M, N);
  CUtensorMap ta, tb, to, to2;
  if (int e = get_tmap(A, M, K, lda, 128, &ta)) return e;
  if (int e = get_tmap(B, N, K, ldb, bn, &tb)) return e;
  if (int e = get_tmap(pre, M, N, ldpre, 128, &to, 0)) return e;
  if (int e = get_tmap(act, M, N, ldact, 128, &to2, 0)) return e;
  const int group_m = pick_group_m(M, N, K, bn, false);
  GemmEpilogue ep{pre, ldpre, 0, bias, nullptr, 0, 0, 0, 1.f, M, N, K, make_drop(0.f, 0, 0, nullptr), group_m, 3, nullptr, nullptr, 0, 0,
                  pick_l2_hints(M, N, K, bn, group_m, false)};
  cudaStream_t st = (cudaStream_t)stream;
  if (bn == 256) return launch_gemm<256, 0, 3>(ta, tb, to, ep, 0, st, &to2);
  if (bn == 128) return launch_gemm<128, 0, 3>(ta, tb, to, ep, 0, st, &to2);
  return launch_gemm<64, 0, 3>(ta, tb, to, ep, 0, st, &to2);
}

// fused q|k|v projection + rotary embedding: out[M,N] = A[M,K] B[N,K]^T with HF's rotate_half RoPE (head_dim 128) applied to the
// output columns [0, rope_cols) in the epilogue. cos / sin: fp32 [L, 64]; row m sits at position m % L (token-major [B*L] rows).
extern "C" int dalm_b200_gemm_bf16_rope(const void* A, long long lda, const void* B, long long ldb, void* out, long long ldo, int M,
                                        int N, int K, const float* cos_t, const float* sin_t, int L, int rope_cols, void* stream) {
  DALM_REQUIRE(M > 0 && N > 0 && K > 0 && (N % 8) == 0 && (K % 8) == 0, "gemm_rope: bad shape M=%d N=%d K=%d", M, N, K);
  DALM_REQUIRE(rope_cols > 0 && rope_cols <= N && (rope_cols % 256) == 0, "gemm_rope: rope_cols=%d must be a multiple of 256 (whole 128-wide heads per tile)", rope_cols);
  DALM_REQUIRE(L > 0 && cos_t != nullptr && sin_t != nullptr && ((uintptr_t)cos_t & 15) == 0 && ((uintptr_t)sin_t & 15) == 0, "gemm_rope: cos / sin tables");
  DALM_REQUIRE(lda >= K && ldb >= K && ldo >= N && (ldo % 8) == 0 && ((uintptr_t)out & 15) == 0, "gemm_rope: leading dimensions / alignment");
  CUtensorMap ta, tb, to;
  if (int e = get_tmap(A, M, K, lda, 128, &ta)) return e;
  if (int e = get_tmap(B, N, K, ldb, 256, &tb)) return e;
  if (int e = get_tmap(out, M, N, ldo, 128, &to, 0)) return e;
  const int group_m = pick_group_m(M, N, K, 256, false);
  GemmEpilogue ep{out, ldo, 0, nullptr, nullptr, 0, 0, 0, 1.f, M, N, K, make_drop(0.f, 0, 0, nullptr), group_m, 2, cos_t, sin_t, L, rope_cols,
                  pick_l2_hints(M, N, K, 256, group_m, false)};
  return launch_gemm<256>(ta, tb, to, ep, 0, (cudaStream_t)stream);
}

// drop cached tensor maps (call when operand buffers are freed / re-allocated at the same address with other shapes)
extern "C" void dalm_b200_gemm_clear_cache() {
  std::lock_guard<std::mutex> g(g_tmap_mu);
  g_tmaps.clear();
}
