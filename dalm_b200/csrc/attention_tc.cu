// dalm_b200 — attention on the 5th-gen tensor cores (tcgen05.mma + TMEM accumulators + TMA), head_dim 128 or 64,
// causal / bidirectional + key-padding mask, MHA / GQA / MQA, optional attention-probability dropout. Forward and backward
// (dKdV, dQ).
//
// Replaces the attention inside HF LlamaForCausalLM (head_dim 128), BertModel (bge-large: 16 x 64, bidirectional,
// attention_probs_dropout_prob 0.1) and FalconForCausalLM (71 q heads / 1 kv head x 64, causal) reached through
// dalm/models/rag_e2e_base_model.py:93,105 / retriever_only_base_model.py:58 (reference) and their autograd backward; same
// math as csrc/attention.cu (flash-style, scores never in HBM, LSE saved, P recomputed).
//
// Every contraction is a 128 x 128 x 128 tcgen05 tile issued by ONE thread; the softmax / dS elementwise work runs on
// four warps that own one TMEM lane (= one query or key row) each:
//     forward : S = Q K^T        -> TMEM | P = exp2(S - m) -> bf16 -> 128B-swizzled smem | PV = P V          -> TMEM -> O (regs)
//     dKdV    : S^T = K Q^T, dP^T = V dO^T -> TMEM | P^T, dS^T -> smem | dV += P^T dO, dK += dS^T Q (TMEM accumulators)
//     dQ      : S = Q K^T, dP = dO V^T     -> TMEM | dS -> smem          | dQ += dS K                  (TMEM accumulator)
// The token-major [tokens][d] tiles that TMA drops into shared memory serve BOTH as K-major operands (contraction over d:
// QK^T, dO V^T) and, through MN-major descriptors, as the B operand of the contractions over tokens (P V, dS K, dS^T Q,
// P^T dO) - no transposes anywhere.
#include "common.cuh"
#include "ptx.cuh"

namespace dalm {
using namespace ptx;

struct AttnTcParams {
  const int64_t* mask;              // [B,L] key-padding mask or nullptr
  __nv_bfloat16* o; long long ldo;  // forward output [B*L, Hq*128]
  float* lse;                       // [B,Hq,L]
  const float* delta;               // [B,Hq,Lp] (backward)  rowsum(dO * O); Lp = L rounded up to 64, pad entries 0
  const float* nl2;                 // [B,Hq,Lp] (backward)  -lse * log2(e); pad entries (and fully masked rows) -inf
  int Lp;
  __nv_bfloat16* dq; __nv_bfloat16* dk; __nv_bfloat16* dv;
  long long lddq, lddk, lddv;
  int B, L, Hq, Hkv;
  float scale;
  int causal;
  int qcol0, kcol0, vcol0, ocol0;   // column of head 0 inside the q / k / v / dO tensor maps
  long long* dbg;                   // optional [64] clock64 timestamps of one CTA's phases (tuning aid), or nullptr
  DropCfg drop;                     // attention-probability dropout (BERT); p = 0 => off. Same element indexing as
                                    // csrc/attention.cu: group ((b*Hq + h)*L + q) * ceil(L/8) + key/8, component key%8
};

constexpr int TB = 128;                      // tile: 128 queries x 128 keys
constexpr int HALF_BYTES = TB * 64 * 2;      // one [128 rows x 64 cols] bf16 box = 16 KB
constexpr int TILE_BYTES = 2 * HALF_BYTES;   // [128 x 128] bf16 as two 64-column halves (P / dS tiles; Q/K/V/dO tiles at D = 128)
template <int D> struct TcD {                // per head_dim constants: a [128 tokens x D] tile is D/64 halves
  static constexpr int KS = D / 16;          // k-steps of the contractions over d (Q K^T, dO V^T)
  static constexpr int HALVES = D / 64;
  static constexpr int TILE = HALVES * HALF_BYTES;
};
template <int D>
__device__ __forceinline__ void load_tile(unsigned char* dst, const CUtensorMap* tm, uint64_t* bar, int col, int row) {
#pragma unroll
  for (int hh = 0; hh < D / 64; ++hh) tma_load_2d(dst + hh * HALF_BYTES, tm, bar, col + hh * 64, row);
}

// A/B K-major descriptors for k-step kk (0..7) of a [128 x 128] tile stored as two 64-col halves
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t tile, int kk) {
  return make_sw128_kmajor_desc(tile + (kk >> 2) * HALF_BYTES) + (uint64_t)(2 * (kk & 3));
}
// B MN-major descriptor for k-step kk (16 token rows) of a [128 tokens x 128 d] tile stored as two 64-col halves
__device__ __forceinline__ uint64_t mnmajor_desc(uint32_t tile, int kk) {
  return make_sw128_mnmajor_desc(tile + kk * 16 * 128, HALF_BYTES, 1024);
}
// store 8 bf16 (one 16-byte piece `piece` of row `row`) into a 128B-swizzled [128 x 64] half tile
__device__ __forceinline__ void st_sw128(unsigned char* half_tile, int row, int piece, const bf16x8& v) {
  *reinterpret_cast<bf16x8*>(half_tile + row * 128 + ((piece ^ (row & 7)) << 4)) = v;
}

// Drain one [128 rows x 128 cols] fp32 TMEM accumulator to a bf16 token-major matrix. Full tiles go through a
// 128B-swizzled staging tile (two 64-column halves) and two TMA stores - full 128-byte lines instead of 32 partial sectors
// per warp store (measured: the per-thread row stores cost 10k of a forward CTA's 30k cycles). Ragged tiles (rows beyond
// the sequence end belong to the NEXT sequence, so the TMA box must not be used) fall back to predicated row stores.
template <int TD>
__device__ __forceinline__ void drain_tile_bf16(uint32_t tacc, float scale, unsigned char* stage, const CUtensorMap* tm,
                                                int col0, int row0_global, bool full_tile, bool row_ok, int r,
                                                __nv_bfloat16* fallback_row) {
#pragma unroll 1
  for (int c = 0; c < TD; c += 32) {
    uint32_t v[32]; float f[32];
    tmem_ld_32x32(tacc + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) * scale;
    if (full_tile) {
#pragma unroll
      for (int g = 0; g < 4; ++g) st_sw128(stage + (c >> 6) * HALF_BYTES, r, ((c & 63) >> 3) + g, pack8(f + g * 8));
    } else if (row_ok) {
#pragma unroll
      for (int g = 0; g < 4; ++g) *reinterpret_cast<bf16x8*>(fallback_row + c + g * 8) = pack8(f + g * 8);
    }
  }
  if (full_tile) {
    fence_proxy_async();
    named_bar_sync(1, 128);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int hh = 0; hh < TD / 64; ++hh) tma_store_2d(tm, stage + hh * HALF_BYTES, col0 + hh * 64, row0_global);
      bulk_commit();
      bulk_wait<0>();
    }
  }
}

// key-validity bits of a 128-key tile, replicated in every lane (no shared memory, no barrier): word w, bit i = key
// kv0 + 32 w + i is a real, un-masked key
__device__ __forceinline__ void key_bits(const int64_t* mask_row, int kv0, int L, int lane, uint32_t* bits) {
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const int key = kv0 + w * 32 + lane;
    bool keep = key < L;
    if (keep && mask_row) keep = mask_row[key] != 0;
    bits[w] = __ballot_sync(0xffffffffu, keep);
  }
}

// ============================================================================================================
// forward: grid (ceil(L/128), Hq, B), 160 threads: warps 0-3 softmax (thread = query row), warp 4 control.
// O accumulates in TMEM (P V with the accumulate flag); the online-softmax correction rescales it in place.
// ============================================================================================================
template <int TD, bool DROP>
__global__ void __launch_bounds__(160, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, const AttnTcParams p) {
  using C = TcD<TD>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;
  unsigned char* sK = sQ + C::TILE;
  unsigned char* sV = sK + C::TILE;
  // D = 128: P overwrites K (K is dead once S = Q K^T has retired (s_full), and K is only reloaded after P V has retired
  // (kv_free)): 96 KB per CTA => two CTAs per SM. D = 64: K is only 16 KB, P (32 KB) gets its own tile: 80 KB per CTA.
  unsigned char* sP = TD == 128 ? sK : sV + C::TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>((TD == 128 ? sV + C::TILE : sP + TILE_BYTES));
  uint64_t *q_full = bars, *kv_full = bars + 1, *kv_free = bars + 2, *s_full = bars + 3, *p_ready = bars + 4, *pv_full = bars + 5;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.Hq / p.Hkv);
  const int L = p.L, q0 = qb * TB;
  const int tok0 = b * L;
  const int nkv = p.causal ? (min(L, q0 + TB) + TB - 1) / TB : (L + TB - 1) / TB;
  const bool dbg = p.dbg != nullptr && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0;
#define TS(slot) do { if (dbg) p.dbg[slot] = clock64(); } while (0)
  if (threadIdx.x == 0) TS(0);

  if (threadIdx.x == 128) {
    mbar_init(q_full, 1); mbar_init(kv_full, 1); mbar_init(kv_free, 1); mbar_init(s_full, 1);
    mbar_init(p_ready, 128); mbar_init(pv_full, 1);
    fence_mbar_init();
    prefetch_tmap(&tm_q); prefetch_tmap(&tm_k); prefetch_tmap(&tm_v);
  }
  if (warp == 0) { tmem_alloc(tmem_holder, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;
  const uint32_t tS = tmem, tO = tmem + 128;
  if (threadIdx.x == 0) TS(1);

  if (warp == 4) {
    if (lane == 0) {
      // ---------------- control: TMA loads + MMA issue ----------------
      mbar_arrive_expect_tx(q_full, C::TILE);
      load_tile<TD>(sQ, &tm_q, q_full, p.qcol0 + h * TD, tok0 + q0);
      constexpr uint32_t idesc_kk = make_idesc_bf16(TB, TB);          // S = Q K^T   (both K-major)
      constexpr uint32_t idesc_mn = make_idesc_bf16_bmn(TB, TD);      // O += P V    (B = V MN-major)
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
      for (int j = 0; j < nkv; ++j) {
        const uint32_t ph = j & 1;
        mbar_wait(kv_free, ph ^ 1);
        mbar_arrive_expect_tx(kv_full, 2 * C::TILE);
        load_tile<TD>(sK, &tm_k, kv_full, p.kcol0 + hk * TD, tok0 + j * TB);
        load_tile<TD>(sV, &tm_v, kv_full, p.vcol0 + hk * TD, tok0 + j * TB);
        TS(2 + j * 8);
        if (j == 0) mbar_wait(q_full, 0);
        mbar_wait(kv_full, ph);
        TS(3 + j * 8);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < C::KS; ++kk) umma_f16(tS, kmajor_desc(aQ, kk), kmajor_desc(aK, kk), idesc_kk, kk != 0);
        umma_commit(s_full);
        TS(4 + j * 8);
        mbar_wait(p_ready, ph);                                         // P written, O rescaled
        TS(5 + j * 8);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma_f16(tO, kmajor_desc(aP, kk), mnmajor_desc(aV, kk), idesc_mn, (j | kk) != 0);
        umma_commit(pv_full);
        umma_commit(kv_free);
        TS(6 + j * 8);
      }
    }
  } else {
    // ---------------- softmax warps: thread = query row; the whole 128-wide S row lives in registers ----------------
    const int r = warp * 32 + lane;
    const int qrow = q0 + r;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    const int64_t* mask_row = p.mask ? p.mask + (size_t)tok0 : nullptr;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkv; ++j) {
      const uint32_t ph = j & 1;
      const int kv0 = j * TB;
      uint32_t kb[4];
      key_bits(mask_row, kv0, L, lane, kb);
      const bool all_keys = (kb[0] & kb[1] & kb[2] & kb[3]) == 0xffffffffu;
      const bool diag = p.causal && (kv0 + TB - 1 > q0);              // only the diagonal tile needs the causal compare
      mbar_wait(s_full, ph);
      if (threadIdx.x == 0) TS(20 + j * 8);
      tc_fence_after();
      float sv[TB];
      {
        uint32_t* raw = reinterpret_cast<uint32_t*>(sv);
#pragma unroll
        for (int c = 0; c < TB; c += 32) tmem_ld_32x32(tS + lane_off + c, raw + c);
        tmem_ld_wait();
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (all_keys && !diag) {
#pragma unroll
        for (int i = 0; i < TB; ++i) { sv[i] *= sl2; mx4[i & 3] = fmaxf(mx4[i & 3], sv[i]); }
      } else {
        // branch-free masking (per-lane divergent selects along the diagonal were measured 5x slower): build an
        // all-ones integer mask for dropped elements and OR in the bit pattern of -inf
        const int dcol = diag ? (qrow - kv0) : 0x7fffffff;                 // columns > dcol are in the causal future
#pragma unroll
        for (int i = 0; i < TB; ++i) {
          const int drop = ((dcol - i) >> 31) | (int)(((kb[i >> 5] >> (i & 31)) & 1u) - 1u);   // -1 if dropped, else 0
          const uint32_t bits = (__float_as_uint(sv[i] * sl2) & ~(uint32_t)drop) | (0xff800000u & (uint32_t)drop);
          sv[i] = __uint_as_float(bits);
          mx4[i & 3] = fmaxf(mx4[i & 3], sv[i]);
        }
      }
      if (threadIdx.x == 0) TS(21 + j * 8);
      const float m_new = fmaxf(m_run, fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])));
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      const float corr = ex2_approx(m_run - m_safe);                   // m_run = -inf -> 0
      m_run = m_new;
      if (j > 0) {
        // O of the previous tiles (in TMEM) must carry the new maximum: rescale in place once P V (j-1) has retired
        mbar_wait(pv_full, ph ^ 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < TD; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tO + lane_off + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * corr);
          tmem_st_32x32(tO + lane_off + c, v);
        }
        tmem_st_wait();
      }
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
      unsigned long long dstream = 0, dgrp0 = 0;
      if constexpr (DROP) {
        dstream = drop_stream(p.drop);
        dgrp0 = (((unsigned long long)b * p.Hq + h) * L + (unsigned long long)qrow) * (unsigned long long)((L + 7) >> 3) + (unsigned long long)(kv0 >> 3);
      }
#pragma unroll
      for (int g = 0; g < TB / 8; ++g) {
        float pv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { pv[i] = ex2_approx(sv[g * 8 + i] - m_safe); rs4[i & 3] += pv[i]; }
        if constexpr (DROP) {
          // dropout acts on the normalised probabilities; the row sum above uses the un-dropped values (HF: softmax -> dropout -> @V)
          float sc[8];
          drop_scale8(p.drop, dstream, dgrp0 + g, sc);
#pragma unroll
          for (int i = 0; i < 8; ++i) pv[i] *= sc[i];
        }
        st_sw128(sP + (g >> 3) * HALF_BYTES, r, g & 7, pack8(pv));
      }
      l_run = l_run * corr + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(p_ready);
      if (threadIdx.x == 0) TS(22 + j * 8);
    }
    // last P V retired -> normalise and store O, store LSE
    mbar_wait(pv_full, (nkv - 1) & 1);
    tc_fence_after();
    if (threadIdx.x == 0) TS(23);
    const float inv_l = l_run > 0.f ? 1.f / l_run : 0.f;
    __nv_bfloat16* orow = p.o + (size_t)(tok0 + (qrow < L ? qrow : 0)) * p.ldo + (size_t)h * TD;
    drain_tile_bf16<TD>(tO + lane_off, inv_l, sQ /* Q is dead after the last S MMA */, &tm_o, p.ocol0 + h * TD, tok0 + q0,
                        q0 + TB <= L, qrow < L, r, orow);
    if (qrow < L)
      p.lse[((size_t)b * p.Hq + h) * L + qrow] = l_run > 0.f ? (m_run + log2f(l_run)) * 0.6931471805599453f : INFINITY;
  }
  if (threadIdx.x == 0) TS(40);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
  if (threadIdx.x == 0) TS(41);
#undef TS
}

// ============================================================================================================
// backward pre-pass: delta[b,h,i] = sum_d dO[i,d] * O[i,d]   (one warp per (token, head))
// ============================================================================================================
template <int TD>
__global__ void attn_tc_delta_kernel(const __nv_bfloat16* __restrict__ o, long long ldo, const __nv_bfloat16* __restrict__ d_o,
                                     long long lddo, const float* __restrict__ lse, float* __restrict__ delta,
                                     float* __restrict__ nl2, int B, int L, int Lp, int Hq) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= B * L * Hq) return;
  const int tok = gw / Hq, h = gw - tok * Hq;
  float acc;
  if constexpr (TD == 128) {
    float a[4], c[4];
    const uint2 ra = *reinterpret_cast<const uint2*>(o + (size_t)tok * ldo + (size_t)h * TD + lane * 4);
    const uint2 rc = *reinterpret_cast<const uint2*>(d_o + (size_t)tok * lddo + (size_t)h * TD + lane * 4);
    float2 t;
    t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ra.x)); a[0] = t.x; a[1] = t.y;
    t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ra.y)); a[2] = t.x; a[3] = t.y;
    t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rc.x)); c[0] = t.x; c[1] = t.y;
    t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rc.y)); c[2] = t.x; c[3] = t.y;
    acc = a[0] * c[0] + a[1] * c[1] + a[2] * c[2] + a[3] * c[3];
  } else {
    const __nv_bfloat162 ra = *reinterpret_cast<const __nv_bfloat162*>(o + (size_t)tok * ldo + (size_t)h * TD + lane * 2);
    const __nv_bfloat162 rc = *reinterpret_cast<const __nv_bfloat162*>(d_o + (size_t)tok * lddo + (size_t)h * TD + lane * 2);
    const float2 a = __bfloat1622float2(ra), c = __bfloat1622float2(rc);
    acc = a.x * c.x + a.y * c.y;
  }
  acc = warp_sum(acc);
  const int b = tok / L, l = tok - b * L;
  const size_t row = ((size_t)b * Hq + h) * Lp;
  if (lane == 0) {
    delta[row + l] = acc;
    nl2[row + l] = -lse[((size_t)b * Hq + h) * L + l] * 1.4426950408889634f;       // lse = +inf (fully masked query) -> -inf
  }
  if (l == L - 1 && L + lane < Lp) {                               // pad entries [L, Lp): queries that do not exist contribute nothing
    for (int lp = L + lane; lp < Lp; lp += 32) { delta[row + lp] = 0.f; nl2[row + lp] = -INFINITY; }
  }
}

// ============================================================================================================
// backward dK, dV: grid (ceil(L/128) key tiles, Hkv, B); thread = key row. Loops over the q heads of the group and the
// query tiles that can see this key tile. dV and dK accumulate in TMEM across ALL of them.
// ============================================================================================================
template <int TD, bool DROP>
__global__ void __launch_bounds__(160, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                       const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                       const __grid_constant__ CUtensorMap tm_dk, const __grid_constant__ CUtensorMap tm_dv,
                       const AttnTcParams p) {
  using C = TcD<TD>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sK = smem;
  unsigned char* sV = sK + C::TILE;
  unsigned char* sQ = sV + C::TILE;
  unsigned char* sdO = sQ + C::TILE;
  unsigned char* sPt = sdO + C::TILE;
  unsigned char* sdSt = sPt + TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdSt + TILE_BYTES);
  uint64_t *kv_full = bars, *qdo_full = bars + 1, *st_full = bars + 2, *pt_ready = bars + 3, *mma2_done = bars + 4;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 8);
  float* sLse = reinterpret_cast<float*>(bars + 10);            // [128] base-2 LSE of the current query tile
  float* sDelta = sLse + 128;                                   // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb = blockIdx.x, hk = blockIdx.y, b = blockIdx.z;
  const int group = p.Hq / p.Hkv;
  const int L = p.L, kv0 = kb * TB, tok0 = b * L;
  const int nq_tiles = (L + TB - 1) / TB;
  const int i_begin = p.causal ? kb : 0;                        // query tiles before the key tile see none of it
  const int n_iter = group * (nq_tiles - i_begin);

  if (threadIdx.x == 128) {
    mbar_init(kv_full, 1); mbar_init(qdo_full, 1); mbar_init(st_full, 1); mbar_init(pt_ready, 128); mbar_init(mma2_done, 1);
    fence_mbar_init();
    prefetch_tmap(&tm_q); prefetch_tmap(&tm_k); prefetch_tmap(&tm_v); prefetch_tmap(&tm_do);
  }
  if (warp == 0) { tmem_alloc(tmem_holder, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;
  const uint32_t tSt = tmem, tdPt = tmem + 128, tdV = tmem + 256, tdK = tmem + 256 + TD;

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * C::TILE);
      load_tile<TD>(sK, &tm_k, kv_full, p.kcol0 + hk * TD, tok0 + kv0);
      load_tile<TD>(sV, &tm_v, kv_full, p.vcol0 + hk * TD, tok0 + kv0);
      constexpr uint32_t idesc_kk = make_idesc_bf16(TB, TB);
      constexpr uint32_t idesc_mn = make_idesc_bf16_bmn(TB, TD);
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ = smem_u32(sQ), aDO = smem_u32(sdO), aPt = smem_u32(sPt), aDSt = smem_u32(sdSt);
      int it = 0;
      for (int hq = hk * group; hq < (hk + 1) * group; ++hq) {
        for (int i = i_begin; i < nq_tiles; ++i, ++it) {
          const uint32_t ph = it & 1;
          mbar_wait(mma2_done, ph ^ 1);                          // previous iteration's dV/dK MMAs have consumed Q/dO
          mbar_arrive_expect_tx(qdo_full, 2 * C::TILE);
          load_tile<TD>(sQ, &tm_q, qdo_full, p.qcol0 + hq * TD, tok0 + i * TB);
          load_tile<TD>(sdO, &tm_do, qdo_full, p.ocol0 + hq * TD, tok0 + i * TB);
          if (it == 0) mbar_wait(kv_full, 0);
          mbar_wait(qdo_full, ph);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < C::KS; ++kk) umma_f16(tSt, kmajor_desc(aK, kk), kmajor_desc(aQ, kk), idesc_kk, kk != 0);    // S^T  = K Q^T
#pragma unroll
          for (int kk = 0; kk < C::KS; ++kk) umma_f16(tdPt, kmajor_desc(aV, kk), kmajor_desc(aDO, kk), idesc_kk, kk != 0);  // dP^T = V dO^T
          umma_commit(st_full);
          mbar_wait(pt_ready, ph);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_f16(tdV, kmajor_desc(aPt, kk), mnmajor_desc(aDO, kk), idesc_mn, (it | kk) != 0);   // dV += P^T dO
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_f16(tdK, kmajor_desc(aDSt, kk), mnmajor_desc(aQ, kk), idesc_mn, (it | kk) != 0);   // dK += dS^T Q
          umma_commit(mma2_done);
        }
      }
    }
  } else {
    const int r = warp * 32 + lane;                              // key row within the tile == TMEM lane
    const int key = kv0 + r;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    bool key_ok = key < L;
    if (key_ok && p.mask) key_ok = p.mask[(size_t)tok0 + key] != 0;
    unsigned long long dstream = 0;
    const int lp8 = (L + 7) >> 3;
    if constexpr (DROP) dstream = drop_stream(p.drop);
    int it = 0;
    for (int hq = hk * group; hq < (hk + 1) * group; ++hq) {
      for (int i = i_begin; i < nq_tiles; ++i, ++it) {
        const uint32_t ph = it & 1;
        const int q0 = i * TB;
        {
          const int qi = q0 + r;
          const size_t idx = ((size_t)b * p.Hq + hq) * L + qi;
          named_bar_sync(1, 128);                                // readers of the previous tile's lse/delta are done
          sLse[r] = qi < L ? p.lse[idx] * 1.4426950408889634f : INFINITY;
          sDelta[r] = qi < L ? p.delta[((size_t)b * p.Hq + hq) * p.Lp + qi] : 0.f;
          named_bar_sync(1, 128);
        }
        mbar_wait(st_full, ph);
        tc_fence_after();
        mbar_wait(mma2_done, ph ^ 1);                            // previous P^T / dS^T tiles consumed by their MMAs
        const bool diag = p.causal && (kv0 + TB - 1 > q0);       // only the diagonal tile needs the causal compare
        const int dropkey = key_ok ? 0 : -1;
#pragma unroll 1
        for (int c = 0; c < TB; c += 32) {
          uint32_t vs[32], vp[32];
          tmem_ld_32x32(tSt + lane_off + c, vs);
          tmem_ld_32x32(tdPt + lane_off + c, vp);
          float lse_c[32], del_c[32];
#pragma unroll
          for (int x = 0; x < 32; x += 4) {                      // per-query statistics: 128-bit broadcast reads
            *reinterpret_cast<float4*>(lse_c + x) = *reinterpret_cast<const float4*>(sLse + c + x);
            *reinterpret_cast<float4*>(del_c + x) = *reinterpret_cast<const float4*>(sDelta + c + x);
          }
          // dropout keep bits of (query q0+c+x, this thread's key), x = 0..31. The mask is generated in groups of 8 KEYS of
          // one query (csrc/attention.cu's indexing), i.e. ACROSS the 8 lanes that share key/8 here: each lane draws the
          // groups of 4 queries (x = (lane & 7) + 8t), keeps only the 8 keep-bits of each draw, and the lanes of an octet
          // exchange them with 8 shuffles - 4 Philox calls per lane per 32 queries instead of 32
          uint32_t keepw = 0xffffffffu;
          if constexpr (DROP) {
            uint32_t mine = 0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int qx = q0 + c + (lane & 7) + 8 * t;
              float sc[8];
              drop_scale8(p.drop, dstream, (((unsigned long long)b * p.Hq + hq) * L + (unsigned long long)qx) * (unsigned long long)lp8 + (unsigned long long)(key >> 3), sc);
#pragma unroll
              for (int j2 = 0; j2 < 8; ++j2) mine |= (sc[j2] != 0.f ? 1u : 0u) << (8 * t + j2);
            }
            keepw = 0;
#pragma unroll
            for (int i2 = 0; i2 < 8; ++i2) {
              const uint32_t v = __shfl_sync(0xffffffffu, mine, (lane & ~7) | i2);     // draws of queries x = i2 + 8t
#pragma unroll
              for (int t = 0; t < 4; ++t) keepw |= ((v >> (8 * t + (lane & 7))) & 1u) << (8 * t + i2);
            }
          }
          tmem_ld_wait();
          float pt[32], ds[32];
          const int qbase = q0 + c;
#pragma unroll
          for (int x = 0; x < 32; ++x) {
            // branch-free masking: dropped (padding key, or key in the query's causal future) -> -inf
            const int drop = dropkey | (diag ? ((qbase + x - key) >> 31) : 0);
            const uint32_t bits = (__float_as_uint(__uint_as_float(vs[x]) * sl2) & ~(uint32_t)drop) | (0xff800000u & (uint32_t)drop);
            const float pr = ex2_approx(__uint_as_float(bits) - lse_c[x]);      // lse = +inf (padding / fully masked query) -> 0
            if constexpr (DROP) {
              const float sc = ((keepw >> x) & 1u) ? p.drop.inv_keep : 0.f;     // P_drop = sc * P ; dP = sc * dP_drop
              pt[x] = pr * sc;
              ds[x] = pr * (__uint_as_float(vp[x]) * sc - del_c[x]) * p.scale;
            } else {
              pt[x] = pr;
              ds[x] = pr * (__uint_as_float(vp[x]) - del_c[x]) * p.scale;
            }
          }
          unsigned char* hp = sPt + (c >> 6) * HALF_BYTES;
          unsigned char* hd = sdSt + (c >> 6) * HALF_BYTES;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            st_sw128(hp, r, ((c & 63) >> 3) + g, pack8(pt + g * 8));
            st_sw128(hd, r, ((c & 63) >> 3) + g, pack8(ds + g * 8));
          }
        }
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(pt_ready);
      }
    }
    // all dV / dK MMAs retired -> drain the accumulators
    mbar_wait(mma2_done, (n_iter - 1) & 1);
    tc_fence_after();
    {
      const bool st_ok = key < L, full = kv0 + TB <= L;
      __nv_bfloat16* dvrow = p.dv + (size_t)(tok0 + (st_ok ? key : 0)) * p.lddv + (size_t)hk * TD;
      __nv_bfloat16* dkrow = p.dk + (size_t)(tok0 + (st_ok ? key : 0)) * p.lddk + (size_t)hk * TD;
      // P^T / dS^T staging tiles are free once the last MMAs have retired
      drain_tile_bf16<TD>(tdV + lane_off, 1.f, sPt, &tm_dv, hk * TD, tok0 + kv0, full, st_ok, r, dvrow);
      drain_tile_bf16<TD>(tdK + lane_off, 1.f, sdSt, &tm_dk, hk * TD, tok0 + kv0, full, st_ok, r, dkrow);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ============================================================================================================
// backward dQ: grid (ceil(L/128) query tiles, Hq, B); thread = query row; dQ accumulates in TMEM over the key tiles
// ============================================================================================================
template <int TD, bool DROP>
__global__ void __launch_bounds__(160, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                      const __grid_constant__ CUtensorMap tm_dq, const AttnTcParams p) {
  using C = TcD<TD>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;
  unsigned char* sdO = sQ + C::TILE;
  unsigned char* sK = sdO + C::TILE;
  unsigned char* sV = sK + C::TILE;
  unsigned char* sdS = sV + C::TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + TILE_BYTES);
  uint64_t *qdo_full = bars, *kv_full = bars + 1, *s_full = bars + 2, *ds_ready = bars + 3, *mma2_done = bars + 4;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.Hq / p.Hkv);
  const int L = p.L, q0 = qb * TB, tok0 = b * L;
  const int nkv = p.causal ? (min(L, q0 + TB) + TB - 1) / TB : (L + TB - 1) / TB;

  if (threadIdx.x == 128) {
    mbar_init(qdo_full, 1); mbar_init(kv_full, 1); mbar_init(s_full, 1); mbar_init(ds_ready, 128); mbar_init(mma2_done, 1);
    fence_mbar_init();
    prefetch_tmap(&tm_q); prefetch_tmap(&tm_k); prefetch_tmap(&tm_v); prefetch_tmap(&tm_do);
  }
  if (warp == 0) { tmem_alloc(tmem_holder, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;
  const uint32_t tS = tmem, tdP = tmem + 128, tdQ = tmem + 256;

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(qdo_full, 2 * C::TILE);
      load_tile<TD>(sQ, &tm_q, qdo_full, p.qcol0 + h * TD, tok0 + q0);
      load_tile<TD>(sdO, &tm_do, qdo_full, p.ocol0 + h * TD, tok0 + q0);
      constexpr uint32_t idesc_kk = make_idesc_bf16(TB, TB);
      constexpr uint32_t idesc_mn = make_idesc_bf16_bmn(TB, TD);
      const uint32_t aQ = smem_u32(sQ), aDO = smem_u32(sdO), aK = smem_u32(sK), aV = smem_u32(sV), aDS = smem_u32(sdS);
      for (int j = 0; j < nkv; ++j) {
        const uint32_t ph = j & 1;
        mbar_wait(mma2_done, ph ^ 1);                            // previous dQ MMAs consumed K (and dS)
        mbar_arrive_expect_tx(kv_full, 2 * C::TILE);
        load_tile<TD>(sK, &tm_k, kv_full, p.kcol0 + hk * TD, tok0 + j * TB);
        load_tile<TD>(sV, &tm_v, kv_full, p.vcol0 + hk * TD, tok0 + j * TB);
        if (j == 0) mbar_wait(qdo_full, 0);
        mbar_wait(kv_full, ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < C::KS; ++kk) umma_f16(tS, kmajor_desc(aQ, kk), kmajor_desc(aK, kk), idesc_kk, kk != 0);      // S  = Q K^T
#pragma unroll
        for (int kk = 0; kk < C::KS; ++kk) umma_f16(tdP, kmajor_desc(aDO, kk), kmajor_desc(aV, kk), idesc_kk, kk != 0);    // dP = dO V^T
        umma_commit(s_full);
        mbar_wait(ds_ready, ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma_f16(tdQ, kmajor_desc(aDS, kk), mnmajor_desc(aK, kk), idesc_mn, (j | kk) != 0);   // dQ += dS K
        umma_commit(mma2_done);
      }
    }
  } else {
    const int r = warp * 32 + lane;
    const int qrow = q0 + r;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    const size_t idx = ((size_t)b * p.Hq + h) * L + qrow;
    const float lse2 = qrow < L ? p.lse[idx] * 1.4426950408889634f : INFINITY;
    const float dl = qrow < L ? p.delta[((size_t)b * p.Hq + h) * p.Lp + qrow] : 0.f;
    unsigned long long dstream = 0, dgrow = 0;
    if constexpr (DROP) {
      dstream = drop_stream(p.drop);
      dgrow = (((unsigned long long)b * p.Hq + h) * L + (unsigned long long)qrow) * (unsigned long long)((L + 7) >> 3);
    }
    for (int j = 0; j < nkv; ++j) {
      const uint32_t ph = j & 1;
      const int kv0 = j * TB;
      uint32_t kbits[4];
      key_bits(p.mask ? p.mask + (size_t)tok0 : nullptr, kv0, L, lane, kbits);
      const bool diag = p.causal && (kv0 + TB - 1 > q0);
      const int dcol = diag ? (qrow - kv0) : 0x7fffffff;           // columns > dcol are in the causal future
      mbar_wait(s_full, ph);
      tc_fence_after();
      mbar_wait(mma2_done, ph ^ 1);                              // previous dS tile consumed
#pragma unroll 1
      for (int c = 0; c < TB; c += 32) {
        uint32_t vs[32], vp[32];
        tmem_ld_32x32(tS + lane_off + c, vs);
        tmem_ld_32x32(tdP + lane_off + c, vp);
        float dsc[32];
        if constexpr (DROP) {
#pragma unroll
          for (int g = 0; g < 4; ++g) drop_scale8(p.drop, dstream, dgrow + (unsigned long long)(((kv0 + c) >> 3) + g), dsc + g * 8);
        }
        tmem_ld_wait();
        const uint32_t kw = kbits[c >> 5];
        float ds[32];
#pragma unroll
        for (int x = 0; x < 32; ++x) {
          const int drop = ((dcol - (c + x)) >> 31) | (int)(((kw >> x) & 1u) - 1u);
          const uint32_t bits = (__float_as_uint(__uint_as_float(vs[x]) * sl2) & ~(uint32_t)drop) | (0xff800000u & (uint32_t)drop);
          const float pr = ex2_approx(__uint_as_float(bits) - lse2);
          if constexpr (DROP) ds[x] = pr * (__uint_as_float(vp[x]) * dsc[x] - dl) * p.scale;
          else                ds[x] = pr * (__uint_as_float(vp[x]) - dl) * p.scale;
        }
        unsigned char* hd = sdS + (c >> 6) * HALF_BYTES;
#pragma unroll
        for (int g = 0; g < 4; ++g) st_sw128(hd, r, ((c & 63) >> 3) + g, pack8(ds + g * 8));
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(ds_ready);
    }
    mbar_wait(mma2_done, (nkv - 1) & 1);
    tc_fence_after();
    __nv_bfloat16* dqrow = p.dq + (size_t)(tok0 + (qrow < L ? qrow : 0)) * p.lddq + (size_t)h * TD;
    drain_tile_bf16<TD>(tdQ + lane_off, 1.f, sdS /* free after the last dQ MMA */, &tm_dq, h * TD, tok0 + q0, q0 + TB <= L,
                        qrow < L, r, dqrow);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ============================================================================================================
// PIPELINED, PERSISTENT backward (round 2). The two kernels above walk one hand-off chain per CTA (TMA -> MMA -> math ->
// MMA, 1 CTA/SM because of TMEM): profiles/r01_attention_tc_ncu_details.csv shows issue slots 12-23 % busy. Here the
// streamed operand is cut into 64-row HALF steps so that everything that repeats is multi-buffered inside the 512 TMEM
// columns and 227 KB of shared memory, and four roles run concurrently in one persistent CTA per SM:
//     warp 0      producer : TMA loads of the resident tiles (per item) and of the streamed halves (3-4 stage ring); in dKdV
//                            its 32 lanes also stage the halves' per-query statistics (-lse, delta) into the ring
//     warp 1      S/dP MMAs: runs up to two halves ahead of the math (a TMEM buffer is re-armed as soon as the math warps have copied
//                            it into registers), straight across item boundaries
//     warp 10     accumulating MMAs: its own issuing thread, so neither stream ever waits behind the other's dependencies
//     warps 2-5   math WG0 : columns  0-31    } of EVERY half; thread = TMEM lane (key row in dKdV, query row in dQ);
//     warps 6-9   math WG1 : columns 32-63    } P / dS written as bf16 into 128B-swizzled tiles that feed the second MMAs
// Items (dKdV: key tile x kv head x batch; dQ: query tile x head x batch) are strided over the persistent grid; the next
// item's resident tiles are fetched while the current item's accumulators drain (WG0 / WG1 drain one accumulator each).
// v1 -> v2 (profiles/r02_attn_bwd_pipe_v1_timeline_*.txt: 17 000 cycles per 4-half item): per-half global loads of lse /
// delta sat exposed in the math warps (~2 500 cycles) -> staged by the producer; 1 300 instructions per thread-half (two warps
// per scheduler => 2 600 cycles) -> mask-free fast path off the diagonal, softmax scale applied once in the drain; the drain
// waited for the TMA store's global completion -> only for its shared-memory read; one item's first S/dP waited for the
// previous item's last accumulate -> continuous S/dP stream; head_dim 64 double-buffers the resident tiles.
// ============================================================================================================
constexpr int CH64 = 64 * 64 * 2;            // one [64 rows x 64 cols] bf16 box = 8 KB (a 64-row tile is D/64 of these)
__device__ __forceinline__ uint64_t kmajor_desc_r64(uint32_t tile, int kk) {      // 64-row tile, K-major, k-step kk over d
  return make_sw128_kmajor_desc(tile + (kk >> 2) * CH64) + (uint64_t)(2 * (kk & 3));
}
__device__ __forceinline__ uint64_t mnmajor_desc_r64(uint32_t tile, int kk) {     // 64-row tile as MN-major B, k-step kk over its rows
  return make_sw128_mnmajor_desc(tile + kk * 16 * 128, CH64, 1024);
}
template <int D>
__device__ __forceinline__ void load_tile64(unsigned char* dst, const CUtensorMap* tm64, uint64_t* bar, int col, int row) {
#pragma unroll
  for (int hh = 0; hh < D / 64; ++hh) tma_load_2d(dst + hh * CH64, tm64, bar, col + hh * 64, row);
}
// drain one [128 rows x NC cols] fp32 TMEM accumulator (NC = 64 or 128), times `scale`, to bf16 HBM: one math warpgroup (128
// threads, named barrier `bar_id`), swizzled staging + TMA stores for full tiles (the issuer waits only until the store has
// READ the staging tile), predicated row stores for ragged ones
template <int NC>
__device__ __forceinline__ void drain_acc_wg(uint32_t tacc, float scale, unsigned char* stage, const CUtensorMap* tm, int col0,
                                             int row0_global, bool full_tile, bool row_ok, int r, __nv_bfloat16* fallback_row,
                                             int bar_id, bool issuer) {
#pragma unroll 1
  for (int c = 0; c < NC; c += 32) {
    uint32_t v[32]; float f[32];
    tmem_ld_32x32(tacc + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) * scale;
    if (full_tile) {
#pragma unroll
      for (int g = 0; g < 4; ++g) st_sw128(stage + (c >> 6) * HALF_BYTES, r, ((c & 63) >> 3) + g, pack8(f + g * 8));
    } else if (row_ok) {
#pragma unroll
      for (int g = 0; g < 4; ++g) *reinterpret_cast<bf16x8*>(fallback_row + c + g * 8) = pack8(f + g * 8);
    }
  }
  if (full_tile) {
    fence_proxy_async();
    named_bar_sync(bar_id, 128);
    if (issuer) {
#pragma unroll
      for (int hh = 0; hh < NC / 64; ++hh) tma_store_2d(tm, stage + hh * HALF_BYTES, col0 + hh * 64, row0_global);
      bulk_commit();
      bulk_wait_read<0>();
    }
  }
}

struct alignas(128) PipeBars {                 // shared-memory barrier block of the pipelined kernels
  uint64_t res_full[2], res_free[2];           // resident tiles of an item (K,V / Q,dO): loaded / no longer read by S / dP MMAs
  uint64_t str_full[4], str_free[4];           // streamed halves landed (stage = n % NS) / consumed by the accumulating MMAs
  uint64_t aux_full[4];                        // dKdV: -lse / delta of the half's 64 queries staged next to the stage (32 arrivals)
  uint64_t s_full[2];                          // S / dP of half n complete in TMEM (buffer = n & 1)
  uint64_t s_free[2];                          // every math thread has pulled its S / dP columns into registers: the buffer can take half n+2
  uint64_t p_ready[2];                         // math wrote the bf16 tiles of half n (and has finished reading S / dP)
  uint64_t mma2_done[2];                       // accumulating MMAs of half n retired: bf16 tiles reusable
  uint64_t acc_full, acc_free;                 // item's accumulators complete / drained
  uint32_t tmem_holder, pad;
};
constexpr int kMaxKeyWords = 128;             // dQ kernel: key-validity bits staged per item for rows up to 4096 keys
constexpr bool kShareHalf = true;             // true: both math warpgroups split every half's columns; false: they alternate halves
constexpr int kPipeThreads = 352;             // warp 0 producer, warp 1 S/dP issuer, warps 2-9 math, warp 10 accumulate issuer
template <int D> constexpr int pipe_stages() { return D == 128 ? 3 : 4; }   // depth of the streamed ring
template <int D> constexpr int pipe_res() { return D == 128 ? 1 : 2; }      // resident-tile buffers (head_dim 64 has the room)

__device__ __forceinline__ void pipe_bars_init(PipeBars* bars) {
  for (int i = 0; i < 2; ++i) {
    mbar_init(&bars->res_full[i], 1); mbar_init(&bars->res_free[i], 1);
    mbar_init(&bars->s_full[i], 1); mbar_init(&bars->s_free[i], kShareHalf ? 256 : 128); mbar_init(&bars->p_ready[i], kShareHalf ? 256 : 128);
    mbar_init(&bars->mma2_done[i], 1);
  }
  for (int i = 0; i < 4; ++i) { mbar_init(&bars->str_full[i], 1); mbar_init(&bars->str_free[i], 1); mbar_init(&bars->aux_full[i], 32); }
  mbar_init(&bars->acc_full, 1); mbar_init(&bars->acc_free, 256);
  fence_mbar_init();
}

// ---- dK, dV ------------------------------------------------------------------------------------------------------
template <int TD, bool DROP>
__global__ void __launch_bounds__(kPipeThreads, 1)
attn_bwd_dkv_pipe_kernel(const __grid_constant__ CUtensorMap tm_q64, const __grid_constant__ CUtensorMap tm_k,
                         const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do64,
                         const __grid_constant__ CUtensorMap tm_dk, const __grid_constant__ CUtensorMap tm_dv,
                         const AttnTcParams p) {
  using C = TcD<TD>;
  constexpr int HT = (TD / 64) * CH64;                           // bytes of a [64 rows x TD] streamed half
  constexpr int NS = pipe_stages<TD>(), RES = pipe_res<TD>();
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sK = smem;                                      // [RES][TILE]
  unsigned char* sV = sK + RES * C::TILE;                        // [RES][TILE]
  unsigned char* sQh = sV + RES * C::TILE;                       // [NS][HT]
  unsigned char* sdOh = sQh + NS * HT;                           // [NS][HT]
  unsigned char* sPT = sdOh + NS * HT;                           // [2][16 KB]  P^T  [128 keys x 64 queries]
  unsigned char* sdST = sPT + 2 * HALF_BYTES;                    // [2][16 KB]  dS^T
  PipeBars* bars = reinterpret_cast<PipeBars*>(sdST + 2 * HALF_BYTES);
  float* sNl = reinterpret_cast<float*>(bars + 1);              // [NS][64]  -lse * log2(e) of the half's queries (-inf beyond L)
  float* sDl = sNl + NS * 64;                                    // [NS][64]  delta

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L, group = p.Hq / p.Hkv;
  const int ntiles = (L + TB - 1) / TB;
  const int n_items = ntiles * p.Hkv * p.B;
  const int nqh = (L + 63) >> 6;                                 // 64-query halves of one sequence
  const bool dbg = p.dbg != nullptr && blockIdx.x == 0;         // tuning aid: clock64 timeline of CTA 0's first halves
#define PTS(slot) do { if (dbg && (slot) < 64) p.dbg[slot] = clock64(); } while (0)

  if (threadIdx.x == 0) {
    pipe_bars_init(bars);
    prefetch_tmap(&tm_q64); prefetch_tmap(&tm_k); prefetch_tmap(&tm_v); prefetch_tmap(&tm_do64);
  }
  if (warp == 1) { tmem_alloc(&bars->tmem_holder, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_holder;
  const uint32_t tdV = tmem + 256, tdK = tmem + 256 + TD;        // S^T / dP^T buffers: tmem + t*128 (+64)

  if (warp == 0) {
    // ================= producer (lane 0: TMA; all lanes: cp.async of the halves' query statistics into the stage) =================
    int n = 0, c = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++c) {
      const int kb = item % ntiles, hk = (item / ntiles) % p.Hkv, b = item / (ntiles * p.Hkv);
      const int tok0 = b * L, kv0 = kb * TB, rb = c % RES;
      if (lane == 0) {
        mbar_wait(&bars->res_free[rb], ((c / RES) & 1) ^ 1);     // the item that used this buffer no longer reads K, V
        mbar_arrive_expect_tx(&bars->res_full[rb], 2 * C::TILE);
        load_tile<TD>(sK + rb * C::TILE, &tm_k, &bars->res_full[rb], p.kcol0 + hk * TD, tok0 + kv0);
        load_tile<TD>(sV + rb * C::TILE, &tm_v, &bars->res_full[rb], p.vcol0 + hk * TD, tok0 + kv0);
      }
      const int h_begin = p.causal ? 2 * kb : 0;                 // query halves before the key tile see none of it
      for (int hq = hk * group; hq < (hk + 1) * group; ++hq) {
        for (int qh = h_begin; qh < nqh; ++qh, ++n) {
          const int s = n % NS, u = n / NS;
          if (lane == 0) {
            mbar_wait(&bars->str_free[s], (u & 1) ^ 1);          // stage s consumed by the accumulating MMAs of half n-NS
            if (n < 8) PTS(56 + n);
            mbar_arrive_expect_tx(&bars->str_full[s], 2 * HT);
            load_tile64<TD>(sQh + s * HT, &tm_q64, &bars->str_full[s], p.qcol0 + hq * TD, tok0 + qh * 64);
            load_tile64<TD>(sdOh + s * HT, &tm_do64, &bars->str_full[s], p.ocol0 + hq * TD, tok0 + qh * 64);
          }
          __syncwarp();                                          // stage s is free for everybody
          {
            // -lse*log2e and delta of the half's 64 queries: 8 bytes per lane and array, global -> shared asynchronously (the
            // padded [B,Hq,Lp] layout written by the delta kernel makes every half complete); each lane's cp.async group
            // arrives on the stage's aux barrier when it has landed - the producer never waits for a load
            const size_t base = ((size_t)b * p.Hq + hq) * p.Lp + (size_t)qh * 64 + lane * 2;
            const uint32_t d0 = smem_u32(sNl + s * 64 + lane * 2), d1 = smem_u32(sDl + s * 64 + lane * 2);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d0), "l"(p.nl2 + base) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d1), "l"(p.delta + base) : "memory");
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars->aux_full[s])) : "memory");
          }
        }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(TB, 64);          // S^T / dP^T half: [128 keys x 64 queries]
      auto nh_of = [&](int item) { return group * (nqh - (p.causal ? 2 * (item % ntiles) : 0)); };
      // the S/dP stream: one half ahead of the accumulating stream, straight across item boundaries
      int sp_item = blockIdx.x, sp_c = 0, sp_k = 0, sp_n = 0;
      auto sp_issue = [&]() {
        const int rb = sp_c % RES;
        if (sp_k == 0) mbar_wait(&bars->res_full[rb], (sp_c / RES) & 1);
        const int st = sp_n % NS, t = sp_n & 1;
        mbar_wait(&bars->str_full[st], (sp_n / NS) & 1);
        if (sp_n < 8) PTS(3 * sp_n);
        tc_fence_after();
        const uint32_t aK = smem_u32(sK + rb * C::TILE), aV = smem_u32(sV + rb * C::TILE);
        const uint32_t tS = tmem + (uint32_t)(t * 128), aQ = smem_u32(sQh + st * HT), aDO = smem_u32(sdOh + st * HT);
#pragma unroll
        for (int kk = 0; kk < C::KS; ++kk) umma_f16(tS, kmajor_desc(aK, kk), kmajor_desc_r64(aQ, kk), idesc_s, kk != 0);        // S^T  = K Q^T
#pragma unroll
        for (int kk = 0; kk < C::KS; ++kk) umma_f16(tS + 64, kmajor_desc(aV, kk), kmajor_desc_r64(aDO, kk), idesc_s, kk != 0);  // dP^T = V dO^T
        umma_commit(&bars->s_full[t]);
        ++sp_n;
        if (++sp_k == nh_of(sp_item)) {                          // every S / dP MMA of that item has been issued
          umma_commit(&bars->res_free[rb]);
          sp_item += gridDim.x; ++sp_c; sp_k = 0;
        }
      };
      // S/dP stream: gated only by its own inputs (resident tiles, streamed stage, a TMEM buffer the math warps have emptied)
      while (sp_item < n_items) {
        if (sp_n >= 2) {
          mbar_wait(&bars->s_free[sp_n & 1], ((sp_n >> 1) - 1) & 1);      // half sp_n-2 has been copied out of this buffer
          tc_fence_after();
        }
        sp_issue();
      }
    }
  } else if (warp == 10) {
    // ================= accumulate issuer: dV += P^T dO, dK += dS^T Q as soon as the math warps publish a half =================
    if (lane == 0) {
      constexpr uint32_t idesc_acc = make_idesc_bf16_bmn(TB, TD);    // B = dO / Q half, MN-major
      int n0 = 0, c = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++c) {
        const int NH = group * (nqh - (p.causal ? 2 * (item % ntiles) : 0));
        for (int k = 0; k < NH; ++k) {
          const int n = n0 + k, s = n & 1, u = n >> 1;
          mbar_wait(&bars->p_ready[s], u & 1);
          if (n < 8) PTS(3 * n + 1);
          if (k == 0) mbar_wait(&bars->acc_free, (c & 1) ^ 1);   // previous item's accumulators drained
          if (n < 8) PTS(3 * n + 2);
          tc_fence_after();
          const int st = n % NS;
          const uint32_t aPT = smem_u32(sPT + s * HALF_BYTES), aDST = smem_u32(sdST + s * HALF_BYTES);
          const uint32_t aQ = smem_u32(sQh + st * HT), aDO = smem_u32(sdOh + st * HT);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_f16(tdV, kmajor_desc(aPT, kk), mnmajor_desc_r64(aDO, kk), idesc_acc, (k | kk) != 0);   // dV += P^T dO
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_f16(tdK, kmajor_desc(aDST, kk), mnmajor_desc_r64(aQ, kk), idesc_acc, (k | kk) != 0);   // dK += dS^T Q
          umma_commit(&bars->mma2_done[s]);
          umma_commit(&bars->str_free[st]);
        }
        umma_commit(&bars->acc_full);
        n0 += NH;
      }
    }
  } else if (warp < 10) {
    // ================= math warpgroups =================
    const int g = (warp - 2) >> 2;                               // warpgroup 0 / 1 <-> buffer n & 1
    const int r = (warp & 3) * 32 + lane;                        // TMEM lane == key row of the tile (lane quarter = warp % 4)
    const int wt = threadIdx.x - 64 - g * 128;                   // 0..127 inside the warpgroup
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    const int lp8 = (L + 7) >> 3;
    unsigned long long dstream = 0;
    if constexpr (DROP) dstream = drop_stream(p.drop);
    // this row's key validity for the NEXT item is fetched while the current item is processed (a per-item global load in
    // front of the first half was ~2 000 exposed cycles per item: profiles/r02_attn_bwd_pipe_v1_timeline_llama.txt)
    auto key_valid = [&](int item) {
      const int kb = item % ntiles, b = item / (ntiles * p.Hkv);
      const int key = kb * TB + r;
      bool ok = key < L;
      if (ok && p.mask) ok = __ldg(p.mask + (size_t)b * L + key) != 0;
      return ok;
    };
    bool key_ok_next = (int)blockIdx.x < n_items ? key_valid(blockIdx.x) : false;
    int n0 = 0, c = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++c) {
      const int kb = item % ntiles, hk = (item / ntiles) % p.Hkv, b = item / (ntiles * p.Hkv);
      const int tok0 = b * L, kv0 = kb * TB;
      const int key = kv0 + r;
      const bool key_ok = key_ok_next;
      if (item + (int)gridDim.x < n_items) key_ok_next = key_valid(item + gridDim.x);
      const int dropkey = key_ok ? 0 : -1;
      const bool warp_keys_ok = __all_sync(0xffffffffu, key_ok);  // warp-uniform: the mask-free path needs all 32 key rows real
      const int h_begin = p.causal ? 2 * kb : 0;
      const int per_head = nqh - h_begin;
      const int NH = group * per_head;
      for (int k = 0; k < NH; ++k) {
        // EVERY half is shared by both warpgroups (WG g takes query columns [32g, 32g + 32)): two warps per scheduler work on the
        // same half, so its math latency - which the S/dP -> math -> accumulate chain is made of - halves
        const int n = n0 + k, t = n & 1;
        if (!kShareHalf && t != g) continue;
        const int u = n >> 1, st = n % NS;
        const int hq = hk * group + k / per_head, qh = h_begin + k % per_head;
        const int qbase = qh * 64;
        mbar_wait(&bars->aux_full[st], (n / NS) & 1);            // -lse / delta of the half's queries are in the stage
        mbar_wait(&bars->s_full[t], u & 1);
        if (wt == 0 && g == 0 && n < 8) PTS(24 + 2 * n);
        tc_fence_after();
        mbar_wait(&bars->mma2_done[t], (u & 1) ^ 1);             // P^T / dS^T tiles of half n-2 consumed by their MMAs
        const bool diag = p.causal && (kv0 + TB - 1 > qbase);    // only halves on / below the diagonal band need the compare
        const bool fast = warp_keys_ok && !diag;
        const uint32_t tS = tmem + (uint32_t)(t * 128);
        unsigned char* hp = sPT + t * HALF_BYTES;
        unsigned char* hd = sdST + t * HALF_BYTES;
        const float* nlq = sNl + st * 64;
        const float* dlq = sDl + st * 64;
#pragma unroll 1
        for (int cc = kShareHalf ? g * 32 : 0; cc < (kShareHalf ? g * 32 + 32 : 64); cc += 32) {
          uint32_t vs[32], vp[32];
          tmem_ld_32x32(tS + lane_off + cc, vs);
          tmem_ld_32x32(tS + 64 + lane_off + cc, vp);
          float nl_c[32], del_c[32];
#pragma unroll
          for (int x = 0; x < 32; x += 4) {                      // per-query statistics: 128-bit broadcast reads
            *reinterpret_cast<float4*>(nl_c + x) = *reinterpret_cast<const float4*>(nlq + cc + x);
            *reinterpret_cast<float4*>(del_c + x) = *reinterpret_cast<const float4*>(dlq + cc + x);
          }
          uint32_t keepw = 0xffffffffu;
          if constexpr (DROP) {                                  // keep bits of (query qbase+cc+x, this key): see attn_bwd_dkv_tc_kernel
            uint32_t mine = 0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int qx = qbase + cc + (lane & 7) + 8 * t;
              float sc[8];
              drop_scale8(p.drop, dstream, (((unsigned long long)b * p.Hq + hq) * L + (unsigned long long)qx) * (unsigned long long)lp8 + (unsigned long long)(key >> 3), sc);
#pragma unroll
              for (int j2 = 0; j2 < 8; ++j2) mine |= (sc[j2] != 0.f ? 1u : 0u) << (8 * t + j2);
            }
            keepw = 0;
#pragma unroll
            for (int i2 = 0; i2 < 8; ++i2) {
              const uint32_t v = __shfl_sync(0xffffffffu, mine, (lane & ~7) | i2);
#pragma unroll
              for (int t = 0; t < 4; ++t) keepw |= ((v >> (8 * t + (lane & 7))) & 1u) << (8 * t + i2);
            }
          }
          tmem_ld_wait();
          if (kShareHalf || cc == 32) {
            tc_fence_before();
            mbar_arrive(&bars->s_free[t]);                       // S^T / dP^T columns are in registers: the TMEM buffer may be overwritten
          }
          float pt[32], ds[32];
          if (fast) {                                            // no padding key in this warp, no causal boundary in this half
#pragma unroll
            for (int x = 0; x < 32; ++x) {
              const float pr = ex2_approx(fmaf(__uint_as_float(vs[x]), sl2, nl_c[x]));     // -lse = -inf (query beyond L / fully masked) -> 0
              if constexpr (DROP) {
                const float sc = ((keepw >> x) & 1u) ? p.drop.inv_keep : 0.f;              // P_drop = sc * P ; dP = sc * dP_drop
                pt[x] = pr * sc;
                ds[x] = pr * fmaf(__uint_as_float(vp[x]), sc, -del_c[x]);
              } else {
                pt[x] = pr;
                ds[x] = pr * (__uint_as_float(vp[x]) - del_c[x]);
              }
            }
          } else {
#pragma unroll
            for (int x = 0; x < 32; ++x) {
              // branch-free masking: dropped (padding key, or key in the query's causal future) -> -inf
              const int drop = dropkey | (diag ? ((qbase + cc + x - key) >> 31) : 0);
              const uint32_t bits = (__float_as_uint(__uint_as_float(vs[x]) * sl2) & ~(uint32_t)drop) | (0xff800000u & (uint32_t)drop);
              const float pr = ex2_approx(__uint_as_float(bits) + nl_c[x]);
              if constexpr (DROP) {
                const float sc = ((keepw >> x) & 1u) ? p.drop.inv_keep : 0.f;
                pt[x] = pr * sc;
                ds[x] = pr * fmaf(__uint_as_float(vp[x]), sc, -del_c[x]);
              } else {
                pt[x] = pr;
                ds[x] = pr * (__uint_as_float(vp[x]) - del_c[x]);
              }
            }
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            st_sw128(hp, r, (cc >> 3) + q4, pack8(pt + q4 * 8));
            st_sw128(hd, r, (cc >> 3) + q4, pack8(ds + q4 * 8));
          }
        }
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(&bars->p_ready[t]);
        if (wt == 0 && g == 0 && n < 8) PTS(24 + 2 * n + 1);
      }
      // ---- item end: all accumulating MMAs retired -> WG0 drains dV, WG1 drains dK (dS was left unscaled: dK *= scale here) ----
      mbar_wait(&bars->acc_full, c & 1);
      if (wt == 0 && g == 0 && c < 4) PTS(40 + 2 * c);
      tc_fence_after();
      {
        const bool st_ok = key < L, full = kv0 + TB <= L;
        const size_t rowoff = (size_t)(tok0 + (st_ok ? key : 0));
        if (g == 0) drain_acc_wg<TD>(tdV + lane_off, 1.f, sPT, &tm_dv, hk * TD, tok0 + kv0, full, st_ok, r,
                                     p.dv + rowoff * p.lddv + (size_t)hk * TD, 1, wt == 0);
        else        drain_acc_wg<TD>(tdK + lane_off, p.scale, sdST, &tm_dk, hk * TD, tok0 + kv0, full, st_ok, r,
                                     p.dk + rowoff * p.lddk + (size_t)hk * TD, 2, wt == 0);
      }
      tc_fence_before();
      mbar_arrive(&bars->acc_free);
      named_bar_sync(3, 256);                                    // both staging areas are free before the next item's halves write them
      if (wt == 0 && g == 0 && c < 4) PTS(40 + 2 * c + 1);
      n0 += NH;
    }
    if (wt == 0) bulk_wait<0>();                                 // the last TMA stores have left before the CTA's shared memory goes away
  }
#undef PTS
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---- dQ ----------------------------------------------------------------------------------------------------------
template <int TD, bool DROP>
__global__ void __launch_bounds__(kPipeThreads, 1)
attn_bwd_dq_pipe_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k64,
                        const __grid_constant__ CUtensorMap tm_v64, const __grid_constant__ CUtensorMap tm_do,
                        const __grid_constant__ CUtensorMap tm_dq, const AttnTcParams p) {
  using C = TcD<TD>;
  constexpr int HT = (TD / 64) * CH64;
  constexpr int NS = pipe_stages<TD>(), RES = pipe_res<TD>();
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;                                      // [RES][TILE]
  unsigned char* sdO = sQ + RES * C::TILE;                       // [RES][TILE]
  unsigned char* sKh = sdO + RES * C::TILE;                      // [NS][HT]
  unsigned char* sVh = sKh + NS * HT;                            // [NS][HT]
  unsigned char* sdS = sVh + NS * HT;                            // [2][16 KB]  dS [128 queries x 64 keys]
  PipeBars* bars = reinterpret_cast<PipeBars*>(sdS + 2 * HALF_BYTES);
  uint32_t* sKeyBits = reinterpret_cast<uint32_t*>(bars + 1);   // [2][kMaxKeyWords] validity bits of the item's key row (word w: keys 32w..)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L, group = p.Hq / p.Hkv;
  const int ntiles = (L + TB - 1) / TB;
  const int n_items = ntiles * p.Hq * p.B;

  if (threadIdx.x == 0) {
    pipe_bars_init(bars);
    prefetch_tmap(&tm_q); prefetch_tmap(&tm_k64); prefetch_tmap(&tm_v64); prefetch_tmap(&tm_do);
  }
  if (warp == 1) { tmem_alloc(&bars->tmem_holder, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_holder;
  const uint32_t tdQ = tmem + 256;
  // number of 64-key halves query tile qb attends to
  auto n_halves = [&](int qb) { const int kend = p.causal ? min(L, qb * TB + TB) : L; return (kend + 63) >> 6; };

  if (warp == 0) {
    if (lane == 0) {
      int n = 0, c = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++c) {
        const int qb = item % ntiles, h = (item / ntiles) % p.Hq, b = item / (ntiles * p.Hq);
        const int hk = h / group, tok0 = b * L, q0 = qb * TB, rb = c % RES;
        mbar_wait(&bars->res_free[rb], ((c / RES) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->res_full[rb], 2 * C::TILE);
        load_tile<TD>(sQ + rb * C::TILE, &tm_q, &bars->res_full[rb], p.qcol0 + h * TD, tok0 + q0);
        load_tile<TD>(sdO + rb * C::TILE, &tm_do, &bars->res_full[rb], p.ocol0 + h * TD, tok0 + q0);
        const int NH = n_halves(qb);
        for (int j = 0; j < NH; ++j, ++n) {
          const int s = n % NS, u = n / NS;
          mbar_wait(&bars->str_free[s], (u & 1) ^ 1);
          mbar_arrive_expect_tx(&bars->str_full[s], 2 * HT);
          load_tile64<TD>(sKh + s * HT, &tm_k64, &bars->str_full[s], p.kcol0 + hk * TD, tok0 + j * 64);
          load_tile64<TD>(sVh + s * HT, &tm_v64, &bars->str_full[s], p.vcol0 + hk * TD, tok0 + j * 64);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(TB, 64);
      int sp_item = blockIdx.x, sp_c = 0, sp_k = 0, sp_n = 0;
      auto sp_issue = [&]() {
        const int rb = sp_c % RES;
        if (sp_k == 0) mbar_wait(&bars->res_full[rb], (sp_c / RES) & 1);
        const int st = sp_n % NS, t = sp_n & 1;
        mbar_wait(&bars->str_full[st], (sp_n / NS) & 1);
        tc_fence_after();
        const uint32_t aQ = smem_u32(sQ + rb * C::TILE), aDO = smem_u32(sdO + rb * C::TILE);
        const uint32_t tS = tmem + (uint32_t)(t * 128), aK = smem_u32(sKh + st * HT), aV = smem_u32(sVh + st * HT);
#pragma unroll
        for (int kk = 0; kk < C::KS; ++kk) umma_f16(tS, kmajor_desc(aQ, kk), kmajor_desc_r64(aK, kk), idesc_s, kk != 0);        // S  = Q K^T
#pragma unroll
        for (int kk = 0; kk < C::KS; ++kk) umma_f16(tS + 64, kmajor_desc(aDO, kk), kmajor_desc_r64(aV, kk), idesc_s, kk != 0);  // dP = dO V^T
        umma_commit(&bars->s_full[t]);
        ++sp_n;
        if (++sp_k == n_halves(sp_item % ntiles)) {
          umma_commit(&bars->res_free[rb]);
          sp_item += gridDim.x; ++sp_c; sp_k = 0;
        }
      };
      while (sp_item < n_items) {
        if (sp_n >= 2) {
          mbar_wait(&bars->s_free[sp_n & 1], ((sp_n >> 1) - 1) & 1);
          tc_fence_after();
        }
        sp_issue();
      }
    }
  } else if (warp == 10) {
    if (lane == 0) {
      constexpr uint32_t idesc_acc = make_idesc_bf16_bmn(TB, TD);
      int n0 = 0, c = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++c) {
        const int NH = n_halves(item % ntiles);
        for (int k = 0; k < NH; ++k) {
          const int n = n0 + k, s = n & 1, u = n >> 1;
          mbar_wait(&bars->p_ready[s], u & 1);
          if (k == 0) mbar_wait(&bars->acc_free, (c & 1) ^ 1);
          tc_fence_after();
          const int st = n % NS;
          const uint32_t aDS = smem_u32(sdS + s * HALF_BYTES), aK = smem_u32(sKh + st * HT);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_f16(tdQ, kmajor_desc(aDS, kk), mnmajor_desc_r64(aK, kk), idesc_acc, (k | kk) != 0);   // dQ += dS K
          umma_commit(&bars->mma2_done[s]);
          umma_commit(&bars->str_free[st]);
        }
        umma_commit(&bars->acc_full);
        n0 += NH;
      }
    }
  } else if (warp < 10) {
    const int g = (warp - 2) >> 2;
    const int r = (warp & 3) * 32 + lane;                        // query row of the tile == TMEM lane
    const int wt = threadIdx.x - 64 - g * 128;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    unsigned long long dstream = 0;
    if constexpr (DROP) dstream = drop_stream(p.drop);
    // this row's statistics of the NEXT item are fetched while the current item is processed
    auto stats = [&](int item, float& nl, float& dl) {
      const int qb = item % ntiles, h = (item / ntiles) % p.Hq, b = item / (ntiles * p.Hq);
      const int qrow = qb * TB + r;
      const size_t idx = ((size_t)b * p.Hq + h) * p.Lp + qrow;     // padded rows: no bounds test for rows of a ragged tile...
      nl = qrow < p.Lp ? p.nl2[idx] : -INFINITY;                   // ...inside the pad; beyond it (Lp % 128 == 64) nothing exists
      dl = qrow < p.Lp ? p.delta[idx] : 0.f;
    };
    // key-validity bits of an item's batch row -> sKeyBits[buf]: warp w of the 8 math warps ballots words w, w+8, ...
    const int mw = warp - 2;
    auto key_bits_row = [&](int item, int buf) {
      const int b = item / (ntiles * p.Hq);
      const int64_t* mrow = p.mask ? p.mask + (size_t)b * L : nullptr;
      const int nwords = ((L + 63) >> 6) * 2;                    // two words per 64-key half, the last half may be partly beyond L
      for (int w = mw; w < nwords && w < kMaxKeyWords; w += 8) {
        const int key = w * 32 + lane;
        bool keep = key < L;
        if (keep && mrow) keep = __ldg(mrow + key) != 0;
        const uint32_t bits = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) sKeyBits[buf * kMaxKeyWords + w] = bits;
      }
    };
    float nl_next = 0.f, dl_next = 0.f;
    if ((int)blockIdx.x < n_items) { stats(blockIdx.x, nl_next, dl_next); key_bits_row(blockIdx.x, 0); }
    named_bar_sync(3, 256);
    int n0 = 0, c = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++c) {
      const int qb = item % ntiles, h = (item / ntiles) % p.Hq, b = item / (ntiles * p.Hq);
      const int tok0 = b * L, q0 = qb * TB, qrow = q0 + r;
      const float nl = nl_next, dl = dl_next;
      if (item + (int)gridDim.x < n_items) {                     // next item's statistics and key bits: in flight during this item
        stats(item + gridDim.x, nl_next, dl_next);
        key_bits_row(item + gridDim.x, (c + 1) & 1);
      }
      const uint32_t* kbrow = sKeyBits + (c & 1) * kMaxKeyWords;
      unsigned long long dgrow = 0;
      if constexpr (DROP) dgrow = (((unsigned long long)b * p.Hq + h) * L + (unsigned long long)qrow) * (unsigned long long)((L + 7) >> 3);
      const int NH = n_halves(qb);
      for (int k = 0; k < NH; ++k) {
        const int n = n0 + k, t = n & 1;
        if (!kShareHalf && t != g) continue;
        const int u = n >> 1, kv0 = k * 64;
        uint32_t kbits[2];
        if (2 * k + 1 < kMaxKeyWords) { kbits[0] = kbrow[2 * k]; kbits[1] = kbrow[2 * k + 1]; }
        else {                                                   // rows longer than the staged bits (L > 32 * kMaxKeyWords): direct
          const int64_t* mrow = p.mask ? p.mask + (size_t)tok0 : nullptr;
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            const int key = kv0 + w * 32 + lane;
            bool keep = key < L;
            if (keep && mrow) keep = mrow[key] != 0;
            kbits[w] = __ballot_sync(0xffffffffu, keep);
          }
        }
        const bool diag = p.causal && (kv0 + 63 > q0);
        const bool fast = !diag && (kbits[0] & kbits[1]) == 0xffffffffu;
        const int dcol = diag ? (qrow - kv0) : 0x7fffffff;
        mbar_wait(&bars->s_full[t], u & 1);
        tc_fence_after();
        mbar_wait(&bars->mma2_done[t], (u & 1) ^ 1);             // dS tile of half n-2 consumed
        const uint32_t tS = tmem + (uint32_t)(t * 128);
        unsigned char* hd = sdS + t * HALF_BYTES;
#pragma unroll 1
        for (int cc = kShareHalf ? g * 32 : 0; cc < (kShareHalf ? g * 32 + 32 : 64); cc += 32) {
          uint32_t vs[32], vp[32];
          tmem_ld_32x32(tS + lane_off + cc, vs);
          tmem_ld_32x32(tS + 64 + lane_off + cc, vp);
          float dsc[32];
          if constexpr (DROP) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) drop_scale8(p.drop, dstream, dgrow + (unsigned long long)(((kv0 + cc) >> 3) + q4), dsc + q4 * 8);
          }
          tmem_ld_wait();
          if (kShareHalf || cc == 32) {
            tc_fence_before();
            mbar_arrive(&bars->s_free[t]);
          }
          float ds[32];
          if (fast) {
#pragma unroll
            for (int x = 0; x < 32; ++x) {
              const float pr = ex2_approx(fmaf(__uint_as_float(vs[x]), sl2, nl));
              if constexpr (DROP) ds[x] = pr * fmaf(__uint_as_float(vp[x]), dsc[x], -dl);
              else                ds[x] = pr * (__uint_as_float(vp[x]) - dl);
            }
          } else {
            const uint32_t kw = kbits[cc >> 5];
#pragma unroll
            for (int x = 0; x < 32; ++x) {
              const int drop = ((dcol - (cc + x)) >> 31) | (int)(((kw >> x) & 1u) - 1u);
              const uint32_t bits = (__float_as_uint(__uint_as_float(vs[x]) * sl2) & ~(uint32_t)drop) | (0xff800000u & (uint32_t)drop);
              const float pr = ex2_approx(__uint_as_float(bits) + nl);
              if constexpr (DROP) ds[x] = pr * fmaf(__uint_as_float(vp[x]), dsc[x], -dl);
              else                ds[x] = pr * (__uint_as_float(vp[x]) - dl);
            }
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) st_sw128(hd, r, (cc >> 3) + q4, pack8(ds + q4 * 8));
        }
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(&bars->p_ready[t]);
      }
      // ---- item end: dQ complete -> WG g drains columns [g*64, g*64+64) (TD = 64: WG0 alone); dS was unscaled: dQ *= scale ----
      mbar_wait(&bars->acc_full, c & 1);
      tc_fence_after();
      if (g * 64 < TD) {
        __nv_bfloat16* dqrow = p.dq + (size_t)(tok0 + (qrow < L ? qrow : 0)) * p.lddq + (size_t)h * TD + g * 64;
        drain_acc_wg<64>(tdQ + lane_off + (uint32_t)(g * 64), p.scale, sdS + g * HALF_BYTES, &tm_dq, h * TD + g * 64, tok0 + q0,
                         q0 + TB <= L, qrow < L, r, dqrow, 1 + g, wt == 0);
      }
      tc_fence_before();
      mbar_arrive(&bars->acc_free);
      named_bar_sync(3, 256);
      n0 += NH;
    }
    if (wt == 0) bulk_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int D> constexpr int dkv_pipe_smem() {
  return 2 * pipe_res<D>() * TcD<D>::TILE + 2 * pipe_stages<D>() * (D / 64) * CH64 + 4 * HALF_BYTES + 1024 + 256 + 2 * pipe_stages<D>() * 256;
}
template <int D> constexpr int dq_pipe_smem() {
  return 2 * pipe_res<D>() * TcD<D>::TILE + 2 * pipe_stages<D>() * (D / 64) * CH64 + 2 * HALF_BYTES + 1024 + 256 + 2 * kMaxKeyWords * 4;
}
static_assert(dkv_pipe_smem<128>() <= 232448 && dq_pipe_smem<128>() <= 232448, "pipelined attention backward exceeds 227 KB of shared memory");

template <int D> constexpr int dkv_smem() { return 4 * TcD<D>::TILE + 2 * TILE_BYTES + 1024 + 2048; }
template <int D> constexpr int dq_smem() { return 4 * TcD<D>::TILE + TILE_BYTES + 1024 + 1024; }
template <int D> constexpr int fwd_smem() { return 3 * TcD<D>::TILE + (D == 128 ? 0 : TILE_BYTES) + 1024 + 1024; }

}  // namespace dalm

using namespace dalm;

static long long* g_attn_dbg = nullptr;
static int g_attn_bwd_pipe = 1;               // 1: pipelined persistent backward kernels (default), 0: the one-chain-per-CTA kernels
extern "C" void dalm_b200_attention_tc_set_mode(int pipelined_backward) { g_attn_bwd_pipe = pipelined_backward; }
extern "C" void dalm_b200_attention_tc_set_debug(void* dev_buffer_64xint64) { g_attn_dbg = (long long*)dev_buffer_64xint64; }

static int tc_maps(const void* ptr, long long rows, long long cols, long long ld, CUtensorMap* m) {
  return get_tmap(ptr, rows, cols, ld, 128, m, 0);
}

template <int D, bool DROP>
static int launch_tc_fwd(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mo,
                         const AttnTcParams& p, cudaStream_t st) {
  static bool attr = false;
  if (!attr) { DALM_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<D, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem<D>())); attr = true; }
  dim3 grid((p.L + TB - 1) / TB, p.Hq, p.B);
  attn_fwd_tc_kernel<D, DROP><<<grid, 160, fwd_smem<D>(), st>>>(mq, mk, mv, mo, p);
  count_launch();
  return check_launch("attn_fwd_tc_kernel");
}

// q/k/v: bf16 token-major 2-D matrices [B*L, ncols] (row stride ld*), head h of q at column qcol0 + h*D etc.
extern "C" int dalm_b200_attention_tc_fwd(const void* q, long long ldq, long long qcols, int qcol0, const void* k,
                                          long long ldk, long long kcols, int kcol0, const void* v, long long ldv,
                                          long long vcols, int vcol0, const int64_t* mask, void* out, long long ldo,
                                          float* lse, int B, int L, int Hq, int Hkv, int D, float scale, int causal,
                                          float drop_p, unsigned long long drop_seed, unsigned long long drop_stream_id,
                                          const void* drop_offset, void* stream) {
  DALM_REQUIRE(D == 128 || D == 64, "attention_tc: head_dim must be 64 or 128 (got %d)", D);
  DALM_REQUIRE(B > 0 && L > 0 && Hq > 0 && Hkv > 0 && Hq % Hkv == 0, "attention_tc: bad shape");
  DALM_REQUIRE((ldo % 8) == 0 && ((uintptr_t)out & 15) == 0, "attention_tc: output alignment");
  DALM_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "attention_tc: dropout p must be in [0,1)");
  DALM_REQUIRE(drop_p == 0.f || D == 64, "attention_tc: probability dropout is built for head_dim 64 (BERT encoder); Llama has attention_dropout = 0");
  CUtensorMap mq, mk, mv, mo;
  const long long rows = (long long)B * L;
  if (int e = tc_maps(q, rows, qcols, ldq, &mq)) return e;
  if (int e = tc_maps(k, rows, kcols, ldk, &mk)) return e;
  if (int e = tc_maps(v, rows, vcols, ldv, &mv)) return e;
  if (int e = tc_maps(out, rows, (long long)Hq * D, ldo, &mo)) return e;
  AttnTcParams p{};
  p.mask = mask; p.o = (__nv_bfloat16*)out; p.ldo = ldo; p.lse = lse; p.B = B; p.L = L; p.Hq = Hq; p.Hkv = Hkv;
  p.scale = scale; p.causal = causal; p.qcol0 = qcol0; p.kcol0 = kcol0; p.vcol0 = vcol0;
  p.dbg = g_attn_dbg;
  p.drop = make_drop(drop_p, drop_seed, drop_stream_id, drop_offset);
  cudaStream_t st = (cudaStream_t)stream;
  if (D == 128) return launch_tc_fwd<128, false>(mq, mk, mv, mo, p, st);
  if (drop_p > 0.f) return launch_tc_fwd<64, true>(mq, mk, mv, mo, p, st);
  return launch_tc_fwd<64, false>(mq, mk, mv, mo, p, st);
}

template <int D, bool DROP>
static int launch_tc_bwd(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mdo,
                         const CUtensorMap& mq64, const CUtensorMap& mk64, const CUtensorMap& mv64, const CUtensorMap& mdo64,
                         const CUtensorMap& mdq, const CUtensorMap& mdk, const CUtensorMap& mdv, const AttnTcParams& p,
                         const void* out, long long ldo, const void* d_out, long long lddo, float* delta, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    DALM_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<D, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, dkv_smem<D>()));
    DALM_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<D, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, dq_smem<D>()));
    attr = true;
  }
  const int total_warps = p.B * p.L * p.Hq;
  attn_tc_delta_kernel<D><<<(total_warps * 32 + 255) / 256, 256, 0, st>>>((const __nv_bfloat16*)out, ldo, (const __nv_bfloat16*)d_out,
                                                                         lddo, p.lse, delta, delta + (size_t)p.B * p.Hq * p.Lp,
                                                                         p.B, p.L, p.Lp, p.Hq);
  if (int e = check_launch("attn_tc_delta_kernel")) return e;
  const int ntiles = (p.L + TB - 1) / TB;
  if (g_attn_bwd_pipe) {
    static bool attr2 = false;
    if (!attr2) {
      DALM_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_pipe_kernel<D, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, dkv_pipe_smem<D>()));
      DALM_CUDA(cudaFuncSetAttribute(attn_bwd_dq_pipe_kernel<D, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, dq_pipe_smem<D>()));
      attr2 = true;
    }
    const int items_kv = ntiles * p.Hkv * p.B, items_q = ntiles * p.Hq * p.B;
    attn_bwd_dkv_pipe_kernel<D, DROP><<<items_kv < kNumSMs ? items_kv : kNumSMs, kPipeThreads, dkv_pipe_smem<D>(), st>>>(
        mq64, mk, mv, mdo64, mdk, mdv, p);
    if (int e = check_launch("attn_bwd_dkv_pipe_kernel")) return e;
    attn_bwd_dq_pipe_kernel<D, DROP><<<items_q < kNumSMs ? items_q : kNumSMs, kPipeThreads, dq_pipe_smem<D>(), st>>>(
        mq, mk64, mv64, mdo, mdq, p);
    count_launch(3);
    return check_launch("attn_bwd_dq_pipe_kernel");
  }
  attn_bwd_dkv_tc_kernel<D, DROP><<<dim3(ntiles, p.Hkv, p.B), 160, dkv_smem<D>(), st>>>(mq, mk, mv, mdo, mdk, mdv, p);
  if (int e = check_launch("attn_bwd_dkv_tc_kernel")) return e;
  attn_bwd_dq_tc_kernel<D, DROP><<<dim3(ntiles, p.Hq, p.B), 160, dq_smem<D>(), st>>>(mq, mk, mv, mdo, mdq, p);
  count_launch(3);
  return check_launch("attn_bwd_dq_tc_kernel");
}

// backward: d_out bf16 [B*L, Hq*D (docols)], delta: fp32 workspace of 2*B*Hq*Lp floats (Lp = L rounded up to 64); dq/dk/dv bf16 token-major outputs
extern "C" int dalm_b200_attention_tc_bwd(const void* q, long long ldq, long long qcols, const void* k, long long ldk,
                                          long long kcols, const void* v, long long ldv, long long vcols,
                                          const int64_t* mask, const void* out, long long ldo, const float* lse,
                                          const void* d_out, long long lddo, long long docols, float* delta, void* dq,
                                          long long lddq, void* dk, long long lddk, void* dv, long long lddv, int B, int L,
                                          int Hq, int Hkv, int D, float scale, int causal, float drop_p,
                                          unsigned long long drop_seed, unsigned long long drop_stream_id,
                                          const void* drop_offset, void* stream) {
  DALM_REQUIRE(D == 128 || D == 64, "attention_tc_bwd: head_dim must be 64 or 128 (got %d)", D);
  DALM_REQUIRE(B > 0 && L > 0 && Hq > 0 && Hkv > 0 && Hq % Hkv == 0, "attention_tc_bwd: bad shape");
  DALM_REQUIRE((lddq % 8) == 0 && (lddk % 8) == 0 && (lddv % 8) == 0 && (ldo % 4) == 0 && (lddo % 4) == 0, "attention_tc_bwd: strides");
  DALM_REQUIRE(((uintptr_t)dq & 15) == 0 && ((uintptr_t)dk & 15) == 0 && ((uintptr_t)dv & 15) == 0, "attention_tc_bwd: output alignment");
  DALM_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || D == 64), "attention_tc_bwd: dropout needs p in [0,1) and head_dim 64");
  CUtensorMap mq, mk, mv, mdo;
  const long long rows = (long long)B * L;
  if (int e = tc_maps(q, rows, qcols, ldq, &mq)) return e;
  if (int e = tc_maps(k, rows, kcols, ldk, &mk)) return e;
  if (int e = tc_maps(v, rows, vcols, ldv, &mv)) return e;
  if (int e = tc_maps(d_out, rows, docols, lddo, &mdo)) return e;
  CUtensorMap mq64, mk64, mv64, mdo64;                         // 64-row boxes: the streamed halves of the pipelined kernels
  if (int e = get_tmap(q, rows, qcols, ldq, 64, &mq64, 0)) return e;
  if (int e = get_tmap(k, rows, kcols, ldk, 64, &mk64, 0)) return e;
  if (int e = get_tmap(v, rows, vcols, ldv, 64, &mv64, 0)) return e;
  if (int e = get_tmap(d_out, rows, docols, lddo, 64, &mdo64, 0)) return e;
  CUtensorMap mdq, mdk, mdv;
  if (int e = tc_maps(dq, rows, (long long)Hq * D, lddq, &mdq)) return e;
  if (int e = tc_maps(dk, rows, (long long)Hkv * D, lddk, &mdk)) return e;
  if (int e = tc_maps(dv, rows, (long long)Hkv * D, lddv, &mdv)) return e;
  AttnTcParams p{};
  p.mask = mask; p.lse = const_cast<float*>(lse); p.delta = delta; p.B = B; p.L = L; p.Hq = Hq; p.Hkv = Hkv;
  p.Lp = (L + 63) / 64 * 64;
  p.nl2 = delta + (size_t)B * Hq * p.Lp;                          // second half of the workspace (see attn_tc_delta_kernel)
  p.scale = scale; p.causal = causal;
  p.dq = (__nv_bfloat16*)dq; p.dk = (__nv_bfloat16*)dk; p.dv = (__nv_bfloat16*)dv; p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  p.dbg = g_attn_dbg;
  p.drop = make_drop(drop_p, drop_seed, drop_stream_id, drop_offset);
  cudaStream_t st = (cudaStream_t)stream;
  if (D == 128) return launch_tc_bwd<128, false>(mq, mk, mv, mdo, mq64, mk64, mv64, mdo64, mdq, mdk, mdv, p, out, ldo, d_out, lddo, delta, st);
  if (drop_p > 0.f) return launch_tc_bwd<64, true>(mq, mk, mv, mdo, mq64, mk64, mv64, mdo64, mdq, mdk, mdv, p, out, ldo, d_out, lddo, delta, st);
  return launch_tc_bwd<64, false>(mq, mk, mv, mdo, mq64, mk64, mv64, mdo64, mdq, mdk, mdv, p, out, ldo, d_out, lddo, delta, st);
}
