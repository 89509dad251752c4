// dalm_b200 — fused multi-head attention, forward and backward, for the encoder (bidirectional + key-padding mask,
// head_dim 32/64) and the decoder (causal + key-padding mask, head_dim 64/128, MHA / GQA / MQA).
//
// Replaces the attention inside HF BertModel / LlamaForCausalLM / FalconForCausalLM that the reference reaches through
// dalm/models/rag_e2e_base_model.py:93,105 and its autograd backward. Flash-style: scores never touch HBM; the forward
// stores only the per-row log-sum-exp, the backward recomputes P tile by tile.
//
// These are the warp-level mma.sync (HMMA) kernels with ldmatrix-fed fragments. The training paths at head_dim 64 / 128
// run the tcgen05/TMEM kernels of csrc/attention_tc.cu; this file serves head_dim 32 (bge-small, cfg-1) and is the
// independent cross-check the tcgen05 kernels are tested against (same masks, same dropout element indexing).
//
//   forward : grid (ceil(L/64), Hq, B), 4 warps, each warp owns 16 query rows, KV streamed in 64-key tiles
//   dKdV    : grid (ceil(L/64), Hkv, B), each warp owns 16 keys, loops over the q heads of its group and 32-query tiles
//   dQ      : grid (ceil(L/64), Hq, B), each warp owns 16 queries, loops over 64-key tiles
#include "common.cuh"

namespace dalm {

struct AttnParams {
  const __nv_bfloat16* q; const __nv_bfloat16* k; const __nv_bfloat16* v;   // token-major: row = b*L + l
  long long ldq, ldk, ldv;          // row strides (elements); head h lives at column h*D (k/v: (h/group)*D)
  const int64_t* mask;              // [B,L] key-padding mask (1 keep / 0 drop) or nullptr
  __nv_bfloat16* o; long long ldo;  // [B*L, Hq*D]
  float* lse;                       // [B,Hq,L]
  // backward only
  const __nv_bfloat16* d_o; long long lddo;
  float* delta;                     // [B,Hq,L] rowsum(dO * O)
  __nv_bfloat16* dq; __nv_bfloat16* dk; __nv_bfloat16* dv;
  long long lddq, lddk, lddv;
  int B, L, Hq, Hkv;
  float scale;                      // 1/sqrt(D)
  int causal;
  DropCfg drop;                     // attention-probability dropout (BERT attention_probs_dropout_prob); p = 0 => off.
                                    // element index of P[b,h,i,j] = ((b*Hq + h)*L + i)*Lp + j, Lp = L rounded up to 8
                                    // (rows start on a Philox group of 8, so one call covers an 8-key MMA n-tile)
};

// ---------------------------------------------------------------- fragment helpers
__device__ __forceinline__ void ldsm_x4(uint32_t* r, const void* smem_ptr) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_ptr));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t* r, const void* smem_ptr) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_ptr));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
// D(16x8, f32) += A(16x16, bf16 row) * B(16x8, bf16 col)
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// cooperative copy of a [ROWS x D] bf16 tile (global row stride ld) into padded smem (row stride D+8); rows >= nvalid
// are zero-filled. 16-byte vector accesses (D % 8 == 0, ld % 8 == 0, 16B-aligned base).
template <int ROWS, int D, int NT>
__device__ __forceinline__ void load_tile(__nv_bfloat16* s, const __nv_bfloat16* g, long long ld, int nvalid) {
  constexpr int CH = D / 8;
  for (int i = threadIdx.x; i < ROWS * CH; i += NT) {
    const int r = i / CH, c = i - r * CH;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < nvalid) v = __ldg(reinterpret_cast<const uint4*>(g + (size_t)r * ld + c * 8));
    *reinterpret_cast<uint4*>(s + r * (D + 8) + c * 8) = v;
  }
}

// A-operand fragments (16 rows x 16 cols at [row0, col0]) from a padded smem tile
template <int D>
__device__ __forceinline__ void load_a_frag(uint32_t* a, const __nv_bfloat16* s, int row0, int col0, int lane) {
  ldsm_x4(a, s + (row0 + (lane & 15)) * (D + 8) + col0 + ((lane >> 4) << 3));
}
// B-operand fragments for TWO adjacent n-tiles (16 "n" rows at n0) x one k-step (16 cols at k0), tile stored [n][k]:
//   b[0],b[1] -> n-tile n0 ; b[2],b[3] -> n-tile n0+8
template <int D>
__device__ __forceinline__ void load_b_frag_nk(uint32_t* b, const __nv_bfloat16* s, int n0, int k0, int lane) {
  ldsm_x4(b, s + (n0 + (lane & 7) + ((lane >> 4) << 3)) * (D + 8) + k0 + (((lane >> 3) & 1) << 3));
}
// B-operand fragments for TWO adjacent n-tiles (16 cols at n0) x one k-step (16 "k" rows at k0), tile stored [k][n]:
//   b[0],b[1] -> n-tile n0 ; b[2],b[3] -> n-tile n0+8
template <int D>
__device__ __forceinline__ void load_b_frag_kn(uint32_t* b, const __nv_bfloat16* s, int k0, int n0, int lane) {
  ldsm_x4_t(b, s + (k0 + (lane & 7) + (((lane >> 3) & 1) << 3)) * (D + 8) + n0 + ((lane >> 4) << 3));
}

// ============================================================================================================
// forward
// ============================================================================================================
template <int D, bool DROP>
__global__ void __launch_bounds__(128) attn_fwd_kernel(AttnParams p) {
  constexpr int BQ = 64, BKV = 64, LDS = D + 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sK = sQ + BQ * LDS;
  __nv_bfloat16* sV = sK + BKV * LDS;
  float* sMask = reinterpret_cast<float*>(sV + BKV * LDS);      // [BKV] additive 0 / -inf

  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.Hq / p.Hkv);
  const int L = p.L, q0 = qb * BQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const size_t tok0 = (size_t)b * L;

  load_tile<BQ, D, 128>(sQ, p.q + (tok0 + q0) * p.ldq + (size_t)h * D, p.ldq, min(BQ, L - q0));

  float o_acc[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const float sl2 = p.scale * 1.4426950408889634f;               // scores are exponentiated in base 2
  const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;       // the two query rows this thread holds

  const int kv_end = p.causal ? min(L, q0 + BQ) : L;
  for (int kv0 = 0; kv0 < kv_end; kv0 += BKV) {
    __syncthreads();                                            // previous tile fully consumed
    const int nvalid = min(BKV, L - kv0);
    load_tile<BKV, D, 128>(sK, p.k + (tok0 + kv0) * p.ldk + (size_t)hk * D, p.ldk, nvalid);
    load_tile<BKV, D, 128>(sV, p.v + (tok0 + kv0) * p.ldv + (size_t)hk * D, p.ldv, nvalid);
    if (threadIdx.x < BKV) {
      const int key = kv0 + threadIdx.x;
      bool keep = key < L;
      if (keep && p.mask) keep = p.mask[tok0 + key] != 0;
      sMask[threadIdx.x] = keep ? 0.f : -INFINITY;
    }
    __syncthreads();

    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[BKV / 8][4];
#pragma unroll
    for (int i = 0; i < BKV / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      uint32_t a[4];
      load_a_frag<D>(a, sQ, warp * 16, kk * 16, lane);
#pragma unroll
      for (int nt = 0; nt < BKV / 16; ++nt) {
        uint32_t bf[4];
        load_b_frag_nk<D>(bf, sK, nt * 16, kk * 16, lane);
        mma16816(s[2 * nt], a, bf[0], bf[1]);
        mma16816(s[2 * nt + 1], a, bf[2], bf[3]);
      }
    }
    // ---- mask + online softmax (base-2) ----
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < BKV / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kc = nt * 8 + t * 2 + (e & 1);
        const int qr = (e < 2) ? row_a : row_b;
        float val = s[nt][e] * sl2 + sMask[kc];
        if (p.causal && (kv0 + kc) > qr) val = -INFINITY;
        s[nt][e] = val;
        mx[e >> 1] = fmaxf(mx[e >> 1], val);
      }
    }
    float corr[2], mnew[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      mnew[r] = fmaxf(m_run[r], mx[r]);
      const float msafe = (mnew[r] == -INFINITY) ? 0.f : mnew[r];
      corr[r] = exp2f(m_run[r] - msafe);                        // m_run = -inf -> 0
      m_run[r] = mnew[r];
      mnew[r] = msafe;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < BKV / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = exp2f(s[nt][e] - mnew[e >> 1]);
        s[nt][e] = pv;
        rs[e >> 1] += pv;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
      l_run[r] = l_run[r] * corr[r] + rs[r];
    }
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      o_acc[i][0] *= corr[0]; o_acc[i][1] *= corr[0];
      o_acc[i][2] *= corr[1]; o_acc[i][3] *= corr[1];
    }
    if (DROP) {
      // dropout acts on the normalised probabilities; the row sum above used the un-dropped values.
      // one Philox call per (row, 8-key n-tile); this thread uses components t*2, t*2+1
      const unsigned long long dstream = drop_stream(p.drop);
      const unsigned long long rbase = ((unsigned long long)b * p.Hq + h) * L;
      const int lp8 = (L + 7) >> 3;
#pragma unroll
      for (int nt = 0; nt < BKV / 8; ++nt) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int qr = r == 0 ? row_a : row_b;
          float sc[8];
          drop_scale8(p.drop, dstream, (rbase + qr) * lp8 + ((kv0 >> 3) + nt), sc);
          float s0 = sc[0], s1 = sc[1];
#pragma unroll
          for (int j = 1; j < 4; ++j) if (t == j) { s0 = sc[2 * j]; s1 = sc[2 * j + 1]; }
          s[nt][2 * r] *= s0; s[nt][2 * r + 1] *= s1;
        }
      }
    }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < BKV / 16; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int nt = 0; nt < D / 16; ++nt) {
        uint32_t bf[4];
        load_b_frag_kn<D>(bf, sV, kk * 16, nt * 16, lane);
        mma16816(o_acc[2 * nt], a, bf[0], bf[1]);
        mma16816(o_acc[2 * nt + 1], a, bf[2], bf[3]);
      }
    }
  }

  // ---- normalise, write O and LSE (natural log of sum exp(scale * s)) ----
  const float inv_l[2] = {l_run[0] > 0.f ? 1.f / l_run[0] : 0.f, l_run[1] > 0.f ? 1.f / l_run[1] : 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = r == 0 ? row_a : row_b;
    if (row < L) {
      __nv_bfloat16* orow = p.o + (tok0 + row) * p.ldo + (size_t)h * D;
#pragma unroll
      for (int nt = 0; nt < D / 8; ++nt) {
        const uint32_t pk = pack_bf16(o_acc[nt][2 * r] * inv_l[r], o_acc[nt][2 * r + 1] * inv_l[r]);
        *reinterpret_cast<uint32_t*>(orow + nt * 8 + t * 2) = pk;
      }
      if (t == 0) {
        // fully masked row: +inf makes the backward's exp(s - lse) vanish
        const float lse = l_run[r] > 0.f ? (m_run[r] + log2f(l_run[r])) * 0.6931471805599453f : INFINITY;
        p.lse[((size_t)b * p.Hq + h) * L + row] = lse;
      }
    }
  }
}

// ============================================================================================================
// backward pre-pass: delta[b,h,i] = sum_d dO[i,d] * O[i,d]
// ============================================================================================================
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, long long ldo, const __nv_bfloat16* __restrict__ d_o,
                                  long long lddo, float* __restrict__ delta, int B, int L, int Hq, int D) {
  // one warp per (token, head)
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int total = B * L * Hq;
  if (gw >= total) return;
  const int tok = gw / Hq, h = gw - tok * Hq;
  const __nv_bfloat16* a = o + (size_t)tok * ldo + (size_t)h * D;
  const __nv_bfloat16* c = d_o + (size_t)tok * lddo + (size_t)h * D;
  float acc = 0.f;
  for (int d = lane * 2; d < D; d += 64) {
    const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(a + d));
    const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(c + d));
    acc += x.x * y.x + x.y * y.y;
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    const int b = tok / L, l = tok - b * L;
    delta[((size_t)b * Hq + h) * L + l] = acc;
  }
}

// ============================================================================================================
// backward: dK, dV.  Each warp owns 16 keys; queries streamed in 32-row tiles.
// ============================================================================================================
template <int D, bool DROP>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(AttnParams p) {
  constexpr int BKV = 64, BQ = 32, LDS = D + 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* sK  = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sV  = sK + BKV * LDS;
  __nv_bfloat16* sQ  = sV + BKV * LDS;
  __nv_bfloat16* sdO = sQ + BQ * LDS;
  float* sLse   = reinterpret_cast<float*>(sdO + BQ * LDS);     // [BQ]
  float* sDelta = sLse + BQ;                                    // [BQ]
  float* sMask  = sDelta + BQ;                                  // [BKV]

  const int kb = blockIdx.x, hk = blockIdx.y, b = blockIdx.z;
  const int group = p.Hq / p.Hkv;
  const int L = p.L, kv0 = kb * BKV;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const size_t tok0 = (size_t)b * L;
  const float sl2 = p.scale * 1.4426950408889634f;

  const int nvalid_kv = min(BKV, L - kv0);
  load_tile<BKV, D, 128>(sK, p.k + (tok0 + kv0) * p.ldk + (size_t)hk * D, p.ldk, nvalid_kv);
  load_tile<BKV, D, 128>(sV, p.v + (tok0 + kv0) * p.ldv + (size_t)hk * D, p.ldv, nvalid_kv);
  if (threadIdx.x < BKV) {
    const int key = kv0 + threadIdx.x;
    bool keep = key < L;
    if (keep && p.mask) keep = p.mask[tok0 + key] != 0;
    sMask[threadIdx.x] = keep ? 0.f : -INFINITY;
  }

  float dk_acc[D / 8][4], dv_acc[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) {
    dk_acc[i][0] = dk_acc[i][1] = dk_acc[i][2] = dk_acc[i][3] = 0.f;
    dv_acc[i][0] = dv_acc[i][1] = dv_acc[i][2] = dv_acc[i][3] = 0.f;
  }
  const int key_a = kv0 + warp * 16 + g, key_b = key_a + 8;     // the two keys (rows of S^T) this thread holds
  const int q_begin = p.causal ? (kv0 / BQ) * BQ : 0;            // queries before the key tile see none of it

  for (int hq = hk * group; hq < (hk + 1) * group; ++hq) {
    for (int q0 = q_begin; q0 < L; q0 += BQ) {
      __syncthreads();
      const int nq = min(BQ, L - q0);
      load_tile<BQ, D, 128>(sQ, p.q + (tok0 + q0) * p.ldq + (size_t)hq * D, p.ldq, nq);
      load_tile<BQ, D, 128>(sdO, p.d_o + (tok0 + q0) * p.lddo + (size_t)hq * D, p.lddo, nq);
      if (threadIdx.x < BQ) {
        const int qi = q0 + threadIdx.x;
        const size_t idx = ((size_t)b * p.Hq + hq) * L + qi;
        sLse[threadIdx.x]   = qi < L ? p.lse[idx] * 1.4426950408889634f : INFINITY;   // base-2 units
        sDelta[threadIdx.x] = qi < L ? p.delta[idx] : 0.f;
      }
      __syncthreads();

      // ---- S^T = K Q^T : 16 keys x 32 queries per warp ----
      float st[BQ / 8][4];
#pragma unroll
      for (int i = 0; i < BQ / 8; ++i) { st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        uint32_t a[4];
        load_a_frag<D>(a, sK, warp * 16, kk * 16, lane);
#pragma unroll
        for (int nt = 0; nt < BQ / 16; ++nt) {
          uint32_t bf[4];
          load_b_frag_nk<D>(bf, sQ, nt * 16, kk * 16, lane);
          mma16816(st[2 * nt], a, bf[0], bf[1]);
          mma16816(st[2 * nt + 1], a, bf[2], bf[3]);
        }
      }
      // ---- P^T = exp2(S^T * sl2 - lse2[q]) with masks ----
      const float mk_a = sMask[warp * 16 + g], mk_b = sMask[warp * 16 + g + 8];
#pragma unroll
      for (int nt = 0; nt < BQ / 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qc = nt * 8 + t * 2 + (e & 1);
          const int key = (e < 2) ? key_a : key_b;
          float val = st[nt][e] * sl2 + ((e < 2) ? mk_a : mk_b);
          if (p.causal && key > (q0 + qc)) val = -INFINITY;
          st[nt][e] = exp2f(val - sLse[qc]);                     // -inf - x -> 0 ; x - (+inf) -> 0
        }
      }
      // dropout scale of each (key, query) element this thread holds (1 when dropout is off)
      float ms[DROP ? BQ / 8 : 1][4];
      if (DROP) {
        const unsigned long long dstream = drop_stream(p.drop);
        const unsigned long long rbase = ((unsigned long long)b * p.Hq + hq) * L;
        const int lp8 = (L + 7) >> 3;
        const int kg = (kv0 >> 3) + warp * 2;                    // Philox group of key_a (key_b is the next group), component g
#pragma unroll
        for (int nt = 0; nt < BQ / 8; ++nt) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int qi = q0 + nt * 8 + t * 2 + c;
            float sa[8], sb[8];
            drop_scale8(p.drop, dstream, (rbase + qi) * lp8 + kg, sa);
            drop_scale8(p.drop, dstream, (rbase + qi) * lp8 + kg + 1, sb);
            float va = sa[0], vb = sb[0];
#pragma unroll
            for (int j = 1; j < 8; ++j) if (g == j) { va = sa[j]; vb = sb[j]; }
            ms[nt][c] = va; ms[nt][2 + c] = vb;                  // e = c: key_a ; e = 2 + c: key_b
          }
        }
      }
      // ---- dV += P_drop^T dO ----
      uint32_t pa[BQ / 16][4];
#pragma unroll
      for (int kk = 0; kk < BQ / 16; ++kk) {
        if (DROP) {
          pa[kk][0] = pack_bf16(st[2 * kk][0] * ms[2 * kk][0], st[2 * kk][1] * ms[2 * kk][1]);
          pa[kk][1] = pack_bf16(st[2 * kk][2] * ms[2 * kk][2], st[2 * kk][3] * ms[2 * kk][3]);
          pa[kk][2] = pack_bf16(st[2 * kk + 1][0] * ms[2 * kk + 1][0], st[2 * kk + 1][1] * ms[2 * kk + 1][1]);
          pa[kk][3] = pack_bf16(st[2 * kk + 1][2] * ms[2 * kk + 1][2], st[2 * kk + 1][3] * ms[2 * kk + 1][3]);
        } else {
          pa[kk][0] = pack_bf16(st[2 * kk][0], st[2 * kk][1]);
          pa[kk][1] = pack_bf16(st[2 * kk][2], st[2 * kk][3]);
          pa[kk][2] = pack_bf16(st[2 * kk + 1][0], st[2 * kk + 1][1]);
          pa[kk][3] = pack_bf16(st[2 * kk + 1][2], st[2 * kk + 1][3]);
        }
      }
#pragma unroll
      for (int kk = 0; kk < BQ / 16; ++kk) {
#pragma unroll
        for (int nt = 0; nt < D / 16; ++nt) {
          uint32_t bf[4];
          load_b_frag_kn<D>(bf, sdO, kk * 16, nt * 16, lane);
          mma16816(dv_acc[2 * nt], pa[kk], bf[0], bf[1]);
          mma16816(dv_acc[2 * nt + 1], pa[kk], bf[2], bf[3]);
        }
      }
      // ---- dP^T = V dO^T : 16 keys x 32 queries ----
      float dpt[BQ / 8][4];
#pragma unroll
      for (int i = 0; i < BQ / 8; ++i) { dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        uint32_t a[4];
        load_a_frag<D>(a, sV, warp * 16, kk * 16, lane);
#pragma unroll
        for (int nt = 0; nt < BQ / 16; ++nt) {
          uint32_t bf[4];
          load_b_frag_nk<D>(bf, sdO, nt * 16, kk * 16, lane);
          mma16816(dpt[2 * nt], a, bf[0], bf[1]);
          mma16816(dpt[2 * nt + 1], a, bf[2], bf[3]);
        }
      }
      // ---- dS^T = P^T * (dP^T - delta[q]) * scale ;  dK += dS^T Q ----
      uint32_t da[BQ / 16][4];
#pragma unroll
      for (int kk = 0; kk < BQ / 16; ++kk) {
        float ds[2][4];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int nt = 2 * kk + half;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int qc = nt * 8 + t * 2 + (e & 1);
            const float dpv = DROP ? dpt[nt][e] * ms[nt][e] : dpt[nt][e];
            ds[half][e] = st[nt][e] * (dpv - sDelta[qc]) * p.scale;
          }
        }
        da[kk][0] = pack_bf16(ds[0][0], ds[0][1]);
        da[kk][1] = pack_bf16(ds[0][2], ds[0][3]);
        da[kk][2] = pack_bf16(ds[1][0], ds[1][1]);
        da[kk][3] = pack_bf16(ds[1][2], ds[1][3]);
      }
#pragma unroll
      for (int kk = 0; kk < BQ / 16; ++kk) {
#pragma unroll
        for (int nt = 0; nt < D / 16; ++nt) {
          uint32_t bf[4];
          load_b_frag_kn<D>(bf, sQ, kk * 16, nt * 16, lane);
          mma16816(dk_acc[2 * nt], da[kk], bf[0], bf[1]);
          mma16816(dk_acc[2 * nt + 1], da[kk], bf[2], bf[3]);
        }
      }
    }
  }
  // ---- write dK, dV ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int key = r == 0 ? key_a : key_b;
    if (key < L) {
      __nv_bfloat16* dkrow = p.dk + (tok0 + key) * p.lddk + (size_t)hk * D;
      __nv_bfloat16* dvrow = p.dv + (tok0 + key) * p.lddv + (size_t)hk * D;
#pragma unroll
      for (int nt = 0; nt < D / 8; ++nt) {
        *reinterpret_cast<uint32_t*>(dkrow + nt * 8 + t * 2) = pack_bf16(dk_acc[nt][2 * r], dk_acc[nt][2 * r + 1]);
        *reinterpret_cast<uint32_t*>(dvrow + nt * 8 + t * 2) = pack_bf16(dv_acc[nt][2 * r], dv_acc[nt][2 * r + 1]);
      }
    }
  }
}

// ============================================================================================================
// backward: dQ.  Each warp owns 16 queries; keys streamed in 64-row tiles.
// ============================================================================================================
template <int D, bool DROP>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(AttnParams p) {
  constexpr int BQ = 64, BKV = 64, LDS = D + 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* sQ  = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sdO = sQ + BQ * LDS;
  __nv_bfloat16* sK  = sdO + BQ * LDS;
  __nv_bfloat16* sV  = sK + BKV * LDS;
  float* sMask = reinterpret_cast<float*>(sV + BKV * LDS);      // [BKV]

  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.Hq / p.Hkv);
  const int L = p.L, q0 = qb * BQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const size_t tok0 = (size_t)b * L;
  const float sl2 = p.scale * 1.4426950408889634f;

  const int nq = min(BQ, L - q0);
  load_tile<BQ, D, 128>(sQ, p.q + (tok0 + q0) * p.ldq + (size_t)h * D, p.ldq, nq);
  load_tile<BQ, D, 128>(sdO, p.d_o + (tok0 + q0) * p.lddo + (size_t)h * D, p.lddo, nq);

  const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;
  float lse2[2], dl[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = r == 0 ? row_a : row_b;
    const size_t idx = ((size_t)b * p.Hq + h) * L + row;
    lse2[r] = row < L ? p.lse[idx] * 1.4426950408889634f : INFINITY;
    dl[r]   = row < L ? p.delta[idx] : 0.f;
  }
  float dq_acc[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) { dq_acc[i][0] = dq_acc[i][1] = dq_acc[i][2] = dq_acc[i][3] = 0.f; }

  const int kv_end = p.causal ? min(L, q0 + BQ) : L;
  for (int kv0 = 0; kv0 < kv_end; kv0 += BKV) {
    __syncthreads();
    const int nvalid = min(BKV, L - kv0);
    load_tile<BKV, D, 128>(sK, p.k + (tok0 + kv0) * p.ldk + (size_t)hk * D, p.ldk, nvalid);
    load_tile<BKV, D, 128>(sV, p.v + (tok0 + kv0) * p.ldv + (size_t)hk * D, p.ldv, nvalid);
    if (threadIdx.x < BKV) {
      const int key = kv0 + threadIdx.x;
      bool keep = key < L;
      if (keep && p.mask) keep = p.mask[tok0 + key] != 0;
      sMask[threadIdx.x] = keep ? 0.f : -INFINITY;
    }
    __syncthreads();

    float s[BKV / 8][4], dp[BKV / 8][4];
#pragma unroll
    for (int i = 0; i < BKV / 8; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      uint32_t a[4], ad[4];
      load_a_frag<D>(a, sQ, warp * 16, kk * 16, lane);
      load_a_frag<D>(ad, sdO, warp * 16, kk * 16, lane);
#pragma unroll
      for (int nt = 0; nt < BKV / 16; ++nt) {
        uint32_t bf[4];
        load_b_frag_nk<D>(bf, sK, nt * 16, kk * 16, lane);
        mma16816(s[2 * nt], a, bf[0], bf[1]);
        mma16816(s[2 * nt + 1], a, bf[2], bf[3]);
        load_b_frag_nk<D>(bf, sV, nt * 16, kk * 16, lane);
        mma16816(dp[2 * nt], ad, bf[0], bf[1]);
        mma16816(dp[2 * nt + 1], ad, bf[2], bf[3]);
      }
    }
    float msq[DROP ? BKV / 8 : 1][4];
    if (DROP) {
      const unsigned long long dstream = drop_stream(p.drop);
      const unsigned long long rbase = ((unsigned long long)b * p.Hq + h) * L;
      const int lp8 = (L + 7) >> 3;
#pragma unroll
      for (int nt = 0; nt < BKV / 8; ++nt) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int qr = r == 0 ? row_a : row_b;
          float sc[8];
          drop_scale8(p.drop, dstream, (rbase + qr) * lp8 + ((kv0 >> 3) + nt), sc);
          float s0 = sc[0], s1 = sc[1];
#pragma unroll
          for (int j = 1; j < 4; ++j) if (t == j) { s0 = sc[2 * j]; s1 = sc[2 * j + 1]; }
          msq[nt][2 * r] = s0; msq[nt][2 * r + 1] = s1;
        }
      }
    }
    // dS = P * (dP - delta) * scale
#pragma unroll
    for (int nt = 0; nt < BKV / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kc = nt * 8 + t * 2 + (e & 1);
        const int qr = (e < 2) ? row_a : row_b;
        float val = s[nt][e] * sl2 + sMask[kc];
        if (p.causal && (kv0 + kc) > qr) val = -INFINITY;
        const float pv = exp2f(val - lse2[e >> 1]);
        float dpv = dp[nt][e];
        if (DROP) dpv *= msq[nt][e];
        s[nt][e] = pv * (dpv - dl[e >> 1]) * p.scale;
      }
    }
    // dQ += dS K
#pragma unroll
    for (int kk = 0; kk < BKV / 16; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int nt = 0; nt < D / 16; ++nt) {
        uint32_t bf[4];
        load_b_frag_kn<D>(bf, sK, kk * 16, nt * 16, lane);
        mma16816(dq_acc[2 * nt], a, bf[0], bf[1]);
        mma16816(dq_acc[2 * nt + 1], a, bf[2], bf[3]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = r == 0 ? row_a : row_b;
    if (row < L) {
      __nv_bfloat16* dqrow = p.dq + (tok0 + row) * p.lddq + (size_t)h * D;
#pragma unroll
      for (int nt = 0; nt < D / 8; ++nt)
        *reinterpret_cast<uint32_t*>(dqrow + nt * 8 + t * 2) = pack_bf16(dq_acc[nt][2 * r], dq_acc[nt][2 * r + 1]);
    }
  }
}

template <int D> static size_t fwd_smem() { return (size_t)(64 * 3) * (D + 8) * 2 + 64 * 4; }
template <int D> static size_t dkv_smem() { return (size_t)(64 * 2 + 32 * 2) * (D + 8) * 2 + (32 + 32 + 64) * 4; }
template <int D> static size_t dq_smem()  { return (size_t)(64 * 4) * (D + 8) * 2 + 64 * 4; }

template <int D, bool DROP> static int launch_fwd(const AttnParams& p, cudaStream_t st) {
  static bool attr = false;
  if (!attr) { DALM_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<D, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem<D>())); attr = true; }
  dim3 grid((p.L + 63) / 64, p.Hq, p.B);
  attn_fwd_kernel<D, DROP><<<grid, 128, fwd_smem<D>(), st>>>(p);
  count_launch();
  return check_launch("attn_fwd_kernel");
}
template <int D, bool DROP> static int launch_bwd(const AttnParams& p, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    DALM_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<D, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dkv_smem<D>()));
    DALM_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<D, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dq_smem<D>()));
    attr = true;
  }
  const int total_warps = p.B * p.L * p.Hq;
  attn_delta_kernel<<<(total_warps * 32 + 255) / 256, 256, 0, st>>>(p.o, p.ldo, p.d_o, p.lddo, p.delta, p.B, p.L, p.Hq, D);
  if (int e = check_launch("attn_delta_kernel")) return e;
  dim3 gkv((p.L + 63) / 64, p.Hkv, p.B);
  attn_bwd_dkv_kernel<D, DROP><<<gkv, 128, dkv_smem<D>(), st>>>(p);
  if (int e = check_launch("attn_bwd_dkv_kernel")) return e;
  dim3 gq((p.L + 63) / 64, p.Hq, p.B);
  attn_bwd_dq_kernel<D, DROP><<<gq, 128, dq_smem<D>(), st>>>(p);
  count_launch(3);
  return check_launch("attn_bwd_dq_kernel");
}

static int check_common(const AttnParams& p, int D) {
  DALM_REQUIRE(D == 32 || D == 64 || D == 128, "attention: head_dim %d unsupported (32/64/128)", D);
  DALM_REQUIRE(p.B > 0 && p.L > 0 && p.Hq > 0 && p.Hkv > 0 && p.Hq % p.Hkv == 0, "attention: bad shape B=%d L=%d Hq=%d Hkv=%d", p.B, p.L, p.Hq, p.Hkv);
  DALM_REQUIRE(p.ldq % 8 == 0 && p.ldk % 8 == 0 && p.ldv % 8 == 0 && p.ldo % 2 == 0, "attention: strides must keep 16-byte row alignment");
  DALM_REQUIRE(((uintptr_t)p.q & 15) == 0 && ((uintptr_t)p.k & 15) == 0 && ((uintptr_t)p.v & 15) == 0, "attention: q/k/v must be 16-byte aligned");
  return 0;
}
}  // namespace dalm

using namespace dalm;

// q/k/v: bf16 token-major views (row b*L+l, head h at column h*D); mask: int64 [B,L] or NULL; out: bf16; lse: fp32 [B,Hq,L]
extern "C" int dalm_b200_attention_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                                       long long ldv, const int64_t* mask, void* out, long long ldo, float* lse, int B,
                                       int L, int Hq, int Hkv, int D, float scale, int causal, float drop_p,
                                       unsigned long long drop_seed, unsigned long long drop_stream_id,
                                       const void* drop_offset, void* stream) {
  AttnParams p{};
  p.q = (const __nv_bfloat16*)q; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.mask = mask; p.o = (__nv_bfloat16*)out; p.ldo = ldo; p.lse = lse;
  p.B = B; p.L = L; p.Hq = Hq; p.Hkv = Hkv; p.scale = scale; p.causal = causal;
  p.drop = make_drop(drop_p, drop_seed, drop_stream_id, drop_offset);
  if (int e = check_common(p, D)) return e;
  DALM_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "attention: dropout p must be in [0,1)");
  DALM_REQUIRE(drop_p == 0.f || D <= 64, "attention: probability dropout is built for head_dim <= 64 (encoder); Llama has attention_dropout = 0");
  cudaStream_t st = (cudaStream_t)stream;
  if (drop_p > 0.f) return D == 32 ? launch_fwd<32, true>(p, st) : launch_fwd<64, true>(p, st);
  if (D == 32) return launch_fwd<32, false>(p, st);
  if (D == 64) return launch_fwd<64, false>(p, st);
  return launch_fwd<128, false>(p, st);
}

// delta: fp32 workspace [B,Hq,L]; dq/dk/dv: bf16 token-major outputs (dk/dv have Hkv heads)
extern "C" int dalm_b200_attention_bwd(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                                       long long ldv, const int64_t* mask, const void* out, long long ldo,
                                       const float* lse, const void* d_out, long long lddo, float* delta, void* dq,
                                       long long lddq, void* dk, long long lddk, void* dv, long long lddv, int B, int L,
                                       int Hq, int Hkv, int D, float scale, int causal, float drop_p,
                                       unsigned long long drop_seed, unsigned long long drop_stream_id,
                                       const void* drop_offset, void* stream) {
  AttnParams p{};
  p.q = (const __nv_bfloat16*)q; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.mask = mask; p.o = (__nv_bfloat16*)const_cast<void*>(out); p.ldo = ldo;
  p.lse = const_cast<float*>(lse); p.d_o = (const __nv_bfloat16*)d_out; p.lddo = lddo; p.delta = delta;
  p.dq = (__nv_bfloat16*)dq; p.dk = (__nv_bfloat16*)dk; p.dv = (__nv_bfloat16*)dv;
  p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  p.B = B; p.L = L; p.Hq = Hq; p.Hkv = Hkv; p.scale = scale; p.causal = causal;
  if (int e = check_common(p, D)) return e;
  DALM_REQUIRE(lddo % 8 == 0 && ((uintptr_t)d_out & 15) == 0, "attention_bwd: d_out alignment");
  DALM_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || D <= 64), "attention_bwd: dropout needs p in [0,1) and head_dim <= 64");
  p.drop = make_drop(drop_p, drop_seed, drop_stream_id, drop_offset);
  cudaStream_t st = (cudaStream_t)stream;
  if (drop_p > 0.f) return D == 32 ? launch_bwd<32, true>(p, st) : launch_bwd<64, true>(p, st);
  if (D == 32) return launch_bwd<32, false>(p, st);
  if (D == 64) return launch_bwd<64, false>(p, st);
  return launch_bwd<128, false>(p, st);
}
