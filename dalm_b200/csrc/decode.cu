// dalm_b200 — autoregressive greedy decoding for the evaluation path (reference dalm/eval/eval_rag.py:126-140:
// `model.generate(**inputs, max_length=max_length, early_stopping=True)` on the generator, then exact match :268-277).
//
//   rope_pos_kernel      : RoPE at explicit per-token position ids (HF generate derives them from the attention mask:
//                          cumsum(mask) - 1, so left / right padded prompts rotate differently from arange)
//   attn_decode_kernel   : one query token per sequence against the KV cache. HBM-bound: reads the cached K and V of
//                          one (sequence, kv head) once = 2 * T * D * 2 bytes; the current token's K / V rows are
//                          appended to the cache by the same launch (no separate copy kernel)
//   greedy_step_kernel   : argmax over the vocabulary row + HF's finished-sequence bookkeeping (pad after EOS), writes
//                          the token, its attention-mask bit and its position id for the next step
//
// All state a step needs (next token ids, position ids, finished flags, per-step alive counts, and — in device-column mode —
// each row's current column) lives in device memory: a decode step is then the SAME launch sequence with the same
// arguments for every token and can be captured once as a CUDA graph and replayed (292 launches per token at Llama-2-7B).
// The host reads back one "is anyone still generating" counter every few steps.
// Measured (profiles/README.md, r01_decode_bench.jsonl): a decode step costs ~3.3 ms + 0.029 ms x cached tokens; the second
// term is attn_decode_kernel's PV pass (pass 3 below walks the keys serially per thread, one dependent 2-byte load each):
// the known limiter of this file, first item of the next round.
#include "common.cuh"
#include <limits.h>
#include <stdlib.h>

namespace dalm {

__global__ void rope_pos_kernel(__nv_bfloat16* __restrict__ buf, long long ld, int col0, int nheads, int D,
                                const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                const int64_t* __restrict__ pos, int T) {
  // one CTA per token row; HF rotate_half convention, same arithmetic as rope_kernel (rowwise.cu)
  const size_t r = blockIdx.x;
  long long p = pos[r];
  p = p < 0 ? 0 : (p >= T ? T - 1 : p);
  const int half = D / 2;
  __nv_bfloat16* base = buf + r * ld + col0;
  for (int i = threadIdx.x; i < nheads * half; i += blockDim.x) {
    const int h = i / half, j = i - h * half;
    __nv_bfloat16* q = base + h * D + j;
    const float x1 = __bfloat162float(q[0]), x2 = __bfloat162float(q[half]);
    const float c = cos_t[(size_t)p * half + j], s = sin_t[(size_t)p * half + j];
    q[0] = __float2bfloat16(x1 * c - x2 * s);
    q[half] = __float2bfloat16(x2 * c + x1 * s);
  }
}

// grid (Hq, B), 128 threads. Shared memory: q[D] | p[cur+1] | red[32] | part[128] floats.
//   pass 1: thread-per-key scores (each thread reads whole K rows with 16-byte loads, q broadcast from smem)
//   pass 2: block max / exp / sum
//   pass 3: thread-per-output-column PV (128 / D thread groups split the keys, combined through smem)
template <int D>
__global__ void __launch_bounds__(128) attn_decode_kernel(const __nv_bfloat16* __restrict__ qkv, long long ldq, int q_col,
                                                          int k_col, int v_col, __nv_bfloat16* __restrict__ cache_k,
                                                          __nv_bfloat16* __restrict__ cache_v, long long cache_sb,
                                                          long long cache_st, const int64_t* __restrict__ mask,
                                                          long long ldm, __nv_bfloat16* __restrict__ out, long long ldo,
                                                          int Hq, int Hkv, int cur_host, const int* __restrict__ cur_dev,
                                                          int sp_cap, float scale) {
  extern __shared__ float sm[];
  float* sq = sm;
  float* sp = sm + D;
  float* red = sp + sp_cap;                      // sp_cap >= cur + 1, a multiple of 4 (host-side: the cache length in
  float* part = red + 32;                        // device-column mode, so one captured launch serves every step)
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int cur = cur_dev ? min(cur_dev[b], sp_cap - 1) : cur_host;
  const int group = Hq / Hkv, kvh = h / group;
  const __nv_bfloat16* qrow = qkv + (size_t)b * ldq + q_col + h * D;
  const __nv_bfloat16* krow = qkv + (size_t)b * ldq + k_col + kvh * D;
  const __nv_bfloat16* vrow = qkv + (size_t)b * ldq + v_col + kvh * D;
  __nv_bfloat16* ck = cache_k + (size_t)b * cache_sb + kvh * D;
  __nv_bfloat16* cv = cache_v + (size_t)b * cache_sb + kvh * D;
  if (tid < D) {
    sq[tid] = __bfloat162float(qrow[tid]) * scale;
    if (h % group == 0) {                       // the first query head of each kv group appends this token's K / V
      ck[(size_t)cur * cache_st + tid] = krow[tid];
      cv[(size_t)cur * cache_st + tid] = vrow[tid];
    }
  }
  __syncthreads();

  float lmax = -INFINITY;
  for (int t = tid; t <= cur; t += 128) {
    // column `cur` is the token being decoded: always visible, read from the qkv row (its cache slot is written above
    // by ANOTHER CTA of this launch, so nobody reads it back here)
    const bool valid = (t == cur) || mask[(size_t)b * ldm + t] != 0;
    float s = -INFINITY;
    if (valid) {
      const __nv_bfloat16* kp = (t == cur) ? krow : ck + (size_t)t * cache_st;
      s = 0.f;
#pragma unroll
      for (int j = 0; j < D; j += 8) {
        float f[8];
        unpack8(*reinterpret_cast<const bf16x8*>(kp + j), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) s += sq[j + e] * f[e];
      }
    }
    sp[t] = s;
    lmax = fmaxf(lmax, s);
  }
  const float m = block_max(lmax, red);         // finite: column `cur` is always valid
  float lsum = 0.f;
  for (int t = tid; t <= cur; t += 128) {
    const float s = sp[t];
    const float p = (s == -INFINITY) ? 0.f : __expf(s - m);
    sp[t] = p;
    lsum += p;
  }
  const float sum = block_sum(lsum, red);
  __syncthreads();

  constexpr int G = 128 / D;                    // thread groups splitting the keys: 1 (D=128), 2 (64), 4 (32)
  const int d = tid % D, g = tid / D;
  float acc = 0.f;
  for (int t = g; t <= cur; t += G) {
    const float p = sp[t];
    if (p != 0.f) {
      const __nv_bfloat16* vp = (t == cur) ? vrow : cv + (size_t)t * cache_st;
      acc += p * __bfloat162float(vp[d]);
    }
  }
  if (G > 1) {
    part[tid] = acc;
    __syncthreads();
    if (g == 0)
      for (int gg = 1; gg < G; ++gg) acc += part[gg * D + d];
  }
  if (g == 0) out[(size_t)b * ldo + h * D + d] = __float2bfloat16(acc / sum);
}

// DEFAULT decode attention: the kernel above with a parallel PV pass. Measured on the first kernel: 0.9 us per cached key per
// layer, because its pass 3 walks the keys serially per thread with the V load behind `if (p != 0)` (profiles/README.md,
// r01_decode_bench.jsonl line 4). Here 128 threads = KG key groups x D/8 lanes, every lane loads 16 bytes of V unconditionally
// (masked keys carry p = 0; their cache rows are initialised memory) and the KG partial rows meet in shared memory.
// Measured (profiles/r02b_decode_bench.jsonl, Llama-2-7B shape, 16 x 192 -> 256 tokens): 13.3 / 9.1 ms (graph / eager) per token
// step with the first kernel -> 4.55 / 5.0 ms = 3 514 tokens/s = 0.50 of the HBM floor. DALM_B200_DECODE_ATTN=1 selects the first
// kernel (kept as the cross-check of tests/test_generate_gpu.py).
// one explicit 16-byte read-only load -> 8 floats (the struct-typed loads above are split into 32-bit loads by the compiler)
__device__ __forceinline__ void load8_nc(const __nv_bfloat16* p, float* f) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}

template <int D>
__global__ void __launch_bounds__(128) attn_decode_v2_kernel(const __nv_bfloat16* __restrict__ qkv, long long ldq, int q_col,
                                                             int k_col, int v_col, __nv_bfloat16* __restrict__ cache_k,
                                                             __nv_bfloat16* __restrict__ cache_v, long long cache_sb,
                                                             long long cache_st, const int64_t* __restrict__ mask,
                                                             long long ldm, __nv_bfloat16* __restrict__ out, long long ldo,
                                                             int Hq, int Hkv, int cur_host, const int* __restrict__ cur_dev,
                                                             int sp_cap, float scale) {
  extern __shared__ float sm[];
  float* sq = sm;
  float* sp = sm + D;
  float* red = sp + sp_cap;
  float* part = red + 32;                        // [KG][D] = 1024 floats for every supported D
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int cur = cur_dev ? min(cur_dev[b], sp_cap - 1) : cur_host;
  const int group = Hq / Hkv, kvh = h / group;
  const __nv_bfloat16* qrow = qkv + (size_t)b * ldq + q_col + h * D;
  const __nv_bfloat16* krow = qkv + (size_t)b * ldq + k_col + kvh * D;
  const __nv_bfloat16* vrow = qkv + (size_t)b * ldq + v_col + kvh * D;
  __nv_bfloat16* ck = cache_k + (size_t)b * cache_sb + kvh * D;
  __nv_bfloat16* cv = cache_v + (size_t)b * cache_sb + kvh * D;
  if (tid < D) {
    sq[tid] = __bfloat162float(qrow[tid]) * scale;
    if (h % group == 0) {
      ck[(size_t)cur * cache_st + tid] = krow[tid];
      cv[(size_t)cur * cache_st + tid] = vrow[tid];
    }
  }
  __syncthreads();

  float lmax = -INFINITY;
  for (int t = tid; t <= cur; t += 128) {
    const bool valid = (t == cur) || mask[(size_t)b * ldm + t] != 0;
    const __nv_bfloat16* kp = (t == cur) ? krow : ck + (size_t)t * cache_st;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < D; j += 8) {
      float f[8];
      load8_nc(kp + j, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += sq[j + e] * f[e];
    }
    s = valid ? s : -INFINITY;
    sp[t] = s;
    lmax = fmaxf(lmax, s);
  }
  const float m = block_max(lmax, red);
  float lsum = 0.f;
  for (int t = tid; t <= cur; t += 128) {
    const float s = sp[t];
    const float p = (s == -INFINITY) ? 0.f : __expf(s - m);
    sp[t] = p;
    lsum += p;
  }
  const float sum = block_sum(lsum, red);
  __syncthreads();

  constexpr int LPK = D / 8, KG = 128 / LPK;    // lanes per key (16 bytes each), key groups: 16 x 8, 8 x 16, 4 x 32
  const int kg = tid / LPK, dl = tid % LPK;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll 4
  for (int t = kg; t <= cur; t += KG) {
    const __nv_bfloat16* vp = (t == cur) ? vrow : cv + (size_t)t * cache_st;
    float f[8];
    load8_nc(vp + 8 * dl, f);
    const float p = sp[t];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += p * f[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[kg * D + 8 * dl + e] = acc[e];
  __syncthreads();
  if (tid < D) {
    float o = 0.f;
#pragma unroll
    for (int k = 0; k < KG; ++k) o += part[k * D + tid];
    out[(size_t)b * ldo + h * D + tid] = __float2bfloat16(o / sum);
  }
}

// grid B, 256 threads. argmax over logits[b, 0..V) (ties -> lowest index, like torch.argmax), then HF's greedy bookkeeping
// (transformers generation/utils.py _sample, do_sample=False): finished rows emit pad, a row finishes when it emits an EOS id.
__global__ void __launch_bounds__(256) greedy_step_kernel(const __nv_bfloat16* __restrict__ logits, long long ld, int V,
                                                          const int64_t* __restrict__ eos_ids, int n_eos, long long pad_id,
                                                          int* __restrict__ unfinished, int64_t* __restrict__ tokens,
                                                          long long ldt, int64_t* __restrict__ mask, long long ldm,
                                                          int col_host, int* __restrict__ cur_dev, int T,
                                                          int64_t* __restrict__ next_ids, int64_t* __restrict__ pos,
                                                          int* __restrict__ alive) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // device-column mode (CUDA-graph replays): this row's new token goes one column after the token just decoded, and the
  // row's own counter advances — no cross-CTA state, so no ordering between the CTAs of this launch is needed
  const int col = cur_dev ? cur_dev[b] + 1 : col_host;
  if (col >= T) return;                          // a replay past the end of the buffers is a no-op
  const __nv_bfloat16* row = logits + (size_t)b * ld;
  float best = -INFINITY;
  int bi = INT_MAX;
  for (int i = tid; i < V; i += 256) {
    const float v = __bfloat162float(row[i]);
    if (bi == INT_MAX || v > best) { best = v; bi = i; }      // ascending i: strict > keeps the lowest index
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi != INT_MAX && (bi == INT_MAX || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
  }
  __shared__ float sv[8];
  __shared__ int si[8];
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) {
      const float ov = sv[w];
      const int oi = si[w];
      if (oi != INT_MAX && (bi == INT_MAX || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
    }
    int unf = unfinished[b];
    const long long tok = unf ? (long long)bi : pad_id;
    tokens[(size_t)b * ldt + col] = tok;
    mask[(size_t)b * ldm + col] = 1;             // HF appends ones to the attention mask for every generated column
    next_ids[b] = tok;
    pos[b] += 1;                                 // position id of the new token = cumsum(mask) - 1
    if (unf)
      for (int e = 0; e < n_eos; ++e)
        if (tok == eos_ids[e]) unf = 0;
    unfinished[b] = unf;
    if (unf) atomicAdd(alive + col, 1);
    if (cur_dev) cur_dev[b] = col;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Weight-streaming GEMM for the decode step: out[m, n] = act(sum_k A[m,k] W[n,k]) + resid[m,n] with M <= 16 token rows.
// HBM-bound: every weight is read exactly once (N*K*2 bytes), the activations (16 x K bf16) stay in L2. The 128-row tcgen05
// tile of the training GEMM wastes 7/8 of its A tile here and runs N = 4096 on 64 CTAs (measured 10.2 ms per token at
// Llama-2-7B against a 2.3 ms HBM floor), so this path uses one `mma.sync.m16n8k16` row tile = the whole batch instead:
//   CTA = 16 output columns (2 n-tiles) x all of K, 256 threads; the 8 warps interleave over 32-wide k-chunks (split-K inside
//   the CTA), partial sums meet in shared memory. Each lane loads 16 bytes (8 consecutive k) of one weight row and of two
//   activation rows per chunk straight from global memory into MMA fragments: lane (g, t) takes k = chunk*32 + 8t .. +7, and
//   the SAME permutation of k inside the chunk is applied to A and W (a contraction does not care about the order), so no
//   cross-lane exchange or ldmatrix is needed. The pieces travel through a per-lane cp.async ring (6 chunks deep).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// 16-byte asynchronous global -> shared copy (zero-fill when !ok). L1-allocating for the activations (every CTA of an SM
// reads the same 16 x K block), L2-only for the weight stream.
__device__ __forceinline__ void cp_async16_ca(void* smem, const void* gmem, bool ok) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(ok ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async16_cg(void* smem, const void* gmem, bool ok) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(ok ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int DG_NT = 2;                 // 8-column n-tiles per CTA
constexpr int DG_STAGES = 6;             // k-chunks in flight per lane (each: 2 activation + DG_NT weight 16-byte pieces)
constexpr int DG_SLOTS = 2 + DG_NT;
constexpr int DG_SMEM = DG_STAGES * 8 * DG_SLOTS * 32 * 16;     // 96 KB: [stage][warp][slot][lane] of 16 bytes

// Every lane prefetches ITS OWN fragment pieces through a private ring in shared memory (it reads back exactly the 16-byte
// slots it copied), so the ring needs no barrier at all: `cp.async.wait_group` orders a thread's own copies. Plain register
// loads were tried first: ptxas sank each load next to its MMA (3 loads in flight per lane instead of 16).
__global__ void __launch_bounds__(256) decode_gemm_kernel(const __nv_bfloat16* __restrict__ A, long long lda,
                                                          const __nv_bfloat16* __restrict__ W, long long ldw,
                                                          void* __restrict__ out, long long ldo, int out_f32,
                                                          const void* __restrict__ resid, long long ldr, int resid_f32, int act,
                                                          int M, int N, int K) {
  extern __shared__ __align__(16) unsigned char dg_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * (DG_NT * 8);
  const int nchunks = (K + 31) / 32;
  const int my_chunks = nchunks > warp ? (nchunks - warp + 7) / 8 : 0;    // this warp takes chunks warp, warp + 8, ...
  uint4* ring = reinterpret_cast<uint4*>(dg_smem) + (size_t)warp * DG_SLOTS * 32 + lane;   // + stage*8*SLOTS*32 + slot*32
  constexpr int STAGE_STRIDE = 8 * DG_SLOTS * 32;
  float acc[DG_NT][4];
#pragma unroll
  for (int i = 0; i < DG_NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  const bool row_lo = g < M, row_hi = g + 8 < M;
  const __nv_bfloat16* a_lo_p = A + (size_t)(row_lo ? g : 0) * lda;
  const __nv_bfloat16* a_hi_p = A + (size_t)(row_hi ? g + 8 : 0) * lda;
  const __nv_bfloat16* w_p[DG_NT];
  bool w_ok[DG_NT];
#pragma unroll
  for (int i = 0; i < DG_NT; ++i) {
    const int n = n0 + i * 8 + g;
    w_ok[i] = n < N;
    w_p[i] = W + (size_t)(w_ok[i] ? n : 0) * ldw;
  }
  auto issue = [&](int i) {                              // i-th chunk of this warp -> ring stage i % DG_STAGES
    if (i < my_chunks) {
      const int k = (warp + 8 * i) * 32 + 8 * t;         // K % 8 == 0: a lane's 8 elements are all inside or all outside K
      const bool kin = k < K;
      const int ks = kin ? k : 0;
      uint4* st = ring + (size_t)(i % DG_STAGES) * STAGE_STRIDE;
      cp_async16_ca(st, a_lo_p + ks, kin && row_lo);
      cp_async16_ca(st + 32, a_hi_p + ks, kin && row_hi);
#pragma unroll
      for (int j = 0; j < DG_NT; ++j) cp_async16_cg(st + (2 + j) * 32, w_p[j] + ks, kin && w_ok[j]);
    }
    cp_async_commit_group();                             // one group per call, empty or not: the wait below counts groups
  };
#pragma unroll
  for (int i = 0; i < DG_STAGES - 1; ++i) issue(i);
  for (int i = 0; i < my_chunks; ++i) {
    issue(i + DG_STAGES - 1);                            // refills the stage this lane consumed in the previous iteration
    cp_async_wait_group<DG_STAGES - 1>();                // chunk i has landed
    const uint4* st = ring + (size_t)(i % DG_STAGES) * STAGE_STRIDE;
    const uint4 alo = st[0], ahi = st[32];
#pragma unroll
    for (int j = 0; j < DG_NT; ++j) {
      const uint4 b = st[(2 + j) * 32];
      // fragment k-slots (2t,2t+1) and (2t+8,2t+9) are fed actual k (8t+0,1),(8t+2,3), then (8t+4,5),(8t+6,7), for A and W alike
      mma16816(acc[j], alo.x, ahi.x, alo.y, ahi.y, b.x, b.y);
      mma16816(acc[j], alo.z, ahi.z, alo.w, ahi.w, b.z, b.w);
    }
  }
  cp_async_wait_group<0>();
  __syncthreads();                                       // every warp is done with its ring: reuse the memory for the partials
  float(*part)[16][DG_NT * 8 + 1] = reinterpret_cast<float(*)[16][DG_NT * 8 + 1]>(dg_smem);
  // accumulator fragment: c0,c1 = (row g, cols 2t, 2t+1), c2,c3 = (row g+8, same cols)
#pragma unroll
  for (int i = 0; i < DG_NT; ++i) {
    part[warp][g][i * 8 + 2 * t] = acc[i][0];
    part[warp][g][i * 8 + 2 * t + 1] = acc[i][1];
    part[warp][g + 8][i * 8 + 2 * t] = acc[i][2];
    part[warp][g + 8][i * 8 + 2 * t + 1] = acc[i][3];
  }
  __syncthreads();
  const int m = tid >> 4, nn = tid & 15, n = n0 + nn;   // 256 threads = 16 rows x 16 columns
  if (m < M && n < N) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += part[w][m][nn];
    if (act == 1) v = gelu_erf(v);
    if (resid) v += resid_f32 ? static_cast<const float*>(resid)[(size_t)m * ldr + n]
                              : __bfloat162float(static_cast<const __nv_bfloat16*>(resid)[(size_t)m * ldr + n]);
    if (out_f32) static_cast<float*>(out)[(size_t)m * ldo + n] = v;
    else static_cast<__nv_bfloat16*>(out)[(size_t)m * ldo + n] = __float2bfloat16(v);
  }
}

}  // namespace dalm

using namespace dalm;
#define ST(s) ((cudaStream_t)(s))

extern "C" int dalm_b200_decode_gemm(const void* A, long long lda, const void* W, long long ldw, void* out, long long ldo,
                                     int out_f32, const void* resid, long long ldr, int resid_f32, int act, int M, int N, int K,
                                     void* stream) {
  DALM_REQUIRE(M > 0 && M <= 16 && N > 0 && K > 0, "decode_gemm: needs 1 <= M <= 16 token rows (got M=%d N=%d K=%d)", M, N, K);
  DALM_REQUIRE((K % 8) == 0 && (lda % 8) == 0 && (ldw % 8) == 0 && lda >= K && ldw >= K && ldo >= N,
               "decode_gemm: K and the operand row strides must be multiples of 8 elements");
  DALM_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0, "decode_gemm: operands must be 16-byte aligned");
  DALM_REQUIRE(act == 0 || act == 1, "decode_gemm: act must be 0 (none) or 1 (gelu)");
  DALM_REQUIRE(resid == nullptr || ldr >= N, "decode_gemm: residual row stride");
  static bool attr = false;
  if (!attr) { DALM_CUDA(cudaFuncSetAttribute(decode_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_SMEM)); attr = true; }
  decode_gemm_kernel<<<(N + DG_NT * 8 - 1) / (DG_NT * 8), 256, DG_SMEM, ST(stream)>>>(
      (const __nv_bfloat16*)A, lda, (const __nv_bfloat16*)W, ldw, out, ldo, out_f32, resid, ldr, resid_f32, act, M, N, K);
  count_launch();
  return check_launch("decode_gemm_kernel");
}

extern "C" int dalm_b200_rope_pos(void* buf, long long ld, int col0, int nheads, int D, const float* cos_t,
                                  const float* sin_t, const int64_t* pos, int M, int T, void* stream) {
  DALM_REQUIRE(M > 0 && T > 0 && nheads > 0 && (D % 2) == 0, "rope_pos: bad shape M=%d T=%d heads=%d D=%d", M, T, nheads, D);
  DALM_REQUIRE(pos != nullptr && cos_t != nullptr && sin_t != nullptr, "rope_pos: null table / positions");
  rope_pos_kernel<<<M, 256, 0, ST(stream)>>>((__nv_bfloat16*)buf, ld, col0, nheads, D, cos_t, sin_t, pos, T);
  count_launch();
  return check_launch("rope_pos_kernel");
}

extern "C" int dalm_b200_attention_decode(const void* qkv, long long ldq, int q_col, int k_col, int v_col, void* cache_k,
                                          void* cache_v, long long cache_sb, long long cache_st, const int64_t* mask,
                                          long long ldm, void* out, long long ldo, int B, int Hq, int Hkv, int D, int cur,
                                          const int* cur_dev, int T, float scale, void* stream) {
  DALM_REQUIRE(B > 0 && Hq > 0 && Hkv > 0 && (Hq % Hkv) == 0, "attention_decode: bad heads B=%d Hq=%d Hkv=%d", B, Hq, Hkv);
  DALM_REQUIRE(D == 32 || D == 64 || D == 128, "attention_decode: head_dim %d unsupported (32/64/128)", D);
  DALM_REQUIRE(T > 0 && T <= 8192 && (cur_dev != nullptr || (cur >= 0 && cur < T)),
               "attention_decode: column %d outside the cache of %d tokens (max 8192)", cur, T);
  DALM_REQUIRE((ldq % 8) == 0 && (q_col % 8) == 0 && (k_col % 8) == 0 && (v_col % 8) == 0 && (cache_st % 8) == 0 &&
                   (cache_sb % 8) == 0, "attention_decode: rows must be 16-byte aligned");
  DALM_REQUIRE(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)cache_k & 15) == 0 && ((uintptr_t)cache_v & 15) == 0,
               "attention_decode: pointers must be 16-byte aligned");
  DALM_REQUIRE(mask != nullptr && cache_st >= (long long)Hkv * D && cache_sb >= cache_st * T, "attention_decode: cache layout");
  const int sp_cap = ((cur_dev ? T : cur + 1) + 3) & ~3;
  const char* variant = getenv("DALM_B200_DECODE_ATTN");
  const bool v2 = !(variant != nullptr && variant[0] == '1');       // default: parallel-PV kernel; 1 = the first (serial-PV) kernel
  const size_t smem = (size_t)(D + sp_cap + 32 + (v2 ? 1024 : 128)) * sizeof(float);
  dim3 grid(Hq, B);
  if (v2) {
#define DALM_DECODE2(DD)                                                                                                 \
  attn_decode_v2_kernel<DD><<<grid, 128, smem, ST(stream)>>>((const __nv_bfloat16*)qkv, ldq, q_col, k_col, v_col,        \
                                                             (__nv_bfloat16*)cache_k, (__nv_bfloat16*)cache_v, cache_sb, \
                                                             cache_st, mask, ldm, (__nv_bfloat16*)out, ldo, Hq, Hkv, cur, \
                                                             cur_dev, sp_cap, scale)
    if (D == 128) DALM_DECODE2(128); else if (D == 64) DALM_DECODE2(64); else DALM_DECODE2(32);
#undef DALM_DECODE2
    count_launch();
    return check_launch("attn_decode_v2_kernel");
  }
#define DALM_DECODE(DD)                                                                                                  \
  attn_decode_kernel<DD><<<grid, 128, smem, ST(stream)>>>((const __nv_bfloat16*)qkv, ldq, q_col, k_col, v_col,           \
                                                          (__nv_bfloat16*)cache_k, (__nv_bfloat16*)cache_v, cache_sb,    \
                                                          cache_st, mask, ldm, (__nv_bfloat16*)out, ldo, Hq, Hkv, cur,     \
                                                          cur_dev, sp_cap, scale)
  if (D == 128) DALM_DECODE(128); else if (D == 64) DALM_DECODE(64); else DALM_DECODE(32);
#undef DALM_DECODE
  count_launch();
  return check_launch("attn_decode_kernel");
}

extern "C" int dalm_b200_greedy_step(const void* logits, long long ld, int B, int V, const int64_t* eos_ids, int n_eos,
                                     long long pad_id, int* unfinished, int64_t* tokens, long long ldt, int64_t* mask,
                                     long long ldm, int col, int* cur_dev, int T, int64_t* next_ids, int64_t* pos, int* alive,
                                     void* stream) {
  DALM_REQUIRE(B > 0 && V > 0 && T > 0 && T <= ldt && T <= ldm, "greedy_step: bad shape B=%d V=%d T=%d", B, V, T);
  DALM_REQUIRE(cur_dev != nullptr || (col >= 0 && col < T), "greedy_step: column %d outside the %d-token buffers", col, T);
  DALM_REQUIRE(n_eos >= 0 && (n_eos == 0 || eos_ids != nullptr), "greedy_step: eos list");
  DALM_REQUIRE(unfinished && tokens && mask && next_ids && pos && alive, "greedy_step: null state pointer");
  greedy_step_kernel<<<B, 256, 0, ST(stream)>>>((const __nv_bfloat16*)logits, ld, V, eos_ids, n_eos, pad_id, unfinished, tokens,
                                                ldt, mask, ldm, col, cur_dev, T, next_ids, pos, alive);
  count_launch();
  return check_launch("greedy_step_kernel");
}
