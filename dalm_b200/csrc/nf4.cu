// dalm_b200 — NF4 (bitsandbytes 4-bit NormalFloat) quantise -> dequantise round trip of a weight tensor, in place.
//
// `use_bnb` in the reference (dalm/models/rag_e2e_base_model.py:136-142, retriever_only_base_model.py:85-91) loads the
// nn.Linear weights through BitsAndBytesConfig(load_in_4bit=True, bnb_4bit_quant_type="nf4",
// bnb_4bit_compute_dtype=bfloat16): weights are cast to fp16, split into blocks of 64 consecutive elements, each block
// stores absmax (fp32) and sixteen-level codes of x / absmax; every forward dequantises code * absmax back to fp16 and runs
// the matmul in bf16. The values the GEMM sees are therefore a pure function of the checkpoint — this kernel computes them
// once at load time, and the product keeps them resident as bf16 (a B200 has the HBM; the reference quantises to fit 7B
// models on smaller parts). Same numerics as bnb's forward, none of its per-step dequantisation.
#include "common.cuh"
#include <cuda_fp16.h>

namespace dalm {

__constant__ float kNF4Code[16] = {-1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f,
                                   -0.28444138169288635f, -0.18477343022823334f, -0.09105003625154495f, 0.0f,
                                   0.07958029955625534f, 0.16093020141124725f, 0.24611230194568634f, 0.33791524171829224f,
                                   0.44070982933044434f, 0.5626170039176941f, 0.7229568362236023f, 1.0f};

// nearest NF4 level; decision boundaries are the midpoints between adjacent levels, a value exactly on a boundary goes down
// (bitsandbytes' dQuantizeNF4 decision tree tests `x > boundary`)
__device__ __forceinline__ int nf4_index(float x) {
  int idx = 0;
#pragma unroll
  for (int i = 0; i < 15; ++i) idx += (x > 0.5f * (kNF4Code[i] + kNF4Code[i + 1])) ? 1 : 0;
  return idx;
}

// one warp per block of 64 elements (2 per lane)
__global__ void nf4_roundtrip_kernel(float* __restrict__ w, long long n, unsigned char* __restrict__ codes, float* __restrict__ absmax) {
  const long long blk = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long base = blk * 64;
  if (base >= n) return;
  float v[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const long long i = base + lane + 32 * j;
    v[j] = i < n ? __half2float(__float2half_rn(w[i])) : 0.f;          // the checkpoint is cast to fp16 before quantisation
  }
  float m = fmaxf(fabsf(v[0]), fabsf(v[1]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float inv = 1.0f / m;                                           // bnb multiplies by the reciprocal
  if (lane == 0 && absmax) absmax[blk] = m;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const long long i = base + lane + 32 * j;
    if (i >= n) continue;
    const int q = m > 0.f ? nf4_index(v[j] * inv) : 7;                  // an all-zero block: level 0.0
    if (codes) codes[i] = (unsigned char)q;
    w[i] = __half2float(__float2half_rn(kNF4Code[q] * m));              // dequantised to fp16, as the forward sees it
  }
}

}  // namespace dalm

using namespace dalm;

// w: fp32 [n] (a row-major weight, flattened) overwritten with its NF4 round trip. codes (uint8 [n]) and absmax
// (fp32 [ceil(n/64)]) are optional outputs for inspection / tests.
extern "C" int dalm_b200_nf4_roundtrip(float* w, long long n, void* codes, float* absmax, void* stream) {
  DALM_REQUIRE(n > 0, "nf4_roundtrip: empty tensor");
  const long long blocks64 = (n + 63) / 64;
  const long long threads = blocks64 * 32;
  nf4_roundtrip_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, n, (unsigned char*)codes, absmax);
  count_launch();
  return check_launch("nf4_roundtrip_kernel");
}
