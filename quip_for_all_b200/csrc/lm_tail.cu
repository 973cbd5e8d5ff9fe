// Tail of a bs=1 greedy decode step in ONE launch: final RMSNorm -> lm_head GEMV (fp16 weights, fp32 accumulate, fp16
// logits) -> argmax -> next token id, its embedding row (the next step's input) and the position counter.
//
// Engine glue, not a reference operator: the reference runs these as HF's LlamaRMSNorm + nn.Linear (cuBLAS) + torch.argmax
// + nn.Embedding + `input_pos += 1` (example_generate.py:29-56), ~12 small launches per token.  The lm_head matrix is the
// only large operand (vocab x hidden fp16 = 262 MB for Llama-2-7B): HBM-bound, 16 bytes per lane per request, eight
// requests in flight per thread.
//
// Rounding points follow HF: x_norm = w * fp16(x * rstd) in fp16; logits are rounded to fp16 before the comparison and ties
// go to the lowest index (torch.argmax).  The accumulation order differs from cuBLAS, so a logit may differ by one fp16
// ulp; tests compare the logits within that tolerance and the token where the top-2 gap is larger.
#include "common.cuh"

namespace qb {

constexpr int LT_THREADS = 512;
constexpr int LT_WARPS = LT_THREADS / 32;
constexpr int LT_U = 8;              // 16-byte requests in flight per lane

struct LtParams {
  const __half* h;        // [hidden] hidden state after the last decoder layer
  const __half* norm_w;   // [hidden]
  const __half* W;        // [vocab][hidden] lm_head
  const __half* emb;      // [vocab][hidden] embedding table (may be null: no gather)
  long long* tok;         // out: next token id
  __half* h_next;         // out: emb[tok] (may be null)
  long long* pos;         // in/out: position counter, incremented (may be null)
  __half* logits;         // optional out: [vocab] fp16 logits (tests)
  unsigned long long* part;   // [grid] packed (ordered fp16 logit << 32 | ~index) partial maxima
  unsigned int* ticket;       // arrival counter, self-resetting
  int hidden, vocab;
  float eps;
};

// order-preserving map of an fp16 bit pattern to an unsigned integer (larger value = larger float)
__device__ __forceinline__ uint32_t h16_ordered(uint32_t b) { return (b & 0x8000u) ? (~b & 0xffffu) : (b | 0x8000u); }

__global__ void __launch_bounds__(LT_THREADS, 1) lm_tail_kernel(const __grid_constant__ LtParams p) {
  extern __shared__ __align__(16) unsigned char lt_smem[];
  __half* xs = reinterpret_cast<__half*>(lt_smem);                 // [hidden] normalised hidden state
  __shared__ float sred[LT_WARPS];
  __shared__ unsigned long long sbest[LT_WARPS];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.hidden, noct = H >> 3;

  // ---- LlamaRMSNorm (every CTA: 8 KB from L2) ----
  float ss = 0.f;
  for (int o = tid; o < noct; o += LT_THREADS) {
    float f[8];
    unpack_h8(__ldg(reinterpret_cast<const uint4*>(p.h) + o), f);
#pragma unroll
    for (int j = 0; j < 8; j++) ss = fmaf(f[j], f[j], ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) sred[warp] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < LT_WARPS; w++) tot += sred[w];
  const float rstd = rsqrtf(tot / (float)H + p.eps);
  for (int o = tid; o < noct; o += LT_THREADS) {
    float f[8], w8[8];
    unpack_h8(__ldg(reinterpret_cast<const uint4*>(p.h) + o), f);
    unpack_h8(__ldg(reinterpret_cast<const uint4*>(p.norm_w) + o), w8);
#pragma unroll
    for (int j = 0; j < 8; j++) f[j] = f16_round(w8[j] * f16_round(f[j] * rstd));
    reinterpret_cast<uint4*>(xs)[o] = pack_h8(f);
  }
  __syncthreads();

  // ---- GEMV: one warp per row, rows strided over all warps of the grid ----
  const int gw = blockIdx.x * LT_WARPS + warp, nw = gridDim.x * LT_WARPS;
  const int iters = (noct + 31) >> 5;                      // 16-byte pieces per lane per row
  const uint64_t pol = l2_evict_first_policy();            // 262 MB pass through the L2 once: do not evict the KV cache for it
  unsigned long long best = 0ull;
  for (int row = gw; row < p.vocab; row += nw) {
    const uint4* wr = reinterpret_cast<const uint4*>(p.W + (size_t)row * H);
    float acc = 0.f;
    for (int i0 = 0; i0 < iters; i0 += LT_U) {
      uint4 wv[LT_U];
#pragma unroll
      for (int u = 0; u < LT_U; u++) {
        const int o = (i0 + u) * 32 + lane;
        wv[u] = make_uint4(0, 0, 0, 0);
        if (i0 + u < iters && o < noct) wv[u] = ldg_stream_v4(wr + o, pol);
      }
#pragma unroll
      for (int u = 0; u < LT_U; u++) {
        const int o = (i0 + u) * 32 + lane;
        if (i0 + u < iters && o < noct) {
          const uint4 xv = reinterpret_cast<const uint4*>(xs)[o];
          const uint32_t ww[4] = {wv[u].x, wv[u].y, wv[u].z, wv[u].w}, xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&ww[q]));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&xx[q]));
            acc = fmaf(a.x, b.x, acc);
            acc = fmaf(a.y, b.y, acc);
          }
        }
      }
    }
    acc = warp_sum(acc);
    const __half lg = __float2half_rn(acc);                // nn.Linear output dtype
    if (lane == 0) {
      if (p.logits) p.logits[row] = lg;
      const uint32_t bits = (uint32_t)__half_as_ushort(lg);
      const bool nan = (bits & 0x7fffu) > 0x7c00u;
      const unsigned long long key = nan ? 0ull : (((unsigned long long)h16_ordered(bits) << 32) | (uint32_t)(~(uint32_t)row));
      best = key > best ? key : best;                      // larger logit wins; equal logits: smaller row (larger ~row)
    }
  }
  if (lane == 0) sbest[warp] = best;
  __syncthreads();
  if (tid == 0) {
    unsigned long long b = 0ull;
#pragma unroll
    for (int w = 0; w < LT_WARPS; w++) b = sbest[w] > b ? sbest[w] : b;
    p.part[blockIdx.x] = b;
    __threadfence();
    const unsigned int t = atomicAdd(p.ticket, 1u);
    s_last = (t == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  // ---- last CTA: reduce the partial maxima, publish the token, gather its embedding, advance the position ----
  __threadfence();
  unsigned long long b = 0ull;
  for (int i = tid; i < (int)gridDim.x; i += LT_THREADS) {
    const unsigned long long v = __ldcg(p.part + i);
    b = v > b ? v : b;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long v = __shfl_xor_sync(0xffffffffu, b, o);
    b = v > b ? v : b;
  }
  if (lane == 0) sbest[warp] = b;
  __syncthreads();
  b = 0ull;
#pragma unroll
  for (int w = 0; w < LT_WARPS; w++) b = sbest[w] > b ? sbest[w] : b;
  const int tokid = (int)(~(uint32_t)(b & 0xffffffffull));
  if (tid == 0) {
    *p.tok = (long long)tokid;
    if (p.pos) *p.pos = *p.pos + 1;
    *p.ticket = 0u;                                         // ready for the next launch (stream-ordered)
  }
  if (p.emb && p.h_next) {
    const uint4* er = reinterpret_cast<const uint4*>(p.emb + (size_t)tokid * H);
    for (int o = tid; o < noct; o += LT_THREADS) reinterpret_cast<uint4*>(p.h_next)[o] = __ldg(er + o);
  }
}

}  // namespace qb

using namespace qb;

extern "C" size_t quipb200_lm_tail_workspace_bytes(void) {
  const int sms = quipb200_sm_count();
  return sms < 1 ? 0 : 256 + (size_t)sms * 8;
}

extern "C" int quipb200_lm_tail(const void* h, const void* norm_w, float eps, const void* lm_head, const void* emb, int hidden,
                                int vocab, int64_t* tok_out, void* h_next, int64_t* pos, void* logits_out, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (!h || !norm_w || !lm_head || !tok_out || !workspace) return QUIPB200_EINVAL;
  if (hidden < 8 || (hidden & 7) || vocab < 1) return QUIPB200_EUNSUPPORTED;
  const void* ptrs[] = {h, norm_w, lm_head, emb, h_next};
  for (const void* q : ptrs)
    if (!aligned16(q)) return QUIPB200_EALIGN;
  if ((uintptr_t)workspace & 255) return QUIPB200_EALIGN;
  const int sms = quipb200_sm_count();
  if (sms < 1) return (int)cudaErrorNoDevice;
  if (workspace_bytes < 256 + (size_t)sms * 8) return QUIPB200_EWORKSPACE;
  const size_t smem = (size_t)hidden * sizeof(__half);
  if (smem > 200 * 1024) return QUIPB200_EUNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(lm_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  LtParams p{};
  p.h = (const __half*)h; p.norm_w = (const __half*)norm_w; p.W = (const __half*)lm_head; p.emb = (const __half*)emb;
  p.tok = (long long*)tok_out; p.h_next = (__half*)h_next; p.pos = (long long*)pos; p.logits = (__half*)logits_out;
  p.ticket = (unsigned int*)workspace;                      // zero-filled once by the caller; self-resetting afterwards
  p.part = (unsigned long long*)((unsigned char*)workspace + 256);
  p.hidden = hidden; p.vocab = vocab; p.eps = eps;
  lm_tail_kernel<<<sms, LT_THREADS, smem, (cudaStream_t)stream>>>(p);
  QB_LAUNCH_CHECK();
  return 0;
}
