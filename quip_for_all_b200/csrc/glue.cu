// Decode-step glue around the QuantLinear hot path (engine side, bs=1): RoPE + KV-cache append +
// single-query attention in one launch.  Not part of the reference's operator surface -- it replaces
// the ~25 small eager kernels HF's LlamaAttention issues per layer between the q/k/v and o projections
// (the reference hides those behind torch.compile CUDA graphs, example_generate.py:68-70).
//
// One CTA per query head.  K/V rows are read once with 8-byte loads (a warp covers one 256-byte row);
// scores live in shared memory; fp32 math, fp16 in/out.  The position comes from device memory so the
// launch is CUDA-graph replayable.
#include "common.cuh"

namespace qb {

constexpr int ATT_THREADS = 1024;   // latency bound: positions are spread over 32 warps / 16 PV groups
constexpr int ATT_GROUPS = ATT_THREADS / 64;

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attn_decode_kernel(
    const __half* __restrict__ q, const __half* __restrict__ k, const __half* __restrict__ v,
    __half* __restrict__ k_cache, __half* __restrict__ v_cache,   // [nkv][max_len][HD]
    const __half* __restrict__ cos_t, const __half* __restrict__ sin_t,   // [max_len][HD]
    const long long* __restrict__ pos_ptr, __half* __restrict__ out, int nh, int nkv, int max_len, float scale) {
  extern __shared__ float sc[];   // [max_len] scores / probabilities
  __shared__ float sq[HD];
  __shared__ float sk[HD];
  __shared__ float sred[ATT_THREADS / 32];
  __shared__ float sout[ATT_GROUPS][HD];
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = blockIdx.x;
  const int group = nh / nkv;
  const int kvh = h / group;
  int pos = (int)(*pos_ptr);
  if (pos >= max_len) pos = max_len - 1;
  __half* kc = k_cache + (size_t)kvh * max_len * HD;
  __half* vc = v_cache + (size_t)kvh * max_len * HD;

  // ---- RoPE on q and the new k (HF convention: x*cos + rotate_half(x)*sin), append k/v ----
  if (tid < HD) {
    const int d = tid;
    const float c = __half2float(cos_t[(size_t)pos * HD + d]);
    const float s = __half2float(sin_t[(size_t)pos * HD + d]);
    const int dr = (d < HD / 2) ? d + HD / 2 : d - HD / 2;
    const float sgn = (d < HD / 2) ? -1.f : 1.f;
    const float qv = __half2float(q[h * HD + d]), qr = sgn * __half2float(q[h * HD + dr]);
    const float kv = __half2float(k[kvh * HD + d]), kr = sgn * __half2float(k[kvh * HD + dr]);
    sq[d] = f16_round(qv * c + qr * s) * scale;
    const __half kn = __float2half_rn(kv * c + kr * s);
    sk[d] = __half2float(kn);
    if (h % group == 0) {
      kc[(size_t)pos * HD + d] = kn;
      vc[(size_t)pos * HD + d] = v[kvh * HD + d];
    }
  }
  __syncthreads();

  // ---- scores: warp w takes positions w, w+8, ...; lane covers 4 dims of the 128-wide row ----
  static_assert(HD == 128, "lane mapping assumes head_dim 128");
  const float q0 = sq[lane * 4], q1 = sq[lane * 4 + 1], q2 = sq[lane * 4 + 2], q3 = sq[lane * 4 + 3];
  float lmax = -INFINITY;
  constexpr int NW = ATT_THREADS / 32, UN = 8;   // 8 K rows in flight per warp
  for (int t0 = warp; t0 <= pos; t0 += NW * UN) {
    uint2 raw[UN];
#pragma unroll
    for (int u = 0; u < UN; u++) {
      const int t = t0 + u * NW;
      raw[u] = make_uint2(0, 0);
      if (t < pos) raw[u] = *reinterpret_cast<const uint2*>(kc + (size_t)t * HD + lane * 4);
    }
    float d[UN];
#pragma unroll
    for (int u = 0; u < UN; u++) {
      const int t = t0 + u * NW;
      if (t == pos) {
        d[u] = q0 * sk[lane * 4] + q1 * sk[lane * 4 + 1] + q2 * sk[lane * 4 + 2] + q3 * sk[lane * 4 + 3];
      } else {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw[u].x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw[u].y));
        d[u] = q0 * a.x + q1 * a.y + q2 * b.x + q3 * b.y;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < UN; u++) d[u] += __shfl_xor_sync(0xffffffffu, d[u], o);
    }
#pragma unroll
    for (int u = 0; u < UN; u++) {
      const int t = t0 + u * NW;
      if (t <= pos) {
        if (lane == 0) sc[t] = d[u];
        lmax = fmaxf(lmax, d[u]);
      }
    }
  }
  if (lane == 0) sred[warp] = lmax;
  __syncthreads();
  float mx = -INFINITY;
#pragma unroll
  for (int w = 0; w < ATT_THREADS / 32; w++) mx = fmaxf(mx, sred[w]);
  __syncthreads();
  float lsum = 0.f;
  for (int t = tid; t <= pos; t += ATT_THREADS) {
    const float p = __expf(sc[t] - mx);
    sc[t] = p;
    lsum += p;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  if (lane == 0) sred[warp] = lsum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < ATT_THREADS / 32; w++) tot += sred[w];
  const float inv = 1.0f / tot;

  // ---- out = P . V : groups of 64 threads split the positions, each thread owns 2 dims ----
  const int d2 = tid & 63, tg = tid >> 6;
  float o0 = 0.f, o1 = 0.f;
  constexpr int UV = 8;
  for (int t0 = tg; t0 <= pos; t0 += ATT_GROUPS * UV) {
    __half2 vr[UV];
#pragma unroll
    for (int u = 0; u < UV; u++) {
      const int t = t0 + u * ATT_GROUPS;
      vr[u] = __float2half2_rn(0.f);
      if (t < pos) vr[u] = *reinterpret_cast<const __half2*>(vc + (size_t)t * HD + d2 * 2);
      else if (t == pos) vr[u] = *reinterpret_cast<const __half2*>(v + kvh * HD + d2 * 2);
    }
#pragma unroll
    for (int u = 0; u < UV; u++) {
      const int t = t0 + u * ATT_GROUPS;
      if (t <= pos) {
        const float2 vv = __half22float2(vr[u]);
        const float p = sc[t];
        o0 = fmaf(p, vv.x, o0);
        o1 = fmaf(p, vv.y, o1);
      }
    }
  }
  sout[tg][d2 * 2] = o0;
  sout[tg][d2 * 2 + 1] = o1;
  __syncthreads();
  if (tid < HD) {
    float r = 0.f;
#pragma unroll
    for (int g2 = 0; g2 < ATT_GROUPS; g2++) r += sout[g2][tid];
    out[h * HD + tid] = __float2half_rn(r * inv);
  }
}

}  // namespace qb

using namespace qb;

extern "C" int quipb200_attn_decode(const void* q, const void* k, const void* v, void* k_cache, void* v_cache,
                                    const void* cos_t, const void* sin_t, const int64_t* pos, void* out,
                                    int n_heads, int n_kv_heads, int head_dim, int max_len, void* stream) {
  if (!q || !k || !v || !k_cache || !v_cache || !cos_t || !sin_t || !pos || !out) return QUIPB200_EINVAL;
  if (head_dim != 128 || n_heads < 1 || n_kv_heads < 1 || n_heads % n_kv_heads || max_len < 1)
    return QUIPB200_EUNSUPPORTED;
  const size_t smem = (size_t)max_len * sizeof(float);
  if (smem > 160 * 1024) return QUIPB200_EUNSUPPORTED;
  if (smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(attn_decode_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  float scale = 1.0f / sqrtf((float)head_dim);
  void* args[] = {&q, &k, &v, &k_cache, &v_cache, &cos_t, &sin_t, &pos, &out, &n_heads, &n_kv_heads, &max_len, &scale};
  cudaError_t e = launch_kernel((const void*)attn_decode_kernel<128>, dim3(n_heads), dim3(ATT_THREADS), args, smem,
                                (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  QB_LAUNCH_CHECK();
  return 0;
}
