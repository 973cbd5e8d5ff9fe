// Dense dequantisation kernels (quip_lib::decompress_*_origorder) -- bit-exact with the reference
// kernels K6-K10 (quip_cuda/origin_order.cu:794-1074), any size, write-bandwidth bound.
//
// Mapping: one thread per code, consecutive lanes take consecutive codes, so every warp store
// instruction writes 512 contiguous bytes (E8P: 16 B / code).  HBM traffic per code: 2 B read +
// 16 B written (E8P12), i.e. the kernel is bounded by the fp16 output it has to materialise.
#include "common.cuh"

namespace qb {

constexpr int DEC_THREADS = 256;
constexpr int DEC_UNROLL = 4;

__device__ __forceinline__ uint4 e8p_q_to_f16x8(uint2 v) {
  __half2 e0, o0, e1, o1;
  q4_to_half2(v.x, e0, o0);  // (w0,w1)=(b0,b2)  (w2,w3)=(b1,b3)
  q4_to_half2(v.y, e1, o1);  // (w4,w5)=(b4,b6)  (w6,w7)=(b5,b7)
  uint4 r;
  r.x = *reinterpret_cast<uint32_t*>(&e0);
  r.y = *reinterpret_cast<uint32_t*>(&o0);
  r.z = *reinterpret_cast<uint32_t*>(&e1);
  r.w = *reinterpret_cast<uint32_t*>(&o1);
  return r;
}

__device__ __forceinline__ uint4 hfma2x4(__half2 s, uint4 a, uint4 c) {
  uint4 r;
  const __half2* pa = reinterpret_cast<const __half2*>(&a);
  const __half2* pc = reinterpret_cast<const __half2*>(&c);
  __half2* pr = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; i++) pr[i] = __hfma2(s, pa[i], pc[i]);
  return r;
}

// mode 0: E8P12 (u16 codes)  1: RVQ4B (u32 codes)  3: RVQ3B (byte triplets)
template <int MODE>
__global__ void __launch_bounds__(DEC_THREADS) decompress_e8_kernel(
    const void* __restrict__ q, const uint2* __restrict__ tab, const uint32_t* __restrict__ cb2,
    uint4* __restrict__ out, int64_t ncodes, float resid_scale) {
  __shared__ uint2 tab1[256];
  __shared__ uint32_t cb2s[256];
  {
    uint2 t = tab[threadIdx.x];
    t.x |= 0x01010101u;
    t.y |= 0x01010101u;
    tab1[threadIdx.x] = t;
    if (MODE == 3) cb2s[threadIdx.x] = cb2[threadIdx.x];
  }
  __syncthreads();
  const __half2 s2 = __float2half2_rn(resid_scale);
  const int64_t stride = (int64_t)gridDim.x * DEC_THREADS;
  int64_t i0 = (int64_t)blockIdx.x * DEC_THREADS + threadIdx.x;
  for (; i0 < ncodes; i0 += stride * DEC_UNROLL) {
    uint32_t code[DEC_UNROLL];
    uint32_t rem[DEC_UNROLL];
#pragma unroll
    for (int u = 0; u < DEC_UNROLL; u++) {
      const int64_t i = i0 + u * stride;
      code[u] = 0;
      rem[u] = 0;
      if (i < ncodes) {
        if (MODE == 0) {
          code[u] = reinterpret_cast<const uint16_t*>(q)[i];
        } else if (MODE == 1) {
          const uint32_t w = reinterpret_cast<const uint32_t*>(q)[i];
          code[u] = w >> 16;
          rem[u] = w & 0xffffu;
        } else {
          const uint8_t* p = reinterpret_cast<const uint8_t*>(q) + 3 * i;
          rem[u] = p[0];
          code[u] = (uint32_t)p[1] | ((uint32_t)p[2] << 8);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < DEC_UNROLL; u++) {
      const int64_t i = i0 + u * stride;
      if (i >= ncodes) continue;
      uint4 w = e8p_q_to_f16x8(e8p_decode_q(tab1[code[u] >> 8], code[u]));
      if (MODE == 1) {
        const uint4 r = e8p_q_to_f16x8(e8p_decode_q(tab1[rem[u] >> 8], rem[u]));
        w = hfma2x4(s2, r, w);
      } else if (MODE == 3) {
        // residual: 8 nibbles of 2*v (two's complement), nibble j+4*h <-> element 2j+h
        const uint32_t c = cb2s[rem[u]];
        const __half2 adj = __float2half2_rn(-516.0f);
        uint4 r;
        uint32_t* pr = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          uint32_t b = ((c >> (4 * j)) & 0x000f000fu) ^ 0x60086008u;  // 512 + (nib^8)/2
          __half2 h = __hadd2(*reinterpret_cast<__half2*>(&b), adj);
          pr[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        w = hfma2x4(s2, r, w);
      }
      out[i] = w;
    }
  }
}

__global__ void __launch_bounds__(DEC_THREADS) decompress_d4_kernel(
    const uint8_t* __restrict__ q, const uint2* __restrict__ cb, uint2* __restrict__ out, int64_t ncodes) {
  __shared__ uint2 cbs[256];
  cbs[threadIdx.x] = cb[threadIdx.x];
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * DEC_THREADS;
  for (int64_t i = (int64_t)blockIdx.x * DEC_THREADS + threadIdx.x; i < ncodes; i += stride)
    out[i] = cbs[q[i]];
}

__global__ void __launch_bounds__(DEC_THREADS) decompress_hi_kernel(
    const uint32_t* __restrict__ q, uint4* __restrict__ out, int64_t ncodes) {
  const int64_t stride = (int64_t)gridDim.x * DEC_THREADS;
  const uint32_t c0 = 0x64086408u;  // 1024 + 8 (+16*nibble)
  const __half2 y16 = __float2half2_rn(1.0f / 16.0f);
  const __half2 z16 = __float2half2_rn(-1024.0f / 16.0f - 8.0f);
  for (int64_t i = (int64_t)blockIdx.x * DEC_THREADS + threadIdx.x; i < ncodes; i += stride) {
    uint32_t qa = q[i];
    uint32_t w[4];
    w[0] = ((qa & 0x000f000fu) << 4) | c0;
    w[1] = (qa & 0x00f000f0u) | c0;
    qa >>= 8;
    w[2] = ((qa & 0x000f000fu) << 4) | c0;
    w[3] = (qa & 0x00f000f0u) | c0;
    uint4 r;
    uint32_t* pr = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      __half2 h = __hfma2(*reinterpret_cast<__half2*>(&w[j]), y16, z16);
      pr[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    out[i] = r;
  }
}

static int dec_grid(int64_t ncodes, int per_thread) {
  int64_t blocks = (ncodes + (int64_t)DEC_THREADS * per_thread - 1) / ((int64_t)DEC_THREADS * per_thread);
  const int64_t cap = (int64_t)quipb200_sm_count() * 8 * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace qb

using namespace qb;

extern "C" int quipb200_decompress_e8p(const int16_t* qidxs, const int64_t* grid, void* out, int64_t rows,
                                       int64_t cpr, void* stream) {
  if (!qidxs || !grid || !out || rows < 0 || cpr < 0) return QUIPB200_EINVAL;
  if (!aligned16(out) || !aligned16(grid)) return QUIPB200_EALIGN;
  const int64_t n = rows * cpr;
  if (n == 0) return 0;
  decompress_e8_kernel<0><<<dec_grid(n, DEC_UNROLL), DEC_THREADS, 0, (cudaStream_t)stream>>>(
      qidxs, reinterpret_cast<const uint2*>(grid), nullptr, reinterpret_cast<uint4*>(out), n, 0.f);
  QB_LAUNCH_CHECK();
  return 0;
}

extern "C" int quipb200_decompress_e8prvq4(const int32_t* qidxs, const int64_t* grid, void* out, int64_t rows,
                                           int64_t cpr, float resid_scale, void* stream) {
  if (!qidxs || !grid || !out || rows < 0 || cpr < 0) return QUIPB200_EINVAL;
  if (!aligned16(out) || !aligned16(grid)) return QUIPB200_EALIGN;
  const int64_t n = rows * cpr;
  if (n == 0) return 0;
  decompress_e8_kernel<1><<<dec_grid(n, DEC_UNROLL), DEC_THREADS, 0, (cudaStream_t)stream>>>(
      qidxs, reinterpret_cast<const uint2*>(grid), nullptr, reinterpret_cast<uint4*>(out), n, resid_scale);
  QB_LAUNCH_CHECK();
  return 0;
}

extern "C" int quipb200_decompress_e8prvq3(const int32_t* qidxs, const int64_t* grid, const int32_t* e81b,
                                           void* out, int64_t rows, int64_t cpr, float resid_scale,
                                           void* stream) {
  if (!qidxs || !grid || !e81b || !out || rows < 0 || cpr < 0) return QUIPB200_EINVAL;
  if (!aligned16(out) || !aligned16(grid)) return QUIPB200_EALIGN;
  const int64_t n = rows * cpr;
  if (n == 0) return 0;
  decompress_e8_kernel<3><<<dec_grid(n, DEC_UNROLL), DEC_THREADS, 0, (cudaStream_t)stream>>>(
      qidxs, reinterpret_cast<const uint2*>(grid), reinterpret_cast<const uint32_t*>(e81b),
      reinterpret_cast<uint4*>(out), n, resid_scale);
  QB_LAUNCH_CHECK();
  return 0;
}

extern "C" int quipb200_decompress_d4(const uint8_t* qidxs, const void* grid, void* out, int64_t rows,
                                      int64_t cpr, void* stream) {
  if (!qidxs || !grid || !out || rows < 0 || cpr < 0) return QUIPB200_EINVAL;
  if (!aligned16(out) || !aligned16(grid)) return QUIPB200_EALIGN;
  const int64_t n = rows * cpr;
  if (n == 0) return 0;
  decompress_d4_kernel<<<dec_grid(n, 4), DEC_THREADS, 0, (cudaStream_t)stream>>>(
      qidxs, reinterpret_cast<const uint2*>(grid), reinterpret_cast<uint2*>(out), n);
  QB_LAUNCH_CHECK();
  return 0;
}

extern "C" int quipb200_decompress_hi(const int32_t* qidxs, void* out, int64_t rows, int64_t cpr,
                                      void* stream) {
  if (!qidxs || !out || rows < 0 || cpr < 0) return QUIPB200_EINVAL;
  if (!aligned16(out)) return QUIPB200_EALIGN;
  const int64_t n = rows * cpr;
  if (n == 0) return 0;
  decompress_hi_kernel<<<dec_grid(n, 4), DEC_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint32_t*>(qidxs), reinterpret_cast<uint4*>(out), n);
  QB_LAUNCH_CHECK();
  return 0;
}
