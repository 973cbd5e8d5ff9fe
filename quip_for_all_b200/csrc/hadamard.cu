// quip_lib::hadamard -- Walsh-Hadamard transform over the last dim (Sylvester order), fp32 internal,
// y = H_n x * scale.  Replaces the third-party fast_hadamard_transform_cuda (register_lib.py:18-20);
// semantics pinned by the reference's own butterfly quant.py:50-59.
//
// One CTA transforms `rows_per_cta` rows held in shared memory.  Rows are read and written with
// 128-bit accesses (8 fp16 values per thread); n <= 8192 uses the Stockham ping-pong passes of
// common.cuh (strided reads, contiguous 128-bit writes, last radix-8 pass in registers feeding the global
// store directly); larger n falls back to in-place padded butterflies.  HBM traffic: read + write of the
// activation tensor, once -- the kernel is bandwidth bound for the prefill-sized inputs it exists for.
#include "common.cuh"

namespace qb {

template <typename T>
struct Vec8;
template <>
struct Vec8<__half> {
  static __device__ __forceinline__ void load(const __half* p, float (&f)[8]) {
    unpack_h8(*reinterpret_cast<const uint4*>(p), f);
  }
  static __device__ __forceinline__ void store(__half* p, const float (&f)[8]) {
    *reinterpret_cast<uint4*>(p) = pack_h8(f);
  }
  static __device__ __forceinline__ float to(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from(float v) { return __float2half_rn(v); }
};
template <>
struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = v;
  }
  static __device__ __forceinline__ float to(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from(float v) { return __float2bfloat16_rn(v); }
};
template <>
struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&f)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  static __device__ __forceinline__ float to(float v) { return v; }
  static __device__ __forceinline__ float from(float v) { return v; }
};

// MODE 0: Stockham ping-pong (n >= 8, two fp32 copies fit); MODE 1: in-place padded butterflies (any n)
template <typename T, int MODE>
__global__ void __launch_bounds__(512) hadamard_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t rows,
                                                       int n, int log2n, int rows_per_cta, float scale, int vec) {
  extern __shared__ __align__(16) float s[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
  int rows_here = rows_per_cta;
  if (row0 + rows_here > rows) rows_here = (int)(rows - row0);
  const int total = rows_here * n;
  const T* xin = x + row0 * n;
  T* yout = y + row0 * n;
  if (MODE == 0) {
    float* a = s;
    float* b = s + (size_t)rows_per_cta * n;
    for (int o = tid; o < (total >> 3); o += nt) {
      float f[8];
      if (vec) {
        Vec8<T>::load(xin + (size_t)o * 8, f);
      } else {
#pragma unroll
        for (int j = 0; j < 8; j++) f[j] = Vec8<T>::to(xin[(size_t)o * 8 + j]);
      }
      reinterpret_cast<float4*>(a + o * 8)[0] = make_float4(f[0], f[1], f[2], f[3]);
      reinterpret_cast<float4*>(a + o * 8)[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    __syncthreads();
    const float* cur = stockham_hi(a, b, total, log2n, tid, nt);
    for (int o = tid; o < (total >> 3); o += nt) {
      float f[8];
      stockham_last(cur, log2n, o, f);
#pragma unroll
      for (int j = 0; j < 8; j++) f[j] *= scale;
      if (vec) {
        Vec8<T>::store(yout + (size_t)o * 8, f);
      } else {
#pragma unroll
        for (int j = 0; j < 8; j++) yout[(size_t)o * 8 + j] = Vec8<T>::from(f[j]);
      }
    }
  } else {
    for (int i = tid; i < total; i += nt) s[spad(i)] = Vec8<T>::to(xin[i]);
    __syncthreads();
    fwht_smem(s, total, log2n, 0, tid, nt);
    for (int i = tid; i < total; i += nt) yout[i] = Vec8<T>::from(s[spad(i)] * scale);
  }
}

}  // namespace qb

using namespace qb;

extern "C" int quipb200_hadamard(const void* x, void* y, int64_t rows, int n, float scale, int dtype,
                                 void* stream) {
  if (rows < 0 || n < 1 || (n & (n - 1)) != 0 || n > 32768) return QUIPB200_EINVAL;
  if (dtype < 0 || dtype > 2) return QUIPB200_EINVAL;
  if (rows == 0) return 0;
  if (!x || !y) return QUIPB200_EINVAL;
  int log2n = 0;
  while ((1 << log2n) < n) log2n++;
  const bool pingpong = n >= 8 && n <= 8192;
  int rows_per_cta = 1;
  if (n < 2048) rows_per_cta = 2048 / n;
  if (rows_per_cta > rows) rows_per_cta = (int)rows;
  const int total = rows_per_cta * n;
  int threads = total / 8;
  if (threads < 32) threads = 32;
  if (threads > 512) threads = 512;
  threads = (threads + 31) / 32 * 32;
  const size_t smem = pingpong ? (size_t)total * 8 : spad_host((size_t)total) * sizeof(float);
  const int64_t grid = (rows + rows_per_cta - 1) / rows_per_cta;
  if (grid > 0x7fffffff) return QUIPB200_EINVAL;
  const int vec = aligned16(x) && aligned16(y) ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_HAD(T, MODE)                                                                          \
  do {                                                                                               \
    if (smem > 48 * 1024) {                                                                          \
      cudaError_t e = cudaFuncSetAttribute(hadamard_kernel<T, MODE>,                                 \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
      if (e != cudaSuccess) return (int)e;                                                           \
    }                                                                                                \
    hadamard_kernel<T, MODE><<<(unsigned)grid, threads, smem, st>>>((const T*)x, (T*)y, rows, n, log2n, \
                                                                    rows_per_cta, scale, vec);       \
  } while (0)
  if (pingpong) {
    if (dtype == 0) LAUNCH_HAD(__half, 0);
    else if (dtype == 1) LAUNCH_HAD(__nv_bfloat16, 0);
    else LAUNCH_HAD(float, 0);
  } else {
    if (dtype == 0) LAUNCH_HAD(__half, 1);
    else if (dtype == 1) LAUNCH_HAD(__nv_bfloat16, 1);
    else LAUNCH_HAD(float, 1);
  }
#undef LAUNCH_HAD
  QB_LAUNCH_CHECK();
  return 0;
}
