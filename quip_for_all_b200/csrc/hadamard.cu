// quip_lib::hadamard -- Walsh-Hadamard transform over the last dim (Sylvester order), fp32 internal,
// y = H_n x * scale.  Replaces the third-party fast_hadamard_transform_cuda (register_lib.py:18-20);
// semantics pinned by the reference's own butterfly quant.py:50-59.
//
// One CTA transforms ROWS_PER_CTA rows held in (padded) shared memory: the first radix-8 pass runs
// in registers on the 8 contiguous elements each thread loads with one 128-bit access, the remaining
// passes are conflict-free radix-8 sweeps over shared memory.  HBM traffic: read + write of the
// activation tensor, once.
#include "common.cuh"

namespace qb {

template <typename T>
struct Cvt;
template <>
struct Cvt<__half> {
  static __device__ __forceinline__ float to(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from(float v) { return __float2half_rn(v); }
};
template <>
struct Cvt<__nv_bfloat16> {
  static __device__ __forceinline__ float to(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from(float v) { return __float2bfloat16_rn(v); }
};
template <>
struct Cvt<float> {
  static __device__ __forceinline__ float to(float v) { return v; }
  static __device__ __forceinline__ float from(float v) { return v; }
};

// total = rows_here * n elements; blocks of n
template <typename T>
__global__ void __launch_bounds__(512) hadamard_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                       int64_t rows, int n, int log2n, int rows_per_cta,
                                                       float scale) {
  extern __shared__ __align__(16) float s[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
  int rows_here = rows_per_cta;
  if (row0 + rows_here > rows) rows_here = (int)(rows - row0);
  const int total = rows_here * n;
  const T* xin = x + row0 * n;
  T* yout = y + row0 * n;

  int b0 = 0;
  if (log2n >= 3) {
    // radix-8 on contiguous octets straight from global memory
    for (int g = tid; g < (total >> 3); g += nt) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = Cvt<T>::to(xin[(size_t)g * 8 + j]);
      butterfly_regs<3>(v);
#pragma unroll
      for (int j = 0; j < 8; j++) s[spad(g * 8 + j)] = v[j];
    }
    b0 = 3;
  } else {
    for (int i = tid; i < total; i += nt) s[spad(i)] = Cvt<T>::to(xin[i]);
  }
  __syncthreads();
  fwht_smem(s, total, log2n, b0, tid, nt);
  for (int i = tid; i < total; i += nt) yout[i] = Cvt<T>::from(s[spad(i)] * scale);
}

}  // namespace qb

using namespace qb;

extern "C" int quipb200_hadamard(const void* x, void* y, int64_t rows, int n, float scale, int dtype,
                                 void* stream) {
  if (!x || !y || rows < 0 || n < 1 || (n & (n - 1)) != 0 || n > 32768) return QUIPB200_EINVAL;
  if (dtype < 0 || dtype > 2) return QUIPB200_EINVAL;
  if (rows == 0) return 0;
  int log2n = 0;
  while ((1 << log2n) < n) log2n++;
  int rows_per_cta = 1;
  if (n < 2048) rows_per_cta = 2048 / n;
  if (rows_per_cta > rows) rows_per_cta = (int)rows;
  const int total = rows_per_cta * n;
  int threads = total / 8;
  if (threads < 32) threads = 32;
  if (threads > 512) threads = 512;
  threads = (threads + 31) / 32 * 32;
  const size_t smem = spad_host((size_t)total) * sizeof(float);
  const int64_t grid = (rows + rows_per_cta - 1) / rows_per_cta;
  if (grid > 0x7fffffff) return QUIPB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_HAD(T)                                                                               \
  do {                                                                                              \
    if (smem > 48 * 1024) {                                                                         \
      cudaError_t e = cudaFuncSetAttribute(hadamard_kernel<T>,                                      \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return (int)e;                                                          \
    }                                                                                               \
    hadamard_kernel<T><<<(unsigned)grid, threads, smem, st>>>((const T*)x, (T*)y, rows, n, log2n,   \
                                                              rows_per_cta, scale);                 \
  } while (0)
  if (dtype == 0) LAUNCH_HAD(__half);
  else if (dtype == 1) LAUNCH_HAD(__nv_bfloat16);
  else LAUNCH_HAD(float);
#undef LAUNCH_HAD
  QB_LAUNCH_CHECK();
  return 0;
}
