// Barrier / shared-memory pass latency probe (what does one FWHT pass really cost?)
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 200
template <int MODE>
__global__ void k(float* out, long long* cyc) {
  extern __shared__ float s[];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < 8192; i += nt) s[i] = i * 0.5f;
  __syncthreads();
  long long t0 = clock64();
  float acc = 0.f;
  for (int it = 0; it < ITERS; it++) {
    if (MODE == 0) {            // barrier only
      __syncthreads();
    } else if (MODE == 1) {     // radix-8 pass, strided read, 8 scalar stores, in place (stride 512)
      if (tid < 512) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = s[tid + j * 512];
#pragma unroll
        for (int h = 1; h < 8; h <<= 1)
#pragma unroll
          for (int j = 0; j < 8; j++) if (!(j & h)) { float a = v[j], c = v[j | h]; v[j] = a + c; v[j | h] = (a - c) * 0.5f; }
#pragma unroll
        for (int j = 0; j < 8; j++) s[tid + j * 512] = v[j];
      }
      __syncthreads();
    } else if (MODE == 2) {     // radix-8 pass, strided read, 2 x 128-bit contiguous stores (ping-pong)
      const float* src = s + (it & 1) * 4096;
      float* dst = s + ((it + 1) & 1) * 4096;
      if (tid < 512) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = src[tid + j * 512];
#pragma unroll
        for (int h = 1; h < 8; h <<= 1)
#pragma unroll
          for (int j = 0; j < 8; j++) if (!(j & h)) { float a = v[j], c = v[j | h]; v[j] = a + c; v[j | h] = (a - c) * 0.5f; }
        reinterpret_cast<float4*>(dst + tid * 8)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(dst + tid * 8)[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
      __syncthreads();
    } else if (MODE == 3) {     // radix-4 pass with every thread (1024 threads x 4)
      const float* src = s + (it & 1) * 4096;
      float* dst = s + ((it + 1) & 1) * 4096;
      if (tid < 1024) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = src[(tid & 1023) + j * 1024];
        float a = v[0] + v[1], b = v[0] - v[1], c = v[2] + v[3], d = v[2] - v[3];
        reinterpret_cast<float4*>(dst + (tid & 1023) * 4)[0] = make_float4(a + c, b + d, (a - c) * 0.5f, (b - d) * 0.5f);
      }
      __syncthreads();
    }
  }
  long long t1 = clock64();
  acc += s[tid];
  out[blockIdx.x * nt + tid] = acc;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int threads, int blocks) {
  float* out; long long* cyc;
  cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<MODE><<<blocks, threads, 65536>>>(out, cyc);
  k<MODE><<<blocks, threads, 65536>>>(out, cyc);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
  double a = 0; for (int i = 0; i < blocks; i++) a += h[i];
  printf("%-44s thr %4d blocks %3d: %7.1f cycles / iteration (%s)\n", name, threads, blocks, a / blocks / ITERS, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  run<0>("barrier only", 512, 148); run<0>("barrier only", 1024, 148); run<0>("barrier only", 512, 1);
  run<1>("radix-8 pass, 8 STS.32 + barrier", 512, 148); run<1>("radix-8 pass, 8 STS.32 + barrier", 1024, 148);
  run<2>("radix-8 pass, 2 STS.128 + barrier", 512, 148); run<2>("radix-8 pass, 2 STS.128 + barrier", 1024, 148);
  run<3>("radix-4 pass (all 1024 thr), 1 STS.128 + bar", 1024, 148);
  return 0;
}
