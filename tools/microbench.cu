// Instruction-throughput probes for the decode inner loop (run on the B200 box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench tools/microbench.cu && gpurun_out/microbench
// Reports thread-level results per clock per SM for each op class, so the per-code instruction budget
// of the E8P GEMV can be turned into an expected codes/clk/SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define NREG 8

#define DEF_KERNEL(NAME, BODY)                                                         \
  __global__ void __launch_bounds__(1024) NAME(uint32_t* out, long long* cyc, uint32_t seed) { \
    uint32_t r[NREG];                                                                  \
    _Pragma("unroll") for (int i = 0; i < NREG; i++) r[i] = seed + threadIdx.x * 17 + i; \
    uint32_t a = seed | 1, b = seed * 3 + threadIdx.x;                                 \
    __syncthreads();                                                                   \
    long long t0 = clock64();                                                          \
    for (int it = 0; it < ITERS; it++) {                                               \
      _Pragma("unroll") for (int i = 0; i < NREG; i++) { BODY }                        \
    }                                                                                  \
    long long t1 = clock64();                                                          \
    uint32_t acc = 0;                                                                  \
    _Pragma("unroll") for (int i = 0; i < NREG; i++) acc ^= r[i];                      \
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + a + b;                          \
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                   \
  }

DEF_KERNEL(k_dp4a_ss, asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(a), "r"(b));)
DEF_KERNEL(k_dp4a_su, asm volatile("dp4a.s32.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(a), "r"(b));)
DEF_KERNEL(k_imad, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));)
DEF_KERNEL(k_lop3, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(a), "r"(b));)
DEF_KERNEL(k_prmt, asm volatile("prmt.b32 %0, %0, %1, 0xba98;" : "+r"(r[i]) : "r"(a));)
DEF_KERNEL(k_popc, asm volatile("popc.b32 %0, %0;" : "+r"(r[i]));)
DEF_KERNEL(k_shf, asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(r[i]) : "r"(a));)
DEF_KERNEL(k_iadd, asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(a));)
DEF_KERNEL(k_hfma2, asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));)
DEF_KERNEL(k_ffma, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&r[i]) : "f"(1.0001f), "f"(0.5f));)
// mixes: one fma-pipe op + one alu-pipe op per slot
DEF_KERNEL(k_mix_dp4a_lop3, asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(a), "r"(b));
           asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[(i + 4) % NREG]) : "r"(a), "r"(b));)
DEF_KERNEL(k_mix_imad_prmt, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
           asm volatile("prmt.b32 %0, %0, %1, 0xba98;" : "+r"(r[(i + 4) % NREG]) : "r"(a));)
DEF_KERNEL(k_mix_dp4a_imad, asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(a), "r"(b));
           asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[(i + 4) % NREG]) : "r"(a), "r"(b));)

// shared-memory 64-bit lookups: conflict-free replicated layout vs random addresses in a 2 KB table
__global__ void __launch_bounds__(1024) k_lds64(uint32_t* out, long long* cyc, uint32_t seed, int mode) {
  extern __shared__ unsigned char sm[];
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = i * 2654435761u;
  __syncthreads();
  uint32_t idx[NREG];
  const uint32_t lane_off = (threadIdx.x & 15) << 3;
  for (int i = 0; i < NREG; i++) idx[i] = (seed * 31 + threadIdx.x * 7 + i * 13) & 255;
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < NREG; i++) {
      uint32_t addr = mode == 0 ? ((idx[i] << 7) | lane_off) : (idx[i] << 3);
      uint2 v = *(const uint2*)(sm + addr);
      acc += v.x ^ v.y;
      idx[i] = (idx[i] * 5 + (v.x & 3) + 1) & 255;   // data-dependent next index (random walk)
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// legacy tensor path: int8 and fp16 mma.sync
__global__ void __launch_bounds__(1024) k_imma(uint32_t* out, long long* cyc, uint32_t seed) {
  int c[4][4] = {};
  uint32_t a0 = seed, a1 = seed * 3, a2 = seed * 5, a3 = seed * 7, b0 = threadIdx.x, b1 = threadIdx.x * 3;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  int acc = 0;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) acc ^= c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void __launch_bounds__(1024) k_hmma(uint32_t* out, long long* cyc, uint32_t seed) {
  float c[4][4] = {};
  uint32_t a0 = 0x3c003c00, a1 = a0, a2 = a0, a3 = a0, b0 = 0x3c003c00, b1 = b0;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  float acc = 0;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) acc += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)acc + seed;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
static void run(const char* name, F launch, int threads, double ops_per_thread, int blocks_per_sm, int sms) {
  uint32_t* out; long long* cyc;
  int blocks = sms * blocks_per_sm;
  cudaMalloc(&out, (size_t)blocks * threads * 4);
  cudaMalloc(&cyc, blocks * 8);
  launch(blocks, threads, out, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  launch(blocks, threads, out, cyc);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long* h = new long long[blocks];
  cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
  double per_clk_sm = ops_per_thread * threads * blocks_per_sm / avg;
  printf("%-22s thr/blk %4d blk/SM %d : %8.1f cyc  -> %7.2f thread-ops/clk/SM  (%.3f ms, %s)\n", name, threads,
         blocks_per_sm, avg, per_clk_sm, ms, cudaGetErrorString(err));
  delete[] h; cudaFree(out); cudaFree(cyc);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("device %s, %d SMs, clock %.0f MHz\n", p.name, sms, p.clockRate / 1000.0);
  const double N = (double)ITERS * NREG;
#define RUN1(K, OPS) run(#K, [](int b, int t, uint32_t* o, long long* c) { K<<<b, t>>>(o, c, 12345u); }, 1024, OPS, 1, sms)
  RUN1(k_dp4a_ss, N); RUN1(k_dp4a_su, N); RUN1(k_imad, N); RUN1(k_lop3, N); RUN1(k_prmt, N); RUN1(k_popc, N);
  RUN1(k_shf, N); RUN1(k_iadd, N); RUN1(k_hfma2, N); RUN1(k_ffma, N);
  RUN1(k_mix_dp4a_lop3, 2 * N); RUN1(k_mix_imad_prmt, 2 * N); RUN1(k_mix_dp4a_imad, 2 * N);
  cudaFuncSetAttribute(k_lds64, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  run("k_lds64_replicated", [](int b, int t, uint32_t* o, long long* c) { k_lds64<<<b, t, 32768>>>(o, c, 7u, 0); }, 1024, N, 1, sms);
  run("k_lds64_random2KB", [](int b, int t, uint32_t* o, long long* c) { k_lds64<<<b, t, 32768>>>(o, c, 7u, 1); }, 1024, N, 1, sms);
  run("k_imma_m16n8k32 (warp-instr x32)", [](int b, int t, uint32_t* o, long long* c) { k_imma<<<b, t>>>(o, c, 3u); }, 1024, (double)ITERS * 4, 1, sms);
  run("k_hmma_m16n8k16 (warp-instr x32)", [](int b, int t, uint32_t* o, long long* c) { k_hmma<<<b, t>>>(o, c, 3u); }, 1024, (double)ITERS * 4, 1, sms);
  return 0;
}
