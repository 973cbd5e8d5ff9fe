// E8P12 bs=1 GEMV inner-loop probe: computed sign decode (round 1) against replicated lookup tables.
//
//   VAR 0  abs table 256 x 8 B (bank conflicts on the random index), signs computed (popc, 2 x imad + prmt, parity term)
//   VAR 1  abs table and sign-mask table replicated per half-warp lane (16 copies, 128 B apart, interleaved in 256-byte
//          rows): address = one PRMT, conflict-free LDS.64, weights = table ^ mask (parity folded into the mask as ^0x02)
//   VAR 2  same with 32 copies (256 B per table row, two separate tables)
//
// A CTA owns a contiguous row range, a warp a 512-byte column chunk (activation records in registers), U rows in flight
// plus U prefetched.  Prints GB/s of packed codes and checks all variants against each other.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/gemv_lut_bench tools/gemv_lut_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ int dp4a_ss(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp4a_su(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint4 ldg_stream_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
  return r;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int THREADS = 512, WARPS = 16;

// sign-mask entry for sign byte s8: bytes 0..3 in .x, 4..7 in .y; 0xfc where negated, | 0x02 when parity is odd
__host__ __device__ inline void sign_mask(uint32_t s8, uint32_t& lo, uint32_t& hi) {
  uint32_t par = 0;
  for (int b = 0; b < 8; b++) par ^= (s8 >> b) & 1u;
  const uint32_t s = s8 ^ par;
  lo = hi = 0;
  for (int j = 0; j < 4; j++) {
    if ((s >> (7 - j)) & 1u) lo |= 0xfcu << (8 * j);
    if ((s >> (3 - j)) & 1u) hi |= 0xfcu << (8 * j);
  }
  if (par) { lo |= 0x02020202u; hi |= 0x02020202u; }
}

template <int VAR, int U>
__global__ void __launch_bounds__(THREADS, 1) gemv_kernel(const unsigned char* __restrict__ q, int nrows, int nseg,
                                                           const uint4* __restrict__ xq_g, const uint2* __restrict__ grid,
                                                           int* __restrict__ out) {
  extern __shared__ __align__(256) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // layout: [tables][xq]
  constexpr int TAB_BYTES = VAR == 0 ? 2048 : (VAR == 1 ? 65536 : 131072);
  uint4* xq = reinterpret_cast<uint4*>(smem + TAB_BYTES);
  if (VAR == 0) {
    if (tid < 256) {
      uint2 t = grid[tid];
      t.x |= 0x01010101u; t.y |= 0x01010101u;
      reinterpret_cast<uint2*>(smem)[tid] = t;
    }
  } else if (VAR == 1) {
    // row i (256 B): [abs entry i x 16 copies][sign entry i x 16 copies]
    for (int e = tid; e < 256 * 32; e += THREADS) {
      const int i = e >> 5, c = e & 31;
      uint2 v;
      if (c < 16) { v = grid[i]; v.x |= 0x01010101u; v.y |= 0x01010101u; }
      else sign_mask((uint32_t)i, v.x, v.y);
      reinterpret_cast<uint2*>(smem)[e] = v;
    }
  } else {
    for (int e = tid; e < 256 * 32; e += THREADS) {
      const int i = e >> 5;
      uint2 v = grid[i]; v.x |= 0x01010101u; v.y |= 0x01010101u;
      reinterpret_cast<uint2*>(smem)[e] = v;
      sign_mask((uint32_t)i, v.x, v.y);
      reinterpret_cast<uint2*>(smem + 65536)[e] = v;
    }
  }
  for (int i = tid; i < nseg; i += THREADS) xq[i] = xq_g[i];
  __syncthreads();

  const int row_bytes = nseg * 2;
  const int C = (nseg / 8 + 31) / 32;            // 512-byte column chunks per row
  const int g = WARPS / C;                       // row phases
  const int chunk = warp / g, sub = warp % g;
  const int base = nrows / gridDim.x, rem = nrows % gridDim.x;
  const int row_begin = blockIdx.x * base + min((int)blockIdx.x, rem);
  const int my_rows = base + ((int)blockIdx.x < rem ? 1 : 0);
  const int seg0 = (chunk * 32 + lane) * 8;
  const bool lane_valid = chunk < C && seg0 < nseg;
  uint32_t xs[8][4];
  int xsum[8];
#pragma unroll
  for (int s = 0; s < 8; s++) {
    uint4 r = make_uint4(0, 0, 0, 0);
    if (lane_valid) r = xq[seg0 + s];
    xs[s][0] = r.x; xs[s][1] = r.y; xs[s][2] = r.z; xs[s][3] = r.w;
    const int sh = dp4a_ss(r.x, 0x01010101u, dp4a_ss(r.y, 0x01010101u, 0));
    const int sl = dp4a_su(0x01010101u, r.z, dp4a_su(0x01010101u, r.w, 0));
    xsum[s] = sh * 256 + sl;
  }
  const unsigned char* colp = q + (size_t)(chunk * 32 + lane) * 16 + (size_t)row_begin * row_bytes;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t offA = VAR == 1 ? (uint32_t)(lane & 15) * 8 : (uint32_t)lane * 8;
  const uint32_t offS = VAR == 1 ? 128u + (uint32_t)(lane & 15) * 8 : (uint32_t)lane * 8;
  const uint32_t sbaseS = VAR == 2 ? sbase + 65536 : sbase;

  uint4 cw[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const int r = sub + u * g;
    cw[u] = make_uint4(0, 0, 0, 0);
    if (lane_valid && r < my_rows) cw[u] = ldg_stream_v4(colp + (size_t)r * row_bytes);
  }
  for (int r0 = sub; r0 < my_rows; r0 += g * U) {
    uint4 nx[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int r = r0 + g * U + u * g;
      nx[u] = make_uint4(0, 0, 0, 0);
      if (lane_valid && r < my_rows) nx[u] = ldg_stream_v4(colp + (size_t)r * row_bytes);
    }
    int tot[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t w[4] = {cw[u].x, cw[u].y, cw[u].z, cw[u].w};
      int aH = 0, aL = 0, aP = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int s = 2 * i + h;
          uint2 v;
          if (VAR == 0) {
            const uint32_t absoff = h ? ((w[i] >> 21) & 0x7f8u) : ((w[i] >> 5) & 0x7f8u);
            const uint32_t sgn = h ? prmt(w[i], 0, 0x4442) : (w[i] & 0xffu);
            const uint2 t1 = *reinterpret_cast<const uint2*>(smem + absoff);
            const uint32_t par = __popc(sgn) & 1u;
            const uint32_t sg = sgn ^ par;
            const uint32_t m_lo = prmt(sg * 0x08040201u, 0u, 0xba98u);
            const uint32_t m_hi = prmt(sg * 0x80402010u, 0u, 0xba98u);
            v.x = t1.x ^ (m_lo & 0xfcfcfcfcu);
            v.y = t1.y ^ (m_hi & 0xfcfcfcfcu);
            aP += (int)par * xsum[s];
          } else {
            // address = (index byte << 8) | lane offset: one PRMT (bytes: [0, 0, idx, off])
            const uint32_t aa = prmt(w[i], offA, h ? 0x5534u : 0x5514u);
            const uint32_t sa = prmt(w[i], offS, h ? 0x5524u : 0x5504u);
            const uint2 t1 = lds64(sbase + aa);
            const uint2 m = lds64(sbaseS + sa);
            v.x = t1.x ^ m.x;
            v.y = t1.y ^ m.y;
          }
          aH = dp4a_ss(v.x, xs[s][0], aH);
          aH = dp4a_ss(v.y, xs[s][1], aH);
          aL = dp4a_su(v.x, xs[s][2], aL);
          aL = dp4a_su(v.y, xs[s][3], aL);
        }
      }
      tot[u] = aH * 256 + aL - 2 * aP;
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int r = r0 + u * g;
      const int t = __reduce_add_sync(0xffffffffu, tot[u]);
      if (lane == 0 && r < my_rows && chunk < C) atomicAdd(out + row_begin + r, t);
    }
#pragma unroll
    for (int u = 0; u < U; u++) cw[u] = nx[u];
  }
}

template <int VAR, int U>
float run(const char* name, const unsigned char* q, int nrows, int nseg, const uint4* xq, const uint2* grid, int* out,
          std::vector<int>& host_out, int reps) {
  const int tab = VAR == 0 ? 2048 : (VAR == 1 ? 65536 : 131072);
  const int smem = tab + nseg * 16;
  CK(cudaFuncSetAttribute(gemv_kernel<VAR, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  CK(cudaMemset(out, 0, nrows * 4));
  gemv_kernel<VAR, U><<<148, THREADS, smem>>>(q, nrows, nseg, xq, grid, out);
  CK(cudaDeviceSynchronize());
  host_out.resize(nrows);
  CK(cudaMemcpy(host_out.data(), out, nrows * 4, cudaMemcpyDeviceToHost));
  float best = 1e9f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0);
    gemv_kernel<VAR, U><<<148, THREADS, smem>>>(q, nrows, nseg, xq, grid, out);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  const double bytes = (double)nrows * nseg * 2;
  printf("%-34s rows %8d K %6d: %8.1f us  %7.1f GB/s\n", name, nrows, nseg * 8, best * 1e3, bytes / best / 1e6);
  return best;
}

int main(int argc, char** argv) {
  const int K = argc > 1 ? atoi(argv[1]) : 4096;
  const long long total_mb = argc > 2 ? atoll(argv[2]) : 512;   // packed bytes (> L2) so the stream comes from HBM
  const int nseg = K / 8;
  const int nrows = (int)(total_mb * 1024 * 1024 / (nseg * 2));
  std::vector<uint16_t> hq((size_t)nrows * nseg);
  uint64_t st = 0x9e3779b97f4a7c15ull;
  for (auto& v : hq) { st = st * 6364136223846793005ull + 1442695040888963407ull; v = (uint16_t)(st >> 40); }
  std::vector<uint32_t> hx((size_t)nseg * 4);
  for (auto& v : hx) { st = st * 6364136223846793005ull + 1442695040888963407ull; v = (uint32_t)(st >> 32); }
  // abs table stand-in: bytes = 2 mod 4 (2, 6, 10, 14), byte 7 possibly negative; the probe only needs the bit pattern class
  std::vector<uint64_t> hg(256);
  for (int i = 0; i < 256; i++) {
    uint64_t e = 0;
    for (int j = 0; j < 8; j++) {
      st = st * 6364136223846793005ull + 1442695040888963407ull;
      int v = 2 + 4 * (int)((st >> 50) & 3);
      if (j == 7 && ((st >> 60) & 1)) v = -v;
      e |= (uint64_t)(uint8_t)v << (8 * j);
    }
    hg[i] = e;
  }
  unsigned char* dq; uint4* dx; uint2* dg; int* dout;
  CK(cudaMalloc(&dq, hq.size() * 2)); CK(cudaMalloc(&dx, hx.size() * 4)); CK(cudaMalloc(&dg, 2048)); CK(cudaMalloc(&dout, nrows * 4));
  CK(cudaMemcpy(dq, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dg, hg.data(), 2048, cudaMemcpyHostToDevice));
  std::vector<int> o0, o1;
  auto same = [&](const std::vector<int>& a, const std::vector<int>& b) {
    for (size_t i = 0; i < a.size(); i++) if (a[i] != b[i]) { printf("  MISMATCH at row %zu: %d vs %d\n", i, a[i], b[i]); return false; }
    return true;
  };
  run<0, 2>("computed signs, U=2", dq, nrows, nseg, dx, dg, dout, o0, 5);
  run<0, 4>("computed signs, U=4", dq, nrows, nseg, dx, dg, dout, o1, 5); same(o0, o1);
  run<1, 1>("LUT 16 copies, U=1", dq, nrows, nseg, dx, dg, dout, o1, 5); printf("  %s\n", same(o0, o1) ? "bit-identical" : "DIFFERENT");
  run<1, 2>("LUT 16 copies, U=2", dq, nrows, nseg, dx, dg, dout, o1, 5); printf("  %s\n", same(o0, o1) ? "bit-identical" : "DIFFERENT");
  run<1, 4>("LUT 16 copies, U=4", dq, nrows, nseg, dx, dg, dout, o1, 5); printf("  %s\n", same(o0, o1) ? "bit-identical" : "DIFFERENT");
  run<2, 2>("LUT 32 copies, U=2", dq, nrows, nseg, dx, dg, dout, o1, 5); printf("  %s\n", same(o0, o1) ? "bit-identical" : "DIFFERENT");
  run<2, 4>("LUT 32 copies, U=4", dq, nrows, nseg, dx, dg, dout, o1, 5); printf("  %s\n", same(o0, o1) ? "bit-identical" : "DIFFERENT");
  return 0;
}
