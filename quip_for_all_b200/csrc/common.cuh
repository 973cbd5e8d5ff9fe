// Shared device helpers for libquipb200 (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/quip_b200.h"

namespace qb {

extern int64_t g_launch_count;

// Enqueue-side error check: returns the cudaError_t as a positive int.
#define QB_LAUNCH_CHECK()                              \
  do {                                                 \
    ::qb::g_launch_count++;                            \
    cudaError_t e__ = cudaGetLastError();              \
    if (e__ != cudaSuccess) return (int)e__;           \
  } while (0)

extern int g_opt_pdl;

// Programmatic dependent launch: the kernel may start while its predecessor drains; it must call
// pdl_wait() before touching anything a predecessor wrote (and before its own first global write).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

static inline cudaError_t launch_kernel(const void* fn, dim3 grid, dim3 block, void** args, size_t smem,
                                        cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_opt_pdl ? 1 : 0;
  return cudaLaunchKernelExC(&cfg, fn, args);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------------------------
// memory helpers
// ---------------------------------------------------------------------------------------------
// streaming 128-bit load: weights are read exactly once per call -> keep them out of L1
__device__ __forceinline__ uint4 ldg_stream_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// L2 evict-first policy for the weight stream: 1.6 GB of codes pass through the 126 MB L2 every token;
// without the hint they evict everything else (kernel code, per-layer scale vectors, KV cache).
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint4 ldg_stream_v4(const void* p, uint64_t pol) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint2 ldg_stream_v2(const void* p, uint64_t pol) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;"
               : "=r"(r.x), "=r"(r.y)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void stg_stream_v4(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// prmt with the sign-replicate selector bit (PTX ISA prmt default mode, selector bit 3)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

// dp4a: 4 x int8 multiply-accumulate.  ss: signed x signed, su: signed a x unsigned b.
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

__device__ __forceinline__ float f16_round(float v) { return __half2float(__float2half_rn(v)); }

// ---------------------------------------------------------------------------------------------
// E8P decode (bit-exact restatement of the codebook definition, codebook/e8p12.py:82-103):
// code c: abs index c>>8 into the 256-entry packed-int8 table, sign byte c&0xff.
// Returns 8 packed int8 in units of 1/4 (lo = packed bytes 0..3, hi = bytes 4..7);
// weight i of the 8-vector is packed byte {0,2,1,3,4,6,5,7}[i].
// `tab1` entries are the reference table OR 0x01 per byte ("+1/4" pre-applied).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint2 e8p_apply_signs(uint2 t1, uint32_t sign_byte, uint32_t& parity) {
  parity = __popc(sign_byte) & 1u;
  const uint32_t s = sign_byte ^ parity;  // bit0 (-> packed byte 7) carries the parity fix
  // bit (7-j) of s -> msb of byte j  (lo: bytes 0..3 <- bits 7..4 ; hi: bytes 4..7 <- bits 3..0)
  const uint32_t m_lo = prmt(s * 0x08040201u, 0u, 0xba98u);  // 0xff where the byte is negated
  const uint32_t m_hi = prmt(s * 0x80402010u, 0u, 0xba98u);
  uint2 v;
  // negate a value whose low two bits are '11' (a|1): xor with 0xfc gives -(a)+1 ... i.e. s*a + 1
  v.x = t1.x ^ (m_lo & 0xfcfcfcfcu);
  v.y = t1.y ^ (m_hi & 0xfcfcfcfcu);
  return v;  // still needs "- 2 per byte" when parity == 1
}

// full decode to packed int8 (quarter units), bytes in packed order
__device__ __forceinline__ uint2 e8p_decode_q(uint2 t1, uint32_t code16) {
  uint32_t par;
  uint2 v = e8p_apply_signs(t1, code16 & 0xffu, par);
  // every byte is >= 2 as unsigned (values 3,7,11,15 or 0xf3..0xff) so no borrow crosses bytes
  v.x -= par * 0x02020202u;
  v.y -= par * 0x02020202u;
  return v;
}

// packed int8 (units 1/4) -> fp16 pairs, exact: 0x5c80 ^ byte == 288 + int8/4 in fp16, then -288.
// out[0]=(b0,b2) out[1]=(b1,b3) of the 32-bit word, as half2 bit patterns.
__device__ __forceinline__ void q4_to_half2(uint32_t w, __half2& even, __half2& odd) {
  const uint32_t e = (w & 0x00ff00ffu) ^ 0x5c805c80u;
  const uint32_t o = ((w >> 8) & 0x00ff00ffu) ^ 0x5c805c80u;
  const __half2 adj = __float2half2_rn(-288.0f);
  even = __hadd2(*reinterpret_cast<const __half2*>(&e), adj);
  odd = __hadd2(*reinterpret_cast<const __half2*>(&o), adj);
}

// ---------------------------------------------------------------------------------------------
// In-CTA Walsh-Hadamard butterflies on a padded fp32 shared-memory array.
// Layout: element i lives at s[spad(i)] (8 floats of padding every 64 -> radix-8 passes with
// strides 8, 64, 512.. are bank-conflict free).  The array holds `total` = K*L elements organised
// as K contiguous blocks of L = 2^log2L; butterflies never cross a block.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int spad(int i) { return i + ((i >> 6) << 3); }
static inline size_t spad_host(size_t n) { return n + ((n >> 6) << 3) + 8; }

template <int R>
__device__ __forceinline__ void butterfly_regs(float (&v)[1 << R]) {
#pragma unroll
  for (int h = 1; h < (1 << R); h <<= 1) {
#pragma unroll
    for (int j = 0; j < (1 << R); j++) {
      if (!(j & h)) {
        const float a = v[j], c = v[j | h];
        v[j] = a + c;
        v[j | h] = a - c;
      }
    }
  }
}

template <int R>
__device__ __forceinline__ void fwht_pass(float* s, int total, int b, int tid, int nthreads) {
  const int ngroups = total >> R;
  const int lomask = (1 << b) - 1;
  for (int g = tid; g < ngroups; g += nthreads) {
    const int lo = g & lomask, hi = g >> b;
    const int base = (hi << (b + R)) | lo;
    float v[1 << R];
#pragma unroll
    for (int j = 0; j < (1 << R); j++) v[j] = s[spad(base + (j << b))];
    butterfly_regs<R>(v);
#pragma unroll
    for (int j = 0; j < (1 << R); j++) s[spad(base + (j << b))] = v[j];
  }
}

// all butterfly stages for bits [b0, log2L); ends with a __syncthreads()
__device__ __forceinline__ void fwht_smem(float* s, int total, int log2L, int b0, int tid, int nthreads) {
  int b = b0;
  while (b < log2L) {
    const int r = (log2L - b >= 3) ? 3 : (log2L - b);
    if (r == 3) fwht_pass<3>(s, total, b, tid, nthreads);
    else if (r == 2) fwht_pass<2>(s, total, b, tid, nthreads);
    else fwht_pass<1>(s, total, b, tid, nthreads);
    b += r;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// legacy tensor path (mma.sync m16n8k16 fp16 -> fp32) for the small K x K orthogonal-block mixes
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
               : "=r"(r[0]), "=r"(r[1])
               : "r"(smem_u32(p)));
}
// ---------------------------------------------------------------------------------------------
// mbarrier / bulk-copy (TMA, 1-D) wrappers.  Barrier and buffer operands are 32-bit shared-memory addresses.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// global -> shared, `bytes` % 16 == 0, both addresses 16-byte aligned; completion is signalled on `bar` (complete_tx)
__device__ __forceinline__ void bulk_load(uint32_t sdst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(sdst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
// shared -> global as one bulk group of the issuing thread
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (bulk stores / the next bulk load into the buffer)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// d = a b (zero accumulator input: no registers to clear)
__device__ __forceinline__ void mma_16816_zero(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "f"(0.f));
}

__device__ __forceinline__ void unpack_h8(const uint4& v, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack_h8(const float (&f)[8]) {
  uint4 v;
  __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; i++) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

// Butterfly passes with compile-time bit position: all shared-memory offsets become immediates.
template <int R, int B>
__device__ __forceinline__ void fwht_pass_c(float* s, int total, int tid, int nthreads) {
  const int ngroups = total >> R;
  for (int g = tid; g < ngroups; g += nthreads) {
    const int lo = g & ((1 << B) - 1), hi = g >> B;
    float* p = s + spad((hi << (B + R)) | lo);
    float v[1 << R];
#pragma unroll
    for (int j = 0; j < (1 << R); j++) v[j] = p[(j << B) + (B >= 6 ? ((j << (B - 6)) << 3) : 0)];
    butterfly_regs<R>(v);
#pragma unroll
    for (int j = 0; j < (1 << R); j++) p[(j << B) + (B >= 6 ? ((j << (B - 6)) << 3) : 0)] = v[j];
  }
}
template <int R>
__device__ __forceinline__ void fwht_pass_b(float* s, int total, int B, int tid, int nt) {
  switch (B) {
    case 3: fwht_pass_c<R, 3>(s, total, tid, nt); break;
    case 6: fwht_pass_c<R, 6>(s, total, tid, nt); break;
    case 9: fwht_pass_c<R, 9>(s, total, tid, nt); break;
    default: fwht_pass_c<R, 12>(s, total, tid, nt); break;
  }
}
// Butterfly stages for bits [3, log2L), highest bits first, each pass followed by a barrier.
// Bits 0..2 are left to the caller, which finishes them in registers on contiguous octets.
static __device__ __noinline__ void fwht_hi(float* s, int total, int log2L, int tid, int nt) {
  const int nb = log2L - 3;
  if (nb <= 0) return;
  const int r0 = nb % 3;
  int B = log2L - r0;
  if (r0 == 1) { fwht_pass_b<1>(s, total, B, tid, nt); __syncthreads(); }
  else if (r0 == 2) { fwht_pass_b<2>(s, total, B, tid, nt); __syncthreads(); }
  for (B -= 3; B >= 3; B -= 3) {
    fwht_pass_b<3>(s, total, B, tid, nt);
    __syncthreads();
  }
}
// ---------------------------------------------------------------------------------------------
// Stockham-style (autosort) block Walsh-Hadamard passes between two UNPADDED fp32 buffers.
// A pass of radix 2^R on blocks of L = 2^p reads 2^R values at stride L/2^R (consecutive threads ->
// consecutive words: conflict free) and writes its 2^R results CONTIGUOUSLY (one or two 128-bit
// stores).  That matters on Blackwell: a CTA barrier drains every pending shared-memory store, and its
// cost grows with the number of STS instructions in flight per warp -- 2 wide stores instead of 8
// narrow ones per pass.  After all passes the data are back in natural order; the last radix-8 pass
// is left to the caller, which keeps the finished octet in registers.
// ---------------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void stockham_pass(const float* __restrict__ src, float* __restrict__ dst, int total,
                                              int p, int tid, int nt) {
  const int sh = p - R;                      // log2 of the read stride
  for (int g = tid; g < (total >> R); g += nt) {
    const int k = g >> sh, rest = g & ((1 << sh) - 1);
    const float* in = src + (k << p) + rest;
    float v[1 << R];
#pragma unroll
    for (int j = 0; j < (1 << R); j++) v[j] = in[j << sh];
    butterfly_regs<R>(v);
    float* out = dst + (k << p) + (rest << R);
    if (R == 3) {
      reinterpret_cast<float4*>(out)[0] = make_float4(v[0], v[1], v[2], v[3]);
      reinterpret_cast<float4*>(out)[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else if (R == 2) {
      reinterpret_cast<float4*>(out)[0] = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      reinterpret_cast<float2*>(out)[0] = make_float2(v[0], v[1]);
    }
  }
}
// All passes except the final radix-8 one (requires p >= 3).  Returns the buffer holding the result.
static __device__ __noinline__ const float* stockham_hi(float* a, float* b, int total, int p, int tid, int nt) {
  float* cur = a;
  float* nxt = b;
  const int r0 = p % 3;
  if (r0 == 1) { stockham_pass<1>(cur, nxt, total, p, tid, nt); __syncthreads(); float* t = cur; cur = nxt; nxt = t; }
  else if (r0 == 2) { stockham_pass<2>(cur, nxt, total, p, tid, nt); __syncthreads(); float* t = cur; cur = nxt; nxt = t; }
  for (int i = 0; i < p / 3 - 1; i++) {
    stockham_pass<3>(cur, nxt, total, p, tid, nt);
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
  return cur;
}
// final radix-8 pass of octet o (natural order result for elements 8o..8o+7)
__device__ __forceinline__ void stockham_last(const float* __restrict__ src, int p, int o, float (&v)[8]) {
  const int sh = p - 3;
  const int k = o >> sh, rest = o & ((1 << sh) - 1);
  const float* in = src + (k << p) + rest;
#pragma unroll
  for (int j = 0; j < 8; j++) v[j] = in[j << sh];
  butterfly_regs<3>(v);
}

// 256-point WHT of one block held by ONE WARP (lane l owns elements 8l..8l+7): bits 0..2 in registers,
// bits 3..7 across lanes with shuffles.  No shared memory, no barrier -- used for the 43 x 256 / 172 x 64-style
// block rotations where blocks are independent.
__device__ __forceinline__ void warp_fwht256(float (&v)[8], int lane) {
  butterfly_regs<3>(v);
#pragma unroll
  for (int b = 0; b < 5; b++) {
    const float sg = ((lane >> b) & 1) ? -1.f : 1.f;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float p = __shfl_xor_sync(0xffffffffu, v[j], 1 << b);
      v[j] = fmaf(sg, v[j], p);
    }
  }
}

// last stages (bits 0..nbits-1) on 8 contiguous values held in registers
__device__ __forceinline__ void butterfly_low(float (&v)[8], int nbits) {
  if (nbits >= 3) butterfly_regs<3>(v);
  else if (nbits == 2) {
#pragma unroll
    for (int h = 1; h < 4; h <<= 1)
#pragma unroll
      for (int j = 0; j < 8; j++)
        if (!(j & h)) { const float a = v[j], c = v[j | h]; v[j] = a + c; v[j | h] = a - c; }
  } else if (nbits == 1) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) { const float a = v[j], c = v[j + 1]; v[j] = a + c; v[j + 1] = a - c; }
  }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide reductions through a 32-float scratch; every thread gets the result (2 barriers)
__device__ __forceinline__ float block_max(float v, float* red, int tid, int nt) {
  v = warp_max(v);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < (nt >> 5); w++) r = fmaxf(r, red[w]);
  return r;
}
// single-barrier variant: `red` must not be in use by a reduction that other warps may still be reading
__device__ __forceinline__ float block_max1(float v, float* red, int tid, int nt) {
  v = warp_max(v);
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  float r = ((tid & 31) < (nt >> 5)) ? red[tid & 31] : 0.f;   // callers reduce |x| >= 0
  return warp_max(r);
}
__device__ __forceinline__ float block_sum(float v, float* red, int tid, int nt) {
  v = warp_sum(v);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < (nt >> 5); w++) r += red[w];
  return r;
}

}  // namespace qb
