// Walsh-Hadamard transforms on the legacy tensor path (mma.sync m16n8k16 with A = H_16 / 4), register resident.
// Shared by decode_step.cu (bs = 1 persistent kernel) and rotate_batched.cu (batched rotations of the M >= 17 path).
#pragma once
#include "common.cuh"

namespace qb {

// ---------------------------------------------------------------------------------------------
// Walsh-Hadamard transforms on the (legacy) tensor path, register resident.
//   H_256 = H_16 (x) H_16 and H_4096 = H_16 (x) H_16 (x) H_16; one factor = one mma.sync.m16n8k16 with A = H_16 / 4
//   (+-0.25: exact in fp16, and three factors give exactly the 1/sqrt(4096) of quant.py:75, two the 1/sqrt(256)).
//   A 16x16 tile M[x][y] lives in a warp in "F(x,y)" layout = the C fragments of its two 8-column halves:
//       r0,r1 = M[g][2t,2t+1]   r2,r3 = M[g+8][2t,2t+1]   r4,r5 = M[g][2t+8,2t+9]   r6,r7 = M[g+8][2t+8,2t+9]
//   (g = lane/4, t = lane%4).  Re-packed pairwise these registers ARE the B fragments of M^T, so
//       hT:  F(x,y) -> F(y',x)   transforms the y index with no data movement at all;
//   applying it twice transforms both indices and restores the layout.  fp32 intermediates are fed back as
//   hi + lo fp16 pairs (two mma), so every factor is exact to ~22 bits: the arithmetic is at least as
//   accurate as the fp32 butterflies it replaces.
// ---------------------------------------------------------------------------------------------
struct HFrag { uint32_t a[4]; };
__device__ __forceinline__ HFrag make_hfrag(int lane) {
  const int g = lane >> 2, t = lane & 3;
  auto h = [](int r, int c) -> uint32_t { return (__popc(r & c) & 1) ? 0xB400u : 0x3400u; };   // -+0.25
  HFrag f;
  f.a[0] = h(g, 2 * t) | (h(g, 2 * t + 1) << 16);
  f.a[1] = h(g + 8, 2 * t) | (h(g + 8, 2 * t + 1) << 16);
  f.a[2] = h(g, 2 * t + 8) | (h(g, 2 * t + 9) << 16);
  f.a[3] = h(g + 8, 2 * t + 8) | (h(g + 8, 2 * t + 9) << 16);
  return f;
}
__device__ __forceinline__ uint32_t pk_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// packed fp16 pairs (p[0] = r0r1, p[1] = r2r3, p[2] = r4r5, p[3] = r6r7) -> transformed tile (fp32)
__device__ __forceinline__ void hT_packed(const uint32_t (&p)[4], const HFrag& A, float (&r)[8]) {
  float d0[4], d1[4];
  const uint32_t b0[2] = {p[0], p[2]}, b1[2] = {p[1], p[3]};
  mma_16816_zero(d0, A.a, b0);
  mma_16816_zero(d1, A.a, b1);
#pragma unroll
  for (int j = 0; j < 4; j++) { r[j] = d0[j]; r[4 + j] = d1[j]; }
}
// fp32 tile -> transformed tile, inputs split into hi + lo fp16 parts
__device__ __forceinline__ void hT_split(float (&r)[8], const HFrag& A) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const __half2 h = __floats2half2_rn(r[2 * q], r[2 * q + 1]);
    const float2 hf = __half22float2(h);
    hi[q] = *reinterpret_cast<const uint32_t*>(&h);
    lo[q] = pk_h2(r[2 * q] - hf.x, r[2 * q + 1] - hf.y);
  }
  float d0[4], d1[4];
  const uint32_t b0h[2] = {hi[0], hi[2]}, b1h[2] = {hi[1], hi[3]};
  const uint32_t b0l[2] = {lo[0], lo[2]}, b1l[2] = {lo[1], lo[3]};
  mma_16816_zero(d0, A.a, b0h);
  mma_16816_zero(d1, A.a, b1h);
  mma_16816(d0, A.a, b0l);
  mma_16816(d1, A.a, b1l);
#pragma unroll
  for (int j = 0; j < 4; j++) { r[j] = d0[j]; r[4 + j] = d1[j]; }
}
// 256-point transform (x 1/16) of a block whose fp16 pairs are given in F(n, k) layout, element = 16 n + k
__device__ __forceinline__ void fwht256_frag(const uint32_t (&p)[4], const HFrag& A, float (&r)[8]) {
  hT_packed(p, A, r);
  hT_split(r, A);
}
// position of register pair q (q = 0..3 <-> r[2q], r[2q+1]) inside a 16x16 tile in F layout: x*16 + y
__device__ __forceinline__ int frag_x(int lane, int q) { return (lane >> 2) + 8 * (q & 1); }
__device__ __forceinline__ int frag_y(int lane, int q) { return 2 * (lane & 3) + 8 * (q >> 1); }

// 4096-point transform (x 1/64) across the 16 warps of the CTA.  In: fp16 pairs of tile F(x, y) of warp w.
// The two tile indices are transformed in registers, the warp index after one exchange through shared memory
// (S: 16 rows of DS_XROW floats; bank-conflict free both ways).  Out: r = F(w', x') of warp x'' ... precisely: the
// value at (row index = warp-index group, x, y) ends up as r of the lane holding (x' = g(+8): old warp group,
// y' = pairs: old x group) in warp = old y group.  Callers use idx_out / idx_in below.
constexpr int DS_XROW = 388;
// `active` is warp-uniform: CTAs with more than 16 warps pass false for the extra warps, which only take part in the barrier.
__device__ __forceinline__ void fwht4096_frag(const uint32_t (&p)[4], const HFrag& A, float* S, int warp, int lane,
                                              float (&r)[8], bool active = true) {
  if (active) {
    hT_packed(p, A, r);
    hT_split(r, A);
#pragma unroll
    for (int q = 0; q < 4; q++)
      *reinterpret_cast<float2*>(S + warp * DS_XROW + frag_x(lane, q) * 24 + frag_y(lane, q)) = make_float2(r[2 * q], r[2 * q + 1]);
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      r[2 * q] = S[frag_y(lane, q) * DS_XROW + warp * 24 + frag_x(lane, q)];
      r[2 * q + 1] = S[(frag_y(lane, q) + 1) * DS_XROW + warp * 24 + frag_x(lane, q)];
    }
    hT_split(r, A);
  } else {
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = 0.f;
  }
}
// Element index of register pair q (first element; the second is +1).  Natural placement (see frag_to_octet below):
// on the block side lane l of warp w holds octet 32 w + l, its 16-byte word j feeding p[{0,2,1,3}[j]]; the transform
// then leaves the result on the spread side, where pairs q and q + 2 are adjacent (4 halfs = 8 bytes):
//   block layout   : input of the output-side rotation, output of the input-side rotation (records: octet = thread)
//   spread layout  : output of the output-side rotation, input of the input-side rotation
// (tools/emu_fwht_frag.py checks both maps against a dense H_4096 with a numpy model of the fragments.)
__device__ __forceinline__ int idx_block(int warp, int lane, int q) { return warp * 256 + lane * 8 + 2 * (((q & 1) << 1) | (q >> 1)); }
__device__ __forceinline__ int idx_spread(int warp, int lane, int q) {
  return ((lane >> 2) + 8 * (q & 1)) * 256 + (warp & 7) * 32 + (lane & 3) * 8 + (warp >> 3) * 4 + (q >> 1) * 2;
}

// A 4096-element fp16 vector in shared memory that is accessed at spread-layout positions (4 consecutive halfs per lane,
// the eight row groups of a warp 512 bytes apart = the same banks) keeps its 16-byte chunk c at chunk c ^ ((c >> 5) & 7).
__device__ __forceinline__ int stg_chunk(int c) { return c ^ ((c >> 5) & 7); }

__device__ __forceinline__ uint32_t ldg_h2(const __half* p, int i) { return __ldg(reinterpret_cast<const unsigned int*>(p + i)); }
__device__ __forceinline__ __half2 as_h2(uint32_t v) { return *reinterpret_cast<const __half2*>(&v); }
__device__ __forceinline__ uint32_t as_u32(__half2 v) { return *reinterpret_cast<const uint32_t*>(&v); }

__device__ __forceinline__ uint4 hmul2x4(const uint4& a, const uint4& b) {
  uint4 r;
  r.x = as_u32(__hmul2(as_h2(a.x), as_h2(b.x)));
  r.y = as_u32(__hmul2(as_h2(a.y), as_h2(b.y)));
  r.z = as_u32(__hmul2(as_h2(a.z), as_h2(b.z)));
  r.w = as_u32(__hmul2(as_h2(a.w), as_h2(b.w)));
  return r;
}
// (the _rn form keeps ptxas from contracting a preceding fp16 multiply and this add into one fma: the reference rounds
// `out * SV` and `+ bias` separately, qlinear.py:112-114)
__device__ __forceinline__ uint4 hadd2x4(const uint4& a, const uint4& b) {
  uint4 r;
  r.x = as_u32(__hadd2_rn(as_h2(a.x), as_h2(b.x)));
  r.y = as_u32(__hadd2_rn(as_h2(a.y), as_h2(b.y)));
  r.z = as_u32(__hadd2_rn(as_h2(a.z), as_h2(b.z)));
  r.w = as_u32(__hadd2_rn(as_h2(a.w), as_h2(b.w)));
  return r;
}

// Natural placement.  F(x, y) is a bit permutation of the natural index of a 256-element block, and H_256 (Sylvester)
// is invariant under a simultaneous bit permutation of its row and column index, so lane l may hand the words of its
// own 16-byte octet (elements 8l .. 8l+7) to fwht256_frag as p = {x, z, y, w} (x, y and z, w stay register pairs: the
// two B fragments) and gets that octet's transformed elements back in r; this packs them (x scale) in memory order.
__device__ __forceinline__ uint4 frag_to_octet(const float (&r)[8], float scale) {
  uint4 v;
  v.x = pk_h2(r[0] * scale, r[1] * scale);
  v.z = pk_h2(r[2] * scale, r[3] * scale);
  v.y = pk_h2(r[4] * scale, r[5] * scale);
  v.w = pk_h2(r[6] * scale, r[7] * scale);
  return v;
}

}  // namespace qb
