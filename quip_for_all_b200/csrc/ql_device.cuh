// Device-side building blocks of the fused QuantLinear path (rotations, input/output sides, E8P integer
// dot product).  Shared by quantlinear.cu (one launch per linear / group) and decode_step.cu (the
// persistent whole-step kernel).  See quantlinear.cu for the design notes.
#pragma once
#include "common.cuh"
#include "fwht_mma.cuh"

namespace qb {

constexpr int PRO_THREADS = 512;    // (1024 threads / 64 registers measured slower for the rotation kernels too)
constexpr int GEMV_MAX_WARPS = 16;   // 512 threads x <=128 registers (a 1024-thread / 64-register variant measured 40 % slower: spills)
constexpr int GEMV_UNROLL = 4;
constexpr int COUNTER_SLOTS = 4096;
// octets (8 elements) a thread keeps in flight per round in the rotations: 1 in the lean instantiations (code
// size), 3 in the general ones (11008 / 512 threads = 2.7 octets per thread: one load round instead of three)
template <bool LEAN> struct Rounds { static constexpr int CH = LEAN ? 1 : 3; };

// ---------------------------------------------------------------------------------------------
// rotation workspace in shared memory
//   s  : fp32 butterfly array, padded (spad)
//   t  : fp16 result.  K == 1: t[i].  K > 1: K rows of stride Ls = L + 8 halfs (+ one zero row) --
//        the layout ldmatrix wants; the orthogonal mix runs IN PLACE on it.
//   hk : fp16 [Kp][Kp] coefficient matrix M[k_out][k_in], zero padded to Kp = roundup16(K)
// ---------------------------------------------------------------------------------------------
struct RotSmem {
  float* s;     // pp == 0: padded (spad) in-place butterfly array;  pp == 1: unpadded ping buffer
  float* s2;    // pp == 1: pong buffer
  int pp;
  __half* t;
  __half* hk;
  float* red;
  int Ls;       // row stride of t (halfs)
  int log2L;
};

__host__ __device__ static inline int kpad(int K) { return (K + 15) / 16 * 16; }
static inline size_t rot_t_halfs(int q, int K) { return K > 1 ? (size_t)(K + 1) * (q / K + 8) : (size_t)q; }
// ping-pong (Stockham) layout whenever two fp32 copies fit comfortably; else the in-place padded layout
static inline bool rot_pingpong(int q, int K) {
  int L = q / (K > 0 ? K : 1);
  return L >= 8 && (size_t)q * 8 <= 96 * 1024;
}
static inline size_t rot_smem_bytes(int q, int K) {
  size_t b = rot_pingpong(q, K) ? (size_t)q * 8 : (spad_host((size_t)q) * sizeof(float) + 15) / 16 * 16;
  b += (rot_t_halfs(q, K) * sizeof(__half) + 15) / 16 * 16;
  if (K > 1) b += (size_t)kpad(K) * kpad(K) * sizeof(__half);
  b += 64 * sizeof(float);
  return (b + 15) / 16 * 16;
}

__device__ __forceinline__ RotSmem rot_carve(unsigned char* base, int q, int K, int log2L) {
  RotSmem r;
  r.s = reinterpret_cast<float*>(base);
  const int Lb = q / (K > 0 ? K : 1);
  r.pp = (Lb >= 8 && (size_t)q * 8 <= 96 * 1024) ? 1 : 0;
  r.s2 = r.s + q;
  size_t off = r.pp ? (size_t)q * 8 : (((size_t)(q + ((q >> 6) << 3) + 8)) * sizeof(float) + 15) / 16 * 16;
  r.t = reinterpret_cast<__half*>(base + off);
  const size_t th = K > 1 ? (size_t)(K + 1) * ((q / K) + 8) : (size_t)q;
  off += (th * sizeof(__half) + 15) / 16 * 16;
  r.hk = reinterpret_cast<__half*>(base + off);
  if (K > 1) {
    const int Kp = (K + 15) / 16 * 16;
    off += (size_t)Kp * Kp * sizeof(__half);
  }
  r.red = reinterpret_cast<float*>(base + off);
  r.Ls = K > 1 ? (q / K) + 8 : q;
  r.log2L = log2L;
  return r;
}

__device__ __forceinline__ int s_index(const RotSmem& sm, int i) { return sm.pp ? i : spad(i); }

__device__ __forceinline__ int t_index(const RotSmem& sm, int K, int i) {
  return K > 1 ? (i >> sm.log2L) * sm.Ls + (i & ((1 << sm.log2L) - 1)) : i;
}

// coefficient matrix M[k_out][k_in] = hadK[k_out][k_in] (output side) or hadK[k_in][k_out] (input side,
// hadK^T; quant.py:79-80), zero padded
__device__ __forceinline__ void load_hadK(const RotSmem& sm, const __half* hadK, int K, int transpose, int tid,
                                          int nt) {
  if (K <= 1 || hadK == nullptr) return;
  const int Kp = (K + 15) / 16 * 16;
  for (int i = tid; i < Kp * Kp; i += nt) {
    const int ko = i / Kp, ki = i - ko * Kp;
    __half v = __float2half_rn(0.f);
    if (ko < K && ki < K) v = transpose ? hadK[ki * K + ko] : hadK[ko * K + ki];
    sm.hk[i] = v;
  }
}

// In-place mix t <- M t on the tensor path.  One warp owns an 8-column tile: it reads every B fragment
// of that tile before it writes, so in-place is safe.  fp16 operands, fp32 accumulate, one fp16
// rounding -- the arithmetic of the reference's `hadK @ input` fp16 GEMM (quant.py:83).
template <int MT>   // MT = Kp / 16 (1..4); larger blocks (use_rand=False K=172) take the CUDA-core mix
__device__ __forceinline__ void mix_mma_tiles(const RotSmem& sm, int K, int warp, int lane, int nwarps, int ntiles) {
  const int Kp = MT * 16;
  const int g = lane >> 2, tq = lane & 3;
  for (int nt_i = warp; nt_i < ntiles; nt_i += nwarps) {
    const int c0 = nt_i << 3;
    uint32_t bf[MT][2];
#pragma unroll
    for (int kt = 0; kt < MT; kt++) {
      int r = kt * 16 + (lane & 15);
      if (r >= K) r = K;                                   // the zero row
      ldmatrix_x2_trans(bf[kt], sm.t + (size_t)r * sm.Ls + c0);
    }
    float acc[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
#pragma unroll
      for (int j = 0; j < 4; j++) acc[mt][j] = 0.f;
#pragma unroll
      for (int kt = 0; kt < MT; kt++) {
        uint32_t af[4];
        ldmatrix_x4(af, sm.hk + (size_t)(mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * Kp + kt * 16 + (lane >> 4) * 8);
        mma_16816(acc[mt], af, bf[kt]);
      }
    }
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      if (r0 < K)
        *reinterpret_cast<__half2*>(sm.t + (size_t)r0 * sm.Ls + c0 + tq * 2) = __floats2half2_rn(acc[mt][0], acc[mt][1]);
      if (r1 < K)
        *reinterpret_cast<__half2*>(sm.t + (size_t)r1 * sm.Ls + c0 + tq * 2) = __floats2half2_rn(acc[mt][2], acc[mt][3]);
    }
  }
}

// Orthogonal-block mix (K > 1) of the fp16 rows already in t: cold path, kept out of line so the common
// power-of-two path stays compact in the instruction cache.
static __device__ __noinline__ void rotate_mix(RotSmem sm, int q, int K, int tid, int nt) {
  const int L = 1 << sm.log2L;
  const int Kp = (K + 15) / 16 * 16;
  if (L >= 8 && Kp <= 64) {
    for (int i = tid; i < sm.Ls; i += nt) sm.t[(size_t)K * sm.Ls + i] = __float2half_rn(0.f);
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
    switch (Kp >> 4) {
      case 1: mix_mma_tiles<1>(sm, K, warp, lane, nwarps, L >> 3); break;
      case 2: mix_mma_tiles<2>(sm, K, warp, lane, nwarps, L >> 3); break;
      case 3: mix_mma_tiles<3>(sm, K, warp, lane, nwarps, L >> 3); break;
      default: mix_mma_tiles<4>(sm, K, warp, lane, nwarps, L >> 3); break;
    }
    __syncthreads();
    return;
  }
  // generic CUDA-core mix (tiny blocks / very large K): t -> s (fp32 copy) -> t
  __syncthreads();
  for (int i = tid; i < q; i += nt) sm.s[s_index(sm, i)] = __half2float(sm.t[t_index(sm, K, i)]);
  __syncthreads();
  for (int i = tid; i < q; i += nt) {
    const int ko = i >> sm.log2L, c = i & (L - 1);
    float acc = 0.f;
    for (int kp = 0; kp < K; kp++)
      acc = fmaf(__half2float(sm.hk[ko * Kp + kp]), sm.s[s_index(sm, (kp << sm.log2L) + c)], acc);
    sm.t[t_index(sm, K, i)] = __float2half_rn(acc);
  }
  __syncthreads();
}

// Pure power-of-two rotations of 4096 / 8192 points on 512 threads (hidden sizes of Llama-2-7B / 70B): instead of
// 3-4 radix-8 passes through shared memory, every warp transforms its 256-wide sub-blocks in registers + shuffles as
// it loads them (warp_fwht256), parks them as rows of an [NO][256 (+16 pad)] fp32 array, and ONE pass finishes the
// log2(NO) remaining levels: thread (column w, half h) reads rows 2j + h, runs the levels over j in registers, the
// last level against its partner lane (xor 16), scales, rounds and writes the fp16 result.  The row pad puts the two
// half-warps on disjoint banks.
constexpr int WK1_STRIDE = 272;
template <int NPT>   // rows per thread = NO / 2: 8 (4096 points) or 16 (8192)
__device__ __forceinline__ void cross_rows_to_t(const float* s, __half* t, float sc, int tid) {
  const int w = (tid >> 5) * 16 + (tid & 15), h = (tid >> 4) & 1;
  float v[NPT];
#pragma unroll
  for (int j = 0; j < NPT; j++) v[j] = s[(2 * j + h) * WK1_STRIDE + w];
#pragma unroll
  for (int d = 1; d < NPT; d <<= 1)
#pragma unroll
    for (int j = 0; j < NPT; j++)
      if (!(j & d)) { const float a = v[j], b = v[j | d]; v[j] = a + b; v[j | d] = a - b; }
  const float sg = h ? -1.f : 1.f;
#pragma unroll
  for (int j = 0; j < NPT; j++) {
    const float p = __shfl_xor_sync(0xffffffffu, v[j], 16);
    t[(2 * j + h) * 256 + w] = __float2half_rn(fmaf(sg, v[j], p) * sc);
  }
}
__device__ __forceinline__ int wk1_index(int oct) { return (oct >> 5) * WK1_STRIDE + (oct & 31) * 8; }

// 512..4096-wide blocks with an orthogonal mix (Llama-2-70B: 28672 = 7 x 4096).  The caller has already run
// warp_fwht256 on every 256-wide sub-block (registers + shuffles) and stored the fp32 result in s; this finishes
// the block transform with ONE shared-memory pass -- the H_NO butterflies across the NO = L / 256 sub-blocks, two
// adjacent columns per thread -- then scales, rounds and writes the fp16 rows the mix reads.
template <int NO>
__device__ __forceinline__ void cross256_tile(const RotSmem& sm, int K, float scale, int tid, int nt) {
  constexpr int L = NO * 256;
  for (int it = tid; it < K * 128; it += nt) {
    const int base = (it >> 7) * L + (it & 127) * 2;
    float2 v[NO];
#pragma unroll
    for (int c = 0; c < NO; c++) v[c] = *reinterpret_cast<const float2*>(sm.s + s_index(sm, base + c * 256));
#pragma unroll
    for (int h = 1; h < NO; h <<= 1)
#pragma unroll
      for (int c = 0; c < NO; c++)
        if (!(c & h)) {
          const float2 a = v[c], b = v[c | h];
          v[c] = make_float2(a.x + b.x, a.y + b.y);
          v[c | h] = make_float2(a.x - b.x, a.y - b.y);
        }
#pragma unroll
    for (int c = 0; c < NO; c++)
      *reinterpret_cast<__half2*>(sm.t + t_index(sm, K, base + c * 256)) = __floats2half2_rn(v[c].x * scale, v[c].y * scale);
  }
}

static __device__ __noinline__ void rotate_cross256(RotSmem sm, int K, float scale, int tid, int nt) {
  switch (sm.log2L) {
    case 9: cross256_tile<2>(sm, K, scale, tid, nt); break;
    case 10: cross256_tile<4>(sm, K, scale, tid, nt); break;
    case 11: cross256_tile<8>(sm, K, scale, tid, nt); break;
    default: cross256_tile<16>(sm, K, scale, tid, nt); break;
  }
  __syncthreads();
}

// Rotation: in: s[spad(i)] (fp32, untransformed); out: t (fp16) = round( (M (x) H_L) s * scale ).
// Butterfly order: bits [3, log2L) in shared memory (highest first), bits 0..2 last in registers on
// the contiguous octet each thread then rounds and stores.  Rounding points follow the reference: fp16
// after the FWHT*scale (register_lib.py:20), fp16 after hadK@ (quant.py:83).  transform == 0: t = round(s).
static __device__ __noinline__ void rotate_smem(RotSmem sm, int q, int K, float scale, int transform, int tid, int nt) {
  const float sc = transform ? scale : 1.0f;
  const int L = 1 << sm.log2L;
  const bool rowvec = (K == 1) || (L >= 8);
  const float* fin = sm.s;
  if (transform && sm.pp) fin = stockham_hi(sm.s, sm.s2, q, sm.log2L, tid, nt);
  else if (transform) fwht_hi(sm.s, q, sm.log2L, tid, nt);
  const int nbits = transform ? (sm.log2L < 3 ? sm.log2L : 3) : 0;
  for (int o = tid; o < (q >> 3); o += nt) {
    float f[8];
    if (transform && sm.pp) {
      stockham_last(fin, sm.log2L, o, f);
    } else {
      const float4* sp = reinterpret_cast<const float4*>(sm.s + s_index(sm, o * 8));
      const float4 a = sp[0], b = sp[1];
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
      butterfly_low(f, nbits);
    }
#pragma unroll
    for (int j = 0; j < 8; j++) f[j] *= sc;
    if (rowvec) {
      *reinterpret_cast<uint4*>(sm.t + t_index(sm, K, o * 8)) = pack_h8(f);
    } else {
#pragma unroll
      for (int j = 0; j < 8; j++) sm.t[t_index(sm, K, o * 8 + j)] = __float2half_rn(f[j]);
    }
  }
  if (K == 1 || !transform) {
    __syncthreads();
    return;
  }
  rotate_mix(sm, q, K, tid, nt);
}

// finished (fully transformed, unscaled) octet o of a K == 1 rotation whose high passes are done
__device__ __forceinline__ void final_octet(const RotSmem& sm, const float* fin, int transform, int o, float (&f)[8]) {
  if (transform && sm.pp) {
    stockham_last(fin, sm.log2L, o, f);
  } else {
    const float4* sp = reinterpret_cast<const float4*>(sm.s + s_index(sm, o * 8));
    const float4 a = sp[0], b = sp[1];
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    if (transform) butterfly_low(f, sm.log2L < 3 ? sm.log2L : 3);
  }
}

__device__ __forceinline__ void pack_record(const float (&f)[8], float inv, uint4& r) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    int v = __float2int_rn(f[j] * inv);
    v = max(-32767, min(32767, v));
    hi[j] = (uint32_t)(v >> 8) & 0xffu;
    lo[j] = (uint32_t)v & 0xffu;
  }
  // byte order (0,2,1,3 | 4,6,5,7) = the packed-byte order of the decoded E8P word (D4 table follows suit)
  r.x = hi[0] | (hi[2] << 8) | (hi[1] << 16) | (hi[3] << 24);
  r.y = hi[4] | (hi[6] << 8) | (hi[5] << 16) | (hi[7] << 24);
  r.z = lo[0] | (lo[2] << 8) | (lo[1] << 16) | (lo[3] << 24);
  r.w = lo[4] | (lo[6] << 8) | (lo[5] << 16) | (lo[7] << 24);
}

__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

// ---------------------------------------------------------------------------------------------
// input side: [rmsnorm] [silu(gate)*x] x*SU -> rotation -> 16-bit fixed point records
// ---------------------------------------------------------------------------------------------
struct PrologueArgs {
  const __half* x;
  int64_t ldx;
  const __half* gate;     // optional: x <- silu(gate) * x
  int64_t ldgate;
  const __half* norm_w;   // optional: x <- rmsnorm(x) * norm_w
  float norm_eps;
  const __half* SU;
  const __half* hadK;
  int K, in_features, q_in, log2L, transform;
  float scale;
  uint4* xq;       // [M][q_in/8] records {H(0,2,1,3), H(4,6,5,7), L(0,2,1,3), L(4,6,5,7)}  (global)
  float* xscale;   // [M]
};

__device__ __forceinline__ bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// element-wise pre-ops on 8 consecutive inputs (rounding points = the fp16 tensor ops HF / the reference issue).
// Flags are tested once per octet, not per element, so the executed instruction stream stays dense.
__device__ __forceinline__ void pre_ops(float (&f)[8], const uint4& gv, bool has_gate, const uint4& wv, bool has_norm,
                                        float rstd, const uint4& sv, bool has_su) {
  if (has_gate) {   // LlamaMLP: act_fn(gate) * up
    float g[8];
    unpack_h8(gv, g);
#pragma unroll
    for (int j = 0; j < 8; j++) f[j] = f16_round(f16_round(silu_f(g[j])) * f[j]);
  }
  if (has_norm) {   // LlamaRMSNorm: weight * (x * rstd).to(fp16)
    float w[8];
    unpack_h8(wv, w);
#pragma unroll
    for (int j = 0; j < 8; j++) f[j] = f16_round(f16_round(f[j] * rstd) * w[j]);
  }
  if (has_su) {     // qlinear.py:91
    float su[8];
    unpack_h8(sv, su);
#pragma unroll
    for (int j = 0; j < 8; j++) f[j] = f16_round(f[j] * su[j]);
  }
}

struct PrologueArgs;
static __device__ __noinline__ void prologue_scalar_fill(const PrologueArgs& a, RotSmem sm, const __half* xr, const __half* gr,
                                                  int tid, int nt);

// 16-byte records in shared memory are XOR-swizzled so that the GEMV lanes (which each read 8
// consecutive records, i.e. a 128-byte stride between lanes) hit distinct banks
__device__ __forceinline__ int swz(int seg) { return seg ^ ((seg >> 3) & 7); }

// Computes the records into `dst` (shared: swizzled; global: linear) and returns the fixed-point scale.
#define QB_DSTAMP(i) do { if (dbg && tid == 0) dbg[i] = clock64(); } while (0)
// LEAN: the host has verified alignment, in_features % 8 == 0, K == 1, q_in / 8 <= threads and L >= 8, so only
// the vectorised, register-resident path is compiled in (instruction-cache footprint of the hot kernel).
template <bool LEAN>
__device__ __forceinline__ float prologue_body(const PrologueArgs& a, unsigned char* rot_base, uint4* dst, int m,
                                               int tid, int nt, bool swizzle, long long* dbg = nullptr) {
  constexpr int CH = Rounds<LEAN>::CH;
  const RotSmem sm = rot_carve(rot_base, a.q_in, a.K, a.log2L);
  const __half* xr = a.x + (size_t)m * a.ldx;
  const __half* gr = a.gate ? a.gate + (size_t)m * a.ldgate : nullptr;
  load_hadK(sm, a.hadK, a.K, /*transpose=*/1, tid, nt);
  const int noct = a.q_in >> 3;
  const int noct_in = a.in_features >> 3;
  const bool vec = LEAN || ((a.in_features & 7) == 0 && al16(xr) && (!gr || al16(gr)) && (!a.SU || al16(a.SU)) &&
                           (!a.norm_w || al16(a.norm_w)));
  float rstd = 1.f;
  // 256-wide blocks (11008 = 43 x 256): each warp transforms whole blocks in registers + shuffles
  const bool wf = !LEAN && vec && a.transform && a.K > 1 && a.log2L == 8 && (nt & 31) == 0;
  // wider blocks (28672 = 7 x 4096): sub-blocks of 256 in registers here, one shared-memory pass (rotate_cross256) after
  const bool wfb = !LEAN && vec && a.transform && a.K > 1 && a.log2L >= 9 && a.log2L <= 12 && (nt & 31) == 0;
  // 4096 / 8192-point pure FWHT: sub-blocks in registers on load, one cross pass (cross_rows_to_t)
  const bool wk1 = vec && a.transform && a.K == 1 && nt == 512 && (a.log2L == 12 || a.log2L == 13);
  if (vec) {
    const bool single = noct <= nt * CH;    // everything fits one round: no re-read for the norm
    uint4 xv[CH];
    if (a.norm_w) {
      float ss = 0.f;
      for (int base = 0; base < noct_in; base += nt * CH) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
          const int idx = base + c * nt + tid;
          xv[c] = make_uint4(0, 0, 0, 0);
          if (idx < noct_in) xv[c] = *reinterpret_cast<const uint4*>(xr + (size_t)idx * 8);
        }
#pragma unroll
        for (int c = 0; c < CH; c++) {
          float f[8];
          unpack_h8(xv[c], f);
          if (gr) {   // the norm never follows a gate in Llama; keep the generic order anyway
            const int idx = base + c * nt + tid;
            if (idx < noct_in)
              pre_ops(f, *reinterpret_cast<const uint4*>(gr + (size_t)idx * 8), true, xv[c], false, 1.f, xv[c], false);
          }
#pragma unroll
          for (int j = 0; j < 8; j++) ss = fmaf(f[j], f[j], ss);
        }
      }
      ss = block_sum(ss, sm.red, tid, nt);
      rstd = rsqrtf(ss / (float)a.in_features + a.norm_eps);
    }
    for (int base = 0; base < noct; base += nt * CH) {
      uint4 gv[CH], wv[CH], sv[CH];
#pragma unroll
      for (int c = 0; c < CH; c++) {
        const int idx = base + c * nt + tid;
        const bool in = idx < noct_in;
        if (!(a.norm_w && single)) {
          xv[c] = make_uint4(0, 0, 0, 0);
          if (in) xv[c] = *reinterpret_cast<const uint4*>(xr + (size_t)idx * 8);
        }
        gv[c] = wv[c] = sv[c] = make_uint4(0, 0, 0, 0);
        if (in && gr) gv[c] = *reinterpret_cast<const uint4*>(gr + (size_t)idx * 8);
        if (in && a.norm_w) wv[c] = *reinterpret_cast<const uint4*>(a.norm_w + (size_t)idx * 8);
        if (in && a.SU) sv[c] = *reinterpret_cast<const uint4*>(a.SU + (size_t)idx * 8);
      }
#pragma unroll
      for (int c = 0; c < CH; c++) {
        const int idx = base + c * nt + tid;
        if (idx < noct) {
          float f[8];
          unpack_h8(xv[c], f);
          if (idx < noct_in) pre_ops(f, gv[c], gr != nullptr, wv[c], a.norm_w != nullptr, rstd, sv[c], a.SU != nullptr);
          if (wf) {   // idx / 32 is warp-uniform: the warp owns block idx >> 5
            warp_fwht256(f, tid & 31);
#pragma unroll
            for (int j = 0; j < 8; j++) f[j] *= a.scale;
            *reinterpret_cast<uint4*>(sm.t + t_index(sm, a.K, idx * 8)) = pack_h8(f);
          } else {
            if (wfb || wk1) warp_fwht256(f, tid & 31);
            float4* d = reinterpret_cast<float4*>(sm.s + (wk1 ? wk1_index(idx) : s_index(sm, idx * 8)));
            d[0] = make_float4(f[0], f[1], f[2], f[3]);
            d[1] = make_float4(f[4], f[5], f[6], f[7]);
          }
        }
      }
    }
  } else {
    prologue_scalar_fill(a, sm, xr, gr, tid, nt);
  }
  QB_DSTAMP(9);
  __syncthreads();
  if (LEAN || (a.K == 1 && noct <= nt * CH && (!a.transform || a.log2L >= 3))) {
    // register-resident tail: last butterflies, abs-max and quantisation without another smem round trip
    const float* fin = sm.s;
    const float sc = a.transform ? a.scale : 1.0f;
    if (wk1) {
      if (LEAN || a.log2L == 12) cross_rows_to_t<8>(sm.s, sm.t, sc, tid);   // (lean instantiations: q <= 4096)
      else cross_rows_to_t<16>(sm.s, sm.t, sc, tid);
      __syncthreads();
    } else if (a.transform) {
      if (LEAN || sm.pp) fin = stockham_hi(sm.s, sm.s2, a.q_in, a.log2L, tid, nt);
      else fwht_hi(sm.s, a.q_in, a.log2L, tid, nt);
    }
    float f[CH][8];
    float mx = 0.f;
#pragma unroll
    for (int c = 0; c < CH; c++) {
      const int o = c * nt + tid;
      if (o < noct) {
        if (wk1) {
          unpack_h8(*reinterpret_cast<const uint4*>(sm.t + (size_t)o * 8), f[c]);   // already scaled and rounded
#pragma unroll
          for (int j = 0; j < 8; j++) mx = fmaxf(mx, fabsf(f[c][j]));
        } else {
          final_octet(sm, fin, a.transform, o, f[c]);
#pragma unroll
          for (int j = 0; j < 8; j++) {
            f[c][j] = f16_round(f[c][j] * sc);   // the rotated vector is an fp16 tensor in the reference
            mx = fmaxf(mx, fabsf(f[c][j]));
          }
        }
      }
    }
    QB_DSTAMP(10);
    mx = block_max1(mx, sm.red + 32, tid, nt);
    QB_DSTAMP(11);
    const float inv = (mx > 0.f) ? 32767.0f / mx : 0.f;
#pragma unroll
    for (int c = 0; c < CH; c++) {
      const int o = c * nt + tid;
      if (o < noct) {
        uint4 r;
        pack_record(f[c], inv, r);
        dst[swizzle ? swz(o) : o] = r;
      }
    }
    return (mx > 0.f) ? mx / 32767.0f : 0.f;
  }
  if (LEAN) return 0.f;   // unreachable
  if (wfb) rotate_cross256(sm, a.K, a.scale, tid, nt);
  if (wf || wfb) rotate_mix(sm, a.q_in, a.K, tid, nt);
  else rotate_smem(sm, a.q_in, a.K, a.scale, a.transform, tid, nt);
  QB_DSTAMP(10);

  // abs-max -> 16-bit fixed-point scale, then the records
  const int L = 1 << a.log2L;
  const bool rowvec = (a.K == 1) || (L >= 8);    // 8 consecutive outputs are contiguous (and 16-B aligned) in t
  float mx = 0.f;
  if (rowvec) {
    for (int sgi = tid; sgi < noct; sgi += nt) {
      float f[8];
      unpack_h8(*reinterpret_cast<const uint4*>(sm.t + t_index(sm, a.K, sgi * 8)), f);
#pragma unroll
      for (int j = 0; j < 8; j++) mx = fmaxf(mx, fabsf(f[j]));
    }
  } else {
    for (int i = tid; i < a.q_in; i += nt) mx = fmaxf(mx, fabsf(__half2float(sm.t[t_index(sm, a.K, i)])));
  }
  mx = block_max(mx, sm.red, tid, nt);
  QB_DSTAMP(11);
  const float inv = (mx > 0.f) ? 32767.0f / mx : 0.f;
  for (int sgi = tid; sgi < noct; sgi += nt) {
    float f[8];
    if (rowvec) {
      unpack_h8(*reinterpret_cast<const uint4*>(sm.t + t_index(sm, a.K, sgi * 8)), f);
    } else {
#pragma unroll
      for (int j = 0; j < 8; j++) f[j] = __half2float(sm.t[t_index(sm, a.K, sgi * 8 + j)]);
    }
    uint4 r;
    pack_record(f, inv, r);
    dst[swizzle ? swz(sgi) : sgi] = r;
  }
  return (mx > 0.f) ? mx / 32767.0f : 0.f;
}

// cold path: unaligned rows or in_features % 8 != 0
static __device__ __noinline__ void prologue_scalar_fill(const PrologueArgs& a, RotSmem sm, const __half* xr, const __half* gr,
                                                  int tid, int nt) {
  float rstd = 1.f;
  if (a.norm_w) {
    float ss = 0.f;
    for (int i = tid; i < a.in_features; i += nt) {
      float v = __half2float(xr[i]);
      if (gr) v = f16_round(f16_round(silu_f(__half2float(gr[i]))) * v);
      ss = fmaf(v, v, ss);
    }
    ss = block_sum(ss, sm.red, tid, nt);
    rstd = rsqrtf(ss / (float)a.in_features + a.norm_eps);
  }
  for (int i = tid; i < a.q_in; i += nt) {
    float v = 0.f;
    if (i < a.in_features) {
      v = __half2float(xr[i]);
      if (gr) v = f16_round(f16_round(silu_f(__half2float(gr[i]))) * v);
      if (a.norm_w) v = f16_round(f16_round(v * rstd) * __half2float(a.norm_w[i]));
      if (a.SU) v = f16_round(v * __half2float(a.SU[i]));
    }
    sm.s[s_index(sm, i)] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// output side
// ---------------------------------------------------------------------------------------------
struct EpilogueArgs {
  const float* acc;      // [M][q_out] integer dot products (main)
  const float* acc2;     // [M][q_out] residual codebook dot products or NULL
  const float* xscale;   // [M]
  float unit;            // weight unit: 0.25 (E8P) / 0.5 (D4)
  float resid_scale;     // fp16-rounded residual scale (RVQ)
  const __half* wscale_pc;
  const __half* hadK;
  int K, q_out, out_features, log2L, transform;
  float scale;           // 1/sqrt(L)
  const __half* SV;
  const __half* bias;
  const __half* residual;   // optional skip connection added after the bias (may alias y)
  int64_t ldres;
  __half* y;
  int64_t ldy;
  int mma;               // host: K == 1 and q_out == 4096 -> the 4096-point rotation runs on the tensor path (fwht_mma.cuh)
};

template <bool LEAN>
__device__ __forceinline__ void epilogue_body(const EpilogueArgs& a, unsigned char* rot_base, int m, float xscale,
                                              int tid, int nt, long long* dbg = nullptr) {
  constexpr int CH = Rounds<LEAN>::CH;
  const RotSmem sm = rot_carve(rot_base, a.q_out, a.K, a.log2L);
  load_hadK(sm, a.hadK, a.K, /*transpose=*/0, tid, nt);
  const float xs = xscale * a.unit;
  const float* ar = a.acc + (size_t)m * a.q_out;
  const float* ar2 = a.acc2 ? a.acc2 + (size_t)m * a.q_out : nullptr;
  const int noct = a.q_out >> 3;
  const bool vec_in = LEAN || ((a.q_out & 7) == 0 && al16(ar) && (!ar2 || al16(ar2)));
  __half* yr = a.y + (size_t)m * a.ldy;
  const __half* rr = a.residual ? a.residual + (size_t)m * a.ldres : nullptr;
  const int L = 1 << a.log2L;
  const bool vec_out = LEAN || ((a.out_features & 7) == 0 && al16(yr) && (!a.SV || al16(a.SV)) &&
                               (!a.bias || al16(a.bias)) && (!rr || al16(rr)) && ((a.K == 1) || L >= 8));
  const int noct_out = a.out_features >> 3;
  if (a.mma && nt == 512 && vec_in && vec_out) {
    // 4096- and 8192-point output rotations as H_16 factors on the (legacy) tensor path, register resident, ONE exchange
    // through shared memory per 4096 points (fwht4096_frag; decode_step.cu's out_side_m for a standalone linear).  The
    // shuffle form below needs 40 shuffles per thread and two passes through shared memory with two CTA barriers.
    // Block layout in (thread = octet tid of each 4096-point half), spread layout out (4 consecutive halfs per position),
    // same fp16 rounding points as the path below.  8192 = H_2 (x) H_4096: both halves are transformed, then one butterfly.
    const int warp = tid >> 5, lane = tid & 31;
    const int NH = a.mma;                         // 1: q_out = 4096, 2: q_out = 8192
    const HFrag A = make_hfrag(lane);
    float r[2][8];
    // SV / bias / skip connection at this thread's output positions: requested first, they travel during the rotation
    uint2 sv2[2][2], bi2[2][2], re2[2][2];
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
#pragma unroll
      for (int xh = 0; xh < 2; xh++) {
        sv2[hh][xh] = bi2[hh][xh] = re2[hh][xh] = make_uint2(0, 0);
        const int pos = hh * 4096 + idx_spread(warp, lane, xh);
        if (hh < NH && pos < a.out_features) {
          if (a.SV) sv2[hh][xh] = __ldg(reinterpret_cast<const uint2*>(a.SV + pos));
          if (a.bias) bi2[hh][xh] = __ldg(reinterpret_cast<const uint2*>(a.bias + pos));
          if (rr) re2[hh][xh] = __ldcg(reinterpret_cast<const uint2*>(rr + pos));
        }
      }
    }
    QB_DSTAMP(13);
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      if (hh < NH) {
        const size_t o = (size_t)hh * 512 + tid;
        const float4 v0 = __ldcg(reinterpret_cast<const float4*>(ar + o * 8));
        const float4 v1 = __ldcg(reinterpret_cast<const float4*>(ar + o * 8) + 1);
        float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        if (ar2) {
          const float4 w0 = __ldcg(reinterpret_cast<const float4*>(ar2 + o * 8));
          const float4 w1 = __ldcg(reinterpret_cast<const float4*>(ar2 + o * 8) + 1);
          const float r2[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int j = 0; j < 8; j++) f[j] = fmaf(a.resid_scale, r2[j], f[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) f[j] = f16_round(f[j] * xs);                        // origin_order.cu:129
        if (a.wscale_pc) {
          float w8[8];
          unpack_h8(__ldg(reinterpret_cast<const uint4*>(a.wscale_pc) + o), w8);
#pragma unroll
          for (int j = 0; j < 8; j++) f[j] = f16_round(f[j] * w8[j]);                   // qlinear.py:107
        }
        const uint4 oct = pack_h8(f);
        const uint32_t p[4] = {oct.x, oct.z, oct.y, oct.w};
        if (hh) __syncthreads();                                                        // the exchange buffer is reused
        fwht4096_frag(p, A, sm.s, warp, lane, r[hh]);                                   // x 1/64
      }
    }
    if (NH == 2) {                                                                      // x 1/sqrt(2): 1/sqrt(8192) in total
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float u = r[0][j], v = r[1][j];
        r[0][j] = (u + v) * 0.70710678118654752f;
        r[1][j] = (u - v) * 0.70710678118654752f;
      }
    }
    QB_DSTAMP(14);
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      if (hh < NH) {
#pragma unroll
        for (int xh = 0; xh < 2; xh++) {
          const int pos = hh * 4096 + idx_spread(warp, lane, xh);
          if (pos < a.out_features) {
            uint32_t ow[2];
#pragma unroll
            for (int yh = 0; yh < 2; yh++) {
              const int q = xh + 2 * yh;
              __half2 h = __floats2half2_rn(r[hh][2 * q], r[hh][2 * q + 1]);
              if (a.SV) h = __hmul2(h, as_h2(yh ? sv2[hh][xh].y : sv2[hh][xh].x));       // qlinear.py:112
              if (a.bias) h = __hadd2_rn(h, as_h2(yh ? bi2[hh][xh].y : bi2[hh][xh].x));  // qlinear.py:114
              if (rr) {
                float2 v = __half22float2(h);
                const float2 e = __half22float2(as_h2(yh ? re2[hh][xh].y : re2[hh][xh].x));
                v.x += e.x;
                v.y += e.y;
                h = __floats2half2_rn(v.x, v.y);
              }
              ow[yh] = as_u32(h);
            }
            *reinterpret_cast<uint2*>(yr + pos) = make_uint2(ow[0], ow[1]);
          }
        }
      }
    }
    return;
  }
  const bool pre_out = LEAN || (vec_out && noct_out <= nt * CH);   // prefetch: loads independent of the rotation
  uint4 psv[CH], pbv[CH], prv[CH];
  if (pre_out) {
#pragma unroll
    for (int c = 0; c < CH; c++) {
      const int idx = c * nt + tid;
      psv[c] = pbv[c] = prv[c] = make_uint4(0, 0, 0, 0);
      if (idx < noct_out) {
        if (a.SV) psv[c] = *reinterpret_cast<const uint4*>(a.SV + (size_t)idx * 8);
        if (a.bias) pbv[c] = *reinterpret_cast<const uint4*>(a.bias + (size_t)idx * 8);
        if (rr) prv[c] = __ldcg(reinterpret_cast<const uint4*>(rr + (size_t)idx * 8));
      }
    }
  }
  const bool wf = !LEAN && vec_in && a.transform && a.K > 1 && a.log2L == 8 && (nt & 31) == 0;
  const bool wfb = !LEAN && vec_in && a.transform && a.K > 1 && a.log2L >= 9 && a.log2L <= 12 && (nt & 31) == 0;
  const bool wk1 = vec_in && (LEAN || pre_out) && a.transform && a.K == 1 && nt == 512 && (a.log2L == 12 || a.log2L == 13);
  if (vec_in) {
    for (int base = 0; base < noct; base += nt * CH) {
      float4 v0[CH], v1[CH], w0[CH], w1[CH];
#pragma unroll
      for (int c = 0; c < CH; c++) {
        const int idx = base + c * nt + tid;
        v0[c] = v1[c] = w0[c] = w1[c] = make_float4(0, 0, 0, 0);
        if (idx < noct) {
          v0[c] = __ldcg(reinterpret_cast<const float4*>(ar + (size_t)idx * 8));
          v1[c] = __ldcg(reinterpret_cast<const float4*>(ar + (size_t)idx * 8) + 1);
          if (ar2) {
            w0[c] = __ldcg(reinterpret_cast<const float4*>(ar2 + (size_t)idx * 8));
            w1[c] = __ldcg(reinterpret_cast<const float4*>(ar2 + (size_t)idx * 8) + 1);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < CH; c++) {
        const int idx = base + c * nt + tid;
        if (idx < noct) {
          float f[8] = {v0[c].x, v0[c].y, v0[c].z, v0[c].w, v1[c].x, v1[c].y, v1[c].z, v1[c].w};
          if (ar2) {
            const float r[8] = {w0[c].x, w0[c].y, w0[c].z, w0[c].w, w1[c].x, w1[c].y, w1[c].z, w1[c].w};
#pragma unroll
            for (int j = 0; j < 8; j++) f[j] = fmaf(a.resid_scale, r[j], f[j]);
          }
#pragma unroll
          for (int j = 0; j < 8; j++) f[j] = f16_round(f[j] * xs);                      // origin_order.cu:129
          if (a.wscale_pc) {
#pragma unroll
            for (int j = 0; j < 8; j++) f[j] = f16_round(f[j] * __half2float(a.wscale_pc[idx * 8 + j]));   // qlinear.py:107
          }
          if (wf) {
            warp_fwht256(f, tid & 31);
#pragma unroll
            for (int j = 0; j < 8; j++) f[j] *= a.scale;
            *reinterpret_cast<uint4*>(sm.t + t_index(sm, a.K, idx * 8)) = pack_h8(f);
          } else {
            if (wfb || wk1) warp_fwht256(f, tid & 31);
            float4* d = reinterpret_cast<float4*>(sm.s + (wk1 ? wk1_index(idx) : s_index(sm, idx * 8)));
            d[0] = make_float4(f[0], f[1], f[2], f[3]);
            d[1] = make_float4(f[4], f[5], f[6], f[7]);
          }
        }
      }
    }
  } else {
    for (int i = tid; i < a.q_out; i += nt) {
      float v = __ldcg(ar + i);
      if (ar2) v = fmaf(a.resid_scale, __ldcg(ar2 + i), v);
      v = f16_round(v * xs);
      if (a.wscale_pc) v = f16_round(v * __half2float(a.wscale_pc[i]));
      sm.s[s_index(sm, i)] = v;
    }
  }
  QB_DSTAMP(13);
  __syncthreads();
  if (LEAN || (a.K == 1 && pre_out && noct <= nt * CH && (!a.transform || a.log2L >= 3))) {
    const float* fin = sm.s;
    const float sc = a.transform ? a.scale : 1.0f;
    if (wk1) {
      if (LEAN || a.log2L == 12) cross_rows_to_t<8>(sm.s, sm.t, sc, tid);   // (lean instantiations: q <= 4096)
      else cross_rows_to_t<16>(sm.s, sm.t, sc, tid);
      __syncthreads();
    } else if (a.transform) {
      if (LEAN || sm.pp) fin = stockham_hi(sm.s, sm.s2, a.q_out, a.log2L, tid, nt);
      else fwht_hi(sm.s, a.q_out, a.log2L, tid, nt);
    }
    QB_DSTAMP(14);
#pragma unroll
    for (int c = 0; c < CH; c++) {
      const int o = c * nt + tid;
      if (o < noct_out) {
        float f[8], o8[8];
        if (wk1) {
          unpack_h8(*reinterpret_cast<const uint4*>(sm.t + (size_t)o * 8), f);   // already scaled and rounded
        } else {
          final_octet(sm, fin, a.transform, o, f);
#pragma unroll
          for (int j = 0; j < 8; j++) f[j] = f16_round(f[j] * sc);
        }
        if (a.SV) {
          unpack_h8(psv[c], o8);
#pragma unroll
          for (int j = 0; j < 8; j++) f[j] = f16_round(f[j] * o8[j]);
        }
        if (a.bias) {
          unpack_h8(pbv[c], o8);
#pragma unroll
          for (int j = 0; j < 8; j++) f[j] = f16_round(f[j] + o8[j]);
        }
        if (rr) {
          unpack_h8(prv[c], o8);
#pragma unroll
          for (int j = 0; j < 8; j++) f[j] += o8[j];
        }
        *reinterpret_cast<uint4*>(yr + (size_t)o * 8) = pack_h8(f);
      }
    }
    return;
  }
  if (LEAN) return;   // unreachable
  if (wfb) rotate_cross256(sm, a.K, a.scale, tid, nt);
  if (wf || wfb) rotate_mix(sm, a.q_out, a.K, tid, nt);
  else rotate_smem(sm, a.q_out, a.K, a.scale, a.transform, tid, nt);
  QB_DSTAMP(14);

  if (vec_out) {
    for (int base = 0; base < noct_out; base += nt * CH) {
      uint4 sv[CH], bv[CH], rv[CH];
#pragma unroll
      for (int c = 0; c < CH; c++) {
        const int idx = base + c * nt + tid;
        sv[c] = bv[c] = rv[c] = make_uint4(0, 0, 0, 0);
        if (pre_out) {
          sv[c] = psv[c]; bv[c] = pbv[c]; rv[c] = prv[c];
        } else if (idx < noct_out) {
          if (a.SV) sv[c] = *reinterpret_cast<const uint4*>(a.SV + (size_t)idx * 8);
          if (a.bias) bv[c] = *reinterpret_cast<const uint4*>(a.bias + (size_t)idx * 8);
          if (rr) rv[c] = __ldcg(reinterpret_cast<const uint4*>(rr + (size_t)idx * 8));
        }
      }
#pragma unroll
      for (int c = 0; c < CH; c++) {
        const int idx = base + c * nt + tid;
        if (idx < noct_out) {
          float f[8], o8[8];
          unpack_h8(*reinterpret_cast<const uint4*>(sm.t + t_index(sm, a.K, idx * 8)), f);
          if (a.SV) {      // qlinear.py:112
            unpack_h8(sv[c], o8);
#pragma unroll
            for (int j = 0; j < 8; j++) f[j] = f16_round(f[j] * o8[j]);
          }
          if (a.bias) {    // qlinear.py:114
            unpack_h8(bv[c], o8);
#pragma unroll
            for (int j = 0; j < 8; j++) f[j] = f16_round(f[j] + o8[j]);
          }
          if (rr) {        // decoder-layer residual (fusion hook)
            unpack_h8(rv[c], o8);
#pragma unroll
            for (int j = 0; j < 8; j++) f[j] += o8[j];
          }
          *reinterpret_cast<uint4*>(yr + (size_t)idx * 8) = pack_h8(f);
        }
      }
    }
  } else {
    for (int i = tid; i < a.out_features; i += nt) {
      float v = __half2float(sm.t[t_index(sm, a.K, i)]);
      if (a.SV) v = f16_round(v * __half2float(a.SV[i]));
      if (a.bias) v = f16_round(v + __half2float(a.bias[i]));
      if (rr) v = v + __half2float(__ldcg(rr + i));
      yr[i] = __float2half_rn(v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GEMV: integer dp4a against 16-bit fixed-point activations
// ---------------------------------------------------------------------------------------------
struct GemvArgs {
  const unsigned char* qidxs;  // packed codes, row pitch row_bytes
  int64_t row_bytes;
  const void* table;           // E8P: uint2[256]; D4: fp16 [256][4]
  int N, nseg, C, g;           // rows, 8-element segments per row, chunks per row, warps per chunk
  int rows_base, rows_rem;     // CTA b owns rows [b*base + min(b,rem), ...): N = G*base + rem
  int fuse_pro, fuse_epi;
  PrologueArgs pro;            // fuse_pro: computed in-kernel; else pro.xq / pro.xscale are read
  EpilogueArgs epi;            // epi.acc / epi.acc2 are this kernel's outputs
  unsigned int* counters;      // [M] tickets (fuse_epi)
  uint32_t xq_off, rot_off;    // shared-memory byte offsets of the x records / rotation workspace
};

struct GroupArgs {
  long long* dbg;   // optional [ctas][16] clock stamps (tools/timeline.py)
  int phase0;
  int n;
  int cta_begin[QUIPB200_MAX_GROUP + 1];
  GemvArgs a[QUIPB200_MAX_GROUP];
};

// element order inside a 4-byte x word after this permute matches the packed-byte order of the
// E8P decode: bytes (0,2,1,3)
__device__ __forceinline__ uint32_t perm_0213(uint32_t w) { return __byte_perm(w, 0, 0x3120); }

template <int CB>
struct CbTraits;
template <>
struct CbTraits<QUIPB200_CB_E8P12> {
  static constexpr int SEGS = 8;        // 8 codes x 2 B = 16 B per lane
  static constexpr int ACCS = 1;
  static constexpr int TAB_BYTES = 2048;
};
template <>
struct CbTraits<QUIPB200_CB_E8P12RVQ4B> {
  static constexpr int SEGS = 4;        // 4 codes x 4 B
  static constexpr int ACCS = 2;
  static constexpr int TAB_BYTES = 2048;
};
template <>
struct CbTraits<QUIPB200_CB_D4> {
  static constexpr int SEGS = 8;        // 16 codes x 1 B, 2 codes per 8-element segment
  static constexpr int ACCS = 1;
  static constexpr int TAB_BYTES = 1024;
};

// one E8P code against one x segment; accumulates hi/lo planes and the parity correction.
// `absoff` = abs index * 8 (byte offset into the table), `sgn` = sign byte.
__device__ __forceinline__ void e8p_dot(uint32_t absoff, uint32_t sgn, const unsigned char* tab,
                                        const uint32_t (&xs)[4], int xsum, int& aH, int& aL, int& aP) {
  const uint2 t1 = *reinterpret_cast<const uint2*>(tab + absoff);
  uint32_t par;
  const uint2 v = e8p_apply_signs(t1, sgn, par);
  aH = dp4a_ss(v.x, xs[0], aH);
  aH = dp4a_ss(v.y, xs[1], aH);
  aL = dp4a_su(v.x, xs[2], aL);   // signed weights x unsigned low bytes
  aL = dp4a_su(v.y, xs[3], aL);
  aP += (int)par * xsum;          // "- 2 per byte when parity odd" folded out: sum_j x_j
}

// ---------------------------------------------------------------------------------------------
// Replicated lookup tables for the decode (the scheme of decode_step.cu, see the comment there): row i (256 bytes) =
// [abs entry i x 16 copies][sign-mask entry of sign byte i x 16 copies] (D4: [int8x4 entry i x 32 copies][unused]); a
// lane reads its own copy, the address is one PRMT on the packed word, the LDS.64 is conflict-free, and
// table ^ mask is the decoded int8 x 8 word pair (parity shift folded into the mask as ^0x02).
// ---------------------------------------------------------------------------------------------
constexpr int LUT_TAB_BYTES = 65536;
__device__ __forceinline__ uint2 lut_sign_mask(uint32_t s8) {
  const uint32_t par = __popc(s8) & 1u;
  const uint32_t s = s8 ^ par;
  uint2 m;
  m.x = (prmt(s * 0x08040201u, 0u, 0xba98u) & 0xfcfcfcfcu) | (par * 0x02020202u);
  m.y = (prmt(s * 0x80402010u, 0u, 0xba98u) & 0xfcfcfcfcu) | (par * 0x02020202u);
  return m;
}
template <int CB>
__device__ __forceinline__ void lut_build_tables(unsigned char* tab, const void* grid, int tid, int nt) {
  for (int e = tid; e < 256 * 32; e += nt) {
    const int i = e >> 5, c = e & 31;
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(grid) + i);
    if (CB == QUIPB200_CB_D4) {     // fp16 [4] -> int8 (units of 1/2), byte order (0,2,1,3) as the activation records
      const __half2 h01 = *reinterpret_cast<const __half2*>(&t.x), h23 = *reinterpret_cast<const __half2*>(&t.y);
      const int v0 = __float2int_rn(__low2float(h01) * 2.0f) & 0xff, v1 = __float2int_rn(__high2float(h01) * 2.0f) & 0xff;
      const int v2 = __float2int_rn(__low2float(h23) * 2.0f) & 0xff, v3 = __float2int_rn(__high2float(h23) * 2.0f) & 0xff;
      reinterpret_cast<uint32_t*>(tab)[i * 64 + c] = (uint32_t)v0 | ((uint32_t)v2 << 8) | ((uint32_t)v1 << 16) | ((uint32_t)v3 << 24);
    } else {
      uint2 v;
      if (c < 16) v = make_uint2(t.x | 0x01010101u, t.y | 0x01010101u);
      else v = lut_sign_mask((uint32_t)i);
      reinterpret_cast<uint2*>(tab)[e] = v;
    }
  }
}
// Same E8P tables, ~15x fewer instructions: half-row h = 2 i + half per thread -- half 0 replicates codebook entry i, half 1
// the sign mask of sign byte i (computed, no load); 128-bit stores with the chunk order rotated by the lane so a
// quarter-warp covers all 32 banks.  (Used by the lean kernel instantiations only: in the general ones ptxas answers
// this body with a GEMV loop that rematerialises shared-memory addresses, 133 instead of 114 instructions per 8 codes.)
__device__ __forceinline__ void lut_build_tables_e8p_fast(unsigned char* tab, const void* grid, int tid, int nt) {
  for (int h = tid; h < 512; h += nt) {
    const int i = h >> 1, half = h & 1;
    uint2 v;
    if (half) {
      v = lut_sign_mask((uint32_t)i);
    } else {
      const uint2 t = __ldg(reinterpret_cast<const uint2*>(grid) + i);
      v = make_uint2(t.x | 0x01010101u, t.y | 0x01010101u);
    }
    uint4* d = reinterpret_cast<uint4*>(tab + i * 256 + half * 128);
#pragma unroll
    for (int k = 0; k < 8; k++) d[(k + tid) & 7] = make_uint4(v.x, v.y, v.x, v.y);
  }
}
// one E8P code (index bytes AB / SB of packed word w) against one x segment: hi and lo activation planes
template <int AB, int SB>
__device__ __forceinline__ void lut_e8p_dot(uint32_t w, const unsigned char* tab, uint32_t offA, uint32_t offS,
                                            const uint32_t (&xs)[4], int& aH, int& aL) {
  const uint2 t = *reinterpret_cast<const uint2*>(tab + prmt(w, offA, 0x5504u | (AB << 4)));
  const uint2 m = *reinterpret_cast<const uint2*>(tab + prmt(w, offS, 0x5504u | (SB << 4)));
  const uint32_t vx = t.x ^ m.x, vy = t.y ^ m.y;
  aH = dp4a_ss(vx, xs[0], aH);
  aH = dp4a_ss(vy, xs[1], aH);
  aL = dp4a_su(vx, xs[2], aL);   // signed weights x unsigned low bytes
  aL = dp4a_su(vy, xs[3], aL);
}

// both 16-bit codes of a 32-bit word
__device__ __forceinline__ void e8p_dot2(uint32_t w, const unsigned char* tab, const uint32_t (&x0)[4], int s0,
                                         const uint32_t (&x1)[4], int s1, int& aH, int& aL, int& aP) {
  e8p_dot((w >> 5) & 0x7f8u, w & 0xffu, tab, x0, s0, aH, aL, aP);
  e8p_dot((w >> 21) & 0x7f8u, __byte_perm(w, 0, 0x4442), tab, x1, s1, aH, aL, aP);
}

}  // namespace qb
