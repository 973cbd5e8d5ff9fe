// Fused QuantLinear.forward for the decode regime (small M) -- the hot path.
//
//   ONE kernel per call when the input dim is a power of two (q,k,v,o,gate,up of Llama-2):
//     phase 0  every warp issues the 128-bit loads of its first packed-code rows (HBM latency starts now)
//     phase 1  every CTA redundantly computes x' = H (SU.x) * wscale/sqrt(n) in shared memory and
//              quantises it to 16-bit fixed point (hidden behind the phase-0 loads)
//     phase 2  GEMV: int8-decode(Qidxs) . x_q as exact integer dp4a on CUDA cores, full rows per CTA
//     phase 3  the last CTA to finish (atomic ticket) runs the output side:
//              *scale -> [*Wscale_pc] -> (hadK (x) H)/sqrt(L) -> [:out] -> *SV -> +bias
//   Non power-of-two input dims (down_proj: 11008 = 43 * 256) run the input rotation in a 1-CTA
//   prologue kernel first (2 launches).
//
// Reference chain replaced: qlinear.py:87-115 -> quant.py:72-88 -> register_lib.py:18-38 ->
// origin_order.cu:388-555 (K1) / fast_hadamard_transform_cuda: 5-9 launches per call.
//
// Why integer arithmetic: every E8P/D4 weight is an odd multiple of 1/4 (resp. 1/2) in [-15/4, 15/4],
// i.e. an int8.  On B200 a 2-bit GEMV is INSTRUCTION bound, not bandwidth bound (6.5 TB/s of int16
// codes = 3.3 T codes/s; the ALU and FMA pipes each retire 64 thread-ops/clk/SM => ~10 ops per code
// per pipe at the roofline), so the int8 lattice point is never converted to fp16: the activation row
// is quantised once to 16-bit fixed point (hi/lo byte planes) and each code costs 4 dp4a.  The integer
// dot products are exact; the only approximation is the fixed-point activation (|err| <= max|x| *
// 2^-16 per element, below the fp16 rounding the reference applies to the same vector).
#include "common.cuh"

namespace qb {

// ---------------------------------------------------------------------------------------------
// options (bench / tests)
// ---------------------------------------------------------------------------------------------
int g_opt_table_repl = 1;    // kept for set_option compatibility; the plain 2 KB table is used
int g_opt_gemv_warps = 0;    // 0: auto (16)
int g_opt_gemv_ctas_per_sm = 1;
int g_opt_stage_mask = 7;    // bench only: bit0 prologue, bit1 gemv, bit2 epilogue
int g_opt_fuse = 3;          // bit0: fuse prologue into the GEMV kernel, bit1: last-CTA epilogue

constexpr int PRO_THREADS = 512;
constexpr int GEMV_MAX_WARPS = 16;
constexpr int GEMV_UNROLL = 4;
constexpr int COUNTER_SLOTS = 4096;

// tickets for the last-CTA-done epilogue: zero at module load, reset by the CTA that consumes them.
// One slot per launch, handed out round-robin by the host; launches that share a slot must not
// overlap in time (4096 slots; a captured 7B decode step uses 224).
__device__ unsigned int g_counters[COUNTER_SLOTS * QUIPB200_MM_MAX_M];
static unsigned int g_next_slot = 0;

// ---------------------------------------------------------------------------------------------
// shared layout helpers for the rotations
// ---------------------------------------------------------------------------------------------
struct RotSmem {
  float* s;      // spad(q) floats (butterfly workspace)
  __half* t;     // q halves (rotated vector, fp16-rounded)
  __half* hk;    // K*K halves, laid out [k_in][k_out]
  float* red;    // 32 floats
};

static inline size_t rot_smem_bytes(int q, int K) {
  size_t b = spad_host((size_t)q) * sizeof(float);
  b += ((size_t)q * sizeof(__half) + 15) / 16 * 16;
  b += ((size_t)K * K * sizeof(__half) + 15) / 16 * 16;
  b += 32 * sizeof(float);
  return (b + 15) / 16 * 16;
}

__device__ __forceinline__ RotSmem rot_carve(unsigned char* base, int q, int K) {
  RotSmem r;
  r.s = reinterpret_cast<float*>(base);
  size_t off = ((size_t)(q + ((q >> 6) << 3) + 8)) * sizeof(float);
  r.t = reinterpret_cast<__half*>(base + off);
  off += ((size_t)q * sizeof(__half) + 15) / 16 * 16;
  r.hk = reinterpret_cast<__half*>(base + off);
  off += ((size_t)K * K * sizeof(__half) + 15) / 16 * 16;
  r.red = reinterpret_cast<float*>(base + off);
  return r;
}

// Rotation: in: s[spad(i)] (fp32), out: t[i] (fp16) = round( (hadK' (x) H_L) s * scale ).
// Rounding points follow the reference: fp16 after the FWHT*scale (register_lib.py:20), fp16 after
// hadK@ (quant.py:83).  `hk` holds coef[k_in][k_out].  transform == 0: t = round(s).
__device__ __forceinline__ void rotate_smem(const RotSmem& sm, int q, int K, int log2L, float scale,
                                            int transform, int tid, int nt) {
  if (!transform) {
    for (int i = tid; i < q; i += nt) sm.t[i] = __float2half_rn(sm.s[spad(i)]);
    __syncthreads();
    return;
  }
  fwht_smem(sm.s, q, log2L, 0, tid, nt);
  if (K == 1) {
    for (int i = tid; i < q; i += nt) sm.t[i] = __float2half_rn(sm.s[spad(i)] * scale);
    __syncthreads();
    return;
  }
  for (int i = tid; i < q; i += nt) sm.s[spad(i)] = f16_round(sm.s[spad(i)] * scale);
  __syncthreads();
  const int L = 1 << log2L;
  const int ktiles = (K + 7) >> 3;
  const int ntasks = ktiles << log2L;
  for (int task = tid; task < ntasks; task += nt) {
    const int c = task & (L - 1);
    const int k0 = (task >> log2L) << 3;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
    for (int kp = 0; kp < K; kp++) {
      const float tv = sm.s[spad((kp << log2L) + c)];
      const __half* row = sm.hk + kp * K + k0;
#pragma unroll
      for (int j = 0; j < 8; j++)
        if (k0 + j < K) acc[j] = fmaf(__half2float(row[j]), tv, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (k0 + j < K) sm.t[((k0 + j) << log2L) + c] = __float2half_rn(acc[j]);
  }
  __syncthreads();
}

__device__ __forceinline__ void load_hadK(const RotSmem& sm, const __half* hadK, int K, int transpose,
                                          int tid, int nt) {
  // want coef[k_in][k_out]; y[k_out] = sum_kin M[k_out][k_in] t[k_in], M = hadK (or hadK^T)
  if (K <= 1 || hadK == nullptr) return;
  for (int i = tid; i < K * K; i += nt) {
    const int kin = i / K, kout = i - kin * K;
    sm.hk[i] = transpose ? hadK[kin * K + kout] : hadK[kout * K + kin];
  }
}

// ---------------------------------------------------------------------------------------------
// input side: x*SU -> rotation -> 16-bit fixed point records
// ---------------------------------------------------------------------------------------------
struct PrologueArgs {
  const __half* x;
  int64_t ldx;
  const __half* SU;
  const __half* hadK;
  int K, in_features, q_in, log2L, transform;
  float scale;
  uint4* xq;       // [M][q_in/8] records {H(0..3), H(4..7), L(0..3), L(4..7)}  (global)
  float* xscale;   // [M]
};

// Computes the records into `dst` (shared or global) and returns the fixed-point scale.
__device__ __forceinline__ float prologue_body(const PrologueArgs& a, unsigned char* rot_base, uint4* dst, int m,
                                               int tid, int nt) {
  const RotSmem sm = rot_carve(rot_base, a.q_in, a.K);
  const __half* xr = a.x + (size_t)m * a.ldx;
  load_hadK(sm, a.hadK, a.K, /*transpose=*/1, tid, nt);
  for (int i = tid; i < a.q_in; i += nt) {
    float v = 0.f;
    if (i < a.in_features) {
      v = __half2float(xr[i]);
      if (a.SU) v = f16_round(v * __half2float(a.SU[i]));  // qlinear.py:91 (fp16 tensor op)
    }
    sm.s[spad(i)] = v;
  }
  __syncthreads();
  rotate_smem(sm, a.q_in, a.K, a.log2L, a.scale, a.transform, tid, nt);

  // abs-max -> 16-bit fixed-point scale
  float mx = 0.f;
  for (int i = tid; i < a.q_in; i += nt) mx = fmaxf(mx, fabsf(__half2float(sm.t[i])));
  mx = warp_max(mx);
  if ((tid & 31) == 0) sm.red[tid >> 5] = mx;
  __syncthreads();
  if (tid < 32) {
    float v = (tid < (nt >> 5)) ? sm.red[tid] : 0.f;
    v = warp_max(v);
    if (tid == 0) sm.red[0] = v;
  }
  __syncthreads();
  mx = sm.red[0];
  const float inv = (mx > 0.f) ? 32767.0f / mx : 0.f;

  const int nseg = a.q_in >> 3;
  for (int sgi = tid; sgi < nseg; sgi += nt) {
    int qv[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      int v = __float2int_rn(__half2float(sm.t[sgi * 8 + j]) * inv);
      qv[j] = max(-32767, min(32767, v));
    }
    uint4 r;
    r.x = ((uint32_t)(qv[0] >> 8) & 0xffu) | (((uint32_t)(qv[1] >> 8) & 0xffu) << 8) |
          (((uint32_t)(qv[2] >> 8) & 0xffu) << 16) | (((uint32_t)(qv[3] >> 8) & 0xffu) << 24);
    r.y = ((uint32_t)(qv[4] >> 8) & 0xffu) | (((uint32_t)(qv[5] >> 8) & 0xffu) << 8) |
          (((uint32_t)(qv[6] >> 8) & 0xffu) << 16) | (((uint32_t)(qv[7] >> 8) & 0xffu) << 24);
    r.z = ((uint32_t)qv[0] & 0xffu) | (((uint32_t)qv[1] & 0xffu) << 8) | (((uint32_t)qv[2] & 0xffu) << 16) |
          (((uint32_t)qv[3] & 0xffu) << 24);
    r.w = ((uint32_t)qv[4] & 0xffu) | (((uint32_t)qv[5] & 0xffu) << 8) | (((uint32_t)qv[6] & 0xffu) << 16) |
          (((uint32_t)qv[7] & 0xffu) << 24);
    dst[sgi] = r;
  }
  return (mx > 0.f) ? mx / 32767.0f : 0.f;
}

__global__ void __launch_bounds__(PRO_THREADS) ql_prologue_kernel(PrologueArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int m = blockIdx.x;
  const float xs = prologue_body(a, smem_raw, a.xq + (size_t)m * (a.q_in >> 3), m, threadIdx.x, PRO_THREADS);
  if (threadIdx.x == 0) a.xscale[m] = xs;
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
  __half* y;
  int64_t ldy;
};

__device__ __forceinline__ void epilogue_body(const EpilogueArgs& a, unsigned char* rot_base, int m, float xscale,
                                              int tid, int nt) {
  const RotSmem sm = rot_carve(rot_base, a.q_out, a.K);
  load_hadK(sm, a.hadK, a.K, /*transpose=*/0, tid, nt);
  const float xs = xscale * a.unit;
  const float* ar = a.acc + (size_t)m * a.q_out;
  const float* ar2 = a.acc2 ? a.acc2 + (size_t)m * a.q_out : nullptr;
  for (int i = tid; i < a.q_out; i += nt) {
    float v = __ldcg(ar + i);
    if (ar2) v = fmaf(a.resid_scale, __ldcg(ar2 + i), v);
    v = f16_round(v * xs);                                             // mm output is fp16 (origin_order.cu:129)
    if (a.wscale_pc) v = f16_round(v * __half2float(a.wscale_pc[i]));  // qlinear.py:107
    sm.s[spad(i)] = v;
  }
  __syncthreads();
  rotate_smem(sm, a.q_out, a.K, a.log2L, a.scale, a.transform, tid, nt);
  __half* yr = a.y + (size_t)m * a.ldy;
  for (int i = tid; i < a.out_features; i += nt) {
    float v = __half2float(sm.t[i]);
    if (a.SV) v = f16_round(v * __half2float(a.SV[i]));  // qlinear.py:112
    if (a.bias) v = v + __half2float(a.bias[i]);         // qlinear.py:114
    yr[i] = __float2half_rn(v);
  }
}

__global__ void __launch_bounds__(PRO_THREADS) ql_epilogue_kernel(EpilogueArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  epilogue_body(a, smem_raw, blockIdx.x, a.xscale[blockIdx.x], threadIdx.x, PRO_THREADS);
}

// ---------------------------------------------------------------------------------------------
// GEMV: integer dp4a against 16-bit fixed-point activations
// ---------------------------------------------------------------------------------------------
struct GemvArgs {
  const unsigned char* qidxs;  // packed codes, row pitch row_bytes
  int64_t row_bytes;
  const void* table;           // E8P: uint2[256]; D4: fp16 [256][4]
  int N, nseg, C, g;           // rows, 8-element segments per row, chunks per row, warps per chunk
  int rows_per_cta_max;
  int fuse_pro, fuse_epi;
  PrologueArgs pro;            // fuse_pro: computed in-kernel; else pro.xq / pro.xscale are read
  EpilogueArgs epi;            // epi.acc / epi.acc2 are this kernel's outputs
  unsigned int* counters;      // [M] tickets (fuse_epi)
  uint32_t xq_off, rot_off;    // shared-memory byte offsets of the x records / rotation workspace
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

// both 16-bit codes of a 32-bit word
__device__ __forceinline__ void e8p_dot2(uint32_t w, const unsigned char* tab, const uint32_t (&x0)[4], int s0,
                                         const uint32_t (&x1)[4], int s1, int& aH, int& aL, int& aP) {
  e8p_dot((w >> 5) & 0x7f8u, w & 0xffu, tab, x0, s0, aH, aL, aP);
  e8p_dot((w >> 21) & 0x7f8u, __byte_perm(w, 0, 0x4442), tab, x1, s1, aH, aL, aP);
}

template <int CB>
__global__ void __launch_bounds__(GEMV_MAX_WARPS * 32, 1) ql_gemv_kernel(GemvArgs a) {
  using T = CbTraits<CB>;
  constexpr int UNROLL = GEMV_UNROLL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [table][red: rows_per_cta_max * C * ACCS ints][x records][rotation workspace]
  unsigned char* tab = smem_raw;
  int* red = reinterpret_cast<int*>(smem_raw + T::TAB_BYTES);
  __shared__ float s_xscale;
  __shared__ int s_last;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  const int m = blockIdx.y;

  // ---- rows of this CTA ----
  const int G = gridDim.x;
  const int row_begin = (int)(((int64_t)blockIdx.x * a.N) / G);
  const int row_end = (int)(((int64_t)(blockIdx.x + 1) * a.N) / G);
  const int nrows = row_end - row_begin;
  const int units = a.C * a.g;

  // ---- phase 0: first batch of code loads for this warp's first unit ----
  uint4 cw[UNROLL];
  int unit = warp;
  {
    const int chunk = unit / a.g, sub = unit - chunk * a.g;
    const bool lv = unit < units && (chunk * 32 + lane) * T::SEGS < a.nseg;
    const unsigned char* colp = a.qidxs + (size_t)(chunk * 32 + lane) * 16 + (size_t)row_begin * a.row_bytes;
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const int r = sub + u * a.g;
      cw[u] = make_uint4(0, 0, 0, 0);
      if (lv && r < nrows) cw[u] = ldg_stream_v4(colp + (size_t)r * a.row_bytes);
    }
  }

  // ---- table ----
  if (CB == QUIPB200_CB_D4) {
    // fp16 [256][4] -> int8 (units of 1/2), byte order (0,2,1,3) to match perm_0213'd activations
    const __half* g = reinterpret_cast<const __half*>(a.table);
    for (int i = tid; i < 256; i += nt) {
      int v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = __float2int_rn(__half2float(g[i * 4 + j]) * 2.0f) & 0xff;
      reinterpret_cast<uint32_t*>(tab)[i] = (uint32_t)v[0] | ((uint32_t)v[2] << 8) | ((uint32_t)v[1] << 16) |
                                            ((uint32_t)v[3] << 24);
    }
  } else {
    const uint2* g = reinterpret_cast<const uint2*>(a.table);
    for (int i = tid; i < 256; i += nt) {
      uint2 t = g[i];
      t.x |= 0x01010101u;
      t.y |= 0x01010101u;
      reinterpret_cast<uint2*>(tab)[i] = t;
    }
  }

  // ---- phase 1: activation records ----
  const uint4* xq;
  float xscale;
  if (a.fuse_pro) {
    uint4* xq_s = reinterpret_cast<uint4*>(smem_raw + a.xq_off);
    xscale = prologue_body(a.pro, smem_raw + a.rot_off, xq_s, m, tid, nt);
    xq = xq_s;
    if (!a.fuse_epi && blockIdx.x == 0 && tid == 0) a.pro.xscale[m] = xscale;   // for the epilogue kernel
  } else {
    xq = a.pro.xq + (size_t)m * a.nseg;
    xscale = a.pro.xscale[m];
  }
  __syncthreads();

  // ---- phase 2: GEMV ----
  while (unit < units) {
    const int chunk = unit / a.g;
    const int sub = unit - chunk * a.g;
    const int seg0 = (chunk * 32 + lane) * T::SEGS;
    const bool lane_valid = seg0 < a.nseg;   // row pitch is a multiple of 16 B => all-or-nothing
    uint32_t xs[T::SEGS][4];
    int xsum[T::SEGS];
#pragma unroll
    for (int sgi = 0; sgi < T::SEGS; sgi++) {
      uint4 r = make_uint4(0, 0, 0, 0);
      if (lane_valid) r = xq[seg0 + sgi];
      xs[sgi][0] = perm_0213(r.x);
      xs[sgi][1] = perm_0213(r.y);
      xs[sgi][2] = perm_0213(r.z);
      xs[sgi][3] = perm_0213(r.w);
      const int sh = dp4a_ss(r.x, 0x01010101u, dp4a_ss(r.y, 0x01010101u, 0));
      const int sl = dp4a_su(0x01010101u, r.z, dp4a_su(0x01010101u, r.w, 0));
      xsum[sgi] = sh * 256 + sl;
    }
    const unsigned char* colp = a.qidxs + (size_t)(chunk * 32 + lane) * 16 + (size_t)row_begin * a.row_bytes;

    for (int r0 = sub; r0 < nrows; r0 += a.g * UNROLL) {
      // software pipeline: fetch the next batch before decoding the current one
      uint4 nx[UNROLL];
      const int rn = r0 + a.g * UNROLL;
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        const int r = rn + u * a.g;
        nx[u] = make_uint4(0, 0, 0, 0);
        if (lane_valid && r < nrows) nx[u] = ldg_stream_v4(colp + (size_t)r * a.row_bytes);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        const int r = r0 + u * a.g;
        if (r < nrows) {   // warp-uniform
          int aH = 0, aL = 0, aP = 0, bH = 0, bL = 0, bP = 0;
          const uint32_t w[4] = {cw[u].x, cw[u].y, cw[u].z, cw[u].w};
          if (CB == QUIPB200_CB_E8P12) {
#pragma unroll
            for (int i = 0; i < 4; i++)
              e8p_dot2(w[i], tab, xs[2 * i], xsum[2 * i], xs[2 * i + 1], xsum[2 * i + 1], aH, aL, aP);
          } else if (CB == QUIPB200_CB_E8P12RVQ4B) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
              // main code = hi16, residual code = lo16, same x segment
              e8p_dot((w[i] >> 21) & 0x7f8u, __byte_perm(w[i], 0, 0x4442), tab, xs[i], xsum[i], aH, aL, aP);
              e8p_dot((w[i] >> 5) & 0x7f8u, w[i] & 0xffu, tab, xs[i], xsum[i], bH, bL, bP);
            }
          } else {  // D4: byte c -> 4 weights; two codes per 8-element segment
            const uint32_t* t4 = reinterpret_cast<const uint32_t*>(tab);
#pragma unroll
            for (int i = 0; i < 4; i++) {
#pragma unroll
              for (int b = 0; b < 4; b++) {
                const uint32_t v = t4[(w[i] >> (8 * b)) & 0xffu];
                const int sgi = i * 2 + (b >> 1), half = b & 1;
                aH = dp4a_ss(v, xs[sgi][half], aH);
                aL = dp4a_su(v, xs[sgi][2 + half], aL);
              }
            }
          }
          int tot = aH * 256 + aL - 2 * aP;
          tot = __reduce_add_sync(0xffffffffu, tot);
          int tot2 = 0;
          if (T::ACCS == 2) {
            tot2 = bH * 256 + bL - 2 * bP;
            tot2 = __reduce_add_sync(0xffffffffu, tot2);
          }
          if (lane == 0) {
            red[(r * a.C + chunk) * T::ACCS] = tot;
            if (T::ACCS == 2) red[(r * a.C + chunk) * T::ACCS + 1] = tot2;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; u++) cw[u] = nx[u];
    }
    unit += nwarps;
    if (unit < units) {   // only when there are more (chunk, sub) units than warps (very wide rows)
      const int chunk2 = unit / a.g, sub2 = unit - chunk2 * a.g;
      const bool lv = (chunk2 * 32 + lane) * T::SEGS < a.nseg;
      const unsigned char* colp2 = a.qidxs + (size_t)(chunk2 * 32 + lane) * 16 + (size_t)row_begin * a.row_bytes;
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        const int r = sub2 + u * a.g;
        cw[u] = make_uint4(0, 0, 0, 0);
        if (lv && r < nrows) cw[u] = ldg_stream_v4(colp2 + (size_t)r * a.row_bytes);
      }
    }
  }
  __syncthreads();
  // ---- combine the C chunk partials of every row, coalesced store ----
  float* acc = const_cast<float*>(a.epi.acc);
  float* acc2 = const_cast<float*>(a.epi.acc2);
  for (int r = tid; r < nrows; r += nt) {
    long long s1 = 0, s2 = 0;   // a chunk partial fits int32; a whole row of 28672 may not
    for (int c = 0; c < a.C; c++) {
      s1 += red[(r * a.C + c) * T::ACCS];
      if (T::ACCS == 2) s2 += red[(r * a.C + c) * T::ACCS + 1];
    }
    __stcg(acc + (size_t)m * a.N + row_begin + r, (float)s1);
    if (T::ACCS == 2) __stcg(acc2 + (size_t)m * a.N + row_begin + r, (float)s2);
  }

  // ---- phase 3: the last CTA of this row of the grid runs the output side ----
  if (!a.fuse_epi) return;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int prev = atomicAdd(a.counters + m, 1u);
    s_last = (prev == (unsigned int)(G - 1));
    if (a.fuse_pro) s_xscale = xscale;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid == 0) a.counters[m] = 0;   // ready for the next launch that draws this slot
  epilogue_body(a.epi, smem_raw + a.rot_off, m, a.fuse_pro ? s_xscale : a.pro.xscale[m], tid, nt);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) l++;
  return ((1 << l) == v) ? l : -1;
}

struct GemvPlan {
  int C, g, warps, G, rows_per_cta_max, accs, tab_bytes;
  int64_t row_bytes;
};

static int gemv_plan(int codebook, int N, int K, GemvPlan* p) {
  int segs, accs, tab;
  int64_t row_bytes;
  if (codebook == QUIPB200_CB_E8P12) { segs = 8; accs = 1; tab = 2048; row_bytes = (int64_t)K / 8 * 2; }
  else if (codebook == QUIPB200_CB_E8P12RVQ4B) { segs = 4; accs = 2; tab = 2048; row_bytes = (int64_t)K / 8 * 4; }
  else if (codebook == QUIPB200_CB_D4) { segs = 8; accs = 1; tab = 1024; row_bytes = (int64_t)K / 4; }
  else return QUIPB200_EUNSUPPORTED;
  if (K % 8 != 0 || row_bytes % 16 != 0 || N < 1) return QUIPB200_EUNSUPPORTED;
  const int nseg = K / 8;
  const int lanes = (nseg + segs - 1) / segs;
  p->C = (lanes + 31) / 32;
  p->accs = accs;
  p->tab_bytes = tab;
  p->row_bytes = row_bytes;
  int wmax = g_opt_gemv_warps > 0 ? g_opt_gemv_warps : 16;
  if (wmax > GEMV_MAX_WARPS) wmax = GEMV_MAX_WARPS;
  if (p->C >= wmax) { p->g = 1; p->warps = wmax; }
  else { p->g = wmax / p->C; p->warps = p->g * p->C; }
  const int sms = quipb200_sm_count();
  if (sms < 1) return (int)cudaErrorNoDevice;
  int G = sms * (g_opt_gemv_ctas_per_sm > 0 ? g_opt_gemv_ctas_per_sm : 1);
  // keep at least ~2 rows per warp-slot so tiny layers do not launch idle CTAs
  const int min_rows = p->g * 2;
  if ((int64_t)G * min_rows > N) G = (N + min_rows - 1) / min_rows;
  if (G < 1) G = 1;
  p->G = G;
  p->rows_per_cta_max = (N + G - 1) / G + 1;
  return 0;
}

// workspace carve: [xq: M*nseg*16][xscale: M*4 -> 256 aligned][acc: M*N*4][acc2: M*N*4]
struct Workspace {
  uint4* xq;
  float* xscale;
  float* acc;
  float* acc2;
  size_t bytes;
};
static Workspace carve_ws(void* base, int M, int N, int K, bool two_accs) {
  Workspace w;
  size_t off = 0;
  auto take = [&](size_t b) {
    size_t o = off;
    off += (b + 255) / 256 * 256;
    return o;
  };
  unsigned char* p = reinterpret_cast<unsigned char*>(base);
  w.xq = reinterpret_cast<uint4*>(p + take((size_t)M * (K / 8) * 16));
  w.xscale = reinterpret_cast<float*>(p + take((size_t)M * 4));
  w.acc = reinterpret_cast<float*>(p + take((size_t)M * N * 4));
  w.acc2 = two_accs ? reinterpret_cast<float*>(p + take((size_t)M * N * 4)) : nullptr;
  w.bytes = off;
  return w;
}

static int set_smem_attr(const void* fn, size_t smem) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

static float f16_round_host(float v) { return __half2float(__float2half_rn(v)); }

constexpr size_t SMEM_LIMIT = 220 * 1024;

// Enqueue the whole chain.  `pa`/`ea` describe the two rotations; gemv fields are filled here.
static int run_chain(int codebook, const void* qidxs, const void* grid, int N, int K, int M, PrologueArgs pa,
                     EpilogueArgs ea, void* workspace, size_t ws_bytes, cudaStream_t st) {
  GemvPlan plan;
  int rc = gemv_plan(codebook, N, K, &plan);
  if (rc) return rc;
  const bool two = plan.accs == 2;
  Workspace ws = carve_ws(workspace, M, N, K, two);
  if (!workspace || ws_bytes < ws.bytes) return QUIPB200_EWORKSPACE;
  pa.xq = ws.xq; pa.xscale = ws.xscale;
  ea.acc = ws.acc; ea.acc2 = ws.acc2; ea.xscale = ws.xscale;

  const size_t pro_smem = rot_smem_bytes(pa.q_in, pa.K);
  const size_t epi_smem = rot_smem_bytes(ea.q_out, ea.K);
  if (pro_smem > SMEM_LIMIT || epi_smem > SMEM_LIMIT) return QUIPB200_EUNSUPPORTED;

  // fused prologue only for pure-FWHT (or identity) input sides: the hadK mix is too much work to
  // repeat in every CTA
  const size_t red_bytes = ((size_t)plan.rows_per_cta_max * plan.C * plan.accs * sizeof(int) + 15) / 16 * 16;
  const size_t xq_bytes = (size_t)(K / 8) * 16;
  bool fuse_pro = (g_opt_fuse & 1) && pa.K == 1;
  bool fuse_epi = (g_opt_fuse & 2) != 0;
  size_t smem;
  for (;;) {
    size_t rot = 0;
    if (fuse_pro) rot = pro_smem;
    if (fuse_epi && epi_smem > rot) rot = epi_smem;
    smem = plan.tab_bytes + red_bytes + (fuse_pro ? xq_bytes : 0) + rot;
    if (smem <= SMEM_LIMIT) break;
    if (fuse_epi) fuse_epi = false;          // drop the larger consumer first
    else if (fuse_pro) fuse_pro = false;
    else return QUIPB200_EUNSUPPORTED;
  }

  if (!fuse_pro && (g_opt_stage_mask & 1)) {
    if ((rc = set_smem_attr((const void*)ql_prologue_kernel, pro_smem))) return rc;
    ql_prologue_kernel<<<M, PRO_THREADS, pro_smem, st>>>(pa);
    QB_LAUNCH_CHECK();
  }

  GemvArgs ga{};
  ga.qidxs = (const unsigned char*)qidxs; ga.row_bytes = plan.row_bytes; ga.table = grid;
  ga.N = N; ga.nseg = K / 8; ga.C = plan.C; ga.g = plan.g; ga.rows_per_cta_max = plan.rows_per_cta_max;
  ga.fuse_pro = fuse_pro ? 1 : 0; ga.fuse_epi = fuse_epi ? 1 : 0;
  ga.pro = pa; ga.epi = ea;
  ga.xq_off = (uint32_t)(plan.tab_bytes + red_bytes);
  ga.rot_off = (uint32_t)(plan.tab_bytes + red_bytes + (fuse_pro ? xq_bytes : 0));
  unsigned int* counters = nullptr;
  if (fuse_epi) {
    cudaError_t e = cudaGetSymbolAddress((void**)&counters, g_counters);
    if (e != cudaSuccess) return (int)e;
    counters += (size_t)(__atomic_fetch_add(&g_next_slot, 1u, __ATOMIC_RELAXED) % COUNTER_SLOTS) * QUIPB200_MM_MAX_M;
  }
  ga.counters = counters;
  if (g_opt_stage_mask & 2) {
    const void* fn = nullptr;
    switch (codebook) {
      case QUIPB200_CB_E8P12: fn = (const void*)ql_gemv_kernel<QUIPB200_CB_E8P12>; break;
      case QUIPB200_CB_E8P12RVQ4B: fn = (const void*)ql_gemv_kernel<QUIPB200_CB_E8P12RVQ4B>; break;
      case QUIPB200_CB_D4: fn = (const void*)ql_gemv_kernel<QUIPB200_CB_D4>; break;
      default: return QUIPB200_EUNSUPPORTED;
    }
    if ((rc = set_smem_attr(fn, smem))) return rc;
    void* args[] = {&ga};
    cudaError_t e = cudaLaunchKernel(fn, dim3(plan.G, M), dim3(plan.warps * 32), args, smem, st);
    if (e != cudaSuccess) return (int)e;
    QB_LAUNCH_CHECK();
  }
  if (!fuse_epi && (g_opt_stage_mask & 4)) {
    if ((rc = set_smem_attr((const void*)ql_epilogue_kernel, epi_smem))) return rc;
    ql_epilogue_kernel<<<M, PRO_THREADS, epi_smem, st>>>(ea);
    QB_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace qb

using namespace qb;

extern "C" size_t quipb200_mm_workspace_bytes(int M, int N, int K) {
  if (M < 1 || N < 1 || K < 8) return 0;
  return carve_ws(nullptr, M, N, K, true).bytes;
}

extern "C" int quipb200_mm(int codebook, const void* x, const void* qidxs, const void* grid, float resid_scale,
                           void* out, int M, int N, int K, void* workspace, size_t ws_bytes, void* stream) {
  if (M < 0 || N < 1 || K < 8) return QUIPB200_EINVAL;
  if (M == 0) return 0;
  if (!x || !qidxs || !grid || !out) return QUIPB200_EINVAL;
  if (M > QUIPB200_MM_MAX_M) return QUIPB200_EUNSUPPORTED;
  if (!aligned16(x) || !aligned16(qidxs) || !aligned16(grid) || !aligned16(workspace)) return QUIPB200_EALIGN;
  PrologueArgs pa{};
  pa.x = (const __half*)x; pa.ldx = K; pa.SU = nullptr; pa.hadK = nullptr; pa.K = 1;
  pa.in_features = K; pa.q_in = K; pa.log2L = 0; pa.transform = 0; pa.scale = 1.f;
  EpilogueArgs ea{};
  ea.unit = (codebook == QUIPB200_CB_D4) ? 0.5f : 0.25f;
  ea.resid_scale = f16_round_host(resid_scale);
  ea.wscale_pc = nullptr; ea.hadK = nullptr; ea.K = 1; ea.q_out = N; ea.out_features = N; ea.log2L = 0;
  ea.transform = 0; ea.scale = 1.f; ea.SV = nullptr; ea.bias = nullptr; ea.y = (__half*)out; ea.ldy = N;
  return run_chain(codebook, qidxs, grid, N, K, M, pa, ea, workspace, ws_bytes, (cudaStream_t)stream);
}

extern "C" size_t quipb200_linear_workspace_bytes(const quipb200_linear_t* L, int M) {
  if (!L || M < 1) return 0;
  return carve_ws(nullptr, M, L->q_out, L->q_in, true).bytes;
}

extern "C" int quipb200_linear_forward(const quipb200_linear_t* L, const void* x, int64_t ldx, void* y,
                                       int64_t ldy, int M, void* workspace, size_t ws_bytes, void* stream) {
  if (!L || M < 0) return QUIPB200_EINVAL;
  if (M == 0) return 0;
  if (!x || !y || !L->qidxs || !L->grid) return QUIPB200_EINVAL;
  if (M > QUIPB200_MM_MAX_M) return QUIPB200_EUNSUPPORTED;
  if (L->K_left < 1 || L->K_right < 1 || L->q_in % L->K_left || L->q_out % L->K_right) return QUIPB200_EINVAL;
  if (L->in_features > L->q_in || L->out_features > L->q_out) return QUIPB200_EINVAL;
  if ((L->K_left > 1 && !L->had_left) || (L->K_right > 1 && !L->had_right)) return QUIPB200_EINVAL;
  const int log2Lin = ilog2_exact(L->q_in / L->K_left), log2Lout = ilog2_exact(L->q_out / L->K_right);
  if (log2Lin < 0 || log2Lout < 0) return QUIPB200_EINVAL;
  if (!aligned16(L->qidxs) || !aligned16(L->grid) || !aligned16(workspace)) return QUIPB200_EALIGN;

  PrologueArgs pa{};
  pa.x = (const __half*)x; pa.ldx = ldx; pa.SU = (const __half*)L->SU; pa.hadK = (const __half*)L->had_left;
  pa.K = L->K_left; pa.in_features = L->in_features; pa.q_in = L->q_in; pa.log2L = log2Lin; pa.transform = 1;
  pa.scale = L->wscale_float / sqrtf((float)(L->q_in / L->K_left));   // quant.py:75
  EpilogueArgs ea{};
  ea.unit = (L->codebook == QUIPB200_CB_D4) ? 0.5f : 0.25f;
  ea.resid_scale = f16_round_host(L->resid_scale);
  ea.wscale_pc = (const __half*)L->wscale_pc; ea.hadK = (const __half*)L->had_right; ea.K = L->K_right;
  ea.q_out = L->q_out; ea.out_features = L->out_features; ea.log2L = log2Lout; ea.transform = 1;
  ea.scale = 1.0f / sqrtf((float)(L->q_out / L->K_right));
  ea.SV = (const __half*)L->SV; ea.bias = (const __half*)L->bias; ea.y = (__half*)y; ea.ldy = ldy;
  return run_chain(L->codebook, L->qidxs, L->grid, L->q_out, L->q_in, M, pa, ea, workspace, ws_bytes,
                   (cudaStream_t)stream);
}
