// Fused QuantLinear.forward for the decode regime (small M) -- the hot path.
//
//   prologue : x*SU -> (hadK^T (x) H_L) -> *wscale/sqrt(L) -> 16-bit fixed point        (1 CTA / row)
//   gemv     : int8-decode(Qidxs) . x_q  as exact integer dp4a on CUDA cores            (1 CTA / SM)
//   epilogue : *scale -> [*Wscale_pc] -> (hadK (x) H_L)/sqrt(L) -> [:out] -> *SV -> +bias (1 CTA / row)
//
// Reference chain replaced: qlinear.py:87-115 -> quant.py:72-88 -> register_lib.py:18-38 ->
// origin_order.cu:388-555 (K1) / fast_hadamard_transform_cuda, 5-9 launches per call.
//
// Why integer arithmetic: every E8P/D4 weight is an odd multiple of 1/4 (resp. 1/2) in [-15/4, 15/4],
// i.e. an int8.  On B200 a 2-bit GEMV is INSTRUCTION bound, not bandwidth bound (6.5 TB/s of int16
// codes = 3.3 T codes/s vs ~34 T thread-instr/s => ~10 instructions per code), so the int8 lattice
// point is never converted to fp16: the activation row is quantised once to 16-bit fixed point
// (hi/lo byte planes) and each code costs 4 dp4a.  The integer dot products are exact; the only
// approximation is the fixed-point activation (|err| <= max|x| * 2^-16 per element, below the fp16
// rounding the reference applies to the same vector).  See DESIGN.md "tolerance".
#include "common.cuh"

namespace qb {

// ---------------------------------------------------------------------------------------------
// options
// ---------------------------------------------------------------------------------------------
int g_opt_table_repl = 16;   // 1: plain 2 KB table, 16: bank-conflict-free replicated table (32 KB)
int g_opt_gemv_warps = 0;    // 0: auto
int g_opt_gemv_ctas_per_sm = 1;
int g_opt_stage_mask = 7;    // bench only: bit0 prologue, bit1 gemv, bit2 epilogue of quipb200_linear_forward

constexpr int PRO_THREADS = 512;
constexpr int GEMV_MAX_WARPS = 24;

// ---------------------------------------------------------------------------------------------
// shared layout helpers for prologue / epilogue
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
  return b;
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

// Rotation: in: s[spad(i)] (fp32, any), out: t[i] (fp16) = round( (hadK' (x) H_L) s * scale ).
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
// prologue
// ---------------------------------------------------------------------------------------------
struct PrologueArgs {
  const __half* x;
  int64_t ldx;
  const __half* SU;
  const __half* hadK;
  int K, in_features, q_in, log2L, transform;
  float scale;
  uint4* xq;       // [M][q_in/8] records {H(0..3), H(4..7), L(0..3), L(4..7)}
  float* xscale;   // [M]
};

__global__ void __launch_bounds__(PRO_THREADS) ql_prologue_kernel(PrologueArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const RotSmem sm = rot_carve(smem_raw, a.q_in, a.K);
  const int tid = threadIdx.x, nt = PRO_THREADS;
  const int m = blockIdx.x;
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
    float v = (tid < nt / 32) ? sm.red[tid] : 0.f;
    v = warp_max(v);
    if (tid == 0) sm.red[0] = v;
  }
  __syncthreads();
  mx = sm.red[0];
  const float inv = (mx > 0.f) ? 32767.0f / mx : 0.f;
  if (tid == 0) a.xscale[m] = (mx > 0.f) ? mx / 32767.0f : 0.f;

  const int nseg = a.q_in >> 3;
  uint4* dst = a.xq + (size_t)m * nseg;
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
}

// ---------------------------------------------------------------------------------------------
// epilogue
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

__global__ void __launch_bounds__(PRO_THREADS) ql_epilogue_kernel(EpilogueArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const RotSmem sm = rot_carve(smem_raw, a.q_out, a.K);
  const int tid = threadIdx.x, nt = PRO_THREADS;
  const int m = blockIdx.x;
  load_hadK(sm, a.hadK, a.K, /*transpose=*/0, tid, nt);
  const float xs = a.xscale[m] * a.unit;
  const float* ar = a.acc + (size_t)m * a.q_out;
  const float* ar2 = a.acc2 ? a.acc2 + (size_t)m * a.q_out : nullptr;
  for (int i = tid; i < a.q_out; i += nt) {
    float v = ar[i];
    if (ar2) v = fmaf(a.resid_scale, ar2[i], v);
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

// ---------------------------------------------------------------------------------------------
// GEMV: integer dp4a against 16-bit fixed-point activations
// ---------------------------------------------------------------------------------------------
struct GemvArgs {
  const unsigned char* qidxs;  // packed codes, row pitch row_bytes
  int64_t row_bytes;
  const void* table;           // E8P: uint2[256]; D4: fp16 [256][4]
  const uint4* xq;             // [M][nseg]
  float* acc;                  // [M][N] exact integer dot products (as float)
  float* acc2;                 // RVQ residual sums or NULL
  int N, nseg, C, g;           // rows, 8-element segments per row, chunks per row, warps per chunk
  int rows_per_cta_max;
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
};
template <>
struct CbTraits<QUIPB200_CB_E8P12RVQ4B> {
  static constexpr int SEGS = 4;        // 4 codes x 4 B
  static constexpr int ACCS = 2;
};
template <>
struct CbTraits<QUIPB200_CB_D4> {
  static constexpr int SEGS = 8;        // 16 codes x 1 B, 2 codes per 8-element segment
  static constexpr int ACCS = 1;
};

// one E8P code against one x segment; accumulates hi/lo planes and the parity correction
template <int REPL>
__device__ __forceinline__ void e8p_dot(uint32_t code16, const unsigned char* tab, uint32_t lane_off,
                                        const uint32_t (&xs)[4], int xsum, int& aH, int& aL, int& aP) {
  const uint32_t absi = code16 >> 8;
  const uint2 t1 = *reinterpret_cast<const uint2*>(tab + (REPL == 16 ? ((absi << 7) | lane_off) : (absi << 3)));
  uint32_t par;
  const uint2 v = e8p_apply_signs(t1, code16 & 0xffu, par);
  aH = dp4a_ss(v.x, xs[0], aH);
  aH = dp4a_ss(v.y, xs[1], aH);
  aL = dp4a_su(v.x, xs[2], aL);   // signed weights x unsigned low bytes
  aL = dp4a_su(v.y, xs[3], aL);
  aP += (int)par * xsum;              // "- 2 per byte when parity odd" folded out: sum_j x_j
}

template <int CB, int REPL, int UNROLL>
__global__ void __launch_bounds__(GEMV_MAX_WARPS * 32, 1) ql_gemv_kernel(GemvArgs a) {
  using T = CbTraits<CB>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [table][red: rows_per_cta_max * C * ACCS ints]
  constexpr int TAB_BYTES = (CB == QUIPB200_CB_D4) ? 1024 : (REPL == 16 ? 32768 : 2048);
  unsigned char* tab = smem_raw;
  int* red = reinterpret_cast<int*>(smem_raw + TAB_BYTES);
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  const int m = blockIdx.y;

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
    if (REPL == 16) {
      for (int i = tid; i < 256 * 16; i += nt) {
        uint2 t = g[i >> 4];
        t.x |= 0x01010101u;
        t.y |= 0x01010101u;
        reinterpret_cast<uint2*>(tab)[i] = t;   // entry e, copy l at byte e*128 + l*8
      }
    } else {
      for (int i = tid; i < 256; i += nt) {
        uint2 t = g[i];
        t.x |= 0x01010101u;
        t.y |= 0x01010101u;
        reinterpret_cast<uint2*>(tab)[i] = t;
      }
    }
  }
  __syncthreads();
  const uint32_t lane_off = (lane & 15) << 3;

  // ---- rows of this CTA ----
  const int G = gridDim.x;
  const int row_begin = (int)(((int64_t)blockIdx.x * a.N) / G);
  const int row_end = (int)(((int64_t)(blockIdx.x + 1) * a.N) / G);
  const int nrows = row_end - row_begin;
  const uint4* xq = a.xq + (size_t)m * a.nseg;

  const int units = a.C * a.g;
  for (int unit = warp; unit < units; unit += nwarps) {
    const int chunk = unit / a.g;
    const int sub = unit - chunk * a.g;
    // this lane's segments: [seg0, seg0 + SEGS)
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
    const unsigned char* colp = a.qidxs + (size_t)(chunk * 32 + lane) * 16;

    for (int r0 = sub; r0 < nrows; r0 += a.g * UNROLL) {
      uint4 cw[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        const int r = r0 + u * a.g;
        cw[u] = make_uint4(0, 0, 0, 0);
        if (lane_valid && r < nrows) cw[u] = ldg_stream_v4(colp + (size_t)(row_begin + r) * a.row_bytes);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        const int r = r0 + u * a.g;
        if (r >= nrows) break;   // warp-uniform
        int aH = 0, aL = 0, aP = 0, bH = 0, bL = 0, bP = 0;
        const uint32_t w[4] = {cw[u].x, cw[u].y, cw[u].z, cw[u].w};
        if (CB == QUIPB200_CB_E8P12) {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            e8p_dot<REPL>(w[i] & 0xffffu, tab, lane_off, xs[2 * i], xsum[2 * i], aH, aL, aP);
            e8p_dot<REPL>(w[i] >> 16, tab, lane_off, xs[2 * i + 1], xsum[2 * i + 1], aH, aL, aP);
          }
        } else if (CB == QUIPB200_CB_E8P12RVQ4B) {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            e8p_dot<REPL>(w[i] >> 16, tab, lane_off, xs[i], xsum[i], aH, aL, aP);      // main  (hi16)
            e8p_dot<REPL>(w[i] & 0xffffu, tab, lane_off, xs[i], xsum[i], bH, bL, bP);  // resid (lo16)
          }
        } else {  // D4: byte c -> 4 weights; two codes per 8-element segment
          const uint32_t* t4 = reinterpret_cast<const uint32_t*>(tab);
#pragma unroll
          for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int b = 0; b < 4; b++) {
              const uint32_t code = (w[i] >> (8 * b)) & 0xffu;
              const uint32_t v = t4[code];
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
  }
  __syncthreads();
  // ---- combine the C chunk partials of every row, coalesced store ----
  for (int r = tid; r < nrows; r += nt) {
    long long s1 = 0, s2 = 0;   // a chunk partial fits int32; a whole row of 28672 may not
    for (int c = 0; c < a.C; c++) {
      s1 += red[(r * a.C + c) * T::ACCS];
      if (T::ACCS == 2) s2 += red[(r * a.C + c) * T::ACCS + 1];
    }
    a.acc[(size_t)m * a.N + row_begin + r] = (float)s1;
    if (T::ACCS == 2) a.acc2[(size_t)m * a.N + row_begin + r] = (float)s2;
  }
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
  int C, g, warps, G, rows_per_cta_max, segs_per_lane, accs;
  int64_t row_bytes;
  size_t smem;
};

static int gemv_plan(int codebook, int N, int K, GemvPlan* p) {
  int segs, accs;
  int64_t row_bytes;
  if (codebook == QUIPB200_CB_E8P12) { segs = 8; accs = 1; row_bytes = (int64_t)K / 8 * 2; }
  else if (codebook == QUIPB200_CB_E8P12RVQ4B) { segs = 4; accs = 2; row_bytes = (int64_t)K / 8 * 4; }
  else if (codebook == QUIPB200_CB_D4) { segs = 8; accs = 1; row_bytes = (int64_t)K / 4; }
  else return QUIPB200_EUNSUPPORTED;
  if (K % 8 != 0 || row_bytes % 16 != 0 || N < 1) return QUIPB200_EUNSUPPORTED;
  const int nseg = K / 8;
  const int lanes = (nseg + segs - 1) / segs;
  p->C = (lanes + 31) / 32;
  p->segs_per_lane = segs;
  p->accs = accs;
  p->row_bytes = row_bytes;
  int wmax = g_opt_gemv_warps > 0 ? g_opt_gemv_warps : 16;
  if (wmax > GEMV_MAX_WARPS) wmax = GEMV_MAX_WARPS;
  if (p->C >= wmax) { p->g = 1; p->warps = wmax; }
  else { p->g = wmax / p->C; p->warps = p->g * p->C; }
  const int sms = quipb200_sm_count();
  int G = sms * (g_opt_gemv_ctas_per_sm > 0 ? g_opt_gemv_ctas_per_sm : 1);
  // keep at least ~2 rows per warp-slot so tiny layers do not launch idle CTAs
  const int min_rows = p->g * 2;
  if ((int64_t)G * min_rows > N) G = (N + min_rows - 1) / min_rows;
  if (G < 1) G = 1;
  p->G = G;
  p->rows_per_cta_max = (N + G - 1) / G + 1;
  const int tab_bytes = (codebook == QUIPB200_CB_D4) ? 1024 : (g_opt_table_repl == 16 ? 32768 : 2048);
  p->smem = (size_t)tab_bytes + (size_t)p->rows_per_cta_max * p->C * accs * sizeof(int);
  if (p->smem > 200 * 1024) return QUIPB200_EUNSUPPORTED;
  return 0;
}

template <int CB, int REPL>
static int launch_gemv_t(const GemvArgs& a, const GemvPlan& p, int M, cudaStream_t st) {
  auto kern = ql_gemv_kernel<CB, REPL, 4>;
  if (p.smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<dim3(p.G, M), p.warps * 32, p.smem, st>>>(a);
  QB_LAUNCH_CHECK();
  return 0;
}

static int launch_gemv(int codebook, const GemvArgs& a, const GemvPlan& p, int M, cudaStream_t st) {
  const bool repl = g_opt_table_repl == 16;
  switch (codebook) {
    case QUIPB200_CB_E8P12:
      return repl ? launch_gemv_t<QUIPB200_CB_E8P12, 16>(a, p, M, st) : launch_gemv_t<QUIPB200_CB_E8P12, 1>(a, p, M, st);
    case QUIPB200_CB_E8P12RVQ4B:
      return repl ? launch_gemv_t<QUIPB200_CB_E8P12RVQ4B, 16>(a, p, M, st)
                  : launch_gemv_t<QUIPB200_CB_E8P12RVQ4B, 1>(a, p, M, st);
    case QUIPB200_CB_D4:
      return launch_gemv_t<QUIPB200_CB_D4, 1>(a, p, M, st);
  }
  return QUIPB200_EUNSUPPORTED;
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

}  // namespace qb

using namespace qb;

extern "C" size_t quipb200_mm_workspace_bytes(int M, int N, int K) {
  if (M < 1 || N < 1 || K < 8) return 0;
  return carve_ws(nullptr, M, N, K, true).bytes;
}

extern "C" int quipb200_mm(int codebook, const void* x, const void* qidxs, const void* grid, float resid_scale,
                           void* out, int M, int N, int K, void* workspace, size_t ws_bytes, void* stream) {
  if (!x || !qidxs || !grid || !out || M < 0 || N < 1 || K < 8) return QUIPB200_EINVAL;
  if (M == 0) return 0;
  if (M > QUIPB200_MM_MAX_M) return QUIPB200_EUNSUPPORTED;
  GemvPlan plan;
  int rc = gemv_plan(codebook, N, K, &plan);
  if (rc) return rc;
  if (!aligned16(x) || !aligned16(qidxs) || !aligned16(grid) || !aligned16(workspace)) return QUIPB200_EALIGN;
  const bool two = plan.accs == 2;
  Workspace ws = carve_ws(workspace, M, N, K, two);
  if (!workspace || ws_bytes < ws.bytes) return QUIPB200_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;

  PrologueArgs pa{};
  pa.x = (const __half*)x; pa.ldx = K; pa.SU = nullptr; pa.hadK = nullptr; pa.K = 1;
  pa.in_features = K; pa.q_in = K; pa.log2L = 0; pa.transform = 0; pa.scale = 1.f;
  pa.xq = ws.xq; pa.xscale = ws.xscale;
  size_t smem = rot_smem_bytes(K, 1);
  if (smem > 220 * 1024) return QUIPB200_EUNSUPPORTED;
  if ((rc = set_smem_attr((const void*)ql_prologue_kernel, smem))) return rc;
  ql_prologue_kernel<<<M, PRO_THREADS, smem, st>>>(pa);
  QB_LAUNCH_CHECK();

  GemvArgs ga{};
  ga.qidxs = (const unsigned char*)qidxs; ga.row_bytes = plan.row_bytes; ga.table = grid;
  ga.xq = ws.xq; ga.acc = ws.acc; ga.acc2 = ws.acc2; ga.N = N; ga.nseg = K / 8; ga.C = plan.C; ga.g = plan.g;
  ga.rows_per_cta_max = plan.rows_per_cta_max;
  if ((rc = launch_gemv(codebook, ga, plan, M, st))) return rc;

  EpilogueArgs ea{};
  ea.acc = ws.acc; ea.acc2 = ws.acc2; ea.xscale = ws.xscale;
  ea.unit = (codebook == QUIPB200_CB_D4) ? 0.5f : 0.25f;
  ea.resid_scale = f16_round_host(resid_scale);
  ea.wscale_pc = nullptr; ea.hadK = nullptr; ea.K = 1; ea.q_out = N; ea.out_features = N; ea.log2L = 0;
  ea.transform = 0; ea.scale = 1.f; ea.SV = nullptr; ea.bias = nullptr; ea.y = (__half*)out; ea.ldy = N;
  smem = rot_smem_bytes(N, 1);
  if (smem > 220 * 1024) return QUIPB200_EUNSUPPORTED;
  if ((rc = set_smem_attr((const void*)ql_epilogue_kernel, smem))) return rc;
  ql_epilogue_kernel<<<M, PRO_THREADS, smem, st>>>(ea);
  QB_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t quipb200_linear_workspace_bytes(const quipb200_linear_t* L, int M) {
  if (!L || M < 1) return 0;
  return carve_ws(nullptr, M, L->q_out, L->q_in, true).bytes;
}

extern "C" int quipb200_linear_forward(const quipb200_linear_t* L, const void* x, int64_t ldx, void* y,
                                       int64_t ldy, int M, void* workspace, size_t ws_bytes, void* stream) {
  if (!L || !x || !y || M < 0 || !L->qidxs || !L->grid) return QUIPB200_EINVAL;
  if (M == 0) return 0;
  if (M > QUIPB200_MM_MAX_M) return QUIPB200_EUNSUPPORTED;
  if (L->K_left < 1 || L->K_right < 1 || L->q_in % L->K_left || L->q_out % L->K_right) return QUIPB200_EINVAL;
  if (L->in_features > L->q_in || L->out_features > L->q_out) return QUIPB200_EINVAL;
  if ((L->K_left > 1 && !L->had_left) || (L->K_right > 1 && !L->had_right)) return QUIPB200_EINVAL;
  const int log2Lin = ilog2_exact(L->q_in / L->K_left), log2Lout = ilog2_exact(L->q_out / L->K_right);
  if (log2Lin < 0 || log2Lout < 0) return QUIPB200_EINVAL;
  GemvPlan plan;
  int rc = gemv_plan(L->codebook, L->q_out, L->q_in, &plan);
  if (rc) return rc;
  if (!aligned16(L->qidxs) || !aligned16(L->grid) || !aligned16(workspace)) return QUIPB200_EALIGN;
  const bool two = plan.accs == 2;
  Workspace ws = carve_ws(workspace, M, L->q_out, L->q_in, two);
  if (!workspace || ws_bytes < ws.bytes) return QUIPB200_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;

  PrologueArgs pa{};
  pa.x = (const __half*)x; pa.ldx = ldx; pa.SU = (const __half*)L->SU; pa.hadK = (const __half*)L->had_left;
  pa.K = L->K_left; pa.in_features = L->in_features; pa.q_in = L->q_in; pa.log2L = log2Lin; pa.transform = 1;
  pa.scale = L->wscale_float / sqrtf((float)(L->q_in / L->K_left));   // quant.py:75
  pa.xq = ws.xq; pa.xscale = ws.xscale;
  size_t smem = rot_smem_bytes(L->q_in, L->K_left);
  if (smem > 220 * 1024) return QUIPB200_EUNSUPPORTED;
  if ((rc = set_smem_attr((const void*)ql_prologue_kernel, smem))) return rc;
  if (g_opt_stage_mask & 1) {
    ql_prologue_kernel<<<M, PRO_THREADS, smem, st>>>(pa);
    QB_LAUNCH_CHECK();
  }

  GemvArgs ga{};
  ga.qidxs = (const unsigned char*)L->qidxs; ga.row_bytes = plan.row_bytes; ga.table = L->grid;
  ga.xq = ws.xq; ga.acc = ws.acc; ga.acc2 = ws.acc2; ga.N = L->q_out; ga.nseg = L->q_in / 8; ga.C = plan.C;
  ga.g = plan.g; ga.rows_per_cta_max = plan.rows_per_cta_max;
  if (g_opt_stage_mask & 2)
    if ((rc = launch_gemv(L->codebook, ga, plan, M, st))) return rc;

  EpilogueArgs ea{};
  ea.acc = ws.acc; ea.acc2 = ws.acc2; ea.xscale = ws.xscale;
  ea.unit = (L->codebook == QUIPB200_CB_D4) ? 0.5f : 0.25f;
  ea.resid_scale = f16_round_host(L->resid_scale);
  ea.wscale_pc = (const __half*)L->wscale_pc; ea.hadK = (const __half*)L->had_right; ea.K = L->K_right;
  ea.q_out = L->q_out; ea.out_features = L->out_features; ea.log2L = log2Lout; ea.transform = 1;
  ea.scale = 1.0f / sqrtf((float)(L->q_out / L->K_right));
  ea.SV = (const __half*)L->SV; ea.bias = (const __half*)L->bias; ea.y = (__half*)y; ea.ldy = ldy;
  smem = rot_smem_bytes(L->q_out, L->K_right);
  if (smem > 220 * 1024) return QUIPB200_EUNSUPPORTED;
  if ((rc = set_smem_attr((const void*)ql_epilogue_kernel, smem))) return rc;
  if (g_opt_stage_mask & 4) {
    ql_epilogue_kernel<<<M, PRO_THREADS, smem, st>>>(ea);
    QB_LAUNCH_CHECK();
  }
  return 0;
}
