// Whole bs=1 decode step of a Llama-style stack of E8P12 QuantLinears as ONE persistent cooperative
// kernel (quipb200_decode_step, include/quip_b200.h).
//
// Why: on B200 a 4096x4096 2-bit linear is 4 MiB = 0.65 us of HBM time, but a kernel boundary costs
// 2-3 us and the "last CTA rotates the output" hand-off another 2-3 us, so the per-linear launches of
// quantlinear.cu spend ~85 % of a decode step waiting (profiles/README.md, session 2).  Here one CTA
// per SM stays resident for the whole step; a decoder layer is five stages separated by grid-wide
// barriers (a release/acquire counter in L2, ~1 us):
//
//   A  [rot_out(down of the previous layer) + residual] -> RMSNorm -> SU -> rot_in -> q,k,v GEMV
//   B  rot_out slices of q,k,v for one head -> RoPE -> KV append -> split-KV attention partials
//   C  combine partials -> SU -> rot_in -> o_proj GEMV
//   D  rot_out(o) + residual -> RMSNorm -> SU -> rot_in -> gate,up GEMV
//   E  rot_out(gate), rot_out(up) -> silu(gate)*up -> SU -> rot_in (43x256 blocks + mix) -> down GEMV
//
// Every CTA recomputes the (tiny) rotations it needs from the previous stage's raw integer dot
// products, so no stage has a single-CTA serial section, and each warp's first packed-code rows are
// already in flight while the rotation runs.  The rotations here are register-resident restatements
// of quantlinear.cu's input / output sides (one octet per thread, constant-geometry radix-8 passes,
// the first pass fused with the global load) with the SAME arithmetic and rounding points: fixed-point
// activations, exact int32 dp4a, fp16 rounding wherever the reference holds an fp16 tensor.
//
// Reference chain replaced: example_generate.py:29-32 (decode_one_tokens) -> HF LlamaDecoderLayer ->
// 7 x qlinear.py:87-115 per layer.
#include <algorithm>

#include "fwht_mma.cuh"
#include "ql_device.cuh"

namespace qb {

constexpr int DS_THREADS = 512;
constexpr int DS_WARPS = DS_THREADS / 32;
constexpr int DS_UNROLL = 2;
constexpr int DS_MAX_SPLITS = 4;
constexpr int DS_HD = 128;
constexpr int DS_PV_GROUPS = DS_THREADS / 16;   // 16 threads (8 dims each) per cached position in the P.V pass

enum { SL_Q = 0, SL_K, SL_V, SL_O, SL_G, SL_U, SL_D, SL_N };

struct DsWs {            // global scratch (device pointers)
  unsigned int* bar;     // [0] arrival counter, [32] epoch base (zero-filled once by the caller)
  float* xscale;         // [SL_N] fixed-point scale of each linear's input vector
  __half* hA;            // layer input (residual of the attention block)
  __half* hB;            // post-attention hidden (residual of the MLP block)
  __half* acc[SL_N];     // f16(integer dot product * xscale/4) of each linear: the reference's fp16 mm output
  __half* att_o;         // [nh][S][hd] normalised partial attention outputs (fp16: halves the combine stage's traffic)
  float* att_ml;         // [nh][S][2]  running max / sum of each partial
};

// byte offsets into dynamic shared memory.  scr: [A: fp32 n][B: fp32 n][fred: 64 floats][Tg][Tu][hk x 3];
// the attention stage aliases scr from its base.
struct DsSmem { uint32_t tab, red, xq, scr, total, nmax, t_halfs, hk_halfs, mid_halfs; };

struct DsGeom {          // CTAs per member of each GEMV stage (host-computed; every layer has the same shapes)
  int G_A[3], G_C, G_D[2], G_E;
};

struct DsParams {
  quipb200_decode_plan_t plan;
  DsWs ws;
  DsSmem sm;
  DsGeom geo;
  const __half* h_in;
  __half* h_out;
  int kv_splits;
  int use_mma;           // hidden-side rotations are 4096-point: tensor-path transforms
  int flags;             // tuning / A-B switches (option "ds_flags")
  long long* dbg;        // optional [64] clock stamps of one CTA (tools/ds_timeline.py)
  int dbg_cta;
  int dbg_layer;         // which layer's stage stamps are recorded (default 1)
};

// ---------------------------------------------------------------------------------------------
// grid-wide barrier: monotonically increasing arrival counter, release/acquire at gpu scope
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& target, unsigned int nblk) {
  __syncthreads();   // every thread's global writes happen-before thread 0's release (cumulativity)
  if (threadIdx.x == 0) {
    target += nblk;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while ((int)(v - target) < 0);
  }
  __syncthreads();
}

// Split-phase form: arrive (release) first, then issue loads that do not depend on the other CTAs' current stage, then
// wait.  (Requests issued BEFORE the release are waited for by its fence and delay the arrival: measured +1.9 us.)
__device__ __forceinline__ void grid_arrive(unsigned int* counter, unsigned int& target, unsigned int nblk) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblk;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
  }
}
__device__ __forceinline__ void grid_wait(unsigned int* counter, unsigned int target) {
  if (threadIdx.x == 0) {
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while ((int)(v - target) < 0);
  }
  __syncthreads();
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// GEMV over a contiguous row range (same inner loop as ql_gemv_kernel<E8P12>)
// ---------------------------------------------------------------------------------------------
struct GemvCfg {
  const unsigned char* q;
  int64_t row_bytes;
  int nseg, C, g, row_begin, nrows;
  // local row r of the CTA -> matrix row  row_begin + (r / nu) * rstride + r % nu.  Rows are 1 .. 3.5 KB contiguous each, so
  // a CTA may own ANY set of rows at full load efficiency: contiguous ranges (nu = rstride = 1) or, for gate / up, every
  // row k * 256 + c of a few columns c of the K x 256 output blocks (nu columns, rstride = 256) -- then the K x K
  // orthogonal mix over k of the output-side rotation is local to the producing CTA (gemv_store_mix).
  int nu, rstride;
  uint32_t inv_nu;       // ceil(65536 / nu): r / nu == (r * inv_nu) >> 16 for r < 4096, nu <= 8
};
__device__ __forceinline__ int grow(const GemvCfg& c, int r) {
  const int q = (int)(((uint32_t)r * c.inv_nu) >> 16);
  return q * c.rstride + (r - q * c.nu);
}

// work unit -> (column chunk, row phase)
__device__ __forceinline__ void unit_of(const GemvCfg& c, int unit, int& chunk, int& sub) {
  chunk = unit / c.g;
  sub = unit - chunk * c.g;
}

__device__ __forceinline__ int ilog2_dev(int v) { return 31 - __clz(v); }

__device__ __forceinline__ GemvCfg make_cfg(const quipb200_linear_t& L, int bx, int G) {
  GemvCfg c;
  c.q = reinterpret_cast<const unsigned char*>(L.qidxs);
  c.nseg = L.q_in >> 3;
  const int segs = (L.codebook == QUIPB200_CB_E8P12RVQ4B) ? 4 : 8;      // 8-element segments covered by one 16-byte load
  c.row_bytes = (int64_t)c.nseg * ((L.codebook == QUIPB200_CB_E8P12RVQ4B) ? 4 : 2);
  const int lanes = (c.nseg + segs - 1) / segs;
  c.C = (lanes + 31) >> 5;
  c.g = c.C >= DS_WARPS ? 1 : DS_WARPS / c.C;
  // (11008 columns: 6 chunks x 2 row phases = 12 units on 16 warps.  Splitting into 8 row phases per chunk, several units
  // per warp, measured 11 % slower: the stage is bound by instruction issue, not by idle warps.)
  const int base = L.q_out / G, rem = L.q_out % G;
  c.row_begin = bx * base + min(bx, rem);
  c.nrows = base + (bx < rem ? 1 : 0);
  c.nu = 1; c.rstride = 1; c.inv_nu = 65536u;
  return c;
}
// column-unit ownership for a linear whose output side is K blocks of Lb: CTA bx of G owns columns [c0, c0 + nu) of every block
__device__ __forceinline__ GemvCfg make_cfg_cols(const quipb200_linear_t& L, int bx, int G) {
  GemvCfg c = make_cfg(L, bx, G);
  const int Lb = L.q_out / L.K_right;
  const int base = Lb / G, rem = Lb % G;
  c.row_begin = bx * base + min(bx, rem);
  c.nu = base + (bx < rem ? 1 : 0);
  c.nrows = c.nu * L.K_right;
  c.rstride = Lb;
  c.inv_nu = (65536u + (uint32_t)c.nu - 1u) / (uint32_t)c.nu;
  return c;
}

template <int CB>
__device__ __forceinline__ void gemv_first(uint4 (&cw)[DS_UNROLL], const GemvCfg& c, int warp, int lane, uint64_t pol) {
  constexpr int SEGS = CbTraits<CB>::SEGS;
  const int units = c.C * c.g;
  int chunk, sub;
  unit_of(c, warp, chunk, sub);
  const bool lv = warp < units && (chunk * 32 + lane) * SEGS < c.nseg;
  const unsigned char* colp = c.q + (size_t)(chunk * 32 + lane) * 16 + (size_t)c.row_begin * c.row_bytes;
#pragma unroll
  for (int u = 0; u < DS_UNROLL; u++) {
    const int r = sub + u * c.g;
    cw[u] = make_uint4(0, 0, 0, 0);
    if (lv && r < c.nrows) cw[u] = ldg_stream_v4(colp + (size_t)grow(c, r) * c.row_bytes, pol);
  }
}

// pull this CTA's whole row range into L2 (no registers held): used when the stage's prologue is long
__device__ __forceinline__ void gemv_prefetch_l2(const GemvCfg& c, int tid) {
  // runs of nu consecutive rows, rstride rows apart (one run for contiguous ownership: nu = 1 is folded below)
  const int runs = (c.nu == 1 && c.rstride == 1) ? 1 : c.nrows / c.nu;
  const size_t run_bytes = (size_t)(runs == 1 ? c.nrows : c.nu) * c.row_bytes;
  const int lines = (int)((run_bytes + 127) >> 7);
  const unsigned char* base = c.q + (size_t)c.row_begin * c.row_bytes;
  for (int i = tid; i < lines * runs; i += DS_THREADS) {
    const int run = i / lines, li = i - run * lines;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)run * c.rstride * c.row_bytes + (size_t)li * 128));
  }
}

// Packed codes of this CTA's share of a later stage -> L2.  Issued when the current stage's GEMV starts, so the HBM
// fetch of stage X+1 runs under the (compute-bound) GEMV of stage X and never in front of a prologue's own requests.
__device__ __forceinline__ void prefetch_stage(const quipb200_linear_t* const* mem, const int* G, int n, int bid, int tid,
                                               bool cols = false) {
  int j, bx = 0;
  j = -1;
  int begin = 0;
  for (int i = 0; i < n; i++) {
    if (bid >= begin && bid < begin + G[i]) { j = i; bx = bid - begin; }
    begin += G[i];
  }
  if (j < 0) return;
  const GemvCfg c = cols ? make_cfg_cols(*mem[j], bx, G[j]) : make_cfg(*mem[j], bx, G[j]);
  gemv_prefetch_l2(c, tid);
}

// ---------------------------------------------------------------------------------------------
// Replicated lookup tables (64 KB of shared memory, built once per launch).  Row i (256 bytes) holds
//   E8P family: [abs-table entry i, "+1/4" pre-applied, x 16 copies][sign-mask entry of sign byte i x 16 copies]
//   D4        : [int8x4 entry i x 32 copies][unused]
// so a lane reads its own copy: the address is (index byte << 8) | lane offset = ONE PRMT on the packed word, and the
// LDS is bank-conflict free whatever the indices are (a half-warp's 16 lanes hit 16 distinct bank pairs).  The
// sign-mask entry is 0xfc in every negated byte, | 0x02 in every byte when the code's parity is odd: table ^ mask
// negates (the low bits of every table byte are 11) and subtracts the parity shift of 2 quarter-units in one LOP3 per
// word -- codebook/e8p12.py:82-103 restated; round 1 computed the mask with popc / 2 x (imad + prmt) and carried a
// separate parity * sum(x) term (21 instructions per code; 10 now, profiles/r02_gemv_lut_microbench.txt).
// ---------------------------------------------------------------------------------------------
constexpr int DS_TAB_BYTES = 65536;

__device__ __forceinline__ uint2 e8p_sign_mask(uint32_t s8) {
  const uint32_t par = __popc(s8) & 1u;
  const uint32_t s = s8 ^ par;
  uint2 m;
  m.x = (prmt(s * 0x08040201u, 0u, 0xba98u) & 0xfcfcfcfcu) | (par * 0x02020202u);
  m.y = (prmt(s * 0x80402010u, 0u, 0xba98u) & 0xfcfcfcfcu) | (par * 0x02020202u);
  return m;
}

template <int CB>
__device__ __forceinline__ void build_tables(unsigned char* tab, const void* grid, int tid) {
  if (CB == QUIPB200_CB_D4) {
    // fp16 [4] -> int8 (units of 1/2), byte order (0,2,1,3) to match the record layout of the activations
    for (int e = tid; e < 256 * 32; e += DS_THREADS) {
      const int i = e >> 5, c = e & 31;
      const uint2 t = reinterpret_cast<const uint2*>(grid)[i];
      const __half2 h01 = *reinterpret_cast<const __half2*>(&t.x), h23 = *reinterpret_cast<const __half2*>(&t.y);
      const int v0 = __float2int_rn(__low2float(h01) * 2.0f) & 0xff, v1 = __float2int_rn(__high2float(h01) * 2.0f) & 0xff;
      const int v2 = __float2int_rn(__low2float(h23) * 2.0f) & 0xff, v3 = __float2int_rn(__high2float(h23) * 2.0f) & 0xff;
      reinterpret_cast<uint32_t*>(tab)[i * 64 + c] = (uint32_t)v0 | ((uint32_t)v2 << 8) | ((uint32_t)v1 << 16) | ((uint32_t)v3 << 24);
    }
  } else {
    for (int e = tid; e < 256 * 32; e += DS_THREADS) {
      const int i = e >> 5, c = e & 31;
      uint2 v;
      if (c < 16) {
        v = reinterpret_cast<const uint2*>(grid)[i];
        v.x |= 0x01010101u;
        v.y |= 0x01010101u;
      } else {
        v = e8p_sign_mask((uint32_t)i);
      }
      reinterpret_cast<uint2*>(tab)[e] = v;
    }
  }
}

// one E8P code (index bytes `ab` / `sb` of packed word w) against one x segment: hi and lo activation planes
template <int AB, int SB>
__device__ __forceinline__ void e8p_dot_lut(uint32_t w, const unsigned char* tab, uint32_t offA, uint32_t offS,
                                            const uint32_t (&xs)[4], int& aH, int& aL) {
  const uint2 t = *reinterpret_cast<const uint2*>(tab + prmt(w, offA, 0x5504u | (AB << 4)));
  const uint2 m = *reinterpret_cast<const uint2*>(tab + prmt(w, offS, 0x5504u | (SB << 4)));
  const uint32_t vx = t.x ^ m.x, vy = t.y ^ m.y;
  aH = dp4a_ss(vx, xs[0], aH);
  aH = dp4a_ss(vy, xs[1], aH);
  aL = dp4a_su(vx, xs[2], aL);   // signed weights x unsigned low bytes
  aL = dp4a_su(vy, xs[3], aL);
}

// xq: swizzled 16-byte activation records in shared memory; red: [nrows][C] chunk partials
template <int CB>
__device__ __forceinline__ void gemv_run(uint4 (&cw)[DS_UNROLL], const GemvCfg& c, const uint4* xq,
                                         const unsigned char* tab, int* red, int warp, int lane, uint64_t pol) {
  using T = CbTraits<CB>;
  const int units = c.C * c.g;
  const uint32_t offA = (CB == QUIPB200_CB_D4) ? (uint32_t)lane * 4u : (uint32_t)(lane & 15) * 8u;
  const uint32_t offS = 128u + offA;
  int unit = warp;
  while (unit < units) {
    int chunk, sub;
    unit_of(c, unit, chunk, sub);
    const int seg0 = (chunk * 32 + lane) * T::SEGS;
    const bool lane_valid = seg0 < c.nseg;
    uint32_t xs[T::SEGS][4];
#pragma unroll
    for (int sgi = 0; sgi < T::SEGS; sgi++) {
      uint4 r = make_uint4(0, 0, 0, 0);
      if (lane_valid) r = xq[swz(seg0 + sgi)];
      xs[sgi][0] = r.x; xs[sgi][1] = r.y; xs[sgi][2] = r.z; xs[sgi][3] = r.w;
    }
    const unsigned char* colp = c.q + (size_t)(chunk * 32 + lane) * 16 + (size_t)c.row_begin * c.row_bytes;
    for (int r0 = sub; r0 < c.nrows; r0 += c.g * DS_UNROLL) {
      uint4 nx[DS_UNROLL];
      const int rn = r0 + c.g * DS_UNROLL;
#pragma unroll
      for (int u = 0; u < DS_UNROLL; u++) {
        const int r = rn + u * c.g;
        nx[u] = make_uint4(0, 0, 0, 0);
        if (lane_valid && r < c.nrows) nx[u] = ldg_stream_v4(colp + (size_t)grow(c, r) * c.row_bytes, pol);
      }
      int tot[DS_UNROLL], tot2[DS_UNROLL];
#pragma unroll
      for (int u = 0; u < DS_UNROLL; u++) {      // rows past the end hold zero words: harmless, not stored
        int aH = 0, aL = 0, cH = 0, cL = 0;
        const uint32_t w[4] = {cw[u].x, cw[u].y, cw[u].z, cw[u].w};
        if (CB == QUIPB200_CB_E8P12) {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            e8p_dot_lut<1, 0>(w[i], tab, offA, offS, xs[2 * i], aH, aL);
            e8p_dot_lut<3, 2>(w[i], tab, offA, offS, xs[2 * i + 1], cH, cL);
          }
          aH += cH; aL += cL;
        } else if (CB == QUIPB200_CB_E8P12RVQ4B) {     // main code = hi16 (a*), residual code = lo16 (c*), same x segment
#pragma unroll
          for (int i = 0; i < 4; i++) {
            e8p_dot_lut<3, 2>(w[i], tab, offA, offS, xs[i], aH, aL);
            e8p_dot_lut<1, 0>(w[i], tab, offA, offS, xs[i], cH, cL);
          }
        } else {                                        // D4: byte -> 4 weights; two codes per 8-element segment
#pragma unroll
          for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int b = 0; b < 4; b++) {
              const uint32_t v = *reinterpret_cast<const uint32_t*>(tab + prmt(w[i], offA, 0x5504u | (b << 4)));
              const int sgi = i * 2 + (b >> 1), half = b & 1;
              aH = dp4a_ss(v, xs[sgi][half], aH);
              aL = dp4a_su(v, xs[sgi][2 + half], aL);
            }
          }
        }
        tot[u] = aH * 256 + aL;
        tot2[u] = cH * 256 + cL;
      }
#pragma unroll
      for (int u = 0; u < DS_UNROLL; u++) {
        const int r = r0 + u * c.g;
        const int t1 = __reduce_add_sync(0xffffffffu, tot[u]);
        int t2 = 0;
        if (T::ACCS == 2) t2 = __reduce_add_sync(0xffffffffu, tot2[u]);
        if (lane == 0 && r < c.nrows) {
          red[(r * c.C + chunk) * T::ACCS] = t1;
          if (T::ACCS == 2) red[(r * c.C + chunk) * T::ACCS + 1] = t2;
        }
      }
#pragma unroll
      for (int u = 0; u < DS_UNROLL; u++) cw[u] = nx[u];
    }
    unit += DS_WARPS;
    if (unit < units) {
      int chunk2, sub2;
      unit_of(c, unit, chunk2, sub2);
      const bool lv = (chunk2 * 32 + lane) * T::SEGS < c.nseg;
      const unsigned char* colp2 = c.q + (size_t)(chunk2 * 32 + lane) * 16 + (size_t)c.row_begin * c.row_bytes;
#pragma unroll
      for (int u = 0; u < DS_UNROLL; u++) {
        const int r = sub2 + u * c.g;
        cw[u] = make_uint4(0, 0, 0, 0);
        if (lv && r < c.nrows) cw[u] = ldg_stream_v4(colp2 + (size_t)grow(c, r) * c.row_bytes, pol);
      }
    }
  }
}

// chunk partials -> global accumulator (fp32 image of the integer dot product)
template <int CB>
__device__ __forceinline__ void gemv_store(const GemvCfg& c, const int* red, __half* acc, float xscale, float resid_scale,
                                           int tid) {
  using T = CbTraits<CB>;
  __syncthreads();
  const float xs = xscale * (CB == QUIPB200_CB_D4 ? 0.5f : 0.25f);     // weight unit: 1/4 (E8P) or 1/2 (D4)
  const float rs = __half2float(__float2half_rn(resid_scale));         // the reference's fp16 hfma operand (origin_order.cu:378)
  for (int r = tid; r < c.nrows; r += DS_THREADS) {
    long long s = 0, s2 = 0;
    for (int k = 0; k < c.C; k++) {
      s += red[(r * c.C + k) * T::ACCS];
      if (T::ACCS == 2) s2 += red[(r * c.C + k) * T::ACCS + 1];
    }
    float f = (float)s;
    if (T::ACCS == 2) f = fmaf(rs, (float)s2, f);
    acc[c.row_begin + grow(c, r)] = __float2half_rn(f * xs);      // origin_order.cu:129 (single fp16 rounding of the mm)
  }
}

// Column-unit variant for gate / up (make_cfg_cols): the CTA holds every block's row of its nu columns, so it applies the
// K x K orthogonal factor of the output-side rotation itself:  t[k'][c] = f16( sum_k M[k'][k] * v[k][c] ),  v = the fp16 mm
// output (origin_order.cu:129) [* per-channel Wscale, qlinear.py:106-107].  fp16 operands, fp32 accumulate, one fp16
// rounding -- the arithmetic of the reference's `hadK @ y` (quant.py:83); the 256-point transform over c then runs in the
// consuming stage.  (The reference transforms first and mixes second; the two factors of M (x) H commute, the rounding
// point between them moves.)  M: fp16 [Kp][Kp] in shared memory; vbuf: float [nu <= 8][64] scratch.
template <int CB>
__device__ __forceinline__ void gemv_store_mix(const GemvCfg& c, const int* red, __half* acc, float xscale, float resid_scale,
                                               const __half* wpc, const __half* M, int K, float* vbuf, int tid) {
  using T = CbTraits<CB>;
  const int Kp = (K + 15) / 16 * 16;
  if (tid < 8 * (Kp - K)) vbuf[(tid / (Kp - K)) * 64 + K + tid % (Kp - K)] = 0.f;   // the padded columns of M meet zeros
  __syncthreads();
  const float xs = xscale * (CB == QUIPB200_CB_D4 ? 0.5f : 0.25f);
  const float rs = __half2float(__float2half_rn(resid_scale));
  for (int r = tid; r < c.nrows; r += DS_THREADS) {
    long long s = 0, s2 = 0;
    for (int k = 0; k < c.C; k++) {
      s += red[(r * c.C + k) * T::ACCS];
      if (T::ACCS == 2) s2 += red[(r * c.C + k) * T::ACCS + 1];
    }
    float f = (float)s;
    if (T::ACCS == 2) f = fmaf(rs, (float)s2, f);
    f = f16_round(f * xs);
    const int k = (int)(((uint32_t)r * c.inv_nu) >> 16), i = r - k * c.nu;
    if (wpc) f = f16_round(f * __half2float(wpc[k * c.rstride + c.row_begin + i]));
    vbuf[i * 64 + k] = f;
  }
  cp_async_wait_all();      // M was requested when the stage's GEMV started
  __syncthreads();
  const int i = tid >> 6, ko = tid & 63;
  if (i < c.nu && ko < K) {
    // row ko of M as 16-byte loads (row pitch Kp halfs = a multiple of 32 bytes), v[.][i] as broadcast float4 loads: the
    // 2-byte form of this loop ran with 8-way bank conflicts (lane stride 96 bytes) on the critical path of the grid barrier
    float a = 0.f;
    const uint4* mrow = reinterpret_cast<const uint4*>(M + ko * Kp);
    const float4* vrow = reinterpret_cast<const float4*>(vbuf + i * 64);
    for (int k8 = 0; k8 < Kp; k8 += 8) {          // columns K .. Kp-1 of M are zero (padded blob), vbuf is finite there
      const uint4 m8 = mrow[k8 >> 3];
      const float4 v0 = vrow[k8 >> 2], v1 = vrow[(k8 >> 2) + 1];
      const float2 m01 = __half22float2(as_h2(m8.x)), m23 = __half22float2(as_h2(m8.y));
      const float2 m45 = __half22float2(as_h2(m8.z)), m67 = __half22float2(as_h2(m8.w));
      a = fmaf(m01.x, v0.x, a); a = fmaf(m01.y, v0.y, a); a = fmaf(m23.x, v0.z, a); a = fmaf(m23.y, v0.w, a);
      a = fmaf(m45.x, v1.x, a); a = fmaf(m45.y, v1.y, a); a = fmaf(m67.x, v1.z, a); a = fmaf(m67.y, v1.w, a);
    }
    acc[ko * c.rstride + c.row_begin + i] = __float2half_rn(a);
  }
}

// member / index of this CTA inside a GEMV stage whose members own G[0], G[1], .. consecutive CTAs
__device__ __forceinline__ void which_member(const int* G, int n, int bid, int& j, int& bx) {
  j = -1; bx = 0;
  int begin = 0;
  for (int i = 0; i < n; i++) {
    if (bid >= begin && bid < begin + G[i]) { j = i; bx = bid - begin; }
    begin += G[i];
  }
}

// ---------------------------------------------------------------------------------------------
// register-resident rotations: thread t owns octet t (elements 8t .. 8t+7) of a vector of n <= 8*512
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void round_h8(float (&f)[8]) {   // f <- fp32(fp16(f)): an fp16 tensor in the reference
  const uint4 v = pack_h8(f);
  unpack_h8(v, f);
}
__device__ __forceinline__ void mul_round_h8(float (&f)[8], const uint4& hv) {
  float o[8];
  unpack_h8(hv, o);
#pragma unroll
  for (int j = 0; j < 8; j++) f[j] *= o[j];
  round_h8(f);
}
__device__ __forceinline__ void add_round_h8(float (&f)[8], const uint4& hv) {
  float o[8];
  unpack_h8(hv, o);
#pragma unroll
  for (int j = 0; j < 8; j++) f[j] += o[j];
  round_h8(f);
}
__device__ __forceinline__ uint4 ldg_u4(const __half* p, int oct) { return __ldg(reinterpret_cast<const uint4*>(p) + oct); }

struct RotBuf { float* A; float* B; float* fred; };

// (H_n v)[octet tid] with v[i] = f16(acc[i] * xs) [* wscale_pc]: the output-side rotation of a linear whose
// raw dot products are in global memory (qlinear.py:106-109, origin_order.cu:129).  n = 2^p.
__device__ __forceinline__ void rot_out_k1(const __half* acc, const __half* wpc, int n, int p, const RotBuf& rb,
                                           int tid, float (&f)[8]) {
  const int noct = n >> 3;
  const float* fin;
  if (p % 3 == 0 && p >= 6) {
    // first constant-geometry radix-8 pass fused with the (coalesced, strided) global load
    const int sh = p - 3;
    if (tid < noct) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = __half2float(__ldcg(acc + tid + (j << sh)));
      if (wpc) {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = f16_round(v[j] * __half2float(wpc[tid + (j << sh)]));
      }
      butterfly_regs<3>(v);
      float4* d = reinterpret_cast<float4*>(rb.A + tid * 8);
      d[0] = make_float4(v[0], v[1], v[2], v[3]);
      d[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
    float* cur = rb.A;
    float* nxt = rb.B;
    for (int i = 0; i < p / 3 - 2; i++) {
      stockham_pass<3>(cur, nxt, n, p, tid, DS_THREADS);
      __syncthreads();
      float* t = cur; cur = nxt; nxt = t;
    }
    fin = cur;
  } else {
    if (tid < noct) {
      float v[8];
      unpack_h8(__ldcg(reinterpret_cast<const uint4*>(acc) + tid), v);
      if (wpc) mul_round_h8(v, ldg_u4(wpc, tid));
      float4* d = reinterpret_cast<float4*>(rb.A + tid * 8);
      d[0] = make_float4(v[0], v[1], v[2], v[3]);
      d[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
    fin = stockham_hi(rb.A, rb.B, n, p, tid, DS_THREADS);
  }
#pragma unroll
  for (int j = 0; j < 8; j++) f[j] = 0.f;
  if (tid < noct) stockham_last(fin, p, tid, f);
}

// in-register octet -> rotated octet (input side, quant.py:72-84 with K == 1)
__device__ __forceinline__ void rot_in_k1(float (&f)[8], int n, int p, const RotBuf& rb, int tid) {
  const int noct = n >> 3;
  __syncthreads();   // earlier readers of A / B are done
  if (tid < noct) {
    float4* d = reinterpret_cast<float4*>(rb.A + tid * 8);
    d[0] = make_float4(f[0], f[1], f[2], f[3]);
    d[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  __syncthreads();
  const float* fin = stockham_hi(rb.A, rb.B, n, p, tid, DS_THREADS);
  if (tid < noct) stockham_last(fin, p, tid, f);
}

// Output side of linear Lp (K_right == 1): rotation, 1/sqrt(n), SV, bias, residual (qlinear.py:108-114 plus the
// decoder-layer skip connection).  f = octet tid of the fp32 result BEFORE the final fp16 rounding (zeros beyond
// out_features).
__device__ __forceinline__ void out_side_k1(const quipb200_linear_t& Lp, const __half* acc, const __half* resid,
                                            const RotBuf& rb, int tid, float (&f)[8]) {
  const bool vout = tid < (Lp.out_features >> 3);
  const __half* SV = reinterpret_cast<const __half*>(Lp.SV);
  const __half* bias = reinterpret_cast<const __half*>(Lp.bias);
  uint4 svv = make_uint4(0, 0, 0, 0), bv = svv, rv = svv;
  if (vout) {
    if (SV) svv = ldg_u4(SV, tid);
    if (bias) bv = ldg_u4(bias, tid);
    if (resid) rv = __ldcg(reinterpret_cast<const uint4*>(resid) + tid);
  }
  const int p = ilog2_dev(Lp.q_out);
  rot_out_k1(acc, reinterpret_cast<const __half*>(Lp.wscale_pc), Lp.q_out, p, rb, tid, f);
  const float sc = 1.0f / sqrtf((float)Lp.q_out);
#pragma unroll
  for (int j = 0; j < 8; j++) f[j] *= sc;
  round_h8(f);
  if (SV) mul_round_h8(f, svv);
  if (bias) add_round_h8(f, bv);
  if (resid) {
    float o[8];
    unpack_h8(rv, o);
#pragma unroll
    for (int j = 0; j < 8; j++) f[j] += o[j];
  }
  if (!vout) {
#pragma unroll
    for (int j = 0; j < 8; j++) f[j] = 0.f;
  }
}

// Input side of linear L (K_left == 1) from an fp16-valued octet held in registers: [RMSNorm] -> SU -> rotation ->
// * wscale/sqrt(n) -> 16-bit fixed point record in xq (qlinear.py:90-102).  Returns the fixed-point scale.
__device__ __forceinline__ float in_side_k1(float (&f)[8], const __half* norm_w, float eps, const quipb200_linear_t& L,
                                            const RotBuf& rb, uint4* xq, int tid) {
  const bool vin = tid < (L.in_features >> 3);
  const int noct = L.q_in >> 3;
  const __half* SU = reinterpret_cast<const __half*>(L.SU);
  uint4 wv = make_uint4(0, 0, 0, 0), sv = wv;
  if (vin && norm_w) wv = ldg_u4(norm_w, tid);
  if (vin && SU) sv = ldg_u4(SU, tid);
  if (norm_w) {   // LlamaRMSNorm: weight * (x * rstd).to(fp16)
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) ss = fmaf(f[j], f[j], ss);
    ss = block_sum(ss, rb.fred, tid, DS_THREADS);
    const float rstd = rsqrtf(ss / (float)L.in_features + eps);
#pragma unroll
    for (int j = 0; j < 8; j++) f[j] *= rstd;
    round_h8(f);
    mul_round_h8(f, wv);
  }
  if (SU) mul_round_h8(f, sv);
  rot_in_k1(f, L.q_in, ilog2_dev(L.q_in), rb, tid);
  const float scale = L.wscale_float / sqrtf((float)L.q_in);
  float mx = 0.f;
#pragma unroll
  for (int j = 0; j < 8; j++) f[j] *= scale;
  round_h8(f);
  if (tid < noct) {
#pragma unroll
    for (int j = 0; j < 8; j++) mx = fmaxf(mx, fabsf(f[j]));
  }
  mx = block_max1(mx, rb.fred + 32, tid, DS_THREADS);
  const float inv = (mx > 0.f) ? 32767.0f / mx : 0.f;
  if (tid < noct) {
    uint4 r;
    pack_record(f, inv, r);
    xq[swz(tid)] = r;
  }
  __syncthreads();
  return (mx > 0.f) ? mx / 32767.0f : 0.f;
}

// ---------------------------------------------------------------------------------------------
// output / input side of a linear with a 4096-point rotation, tensor-path version (same rounding points
// as out_side_k1 / in_side_k1; values live as register pairs at idx_spread / idx_block positions)
// ---------------------------------------------------------------------------------------------

// Staging area for the small per-layer vectors of a hidden-side stage.  Everything a stage reads from global memory
// is requested with 16-byte cp.async (one request per 8 elements) right after the grid barrier and consumed from shared
// memory in whatever register layout the tensor-path rotations need: on B200 the number of (narrow, same-address)
// requests that 148 CTAs throw at L2 at the same instant is what these stages were bound by.
struct Stg {
  __half* sv;      // [4096] SV of the producing linear
  __half* bias;    // [4096]
  __half* resid;   // [4096] skip connection
  __half* nw;      // [4096] RMSNorm weight
  __half* su;      // [4096] SU of the consuming linear
  __half* atto;    // [n_heads * S * 128] split-KV partial outputs (stage C; aliases sv .. nw)
};
// The staged vectors are read at spread-layout positions, 4 consecutive halfs per lane: stored with the chunk swizzle of
// fwht_mma.cuh (4 wavefronts per 8-byte read instead of 8 per 4-byte read).
__device__ __forceinline__ void stg_vec(__half* dst, const __half* src, int n, int tid) {
  if (src != nullptr && tid * 8 < n) cp_async16(dst + stg_chunk(tid) * 8, src + tid * 8);
}
// halfs i .. i+3 (i % 4 == 0) of a staged vector
__device__ __forceinline__ uint2 stg_ld4(const __half* v, int i) {
  return *reinterpret_cast<const uint2*>(v + stg_chunk(i >> 3) * 8 + (i & 7));
}

// f[2q], f[2q+1]: fp32 result at elements idx_spread(warp, lane, q) (+1) BEFORE the final fp16 rounding.
// The caller has issued stg_vec() for sv / bias / resid of this linear (and whatever the input side needs).
__device__ __forceinline__ uint4 out_side_load(const __half* acc, int warp, int lane) {   // block `warp`, octet `lane`
  return __ldcg(reinterpret_cast<const uint4*>(acc) + warp * 32 + lane);
}
__device__ __forceinline__ void out_side_m(const quipb200_linear_t& Lp, const uint4& oct_in, bool has_resid, const Stg& st,
                                           const HFrag& A, float* S, int warp, int lane, float (&f)[8],
                                           long long* dbg = nullptr) {
  const bool has_sv = Lp.SV != nullptr, has_bias = Lp.bias != nullptr;
  const __half* wpc = reinterpret_cast<const __half*>(Lp.wscale_pc);
  uint4 oct = oct_in;
  if (dbg) dbg[40] = clock64() + (oct.x & 0);
  if (wpc) oct = hmul2x4(oct, __ldg(reinterpret_cast<const uint4*>(wpc) + warp * 32 + lane));   // qlinear.py:107
  const uint32_t p[4] = {oct.x, oct.z, oct.y, oct.w};                                // natural placement (fwht_mma.cuh)
  cp_async_wait_all();                                                               // staged vectors: visible after the
  if (dbg) dbg[41] = clock64();
  fwht4096_frag(p, A, S, warp, lane, f);                                             // exchange barrier inside; x 1/64
  if (dbg) dbg[42] = clock64() + (__float_as_int(f[0]) & 0);
#pragma unroll
  for (int xh = 0; xh < 2; xh++) {
    const int i = idx_spread(warp, lane, xh);                                        // pairs xh and xh + 2: halfs i .. i+3
    const bool ok = i < Lp.out_features;
    uint2 sv = make_uint2(0, 0), bi = sv, re = sv;
    if (has_sv) sv = stg_ld4(st.sv, i);
    if (has_bias) bi = stg_ld4(st.bias, i);
    if (has_resid) re = stg_ld4(st.resid, i);
#pragma unroll
    for (int yh = 0; yh < 2; yh++) {
      const int q = xh + 2 * yh;
      __half2 h = __floats2half2_rn(f[2 * q], f[2 * q + 1]);
      if (has_sv) h = __hmul2(h, as_h2(yh ? sv.y : sv.x));                            // qlinear.py:112
      if (has_bias) h = __hadd2_rn(h, as_h2(yh ? bi.y : bi.x));                          // qlinear.py:114
      float2 v = __half22float2(h);
      if (has_resid) {
        const float2 r = __half22float2(as_h2(yh ? re.y : re.x));
        v.x += r.x;
        v.y += r.y;
      }
      f[2 * q] = ok ? v.x : 0.f;
      f[2 * q + 1] = ok ? v.y : 0.f;
    }
  }
}

// f: fp16-valued pairs at idx_spread positions (zeros beyond in_features): [RMSNorm] -> SU -> rotation -> * wscale/64 ->
// 16-bit fixed point records (each thread ends up holding its own record).  Returns the fixed-point scale.
// The caller has issued stg_vec() for nw / su and guarantees a CTA barrier after this thread's cp.async.wait_all
// (staged == true), or passes staged == false and the function synchronises itself.
__device__ __forceinline__ float in_side_m(const float (&f)[8], bool has_norm, float eps, const quipb200_linear_t& L,
                                           const Stg& st, bool staged, const HFrag& A, float* S, float* fred,
                                           uint4* xq, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  const bool has_su = L.SU != nullptr;
  if (!staged) {
    cp_async_wait_all();
    __syncthreads();
  }
  float rstd = 1.f;
  if (has_norm) {   // LlamaRMSNorm: weight * (x * rstd).to(fp16)
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) ss = fmaf(f[j], f[j], ss);
    ss = block_sum(ss, fred, tid, DS_THREADS);
    rstd = rsqrtf(ss / (float)L.in_features + eps);
  }
  uint32_t p[4];
#pragma unroll
  for (int xh = 0; xh < 2; xh++) {
    const int i = idx_spread(warp, lane, xh);                                        // pairs xh and xh + 2: halfs i .. i+3
    const bool ok = i < L.in_features;
    uint2 nw = make_uint2(0, 0), su = nw;
    if (has_norm) nw = stg_ld4(st.nw, i);
    if (has_su) su = stg_ld4(st.su, i);
#pragma unroll
    for (int yh = 0; yh < 2; yh++) {
      const int q = xh + 2 * yh;
      __half2 h = __floats2half2_rn(f[2 * q] * rstd, f[2 * q + 1] * rstd);
      if (has_norm) h = __hmul2(as_h2(yh ? nw.y : nw.x), h);
      if (has_su) h = __hmul2(h, as_h2(yh ? su.y : su.x));                            // qlinear.py:91
      p[q] = ok ? as_u32(h) : 0u;
    }
  }
  float r[8];
  fwht4096_frag(p, A, S, warp, lane, r);                                             // block layout, 1/64 included
  // block layout with natural placement: this thread now holds octet `tid` of the rotated vector, i.e. its own record
  const float ws = L.wscale_float;
  float mx = 0.f;
  float o8[8];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const __half2 h = __floats2half2_rn(r[2 * q] * ws, r[2 * q + 1] * ws);           // register_lib.py:20 (fp16 out)
    const float2 v = __half22float2(h);
    mx = fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y)));
    const int e = idx_block(0, 0, q);
    o8[e] = v.x;
    o8[e + 1] = v.y;
  }
  mx = block_max1(mx, fred + 32, tid, DS_THREADS);
  const float inv = (mx > 0.f) ? 32767.0f / mx : 0.f;
  uint4 rec;
  pack_record(o8, inv, rec);
  xq[swz(tid)] = rec;
  __syncthreads();
  return (mx > 0.f) ? mx / 32767.0f : 0.f;
}

// ---------------------------------------------------------------------------------------------
// block rotations (K blocks of 256, K x K orthogonal mix): gate / up output side, down input side
// ---------------------------------------------------------------------------------------------
struct BlkBuf { __half* Tg; __half* Tu; __half* hkg; __half* hku; __half* hkd; __half* vSVg; __half* vSVu; __half* vSUd; int LS; };

// In-place K x K mix  T <- M T  of one or two block buffers (256 columns, row stride LS) on the tensor path; both
// buffers share one barrier pair.  fp16 operands, fp32 accumulate, one fp16 rounding: the reference's `hadK @ y`
// (quant.py:83).  MT = Kp / 16.
template <int MT>
__device__ __forceinline__ void mix_tiles_one(__half* T, const __half* hk, int K, int LS, int warp, int lane) {
  constexpr int Kp = MT * 16;
  const int g = lane >> 2, tq = lane & 3;
  for (int nt_i = warp; nt_i < 32; nt_i += DS_WARPS) {
    const int c0 = nt_i << 3;
    uint32_t bf[MT][2];
#pragma unroll
    for (int kt = 0; kt < MT; kt++) {
      int r = kt * 16 + (lane & 15);
      if (r >= K) r = K;                                   // the zero row
      ldmatrix_x2_trans(bf[kt], T + (size_t)r * LS + c0);
    }
    float acc[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
#pragma unroll
      for (int j = 0; j < 4; j++) acc[mt][j] = 0.f;
#pragma unroll
      for (int kt = 0; kt < MT; kt++) {      // A fragments are re-read per tile: keeping all MT*MT of them live spilled
        uint32_t af[4];
        ldmatrix_x4(af, hk + (size_t)(mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * Kp + kt * 16 + (lane >> 4) * 8);
        mma_16816(acc[mt], af, bf[kt]);
      }
    }
    __syncwarp();                                          // every lane has read its B fragments of this tile
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      if (r0 < K) *reinterpret_cast<__half2*>(T + (size_t)r0 * LS + c0 + tq * 2) = __floats2half2_rn(acc[mt][0], acc[mt][1]);
      if (r1 < K) *reinterpret_cast<__half2*>(T + (size_t)r1 * LS + c0 + tq * 2) = __floats2half2_rn(acc[mt][2], acc[mt][3]);
    }
  }
}
template <int MT>
__device__ __forceinline__ void mix_tiles(__half* T0, const __half* hk0, __half* T1, const __half* hk1, int K, int LS, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < LS; i += DS_THREADS) {             // the zero row read in place of rows >= K
    T0[(size_t)K * LS + i] = __float2half_rn(0.f);
    if (T1) T1[(size_t)K * LS + i] = __float2half_rn(0.f);
  }
  __syncthreads();
  mix_tiles_one<MT>(T0, hk0, K, LS, warp, lane);
  if (T1) mix_tiles_one<MT>(T1, hk1, K, LS, warp, lane);
  __syncthreads();
}
__device__ __forceinline__ void mix_blocks(__half* T0, const __half* hk0, __half* T1, const __half* hk1, int K, int LS, int tid) {
  switch ((K + 15) >> 4) {
    case 1: mix_tiles<1>(T0, hk0, T1, hk1, K, LS, tid); break;
    case 2: mix_tiles<2>(T0, hk0, T1, hk1, K, LS, tid); break;
    case 3: mix_tiles<3>(T0, hk0, T1, hk1, K, LS, tid); break;
    default: mix_tiles<4>(T0, hk0, T1, hk1, K, LS, tid); break;
  }
}

// The layer-constant vectors of stage E (SV of gate / up, SU of down, the coefficient blob): requested between the arrival
// at stage D's grid barrier and the wait, so that only the raw dot products remain to be fetched after the barrier.
__device__ __forceinline__ void stage_e_static(const quipb200_linear_t& Lg, const quipb200_linear_t& Lu,
                                               const quipb200_linear_t& Ld, const __half* hk_blob, const BlkBuf& bb, int tid) {
  const int K = Lg.K_right, Kp = (K + 15) / 16 * 16;
  const int noct_mid = Lg.out_features >> 3;
  const __half* SVg = reinterpret_cast<const __half*>(Lg.SV);
  const __half* SVu = reinterpret_cast<const __half*>(Lu.SV);
  const __half* SUd = reinterpret_cast<const __half*>(Ld.SU);
  for (int o = tid; o < noct_mid; o += DS_THREADS) {
    if (SVg) cp_async16(bb.vSVg + o * 8, SVg + o * 8);
    if (SVu) cp_async16(bb.vSVu + o * 8, SVu + o * 8);
    if (SUd) cp_async16(bb.vSUd + o * 8, SUd + o * 8);
  }
  // slot 2 of the layer's coefficient blob: M[k_out][k_in] = had_left^T of down (the input-side mix; slots 0 / 1, the
  // factors of gate / up, are applied by the producing CTAs)
  const int n16 = (Kp * Kp) >> 3;
  const __half* src = hk_blob + (size_t)2 * Kp * Kp;
  for (int i = tid; i < n16; i += DS_THREADS) cp_async16(bb.hkd + i * 8, src + i * 8);
}

// Stage-E input construction for K > 1:  x = rot_in_down( SU_d . silu(out(gate)) * out(up) ), records -> xq.
// Everything the stage reads from global memory is requested up front (one L2 round trip): the raw dot
// products of this warp's blocks into registers, SV_gate / SV_up / SU_down into shared memory with cp.async,
// the three K x K coefficient matrices with batched 2-byte loads.
constexpr int DS_EB = 4;   // blocks per warp (K <= 64: all of them in one round)
__device__ __forceinline__ float stage_e_blocks(const quipb200_linear_t& Lg, const quipb200_linear_t& Lu,
                                                const quipb200_linear_t& Ld, const __half* acc_g, const __half* acc_u,
                                                const __half* hk_blob, const BlkBuf& bb, const HFrag& A,
                                                float* fred, uint4* xq, int tid, bool staged, long long* dbg) {
#define DS_E(i) do { if (dbg) dbg[i] = clock64(); } while (0)
  const int lane = tid & 31, warp = tid >> 5;
  const int K = Lg.K_right, LS = bb.LS;
  const int noct_mid = Lg.out_features >> 3;
  const __half* SVg = reinterpret_cast<const __half*>(Lg.SV);
  const __half* SVu = reinterpret_cast<const __half*>(Lu.SV);
  const __half* SUd = reinterpret_cast<const __half*>(Ld.SU);
  const __half* bg = reinterpret_cast<const __half*>(Lg.bias);
  const __half* bu = reinterpret_cast<const __half*>(Lu.bias);
  if (!staged) stage_e_static(Lg, Lu, Ld, hk_blob, bb, tid);
  DS_E(30);
  // acc_g / acc_u hold t = (M (x) I) v: the producing CTAs applied the K x K factor (gemv_store_mix).  What is left of the
  // output-side rotation is the 256-point transform of every block, so one pass per block takes the dot products all the
  // way to down's block-transformed input: H_256 (x 1/16) -> SV / bias -> silu(gate) * up -> SU -> H_256 (x wscale/16).
  const float ws_d = Ld.wscale_float;
  uint4 ga[DS_EB], ua[DS_EB];      // one 16-byte octet per lane per block; all of the warp's blocks requested up front
#pragma unroll
  for (int r = 0; r < DS_EB; r++) {
    const int b = warp + r * DS_WARPS;
    ga[r] = ua[r] = make_uint4(0, 0, 0, 0);
    if (b < K) {
      ga[r] = __ldcg(reinterpret_cast<const uint4*>(acc_g) + b * 32 + lane);
      ua[r] = __ldcg(reinterpret_cast<const uint4*>(acc_u) + b * 32 + lane);
    }
  }
  cp_async_wait_all();             // staged SV / SU / coefficients: every thread's requests, then the CTA barrier
  __syncthreads();
#pragma unroll
  for (int r = 0; r < DS_EB; r++) {
    const int b = warp + r * DS_WARPS;
    if (b < K) {
      const int o = b * 32 + lane;               // octet of the intermediate vector
      // the lane's octet is already a valid fragment (F(x, y) is a bit permutation of the natural index and H_256 is
      // invariant under it, tools/emu_fwht_frag.py): pairs in, the same octet's transformed pairs out
      const uint32_t pg[4] = {ga[r].x, ga[r].z, ga[r].y, ga[r].w}, pu[4] = {ua[r].x, ua[r].z, ua[r].y, ua[r].w};
      float g[8], u[8];
      fwht256_frag(pg, A, g);
      fwht256_frag(pu, A, u);
      const uint4 g4 = frag_to_octet(g, 1.0f), u4 = frag_to_octet(u, 1.0f);     // fp16: the rotated mm outputs
      uint4 sv_g = make_uint4(0, 0, 0, 0), sv_u = sv_g, su_d = sv_g, b_g = sv_g, b_u = sv_g;
      if (SVg) sv_g = *reinterpret_cast<const uint4*>(bb.vSVg + o * 8);
      if (SVu) sv_u = *reinterpret_cast<const uint4*>(bb.vSVu + o * 8);
      if (SUd) su_d = *reinterpret_cast<const uint4*>(bb.vSUd + o * 8);
      if (bg && o < noct_mid) b_g = __ldg(reinterpret_cast<const uint4*>(bg) + o);
      if (bu && o < noct_mid) b_u = __ldg(reinterpret_cast<const uint4*>(bu) + o);
      const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w}, uw[4] = {u4.x, u4.y, u4.z, u4.w};
      const uint32_t svg[4] = {sv_g.x, sv_g.y, sv_g.z, sv_g.w}, svu[4] = {sv_u.x, sv_u.y, sv_u.z, sv_u.w};
      const uint32_t sud[4] = {su_d.x, su_d.y, su_d.z, su_d.w};
      const uint32_t bgw[4] = {b_g.x, b_g.y, b_g.z, b_g.w}, buw[4] = {b_u.x, b_u.y, b_u.z, b_u.w};
      uint32_t aw[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        __half2 gg = as_h2(gw[q]);
        __half2 uu = as_h2(uw[q]);
        if (SVg) gg = __hmul2(gg, as_h2(svg[q]));
        if (bg) gg = __hadd2_rn(gg, as_h2(bgw[q]));
        if (SVu) uu = __hmul2(uu, as_h2(svu[q]));
        if (bu) uu = __hadd2_rn(uu, as_h2(buw[q]));
        const float2 gf = __half22float2(gg);
        const __half2 sg = __floats2half2_rn(silu_f(gf.x), silu_f(gf.y));
        __half2 a = __hmul2(sg, uu);                                   // LlamaMLP: act_fn(gate) * up
        if (SUd) a = __hmul2(a, as_h2(sud[q]));
        aw[q] = (o < noct_mid) ? as_u32(a) : 0u;
      }
      const uint32_t pa[4] = {aw[0], aw[2], aw[1], aw[3]};
      float w8[8];
      fwht256_frag(pa, A, w8);
      *reinterpret_cast<uint4*>(bb.Tu + b * LS + lane * 8) = frag_to_octet(w8, ws_d);
    }
  }
  DS_E(33);
  mix_blocks(bb.Tu, bb.hkd, nullptr, nullptr, K, LS, tid);
  DS_E(34);
  float mx = 0.f;
  for (int b = warp; b < K; b += DS_WARPS) {
    float u[8];
    unpack_h8(*reinterpret_cast<const uint4*>(bb.Tu + b * LS + lane * 8), u);
#pragma unroll
    for (int j = 0; j < 8; j++) mx = fmaxf(mx, fabsf(u[j]));
  }
  mx = block_max(mx, fred, tid, DS_THREADS);
  const float inv = (mx > 0.f) ? 32767.0f / mx : 0.f;
  for (int b = warp; b < K; b += DS_WARPS) {
    float u[8];
    unpack_h8(*reinterpret_cast<const uint4*>(bb.Tu + b * LS + lane * 8), u);
    uint4 r;
    pack_record(u, inv, r);
    xq[swz(b * 32 + lane)] = r;
  }
  __syncthreads();
  DS_E(35);
#undef DS_E
  return (mx > 0.f) ? mx / 32767.0f : 0.f;
}

// ---------------------------------------------------------------------------------------------
// stage B helpers
// ---------------------------------------------------------------------------------------------
// 128 outputs [128*j, 128*j+128) of H_n f (n = 128*nb, Sylvester order): first the nb blocks are combined
// with the signs of row j of H_nb, then one 128-point transform.  part: [4][128] floats.
// Thread t loads one 16-byte octet: block ih = t / 16, elements 8*(t % 16) .. +7 (n <= 4096: one load per thread).
// Blocks are combined with the signs of row j of H_nb: first the two blocks of a warp (shuffle), then the 16 warps
// through shared memory (part: [16][128] floats, summed by slice_finish).
__device__ __forceinline__ uint4 slice_load(int nb, const __half* acc, int tid) {
  uint4 v = make_uint4(0, 0, 0, 0);
  if ((tid >> 4) < nb) v = __ldcg(reinterpret_cast<const uint4*>(acc) + tid);
  return v;
}
__device__ __forceinline__ void slice_sum(const quipb200_linear_t& L, int nb, const uint4& raw, int j, float* part, int tid) {
  const int ih = tid >> 4, o = tid & 15, warp = tid >> 5;
  const __half* wpc = reinterpret_cast<const __half*>(L.wscale_pc);
  float f[8];
  unpack_h8(raw, f);
  if (wpc && ih < nb) mul_round_h8(f, ldg_u4(wpc, tid));
  const float sg = (ih < nb) ? ((__popc(j & ih) & 1) ? -1.f : 1.f) : 0.f;
#pragma unroll
  for (int e = 0; e < 8; e++) {
    f[e] *= sg;
    f[e] += __shfl_xor_sync(0xffffffffu, f[e], 16);     // the warp's other block
  }
  if ((tid & 16) == 0) {
    float4* d = reinterpret_cast<float4*>(part + warp * 128 + o * 8);
    d[0] = make_float4(f[0], f[1], f[2], f[3]);
    d[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
}

// warp-level 128-point WHT: lane holds elements 4*lane .. 4*lane+3
__device__ __forceinline__ void warp_fwht128(float (&v)[4], int lane) {
  { const float a = v[0], b = v[1], c = v[2], d = v[3]; v[0] = a + b; v[1] = a - b; v[2] = c + d; v[3] = c - d; }
  { const float a = v[0], b = v[1], c = v[2], d = v[3]; v[0] = a + c; v[1] = b + d; v[2] = a - c; v[3] = b - d; }
#pragma unroll
  for (int b = 0; b < 5; b++) {
    const float sg = ((lane >> b) & 1) ? -1.f : 1.f;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float p = __shfl_xor_sync(0xffffffffu, v[j], 1 << b);
      v[j] = fmaf(sg, v[j], p);
    }
  }
}

// finish one head slice in a single warp: transform, output-side scalings of the linear (qlinear.py:108-114).
// svv / bv: the 4 SV / bias halfs of this lane's outputs (j * 128 + 4 * lane ..), loaded by the caller ahead of time.
__device__ __forceinline__ void slice_finish(const quipb200_linear_t& L, const float* part, uint2 svv, uint2 bv, int lane,
                                             float (&v)[4]) {
#pragma unroll
  for (int e = 0; e < 4; e++) v[e] = 0.f;
#pragma unroll
  for (int w = 0; w < DS_WARPS; w++) {
    const float4 t4 = *reinterpret_cast<const float4*>(part + w * 128 + lane * 4);
    v[0] += t4.x; v[1] += t4.y; v[2] += t4.z; v[3] += t4.w;
  }
  warp_fwht128(v, lane);
  const float sc = 1.0f / sqrtf((float)L.q_out);
  const float2 s01 = __half22float2(as_h2(svv.x)), s23 = __half22float2(as_h2(svv.y));
  const float2 b01 = __half22float2(as_h2(bv.x)), b23 = __half22float2(as_h2(bv.y));
  const float sv4[4] = {s01.x, s01.y, s23.x, s23.y}, b4[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
  for (int e = 0; e < 4; e++) {
    v[e] = f16_round(v[e] * sc);
    if (L.SV) v[e] = f16_round(v[e] * sv4[e]);
    if (L.bias) v[e] = f16_round(v[e] + b4[e]);
  }
}

constexpr int DS_ATT_UN = 4;       // positions per half-warp (scores) / per 16-thread group (P.V) in one pass

// Layer-constant vectors of stages D and A (SV / bias of the producing linear, the skip connection written two barriers
// earlier, norm weight, SU of the consuming linear): requested between the arrival at the previous stage's grid barrier
// and the wait (see grid_arrive).
__device__ __forceinline__ void stage_d_static(const DsParams& p, const quipb200_decode_layer_t& Ly, const Stg& stg, int bid, int tid) {
  int j2, bx2;
  which_member(p.geo.G_D, 2, bid, j2, bx2);
  if (j2 < 0) return;
  const quipb200_linear_t& Lp = Ly.o;
  const quipb200_linear_t& Ln = (j2 == 0) ? Ly.gate : Ly.up;
  stg_vec(stg.sv, reinterpret_cast<const __half*>(Lp.SV), Lp.out_features, tid);
  stg_vec(stg.bias, reinterpret_cast<const __half*>(Lp.bias), Lp.out_features, tid);
  stg_vec(stg.resid, p.ws.hA, Lp.out_features, tid);
  stg_vec(stg.nw, reinterpret_cast<const __half*>(Ly.post_norm_w), Ln.in_features, tid);
  stg_vec(stg.su, reinterpret_cast<const __half*>(Ln.SU), Ln.in_features, tid);
}
__device__ __forceinline__ void stage_a_static(const DsParams& p, const quipb200_decode_layer_t& Ly, const quipb200_decode_layer_t& Lnx,
                                               const Stg& stg, int bid, int tid) {
  int j2, bx2;
  which_member(p.geo.G_A, 3, bid, j2, bx2);
  if (j2 < 0) return;
  const quipb200_linear_t& Lp = Ly.down;
  const quipb200_linear_t& Ln = (j2 == 0) ? Lnx.q : (j2 == 1 ? Lnx.k : Lnx.v);
  stg_vec(stg.sv, reinterpret_cast<const __half*>(Lp.SV), Lp.out_features, tid);
  stg_vec(stg.bias, reinterpret_cast<const __half*>(Lp.bias), Lp.out_features, tid);
  stg_vec(stg.resid, p.ws.hB, Lp.out_features, tid);
  stg_vec(stg.nw, reinterpret_cast<const __half*>(Lnx.input_norm_w), Ln.in_features, tid);
  stg_vec(stg.su, reinterpret_cast<const __half*>(Ln.SU), Ln.in_features, tid);
}

template <int CB>
__global__ void __launch_bounds__(DS_THREADS, 1) decode_step_kernel(const __grid_constant__ DsParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = gridDim.x, bid = blockIdx.x;
  const quipb200_decode_plan_t& P = p.plan;
  unsigned char* tab = smem + p.sm.tab;
  int* red = reinterpret_cast<int*>(smem + p.sm.red);
  uint4* xq = reinterpret_cast<uint4*>(smem + p.sm.xq);
  unsigned char* scr = smem + p.sm.scr;
  RotBuf rb;
  rb.A = reinterpret_cast<float*>(scr);
  rb.B = rb.A + p.sm.nmax;
  rb.fred = rb.B + p.sm.nmax;
  BlkBuf bb;
  bb.Tg = reinterpret_cast<__half*>(rb.fred + 64);
  bb.Tu = bb.Tg + p.sm.t_halfs;
  bb.hkg = bb.Tu + p.sm.t_halfs;
  bb.hku = bb.hkg + p.sm.hk_halfs;
  bb.hkd = bb.hku + p.sm.hk_halfs;
  bb.vSVg = bb.hkd + p.sm.hk_halfs;
  bb.vSVu = bb.vSVg + p.sm.mid_halfs;
  bb.vSUd = reinterpret_cast<__half*>(rb.A);               // A | B is idle in stage E when the MLP side is blocked (K > 1)
  bb.LS = 256 + 8;
  const HFrag hfrag = make_hfrag(lane);
  float* const XS = rb.A;                                   // exchange buffer of fwht4096_frag (16 x DS_XROW floats)
  Stg stg;                                                   // staging vectors: alias the block buffers (unused in A / C / D)
  stg.sv = reinterpret_cast<__half*>(rb.fred + 64) + 4096; stg.bias = stg.sv + 4096; stg.resid = stg.bias + 4096; stg.nw = stg.resid + 4096;
  stg.atto = stg.sv;                                          // 32 KB (n_heads * S * 128 halfs <= 16384)
  stg.su = stg.sv + 32768;
  long long* dbg = (p.dbg && bid == p.dbg_cta && tid == 0) ? p.dbg : nullptr;
#define DS_STAMP(i) do { if (dbg) dbg[i] = clock64(); } while (0)
#define DS_ST(i) do { if (dbg && l == p.dbg_layer) dbg[i] = clock64(); } while (0)
  DS_STAMP(0);

  // layer descriptors: two slots in shared memory, layer l+1 is fetched (cp.async) while layer l runs
  __shared__ __align__(16) quipb200_decode_layer_t s_desc[2];
  constexpr int DESC_U2 = (int)(sizeof(quipb200_decode_layer_t) / 8);
  static_assert(sizeof(quipb200_decode_layer_t) % 8 == 0, "descriptor copy granularity");
  if (tid < DESC_U2) reinterpret_cast<uint2*>(&s_desc[0])[tid] = reinterpret_cast<const uint2*>(&P.layers[0])[tid];

  unsigned int bar_target = 0;
  if (tid == 0) bar_target = *reinterpret_cast<volatile unsigned int*>(p.ws.bar + 32);
  const uint64_t pol = l2_evict_first_policy();

  // the codebook tables (identical for every linear)
  build_tables<CB>(tab, P.layers[0].q.grid, tid);
  const int hid8 = P.hidden >> 3;
  if (bid == 0 && tid < hid8) reinterpret_cast<uint4*>(p.ws.hA)[tid] = reinterpret_cast<const uint4*>(p.h_in)[tid];
  __syncthreads();

  {
    const quipb200_linear_t* nx[3] = {&s_desc[0].q, &s_desc[0].k, &s_desc[0].v};
    prefetch_stage(nx, p.geo.G_A, 3, bid, tid);
  }
  int pos = (int)(*P.pos);
  if (pos >= P.max_len) pos = P.max_len - 1;
  if (pos < 0) pos = 0;
  const int S = p.kv_splits;
  const float attn_scale = 1.0f / sqrtf((float)DS_HD);

  for (int l = 0; l < P.n_layers; l++) {
    const quipb200_decode_layer_t& Ly = s_desc[l & 1];
    uint4 cw[DS_UNROLL];
    // ======================= stage A: q, k, v =======================
    {
      int j, bx;
      which_member(p.geo.G_A, 3, bid, j, bx);
      if (j >= 0) {
        const quipb200_linear_t L = (j == 0) ? Ly.q : (j == 1 ? Ly.k : Ly.v);
        const GemvCfg c = make_cfg(L, bx, p.geo.G_A[j]);
        DS_ST(1);
        float f[8];
        float xs;
        if (p.use_mma) {
          if (l == 0) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
              const float2 v = __half22float2(as_h2(ldg_h2(p.h_in, idx_spread(warp, lane, q))));
              f[2 * q] = v.x;
              f[2 * q + 1] = v.y;
            }
          } else {
            const quipb200_linear_t Lp = s_desc[(l - 1) & 1].down;
            if (p.flags & 1) {      // (else staged during the wait of the previous layer's last barrier: stage_a_static)
              stg_vec(stg.sv, reinterpret_cast<const __half*>(Lp.SV), Lp.out_features, tid);
              stg_vec(stg.bias, reinterpret_cast<const __half*>(Lp.bias), Lp.out_features, tid);
              stg_vec(stg.resid, p.ws.hB, Lp.out_features, tid);
              stg_vec(stg.nw, reinterpret_cast<const __half*>(Ly.input_norm_w), L.in_features, tid);
              stg_vec(stg.su, reinterpret_cast<const __half*>(L.SU), L.in_features, tid);
            }
            const uint4 oct = out_side_load(p.ws.acc[SL_D], warp, lane);
            out_side_m(Lp, oct, true, stg, hfrag, XS, warp, lane, f);
#pragma unroll
            for (int q = 0; q < 4; q++) {
              const __half2 h = __floats2half2_rn(f[2 * q], f[2 * q + 1]);
              if (bid == 0) *reinterpret_cast<__half2*>(p.ws.hA + idx_spread(warp, lane, q)) = h;
              const float2 v = __half22float2(h);
              f[2 * q] = v.x;
              f[2 * q + 1] = v.y;
            }
          }
          DS_ST(2);
          if (l == 0) {
            stg_vec(stg.nw, reinterpret_cast<const __half*>(Ly.input_norm_w), L.in_features, tid);
            stg_vec(stg.su, reinterpret_cast<const __half*>(L.SU), L.in_features, tid);
          }
          xs = in_side_m(f, true, P.norm_eps, L, stg, l != 0, hfrag, XS, rb.fred, xq, tid);
        } else {
          if (l == 0) {
            uint4 hv = make_uint4(0, 0, 0, 0);
            if (tid < hid8) hv = reinterpret_cast<const uint4*>(p.h_in)[tid];
            unpack_h8(hv, f);
          } else {
            const quipb200_linear_t Lp = s_desc[(l - 1) & 1].down;
            out_side_k1(Lp, p.ws.acc[SL_D], p.ws.hB, rb, tid, f);
            const uint4 hv = pack_h8(f);
            if (bid == 0 && tid < hid8) reinterpret_cast<uint4*>(p.ws.hA)[tid] = hv;
            unpack_h8(hv, f);
          }
          DS_ST(2);
          xs = in_side_k1(f, reinterpret_cast<const __half*>(Ly.input_norm_w), P.norm_eps, L, rb, xq, tid);
        }
        if (bx == 0 && tid == 0) p.ws.xscale[SL_Q + j] = xs;
        DS_ST(3);
        {
          const quipb200_linear_t* nx[1] = {&Ly.o};
          prefetch_stage(nx, &p.geo.G_C, 1, bid, tid);
        }
        gemv_first<CB>(cw, c, warp, lane, pol);
        gemv_run<CB>(cw, c, xq, tab, red, warp, lane, pol);
        DS_ST(4);
        gemv_store<CB>(c, red, p.ws.acc[SL_Q + j], xs, L.resid_scale, tid);
        DS_ST(5);
      } else {   // idle in this stage, not in the next
        const quipb200_linear_t* nx[1] = {&Ly.o};
        prefetch_stage(nx, &p.geo.G_C, 1, bid, tid);
      }
      grid_barrier(p.ws.bar, bar_target, nblk);
      DS_ST(6);
    }
    // ======================= stage B: attention =======================
    {
      const int nh = P.n_heads, nkv = P.n_kv_heads, group = nh / nkv;
      if (l + 1 < P.n_layers && tid < DESC_U2)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(reinterpret_cast<uint2*>(&s_desc[(l + 1) & 1]) + tid)),
                     "l"(reinterpret_cast<const uint2*>(&P.layers[l + 1]) + tid) : "memory");
      if (bid < nh * S) {
        const int h = bid / S, s = bid - h * S, kvh = h / group;
        // cached positions [0, pos) are split evenly over the first S-1 CTAs of the head; the last one takes only the
        // new token, whose k / v slices it has to rotate first (balances the stage: 3 slices ~ a third of the cache)
        int t_begin, t_end;
        if (S == 1) { t_begin = 0; t_end = pos + 1; }
        else if (s == S - 1) { t_begin = pos; t_end = pos + 1; }
        else {
          const int chunk = (pos + S - 2) / (S - 1);
          t_begin = s * chunk;
          t_end = min(pos, t_begin + chunk);
        }
        const bool has_new = (t_begin <= pos) && (pos < t_end);
        float* part = reinterpret_cast<float*>(scr);          // [3][16][128]
        float* sq = part + 3 * DS_WARPS * 128;                           // [128] rotated, scaled query
        float* sk = sq + 128;                                 // [128] new key (post RoPE)
        float* sv = sk + 128;                                 // [128] new value
        float* sred = sv + 128;                               // [32]
        float* sout = sred + 32;                              // [DS_PV_GROUPS][128]
        float* sc = sout + DS_PV_GROUPS * 128;                // [chunk] scores
        __half* kc = reinterpret_cast<__half*>(Ly.k_cache) + (size_t)kvh * P.max_len * DS_HD;
        __half* vc = reinterpret_cast<__half*>(Ly.v_cache) + (size_t)kvh * P.max_len * DS_HD;
        // Everything this CTA needs that does not depend on the q/k/v stage is requested first and travels while the head
        // slices are rotated: cached K rows of the first scores pass, the RoPE table row, SV / bias of the slices.
        uint4 kfirst[DS_ATT_UN], vfirst[DS_ATT_UN];
#pragma unroll
        for (int u = 0; u < DS_ATT_UN; u++) {
          const int tk = t_begin + warp * 2 + (lane >> 4) + u * DS_WARPS * 2;
          kfirst[u] = vfirst[u] = make_uint4(0, 0, 0, 0);
          if (tk < t_end && tk != pos) {
            kfirst[u] = __ldcg(reinterpret_cast<const uint4*>(kc + (size_t)tk * DS_HD) + (lane & 15));
            vfirst[u] = __ldcg(reinterpret_cast<const uint4*>(vc + (size_t)tk * DS_HD) + (lane & 15));
          }
        }
        uint2 rope_c = make_uint2(0, 0), rope_s = rope_c, sl_sv = rope_c, sl_b = rope_c;
        if (warp < 2) {
          rope_c = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(P.cos_t) + (size_t)pos * DS_HD) + lane);
          rope_s = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(P.sin_t) + (size_t)pos * DS_HD) + lane);
        }
        if (warp < 3) {
          const quipb200_linear_t& Ls = warp == 0 ? Ly.q : (warp == 1 ? Ly.k : Ly.v);
          const int jj = warp == 0 ? h : kvh;
          if (Ls.SV) sl_sv = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(Ls.SV) + jj * 128) + lane);
          if (Ls.bias) sl_b = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(Ls.bias) + jj * 128) + lane);
        }
        {
          const int nbq = Ly.q.q_out >> 7, nbk = Ly.k.q_out >> 7, nbv = Ly.v.q_out >> 7;
          const uint4 aq = slice_load(nbq, p.ws.acc[SL_Q], tid);
          uint4 ak = make_uint4(0, 0, 0, 0), av = ak;
          if (has_new) {
            ak = slice_load(nbk, p.ws.acc[SL_K], tid);
            av = slice_load(nbv, p.ws.acc[SL_V], tid);
          }
          slice_sum(Ly.q, nbq, aq, h, part, tid);
          if (has_new) {
            slice_sum(Ly.k, nbk, ak, kvh, part + DS_WARPS * 128, tid);
            slice_sum(Ly.v, nbv, av, kvh, part + 2 * DS_WARPS * 128, tid);
          }
        }
        __syncthreads();
        DS_ST(7);
        if (warp < (has_new ? 3 : 1)) {
          float v[4];
          const quipb200_linear_t& L = warp == 0 ? Ly.q : (warp == 1 ? Ly.k : Ly.v);
          slice_finish(L, part + warp * DS_WARPS * 128, sl_sv, sl_b, lane, v);
          if (warp < 2) {   // RoPE, HF rotate_half convention: x*cos + rotate_half(x)*sin
            const float2 c01 = __half22float2(as_h2(rope_c.x)), c23 = __half22float2(as_h2(rope_c.y));
            const float2 s01 = __half22float2(as_h2(rope_s.x)), s23 = __half22float2(as_h2(rope_s.y));
            const float cv[4] = {c01.x, c01.y, c23.x, c23.y}, sv4[4] = {s01.x, s01.y, s23.x, s23.y};
            const float sgn = (lane < 16) ? -1.f : 1.f;
            float r4[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const float pv = __shfl_xor_sync(0xffffffffu, v[e], 16);
              r4[e] = f16_round(v[e] * cv[e] + sgn * pv * sv4[e]);
            }
            if (warp == 0) {
              *reinterpret_cast<float4*>(sq + lane * 4) = make_float4(r4[0] * attn_scale, r4[1] * attn_scale, r4[2] * attn_scale, r4[3] * attn_scale);
            } else {
              *reinterpret_cast<float4*>(sk + lane * 4) = make_float4(r4[0], r4[1], r4[2], r4[3]);
              if (h % group == 0) {
                uint2 o;
                o.x = pk_h2(r4[0], r4[1]);
                o.y = pk_h2(r4[2], r4[3]);
                *reinterpret_cast<uint2*>(kc + (size_t)pos * DS_HD + lane * 4) = o;
              }
            }
          } else {
            *reinterpret_cast<float4*>(sv + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
            if (h % group == 0) {
              uint2 o;
              o.x = pk_h2(v[0], v[1]);
              o.y = pk_h2(v[2], v[3]);
              *reinterpret_cast<uint2*>(vc + (size_t)pos * DS_HD + lane * 4) = o;
            }
          }
        }
        __syncthreads();
        DS_ST(8);
        // ---- scores, softmax and P.V in one sweep: a half-warp owns a cached position (lane l16: dims 8 * l16 .. + 7 of its K
        // and V rows, 16-byte loads) and keeps a running (max, sum, weighted V) of its own positions; the 32 half-warp
        // partials meet once in shared memory.  (Round 1 ran scores -> block max -> exp / block sum -> P.V with four block
        // barriers and a second mapping for V.)
        const int l16 = lane & 15, hw = lane >> 4;
        float qv[8];
#pragma unroll
        for (int e = 0; e < 8; e++) qv[e] = sq[l16 * 8 + e];
        float m_run = -INFINITY, l_run = 0.f;
        float oacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        constexpr int UN = DS_ATT_UN;
        for (int tb = t_begin + warp * 2; tb < t_end; tb += DS_WARPS * 2 * UN) {   // warp-uniform trip count (full-mask shuffles inside)
          const int t0 = tb + hw;
          uint4 raw[UN], vr[UN];
          const bool first = tb == t_begin + warp * 2;                              // first pass: K requested at the top of the stage
#pragma unroll
          for (int u = 0; u < UN; u++) {
            const int t = t0 + u * DS_WARPS * 2;
            raw[u] = kfirst[u];
            vr[u] = vfirst[u];
            if (!first && t < t_end && t != pos) {
              raw[u] = __ldcg(reinterpret_cast<const uint4*>(kc + (size_t)t * DS_HD) + l16);
              vr[u] = __ldcg(reinterpret_cast<const uint4*>(vc + (size_t)t * DS_HD) + l16);
            }
          }
          float d[UN];
#pragma unroll
          for (int u = 0; u < UN; u++) {
            const int t = t0 + u * DS_WARPS * 2;
            float kf[8];
            if (t == pos) {
#pragma unroll
              for (int e = 0; e < 8; e++) kf[e] = sk[l16 * 8 + e];
            } else {
              unpack_h8(raw[u], kf);
            }
            float acc = 0.f;
#pragma unroll
            for (int e = 0; e < 8; e++) acc = fmaf(qv[e], kf[e], acc);
            d[u] = acc;
          }
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < UN; u++) d[u] += __shfl_xor_sync(0xffffffffu, d[u], o);
          }
          // online softmax over this half-warp's (up to UN) positions of the pass
          float mx = m_run;
#pragma unroll
          for (int u = 0; u < UN; u++)
            if (t0 + u * DS_WARPS * 2 < t_end) mx = fmaxf(mx, d[u]);
          const float resc = (m_run == -INFINITY) ? 0.f : __expf(m_run - mx);
          l_run *= resc;
#pragma unroll
          for (int e = 0; e < 8; e++) oacc[e] *= resc;
#pragma unroll
          for (int u = 0; u < UN; u++) {
            const int t = t0 + u * DS_WARPS * 2;
            if (t < t_end) {
              const float pr = __expf(d[u] - mx);
              l_run += pr;
              float vf[8];
              if (t == pos) {
#pragma unroll
                for (int e = 0; e < 8; e++) vf[e] = sv[l16 * 8 + e];
              } else {
                unpack_h8(vr[u], vf);
              }
#pragma unroll
              for (int e = 0; e < 8; e++) oacc[e] = fmaf(pr, vf[e], oacc[e]);
            }
          }
          m_run = mx;
        }
        // the warp's two half-warps merge through shuffles, then warp partials -> shared memory: sout[16][128] weighted V,
        // sc[16][2] (max, sum)
        {
          const float m_o = __shfl_xor_sync(0xffffffffu, m_run, 16), l_o = __shfl_xor_sync(0xffffffffu, l_run, 16);
          const float mw = fmaxf(m_run, m_o);
          const float wa = (m_run == -INFINITY) ? 0.f : __expf(m_run - mw), wb = (m_o == -INFINITY) ? 0.f : __expf(m_o - mw);
#pragma unroll
          for (int e = 0; e < 8; e++) {
            const float o_o = __shfl_xor_sync(0xffffffffu, oacc[e], 16);
            oacc[e] = wa * oacc[e] + wb * o_o;
          }
          if (hw == 0) {
            float4* so = reinterpret_cast<float4*>(sout + warp * 128 + l16 * 8);
            so[0] = make_float4(oacc[0], oacc[1], oacc[2], oacc[3]);
            so[1] = make_float4(oacc[4], oacc[5], oacc[6], oacc[7]);
            if (l16 == 0) {
              sc[2 * warp] = mw;
              sc[2 * warp + 1] = wa * l_run + wb * l_o;
            }
          }
        }
        __syncthreads();
        DS_ST(9);
        if (tid < DS_HD) {
          float mx = -INFINITY;
#pragma unroll
          for (int g2 = 0; g2 < DS_WARPS; g2++) mx = fmaxf(mx, sc[2 * g2]);
          float tot = 0.f, r = 0.f;
#pragma unroll
          for (int g2 = 0; g2 < DS_WARPS; g2++) {
            const float mg = sc[2 * g2];
            const float w = (mg == -INFINITY) ? 0.f : __expf(mg - mx);
            tot = fmaf(w, sc[2 * g2 + 1], tot);
            r = fmaf(w, sout[g2 * 128 + tid], r);
          }
          p.ws.att_o[(size_t)(h * S + s) * DS_HD + tid] = __float2half_rn(tot > 0.f ? r / tot : 0.f);
          if (tid == 0) {
            __stcg(p.ws.att_ml + (size_t)(h * S + s) * 2, mx);
            __stcg(p.ws.att_ml + (size_t)(h * S + s) * 2 + 1, tot);
          }
        }
        DS_ST(10);
      }
      cp_async_wait_all();
      grid_barrier(p.ws.bar, bar_target, nblk);
      DS_ST(11);
    }
    // ======================= stage C: o_proj =======================
    {
      int j, bx;
      which_member(&p.geo.G_C, 1, bid, j, bx);
      if (j >= 0) {
        const quipb200_linear_t L = Ly.o;
        const GemvCfg c = make_cfg(L, bx, p.geo.G_C);
        // combine the split-KV partials into the fp16 attention output.  Split weights exp(m_s - M) / sum go through
        // shared memory: [head][split] floats, computed once per CTA
        float* wsm = rb.fred + 64;   // aliases the block buffers (unused in this stage); n_heads * S <= 512 floats
        if (p.use_mma) {             // partial outputs and SU of o_proj -> shared memory, 16 bytes per request
          const int n8 = (P.n_heads * S * DS_HD) >> 3;
          for (int i = tid; i < n8; i += DS_THREADS) {           // chunk swizzle: see the combine below
            const int row = i >> 4, hh = row / S;
            cp_async16(stg.atto + (((row << 4) | ((i & 15) ^ ((hh >> 1) & 7))) << 3), p.ws.att_o + i * 8);
          }
          stg_vec(stg.su, reinterpret_cast<const __half*>(L.SU), L.in_features, tid);
        }
        if (tid < P.n_heads) {
          float m[DS_MAX_SPLITS], lsum[DS_MAX_SPLITS], M = -INFINITY, den = 0.f;
#pragma unroll
          for (int s = 0; s < DS_MAX_SPLITS; s++) {
            m[s] = (s < S) ? __ldcg(p.ws.att_ml + (size_t)(tid * S + s) * 2) : -INFINITY;
            lsum[s] = (s < S) ? __ldcg(p.ws.att_ml + (size_t)(tid * S + s) * 2 + 1) : 0.f;
            M = fmaxf(M, m[s]);
          }
#pragma unroll
          for (int s = 0; s < DS_MAX_SPLITS; s++) {
            m[s] = (m[s] == -INFINITY) ? 0.f : __expf(m[s] - M);
            den = fmaf(m[s], lsum[s], den);
          }
          const float inv = 1.0f / den;
#pragma unroll
          for (int s = 0; s < DS_MAX_SPLITS; s++)
            if (s < S) wsm[tid * S + s] = m[s] * lsum[s] * inv;      // partial outputs are normalised by their own sum
        }
        cp_async_wait_all();
        __syncthreads();
        float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float xs;
        if (p.use_mma) {
          // register pairs at idx_spread positions: pairs xh and xh + 2 are 4 consecutive halfs of one head.  The row
          // groups of a warp are 2 heads apart (the same banks), so 16-byte chunk c of head row r = h * S + s is staged
          // at chunk c ^ ((h >> 1) & 7) of that row.
#pragma unroll
          for (int xh = 0; xh < 2; xh++) {
            const int i0 = idx_spread(warp, lane, xh);
            if (i0 < P.n_heads * DS_HD) {
              const int h = i0 >> 7, d0 = i0 & (DS_HD - 1);
              const int off = ((((d0 >> 3) ^ ((h >> 1) & 7))) << 3) + (d0 & 7);
              float ax0 = 0.f, ay0 = 0.f, ax1 = 0.f, ay1 = 0.f;
#pragma unroll
              for (int s = 0; s < DS_MAX_SPLITS; s++) {
                if (s < S) {
                  const float w = wsm[h * S + s];
                  const uint2 a = *reinterpret_cast<const uint2*>(stg.atto + (h * S + s) * DS_HD + off);
                  const float2 a0 = __half22float2(as_h2(a.x)), a1 = __half22float2(as_h2(a.y));
                  ax0 = fmaf(w, a0.x, ax0);
                  ay0 = fmaf(w, a0.y, ay0);
                  ax1 = fmaf(w, a1.x, ax1);
                  ay1 = fmaf(w, a1.y, ay1);
                }
              }
              const float2 v0 = __half22float2(__floats2half2_rn(ax0, ay0));
              const float2 v1 = __half22float2(__floats2half2_rn(ax1, ay1));
              f[2 * xh] = v0.x;
              f[2 * xh + 1] = v0.y;
              f[2 * xh + 4] = v1.x;
              f[2 * xh + 5] = v1.y;
            }
          }
          DS_ST(12);
          xs = in_side_m(f, false, 0.f, L, stg, true, hfrag, XS, rb.fred, xq, tid);
        } else {
          if (tid < ((P.n_heads * DS_HD) >> 3)) {
            const int h = (tid * 8) / DS_HD, d = (tid * 8) % DS_HD;
#pragma unroll
            for (int s = 0; s < DS_MAX_SPLITS; s++) {
              if (s < S) {
                const float w = wsm[h * S + s];
                float a8[8];
                unpack_h8(__ldcg(reinterpret_cast<const uint4*>(p.ws.att_o + (size_t)(h * S + s) * DS_HD + d)), a8);
#pragma unroll
                for (int e = 0; e < 8; e++) f[e] = fmaf(w, a8[e], f[e]);
              }
            }
            round_h8(f);
          }
          DS_ST(12);
          xs = in_side_k1(f, nullptr, 0.f, L, rb, xq, tid);
        }
        if (bx == 0 && tid == 0) p.ws.xscale[SL_O] = xs;
        DS_ST(13);
        {
          const quipb200_linear_t* nx[2] = {&Ly.gate, &Ly.up};
          prefetch_stage(nx, p.geo.G_D, 2, bid, tid, Ly.gate.K_right > 1);
        }
        gemv_first<CB>(cw, c, warp, lane, pol);
        gemv_run<CB>(cw, c, xq, tab, red, warp, lane, pol);
        DS_ST(14);
        gemv_store<CB>(c, red, p.ws.acc[SL_O], xs, L.resid_scale, tid);
      } else {
        const quipb200_linear_t* nx[2] = {&Ly.gate, &Ly.up};
        prefetch_stage(nx, p.geo.G_D, 2, bid, tid, Ly.gate.K_right > 1);
      }
      grid_arrive(p.ws.bar, bar_target, nblk);
      if (p.use_mma && !(p.flags & 1)) stage_d_static(p, Ly, stg, bid, tid);
      grid_wait(p.ws.bar, bar_target);
      DS_ST(15);
    }
    // ======================= stage D: gate, up =======================
    {
      int j, bx;
      which_member(p.geo.G_D, 2, bid, j, bx);
      if (j >= 0) {
        const quipb200_linear_t L = (j == 0) ? Ly.gate : Ly.up;
        const bool colmix = L.K_right > 1;       // column-unit ownership + local K x K mix (gemv_store_mix)
        const GemvCfg c = colmix ? make_cfg_cols(L, bx, p.geo.G_D[j]) : make_cfg(L, bx, p.geo.G_D[j]);
        float f[8];
        float xs;
        const quipb200_linear_t Lp = Ly.o;
        if (p.use_mma) {
          if (p.flags & 1) {        // (else staged during the wait of stage C's barrier: stage_d_static)
            stg_vec(stg.sv, reinterpret_cast<const __half*>(Lp.SV), Lp.out_features, tid);
            stg_vec(stg.bias, reinterpret_cast<const __half*>(Lp.bias), Lp.out_features, tid);
            stg_vec(stg.resid, p.ws.hA, Lp.out_features, tid);
            stg_vec(stg.nw, reinterpret_cast<const __half*>(Ly.post_norm_w), L.in_features, tid);
            stg_vec(stg.su, reinterpret_cast<const __half*>(L.SU), L.in_features, tid);
          }
          const uint4 oct = out_side_load(p.ws.acc[SL_O], warp, lane);
          out_side_m(Lp, oct, true, stg, hfrag, XS, warp, lane, f, (dbg && l == p.dbg_layer) ? dbg : nullptr);
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const __half2 h = __floats2half2_rn(f[2 * q], f[2 * q + 1]);
            if (bid == 0) *reinterpret_cast<__half2*>(p.ws.hB + idx_spread(warp, lane, q)) = h;
            const float2 v = __half22float2(h);
            f[2 * q] = v.x;
            f[2 * q + 1] = v.y;
          }
          DS_ST(16);
          xs = in_side_m(f, true, P.norm_eps, L, stg, true, hfrag, XS, rb.fred, xq, tid);
        } else {
          out_side_k1(Lp, p.ws.acc[SL_O], p.ws.hA, rb, tid, f);
          const uint4 hv = pack_h8(f);
          if (bid == 0 && tid < hid8) reinterpret_cast<uint4*>(p.ws.hB)[tid] = hv;
          unpack_h8(hv, f);
          DS_ST(16);
          xs = in_side_k1(f, reinterpret_cast<const __half*>(Ly.post_norm_w), P.norm_eps, L, rb, xq, tid);
        }
        if (bx == 0 && tid == 0) p.ws.xscale[SL_G + j] = xs;
        DS_ST(17);
        {
          const quipb200_linear_t* nx[1] = {&Ly.down};
          prefetch_stage(nx, &p.geo.G_E, 1, bid, tid);
        }
        if (colmix) {     // this member's padded K x K factor (gate: slot 0, up: slot 1 of the layer's blob) -> shared memory
          const int Kp = (L.K_right + 15) / 16 * 16;
          const __half* src = reinterpret_cast<const __half*>(Ly.mlp_hk) + (size_t)j * Kp * Kp;
          for (int i = tid; i < (Kp * Kp) >> 3; i += DS_THREADS) cp_async16(bb.hkg + i * 8, src + i * 8);
        }
        gemv_first<CB>(cw, c, warp, lane, pol);
        gemv_run<CB>(cw, c, xq, tab, red, warp, lane, pol);
        DS_ST(18);
        if (colmix)
          gemv_store_mix<CB>(c, red, p.ws.acc[SL_G + j], xs, L.resid_scale, reinterpret_cast<const __half*>(L.wscale_pc), bb.hkg,
                             L.K_right, reinterpret_cast<float*>(bb.hku), tid);
        else
          gemv_store<CB>(c, red, p.ws.acc[SL_G + j], xs, L.resid_scale, tid);
      } else {
        const quipb200_linear_t* nx[1] = {&Ly.down};
        prefetch_stage(nx, &p.geo.G_E, 1, bid, tid);
      }
      grid_arrive(p.ws.bar, bar_target, nblk);
      if (!(p.flags & 1) && Ly.down.K_left > 1)
        stage_e_static(Ly.gate, Ly.up, Ly.down, reinterpret_cast<const __half*>(Ly.mlp_hk), bb, tid);
      grid_wait(p.ws.bar, bar_target);
      DS_ST(19);
    }
    // ======================= stage E: down =======================
    {
      int j, bx;
      which_member(&p.geo.G_E, 1, bid, j, bx);
      if (j >= 0) {
        const quipb200_linear_t L = Ly.down;
        const GemvCfg c = make_cfg(L, bx, p.geo.G_E);
        float xs;
        if (L.K_left > 1) {
          xs = stage_e_blocks(Ly.gate, Ly.up, L, p.ws.acc[SL_G], p.ws.acc[SL_U], reinterpret_cast<const __half*>(Ly.mlp_hk), bb,
                              hfrag, rb.fred, xq, tid, !(p.flags & 1), (dbg && l == p.dbg_layer) ? dbg : nullptr);
        } else {
          float g[8], u[8];
          {
            const quipb200_linear_t Lg = Ly.gate;
            out_side_k1(Lg, p.ws.acc[SL_G], nullptr, rb, tid, g);
            round_h8(g);
          }
          __syncthreads();
          {
            const quipb200_linear_t Lu = Ly.up;
            out_side_k1(Lu, p.ws.acc[SL_U], nullptr, rb, tid, u);
            round_h8(u);
          }
#pragma unroll
          for (int e = 0; e < 8; e++) g[e] = silu_f(g[e]);
          round_h8(g);
#pragma unroll
          for (int e = 0; e < 8; e++) u[e] *= g[e];
          round_h8(u);
          xs = in_side_k1(u, nullptr, 0.f, L, rb, xq, tid);
        }
        DS_ST(22);
        if (bx == 0 && tid == 0) p.ws.xscale[SL_D] = xs;
        if (l + 1 < P.n_layers) {
          const quipb200_decode_layer_t& Ln = s_desc[(l + 1) & 1];
          const quipb200_linear_t* nx[3] = {&Ln.q, &Ln.k, &Ln.v};
          prefetch_stage(nx, p.geo.G_A, 3, bid, tid);
        }
        gemv_first<CB>(cw, c, warp, lane, pol);
        gemv_run<CB>(cw, c, xq, tab, red, warp, lane, pol);
        DS_ST(23);
        gemv_store<CB>(c, red, p.ws.acc[SL_D], xs, L.resid_scale, tid);
      } else if (l + 1 < P.n_layers) {
        const quipb200_decode_layer_t& Ln = s_desc[(l + 1) & 1];
        const quipb200_linear_t* nx[3] = {&Ln.q, &Ln.k, &Ln.v};
        prefetch_stage(nx, p.geo.G_A, 3, bid, tid);
      }
      grid_arrive(p.ws.bar, bar_target, nblk);
      if (p.use_mma && !(p.flags & 1) && l + 1 < P.n_layers) stage_a_static(p, Ly, s_desc[(l + 1) & 1], stg, bid, tid);
      grid_wait(p.ws.bar, bar_target);
      DS_ST(24);
    }
  }
  // ---- output of the last layer: rot_out(down) + residual -> h_out ----
  if (bid == 0) {
    float f[8];
    const quipb200_linear_t Lp = s_desc[(P.n_layers - 1) & 1].down;
    if (p.use_mma) {
      stg_vec(stg.sv, reinterpret_cast<const __half*>(Lp.SV), Lp.out_features, tid);
      stg_vec(stg.bias, reinterpret_cast<const __half*>(Lp.bias), Lp.out_features, tid);
      stg_vec(stg.resid, p.ws.hB, Lp.out_features, tid);
      out_side_m(Lp, out_side_load(p.ws.acc[SL_D], warp, lane), true, stg, hfrag, XS, warp, lane, f);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int i = idx_spread(warp, lane, q);
        if (i < P.hidden) *reinterpret_cast<__half2*>(p.h_out + i) = __floats2half2_rn(f[2 * q], f[2 * q + 1]);
      }
    } else {
      out_side_k1(Lp, p.ws.acc[SL_D], p.ws.hB, rb, tid, f);
      if (tid < hid8) reinterpret_cast<uint4*>(p.h_out)[tid] = pack_h8(f);
    }
    if (tid == 0) *reinterpret_cast<volatile unsigned int*>(p.ws.bar + 32) = bar_target;
  }
  DS_STAMP(63);
#undef DS_STAMP
#undef DS_ST
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int ilog2_exact_h(int v) {
  int l = 0;
  while ((1 << l) < v) l++;
  return ((1 << l) == v) ? l : -1;
}

static bool linear_ok(const quipb200_linear_t& L) {
  if ((L.codebook != QUIPB200_CB_E8P12 && L.codebook != QUIPB200_CB_E8P12RVQ4B && L.codebook != QUIPB200_CB_D4) ||
      !L.qidxs || !L.grid)
    return false;
  if (L.K_left < 1 || L.K_right < 1 || L.q_in % L.K_left || L.q_out % L.K_right) return false;
  if (L.in_features > L.q_in || L.out_features > L.q_out) return false;
  if ((L.K_left > 1 && !L.had_left) || (L.K_right > 1 && !L.had_right)) return false;
  if (ilog2_exact_h(L.q_in / L.K_left) < 0 || ilog2_exact_h(L.q_out / L.K_right) < 0) return false;
  if (L.q_in % 64 != 0) return false;                         // 16-byte row pitch
  if ((L.in_features & 7) || (L.out_features & 7)) return false;
  const void* ptrs[] = {L.qidxs, L.grid, L.SU, L.SV, L.bias, L.had_left, L.had_right, L.wscale_pc};
  for (const void* q : ptrs)
    if (!aligned16(q)) return false;
  return true;
}

static int group_ctas(const quipb200_linear_t* const* mem, int n, int nblk, int* G) {
  long long w[3], tot = 0;
  for (int i = 0; i < n; i++) { w[i] = (long long)mem[i]->q_out * mem[i]->q_in; tot += w[i]; }
  for (int i = 0; i < n; i++) {
    int Gi = (int)((long long)nblk * w[i] / tot);
    if (Gi > mem[i]->q_out) Gi = mem[i]->q_out;
    if (Gi < 1) Gi = 1;
    G[i] = Gi;
  }
  return 0;
}

int g_ds_splits = 0;   // test / tuning hook: force the number of KV splits (0 = automatic)
int g_ds_flags = 0;    // option "ds_flags"

struct DsLayout {
  DsSmem sm;
  DsGeom geo;
  size_t ws_bytes;
  size_t off_bar, off_xscale, off_hA, off_hB, off_acc[SL_N], off_att_o, off_att_ml;
  int splits;
  int use_mma;
};

static bool same_shape(const quipb200_linear_t& a, const quipb200_linear_t& b) {
  return a.in_features == b.in_features && a.out_features == b.out_features && a.q_in == b.q_in && a.q_out == b.q_out &&
         a.K_left == b.K_left && a.K_right == b.K_right;
}

static int ds_layout(const quipb200_decode_plan_t* P, const quipb200_decode_layer_t* hl, int nblk, DsLayout* out) {
  if (!P || !hl || P->n_layers < 1 || P->head_dim != DS_HD || P->n_heads < 1 || P->n_kv_heads < 1 ||
      P->n_heads % P->n_kv_heads || P->max_len < 1 || (P->hidden & 7))
    return QUIPB200_EUNSUPPORTED;
  const quipb200_decode_layer_t& Y = hl[0];
  for (int l = 0; l < P->n_layers; l++) {
    const quipb200_decode_layer_t& Z = hl[l];
    const quipb200_linear_t* all[SL_N] = {&Z.q, &Z.k, &Z.v, &Z.o, &Z.gate, &Z.up, &Z.down};
    const quipb200_linear_t* ref[SL_N] = {&Y.q, &Y.k, &Y.v, &Y.o, &Y.gate, &Y.up, &Y.down};
    for (int i = 0; i < SL_N; i++)
      if (!linear_ok(*all[i]) || !same_shape(*all[i], *ref[i]) || all[i]->codebook != Y.q.codebook) return QUIPB200_EUNSUPPORTED;
    if (!Z.input_norm_w || !Z.post_norm_w || !Z.k_cache || !Z.v_cache) return QUIPB200_EINVAL;
    if (Z.gate.K_right > 1 && !Z.mlp_hk) return QUIPB200_EINVAL;      // the padded coefficient blob is required for blocked MLP dims
  }
  // shape chain of a Llama decoder layer
  if (Y.q.in_features != P->hidden || Y.k.in_features != P->hidden || Y.v.in_features != P->hidden) return QUIPB200_EUNSUPPORTED;
  if (Y.q.out_features != P->n_heads * DS_HD || Y.k.out_features != P->n_kv_heads * DS_HD ||
      Y.v.out_features != P->n_kv_heads * DS_HD)
    return QUIPB200_EUNSUPPORTED;
  if (Y.o.in_features != P->n_heads * DS_HD || Y.o.out_features != P->hidden) return QUIPB200_EUNSUPPORTED;
  if (Y.gate.in_features != P->hidden || Y.up.in_features != P->hidden) return QUIPB200_EUNSUPPORTED;
  if (Y.gate.out_features != Y.up.out_features || Y.gate.q_out != Y.up.q_out || Y.down.in_features != Y.gate.out_features ||
      Y.down.q_in != Y.gate.q_out || Y.down.out_features != P->hidden)
    return QUIPB200_EUNSUPPORTED;
  // rotations covered by the register-resident code: pure power-of-two (<= 4096 points) everywhere except the MLP
  // intermediate, which may be K blocks of 256 with a K x K mix (K <= 64)
  const quipb200_linear_t* k1[4] = {&Y.q, &Y.k, &Y.v, &Y.o};
  size_t nmax = 64;
  for (const quipb200_linear_t* L : k1) {
    if (L->K_left != 1 || L->K_right != 1) return QUIPB200_EUNSUPPORTED;
    nmax = std::max(nmax, (size_t)std::max(L->q_in, L->q_out));
  }
  if (Y.q.q_out < 128 || Y.k.q_out < 128 || Y.v.q_out < 128) return QUIPB200_EUNSUPPORTED;
  if (Y.gate.K_left != 1 || Y.up.K_left != 1 || Y.down.K_right != 1) return QUIPB200_EUNSUPPORTED;
  nmax = std::max(nmax, (size_t)std::max(Y.gate.q_in, Y.down.q_out));
  const int K = Y.gate.K_right;
  if (Y.up.K_right != K || Y.down.K_left != K) return QUIPB200_EUNSUPPORTED;
  size_t t_halfs = 0, hk_halfs = 0, mid_halfs = 0;
  if (K == 1) {
    nmax = std::max(nmax, (size_t)Y.gate.q_out);
  } else {
    if (Y.gate.q_out / K != 256 || K > 64) return QUIPB200_EUNSUPPORTED;
    t_halfs = (size_t)(K + 1) * (256 + 8);
    const size_t Kp = (K + 15) / 16 * 16;
    hk_halfs = Kp * Kp;
    mid_halfs = ((size_t)Y.gate.out_features + 7) / 8 * 8;
    nmax = std::max(nmax, (size_t)2048);   // A | B doubles as the per-warp octet -> fragment scratch (16 KB)
    nmax = std::max(nmax, (mid_halfs + 3) / 4);   // ... and holds the staged SU of down (mid_halfs fp16) in stage E
  }
  if (nmax > 8 * DS_THREADS) return QUIPB200_EUNSUPPORTED;

  const quipb200_linear_t* gA[3] = {&Y.q, &Y.k, &Y.v};
  const quipb200_linear_t* gC[1] = {&Y.o};
  const quipb200_linear_t* gD[2] = {&Y.gate, &Y.up};
  const quipb200_linear_t* gE[1] = {&Y.down};
  DsGeom geo;
  group_ctas(gA, 3, nblk, geo.G_A);
  group_ctas(gC, 1, nblk, &geo.G_C);
  group_ctas(gD, 2, nblk, geo.G_D);
  if (K > 1)
    for (int i = 0; i < 2; i++) geo.G_D[i] = std::min(geo.G_D[i], Y.gate.q_out / K);
  group_ctas(gE, 1, nblk, &geo.G_E);
  out->geo = geo;
  struct { const quipb200_linear_t* const* m; int n; const int* G; } groups[4] = {
      {gA, 3, geo.G_A}, {gC, 1, &geo.G_C}, {gD, 2, geo.G_D}, {gE, 1, &geo.G_E}};
  size_t red = 0, xq = 0, accb[SL_N];
  for (auto& g : groups) {
    for (int i = 0; i < g.n; i++) {
      const quipb200_linear_t& L = *g.m[i];
      const int segs = L.codebook == QUIPB200_CB_E8P12RVQ4B ? 4 : 8, accs = L.codebook == QUIPB200_CB_E8P12RVQ4B ? 2 : 1;
      const int nseg = L.q_in / 8, lanes = (nseg + segs - 1) / segs, C = (lanes + 31) / 32;
      size_t rows = (size_t)L.q_out / g.G[i] + 1;
      if (g.m == gD && L.K_right > 1) {              // column-unit ownership (make_cfg_cols): ceil(Lb / G) columns x K blocks
        const size_t Lb = (size_t)L.q_out / L.K_right;
        rows = (Lb + g.G[i] - 1) / g.G[i] * L.K_right;
        if (rows / L.K_right > 8 || rows > 4096) return QUIPB200_EUNSUPPORTED;      // vbuf rows / the grow() magic division
      }
      red = std::max(red, rows * C * accs * sizeof(int));
      xq = std::max(xq, (size_t)((nseg + 7) / 8 * 8) * 16);
    }
  }
  const quipb200_linear_t* all[SL_N] = {&Y.q, &Y.k, &Y.v, &Y.o, &Y.gate, &Y.up, &Y.down};
  for (int i = 0; i < SL_N; i++) accb[i] = (size_t)all[i]->q_out * sizeof(__half);

  // KV splits per head: one per ~128 cached positions, at most 4 (measured: 4 splits beat 1 at 384 positions)
  int S = nblk / P->n_heads;
  if (S > DS_MAX_SPLITS) S = DS_MAX_SPLITS;
  if (S < 1) return QUIPB200_EUNSUPPORTED;        // fewer CTAs than heads
  if (g_ds_splits > 0) S = std::min(S, g_ds_splits);
  else S = std::min(S, std::max(1, (P->max_len + 127) / 128));
  if (P->n_heads * S > 512 || P->n_heads * S * DS_HD > 16384) return QUIPB200_EUNSUPPORTED;
  out->splits = S;
  const size_t chunk = S > 1 ? ((size_t)P->max_len + S - 2) / (S - 1) : (size_t)P->max_len;
  const size_t attn = (3 * DS_WARPS * 128 + 3 * 128 + 32 + DS_PV_GROUPS * 128 + std::max(chunk, (size_t)(4 * DS_WARPS))) * sizeof(float);   // sc: scores (grouped path) or the 32 half-warp (max, sum) pairs
  // tensor-path rotations when every hidden-side rotation has exactly 4096 points
  const bool use_mma = Y.q.q_in == 4096 && Y.k.q_in == 4096 && Y.v.q_in == 4096 && Y.o.q_in == 4096 && Y.o.q_out == 4096 &&
                       Y.gate.q_in == 4096 && Y.up.q_in == 4096 && Y.down.q_out == 4096 && P->hidden <= 4096;
  out->use_mma = use_mma ? 1 : 0;
  if (use_mma) nmax = std::max(nmax, (size_t)4096);   // the exchange buffer (16 x 388 floats) lives in A | B
  size_t blk = 2 * t_halfs * 2 + 3 * hk_halfs * 2 + 2 * mid_halfs * 2;
  if (use_mma) blk = std::max(blk, (size_t)90112);    // 8 KB spare + staged vectors / attention partials (72 KB) + 8 KB spare
  size_t scr = 2 * nmax * sizeof(float) + 64 * sizeof(float) + blk;
  scr = std::max(scr, attn);
  auto up16 = [](size_t v) { return (v + 15) / 16 * 16; };
  DsSmem sm;
  size_t off = 0;
  sm.tab = 0; off += DS_TAB_BYTES;
  sm.red = (uint32_t)off; off += up16(red);
  sm.xq = (uint32_t)off; off += up16(xq);
  sm.scr = (uint32_t)off; off += up16(scr);
  sm.total = (uint32_t)off;
  sm.nmax = (uint32_t)nmax;
  sm.t_halfs = (uint32_t)t_halfs;
  sm.hk_halfs = (uint32_t)hk_halfs;
  sm.mid_halfs = (uint32_t)mid_halfs;
  if (off > 227 * 1024 - 2560) return QUIPB200_EUNSUPPORTED;   // static descriptors (1.5 KB) + the 1 KB the system reserves per CTA
  out->sm = sm;
  size_t w = 0;
  auto take = [&](size_t b) { size_t o = w; w += (b + 255) / 256 * 256; return o; };
  out->off_bar = take(256);
  out->off_xscale = take(SL_N * sizeof(float));
  out->off_hA = take((size_t)P->hidden * 2);
  out->off_hB = take((size_t)P->hidden * 2);
  for (int i = 0; i < SL_N; i++) out->off_acc[i] = take(accb[i]);
  out->off_att_o = take((size_t)P->n_heads * S * DS_HD * sizeof(__half));
  out->off_att_ml = take((size_t)P->n_heads * S * 2 * sizeof(float));
  out->ws_bytes = w;
  return 0;
}

long long* g_ds_dbg = nullptr;
int g_ds_dbg_cta = 0;
int g_ds_dbg_layer = 1;

}  // namespace qb

using namespace qb;

extern "C" int quipb200_decode_step_set_splits(int splits) {
  if (splits < 0 || splits > DS_MAX_SPLITS) return QUIPB200_EINVAL;
  g_ds_splits = splits;
  return 0;
}

extern "C" int quipb200_decode_step_debug(void* device_int64_buffer) {
  g_ds_dbg = (long long*)device_int64_buffer;
  return 0;
}

extern "C" int quipb200_decode_step_debug_cta(int cta) {
  g_ds_dbg_cta = cta < 0 ? 0 : (cta & 0xffff);
  g_ds_dbg_layer = cta < 0 ? 1 : ((cta >> 16) ? (cta >> 16) - 1 : 1);      // bits 16..: layer + 1 (0 = default layer 1)
  return 0;
}

extern "C" size_t quipb200_decode_step_workspace_bytes(const quipb200_decode_plan_t* plan,
                                                       const quipb200_decode_layer_t* host_layers) {
  const int sms = quipb200_sm_count();
  if (sms < 1) return 0;
  DsLayout lay;
  if (ds_layout(plan, host_layers, sms, &lay)) return 0;
  return lay.ws_bytes;
}

extern "C" int quipb200_decode_step(const quipb200_decode_plan_t* plan, const quipb200_decode_layer_t* host_layers,
                                    const void* h_in, void* h_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!plan || !host_layers || !h_in || !h_out || !workspace || !plan->layers || !plan->cos_t || !plan->sin_t || !plan->pos)
    return QUIPB200_EINVAL;
  if (!aligned16(h_in) || !aligned16(h_out) || ((uintptr_t)workspace & 255)) return QUIPB200_EALIGN;
  const int sms = quipb200_sm_count();
  if (sms < 1) return (int)cudaErrorNoDevice;
  DsLayout lay;
  int rc = ds_layout(plan, host_layers, sms, &lay);
  if (rc) return rc;
  if (workspace_bytes < lay.ws_bytes) return QUIPB200_EWORKSPACE;
  static int checked_dev[5] = {-1, -1, -1, -1, -1};
  int dev = 0;
  cudaGetDevice(&dev);
  const int cb = host_layers[0].q.codebook;
  const void* fn = cb == QUIPB200_CB_E8P12        ? (const void*)decode_step_kernel<QUIPB200_CB_E8P12>
                   : cb == QUIPB200_CB_E8P12RVQ4B ? (const void*)decode_step_kernel<QUIPB200_CB_E8P12RVQ4B>
                                                  : (const void*)decode_step_kernel<QUIPB200_CB_D4>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.sm.total);
  if (e != cudaSuccess) return (int)e;
  if (checked_dev[cb] != dev) {
    int coop = 0, occ = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, DS_THREADS, lay.sm.total);
    if (e != cudaSuccess) return (int)e;
    if (!coop || occ < 1) return QUIPB200_EUNSUPPORTED;
    checked_dev[cb] = dev;
  }
  DsParams p{};
  p.plan = *plan;
  unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
  p.ws.bar = reinterpret_cast<unsigned int*>(w + lay.off_bar);
  p.ws.xscale = reinterpret_cast<float*>(w + lay.off_xscale);
  p.ws.hA = reinterpret_cast<__half*>(w + lay.off_hA);
  p.ws.hB = reinterpret_cast<__half*>(w + lay.off_hB);
  for (int i = 0; i < SL_N; i++) p.ws.acc[i] = reinterpret_cast<__half*>(w + lay.off_acc[i]);
  p.ws.att_o = reinterpret_cast<__half*>(w + lay.off_att_o);
  p.ws.att_ml = reinterpret_cast<float*>(w + lay.off_att_ml);
  p.sm = lay.sm;
  p.geo = lay.geo;
  p.h_in = reinterpret_cast<const __half*>(h_in);
  p.h_out = reinterpret_cast<__half*>(h_out);
  p.kv_splits = lay.splits;
  p.use_mma = lay.use_mma;
  p.flags = g_ds_flags;
  p.dbg = g_ds_dbg;
  p.dbg_cta = g_ds_dbg_cta;
  p.dbg_layer = g_ds_dbg_layer;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(sms);
  cfg.blockDim = dim3(DS_THREADS);
  cfg.dynamicSmemBytes = lay.sm.total;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  void* args[] = {&p};
  e = cudaLaunchKernelExC(&cfg, fn, args);
  if (e != cudaSuccess) return (int)e;
  QB_LAUNCH_CHECK();
  return 0;
}
