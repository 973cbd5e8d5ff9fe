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
// already in flight while the rotation runs.  The arithmetic (fixed-point activations, exact int32
// dp4a, rounding points) is that of quantlinear.cu: the same device functions are used.
//
// Reference chain replaced: example_generate.py:29-32 (decode_one_tokens) -> HF LlamaDecoderLayer ->
// 7 x qlinear.py:87-115 per layer.
#include <algorithm>

#include "ql_device.cuh"

namespace qb {

constexpr int DS_THREADS = 512;
constexpr int DS_WARPS = DS_THREADS / 32;
constexpr int DS_UNROLL = 4;
constexpr int DS_MAX_SPLITS = 4;
constexpr int DS_HD = 128;
constexpr int DS_PV_GROUPS = DS_THREADS / 64;

enum { SL_Q = 0, SL_K, SL_V, SL_O, SL_G, SL_U, SL_D, SL_N };

struct DsWs {            // global scratch (device pointers)
  unsigned int* bar;     // [0] arrival counter, [32] epoch base (zero-filled once by the caller)
  float* xscale;         // [SL_N] fixed-point scale of each linear's input vector
  __half* hA;            // layer input (residual of the attention block)
  __half* hB;            // post-attention hidden (residual of the MLP block)
  float* acc[SL_N];      // raw integer dot products of each linear
  float* att_o;          // [nh][S][hd] un-normalised partial attention outputs
  float* att_ml;         // [nh][S][2]  running max / sum of each partial
};

struct DsSmem { uint32_t tab, red, xq, vh, vg, vu, rot, total; };   // byte offsets into dynamic smem

struct DsParams {
  quipb200_decode_plan_t plan;
  DsWs ws;
  DsSmem sm;
  const __half* h_in;
  __half* h_out;
  int kv_splits;
  long long* dbg;        // optional [16] clock stamps of CTA 0 (tools/timeline)
};

// ---------------------------------------------------------------------------------------------
// grid-wide barrier: monotonically increasing arrival counter, release/acquire at gpu scope
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& target, unsigned int nblk) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblk;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while ((int)(v - target) < 0);
    __threadfence();
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// GEMV over a contiguous row range (same inner loop as ql_gemv_kernel<E8P12>)
// ---------------------------------------------------------------------------------------------
struct GemvCfg {
  const unsigned char* q;
  int64_t row_bytes;
  int nseg, C, g, row_begin, nrows;
};

__device__ __forceinline__ int ilog2_dev(int v) { return 31 - __clz(v); }

__device__ __forceinline__ GemvCfg make_cfg(const quipb200_linear_t& L, int bx, int G) {
  GemvCfg c;
  c.q = reinterpret_cast<const unsigned char*>(L.qidxs);
  c.nseg = L.q_in >> 3;
  c.row_bytes = (int64_t)c.nseg * 2;
  const int lanes = (c.nseg + 7) >> 3;
  c.C = (lanes + 31) >> 5;
  c.g = c.C >= DS_WARPS ? 1 : DS_WARPS / c.C;
  const int base = L.q_out / G, rem = L.q_out % G;
  c.row_begin = bx * base + min(bx, rem);
  c.nrows = base + (bx < rem ? 1 : 0);
  return c;
}

__device__ __forceinline__ void gemv_first(uint4 (&cw)[DS_UNROLL], const GemvCfg& c, int warp, int lane, uint64_t pol) {
  const int units = c.C * c.g;
  const int chunk = warp / c.g, sub = warp - chunk * c.g;
  const bool lv = warp < units && (chunk * 32 + lane) * 8 < c.nseg;
  const unsigned char* colp = c.q + (size_t)(chunk * 32 + lane) * 16 + (size_t)c.row_begin * c.row_bytes;
#pragma unroll
  for (int u = 0; u < DS_UNROLL; u++) {
    const int r = sub + u * c.g;
    cw[u] = make_uint4(0, 0, 0, 0);
    if (lv && r < c.nrows) cw[u] = ldg_stream_v4(colp + (size_t)r * c.row_bytes, pol);
  }
}

// xq: swizzled 16-byte activation records in shared memory; red: [nrows][C] chunk partials
__device__ __forceinline__ void gemv_run(uint4 (&cw)[DS_UNROLL], const GemvCfg& c, const uint4* xq,
                                         const unsigned char* tab, int* red, int warp, int lane, uint64_t pol) {
  const int units = c.C * c.g;
  int unit = warp;
  while (unit < units) {
    const int chunk = unit / c.g;
    const int sub = unit - chunk * c.g;
    const int seg0 = (chunk * 32 + lane) * 8;
    const bool lane_valid = seg0 < c.nseg;
    uint32_t xs[8][4];
    int xsum[8];
#pragma unroll
    for (int sgi = 0; sgi < 8; sgi++) {
      uint4 r = make_uint4(0, 0, 0, 0);
      if (lane_valid) r = xq[swz(seg0 + sgi)];
      xs[sgi][0] = r.x; xs[sgi][1] = r.y; xs[sgi][2] = r.z; xs[sgi][3] = r.w;
      const int sh = dp4a_ss(r.x, 0x01010101u, dp4a_ss(r.y, 0x01010101u, 0));
      const int sl = dp4a_su(0x01010101u, r.z, dp4a_su(0x01010101u, r.w, 0));
      xsum[sgi] = sh * 256 + sl;
    }
    const unsigned char* colp = c.q + (size_t)(chunk * 32 + lane) * 16 + (size_t)c.row_begin * c.row_bytes;
    for (int r0 = sub; r0 < c.nrows; r0 += c.g * DS_UNROLL) {
      uint4 nx[DS_UNROLL];
      const int rn = r0 + c.g * DS_UNROLL;
#pragma unroll
      for (int u = 0; u < DS_UNROLL; u++) {
        const int r = rn + u * c.g;
        nx[u] = make_uint4(0, 0, 0, 0);
        if (lane_valid && r < c.nrows) nx[u] = ldg_stream_v4(colp + (size_t)r * c.row_bytes, pol);
      }
#pragma unroll 1
      for (int u = 0; u < DS_UNROLL; u++) {
        const int r = r0 + u * c.g;
        if (r < c.nrows) {
          int aH = 0, aL = 0, aP = 0, cH = 0, cL = 0, cP = 0;
          const uint32_t w[4] = {cw[0].x, cw[0].y, cw[0].z, cw[0].w};
#pragma unroll
          for (int i = 0; i < 4; i++) {
            e8p_dot((w[i] >> 5) & 0x7f8u, w[i] & 0xffu, tab, xs[2 * i], xsum[2 * i], aH, aL, aP);
            e8p_dot((w[i] >> 21) & 0x7f8u, __byte_perm(w[i], 0, 0x4442), tab, xs[2 * i + 1], xsum[2 * i + 1], cH, cL, cP);
          }
          aH += cH; aL += cL; aP += cP;
          int tot = aH * 256 + aL - 2 * aP;
          tot = __reduce_add_sync(0xffffffffu, tot);
          if (lane == 0) red[r * c.C + chunk] = tot;
        }
#pragma unroll
        for (int v = 0; v + 1 < DS_UNROLL; v++) cw[v] = cw[v + 1];
      }
#pragma unroll
      for (int u = 0; u < DS_UNROLL; u++) cw[u] = nx[u];
    }
    unit += DS_WARPS;
    if (unit < units) {
      const int chunk2 = unit / c.g, sub2 = unit - chunk2 * c.g;
      const bool lv = (chunk2 * 32 + lane) * 8 < c.nseg;
      const unsigned char* colp2 = c.q + (size_t)(chunk2 * 32 + lane) * 16 + (size_t)c.row_begin * c.row_bytes;
#pragma unroll
      for (int u = 0; u < DS_UNROLL; u++) {
        const int r = sub2 + u * c.g;
        cw[u] = make_uint4(0, 0, 0, 0);
        if (lv && r < c.nrows) cw[u] = ldg_stream_v4(colp2 + (size_t)r * c.row_bytes, pol);
      }
    }
  }
}

// CTA -> (member, index within member, CTAs of the member): CTAs are split in proportion to code bytes
__device__ __forceinline__ void split_ctas(const quipb200_linear_t* const* mem, int n, int nblk, int bid, int& j, int& bx,
                                           int& G) {
  long long w[3], tot = 0;
  for (int i = 0; i < n; i++) { w[i] = (long long)mem[i]->q_out * mem[i]->q_in; tot += w[i]; }
  int begin = 0;
  j = -1; bx = 0; G = 1;
  for (int i = 0; i < n; i++) {
    int Gi = (int)((long long)nblk * w[i] / tot);
    if (Gi > mem[i]->q_out) Gi = mem[i]->q_out;
    if (Gi < 1) Gi = 1;
    if (bid >= begin && bid < begin + Gi) { j = i; bx = bid - begin; G = Gi; }
    begin += Gi;
  }
}

__device__ __forceinline__ void fill_pro(PrologueArgs& pa, const quipb200_linear_t& L, const __half* x, const __half* gate,
                                         const __half* norm_w, float eps) {
  pa.x = x; pa.ldx = 0; pa.gate = gate; pa.ldgate = 0; pa.norm_w = norm_w; pa.norm_eps = eps;
  pa.SU = reinterpret_cast<const __half*>(L.SU); pa.hadK = reinterpret_cast<const __half*>(L.had_left);
  pa.K = L.K_left; pa.in_features = L.in_features; pa.q_in = L.q_in;
  pa.log2L = ilog2_dev(L.q_in / L.K_left); pa.transform = 1;
  pa.scale = L.wscale_float / sqrtf((float)(L.q_in / L.K_left));
  pa.xq = nullptr; pa.xscale = nullptr;
}

__device__ __forceinline__ void fill_epi(EpilogueArgs& ea, const quipb200_linear_t& L, const float* acc, const __half* residual,
                                         __half* y) {
  ea.acc = acc; ea.acc2 = nullptr; ea.xscale = nullptr; ea.unit = 0.25f; ea.resid_scale = 0.f;
  ea.wscale_pc = reinterpret_cast<const __half*>(L.wscale_pc); ea.hadK = reinterpret_cast<const __half*>(L.had_right);
  ea.K = L.K_right; ea.q_out = L.q_out; ea.out_features = L.out_features;
  ea.log2L = ilog2_dev(L.q_out / L.K_right); ea.transform = 1;
  ea.scale = 1.0f / sqrtf((float)(L.q_out / L.K_right));
  ea.SV = reinterpret_cast<const __half*>(L.SV); ea.bias = reinterpret_cast<const __half*>(L.bias);
  ea.residual = residual; ea.ldres = 0; ea.y = y; ea.ldy = 0;
}

// one copy of each side in the instruction stream
__device__ __noinline__ float ds_prologue(const PrologueArgs& pa, unsigned char* rot, uint4* xq, int tid) {
  const float xs = prologue_body<false>(pa, rot, xq, 0, tid, DS_THREADS, true);
  __syncthreads();
  return xs;
}
__device__ __noinline__ void ds_epilogue(const EpilogueArgs& ea, unsigned char* rot, float xscale, int tid) {
  epilogue_body<false>(ea, rot, 0, xscale, tid, DS_THREADS);
  __syncthreads();
}

// chunk partials -> global accumulator (fp32 image of the integer dot product)
__device__ __forceinline__ void gemv_store(const GemvCfg& c, const int* red, float* acc, int tid) {
  __syncthreads();
  for (int r = tid; r < c.nrows; r += DS_THREADS) {
    long long s = 0;
    for (int k = 0; k < c.C; k++) s += red[r * c.C + k];
    __stcg(acc + c.row_begin + r, (float)s);
  }
}

// ---------------------------------------------------------------------------------------------
// stage B helpers
// ---------------------------------------------------------------------------------------------
// 128 outputs [128*j, 128*j+128) of H_n f (n = 128*nb, Sylvester order): first the nb blocks are combined
// with the signs of row j of H_nb, then one 128-point transform.  part: [4][128] floats.
__device__ __forceinline__ void slice_partial(const quipb200_linear_t& L, const float* acc, float xs, int j, float* part,
                                              int tid) {
  const int lo = tid & 127, prt = tid >> 7;
  const int nb = L.q_out >> 7;
  const __half* wpc = reinterpret_cast<const __half*>(L.wscale_pc);
  float s = 0.f;
  for (int ih = prt; ih < nb; ih += 4) {
    float f = f16_round(__ldcg(acc + ih * 128 + lo) * xs);
    if (wpc) f = f16_round(f * __half2float(wpc[ih * 128 + lo]));
    s += (__popc(j & ih) & 1) ? -f : f;
  }
  part[prt * 128 + lo] = s;
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

// finish one head slice in a single warp: transform, output-side scalings of the linear (qlinear.py:108-114)
__device__ __forceinline__ void slice_finish(const quipb200_linear_t& L, const float* part, int j, int lane, float (&v)[4]) {
#pragma unroll
  for (int e = 0; e < 4; e++)
    v[e] = part[lane * 4 + e] + part[128 + lane * 4 + e] + part[256 + lane * 4 + e] + part[384 + lane * 4 + e];
  warp_fwht128(v, lane);
  const float sc = 1.0f / sqrtf((float)L.q_out);
  const __half* SV = reinterpret_cast<const __half*>(L.SV);
  const __half* bias = reinterpret_cast<const __half*>(L.bias);
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const int i = j * 128 + lane * 4 + e;
    v[e] = f16_round(v[e] * sc);
    if (SV) v[e] = f16_round(v[e] * __half2float(SV[i]));
    if (bias) v[e] = f16_round(v[e] + __half2float(bias[i]));
  }
}

__global__ void __launch_bounds__(DS_THREADS, 1) decode_step_kernel(const __grid_constant__ DsParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = gridDim.x, bid = blockIdx.x;
  const quipb200_decode_plan_t& P = p.plan;
  unsigned char* tab = smem + p.sm.tab;
  int* red = reinterpret_cast<int*>(smem + p.sm.red);
  uint4* xq = reinterpret_cast<uint4*>(smem + p.sm.xq);
  __half* vh = reinterpret_cast<__half*>(smem + p.sm.vh);
  __half* vg = reinterpret_cast<__half*>(smem + p.sm.vg);
  __half* vu = reinterpret_cast<__half*>(smem + p.sm.vu);
  unsigned char* rot = smem + p.sm.rot;
  long long* dbg = (p.dbg && bid == 0 && tid == 0) ? p.dbg : nullptr;
#define DS_STAMP(i) do { if (dbg) dbg[i] = clock64(); } while (0)
#define DS_ST(i) do { if (dbg && l == 1) dbg[i] = clock64(); } while (0)
  DS_STAMP(0);

  unsigned int bar_target = 0;
  if (tid == 0) bar_target = *reinterpret_cast<volatile unsigned int*>(p.ws.bar + 32);
  const uint64_t pol = l2_evict_first_policy();

  // the E8P abs table (identical for every linear), "+1/4" pre-applied
  if (tid < 256) {
    uint2 t = reinterpret_cast<const uint2*>(P.layers[0].q.grid)[tid];
    t.x |= 0x01010101u;
    t.y |= 0x01010101u;
    reinterpret_cast<uint2*>(tab)[tid] = t;
  }
  const int hid8 = P.hidden >> 3;
  if (bid == 0)
    for (int i = tid; i < hid8; i += DS_THREADS)
      reinterpret_cast<uint4*>(p.ws.hA)[i] = reinterpret_cast<const uint4*>(p.h_in)[i];
  __syncthreads();

  int pos = (int)(*P.pos);
  if (pos >= P.max_len) pos = P.max_len - 1;
  if (pos < 0) pos = 0;
  const int S = p.kv_splits;
  const float attn_scale = 1.0f / sqrtf((float)DS_HD);

  for (int l = 0; l < P.n_layers; l++) {
    const quipb200_decode_layer_t& Ly = P.layers[l];
    uint4 cw[DS_UNROLL];
    // ======================= stage A: q, k, v =======================
    {
      const quipb200_linear_t* mem[3] = {&Ly.q, &Ly.k, &Ly.v};
      int j, bx, G;
      split_ctas(mem, 3, nblk, bid, j, bx, G);
      if (j >= 0) {
        const quipb200_linear_t& L = *mem[j];
        const GemvCfg c = make_cfg(L, bx, G);
        DS_ST(1);
        gemv_first(cw, c, warp, lane, pol);
        if (l == 0) {
          for (int i = tid; i < hid8; i += DS_THREADS)
            reinterpret_cast<uint4*>(vh)[i] = reinterpret_cast<const uint4*>(p.h_in)[i];
          __syncthreads();
        } else {
          EpilogueArgs ea;
          fill_epi(ea, P.layers[l - 1].down, p.ws.acc[SL_D], p.ws.hB, vh);
          ds_epilogue(ea, rot, __ldcg(p.ws.xscale + SL_D), tid);
          if (bid == 0)
            for (int i = tid; i < hid8; i += DS_THREADS)
              reinterpret_cast<uint4*>(p.ws.hA)[i] = reinterpret_cast<const uint4*>(vh)[i];
        }
        DS_ST(2);
        PrologueArgs pa;
        fill_pro(pa, L, vh, nullptr, reinterpret_cast<const __half*>(Ly.input_norm_w), P.norm_eps);
        const float xs = ds_prologue(pa, rot, xq, tid);
        if (bx == 0 && tid == 0) p.ws.xscale[SL_Q + j] = xs;
        DS_ST(3);
        gemv_run(cw, c, xq, tab, red, warp, lane, pol);
        DS_ST(4);
        gemv_store(c, red, p.ws.acc[SL_Q + j], tid);
        DS_ST(5);
      }
      grid_barrier(p.ws.bar, bar_target, nblk);
      DS_ST(6);
    }
    // ======================= stage B: attention =======================
    {
      const int nh = P.n_heads, nkv = P.n_kv_heads, group = nh / nkv;
      if (bid < nh * S) {
        const int h = bid / S, s = bid - h * S, kvh = h / group;
        const int T = pos + 1, chunk = (T + S - 1) / S;
        const int t_begin = s * chunk, t_end = min(T, t_begin + chunk);
        const bool has_new = (t_begin <= pos) && (pos < t_end);
        float* part = reinterpret_cast<float*>(rot);          // [3][4][128]
        float* sq = part + 3 * 512;                           // [128] rotated, scaled query
        float* sk = sq + 128;                                 // [128] new key (post RoPE)
        float* sv = sk + 128;                                 // [128] new value
        float* sred = sv + 128;                               // [32]
        float* sout = sred + 32;                              // [DS_PV_GROUPS][128]
        float* sc = sout + DS_PV_GROUPS * 128;                // [chunk] scores
        __half* kc = reinterpret_cast<__half*>(Ly.k_cache) + (size_t)kvh * P.max_len * DS_HD;
        __half* vc = reinterpret_cast<__half*>(Ly.v_cache) + (size_t)kvh * P.max_len * DS_HD;
        slice_partial(Ly.q, p.ws.acc[SL_Q], __ldcg(p.ws.xscale + SL_Q) * 0.25f, h, part, tid);
        if (has_new) {
          slice_partial(Ly.k, p.ws.acc[SL_K], __ldcg(p.ws.xscale + SL_K) * 0.25f, kvh, part + 512, tid);
          slice_partial(Ly.v, p.ws.acc[SL_V], __ldcg(p.ws.xscale + SL_V) * 0.25f, kvh, part + 1024, tid);
        }
        __syncthreads();
        DS_ST(7);
        if (warp < (has_new ? 3 : 1)) {
          float v[4];
          const quipb200_linear_t& L = warp == 0 ? Ly.q : (warp == 1 ? Ly.k : Ly.v);
          slice_finish(L, part + warp * 512, warp == 0 ? h : kvh, lane, v);
          if (warp < 2) {   // RoPE, HF rotate_half convention: x*cos + rotate_half(x)*sin
            const __half* ct = reinterpret_cast<const __half*>(P.cos_t) + (size_t)pos * DS_HD;
            const __half* st = reinterpret_cast<const __half*>(P.sin_t) + (size_t)pos * DS_HD;
            const float sgn = (lane < 16) ? -1.f : 1.f;
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const int d = lane * 4 + e;
              const float pv = __shfl_xor_sync(0xffffffffu, v[e], 16);
              const float r = f16_round(v[e] * __half2float(ct[d]) + sgn * pv * __half2float(st[d]));
              if (warp == 0) sq[d] = r * attn_scale;
              else {
                sk[d] = r;
                if (h % group == 0) kc[(size_t)pos * DS_HD + d] = __float2half_rn(r);
              }
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const int d = lane * 4 + e;
              sv[d] = v[e];
              if (h % group == 0) vc[(size_t)pos * DS_HD + d] = __float2half_rn(v[e]);
            }
          }
        }
        __syncthreads();
        DS_ST(8);
        // ---- scores over [t_begin, t_end) ----
        const float q0 = sq[lane * 4], q1 = sq[lane * 4 + 1], q2 = sq[lane * 4 + 2], q3 = sq[lane * 4 + 3];
        float lmax = -INFINITY;
        constexpr int UN = 8;
        for (int t0 = t_begin + warp; t0 < t_end; t0 += DS_WARPS * UN) {
          uint2 raw[UN];
#pragma unroll
          for (int u = 0; u < UN; u++) {
            const int t = t0 + u * DS_WARPS;
            raw[u] = make_uint2(0, 0);
            if (t < t_end && t != pos) raw[u] = __ldcg(reinterpret_cast<const uint2*>(kc + (size_t)t * DS_HD + lane * 4));
          }
          float d[UN];
#pragma unroll
          for (int u = 0; u < UN; u++) {
            const int t = t0 + u * DS_WARPS;
            if (t == pos) {
              d[u] = q0 * sk[lane * 4] + q1 * sk[lane * 4 + 1] + q2 * sk[lane * 4 + 2] + q3 * sk[lane * 4 + 3];
            } else {
              const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw[u].x));
              const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw[u].y));
              d[u] = q0 * a.x + q1 * a.y + q2 * b.x + q3 * b.y;
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < UN; u++) d[u] += __shfl_xor_sync(0xffffffffu, d[u], o);
          }
#pragma unroll
          for (int u = 0; u < UN; u++) {
            const int t = t0 + u * DS_WARPS;
            if (t < t_end) {
              if (lane == 0) sc[t - t_begin] = d[u];
              lmax = fmaxf(lmax, d[u]);
            }
          }
        }
        if (lane == 0) sred[warp] = lmax;
        __syncthreads();
        float mx = -INFINITY;
#pragma unroll
        for (int w = 0; w < DS_WARPS; w++) mx = fmaxf(mx, sred[w]);
        __syncthreads();
        float lsum = 0.f;
        for (int t = t_begin + tid; t < t_end; t += DS_THREADS) {
          const float pr = __expf(sc[t - t_begin] - mx);
          sc[t - t_begin] = pr;
          lsum += pr;
        }
        lsum = warp_sum(lsum);
        if (lane == 0) sred[warp] = lsum;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < DS_WARPS; w++) tot += sred[w];
        DS_ST(9);
        // ---- partial out = P . V ----
        const int d2 = tid & 63, tg = tid >> 6;
        float o0 = 0.f, o1 = 0.f;
        constexpr int UV = 8;
        for (int t0 = t_begin + tg; t0 < t_end; t0 += DS_PV_GROUPS * UV) {
          __half2 vr[UV];
#pragma unroll
          for (int u = 0; u < UV; u++) {
            const int t = t0 + u * DS_PV_GROUPS;
            vr[u] = __float2half2_rn(0.f);
            if (t < t_end && t != pos) {
              const unsigned int w32 = __ldcg(reinterpret_cast<const unsigned int*>(vc + (size_t)t * DS_HD + d2 * 2));
              vr[u] = *reinterpret_cast<const __half2*>(&w32);
            } else if (t == pos) {
              vr[u] = __floats2half2_rn(sv[d2 * 2], sv[d2 * 2 + 1]);
            }
          }
#pragma unroll
          for (int u = 0; u < UV; u++) {
            const int t = t0 + u * DS_PV_GROUPS;
            if (t < t_end) {
              const float2 vv = __half22float2(vr[u]);
              const float pr = sc[t - t_begin];
              o0 = fmaf(pr, vv.x, o0);
              o1 = fmaf(pr, vv.y, o1);
            }
          }
        }
        sout[tg * 128 + d2 * 2] = o0;
        sout[tg * 128 + d2 * 2 + 1] = o1;
        __syncthreads();
        if (tid < DS_HD) {
          float r = 0.f;
#pragma unroll
          for (int g2 = 0; g2 < DS_PV_GROUPS; g2++) r += sout[g2 * 128 + tid];
          __stcg(p.ws.att_o + (size_t)(h * S + s) * DS_HD + tid, r);
        }
        if (tid == 0) {
          __stcg(p.ws.att_ml + (size_t)(h * S + s) * 2, mx);
          __stcg(p.ws.att_ml + (size_t)(h * S + s) * 2 + 1, tot);
        }
        DS_ST(10);
      }
      grid_barrier(p.ws.bar, bar_target, nblk);
      DS_ST(11);
    }
    // ======================= stage C: o_proj =======================
    {
      const quipb200_linear_t* mem[1] = {&Ly.o};
      int j, bx, G;
      split_ctas(mem, 1, nblk, bid, j, bx, G);
      if (j >= 0) {
        const quipb200_linear_t& L = Ly.o;
        const GemvCfg c = make_cfg(L, bx, G);
        gemv_first(cw, c, warp, lane, pol);
        // combine the split-KV partials into the fp16 attention output
        const int nh = P.n_heads;
        for (int o = tid; o < (nh * DS_HD) >> 3; o += DS_THREADS) {
          const int h = (o * 8) / DS_HD, d = (o * 8) % DS_HD;
          float m[DS_MAX_SPLITS], M = -INFINITY, den = 0.f;
          for (int s = 0; s < S; s++) {
            m[s] = __ldcg(p.ws.att_ml + (size_t)(h * S + s) * 2);
            M = fmaxf(M, m[s]);
          }
          float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int s = 0; s < S; s++) {
            const float w = (m[s] == -INFINITY) ? 0.f : __expf(m[s] - M);
            den = fmaf(w, __ldcg(p.ws.att_ml + (size_t)(h * S + s) * 2 + 1), den);
            const float4 a = __ldcg(reinterpret_cast<const float4*>(p.ws.att_o + (size_t)(h * S + s) * DS_HD + d));
            const float4 b = __ldcg(reinterpret_cast<const float4*>(p.ws.att_o + (size_t)(h * S + s) * DS_HD + d) + 1);
            f[0] = fmaf(w, a.x, f[0]); f[1] = fmaf(w, a.y, f[1]); f[2] = fmaf(w, a.z, f[2]); f[3] = fmaf(w, a.w, f[3]);
            f[4] = fmaf(w, b.x, f[4]); f[5] = fmaf(w, b.y, f[5]); f[6] = fmaf(w, b.z, f[6]); f[7] = fmaf(w, b.w, f[7]);
          }
          const float inv = 1.0f / den;
#pragma unroll
          for (int e = 0; e < 8; e++) f[e] *= inv;
          reinterpret_cast<uint4*>(vh)[o] = pack_h8(f);
        }
        __syncthreads();
        DS_ST(12);
        PrologueArgs pa;
        fill_pro(pa, L, vh, nullptr, nullptr, 0.f);
        const float xs = ds_prologue(pa, rot, xq, tid);
        if (bx == 0 && tid == 0) p.ws.xscale[SL_O] = xs;
        DS_ST(13);
        gemv_run(cw, c, xq, tab, red, warp, lane, pol);
        DS_ST(14);
        gemv_store(c, red, p.ws.acc[SL_O], tid);
      }
      grid_barrier(p.ws.bar, bar_target, nblk);
      DS_ST(15);
    }
    // ======================= stage D: gate, up =======================
    {
      const quipb200_linear_t* mem[2] = {&Ly.gate, &Ly.up};
      int j, bx, G;
      split_ctas(mem, 2, nblk, bid, j, bx, G);
      if (j >= 0) {
        const quipb200_linear_t& L = *mem[j];
        const GemvCfg c = make_cfg(L, bx, G);
        gemv_first(cw, c, warp, lane, pol);
        EpilogueArgs ea;
        fill_epi(ea, Ly.o, p.ws.acc[SL_O], p.ws.hA, vh);
        ds_epilogue(ea, rot, __ldcg(p.ws.xscale + SL_O), tid);
        if (bid == 0)
          for (int i = tid; i < hid8; i += DS_THREADS)
            reinterpret_cast<uint4*>(p.ws.hB)[i] = reinterpret_cast<const uint4*>(vh)[i];
        DS_ST(16);
        PrologueArgs pa;
        fill_pro(pa, L, vh, nullptr, reinterpret_cast<const __half*>(Ly.post_norm_w), P.norm_eps);
        const float xs = ds_prologue(pa, rot, xq, tid);
        if (bx == 0 && tid == 0) p.ws.xscale[SL_G + j] = xs;
        DS_ST(17);
        gemv_run(cw, c, xq, tab, red, warp, lane, pol);
        DS_ST(18);
        gemv_store(c, red, p.ws.acc[SL_G + j], tid);
      }
      grid_barrier(p.ws.bar, bar_target, nblk);
      DS_ST(19);
    }
    // ======================= stage E: down =======================
    {
      const quipb200_linear_t* mem[1] = {&Ly.down};
      int j, bx, G;
      split_ctas(mem, 1, nblk, bid, j, bx, G);
      if (j >= 0) {
        const quipb200_linear_t& L = Ly.down;
        const GemvCfg c = make_cfg(L, bx, G);
        gemv_first(cw, c, warp, lane, pol);
        EpilogueArgs ea;
        fill_epi(ea, Ly.gate, p.ws.acc[SL_G], nullptr, vg);
        ds_epilogue(ea, rot, __ldcg(p.ws.xscale + SL_G), tid);
        DS_ST(20);
        fill_epi(ea, Ly.up, p.ws.acc[SL_U], nullptr, vu);
        ds_epilogue(ea, rot, __ldcg(p.ws.xscale + SL_U), tid);
        DS_ST(21);
        PrologueArgs pa;
        fill_pro(pa, L, vu, vg, nullptr, 0.f);
        const float xs = ds_prologue(pa, rot, xq, tid);
        if (bx == 0 && tid == 0) p.ws.xscale[SL_D] = xs;
        DS_ST(22);
        gemv_run(cw, c, xq, tab, red, warp, lane, pol);
        DS_ST(23);
        gemv_store(c, red, p.ws.acc[SL_D], tid);
      }
      grid_barrier(p.ws.bar, bar_target, nblk);
      DS_ST(24);
    }
  }
  // ---- output of the last layer: rot_out(down) + residual -> h_out ----
  if (bid == 0) {
    EpilogueArgs ea;
    fill_epi(ea, P.layers[P.n_layers - 1].down, p.ws.acc[SL_D], p.ws.hB, p.h_out);
    ds_epilogue(ea, rot, __ldcg(p.ws.xscale + SL_D), tid);
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
  if (L.codebook != QUIPB200_CB_E8P12 || !L.qidxs || !L.grid) return false;
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

struct DsLayout {
  DsSmem sm;
  size_t ws_bytes;
  size_t off_bar, off_xscale, off_hA, off_hB, off_acc[SL_N], off_att_o, off_att_ml;
  int splits;
};

static int ds_layout(const quipb200_decode_plan_t* P, const quipb200_decode_layer_t* hl, int nblk, DsLayout* out) {
  if (!P || !hl || P->n_layers < 1 || P->head_dim != DS_HD || P->n_heads < 1 || P->n_kv_heads < 1 ||
      P->n_heads % P->n_kv_heads || P->max_len < 1 || (P->hidden & 7))
    return QUIPB200_EUNSUPPORTED;
  size_t red = 0, xq = 0, rotb = 0, vh = 0, vmid = 0, accb[SL_N] = {0, 0, 0, 0, 0, 0, 0};
  for (int l = 0; l < P->n_layers; l++) {
    const quipb200_decode_layer_t& Y = hl[l];
    const quipb200_linear_t* all[SL_N] = {&Y.q, &Y.k, &Y.v, &Y.o, &Y.gate, &Y.up, &Y.down};
    for (int i = 0; i < SL_N; i++)
      if (!linear_ok(*all[i])) return QUIPB200_EUNSUPPORTED;
    if (!Y.input_norm_w || !Y.post_norm_w || !Y.k_cache || !Y.v_cache) return QUIPB200_EINVAL;
    // shape chain of a Llama decoder layer
    if (Y.q.in_features != P->hidden || Y.k.in_features != P->hidden || Y.v.in_features != P->hidden) return QUIPB200_EUNSUPPORTED;
    if (Y.q.out_features != P->n_heads * DS_HD || Y.k.out_features != P->n_kv_heads * DS_HD ||
        Y.v.out_features != P->n_kv_heads * DS_HD)
      return QUIPB200_EUNSUPPORTED;
    if (Y.q.K_right != 1 || Y.k.K_right != 1 || Y.v.K_right != 1) return QUIPB200_EUNSUPPORTED;
    if (Y.q.q_out < 128 || Y.k.q_out < 128 || Y.v.q_out < 128) return QUIPB200_EUNSUPPORTED;
    if (Y.o.in_features != P->n_heads * DS_HD || Y.o.out_features != P->hidden) return QUIPB200_EUNSUPPORTED;
    if (Y.gate.in_features != P->hidden || Y.up.in_features != P->hidden) return QUIPB200_EUNSUPPORTED;
    if (Y.gate.out_features != Y.up.out_features || Y.down.in_features != Y.gate.out_features ||
        Y.down.out_features != P->hidden)
      return QUIPB200_EUNSUPPORTED;
    const quipb200_linear_t* gA[3] = {&Y.q, &Y.k, &Y.v};
    const quipb200_linear_t* gC[1] = {&Y.o};
    const quipb200_linear_t* gD[2] = {&Y.gate, &Y.up};
    const quipb200_linear_t* gE[1] = {&Y.down};
    struct { const quipb200_linear_t* const* m; int n; } groups[4] = {{gA, 3}, {gC, 1}, {gD, 2}, {gE, 1}};
    for (auto& g : groups) {
      int G[3];
      group_ctas(g.m, g.n, nblk, G);
      for (int i = 0; i < g.n; i++) {
        const quipb200_linear_t& L = *g.m[i];
        const int nseg = L.q_in / 8, lanes = (nseg + 7) / 8, C = (lanes + 31) / 32;
        const size_t rows = (size_t)L.q_out / G[i] + 1;
        red = std::max(red, rows * C * sizeof(int));
        xq = std::max(xq, (size_t)((nseg + 7) / 8 * 8) * 16);
        rotb = std::max(rotb, rot_smem_bytes(L.q_in, L.K_left));
        rotb = std::max(rotb, rot_smem_bytes(L.q_out, L.K_right));
      }
    }
    for (int i = 0; i < SL_N; i++) accb[i] = std::max(accb[i], (size_t)all[i]->q_out * sizeof(float));
    vh = std::max(vh, (size_t)std::max(P->hidden, P->n_heads * DS_HD) * 2);
    vmid = std::max(vmid, (size_t)Y.gate.out_features * 2);
  }
  int S = nblk / P->n_heads;
  if (S > DS_MAX_SPLITS) S = DS_MAX_SPLITS;
  if (S < 1) return QUIPB200_EUNSUPPORTED;        // fewer CTAs than heads
  out->splits = S;
  const size_t chunk = ((size_t)P->max_len + S - 1) / S;
  const size_t attn = (3 * 512 + 3 * 128 + 32 + DS_PV_GROUPS * 128 + chunk) * sizeof(float);
  rotb = std::max(rotb, attn);
  auto up16 = [](size_t v) { return (v + 15) / 16 * 16; };
  DsSmem sm;
  size_t off = 0;
  sm.tab = 0; off += 2048;
  sm.red = (uint32_t)off; off += up16(red);
  sm.xq = (uint32_t)off; off += up16(xq);
  sm.vh = (uint32_t)off; off += up16(vh);
  sm.vg = (uint32_t)off; off += up16(vmid);
  sm.vu = (uint32_t)off; off += up16(vmid);
  sm.rot = (uint32_t)off; off += up16(rotb);
  sm.total = (uint32_t)off;
  if (off > 227 * 1024) return QUIPB200_EUNSUPPORTED;
  out->sm = sm;
  size_t w = 0;
  auto take = [&](size_t b) { size_t o = w; w += (b + 255) / 256 * 256; return o; };
  out->off_bar = take(256);
  out->off_xscale = take(SL_N * sizeof(float));
  out->off_hA = take((size_t)P->hidden * 2);
  out->off_hB = take((size_t)P->hidden * 2);
  for (int i = 0; i < SL_N; i++) out->off_acc[i] = take(accb[i]);
  out->off_att_o = take((size_t)P->n_heads * S * DS_HD * sizeof(float));
  out->off_att_ml = take((size_t)P->n_heads * S * 2 * sizeof(float));
  out->ws_bytes = w;
  return 0;
}

long long* g_ds_dbg = nullptr;

}  // namespace qb

using namespace qb;

extern "C" int quipb200_decode_step_debug(void* device_int64_buffer) {
  g_ds_dbg = (long long*)device_int64_buffer;
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
  static int checked_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaError_t e = cudaFuncSetAttribute(decode_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.sm.total);
  if (e != cudaSuccess) return (int)e;
  if (checked_dev != dev) {
    int coop = 0, occ = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, decode_step_kernel, DS_THREADS, lay.sm.total);
    if (e != cudaSuccess) return (int)e;
    if (!coop || occ < 1) return QUIPB200_EUNSUPPORTED;
    checked_dev = dev;
  }
  DsParams p{};
  p.plan = *plan;
  unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
  p.ws.bar = reinterpret_cast<unsigned int*>(w + lay.off_bar);
  p.ws.xscale = reinterpret_cast<float*>(w + lay.off_xscale);
  p.ws.hA = reinterpret_cast<__half*>(w + lay.off_hA);
  p.ws.hB = reinterpret_cast<__half*>(w + lay.off_hB);
  for (int i = 0; i < SL_N; i++) p.ws.acc[i] = reinterpret_cast<float*>(w + lay.off_acc[i]);
  p.ws.att_o = reinterpret_cast<float*>(w + lay.off_att_o);
  p.ws.att_ml = reinterpret_cast<float*>(w + lay.off_att_ml);
  p.sm = lay.sm;
  p.h_in = reinterpret_cast<const __half*>(h_in);
  p.h_out = reinterpret_cast<__half*>(h_out);
  p.kv_splits = lay.splits;
  p.dbg = g_ds_dbg;
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
  e = cudaLaunchKernelExC(&cfg, (const void*)decode_step_kernel, args);
  if (e != cudaSuccess) return (int)e;
  QB_LAUNCH_CHECK();
  return 0;
}
