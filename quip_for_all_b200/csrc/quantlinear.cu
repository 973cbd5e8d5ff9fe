// Fused QuantLinear.forward for the decode regime (small M) -- the hot path.
//
//   ONE kernel per call (or per group of up to 3 linears sharing an input: q/k/v, gate/up):
//     phase 0  every warp issues the 128-bit loads of its first packed-code rows (HBM latency starts now)
//     phase 1  every CTA redundantly computes x' = H (SU.x) * wscale/sqrt(n) in shared memory and
//              quantises it to 16-bit fixed point (hidden behind the phase-0 loads)
//     phase 2  GEMV: int8-decode(Qidxs) . x_q as exact integer dp4a on CUDA cores, full rows per CTA
//     phase 3  the last CTA to finish (atomic ticket) runs the output side:
//              *scale -> [*Wscale_pc] -> (hadK (x) H)/sqrt(L) -> [:out] -> *SV -> +bias [-> +residual]
//   Non power-of-two input dims (down_proj: 11008 = 43 * 256) run the input rotation -- FWHT blocks
//   plus a 43x43 orthogonal mix on the legacy tensor path (mma.sync) -- in a 1-CTA prologue kernel
//   first (2 launches).
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
#include "ql_device.cuh"
#include "rot_cluster.cuh"

namespace qb {

// ---------------------------------------------------------------------------------------------
// options (bench / tests)
// ---------------------------------------------------------------------------------------------
int g_opt_table_repl = 16;   // 16: replicated lookup tables (64 KB) when they fit; 1: always the plain 2 KB table + computed signs
int g_opt_gemv_warps = 0;    // 0: auto (16)
int g_opt_gemv_ctas_per_sm = 1;
int g_opt_stage_mask = 7;    // bench only: bit0 prologue, bit1 gemv, bit2 epilogue
int g_opt_fuse = 3;          // bit0: fuse prologue into the GEMV kernel, bit1: last-CTA epilogue
int g_opt_rot_cluster = 1;  // block rotations with an orthogonal mix run on a thread-block cluster (rot_cluster.cuh); 0: one CTA
int g_opt_lean = 1;          // use the instruction-cache-lean kernel instantiation when eligible
int g_opt_phase0 = 1;        // experiment: 0 = no early code prefetch
int g_opt_epi_mma = 1;       // 4096-point output rotations of the per-linear kernels on the tensor path (0: shuffle form)
long long* g_dbg_timeline = nullptr;   // profiling hook: per-CTA clock64 stamps of the GEMV kernel phases

// tickets for the last-CTA-done epilogue: zero at module load, reset by the CTA that consumes them.
// One slot per launched linear, handed out round-robin by the host; launches that share a slot must
// not overlap in time (4096 slots; a captured 7B decode step uses 224).
__device__ unsigned int g_counters[COUNTER_SLOTS * QUIPB200_MM_MAX_M];
static unsigned int g_next_slot = 0;

__global__ void __launch_bounds__(PRO_THREADS) ql_prologue_kernel(PrologueArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x;
  const float xs = prologue_body<false>(a, smem_raw, a.xq + (size_t)m * (a.q_in >> 3), m, threadIdx.x, PRO_THREADS, false);
  if (threadIdx.x == 0) a.xscale[m] = xs;
}

__global__ void __launch_bounds__(PRO_THREADS) ql_epilogue_kernel(EpilogueArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_launch_dependents();
  pdl_wait();
  epilogue_body<false>(a, smem_raw, blockIdx.x, a.xscale[blockIdx.x], threadIdx.x, PRO_THREADS);
}

// LUT: the decode uses the replicated abs + sign-mask tables (64 KB, ql_device.cuh) instead of the 2 KB table + computed
// signs: 10 instead of 21 instructions per code; chosen by the host when the shared memory is there (every Llama shape).
template <int CB, bool LEAN, bool LEAN_EPI, bool LUT>
__global__ void __launch_bounds__(GEMV_MAX_WARPS * 32, 1) ql_gemv_kernel(const __grid_constant__ GroupArgs ga) {
  using T = CbTraits<CB>;
  constexpr int UNROLL = GEMV_UNROLL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [table][red: rows_per_cta * C * ACCS ints][x records][rotation workspace]
  unsigned char* tab = smem_raw;
  int* red = reinterpret_cast<int*>(smem_raw + (LUT ? LUT_TAB_BYTES : T::TAB_BYTES));
  __shared__ float s_xscale;
  __shared__ int s_last;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  const int m = blockIdx.y;

  pdl_launch_dependents();   // the next kernel may start prefetching its own weights while we run
  long long* dbg = ga.dbg ? ga.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 : nullptr;
#define QB_STAMP(i) do { if (dbg && threadIdx.x == 0) dbg[i] = clock64(); } while (0)
  QB_STAMP(0);

  // ---- which member of the group, which rows ----
  int j = 0;
  if (ga.n > 1 && (int)blockIdx.x >= ga.cta_begin[1]) j = 1;
  if (ga.n > 2 && (int)blockIdx.x >= ga.cta_begin[2]) j = 2;
  const GemvArgs& a = ga.a[j];
  const int G = ga.cta_begin[j + 1] - ga.cta_begin[j];
  const int bx = blockIdx.x - ga.cta_begin[j];
  const int row_begin = bx * a.rows_base + min(bx, a.rows_rem);
  const int nrows = a.rows_base + (bx < a.rows_rem ? 1 : 0);
  const int units = a.C * a.g;

  // ---- warm L2 with the small per-layer vectors this CTA will need after the dependency wait ----
  {
    const char* pf = nullptr;
    int lines = 0;
    if (tid < 128) { pf = reinterpret_cast<const char*>(a.pro.SU); lines = (a.pro.in_features * 2 + 127) >> 7; }
    else if (tid < 256) { pf = reinterpret_cast<const char*>(a.pro.norm_w); lines = (a.pro.in_features * 2 + 127) >> 7; }
    else if (tid < 448) { pf = reinterpret_cast<const char*>(a.epi.SV); lines = (a.epi.out_features * 2 + 127) >> 7; }
    const int li = tid < 128 ? tid : (tid < 256 ? tid - 128 : tid - 256);
    if (pf != nullptr && li < lines) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + (size_t)li * 128));
  }

  // ---- phase 0: first batch of code loads for this warp's first unit ----
  const uint64_t pol = l2_evict_first_policy();
  uint4 cw[UNROLL];
  int unit = warp;
  {
    const int chunk = unit / a.g, sub = unit - chunk * a.g;
    const bool lv = unit < units && (chunk * 32 + lane) * T::SEGS < a.nseg;
    const unsigned char* colp = a.qidxs + (size_t)(chunk * 32 + lane) * 16 + (size_t)row_begin * a.row_bytes;
    if (ga.phase0) {
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        const int r = sub + u * a.g;
        cw[u] = make_uint4(0, 0, 0, 0);
        if (lv && r < nrows) cw[u] = ldg_stream_v4(colp + (size_t)r * a.row_bytes, pol);
      }
    }
  }

  // ---- table entry for this thread: issue the load now, park it in shared memory after phase 1 ----
  uint2 tabv = make_uint2(0, 0);
  if (!LUT && tid < 256) tabv = reinterpret_cast<const uint2*>(a.table)[tid];   // E8P: int8x8 entry; D4: 4 x fp16

  QB_STAMP(1);
  // everything above read only weights; from here on we consume what earlier kernels produced
  pdl_wait();
  QB_STAMP(2);

  // ---- phase 1: activation records ----
  const uint4* xq;
  float xscale;
  if (LEAN || a.fuse_pro) {
    uint4* xq_s = reinterpret_cast<uint4*>(smem_raw + a.xq_off);
    xscale = prologue_body<LEAN>(a.pro, smem_raw + a.rot_off, xq_s, m, tid, nt, true, dbg);
    xq = xq_s;
    if (!a.fuse_epi && bx == 0 && tid == 0) a.pro.xscale[m] = xscale;   // for the epilogue kernel
  } else {
    xq = a.pro.xq + (size_t)m * a.nseg;
    xscale = a.pro.xscale[m];
  }
  if (LUT) {
    if (LEAN && CB == QUIPB200_CB_E8P12) lut_build_tables_e8p_fast(tab, a.table, tid, nt);
    else lut_build_tables<CB>(tab, a.table, tid, nt);
  } else if (tid < 256) {
    if (CB == QUIPB200_CB_D4) {
      // fp16 [4] -> int8 (units of 1/2), byte order (0,2,1,3) to match perm_0213'd activations
      const __half2 h01 = *reinterpret_cast<const __half2*>(&tabv.x), h23 = *reinterpret_cast<const __half2*>(&tabv.y);
      const int v0 = __float2int_rn(__low2float(h01) * 2.0f) & 0xff, v1 = __float2int_rn(__high2float(h01) * 2.0f) & 0xff;
      const int v2 = __float2int_rn(__low2float(h23) * 2.0f) & 0xff, v3 = __float2int_rn(__high2float(h23) * 2.0f) & 0xff;
      reinterpret_cast<uint32_t*>(tab)[tid] = (uint32_t)v0 | ((uint32_t)v2 << 8) | ((uint32_t)v1 << 16) | ((uint32_t)v3 << 24);
    } else {
      tabv.x |= 0x01010101u;
      tabv.y |= 0x01010101u;
      reinterpret_cast<uint2*>(tab)[tid] = tabv;
    }
  }
  __syncthreads();
  QB_STAMP(3);
  if (!ga.phase0) {
    const int chunk = unit / a.g, sub = unit - chunk * a.g;
    const bool lv = unit < units && (chunk * 32 + lane) * T::SEGS < a.nseg;
    const unsigned char* colp = a.qidxs + (size_t)(chunk * 32 + lane) * 16 + (size_t)row_begin * a.row_bytes;
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const int r = sub + u * a.g;
      cw[u] = make_uint4(0, 0, 0, 0);
      if (lv && r < nrows) cw[u] = ldg_stream_v4(colp + (size_t)r * a.row_bytes, pol);
    }
  }

  // ---- phase 2: GEMV ----
  bool first_unit = true;
  while (unit < units) {
    const int chunk = unit / a.g;
    const int sub = unit - chunk * a.g;
    const int seg0 = (chunk * 32 + lane) * T::SEGS;
    const bool lane_valid = seg0 < a.nseg;   // row pitch is a multiple of 16 B => all-or-nothing
    uint32_t xs[T::SEGS][4];
    int xsum[T::SEGS];
    const uint32_t offA = (CB == QUIPB200_CB_D4) ? (uint32_t)lane * 4u : (uint32_t)(lane & 15) * 8u, offS = 128u + offA;
#pragma unroll
    for (int sgi = 0; sgi < T::SEGS; sgi++) {
      uint4 r = make_uint4(0, 0, 0, 0);
      if (lane_valid) r = xq[(LEAN || a.fuse_pro) ? swz(seg0 + sgi) : seg0 + sgi];
      xs[sgi][0] = r.x;
      xs[sgi][1] = r.y;
      xs[sgi][2] = r.z;
      xs[sgi][3] = r.w;
      xsum[sgi] = 0;
      if (!LUT) {      // (the table path folds the parity shift into the sign mask: no sum(x) term)
        const int sh = dp4a_ss(r.x, 0x01010101u, dp4a_ss(r.y, 0x01010101u, 0));
        const int sl = dp4a_su(0x01010101u, r.z, dp4a_su(0x01010101u, r.w, 0));
        xsum[sgi] = sh * 256 + sl;
      }
    }
    const unsigned char* colp = a.qidxs + (size_t)(chunk * 32 + lane) * 16 + (size_t)row_begin * a.row_bytes;
    if (first_unit) { QB_STAMP(4); first_unit = false; }

    for (int r0 = sub; r0 < nrows; r0 += a.g * UNROLL) {
      // software pipeline: fetch the next batch before decoding the current one
      uint4 nx[UNROLL];
      const int rn = r0 + a.g * UNROLL;
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        const int r = rn + u * a.g;
        nx[u] = make_uint4(0, 0, 0, 0);
        if (lane_valid && r < nrows) nx[u] = ldg_stream_v4(colp + (size_t)r * a.row_bytes, pol);
      }
      // one copy of the 8-code decode in the instruction stream (the kernel's hot code must stay well
      // inside the 32 KB instruction cache: every launch runs it only a handful of times); the
      // register queue is rotated instead of indexed
#pragma unroll 1
      for (int u = 0; u < UNROLL; u++) {
        const int r = r0 + u * a.g;
        if (r < nrows) {   // warp-uniform
          int aH = 0, aL = 0, aP = 0, bH = 0, bL = 0, bP = 0;
          const uint32_t w[4] = {cw[0].x, cw[0].y, cw[0].z, cw[0].w};
          if (CB == QUIPB200_CB_E8P12) {
            // four independent accumulation chains (two per word half) keep the dp4a pipe fed
            int cH = 0, cL = 0, cP = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
              if (LUT) {
                lut_e8p_dot<1, 0>(w[i], tab, offA, offS, xs[2 * i], aH, aL);
                lut_e8p_dot<3, 2>(w[i], tab, offA, offS, xs[2 * i + 1], cH, cL);
              } else {
                e8p_dot((w[i] >> 5) & 0x7f8u, w[i] & 0xffu, tab, xs[2 * i], xsum[2 * i], aH, aL, aP);
                e8p_dot((w[i] >> 21) & 0x7f8u, __byte_perm(w[i], 0, 0x4442), tab, xs[2 * i + 1], xsum[2 * i + 1], cH, cL, cP);
              }
            }
            aH += cH; aL += cL; aP += cP;
          } else if (CB == QUIPB200_CB_E8P12RVQ4B) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
              // main code = hi16, residual code = lo16, same x segment
              if (LUT) {
                lut_e8p_dot<3, 2>(w[i], tab, offA, offS, xs[i], aH, aL);
                lut_e8p_dot<1, 0>(w[i], tab, offA, offS, xs[i], bH, bL);
              } else {
                e8p_dot((w[i] >> 21) & 0x7f8u, __byte_perm(w[i], 0, 0x4442), tab, xs[i], xsum[i], aH, aL, aP);
                e8p_dot((w[i] >> 5) & 0x7f8u, w[i] & 0xffu, tab, xs[i], xsum[i], bH, bL, bP);
              }
            }
          } else {  // D4: byte c -> 4 weights; two codes per 8-element segment
            const uint32_t* t4 = reinterpret_cast<const uint32_t*>(tab);
#pragma unroll
            for (int i = 0; i < 4; i++) {
#pragma unroll
              for (int b = 0; b < 4; b++) {
                const uint32_t v = LUT ? *reinterpret_cast<const uint32_t*>(tab + prmt(w[i], offA, 0x5504u | (b << 4)))
                                       : t4[(w[i] >> (8 * b)) & 0xffu];
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
#pragma unroll
        for (int v = 0; v + 1 < UNROLL; v++) cw[v] = cw[v + 1];
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
        if (lv && r < nrows) cw[u] = ldg_stream_v4(colp2 + (size_t)r * a.row_bytes, pol);
      }
    }
  }
  QB_STAMP(5);
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

  QB_STAMP(6);
  // ---- phase 3: the last CTA of this member runs the output side ----
  if (!LEAN_EPI && !a.fuse_epi) return;
  __syncthreads();   // every thread's acc stores are ordered before thread 0's fence (fences are cumulative)
  if (tid == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(a.counters + m, 1u);
    s_last = (prev == (unsigned int)(G - 1));
    if (LEAN || a.fuse_pro) s_xscale = xscale;
  }
  __syncthreads();
  QB_STAMP(7);
  if (!s_last) return;
  if (tid == 0) a.counters[m] = 0;   // ready for the next launch that draws this slot
  epilogue_body<LEAN_EPI>(a.epi, smem_raw + a.rot_off, m, (LEAN || a.fuse_pro) ? s_xscale : a.pro.xscale[m], tid, nt, dbg);
  __syncthreads();
  QB_STAMP(8);
#undef QB_STAMP
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
  int C, g, warps, accs, tab_bytes;
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
  int wmax = g_opt_gemv_warps > 0 ? g_opt_gemv_warps : GEMV_MAX_WARPS;
  if (wmax > GEMV_MAX_WARPS) wmax = GEMV_MAX_WARPS;
  if (p->C >= wmax) { p->g = 1; p->warps = wmax; }
  else { p->g = wmax / p->C; p->warps = p->g * p->C; }
  return 0;
}

// per-linear workspace carve: [xq: M*nseg*16][xscale: M*4 -> 256 aligned][acc: M*N*4][acc2: M*N*4]
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

struct Member {
  int codebook, N, K;
  const void* qidxs;
  const void* grid;
  PrologueArgs pa;
  EpilogueArgs ea;
};

// Enqueue a group of 1..3 linears of the same codebook in one GEMV launch (+ prologue / epilogue
// kernels where they cannot be fused).
// the cluster rotation kernels take only the vectorised layout: whole octets, 16-byte aligned rows
static bool rotc_pro_ok(const PrologueArgs& a, int M) {
  return g_opt_rot_cluster && a.transform && rotc_supported(a.q_in, a.K, a.log2L) && (a.in_features & 7) == 0 && aligned16(a.x) &&
         aligned16(a.gate) && aligned16(a.SU) && aligned16(a.norm_w) && aligned16(a.hadK) &&
         (M == 1 || ((a.ldx & 7) == 0 && (a.ldgate & 7) == 0));
}
static bool rotc_epi_ok(const EpilogueArgs& a, int M) {
  return g_opt_rot_cluster && a.transform && rotc_supported(a.q_out, a.K, a.log2L) && (a.out_features & 7) == 0 &&
         aligned16(a.wscale_pc) && aligned16(a.SV) && aligned16(a.bias) && aligned16(a.residual) && aligned16(a.y) &&
         (M == 1 || ((a.ldres & 7) == 0 && (a.ldy & 7) == 0));
}

static int run_group(Member* mem, int n, int M, void* workspace, size_t ws_bytes, cudaStream_t st) {
  if (n < 1 || n > QUIPB200_MAX_GROUP) return QUIPB200_EINVAL;
  const int sms = quipb200_sm_count();
  if (sms < 1) return (int)cudaErrorNoDevice;
  GemvPlan plan[QUIPB200_MAX_GROUP];
  GroupArgs ga{};
  ga.dbg = g_dbg_timeline;
  ga.phase0 = g_opt_phase0;
  ga.n = n;
  int rc;
  size_t ws_off = 0;
  long long total_rows_w = 0;
  for (int j = 0; j < n; j++) {
    if (mem[j].codebook != mem[0].codebook) return QUIPB200_EUNSUPPORTED;
    if ((rc = gemv_plan(mem[j].codebook, mem[j].N, mem[j].K, &plan[j]))) return rc;
    if (plan[j].warps != plan[0].warps) return QUIPB200_EUNSUPPORTED;
    total_rows_w += (long long)mem[j].N * mem[j].K;
  }
  // CTAs: one per SM in total, split between the members in proportion to their code bytes
  int Gtot = sms * (g_opt_gemv_ctas_per_sm > 0 ? g_opt_gemv_ctas_per_sm : 1);
  int cta = 0;
  size_t smem = 0;
  bool any_unfused_pro = false, any_unfused_epi = false, lut_fits = true;
  uint32_t xq_rel[3] = {0, 0, 0}, rot_rel[3] = {0, 0, 0};
  for (int j = 0; j < n; j++) {
    Member& mj = mem[j];
    const bool two = plan[j].accs == 2;
    Workspace ws = carve_ws(workspace ? (unsigned char*)workspace + ws_off : nullptr, M, mj.N, mj.K, two);
    ws_off += ws.bytes;
    if (!workspace || ws_bytes < ws_off) return QUIPB200_EWORKSPACE;
    mj.pa.xq = ws.xq; mj.pa.xscale = ws.xscale;
    mj.ea.acc = ws.acc; mj.ea.acc2 = ws.acc2; mj.ea.xscale = ws.xscale;

    int G = (int)((long long)Gtot * mj.N * mj.K / total_rows_w);
    const int min_rows = plan[j].g * 2;   // keep ~2 rows per warp slot so tiny layers do not launch idle CTAs
    if ((long long)G * min_rows > mj.N) G = (mj.N + min_rows - 1) / min_rows;
    if (G < 1) G = 1;
    ga.cta_begin[j] = cta;
    cta += G;
    ga.cta_begin[j + 1] = cta;
    const int rows_per_cta = (mj.N + G - 1) / G + 1;

    const size_t pro_smem = rot_smem_bytes(mj.pa.q_in, mj.pa.K);
    const size_t epi_smem = rot_smem_bytes(mj.ea.q_out, mj.ea.K);
    if (pro_smem > SMEM_LIMIT || epi_smem > SMEM_LIMIT) return QUIPB200_EUNSUPPORTED;
    const size_t red_bytes = ((size_t)rows_per_cta * plan[j].C * plan[j].accs * sizeof(int) + 15) / 16 * 16;
    const size_t xq_bytes = (size_t)((mj.K / 8 + 7) / 8 * 8) * 16;   // swizzle stays inside 8-record blocks
    // fused prologue only for pure-FWHT (or identity) input sides: the orthogonal-block mix is too
    // much work to repeat in every CTA
    bool fuse_pro = (g_opt_fuse & 1) && mj.pa.K == 1;
    // orthogonal-mix output sides go to the cluster kernel (one CTA would spend longer on a 28672-point rotation than
    // the whole GEMV takes); the last-CTA epilogue keeps the pure power-of-two case
    // ... and only behind a fused input side: when the input rotation is its own kernel anyway (K_left > 1), a separate
    // epilogue kernel -- launched early through PDL, its vector loads issued before the dependency wait -- measured
    // 3-4 us faster than the ticketed last CTA (11008x4096: 19.2 vs 22.0 us, 28672x8192: 41.2 vs 45.3 us)
    bool fuse_epi = (g_opt_fuse & 2) != 0 && !rotc_epi_ok(mj.ea, M) && (fuse_pro || g_opt_fuse == 2);
    size_t need;
    for (;;) {
      size_t rot = 0;
      if (fuse_pro) rot = pro_smem;
      if (fuse_epi && epi_smem > rot) rot = epi_smem;
      need = plan[j].tab_bytes + red_bytes + (fuse_pro ? xq_bytes : 0) + rot;
      if (need <= SMEM_LIMIT) break;
      if (fuse_epi) fuse_epi = false;
      else if (fuse_pro) fuse_pro = false;
      else return QUIPB200_EUNSUPPORTED;
    }
    if (need > smem) smem = need;
    if (need - plan[j].tab_bytes + LUT_TAB_BYTES > SMEM_LIMIT) lut_fits = false;   // same fusion choices must still fit
    xq_rel[j] = (uint32_t)red_bytes;
    rot_rel[j] = (uint32_t)(red_bytes + (fuse_pro ? xq_bytes : 0));
    any_unfused_pro |= !fuse_pro;
    any_unfused_epi |= !fuse_epi;

    GemvArgs& g = ga.a[j];
    g.qidxs = (const unsigned char*)mj.qidxs; g.row_bytes = plan[j].row_bytes; g.table = mj.grid;
    g.N = mj.N; g.nseg = mj.K / 8; g.C = plan[j].C; g.g = plan[j].g;
    g.rows_base = mj.N / G; g.rows_rem = mj.N % G;
    g.fuse_pro = fuse_pro ? 1 : 0; g.fuse_epi = fuse_epi ? 1 : 0;
    g.pro = mj.pa; g.epi = mj.ea;
    g.counters = nullptr;
    if (fuse_epi) {
      unsigned int* counters = nullptr;
      cudaError_t e = cudaGetSymbolAddress((void**)&counters, g_counters);
      if (e != cudaSuccess) return (int)e;
      g.counters = counters + (size_t)(__atomic_fetch_add(&g_next_slot, 1u, __ATOMIC_RELAXED) % COUNTER_SLOTS) *
                                  QUIPB200_MM_MAX_M;
    }
  }

  // replicated lookup tables (64 KB) when every member keeps its fusion choices with them and the layers are big enough
  // to repay the ~0.5 us table build (option "gemv_table_repl": 1 = never)
  long long weights = 0;
  for (int j = 0; j < n; j++) weights += (long long)mem[j].N * mem[j].K;
  const bool use_lut = lut_fits && g_opt_table_repl != 1 && weights >= (8ll << 20);
  for (int j = 0; j < n; j++) {
    const uint32_t tb = use_lut ? (uint32_t)LUT_TAB_BYTES : (uint32_t)plan[j].tab_bytes;
    ga.a[j].xq_off = tb + xq_rel[j];
    ga.a[j].rot_off = tb + rot_rel[j];
  }
  if (use_lut) smem = smem - plan[0].tab_bytes + LUT_TAB_BYTES;

  // lean instantiations (instruction-cache footprint): fused, pure power-of-two rotation, every vector 16-byte
  // aligned, one octet per thread.  Input and output side qualify independently (gate/up: lean in, 43-block out).
  bool lean_pro = g_opt_lean && mem[0].codebook == QUIPB200_CB_E8P12;
  bool lean_epi = lean_pro;
  for (int j = 0; j < n; j++) {
    const PrologueArgs& pa = ga.a[j].pro;
    const EpilogueArgs& ea = ga.a[j].epi;
    const int nt = plan[j].warps * 32;
    lean_pro = lean_pro && ga.a[j].fuse_pro && pa.K == 1 && (pa.in_features & 7) == 0 && (pa.q_in >> 3) <= nt &&
               (!pa.transform || pa.log2L >= 3) && rot_pingpong(pa.q_in, 1) && aligned16(pa.x) && aligned16(pa.gate) &&
               aligned16(pa.SU) && aligned16(pa.norm_w) && (M == 1 || ((pa.ldx & 7) == 0 && (pa.ldgate & 7) == 0));
    lean_epi = lean_epi && ga.a[j].fuse_epi && ea.K == 1 && (ea.q_out >> 3) <= nt && (ea.out_features & 7) == 0 &&
               (!ea.transform || ea.log2L >= 3) && rot_pingpong(ea.q_out, 1) && aligned16(ea.SV) && aligned16(ea.bias) &&
               aligned16(ea.residual) && aligned16(ea.y) && aligned16(ea.acc) &&
               (M == 1 || ((ea.ldres & 7) == 0 && (ea.ldy & 7) == 0));
  }

  if (any_unfused_pro && (g_opt_stage_mask & 1)) {
    // cluster kernel: members of the same block shape share one launch
    bool done[QUIPB200_MAX_GROUP] = {false, false, false};
    for (int j = 0; j < n; j++) {
      if (ga.a[j].fuse_pro || done[j] || !rotc_pro_ok(ga.a[j].pro, M)) continue;
      RotcProGroup grp{};
      int cnt = 0;
      for (int i = j; i < n; i++) {
        if (ga.a[i].fuse_pro || done[i] || !rotc_pro_ok(ga.a[i].pro, M)) continue;
        if (ga.a[i].pro.K != ga.a[j].pro.K || ga.a[i].pro.log2L != ga.a[j].pro.log2L) continue;
        grp.a[cnt++] = ga.a[i].pro;
        done[i] = true;
      }
      RotcPlan rp = rotc_plan(ga.a[j].pro.K, ga.a[j].pro.log2L);
      const size_t csm = rotc_smem_bytes(ga.a[j].pro.K, ga.a[j].pro.log2L, rp);
      if ((rc = set_smem_attr((const void*)ql_prologue_cluster_kernel, csm))) return rc;
      void* cargs[] = {&grp, &rp};
      cudaError_t ce = launch_cluster_kernel((const void*)ql_prologue_cluster_kernel, dim3(rp.C, M, cnt), rp.C, cargs, csm, st);
      if (ce != cudaSuccess) return (int)ce;
      QB_LAUNCH_CHECK();
    }
    for (int j = 0; j < n; j++) {
      if (ga.a[j].fuse_pro || done[j]) continue;
      const size_t pro_smem = rot_smem_bytes(mem[j].pa.q_in, mem[j].pa.K);
      if ((rc = set_smem_attr((const void*)ql_prologue_kernel, pro_smem))) return rc;
      void* pargs[] = {&ga.a[j].pro};
      cudaError_t pe = launch_kernel((const void*)ql_prologue_kernel, dim3(M), dim3(PRO_THREADS), pargs, pro_smem, st);
      if (pe != cudaSuccess) return (int)pe;
      QB_LAUNCH_CHECK();
    }
  }
  if (g_opt_stage_mask & 2) {
    const void* fn = nullptr;
    switch (mem[0].codebook) {
      case QUIPB200_CB_E8P12:
        if (use_lut) {
          if (lean_pro && lean_epi) fn = (const void*)ql_gemv_kernel<QUIPB200_CB_E8P12, true, true, true>;
          else if (lean_pro) fn = (const void*)ql_gemv_kernel<QUIPB200_CB_E8P12, true, false, true>;
          else if (lean_epi) fn = (const void*)ql_gemv_kernel<QUIPB200_CB_E8P12, false, true, true>;
          else fn = (const void*)ql_gemv_kernel<QUIPB200_CB_E8P12, false, false, true>;
        } else {
          if (lean_pro && lean_epi) fn = (const void*)ql_gemv_kernel<QUIPB200_CB_E8P12, true, true, false>;
          else if (lean_pro) fn = (const void*)ql_gemv_kernel<QUIPB200_CB_E8P12, true, false, false>;
          else if (lean_epi) fn = (const void*)ql_gemv_kernel<QUIPB200_CB_E8P12, false, true, false>;
          else fn = (const void*)ql_gemv_kernel<QUIPB200_CB_E8P12, false, false, false>;
        }
        break;
      case QUIPB200_CB_E8P12RVQ4B:
        fn = use_lut ? (const void*)ql_gemv_kernel<QUIPB200_CB_E8P12RVQ4B, false, false, true>
                     : (const void*)ql_gemv_kernel<QUIPB200_CB_E8P12RVQ4B, false, false, false>;
        break;
      case QUIPB200_CB_D4:
        fn = use_lut ? (const void*)ql_gemv_kernel<QUIPB200_CB_D4, false, false, true>
                     : (const void*)ql_gemv_kernel<QUIPB200_CB_D4, false, false, false>;
        break;
      default: return QUIPB200_EUNSUPPORTED;
    }
    if ((rc = set_smem_attr(fn, smem))) return rc;
    void* args[] = {&ga};
    cudaError_t e = launch_kernel(fn, dim3(cta, M), dim3(plan[0].warps * 32), args, smem, st);
    if (e != cudaSuccess) return (int)e;
    QB_LAUNCH_CHECK();
  }
  if (any_unfused_epi && (g_opt_stage_mask & 4)) {
    bool done[QUIPB200_MAX_GROUP] = {false, false, false};
    for (int j = 0; j < n; j++) {
      if (ga.a[j].fuse_epi || done[j] || !rotc_epi_ok(ga.a[j].epi, M)) continue;
      RotcEpiGroup grp{};
      int cnt = 0;
      for (int i = j; i < n; i++) {
        if (ga.a[i].fuse_epi || done[i] || !rotc_epi_ok(ga.a[i].epi, M)) continue;
        if (ga.a[i].epi.K != ga.a[j].epi.K || ga.a[i].epi.log2L != ga.a[j].epi.log2L) continue;
        grp.a[cnt++] = ga.a[i].epi;
        done[i] = true;
      }
      RotcPlan rp = rotc_plan(ga.a[j].epi.K, ga.a[j].epi.log2L);
      const size_t csm = rotc_smem_bytes(ga.a[j].epi.K, ga.a[j].epi.log2L, rp);
      if ((rc = set_smem_attr((const void*)ql_epilogue_cluster_kernel, csm))) return rc;
      void* cargs[] = {&grp, &rp};
      cudaError_t ce = launch_cluster_kernel((const void*)ql_epilogue_cluster_kernel, dim3(rp.C, M, cnt), rp.C, cargs, csm, st);
      if (ce != cudaSuccess) return (int)ce;
      QB_LAUNCH_CHECK();
    }
    for (int j = 0; j < n; j++) {
      if (ga.a[j].fuse_epi || done[j]) continue;
      const size_t epi_smem = rot_smem_bytes(mem[j].ea.q_out, mem[j].ea.K);
      if ((rc = set_smem_attr((const void*)ql_epilogue_kernel, epi_smem))) return rc;
      void* eargs[] = {&ga.a[j].epi};
      cudaError_t ee = launch_kernel((const void*)ql_epilogue_kernel, dim3(M), dim3(PRO_THREADS), eargs, epi_smem, st);
      if (ee != cudaSuccess) return (int)ee;
      QB_LAUNCH_CHECK();
    }
  }
  return 0;
}

static int member_from_layer(const quipb200_linear_t* L, const quipb200_fusion_t* fu, const void* x, int64_t ldx,
                             void* y, int64_t ldy, Member* m) {
  if (!L->qidxs || !L->grid || !x || !y) return QUIPB200_EINVAL;
  if (L->K_left < 1 || L->K_right < 1 || L->q_in % L->K_left || L->q_out % L->K_right) return QUIPB200_EINVAL;
  if (L->in_features > L->q_in || L->out_features > L->q_out) return QUIPB200_EINVAL;
  if ((L->K_left > 1 && !L->had_left) || (L->K_right > 1 && !L->had_right)) return QUIPB200_EINVAL;
  const int log2Lin = ilog2_exact(L->q_in / L->K_left), log2Lout = ilog2_exact(L->q_out / L->K_right);
  if (log2Lin < 0 || log2Lout < 0) return QUIPB200_EINVAL;
  if (!aligned16(L->qidxs) || !aligned16(L->grid)) return QUIPB200_EALIGN;
  m->codebook = L->codebook; m->N = L->q_out; m->K = L->q_in; m->qidxs = L->qidxs; m->grid = L->grid;
  PrologueArgs pa{};
  pa.x = (const __half*)x; pa.ldx = ldx; pa.SU = (const __half*)L->SU; pa.hadK = (const __half*)L->had_left;
  pa.K = L->K_left; pa.in_features = L->in_features; pa.q_in = L->q_in; pa.log2L = log2Lin; pa.transform = 1;
  pa.scale = L->wscale_float / sqrtf((float)(L->q_in / L->K_left));   // quant.py:75
  if (fu) {
    pa.gate = (const __half*)fu->gate; pa.ldgate = fu->ldgate;
    pa.norm_w = (const __half*)fu->pre_norm_weight; pa.norm_eps = fu->pre_norm_eps;
  }
  EpilogueArgs ea{};
  ea.unit = (L->codebook == QUIPB200_CB_D4) ? 0.5f : 0.25f;
  ea.resid_scale = f16_round_host(L->resid_scale);
  ea.wscale_pc = (const __half*)L->wscale_pc; ea.hadK = (const __half*)L->had_right; ea.K = L->K_right;
  ea.q_out = L->q_out; ea.out_features = L->out_features; ea.log2L = log2Lout; ea.transform = 1;
  ea.scale = 1.0f / sqrtf((float)(L->q_out / L->K_right));
  ea.SV = (const __half*)L->SV; ea.bias = (const __half*)L->bias; ea.y = (__half*)y; ea.ldy = ldy;
  if (fu) { ea.residual = (const __half*)fu->residual; ea.ldres = fu->ldres; }
  ea.mma = (g_opt_epi_mma && L->K_right == 1) ? (L->q_out == 4096 ? 1 : (L->q_out == 8192 ? 2 : 0)) : 0;
  m->pa = pa; m->ea = ea;
  return 0;
}

}  // namespace qb

using namespace qb;

extern "C" int quipb200_debug_timeline(void* device_int64_buffer) {
  g_dbg_timeline = (long long*)device_int64_buffer;
  return 0;
}

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
  Member mem{};
  mem.codebook = codebook; mem.N = N; mem.K = K; mem.qidxs = qidxs; mem.grid = grid;
  PrologueArgs pa{};
  pa.x = (const __half*)x; pa.ldx = K; pa.K = 1;
  pa.in_features = K; pa.q_in = K; pa.log2L = 0; pa.transform = 0; pa.scale = 1.f;
  EpilogueArgs ea{};
  ea.unit = (codebook == QUIPB200_CB_D4) ? 0.5f : 0.25f;
  ea.resid_scale = f16_round_host(resid_scale);
  ea.K = 1; ea.q_out = N; ea.out_features = N; ea.log2L = 0;
  ea.transform = 0; ea.scale = 1.f; ea.y = (__half*)out; ea.ldy = N;
  mem.pa = pa; mem.ea = ea;
  return run_group(&mem, 1, M, workspace, ws_bytes, (cudaStream_t)stream);
}

extern "C" size_t quipb200_linear_workspace_bytes(const quipb200_linear_t* L, int M) {
  if (!L || M < 1) return 0;
  return carve_ws(nullptr, M, L->q_out, L->q_in, true).bytes;
}

extern "C" size_t quipb200_linear_group_workspace_bytes(const quipb200_linear_t* layers, int n, int M) {
  if (!layers || n < 1 || M < 1) return 0;
  size_t b = 0;
  for (int j = 0; j < n; j++) b += carve_ws(nullptr, M, layers[j].q_out, layers[j].q_in, true).bytes;
  return b;
}

extern "C" int quipb200_linear_group_forward(const quipb200_linear_t* layers, int n, const quipb200_fusion_t* fusion,
                                             const void* x, int64_t ldx, void* const* y, const int64_t* ldy, int M,
                                             void* workspace, size_t ws_bytes, void* stream) {
  if (!layers || !y || !ldy || n < 1 || n > QUIPB200_MAX_GROUP || M < 0) return QUIPB200_EINVAL;
  if (M == 0) return 0;
  if (M > QUIPB200_MM_MAX_M) return QUIPB200_EUNSUPPORTED;
  if (!aligned16(workspace)) return QUIPB200_EALIGN;
  Member mem[QUIPB200_MAX_GROUP];
  for (int j = 0; j < n; j++) {
    int rc = member_from_layer(&layers[j], fusion, x, ldx, y[j], ldy[j], &mem[j]);
    if (rc) return rc;
  }
  return run_group(mem, n, M, workspace, ws_bytes, (cudaStream_t)stream);
}

extern "C" int quipb200_linear_forward(const quipb200_linear_t* L, const void* x, int64_t ldx, void* y,
                                       int64_t ldy, int M, void* workspace, size_t ws_bytes, void* stream) {
  if (!L) return QUIPB200_EINVAL;
  void* ys[1] = {y};
  int64_t lds[1] = {ldy};
  return quipb200_linear_group_forward(L, 1, nullptr, x, ldx, ys, lds, M, workspace, ws_bytes, stream);
}
