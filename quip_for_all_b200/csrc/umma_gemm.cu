// Batched (prefill-side) E8P12 decode + GEMM on the 5th-generation tensor cores:
//
//     out[M, N] (fp16) = x[M, K] (fp16) . decode(Qidxs[N, K/8])^T         17 <= M <= 256
//
// Replaces, for moderate M, the reference's "decompress to a dense fp16 matrix, then cuBLAS" path
// (codebook/e8p12.py:153-155 -> origin_order.cu:837-885 + input @ W.T), which writes and re-reads
// 2*N*K bytes of dense weights per call (32 MiB for 4096 x 4096) to do a few GFLOP of math.  Here the
// packed codes (N*K/4 bytes) are the only weight traffic: every CTA decodes its 128 x 64 weight tile
// straight into the UMMA-canonical shared-memory layout (K-major, 128-byte swizzle) and one elected
// thread feeds it to tcgen05.mma (kind::f16, fp32 accumulators in TMEM).
//
//   MMA shape : D[128 weight rows x NTOK tokens] += A[128 x 16] . B[NTOK x 16]^T   (cta_group::1, M = 128)
//   grid      : (N / 128 row tiles) x (split-K <= 4 so that ~all SMs have a CTA); the K splits of a row tile form a
//               thread-block cluster (1, ksplit, 1)
//   pipeline  : STAGES shared-memory slots of 128 k.  16 producer warps decode the packed codes of the slot (table
//               lookup + sign decode + int8 -> fp16 with the reference's exact 0x5c80 trick, one 16-byte swizzled store
//               per code; codes ride three stages ahead in registers) and arrive on full[s]; the activation tile of the
//               slot comes in by TMA (cp.async.bulk.tensor, tensor map over x with SWIZZLE_128B, rows beyond M zero
//               filled), its bytes counted on the same barrier; a 17th warp waits on full[s], issues the 8 MMAs of the
//               stage and tcgen05.commit's to empty[s], which gates the slot's reuse.  No CTA-wide barrier, no thread
//               waiting on activation data in the main loop.
//   epilogue  : tcgen05.ld (32 lanes x 32 bit x 16 columns) -> fp16 store; with split-K the cluster reduce-scatters
//               the partial tiles through distributed shared memory (each CTA owns NTOK / ksplit token columns).
//
// Numerics: fp16 x fp16 products accumulated in fp32 (hardware order), one fp16 rounding -- the same
// class as the reference's mma.sync kernel and cuBLAS path; decoded weights are bit-exact.
#include <cooperative_groups.h>
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"

namespace qb {

namespace cg = cooperative_groups;
extern int g_opt_umma_ksplit;
extern int g_opt_umma_rt;

constexpr int UG_THREADS = 256;
constexpr int UG_BM = 128;   // weight rows per CTA = UMMA M
constexpr int UG_BK = 64;    // k per stage = one 128-byte swizzle row of fp16 = 8 codes

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)1 << 16;     // leading byte offset (unused for swizzled K-major): 1
  d |= (uint64_t)64 << 32;    // stride byte offset: 1024 >> 4
  d |= (uint64_t)1 << 46;     // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;     // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t sdst, const void* gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdst), "l"(gsrc), "r"(src_bytes) : "memory");
}

// TMA: one [box rows x 64 k] fp16 tile of the activation matrix (tensor map with SWIZZLE_128B: the layout the UMMA
// descriptor expects), rows beyond M zero-filled; completion is counted in bytes on `bar`
__device__ __forceinline__ void tma_load_2d(uint32_t sdst, const CUtensorMap* tmap, uint32_t bar, int crd0, int crd1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(sdst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(crd0), "r"(crd1) : "memory");
}

struct UmmaArgs {
  const unsigned char* codes;   // [N][K/8] int16 (E8P12) / int32 (E8P12RVQ4B) or [N][K/4] uint8 (D4)
  const __half* x;              // [M][K]
  const uint2* table;           // E8P family: int64[256] abs table; D4: fp16 [256][4]; HI: unused
  const uint32_t* table2;       // E8P12RVQ3B: e81b residual table, int32[256] (8 nibbles of 2 * v)
  float resid_scale;            // E8P12RVQ4B / E8P12RVQ3B
  __half* out;                  // [M][N]
  float* ws;                    // [256][N] fp32 split-K partials (zero on entry, zero on exit)
  unsigned int* tickets;        // [N/128]
  int M, N, K, ksplit, kb_per_split;
};

constexpr int UG_PRODUCERS = 512;               // 16 producer warps
constexpr int UG_THREADS2 = UG_PRODUCERS + 32;  // + 1 MMA-issue warp
constexpr int UG_BK2 = 128;                     // k per stage: two 64-wide swizzle tiles
constexpr int UG_WS_LD = 256;                   // token pitch of the split-K workspace

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Warp-specialised: 16 producer warps fill the stage slots (no CTA-wide barrier in the main loop), warp 16 waits on
// full[s], issues the 8 MMAs of the stage and commits to empty[s].
// CB: the producers' decode (QUIPB200_CB_E8P12 / _E8P12RVQ4B / _D4 / _E8P12RVQ3B / _HI); everything else is shared.
// RT: 128-row weight tiles per CTA.  RT = 1: a stage is 128 rows x 128 k (two 64-wide swizzle tiles of A and of B).
// RT = 2 (M > 64): a stage is 256 rows x 64 k -- two A tiles against ONE activation tile, two TMEM accumulators -- which
// halves the activation bytes every CTA pulls from L2 per weight (at M = 256 that re-read, not the decode or the MMAs,
// bounded the kernel: 64 KB per stage per CTA).
template <int CB, int NTOK, int STAGES, int RT>
__global__ void __launch_bounds__(UG_THREADS2, 1) e8p_umma_kernel(const __grid_constant__ UmmaArgs a,
                                                                  const __grid_constant__ CUtensorMap tmap_x) {
  extern __shared__ unsigned char smem_raw[];
  constexpr uint32_t A_SUB = UG_BM * 128, B_SUB = NTOK * 128;           // one 64-wide swizzle tile
  constexpr uint32_t A_BYTES = 2 * A_SUB, B_BYTES = (RT == 2 ? 1 : 2) * B_SUB;
  constexpr int BK = RT == 2 ? 64 : 128;                                // k per stage
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                 // swizzle atoms need 1024-byte alignment
  unsigned char* gbase = smem_raw + (base - raw);
  const uint32_t sA = base, sB = base + STAGES * A_BYTES;
  unsigned char* tab = gbase + STAGES * (A_BYTES + B_BYTES);    // 2 KB table + 1 KB residual table (RVQ3B)
  const uint32_t* tab2 = reinterpret_cast<const uint32_t*>(tab + 2048);
  const uint32_t sbar = base + STAGES * (A_BYTES + B_BYTES) + 3072;   // full[STAGES], empty[STAGES], done
  const uint32_t bar_full = sbar, bar_empty = sbar + 8 * STAGES, bar_done = sbar + 16 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + STAGES * (A_BYTES + B_BYTES) + 3072 + 8 * (2 * STAGES + 1));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * (UG_BM * RT);
  const int kb_begin = blockIdx.y * a.kb_per_split;                       // in units of BK
  const int nit = min(a.kb_per_split, a.K / BK - kb_begin);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) {
      mbar_init(bar_full + 8 * s, UG_PRODUCERS / 32 + 1);    // one arrive per producer warp + the TMA issuer's expect_tx
      mbar_init(bar_empty + 8 * s, 1);                   // tcgen05.commit
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), RT * NTOK);
  if (tid < 256 && CB != QUIPB200_CB_HI) {
    uint2 t = a.table[tid];
    if (CB != QUIPB200_CB_D4) {      // E8P abs entries with the "+1/4" pre-applied; D4: the fp16 grid rows as they are
      t.x |= 0x01010101u;
      t.y |= 0x01010101u;
    }
    reinterpret_cast<uint2*>(tab)[tid] = t;
    if (CB == QUIPB200_CB_E8P12RVQ3B) reinterpret_cast<uint32_t*>(tab + 2048)[tid] = a.table2[tid];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // instruction descriptor: D = F32, A = B = F16, both K-major, N = NTOK, M = 128 (cute::UMMA::InstrDescriptor)
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NTOK >> 3) << 17) | ((uint32_t)(UG_BM >> 4) << 24);

  if (warp == UG_PRODUCERS / 32) {
    // ===== MMA issuer =====
    if (lane == 0) {
      for (int it = 0; it < nit; it++) {
        const int s = it % STAGES;
        mbar_wait(bar_full + 8 * s, (uint32_t)((it / STAGES) & 1));
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; h++) {      // RT = 1: the two k halves of the stage; RT = 2: the two row tiles (accumulator h)
          const uint64_t ad = umma_smem_desc(sA + s * A_BYTES + h * A_SUB);
          const uint64_t bd = umma_smem_desc(sB + s * B_BYTES + (RT == 2 ? 0 : h) * B_SUB);
          const uint32_t td = tmem_base + (RT == 2 ? (uint32_t)(h * NTOK) : 0u);
#pragma unroll
          for (int k = 0; k < 4; k++)    // +32 bytes along K inside the swizzled row = +2 in the address field
            umma_f16(td, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), IDESC, (it > 0 || (RT == 1 && h > 0) || k > 0) ? 1u : 0u);
        }
        umma_commit(bar_empty + 8 * s);          // slot s is free again when these MMAs retire
      }
      umma_commit(bar_done);                     // all MMAs of this CTA
    }
    __syncwarp();
  } else {
    // ===== producers: thread -> (weight row, 32 of its 128 weights of the stage = 4 swizzle chunks); activations: 16-byte
    // chunks round-robin.  Packed bytes per thread per stage: 8 (E8P12: 4 codes, D4: 8 codes) or 16 (RVQ4B: 4 codes).
    // Packed bytes per thread per stage: 8 (E8P12 4 codes, D4 8 codes), 16 (RVQ4B 4 codes, HI 4 words), 12 (RVQ3B 4 codes).
    constexpr int CBYTES = (CB == QUIPB200_CB_E8P12RVQ4B || CB == QUIPB200_CB_HI) ? 16 : (CB == QUIPB200_CB_E8P12RVQ3B ? 12 : 8);
    // RT = 1: thread -> row tid / 4, k quarter tid % 4 of the stage's 128;  RT = 2: row tid / 2 (of 256), k half tid % 2 of 64
    const int wrow = RT == 2 ? (tid >> 1) : (tid >> 2), wq = RT == 2 ? (tid & 1) : (tid & 3);
    const int asub = RT == 2 ? (wrow >> 7) : (wq >> 1);              // which 16 KB A tile of the stage
    const size_t row_bytes = (size_t)(a.K >> 5) * CBYTES;            // CBYTES per 32 weights
    const unsigned char* wsrc = a.codes + (size_t)(n0 + wrow) * row_bytes + wq * CBYTES;
    const uint64_t pol = l2_evict_first_policy();
    auto load_codes = [&](int it) -> uint4 {
      const unsigned char* p = wsrc + (size_t)(kb_begin + it) * ((BK / 32) * CBYTES);
      if (CBYTES == 16) return ldg_stream_v4(p, pol);
      if (CBYTES == 12) {      // 4-byte aligned only (row pitch 3K/8)
        const uint32_t* p4 = reinterpret_cast<const uint32_t*>(p);
        return make_uint4(__ldg(p4), __ldg(p4 + 1), __ldg(p4 + 2), 0u);
      }
      const uint2 v = ldg_stream_v2(p, pol);
      return make_uint4(v.x, v.y, 0u, 0u);
    };
    // activations: thread 0 posts the stage's byte count on full[s] and issues the tile loads; the copy engine does the
    // rest (address generation, swizzle, zero fill of the rows beyond M), so no thread ever waits on activation data
    auto issue_acts = [&](int it) {
      if (tid != 0) return;
      const int s = it % STAGES;
      mbar_arrive_expect_tx(bar_full + 8 * s, B_BYTES);
#pragma unroll
      for (int h = 0; h < (RT == 2 ? 1 : 2); h++)
        tma_load_2d(sB + s * B_BYTES + h * B_SUB, &tmap_x, bar_full + 8 * s, (kb_begin + it) * BK + h * 64, 0);
    };
    // packed codes come from HBM (read once, ~1 us away): three stages of them ride in registers ahead of the decode
    uint4 cur = make_uint4(0, 0, 0, 0), nx1 = cur, nx2 = cur;
    const __half2 rs2 = __float2half2_rn(a.resid_scale);      // the reference's fp16 hfma2 operand (origin_order.cu:378)
    if (nit > 0) {
      cur = load_codes(0);
      issue_acts(0);
    }
    if (nit > 1) nx1 = load_codes(1);
    if (nit > 2) nx2 = load_codes(2);
    for (int it = 0; it < nit; it++) {
      const int s = it % STAGES;
      uint4 nx3 = make_uint4(0, 0, 0, 0);
      if (it + 3 < nit) nx3 = load_codes(it + 3);
      if (it + 1 < nit) {
        if (it + 1 >= STAGES) mbar_wait(bar_empty + 8 * ((it + 1) % STAGES), (uint32_t)(((it + 1) / STAGES - 1) & 1));
        issue_acts(it + 1);
      }
      // weights of stage `it` -> slot s (free: empty[s] was waited on one iteration ago, or it < STAGES)
      {
        const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
        unsigned char* arow = gbase + (size_t)s * A_BYTES + asub * A_SUB + (wrow & 127) * 128;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          uint4 v;
          if (CB == QUIPB200_CB_HI) {            // 8 nibbles of one word: w = nib - 7.5, elements (0,2,4,6,1,3,5,7) (hi.py:41-50)
            uint32_t qa = w[j];
            const uint32_t c0 = 0x64086408u;     // 1024 + 8 (+ 16 * nibble); the reference's constants (origin_order.cu:1028-1051)
            const __half2 y16 = __float2half2_rn(1.0f / 16.0f), z16 = __float2half2_rn(-1024.0f / 16.0f - 8.0f);
            uint32_t hw4[4];
            hw4[0] = ((qa & 0x000f000fu) << 4) | c0;
            hw4[1] = (qa & 0x00f000f0u) | c0;
            qa >>= 8;
            hw4[2] = ((qa & 0x000f000fu) << 4) | c0;
            hw4[3] = (qa & 0x00f000f0u) | c0;
            uint32_t o4[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const __half2 h = __hfma2(*reinterpret_cast<const __half2*>(&hw4[e]), y16, z16);
              o4[e] = *reinterpret_cast<const uint32_t*>(&h);
            }
            v = make_uint4(o4[0], o4[1], o4[2], o4[3]);
          } else if (CB == QUIPB200_CB_E8P12RVQ3B) {   // byte triplets [resid, main lo, main hi] (e8p12_rvq3.py:97-107)
            const uint32_t lo = w[(3 * j) >> 2], hi = w[((3 * j) >> 2) + 1 < 3 ? ((3 * j) >> 2) + 1 : 2];
            const uint32_t c24 = __funnelshift_r(lo, hi, ((3 * j) & 3) * 8) & 0xffffffu;
            const uint32_t rc = c24 & 0xffu, code = c24 >> 8;
            const uint2 t1 = *reinterpret_cast<const uint2*>(tab + ((code >> 8) << 3));
            const uint2 q = e8p_decode_q(t1, code);
            __half2 e0, o0, e1, o1;
            q4_to_half2(q.x, e0, o0);
            q4_to_half2(q.y, e1, o1);
            const uint32_t c = tab2[rc];          // 8 nibbles of 2 * v (two's complement), nibble j + 4 h <-> element 2 j + h
            const __half2 adj = __float2half2_rn(-516.0f);
            __half2 r4[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
              uint32_t b = ((c >> (4 * e)) & 0x000f000fu) ^ 0x60086008u;      // 512 + (nib ^ 8) / 2 (origin_order.cu:323-327)
              r4[e] = __hadd2(*reinterpret_cast<const __half2*>(&b), adj);
            }
            e0 = __hfma2(rs2, r4[0], e0); o0 = __hfma2(rs2, r4[1], o0);
            e1 = __hfma2(rs2, r4[2], e1); o1 = __hfma2(rs2, r4[3], o1);
            v.x = *reinterpret_cast<const uint32_t*>(&e0);
            v.y = *reinterpret_cast<const uint32_t*>(&o0);
            v.z = *reinterpret_cast<const uint32_t*>(&e1);
            v.w = *reinterpret_cast<const uint32_t*>(&o1);
          } else if (CB == QUIPB200_CB_D4) {            // two 1-byte codes -> 2 x 4 fp16 weights (table rows, natural order)
            const uint32_t b0 = (w[j >> 1] >> ((j & 1) * 16)) & 0xffu, b1 = (w[j >> 1] >> ((j & 1) * 16 + 8)) & 0xffu;
            const uint2 g0 = *reinterpret_cast<const uint2*>(tab + (b0 << 3)), g1 = *reinterpret_cast<const uint2*>(tab + (b1 << 3));
            v = make_uint4(g0.x, g0.y, g1.x, g1.y);
          } else {
            const uint32_t code = (CB == QUIPB200_CB_E8P12RVQ4B) ? (w[j] >> 16) : ((w[j >> 1] >> ((j & 1) * 16)) & 0xffffu);
            const uint2 t1 = *reinterpret_cast<const uint2*>(tab + ((code >> 8) << 3));
            const uint2 q = e8p_decode_q(t1, code);
            __half2 e0, o0, e1, o1;
            q4_to_half2(q.x, e0, o0);     // weights (0,1) = bytes (0,2); (2,3) = bytes (1,3)
            q4_to_half2(q.y, e1, o1);
            if (CB == QUIPB200_CB_E8P12RVQ4B) {   // W = g[main] + fp16(scale) * g[resid], one fp16 fma rounding (e8p12_rvq4.py:23)
              const uint32_t rc = w[j] & 0xffffu;
              const uint2 t2 = *reinterpret_cast<const uint2*>(tab + ((rc >> 8) << 3));
              const uint2 q2 = e8p_decode_q(t2, rc);
              __half2 re0, ro0, re1, ro1;
              q4_to_half2(q2.x, re0, ro0);
              q4_to_half2(q2.y, re1, ro1);
              e0 = __hfma2(rs2, re0, e0); o0 = __hfma2(rs2, ro0, o0);
              e1 = __hfma2(rs2, re1, e1); o1 = __hfma2(rs2, ro1, o1);
            }
            v.x = *reinterpret_cast<const uint32_t*>(&e0);
            v.y = *reinterpret_cast<const uint32_t*>(&o0);
            v.z = *reinterpret_cast<const uint32_t*>(&e1);
            v.w = *reinterpret_cast<const uint32_t*>(&o1);
          }
          const int ch = (wq & 1) * 4 + j;
          *reinterpret_cast<uint4*>(arow + ((ch ^ (wrow & 7)) << 4)) = v;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * s);
      cur = nx1; nx1 = nx2; nx2 = nx3;
    }
  }
  mbar_wait(bar_done, 0);
  tc_fence_after();

  // ---- epilogue: TMEM lane = weight row (of row tile `sub`), column = sub * NTOK + token.  Producer warp w: lane quadrant
  // w % 4, column quarter w / 4.
  const int quad = warp & 3, cq = warp >> 2;
  constexpr int CW = NTOK / 4 < 16 ? 16 : NTOK / 4;        // columns per warp (>= one 16-column load)
  if (a.ksplit == 1) {
    if (warp < UG_PRODUCERS / 32 && nit > 0) {
#pragma unroll
      for (int sub = 0; sub < RT; sub++) {
        const int n = n0 + sub * UG_BM + quad * 32 + lane;
        for (int c0 = cq * CW; c0 < (cq + 1) * CW && c0 < NTOK; c0 += 16) {
          if (c0 >= a.M) break;                        // warp-uniform: token columns beyond M hold zeros
          float v[16];
          tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(sub * NTOK + c0), v);
#pragma unroll
          for (int j = 0; j < 16; j++)
            if (c0 + j < a.M) a.out[(size_t)(c0 + j) * a.N + n] = __float2half_rn(v[j]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, RT * NTOK);
    return;
  }
  // ---- split-K: the ksplit CTAs of a row tile form a cluster (1, ksplit, 1) and reduce-scatter their accumulators through
  // distributed shared memory: CTA r owns the token columns [r * NTOK / ksplit, ...), every CTA pushes each 16-column
  // group of its TMEM tile into the owner's receive buffer (the stage ring, idle now), the owner adds the ksplit partial
  // tiles and stores fp16.  No atomics, no workspace, no second pass through L2.  (RT = 2: one round per row tile.)
  {
    cg::cluster_group cluster = cg::this_cluster();
    const int r = (int)cluster.block_rank();
    const int cols_per = NTOK / a.ksplit;                       // >= 4 (NTOK >= 32, ksplit <= 8)
    float* recv = reinterpret_cast<float*>(gbase);              // [ksplit sources][cols_per][128 rows]
    const int tok0 = r * cols_per;
#pragma unroll 1
    for (int sub = 0; sub < RT; sub++) {
      cluster.sync();                                           // every CTA is done with its stage ring / the previous round
      if (warp < UG_PRODUCERS / 32) {
        const int row = quad * 32 + lane;
        for (int c0 = cq * CW; c0 < (cq + 1) * CW && c0 < NTOK; c0 += 16) {
          if (c0 >= a.M) break;
          float v[16];
          if (nit > 0) {
            tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(sub * NTOK + c0), v);
          } else {                                               // (a split with no k-blocks contributes zeros)
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 16; j++) {
            const int c = c0 + j;
            const int d = c / cols_per, lc = c - d * cols_per;
            float* dst = cluster.map_shared_rank(recv, d) + ((size_t)(r * cols_per + lc) << 7) + row;
            *dst = v[j];
          }
        }
      }
      cluster.sync();                                           // all partial tiles of this round delivered
      for (int i = tid; i < cols_per * UG_BM; i += UG_THREADS2) {
        const int lc = i >> 7, row = i & 127;
        const int tok = tok0 + lc;
        if (tok >= a.M) break;
        float acc = 0.f;
        for (int src = 0; src < a.ksplit; src++) acc += recv[((size_t)(src * cols_per + lc) << 7) + row];
        a.out[(size_t)tok * a.N + n0 + sub * UG_BM + row] = __float2half_rn(acc);
      }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, RT * NTOK);
    cluster.sync();                                             // nobody exits while a peer may still read its buffer
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (the library does not link libcuda)
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmapEncodeFn tmap_encoder() {
  static TmapEncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (TmapEncodeFn)p;
  }
  return fn;
}

template <int CB, int NTOK, int STAGES, int RT>
static int launch_umma(const UmmaArgs& a, dim3 grid, cudaStream_t st) {
  const size_t smem = (size_t)STAGES * (2 * UG_BM * 128 + (RT == 2 ? 1 : 2) * NTOK * 128) + 3072 + 8 * (2 * STAGES + 1) + 16 + 1024;
  const void* fn = (const void*)e8p_umma_kernel<CB, NTOK, STAGES, RT>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  // x as a 2-D tensor {K (contiguous), M}; one box = 64 k x NTOK token rows = one 128-byte-swizzled UMMA B tile
  TmapEncodeFn enc = tmap_encoder();
  if (!enc) return QUIPB200_EUNSUPPORTED;
  alignas(64) CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)a.K, (cuuint64_t)a.M};
  const cuuint64_t gstride[1] = {(cuuint64_t)a.K * sizeof(__half)};
  const cuuint32_t box[2] = {64u, (cuuint32_t)NTOK};
  const cuuint32_t estr[2] = {1u, 1u};
  if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(a.x), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return QUIPB200_EINVAL;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(UG_THREADS2);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;      // the split-K CTAs of one row tile
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = grid.y;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  void* args[] = {const_cast<UmmaArgs*>(&a), &tmap};
  e = cudaLaunchKernelExC(&cfg, fn, args);
  if (e != cudaSuccess) return (int)e;
  QB_LAUNCH_CHECK();
  return 0;
}
// rt: row tiles per CTA chosen by the host (2 needs N % 256 == 0)
template <int CB>
static int launch_umma_m(const UmmaArgs& a, dim3 grid, int rt, cudaStream_t st) {
  if (rt == 2) {
    if (a.M <= 128) return launch_umma<CB, 128, 4, 2>(a, grid, st);
    return launch_umma<CB, 256, 3, 2>(a, grid, st);
  }
  if (a.M <= 32) return launch_umma<CB, 32, 4, 1>(a, grid, st);
  if (a.M <= 64) return launch_umma<CB, 64, 4, 1>(a, grid, st);
  if (a.M <= 128) return launch_umma<CB, 128, 3, 1>(a, grid, st);
  return launch_umma<CB, 256, 2, 1>(a, grid, st);
}

}  // namespace qb

using namespace qb;

extern "C" size_t quipb200_e8p_mm_umma_workspace_bytes(int M, int N, int K) {
  (void)M; (void)N; (void)K;
  return 0;   // split-K partial tiles are reduced through distributed shared memory: no workspace since round 2
}

extern "C" int quipb200_mm_umma(int codebook, const void* x, const void* qidxs, const void* grid, const void* grid2, float scale,
                                void* out, int M, int N, int K, void* workspace, size_t ws_bytes, void* stream) {
  if (!x || !qidxs || !out) return QUIPB200_EINVAL;
  if (codebook < QUIPB200_CB_E8P12 || codebook > QUIPB200_CB_HI) return QUIPB200_EUNSUPPORTED;
  if (codebook != QUIPB200_CB_HI && !grid) return QUIPB200_EINVAL;
  if (codebook == QUIPB200_CB_E8P12RVQ3B && !grid2) return QUIPB200_EINVAL;
  if (M < 1 || M > 256 || N < UG_BM || N % UG_BM || K < UG_BK2 || K % UG_BK2) return QUIPB200_EUNSUPPORTED;
  if (!aligned16(x) || !aligned16(qidxs) || !aligned16(grid) || !aligned16(out) || ((uintptr_t)grid2 & 3)) return QUIPB200_EALIGN;
  const int sms = quipb200_sm_count();
  if (sms < 1) return (int)cudaErrorNoDevice;
  // (option umma_rt=2: 256 rows x 64 k per stage.  Measured slower at every shape -- the pipeline is bound by per-stage
  // latency, and halving k per stage doubles the stage count -- so one row tile per CTA is the default.)
  const int rt = (g_opt_umma_rt == 2 && M > 64 && N % (2 * UG_BM) == 0) ? 2 : 1;
  const int tiles = N / (UG_BM * rt), nkb = K / (rt == 2 ? 64 : UG_BK2);
  int ksplit = 1;
  // <= 4: clusters of 8 of these one-per-SM CTAs do not all fit in one wave (measured 2x slower at 4096 x 4096)
  while (ksplit < 4 && tiles * ksplit * 2 <= sms && nkb / (ksplit * 2) >= 4 * rt) ksplit *= 2;
  if (g_opt_umma_ksplit > 0 && g_opt_umma_ksplit <= 8 && !(g_opt_umma_ksplit & (g_opt_umma_ksplit - 1)) &&
      nkb / g_opt_umma_ksplit >= 1)
    ksplit = g_opt_umma_ksplit;
  UmmaArgs a{};
  a.codes = (const unsigned char*)qidxs; a.x = (const __half*)x; a.table = (const uint2*)grid; a.out = (__half*)out;
  a.table2 = (const uint32_t*)grid2;
  a.resid_scale = scale;
  a.M = M; a.N = N; a.K = K;
  a.ksplit = ksplit;
  a.kb_per_split = (nkb + ksplit - 1) / ksplit;
  (void)workspace; (void)ws_bytes;   // split-K partial tiles are reduced through distributed shared memory: no workspace
  const dim3 grid_dim(tiles, ksplit);
  cudaStream_t st = (cudaStream_t)stream;
  if (codebook == QUIPB200_CB_E8P12) return launch_umma_m<QUIPB200_CB_E8P12>(a, grid_dim, rt, st);
  if (codebook == QUIPB200_CB_E8P12RVQ4B) return launch_umma_m<QUIPB200_CB_E8P12RVQ4B>(a, grid_dim, rt, st);
  if (codebook == QUIPB200_CB_E8P12RVQ3B) return launch_umma_m<QUIPB200_CB_E8P12RVQ3B>(a, grid_dim, rt, st);
  if (codebook == QUIPB200_CB_HI) return launch_umma_m<QUIPB200_CB_HI>(a, grid_dim, rt, st);
  return launch_umma_m<QUIPB200_CB_D4>(a, grid_dim, rt, st);
}

extern "C" int quipb200_e8p_mm_umma(const void* x, const void* qidxs, const void* grid, void* out, int M, int N, int K,
                                    void* workspace, size_t ws_bytes, void* stream) {
  return quipb200_mm_umma(QUIPB200_CB_E8P12, x, qidxs, grid, nullptr, 0.f, out, M, N, K, workspace, ws_bytes, stream);
}
