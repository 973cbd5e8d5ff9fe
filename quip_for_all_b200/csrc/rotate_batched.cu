// Batched fused rotation for the M >= 17 path of QuantLinear.forward (qlinear.py:90-114):
//
//   input side :  y = f16( (hadK^T (x) H_L) f16(x * SU) * scale )                       quant.py:72-88, qlinear.py:91,99
//   output side:  y = f16( f16( (hadK (x) H_L) x * scale )[:out] * SV ) + bias          qlinear.py:108-114
//
// The reference runs these as 3-6 separate passes over the [M, n] activation tensor (x*SU, F.pad, the external
// fast_hadamard_transform kernel, hadK.T.contiguous(), a batched K x K matmul, the slice, *SV, +bias): at prefill
// sizes (M = 65 536) that is several GB of HBM traffic per linear and costs more than the GEMM itself
// (profiles/README.md).  Here each row is read once and written once: one 512-thread CTA per row, the transform on
// the legacy tensor path (fwht_mma.cuh: H_4096 = H_16^(x)3 with one shared-memory exchange, or K blocks of H_256 =
// H_16^(x)2 held by one warp each), the K x K mix as mma.sync tiles in shared memory, SU / SV / bias applied in
// registers with the reference's fp16 rounding points.  HBM-bound by construction: 2 * n * 2 bytes per row.
//
// Covered: n == 4096 with K == 1, and n == K * 256 with K <= 64 (Llama-2-7B: 4096 and 11008 = 43 * 256).  Anything
// else returns QUIPB200_EUNSUPPORTED and the caller keeps the reference's op sequence (quipb200_hadamard + matmul).
#include "fwht_mma.cuh"
#include "ql_device.cuh"

namespace qb {

constexpr int RB_THREADS = 512;
constexpr int RB_WARPS = RB_THREADS / 32;
constexpr int RB_LS = 256 + 8;      // row stride (halfs) of the block buffer: ldmatrix friendly

int g_opt_rot_pipe_rows = 1024;   // rows from which the K x 256 rotation streams rows through shared memory (rotblk_pipe_kernel)
int g_opt_rot_warp_rows = 2048;   // rows from which the n = 4096 rotation runs one warp per row (rot4096w_kernel)

struct RotBArgs {
  const __half* x; int64_t ldx;     // [M][in_feat]
  __half* y; int64_t ldy;           // [M][out_feat]
  const __half* pre;                // [in_feat]  or NULL  (SU)
  const __half* post;               // [out_feat] or NULL  (SV)
  const __half* bias;               // [out_feat] or NULL
  const __half* hk;                 // [Kp][Kp] zero-padded coefficient matrix M[k_out][k_in], or NULL when K == 1
  int M, in_feat, out_feat, n, K;
  float post_scale;                 // scale * sqrt(n / K): what is left after the 1/sqrt(L) built into the H_16/4 factors
};

// ---- n == 4096, K == 1: thread t owns octet t of the row --------------------------------------------------
__global__ void __launch_bounds__(RB_THREADS) rot4096_kernel(const __grid_constant__ RotBArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* S = reinterpret_cast<float*>(smem);                                   // 16 x DS_XROW floats
  __half* V = reinterpret_cast<__half*>(smem + 16 * DS_XROW * sizeof(float));  // [4096]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const HFrag hf = make_hfrag(lane);
  const bool vin = tid * 8 < a.in_feat, vout = tid * 8 < a.out_feat;
  uint4 pre = make_uint4(0, 0, 0, 0), post = pre, bias = pre;
  if (a.pre && vin) pre = __ldg(reinterpret_cast<const uint4*>(a.pre) + tid);
  if (a.post && vout) post = __ldg(reinterpret_cast<const uint4*>(a.post) + tid);
  if (a.bias && vout) bias = __ldg(reinterpret_cast<const uint4*>(a.bias) + tid);
  const float sc = a.post_scale;
  uint4 nxt = make_uint4(0, 0, 0, 0);
  if (vin && (int)blockIdx.x < a.M) nxt = ldg_stream_v4(a.x + (size_t)blockIdx.x * a.ldx + tid * 8);
  for (int row = blockIdx.x; row < a.M; row += gridDim.x) {
    uint4 oct = nxt;
    const int rown = row + gridDim.x;      // the next row's octet is in flight while this one is transformed
    if (vin && rown < a.M) nxt = ldg_stream_v4(a.x + (size_t)rown * a.ldx + tid * 8);
    if (a.pre) oct = hmul2x4(oct, pre);                                        // qlinear.py:91 (fp16 tensor)
    const uint32_t p[4] = {oct.x, oct.z, oct.y, oct.w};                        // block layout, natural placement (fwht_mma.cuh)
    float r[8];
    fwht4096_frag(p, hf, S, warp, lane, r);                                    // spread layout, x 1/64
#pragma unroll
    for (int xh = 0; xh < 2; xh++) {                                           // pairs xh and xh + 2 are adjacent
      const int i = idx_spread(warp, lane, xh);
      *reinterpret_cast<uint2*>(V + stg_chunk(i >> 3) * 8 + (i & 7)) =
          make_uint2(pk_h2(r[2 * xh] * sc, r[2 * xh + 1] * sc), pk_h2(r[2 * xh + 4] * sc, r[2 * xh + 5] * sc));
    }
    __syncthreads();
    if (vout) {
      uint4 o = *reinterpret_cast<const uint4*>(V + stg_chunk(tid) * 8);
      if (a.post) o = hmul2x4(o, post);                                        // qlinear.py:112
      if (a.bias) o = hadd2x4(o, bias);                                        // qlinear.py:114
      stg_stream_v4(a.y + (size_t)row * a.ldy + tid * 8, o);
    }
  }
}

// ---- n == 4096, K == 1, many rows: one WARP per row, no shared-memory exchange, no barrier ------------------------
// The F(x, y) register layout of fwht_mma.cuh is a bit permutation of the natural index, and a Sylvester Hadamard
// matrix is invariant under a simultaneous bit permutation of its row and column index (H[i][j] = (-1)^popc(i & j)).
// So a lane that loads octet `lane` of a 256-element block (16 bytes) already holds a valid fragment: its pairs go into
// fwht256_frag as they are and come back as the transformed elements of the same octet (tools/emu_fwht_frag.py checks
// this with a numpy model of the mma fragments).  A warp holds the 16 blocks of a row, transforms each on the tensor
// path and finishes with a radix-16 butterfly across the blocks in fp32 registers.  Rows in flight per SM: 12 (against
// 2 for the CTA-per-row kernel, which stays for small M where one row per warp would leave most SMs idle).
constexpr int RW_THREADS = 128;
constexpr int RW_WARPS = RW_THREADS / 32;
__global__ void __launch_bounds__(RW_THREADS, 3) rot4096w_kernel(const __grid_constant__ RotBArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint4* s_pre = reinterpret_cast<uint4*>(smem);        // [512] octets each; only the vectors that exist are filled
  uint4* s_post = s_pre + 512;
  uint4* s_bias = s_post + 512;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int noct_in = a.in_feat >> 3, noct_out = a.out_feat >> 3;
  for (int i = tid; i < 512; i += RW_THREADS) {
    const uint4 z4 = make_uint4(0, 0, 0, 0);
    if (a.pre) s_pre[i] = (i < noct_in) ? __ldg(reinterpret_cast<const uint4*>(a.pre) + i) : z4;
    if (a.post) s_post[i] = (i < noct_out) ? __ldg(reinterpret_cast<const uint4*>(a.post) + i) : z4;
    if (a.bias) s_bias[i] = (i < noct_out) ? __ldg(reinterpret_cast<const uint4*>(a.bias) + i) : z4;
  }
  __syncthreads();
  const HFrag hf = make_hfrag(lane);
  const float sc = a.post_scale * 0.25f;      // two H_16/4 factors on the tensor path; the third factor is the plain butterfly
  for (int row = blockIdx.x * RW_WARPS + warp; row < a.M; row += gridDim.x * RW_WARPS) {
    const __half* xr = a.x + (size_t)row * a.ldx;
    uint4 oct[16];
#pragma unroll
    for (int z = 0; z < 16; z++) {
      const int o = z * 32 + lane;
      oct[z] = make_uint4(0, 0, 0, 0);
      if (o < noct_in) oct[z] = ldg_stream_v4(xr + o * 8);
    }
    float R[16][8];
#pragma unroll
    for (int z = 0; z < 16; z++) {
      uint4 v = oct[z];
      if (a.pre) v = hmul2x4(v, s_pre[z * 32 + lane]);                          // qlinear.py:91 (fp16 tensor)
      const uint32_t p[4] = {v.x, v.z, v.y, v.w};      // (x, y) and (z, w) stay register pairs: the two B fragments
      fwht256_frag(p, hf, R[z]);                                                // x 1/16
    }
#pragma unroll
    for (int h = 1; h < 16; h <<= 1) {
#pragma unroll
      for (int z = 0; z < 16; z++) {
        if (z & h) continue;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float u = R[z][j], w = R[z | h][j];
          R[z][j] = u + w;
          R[z | h][j] = u - w;
        }
      }
    }
    __half* yr = a.y + (size_t)row * a.ldy;
#pragma unroll
    for (int z = 0; z < 16; z++) {
      const int o = z * 32 + lane;
      uint4 v = frag_to_octet(R[z], sc);                                        // register_lib.py:20 (fp16 out)
      if (a.post) v = hmul2x4(v, s_post[o]);                                    // qlinear.py:112
      if (a.bias) v = hadd2x4(v, s_bias[o]);                                    // qlinear.py:114
      if (o < noct_out) stg_stream_v4(yr + o * 8, v);
    }
  }
}

// ---- n == K * 256: warp w owns blocks w, w + 16, ..; K x K mix on the tensor path in shared memory ----------
__global__ void __launch_bounds__(RB_THREADS) rotblk_kernel(const __grid_constant__ RotBArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int K = a.K, Kp = (K + 15) / 16 * 16;
  __half* T = reinterpret_cast<__half*>(smem);                                 // [(K + 1)][RB_LS]
  __half* hk = T + (size_t)(K + 1) * RB_LS;                                    // [Kp][Kp]
  float* dummy = reinterpret_cast<float*>(hk + (size_t)Kp * Kp);               // 64 floats (unused reduction scratch)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const HFrag hf = make_hfrag(lane);
  if (K > 1)
    for (int i = tid; i < (Kp * Kp) >> 3; i += RB_THREADS)
      reinterpret_cast<uint4*>(hk)[i] = __ldg(reinterpret_cast<const uint4*>(a.hk) + i);
  const int noct_in = a.in_feat >> 3, noct_out = a.out_feat >> 3;
  const float sc = a.post_scale;
  RotSmem rs;
  rs.s = dummy; rs.s2 = dummy; rs.pp = 1; rs.t = T; rs.hk = hk; rs.red = dummy; rs.Ls = RB_LS; rs.log2L = 8;
  constexpr int NB = 4;                      // blocks per warp (K <= 64)
  uint4 nxt[NB];
#pragma unroll
  for (int j = 0; j < NB; j++) {
    const int oi = (warp + j * RB_WARPS) * 32 + lane;
    nxt[j] = make_uint4(0, 0, 0, 0);
    if (warp + j * RB_WARPS < K && oi < noct_in && (int)blockIdx.x < a.M)
      nxt[j] = ldg_stream_v4(a.x + (size_t)blockIdx.x * a.ldx + oi * 8);
  }
  for (int row = blockIdx.x; row < a.M; row += gridDim.x) {
    uint4 cur[NB];
    const int rown = row + gridDim.x;        // the next row's octets are in flight while this row is transformed
#pragma unroll
    for (int j = 0; j < NB; j++) {
      cur[j] = nxt[j];
      const int oi = (warp + j * RB_WARPS) * 32 + lane;
      nxt[j] = make_uint4(0, 0, 0, 0);
      if (warp + j * RB_WARPS < K && oi < noct_in && rown < a.M) nxt[j] = ldg_stream_v4(a.x + (size_t)rown * a.ldx + oi * 8);
    }
#pragma unroll
    for (int j = 0; j < NB; j++) {
      const int b = warp + j * RB_WARPS;
      if (b >= K) break;
      const int oi = b * 32 + lane;
      uint4 oct = cur[j];
      if (a.pre && oi < noct_in) oct = hmul2x4(oct, __ldg(reinterpret_cast<const uint4*>(a.pre) + oi));
      // the octet is already a valid fragment (see rot4096w_kernel): pairs in, the same octet's transformed pairs out
      const uint32_t p[4] = {oct.x, oct.z, oct.y, oct.w};
      float r[8];
      fwht256_frag(p, hf, r);                                                  // x 1/16
      *reinterpret_cast<uint4*>(T + b * RB_LS + lane * 8) = frag_to_octet(r, sc);   // register_lib.py:20 (fp16 out)
    }
    if (K > 1) rotate_mix(rs, K * 256, K, tid, RB_THREADS);                    // barrier, mma.sync tiles (quant.py:83), barrier
    else __syncthreads();
    __half* yr = a.y + (size_t)row * a.ldy;
    for (int o = tid; o < noct_out; o += RB_THREADS) {
      uint4 v = *reinterpret_cast<const uint4*>(T + (o >> 5) * RB_LS + (o & 31) * 8);
      if (a.post) v = hmul2x4(v, __ldg(reinterpret_cast<const uint4*>(a.post) + o));
      if (a.bias) v = hadd2x4(v, __ldg(reinterpret_cast<const uint4*>(a.bias) + o));
      stg_stream_v4(yr + o * 8, v);
    }
    __syncthreads();      // T is rewritten by the next row
  }
}

// ---- n == K * 256, many rows: persistent CTA per SM, rows streamed through shared memory by bulk copies (TMA) --------
// One I/O warp moves whole rows: cp.async.bulk global -> shared (mbarrier complete_tx) into one of RP_SLOTS row buffers,
// cp.async.bulk shared -> global for the finished row.  The 16 compute warps form two groups of 8 that take alternate
// rows (the barrier waits of one group are filled by the other) and work in place on the row buffer:
//   P1  H_256 of every block (one warp per block, natural placement), x scale, fp16; block b stores its 16-byte chunk c
//       at chunk c ^ (b & 7) so the K x K mix can ldmatrix a column tile without bank conflicts on a dense buffer
//   P2  K x K mix of each 8-column tile (mma.sync, the A fragments of the coefficient matrix stay in registers)
//   P3  chunk un-swizzle + SV / bias, fence to the async proxy, arrive on the slot's `done` barrier
// Two named barriers per row inside a group; loads, stores and arithmetic of different rows overlap through the slot
// ring.  Same arithmetic and rounding points as rotblk_kernel.
constexpr int RP_SLOTS = 4;
constexpr int RP_COMPUTE = 512;
constexpr int RP_THREADS = RP_COMPUTE + 32;

constexpr int RP_GROUPS = 2;                          // row groups inside the CTA
constexpr int RP_GW = RP_COMPUTE / 32 / RP_GROUPS;    // warps per group
static_assert(RP_GW == 8, "block b of a warp must keep (b & 7) == warp-in-group");
__device__ __forceinline__ void rp_bar(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(RP_GW * 32) : "memory"); }

template <int MT>
__global__ void __launch_bounds__(RP_THREADS, 1) rotblk_pipe_kernel(const __grid_constant__ RotBArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int Kp = MT * 16;
  const int K = a.K;
  const int row_halfs = K * 256;
  __half* buf = reinterpret_cast<__half*>(smem);                                  // [RP_SLOTS][K][256], dense
  __half* zrow = buf + (size_t)RP_SLOTS * row_halfs;                              // [256] zeros: the rows >= K of a tile
  __half* hk = zrow + 256;                                                        // [Kp][Kp]
  __half* s_pre = hk + Kp * Kp;                                                   // [K * 256] each, if present
  __half* s_post = s_pre + (a.pre ? row_halfs : 0);
  __half* s_bias = s_post + (a.post ? row_halfs : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + (a.bias ? row_halfs : 0));   // full[RP_SLOTS], done[RP_SLOTS]
  const uint32_t bar_full = smem_u32(bars), bar_done = smem_u32(bars + RP_SLOTS);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int noct_in = a.in_feat >> 3, noct_out = a.out_feat >> 3, noct = row_halfs >> 3;

  if (tid == 0) {
    for (int s = 0; s < RP_SLOTS; s++) {
      mbar_init(bar_full + 8 * s, 1);                    // the I/O thread's arrive.expect_tx
      mbar_init(bar_done + 8 * s, RP_GW);                // one arrive per warp of the group that owns the row
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < (Kp * Kp) >> 3; i += RP_THREADS)
    reinterpret_cast<uint4*>(hk)[i] = __ldg(reinterpret_cast<const uint4*>(a.hk) + i);
  for (int i = tid; i < 32; i += RP_THREADS) reinterpret_cast<uint4*>(zrow)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < noct; i += RP_THREADS) {
    const uint4 z4 = make_uint4(0, 0, 0, 0);
    if (a.pre) reinterpret_cast<uint4*>(s_pre)[i] = (i < noct_in) ? __ldg(reinterpret_cast<const uint4*>(a.pre) + i) : z4;
    if (a.post) reinterpret_cast<uint4*>(s_post)[i] = (i < noct_out) ? __ldg(reinterpret_cast<const uint4*>(a.post) + i) : z4;
    if (a.bias) reinterpret_cast<uint4*>(s_bias)[i] = (i < noct_out) ? __ldg(reinterpret_cast<const uint4*>(a.bias) + i) : z4;
  }
  __syncthreads();

  const int n_my = (a.M - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // rows blockIdx.x, + gridDim.x, ..
  const uint32_t bytes_in = (uint32_t)a.in_feat * 2u, bytes_out = (uint32_t)a.out_feat * 2u;

  if (warp == RP_COMPUTE / 32) {
    // ---------------- I/O warp: one thread drives the slot ring ----------------
    if (lane == 0) {
      int issued = 0;
      for (; issued < n_my && issued < RP_SLOTS; issued++) {
        const size_t row = (size_t)blockIdx.x + (size_t)issued * gridDim.x;
        mbar_arrive_expect_tx(bar_full + 8 * issued, bytes_in);
        bulk_load(smem_u32(buf + (size_t)issued * row_halfs), a.x + row * a.ldx, bytes_in, bar_full + 8 * issued);
      }
      for (int done = 0; done < n_my; done++) {
        const int s = done % RP_SLOTS;
        mbar_wait(bar_done + 8 * s, (uint32_t)(done / RP_SLOTS) & 1u);
        const size_t row = (size_t)blockIdx.x + (size_t)done * gridDim.x;
        bulk_store(a.y + row * a.ldy, smem_u32(buf + (size_t)s * row_halfs), bytes_out);
        if (issued < n_my) {                             // slot s is refilled with row `issued` == done + RP_SLOTS
          bulk_store_wait_read();
          const size_t rown = (size_t)blockIdx.x + (size_t)issued * gridDim.x;
          mbar_arrive_expect_tx(bar_full + 8 * s, bytes_in);
          bulk_load(smem_u32(buf + (size_t)s * row_halfs), a.x + rown * a.ldx, bytes_in, bar_full + 8 * s);
          issued++;
        }
      }
      bulk_store_wait_all();
    }
    return;
  }

  // ---------------- compute warps: two groups of 8, each with its own rows (even / odd) and named barrier ----------------
  const HFrag hf = make_hfrag(lane);
  const float sc = a.post_scale;
  const int g = lane >> 2, tq = lane & 3;
  const int grp = warp / RP_GW, wg = warp % RP_GW;       // every block b of this warp has (b & 7) == wg
  uint32_t af[MT][MT][4];                                // coefficient matrix as mma A fragments
#pragma unroll
  for (int mt = 0; mt < MT; mt++)
#pragma unroll
    for (int kt = 0; kt < MT; kt++)
      ldmatrix_x4(af[mt][kt], hk + (size_t)(mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * Kp + kt * 16 + (lane >> 4) * 8);
  const int swz_chunk = (lane ^ wg) << 3;                // this lane's chunk inside its (swizzled) blocks

  for (int it = grp; it < n_my; it += RP_GROUPS) {
    const int s = it % RP_SLOTS;
    __half* T = buf + (size_t)s * row_halfs;
    mbar_wait(bar_full + 8 * s, (uint32_t)(it / RP_SLOTS) & 1u);
    // P1: two blocks per pass for instruction-level parallelism
    for (int b0 = wg; b0 < K; b0 += 2 * RP_GW) {
      uint4 oct[2];
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int b = b0 + j * RP_GW, o = b * 32 + lane;
        oct[j] = make_uint4(0, 0, 0, 0);
        if (b < K && o < noct_in) oct[j] = *reinterpret_cast<const uint4*>(T + o * 8);
        if (a.pre && b < K) oct[j] = hmul2x4(oct[j], *reinterpret_cast<const uint4*>(s_pre + o * 8));   // qlinear.py:91
      }
      float r[2][8];
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const uint32_t p[4] = {oct[j].x, oct[j].z, oct[j].y, oct[j].w};
        fwht256_frag(p, hf, r[j]);                                                      // x 1/16; every lane has loaded
      }
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int b = b0 + j * RP_GW;
        if (b < K) *reinterpret_cast<uint4*>(T + b * 256 + swz_chunk) = frag_to_octet(r[j], sc);   // register_lib.py:20
      }
    }
    rp_bar(grp);
    // P2: T <- M T, 8-column tile nt = chunk nt of every block (quant.py:83: fp16 operands, fp32 accumulate, fp16 out)
#pragma unroll 2
    for (int ti = 0; ti < 32 / RP_GW; ti++) {
      const int nt = wg + ti * RP_GW;
      uint32_t bf[MT][2];
#pragma unroll
      for (int kt = 0; kt < MT; kt++) {
        const int r = kt * 16 + (lane & 15);
        const __half* src = (r < K) ? T + r * 256 + ((nt ^ (r & 7)) << 3) : zrow;
        ldmatrix_x2_trans(bf[kt], src);
      }
      float acc[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; mt++) {
        mma_16816_zero(acc[mt], af[mt][0], bf[0]);
#pragma unroll
        for (int kt = 1; kt < MT; kt++) mma_16816(acc[mt], af[mt][kt], bf[kt]);
      }
      __syncwarp();                                        // every lane has read its B fragments of this tile
#pragma unroll
      for (int mt = 0; mt < MT; mt++) {
        const int r0 = mt * 16 + g, r1 = r0 + 8;
        const int off = ((nt ^ g) << 3) + tq * 2;          // (r0 & 7) == (r1 & 7) == g
        if (r0 < K) *reinterpret_cast<__half2*>(T + r0 * 256 + off) = __floats2half2_rn(acc[mt][0], acc[mt][1]);
        if (r1 < K) *reinterpret_cast<__half2*>(T + r1 * 256 + off) = __floats2half2_rn(acc[mt][2], acc[mt][3]);
      }
    }
    rp_bar(grp);
    // P3
    for (int b0 = wg; b0 < K; b0 += 2 * RP_GW) {
      uint4 v[2];
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int b = b0 + j * RP_GW, o = b * 32 + lane;
        if (b < K) {
          v[j] = *reinterpret_cast<const uint4*>(T + b * 256 + swz_chunk);
          if (a.post) v[j] = hmul2x4(v[j], *reinterpret_cast<const uint4*>(s_post + o * 8));      // qlinear.py:112
          if (a.bias) v[j] = hadd2x4(v[j], *reinterpret_cast<const uint4*>(s_bias + o * 8));      // qlinear.py:114
        }
      }
      __syncwarp();                                        // both blocks have been read by every lane
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int b = b0 + j * RP_GW;
        if (b < K) *reinterpret_cast<uint4*>(T + (b * 32 + lane) * 8) = v[j];
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_done + 8 * s);
  }
}

}  // namespace qb

using namespace qb;

extern "C" int quipb200_rotate_batched(const void* x, int64_t ldx, void* y, int64_t ldy, const void* pre, const void* post,
                                       const void* bias, const void* hk_padded, int M, int in_features, int out_features,
                                       int n, int K, float scale, void* stream) {
  if (!x || !y || M < 0 || K < 1 || n < 1 || n % K) return QUIPB200_EINVAL;
  if (M == 0) return 0;
  if (in_features > n || out_features > n || (in_features & 7) || (out_features & 7) || (ldx & 7) || (ldy & 7))
    return QUIPB200_EUNSUPPORTED;
  if (!aligned16(x) || !aligned16(y) || !aligned16(pre) || !aligned16(post) || !aligned16(bias) || !aligned16(hk_padded))
    return QUIPB200_EALIGN;
  const int L = n / K;
  RotBArgs a{};
  a.x = (const __half*)x; a.ldx = ldx; a.y = (__half*)y; a.ldy = ldy;
  a.pre = (const __half*)pre; a.post = (const __half*)post; a.bias = (const __half*)bias; a.hk = (const __half*)hk_padded;
  a.M = M; a.in_feat = in_features; a.out_feat = out_features; a.n = n; a.K = K;
  a.post_scale = scale * sqrtf((float)L);
  const int sms = quipb200_sm_count();
  if (sms < 1) return (int)cudaErrorNoDevice;
  cudaStream_t st = (cudaStream_t)stream;
  if (K == 1 && n == 4096) {
    const size_t smem = 16 * DS_XROW * sizeof(float) + 4096 * 2;
    cudaError_t e = cudaFuncSetAttribute(rot4096_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    if (M >= qb::g_opt_rot_warp_rows) {      // enough rows to give every resident warp one
      const size_t smem_w = 3 * 512 * sizeof(uint4);
      const int ctas = (M + RW_WARPS - 1) / RW_WARPS;
      const int grid = ctas < sms * 3 ? ctas : sms * 3;
      rot4096w_kernel<<<grid, RW_THREADS, smem_w, st>>>(a);
    } else {
      const int grid = M < sms * 4 ? M : sms * 4;
      rot4096_kernel<<<grid, RB_THREADS, smem, st>>>(a);
    }
  } else if (L == 256 && K > 1 && K <= 48 && hk_padded && M >= qb::g_opt_rot_pipe_rows) {
    // persistent CTA per SM, rows streamed by bulk copies
    const int Kp = (K + 15) / 16 * 16;
    const int vecs = (pre ? 1 : 0) + (post ? 1 : 0) + (bias ? 1 : 0);
    const size_t smem = ((size_t)RP_SLOTS + vecs) * K * 512 + 512 + (size_t)Kp * Kp * 2 + 2 * RP_SLOTS * 8;
    const int grid = M < sms ? M : sms;
    cudaError_t e = cudaSuccess;
    switch (Kp >> 4) {
      case 1:
        e = cudaFuncSetAttribute(rotblk_pipe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) rotblk_pipe_kernel<1><<<grid, RP_THREADS, smem, st>>>(a);
        break;
      case 2:
        e = cudaFuncSetAttribute(rotblk_pipe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) rotblk_pipe_kernel<2><<<grid, RP_THREADS, smem, st>>>(a);
        break;
      default:
        e = cudaFuncSetAttribute(rotblk_pipe_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) rotblk_pipe_kernel<3><<<grid, RP_THREADS, smem, st>>>(a);
        break;
    }
    if (e != cudaSuccess) return (int)e;
  } else if (L == 256 && K <= 64 && (K == 1 || hk_padded)) {
    const int Kp = (K + 15) / 16 * 16;
    const size_t smem = (size_t)(K + 1) * RB_LS * 2 + (size_t)Kp * Kp * 2 + 64 * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(rotblk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = (int)((200 * 1024) / (smem + 1024));
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    const int grid = M < sms * per_sm ? M : sms * per_sm;
    rotblk_kernel<<<grid, RB_THREADS, smem, st>>>(a);
  } else {
    return QUIPB200_EUNSUPPORTED;
  }
  QB_LAUNCH_CHECK();
  return 0;
}
