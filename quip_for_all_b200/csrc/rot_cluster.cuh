// Block rotations with an orthogonal mix (11008 = 43 x 256, 28672 = 7 x 4096, 5120 = 5 x 1024, ...) for ONE
// activation row, spread over a thread-block cluster.
//
// A single CTA needs ~19 us for the 28672-point input/output side of a Llama-2-70B MLP linear (75 k warp
// instructions, every phase latency bound at 16 warps): more than the GEMV it brackets.  Here a cluster of
// C = min(8, K) CTAs shares the row:
//   step 1  CTA r owns the blocks k = r, r + C, ...: loads + element-wise pre-ops, FWHT of each 256-wide sub-block
//           in registers/shuffles, the remaining log2(L / 256) butterfly levels in one pass over local shared memory,
//           scale, round to fp16 -- and PUSHES every 8-column tile of the finished row k into the shared memory of the
//           CTA that owns that column range (st.shared::cluster through DSMEM)
//   sync    barrier.cluster: every CTA now holds all K rows of its own column slice
//   step 2  K x K mix of the slice on the tensor path (mma.sync m16n8k16, in place) -- columns are independent
//   step 3  input side: cluster-wide abs-max (one float per CTA exchanged through DSMEM), 16-bit fixed-point records;
//           output side: * SV + bias (+ residual) and the fp16 store
// Same arithmetic and rounding points as the one-CTA kernels in ql_device.cuh (fp32 butterflies, fp16 after the
// scaled FWHT, fp16 operands / fp32 accumulate in the mix).
#pragma once
#include <cooperative_groups.h>

#include "ql_device.cuh"

namespace qb {

namespace cg = cooperative_groups;

constexpr int ROTC_THREADS = 512;
constexpr int ROTC_MAX_CLUSTER = 8;

struct RotcPlan {
  int C;            // cluster size (grid.x)
  int tiles_base;   // 8-column tiles per CTA: rank r owns [r*base + min(r, rem), +base + (r < rem))
  int tiles_rem;
  int Lsd;          // row stride (halfs) of the slice array: 8 * (tiles_base + (rem > 0)) + 8
  int nblk_max;     // ceil(K / C)
};

struct PrologueArgs;
struct EpilogueArgs;
// up to QUIPB200_MAX_GROUP rows of the same shape in one launch (blockIdx.z): the members of a q/k/v or gate/up group
struct RotcProGroup { PrologueArgs a[QUIPB200_MAX_GROUP]; };
struct RotcEpiGroup { EpilogueArgs a[QUIPB200_MAX_GROUP]; };

static inline bool rotc_supported(int q, int K, int log2L) { return K > 1 && K <= 64 && log2L >= 8 && log2L <= 12 && q == (K << log2L); }

static inline RotcPlan rotc_plan(int K, int log2L) {
  RotcPlan p;
  p.C = K < ROTC_MAX_CLUSTER ? K : ROTC_MAX_CLUSTER;
  const int tiles = 1 << (log2L - 3);
  p.tiles_base = tiles / p.C;
  p.tiles_rem = tiles % p.C;
  p.Lsd = 8 * (p.tiles_base + (p.tiles_rem ? 1 : 0)) + 8;
  p.nblk_max = (K + p.C - 1) / p.C;
  return p;
}

static inline size_t rotc_smem_bytes(int K, int log2L, const RotcPlan& p) {
  size_t b = 0;
  if (log2L > 8) b += (size_t)p.nblk_max * sizeof(float) << log2L;    // fp32 rows of the local blocks
  b += ((size_t)(K + 1) * p.Lsd * sizeof(__half) + 15) / 16 * 16;      // slice (+ one zero row for the mma padding)
  b += (size_t)kpad(K) * kpad(K) * sizeof(__half);
  b += (64 + 16) * sizeof(float);
  return (b + 15) / 16 * 16;
}

struct RotcSmem {
  float* s;
  __half* t;
  __half* hk;
  float* red;
  float* mxslot;
};

__device__ __forceinline__ RotcSmem rotc_carve(unsigned char* base, int K, int log2L, const RotcPlan& p) {
  RotcSmem r;
  size_t off = 0;
  r.s = reinterpret_cast<float*>(base);
  if (log2L > 8) off += (size_t)p.nblk_max * sizeof(float) << log2L;
  r.t = reinterpret_cast<__half*>(base + off);
  off += ((size_t)(K + 1) * p.Lsd * sizeof(__half) + 15) / 16 * 16;
  r.hk = reinterpret_cast<__half*>(base + off);
  off += (size_t)kpad(K) * kpad(K) * sizeof(__half);
  r.red = reinterpret_cast<float*>(base + off);
  r.mxslot = r.red + 64;
  return r;
}

// owner CTA of column tile `tl` and the tile's index inside that CTA's slice
__device__ __forceinline__ void rotc_owner(const RotcPlan& p, int tl, int& d, int& lt) {
  const int big = p.tiles_rem * (p.tiles_base + 1);
  if (tl < big) {
    d = tl / (p.tiles_base + 1);
    lt = tl - d * (p.tiles_base + 1);
  } else {
    const int u = tl - big;
    const int dd = u / p.tiles_base;
    d = p.tiles_rem + dd;
    lt = u - dd * p.tiles_base;
  }
}
__device__ __forceinline__ int rotc_tile_begin(const RotcPlan& p, int r) { return r * p.tiles_base + min(r, p.tiles_rem); }
__device__ __forceinline__ int rotc_tile_count(const RotcPlan& p, int r) { return p.tiles_base + (r < p.tiles_rem ? 1 : 0); }

// remaining butterfly levels across the NO = L / 256 sub-blocks of the local blocks, then scale, round and push
template <int NO>
__device__ __forceinline__ void rotc_cross_push(cg::cluster_group& cluster, const RotcSmem& sm, const RotcPlan& p, int K, int r,
                                                float scale, int tid) {
  constexpr int L = NO * 256;
  const int nloc = (K - r + p.C - 1) / p.C;
  for (int it = tid; it < nloc * 128; it += ROTC_THREADS) {
    const int j = it >> 7, cp = (it & 127) * 2;
    const int k = r + j * p.C;
    float2 v[NO];
#pragma unroll
    for (int c = 0; c < NO; c++) v[c] = *reinterpret_cast<const float2*>(sm.s + j * L + c * 256 + cp);
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
    for (int c = 0; c < NO; c++) {
      const int col = c * 256 + cp;
      int d, lt;
      rotc_owner(p, col >> 3, d, lt);
      __half* dst = cluster.map_shared_rank(sm.t, d) + (size_t)k * p.Lsd + lt * 8 + (col & 7);
      *reinterpret_cast<__half2*>(dst) = __floats2half2_rn(v[c].x * scale, v[c].y * scale);
    }
  }
}

// after warp_fwht256 of sub-block c of local block j: either the finished row tile (L == 256) is pushed, or the fp32
// values are parked for rotc_cross_push
__device__ __forceinline__ void rotc_put(cg::cluster_group& cluster, const RotcSmem& sm, const RotcPlan& p, int log2L, int k, int j,
                                         int c, int lane, float (&f)[8], float scale) {
  if (log2L == 8) {
#pragma unroll
    for (int e = 0; e < 8; e++) f[e] *= scale;
    int d, lt;
    rotc_owner(p, lane, d, lt);
    __half* dst = cluster.map_shared_rank(sm.t, d) + (size_t)k * p.Lsd + lt * 8;
    *reinterpret_cast<uint4*>(dst) = pack_h8(f);
  } else {
    float4* dp = reinterpret_cast<float4*>(sm.s + ((size_t)j << log2L) + c * 256 + lane * 8);
    dp[0] = make_float4(f[0], f[1], f[2], f[3]);
    dp[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
}

__device__ __forceinline__ void rotc_finish_rows(cg::cluster_group& cluster, const RotcSmem& sm, const RotcPlan& p, int K,
                                                 int log2L, int r, float scale, int tid) {
  if (log2L > 8) {
    __syncthreads();
    switch (log2L) {
      case 9: rotc_cross_push<2>(cluster, sm, p, K, r, scale, tid); break;
      case 10: rotc_cross_push<4>(cluster, sm, p, K, r, scale, tid); break;
      case 11: rotc_cross_push<8>(cluster, sm, p, K, r, scale, tid); break;
      default: rotc_cross_push<16>(cluster, sm, p, K, r, scale, tid); break;
    }
  }
  cluster.sync();   // every CTA now holds all K rows of its column slice
}

__device__ __forceinline__ void rotc_mix(const RotcSmem& sm, const RotcPlan& p, int K, int ntiles, int tid) {
  RotSmem rs;
  rs.s = rs.s2 = nullptr; rs.pp = 0; rs.t = sm.t; rs.hk = sm.hk; rs.red = sm.red; rs.Ls = p.Lsd; rs.log2L = 0;
  const int warp = tid >> 5, lane = tid & 31, nwarps = ROTC_THREADS >> 5;
  switch (kpad(K) >> 4) {
    case 1: mix_mma_tiles<1>(rs, K, warp, lane, nwarps, ntiles); break;
    case 2: mix_mma_tiles<2>(rs, K, warp, lane, nwarps, ntiles); break;
    case 3: mix_mma_tiles<3>(rs, K, warp, lane, nwarps, ntiles); break;
    default: mix_mma_tiles<4>(rs, K, warp, lane, nwarps, ntiles); break;
  }
  __syncthreads();
}

__device__ __forceinline__ void rotc_setup(const RotcSmem& sm, const RotcPlan& p, const __half* hadK, int K, int transpose, int tid) {
  RotSmem rs;
  rs.hk = sm.hk;
  load_hadK(rs, hadK, K, transpose, tid, ROTC_THREADS);
  for (int i = tid; i < p.Lsd; i += ROTC_THREADS) sm.t[(size_t)K * p.Lsd + i] = __float2half_rn(0.f);
}

// ---------------------------------------------------------------------------------------------
// input side.  grid (C, M, members), cluster (C, 1, 1).  Host guarantees: rotc_supported, in_features % 8 == 0, x / gate / SU /
// norm_w rows 16-byte aligned.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROTC_THREADS) ql_prologue_cluster_kernel(const __grid_constant__ RotcProGroup grp, RotcPlan p) {
  const PrologueArgs& a = grp.a[blockIdx.z];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int r = (int)cluster.block_rank();
  const int m = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const RotcSmem sm = rotc_carve(smem_raw, a.K, a.log2L, p);
  pdl_launch_dependents();
  rotc_setup(sm, p, a.hadK, a.K, /*transpose=*/1, tid);
  pdl_wait();
  cluster.sync();   // peers are resident (DSMEM stores below) and their zero rows / coefficient loads are issued

  const __half* xr = a.x + (size_t)m * a.ldx;
  const __half* gr = a.gate ? a.gate + (size_t)m * a.ldgate : nullptr;
  const int noct_in = a.in_features >> 3;
  float rstd = 1.f;
  if (a.norm_w) {   // every CTA needs the row's rms: recomputed per CTA (57 KB of L2 reads at most)
    float ss = 0.f;
    for (int idx = tid; idx < noct_in; idx += ROTC_THREADS) {
      float f[8];
      const uint4 xv = *reinterpret_cast<const uint4*>(xr + (size_t)idx * 8);
      unpack_h8(xv, f);
      if (gr) pre_ops(f, *reinterpret_cast<const uint4*>(gr + (size_t)idx * 8), true, xv, false, 1.f, xv, false);
#pragma unroll
      for (int e = 0; e < 8; e++) ss = fmaf(f[e], f[e], ss);
    }
    ss = block_sum(ss, sm.red, tid, ROTC_THREADS);
    rstd = rsqrtf(ss / (float)a.in_features + a.norm_eps);
  }

  // step 1: one warp per 256-wide sub-block of the local blocks
  const int NO = 1 << (a.log2L - 8);
  const int nloc = (a.K - r + p.C - 1) / p.C;
  for (int u = warp; u < nloc * NO; u += ROTC_THREADS / 32) {
    const int j = u >> (a.log2L - 8), c = u & (NO - 1);
    const int k = r + j * p.C;
    const int idx = ((k << a.log2L) >> 3) + c * 32 + lane;
    const bool in = idx < noct_in;
    uint4 xv = make_uint4(0, 0, 0, 0), gv = xv, wv = xv, sv = xv;
    if (in) xv = *reinterpret_cast<const uint4*>(xr + (size_t)idx * 8);
    if (in && gr) gv = *reinterpret_cast<const uint4*>(gr + (size_t)idx * 8);
    if (in && a.norm_w) wv = *reinterpret_cast<const uint4*>(a.norm_w + (size_t)idx * 8);
    if (in && a.SU) sv = *reinterpret_cast<const uint4*>(a.SU + (size_t)idx * 8);
    float f[8];
    unpack_h8(xv, f);
    if (in) pre_ops(f, gv, gr != nullptr, wv, a.norm_w != nullptr, rstd, sv, a.SU != nullptr);
    warp_fwht256(f, lane);
    rotc_put(cluster, sm, p, a.log2L, k, j, c, lane, f, a.scale);
  }
  rotc_finish_rows(cluster, sm, p, a.K, a.log2L, r, a.scale, tid);

  // step 2
  const int ntiles = rotc_tile_count(p, r), tile0 = rotc_tile_begin(p, r);
  rotc_mix(sm, p, a.K, ntiles, tid);

  // step 3: abs-max over the whole row -> 16-bit fixed-point records of the own slice
  float mx = 0.f;
  for (int it = tid; it < a.K * ntiles; it += ROTC_THREADS) {
    const int k = it / ntiles, tl = it - k * ntiles;
    float f[8];
    unpack_h8(*reinterpret_cast<const uint4*>(sm.t + (size_t)k * p.Lsd + tl * 8), f);
#pragma unroll
    for (int e = 0; e < 8; e++) mx = fmaxf(mx, fabsf(f[e]));
  }
  mx = block_max(mx, sm.red, tid, ROTC_THREADS);
  if (tid < p.C) cluster.map_shared_rank(sm.mxslot, tid)[r] = mx;
  cluster.sync();
  mx = 0.f;
  for (int i = 0; i < p.C; i++) mx = fmaxf(mx, sm.mxslot[i]);
  const float inv = (mx > 0.f) ? 32767.0f / mx : 0.f;
  uint4* dst = a.xq + (size_t)m * (a.q_in >> 3);
  for (int it = tid; it < a.K * ntiles; it += ROTC_THREADS) {
    const int k = it / ntiles, tl = it - k * ntiles;
    float f[8];
    unpack_h8(*reinterpret_cast<const uint4*>(sm.t + (size_t)k * p.Lsd + tl * 8), f);
    uint4 rec;
    pack_record(f, inv, rec);
    dst[((k << a.log2L) >> 3) + tile0 + tl] = rec;
  }
  if (r == 0 && tid == 0) a.xscale[m] = (mx > 0.f) ? mx / 32767.0f : 0.f;
}

// ---------------------------------------------------------------------------------------------
// output side.  Host guarantees: rotc_supported, out_features % 8 == 0, acc / acc2 / wscale_pc / SV / bias / residual /
// y rows 16-byte aligned.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROTC_THREADS) ql_epilogue_cluster_kernel(const __grid_constant__ RotcEpiGroup grp, RotcPlan p) {
  const EpilogueArgs& a = grp.a[blockIdx.z];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int r = (int)cluster.block_rank();
  const int m = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const RotcSmem sm = rotc_carve(smem_raw, a.K, a.log2L, p);
  pdl_launch_dependents();
  rotc_setup(sm, p, a.hadK, a.K, /*transpose=*/0, tid);
  pdl_wait();
  cluster.sync();

  const float xs = a.xscale[m] * a.unit;
  const float* ar = a.acc + (size_t)m * a.q_out;
  const float* ar2 = a.acc2 ? a.acc2 + (size_t)m * a.q_out : nullptr;
  const int NO = 1 << (a.log2L - 8);
  const int nloc = (a.K - r + p.C - 1) / p.C;
  for (int u = warp; u < nloc * NO; u += ROTC_THREADS / 32) {
    const int j = u >> (a.log2L - 8), c = u & (NO - 1);
    const int k = r + j * p.C;
    const int idx = ((k << a.log2L) >> 3) + c * 32 + lane;
    const float4 v0 = __ldcg(reinterpret_cast<const float4*>(ar + (size_t)idx * 8));
    const float4 v1 = __ldcg(reinterpret_cast<const float4*>(ar + (size_t)idx * 8) + 1);
    float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    if (ar2) {
      const float4 w0 = __ldcg(reinterpret_cast<const float4*>(ar2 + (size_t)idx * 8));
      const float4 w1 = __ldcg(reinterpret_cast<const float4*>(ar2 + (size_t)idx * 8) + 1);
      const float rr[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int e = 0; e < 8; e++) f[e] = fmaf(a.resid_scale, rr[e], f[e]);
    }
#pragma unroll
    for (int e = 0; e < 8; e++) f[e] = f16_round(f[e] * xs);                          // origin_order.cu:129
    if (a.wscale_pc) {
      float w[8];
      unpack_h8(*reinterpret_cast<const uint4*>(a.wscale_pc + (size_t)idx * 8), w);
#pragma unroll
      for (int e = 0; e < 8; e++) f[e] = f16_round(f[e] * w[e]);                       // qlinear.py:107
    }
    warp_fwht256(f, lane);
    rotc_put(cluster, sm, p, a.log2L, k, j, c, lane, f, a.scale);
  }
  rotc_finish_rows(cluster, sm, p, a.K, a.log2L, r, a.scale, tid);

  const int ntiles = rotc_tile_count(p, r), tile0 = rotc_tile_begin(p, r);
  rotc_mix(sm, p, a.K, ntiles, tid);

  __half* yr = a.y + (size_t)m * a.ldy;
  const __half* rr = a.residual ? a.residual + (size_t)m * a.ldres : nullptr;
  const int noct_out = a.out_features >> 3;
  for (int it = tid; it < a.K * ntiles; it += ROTC_THREADS) {
    const int k = it / ntiles, tl = it - k * ntiles;
    const int idx = ((k << a.log2L) >> 3) + tile0 + tl;
    if (idx >= noct_out) continue;
    float f[8], o8[8];
    unpack_h8(*reinterpret_cast<const uint4*>(sm.t + (size_t)k * p.Lsd + tl * 8), f);
    if (a.SV) {      // qlinear.py:112
      unpack_h8(*reinterpret_cast<const uint4*>(a.SV + (size_t)idx * 8), o8);
#pragma unroll
      for (int e = 0; e < 8; e++) f[e] = f16_round(f[e] * o8[e]);
    }
    if (a.bias) {    // qlinear.py:114
      unpack_h8(*reinterpret_cast<const uint4*>(a.bias + (size_t)idx * 8), o8);
#pragma unroll
      for (int e = 0; e < 8; e++) f[e] = f16_round(f[e] + o8[e]);
    }
    if (rr) {        // decoder-layer residual (fusion hook)
      unpack_h8(__ldcg(reinterpret_cast<const uint4*>(rr + (size_t)idx * 8)), o8);
#pragma unroll
      for (int e = 0; e < 8; e++) f[e] += o8[e];
    }
    *reinterpret_cast<uint4*>(yr + (size_t)idx * 8) = pack_h8(f);
  }
}

// launch with the cluster dimension (and programmatic dependent launch, like every other kernel of the chain)
static inline cudaError_t launch_cluster_kernel(const void* fn, dim3 grid, int cluster_x, void** args, size_t smem,
                                                cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(ROTC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_opt_pdl ? 2 : 1;
  return cudaLaunchKernelExC(&cfg, fn, args);
}

}  // namespace qb
