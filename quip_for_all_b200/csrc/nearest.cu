// Quantise-time nearest-codeword search for the E8P12 family.
//
// Reference: `E8P12_codebook.round` (codebook/e8p12.py:125-128) -- `(2 * X @ grid.T - grid_norm).argmax(-1)` over
// the 65 536 x 8 codeword table -- and `E8P12RVQ4B_codebook.quantize` (codebook/e8p12_rvq4.py:37-46), which runs it
// twice (second time on `(X - init_vals) / opt_resid_scale`).  The reference materialises the [m, 65536] fp32 score
// matrix with a library GEMM (1 GB for the 4096 rows of one LDLQ step) and runs argmax over it.  Here the scores are
// never stored: a CTA decodes a chunk of 1024 codewords into shared memory once (same decode as the inference
// kernels, common.cuh `e8p_decode_q`), every thread keeps 4 input vectors in registers, streams the chunk through
// broadcast shared-memory reads (8 FMA + compare per pair) and folds its running (score, index) into one 64-bit
// atomicMax per vector.  Key = (order-preserving image of the fp32 score) << 32 | (0xffff - index): the maximum is
// the best score and, on equal scores, the LOWEST index -- torch.argmax's first-occurrence rule.
//
// Two kernels compute the argmax: `e8p_nearest_kernel` (every dot product, below) and `e8p_nearest_struct_kernel` (512
// structured candidates, further down; the default for the E8P12 stages: 10x faster again at 28 672 rows).
//
// Arithmetic of the brute-force kernel (stated so that the CPU oracle can restate it): score(c) = fl(sum_j (2 x_j) g_j) - fl(fl(sqrt(|g|^2))^2),
// the sum as an fp32 fma chain over j = 0..7 in element order; |g|^2 is exact in fp32 (multiples of 1/16 below 17), so
// the norm term equals torch's `grid.norm(dim=-1) ** 2` bit for bit.
#include <cuda_fp16.h>

#include "common.cuh"

namespace qb {

constexpr int NQ_THREADS = 128;
constexpr int NQ_VT = 4;        // vectors per thread
constexpr int NQ_CHUNK = 1024;  // codewords per CTA
constexpr int NQ_NCHUNK = 65536 / NQ_CHUNK;

__device__ __forceinline__ unsigned long long nq_key(float s, uint32_t c) {
  uint32_t u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned long long)(0xffffu - c);
}

// codeword c -> 8 floats in element order (weight i = packed byte {0,2,1,3,4,6,5,7}[i])
__device__ __forceinline__ void nq_decode(const uint2* __restrict__ tab, uint32_t c, float (&e)[8]) {
  uint2 t = __ldg(tab + (c >> 8));
  t.x |= 0x01010101u;
  t.y |= 0x01010101u;
  const uint2 q = e8p_decode_q(t, c & 0xffffu);
  float w[8];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    w[j] = (float)(int)(signed char)((q.x >> (8 * j)) & 0xffu) * 0.25f;
    w[4 + j] = (float)(int)(signed char)((q.y >> (8 * j)) & 0xffu) * 0.25f;
  }
  e[0] = w[0]; e[1] = w[2]; e[2] = w[1]; e[3] = w[3];
  e[4] = w[4]; e[5] = w[6]; e[6] = w[5]; e[7] = w[7];
}

// TABLE = false: the 65 536 E8P12 codewords, decoded from the abs table (tab), 64 chunks in grid.y.
// TABLE = true : an explicit codebook `table` = float [ncode][8] (ncode <= NQ_CHUNK: one chunk), e.g. the 256-entry e81b
//                residual grid of E8P12RVQ3B (codebook/e8p12_rvq3.py:16-50).
template <bool TABLE>
__global__ void __launch_bounds__(NQ_THREADS) e8p_nearest_kernel(const float* __restrict__ x, int64_t m,
                                                                 const uint2* __restrict__ tab,
                                                                 const float* __restrict__ table, int ncode,
                                                                 unsigned long long* __restrict__ keys) {
  __shared__ __align__(16) float g[NQ_CHUNK][8];
  __shared__ float gn[NQ_CHUNK];
  const int tid = threadIdx.x;
  const uint32_t c0 = blockIdx.y * NQ_CHUNK;
  const int nk = TABLE ? ncode : NQ_CHUNK;
  for (int k = tid; k < nk; k += NQ_THREADS) {
    float e[8];
    if (TABLE) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(table + (size_t)k * 8));
      const float4 b = __ldg(reinterpret_cast<const float4*>(table + (size_t)k * 8) + 1);
      e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w; e[4] = b.x; e[5] = b.y; e[6] = b.z; e[7] = b.w;
    } else {
      nq_decode(tab, c0 + k, e);
    }
    float n2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) n2 = fmaf(e[i], e[i], n2);
    const float nr = __fsqrt_rn(n2);
    gn[k] = __fmul_rn(nr, nr);
    *reinterpret_cast<float4*>(&g[k][0]) = make_float4(e[0], e[1], e[2], e[3]);
    *reinterpret_cast<float4*>(&g[k][4]) = make_float4(e[4], e[5], e[6], e[7]);
  }
  // this thread's vectors: v = (blockIdx.x * NQ_VT + i) * NQ_THREADS + tid  (consecutive lanes -> consecutive 32-byte rows)
  float xv[NQ_VT][8];
  int64_t vi[NQ_VT];
#pragma unroll
  for (int i = 0; i < NQ_VT; i++) {
    vi[i] = ((int64_t)blockIdx.x * NQ_VT + i) * NQ_THREADS + tid;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (vi[i] < m) {
      a = __ldg(reinterpret_cast<const float4*>(x + vi[i] * 8));
      b = __ldg(reinterpret_cast<const float4*>(x + vi[i] * 8) + 1);
    }
    xv[i][0] = 2.f * a.x; xv[i][1] = 2.f * a.y; xv[i][2] = 2.f * a.z; xv[i][3] = 2.f * a.w;
    xv[i][4] = 2.f * b.x; xv[i][5] = 2.f * b.y; xv[i][6] = 2.f * b.z; xv[i][7] = 2.f * b.w;
  }
  __syncthreads();
  float best[NQ_VT];
  int bk[NQ_VT];
#pragma unroll
  for (int i = 0; i < NQ_VT; i++) { best[i] = -INFINITY; bk[i] = 0; }
#pragma unroll 4
  for (int k = 0; k < nk; k++) {
    const float4 a = *reinterpret_cast<const float4*>(&g[k][0]);
    const float4 b = *reinterpret_cast<const float4*>(&g[k][4]);
    const float nn = gn[k];
#pragma unroll
    for (int i = 0; i < NQ_VT; i++) {
      float s = __fmul_rn(xv[i][0], a.x);
      s = fmaf(xv[i][1], a.y, s);
      s = fmaf(xv[i][2], a.z, s);
      s = fmaf(xv[i][3], a.w, s);
      s = fmaf(xv[i][4], b.x, s);
      s = fmaf(xv[i][5], b.y, s);
      s = fmaf(xv[i][6], b.z, s);
      s = fmaf(xv[i][7], b.w, s);
      s = __fsub_rn(s, nn);
      if (s > best[i]) { best[i] = s; bk[i] = k; }
    }
  }
#pragma unroll
  for (int i = 0; i < NQ_VT; i++)
    if (vi[i] < m) atomicMax(keys + vi[i], nq_key(best[i], c0 + (uint32_t)bk[i]));
}

// mode 0: single stage          vals = g[c]                          idx = c
// mode 1: first of two stages   vals = g[c] (held for mode 2)        idx = c (held), xr = (x - g[c]) / resid_scale, key reset
// mode 2: second of two stages  vals = vals + g[c] * resid_scale     idx = (idx << 16) + c
// mode 3: second stage against an explicit table (E8P12RVQ3B): vals = vals + table[c] * resid_scale, idx = (idx << 8) + c
__global__ void e8p_nearest_finish_kernel(const float* __restrict__ x, int64_t m, const uint2* __restrict__ tab,
                                          unsigned long long* __restrict__ keys, float* __restrict__ vals,
                                          long long* __restrict__ idx, float* __restrict__ xr, float resid_scale, int mode,
                                          const float* __restrict__ table) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= m) return;
  const uint32_t c = 0xffffu - (uint32_t)(keys[v] & 0xffffull);
  float e[8];
  if (mode == 3) {
#pragma unroll
    for (int j = 0; j < 8; j++) e[j] = __ldg(table + (size_t)c * 8 + j);
  } else {
    nq_decode(tab, c, e);
  }
  float4* vo = reinterpret_cast<float4*>(vals + v * 8);
  if (mode == 0) {
    vo[0] = make_float4(e[0], e[1], e[2], e[3]);
    vo[1] = make_float4(e[4], e[5], e[6], e[7]);
    idx[v] = (long long)c;
  } else if (mode == 1) {
    vo[0] = make_float4(e[0], e[1], e[2], e[3]);
    vo[1] = make_float4(e[4], e[5], e[6], e[7]);
    idx[v] = (long long)c;
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + v * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(x + v * 8) + 1);
    float4* ro = reinterpret_cast<float4*>(xr + v * 8);
    ro[0] = make_float4(__fdiv_rn(__fsub_rn(a.x, e[0]), resid_scale), __fdiv_rn(__fsub_rn(a.y, e[1]), resid_scale),
                        __fdiv_rn(__fsub_rn(a.z, e[2]), resid_scale), __fdiv_rn(__fsub_rn(a.w, e[3]), resid_scale));
    ro[1] = make_float4(__fdiv_rn(__fsub_rn(b.x, e[4]), resid_scale), __fdiv_rn(__fsub_rn(b.y, e[5]), resid_scale),
                        __fdiv_rn(__fsub_rn(b.z, e[6]), resid_scale), __fdiv_rn(__fsub_rn(b.w, e[7]), resid_scale));
    keys[v] = 0ull;
  } else {
    const float4 a = vo[0], b = vo[1];
    vo[0] = make_float4(__fadd_rn(a.x, __fmul_rn(e[0], resid_scale)), __fadd_rn(a.y, __fmul_rn(e[1], resid_scale)),
                        __fadd_rn(a.z, __fmul_rn(e[2], resid_scale)), __fadd_rn(a.w, __fmul_rn(e[3], resid_scale)));
    vo[1] = make_float4(__fadd_rn(b.x, __fmul_rn(e[4], resid_scale)), __fadd_rn(b.y, __fmul_rn(e[5], resid_scale)),
                        __fadd_rn(b.z, __fmul_rn(e[6], resid_scale)), __fadd_rn(b.w, __fmul_rn(e[7], resid_scale)));
    idx[v] = (idx[v] << (mode == 3 ? 8 : 16)) + (long long)c;
  }
}


// ---------------------------------------------------------------------------------------------
// Structured search: the same argmax from the STRUCTURE of the codebook (codebook/e8p12.py:82-103).  A codeword is
// g = (sigma . t + delta) / 4 with t one of the 256 signed abs-table rows (quarter units), sigma an even-weight sign vector
// and delta = +-1 (odd parity of the sign byte <-> the -1/4 shift), so for a fixed (t, delta)
//     2 x.g - |g|^2 = sum_j sigma_j w_j + delta sum(x) / 2 - (|t|^2 + 8) / 16,      w_j = t_j (x_j / 2 - delta / 8),
// maximised by sigma_j = sign(w_j), flipping the smallest |w_j| when the number of negations is odd: 512 candidates of
// ~40 operations instead of 65 536 dot products (oracle: quip_oracle.e8p_nearest_structured, checked against the brute
// force).  8 threads share a vector (32 abs rows each, ascending), winners meet through shuffles; equal scores resolve
// to the lower abs row, and inside a row to the lower sign byte -- the lowest index, as torch.argmax.
// ---------------------------------------------------------------------------------------------
constexpr int NS_THREADS = 256;
constexpr int NS_TPV = 8;                       // threads per vector
constexpr int NS_VPB = NS_THREADS / NS_TPV;     // vectors per block

__device__ __forceinline__ float ns_eval(const float (&t)[8], const float (&u)[8], float lin, float t2, uint32_t& negmask,
                                         int& jmin, bool& odd) {
  float sum = 0.f, mn = INFINITY;
  negmask = 0u;
  jmin = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const float w = t[j] * u[j];
    const float aw = fabsf(w);
    if (w < 0.f) negmask |= 1u << j;
    sum += aw;
    if (aw < mn) { mn = aw; jmin = j; }
  }
  odd = (__popc(negmask) & 1) != 0;
  const float val = odd ? sum - 2.0f * mn : sum;
  return val + lin - (t2 + 8.0f) * 0.0625f;
}

// element-order negation mask (+ parity fix) -> index of the codeword
__device__ __forceinline__ uint32_t ns_code(int a, uint32_t negmask, int jmin, bool odd, bool plus) {
  if (odd) negmask ^= 1u << jmin;
  // element i <-> packed byte {0,2,1,3,4,6,5,7}[i] <-> sign bit 7 - byte
  const int bitpos[8] = {7, 5, 6, 4, 3, 1, 2, 0};
  uint32_t sgn = 0u;
#pragma unroll
  for (int i = 0; i < 8; i++) sgn |= ((negmask >> i) & 1u) << bitpos[i];
  if (!plus) sgn ^= 1u;
  return ((uint32_t)a << 8) | sgn;
}

__global__ void __launch_bounds__(NS_THREADS) e8p_nearest_struct_kernel(const float* __restrict__ x, int64_t m,
                                                                        const uint2* __restrict__ tab,
                                                                        unsigned long long* __restrict__ keys) {
  __shared__ __align__(16) float st[256][8];
  __shared__ float st2[256];
  const int tid = threadIdx.x;
  {
    const uint2 q = __ldg(tab + tid);
    float w[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      w[j] = (float)(int)(signed char)((q.x >> (8 * j)) & 0xffu);
      w[4 + j] = (float)(int)(signed char)((q.y >> (8 * j)) & 0xffu);
    }
    const float e[8] = {w[0], w[2], w[1], w[3], w[4], w[6], w[5], w[7]};
    float n2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) n2 = fmaf(e[i], e[i], n2);
    st2[tid] = n2;
    *reinterpret_cast<float4*>(&st[tid][0]) = make_float4(e[0], e[1], e[2], e[3]);
    *reinterpret_cast<float4*>(&st[tid][4]) = make_float4(e[4], e[5], e[6], e[7]);
  }
  __syncthreads();
  const int64_t v = (int64_t)blockIdx.x * NS_VPB + (tid >> 3);
  const int part = tid & 7;
  float xv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (v < m) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + v * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(x + v * 8) + 1);
    xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w; xv[4] = b.x; xv[5] = b.y; xv[6] = b.z; xv[7] = b.w;
  }
  float up[8], um[8], xsum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    up[j] = xv[j] * 0.5f - 0.125f;      // delta = +1
    um[j] = xv[j] * 0.5f + 0.125f;      // delta = -1
    xsum += xv[j];
  }
  const float linp = 0.5f * xsum, linm = -0.5f * xsum;
  float best = -INFINITY;
  int ba = 0;
  bool bplus = true;
  for (int r = 0; r < 32; r++) {
    const int a = part * 32 + r;
    float t[8];
    const float4 t0 = *reinterpret_cast<const float4*>(&st[a][0]);
    const float4 t1 = *reinterpret_cast<const float4*>(&st[a][4]);
    t[0] = t0.x; t[1] = t0.y; t[2] = t0.z; t[3] = t0.w; t[4] = t1.x; t[5] = t1.y; t[6] = t1.z; t[7] = t1.w;
    const float t2 = st2[a];
    uint32_t nm;
    int jm;
    bool od;
    const float sp = ns_eval(t, up, linp, t2, nm, jm, od);
    uint32_t nm2;
    int jm2;
    bool od2;
    const float sm = ns_eval(t, um, linm, t2, nm2, jm2, od2);
    float sc = sp;
    bool plus = true;
    if (sm > sp) {
      sc = sm;
      plus = false;
    } else if (sm == sp) {            // same abs row, equal scores: the lower sign byte wins
      if (ns_code(a, nm2, jm2, od2, false) < ns_code(a, nm, jm, od, true)) plus = false;
    }
    if (sc > best) { best = sc; ba = a; bplus = plus; }
  }
  // the 8 threads of a vector: higher score, then lower abs row
#pragma unroll
  for (int o = 1; o < NS_TPV; o <<= 1) {
    const float os = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, ba, o);
    const int op = __shfl_xor_sync(0xffffffffu, (int)bplus, o);
    if (os > best || (os == best && oa < ba)) { best = os; ba = oa; bplus = op != 0; }
  }
  if (part == 0 && v < m) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; j++) t[j] = st[ba][j];
    uint32_t nm;
    int jm;
    bool od;
    (void)ns_eval(t, bplus ? up : um, bplus ? linp : linm, st2[ba], nm, jm, od);
    keys[v] = nq_key(best, ns_code(ba, nm, jm, od, bplus));
  }
}

int g_opt_nearest_struct = 1;   // 1: structured 512-candidate search for the E8P12 stages; 0: all 65 536 dot products
                                // (option "nearest_struct"; the brute-force kernel stays as the cross-check)

}  // namespace qb

using namespace qb;

extern "C" size_t quipb200_e8p_quantize_workspace_bytes(int64_t m) {
  if (m < 1) return 0;
  return (((size_t)m * sizeof(unsigned long long) + 15) & ~(size_t)15) + (size_t)m * 8 * sizeof(float);
}

static int nearest_run(const float* x, int64_t m, const int64_t* grid_packed_abs, int n_stages, float resid_scale,
                       const float* table2, int ncode2, float* vals_out, int64_t* idx_out, void* workspace,
                       size_t workspace_bytes, void* stream) {
  if (!x || !grid_packed_abs || !vals_out || !idx_out || !workspace || m < 1 || (n_stages != 1 && n_stages != 2))
    return QUIPB200_EINVAL;
  if (n_stages == 2 && !(resid_scale != 0.f && resid_scale == resid_scale)) return QUIPB200_EINVAL;   // (the reference's
  // quantizer default of -1 reaches the codebook unchanged, quantizer.py:69,127: negative scales are legal)
  if (table2 && (n_stages != 2 || ncode2 < 1 || ncode2 > NQ_CHUNK || !aligned16(table2))) return QUIPB200_EINVAL;
  if (!aligned16(x) || !aligned16(vals_out) || !aligned16(workspace) || ((uintptr_t)idx_out & 7)) return QUIPB200_EALIGN;
  if (workspace_bytes < quipb200_e8p_quantize_workspace_bytes(m)) return QUIPB200_EWORKSPACE;
  const int64_t gx = (m + NQ_THREADS * NQ_VT - 1) / (NQ_THREADS * NQ_VT);
  if (gx > 0x7fffffffLL) return QUIPB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(workspace);
  float* xr = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) +
                                       (((size_t)m * sizeof(unsigned long long) + 15) & ~(size_t)15));   // 16-byte aligned rows
  const uint2* tab = reinterpret_cast<const uint2*>(grid_packed_abs);
  cudaError_t e = cudaMemsetAsync(keys, 0, (size_t)m * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return (int)e;
  const dim3 grid((unsigned)gx, NQ_NCHUNK);
  const int fb = 256;
  const unsigned fgrid = (unsigned)((m + fb - 1) / fb);
  long long* idx = reinterpret_cast<long long*>(idx_out);
  const unsigned sgrid = (unsigned)((m + NS_VPB - 1) / NS_VPB);
  if (g_opt_nearest_struct) e8p_nearest_struct_kernel<<<sgrid, NS_THREADS, 0, st>>>(x, m, tab, keys);
  else e8p_nearest_kernel<false><<<grid, NQ_THREADS, 0, st>>>(x, m, tab, nullptr, 0, keys);
  QB_LAUNCH_CHECK();
  e8p_nearest_finish_kernel<<<fgrid, fb, 0, st>>>(x, m, tab, keys, vals_out, idx, xr, resid_scale, n_stages == 1 ? 0 : 1, nullptr);
  QB_LAUNCH_CHECK();
  if (n_stages == 2 && !table2) {
    if (g_opt_nearest_struct) e8p_nearest_struct_kernel<<<sgrid, NS_THREADS, 0, st>>>(xr, m, tab, keys);
    else e8p_nearest_kernel<false><<<grid, NQ_THREADS, 0, st>>>(xr, m, tab, nullptr, 0, keys);
    QB_LAUNCH_CHECK();
    e8p_nearest_finish_kernel<<<fgrid, fb, 0, st>>>(xr, m, tab, keys, vals_out, idx, xr, resid_scale, 2, nullptr);
    QB_LAUNCH_CHECK();
  } else if (n_stages == 2) {
    e8p_nearest_kernel<true><<<dim3((unsigned)gx, 1), NQ_THREADS, 0, st>>>(xr, m, tab, table2, ncode2, keys);
    QB_LAUNCH_CHECK();
    e8p_nearest_finish_kernel<<<fgrid, fb, 0, st>>>(xr, m, tab, keys, vals_out, idx, xr, resid_scale, 3, table2);
    QB_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int quipb200_e8p_quantize(const float* x, int64_t m, const int64_t* grid_packed_abs, int n_stages,
                                     float resid_scale, float* vals_out, int64_t* idx_out, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  return nearest_run(x, m, grid_packed_abs, n_stages, resid_scale, nullptr, 0, vals_out, idx_out, workspace, workspace_bytes,
                     stream);
}

extern "C" int quipb200_e8prvq3_quantize(const float* x, int64_t m, const int64_t* grid_packed_abs, const float* e81b_grid,
                                         float resid_scale, float* vals_out, int64_t* idx_out, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  if (!e81b_grid) return QUIPB200_EINVAL;
  return nearest_run(x, m, grid_packed_abs, 2, resid_scale, e81b_grid, 256, vals_out, idx_out, workspace, workspace_bytes,
                     stream);
}
