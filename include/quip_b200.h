/*
 * quip_b200.h -- C ABI of libquipb200.so: the B200 (sm_100a) replacement for the native layer under
 * the reference's `torch.ops.quip_lib.*` operator registry (chu-tianxiang/QuIP-for-all @ 04754a4).
 *
 * Drop-in boundary.  The reference binds its native code through a pybind11 module `quiptools_cuda`
 * (quip_cuda/quiptools_wrapper.cpp:87-100) plus the third-party `fast_hadamard_transform_cuda`
 * (register_lib.py:5,18-20).  Every entry point below names the reference interface it replaces.
 * Conventions (all entry points):
 *   - plain pointers + sizes, no torch/ATen types; every buffer is CALLER-OWNED device memory,
 *     contiguous row-major, 16-byte aligned;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*), never synchronised, never
 *     allocates -> CUDA-graph capturable (the reference relies on this: origin_order.cu:608-609);
 *   - returns 0 on success, a positive cudaError_t value for CUDA failures, or a negative QUIPB200_E*
 *     code for argument errors.  Never throws, never exits (contrast e8p_gemv.cu:36-43 `exit()`).
 *   - `Qidxs` tensors are the reference's on-disk format unchanged: (q_out, q_in/(codesz*packsz))
 *     (qlinear.py:52-57); signed storage is reinterpreted as unsigned (origin_order.cu:268, :841).
 */
#ifndef QUIP_B200_H
#define QUIP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QUIPB200_ABI_VERSION 1

/* argument-error codes (negative); CUDA errors are returned as positive cudaError_t values */
#define QUIPB200_EINVAL     (-1)  /* bad shape / null pointer / unsupported combination */
#define QUIPB200_EALIGN     (-2)  /* pointer or row pitch not 16-byte aligned */
#define QUIPB200_EWORKSPACE (-3)  /* workspace too small (see quipb200_linear_workspace_bytes) */
#define QUIPB200_EUNSUPPORTED (-4) /* shape outside what the fused path covers; caller uses the dense path */

/* codebooks: codebook/__init__.py:7-13 (codebook_id) */
enum quipb200_codebook {
  QUIPB200_CB_E8P12 = 0,      /* codebook/e8p12.py      int16 codes, 8 weights / code, 2 bit */
  QUIPB200_CB_E8P12RVQ4B = 1, /* codebook/e8p12_rvq4.py int32 codes (hi16 main | lo16 residual), 4 bit */
  QUIPB200_CB_D4 = 2,         /* codebook/d4.py         uint8 codes, 4 weights / code, 2 bit */
  QUIPB200_CB_E8P12RVQ3B = 3, /* codebook/e8p12_rvq3.py 3-byte codes, 3 bit */
  QUIPB200_CB_HI = 4          /* codebook/hi.py         int32 = 8 nibbles, 4 bit scalar */
};

int quipb200_abi_version(void);
/* Human-readable text for a code returned by any entry point (static storage). */
const char* quipb200_strerror(int code);
/* Number of SMs of the current device (cached); <0 on error. */
int quipb200_sm_count(void);

/* ---------------------------------------------------------------------------------------------
 * quip_lib::hadamard(Tensor x, float scale) -> Tensor        register_lib.py:10-20
 *   replaces fast_hadamard_transform_cuda.fast_hadamard_transform: unnormalised Sylvester-order
 *   Walsh-Hadamard transform over the last dim (n = power of two, <= 32768), fp32 internal math,
 *   y = H_n x * scale, output dtype = input dtype.  dtype: 0 = fp16, 1 = bf16, 2 = fp32.
 * ------------------------------------------------------------------------------------------- */
int quipb200_hadamard(const void* x, void* y, int64_t rows, int n, float scale, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dense dequantisation: quip_lib::decompress_*_origorder       register_lib.py:109-185
 *   replaces quiptools_cuda.decompress_{e8p,e8prvq4,d4,e8prvq3,hi}_origorder
 *   (origin_order.cu:837-885, :956-1024, :794-833, :887-954, :1028-1074).  Bit-exact with those kernels.
 *   out: fp16 [rows, cols] with cols = 8*codes_per_row (4* for D4).  Unlike the reference (which
 *   silently requires rows*cols % 2048 == 0) any size is handled.
 * ------------------------------------------------------------------------------------------- */
int quipb200_decompress_e8p(const int16_t* qidxs, const int64_t* grid_packed_abs, void* out_f16,
                            int64_t rows, int64_t codes_per_row, void* stream);
int quipb200_decompress_e8prvq4(const int32_t* qidxs, const int64_t* grid_packed_abs, void* out_f16,
                                int64_t rows, int64_t codes_per_row, float resid_scale, void* stream);
int quipb200_decompress_d4(const uint8_t* qidxs, const void* grid_f16_256x4, void* out_f16,
                           int64_t rows, int64_t codes_per_row, void* stream);
/* qidxs: int32 [rows, 3*codes_per_row/4] viewed as byte triplets */
int quipb200_decompress_e8prvq3(const int32_t* qidxs, const int64_t* grid_packed_abs,
                                const int32_t* e81b_packed, void* out_f16,
                                int64_t rows, int64_t codes_per_row, float resid_scale, void* stream);
int quipb200_decompress_hi(const int32_t* qidxs, void* out_f16, int64_t rows, int64_t codes_per_row,
                           void* stream);

/* ---------------------------------------------------------------------------------------------
 * Decode + matmul: quip_lib::{e8p,e8prvq4,d4}_mm_origorder       register_lib.py:22-38, :58-92
 *   replaces quiptools_cuda.*_mm_origorder (origin_order.cu:557-743):
 *     out[M,N] (fp16) = x[M,K] (fp16) . decode(Qidxs[N, K/codesz])^T
 *   Small-M (decode) path: activations are quantised per row to 16-bit fixed point, the dot products
 *   run as exact integer dp4a on CUDA cores, one fp16 rounding at the end (tolerance: DESIGN.md).
 *   Requires the packed row pitch to be a multiple of 16 bytes (K % 64 == 0; RVQ4B: K % 32 == 0)
 *   and M <= QUIPB200_MM_MAX_M; otherwise returns QUIPB200_EUNSUPPORTED and the caller takes the
 *   decompress + dense-GEMM path (exactly what the reference does for M >= 32, e8p12.py:153-155).
 *   workspace: >= quipb200_mm_workspace_bytes(M, N, K) bytes of device scratch.
 * ------------------------------------------------------------------------------------------- */
#define QUIPB200_MM_MAX_M 16
size_t quipb200_mm_workspace_bytes(int M, int N, int K);
int quipb200_mm(int codebook, const void* x_f16, const void* qidxs, const void* grid,
                float resid_scale, void* out_f16, int M, int N, int K,
                void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Batched fused rotation (M >= 17 path of QuantLinear.forward): one pass over [M, n] instead of the reference's
 * x*SU -> F.pad -> hadamard -> hadK.T.contiguous() -> hadK @ .. (input side, qlinear.py:91,99-100, quant.py:72-88) or
 * hadamard -> hadK @ .. -> [:out_features] -> *SV -> +bias (output side, qlinear.py:108-114):
 *     y[M, out_features] = f16( f16( (M_K (x) H_{n/K}) f16(x * pre) * scale )[:out_features] * post ) + bias
 *   pre / post / bias: fp16 vectors or NULL; hk_padded: fp16 [Kp][Kp] (Kp = roundup16(K)), zero padded, holding the
 *   coefficient matrix M_K[k_out][k_in] (hadK for the output side, hadK^T for the input side) or NULL when K == 1;
 *   scale as passed to quip_lib::hadamard (quant.py:75: scale / sqrt(n/K)).
 *   Covered: n == 4096 with K == 1, and n == 256*K with K <= 64; else QUIPB200_EUNSUPPORTED (caller keeps the
 *   reference's op sequence over quipb200_hadamard).  in/out_features % 8 == 0, row pitches % 8 == 0.
 *   Kernel choice by M (options "rot_warp_rows", "rot_pipe_rows"): a CTA per row for few rows; for many rows one warp
 *   per row (n == 4096) or persistent CTAs streaming rows through shared memory with bulk copies (n == 256*K, K <= 48).
 *   All variants give the same result up to the fp32 summation order before the fp16 roundings.
 * ------------------------------------------------------------------------------------------- */
int quipb200_rotate_batched(const void* x_f16, int64_t ldx, void* y_f16, int64_t ldy, const void* pre, const void* post,
                            const void* bias, const void* hk_padded, int M, int in_features, int out_features,
                            int n, int K, float scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Batched decode + GEMM on the 5th-generation tensor cores (tcgen05.mma, TMEM accumulators), 1 <= M <= 256:
 *   same contract as quipb200_mm for E8P12, replacing the reference's M >= 32 route
 *   "decompress_e8p_origorder + input @ W.T" (codebook/e8p12.py:153-155, origin_order.cu:837-885) without
 *   ever materialising the dense weight: packed codes are decoded straight into the UMMA shared-memory
 *   layout.  Requires N % 128 == 0 and K % 64 == 0 (else QUIPB200_EUNSUPPORTED -> caller uses the dense path).
 *   The activation tile is staged by TMA (tensor map over x, encoded per call) and the split-K partial tiles are reduced
 *   through distributed shared memory inside a thread-block cluster, so the call needs NO workspace:
 *   quipb200_e8p_mm_umma_workspace_bytes returns 0 and workspace / workspace_bytes are ignored (kept for ABI stability).
 *   Concurrent launches on different streams share nothing.
 * ------------------------------------------------------------------------------------------- */
size_t quipb200_e8p_mm_umma_workspace_bytes(int M, int N, int K);
/* The same kernel with the producers' decode templated on the codebook, covering the reference's small-M mm kernels:
 *   QUIPB200_CB_E8P12       K1  origin_order.cu:388-555
 *   QUIPB200_CB_E8P12RVQ4B  K2  :337-385, :698-743   `scale` = residual scale, rounded to fp16, one fp16 fma (as the reference)
 *   QUIPB200_CB_D4          K3  :143-168, :557-602   grid = fp16 [256][4]
 *   QUIPB200_CB_E8P12RVQ3B  K4  :287-335, :650-696   grid2 = e81b residual table, int32[256]; `scale` as for RVQ4B
 *   QUIPB200_CB_HI          K5  :170-206, :745-788   no table (grid may be NULL)
 * grid2 is NULL except for RVQ3B.  Same shape rules as quipb200_e8p_mm_umma. */
int quipb200_mm_umma(int codebook, const void* x_f16, const void* qidxs, const void* grid, const void* grid2, float scale,
                     void* out_f16, int M, int N, int K, void* workspace, size_t workspace_bytes, void* stream);
int quipb200_e8p_mm_umma(const void* x_f16, const void* qidxs, const void* grid_packed_abs, void* out_f16,
                         int M, int N, int K, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused QuantLinear.forward (eval branch)                          qlinear.py:87-115
 *   y = [(hadK_R (x) H) ( decode(Qidxs) . ((hadK_L^T (x) H) (SU . x)) * wscale ) * Wscale_pc][:out] . SV + bias
 *   i.e. the whole chain  x*SU -> matmul_hadUt_cuda -> codebook(x, Qidxs) -> [*Wscale] ->
 *   matmul_hadU_cuda -> [:out_features] -> *SV -> +bias  (quant.py:72-88) in three launches
 *   (prologue / GEMV / epilogue) instead of the reference's 5-9 launches + library calls.
 * ------------------------------------------------------------------------------------------- */
typedef struct quipb200_linear {
  int32_t codebook;          /* enum quipb200_codebook (E8P12, E8P12RVQ4B, D4 on the fused path) */
  int32_t in_features;       /* qlinear.py:20 */
  int32_t out_features;      /* qlinear.py:21 */
  int32_t q_in;              /* padded in  dim (qlinear.py:29) */
  int32_t q_out;             /* padded out dim (qlinear.py:30) */
  int32_t K_left;            /* qlinear.py:29  (1 => pure FWHT) */
  int32_t K_right;           /* qlinear.py:30 */
  float wscale_float;        /* qlinear.py:75, quantizer.py:837 */
  float resid_scale;         /* codebook.opt_resid_scale (RVQ only) */
  const void* qidxs;         /* Qidxs buffer, reference layout (qlinear.py:52-57) */
  const void* grid;          /* E8P*: int64[256] grid_packed_abs; D4: fp16 [256,4] */
  const void* SU;            /* fp16 [in_features]  or NULL (qlinear.py:90, quantizer.py:840-844) */
  const void* SV;            /* fp16 [out_features] or NULL */
  const void* bias;          /* fp16 [out_features] or NULL */
  const void* had_left;      /* fp16 [K_left, K_left]   or NULL when K_left == 1 */
  const void* had_right;     /* fp16 [K_right, K_right] or NULL when K_right == 1 */
  const void* wscale_pc;     /* fp16 [q_out] per-channel Wscale (already mean-normalised) or NULL */
} quipb200_linear_t;

/* Fused-path launches take a "last CTA finishes the output side" ticket slot from a device-global table of 4096 slots,
 * handed out round-robin by the host and reset by the consuming CTA: launches that could land on the same slot must not
 * overlap in time, i.e. issue fused forwards of one device from one stream (or serialise streams with events).  A captured
 * 7B decode step uses 224 slots. */
size_t quipb200_linear_workspace_bytes(const quipb200_linear_t* layer, int M);
/* x: fp16 [M, in_features] with row pitch ldx elements; y: fp16 [M, out_features] pitch ldy. */
int quipb200_linear_forward(const quipb200_linear_t* layer, const void* x_f16, int64_t ldx,
                            void* y_f16, int64_t ldy, int M,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Grouped / hooked variant used by the decode engine (no reference counterpart: the reference leaves
 * these element-wise ops to HF modules and hides their launches behind torch.compile CUDA graphs,
 * example_generate.py:68-70).  Up to QUIPB200_MAX_GROUP QuantLinears that read the SAME input
 * (q/k/v or gate/up of a decoder layer) run in ONE launch, optionally with the ops that surround them
 * folded in:
 *   pre_norm_weight : x <- LlamaRMSNorm(x) (fp32 statistics, fp16 result times fp16 weight)   before SU
 *   gate            : x <- silu(gate) * x   (LlamaMLP: act_fn(gate_proj(h)) * up_proj(h))      before SU
 *   residual        : y <- residual + y     (decoder-layer skip connection)                    after bias
 * ------------------------------------------------------------------------------------------- */
#define QUIPB200_MAX_GROUP 3
typedef struct quipb200_fusion {
  const void* pre_norm_weight;   /* fp16 [in_features] or NULL */
  float pre_norm_eps;
  int32_t reserved;
  const void* gate;              /* fp16 [M, in_features] (row pitch ldgate) or NULL */
  int64_t ldgate;
  const void* residual;          /* fp16 [M, out_features] (row pitch ldres) or NULL; may alias y */
  int64_t ldres;
} quipb200_fusion_t;

size_t quipb200_linear_group_workspace_bytes(const quipb200_linear_t* layers, int n_layers, int M);
int quipb200_linear_group_forward(const quipb200_linear_t* layers, int n_layers,
                                  const quipb200_fusion_t* fusion /* may be NULL */,
                                  const void* x_f16, int64_t ldx, void* const* y_f16, const int64_t* ldy,
                                  int M, void* workspace, size_t workspace_bytes, void* stream);

/* RoPE (HF rotate_half convention) + KV-cache append + single-query attention, one CTA per head.
 * q [n_heads*head_dim], k/v [n_kv_heads*head_dim] fp16; caches [n_kv_heads, max_len, head_dim] fp16;
 * cos/sin [max_len, head_dim] fp16; *pos = index of the new token (device memory, graph-replayable). */
int quipb200_attn_decode(const void* q, const void* k, const void* v, void* k_cache, void* v_cache,
                         const void* cos_t, const void* sin_t, const int64_t* pos, void* out,
                         int n_heads, int n_kv_heads, int head_dim, int max_len, void* stream);

/* Tail of a bs=1 greedy decode step in one launch: LlamaRMSNorm(h) -> lm_head (fp16 [vocab, hidden], fp32 accumulate, fp16
 * logits) -> argmax (ties: lowest index, as torch.argmax) -> *tok_out; optionally h_next = emb[tok] and *pos += 1.
 * Replaces HF's final norm + nn.Linear + torch.argmax (+ nn.Embedding + `input_pos += 1`) of example_generate.py:29-56.
 * emb / h_next / pos / logits_out may be NULL.  workspace: quipb200_lm_tail_workspace_bytes() bytes, 256-byte aligned,
 * zero-filled ONCE by the caller (the arrival counter resets itself); one launch at a time per workspace. */
size_t quipb200_lm_tail_workspace_bytes(void);
int quipb200_lm_tail(const void* h_f16, const void* norm_w_f16, float eps, const void* lm_head_f16, const void* emb_f16,
                     int hidden, int vocab, int64_t* tok_out, void* h_next_f16, int64_t* pos, void* logits_out_f16,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Whole decode step of a Llama-style stack of QuantLinears in ONE persistent cooperative kernel
 * (bs = 1).  Replaces, per token, what the reference runs as 32 x (7 QuantLinear.forward + HF
 * attention / norm / MLP glue) hidden behind torch.compile CUDA graphs (example_generate.py:29-32,
 * :68-70; qlinear.py:87-115 per linear).  One CTA per SM stays resident for all layers; the five
 * data-dependent stages of a decoder layer ([norm + q,k,v] -> [RoPE + KV append + attention] ->
 * [o_proj + residual] -> [norm + gate,up] -> [silu*up + down_proj + residual]) are separated by
 * grid-wide barriers instead of kernel boundaries, every CTA recomputes the (tiny) rotations it needs
 * from the previous stage's raw integer dot products, and the packed codes of a stage are already in
 * flight while the rotation runs.  Arithmetic and rounding points are those of
 * quipb200_linear_group_forward / quipb200_attn_decode.
 *
 * Restrictions (else QUIPB200_EUNSUPPORTED and the caller uses the per-linear entry points): one codebook
 * (E8P12, E8P12RVQ4B or D4) for every linear, head_dim 128, fp16 everywhere, all seven linears of a layer present,
 * power-of-two padded attention dims (K_right == 1 for q/k/v, K_left == 1 for o).
 * ------------------------------------------------------------------------------------------- */
typedef struct quipb200_decode_layer {
  quipb200_linear_t q, k, v, o, gate, up, down;   /* device pointers inside */
  const void* input_norm_w;   /* fp16 [hidden]  LlamaRMSNorm before attention */
  const void* post_norm_w;    /* fp16 [hidden]  LlamaRMSNorm before the MLP */
  void* k_cache;              /* fp16 [n_kv_heads, max_len, head_dim] */
  void* v_cache;
  const void* mlp_hk;         /* optional (NULL ok): fp16 [3][Kp][Kp], Kp = roundup16(K_right of gate): gate.had_right,
                                 up.had_right and down.had_left^T, each zero padded -- lets the kernel fetch the three
                                 coefficient matrices with 16-byte requests */
} quipb200_decode_layer_t;

typedef struct quipb200_decode_plan {
  int32_t n_layers, hidden, n_heads, n_kv_heads, head_dim, max_len;
  float norm_eps;
  int32_t reserved;
  const quipb200_decode_layer_t* layers;   /* DEVICE array [n_layers] */
  const void* cos_t;                       /* fp16 [max_len, head_dim] */
  const void* sin_t;
  const int64_t* pos;                      /* device: index of the new token (graph-replayable) */
} quipb200_decode_plan_t;

/* `host_layers`: a HOST copy of the layer array (shapes are validated and scratch is sized from it).
 * The workspace must be zero-filled once before its first use and must not be shared between plans
 * whose launches can overlap. */
size_t quipb200_decode_step_workspace_bytes(const quipb200_decode_plan_t* plan,
                                            const quipb200_decode_layer_t* host_layers);
/* h_in: fp16 [hidden] (embedding of the current token); h_out: fp16 [hidden] (input of the final norm). */
int quipb200_decode_step(const quipb200_decode_plan_t* plan, const quipb200_decode_layer_t* host_layers,
                         const void* h_in, void* h_out, void* workspace, size_t workspace_bytes, void* stream);

/* profiling hook: when non-NULL, CTA 0 of the decode-step kernel writes int64 clock64() stamps of the
 * first layer's stages into buffer[0..15]; pass NULL to switch it off. */
int quipb200_decode_step_debug(void* device_int64_buffer);
int quipb200_decode_step_debug_cta(int cta);   /* which CTA writes the stamps (default 0) */
/* tuning / test hook: force the number of KV splits per head of the attention stage (0 = automatic, <= 4);
 * takes effect for workspaces sized and steps launched afterwards. */
int quipb200_decode_step_set_splits(int splits);

/* ---------------------------------------------------------------------------------------------
 * Quantise-time nearest-codeword search
 *   replaces `E8P12_codebook.round` / `.quantize` (codebook/e8p12.py:125-134: `(2 * X @ grid.T - grid_norm).argmax(-1)`
 *   over the 65 536 x 8 table, then `grid[idx]`) and, with n_stages = 2, `E8P12RVQ4B_codebook.quantize`
 *   (codebook/e8p12_rvq4.py:37-46: a second search on `(X - init_vals) / opt_resid_scale`,
 *   vals = init + resid * scale, idx = (init << 16) + resid) -- the call LDLQ makes once per 8 columns (quant.py:128-129).
 *   x: fp32 [m, 8]; vals_out: fp32 [m, 8]; idx_out: int64 [m] (torch.argmax's dtype; equal scores resolve to the lowest
 *   index, torch's first-occurrence rule).  The [m, 65536] score matrix of the reference is never materialised.
 *   workspace: quipb200_e8p_quantize_workspace_bytes(m) bytes, 16-byte aligned.
 * ------------------------------------------------------------------------------------------- */
size_t quipb200_e8p_quantize_workspace_bytes(int64_t m);
int quipb200_e8p_quantize(const float* x, int64_t m, const int64_t* grid_packed_abs, int n_stages, float resid_scale,
                          float* vals_out, int64_t* idx_out, void* workspace, size_t workspace_bytes, void* stream);
/* `E8P12RVQ3B_codebook.quantize` (codebook/e8p12_rvq3.py:91-101): E8P12 search, then the residual `(X - init) / scale`
 * against the 256-entry e81b grid (fp32 [256, 8], :16-50); vals = init + resid * scale, idx = (init << 8) + resid.
 * Same workspace size. */
int quipb200_e8prvq3_quantize(const float* x, int64_t m, const int64_t* grid_packed_abs, const float* e81b_grid,
                              float resid_scale, float* vals_out, int64_t* idx_out, void* workspace, size_t workspace_bytes,
                              void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU layer pipeline: stage-to-stage hand-off over NVLink peer memory (the reference's only
 * multi-device facility keeps whole decoder blocks per device, quantizer.py:180-191, :831, and moves the activation with
 * accelerate's hooks).  A mailbox is device memory of THIS process exported by CUDA IPC; the upstream process maps it and
 * stores the payload + a sequence number into it (send); the owner polls the sequence number locally and copies the
 * payload into its engine's input buffer (wait).  Both are enqueued on `stream`, involve no host synchronisation and
 * advance a caller-owned device counter (uint64, zero-initialised unless the first wait is to pass without a send: -1).
 * bytes: multiple of 8 (8-byte aligned buffers); a wait whose sequence number is 0 copies nothing.  A wait not satisfied within ~2 s of GPU time increments *err_flag (uint32) and returns.
 * ------------------------------------------------------------------------------------------- */
int quipb200_mailbox_create(size_t bytes, void** dev_ptr, void* ipc_handle_64);   /* zero-filled; 64-byte cudaIpcMemHandle_t out */
int quipb200_mailbox_open(const void* ipc_handle_64, void** peer_ptr);
int quipb200_mailbox_close(void* peer_ptr);
int quipb200_mailbox_destroy(void* dev_ptr);
int quipb200_handoff_send(const void* src, void* peer_dst, size_t bytes, void* peer_flag, void* seq_counter, void* stream);
int quipb200_handoff_wait(const void* flag, void* seq_counter, const void* inbox, void* dst, size_t bytes, void* err_flag,
                          void* stream);

/* Tuning / introspection hooks used by bench.py and the tests (not part of the reference surface). */
/* options: "umma" 0|1|2 (tcgen05 decode+GEMM never / whenever covered / where measured faster; default 2),
 *   "rot_warp_rows", "rot_pipe_rows" (row counts from which the many-rows rotation kernels are used),
 *   "pdl" 0|1, "fuse" 0..3, "stage_mask" 0..7 (bench only), "gemv_warps", "gemv_ctas_per_sm", "lean", "phase0"
 *   (tuning hooks of the per-linear launches).  Unknown names / out-of-range values: QUIPB200_EINVAL. */
int quipb200_set_option(const char* name, int value);
int quipb200_get_option(const char* name);
/* profiling hook: when non-NULL, every CTA of the GEMV kernel writes 16 int64 clock64() stamps of its
 * phases into buffer[cta*16 ..] (tools/timeline.py); pass NULL to switch it off. */
int quipb200_debug_timeline(void* device_int64_buffer);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches evidence) */
int64_t quipb200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* QUIP_B200_H */
