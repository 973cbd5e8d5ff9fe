"""`torch.ops.quip_lib.*` -- the reference's operator registry (register_lib.py), re-implemented over
the C-ABI library libquipb200.so.  Same library name, op names and schemas as the reference
(register_lib.py:8-185), each with a fake (shape) impl so the ops trace / capture, and a CUDA-only
impl: CPU tensors raise NotImplementedError exactly as the reference's stubs do
(register_lib.py:12, 24, ...).  The binding is "torch.ops binding only": tensors are unwrapped to raw
pointers + the current stream and handed to `extern "C"` entry points.

New op (not in the reference): quip_lib::quantlinear_fwd -- the fused eval-mode QuantLinear.forward.
"""
import ctypes
import math

import torch
from torch import Tensor

from . import _native
from ._native import CODEBOOK_ENUM, LinearDesc, check, lib

_LIB = torch.library.Library("quip_lib", "DEF")
_DT = {torch.float16: 0, torch.bfloat16: 1, torch.float32: 2}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _contig(t):
    return t if t.is_contiguous() else t.contiguous()


def _define(schema, cuda_impl, fake_impl):
    name = schema.split("(")[0]
    _LIB.define(schema)
    _LIB.impl(name, cuda_impl, "CUDA")
    torch.library.register_fake(f"quip_lib::{name}", fake_impl, lib=_LIB)


# ------------------------------------------------------------------------------------------------
# hadamard                                                                 register_lib.py:10-20
# ------------------------------------------------------------------------------------------------
def _hadamard_cuda(x: Tensor, scale: float) -> Tensor:
    n = x.shape[-1]
    if n & (n - 1) or x.dtype not in _DT:
        raise RuntimeError(f"quip_lib::hadamard: last dim must be a power of two and dtype fp16/bf16/fp32, "
                           f"got {tuple(x.shape)} {x.dtype}")
    xc = _contig(x)
    y = torch.empty_like(xc)
    rows = xc.numel() // n if n else 0
    if rows == 0:
        return y
    with torch.cuda.device(x.device):
        check(lib().quipb200_hadamard(_ptr(xc), _ptr(y), rows, n, float(scale), _DT[x.dtype], _stream()),
              "hadamard")
    return y


_define("hadamard(Tensor x, float scale) -> Tensor", _hadamard_cuda, lambda x, scale: torch.empty_like(x))


# ------------------------------------------------------------------------------------------------
# decompress_*_origorder                                                  register_lib.py:109-185
# ------------------------------------------------------------------------------------------------
def _check_q(Qidxs, dtype, name):
    if Qidxs.dim() != 2 or Qidxs.dtype != dtype:
        raise RuntimeError(f"quip_lib::{name}: Qidxs must be 2-D {dtype}, got {tuple(Qidxs.shape)} {Qidxs.dtype}")


def _decompress_e8p(Qidxs: Tensor, grid: Tensor) -> Tensor:
    _check_q(Qidxs, torch.int16, "decompress_e8p_origorder")
    q = _contig(Qidxs)
    out = torch.empty((q.shape[0], q.shape[1] * 8), dtype=torch.float16, device=q.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(q.device):
        check(lib().quipb200_decompress_e8p(_ptr(q), _ptr(grid), _ptr(out), q.shape[0], q.shape[1], _stream()),
              "decompress_e8p_origorder")
    return out


def _decompress_e8prvq4(Qidxs: Tensor, grid: Tensor, scale: float) -> Tensor:
    _check_q(Qidxs, torch.int32, "decompress_e8prvq4_origorder")
    q = _contig(Qidxs)
    out = torch.empty((q.shape[0], q.shape[1] * 8), dtype=torch.float16, device=q.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(q.device):
        check(lib().quipb200_decompress_e8prvq4(_ptr(q), _ptr(grid), _ptr(out), q.shape[0], q.shape[1],
                                                float(scale), _stream()), "decompress_e8prvq4_origorder")
    return out


def _decompress_e8prvq3(Qidxs: Tensor, grid: Tensor, grid2: Tensor, scale: float) -> Tensor:
    _check_q(Qidxs, torch.int32, "decompress_e8prvq3_origorder")
    q = _contig(Qidxs)
    cols = q.shape[1] * 32 // 3
    out = torch.empty((q.shape[0], cols), dtype=torch.float16, device=q.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(q.device):
        check(lib().quipb200_decompress_e8prvq3(_ptr(q), _ptr(grid), _ptr(grid2), _ptr(out), q.shape[0],
                                                cols // 8, float(scale), _stream()),
              "decompress_e8prvq3_origorder")
    return out


def _decompress_d4(Qidxs: Tensor, grid: Tensor) -> Tensor:
    _check_q(Qidxs, torch.uint8, "decompress_d4_origorder")
    q = _contig(Qidxs)
    g = _d4_grid_f16(grid)
    out = torch.empty((q.shape[0], q.shape[1] * 4), dtype=torch.float16, device=q.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(q.device):
        check(lib().quipb200_decompress_d4(_ptr(q), _ptr(g), _ptr(out), q.shape[0], q.shape[1], _stream()),
              "decompress_d4_origorder")
    return out


def _decompress_hi(Qidxs: Tensor) -> Tensor:
    _check_q(Qidxs, torch.int32, "decompress_hi_origorder")
    q = _contig(Qidxs)
    out = torch.empty((q.shape[0], q.shape[1] * 8), dtype=torch.float16, device=q.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(q.device):
        check(lib().quipb200_decompress_hi(_ptr(q), _ptr(out), q.shape[0], q.shape[1], _stream()),
              "decompress_hi_origorder")
    return out


def _d4_grid_f16(grid):
    """The D4 kernel contract is fp16 [256,4] (origin_order.cu:796 reads it as uint64[256]); the
    reference never casts the per-layer buffer (SURVEY A.3 caution) -- do it here."""
    if grid.dtype != torch.float16:
        grid = grid.to(torch.float16)
    return _contig(grid)


def _fake_dec(mult):
    def f(Qidxs, *a):
        return Qidxs.new_empty((Qidxs.shape[0], Qidxs.shape[1] * mult), dtype=torch.float16)
    return f


_define("decompress_e8p_origorder(Tensor Qidxs, Tensor grid) -> Tensor", _decompress_e8p, _fake_dec(8))
_define("decompress_e8prvq4_origorder(Tensor Qidxs, Tensor grid, float scale) -> Tensor", _decompress_e8prvq4,
        _fake_dec(8))
_define("decompress_e8prvq3_origorder(Tensor Qidxs, Tensor grid, Tensor grid2, float scale) -> Tensor",
        _decompress_e8prvq3,
        lambda Q, g, g2, s: Q.new_empty((Q.shape[0], Q.shape[1] * 32 // 3), dtype=torch.float16))
_define("decompress_d4_origorder(Tensor Qidxs, Tensor grid) -> Tensor", _decompress_d4, _fake_dec(4))
_define("decompress_hi_origorder(Tensor Qidxs) -> Tensor", _decompress_hi, _fake_dec(8))


# ------------------------------------------------------------------------------------------------
# *_mm_origorder                                                           register_lib.py:22-107
# ------------------------------------------------------------------------------------------------
def _mm_fused(codebook, x, Qidxs, grid, scale, K):
    """Small-M integer-dp4a GEMV path; returns None if the shape is outside the fused path."""
    M, N = x.shape[0], Qidxs.shape[0]
    if M == 0:
        return x.new_empty((0, N))
    if M > _native.MM_MAX_M:
        return None
    L = lib()
    ws_bytes = L.quipb200_mm_workspace_bytes(M, N, K)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    out = torch.empty((M, N), dtype=torch.float16, device=x.device)
    with torch.cuda.device(x.device):
        rc = L.quipb200_mm(codebook, _ptr(x), _ptr(Qidxs), _ptr(grid), float(scale), _ptr(out), M, N, K,
                           _ptr(ws), ws_bytes, _stream())
    if rc == _native.EUNSUPPORTED:
        return None
    check(rc, "mm_origorder")
    return out




def umma_preferred(M: int, N: int, K: int, codebook=None) -> bool:
    """Dispatch policy of the codebook mm ops (measured on B200, profiles/README.md): the tcgen05 kernel decodes every code
    once for all rows, the integer-dp4a GEMV once per row (1 row: 10 us, 4 rows: 37 us, 16 rows: 121 us at 4096 x 4096),
    and decompress + cuBLAS catches up at ~128 rows for 4096 x 4096 and beyond 128 rows for the larger layers.  RVQ4B / D4: the reference's own small-M kernels (K2 / K3) stop at
    32 / 24 rows; the tcgen05 route covers 4 .. 32 rows for them."""
    opt = _native.get_option("umma")
    if opt == 0 or N % 128 or K % 128 or M < 1 or M > 256:
        return False
    if codebook in (_native.CB_E8P12RVQ3B, _native.CB_HI):
        return M <= 32           # no integer-GEMV route for these: the tcgen05 kernel serves the whole K4 / K5 range
    if codebook is not None and codebook != _native.CB_E8P12:
        return 4 <= M <= 32
    if opt == 1:
        return M > 16
    if 4 <= M <= 64:                 # (1 .. 3 rows: the integer GEMV, one pass per row)
        return True
    # beyond 64 rows the activation tile every CTA re-reads from L2 grows with M; decompress + cuBLAS pays 2 N K bytes of
    # dense weights instead, so the crossover moves out with the layer size (r02_umma_bench.json)
    return M <= 128 and N * K >= (32 << 20)      # (129 .. 256 rows: 0.7-0.9x of the dense route at every shape measured)


def _mm_umma(x, Qidxs, grid, K, codebook=None, scale=0.0, grid2=None):
    """1 <= M <= 256: in-kernel decode + tcgen05 GEMM (csrc/umma_gemm.cu); None if the shape is not covered."""
    M, N = x.shape[0], Qidxs.shape[0]
    if M < 1 or M > 256 or N % 128 or K % 128:
        return None
    if codebook is None:
        codebook = _native.CB_E8P12
    L = lib()
    # (no workspace: split-K partial tiles are reduced inside the kernel's cluster, so concurrent launches on different
    # streams share nothing)
    out = torch.empty((M, N), dtype=torch.float16, device=x.device)
    with torch.cuda.device(x.device):
        rc = L.quipb200_mm_umma(int(codebook), _ptr(x), _ptr(Qidxs), _ptr(grid), _ptr(grid2), float(scale), _ptr(out), M, N, K,
                                None, 0, _stream())
    if rc == _native.EUNSUPPORTED:
        return None
    check(rc, "mm_umma")
    return out


def _mm(codebook, name, x, Qidxs, grid, scale, K, dense, grid2=None):
    if x.dim() != 2 or Qidxs.dim() != 2:
        raise RuntimeError(f"quip_lib::{name}: x and Qidxs must be 2-D")
    if x.shape[1] != K:
        raise RuntimeError(f"quip_lib::{name}: x has {x.shape[1]} columns, Qidxs decodes to {K}")
    if x.device != Qidxs.device:
        raise RuntimeError(f"quip_lib::{name}: x and Qidxs on different devices")
    xh = _contig(x if x.dtype == torch.float16 else x.to(torch.float16))
    q = _contig(Qidxs)
    out = None
    M = xh.shape[0]
    if codebook is not None and umma_preferred(M, q.shape[0], K, codebook):
        out = _mm_umma(xh, q, grid, K, codebook, scale, grid2)   # tcgen05: decode once, all rows (the dp4a path re-decodes per row)
    if out is None and codebook in (_native.CB_E8P12, _native.CB_E8P12RVQ4B, _native.CB_D4):
        out = _mm_fused(codebook, xh, q, grid, scale, K)
    if out is None and codebook == _native.CB_E8P12 and _native.get_option("umma") == 1:
        out = _mm_umma(xh, q, grid, K)
    if out is None:
        # decompress + dense GEMM: what the reference itself does for M >= 32 (codebook/e8p12.py:153-155)
        out = xh @ dense(q).T
    return out if x.dtype == torch.float16 else out.to(x.dtype)


def _e8p_mm(x: Tensor, Qidxs: Tensor, grid: Tensor) -> Tensor:
    return _mm(_native.CB_E8P12, "e8p_mm_origorder", x, Qidxs, _contig(grid), 0.0, Qidxs.shape[1] * 8,
               lambda q: _decompress_e8p(q, grid))


def _e8prvq4_mm(x: Tensor, Qidxs: Tensor, grid: Tensor, scale: float) -> Tensor:
    return _mm(_native.CB_E8P12RVQ4B, "e8prvq4_mm_origorder", x, Qidxs, _contig(grid), scale,
               Qidxs.shape[1] * 8, lambda q: _decompress_e8prvq4(q, grid, scale))


def _d4_mm(x: Tensor, Qidxs: Tensor, grid: Tensor) -> Tensor:
    g = _d4_grid_f16(grid)
    return _mm(_native.CB_D4, "d4_mm_origorder", x, Qidxs, g, 0.0, Qidxs.shape[1] * 4,
               lambda q: _decompress_d4(q, g))


def _e8prvq3_mm(x: Tensor, Qidxs: Tensor, grid: Tensor, grid2: Tensor, scale: float) -> Tensor:
    return _mm(_native.CB_E8P12RVQ3B, "e8prvq3_mm_origorder", x, Qidxs, _contig(grid), scale, Qidxs.shape[1] * 32 // 3,
               lambda q: _decompress_e8prvq3(q, grid, grid2, scale), grid2=_contig(grid2))


def _hi_mm(x: Tensor, Qidxs: Tensor) -> Tensor:
    return _mm(_native.CB_HI, "hi_mm_origorder", x, Qidxs, None, 0.0, Qidxs.shape[1] * 8, _decompress_hi)


def _fake_mm(x, Qidxs, *a):
    return x.new_empty((x.shape[0], Qidxs.shape[0]), dtype=x.dtype)


_define("e8p_mm_origorder(Tensor x, Tensor Qidxs, Tensor grid) -> Tensor", _e8p_mm, _fake_mm)
_define("e8prvq4_mm_origorder(Tensor x, Tensor Qidxs, Tensor grid, float scale) -> Tensor", _e8prvq4_mm, _fake_mm)
_define("e8prvq3_mm_origorder(Tensor x, Tensor Qidxs, Tensor grid, Tensor grid2, float scale) -> Tensor",
        _e8prvq3_mm, _fake_mm)
_define("d4_mm_origorder(Tensor x, Tensor Qidxs, Tensor grid) -> Tensor", _d4_mm, _fake_mm)
_define("hi_mm_origorder(Tensor x, Tensor Qidxs) -> Tensor", _hi_mm, _fake_mm)


# ------------------------------------------------------------------------------------------------
# batched fused rotation (new)                          quant.py:72-88 + qlinear.py:91, :108-114
# ------------------------------------------------------------------------------------------------
def rotate_supported(n: int, K: int) -> bool:
    return (K == 1 and n == 4096) or (n == 256 * K and K <= 64)


def _rotate_fused(x: Tensor, pre, hk, post, bias, n: int, K: int, out_features: int, scale: float) -> Tensor:
    if x.dim() != 2 or x.dtype != torch.float16:
        raise RuntimeError("quip_lib::rotate_fused: x must be fp16 [M, in_features]")
    M, fin = x.shape
    y = torch.empty((M, out_features), dtype=torch.float16, device=x.device)
    if M == 0:
        return y
    if x.stride(1) != 1:
        x = x.contiguous()
    with torch.cuda.device(x.device):
        check(lib().quipb200_rotate_batched(_ptr(x), x.stride(0), _ptr(y), y.stride(0), _ptr(pre), _ptr(post), _ptr(bias),
                                            _ptr(hk), M, fin, out_features, n, K, float(scale), _stream()),
              "rotate_fused")
    return y


_define("rotate_fused(Tensor x, Tensor? pre, Tensor? hk, Tensor? post, Tensor? bias, int n, int K, int out_features, "
        "float scale) -> Tensor", _rotate_fused,
        lambda x, pre, hk, post, bias, n, K, out_features, scale: x.new_empty((x.shape[0], out_features)))


# ------------------------------------------------------------------------------------------------
# fused QuantLinear.forward (new)                                            qlinear.py:87-115
# ------------------------------------------------------------------------------------------------
FUSED_CODEBOOKS = ("E8P12", "E8P12RVQ4B", "D4")


def fused_supported(codebook_id: str, q_in: int, M: int) -> bool:
    if codebook_id not in FUSED_CODEBOOKS or M > _native.MM_MAX_M:
        return False
    return q_in % (32 if codebook_id == "E8P12RVQ4B" else 64) == 0   # packed row pitch multiple of 16 bytes


def _quantlinear_fwd(x: Tensor, Qidxs: Tensor, grid: Tensor, SU, SV, bias, had_left, had_right, wscale_pc,
                     codebook: int, in_features: int, out_features: int, q_in: int, q_out: int,
                     K_left: int, K_right: int, wscale: float, resid_scale: float) -> Tensor:
    if x.dim() != 2 or x.shape[1] != in_features or x.dtype != torch.float16:
        raise RuntimeError("quip_lib::quantlinear_fwd: x must be fp16 [M, in_features]")
    M = x.shape[0]
    y = torch.empty((M, out_features), dtype=torch.float16, device=x.device)
    if M == 0:
        return y
    if x.stride(1) != 1:
        x = x.contiguous()
    d = LinearDesc(codebook, in_features, out_features, q_in, q_out, K_left, K_right, wscale, resid_scale,
                   Qidxs.data_ptr(), grid.data_ptr(),
                   SU.data_ptr() if SU is not None else None, SV.data_ptr() if SV is not None else None,
                   bias.data_ptr() if bias is not None else None,
                   had_left.data_ptr() if had_left is not None else None,
                   had_right.data_ptr() if had_right is not None else None,
                   wscale_pc.data_ptr() if wscale_pc is not None else None)
    L = lib()
    ws_bytes = L.quipb200_linear_workspace_bytes(ctypes.byref(d), M)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        check(L.quipb200_linear_forward(ctypes.byref(d), _ptr(x), x.stride(0), _ptr(y), y.stride(0), M,
                                        _ptr(ws), ws_bytes, _stream()), "quantlinear_fwd")
    return y


_define("quantlinear_fwd(Tensor x, Tensor Qidxs, Tensor grid, Tensor? SU, Tensor? SV, Tensor? bias, "
        "Tensor? had_left, Tensor? had_right, Tensor? wscale_pc, int codebook, int in_features, "
        "int out_features, int q_in, int q_out, int K_left, int K_right, float wscale, float resid_scale) -> Tensor",
        _quantlinear_fwd,
        lambda x, Qidxs, grid, SU, SV, bias, hl, hr, wpc, cb, fin, fout, *a: x.new_empty((x.shape[0], fout)))
