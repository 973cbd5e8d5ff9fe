"""Hadamard helpers of the QuantLinear path (reference: quant.py:1-88).

  y = (hadK (x) H_{n/K}) x / sqrt(n/K),   flat index = k*(n/K) + c,   H = Sylvester Walsh-Hadamard

`matmul_hadU` / `matmul_hadUt` are the device-agnostic pure-torch transforms (the reference's only
CPU-runnable ones, used at quantisation time); `matmul_hadU_cuda` / `matmul_hadUt_cuda` run the
`quip_lib::hadamard` CUDA op.  The LDLQ solvers of the reference's quant.py (:91-232) live in `ldlq.py`.
"""
import math
import os

import torch

_HAD_TABLES = None
_HAD_FILE = "hadamard.safetensors"


def _had_tables():
    """+-1 Hadamard matrices of order 4*odd used when use_rand=False (quant.py:8).  The 2.8 MB table
    file is a data asset of the reference checkpoint format; it is looked up in $QUIP_HADAMARD_PATH,
    next to this package, or the working directory (where the reference itself expects it)."""
    global _HAD_TABLES
    if _HAD_TABLES is None:
        cands = [os.environ.get("QUIP_HADAMARD_PATH"),
                 os.path.join(os.path.dirname(os.path.abspath(__file__)), _HAD_FILE),
                 os.path.join(os.getcwd(), _HAD_FILE)]
        _HAD_TABLES = {}
        for c in cands:
            if c and os.path.isfile(c):
                from safetensors.torch import load_file
                _HAD_TABLES = load_file(c)
                break
    return _HAD_TABLES


def register_had_table(order: int, matrix: torch.Tensor):
    """Install a +-1 Hadamard matrix of the given order (tests / users without the table file)."""
    _had_tables()[str(order)] = matrix.to(torch.float32)


def next_power_of_2(n):
    return 1 if n == 0 else 2 ** math.ceil(math.log(n, 2))


def get_power_of_2(n):
    """(k, odd) with n == odd * 2**k."""
    k = 0
    while n % 2 == 0:
        n //= 2
        k += 1
    return k, n


def get_hadK(n, use_rand=True):
    """(hadK or None, K, padded_n) -- reference: quant.py:26-39.
    power of two: pure FWHT.  use_rand: random orthogonal odd x odd block (overwritten from the
    checkpoint at load).  Otherwise a tabulated Hadamard of order 4*odd, or zero-padding to 2^k."""
    exp, base = get_power_of_2(n)
    if base == 1:
        return None, 1, n
    if use_rand:
        import scipy.stats
        return torch.tensor(scipy.stats.special_ortho_group.rvs(base)).to(torch.float32), base, n
    order = base * 4
    if exp < 2 or str(order) not in _had_tables():
        if exp >= 2 and not _had_tables() and 12 <= order <= 252:
            raise FileNotFoundError(
                f"use_rand=False needs the order-{order} Hadamard table for n={n}: put the checkpoint's "
                f"{_HAD_FILE} in the working directory or set QUIP_HADAMARD_PATH")
        return None, 1, next_power_of_2(n)
    return _had_tables()[str(order)] / math.sqrt(order), order, n


def _fwht_lastdim(x):
    """Unnormalised Sylvester-order FWHT over the last dim (power of two)."""
    n = x.shape[-1]
    lead = x.shape[:-1]
    h = 1
    while h < n:
        x = x.reshape(*lead, n // (2 * h), 2, h)
        x = torch.stack((x[..., 0, :] + x[..., 1, :], x[..., 0, :] - x[..., 1, :]), dim=-2)
        x = x.reshape(*lead, n)
        h *= 2
    return x


def matmul_hadU(X, hadK, K, padN, transpose=False):
    n = X.shape[-1]
    if padN != n:
        X = torch.nn.functional.pad(X, (0, padN - n))
    lead = X.shape[:-1]
    y = _fwht_lastdim(X.reshape(-1, K, padN // K))
    if K > 1:
        hk = hadK.T if transpose else hadK
        y = hk.to(device=y.device, dtype=y.dtype) @ y
    # the reference divides by `torch.tensor(padN / K).sqrt()` (quant.py:65), an fp32 scalar: for block lengths that are not
    # powers of 4 (128, 8192, ...) the divisor is the fp32-rounded root also when X is fp64 (quantise-time Hessians)
    return y.reshape(*lead, padN) / float(torch.tensor(padN / K).sqrt())


def matmul_hadUt(X, hadK, K, padN):
    return matmul_hadU(X, hadK, K, padN, transpose=True)


def matmul_hadU_cuda(X, hadK, K, n, scale=None, transpose=False):
    """reference: quant.py:72-84 (same op sequence: pad -> hadamard op -> hadK @)."""
    if n != X.shape[-1]:
        X = torch.nn.functional.pad(X, (0, n - X.shape[-1]))
    had_scale = (1.0 if scale is None else scale) / math.sqrt(n // K)
    if K == 1:
        return torch.ops.quip_lib.hadamard(X, had_scale)
    if transpose:
        hadK = hadK.T.contiguous()
    y = torch.ops.quip_lib.hadamard(X.reshape(-1, K, n // K), had_scale)
    return (hadK @ y).reshape(X.shape)


def matmul_hadUt_cuda(X, hadK, K, n, scale=None):
    return matmul_hadU_cuda(X, hadK, K, n, scale=scale, transpose=True)
