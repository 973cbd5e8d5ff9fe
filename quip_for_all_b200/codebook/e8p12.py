"""E8P12 codebook: 2 bits/weight, 8-vector codes, 16-bit index (reference: codebook/e8p12.py).

The codebook is a subset of E8 + 1/4: 256 absolute-value patterns (227 points of |D8^| with norm^2 <= 10
plus 29 norm-12 points) x 2^7 even sign flips x (+-1/4 shift) = 2^16 codewords.
Index c: abs pattern c >> 8, sign byte c & 0xff.  `grid_packed_abs` is the int64[256] table the CUDA
kernels consume (8 packed int8 in units of 1/4).
"""
from functools import lru_cache

import torch
from torch import nn

_E8P_CODESZ = 8
# packed byte j of a table entry / decoded word holds weight element _PERM[j]
_PERM = (0, 2, 1, 3, 4, 6, 5, 7)

# the 29 extra patterns of squared norm 12 (five 3/2 and three 1/2), as numerators over 2
# (data table of the codebook definition, codebook/e8p12.py:28-60)
_NORM12_HEX = ("31113333 13113333 11313333 11133333 33313311 33313131 33311331 33313113 33311313 33311133 "
               "33133311 33133131 33131331 33133113 33131313 33131133 31333311 31333131 31331331 31333113 "
               "31331313 13331133 13333311 13333131 13331331 13333113 13331313 11331333 33113331").split()


def _d8_abs_patterns():
    """All patterns of odd numerators (1,3,5,7)/2 with sum of squares <= 10, lexicographic order (227 rows)."""
    rows = []
    stack = [((), 0)]
    while stack:
        prefix, n2 = stack.pop()
        if len(prefix) == 8:
            rows.append(prefix)
            continue
        rest = 7 - len(prefix)
        # push in descending order so that pops come out ascending (lexicographic)
        for v in (7, 5, 3, 1):
            m = n2 + v * v
            if m + rest <= 40:
                stack.append((prefix + (v,), m))
    return rows


@lru_cache(maxsize=None)
def get_packed_abs_grid() -> torch.Tensor:
    """int64[256] `grid_packed_abs` (reference: codebook/e8p12.py:63-79)."""
    rows = _d8_abs_patterns() + [tuple(int(ch) for ch in s) for s in _NORM12_HEX]
    assert len(rows) == 256
    a = torch.tensor(rows, dtype=torch.int64)            # numerators over 2
    a = a[:, list(_PERM)]
    odd_sum = (a.sum(1) // 2) % 2                         # parity of the (integer) coordinate sum
    a[:, 7] = a[:, 7] * (1 - 2 * odd_sum)
    q = a * 2                                             # units of 1/4: +-2, 6, 10, 14
    packed = torch.zeros(256, dtype=torch.int64)
    for j in range(8):
        packed |= (q[:, j] & 0xFF) << (8 * j)
    # byte 7 negative => the int64 is negative (sign-extended OR in the reference)
    return packed


@lru_cache(maxsize=None)
def get_full_grid():
    """(float32 [65536, 8] codewords, int64 [65536] indices).  Vectorised statement of the codebook
    definition (reference: codebook/e8p12.py:82-103)."""
    tab = get_packed_abs_grid()
    c = torch.arange(1 << 16, dtype=torch.int64)
    signs = c & 255
    parity = torch.zeros_like(c)
    for i in range(8):
        parity ^= (signs >> i) & 1
    signs = signs ^ parity
    packed = tab[c >> 8]
    grid = torch.zeros(1 << 16, 8)
    for i in range(8):
        j = _PERM[i]
        byte = (packed >> (8 * j)) & 0xFF
        val = torch.where(byte >= 128, byte - 256, byte).float() / 4
        neg = ((signs >> (7 - j)) & 1).bool()
        grid[:, i] = torch.where(neg, -val, val)
    grid += torch.where(parity.bool(), -0.25, 0.25).unsqueeze(1)
    return grid, c


class E8P12_codebook(nn.Module):
    """Same attributes / methods as the reference class (codebook/e8p12.py:106-156)."""

    def __init__(self, inference=False, **kwargs):
        super().__init__()
        self.id = "E8P12"
        self.opt_scale = 1.03
        self.codesz = _E8P_CODESZ
        self.idx_dtype = torch.int16
        self.packsz = 1
        self.pack_out = False
        self.version = 1
        self.register_buffer("grid_packed_abs", get_packed_abs_grid().clone(), persistent=False)
        if not inference:
            grid, _ = get_full_grid()
            self.register_buffer("grid", grid.clone(), persistent=False)
            self.register_buffer("grid_norm", grid.norm(dim=-1) ** 2, persistent=False)

    def round(self, X, grid, grid_norm):
        assert X.shape[-1] == self.codesz
        Xqidx = (2 * X @ grid.T - grid_norm).argmax(-1)
        return grid[Xqidx], Xqidx

    def quantize(self, X, return_idx=True):
        # CUDA fp32 rows (what LDLQ feeds, quant.py:128-129) go to the fused search kernel (csrc/nearest.cu), which never
        # materialises the [m, 65536] score matrix; other dtypes / devices evaluate the reference's torch expression
        from ..nearest import e8p_quantize, native_ok
        if native_ok(X):
            vals, idxs = e8p_quantize(X, self.grid_packed_abs)
        else:
            vals, idxs = self.round(X, self.grid, self.grid_norm)
        return (vals, idxs) if return_idx else vals

    def maybe_pack_idxs(self, idxs):
        return idxs

    def decompress_weight(self, Qidxs):
        return torch.ops.quip_lib.decompress_e8p_origorder(Qidxs, self.grid_packed_abs)

    def forward(self, input, Qidxs):
        # the reference's dispatch rule is M < 32 -> custom kernel, else decompress + GEMM (codebook/e8p12.py:147-155);
        # here the mm op also covers 17 <= M <= 256 with an in-kernel-decode tcgen05 GEMM (and falls back to the
        # dense path by itself for shapes it does not cover), so only large M re-materialises the dense weight
        if input.size(0) <= 256:
            return torch.ops.quip_lib.e8p_mm_origorder(input, Qidxs, self.grid_packed_abs)
        return input @ self.decompress_weight(Qidxs).T
