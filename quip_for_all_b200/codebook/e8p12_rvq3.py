"""E8P12RVQ3B: 3 bits/weight = E8P12 code (16 bit) + 8-bit index into a 256-point E8 residual table
(reference: codebook/e8p12_rvq3.py).  Packed as byte triplets [resid, idx_lo, idx_hi] inside int32."""
from fractions import Fraction
from functools import lru_cache
from itertools import combinations, product

import torch
from torch import nn

from .e8p12 import _E8P_CODESZ, get_full_grid, get_packed_abs_grid


@lru_cache(maxsize=None)
def get_e81bgrid() -> torch.Tensor:
    """float32 [256, 8]: E8 points of squared norm <= 2 (integer points first, then half-integer
    points, each block in lexicographic order) + 15 axis points of norm 4 (codebook/e8p12_rvq3.py:16-50)."""
    ints = [(0.0,) * 8]
    for i, j in combinations(range(8), 2):
        for si, sj in product((-1.0, 1.0), repeat=2):
            v = [0.0] * 8
            v[i], v[j] = si, sj
            ints.append(tuple(v))
    halves = [s for s in product((-0.5, 0.5), repeat=8) if sum(s) % 2 == 0]
    axis = []
    for sgn in (2.0, -2.0):
        for i in range(8):
            if sgn < 0 and i == 7:
                continue          # the reference leaves this one out to stay at 256 entries
            v = [0.0] * 8
            v[i] = sgn
            axis.append(tuple(v))
    return torch.tensor(sorted(ints) + sorted(halves) + axis, dtype=torch.float32)


@lru_cache(maxsize=None)
def _pack_e81b_cached():
    return pack_e81b(get_e81bgrid())


def pack_e81b(cba: torch.Tensor) -> torch.Tensor:
    """int32 [256]: nibble j = 2*v[[0,2,4,6,1,3,5,7][j]] & 0xf (codebook/e8p12_rvq3.py:53-62)."""
    q = (cba[:, [0, 2, 4, 6, 1, 3, 5, 7]] * 2).to(torch.int64) & 0xF
    acc = torch.zeros(q.shape[0], dtype=torch.int64)
    for j in range(8):
        acc |= q[:, j] << (4 * j)
    acc = torch.where(acc >= 2 ** 31, acc - 2 ** 32, acc)
    return acc.to(torch.int32)


class E8P12RVQ3B_codebook(nn.Module):

    def __init__(self, inference=False, opt_resid_scale=None, **kwargs):
        super().__init__()
        self.id = "E8P12RVQ3B"
        self.opt_scale = 0.98
        self.codesz = _E8P_CODESZ
        self.idx_dtype = torch.int32
        self.packsz = Fraction(4, 3)
        self.pack_out = False
        self.version = 0
        self.opt_resid_scale = 1 / 2.04 if opt_resid_scale is None else opt_resid_scale
        self.register_buffer("grid_packed_abs", get_packed_abs_grid().clone(), persistent=False)
        self.register_buffer("e81b_grid", get_e81bgrid().clone(), persistent=False)
        self.register_buffer("e81b_grid_packed", _pack_e81b_cached().clone(), persistent=False)
        if not inference:
            grid, _ = get_full_grid()
            self.register_buffer("grid", grid.clone(), persistent=False)
            self.register_buffer("grid_norm", grid.norm(dim=-1) ** 2, persistent=False)
            self.register_buffer("e81b_grid_norm", self.e81b_grid.norm(dim=-1) ** 2, persistent=False)

    def round(self, X, grid, grid_norm):
        assert X.shape[-1] == self.codesz
        Xqidx = (2 * X @ grid.T - grid_norm).argmax(-1)
        return grid[Xqidx], Xqidx

    def quantize(self, X, return_idx=True):
        from ..nearest import e8p_quantize, native_ok
        if native_ok(X):      # E8P12 search + the 256-entry residual search in the fused kernels (csrc/nearest.cu)
            final_vals, final_idxs = e8p_quantize(X, self.grid_packed_abs, 2, self.opt_resid_scale, resid_grid=self.e81b_grid)
            return (final_vals, final_idxs) if return_idx else final_vals
        init_vals, init_idxs = self.round(X, self.grid, self.grid_norm)
        resid = (X - init_vals) / self.opt_resid_scale
        resid_vals, resid_idxs = self.round(resid, self.e81b_grid, self.e81b_grid_norm)
        final_vals = init_vals + resid_vals * self.opt_resid_scale
        final_idxs = (init_idxs << 8) + resid_idxs
        return (final_vals, final_idxs) if return_idx else final_vals

    def maybe_pack_idxs(self, idxs):
        # keep the low 3 bytes of every (little-endian) int32 code
        b = idxs.to(torch.int32).contiguous().view(torch.int8).view(idxs.shape[0], idxs.shape[1], -1)
        return b[..., :3].reshape(idxs.shape[0], -1).contiguous().view(torch.int32)

    def decompress_weight(self, Qidxs):
        return torch.ops.quip_lib.decompress_e8prvq3_origorder(
            Qidxs, self.grid_packed_abs, self.e81b_grid_packed, self.opt_resid_scale)

    def forward(self, input, Qidxs):
        if input.size(0) < 32:
            return torch.ops.quip_lib.e8prvq3_mm_origorder(
                input, Qidxs, self.grid_packed_abs, self.e81b_grid_packed, self.opt_resid_scale)
        return input @ self.decompress_weight(Qidxs).T
