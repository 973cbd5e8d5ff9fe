"""D4 codebook: 2 bits/weight, 4-vector codes, 8-bit index (reference: codebook/d4.py).

256 points of the deep-hole-centred D4 lattice (half-integer coordinates, even coordinate sum) with
squared norm <= 9.  Index byte: low 5 bits select the absolute pattern, bits 5/6 flip elements 2/1,
element 3 is flipped to make the coordinate sum even, bit 7 negates everything."""
from functools import lru_cache

import torch
from torch import nn

_D4_CODESZ = 4


def _abs_pattern(i5):
    h, t, f = 0.5, 1.5, 2.5
    if i5 == 0:
        return [h, h, h, h]
    if i5 == 1:
        return [t, t, t, t]
    if i5 < 8:                      # two of one kind, two of the other; element 0 and element i5>>1 paired
        a, b = (t, h) if i5 & 1 else (h, t)
        x = [b] * 4
        x[0] = a
        x[i5 >> 1] = a
        return x
    if i5 < 16:                     # one odd element out
        a, b = (t, h) if i5 < 12 else (h, t)
        x = [b] * 4
        x[i5 & 3] = a
        return x
    if i5 < 20:                     # a single 5/2
        x = [h] * 4
        x[i5 & 3] = f
        return x
    r = i5 - 20                     # one 3/2 and one 5/2 at distinct positions
    p15, p25 = r & 3, r >> 2
    if p25 >= p15:
        p25 += 1
    x = [h] * 4
    x[p15] = t
    x[p25] = f
    return x


def code8_to_d4(i8):
    x = _abs_pattern(i8 & 31)
    if i8 & 32:
        x[2] = -x[2]
    if i8 & 64:
        x[1] = -x[1]
    if sum(x) % 2 != 0:
        x[3] = -x[3]
    if i8 & 128:
        x = [-v for v in x]
    return x


@lru_cache(maxsize=None)
def build_D4_CB() -> torch.Tensor:
    return torch.tensor([code8_to_d4(i) for i in range(256)], dtype=torch.float32)


class D4_codebook(nn.Module):

    def __init__(self, inference=False, **kwargs):
        super().__init__()
        self.id = "D4"
        self.register_buffer("grid", build_D4_CB().clone(), persistent=False)
        if not inference:
            self.register_buffer("grid_norm", (self.grid @ self.grid.T).diag(), persistent=False)
        self.codesz = _D4_CODESZ
        self.opt_scale = 1.21
        self.idx_dtype = torch.uint8
        self.packsz = 1
        self.pack_out = False
        self.version = 0

    def _quantize_noscale(self, X, return_idx=True):
        Xqidx = (2 * X @ self.grid.T - self.grid_norm).argmax(1)
        if return_idx:
            return self.grid[Xqidx, :], Xqidx.to(self.idx_dtype)
        return self.grid[Xqidx, :]

    def quantize(self, X, return_idx=True):
        assert X.shape[-1] == self.codesz
        return self._quantize_noscale(X, return_idx=return_idx)

    def maybe_pack_idxs(self, idxs):
        return idxs

    def decompress_weight(self, Qidxs):
        return torch.ops.quip_lib.decompress_d4_origorder(Qidxs, self.grid)

    def forward(self, input, Qidxs):
        if input.shape[0] < 24:      # D4's own threshold (codebook/d4.py:134)
            return torch.ops.quip_lib.d4_mm_origorder(input, Qidxs, self.grid)
        return input @ self.decompress_weight(Qidxs).t()
