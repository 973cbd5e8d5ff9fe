"""HI4B1C: 4-bit scalar half-integer grid -- weight = nibble - 7.5, eight nibbles per int32
(reference: codebook/hi.py).  No lookup table on the inference path: the kernels compute the value."""
import torch
from torch import nn

_NIBBLE_TO_ELEMENT = (0, 2, 4, 6, 1, 3, 5, 7)   # nibble j of a packed int32 holds element [j] of the group


def get_grid():
    """float32 [16, 1]: -7.5 ... 7.5"""
    return (torch.arange(16, dtype=torch.float32) - 7.5).unsqueeze(-1)


class HI4B1C_codebook(nn.Module):
    id = "HI"
    opt_scale = 2.97
    codesz = 1
    idx_dtype = torch.int32
    packsz = 8
    pack_out = False
    version = 0

    def __init__(self, inference=False, **kwargs):
        super().__init__()
        if not inference:
            g = get_grid()
            self.register_buffer("grid", g, persistent=False)
            self.register_buffer("grid_norm", (g * g).sum(-1), persistent=False)

    def round(self, X, grid, grid_norm):
        assert X.shape[-1] == self.codesz
        best = (2 * X @ grid.T - grid_norm).argmax(-1)
        return grid[best], best

    def quantize(self, X, return_idx=True):
        vals, idx = self.round(X, self.grid, self.grid_norm)
        return (vals, idx.to(self.idx_dtype)) if return_idx else vals

    def maybe_pack_idxs(self, idxs):
        packed = torch.zeros_like(idxs[:, ::self.packsz])
        for j, e in enumerate(_NIBBLE_TO_ELEMENT):
            packed = packed + (idxs[:, e::self.packsz] << (4 * j))
        return packed

    def decompress_weight(self, Qidxs):
        return torch.ops.quip_lib.decompress_hi_origorder(Qidxs)

    def forward(self, input, Qidxs):
        if input.shape[0] < 32:
            return torch.ops.quip_lib.hi_mm_origorder(input, Qidxs)
        return input @ self.decompress_weight(Qidxs).t()
