"""E8P12RVQ4B: 4 bits/weight = two E8P12 codes (main + scaled residual) per 8 weights
(reference: codebook/e8p12_rvq4.py).  Qidxs int32: (main << 16) + residual."""
import torch
from torch import nn

from .e8p12 import _E8P_CODESZ, get_full_grid, get_packed_abs_grid


class E8P12RVQ4B_codebook(nn.Module):

    def __init__(self, inference=False, opt_resid_scale=None, **kwargs):
        super().__init__()
        self.id = "E8P12RVQ4B"
        self.opt_scale = 1.03
        self.codesz = _E8P_CODESZ
        self.idx_dtype = torch.int32
        self.packsz = 1
        self.pack_out = False
        self.version = 0
        self.opt_resid_scale = 1 / 3.45 if opt_resid_scale is None else opt_resid_scale
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
        from ..nearest import e8p_quantize, native_ok
        if native_ok(X):      # both searches + the residual arithmetic in the fused kernels (csrc/nearest.cu)
            final_vals, final_idxs = e8p_quantize(X, self.grid_packed_abs, 2, self.opt_resid_scale)
            return (final_vals, final_idxs) if return_idx else final_vals
        init_vals, init_idxs = self.round(X, self.grid, self.grid_norm)
        resid = (X - init_vals) / self.opt_resid_scale
        resid_vals, resid_idxs = self.round(resid, self.grid, self.grid_norm)
        final_vals = init_vals + resid_vals * self.opt_resid_scale
        final_idxs = (init_idxs << 16) + resid_idxs
        return (final_vals, final_idxs) if return_idx else final_vals

    def maybe_pack_idxs(self, idxs):
        return idxs

    def decompress_weight(self, Qidxs):
        return torch.ops.quip_lib.decompress_e8prvq4_origorder(Qidxs, self.grid_packed_abs, self.opt_resid_scale)

    def forward(self, input, Qidxs):
        if input.size(0) < 32:
            return torch.ops.quip_lib.e8prvq4_mm_origorder(input, Qidxs, self.grid_packed_abs,
                                                           self.opt_resid_scale)
        return input @ self.decompress_weight(Qidxs).T
