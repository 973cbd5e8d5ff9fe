"""QuantLinear -- drop-in replacement for the reference module (qlinear.py:8-159).

Same constructor, attributes, buffers / parameters and state-dict keys
(`SU, SV, Qidxs, Wscale, weight, [bias], [had_left], [had_right]`), same `forward / pack / calc_weight`.
The eval forward computes

    y = SV * [(hadK_R (x) H) ( decode(Qidxs) ((hadK_L^T (x) H)(SU * x)) * wscale )][:out] + bias

For decode-sized inputs (M <= 16 rows) on CUDA the whole chain is ONE fused op
(`quip_lib::quantlinear_fwd`, three kernels); otherwise the reference's op sequence is issued through the
individual `quip_lib` ops (hadamard -> codebook mm / decompress + GEMM -> hadamard).
There is no CPU path: the ops are CUDA-only, as in the reference.
"""
import torch
import torch.nn as nn

from . import register_lib  # noqa: F401  (defines torch.ops.quip_lib)
from ._native import CODEBOOK_ENUM
from .quant import get_hadK, matmul_hadU_cuda, matmul_hadUt_cuda
from .register_lib import fused_supported


class QuantLinear(nn.Module):

    def __init__(self, in_features, out_features, codebook, bias=True, use_rand=True, per_channel=False,
                 weight_dtype=torch.float16):
        super().__init__()
        # peft looks for infeatures / outfeatures
        self.in_features = self.infeatures = in_features
        self.out_features = self.outfeatures = out_features
        self.codebook = codebook
        self.use_rand = use_rand
        self.per_channel = per_channel
        self.weight_dtype = weight_dtype

        had_left, self.K_left, self.q_in_features = get_hadK(in_features, use_rand)
        had_right, self.K_right, self.q_out_features = get_hadK(out_features, use_rand)
        # random orthogonal blocks are part of the checkpoint; tabulated Hadamards are not
        for name, mat in (("had_left", had_left), ("had_right", had_right)):
            if mat is not None:
                self.register_buffer(name, mat.to(weight_dtype), persistent=use_rand)
            else:
                setattr(self, name, None)

        per_idx = codebook.codesz * codebook.packsz
        if codebook.pack_out:
            shape = (self.q_out_features // codebook.packsz, self.q_in_features // codebook.codesz)
        else:
            shape = (self.q_out_features, int(self.q_in_features // per_idx))
        self.register_buffer("Qidxs", torch.zeros(shape, dtype=codebook.idx_dtype))

        self.SU = nn.Parameter(torch.ones(in_features, dtype=weight_dtype), requires_grad=True)
        self.SV = nn.Parameter(torch.ones(out_features, dtype=weight_dtype), requires_grad=True)
        if per_channel:
            self.register_buffer("Wscale", torch.ones(self.q_out_features, dtype=weight_dtype))
        else:
            self.register_buffer("Wscale", torch.ones((), dtype=torch.float))
        self.wscale_float = 1.0
        # transformers reads `.weight` to find the module's device / dtype
        self.register_buffer("weight", torch.zeros((), dtype=weight_dtype))
        if bias:
            self.register_buffer("bias", torch.zeros(out_features, dtype=weight_dtype))
        else:
            self.bias = None

    # ------------------------------------------------------------------------------------------
    def _fused_ok(self, x):
        if not (x.is_cuda and fused_supported(self.codebook.id, self.q_in_features, x.shape[0])):
            return False
        # the fused kernels read SU / SV / bias / hadK / per-channel Wscale as fp16
        for t in (self.SU, self.SV, self.bias, self.had_left, self.had_right,
                  self.Wscale if self.per_channel else None):
            if t is not None and t.dtype != torch.float16:
                return False
        return self.Qidxs.is_contiguous()

    def _grid_tensor(self):
        cb = self.codebook
        if cb.id == "D4":
            g = cb.grid
            if g.dtype != torch.float16:          # kernel contract: fp16 [256, 4]
                g = g.to(torch.float16)
                cb.grid = g
            return g
        return cb.grid_packed_abs

    def _batched_fused_ok(self, x):
        from .register_lib import rotate_supported
        if not x.is_cuda or x.dtype != torch.float16 or self.weight_dtype != torch.float16:
            return False
        if self.in_features % 8 or self.out_features % 8:
            return False
        for t in (self.SU, self.SV, self.bias, self.had_left, self.had_right):
            if t is not None and t.dtype != torch.float16:
                return False
        return rotate_supported(self.q_in_features, self.K_left) and rotate_supported(self.q_out_features, self.K_right)

    def _batched_umma_ok(self, x):
        """4 <= M <= 16 rows: rotations as one pass each + the tcgen05 mm (codes decoded once for all rows) beats the
        single fused launch, whose GEMV runs once per row."""
        from .register_lib import umma_preferred
        cb = CODEBOOK_ENUM.get(self.codebook.id)
        return (self.codebook.id in ("E8P12", "E8P12RVQ4B", "D4") and not self.per_channel and self._batched_fused_ok(x)
                and umma_preferred(x.shape[0], self.q_out_features, self.q_in_features, cb))

    def _hk_padded(self, side):
        """Zero-padded [Kp, Kp] coefficient matrix M[k_out][k_in] of the K x K block mix: hadK^T on the input side
        (quant.py:79-80), hadK on the output side.  Cached; rebuilt when the buffer changes."""
        had = self.had_left if side == "left" else self.had_right
        K = self.K_left if side == "left" else self.K_right
        if K == 1 or had is None:
            return None
        cache = self.__dict__.setdefault("_hk_cache", {})
        key = (side, had.data_ptr(), had._version, had.device)
        hit = cache.get(side)
        if hit is not None and hit[0] == key:
            return hit[1]
        Kp = (K + 15) // 16 * 16
        pad = torch.zeros(Kp, Kp, dtype=torch.float16, device=had.device)
        pad[:K, :K] = had.t() if side == "left" else had
        cache[side] = (key, pad)
        return pad

    def forward(self, input):
        x = input.reshape(-1, input.shape[-1])
        # the kernels want 16-byte aligned rows (8-element row pitch): sliced / offset activation views that the reference
        # accepts are copied once instead of being rejected by the C ABI
        if x.is_cuda and (x.stride(-1) != 1 or x.data_ptr() % 16 or (x.shape[0] > 1 and (x.stride(0) * x.element_size()) % 16)):
            x = x.clone(memory_format=torch.contiguous_format)
        x_dtype = x.dtype
        if self.training:
            if self.SU is not None:
                x = x * self.SU
            if x.shape[-1] != self.q_in_features:
                x = torch.nn.functional.pad(x, (0, self.q_in_features - x.shape[-1]))
            W = self.W if hasattr(self, "W") else self.calc_weight(cache=False).to(x.dtype)
            out = (x @ W)[..., :self.out_features]
            if self.SV is not None:
                out = out * self.SV
        elif self._fused_ok(x) and not (x.shape[0] >= 4 and self._batched_umma_ok(x)):
            xh = x if x_dtype == torch.float16 else x.to(torch.float16)
            cb = self.codebook
            out = torch.ops.quip_lib.quantlinear_fwd(
                xh, self.Qidxs, self._grid_tensor(),
                self.SU, self.SV, self.bias, self.had_left, self.had_right,
                self.Wscale if self.per_channel else None,
                CODEBOOK_ENUM[cb.id], self.in_features, self.out_features,
                self.q_in_features, self.q_out_features, self.K_left, self.K_right,
                float(self.wscale_float), float(getattr(cb, "opt_resid_scale", 0.0) or 0.0))
            if x_dtype != torch.float16:
                out = out.to(x_dtype)
            return out.view(*input.shape[:-1], out.shape[-1])   # SV and bias already applied
        elif self._batched_fused_ok(x):
            # M >= 17: same chain, the two rotations (with SU / SV / bias and the slice folded in) as one pass each
            import math
            out = torch.ops.quip_lib.rotate_fused(
                x, self.SU, self._hk_padded("left"), None, None, self.q_in_features, self.K_left, self.q_in_features,
                float(self.wscale_float) / math.sqrt(self.q_in_features // self.K_left))
            out = self.codebook(out, self.Qidxs)
            if self.per_channel:
                out = out * self.Wscale
            out = torch.ops.quip_lib.rotate_fused(
                out, None, self._hk_padded("right"), self.SV, self.bias, self.q_out_features, self.K_right,
                self.out_features, 1.0 / math.sqrt(self.q_out_features // self.K_right))
            return out.view(*input.shape[:-1], out.shape[-1])   # SV and bias already applied
        else:
            # the reference's op sequence (qlinear.py:90-112)
            if self.SU is not None:
                x = x * self.SU
            x = matmul_hadUt_cuda(x, self.had_left, self.K_left, self.q_in_features, self.wscale_float)
            if x_dtype != torch.float16:
                x = x.to(torch.float16)
            out = self.codebook(x, self.Qidxs)
            if x_dtype != torch.float16:
                out = out.to(dtype=x_dtype)
            if self.per_channel:
                out = out * self.Wscale
            out = matmul_hadU_cuda(out, self.had_right, self.K_right,
                                   self.q_out_features)[..., :self.out_features]
            if self.SV is not None:
                out = out * self.SV
        out = out.view(*input.shape[:-1], out.shape[-1])
        return out + self.bias if self.bias is not None else out

    # ------------------------------------------------------------------------------------------
    def pack(self, linear, attr):
        """Fill the buffers from a quantiser result dict (reference: qlinear.py:117-142)."""
        scaleWH, SU, SV = attr["scaleWH"], attr["SU"], attr["SV"]
        if scaleWH is not None:
            self.SU.data.copy_(scaleWH if attr["merge_su"] else SU * scaleWH)
        elif not attr["merge_su"]:
            self.SU.data.copy_(SU)
        else:
            self.SU = None
        if attr["merge_sv"]:
            self.SV = None
        else:
            self.SV.data.copy_(SV)
        self.Qidxs.copy_(attr["Qidxs"])
        self.Wscale.copy_(attr["w_scale"].squeeze() if self.per_channel else attr["w_scale"])
        if attr["left_hadK"] is not None:
            self.had_left.copy_(attr["left_hadK"])
        if attr["right_hadK"] is not None:
            self.had_right.copy_(attr["right_hadK"])
        if linear.bias is not None:
            self.bias.copy_(linear.bias / SV if attr["merge_sv"] else linear.bias)

    @torch.no_grad()
    def calc_weight(self, cache=True):
        """Dense (q_in, q_out) weight for `x @ W` in training mode (reference: qlinear.py:144-159)."""
        weight = self.codebook.decompress_weight(self.Qidxs)
        wscale_float = self.Wscale.mean().float().item()
        had_left = self.had_left.to(weight.dtype) if self.had_left is not None else None
        had_right = self.had_right.to(weight.dtype) if self.had_right is not None else None
        rows = matmul_hadU_cuda(weight, had_left, self.K_left, self.q_in_features, wscale_float)
        W = matmul_hadU_cuda(rows.T, had_right, self.K_right, self.q_out_features).to(self.weight_dtype)
        if self.per_channel:
            W = W * self.Wscale / self.Wscale.mean()
        if cache:
            self.register_buffer("W", W, persistent=False)
        return W


QuipLinear = QuantLinear   # BASELINE.json's name for the same class
