"""Quantise-time side of the path: LDLQ adaptive rounding with codebook feedback and the per-layer
driver that turns an `nn.Linear` + a proxy Hessian into a packed `QuantLinear`.

What it restates from the reference (behaviour, not code):
  * `block_LDL`               quant.py:91-104    unit block-lower factor of the Cholesky factor
  * `LDLQ` / `LDLQ_buffered`  quant.py:107-232   hatW_k = Q(W_k + (W - hatW)_{>k} L[>k, k]), last 8-column group first;
                                                 optional re-rounding sweeps (`quip_tune_iters`)
  * `QUIP.add_batch`          quip.py:43-69      H = (2 / n_batches) sum_b X_b^T X_b, fp64
  * `QUIP.quant`              quip.py:71-184     dead columns, trace normalisation, optional W/H rescaling, sign vectors +
                                                 two-sided Hadamard incoherence processing, damped Cholesky, RMS scale /
                                                 `opt_scale`, LDLQ, the de-rotated weight and the `attr` dict `pack` takes

The per-group rounding call `cb.quantize` is the GPU search kernel (csrc/nearest.cu) for CUDA fp32 rows.  The column
sweep is organised differently from the reference's two variants: one running error matrix E = W - hatW, a GEMM per
block of groups for the feedback of everything already rounded to the right of the block, and a short in-block
recurrence -- the same sums, grouped so that the sequential part touches `block_cols` columns instead of n.
"""
import math
from typing import Optional

import torch
from torch import nn

from .quant import get_hadK, matmul_hadU, matmul_hadUt


def block_ldl(L: torch.Tensor, b: int) -> torch.Tensor:
    """L (lower Cholesky factor, n x n) -> L @ blockdiag(L_11, L_22, ...)^-1 with b x b diagonal blocks: the diagonal
    blocks become identity, so `Lb - I` is the strictly-block-lower feedback matrix (quant.py:91-104)."""
    n = L.shape[0]
    if n % b:
        raise ValueError("block_ldl: n must be a multiple of the code size")
    g = n // b
    blocks = L.reshape(g, b, g, b)
    diag = torch.stack([blocks[i, :, i, :] for i in range(g)])             # [g, b, b]
    inv = torch.linalg.inv(diag)
    out = torch.matmul(L.reshape(n, g, b).transpose(0, 1), inv)            # [g, n, b]: column block i times inv(L_ii)
    out = out.transpose(0, 1).reshape(n, n)
    if torch.isnan(out).any():
        raise ValueError("Hessian is not invertible")
    return out


@torch.no_grad()
def ldlq(W: torch.Tensor, H: torch.Tensor, L: torch.Tensor, cb, tune_iters: int = 0, block_cols: int = 128):
    """Returns (hatW [m, n] in W's dtype, Qidxs [m, n / codesz] in `cb.idx_dtype`).  W is already rotated and scaled."""
    m, n = W.shape
    cs = cb.codesz
    if n % cs:
        raise ValueError("ldlq: columns must be a multiple of the code size")
    G = n // cs
    Lb = block_ldl(L.clone(), cs)
    hatW = torch.zeros_like(W)
    E = torch.zeros_like(W)                       # W - hatW for columns already rounded, 0 elsewhere
    Q = torch.zeros(m, G, dtype=cb.idx_dtype, device=W.device)
    gpb = max(1, block_cols // cs)                # groups per block
    for g_hi in range(G, 0, -gpb):
        g_lo = max(0, g_hi - gpb)
        c_lo, c_hi = g_lo * cs, g_hi * cs
        # feedback of every column to the right of the block, one GEMM
        carry = E[:, c_hi:] @ Lb[c_hi:, c_lo:c_hi] if c_hi < n else torch.zeros(m, c_hi - c_lo, dtype=W.dtype, device=W.device)
        for g in range(g_hi - 1, g_lo - 1, -1):
            a, b = g * cs, (g + 1) * cs
            target = W[:, a:b] + carry[:, a - c_lo:b - c_lo]
            if b < c_hi:
                target = target + E[:, b:c_hi] @ Lb[b:c_hi, a:b]
            vals, idx = cb.quantize(target.contiguous())
            hatW[:, a:b] = vals.to(W.dtype)
            E[:, a:b] = W[:, a:b] - hatW[:, a:b]
            Q[:, g] = idx.to(cb.idx_dtype)
    for _ in range(tune_iters):
        for g in range(G - 1, -1, -1):
            a, b = g * cs, (g + 1) * cs
            Hgg_inv = torch.linalg.inv(H[a:b, a:b])
            target = hatW[:, a:b] + (W - hatW) @ H[:, a:b] @ Hgg_inv
            vals, idx = cb.quantize(target.contiguous())
            hatW[:, a:b] = vals.to(W.dtype)
            Q[:, g] = idx.to(cb.idx_dtype)
    return hatW, Q


def proxy_loss(W: torch.Tensor, hatW: torch.Tensor, H: torch.Tensor) -> float:
    """tr((W - hatW) H (W - hatW)^T) / tr(W H W^T): the quantity LDLQ minimises, for reporting / tests."""
    D = (W - hatW).to(H.dtype)
    Wd = W.to(H.dtype)
    return float(((D @ H) * D).sum() / ((Wd @ H) * Wd).sum())


class LayerQuantizer:
    """Collects the proxy Hessian of one linear layer and quantises it (reference class: quip.py `QUIP`)."""

    def __init__(self, layer: nn.Linear, cb):
        if not isinstance(layer, nn.Linear):
            raise TypeError("LayerQuantizer covers nn.Linear (the decoder-block linears of the path)")
        self.layer = layer
        self.dev = layer.weight.device
        self.rows, self.columns = layer.weight.shape
        self.H = torch.zeros(self.columns, self.columns, dtype=torch.float64, device=self.dev)
        self.mu = torch.zeros(self.columns, dtype=torch.float64, device=self.dev)
        self.nsamples = 0
        self.cb = cb.to(self.dev)

    @torch.no_grad()
    def add_batch(self, inp: torch.Tensor, out=None):
        """Running mean over calls of 2 X^T X (X = the call's tokens x columns), and of the column sums (quip.py:43-69)."""
        x = inp.reshape(-1, inp.shape[-1]).to(device=self.dev, dtype=torch.float64)
        calls = 1 if inp.dim() <= 2 else inp.shape[0]
        keep = self.nsamples / (self.nsamples + calls)
        self.H *= keep
        self.mu *= keep
        self.nsamples += calls
        self.mu += x.sum(dim=0) / self.nsamples
        xs = math.sqrt(2.0 / self.nsamples) * x
        self.H += xs.T @ xs

    def _damped_cholesky(self, H, sigma_reg):
        idx = torch.arange(H.shape[0], device=H.device)
        for _ in range(10):          # every attempt adds sigma_reg to the diagonal again (quip.py:131-143)
            H[idx, idx] += sigma_reg
            try:
                L = torch.linalg.cholesky(H)
            except RuntimeError:
                continue
            if not torch.isnan(L).any():
                return L
        raise ValueError("Hessian is not invertible")

    @torch.no_grad()
    def quantize(self, rescale_WH=False, use_fp64=False, sigma_reg=0.01, scale_override=0, use_rand=True,
                 per_channel=False, quip_tune_iters=0, SU: Optional[torch.Tensor] = None,
                 SV: Optional[torch.Tensor] = None, block_cols=128):
        """Returns the `attr` dict of the reference (`QuantLinear.pack` input) and overwrites `layer.weight` with the
        de-rotated quantised weight, as quip.py:160-168 does."""
        dt = torch.float64 if use_fp64 else torch.float32
        H = self.H.to(dt).clone()
        w = self.layer.weight.data.clone().to(dt)
        dead = torch.diag(H) == 0
        H[dead, dead] = 1
        w[:, dead] = 0
        H /= torch.diag(H).mean()
        scaleWH = None
        if rescale_WH:
            H /= H.abs().max()
            dH = torch.diag(H).clamp(min=1e-8)
            dW = torch.diag(w.T @ w).clamp(min=1e-8)      # (the reference's expression, quip.py:104: same summation order)
            scaleWH = (dH / dW).sqrt().sqrt().to(torch.float32).clamp(min=1e-8)
            w *= scaleWH[None, :]
            H /= scaleWH[None, :]
            H /= scaleWH[:, None]
        merge_su = SU is not None or hasattr(self.layer, "SU")
        merge_sv = SV is not None or hasattr(self.layer, "SV")
        if SU is None:
            SU = self.layer.SU if hasattr(self.layer, "SU") else (torch.randn(self.columns, device=self.dev).sign() + 1e-5).sign()
        if SV is None:
            SV = self.layer.SV if hasattr(self.layer, "SV") else (torch.randn(self.rows, device=self.dev).sign() + 1e-5).sign()
        SU, SV = SU.to(device=self.dev, dtype=dt), SV.to(device=self.dev, dtype=dt)
        hl, Kl, Nl = get_hadK(self.columns, use_rand=use_rand)
        hr, Kr, Nr = get_hadK(self.rows, use_rand=use_rand)
        # incoherence processing: H <- U (SU H SU) U^T,  W <- V (SV W SU) U^T   (U, V: orthogonal Hadamard-type transforms)
        H = matmul_hadUt(matmul_hadUt(H * SU, hl, Kl, Nl).T * SU, hl, Kl, Nl)
        w = matmul_hadUt(matmul_hadUt(w.T * SV, hr, Kr, Nr).T * SU, hl, Kl, Nl)
        L = self._damped_cholesky(H, sigma_reg)
        if per_channel:
            w_scale = w.square().mean(dim=1, keepdim=True).sqrt()
        else:
            w_scale = w.square().mean().sqrt()
        w_scale = w_scale / (scale_override if scale_override > 0 else self.cb.opt_scale)
        w = w / w_scale
        hat_w, Qidxs = ldlq(w, H, L, self.cb, quip_tune_iters, block_cols=block_cols)
        self.last_proxy_loss = proxy_loss(w, hat_w, H)
        hat_w = hat_w * w_scale
        back = matmul_hadU(hat_w, hl, Kl, Nl)[..., :self.columns] * SU
        back = (matmul_hadU(back.T, hr, Kr, Nr)[..., :self.rows] * SV).T
        if rescale_WH:
            back = back / scaleWH[None, :]
        self.layer.weight.data = back.reshape(self.layer.weight.shape).to(self.layer.weight.dtype)
        return {
            "left_hadK": hl.cpu() if (use_rand and hl is not None) else None,
            "right_hadK": hr.cpu() if (use_rand and hr is not None) else None,
            "Qidxs": self.cb.maybe_pack_idxs(Qidxs).cpu(),
            "w_scale": w_scale.cpu(),
            "SU": SU.cpu(), "SV": SV.cpu(),
            "merge_su": merge_su, "merge_sv": merge_sv,
            "scaleWH": scaleWH.cpu() if rescale_WH else None,
        }


@torch.no_grad()
def quantize_linear(layer: nn.Linear, calib_inputs, codebook: str = "E8P12", weight_dtype=torch.float16, **kw):
    """nn.Linear + calibration activations ([..., in_features] tensors) -> packed `QuantLinear` on the layer's device,
    ready for `forward` (the tail of the reference's per-layer loop, quantizer.py:560-600: quant -> pack)."""
    from .codebook import codebook_id
    from .qlinear import QuantLinear
    dev = layer.weight.device
    cb = codebook_id[codebook](inference=False)
    lq = LayerQuantizer(layer, cb)
    for x in calib_inputs:
        lq.add_batch(x)
    use_rand = kw.get("use_rand", True)
    per_channel = kw.get("per_channel", False)
    attr = lq.quantize(**kw)
    # with fresh random sign vectors nothing was merged into neighbouring layers, so SU / SV stay in the module
    attr["merge_su"], attr["merge_sv"] = False, False
    ql = QuantLinear(layer.in_features, layer.out_features, codebook_id[codebook](inference=True),
                     bias=layer.bias is not None, use_rand=use_rand, per_channel=per_channel, weight_dtype=weight_dtype)
    ql.pack(layer, attr)
    ql = ql.to(dev).train(layer.training)
    ql.wscale_float = float(ql.Wscale.float().mean().item())        # quantizer.py:837
    if per_channel:
        ql.Wscale.data = ql.Wscale.data / ql.wscale_float
    ql.proxy_loss = lq.last_proxy_loss
    return ql
