"""Shared test helpers: oracle-side evaluation of a product QuantLinear module."""
import numpy as np
import torch

import quip_oracle as qo


def oracle_w_hat(codebook_id, qidxs_np, resid_scale=None):
    if codebook_id == "E8P12":
        return qo.decompress_e8p(qidxs_np)
    if codebook_id == "E8P12RVQ4B":
        return qo.decompress_e8prvq4(qidxs_np, resid_scale if resid_scale is not None else qo.RVQ4_DEFAULT_RESID_SCALE)
    if codebook_id == "D4":
        return qo.decompress_d4(qidxs_np)
    if codebook_id == "E8P12RVQ3B":
        return qo.decompress_e8prvq3(qidxs_np, resid_scale if resid_scale is not None else qo.RVQ3_DEFAULT_RESID_SCALE)
    if codebook_id == "HI":
        return qo.decompress_hi(qidxs_np)
    raise KeyError(codebook_id)


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def oracle_forward(layer, x, rounding="reference"):
    """Oracle QuantLinear.forward for a product `QuantLinear` module (any device) and fp16 input x."""
    cb = layer.codebook
    W_hat = oracle_w_hat(cb.id, _np(layer.Qidxs), getattr(cb, "opt_resid_scale", None))
    return qo.quantlinear_forward(
        _np(x), W_hat=W_hat, in_features=layer.in_features, out_features=layer.out_features,
        q_in=layer.q_in_features, q_out=layer.q_out_features,
        SU=_np(layer.SU), SV=_np(layer.SV), bias=_np(layer.bias),
        wscale_float=layer.wscale_float,
        Wscale_per_channel=_np(layer.Wscale) if layer.per_channel else None,
        had_left=_np(layer.had_left), K_left=layer.K_left, had_right=_np(layer.had_right), K_right=layer.K_right,
        rounding=rounding)


def make_layer(fin, fout, codebook="E8P12", bias=False, use_rand=True, per_channel=False, seed=0, device="cpu",
               trained_scales=True):
    """A randomly filled product QuantLinear (post-load tricks applied)."""
    from quip_for_all_b200 import QuantLinear, codebook_id
    from quip_for_all_b200.quantizer import apply_load_time_tricks
    g = torch.Generator().manual_seed(seed)
    cb = codebook_id[codebook](inference=True)
    layer = QuantLinear(fin, fout, cb, bias=bias, use_rand=use_rand, per_channel=per_channel)
    info = torch.iinfo(cb.idx_dtype)
    layer.Qidxs.copy_(torch.randint(info.min, info.max + 1, layer.Qidxs.shape, dtype=torch.int64, generator=g)
                      .to(cb.idx_dtype))
    sgn = lambda n: (torch.randint(0, 2, (n,), generator=g) * 2 - 1).float()
    amp = (lambda n: 1 + 0.1 * torch.randn(n, generator=g)) if trained_scales else (lambda n: torch.ones(n))
    layer.SU.data.copy_((sgn(fin) * amp(fin)).half())
    layer.SV.data.copy_((sgn(fout) * amp(fout)).half())
    if per_channel:
        layer.Wscale.copy_((0.02 * (1 + 0.2 * torch.rand(layer.q_out_features, generator=g))).half())
    else:
        layer.Wscale.fill_(0.02 / 1.09375)
    if bias:
        layer.bias.copy_((0.1 * torch.randn(fout, generator=g)).half())
    apply_load_time_tricks(torch.nn.Sequential(layer))
    return layer.to(device).eval()


def tol_of(ref):
    """Stated fp16 tolerance of the fused forward against the reference-rounding oracle: the
    reference chains up to 8 fp16 roundings on values of the output's magnitude; 2^-8 of the largest
    output covers that chain plus the 16-bit fixed-point activation (DESIGN.md 'tolerance')."""
    return 2.0 ** -8 * float(np.abs(ref).max())
