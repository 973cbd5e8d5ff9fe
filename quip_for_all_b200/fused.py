"""Engine-side bindings of the grouped / hooked C-ABI entry points (include/quip_b200.h):
`quipb200_linear_group_forward` (up to 3 QuantLinears sharing an input in one launch, with RMSNorm /
silu-gate / residual folded in) and `quipb200_attn_decode`.  Used by `modeling.LlamaDecodeEngine`;
the drop-in module path (`QuantLinear.forward`) does not depend on this file."""
import ctypes

import torch

from . import _native
from ._native import CODEBOOK_ENUM, Fusion, LinearDesc, check, lib
from .register_lib import fused_supported


def _p(t):
    return t.data_ptr() if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def linear_desc(layer) -> LinearDesc:
    cb = layer.codebook
    grid = layer._grid_tensor()
    return LinearDesc(CODEBOOK_ENUM[cb.id], layer.in_features, layer.out_features, layer.q_in_features,
                      layer.q_out_features, layer.K_left, layer.K_right, float(layer.wscale_float),
                      float(getattr(cb, "opt_resid_scale", 0.0) or 0.0),
                      _p(layer.Qidxs), _p(grid), _p(layer.SU), _p(layer.SV), _p(layer.bias),
                      _p(layer.had_left), _p(layer.had_right), _p(layer.Wscale) if layer.per_channel else None)


class LinearGroup:
    """1..3 QuantLinears that read the same [M, in_features] fp16 input, launched together."""

    def __init__(self, layers, max_m=1):
        assert 1 <= len(layers) <= 3
        l0 = layers[0]
        for l in layers:
            if not fused_supported(l.codebook.id, l.q_in_features, max_m) or l.codebook.id != l0.codebook.id \
                    or l.in_features != l0.in_features:
                raise ValueError("layers cannot be grouped on the fused path")
            for t in (l.SU, l.SV, l.bias, l.had_left, l.had_right):
                if t is not None and t.dtype != torch.float16:
                    raise ValueError("fused path needs fp16 scale / bias / hadK tensors")
        self.layers = layers
        self.n = len(layers)
        self.dev = l0.Qidxs.device
        self.descs = (LinearDesc * self.n)(*[linear_desc(l) for l in layers])
        self.max_m = max_m
        L = lib()
        self.ws_bytes = L.quipb200_linear_group_workspace_bytes(self.descs, self.n, max_m)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.dev)
        self.out = [torch.empty(max_m, l.out_features, dtype=torch.float16, device=self.dev) for l in layers]
        self._yptr = (ctypes.c_void_p * self.n)(*[o.data_ptr() for o in self.out])
        self._ldy = (ctypes.c_int64 * self.n)(*[o.stride(0) for o in self.out])

    def __call__(self, x, norm_w=None, eps=0.0, gate=None, residual=None):
        """x: fp16 [M, in]; returns the list of static output tensors (overwritten by the next call)."""
        M = x.shape[0]
        assert M <= self.max_m and x.dtype == torch.float16 and x.stride(1) == 1
        fu = Fusion(_p(norm_w), float(eps), 0, _p(gate), gate.stride(0) if gate is not None else 0,
                    _p(residual), residual.stride(0) if residual is not None else 0)
        check(lib().quipb200_linear_group_forward(self.descs, self.n, ctypes.byref(fu), x.data_ptr(), x.stride(0),
                                                  self._yptr, self._ldy, M, self.ws.data_ptr(), self.ws_bytes,
                                                  _stream()), "linear_group_forward")
        return self.out


def attn_decode(q, k, v, k_cache, v_cache, cos, sin, pos, out, n_heads, n_kv_heads, head_dim):
    """RoPE + KV append + single-query attention (one launch). k_cache/v_cache: [1, n_kv, max_len, hd]."""
    max_len = k_cache.shape[-2]
    check(lib().quipb200_attn_decode(q.data_ptr(), k.data_ptr(), v.data_ptr(), k_cache.data_ptr(),
                                     v_cache.data_ptr(), cos.data_ptr(), sin.data_ptr(), pos.data_ptr(),
                                     out.data_ptr(), n_heads, n_kv_heads, head_dim, max_len, _stream()),
          "attn_decode")
    return out


class LmTail:
    """Final RMSNorm -> lm_head -> argmax (-> position increment) in one launch (csrc/lm_tail.cu)."""

    def __init__(self, norm_weight, eps, lm_head_weight):
        if lm_head_weight.dtype != torch.float16 or norm_weight.dtype != torch.float16 or not lm_head_weight.is_contiguous():
            raise ValueError("lm_tail needs contiguous fp16 lm_head / norm weights")
        self.nw, self.eps, self.W = norm_weight, float(eps), lm_head_weight
        self.vocab, self.hidden = lm_head_weight.shape
        nbytes = lib().quipb200_lm_tail_workspace_bytes()
        self.ws = torch.zeros(nbytes, dtype=torch.uint8, device=lm_head_weight.device)

    def __call__(self, h, tok_out, pos=None, logits_out=None):
        check(lib().quipb200_lm_tail(h.data_ptr(), self.nw.data_ptr(), self.eps, self.W.data_ptr(), None, self.hidden,
                                     self.vocab, tok_out.data_ptr(), None, _p(pos), _p(logits_out), self.ws.data_ptr(),
                                     self.ws.numel(), _stream()), "lm_tail")
        return tok_out
