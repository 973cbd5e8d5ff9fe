"""Engine-side binding of `quipb200_decode_step` (include/quip_b200.h): one persistent cooperative
kernel per bs=1 decode step over all decoder layers of a Llama-style stack of E8P12 QuantLinears.
Used by `modeling.LlamaDecodeEngine` when the model qualifies; otherwise the engine falls back to
the per-group launches of `fused.py` (still CUDA, never CPU)."""
import ctypes
from ctypes import POINTER, Structure, c_float, c_int32, c_size_t, c_void_p

import torch

from ._native import EUNSUPPORTED, LinearDesc, QuipB200Error, check, lib
from .fused import linear_desc


class DecodeLayer(Structure):
    """struct quipb200_decode_layer."""
    _fields_ = [("q", LinearDesc), ("k", LinearDesc), ("v", LinearDesc), ("o", LinearDesc),
                ("gate", LinearDesc), ("up", LinearDesc), ("down", LinearDesc),
                ("input_norm_w", c_void_p), ("post_norm_w", c_void_p), ("k_cache", c_void_p), ("v_cache", c_void_p),
                ("mlp_hk", c_void_p)]


class DecodePlan(Structure):
    """struct quipb200_decode_plan."""
    _fields_ = [("n_layers", c_int32), ("hidden", c_int32), ("n_heads", c_int32), ("n_kv_heads", c_int32),
                ("head_dim", c_int32), ("max_len", c_int32), ("norm_eps", c_float), ("reserved", c_int32),
                ("layers", c_void_p), ("cos_t", c_void_p), ("sin_t", c_void_p), ("pos", c_void_p)]


_bound = False


def _bind():
    global _bound
    L = lib()
    if not _bound:
        L.quipb200_decode_step_workspace_bytes.restype = c_size_t
        L.quipb200_decode_step_workspace_bytes.argtypes = [POINTER(DecodePlan), POINTER(DecodeLayer)]
        L.quipb200_decode_step.restype = ctypes.c_int
        L.quipb200_decode_step.argtypes = [POINTER(DecodePlan), POINTER(DecodeLayer), c_void_p, c_void_p, c_void_p,
                                           c_size_t, c_void_p]
        L.quipb200_decode_step_set_splits.restype = ctypes.c_int
        L.quipb200_decode_step_set_splits.argtypes = [ctypes.c_int]
        L.quipb200_decode_step_debug.restype = ctypes.c_int
        L.quipb200_decode_step_debug.argtypes = [c_void_p]
        _bound = True
    return L


class PersistentDecodeStep:
    """Holds the device-side layer table + scratch of one engine; `__call__(h_in, h_out)` enqueues the step.

    layers: HF LlamaDecoderLayer modules whose seven block linears are E8P12 QuantLinears.
    k_cache / v_cache: fp16 [L, 1, n_kv, max_len, head_dim]; cos / sin: fp16 [max_len, head_dim]; pos: int64 [1].
    Raises ValueError when the stack is outside what the kernel covers."""

    def __init__(self, layers, k_cache, v_cache, cos, sin, pos, n_heads, n_kv_heads, head_dim, hidden, eps):
        L = _bind()
        self.dev = k_cache.device
        n = len(layers)
        self.host_layers = (DecodeLayer * n)()
        self._keep = []
        for i, lyr in enumerate(layers):
            at, mlp = lyr.self_attn, lyr.mlp
            mods = (at.q_proj, at.k_proj, at.v_proj, at.o_proj, mlp.gate_proj, mlp.up_proj, mlp.down_proj)
            for m in mods:
                cbid = getattr(getattr(m, "codebook", None), "id", None)
                if cbid not in ("E8P12", "E8P12RVQ4B", "D4") or cbid != mods[0].codebook.id or getattr(m, "per_channel", False):
                    raise ValueError("persistent decode step: every block linear must be a QuantLinear of one codebook "
                                     "(E8P12, E8P12RVQ4B or D4)")
                for t in (m.SU, m.SV, m.bias, m.had_left, m.had_right):
                    if t is not None and t.dtype != torch.float16:
                        raise ValueError("persistent decode step needs fp16 scale / bias / hadK tensors")
            d = self.host_layers[i]
            d.q, d.k, d.v, d.o, d.gate, d.up, d.down = [linear_desc(m) for m in mods]
            nw1, nw2 = lyr.input_layernorm.weight, lyr.post_attention_layernorm.weight
            if nw1.dtype != torch.float16 or nw2.dtype != torch.float16:
                raise ValueError("persistent decode step needs fp16 norm weights")
            d.input_norm_w, d.post_norm_w = nw1.data_ptr(), nw2.data_ptr()
            d.k_cache, d.v_cache = k_cache[i].data_ptr(), v_cache[i].data_ptr()
            hk = None
            K = mlp.gate_proj.K_right
            if K > 1 and mlp.up_proj.K_right == K and mlp.down_proj.K_left == K and mlp.gate_proj.had_right is not None:
                Kp = (K + 15) // 16 * 16
                hk = torch.zeros(3, Kp, Kp, dtype=torch.float16, device=self.dev)
                hk[0, :K, :K] = mlp.gate_proj.had_right
                hk[1, :K, :K] = mlp.up_proj.had_right
                hk[2, :K, :K] = mlp.down_proj.had_left.t()      # input side: M[k_out][k_in] = hadK^T (quant.py:79-80)
            d.mlp_hk = hk.data_ptr() if hk is not None else None
            self._keep.append((mods, nw1, nw2, hk))
        raw = bytes(self.host_layers)
        self.dev_layers = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.dev)
        self.plan = DecodePlan(n, hidden, n_heads, n_kv_heads, head_dim, k_cache.shape[-2], float(eps), 0,
                               self.dev_layers.data_ptr(), cos.data_ptr(), sin.data_ptr(), pos.data_ptr())
        self._refs = (k_cache, v_cache, cos, sin, pos)
        self.ws_bytes = L.quipb200_decode_step_workspace_bytes(ctypes.byref(self.plan), self.host_layers)
        if self.ws_bytes == 0:
            raise ValueError("persistent decode step: model shape not covered")
        self.ws = torch.zeros(self.ws_bytes + 256, dtype=torch.uint8, device=self.dev)
        off = (-self.ws.data_ptr()) % 256
        self.ws_ptr = self.ws.data_ptr() + off

    def __call__(self, h_in, h_out):
        assert h_in.dtype == torch.float16 and h_out.dtype == torch.float16 and h_in.is_contiguous()
        rc = lib().quipb200_decode_step(ctypes.byref(self.plan), self.host_layers, h_in.data_ptr(), h_out.data_ptr(),
                                        self.ws_ptr, self.ws_bytes,
                                        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc == EUNSUPPORTED:
            raise QuipB200Error("decode_step: unsupported shape")
        check(rc, "decode_step")
        return h_out
