"""Model-level harness around the QuantLinear hot path.

* `make_random_quantized_llama` -- a HF `LlamaForCausalLM` whose block linears are `QuantLinear`s
  with random-init packed weights (BASELINE configs 2-5: no checkpoints / network are available).
  Synthetic-input recipe: SURVEY.md section 8(d) / BASELINE.md section 4.
* `LlamaDecodeEngine` -- bs=1 greedy decode loop over such a model (or one loaded with
  `load_quantized_model`): static KV cache + whole-step CUDA graph, the stand-in for the reference's
  `example_generate.py --compile` (StaticCache + torch.compile(reduce-overhead), example_generate.py:62-70).
  It can own a contiguous slice of the decoder layers, which is the unit of the multi-GPU layer pipeline.
"""
import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from .codebook import codebook_id
from .qlinear import QuantLinear
from .quantizer import QuipQuantizer, apply_load_time_tricks

LLAMA2_SHAPES = {
    # hidden, intermediate, layers, heads, kv_heads
    "llama2-7b": dict(hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                      num_attention_heads=32, num_key_value_heads=32),
    "llama2-13b": dict(hidden_size=5120, intermediate_size=13824, num_hidden_layers=40,
                       num_attention_heads=40, num_key_value_heads=40),
    "llama2-70b": dict(hidden_size=8192, intermediate_size=28672, num_hidden_layers=80,
                       num_attention_heads=64, num_key_value_heads=8),
    "tiny": dict(hidden_size=256, intermediate_size=704, num_hidden_layers=2,
                 num_attention_heads=4, num_key_value_heads=4),
    # head_dim 128 (fused attention path), GQA, non power-of-two intermediate (11 * 128)
    "tiny128": dict(hidden_size=512, intermediate_size=1408, num_hidden_layers=2,
                    num_attention_heads=4, num_key_value_heads=2),
    # shapes the persistent decode-step kernel covers: intermediate = 11 blocks of 256 / a pure power of two
    "tiny256": dict(hidden_size=512, intermediate_size=2816, num_hidden_layers=2,
                    num_attention_heads=4, num_key_value_heads=2),
    "tinypow2": dict(hidden_size=256, intermediate_size=1024, num_hidden_layers=2,
                     num_attention_heads=2, num_key_value_heads=1),
}


def llama_config(name_or_cfg, vocab_size=32000, **overrides):
    from transformers import LlamaConfig
    if not isinstance(name_or_cfg, str):
        return name_or_cfg
    kw = dict(LLAMA2_SHAPES[name_or_cfg], vocab_size=vocab_size, max_position_embeddings=4096,
              rms_norm_eps=1e-5, tie_word_embeddings=False)
    kw.update(overrides)
    return LlamaConfig(**kw)


E8P_GRID_RMS = 1.09375 ** 0.5 * 1.09375 ** 0.5  # E8P grid mean square per element = 1.09375


@torch.no_grad()
def randomize_quantlinear(layer: QuantLinear, gen: torch.Generator, weight_std: float = 0.02):
    """Random packed weights: codes uniform over the full index range, SU/SV random signs,
    Wscale chosen so that the effective weight std is `weight_std`."""
    dev = layer.Qidxs.device
    info = torch.iinfo(layer.Qidxs.dtype)
    q = torch.randint(info.min, info.max + 1, layer.Qidxs.shape, dtype=torch.int64, device=dev, generator=gen)
    layer.Qidxs.copy_(q.to(layer.Qidxs.dtype))
    for p in (layer.SU, layer.SV):
        s = torch.randint(0, 2, p.shape, device=dev, generator=gen).to(p.dtype) * 2 - 1
        p.data.copy_(s)
    rms = {"E8P12": 1.09375 ** 0.5, "E8P12RVQ4B": 1.09375 ** 0.5 * 1.04, "E8P12RVQ3B": 1.09375 ** 0.5 * 1.06,
           "D4": 1.2990, "HI": 4.61}[layer.codebook.id]
    layer.Wscale.fill_(weight_std / rms)
    if layer.bias is not None:
        layer.bias.zero_()
    for name in ("had_left", "had_right"):
        h = getattr(layer, name)
        if h is not None and layer.use_rand:
            k = h.shape[0]
            a = torch.randn(k, k, device=dev, generator=gen, dtype=torch.float32)
            qm, r = torch.linalg.qr(a)
            qm = qm * torch.sign(torch.diagonal(r)).unsqueeze(0)
            h.copy_(qm.to(h.dtype))


@torch.no_grad()
def make_random_quantized_llama(config="llama2-7b", codebook="E8P12", seed=0, device="cuda",
                                dtype=torch.float16, use_rand=True, per_channel=False,
                                layer_range: Optional[range] = None):
    """LlamaForCausalLM with random-init QuantLinear blocks, built directly on `device`.
    `layer_range` keeps only a slice of the decoder layers (pipeline stage); embeddings / final norm /
    lm_head are always created (they are small) so any stage can be first or last."""
    from transformers import LlamaForCausalLM
    cfg = llama_config(config)
    if layer_range is not None:
        import copy
        cfg = copy.deepcopy(cfg)
        cfg.num_hidden_layers = len(layer_range)
    from .quantizer import _materialize, init_empty_weights
    with init_empty_weights():           # parameters on meta, computed buffers (rotary inv_freq) real
        model = LlamaForCausalLM(cfg).to(dtype)
    # opt_resid_scale=None: the codebook's own default (1/3.45 for RVQ4B, BASELINE config 4); the quantizer's default of -1
    # (reference quantizer.py:69) is handed to the codebook as a literal scale
    quantizer = QuipQuantizer(codebook=codebook, use_rand=use_rand, per_channel=per_channel, inference=True,
                              ft_epochs=0, opt_resid_scale=None)
    model = quantizer.convert_model(model)
    _materialize(model, device, dtype)
    model = model.to(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    std = getattr(cfg, "initializer_range", 0.02)
    for name, mod in model.named_modules():
        if isinstance(mod, QuantLinear):
            randomize_quantlinear(mod, gen)
        elif isinstance(mod, nn.Embedding):
            mod.weight.normal_(0.0, std, generator=gen)
        elif isinstance(mod, nn.Linear):
            mod.weight.normal_(0.0, std, generator=gen)
            if mod.bias is not None:
                mod.bias.zero_()
        elif mod.__class__.__name__.endswith("RMSNorm"):
            mod.weight.fill_(1.0)
    apply_load_time_tricks(model, merge_suv=False)
    model.is_quantized = True
    model.eval()
    return model


def quantized_bytes(model) -> int:
    """Packed code bytes of all QuantLinears = the algorithmic HBM bytes of one bs=1 decode step."""
    return sum(m.Qidxs.numel() * m.Qidxs.element_size() for m in model.modules() if isinstance(m, QuantLinear))


class LlamaDecodeEngine:
    """bs=1 greedy decoding over (a slice of) a Llama model with QuantLinear blocks.

    step graph:  tok -> [embed] -> layers -> [norm -> lm_head -> argmax -> tok]   (one CUDA graph)
    """

    def __init__(self, model, max_cache_len: int = 512, first_stage: bool = True, last_stage: bool = True,
                 use_cuda_graph: bool = True, fused: bool = True, persistent: bool = True):
        self.model = model
        self.cfg = model.config
        self.layers = list(model.model.layers)
        self.first, self.last = first_stage, last_stage
        self.dev = next(model.parameters()).device
        cfg = self.cfg
        self.nh, self.nkv = cfg.num_attention_heads, cfg.num_key_value_heads
        self.hd = getattr(cfg, "head_dim", None) or cfg.hidden_size // cfg.num_attention_heads
        self.eps = cfg.rms_norm_eps
        self.max_len = max_cache_len
        dt = torch.float16
        L = len(self.layers)
        self.k_cache = torch.zeros(L, 1, self.nkv, max_cache_len, self.hd, dtype=dt, device=self.dev)
        self.v_cache = torch.zeros_like(self.k_cache)
        rp = getattr(cfg, "rope_parameters", None) or {}
        theta = rp.get("rope_theta", getattr(cfg, "rope_theta", 10000.0))
        inv = 1.0 / (theta ** (torch.arange(0, self.hd, 2, dtype=torch.float32, device=self.dev) / self.hd))
        t = torch.arange(max_cache_len, dtype=torch.float32, device=self.dev)
        fr = torch.outer(t, inv)
        emb = torch.cat((fr, fr), dim=-1)
        self.cos, self.sin = emb.cos().to(dt), emb.sin().to(dt)
        self.arange = torch.arange(max_cache_len, device=self.dev)
        # static step buffers
        self.tok = torch.zeros(1, 1, dtype=torch.long, device=self.dev)
        self.pos = torch.zeros(1, dtype=torch.long, device=self.dev)
        self.hidden_in = torch.zeros(1, 1, cfg.hidden_size, dtype=dt, device=self.dev)
        self.hidden_out = torch.zeros(1, 1, cfg.hidden_size, dtype=dt, device=self.dev)
        self.use_graph = use_cuda_graph
        self.graph = None
        self.launches_per_step = None
        self.fused = None
        self.persistent = None
        if fused:
            self._build_fused()
        if fused and persistent and self.fused is not None:
            self._build_persistent()
        self.tail = None
        if self.persistent is not None and self.last:
            self._build_tail()
        self._host_pos = 0          # host mirror of self.pos (cache-full check without a device sync)

    def _build_fused(self):
        """Per-layer launch groups for the fused decode step (5 launches of ours per layer + 1 prologue):
        [norm+q,k,v] -> [rope+append+attention] -> [o + residual] -> [norm+gate,up] -> [silu*up+down + residual]."""
        from .fused import LinearGroup
        try:
            groups = []
            for lyr in self.layers:
                at, mlp = lyr.self_attn, lyr.mlp
                if not all(isinstance(m, QuantLinear) for m in
                           (at.q_proj, at.k_proj, at.v_proj, at.o_proj, mlp.gate_proj, mlp.up_proj, mlp.down_proj)):
                    raise ValueError("not all block linears are QuantLinear")
                groups.append(dict(qkv=LinearGroup([at.q_proj, at.k_proj, at.v_proj]), o=LinearGroup([at.o_proj]),
                                   gu=LinearGroup([mlp.gate_proj, mlp.up_proj]), down=LinearGroup([mlp.down_proj])))
            if self.hd != 128:
                raise ValueError("attention kernel needs head_dim 128")
            self.fused = groups
            self.attn_out = torch.empty(1, self.nh * self.hd, dtype=torch.float16, device=self.dev)
        except ValueError:
            self.fused = None

    def _build_persistent(self):
        """Whole-step persistent kernel (decode_step.cu) when every block linear is E8P12 and the shapes qualify."""
        from .decode_step import PersistentDecodeStep
        try:
            self.persistent = PersistentDecodeStep(self.layers, self.k_cache, self.v_cache, self.cos, self.sin, self.pos,
                                                   self.nh, self.nkv, self.hd, self.cfg.hidden_size, self.eps)
            self.h_step_out = torch.empty(1, self.cfg.hidden_size, dtype=torch.float16, device=self.dev)
        except ValueError:
            self.persistent = None

    def _build_tail(self):
        """Final norm -> lm_head -> argmax -> position increment in one launch (lm_tail.cu) for fp16 bias-free lm_heads."""
        from .fused import LmTail
        head = self.model.lm_head
        try:
            if getattr(head, "bias", None) is not None:
                raise ValueError("lm_head has a bias")
            self.tail = LmTail(self.model.model.norm.weight, self.eps, head.weight)
        except ValueError:
            self.tail = None

    def _layer_fused(self, li, h):
        """h: fp16 [1, hidden]; one decode position (self.pos)."""
        from .fused import attn_decode
        lyr, g = self.layers[li], self.fused[li]
        q, k, v = g["qkv"](h, norm_w=lyr.input_layernorm.weight, eps=self.eps)
        attn_decode(q, k, v, self.k_cache[li], self.v_cache[li], self.cos, self.sin, self.pos, self.attn_out,
                    self.nh, self.nkv, self.hd)
        h = g["o"](self.attn_out, residual=h)[0]
        gate, up = g["gu"](h, norm_w=lyr.post_attention_layernorm.weight, eps=self.eps)
        return g["down"](up, gate=gate, residual=h)[0]

    # ---- building blocks --------------------------------------------------------------------
    def _rms(self, x, w):
        v = x.float()
        v = v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + self.eps)
        return w * v.to(x.dtype)

    @staticmethod
    def _rot(x):
        h = x.shape[-1] // 2
        return torch.cat((-x[..., h:], x[..., :h]), dim=-1)

    def _layer(self, li, h, pos, cos, sin, mask):
        lyr = self.layers[li]
        at = lyr.self_attn
        T = h.shape[1]
        x = self._rms(h, lyr.input_layernorm.weight)
        q = at.q_proj(x).view(1, T, self.nh, self.hd).transpose(1, 2)
        k = at.k_proj(x).view(1, T, self.nkv, self.hd).transpose(1, 2)
        v = at.v_proj(x).view(1, T, self.nkv, self.hd).transpose(1, 2)
        q = q * cos + self._rot(q) * sin
        k = k * cos + self._rot(k) * sin
        self.k_cache[li].index_copy_(2, pos, k)
        self.v_cache[li].index_copy_(2, pos, v)
        kk, vv = self.k_cache[li], self.v_cache[li]
        o = F.scaled_dot_product_attention(q, kk, vv, attn_mask=mask, enable_gqa=(self.nkv != self.nh))
        o = o.transpose(1, 2).reshape(1, T, self.nh * self.hd)
        h = h + at.o_proj(o)
        x = self._rms(h, lyr.post_attention_layernorm.weight)
        mlp = lyr.mlp
        h = h + mlp.down_proj(F.silu(mlp.gate_proj(x)) * mlp.up_proj(x))
        return h

    def _forward(self, tok_or_hidden, pos):
        """pos: long [T] absolute positions. Returns next-token ids [1,1] (last stage) or hidden."""
        h = self.model.model.embed_tokens(tok_or_hidden) if self.first else tok_or_hidden
        if self.fused is not None and h.shape[1] == 1 and pos is self.pos:
            h2 = h.view(1, -1)
            if self.persistent is not None:
                h2 = self.persistent(h2.contiguous(), self.h_step_out)
            else:
                for li in range(len(self.layers)):
                    h2 = self._layer_fused(li, h2)
            h = h2.view(1, 1, -1)
            if not self.last:
                return h
            h = self._rms(h, self.model.model.norm.weight)
            return self.model.lm_head(h).argmax(-1)
        cos = self.cos.index_select(0, pos)[None, None]
        sin = self.sin.index_select(0, pos)[None, None]
        mask = (self.arange[None, :] <= pos[:, None])[None, None]     # [1,1,T,max_len]
        for li in range(len(self.layers)):
            h = self._layer(li, h, pos, cos, sin, mask)
        if not self.last:
            return h
        h = self._rms(h[:, -1:], self.model.model.norm.weight)
        logits = self.model.lm_head(h)
        return logits.argmax(-1)

    # ---- public API -------------------------------------------------------------------------
    @torch.no_grad()
    def prefill(self, ids_or_hidden):
        T = ids_or_hidden.shape[1]
        assert T <= self.max_len
        pos = torch.arange(T, device=self.dev)
        out = self._forward(ids_or_hidden, pos)
        self.pos.fill_(T)
        self._host_pos = T
        if self.last:
            self.tok.copy_(out)
        return out

    @torch.no_grad()
    def _step_body(self):
        if self.tail is not None:        # [embedding] -> persistent decode step -> fused tail (token id + position)
            h = self.model.model.embed_tokens(self.tok).view(1, -1) if self.first else self.hidden_in.view(1, -1)
            self.persistent(h.contiguous(), self.h_step_out)
            self.tail(self.h_step_out, self.tok, pos=self.pos)
            return
        out = self._forward(self.tok if self.first else self.hidden_in, self.pos)
        if self.last:
            self.tok.copy_(out)
        else:
            self.hidden_out.copy_(out)
        self.pos.add_(1)

    @torch.no_grad()
    def capture(self):
        """Warm up and capture one decode step into a CUDA graph (state is restored afterwards)."""
        if not self.use_graph:
            return
        if self._host_pos + 3 > self.max_len:      # two warm-up steps and the captured one write at pos, pos+1, pos+2
            raise RuntimeError(f"KV cache too small to capture a decode step: position {self._host_pos}, "
                               f"max_cache_len {self.max_len}")
        saved = (self.tok.clone(), self.pos.clone(), self.k_cache.clone(), self.v_cache.clone())
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(2):
                self._step_body()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        self.pos.copy_(saved[1])
        from . import _native
        g = torch.cuda.CUDAGraph()
        lc0 = _native.launch_count()
        with torch.cuda.graph(g):
            self._step_body()
        self.launches_per_step = _native.launch_count() - lc0     # kernels of ours enqueued per replay
        self.graph = g
        self.tok.copy_(saved[0]); self.pos.copy_(saved[1])
        self.k_cache.copy_(saved[2]); self.v_cache.copy_(saved[3])

    @torch.no_grad()
    def capture_host_io(self):
        """A second CUDA graph for callers that keep the token on the host: [pinned host token -> device] -> the decode step ->
        [next token -> pinned host], i.e. the two 8-byte copies ride inside the replay instead of being two more API calls
        per step.  Returns the pinned (input, output) buffers; use `step_host()`."""
        assert self.first and self.last, "host token I/O belongs to an engine that holds the whole model"
        if getattr(self, "graph_io", None) is not None:
            return self.h_tok_in, self.h_tok_out
        if self.use_graph and self.graph is None:
            self.capture()
        self.h_tok_in = torch.zeros(1, 1, dtype=torch.long).pin_memory()
        self.h_tok_out = torch.zeros(1, 1, dtype=torch.long).pin_memory()
        self.h_tok_in.copy_(self.tok.cpu())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.tok.copy_(self.h_tok_in, non_blocking=True)
            self._step_body()
            self.h_tok_out.copy_(self.tok, non_blocking=True)
        self.graph_io = g
        return self.h_tok_in, self.h_tok_out

    @torch.no_grad()
    def step_host(self, token_id=None):
        """One decode step with host-resident token ids: feeds `token_id` (default: the previous output), returns the next
        token id as a Python int after a stream synchronisation."""
        if getattr(self, "graph_io", None) is None:
            self.capture_host_io()
        if self._host_pos >= self.max_len:
            raise RuntimeError(f"KV cache is full ({self.max_len} positions): allocate a larger max_cache_len")
        self._host_pos += 1
        if token_id is not None:
            self.h_tok_in[0, 0] = int(token_id)
        self.graph_io.replay()
        torch.cuda.current_stream(self.dev).synchronize()
        nxt = int(self.h_tok_out[0, 0])
        self.h_tok_in[0, 0] = nxt
        return nxt

    @torch.no_grad()
    def step(self):
        """One decode step: consumes self.tok / self.pos, leaves the next token in self.tok."""
        if self._host_pos >= self.max_len:          # the kernels clamp the position: refuse instead of overwriting the last slot
            raise RuntimeError(f"KV cache is full ({self.max_len} positions): allocate a larger max_cache_len")
        self._host_pos += 1
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_body()

    @torch.no_grad()
    def generate(self, input_ids, max_new_tokens):
        """Greedy decode; returns the generated ids [1, max_new_tokens] (device tensor)."""
        assert self.first and self.last
        if input_ids.shape[1] + max_new_tokens - 1 > self.max_len:
            raise ValueError(f"prompt ({input_ids.shape[1]}) + max_new_tokens ({max_new_tokens}) exceeds max_cache_len {self.max_len}")
        self.prefill(input_ids.to(self.dev))
        if self.use_graph and self.graph is None:
            self.capture()
        out = torch.empty(1, max_new_tokens, dtype=torch.long, device=self.dev)
        out[:, 0] = self.tok[:, 0]
        for i in range(1, max_new_tokens):
            self.step()
            out[:, i] = self.tok[:, 0]
        return out
