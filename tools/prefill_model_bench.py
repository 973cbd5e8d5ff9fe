"""BASELINE config 3: Llama-2-7B E8P12, bs=32 x seq=2048 prefill (65 536 rows) through the unmodified HF forward with
QuantLinear blocks (batched path: fused rotations + decompress + library GEMM).  Reports tokens/s and the share of the
linears.  Usage (GPU box): python tools/prefill_model_bench.py [bs] [seq] [n_layers]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quip_for_all_b200 import QuantLinear  # noqa: E402
from quip_for_all_b200.modeling import llama_config, make_random_quantized_llama  # noqa: E402

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
seq = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
nl = int(sys.argv[3]) if len(sys.argv) > 3 else 32
dev = torch.device("cuda:0")
cfg = llama_config("llama2-7b", num_hidden_layers=nl)
model = make_random_quantized_llama(cfg, "E8P12", seed=0, device=dev)
ids = torch.randint(0, 32000, (bs, seq), generator=torch.Generator().manual_seed(0)).to(dev)
lin_ms = [0.0]


def timed(mod):
    orig = mod.forward

    def f(x):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = orig(x)
        e1.record()
        evs.append((e0, e1))
        return y
    return f


with torch.no_grad():
    for _ in range(2):
        h = model.model(ids).last_hidden_state          # decoder stack only (the lm_head of a prefill sees 1 row per sequence)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        h = model.model(ids).last_hidden_state
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    evs = []
    for m in model.modules():
        if isinstance(m, QuantLinear):
            m.forward = timed(m)
    model.model(ids)
    torch.cuda.synchronize()
    lin = sum(a.elapsed_time(b) for a, b in evs)
rec = {"config": f"llama2-7b E8P12 prefill bs={bs} seq={seq} layers={nl}", "ms": round(ms, 2),
       "tokens_per_s": round(bs * seq / (ms * 1e-3), 1), "quantlinear_ms": round(lin, 2),
       "quantlinear_share": round(lin / ms, 3),
       "gemm_tflops_equiv": round(2.0 * bs * seq * sum(m.in_features * m.out_features for m in model.modules()
                                                        if isinstance(m, QuantLinear)) / (ms * 1e-3) / 1e12, 1)}
print(json.dumps(rec))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rec, open(os.path.join(ROOT, "gpurun_out", "prefill_model_bench.json"), "w"), indent=1)
