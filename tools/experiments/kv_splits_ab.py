"""A/B of the number of KV splits per head in the attention stage of the persistent decode step (BASELINE config 2,
context 128 -> ~400).  Usage (GPU box): python tools/experiments/kv_splits_ab.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from quip_for_all_b200.decode_step import _bind  # noqa: E402
from quip_for_all_b200.modeling import LlamaDecodeEngine, make_random_quantized_llama  # noqa: E402

dev = torch.device("cuda:0")
model = make_random_quantized_llama("llama2-7b", "E8P12", seed=0, device=dev)
ids = torch.randint(0, 32000, (1, 128), generator=torch.Generator().manual_seed(0)).to(dev)
for rep in range(2):
    for S in (4, 3, 2, 1):
        _bind().quipb200_decode_step_set_splits(S)
        eng = LlamaDecodeEngine(model, max_cache_len=688)
        eng.prefill(ids)
        eng.capture()
        for _ in range(16):
            eng.step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(256):
            eng.step()
        e1.record()
        e1.synchronize()
        print(f"splits {S}: {256 / (e0.elapsed_time(e1) * 1e-3):.1f} tok/s", flush=True)
        del eng
_bind().quipb200_decode_step_set_splits(0)
