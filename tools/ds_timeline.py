"""Stage timeline of the persistent decode-step kernel (clock64 stamps of CTA 0, first layer).
Usage (GPU box): python tools/ds_timeline.py [n_layers] [ctx] [cta] [option=value ...]   (options: quipb200_set_option)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quip_for_all_b200 import _native  # noqa: E402
from quip_for_all_b200.decode_step import _bind  # noqa: E402
from quip_for_all_b200.modeling import LlamaDecodeEngine, llama_config, make_random_quantized_llama  # noqa: E402

nl = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 384
cta = int(sys.argv[3]) if len(sys.argv) > 3 else 0      # cta | (layer + 1) << 16 selects the stamped layer
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    _native.set_option(k, int(v))
    print(f"option {k} = {v}")
dev = torch.device("cuda:0")
model = make_random_quantized_llama(llama_config("llama2-7b", num_hidden_layers=nl), "E8P12", seed=0, device=dev)
eng = LlamaDecodeEngine(model, max_cache_len=ctx + 64, use_cuda_graph=False)
assert eng.persistent is not None
ids = torch.randint(0, 32000, (1, ctx))
eng.prefill(ids.to(dev))
buf = torch.zeros(64, dtype=torch.int64, device=dev)
L = _bind()
for _ in range(5):
    eng.step()
torch.cuda.synchronize()
L.quipb200_decode_step_debug_cta(cta)
L.quipb200_decode_step_debug(buf.data_ptr())
eng.step()
torch.cuda.synchronize()
L.quipb200_decode_step_debug(None)
t = buf.cpu().tolist()
names = ["start"] + [f"stamp{i}" for i in range(1, 40)]
base = t[0]
prev = base
for i, v in enumerate(t):
    if v:
        print(f"  [{i:2d}] +{(v - prev) / 1965.0:8.2f} us   (t = {(v - base) / 1965.0:8.2f} us)")
        prev = v
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    eng.step()
e1.record()
torch.cuda.synchronize()
print(f"eager step (incl. embed/lm_head/launch gaps), {nl} layers: {e0.elapsed_time(e1) / 20 * 1000:.1f} us")
