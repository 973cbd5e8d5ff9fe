"""Arrival skew at the grid barriers of the persistent decode-step kernel: stamps of EVERY CTA (one step per CTA, the
kernel records one CTA at a time), reported per stage as the time each CTA spent in the stage's work and at its barrier.
The CTA with the shortest barrier wait is the one the others waited for.
Usage (GPU box): python tools/ds_skew.py [n_layers] [ctx]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quip_for_all_b200.decode_step import _bind  # noqa: E402
from quip_for_all_b200.modeling import LlamaDecodeEngine, llama_config, make_random_quantized_llama  # noqa: E402

nl = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 384
dev = torch.device("cuda:0")
model = make_random_quantized_llama(llama_config("llama2-7b", num_hidden_layers=nl), "E8P12", seed=0, device=dev)
eng = LlamaDecodeEngine(model, max_cache_len=ctx + 400, use_cuda_graph=False)
eng.prefill(torch.randint(0, 32000, (1, ctx)).to(dev))
L = _bind()
buf = torch.zeros(64, dtype=torch.int64, device=dev)
for _ in range(5):
    eng.step()
rows = []
for cta in range(148):
    buf.zero_()
    L.quipb200_decode_step_debug_cta(cta)
    L.quipb200_decode_step_debug(buf.data_ptr())
    eng.step()
    torch.cuda.synchronize()
    rows.append(buf.cpu().tolist())
L.quipb200_decode_step_debug(None)
us = lambda a, b: (b - a) / 1965.0 if a and b else float("nan")
# (work begin, work end = barrier arrival, barrier released) stamp triples per stage
stages = {"A": (1, 5, 6), "B": (6, 10, 11), "C": (11, 14, 15), "D": (15, 18, 19), "E": (19, 23, 24)}
for name, (b, e, r) in stages.items():
    work = [us(t[b], t[e]) for t in rows]
    wait = [us(t[e], t[r]) for t in rows]
    ok = [i for i in range(148) if work[i] == work[i] and wait[i] == wait[i]]
    if not ok:
        continue
    ws = sorted(ok, key=lambda i: wait[i])
    mean_work = sum(work[i] for i in ok) / len(ok)
    print(f"stage {name}: work mean {mean_work:5.2f} max {max(work[i] for i in ok):5.2f} min {min(work[i] for i in ok):5.2f} us | "
          f"barrier wait mean {sum(wait[i] for i in ok) / len(ok):5.2f} min {wait[ws[0]]:5.2f} us | last arrivers (cta: work / wait): "
          + ", ".join(f"{i}: {work[i]:.2f}/{wait[i]:.2f}" for i in ws[:6]))
    hist = {}
    for i in ok:
        hist.setdefault(round(work[i]), []).append(i)
    print("    work histogram (us -> #CTAs): " + ", ".join(f"{k}: {len(v)}" for k, v in sorted(hist.items())))
