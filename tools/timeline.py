"""Per-phase timeline of the fused GEMV kernel (clock64 stamps written by CTA thread 0).
Usage (GPU box): python tools/timeline.py [fin fout]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quip_for_all_b200 import QuantLinear, _native, codebook_id  # noqa: E402
from quip_for_all_b200.modeling import randomize_quantlinear  # noqa: E402
from quip_for_all_b200.quantizer import apply_load_time_tricks  # noqa: E402

fin, fout = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 4096)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    _native.set_option(k, int(v))
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev)
gen.manual_seed(0)
layers = []
for i in range(40):
    L = QuantLinear(fin, fout, codebook_id["E8P12"](inference=True), bias=False).to(dev)
    randomize_quantlinear(L, gen)
    layers.append(L.eval())
apply_load_time_tricks(torch.nn.ModuleList(layers))
x = torch.randn(1, fin, device=dev, dtype=torch.float16)
buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
names = ["entry", "p0 issued", "pdl_wait", "phase1 done", "x regs", "gemv done", "acc stored", "ticket", "epilogue",
         "pro: loaded+sts", "pro: fwht", "pro: absmax", "-", "epi: loaded+sts", "epi: rotated", "-"]
with torch.no_grad():
    for L in layers:           # warm, cold-L2 for the last ones
        L(x)
    torch.cuda.synchronize()
    _native.lib().quipb200_debug_timeline(buf.data_ptr())
    layers[-1](x)
    torch.cuda.synchronize()
    _native.lib().quipb200_debug_timeline(None)
t = buf.view(148, 16).cpu()
t = t[t[:, 0] > 0]
d = (t[:, 1:16] - t[:, 0:1]).float() / 1965.0    # us at 1965 MHz (clock64 is per-SM; deltas only)
print(f"{fin}x{fout}: {t.shape[0]} CTAs; us since CTA entry (median / max over CTAs)")
for i, n in enumerate(names[1:]):
    col = d[:, i]
    valid = t[:, i + 1] > 0
    if valid.any():
        c = col[valid]
        print(f"  {n:12s} median {c.median():7.2f}  max {c.max():7.2f}  (n={int(valid.sum())})")
