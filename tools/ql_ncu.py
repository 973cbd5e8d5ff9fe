"""A few fused bs=1 forwards of one QuantLinear shape for an ncu capture of the per-linear kernel (GPU box):
ncu --set full --clock-control none --import-source on -k regex:ql_gemv -s 6 -c 1 -o gpurun_out/ql_gemv python tools/ql_ncu.py 4096x4096"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quip_for_all_b200 import QuantLinear, codebook_id  # noqa: E402
from quip_for_all_b200.modeling import randomize_quantlinear  # noqa: E402
from quip_for_all_b200.quantizer import apply_load_time_tricks  # noqa: E402

dev = torch.device("cuda:0")
gen = torch.Generator(device=dev)
gen.manual_seed(0)
fin, fout = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "4096x4096").split("x"))
layers = []
for _ in range(40):          # > L2 of distinct packed weights: every launch streams its codes from HBM
    L = QuantLinear(fin, fout, codebook_id["E8P12"](inference=True), bias=False).to(dev)
    randomize_quantlinear(L, gen)
    L.eval()
    layers.append(L)
apply_load_time_tricks(torch.nn.ModuleList(layers))
x = torch.randn(1, fin, device=dev, dtype=torch.float16)
with torch.no_grad():
    for L in layers:
        y = L(x)
torch.cuda.synchronize()
