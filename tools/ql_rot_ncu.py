"""A few unfused bs=1 forwards of the Llama-2-70B MLP shapes for an ncu capture of the 1-CTA rotation kernels (GPU box):
ncu --set full --clock-control none --import-source on -k regex:'ql_prologue|ql_epilogue' -s 4 -c 4 -o gpurun_out/ql_rot python tools/ql_rot_ncu.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quip_for_all_b200 import QuantLinear, _native, codebook_id  # noqa: E402
from quip_for_all_b200.modeling import randomize_quantlinear  # noqa: E402
from quip_for_all_b200.quantizer import apply_load_time_tricks  # noqa: E402

dev = torch.device("cuda:0")
gen = torch.Generator(device=dev)
gen.manual_seed(0)
shapes = [(28672, 8192), (8192, 28672)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[1].split(",")]
_native.set_option("fuse", 0)
for fin, fout in shapes:
    L = QuantLinear(fin, fout, codebook_id["E8P12"](inference=True), bias=False).to(dev)
    randomize_quantlinear(L, gen)
    L.eval()
    apply_load_time_tricks(torch.nn.ModuleList([L]))
    x = torch.randn(1, fin, device=dev, dtype=torch.float16)
    with torch.no_grad():
        for _ in range(3):
            y = L(x)
    torch.cuda.synchronize()
