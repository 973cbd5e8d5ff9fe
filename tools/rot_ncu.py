"""Two launches for an ncu capture of the many-rows rotation kernels at M = 65 536 (GPU box):
ncu --set full --clock-control none -k regex:'rot4096w|rotblk_pipe' -c 2 -o gpurun_out/rot_full python tools/rot_ncu.py"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quip_for_all_b200.register_lib  # noqa: E402,F401

dev = torch.device("cuda:0")
M = 65536
g = torch.Generator().manual_seed(0)
for n, K in ((4096, 1), (11008, 43)):
    x = torch.randn(M, n, device=dev, dtype=torch.float16)
    vec = (1 + 0.1 * torch.randn(n, generator=g)).half().to(dev)
    hk = None
    if K > 1:
        qm, _ = torch.linalg.qr(torch.randn(K, K, generator=g))
        hk = torch.zeros(48, 48, dtype=torch.float16, device=dev)
        hk[:K, :K] = qm.half().to(dev)
    for _ in range(2):      # the second launch of each shape is the warm one (-c picks by regex order)
        y = torch.ops.quip_lib.rotate_fused(x, vec, hk, None, None, n, K, n, 0.37 / math.sqrt(n // K))
    torch.cuda.synchronize()
    del x, y
