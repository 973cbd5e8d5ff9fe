"""Where does the row-streaming rotation differ from the CTA-per-row kernel?  (GPU box; debugging aid)"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quip_for_all_b200 import _native  # noqa: E402
import quip_for_all_b200.register_lib  # noqa: E402,F401

dev = torch.device("cuda:0")
n, K = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2816, 11)
for M in [int(v) for v in sys.argv[3:]] or [3, 300, 1500]:
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, n, generator=g).half().to(dev)
    qm, _ = torch.linalg.qr(torch.randn(K, K, generator=g))
    Kp = (K + 15) // 16 * 16
    hk = torch.zeros(Kp, Kp, dtype=torch.float16, device=dev)
    hk[:K, :K] = qm.half().to(dev)
    scale = 0.37 / math.sqrt(n // K)
    _native.set_option("rot_pipe_rows", 1)
    y_p = torch.ops.quip_lib.rotate_fused(x, None, hk, None, None, n, K, n, scale)
    torch.cuda.synchronize()
    _native.set_option("rot_pipe_rows", 1 << 30)
    y_c = torch.ops.quip_lib.rotate_fused(x, None, hk, None, None, n, K, n, scale)
    torch.cuda.synchronize()
    bad = (y_p != y_c)
    rows = bad.any(dim=1).nonzero().flatten().tolist()
    print(f"M={M}: {len(rows)} bad rows of {M}; first {rows[:12]}")
    if rows:
        sms = 148
        grid = min(M, sms)
        its = sorted({r // grid for r in rows})
        print("   row-in-CTA indices (it) of bad rows:", its[:20])
        r = rows[0]
        blk = bad[r].view(K, 256).any(dim=1).nonzero().flatten().tolist()
        print(f"   row {r}: bad blocks {blk}; bad elements in first bad block:",
              bad[r].view(K, 256)[blk[0]].nonzero().flatten().tolist()[:24])
        print("   max abs diff", (y_p.float() - y_c.float()).abs().max().item(), "ref max", y_c.float().abs().max().item())
