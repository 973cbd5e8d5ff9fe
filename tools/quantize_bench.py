"""Quantise-time nearest-codeword search: the fused kernel (csrc/nearest.cu) against the reference's expression
`(2 * X @ grid.T - grid_norm).argmax(-1)` (codebook/e8p12.py:125-128) evaluated by torch / cuBLAS on the same GPU, at the
row counts LDLQ feeds (one call per 8 columns, m = out_features), plus one whole-layer LDLQ.
Usage (GPU box): python tools/quantize_bench.py [out.json]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quip_for_all_b200 import codebook_id  # noqa: E402
from quip_for_all_b200.ldlq import ldlq, proxy_loss  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    dev = torch.device("cuda:0")
    res = {"search_us": [], "ldlq": []}
    for name in ("E8P12", "E8P12RVQ4B"):
        cb = codebook_id[name](inference=False).to(dev)
        for m in (1024, 4096, 11008, 28672):
            x = torch.randn(m, 8, device=dev) * 1.03
            ours = timed(lambda: cb.quantize(x), 20)
            if name == "E8P12":
                ref = timed(lambda: cb.round(x, cb.grid, cb.grid_norm), 5)
            else:
                def two():
                    v0, i0 = cb.round(x, cb.grid, cb.grid_norm)
                    r = (x - v0) / cb.opt_resid_scale
                    v1, i1 = cb.round(r, cb.grid, cb.grid_norm)
                    return v0 + v1 * cb.opt_resid_scale, (i0 << 16) + i1
                ref = timed(two, 5)
            pairs = m * 65536 * (2 if name != "E8P12" else 1)
            res["search_us"].append({"codebook": name, "m": m, "fused_us": round(ours, 1), "torch_expr_us": round(ref, 1),
                                     "speedup": round(ref / ours, 2), "fused_Gpairs_s": round(pairs / ours * 1e-3, 1)})
            print(res["search_us"][-1], flush=True)
    # whole-layer LDLQ, 4096 x 4096, E8P12 (512 sequential rounding calls of 4096 rows)
    cb = codebook_id["E8P12"](inference=False).to(dev)
    n = 4096
    A = torch.randn(n, 2 * n, device=dev)
    H = A @ A.T / (2 * n)
    H /= torch.diag(H).mean()
    H[torch.arange(n), torch.arange(n)] += 0.01
    L = torch.linalg.cholesky(H)
    W = torch.randn(n, n, device=dev) * 1.03

    class TorchCb:                      # the same codebook object with the reference's rounding expression
        codesz, idx_dtype = cb.codesz, cb.idx_dtype

        @staticmethod
        def quantize(X, return_idx=True):
            return cb.round(X, cb.grid, cb.grid_norm)

    for tag, c in (("fused", cb), ("torch_expr", TorchCb)):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hat, Q = ldlq(W, H, L, c, 0)
        e1.record()
        torch.cuda.synchronize()
        res["ldlq"].append({"shape": [n, n], "rounding": tag, "ms": round(e0.elapsed_time(e1), 1),
                            "proxy_loss": round(proxy_loss(W, hat, H), 5)})
        print(res["ldlq"][-1], flush=True)
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
