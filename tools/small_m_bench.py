"""Codebook mm ops at 1 <= M <= 32 on the same packed weights: this repository's routes (integer-dp4a GEMV per row for
M <= 3, the codebook-templated tcgen05 decode+GEMM from 4 rows on) against the reference's own small-M kernels K1 / K2 /
K3 (origin_order.cu:388-555, :337-385, :143-168) recompiled for sm_100a (oracle/_ref).  NL distinct weight matrices (> L2)
inside one CUDA graph; us per call.   Usage (GPU box): python tools/small_m_bench.py [M,M,...] [NxK,...]   (RVQ3B and HI included)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from quip_for_all_b200 import _native, codebook_id  # noqa: E402
from quip_for_all_b200.codebook.d4 import build_D4_CB  # noqa: E402
from umma_bench import graph_time  # noqa: E402


def main():
    import build_ref
    ref = build_ref.load_ref_module()
    dev = torch.device("cuda:0")
    Ms = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["1", "4", "8", "16", "31"])]
    shapes = [tuple(int(v) for v in s.split("x")) for s in
              (sys.argv[2].split(",") if len(sys.argv) > 2 else ["4096x4096", "11008x4096"])]
    grid = codebook_id["E8P12"](inference=True).to(dev).grid_packed_abs
    d4g = build_D4_CB().half().to(dev)
    from quip_for_all_b200.codebook.e8p12_rvq3 import get_e81bgrid, pack_e81b
    e81b = pack_e81b(get_e81bgrid()).to(dev)
    out = []
    for N, K in shapes:
        for cb in ("E8P12", "E8P12RVQ4B", "D4", "E8P12RVQ3B", "HI"):
            per = N * K // (2 if cb in ("E8P12RVQ4B", "HI") else 4) if cb != "E8P12RVQ3B" else N * K * 3 // 8
            NL = max(4, min(48, (256 << 20) // per + 1))
            if cb == "E8P12":
                qs = [torch.randint(-32768, 32768, (N, K // 8), device=dev, dtype=torch.int32).to(torch.int16) for _ in range(NL)]
                ours = lambda x, q: torch.ops.quip_lib.e8p_mm_origorder(x, q, grid)
                theirs = (lambda x, q: ref.e8p_mm_origorder(x, q, grid)) if ref else None
            elif cb == "E8P12RVQ4B":
                qs = [torch.randint(-2**31, 2**31, (N, K // 8), device=dev, dtype=torch.int64).to(torch.int32) for _ in range(NL)]
                ours = lambda x, q: torch.ops.quip_lib.e8prvq4_mm_origorder(x, q, grid, 1 / 3.45)
                theirs = (lambda x, q: ref.e8prvq4_mm_origorder(x, q, grid, 1 / 3.45)) if ref else None
            elif cb == "D4":
                qs = [torch.randint(0, 256, (N, K // 4), device=dev, dtype=torch.int32).to(torch.uint8) for _ in range(NL)]
                ours = lambda x, q: torch.ops.quip_lib.d4_mm_origorder(x, q, d4g)
                theirs = (lambda x, q: ref.d4_mm_origorder(x, q, d4g)) if ref else None
            elif cb == "E8P12RVQ3B":
                qs = [torch.randint(-2**31, 2**31, (N, 3 * K // 32), device=dev, dtype=torch.int64).to(torch.int32) for _ in range(NL)]
                ours = lambda x, q: torch.ops.quip_lib.e8prvq3_mm_origorder(x, q, grid, e81b, 1 / 2.04)
                theirs = (lambda x, q: ref.e8prvq3_mm_origorder(x, q, grid, e81b, 1 / 2.04)) if ref else None
            else:
                qs = [torch.randint(-2**31, 2**31, (N, K // 8), device=dev, dtype=torch.int64).to(torch.int32) for _ in range(NL)]
                ours = lambda x, q: torch.ops.quip_lib.hi_mm_origorder(x, q)
                theirs = (lambda x, q: ref.hi_mm_origorder(x, q)) if ref else None
            for M in Ms:
                x = torch.randn(M, K, device=dev, dtype=torch.float16)
                ours(x, qs[0])                      # workspace creation outside the capture
                rec = {"codebook": cb, "N": N, "K": K, "M": M, "layers": NL, "code_bytes": per}
                rec["ours_us"] = round(1000 * graph_time(lambda: [ours(x, q) for q in qs]) / NL, 2)
                if theirs is not None:
                    rec["reference_kernel_us"] = round(1000 * graph_time(lambda: [theirs(x, q) for q in qs]) / NL, 2)
                    rec["speedup"] = round(rec["reference_kernel_us"] / rec["ours_us"], 2)
                rec["ours_code_gbs"] = round(per / (rec["ours_us"] * 1e-6) / 1e9, 1)
                print(json.dumps(rec), flush=True)
                out.append(rec)
            del qs
            torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "small_m_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
