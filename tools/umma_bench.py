"""tcgen05 decode+GEMM (umma_gemm.cu) vs the reference's M >= 32 route (decompress + cuBLAS GEMM) on the same
packed weights.  Sweeps NL distinct layers (> L2) inside one CUDA graph; reports us/call and TFLOP/s.
Usage (GPU box): python tools/umma_bench.py [M,M,...] [NxK,NxK,...]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quip_for_all_b200 import _native, codebook_id  # noqa: E402
import quip_for_all_b200  # noqa: E402,F401


def graph_time(fn, reps=5):
    with torch.no_grad():
        fn()
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(reps):
                g.replay()
            e1.record(s)
            e1.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda:0")
    Ms = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["32", "64", "128", "256"])]
    shapes = [tuple(int(v) for v in s.split("x")) for s in
              (sys.argv[2].split(",") if len(sys.argv) > 2 else ["4096x4096", "11008x4096", "4096x11008"])]
    cb = codebook_id["E8P12"](inference=True).to(dev)
    grid = cb.grid_packed_abs
    out = []
    for N, K in shapes:
        per = N * K // 4
        NL = max(4, min(48, (256 << 20) // per + 1))
        qs = [torch.randint(-32768, 32768, (N, K // 8), device=dev, dtype=torch.int32).to(torch.int16) for _ in range(NL)]
        for M in Ms:
            x = torch.randn(M, K, device=dev, dtype=torch.float16)
            rec = {"N": N, "K": K, "M": M, "layers": NL, "code_bytes": per, "gflop": 2.0 * M * N * K / 1e9}
            for name, flag in (("umma", 1 if M > 16 else 2), ("dense", 0)):
                _native.set_option("umma", flag)
                try:
                    ms = graph_time(lambda: [torch.ops.quip_lib.e8p_mm_origorder(x, q, grid) for q in qs])
                finally:
                    _native.set_option("umma", 2)
                us = 1000 * ms / NL
                rec[name + "_us"] = round(us, 2)
                rec[name + "_tflops"] = round(rec["gflop"] / us / 1e3, 1)
            rec["speedup"] = round(rec["dense_us"] / rec["umma_us"], 2)
            rec["umma_code_gbs"] = round(per / (rec["umma_us"] * 1e-6) / 1e9, 1)
            print(json.dumps(rec), flush=True)
            out.append(rec)
        del qs
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "umma_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
