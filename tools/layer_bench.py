"""Per-shape QuantLinear timings on the GPU (not the headline bench; feeds DESIGN.md / profiles/).
For each Llama-2-7B linear shape, sweeps NL distinct layers (> L2) inside one CUDA graph and reports
us/call and GB/s of packed codes for: the fused op, its separate stages, and the reference's own kernel."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from quip_for_all_b200 import QuantLinear, _native, codebook_id  # noqa: E402
from quip_for_all_b200.modeling import randomize_quantlinear  # noqa: E402
from quip_for_all_b200.quantizer import apply_load_time_tricks  # noqa: E402


def graph_time(fn, reps=10):
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
    cb_name = sys.argv[1] if len(sys.argv) > 1 else "E8P12"
    Ms = [int(v) for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["1"])]
    shapes = [(4096, 4096), (4096, 11008), (11008, 4096), (8192, 8192), (8192, 28672), (28672, 8192)]
    if len(sys.argv) > 3:
        shapes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[3].split(",")]
    try:
        import build_ref
        ref = build_ref.load_ref_module()
    except Exception:
        ref = None
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    out = []
    for fin, fout in shapes:
        per = fin * fout // 4 if cb_name != "E8P12RVQ4B" else fin * fout // 2
        NL = max(4, min(48, (192 << 20) // per + 1))
        layers = []
        for i in range(NL):
            L = QuantLinear(fin, fout, codebook_id[cb_name](inference=True), bias=False).to(dev)
            randomize_quantlinear(L, gen)
            layers.append(L.eval())
        apply_load_time_tricks(torch.nn.ModuleList(layers))
        code_bytes = layers[0].Qidxs.numel() * layers[0].Qidxs.element_size()
        for M in Ms:
            x = torch.randn(M, fin, device=dev, dtype=torch.float16)
            rec = {"shape": f"{fin}x{fout}", "codebook": cb_name, "M": M, "layers": NL, "code_bytes": code_bytes}

            def sweep():
                for L in layers:
                    L(x)
            for name, fuse, mask, pdl in (("fused", 3, 7, 1), ("fused_nopdl", 3, 7, 0), ("unfused_all", 0, 7, 1), ("fused_pro_only", 1, 7, 1),
                                          ("gemv_only", 0, 2, 0), ("prologue_only", 0, 1, 0), ("epilogue_only", 0, 4, 0)):
                _native.set_option("fuse", fuse)
                _native.set_option("stage_mask", mask)
                _native.set_option("pdl", pdl)
                try:
                    ms = graph_time(sweep)
                finally:
                    _native.set_option("fuse", 3)
                    _native.set_option("stage_mask", 7)
                    _native.set_option("pdl", 1)
                rec[name + "_us"] = round(1000 * ms / NL, 3)
            rec["fused_gbs"] = round(code_bytes / (rec["fused_us"] * 1e-6) / 1e9, 1)
            rec["gemv_only_gbs"] = round(code_bytes / (rec["gemv_only_us"] * 1e-6) / 1e9, 1)
            if ref is not None and cb_name == "E8P12" and M < 32:
                xr = torch.randn(M, layers[0].q_in_features, device=dev, dtype=torch.float16)

                def ref_sweep():
                    for L in layers:
                        ref.e8p_mm_origorder(xr, L.Qidxs, L.codebook.grid_packed_abs)
                ms = graph_time(ref_sweep)
                rec["ref_mm_only_us"] = round(1000 * ms / NL, 3)
                rec["ref_mm_only_gbs"] = round(code_bytes / (rec["ref_mm_only_us"] * 1e-6) / 1e9, 1)
            print(json.dumps(rec), flush=True)
            out.append(rec)
        del layers
        torch.cuda.empty_cache()
    # grouped launches of a Llama-2-7B decoder layer (the engine's fused decode step)
    if cb_name == "E8P12" and len(sys.argv) <= 3:
        from quip_for_all_b200.fused import LinearGroup, attn_decode
        NL = 16

        def mk(fin, fout):
            L = QuantLinear(fin, fout, codebook_id[cb_name](inference=True), bias=False).to(dev)
            randomize_quantlinear(L, gen)
            return L.eval()
        qkv, og, gu, dn = [], [], [], []
        for i in range(NL):
            ls = [mk(4096, 4096) for _ in range(4)] + [mk(4096, 11008) for _ in range(2)] + [mk(11008, 4096)]
            apply_load_time_tricks(torch.nn.ModuleList(ls))
            qkv.append(LinearGroup(ls[0:3])); og.append(LinearGroup(ls[3:4])); gu.append(LinearGroup(ls[4:6]))
            dn.append(LinearGroup(ls[6:7]))
        h = torch.randn(1, 4096, device=dev, dtype=torch.float16)
        w = torch.ones(4096, device=dev, dtype=torch.float16)
        u = torch.randn(1, 11008, device=dev, dtype=torch.float16)
        kc = torch.zeros(1, 32, 512, 128, device=dev, dtype=torch.float16)
        vc = torch.zeros_like(kc)
        cs = torch.ones(512, 128, device=dev, dtype=torch.float16)
        pos = torch.tensor([384], device=dev)
        ao = torch.empty(1, 4096, device=dev, dtype=torch.float16)
        rec = {"shape": "llama2-7b decoder layer, grouped", "layers": NL}
        rec["qkv_group_us"] = round(1000 * graph_time(lambda: [g(h, norm_w=w, eps=1e-5) for g in qkv]) / NL, 3)
        rec["o_resid_us"] = round(1000 * graph_time(lambda: [g(h, residual=h) for g in og]) / NL, 3)
        rec["gate_up_group_us"] = round(1000 * graph_time(lambda: [g(h, norm_w=w, eps=1e-5) for g in gu]) / NL, 3)
        rec["down_silu_resid_us"] = round(1000 * graph_time(lambda: [g(u, gate=u, residual=h) for g in dn]) / NL, 3)
        rec["attn_ctx384_us"] = round(1000 * graph_time(lambda: [attn_decode(h, h, h, kc, vc, cs, cs, pos, ao, 32, 32, 128) for _ in range(NL)]) / NL, 3)
        rec["layer_total_us"] = round(sum(v for k, v in rec.items() if k.endswith("_us")), 3)

        def whole_layer():
            for i in range(NL):
                q, k, v = qkv[i](h, norm_w=w, eps=1e-5)
                attn_decode(q, k, v, kc, vc, cs, cs, pos, ao, 32, 32, 128)
                h2 = og[i](ao, residual=h)[0]
                g_, u_ = gu[i](h2, norm_w=w, eps=1e-5)
                dn[i](u_, gate=g_, residual=h2)
        rec["layer_chained_us"] = round(1000 * graph_time(whole_layer) / NL, 3)
        _native.set_option("pdl", 0)
        rec["layer_chained_nopdl_us"] = round(1000 * graph_time(whole_layer) / NL, 3)
        _native.set_option("pdl", 1)
        print(json.dumps(rec), flush=True)
        out.append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"layer_bench_{cb_name}.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
