"""BASELINE config 3 (batched / prefill path): QuantLinear.forward at M = bs*seq rows, i.e. the reference's
M >= 32 route -- hadamard op -> decompress_*_origorder -> dense fp16 GEMM (cuBLAS, a plain library GEMM) ->
hadamard op -- on the three Llama-2-7B linear shapes.  Reports per-stage time and TFLOP/s of the contraction."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quip_for_all_b200 import QuantLinear, codebook_id  # noqa: E402
from quip_for_all_b200.modeling import randomize_quantlinear  # noqa: E402
from quip_for_all_b200.quantizer import apply_load_time_tricks  # noqa: E402


def ev_time(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda:0")
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    cb = sys.argv[2] if len(sys.argv) > 2 else "E8P12"
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    out = []
    for fin, fout in ((4096, 4096), (4096, 11008), (11008, 4096)):
        L = QuantLinear(fin, fout, codebook_id[cb](inference=True), bias=False).to(dev)
        randomize_quantlinear(L, gen)
        apply_load_time_tricks(torch.nn.ModuleList([L]))
        L.eval()
        x = torch.randn(M, fin, device=dev, dtype=torch.float16)
        with torch.no_grad():
            t_all = ev_time(lambda: L(x))
            t_dec = ev_time(lambda: L.codebook.decompress_weight(L.Qidxs))
            W = L.codebook.decompress_weight(L.Qidxs)
            xp = torch.randn(M, L.q_in_features, device=dev, dtype=torch.float16)
            t_mm = ev_time(lambda: xp @ W.T)
            t_hin = ev_time(lambda: torch.ops.quip_lib.hadamard(xp.view(-1, L.q_in_features // L.K_left), 0.01))
            yo = torch.randn(M, L.q_out_features, device=dev, dtype=torch.float16)
            t_hout = ev_time(lambda: torch.ops.quip_lib.hadamard(yo.view(-1, L.q_out_features // L.K_right), 0.01))
            # the one-pass fused rotations the forward actually uses (rotate_batched.cu), both sides, with the n = 4096
            # side timed on the CTA-per-row kernel and on the warp-per-row kernel
            from quip_for_all_b200 import _native
            rot = {}
            xs_in = x
            for side, n, K, src in (("in", L.q_in_features, L.K_left, xs_in), ("out", L.q_out_features, L.K_right, yo)):
                hk = L._hk_padded("left" if side == "in" else "right") if K > 1 else None
                vec = torch.ones(src.shape[1], device=dev, dtype=torch.float16)
                call = lambda: torch.ops.quip_lib.rotate_fused(src, vec if side == "in" else None, hk,
                                                               vec if side == "out" else None, None, n, K, src.shape[1], 0.01)
                defaults = {k: _native.get_option(k) for k in ("rot_warp_rows", "rot_pipe_rows")}
                for name in ("cta_per_row", "default"):
                    for k, v in defaults.items():
                        _native.set_option(k, (1 << 30) if name == "cta_per_row" else v)
                    t = ev_time(call)
                    rot[f"{side}_{name}_ms"] = round(t, 3)
                    rot[f"{side}_{name}_gbs"] = round(2 * src.numel() * 2 / (t * 1e-3) / 1e9, 1)
                for k, v in defaults.items():
                    _native.set_option(k, v)
        flops = 2.0 * M * L.q_in_features * L.q_out_features
        rec = {"shape": f"{fin}x{fout}", "M": M, "codebook": cb, "forward_ms": round(t_all, 3),
               "decompress_ms": round(t_dec, 4), "gemm_ms": round(t_mm, 3), "hadamard_in_ms": round(t_hin, 3),
               "hadamard_out_ms": round(t_hout, 3),
               "gemm_tflops": round(flops / (t_mm * 1e-3) / 1e12, 1), "forward_tflops": round(flops / (t_all * 1e-3) / 1e12, 1),
               "decompress_gbs_written": round(L.q_in_features * L.q_out_features * 2 / (t_dec * 1e-3) / 1e9, 1),
               "hadamard_in_gbs": round(2 * xp.numel() * 2 / (t_hin * 1e-3) / 1e9, 1), "rotate_fused": rot}
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del x, xp, yo, W
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"prefill_bench_{cb}_M{M}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
