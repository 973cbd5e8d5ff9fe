#!/usr/bin/env python
"""bench.py -- headline benchmark: Llama-2-7B E8P12 2-bit, bs=1 greedy decode tokens/s on B200
(BASELINE.json metric, configs[1]) + QuantLinear GEMV GB/s against the HBM roofline.

    python bench.py --gpus N --steps K --warmup W            (ours)
    python bench.py --impl reference --gpus N --steps K ...  (reference arm: CPU port of the path)

A "step" is one decode token = one pass of the hot path (224 QuantLinear GEMVs) over the model.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how every field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decode tokens/s bs=1 Llama-2-7B E8P12; QuipLinear GEMV GB/s vs HBM roofline"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default=None,
                    help="default: llama2-7b on one GPU (BASELINE config 2), llama2-70b for the multi-GPU layer pipeline (config 5)")
    ap.add_argument("--no-secondary", dest="no_secondary", action="store_true",
                    help="N > 1: skip the additional 7B pipeline measurement")
    ap.add_argument("--codebook", default="E8P12")
    ap.add_argument("--prompt-len", type=int, default=128)
    ap.add_argument("--cache-len", type=int, default=0, help="KV cache length (0: prompt + steps + warmup + 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-bench", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--no-hf-dropin", action="store_true")
    ap.add_argument("--no-70b", action="store_true", help="N = 1: skip the single-GPU Llama-2-70B leg (the base of the N > 1 scaling)")
    ap.add_argument("--handoff", default="peer", choices=["peer", "nccl"],
                    help="N > 1: stage hand-off through NVLink peer memory (device-side store + flag) or host-issued NCCL p2p")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--engine", default="auto", choices=["auto", "persistent", "grouped"],
                    help="decode engine: one persistent whole-step kernel, or one launch per linear group")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle's C port of the path, on the host cores
# --------------------------------------------------------------------------------------------------
LLAMA_LINEARS = {
    "llama2-7b": (32, [(4096, 4096)] * 4 + [(4096, 11008)] * 2 + [(11008, 4096)]),
    "llama2-70b": (80, [(8192, 8192), (8192, 1024), (8192, 1024), (8192, 8192), (8192, 28672), (8192, 28672),
                        (28672, 8192)]),
    "tiny": (2, [(256, 256)] * 4 + [(256, 704)] * 2 + [(704, 256)]),
}


def _host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class CpuTokenPass:
    """The QuantLinear forwards of ONE decode token (bs=1, decode-every-call: 7 per decoder layer, every layer with its own
    packed weights) through oracle/quip_oracle.c on all host threads.  Attention / norms / lm_head are not part of the
    reference's quantised path and are not timed (they are < 2 % of the CPU time of a token).
    `n_run` <= n_layers distinct layers are allocated and run per pass; tokens/s = 1 / (t_pass * n_layers / n_run)."""

    def __init__(self, model_name, n_run=None):
        # torchrun exports OMP_NUM_THREADS=1: the CPU arm is meant to use the host, so the thread count is set explicitly
        os.environ["OMP_NUM_THREADS"] = str(_host_threads())
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import numpy as np
        import quip_oracle as qo
        import quip_oracle_c as qc
        self.np, self.qo, self.qc = np, qo, qc
        self.n_layers, shapes = LLAMA_LINEARS[model_name]
        self.n_run = self.n_layers if n_run is None else max(1, min(n_run, self.n_layers))
        rng = np.random.default_rng(0)
        self.tab = qo.e8p_abs_table()
        self.layers = []
        orth = lambda k: np.linalg.qr(rng.standard_normal((k, k)))[0].astype(np.float32) if k > 1 else None
        base = []
        for fin, fout in shapes:
            Kl, q_in, _ = qo.hadK_shape(fin, True)
            Kr, q_out, _ = qo.hadK_shape(fout, True)
            base.append(dict(fin=fin, fout=fout, q_in=q_in, q_out=q_out, Kl=Kl, Kr=Kr, hl=orth(Kl), hr=orth(Kr),
                             q=rng.integers(-32768, 32768, (q_out, q_in // 8)).astype(np.int16),
                             SU=np.sign(rng.standard_normal(fin)).astype(np.float32),
                             SV=np.sign(rng.standard_normal(fout)).astype(np.float32),
                             x=rng.standard_normal((1, fin)).astype(np.float16)))
        for li in range(self.n_run):      # distinct weight arrays per layer (rolled copies: same statistics, no cache reuse)
            self.layers.append([dict(L, q=np.roll(L["q"], li + 1, axis=0).copy() if li else L["q"]) for L in base])
        self.cores = qc.num_threads()

    def run_once(self):
        t0 = time.perf_counter()
        for lin in self.layers:
            for L in lin:
                self.qc.quantlinear_forward_e8p(L["x"], L["q"], self.tab, L["fin"], L["fout"], L["q_in"], L["q_out"],
                                                SU=L["SU"], SV=L["SV"], wscale_float=0.0183,
                                                had_left=L["hl"], K_left=L["Kl"], had_right=L["hr"], K_right=L["Kr"])
        return time.perf_counter() - t0

    def tok_s(self, t_pass):
        return 1.0 / (t_pass * self.n_layers / self.n_run)

    def describe(self):
        what = (f"all {self.n_layers} decoder layers" if self.n_run == self.n_layers else
                f"{self.n_run} of {self.n_layers} decoder layers (distinct weights; tokens/s scaled by {self.n_layers}/{self.n_run})")
        return (f"one decode token = the 7 QuantLinear forwards (decode-every-call, bs=1, E8P12) of {what} via "
                f"oracle/quip_oracle.c, OpenMP, {self.cores} threads")


def _cpu_layers_for_budget(model_name, steps, warmup, budget_s=150.0):
    """How many decoder layers a pass may run so that (steps + warmup) passes fit the time budget: probed with a
    two-layer pass on this host."""
    probe = CpuTokenPass(model_name, n_run=2)
    probe.run_once()
    t2 = min(probe.run_once() for _ in range(2))
    per_layer = t2 / 2
    n_layers = probe.n_layers
    fit = int(budget_s / max(1e-9, per_layer * max(1, steps + warmup)))
    return max(1, min(n_layers, fit)), per_layer


def run_reference_arm(a):
    """`--impl reference`: the reference has no CPU implementation of its own ops (register_lib.py: CUDA-only), so the
    CPU statement of the path is the oracle port; timed on all host threads.  Every step is a whole-token pass when
    `steps + warmup` of them fit the time budget on this host, else the largest whole number of layers that does (stated
    in `sample`); `steps` and `warmup` are used as given."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_run, _ = _cpu_layers_for_budget(a.model, a.steps, a.warmup)
    s = CpuTokenPass(a.model, n_run=n_run)
    for _ in range(a.warmup):
        s.run_once()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        s.run_once()
    t_pass = (time.perf_counter() - t0) / a.steps
    val = s.tok_s(t_pass)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "tokens/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000.0 * t_pass * s.n_layers / s.n_run,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int8xint16 fixed point (ours) / fp16->fp32 (this port)", "data": "synthetic",
        "config": {"workload": f"{a.model} {a.codebook} bs=1 greedy decode, random-init packed weights "
                               "(CPU port of the QuantLinear path on the host cores; see `sampled`)",
                   "sampled": s.describe(), "timed_s": t_pass * a.steps},
        "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": s.cores, "kind": "port", "sample": s.describe()},
        "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# ours
# --------------------------------------------------------------------------------------------------
def kernel_bench(model, torch, peak, peak_src, reps=5):
    """GEMV-only sweep over every QuantLinear of the model (distinct weights, 1.6 GB >> L2), timed with
    CUDA events on the launching stream; plus the same sweep for prologue / epilogue and the full op."""
    from quip_for_all_b200 import QuantLinear, _native
    layers = [m for m in model.modules() if isinstance(m, QuantLinear)]
    dev = layers[0].Qidxs.device
    xs = {}
    for L in layers:
        if L.in_features not in xs:
            xs[L.in_features] = torch.randn(1, L.in_features, device=dev, dtype=torch.float16)
    bytes_total = sum(L.Qidxs.numel() * L.Qidxs.element_size() for L in layers)
    out = {}
    for name, fuse, mask in (("fused_op", 3, 7), ("gemv_only", 0, 2), ("prologue_only", 0, 1),
                             ("epilogue_only", 0, 4)):
        _native.set_option("fuse", fuse)
        _native.set_option("stage_mask", mask)
        try:
            with torch.no_grad():
                for L in layers[:8]:
                    L(xs[L.in_features])
                torch.cuda.synchronize()
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        for L in layers:
                            L(xs[L.in_features])
                    g.replay()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(s)
                    for _ in range(reps):
                        g.replay()
                    e1.record(s)
                    e1.synchronize()
                ms = e0.elapsed_time(e1) / reps
        finally:
            _native.set_option("stage_mask", 7)
            _native.set_option("fuse", 3)
        out[name] = {"us_per_launch_avg": 1000.0 * ms / len(layers), "launches": len(layers), "ms_per_sweep": ms}
    us = out["fused_op"]["us_per_launch_avg"]
    achieved = bytes_total / len(layers) / (us * 1e-6) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "kernel": "ql_gemv_kernel<E8P12> (fused: x rotation + GEMV + last-CTA output rotation; "
                                       "down_proj adds a 1-CTA prologue launch, included in the time)",
            "bytes_per_launch": bytes_total / len(layers), "us_per_launch": us, "peak_source": peak_src,
            "how": f"CUDA events around a graph of {len(layers)} back-to-back fused QuantLinear calls over all "
                   f"QuantLinears of the model ({bytes_total/1e9:.3f} GB of distinct codes, >> L2), {reps} replays"}
    return roof, out


def decode_step_kernel_bench(eng, model, torch, code_bytes, peak, peak_src, reps=20):
    """The dominant kernel of the headline engine: ONE launch of the persistent decode-step kernel processes the
    packed codes of every QuantLinear of the model (1.6 GB >> L2).  Timed alone with CUDA events on the launching
    stream; `traffic` is the ncu dram__bytes figure of the same kernel when a capture has been committed."""
    with torch.no_grad():
        h = model.model.embed_tokens(eng.tok).view(1, -1).contiguous()
        for _ in range(3):
            eng.persistent(h, eng.h_step_out)
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):          # a graph holding ONLY this kernel: replayed like the decode step replays it
                eng.persistent(h, eng.h_step_out)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(reps):
                g.replay()
            e1.record(s)
            e1.synchronize()
    us = 1000.0 * e0.elapsed_time(e1) / reps
    achieved = code_bytes / (us * 1e-6) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("decode_step_kernel", {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic,
            "kernel": "decode_step_kernel (persistent: all decoder layers of one bs=1 step; rotations, attention and "
                      "GEMVs of the 224 QuantLinears)",
            "bytes_per_launch": code_bytes, "us_per_launch": us, "peak_source": peak_src,
            "how": f"CUDA events around {reps} back-to-back graph replays of the kernel alone (same position, "
                   f"{code_bytes/1e9:.3f} GB of distinct packed codes per launch, >> L2)"}


def ref_cuda_bench(model, torch, reps=3):
    """Informational: the reference's own kernels (quip_cuda, recompiled for sm_100a, oracle/_ref) on the
    same packed weights: e8p_mm_origorder at M=1 swept over every layer ("kernel to beat")."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import build_ref
        ref = build_ref.load_ref_module()
    except Exception as e:
        return {"unavailable": str(e)[:200]}
    if ref is None:
        return {"unavailable": "oracle/_ref/quiptools_cuda.so not present"}
    from quip_for_all_b200 import QuantLinear
    layers = [m for m in model.modules() if isinstance(m, QuantLinear) and m.codebook.id == "E8P12"]
    if not layers:
        return {"unavailable": "no E8P12 layers"}
    dev = layers[0].Qidxs.device
    xs = {q: torch.randn(1, q, device=dev, dtype=torch.float16) for q in {L.q_in_features for L in layers}}
    bytes_total = sum(L.Qidxs.numel() * 2 for L in layers)
    with torch.no_grad():
        for L in layers[:4]:
            ref.e8p_mm_origorder(xs[L.q_in_features], L.Qidxs, L.codebook.grid_packed_abs)
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for L in layers:
                    ref.e8p_mm_origorder(xs[L.q_in_features], L.Qidxs, L.codebook.grid_packed_abs)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(reps):
                g.replay()
            e1.record(s)
            e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"kernel": "tinygemm_m16n8k16_chunk_kernel<BLayout_E8> via quiptools_cuda.e8p_mm_origorder (M=1, mm ONLY -- no "
                      "Hadamards / scalings --, CUDA-graph replay)",
            "us_per_launch_avg": 1000.0 * ms / len(layers), "gbs": bytes_total / (ms * 1e-3) / 1e9,
            "launches": len(layers)}


def hf_dropin_bench(model, torch, prompt, steps, warmup):
    """The API the north_star names: UNMODIFIED HF LlamaForCausalLM.forward over drop-in QuantLinear modules, static KV
    cache, one token per forward, the step captured in a CUDA graph (what the reference's example_generate.py:29-56,
    62-70 does with torch.compile(mode="reduce-overhead")).  Greedy, bs=1.  Returns tokens/s (CUDA events)."""
    try:
        from transformers import StaticCache
        dev = prompt.device
        T = prompt.shape[1]
        cache = StaticCache(config=model.config, max_cache_len=T + 2 * (steps + warmup) + 16)
        with torch.no_grad():
            out = model(input_ids=prompt, past_key_values=cache, use_cache=True,
                        cache_position=torch.arange(T, device=dev))
            tok = out.logits[:, -1:].argmax(-1)                      # static buffers of the captured step
            pos = torch.full((1,), T, dtype=torch.long, device=dev)

            def step():
                logits = model(input_ids=tok, past_key_values=cache, use_cache=True, cache_position=pos).logits
                tok.copy_(logits[:, -1:].argmax(-1))
                pos.add_(1)

            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3):
                    step()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step()
            for _ in range(warmup):
                g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                g.replay()
            e1.record()
            e1.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": 1000.0 / ms, "unit": "tokens/s", "ms_per_step": ms, "steps": steps,
                "how": "transformers LlamaForCausalLM.forward (unmodified) with quip_for_all_b200.QuantLinear modules, "
                       "StaticCache, one CUDA graph per decode step (torch.cuda.graph), greedy, bs=1"}
    except Exception as e:       # informational leg: never fail the headline line
        return {"unavailable": f"{type(e).__name__}: {str(e)[:300]}"}


def single_gpu_70b(a, torch, dev, peak, model_name=None):
    """BASELINE config 5 on ONE GPU (17.1 GB of packed codes fit): the N = 1 point of the 70B layer-pipeline scaling that
    `--gpus N` (N > 1) measures.  Same engine as a pipeline stage (grouped launches: the persistent kernel does not cover
    hidden 8192 / 7 x 4096 blocks), CUDA-graph replay, CUDA events."""
    try:
        from quip_for_all_b200.modeling import LlamaDecodeEngine, make_random_quantized_llama, quantized_bytes
        steps, warm = min(a.steps, 48), 4
        model_name = model_name or "llama2-70b"
        model = make_random_quantized_llama(model_name, a.codebook, seed=0, device=dev)
        code_bytes = quantized_bytes(model)
        eng = LlamaDecodeEngine(model, max_cache_len=a.prompt_len + 2 * (steps + warm) + 16, use_cuda_graph=not a.no_graph)
        g = torch.Generator().manual_seed(0)
        eng.prefill(torch.randint(0, model.config.vocab_size, (1, a.prompt_len), generator=g).to(dev))
        eng.capture()
        for _ in range(warm):
            eng.step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            eng.step()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out = {"workload": f"{model_name} {a.codebook} bs=1 greedy decode on one GPU, random-init packed weights",
               "value": 1000.0 / ms, "unit": "tokens/s", "ms_per_step": ms, "steps": steps,
               "engine": "persistent whole-step kernel" if eng.persistent is not None else "grouped launches (6 per decoder layer)",
               "packed_code_bytes_per_token": code_bytes, "frac_of_hbm_roofline": (1000.0 / ms) * code_bytes / (peak * 1e9)}
        del eng, model
        torch.cuda.empty_cache()
        return out
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {str(e)[:300]}"}


def run_ours(a):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from quip_for_all_b200.parallel import run_pipeline_bench
        return run_pipeline_bench(a, METRIC, ClockSampler, peak_gbs=measured_peaks()[0],
                                  single_gpu_fn=None if a.no_70b else single_gpu_70b)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from quip_for_all_b200 import _native
    from quip_for_all_b200.modeling import LlamaDecodeEngine, make_random_quantized_llama, quantized_bytes
    _native.lib()
    peak, peak_src = measured_peaks()

    t0 = time.time()
    model = make_random_quantized_llama(a.model, a.codebook, seed=0, device=dev)
    code_bytes = quantized_bytes(model)
    lm_head_bytes = model.lm_head.weight.numel() * 2
    cache_len = a.cache_len or (a.prompt_len + a.steps * 2 + a.warmup * 2 + 16)
    eng = LlamaDecodeEngine(model, max_cache_len=cache_len, use_cuda_graph=not a.no_graph,
                            persistent=(a.engine != "grouped"))
    if a.engine == "persistent" and eng.persistent is None:
        raise SystemExit("--engine persistent: model not covered by the persistent decode-step kernel")
    g = torch.Generator().manual_seed(0)
    prompt = torch.randint(0, model.config.vocab_size, (1, a.prompt_len), generator=g)
    pinned_in = prompt.clone().pin_memory()
    eng.prefill(pinned_in.to(dev, non_blocking=True))
    lc0 = _native.launch_count()
    eng.capture()
    launches_per_step = (_native.launch_count() - lc0) // 3 if not a.no_graph else None  # 2 warm-up + 1 captured
    torch.cuda.synchronize()
    build_s = time.time() - t0

    # ---- device-resident timing: K graph replays, CUDA events, clocks sampled alongside --------------
    for _ in range(a.warmup):
        eng.step()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()          # ncu --profile-from-start off captures only the timed region
    e0.record()
    for _ in range(a.steps):
        eng.step()
    e1.record()
    e1.synchronize()
    torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    ck = clocks.stop()
    tok_s = a.steps / (ms * 1e-3)

    # ---- e2e: token id travels host -> device and back every step (pinned buffers) --------------------
    # public API: LlamaDecodeEngine.step_host() -- one graph replay whose first / last nodes are the 8-byte copies from /
    # to pinned host memory, then a stream synchronisation and the host-side read of the token
    e2e_steps = a.steps
    e2e_how = "per step: pinned host token id -> device, graph replay, next token id -> pinned host, stream sync"
    torch.cuda.synchronize()
    if not a.no_graph:
        eng.capture_host_io()
        eng.step_host()
        torch.cuda.synchronize()
        t_e2e0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng.step_host()                                 # H2D of this step's token, the step, D2H of the sampled token, sync
        t_e2e = time.perf_counter() - t_e2e0
        e2e_how = ("per step: LlamaDecodeEngine.step_host(): ONE graph replay = [pinned host token id -> device] + decode step + "
                   "[next token id -> pinned host], stream sync, host reads the token")
    else:
        h_tok_in = torch.zeros(1, 1, dtype=torch.long).pin_memory()
        h_tok_out = torch.zeros(1, 1, dtype=torch.long).pin_memory()
        h_tok_in.copy_(eng.tok.cpu())
        t_e2e0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng.tok.copy_(h_tok_in, non_blocking=True)      # H2D: this step's input token
            eng.step()
            h_tok_out.copy_(eng.tok, non_blocking=True)     # D2H: the sampled token
            torch.cuda.current_stream().synchronize()
            h_tok_in.copy_(h_tok_out)
        t_e2e = time.perf_counter() - t_e2e0
    e2e_tok_s = e2e_steps / t_e2e

    line = {
        "metric": METRIC, "value": tok_s, "unit": "tokens/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int8 weights x int16 fixed-point activations (exact int32 dp4a), fp16 in/out", "data": "synthetic",
        "config": {"workload": f"{a.model} {a.codebook} bs=1 greedy decode, random-init packed weights, "
                               f"synthetic {a.prompt_len}-token prompt, static KV cache {cache_len}",
                   "l2": "inputs larger than L2: every step streams %.3f GB of packed codes + %.3f GB fp16 lm_head"
                         % (code_bytes / 1e9, lm_head_bytes / 1e9),
                   "cuda_graph": not a.no_graph,
                   "engine": "quip_for_all_b200.modeling.LlamaDecodeEngine (%s)" %
                             ("persistent whole-step kernel" if eng.persistent is not None else "grouped launches")},
        "clocks": ck,
        "e2e": {"value": e2e_tok_s, "unit": "tokens/s", "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 8,
                "how": e2e_how},
        "gpu_launches": (launches_per_step * a.steps) if launches_per_step else None,
        "launches_per_step_own_kernels": launches_per_step,
        "model_roofline": {"packed_code_bytes_per_token": code_bytes,
                           "tok_s_at_hbm_peak": peak * 1e9 / code_bytes,
                           "frac_of_hbm_roofline": tok_s * code_bytes / (peak * 1e9)},
        "build_s": build_s,
    }
    if not a.no_kernel_bench:
        roof, stages = kernel_bench(model, torch, peak, peak_src)
        line["stage_us"] = {k: v["us_per_launch_avg"] for k, v in stages.items()}
        if eng.persistent is not None:
            line["roofline"] = decode_step_kernel_bench(eng, model, torch, code_bytes, peak, peak_src)
            line["roofline_per_linear_kernel"] = roof      # the drop-in op path (one launch per QuantLinear)
        else:
            line["roofline"] = roof
    if not a.no_hf_dropin:
        line["hf_dropin"] = hf_dropin_bench(model, torch, prompt.to(dev), min(a.steps, 64), 4)
    if not a.no_ref_cuda:
        line["ref_cuda"] = ref_cuda_bench(model, torch)
    if not a.no_70b and a.model == "llama2-7b":
        del eng
        model = None
        torch.cuda.empty_cache()
        line["llama2_70b_1gpu"] = single_gpu_70b(a, torch, dev, peak)
    if not a.no_cpu_baseline:
        mname = a.model if a.model in LLAMA_LINEARS else "llama2-7b"
        n_run, _ = _cpu_layers_for_budget(mname, 3, 1, budget_s=20.0)     # ~10-30 s of CPU work
        s = CpuTokenPass(mname, n_run=n_run)
        s.run_once()
        t_pass = min(s.run_once() for _ in range(3))
        line["cpu_baseline"] = {"value": s.tok_s(t_pass), "unit": "tokens/s", "cores": s.cores,
                                "kind": "port", "sample": s.describe()}
    print(json.dumps(line))


def main():
    a = parse()
    if a.model is None:
        a.model = "llama2-70b" if max(a.gpus, int(os.environ.get("WORLD_SIZE", "1"))) > 1 else "llama2-7b"
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
