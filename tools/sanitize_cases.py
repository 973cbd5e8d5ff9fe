"""Small invocations of every new kernel for compute-sanitizer (memcheck / racecheck); no timing, tiny shapes.
Usage (GPU box): compute-sanitizer --tool memcheck python tools/sanitize_cases.py"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quip_for_all_b200  # noqa: E402,F401
from quip_for_all_b200 import _native, codebook_id  # noqa: E402
from quip_for_all_b200.decode_step import _bind  # noqa: E402
from quip_for_all_b200.modeling import LlamaDecodeEngine, llama_config, make_random_quantized_llama  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
grid = codebook_id["E8P12"](inference=True).to(dev).grid_packed_abs

# persistent decode step: block path (11 x 256), power-of-two path, 1 and 4 KV splits, a few positions
for name in ("tiny256", "tinypow2"):
    for splits in (0, 4):
        _bind().quipb200_decode_step_set_splits(splits)
        model = make_random_quantized_llama(llama_config(name, num_hidden_layers=2), "E8P12", seed=1, device=dev)
        eng = LlamaDecodeEngine(model, max_cache_len=40, persistent=True, use_cuda_graph=False)
        assert eng.persistent is not None
        eng.prefill(torch.randint(0, 32000, (1, 3), generator=g).to(dev))
        for _ in range(3):
            eng.step()
        torch.cuda.synchronize()
        print("decode_step", name, "splits", splits, "ok", int(eng.tok.item()))
_bind().quipb200_decode_step_set_splits(0)

# tcgen05 decode + GEMM: no split-K (1 k-block), split-K, partial token tile
for (M, N, K) in ((5, 128, 128), (17, 256, 1024), (40, 128, 2048), (130, 128, 256)):
    _native.set_option("umma", 1 if M > 16 else 2)
    q = torch.randint(-32768, 32768, (N, K // 8), generator=g).to(torch.int16).to(dev)
    x = torch.randn(M, K, generator=g).half().to(dev)
    y = torch.ops.quip_lib.e8p_mm_origorder(x, q, grid)
    torch.cuda.synchronize()
    print("umma", M, N, K, "ok", float(y.float().abs().max()))
_native.set_option("umma", 2)

# batched rotations: 4096-point, 11 x 256 blocks with mix, padded in/out features
for (n, K, fin, fout, M) in ((4096, 1, 4000, 3968, 3), (2816, 11, 2816, 2800, 5), (256, 1, 256, 256, 2)):
    x = torch.randn(M, fin, generator=g).half().to(dev)
    pre = torch.randn(fin, generator=g).half().to(dev)
    post = torch.randn(fout, generator=g).half().to(dev)
    hk = None
    if K > 1:
        Kp = (K + 15) // 16 * 16
        hk = torch.zeros(Kp, Kp, dtype=torch.float16, device=dev)
        hk[:K, :K] = torch.linalg.qr(torch.randn(K, K, generator=g))[0].half().to(dev)
    y = torch.ops.quip_lib.rotate_fused(x, pre, hk, post, post, n, K, fout, 1.0 / math.sqrt(n // K))
    torch.cuda.synchronize()
    print("rotate", n, K, M, "ok", float(y.float().abs().max()))

# per-linear fused op (round 2 paths): cluster rotations on both sides (K = 3 / 5 / 7 blocks of 256 .. 4096 points), the
# 4096 / 8192-point warp-first FWHT, replicated-table GEMV (>= 8 Mi weights), grouped launch with a shared input
from quip_for_all_b200 import QuantLinear  # noqa: E402
from quip_for_all_b200.modeling import randomize_quantlinear  # noqa: E402
from quip_for_all_b200.quantizer import apply_load_time_tricks  # noqa: E402
from quip_for_all_b200.fused import LinearGroup  # noqa: E402

gd = torch.Generator(device=dev)
gd.manual_seed(0)
for (fin, fout, M) in ((768, 1280, 1), (3 * 1024, 5 * 512, 2), (7 * 4096, 4096, 1), (4096, 7 * 2048, 1), (8192, 4096, 1), (4096, 4096, 3)):
    L = QuantLinear(fin, fout, codebook_id["E8P12"](inference=True), bias=True).to(dev)
    randomize_quantlinear(L, gd)
    L.eval()
    apply_load_time_tricks(torch.nn.ModuleList([L]))
    x = torch.randn(M, fin, generator=g).half().to(dev)
    with torch.no_grad():
        y = L(x)
    torch.cuda.synchronize()
    print("quantlinear", fin, fout, M, "ok", float(y.float().abs().max()))
ls = [QuantLinear(1024, 3 * 256, codebook_id["E8P12"](inference=True), bias=False).to(dev) for _ in range(2)]
for L in ls:
    randomize_quantlinear(L, gd)
    L.eval()
apply_load_time_tricks(torch.nn.ModuleList(ls))
grp = LinearGroup(ls)
h = torch.randn(1, 1024, generator=g).half().to(dev)
w = torch.ones(1024, dtype=torch.float16, device=dev)
with torch.no_grad():
    a, b = grp(h, norm_w=w, eps=1e-5)
torch.cuda.synchronize()
print("group (cluster epilogue, 2 members in one launch) ok", float(a.float().abs().max()), float(b.float().abs().max()))

# quantise-time nearest-codeword search (csrc/nearest.cu): partial thread tiles, several CTAs in x, both RVQ4B stages,
# the explicit-table second stage of RVQ3B
for name in ("E8P12", "E8P12RVQ4B", "E8P12RVQ3B"):
    cb = codebook_id[name](inference=False).to(dev)
    for m in (3, 700):
        xq = (torch.randn(m, 8, generator=g) * 1.1).to(dev)
        vals, idx = cb.quantize(xq)
        torch.cuda.synchronize()
        print("nearest", name, m, "ok", int(idx.max()), float((vals - xq).square().mean()))
