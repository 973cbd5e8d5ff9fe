"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN PYTHON in the build container.

Run once (already done; outputs are committed):   python tests/golden/gen_golden.py
Needs /root/reference (read-only).  It does not exist on the GPU box, so nothing at test time reads
it -- tests only read the committed .npz files.

What is executed from the reference (imported with cwd=/root/reference because quant.py:8 loads
hadamard.safetensors by relative path):
  * codebook/e8p12.py   get_packed_abs_grid(), get_full_grid()         (E8P tables)
  * codebook/d4.py      build_D4_CB()
  * codebook/e8p12_rvq3.py  get_e81bgrid(), pack_e81b()
  * quant.py            get_hadK(), matmul_hadU()  (the only CPU-runnable transform, quant.py:42-65)
  * qlinear.py          QuantLinear.__init__/forward  (eval branch, qlinear.py:87-115)
The reference has NO CPU implementation of its `quip_lib` ops (register_lib.py: CUDA-only impls), so
for the QuantLinear goldens the ops are bound here to glue made ONLY of reference pieces:
  hadamard(x, s)              := matmul_hadU(x.float(), None, 1, n) * sqrt(n) * s  -> x.dtype
  e8p_mm_origorder(x, Q, cb)  := (x.float() @ get_full_grid()[Q].T).half()   (fp32 accumulate, 1 rounding)
  decompress_*                := table gather from get_full_grid() / build_D4_CB()
Shim: codebook/e8p12.py:96 does np.int8(250), which raises OverflowError under numpy>=2; the module's
`np` is replaced by a proxy whose int8() wraps (numpy-1 behaviour the reference was written against).
"""
import hashlib
import math
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


class _NpProxy:
    def __getattr__(self, k):
        return getattr(np, k)

    @staticmethod
    def int8(v):
        return np.int8(((int(v) + 128) % 256) - 128)


def _import_reference():
    os.chdir(REF)
    sys.path.insert(0, REF)
    import codebook.e8p12 as e8p12
    e8p12.np = _NpProxy()
    import codebook.d4 as d4
    import codebook.e8p12_rvq3 as rvq3
    import codebook.e8p12_rvq4 as rvq4
    import codebook.hi as hi
    import quant
    import qlinear
    return e8p12, d4, rvq3, rvq4, hi, quant, qlinear


def main():
    e8p12, d4, rvq3, rvq4, hi, quant, qlinear = _import_reference()
    torch.manual_seed(0)
    np.random.seed(0)

    # ---------------- tables ----------------
    abs_tab = e8p12.get_packed_abs_grid()
    full_grid, _ = e8p12.get_full_grid(abs_tab)
    d4_cb = d4.build_D4_CB()
    e81b = rvq3.get_e81bgrid()
    e81b_packed = rvq3.pack_e81b(e81b)
    np.savez_compressed(
        os.path.join(OUT, "tables.npz"),
        e8p_abs=abs_tab.numpy().astype(np.int64),
        e8p_full_grid_f16=full_grid.numpy().astype(np.float16),
        d4_grid=d4_cb.numpy().astype(np.float32),
        e81b_grid=e81b.numpy().astype(np.float32),
        e81b_packed=e81b_packed.numpy().astype(np.int32),
        e8p_abs_sha256=hashlib.sha256(abs_tab.numpy().astype("<i8").tobytes()).hexdigest(),
        e8p_full_grid_f16_sha256=hashlib.sha256(full_grid.numpy().astype(np.float16).tobytes()).hexdigest(),
    )

    # ---------------- get_hadK / matmul_hadU ----------------
    had = {}
    shapes = []
    for n in (64, 96, 160, 4096, 11008, 28672, 1024, 8192, 24, 344, 6):
        for use_rand in (True, False):
            torch.manual_seed(n)
            np.random.seed(n)
            hk, K, padn = quant.get_hadK(n, use_rand)
            shapes.append((n, int(use_rand), K, padn, 0 if hk is None else 1))
            if hk is not None and n <= 400:
                had[f"hadK_n{n}_r{int(use_rand)}"] = hk.numpy().astype(np.float32)
                for tr in (False, True):
                    x = torch.randn(3, n, dtype=torch.float32)
                    y = quant.matmul_hadU(x, hk, K, padn, transpose=tr)
                    had[f"x_n{n}_r{int(use_rand)}_t{int(tr)}"] = x.numpy()
                    had[f"y_n{n}_r{int(use_rand)}_t{int(tr)}"] = y.numpy()
            elif hk is None and padn <= 4096:
                x = torch.randn(2, n, dtype=torch.float32)
                y = quant.matmul_hadU(x, None, 1, padn)
                had[f"x_n{n}_r{int(use_rand)}_t0"] = x.numpy()
                had[f"y_n{n}_r{int(use_rand)}_t0"] = y.numpy()
    had["shapes"] = np.array(shapes, dtype=np.int64)
    # the two Hadamard-table blocks Llama-2 needs when use_rand=False (11008 -> 172, 28672 -> 28)
    had["table_172"] = quant.had_tensors["172"].numpy().astype(np.int8)
    had["table_28"] = quant.had_tensors["28"].numpy().astype(np.int8)
    had["table_12"] = quant.had_tensors["12"].numpy().astype(np.int8)
    had["table_20"] = quant.had_tensors["20"].numpy().astype(np.int8)
    had["table_keys"] = np.array(sorted(int(k) for k in quant.had_tensors.keys()), dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "hadamard.npz"), **had)

    # ---------------- quip_lib ops bound to reference-only CPU glue ----------------
    lib = torch.library.Library("quip_lib", "DEF")
    d4_grid = d4_cb

    def _hadamard(x, scale):
        n = x.shape[-1]
        y = quant.matmul_hadU(x.float().reshape(-1, n), None, 1, n) * math.sqrt(n) * scale
        return y.reshape(x.shape).to(x.dtype)

    def _e8p_w(Q):
        idx = Q.to(torch.int64) & 0xFFFF
        return full_grid[idx].reshape(Q.shape[0], -1)

    def _rvq4_w(Q, scale):
        q = Q.to(torch.int64) & 0xFFFFFFFF
        hi_ = full_grid[q >> 16].half()
        lo_ = full_grid[q & 0xFFFF].half()
        s = torch.tensor(scale, dtype=torch.float32).half()
        # __hfma2: single rounding; operands exact in float64
        w = (s.double() * lo_.double() + hi_.double()).half()
        return w.reshape(Q.shape[0], -1)

    def _d4_w(Q, grid):
        return grid.half()[Q.to(torch.int64)].reshape(Q.shape[0], -1)

    lib.define("hadamard(Tensor x, float scale) -> Tensor")
    lib.impl("hadamard", _hadamard, "CPU")
    lib.define("e8p_mm_origorder(Tensor x, Tensor Qidxs, Tensor grid) -> Tensor")
    lib.impl("e8p_mm_origorder", lambda x, Q, g: (x.float() @ _e8p_w(Q).T).half(), "CPU")
    lib.define("decompress_e8p_origorder(Tensor Qidxs, Tensor grid) -> Tensor")
    lib.impl("decompress_e8p_origorder", lambda Q, g: _e8p_w(Q).half(), "CPU")
    lib.define("e8prvq4_mm_origorder(Tensor x, Tensor Qidxs, Tensor grid, float scale) -> Tensor")
    lib.impl("e8prvq4_mm_origorder", lambda x, Q, g, s: (x.float() @ _rvq4_w(Q, s).float().T).half(), "CPU")
    lib.define("decompress_e8prvq4_origorder(Tensor Qidxs, Tensor grid, float scale) -> Tensor")
    lib.impl("decompress_e8prvq4_origorder", lambda Q, g, s: _rvq4_w(Q, s), "CPU")
    lib.define("d4_mm_origorder(Tensor x, Tensor Qidxs, Tensor grid) -> Tensor")
    lib.impl("d4_mm_origorder", lambda x, Q, g: (x.float() @ _d4_w(Q, g).float().T).half(), "CPU")
    lib.define("decompress_d4_origorder(Tensor Qidxs, Tensor grid) -> Tensor")
    lib.impl("decompress_d4_origorder", lambda Q, g: _d4_w(Q, g), "CPU")

    # ---------------- QuantLinear.forward goldens ----------------
    cases = [
        # name, in, out, codebook, bias, use_rand, per_channel, M, drop_SU, drop_SV
        ("e8p_64x32", 64, 32, "E8P12", False, True, False, 1, False, False),
        ("e8p_128x256_b", 128, 256, "E8P12", True, True, False, 3, False, False),
        ("e8p_96x80_rand", 96, 80, "E8P12", True, True, False, 2, False, False),      # K_left=3, K_right=5 random orthogonal
        ("e8p_96x160_tab", 96, 160, "E8P12", False, False, False, 2, False, False),   # use_rand=False: 96->K=12, 160->K=20 tables
        ("e8p_66x50_pad", 66, 50, "E8P12", True, False, False, 2, False, False),      # use_rand=False, exp<2 -> zero-pad to 128 / 64
        ("e8p_64x64_pc", 64, 64, "E8P12", True, True, True, 2, False, False),         # per_channel
        ("e8p_64x64_nosuv", 64, 64, "E8P12", False, True, False, 1, True, True),      # SU = SV = None
        ("e8p_512x512_m40", 512, 512, "E8P12", False, True, False, 40, False, False), # M >= 32 -> decompress + matmul
        ("rvq4_64x64", 64, 64, "E8P12RVQ4B", True, True, False, 2, False, False),
        ("rvq4_96x32", 96, 32, "E8P12RVQ4B", False, True, False, 1, False, False),
        ("d4_64x64", 64, 64, "D4", True, True, False, 2, False, False),
        ("d4_64x96_m30", 64, 96, "D4", False, True, False, 30, False, False),          # M >= 24 -> decompress path
    ]
    import codebook as cbmod
    out = {"names": np.array([c[0] for c in cases])}
    for (name, fin, fout, cbid, bias, use_rand, pc, M, drop_su, drop_sv) in cases:
        seed = int(hashlib.sha256(name.encode()).hexdigest()[:8], 16)
        torch.manual_seed(seed)
        np.random.seed(seed % (2**31))
        cb = cbmod.codebook_id[cbid](inference=True)
        layer = qlinear.QuantLinear(fin, fout, cb, bias=bias, use_rand=use_rand, per_channel=pc,
                                    weight_dtype=torch.float16)
        info = torch.iinfo(cb.idx_dtype)
        layer.Qidxs.copy_(torch.randint(info.min, info.max + 1, layer.Qidxs.shape, dtype=torch.int64).to(cb.idx_dtype))
        # trained-looking (non +-1) fp16 scale vectors, as after fine-tuning
        layer.SU.data.copy_(((torch.randn(fin).sign() + 1e-5).sign() * (1 + 0.1 * torch.randn(fin))).half())
        layer.SV.data.copy_(((torch.randn(fout).sign() + 1e-5).sign() * (1 + 0.1 * torch.randn(fout))).half())
        if pc:
            layer.Wscale.copy_((0.02 * (1 + 0.2 * torch.rand(layer.q_out_features))).half())
        else:
            layer.Wscale.copy_(torch.tensor(0.02 / 1.09375))
        if bias:
            layer.bias.copy_((0.1 * torch.randn(fout)).half())
        # post-load tricks, quantizer.py:836-844
        layer.wscale_float = layer.Wscale.mean().float().item()
        if pc:
            layer.Wscale = layer.Wscale / layer.Wscale.mean()
        if drop_su:
            layer.SU = None
        if drop_sv:
            layer.SV = None
        layer.eval()
        x = torch.randn(M, fin).half()
        with torch.no_grad():
            y = layer(x)
        sd = {k: v for k, v in layer.state_dict().items()}
        pre = name + "/"
        out[pre + "x"] = x.numpy()
        out[pre + "y"] = y.numpy()
        out[pre + "meta"] = np.array([fin, fout, int(bias), int(use_rand), int(pc), M,
                                      layer.K_left, layer.K_right, layer.q_in_features, layer.q_out_features],
                                     dtype=np.int64)
        out[pre + "codebook"] = np.array(cbid)
        out[pre + "wscale_float"] = np.array(layer.wscale_float, dtype=np.float64)
        out[pre + "Qidxs"] = layer.Qidxs.numpy()
        out[pre + "Wscale"] = layer.Wscale.numpy()
        if layer.SU is not None:
            out[pre + "SU"] = layer.SU.detach().numpy()
        if layer.SV is not None:
            out[pre + "SV"] = layer.SV.detach().numpy()
        if layer.bias is not None:
            out[pre + "bias"] = layer.bias.numpy()
        if layer.had_left is not None:
            out[pre + "had_left"] = layer.had_left.numpy()
        if layer.had_right is not None:
            out[pre + "had_right"] = layer.had_right.numpy()
        out[pre + "state_keys"] = np.array(sorted(sd.keys()))
        # the decompressed weight the reference's table produces for these codes
        out[pre + "W_hat"] = cb.decompress_weight(layer.Qidxs).numpy()
    np.savez_compressed(os.path.join(OUT, "quantlinear.npz"), **out)
    for f in ("tables.npz", "hadamard.npz", "quantlinear.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
