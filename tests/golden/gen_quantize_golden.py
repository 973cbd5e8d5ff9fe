"""Generate tests/golden/quantize.npz by RUNNING THE REFERENCE'S OWN PYTHON (quantise-time path).

Run once (outputs are committed):   python tests/golden/gen_quantize_golden.py
Needs /root/reference (read-only; absent on the GPU box -- tests only read the committed .npz).

Executed from the reference, on CPU tensors (these pieces are pure torch):
  * codebook/e8p12.py        E8P12_codebook(inference=False).quantize            (:125-134)
  * codebook/e8p12_rvq4.py   E8P12RVQ4B_codebook(inference=False).quantize       (:32-46)
  * codebook/e8p12_rvq3.py   E8P12RVQ3B_codebook(inference=False).quantize       (:91-101), maybe_pack_idxs (:103-108)
  * codebook/d4.py, hi.py    D4_codebook.quantize, HI4B1C_codebook.quantize / maybe_pack_idxs
  * quant.py                 LDLQ (:107-139), LDLQ_buffered (:142-232), block_LDL (:91-104)
  * quip.py                  QUIP.add_batch / QUIP.quant                         (:43-184)
The same numpy-2 shim as gen_golden.py is applied to codebook/e8p12.py:96.
"""
import os
import sys

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


def main():
    os.chdir(REF)
    sys.path.insert(0, REF)
    import codebook.e8p12 as e8p12
    e8p12.np = _NpProxy()
    import codebook.e8p12_rvq4 as rvq4
    import codebook.e8p12_rvq3 as rvq3
    import codebook.d4 as d4
    import codebook.hi as hi
    import quant
    import quip
    torch.manual_seed(0)
    np.random.seed(0)
    out = {}

    # ---- nearest-codeword search: Gaussian vectors at the scale LDLQ feeds (unit-variance weights / opt_scale),
    # ---- plus outliers, exact lattice points (ties between +-1/4 shifts are impossible there) and zeros
    cb = e8p12.E8P12_codebook(inference=False)
    X = torch.randn(1536, 8) * 1.03
    X[1400:1464] *= 4.0                                       # far outside the ball: norm-12 shell candidates
    X[1464:1528] = cb.grid[torch.randint(0, 65536, (64,))]    # exact codewords
    X[1528:] = 0.0                                            # all-zero rows: 256-way tie -> first index
    vals, idx = cb.quantize(X)
    out["nearest_x"] = X.numpy()
    out["nearest_idx"] = idx.numpy().astype(np.int64)
    out["nearest_vals"] = vals.numpy()
    cb4 = rvq4.E8P12RVQ4B_codebook(inference=False)
    vals4, idx4 = cb4.quantize(X)
    out["nearest_rvq4_idx"] = idx4.numpy().astype(np.int64)
    out["nearest_rvq4_vals"] = vals4.numpy()
    out["rvq4_resid_scale"] = np.float64(cb4.opt_resid_scale)
    cb3 = rvq3.E8P12RVQ3B_codebook(inference=False)
    vals3, idx3 = cb3.quantize(X)
    out["nearest_rvq3_idx"] = idx3.numpy().astype(np.int64)
    out["nearest_rvq3_vals"] = vals3.numpy()
    out["rvq3_resid_scale"] = np.float64(cb3.opt_resid_scale)

    # ---- LDLQ on a small problem, float64 (ties / near-ties out of the picture) and float32
    m, n = 48, 256
    A = torch.randn(n, 4 * n, dtype=torch.float64)
    H = A @ A.T / (4 * n)
    H = H / torch.diag(H).mean()
    H[torch.arange(n), torch.arange(n)] += 0.01
    W = torch.randn(m, n, dtype=torch.float64) * 1.03
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        Hd, Wd = H.to(dt), W.to(dt)
        L = torch.linalg.cholesky(Hd)
        cbq = e8p12.E8P12_codebook(inference=False)
        cbq.grid = cbq.grid.to(dt)
        cbq.grid_norm = cbq.grid_norm.to(dt)
        hat, Q = quant.LDLQ(Wd.clone(), Hd.clone(), L.clone(), cbq, 0)
        hat_b, Q_b = quant.LDLQ_buffered(Wd.clone(), Hd.clone(), L.clone(), cbq, 0, buf_cols=128)
        out[f"ldlq_{tag}_hat"] = hat.numpy()
        out[f"ldlq_{tag}_Q"] = Q.numpy().astype(np.int16)
        out[f"ldlq_{tag}_buffered_same"] = np.bool_(torch.equal(Q, Q_b))
        hat_t, Q_t = quant.LDLQ(Wd.clone(), Hd.clone(), L.clone(), cbq, 1)
        out[f"ldlq_{tag}_tune1_Q"] = Q_t.numpy().astype(np.int16)
    out["ldlq_W"] = W.numpy()
    out["ldlq_H"] = H.numpy()

    # ---- the per-layer driver QUIP.quant on nn.Linear(256 -> 64), fixed SU / SV (taken from the layer when present)
    lin = torch.nn.Linear(256, 64, bias=True)
    lin.weight.data = torch.randn(64, 256) * 0.02
    SU = (torch.randn(256).sign() + 1e-5).sign()
    SV = (torch.randn(64).sign() + 1e-5).sign()
    lin.SU = SU.clone()
    lin.SV = SV.clone()
    w0 = lin.weight.data.clone()
    cb64 = e8p12.E8P12_codebook(inference=False)
    cb64.grid = cb64.grid.double()              # use_fp64=True feeds float64 rows to cb.quantize
    cb64.grid_norm = cb64.grid_norm.double()
    q = quip.QUIP(lin, cb64)
    calib = torch.randn(4, 96, 256)
    for b in range(4):
        q.add_batch(calib[b], None)
    attr = q.quant(rescale_WH=False, use_fp64=True, sigma_reg=0.01, scale_override=0, use_buffered=True, use_rand=True,
                   per_channel=False, quip_tune_iters=0)
    out["quip_w"] = w0.numpy()
    out["quip_bias"] = lin.bias.data.numpy()
    out["quip_calib"] = calib.numpy()
    out["quip_SU"] = SU.numpy()
    out["quip_SV"] = SV.numpy()
    out["quip_Qidxs"] = attr["Qidxs"].numpy().astype(np.int16)
    out["quip_w_scale"] = np.float64(attr["w_scale"].item())
    out["quip_w_hat"] = lin.weight.data.numpy()         # the de-rotated quantised weight QUIP.quant writes back (:160-168)
    # ---- the two small codebooks and the index packing of the two packed formats
    torch.manual_seed(7)
    cbd = d4.D4_codebook(inference=False)
    X4 = torch.randn(600, 4) * 1.21
    X4[:16] = cbd.grid[torch.randint(0, 256, (16,))]
    vd, idd = cbd.quantize(X4)
    out["d4_x"], out["d4_idx"], out["d4_vals"] = X4.numpy(), idd.numpy().astype(np.int64), vd.numpy()
    cbh = hi.HI4B1C_codebook(inference=False)
    X1 = torch.randn(800, 1) * 2.97
    vh, idh = cbh.quantize(X1)
    out["hi_x"], out["hi_idx"], out["hi_vals"] = X1.numpy(), idh.numpy().astype(np.int64), vh.numpy()
    raw_hi = torch.randint(0, 16, (6, 64), dtype=torch.int32)
    out["hi_pack_in"], out["hi_pack_out"] = raw_hi.numpy(), cbh.maybe_pack_idxs(raw_hi).numpy()
    raw3 = ((torch.randint(0, 65536, (5, 16)) << 8) + torch.randint(0, 256, (5, 16))).to(torch.int32)
    out["rvq3_pack_in"], out["rvq3_pack_out"] = raw3.numpy(), cb3.maybe_pack_idxs(raw3).numpy()

    # ---- second driver case: W/H rescaling, per-channel scales, one re-rounding sweep (unbuffered LDLQ)
    torch.manual_seed(1)
    lin2 = torch.nn.Linear(128, 48, bias=False)
    lin2.weight.data = torch.randn(48, 128) * 0.02 * (1 + torch.rand(48, 1))
    SU2 = (torch.randn(128).sign() + 1e-5).sign()
    SV2 = (torch.randn(48).sign() + 1e-5).sign()      # 48 = 3 * 16: random orthogonal 3 x 3 block on the output side
    lin2.SU, lin2.SV = SU2.clone(), SV2.clone()
    w2 = lin2.weight.data.clone()
    cb64b = e8p12.E8P12_codebook(inference=False)
    cb64b.grid = cb64b.grid.double()
    cb64b.grid_norm = cb64b.grid_norm.double()
    q2 = quip.QUIP(lin2, cb64b)
    calib2 = torch.randn(3, 200, 128) * (0.5 + torch.rand(128))
    for b in range(3):
        q2.add_batch(calib2[b], None)
    np.random.seed(3)                                  # get_hadK(use_rand=True) draws scipy special_ortho_group for 48 rows
    attr2 = q2.quant(rescale_WH=True, use_fp64=True, sigma_reg=0.01, scale_override=0, use_buffered=False, use_rand=True,
                     per_channel=True, quip_tune_iters=1)
    out["quip2_w"] = w2.numpy()
    out["quip2_calib"] = calib2.numpy()
    out["quip2_SU"] = SU2.numpy()
    out["quip2_SV"] = SV2.numpy()
    out["quip2_Qidxs"] = attr2["Qidxs"].numpy().astype(np.int16)
    out["quip2_w_scale"] = attr2["w_scale"].numpy().astype(np.float64)
    out["quip2_scaleWH"] = attr2["scaleWH"].numpy().astype(np.float64)
    out["quip2_right_hadK"] = attr2["right_hadK"].numpy().astype(np.float64)
    out["quip2_w_hat"] = lin2.weight.data.numpy()
    np.savez_compressed(os.path.join(OUT, "quantize.npz"), **out)
    for k, v in out.items():
        print(k, getattr(v, "shape", v), getattr(v, "dtype", ""))


if __name__ == "__main__":
    main()
