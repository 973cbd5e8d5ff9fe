"""Pins the C restatement of the oracle (oracle/quip_oracle.c) against the numpy oracle and the goldens."""
import os

import numpy as np

import quip_oracle as qo
import quip_oracle_c as qc
from helpers import make_layer, oracle_forward


def test_c_decompress_all_codes_bit_exact():
    codes = np.arange(65536, dtype=np.uint16).reshape(256, 256)
    w = qc.decompress_e8p(codes, qo.e8p_abs_table())
    assert np.array_equal(w.view(np.uint16), qo.e8p_full_grid().astype(np.float16).reshape(256, 2048).view(np.uint16))


def test_c_mm_matches_numpy():
    rng = np.random.default_rng(0)
    q = rng.integers(-32768, 32768, (96, 64)).astype(np.int16)
    x = rng.standard_normal((3, 512)).astype(np.float16)
    y = qc.e8p_mm(x, q, qo.e8p_abs_table())
    ref = x.astype(np.float64) @ qo.decompress_e8p(q).astype(np.float64).T
    assert np.abs(y - ref).max() <= 1e-4 * np.abs(ref).max()


def test_c_forward_matches_numpy_oracle_and_golden(golden_dir):
    for fin, fout, bias in ((256, 512, True), (96, 80, True), (448, 320, False)):
        layer = make_layer(fin, fout, "E8P12", bias=bias, seed=fin)
        x = np.random.default_rng(fin).standard_normal((2, fin)).astype(np.float16)
        import torch
        ref = oracle_forward(layer, torch.tensor(x))
        npf = lambda t: None if t is None else t.detach().float().numpy()
        y = qc.quantlinear_forward_e8p(
            x, layer.Qidxs.numpy(), qo.e8p_abs_table(), fin, fout, layer.q_in_features, layer.q_out_features,
            SU=npf(layer.SU), SV=npf(layer.SV), bias=npf(layer.bias), wscale_float=layer.wscale_float,
            had_left=npf(layer.had_left), K_left=layer.K_left, had_right=npf(layer.had_right), K_right=layer.K_right)
        assert np.abs(y - ref).max() <= 2.0 ** -8 * np.abs(ref).max()
    ql = np.load(os.path.join(golden_dir, "quantlinear.npz"))
    pre = "e8p_128x256_b/"
    y = qc.quantlinear_forward_e8p(ql[pre + "x"], ql[pre + "Qidxs"], qo.e8p_abs_table(), 128, 256, 128, 256,
                                   SU=ql[pre + "SU"], SV=ql[pre + "SV"], bias=ql[pre + "bias"],
                                   wscale_float=float(ql[pre + "wscale_float"]))
    ref = ql[pre + "y"].astype(np.float64)
    assert np.abs(y - ref).max() <= 2.0 ** -8 * np.abs(ref).max()
