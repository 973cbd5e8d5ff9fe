"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C-ABI library
via torch.ops.quip_lib; the CPU oracle (oracle/quip_oracle.py) is the checker.

Bars:  dequantised weights / integer work: BIT-EXACT.   Floating-point forward: |y - y_oracle| <=
2^-8 * max|y_oracle| against the oracle evaluated with the reference's fp16 rounding points (helpers.tol_of).
"""
import os

import numpy as np
import pytest
import torch

import quip_oracle as qo
from helpers import make_layer, oracle_forward, oracle_w_hat, tol_of

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _lib_loaded():
    import quip_for_all_b200  # noqa: F401
    from quip_for_all_b200 import _native
    _native.lib()          # fail loudly if the CUDA library is missing
    yield


def _grid(dev=DEV):
    from quip_for_all_b200.codebook.e8p12 import get_packed_abs_grid
    return get_packed_abs_grid().to(dev)


# ------------------------------------------------------------------------------------------------
# dequantisation: bit-exact
# ------------------------------------------------------------------------------------------------
def test_decompress_e8p_all_65536_codes_bit_exact():
    codes = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(256, 256).to(DEV)
    w = torch.ops.quip_lib.decompress_e8p_origorder(codes, _grid())
    ref = qo.e8p_full_grid().astype(np.float16).reshape(256, 2048)
    assert np.array_equal(w.cpu().numpy().view(np.uint16), ref.view(np.uint16))


@pytest.mark.parametrize("shape", [(1, 1), (3, 5), (7, 33), (64, 512), (4096, 512), (0, 8), (5, 0)])
def test_decompress_e8p_shapes(shape):
    g = torch.Generator().manual_seed(shape[0] * 1000 + shape[1])
    q = torch.randint(-32768, 32768, shape, generator=g).to(torch.int16)
    w = torch.ops.quip_lib.decompress_e8p_origorder(q.to(DEV), _grid())
    assert w.shape == (shape[0], shape[1] * 8) and w.dtype == torch.float16
    if q.numel():
        assert np.array_equal(w.cpu().numpy().view(np.uint16), qo.decompress_e8p(q.numpy()).view(np.uint16))


@pytest.mark.parametrize("shape", [(3, 5), (128, 64), (1000, 129)])
def test_decompress_rvq4_d4_rvq3_hi_bit_exact(shape):
    from quip_for_all_b200.codebook.d4 import build_D4_CB
    from quip_for_all_b200.codebook.e8p12_rvq3 import get_e81bgrid, pack_e81b
    g = torch.Generator().manual_seed(7)
    q32 = torch.randint(-2**31, 2**31, shape, generator=g, dtype=torch.int64).to(torch.int32)
    for scale in (1 / 3.45, 0.37):
        w = torch.ops.quip_lib.decompress_e8prvq4_origorder(q32.to(DEV), _grid(), scale)
        assert np.array_equal(w.cpu().numpy().view(np.uint16), qo.decompress_e8prvq4(q32.numpy(), scale).view(np.uint16))
    q8 = torch.randint(0, 256, shape, generator=g).to(torch.uint8)
    for gd in (build_D4_CB().half(), build_D4_CB()):     # fp16 contract; fp32 grid is cast by the binding
        w = torch.ops.quip_lib.decompress_d4_origorder(q8.to(DEV), gd.to(DEV))
        assert np.array_equal(w.cpu().numpy().view(np.uint16), qo.decompress_d4(q8.numpy()).view(np.uint16))
    w = torch.ops.quip_lib.decompress_hi_origorder(q32.to(DEV))
    assert np.array_equal(w.cpu().numpy().view(np.uint16), qo.decompress_hi(q32.numpy()).view(np.uint16))
    q3 = torch.randint(-2**31, 2**31, (shape[0], 3 * ((shape[1] + 3) // 4)), generator=g, dtype=torch.int64).to(torch.int32)
    cb2 = pack_e81b(get_e81bgrid()).to(DEV)
    w = torch.ops.quip_lib.decompress_e8prvq3_origorder(q3.to(DEV), _grid(), cb2, 1 / 2.04)
    assert np.array_equal(w.cpu().numpy().view(np.uint16), qo.decompress_e8prvq3(q3.numpy(), 1 / 2.04).view(np.uint16))


def test_decompress_vs_golden_reference_weights(golden_dir):
    ql = np.load(os.path.join(golden_dir, "quantlinear.npz"))
    from quip_for_all_b200 import codebook_id
    for name in ql["names"]:
        pre = str(name) + "/"
        cb = codebook_id[str(ql[pre + "codebook"])](inference=True).to(DEV)
        w = cb.decompress_weight(torch.tensor(ql[pre + "Qidxs"]).to(DEV))
        assert np.array_equal(w.cpu().numpy().view(np.uint16), ql[pre + "W_hat"].view(np.uint16)), name


# ------------------------------------------------------------------------------------------------
# hadamard op
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 4, 8, 64, 256, 1024, 4096, 8192, 32768])
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.bfloat16])
def test_hadamard_matches_oracle(n, dtype):
    g = torch.Generator().manual_seed(n)
    rows = 5 if n <= 4096 else 2
    x = torch.randn(rows, n, generator=g).to(dtype)
    scale = 1.0 / np.sqrt(n)
    y = torch.ops.quip_lib.hadamard(x.to(DEV), scale)
    assert y.dtype == dtype and y.shape == x.shape
    ref = qo.fwht(x.float().numpy(), scale)
    eps = {torch.float16: 2.0 ** -10, torch.bfloat16: 2.0 ** -7, torch.float32: 2.0 ** -20}[dtype]
    assert np.abs(y.float().cpu().numpy() - ref).max() <= eps * max(1.0, np.abs(ref).max())


def test_hadamard_batched_3d_and_involution():
    x = torch.randn(4, 43, 256, device=DEV, dtype=torch.float32)
    y = torch.ops.quip_lib.hadamard(x, 1 / 16.0)
    ref = qo.fwht(x.cpu().numpy(), 1 / 16.0)
    assert np.abs(y.cpu().numpy() - ref).max() < 1e-4
    z = torch.ops.quip_lib.hadamard(y, 1 / 16.0)           # H H = n I
    assert torch.allclose(z, x, atol=1e-4)
    assert torch.ops.quip_lib.hadamard(torch.zeros(0, 64, device=DEV), 1.0).shape == (0, 64)
    with pytest.raises(RuntimeError):
        torch.ops.quip_lib.hadamard(torch.zeros(2, 24, device=DEV), 1.0)


# ------------------------------------------------------------------------------------------------
# decode + matmul ops
# ------------------------------------------------------------------------------------------------
def _mm_check(out, x, W_hat, slack=1.0):
    ref64 = x.double().cpu().numpy() @ W_hat.astype(np.float64).T
    err = np.abs(out.float().cpu().numpy() - ref64)
    # fp16 output rounding (2^-11 relative) + 16-bit fixed-point activations: 2^-9 of the output scale
    tol = slack * 2.0 ** -9 * np.abs(ref64).max()
    assert err.max() <= tol, (err.max(), tol)
    return err.max() / np.abs(ref64).max()


@pytest.mark.parametrize("M", [1, 2, 7, 16, 17, 40])
@pytest.mark.parametrize("N,K", [(512, 4096), (64, 64), (8, 128), (96, 2048 + 64)])
def test_e8p_mm(M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N)
    q = torch.randint(-32768, 32768, (N, K // 8), generator=g).to(torch.int16)
    x = torch.randn(M, K, generator=g).half()
    out = torch.ops.quip_lib.e8p_mm_origorder(x.to(DEV), q.to(DEV), _grid())
    assert out.shape == (M, N) and out.dtype == torch.float16
    _mm_check(out, x, qo.decompress_e8p(q.numpy()))


def test_e8p_mm_row_pitch_not_16B_takes_dense_path():
    q = torch.randint(-32768, 32768, (24, 12)).to(torch.int16)       # K = 96: 24-byte rows
    x = torch.randn(3, 96).half()
    out = torch.ops.quip_lib.e8p_mm_origorder(x.to(DEV), q.to(DEV), _grid())
    _mm_check(out, x, qo.decompress_e8p(q.numpy()))


def test_e8p_mm_edge_inputs():
    q = torch.randint(-32768, 32768, (64, 64)).to(torch.int16).to(DEV)
    W = qo.decompress_e8p(q.cpu().numpy())
    z = torch.ops.quip_lib.e8p_mm_origorder(torch.zeros(2, 512, device=DEV, dtype=torch.float16), q, _grid())
    assert torch.count_nonzero(z) == 0                                  # all-zero row: scale 0 guard
    assert torch.ops.quip_lib.e8p_mm_origorder(torch.zeros(0, 512, device=DEV, dtype=torch.float16), q, _grid()).shape == (0, 64)
    # wide dynamic range inside one row: one element 1000x larger than the rest
    x = torch.randn(1, 512).half()
    x[0, 3] = 800.0
    out = torch.ops.quip_lib.e8p_mm_origorder(x.to(DEV), q, _grid())
    _mm_check(out, x, W)
    # tiny magnitudes (fp16 subnormal region scaled up by the per-row scale)
    x = (torch.randn(2, 512) * 1e-4).half()
    out = torch.ops.quip_lib.e8p_mm_origorder(x.to(DEV), q, _grid())
    _mm_check(out, x, W, slack=2.0)
    with pytest.raises(RuntimeError):
        torch.ops.quip_lib.e8p_mm_origorder(torch.zeros(1, 256, device=DEV, dtype=torch.float16), q, _grid())


@pytest.mark.parametrize("M", [1, 5, 30])
def test_rvq4_and_d4_mm(M):
    from quip_for_all_b200.codebook.d4 import build_D4_CB
    g = torch.Generator().manual_seed(M)
    N, K = 256, 1024
    x = torch.randn(M, K, generator=g).half()
    q = torch.randint(-2**31, 2**31, (N, K // 8), generator=g, dtype=torch.int64).to(torch.int32)
    out = torch.ops.quip_lib.e8prvq4_mm_origorder(x.to(DEV), q.to(DEV), _grid(), 1 / 3.45)
    _mm_check(out, x, qo.decompress_e8prvq4(q.numpy()), slack=1.5)
    q8 = torch.randint(0, 256, (N, K // 4), generator=g).to(torch.uint8)
    out = torch.ops.quip_lib.d4_mm_origorder(x.to(DEV), q8.to(DEV), build_D4_CB().to(DEV))
    _mm_check(out, x, qo.decompress_d4(q8.numpy()))


def test_rvq3_and_hi_mm_dense_path():
    from quip_for_all_b200.codebook.e8p12_rvq3 import get_e81bgrid, pack_e81b
    g = torch.Generator().manual_seed(3)
    N, K = 64, 256
    x = torch.randn(4, K, generator=g).half()
    q3 = torch.randint(-2**31, 2**31, (N, 3 * K // 32), generator=g, dtype=torch.int64).to(torch.int32)
    out = torch.ops.quip_lib.e8prvq3_mm_origorder(x.to(DEV), q3.to(DEV), _grid(), pack_e81b(get_e81bgrid()).to(DEV), 1 / 2.04)
    _mm_check(out, x, qo.decompress_e8prvq3(q3.numpy()), slack=2.0)
    qh = torch.randint(-2**31, 2**31, (N, K // 8), generator=g, dtype=torch.int64).to(torch.int32)
    out = torch.ops.quip_lib.hi_mm_origorder(x.to(DEV), qh.to(DEV))
    _mm_check(out, x, qo.decompress_hi(qh.numpy()), slack=2.0)


@pytest.mark.parametrize("M", [1, 3, 8, 17, 31, 32, 40])
@pytest.mark.parametrize("N,K", [(128, 128), (4096, 4096), (1408, 1152)])
def test_rvq3_and_hi_mm_tcgen05_route(M, N, K):
    """K4 / K5 replacements (origin_order.cu:287-335 / :170-206 with hosts :650-696 / :745-788): for M <= 32 the RVQ3B and HI
    mm ops are ONE launch of the codebook-templated tcgen05 kernel (3-byte codes + nibble residual table with one fp16
    fma; 4-bit scalar codes); above that the reference's own route, decompress + GEMM.  Against the oracle's dequantised
    weights (bit-exact on their own, test_decompress_rvq4_d4_rvq3_hi_bit_exact)."""
    from quip_for_all_b200 import _native
    from quip_for_all_b200.codebook.e8p12_rvq3 import get_e81bgrid, pack_e81b
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).half()
    xd = x.to(DEV)
    e81b = pack_e81b(get_e81bgrid()).to(DEV)
    q3 = torch.randint(-2**31, 2**31, (N, 3 * K // 32), generator=g, dtype=torch.int64).to(torch.int32)
    qh = torch.randint(-2**31, 2**31, (N, K // 8), generator=g, dtype=torch.int64).to(torch.int32)
    lc0 = _native.launch_count()
    o3 = torch.ops.quip_lib.e8prvq3_mm_origorder(xd, q3.to(DEV), _grid(), e81b, 1 / 2.04)
    oh = torch.ops.quip_lib.hi_mm_origorder(xd, qh.to(DEV))
    n_launch = _native.launch_count() - lc0
    assert n_launch == (2 if M <= 32 else 2)      # M <= 32: one tcgen05 launch each; above: one decompress launch each (+ cuBLAS)
    _mm_check(o3, x, qo.decompress_e8prvq3(q3.numpy(), 1 / 2.04))
    _mm_check(oh, x, qo.decompress_hi(qh.numpy()))
    if M <= 32:        # and against the dense route on the same device
        _native.set_option("umma", 0)
        try:
            d3 = torch.ops.quip_lib.e8prvq3_mm_origorder(xd, q3.to(DEV), _grid(), e81b, 1 / 2.04)
            dh = torch.ops.quip_lib.hi_mm_origorder(xd, qh.to(DEV))
        finally:
            _native.set_option("umma", 2)
        for a_, b_ in ((o3, d3), (oh, dh)):
            assert (a_.float() - b_.float()).abs().max().item() <= 2.0 ** -9 * b_.float().abs().max().item() + 1e-6


# ------------------------------------------------------------------------------------------------
# fused QuantLinear.forward
# ------------------------------------------------------------------------------------------------
def _layer_from_golden(ql, name):
    from quip_for_all_b200 import QuantLinear, codebook_id, quant
    pre = name + "/"
    fin, fout, bias, use_rand, pc, M, K_left, K_right, q_in, q_out = [int(v) for v in ql[pre + "meta"]]
    cb = codebook_id[str(ql[pre + "codebook"])](inference=True)
    layer = QuantLinear(fin, fout, cb, bias=bool(bias), use_rand=bool(use_rand), per_channel=bool(pc))
    assert (layer.K_left, layer.K_right, layer.q_in_features, layer.q_out_features) == (K_left, K_right, q_in, q_out)
    layer.Qidxs.copy_(torch.tensor(ql[pre + "Qidxs"]))
    layer.Wscale.copy_(torch.tensor(ql[pre + "Wscale"]))
    for attr in ("SU", "SV"):
        if pre + attr in ql.files:
            getattr(layer, attr).data.copy_(torch.tensor(ql[pre + attr]))
        else:
            setattr(layer, attr, None)
    if bias:
        layer.bias.copy_(torch.tensor(ql[pre + "bias"]))
    for attr in ("had_left", "had_right"):
        if pre + attr in ql.files:
            getattr(layer, attr).copy_(torch.tensor(ql[pre + attr]))
    layer.wscale_float = float(ql[pre + "wscale_float"])
    return layer.to(DEV).eval(), torch.tensor(ql[pre + "x"]), ql[pre + "y"]


def test_quantlinear_forward_vs_reference_golden(golden_dir):
    """Product forward on the GPU vs outputs of the reference's own QuantLinear.forward (gen_golden.py)."""
    from quip_for_all_b200 import quant
    had = np.load(os.path.join(golden_dir, "hadamard.npz"))
    for k in (12, 20, 28, 172):
        quant.register_had_table(k, torch.tensor(had[f"table_{k}"].astype(np.float32)))
    ql = np.load(os.path.join(golden_dir, "quantlinear.npz"))
    for name in ql["names"]:
        layer, x, yref = _layer_from_golden(ql, str(name))
        with torch.no_grad():
            y = layer(x.to(DEV))
        assert y.shape == yref.shape and y.dtype == torch.float16
        err = np.abs(y.float().cpu().numpy() - yref.astype(np.float64)).max()
        assert err <= tol_of(yref), (name, err, tol_of(yref))


CASES = [
    # fin, fout, codebook, bias, use_rand, per_channel, M
    (4096, 4096, "E8P12", False, True, False, 1),        # BASELINE config 1 / q,k,v,o of Llama-2-7B
    (4096, 11008, "E8P12", False, True, False, 1),       # gate/up: K_right = 43
    (11008, 4096, "E8P12", False, True, False, 1),       # down: K_left = 43
    (1024, 8192, "E8P12", True, True, False, 3),
    (8192, 1024, "E8P12", True, True, True, 2),          # per-channel
    (4096, 4096, "E8P12RVQ4B", False, True, False, 1),
    (4096, 4096, "D4", True, True, False, 2),
    (2048, 2048, "E8P12", False, True, False, 16),
    (2048, 2048, "E8P12", False, True, False, 40),       # M >= 32: reference op sequence (decompress + GEMM)
    (448, 320, "E8P12", True, True, False, 4),           # K_left = 7, K_right = 5, 64 | q_in
    # BASELINE config 4 codebooks on the 7B MLP shapes (K_right / K_left = 43)
    (4096, 11008, "E8P12RVQ4B", False, True, False, 1),
    (11008, 4096, "E8P12RVQ4B", False, True, False, 1),
    (4096, 11008, "D4", False, True, False, 1),
    (11008, 4096, "D4", True, True, False, 1),
    # BASELINE config 5: every Llama-2-70B linear: 8192-point rotations, 28672 = 7 x 4096
    (8192, 8192, "E8P12", False, True, False, 1),        # q, o
    (8192, 1024, "E8P12", False, True, False, 1),        # k, v
    (8192, 28672, "E8P12", False, True, False, 1),       # gate, up: K_right = 7
    (28672, 8192, "E8P12", False, True, False, 1),       # down: K_left = 7
    (8192, 8192, "E8P12", False, True, False, 40),
    (8192, 1024, "E8P12", False, True, False, 40),
    (8192, 28672, "E8P12", False, True, False, 40),
    (28672, 8192, "E8P12", False, True, False, 40),
    # Llama-2-13B dims: 5120 = 5 x 1024, 13824 = 27 x 512 (wide rotation blocks with an orthogonal mix), and 3 x 2048
    (5120, 5120, "E8P12", False, True, False, 1),
    (5120, 13824, "E8P12", False, True, False, 1),
    (13824, 5120, "E8P12", True, True, False, 2),
    (6144, 6144, "E8P12", False, True, True, 1),
]


@pytest.mark.parametrize("fin,fout,cbid,bias,use_rand,pc,M", CASES)
def test_quantlinear_forward_vs_oracle(fin, fout, cbid, bias, use_rand, pc, M):
    layer = make_layer(fin, fout, cbid, bias=bias, use_rand=use_rand, per_channel=pc, seed=fin + fout + M, device=DEV)
    x = torch.randn(M, fin, generator=torch.Generator().manual_seed(M)).half()
    with torch.no_grad():
        y = layer(x.to(DEV))
    ref = oracle_forward(layer, x, rounding="reference")
    err = np.abs(y.float().cpu().numpy() - ref)
    assert err.max() <= tol_of(ref), (err.max(), tol_of(ref))
    # and against the unrounded fp64 forward
    ref64 = oracle_forward(layer, x, rounding="none")
    assert np.abs(y.float().cpu().numpy() - ref64).max() <= tol_of(ref64)
    # relative RMS error well inside fp16 resolution
    assert np.sqrt((err ** 2).mean()) <= 2.0 ** -10 * np.sqrt((ref ** 2).mean()) * 2


def test_fused_matches_reference_op_sequence():
    """The fused op and the reference's own op sequence (through the individual quip_lib ops) agree."""
    layer = make_layer(4096, 11008, "E8P12", bias=True, seed=5, device=DEV)
    x = torch.randn(2, 4096, generator=torch.Generator().manual_seed(0)).half().to(DEV)
    with torch.no_grad():
        y_fused = layer(x)
        from quip_for_all_b200 import register_lib
        old = register_lib.fused_supported
        try:
            import quip_for_all_b200.qlinear as qmod
            qmod.fused_supported = lambda *a: False
            y_seq = layer(x)
        finally:
            qmod.fused_supported = old
    assert (y_fused - y_seq).abs().max().item() <= 2.0 ** -8 * y_seq.abs().max().item()


def test_properties_at_full_size():
    """Size-independent properties at BASELINE's full layer size: linearity and sign symmetry."""
    layer = make_layer(4096, 4096, "E8P12", seed=11, device=DEV)
    x = torch.randn(1, 4096, generator=torch.Generator().manual_seed(1)).half().to(DEV)
    with torch.no_grad():
        y1, y2, yn = layer(x), layer(2 * x), layer(-x)
        z = layer(torch.zeros_like(x))
    assert torch.count_nonzero(z) == 0
    assert torch.equal(yn, -y1)                                   # exact: integer arithmetic is sign-symmetric
    assert (y2 - 2 * y1).abs().max().item() <= 2.0 ** -9 * y2.abs().max().item()
    # batch invariance on the integer-dp4a path (M <= 3): row i of a batch == the same row alone, bit-exact
    xb = torch.randn(5, 4096, generator=torch.Generator().manual_seed(2)).half().to(DEV)
    with torch.no_grad():
        yb3 = layer(xb[:3])
        for i in range(3):
            assert torch.equal(yb3[i:i + 1], layer(xb[i:i + 1]))
        # M >= 4 takes the rotations + tcgen05 mm route (fp16 x fp16 -> fp32): equal to fp16 noise
        yb = layer(xb)
        for i in range(5):
            assert (yb[i:i + 1] - layer(xb[i:i + 1])).abs().max().item() <= 2.0 ** -8 * yb.abs().max().item()


def test_non_fp16_activations_and_3d_input():
    layer = make_layer(512, 256, "E8P12", bias=True, seed=3, device=DEV)
    x = torch.randn(2, 3, 512, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        y16 = layer(x.half().to(DEV))
        y32 = layer(x.to(DEV))
        ybf = layer(x.bfloat16().to(DEV))
    assert y16.shape == (2, 3, 256) and y32.dtype == torch.float32 and ybf.dtype == torch.bfloat16
    assert (y32 - y16.float()).abs().max() <= 2.0 ** -8 * y16.abs().max()


@pytest.mark.parametrize("fin,fout,cbid,pc", [(512, 768, "E8P12", False), (768, 512, "E8P12RVQ4B", True), (1024, 1024, "D4", False)])
def test_calc_weight_vs_oracle_dense_weight(fin, fout, cbid, pc):
    """calc_weight (qlinear.py:144-159) against the oracle's dense matrix M_L W_hat^T M_R^T * Wscale: the float64
    forward of the identity without SU / SV / bias, and the training-mode forward against the oracle."""
    layer = make_layer(fin, fout, cbid, bias=True, per_channel=pc, seed=fin + 3, device=DEV)
    with torch.no_grad():
        W = layer.calc_weight(cache=False).float().cpu().numpy()   # [q_in, q_out] for `x @ W`; SU / SV / bias are applied around it
    qi, qo_ = layer.q_in_features, layer.q_out_features
    assert W.shape == (qi, qo_)
    cb = layer.codebook
    W_hat = oracle_w_hat(cb.id, layer.Qidxs.cpu().numpy(), getattr(cb, "opt_resid_scale", None))
    npy = lambda t: None if t is None else t.detach().float().cpu().numpy()
    # calc_weight takes its global scale from Wscale.mean() and, per channel, multiplies the ROTATED matrix by
    # Wscale / Wscale.mean() (qlinear.py:146,155-156): after the load-time normalisation (quantizer.py:838-839) that mean is
    # 1, not wscale_float, and the per-channel factor sits after the output rotation, not before it as in the eval forward
    # (qlinear.py:106-109).  Reference behaviour, restated as is.
    wsf = float(layer.Wscale.mean())
    W_ref = qo.quantlinear_forward(np.eye(qi, dtype=np.float16), W_hat=W_hat, in_features=qi, out_features=qo_, q_in=qi,
                                   q_out=qo_, SU=None, SV=None, bias=None, wscale_float=wsf, Wscale_per_channel=None,
                                   had_left=npy(layer.had_left), K_left=layer.K_left, had_right=npy(layer.had_right),
                                   K_right=layer.K_right, rounding="none")     # rows of the identity through the path
    if pc:
        W_ref = W_ref * npy(layer.Wscale / layer.Wscale.mean())[None, :]
    # fp16 dense weight: two fp16 Hadamard passes + scalings, 2^-8 of the largest entry
    assert np.abs(W - W_ref).max() <= 2.0 ** -8 * np.abs(W_ref).max(), (np.abs(W - W_ref).max(), np.abs(W_ref).max())
    x = torch.randn(5, fin, generator=torch.Generator().manual_seed(4)).half()
    with torch.no_grad():
        layer.train()
        y_train = layer(x.to(DEV)).float().cpu().numpy()
        layer.eval()
    xs = np.zeros((5, qi))
    xs[:, :fin] = x.double().numpy() * npy(layer.SU).astype(np.float64)
    ref = (xs @ W_ref)[:, :fout] * npy(layer.SV).astype(np.float64) + npy(layer.bias).astype(np.float64)   # qlinear.py:93-97,111-114
    assert np.abs(y_train - ref).max() <= 2.0 ** -7 * np.abs(ref).max()


def test_training_mode_dense_weight_matches_eval():
    """calc_weight (dense dequant + two-sided Hadamard, qlinear.py:144-159) reproduces the eval forward."""
    layer = make_layer(512, 768, "E8P12", bias=True, seed=9, device=DEV)   # 768 = 3 * 256
    x = torch.randn(4, 512, generator=torch.Generator().manual_seed(4)).half().to(DEV)
    with torch.no_grad():
        y_eval = layer(x)
        layer.train()
        y_train = layer(x)
        layer.eval()
    assert (y_eval - y_train).abs().max().item() <= 2.0 ** -6 * y_eval.abs().max().item()


def test_cuda_graph_capture_of_fused_forward():
    layer = make_layer(4096, 4096, "E8P12", seed=2, device=DEV)
    x = torch.randn(1, 4096, device=DEV, dtype=torch.float16)
    with torch.no_grad():
        y_eager = layer(x).clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            layer(x)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            y_g = layer(x)
        x.copy_(torch.randn(1, 4096, device=DEV, dtype=torch.float16))
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(y_g, layer(x))
    assert not torch.equal(y_g, y_eager)


# ------------------------------------------------------------------------------------------------
# vs the reference's own CUDA kernels (oracle/_ref/quiptools_cuda.so), when the build shipped it
# ------------------------------------------------------------------------------------------------
def _ref_mod():
    import build_ref
    m = build_ref.load_ref_module()
    if m is None:      # the compiled reference extension is part of the snapshot (oracle/build_ref.py via build()): its absence on
        pytest.fail("oracle/_ref/quiptools_cuda.so not present: run __graft_entry__.build() where /root/reference exists")
    return m


def test_vs_reference_cuda_kernels():
    from quip_for_all_b200.codebook.d4 import build_D4_CB
    ref = _ref_mod()
    g = torch.Generator().manual_seed(0)
    N, K = 4096, 4096
    grid = _grid()
    q = torch.randint(-32768, 32768, (N, K // 8), generator=g).to(torch.int16).to(DEV)
    Y = torch.empty(N, K, dtype=torch.float16, device=DEV)
    ref.decompress_e8p_origorder(q, grid, Y)
    assert torch.equal(Y, torch.ops.quip_lib.decompress_e8p_origorder(q, grid))       # K6 bit-exact
    q4 = torch.randint(-2**31, 2**31, (N, K // 8), generator=g, dtype=torch.int64).to(torch.int32).to(DEV)
    ref.decompress_e8prvq4_origorder(q4, grid, Y, 1 / 3.45)
    assert torch.equal(Y, torch.ops.quip_lib.decompress_e8prvq4_origorder(q4, grid, 1 / 3.45))   # K7
    q8 = torch.randint(0, 256, (N, K // 4), generator=g).to(torch.uint8).to(DEV)
    cb = build_D4_CB().half().to(DEV)
    ref.decompress_d4_origorder(q8, cb, Y)
    assert torch.equal(Y, torch.ops.quip_lib.decompress_d4_origorder(q8, cb))         # K8
    for M in (1, 4, 16):
        x = torch.randn(M, K, generator=g).half().to(DEV)
        for mine, theirs in (
            (torch.ops.quip_lib.e8p_mm_origorder(x, q, grid), ref.e8p_mm_origorder(x, q, grid)),                  # K1
            (torch.ops.quip_lib.e8prvq4_mm_origorder(x, q4, grid, 1 / 3.45), ref.e8prvq4_mm_origorder(x, q4, grid, 1 / 3.45)),  # K2
            (torch.ops.quip_lib.d4_mm_origorder(x, q8, cb), ref.d4_mm_origorder(x, q8, cb)),                      # K3
        ):
            d = (mine.float() - theirs.float()).abs().max().item()
            assert d <= 2.0 ** -9 * theirs.float().abs().max().item(), d


# ------------------------------------------------------------------------------------------------
# drop-in into the HF forward pass, and the decode engine
# ------------------------------------------------------------------------------------------------
def _tiny_model(codebook="E8P12"):
    from quip_for_all_b200.modeling import make_random_quantized_llama
    return make_random_quantized_llama("tiny", codebook, seed=0, device=DEV)


def _densify(model):
    """Replace every QuantLinear by an nn.Linear holding its dense dequantised weight."""
    import copy
    from quip_for_all_b200 import QuantLinear
    dense = copy.deepcopy(model)
    dense.is_quantized = False
    for name, mod in list(dense.named_modules()):
        for cname, child in list(mod.named_children()):
            if isinstance(child, QuantLinear):
                W = child.calc_weight(cache=False)                       # (q_in, q_out), x @ W
                W = W[:child.in_features, :child.out_features].float()
                if child.SU is not None:
                    W = child.SU.float()[:, None] * W
                if child.SV is not None:
                    W = W * child.SV.float()[None, :]
                lin = torch.nn.Linear(child.in_features, child.out_features, bias=child.bias is not None,
                                      device=DEV, dtype=torch.float32)
                lin.weight.data.copy_(W.T)
                if child.bias is not None:
                    lin.bias.data.copy_(child.bias.float())
                setattr(mod, cname, lin)
    return dense.float()


def test_hf_forward_dropin_decode_and_prefill():
    model = _tiny_model()
    dense = _densify(model)
    ids = torch.randint(0, 32000, (1, 40), generator=torch.Generator().manual_seed(0)).to(DEV)
    with torch.no_grad():
        lq = model(ids).logits.float()           # M = 40 rows: reference op sequence
        ld = dense(ids).logits
        assert (lq - ld).abs().max() <= 0.03 * ld.abs().max()
        lq1 = model(ids[:, :5]).logits.float()   # M = 5 rows: fused path
        ld1 = dense(ids[:, :5]).logits
        assert (lq1 - ld1).abs().max() <= 0.03 * ld1.abs().max()


def test_decode_engine_matches_hf_greedy():
    from quip_for_all_b200.modeling import LlamaDecodeEngine
    model = _tiny_model()
    ids = torch.randint(0, 32000, (1, 12), generator=torch.Generator().manual_seed(1)).to(DEV)
    eng = LlamaDecodeEngine(model, max_cache_len=64)
    out = eng.generate(ids, 8)
    eng2 = LlamaDecodeEngine(model, max_cache_len=64, use_cuda_graph=False)
    out2 = eng2.generate(ids, 8)
    assert torch.equal(out, out2)                                 # graph replay == eager
    with torch.no_grad():
        seq = ids
        hf = []
        for _ in range(8):
            nxt = model(seq).logits[:, -1].argmax(-1, keepdim=True)
            hf.append(nxt)
            seq = torch.cat([seq, nxt], 1)
    hf = torch.cat(hf, 1)
    # identical up to fp16 noise flipping a near-tie: require the first tokens to agree
    assert torch.equal(out[:, :2], hf[:, :2])


# ------------------------------------------------------------------------------------------------
# grouped / hooked launches and the fused decode step
# ------------------------------------------------------------------------------------------------
def test_linear_group_matches_individual_layers():
    from quip_for_all_b200.fused import LinearGroup
    layers = [make_layer(512, n, "E8P12", bias=(i == 1), seed=30 + i, device=DEV) for i, n in enumerate((512, 256, 768))]
    x = torch.randn(1, 512, generator=torch.Generator().manual_seed(9)).half().to(DEV)
    with torch.no_grad():
        outs = [o.clone() for o in LinearGroup(layers)(x)]
        for o, l in zip(outs, layers):
            assert torch.equal(o, l(x))                      # same kernels, same integer sums: bit-identical


def test_fusion_hooks_norm_gate_residual():
    from quip_for_all_b200.fused import LinearGroup
    g = torch.Generator().manual_seed(4)
    lay = make_layer(1408, 512, "E8P12", bias=True, seed=77, device=DEV)       # K_left = 11
    lay2 = make_layer(512, 1408, "E8P12", seed=78, device=DEV)                 # K_right = 11
    x = torch.randn(1, 1408, generator=g).half().to(DEV)
    gate = torch.randn(1, 1408, generator=g).half().to(DEV)
    res = torch.randn(1, 512, generator=g).half().to(DEV)
    w = (1 + 0.1 * torch.randn(512, generator=g)).half().to(DEV)
    with torch.no_grad():
        y = LinearGroup([lay])(x, gate=gate, residual=res)[0].clone()
        ref = lay(torch.nn.functional.silu(gate) * x) + res
        assert (y - ref).abs().max().item() <= 2.0 ** -8 * ref.abs().max().item()
        h = torch.randn(1, 512, generator=g).half().to(DEV)
        y2 = LinearGroup([lay2])(h, norm_w=w, eps=1e-5)[0].clone()
        hn = h.float()
        hn = (hn * torch.rsqrt(hn.pow(2).mean(-1, keepdim=True) + 1e-5)).half()
        ref2 = lay2(w * hn)
        assert (y2 - ref2).abs().max().item() <= 2.0 ** -8 * ref2.abs().max().item()


def test_attn_decode_kernel_matches_sdpa():
    from quip_for_all_b200.fused import attn_decode
    nh, nkv, hd, L, pos = 8, 2, 128, 96, 37
    g = torch.Generator().manual_seed(5)
    q = torch.randn(1, nh * hd, generator=g).half().to(DEV)
    k = torch.randn(1, nkv * hd, generator=g).half().to(DEV)
    v = torch.randn(1, nkv * hd, generator=g).half().to(DEV)
    kc = torch.randn(1, nkv, L, hd, generator=g).half().to(DEV)
    vc = torch.randn(1, nkv, L, hd, generator=g).half().to(DEV)
    inv = 1.0 / (10000 ** (torch.arange(0, hd, 2).float() / hd))
    fr = torch.outer(torch.arange(L).float(), inv)
    emb = torch.cat((fr, fr), -1)
    cos, sin = emb.cos().half().to(DEV), emb.sin().half().to(DEV)
    p = torch.tensor([pos], device=DEV)
    out = torch.empty(1, nh * hd, dtype=torch.float16, device=DEV)
    kc2, vc2 = kc.clone(), vc.clone()
    attn_decode(q, k, v, kc2, vc2, cos, sin, p, out, nh, nkv, hd)
    rot = lambda x: torch.cat((-x[..., hd // 2:], x[..., :hd // 2]), -1)
    qh = q.view(1, nh, 1, hd).float()
    kh = k.view(1, nkv, 1, hd).float()
    c, s = cos[pos].float(), sin[pos].float()
    qr = (qh * c + rot(qh) * s).half().float()
    kr = (kh * c + rot(kh) * s).half()
    kref, vref = kc.clone(), vc.clone()
    kref[:, :, pos] = kr[:, :, 0]
    vref[:, :, pos] = v.view(1, nkv, hd)
    assert torch.equal(kc2[:, :, :pos + 1], kref[:, :, :pos + 1]) and torch.equal(vc2[:, :, :pos + 1], vref[:, :, :pos + 1])
    ref = torch.nn.functional.scaled_dot_product_attention(
        qr, kref[:, :, :pos + 1].float(), vref[:, :, :pos + 1].float(), enable_gqa=True)
    assert (out.float() - ref.reshape(1, -1)).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3


def test_fused_decode_step_matches_unfused_engine():
    from quip_for_all_b200.modeling import LlamaDecodeEngine, make_random_quantized_llama
    model = make_random_quantized_llama("tiny128", "E8P12", seed=3, device=DEV)
    ids = torch.randint(0, 32000, (1, 10), generator=torch.Generator().manual_seed(2)).to(DEV)
    e1 = LlamaDecodeEngine(model, max_cache_len=64, fused=True, persistent=False)
    assert e1.fused is not None
    e2 = LlamaDecodeEngine(model, max_cache_len=64, fused=False, use_cuda_graph=False)
    e1.prefill(ids)
    e2.prefill(ids)
    assert torch.equal(e1.tok, e2.tok)
    # one step from identical state: hidden states agree to fp16 noise
    with torch.no_grad():
        h1 = e1.model.model.embed_tokens(e1.tok).view(1, -1)
        h2 = h1.clone()
        for li in range(len(e1.layers)):
            h1 = e1._layer_fused(li, h1)
        cos = e2.cos.index_select(0, e2.pos)[None, None]
        sin = e2.sin.index_select(0, e2.pos)[None, None]
        mask = (e2.arange[None, :] <= e2.pos[:, None])[None, None]
        h2 = h2.view(1, 1, -1)
        for li in range(len(e2.layers)):
            h2 = e2._layer(li, h2, e2.pos, cos, sin, mask)
    d = (h1.float() - h2.view(1, -1).float()).abs().max().item()
    assert d <= 2.0 ** -6 * h2.float().abs().max().item(), d
    out = e1.generate(ids, 6)          # graph-captured fused path runs end to end
    assert out.shape == (1, 6)


# ------------------------------------------------------------------------------------------------
# persistent whole-step kernel (decode_step.cu) against the per-group launches it replaces
# ------------------------------------------------------------------------------------------------
@pytest.fixture
def kv_splits(request):
    from quip_for_all_b200.decode_step import _bind
    L = _bind()
    assert L.quipb200_decode_step_set_splits(request.param) == 0
    yield request.param
    L.quipb200_decode_step_set_splits(0)


@pytest.mark.parametrize("kv_splits", [0, 4], indirect=True)
@pytest.mark.parametrize("name,n_layers", [("tiny256", 2), ("tinypow2", 3), ("llama2-7b", 2)])
def test_persistent_decode_step_matches_grouped_engine(name, n_layers, kv_splits):
    from quip_for_all_b200.modeling import LlamaDecodeEngine, llama_config, make_random_quantized_llama
    cfg = llama_config(name, num_hidden_layers=n_layers)
    model = make_random_quantized_llama(cfg, "E8P12", seed=5, device=DEV)
    e1 = LlamaDecodeEngine(model, max_cache_len=96, persistent=True, use_cuda_graph=False)
    assert e1.persistent is not None
    e2 = LlamaDecodeEngine(model, max_cache_len=96, persistent=False, use_cuda_graph=False)
    assert e2.persistent is None and e2.fused is not None
    g = torch.Generator().manual_seed(11)
    for plen in (1, 3, 37):
        ids = torch.randint(0, 32000, (1, plen), generator=g).to(DEV)
        e1.prefill(ids)
        e2.prefill(ids)
        assert torch.equal(e1.tok, e2.tok)
        for _ in range(3):
            with torch.no_grad():
                h = model.model.embed_tokens(e1.tok).view(1, -1).contiguous()
                h1 = e1.persistent(h, e1.h_step_out).clone()
                h2 = h.clone()
                for li in range(len(e2.layers)):
                    h2 = e2._layer_fused(li, h2)
            torch.cuda.synchronize()
            ref = h2.float()
            d = (h1.float() - ref).abs().max().item()
            # same device functions and rounding points; only fp32 summation order differs in the head slices
            assert d <= 2.0 ** -8 * ref.abs().max().item(), (name, plen, d, ref.abs().max().item())
            p = int(e1.pos.item())
            dk = (e1.k_cache[:, :, :, :p + 1].float() - e2.k_cache[:, :, :, :p + 1].float()).abs().max().item()
            dv = (e1.v_cache[:, :, :, :p + 1].float() - e2.v_cache[:, :, :, :p + 1].float()).abs().max().item()
            assert dk <= 2.0 ** -8 * e2.k_cache.float().abs().max().item(), (name, plen, dk)
            assert dv <= 2.0 ** -8 * e2.v_cache.float().abs().max().item(), (name, plen, dv)
            e1.pos.add_(1)
            e2.pos.add_(1)


def test_persistent_decode_step_under_cuda_graph_generates_same_tokens():
    from quip_for_all_b200.modeling import LlamaDecodeEngine, make_random_quantized_llama
    model = make_random_quantized_llama("tiny256", "E8P12", seed=3, device=DEV)
    ids = torch.randint(0, 32000, (1, 10), generator=torch.Generator().manual_seed(2)).to(DEV)
    e1 = LlamaDecodeEngine(model, max_cache_len=64, persistent=True)       # cooperative launch inside a CUDA graph
    assert e1.persistent is not None
    e2 = LlamaDecodeEngine(model, max_cache_len=64, persistent=False)
    o1 = e1.generate(ids, 12)
    o2 = e2.generate(ids, 12)
    assert torch.equal(o1[:, :4], o2[:, :4])      # later tokens may differ by an fp16 near-tie


def test_decode_engine_host_token_io_graph():
    """`step_host()`: the token id comes from / returns to pinned host memory inside ONE graph replay (what bench.py's e2e
    leg times); same tokens as the device-resident loop, and an overridden input token is what the step consumes."""
    from quip_for_all_b200.modeling import LlamaDecodeEngine, make_random_quantized_llama
    model = make_random_quantized_llama("tiny256", "E8P12", seed=3, device=DEV)
    ids = torch.randint(0, 32000, (1, 10), generator=torch.Generator().manual_seed(2)).to(DEV)
    e1 = LlamaDecodeEngine(model, max_cache_len=64, persistent=True)
    ref = e1.generate(ids, 9)[0].tolist()
    e2 = LlamaDecodeEngine(model, max_cache_len=64, persistent=True)
    e2.prefill(ids)
    e2.capture()
    first = int(e2.tok.item())
    got = [first] + [e2.step_host() for _ in range(8)]
    assert got == ref
    # feed a different token: the continuation must equal the device loop started from that token
    e3 = LlamaDecodeEngine(model, max_cache_len=64, persistent=True)
    e3.prefill(ids)
    e3.capture()
    e3.tok.fill_(123)
    e3.step()
    want = int(e3.tok.item())
    e4 = LlamaDecodeEngine(model, max_cache_len=64, persistent=True)
    e4.prefill(ids)
    assert e4.step_host(123) == want
    with pytest.raises(RuntimeError):
        for _ in range(100):
            e4.step_host()


# ------------------------------------------------------------------------------------------------
# tcgen05 decode + GEMM (umma_gemm.cu): 17 <= M <= 256
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M", [4, 9, 16, 17, 32, 33, 64, 100, 128, 200, 256])
@pytest.mark.parametrize("N,K", [(128, 128), (512, 4096), (4096, 4096), (1408, 1024 + 128)])
def test_e8p_mm_umma_matches_oracle(M, N, K):
    from quip_for_all_b200 import _native
    _native.set_option("umma", 2 if M <= 32 else 1)      # auto policy takes 4 <= M <= 32; force it above
    g = torch.Generator().manual_seed(M * 11 + N + K)
    q = torch.randint(-32768, 32768, (N, K // 8), generator=g).to(torch.int16)
    x = torch.randn(M, K, generator=g).half()
    lc0 = _native.launch_count()
    qd, xd = q.to(DEV), x.to(DEV)
    try:
        out = torch.ops.quip_lib.e8p_mm_origorder(xd, qd, _grid())
        out2 = torch.ops.quip_lib.e8p_mm_origorder(xd, qd, _grid())      # split-K workspace is self-cleaning
    finally:
        _native.set_option("umma", 2)
    assert _native.launch_count() - lc0 == 2                          # one launch of ours per call: no dense path
    assert out.shape == (M, N) and out.dtype == torch.float16
    W = qo.decompress_e8p(q.numpy())
    _mm_check(out, x, W)
    _mm_check(out2, x, W)
    # against the dense route on the same device (decompress + cuBLAS): both fp32-accumulate, fp16 out
    _native.set_option("umma", 0)
    try:
        other = torch.ops.quip_lib.e8p_mm_origorder(xd, qd, _grid())      # dp4a GEMV (M <= 16) or decompress + cuBLAS
    finally:
        _native.set_option("umma", 2)
    dense = other
    d = (out.float() - dense.float()).abs().max().item()
    assert d <= 2.0 ** -9 * dense.float().abs().max().item() + 1e-6, d


@pytest.mark.parametrize("M", [4, 8, 16, 23, 31, 32])
@pytest.mark.parametrize("cb", ["E8P12RVQ4B", "D4"])
@pytest.mark.parametrize("N,K", [(4096, 4096), (11008, 4096), (4096, 11008 + 128 - 11008 % 128)])
def test_rvq4_d4_small_m_tcgen05_route(cb, M, N, K):
    """4 <= M <= 32 for the RVQ4B / D4 mm ops (the reference's K2 / K3 range, origin_order.cu:337-385, :143-168): ONE launch
    of the codebook-templated tcgen05 kernel (codes decoded once for all rows) against the oracle and the dense route."""
    from quip_for_all_b200 import _native
    from quip_for_all_b200.codebook.d4 import build_D4_CB
    g = torch.Generator().manual_seed(M * 7 + N + K)
    x = torch.randn(M, K, generator=g).half()
    xd = x.to(DEV)
    lc0 = _native.launch_count()
    if cb == "D4":
        q = torch.randint(0, 256, (N, K // 4), generator=g).to(torch.uint8)
        grid = build_D4_CB().half().to(DEV)
        call = lambda: torch.ops.quip_lib.d4_mm_origorder(xd, q.to(DEV), grid)
        W = qo.decompress_d4(q.numpy())
    else:
        q = torch.randint(-2**31, 2**31, (N, K // 8), generator=g, dtype=torch.int64).to(torch.int32)
        call = lambda: torch.ops.quip_lib.e8prvq4_mm_origorder(xd, q.to(DEV), _grid(), 1 / 3.45)
        W = qo.decompress_e8prvq4(q.numpy(), 1 / 3.45)
    out = call()
    out2 = call()
    assert _native.launch_count() - lc0 == 2                  # one launch of ours per call
    _mm_check(out, x, W)
    _mm_check(out2, x, W)
    _native.set_option("umma", 0)
    try:
        dense = call()                                        # per-row dp4a GEMV / decompress + cuBLAS
    finally:
        _native.set_option("umma", 2)
    d = (out.float() - dense.float()).abs().max().item()
    assert d <= 2.0 ** -9 * dense.float().abs().max().item() + 1e-6, d


def test_e8p_mm_umma_unsupported_shapes_take_dense_path():
    g = torch.Generator().manual_seed(5)
    q = torch.randint(-32768, 32768, (96, 32), generator=g).to(torch.int16)      # N % 128 != 0
    x = torch.randn(40, 256, generator=g).half()
    out = torch.ops.quip_lib.e8p_mm_origorder(x.to(DEV), q.to(DEV), _grid())
    _mm_check(out, x, qo.decompress_e8p(q.numpy()))


# ------------------------------------------------------------------------------------------------
# batched fused rotation (rotate_batched.cu) and the M >= 17 forward built on it
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,K,fin,fout", [(4096, 1, 4096, 4096), (4096, 1, 4000, 3968), (11008, 43, 11008, 11008),
                                          (2816, 11, 2816, 2800), (256, 1, 256, 256)])
@pytest.mark.parametrize("M", [1, 17, 300])
def test_rotate_fused_matches_unfused_op_sequence(n, K, fin, fout, M):
    """One-pass rotation == the reference's pass-per-op sequence (x*pre -> pad -> hadamard -> hadK @ -> slice -> *post
    -> +bias) built from the individual quip_lib ops, to fp16 rounding of the last step."""
    import math
    from quip_for_all_b200.quant import matmul_hadU_cuda
    g = torch.Generator().manual_seed(n + K + M)
    x = torch.randn(M, fin, generator=g).half().to(DEV)
    pre = (1 + 0.1 * torch.randn(fin, generator=g)).half().to(DEV)
    post = (1 + 0.1 * torch.randn(fout, generator=g)).half().to(DEV)
    bias = (0.1 * torch.randn(fout, generator=g)).half().to(DEV)
    hadK = None
    hk = None
    if K > 1:
        qm, _ = torch.linalg.qr(torch.randn(K, K, generator=g))
        hadK = qm.half().to(DEV)
        Kp = (K + 15) // 16 * 16
        hk = torch.zeros(Kp, Kp, dtype=torch.float16, device=DEV)
        hk[:K, :K] = hadK
    scale = 0.37 / math.sqrt(n // K)
    y = torch.ops.quip_lib.rotate_fused(x, pre, hk, post, bias, n, K, fout, scale)
    ref = matmul_hadU_cuda(x * pre, hadK, K, n, scale=0.37)[..., :fout] * post + bias
    assert y.shape == ref.shape
    d = (y.float() - ref.float()).abs().max().item()
    assert d <= 2.0 ** -9 * ref.float().abs().max().item(), (d, ref.float().abs().max().item())
    # no pre / post / bias: plain rotation
    y2 = torch.ops.quip_lib.rotate_fused(x, None, hk, None, None, n, K, n, scale)
    ref2 = matmul_hadU_cuda(x, hadK, K, n, scale=0.37)
    assert (y2.float() - ref2.float()).abs().max().item() <= 2.0 ** -9 * ref2.float().abs().max().item()


@pytest.mark.parametrize("fin,fout", [(4096, 4096), (4000, 3968)])
@pytest.mark.parametrize("M,force", [(2301, False), (5, True), (148 * 12 * 2 + 3, False)])
def test_rotate_fused_warp_per_row_kernel(fin, fout, M, force):
    """The many-rows variant of the n = 4096 rotation (one warp per row, rotate_batched.cu::rot4096w_kernel) against the
    CTA-per-row kernel (bit-identical inputs; both round once to fp16) and the unfused op sequence."""
    import math
    from quip_for_all_b200 import _native
    from quip_for_all_b200.quant import matmul_hadU_cuda
    g = torch.Generator().manual_seed(M + fin)
    x = torch.randn(M, fin, generator=g).half().to(DEV)
    pre = (1 + 0.1 * torch.randn(fin, generator=g)).half().to(DEV)
    post = (1 + 0.1 * torch.randn(fout, generator=g)).half().to(DEV)
    bias = (0.1 * torch.randn(fout, generator=g)).half().to(DEV)
    scale = 0.37 / math.sqrt(4096)
    default = _native.get_option("rot_warp_rows")
    try:
        _native.set_option("rot_warp_rows", 1 if force else default)
        y_w = torch.ops.quip_lib.rotate_fused(x, pre, None, post, bias, 4096, 1, fout, scale)
        y_w2 = torch.ops.quip_lib.rotate_fused(x, None, None, None, None, 4096, 1, 4096, scale)
        _native.set_option("rot_warp_rows", 1 << 30)
        y_c = torch.ops.quip_lib.rotate_fused(x, pre, None, post, bias, 4096, 1, fout, scale)
        y_c2 = torch.ops.quip_lib.rotate_fused(x, None, None, None, None, 4096, 1, 4096, scale)
    finally:
        _native.set_option("rot_warp_rows", default)
    ref = matmul_hadU_cuda(x * pre, None, 1, 4096, scale=0.37)[..., :fout] * post + bias
    tol = 2.0 ** -9 * ref.float().abs().max().item()
    assert y_w.shape == ref.shape and y_w2.shape == (M, 4096)
    assert (y_w.float() - ref.float()).abs().max().item() <= tol
    assert (y_w.float() - y_c.float()).abs().max().item() <= tol
    assert (y_w2.float() - y_c2.float()).abs().max().item() <= tol
    # the two kernels differ only in fp32 summation order before the single fp16 rounding: all but a few elements equal
    assert (y_w != y_c).float().mean().item() < 0.02


@pytest.mark.parametrize("n,K,fin,fout", [(11008, 43, 11008, 11008), (2816, 11, 2816, 2800), (2816, 11, 2808, 2816),
                                          (4096, 16, 4096, 4096), (5120, 20, 5120, 5112)])
@pytest.mark.parametrize("M,force", [(1500, False), (3, True), (597, True)])
def test_rotate_fused_row_streaming_kernel(n, K, fin, fout, M, force):
    """The many-rows variant of the K x 256 rotation (persistent CTAs, rows streamed through shared memory with bulk
    copies, rotate_batched.cu::rotblk_pipe_kernel) against the CTA-per-row kernel -- same arithmetic in the same order,
    so bit-identical -- and against the unfused op sequence."""
    import math
    from quip_for_all_b200 import _native
    from quip_for_all_b200.quant import matmul_hadU_cuda
    g = torch.Generator().manual_seed(M + n + fin)
    x = torch.randn(M, fin, generator=g).half().to(DEV)
    pre = (1 + 0.1 * torch.randn(fin, generator=g)).half().to(DEV)
    post = (1 + 0.1 * torch.randn(fout, generator=g)).half().to(DEV)
    bias = (0.1 * torch.randn(fout, generator=g)).half().to(DEV)
    qm, _ = torch.linalg.qr(torch.randn(K, K, generator=g))
    hadK = qm.half().to(DEV)
    Kp = (K + 15) // 16 * 16
    hk = torch.zeros(Kp, Kp, dtype=torch.float16, device=DEV)
    hk[:K, :K] = hadK
    scale = 0.37 / math.sqrt(n // K)
    default = _native.get_option("rot_pipe_rows")
    try:
        _native.set_option("rot_pipe_rows", 1 if force else default)
        y_p = torch.ops.quip_lib.rotate_fused(x, pre, hk, post, bias, n, K, fout, scale)
        y_p2 = torch.ops.quip_lib.rotate_fused(x, None, hk, None, None, n, K, n, scale)
        torch.cuda.synchronize()
        _native.set_option("rot_pipe_rows", 1 << 30)
        y_c = torch.ops.quip_lib.rotate_fused(x, pre, hk, post, bias, n, K, fout, scale)
        y_c2 = torch.ops.quip_lib.rotate_fused(x, None, hk, None, None, n, K, n, scale)
    finally:
        _native.set_option("rot_pipe_rows", default)
    assert y_p.shape == (M, fout) and y_p2.shape == (M, n)
    assert torch.equal(y_p, y_c)
    assert torch.equal(y_p2, y_c2)
    ref = matmul_hadU_cuda(x * pre, hadK, K, n, scale=0.37)[..., :fout] * post + bias
    assert (y_p.float() - ref.float()).abs().max().item() <= 2.0 ** -9 * ref.float().abs().max().item()


@pytest.mark.parametrize("fin,fout,bias", [(4096, 11008, True), (11008, 4096, False), (4096, 4096, True)])
@pytest.mark.parametrize("M", [17, 40, 300])
def test_batched_forward_vs_oracle(fin, fout, bias, M):
    layer = make_layer(fin, fout, "E8P12", bias=bias, seed=fin + M, device=DEV)
    assert layer._batched_fused_ok(torch.empty(M, fin, dtype=torch.float16, device=DEV))
    x = torch.randn(M, fin, generator=torch.Generator().manual_seed(M)).half()
    with torch.no_grad():
        y = layer(x.to(DEV))
    ref = oracle_forward(layer, x, rounding="reference")
    err = np.abs(y.float().cpu().numpy() - ref)
    assert err.max() <= tol_of(ref), (err.max(), tol_of(ref))
    ref64 = oracle_forward(layer, x, rounding="none")
    assert np.abs(y.float().cpu().numpy() - ref64).max() <= tol_of(ref64)


def _oracle_layer_params(lyr):
    def lin(m):
        return dict(W_hat=qo.decompress_e8p(m.Qidxs.cpu().numpy()), in_features=m.in_features, out_features=m.out_features,
                    q_in=m.q_in_features, q_out=m.q_out_features,
                    SU=None if m.SU is None else m.SU.detach().float().cpu().numpy(),
                    SV=None if m.SV is None else m.SV.detach().float().cpu().numpy(),
                    bias=None if m.bias is None else m.bias.float().cpu().numpy(), wscale_float=m.wscale_float,
                    had_left=None if m.had_left is None else m.had_left.float().cpu().numpy(), K_left=m.K_left,
                    had_right=None if m.had_right is None else m.had_right.float().cpu().numpy(), K_right=m.K_right)
    at, mlp = lyr.self_attn, lyr.mlp
    return dict(input_norm=lyr.input_layernorm.weight.float().cpu().numpy(),
                post_norm=lyr.post_attention_layernorm.weight.float().cpu().numpy(),
                q=lin(at.q_proj), k=lin(at.k_proj), v=lin(at.v_proj), o=lin(at.o_proj),
                gate=lin(mlp.gate_proj), up=lin(mlp.up_proj), down=lin(mlp.down_proj))


@pytest.mark.parametrize("name", ["tiny256", "tinypow2"])
def test_persistent_decode_step_matches_cpu_oracle(name):
    """The persistent whole-step kernel against the CPU oracle of the decode loop body (oracle/quip_oracle.py:
    llama_decoder_layer_step): hidden state after every layer's worth of work and the appended K/V rows."""
    from quip_for_all_b200.modeling import LlamaDecodeEngine, llama_config, make_random_quantized_llama
    cfg = llama_config(name, num_hidden_layers=2)
    model = make_random_quantized_llama(cfg, "E8P12", seed=9, device=DEV)
    eng = LlamaDecodeEngine(model, max_cache_len=48, persistent=True, use_cuda_graph=False)
    assert eng.persistent is not None
    ids = torch.randint(0, 32000, (1, 5), generator=torch.Generator().manual_seed(4)).to(DEV)
    eng.prefill(ids)
    nh, nkv, hd = cfg.num_attention_heads, cfg.num_key_value_heads, eng.hd
    params = [_oracle_layer_params(l) for l in model.model.layers]
    for _ in range(2):
        pos = int(eng.pos.item())
        kc = [eng.k_cache[i, 0].float().cpu().numpy().astype(np.float64) for i in range(2)]
        vc = [eng.v_cache[i, 0].float().cpu().numpy().astype(np.float64) for i in range(2)]
        with torch.no_grad():
            h = model.model.embed_tokens(eng.tok).view(1, -1).contiguous()
            got = eng.persistent(h, eng.h_step_out).float().cpu().numpy()[0]
        torch.cuda.synchronize()
        ref = h.float().cpu().numpy()[0].astype(np.float64)
        for i in range(2):
            ref = qo.llama_decoder_layer_step(ref, params[i], kc[i], vc[i], pos, n_heads=nh, n_kv_heads=nkv, head_dim=hd,
                                              eps=cfg.rms_norm_eps)
        tol = 2.0 ** -6 * np.abs(ref).max()        # two layers of chained fp16 rounding points (cf. the unfused-engine test)
        assert np.abs(got - ref).max() <= tol, (np.abs(got - ref).max(), tol)
        for i in range(2):                          # the appended rows of the KV cache
            gk = eng.k_cache[i, 0, :, pos].float().cpu().numpy()
            gv = eng.v_cache[i, 0, :, pos].float().cpu().numpy()
            assert np.abs(gk - kc[i][:, pos]).max() <= 2.0 ** -7 * np.abs(kc[i][:, pos]).max() + 1e-3
            assert np.abs(gv - vc[i][:, pos]).max() <= 2.0 ** -7 * np.abs(vc[i][:, pos]).max() + 1e-3
        eng.pos.add_(1)


def test_persistent_decode_step_matches_cpu_oracle_at_llama2_7b_dims():
    """The headline kernel at BASELINE config 2's layer shapes (hidden 4096, intermediate 11008 = 43 x 256, 32 heads),
    two decoder layers, DIRECTLY against the CPU oracle of the decode loop body (not via the grouped launches)."""
    from quip_for_all_b200.modeling import LlamaDecodeEngine, llama_config, make_random_quantized_llama
    cfg = llama_config("llama2-7b", num_hidden_layers=2)
    model = make_random_quantized_llama(cfg, "E8P12", seed=21, device=DEV)
    eng = LlamaDecodeEngine(model, max_cache_len=160, persistent=True, use_cuda_graph=False)
    assert eng.persistent is not None
    ids = torch.randint(0, 32000, (1, 131), generator=torch.Generator().manual_seed(5)).to(DEV)
    eng.prefill(ids)
    nh, nkv, hd = cfg.num_attention_heads, cfg.num_key_value_heads, eng.hd
    params = [_oracle_layer_params(l) for l in model.model.layers]
    for _ in range(2):
        pos = int(eng.pos.item())
        kc = [eng.k_cache[i, 0].float().cpu().numpy().astype(np.float64) for i in range(2)]
        vc = [eng.v_cache[i, 0].float().cpu().numpy().astype(np.float64) for i in range(2)]
        with torch.no_grad():
            h = model.model.embed_tokens(eng.tok).view(1, -1).contiguous()
            got = eng.persistent(h, eng.h_step_out).float().cpu().numpy()[0]
        torch.cuda.synchronize()
        ref = h.float().cpu().numpy()[0].astype(np.float64)
        for i in range(2):
            ref = qo.llama_decoder_layer_step(ref, params[i], kc[i], vc[i], pos, n_heads=nh, n_kv_heads=nkv, head_dim=hd,
                                              eps=cfg.rms_norm_eps)
        tol = 2.0 ** -7 * np.abs(ref).max()
        assert np.abs(got - ref).max() <= tol, (np.abs(got - ref).max(), tol)
        for i in range(2):
            gk = eng.k_cache[i, 0, :, pos].float().cpu().numpy()
            gv = eng.v_cache[i, 0, :, pos].float().cpu().numpy()
            assert np.abs(gk - kc[i][:, pos]).max() <= 2.0 ** -7 * np.abs(kc[i][:, pos]).max() + 1e-3
            assert np.abs(gv - vc[i][:, pos]).max() <= 2.0 ** -7 * np.abs(vc[i][:, pos]).max() + 1e-3
        eng.pos.add_(1)


@pytest.mark.parametrize("cb", ["E8P12RVQ4B", "D4"])
def test_persistent_decode_step_other_codebooks(cb):
    """BASELINE config 4 codebooks through the persistent kernel (two integer accumulators + fp16 residual scale for
    RVQ4B; 1-byte codes and the half-unit table for D4) against the per-linear launches."""
    from quip_for_all_b200.modeling import LlamaDecodeEngine, llama_config, make_random_quantized_llama
    model = make_random_quantized_llama(llama_config("tiny256", num_hidden_layers=2), cb, seed=6, device=DEV)
    e1 = LlamaDecodeEngine(model, max_cache_len=64, persistent=True, use_cuda_graph=False)
    assert e1.persistent is not None
    e2 = LlamaDecodeEngine(model, max_cache_len=64, persistent=False, use_cuda_graph=False)
    ids = torch.randint(0, 32000, (1, 7), generator=torch.Generator().manual_seed(8)).to(DEV)
    e1.prefill(ids)
    e2.prefill(ids)
    for _ in range(2):
        with torch.no_grad():
            h = model.model.embed_tokens(e1.tok).view(1, -1).contiguous()
            h1 = e1.persistent(h, e1.h_step_out).clone()
            h2 = h.clone()
            for li in range(len(e2.layers)):
                h2 = e2._layer_fused(li, h2)
        torch.cuda.synchronize()
        d = (h1.float() - h2.float()).abs().max().item()
        assert d <= 2.0 ** -8 * h2.float().abs().max().item(), (cb, d)
        e1.pos.add_(1)
        e2.pos.add_(1)


@pytest.mark.parametrize("hidden,vocab", [(4096, 32000), (256, 1000), (8192, 517)])
def test_lm_tail_matches_torch_norm_head_argmax(hidden, vocab):
    """Fused tail (csrc/lm_tail.cu): LlamaRMSNorm -> fp16 lm_head -> argmax -> position increment, against the torch ops
    HF runs.  Logits: one fp16 ulp (accumulation order differs from cuBLAS); the token must be a maximiser of the torch
    logits up to that ulp; exact ties go to the lowest index."""
    from quip_for_all_b200.fused import LmTail
    g = torch.Generator().manual_seed(hidden + vocab)
    W = (0.02 * torch.randn(vocab, hidden, generator=g)).half().to(DEV)
    nw = (1 + 0.1 * torch.randn(hidden, generator=g)).half().to(DEV)
    h = torch.randn(1, hidden, generator=g).half().to(DEV)
    tail = LmTail(nw, 1e-5, W)
    tok = torch.zeros(1, 1, dtype=torch.long, device=DEV)
    pos = torch.full((1,), 41, dtype=torch.long, device=DEV)
    logits = torch.empty(vocab, dtype=torch.float16, device=DEV)
    for rep in range(2):                                   # second launch: the arrival counter reset itself
        tail(h, tok, pos=pos, logits_out=logits)
    torch.cuda.synchronize()
    v = h.float()
    x = nw * (v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + 1e-5)).half()
    ref = (x.float() @ W.float().t()).half().view(-1)
    d = (logits.float() - ref.float()).abs().max().item()
    assert d <= 2.0 ** -10 * ref.float().abs().max().item() + 1e-6, d
    assert int(pos.item()) == 43
    t = int(tok.item())
    assert 0 <= t < vocab
    assert logits[t].item() == logits.max().item() and t == int(torch.nonzero(logits == logits.max())[0].item())
    assert ref[t].item() >= ref.max().item() - 2.0 ** -9 * abs(ref.max().item())
    # exact ties: duplicated rows -> the lowest index wins (torch.argmax)
    W2 = W.clone()
    W2[vocab - 1] = W2[t]
    W2[min(t + 3, vocab - 2)] = W2[t]
    tail2 = LmTail(nw, 1e-5, W2)
    tail2(h, tok)
    assert int(tok.item()) == t


def test_forward_accepts_sliced_and_offset_activation_views():
    """activation views that are not 16-byte aligned / have an odd row pitch work in the reference; the fast paths
    must take them too (copied once) instead of raising from the C ABI."""
    layer = make_layer(512, 256, "E8P12", bias=True, seed=13, device=DEV)
    big = torch.randn(24, 515, generator=torch.Generator().manual_seed(5)).half().to(DEV)
    with torch.no_grad():
        for M in (1, 3, 20):
            xv = big[:M, 3:]                   # offset by 6 bytes, row pitch 515 elements
            assert xv.data_ptr() % 16 != 0
            y = layer(xv)
            y_ref = layer(xv.contiguous())
            assert torch.equal(y, y_ref)
