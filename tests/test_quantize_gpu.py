"""Quantise-time path on the GPU: the fused nearest-codeword kernel (csrc/nearest.cu) through the
C ABI against the reference's own outputs (tests/golden/quantize.npz), the float64 oracle and the reference's torch
expression evaluated on the same device; LDLQ and the per-layer driver on top of it; and the round trip
nn.Linear -> quantize -> packed QuantLinear.forward == dense linear with the de-rotated quantised weight."""
import os

import numpy as np
import pytest
import torch

import quip_oracle as qo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gq(golden_dir):
    return np.load(os.path.join(golden_dir, "quantize.npz"))


def _near_tie_ok(x, got, want, frac=0.005):
    got, want = np.asarray(got, dtype=np.int64), np.asarray(want, dtype=np.int64)
    diff = np.nonzero(got != want)[0]
    if diff.size:
        gap = np.abs(qo.e8p_score(x[diff], got[diff]) - qo.e8p_score(x[diff], want[diff]))
        assert gap.max() < 1e-5, f"{diff.size} rows differ, worst score gap {gap.max()}"
        assert diff.size <= frac * got.size
    return diff.size


def test_nearest_kernel_vs_reference_golden(gq):
    from quip_for_all_b200 import codebook_id
    dev = torch.device("cuda:0")
    cb = codebook_id["E8P12"](inference=False).to(dev)
    x = gq["nearest_x"]
    vals, idx = cb.quantize(torch.from_numpy(x).to(dev))
    assert idx.dtype == torch.int64 and vals.dtype == torch.float32 and vals.shape == (x.shape[0], 8)
    idx, vals = idx.cpu().numpy(), vals.cpu().numpy()
    _near_tie_ok(x, idx, gq["nearest_idx"])
    assert np.array_equal(idx[1464:], gq["nearest_idx"][1464:])          # exact codewords; zero rows (first index of a tie)
    np.testing.assert_array_equal(vals, qo.e8p_full_grid()[idx])         # vals are exactly the decoded codewords
    oidx, _ = qo.e8p_nearest(x)
    _near_tie_ok(x, idx, oidx)


def test_nearest_kernel_rvq4_vs_reference_golden(gq):
    from quip_for_all_b200 import codebook_id
    dev = torch.device("cuda:0")
    cb = codebook_id["E8P12RVQ4B"](inference=False).to(dev)
    x = gq["nearest_x"]
    vals, idx = cb.quantize(torch.from_numpy(x).to(dev))
    idx, vals = idx.cpu().numpy(), vals.cpu().numpy()
    same = idx == gq["nearest_rvq4_idx"]
    assert same.mean() > 0.99
    np.testing.assert_array_equal(vals[same], gq["nearest_rvq4_vals"][same])
    ovals, oidx, _ = qo.e8prvq4_quantize(x, float(gq["rvq4_resid_scale"]))
    osame = idx == oidx
    assert osame.mean() > 0.99
    np.testing.assert_array_equal(vals[osame], ovals[osame])
    # whatever the near-ties, the two-stage result is never worse than the golden's by more than fp32 noise
    err = ((vals - x) ** 2).sum(1)
    gerr = ((gq["nearest_rvq4_vals"] - x) ** 2).sum(1)
    assert (err <= gerr + 1e-4).all()


def test_nearest_kernel_rvq3_vs_reference_golden(gq):
    """E8P12RVQ3B: the second search runs against the 256-entry e81b table (csrc/nearest.cu, TABLE variant)"""
    from quip_for_all_b200 import codebook_id
    dev = torch.device("cuda:0")
    cb = codebook_id["E8P12RVQ3B"](inference=False).to(dev)
    x = gq["nearest_x"]
    vals, idx = cb.quantize(torch.from_numpy(x).to(dev))
    idx, vals = idx.cpu().numpy(), vals.cpu().numpy()
    same = idx == gq["nearest_rvq3_idx"]
    assert same.mean() > 0.99
    np.testing.assert_array_equal(vals[same], gq["nearest_rvq3_vals"][same])
    ovals, oidx = qo.e8prvq3_quantize(x, float(gq["rvq3_resid_scale"]))
    osame = idx == oidx
    assert osame.mean() > 0.99
    np.testing.assert_array_equal(vals[osame], ovals[osame])
    assert (idx >> 8).max() < 65536 and (idx & 0xff).max() < 256
    err = ((vals - x) ** 2).sum(1)
    gerr = ((gq["nearest_rvq3_vals"] - x) ** 2).sum(1)
    assert (err <= gerr + 1e-4).all()
    # ragged size against the torch expression of the same device
    xr = (torch.randn(777, 8, generator=torch.Generator().manual_seed(5)) * 1.1).to(dev)
    v1, i1 = cb.quantize(xr)
    iv, ii = cb.round(xr, cb.grid, cb.grid_norm)
    rr = (xr - iv) / cb.opt_resid_scale
    rv, ri = cb.round(rr, cb.e81b_grid, cb.e81b_grid_norm)
    assert ((i1 == (ii << 8) + ri).float().mean()) > 0.99


def test_structured_search_kernel_equals_brute_force_kernel(gq):
    """the default 512-candidate kernel against the kernel that evaluates all 65 536 codewords (option nearest_struct = 0)
    on the golden rows and on a large random batch; exact-tie rows (zeros) and exact codewords included"""
    from quip_for_all_b200 import _native, codebook_id
    dev = torch.device("cuda:0")
    cb = codebook_id["E8P12"](inference=False).to(dev)
    xs = [torch.from_numpy(gq["nearest_x"]), torch.randn(20000, 8, generator=torch.Generator().manual_seed(11)) * 1.3]
    for x in xs:
        xd = x.to(dev)
        assert _native.get_option("nearest_struct") == 1
        v1, i1 = cb.quantize(xd)
        _native.set_option("nearest_struct", 0)
        try:
            v0, i0 = cb.quantize(xd)
        finally:
            _native.set_option("nearest_struct", 1)
        _near_tie_ok(x.numpy(), i1.cpu().numpy(), i0.cpu().numpy(), frac=0.002)
        same = (i1 == i0)
        assert torch.equal(v1[same], v0[same])
    x = gq["nearest_x"]
    _, i1 = cb.quantize(torch.from_numpy(x).to(dev))
    assert np.array_equal(i1.cpu().numpy()[1464:], gq["nearest_idx"][1464:])


@pytest.mark.parametrize("m", [1, 7, 511, 513, 5000])
def test_nearest_kernel_vs_torch_expression_same_device(m):
    """ragged sizes (one vector, partial thread tiles, several CTAs in x) against `round` (codebook/e8p12.py:125-128)
    evaluated by torch on the same GPU"""
    from quip_for_all_b200 import codebook_id
    dev = torch.device("cuda:0")
    cb = codebook_id["E8P12"](inference=False).to(dev)
    g = torch.Generator().manual_seed(m)
    x = (torch.randn(m, 8, generator=g) * 1.2).to(dev)
    vals, idx = cb.quantize(x)
    rvals, ridx = cb.round(x, cb.grid, cb.grid_norm)
    _near_tie_ok(x.cpu().numpy(), idx.cpu().numpy(), ridx.cpu().numpy(), frac=0.01 if m > 100 else 1.0)
    # idempotence: a codeword quantises to itself
    v2, i2 = cb.quantize(vals)
    assert torch.equal(i2, idx) and torch.equal(v2, vals)
    # non-contiguous / offset views are accepted (LDLQ hands over column slices)
    big = torch.zeros(m, 24, device=dev)
    big[:, 8:16] = x
    v3, i3 = cb.quantize(big[:, 8:16])
    assert torch.equal(i3, idx)


def test_ldlq_on_gpu_vs_reference_golden(gq):
    from quip_for_all_b200 import codebook_id
    from quip_for_all_b200.ldlq import ldlq, proxy_loss
    dev = torch.device("cuda:0")
    cb = codebook_id["E8P12"](inference=False).to(dev)
    W = torch.from_numpy(gq["ldlq_W"]).float().to(dev)
    H = torch.from_numpy(gq["ldlq_H"]).float().to(dev)
    L = torch.linalg.cholesky(H)
    hat, Q = ldlq(W, H, L, cb, 0)
    Qn = Q.cpu().numpy()
    # rows are independent problems; an fp32 near-tie changes the rest of its row only
    rows_same_f32 = (Qn == gq["ldlq_f32_Q"]).all(1).mean()
    rows_same_f64 = (Qn == gq["ldlq_f64_Q"]).all(1).mean()
    assert max(rows_same_f32, rows_same_f64) >= 0.85, (rows_same_f32, rows_same_f64)
    W64, H64 = torch.from_numpy(gq["ldlq_W"]), torch.from_numpy(gq["ldlq_H"])
    ours = proxy_loss(W64, hat.double().cpu(), H64)
    ref = proxy_loss(W64, torch.from_numpy(gq["ldlq_f64_hat"]), H64)
    assert abs(ours - ref) < 0.02 * ref, (ours, ref)
    assert torch.equal(hat.cpu(), torch.from_numpy(qo.e8p_full_grid())[Q.cpu().long() & 0xffff].reshape(hat.shape))


@pytest.mark.parametrize("codebook,fin,fout", [("E8P12", 1024, 512), ("E8P12RVQ4B", 512, 256), ("E8P12", 1408, 512),
                                               ("E8P12RVQ3B", 512, 256)])
def test_quantize_linear_round_trip(codebook, fin, fout):
    """nn.Linear -> LayerQuantizer (GPU search kernel) -> packed QuantLinear: the inference path on the packed codes
    reproduces the dense linear whose weight is the de-rotated quantised matrix, and the quantised layer approximates
    the original one at the codebook's rate (1408 = 11 * 128: random-orthogonal 11 x 11 block on the input side)."""
    from quip_for_all_b200.ldlq import quantize_linear
    from helpers import oracle_forward
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    lin = torch.nn.Linear(fin, fout, bias=True).to(dev)
    w0 = lin.weight.data.clone()
    calib = [torch.randn(8, 256, fin, device=dev) for _ in range(3)]      # 6144 tokens: a full-rank proxy Hessian
    ql = quantize_linear(lin, calib, codebook=codebook).eval()
    x = torch.randn(5, fin, device=dev).half()
    with torch.no_grad():
        y_q = ql(x).float()
    y_hat = torch.nn.functional.linear(x.float(), lin.weight.data.float(), lin.bias.data.float())
    y_0 = torch.nn.functional.linear(x.float(), w0.float(), lin.bias.data.float())
    tol = 2.0 ** -7 * float(y_hat.abs().max())
    assert float((y_q - y_hat).abs().max()) < tol
    ref = oracle_forward(ql, x)
    assert float(np.abs(y_q.cpu().numpy() - ref.astype(np.float32)).max()) < 2.0 ** -8 * float(np.abs(ref).max())
    rel = float((y_hat - y_0).norm() / (y_0 - lin.bias.data.float()).norm())
    # relative error on iid weights: 2-bit ~0.30, 3-bit ~0.14, 4-bit ~0.07
    assert rel < {"E8P12": 0.40, "E8P12RVQ3B": 0.22, "E8P12RVQ4B": 0.12}[codebook], rel
    assert ql.proxy_loss < {"E8P12": 0.13, "E8P12RVQ3B": 0.05, "E8P12RVQ4B": 0.015}[codebook]


@pytest.mark.parametrize("codebook,min_corr", [("E8P12RVQ4B", 0.97), ("E8P12", 0.70)])
def test_quantize_model_end_to_end(tmp_path, codebook, min_corr):
    """tiny Llama: quantize_model on the GPU (search kernel inside LDLQ) -> logits through the CUDA inference path track the
    fp32 model's at the codebook's rate; save -> load_quantized_model reproduces the quantised model's logits exactly."""
    from transformers import AutoModelForCausalLM, LlamaConfig
    from quip_for_all_b200 import QuipQuantizer, load_quantized_model
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    cfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                      num_key_value_heads=2, vocab_size=128, max_position_embeddings=64)
    # built the way the loader builds it (fp16 parameters, computed buffers such as the rotary inv_freq stay fp32);
    # `.half()` on an fp32 model would round inv_freq and the two models would differ in RoPE, not in the linears
    m = AutoModelForCausalLM.from_config(cfg, dtype=torch.float16).eval().to(dev)
    ids = torch.randint(0, 128, (2, 24), device=dev)
    with torch.no_grad():
        ref = m(ids).logits.float()
    qz = QuipQuantizer(codebook, quip_tune_iters=0, ft_epochs=0, inference=False, opt_resid_scale=None)
    calib = [torch.randint(0, 128, (4, 32)) for _ in range(4)]
    m = qz.quantize_model(m, calib, save_dir=str(tmp_path))
    with torch.no_grad():
        got = m(ids).logits.float()
    assert torch.isfinite(got).all()
    a, b = (got - got.mean()).flatten(), (ref - ref.mean()).flatten()
    corr = float((a @ b) / (a.norm() * b.norm()))
    assert corr > min_corr, corr
    m2 = load_quantized_model(str(tmp_path)).to(dev).eval()
    with torch.no_grad():
        again = m2(ids).logits.float()
    assert torch.equal(again, got)
