"""Quantise-time path on CPU: the oracle's nearest-codeword search / LDLQ against vectors produced by the reference's
own Python (tests/golden/quantize.npz, tests/golden/gen_quantize_golden.py), and the product's host-side LDLQ driver
(`quip_for_all_b200.ldlq`, CPU tensors take the reference's torch rounding expression) against the same vectors."""
import os

import numpy as np
import pytest
import torch

import quip_oracle as qo


@pytest.fixture(scope="module")
def gq(golden_dir):
    return np.load(os.path.join(golden_dir, "quantize.npz"))


def _check_idx(x, got, want, max_near_ties=0.005):
    """Indices equal, except rows where both choices score within 1e-5 of each other (fp32 near-ties)."""
    got, want = np.asarray(got, dtype=np.int64), np.asarray(want, dtype=np.int64)
    diff = np.nonzero(got != want)[0]
    if diff.size:
        gap = np.abs(qo.e8p_score(x[diff], got[diff]) - qo.e8p_score(x[diff], want[diff]))
        assert gap.max() < 1e-5, f"{diff.size} rows differ, worst score gap {gap.max()}"
        assert diff.size <= max_near_ties * got.size
    return diff.size


def test_oracle_nearest_matches_reference_quantize(gq):
    x = gq["nearest_x"]
    idx, _ = qo.e8p_nearest(x)
    _check_idx(x, idx, gq["nearest_idx"])
    assert np.array_equal(idx[1464:1528], gq["nearest_idx"][1464:1528])      # exact codewords: unique maximum
    assert np.array_equal(idx[1528:], gq["nearest_idx"][1528:])              # all-zero rows: first index of the tie
    np.testing.assert_array_equal(qo.e8p_full_grid()[idx[:1400]][idx[:1400] == gq["nearest_idx"][:1400]],
                                  gq["nearest_vals"][:1400][idx[:1400] == gq["nearest_idx"][:1400]])


def test_oracle_rvq4_quantize_matches_reference(gq):
    x = gq["nearest_x"]
    vals, idx, _ = qo.e8prvq4_quantize(x, float(gq["rvq4_resid_scale"]))
    same = idx == gq["nearest_rvq4_idx"]
    assert same.mean() > 0.995
    np.testing.assert_array_equal(vals[same], gq["nearest_rvq4_vals"][same])


def test_oracle_ldlq_matches_reference(gq):
    W, H = gq["ldlq_W"], gq["ldlq_H"]
    L = np.linalg.cholesky(H)
    hat, Q = qo.ldlq(W, H, L)
    assert np.array_equal(Q.astype(np.int16), gq["ldlq_f64_Q"])
    np.testing.assert_allclose(hat, gq["ldlq_f64_hat"], rtol=0, atol=0)
    assert bool(gq["ldlq_f64_buffered_same"])
    _, Qt = qo.ldlq(W, H, L, tune_iters=1)
    assert np.array_equal(Qt.astype(np.int16), gq["ldlq_f64_tune1_Q"])


def test_product_ldlq_cpu_matches_reference(gq):
    from quip_for_all_b200 import codebook_id
    from quip_for_all_b200.ldlq import ldlq, proxy_loss
    cb = codebook_id["E8P12"](inference=False)
    cb.grid = cb.grid.double()
    cb.grid_norm = cb.grid_norm.double()
    W, H = torch.from_numpy(gq["ldlq_W"]), torch.from_numpy(gq["ldlq_H"])
    L = torch.linalg.cholesky(H)
    for bc in (128, 64, 8, 256):                       # the blocking of the sweep does not change the result
        hat, Q = ldlq(W, H, L, cb, 0, block_cols=bc)
        assert np.array_equal(Q.numpy(), gq["ldlq_f64_Q"]), bc
        assert np.array_equal(hat.numpy(), gq["ldlq_f64_hat"])
    _, Qt = ldlq(W, H, L, cb, 1)
    assert np.array_equal(Qt.numpy(), gq["ldlq_f64_tune1_Q"])
    # LDLQ beats plain nearest rounding on the proxy objective it minimises
    near = cb.quantize(W.reshape(-1, 8), return_idx=False).reshape(W.shape)
    assert proxy_loss(W, hat, H) < proxy_loss(W, near, H)


def test_product_layer_quantizer_cpu_matches_reference(gq):
    """`LayerQuantizer` (reference: quip.py QUIP.add_batch / .quant) on the reference's layer, calibration batches and
    sign vectors: same packed indices, scale and de-rotated weight."""
    from quip_for_all_b200 import codebook_id
    from quip_for_all_b200.ldlq import LayerQuantizer
    lin = torch.nn.Linear(256, 64, bias=True)
    lin.weight.data = torch.from_numpy(gq["quip_w"]).clone()
    lin.bias.data = torch.from_numpy(gq["quip_bias"]).clone()
    cb = codebook_id["E8P12"](inference=False)
    cb.grid = cb.grid.double()
    cb.grid_norm = cb.grid_norm.double()
    lq = LayerQuantizer(lin, cb)
    calib = torch.from_numpy(gq["quip_calib"])
    for b in range(calib.shape[0]):
        lq.add_batch(calib[b])
    attr = lq.quantize(use_fp64=True, sigma_reg=0.01, SU=torch.from_numpy(gq["quip_SU"]), SV=torch.from_numpy(gq["quip_SV"]))
    assert np.array_equal(attr["Qidxs"].numpy(), gq["quip_Qidxs"])
    assert abs(float(attr["w_scale"]) - float(gq["quip_w_scale"])) < 1e-12 * float(gq["quip_w_scale"]) + 1e-15
    np.testing.assert_allclose(lin.weight.data.numpy(), gq["quip_w_hat"], rtol=0, atol=1e-7)
    assert attr["merge_su"] and attr["merge_sv"] and attr["left_hadK"] is None and attr["scaleWH"] is None
    assert 0 < lq.last_proxy_loss < 0.2


def _tiny_llama(dtype=torch.float32):
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(0)
    cfg = LlamaConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                      num_key_value_heads=2, vocab_size=100, max_position_embeddings=64)
    return LlamaForCausalLM(cfg).eval().to(dtype)


def test_quantize_model_save_load_round_trip_cpu(tmp_path):
    """`QuipQuantizer.quantize_model` (reference: quantizer.py:250-600, without the fine-tuning stage) on a tiny Llama with
    ready calibration batches, `save`, then `load_quantized_model`'s body: same module tree, same tensors, the
    reference's state-dict key layout."""
    from quip_for_all_b200 import QuantLinear, QuipQuantizer
    from quip_for_all_b200.quantizer import _load_quantized_model
    m = _tiny_llama()
    qz = QuipQuantizer("E8P12", quip_tune_iters=0, ft_epochs=0, inference=False)
    calib = [torch.randint(0, 100, (2, 16), generator=torch.Generator().manual_seed(i)) for i in range(3)]
    m = qz.quantize_model(m, calib, save_dir=str(tmp_path))
    qls = {n: x for n, x in m.named_modules() if isinstance(x, QuantLinear)}
    assert len(qls) == 14 and not any(isinstance(x, torch.nn.Linear) for n, x in m.named_modules() if "layers" in n)
    for n, q in qls.items():
        assert q.Wscale.dtype == torch.float32 and abs(q.wscale_float - float(q.Wscale)) < 1e-9
        assert 0 < q.proxy_loss < 0.2 and q.Qidxs.abs().max() > 0
    assert sorted(os.listdir(tmp_path)) == ["config.json", "pytorch_model.bin", "quantization_config.json"]
    sd = torch.load(tmp_path / "pytorch_model.bin")
    pre = "model.layers.0.self_attn.q_proj."
    assert {k[len(pre):] for k in sd if k.startswith(pre)} == {"SU", "SV", "Qidxs", "Wscale", "weight"}
    m2 = _load_quantized_model(str(tmp_path), torch_dtype=torch.float32)
    sd1, sd2 = m.state_dict(), m2.state_dict()
    assert sd1.keys() == sd2.keys()
    for k in sd1:
        assert torch.equal(sd1[k], sd2[k]), k
    with pytest.raises(NotImplementedError):
        QuipQuantizer("E8P12", ft_epochs=2, inference=False).quantize_model(_tiny_llama(), calib)
    with pytest.raises(ValueError):
        QuipQuantizer("E8P12", ft_epochs=0, inference=False).quantize_model(_tiny_llama(), "wikitext2")


def test_oracle_rvq3_quantize_matches_reference(gq):
    x = gq["nearest_x"]
    vals, idx = qo.e8prvq3_quantize(x, float(gq["rvq3_resid_scale"]))
    same = idx == gq["nearest_rvq3_idx"]
    assert same.mean() > 0.995
    np.testing.assert_array_equal(vals[same], gq["nearest_rvq3_vals"][same])


def test_product_layer_quantizer_rescale_per_channel_tune_matches_reference(gq):
    """second driver case of the reference (quip.py QUIP.quant): W/H rescaling, per-channel scales, a random orthogonal
    3 x 3 block on the 48-row output side (same numpy seed -> same draw), one re-rounding sweep, unbuffered LDLQ."""
    from quip_for_all_b200 import codebook_id
    from quip_for_all_b200.ldlq import LayerQuantizer
    lin = torch.nn.Linear(128, 48, bias=False)
    lin.weight.data = torch.from_numpy(gq["quip2_w"]).clone()
    cb = codebook_id["E8P12"](inference=False)
    cb.grid = cb.grid.double()
    cb.grid_norm = cb.grid_norm.double()
    lq = LayerQuantizer(lin, cb)
    calib = torch.from_numpy(gq["quip2_calib"])
    for b in range(calib.shape[0]):
        lq.add_batch(calib[b])
    np.random.seed(3)
    attr = lq.quantize(rescale_WH=True, use_fp64=True, sigma_reg=0.01, per_channel=True, quip_tune_iters=1,
                       SU=torch.from_numpy(gq["quip2_SU"]), SV=torch.from_numpy(gq["quip2_SV"]))
    np.testing.assert_array_equal(attr["right_hadK"].double().numpy(), gq["quip2_right_hadK"])
    np.testing.assert_allclose(attr["scaleWH"].double().numpy(), gq["quip2_scaleWH"], rtol=1e-6)
    np.testing.assert_allclose(attr["w_scale"].double().numpy(), gq["quip2_w_scale"], rtol=1e-12)
    same_rows = (attr["Qidxs"].numpy() == gq["quip2_Qidxs"]).all(1).mean()
    assert same_rows == 1.0, same_rows
    np.testing.assert_allclose(lin.weight.data.numpy(), gq["quip2_w_hat"], rtol=0, atol=2e-7)
    assert attr["left_hadK"] is None and attr["w_scale"].shape == (48, 1)


def test_small_codebooks_and_index_packing_match_reference(gq):
    """D4 / HI `quantize` and the packed index formats of HI (8 nibbles per int32) and E8P12RVQ3B (3 bytes per code)
    against outputs of the reference's codebook classes."""
    from quip_for_all_b200 import codebook_id
    d = codebook_id["D4"](inference=False)
    vals, idx = d.quantize(torch.from_numpy(gq["d4_x"]))
    assert np.array_equal(idx.numpy().astype(np.int64), gq["d4_idx"]) and np.array_equal(vals.numpy(), gq["d4_vals"])
    h = codebook_id["HI"](inference=False)
    vals, idx = h.quantize(torch.from_numpy(gq["hi_x"]))
    assert np.array_equal(idx.numpy().astype(np.int64), gq["hi_idx"]) and np.array_equal(vals.numpy(), gq["hi_vals"])
    assert np.array_equal(h.maybe_pack_idxs(torch.from_numpy(gq["hi_pack_in"])).numpy(), gq["hi_pack_out"])
    r3 = codebook_id["E8P12RVQ3B"](inference=True)
    assert np.array_equal(r3.maybe_pack_idxs(torch.from_numpy(gq["rvq3_pack_in"])).numpy(), gq["rvq3_pack_out"])


def test_structured_search_equals_brute_force(gq):
    """the 512-candidate search derived from the codebook's structure picks the brute-force argmax on generic inputs
    (random and outlier rows; exact-tie rows are excluded by construction of the method)"""
    x = gq["nearest_x"][:1464]
    i_bf, s_bf = qo.e8p_nearest(x)
    i_st, s_st = qo.e8p_nearest_structured(x)
    np.testing.assert_allclose(s_st, s_bf, rtol=0, atol=1e-9)
    diff = np.nonzero(i_st != i_bf)[0]
    assert diff.size == 0 or np.abs(qo.e8p_score(x[diff], i_st[diff]) - s_bf[diff]).max() < 1e-9
    assert (i_st[1400:] == i_bf[1400:]).all()                # outliers and exact codewords
