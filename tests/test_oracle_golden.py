"""Pins the CPU oracle (oracle/quip_oracle.py) against the reference's own Python outputs
(tests/golden/*.npz, produced by tests/golden/gen_golden.py) and the SURVEY.md A.7 hashes."""
import hashlib
import os

import numpy as np
import pytest

import quip_oracle as qo


@pytest.fixture(scope="module")
def tables(golden_dir):
    return np.load(os.path.join(golden_dir, "tables.npz"))


@pytest.fixture(scope="module")
def had(golden_dir):
    return np.load(os.path.join(golden_dir, "hadamard.npz"))


@pytest.fixture(scope="module")
def ql(golden_dir):
    return np.load(os.path.join(golden_dir, "quantlinear.npz"))


def test_e8p_abs_table_matches_reference(tables):
    t = qo.e8p_abs_table()
    assert np.array_equal(t, tables["e8p_abs"])
    # SURVEY.md A.7 pins
    assert hashlib.sha256(t.astype("<i8").tobytes()).hexdigest() == \
        "81ec757eccfb81c367a30c2227a121386c2416dad5f1e16bbd1594181472111c"
    u = t.view(np.uint64)
    assert [int(v) for v in u[:4]] == [0x0202020202020202, 0xfa02020202020202, 0x0a02020202020202, 0xfe02060202020202]
    assert int(u[255]) == 0xfe06060602060206


def test_e8p_full_grid_bit_exact(tables):
    ref = tables["e8p_full_grid_f16"]
    g = qo.e8p_full_grid()
    assert np.array_equal(g.astype(np.float16).view(np.uint16), ref.view(np.uint16))
    # independent bit-level formulation (origin_order.cu:211-231) agrees for ALL 65536 codes
    d = qo.e8p_decode(np.arange(65536, dtype=np.uint16))
    assert np.array_equal(d.astype(np.float16).view(np.uint16), ref.view(np.uint16))
    assert hashlib.sha256(g.astype(np.float16).tobytes()).hexdigest() == \
        "07979702c06864796a4a049c1155c96d578d131de7e9117dff9cd2b12dc1458e"
    assert g[0x1234].tolist() == [0.25, -0.75, 0.25, -0.75, 1.25, 0.25, -0.75, -1.75]
    assert g[0xffff].tolist() == [-1.25, -1.25, -0.25, -0.25, -1.25, -1.25, -1.25, 0.75]


def test_e8p_grid_properties():
    g = qo.e8p_full_grid()
    assert len(np.unique(g, axis=0)) == 65536                 # all codewords distinct
    assert abs(float(np.sqrt((g.astype(np.float64) ** 2).mean())) - 1.09375 ** 0.5 * 1.09375 ** 0.5) < 0.05
    q = g * 4
    assert np.array_equal(q, np.round(q)) and np.all(np.abs(q) % 2 == 1) and np.abs(q).max() <= 15


def test_signed_int16_codes_reinterpreted():
    codes = np.array([[-1, -32768, 0x1234, 0]], dtype=np.int16)
    w = qo.decompress_e8p(codes)
    g = qo.e8p_full_grid()
    exp = np.concatenate([g[0xffff], g[0x8000], g[0x1234], g[0]]).astype(np.float16)
    assert np.array_equal(w[0], exp)


def test_d4_and_e81b_tables(tables):
    assert np.array_equal(qo.d4_grid(), tables["d4_grid"])
    assert hashlib.sha256(qo.d4_grid().astype(np.float16).tobytes()).hexdigest() == \
        "3055b7ccb5181fb734c0f0a5bf566f79481d6efbaf0bcc24c9dbe259c95b7968"
    assert np.array_equal(qo.e81b_grid(), tables["e81b_grid"])
    assert np.array_equal(qo.e81b_packed(), tables["e81b_packed"])


def test_hadK_shapes(had):
    for n, use_rand, K, padn, has in had["shapes"]:
        k, p, kind = qo.hadK_shape(int(n), bool(use_rand))
        assert (k, p) == (int(K), int(padn)), (n, use_rand)
        assert (kind is not None) == bool(has)
    assert set(int(k) for k in had["table_keys"]) == set(qo.HAD_TABLE_SIZES)


def test_matmul_hadU_matches_reference(had):
    n_checked = 0
    for n, use_rand, K, padn, has in had["shapes"]:
        for tr in (0, 1):
            key = f"x_n{n}_r{use_rand}_t{tr}"
            if key not in had.files:
                continue
            hk = had[f"hadK_n{n}_r{use_rand}"] if has else None
            y = qo.matmul_hadU(had[key], hk, int(K), int(padn), transpose=bool(tr))
            ref = had[f"y_n{n}_r{use_rand}_t{tr}"]
            assert y.shape == ref.shape
            np.testing.assert_allclose(y, ref, rtol=0, atol=2e-5 * max(1.0, np.abs(ref).max()))
            n_checked += 1
    assert n_checked >= 10


def test_fwht_is_sylvester():
    from scipy.linalg import hadamard
    rng = np.random.default_rng(1)
    for n in (1, 2, 8, 64, 256):
        x = rng.standard_normal((3, n))
        np.testing.assert_allclose(qo.fwht(x, 0.5), x @ hadamard(n).T * 0.5, atol=1e-9)


def test_tabulated_hadamards_are_hadamard(had):
    for k in (12, 20, 28, 172):
        h = had[f"table_{k}"].astype(np.int64)
        assert set(np.unique(h)) == {-1, 1}
        assert np.array_equal(h @ h.T, k * np.eye(k, dtype=np.int64))


def _case(ql, name):
    pre = name + "/"
    g = {k[len(pre):]: ql[k] for k in ql.files if k.startswith(pre)}
    return g


def _w_hat(c):
    cb = str(c["codebook"])
    if cb == "E8P12":
        return qo.decompress_e8p(c["Qidxs"])
    if cb == "E8P12RVQ4B":
        return qo.decompress_e8prvq4(c["Qidxs"])
    if cb == "D4":
        return qo.decompress_d4(c["Qidxs"])
    raise AssertionError(cb)


def test_decompress_bit_exact_vs_reference(ql):
    for name in ql["names"]:
        c = _case(ql, str(name))
        w = _w_hat(c)
        assert w.dtype == np.float16
        assert np.array_equal(w.view(np.uint16), c["W_hat"].view(np.uint16)), name


def test_quantlinear_forward_vs_reference(ql):
    """Oracle forward (reference rounding points) vs the reference module's eval forward.
    Tolerance: the reference chains up to 8 fp16 roundings; the oracle rounds at the same points but
    accumulates exactly, so they agree to a few fp16 ulp of the largest output."""
    for name in ql["names"]:
        c = _case(ql, str(name))
        fin, fout, bias, use_rand, pc, M, K_left, K_right, q_in, q_out = [int(v) for v in c["meta"]]
        y = qo.quantlinear_forward(
            c["x"], W_hat=_w_hat(c), in_features=fin, out_features=fout, q_in=q_in, q_out=q_out,
            SU=c.get("SU"), SV=c.get("SV"), bias=c.get("bias"),
            wscale_float=float(c["wscale_float"]),
            Wscale_per_channel=c["Wscale"] if pc else None,
            had_left=c.get("had_left"), K_left=K_left, had_right=c.get("had_right"), K_right=K_right,
            rounding="reference")
        ref = c["y"].astype(np.float64)
        tol = 2.0 ** -8 * np.abs(ref).max()
        assert np.abs(y - ref).max() <= tol, (name, np.abs(y - ref).max(), tol)
        # and the fp64 "truth" is within the same band
        y64 = qo.quantlinear_forward(
            c["x"], W_hat=_w_hat(c), in_features=fin, out_features=fout, q_in=q_in, q_out=q_out,
            SU=c.get("SU"), SV=c.get("SV"), bias=c.get("bias"),
            wscale_float=float(c["wscale_float"]),
            Wscale_per_channel=c["Wscale"] if pc else None,
            had_left=c.get("had_left"), K_left=K_left, had_right=c.get("had_right"), K_right=K_right,
            rounding="none")
        assert np.abs(y64 - ref).max() <= tol, (name, "fp64")


def test_state_dict_keys(ql):
    c = _case(ql, "e8p_96x80_rand")
    assert set(c["state_keys"]) == {"SU", "SV", "Qidxs", "Wscale", "weight", "bias", "had_left", "had_right"}
    c = _case(ql, "e8p_96x160_tab")   # use_rand=False: hadK buffers are non-persistent
    assert set(c["state_keys"]) == {"SU", "SV", "Qidxs", "Wscale", "weight"}


def test_qidxs_shapes():
    assert qo.qidxs_shape(4096, 4096, "E8P12") == (4096, 512)
    assert qo.qidxs_shape(4096, 11008, "E8P12") == (11008, 512)
    assert qo.qidxs_shape(11008, 4096, "E8P12") == (4096, 1376)
    assert qo.qidxs_shape(8192, 28672, "E8P12RVQ4B") == (28672, 1024)
    assert qo.qidxs_shape(4096, 4096, "D4") == (4096, 1024)
    assert qo.qidxs_shape(4096, 4096, "E8P12RVQ3B") == (4096, 384)
    assert qo.qidxs_shape(66, 50, "E8P12", use_rand=False) == (64, 16)   # exp<2: padded to 128 / 64
    assert qo.qidxs_shape(88, 40, "E8P12", use_rand=False) == (40, 11)   # tables 44 and 20 exist


def test_decoder_layer_step_oracle_matches_hf_llama_layer():
    """Pin the decode-loop-body oracle (llama_decoder_layer_step: RMSNorm, rotate_half RoPE, GQA attention over the
    KV cache, residuals, SiLU MLP) on HuggingFace's own LlamaModel: one layer, dense nn.Linear weights equal to the
    QuantLinears' effective matrices, a 6-token causal forward vs 6 oracle steps (no fp16 rounding, float64)."""
    import math
    import torch
    from transformers import LlamaConfig, LlamaModel
    rng = np.random.default_rng(3)
    H, I, nh, nkv, hd, T = 128, 256, 2, 1, 64, 6

    def lin(fi, fo):
        q = rng.integers(-32768, 32768, (fo, fi // 8)).astype(np.int16)
        return dict(W_hat=qo.decompress_e8p(q), in_features=fi, out_features=fo, q_in=fi, q_out=fo,
                    SU=np.sign(rng.standard_normal(fi)), SV=np.sign(rng.standard_normal(fo)), wscale_float=0.05)

    layer = dict(input_norm=1 + 0.1 * rng.standard_normal(H), post_norm=1 + 0.1 * rng.standard_normal(H),
                 q=lin(H, nh * hd), k=lin(H, nkv * hd), v=lin(H, nkv * hd), o=lin(nh * hd, H),
                 gate=lin(H, I), up=lin(H, I), down=lin(I, H))
    # two layers: hidden_states[1] is then layer 0's raw output (HF applies the final norm to the LAST entry only)
    cfg = LlamaConfig(hidden_size=H, intermediate_size=I, num_hidden_layers=2, num_attention_heads=nh,
                      num_key_value_heads=nkv, vocab_size=32, rms_norm_eps=1e-5, max_position_embeddings=64,
                      attn_implementation="eager")
    model = LlamaModel(cfg).double().eval()
    hf = model.layers[0]
    with torch.no_grad():
        for name, mod in (("q", hf.self_attn.q_proj), ("k", hf.self_attn.k_proj), ("v", hf.self_attn.v_proj),
                          ("o", hf.self_attn.o_proj), ("gate", hf.mlp.gate_proj), ("up", hf.mlp.up_proj),
                          ("down", hf.mlp.down_proj)):
            p = layer[name]
            W = qo.quantlinear_forward(np.eye(p["in_features"]), rounding="none", **p).T      # [out, in]
            mod.weight.copy_(torch.from_numpy(W))
        hf.input_layernorm.weight.copy_(torch.from_numpy(layer["input_norm"]))
        hf.post_attention_layernorm.weight.copy_(torch.from_numpy(layer["post_norm"]))
        x = rng.standard_normal((1, T, H))
        ref = model(inputs_embeds=torch.from_numpy(x), output_hidden_states=True).hidden_states[1][0].numpy()
    kc, vc = np.zeros((nkv, 16, hd)), np.zeros((nkv, 16, hd))
    for t in range(T):
        got = qo.llama_decoder_layer_step(x[0, t], layer, kc, vc, t, n_heads=nh, n_kv_heads=nkv, head_dim=hd, eps=1e-5,
                                          rounding="none")
        # HF's LlamaRMSNorm computes its statistics in float32 whatever the module dtype: 1e-6 relative
        assert np.abs(got - ref[t]).max() <= 2e-6 * max(1.0, np.abs(ref[t]).max()), (t, np.abs(got - ref[t]).max())
