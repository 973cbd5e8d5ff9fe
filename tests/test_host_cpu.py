"""CPU-side tests of the product host code: tables, module geometry, op registry, C-ABI surface, loader."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

import quip_oracle as qo
from helpers import make_layer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_tables_equal_oracle_and_reference(golden_dir):
    from quip_for_all_b200.codebook import e8p12, d4, e8p12_rvq3
    t = np.load(os.path.join(golden_dir, "tables.npz"))
    assert np.array_equal(e8p12.get_packed_abs_grid().numpy(), t["e8p_abs"])
    assert np.array_equal(e8p12.get_packed_abs_grid().numpy(), qo.e8p_abs_table())
    g, idx = e8p12.get_full_grid()
    assert np.array_equal(g.numpy().astype(np.float16).view(np.uint16), t["e8p_full_grid_f16"].view(np.uint16))
    assert np.array_equal(d4.build_D4_CB().numpy(), t["d4_grid"])
    assert np.array_equal(e8p12_rvq3.get_e81bgrid().numpy(), t["e81b_grid"])
    assert np.array_equal(e8p12_rvq3.pack_e81b(e8p12_rvq3.get_e81bgrid()).numpy(), t["e81b_packed"])


def test_codebook_attributes():
    from quip_for_all_b200 import codebook_id
    from fractions import Fraction
    exp = {"E8P12": (8, 1, torch.int16, 1.03), "E8P12RVQ4B": (8, 1, torch.int32, 1.03),
           "E8P12RVQ3B": (8, Fraction(4, 3), torch.int32, 0.98), "D4": (4, 1, torch.uint8, 1.21),
           "HI": (1, 8, torch.int32, 2.97)}
    assert set(codebook_id) == set(exp)
    for k, (codesz, packsz, dt, opt) in exp.items():
        cb = codebook_id[k](inference=True)
        assert (cb.id, cb.codesz, cb.packsz, cb.idx_dtype, cb.opt_scale, cb.pack_out) == (k, codesz, packsz, dt, opt, False)
        assert cb.state_dict() == {}          # only non-persistent buffers
    assert abs(codebook_id["E8P12RVQ4B"](inference=True).opt_resid_scale - 1 / 3.45) < 1e-12
    assert codebook_id["E8P12RVQ4B"](inference=True, opt_resid_scale=0.3).opt_resid_scale == 0.3


def test_quantize_roundtrip_cpu():
    """quantize (nearest codeword) of a codeword returns its index: the grid is self-consistent."""
    from quip_for_all_b200 import codebook_id
    cb = codebook_id["E8P12"](inference=False)
    idx = torch.randint(0, 65536, (200,))
    vals, got = cb.quantize(cb.grid[idx])
    assert torch.equal(got, idx) and torch.equal(vals, cb.grid[idx])
    d = codebook_id["D4"](inference=False)
    i = torch.arange(256)
    assert torch.equal(d.quantize(d.grid[i])[1].long(), i)


def test_quantlinear_geometry_and_state_dict():
    lay = make_layer(4096, 11008, "E8P12", bias=True)
    sd = lay.state_dict()
    assert set(sd) == {"SU", "SV", "Qidxs", "Wscale", "weight", "bias", "had_right"}
    assert sd["Qidxs"].shape == (11008, 512) and sd["Qidxs"].dtype == torch.int16
    assert sd["had_right"].shape == (43, 43) and sd["had_right"].dtype == torch.float16
    assert sd["Wscale"].shape == () and sd["Wscale"].dtype == torch.float32
    assert sd["weight"].shape == () and lay.K_left == 1 and lay.K_right == 43
    assert lay.infeatures == 4096 and lay.outfeatures == 11008
    assert abs(lay.wscale_float - 0.02 / 1.09375) < 1e-7
    lay = make_layer(11008, 4096, "E8P12RVQ4B")
    assert lay.Qidxs.shape == (4096, 1376) and lay.Qidxs.dtype == torch.int32 and lay.K_left == 43
    lay = make_layer(4096, 4096, "D4")
    assert lay.Qidxs.shape == (4096, 1024) and lay.Qidxs.dtype == torch.uint8
    lay = make_layer(4096, 4096, "E8P12RVQ3B")
    assert lay.Qidxs.shape == (4096, 384)
    lay = make_layer(4096, 4096, "HI")
    assert lay.Qidxs.shape == (4096, 512)
    from quip_for_all_b200 import QuipLinear, QuantLinear
    assert QuipLinear is QuantLinear


def test_had_orthonormal_random_block():
    lay = make_layer(96, 80, "E8P12")
    for h in (lay.had_left, lay.had_right):
        hf = h.float()
        assert torch.allclose(hf @ hf.T, torch.eye(hf.shape[0]), atol=5e-3)


def test_matmul_hadU_cpu_matches_golden(golden_dir):
    from quip_for_all_b200.quant import matmul_hadU
    had = np.load(os.path.join(golden_dir, "hadamard.npz"))
    n_checked = 0
    for n, use_rand, K, padn, has in had["shapes"]:
        for tr in (0, 1):
            key = f"x_n{n}_r{use_rand}_t{tr}"
            if key not in had.files:
                continue
            hk = torch.tensor(had[f"hadK_n{n}_r{use_rand}"]) if has else None
            y = matmul_hadU(torch.tensor(had[key]), hk, int(K), int(padn), transpose=bool(tr))
            ref = torch.tensor(had[f"y_n{n}_r{use_rand}_t{tr}"])
            assert torch.allclose(y, ref, atol=2e-5 * max(1.0, ref.abs().max().item()))
            n_checked += 1
    assert n_checked >= 10


def test_get_hadK_shapes_with_tables(golden_dir):
    from quip_for_all_b200 import quant
    had = np.load(os.path.join(golden_dir, "hadamard.npz"))
    for k in (12, 20, 28, 172):
        quant.register_had_table(k, torch.tensor(had[f"table_{k}"].astype(np.float32)))
    hk, K, n = quant.get_hadK(11008, use_rand=False)
    assert (K, n) == (172, 11008) and torch.allclose(hk @ hk.T, torch.eye(172), atol=1e-5)
    hk, K, n = quant.get_hadK(28672, use_rand=False)
    assert (K, n) == (28, 28672)
    assert quant.get_hadK(66, use_rand=False)[1:] == (1, 128)       # exp < 2 -> zero-pad
    assert quant.get_hadK(4096, use_rand=False) == (None, 1, 4096)
    hk, K, n = quant.get_hadK(11008, use_rand=True)
    assert (K, n) == (43, 11008) and hk.shape == (43, 43)


REF_SCHEMAS = {
    "hadamard": "quip_lib::hadamard(Tensor x, float scale) -> Tensor",
    "e8p_mm_origorder": "quip_lib::e8p_mm_origorder(Tensor x, Tensor Qidxs, Tensor grid) -> Tensor",
    "e8prvq3_mm_origorder": "quip_lib::e8prvq3_mm_origorder(Tensor x, Tensor Qidxs, Tensor grid, Tensor grid2, float scale) -> Tensor",
    "e8prvq4_mm_origorder": "quip_lib::e8prvq4_mm_origorder(Tensor x, Tensor Qidxs, Tensor grid, float scale) -> Tensor",
    "d4_mm_origorder": "quip_lib::d4_mm_origorder(Tensor x, Tensor Qidxs, Tensor grid) -> Tensor",
    "hi_mm_origorder": "quip_lib::hi_mm_origorder(Tensor x, Tensor Qidxs) -> Tensor",
    "decompress_e8p_origorder": "quip_lib::decompress_e8p_origorder(Tensor Qidxs, Tensor grid) -> Tensor",
    "decompress_e8prvq3_origorder": "quip_lib::decompress_e8prvq3_origorder(Tensor Qidxs, Tensor grid, Tensor grid2, float scale) -> Tensor",
    "decompress_e8prvq4_origorder": "quip_lib::decompress_e8prvq4_origorder(Tensor Qidxs, Tensor grid, float scale) -> Tensor",
    "decompress_d4_origorder": "quip_lib::decompress_d4_origorder(Tensor Qidxs, Tensor grid) -> Tensor",
    "decompress_hi_origorder": "quip_lib::decompress_hi_origorder(Tensor Qidxs) -> Tensor",
}


def test_op_registry_schemas_match_reference():
    import quip_for_all_b200  # noqa: F401
    for name, schema in REF_SCHEMAS.items():
        op = getattr(torch.ops.quip_lib, name).default
        assert str(op._schema) == schema
    assert hasattr(torch.ops.quip_lib, "quantlinear_fwd")


def test_ops_are_cuda_only_like_reference():
    import quip_for_all_b200  # noqa: F401
    with pytest.raises(NotImplementedError):
        torch.ops.quip_lib.hadamard(torch.zeros(2, 8), 1.0)
    with pytest.raises(NotImplementedError):
        torch.ops.quip_lib.decompress_e8p_origorder(torch.zeros(2, 8, dtype=torch.int16), torch.zeros(256, dtype=torch.int64))
    lay = make_layer(64, 64, "E8P12")
    with pytest.raises(NotImplementedError):
        lay(torch.zeros(1, 64, dtype=torch.float16))


def test_fake_impls_give_shapes():
    import quip_for_all_b200  # noqa: F401
    q = torch.empty(32, 16, dtype=torch.int16, device="meta")
    g = torch.empty(256, dtype=torch.int64, device="meta")
    x = torch.empty(3, 128, dtype=torch.float16, device="meta")
    assert torch.ops.quip_lib.decompress_e8p_origorder(q, g).shape == (32, 128)
    assert torch.ops.quip_lib.e8p_mm_origorder(x, q, g).shape == (3, 32)
    assert torch.ops.quip_lib.hadamard(x, 0.5).shape == (3, 128)
    q3 = torch.empty(32, 12, dtype=torch.int32, device="meta")
    assert torch.ops.quip_lib.decompress_e8prvq3_origorder(q3, g, g, 0.5).shape == (32, 128)
    assert torch.ops.quip_lib.decompress_d4_origorder(torch.empty(8, 4, dtype=torch.uint8, device="meta"), g).shape == (8, 16)


def test_cabi_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports everything include/quip_b200.h declares."""
    from quip_for_all_b200 import _native
    from quip_for_all_b200.build import build
    build()
    hdr = open(os.path.join(ROOT, "include", "quip_b200.h")).read()
    declared = set(re.findall(r"^(?:int|size_t|int64_t|const char\*)\s+(quipb200_\w+)\s*\(", hdr, re.M))
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    L = ctypes.CDLL(_native.LIB_PATH)
    for s in declared:
        assert hasattr(L, s), s
    L.quipb200_abi_version.restype = ctypes.c_int
    assert L.quipb200_abi_version() == 1
    L.quipb200_strerror.restype = ctypes.c_char_p
    assert b"workspace" in L.quipb200_strerror(-3)
    assert ctypes.sizeof(_native.LinearDesc) == 104     # 9 x 4 B (+4 pad) + 8 pointers


def test_missing_library_fails_loudly(monkeypatch):
    from quip_for_all_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", "/nonexistent/libquipb200.so")
    with pytest.raises(_native.QuipB200Error):
        _native.lib()


def _tiny_llama():
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(hidden_size=64, intermediate_size=192, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=128, max_position_embeddings=64)
    return cfg, LlamaForCausalLM


def test_convert_model_replaces_block_linears_only():
    from quip_for_all_b200 import QuantLinear, QuipQuantizer
    cfg, cls = _tiny_llama()
    model = cls(cfg).half()
    q = QuipQuantizer(codebook="E8P12", inference=True, modules_to_not_convert=["down_proj"])
    q.convert_model(model)
    assert q.block_name_to_quantize == "model.layers"
    assert q.get_no_split_module_classes(model) == ["LlamaDecoderLayer"]
    kinds = {n: type(m).__name__ for n, m in model.named_modules() if n.endswith("_proj") or n == "lm_head"}
    assert kinds["lm_head"] == "Linear"
    assert kinds["model.layers.0.self_attn.q_proj"] == "QuantLinear"
    assert kinds["model.layers.1.mlp.down_proj"] == "Linear"      # skipped by pattern
    ql = model.model.layers[0].self_attn.k_proj
    assert isinstance(ql, QuantLinear) and (ql.in_features, ql.out_features) == (64, 32) and ql.bias is None
    d = q.to_dict()
    assert d["codebook"] == "E8P12" and d["codesz"] == 8 and d["quant_method"] == "QUiP"
    q2 = QuipQuantizer.from_dict(dict(d, inference=True))      # extra keys swallowed
    assert q2.codebook.id == "E8P12"
    with pytest.raises(NotImplementedError):
        q.quantize_model(model, None)


@pytest.mark.parametrize("fmt", ["bin", "safetensors"])
def test_checkpoint_roundtrip_cpu(tmp_path, fmt):
    """Write a reference-format checkpoint folder (sharded, SU merged away on one layer), load it back."""
    from quip_for_all_b200 import QuantLinear, QuipQuantizer
    from quip_for_all_b200 import quantizer as qz
    from quip_for_all_b200.modeling import randomize_quantlinear
    cfg, cls = _tiny_llama()
    torch.manual_seed(0)
    src = cls(cfg).half()
    QuipQuantizer(codebook="E8P12", inference=True).convert_model(src)
    gen = torch.Generator().manual_seed(1)
    for m in src.modules():
        if isinstance(m, QuantLinear):
            randomize_quantlinear(m, gen)
    src.model.layers[0].self_attn.q_proj.SU = None           # merged at pack time (qlinear.py:125)
    sd = {k: v.contiguous() for k, v in src.state_dict().items()}
    keys = sorted(sd)
    shards = [dict((k, sd[k]) for k in keys[::2]), dict((k, sd[k]) for k in keys[1::2])]
    ext = "bin" if fmt == "bin" else "safetensors"
    base = "pytorch_model" if fmt == "bin" else "model"
    wm = {}
    for i, sh in enumerate(shards):
        name = f"{base}-{i+1:05d}-of-00002.{ext}"
        if fmt == "bin":
            torch.save(sh, tmp_path / name)
        else:
            from safetensors.torch import save_file
            save_file(sh, str(tmp_path / name))
        wm.update({k: name for k in sh})
    (tmp_path / f"{base}.{ext}.index.json").write_text(json.dumps({"metadata": {}, "weight_map": wm}))
    files = qz._checkpoint_files(str(tmp_path), use_safetensors=(fmt == "safetensors"))
    assert len(files) == 2
    with qz.init_empty_weights():
        dst = cls(cfg).half()
    assert dst.model.embed_tokens.weight.is_meta and not dst.model.rotary_emb.inv_freq.is_meta
    QuipQuantizer(codebook="E8P12", inference=True).convert_model(dst)
    assert qz._materialize(dst, "cpu", torch.float16) == []
    missing = qz.load_state_into(dst, files, merge_suv=False)
    assert missing == ["model.layers.0.self_attn.q_proj.SU"]       # not merged away by config: reported, not dropped
    missing = qz.load_state_into(dst, files, merge_suv=True)
    assert missing == []
    qz.apply_load_time_tricks(dst)
    assert dst.model.layers[0].self_attn.q_proj.SU is None
    assert dst.model.layers[0].self_attn.k_proj.SU is not None
    for k, v in dst.state_dict().items():
        assert torch.equal(v, sd[k]), k
    l0 = dst.model.layers[1].mlp.up_proj
    assert abs(l0.wscale_float - float(l0.Wscale)) < 1e-9
    # computed buffers keep their values (parameters only go through the meta device)
    assert torch.isfinite(dst.model.rotary_emb.inv_freq).all() and dst.model.rotary_emb.inv_freq[0] == 1.0


def test_load_quantized_model_requires_cuda(tmp_path):
    from quip_for_all_b200 import load_quantized_model
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="No GPU found"):
        load_quantized_model(str(tmp_path))


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The ctypes mirrors used by the Python host side have the layout the C ABI declares (include/quip_b200.h):
    compile a probe with gcc and compare sizeof / offsetof."""
    import ctypes
    import subprocess
    from quip_for_all_b200._native import Fusion, LinearDesc
    from quip_for_all_b200.decode_step import DecodeLayer, DecodePlan
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "probe.c"
    src.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "quip_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu\\n", sizeof(quipb200_linear_t), sizeof(quipb200_fusion_t), sizeof(quipb200_decode_layer_t),
         sizeof(quipb200_decode_plan_t));
  printf("%zu %zu %zu %zu\\n", offsetof(quipb200_linear_t, qidxs), offsetof(quipb200_linear_t, wscale_pc),
         offsetof(quipb200_decode_layer_t, input_norm_w), offsetof(quipb200_decode_layer_t, mlp_hk));
  printf("%zu %zu\\n", offsetof(quipb200_decode_plan_t, layers), offsetof(quipb200_decode_plan_t, pos));
  return 0;
}
''')
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    got = [int(v) for v in out]
    want = [ctypes.sizeof(LinearDesc), ctypes.sizeof(Fusion), ctypes.sizeof(DecodeLayer), ctypes.sizeof(DecodePlan),
            LinearDesc.qidxs.offset, LinearDesc.wscale_pc.offset, DecodeLayer.input_norm_w.offset, DecodeLayer.mlp_hk.offset,
            DecodePlan.layers.offset, DecodePlan.pos.offset]
    assert got == want, (got, want)


def test_mm_dispatch_policy_and_rotation_coverage():
    """Host-side dispatch rules of the batched path (register_lib): which (M, N, K) go to the tcgen05 kernel and which
    rotation lengths the one-pass kernels cover (everything else keeps the reference's op sequence)."""
    from quip_for_all_b200 import _native, register_lib as rl
    lib_present = os.path.exists(_native.LIB_PATH)
    assert rl.rotate_supported(4096, 1) and rl.rotate_supported(11008, 43) and rl.rotate_supported(256, 1)
    assert not rl.rotate_supported(8192, 1) and not rl.rotate_supported(28672, 7) and not rl.rotate_supported(11008, 172)
    if lib_present:
        assert _native.get_option("umma") == 2
        assert not rl.umma_preferred(1, 4096, 4096) and not rl.umma_preferred(3, 4096, 4096)
        assert rl.umma_preferred(4, 4096, 4096) and rl.umma_preferred(32, 4096, 4096)
        assert rl.umma_preferred(64, 4096, 4096) and rl.umma_preferred(64, 4096, 11008)
        # 65 .. 128 rows: layers of >= 32 Mi weights; beyond that the dense route
        assert not rl.umma_preferred(128, 4096, 4096) and rl.umma_preferred(128, 4096, 11008)
        assert not rl.umma_preferred(129, 8192, 8192) and not rl.umma_preferred(256, 8192, 28672)
        assert not rl.umma_preferred(16, 4096 + 64, 4096) and not rl.umma_preferred(300, 4096, 4096)


def test_tensor_path_fragment_placement_model():
    """The register-layout claims csrc/fwht_mma.cuh relies on, checked on a numpy model of the mma.sync fragments against
    dense Sylvester matrices: a lane's own 16-byte octet is a valid fragment of the 256-point transform ("natural
    placement"), the warp-per-row 4096-point transform, and the block / spread index maps of the CTA-wide transform.
    The same index formulas must be the ones written in the header."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("emu_fwht_frag", os.path.join(ROOT, "tools", "emu_fwht_frag.py"))
    emu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(emu)
    emu.main()
    emu.check_4096()
    hdr = open(os.path.join(ROOT, "quip_for_all_b200", "csrc", "fwht_mma.cuh")).read()
    blk = re.search(r"int idx_block\(int warp, int lane, int q\) \{ return ([^;]+); \}", hdr).group(1)
    spr = re.search(r"int idx_spread\(int warp, int lane, int q\) \{\s*return ([^;]+);", hdr).group(1)
    for warp in (0, 5, 9, 15):
        for lane in (0, 7, 18, 31):
            for q in range(4):
                env = {"warp": warp, "lane": lane, "q": q}
                assert eval(blk, {}, env) == emu.idx_block(warp, lane, q)
                assert eval(spr, {}, env) == emu.idx_spread(warp, lane, q)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (CPU port of the path on the host cores) prints one JSON line with the keys the
    driver reads; ranks other than 0 print nothing and exit 0."""
    import subprocess
    import sys
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--model", "tiny"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "tokens/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("decode tokens/s bs=1 Llama-2-7B E8P12")
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert "workload" in line["config"]
    assert line["steps"] == 1 and line["warmup"] == 0                     # used as given, not capped
    assert "all 2 decoder layers" in line["config"]["sampled"]            # a whole-token pass, not one layer x n_layers
    # under torchrun (OMP_NUM_THREADS=1 exported) the arm still uses the host's cores
    r3 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                         "--model", "tiny"], capture_output=True, text=True, env=dict(env, OMP_NUM_THREADS="1"), timeout=300)
    assert json.loads(r3.stdout.strip().splitlines()[-1])["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                         "--model", "tiny"], capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2"), timeout=300)
    assert r2.returncode == 0 and r2.stdout.strip() == ""


# ------------------------------------------------------------------------------------------------
# checkpoint folders written with the REFERENCE's module tree / key layout / config (tests/golden/gen_checkpoint.py)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("folder,st", [("ref_ckpt_e8p12_bin", False), ("ref_ckpt_e8p12_sharded_st", True),
                                       ("ref_ckpt_e8p12rvq4b_bin", False), ("ref_ckpt_d4_bin", False),
                                       ("ref_ckpt_hi_bin", False), ("ref_ckpt_e8p12rvq3b_bin", False)])
def test_loader_consumes_reference_written_checkpoint(golden_dir, folder, st):
    import json
    from quip_for_all_b200 import quantizer as qz
    path = os.path.join(golden_dir, folder)
    model = qz._load_quantized_model(path, use_safetensors=st)
    files = qz._checkpoint_files(path, st)
    ref = {}
    for f in files:
        ref.update(qz._read_shard(f))
    with open(os.path.join(golden_dir, "ref_ckpt_manifest.json")) as f:
        manifest = json.load(f)
    if "e8p12_" in folder:
        assert set(ref) == set(manifest)
    sd = model.state_dict()
    # every tensor the reference's model exposes has a destination of the same shape / dtype and arrives bit-exact
    for k, v in ref.items():
        assert k in sd, k
        assert tuple(sd[k].shape) == tuple(v.shape) and sd[k].dtype == v.dtype, k
        assert torch.equal(sd[k], v), k
    assert set(sd) == set(ref)
    from quip_for_all_b200 import QuantLinear
    ql = [m for m in model.modules() if isinstance(m, QuantLinear)]
    assert len(ql) == 14
    cb = {"e8p12": "E8P12", "e8p12rvq4b": "E8P12RVQ4B", "d4": "D4", "hi": "HI", "e8p12rvq3b": "E8P12RVQ3B"}[folder.split("_")[2]]
    for m in ql:
        assert m.codebook.id == cb
        assert m.SU is not None and m.SV is not None              # merge_suv = false in the reference config
        assert abs(m.wscale_float - float(m.Wscale)) < 1e-9        # quantizer.py:837
    mlp = model.model.layers[1].mlp                                # 768 = 3 * 256: use_rand blocks, persistent buffers
    assert mlp.gate_proj.K_right == 3 and tuple(mlp.gate_proj.had_right.shape) == (3, 3)
    assert mlp.down_proj.K_left == 3 and torch.equal(mlp.down_proj.had_left, ref["model.layers.1.mlp.down_proj.had_left"])
    if cb == "E8P12RVQ4B":       # the reference hands its default opt_resid_scale = -1 to the codebook as a literal scale
        assert ql[0].codebook.opt_resid_scale == -1                # (quantizer.py:69,232; e8p12_rvq4.py:23): kept as is
    assert torch.isfinite(model.model.rotary_emb.inv_freq).all()


def test_loader_reports_missing_scale_vectors(golden_dir, tmp_path):
    """an SU / SV absent from a checkpoint whose config says merge_suv = false is an error, not a silent None."""
    import shutil
    from quip_for_all_b200 import quantizer as qz
    src = os.path.join(golden_dir, "ref_ckpt_e8p12_bin")
    dst = tmp_path / "broken"
    shutil.copytree(src, dst)
    sd = torch.load(dst / "pytorch_model.bin", map_location="cpu", weights_only=True)
    del sd["model.layers.0.mlp.up_proj.SV"]
    torch.save(sd, dst / "pytorch_model.bin")
    with pytest.raises(RuntimeError, match="missing 1 tensors"):
        qz._load_quantized_model(str(dst))
