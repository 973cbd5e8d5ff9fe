"""Generate tests/golden/ref_ckpt_*/ : tiny quantized-Llama checkpoint folders whose module tree, state-dict keys,
tensor shapes / dtypes and quantization config are produced by the REFERENCE'S OWN CODE, imported in the build container.

Run once (already done; outputs are committed):   python tests/golden/gen_checkpoint.py
Needs /root/reference (read-only); nothing at test time reads it.

What runs from the reference:
  * quantizer.py  QuipQuantizer.__init__ / convert_model / _replace_by_quant_layers / to_dict   (quantizer.py:60-260)
  * qlinear.py    QuantLinear.__init__  (buffers + parameters = the state-dict key layout, qlinear.py:10-84)
  * codebook/*    the codebook modules attached to every layer
  * constants.py  QUIP_CONFIG (file name of the quantization config)
Absent third-party packages are stubbed FOR IMPORT ONLY (no function of a stub is called on the path above):
accelerate, fast_hadamard_transform_cuda, quiptools_cuda.  `QuipQuantizer.save` (quantizer.py:718-756) is
`Accelerator().save_model(model, dir)` + `config.save_pretrained` + a JSON dump of `to_dict()`; accelerate is not
installed here, so the state dict the reference model exposes is written with the file names accelerate uses for an
unsharded checkpoint (`pytorch_model.bin`, `model.safetensors`) and, for the sharded variant, its index-file layout
(`pytorch_model.bin.index.json` with "weight_map").  Packed codes, sign vectors and scales are random (seeded): the
fixture pins the KEY / SHAPE / DTYPE / CONFIG contract of a reference-written folder, which is what the loader consumes.
"""
import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _import_reference_quantizer():
    os.chdir(REF)
    sys.path.insert(0, REF)
    import transformers  # noqa: F401  (before the stubs: transformers probes `accelerate` through importlib at import)
    from transformers import AutoConfig, AutoModelForCausalLM, AutoTokenizer, LlamaForCausalLM  # noqa: F401
    import transformers.pytorch_utils  # noqa: F401
    nope = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stub: not available in the build container"))
    _stub("accelerate", Accelerator=nope, cpu_offload_with_hook=nope, init_empty_weights=nope,
          load_checkpoint_and_dispatch=nope)
    _stub("accelerate.hooks", remove_hook_from_module=nope)
    _stub("fast_hadamard_transform_cuda")
    _stub("quiptools_cuda")
    sys.path.insert(0, OUT)
    from gen_golden import _NpProxy
    import codebook.e8p12 as e8p12
    e8p12.np = _NpProxy()          # numpy >= 2 shim for np.int8(250), see gen_golden.py
    import quantizer
    import constants
    return quantizer, constants


def main():
    quantizer, constants = _import_reference_quantizer()
    from transformers import LlamaConfig, LlamaForCausalLM
    for cb_name in ("E8P12", "E8P12RVQ4B", "D4", "HI", "E8P12RVQ3B"):
        torch.manual_seed(0)
        cfg = LlamaConfig(hidden_size=256, intermediate_size=768, num_hidden_layers=2, num_attention_heads=2,
                          num_key_value_heads=2, vocab_size=128, max_position_embeddings=64, tie_word_embeddings=False)
        model = LlamaForCausalLM(cfg).half()
        qz = quantizer.QuipQuantizer(codebook=cb_name, dataset="c4", inference=True, ft_epochs=0)
        model = qz.convert_model(model)                      # the reference swaps nn.Linear -> its QuantLinear
        g = torch.Generator().manual_seed(1)
        import qlinear
        n_q = 0
        for name, mod in model.named_modules():
            if isinstance(mod, qlinear.QuantLinear):
                n_q += 1
                info = torch.iinfo(mod.Qidxs.dtype)
                mod.Qidxs.copy_(torch.randint(info.min, info.max + 1, mod.Qidxs.shape, dtype=torch.int64, generator=g)
                                .to(mod.Qidxs.dtype))
                sgn = lambda n: (torch.randint(0, 2, (n,), generator=g) * 2 - 1).float()
                mod.SU.data.copy_((sgn(mod.in_features) * (1 + 0.1 * torch.randn(mod.in_features, generator=g))).half())
                mod.SV.data.copy_((sgn(mod.out_features) * (1 + 0.1 * torch.randn(mod.out_features, generator=g))).half())
                mod.Wscale.fill_(0.02 / 1.09375)
        assert n_q == 14, n_q
        sd = {k: v.detach().clone().contiguous() for k, v in model.state_dict().items()}
        qcfg = qz.to_dict()

        # (a) unsharded torch checkpoint, the reference's default (safe_serialization=False)
        d = os.path.join(OUT, f"ref_ckpt_{cb_name.lower()}_bin")
        os.makedirs(d, exist_ok=True)
        torch.save(sd, os.path.join(d, "pytorch_model.bin"))
        model.config.save_pretrained(d)
        with open(os.path.join(d, constants.QUIP_CONFIG), "w", encoding="utf-8") as f:
            json.dump(qcfg, f, indent=2)
        if cb_name != "E8P12":
            continue
        # (b) sharded safetensors (accelerate's index layout), quantization config inside config.json as the
        #     reference leaves it after quantize_model (quantizer.py:707-708)
        from safetensors.torch import save_file
        d = os.path.join(OUT, "ref_ckpt_e8p12_sharded_st")
        os.makedirs(d, exist_ok=True)
        keys = sorted(sd)
        half = len(keys) // 2
        shards = {"model-00001-of-00002.safetensors": keys[:half], "model-00002-of-00002.safetensors": keys[half:]}
        weight_map = {}
        for fn, ks in shards.items():
            save_file({k: sd[k] for k in ks}, os.path.join(d, fn), metadata={"format": "pt"})
            weight_map.update({k: fn for k in ks})
        with open(os.path.join(d, "model.safetensors.index.json"), "w") as f:
            json.dump({"metadata": {"total_size": int(sum(v.numel() * v.element_size() for v in sd.values()))},
                       "weight_map": weight_map}, f, indent=2)
        model.config.quantization_config = qcfg
        model.config.save_pretrained(d)
        with open(os.path.join(d, constants.QUIP_CONFIG), "w", encoding="utf-8") as f:
            json.dump(qcfg, f, indent=2)
        # the key/shape/dtype manifest as a small text file for the CPU tests
        with open(os.path.join(OUT, "ref_ckpt_manifest.json"), "w") as f:
            json.dump({k: [list(v.shape), str(v.dtype)] for k, v in sd.items()}, f, indent=1)
    print("written")


if __name__ == "__main__":
    main()
