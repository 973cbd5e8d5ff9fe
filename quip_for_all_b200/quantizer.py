"""Model surgery, checkpoint loading and the offline quantisation driver (reference: quantizer.py:53-848).

`QuipQuantizer` keeps the reference's constructor keywords / `to_dict` / `from_dict` / `convert_model` /
`get_no_split_module_classes`.  `quantize_model` (block-wise calibration -> LDLQ per sub-layer group with the GPU
codeword search, `ldlq.py` / `csrc/nearest.cu` -> packed `QuantLinear`s) and `save` (the reference's checkpoint layout) are
here too; what is NOT part of this build: fetching / tokenising a calibration dataset (no network: ready batches are
passed in), block-wise fine-tuning (`ft_epochs > 0`) and `merge_suv` at quantisation time -- each raises.

`load_quantized_model` has the reference signature but does not need `accelerate`: the HF model is
instantiated on the meta device, every block linear is swapped for a `QuantLinear`, and the
checkpoint shards (`pytorch_model*.bin` or `*.safetensors`, with or without an index json -- what
`Accelerator.save_model` writes, quantizer.py:718-756) are streamed into it.
"""
import contextlib
import json
import os
from logging import getLogger
from typing import Any, Dict, List, Optional, Union

import torch
from torch import nn

from .codebook import codebook_id
from .constants import QUIP_CONFIG
from .qlinear import QuantLinear
from .utils import Conv1D, get_block_name_with_pattern, get_layers, recurse_getattr

logger = getLogger(__name__)


class QuipQuantizer(object):
    """Configuration holder + layer replacement.  Keyword set = reference quantizer.py:58-79."""

    def __init__(self, codebook: str, dataset: str = "redpajama", nsamples: int = 4096,
                 model_seqlen: int = 2048, quip_tune_iters: int = 10, use_rand: bool = True,
                 rescale_WH: bool = False, sigma_reg: float = 1e-2, sigma_reg2: float = 1e-2,
                 modules_to_not_convert: Optional[List] = None, block_name_to_quantize: Optional[str] = None,
                 merge_suv: bool = False, per_channel: bool = False, opt_resid_scale: Optional[float] = -1,
                 inference: bool = False, ft_epochs: int = 5, ft_lr: float = 5e-5, ft_susv_lr: float = 5e-4,
                 ft_valid_size: int = 128, ft_bs: int = 8, ft_update_freq: int = 2, ft_early_stop: int = 3,
                 cache_on_gpu: bool = False, scale_override: float = -1, *args, **kwargs):
        if codebook not in codebook_id:
            raise ValueError(f"unknown codebook {codebook!r}; expected one of {sorted(codebook_id)}")
        self.codebook = codebook_id[codebook](inference=inference, opt_resid_scale=opt_resid_scale)
        self.dataset = dataset
        self.nsamples = nsamples
        self.model_seqlen = model_seqlen
        self.quip_tune_iters = quip_tune_iters
        self.use_rand = use_rand
        self.rescale_WH = rescale_WH
        self.sigma_reg = sigma_reg
        self.sigma_reg2 = sigma_reg2
        self.modules_to_not_convert = modules_to_not_convert or []
        self.block_name_to_quantize = block_name_to_quantize
        self.merge_suv = merge_suv
        self.per_channel = per_channel
        self.opt_resid_scale = opt_resid_scale
        self.inference = inference
        self.ft_epochs, self.ft_lr, self.ft_susv_lr = ft_epochs, ft_lr, ft_susv_lr
        self.ft_valid_size, self.ft_bs = ft_valid_size, ft_bs
        self.ft_update_freq, self.ft_early_stop = ft_update_freq, ft_early_stop
        self.quant_method = "QUiP"
        self.cache_on_gpu = cache_on_gpu        # calibration activations stay on the device between blocks (quantizer.py:75)
        self.scale_override = scale_override

    def to_dict(self):
        """quantization_config.json contents (reference: quantizer.py:132-147)."""
        return {
            "quant_method": "QUiP",
            "rescale_WH": self.rescale_WH,
            "use_rand": self.use_rand,
            "codebook": self.codebook.id,
            "codesz": self.codebook.codesz,
            "idx_dtype": str(self.codebook.idx_dtype),
            "merge_suv": self.merge_suv,
            "per_channel": self.per_channel,
            "opt_resid_scale": self.opt_resid_scale,
            "modules_to_not_convert": self.modules_to_not_convert,
        }

    @classmethod
    def from_dict(cls, config_dict: Dict[str, Any]):
        return cls(**config_dict)   # extra keys (codesz, idx_dtype, quant_method) fall into **kwargs

    def convert_model(self, model: nn.Module):
        """Swap every Linear / Conv1D / Conv2d under the block prefix for a QuantLinear."""
        if self.block_name_to_quantize is None:
            self.block_name_to_quantize = get_block_name_with_pattern(model)
        targets = get_layers(model, prefix=self.block_name_to_quantize, skip=self.modules_to_not_convert)
        self._replace_by_quant_layers(model, targets)
        return model

    def get_no_split_module_classes(self, model):
        block = recurse_getattr(model, self.block_name_to_quantize)[0]
        return [block.__class__.__name__]

    def _replace_by_quant_layers(self, model: nn.Module, targets: Dict[str, nn.Module]):
        for name, layer in targets.items():
            if isinstance(layer, QuantLinear):
                continue
            if isinstance(layer, nn.Linear):
                fin, fout = layer.in_features, layer.out_features
            elif isinstance(layer, nn.Conv2d):
                fin, fout = layer.in_channels, layer.out_channels
            elif isinstance(layer, Conv1D):
                fin, fout = layer.weight.shape[0], layer.weight.shape[1]
            else:
                continue
            device = layer.weight.device
            # a fresh codebook object per layer, as the reference does (quantizer.py:230-233)
            cb = codebook_id[self.codebook.id](inference=True, opt_resid_scale=self.opt_resid_scale)
            with torch.device("cpu"):
                new = QuantLinear(fin, fout, cb, bias=(layer.bias is not None), use_rand=self.use_rand,
                                  per_channel=self.per_channel, weight_dtype=layer.weight.dtype)
            if device.type != "meta":
                new = new.to(device)
            parent_name, _, attr = name.rpartition(".")
            parent = recurse_getattr(model, parent_name) if parent_name else model
            setattr(parent, attr, new)

    # ---------------------------------------------------------------------------------------------
    # offline quantisation (reference: quantizer.py:250-600).  Calibration data: the reference tokenises a hub dataset
    # (data.py); there is no network here, so `calib` is an iterable of ready batches -- LongTensor input_ids [B, T] or
    # dicts of model kwargs.  Block-wise fine-tuning (ft_epochs > 0, quantizer.py:501-567) is not part of this build.
    # ---------------------------------------------------------------------------------------------
    @staticmethod
    def _sublayer_groups(names):
        """Order in which the linears of a block are quantised: q/k/v -> attention out -> MLP in -> MLP out, each group
        seeing the already-quantised groups before it (reference: utils.split_block_to_sublayers); unknown layouts are
        treated as one group."""
        roles = (("q_proj", "k_proj", "v_proj", "query_key_value", "c_attn", "qkv_proj", "W_pack"),
                 ("o_proj", "out_proj", "attn.c_proj", "attention.dense", "self_attention.dense"),
                 ("gate_proj", "up_proj", "dense_h_to_4h", "c_fc", "fc_in", "fc1"),
                 ("down_proj", "dense_4h_to_h", "mlp.c_proj", "fc_out", "fc2"))
        groups = [[n for n in names if any(n.endswith(r) for r in role)] for role in roles]
        if sum(len(g) for g in groups) != len(names) or len({n for g in groups for n in g}) != len(names):
            return [list(names)]
        return [g for g in groups if g]

    @torch.no_grad()
    def quantize_model(self, model: nn.Module, calib, save_dir: str = ""):
        from .ldlq import LayerQuantizer
        if isinstance(calib, str) or hasattr(calib, "encode") or hasattr(calib, "tokenize"):
            raise ValueError("quantize_model: pass calibration batches (input_ids tensors or kwargs dicts); fetching and "
                             "tokenising a hub dataset (reference data.py) needs network access")
        if self.ft_epochs and self.ft_epochs > 0:
            raise NotImplementedError("block-wise fine-tuning (ft_epochs > 0) is not part of this build: pass ft_epochs=0")
        if self.merge_suv:
            raise NotImplementedError("merge_suv=True at quantisation time (sharing sign vectors between consecutive "
                                      "layers, quantizer.py:409-421) is not part of this build; checkpoints written that "
                                      "way by the reference load fine")
        if getattr(self.codebook, "grid", None) is None:
            raise ValueError("quantize_model needs the full codebook: construct QuipQuantizer(..., inference=False)")
        model.eval()
        use_cache = getattr(getattr(model, "config", None), "use_cache", None)
        if use_cache is not None:
            model.config.use_cache = False
        if self.block_name_to_quantize is None:
            self.block_name_to_quantize = get_block_name_with_pattern(model)
        blocks = recurse_getattr(model, self.block_name_to_quantize)
        dev = next(model.parameters()).device
        origin_dtype = next(model.parameters()).dtype

        # inputs of the first block: run the model until the block is entered
        class _Stop(Exception):
            pass

        inputs, in_kwargs = [], []

        keep = (lambda t: t) if self.cache_on_gpu else (lambda t: t.cpu())   # host RAM unless cache_on_gpu, as the reference

        def grab(_, args, kwargs):
            x = args[0] if args else kwargs["hidden_states"]
            inputs.append(keep(x.detach()))
            in_kwargs.append({k: v for k, v in kwargs.items() if k != "hidden_states"})
            raise _Stop

        h = blocks[0].register_forward_pre_hook(grab, with_kwargs=True)
        for batch in calib:
            data = batch if isinstance(batch, dict) else {"input_ids": batch}
            try:
                model(**{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in data.items()})
            except _Stop:
                pass
        h.remove()
        if not inputs:
            raise ValueError("quantize_model: no calibration batch reached the first block")

        for bi, block in enumerate(blocks):
            block.float()
            layers = get_layers(block, skip=self.modules_to_not_convert)
            acc = {n: LayerQuantizer(l, codebook_id[self.codebook.id](inference=False, opt_resid_scale=self.opt_resid_scale))
                   for n, l in layers.items() if isinstance(l, nn.Linear)}
            hooks = [layers[n].register_forward_hook(lambda _, i, o, n=n: acc[n].add_batch(i[0].data)) for n in acc]
            outs = []
            for x, kw in zip(inputs, in_kwargs):
                y = block(x.to(dev).float(), **kw)
                outs.append(keep((y[0] if isinstance(y, tuple) else y).detach()))
            for hk in hooks:
                hk.remove()
            for group in self._sublayer_groups(list(acc)):
                for n in group:
                    attr = acc[n].quantize(rescale_WH=self.rescale_WH, sigma_reg=self.sigma_reg,
                                           quip_tune_iters=self.quip_tune_iters,
                                           scale_override=self.scale_override if self.scale_override and self.scale_override > 0 else 0,
                                           use_rand=self.use_rand, per_channel=self.per_channel)
                    lin = layers[n]
                    self._replace_by_quant_layers(block, {n: lin})
                    q = recurse_getattr(block, n)
                    q.cpu()
                    q.pack(lin.cpu(), attr)
                    q.to(dev)
                    q.train(block.training)        # a fresh module is in training mode (the dense calc_weight branch)
                    q.proxy_loss = acc[n].last_proxy_loss
                    q._w_scale_f32 = attr["w_scale"].to(torch.float32)
                    acc[n].H = None
            block.to(origin_dtype)
            for q in get_layers(block, [QuantLinear]).values():
                q.weight_dtype = origin_dtype
                if not q.per_channel and hasattr(q, "_w_scale_f32"):      # the scalar scale stays fp32
                    q.Wscale = q._w_scale_f32.to(dev).reshape(())
                    del q._w_scale_f32
            apply_load_time_tricks(block, self.merge_suv)
            inputs = outs            # the next block is calibrated on the un-quantised stream, as the reference does
        model.is_quantized = True
        if use_cache is not None:
            model.config.use_cache = use_cache
        if hasattr(model, "config"):
            model.config.quantization_config = self.to_dict()
        if save_dir:
            self.save(model, save_dir)
        return model

    def save(self, model: nn.Module, save_dir: str, max_shard_size: str = "10GB", safe_serialization: bool = False):
        """Write the checkpoint folder `load_quantized_model` reads: weights under the model's state-dict keys
        (`pytorch_model.bin` or `model.safetensors`, what `Accelerator.save_model` writes for a model below the shard
        size, reference quantizer.py:718-756), the HF config and quantization_config.json."""
        import json
        os.makedirs(save_dir, exist_ok=True)
        sd = {}
        for k, v in model.state_dict().items():
            t = v.detach().cpu().contiguous()
            lay = recurse_getattr(model, k.rpartition(".")[0]) if "." in k else model
            if isinstance(lay, QuantLinear) and k.endswith(".Wscale") and lay.per_channel:
                t = t * lay.wscale_float          # undo the load-time normalisation (quantizer.py:838-839)
            sd[k] = t
        if safe_serialization:
            from safetensors.torch import save_file
            save_file(sd, os.path.join(save_dir, "model.safetensors"), metadata={"format": "pt"})
        else:
            torch.save(sd, os.path.join(save_dir, "pytorch_model.bin"))
        if hasattr(model, "config"):
            qc = getattr(model.config, "quantization_config", None)
            if qc is not None:
                del model.config.quantization_config      # read back from quantization_config.json; recent transformers
                                                          # validate `quant_method` of a config.json entry on load
            model.config.save_pretrained(save_dir)
            if qc is not None:
                model.config.quantization_config = qc
        with open(os.path.join(save_dir, QUIP_CONFIG), "w", encoding="utf-8") as f:
            json.dump(self.to_dict(), f, indent=2)


# --------------------------------------------------------------------------------------------------
# checkpoint loading
# --------------------------------------------------------------------------------------------------
def load_config(model_path, safetensors=True, trust_remote_code=True, revision=None):
    from transformers import AutoConfig
    if not os.path.isdir(model_path):
        from huggingface_hub import snapshot_download
        ignore = ["*msgpack*", "*h5*", "optimizer.pt"]
        ignore += ["*.pt*", "*.bin*", "consolidated*"] if safetensors else ["*.safetensors*"]
        model_path = snapshot_download(model_path, ignore_patterns=ignore, revision=revision)
    config = AutoConfig.from_pretrained(model_path, trust_remote_code=trust_remote_code, revision=revision)
    return model_path, config


def _checkpoint_files(folder: str, use_safetensors: bool) -> List[str]:
    names = (["model.safetensors.index.json", "model.safetensors"] if use_safetensors else []) + \
            ["pytorch_model.bin.index.json", "pytorch_model.bin", "model.safetensors.index.json",
             "model.safetensors"]
    for n in names:
        p = os.path.join(folder, n)
        if not os.path.isfile(p):
            continue
        if n.endswith(".index.json"):
            with open(p) as f:
                shard_names = sorted(set(json.load(f)["weight_map"].values()))
            return [os.path.join(folder, s) for s in shard_names]
        return [p]
    raise FileNotFoundError(f"no pytorch_model*.bin / model*.safetensors checkpoint found in {folder}")


def _read_shard(path: str) -> Dict[str, torch.Tensor]:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    return torch.load(path, map_location="cpu", weights_only=True)


@contextlib.contextmanager
def init_empty_weights():
    """Parameters are created on the meta device, buffers keep real storage and their computed values (rotary
    inv_freq, causal masks, ...): the behaviour of accelerate's `init_empty_weights(include_buffers=False)` that the
    reference relies on (quantizer.py:811), without the accelerate dependency."""
    old_register = nn.Module.register_parameter

    def register_parameter(module, name, param):
        old_register(module, name, param)
        if param is not None:
            cur = module._parameters[name]
            kw = dict(cur.__dict__)
            kw["requires_grad"] = cur.requires_grad
            module._parameters[name] = cur.__class__(cur.to("meta"), **kw)

    nn.Module.register_parameter = register_parameter
    try:
        yield
    finally:
        nn.Module.register_parameter = old_register


def _materialize(model: nn.Module, device, dtype):
    """Give real storage to the parameters that are still on the meta device.  Returns the names of buffers that had no
    real storage (none when the model was built under init_empty_weights): they must come from the checkpoint."""
    meta_buffers = []
    for mname, mod in model.named_modules():
        for name, p in list(mod._parameters.items()):
            if p is not None and p.is_meta:
                mod._parameters[name] = nn.Parameter(torch.empty(p.shape, dtype=p.dtype, device=device),
                                                     requires_grad=False)
        for name, b in list(mod._buffers.items()):
            if b is not None and b.is_meta:
                mod._buffers[name] = torch.empty(b.shape, dtype=b.dtype, device=device)
                meta_buffers.append(f"{mname}.{name}" if mname else name)
    return meta_buffers


def load_state_into(model: nn.Module, files: List[str], merge_suv: bool = True):
    """Stream checkpoint shards into the model.  With `merge_suv` (quantization config), QuantLinear SU / SV vectors
    that the checkpoint lacks were merged away at pack time (qlinear.py:125-131) and are set to None; without it an
    absent SU / SV is a missing tensor like any other."""
    own = dict(model.state_dict(keep_vars=True))
    seen = set()
    for f in files:
        shard = _read_shard(f)
        for k, v in shard.items():
            if k not in own:
                logger.warning("checkpoint key %s has no destination", k)
                continue
            dst = own[k]
            if dst.shape != v.shape:
                raise RuntimeError(f"shape mismatch for {k}: checkpoint {tuple(v.shape)} vs model {tuple(dst.shape)}")
            with torch.no_grad():
                dst.copy_(v.to(dst.dtype) if dst.dtype.is_floating_point and v.dtype.is_floating_point else v)
            seen.add(k)
        del shard
    if merge_suv:
        for name, mod in model.named_modules():
            if isinstance(mod, QuantLinear):
                for attr in ("SU", "SV"):
                    if f"{name}.{attr}" not in seen:
                        setattr(mod, attr, None)
    missing = [k for k in own if k not in seen and not (merge_suv and k.endswith((".SU", ".SV")))]
    return missing


def apply_load_time_tricks(model: nn.Module, merge_suv: bool = False):
    """Post-load normalisation (reference: quantizer.py:836-844)."""
    for layer in get_layers(model, [QuantLinear]).values():
        layer.wscale_float = layer.Wscale.mean().float().item()
        if layer.per_channel:
            layer.Wscale = layer.Wscale / layer.Wscale.mean()
        if merge_suv:
            if layer.SU is not None and torch.all(layer.SU > 0):
                layer.SU = None
            if layer.SV is not None and torch.all(layer.SV > 0):
                layer.SV = None


def load_quantized_model(save_folder: str, revision: Optional[str] = None,
                         torch_dtype: Optional[Union[str, torch.dtype]] = torch.float16,
                         trust_remote_code: bool = True, use_safetensors: bool = False,
                         device_map: Optional[Union[str, dict]] = None):
    """Load a QuIP-for-all checkpoint folder into a HF model whose block linears are QuantLinears.
    Signature and behaviour follow the reference (quantizer.py:779-848), including the hard CUDA
    requirement -- there is no CPU inference path."""
    if not torch.cuda.is_available():
        raise RuntimeError("No GPU found. A GPU is needed to run quantized model.")
    return _load_quantized_model(save_folder, revision, torch_dtype, trust_remote_code, use_safetensors, device_map)


def _load_quantized_model(save_folder, revision=None, torch_dtype=torch.float16, trust_remote_code=True,
                          use_safetensors=False, device_map=None):
    """Body of load_quantized_model (host-side work only; callable without a device by the CPU tests)."""
    from transformers import AutoModelForCausalLM
    folder, config = load_config(save_folder, trust_remote_code=trust_remote_code,
                                 safetensors=use_safetensors, revision=revision)
    if isinstance(torch_dtype, str):
        torch_dtype = getattr(torch, torch_dtype)
    with init_empty_weights():
        model = AutoModelForCausalLM.from_config(config, trust_remote_code=trust_remote_code, dtype=torch_dtype)

    qcfg = getattr(config, "quantization_config", None)
    if qcfg is None:
        with open(os.path.join(folder, QUIP_CONFIG)) as f:
            qcfg = json.load(f)
    qcfg = dict(qcfg if isinstance(qcfg, dict) else qcfg.to_dict())
    qcfg["inference"] = True
    qcfg["ft_epochs"] = 0
    quantizer = QuipQuantizer.from_dict(qcfg)
    model = quantizer.convert_model(model)

    device = "cpu"
    if isinstance(device_map, str) and device_map not in ("auto", "balanced", "sequential"):
        device = device_map
    elif isinstance(device_map, dict) and set(device_map) == {""}:
        device = device_map[""]
    elif device_map is not None:
        device = "cuda"      # whole model on the current GPU; multi-GPU goes through parallel.LayerPipeline
    meta_buffers = _materialize(model, "cpu", torch_dtype)
    missing = load_state_into(model, _checkpoint_files(folder, use_safetensors), merge_suv=quantizer.merge_suv)
    persistent = set(model.state_dict().keys())
    lost = [b for b in meta_buffers if b not in persistent]
    if lost:      # a computed, non-persistent buffer without storage would be read as uninitialised memory
        raise RuntimeError(f"{len(lost)} non-persistent buffers have no values, e.g. {lost[:5]}")
    if hasattr(model, "tie_weights"):
        model.tie_weights()
    missing = [k for k in missing if not ("lm_head" in k and getattr(config, "tie_word_embeddings", False))]
    if missing:
        raise RuntimeError(f"checkpoint is missing {len(missing)} tensors, e.g. {missing[:5]}")
    apply_load_time_tricks(model, quantizer.merge_suv)
    if device != "cpu":
        model = model.to(device)
    model.is_quantized = True
    model.eval()
    return model
