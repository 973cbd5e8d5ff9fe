"""Model-structure helpers needed by the loader (reference: utils.py:35-93 and the recurse helpers)."""
import functools
from typing import List, Optional

from torch import nn

from .constants import BLOCK_PATTERNS

try:
    from transformers.pytorch_utils import Conv1D
except Exception:  # pragma: no cover
    class Conv1D(nn.Module):
        pass


def get_layers(module: nn.Module, layers=None, prefix: Optional[str] = None, skip: Optional[List] = None,
               name: str = ""):
    """{qualified name: module} for every module of one of the `layers` types whose name starts with
    `prefix` and contains none of the `skip` patterns."""
    if layers is None:
        layers = [Conv1D, nn.Conv2d, nn.Linear]
    skip = skip or []
    if isinstance(module, tuple(layers)):
        ok_prefix = prefix is None or name.startswith(prefix)
        if ok_prefix and not any(pat in name for pat in skip):
            return {name: module}
        if prefix is not None or skip:
            # matching type but filtered out: still descend (mirrors the reference's fall-through)
            pass
    found = {}
    for child_name, child in module.named_children():
        found.update(get_layers(child, layers=layers, prefix=prefix, skip=skip,
                                name=f"{name}.{child_name}" if name else child_name))
    return found


def get_block_name_with_pattern(model: nn.Module):
    names = [n for n, _ in model.named_modules()]
    for pattern in BLOCK_PATTERNS:
        if any(pattern in n for n in names):
            return pattern
    raise ValueError("Block pattern could not be match. Pass `block_name_to_quantize` argument in `quantize_model`")


def recurse_getattr(obj, attr: str):
    """getattr through a dotted path."""
    return functools.reduce(getattr, [obj] + attr.split("."))


def recurse_setattr(module, name, value):
    if "." not in name:
        setattr(module, name, value)
    else:
        head, rest = name.split(".", 1)
        recurse_setattr(getattr(module, head), rest, value)
