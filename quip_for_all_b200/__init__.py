"""quip_for_all_b200 -- B200-native (sm_100a) inference path for QuIP#-quantised linears, behind the
API of chu-tianxiang/QuIP-for-all: `QuantLinear` (alias `QuipLinear`), `codebook_id`,
`torch.ops.quip_lib.*`, `QuipQuantizer.convert_model`, `load_quantized_model`.

The compute lives in hand-written CUDA (csrc/*.cu -> lib/libquipb200.so, C ABI in include/quip_b200.h);
this package is the thin host side.  There is no CPU or eager fallback for the ops.
"""
from . import register_lib  # noqa: F401  registers torch.ops.quip_lib.*
from .codebook import codebook_id
from .qlinear import QuantLinear, QuipLinear
from .quantizer import QuipQuantizer, load_quantized_model

__all__ = ["QuantLinear", "QuipLinear", "QuipQuantizer", "codebook_id", "load_quantized_model"]
