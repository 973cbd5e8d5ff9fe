"""Name patterns used to locate the transformer blocks of a HF model (reference: constants.py:19-26)."""

BLOCK_PATTERNS = [
    "transformer.h",
    "model.decoder.layers",
    "gpt_neox.layers",
    "model.layers",
]

QUIP_CONFIG = "quantization_config.json"
