"""ctypes binding of the C-ABI library libquipb200.so (include/quip_b200.h).

There is NO fallback: if the library is missing or does not load, importing the ops fails loudly.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libquipb200.so")

CB_E8P12, CB_E8P12RVQ4B, CB_D4, CB_E8P12RVQ3B, CB_HI = 0, 1, 2, 3, 4
CODEBOOK_ENUM = {"E8P12": CB_E8P12, "E8P12RVQ4B": CB_E8P12RVQ4B, "D4": CB_D4,
                 "E8P12RVQ3B": CB_E8P12RVQ3B, "HI": CB_HI}
EUNSUPPORTED = -4
MM_MAX_M = 16

EXPORTS = [
    "quipb200_abi_version", "quipb200_strerror", "quipb200_sm_count",
    "quipb200_hadamard",
    "quipb200_decompress_e8p", "quipb200_decompress_e8prvq4", "quipb200_decompress_d4",
    "quipb200_decompress_e8prvq3", "quipb200_decompress_hi",
    "quipb200_mm_workspace_bytes", "quipb200_mm",
    "quipb200_linear_workspace_bytes", "quipb200_linear_forward",
    "quipb200_linear_group_workspace_bytes", "quipb200_linear_group_forward", "quipb200_attn_decode",
    "quipb200_decode_step_workspace_bytes", "quipb200_decode_step", "quipb200_decode_step_debug", "quipb200_decode_step_debug_cta", "quipb200_decode_step_set_splits",
    "quipb200_e8p_mm_umma_workspace_bytes", "quipb200_e8p_mm_umma", "quipb200_mm_umma", "quipb200_rotate_batched",
    "quipb200_lm_tail_workspace_bytes", "quipb200_lm_tail",
    "quipb200_e8p_quantize_workspace_bytes", "quipb200_e8p_quantize", "quipb200_e8prvq3_quantize",
    "quipb200_mailbox_create", "quipb200_mailbox_open", "quipb200_mailbox_close", "quipb200_mailbox_destroy",
    "quipb200_handoff_send", "quipb200_handoff_wait",
    "quipb200_set_option", "quipb200_get_option", "quipb200_launch_count", "quipb200_debug_timeline",
]


class LinearDesc(Structure):
    """struct quipb200_linear (include/quip_b200.h)."""
    _fields_ = [
        ("codebook", c_int32), ("in_features", c_int32), ("out_features", c_int32),
        ("q_in", c_int32), ("q_out", c_int32), ("K_left", c_int32), ("K_right", c_int32),
        ("wscale_float", c_float), ("resid_scale", c_float),
        ("qidxs", c_void_p), ("grid", c_void_p), ("SU", c_void_p), ("SV", c_void_p), ("bias", c_void_p),
        ("had_left", c_void_p), ("had_right", c_void_p), ("wscale_pc", c_void_p),
    ]


class Fusion(Structure):
    """struct quipb200_fusion (include/quip_b200.h)."""
    _fields_ = [("pre_norm_weight", c_void_p), ("pre_norm_eps", c_float), ("reserved", c_int32),
                ("gate", c_void_p), ("ldgate", c_int64), ("residual", c_void_p), ("ldres", c_int64)]


class QuipB200Error(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QuipB200Error(
            f"{LIB_PATH} not found: the CUDA library is not built. Run `python -m quip_for_all_b200.build` "
            "(or __graft_entry__.build()). There is no CPU / eager fallback for the quip_lib ops.")
    L = ctypes.CDLL(LIB_PATH)
    vp = c_void_p
    L.quipb200_abi_version.restype = c_int
    L.quipb200_strerror.restype = c_char_p
    L.quipb200_strerror.argtypes = [c_int]
    L.quipb200_sm_count.restype = c_int
    L.quipb200_hadamard.argtypes = [vp, vp, c_int64, c_int, c_float, c_int, vp]
    L.quipb200_decompress_e8p.argtypes = [vp, vp, vp, c_int64, c_int64, vp]
    L.quipb200_decompress_e8prvq4.argtypes = [vp, vp, vp, c_int64, c_int64, c_float, vp]
    L.quipb200_decompress_d4.argtypes = [vp, vp, vp, c_int64, c_int64, vp]
    L.quipb200_decompress_e8prvq3.argtypes = [vp, vp, vp, vp, c_int64, c_int64, c_float, vp]
    L.quipb200_decompress_hi.argtypes = [vp, vp, c_int64, c_int64, vp]
    L.quipb200_mm_workspace_bytes.restype = c_size_t
    L.quipb200_mm_workspace_bytes.argtypes = [c_int, c_int, c_int]
    L.quipb200_mm.argtypes = [c_int, vp, vp, vp, c_float, vp, c_int, c_int, c_int, vp, c_size_t, vp]
    L.quipb200_linear_workspace_bytes.restype = c_size_t
    L.quipb200_linear_workspace_bytes.argtypes = [POINTER(LinearDesc), c_int]
    L.quipb200_linear_forward.argtypes = [POINTER(LinearDesc), vp, c_int64, vp, c_int64, c_int, vp, c_size_t, vp]
    L.quipb200_linear_group_workspace_bytes.restype = c_size_t
    L.quipb200_linear_group_workspace_bytes.argtypes = [POINTER(LinearDesc), c_int, c_int]
    L.quipb200_linear_group_forward.argtypes = [POINTER(LinearDesc), c_int, POINTER(Fusion), vp, c_int64,
                                                POINTER(c_void_p), POINTER(c_int64), c_int, vp, c_size_t, vp]
    L.quipb200_attn_decode.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, c_int, c_int, c_int, c_int, vp]
    L.quipb200_e8p_mm_umma_workspace_bytes.restype = c_size_t
    L.quipb200_e8p_mm_umma_workspace_bytes.argtypes = [c_int, c_int, c_int]
    L.quipb200_e8p_mm_umma.restype = c_int
    L.quipb200_e8p_mm_umma.argtypes = [vp, vp, vp, vp, c_int, c_int, c_int, vp, c_size_t, vp]
    L.quipb200_mm_umma.restype = c_int
    L.quipb200_mm_umma.argtypes = [c_int, vp, vp, vp, vp, c_float, vp, c_int, c_int, c_int, vp, c_size_t, vp]
    L.quipb200_rotate_batched.restype = c_int
    L.quipb200_rotate_batched.argtypes = [vp, c_int64, vp, c_int64, vp, vp, vp, vp, c_int, c_int, c_int, c_int, c_int,
                                          c_float, vp]
    L.quipb200_lm_tail_workspace_bytes.restype = c_size_t
    L.quipb200_lm_tail.restype = c_int
    L.quipb200_lm_tail.argtypes = [vp, vp, c_float, vp, vp, c_int, c_int, vp, vp, vp, vp, vp, c_size_t, vp]
    L.quipb200_e8p_quantize_workspace_bytes.restype = c_size_t
    L.quipb200_e8p_quantize_workspace_bytes.argtypes = [c_int64]
    L.quipb200_e8p_quantize.restype = c_int
    L.quipb200_e8p_quantize.argtypes = [vp, c_int64, vp, c_int, c_float, vp, vp, vp, c_size_t, vp]
    L.quipb200_e8prvq3_quantize.restype = c_int
    L.quipb200_e8prvq3_quantize.argtypes = [vp, c_int64, vp, vp, c_float, vp, vp, vp, c_size_t, vp]
    L.quipb200_mailbox_create.argtypes = [c_size_t, POINTER(c_void_p), vp]
    L.quipb200_mailbox_open.argtypes = [vp, POINTER(c_void_p)]
    L.quipb200_mailbox_close.argtypes = [vp]
    L.quipb200_mailbox_destroy.argtypes = [vp]
    L.quipb200_handoff_send.argtypes = [vp, vp, c_size_t, vp, vp, vp]
    L.quipb200_handoff_wait.argtypes = [vp, vp, vp, vp, c_size_t, vp, vp]
    for fn in ("quipb200_mailbox_create", "quipb200_mailbox_open", "quipb200_mailbox_close", "quipb200_mailbox_destroy",
               "quipb200_handoff_send", "quipb200_handoff_wait"):
        getattr(L, fn).restype = c_int
    L.quipb200_debug_timeline.argtypes = [vp]
    L.quipb200_debug_timeline.restype = c_int
    L.quipb200_set_option.argtypes = [c_char_p, c_int]
    L.quipb200_get_option.argtypes = [c_char_p]
    L.quipb200_launch_count.restype = c_int64
    for fn in ("quipb200_hadamard", "quipb200_decompress_e8p", "quipb200_decompress_e8prvq4",
               "quipb200_decompress_d4", "quipb200_decompress_e8prvq3", "quipb200_decompress_hi",
               "quipb200_mm", "quipb200_linear_forward", "quipb200_linear_group_forward", "quipb200_attn_decode",
               "quipb200_decode_step_workspace_bytes", "quipb200_decode_step", "quipb200_decode_step_debug", "quipb200_decode_step_debug_cta", "quipb200_decode_step_set_splits",
    "quipb200_set_option", "quipb200_get_option"):
        getattr(L, fn).restype = c_int
    if L.quipb200_abi_version() != 1:
        raise QuipB200Error("libquipb200.so ABI version mismatch")
    _lib = L
    # tuning switches for A/B runs (tools/ab_bench.sh): QUIPB200_OPTIONS="name=value,name=value"
    for kv in filter(None, os.environ.get("QUIPB200_OPTIONS", "").split(",")):
        k, _, v = kv.partition("=")
        L.quipb200_set_option(k.strip().encode(), int(v))
    return L


def check(rc, what=""):
    """Raise on a non-zero return code (negative: argument error, positive: cudaError_t)."""
    if rc != 0:
        msg = lib().quipb200_strerror(rc).decode()
        raise QuipB200Error(f"{what}: {msg} (code {rc})")


def set_option(name, value):
    check(lib().quipb200_set_option(name.encode(), int(value)), f"set_option({name})")


def get_option(name):
    return int(lib().quipb200_get_option(name.encode()))


def launch_count():
    return int(lib().quipb200_launch_count())
