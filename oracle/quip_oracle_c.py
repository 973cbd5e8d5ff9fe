"""ctypes wrapper of oracle/libquip_oracle.so (C restatement; TEST INFRASTRUCTURE, see quip_oracle.c)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libquip_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "quip_oracle.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = ctypes.CDLL(_SO)
        _lib.qo_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def num_threads():
    return lib().qo_num_threads()


def decompress_e8p(qidxs, table):
    q = np.ascontiguousarray(qidxs).view(np.uint16)
    N, K = q.shape[0], q.shape[1] * 8
    out = np.empty((N, K), dtype=np.uint16)
    t = np.ascontiguousarray(table).view(np.uint64)
    lib().qo_decompress_e8p(_p(q), _p(t), _p(out), ctypes.c_int64(N), ctypes.c_int64(K))
    return out.view(np.float16)


def e8p_mm(x16, qidxs, table):
    x = np.ascontiguousarray(x16, dtype=np.float16)
    q = np.ascontiguousarray(qidxs).view(np.uint16)
    M, K = x.shape
    N = q.shape[0]
    y = np.empty((M, N), dtype=np.float32)
    t = np.ascontiguousarray(table).view(np.uint64)
    lib().qo_e8p_mm(_p(x.view(np.uint16)), _p(q), _p(t), _p(y), ctypes.c_int64(M), ctypes.c_int64(N), ctypes.c_int64(K))
    return y


def quantlinear_forward_e8p(x16, qidxs, table, in_features, out_features, q_in, q_out, SU=None, SV=None,
                            bias=None, wscale_float=1.0, had_left=None, K_left=1, had_right=None, K_right=1):
    x = np.ascontiguousarray(x16, dtype=np.float16).reshape(-1, in_features)
    M = x.shape[0]
    q = np.ascontiguousarray(qidxs).view(np.uint16)
    t = np.ascontiguousarray(table).view(np.uint64)
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
    SU, SV, bias, hl, hr = f32(SU), f32(SV), f32(bias), f32(had_left), f32(had_right)
    y = np.empty((M, out_features), dtype=np.float32)
    i64 = ctypes.c_int64
    lib().qo_quantlinear_forward_e8p(_p(x.view(np.uint16)), _p(y), i64(M), i64(in_features), i64(out_features),
                                     i64(q_in), i64(q_out), _p(q), _p(t), _p(SU), _p(SV), _p(bias),
                                     ctypes.c_float(wscale_float), _p(hl), i64(K_left), _p(hr), i64(K_right))
    return y
