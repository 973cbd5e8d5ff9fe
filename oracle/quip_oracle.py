"""CPU oracle (numpy) for the QuIP# inference hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A from-scratch restatement of the reference's algorithm for the `QuantLinear.forward` path
(chu-tianxiang/QuIP-for-all @ 04754a4).  Every function cites the reference file:line it follows
(paths relative to /root/reference).  Nothing here is imported by the product package
`quip_for_all_b200`; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it, and only as the checker.

Parity pinning: the reference ships NO tests, golden vectors or fixtures for this path
(SURVEY.md section 4), and its Hadamard arithmetic lives in the un-vendored, un-pinned third-party
package `fast-hadamard-transform` (requirements.txt:5; call site register_lib.py:18-20).  The oracle
is therefore pinned against outputs of the reference's own Python run in the build container
(`tests/golden/gen_golden.py` imports /root/reference/{quant,qlinear,codebook/*}.py, stubs the
CUDA-only `quip_lib` ops with the reference's own CPU-runnable pieces -- `get_full_grid` table
gather and the pure-torch `matmul_hadU` -- and commits the vectors under tests/golden/), plus the
sha256 / sample-codeword pins extracted in SURVEY.md appendix A.7.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np

# --------------------------------------------------------------------------------------
# E8P12 codebook  (codebook/e8p12.py)
# --------------------------------------------------------------------------------------

E8P_CODESZ = 8
_E8P_COL_PERM = [0, 2, 1, 3, 4, 6, 5, 7]  # codebook/e8p12.py:72 and :84 (shuffle_map)

# codebook/e8p12.py:28-60 -- the 29 norm-12 rows (numerators; the reference divides by 2).
_NORM12_X2 = [
    [3, 1, 1, 1, 3, 3, 3, 3], [1, 3, 1, 1, 3, 3, 3, 3], [1, 1, 3, 1, 3, 3, 3, 3],
    [1, 1, 1, 3, 3, 3, 3, 3], [3, 3, 3, 1, 3, 3, 1, 1], [3, 3, 3, 1, 3, 1, 3, 1],
    [3, 3, 3, 1, 1, 3, 3, 1], [3, 3, 3, 1, 3, 1, 1, 3], [3, 3, 3, 1, 1, 3, 1, 3],
    [3, 3, 3, 1, 1, 1, 3, 3], [3, 3, 1, 3, 3, 3, 1, 1], [3, 3, 1, 3, 3, 1, 3, 1],
    [3, 3, 1, 3, 1, 3, 3, 1], [3, 3, 1, 3, 3, 1, 1, 3], [3, 3, 1, 3, 1, 3, 1, 3],
    [3, 3, 1, 3, 1, 1, 3, 3], [3, 1, 3, 3, 3, 3, 1, 1], [3, 1, 3, 3, 3, 1, 3, 1],
    [3, 1, 3, 3, 1, 3, 3, 1], [3, 1, 3, 3, 3, 1, 1, 3], [3, 1, 3, 3, 1, 3, 1, 3],
    [1, 3, 3, 3, 1, 1, 3, 3], [1, 3, 3, 3, 3, 3, 1, 1], [1, 3, 3, 3, 3, 1, 3, 1],
    [1, 3, 3, 3, 1, 3, 3, 1], [1, 3, 3, 3, 3, 1, 1, 3], [1, 3, 3, 3, 1, 3, 1, 3],
    [1, 1, 3, 3, 1, 3, 3, 3], [3, 3, 1, 1, 3, 3, 3, 1],
]


def e8p_abs_rows_x2() -> np.ndarray:
    """The 256 absolute-value rows (units of 1/2, i.e. odd integers 1,3,5,7), BEFORE the column
    permutation.  codebook/e8p12.py:65-71: all |v| for v in (Z+1/2)^8, |v_i|<=3.5, even coordinate
    sum, ||v||^2 <= 10, de-duplicated and sorted lexicographically (torch.unique(dim=0)), then the
    29 norm-12 rows appended.  Any half-integer abs pattern admits an even-sum sign assignment
    (flipping one sign changes the sum by an odd integer), so the parity filter removes nothing
    after abs(); enumerating odd numerators in ascending nested order IS lexicographic order."""
    rows = []
    vals = (1, 3, 5, 7)

    def rec(prefix, norm_x4):
        if len(prefix) == 8:
            rows.append(prefix)
            return
        for v in vals:
            n2 = norm_x4 + v * v
            # remaining coordinates are at least 1 each (=0.25 each in true units)
            if n2 + (7 - len(prefix)) > 40:  # ||v||^2 <= 10  <=>  sum (2v)^2 <= 40
                break
            rec(prefix + [v], n2)

    rec([], 0)
    d8abs = np.array(rows, dtype=np.int64)
    assert d8abs.shape == (227, 8), d8abs.shape
    return np.concatenate([d8abs, np.array(_NORM12_X2, dtype=np.int64)], axis=0)


def e8p_abs_table() -> np.ndarray:
    """`grid_packed_abs`: int64[256].  codebook/e8p12.py:63-79."""
    cba = e8p_abs_rows_x2().astype(np.float64) / 2.0          # true values .5,1.5,2.5,3.5
    cba = cba[:, _E8P_COL_PERM]                                # :72
    odd = (cba.sum(1) % 2).astype(np.int64)                    # :73  (row sum is an integer)
    cba[:, 7] *= (1 - 2 * odd)
    q = (cba * 4).astype(np.int64)                             # :74-75  -> +-2,6,10,14
    acc = q[:, 0].copy()                                       # :76-78 (sign-extending OR, as torch int64)
    for i in range(7):
        acc = acc | (q[:, i + 1] << ((i + 1) * 8))
    return acc


def e8p_decode_packed(codes: np.ndarray, table: np.ndarray | None = None) -> np.ndarray:
    """Bit-level decode of uint16 codes -> uint64 of 8 packed int8 (units of 1/4), packed byte order.
    quip_cuda/origin_order.cu:211-231 (decode8weights, 64-bit form)."""
    if table is None:
        table = e8p_abs_table()
    c = np.asarray(codes).astype(np.uint16, copy=False).astype(np.uint64)
    tab = np.asarray(table).astype(np.int64).view(np.uint64)
    bits_sign = c & np.uint64(0xFF)
    # popcount parity of the sign byte
    p = bits_sign.copy()
    p ^= p >> np.uint64(4)
    p ^= p >> np.uint64(2)
    p ^= p >> np.uint64(1)
    parity = p & np.uint64(1)
    sign_vec = bits_sign ^ parity
    packed = tab[(c >> np.uint64(8)).astype(np.int64)]
    with np.errstate(over="ignore"):
        decoded_sign = sign_vec * np.uint64(0x8040201008040201)
        decoded_sign &= np.uint64(0x8080808080808080)
        decoded_sign >>= np.uint64(7)
        decoded_sign *= np.uint64(252)
        packed = packed ^ decoded_sign
        packed |= np.uint64(0x0101010101010101)
        packed = packed - parity * np.uint64(0x0202020202020202)
    return packed


def e8p_packed_to_weights_q(packed: np.ndarray) -> np.ndarray:
    """uint64 packed int8 -> int8[..., 8] in WEIGHT order (weight i = packed byte [0,2,1,3,4,6,5,7][i]).
    quip_cuda/origin_order.cu:846-856 (half2 stores 01,45,23,67 from even/odd byte lanes)."""
    p = np.asarray(packed, dtype=np.uint64)
    b = np.stack([((p >> np.uint64(8 * j)) & np.uint64(0xFF)).astype(np.uint8) for j in range(8)],
                 axis=-1).view(np.int8)
    return b[..., _E8P_COL_PERM]


def e8p_decode(codes: np.ndarray, table: np.ndarray | None = None) -> np.ndarray:
    """codes (any int dtype, reinterpreted as uint16) -> float32[..., 8] weights (exact multiples of 1/4)."""
    return e8p_packed_to_weights_q(e8p_decode_packed(codes, table)).astype(np.float32) / 4.0


def e8p_full_grid(table: np.ndarray | None = None) -> np.ndarray:
    """float32[65536, 8]: the full codebook.  Restates codebook/e8p12.py:82-103 (get_full_grid) literally
    (per-element sign from bit (7-ii), +-1/4 by parity) -- deliberately a DIFFERENT formulation from
    e8p_decode_packed so the two pin each other.  (The reference's np.int8(250) at :96 overflows under
    numpy>=2; the intended two's-complement wrap is used here.)"""
    if table is None:
        table = e8p_abs_table()
    tab = np.asarray(table).astype(np.int64).view(np.uint64)
    c = np.arange(1 << 16, dtype=np.int64)
    signs = c & 255
    absi = c >> 8
    parity = np.zeros_like(c)
    for i in range(8):
        parity ^= (signs >> i) & 1
    signs = signs ^ parity
    abs_code = tab[absi]
    out = np.zeros((1 << 16, 8), dtype=np.float32)
    for i in range(8):
        ii = _E8P_COL_PERM[i]
        byte = ((abs_code >> np.uint64(8 * ii)) & np.uint64(255)).astype(np.uint8).view(np.int8)
        v = byte.astype(np.float32) / 4.0
        neg = ((signs >> (7 - ii)) & 1).astype(bool)
        out[:, i] = np.where(neg, -v, v)
    out += np.where(parity.astype(bool), -0.25, 0.25)[:, None].astype(np.float32)
    return out


def decompress_e8p(qidxs: np.ndarray, table: np.ndarray | None = None) -> np.ndarray:
    """Qidxs int16 [N, K/8] -> float16 [N, K].  quip_cuda/origin_order.cu:837-885 (kernel K6)."""
    q = np.ascontiguousarray(qidxs)
    assert q.ndim == 2
    w = e8p_decode(q.view(np.uint16) if q.dtype.itemsize == 2 else q, table)
    return w.reshape(q.shape[0], q.shape[1] * 8).astype(np.float16)


# --------------------------------------------------------------------------------------
# E8P12RVQ4B  (codebook/e8p12_rvq4.py)
# --------------------------------------------------------------------------------------

RVQ4_DEFAULT_RESID_SCALE = 1 / 3.45  # codebook/e8p12_rvq4.py:23


def _hfma_f16(a16: np.ndarray, b16: np.ndarray, c16: np.ndarray) -> np.ndarray:
    """Emulates __hfma2 lane-wise: RN_fp16(a*b + c), single rounding.  Operands here carry <= 16
    significant bits so the float64 product/sum is exact and numpy's float64->float16 cast rounds once."""
    return (a16.astype(np.float64) * b16.astype(np.float64) + c16.astype(np.float64)).astype(np.float16)


def decompress_e8prvq4(qidxs: np.ndarray, scale: float = RVQ4_DEFAULT_RESID_SCALE,
                       table: np.ndarray | None = None) -> np.ndarray:
    """Qidxs int32 [N, K/8] -> float16 [N, K];  W = g[hi16] + fp16(scale) * g[lo16], one fp16 fma rounding.
    quip_cuda/origin_order.cu:956-995 (kernel K7); code layout codebook/e8p12_rvq4.py:42."""
    q = np.ascontiguousarray(qidxs).view(np.uint32)
    main = (q >> np.uint32(16)).astype(np.uint16)
    resid = (q & np.uint32(0xFFFF)).astype(np.uint16)
    w_hi = e8p_decode(main, table).astype(np.float16)
    w_lo = e8p_decode(resid, table).astype(np.float16)
    s16 = np.float16(np.float32(scale))      # __float2half2_rn(scale) from a C float
    w = _hfma_f16(np.broadcast_to(s16, w_lo.shape), w_lo, w_hi)
    return w.reshape(q.shape[0], q.shape[1] * 8)


# --------------------------------------------------------------------------------------
# D4  (codebook/d4.py)
# --------------------------------------------------------------------------------------

D4_CODESZ = 4


def _d4_code3_signs(i3: int, x: list) -> list:
    """codebook/d4.py:26-37."""
    if i3 & (1 << 5):
        x[2] *= -1
    if i3 & (1 << 6):
        x[1] *= -1
    if sum(x) % 2 != 0:
        x[3] *= -1
    if i3 & (1 << 7):
        x = [-v for v in x]
    assert sum(x) % 2 == 0
    return x


def _d4_code8(i8: int) -> list:
    """codebook/d4.py:40-86 (code8_to_d4)."""
    i3 = i8 & (7 << 5)
    i8 &= 31
    h, t, f = 0.5, 1.5, 2.5
    if i8 < 2:
        x = [h] * 4 if i8 == 0 else [t] * 4
    elif i8 < 8:
        ibx = i8 >> 1
        if i8 & 1:
            x = [h] * 4
            x[0] = t
            x[ibx] = t
        else:
            x = [t] * 4
            x[0] = h
            x[ibx] = h
    elif i8 < 16:
        ibx = i8 & 3
        if i8 < 12:
            x = [h] * 4
            x[ibx] = t
        else:
            x = [t] * 4
            x[ibx] = h
    elif i8 < 20:
        x = [h] * 4
        x[i8 & 3] = f
    else:
        ibx = i8 - 20
        ib4 = ibx & 3
        ib3 = ibx >> 2
        x = [h] * 4
        x[ib4] = t
        if ib3 >= ib4:
            ib3 += 1
        x[ib3] = f
    return _d4_code3_signs(i3, x)


def d4_grid() -> np.ndarray:
    """float32[256, 4].  codebook/d4.py:89-96 (build_D4_CB)."""
    return np.array([_d4_code8(i) for i in range(256)], dtype=np.float32)


def decompress_d4(qidxs: np.ndarray, grid: np.ndarray | None = None) -> np.ndarray:
    """Qidxs uint8 [N, K/4] -> float16 [N, K] (table copy).  quip_cuda/origin_order.cu:794-833 (K8)."""
    if grid is None:
        grid = d4_grid()
    q = np.ascontiguousarray(qidxs).view(np.uint8)
    g16 = np.asarray(grid).astype(np.float16)
    return g16[q.astype(np.int64)].reshape(q.shape[0], q.shape[1] * 4)


# --------------------------------------------------------------------------------------
# HI (4-bit scalar) and E8P12RVQ3B  (codebook/hi.py, codebook/e8p12_rvq3.py) -- "next" rows (section 8f)
# --------------------------------------------------------------------------------------

_HI_NIBBLE_ORDER = [0, 2, 4, 6, 1, 3, 5, 7]  # codebook/hi.py:41-50: nibble j holds element order[j]


def decompress_hi(qidxs: np.ndarray) -> np.ndarray:
    """Qidxs int32 [N, K/8] -> float16 [N, K]; w = nibble - 7.5.
    quip_cuda/origin_order.cu:1028-1051 (K10): stores half2 pairs (n0,n4),(n1,n5),(n2,n6),(n3,n7)."""
    q = np.ascontiguousarray(qidxs).view(np.uint32)
    out = np.zeros(q.shape + (8,), dtype=np.float32)
    for j in range(8):
        nib = ((q >> np.uint32(4 * j)) & np.uint32(0xF)).astype(np.float32)
        out[..., _HI_NIBBLE_ORDER[j]] = nib - 7.5
    return out.reshape(q.shape[0], q.shape[1] * 8).astype(np.float16)


def e81b_grid() -> np.ndarray:
    """float32 [256, 8].  codebook/e8p12_rvq3.py:16-50 (get_e81bgrid): the E8 points (Z^8 and Z^8+1/2, even
    coordinate sum) with ||v||^2 <= 2 in cartesian_prod order -- integer points first, each group in
    lexicographic order -- i.e. 113 integer points (0 and the +-1 pairs) then the 128 even-sum +-1/2
    points, then 15 norm-4 axis points (the -2 on the last axis is commented out in the reference)."""
    import itertools
    ints = [np.zeros(8)]
    for i, j in itertools.combinations(range(8), 2):
        for si in (-1.0, 1.0):
            for sj in (-1.0, 1.0):
                v = np.zeros(8)
                v[i], v[j] = si, sj
                ints.append(v)
    halves = [np.array(sg) for sg in itertools.product((-0.5, 0.5), repeat=8) if sum(sg) % 2 == 0]

    def lex(rows):
        a = np.array(rows)
        return a[np.lexsort(a.T[::-1])]

    norm4 = []
    for sgn in (2.0, -2.0):
        for i in range(8):
            if sgn < 0 and i == 7:
                continue
            v = np.zeros(8)
            v[i] = sgn
            norm4.append(v)
    g = np.concatenate([lex(ints), lex(halves), np.array(norm4)], axis=0).astype(np.float32)
    assert g.shape == (256, 8)
    return g


def e81b_packed(grid: np.ndarray | None = None) -> np.ndarray:
    """int32 [256]: 8 nibbles of (2*v & 0xf) in element order [0,2,4,6,1,3,5,7].  codebook/e8p12_rvq3.py:53-62."""
    if grid is None:
        grid = e81b_grid()
    cba = (np.asarray(grid)[:, [0, 2, 4, 6, 1, 3, 5, 7]] * 2).astype(np.int64) & 0xF
    acc = cba[:, 0].copy()
    for i in range(7):
        acc |= cba[:, i + 1] << ((i + 1) * 4)
    return acc.astype(np.uint32).view(np.int32)


RVQ3_DEFAULT_RESID_SCALE = 1 / 2.04  # codebook/e8p12_rvq3.py:75


def decompress_e8prvq3(qidxs: np.ndarray, scale: float = RVQ3_DEFAULT_RESID_SCALE,
                       table: np.ndarray | None = None, cb2: np.ndarray | None = None) -> np.ndarray:
    """Qidxs int32 [N, 3K/32] (byte triplets [resid, idx_lo, idx_hi]) -> float16 [N, K];
    W = g[idx] + fp16(scale) * e81b[resid], one fp16 fma rounding.  quip_cuda/origin_order.cu:887-923 (K9)."""
    if cb2 is None:
        cb2 = e81b_packed()
    q = np.ascontiguousarray(qidxs)
    b = q.view(np.uint8).reshape(q.shape[0], -1, 3)
    resid = b[..., 0].astype(np.int64)
    idx = (b[..., 1].astype(np.uint16) | (b[..., 2].astype(np.uint16) << np.uint16(8)))
    w_hi = e8p_decode(idx, table).astype(np.float16)
    c = np.asarray(cb2).view(np.uint32)[resid]
    w_lo = np.zeros(c.shape + (8,), dtype=np.float16)
    for i in range(4):  # half2 i = (nibble i, nibble i+4) = elements (2i, 2i+1)   origin_order.cu:908-911
        lo = ((c >> np.uint32(4 * i)) & np.uint32(0xF)).astype(np.int64)
        hi = ((c >> np.uint32(4 * i + 16)) & np.uint32(0xF)).astype(np.int64)
        w_lo[..., 2 * i] = (((lo ^ 8) - 8) / 2.0).astype(np.float16)
        w_lo[..., 2 * i + 1] = (((hi ^ 8) - 8) / 2.0).astype(np.float16)
    s16 = np.float16(np.float32(scale))
    w = _hfma_f16(np.broadcast_to(s16, w_lo.shape), w_lo, w_hi)
    return w.reshape(q.shape[0], -1)


# --------------------------------------------------------------------------------------
# Hadamard  (quant.py)
# --------------------------------------------------------------------------------------

def next_power_of_2(n: int) -> int:
    """quant.py:11-14."""
    return 1 if n == 0 else 2 ** math.ceil(math.log(n, 2))


def get_power_of_2(n: int):
    """quant.py:17-23: (exponent, odd base) with n = base * 2**exponent."""
    k = 0
    while n % 2 == 0:
        n //= 2
        k += 1
    return k, n


def hadK_shape(n: int, use_rand: bool = True, table_sizes=None):
    """Shape logic of quant.py:26-39 (get_hadK) without drawing the matrix:
    returns (K, padded_n, kind) with kind in {None, 'rand', 'table'}."""
    exp, base = get_power_of_2(n)
    if base == 1:
        return 1, n, None
    if use_rand:
        return base, n, "rand"
    pad_n = next_power_of_2(n)
    if table_sizes is None:
        table_sizes = HAD_TABLE_SIZES
    if exp < 2 or (base * 4) not in table_sizes:
        return 1, pad_n, None
    return base * 4, n, "table"


# keys of /root/reference/hadamard.safetensors (1, 2, 4, 12, 20, 28, ..., 252)
HAD_TABLE_SIZES = frozenset([1, 2, 4] + list(range(12, 253, 8)))


def fwht(x: np.ndarray, scale: float = 1.0, dtype=np.float64) -> np.ndarray:
    """Unnormalised Sylvester (natural-order) Walsh-Hadamard transform over the last axis, times `scale`.
    Semantics of fast_hadamard_transform(x, scale) (register_lib.py:18-20) as pinned by the butterfly
    in quant.py:50-59: y = x @ H_n^T * scale with H = scipy.linalg.hadamard(n)."""
    y = np.array(x, dtype=dtype, copy=True)
    n = y.shape[-1]
    assert n & (n - 1) == 0, "FWHT length must be a power of two"
    h = 1
    lead = y.shape[:-1]
    while h < n:
        y = y.reshape(lead + (n // (2 * h), 2, h))
        a = y[..., 0, :].copy()
        b = y[..., 1, :].copy()
        y[..., 0, :] = a + b
        y[..., 1, :] = a - b
        y = y.reshape(lead + (n,))
        h *= 2
    return y * dtype(scale)


def matmul_hadU(X: np.ndarray, hadK, K: int, padN: int, transpose: bool = False,
                scale: float | None = None, dtype=np.float64) -> np.ndarray:
    """quant.py:42-65 / :72-84:  y = (hadK (x) H_{n/K}) x * s / sqrt(n/K), index = k*(n/K) + c; pads to padN.
    `scale=None` -> s = 1 (matmul_hadU_cuda's default)."""
    X = np.asarray(X, dtype=dtype)
    n = X.shape[-1]
    if padN != n:
        pad = [(0, 0)] * (X.ndim - 1) + [(0, padN - n)]
        X = np.pad(X, pad)
    s = (1.0 if scale is None else scale) / math.sqrt(padN // K)
    lead = X.shape[:-1]
    if K == 1:
        return fwht(X, s, dtype)
    y = fwht(X.reshape(lead + (K, padN // K)), s, dtype)
    hk = np.asarray(hadK, dtype=dtype)
    if transpose:
        hk = hk.T
    y = np.einsum("ij,...jc->...ic", hk, y)
    return y.reshape(lead + (padN,))


# --------------------------------------------------------------------------------------
# QuantLinear.forward  (qlinear.py:87-115), eval branch
# --------------------------------------------------------------------------------------

def _r16(a):
    return np.asarray(a).astype(np.float16)


def quantlinear_forward(x, *, W_hat, in_features, out_features, q_in, q_out,
                        SU=None, SV=None, bias=None, wscale_float=1.0, Wscale_per_channel=None,
                        had_left=None, K_left=1, had_right=None, K_right=1,
                        rounding="reference"):
    """Eval-mode forward of QuantLinear for fp16 activations.

    W_hat: float16/32 [q_out, q_in] = decode(Qidxs) (codebook decompress).
    rounding="reference": rounds to fp16 at the reference's rounding points -- x*SU (qlinear.py:91),
        hadamard output (register_lib.py:20), hadK@ (quant.py:83), mm output (origin_order.cu:129-130),
        per-channel scale (:107), output hadamard (+hadK@), *SV (:112), +bias (:114) -- with exact (float64)
        accumulation inside each op (the kernels accumulate in fp32).
    rounding="none": float64 throughout (the mathematical forward; SURVEY.md A.5 "Forward identity").
    Returns float64 [M, out_features] (values are fp16-representable when rounding="reference").
    """
    rd = _r16 if rounding == "reference" else (lambda a: np.asarray(a, dtype=np.float64))
    f64 = np.float64
    x = np.asarray(x)
    x2 = x.reshape(-1, x.shape[-1]).astype(f64)
    assert x2.shape[-1] == in_features
    if SU is not None:
        x2 = rd(x2 * np.asarray(SU, dtype=f64)).astype(f64)
    # matmul_hadUt_cuda(x, had_left, K_left, q_in, wscale_float)   qlinear.py:99-100 -> quant.py:72-88
    if q_in != in_features:
        x2 = np.pad(x2, [(0, 0), (0, q_in - in_features)])
    s_in = wscale_float / math.sqrt(q_in // K_left)
    if K_left == 1:
        x2 = rd(fwht(x2, s_in)).astype(f64)
    else:
        t = rd(fwht(x2.reshape(-1, K_left, q_in // K_left), s_in)).astype(f64)
        hk = np.asarray(had_left, dtype=f64).T          # transpose=True -> hadK.T   quant.py:79-80
        x2 = rd(np.einsum("ij,mjc->mic", hk, t)).astype(f64).reshape(-1, q_in)
    out = rd(x2 @ np.asarray(W_hat, dtype=f64).T).astype(f64)      # codebook(x, Qidxs)   qlinear.py:103
    if Wscale_per_channel is not None:
        out = rd(out * np.asarray(Wscale_per_channel, dtype=f64)).astype(f64)   # qlinear.py:106-107
    s_out = 1.0 / math.sqrt(q_out // K_right)
    if K_right == 1:
        out = rd(fwht(out, s_out)).astype(f64)
    else:
        t = rd(fwht(out.reshape(-1, K_right, q_out // K_right), s_out)).astype(f64)
        hk = np.asarray(had_right, dtype=f64)
        out = rd(np.einsum("ij,mjc->mic", hk, t)).astype(f64).reshape(-1, q_out)
    out = out[:, :out_features]                                                  # qlinear.py:109
    if SV is not None:
        out = rd(out * np.asarray(SV, dtype=f64)).astype(f64)                    # :111-112
    if bias is not None:
        out = rd(out + np.asarray(bias, dtype=f64)).astype(f64)                  # :114
    return out.reshape(x.shape[:-1] + (out_features,))


def codebook_mm(x16: np.ndarray, W_hat: np.ndarray) -> np.ndarray:
    """`x @ decode(Qidxs)^T`: fp16 in, exact accumulate, one fp16 rounding (kernels K1-K3; origin_order.cu:523-553)."""
    return (np.asarray(x16, dtype=np.float64) @ np.asarray(W_hat, dtype=np.float64).T).astype(np.float16)


# --------------------------------------------------------------------------------------
# QuantLinear buffer geometry  (qlinear.py:29-57)
# --------------------------------------------------------------------------------------

CODEBOOK_SPECS = {
    # id: (codesz, packsz, numpy idx dtype, bits/weight)
    "E8P12": (8, Fraction(1), np.int16, 2),
    "E8P12RVQ4B": (8, Fraction(1), np.int32, 4),
    "E8P12RVQ3B": (8, Fraction(4, 3), np.int32, 3),
    "D4": (4, Fraction(1), np.uint8, 2),
    "HI": (8, Fraction(1), np.int32, 4),
}


def qidxs_shape(in_features: int, out_features: int, codebook: str, use_rand: bool = True):
    """Shape of the `Qidxs` buffer: (q_out, q_in // (codesz*packsz)).  qlinear.py:52-57."""
    codesz, packsz, _, _ = CODEBOOK_SPECS[codebook]
    _, q_in, _ = hadK_shape(in_features, use_rand)
    _, q_out, _ = hadK_shape(out_features, use_rand)
    return q_out, int(q_in // (codesz * packsz))


# --------------------------------------------------------------------------------------------------
# decode step of a Llama decoder layer built from QuantLinears (the loop body of the reference's generation example)
# --------------------------------------------------------------------------------------------------
def llama_decoder_layer_step(h, layer, k_cache, v_cache, pos, *, n_heads, n_kv_heads, head_dim, eps, theta=10000.0,
                             rounding="reference"):
    """One bs=1 decode position through one decoder layer whose seven projections are QuantLinears.

    Follows what the reference's decode loop executes per layer (example_generate.py:29-32 ->
    HF LlamaDecoderLayer.forward with the block linears swapped for QuantLinear, quantizer.py:193-248):
        x = RMSNorm(h); q,k,v = QuantLinear(x); RoPE(q,k) (rotate_half); append k,v at `pos`;
        o = softmax(q k^T / sqrt(d)) v over positions 0..pos (GQA: query head j uses kv head j // group);
        h = h + o_proj(o); x = RMSNorm(h); h = h + down(silu(gate(x)) * up(x)).
    h: [hidden] fp16-valued; layer: dict with 'input_norm', 'post_norm' (fp16 vectors) and 'q','k','v','o','gate','up',
    'down' -> kwargs of quantlinear_forward (W_hat, dims, SU, SV, ...); k_cache / v_cache: float arrays
    [n_kv_heads, max_len, head_dim] holding fp16 values, updated in place at `pos`.
    rounding="reference" keeps fp16 tensors where HF holds fp16 tensors (norm output, q/k after RoPE, attention
    output, residual sums, silu and its product); accumulations are float64.  Returns the new hidden state (float64).
    """
    rd = _r16 if rounding == "reference" else (lambda a: np.asarray(a, dtype=np.float64))
    f64 = np.float64

    def rms(x, w):
        v = np.asarray(x, dtype=f64)
        v = rd(v / np.sqrt((v * v).mean() + eps))
        return rd(np.asarray(w, dtype=f64) * v)

    def lin(name, x):
        return quantlinear_forward(np.asarray(x, dtype=f64)[None, :], rounding=rounding, **layer[name])[0]

    h = np.asarray(h, dtype=f64)
    x = rms(h, layer["input_norm"])
    q = lin("q", x).reshape(n_heads, head_dim)
    k = lin("k", x).reshape(n_kv_heads, head_dim)
    v = lin("v", x).reshape(n_kv_heads, head_dim)
    inv = 1.0 / (theta ** (np.arange(0, head_dim, 2, dtype=f64) / head_dim))
    ang = np.concatenate([pos * inv, pos * inv])
    cos, sin = rd(np.cos(ang)), rd(np.sin(ang))                    # HF keeps the rotary tables in the model dtype

    def rope(t):
        half = head_dim // 2
        rot = np.concatenate([-t[:, half:], t[:, :half]], axis=1)
        return rd(t * cos + rot * sin)

    q, k = rope(q), rope(k)
    k_cache[:, pos] = k
    v_cache[:, pos] = v
    group = n_heads // n_kv_heads
    o = np.zeros((n_heads, head_dim), dtype=f64)
    for j in range(n_heads):
        kk, vv = k_cache[j // group, :pos + 1].astype(f64), v_cache[j // group, :pos + 1].astype(f64)
        s = kk @ q[j] / math.sqrt(head_dim)
        p = np.exp(s - s.max())
        o[j] = (p / p.sum()) @ vv
    o = rd(o).reshape(-1)
    h = rd(h + lin("o", o))
    x = rms(h, layer["post_norm"])
    g, u = lin("gate", x), lin("up", x)
    act = rd(rd(g / (1.0 + np.exp(-g))) * u)
    return rd(h + lin("down", act))


# --------------------------------------------------------------------------------------
# Quantise-time path: nearest codeword + LDLQ  -- float64 statements
# --------------------------------------------------------------------------------------

def e8p_nearest(x: np.ndarray, table: np.ndarray | None = None, chunk: int = 256):
    """codebook/e8p12.py:125-128 (`round`): argmax_c 2 x.g_c - |g_c|^2 over the 65 536 codewords, FIRST index on equal
    scores (torch.argmax).  Evaluated in float64 (the codewords are multiples of 1/4: products and norms are exact for
    fp32 inputs up to the final sum).  Returns (idx int64 [m], best score float64 [m], scores of a given index via
    `e8p_score`)."""
    g = e8p_full_grid(table).astype(np.float64)
    gn = (g * g).sum(1)
    x = np.asarray(x, dtype=np.float64).reshape(-1, 8)
    idx = np.empty(x.shape[0], dtype=np.int64)
    best = np.empty(x.shape[0], dtype=np.float64)
    for a in range(0, x.shape[0], chunk):
        s = 2.0 * x[a:a + chunk] @ g.T - gn
        idx[a:a + chunk] = s.argmax(1)
        best[a:a + chunk] = s.max(1)
    return idx, best


def e8p_score(x: np.ndarray, idx: np.ndarray, table: np.ndarray | None = None) -> np.ndarray:
    """float64 score 2 x.g - |g|^2 of the given codeword indices (one per row of x)."""
    g = e8p_full_grid(table).astype(np.float64)[np.asarray(idx, dtype=np.int64)]
    x = np.asarray(x, dtype=np.float64).reshape(-1, 8)
    return 2.0 * (x * g).sum(1) - (g * g).sum(1)


def e8prvq4_quantize(x: np.ndarray, scale: float = RVQ4_DEFAULT_RESID_SCALE):
    """codebook/e8p12_rvq4.py:37-46 with fp32 tensors: init = round(X); resid = (X - init) / fl32(scale) (fp32);
    resid code = round(resid); vals = init + resid_vals * fl32(scale) (fp32 mul, fp32 add); idx = (init << 16) + resid."""
    g32 = e8p_full_grid()
    x32 = np.asarray(x, dtype=np.float32).reshape(-1, 8)
    s32 = np.float32(scale)
    i0, _ = e8p_nearest(x32)
    v0 = g32[i0]
    r = ((x32 - v0) / s32).astype(np.float32)
    i1, _ = e8p_nearest(r)
    vals = (v0 + (g32[i1] * s32).astype(np.float32)).astype(np.float32)
    return vals, (i0 << 16) + i1, r


def block_ldl(L: np.ndarray, b: int) -> np.ndarray:
    """quant.py:91-104: every b-column block of L is multiplied by the inverse of its diagonal block."""
    n = L.shape[0]
    out = np.array(L, dtype=np.float64, copy=True)
    for i in range(n // b):
        out[:, i * b:(i + 1) * b] = out[:, i * b:(i + 1) * b] @ np.linalg.inv(L[i * b:(i + 1) * b, i * b:(i + 1) * b])
    return out


def ldlq(W: np.ndarray, H: np.ndarray, L: np.ndarray, tune_iters: int = 0):
    """quant.py:107-139 (`LDLQ`, E8P12 codebook), float64: the last 8-column group is rounded first, each group sees
    the error of everything to its right through the block-LDL factor; optional re-rounding sweeps (:131-139)."""
    g = e8p_full_grid().astype(np.float64)
    W = np.asarray(W, dtype=np.float64)
    m, n = W.shape
    Lb = block_ldl(np.asarray(L, dtype=np.float64), 8)
    hat = np.zeros_like(W)
    Q = np.zeros((m, n // 8), dtype=np.int64)
    for k in range(n // 8 - 1, -1, -1):
        a, b = 8 * k, 8 * k + 8
        t = W[:, a:b] + (W[:, b:] - hat[:, b:]) @ Lb[b:, a:b]
        Q[:, k], _ = e8p_nearest(t)
        hat[:, a:b] = g[Q[:, k]]
    for _ in range(tune_iters):
        for k in range(n // 8 - 1, -1, -1):
            a, b = 8 * k, 8 * k + 8
            t = hat[:, a:b] + (W - hat) @ H[:, a:b] @ np.linalg.inv(H[a:b, a:b])
            Q[:, k], _ = e8p_nearest(t)
            hat[:, a:b] = g[Q[:, k]]
    return hat, Q


def e8prvq3_quantize(x: np.ndarray, scale: float = RVQ3_DEFAULT_RESID_SCALE):
    """codebook/e8p12_rvq3.py:91-101 with fp32 tensors: init = E8P12 round(X); resid = (X - init) / fl32(scale); the
    residual code is the nearest of the 256 e81b points (argmax 2 r.g - |g|^2, first index on ties);
    vals = init + e81b[resid] * fl32(scale); idx = (init << 8) + resid."""
    g32 = e8p_full_grid()
    e = e81b_grid().astype(np.float64)
    x32 = np.asarray(x, dtype=np.float32).reshape(-1, 8)
    s32 = np.float32(scale)
    i0, _ = e8p_nearest(x32)
    v0 = g32[i0]
    r = ((x32 - v0) / s32).astype(np.float32)
    sc = 2.0 * r.astype(np.float64) @ e.T - (e * e).sum(1)
    i1 = sc.argmax(1)
    vals = (v0 + (e81b_grid().astype(np.float32)[i1] * s32).astype(np.float32)).astype(np.float32)
    return vals, (i0 << 8) + i1


def e8p_nearest_structured(x: np.ndarray, table: np.ndarray | None = None):
    """The same argmax as `e8p_nearest` from the STRUCTURE of the codebook (codebook/e8p12.py:82-103) instead of all
    65 536 dot products: a codeword is g = (sigma . t + delta) / 4 with t one of the 256 signed abs-table rows (quarter
    units), sigma an even-weight sign vector and delta = +-1, so for a fixed (t, delta)
        2 x.g - |g|^2 = sum_j sigma_j w_j + delta sum(x) / 2 - (|t|^2 + 8) / 16,    w_j = t_j (x_j / 2 - delta / 8),
    which is maximised by sigma_j = sign(w_j), flipping the smallest |w_j| when the number of negations is odd.
    512 candidates of 8 operations each.  Exact ties (w_j = 0, equal minima, equal candidates) are not resolved to the
    first index here -- this is a cross-check of the algebra on generic inputs, and the model of a ~40x cheaper GPU search.
    Returns (idx int64 [m], score float64 [m])."""
    if table is None:
        table = e8p_abs_table()
    tab = np.asarray(table).astype(np.int64).view(np.uint64)
    # t[a, i]: element i of abs row a in quarter units (element i = packed byte _E8P_COL_PERM[i], signed)
    t = np.zeros((256, 8), dtype=np.float64)
    for i in range(8):
        b = ((tab >> np.uint64(8 * _E8P_COL_PERM[i])) & np.uint64(255)).astype(np.uint8).view(np.int8)
        t[:, i] = b.astype(np.float64)
    t2 = (t * t).sum(1)
    x = np.asarray(x, dtype=np.float64).reshape(-1, 8)
    m = x.shape[0]
    best_s = np.full(m, -np.inf)
    best_c = np.zeros(m, dtype=np.int64)
    bit_of_elem = np.array([7 - _E8P_COL_PERM[i] for i in range(8)])      # sign bit (7 - packed byte) of element i
    for delta in (1.0, -1.0):
        w = t[None, :, :] * (x[:, None, :] / 2.0 - delta / 8.0)          # [m, 256, 8]
        neg = w < 0
        aw = np.abs(w)
        odd = (neg.sum(2) & 1).astype(bool)
        jmin = aw.argmin(2)
        val = aw.sum(2) - np.where(odd, 2.0 * np.take_along_axis(aw, jmin[..., None], 2)[..., 0], 0.0)
        score = val + delta * x.sum(1)[:, None] / 2.0 - (t2[None, :] + 8.0) / 16.0
        a = score.argmax(1)
        sc = score[np.arange(m), a]
        ng = neg[np.arange(m), a].copy()                                  # [m, 8]
        fl = odd[np.arange(m), a]
        jm = jmin[np.arange(m), a]
        ng[np.arange(m)[fl], jm[fl]] ^= True
        s = (ng.astype(np.int64) << bit_of_elem[None, :]).sum(1)           # even-weight sign word
        sign_byte = s if delta > 0 else (s ^ 1)                            # odd parity <-> the -1/4 shift
        c = (a.astype(np.int64) << 8) | sign_byte
        take = (sc > best_s) | ((sc == best_s) & (c < best_c))
        best_s = np.where(take, sc, best_s)
        best_c = np.where(take, c, best_c)
    return best_c, best_s
