"""numpy model of an IMMA (mma.sync.m16n8k32, u8 x s8 -> s32) formulation of the bs=1 E8P12 GEMV inner loop.

Plan for the next round: the dp4a loop spends 32 of its 175 instructions per 16-byte load on
IDP.4A and 16 more on the parity correction.  One IMMA takes 8 weight rows x 4 codes; its B operand is the decoded
(sign-applied, parity NOT applied) int8 words of one code per lane, its A operand the activation records of the lane's
own 8 segments: row 0 = high bytes + 128 (unsigned), row 1 = low bytes, row 2 = ones (yields sum(w) to undo the +128).
The parity term -2 * par * sum(x_segment) becomes ONE more IMMA per 32 codes: B' = the parity bytes of the lane's 8
codes, A' = the three bytes of (sum(x_segment) + 2^18) plus a row of ones.  Everything stays exact integer arithmetic.

This script checks the fragment / k-slot mapping and the bias algebra against the oracle's decode.  No GPU needed.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import quip_oracle as qo  # noqa: E402

PERM = [0, 2, 1, 3, 4, 6, 5, 7]          # weight i of a code = packed byte PERM[i]; records use the same byte order


def imma_m16n8k32(A_frag, B_frag, C_frag):
    """A_frag[lane][4] uint32 (u8 x4 each), B_frag[lane][2] uint32 (s8 x4 each), C_frag[lane][4] int64 -> D_frag."""
    A = np.zeros((16, 32), dtype=np.int64)
    B = np.zeros((32, 8), dtype=np.int64)
    C = np.zeros((16, 8), dtype=np.int64)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for i in range(4):
            A[g, 4 * t + i] = (int(A_frag[lane][0]) >> (8 * i)) & 0xFF
            A[g + 8, 4 * t + i] = (int(A_frag[lane][1]) >> (8 * i)) & 0xFF
            A[g, 16 + 4 * t + i] = (int(A_frag[lane][2]) >> (8 * i)) & 0xFF
            A[g + 8, 16 + 4 * t + i] = (int(A_frag[lane][3]) >> (8 * i)) & 0xFF
            for j in range(2):
                b = (int(B_frag[lane][j]) >> (8 * i)) & 0xFF
                B[16 * j + 4 * t + i, g] = b - 256 if b >= 128 else b
        C[g, 2 * t], C[g, 2 * t + 1], C[g + 8, 2 * t], C[g + 8, 2 * t + 1] = C_frag[lane]
    D = A @ B + C
    out = np.zeros((32, 4), dtype=np.int64)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        out[lane] = [D[g, 2 * t], D[g, 2 * t + 1], D[g + 8, 2 * t], D[g + 8, 2 * t + 1]]
    return out


def main():
    rng = np.random.default_rng(0)
    rows, codes_per_row = 8, 32                       # one warp step: 8 weight rows x 32 codes (lane (g, t): row g, codes 8t..8t+7)
    q = rng.integers(0, 65536, (rows, codes_per_row)).astype(np.uint16)
    x = rng.integers(-32767, 32768, codes_per_row * 8).astype(np.int64)
    # exact answer from the oracle decode (int8 quarter units, weight order)
    packed = qo.e8p_decode_packed(q)                                        # uint64, packed byte order, parity applied
    wq = qo.e8p_packed_to_weights_q(packed).astype(np.int64)               # [rows][codes][8]
    want = (wq.reshape(rows, -1) * x[None, :]).sum(axis=1)
    # what the kernel would hold: decoded words WITHOUT the parity shift, parity bits, activation records
    sign = (q & 0xFF).astype(np.uint64)
    par = np.array([[bin(int(v)).count("1") & 1 for v in r] for r in sign], dtype=np.uint64)
    with np.errstate(over="ignore"):
        v_nopar = packed + par * np.uint64(0x0202020202020202)             # every byte is ..01 after the shift: no carries
    seg = x.reshape(codes_per_row, 8)[:, :]                                 # [segment][element]
    xb = np.zeros((codes_per_row, 8), dtype=np.int64)
    for i in range(8):
        xb[:, PERM[i]] = seg[:, i]                                          # packed byte order
    hi_u = ((xb >> 8) + 128).astype(np.int64)                               # 0..255
    lo_u = (xb & 0xFF).astype(np.int64)
    assert hi_u.min() >= 0 and hi_u.max() <= 255
    xsum = seg.sum(axis=1) + (1 << 18)
    assert xsum.min() >= 0 and xsum.max() < (1 << 19)

    def word(bytes4):
        return sum(int(b) << (8 * i) for i, b in enumerate(bytes4))

    D = np.zeros((32, 4), dtype=np.int64)
    for j in range(8):                                                       # the 8 IMMAs of one 16-byte load per lane
        A_frag = np.zeros((32, 4), dtype=np.uint64)
        B_frag = np.zeros((32, 2), dtype=np.uint64)
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            s = 8 * t + j                                                    # this lane's code / segment for IMMA j
            w = int(v_nopar[g, s])
            B_frag[lane] = [w & 0xFFFFFFFF, w >> 32]
            plane = [hi_u[s], lo_u[s], np.ones(8, dtype=np.int64)][g] if g < 3 else np.zeros(8, dtype=np.int64)
            A_frag[lane] = [word(plane[:4]), 0, word(plane[4:]), 0]
        D = imma_m16n8k32(A_frag, B_frag, D)
    # parity IMMA: B' = parity bytes of the lane's 8 codes, A' = bytes of (xsum + 2^18) of its 8 segments and a row of ones
    A_frag = np.zeros((32, 4), dtype=np.uint64)
    B_frag = np.zeros((32, 2), dtype=np.uint64)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        pj = [int(par[g, 8 * t + j]) for j in range(8)]
        B_frag[lane] = [word(pj[:4]), word(pj[4:])]
        xs8 = xsum[8 * t:8 * t + 8]
        planes = [xs8 & 0xFF, (xs8 >> 8) & 0xFF, (xs8 >> 16) & 0xFF, np.ones(8, dtype=np.int64)]
        pl = planes[g] if g < 4 else np.zeros(8, dtype=np.int64)
        A_frag[lane] = [word(pl[:4]), 0, word(pl[4:]), 0]
    P = imma_m16n8k32(A_frag, B_frag, np.zeros((32, 4), dtype=np.int64))
    # lanes g = 0..2 hold the planes of weight rows 2t, 2t+1 (c0, c1); combine per weight row n
    got = np.zeros(rows, dtype=np.int64)
    for n in range(rows):
        t, e = n >> 1, n & 1
        H, L, S = D[0 * 4 + t][e], D[1 * 4 + t][e], D[2 * 4 + t][e]
        P0, P1, P2, P1s = P[0 * 4 + t][e], P[1 * 4 + t][e], P[2 * 4 + t][e], P[3 * 4 + t][e]
        parx = P0 + 256 * P1 + 65536 * P2 - (1 << 18) * P1s                  # sum over codes of par * sum(x_segment)
        got[n] = 256 * (H - 128 * S) + L - 2 * parx
    print("IMMA formulation vs oracle decode:", "exact" if np.array_equal(got, want) else "MISMATCH", got[:4], want[:4])
    assert np.array_equal(got, want)


if __name__ == "__main__":
    main()
