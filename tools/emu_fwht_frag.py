"""numpy emulation of the mma.sync fragment algebra in csrc/fwht_mma.cuh (no GPU needed).

Checks the claim the warp-per-row rotation kernel relies on: a lane that loads the 16-byte octet `lane` of a
256-element block (elements 8*lane .. 8*lane+7) and hands its four fp16 pairs to fwht256_frag as p[0..3] gets back, in
r[2q], r[2q+1], the transformed elements 8*lane + 2q, 8*lane + 2q + 1 -- i.e. the F(x, y) register layout is a
bit permutation of the natural index, and the Sylvester Hadamard matrix is invariant under simultaneous bit
permutations of its row and column index.
"""
import numpy as np
from scipy.linalg import hadamard


def mma_16816(A_frag, B_frag, C_frag):
    """A_frag[lane][4 regs][2], B_frag[lane][2 regs][2], C_frag[lane][4] -> D_frag[lane][4] (fp32)."""
    A = np.zeros((16, 16)); B = np.zeros((16, 8)); C = np.zeros((16, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for e in range(2):
            A[g, 2 * t + e] = A_frag[lane][0][e]
            A[g + 8, 2 * t + e] = A_frag[lane][1][e]
            A[g, 2 * t + 8 + e] = A_frag[lane][2][e]
            A[g + 8, 2 * t + 8 + e] = A_frag[lane][3][e]
            B[2 * t + e, g] = B_frag[lane][0][e]
            B[2 * t + 8 + e, g] = B_frag[lane][1][e]
            C[g, 2 * t + e] = C_frag[lane][e]
            C[g + 8, 2 * t + e] = C_frag[lane][2 + e]
    D = A @ B + C
    out = np.zeros((32, 4))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        out[lane] = [D[g, 2 * t], D[g, 2 * t + 1], D[g + 8, 2 * t], D[g + 8, 2 * t + 1]]
    return out


def make_hfrag():
    H = hadamard(16) * 0.25
    f = np.zeros((32, 4, 2))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for e in range(2):
            f[lane][0][e] = H[g, 2 * t + e]
            f[lane][1][e] = H[g + 8, 2 * t + e]
            f[lane][2][e] = H[g, 2 * t + 8 + e]
            f[lane][3][e] = H[g + 8, 2 * t + 8 + e]
    return f


def hT(p, A):
    """p[lane][4][2] -> r[lane][8]   (hT_packed; hT_split is the same algebra with hi + lo parts)."""
    z = np.zeros((32, 4))
    d0 = mma_16816(A, p[:, [0, 2]], z)
    d1 = mma_16816(A, p[:, [1, 3]], z)
    return np.concatenate([d0, d1], axis=1)


def fwht256_frag(p, A):
    r = hT(p, A)
    return hT(r.reshape(32, 4, 2), A)


def main():
    rng = np.random.default_rng(0)
    A = make_hfrag()
    x = rng.standard_normal(256)
    p = x.reshape(32, 4, 2)                       # lane l: octet l, pair q = elements 8l + 2q, 8l + 2q + 1
    r = fwht256_frag(p, A)
    want = hadamard(256) @ x / 16.0
    err = np.abs(r.reshape(256) - want).max()
    print("fwht256 with natural octet placement: max err", err)
    assert err < 1e-12
    # 4096 = 16 blocks of 256 transformed in registers, then a radix-16 butterfly across the blocks
    x = rng.standard_normal(4096)
    R = np.stack([fwht256_frag(x[z * 256:(z + 1) * 256].reshape(32, 4, 2), A) for z in range(16)])   # [z][lane][8]
    h = 1
    while h < 16:
        for z in range(16):
            if not z & h:
                a, b = R[z].copy(), R[z | h].copy()
                R[z], R[z | h] = a + b, a - b
        h <<= 1
    got = R.reshape(4096) * 0.25
    want = hadamard(4096) @ x / 64.0
    err = np.abs(got - want).max()
    print("warp-per-row 4096: max err", err)
    assert err < 1e-11


if __name__ == "__main__":
    main()


# ---------------------------------------------------------------------------------------------------------
# CTA-wide 4096-point transform (fwht4096_frag: 16 warps, one shared-memory exchange) with natural placement
# ---------------------------------------------------------------------------------------------------------
def frag_x(lane, q): return (lane >> 2) + 8 * (q & 1)
def frag_y(lane, q): return 2 * (lane & 3) + 8 * (q >> 1)
WQ = [0, 2, 1, 3]            # p[q] = word WQ[q] of the lane's octet  ({x, z, y, w})


def idx_block(warp, lane, q):      # element index of r[2q] (r[2q+1] is +1): input of rot_out / output of rot_in
    return warp * 256 + lane * 8 + 2 * WQ[q]


def idx_spread(warp, lane, q):     # output of rot_out / input of rot_in
    g, t = lane >> 2, lane & 3
    return (g + 8 * (q & 1)) * 256 + (warp & 7) * 32 + t * 8 + (warp >> 3) * 4 + (q >> 1) * 2


def fwht4096_frag(P, A):
    """P[warp][lane][4][2] -> R[warp][lane][8]."""
    XROW = 388
    S = np.zeros(16 * XROW)
    for w in range(16):
        r = hT(hT(P[w], A).reshape(32, 4, 2), A)
        for lane in range(32):
            for q in range(4):
                S[w * XROW + frag_x(lane, q) * 24 + frag_y(lane, q)] = r[lane][2 * q]
                S[w * XROW + frag_x(lane, q) * 24 + frag_y(lane, q) + 1] = r[lane][2 * q + 1]
    R = np.zeros((16, 32, 8))
    for w in range(16):
        r = np.zeros((32, 8))
        for lane in range(32):
            for q in range(4):
                r[lane][2 * q] = S[frag_y(lane, q) * XROW + w * 24 + frag_x(lane, q)]
                r[lane][2 * q + 1] = S[(frag_y(lane, q) + 1) * XROW + w * 24 + frag_x(lane, q)]
        R[w] = hT(r.reshape(32, 4, 2), A)
    return R


def check_4096():
    rng = np.random.default_rng(1)
    A = make_hfrag()
    x = rng.standard_normal(4096)
    want = hadamard(4096) @ x / 64.0
    # block -> spread (output-side rotation)
    P = np.zeros((16, 32, 4, 2))
    for w in range(16):
        for lane in range(32):
            for q in range(4):
                i = idx_block(w, lane, q)
                P[w, lane, q] = x[i:i + 2]
    R = fwht4096_frag(P, A)
    err = 0.0
    seen = set()
    for w in range(16):
        for lane in range(32):
            for q in range(4):
                i = idx_spread(w, lane, q)
                seen.update((i, i + 1))
                err = max(err, abs(R[w, lane, 2 * q] - want[i]), abs(R[w, lane, 2 * q + 1] - want[i + 1]))
    assert len(seen) == 4096
    print("block -> spread, natural placement: max err", err)
    assert err < 1e-11
    # spread -> block (input-side rotation)
    for w in range(16):
        for lane in range(32):
            for q in range(4):
                i = idx_spread(w, lane, q)
                P[w, lane, q] = x[i:i + 2]
    R = fwht4096_frag(P, A)
    err = 0.0
    for w in range(16):
        for lane in range(32):
            for q in range(4):
                i = idx_block(w, lane, q)
                err = max(err, abs(R[w, lane, 2 * q] - want[i]), abs(R[w, lane, 2 * q + 1] - want[i + 1]))
    print("spread -> block, natural placement: max err", err)
    assert err < 1e-11


if __name__ == "__main__":
    check_4096()
