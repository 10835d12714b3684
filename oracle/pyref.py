"""ORACLE — TEST INFRASTRUCTURE ONLY.  Pure-Python second opinion for tiny cases.

Written from the definitions only (no FFT, no shared code with oracle.c): coefficients by the O(n^2) inverse
DFT, every LDE value by Horner evaluation at 7 * w_N^i, leaves by the bit-reversal rule, digests by the
index formula of plonky2's `MerkleTree::prove` (not by the recursive fill).  Poseidon comes from
tools/poseidon_derive.py (its own ChaCha8 + naive permutation in Python ints).  SURVEY.md 8a A1-A11.
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import poseidon_derive as pd  # noqa: E402

P = pd.P
G2 = 1753635133440165772


def root(n_log):
    return pow(G2, 1 << (32 - n_log), P)


def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


_RC, _M = None, None


def permute(s):
    global _RC, _M
    if _RC is None:
        _RC, _M = pd.round_constants(), pd.mds_matrix()
    return pd.permute_naive(s, _RC, _M)


def hash_no_pad(x):
    st = [0] * 12
    for off in range(0, len(x), 8):
        chunk = x[off:off + 8]
        st[:len(chunk)] = [v % P for v in chunk]
        st = permute(st)
    return st[:4]


def hash_or_noop(x):
    return [v % P for v in x] + [0] * (4 - len(x)) if len(x) <= 4 else hash_no_pad(x)


def two_to_one(l, r):
    return permute(list(l) + list(r) + [0] * 4)[:4]


def idft(values):
    n = len(values)
    n_log = n.bit_length() - 1
    w_inv = pow(root(n_log), P - 2, P)
    n_inv = pow(n, P - 2, P)
    return [sum(v * pow(w_inv, i * j, P) for i, v in enumerate(values)) * n_inv % P for j in range(n)]


def horner(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % P
    return acc


def commit(cols, rate_bits, cap_height, is_coeffs=False, salt=None):
    """cols: k lists of n ints.  Returns dict(coeffs, leaves, digests, cap) of Python ints."""
    n = len(cols[0])
    n_log = n.bit_length() - 1
    N_log = n_log + rate_bits
    N = 1 << N_log
    coeffs = [[v % P for v in c] if is_coeffs else idft([v % P for v in c]) for c in cols]
    wN = root(N_log)
    leaves = []
    for j in range(N):
        i = bitrev(j, N_log)
        x = 7 * pow(wN, i, P) % P
        row = [horner(c, x) for c in coeffs]
        if salt is not None:
            row += [s[i] % P for s in salt]
        leaves.append(row)
    sub_log = N_log - cap_height
    sub = 1 << sub_log
    L = 2 * (sub - 1)
    digests = [None] * (L << cap_height)
    cap = []
    for s in range(1 << cap_height):
        layer = [hash_or_noop(leaves[s * sub + m]) for m in range(sub)]
        for i in range(sub_log):
            for m, d in enumerate(layer):
                digests[s * L + 2 * (((m >> 1) << (i + 1)) + (1 << i) - 1) + (m & 1)] = d
            layer = [two_to_one(layer[2 * q], layer[2 * q + 1]) for q in range(len(layer) // 2)]
        cap.append(layer[0])
    assert all(d is not None for d in digests)
    return dict(coeffs=coeffs, leaves=leaves, digests=digests, cap=cap)
