#!/usr/bin/env python3
"""Exact integer model of the partial rounds as the CUDA permutation runs them (csrc/poseidon.cuh), used to
(1) check the algebra against the naive permutation (tools/poseidon_derive.py) and (2) bound every 32-bit limb by
interval arithmetic so that no signed 32-bit intermediate can overflow.

Idea.  The MDS matrix is circ(K) + diag(8, 0, ..): multiplication by K in Z[t] / (t^12 - 1).  The ring splits as
(t^3 - 1)(t^3 + 1)(t^6 + 1); in that basis (U, V, W) the matrix is three small products whose constants are all +-2^j:
    U' = U (*) (64, 128, 64)  mod t^3 - 1,   V' = V (*) (-4, -32, 8)  mod t^3 + 1,   W' = W (*) 2Q  mod t^6 + 1.
A full round has to come back to the word basis for its twelve S-boxes; a partial round only needs word 0,
    x0 = (U0 + V0 + 2 W0) / 4,
so the 22 partial rounds stay in the split basis: no butterflies, no recombination, and the words are kept as three
signed limbs (22 + 21 + 21 bits) that are only re-normalised (carries + one fold of 2^64 = 2^32 - 1), never packed.

    python tools/poseidon_crt_model.py        # self-check + limb bounds
"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import poseidon_derive as PD

P = PD.P
M22, M21 = (1 << 22) - 1, (1 << 21) - 1
Q = [2, -4, 16, 1, -1, -1]
BIAS_X = 1 << 23          # added to every limb of 4 x0 before packing (limbs may be slightly negative)
BIAS_OUT_LOG = 30         # added to every limb of a word that leaves the split basis


class Iv:
    """closed integer interval with the arithmetic the limb code uses; asserts the int32 range on every result"""
    __slots__ = ("lo", "hi")

    def __init__(self, lo, hi=None):
        self.lo, self.hi = lo, (lo if hi is None else hi)
        assert -(1 << 31) <= self.lo and self.hi < (1 << 31), (self.lo, self.hi)

    def __add__(self, o):
        o = o if isinstance(o, Iv) else Iv(o)
        return Iv(self.lo + o.lo, self.hi + o.hi)
    __radd__ = __add__

    def __sub__(self, o):
        o = o if isinstance(o, Iv) else Iv(o)
        return Iv(self.lo - o.hi, self.hi - o.lo)

    def __rsub__(self, o):
        return Iv(o) - self

    def __mul__(self, c):
        assert isinstance(c, int)
        a, b = self.lo * c, self.hi * c
        return Iv(min(a, b), max(a, b))
    __rmul__ = __mul__

    def __neg__(self):
        return Iv(-self.hi, -self.lo)

    def __rshift__(self, s):
        return Iv(self.lo >> s, self.hi >> s)

    def __and__(self, m):
        return Iv(0, m)

    def __lshift__(self, s):
        return self * (1 << s)

    def __repr__(self):
        return f"[{self.lo}, {self.hi}]"

    def join(self, o):
        return Iv(min(self.lo, o.lo), max(self.hi, o.hi))


def split(x):
    return [x & M22, (x >> 22) & M21, x >> 43]


def value(l):
    return l[0] + (l[1] << 22) + (l[2] << 43)


def fwd_crt(l):
    """one limb plane, 12 words -> U[3], V[3], W[6]"""
    sp = [l[i] + l[i + 6] for i in range(6)]
    W = [l[i] - l[i + 6] for i in range(6)]
    U = [sp[i] + sp[i + 3] for i in range(3)]
    V = [sp[i] - sp[i + 3] for i in range(3)]
    return U, V, W


def ring_B(W):
    """B = W (*) Q mod (t^6 + 1)"""
    out = []
    for k in range(6):
        acc = 0
        for i in range(6):
            q = Q[k - i] if k >= i else -Q[k - i + 6]
            acc = acc + W[i] * q
        out.append(acc)
    return out


def ring_CqD(U, V):
    T = U[0] + U[1] + U[2]
    Cq = [T + U[2], T + U[0], T + U[1]]                       # U (*) (1, 2, 1)
    D = [V[2] * 8 - V[0] - V[1] * 2, -(V[0] * 8) - V[1] - V[2] * 2, V[0] * 2 - V[1] * 8 - V[2]]   # V (*) (-1, -8, 2)
    return Cq, D


def mult_stay(U, V, W, z8):
    """(U, V, W) <- K (*) (U, V, W) + 8 z (1, 1, 1), one limb plane; z8 = 8 * limb of the new word 0"""
    Cq, D = ring_CqD(U, V)
    B = ring_B(W)
    U2 = [Cq[k] * 64 for k in range(3)]
    V2 = [D[k] * 4 for k in range(3)]
    W2 = [B[k] * 2 for k in range(6)]
    U2[0] = U2[0] + z8
    V2[0] = V2[0] + z8
    W2[0] = W2[0] + z8
    return U2, V2, W2


def mult_leave(U, V, W, z8):
    """time-domain limbs of M x from the split basis (the tail of the ordinary MDS layer)"""
    Cq, D = ring_CqD(U, V)
    B = ring_B(W)
    A = [Cq[k] * 16 + D[k] for k in range(3)] + [Cq[k] * 16 - D[k] for k in range(3)]
    out = [A[k] + B[k] for k in range(6)] + [A[k] - B[k] for k in range(6)]
    out[0] = out[0] + z8
    return out


def normalise(o):
    """three signed limbs of one word -> limbs back near 22 / 21 / 21 bits, same value mod p"""
    c0 = o[0] >> 22
    n0 = o[0] & M22
    t1 = o[1] + c0
    c1 = t1 >> 21
    n1 = t1 & M21
    t2 = o[2] + c1
    top = t2 >> 21
    n2 = t2 & M21
    return [n0 - top, n1 + (top << 10), n2]


def div4(v):
    """v / 4 mod p for a u64 v: q + r * 4^-1 with r * 4^-1 = ((4 - r) << 62) - ((4 - r) << 30) + 1 (r != 0; r = 0 gives p)"""
    q, r = v >> 2, v & 3
    sh = (r << 30) & 0xFFFFFFFF
    t = (((~sh) & 0xFFFFFFFF) << 32) | ((sh + 1) & 0xFFFFFFFF)
    return q + t            # may exceed 2^64 by less than p only when r = 0; the device code folds the carry


INV4 = pow(4, P - 2, P)
BIAS_X_VALUE = BIAS_X * (1 + (1 << 22) + (1 << 43))
BIAS_OUT_VALUE = (1 << BIAS_OUT_LOG) * (1 + (1 << 22) + (1 << 43))


def partial_constants():
    """word-0 constant of partial round i with the packing bias of x0 taken out, and the vector added in front of
    the first of the last four full rounds with the packing bias of the leaving words taken out"""
    scal, tail = PD.pushed_partial_constants()
    sc = [(c - BIAS_X_VALUE * INV4) % P for c in scal]
    tv = [(c - BIAS_OUT_VALUE) % P for c in tail]
    return sc, tv


def partial_rounds_model(s, track=None):
    """s: 12 canonical words holding the state after the 4th full round's S-boxes (before its MDS).
    Returns the state in front of the S-boxes of full round 26 (constants added)."""
    sc, tv = partial_constants()
    planes = [[split(x)[L] for x in s] for L in range(3)]
    z = split(s[0])
    st = []
    crt = [fwd_crt(planes[L]) for L in range(3)]
    for k in range(3):                                        # U = sums of four limbs: 24 bits, too wide for the x256 gain
        limbs = normalise([crt[L][0][k] for L in range(3)])
        for L in range(3):
            crt[L][0][k] = limbs[L]
    for L in range(3):
        st.append(mult_stay(*crt[L], z[L] * 8))
    # st[L] = (U, V, W) un-normalised
    for i in range(PD.N_PARTIAL):
        # normalise every word
        for j, n in ((0, 3), (1, 3), (2, 6)):
            for k in range(n):
                limbs = normalise([st[L][j][k] for L in range(3)])
                for L in range(3):
                    st[L][j][k] = limbs[L]
                if track is not None:
                    for L in range(3):
                        track(f"norm{L}", limbs[L])
        E = [st[L][0][0] + st[L][1][0] + st[L][2][0] * 2 + BIAS_X for L in range(3)]
        for L in range(3):
            assert 0 <= E[L] < (1 << 32)
        v = value(E) % P                                      # combine3: any representative
        e = div4(v) % P
        assert (4 * e - v) % P == 0
        a = (e + sc[i]) % P
        zz = PD.sbox(a)
        zl, el = split(zz), split(e)
        for L in range(3):
            d = zl[L] - el[L] + (BIAS_X >> 2)
            for j in range(3):
                st[L][j][0] = st[L][j][0] + d
        if i + 1 < PD.N_PARTIAL:
            for L in range(3):
                st[L] = mult_stay(*st[L], zl[L] * 8)
        else:
            outl = [mult_leave(*st[L], zl[L] * 8) for L in range(3)]
            res = []
            for w in range(12):
                O = [outl[L][w] + (1 << BIAS_OUT_LOG) for L in range(3)]
                for L in range(3):
                    assert 0 <= O[L] < (1 << 32), O
                res.append((value(O) + tv[w]) % P)
            return res


def permute_model(state):
    rc, M = PD.round_constants(), PD.mds_matrix()
    s = [x % P for x in state]
    for r in range(PD.N_FULL_HALF):
        s = [(x + rc[r * 12 + i]) % P for i, x in enumerate(s)]
        s = [PD.sbox(x) for x in s]
        if r + 1 < PD.N_FULL_HALF:
            s = PD.mat_vec(M, s)
    s = partial_rounds_model(s)
    for r in range(PD.N_FULL_HALF + PD.N_PARTIAL, PD.N_ROUNDS):
        if r > PD.N_FULL_HALF + PD.N_PARTIAL:
            s = [(x + rc[r * 12 + i]) % P for i, x in enumerate(s)]
        s = PD.mat_vec(M, [PD.sbox(x) for x in s])
    return s


def bounds(lean=False):
    """interval analysis of the limb pipeline: fixed point of the normalised-limb ranges over the rounds.
    lean: the B200ZKP_LEAN tuning build (word 0 packed from the previous layer's ring products: the injected
    difference carries no 2^21 offset, and X0 = leaving limb of word 0 + 2^30 must fit a signed 32-bit register)"""
    full = [Iv(0, M22), Iv(0, M21), Iv(0, M21)]
    # entry: time-domain limbs of arbitrary u64 words
    st = []
    crt = [fwd_crt([full[L]] * 12) for L in range(3)]
    for k in range(3):
        limbs = normalise([crt[L][0][k] for L in range(3)])
        for L in range(3):
            crt[L][0][k] = limbs[L]
    for L in range(3):
        st.append(mult_stay(*crt[L], full[L] * 8))
    report = {"entry, un-normalised limb": [st[L][0][0].join(st[L][1][0]).join(st[L][2][0]) for L in range(3)]}
    for it in range(6):
        new = []
        for j, n in ((0, 3), (1, 3), (2, 6)):
            for k in range(n):
                limbs = normalise([st[L][j][k] for L in range(3)])
                new.append(limbs)
                for L in range(3):
                    st[L][j][k] = limbs[L]
        lim = [new[0][L] for L in range(3)]
        for w in new:
            lim = [lim[L].join(w[L]) for L in range(3)]
        report["normalised limb"] = lim
        # widen all words to the joined range (fixed point, independent of position)
        for L in range(3):
            for j, n in ((0, 3), (1, 3), (2, 6)):
                for k in range(n):
                    st[L][j][k] = lim[L]
        E = [st[L][0][0] + st[L][1][0] + st[L][2][0] * 2 + BIAS_X for L in range(3)]
        report["4 x0 + bias"] = E
        assert all(e.lo >= 0 for e in E)
        for L in range(3):
            d = full[L] - full[L] + (0 if lean else (BIAS_X >> 2))
            for j in range(3):
                st[L][j][0] = st[L][j][0] + d
        report["word 0 after injection"] = [st[L][0][0] for L in range(3)]
        outl = [mult_leave(*st[L], full[L] * 8) for L in range(3)]
        o = [outl[L][0] for L in range(3)]
        for L in range(3):
            for w in range(12):
                o[L] = o[L].join(outl[L][w])
        report["leaving limb"] = o
        for L in range(3):
            assert o[L].lo + (1 << BIAS_OUT_LOG) >= 0 and o[L].hi + (1 << BIAS_OUT_LOG) < (1 << 32), o
            if lean:
                assert o[L].hi + (1 << BIAS_OUT_LOG) < (1 << 31), o      # X0 is kept in a signed register
        st = [mult_stay(*st[L], full[L] * 8) for L in range(3)]
        m = [st[L][0][0] for L in range(3)]
        for L in range(3):
            for j, n in ((0, 3), (1, 3), (2, 6)):
                for k in range(n):
                    m[L] = m[L].join(st[L][j][k])
        report["un-normalised limb"] = m
    return report


def self_check(trials=6):
    rnd = random.Random(99)
    vecs = [[0] * 12, list(range(12)), [P - 1] * 12] + [[rnd.randrange(P) for _ in range(12)] for _ in range(trials)]
    for v in vecs:
        assert permute_model(v) == PD.permute_naive(v), "split-basis partial rounds disagree with the naive permutation"
    return True


if __name__ == "__main__":
    for k, v in bounds().items():
        print(f"{k:28s} {v}")
    for k, v in bounds(lean=True).items():
        print(f"lean: {k:22s} {v}")
    self_check()
    print("split-basis partial rounds == naive permutation on", 9, "vectors")
