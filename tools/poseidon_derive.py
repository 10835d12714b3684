"""Derivation of every Poseidon-Goldilocks table used in this repo, from first principles.

Nothing here is copied from plonky2 (whose source is not on this machine): the round
constants are regenerated from their published recipe (ChaCha8Rng::seed_from_u64(0) + rand 0.8
gen_range(0..p)) and the "fast partial round" tables are re-derived from the MDS
matrix by the sparse factorisation of the Poseidon paper (App. B of eprint 2019/458).  The derived
fast form is checked against the naive form on random states before anything is emitted.

Used by:
  * tools/gen_tables.py        -> intmax_zkp_core_b200/csrc/poseidon_tables.cuh  (product)
  *                            -> oracle/poseidon_fast_tables.h                  (CPU baseline only)
  * tests/test_oracle_poseidon.py (cross-check of the C oracle's own run-time ChaCha8 derivation)
"""
from __future__ import annotations

P = 0xFFFFFFFF00000001
WIDTH = 12
N_FULL_HALF = 4
N_PARTIAL = 22
N_ROUNDS = 2 * N_FULL_HALF + N_PARTIAL
MDS_CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
MDS_DIAG = [8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]
M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF


# ----------------------------------------------------------------------------- ChaCha8 / rand 0.8
def _pcg32_seed_words(state: int):
    """rand_core::SeedableRng::seed_from_u64: PCG32 output words."""
    MUL, INC = 6364136223846793005, 11634580027462260723
    out = []
    for _ in range(8):
        state = (state * MUL + INC) & M64
        x = ((((state >> 18) ^ state) >> 27)) & M32
        rot = state >> 59
        out.append(((x >> rot) | (x << ((32 - rot) & 31))) & M32 if rot else x)
    return out


def _rotl(x, n):
    return ((x << n) | (x >> (32 - n))) & M32


def _chacha_block(key, counter, rounds=8):
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key) + [
        counter & M32, (counter >> 32) & M32, 0, 0]
    w = list(st)

    def qr(a, b, c, d):
        w[a] = (w[a] + w[b]) & M32; w[d] = _rotl(w[d] ^ w[a], 16)
        w[c] = (w[c] + w[d]) & M32; w[b] = _rotl(w[b] ^ w[c], 12)
        w[a] = (w[a] + w[b]) & M32; w[d] = _rotl(w[d] ^ w[a], 8)
        w[c] = (w[c] + w[d]) & M32; w[b] = _rotl(w[b] ^ w[c], 7)

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(w[i] + st[i]) & M32 for i in range(16)]


def round_constants():
    """ALL_ROUND_CONSTANTS[30*12]: ChaCha8Rng::seed_from_u64(0), gen_range(0..p) x 360."""
    key = _pcg32_seed_words(0)
    words, ctr = [], 0

    def next_u64():
        nonlocal words, ctr
        if len(words) < 2:
            words += _chacha_block(key, ctr); ctr += 1
        lo, hi = words[0], words[1]
        words = words[2:]
        return lo | (hi << 32)

    out = []
    zone = P - 1  # (p << p.leading_zeros()) - 1 with leading_zeros == 0
    while len(out) < N_ROUNDS * WIDTH:
        v = next_u64()
        prod = v * P
        if (prod & M64) <= zone:
            out.append(prod >> 64)
    return out


# ----------------------------------------------------------------------------- field / linear algebra
def inv(a):
    return pow(a, P - 2, P)


def mds_matrix():
    """M[r][c] such that out[r] = sum_c M[r][c] * in[c]  (SURVEY.md A9)."""
    M = [[0] * WIDTH for _ in range(WIDTH)]
    for r in range(WIDTH):
        for i in range(WIDTH):
            M[r][(i + r) % WIDTH] = (M[r][(i + r) % WIDTH] + MDS_CIRC[i]) % P
        M[r][r] = (M[r][r] + MDS_DIAG[r]) % P
    return M


def mat_mul(A, B):
    n, m, k = len(A), len(B[0]), len(B)
    return [[sum(A[i][t] * B[t][j] for t in range(k)) % P for j in range(m)] for i in range(n)]


def mat_vec(A, v):
    return [sum(a * b for a, b in zip(row, v)) % P for row in A]


def mat_inv(A):
    n = len(A)
    aug = [list(row) + [1 if i == j else 0 for j in range(n)] for i, row in enumerate(A)]
    for c in range(n):
        piv = next(r for r in range(c, n) if aug[r][c] % P)
        aug[c], aug[piv] = aug[piv], aug[c]
        s = inv(aug[c][c])
        aug[c] = [x * s % P for x in aug[c]]
        for r in range(n):
            if r != c and aug[r][c]:
                f = aug[r][c]
                aug[r] = [(x - f * y) % P for x, y in zip(aug[r], aug[c])]
    return [row[n:] for row in aug]


# ----------------------------------------------------------------------------- permutation, naive
def sbox(x):
    return pow(x, 7, P)


def permute_naive(state, rc=None, M=None):
    rc = rc or round_constants()
    M = M or mds_matrix()
    s = [x % P for x in state]
    for r in range(N_ROUNDS):
        s = [(x + rc[r * WIDTH + i]) % P for i, x in enumerate(s)]
        if r < N_FULL_HALF or r >= N_FULL_HALF + N_PARTIAL:
            s = [sbox(x) for x in s]
        else:
            s[0] = sbox(s[0])
        s = mat_vec(M, s)
    return s


# ----------------------------------------------------------------------------- fast partial rounds
def fast_partial_tables(rc=None, M=None):
    """Equivalent form of the 22 partial rounds:

        state += first_vec                      (12 constants)
        state[1:] = init_mat @ state[1:]        (dense 11x11, once)
        for i in 0..22:
            state[0] = sbox(state[0]) + post_scalar[i]
            new0     = 25*state[0] + sum_j vhat[i][j] * state[1+j]
            state[1+j] += what[i][j] * state[0]
            state[0] = new0

    Derivation: (1) a round-constant vector added before round i+1 equals M^-1 of it added before the
    linear layer of round i; its coordinates 1..11 commute with the one-word S-box and migrate back to
    round i's own constant, leaving a scalar on word 0.  (2) a dense matrix D = [[d00, v^T],[w, Dh]]
    factors as [[d00, v^T Dh^-1],[w, I]] . [[1,0],[0,Dh]]; the block-diagonal factor commutes with the
    one-word S-box and is absorbed into the previous round's matrix.
    """
    rc = rc or round_constants()
    M = M or mds_matrix()
    Minv = mat_inv(M)
    c = [rc[(N_FULL_HALF + i) * WIDTH:(N_FULL_HALF + i + 1) * WIDTH] for i in range(N_PARTIAL)]
    post = [0] * N_PARTIAL
    acc = list(c[N_PARTIAL - 1])
    for i in range(N_PARTIAL - 1, 0, -1):
        e = mat_vec(Minv, acc)
        post[i - 1] = e[0]
        acc = [c[i - 1][0]] + [(c[i - 1][j] + e[j]) % P for j in range(1, WIDTH)]
    first_vec = acc

    vhat = [None] * N_PARTIAL
    what = [None] * N_PARTIAL
    D = [row[:] for row in M]
    for i in range(N_PARTIAL - 1, -1, -1):
        Dh = [row[1:] for row in D[1:]]
        Dh_inv = mat_inv(Dh)
        v = D[0][1:]
        w = [D[r][0] for r in range(1, WIDTH)]
        assert D[0][0] == 25
        vhat[i] = [sum(v[t] * Dh_inv[t][j] for t in range(WIDTH - 1)) % P for j in range(WIDTH - 1)]
        what[i] = w
        Mp = [[1] + [0] * (WIDTH - 1)] + [[0] + Dh[r] for r in range(WIDTH - 1)]
        D = mat_mul(Mp, M)
    init_mat = [row[1:] for row in Mp[1:]]
    return dict(first_vec=first_vec, post=post, vhat=vhat, what=what, init_mat=init_mat)


def permute_fast(state, rc=None, M=None, T=None):
    rc = rc or round_constants()
    M = M or mds_matrix()
    T = T or fast_partial_tables(rc, M)
    s = [x % P for x in state]
    for r in range(N_FULL_HALF):
        s = [(x + rc[r * WIDTH + i]) % P for i, x in enumerate(s)]
        s = mat_vec(M, [sbox(x) for x in s])
    s = [(x + y) % P for x, y in zip(s, T["first_vec"])]
    s = [s[0]] + mat_vec(T["init_mat"], s[1:])
    for i in range(N_PARTIAL):
        s0 = (sbox(s[0]) + T["post"][i]) % P
        new0 = (25 * s0 + sum(a * b for a, b in zip(T["vhat"][i], s[1:]))) % P
        s = [new0] + [(x + w * s0) % P for x, w in zip(s[1:], T["what"][i])]
    for r in range(N_FULL_HALF + N_PARTIAL, N_ROUNDS):
        s = [(x + rc[r * WIDTH + i]) % P for i, x in enumerate(s)]
        s = mat_vec(M, [sbox(x) for x in s])
    return s


def pushed_partial_constants(rc=None, M=None):
    """Forward-pushed form that keeps the small-constant MDS in every partial round:

        for i in 0..22: state[0] = sbox(state[0] + scal[i]); state = M @ state
        then the first of the last four full rounds adds tail_vec instead of its own constants.

    (word-0 scalar only; the other 11 words of each round constant are pushed *forward* through M
    and end up in tail_vec.)  Returned so the CUDA side can pick either form; both are checked.
    """
    rc = rc or round_constants()
    M = M or mds_matrix()
    scal = []
    carry = [0] * WIDTH
    for i in range(N_PARTIAL):
        c = rc[(N_FULL_HALF + i) * WIDTH:(N_FULL_HALF + i + 1) * WIDTH]
        d = [(a + b) % P for a, b in zip(c, carry)]
        scal.append(d[0])
        carry = mat_vec(M, [0] + d[1:])
    r = N_FULL_HALF + N_PARTIAL
    tail_vec = [(a + b) % P for a, b in zip(rc[r * WIDTH:(r + 1) * WIDTH], carry)]
    return scal, tail_vec


def permute_pushed(state, rc=None, M=None):
    rc = rc or round_constants()
    M = M or mds_matrix()
    scal, tail_vec = pushed_partial_constants(rc, M)
    s = [x % P for x in state]
    for r in range(N_FULL_HALF):
        s = [(x + rc[r * WIDTH + i]) % P for i, x in enumerate(s)]
        s = mat_vec(M, [sbox(x) for x in s])
    for i in range(N_PARTIAL):
        s[0] = sbox((s[0] + scal[i]) % P)
        s = mat_vec(M, s)
    for r in range(N_FULL_HALF + N_PARTIAL, N_ROUNDS):
        cv = tail_vec if r == N_FULL_HALF + N_PARTIAL else rc[r * WIDTH:(r + 1) * WIDTH]
        s = [(x + cv[i]) % P for i, x in enumerate(s)]
        s = mat_vec(M, [sbox(x) for x in s])
    return s


def self_check(trials=8):
    import random
    rnd = random.Random(1234)
    rc, M = round_constants(), mds_matrix()
    assert rc[0] == 0xB585F766F2144405 and rc[359] == 0xBC8DFB627FE558FC, "constants drifted (SURVEY App. A)"
    T = fast_partial_tables(rc, M)
    vecs = [[0] * 12, list(range(12)), [P - 1] * 12] + [[rnd.randrange(P) for _ in range(12)] for _ in range(trials)]
    for v in vecs:
        a = permute_naive(v, rc, M)
        assert a == permute_fast(v, rc, M, T), "fast partial rounds disagree with naive form"
        assert a == permute_pushed(v, rc, M), "pushed-constant form disagrees with naive form"
    # SURVEY.md App. B self-derived KAT
    assert permute_naive(list(range(12)), rc, M)[0] == 15442313428170673822
    return True


if __name__ == "__main__":
    print("self_check:", self_check())
