"""ORACLE — TEST INFRASTRUCTURE ONLY (row N1a of SURVEY.md section 8f).

Pure-Python restatement of the permutation-argument step of plonky2's prover, the computation between the wires
commitment and the `Z + partial products` commitment of prove():

    plonky2 @ f99ed9c  plonky2/src/plonk/prover.rs   all_wires_permutation_partial_products,
                                                     wires_permutation_partial_products_and_zs
                       plonky2/src/plonk/vanishing_poly.rs / plonk_common.rs
                                                     quotient_chunk_products, partial_products_and_z_gx
                       plonky2/src/plonk/permutation_argument.rs + circuit_builder.rs
                                                     sigma polynomials, k_is = get_unique_coset_shifts (7^j)

reached from the reference through every prove() (/root/reference/src/transaction/circuits/mod.rs:453,
src/zkdsa/circuits/mod.rs:326, src/rollup/circuits/mod.rs:1247).  The crate source is not on this machine, so this
follows the published algorithm; **parity unpinned** (no fixture in the reference holds a Z polynomial).  What pins the
conventions here is the argument itself: `check_recurrences` re-derives every partial product from its definition, and
for wires that satisfy the copy constraints of sigma the running product returns to 1 after the last row
(`valid_permutation_instance` builds such an instance).

Per row i (x = w_n^i), per challenge (beta, gamma), with R routed wires and chunks of `degree` wires:
    num_j = wire[j][i] + beta * k_j * x + gamma          den_j = wire[j][i] + beta * sigma[j][i] + gamma
    q_l   = prod_{j in chunk l} num_j / den_j            l = 0 .. ceil(R / degree) - 1
    pp_l  = Z(x) * q_0 * ... * q_l   (l < num_prods = chunks - 1)        Z(g x) = Z(x) * q_0 * ... * q_last,  Z(1) = 1
Output columns in the order prove() commits them: Z of every challenge first, then the partial products of challenge 0,
of challenge 1, ...
"""
from __future__ import annotations

import random
from typing import List, Sequence

P = 0xFFFFFFFF00000001
GENERATOR = 7
POWER_OF_TWO_GENERATOR = 1753635133440165772


def root(n_log: int) -> int:
    return pow(POWER_OF_TWO_GENERATOR, 1 << (32 - n_log), P)


def coset_shifts(num_routed: int) -> List[int]:
    """k_is: the identity permutation sends (wire j, row i) to k_j * w^i with k_j = 7^j"""
    return [pow(GENERATOR, j, P) for j in range(num_routed)]


def num_partial_products(num_routed: int, degree: int) -> int:
    return (num_routed + degree - 1) // degree - 1


def partial_products_and_zs(wires: Sequence[Sequence[int]], sigmas: Sequence[Sequence[int]], k_is: Sequence[int],
                            betas: Sequence[int], gammas: Sequence[int], degree: int) -> List[List[int]]:
    """wires, sigmas: [num_routed][n] (column = wire).  Returns the committed columns, each of length n."""
    R, n = len(wires), len(wires[0])
    n_log = n.bit_length() - 1
    w = root(n_log)
    chunks = [list(range(s, min(s + degree, R))) for s in range(0, R, degree)]
    num_prods = len(chunks) - 1
    zs, pps = [], []
    for beta, gamma in zip(betas, gammas):
        z_col = [0] * n
        pp_cols = [[0] * n for _ in range(num_prods)]
        z, x = 1, 1
        for i in range(n):
            z_col[i] = z
            acc = z
            for l, ch in enumerate(chunks):
                num = den = 1
                for j in ch:
                    wv = wires[j][i] % P
                    num = num * ((wv + beta * k_is[j] % P * x + gamma) % P) % P
                    den = den * ((wv + beta * (sigmas[j][i] % P) + gamma) % P) % P
                acc = acc * num % P * pow(den, P - 2, P) % P
                if l < num_prods:
                    pp_cols[l][i] = acc
            z = acc
            x = x * w % P
        zs.append(z_col)
        pps.extend(pp_cols)
    return zs + pps


def check_recurrences(cols: Sequence[Sequence[int]], wires, sigmas, k_is, betas, gammas, degree: int) -> bool:
    """the identities plonky2's vanishing polynomial enforces on these columns (vanishing_poly.rs, check_partial_products)"""
    R, n = len(wires), len(wires[0])
    w = root(n.bit_length() - 1)
    chunks = [list(range(s, min(s + degree, R))) for s in range(0, R, degree)]
    num_prods = len(chunks) - 1
    C = len(betas)
    for c, (beta, gamma) in enumerate(zip(betas, gammas)):
        z = cols[c]
        pp = cols[C + c * num_prods:C + (c + 1) * num_prods]
        if z[0] != 1:
            return False
        x = 1
        for i in range(n):
            prev = z[i]
            for l, ch in enumerate(chunks):
                num = den = 1
                for j in ch:
                    num = num * ((wires[j][i] + beta * k_is[j] % P * x + gamma) % P) % P
                    den = den * ((wires[j][i] + beta * sigmas[j][i] + gamma) % P) % P
                nxt = pp[l][i] if l < num_prods else z[(i + 1) % n]
                if i == n - 1 and l == num_prods:
                    break                                   # wrap-around holds only for a satisfied permutation
                if nxt * den % P != prev * num % P:
                    return False
                prev = nxt
            x = x * w % P
    return True


def valid_permutation_instance(num_routed: int, n_log: int, seed: int = 0):
    """random copy constraints: a permutation of the R * n wire slots, wire values constant on its cycles, and the
    sigma columns sigma[j][i] = k_j' * w^i' for slot (j, i) -> (j', i')"""
    rnd = random.Random(seed)
    n = 1 << n_log
    k_is = coset_shifts(num_routed)
    w = root(n_log)
    xs = [1] * n
    for i in range(1, n):
        xs[i] = xs[i - 1] * w % P
    slots = [(j, i) for j in range(num_routed) for i in range(n)]
    perm = slots[:]
    rnd.shuffle(perm)
    # keep about half of the slots fixed so cycles stay short, as in a real circuit
    for t in range(len(slots)):
        if rnd.random() < 0.5:
            perm[t] = None
    free = [s for s, p in zip(slots, perm) if p is not None]
    targets = free[:]
    rnd.shuffle(targets)
    mapping = {s: s for s in slots}
    for s, t in zip(free, targets):
        mapping[s] = t
    wires = [[0] * n for _ in range(num_routed)]
    seen = set()
    for s in slots:
        if s in seen:
            continue
        v = rnd.randrange(P)
        cur = s
        while cur not in seen:
            seen.add(cur)
            wires[cur[0]][cur[1]] = v
            cur = mapping[cur]
    sigmas = [[k_is[mapping[(j, i)][0]] * xs[mapping[(j, i)][1]] % P for i in range(n)] for j in range(num_routed)]
    return wires, sigmas, k_is


def final_product(cols, wires, sigmas, k_is, beta, gamma, degree: int, challenge: int = 0) -> int:
    """Z(w^n) = Z(w^(n-1)) * prod of the last row's chunks: 1 for a satisfied permutation"""
    R, n = len(wires), len(wires[0])
    w = root(n.bit_length() - 1)
    x = pow(w, n - 1, P)
    acc = cols[challenge][n - 1]
    for j in range(R):
        num = (wires[j][n - 1] + beta * k_is[j] % P * x + gamma) % P
        den = (wires[j][n - 1] + beta * sigmas[j][n - 1] + gamma) % P
        acc = acc * num % P * pow(den, P - 2, P) % P
    return acc
