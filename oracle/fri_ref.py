"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of plonky2's opening proof (rows N2 + N3 of SURVEY.md section 8f), prover AND verifier, written from the
published algorithm of plonky2 @ f99ed9c (un-vendored dependency of the reference, Cargo.toml:12):
    prover    fri/oracle.rs `PolynomialBatch::prove_openings`, fri/prover.rs `fri_proof`, `fri_committed_trees`,
              `fri_proof_of_work`, `fri_prover_query_round`, util/reducing.rs `ReducingFactor`,
              field/src/polynomial/division.rs `divide_by_linear`, iop/challenger.rs `Challenger`
    verifier  fri/verifier.rs `verify_fri_proof`, `fri_verifier_query_round`, `fri_combine_initial`, `compute_evaluation`,
              `PrecomputedReducedOpenings::from_os_and_alpha`, fri/challenges.rs `fri_challenges`

PARITY UNPINNED: the reference holds no proof fixture (its tests prove and verify in-process) and cannot be built here.
What anchors this file instead: the verifier below is written from the verifier's side of the protocol only (Lagrange
interpolation of each coset, Horner evaluation of final_poly), and it must accept every proof the prover restatement
and the CUDA path produce and reject tampered ones; the Poseidon / Merkle primitives underneath are pinned (oracle.c).
Pure-Python field arithmetic: use at small sizes only.  Only tests/ may import this module.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from . import oracle as O

P = 0xFFFFFFFF00000001
W = 7                       # X^2 = 7
G = 7                       # multiplicative generator / coset shift
G2 = 1753635133440165772    # generator of the 2^32 subgroup
Ext = Tuple[int, int]


def root(n_log: int) -> int:
    w = G2
    for _ in range(n_log, 32):
        w = w * w % P
    return w


def bitrev(x: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


# ------------------------------------------------------------------ quadratic extension
def eadd(x: Ext, y: Ext) -> Ext:
    return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)


def esub(x: Ext, y: Ext) -> Ext:
    return ((x[0] - y[0]) % P, (x[1] - y[1]) % P)


def emul(x: Ext, y: Ext) -> Ext:
    return ((x[0] * y[0] + W * x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)


def escale(x: Ext, c: int) -> Ext:
    return (x[0] * c % P, x[1] * c % P)


def epow(x: Ext, e: int) -> Ext:
    r = (1, 0)
    while e:
        if e & 1:
            r = emul(r, x)
        x = emul(x, x)
        e >>= 1
    return r


def einv(x: Ext) -> Ext:
    # (a + bX)^-1 = (a - bX) / (a^2 - 7 b^2)
    d = pow((x[0] * x[0] - W * x[1] * x[1]) % P, P - 2, P)
    return (x[0] * d % P, (-x[1]) * d % P)


def eval_poly(coeffs: Sequence[Ext], z: Ext) -> Ext:
    acc = (0, 0)
    for c in reversed(coeffs):
        acc = eadd(emul(acc, z), c)
    return acc


# ------------------------------------------------------------------ Challenger (iop/challenger.rs)
class Challenger:
    def __init__(self):
        self.state = np.zeros(12, dtype=np.uint64)
        self.inp: List[int] = []
        self.out: List[int] = []

    def observe_element(self, e: int):
        self.out = []
        self.inp.append(int(e) % P)
        if len(self.inp) == 8:
            self._duplex()

    def observe_elements(self, es):
        for e in np.asarray(es, dtype=np.uint64).reshape(-1).tolist():
            self.observe_element(e)

    def observe_cap(self, cap):
        self.observe_elements(cap)

    def observe_extension_elements(self, es):
        for a, b in es:
            self.observe_element(a)
            self.observe_element(b)

    def get_challenge(self) -> int:
        if self.inp or not self.out:
            self._duplex()
        return self.out.pop()

    def get_hash(self) -> List[int]:
        return [self.get_challenge() for _ in range(4)]

    def get_extension_challenge(self) -> Ext:
        a = self.get_challenge()
        b = self.get_challenge()
        return (a, b)

    def _duplex(self):
        for i, v in enumerate(self.inp):
            self.state[i] = v
        self.inp = []
        self.state = O.permute(self.state)
        self.out = [int(v) for v in self.state[:8]]


# ------------------------------------------------------------------ prover
def reduce_polys_base(polys: Sequence[Sequence[int]], alpha: Ext) -> List[Ext]:
    """ReducingFactor::reduce_polys_base: sum_i alpha^i p_i."""
    n = max(len(p) for p in polys)
    out = [(0, 0)] * n
    pw = (1, 0)
    for p in polys:
        out = [eadd(o, escale(pw, int(c))) for o, c in zip(out, p)]
        pw = emul(pw, alpha)
    return out


def divide_by_linear(coeffs: Sequence[Ext], z: Ext) -> List[Ext]:
    """(p(X) - p(z)) / (X - z): scan from the top, drop p(z)."""
    bs = []
    acc = (0, 0)
    for c in reversed(coeffs):
        acc = eadd(emul(acc, z), c)
        bs.append(acc)
    bs.pop()
    bs.reverse()
    return bs


def final_poly_of(oracle_polys: Sequence[np.ndarray], batches, alpha: Ext, mul_by_x: bool) -> List[Ext]:
    """prove_openings' final_poly.  oracle_polys[o] = (k_o, n) coefficients; batches = [(point, [(o, i), ...]), ...]."""
    n = oracle_polys[0].shape[1]
    final: List[Ext] = []
    for point, polys in batches:
        comp = reduce_polys_base([oracle_polys[o][i].tolist() for o, i in polys], alpha) if polys else [(0, 0)] * n
        quotient = divide_by_linear(comp, point)
        if not mul_by_x:
            quotient.append((0, 0))                    # "pad back to power of two" (later plonky2)
        shift = epow(alpha, len(polys))                # alpha.shift_poly(&mut final_poly)
        final = [emul(c, shift) for c in final]
        if len(final) < len(quotient):
            final = final + [(0, 0)] * (len(quotient) - len(final))
        final = [eadd(a, b) for a, b in zip(final, quotient)]
    if mul_by_x:
        final = [(0, 0)] + final                       # final_poly.coeffs.insert(0, ZERO): multiply by X (2022 plonky2, PR 436)
    assert len(final) == n
    return final


def ext_fft(coeffs: Sequence[Ext]) -> List[Ext]:
    """FFT of extension coefficients over a base-field subgroup: componentwise."""
    a = O.fft(np.array([c[0] for c in coeffs], dtype=np.uint64))
    b = O.fft(np.array([c[1] for c in coeffs], dtype=np.uint64))
    return [(int(x), int(y)) for x, y in zip(a, b)]


def coset_fft(coeffs: Sequence[Ext], shift: int) -> List[Ext]:
    pw, scaled = 1, []
    for c in coeffs:
        scaled.append(escale(c, pw))
        pw = pw * shift % P
    return ext_fft(scaled)


def reverse_index_bits(v: list) -> list:
    bits = len(v).bit_length() - 1
    return [v[bitrev(i, bits)] for i in range(len(v))]


def flatten(chunk: Sequence[Ext]) -> List[int]:
    return [x for e in chunk for x in e]


def proof_of_work(current_hash: Sequence[int], pow_bits: int) -> int:
    """Smallest w with leading_zeros(hash_no_pad(current_hash || w)[0]) >= pow_bits (serial search)."""
    w = 0
    while True:
        r = int(O.hash_no_pad(list(current_hash) + [w])[0])
        if (64 - r.bit_length()) >= pow_bits:
            return w
        w += 1


def fri_proof(initial_trees, final_coeffs: List[Ext], ch: Challenger, rate_bits: int, cap_height: int, arities: Sequence[int],
              pow_bits: int, num_queries: int) -> dict:
    """initial_trees: list of dicts from O.commit (leaves, digests, cap).  final_coeffs: length n (degree side)."""
    n = len(final_coeffs)
    N = n << rate_bits
    coeffs = list(final_coeffs) + [(0, 0)] * (N - n)               # final_poly.lde(rate_bits)
    values = coset_fft(coeffs, G)
    shift = G
    trees, caps = [], []
    for ab in arities:
        arity = 1 << ab
        values = reverse_index_bits(values)
        leaves = np.array([flatten(values[j * arity:(j + 1) * arity]) for j in range(len(values) // arity)], dtype=np.uint64)
        dig, cap = O.merkle_new(leaves, cap_height)
        trees.append(dict(leaves=leaves, digests=dig, cap=cap))
        ch.observe_cap(cap)
        caps.append(cap)
        beta = ch.get_extension_challenge()
        folded = []
        for j in range(len(coeffs) // arity):                       # reduce_with_powers(chunk, beta)
            acc = (0, 0)
            for c in reversed(coeffs[j * arity:(j + 1) * arity]):
                acc = eadd(emul(acc, beta), c)
            folded.append(acc)
        coeffs = folded
        shift = pow(shift, arity, P)
        values = coset_fft(coeffs, shift)
    assert all(c == (0, 0) for c in coeffs[len(coeffs) >> rate_bits:]), "the removed coefficients should always be zero"
    coeffs = coeffs[:len(coeffs) >> rate_bits]
    ch.observe_extension_elements(coeffs)
    pow_witness = proof_of_work(ch.get_hash(), pow_bits)
    rounds = []
    for _ in range(num_queries):
        x_index = x0 = ch.get_challenge() % N
        initial = []
        for t in initial_trees:
            Nt = t["leaves"].shape[0]
            initial.append((t["leaves"][x_index].copy(), O.merkle_prove(t["digests"], Nt, int(np.log2(t["cap"].shape[0])), x_index)))
        steps = []
        for t, ab in zip(trees, arities):
            x_index >>= ab
            nl = t["leaves"].shape[0]
            steps.append((t["leaves"][x_index].reshape(-1, 2).copy(), O.merkle_prove(t["digests"], nl, cap_height, x_index)))
        rounds.append(dict(x_index=x0, initial=initial, steps=steps))
    return dict(caps=caps, rounds=rounds, final_poly=coeffs, pow_witness=pow_witness)


def prove_openings(oracle_commits, batches, ch: Challenger, rate_bits: int, cap_height: int, arities, pow_bits: int,
                   num_queries: int, mul_by_x: bool = True) -> dict:
    alpha = ch.get_extension_challenge()
    final = final_poly_of([c["coeffs"] for c in oracle_commits], batches, alpha, mul_by_x)
    proof = fri_proof(oracle_commits, final, ch, rate_bits, cap_height, arities, pow_bits, num_queries)
    proof["alpha"] = alpha
    proof["final_poly_before_lde"] = final
    return proof


# ------------------------------------------------------------------ verifier (written from the verifier's side)
def interpolate_at(points: Sequence[Tuple[Ext, Ext]], x: Ext) -> Ext:
    """Plain Lagrange interpolation of {(x_i, y_i)} evaluated at x."""
    acc = (0, 0)
    for i, (xi, yi) in enumerate(points):
        num, den = (1, 0), (1, 0)
        for j, (xj, _) in enumerate(points):
            if i != j:
                num = emul(num, esub(x, xj))
                den = emul(den, esub(xi, xj))
        acc = eadd(acc, emul(yi, emul(num, einv(den))))
    return acc


def compute_evaluation(x: int, x_index_within_coset: int, arity_bits: int, evals: Sequence[Ext], beta: Ext) -> Ext:
    arity = 1 << arity_bits
    g = root(arity_bits)
    ev = reverse_index_bits(list(evals))
    rev = bitrev(x_index_within_coset, arity_bits)
    coset_start = x * pow(g, arity - rev, P) % P
    pts = [((coset_start * pow(g, i, P) % P, 0), ev[i]) for i in range(arity)]
    return interpolate_at(pts, beta)


def verify(proof: dict, initial_caps: Sequence[np.ndarray], batches, openings: Sequence[Sequence[Ext]], ch: Challenger,
           n_log: int, rate_bits: int, cap_height: int, arities, pow_bits: int, num_queries: int, mul_by_x: bool = True) -> bool:
    """verify_fri_proof.  openings[b][j] = claimed value of batch b's j-th polynomial at its point.  `ch` must be in the
    state the prover's challenger had when prove_openings was called."""
    N_log = n_log + rate_bits
    N = 1 << N_log
    alpha = ch.get_extension_challenge()
    betas = []
    for cap in proof["caps"]:
        ch.observe_cap(cap)
        betas.append(ch.get_extension_challenge())
    ch.observe_extension_elements(proof["final_poly"])
    pow_response = int(O.hash_no_pad(ch.get_hash() + [proof["pow_witness"]])[0])
    if 64 - pow_response.bit_length() < pow_bits:
        return False
    if len(proof["rounds"]) != num_queries or len(proof["caps"]) != len(arities):
        return False
    # PrecomputedReducedOpenings::from_os_and_alpha
    reduced_openings = []
    for vals in openings:
        acc = (0, 0)
        for v in reversed(vals):
            acc = eadd(emul(acc, alpha), v)
        reduced_openings.append(acc)
    for rnd in proof["rounds"]:
        x_index = ch.get_challenge() % N
        # fri_verify_initial_proof
        for (row, sib), cap in zip(rnd["initial"], initial_caps):
            if not O.merkle_verify(row, x_index, sib, cap):
                return False
        subgroup_x = G * pow(root(N_log), bitrev(x_index, N_log), P) % P
        # fri_combine_initial
        total = (0, 0)
        for (point, polys), red_open in zip(batches, reduced_openings):
            acc = (0, 0)
            for o, i in reversed(polys):
                acc = eadd(emul(acc, alpha), (int(rnd["initial"][o][0][i]), 0))
            num = esub(acc, red_open)
            den = esub((subgroup_x, 0), point)
            total = emul(total, epow(alpha, len(polys)))
            total = eadd(total, emul(num, einv(den)))
        old_eval = escale(total, subgroup_x) if mul_by_x else total
        for i, ab in enumerate(arities):
            evals = [(int(a), int(b)) for a, b in rnd["steps"][i][0]]
            coset_index = x_index >> ab
            within = x_index & ((1 << ab) - 1)
            if evals[within] != old_eval:
                return False
            old_eval = compute_evaluation(subgroup_x, within, ab, evals, betas[i])
            if not O.merkle_verify(np.array(flatten(evals), dtype=np.uint64), coset_index, rnd["steps"][i][1], proof["caps"][i]):
                return False
            subgroup_x = pow(subgroup_x, 1 << ab, P)
            x_index = coset_index
        if eval_poly(proof["final_poly"], (subgroup_x, 0)) != old_eval:
            return False
    return True
