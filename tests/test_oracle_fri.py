"""CPU: the opening-proof restatement (oracle/fri_ref.py) is self-consistent — its prover's proofs pass its independently
written verifier, tampered proofs do not.  PARITY UNPINNED (no proof fixture in the reference); see the module header."""
import copy

import numpy as np
import pytest

from helpers import P, rand_field


def make_case(O, F, rng, n_log, ks, rate_bits, cap_height):
    n = 1 << n_log
    commits = [O.commit(rand_field(rng, (k, n)), rate_bits, cap_height, is_coeffs=True) for k in ks]
    zeta = (int(rng.integers(1, P, dtype=np.uint64)), int(rng.integers(1, P, dtype=np.uint64)))
    g = F.root(n_log)
    gzeta = F.escale(zeta, g)
    all_polys = [(o, i) for o, k in enumerate(ks) for i in range(k)]
    zs = [(len(ks) - 1, i) for i in range(min(2, ks[-1]))]
    batches = [(zeta, all_polys), (gzeta, zs)]
    openings = [[tuple(int(v) for v in O.eval_ext2(commits[o]["coeffs"][i], point)) for o, i in polys] for point, polys in batches]
    return commits, batches, openings


def seeded_challenger(F, commits):
    ch = F.Challenger()
    for c in commits:
        ch.observe_cap(c["cap"])
    return ch


@pytest.mark.parametrize("mul_by_x", [True, False])
@pytest.mark.parametrize("n_log,ks,rate_bits,cap_height,arities", [
    (5, (3, 2), 2, 1, (2, 1)),
    (4, (1, 4, 2), 3, 0, (1, 1, 1)),
    (6, (5,), 1, 2, (3,)),
    (3, (2, 2), 3, 2, ()),
])
def test_prover_restatement_passes_the_verifier(oracle, n_log, ks, rate_bits, cap_height, arities, mul_by_x):
    from oracle import fri_ref as F
    O = oracle
    rng = np.random.default_rng(1000 + n_log)
    commits, batches, openings = make_case(O, F, rng, n_log, ks, rate_bits, cap_height)
    pow_bits, nq = 5, 6
    proof = F.prove_openings(commits, batches, seeded_challenger(F, commits), rate_bits, cap_height, arities, pow_bits, nq, mul_by_x)
    caps = [c["cap"] for c in commits]
    args = (n_log, rate_bits, cap_height, arities, pow_bits, nq, mul_by_x)
    assert F.verify(proof, caps, batches, openings, seeded_challenger(F, commits), *args)
    assert len(proof["final_poly"]) == (1 << n_log) >> sum(arities)

    # the other convention must not verify (the two differ by a factor X on every evaluation)
    assert not F.verify(proof, caps, batches, openings, seeded_challenger(F, commits), *args[:-1], not mul_by_x)

    # a wrong opening, a flipped final coefficient, a wrong witness, a swapped query answer: all rejected
    bad = copy.deepcopy(openings)
    bad[0][0] = ((bad[0][0][0] + 1) % P, bad[0][0][1])
    assert not F.verify(proof, caps, batches, bad, seeded_challenger(F, commits), *args)
    p2 = copy.deepcopy(proof)
    p2["final_poly"][0] = ((p2["final_poly"][0][0] + 1) % P, p2["final_poly"][0][1])
    assert not F.verify(p2, caps, batches, openings, seeded_challenger(F, commits), *args)
    p3 = copy.deepcopy(proof)
    p3["pow_witness"] += 1
    assert not F.verify(p3, caps, batches, openings, seeded_challenger(F, commits), *args)
    if arities:
        p4 = copy.deepcopy(proof)
        p4["rounds"][0]["steps"][0][0][0, 0] ^= np.uint64(1)
        assert not F.verify(p4, caps, batches, openings, seeded_challenger(F, commits), *args)


def test_quotient_is_exact(oracle):
    """(F(X) - F(z)) / (X - z) * (X - z) + F(z) = F(X): divide_by_linear against plain multiplication."""
    from oracle import fri_ref as F
    rng = np.random.default_rng(7)
    coeffs = [(int(a), int(b)) for a, b in rand_field(rng, (33, 2))]
    z = (int(rng.integers(1, P, dtype=np.uint64)), int(rng.integers(1, P, dtype=np.uint64)))
    q = F.divide_by_linear(coeffs, z)
    assert len(q) == len(coeffs) - 1
    back = [(0, 0)] * len(coeffs)
    for i, c in enumerate(q):                      # q(X) * (X - z)
        back[i + 1] = F.eadd(back[i + 1], c)
        back[i] = F.esub(back[i], F.emul(c, z))
    back[0] = F.eadd(back[0], F.eval_poly(coeffs, z))
    assert back == coeffs


def test_pow_witness_is_minimal(oracle):
    from oracle import fri_ref as F
    h = [1, 2, 3, 4]
    w = F.proof_of_work(h, 6)
    lz = lambda x: 64 - int(oracle.hash_no_pad(h + [x])[0]).bit_length()
    assert lz(w) >= 6 and all(lz(x) < 6 for x in range(w))


def test_challenger_matches_sponge_definition(oracle):
    """8 observed elements then one challenge = last rate word of one permutation of the overwritten zero state."""
    from oracle import fri_ref as F
    ch = F.Challenger()
    ch.observe_elements(list(range(1, 9)))
    st = np.zeros(12, np.uint64); st[:8] = np.arange(1, 9, dtype=np.uint64)
    out = oracle.permute(st)
    assert ch.get_challenge() == int(out[7]) and ch.get_challenge() == int(out[6])
    ch.observe_element(5)                                   # partial buffer: overwrite word 0 only, then permute
    st2 = out.copy(); st2[0] = 5
    assert ch.get_challenge() == int(oracle.permute(st2)[7])


def golden_case_inputs(O, F, case):
    n = 1 << case["n_log"]
    coeffs = [O.synthetic_values(k, n, seed=o + 1) for o, k in enumerate(case["ks"])]
    zeta = tuple(case["zeta"])
    ks = case["ks"]
    batches = [(zeta, [(o, i) for o, k in enumerate(ks) for i in range(k)]),
               (F.escale(zeta, F.root(case["n_log"])), [(len(ks) - 1, 0)])]
    return coeffs, batches


def test_golden_opening_proof_vectors(oracle, golden):
    """The committed vectors (tests/golden/make_golden.py) freeze the restatement: regenerate and compare."""
    from oracle import fri_ref as F
    sys_path_golden = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden")
    __import__("sys").path.insert(0, sys_path_golden)
    import make_golden as G
    for case in golden["opening_proof_vectors"]["cases"]:
        coeffs, batches = golden_case_inputs(oracle, F, case)
        r, h = case["rate_bits"], case["cap_height"]
        commits = [oracle.commit(c, r, h, is_coeffs=True) for c in coeffs]
        ch = F.Challenger()
        for c in commits:
            ch.observe_cap(c["cap"])
        proof = F.prove_openings(commits, batches, ch, r, h, case["arities"], case["pow_bits"], case["num_queries"], case["mul_by_x"])
        assert [[int(a), int(b)] for a, b in proof["final_poly"]] == case["final_poly"]
        assert proof["pow_witness"] == case["pow_witness"]
        assert [rnd["x_index"] for rnd in proof["rounds"]] == case["x_indices"]
        assert "%016x" % G.fnv(G.proof_words(proof)) == case["proof_fnv1a64"]
