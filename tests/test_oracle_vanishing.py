"""oracle/vanishing_ref.py (row N1b restatement) against the argument it restates: for a satisfied synthetic circuit the
quotient computed point by point on the coset interpolates to a polynomial that fulfils the verifier's identity at a random
point, and a single wrong wire, sigma value or partial product breaks it.  CPU only."""
import random

import numpy as np
import pytest

from oracle import perm_ref as PR
from oracle import vanishing_ref as V

P = V.P


def build_instance(oracle, n_log, seed, rate_bits=3, degree=8, extended=False):
    c = V.Circuit(n_log, seed=seed, extended=extended)
    rnd = random.Random(seed + 1)
    C = 2
    betas, gammas, alphas = ([rnd.randrange(P) for _ in range(C)] for _ in range(3))
    routed = [c.wires[j] for j in range(V.NUM_ROUTED)]
    zs_pp = PR.partial_products_and_zs(routed, c.sigmas, c.k_is, betas, gammas, degree)
    values = dict(cs=c.constants + c.sigmas, wires=c.wires, zpp=zs_pp)
    coeffs = {k: [oracle.ifft(np.array(col, np.uint64)) for col in cols] for k, cols in values.items()}
    ldes = {k: [oracle.coset_lde(col, rate_bits) for col in cols] for k, cols in coeffs.items()}
    return c, betas, gammas, alphas, coeffs, ldes


def coset_ifft(oracle, vals):
    co = oracle.ifft(np.array(vals, np.uint64))
    inv7 = pow(7, P - 2, P)
    s, out = 1, []
    for a in co:
        out.append(int(a) * s % P)
        s = s * inv7 % P
    return out


def test_poseidon_gate_witness_satisfies_constraints(oracle):
    rnd = random.Random(5)
    for swap in (0, 1):
        inp = [rnd.randrange(P) for _ in range(12)]
        w = V.poseidon_gate_witness(inp, swap)
        out = []
        V.poseidon_gate_constraints(w, out)
        assert len(out) == 123 and not any(out)
        # the gate computes the reference's permutation (with the first two 4-word halves swapped when swap = 1)
        st = inp[4:8] + inp[:4] + inp[8:] if swap else inp
        assert [int(x) for x in oracle.permute(np.array(st, np.uint64))] == w[V.W_OUT:V.W_OUT + 12]
        w[V.W_PARTIAL + 7] = (w[V.W_PARTIAL + 7] + 1) % P
        out = []
        V.poseidon_gate_constraints(w, out)
        assert any(out)


def test_selector_groups():
    sel, groups = V.selector_groups([V.NOOP, V.CONSTANT, V.PUBLIC_INPUT, V.ARITHMETIC, V.POSEIDON], 8)
    assert groups == [(0, 4), (4, 5)] and sel == [0, 0, 0, 0, 1]
    assert V.selector_groups([V.NOOP, V.CONSTANT], 8) == ([0, 0], [(0, 2)])
    # a filter vanishes on every other gate of its group and on rows that belong to another group
    for row in range(4):
        for other in list(range(4)) + [V.UNUSED_SELECTOR]:
            f = V.compute_filter(row, (0, 4), other, True)
            assert (f != 0) == (other == row)


@pytest.mark.parametrize("n_log,seed", [(3, 1), (4, 2)])
def test_quotient_satisfies_the_verifier_identity(oracle, n_log, seed):
    c, betas, gammas, alphas, coeffs, ldes = build_instance(oracle, n_log, seed)
    # the permutation argument closes for this witness
    z = [int(v) for v in coeffs["zpp"][0]]
    q = V.quotient_values(c, ldes["cs"], ldes["wires"], ldes["zpp"], betas, gammas, alphas, 8, 3, 3)
    qc = [coset_ifft(oracle, col) for col in q]
    rnd = random.Random(99)
    for _ in range(2):
        zeta = rnd.randrange(2, P)
        assert V.check_quotient_identity(c, coeffs["cs"], coeffs["wires"], coeffs["zpp"], qc, betas, gammas, alphas, 8, zeta)
    # quotient_degree_bits < rate_bits reads every 2nd LDE row; the smaller coset cannot hold the degree-8n quotient of this
    # circuit, so only its values are compared: they are the same function on the sub-coset
    q2 = V.quotient_values(c, ldes["cs"], ldes["wires"], ldes["zpp"], betas, gammas, alphas, 8, 3, 2)
    assert all(q2[k][i] == q[k][2 * i] for k in range(2) for i in range(len(q2[0])))
    # one wrong wire: the "quotient" no longer satisfies the identity
    bad = [list(col) for col in c.wires]
    bad[V.W_PARTIAL + 3][0] = (bad[V.W_PARTIAL + 3][0] + 1) % P
    bad_coeffs = [oracle.ifft(np.array(col, np.uint64)) for col in bad]
    bad_lde = [oracle.coset_lde(col, 3) for col in bad_coeffs]
    qb = V.quotient_values(c, ldes["cs"], bad_lde, ldes["zpp"], betas, gammas, alphas, 8, 3, 3)
    qbc = [coset_ifft(oracle, col) for col in qb]
    assert not V.check_quotient_identity(c, coeffs["cs"], bad_coeffs, coeffs["zpp"], qbc, betas, gammas, alphas, 8, 12345)


def test_other_gate_witnesses_satisfy_their_constraints():
    """ArithmeticExtension, MulExtension, BaseSum, Reducing, ReducingExtension, RandomAccess, Exponentiation, PoseidonMds: the
    generated row satisfies eval_unfiltered, the constraint count is Gate::num_constraints, and a changed wire breaks it"""
    rnd = random.Random(7)
    for g in (V.ARITHMETIC_EXTENSION, V.MUL_EXTENSION, V.BASE_SUM, V.REDUCING, V.REDUCING_EXTENSION, V.RANDOM_ACCESS,
              V.EXPONENTIATION, V.POSEIDON_MDS):
        pr = V.EXT_GATE_PARAMS.get(g, ())
        cc = [rnd.randrange(P) for _ in range(2)]
        w = V.gate_witness(g, rnd, cc, pr)
        out = V.gate_constraints(g, w, cc, [0] * 4, pr)
        assert len(out) == V.gate_num_constraints(g, pr) and not any(out)
        w[0] = (w[0] + 1) % P
        assert any(V.gate_constraints(g, w, cc, [0] * 4, pr))
    # known values: 3^0b101 = 243 through the exponentiation gate, a base-2 sum, one quadratic-extension product
    w = [0] * V.NUM_WIRES
    w[0], w[1], w[3] = 3, 1, 1
    w[2 + 3:2 + 6] = [3, 9, 243]
    w[1 + 3] = 243
    assert not any(V.gate_constraints(V.EXPONENTIATION, w, [0, 0], [0] * 4, (3,)))
    assert V.ext_mul((2, 3), (5, 11)) == (2 * 5 + 7 * 3 * 11, 2 * 11 + 3 * 5)


def test_quotient_identity_with_every_gate_kind(oracle):
    """the same identity for a circuit that also holds rows of the other eight gate kinds (13 gates, 4 selector polynomials)"""
    c, betas, gammas, alphas, coeffs, ldes = build_instance(oracle, 5, 11, extended=True)
    assert c.num_selectors == 4 and len(c.gates) == 13
    q = V.quotient_values(c, ldes["cs"], ldes["wires"], ldes["zpp"], betas, gammas, alphas, 8, 3, 3)
    qc = [coset_ifft(oracle, col) for col in q]
    assert V.check_quotient_identity(c, coeffs["cs"], coeffs["wires"], coeffs["zpp"], qc, betas, gammas, alphas, 8, 987654321)
    assert V.check_quotient_identity(c, coeffs["cs"], coeffs["wires"], coeffs["zpp"], qc, betas, gammas, alphas, 8, 31337)
    # one wrong wire on a ReducingGate row
    row = c.row_gate.index(c.gates.index(V.REDUCING))
    bad = [list(col) for col in c.wires]
    bad[9][row] = (bad[9][row] + 1) % P
    bad_coeffs = [oracle.ifft(np.array(col, np.uint64)) for col in bad]
    bad_lde = [oracle.coset_lde(col, 3) for col in bad_coeffs]
    qb = V.quotient_values(c, ldes["cs"], bad_lde, ldes["zpp"], betas, gammas, alphas, 8, 3, 3)
    qbc = [coset_ifft(oracle, col) for col in qb]
    assert not V.check_quotient_identity(c, coeffs["cs"], bad_coeffs, coeffs["zpp"], qbc, betas, gammas, alphas, 8, 12345)
