"""Oracle self-consistency on the part of the path the reference does NOT pin (NTT / LDE / Merkle cap):
definitional DFT vs radix-2, Horner spot checks, algebraic properties, the pure-Python twin, the SURVEY.md
App. C vectors, the committed golden vectors and the multithreaded CPU baseline.  CPU only."""
import numpy as np
import pytest

from helpers import P, bitrev, bitrev_perm, hostile_columns, rand_field


def test_fft_matches_definition(oracle):
    rng = np.random.default_rng(0)
    for n_log in range(0, 9):
        v = rand_field(rng, 1 << n_log)
        assert (oracle.fft(v) == oracle.dft(v)).all()


def test_ifft_inverts_fft_and_noncanonical_inputs(oracle):
    rng = np.random.default_rng(1)
    for n_log in (0, 1, 5, 10):
        v = rand_field(rng, 1 << n_log)
        assert (oracle.ifft(oracle.fft(v)) == v).all()
        assert (oracle.fft(oracle.ifft(v)) == v).all()
    nc = np.array([2**64 - 1, P, P + 5, 0], dtype=np.uint64)
    assert (oracle.fft(nc) == oracle.fft(nc % np.uint64(P))).all()


def test_lde_is_evaluation_on_coset(oracle):
    rng = np.random.default_rng(2)
    for n_log, r in ((3, 3), (5, 1), (4, 0), (6, 3)):
        c = rand_field(rng, 1 << n_log)
        lde = oracle.coset_lde(c, r)
        for i in (0, 1, (1 << (n_log + r)) - 1, 5 % (1 << (n_log + r))):
            assert int(lde[i]) == oracle.eval_at_lde_point(c, r, i)


def test_lde_rows_are_coset_ntts(oracle):
    """natural rows i = t (mod 2^r) are the size-n transform of coeff_j * (7 w_N^t)^j."""
    rng = np.random.default_rng(3)
    n_log, r = 5, 3
    n = 1 << n_log
    c = rand_field(rng, n)
    lde = oracle.coset_lde(c, r)
    g2 = 1753635133440165772
    wN = pow(g2, 1 << (32 - n_log - r), P)
    for t in range(1 << r):
        s = 7 * pow(wN, t, P) % P
        scaled = np.array([int(c[j]) * pow(s, j, P) % P for j in range(n)], dtype=np.uint64)
        assert (oracle.fft(scaled) == lde[t::1 << r]).all()


def test_commit_linearity(oracle):
    rng = np.random.default_rng(4)
    a, b = rand_field(rng, (3, 16)), rand_field(rng, (3, 16))
    s = ((a.astype(object) + b.astype(object)) % P).astype(np.uint64)
    ra, rb, rs = (oracle.commit(x, 2, 1) for x in (a, b, s))
    assert (((ra["leaves"].astype(object) + rb["leaves"].astype(object)) % P).astype(np.uint64) == rs["leaves"]).all()
    assert (((ra["coeffs"].astype(object) + rb["coeffs"].astype(object)) % P).astype(np.uint64) == rs["coeffs"]).all()


def test_survey_appendix_c(oracle):
    v = oracle.synthetic_values(9, 8)
    assert [int(x) for x in v[0][:3]] == [16294208416658607535, 7960286522194355700, 487617019471545679]
    assert int(v[8][7]) == 15781199770154346095
    r = oracle.commit(v, 3, 2)
    assert [int(x) for x in r["coeffs"][0][:3]] == [8511423253799370256, 3699637470157816969, 12653192862501000485]
    assert [int(x) for x in r["leaves"][0][:3]] == [1132379675625856675, 10298623730004040521, 13308988100810897799]
    assert [int(x) for x in r["leaves"][1][:3]] == [4462446298076148731, 10303375107340728390, 16635710509762984575]
    assert int(r["leaves"][63][8]) == 8562002592535086970
    assert [[int(x) for x in row] for row in r["cap"]] == [
        [9531016979423918488, 17086599980695735262, 12854109491395286945, 1436292215001049984],
        [7836116892235451435, 1036240376415213189, 7578210113274565172, 15446728671533160986],
        [17059806711847136650, 17584175406166898310, 11149053496755269788, 12423130825012369409],
        [7263687849510528849, 3413957559828268007, 13767290799342735489, 13701751426046424270]]
    r2 = oracle.commit(oracle.synthetic_values(3, 4), 1, 0)     # leaves <= 4 elements: not hashed
    assert [int(x) for x in r2["leaves"][0]] == [13501206060863318524, 10258701369751965010, 2389896691979806751]
    assert [int(x) for x in r2["cap"][0]] == [17542564077989784265, 408585749171979391, 12611990736760798239, 1539051007847932482]


def test_golden_commit_vectors(oracle, golden):
    for case in golden["commit_vectors"]["cases"]:
        n = 1 << case["n_log"]
        v = oracle.synthetic_values(case["k"], n)
        salt = oracle.synthetic_values(4, n << case["rate_bits"], seed=case["salt_seed"]) if case["salt_seed"] else None
        r = oracle.commit(v, case["rate_bits"], case["cap_height"], is_coeffs=case["is_coeffs"], salt=salt)
        assert [[int(x) for x in row] for row in r["cap"]] == case["cap"]
        assert [int(x) for x in r["coeffs"][0][:3]] == case["coeffs_col0_head"]
        assert [int(x) for x in r["leaves"][0][:3]] == case["leaf0_head"]


@pytest.mark.parametrize("n_log,k,r,h,salted", [(2, 3, 1, 0, False), (3, 9, 2, 2, False), (2, 5, 2, 4, False),
                                               (0, 6, 2, 1, False), (3, 2, 1, 2, True)])
def test_python_twin_agrees(oracle, n_log, k, r, h, salted):
    from oracle import pyref
    v = oracle.synthetic_values(k, 1 << n_log, seed=3)
    salt = oracle.synthetic_values(4, 1 << (n_log + r), seed=9) if salted else None
    a = oracle.commit(v, r, h, salt=salt)
    b = pyref.commit([[int(x) for x in c] for c in v], r, h,
                     salt=None if salt is None else [[int(x) for x in s] for s in salt])
    assert [[int(x) for x in c] for c in a["coeffs"]] == b["coeffs"]
    assert [[int(x) for x in row] for row in a["leaves"]] == b["leaves"]
    assert [[int(x) for x in row] for row in a["cap"]] == b["cap"]
    assert [[int(x) for x in row] for row in a["digests"]] == b["digests"]     # recursive fill == index formula


def test_merkle_prove_verify_every_leaf(oracle):
    rng = np.random.default_rng(5)
    for (N, L, h) in ((16, 7, 0), (16, 7, 2), (8, 3, 1), (4, 9, 2), (32, 13, 5)):
        leaves = rand_field(rng, (N, L))
        dig, cap = oracle.merkle_new(leaves, h)
        assert dig.shape[0] == 2 * (N - (1 << h))
        for i in range(N):
            sib = oracle.merkle_prove(dig, N, h, i)
            assert sib.shape[0] == (N.bit_length() - 1) - h
            assert oracle.merkle_verify(leaves[i], i, sib, cap)
            if sib.shape[0]:
                bad = sib.copy(); bad[0, 0] ^= np.uint64(1)
                assert not oracle.merkle_verify(leaves[i], i, bad, cap)


def test_merkle_bad_arguments(oracle):
    with pytest.raises(ValueError):
        oracle.merkle_new(np.zeros((8, 5), np.uint64), 4)      # cap_height > log2(N)
    with pytest.raises(ValueError):
        oracle.merkle_new(np.zeros((6, 5), np.uint64), 1)      # not a power of two


def test_hostile_inputs_commit(oracle):
    v = hostile_columns(8)
    r = oracle.commit(v, 3, 2)
    canon = np.where(v >= np.uint64(P), v - np.uint64(P), v)
    r2 = oracle.commit(canon, 3, 2)
    for key in ("coeffs", "leaves", "cap", "digests"):
        assert (r[key] == r2[key]).all()
    assert (r["leaves"][:, 0] == 0).all() and (r["leaves"][:, 4] == 0).all()     # zero columns stay zero


@pytest.mark.parametrize("n_log,k,r,h", [(6, 7, 3, 4), (10, 20, 3, 4), (4, 135, 3, 4), (3, 3, 1, 0), (5, 16, 3, 8)])
def test_cpu_baseline_matches_oracle(oracle, n_log, k, r, h):
    v = oracle.synthetic_values(k, 1 << n_log, seed=1)
    a = oracle.commit(v, r, h)
    b, times = oracle.baseline_commit(v, r, h)
    for key in ("coeffs", "leaves", "digests", "cap"):
        assert (a[key] == b[key]).all(), key
    assert len(times) == 5 and times[4] >= 0
    c, _ = oracle.baseline_commit(a["coeffs"], r, h, is_coeffs=True)
    assert (c["cap"] == a["cap"]).all()


def test_eval_ext2_matches_definition(oracle):
    """Quadratic-extension evaluation (row N3) against Python integers: X^2 = 7."""
    rng = np.random.default_rng(8)
    for n in (1, 2, 7, 64):
        c = [int(x) % P for x in rng.integers(0, 2**64, size=n, dtype=np.uint64)]
        za, zb = (int(x) % P for x in rng.integers(0, 2**64, size=2, dtype=np.uint64))
        ra, rb, pa, pb = 0, 0, 1, 0          # result, running power of zeta
        for ci in c:
            ra, rb = (ra + ci * pa) % P, (rb + ci * pb) % P
            pa, pb = (pa * za + 7 * pb * zb) % P, (pa * zb + pb * za) % P
        got = oracle.eval_ext2(np.array(c, dtype=np.uint64), np.array([za, zb], dtype=np.uint64))
        assert [int(x) for x in got] == [ra, rb]
    # a base-field point gives the base-field value: compare with the LDE (zeta = 7 * w_N^i)
    c = rand_field(rng, 32)
    lde = oracle.coset_lde(c, 2)
    g2 = 1753635133440165772
    x = 7 * pow(pow(g2, 1 << (32 - 7), P), 5, P) % P
    got = oracle.eval_ext2(c, np.array([x, 0], dtype=np.uint64))
    assert int(got[0]) == int(lde[5]) and int(got[1]) == 0
