"""oracle/perm_ref.py (row N1a, permutation argument): the restatement closes on satisfied copy constraints, opens on a
wrong sigma, and its columns satisfy the recurrences plonky2's verifier checks."""
import random

from oracle import perm_ref as PR

P = PR.P


def test_coset_shifts_and_counts():
    assert PR.coset_shifts(4) == [1, 7, 49, 343]
    assert PR.num_partial_products(80, 8) == 9 and PR.num_partial_products(12, 8) == 1 and PR.num_partial_products(8, 8) == 0
    assert pow(PR.root(5), 32, P) == 1 and pow(PR.root(5), 16, P) == P - 1


def test_argument_closes_on_a_satisfied_permutation():
    rnd = random.Random(2)
    for R, degree, n_log in ((12, 8, 4), (7, 3, 3), (20, 8, 5)):
        wires, sigmas, k_is = PR.valid_permutation_instance(R, n_log, seed=R)
        betas = [rnd.randrange(P) for _ in range(2)]
        gammas = [rnd.randrange(P) for _ in range(2)]
        cols = PR.partial_products_and_zs(wires, sigmas, k_is, betas, gammas, degree)
        chunks = -(-R // degree)
        assert len(cols) == 2 * chunks and all(len(c) == 1 << n_log for c in cols)
        assert PR.check_recurrences(cols, wires, sigmas, k_is, betas, gammas, degree)
        for c in range(2):
            assert PR.final_product(cols, wires, sigmas, k_is, betas[c], gammas[c], degree, c) == 1


def test_argument_opens_on_a_wrong_sigma_and_detects_a_tampered_column():
    R, degree, n_log = 12, 8, 4
    wires, sigmas, k_is = PR.valid_permutation_instance(R, n_log, seed=5)
    betas, gammas = [123456789], [987654321]
    sigmas[2][3] = (sigmas[2][3] + 1) % P
    cols = PR.partial_products_and_zs(wires, sigmas, k_is, betas, gammas, degree)
    assert PR.check_recurrences(cols, wires, sigmas, k_is, betas, gammas, degree)      # still the honest prover's columns
    assert PR.final_product(cols, wires, sigmas, k_is, betas[0], gammas[0], degree) != 1
    cols[1][5] = (cols[1][5] + 1) % P
    assert not PR.check_recurrences(cols, wires, sigmas, k_is, betas, gammas, degree)


def test_identity_permutation_gives_all_ones():
    R, degree, n_log = 9, 4, 3
    n = 1 << n_log
    k_is = PR.coset_shifts(R)
    w = PR.root(n_log)
    rnd = random.Random(8)
    wires = [[rnd.randrange(P) for _ in range(n)] for _ in range(R)]
    sigmas = [[k_is[j] * pow(w, i, P) % P for i in range(n)] for j in range(R)]
    cols = PR.partial_products_and_zs(wires, sigmas, k_is, [5, 6], [7, 8], degree)
    assert all(v == 1 for c in cols for v in c)
