"""GPU parity for row N1a (permutation argument: Z and partial products) against oracle/perm_ref.py, through the C ABI."""
import random

import numpy as np
import pytest

from oracle import perm_ref as PR

pytestmark = pytest.mark.gpu
P = PR.P


@pytest.fixture(scope="module")
def ctx():
    import intmax_zkp_core_b200 as z
    c = z.Context(0)
    yield c
    c.close()


def _instance(R, n_log, Cn, seed):
    rnd = random.Random(seed)
    wires, sigmas, k_is = PR.valid_permutation_instance(R, n_log, seed=seed)
    betas = [rnd.randrange(P) for _ in range(Cn)]
    gammas = [rnd.randrange(P) for _ in range(Cn)]
    return wires, sigmas, k_is, betas, gammas


@pytest.mark.parametrize("R,degree,n_log,Cn", [(80, 8, 6, 2), (12, 8, 4, 2), (7, 3, 5, 1), (5, 8, 0, 3), (80, 8, 11, 2), (33, 4, 7, 2)])
def test_zs_partial_products_match_restatement(ctx, R, degree, n_log, Cn):
    from intmax_zkp_core_b200 import prover as Z
    wires, sigmas, k_is, betas, gammas = _instance(R, n_log, Cn, seed=R * 100 + n_log)
    got = Z.zs_partial_products(np.array(wires, np.uint64), np.array(sigmas, np.uint64), np.array(k_is, np.uint64), betas, gammas,
                                degree, ctx)
    ref = PR.partial_products_and_zs(wires, sigmas, k_is, betas, gammas, degree)
    assert got.shape == (len(ref), 1 << n_log)
    assert (got == np.array(ref, np.uint64)).all()
    assert (Z.get_unique_coset_shifts(R) == np.array(k_is, np.uint64)).all()
    assert Z.num_partial_products(R, degree) == PR.num_partial_products(R, degree)


def test_argument_closes_and_detects_a_broken_copy_constraint(ctx):
    """a satisfied permutation brings the running product back to 1 after the last row; a wrong sigma does not"""
    from intmax_zkp_core_b200 import prover as Z
    R, degree, n_log = 80, 8, 8
    wires, sigmas, k_is, betas, gammas = _instance(R, n_log, 2, seed=7)
    got = Z.zs_partial_products(np.array(wires, np.uint64), np.array(sigmas, np.uint64), np.array(k_is, np.uint64), betas, gammas,
                                degree, ctx)
    cols = [[int(v) for v in row] for row in got]
    assert PR.check_recurrences(cols, wires, sigmas, k_is, betas, gammas, degree)
    for c in range(2):
        assert PR.final_product(cols, wires, sigmas, k_is, betas[c], gammas[c], degree, c) == 1
    sigmas[3][5] = (sigmas[3][5] + 1) % P
    bad = Z.zs_partial_products(np.array(wires, np.uint64), np.array(sigmas, np.uint64), np.array(k_is, np.uint64), betas, gammas,
                                degree, ctx)
    cols = [[int(v) for v in row] for row in bad]
    assert PR.final_product(cols, wires, sigmas, k_is, betas[0], gammas[0], degree, 0) != 1


def test_non_canonical_inputs_and_plonky2_row_order(ctx):
    from intmax_zkp_core_b200 import prover as Z
    R, degree, n_log = 16, 8, 5
    wires, sigmas, k_is, betas, gammas = _instance(R, n_log, 2, seed=11)
    w = np.array(wires, np.uint64)
    s = np.array(sigmas, np.uint64)
    ref = np.array(PR.partial_products_and_zs(wires, sigmas, k_is, betas, gammas, degree), np.uint64)
    small = w < np.uint64(2**32 - 1)
    w_nc = np.where(small, w + np.uint64(P), w)          # same residues, not canonical
    got = Z.zs_partial_products(w_nc, s, np.array(k_is, np.uint64), betas, gammas, degree, ctx)
    assert (got == ref).all()
    per = Z.all_wires_permutation_partial_products(w, s, np.array(k_is, np.uint64), betas, gammas, degree, ctx)
    num_prods = Z.num_partial_products(R, degree)
    for c in range(2):
        assert (per[c][-1] == ref[c]).all()                                   # Z last, as prover.rs returns it
        assert (per[c][:-1] == ref[2 + c * num_prods:2 + (c + 1) * num_prods]).all()
    one = Z.wires_permutation_partial_products_and_zs(w, s, np.array(k_is, np.uint64), betas[1], gammas[1], degree, ctx)
    assert (one == per[1]).all()


def test_device_resident_feeds_the_commitment(ctx):
    """wires in HBM -> Z batch in HBM -> PolynomialBatch::from_values on the device, equal to the host path"""
    import torch
    from intmax_zkp_core_b200 import device as D, prover as Z
    from oracle import oracle as O
    R, degree, n_log = 80, 8, 9
    wires, sigmas, k_is, betas, gammas = _instance(R, n_log, 2, seed=5)
    tctx = D.torch_context(0)
    witness = torch.zeros((135, 1 << n_log), dtype=torch.int64, device="cuda")
    witness[:R] = torch.from_numpy(np.array(wires, np.uint64).view(np.int64)).cuda()
    sg = torch.from_numpy(np.array(sigmas, np.uint64).view(np.int64)).cuda()
    batch = Z.zs_partial_products_device(tctx, witness[:R], sg, np.array(k_is, np.uint64), betas, gammas, degree)
    ref = np.array(PR.partial_products_and_zs(wires, sigmas, k_is, betas, gammas, degree), np.uint64)
    assert (batch.cpu().numpy().view(np.uint64) == ref).all()
    com = D.commit_device(tctx, batch, 3, 4)
    torch.cuda.synchronize()
    assert (com.cap.cpu().numpy().view(np.uint64) == O.commit(ref, 3, 4)["cap"]).all()


def test_bad_arguments(ctx):
    from intmax_zkp_core_b200 import prover as Z
    from intmax_zkp_core_b200._lib import B200ZkpError
    w = np.ones((4, 8), np.uint64)
    with pytest.raises(ValueError):
        Z.zs_partial_products(w, np.ones((4, 4), np.uint64), np.ones(4, np.uint64), [1], [1], 8, ctx)
    with pytest.raises(B200ZkpError):
        Z.zs_partial_products(np.ones((40, 8), np.uint64), np.ones((40, 8), np.uint64), np.ones(40, np.uint64), [1], [1], 1, ctx)


def test_coset_ifft_inverts_the_coset_transform(ctx):
    """PolynomialValues::coset_ifft: against the oracle's coset transform (shift 7) and the definition (other shifts)"""
    import intmax_zkp_core_b200 as z
    from oracle import oracle as O
    for n_log, k in ((0, 2), (1, 3), (5, 4), (11, 3), (13, 2)):
        c = O.synthetic_values(k, 1 << n_log, seed=n_log)
        vals = np.stack([O.coset_lde(row, 0) for row in c])   # p(7 w^i), natural order, one polynomial per row
        assert (z.coset_ifft_batch(vals, 7, ctx) == c).all()
    rnd = random.Random(4)
    n_log, shift = 4, 0xDEADBEEF12345
    n = 1 << n_log
    coeffs = [rnd.randrange(P) for _ in range(n)]
    w = PR.root(n_log)
    vals = []
    for i in range(n):
        x = shift * pow(w, i, P) % P
        acc = 0
        for cf in reversed(coeffs):
            acc = (acc * x + cf) % P
        vals.append(acc)
    got = z.PolynomialValues(np.array(vals, np.uint64)).coset_ifft(shift, ctx)
    assert [int(v) for v in got.coeffs] == coeffs
    # a non-canonical shift is reduced first: 5 + p names the coset of 5
    small = np.array([[3, 1, 4, 1, 5, 9, 2, 6]], np.uint64)
    assert (z.coset_ifft_batch(small, 5 + P, ctx) == z.coset_ifft_batch(small, 5, ctx)).all()
    from intmax_zkp_core_b200._lib import B200ZkpError
    with pytest.raises(B200ZkpError):
        z.coset_ifft_batch(np.ones((1, 4), np.uint64), P, ctx)


def test_quotient_chunks_commitment(ctx):
    """row N1c: quotient values on the 8n-point coset -> coset_ifft -> 8 chunks per challenge -> from_coeffs, host and device"""
    import torch
    import intmax_zkp_core_b200 as z
    from intmax_zkp_core_b200 import device as D, prover as Z
    from oracle import oracle as O
    degree_bits, q_bits, Cn = 7, 3, 2
    n = 1 << degree_bits
    qpoly = O.synthetic_values(Cn, n << q_bits, seed=9)       # the quotient polynomials' coefficients
    qvals = np.stack([O.coset_lde(row, 0) for row in qpoly])  # their values on 7 <w_8n>
    chunks = Z.quotient_poly_chunks(qvals, degree_bits, ctx=ctx)
    assert chunks.shape == (Cn << q_bits, n)
    assert (chunks == qpoly.reshape(Cn << q_bits, n)).all()
    batch = Z.commit_quotient(qvals, degree_bits, 3, False, 4, ctx=ctx)
    ref = O.commit(qpoly.reshape(Cn << q_bits, n), 3, 4, is_coeffs=True)
    assert (batch.merkle_tree.cap.elements == ref["cap"]).all()
    tctx = D.torch_context(0)
    dv = torch.from_numpy(qvals.view(np.int64)).cuda()
    com = Z.commit_quotient_device(tctx, dv, degree_bits, 3, 4)
    torch.cuda.synchronize()
    assert (com.cap.cpu().numpy().view(np.uint64) == ref["cap"]).all()
    assert (com.coeffs.cpu().numpy().view(np.uint64) == qpoly.reshape(Cn << q_bits, n)).all()
    # quotient_degree_factor that is not a power of two (plonky2 picks it in min..=max): 6 chunks per challenge are committed and
    # a quotient with non-zero coefficients beyond 6n is refused like plonky2's trim_to_len panic
    with pytest.raises(Z.QuotientError):
        Z.quotient_poly_chunks(qvals, degree_bits, 6, ctx)
    with pytest.raises(Z.QuotientError):
        Z.commit_quotient_device(tctx, dv, degree_bits, 3, 4, quotient_degree_factor=6)
    qpoly6 = qpoly.copy()
    qpoly6[:, 6 * n:] = 0
    qvals6 = np.stack([O.coset_lde(row, 0) for row in qpoly6])
    chunks6 = Z.quotient_poly_chunks(qvals6, degree_bits, 6, ctx)
    assert chunks6.shape == (Cn * 6, n) and (chunks6 == qpoly6[:, :6 * n].reshape(Cn * 6, n)).all()
    ref6 = O.commit(chunks6, 3, 4, is_coeffs=True)
    assert (Z.commit_quotient(qvals6, degree_bits, 3, False, 4, ctx=ctx, quotient_degree_factor=6).merkle_tree.cap.elements == ref6["cap"]).all()
    com6 = Z.commit_quotient_device(tctx, torch.from_numpy(qvals6.view(np.int64)).cuda(), degree_bits, 3, 4, quotient_degree_factor=6)
    torch.cuda.synchronize()
    assert (com6.cap.cpu().numpy().view(np.uint64) == ref6["cap"]).all()
