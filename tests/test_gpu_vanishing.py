"""Row N1b on the GPU: b200zkp_dev_quotient_values against oracle/vanishing_ref.py bit for bit, and the whole middle of prove()
on the device — wires commitment -> Z / partial products (N1a) -> their commitment -> quotient values (N1b) -> coset_ifft + chunk
commitment (N1c) — with only challenges and caps crossing PCIe; the committed quotient satisfies the verifier's identity."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P = 0xFFFFFFFF00000001


def _common(zp, c):
    return zp.CommonCircuitData(c.n_log, [(g, c.selector_indices[i], c.groups[c.selector_indices[i]], c.gate_params[i])
                                          for i, g in enumerate(c.gates)],
                                c.num_selectors, quotient_degree_factor=8, k_is=np.array(c.k_is, np.uint64))


@pytest.mark.parametrize("n_log,seed,extended", [(3, 1, False), (5, 2, False), (7, 3, False), (5, 4, True), (8, 5, True)])
def test_quotient_values_match_oracle_and_close_the_identity(oracle, n_log, seed, extended):
    """extended: the circuit also holds PoseidonMds, BaseSum, Reducing, ReducingExtension, ArithmeticExtension, MulExtension,
    Exponentiation and RandomAccess rows (13 gate kinds, 4 selector polynomials)"""
    import torch
    from intmax_zkp_core_b200 import device as D, prover as zp
    from oracle import vanishing_ref as V
    c = V.Circuit(n_log, seed=seed, extended=extended)
    n, r, h = 1 << n_log, 3, min(4, n_log)
    rnd = random.Random(seed + 7)
    betas, gammas, alphas = ([rnd.randrange(P) for _ in range(2)] for _ in range(3))
    ctx = D.torch_context(0)
    dev = lambda cols: torch.from_numpy(np.array(cols, np.uint64).view(np.int64)).cuda()
    cs_vals, wire_vals = dev(c.constants + c.sigmas), dev(c.wires)
    com_cs = D.commit_device(ctx, cs_vals, r, h)
    com_w = D.commit_device(ctx, wire_vals, r, h)
    k_is = np.array(c.k_is, np.uint64)
    zpp_vals = zp.zs_partial_products_device(ctx, wire_vals[:80], dev(c.sigmas), k_is, betas, gammas, 8)
    com_z = D.commit_device(ctx, zpp_vals, r, h)
    common = _common(zp, c)
    q = zp.compute_quotient_values_device(ctx, common, com_cs.lde, com_w.lde, com_z.lde, betas, gammas, alphas, c.pi_hash)
    ctx.synchronize()
    got = q.cpu().numpy().view(np.uint64)
    if n_log <= 5:
        from helpers import bitrev_perm
        inv = np.argsort(bitrev_perm(n_log + r))
        nat = lambda com: [row[inv] for row in com.lde.cpu().numpy().view(np.uint64)]      # leaf order -> natural order
        want = V.quotient_values(c, nat(com_cs), nat(com_w), nat(com_z), betas, gammas, alphas, 8, r, 3)
        assert (got == np.array(want, np.uint64)).all()
    # N1c on the device, then the verifier's identity at a random point from the committed coefficients
    com_q = zp.commit_quotient_device(ctx, q, n_log, r, h)
    ctx.synchronize()
    qc = com_q.coeffs.cpu().numpy().view(np.uint64).reshape(2, 8 * n)
    co = lambda com: com.coeffs.cpu().numpy().view(np.uint64)
    zeta = rnd.randrange(2, P)
    assert V.check_quotient_identity(c, co(com_cs), co(com_w), co(com_z), qc, betas, gammas, alphas, 8, zeta)
    # a witness that breaks one Poseidon wire does not
    bad = np.array(c.wires, np.uint64)
    bad[V.W_FULL1 + 5][0] = (int(bad[V.W_FULL1 + 5][0]) + 1) % P
    com_b = D.commit_device(ctx, torch.from_numpy(bad.view(np.int64)).cuda(), r, h)
    qb = zp.compute_quotient_values_device(ctx, common, com_cs.lde, com_b.lde, com_z.lde, betas, gammas, alphas, c.pi_hash)
    com_qb = zp.commit_quotient_device(ctx, qb, n_log, r, h)
    ctx.synchronize()
    qbc = com_qb.coeffs.cpu().numpy().view(np.uint64).reshape(2, 8 * n)
    assert not V.check_quotient_identity(c, co(com_cs), co(com_b), co(com_z), qbc, betas, gammas, alphas, 8, zeta)
    ctx.close()


def test_quotient_argument_errors():
    import ctypes as C
    import torch
    from intmax_zkp_core_b200 import device as D, prover as zp
    from intmax_zkp_core_b200._lib import B200ZkpError
    ctx = D.torch_context(0)
    common = zp.CommonCircuitData(3, [(zp.GATE_NOOP, 0, (0, 2)), (13, 0, (0, 2))], 1)       # unknown gate kind
    t = lambda k: torch.zeros((k, 64), dtype=torch.int64, device="cuda")
    with pytest.raises(B200ZkpError) as e:
        zp.compute_quotient_values_device(ctx, common, t(83), t(135), t(20), [1, 2], [3, 4], [5, 6], [0, 0, 0, 0])
    assert e.value.code == -4
    # a gate whose parameters need more than 135 wires
    common = zp.CommonCircuitData(3, [(zp.GATE_NOOP, 0, (0, 2)), (zp.GATE_EXPONENTIATION, 0, (0, 2), (70,))], 1)
    with pytest.raises(B200ZkpError) as e:
        zp.compute_quotient_values_device(ctx, common, t(83), t(135), t(20), [1, 2], [3, 4], [5, 6], [0, 0, 0, 0])
    assert e.value.code == -1
    common = zp.CommonCircuitData(3, [(zp.GATE_NOOP, 0, (0, 1))], 1)
    with pytest.raises(ValueError):
        zp.compute_quotient_values_device(ctx, common, t(83), t(135), t(19), [1, 2], [3, 4], [5, 6], [0, 0, 0, 0])
    ctx.close()
