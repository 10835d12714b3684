"""GPU parity of the opening proof (rows N2 + N3): intmax_zkp_core_b200/fri.py through the C ABI against the CPU restatement
oracle/fri_ref.py on the same transcript — bit-exact final_poly, commit-phase caps, proof-of-work witness and query answers at
small sizes; at circuit-like sizes the proof produced on the GPU must pass the oracle's verifier."""
import ctypes as C

import numpy as np
import pytest

from helpers import P, rand_field

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zf():
    import intmax_zkp_core_b200.fri as zf
    return zf


def rand_ext(rng):
    return (int(rng.integers(1, P, dtype=np.uint64)), int(rng.integers(1, P, dtype=np.uint64)))


def to_pairs(a):
    return [(int(x), int(y)) for x, y in np.asarray(a).reshape(-1, 2)]


def build_case(zf, ctx, O, F, rng, n_log, ks, rate_bits, cap_height, with_oracle_commits=True):
    import intmax_zkp_core_b200 as z
    n = 1 << n_log
    data = [rand_field(rng, (k, n), canonical=False) for k in ks]
    gpu = [z.PolynomialBatch.from_coeffs(d, rate_bits, False, cap_height, ctx=ctx) for d in data]
    cpu = [O.commit(d, rate_bits, cap_height, is_coeffs=True) for d in data] if with_oracle_commits else None
    zeta = rand_ext(rng)
    gzeta = F.escale(zeta, F.root(n_log))
    all_polys = [(o, i) for o, k in enumerate(ks) for i in range(k)]
    zs = [(len(ks) - 1, i) for i in range(min(2, ks[-1]))]
    batches = [(zeta, all_polys), (gzeta, zs)]
    instance = zf.FriInstanceInfo([zf.FriBatchInfo(pt, [zf.FriPolynomialInfo(o, i) for o, i in polys]) for pt, polys in batches])
    return data, gpu, cpu, batches, instance


@pytest.mark.parametrize("mul_by_x", [True, False])
@pytest.mark.parametrize("n_log,ks,rate_bits", [(0, (2,), 1), (1, (1, 1), 0), (5, (3, 2), 2), (10, (4, 1, 3), 3), (12, (2, 3), 1),
                                                (13, (1, 2), 2)])
def test_final_poly_matches_oracle(zf, ctx, oracle, n_log, ks, rate_bits, mul_by_x):
    """reduce_polys_base + divide_by_linear + shift_poly (+ the factor X): the device scan against the serial definition,
    including degrees above one scan segment (1024) and non-canonical inputs."""
    from oracle import fri_ref as F
    rng = np.random.default_rng(31 * n_log + len(ks))
    data, gpu, _, batches, instance = build_case(zf, ctx, oracle, F, rng, n_log, ks, rate_bits, 0, with_oracle_commits=False)
    alpha = rand_ext(rng)
    canon = [np.where(d >= np.uint64(P), d - np.uint64(P), d) for d in data]
    want = F.final_poly_of(canon, batches, alpha, mul_by_x)
    st = zf.FriCommitPhase.from_oracles(instance, gpu, alpha, mul_by_x, ctx=ctx)
    got = st.coeffs()
    n = 1 << n_log
    assert to_pairs(got[:n]) == want
    assert not got[n:].any()
    st.close()


@pytest.mark.parametrize("n_log,rate_bits,cap_height,arities", [(5, 2, 1, (2, 1)), (4, 3, 0, (1, 1, 1)), (6, 1, 2, (3,)),
                                                                (8, 3, 4, (4,)), (3, 3, 2, ()), (7, 0, 0, (4, 3))])
def test_commit_phase_matches_oracle(zf, ctx, oracle, n_log, rate_bits, cap_height, arities):
    """fri_committed_trees: layer caps, folded coefficients, final_poly and every layer's query answers."""
    from oracle import fri_ref as F
    rng = np.random.default_rng(5 + n_log)
    coeffs = rand_field(rng, (1 << n_log, 2), canonical=False)
    canon = np.where(coeffs >= np.uint64(P), coeffs - np.uint64(P), coeffs)
    want = F.fri_proof([], to_pairs(canon), F.Challenger(), rate_bits, cap_height, arities, 3, 4)
    st = zf.FriCommitPhase.from_coeffs(coeffs, rate_bits, ctx=ctx)
    ch = zf.Challenger(ctx)
    for i, ab in enumerate(arities):
        cap = st.commit_layer(ab, cap_height)
        assert (cap.flatten().reshape(-1, 4) == want["caps"][i]).all(), f"layer {i} cap"
        ch.observe_cap(cap)
        st.fold(ch.get_extension_challenge())
    fp = st.final_poly()
    assert to_pairs(fp) == want["final_poly"]
    assert st.shape()["layers"] == len(arities)
    lde_bits = n_log + rate_bits
    for rnd in want["rounds"]:
        x, bits = rnd["x_index"], lde_bits
        for i, ab in enumerate(arities):
            x >>= ab
            bits -= ab
            evals, sib = st.query(i, [x], ab, bits - cap_height)
            assert (evals[0] == rnd["steps"][i][0]).all() and (sib[0] == rnd["steps"][i][1]).all()
    st.close()


@pytest.mark.parametrize("mul_by_x", [True, False])
@pytest.mark.parametrize("n_log,ks,rate_bits,cap_height,arities", [
    (5, (3, 2), 2, 1, (2, 1)),
    (4, (1, 4, 2), 3, 0, (1, 1, 1)),
    (6, (5,), 1, 2, (3,)),
    (3, (2, 2), 3, 2, ()),
    (8, (6, 3), 3, 4, (4,)),
])
def test_prove_openings_matches_oracle(zf, ctx, oracle, n_log, ks, rate_bits, cap_height, arities, mul_by_x):
    from oracle import fri_ref as F
    rng = np.random.default_rng(77 + n_log)
    data, gpu, cpu, batches, instance = build_case(zf, ctx, oracle, F, rng, n_log, ks, rate_bits, cap_height)
    pow_bits, nq = 6, 7
    cfg = zf.FriConfig(rate_bits=rate_bits, cap_height=cap_height, proof_of_work_bits=pow_bits, num_query_rounds=nq)
    params = zf.FriParams(config=cfg, hiding=False, degree_bits=n_log, reduction_arity_bits=list(arities))

    ch_g = zf.Challenger(ctx)
    ch_c = F.Challenger()
    for b, c in zip(gpu, cpu):
        assert (b._cap.flatten().reshape(-1, 4) == c["cap"]).all()
        ch_g.observe_cap(b._cap)
        ch_c.observe_cap(c["cap"])
    got = zf.prove_openings(instance, gpu, ch_g, params, mul_by_x)
    want = F.prove_openings(cpu, batches, ch_c, rate_bits, cap_height, arities, pow_bits, nq, mul_by_x)

    assert len(got.commit_phase_merkle_caps) == len(arities)
    for a, b in zip(got.commit_phase_merkle_caps, want["caps"]):
        assert (a.flatten().reshape(-1, 4) == b).all()
    assert to_pairs(got.final_poly) == want["final_poly"]
    assert got.pow_witness == want["pow_witness"]
    assert len(got.query_round_proofs) == nq
    for rg, rc in zip(got.query_round_proofs, want["rounds"]):
        for (row_g, proof_g), (row_c, sib_c) in zip(rg.initial_trees_proof.evals_proofs, rc["initial"]):
            assert (row_g == row_c).all() and (np.asarray(proof_g.siblings).reshape(-1, 4) == sib_c).all()
        for sg, (ev_c, sib_c) in zip(rg.steps, rc["steps"]):
            assert (sg.evals == ev_c).all() and (np.asarray(sg.merkle_proof.siblings).reshape(-1, 4) == sib_c).all()
    # both transcripts end in the same state
    assert ch_g.get_challenge() == ch_c.get_challenge()


def proof_to_oracle_form(proof):
    return dict(
        caps=[c.flatten().reshape(-1, 4) for c in proof.commit_phase_merkle_caps],
        final_poly=to_pairs(proof.final_poly), pow_witness=proof.pow_witness,
        rounds=[dict(initial=[(row, np.asarray(p.siblings).reshape(-1, 4)) for row, p in r.initial_trees_proof.evals_proofs],
                     steps=[(s.evals, np.asarray(s.merkle_proof.siblings).reshape(-1, 4)) for s in r.steps])
                for r in proof.query_round_proofs])


@pytest.mark.parametrize("n_log,ks", [(12, (85, 135, 20, 16)), (14, (9, 5, 4)), (16, (3, 2))])
def test_standard_config_proof_passes_the_oracle_verifier(zf, ctx, oracle, n_log, ks):
    """The reference's only configuration (standard_recursion_config: rate 3, cap 4, arity 16, 16 PoW bits, 28 queries) at the
    sizes its circuits have: the opening proof made on the GPU is accepted by the CPU verifier restatement, with the openings
    themselves evaluated on the GPU; a corrupted opening is rejected."""
    from oracle import fri_ref as F
    rng = np.random.default_rng(n_log)
    cfg = zf.standard_recursion_fri_config()
    params = cfg.fri_params(n_log)
    assert params.reduction_arity_bits == [4] * ((n_log - 5 + 3) // 4)
    _, gpu, _, batches, instance = build_case(zf, ctx, oracle, F, rng, n_log, ks, cfg.rate_bits, cfg.cap_height, with_oracle_commits=False)
    ch = zf.Challenger(ctx)
    for b in gpu:
        ch.observe_cap(b._cap)
    evals = [b.eval_ext2(np.array(batches[0][0], dtype=np.uint64)) for b in gpu]
    evals_g = gpu[-1].eval_ext2(np.array(batches[1][0], dtype=np.uint64))
    openings = [[tuple(int(v) for v in evals[o][i]) for o, i in batches[0][1]], [tuple(int(v) for v in evals_g[i]) for _, i in batches[1][1]]]
    proof = zf.prove_openings(instance, gpu, ch, params, True)
    assert 64 - int(oracle.hash_no_pad(list(_pow_hash(F, gpu, proof)) + [proof.pow_witness])[0]).bit_length() >= 16

    def fresh():
        c = F.Challenger()
        for b in gpu:
            c.observe_cap(b._cap.flatten())
        return c
    caps = [b._cap.flatten().reshape(-1, 4) for b in gpu]
    args = (n_log, cfg.rate_bits, cfg.cap_height, params.reduction_arity_bits, cfg.proof_of_work_bits, cfg.num_query_rounds, True)
    form = proof_to_oracle_form(proof)
    assert F.verify(form, caps, batches, openings, fresh(), *args)
    openings[0][3] = ((openings[0][3][0] + 1) % P, openings[0][3][1])
    assert not F.verify(form, caps, batches, openings, fresh(), *args)


def _pow_hash(F, gpu, proof):
    """Replays the transcript up to the proof-of-work hash."""
    c = F.Challenger()
    for b in gpu:
        c.observe_cap(b._cap.flatten())
    c.get_extension_challenge()
    for cap in proof.commit_phase_merkle_caps:
        c.observe_cap(cap.flatten())
        c.get_extension_challenge()
    c.observe_extension_elements(to_pairs(proof.final_poly))
    return c.get_hash()


def test_pow_grind_smallest_witness(zf, ctx, oracle):
    from oracle import fri_ref as F
    import intmax_zkp_core_b200 as z
    for seed, bits in [(1, 0), (2, 5), (3, 9), (4, 12)]:
        rng = np.random.default_rng(seed)
        h = z.HashOut(rand_field(rng, 4))
        w = zf.fri_proof_of_work(h, zf.FriConfig(proof_of_work_bits=bits), ctx)
        assert w == F.proof_of_work([int(v) for v in h.elements], bits)
    # 16 bits (the reference's setting): the witness qualifies and nothing below it does (checked with the batched GPU hash
    # against the oracle on a sample)
    h = z.HashOut(rand_field(np.random.default_rng(9), 4))
    w = zf.fri_proof_of_work(h, zf.FriConfig(proof_of_work_bits=16), ctx)
    lz = lambda x: 64 - int(oracle.hash_no_pad([int(v) for v in h.elements] + [x])[0]).bit_length()
    assert lz(w) >= 16
    rows = np.tile(np.concatenate([np.asarray(h.elements, np.uint64), np.zeros(1, np.uint64)]), (w, 1))
    rows[:, 4] = np.arange(w, dtype=np.uint64)
    first = z.PoseidonHash.hash_no_pad_batch(rows, ctx)[:, 0] if w else np.zeros(0, np.uint64)
    assert not (first < np.uint64(1 << 48)).any()
    # the duplex form (state position 5, response word 7) and a bound that is too small
    state = rand_field(np.random.default_rng(10), 12)
    wit = C.c_uint64()
    ctx.check(ctx._lib.b200zkp_pow_grind(ctx._h, state.ctypes.data_as(C.c_void_p), 5, 7, 8, 0, C.byref(wit)))
    s = state.copy(); s[5] = wit.value
    assert int(oracle.permute(s)[7]) >> 56 == 0
    for x in range(wit.value):
        s[5] = x
        assert int(oracle.permute(s)[7]) >> 56 != 0
    rc = ctx._lib.b200zkp_pow_grind(ctx._h, state.ctypes.data_as(C.c_void_p), 5, 7, 40, 1000, C.byref(wit))
    assert rc == -4


def test_fri_argument_checks(zf, ctx):
    import intmax_zkp_core_b200 as z
    st = zf.FriCommitPhase.from_coeffs(np.ones((8, 2), np.uint64), 1, ctx=ctx)
    with pytest.raises(z.B200ZkpError):
        st.fold((1, 0))                       # nothing committed yet
    with pytest.raises(z.B200ZkpError):
        st.commit_layer(5, 0)                 # arity above the LDE size
    with pytest.raises(z.B200ZkpError):
        st.commit_layer(2, 3)                 # cap above the leaf count
    st.commit_layer(2, 1)
    with pytest.raises(z.B200ZkpError):
        st.commit_layer(1, 0)                 # previous layer not folded
    with pytest.raises(z.B200ZkpError):
        st.final_poly()
    st.fold((3, 4))
    assert st.final_poly().shape == (2, 2)
    with pytest.raises(z.B200ZkpError):
        st.query(0, [4], 2, 1)                # leaf index out of range
    st.close()
    a = z.PolynomialBatch.from_coeffs(np.ones((1, 8), np.uint64), 1, False, 0, ctx=ctx)
    b = z.PolynomialBatch.from_coeffs(np.ones((1, 16), np.uint64), 1, False, 0, ctx=ctx)
    inst = zf.FriInstanceInfo([zf.FriBatchInfo((1, 2), [zf.FriPolynomialInfo(0, 0), zf.FriPolynomialInfo(1, 0)])])
    with pytest.raises(z.B200ZkpError):
        zf.FriCommitPhase.from_oracles(inst, [a, b], (5, 6), True, ctx=ctx)      # degrees differ
    inst = zf.FriInstanceInfo([zf.FriBatchInfo((1, 2), [zf.FriPolynomialInfo(0, 1)])])
    with pytest.raises(z.B200ZkpError):
        zf.FriCommitPhase.from_oracles(inst, [a], (5, 6), True, ctx=ctx)         # polynomial index out of range


def test_hypothesis_opening_proofs(zf, ctx, oracle):
    """Random (degree, oracle shapes, rate, cap height, reduction arities, X-factor convention, opening points incl. base-field
    and zero ones) against the restated prover, bit for bit, and through the restated verifier."""
    from hypothesis import given, settings, strategies as st, HealthCheck
    from oracle import fri_ref as F
    import intmax_zkp_core_b200 as z

    @settings(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck))
    @given(st.integers(1, 8), st.lists(st.integers(1, 5), min_size=1, max_size=3), st.integers(0, 3), st.integers(0, 3),
           st.lists(st.integers(1, 4), max_size=3), st.booleans(), st.integers(0, 2**32 - 1), st.integers(0, 3))
    def run(n_log, ks, r, h, arities, mul_by_x, seed, point_kind):
        n = 1 << n_log
        # keep every layer tree non-degenerate the way plonky2's strategies do: arity <= remaining LDE bits - cap height
        arities, bits = list(arities), n_log + r
        valid = []
        for ab in arities:
            if ab <= bits - h and ab <= n_log - sum(valid):
                valid.append(ab)
                bits -= ab
        h = min(h, n_log + r)
        rng = np.random.default_rng(seed)
        data = [rand_field(rng, (k, n), canonical=False) for k in ks]
        gpu = [z.PolynomialBatch.from_coeffs(d, r, False, h, ctx=ctx) for d in data]
        cpu = [oracle.commit(d, r, h, is_coeffs=True) for d in data]
        zeta = [rand_ext(rng), (int(rng.integers(1, P, dtype=np.uint64)), 0), (0, 0), (P - 1, P - 1)][point_kind]
        batches = [(zeta, [(o, i) for o, k in enumerate(ks) for i in range(k)]),
                   (F.escale(zeta, F.root(n_log)), [(len(ks) - 1, 0)])]
        inst = zf.FriInstanceInfo([zf.FriBatchInfo(pt, [zf.FriPolynomialInfo(o, i) for o, i in polys]) for pt, polys in batches])
        cfg = zf.FriConfig(rate_bits=r, cap_height=h, proof_of_work_bits=3, num_query_rounds=3)
        params = zf.FriParams(config=cfg, hiding=False, degree_bits=n_log, reduction_arity_bits=valid)
        ch_g, ch_c = zf.Challenger(ctx), F.Challenger()
        for b, c in zip(gpu, cpu):
            ch_g.observe_cap(b._cap)
            ch_c.observe_cap(c["cap"])
        got = zf.prove_openings(inst, gpu, ch_g, params, mul_by_x)
        want = F.prove_openings(cpu, batches, ch_c, r, h, valid, 3, 3, mul_by_x)
        form = proof_to_oracle_form(got)
        assert form["final_poly"] == want["final_poly"] and form["pow_witness"] == want["pow_witness"]
        assert all((a == b).all() for a, b in zip(form["caps"], want["caps"]))
        for rg, rc in zip(form["rounds"], want["rounds"]):
            assert all((a[0] == b[0]).all() and (a[1] == b[1]).all() for a, b in zip(rg["initial"], rc["initial"]))
            assert all((a[0] == b[0]).all() and (a[1] == b[1]).all() for a, b in zip(rg["steps"], rc["steps"]))
        # (x - z = 0 needs z on the LDE coset 7<w_N>: neither the random nor the special points above are)
        openings = [[tuple(int(v) for v in oracle.eval_ext2(cpu[o]["coeffs"][i], np.array(pt, dtype=np.uint64))) for o, i in polys]
                    for pt, polys in batches]
        fresh = F.Challenger()
        for c in cpu:
            fresh.observe_cap(c["cap"])
        if True:
            assert F.verify(form, [c["cap"] for c in cpu], batches, openings, fresh, n_log, r, h, valid, 3, 3, mul_by_x)

    run()


def test_golden_opening_proof_vectors_on_gpu(zf, ctx, oracle, golden):
    """The committed opening-proof vectors (tests/golden/opening_proof_vectors.json) reproduced through the C ABI."""
    import os
    import sys
    import intmax_zkp_core_b200 as z
    from oracle import fri_ref as F
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden as G
    for case in golden["opening_proof_vectors"]["cases"]:
        n_log, ks, r, h = case["n_log"], case["ks"], case["rate_bits"], case["cap_height"]
        gpu = [z.PolynomialBatch.from_coeffs(oracle.synthetic_values(k, 1 << n_log, seed=o + 1), r, False, h, ctx=ctx)
               for o, k in enumerate(ks)]
        zeta = tuple(case["zeta"])
        inst = zf.FriInstanceInfo([
            zf.FriBatchInfo(zeta, [zf.FriPolynomialInfo(o, i) for o, k in enumerate(ks) for i in range(k)]),
            zf.FriBatchInfo(F.escale(zeta, F.root(n_log)), [zf.FriPolynomialInfo(len(ks) - 1, 0)])])
        cfg = zf.FriConfig(rate_bits=r, cap_height=h, proof_of_work_bits=case["pow_bits"], num_query_rounds=case["num_queries"])
        params = zf.FriParams(config=cfg, hiding=False, degree_bits=n_log, reduction_arity_bits=list(case["arities"]))
        ch = zf.Challenger(ctx)
        for b in gpu:
            ch.observe_cap(b._cap)
        proof = proof_to_oracle_form(zf.prove_openings(inst, gpu, ch, params, case["mul_by_x"]))
        assert [list(e) for e in proof["final_poly"]] == case["final_poly"]
        assert proof["pow_witness"] == case["pow_witness"]
        assert [[int(x) for x in cap[0]] for cap in proof["caps"]] == case["commit_phase_caps_row0"]
        assert "%016x" % G.fnv(G.proof_words(proof)) == case["proof_fnv1a64"]
