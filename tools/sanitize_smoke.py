#!/usr/bin/env python3
"""Small commitments + accessors for compute-sanitizer (memcheck / racecheck) runs on the GPU box:
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import intmax_zkp_core_b200 as z
from oracle import oracle as O

ctx = z.Context(0)
ONLY = os.environ.get("SANITIZE_ONLY")      # "new": only the kernels added last (latency-form hashing, permutation argument, coset_ifft)
shapes = [(0, 3, 3, 2, False, False), (3, 9, 3, 2, False, False), (7, 5, 1, 0, True, False),
          (9, 20, 3, 4, False, True), (10, 135, 3, 4, False, False), (13, 4, 2, 6, False, False)]
if ONLY == "new":
    shapes = [(3, 9, 3, 2, False, False), (9, 20, 3, 4, False, False)]
for (n_log, k, r, h, coeffs, salted) in shapes:
    v = O.synthetic_values(k, 1 << n_log, seed=1)
    salt = O.synthetic_values(4, 1 << (n_log + r), seed=2) if salted else None
    ctor = z.PolynomialBatch.from_coeffs if coeffs else z.PolynomialBatch.from_values
    b = ctor(v, r, salted, h, salt=salt, ctx=ctx, copy_back=(n_log == 9))
    ref = O.commit(v, r, h, is_coeffs=coeffs, salt=salt)
    assert (b.merkle_tree.cap.elements == ref["cap"]).all()
    assert (b.merkle_tree.leaves == ref["leaves"]).all()
    assert (b.merkle_tree.digests == ref["digests"]).all()
    N = 1 << (n_log + r)
    rows, sib = b.rows([0, N - 1, N // 2])
    assert (rows[1] == ref["leaves"][N - 1]).all()
    b.get_lde_values(0)
x = np.arange(5 * 64, dtype=np.uint64).reshape(5, 64)
assert (z.fft_batch(z.ifft_batch(x, ctx), ctx) == x).all()
z.coset_lde_batch(x, 2, ctx)
t = z.MerkleTree.new(np.arange(64 * 7, dtype=np.uint64).reshape(64, 7), 2, ctx=ctx)
t.prove(5); t.digests
z.PoseidonHash.hash_no_pad_batch(np.arange(40, dtype=np.uint64).reshape(4, 10), ctx)
z.PoseidonHash.two_to_one_batch(np.arange(12, dtype=np.uint64).reshape(3, 4), np.arange(12, 24, dtype=np.uint64).reshape(3, 4), ctx)
z.PoseidonPermutation.permute(np.arange(12, dtype=np.uint64), ctx)
assert (z.coset_ifft_batch(np.stack([O.coset_lde(row, 0) for row in x]), 7, ctx) == x).all()
# opening proof (rows N2 + N3): evaluation, alpha-reduction, blocked suffix scan over > 1 segment, FRI layers, PoW, gathers
import intmax_zkp_core_b200.fri as zf
fri_cases = [(12, (3, 2), 1, 2, (4, 3), True), (5, (2,), 3, 0, (2, 1), False), (1, (1, 1), 0, 0, (), True)]
for (n_log, ks, r, h, arities, mul_by_x) in ([] if ONLY == "new" else fri_cases):
    batches = [z.PolynomialBatch.from_coeffs(O.synthetic_values(k, 1 << n_log, seed=3 + i), r, False, h, ctx=ctx) for i, k in enumerate(ks)]
    batches[0].eval_ext2(np.array([5, 6], np.uint64))
    inst = zf.FriInstanceInfo([zf.FriBatchInfo((11, 12), [zf.FriPolynomialInfo(o, i) for o, k in enumerate(ks) for i in range(k)]),
                               zf.FriBatchInfo((13, 0), [zf.FriPolynomialInfo(0, 0)])])
    cfg = zf.FriConfig(rate_bits=r, cap_height=h, proof_of_work_bits=5, num_query_rounds=4)
    params = zf.FriParams(config=cfg, hiding=False, degree_bits=n_log, reduction_arity_bits=list(arities))
    ch = zf.Challenger(ctx)
    for b in batches:
        ch.observe_cap(b._cap)
    proof = zf.prove_openings(inst, batches, ch, params, mul_by_x)
    assert len(proof.query_round_proofs) == 4
# permutation argument (row N1a): ragged last chunk, more than one scan block, a single row
from intmax_zkp_core_b200 import prover as zp
from oracle import perm_ref as PR
for (R, deg, n_log, Cn) in [(13, 4, 11, 2), (80, 8, 5, 2), (3, 8, 0, 1)]:
    wires, sigmas, k_is = PR.valid_permutation_instance(R, n_log, seed=R)
    got = zp.zs_partial_products(np.array(wires, np.uint64), np.array(sigmas, np.uint64), np.array(k_is, np.uint64),
                                 list(range(5, 5 + Cn)), list(range(9, 9 + Cn)), deg, ctx)
    ref = PR.partial_products_and_zs(wires, sigmas, k_is, list(range(5, 5 + Cn)), list(range(9, 9 + Cn)), deg)
    assert (got == np.array(ref, np.uint64)).all()
print("sanitize smoke ok, launches:", ctx.launch_count)
ctx.close()
