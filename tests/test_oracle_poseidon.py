"""Pins the oracle's Poseidon against every fixture the reference holds for this path
and against the independent pure-Python derivation in tools/poseidon_derive.py.  CPU only."""
import os
import sys

import numpy as np

from helpers import P, hex_to_elements

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import poseidon_derive as pd  # noqa: E402

# SURVEY.md App. A, first and last rows of ALL_ROUND_CONSTANTS (hex)
ROW00 = "b585f766f2144405 7746a55f43921ad7 b2fb0d31cee799b4 0f6760a4803427d7 e10d666650f4e012 8cae14cb07d09bf1 d438539c95f63e9f ef781c7ce35b4c3d cdc4a239b0c44426 277fa208bf337bff e17653a29da578a1 c54302f225db2c76"
ROW29 = "f3c12fe54d5c653b 40b9e922ed9771e2 551f5b0fbe7b1840 25032aa7c4cb1811 aaed34074b164346 8ffd96bbf9c9c81d 70fc91eb5937085c 7f795e2a5f915440 4543d9df5476d3cb f172d73e004fc90d dfd1c4febcc81238 bc8dfb627fe558fc"


def test_round_constants_regenerated(oracle):
    rc = oracle.round_constants()
    assert [int(x) for x in rc[:12]] == [int(h, 16) for h in ROW00.split()]
    assert [int(x) for x in rc[348:]] == [int(h, 16) for h in ROW29.split()]
    assert [int(x) for x in rc] == pd.round_constants()      # C ChaCha8 == Python ChaCha8
    assert all(int(x) < P for x in rc)


def test_reference_two_to_one_zero(oracle, golden):
    k = golden["reference_poseidon_kats"]
    z = np.zeros(4, np.uint64)
    assert [int(x) for x in oracle.two_to_one(z, z)] == k["two_to_one_zero_zero"]["elements"]


def test_reference_zero_hash_chain(oracle, golden):
    chain = golden["reference_poseidon_kats"]["zero_hash_chain"]["hex"]
    s = np.zeros(4, np.uint64)
    for i, h in enumerate(chain):
        assert (s == hex_to_elements(h)).all(), f"chain entry {i}"
        s = oracle.two_to_one(s, s)


def _zero_tree_root(oracle, depth, z):
    """get_merkle_proof_with_zero over an empty tree (/root/reference/src/merkle_tree/tree.rs:55-91)."""
    root, s = z.copy(), z.copy()
    for _ in range(depth):
        root = oracle.two_to_one(root, s)
        s = oracle.two_to_one(s, s)
    return root


def test_reference_prev_block_header_roots(oracle, golden):
    h = golden["reference_poseidon_kats"]["prev_block_header"]
    z = np.zeros(4, np.uint64)
    assert (_zero_tree_root(oracle, 32, z) == hex_to_elements(h["block_headers_digest"])).all()
    assert (_zero_tree_root(oracle, 4, z) == hex_to_elements(h["deposit_digest"])).all()
    assert (_zero_tree_root(oracle, 4, oracle.two_to_one(z, z)) == hex_to_elements(h["transactions_digest"])).all()


def test_reference_account_addresses(oracle, golden):
    a = golden["reference_poseidon_kats"]["accounts"]
    for sk, hx in zip(a["sk"], a["address_hex"]):
        sk = np.array(sk, dtype=np.uint64)
        assert (oracle.two_to_one(sk, sk) == hex_to_elements(hx)).all()   # full-range inputs


def test_reference_tx_hashes_multi_permutation_sponge(oracle, golden):
    """SMT leaf = hash_pad([key, value, 1]) = a 12-element hash_no_pad (two permutations) -> pins the
    overwrite-mode sponge and the hash_pad rule."""
    t = golden["reference_poseidon_kats"]["tx_hashes"]

    def u(x):
        return np.array([x, 0, 0, 0], dtype=np.uint64)

    def leaf(k, v):
        return oracle.hash_pad(np.concatenate([k, v, np.array([1], np.uint64)]))

    def leaf_explicit(k, v):
        return oracle.hash_no_pad(np.concatenate([k, v, np.array([1, 1, 0, 1], np.uint64)]))

    e3, e4 = t["smt_entries"]
    r2a = leaf(u(e3["key"][1]), leaf(u(e3["key"][2]), u(e3["value"])))
    r2b = leaf(u(e4["key"][1]), leaf(u(e4["key"][2]), u(e4["value"])))
    assert (leaf(u(1), u(2)) == leaf_explicit(u(1), u(2))).all()
    # keys 407 (bit0 = 1) and 832 (bit0 = 0) split at depth 0
    diff_root = oracle.two_to_one(leaf(u(e4["key"][0]), r2b), leaf(u(e3["key"][0]), r2a))
    for nonce, hx in zip(t["nonces"], t["tx_hex"]):
        got = oracle.two_to_one(diff_root, np.array(nonce, dtype=np.uint64))
        assert (got == hex_to_elements(hx)).all()


def test_survey_appendix_b_self_derived(oracle):
    assert int(oracle.permute(np.arange(12, dtype=np.uint64))[0]) == 15442313428170673822
    assert int(oracle.permute(np.full(12, P - 1, np.uint64))[0]) == 13691089994624172887
    assert [int(x) for x in oracle.hash_no_pad(np.arange(135, dtype=np.uint64))] == [
        4848071992462728551, 7985168359107384293, 2979147297992328185, 11181256925898874940]
    assert [int(x) for x in oracle.hash_no_pad(np.arange(9, dtype=np.uint64))] == [  # [0..8] inclusive: 8 + 1 short chunk
        18007381329477297286, 11010590292829788888, 258931329831288973, 9046877563820385107]
    assert [int(x) for x in oracle.two_to_one(np.array([1, 2, 3, 4], np.uint64), np.array([5, 6, 7, 8], np.uint64))] == [
        15064728126975588673, 10314245681893968020, 11300930272442645327, 2830815762300183090]


def test_python_twin_and_equivalent_forms(oracle):
    """C oracle == pure-Python naive == fast-partial == pushed-constant forms (the tables the CUDA side uses)."""
    rng = np.random.default_rng(7)
    rc, M = pd.round_constants(), pd.mds_matrix()
    T = pd.fast_partial_tables(rc, M)
    for _ in range(6):
        v = rng.integers(0, 2**64, size=12, dtype=np.uint64)
        ref = [int(x) for x in oracle.permute(v)]
        vi = [int(x) for x in v]
        assert pd.permute_naive(vi, rc, M) == ref
        assert pd.permute_fast(vi, rc, M, T) == ref
        assert pd.permute_pushed(vi, rc, M) == ref


def test_cpu_baseline_permutation_matches_oracle(oracle):
    import ctypes as C
    bl = oracle.baseline_lib()
    rng = np.random.default_rng(3)
    for v in [np.zeros(12, np.uint64), np.full(12, 2**64 - 1, np.uint64)] + [rng.integers(0, 2**64, size=12, dtype=np.uint64) for _ in range(20)]:
        s = v.copy()
        bl.cpub_permute(s.ctypes.data_as(C.POINTER(C.c_uint64)))
        assert (s == oracle.permute(v)).all()


def test_hash_or_noop_and_short_inputs(oracle):
    x = np.array([5, 2**64 - 1, 7], dtype=np.uint64)
    assert [int(v) for v in oracle.hash_or_noop(x)] == [5, (2**64 - 1) - P, 7, 0]     # not hashed, canonicalised
    assert (oracle.hash_or_noop(np.arange(5, dtype=np.uint64)) == oracle.hash_no_pad(np.arange(5, dtype=np.uint64))).all()
    assert (oracle.hash_no_pad(np.zeros(0, np.uint64)) == 0).all()
