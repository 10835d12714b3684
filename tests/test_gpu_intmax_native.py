"""Row N4: the reference's own native Poseidon tree helpers (src/merkle_tree/tree.rs, PoseidonNodeHash, BlockHeader)
behind their own interface, computed by the GPU hash kernels and pinned by the reference's fixtures."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_proof(oracle, leaves, index, depth, zero):
    """Straight restatement of get_merkle_proof_with_zero with the oracle's two_to_one."""
    nodes = [np.asarray(x, dtype=np.uint64) for x in leaves] or [zero]
    num = 1 << (len(nodes) - 1).bit_length()
    lg = num.bit_length() - 1
    nodes = nodes + [zero] * (num - len(nodes))
    sib = [zero]
    for _ in range(1, depth):
        sib.append(oracle.two_to_one(sib[-1], sib[-1]))
    rest = index
    for i in range(min(lg, len(sib))):
        sib[i] = nodes[rest ^ 1]
        nodes = [oracle.two_to_one(nodes[2 * j], nodes[2 * j + 1]) for j in range(len(nodes) // 2)]
        rest >>= 1
    root = nodes[0]
    for s in sib[lg:]:
        root = oracle.two_to_one(root, s)
    return sib, root


def test_block_header_fixture(ctx, golden):
    """BlockHeader::new(4) and BlockDetail::new(4) of /root/reference/src/rollup/circuits/mod.rs:93-109."""
    from intmax_zkp_core_b200 import intmax as I
    k = golden["reference_poseidon_kats"]
    hdr = I.BlockHeader.new(4, ctx)
    assert hdr.block_headers_digest.to_hex() == k["prev_block_header"]["block_headers_digest"]
    assert hdr.transactions_digest.to_hex() == k["prev_block_header"]["transactions_digest"]
    assert hdr.deposit_digest.to_hex() == k["prev_block_header"]["deposit_digest"]
    prev_block_hash = I.get_block_hash(hdr, ctx)
    proof = I.get_merkle_proof([prev_block_hash], 0, I.LOG_MAX_N_BLOCKS, ctx)
    assert [s.to_hex() for s in proof.siblings] == k["zero_hash_chain"]["hex"]
    assert I.get_merkle_root(0, prev_block_hash, proof.siblings, ctx) == proof.root


@pytest.mark.parametrize("n_leaves,index,depth", [(0, 0, 5), (1, 0, 0), (1, 0, 3), (2, 1, 1), (5, 3, 6), (8, 7, 3), (13, 4, 32)])
def test_merkle_proof_with_zero_matches_oracle(ctx, oracle, n_leaves, index, depth):
    from intmax_zkp_core_b200 import intmax as I
    rng = np.random.default_rng(n_leaves * 100 + depth)
    leaves = [rng.integers(0, 2**64, size=4, dtype=np.uint64) % np.uint64(0xFFFFFFFF00000001) for _ in range(n_leaves)]
    zero = rng.integers(0, 2**63, size=4, dtype=np.uint64)
    p = I.get_merkle_proof_with_zero(leaves, index, depth, zero, ctx)
    sib, root = _oracle_proof(oracle, leaves, index, depth, zero)
    assert len(p.siblings) == len(sib)
    for a, b in zip(p.siblings, sib):
        assert (a.elements == b).all()
    assert (p.root.elements == root).all()
    if depth > 0:
        assert I.get_merkle_root(index, p.value, p.siblings, ctx) == p.root
    if depth <= 6:     # more leaves than 2^depth: the reference asserts
        with pytest.raises(AssertionError):
            I.get_merkle_proof_with_zero(leaves + [zero] * ((1 << depth) + 1), 0, depth, zero, ctx)


def test_node_hash_fixture(ctx, oracle, golden):
    """PoseidonNodeHash on the SMT entries of src/bin/block_circuit.rs:108-123 reproduces the fixture's tx hashes."""
    from intmax_zkp_core_b200 import intmax as I
    import intmax_zkp_core_b200 as z
    t = golden["reference_poseidon_kats"]["tx_hashes"]

    def u(x):
        return z.HashOut([x, 0, 0, 0])

    e3, e4 = t["smt_entries"]
    N = I.PoseidonNodeHash
    r2a = N.calc_leaf(u(e3["key"][1]), N.calc_leaf(u(e3["key"][2]), u(e3["value"]), ctx), ctx)
    r2b = N.calc_leaf(u(e4["key"][1]), N.calc_leaf(u(e4["key"][2]), u(e4["value"]), ctx), ctx)
    diff_root = N.calc_internal(N.calc_leaf(u(e4["key"][0]), r2b, ctx), N.calc_leaf(u(e3["key"][0]), r2a, ctx), ctx)
    for nonce, hx in zip(t["nonces"], t["tx_hex"]):
        assert z.PoseidonHash.two_to_one(diff_root, nonce, ctx).to_hex() == hx
    keys = np.arange(40, dtype=np.uint64).reshape(10, 4)
    vals = keys * np.uint64(7) + np.uint64(1)
    got = N.calc_leaf_batch(keys, vals, ctx)
    for i in range(10):
        assert (got[i] == oracle.hash_pad(np.concatenate([keys[i], vals[i], np.array([1], np.uint64)]))).all()
