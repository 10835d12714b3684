"""Row N4 of SURVEY.md section 8(f): the reference's OWN native Poseidon tree helpers, on the GPU hash kernels.

Mirrors, with the same names, argument meaning and assertions:
    get_merkle_proof_with_zero / get_merkle_proof / get_merkle_root / MerkleProof   /root/reference/src/merkle_tree/tree.rs:27-128
    PoseidonNodeHash::calc_node_hash (Internal | Leaf)        /root/reference/src/sparse_merkle_tree/goldilocks_poseidon/mod.rs:161-183
    BlockHeader zero-tree digests                              /root/reference/src/transaction/block_header.rs (via get_merkle_proof)

The dense tree over the given leaves is one `b200zkp_merkle_new` call (leaves of 4 elements are not hashed —
hash_or_noop — so the level reduction is exactly the reference's pairwise two_to_one), the padding with zero hashes up to
`depth` is a chain of two_to_one calls.  These trees are tiny in the reference (depth <= 32, a handful of leaves): the point
is parity behind the reference's own interface, pinned by its fixtures, not throughput.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .plonky2 import Context, HashOut, MerkleTree, PoseidonHash, default_context


def log2_ceil(value: int) -> int:
    """tree.rs:9-24 (asserts value != 0)."""
    assert value != 0, "The first argument must be a positive number."
    if value == 1:
        return 0
    log_value, tmp = 1, value - 1
    while tmp > 1:
        tmp //= 2
        log_value += 1
    return log_value


@dataclass
class MerkleProof:
    """tree.rs:28-33: siblings are ordered from the leaf upwards."""
    index: int
    value: HashOut
    siblings: List[HashOut]
    root: HashOut


def _h(x) -> HashOut:
    return x if isinstance(x, HashOut) else HashOut(x)


ZERO = HashOut([0, 0, 0, 0])


def get_merkle_proof_with_zero(leaves: Sequence, index: int, depth: int, zero=ZERO,
                               ctx: Optional[Context] = None) -> MerkleProof:
    """tree.rs:49-99.  `leaves` are filled from the left, the rest of the 2^depth tree is `zero`."""
    ctx = ctx or default_context()
    zero = _h(zero)
    nodes = [_h(x) for x in leaves] if len(leaves) else [zero]
    assert index < len(nodes)
    assert len(nodes) <= 1 << depth
    num_leaves = 1 << (len(nodes) - 1).bit_length()
    log_num_leaves = log2_ceil(num_leaves)
    value = nodes[index]
    nodes = nodes + [zero] * (num_leaves - len(nodes))

    # zero hashes: siblings[i] = hash of an all-zero subtree of height i
    siblings = [zero]
    for _ in range(1, depth):
        last = siblings[-1]
        siblings.append(PoseidonHash.two_to_one(last, last, ctx))

    if log_num_leaves > 0:
        # dense part on the device: 4-element leaves are their own digests, levels are two_to_one
        tree = MerkleTree.new(np.stack([x.elements for x in nodes]), 0, ctx=ctx)
        path = tree.prove(index).siblings
        for i in range(min(log_num_leaves, len(siblings))):
            siblings[i] = HashOut(path[i])
        root = tree.cap[0]
    else:
        root = nodes[0]
    for sib in siblings[log_num_leaves:]:
        root = PoseidonHash.two_to_one(root, sib, ctx)      # above the dense part the sibling is always on the right
    return MerkleProof(index=index, value=value, siblings=siblings, root=root)


def get_merkle_proof(leaves: Sequence, index: int, depth: int, ctx: Optional[Context] = None) -> MerkleProof:
    """tree.rs:101-107."""
    return get_merkle_proof_with_zero(leaves, index, depth, ZERO, ctx)


def get_merkle_root(index: int, value, siblings: Sequence, ctx: Optional[Context] = None) -> HashOut:
    """tree.rs:109-128."""
    ctx = ctx or default_context()
    root, rest = _h(value), index
    for sib in siblings:
        sib = _h(sib)
        root = PoseidonHash.two_to_one(root, sib, ctx) if rest & 1 == 0 else PoseidonHash.two_to_one(sib, root, ctx)
        rest >>= 1
    return root


class PoseidonNodeHash:
    """goldilocks_poseidon/mod.rs:161-183."""

    @staticmethod
    def calc_internal(left, right, ctx: Optional[Context] = None) -> HashOut:
        return PoseidonHash.two_to_one(_h(left), _h(right), ctx)

    @staticmethod
    def calc_leaf(key, value, ctx: Optional[Context] = None) -> HashOut:
        k, v = _h(key), _h(value)
        return PoseidonHash.hash_pad(np.concatenate([k.elements, v.elements, np.array([1], np.uint64)]), ctx)

    @staticmethod
    def calc_leaf_batch(keys, values, ctx: Optional[Context] = None) -> np.ndarray:
        """Many SMT leaf hashes at once: hash_pad([k, v, 1]) = hash_no_pad([k0..3, v0..3, 1, 1, 0, 1])."""
        k = np.asarray(keys, dtype=np.uint64).reshape(-1, 4)
        v = np.asarray(values, dtype=np.uint64).reshape(-1, 4)
        pad = np.tile(np.array([1, 1, 0, 1], dtype=np.uint64), (k.shape[0], 1))
        return PoseidonHash.hash_no_pad_batch(np.concatenate([k, v, pad], axis=1), ctx)


LOG_MAX_N_BLOCKS = 32   # /root/reference/src/rollup/circuits/mod.rs:67


@dataclass
class BlockHeader:
    """/root/reference/src/transaction/block_header.rs:23-32, `new` at :126-150."""
    block_number: int
    prev_block_hash: HashOut
    block_headers_digest: HashOut
    transactions_digest: HashOut
    deposit_digest: HashOut
    proposed_world_state_digest: HashOut
    approved_world_state_digest: HashOut
    latest_account_digest: HashOut

    @staticmethod
    def new(log_num_txs_in_block: int, ctx: Optional[Context] = None) -> "BlockHeader":
        default_tx_hash = PoseidonHash.two_to_one(ZERO, ZERO, ctx)   # MergeAndPurgeTransitionPublicInputs::default().tx_hash
        return BlockHeader(
            block_number=0, prev_block_hash=ZERO,
            block_headers_digest=get_merkle_proof([], 0, LOG_MAX_N_BLOCKS, ctx).root,
            transactions_digest=get_merkle_proof_with_zero([], 0, log_num_txs_in_block, default_tx_hash, ctx).root,
            deposit_digest=get_merkle_proof_with_zero([], 0, log_num_txs_in_block, ZERO, ctx).root,
            proposed_world_state_digest=ZERO, approved_world_state_digest=ZERO, latest_account_digest=ZERO)


def get_block_hash(h: BlockHeader, ctx: Optional[Context] = None) -> HashOut:
    """block_header.rs:157-174."""
    t = PoseidonHash.two_to_one
    a = t(HashOut([h.block_number, 0, 0, 0]), h.latest_account_digest, ctx)
    b = t(h.deposit_digest, h.transactions_digest, ctx)
    c = t(a, b, ctx)
    d = t(h.proposed_world_state_digest, h.approved_world_state_digest, ctx)
    e = t(c, d, ctx)
    return t(h.block_headers_digest, e, ctx)
