"""intmax-zkp-core_b200 — B200-native polynomial commitment (iNTT -> coset LDE -> Poseidon Merkle cap).

One hot path of intmax-zkp-core's plonky2 prover, behind plonky2's own API names:

    from intmax_zkp_core_b200 import PolynomialBatch, MerkleTree, PoseidonHash

The compute lives in csrc/ (hand-written CUDA for sm_100a) behind the C ABI in include/b200zkp.h; this
package is the host-side mirror of the reference interface.  There is no CPU fallback.
"""
from ._lib import B200ZkpError, build, lib  # noqa: F401
from .plonky2 import (  # noqa: F401
    Context, HashOut, MerkleCap, MerkleProof, MerkleTree, PolynomialBatch, PolynomialCoeffs, PolynomialValues,
    PoseidonHash, PoseidonPermutation, SALT_SIZE, coset_ifft_batch, coset_lde_batch, default_context, fft_batch, ifft_batch,
    log2_strict, reverse_bits, verify_merkle_proof_to_cap,
)

__all__ = [
    "B200ZkpError", "Context", "HashOut", "MerkleCap", "MerkleProof", "MerkleTree", "PolynomialBatch",
    "PolynomialCoeffs", "PolynomialValues", "PoseidonHash", "PoseidonPermutation", "SALT_SIZE", "build",
    "coset_ifft_batch", "coset_lde_batch", "default_context", "fft_batch", "ifft_batch", "lib", "log2_strict", "reverse_bits",
    "verify_merkle_proof_to_cap",
]
