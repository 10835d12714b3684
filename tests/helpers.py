"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

P = 0xFFFFFFFF00000001


def hex_to_elements(s: str) -> np.ndarray:
    """The reference's WrappedHashOut hex form -> 4 u64 (32 bytes LE, byte-reversed in the string;
    /root/reference/src/sparse_merkle_tree/goldilocks_poseidon/hash/mod.rs:84-119)."""
    raw = bytes.fromhex(s[2:])[::-1]
    return np.frombuffer(raw, dtype="<u8").astype(np.uint64)


def bitrev(x: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def bitrev_perm(bits: int) -> np.ndarray:
    n = 1 << bits
    idx = np.arange(n, dtype=np.uint64)
    out = np.zeros(n, dtype=np.uint64)
    for b in range(bits):
        out |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(bits - 1 - b)
    return out.astype(np.int64)


def rand_field(rng, shape, canonical=True) -> np.ndarray:
    v = rng.integers(0, 2**64, size=shape, dtype=np.uint64)
    if canonical:
        v = np.where(v >= np.uint64(P), v - np.uint64(P), v)
    return v


def hostile_columns(n: int) -> np.ndarray:
    """All-zero, all p-1, non-canonical 2^64-1, single non-zero, p itself (== 0), alternating extremes."""
    cols = [np.zeros(n, np.uint64), np.full(n, P - 1, np.uint64), np.full(n, 2**64 - 1, np.uint64)]
    one = np.zeros(n, np.uint64); one[n // 2] = 1; cols.append(one)
    cols.append(np.full(n, P, np.uint64))
    alt = np.zeros(n, np.uint64); alt[::2] = np.uint64(2**64 - 1); alt[1::2] = np.uint64(P - 1); cols.append(alt)
    return np.stack(cols)
