"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end of oracle/liboracle.so (oracle.c: the single-threaded C restatement of plonky2's
commitment path) and oracle/libcpubaseline.so (cpu_baseline.c: the multithreaded plonky2-shaped CPU
baseline).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under intmax_zkp_core_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
P = 0xFFFFFFFF00000001
_u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> None:
    """Compile the oracle with gcc (recipe: oracle/Makefile)."""
    targets = ["liboracle.so", "libcpubaseline.so"]
    if force or any(not os.path.exists(os.path.join(_HERE, t)) for t in targets):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


def _ptr(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


_lib = None
_blib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        _lib.orc_eval_at_lde_point.restype = C.c_uint64
        _lib.orc_eval_at_lde_point.argtypes = [_u64p, C.c_uint, C.c_uint, C.c_uint64]
        _lib.orc_merkle_new.argtypes = [_u64p, C.c_uint64, C.c_uint32, C.c_uint32, _u64p, _u64p]
        _lib.orc_merkle_prove.argtypes = [_u64p, C.c_uint64, C.c_uint32, C.c_uint64, _u64p]
        _lib.orc_merkle_verify.argtypes = [_u64p, C.c_uint32, C.c_uint64, _u64p, C.c_uint32, _u64p]
        _lib.orc_commit.argtypes = [_u64p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                    _u64p, _u64p, _u64p, _u64p, _u64p]
        _lib.orc_hash_no_pad.argtypes = [_u64p, C.c_size_t, _u64p]
        _lib.orc_hash_pad.argtypes = [_u64p, C.c_size_t, _u64p]
        _lib.orc_hash_or_noop.argtypes = [_u64p, C.c_size_t, _u64p]
        _lib.orc_two_to_one.argtypes = [_u64p, _u64p, _u64p]
        _lib.orc_permute.argtypes = [_u64p]
        _lib.orc_dft.argtypes = [_u64p, _u64p, C.c_uint]
        _lib.orc_fft.argtypes = [_u64p, C.c_uint]
        _lib.orc_ifft.argtypes = [_u64p, C.c_uint]
        _lib.orc_coset_lde.argtypes = [_u64p, C.c_uint, C.c_uint, _u64p]
        _lib.orc_round_constants.argtypes = [_u64p]
        _lib.orc_eval_ext2.argtypes = [_u64p, C.c_uint64, _u64p, _u64p]
    return _lib


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint64))


def round_constants() -> np.ndarray:
    out = np.zeros(360, dtype=np.uint64)
    lib().orc_round_constants(_ptr(out))
    return out


def permute(state) -> np.ndarray:
    s = _u64(state).copy()
    assert s.shape == (12,)
    lib().orc_permute(_ptr(s))
    return s


def permute_many(states) -> np.ndarray:
    s = _u64(states).copy().reshape(-1, 12)
    L = lib()
    for i in range(s.shape[0]):
        L.orc_permute(s[i].ctypes.data_as(_u64p))
    return s


def hash_no_pad(x) -> np.ndarray:
    x = _u64(x); out = np.zeros(4, dtype=np.uint64)
    lib().orc_hash_no_pad(_ptr(x), x.size, _ptr(out))
    return out


def hash_pad(x) -> np.ndarray:
    x = _u64(x); out = np.zeros(4, dtype=np.uint64)
    lib().orc_hash_pad(_ptr(x), x.size, _ptr(out))
    return out


def hash_or_noop(x) -> np.ndarray:
    x = _u64(x); out = np.zeros(4, dtype=np.uint64)
    lib().orc_hash_or_noop(_ptr(x), x.size, _ptr(out))
    return out


def two_to_one(l, r) -> np.ndarray:
    l, r = _u64(l), _u64(r); out = np.zeros(4, dtype=np.uint64)
    lib().orc_two_to_one(_ptr(l), _ptr(r), _ptr(out))
    return out


def dft(x) -> np.ndarray:
    x = _u64(x); n_log = int(x.size).bit_length() - 1
    out = np.zeros_like(x)
    lib().orc_dft(_ptr(x), _ptr(out), n_log)
    return out


def fft(x) -> np.ndarray:
    x = _u64(x).copy(); n_log = int(x.size).bit_length() - 1
    lib().orc_fft(_ptr(x), n_log)
    return x


def ifft(x) -> np.ndarray:
    x = _u64(x).copy(); n_log = int(x.size).bit_length() - 1
    lib().orc_ifft(_ptr(x), n_log)
    return x


def coset_lde(coeffs, rate_bits: int) -> np.ndarray:
    c = _u64(coeffs); n_log = int(c.size).bit_length() - 1
    out = np.zeros(c.size << rate_bits, dtype=np.uint64)
    lib().orc_coset_lde(_ptr(c), n_log, rate_bits, _ptr(out))
    return out


def eval_at_lde_point(coeffs, rate_bits: int, i: int) -> int:
    c = _u64(coeffs); n_log = int(c.size).bit_length() - 1
    return int(lib().orc_eval_at_lde_point(_ptr(c), n_log, rate_bits, i))


def eval_ext2(coeffs, zeta) -> np.ndarray:
    """p(zeta) in F_p[X]/(X^2 - 7); coeffs (n,), zeta (2,) -> (2,)."""
    c, z = _u64(coeffs), _u64(zeta)
    out = np.zeros(2, dtype=np.uint64)
    lib().orc_eval_ext2(_ptr(c), c.size, _ptr(z), _ptr(out))
    return out


def merkle_new(leaves, cap_height: int):
    """leaves: (N, leaf_len) uint64.  Returns (digests (2(N-2^h),4), cap (2^h,4))."""
    lv = _u64(leaves)
    N, L = lv.shape
    dig = np.zeros((max(2 * (N - (1 << cap_height)), 0), 4), dtype=np.uint64)
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    rc = lib().orc_merkle_new(_ptr(lv), N, L, cap_height, _ptr(dig) if dig.size else None, _ptr(cap))
    if rc != 0:
        raise ValueError("orc_merkle_new: bad arguments")
    return dig, cap


def merkle_prove(digests, n_leaves: int, cap_height: int, leaf_index: int) -> np.ndarray:
    n_sib = (n_leaves.bit_length() - 1) - cap_height
    sib = np.zeros((n_sib, 4), dtype=np.uint64)
    d = _u64(digests)
    if n_sib:
        lib().orc_merkle_prove(_ptr(d), n_leaves, cap_height, leaf_index, _ptr(sib))
    return sib


def merkle_verify(leaf, leaf_index: int, siblings, cap) -> bool:
    leaf, sib, cap = _u64(leaf), _u64(siblings).reshape(-1, 4), _u64(cap)
    return bool(lib().orc_merkle_verify(_ptr(leaf), leaf.size, leaf_index,
                                        _ptr(sib) if sib.size else None, sib.shape[0], _ptr(cap)))


def commit(inp, rate_bits: int, cap_height: int, is_coeffs: bool = False, salt=None):
    """inp: (k, n) uint64 column-major (one row of this array per polynomial).

    Returns dict(coeffs (k,n), leaves (N,k+salt), digests (2(N-2^h),4), cap (2^h,4)).
    """
    x = _u64(inp)
    k, n = x.shape
    n_log = n.bit_length() - 1
    assert 1 << n_log == n
    N = n << rate_bits
    row = k + (4 if salt is not None else 0)
    coeffs = np.zeros((k, n), dtype=np.uint64)
    leaves = np.zeros((N, row), dtype=np.uint64)
    dig = np.zeros((max(2 * (N - (1 << cap_height)), 0), 4), dtype=np.uint64)
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    s = None
    if salt is not None:
        s = _u64(salt)
        assert s.shape == (4, N)
    rc = lib().orc_commit(_ptr(x), int(is_coeffs), n_log, k, rate_bits, cap_height,
                          _ptr(s) if s is not None else None, _ptr(coeffs), _ptr(leaves),
                          _ptr(dig) if dig.size else None, _ptr(cap))
    if rc != 0:
        raise ValueError(f"orc_commit failed rc={rc}")
    return dict(coeffs=coeffs, leaves=leaves, digests=dig, cap=cap)


# ------------------------------------------------------------------ synthetic inputs
def splitmix64(j: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (j.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synthetic_values(k: int, n: int, seed: int = 0) -> np.ndarray:
    """v[c][i] = splitmix64(seed*2^48 + c*n + i) mod p, shape (k, n)."""
    idx = np.arange(k * n, dtype=np.uint64) + np.uint64(seed << 48)
    v = splitmix64(idx)
    v = np.where(v >= np.uint64(P), v - np.uint64(P), v)
    return v.reshape(k, n)


# ------------------------------------------------------------------ multithreaded CPU baseline
def baseline_lib():
    global _blib
    if _blib is None:
        build()
        _blib = C.CDLL(os.path.join(_HERE, "libcpubaseline.so"))
        _blib.cpub_commit.argtypes = [_u64p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                      _u64p, _u64p, _u64p, _u64p, C.POINTER(C.c_double)]
        _blib.cpub_commit.restype = C.c_int
        _blib.cpub_threads.restype = C.c_int
        _blib.cpub_set_threads.argtypes = [C.c_int]
        _blib.cpub_permute.argtypes = [_u64p]
        _blib.cpub_leaf_hash_rows.argtypes = [_u64p, C.c_uint64, C.c_uint32, _u64p]
        _blib.cpub_leaf_hash_rows.restype = C.c_double
    return _blib


def baseline_commit(inp, rate_bits: int, cap_height: int, is_coeffs: bool = False):
    """Multithreaded plonky2-shaped CPU commit.  Returns (dict like commit(), stage seconds[5])."""
    x = _u64(inp)
    k, n = x.shape
    n_log = n.bit_length() - 1
    N = n << rate_bits
    coeffs = np.zeros((k, n), dtype=np.uint64)
    leaves = np.zeros((N, k), dtype=np.uint64)
    dig = np.zeros((max(2 * (N - (1 << cap_height)), 0), 4), dtype=np.uint64)
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    t = (C.c_double * 5)()
    rc = baseline_lib().cpub_commit(_ptr(x), int(is_coeffs), n_log, k, rate_bits, cap_height,
                                    _ptr(coeffs), _ptr(leaves), _ptr(dig) if dig.size else None,
                                    _ptr(cap), t)
    if rc != 0:
        raise ValueError(f"cpub_commit failed rc={rc}")
    return dict(coeffs=coeffs, leaves=leaves, digests=dig, cap=cap), list(t)


def baseline_threads() -> int:
    return int(baseline_lib().cpub_threads())


def baseline_set_threads(n: int) -> int:
    """OpenMP team size of the CPU baseline (torchrun exports OMP_NUM_THREADS=1 to its workers); returns the new size"""
    baseline_lib().cpub_set_threads(int(n))
    return baseline_threads()
