"""Host mirror of the prover steps around the commitments: the permutation argument.

plonky2 @ f99ed9c, plonky2/src/plonk/prover.rs:
    all_wires_permutation_partial_products(witness, betas, gammas, prover_data, common_data)
    wires_permutation_partial_products_and_zs(witness, beta, gamma, prover_data, common_data)
reached from the reference through every prove() (/root/reference/src/transaction/circuits/mod.rs:453,
src/zkdsa/circuits/mod.rs:326, src/rollup/circuits/mod.rs:1247).  prove() then moves every Z to the front of the batch
(`zs_partial_products = [plonk_z_vecs, partial_products.concat()].concat()`) and commits it with
PolynomialBatch::from_values; `zs_partial_products` below returns that batch directly, and on the device path it feeds
the commitment without leaving HBM.  Same names and argument meaning as plonky2; no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from .plonky2 import Context, P, PolynomialBatch, _p, _u64, coset_ifft_batch, default_context, log2_strict

MULTIPLICATIVE_GROUP_GENERATOR = 7


def get_unique_coset_shifts(num_shifts: int) -> np.ndarray:
    """CommonCircuitData::k_is: k_j = g^j (plonky2/src/plonk/permutation_argument.rs / plonk_common.rs)"""
    out, x = [], 1
    for _ in range(num_shifts):
        out.append(x)
        x = x * MULTIPLICATIVE_GROUP_GENERATOR % P
    return np.array(out, dtype=np.uint64)


def num_partial_products(num_routed_wires: int, quotient_degree_factor: int) -> int:
    """CommonCircuitData::num_partial_products = ceil(num_routed_wires / quotient_degree_factor) - 1"""
    return -(-num_routed_wires // quotient_degree_factor) - 1


def zs_partial_products(wires, sigmas, k_is, betas: Sequence[int], gammas: Sequence[int], quotient_degree_factor: int,
                        ctx: Optional[Context] = None) -> np.ndarray:
    """The `Z + partial products` batch of prove(), ready for PolynomialBatch.from_values.

    wires, sigmas: (num_routed_wires, n) — the routed wire columns of the witness and the sigma polynomials' values
    on the subgroup; k_is: (num_routed_wires,); betas, gammas: one per challenge.
    Returns (num_challenges * (1 + num_partial_products), n): row c = Z of challenge c, then the partial products of
    challenge 0, of challenge 1, ..."""
    ctx = ctx or default_context()
    w, s = _u64(wires), _u64(sigmas)
    if w.ndim != 2 or w.shape != s.shape or w.shape[0] == 0:
        raise ValueError("wires and sigmas must be equal-shaped (num_routed_wires, n) arrays")
    R, n = w.shape
    n_log = log2_strict(n)
    k = _u64(k_is)
    b, g = _u64(np.asarray(betas, dtype=np.uint64)), _u64(np.asarray(gammas, dtype=np.uint64))
    if k.shape != (R,) or b.shape != g.shape or b.ndim != 1 or b.size == 0:
        raise ValueError("k_is must have one entry per routed wire, betas / gammas one per challenge")
    chunks = -(-R // quotient_degree_factor)
    out = np.empty((b.size * chunks, n), dtype=np.uint64)
    ctx.check(ctx._lib.b200zkp_partial_products_and_zs(ctx._h, _p(w), _p(s), n_log, R, quotient_degree_factor, _p(k), _p(b), _p(g),
                                                       b.size, _p(out)))
    return out


def wires_permutation_partial_products_and_zs(wires, sigmas, k_is, beta: int, gamma: int, quotient_degree_factor: int,
                                              ctx: Optional[Context] = None) -> np.ndarray:
    """plonky2's per-challenge form: rows [pp_0 .. pp_{num_prods-1}, Z] (Z last, as prover.rs returns it)."""
    cols = zs_partial_products(wires, sigmas, k_is, [beta], [gamma], quotient_degree_factor, ctx)
    return np.concatenate([cols[1:], cols[:1]], axis=0)


def all_wires_permutation_partial_products(wires, sigmas, k_is, betas, gammas, quotient_degree_factor: int,
                                           ctx: Optional[Context] = None):
    """one wires_permutation_partial_products_and_zs result per challenge (one launch sequence for all of them)"""
    cols = zs_partial_products(wires, sigmas, k_is, betas, gammas, quotient_degree_factor, ctx)
    Cn = len(betas)
    num_prods = cols.shape[0] // Cn - 1
    return [np.concatenate([cols[Cn + c * num_prods:Cn + (c + 1) * num_prods], cols[c:c + 1]], axis=0) for c in range(Cn)]


def zs_partial_products_device(ctx: Context, wires, sigmas, k_is, betas, gammas, quotient_degree_factor: int, out=None):
    """Device-resident form on torch int64 tensors (uint64 bit patterns): wires / sigmas (num_routed, n) CUDA tensors
    (rows may be a view of a wider witness matrix as long as each row is contiguous); returns the (C * chunks, n) batch
    in HBM, ready for device.commit_device."""
    import torch
    R, n = wires.shape
    n_log = log2_strict(n)
    assert wires.is_cuda and sigmas.is_cuda and wires.stride(1) == 1 and sigmas.stride(1) == 1 and sigmas.shape == wires.shape
    k = _u64(k_is)
    b, g = _u64(np.asarray(betas, dtype=np.uint64)), _u64(np.asarray(gammas, dtype=np.uint64))
    chunks = -(-R // quotient_degree_factor)
    if out is None:
        out = torch.empty((b.size * chunks, n), dtype=torch.int64, device=wires.device)
    ctx.check(ctx._lib.b200zkp_dev_partial_products_and_zs(ctx._h, C.c_void_p(wires.data_ptr()), wires.stride(0),
                                                           C.c_void_p(sigmas.data_ptr()), sigmas.stride(0), n_log, R,
                                                           quotient_degree_factor, _p(k), _p(b), _p(g), b.size,
                                                           C.c_void_p(out.data_ptr()), out.stride(0)))
    return out


class QuotientError(ValueError):
    """plonky2's `trim_to_len` panic: the vanishing polynomial is not divisible by Z_H (unsatisfied witness)."""


def quotient_poly_chunks(quotient_values, degree_bits: int, quotient_degree_factor: Optional[int] = None,
                         ctx: Optional[Context] = None) -> np.ndarray:
    """Tail of plonky2's compute_quotient_polys + the chunking in prove() (row N1c):

        quotient_values.map(|v| v.coset_ifft(F::coset_shift()))            // one polynomial per challenge
        quotient_poly.trim_to_len(quotient_degree); quotient_poly.chunks(degree)

    quotient_values: (num_challenges, n * 2^quotient_degree_bits) values on the coset 7 <w>, natural order.
    quotient_degree_factor (CommonCircuitData; chosen in min..=max, not necessarily a power of two; default: every chunk):
    coefficients from quotient_degree_factor * n on must vanish — plonky2 panics there, this raises QuotientError — and only
    quotient_degree_factor chunks per challenge are committed.
    Returns (num_challenges * quotient_degree_factor, n) coefficient chunks, the input of PolynomialBatch.from_coeffs."""
    q = _u64(quotient_values)
    if q.ndim != 2:
        raise ValueError("expected (num_challenges, n << quotient_degree_bits)")
    n = 1 << degree_bits
    if q.shape[1] % n or q.shape[1] < n:
        raise ValueError("quotient length must be a multiple of the degree")
    coeffs = coset_ifft_batch(q, MULTIPLICATIVE_GROUP_GENERATOR, ctx)
    total = q.shape[1] // n
    qdf = total if quotient_degree_factor is None else int(quotient_degree_factor)
    if not 0 < qdf <= total:
        raise ValueError("quotient_degree_factor out of range")
    if coeffs[:, qdf * n:].any():
        raise QuotientError("Quotient has failed: coefficients beyond quotient_degree_factor * n are not zero")
    return np.ascontiguousarray(coeffs[:, :qdf * n]).reshape(q.shape[0] * qdf, n)


def commit_quotient(quotient_values, degree_bits: int, rate_bits: int, blinding: bool, cap_height: int, salt=None,
                    ctx: Optional[Context] = None, quotient_degree_factor: Optional[int] = None) -> PolynomialBatch:
    """quotient_polys_commitment of prove(): PolynomialBatch::from_coeffs(all_quotient_poly_chunks, ...)"""
    return PolynomialBatch.from_coeffs(quotient_poly_chunks(quotient_values, degree_bits, quotient_degree_factor, ctx), rate_bits,
                                       blinding, cap_height, salt=salt, ctx=ctx)


def commit_quotient_device(ctx: Context, quotient_values, degree_bits: int, rate_bits: int, cap_height: int,
                           quotient_degree_factor: Optional[int] = None):
    """device-resident form: (C, n << q) CUDA tensor of quotient values -> DeviceCommitment of the C * quotient_degree_factor
    chunks (default 2^q); the coefficients never leave HBM (the divisibility check reads one flag back)"""
    import torch
    from .device import commit_device
    Cn, total = quotient_values.shape
    n = 1 << degree_bits
    assert quotient_values.is_cuda and quotient_values.is_contiguous() and total % n == 0
    coeffs = torch.empty_like(quotient_values)
    scratch = torch.empty_like(quotient_values)
    ctx.check(ctx._lib.b200zkp_dev_coset_intt(ctx._h, C.c_void_p(quotient_values.data_ptr()), total, C.c_void_p(coeffs.data_ptr()), total,
                                              C.c_void_p(scratch.data_ptr()), log2_strict(total), Cn, MULTIPLICATIVE_GROUP_GENERATOR))
    qdf = total // n if quotient_degree_factor is None else int(quotient_degree_factor)
    if not 0 < qdf <= total // n:
        raise ValueError("quotient_degree_factor out of range")
    if qdf < total // n:
        ctx.synchronize()       # (the ctx stream may not be torch's current stream)
        if bool(coeffs[:, qdf * n:].any()):
            raise QuotientError("Quotient has failed: coefficients beyond quotient_degree_factor * n are not zero")
        coeffs = coeffs[:, :qdf * n].contiguous()
    return commit_device(ctx, coeffs.view(Cn * qdf, n), rate_bits, cap_height, is_coeffs=True)


# ------------------------------------------------------------------------------------------------ row N1b: compute_quotient_polys
GATE_NOOP, GATE_CONSTANT, GATE_PUBLIC_INPUT, GATE_ARITHMETIC, GATE_POSEIDON = range(5)
(GATE_ARITHMETIC_EXTENSION, GATE_MUL_EXTENSION, GATE_BASE_SUM, GATE_REDUCING, GATE_REDUCING_EXTENSION, GATE_RANDOM_ACCESS,
 GATE_EXPONENTIATION, GATE_POSEIDON_MDS) = range(5, 13)
MAX_GATES = 16


class _VanishingDesc(C.Structure):
    _fields_ = [("degree_bits", C.c_uint32), ("quotient_degree_bits", C.c_uint32), ("quotient_degree_factor", C.c_uint32),
                ("num_routed_wires", C.c_uint32), ("num_challenges", C.c_uint32), ("num_selectors", C.c_uint32),
                ("n_gates", C.c_uint32),
                ("gate_kind", C.c_uint32 * MAX_GATES), ("gate_selector_index", C.c_uint32 * MAX_GATES),
                ("gate_group_begin", C.c_uint32 * MAX_GATES), ("gate_group_end", C.c_uint32 * MAX_GATES),
                ("gate_params", (C.c_uint32 * 3) * MAX_GATES),
                ("k_is", C.c_void_p), ("betas", C.c_void_p), ("gammas", C.c_void_p), ("alphas", C.c_void_p),
                ("public_inputs_hash", C.c_uint64 * 4)]


class CommonCircuitData:
    """What compute_quotient_polys needs of plonky2's CommonCircuitData: degree_bits, the gate list in CircuitBuilder's order with
    the selector polynomial and selector group of every gate (SelectorsInfo), quotient_degree_factor, k_is.
    gates: [(kind, selector_index, (group_begin, group_end)[, params]), ...]; params: the gate's own parameters
    (BaseSum (B, num_limbs), Reducing / ReducingExtension (num_coeffs,), RandomAccess (bits, num_copies, num_extra_constants),
    Exponentiation (num_power_bits,))"""

    def __init__(self, degree_bits: int, gates, num_selectors: int, quotient_degree_factor: int = 8, num_routed_wires: int = 80,
                 k_is=None):
        self.degree_bits, self.gates, self.num_selectors = degree_bits, list(gates), num_selectors
        self.quotient_degree_factor, self.num_routed_wires = quotient_degree_factor, num_routed_wires
        self.k_is = _u64(k_is) if k_is is not None else get_unique_coset_shifts(num_routed_wires)
        if len(self.gates) > MAX_GATES:
            raise ValueError("too many gate types")

    @property
    def quotient_degree_bits(self) -> int:
        return max(self.quotient_degree_factor - 1, 0).bit_length()        # log2_ceil


def compute_quotient_values_device(ctx: Context, common: CommonCircuitData, constants_sigmas_lde, wires_lde, zs_partial_products_lde,
                                   betas, gammas, alphas, public_inputs_hash, out=None):
    """compute_quotient_polys up to its coset_ifft, on the three LDEs where the commitments left them: (columns, N) CUDA tensors
    in leaf order (DeviceCommitment.lde).  Returns the (num_challenges, n << quotient_degree_bits) quotient values in natural order
    on the coset 7 <w> — the input of commit_quotient_device."""
    import torch
    q_bits = common.quotient_degree_bits
    Q = 1 << (common.degree_bits + q_bits)
    b, g, a = (_u64(np.asarray(v, dtype=np.uint64)) for v in (betas, gammas, alphas))
    if not (b.shape == g.shape == a.shape) or b.ndim != 1:
        raise ValueError("one beta, gamma and alpha per challenge")
    for t in (constants_sigmas_lde, wires_lde, zs_partial_products_lde):
        if not t.is_cuda or t.stride(1) != 1 or t.shape[1] < Q:
            raise ValueError("LDEs must be CUDA tensors of at least n << quotient_degree_bits leaves per column")
    chunks = -(-common.num_routed_wires // common.quotient_degree_factor)
    if constants_sigmas_lde.shape[0] != common.num_selectors + 2 + common.num_routed_wires or wires_lde.shape[0] < 135 \
            or zs_partial_products_lde.shape[0] != b.size * chunks:
        raise ValueError("batch widths do not match the circuit description")
    d = _VanishingDesc()
    d.degree_bits, d.quotient_degree_bits, d.quotient_degree_factor = common.degree_bits, q_bits, common.quotient_degree_factor
    d.num_routed_wires, d.num_challenges, d.num_selectors, d.n_gates = common.num_routed_wires, b.size, common.num_selectors, len(common.gates)
    for i, gate in enumerate(common.gates):
        kind, sel, grp = gate[:3]
        d.gate_kind[i], d.gate_selector_index[i], d.gate_group_begin[i], d.gate_group_end[i] = kind, sel, grp[0], grp[1]
        for j, v in enumerate(gate[3] if len(gate) > 3 else ()):
            d.gate_params[i][j] = int(v)
    k = _u64(common.k_is)
    d.k_is, d.betas, d.gammas, d.alphas = k.ctypes.data, b.ctypes.data, g.ctypes.data, a.ctypes.data
    for i, v in enumerate(public_inputs_hash):
        d.public_inputs_hash[i] = int(v)
    if out is None:
        out = torch.empty((b.size, Q), dtype=torch.int64, device=wires_lde.device)
    ctx.check(ctx._lib.b200zkp_dev_quotient_values(ctx._h, C.byref(d), C.c_void_p(constants_sigmas_lde.data_ptr()), constants_sigmas_lde.stride(0),
                                                   C.c_void_p(wires_lde.data_ptr()), wires_lde.stride(0),
                                                   C.c_void_p(zs_partial_products_lde.data_ptr()), zs_partial_products_lde.stride(0),
                                                   C.c_void_p(out.data_ptr()), out.stride(0)))
    return out
