"""ORACLE — TEST INFRASTRUCTURE ONLY (row N1b of SURVEY.md section 8f).

Pure-Python restatement of the step of plonky2's prover that consumes the three LDEs: evaluating the vanishing polynomial on
the quotient coset and dividing by Z_H,

    plonky2 @ f99ed9c  plonky2/src/plonk/prover.rs          compute_quotient_polys
                       plonky2/src/plonk/vanishing_poly.rs  eval_vanishing_poly_base_batch, evaluate_gate_constraints_base_batch
                       plonky2/src/plonk/plonk_common.rs    ZeroPolyOnCoset, eval_l_1, reduce_with_powers
                       plonky2/src/plonk/vars.rs, gates/selectors.rs (selector polynomials, compute_filter)
                       plonky2/src/gates/{noop, constant, public_input, arithmetic_base, poseidon, poseidon_mds, base_sum,
                                          arithmetic_extension, multiplication_extension, reducing, reducing_extension,
                                          random_access, exponentiation}.rs

reached from the reference through every prove() (/root/reference/src/rollup/circuits/mod.rs:1247,
src/transaction/circuits/mod.rs:453, src/zkdsa/circuits/mod.rs:326); the Poseidon gate is the one the reference's circuits
instantiate most (/root/reference/src/poseidon/gadgets/mod.rs:7-22).  The crate source is not on this machine: this follows the
published algorithm and wire layouts from memory — **parity unpinned** (no fixture of the reference holds a quotient polynomial).
What pins it here is the verifier's own identity: for a witness that satisfies the gates and the copy constraints,
    quotient(zeta) * Z_H(zeta) == sum_t alpha^t * term_t(zeta)
at a random point zeta outside the domain, where the right-hand side is recomputed from polynomial evaluations at zeta and
g * zeta (`check_quotient_identity`), and it fails as soon as one wire is changed.

Vanishing terms, in plonky2's order (C challenges, R routed wires, chunks of `degree` = quotient_degree_factor wires):
    [ L_1(x) (Z_c(x) - 1)                                            for c < C ]
    [ acc_{c,l}(x) prod_{j in chunk l} num_{c,j} - acc_{c,l+1}(x) prod_{j in chunk l} den_{c,j}   for c < C, l < chunks ]
        acc_c = [Z_c(x), pp_{c,0}, .., pp_{c,chunks-2}, Z_c(g x)],  num = w_j + beta_c k_j x + gamma_c,  den = w_j + beta_c sigma_j + gamma_c
    [ sum over gate types of filter_g(x) * constraint_{g,t}(x)       for t < num_gate_constraints ]
and quotient_c(x) = (sum_t alpha_c^t term_t(x)) / Z_H(x) on the coset 7 <w_(n 2^q)>.
"""
from __future__ import annotations

import os
import random
import sys
from typing import Dict, List, Sequence

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.path.join(ROOT, "tools") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tools"))
import poseidon_derive as PD   # round constants / MDS / fast partial-round tables re-derived from their published recipe

P = 0xFFFFFFFF00000001
GENERATOR = 7
POWER_OF_TWO_GENERATOR = 1753635133440165772
UNUSED_SELECTOR = 0xFFFFFFFF

WIDTH, N_FULL_HALF, N_PARTIAL = 12, 4, 22
NUM_WIRES, NUM_ROUTED, NUM_GATE_CONSTANTS = 135, 80, 2       # CircuitConfig::standard_recursion_config()

# gate kinds (ids shared with csrc/vanishing_kernels.cuh).  Gates with parameters carry them as a tuple:
#   BASE_SUM (B, num_limbs) | REDUCING (num_coeffs,) | REDUCING_EXTENSION (num_coeffs,) |
#   RANDOM_ACCESS (bits, num_copies, num_extra_constants) | EXPONENTIATION (num_power_bits,)
NOOP, CONSTANT, PUBLIC_INPUT, ARITHMETIC, POSEIDON = 0, 1, 2, 3, 4
ARITHMETIC_EXTENSION, MUL_EXTENSION, BASE_SUM, REDUCING, REDUCING_EXTENSION, RANDOM_ACCESS, EXPONENTIATION, POSEIDON_MDS = 5, 6, 7, 8, 9, 10, 11, 12
D = 2                      # quadratic extension F[X] / (X^2 - 7)
W_EXT = 7


def gate_degree(kind: int, params=()) -> int:
    """Gate::degree()"""
    fixed = {NOOP: 0, CONSTANT: 1, PUBLIC_INPUT: 1, ARITHMETIC: 3, POSEIDON: 7, ARITHMETIC_EXTENSION: 3, MUL_EXTENSION: 3,
             REDUCING: 2, REDUCING_EXTENSION: 2, EXPONENTIATION: 4, POSEIDON_MDS: 1}
    if kind == BASE_SUM:
        return params[0]
    if kind == RANDOM_ACCESS:
        return params[0] + 1
    return fixed[kind]


def gate_num_constraints(kind: int, params=()) -> int:
    """Gate::num_constraints() for standard_recursion_config (135 wires, 80 routed)"""
    fixed = {NOOP: 0, CONSTANT: NUM_GATE_CONSTANTS, PUBLIC_INPUT: 4, ARITHMETIC: NUM_ROUTED // 4, POSEIDON: 123,
             ARITHMETIC_EXTENSION: (NUM_ROUTED // (4 * D)) * D, MUL_EXTENSION: (NUM_ROUTED // (3 * D)) * D, POSEIDON_MDS: WIDTH * D}
    if kind == BASE_SUM:
        return 1 + params[1]
    if kind in (REDUCING, REDUCING_EXTENSION):
        return D * params[0]
    if kind == RANDOM_ACCESS:
        return params[1] * (params[0] + 2) + params[2]
    if kind == EXPONENTIATION:
        return params[0] + 1
    return fixed[kind]


GATE_DEGREE = {k: gate_degree(k) for k in (NOOP, CONSTANT, PUBLIC_INPUT, ARITHMETIC, POSEIDON)}
GATE_CONSTRAINTS = {k: gate_num_constraints(k) for k in (NOOP, CONSTANT, PUBLIC_INPUT, ARITHMETIC, POSEIDON)}


def ext_mul(a, b):
    return ((a[0] * b[0] + W_EXT * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def ext_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def ext_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def ext_scale(a, c):
    return (a[0] * c % P, a[1] * c % P)

# PoseidonGate wire layout (gates/poseidon.rs)
W_IN, W_OUT, W_SWAP, W_DELTA, W_FULL0, W_PARTIAL, W_FULL1 = 0, 12, 24, 25, 29, 65, 87

_RC = PD.round_constants()
_M = PD.mds_matrix()
_FAST = PD.fast_partial_tables(_RC, _M)


def root(n_log: int) -> int:
    return pow(POWER_OF_TWO_GENERATOR, 1 << (32 - n_log), P)


def sbox(x: int) -> int:
    return pow(x, 7, P)


def mds(s: Sequence[int]) -> List[int]:
    return [sum(_M[r][i] * s[i] for i in range(WIDTH)) % P for r in range(WIDTH)]


def fast_tables() -> Dict[str, list]:
    return _FAST


# ---------------------------------------------------------------------------------------------------- gate constraints
def poseidon_gate_constraints(wires: Sequence[int], out: List[int]) -> None:
    """gates/poseidon.rs eval_unfiltered; appends the 123 constraints to `out`.  `wires`: the 135 wire values at one point."""
    swap = wires[W_SWAP]
    out.append(swap * (swap - 1) % P)
    state = [0] * WIDTH
    for i in range(4):
        lhs, rhs, delta = wires[W_IN + i], wires[W_IN + 4 + i], wires[W_DELTA + i]
        out.append((swap * (rhs - lhs) - delta) % P)
        state[i] = (lhs + delta) % P
        state[i + 4] = (rhs - delta) % P
    for i in range(8, 12):
        state[i] = wires[W_IN + i]
    rnd = 0
    for r in range(N_FULL_HALF):
        state = [(x + _RC[rnd * WIDTH + i]) % P for i, x in enumerate(state)]
        if r:
            for i in range(WIDTH):
                s_in = wires[W_FULL0 + WIDTH * (r - 1) + i]
                out.append((state[i] - s_in) % P)
                state[i] = s_in
        state = mds([sbox(x) for x in state])
        rnd += 1
    state = [(x + y) % P for x, y in zip(state, _FAST["first_vec"])]
    state = [state[0]] + [sum(a * b for a, b in zip(row, state[1:])) % P for row in _FAST["init_mat"]]
    for r in range(N_PARTIAL):
        s_in = wires[W_PARTIAL + r]
        out.append((state[0] - s_in) % P)
        s0 = (sbox(s_in) + _FAST["post"][r]) % P
        new0 = (25 * s0 + sum(a * b for a, b in zip(_FAST["vhat"][r], state[1:]))) % P
        state = [new0] + [(x + w * s0) % P for x, w in zip(state[1:], _FAST["what"][r])]
    rnd += N_PARTIAL
    for r in range(N_FULL_HALF):
        state = [(x + _RC[rnd * WIDTH + i]) % P for i, x in enumerate(state)]
        for i in range(WIDTH):
            s_in = wires[W_FULL1 + WIDTH * r + i]
            out.append((state[i] - s_in) % P)
            state[i] = s_in
        state = mds([sbox(x) for x in state])
        rnd += 1
    for i in range(WIDTH):
        out.append((state[i] - wires[W_OUT + i]) % P)


def poseidon_gate_witness(inputs: Sequence[int], swap: int = 0) -> List[int]:
    """the 135 wires of one PoseidonGate row whose constraints all vanish (what the gate's generator computes)"""
    w = [0] * NUM_WIRES
    for i in range(WIDTH):
        w[W_IN + i] = inputs[i] % P
    w[W_SWAP] = swap
    state = [0] * WIDTH
    for i in range(4):
        lhs, rhs = w[W_IN + i], w[W_IN + 4 + i]
        delta = swap * (rhs - lhs) % P
        w[W_DELTA + i] = delta
        state[i] = (lhs + delta) % P
        state[i + 4] = (rhs - delta) % P
    for i in range(8, 12):
        state[i] = w[W_IN + i]
    rnd = 0
    for r in range(N_FULL_HALF):
        state = [(x + _RC[rnd * WIDTH + i]) % P for i, x in enumerate(state)]
        if r:
            for i in range(WIDTH):
                w[W_FULL0 + WIDTH * (r - 1) + i] = state[i]
        state = mds([sbox(x) for x in state])
        rnd += 1
    state = [(x + y) % P for x, y in zip(state, _FAST["first_vec"])]
    state = [state[0]] + [sum(a * b for a, b in zip(row, state[1:])) % P for row in _FAST["init_mat"]]
    for r in range(N_PARTIAL):
        w[W_PARTIAL + r] = state[0]
        s0 = (sbox(state[0]) + _FAST["post"][r]) % P
        new0 = (25 * s0 + sum(a * b for a, b in zip(_FAST["vhat"][r], state[1:]))) % P
        state = [new0] + [(x + ww * s0) % P for x, ww in zip(state[1:], _FAST["what"][r])]
    rnd += N_PARTIAL
    for r in range(N_FULL_HALF):
        state = [(x + _RC[rnd * WIDTH + i]) % P for i, x in enumerate(state)]
        for i in range(WIDTH):
            w[W_FULL1 + WIDTH * r + i] = state[i]
        state = mds([sbox(x) for x in state])
        rnd += 1
    for i in range(WIDTH):
        w[W_OUT + i] = state[i]
    return w


def random_access_layout(bits: int, num_copies: int, num_extra: int):
    """gates/random_access.rs wire indices: (access_index, claimed_element, list_item(i), bit(i)) per copy, extra constants"""
    vec = 1 << bits
    routed = (2 + vec) * num_copies + num_extra
    return dict(vec=vec, access=lambda c: (2 + vec) * c, claimed=lambda c: (2 + vec) * c + 1,
                item=lambda i, c: (2 + vec) * c + 2 + i, extra=lambda i: (2 + vec) * num_copies + i,
                bit=lambda i, c: routed + c * bits + i)


def gate_constraints(kind: int, wires: Sequence[int], consts: Sequence[int], pi_hash: Sequence[int], params=()) -> List[int]:
    """eval_unfiltered of one gate type at one point; consts = the NUM_GATE_CONSTANTS gate constants (selectors removed)"""
    out: List[int] = []
    ext = lambda j: (wires[j], wires[j + 1])
    if kind == ARITHMETIC_EXTENSION:                       # gates/arithmetic_extension.rs, num_ops = num_routed / (4 D)
        for i in range(NUM_ROUTED // (4 * D)):
            m0, m1, addend, output = (ext(4 * D * i + D * t) for t in range(4))
            computed = ext_add(ext_scale(ext_mul(m0, m1), consts[0]), ext_scale(addend, consts[1]))
            out += list(ext_sub(output, computed))
    elif kind == MUL_EXTENSION:                            # gates/multiplication_extension.rs, num_ops = num_routed / (3 D)
        for i in range(NUM_ROUTED // (3 * D)):
            m0, m1, output = (ext(3 * D * i + D * t) for t in range(3))
            out += list(ext_sub(output, ext_scale(ext_mul(m0, m1), consts[0])))
    elif kind == BASE_SUM:                                 # gates/base_sum.rs: WIRE_SUM = 0, limbs from wire 1, little endian
        base, num_limbs = params
        limbs = wires[1:1 + num_limbs]
        computed = 0
        for l in reversed(limbs):
            computed = (computed * base + l) % P
        out.append((computed - wires[0]) % P)
        for l in limbs:
            prod = 1
            for i in range(base):
                prod = prod * (l - i) % P
            out.append(prod)
    elif kind in (REDUCING, REDUCING_EXTENSION):           # gates/reducing.rs, reducing_extension.rs
        (num_coeffs,) = params
        cw = 1 if kind == REDUCING else D                  # wires per coefficient
        output, alpha, acc = ext(0), ext(D), ext(2 * D)
        start_coeffs = 3 * D
        start_accs = start_coeffs + num_coeffs * cw
        for i in range(num_coeffs):
            coeff = (wires[start_coeffs + i], 0) if kind == REDUCING else ext(start_coeffs + D * i)
            acc_i = output if i == num_coeffs - 1 else ext(start_accs + D * i)
            out += list(ext_sub(ext_add(ext_mul(acc, alpha), coeff), acc_i))
            acc = acc_i
    elif kind == RANDOM_ACCESS:                            # gates/random_access.rs
        bits, num_copies, num_extra = params
        L = random_access_layout(bits, num_copies, num_extra)
        for c in range(num_copies):
            bs = [wires[L["bit"](i, c)] for i in range(bits)]
            for b in bs:
                out.append(b * (b - 1) % P)
            rec = 0
            for b in reversed(bs):
                rec = (2 * rec + b) % P
            out.append((rec - wires[L["access"](c)]) % P)
            items = [wires[L["item"](i, c)] for i in range(L["vec"])]
            for b in bs:
                items = [(items[2 * t] + b * (items[2 * t + 1] - items[2 * t])) % P for t in range(len(items) // 2)]
            out.append((items[0] - wires[L["claimed"](c)]) % P)
        for i in range(num_extra):
            out.append((consts[i] - wires[L["extra"](i)]) % P)
    elif kind == EXPONENTIATION:                           # gates/exponentiation.rs: base 0, power bits 1.., output, intermediates
        (nb,) = params
        base = wires[0]
        bits_ = wires[1:1 + nb]
        output = wires[1 + nb]
        inter = wires[2 + nb:2 + 2 * nb]
        for i in range(nb):
            prev = 1 if i == 0 else inter[i - 1] * inter[i - 1] % P
            cur = bits_[nb - 1 - i]
            out.append((prev * ((cur * base + 1 - cur) % P) - inter[i]) % P)
        out.append((output - inter[nb - 1]) % P)
    elif kind == POSEIDON_MDS:                             # gates/poseidon_mds.rs: the MDS layer on 12 extension elements
        cols = [mds([wires[D * i + comp] for i in range(WIDTH)]) for comp in range(D)]
        for i in range(WIDTH):
            for comp in range(D):
                out.append((wires[D * (WIDTH + i) + comp] - cols[comp][i]) % P)
    elif kind == CONSTANT:                                   # gates/constant.rs: consts_inputs[i] - wire[i]
        for i in range(NUM_GATE_CONSTANTS):
            out.append((consts[i] - wires[i]) % P)
    elif kind == PUBLIC_INPUT:                             # gates/public_input.rs: wire[i] - public_inputs_hash[i]
        for i in range(4):
            out.append((wires[i] - pi_hash[i]) % P)
    elif kind == ARITHMETIC:                               # gates/arithmetic_base.rs, num_ops = num_routed_wires / 4
        for i in range(NUM_ROUTED // 4):
            m0, m1, addend, output = wires[4 * i:4 * i + 4]
            out.append((output - (m0 * m1 % P * consts[0] + addend * consts[1])) % P)
    elif kind == POSEIDON:
        poseidon_gate_constraints(wires, out)
    return out


# ---------------------------------------------------------------------------------------------------- witnesses of the other gates
# parameters of the instances standard_recursion_config circuits hold (Gate::new_from_config; RandomAccess for 16-element lists)
EXT_GATE_PARAMS = {BASE_SUM: (2, 63), REDUCING: (43,), REDUCING_EXTENSION: (32,), RANDOM_ACCESS: (4, 4, 2), EXPONENTIATION: (66,)}


def gate_witness(kind: int, rnd: random.Random, consts: Sequence[int], params=()) -> List[int]:
    """135 wires of one row of `kind` whose constraints vanish (what the gate's generators compute from random inputs)"""
    w = [rnd.randrange(P) for _ in range(NUM_WIRES)]        # unconstrained wires hold anything
    rext = lambda: (rnd.randrange(P), rnd.randrange(P))

    def put(j, e):
        w[j], w[j + 1] = e

    if kind == ARITHMETIC_EXTENSION:
        for i in range(NUM_ROUTED // (4 * D)):
            m0, m1, ad = rext(), rext(), rext()
            put(4 * D * i, m0); put(4 * D * i + D, m1); put(4 * D * i + 2 * D, ad)
            put(4 * D * i + 3 * D, ext_add(ext_scale(ext_mul(m0, m1), consts[0]), ext_scale(ad, consts[1])))
    elif kind == MUL_EXTENSION:
        for i in range(NUM_ROUTED // (3 * D)):
            m0, m1 = rext(), rext()
            put(3 * D * i, m0); put(3 * D * i + D, m1); put(3 * D * i + 2 * D, ext_scale(ext_mul(m0, m1), consts[0]))
    elif kind == BASE_SUM:
        base, num_limbs = params
        limbs = [rnd.randrange(base) for _ in range(num_limbs)]
        w[1:1 + num_limbs] = limbs
        w[0] = sum(l * pow(base, i, P) for i, l in enumerate(limbs)) % P
    elif kind in (REDUCING, REDUCING_EXTENSION):
        (num_coeffs,) = params
        cw = 1 if kind == REDUCING else D
        alpha, acc = rext(), rext()
        put(D, alpha); put(2 * D, acc)
        start_coeffs = 3 * D
        start_accs = start_coeffs + num_coeffs * cw
        for i in range(num_coeffs):
            if kind == REDUCING:
                coeff = (rnd.randrange(P), 0)
                w[start_coeffs + i] = coeff[0]
            else:
                coeff = rext()
                put(start_coeffs + D * i, coeff)
            acc = ext_add(ext_mul(acc, alpha), coeff)
            put(0 if i == num_coeffs - 1 else start_accs + D * i, acc)
    elif kind == RANDOM_ACCESS:
        bits, num_copies, num_extra = params
        L = random_access_layout(bits, num_copies, num_extra)
        for c in range(num_copies):
            idx = rnd.randrange(L["vec"])
            items = [rnd.randrange(P) for _ in range(L["vec"])]
            w[L["access"](c)] = idx
            w[L["claimed"](c)] = items[idx]
            for i, v in enumerate(items):
                w[L["item"](i, c)] = v
            for i in range(bits):
                w[L["bit"](i, c)] = (idx >> i) & 1
        for i in range(num_extra):
            w[L["extra"](i)] = consts[i]
    elif kind == EXPONENTIATION:
        (nb,) = params
        base = rnd.randrange(P)
        bits_ = [rnd.randrange(2) for _ in range(nb)]
        w[0] = base
        w[1:1 + nb] = bits_
        cur = 1
        for i in range(nb):
            prev = 1 if i == 0 else cur * cur % P
            cur = prev * (base if bits_[nb - 1 - i] else 1) % P
            w[2 + nb + i] = cur
        w[1 + nb] = cur
    elif kind == POSEIDON_MDS:
        cols = [mds([w[D * i + comp] for i in range(WIDTH)]) for comp in range(D)]
        for i in range(WIDTH):
            for comp in range(D):
                w[D * (WIDTH + i) + comp] = cols[comp][i]
    else:
        raise ValueError(kind)
    return w


# ---------------------------------------------------------------------------------------------------- selectors
def selector_groups(gates: Sequence[int], max_degree: int):
    """gates/selectors.rs selector_polynomials: gates are sorted by degree; one selector polynomial per group, groups grown
    greedily while  group size + the largest gate degree in it <= max_degree + 1  (one group, no UNUSED factor, if all fit).
    Returns (selector_indices per gate, groups as (begin, end))."""
    n = len(gates)
    degs = [gate_degree(*g) if isinstance(g, tuple) else gate_degree(g) for g in gates]
    assert degs == sorted(degs), "gates must be sorted by degree"
    if max(degs) + n - 1 <= max_degree:
        return [0] * n, [(0, n)]
    groups, sel, begin = [], [0] * n, 0
    while begin < n:
        end = begin + 1
        while end < n and (end + 1 - begin) + degs[end] <= max_degree:
            end += 1
        for i in range(begin, end):
            sel[i] = len(groups)
        groups.append((begin, end))
        begin = end
    return sel, groups


def compute_filter(row: int, group, s: int, many_selectors: bool) -> int:
    f = 1
    for i in range(group[0], group[1]):
        if i != row:
            f = f * (i - s) % P
    if many_selectors:
        f = f * (UNUSED_SELECTOR - s) % P
    return f


# ---------------------------------------------------------------------------------------------------- a synthetic circuit
class Circuit:
    """A satisfied standard_recursion_config-shaped instance: which gate sits on each row, wires, gate constants, copy
    constraints.  Rows: a chain of Poseidon permutations (each row's inputs are wired to the previous row's outputs), a few
    arithmetic rows fed by Poseidon outputs, one constant row, one public-input row, no-ops to the power of two."""

    def __init__(self, n_log: int, seed: int = 0, n_poseidon: int | None = None, extended: bool = False):
        """extended: also rows of PoseidonMds, BaseSum, Reducing, ReducingExtension, ArithmeticExtension, MulExtension,
        Exponentiation and RandomAccess gates (random satisfied instances, not wired to anything)"""
        rnd = random.Random(seed)
        n = 1 << n_log
        self.n_log, self.n = n_log, n
        # sorted by degree, as CircuitBuilder does
        self.gates = [NOOP, CONSTANT, PUBLIC_INPUT, ARITHMETIC, POSEIDON]
        if extended:
            self.gates = [NOOP, CONSTANT, PUBLIC_INPUT, POSEIDON_MDS, BASE_SUM, REDUCING, REDUCING_EXTENSION, ARITHMETIC,
                          ARITHMETIC_EXTENSION, MUL_EXTENSION, EXPONENTIATION, RANDOM_ACCESS, POSEIDON]
        self.gate_params = [EXT_GATE_PARAMS.get(g, ()) for g in self.gates]
        self.max_degree = 8
        self.selector_indices, self.groups = selector_groups(list(zip(self.gates, self.gate_params)), self.max_degree)
        self.num_selectors = len(self.groups)
        self.row_gate = [0] * n                     # index into self.gates
        self.wires = [[0] * n for _ in range(NUM_WIRES)]
        self.gate_consts = [[0] * n for _ in range(NUM_GATE_CONSTANTS)]
        self.pi_hash = [rnd.randrange(P) for _ in range(4)]
        classes: List[List[tuple]] = []             # copy-constraint classes of (wire, row) cells
        n_pos = n_poseidon if n_poseidon is not None else max(1, n // 2)
        n_other = 8 if extended else 0
        if extended and n - n_pos < n_other + 4:
            n_pos = max(1, n - n_other - 4)
        n_arith = max(1, min(n // 8, n - n_pos - 2 - n_other)) if n - n_pos - 2 - n_other > 0 else 0
        row = 0
        state = [rnd.randrange(P) for _ in range(WIDTH)]
        prev_out_row = None
        pos_rows = []
        for _ in range(n_pos):
            w = poseidon_gate_witness(state, swap=rnd.randrange(2))
            for j in range(NUM_WIRES):
                self.wires[j][row] = w[j]
            self.row_gate[row] = self.gates.index(POSEIDON)
            if prev_out_row is not None:
                for t in range(WIDTH):
                    classes.append([(W_OUT + t, prev_out_row), (W_IN + t, row)])
            prev_out_row = row
            pos_rows.append(row)
            state = [w[W_OUT + t] for t in range(WIDTH)]
            row += 1
        for _ in range(n_arith):
            c0, c1 = rnd.randrange(P), rnd.randrange(P)
            self.gate_consts[0][row], self.gate_consts[1][row] = c0, c1
            src = rnd.choice(pos_rows)
            for i in range(NUM_ROUTED // 4):
                m0 = self.wires[W_OUT + (i % WIDTH)][src]          # copied from a Poseidon output
                m1, addend = rnd.randrange(P), rnd.randrange(P)
                self.wires[4 * i][row], self.wires[4 * i + 1][row], self.wires[4 * i + 2][row] = m0, m1, addend
                self.wires[4 * i + 3][row] = (m0 * m1 % P * c0 + addend * c1) % P
                classes.append([(W_OUT + (i % WIDTH), src), (4 * i, row)])
            self.row_gate[row] = self.gates.index(ARITHMETIC)
            row += 1
        if extended:
            for gi, g in enumerate(self.gates):
                if g in (NOOP, CONSTANT, PUBLIC_INPUT, ARITHMETIC, POSEIDON) or row >= n:
                    continue
                cc = [rnd.randrange(P) for _ in range(NUM_GATE_CONSTANTS)]
                w = gate_witness(g, rnd, cc, self.gate_params[gi])
                assert not any(gate_constraints(g, w, cc, self.pi_hash, self.gate_params[gi]))
                for i in range(NUM_GATE_CONSTANTS):
                    self.gate_consts[i][row] = cc[i]
                for j in range(NUM_WIRES):
                    self.wires[j][row] = w[j]
                self.row_gate[row] = gi
                row += 1
        if row < n:
            c = [rnd.randrange(P), rnd.randrange(P)]
            for i in range(NUM_GATE_CONSTANTS):
                self.gate_consts[i][row] = c[i]
                self.wires[i][row] = c[i]
            self.row_gate[row] = self.gates.index(CONSTANT)
            row += 1
        if row < n:
            for i in range(4):
                self.wires[i][row] = self.pi_hash[i]
            self.row_gate[row] = self.gates.index(PUBLIC_INPUT)
            row += 1
        # remaining rows: NoopGate with arbitrary wire values
        for r in range(row, n):
            for j in range(NUM_WIRES):
                self.wires[j][r] = rnd.randrange(P)
            self.row_gate[r] = self.gates.index(NOOP)
        # selector polynomials (values on H) + constants: [selectors..., gate constants...]
        self.constants = []
        for k, grp in enumerate(self.groups):
            self.constants.append([self.row_gate[r] if grp[0] <= self.row_gate[r] < grp[1] else UNUSED_SELECTOR for r in range(n)])
        self.constants += self.gate_consts
        # sigma polynomials from the classes: the cells of a class form one cycle; every other routed cell is a fixed point
        self.k_is = [pow(GENERATOR, j, P) for j in range(NUM_ROUTED)]
        w = root(n_log)
        wp = [1] * n
        for i in range(1, n):
            wp[i] = wp[i - 1] * w % P
        nxt = {}
        merged: Dict[tuple, List[tuple]] = {}
        for cls in classes:                          # union the classes that share a cell
            cells = []
            for cell in cls:
                cells += merged.get(cell, [cell])
            cells = list(dict.fromkeys(cells))
            for cell in cells:
                merged[cell] = cells
        for cells in {id(v): v for v in merged.values()}.values():
            vals = {self.wires[j][r] for (j, r) in cells}
            assert len(vals) == 1, "copy constraint between unequal cells"
            for a, b in zip(cells, cells[1:] + cells[:1]):
                nxt[a] = b
        self.sigmas = [[0] * n for _ in range(NUM_ROUTED)]
        for j in range(NUM_ROUTED):
            for r in range(n):
                tj, tr = nxt.get((j, r), (j, r))
                self.sigmas[j][r] = self.k_is[tj] * wp[tr] % P


# ---------------------------------------------------------------------------------------------------- evaluation at a point
def vanishing_terms_at(c: Circuit, x: int, consts_row, sigmas_row, wires_row, zs_row, pp_rows, next_zs_row, betas, gammas,
                       degree: int) -> List[int]:
    """every vanishing term at one point x, from the values of all polynomials at x (and of Z at g x), in plonky2's order"""
    n = c.n
    C = len(betas)
    zh = (pow(x, n, P) - 1) % P
    l1 = zh * pow(n * (x - 1) % P, P - 2, P) % P
    terms = [l1 * (zs_row[i] - 1) % P for i in range(C)]
    chunks = [list(range(s, min(s + degree, NUM_ROUTED))) for s in range(0, NUM_ROUTED, degree)]
    for i in range(C):
        acc = [zs_row[i]] + list(pp_rows[i]) + [next_zs_row[i]]
        for l, ch in enumerate(chunks):
            num, den = 1, 1
            for j in ch:
                num = num * ((wires_row[j] + betas[i] * c.k_is[j] % P * x + gammas[i]) % P) % P
                den = den * ((wires_row[j] + betas[i] * sigmas_row[j] + gammas[i]) % P) % P
            terms.append((acc[l] * num - acc[l + 1] * den) % P)
    n_gc = max(gate_num_constraints(g, pr) for g, pr in zip(c.gates, c.gate_params))
    gc = [0] * n_gc
    for gi, g in enumerate(c.gates):
        k = c.selector_indices[gi]
        f = compute_filter(gi, c.groups[k], consts_row[k], c.num_selectors > 1)
        for t, v in enumerate(gate_constraints(g, wires_row, consts_row[c.num_selectors:], c.pi_hash, c.gate_params[gi])):
            gc[t] = (gc[t] + f * v) % P
    return terms + gc


def reduce_with_powers(terms: Sequence[int], alpha: int) -> int:
    acc = 0
    for t in reversed(terms):
        acc = (acc * alpha + t) % P
    return acc


def quotient_values(c: Circuit, lde_consts_sigmas, lde_wires, lde_zs_pp, betas, gammas, alphas, degree: int, rate_bits: int,
                    quotient_degree_bits: int) -> List[List[int]]:
    """compute_quotient_polys up to (not including) the coset_ifft.  lde_*: [columns][n << rate_bits] values in NATURAL order on the
    coset 7 <w_N> (column order of the committed batches: constants then sigmas; wires; Z of every challenge then the partial
    products challenge by challenge).  Returns [C][n << quotient_degree_bits] in natural order."""
    n_log, n = c.n_log, c.n
    C = len(betas)
    num_prods = (NUM_ROUTED + degree - 1) // degree - 1
    q_size = n << quotient_degree_bits
    step = 1 << (rate_bits - quotient_degree_bits)
    next_step = 1 << quotient_degree_bits
    wq = root(n_log + quotient_degree_bits)
    n_const = c.num_selectors + NUM_GATE_CONSTANTS
    out = [[0] * q_size for _ in range(C)]
    x = GENERATOR
    for i in range(q_size):
        row = i * step
        nrow = ((i + next_step) % q_size) * step
        consts_row = [int(col[row]) for col in lde_consts_sigmas[:n_const]]
        sigmas_row = [int(col[row]) for col in lde_consts_sigmas[n_const:]]
        wires_row = [int(col[row]) for col in lde_wires]
        zs_row = [int(lde_zs_pp[k][row]) for k in range(C)]
        next_zs = [int(lde_zs_pp[k][nrow]) for k in range(C)]
        pp_rows = [[int(lde_zs_pp[C + k * num_prods + l][row]) for l in range(num_prods)] for k in range(C)]
        terms = vanishing_terms_at(c, x, consts_row, sigmas_row, wires_row, zs_row, pp_rows, next_zs, betas, gammas, degree)
        zh_inv = pow((pow(x, n, P) - 1) % P, P - 2, P)
        for k in range(C):
            out[k][i] = reduce_with_powers(terms, alphas[k]) * zh_inv % P
        x = x * wq % P
    return out


def horner(coeffs: Sequence[int], x: int) -> int:
    acc = 0
    for a in reversed(coeffs):
        acc = (acc * x + int(a)) % P
    return acc


def check_quotient_identity(c: Circuit, coeff_consts_sigmas, coeff_wires, coeff_zs_pp, quotient_coeffs, betas, gammas, alphas,
                            degree: int, zeta: int) -> bool:
    """the verifier's identity at a base-field point zeta: quotient_c(zeta) * Z_H(zeta) == sum_t alpha_c^t term_t(zeta), with every
    term recomputed from polynomial evaluations at zeta and g * zeta (coefficients in, Horner)"""
    C = len(betas)
    num_prods = (NUM_ROUTED + degree - 1) // degree - 1
    g = root(c.n_log)
    n_const = c.num_selectors + NUM_GATE_CONSTANTS
    ev = lambda cols, x: [horner(col, x) for col in cols]
    cs = ev(coeff_consts_sigmas, zeta)
    wires_row = ev(coeff_wires, zeta)
    zpp = ev(coeff_zs_pp, zeta)
    next_zs = [horner(coeff_zs_pp[k], zeta * g % P) for k in range(C)]
    pp_rows = [[zpp[C + k * num_prods + l] for l in range(num_prods)] for k in range(C)]
    terms = vanishing_terms_at(c, zeta, cs[:n_const], cs[n_const:], wires_row, zpp[:C], pp_rows, next_zs, betas, gammas, degree)
    zh = (pow(zeta, c.n, P) - 1) % P
    return all(horner(quotient_coeffs[k], zeta) * zh % P == reduce_with_powers(terms, alphas[k]) for k in range(C))
