"""CPU checks around the Poseidon kernel that need no GPU: the integer model of the split-basis partial rounds
(tools/poseidon_crt_model.py: algebra against the naive permutation, limb bounds by interval arithmetic), the constant
tables derived from it, and two build-time properties of the compiled leaf hash (no register spills, wide-multiply count)."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_split_basis_model_equals_naive_permutation():
    import poseidon_crt_model as M
    assert M.self_check(trials=4)


@pytest.mark.parametrize("lean", [False, True])
def test_limb_bounds_fit_32_bits(lean):
    """every limb of the split-basis rounds stays inside a signed 32-bit register (the Iv class asserts it on every
    operation), the packed words of x0 and of the leaving state are non-negative after their biases"""
    import poseidon_crt_model as M
    rep = M.bounds(lean=lean)
    for name, limbs in rep.items():
        for iv in limbs:
            assert -(1 << 31) <= iv.lo and iv.hi < (1 << 31), (name, iv)
    assert all(iv.lo >= 0 for iv in rep["4 x0 + bias"])
    assert all(iv.lo + (1 << M.BIAS_OUT_LOG) >= 0 for iv in rep["leaving limb"])


def test_split_add_table_matches_the_model():
    """SPLIT_ADD in csrc/poseidon_tables.cuh = pushed constants with the packing biases taken out, plus an all-zero row"""
    import poseidon_crt_model as M
    import poseidon_derive as PD
    sc, tv = M.partial_constants()
    text = open(os.path.join(ROOT, "intmax_zkp_core_b200", "csrc", "poseidon_tables.cuh")).read()
    body = text[text.index("SPLIT_ADD[372]"):]
    vals = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{16})ull", body[:body.index("};")])]
    assert len(vals) == 372
    rc = PD.round_constants()
    assert vals[:48] == rc[:48]
    assert [vals[(4 + i) * 12] for i in range(22)] == sc
    assert all(vals[(4 + i) * 12 + j] == 0 for i in range(22) for j in range(1, 12))
    assert vals[26 * 12:27 * 12] == tv
    assert vals[27 * 12:30 * 12] == rc[27 * 12:]
    assert vals[360:] == [0] * 12


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")
def test_leaf_hash_build_properties():
    """the shipped leaf hash keeps its state in registers (no stack frame at the 256 x 3 CTA shape) and holds the wide
    multiplies the design accounts for; a regression here is a performance bug the GPU tests would not see"""
    import intmax_zkp_core_b200 as z
    from intmax_zkp_core_b200 import _lib
    z.lib()
    out = subprocess.run(["cuobjdump", "-res-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    m = re.search(r"Function _ZN6merkle16leaf_hash_kernel\S*:\s*\n\s*REG:(\d+) STACK:(\d+)", out)
    assert m, "leaf_hash_kernel not found in the library"
    regs, stack = int(m.group(1)), int(m.group(2))
    assert regs <= 85 and stack == 0, (regs, stack)
    import sass_mix as S
    _, ins = S.disasm(_lib.LIB_PATH, "leaf_hash_kernel")
    wide = sum(1 for _, t in ins if S.opcode(t).startswith("IMAD.WIDE"))
    assert 250 <= wide <= 330, wide          # 12 + 1 S-boxes x 20, the packing folds, addressing
    assert not any(S.opcode(t).startswith(("LDL", "STL")) for _, t in ins)
