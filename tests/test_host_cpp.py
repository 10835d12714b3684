"""host/plonky2_api.hpp (the C++ mirror of the plonky2 interface above the C ABI): compiles everywhere, runs on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "test_host_api")


def _build():
    import intmax_zkp_core_b200 as z
    z.lib()   # make sure libb200zkp.so exists
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    libdir = os.path.join(ROOT, "intmax_zkp_core_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-o", BIN, os.path.join(ROOT, "host", "test_host_api.cpp"),
                           "-L" + libdir, "-lb200zkp", "-Wl,-rpath," + libdir])


def test_host_cpp_api_compiles_and_links():
    _build()
    assert os.path.exists(BIN)


@pytest.mark.gpu
def test_host_cpp_api_runs_on_gpu():
    _build()
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host C++ API ok" in out.stdout
    # the opening proof made through host/fri_api.hpp equals the Python mirror's on the same inputs and transcript
    import numpy as np
    import intmax_zkp_core_b200 as z
    import intmax_zkp_core_b200.fri as zf
    from oracle import oracle as O
    ctx = z.Context(0)
    cfg = zf.FriConfig(rate_bits=2, cap_height=1, proof_of_work_bits=7, reduction_strategy=zf.ConstantArityBits(2, 2), num_query_rounds=5)
    params = cfg.fri_params(6)
    o0 = z.PolynomialBatch.from_coeffs(O.synthetic_values(3, 64, seed=5), 2, False, 1, ctx=ctx)
    o1 = z.PolynomialBatch.from_coeffs(O.synthetic_values(2, 64, seed=6), 2, False, 1, ctx=ctx)
    ch = zf.Challenger(ctx)
    ch.observe_cap(o0._cap)
    ch.observe_cap(o1._cap)
    inst = zf.FriInstanceInfo([
        zf.FriBatchInfo((123456789, 987654321), [zf.FriPolynomialInfo(0, i) for i in range(3)] + [zf.FriPolynomialInfo(1, i) for i in range(2)]),
        zf.FriBatchInfo((555, 777), [zf.FriPolynomialInfo(1, 0), zf.FriPolynomialInfo(1, 1)])])
    proof = zf.prove_openings(inst, [o0, o1], ch, params, True)
    d = 0xcbf29ce484222325
    words = []
    for cap in proof.commit_phase_merkle_caps:
        words += cap.flatten().tolist()
    words += proof.final_poly.reshape(-1).tolist() + [proof.pow_witness]
    for r in proof.query_round_proofs:
        for row, p in r.initial_trees_proof.evals_proofs:
            words += row.tolist() + np.asarray(p.siblings).reshape(-1).tolist()
        for st in r.steps:
            words += st.evals.reshape(-1).tolist() + np.asarray(st.merkle_proof.siblings).reshape(-1).tolist()
    words.append(ch.get_challenge())
    for w in words:
        d = ((d ^ int(w)) * 0x100000001b3) % 2**64
    assert f"fri digest {d:016x} pow {proof.pow_witness}" in out.stdout, out.stdout
    ctx.close()
