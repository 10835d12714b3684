"""Times the opening proof (prove_openings: rows N2 + N3) on the device-resident oracles of one plonky2 proof shape:
constants+sigmas k=85, wires k=135, Z+partial products k=20, quotient chunks k=16, standard_recursion_config.
Usage: python tools/fri_timing.py [n_log ...]   (prints one JSON line per size)"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import intmax_zkp_core_b200 as z
import intmax_zkp_core_b200.fri as zf


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [12, 14, 16]
    ctx = z.Context(0)
    for n_log in sizes:
        n = 1 << n_log
        rng = np.random.default_rng(n_log)
        ks = (85, 135, 20, 16)
        oracles = [z.PolynomialBatch.from_coeffs(rng.integers(0, 2**63, size=(k, n), dtype=np.uint64), 3, False, 4, ctx=ctx) for k in ks]
        cfg = zf.standard_recursion_fri_config()
        params = cfg.fri_params(n_log)
        zeta = (int(rng.integers(1, 2**63)), int(rng.integers(1, 2**63)))
        g = pow(1753635133440165772, 1 << (32 - n_log), zf.P)
        gz = (zeta[0] * g % zf.P, zeta[1] * g % zf.P)
        inst = zf.FriInstanceInfo([
            zf.FriBatchInfo(zeta, [zf.FriPolynomialInfo(o, i) for o, k in enumerate(ks) for i in range(k)]),
            zf.FriBatchInfo(gz, zf.FriPolynomialInfo.from_range(2, range(2)))])
        best = {}
        for it in range(4):
            ch = zf.Challenger(ctx)
            for o in oracles:
                ch.observe_cap(o._cap)
            ctx.synchronize()
            t0 = time.perf_counter()
            ev = [o.eval_ext2(np.array(zeta, np.uint64)) for o in oracles] + [oracles[2].eval_ext2(np.array(gz, np.uint64))]
            t1 = time.perf_counter()
            alpha = ch.get_extension_challenge()
            st = zf.FriCommitPhase.from_oracles(inst, oracles, alpha, True)
            ctx.synchronize()
            t2 = time.perf_counter()
            proof = zf.fri_proof(oracles, st, ch, params)
            t3 = time.perf_counter()
            st.close()
            cur = dict(openings_ms=(t1 - t0) * 1e3, final_poly_lde_ms=(t2 - t1) * 1e3, fri_proof_ms=(t3 - t2) * 1e3, total_ms=(t3 - t0) * 1e3)
            if it and (not best or cur["total_ms"] < best["total_ms"]):
                best = cur
        print(json.dumps(dict(workload=f"prove_openings 2^{n_log} x (85+135+20+16), standard_recursion_config", n_log=n_log,
                              layers=len(params.reduction_arity_bits), pow_witness=proof.pow_witness,
                              **{k: round(v, 3) for k, v in best.items()})), flush=True)
        del oracles


if __name__ == "__main__":
    main()
