#!/usr/bin/env python3
"""Commit latency for the real-circuit shapes of SURVEY.md section 8 (n = 2^12..2^16, k in {135, 20, 16, 85}),
device-resident, median of 20 after 5 warm-ups, next to the CPU port on the same box.  GPU box only.
    python tools/latency_sweep.py [--cpu]"""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from intmax_zkp_core_b200 import device as D
from intmax_zkp_core_b200.plonky2 import Context as _Ctx
D.Context = _Ctx


def main():
    with_cpu = "--cpu" in sys.argv
    ctx = D.torch_context(0)
    rows = []
    for n_log in (12, 14, 16):
        for k in (135, 85, 20, 16):
            v = torch.randint(0, 2**62, (k, 1 << n_log), dtype=torch.int64, device="cuda")
            out = D.DeviceCommitment(n_log, k, 3, 4, v.device)
            ts = []
            for it in range(25):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                D.commit_device(ctx, v, 3, 4, out=out, is_coeffs=(k == 16))
                e1.record()
                torch.cuda.synchronize()
                if it >= 5:
                    ts.append(e0.elapsed_time(e1))
            rec = {"n_log": n_log, "k": k, "kind": "from_coeffs" if k == 16 else "from_values",
                   "gpu_ms_median": round(statistics.median(ts), 4), "gpu_ms_min": round(min(ts), 4),
                   "cells_per_s": (k << n_log) / (statistics.median(ts) * 1e-3)}
            if with_cpu:
                from oracle import oracle as O
                x = v.cpu().numpy().view(np.uint64)
                O.baseline_commit(x, 3, 4, is_coeffs=(k == 16))
                t0 = time.perf_counter()
                O.baseline_commit(x, 3, 4, is_coeffs=(k == 16))
                rec["cpu_port_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
                rec["cpu_threads"] = O.baseline_threads()
            rows.append(rec)
            print(json.dumps(rec), flush=True)
    ctx.close()
    # N2 shape: FRI commit-phase trees = MerkleTree::new over row-major leaves of arity * D = 32 elements
    import ctypes as C
    ctx2 = D.torch_context(0)
    for lg in (19, 15, 11):
        N = 1 << lg
        leaves = torch.randint(0, 2**62, (N, 32), dtype=torch.int64, device="cuda")
        dig = torch.empty((2 * (N - 16), 4), dtype=torch.int64, device="cuda")
        cap = torch.empty((16, 4), dtype=torch.int64, device="cuda")
        ts = []
        for it in range(15):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ctx2.check(ctx2._lib.b200zkp_dev_merkle(ctx2._h, C.c_void_p(leaves.data_ptr()), 32, 1, 32, N, 4,
                                                    C.c_void_p(dig.data_ptr()), C.c_void_p(cap.data_ptr())))
            e1.record()
            torch.cuda.synchronize()
            if it >= 5:
                ts.append(e0.elapsed_time(e1))
        print(json.dumps({"fri_layer_tree": True, "leaves": N, "leaf_len": 32, "cap_height": 4, "layout": "row-major",
                          "gpu_ms_median": round(statistics.median(ts), 4)}), flush=True)
    ctx2.close()
    # independent proofs on independent contexts (one stream each): small commitments overlap on the GPU
    for n_log, k in ((12, 135), (14, 135)):
        for n_ctx in (1, 4, 8):
            ctxs = [D.Context(0) for _ in range(n_ctx)]
            vs = [torch.randint(0, 2**62, (k, 1 << n_log), dtype=torch.int64, device="cuda") for _ in range(n_ctx)]
            outs = [D.DeviceCommitment(n_log, k, 3, 4, vs[0].device) for _ in range(n_ctx)]
            torch.cuda.synchronize()
            reps = 20
            for it in range(reps + 3):
                if it == 3:
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                for c, v, o in zip(ctxs, vs, outs):
                    D.commit_device(c, v, 3, 4, out=o)
            for c in ctxs:
                c.synchronize()
            dt = (time.perf_counter() - t0) / reps
            rec = {"n_log": n_log, "k": k, "concurrent_contexts": n_ctx, "ms_per_round_of_commits": round(dt * 1e3, 4),
                   "commits_per_s": round(n_ctx / dt, 1), "cells_per_s": n_ctx * (k << n_log) / dt}
            print(json.dumps(rec), flush=True)
            for c in ctxs:
                c.close()


if __name__ == "__main__":
    main()
