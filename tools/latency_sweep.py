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


if __name__ == "__main__":
    main()
