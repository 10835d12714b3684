#!/usr/bin/env python3
"""Row N1a timing: Z + partial products (80 routed wires, chunks of 8, 2 challenges) on device-resident wires, followed by
the commitment of the resulting 20 columns.  GPU box only.   python tools/perm_timing.py [n_log ...]"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from intmax_zkp_core_b200 import device as D, prover as Z


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [12, 14, 16, 20]
    ctx = D.torch_context(0)
    k_is = Z.get_unique_coset_shifts(80)
    betas, gammas = [0x1234567890ABCDEF % Z.P, 77], [0xFEDCBA0987654321 % Z.P, 99]
    for n_log in sizes:
        n = 1 << n_log
        wires = torch.randint(0, 2**62, (80, n), dtype=torch.int64, device="cuda")
        sig = torch.randint(0, 2**62, (80, n), dtype=torch.int64, device="cuda")
        out = torch.empty((20, n), dtype=torch.int64, device="cuda")
        com = D.DeviceCommitment(n_log, 20, 3, 4, out.device)
        tz, tc = [], []
        for it in range(12):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            Z.zs_partial_products_device(ctx, wires, sig, k_is, betas, gammas, 8, out=out)
            e[1].record()
            D.commit_device(ctx, out, 3, 4, out=com)
            e[2].record()
            torch.cuda.synchronize()
            if it >= 2:
                tz.append(e[0].elapsed_time(e[1]))
                tc.append(e[1].elapsed_time(e[2]))
        mz = statistics.median(tz)
        rec = {"workload": "wires_permutation_partial_products_and_zs, 80 routed wires, degree 8, 2 challenges", "n_log": n_log,
               "zs_partial_products_ms": round(mz, 4), "commit_20_columns_ms": round(statistics.median(tc), 4),
               "rows_per_s": n / (mz * 1e-3), "read_GBps": 2 * 80 * n * 8 / (mz * 1e-3) / 1e9}
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
