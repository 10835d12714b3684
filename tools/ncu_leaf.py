#!/usr/bin/env python3
"""One commitment (2^n_log x k, rate_bits 3, cap_height 4) for an ncu capture of the leaf hash.  GPU box only.
    ncu --set full --import-source on --clock-control none -k regex:leaf_hash_kernel -c 1 -o gpurun_out/x python tools/ncu_leaf.py [n_log] [k]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from intmax_zkp_core_b200 import device as D


def main():
    n_log = int(sys.argv[1]) if len(sys.argv) > 1 else 17
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 135
    ctx = D.torch_context(0)
    v = torch.randint(0, 2**62, (k, 1 << n_log), dtype=torch.int64, device="cuda")
    out = D.DeviceCommitment(n_log, k, 3, 4, v.device)
    for _ in range(2):
        D.commit_device(ctx, v, 3, 4, out=out)
    torch.cuda.synchronize()
    print("done", n_log, k)


if __name__ == "__main__":
    main()
