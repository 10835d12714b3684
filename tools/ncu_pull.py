#!/usr/bin/env python3
"""One partitioned commitment (device inputs) with ONE process driving two GPUs (b200zkp_comm_init_all: the ordering between the
devices is CUDA events, so ncu can replay kernels), the short command the ncu capture of the fused gather pass
(ntc::ct_pull_kernel) runs:
    ncu --set full --import-source on --clock-control none -k regex:ct_pull -c 2 -o gpurun_out/r2_pull python tools/ncu_pull.py
    python tools/ncu_pull.py [n_log] [k] [reps] [gpus]      # without ncu: prints wall-clock ms of whole commits
ncu cannot replay kernels of two devices that wait for each other (it failed with UnknownError on the 2-GPU form), so the capture
under profiles/ is taken with gpus = 1: a one-rank communicator runs the same kernel on the same tiles, all 8 coset blocks per
tile, reading the shard from the local gather buffer instead of a peer's window (no NVLink in that capture)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import intmax_zkp_core_b200 as z
from intmax_zkp_core_b200 import device as D


def main():
    n_log = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 135
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    G = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    ctxs = [z.Context(g) for g in range(G)]
    comm = D.Comm.init_all(ctxs)
    assert comm.peer_exchange or G == 1
    sh = D.ShardedCommitment(comm, n_log, k, 3, 4)
    n = 1 << n_log
    dv = [torch.randint(0, 2**62, (max(l["n_cols"], 1), n), dtype=torch.int64, device=torch.device("cuda", i))
          for i, l in enumerate(sh.layouts)]
    sh.run(dv)
    sh.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(reps):
        sh.run(dv)
    sh.synchronize()
    print(f"2^{n_log} x {k} over {G} GPUs (one process): {(time.perf_counter() - t0) / reps * 1e3:.3f} ms per commit")
    sh.close()
    comm.close()
    for c in ctxs:
        c.close()


if __name__ == "__main__":
    main()
