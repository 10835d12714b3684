#!/usr/bin/env python3
"""Short GPU check of a library build (B200ZKP_LIB selects it): the reference Poseidon fixture, x^7 on edge values, one
2^10 x 135 commitment bit-exact against the oracle (139 k permutations), then the stage times of the 2^20 x 135 commitment.
    B200ZKP_LIB=$PWD/build/variants/x.so python tools/quick_validate.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import __graft_entry__ as G
import intmax_zkp_core_b200 as z
from intmax_zkp_core_b200 import device as D
from oracle import oracle as O

P = O.P


def main():
    G.smoke()
    ctx = z.Context(0)
    edge = [0, 1, 2, P - 1, P - 2, 2**32 - 1, 2**32, 2**32 + 1, 2**63, 2**64 - 1, P, P + 1, 0xFFFFFFFF00000000, 0x00000000FFFFFFFF]
    rng = np.random.default_rng(3)
    a = np.array(edge + [int(x) for x in rng.integers(0, 2**64, size=4096, dtype=np.uint64)], dtype=np.uint64)
    out = np.empty_like(a)
    ctx.check(ctx._lib.b200zkp_field_op(ctx._h, 6, a.ctypes.data, a.ctypes.data, a.size, out.ctypes.data))
    assert [int(v) for v in out] == [pow(int(x) % P, 7, P) for x in a], "x^7 mismatch"
    st = rng.integers(0, 2**64, size=(3000, 12), dtype=np.uint64)       # > 2048: throughput form; the first 100: latency form
    assert (z.PoseidonPermutation.permute(st, ctx) == O.permute_many(st)).all()
    assert (z.PoseidonPermutation.permute(st[:100], ctx) == O.permute_many(st[:100])).all()
    ctx.close()
    tctx = D.torch_context(0)
    tctx.set_timing(True)
    v = torch.randint(0, 2**62, (135, 1 << 20), dtype=torch.int64, device="cuda")
    out = D.DeviceCommitment(20, 135, 3, 4, v.device)
    ts = []
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        D.commit_device(tctx, v, 3, 4, out=out)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    print(json.dumps({"lib": os.environ.get("B200ZKP_LIB", "default"), "commit_ms": sorted(ts)[len(ts) // 2], "all": ts}))


if __name__ == "__main__":
    main()
