#!/usr/bin/env python3
"""One inverse transform + one coset LDE at the benchmark shape (2^20 x 135, rate_bits 3), device resident, no hashing:
the short command the ncu captures of the transform passes run (profiles/README.md).
    python tools/ntt_profile.py [n_log] [k] [reps]     # prints CUDA-event ms per stage"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from intmax_zkp_core_b200 import device as D


def main():
    n_log = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 135
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    r = 3
    n, N = 1 << n_log, 1 << (n_log + r)
    ctx = D.torch_context(0)
    v = torch.randint(0, 2**62, (k, n), dtype=torch.int64, device="cuda")
    coeffs = torch.empty_like(v)
    lde = torch.empty((k, N), dtype=torch.int64, device="cuda")
    lib = ctx._lib

    def intt():
        ctx.check(lib.b200zkp_dev_intt(ctx._h, C.c_void_p(v.data_ptr()), n, C.c_void_p(coeffs.data_ptr()), n, C.c_void_p(lde.data_ptr()), n_log, k))

    def ldef():
        ctx.check(lib.b200zkp_dev_lde(ctx._h, C.c_void_p(coeffs.data_ptr()), n, C.c_void_p(lde.data_ptr()), N, n_log, k, r, 0, 1 << r))
    intt(); ldef()
    torch.cuda.synchronize()
    for name, fn in (("intt", intt), ("lde", ldef)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name} 2^{n_log} x {k}: {e0.elapsed_time(e1) / reps:.3f} ms")
    ctx.close()


if __name__ == "__main__":
    main()
