#!/usr/bin/env python3
"""Integer-pipe micro-benchmarks (b200zkp_int_pipe_bench, every kind): giga thread-instructions/s and cycles per
warp-instruction per SM sub-partition.  GPU box only.   python tools/pipe_bench.py"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import intmax_zkp_core_b200 as z

NAMES = ["imad_wide", "iadd3", "imad", "imad_wide+lop3", "lop3", "imad_hi", "imad+lop3", "iadd3_carry_pair", "imad_wide_noacc",
         "dfma", "dfma+imad_wide", "dfma+imad", "dfma+lop3", "imad_wide+imad", "imad_wide_noacc+lop3", "imad_wide+3lop3",
         "iadd3_3in_ur", "iadd3_3in+imad", "wide:imad:lop3:iadd3=1:2:2:3", "wide:lop3:imad=1:2:1"]


def main():
    import torch
    ctx = z.Context(0)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    mhz = 1965.0
    out = {}
    for kind, name in enumerate(NAMES):
        g = C.c_double()
        ctx.check(ctx._lib.b200zkp_int_pipe_bench(ctx._h, kind, 2000, C.byref(g)))
        cyc = sms * 4 * 32 * mhz * 1e6 / (g.value * 1e9)
        out[name] = {"gips": round(g.value, 1), "cycles_per_warp_instr_at_1965MHz": round(cyc, 3)}
        print(json.dumps({"kind": kind, "name": name, **out[name]}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
