#!/usr/bin/env python3
"""profiles/ntt_mix.json from an ncu report of the transform passes (no GPU needed):
    ncu --set full --import-source on --clock-control none -k regex:ct_pass --launch-skip 6 -c 6 -o rep python tools/ntt_profile.py 20 135
    python tools/ncu_ntt.py rep.ncu-rep [--json profiles/ntt_mix.json]
Per launch: duration, DRAM bytes and throughput, ALU / multiplier pipe utilisation, issue slots, warp instructions, registers;
keyed by the hash of the kernel sources at capture time (bench.py marks the block stale when they changed since)."""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def source_key():
    h = hashlib.sha256()
    for name in ("goldilocks.cuh", "ntt_ct_kernels.cuh"):
        with open(os.path.join(ROOT, "intmax_zkp_core_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = {h: i for i, h in enumerate(rows[0])}

    def col(r, name, default=None):
        i = hdr.get(name)
        try:
            return float(r[i]) if i is not None and r[i] != "" else default
        except ValueError:
            return default
    launches = []
    for r in rows[2:]:
        launches.append({
            "kernel": r[hdr["Kernel Name"]].split("(")[0].replace("void ", ""),
            "grid": r[hdr["Grid Size"]],
            "ms": col(r, "gpu__time_duration.sum"),
            "dram_read_gb": col(r, "dram__bytes_read.sum"), "dram_write_gb": col(r, "dram__bytes_write.sum"),
            "dram_gbs": round((col(r, "dram__bytes_read.sum", 0.0) + col(r, "dram__bytes_write.sum", 0.0)) / max(col(r, "gpu__time_duration.sum", 1.0), 1e-9) * 1e3, 1),
            "alu_pipe_pct": col(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "fmaheavy_pipe_pct": col(r, "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            "issue_pct": col(r, "sm__inst_issued.avg.pct_of_peak_sustained_active"),
            "warp_instructions": col(r, "smsp__inst_executed.sum"),
            "registers": col(r, "launch__registers_per_thread"),
            "warps_active_pct": col(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        })
    blk = {"source_key": source_key(), "shape": "2^20 x 135, rate_bits 3 (tools/ntt_profile.py 20 135): inverse transform passes 1-3, LDE passes 1-3",
           "captured_from": os.path.basename(rep), "launches": launches}
    js = json.dumps(blk, indent=1)
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            f.write(js + "\n")
    print(js)


if __name__ == "__main__":
    main()
