#!/usr/bin/env python3
"""Dynamic instruction mix of a kernel from an ncu report taken with --import-source on (no GPU needed).

    python tools/ncu_mix.py <report.ncu-rep> <kernel substring> [--units N] [--json out.json] [--top K]

Aggregates the source page's "Instructions Executed" (warp instructions) by SASS opcode and by issue pipe (multiplier pipe
"fmaheavy": IMAD*, an IMAD.WIDE / IMAD.HI holds it for two slots; ALU pipe: IADD3 / LOP3 / SHF / LEA / SEL / ISETP / ...).
--units N divides every count by N warp-level work units (e.g. permutations / 32 for the leaf hash).  --json writes the
profiles/leaf_mix.json block bench.py reads (keyed by the hash of the kernel sources at capture time)."""
import csv
import io
import json
import os
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from sass_mix import pipe  # noqa: E402


def kernels(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    cur, rows, res = None, [], []
    for rec in csv.reader(io.StringIO(out)):
        if rec and rec[0] == "Kernel Name":
            if cur:
                res.append((cur, rows))
            cur, rows = rec[1], []
        elif cur:
            rows.append(rec)
    if cur:
        res.append((cur, rows))
    return res


def main():
    report, pat = sys.argv[1], sys.argv[2]
    units = float(sys.argv[sys.argv.index("--units") + 1]) if "--units" in sys.argv else 1.0
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30
    hits = [(k, rows) for k, rows in kernels(report) if pat in k]
    if not hits:
        sys.exit(f"no kernel matching {pat}")
    name, rows = hits[0]
    hdr = rows[0]
    i_src, i_exec = hdr.index("Source"), hdr.index("Instructions Executed")
    ops, pipes = Counter(), Counter()
    for r in rows[1:]:
        try:
            n = float(r[i_exec])
        except (ValueError, IndexError):
            continue
        toks = r[i_src].split()
        if toks and toks[0].startswith("@"):
            toks = toks[1:]
        if not toks:
            continue
        op = toks[0].rstrip(";")
        ops[op] += n
        pipes[pipe(op)] += n
    total = sum(ops.values())
    print(f"{name}: {total / units:.1f} warp-instructions per unit ({units:g} units)")
    wide = pipes["fma2"] / units
    fma1 = pipes["fma1"] / units
    alu = pipes["alu"] / units
    print(f"  multiplier pipe: {wide:.1f} two-slot (IMAD.WIDE/HI) + {fma1:.1f} one-slot = {2 * wide + fma1:.1f} slots;  ALU pipe: {alu:.1f};  other: {(total / units) - wide - fma1 - alu:.1f}")
    for op, n in ops.most_common(top):
        print(f"  {op:32s} {n / units:12.1f}")
    if "--json" in sys.argv:
        import bench
        path = sys.argv[sys.argv.index("--json") + 1]
        block = {"source_key": bench.kernel_source_key(), "kernel": name.split("(")[0],
                 "instructions": round(total / units, 1), "imad_wide": round(wide, 1), "fma_single_slot": round(fma1, 1), "alu": round(alu, 1),
                 "unit": "warp-instructions per warp-permutation (dynamic, ncu source-page counters / permutations)",
                 "captured_from": [os.path.basename(report)]}
        if os.path.exists(path):
            old = json.load(open(path))
            for k in ("ncu", "traffic_bytes", "traffic_note", "note"):
                if k in old and k not in block:
                    block[k] = old[k]
        json.dump(block, open(path, "w"), indent=1)
        print("wrote", path)


if __name__ == "__main__":
    main()
