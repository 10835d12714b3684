#!/usr/bin/env python3
"""Static pipe model of a kernel from its SASS (no GPU needed).

    python tools/sass_mix.py <lib.so|.cubin> <function substring> [--weights a:w,...] [--block-weights w0,w1,...] [--blocks]

Splits the function into basic blocks, classifies every instruction by issue pipe (B200: the multiplier pipe "fmaheavy"
takes IMAD* — an IMAD.WIDE / IMAD.HI holds it for two slots — the ALU pipe takes IADD3 / LOP3 / SHF / LEA / SEL / ISETP /
PRMT / MOV) and prints per-block and weighted totals.  Weights: address ranges (hex, inclusive start of block) with an
execution count, e.g. the permutation loop's full-round-only blocks x8, partial-only x22, common x30.  With --auto-perm
the weights are derived for the Poseidon round loop: the largest loop is found from its backward branch, blocks inside it
that hold more than 100 IMAD.WIDE are the 11 extra S-boxes of a full round (x8), the rest of the loop x30, and everything
outside x1 — the same accounting profiles/README.md uses for the ncu source-page counters.
"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict

FMA2 = ("IMAD.WIDE", "IMAD.HI")                       # two slots on the multiplier pipe
FMA1 = ("IMAD", "IMUL")                                # one slot
ALU = ("IADD3", "IADD", "LOP3", "SHF", "LEA", "SEL", "ISETP", "PRMT", "MOV", "ICMP", "IMNMX", "VIADD", "VIMNMX", "PLOP3",
       "IABS", "FLO", "POPC", "BREV", "SGXT", "BMSK", "P2R", "R2P", "CS2R", "UMOV")


def pipe(op: str) -> str:
    if op.startswith("U") and not op.startswith("UMOV"):
        return "uniform"
    for p in FMA2:
        if op.startswith(p):
            return "fma2"
    for p in FMA1:
        if op.startswith(p):
            return "fma1"
    base = op.split(".")[0]
    if base in ALU:
        return "alu"
    if base in ("LDG", "STG", "LDS", "STS", "LDC", "LDCU", "LD", "ST", "LDL", "STL", "ATOMG", "RED", "SHFL"):
        return "lsu"
    if base in ("BRA", "EXIT", "BSSY", "BSYNC", "CALL", "RET", "NOP", "BAR", "WARPSYNC", "S2R", "S2UR", "BRX", "JMP"):
        return "ctl"
    return "other:" + base


def disasm(path, fn):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, funcs = None, OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
    hits = [k for k in funcs if fn in k]
    if not hits:
        sys.exit(f"no function matching {fn}; have: {list(funcs)[:50]}")
    return hits[0], funcs[hits[0]]


def opcode(text):
    t = text
    if t.startswith("@"):
        t = t.split(None, 1)[1]
    return t.split()[0]


def blocks_of(ins):
    leaders = {ins[0][0]}
    for i, (a, t) in enumerate(ins):
        op = opcode(t)
        if op.startswith("BRA") or op.startswith("EXIT") or op.startswith("BRX"):
            if i + 1 < len(ins):
                leaders.add(ins[i + 1][0])
            m = re.search(r"0x([0-9a-f]+)\s*$", t)
            if m and op.startswith("BRA"):
                leaders.add(int(m.group(1), 16))
    blocks, cur = [], []
    for a, t in ins:
        if a in leaders and cur:
            blocks.append(cur)
            cur = []
        cur.append((a, t))
    if cur:
        blocks.append(cur)
    return blocks


def summarize(c: Counter):
    fma2 = sum(v for k, v in c.items() if pipe(k) == "fma2")
    fma1 = sum(v for k, v in c.items() if pipe(k) == "fma1")
    alu = sum(v for k, v in c.items() if pipe(k) == "alu")
    tot = sum(c.values())
    return dict(total=tot, wide=fma2, fma1=fma1, mul_slots=2 * fma2 + fma1, alu=alu, other=tot - fma2 - fma1 - alu)


def main():
    path, fn = sys.argv[1], sys.argv[2]
    name, ins = disasm(path, fn)
    blocks = blocks_of(ins)
    weights = {}
    if "--weights" in sys.argv:
        for spec in sys.argv[sys.argv.index("--weights") + 1].split(","):
            rng, w = spec.split(":")
            weights[int(rng, 16)] = float(w)
    if "--block-weights" in sys.argv:      # one execution count per basic block, in address order
        ws = [float(x) for x in sys.argv[sys.argv.index("--block-weights") + 1].split(",")]
        assert len(ws) == len(blocks), (len(ws), len(blocks))
        weights = {b[0][0]: w for b, w in zip(blocks, ws)}
    auto = "--auto-perm" in sys.argv
    if auto:
        # largest backward branch = the round loop
        best = None
        for a, t in ins:
            if opcode(t).startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)\s*$", t)
                if m and int(m.group(1), 16) < a and (best is None or a - int(m.group(1), 16) > best[1] - best[0]):
                    best = (int(m.group(1), 16), a)
        lo, hi = best
        for b in blocks:
            a0 = b[0][0]
            if lo <= a0 <= hi:
                nw = sum(1 for _, t in b if opcode(t).startswith("IMAD.WIDE"))
                na = sum(1 for _, t in b if pipe(opcode(t)) == "alu")
                # 11 extra S-boxes of a full round, or the 12 round-constant adds in front of a full round: 8 of 30 trips
                weights[a0] = 8.0 if (nw > 100 or (nw == 0 and na >= 40)) else 30.0
            else:
                weights[a0] = 1.0
    total = Counter()
    print(f"function {name}: {len(ins)} instructions, {len(blocks)} blocks")
    for b in blocks:
        c = Counter(opcode(t) for _, t in b)
        w = weights.get(b[0][0], 1.0)
        if "--blocks" in sys.argv:
            s = summarize(c)
            last = b[-1][1]
            print(f"  {b[0][0]:06x}-{b[-1][0]:06x} x{w:<5g} n={s['total']:5d} wide={s['wide']:4d} fma1={s['fma1']:4d} alu={s['alu']:4d}  | {last}")
        for k, v in c.items():
            total[k] += v * w
    s = summarize(total)
    print("weighted:", {k: round(v) for k, v in s.items()})
    by = sorted(total.items(), key=lambda kv: -kv[1])
    print("  " + ", ".join(f"{k} {v:g}" for k, v in by[:28]))
    bound = max(s["mul_slots"], s["alu"], s["total"] / 2)
    print(f"  model: multiplier slots {s['mul_slots']:.0f}, alu slots {s['alu']:.0f}, issue/2 {s['total'] / 2:.0f} -> bound {bound:.0f} slots")


if __name__ == "__main__":
    main()
