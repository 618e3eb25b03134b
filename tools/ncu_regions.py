#!/usr/bin/env python
"""Per-region instruction and stall-sample totals of one kernel from an ncu report's SASS source page.
    ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv
    python tools/ncu_regions.py src.csv [lo:hi:name ...]      (hex offsets from the kernel's first instruction)
Without regions: the top 40 instructions by samples and the opcode histogram weighted by executions."""
import csv
import sys
from collections import Counter


def load(path):
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    out = []
    for r in rows[hdr_i + 1:]:
        if len(r) <= iex or not r[ia].startswith("0x"):
            continue
        out.append((int(r[ia], 16), r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0)))
    base = out[0][0]
    return [(a - base, s, n, e) for a, s, n, e in out]


def opcode(s):
    t = s.split()
    if t and t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0] if t else "?"


def main():
    ins = load(sys.argv[1])
    tot_e = sum(i[3] for i in ins)
    tot_s = sum(i[2] for i in ins)
    print(f"{len(ins)} instructions, {tot_e} executed, {tot_s} samples")
    regions = []
    for a in sys.argv[2:]:
        lo, hi, name = a.split(":")
        regions.append((int(lo, 16), int(hi, 16), name))
    if regions:
        rest_e, rest_s = tot_e, tot_s
        for lo, hi, name in regions:
            sel = [i for i in ins if lo <= i[0] < hi]
            e, s = sum(i[3] for i in sel), sum(i[2] for i in sel)
            rest_e -= e
            rest_s -= s
            ops = Counter()
            for i in sel:
                ops[opcode(i[1])] += i[3]
            top = ", ".join(f"{k} {v / max(e, 1):.2f}" for k, v in ops.most_common(6))
            print(f"{name:24s} [{lo:#06x},{hi:#06x}) static {len(sel):5d}  executed {e:10d} ({100 * e / tot_e:5.1f} %)  samples {s:7d} ({100 * s / tot_s:5.1f} %)  {top}")
        print(f"{'(elsewhere)':24s} executed {rest_e:10d} ({100 * rest_e / tot_e:5.1f} %)  samples {rest_s:7d} ({100 * rest_s / tot_s:5.1f} %)")
        return
    ops = Counter()
    for i in ins:
        ops[opcode(i[1])] += i[3]
    print("opcode histogram (share of executed warp instructions):")
    for k, v in ops.most_common(24):
        print(f"  {k:12s} {v:10d}  {100 * v / tot_e:5.1f} %")
    print("hottest instructions by samples:")
    for a, s, n, e in sorted(ins, key=lambda i: -i[2])[:40]:
        print(f"  {a:#06x} samples {n:6d} executed {e:9d}  {s}")


if __name__ == "__main__":
    main()
