#!/usr/bin/env python
"""Top source lines of an `ncu --page source --csv --print-source cuda,sass` export, by warp-stall samples and by executed
instructions, with the three largest stall reasons of each line:

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-id :::N > X_source.csv
    python profiles/ncu_source_hotspots.py X_source.csv [top=30] > profiles/...txt
"""
import csv
import sys

src = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
REASONS = ["barrier", "branch_resolving", "dispatch", "drain", "lg", "long_sb", "math", "membar", "mio", "misc", "no_inst",
           "not_selected", "selected", "short_sb", "sleep", "tex", "wait"]
cur, out, kernel = None, [], None
for r in csv.reader(open(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1]
    if len(r) >= 2 and r[0] == "Function Name" and kernel is None:
        kernel = r[1]
    if cur and len(r) > 48 and r[0].isdigit():
        num = lambda x: int(x) if x.isdigit() else 0
        out.append((num(r[6]), num(r[7]), cur.split("/")[-1], int(r[0]), r[1].strip(), [num(x) for x in r[31:48]]))
samples, instr = sum(o[0] for o in out), sum(o[1] for o in out)
print(f"# {src}\n# kernel: {kernel}\n# {samples} warp-stall samples, {instr} warp instructions executed (source-line rows)")
tot = [sum(o[5][i] for o in out) for i in range(len(REASONS))]
print("# stall reasons, all lines: " + ", ".join(f"{n} {100 * t / max(samples, 1):.1f}%" for t, n in sorted(zip(tot, REASONS), reverse=True)[:8]))
print(f"# {'samples':>8} {'instr':>7}  location: source   [largest stall reasons]")
for s, ie, f, ln, text, st in sorted(out, reverse=True)[:top]:
    big = ", ".join(f"{n} {v}" for v, n in sorted(zip(st, REASONS), reverse=True)[:3] if v)
    print(f"{100 * s / max(samples, 1):7.1f}% {100 * ie / max(instr, 1):6.1f}%  {f}:{ln}: {text[:100]}   [{big}]")
