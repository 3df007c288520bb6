#!/usr/bin/env python
"""Opcode histogram of the shipped library's SASS (sm_100a), so the tcgen05 / TMA / TMEM claims can be checked without
rebuilding:

    python profiles/sass_histogram.py > profiles/r02_sass_opcode_histogram.txt

UTCHMMA / UTCIMMA = tcgen05.mma (kind::tf32|f16 / kind::i8), UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store /
reduce-add, UBLKCP = cp.async.bulk, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, LDGSTS = cp.async,
MULTIMEM = multimem.ld_reduce / multimem.st (NVSwitch), REDG / ATOMG = global reductions / atomics."""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "cubecobrarecommender_b200/libcubecobra_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
SPECIAL = re.compile(r"^(UTCHMMA|UTCIMMA|UTCQMMA|UTCOMMA|UTMALDG|UTMASTG|UTMAREDG|UTMAPF|UBLKCP|LDTM|STTM|UTCBAR|UTCCP|UTCATOMSWS|"
                     r"LDGSTS|MULTIMEM|SYNCS|UCGABAR|ACQBULK|CCTL)")
per_kernel = collections.defaultdict(collections.Counter)
total = collections.Counter()
fn = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if not m or fn is None:
        continue
    op = m.group(1)
    total[op.split(".")[0]] += 1
    if SPECIAL.match(op):
        per_kernel[fn][op] += 1
print(f"# cuobjdump -sass {lib} ({os.path.getsize(lib)} bytes), {sum(total.values())} instructions in {len(per_kernel)} kernels with "
      f"Blackwell-specific opcodes")
print("# --- Blackwell-specific opcodes per kernel ---")
for fn in sorted(per_kernel):
    print(fn)
    for op, n in sorted(per_kernel[fn].items(), key=lambda kv: (-kv[1], kv[0])):
        print(f"    {n:5d}  {op}")
print("# --- all opcodes, whole library ---")
for op, n in total.most_common():
    print(f"{n:8d}  {op}")
