#!/usr/bin/env python
"""profiles/roofline_traffic.json (read by bench.py for `roofline.traffic`) from a trimmed ncu capture of the step
(profiles/summarize_ncu.py output): per-launch dram__bytes_read.sum + dram__bytes_write.sum of the big 512 <-> C GEMM
passes (grid 148, 256-wide pair tiles), of the softmax-KL kernel, Adam and the row select.

    python profiles/roofline_traffic_from_ncu.py profiles/r02_step_kernels_ncu_full.csv
"""
import csv
import json
import os
import sys

src = sys.argv[1]
rows = list(csv.reader(open(src)))
hdr = rows[0]
col = {h.split(" [")[0]: i for i, h in enumerate(hdr)}


def traffic(r):
    return (float(r[col["dram__bytes_read.sum"]]) + float(r[col["dram__bytes_write.sum"]])) * 1e9   # the export is in Gbyte


step_end = next(i for i, r in enumerate(rows[1:], 1) if "adam_kernel" in r[1])       # the train step comes first
big = [r for r in rows[1:step_end] if "gemm_tc_kernel<0" in r[1] and "256, 2>" in r[1] and r[col["launch__grid_size"]] == "148"]
out = {}
name = os.path.relpath(src, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out["gemm_tc_kernel<tf32>"] = {
    "dram_bytes_per_launch": int(sum(traffic(r) for r in big) / len(big)), "launches_averaged": len(big),
    "per_launch": [int(traffic(r)) for r in big],
    "source": f"{name} (ncu --set full, one warm train step; mean of dram__bytes_read.sum + dram__bytes_write.sum over the "
              f"{len(big)} 2*4096*512*20884-flop gemm_tc_kernel launches: fw1, fwd+BCE, fwd+store, 2 x (dW, dX), dW1)",
    "algorithmic_bytes_note": "fwd: 8.4 MB A + 42.8 MB W + 344 MB logits out; dW: 344 + 8.4 in, 42.8 out; dX: 344 + 42.8 in, 8.4 out; "
                              "fw1 / dW1: 344 MB dense x + 42.8 MB W1 (or out) + 8.4 MB"}
for key, pat in (("softmax_kl_regs_kernel", "softmax_kl_regs_kernel"), ("adam_kernel", "adam_kernel"),
                 ("topn_rowselect_kernel", "topn_rowselect_kernel"), ("noise_kernel", "noise_kernel")):
    rs = [r for r in rows[1:] if pat in r[1]]
    if rs:
        out[key] = {"dram_bytes_per_launch": int(sum(traffic(r) for r in rs) / len(rs)), "launches_averaged": len(rs),
                    "grid": rs[0][col["launch__grid_size"]], "source": f"{name} (ncu --set full)"}
out["topn_rowselect_kernel"]["cubes"] = 2048
out["topn_rowselect_kernel"]["note"] = "this launch ranks 2048 cubes (profiles/step_driver.py --recommend); 86.1 KB algorithmic per cube"
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "roofline_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
