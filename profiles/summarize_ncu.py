#!/usr/bin/env python
"""Trim an `ncu --page raw --csv` export to the columns the roofline discussion uses.

    python profiles/summarize_ncu.py gpurun_out/r01_prof_step_raw.csv profiles/r01_step_kernels.csv
"""
import csv
import sys

KEEP = [
    "ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def main(src, dst):
    with open(src) as f:
        r = csv.reader(f)
        hdr, units = next(r), next(r)
        idx = [hdr.index(k) for k in KEEP if k in hdr]
        with open(dst, "w", newline="") as g:
            w = csv.writer(g)
            w.writerow([f"{hdr[i]} [{units[i]}]" if units[i] else hdr[i] for i in idx])
            for row in r:
                out = [row[i] for i in idx]
                out[1] = out[1].split("(")[0][:80]
                w.writerow(out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
