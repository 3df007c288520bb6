#!/bin/bash
# Round 2, call aa: launch_dependents in the helper kernels between GEMMs: train tests + bench.
OUT=gpurun_out/r02aa; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q --timeout=200 -k "train or baseline or dropin or kl or bag or chain" > $OUT/pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_sel.log | cut -c1-300
bash profiles/run_ab.sh r02aa
BENCH_ARGS="" bash profiles/run_ab.sh r02aa_2
