#!/bin/bash
# Round 2, call r: full suite (lean BCE epilogue, register-form select as default); ncu source capture of the select; bench A/B.
OUT=gpurun_out/r02r; mkdir -p $OUT
timeout 800 python -m pytest tests -m gpu -x -q --timeout=200 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log | cut -c1-300
timeout 120 python profiles/topn_bench.py > $OUT/topn_bench.jsonl 2> $OUT/topn_bench.err; echo "topn bench rc=$?"; grep -E '"batch": 4096' $OUT/topn_bench.jsonl | cut -c1-260
timeout 60 python profiles/topn_phase_profile.py > $OUT/topn_phase.jsonl 2>> $OUT/topn_bench.err; echo "phase rc=$?"; head -1 $OUT/topn_phase.jsonl | cut -c1-600
TOPN_REPS=1 TOPN_BATCHES=4096 timeout 300 ncu --set full --import-source on --clock-control none \
    --kernel-name-base mangled -k regex:topn_rowselect_kernelILb1ELb1ELb0ELb1E -c 2 -f -o $OUT/topn_regs python profiles/topn_bench.py > $OUT/topn_ncu.log 2>&1
echo "ncu rc=$?"
if [ -f $OUT/topn_regs.ncu-rep ]; then
  ncu -i $OUT/topn_regs.ncu-rep --page source --csv --print-source cuda,sass --kernel-id :::2 > $OUT/topn_regs_source.csv 2>/dev/null
  ncu -i $OUT/topn_regs.ncu-rep --page raw --csv > $OUT/topn_regs_raw.csv 2>/dev/null
  python profiles/ncu_source_hotspots.py $OUT/topn_regs_source.csv 45 > $OUT/topn_regs_stalls.txt; head -40 $OUT/topn_regs_stalls.txt | cut -c1-230
  rm -f $OUT/topn_regs_source.csv
fi
bash profiles/run_ab.sh r02r "CC_PRECISION=bf16"
