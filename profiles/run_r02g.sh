#!/bin/bash
# round 2, call g: new-kernel tests, select / KL micro-benchmarks, ncu source view of the KL kernel
OUT=gpurun_out/r02g; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_graph.py tests/test_gpu_dae.py tests/test_gpu_dropin.py -m gpu -x -q --timeout=200 \
  -k "topn or stream_k or first_layer or web or resume or overflow or row_block" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest.log
timeout 120 python profiles/topn_bench.py > $OUT/topn_bench.jsonl 2> $OUT/topn_bench.err; echo "topn bench rc=$?"; cat $OUT/topn_bench.jsonl | cut -c1-330
timeout 60 python profiles/topn_phase_profile.py > $OUT/topn_phase.jsonl 2>> $OUT/topn_bench.err; echo "phase rc=$?"; cat $OUT/topn_phase.jsonl
timeout 120 python profiles/kl_bench.py > $OUT/kl_bench.jsonl 2> $OUT/kl_bench.err; echo "kl bench rc=$?"; cat $OUT/kl_bench.jsonl
KL_REPS=1 timeout 200 ncu --set full --import-source on --clock-control none -k regex:softmax_kl_regs -c 2 -f -o $OUT/kl_regs \
  python profiles/kl_bench.py > $OUT/kl_ncu.log 2>&1; echo "ncu rc=$?"
if [ -f $OUT/kl_regs.ncu-rep ]; then
  ncu -i $OUT/kl_regs.ncu-rep --page raw --csv > $OUT/kl_regs_raw.csv 2>/dev/null
  ncu -i $OUT/kl_regs.ncu-rep --page source --csv --print-source cuda,sass --kernel-id :::2 > $OUT/kl_regs_source.csv 2>/dev/null
fi
ls -la $OUT
