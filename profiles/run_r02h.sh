#!/bin/bash
OUT=gpurun_out/r02h; mkdir -p $OUT
timeout 500 python -m pytest tests -m gpu -x -q --timeout=200 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log | cut -c1-300
timeout 120 python profiles/topn_bench.py > $OUT/topn_bench.jsonl 2> $OUT/topn_bench.err; echo "topn bench rc=$?"; grep -E "4096" $OUT/topn_bench.jsonl | cut -c1-200
timeout 60 python profiles/topn_phase_profile.py > $OUT/topn_phase.jsonl 2>> $OUT/topn_bench.err; echo "phase rc=$?"; head -2 $OUT/topn_phase.jsonl
timeout 120 python profiles/kl_bench.py > $OUT/kl_bench.jsonl 2> $OUT/kl_bench.err; echo "kl bench rc=$?"; cat $OUT/kl_bench.jsonl | cut -c1-260
bash profiles/run_ab.sh r02h "CC_PRECISION=bf16"
