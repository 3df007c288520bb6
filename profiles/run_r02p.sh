#!/bin/bash
# Round 2, call p: register-form row select v3 (tests, bench, phase profile), the GEMM stream-K / tile-order tests.
OUT=gpurun_out/r02p; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q --timeout=200 -k "topn or select or recommend or stream_k or full_identity" > $OUT/pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_sel.log | cut -c1-300
timeout 120 python profiles/topn_bench.py > $OUT/topn_bench.jsonl 2> $OUT/topn_bench.err; echo "topn bench rc=$?"; grep -E '"batch": 4096' $OUT/topn_bench.jsonl | cut -c1-260
timeout 60 python profiles/topn_phase_profile.py > $OUT/topn_phase.jsonl 2>> $OUT/topn_bench.err; echo "phase rc=$?"; head -1 $OUT/topn_phase.jsonl | cut -c1-600
