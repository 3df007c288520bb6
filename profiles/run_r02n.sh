#!/bin/bash
# Round 2, call n (2 GPUs): register-form row select v2 (tests, bench, phase profile) + p2p_overlap against p2p at N=2.
OUT=gpurun_out/r02n; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q --timeout=200 -k "topn or select or recommend or full_identity or first_layer" > $OUT/pytest_topn.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_topn.log | cut -c1-300
timeout 120 python profiles/topn_bench.py > $OUT/topn_bench.jsonl 2> $OUT/topn_bench.err; echo "topn bench rc=$?"; grep -E '"batch": 4096' $OUT/topn_bench.jsonl | cut -c1-260
timeout 60 python profiles/topn_phase_profile.py > $OUT/topn_phase.jsonl 2>> $OUT/topn_bench.err; echo "phase rc=$?"; cat $OUT/topn_phase.jsonl | cut -c1-600
bash profiles/run_r02_overlap.sh 2
