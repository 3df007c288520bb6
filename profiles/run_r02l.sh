#!/bin/bash
# Round 2, call l: full GPU suite, KL kernel launch shapes, default bench line + reference arm, compute-sanitizer logs.
OUT=gpurun_out/r02l; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q --timeout=200 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log | cut -c1-300
timeout 120 python profiles/kl_bench.py > $OUT/kl_bench.jsonl 2> $OUT/kl_bench.err; echo "kl bench rc=$?"; cat $OUT/kl_bench.jsonl | cut -c1-260
timeout 400 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-600 $OUT/bench.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"; cut -c1-600 $OUT/bench_reference.json
timeout 600 compute-sanitizer --tool memcheck python profiles/sanitizer_driver.py > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $OUT/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python profiles/sanitizer_driver.py > $OUT/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 $OUT/sanitizer_racecheck.log
