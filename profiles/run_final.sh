#!/bin/bash
# End-of-session evidence: GPU test suite, the default bench line, the reference arm, the launch list of the bench command.
OUT=gpurun_out
TAG=${1:-r01d}
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "ref rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $OUT/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu1.log 2>&1
echo "launch list rc=$?"
