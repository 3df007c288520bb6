#!/bin/bash
# Round 2, call y: loss_finalize / first-layer colsum on the side stream: train-step tests + default bench line.
OUT=gpurun_out/r02y; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q --timeout=200 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-260 $OUT/bench.json
bash profiles/run_ab.sh r02y "CC_PRECISION=bf16"
