#!/bin/bash
# Round 2, call u: small-layer chains in the engine: full suite + A/B of the train step (chain on / off, tf32 and bf16).
OUT=gpurun_out/r02u; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q --timeout=200 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log | cut -c1-300
bash profiles/run_ab.sh r02u "CC_SMALL_CHAIN=0" "CC_PRECISION=bf16" "CC_PRECISION=bf16 CC_SMALL_CHAIN=0"
