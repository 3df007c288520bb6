#!/bin/bash
# Round 2, call ac: 3-D TMA maps for MN-major tf32 operands: GEMM / chain tests, layout micro-benchmark (3-D on / off), step A/B.
OUT=gpurun_out/r02ac; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q --timeout=200 -k "gemm or chain or stream_k or bce" > $OUT/pytest_gemm.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gemm.log | cut -c1-300
timeout 100 python profiles/gemm_layout_bench.py > $OUT/gemm_layout_mn3.jsonl 2> $OUT/gemm_layout.err; echo "layout rc=$?"
CC_GEMM_MN3=0 timeout 100 python profiles/gemm_layout_bench.py > $OUT/gemm_layout_boxes.jsonl 2>> $OUT/gemm_layout.err
python - <<'PY'
import json
a=[json.loads(l) for l in open('gpurun_out/r02ac/gemm_layout_mn3.jsonl')]; b=[json.loads(l) for l in open('gpurun_out/r02ac/gemm_layout_boxes.jsonl')]
for x,y in zip(a,b):
    if x['precision']=='tf32': print(f"{x['case'][:72]:72s} 3-D {x['ms']*1000:7.1f} us   boxes {y['ms']*1000:7.1f} us   (3-D operands so far {x['operands_through_3d_maps_so_far']})")
PY
bash profiles/run_ab.sh r02ac "CC_GEMM_MN3=0"
