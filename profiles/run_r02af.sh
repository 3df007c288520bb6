#!/bin/bash
OUT=gpurun_out/r02af; mkdir -p $OUT
for v in "CC_NOOP=1" "CC_GEMM_MN3=0" "CC_SMALL_CHAIN=0"; do
  env $v timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --no-loss-check > $OUT/bench_$v.json 2> $OUT/bench_$v.err
  python -c "
import json;d=json.load(open('$OUT/bench_$v.json'))
e=d['extras']['ml_recommend']; print('$v', round(d['ms_per_step'],4), e['recs_per_s'], [round(x,4) for x in e['seconds_runs']], e['device_recs_per_s'])"
done
