#!/bin/bash
OUT=gpurun_out/r02ah; mkdir -p $OUT
for v in "CC_NOOP=1" "CC_GEMM_READABLE=0"; do
  env $v timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-loss-check > $OUT/bench_$v.json 2> $OUT/bench_$v.err
  python -c "
import json;d=json.load(open('$OUT/bench_$v.json'))
print('$v', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['roofline']['instrumented_ms_per_step'],4), {k: round(v['ms_total']/v['launches'],4) for k,v in d['kernels'].items()}, d['clocks']['samples'])"
done
