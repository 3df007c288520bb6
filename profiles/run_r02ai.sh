#!/bin/bash
# Round 2, call ai: the default bench line + reference arm on the final code.
OUT=gpurun_out/r02ai; mkdir -p $OUT
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['traffic'], d['loss_rel_err'], d['gpu_launches'], d['clocks'])
print({k: round(v['ms_total']/v['launches'],4) for k,v in d['kernels'].items()})
e=d['extras']; print(e['ml_recommend']['recs_per_s'], e['ml_recommend']['device_recs_per_s'], e['ml_recommend']['select_roofline']['frac'], e['graph_build']['device_seconds'], e['graph_build']['count_roofline']['frac'])
r=json.load(open('$OUT/bench_reference.json')); print('reference', r['value'], r['cpu_baseline']['cores'])"
