#!/bin/bash
OUT=gpurun_out/r02ae; mkdir -p $OUT
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])
e=d['extras']['ml_recommend']; print(e['recs_per_s'], e['seconds_runs'], e['device_recs_per_s'])"
timeout 200 python profiles/ml_recommend_profile.py 2>&1 | cut -c1-330
