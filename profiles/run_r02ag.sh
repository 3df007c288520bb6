#!/bin/bash
# Round 2, call ag: 3-D maps for the unpadded Keras kernels (readable-range registration): suite + default bench.
OUT=gpurun_out/r02ag; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q --timeout=200 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['achieved'], d['loss_rel_err'], d['clocks'])
print({k: round(v['ms_total']/v['launches'],4) for k,v in d['kernels'].items()})
e=d['extras']['ml_recommend']; print(e['recs_per_s'], e['seconds_runs'], e['device_recs_per_s'])"
