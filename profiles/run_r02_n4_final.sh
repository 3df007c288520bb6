#!/bin/bash
# Round 2, 4 GPUs, final code: dp_check (auto-relevant modes) + the driver-style weak line.
OUT=gpurun_out/r02_n4_final; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 -m cubecobrarecommender_b200.dp_check --precision tf32 --steps 3 --modes p2p_multicast,p2p_overlap > $OUT/dp_check_tf32.json 2> $OUT/dp_check_tf32.err
echo "dp_check rc=$?"; python -c "
import json
d=json.load(open('$OUT/dp_check_tf32.json')); print('violations', d['violations'])
for m,r in d['modes'].items(): print(m, {k:r[k] for k in ('loss_rel_err','weights_max_abs_diff','replicas_bit_identical','multicast') if k in r})"
timeout 300 $TR --master-port 29542 bench.py --gpus 4 --steps 20 --warmup 5 > $OUT/bench_weak_tf32.json 2> $OUT/bench_weak_tf32.err; echo "weak rc=$?"
python -c "
import json; d=json.load(open('$OUT/bench_weak_tf32.json')); print(round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['config'].get('gradient_exchange'), d['clocks']['reasons'], 'cpu_baseline' in d and d['cpu_baseline'] is not None)"
