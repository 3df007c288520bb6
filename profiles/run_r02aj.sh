#!/bin/bash
# Round 2, call aj (2 GPUs): sanity of the final code at N > 1 (dp_check in the default exchange modes + a short weak line).
OUT=gpurun_out/r02aj; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 120 $TR --master-port 29551 -m cubecobrarecommender_b200.dp_check --precision tf32 --steps 3 --modes p2p_overlap,p2p_unicast > $OUT/dp_check_tf32.json 2> $OUT/dp_check_tf32.err
echo "dp_check rc=$?"; python -c "
import json
d=json.load(open('$OUT/dp_check_tf32.json')); print('violations', d['violations'])
for m,r in d['modes'].items(): print(m, {k:r[k] for k in ('loss_rel_err','weights_max_abs_diff','replicas_bit_identical') if k in r})"
timeout 120 $TR --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > $OUT/bench_weak.json 2> $OUT/bench_weak.err; echo "weak rc=$?"
python -c "
import json; d=json.load(open('$OUT/bench_weak.json')); print(round(d['value']), round(d['ms_per_step'],4), d['config'].get('gradient_exchange'), d['clocks'])"
