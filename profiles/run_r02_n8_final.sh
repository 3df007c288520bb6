#!/bin/bash
# Round 2, 8 GPUs, final code: the driver-style weak-scaling lines (tf32 default exchange mode, bf16).
OUT=gpurun_out/r02_n8_final; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29531 bench.py --gpus 8 --steps 50 --warmup 5 --no-extras --no-cpu-baseline > $OUT/bench_weak_tf32.json 2> $OUT/bench_weak_tf32.err; echo "weak rc=$?"
timeout 300 $TR --master-port 29532 bench.py --gpus 8 --steps 50 --warmup 5 --precision bf16 --no-extras --no-cpu-baseline > $OUT/bench_weak_bf16.json 2> $OUT/bench_weak_bf16.err; echo "weak bf16 rc=$?"
for f in bench_weak_tf32 bench_weak_bf16; do python -c "
import json; d=json.load(open('$OUT/$f.json')); print('$f', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['config'].get('gradient_exchange'), d['dtype'], d['clocks']['reasons'])"; done
