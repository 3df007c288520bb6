#!/bin/bash
# Multi-GPU evidence of round 2 (run under `gpurun --gpus N`):  bash profiles/run_r02_ngpu.sh N [small]
#  1. dp_check: N ranks on a fixed global batch reproduce the 1-GPU step (p2p unicast, multicast, nccl)
#  2. configs[4] scale-up build, all_reduce and reduce-scatter forms
#  3. configs[3] batched top-50 at 1, 2, ... N GPUs
#  4. bench lines: strong scaling (global batch 4096) in tf32, and BASELINE configs[2] as written (bf16, full-I, strong)
N=${1:-8}; SMALL=$2
OUT=gpurun_out/r02_n$N; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ -n "$SMALL" ]; then export SCALE_UP_CUBES=65536 SCALE_UP_CARDS=8192 RECOMMEND_CUBES=8192; fi
timeout 300 $TR --master-port 29501 -m cubecobrarecommender_b200.dp_check --precision tf32 --steps 3 --modes p2p_unicast,p2p_multicast,nccl,p2p_overlap > $OUT/dp_check_tf32.json 2> $OUT/dp_check_tf32.err
echo "dp_check rc=$?"; python -c "
import json,sys
d=json.load(open('$OUT/dp_check_tf32.json')); print('violations', d['violations'])
for m,r in d['modes'].items(): print(m, {k:r[k] for k in ('loss_rel_err','weights_max_abs_diff','replicas_bit_identical','multicast') if k in r})"
timeout 300 $TR --master-port 29502 profiles/scale_up_8gpu.py > $OUT/scale_up.json 2> $OUT/scale_up.err; echo "scale_up rc=$?"; cat $OUT/scale_up.json | cut -c1-1500
timeout 300 $TR --master-port 29503 profiles/ml_recommend_ngpu.py > $OUT/ml_recommend.jsonl 2> $OUT/ml_recommend.err; echo "ml_recommend rc=$?"; cat $OUT/ml_recommend.jsonl | cut -c1-300
if [ -z "$SMALL" ]; then
  timeout 300 $TR --master-port 29504 bench.py --gpus $N --steps 50 --warmup 5 --scaling strong --no-extras --no-cpu-baseline > $OUT/bench_strong_tf32.json 2> $OUT/bench_strong_tf32.err; echo "strong rc=$?"
  timeout 300 $TR --master-port 29505 bench.py --gpus $N --steps 50 --warmup 5 --scaling strong --precision bf16 --reg-mode full --no-extras --no-cpu-baseline > $OUT/bench_strong_bf16_full.json 2> $OUT/bench_strong_bf16_full.err; echo "bf16 full rc=$?"
  timeout 300 $TR --master-port 29506 bench.py --gpus $N --steps 50 --warmup 5 --no-extras --no-cpu-baseline > $OUT/bench_weak_tf32.json 2> $OUT/bench_weak_tf32.err; echo "weak rc=$?"
  CC_DP_MODE=p2p_overlap timeout 300 $TR --master-port 29507 bench.py --gpus $N --steps 50 --warmup 5 --no-extras --no-cpu-baseline > $OUT/bench_weak_tf32_overlap.json 2> $OUT/bench_weak_tf32_overlap.err; echo "weak overlap rc=$?"
  timeout 300 $TR --master-port 29508 bench.py --gpus $N --steps 50 --warmup 5 --precision bf16 --no-extras --no-cpu-baseline > $OUT/bench_weak_bf16.json 2> $OUT/bench_weak_bf16.err; echo "weak bf16 rc=$?"
  for f in bench_strong_tf32 bench_strong_bf16_full bench_weak_tf32 bench_weak_tf32_overlap bench_weak_bf16; do python -c "
import json; d=json.load(open('$OUT/$f.json')); print('$f', d['value'], d['ms_per_step'], d['config']['batch_per_gpu'], d['config']['reg_rows_per_gpu'], d['scaling'], d['dtype'])"; done
fi
ls $OUT
