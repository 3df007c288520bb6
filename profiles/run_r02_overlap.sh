#!/bin/bash
# p2p_overlap (bucket-wise exchange under backward) against p2p, on N GPUs:  bash profiles/run_r02_overlap.sh N
N=${1:-2}
OUT=gpurun_out/r02_overlap_n$N; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 -m cubecobrarecommender_b200.dp_check --precision tf32 --steps 3 --modes p2p_overlap,p2p_unicast \
  > $OUT/dp_check_overlap.json 2> $OUT/dp_check_overlap.err; echo "dp_check rc=$?"; tail -3 $OUT/dp_check_overlap.err | cut -c1-400
python -c "
import json
d=json.load(open('$OUT/dp_check_overlap.json')); print('violations', d['violations'])
for m,r in d['modes'].items(): print(m, {k:r[k] for k in ('loss_rel_err','weights_max_abs_diff','replicas_bit_identical','adam_m_max_rel_diff','multicast') if k in r})"
for mode in p2p p2p_overlap; do
  for sc in weak strong; do
    CC_DP_MODE=$mode timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 --scaling $sc --no-extras --no-cpu-baseline \
      > $OUT/bench_${mode}_$sc.json 2> $OUT/bench_${mode}_$sc.err
    python -c "
import json; d=json.load(open('$OUT/bench_${mode}_$sc.json')); print('$mode $sc', round(d['value']), round(d['ms_per_step'],4), {k: round(v['ms_total']/v['launches'],4) for k,v in d['kernels'].items()})"
  done
done
