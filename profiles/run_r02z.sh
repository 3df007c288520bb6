#!/bin/bash
# Round 2, call z (2 GPUs): the data-parallel paths on the final code (small-layer chains, side-stream helpers).
OUT=gpurun_out/r02z; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout=400 > $OUT/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 $OUT/pytest_multi.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 -m cubecobrarecommender_b200.dp_check --precision tf32 --steps 3 --modes p2p_unicast,p2p_multicast,nccl,p2p_overlap > $OUT/dp_check_tf32.json 2> $OUT/dp_check_tf32.err
echo "dp_check rc=$?"; python -c "
import json
d=json.load(open('$OUT/dp_check_tf32.json')); print('violations', d['violations'])
for m,r in d['modes'].items(): print(m, {k:r[k] for k in ('loss_rel_err','weights_max_abs_diff','replicas_bit_identical','multicast') if k in r})"
for sc in weak strong; do
  timeout 300 $TR --master-port 29522 bench.py --gpus 2 --steps 50 --warmup 5 --scaling $sc --no-extras --no-cpu-baseline > $OUT/bench_$sc.json 2> $OUT/bench_$sc.err
  python -c "
import json; d=json.load(open('$OUT/bench_$sc.json')); print('$sc', round(d['value']), round(d['ms_per_step'],4), d['config'].get('gradient_exchange'), {k: round(v['ms_total']/v['launches'],4) for k,v in d['kernels'].items()})"
done
