#!/bin/bash
# First GPU call of round 2 (run under gpurun from the repo root, ~3 min of box time; outputs land in gpurun_out/):
#   gpurun --timeout 400 -- 'bash profiles/run_round2_first.sh r02a'
# 1. the GPU suite as shipped, 2. the top-N tests again with the opt-in NEXT revision of the row select (algo 4, never run
# on a GPU in round 1), 3. the select micro-benchmark and the per-phase cycle profile of every variant, 4. the default
# bench line (its extras now carry the select's own roofline), 5. an ncu --set full capture of the select kernels.
OUT=gpurun_out
TAG=${1:-r02a}
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/${TAG}_pytest_gpu.log
CC_TOPN_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_graph.py -k topn -x -q > $OUT/${TAG}_pytest_topn_next.log 2>&1
echo "topn tests with algo 4 rc=$?"; tail -3 $OUT/${TAG}_pytest_topn_next.log
CC_TOPN_EXPERIMENTAL=1 timeout 60 python profiles/topn_bench.py > $OUT/${TAG}_topn_bench.jsonl 2> $OUT/${TAG}_topn_bench.err
echo "topn bench rc=$?"; cat $OUT/${TAG}_topn_bench.jsonl
CC_TOPN_EXPERIMENTAL=1 timeout 60 python profiles/topn_phase_profile.py > $OUT/${TAG}_topn_phase.jsonl 2>> $OUT/${TAG}_topn_bench.err
echo "phase profile rc=$?"; cat $OUT/${TAG}_topn_phase.jsonl
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
CC_TOPN_EXPERIMENTAL=1 TOPN_REPS=1 TOPN_BATCHES=4096 timeout 150 ncu --set full --import-source on --clock-control none \
    -k regex:topn_rowselect -c 6 -f -o $OUT/${TAG}_topn_rs python profiles/topn_bench.py > $OUT/${TAG}_topn_ncu.log 2>&1
echo "ncu rc=$?"
[ -f $OUT/${TAG}_topn_rs.ncu-rep ] && ncu -i $OUT/${TAG}_topn_rs.ncu-rep --page raw --csv > $OUT/${TAG}_topn_rs_raw.csv 2>/dev/null
ls -la $OUT | tail -12
# (separately, on 8 GPUs:  gpurun --gpus 8 --timeout 300 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
#   --master-addr 127.0.0.1 --master-port 29511 profiles/scale_up_8gpu.py > gpurun_out/r02_scale_up_8gpu.json')
# (and configs[3] on 8 GPUs:  gpurun --gpus 8 --timeout 300 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
#   --master-addr 127.0.0.1 --master-port 29512 profiles/ml_recommend_ngpu.py > gpurun_out/r02_ml_recommend_8gpu.json')
