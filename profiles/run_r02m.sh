#!/bin/bash
# Round 2, call m: suite after the logit-bound fix, configs[3] taken apart, launch list + full ncu capture of the step.
OUT=gpurun_out/r02m; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q --timeout=200 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log | cut -c1-300
timeout 200 python profiles/ml_recommend_profile.py > $OUT/ml_recommend_profile.jsonl 2> $OUT/ml_recommend_profile.err; echo "recommend profile rc=$?"; cut -c1-700 $OUT/ml_recommend_profile.jsonl
timeout 200 python profiles/ml_recommend_profile.py --trained > $OUT/ml_recommend_profile_trained.jsonl 2>> $OUT/ml_recommend_profile.err; echo "recommend profile (spread logits) rc=$?"; cut -c1-700 $OUT/ml_recommend_profile_trained.jsonl
timeout 120 python profiles/topn_bench.py > $OUT/topn_bench.jsonl 2> $OUT/topn_bench.err; echo "topn bench rc=$?"; grep -E '"batch": 4096' $OUT/topn_bench.jsonl | cut -c1-200
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-loss-check > $OUT/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-loss-check > $OUT/ncu1.log 2>&1
echo "launch list rc=$?"
python profiles/step_driver.py --graph --recommend > $OUT/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o $OUT/prof_step -f \
    python profiles/step_driver.py --graph --recommend > $OUT/ncu2.log 2>&1
echo "full capture rc=$?"
REP=$OUT/prof_step.ncu-rep
if [ -f $REP ]; then
  ncu -i $REP --page raw --csv > $OUT/prof_step_raw.csv 2>/dev/null
  ls -la $REP
  [ $(stat -c %s $REP) -gt 40000000 ] && rm -f $REP
fi
ls -la $OUT
