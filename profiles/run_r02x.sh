#!/bin/bash
# Round 2, call x: final code -- suite, smoke, default bench line + reference arm, configs[3] profile, full ncu capture of the step.
OUT=gpurun_out/r02x; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q --timeout=200 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-260 $OUT/bench.json
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"; cut -c1-160 $OUT/bench_reference.json
timeout 200 python profiles/ml_recommend_profile.py > $OUT/ml_recommend_profile.jsonl 2> $OUT/ml_recommend_profile.err; echo "recommend profile rc=$?"; cut -c1-420 $OUT/ml_recommend_profile.jsonl
python profiles/step_driver.py --graph --recommend > $OUT/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o $OUT/prof_step -f \
    python profiles/step_driver.py --graph --recommend > $OUT/ncu2.log 2>&1
echo "full capture rc=$?"
REP=$OUT/prof_step.ncu-rep
if [ -f $REP ]; then
  ncu -i $REP --page raw --csv > $OUT/prof_step_raw.csv 2>/dev/null
  rm -f $REP
fi
ls $OUT
