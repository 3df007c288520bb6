#!/bin/bash
# Round 2, call ad: the round-end sequence on the final code + launch list + a targeted ncu capture of the step's GEMMs.
OUT=gpurun_out/r02ad; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q --timeout=200 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-260 $OUT/bench.json
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"; cut -c1-160 $OUT/bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-loss-check > $OUT/ncu1.log 2>&1
echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:"gemm_tc_kernel|chain_tc_kernel" -o $OUT/prof_gemms -f \
    python profiles/step_driver.py > $OUT/ncu2.log 2>&1
echo "gemm capture rc=$?"
if [ -f $OUT/prof_gemms.ncu-rep ]; then ncu -i $OUT/prof_gemms.ncu-rep --page raw --csv > $OUT/prof_gemms_raw.csv 2>/dev/null; rm -f $OUT/prof_gemms.ncu-rep; fi
ls $OUT
