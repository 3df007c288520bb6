#!/bin/bash
# Round-1 profiling recipe (run under gpurun from the repo root); outputs land in gpurun_out/.
OUT=gpurun_out
TAG=${1:-r01}
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "ref rc=$?"
# every launch of the bench command with its device time
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu1.log 2>&1
echo "launch list rc=$?"
# one warm train step + one graph build + one batched top-50 call, full metric set
python profiles/step_driver.py --graph --recommend > $OUT/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -o $OUT/${TAG}_prof_step -f \
    python profiles/step_driver.py --graph --recommend > $OUT/ncu2.log 2>&1
echo "full capture rc=$?"
REP=$OUT/${TAG}_prof_step.ncu-rep
if [ -f $REP ]; then
  ncu -i $REP --page raw --csv > $OUT/${TAG}_prof_step_raw.csv 2>/dev/null
  ncu -i $REP --page source --csv -k regex:gemm_tc_kernelILi0ELi1 -c 1 > $OUT/${TAG}_src_gemm_bce.csv 2>/dev/null
  ls -la $REP
  [ $(stat -c %s $REP) -gt 40000000 ] && rm -f $REP
fi
ls -la $OUT
