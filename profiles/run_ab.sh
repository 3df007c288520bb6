#!/bin/bash
# A/B of the train step under environment switches, one bench line per variant (run under gpurun from the repo root):
#   gpurun --timeout 900 -- 'bash profiles/run_ab.sh r02f "CC_FIRST_LAYER=tensor" "CC_GEMM_STREAM_K=0" ...'
# The first line is always the default configuration.  Prints value / ms per step / per-kernel averages.
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 50 --warmup 5 --no-extras --no-cpu-baseline --no-loss-check $BENCH_ARGS \
      > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - "$OUT/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    k = {n: round(v["ms_total"] / v["launches"], 4) for n, v in d["kernels"].items()}
    print(f"{sys.argv[2]:34s} {d['value']/1e6:7.4f} M cubes/s  {d['ms_per_step']:.4f} ms  e2e {d['e2e']['value']/1e6:.4f}  {k}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run default CC_NOOP=1
i=0
for v in "$@"; do
  i=$((i+1))
  run "v${i}_$(echo $v | tr '= ' '__')" $v
done
