#!/bin/bash
# usage: ms_run.sh NAME...   (one-launch multi-step mode at 8192 envs, DR off and on, and C3 at 16384)
for name in "$@"; do
  lib=""; [ "$name" != base ] && lib="SWARM_B200_LIB=$PWD/gpurun_variants/lib_$name.so"
  for cfg in "c4 8192 off" "c4 8192 on" "c3 16384 off"; do
    set -- $cfg
    env $lib python bench.py --steps 400 --warmup 20 --no-cpu --no-e2e --no-dr-off --no-named-sizes --workload $1 --envs-per-gpu $2 --dr $3 --graph 20 > gpurun_out/ms_${name}_$1_$3.json 2> gpurun_out/ms_${name}_$1_$3.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ms_${name}_$1_$3.json"))
    print("%-8s %s %s dr=%s  %.2f us/step frac %.3f" % ("$name", "$1", "$2", "$3", d["ms_per_step"]*1e3, d["roofline"]["frac"]), flush=True)
except Exception as e:
    print("$name FAILED", e)
PY
  done
done
