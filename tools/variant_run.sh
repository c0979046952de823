#!/bin/bash
# usage (GPU box, via gpurun): tools/variant_run.sh [--parity] STEPS WORKLOAD[:extra bench flags] NAME...
# Benches gpurun_variants/lib_NAME.so ("base" = the in-tree library) and prints one line per variant.
parity=0; [ "$1" = "--parity" ] && { parity=1; shift; }
steps=$1; wl=$2; shift 2
w=${wl%%:*}; extra=""; [ "$wl" != "$w" ] && extra=${wl#*:}
for name in "$@"; do
  lib=""; envs=""
  n=$name
  case $name in *@*) envs=${name#*@}; n=${name%%@*};; esac   # NAME@VAR=VALUE,VAR=VALUE
  [ "$n" != base ] && lib=$PWD/gpurun_variants/lib_$n.so
  envcmd="env"; [ -n "$lib" ] && envcmd="$envcmd SWARM_B200_LIB=$lib"
  [ -n "$envs" ] && envcmd="$envcmd ${envs//,/ }"
  if [ $parity = 1 ]; then
    $envcmd python -m pytest tests/test_gpu_parity.py tests/test_domain_randomization.py -m gpu -x -q 2>&1 | tail -2 | tr '\n' ' '
  fi
  $envcmd python bench.py --steps $steps --warmup 10 --no-cpu --no-e2e --workload $w $extra > gpurun_out/vb_${name//[@=,]/_}_$w.json 2> gpurun_out/vb_${name//[@=,]/_}_$w.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/vb_${name//[@=,]/_}_$w.json"))
    off=d.get("dr_off")
    print("%-28s %s %.4g ms %.4f frac %.3f" % ("$name", "$w", d["value"], d["ms_per_step"], d["roofline"]["frac"]),
          ("| dr_off %.4g ms %.4f frac %.3f" % (off["value"], off["ms_per_step"], off["roofline_frac"])) if off else "", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], flush=True)
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/vb_${name//[@=,]/_}_$w.err").read()[-1500:])
PY
done
