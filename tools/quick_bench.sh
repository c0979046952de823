#!/bin/bash
# usage: tools/quick_bench.sh tag [workloads...]   (run on the GPU box via gpurun)
tag=$1; shift
for w in "$@"; do
  python bench.py --steps 300 --warmup 10 --no-cpu --no-e2e --workload $w > gpurun_out/bench_${tag}_$w.json 2>gpurun_out/bench_${tag}_$w.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${tag}_$w.json"))
    off = d.get("dr_off"); dl = d.get("dr_delay")
    print("$w", "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"],
          ("| dr_off %.4g frac %.3f" % (off["value"], off["roofline_frac"])) if off else "",
          ("| dr_delay %.4g frac %.3f" % (dl["value"], dl["roofline_frac"])) if dl else "")
except Exception as e:
    print("$w FAILED", e); print(open("gpurun_out/bench_${tag}_$w.err").read()[-2000:])
PY
done
