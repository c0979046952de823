#!/bin/bash
for name in "$@"; do
  lib=""; [ "$name" != base ] && lib="SWARM_B200_LIB=$PWD/gpurun_variants/lib_$name.so"
  env $lib python bench.py --steps 50 --warmup 5 --no-cpu --no-dr-off --no-named-sizes --e2e-steps 24 > gpurun_out/e2e_$name.json 2> gpurun_out/e2e_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/e2e_$name.json")); e=d["e2e"]
    print("%-8s e2e %.4g  ms %.3f  d2h %.1f GB/s  link %.1f | lean %.4g" % ("$name", e["value"], e["ms_per_step"], e["d2h_gbs"], e["link_d2h_gbs_measured"], e["variants"]["lean_obs_reward_flags"]["value"]), flush=True)
except Exception as ex:
    print("$name FAILED", ex)
PY
done
