#!/bin/bash
# usage (8-GPU box, via gpurun --gpus 8): tools/scaling8.sh TAG -- the 8-GPU lines only (strong scaling of the 65536-env
# config as a Python loop and as one launch per 20 steps; weak scaling with the host-buffer arm): a refresh of scaling.sh's
# n = 8 rows for a later build
tag=$1; out=gpurun_out
common="--steps 400 --warmup 20 --no-cpu --no-named-sizes --no-dr-off"
run() {
  label=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 $common "$@" > $out/scale_${tag}_${label}_n8.json 2> $out/scale_${tag}_${label}_n8.err
  python - <<PY
import json
try:
    d=json.loads(open("$out/scale_${tag}_${label}_n8.json").read().strip().splitlines()[-1])
    e=d.get("e2e") or {}
    print("$label n=8", "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "frac/gpu %.3f" % d["roofline"]["frac"], "envs/gpu", d["config"]["envs_per_gpu"],
          ("| e2e %.4g link %s" % (e["value"], e.get("link_d2h_concurrent"))) if e else "")
except Exception as ex:
    print("$label FAILED", ex); print(open("$out/scale_${tag}_${label}_n8.err").read()[-1500:])
PY
}
run strong_graph --scaling strong --no-e2e --graph 20
run strong_loop --scaling strong --no-e2e
run weak
