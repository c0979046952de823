#!/bin/bash
# usage (8-GPU box, via gpurun --gpus 8): tools/scaling.sh TAG -- strong-scaling table of BASELINE config 4 as written
# (65536 env instances IN TOTAL over N GPUs), as a Python loop over step() and as CUDA-graph replays of 20 steps, plus
# the weak-scaling line at 8 GPUs with the host-buffer arm and the concurrent bare-copy ceiling.
tag=$1; out=gpurun_out
common="--steps 400 --warmup 20 --no-cpu --no-named-sizes --no-dr-off"
run() {  # n, extra flags, label
  n=$1; shift; label=$1; shift
  if [ $n = 1 ]; then python bench.py --gpus 1 $common "$@" > $out/scale_${tag}_${label}_n$n.json 2> $out/scale_${tag}_${label}_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n $common "$@" > $out/scale_${tag}_${label}_n$n.json 2> $out/scale_${tag}_${label}_n$n.err; fi
  python - <<PY
import json
try:
    d=json.loads(open("$out/scale_${tag}_${label}_n$n.json").read().strip().splitlines()[-1])
    e=d.get("e2e") or {}
    print("$label n=$n", "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "frac/gpu %.3f" % d["roofline"]["frac"], "envs/gpu", d["config"]["envs_per_gpu"],
          ("| e2e %.4g link %s" % (e["value"], e.get("link_d2h_concurrent"))) if e else "")
except Exception as ex:
    print("$label n=$n FAILED", ex); print(open("$out/scale_${tag}_${label}_n$n.err").read()[-1500:])
PY
}
for n in 1 2 4 8; do run $n strong_loop --scaling strong --no-e2e; done
for n in 1 2 4 8; do run $n strong_graph --scaling strong --no-e2e --graph 20; done
run 8 weak
run 1 weak
nvidia-smi topo -m > $out/scale_${tag}_topo.txt 2>&1
