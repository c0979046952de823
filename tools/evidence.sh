#!/bin/bash
# usage (on the GPU box, via gpurun): tools/evidence.sh TAG  -- bench lines, launch list and ncu captures for profiles/
tag=$1
out=gpurun_out
tools/quick_bench.sh warm c4 >/dev/null 2>&1
python bench.py > $out/bench_full_$tag.json 2> $out/bench_full_$tag.err
python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
tools/quick_bench.sh $tag c2 c3 c1 c5 c5_w70 c4_w44 phys phys8
# round 2b: randomisation above 32 drones and on the physics env
python bench.py --steps 300 --warmup 10 --no-cpu --no-e2e --no-dr-off --workload c5_w70 --dr on > $out/bench_${tag}_c5_w70_dr.json 2> $out/bench_${tag}_c5_w70_dr.err
python bench.py --steps 300 --warmup 10 --no-cpu --no-e2e --no-dr-off --workload c5 --dr on > $out/bench_${tag}_c5_dr.json 2> $out/bench_${tag}_c5_dr.err
python bench.py --steps 300 --warmup 10 --no-cpu --no-e2e --no-dr-off --workload phys8 --dr on > $out/bench_${tag}_phys8_dr.json 2> $out/bench_${tag}_phys8_dr.err
python tools/facade_latency.py > $out/facade_latency_$tag.json 2> $out/facade_latency_$tag.err
for dr in on off; do
  CMD="python bench.py --steps 12 --warmup 4 --no-cpu --no-e2e --no-dr-off --no-named-sizes --dr $dr"
  $CMD > $out/plain_${tag}_$dr.log 2>&1 || { echo "plain run failed ($dr)"; continue; }
  if [ $dr = on ]; then
    ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_$tag.csv $CMD > $out/ncu_${tag}_list.log 2>&1
  fi
  ncu --set full --clock-control none --import-source on -k regex:swarm_step_rot_kernel -s 16 -c 2 -f -o $out/prof_c4_${tag}_dr$dr $CMD > $out/ncu_${tag}_$dr.log 2>&1
done
CMD="python bench.py --workload c5 --steps 12 --warmup 4 --no-cpu --no-e2e"
$CMD > $out/plain_${tag}_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:swarm_step_rotx -s 16 -c 2 -f -o $out/prof_c5_$tag $CMD > $out/ncu_${tag}_c5.log 2>&1
ls -la $out/*$tag* | awk '{print $5, $9}'
