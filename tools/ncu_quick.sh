#!/bin/bash
# usage (GPU box): tools/ncu_quick.sh TAG "ENV=..,ENV=.." [bench flags...]   -- a few counters of the step kernels, CSV
tag=$1; envs=$2; shift 2
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__icc_request_hit_rate.pct,sm__cycles_active.avg,gpc__cycles_elapsed.max,launch__registers_per_thread,launch__grid_size,smsp__warps_active.avg.per_cycle_active,smsp__warps_eligible.avg.per_cycle_active
for st in no_instruction long_scoreboard wait not_selected short_scoreboard math_pipe_throttle mio_throttle lg_throttle branch_resolving dispatch_stall imc_miss barrier; do
  M=$M,smsp__average_warps_issue_stalled_${st}_per_issue_active.ratio
done
M=$M,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
envcmd="env"; [ -n "$envs" ] && envcmd="env ${envs//,/ }"
$envcmd ncu --metrics $M --clock-control none -k regex:swarm_step_rot -s 10 -c 4 --csv --log-file gpurun_out/nq_$tag.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --no-dr-off "$@" > gpurun_out/nq_$tag.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/nq_$tag.csv")) if len(r)>5]
h=rows[0]
out={}
for r in rows[1:]:
    d=dict(zip(h,r)); out.setdefault((d["ID"],d["Kernel Name"][:48]),{})[d["Metric Name"]]=d["Metric Value"]
for (i,k),m in list(out.items())[:2]:
    print("$tag", k)
    for name,v in m.items():
        print("   %-86s %s" % (name, v))
PY
