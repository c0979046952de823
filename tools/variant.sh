#!/bin/bash
# usage (build container): tools/variant.sh NAME "-DFLAG=1 ..." [file.cu ...]
# Builds gpurun_variants/lib_NAME.so: the listed kernel TUs (default swarm_step_rot.cu) recompiled with the extra
# flags, linked against the objects of the in-tree build.  SWARM_B200_LIB=<that .so> selects it at run time.
set -e
name=$1; flags=$2; shift 2 || true
files=${@:-swarm_step_rot.cu}
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/multi-agent-rl-for-autonomous-drone-swarms_b200/csrc
out=$root/gpurun_variants; tmp=/tmp/swarm_variants/$name
mkdir -p $out $tmp
NVCCFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -I$root/include -I$src"
objs=""
for f in swarm_kernels.cu swarm_step_rot.cu swarm_step_rotx.cu swarm_abi.cu; do
  if echo " $files " | grep -q " $f "; then
    nvcc $NVCCFLAGS $flags -c $src/$f -o $tmp/${f%.cu}.o &
    objs="$objs $tmp/${f%.cu}.o"
  else
    objs="$objs $src/${f%.cu}.o"
  fi
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/lib_$name.so $objs -lcudart -ldl
echo "built $out/lib_$name.so ($flags)"
