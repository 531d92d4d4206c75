#!/bin/bash
# ncu --set full of the kernels matching $2 (first forward after 3 warm-ups), after a plain run.  usage: gpu_ncu_one.sh <tag> <regex> [count]
TAG=$1; RE=$2; N=${3:-3}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $((3*N)) -c $N -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_$TAG.ncu-rep
