#!/bin/bash
# ncu --set full of the kernels matching $3 inside one case of tools/bench_configs.py.  usage: gpu_ncu_cfg.sh <tag> <case substring> <regex> [skip] [count]
TAG=$1; CASE=$2; RE=$3; S=${4:-4}; N=${5:-2}
mkdir -p gpurun_out
timeout 600 python tools/bench_configs.py $TAG --only $CASE > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
cut -c1-600 gpurun_out/plain_$TAG.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $S -c $N -o gpurun_out/prof_$TAG -f \
    python tools/bench_configs.py $TAG --only $CASE > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_$TAG.ncu-rep
