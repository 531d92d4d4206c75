#!/bin/bash
# ncu --set full of kernels matching a regex for an arbitrary python command (plain run first).
# usage: tools/gpu_ncu_cmd.sh <tag> <kernel regex> <skip> <count> <python args...>
TAG=$1; RE=$2; SKIP=$3; N=$4; shift 4
mkdir -p gpurun_out
timeout 600 python "$@" > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
tail -2 gpurun_out/plain_$TAG.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $N -o gpurun_out/prof_$TAG -f \
    python "$@" > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_$TAG.ncu-rep
