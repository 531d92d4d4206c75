#!/bin/bash
# one `ncu --set full` capture of every kernel of ONE forward (the 4th: after 3 warm-up forwards), only after the same
# command has exited 0 without ncu.  usage: tools/gpu_ncu_full.sh <tag>
TAG=${1:-r}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"bicubic_add|conv3x3_tc|gemm_tc|stem_tc|window_stack" -s 27 -c 9 -o gpurun_out/prof_forward_$TAG -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/prof_forward_$TAG.ncu-rep
