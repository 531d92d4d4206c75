#!/bin/bash
# one iteration on the window stack: its parity tests, the phase trace, and an A/B of the whole forward.  usage: tools/gpu_stack_iter.sh <tag> [fast]
TAG=${1:-s}
mkdir -p gpurun_out
tools/gpu_tests_only.sh $TAG "stack or window or transformer_block or golden or bf16"
python tools/probes/stack_phase_trace.py > gpurun_out/stack_phase_trace_$TAG.log 2>&1; grep -A14 "warp 0, first" gpurun_out/stack_phase_trace_$TAG.log
python tools/probes/ab_probe.py stack_var=0,0 n=20 rounds=4 cool=1 > gpurun_out/ab_$TAG.log 2>&1; tail -3 gpurun_out/ab_$TAG.log
if [ -n "$2" ]; then
python tools/probes/ab_probe.py stack_var=0,0 model=FastTransformer frames=4 scale=2 n=40 rounds=3 > gpurun_out/ab_${TAG}_fast.log 2>&1; tail -3 gpurun_out/ab_${TAG}_fast.log
python tools/probes/stack_phase_trace.py model=FastTransformer > gpurun_out/stack_phase_trace_${TAG}_fast.log 2>&1; grep -A13 "warp 0, first" gpurun_out/stack_phase_trace_${TAG}_fast.log
fi
