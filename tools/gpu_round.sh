#!/bin/bash
# One GPU round: parity tests, bench line, ncu launch list and a full capture of the dominant kernel.
# usage: tools/gpu_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cat gpurun_out/bench_ref_$TAG.json
# launch list (only after the plain run above exited 0)
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 3 -c 1 -o gpurun_out/prof_conv_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
