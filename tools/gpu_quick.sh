#!/bin/bash
# quick GPU round: tests (-x), bench, ncu launch list.  usage: tools/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-q}; KEXPR=${2:-}
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -k "$KEXPR" > gpurun_out/pytest_gpu_$TAG.log 2>&1
else
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1
fi
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" gpurun_out/pytest_gpu_$TAG.log | head -30
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launches rc=$?"
