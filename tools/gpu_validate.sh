#!/bin/bash
# usage: tools/gpu_validate.sh [tag]: full GPU parity suite, smoke, bench (+ reference arm), all BASELINE configs, ncu launch list
TAG=${1:-r05i}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_${TAG}.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" gpurun_out/pytest_gpu_${TAG}.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'], d['roofline']['kernel_ms_per_launch'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>/dev/null; cut -c 1-400 gpurun_out/bench_ref_${TAG}.json
timeout 600 python tools/bench_configs.py ${TAG} > gpurun_out/configs_${TAG}.log 2>&1; tail -12 gpurun_out/configs_${TAG}.log | cut -c 1-250
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_${TAG}.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "ncu launches rc=$?"
