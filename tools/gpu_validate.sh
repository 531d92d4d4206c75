#!/bin/bash
# usage: tools/gpu_validate.sh  (tag r05i): full GPU parity suite, smoke, bench (+ reference arm), all BASELINE configs, ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu_r05i.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r05i.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" gpurun_out/pytest_gpu_r05i.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r05i.json 2> gpurun_out/bench_r05i.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r05i.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'], d['roofline']['kernel_ms_per_launch'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r05i.json 2>/dev/null; cut -c 1-400 gpurun_out/bench_ref_r05i.json
timeout 600 python tools/bench_configs.py r05i > gpurun_out/configs_r05i.log 2>&1; tail -12 gpurun_out/configs_r05i.log | cut -c 1-250
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_r05i.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r05i.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_r05i.log 2>&1
echo "ncu launches rc=$?"
