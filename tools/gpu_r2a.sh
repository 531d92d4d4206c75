#!/bin/bash
# round-2 validation: full GPU parity suite, smoke, bench (+ reference arm)
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
grep -E "^FAILED|^ERROR|passed|failed|rc=|Error|assert" gpurun_out/pytest_gpu_$TAG.log | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'], d['roofline']['kernel_ms_per_launch'], d['config'].get('sustained'), d['roofline']['memory_bound_kernels'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cut -c 1-300 gpurun_out/bench_ref_$TAG.json
