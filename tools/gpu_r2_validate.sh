#!/bin/bash
# round-2 validation: GPU parity suite, smoke, bench (+ reference arm), parity report at BASELINE sizes, every BASELINE config on one GPU,
# attention probes, ncu launch list.   usage: tools/gpu_r2_validate.sh <tag>
TAG=${1:-r2v}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" gpurun_out/pytest_gpu_$TAG.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('sustained'), d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'], d['roofline']['kernel_ms_per_launch'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cut -c 1-200 gpurun_out/bench_ref_$TAG.json
timeout 900 python tools/parity_report.py $TAG 2>&1 | cut -c1-420
timeout 600 python tools/bench_configs.py $TAG > gpurun_out/configs_$TAG.log 2>&1; tail -12 gpurun_out/configs_$TAG.log | cut -c 1-330
for b in 1 2 16; do timeout 120 python tools/probes/attn_probe.py $b 3600 | tail -1; done | tee gpurun_out/attn_probe_$TAG.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sustained > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sustained > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launches rc=$?"
