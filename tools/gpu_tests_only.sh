#!/bin/bash
# GPU parity suite only (all failures listed).  usage: tools/gpu_tests_only.sh <tag> [pytest -k expression]
TAG=${1:-t}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider ${2:+-k "$2"} > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
grep -E "^FAILED|^ERROR|passed|failed|rc=|^E  " gpurun_out/pytest_gpu_$TAG.log | head -60
