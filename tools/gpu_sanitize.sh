#!/bin/bash
# compute-sanitizer (ONE tool per call, B200_PROFILING.md) on the small-shape tests of the kernels added in round 2.
# usage: tools/gpu_sanitize.sh <memcheck|racecheck|synccheck> <tag> [pytest -k expression]
TOOL=${1:-memcheck}; TAG=${2:-r2}
mkdir -p gpurun_out
K=${3:-'frames_to_planar or interleaved_uint8 or uint8_frames_on_fast_resize or (global_attention_tcgen05 and 328) or (global_attention_tcgen05 and 1-128) or c_packer'}
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 9 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -q -m gpu -x -p no:cacheprovider -k "$K" > gpurun_out/sanitizer_${TOOL}_$TAG.log 2>&1
echo "sanitizer($TOOL) rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Error|RACECHECK SUMMARY|hazard" gpurun_out/sanitizer_${TOOL}_$TAG.log | head -20
