#!/bin/bash
mkdir -p gpurun_out
for k in fused_window_stack; do
  timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 120 -p no:cacheprovider -k "$k" > gpurun_out/dbg_$k.log 2>&1
  echo "$k rc=$? $(tail -1 gpurun_out/dbg_$k.log)"; grep -E "^E  " gpurun_out/dbg_$k.log | head -8
done
