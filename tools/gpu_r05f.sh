#!/bin/bash
# round r05f: patch embed with L2 prefetch of whole patch rows; three compute streams in the frame pipeline
mkdir -p gpurun_out; L=gpurun_out/r05f.log; : > $L
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "embed" --timeout 120 -p no:cacheprovider 2>&1 | tail -6 | tee -a $L
timeout 150 python tools/probes/ab_probe.py embed_pair=1,9,0 n=20 rounds=5 cool=0.7 2>&1 | tee -a $L
timeout 150 python tools/probes/ab_probe.py model=FastTransformer frames=4 scale=2 embed_pair=0,13 n=10 rounds=4 cool=0.7 2>&1 | tee -a $L
for cs in 3 2; do
TU_COMPUTE_STREAMS=$cs timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench $cs compute streams', round(d['value'],1), round(d['e2e']['value'],1), round(d['e2e']['bf16_host_tensors']['value'],1), d['clocks'])" | tee -a $L
done
