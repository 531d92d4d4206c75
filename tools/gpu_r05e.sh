#!/bin/bash
# round r05e: patch embed variants (two CTAs per SM, + cluster multicast; dim 128 and 192): tests, A/B in the short-bench regime, bench
mkdir -p gpurun_out; L=gpurun_out/r05e.log; : > $L
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "embed" --timeout 120 -p no:cacheprovider 2>&1 | tail -6 | tee -a $L
timeout 150 python tools/probes/ab_probe.py embed_pair=0,1,2 n=20 rounds=5 cool=0.7 2>&1 | tee -a $L
timeout 150 python tools/probes/ab_probe.py model=FastTransformer frames=4 scale=2 embed_pair=0,1,2 n=10 rounds=4 cool=0.7 2>&1 | tee -a $L
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', round(d['value'],1), round(d['e2e']['value'],1), round(d['e2e']['bf16_host_tensors']['value'],1), d['clocks'])" | tee -a $L
TU_COMPUTE_STREAMS=1 timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench 1 compute stream', round(d['value'],1), round(d['e2e']['value'],1), round(d['e2e']['bf16_host_tensors']['value'],1), d['clocks'])" | tee -a $L
