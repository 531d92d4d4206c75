#!/bin/bash
# round r05b: new kernels (bicubic pair, embed pair, stack split): targeted tests, A/B of each switch, bench, compute streams
mkdir -p gpurun_out; L=gpurun_out/r05b.log; : > $L
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "bicubic or embed or stack or dec12" --timeout 120 -p no:cacheprovider 2>&1 | tail -6 | tee -a $L
timeout 300 python -m pytest tests/test_gpu_models.py -q -m gpu -x -k "split or overlap or golden" --timeout 200 -p no:cacheprovider 2>&1 | tail -6 | tee -a $L
for sw in bicubic_pair embed_pair stack_split; do
  timeout 150 python tools/probes/ab_probe.py $sw=0,1 n=20 rounds=5 2>&1 | tee -a $L
done
timeout 150 python tools/probes/ab_probe.py model=FastTransformer frames=4 scale=2 stack_split=0,1 n=10 rounds=4 2>&1 | tee -a $L
for cs in 1 2 1 2; do
  TU_COMPUTE_STREAMS=$cs timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('streams $cs', round(d['value'],1), round(d['e2e']['value'],1), round(d['e2e']['bf16_host_tensors']['value'],1), d['roofline']['kernel_ms_per_launch'])" | tee -a $L
done
