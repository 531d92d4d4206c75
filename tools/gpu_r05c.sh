#!/bin/bash
# round r05c: split tests incl. FastTransformer, A/B of the round's switches in the regime of a short bench run (idle gaps), ncu of the pair bicubic kernel
mkdir -p gpurun_out; L=gpurun_out/r05c.log; : > $L
timeout 300 python -m pytest tests/test_gpu_models.py -q -m gpu -x -k "split or overlap" --timeout 200 -p no:cacheprovider 2>&1 | tail -4 | tee -a $L
for sw in stack_split embed_pair bicubic_pair fuse_dec12; do
  timeout 150 python tools/probes/ab_probe.py $sw=0,1 n=20 rounds=6 cool=0.7 2>&1 | grep -v "^    " | tee -a $L
done
timeout 150 python tools/probes/ab_probe.py model=FastTransformer frames=4 scale=2 stack_split=0,1 n=10 rounds=5 cool=0.7 2>&1 | tee -a $L
timeout 200 ncu --set full --clock-control none --import-source on -k regex:bicubic_add_clamp_pair -s 6 -c 1 -f -o gpurun_out/prof_bicubic_pair_r05c python tools/probes/ab_probe.py bicubic_pair=1 n=2 rounds=1 cool=0.1 > gpurun_out/ncu_bicubic_r05c.log 2>&1; tail -2 gpurun_out/ncu_bicubic_r05c.log | tee -a $L
