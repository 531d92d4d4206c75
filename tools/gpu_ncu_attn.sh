#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probes/attn_shape_probe.py > gpurun_out/plain_attn.log 2>&1 || { tail -3 gpurun_out/plain_attn.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"global_attn_mma_kernel" -s 60 -c 1 -o gpurun_out/prof_attn_bal -f python tools/probes/attn_shape_probe.py > gpurun_out/ncu_attn.log 2>&1
echo rc=$?; ls -la gpurun_out/prof_attn_bal.ncu-rep
