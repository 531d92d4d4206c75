#!/bin/bash
# scaling run on N GPUs of one box: bench.py under torchrun for N = 1, 2, 4, 8 (as many as are visible)
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  [ $N -le $NG ] || continue
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  echo "N=$N rc=$? $(python -c "import json,sys; d=json.loads([l for l in open('gpurun_out/scale_n$N.json') if l.startswith('{')][-1]); print(round(d['value'],1),'fps  e2e',round(d['e2e']['value'],1),' ms/step',round(d['ms_per_step'],4))" 2>&1)"
done
