#!/bin/bash
# 8-GPU round: copy-only ceiling of the end-to-end path at N = 1/2/4/8, bench.py at N = 8 (64 frames sharded, bitwise gather check),
# cfg4 x2 / x6, cfg5a and cfg5b frame-sharded on 8 GPUs.  usage: tools/gpu_scale8.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
: > gpurun_out/copy_ceiling_$TAG.jsonl
for N in 1 2 4 8; do
  [ $N -le $NG ] || continue
  timeout 300 $TR --nproc-per-node $N --master-port $((29700+N)) tools/probes/copy_ceiling.py 2>gpurun_out/copy_$N.err | grep '^{' >> gpurun_out/copy_ceiling_$TAG.jsonl
done
cat gpurun_out/copy_ceiling_$TAG.jsonl | cut -c1-600
for N in 8 4 2; do
  [ $N -le $NG ] || continue
  timeout 600 $TR --nproc-per-node $N --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
  echo "N=$N rc=$? $(python -c "import json,sys; d=json.loads([l for l in open('gpurun_out/scale_${TAG}_n$N.json') if l.startswith('{')][-1]); print(round(d['value'],1),'fps  e2e',round(d['e2e']['value'],1),' ms/step',round(d['ms_per_step'],4), d['config'].get('sharding_check'), d['config'].get('sustained',{}).get('frames_per_s'))" 2>&1)"
done
[ 8 -le $NG ] && timeout 900 $TR --nproc-per-node 8 --master-port 29650 tools/bench_configs.py $TAG --only cfg4_fast_720p_x2,cfg4_fast_720p_x6,cfg5a,cfg5b,cfg2_window_720p_1080p_b8 2>gpurun_out/configs8.err | cut -c1-400
