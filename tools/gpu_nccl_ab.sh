#!/bin/bash
# usage (under gpurun --gpus 2): bash tools/gpu_nccl_ab.sh <tag>  -- the C3 weak 2-GPU step under several NCCL point-to-point settings
TAG=$1
mkdir -p gpurun_out
k=0
for envs in "X=1" "NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32" "NCCL_P2P_USE_CUDA_MEMCPY=1"; do
  k=$((k+1))
  env $envs NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,P2P timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port $((29700+k)) bench.py --gpus 2 --workload c3 --steps 8 --e2e-steps 1 --no-cpu-baseline \
    > gpurun_out/${TAG}_$k.json 2> gpurun_out/${TAG}_$k.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_$k.json').read().strip().splitlines()[-1])
    mg = d.get('multi_gpu_check') or {}
    print('[$envs]', 'ms/step %.3f' % d['ms_per_step'], {k: round(v, 3) for k, v in d['kernel_ms'].items() if v}, '| check', mg.get('pattern_ok'), mg.get('max_rel_K'))
except Exception as ex:
    print('[$envs] failed', ex)
PY
  grep -iE "p2p.*channel|nChannelsPerPeer|P2P Chunksize|via P2P" gpurun_out/${TAG}_$k.err | sort | uniq -c | head -6 | cut -c1-220
done
