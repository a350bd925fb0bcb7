#!/bin/bash
# usage (under gpurun --gpus N): bash tools/gpu_multi_bench.sh <tag> <N> "<bench args>" ["<bench args>" ...]
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
k=0
for a in "$@"; do
  k=$((k+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+k)) \
    bench.py --gpus $N $a > gpurun_out/${TAG}_${N}gpu_$k.json 2> gpurun_out/${TAG}_${N}gpu_$k.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_${N}gpu_$k.json').read().strip().splitlines()[-1])
    mg = d.get('multi_gpu_check') or {}
    print('$a', '| ms/step %.3f' % d['ms_per_step'], 'value %.4g' % d['value'], d['scaling'], 'e2e ms %.1f' % d['e2e']['ms_per_step'],
          '| check', mg.get('pattern_ok'), mg.get('max_rel_K'), mg.get('max_rel_R'), mg.get('exchange'))
except Exception as ex:
    print('$a failed', ex)
PY
  grep -E "Error|error" gpurun_out/${TAG}_${N}gpu_$k.err | head -3 | cut -c1-300
done
