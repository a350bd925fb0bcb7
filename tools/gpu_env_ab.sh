#!/bin/bash
# usage (under gpurun): bash tools/gpu_env_ab.sh <tag> <workload> "ENV=a ENV2=b" "ENV=c" ...  -- bench one workload under several environments
TAG=$1; WL=$2; shift; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 300 python bench.py --workload $WL --steps 5 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_$i.json 2> gpurun_out/${TAG}_$i.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_$i.json').read().strip().splitlines()[-1])
    print('[$envs]', 'ms/step %.3f' % d['ms_per_step'], {k: round(v, 3) for k, v in d['kernel_ms'].items() if v}, d['checks'])
except Exception as ex:
    print('[$envs] bench failed', ex)
PY
  tail -2 gpurun_out/${TAG}_$i.err
done
