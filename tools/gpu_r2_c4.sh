#!/bin/bash
# usage (under gpurun): bash tools/gpu_r2_c4.sh <tag>   -- parity of the hyperelastic paths + c4 bench A/B (v1 serial kernel vs warp-specialised)
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sumfact.py tests/test_gpu_golden.py -x -q -k "sumfact or nh or svk or hyper or c4 or finite" 2>&1 | tail -8
for v in 0 1; do
  if [ $v = 1 ]; then export GFGPU_SF_HYPER_V1=1; else unset GFGPU_SF_HYPER_V1; fi
  timeout 600 python bench.py --workload c4 --steps 5 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_c4_v$v.json 2> gpurun_out/${TAG}_c4_v$v.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_c4_v$v.json').read().strip().splitlines()[-1])
    print('v1=$v ms/step %.3f' % d['ms_per_step'], d['kernel_ms'], 'checks', d['checks'])
except Exception as ex:
    print('v1=$v failed', ex); print(open('gpurun_out/${TAG}_c4_v$v.err').read()[-1500:])
PY
done
