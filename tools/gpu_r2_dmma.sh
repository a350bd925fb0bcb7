#!/bin/bash
# usage (under gpurun): bash tools/gpu_r2_dmma.sh <tag>  -- A/B of the last contraction of the Q4 Laplace kernel: FMA pipe vs fp64 tensor core
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sumfact.py -x -q 2>&1 | tail -3
for v in 0 1 2 3; do
  GFGPU_SF_VARIANT=$v GFGPU_NO_DIRECT=1 timeout 600 python bench.py --workload c5 --steps 5 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_c5_v$v.json 2> gpurun_out/${TAG}_c5_v$v.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_c5_v$v.json').read().strip().splitlines()[-1])
    print('variant $v (staged output) element kernel %.3f ms  step %.3f ms' % (d['kernel_ms']['elem'], d['ms_per_step']), 'checks', d['checks'])
except Exception as ex:
    print('variant $v failed', ex); print(open('gpurun_out/${TAG}_c5_v$v.err').read()[-1500:])
PY
done
timeout 600 python bench.py --workload c5 --steps 5 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_c5_direct.json 2> gpurun_out/${TAG}_c5_direct.err
python - <<PY
import json
d = json.loads(open('gpurun_out/${TAG}_c5_direct.json').read().strip().splitlines()[-1])
print('default (direct mode) step %.3f ms' % d['ms_per_step'], d['kernel_ms'], 'checks', d['checks'])
PY
