#!/bin/bash
# usage (under gpurun): bash tools/gpu_r2_c5b.sh <tag>  -- sum-factorised parity + c5 with and without the direct mode
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sumfact.py tests/test_gpu_golden.py tests/test_gpu_workspace.py tests/test_gpu_halo.py -x -q 2>&1 | tail -4
for mode in direct staged; do
  if [ $mode = staged ]; then export GFGPU_NO_DIRECT=1; else unset GFGPU_NO_DIRECT; fi
  timeout 600 python bench.py --workload c5 --steps 5 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_c5_$mode.json 2> gpurun_out/${TAG}_c5_$mode.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_c5_$mode.json').read().strip().splitlines()[-1])
    print('c5 $mode ms/step %.3f' % d['ms_per_step'], d['kernel_ms'], 'checks', d['checks'], 'dev GB %.1f' % (d['device_bytes'] / 1e9))
except Exception as ex:
    print('c5 $mode failed', ex); print(open('gpurun_out/${TAG}_c5_$mode.err').read()[-1500:])
PY
done
