#!/bin/bash
# usage (under gpurun): bash tools/gpu_r2_c5.sh <tag> [env assignments]  -- parity of the sum-factorised paths + c5 / c4 bench
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sumfact.py tests/test_gpu_golden.py tests/test_gpu_workspace.py tests/test_gpu_halo.py -x -q 2>&1 | tail -5
for wl in c5 c4; do
  env "$@" timeout 600 python bench.py --workload $wl --steps 5 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_$wl.json 2> gpurun_out/${TAG}_$wl.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_$wl.json').read().strip().splitlines()[-1])
    print('$wl ms/step %.3f' % d['ms_per_step'], d['kernel_ms'], 'checks', d['checks'], 'dev GB %.1f' % (d['device_bytes'] / 1e9))
except Exception as ex:
    print('$wl failed', ex); print(open('gpurun_out/${TAG}_$wl.err').read()[-1500:])
PY
done
