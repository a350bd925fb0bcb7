#!/bin/bash
# usage (under gpurun): bash tools/gpu_ab.sh <tag> <workload> lib1 lib2 ...   -- bench the same workload with several library variants
TAG=$1; WL=$2; shift; shift
mkdir -p gpurun_out
for lib in "$@"; do
  name=$(basename $lib .so)
  GFGPU_LIB=$PWD/$lib timeout 300 python bench.py --workload $WL --steps 5 --no-cpu-baseline > gpurun_out/${TAG}_${name}_$WL.json 2> gpurun_out/${TAG}_${name}_$WL.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_${name}_$WL.json').read().strip().splitlines()[-1])
    print('$name', d['config']['workload'][:3], 'ms/step %.3f' % d['ms_per_step'], {k: round(v, 3) for k, v in d['kernel_ms'].items() if v}, d['checks'])
except Exception as ex:
    print('$name bench failed', ex)
PY
  tail -2 gpurun_out/${TAG}_${name}_$WL.err
done
