#!/bin/bash
# uniform kernel variants: one C3 bench line per GFGPU_UT_VARIANT (and extra env), no ncu
TAG=${1:-r2v}; shift
mkdir -p gpurun_out
for v in "$@"; do
  name=$(echo $v | tr ' =,' '___')
  env GFGPU_DEBUG=1 $v timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_bench_$name.json 2> gpurun_out/${TAG}_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms/step %.3f' % d['ms_per_step'], 'tile %.3f' % d['kernel_ms']['recompute'], 'frac %.4f' % d['roofline']['frac'], d['checks'])
except Exception as ex:
    print('$name bench failed', ex)
PY
  grep -E "rror" gpurun_out/${TAG}_bench_$name.err | head -2 | cut -c1-300
done
