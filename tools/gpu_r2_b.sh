#!/bin/bash
# uniform kernel iteration: parity tests, C3 bench line, optional env variants, ncu capture
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_uniform.py -x -q -p no:cacheprovider > gpurun_out/${TAG}_pytest_uniform.log 2>&1
tail -6 gpurun_out/${TAG}_pytest_uniform.log | cut -c1-300
run() {  # name, env...
  local name=$1; shift
  env GFGPU_DEBUG=1 "$@" timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$name.json 2> gpurun_out/${TAG}_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms/step %.3f' % d['ms_per_step'], 'tile %.3f' % d['kernel_ms']['recompute'], 'frac %.4f' % d['roofline']['frac'],
          'sym %.2fs' % d['symbolic_s'], 'setup %.2fs' % d.get('setup_s', -1), 'dev %.1f GB' % (d['device_bytes'] / 1e9), d['checks'])
except Exception as ex:
    print('$name bench failed', ex)
PY
  grep -E "uniform tiles|rror" gpurun_out/${TAG}_bench_$name.err | head -3 | cut -c1-400
}
run default
for v in "$@"; do
  [ "$v" = "$TAG" ] && continue
  run "$(echo $v | tr ' =' '__')" $v
done
env timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_utiles -s 2 -c 1 -f -o gpurun_out/${TAG}_ncu \
  python bench.py --workload c3 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_ncu.log | cut -c1-200
