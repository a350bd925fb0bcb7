#!/bin/bash
# round 2, first GPU call: parity of the class-uniform tile kernel, then C3 bench lines with both per-nonzero kernels
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_uniform.py -x -q -p no:cacheprovider > gpurun_out/${TAG}_pytest_uniform.log 2>&1
tail -15 gpurun_out/${TAG}_pytest_uniform.log | cut -c1-300
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_uniform.py -q -x -p no:cacheprovider -k "c3_elast3d_p2_n2 or switches" > gpurun_out/${TAG}_sanitizer.log 2>&1
grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/${TAG}_sanitizer.log | head -8 | cut -c1-250
for mode in 1 0; do
  GFGPU_DEBUG=1 GFGPU_UNIFORM=$mode timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c3_u$mode.json 2> gpurun_out/${TAG}_bench_c3_u$mode.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_bench_c3_u$mode.json').read().strip().splitlines()[-1])
    print('uniform=$mode', 'ms/step %.3f' % d['ms_per_step'], d['kernel_ms'], 'frac %.4f' % d['roofline']['frac'],
          'sym %.2fs' % d['symbolic_s'], 'setup %.2fs' % d.get('setup_s', -1), 'dev %.1f GB' % (d['device_bytes'] / 1e9), d['checks'])
except Exception as ex:
    print('bench failed', ex)
PY
  grep -E "gfgpu\]|Error|error" gpurun_out/${TAG}_bench_c3_u$mode.err | head -6 | cut -c1-400
done
