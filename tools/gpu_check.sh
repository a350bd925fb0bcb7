#!/bin/bash
# usage (under gpurun): bash tools/gpu_check.sh <tag> [workloads...]   -- parity tests, bench lines, one ncu --set full capture
TAG=${1:-x}; shift
WL=${@:-c3 c2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
for w in $WL; do
  timeout 300 python bench.py --workload $w --steps 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_bench_$w.json').read().strip().splitlines()[-1])
    print(d['config']['workload'][:3], 'ms/step %.3f' % d['ms_per_step'], d['kernel_ms'], 'frac %.4f' % d['roofline']['frac'],
          'sym %.2fs' % d['symbolic_s'], 'dev %.1f GB' % (d['device_bytes'] / 1e9), d['checks'])
except Exception as ex:
    print('bench $w failed', ex)
PY
  tail -2 gpurun_out/${TAG}_bench_$w.err
done
if [ -n "$NCU_KERNEL" ]; then
  ncu --set full --clock-control none --import-source on -k regex:$NCU_KERNEL -s ${NCU_SKIP:-2} -c 1 -o gpurun_out/${TAG}_ncu \
    python bench.py --workload ${NCU_WL:-c3} --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-200
fi
