for w in c1 c2 c3 c4 c5; do
  GFGPU_DEBUG=1 timeout 600 python bench.py --workload $w --steps 5 --e2e-steps 1 --no-cpu-baseline > gpurun_out/s5f_$w.json 2> gpurun_out/s5f_$w.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/s5f_$w.json').read().strip().splitlines()[-1])
    print('$w', 'ne', d['config']['elements'], 'nnz', d['config']['nnz'], 'ms/step %.3f' % d['ms_per_step'], 'Melt/s %.1f' % (d['value']/1e6), {k: round(v, 3) for k, v in d['kernel_ms'].items() if v}, 'frac %.3f fp64 %.3f of %.1f TF' % (d['roofline']['frac'], d['roofline_fp64']['frac'], d['roofline_fp64']['peak']), 'sym %.2f' % d['symbolic_s'], 'GB %.1f' % (d['device_bytes']/1e9))
except Exception as ex:
    print('$w bench failed', ex)
PY
  grep gfgpu gpurun_out/s5f_$w.err | tail -2; tail -1 gpurun_out/s5f_$w.err | cut -c1-300
done
