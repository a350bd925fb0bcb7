mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sumfact.py tests/test_gpu_dropin.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/s6a_pytest.log 2>&1; tail -3 gpurun_out/s6a_pytest.log
timeout 200 python bench.py --workload c5 --steps 5 --e2e-steps 1 --no-cpu-baseline > gpurun_out/s6a_c5.json 2> gpurun_out/s6a_c5.err
python -c "
import json
d=json.loads(open('gpurun_out/s6a_c5.json').read().strip().splitlines()[-1]); print('c5', d['ms_per_step'], d['kernel_ms'], d['checks'], d['roofline']['frac'])"
tail -2 gpurun_out/s6a_c5.err
