#!/bin/bash
# usage (under gpurun): bash tools/gpu_round.sh <tag>  -- parity tests, default bench line, ncu launch list, ncu --set full of the two hot kernels
TAG=${1:-r}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; tail -c 3000 gpurun_out/${TAG}_bench_c3.json; tail -2 gpurun_out/${TAG}_bench_c3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_c3.csv \
  python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
tail -1 gpurun_out/${TAG}_launches.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tiles|k_affine_residual|k_gather_residual' -s 6 -c 3 -o gpurun_out/${TAG}_ncu_c3 \
  python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-300
ls -la gpurun_out
