#!/bin/bash
# usage (under gpurun): bash tools/gpu_next_session.sh <tag>
# First GPU call of the next session: everything that was written after the GPU minutes of round 1 were spent.
#  1. the whole GPU suite with the xfail / xpass report (tests/test_gpu_dropin.py::test_equivalent_spellings_run_on_the_device:
#     forms recognised BY PROBE, verified on CPU only so far -- drop the xfail marker once they pass);
#  2. the drop-in with bricks on EMPTY regions (shim: "an empty region assembles nothing");
#  3. bench lines of the five BASELINE configurations.
TAG=${1:-n}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rxX > gpurun_out/${TAG}_pytest.log 2>&1; tail -12 gpurun_out/${TAG}_pytest.log
for c in "model=elasticity dim=3 n=3 gt=pk k=2 empty_region=1" "model=poisson dim=2 n=12 gt=pk k=1 empty_region=1"; do
  timeout 120 oracle/_ref/model_test $c 2>/dev/null | tail -1 | cut -c1-400
done
bash tools/gpu_all_workloads.sh
