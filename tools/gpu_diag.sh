#!/bin/bash
# usage (under gpurun): bash tools/gpu_diag.sh  -- prints the JSON line of a few model_test cases
mkdir -p gpurun_out
for c in "model=elasticity dim=3 n=3 gt=pk k=2 dirichlet=mult" "model=elasticity dim=3 n=2 gt=pk k=1 dirichlet=mult" "model=elasticity dim=2 n=8 gt=qk k=2 dirichlet=mult" "model=finite_strain dim=3 n=2 gt=pk k=2 dirichlet=mult" "model=elasticity dim=3 n=3 gt=pk k=2 dirichlet=penal" "model=poisson dim=3 n=3 gt=pk k=2 dirichlet=mult"; do
  echo "== $c"
  timeout 120 oracle/_ref/model_test $c 2>gpurun_out/diag.err | tail -1
  grep -v "Trace\|Level" gpurun_out/diag.err | tail -3
done
timeout 300 python -m pytest tests/test_gpu_dropin.py -m gpu -q -k "reduced_mesh_fems" 2>&1 | tail -15
