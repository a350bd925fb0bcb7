"""GPU: drop-in parity with the UNMODIFIED reference IN ONE PROCESS.  oracle/_ref/shim_test links the
reference (libgetfem.so built from /root/reference/src), the C++ shim (getfem_b200/shim, the code a GetFEM
maintainer would add) and libgfgpu.so; it assembles with ga_workspace::assembly() on the CPU and with
getfem_b200::device_assembler::assembly() on the GPU and compares the gmm containers:
CSC pattern identical, values / residual within 1e-12."""
import json
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "oracle", "_ref", "shim_test")

CASES = [
    "dim=3 n=6 gt=pk k=2 q=3 im=4 family=elast",        # BASELINE config 3 (reduced)
    "dim=3 n=10 gt=pk k=1 q=1 im=2 family=laplace",     # config 2 (reduced)
    "dim=2 n=48 gt=pk k=1 q=1 im=2 family=laplace",     # config 1 (reduced)
    "dim=3 n=3 gt=qk k=2 q=3 im=6 family=nh_ciarlet",   # config 4 (reduced)
    "dim=3 n=3 gt=qk k=2 q=3 im=6 family=svk",
    "dim=3 n=4 gt=pk k=2 q=3 im=4 family=nh_bonet",
    "dim=3 n=3 gt=qk k=2 q=3 im=6 family=mooney_rivlin",   # the other laws of add_finite_strain_elasticity_brick
    "dim=3 n=3 gt=pk k=2 q=3 im=4 family=ciarlet_geymonat",
    "dim=3 n=3 gt=pk k=2 q=3 im=4 family=blatz_ko",
    "dim=3 n=5 gt=pk k=2 q=3 im=4 family=mass",
    "dim=3 n=2 gt=qk k=4 q=1 im=8 family=laplace",      # config 5 (reduced); tables come from the reference itself
    "dim=2 n=48 gt=pk k=1 q=1 im=2 family=source",      # the RHS of config 1: "-f*Test_u" (order 1 only, empty tangent)
    "dim=3 n=4 gt=pk k=2 q=3 im=4 family=source",       # "-(f.Test_u)", vector
    # mesh regions read from td.rg: boundary faces and sub-regions of convexes
    "dim=3 n=4 gt=pk k=2 q=3 im=4 family=mass region=outer",     # Robin / Dirichlet penalisation matrix
    "dim=3 n=4 gt=pk k=2 q=3 im=4 family=source region=xmax",    # Neumann load
    "dim=3 n=4 gt=pk k=2 q=3 im=4 family=nsource region=outer",  # "-(Reshape(g,qdim(u),meshdim)*Normal).Test_u"
    "dim=2 n=24 gt=pk k=1 q=1 im=2 family=nsource region=outer", # "((g).Normal)*Test_u"
    "dim=3 n=3 gt=qk k=2 q=1 im=6 family=mass region=outer",
    "dim=3 n=3 gt=qk k=2 q=3 im=6 family=nsource region=xmax",
    "dim=3 n=6 gt=pk k=2 q=3 im=4 family=elast region=half",
    "dim=3 n=3 gt=qk k=2 q=3 im=6 family=nh_ciarlet region=half",
    # model-like workspaces: stiffness + Robin boundary mass + normal source + volumic source, ONE tangent accumulated
    # on the device (gfgpu_matrix_*) and downloaded once
    "dim=3 n=4 gt=pk k=2 q=3 im=4 family=elast model=1",
    "dim=2 n=24 gt=pk k=1 q=1 im=2 family=laplace model=1",
    "dim=3 n=3 gt=qk k=2 q=3 im=6 family=nh_ciarlet model=1",
    # fem-data coefficients read from the workspace (add_fem_constant): heterogeneous material, distributed load
    "dim=3 n=4 gt=pk k=2 q=3 im=4 family=elast coef=fem kd=1",
    "dim=3 n=4 gt=pk k=2 q=1 im=4 family=laplace coef=fem kd=2",
    "dim=3 n=4 gt=pk k=2 q=3 im=4 family=source coef=fem kd=2",
    "dim=3 n=4 gt=pk k=2 q=3 im=4 family=source region=xmax coef=fem kd=1",
    "dim=3 n=3 gt=qk k=2 q=1 im=6 family=mass region=outer coef=fem kd=1",
    "dim=3 n=4 gt=pk k=2 q=3 im=4 family=elast model=1 coef=fem kd=1",
    # non-uniform mesh_fem / mesh_im (first slice): P1 + P2 (Q1 + Q2) convexes with two integration
    # methods in one mesh; one device term per group of like convexes, merged in the workspace tangent
    "dim=3 n=4 gt=pk k=2 q=3 im=4 family=elast mixed=1",
    "dim=2 n=8 gt=pk k=2 q=1 im=4 family=laplace mixed=1",
    "dim=3 n=3 gt=qk k=2 q=3 im=6 family=nh_ciarlet mixed=1",
    "dim=3 n=4 gt=pk k=2 q=3 im=4 family=mass region=outer mixed=1",
    "dim=3 n=4 gt=pk k=2 q=3 im=4 family=elast region=half mixed=1",
    # im_data coefficients (ga_workspace::add_im_data): one value per Gauss point, read on the device as a field on a synthetic
    # one-dof-per-Gauss-point fem
    "dim=3 n=4 gt=pk k=2 q=3 im=4 family=elast coef=imd",
    "dim=3 n=4 gt=pk k=2 q=1 im=4 family=laplace coef=imd",
    "dim=3 n=3 gt=qk k=2 q=1 im=6 family=mass coef=imd",
    "dim=3 n=4 gt=pk k=2 q=3 im=4 family=source coef=imd",
    "dim=2 n=12 gt=pk k=1 q=1 im=2 family=source coef=imd",
    "dim=3 n=4 gt=pk k=2 q=3 im=4 family=elast region=half coef=imd",
    # P4 simplices (the degree of the reference's published table, contrib/opt_assembly/opt_assembly.cc:704-712)
    "dim=3 n=2 gt=pk k=4 q=3 im=8 family=elast",
    "dim=3 n=2 gt=pk k=4 q=1 im=8 family=laplace",
    "dim=2 n=6 gt=pk k=4 q=2 im=8 family=elast",
    "dim=3 n=2 gt=pk k=4 q=3 im=8 family=svk",
    "dim=3 n=2 gt=pk k=4 q=3 im=8 family=mass region=outer",
]


@pytest.mark.parametrize("case", CASES)
def test_shim_matches_reference_in_process(case):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/shim_test not built (needs the reference sources)")
    out = subprocess.run([BIN] + case.split(), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["pattern_ok"] and r["nnz_ref"] == r["nnz_gpu"]
    assert 0 <= r["rel_K"] < 1e-12, r
    assert r["rel_R"] < 1e-12, r
    # the workspace tangent stays on the device between calls: the second assembly(2) downloads values only, bit-identical
    assert r["second_call_same"], r
    assert r["pattern_downloads"] == (0 if r["nnz_ref"] == 0 else 1), r
