"""GPU: the law operators of the NVRTC route that were added after the round's GPU budget was spent -- the compressible
neo-Hookean laws (nh_pk2 / nh_dpk2), the laws given through the invariants of C (iso_pk2 / iso_dpk2: compressible Mooney-Rivlin,
Ciarlet-Geymonat, generalized Blatz-Ko) and the PLANE STRAIN wrappers add_finite_strain_elasticity_brick picks for them in 2D
(adapt_law_name, getfem_nonlinear_elasticity.cc:2271-2298; plane_strain_hyperelastic_law, :906-945).  Their device helper TEXT is
compiled for the host and checked against the pinned material point in tests/test_jit_law_formulas.py (3D and the plane-strain
embedding, 1e-12), and the translated forms compile under NVRTC (tests/test_shim_probe.py); this file is the end-to-end
comparison with the reference in one process, like tests/test_gpu_dropin.py.  (It sorts last among the GPU files on purpose.)"""
import json
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "oracle", "_ref", "model_test")

MODELS = [  # 2D finite strain through the unmodified brick: the plane-strain wrapper of every law
    "model=finite_strain dim=2 n=8 gt=pk k=2 law=Compressible_Neo_Hookean_Ciarlet",
    "model=finite_strain dim=2 n=6 gt=qk k=2 law=Compressible_Neo_Hookean_Bonet",
    "model=finite_strain dim=2 n=8 gt=pk k=2 law=Compressible_Mooney_Rivlin",
    "model=finite_strain dim=2 n=6 gt=qk k=2 law=Ciarlet_Geymonat",
    "model=finite_strain dim=2 n=8 gt=pk k=2 law=Generalized_Blatz_Ko",
]


@pytest.mark.parametrize("case", MODELS)
def test_plane_strain_bricks_run_on_the_device(case):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/model_test not built (needs the reference sources)")
    out = subprocess.run([BIN] + case.split(), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["device_workspace_calls"] >= 4, r
    assert r["max_one_sided_rel"] < 1e-14, r
    assert 0 <= r["rel_K"] < 1e-12, r
    assert r["rel_rhs"] < 1e-12, r


COMPOUND = [  # the 3D operators inside a compound form (alone they are closed-form families): against the reference's AHL wrappers
    ("dim=3 n=2 gt=pk k=2 uscale=0.02", "((Id(3)+Grad_u)*Compressible_Neo_Hookean_Ciarlet_PK2(Grad_u,params)):Grad_Test_u + a*u.Test_u"),
    ("dim=3 n=2 gt=qk k=2 uscale=0.02", "((Id(3)+Grad_u)*Compressible_Neo_Hookean_Bonet_PK2(Grad_u,[1.3;0.7])):Grad_Test_u + a*u.Test_u"),
    ("dim=3 n=2 gt=pk k=2 uscale=0.02", "((Id(3)+Grad_u)*Compressible_Mooney_Rivlin_PK2(Grad_u,[0.8;0.3;2.0])):Grad_Test_u + a*u.Test_u"),
    ("dim=3 n=2 gt=qk k=2 uscale=0.02", "((Id(3)+Grad_u)*Ciarlet_Geymonat_PK2(Grad_u,[1.3;0.7;0.25])):Grad_Test_u + a*u.Test_u"),
    ("dim=3 n=2 gt=pk k=2 uscale=0.02", "((Id(3)+Grad_u)*Generalized_Blatz_Ko_PK2(Grad_u,[1.0;1.0;1.5;-0.5;1.5])):Grad_Test_u + a*u.Test_u"),
]


@pytest.mark.parametrize("mesh,expr", COMPOUND)
def test_law_operators_inside_compound_forms(mesh, expr):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/model_test not built (needs the reference sources)")
    out = subprocess.run([BIN, "model=expr"] + mesh.split() + ["expr=" + expr], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["device_workspace_calls"] >= 2, r
    assert r["pattern_ok"] and r["nnz_ref"] == r["nnz_gpu"], r
    assert 0 <= r["rel_K"] < 1e-12 and r["rel_V"] < 1e-12 and r["norm_V"] > 0, r


NORM_POTENTIALS = [  # p-Laplacian-like energies: Norm / Norm_sqr with their first and second derivatives, the second derivatives of pow
    ("dim=3 n=3 gt=pk k=2 q=1", "pow(Norm_sqr(Grad_u),1.5)/3 + pow(Norm(Grad_u),2.5)/2.5"),
    ("dim=3 n=2 gt=qk k=2", "pow(Norm(Grad_u),2.5)/2.5 + Norm(u)"),
    ("dim=2 n=6 gt=pk k=2", "pow(1+Norm_sqr(Grad_u),0.75) + sqr(Norm(u))/2"),
    # the compressible Mooney-Rivlin energy WRITTEN OUT with Matrix_j1 / Matrix_j2 / Det of the Cauchy-Green tensor (the same value
    # as Compressible_Mooney_Rivlin_potential(Grad_u,[0.8;0.3;2.0])): chain rule through two operators, second derivatives of both
    ("dim=3 n=2 gt=pk k=2 uscale=0.02",
     "0.8*(Matrix_j1(Right_Cauchy_Green(Id(3)+Grad_u))-3) + 0.3*(Matrix_j2(Right_Cauchy_Green(Id(3)+Grad_u))-3)"
     " + 2.0*sqr(sqrt(Det(Right_Cauchy_Green(Id(3)+Grad_u)))-1)"),
]


@pytest.mark.parametrize("mesh,expr", NORM_POTENTIALS)
def test_norm_potentials_run_on_the_device(mesh, expr):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/model_test not built (needs the reference sources)")
    out = subprocess.run([BIN, "model=expr"] + mesh.split() + ["expr=" + expr], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["device_workspace_calls"] >= 3, r
    assert r["pattern_ok"] and r["rel_K"] < 1e-12 and r["rel_V"] < 1e-12, r
    assert r["E_ref"] != 0 and abs(r["E_gpu"] - r["E_ref"]) <= 1e-12 * abs(r["E_ref"]), r


CONSTANTS = [  # fixed-size vector / matrix constants as parameters of a run-time compiled term (anisotropic diffusion, advection)
    ("dim=3 n=3 gt=pk k=2 q=1", "(Reshape(amat,3,3)*Grad_u).Grad_Test_u + (dvec.Grad_u)*Test_u"),
    ("dim=2 n=6 gt=qk k=2", "(Grad_u*Reshape(amat,2,2)):Grad_Test_u + sin(u.dvec)*(dvec.Test_u)"),
    ("dim=3 n=2 gt=qk k=2", "sqr(Norm(u))*Grad_u:Grad_Test_u + (dvec.u)*(dvec.Test_u)"),
    ("dim=3 n=2 gt=pk k=2", "Cross_product(u,dvec).Test_u + Norm_sqr(Cross_product(u,dvec))*(u.Test_u) + Grad_u:Grad_Test_u"),
    # slices of a matrix, M(:,j) and M(i,:) (the spellings of tests/test_assembly.cc with a state-dependent coefficient)
    ("dim=3 n=2 gt=pk k=2", "(1+u.u)*(Grad_u(:,1).Grad_Test_u(:,1) + 2*Grad_u(2,:).Grad_Test_u(2,:)) + Grad_u:Grad_Test_u"),
    # explicit vectors / matrices of expressions (GA_NODE_C_MATRIX): a state-dependent anisotropic tensor, a rotated gradient
    ("dim=2 n=6 gt=pk k=2 q=1", "([1+u*u,0.3*u;0.1*u,2+sin(u)]*Grad_u).Grad_Test_u + [Grad_u(2);-Grad_u(1)].Grad_Test_u"),
    ("dim=3 n=2 gt=qk k=2", "([u(1),0,0;0,u(2),0;0,0,1+u(3)*u(3)]*Grad_u):Grad_Test_u + [u(2);u(3);u(1)].Test_u + Grad_u:Grad_Test_u"),
    ("dim=3 n=2 gt=pk k=2 q=1", "max(u,0.2)*Grad_u.Grad_Test_u + min(u,a)*Test_u + sinc(u)*Test_u + abs(u)*Test_u + neg_part(u)*Test_u"),
]


@pytest.mark.parametrize("mesh,expr", CONSTANTS)
def test_vector_and_matrix_constants(mesh, expr):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/model_test not built (needs the reference sources)")
    out = subprocess.run([BIN, "model=expr"] + mesh.split() + ["expr=" + expr], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["device_workspace_calls"] >= 2, r
    assert r["pattern_ok"] and r["nnz_ref"] == r["nnz_gpu"], r
    assert 0 <= r["rel_K"] < 1e-12 and r["rel_V"] < 1e-12 and r["norm_V"] > 0, r
