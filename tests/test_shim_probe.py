"""CPU: recognition BY PROBE in the C++ shim (getfem_b200/shim/gfgpu_getfem_shim.cc::recognise_by_probe).  The reference's own
tests write one bilinear form in many algebraically equivalent ways (tests/test_assembly.cc:777-866, lambda = 3, mu = 2); the
shim identifies such a directly written order-2 tree numerically, from the reference's interpreter on two convexes, and sends
the FAMILY to the device.  Here, without a GPU: the dispatch patch in GFGPU_DRYRUN mode prints what it would send (family and
fitted parameters) for each spelling, and says "NOT recognised" for forms outside the families."""
import os
import re
import subprocess

import pytest

from conftest import ROOT

BIN = os.path.join(ROOT, "oracle", "_ref", "model_test")
LAPLACE, ELASTICITY, MASS = 0, 1, 5
EL = "mu*(Grad_Test_u'+Grad_Test_u):Grad_Test2_u"

CASES = [  # (mesh args, expression, family, parameters)
    ("dim=3 n=2 gt=pk k=2", "(lambda*Trace(Grad_Test_u)*Id(qdim(u)) + mu*(Grad_Test_u'+Grad_Test_u)):Grad_Test2_u", ELASTICITY, (3, 2)),
    ("dim=3 n=2 gt=pk k=2", "lambda*Div_Test_u*Div_Test2_u + " + EL, ELASTICITY, (3, 2)),
    ("dim=3 n=2 gt=pk k=2", "lambda*((Grad_Test2_u@Grad_Test_u):Id(meshdim)):Id(meshdim) + " + EL, ELASTICITY, (3, 2)),
    ("dim=3 n=2 gt=pk k=2", "lambda*Id(meshdim)@Id(meshdim)*Grad_Test_u:Grad_Test2_u + " + EL, ELASTICITY, (3, 2)),
    ("dim=3 n=2 gt=pk k=2", "lambda*(Id(meshdim)*Id(meshdim))@Id(meshdim)*Grad_Test_u:Grad_Test2_u + " + EL, ELASTICITY, (3, 2)),
    ("dim=3 n=2 gt=pk k=2", "lambda*Trace(Grad_Test_u)*Trace(Grad_Test2_u) +mu*(Grad_Test_u'(:,1)+Grad_Test_u(:,1)):Grad_Test2_u(:,1)"
     "+mu*(Grad_Test_u'(:,2)+Grad_Test_u(:,2)):Grad_Test2_u(:,2)+mu*(Grad_Test_u'(:,3)+Grad_Test_u(:,3)):Grad_Test2_u(:,3)", ELASTICITY, (3, 2)),
    ("dim=3 n=2 gt=pk k=2", "lambda*Trace(Grad_Test_u)*Trace(Grad_Test2_u) + mu*(Grad_Test_u'(1,:)+Grad_Test_u(1,:)):Grad_Test2_u(1,:)"
     "+ mu*(Grad_Test_u'(2,:)+Grad_Test_u(2,:)):Grad_Test2_u(2,:)+mu*(Grad_Test_u'(3,:)+Grad_Test_u(3,:)):Grad_Test2_u(3,:)", ELASTICITY, (3, 2)),
    ("dim=2 n=4 gt=pk k=2", "lambda*Trace(Grad_Test_u)*Trace(Grad_Test2_u) +mu*(Grad_Test_u'(:,1)+Grad_Test_u(:,1)):Grad_Test2_u(:,1)"
     "+mu*(Grad_Test_u'(:,2)+Grad_Test_u(:,2)):Grad_Test2_u(:,2)", ELASTICITY, (3, 2)),
    ("dim=3 n=2 gt=qk k=2", "lambda*Div_Test_u*Div_Test2_u + " + EL, ELASTICITY, (3, 2)),
    ("dim=3 n=2 gt=pk k=2", "2*mu*Sym(Grad_Test_u):Sym(Grad_Test2_u)", ELASTICITY, (0, 2)),
    # scalar forms (q=1)
    ("dim=3 n=2 gt=pk k=2 q=1", "Grad_Test_u(1)*Grad_Test2_u(1) + Grad_Test_u(2)*Grad_Test2_u(2) + Grad_Test_u(3)*Grad_Test2_u(3)", LAPLACE, (1,)),
    ("dim=3 n=2 gt=pk k=2 q=1", "[Grad_Test_u(1); Grad_Test_u(3); Grad_Test_u(2)].[Grad_Test2_u(1); Grad_Test2_u(3); Grad_Test2_u(2)]", LAPLACE, (1,)),
    ("dim=2 n=4 gt=pk k=1 q=1", "a*[Grad_Test_u(2); Grad_Test_u(1)].[Grad_Test2_u(2); Grad_Test2_u(1)]", LAPLACE, (1.7,)),
    ("dim=3 n=2 gt=pk k=2 q=1", "Test2_u*(a*Test_u)*2", MASS, (3.4,)),
    # vector Laplace and vector mass
    ("dim=3 n=2 gt=pk k=1", "a*(Grad_Test_u(1,:).Grad_Test2_u(1,:) + Grad_Test_u(2,:).Grad_Test2_u(2,:) + Grad_Test_u(3,:).Grad_Test2_u(3,:))", LAPLACE, (1.7,)),
    ("dim=2 n=4 gt=pk k=2", "Test_u(1)*Test2_u(1) + Test_u(2)*Test2_u(2)", MASS, (1,)),
]

NOT_FAMILIES = [
    ("dim=3 n=2 gt=pk k=2", "Grad_Test_u(1,:).Grad_Test2_u(1,:)"),                     # one row of the gradient only
    ("dim=3 n=2 gt=pk k=2", "lambda*Div_Test_u*Div_Test2_u"),                           # no shear part: see the test
    ("dim=3 n=2 gt=pk k=2 q=1", "X(1)*Grad_Test_u.Grad_Test2_u"),                       # a coefficient that varies in space
    ("dim=3 n=2 gt=pk k=2 q=1", "Grad_Test_u(1)*Grad_Test2_u(1) + 2*Grad_Test_u(2)*Grad_Test2_u(2) + Grad_Test_u(3)*Grad_Test2_u(3)"),  # anisotropic
]


def _dryrun_order1(mesh, expr):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/model_test not built (needs the reference sources)")
    out = subprocess.run([BIN, "model=expr"] + mesh.split() + ["expr=" + expr], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, GFGPU_DRYRUN="1"))
    assert out.returncode == 0, out.stderr[-1500:]
    lines = [l for l in out.stderr.splitlines() if l.startswith("[gfgpu dryrun] order 1 (assembly order 2)")]
    assert len(lines) == 1, out.stderr[-1500:]
    return lines[0]


def _dryrun(mesh, expr):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/model_test not built (needs the reference sources)")
    out = subprocess.run([BIN, "model=expr"] + mesh.split() + ["expr=" + expr], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, GFGPU_DRYRUN="1"))
    assert out.returncode == 0, out.stderr[-1500:]
    lines = [l for l in out.stderr.splitlines() if l.startswith("[gfgpu dryrun] order 2 (assembly order 2)")]
    assert len(lines) == 1, out.stderr[-1500:]
    return lines[0]


@pytest.mark.parametrize("mesh,expr,family,params", CASES)
def test_equivalent_spellings_are_recognised(mesh, expr, family, params):
    line = _dryrun(mesh, expr)
    m = re.search(r"-> recognised family (\d+) \(([^)]*)\)", line)
    assert m, line
    assert int(m.group(1)) == family, line
    got = [float(x) for x in m.group(2).split()]
    assert len(got) == len(params) and all(abs(g - p) <= 1e-9 * max(1.0, abs(p)) for g, p in zip(got, params)), line


@pytest.mark.parametrize("mesh,expr", NOT_FAMILIES)
def test_other_forms_are_not_recognised(mesh, expr):
    line = _dryrun(mesh, expr)
    if "lambda*Div_Test_u*Div_Test2_u" == expr:  # a pure volumetric form IS lambda D + 0 S: elasticity with mu snapped to 0
        assert "-> recognised family 1 ( 3 0 )" in line, line
        return
    assert "NOT recognised" in line, line


EU = "mu*(Grad_u'+Grad_u):Grad_Test_u"
DERIVED_CASES = [  # written in u: an order-1 tree (identified through its derivative, and checked to be K u) and its tangent
    ("dim=3 n=2 gt=pk k=2", "lambda*Div_u*Div_Test_u + " + EU, ELASTICITY, (3, 2)),
    ("dim=3 n=2 gt=pk k=2", "lambda*Trace(Grad_u)*Trace(Grad_Test_u) + mu*(Grad_u'(:,1)+Grad_u(:,1)):Grad_Test_u(:,1)"
     "+mu*(Grad_u'(:,2)+Grad_u(:,2)):Grad_Test_u(:,2)+mu*(Grad_u'(:,3)+Grad_u(:,3)):Grad_Test_u(:,3)", ELASTICITY, (3, 2)),
    ("dim=2 n=4 gt=qk k=2", "(lambda*Trace(Grad_u)*Id(qdim(u)) + mu*(Grad_u'+Grad_u)):Grad_Test_u", ELASTICITY, (3, 2)),
    # the potentials of tests/test_assembly.cc:777-803: order 0, differentiated twice by add_expression
    ("dim=3 n=2 gt=pk k=2 q=1", "(Grad_u:Grad_u)/2", LAPLACE, (1,)),
    ("dim=3 n=2 gt=pk k=2 q=1", "sqr(Norm(Grad_u))/2", LAPLACE, (1,)),
    ("dim=3 n=2 gt=pk k=2 q=1", "Norm_sqr(Grad_u)/2", LAPLACE, (1,)),
    ("dim=3 n=2 gt=pk k=2 q=1", "a*(sqr(Grad_u(1)) + sqr(Grad_u(2)) + sqr(Grad_u(3)))/2", LAPLACE, (1.7,)),
    ("dim=2 n=4 gt=pk k=1 q=1", "([Grad_u(2); Grad_u(1)].[Grad_u(2); Grad_u(1)])/2", LAPLACE, (1,)),
    ("dim=2 n=4 gt=pk k=2 q=1", "a*u*u/2", MASS, (1.7,)),
]


@pytest.mark.parametrize("mesh,expr,family,params", DERIVED_CASES)
def test_derived_spellings_are_recognised(mesh, expr, family, params):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/model_test not built (needs the reference sources)")
    out = subprocess.run([BIN, "model=expr"] + mesh.split() + ["expr=" + expr], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, GFGPU_DRYRUN="1"))
    assert out.returncode == 0, out.stderr[-1500:]
    seen = set()
    for line in out.stderr.splitlines():
        m = re.match(r"\[gfgpu dryrun\] order ([12]) \(assembly order ([12])\)", line)
        if not m:
            continue
        r = re.search(r"-> recognised family (\d+) \(([^)]*)\)", line)
        assert r, line
        got = [float(x) for x in r.group(2).split()]
        assert int(r.group(1)) == family and len(got) == len(params), line
        assert all(abs(g - p) <= 1e-9 * max(1.0, abs(p)) for g, p in zip(got, params)), line
        seen.add((int(m.group(1)), int(m.group(2))))
    assert seen == {(1, 1), (1, 2), (2, 1), (2, 2)}, (seen, out.stderr[-800:])


def test_a_load_mixed_into_the_tree_is_not_taken_for_the_linear_form():
    """r = K u must hold for the order-1 tree: with a source term summed into the same tree the probe alone never accepts it
    as a linear family -- the tree is refused, or (vector variables, round 2) translated as it stands into ONE run-time
    compiled term (family 11) whose first form keeps the load and whose second form is the reference's own order-2 tree."""
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/model_test not built (needs the reference sources)")
    expr = "lambda*Trace(Grad_u)*Trace(Grad_Test_u) + mu*(Grad_u'+Grad_u):Grad_Test_u + [1;2;3].Test_u"
    out = subprocess.run([BIN, "model=expr", "dim=3", "n=2", "gt=pk", "k=2", "expr=" + expr], capture_output=True, text=True,
                         timeout=300, env=dict(os.environ, GFGPU_DRYRUN="1"))
    lines = [l for l in out.stderr.splitlines() if l.startswith("[gfgpu dryrun] order 1")]
    assert lines and all("NOT recognised" in l or "recognised family 11 " in l for l in lines), out.stderr[-1500:]


SOURCE = 6
LOAD_CASES = [  # constant loads in spellings the printed forms do not know -> family SOURCE with the fitted F
    ("dim=3 n=2 gt=pk k=2", "-(a*[1;2;3]).Test_u", (-1.7, -3.4, -5.1)),
    ("dim=3 n=2 gt=pk k=2", "Test_u(2)*lambda - mu*Test_u(1)", (-2, 3, 0)),
    ("dim=2 n=4 gt=pk k=1 q=1", "(a/2)*Test_u*2", (1.7,)),
    ("dim=3 n=2 gt=qk k=2 region=1", "[0;mu;-lambda].Test_u*2", (0, 4, -6)),   # a Neumann load on the faces x = 1
]
REFUSED = [  # the probe sees two items only: anything that may vary over the region is not its business
    ("dim=3 n=2 gt=pk k=2 q=1", "Grad_Test_u.(c0*Grad_Test2_u)*2"),               # fem-data material: constant on the probe convexes only
    ("dim=3 n=2 gt=pk k=2 q=1", "c0*Test_u*3"),                                   # fem-data load
]


@pytest.mark.parametrize("mesh,expr,F", LOAD_CASES)
def test_constant_loads_are_recognised(mesh, expr, F):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/model_test not built (needs the reference sources)")
    out = subprocess.run([BIN, "model=expr"] + mesh.split() + ["expr=" + expr], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, GFGPU_DRYRUN="1"))
    assert out.returncode == 0, out.stderr[-1500:]
    lines = [l for l in out.stderr.splitlines() if l.startswith("[gfgpu dryrun] order 1")]
    assert lines, out.stderr[-1500:]
    for line in lines:
        r = re.search(r"-> recognised family (\d+) \(([^)]*)\)", line)
        assert r and int(r.group(1)) == SOURCE, line
        got = [float(x) for x in r.group(2).split()]
        assert len(got) == len(F) and all(abs(g - f) <= 1e-9 * max(1.0, abs(f)) for g, f in zip(got, F)), line


@pytest.mark.parametrize("mesh,expr", REFUSED)
def test_the_probe_refuses_what_may_vary_over_the_region(mesh, expr):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/model_test not built (needs the reference sources)")
    out = subprocess.run([BIN, "model=expr"] + mesh.split() + ["expr=" + expr], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, GFGPU_DRYRUN="1"))
    lines = [l for l in out.stderr.splitlines() if l.startswith("[gfgpu dryrun] order")]
    # never a constant-coefficient family fitted on two items: refused, or (a fem-data LOAD, round 2) translated as it stands for
    # the NVRTC route, which evaluates the field at every Gauss point
    assert lines and all("NOT recognised" in l or "recognised family 11 " in l for l in lines), out.stderr[-1500:]


NONLINEAR = [  # forms that contain the unknown: a numerical fit at ONE state says nothing
    ("dim=2 n=4 gt=pk k=1 q=1", "(1+u*u)*Grad_u.Grad_Test_u"),
    ("dim=2 n=4 gt=pk k=1 q=1", "u*u*u*Test_u+Grad_u.Grad_Test_u"),
    ("dim=2 n=4 gt=pk k=1 q=1", "u*Test_u*Test2_u"),       # a "bilinear form" whose coefficient is the state
]


@pytest.mark.parametrize("uzero", ["4", "-1", None])
@pytest.mark.parametrize("mesh,expr", NONLINEAR)
def test_state_dependent_forms_are_refused_whatever_the_state(mesh, expr, uzero):
    """with u = 0 on the probe convexes (uzero=4: the first two triangles) or everywhere (the first Newton step) the old
    probe fitted a plain Laplacian / a constant load; the structural check refuses the tree before any fit"""
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/model_test not built (needs the reference sources)")
    args = [BIN, "model=expr"] + mesh.split() + ["expr=" + expr] + (["uzero=" + uzero] if uzero else [])
    out = subprocess.run(args, capture_output=True, text=True, timeout=300, env=dict(os.environ, GFGPU_DRYRUN="1"))
    lines = [l for l in out.stderr.splitlines() if l.startswith("[gfgpu dryrun] order")]
    # never a constant-coefficient family fitted at one state: either refused, or (scalar variable, round 2) the tree itself
    # translated for the NVRTC route -- family 11, which evaluates the state at every Gauss point
    assert lines and all("NOT recognised" in l or "recognised family 11 " in l for l in lines), out.stderr[-1500:]


def test_a_load_written_with_the_unknown_is_not_a_constant_load():
    """order-1-only tree u*Test_u (added without derivatives it would look like a load): u locally constant must not fit SOURCE"""
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/model_test not built (needs the reference sources)")
    out = subprocess.run([BIN, "model=expr", "dim=2", "n=4", "gt=pk", "k=1", "q=1", "expr=u*Test_u", "uzero=-1"],
                         capture_output=True, text=True, timeout=300, env=dict(os.environ, GFGPU_DRYRUN="1"))
    lines = [l for l in out.stderr.splitlines() if l.startswith("[gfgpu dryrun] order")]
    # u*Test_u has a derivative tree (mass), so it is the derived route: recognised as MASS with r = M u checked, never as a load
    assert lines and not any("recognised family 6" in l for l in lines), out.stderr[-1500:]


def test_coupled_trees_of_the_incompressibility_brick_are_recognised():
    """CPU (dry run of the dispatch patch): the four trees of "-p*Div_Test_u-Test_p*Div_u" as the reference prints them are
    matched as coupled div-pressure parts (shim tag 1000); nothing is left to the "not recognised" path."""
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/model_test not built (needs the reference sources)")
    out = subprocess.run([BIN, "model=incompressible", "dim=3", "n=2", "gt=pk", "k=2"], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, GFGPU_DRYRUN="1"))
    assert out.returncode == 0, out.stderr[-1500:]
    lines = [l for l in out.stderr.splitlines() if l.startswith("[gfgpu dryrun]")]
    assert not [l for l in lines if "NOT recognised" in l], lines
    coupled = [l for l in lines if "family 1000" in l]
    forms = {l.split(": ", 1)[1].split(" -> ")[0] for l in coupled}
    assert forms == {"-(Test_p*Div_u)", "-(Test_p*Div_Test2_u)", "(-p)*Div_Test_u", "(-Test2_p)*Div_Test_u"}, forms


JIT_EXPRS = [
    "(1+sqr(u))*Grad_u.Grad_Test_u + sin(u)*Test_u",
    "a*exp(u)*Grad_u.Grad_Test_u + a*Norm_sqr(Grad_u)*Test_u",
    "pow(1+Norm_sqr(Grad_u),0.75)*Grad_u.Grad_Test_u - a*Test_u",
    "sqrt(1+Norm_sqr(Grad_u))*Test_u + Grad_u(1)*Test_u",
    "([1,2,3].Grad_u)*Test_u + 0.1*Grad_u.Grad_Test_u",
]


@pytest.mark.parametrize("expr", JIT_EXPRS)
def test_general_scalar_expressions_take_the_nvrtc_route(expr):
    """CPU (dry run): an expression of a scalar variable that no normal form and no probe covers is translated from the
    reference's analysed trees into a JIT term (family 11); constants it names become par[k]."""
    line = _dryrun_order1("dim=3 n=2 gt=pk k=2 q=1", expr)
    assert "recognised family 11" in line, line


X_EXPRS = [  # the position X (whole vector or one coordinate): x = sum_g G_g N_g(q) from the geometric transformation's values
    ("dim=3 n=2 gt=pk k=2 q=1", "X(1)*sin(u)*Test_u"),
    ("dim=3 n=2 gt=pk k=2 q=1", "(1+X.X)*Grad_u.Grad_Test_u + X(3)*Test_u"),
    ("dim=3 n=2 gt=pk k=2", "X(1)*Test_u(1)"),                                    # a body load that varies in space: no order-2 tree
    ("dim=2 n=4 gt=qk k=2", "(1+X(2))*Grad_u:Grad_Test_u - X.Test_u"),
    # boundary faces: the probe sees two items only, so what varies over the region (X, Normal on a non-planar boundary, the
    # state) is compiled as it stands -- the run-time kernel takes the face tables and the unit normal (C&E.cc:8836-8847)
    ("dim=3 n=2 gt=pk k=2 region=2", "X(1)*Test_u(1)"),
    ("dim=3 n=2 gt=pk k=2 region=2", "Test_u.(lambda*Normal)*2"),
    ("dim=3 n=2 gt=pk k=2 q=1 region=2", "X(1)*sin(u)*Test_u + (u*u*u*u)*Test_u"),             # radiation-like Robin condition
    ("dim=2 n=4 gt=qk k=2 region=2", "(u.Normal)*(Test_u.Normal)*(1+X(1)) + exp(u(1))*Test_u(2)"),
    # the Saint-Venant Kirchhoff operator and its derivative (the one law the reference defines in any dimension: 2D finite strain)
    ("dim=2 n=4 gt=pk k=2", "((Id(2)+Grad_u)*Saint_Venant_Kirchhoff_PK2(Grad_u,params)):Grad_Test_u"),
    ("dim=3 n=2 gt=pk k=2", "((Id(3)+Grad_u)*Saint_Venant_Kirchhoff_PK2(Grad_u,[1.3;0.7])):Grad_Test_u + a*u.Test_u"),
    ("dim=2 n=4 gt=pk k=2", "((Id(2)+Grad_u)*Plane_Strain_Compressible_Neo_Hookean_Ciarlet_PK2(Grad_u,params)):Grad_Test_u"),
    ("dim=3 n=2 gt=pk k=2", "((Id(3)+Grad_u)*Compressible_Neo_Hookean_Bonet_PK2(Grad_u,params)):Grad_Test_u + a*u.Test_u"),
    ("dim=2 n=4 gt=qk k=2", "((Id(2)+Grad_u)*Plane_Strain_Compressible_Mooney_Rivlin_PK2(Grad_u,[0.8;0.3;2.0])):Grad_Test_u"),
    ("dim=3 n=2 gt=pk k=2", "((Id(3)+Grad_u)*Ciarlet_Geymonat_PK2(Grad_u,[1.3;0.7;0.25])):Grad_Test_u + a*u.Test_u"),
    ("dim=2 n=4 gt=pk k=2", "((Id(2)+Grad_u)*Plane_Strain_Generalized_Blatz_Ko_PK2(Grad_u,[1.0;1.0;1.5;-0.5;1.5])):Grad_Test_u"),
    # nonlinear matrix operators and their derivatives: a compressible neo-Hookean law WRITTEN OUT with Det / Inv, the
    # large-strain helpers
    ("dim=3 n=2 gt=pk k=2", "(mu*((Id(3)+Grad_u) - Inv(Id(3)+Grad_u)') + lambda*log(Det(Id(3)+Grad_u))*Inv(Id(3)+Grad_u)'):Grad_Test_u"),
    ("dim=2 n=4 gt=qk k=2", "(mu*((Id(2)+Grad_u) - Inv(Id(2)+Grad_u)') + lambda*log(Det(Id(2)+Grad_u))*Inv(Id(2)+Grad_u)'):Grad_Test_u"),
    ("dim=3 n=2 gt=pk k=2", "((Id(3)+Grad_u)*(lambda*Trace(Green_Lagrangian(Id(3)+Grad_u))*Id(3)+2*mu*Green_Lagrangian(Id(3)+Grad_u))):Grad_Test_u"),
    ("dim=3 n=2 gt=pk k=2", "(Matrix_i2(Right_Cauchy_Green(Id(3)+Grad_u))*Left_Cauchy_Green(Id(3)+Grad_u)):Grad_Test_u"),
    # a POTENTIAL: the reference differentiates it twice; orders 0, 1 and 2 are translated (second derivatives of Det / log, ...)
    ("dim=3 n=2 gt=pk k=2", "mu/2*(Trace(Right_Cauchy_Green(Id(3)+Grad_u))-3) - mu*log(Det(Id(3)+Grad_u)) + lambda/2*sqr(log(Det(Id(3)+Grad_u)))"),
    ("dim=2 n=4 gt=pk k=2 q=1", "(1+sqr(u))*Norm_sqr(Grad_u)/2 + cos(u) + a*u"),
    # p-Laplacian energies: Norm / Norm_sqr with their derivatives, the second derivatives of pow
    ("dim=3 n=2 gt=pk k=2 q=1", "pow(Norm_sqr(Grad_u),1.5)/3 + pow(Norm(Grad_u),2.5)/2.5"),
    ("dim=3 n=2 gt=pk k=2", "pow(Norm(Grad_u),2.5)/2.5 + Norm(u)"),
    ("dim=3 n=2 gt=pk k=2", "sqr(Norm(u))*Grad_u:Grad_Test_u"),
    ("dim=2 n=4 gt=pk k=2 q=1", "([1+u*u,0.3*u;0.1*u,2+sin(u)]*Grad_u).Grad_Test_u + [Grad_u(2);-Grad_u(1)].Grad_Test_u"),
    ("dim=3 n=2 gt=pk k=2", "Cross_product(u,dvec).Test_u + Norm_sqr(Cross_product(u,dvec))*(u.Test_u)"),
    ("dim=3 n=2 gt=pk k=2 q=1", "max(u,0.2)*Grad_u.Grad_Test_u + min(u,a)*Test_u + sinc(u)*Test_u + abs(u)*Test_u + neg_part(u)*Test_u"),
    ("dim=3 n=2 gt=pk k=2", "0.8*(Matrix_j1(Right_Cauchy_Green(Id(3)+Grad_u))-3) + 0.3*(Matrix_j2(Right_Cauchy_Green(Id(3)+Grad_u))-3)"
                            " + 2.0*sqr(sqrt(Det(Right_Cauchy_Green(Id(3)+Grad_u)))-1)"),
    ("dim=3 n=2 gt=qk k=2", "Matrix_i2(Green_Lagrangian(Id(3)+Grad_u)) + sqr(Trace(Green_Lagrangian(Id(3)+Grad_u))) + tanh(u.u)"),
    # fixed-size VECTOR and MATRIX constants (an anisotropic diffusion tensor through Reshape(A,N,N), an advection direction)
    ("dim=3 n=2 gt=pk k=2 q=1", "(Reshape(amat,3,3)*Grad_u).Grad_Test_u + (dvec.Grad_u)*Test_u"),
    ("dim=2 n=4 gt=qk k=2", "(Grad_u*Reshape(amat,2,2)):Grad_Test_u + sin(u.dvec)*(dvec.Test_u)"),
    # scalar fem-data coefficients inside a translated tree: fld[k], evaluated on the data fem at every Gauss point
    ("dim=3 n=2 gt=pk k=2 q=1", "c0*sin(u)*Test_u + (1+c0)*Grad_u.Grad_Test_u"),
    ("dim=3 n=2 gt=pk k=2", "c0*(1+Norm_sqr(u))*Grad_u:Grad_Test_u + c0*X.Test_u"),
    ("dim=2 n=4 gt=qk k=2 q=1 region=2", "c0*u*u*u*Test_u"),
    # ONE vector-valued fem-data field (vfld): advection by a data velocity
    ("dim=3 n=2 gt=pk k=2 q=1", "(w0.Grad_u)*Test_u + 0.1*Grad_u.Grad_Test_u"),
    ("dim=2 n=4 gt=qk k=2", "(Grad_u*w0).Test_u + (1+Norm_sqr(w0))*Grad_u:Grad_Test_u - w0.Test_u"),
]


@pytest.mark.parametrize("mesh,expr", X_EXPRS)
def test_expressions_of_the_position_take_the_nvrtc_route(mesh, expr):
    line = _dryrun_order1(mesh, expr)
    assert "recognised family 11" in line, line


@pytest.mark.parametrize("mesh,expr", X_EXPRS + [("dim=3 n=2 gt=pk k=2 q=1", e) for e in JIT_EXPRS])
def test_the_translated_forms_compile(mesh, expr):
    """CPU: what the translator emits for the reference's order-1 / order-2 trees is valid source for the run-time kernel
    (NVRTC targets sm_100a without a GPU) -- position, unit normal, parameters and probes included."""
    from getfem_b200 import capi
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/model_test not built (needs the reference sources)")
    out = subprocess.run([BIN, "model=expr"] + mesh.split() + ["expr=" + expr], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, GFGPU_DRYRUN="1"))
    assert out.returncode == 0, out.stderr[-1500:]
    forms = [l.split("\t") for l in out.stderr.splitlines() if l.startswith("[gfgpu dryrun jit]")]
    assert forms, out.stderr[-1500:]
    for head, f1, f2, f0 in forms:
        m = re.search(r"dim=(\d) qdim=(\d)", head)
        capi.jit_check(int(m.group(1)), f1, f2, qdim=int(m.group(2)))
        if f0:  # the order-0 integrand (a scalar without test functions), compiled in the place of the first form
            capi.jit_check(int(m.group(1)), "(" + f0 + ")*dot(tv,tv)", f2, qdim=int(m.group(2)))
    if "potential:" in expr or expr.startswith("mu/2*("):
        assert all(f[3] for f in forms), forms


def test_the_nvrtc_route_refuses_what_it_cannot_express():
    """operators outside the translator's language: no silent approximation -- the tree is reported as not
    recognised"""
    for mesh, expr in (("dim=3 n=2 gt=pk k=2 q=1", "Hess_u:Hess_Test_u"), ("dim=3 n=2 gt=pk k=2", "Expm(Grad_u):Grad_Test_u")):
        line = _dryrun_order1(mesh, expr)
        assert "NOT recognised" in line, line


def test_the_trees_of_the_dirichlet_brick_with_multipliers_are_recognised():
    """CPU (dry run): add_Dirichlet_condition_with_multipliers makes three kinds of workspaces -- the inf-sup filter's
    asm_mass_matrix inside model::actualize_sizes, the constraint matrix asm_mass_matrix(B, mim, mf_mult, mf_u, region) on the
    REDUCED multiplier fem, and asm_source_term on it ("A:Test_u"): the coupled tree is the rectangular mass family (shim tag 1001),
    the load a source term or a run-time compiled term; nothing is left to the "not recognised" path."""
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/model_test not built (needs the reference sources)")
    for case in ("model=poisson dim=2 n=3 gt=pk k=2 dirichlet=mult", "model=elasticity dim=3 n=2 gt=pk k=2 dirichlet=mult"):
        out = subprocess.run([BIN] + case.split(), capture_output=True, text=True, timeout=300, env=dict(os.environ, GFGPU_DRYRUN="1"))
        assert out.returncode == 0, out.stderr[-1500:]
        lines = [l for l in out.stderr.splitlines() if l.startswith("[gfgpu dryrun] order")]
        coupled = [l for l in lines if "Test_u1:Test2_u2" in l]
        assert len(coupled) >= 2 and all("recognised family 1001" in l for l in coupled), coupled
        loads = [l for l in lines if "A:Test_u" in l]
        assert loads and all("recognised family 6 " in l or "recognised family 11 " in l for l in loads), loads
        assert not [l for l in lines if "NOT recognised" in l and "order 1 " in l], lines


@pytest.mark.parametrize("kind", ["asm_mass_rect", "asm_mass_rect_volume", "asm_mass_rect_partial", "asm_mass_partial_both"])
def test_mass_matrices_on_two_fems_are_the_rectangular_family(kind):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/model_test not built (needs the reference sources)")
    out = subprocess.run([BIN, "model=" + kind, "dim=2", "n=3", "gt=pk", "k=2"], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, GFGPU_DRYRUN="1"))
    assert out.returncode == 0, out.stderr[-1500:]
    lines = [l for l in out.stderr.splitlines() if l.startswith("[gfgpu dryrun] order 2")]
    assert lines and all("recognised family 1001" in l for l in lines), lines


def _zz_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("zz_law_operators", os.path.join(ROOT, "tests", "test_gpu_zz_law_operators.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cases = [(["model=expr"] + m.split() + ["expr=" + e]) for m, e in mod.COMPOUND + mod.NORM_POTENTIALS + mod.CONSTANTS]
    return cases + [c.split() for c in mod.MODELS]


@pytest.mark.parametrize("args", _zz_cases(), ids=lambda a: " ".join(a)[:60])
def test_every_case_of_the_last_gpu_file_translates_and_compiles(args):
    """CPU: the end-to-end GPU cases that could not be run on a B200 in this round (tests/test_gpu_zz_law_operators.py) at least
    take the NVRTC route in the dry run, and every form the translator emits for them compiles for sm_100a."""
    from getfem_b200 import capi
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/model_test not built (needs the reference sources)")
    out = subprocess.run([BIN] + args, capture_output=True, text=True, timeout=300, env=dict(os.environ, GFGPU_DRYRUN="1"))
    assert out.returncode == 0, out.stderr[-1500:]
    forms = [l.split("\t") for l in out.stderr.splitlines() if l.startswith("[gfgpu dryrun jit]")]
    assert forms, out.stderr[-1500:]
    assert not [l for l in out.stderr.splitlines() if l.startswith("[gfgpu dryrun] order 1") and "NOT recognised" in l]
    seen = set()
    for head, f1, f2, f0 in forms:
        if (f1, f2, f0) in seen:
            continue
        seen.add((f1, f2, f0))
        m = re.search(r"dim=(\d) qdim=(\d)", head)
        capi.jit_check(int(m.group(1)), f1, f2, qdim=int(m.group(2)))
        if f0:
            capi.jit_check(int(m.group(1)), "(" + f0 + ")*dot(tv,tv)", f2, qdim=int(m.group(2)))
    if args[0] == "model=expr" and ("potential" in " ".join(args) or any(a.startswith("expr=pow(") or a.startswith("expr=0.8*(Matrix_j1") for a in args)):
        assert all(f[3] for f in forms), "the order-0 form is missing"
