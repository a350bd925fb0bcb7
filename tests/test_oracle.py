"""CPU: the plain-C oracle (oracle/asm_oracle.c) against golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  Pattern bit-exact; values 1e-13."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import oracle


@pytest.mark.parametrize("name", golden_names(oracle_only=True))
def test_oracle_matches_reference(name):
    g = load_golden(name)
    ndof = g["meta"]["ndof"]
    w, gt_grad, phi, gphi = g["tables"]
    jc, ir, pr, R = oracle.assemble(g["pts"], g["conn"], g["elem_dof"], ndof, g["Q"], w, gt_grad, phi, gphi,
                                    g["gt_linear"], g["family"], g["fparams"], g["U"], region=g["region"],
                                    nq=g["meta"]["nq"], fields=g["fields"])
    if g["extra_terms"]:  # several expressions in one workspace: every tree adds into the same K / V
        from conftest import csc_sum
        mats = [(jc, ir, pr)]
        for fam, fp, rg, _ in g["extra_terms"]:
            jc2, ir2, pr2, R2 = oracle.assemble(g["pts"], g["conn"], g["elem_dof"], ndof, g["Q"], w, gt_grad, phi, gphi,
                                                g["gt_linear"], fam, fp, g["U"], region=rg, nq=g["meta"]["nq"])
            mats.append((jc2, ir2, pr2))
            R = R + R2
        jc, ir, pr = csc_sum(mats, ndof)
    assert np.array_equal(jc, g["K_jc"]), "column pointers differ (pattern not bit-exact)"
    assert np.array_equal(ir, g["K_ir"]), "row indices differ (pattern not bit-exact)"
    rel = np.linalg.norm(pr - g["K_pr"]) / max(np.linalg.norm(g["K_pr"]), 1e-300)
    assert rel < 1e-13, rel
    relr = np.linalg.norm(R - g["R"]) / max(np.linalg.norm(g["R"]), 1e-300)
    assert relr < 1e-13, relr


def test_oracle_penalty_branch_detF_negative():
    """det(I+Grad_u) <= 0 triggers the +1e200*C term of the Neo-Hookean law
    (getfem_nonlinear_elasticity.cc:655-656); the oracle must reproduce it (values blow up)."""
    g = load_golden("x_nh_ciarlet_p1tet_n2")
    U = g["U"] * 0.0
    xyz = g["dof_xyz"]
    U[0::3] = -3.0 * xyz[0::3, 0]  # u_x = -3x  -> F_xx = -2, det F < 0
    _, _, pr, R = oracle.assemble(g["pts"], g["conn"], g["elem_dof"], g["meta"]["ndof"], 3, g["quad_w"], g["gt_grad"],
                                  g["phi"], g["gphi"], True, "nh_ciarlet", g["fparams"], U)
    assert np.abs(R).max() > 1e190


@pytest.mark.parametrize("family", ["svk", "nh_ciarlet", "nh_bonet", "mooney_rivlin", "ciarlet_geymonat", "blatz_ko"])
def test_oracle_law_derivatives(family):
    """The reference's own check of its hyperelastic laws, abstract_hyperelastic_law::test_derivatives
    (getfem_nonlinear_elasticity.cc:298-347: 100 random states, dsigma.DE against sigma(E+DE)-sigma(E), 1.5e-4 relative),
    restated on the oracle's material point in the variable the GWFL operators use (Grad_u).  A central difference keeps the
    truncation error below the reference's tolerance for every sample."""
    rng = np.random.default_rng(7)
    par = {"mooney_rivlin": np.array([0.8, 0.3, 2.0]), "ciarlet_geymonat": np.array([1.3, 0.7, 0.25]),
           "blatz_ko": np.array([1.0, 1.0, 1.5, -0.5, 1.5])}.get(family, np.array([1.3, 0.7]))
    h, done = 1e-6, 0
    while done < 100:
        Gu = 0.4 * rng.uniform(-1, 1, (3, 3))
        if np.linalg.det(np.eye(3) + Gu) < 0.2:
            continue
        D = rng.uniform(-1, 1, (3, 3))
        S, dS = oracle.hyper_law(family, Gu, par)
        Sp, _ = oracle.hyper_law(family, Gu + h * D, par)
        Sm, _ = oracle.hyper_law(family, Gu - h * D, par)
        fd = (Sp - Sm) / (2 * h)
        an = np.einsum("ijkl,kl->ij", dS, D)
        assert np.abs(an - fd).max() <= 1.5e-4 * np.abs(an).max(), (family, done)
        assert np.abs(S - S.T).max() <= 1e-13 * max(np.abs(S).max(), 1.0)  # PK2 is symmetric
        done += 1


_DRIVER = None


def _ref_driver():
    import os
    from conftest import ROOT
    p = os.path.join(ROOT, "oracle", "_ref", "gf_ref_driver")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref/gf_ref_driver not built (needs the reference sources)")
    return p


@pytest.mark.parametrize("case", [
    "family=laplace dim=3 n=3 k=1 q=1 im=2", "family=laplace dim=2 n=6 k=1 q=1 im=2 a=0.7",
    "family=laplace_vec dim=3 n=2 k=1 q=3 im=2", "family=laplace dim=3 n=2 gt=qk k=2 q=1 im=6",
    "family=elast dim=3 n=3 k=2 q=3 im=4 lambda=1.3 mu=0.7", "family=elast dim=2 n=5 k=2 q=2 im=4",
    "family=elast dim=3 n=2 gt=qk k=2 q=3 im=6", "family=mass dim=3 n=3 k=2 q=1 im=4 a=2.5", "family=mass dim=2 n=4 k=1 q=2 im=2",
])
def test_reference_old_vs_new_assembly(case):
    """The reference's own cross-check for this path (tests/test_assembly.cc:279-349, 773-881, tolerance 1e-10): the GWFL
    expression the golden fixtures are generated from against the LEGACY generic_assembly string of the same form, an
    independent code path of the reference (getfem_assembling_tensors.cc).  Pins the fixture generator, hence the oracle."""
    import json
    import subprocess
    out = subprocess.run([_ref_driver()] + case.split() + ["mode=crosscheck"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-1500:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["cross_rel"] < 1e-10, r


@pytest.mark.parametrize("family", ["svk", "nh_ciarlet", "nh_bonet", "mooney_rivlin", "ciarlet_geymonat", "blatz_ko"])
def test_reference_law_derivative_check(family):
    """abstract_hyperelastic_law::test_derivatives run on the reference's own laws (the ones the golden fixtures of the
    finite-strain families were generated with)."""
    import json
    import subprocess
    out = subprocess.run([_ref_driver(), "family=" + family, "dim=3", "n=1", "gt=qk", "k=2", "q=3", "im=6", "mode=lawcheck",
                          "lambda=1.3", "mu=0.7", "a=0.25"],
                         capture_output=True, text=True, timeout=300)
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert out.returncode == 0 and r["lawcheck"], r


def _potentials():
    import json
    import os
    from conftest import GOLDEN
    return json.load(open(os.path.join(GOLDEN, "potentials.json")))


@pytest.mark.parametrize("name", sorted(_potentials()))
def test_oracle_potential_matches_reference(name):
    """order 0: ga_workspace::assembly(0) of the UNMODIFIED reference (tests/golden/make_potentials.py) against the oracle"""
    g = load_golden(name)
    w, gt_grad, phi, gphi = g["tables"]
    E = oracle.potential(g["pts"], g["conn"], g["elem_dof"], g["meta"]["ndof"], g["Q"], w, gt_grad, phi, gphi,
                         g["gt_linear"], g["family"], g["fparams"], g["U"], region=g["region"], nq=g["meta"]["nq"])
    ref = _potentials()[name]["potential"]
    assert abs(E - ref) <= 1e-12 * max(abs(ref), 1e-300), (E, ref)
