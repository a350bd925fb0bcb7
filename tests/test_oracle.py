"""CPU: the plain-C oracle (oracle/asm_oracle.c) against golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  Pattern bit-exact; values 1e-13."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import oracle


@pytest.mark.parametrize("name", golden_names())
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
