"""GPU: coupled bilinear terms (gfgpu_rect_*): the off-diagonal blocks of mixed formulations, "-(Test2_p*Div_Test_u)" and
its transpose.  Device against the numpy restatement of the reference's element loop + drop rule (oracle.rect_div_pressure;
the restatement itself is pinned on reference fixtures in tests/test_oracle.py): CSC pattern identical, values 1e-12; the
transposed block, the products (residual parts R_u = B p, R_p = B^T u) and the saddle-point matrix [[A, B], [B^T, 0]]
accumulated on the device against scipy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [  # gt, dim, (k_u, k_p), im, nsub, distortion
    ("PK", 3, (2, 1), 4, [2, 2, 3], 0.15),   # Taylor-Hood tetrahedra
    ("PK", 2, (2, 1), 4, [4, 3], 0.15),
    ("QK", 3, (2, 1), 6, [2, 2, 2], 0.12),   # Q2-Q1 hexahedra, genuinely trilinear
    ("QK", 2, (2, 1), 6, [3, 4], 0.12),
    ("PK", 3, (3, 2), 5, [1, 2, 1], 0.1),
    ("PK", 3, (1, 1), 2, [3, 3, 3], 0.0),    # equal order on a regular mesh: exact zeros fall to the drop rule
]


def _setup(gt, dim, ks, im, nsub, distort):
    import getfem_b200 as gf
    from getfem_b200 import capi, fem_tables
    ctx = capi.Context(0)
    m = gf.mesh()
    gf.regular_unit_mesh(m, nsub, "GT_%s(%d,1)" % (gt, dim))
    rng = np.random.default_rng(5)
    if distort:
        m.pts = m.pts + distort / max(nsub) * rng.uniform(-1, 1, m.pts.shape)
        m._dev = {}
    mfu, mfp = gf.mesh_fem(m, dim), gf.mesh_fem(m, 1)
    mfu.set_classical_finite_element(ks[0])
    mfp.set_classical_finite_element(ks[1])
    dmesh = m.device(ctx)
    dfu, dfp = mfu.device(ctx), mfp.device(ctx)
    tu = fem_tables.classical_tables(gt, dim, ks[0], im)
    tp = fem_tables.classical_tables(gt, dim, ks[1], im)
    assert np.array_equal(tu["quad_w"], tp["quad_w"])
    tabu = capi.DeviceTables(ctx, tu["quad_w"], tu["gt_grad"], tu["phi"], tu["gphi"])
    tabp = capi.DeviceTables(ctx, tp["quad_w"], tp["gt_grad"], tp["phi"], tp["gphi"])
    return ctx, m, dmesh, dfu, dfp, tu, tp, tabu, tabp, rng


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s%dd-k%d%d" % (c[0], c[1], c[2][0], c[2][1]))
def test_div_pressure_block_matches_the_restated_reference(case):
    import scipy.sparse as sp
    from getfem_b200 import capi
    from oracle import oracle
    gt, dim, ks, im, nsub, distort = case
    ctx, m, dmesh, dfu, dfp, tu, tp, tabu, tabp, rng = _setup(*case)
    coef, alpha = 1.7, -1.0
    r = capi.DeviceRect(ctx, dmesh, dfu, tabu, dfp, tabp, capi.RECT_DIV_PRESSURE, coef, alpha)
    r.assemble()
    jc, ir, pr = r.export_csc()
    ojc, oir, opr = oracle.rect_div_pressure(m.pts, m.conn, dfu.elem_dof(), dim, dfp.elem_dof(), tu["quad_w"], tu["gt_grad"],
                                             tu["gphi"], tp["phi"], coef)
    assert np.array_equal(jc, ojc) and np.array_equal(ir, oir), "pattern differs from the restated reference"
    assert np.linalg.norm(pr - alpha * opr) <= 1e-12 * np.linalg.norm(opr)
    Bm = sp.csc_matrix((pr, ir, jc), shape=(dfu.ndof, dfp.ndof))
    # the transposed block = the tree with the test functions swapped
    tjc, tir, tpr = r.export_csc(transposed=True)
    Bt = sp.csc_matrix((tpr, tir, tjc), shape=(dfp.ndof, dfu.ndof))
    assert (Bt - Bm.T).nnz == 0 and Bt.has_sorted_indices
    ref = sp.csc_matrix(Bm.T)
    ref.sort_indices()
    assert np.array_equal(tjc, ref.indptr) and np.array_equal(tir, ref.indices) and np.array_equal(tpr, ref.data)
    # a second assembly reuses the pattern and gives the same bits
    r.assemble()
    assert np.array_equal(r.export_csc()[2], pr)
    # residual parts
    p, u = rng.uniform(-1, 1, dfp.ndof), rng.uniform(-1, 1, dfu.ndof)
    assert np.linalg.norm(r.mult(p) - Bm @ p) <= 1e-13 * np.linalg.norm(Bm @ p)
    assert np.linalg.norm(r.mult(u, transposed=True) - Bm.T @ u) <= 1e-13 * np.linalg.norm(Bm.T @ u)
    y0 = rng.uniform(-1, 1, dfu.ndof)
    assert np.linalg.norm(r.mult(p, alpha=-2.0, beta=1.0, y=y0) - (y0 - 2 * (Bm @ p))) <= 1e-13 * np.linalg.norm(y0)
    # div of a constant field vanishes: every column of B sums to (boundary terms aside) -> B^T applied to a constant vector
    # field equals the boundary flux; with u = const the divergence is zero everywhere: B^T u = 0
    uc = np.tile(rng.uniform(-1, 1, dim), dfu.ndof // dim)
    assert np.abs(r.mult(uc, transposed=True)).max() <= 1e-12 * np.abs(pr).max() * 50


def test_saddle_point_matrix_on_the_device():
    """[[A, B], [B^T, 0]] of incompressible elasticity: stiffness term + the coupled block and its transpose, one resident matrix"""
    import scipy.sparse as sp
    from getfem_b200 import capi
    ctx, m, dmesh, dfu, dfp, tu, tp, tabu, tabp, rng = _setup("PK", 3, (2, 1), 4, [2, 2, 2], 0.1)
    nu, npp = dfu.ndof, dfp.ndof
    term = capi.DeviceTerm(ctx, dmesh, dfu, tabu, "elast", [0.0, 1.0])
    term.assemble_host(None, capi.TANGENT, None, None)
    r = capi.DeviceRect(ctx, dmesh, dfu, tabu, dfp, tabp, capi.RECT_DIV_PRESSURE, 1.0, -1.0)
    r.assemble()
    K = capi.DeviceMatrix(ctx, nu + npp)
    K.add_term(term)
    K.add_rect(r, False, 1.0, 0, nu)
    K.add_rect(r, True, 1.0, nu, 0)
    jc, ir, pr = K.export_csc()
    S = sp.csc_matrix((pr, ir, jc), shape=(nu + npp, nu + npp))
    A = sp.csc_matrix(term.export_csc()[::-1][0:1] + (term.export_csc()[1], term.export_csc()[0]), shape=(nu, nu))
    Bm = sp.csc_matrix((r.export_csc()[2], r.export_csc()[1], r.export_csc()[0]), shape=(nu, npp))
    R = sp.bmat([[A, Bm], [Bm.T, None]], format="csc")
    R.sort_indices()
    assert (abs(S - R)).max() == 0.0 and S.nnz == R.nnz
    x = rng.uniform(-1, 1, nu + npp)
    assert np.linalg.norm(K.mult(x) - R @ x) <= 1e-13 * np.linalg.norm(R @ x)
    assert abs(S - S.T).max() <= 1e-15 * abs(S).max()


def test_rect_error_paths():
    from getfem_b200 import capi
    ctx, m, dmesh, dfu, dfp, tu, tp, tabu, tabp, rng = _setup("PK", 3, (2, 1), 4, [1, 1, 1], 0.0)
    with pytest.raises(capi.GfgpuError, match="vector rows"):
        capi.DeviceRect(ctx, dmesh, dfp, tabp, dfp, tabp)
    with pytest.raises(capi.GfgpuError, match="tables do not match"):
        capi.DeviceRect(ctx, dmesh, dfu, tabp, dfp, tabp)
    r = capi.DeviceRect(ctx, dmesh, dfu, tabu, dfp, tabp)
    with pytest.raises(capi.GfgpuError, match="no assembled"):
        r.export_csc()


def _coupled_names():
    import glob
    import os
    from conftest import GOLDEN
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "coupled", "*.npz")))


@pytest.mark.parametrize("name", _coupled_names())
def test_device_blocks_match_the_reference_fixtures(name):
    """C ABI with the reference's own meshes, dof tables and reference tables (tests/golden/coupled, generated by the
    unmodified reference): both blocks of "-p*Div_Test_u - Test_p*Div_u" pattern bit-exact, values 1e-12; residual parts."""
    from getfem_b200 import capi
    from test_coupled_oracle import load_coupled
    g = load_coupled(name)
    dim, nu, npp = g["meta"]["dim"], g["meta"]["ndof_u"], g["meta"]["ndof_p"]
    qk = "QK" in g["meta"]["fem_u"]
    ctx = capi.Context(0)
    mesh = capi.DeviceMesh(ctx, g["pts"], g["conn"], capi.GT_QK if qk else capi.GT_PK)
    ku, kp = int(g["meta"]["fem_u"][-2]), int(g["meta"]["fem_p"][-2])
    kind = capi.FEM_QK if qk else capi.FEM_PK
    fu = capi.DeviceFem(ctx, mesh, kind, ku, dim, g["elem_dof_u"].shape[1], g["elem_dof_u"], nu)
    fp = capi.DeviceFem(ctx, mesh, kind, kp, 1, g["elem_dof_p"].shape[1], g["elem_dof_p"], npp)
    tu = capi.DeviceTables(ctx, g["quad_w"], g["gt_grad"], g["phi_u"], g["gphi_u"])
    tp = capi.DeviceTables(ctx, g["quad_w"], g["gt_grad"], g["phi_p"], g["gphi_p"])
    r = capi.DeviceRect(ctx, mesh, fu, tu, fp, tp, capi.RECT_DIV_PRESSURE, 1.0, -1.0)  # "(-Test2_p)*Div_Test_u"
    r.assemble()
    jc, ir, pr = r.export_csc()
    assert np.array_equal(jc, g["Kup_jc"]) and np.array_equal(ir, g["Kup_ir"]), "pattern differs from the reference"
    assert np.linalg.norm(pr - g["Kup_pr"]) <= 1e-12 * np.linalg.norm(g["Kup_pr"])
    tjc, tir, tpr = r.export_csc(transposed=True)                                      # "-(Test_p*Div_Test2_u)"
    assert np.array_equal(tjc, g["Kpu_jc"]) and np.array_equal(tir, g["Kpu_ir"])
    assert np.linalg.norm(tpr - g["Kpu_pr"]) <= 1e-12 * np.linalg.norm(g["Kpu_pr"])
    R = g["R"]
    assert np.linalg.norm(r.mult(g["P"]) - R[:nu]) <= 1e-12 * np.linalg.norm(R[:nu])
    assert np.linalg.norm(r.mult(g["U"], transposed=True) - R[nu:]) <= 1e-12 * np.linalg.norm(R[nu:])
