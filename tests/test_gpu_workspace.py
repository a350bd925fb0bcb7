"""GPU: the ga_workspace mirror end to end -- own regular mesh, DEVICE dof enumeration, own tables --
against the reference goldens (dof numbering and CSC pattern bit-exact, values 1e-12) and, at sizes
the oracle finishes in seconds, against the CPU oracle; plus size-independent properties."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu

EXPR = {  # {s}: suffix of the constants' names when several expressions share a workspace
    "laplace": "a{s}*Grad_u.Grad_Test_u",
    "mass": "a{s}*u.Test_u",
    "elast": "(Div_u*((lambda{s})*Id(meshdim))+(2*(mu{s}))*Sym(Grad_u)):Grad_Test_u",
    "svk": "((Id(meshdim)+Grad_u)*(Saint_Venant_Kirchhoff_PK2(Grad_u,params{s}))):Grad_Test_u",
    "nh_ciarlet": "((Id(meshdim)+Grad_u)*(Compressible_Neo_Hookean_Ciarlet_PK2(Grad_u,params{s}))):Grad_Test_u",
    "nh_bonet": "((Id(meshdim)+Grad_u)*(Compressible_Neo_Hookean_Bonet_PK2(Grad_u,params{s}))):Grad_Test_u",
    "mooney_rivlin": "((Id(meshdim)+Grad_u)*(Compressible_Mooney_Rivlin_PK2(Grad_u,params{s}))):Grad_Test_u",
    "ciarlet_geymonat": "((Id(meshdim)+Grad_u)*(Ciarlet_Geymonat_PK2(Grad_u,params{s}))):Grad_Test_u",
    "blatz_ko": "((Id(meshdim)+Grad_u)*(Generalized_Blatz_Ko_PK2(Grad_u,params{s}))):Grad_Test_u",
    "source": "-f{s}.Test_u",  # "-f*Test_u" when qdim = 1 (the strings of oracle/ref_driver.cc)
    "nsource": "(Reshape(g{s},qdim(u),meshdim)*Normal).Test_u",  # "((g).Normal)*Test_u" when qdim = 1
}


def add_term(ws, mim, m, Q, family, params, region, sfx="", fields=None):
    from conftest import make_region
    import getfem_b200 as gf
    expr = EXPR[family].format(s=sfx)
    if fields is not None:  # fem-data coefficients on a classical data fem of degree kd: (kd, [nodal values ...], dof table)
        kd, vals, ded = fields
        mfd = gf.mesh_fem(m, Q if family == "source" else 1)
        mfd.set_classical_finite_element(kd)
        if ded is not None:
            assert np.array_equal(mfd.ind_scalar_basic_dof_of_element(), ded), "data fem numbering differs from the reference"
        names = {"laplace": ["a"], "mass": ["a"], "elast": ["lambda", "mu"], "source": ["f"]}[family]
        for nm, v in zip(names, vals):
            ws.add_fem_constant(nm + sfx, mfd, v(mfd) if callable(v) else v)
        if family == "source" and Q == 1:
            expr = "-f%s*Test_u" % sfx
        rg = make_region(m, region)
        ws.add_expression(expr, mim, rg)
        ws.data_fem = mfd
        return rg
    if family == "source":
        ws.add_fixed_size_constant("f" + sfx, [-p for p in params])  # the goldens carry F = -f
        if Q == 1:
            expr = "-f%s*Test_u" % sfx
    elif family == "nsource":
        ws.add_fixed_size_constant("g" + sfx, params)
        if Q == 1:
            expr = "((g%s).Normal)*Test_u" % sfx
    elif family in ("laplace", "mass"):
        ws.add_fixed_size_constant("a" + sfx, [params[0]])
    elif family == "elast":
        ws.add_fixed_size_constant("lambda" + sfx, [params[0]])
        ws.add_fixed_size_constant("mu" + sfx, [params[1]])
    else:
        ws.add_fixed_size_constant("params" + sfx, params)
    rg = make_region(m, region)
    ws.add_expression(expr, mim, rg)
    return rg


def build_ws(dim, nsub, gt, k, Q, im, family, params, U=None, region=None, extra=(), fields=None, pts=None):
    import getfem_b200 as gf
    m = gf.mesh()
    gf.regular_unit_mesh(m, nsub, "GT_%s(%d,1)" % (gt, dim))
    if pts is not None:  # distorted fixture: same numbering (the reference re-adds the convexes in order), moved nodes
        assert pts.shape == m.pts.shape and np.abs(pts - m.pts).max() < 0.5 / min(nsub)
        m.pts = np.ascontiguousarray(pts, np.float64)
    mf = gf.mesh_fem(m, Q)
    mf.set_classical_finite_element(k)
    mim = gf.mesh_im(m)
    mim.set_integration_method(im)
    ws = gf.ga_workspace()
    ndof = mf.nb_dof()
    if U is None:
        U = np.zeros(ndof)
    elif callable(U):
        U = U(mf)
    ws.add_fem_variable("u", mf, slice(0, ndof), U)
    ws.region = add_term(ws, mim, m, Q, family, params, region, fields=fields)
    for t, (fam2, fp2, _, rgname2) in enumerate(extra):
        add_term(ws, mim, m, Q, fam2, list(fp2), rgname2, str(t + 2))
    return ws, mf, m, U


@pytest.mark.parametrize("name", golden_names(mirror=True))
def test_workspace_matches_reference_golden(name):
    g = load_golden(name)
    a = g["args"]
    dim = int(a["dim"])
    nsub = [int(a["n"])] * dim if "n" in a else [int(a["nx"]), int(a["ny"]), int(a["nz"])][:dim]
    ws, mf, m, _ = build_ws(dim, nsub, "PK" if g["gt_linear"] else "QK", int(a["k"]), g["Q"], int(a["im"]),
                            g["family"], g["fparams"], g["U"], a.get("region"), g["extra_terms"],
                            None if g["fields"] is None else
                            (g["fields"]["kd"], [(-v if g["family"] == "source" else v) for v in g["fields"]["vals"]],
                             g["fields"]["d_elem_dof"]),
                            pts=g["pts"] if "noise" in a else None)
    # device first-touch numbering == mesh_fem::enumerate_dof, bit for bit
    assert mf.nb_dof() == g["meta"]["ndof"]
    assert np.array_equal(mf.ind_scalar_basic_dof_of_element(), g["elem_dof"])
    ws.assembly(2)
    ws.assembly(1)
    jc, ir, pr = ws.assembled_matrix()
    assert np.array_equal(jc, g["K_jc"]) and np.array_equal(ir, g["K_ir"])
    # Q4: the reference's own basis tables carry 5e-10 of round-off (see test_host_tables.py)
    tol = 1e-8 if int(a["k"]) >= 4 else 1e-12
    assert np.linalg.norm(pr - g["K_pr"]) / max(np.linalg.norm(g["K_pr"]), 1e-300) < tol
    assert np.linalg.norm(ws.assembled_vector() - g["R"]) / np.linalg.norm(g["R"]) < tol


def smooth_u(amp):
    def f(mf):
        X = mf.basic_dof_nodes()
        dim = X.shape[1]
        k = np.arange(X.shape[0]) % mf.Qdim
        xk = X[np.arange(X.shape[0]), k % dim]
        xn = X[np.arange(X.shape[0]), (k + 1) % dim]
        return amp * np.sin(2 * np.pi * xn) * np.cos(np.pi * xk)
    return f


CASES = [  # dim, nsub, gt, k, Q, im, family, params, U
    (3, [6, 5, 4], "PK", 2, 3, 4, "elast", [1.0, 1.0], "random"),
    (3, [9, 9, 9], "PK", 1, 1, 2, "laplace", [1.0], "random"),
    (2, [40, 30], "PK", 1, 1, 2, "laplace", [2.0], "random"),
    (3, [4, 3, 3], "QK", 2, 3, 6, "nh_ciarlet", [1.0, 1.0], "smooth"),
    (3, [3, 3, 2], "QK", 2, 3, 6, "svk", [1.3, 0.7], "smooth"),
    (3, [4, 4, 3], "PK", 2, 3, 4, "nh_bonet", [1.3, 0.7], "smooth"),
    (3, [2, 2, 1], "QK", 4, 1, 8, "laplace", [1.0], "random"),
    (3, [5, 4, 3], "PK", 2, 3, 4, "mass", [1.5], "random"),
    # low-order scalar-coefficient forms: the column kernel (csrc/recompute_cols.cu) with every (table size, qdim) pair
    (3, [6, 5, 4], "PK", 1, 1, 2, "mass", [1.5], "random"),
    (3, [4, 4, 3], "PK", 1, 3, 2, "mass", [0.8], "random"),
    (2, [12, 9], "PK", 1, 2, 2, "mass", [2.0], "random"),
    (2, [10, 8], "PK", 1, 2, 2, "laplace", [0.7], "random"),
    (3, [4, 3, 5], "PK", 1, 3, 2, "laplace", [1.1], "random"),
    (2, [30, 20], "PK", 1, 1, 2, "source", [-1.5], "random"),
    (3, [3, 2, 2], "QK", 2, 3, 6, "source", [0.5, -1.0, 2.0], "random"),
    # mesh regions: boundary faces (Robin / penalisation mass, Neumann source, normal source) and sub-regions of convexes
    (3, [5, 4, 3], "PK", 2, 3, 4, "mass", [1.5], "random", "outer"),
    (3, [6, 5, 4], "PK", 2, 3, 4, "source", [0.5, -1.0, 2.0], "random", "xmax"),
    (3, [5, 4, 4], "PK", 2, 3, 4, "nsource", [0.3, -0.2, 0.9, 1.1, 0.4, -0.7, 0.6, 0.1, -1.3], "random", "outer"),
    (2, [30, 20], "PK", 1, 1, 2, "nsource", [0.8, -0.4], "random", "outer"),
    (2, [12, 9], "PK", 2, 2, 4, "mass", [2.0], "random", "outer"),
    (3, [4, 3, 3], "QK", 2, 1, 6, "mass", [1.5], "random", "zmin"),
    (3, [3, 3, 2], "QK", 2, 3, 6, "nsource", [0.3, -0.2, 0.9, 1.1, 0.4, -0.7, 0.6, 0.1, -1.3], "random", "outer"),
    (2, [9, 7], "QK", 2, 1, 4, "mass", [1.0], "random", "outer"),
    (3, [6, 5, 4], "PK", 2, 3, 4, "elast", [1.0, 1.0], "random", "half"),
    (3, [9, 9, 9], "PK", 1, 1, 2, "laplace", [1.0], "random", "half"),
    (3, [4, 3, 3], "QK", 2, 3, 6, "nh_ciarlet", [1.0, 1.0], "smooth", "half"),
    (3, [2, 2, 1], "QK", 4, 1, 8, "laplace", [1.0], "random", "half"),
    # fem-data coefficients (add_fem_constant) on a data fem of degree kd (last entry), volume and boundary regions
    (3, [5, 4, 3], "PK", 2, 1, 4, "laplace", [1.0], "random", None, 1),
    (3, [5, 4, 3], "PK", 2, 3, 4, "elast", [1.0, 1.0], "random", None, 2),
    (2, [14, 11], "PK", 2, 2, 4, "elast", [1.0, 1.0], "random", "half", 1),
    (3, [5, 4, 3], "PK", 2, 3, 4, "source", [1.0, 1.0, 1.0], "random", None, 2),
    (3, [5, 4, 3], "PK", 2, 3, 4, "source", [1.0, 1.0, 1.0], "random", "xmax", 1),
    (3, [5, 4, 3], "PK", 2, 1, 4, "mass", [1.0], "random", "outer", 2),
    (3, [3, 3, 2], "QK", 2, 1, 6, "laplace", [1.0], "random", None, 2),
    (3, [3, 3, 2], "QK", 2, 3, 6, "source", [1.0, 1.0, 1.0], "random", "zmin", 1),
]


def field_functions(family, Q):
    """Analytic coefficient fields sampled at the data fem's dof nodes."""
    def scal(c0, kx, ky, kz):
        def f(mfd):
            X = mfd.basic_dof_nodes()
            z = X[:, 2] if X.shape[1] > 2 else 0.0
            return c0 * (1.0 + 0.3 * np.sin(kx * X[:, 0] + ky * X[:, 1] + kz * z))
        return f

    def vec(mfd):
        X = mfd.basic_dof_nodes()
        comp = np.arange(X.shape[0]) % Q
        return 0.7 * (comp + 1) * (1.0 + 0.5 * X[:, 0] - 0.25 * X[:, -1])
    if family == "source":
        return [vec]
    if family == "elast":
        return [scal(1.2, 1.1, -0.7, 0.3), scal(0.8, 1.3, -1.0, 0.0)]
    return [scal(1.3, 1.7, 0.9, 0.4)]


def oracle_region(rg, t, ft):
    """Region + all-point tables in the oracle's layout (volume points, then face after face)."""
    if rg is None:
        return None, (t["quad_w"], t["gt_grad"], t["phi"], t["gphi"])
    cv, fc = rg.items()
    if not rg.is_only_faces():
        return {"items_cv": cv, "items_f": fc}, (t["quad_w"], t["gt_grad"], t["phi"], t["gphi"])
    nq = len(t["quad_w"])
    nf, nqf = ft["quad_w"].shape
    reg = {"items_cv": cv, "items_f": fc, "face_first": nq + nqf * np.arange(nf), "face_nq": np.full(nf, nqf),
           "ref_normals": ft["normals"]}
    cat = lambda a, b: np.concatenate([a, b.reshape((nf * nqf,) + a.shape[1:])])  # noqa: E731
    return reg, (cat(t["quad_w"], ft["quad_w"]), cat(t["gt_grad"], ft["gt_grad"]), cat(t["phi"], ft["phi"]),
                 cat(t["gphi"], ft["gphi"]))


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s_%s%d_q%d_%s%s%s" % (
    c[6], c[2], c[3], c[4], "x".join(map(str, c[1])), "_" + c[9] if len(c) > 9 and c[9] else "",
    "_coefk%d" % c[10] if len(c) > 10 else ""))
def test_workspace_matches_oracle(case):
    from oracle import oracle
    from getfem_b200 import fem_tables
    dim, nsub, gt, k, Q, im, family, params, umode = case[:9]
    region = case[9] if len(case) > 9 else None
    kd = case[10] if len(case) > 10 else None
    if umode == "random":
        rng = np.random.default_rng(7)
        U = lambda mf: rng.uniform(-1, 1, mf.nb_dof())  # noqa: E731
    else:
        U = smooth_u(0.03)
    fields = None if kd is None else (kd, field_functions(family, Q), None)
    ws, mf, m, Uv = build_ws(dim, nsub, gt, k, Q, im, family, params, U, region, fields=fields)
    ws.assembly(2)
    ws.assembly(1)
    jc, ir, pr = ws.assembled_matrix()
    t = fem_tables.classical_tables(gt, dim, k, im)
    ft = fem_tables.classical_face_tables(gt, dim, k, im) if region in ("outer", "xmax", "zmin") else None
    reg, (w, gtg, phi, gphi) = oracle_region(ws.region, t, ft)
    ofields = None
    if kd is not None:  # the data fem's basis at the same points as the other tables; "-f.Test_u" integrates F = -f
        mfd = ws.data_fem
        X = t["quad_x"] if ft is None else np.concatenate([t["quad_x"], ft["quad_x"].reshape(-1, dim)])
        vals = [fn(mfd) for fn in field_functions(family, Q)]
        ofields = {"d_elem_dof": mfd.ind_scalar_basic_dof_of_element(), "d_phi": fem_tables.lagrange_tables(gt, dim, kd, X)[0],
                   "vals": [-v for v in vals] if family == "source" else vals}
    ojc, oir, opr, oR = oracle.assemble(m.pts, m.conn, mf.ind_scalar_basic_dof_of_element(), mf.nb_dof(), Q,
                                        w, gtg, phi, gphi, gt == "PK", family, params, Uv, region=reg,
                                        nq=len(t["quad_w"]), fields=ofields)
    assert np.array_equal(jc, ojc) and np.array_equal(ir, oir)
    assert np.linalg.norm(pr - opr) / max(np.linalg.norm(opr), 1e-300) < 1e-12
    assert np.linalg.norm(ws.assembled_vector() - oR) / np.linalg.norm(oR) < 1e-12
    if family == "nsource" and region == "outer" and Q == dim:
        # divergence theorem on the closed boundary: sum_i int (A n)_b phi_i = int (A n)_b = 0 for a constant A
        assert np.abs(ws.assembled_vector().reshape(-1, Q).sum(0)).max() < 1e-12
    if family == "mass" and region == "outer" and Q == 1 and kd is None:
        # 1^T M 1 = measure of the boundary of the unit square / cube
        import scipy.sparse as sp
        M = sp.csc_matrix((pr, ir, jc), shape=(mf.nb_dof(),) * 2)
        assert abs(M.sum() / params[0] - 2 * dim) < 1e-11


def test_config1_stiffness_plus_rhs():
    """BASELINE config 1 as the reference states it: 2D Poisson P1, stiffness AND right-hand side in one workspace
    ("a*Grad_u.Grad_Test_u" + "-f*Test_u").  assembly(2) = K (the source term has no order-2 tree), assembly(1) =
    K u - F; checked against the oracle term by term."""
    import getfem_b200 as gf
    from getfem_b200 import fem_tables
    from oracle import oracle
    n = 24
    m = gf.mesh()
    gf.regular_unit_mesh(m, [n, n], "GT_PK(2,1)")
    mf = gf.mesh_fem(m, 1)
    mf.set_classical_finite_element(1)
    mim = gf.mesh_im(m)
    mim.set_integration_method(2)
    rng = np.random.default_rng(11)
    U = rng.uniform(-1, 1, mf.nb_dof())
    ws = gf.ga_workspace()
    ws.add_fem_variable("u", mf, slice(0, mf.nb_dof()), U)
    ws.add_fixed_size_constant("a", [1.0])
    ws.add_fixed_size_constant("f", [3.0])
    ws.add_expression("a*Grad_u.Grad_Test_u", mim)
    ws.add_expression("-f*Test_u", mim)
    ws.assembly(2)
    ws.assembly(1)
    jc, ir, pr = ws.assembled_matrix()
    t = fem_tables.classical_tables("PK", 2, 1, 2)
    args = (m.pts, m.conn, mf.ind_scalar_basic_dof_of_element(), mf.nb_dof(), 1, t["quad_w"], t["gt_grad"], t["phi"],
            t["gphi"], True)
    ojc, oir, opr, oR = oracle.assemble(*args, "laplace", [1.0], U)
    sjc, sir, spr, sR = oracle.assemble(*args, "source", [-3.0], U)
    assert sjc[-1] == 0, "a source term must not produce tangent entries"
    assert np.array_equal(jc, ojc) and np.array_equal(ir, oir)
    assert np.linalg.norm(pr - opr) / np.linalg.norm(opr) < 1e-12
    assert np.linalg.norm(ws.assembled_vector() - (oR + sR)) / np.linalg.norm(oR + sR) < 1e-12
    assert abs(sR.sum() + 3.0) < 1e-12  # integral of -f over the unit square


def test_properties_at_size():
    """Size-independent checks on a mesh too large for the oracle to be practical in a test:
    stencil nnz count of the P1 Laplacian (7-point: the drop rule removes the exact zeros), K*1 = 0,
    symmetry, residual == K*u for the linear form."""
    import scipy.sparse as sp
    n = 40
    rng = np.random.default_rng(3)
    ws, mf, m, U = build_ws(3, [n, n, n], "PK", 1, 1, 2, "laplace", [1.0], lambda mf: rng.uniform(-1, 1, mf.nb_dof()))
    ws.assembly(2)
    ws.assembly(1)
    jc, ir, pr = ws.assembled_matrix()
    nd = mf.nb_dof()
    assert nd == (n + 1) ** 3
    assert jc[-1] == 7 * (n + 1) ** 3 - 6 * (n + 1) ** 2
    K = sp.csc_matrix((pr, ir, jc), shape=(nd, nd))
    assert abs(K - K.T).max() < 1e-13 * abs(K).max()
    assert np.abs(K @ np.ones(nd)).max() < 1e-12 * abs(K).max()
    R = ws.assembled_vector()
    assert np.linalg.norm(K @ U - R) / np.linalg.norm(R) < 1e-12
