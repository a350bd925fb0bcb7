"""GPU: JIT terms (gfgpu_term_create_jit, the NVRTC route).  The run-time kernel against (i) the closed-form families it can
also express -- same tables, same drop rule: identical pattern, values 1e-13 -- and (ii) a numpy restatement of the
quadrature for a genuinely nonlinear form (the order-1 / order-2 trees of "(1+sqr(u))*Grad_u.Grad_Test_u + sin(u)*Test_u +
c*Norm_sqr(Grad_u)*Test_u"), plus the size-independent property tangent = d residual / d u.  The reference itself is the
checker one level up: tests/test_gpu_dropin.py runs such expressions through the shim's tree translation."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MESHES = [("PK", 3, 2, 4, [2, 2, 2], 0.15), ("PK", 2, 2, 4, [4, 4], 0.15), ("QK", 3, 2, 6, [2, 2, 2], 0.12), ("QK", 2, 1, 3, [5, 4], 0.1),
          ("PK", 3, 1, 2, [3, 3, 3], 0.0)]


def _setup(gt, dim, k, im, nsub, distort):
    import getfem_b200 as gf
    from getfem_b200 import capi, fem_tables
    ctx = capi.Context(0)
    m = gf.mesh()
    gf.regular_unit_mesh(m, nsub, "GT_%s(%d,1)" % (gt, dim))
    rng = np.random.default_rng(11)
    if distort:
        m.pts = m.pts + distort / max(nsub) * rng.uniform(-1, 1, m.pts.shape)
        m._dev = {}
    mf = gf.mesh_fem(m, 1)
    mf.set_classical_finite_element(k)
    dmesh, dfem = m.device(ctx), mf.device(ctx)
    t = fem_tables.classical_tables(gt, dim, k, im)
    tab = capi.DeviceTables(ctx, t["quad_w"], t["gt_grad"], t["phi"], t["gphi"])
    U = rng.uniform(-1, 1, dfem.ndof)
    return ctx, m, dmesh, dfem, t, tab, U, rng


def _run(term, U, ndof):
    from getfem_b200 import capi
    R = np.empty(ndof)
    term.assemble_host(U, capi.TANGENT | capi.RESIDUAL, None, R)
    return term.export_csc() + (R,)


@pytest.mark.parametrize("mesh", MESHES, ids=lambda c: "%s%dd-k%d" % (c[0], c[1], c[2]))
def test_jit_reproduces_the_closed_form_families(mesh):
    from getfem_b200 import capi
    ctx, m, dmesh, dfem, t, tab, U, rng = _setup(*mesh)
    for fam, f1, f2 in (("laplace", "par[0]*dot(gu,tg)", "par[0]*dot(t2g,tg)"), ("mass", "par[0]*u*tv", "par[0]*t2v*tv")):
        ref = capi.DeviceTerm(ctx, dmesh, dfem, tab, fam, [1.7], 0.5, capi.STRATEGY_STAGED)
        jc, ir, pr, R = _run(ref, U, dfem.ndof)
        jit = capi.DeviceTerm.jit(ctx, dmesh, dfem, tab, f1, f2, [1.7], 0.5, value_dependent=False)
        jjc, jir, jpr, jR = _run(jit, U, dfem.ndof)
        assert np.array_equal(jc, jjc) and np.array_equal(ir, jir), fam
        assert np.linalg.norm(pr - jpr) <= 1e-13 * np.linalg.norm(pr) and np.linalg.norm(R - jR) <= 1e-13 * np.linalg.norm(R), fam
        jjc2, jir2, jpr2, jR2 = _run(jit, U, dfem.ndof)  # second pass: same bits
        assert np.array_equal(jpr, jpr2) and np.array_equal(jR, jR2)


F1 = "(1.0+sqr(u))*dot(gu,tg) + sin(u)*tv + par[0]*normsqr(gu)*tv"
F2 = "(2.0*u*t2v)*dot(gu,tg) + (1.0+sqr(u))*dot(t2g,tg) + cos(u)*t2v*tv + par[0]*2.0*dot(gu,t2g)*tv"


def _numpy_forms(m, ed, t, U, c):
    """dense restatement: K = G^T pc(q), J = |det K|, B = K^-T, physical gradients B ghat, the two forms at every point"""
    ne, nd = ed.shape
    n = int(ed.max()) + 1
    R = np.zeros(n)
    K = np.zeros((n, n))
    for e in range(ne):
        G = m.pts[m.conn[e]]
        for q in range(len(t["quad_w"])):
            Kq = G.T @ t["gt_grad"][q]
            J = abs(np.linalg.det(Kq))
            B = np.linalg.inv(Kq).T
            ph = t["phi"][q]
            dph = t["gphi"][q] @ B.T  # nd x N
            u = U[ed[e]] @ ph
            gu = U[ed[e]] @ dph
            wj = t["quad_w"][q] * J
            r = (1 + u * u) * (dph @ gu) + np.sin(u) * ph + c * (gu @ gu) * ph
            k = (2 * u) * np.outer(dph @ gu, ph) + (1 + u * u) * (dph @ dph.T) + np.cos(u) * np.outer(ph, ph) \
                + 2 * c * np.outer(ph, dph @ gu)
            R[ed[e]] += wj * r
            K[np.ix_(ed[e], ed[e])] += wj * k
    return K, R


@pytest.mark.parametrize("mesh", MESHES[:4], ids=lambda c: "%s%dd-k%d" % (c[0], c[1], c[2]))
def test_jit_nonlinear_form_against_numpy_and_its_own_derivative(mesh):
    import scipy.sparse as sp
    from getfem_b200 import capi
    ctx, m, dmesh, dfem, t, tab, U, rng = _setup(*mesh)
    U = 0.7 * U
    c = 0.3
    term = capi.DeviceTerm.jit(ctx, dmesh, dfem, tab, F1, F2, [c], 1.0, value_dependent=True)
    jc, ir, pr, R = _run(term, U, dfem.ndof)
    n = dfem.ndof
    S = sp.csc_matrix((pr, ir, jc), shape=(n, n)).toarray()
    Kd, Rd = _numpy_forms(m, dfem.elem_dof(), t, U, c)
    assert np.linalg.norm(R - Rd) <= 1e-12 * np.linalg.norm(Rd)
    assert np.linalg.norm(S - Kd) <= 1e-12 * np.linalg.norm(Kd)
    # tangent = derivative of the residual (central differences along a random direction)
    d = rng.uniform(-1, 1, n)
    h = 1e-6
    Rp = _run(term, U + h * d, n)[3]
    Rm = _run(term, U - h * d, n)[3]
    fd = (Rp - Rm) / (2 * h)
    assert np.linalg.norm(S @ d - fd) <= 1e-7 * np.linalg.norm(fd)
    # a state change may move the pattern of a value-dependent term: the deferred check settles it, results stay right
    jc2, ir2, pr2, R2 = _run(term, 0.0 * U, n)
    K0, R0 = _numpy_forms(m, dfem.elem_dof(), t, 0.0 * U, c)
    S2 = sp.csc_matrix((pr2, ir2, jc2), shape=(n, n)).toarray()
    assert np.linalg.norm(S2 - K0) <= 1e-12 * np.linalg.norm(K0) and np.linalg.norm(R2 - R0) <= 1e-12 * max(np.linalg.norm(R0), 1e-300)


def test_jit_errors_are_loud():
    from getfem_b200 import capi
    ctx, m, dmesh, dfem, t, tab, U, rng = _setup("PK", 3, 1, 2, [1, 1, 1], 0.0)
    term = capi.DeviceTerm.jit(ctx, dmesh, dfem, tab, "nonsense(u)*tv", "0.0", [], 1.0)
    with pytest.raises(capi.GfgpuError, match="does not compile"):
        term.assemble_host(U, capi.TANGENT, None, None)


@pytest.mark.parametrize("mesh", [("PK", 3, 2, 4, [2, 2, 2], 0.15), ("PK", 2, 2, 4, [4, 3], 0.15), ("QK", 3, 1, 3, [2, 3, 2], 0.1)],
                         ids=lambda c: "%s%dd-k%d" % (c[0], c[1], c[2]))
def test_jit_vector_variable_reproduces_the_elasticity_family(mesh):
    """vector variable: u / tv are vec, gu / tg are mat; lambda div u div v + mu (grad u + grad u^T) : grad v is the ELASTICITY
    family: identical pattern, values and residual 1e-13"""
    import getfem_b200 as gf
    from getfem_b200 import capi, fem_tables
    gt, dim, k, im, nsub, distort = mesh
    ctx = capi.Context(0)
    m = gf.mesh()
    gf.regular_unit_mesh(m, nsub, "GT_%s(%d,1)" % (gt, dim))
    rng = np.random.default_rng(2)
    m.pts = m.pts + distort / max(nsub) * rng.uniform(-1, 1, m.pts.shape)
    m._dev = {}
    mf = gf.mesh_fem(m, dim)
    mf.set_classical_finite_element(k)
    dmesh, dfem = m.device(ctx), mf.device(ctx)
    t = fem_tables.classical_tables(gt, dim, k, im)
    tab = capi.DeviceTables(ctx, t["quad_w"], t["gt_grad"], t["phi"], t["gphi"])
    U = rng.uniform(-1, 1, dfem.ndof)
    ref = capi.DeviceTerm(ctx, dmesh, dfem, tab, "elast", [1.3, 0.7], 1.0, capi.STRATEGY_STAGED)
    jc, ir, pr, R = _run(ref, U, dfem.ndof)
    f1 = "par[0]*trace(gu)*trace(tg) + par[1]*ddot(gu+transp(gu),tg)"
    f2 = "par[0]*trace(t2g)*trace(tg) + par[1]*ddot(t2g+transp(t2g),tg)"
    jit = capi.DeviceTerm.jit(ctx, dmesh, dfem, tab, f1, f2, [1.3, 0.7], 1.0, value_dependent=False)
    jjc, jir, jpr, jR = _run(jit, U, dfem.ndof)
    assert np.array_equal(jc, jjc) and np.array_equal(ir, jir)
    assert np.linalg.norm(pr - jpr) <= 1e-13 * np.linalg.norm(pr) and np.linalg.norm(R - jR) <= 1e-13 * np.linalg.norm(R)
