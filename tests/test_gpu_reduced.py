"""GPU: reduced mesh_fems behind the C ABI (gfgpu_reduction_*, gfgpu_matrix_add_term_reduced / add_rect_reduced) and the coupled
mass family on regions (GFGPU_RECT_MASS, gfgpu_rect_set_region).  The reference projects what it assembled on the basic dofs
with the extension matrix E: K += E1^T K_basic E2, V += E1^T V_basic, U_basic = E U (workspace.cc:861-935).  Here the device
result is compared with scipy's sparse products of the SAME device blocks (the blocks themselves are pinned on the reference in
test_gpu_golden.py / test_gpu_rect.py; the whole chain against the reference in one process in test_gpu_dropin.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _extension(rng, nb, kind):
    import scipy.sparse as sp
    if kind == "selection":  # partial_mesh_fem: every reduced dof IS one basic dof
        kept = np.sort(rng.choice(nb, size=nb - nb // 3, replace=False))
        E = sp.csr_matrix((np.ones(len(kept)), (kept, np.arange(len(kept)))), shape=(nb, len(kept)))
    else:  # general: identity on the first reduced dofs, the last quarter of the basic dofs are combinations of two or three reduced ones
        nr = nb - nb // 4
        rows, cols, vals = list(range(nr)), list(range(nr)), [1.0] * nr
        for j in range(nr, nb):
            for c in sorted(set(rng.integers(0, nr, size=3).tolist())):
                rows.append(j); cols.append(c); vals.append(float(rng.uniform(-1, 1)))
        E = sp.csr_matrix((vals, (rows, cols)), shape=(nb, nr))
    E.sort_indices()
    return E


@pytest.mark.parametrize("kind", ["selection", "general"])
def test_term_projected_with_the_extension_matrix(kind):
    import scipy.sparse as sp
    from conftest import load_golden
    from getfem_b200 import capi
    g = load_golden("c3_elast3d_p2_n2")
    nb = g["meta"]["ndof"]
    ctx = capi.Context(0)
    mesh = capi.DeviceMesh(ctx, g["pts"], g["conn"], capi.GT_PK)
    fem = capi.DeviceFem(ctx, mesh, capi.FEM_PK, 2, 3, g["elem_dof"].shape[1], g["elem_dof"], nb)
    tab = capi.DeviceTables(ctx, g["quad_w"], g["gt_grad"], g["phi"], g["gphi"])
    term = capi.DeviceTerm(ctx, mesh, fem, tab, "elast", g["fparams"])
    rng = np.random.default_rng(11)
    E = _extension(rng, nb, kind)
    nr = E.shape[1]
    dE = capi.DeviceReduction(ctx, nb, nr, E.indptr, E.indices, E.data)
    # state and residual: U_basic = E U, V = E^T R_basic
    U = rng.uniform(-1, 1, nr)
    Ub = dE.extend(U)
    assert np.linalg.norm(Ub - E @ U) <= 1e-14 * np.linalg.norm(Ub)
    R = np.empty(nb)
    term.assemble_host(Ub, capi.TANGENT | capi.RESIDUAL, None, R)
    V0 = rng.uniform(-1, 1, nr)
    V = dE.restrict_add(R, V0, alpha=0.5)
    assert np.linalg.norm(V - (V0 + 0.5 * (E.T @ R))) <= 1e-13 * np.linalg.norm(V)
    # tangent: K(off.., off..) += alpha E^T K_basic E, twice (the second call finds its pattern in place)
    jc, ir, pr = term.export_csc()
    Kb = sp.csc_matrix((pr, ir, jc), shape=(nb, nb))
    # y = K^T x on the term's own CSC (gfgpu_term_tmult_dev: what bench.py's full-size property checks are made of)
    import torch
    xt = torch.from_numpy(rng.uniform(-1, 1, nb)).cuda()
    yt = torch.full((nb,), 3.0, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    term.tmult_dev(xt.data_ptr(), yt.data_ptr(), alpha=2.0, beta=0.5)
    term.ctx_synchronize()
    want_y = 1.5 + 2.0 * (Kb.T @ xt.cpu().numpy())
    assert np.linalg.norm(yt.cpu().numpy() - want_y) <= 1e-13 * np.linalg.norm(want_y)
    off, alpha = 5, 1.0  # (v + v = 2 v exactly: the selection case compares bit for bit)
    K = capi.DeviceMatrix(ctx, nr + off + 3)
    K.add_term_reduced(term, dE, alpha=alpha, row_off=off, col_off=off)
    gen = K.pattern_generation
    K.add_term_reduced(term, dE, alpha=alpha, row_off=off, col_off=off)
    assert K.pattern_generation == gen
    kjc, kir, kpr = K.export_csc()
    got = sp.csc_matrix((kpr, kir, kjc), shape=(nr + off + 3, nr + off + 3))
    want = (2 * alpha) * (E.T @ Kb @ E)
    want = sp.csc_matrix(want)
    want.eliminate_zeros()
    sub = sp.csc_matrix(got[off:off + nr, off:off + nr])
    assert got.nnz == sub.nnz, "entries outside the block"
    d = sub - want
    assert abs(d).max() <= 1e-12 * abs(want).max()
    if kind == "selection":  # no sums: the stored entries are exactly the nonzero entries of the selected block
        sub.sort_indices(); want.sort_indices()
        assert np.array_equal(sub.indptr, want.indptr) and np.array_equal(sub.indices, want.indices)
        assert np.array_equal(sub.data, want.data)


def _mass_setup(gt, dim, ks, im, nsub, qdim):
    import getfem_b200 as gf
    from getfem_b200 import capi, fem_tables
    ctx = capi.Context(0)
    m = gf.mesh()
    gf.regular_unit_mesh(m, nsub, "GT_%s(%d,1)" % (gt, dim))
    rng = np.random.default_rng(3)
    m.pts = m.pts + 0.1 / max(nsub) * rng.uniform(-1, 1, m.pts.shape)
    m._dev = {}
    mfa, mfb = gf.mesh_fem(m, qdim), gf.mesh_fem(m, qdim)
    mfa.set_classical_finite_element(ks[0])
    mfb.set_classical_finite_element(ks[1])
    ta = fem_tables.classical_tables(gt, dim, ks[0], im)
    tb = fem_tables.classical_tables(gt, dim, ks[1], im)
    taba = capi.DeviceTables(ctx, ta["quad_w"], ta["gt_grad"], ta["phi"], ta["gphi"])
    tabb = capi.DeviceTables(ctx, tb["quad_w"], tb["gt_grad"], tb["phi"], tb["gphi"])
    return ctx, m, m.device(ctx), mfa.device(ctx), mfb.device(ctx), ta, tb, taba, tabb, rng


@pytest.mark.parametrize("case", [("PK", 3, (1, 2), 4, [2, 2, 2], 3), ("QK", 2, (1, 2), 6, [3, 3], 1), ("PK", 2, (2, 2), 4, [3, 4], 2)],
                         ids=lambda c: "%s%dd-k%d%d-q%d" % (c[0], c[1], c[2][0], c[2][1], c[5]))
def test_coupled_mass_block_on_convexes(case):
    """block((i,a),(j,b)) = delta_ab int phi_i psi_j over all convexes and over a subset of them, against a numpy quadrature
    (affine / multilinear geometry from the same tables); off-diagonal components are exact zeros and fall to the drop rule."""
    import scipy.sparse as sp
    from getfem_b200 import capi
    gt, dim, ks, im, nsub, qdim = case
    ctx, m, dmesh, dfa, dfb, ta, tb, taba, tabb, rng = _mass_setup(gt, dim, ks, im, nsub, qdim)
    r = capi.DeviceRect(ctx, dmesh, dfa, taba, dfb, tabb, capi.RECT_MASS, 1.0, 1.0)
    eda, edb = dfa.elem_dof(), dfb.elem_dof()
    w, gtg = ta["quad_w"], ta["gt_grad"]

    def dense(cvs):
        M = np.zeros((dfa.ndof, dfb.ndof))
        for e in cvs:
            G = m.pts[m.conn[e]].T  # dim x ng
            for q in range(len(w)):
                if w[q] == 0.0:
                    continue
                J = abs(np.linalg.det(G @ gtg[q]))
                blk = w[q] * J * np.outer(ta["phi"][q], tb["phi"][q])
                for c in range(qdim):
                    M[np.ix_(eda[e] + c, edb[e] + c)] += blk
        return M

    ne = m.conn.shape[0]
    for cvs in (list(range(ne)), sorted(rng.choice(ne, size=ne // 2, replace=False).tolist())):
        r.set_region(None if len(cvs) == ne else cvs)
        r.assemble()
        jc, ir, pr = r.export_csc()
        got = sp.csc_matrix((pr, ir, jc), shape=(dfa.ndof, dfb.ndof)).toarray()
        want = dense(cvs)
        assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()
        # stored entries: those with a contribution the drop rule keeps -- never a (c != d) component pair
        rows = np.asarray(ir)
        cols = np.repeat(np.arange(dfb.ndof), np.diff(jc))
        if qdim > 1:
            assert np.all(rows % qdim == cols % qdim)
