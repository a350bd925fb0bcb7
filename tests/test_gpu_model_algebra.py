"""GPU: model-level algebra on the resident tangent: Dirichlet conditions with simplification,
the right-hand side accumulated on the device, the CSR hand-off and a device solve -- assemble -> constrain -> solve
without a copy of K to the host.  The reference semantics (model::assembly, getfem_models.cc:2806-2871, restated here
with scipy on the exported matrix) is the checker."""
import numpy as np
import pytest

from conftest import load_golden
from test_gpu_golden import device_assemble

pytestmark = pytest.mark.gpu


def reference_constraints(S, rhs, dofs, go, pr, linear, symmetric):
    """getfem_models.cc:2834-2869, on a scipy matrix"""
    import scipy.sparse as sp
    S = S.tolil(copy=True)
    rhs = rhs.copy()
    if linear:
        if symmetric and np.linalg.norm(go) > 0:
            rhs -= S.tocsc()[:, dofs] @ go
        rhs[dofs] = go
    else:
        rhs[dofs] += go - pr
    S[dofs, :] = 0.0
    if symmetric:
        S[:, dofs] = 0.0
    for d in dofs:
        S[d, d] = 1.0
    return sp.csc_matrix(S), rhs


@pytest.mark.parametrize("name,linear,symmetric", [("c3_elast3d_p2_n2", True, True), ("c3_elast3d_p2_n2", True, False),
                                                   ("c4_nh_ciarlet_q2_n2", False, True), ("m_lap_q2_robin", True, True)])
def test_dirichlet_with_simplification_on_the_device(name, linear, symmetric):
    import scipy.sparse as sp
    import torch
    from getfem_b200 import capi
    g = load_golden(name)
    n = g["meta"]["ndof"]
    jc, ir, pr, R, term = device_assemble(g)
    K = getattr(term, "matrix", None)
    if K is None:
        K = capi.DeviceMatrix(term.ctx, n)
        K.add_term(term)
    jc, ir, pr = K.export_csc()
    S = sp.csc_matrix((pr, ir, jc), shape=(n, n))
    rng = np.random.default_rng(11)
    dofs = np.sort(rng.choice(n, size=max(3, n // 7), replace=False)).astype(np.int64)
    go = rng.uniform(-1, 1, dofs.size)
    present = rng.uniform(-1, 1, dofs.size)
    rhs0 = rng.uniform(-1, 1, n)
    rhs = torch.from_numpy(rhs0).cuda()
    gen = K.pattern_generation
    K.apply_dof_constraints(dofs, go, rhs.data_ptr(), present=present, linear=linear, symmetric=symmetric)
    torch.cuda.synchronize()
    Sref, rref = reference_constraints(S, rhs0, dofs, go, present, linear, symmetric)
    assert K.pattern_generation == gen  # the diagonal slots exist in a finite-element tangent: the pattern does not move
    jc2, ir2, pr2 = K.export_csc()
    assert np.array_equal(jc2, jc) and np.array_equal(ir2, ir)
    S2 = sp.csc_matrix((pr2, ir2, jc2), shape=(n, n))
    D = (S2 - Sref).tocoo()
    assert D.nnz == 0 or np.abs(D.data).max() == 0.0  # cleared entries are exact zeros, the rest is untouched
    assert np.linalg.norm(rhs.cpu().numpy() - rref) <= 1e-13 * np.linalg.norm(rref)


def test_rhs_only_is_refused_for_symmetric_linear_models():
    import torch
    from getfem_b200 import capi
    g = load_golden("c3_elast3d_p2_n2")
    n = g["meta"]["ndof"]
    jc, ir, pr, R, term = device_assemble(g)
    K = capi.DeviceMatrix(term.ctx, n)
    K.add_term(term)
    rhs = torch.zeros(n, dtype=torch.float64, device="cuda")
    with pytest.raises(capi.GfgpuError, match="Rhs only"):  # getfem_models.cc:2843-2845
        K.apply_dof_constraints([0, 1], [1.0, 2.0], rhs.data_ptr(), build_matrix=False)
    K.apply_dof_constraints([0, 1], [0.0, 0.0], rhs.data_ptr(), build_matrix=False)  # homogeneous values: allowed
    with pytest.raises(capi.GfgpuError, match="out of range"):
        K.apply_dof_constraints([n], [1.0], rhs.data_ptr())


def test_missing_diagonal_slot_is_added_to_the_pattern():
    """a rectangular-looking coupling block has no diagonal entries: the constrained dofs get theirs"""
    from getfem_b200 import capi
    g = load_golden("c3_elast3d_p2_n2")
    n = g["meta"]["ndof"]
    jc, ir, pr, R, term = device_assemble(g)
    B = capi.DeviceMatrix(term.ctx, 2 * n)
    B.add_term(term, 1.0, 0, n)  # upper right block only
    gen, nnz = B.pattern_generation, B.nnz
    B.apply_dof_constraints([3, n + 5], [0.0, 0.0], None, symmetric=False)
    assert B.pattern_generation == gen + 1 and B.nnz == nnz + 2
    bjc, bir, bpr = B.export_csc()
    for d in (3, n + 5):
        col = slice(bjc[d], bjc[d + 1])
        k = np.searchsorted(bir[col], d)
        assert bir[col][k] == d and bpr[col][k] == 1.0
        assert np.count_nonzero(bpr[col]) == 1 or d != 3  # column 3 of the upper right block is empty otherwise


@pytest.mark.parametrize("name", ["c3_elast3d_p2_n2", "m_lap_q2_robin", "c3_elast3d_p2_n8"])
def test_csr_hand_off_and_device_solve(name):
    """K u = f with clamped dofs, entirely on the device: residual accumulated with gfgpu_term_residual_add_dev,
    constraints applied, CG; the CSR export equals scipy's conversion bit for bit; the solution equals a direct host solve."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    import torch
    from getfem_b200 import capi
    g = load_golden(name)
    n = g["meta"]["ndof"]
    jc, ir, pr, R, term = device_assemble(g)
    K = getattr(term, "matrix", None)
    if K is None:
        K = capi.DeviceMatrix(term.ctx, n)
        K.add_term(term)
    rng = np.random.default_rng(3)
    f = rng.uniform(-1, 1, n)
    rhs = torch.from_numpy(f).cuda()
    if not g["extra_terms"]:
        term.residual_add_dev(rhs.data_ptr(), alpha=-1.0)  # rrhs = f - R(u0), as a Newton step would start
        f = f - R
    dofs = np.arange(0, n, 5, dtype=np.int64)  # enough clamped dofs to remove the rigid motions / constants
    go = np.zeros(dofs.size)
    K.apply_dof_constraints(dofs, go, rhs.data_ptr())
    jc2, ir2, pr2 = K.export_csc()
    S = sp.csc_matrix((pr2, ir2, jc2), shape=(n, n))
    # CSR hand-off
    rp = torch.empty(n + 1, dtype=torch.int64, device="cuda")
    col = torch.empty(K.nnz, dtype=torch.int32, device="cuda")
    val = torch.empty(K.nnz, dtype=torch.float64, device="cuda")
    K.export_csr_dev(rp.data_ptr(), col.data_ptr(), val.data_ptr())
    torch.cuda.synchronize()
    Sr = sp.csr_matrix(S)
    Sr.sort_indices()
    assert np.array_equal(rp.cpu().numpy(), Sr.indptr) and np.array_equal(col.cpu().numpy(), Sr.indices)
    assert np.array_equal(val.cpu().numpy(), Sr.data)
    # torch consumes the hand-off without a copy: K x through torch's own CSR product
    Kt = torch.sparse_csr_tensor(rp, col.to(torch.int64), val, size=(n, n))
    xt = torch.from_numpy(rng.uniform(-1, 1, n)).cuda()
    assert np.linalg.norm((Kt @ xt).cpu().numpy() - S @ xt.cpu().numpy()) <= 1e-12 * np.linalg.norm(S @ xt.cpu().numpy())
    # device solve
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    its, rel = K.cg_dev(rhs.data_ptr(), x.data_ptr(), rtol=1e-12, max_iter=20000)
    b = rhs.cpu().numpy()
    bref = f.copy()
    bref[dofs] = 0.0
    assert np.linalg.norm(b - bref) <= 1e-13 * np.linalg.norm(bref)
    xh = spl.spsolve(S.tocsc(), b)
    assert rel <= 1e-12 and its > 0
    assert np.linalg.norm(x.cpu().numpy() - xh) <= 1e-8 * np.linalg.norm(xh)
    assert np.abs(x.cpu().numpy()[dofs]).max() == 0.0
    x2 = torch.zeros(n, dtype=torch.float64, device="cuda")
    K.cg_dev(rhs.data_ptr(), x2.data_ptr(), rtol=1e-12, max_iter=20000)
    assert torch.equal(x, x2)  # fixed-order reductions: bitwise reproducible
