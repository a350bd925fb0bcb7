"""Reference-element tables staged once on the device: geometric-transformation gradients, Lagrange
basis values/gradients and quadrature rules for the classical PK / QK families.

What it restates from the reference (paths relative to the GetFEM tree):
  PK_fem_   src/getfem_fem.cc:719-784   Lagrange simplex of degree k, nodes = simplex_of_reference(nc,k)
                                        in lexicographic lattice order, first coordinate fastest (:760-763)
  QK_fem_   src/getfem_fem.cc:791-843, 1010-1047   tensor product of equispaced FEM_PK(1,k);
                                        node index ix + (k+1) iy + (k+1)^2 iz
  GT_PK(N,1) / GT_QK(N,1)  src/bgeot_geometric_trans.cc:600-672  same shape functions with k = 1
  fem_precomp_::val/grad, geotrans_precomp_::grad   getfem_fem.h:653-680, bgeot_geometric_trans.h:292-345
  approx_integration   getfem_integration.h:155-222; Gauss rules getfem_integration.cc:560-741 (tensor rule,
                       point index i1 + i2*n1); classical_approx_im getfem_integration.cc:1245-1296
The simplex cubature points are written from their closed forms (Hammer-Stroud / Keast / Dunavant); they and
every basis table are compared with tables dumped from the reference (tests/golden) in tests/test_host_tables.py.
"""
import itertools
import math

import numpy as np


# ---------------------------------------------------------------- Lagrange bases
def pk_nodes(N, k):
    """Lattice multi-indices (a_1..a_N), sum <= k, lexicographic with the FIRST coordinate fastest."""
    out = []
    for a in itertools.product(range(k + 1), repeat=N):  # a[0] slowest in product(); reverse to make x fastest
        a = a[::-1]
        if sum(a) <= k:
            out.append(a)
    return np.array(out, np.int64)


def _pk_eval(N, k, X):
    """phi [nq, nd], gphi [nq, nd, N] of the degree-k Lagrange basis on the reference simplex."""
    X = np.asarray(X, np.float64)
    nq = X.shape[0]
    nodes = pk_nodes(N, k)
    nd = nodes.shape[0]
    lam = np.concatenate([1.0 - X.sum(1, keepdims=True), X], axis=1)  # [nq, N+1]
    dlam = np.concatenate([-np.ones((1, N)), np.eye(N)], axis=0)  # [N+1, N]
    phi = np.ones((nq, nd))
    gphi = np.zeros((nq, nd, N))
    for i, a in enumerate(nodes):
        full = (k - int(a.sum()),) + tuple(int(v) for v in a)
        # phi = prod_d prod_{m<a_d} (k lam_d - m)/(m+1)
        facs, dfacs = [], []  # each factor f and df/dlam_d
        for d, ad in enumerate(full):
            for m in range(ad):
                facs.append(((k * lam[:, d] - m) / (m + 1), d, k / (m + 1)))
        val = np.ones(nq)
        for f, _, _ in facs:
            val = val * f
        phi[:, i] = val
        g = np.zeros((nq, N))
        for t, (_, d, df) in enumerate(facs):
            rest = np.ones(nq)
            for s, (f2, _, _) in enumerate(facs):
                if s != t:
                    rest = rest * f2
            g += (rest * df)[:, None] * dlam[d][None, :]
        gphi[:, i, :] = g
    return phi, gphi


def _lagrange_1d(k, x):
    """values [nq, k+1] and derivatives of the equispaced 1-D Lagrange basis on [0,1]."""
    x = np.asarray(x, np.float64)
    t = np.arange(k + 1) / float(k)
    v = np.ones((x.shape[0], k + 1))
    dv = np.zeros((x.shape[0], k + 1))
    for i in range(k + 1):
        others = [j for j in range(k + 1) if j != i]
        den = np.prod([t[i] - t[j] for j in others])
        v[:, i] = np.prod([x - t[j] for j in others], axis=0) / den if others else 1.0
        s = np.zeros_like(x)
        for m in others:
            s = s + np.prod([x - t[j] for j in others if j != m], axis=0) if len(others) > 1 else s + 1.0
        dv[:, i] = s / den
    return v, dv


def _qk_eval(N, k, X):
    X = np.asarray(X, np.float64)
    nq = X.shape[0]
    n1 = k + 1
    nd = n1 ** N
    v1 = [_lagrange_1d(k, X[:, d]) for d in range(N)]
    phi = np.empty((nq, nd))
    gphi = np.empty((nq, nd, N))
    for i in range(nd):
        idx = [(i // n1 ** d) % n1 for d in range(N)]
        vals = [v1[d][0][:, idx[d]] for d in range(N)]
        ders = [v1[d][1][:, idx[d]] for d in range(N)]
        phi[:, i] = np.prod(vals, axis=0)
        for d in range(N):
            gphi[:, i, d] = np.prod([ders[e] if e == d else vals[e] for e in range(N)], axis=0)
    return phi, gphi


def lagrange_tables(kind, N, k, X):
    """kind: 'PK' or 'QK'.  Returns (phi, gphi) at the points X [nq, N]."""
    return _pk_eval(N, k, X) if kind == "PK" else _qk_eval(N, k, X)


def nb_dof(kind, N, k):
    return math.comb(N + k, k) if kind == "PK" else (k + 1) ** N


def ref_nodes(kind, N, k):
    if kind == "PK":
        return pk_nodes(N, k) / float(k)
    n1 = k + 1
    return np.array([[((i // n1 ** d) % n1) / float(k) for d in range(N)] for i in range(n1 ** N)])


# ---------------------------------------------------------------- quadrature
def gauss_1d(npts):
    """Gauss-Legendre on [0,1], ascending abscissae (IM_GAUSS1D(2*npts-2))."""
    x, w = np.polynomial.legendre.leggauss(npts)
    return 0.5 * (x + 1.0), 0.5 * w


def _perm3(a, b):  # (a,a,a),(b,a,a),(a,b,a),(a,a,b)
    return [(a, a, a), (b, a, a), (a, b, a), (a, a, b)]


def simplex_rule(N, degree):
    """Returns (name, X, w) of IM_TRIANGLE(k) / IM_TETRAHEDRON(k), k = first available >= degree
    (mim.set_integration_method(degree), getfem_integration.cc:1264-1269)."""
    s5, s15 = math.sqrt(5.0), math.sqrt(15.0)
    if N == 2:
        if degree <= 2:
            a, b = 1.0 / 6.0, 2.0 / 3.0
            return "IM_TRIANGLE(2)", np.array([(a, a), (b, a), (a, b)]), np.full(3, 1.0 / 6.0)
        if degree == 4:  # Dunavant degree 4, 6 points
            r = math.sqrt(38.0 - 44.0 * math.sqrt(0.4))
            a = (8.0 - math.sqrt(10.0) + r) / 18.0
            b = (8.0 - math.sqrt(10.0) - r) / 18.0
            q = math.sqrt(213125.0 - 53320.0 * math.sqrt(10.0))
            wa, wb = (620.0 + q) / 7440.0, (620.0 - q) / 7440.0
            X = [(a, a), (1 - 2 * a, a), (a, 1 - 2 * a), (b, b), (1 - 2 * b, b), (b, 1 - 2 * b)]
            return "IM_TRIANGLE(4)", np.array(X), np.array([wa] * 3 + [wb] * 3)
    if N == 3:
        if degree <= 2:
            a, b = (5.0 - s5) / 20.0, (5.0 + 3.0 * s5) / 20.0
            return "IM_TETRAHEDRON(2)", np.array(_perm3(a, b)), np.full(4, 1.0 / 24.0)
        if degree in (4, 5):  # Keast, 15 points, exact to degree 5
            a2, b2 = (7.0 + s15) / 34.0, (13.0 - 3.0 * s15) / 34.0
            a1, b1 = (7.0 - s15) / 34.0, (13.0 + 3.0 * s15) / 34.0
            c, d = (10.0 - 2.0 * s15) / 40.0, (10.0 + 2.0 * s15) / 40.0
            X = [(0.25, 0.25, 0.25)] + _perm3(a2, b2) + _perm3(a1, b1) + \
                [(d, c, c), (c, d, c), (d, d, c), (c, c, d), (d, c, d), (c, d, d)]
            w = [16.0 / 810.0] + [(2665.0 - 14.0 * s15) / 226800.0] * 4 + [(2665.0 + 14.0 * s15) / 226800.0] * 4 + \
                [10.0 / 1134.0] * 6
            return "IM_TETRAHEDRON(5)", np.array(X), np.array(w)
    raise NotImplementedError("no simplex cubature for dimension %d degree %d" % (N, degree))


def parallelepiped_rule(N, degree):
    """IM_GAUSS_PARALLELEPIPED(N, degree): degree/2+1 Gauss points per direction, first direction fastest."""
    n1 = degree // 2 + 1
    x, w = gauss_1d(n1)
    nq = n1 ** N
    X = np.empty((nq, N))
    W = np.ones(nq)
    for q in range(nq):
        for d in range(N):
            i = (q // n1 ** d) % n1
            X[q, d] = x[i]
            W[q] *= w[i]
    return "IM_GAUSS_PARALLELEPIPED(%d,%d)" % (N, degree), X, W


# ---------------------------------------------------------------- faces of the reference elements
def face_dir_points(gt_kind, N):
    """cvr->dir_points_of_face(f) of the reference simplex / parallelepiped (bgeot_convex_ref.cc): [nf, N, N] (N points
    of dimension N per face) and the reference normals pgt->normals() [nf, N].
    Simplex: face 0 is the oblique face (points e_1..e_N, unit normal (1..1)/sqrt(N)), face f >= 1 is x_{f-1} = 0
    (origin, then the e_j, j != f-1; normal -e_{f-1}).  Parallelepiped: face 2d is x_d = 1, face 2d+1 is x_d = 0
    (base point, then base + e_j, j != d; normals +-e_d)."""
    I = np.eye(N)
    pts, nrm = [], []
    if gt_kind == "PK":
        pts.append(I.copy())
        nrm.append(np.full(N, 1.0 / math.sqrt(N)))
        for f in range(N):
            pts.append(np.array([np.zeros(N)] + [I[j] for j in range(N) if j != f]))
            nrm.append(-I[f])
    else:
        for d in range(N):
            for side in (1.0, 0.0):
                base = side * I[d]
                pts.append(np.array([base] + [base + I[j] for j in range(N) if j != d]))
                nrm.append(I[d] if side else -I[d])
    return np.array(pts), np.array(nrm)


def _face_rule(gt_kind, N, im_name):
    """The (N-1)-dimensional method the reference puts on every face (getfem_im_list.h im_desc_face_meth;
    IM_PRODUCT of Gauss rules for parallelepipeds, getfem_integration.cc:620-647)."""
    if gt_kind == "QK":
        deg = int(im_name.split(",")[1].rstrip(")"))
        if N == 2:
            x, w = gauss_1d(deg // 2 + 1)
            return x[:, None], w
        _, X, w = parallelepiped_rule(N - 1, deg)
        return X, w
    if im_name == "IM_TRIANGLE(2)":
        x, w = gauss_1d(2)
        return x[:, None], w
    if im_name == "IM_TRIANGLE(4)":
        x, w = gauss_1d(3)
        return x[:, None], w
    if im_name == "IM_TETRAHEDRON(2)":
        _, X, w = simplex_rule(2, 2)
        return X, w
    if im_name == "IM_TETRAHEDRON(5)":  # IM_TRIANGLE(5): Radon's 7 points
        s15 = math.sqrt(15.0)
        a, b = (6.0 + s15) / 21.0, (9.0 - 2.0 * s15) / 21.0  # 0.4701.., 0.0597..
        c, d = (6.0 - s15) / 21.0, (9.0 + 2.0 * s15) / 21.0  # 0.1012.., 0.7974..
        X = [(1.0 / 3.0, 1.0 / 3.0), (a, a), (b, a), (a, b), (c, c), (d, c), (c, d)]
        w = [9.0 / 80.0] + [(155.0 + s15) / 2400.0] * 3 + [(155.0 - s15) / 2400.0] * 3
        return np.array(X), np.array(w)
    raise NotImplementedError("no face method for " + im_name)


def face_rule_points(gt_kind, N, im_name):
    """Points [nf, nqf, N] and weights [nf, nqf] of the face part of the approx_integration, as
    approx_integration::add_method_on_face builds them (getfem_integration.cc:320-351): pt = P0 + sum_j (P_{j+1} - P0)
    xi_j, weight = w * sqrt(det(A^T A))."""
    P, nrm = face_dir_points(gt_kind, N)
    xi, w = _face_rule(gt_kind, N, im_name)
    X = np.empty((P.shape[0], xi.shape[0], N))
    W = np.empty((P.shape[0], xi.shape[0]))
    for f in range(P.shape[0]):
        A = (P[f, 1:] - P[f, 0]).T  # N x (N-1)
        det = math.sqrt(abs(np.linalg.det(A.T @ A)))
        X[f] = P[f, 0][None, :] + xi @ A.T
        W[f] = w * det
    return X, W, nrm


def classical_face_tables(gt_kind, N, fem_degree, im_degree):
    """Face tables for gfgpu_tables_set_faces: normals [nf, N], w [nf, nqf], gt_grad [nf, nqf, ng, N], phi, gphi."""
    name = (simplex_rule(N, im_degree) if gt_kind == "PK" else parallelepiped_rule(N, im_degree))[0]
    X, W, nrm = face_rule_points(gt_kind, N, name)
    nf, nqf = W.shape
    flat = X.reshape(nf * nqf, N)
    _, gt_grad = lagrange_tables(gt_kind, N, 1, flat)
    phi, gphi = lagrange_tables(gt_kind, N, fem_degree, flat)
    return {"normals": nrm, "quad_x": X, "quad_w": W, "gt_grad": gt_grad.reshape(nf, nqf, -1, N),
            "phi": phi.reshape(nf, nqf, -1), "gphi": gphi.reshape(nf, nqf, -1, N)}


def classical_tables(gt_kind, N, fem_degree, im_degree):
    """All tables one (geotrans, fem, im) triple needs.  gt_kind 'PK' (simplices, affine) or 'QK'."""
    name, X, w = simplex_rule(N, im_degree) if gt_kind == "PK" else parallelepiped_rule(N, im_degree)
    _, gt_grad = lagrange_tables(gt_kind, N, 1, X)
    phi, gphi = lagrange_tables(gt_kind, N, fem_degree, X)
    return {"im": name, "quad_x": X, "quad_w": w, "gt_grad": gt_grad, "phi": phi, "gphi": gphi}
