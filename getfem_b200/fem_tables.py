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


def classical_tables(gt_kind, N, fem_degree, im_degree):
    """All tables one (geotrans, fem, im) triple needs.  gt_kind 'PK' (simplices, affine) or 'QK'."""
    name, X, w = simplex_rule(N, im_degree) if gt_kind == "PK" else parallelepiped_rule(N, im_degree)
    _, gt_grad = lagrange_tables(gt_kind, N, 1, X)
    phi, gphi = lagrange_tables(gt_kind, N, fem_degree, X)
    return {"im": name, "quad_x": X, "quad_w": w, "gt_grad": gt_grad, "phi": phi, "gphi": gphi}
