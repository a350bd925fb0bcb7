"""ctypes binding of the plain-C assembly oracle (oracle/asm_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never imported by the
getfem_b200 package (tests/test_abi.py::test_product_never_imports_oracle enforces this).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgfo.so")
FAMILIES = {"laplace": 0, "elast": 1, "svk": 2, "nh_ciarlet": 3, "nh_bonet": 4, "mass": 5, "source": 6,
            "nsource": 7, "mooney_rivlin": 8, "ciarlet_geymonat": 9, "blatz_ko": 10}


def build(force=False):
    src = os.path.join(_HERE, "asm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-O2", "-fPIC", "-shared", "-o", _SO, src, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.gfo_assemble.restype = C.c_void_p
        L.gfo_assemble.argtypes = [C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.gfo_assemble_region.restype = C.c_void_p
        L.gfo_assemble_region.argtypes = L.gfo_assemble.argtypes + [C.c_int64] + [C.c_void_p] * 5
        L.gfo_assemble_fields.restype = C.c_void_p
        L.gfo_assemble_fields.argtypes = L.gfo_assemble_region.argtypes + [C.c_int, C.c_int] + [C.c_void_p] * 4
        L.gfo_nnz.restype = C.c_int64
        L.gfo_nnz.argtypes = [C.c_void_p]
        L.gfo_get_csc.argtypes = [C.c_void_p] * 4
        L.gfo_get_residual.argtypes = [C.c_void_p] * 2
        L.gfo_free.argtypes = [C.c_void_p]
        L.gfo_hyper_law.argtypes = [C.c_int] + [C.c_void_p] * 4
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def hyper_law(family, grad_u, params):
    """(S, dS): PK2 stress (3x3) and dS[i,j,k,l] = dS_ij / d(Grad_u)_kl of the law behind `family` at one point."""
    Gu = np.asfortranarray(grad_u, dtype=np.float64)  # the oracle's tensors are column-major
    par = np.ascontiguousarray(params, np.float64)
    S = np.zeros((3, 3), order="F")
    dS = np.zeros((3, 3, 3, 3), order="F")
    lib().gfo_hyper_law(FAMILIES[family], _p(Gu), _p(par), _p(S), _p(dS))
    return np.array(S), np.array(dS)


def potential(pts, conn, elem_dof, ndof, Q, w, gt_grad, phi, gphi, gt_linear, family, params, U, region=None, nq=None):
    """Order 0 (ga_workspace::assembly(0), assembled_potential()): the potential whose first variation is the family's
    order-1 form.  Quadratic forms: 1/2 u.R with R = K u; linear forms: u.R; finite-strain laws: int W (asm_oracle.c
    hyper_energy, accumulated by the element loop under order_mask bit 2)."""
    hyper = family in ("svk", "nh_ciarlet", "nh_bonet", "mooney_rivlin", "ciarlet_geymonat", "blatz_ko")
    _, _, _, R = assemble(pts, conn, elem_dof, ndof, Q, w, gt_grad, phi, gphi, gt_linear, family, params, U,
                          order_mask=5 if hyper else 1, region=region, nq=nq)
    if hyper:
        lib().gfo_last_potential.restype = C.c_double
        return float(lib().gfo_last_potential())
    d = float(np.dot(np.asarray(U, np.float64), R))
    return d if family in ("source", "nsource") else 0.5 * d


def assemble(pts, conn, elem_dof, ndof, Q, w, gt_grad, phi, gphi, gt_linear, family, params, U, order_mask=3,
             region=None, nq=None, fields=None):
    """Returns (jc, ir, pr, R): tangent in CSC (int64 indices) and residual.
    region = None (all convexes) or a dict: items_cv, items_f (face -1 = whole convex) and, when faces are used,
    face_first, face_nq, ref_normals -- the tables then cover ALL integration points and nq is the number of
    volume points.
    fields = None or a dict: d_elem_dof [ne, nd_d], d_phi [points, nd_d] (same point order as the other tables),
    vals = [values of field 0 (, field 1)]: fem-data coefficients replacing params[0 .. len(vals))."""
    pts = np.ascontiguousarray(pts, np.float64)
    conn = np.ascontiguousarray(conn, np.int32)
    elem_dof = np.ascontiguousarray(elem_dof, np.int64)
    w = np.ascontiguousarray(w, np.float64)
    gt_grad = np.ascontiguousarray(gt_grad, np.float64)
    phi = np.ascontiguousarray(phi, np.float64)
    gphi = np.ascontiguousarray(gphi, np.float64)
    params = np.ascontiguousarray(params, np.float64)
    U = np.ascontiguousarray(U, np.float64)
    ne, ng = conn.shape
    nd = elem_dof.shape[1]
    L = lib()
    if fields is not None:
        ded = np.ascontiguousarray(fields["d_elem_dof"], np.int64)
        dphi = np.ascontiguousarray(fields["d_phi"], np.float64)
        vals = [np.ascontiguousarray(v, np.float64) for v in fields["vals"]]
        assert dphi.shape[0] == len(w), "the data fem's table must cover the same points as the other tables"
        icv = ifc = ff = fn = rn = None
        if region is not None:
            icv = np.ascontiguousarray(region["items_cv"], np.int32)
            ifc = np.ascontiguousarray(region.get("items_f", np.full(len(icv), -1)), np.int32)
            if (ifc >= 0).any():
                ff = np.ascontiguousarray(region["face_first"], np.int32)
                fn = np.ascontiguousarray(region["face_nq"], np.int32)
                rn = np.ascontiguousarray(region["ref_normals"], np.float64)
        pp = lambda a: None if a is None else _p(a)  # noqa: E731
        h = L.gfo_assemble_fields(pts.shape[1], ne, ng, _p(pts), _p(conn), nd, Q, _p(elem_dof), ndof,
                                  len(w) if nq is None else nq, _p(w), _p(gt_grad), _p(phi), _p(gphi), int(gt_linear),
                                  FAMILIES[family], _p(params), _p(U), order_mask, ne if icv is None else len(icv), pp(icv),
                                  pp(ifc), pp(ff), pp(fn), pp(rn), len(vals), ded.shape[1], _p(ded), _p(dphi), _p(vals[0]),
                                  _p(vals[1]) if len(vals) > 1 else None)
    elif region is None and nq is not None:  # all-point tables were passed: the volume points come first
        w, gt_grad, phi, gphi = (np.ascontiguousarray(t[:nq]) for t in (w, gt_grad, phi, gphi))
    if fields is not None:
        pass
    elif region is None:
        h = L.gfo_assemble(pts.shape[1], ne, ng, _p(pts), _p(conn), nd, Q, _p(elem_dof), ndof, len(w), _p(w),
                           _p(gt_grad), _p(phi), _p(gphi), int(gt_linear), FAMILIES[family], _p(params), _p(U),
                           order_mask)
    else:
        icv = np.ascontiguousarray(region["items_cv"], np.int32)
        ifc = np.ascontiguousarray(region.get("items_f", np.full(len(icv), -1)), np.int32)
        faces = bool((ifc >= 0).any())
        ff = np.ascontiguousarray(region["face_first"], np.int32) if faces else None
        fn = np.ascontiguousarray(region["face_nq"], np.int32) if faces else None
        rn = np.ascontiguousarray(region["ref_normals"], np.float64) if faces else None
        h = L.gfo_assemble_region(pts.shape[1], ne, ng, _p(pts), _p(conn), nd, Q, _p(elem_dof), ndof,
                                  len(w) if nq is None else nq, _p(w), _p(gt_grad), _p(phi), _p(gphi),
                                  int(gt_linear), FAMILIES[family], _p(params), _p(U), order_mask, len(icv), _p(icv),
                                  _p(ifc), _p(ff) if faces else None, _p(fn) if faces else None,
                                  _p(rn) if faces else None)
    nnz = L.gfo_nnz(h)
    jc = np.empty(ndof + 1, np.int64)
    ir = np.empty(nnz, np.int64)
    pr = np.empty(nnz, np.float64)
    R = np.empty(ndof, np.float64)
    L.gfo_get_csc(h, _p(jc), _p(ir), _p(pr))
    L.gfo_get_residual(h, _p(R))
    L.gfo_free(h)
    return jc, ir, pr, R


def rect_div_pressure(pts, conn, edof_r, Qr, edof_c, quad_w, gt_grad, gphi_r, phi_c, coef=1.0):
    """TEST INFRASTRUCTURE: numpy restatement of a coupled order-2 tree "-(Test2_p*Div_Test_u)" up to its sign and factor:
    block(row (i,a), column j) = coef * int psi_j d(phi_i)/dx_a over every convex, element matrix by element matrix through
    the reference's drop rule (|v| > 1e-14 max|E| per element matrix, C&E.cc:4889,4898,5380-5402; an entry is stored iff one
    of its contributions is kept, sums in ascending convex order like add_elem_matrix, C&E.cc:4853-4936).
    Geometry as ga_exec does it: K = G^T pc(q), J = |det K|, B = K^-T (bgeot_geometric_trans.cc:270-413).
    Returns (jc, ir, pr) of the (ndof_r x ndof_c) block in gmm::csc_matrix layout."""
    pts = np.asarray(pts, float)
    ne, ng = conn.shape
    N = pts.shape[1]
    nq, ndr = gphi_r.shape[0], gphi_r.shape[1]
    ndc = phi_c.shape[1]
    nrows = int(edof_r.max()) + Qr
    ncols = int(edof_c.max()) + 1
    entries = {}
    for e in range(ne):
        G = pts[conn[e]]  # ng x N
        E = np.zeros((ndr * Qr, ndc))
        for q in range(nq):
            if quad_w[q] == 0.0:
                continue
            K = G.T @ gt_grad[q]  # N x N
            J = abs(np.linalg.det(K))
            B = np.linalg.inv(K).T
            dphi = gphi_r[q] @ B.T  # ndr x N : d phi_i / d x_a
            for i in range(ndr):
                for a in range(Qr):
                    E[i * Qr + a, :] += quad_w[q] * J * dphi[i, a] * phi_c[q]
        E *= coef
        vmax = np.abs(E).max()
        for i in range(ndr):
            for a in range(Qr):
                for j in range(ndc):
                    v = E[i * Qr + a, j]
                    if vmax != 0.0 and abs(v) > 1e-14 * vmax:
                        key = (int(edof_c[e, j]), int(edof_r[e, i]) + a)
                        entries[key] = entries.get(key, 0.0) + v
    keys = sorted(entries)
    jc = np.zeros(ncols + 1, np.int64)
    for c, _ in keys:
        jc[c + 1] += 1
    jc = np.cumsum(jc)
    ir = np.array([r for _, r in keys], np.int32)
    pr = np.array([entries[k] for k in keys])
    return jc, ir, pr
