"""ctypes binding of the C ABI declared in include/gfgpu.h (getfem_b200/libgfgpu.so).

This is the only way the Python host side reaches the device: there is no CPU fallback, and a
missing library or a missing GPU raises immediately.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GFGPU_LIB") or os.path.join(_HERE, "libgfgpu.so")  # GFGPU_LIB: kernel-variant experiments

GT_PK, GT_QK = 0, 1
FEM_PK, FEM_QK = 0, 1
LAPLACE, ELASTICITY, SVK, NEOHOOKEAN_CIARLET, NEOHOOKEAN_BONET, MASS, SOURCE, NORMAL_SOURCE = range(8)
MOONEY_RIVLIN, CIARLET_GEYMONAT, BLATZ_KO = 8, 9, 10
RESIDUAL, TANGENT = 1, 2
STRATEGY_AUTO, STRATEGY_STAGED, STRATEGY_RECOMPUTE = 0, 1, 2

FAMILY_BY_NAME = {
    "laplace": LAPLACE, "elast": ELASTICITY, "elasticity": ELASTICITY, "svk": SVK,
    "nh_ciarlet": NEOHOOKEAN_CIARLET, "nh_bonet": NEOHOOKEAN_BONET, "mass": MASS, "source": SOURCE,
    "nsource": NORMAL_SOURCE, "normal_source": NORMAL_SOURCE,
    "mooney_rivlin": MOONEY_RIVLIN, "ciarlet_geymonat": CIARLET_GEYMONAT, "blatz_ko": BLATZ_KO,
}

# symbol -> (restype, argtypes); kept in one table so tests can check it against include/gfgpu.h
_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_i64 = C.c_int64
SIGNATURES = {
    "gfgpu_last_error": (C.c_char_p, []),
    "gfgpu_version": (C.c_int, []),
    "gfgpu_launch_count": (_i64, []),
    "gfgpu_ctx_create": (C.c_int, [C.c_int, _P, _PP]),
    "gfgpu_ctx_destroy": (C.c_int, [_P]),
    "gfgpu_ctx_synchronize": (C.c_int, [_P]),
    "gfgpu_ctx_bytes_in_use": (_i64, [_P]),
    "gfgpu_ctx_measure_fp64_peak": (C.c_int, [_P, _P]),
    "gfgpu_ctx_measure_dmma_peak": (C.c_int, [_P, _P]),
    "gfgpu_mesh_create": (C.c_int, [_P, C.c_int, _i64, _P, _i64, C.c_int, _P, C.c_int, _PP]),
    "gfgpu_mesh_destroy": (C.c_int, [_P]),
    "gfgpu_fem_create": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _i64, _PP]),
    "gfgpu_fem_nb_dof": (_i64, [_P]),
    "gfgpu_fem_get_elem_dof": (C.c_int, [_P, _P]),
    "gfgpu_fem_destroy": (C.c_int, [_P]),
    "gfgpu_tables_create": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _PP]),
    "gfgpu_tables_set_gt_values": (C.c_int, [_P, _P, _P]),
    "gfgpu_tables_set_faces": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "gfgpu_tables_destroy": (C.c_int, [_P]),
    "gfgpu_term_create": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, C.c_int, C.c_double, C.c_int, _PP]),
    "gfgpu_term_create_jit": (C.c_int, [_P, _P, _P, _P, C.c_char_p, C.c_char_p, _P, C.c_int, C.c_double, C.c_int, _PP]),
    "gfgpu_jit_check": (C.c_int, [C.c_int, C.c_int, C.c_char_p, C.c_char_p]),
    "gfgpu_term_set_params": (C.c_int, [_P, _P, C.c_int]),
    "gfgpu_term_destroy": (C.c_int, [_P]),
    "gfgpu_term_set_region": (C.c_int, [_P, _i64, _P, _P]),
    "gfgpu_term_set_fields": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P]),
    "gfgpu_term_update_field": (C.c_int, [_P, C.c_int, _P]),
    "gfgpu_term_set_element_range": (C.c_int, [_P, _i64, _i64]),
    "gfgpu_term_assemble_dev": (C.c_int, [_P, _P, C.c_int]),
    "gfgpu_term_assemble_host": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "gfgpu_term_potential_dev": (C.c_int, [_P, _P, _P]),
    "gfgpu_term_potential_host": (C.c_int, [_P, _P, _P]),
    "gfgpu_term_last_timings": (C.c_int, [_P, _P]),
    "gfgpu_term_strategy": (C.c_int, [_P]),
    "gfgpu_term_kernel_kind": (C.c_int, [_P]),
    "gfgpu_term_nnz": (_i64, [_P]),
    "gfgpu_term_nb_dof": (_i64, [_P]),
    "gfgpu_term_pattern_generation": (_i64, [_P]),
    "gfgpu_term_csc_view": (C.c_int, [_P, _PP, _PP, _PP]),
    "gfgpu_term_residual_view": (C.c_int, [_P, _PP]),
    "gfgpu_term_export_csc_host": (C.c_int, [_P, _P, _P, _P]),
    "gfgpu_term_export_residual_host": (C.c_int, [_P, _P]),
    "gfgpu_matrix_create": (C.c_int, [_P, _i64, _i64, _PP]),
    "gfgpu_matrix_destroy": (C.c_int, [_P]),
    "gfgpu_matrix_clear": (C.c_int, [_P, C.c_int]),
    "gfgpu_matrix_add_term": (C.c_int, [_P, _P, C.c_double, _i64, _i64]),
    "gfgpu_matrix_nnz": (_i64, [_P]),
    "gfgpu_matrix_pattern_generation": (_i64, [_P]),
    "gfgpu_matrix_csc_view": (C.c_int, [_P, _PP, _PP, _PP]),
    "gfgpu_matrix_export_csc_host": (C.c_int, [_P, _P, _P, _P]),
    "gfgpu_matrix_mult_dev": (C.c_int, [_P, C.c_int, C.c_double, _P, C.c_double, _P]),
    "gfgpu_matrix_mult_host": (C.c_int, [_P, C.c_int, C.c_double, _P, C.c_double, _P]),
    "gfgpu_matrix_apply_dof_constraints": (C.c_int, [_P, _i64, _P, _P, _P, _P, C.c_int]),
    "gfgpu_matrix_export_csr_dev": (C.c_int, [_P, _P, _P, _P]),
    "gfgpu_matrix_cg_dev": (C.c_int, [_P, _P, _P, C.c_double, C.c_int, _P, _P]),
    "gfgpu_term_residual_add_dev": (C.c_int, [_P, C.c_double, _P, _i64]),
    "gfgpu_rect_create": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_double, C.c_double, _PP]),
    "gfgpu_rect_destroy": (C.c_int, [_P]),
    "gfgpu_rect_set_region": (C.c_int, [_P, C.c_int64, _P, _P]),
    "gfgpu_term_tmult_dev": (C.c_int, [_P, C.c_double, _P, C.c_double, _P]),
    "gfgpu_term_set_jit_potential": (C.c_int, [_P, C.c_char_p]),
    "gfgpu_reduction_create": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, _PP]),
    "gfgpu_reduction_destroy": (C.c_int, [_P]),
    "gfgpu_reduction_extend_host": (C.c_int, [_P, _P, _P]),
    "gfgpu_reduction_restrict_add_host": (C.c_int, [_P, C.c_double, _P, _P]),
    "gfgpu_matrix_add_term_reduced": (C.c_int, [_P, _P, _P, C.c_double, C.c_int64, C.c_int64]),
    "gfgpu_matrix_add_rect_reduced": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_double, C.c_int64, C.c_int64]),
    "gfgpu_rect_assemble_dev": (C.c_int, [_P]),
    "gfgpu_rect_nnz": (_i64, [_P]),
    "gfgpu_rect_export_csc_host": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "gfgpu_rect_mult_dev": (C.c_int, [_P, C.c_int, C.c_double, _P, C.c_double, _P]),
    "gfgpu_rect_mult_host": (C.c_int, [_P, C.c_int, C.c_double, _P, C.c_double, _P]),
    "gfgpu_matrix_add_rect": (C.c_int, [_P, _P, C.c_int, C.c_double, _i64, _i64]),
    "gfgpu_term_halo_begin": (C.c_int, [_P, _P, _P, _P]),
    "gfgpu_term_halo_ghost_pairs": (C.c_int, [_P, _i64, _i64, _P, _P, _P, _P]),
    "gfgpu_term_halo_add_source": (C.c_int, [_P, C.c_int, _i64, _P, _P, _P, _i64, _i64]),
    "gfgpu_term_halo_commit": (C.c_int, [_P, _i64, _i64]),
    "gfgpu_term_halo_send_view": (C.c_int, [_P, _i64, _i64, _PP, _P, _PP]),
    "gfgpu_term_halo_recv_view": (C.c_int, [_P, C.c_int, _PP, _P, _PP, _P]),
    "gfgpu_term_halo_accumulate": (C.c_int, [_P, C.c_int]),
    "gfgpu_term_halo_add_send": (C.c_int, [_P, C.c_int, _i64, _i64, _i64]),
    "gfgpu_term_halo_exchange": (C.c_int, [_P, _P, C.c_int]),
    "gfgpu_comm_unique_id": (C.c_int, [_P, C.c_int]),
    "gfgpu_comm_create": (C.c_int, [_P, C.c_int, C.c_int, _P, _PP]),
    "gfgpu_comm_destroy": (C.c_int, [_P]),
    "gfgpu_comm_rank": (C.c_int, [_P]),
    "gfgpu_comm_size": (C.c_int, [_P]),
    "gfgpu_term_owned_range": (C.c_int, [_P, _P, _P]),
}


class GfgpuError(RuntimeError):
    """Mirrors gmm::gmm_error: raised for every non-zero status of the C ABI."""


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GfgpuError("%s is missing: run `python -m getfem_b200.build` (there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise GfgpuError(lib().gfgpu_last_error().decode())


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def launch_count():
    return int(lib().gfgpu_launch_count())


class _Handle:
    _destroy = None

    def __init__(self):
        self.h = C.c_void_p()

    def close(self):
        if self.h is not None and self.h.value:
            getattr(lib(), self._destroy)(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context(_Handle):
    _destroy = "gfgpu_ctx_destroy"

    def __init__(self, device=0, stream=None):
        super().__init__()
        check(lib().gfgpu_ctx_create(int(device), C.c_void_p(stream or 0), C.byref(self.h)))
        self.device = device

    def synchronize(self):
        check(lib().gfgpu_ctx_synchronize(self.h))

    def bytes_in_use(self):
        return int(lib().gfgpu_ctx_bytes_in_use(self.h))

    def measure_dmma_peak(self):
        """fp64 tensor-core (mma.sync m8n8k4) peak in TFLOP/s measured on this device."""
        v = C.c_double()
        check(lib().gfgpu_ctx_measure_dmma_peak(self.h, C.byref(v)))
        return float(v.value)

    def measure_fp64_peak(self):
        """fp64 FMA peak in TFLOP/s measured on this device (DFMA chains in registers)."""
        v = C.c_double()
        check(lib().gfgpu_ctx_measure_fp64_peak(self.h, C.byref(v)))
        return float(v.value)


class DeviceMesh(_Handle):
    _destroy = "gfgpu_mesh_destroy"

    def __init__(self, ctx, pts, conn, gt_kind):
        super().__init__()
        pts = np.ascontiguousarray(pts, np.float64)
        conn = np.ascontiguousarray(conn, np.int32)
        self.ctx, self.dim, self.ne, self.ng, self.npts = ctx, pts.shape[1], conn.shape[0], conn.shape[1], pts.shape[0]
        check(lib().gfgpu_mesh_create(ctx.h, self.dim, self.npts, ptr(pts), self.ne, self.ng, ptr(conn), int(gt_kind),
                                      C.byref(self.h)))


class DeviceFem(_Handle):
    _destroy = "gfgpu_fem_destroy"

    def __init__(self, ctx, mesh, fem_kind, degree, qdim, nd, elem_dof=None, ndof=0):
        super().__init__()
        self.ctx, self.mesh, self.qdim, self.nd = ctx, mesh, qdim, nd
        ed = None if elem_dof is None else np.ascontiguousarray(elem_dof, np.int64)
        check(lib().gfgpu_fem_create(ctx.h, mesh.h, int(fem_kind), int(degree), int(qdim), int(nd), ptr(ed), int(ndof),
                                     C.byref(self.h)))
        self.ndof = int(lib().gfgpu_fem_nb_dof(self.h))

    def elem_dof(self):
        out = np.empty((self.mesh.ne, self.nd), np.int64)
        check(lib().gfgpu_fem_get_elem_dof(self.h, ptr(out)))
        return out


class DeviceTables(_Handle):
    _destroy = "gfgpu_tables_destroy"

    def __init__(self, ctx, w, gt_grad, phi, gphi):
        super().__init__()
        w = np.ascontiguousarray(w, np.float64)
        gt_grad = np.ascontiguousarray(gt_grad, np.float64)
        phi = np.ascontiguousarray(phi, np.float64)
        gphi = np.ascontiguousarray(gphi, np.float64)
        nq, ng, dim = gt_grad.shape
        nd = phi.shape[1]
        assert w.shape == (nq,) and gphi.shape == (nq, nd, dim)
        self.nq, self.ng, self.nd, self.dim = nq, ng, nd, dim
        check(lib().gfgpu_tables_create(ctx.h, dim, nq, ng, nd, ptr(w), ptr(gt_grad), ptr(phi), ptr(gphi),
                                        C.byref(self.h)))

    def set_gt_values(self, gt_val, face_gt_val=None):
        """[nq, ng] shape values of the geometric transformation at the volume points (the position X in JIT integrands);
        face_gt_val [nf, nqf, ng]: the same at the face points (after set_faces)"""
        gt_val = np.ascontiguousarray(gt_val, np.float64)
        assert gt_val.shape == (self.nq, self.ng)
        if face_gt_val is not None:
            face_gt_val = np.ascontiguousarray(face_gt_val, np.float64)
            assert face_gt_val.ndim == 3 and face_gt_val.shape[2] == self.ng
        check(lib().gfgpu_tables_set_gt_values(self.h, ptr(gt_val), ptr(face_gt_val) if face_gt_val is not None else None))

    def set_faces(self, normals, w, gt_grad, phi, gphi):
        """Tables at the face points: normals [nf, dim], w [nf, nqf], gt_grad [nf, nqf, ng, dim], phi [nf, nqf, nd],
        gphi [nf, nqf, nd, dim]."""
        normals = np.ascontiguousarray(normals, np.float64)
        w = np.ascontiguousarray(w, np.float64)
        gt_grad = np.ascontiguousarray(gt_grad, np.float64)
        phi = np.ascontiguousarray(phi, np.float64)
        gphi = np.ascontiguousarray(gphi, np.float64)
        nf, nqf = w.shape
        assert normals.shape == (nf, self.dim) and gt_grad.shape == (nf, nqf, self.ng, self.dim)
        assert phi.shape == (nf, nqf, self.nd) and gphi.shape == (nf, nqf, self.nd, self.dim)
        check(lib().gfgpu_tables_set_faces(self.h, nf, nqf, ptr(normals), ptr(w), ptr(gt_grad), ptr(phi), ptr(gphi)))

    def set_faces_from_all_points(self, face_first, face_nq, normals, w, gt_grad, phi, gphi):
        """Same from tables over ALL integration points of an approx_integration (volume points, then face after
        face): the points of face f are [face_first[f], face_first[f] + face_nq[f])."""
        face_first, face_nq = np.asarray(face_first), np.asarray(face_nq)
        if len(set(int(v) for v in face_nq)) != 1:
            raise GfgpuError("the faces of the integration method carry different numbers of points")
        idx = np.array([np.arange(f0, f0 + n) for f0, n in zip(face_first, face_nq)])
        self.set_faces(normals, np.asarray(w)[idx], np.asarray(gt_grad)[idx], np.asarray(phi)[idx], np.asarray(gphi)[idx])


class DeviceTerm(_Handle):
    _destroy = "gfgpu_term_destroy"

    def __init__(self, ctx, mesh, fem, tables, family, params, alpha=1.0, strategy=STRATEGY_AUTO):
        super().__init__()
        self.ctx, self.mesh, self.fem, self.tables = ctx, mesh, fem, tables
        params = np.ascontiguousarray(params, np.float64)
        fam = FAMILY_BY_NAME[family] if isinstance(family, str) else int(family)
        check(lib().gfgpu_term_create(ctx.h, mesh.h, fem.h, tables.h, fam, ptr(params), len(params), float(alpha),
                                      int(strategy), C.byref(self.h)))

    @classmethod
    def jit(cls, ctx, mesh, fem, tables, form1, form2, params=(), alpha=1.0, value_dependent=True):
        """A JIT term (gfgpu_term_create_jit): the order-1 / order-2 forms of a scalar variable as C expressions in
        u, gu, par[k], tv, tg, t2v, t2g; compiled with NVRTC at the first assembly."""
        self = cls.__new__(cls)
        _Handle.__init__(self)
        self.ctx, self.mesh, self.fem, self.tables = ctx, mesh, fem, tables
        params = np.ascontiguousarray(params, np.float64)
        check(lib().gfgpu_term_create_jit(ctx.h, mesh.h, fem.h, tables.h, form1.encode(), form2.encode(),
                                          ptr(params) if len(params) else None, len(params), float(alpha),
                                          1 if value_dependent else 0, C.byref(self.h)))
        return self

    def set_params(self, params):
        params = np.ascontiguousarray(params, np.float64)
        check(lib().gfgpu_term_set_params(self.h, ptr(params), len(params)))

    def set_region(self, cv, face=None):
        """Integrate over the items (cv[k], face[k]) in mr_visitor order; face None / -1 = whole convexes;
        cv None = back to all convexes."""
        if cv is None:
            check(lib().gfgpu_term_set_region(self.h, 0, None, None))
            return
        cv = np.ascontiguousarray(cv, np.int32)
        fc = None if face is None else np.ascontiguousarray(face, np.int32)
        check(lib().gfgpu_term_set_region(self.h, len(cv), ptr(cv), ptr(fc)))

    def set_fields(self, data_fem, phi, vals, phi_faces=None):
        """Fem-data coefficients: vals = [values of field 0 (, field 1)] on data_fem (a DeviceFem of the same mesh);
        phi [nq, nd_d] = its basis at the volume points, phi_faces [nf, nqf, nd_d] at the face points."""
        if data_fem is None:
            check(lib().gfgpu_term_set_fields(self.h, 0, None, None, None, None, None))
            self._dfem = None
            return
        phi = np.ascontiguousarray(phi, np.float64)
        pf = None if phi_faces is None else np.ascontiguousarray(phi_faces, np.float64)
        v = [np.ascontiguousarray(x, np.float64) for x in vals]
        assert phi.shape[1] == data_fem.nd and all(len(x) == data_fem.ndof for x in v)
        check(lib().gfgpu_term_set_fields(self.h, len(v), data_fem.h, ptr(phi), ptr(pf), ptr(v[0]),
                                          ptr(v[1]) if len(v) > 1 else None))
        self._dfem = data_fem  # keep it alive

    def update_field(self, k, vals):
        check(lib().gfgpu_term_update_field(self.h, int(k), ptr(np.ascontiguousarray(vals, np.float64))))

    def set_element_range(self, e0, e1):
        check(lib().gfgpu_term_set_element_range(self.h, int(e0), int(e1)))

    def assemble_dev(self, U_dev_ptr, order_mask):
        check(lib().gfgpu_term_assemble_dev(self.h, C.c_void_p(U_dev_ptr or 0), int(order_mask)))

    def assemble_host(self, U, order_mask, pr_out=None, R_out=None):
        U = None if U is None else np.ascontiguousarray(U, np.float64)
        check(lib().gfgpu_term_assemble_host(self.h, ptr(U), int(order_mask), ptr(pr_out), ptr(R_out)))

    def potential_host(self, U):
        """order 0: the term's potential at the state U (ga_workspace::assembly(0) / assembled_potential())"""
        U = None if U is None else np.ascontiguousarray(U, np.float64)
        E = C.c_double()
        check(lib().gfgpu_term_potential_host(self.h, ptr(U), C.byref(E)))
        return E.value

    def last_timings(self):
        """dict of device ms of the last assemble: elem, gather, rgather, pattern."""
        out = np.zeros(8, np.float32)
        check(lib().gfgpu_term_last_timings(self.h, ptr(out)))
        return dict(zip(("elem", "gather", "rgather", "pattern", "recompute"), (float(v) for v in out[:5])))

    @property
    def strategy(self):
        return int(lib().gfgpu_term_strategy(self.h))

    @property
    def kernel_kind(self):
        """0 generic element kernel, 1 general tile kernel, 2 column kernel, 3 class-uniform tile kernel, 4 direct mode."""
        return int(lib().gfgpu_term_kernel_kind(self.h))

    @property
    def nnz(self):
        return int(lib().gfgpu_term_nnz(self.h))

    @property
    def ndof(self):
        return int(lib().gfgpu_term_nb_dof(self.h))

    @property
    def pattern_generation(self):
        return int(lib().gfgpu_term_pattern_generation(self.h))

    def csc_view(self):
        jc, ir, pr = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib().gfgpu_term_csc_view(self.h, C.byref(jc), C.byref(ir), C.byref(pr)))
        return jc.value, ir.value, pr.value

    def tmult_dev(self, x_dev_ptr, y_dev_ptr, alpha=1.0, beta=0.0):
        """y = beta y + alpha K^T x on the device (the term's own CSC)"""
        check(lib().gfgpu_term_tmult_dev(self.h, float(alpha), C.c_void_p(x_dev_ptr), float(beta), C.c_void_p(y_dev_ptr)))

    def residual_view(self):
        r = C.c_void_p()
        check(lib().gfgpu_term_residual_view(self.h, C.byref(r)))
        return r.value

    def export_csc(self, values=True):
        nnz, n = self.nnz, self.ndof
        jc = np.empty(n + 1, np.int64)
        ir = np.empty(nnz, np.int32)
        pr = np.empty(nnz, np.float64) if values else None
        check(lib().gfgpu_term_export_csc_host(self.h, ptr(jc), ptr(ir), ptr(pr)))
        return jc, ir, pr

    def residual_add_dev(self, rhs_dev_ptr, alpha=1.0, row_off=0):
        """rhs[row_off + i] += alpha R_i on the device (the model's rrhs, getfem_models.cc:2553-2570)."""
        check(lib().gfgpu_term_residual_add_dev(self.h, float(alpha), C.c_void_p(rhs_dev_ptr), int(row_off)))

    def export_residual(self):
        R = np.empty(self.ndof, np.float64)
        check(lib().gfgpu_term_export_residual_host(self.h, ptr(R)))
        return R

    # ---- multi-GPU halo (include/gfgpu.h, "multi-GPU"); getfem_b200/halo.py drives these
    def halo_begin(self, U_dev_ptr=None):
        lo, hi = _i64(), _i64()
        check(lib().gfgpu_term_halo_begin(self.h, C.c_void_p(U_dev_ptr or 0), C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def halo_ghost_pairs(self, dof_lo, dof_hi):
        n = _i64()
        check(lib().gfgpu_term_halo_ghost_pairs(self.h, int(dof_lo), int(dof_hi), C.byref(n), None, None, None))
        J, I, m = np.empty(n.value, np.int32), np.empty(n.value, np.int32), np.empty(n.value, np.uint16)
        if n.value:
            check(lib().gfgpu_term_halo_ghost_pairs(self.h, int(dof_lo), int(dof_hi), C.byref(n), ptr(J), ptr(I), ptr(m)))
        return J, I, m

    def halo_add_source(self, src_rank, J, I, mask, r_lo, r_hi):
        J, I = np.ascontiguousarray(J, np.int32), np.ascontiguousarray(I, np.int32)
        mask = np.ascontiguousarray(mask, np.uint16)
        check(lib().gfgpu_term_halo_add_source(self.h, int(src_rank), len(J), ptr(J), ptr(I), ptr(mask), int(r_lo), int(r_hi)))

    def halo_commit(self, own_lo, own_hi):
        check(lib().gfgpu_term_halo_commit(self.h, int(own_lo), int(own_hi)))

    def halo_send_buffers(self, dof_lo, dof_hi, r_lo):
        """(values of my ghost columns [dof_lo, dof_hi), residual slice [r_lo, dof_hi)) as zero-copy device tensors."""
        pr, R, cnt = C.c_void_p(), C.c_void_p(), _i64()
        check(lib().gfgpu_term_halo_send_view(self.h, int(dof_lo), int(dof_hi), C.byref(pr), C.byref(cnt), C.byref(R)))
        rp = (R.value + 8 * (int(r_lo) - int(dof_lo))) if R.value else None
        return _dev_tensor(pr.value, cnt.value, self.ctx.device), _dev_tensor(rp, int(dof_hi) - int(r_lo), self.ctx.device)

    def halo_recv_buffers(self, src_rank):
        pr, R, cnt, rc = C.c_void_p(), C.c_void_p(), _i64(), _i64()
        check(lib().gfgpu_term_halo_recv_view(self.h, int(src_rank), C.byref(pr), C.byref(cnt), C.byref(R), C.byref(rc)))
        return _dev_tensor(pr.value, cnt.value, self.ctx.device), _dev_tensor(R.value, rc.value, self.ctx.device)

    def halo_accumulate(self, order_mask):
        check(lib().gfgpu_term_halo_accumulate(self.h, int(order_mask)))

    def halo_add_send(self, owner_rank, dof_lo, dof_hi, r_lo):
        check(lib().gfgpu_term_halo_add_send(self.h, int(owner_rank), int(dof_lo), int(dof_hi), int(r_lo)))

    def halo_exchange(self, comm, order_mask):
        """the exchange inside the library: one ncclGroup of sends / receives on the context's stream + accumulation"""
        check(lib().gfgpu_term_halo_exchange(self.h, comm.h, int(order_mask)))

    def ctx_synchronize(self):
        self.ctx.synchronize()

    def owned_range(self):
        lo, hi = _i64(), _i64()
        check(lib().gfgpu_term_owned_range(self.h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value


COMM_ID_BYTES = 128


def comm_unique_id():
    """rendezvous id of a new communicator (rank 0 creates it, every rank passes it to Communicator)"""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    check(lib().gfgpu_comm_unique_id(buf, COMM_ID_BYTES))
    return bytes(buf.raw)


class Communicator(_Handle):
    """gfgpu_comm: NCCL communicator of the ranks that share an assembly (one process per GPU)."""
    _destroy = "gfgpu_comm_destroy"

    def __init__(self, ctx, nranks, rank, unique_id):
        super().__init__()
        assert len(unique_id) == COMM_ID_BYTES
        self.ctx = ctx
        buf = C.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        check(lib().gfgpu_comm_create(ctx.h, int(nranks), int(rank), buf, C.byref(self.h)))

    @property
    def rank(self):
        return int(lib().gfgpu_comm_rank(self.h))

    @property
    def size(self):
        return int(lib().gfgpu_comm_size(self.h))


MODEL_LINEAR, MODEL_SYMMETRIC, BUILD_MATRIX = 1, 2, 4


class DeviceMatrix(_Handle):
    """The workspace-level tangent on the device: the sum of terms at their variable offsets (gfgpu_matrix_*)."""
    _destroy = "gfgpu_matrix_destroy"

    def __init__(self, ctx, nrows, ncols=None):
        super().__init__()
        self.ctx, self.nrows, self.ncols = ctx, int(nrows), int(nrows if ncols is None else ncols)
        check(lib().gfgpu_matrix_create(ctx.h, self.nrows, self.ncols, C.byref(self.h)))

    def clear(self, keep_pattern=False):
        check(lib().gfgpu_matrix_clear(self.h, 1 if keep_pattern else 0))

    def add_term(self, term, alpha=1.0, row_off=0, col_off=0):
        check(lib().gfgpu_matrix_add_term(self.h, term.h, float(alpha), int(row_off), int(col_off)))

    def add_term_reduced(self, term, E, alpha=1.0, row_off=0, col_off=0):
        """K += alpha E^T K_term E (a term assembled on the basic dofs of a reduced mesh_fem, workspace.cc:861-935)"""
        check(lib().gfgpu_matrix_add_term_reduced(self.h, term.h, E.h, float(alpha), int(row_off), int(col_off)))

    def add_rect_reduced(self, rect, E_rows=None, E_cols=None, transposed=False, alpha=1.0, row_off=0, col_off=0):
        check(lib().gfgpu_matrix_add_rect_reduced(self.h, rect.h, 1 if transposed else 0, E_rows.h if E_rows else None,
                                                  E_cols.h if E_cols else None, float(alpha), int(row_off), int(col_off)))

    @property
    def nnz(self):
        return int(lib().gfgpu_matrix_nnz(self.h))

    @property
    def pattern_generation(self):
        return int(lib().gfgpu_matrix_pattern_generation(self.h))

    def csc_view(self):
        jc, ir, pr = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib().gfgpu_matrix_csc_view(self.h, C.byref(jc), C.byref(ir), C.byref(pr)))
        return jc.value, ir.value, pr.value

    def export_csc(self):
        nnz = self.nnz
        jc = np.empty(self.ncols + 1, np.int64)
        ir = np.empty(nnz, np.int32)
        pr = np.empty(nnz, np.float64)
        check(lib().gfgpu_matrix_export_csc_host(self.h, ptr(jc), ptr(ir), ptr(pr)))
        return jc, ir, pr

    def mult(self, x, transposed=False, alpha=1.0, beta=0.0, y=None):
        """y = beta*y + alpha*K x (or K^T x) through host buffers."""
        x = np.ascontiguousarray(x, np.float64)
        nout = self.ncols if transposed else self.nrows
        y = np.zeros(nout) if y is None else np.ascontiguousarray(y, np.float64).copy()
        check(lib().gfgpu_matrix_mult_host(self.h, 1 if transposed else 0, float(alpha), ptr(x), float(beta), ptr(y)))
        return y

    def mult_dev(self, x_dev_ptr, y_dev_ptr, transposed=False, alpha=1.0, beta=0.0):
        check(lib().gfgpu_matrix_mult_dev(self.h, 1 if transposed else 0, float(alpha), C.c_void_p(x_dev_ptr), float(beta),
                                          C.c_void_p(y_dev_ptr)))

    def add_rect(self, rect, transposed=False, alpha=1.0, row_off=0, col_off=0):
        check(lib().gfgpu_matrix_add_rect(self.h, rect.h, 1 if transposed else 0, float(alpha), int(row_off), int(col_off)))

    def apply_dof_constraints(self, dofs, values, rhs_dev_ptr=None, present=None, linear=True, symmetric=True,
                              build_matrix=True):
        """Dirichlet conditions with simplification on the resident tangent and a device right-hand side
        (model::real_dof_constraints, getfem_models.cc:2806-2871)."""
        dofs = np.ascontiguousarray(dofs, np.int64)
        values = np.ascontiguousarray(values, np.float64)
        assert dofs.shape == values.shape
        pres = None if present is None else np.ascontiguousarray(present, np.float64)
        flags = (MODEL_LINEAR if linear else 0) | (MODEL_SYMMETRIC if symmetric else 0) | (BUILD_MATRIX if build_matrix else 0)
        check(lib().gfgpu_matrix_apply_dof_constraints(self.h, dofs.size, ptr(dofs), ptr(values),
                                                       None if pres is None else ptr(pres),
                                                       None if rhs_dev_ptr is None else C.c_void_p(rhs_dev_ptr), flags))

    def export_csr_dev(self, rowptr_dev_ptr, col_dev_ptr, val_dev_ptr):
        """CSR (int64 row pointers, int32 columns, values) into caller-owned device buffers: the solver hand-off."""
        vp = lambda p: None if p is None else C.c_void_p(p)  # noqa: E731
        check(lib().gfgpu_matrix_export_csr_dev(self.h, vp(rowptr_dev_ptr), vp(col_dev_ptr), vp(val_dev_ptr)))

    def cg_dev(self, b_dev_ptr, x_dev_ptr, rtol=1e-10, max_iter=10000):
        """Jacobi-preconditioned CG on the device; returns (iterations, relative residual)."""
        it, rr = C.c_int(0), C.c_double(0.0)
        check(lib().gfgpu_matrix_cg_dev(self.h, C.c_void_p(b_dev_ptr), C.c_void_p(x_dev_ptr), float(rtol), int(max_iter),
                                        C.byref(it), C.byref(rr)))
        return it.value, rr.value


RECT_DIV_PRESSURE = 0
RECT_MASS = 1


def jit_check(dim, form1, form2, qdim=1):
    """Compile the two forms of a JIT term with NVRTC (no GPU needed); raises GfgpuError with the compiler log."""
    check(lib().gfgpu_jit_check(int(dim), int(qdim), form1.encode(), form2.encode()))


class DeviceRect(_Handle):
    """A coupled bilinear term (gfgpu_rect_*): Test on `fem_rows`, Test2 on `fem_cols`; the block and its transpose."""
    _destroy = "gfgpu_rect_destroy"

    def __init__(self, ctx, mesh, fem_rows, tab_rows, fem_cols, tab_cols, family=RECT_DIV_PRESSURE, coef=1.0, alpha=1.0):
        super().__init__()
        self.ctx, self.nrows, self.ncols = ctx, fem_rows.ndof, fem_cols.ndof
        self._keep = (mesh, fem_rows, tab_rows, fem_cols, tab_cols)
        check(lib().gfgpu_rect_create(ctx.h, mesh.h, fem_rows.h, tab_rows.h, fem_cols.h, tab_cols.h, int(family), float(coef),
                                      float(alpha), C.byref(self.h)))

    def set_region(self, cv, face=None):
        cv = None if cv is None else np.ascontiguousarray(cv, np.int32)
        face = None if face is None else np.ascontiguousarray(face, np.int32)
        check(lib().gfgpu_rect_set_region(self.h, 0 if cv is None else len(cv), ptr(cv) if cv is not None else None,
                                          ptr(face) if face is not None else None))

    def assemble(self):
        check(lib().gfgpu_rect_assemble_dev(self.h))

    @property
    def nnz(self):
        return int(lib().gfgpu_rect_nnz(self.h))

    def export_csc(self, transposed=False):
        nnz = max(self.nnz, 0)  # (-1 before the first assembly: the call below then reports it)
        jc = np.empty((self.nrows if transposed else self.ncols) + 1, np.int64)
        ir = np.empty(nnz, np.int32)
        pr = np.empty(nnz, np.float64)
        check(lib().gfgpu_rect_export_csc_host(self.h, 1 if transposed else 0, ptr(jc), ptr(ir), ptr(pr)))
        return jc, ir, pr

    def mult(self, x, transposed=False, alpha=1.0, beta=0.0, y=None):
        x = np.ascontiguousarray(x, np.float64)
        nout = self.ncols if transposed else self.nrows
        y = np.zeros(nout) if y is None else np.ascontiguousarray(y, np.float64).copy()
        check(lib().gfgpu_rect_mult_host(self.h, 1 if transposed else 0, float(alpha), ptr(x), float(beta), ptr(y)))
        return y


class DeviceReduction(_Handle):
    """The extension matrix E (nb_basic_dof x nb_dof) of a reduced mesh_fem, given as a scipy CSR matrix or (rowptr, col, val)."""
    _destroy = "gfgpu_reduction_destroy"

    def __init__(self, ctx, n_basic, n_dof, rowptr, col, val):
        super().__init__()
        self.ctx, self.n_basic, self.n_dof = ctx, int(n_basic), int(n_dof)
        rowptr = np.ascontiguousarray(rowptr, np.int64)
        col = np.ascontiguousarray(col, np.int32)
        val = np.ascontiguousarray(val, np.float64)
        assert len(rowptr) == self.n_basic + 1
        check(lib().gfgpu_reduction_create(ctx.h, self.n_basic, self.n_dof, ptr(rowptr), ptr(col), ptr(val), C.byref(self.h)))

    def extend(self, x):
        x = np.ascontiguousarray(x, np.float64)
        y = np.empty(self.n_basic)
        check(lib().gfgpu_reduction_extend_host(self.h, ptr(x), ptr(y)))
        return y

    def restrict_add(self, x, y, alpha=1.0):
        x = np.ascontiguousarray(x, np.float64)
        y = np.ascontiguousarray(y, np.float64).copy()
        check(lib().gfgpu_reduction_restrict_add_host(self.h, float(alpha), ptr(x), ptr(y)))
        return y


class _DevArray:
    def __init__(self, p, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (p, False), "version": 2}


def _dev_tensor(p, n, device):
    """Zero-copy torch view of `n` doubles of library-owned device memory (torch is plumbing: NCCL p2p on it)."""
    import torch
    if not p or n <= 0:
        return torch.empty(0, dtype=torch.float64, device="cuda:%d" % device)
    return torch.as_tensor(_DevArray(p, n), device="cuda:%d" % device)
