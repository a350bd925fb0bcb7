"""Host-side mirror of the reference interface for the accelerated path.

Same names and argument meaning as GetFEM's C++ API (the parity tests read like tests/test_assembly.cc):

  mesh          regular_unit_mesh(m, nsubdiv, pgt)              src/getfem_regular_meshes.cc:237-284
  mesh_fem      mesh_fem(m, Qdim); set_classical_finite_element(K); nb_dof();
                ind_scalar_basic_dof_of_element(cv)              src/getfem/getfem_mesh_fem.h:459-461
  mesh_im       mesh_im(m); set_integration_method(degree)       src/getfem/getfem_mesh_im.h
  mesh_region   add(cv[, f]), is_only_faces(); outer_faces_of_mesh(m)   src/getfem/getfem_mesh_region.h, getfem_mesh.h:629-645
  ga_workspace  add_fem_variable, add_fixed_size_constant, add_expression, assembly(order),
                assembled_matrix(), assembled_vector()           src/getfem/getfem_generic_assembly.h:262-597

Everything numerical happens on the device through the C ABI (capi.py); there is no CPU fallback:
an expression that is not one of the recognised families raises, exactly like GMM_ASSERT1 would.
"""
import re

import numpy as np

from . import capi, fem_tables
from .regular_mesh import regular_unit_mesh as _regular_unit_mesh

_ctx = {}


def default_context(device=None, stream=None):
    """One context per (device, stream).  The device defaults to LOCAL_RANK (one process per GPU)."""
    import os
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    key = (device, stream)
    if key not in _ctx:
        _ctx[key] = capi.Context(device, stream)
    return _ctx[key]


class mesh:
    """getfem::mesh restricted to what the path reads: points and convex connectivity."""

    def __init__(self, pts=None, conn=None, gt="GT_PK"):
        self.pts = None if pts is None else np.ascontiguousarray(pts, np.float64)
        self.conn = None if conn is None else np.ascontiguousarray(conn, np.int32)
        self.gt = gt  # "GT_PK" (affine simplices) or "GT_QK" (multilinear parallelepipeds), degree 1
        self._dev = {}

    def dim(self):
        return self.pts.shape[1]

    def nb_convex(self):
        return self.conn.shape[0]

    def nb_points(self):
        return self.pts.shape[0]

    def device(self, ctx):
        if ctx not in self._dev:
            self._dev[ctx] = capi.DeviceMesh(ctx, self.pts, self.conn, capi.GT_PK if self.gt == "GT_PK" else capi.GT_QK)
        return self._dev[ctx]

    def face_local_nodes(self):
        """Local geometric nodes of every face of the (degree-1) reference convex: [nf, nodes per face].
        Simplex: face f holds every vertex but vertex f (bgeot_convex_structure.cc simplex_structure); parallelepiped
        (node ix + 2 iy + 4 iz): face 2d is x_d = 1, face 2d+1 is x_d = 0 -- the numbering of fem_tables.face_dir_points."""
        N, ng = self.dim(), self.conn.shape[1]
        if self.gt == "GT_PK":
            return np.array([[v for v in range(ng) if v != f] for f in range(N + 1)])
        out = []
        for d in range(N):
            for side in (1, 0):
                out.append([v for v in range(ng) if ((v >> d) & 1) == side])
        return np.array(out)

    def points_of_face_of_convex(self, cv, f):
        return self.pts[self.conn[cv][self.face_local_nodes()[f]]]


class mesh_region:
    """getfem::mesh_region restricted to what the path reads: a set of convexes or of (convex, face) pairs, walked in
    mr_visitor order (ascending convex, then ascending face; getfem_mesh_region.h)."""

    def __init__(self):
        self._items = set()

    def add(self, cv, f=-1):
        self._items.add((int(cv), int(f)))

    def is_only_faces(self):
        return bool(self._items) and all(f >= 0 for _, f in self._items)

    def is_only_convexes(self):
        return all(f < 0 for _, f in self._items)

    def items(self):
        """(cv [n], face [n]) int32 arrays in visitor order."""
        it = sorted(self._items)
        return np.array([c for c, _ in it], np.int32), np.array([f for _, f in it], np.int32)

    def __len__(self):
        return len(self._items)


def outer_faces_of_mesh(m):
    """getfem::outer_faces_of_mesh (getfem_mesh.h:629-645): the faces that belong to one convex only."""
    fl = m.face_local_nodes()  # [nf, nv]
    nf = fl.shape[0]
    keys = np.sort(m.conn[:, fl], axis=2).reshape(-1, fl.shape[1])  # [ne*nf, nv] sorted vertex ids
    _, inv, cnt = np.unique(keys, axis=0, return_inverse=True, return_counts=True)
    outer = np.nonzero(cnt[inv.reshape(-1)] == 1)[0]
    rg = mesh_region()
    for k in outer:
        rg.add(k // nf, k % nf)
    return rg


def regular_unit_mesh(m, nsubdiv, pgt):
    """pgt: 'GT_PK(N,1)' or 'GT_QK(N,1)' (name_of_geometric_trans syntax)."""
    mt = re.fullmatch(r"GT_(PK|QK)\((\d),1\)", pgt.replace(" ", ""))
    if not mt or int(mt.group(2)) != len(nsubdiv):
        raise capi.GfgpuError("cannot build a regular mesh for " + pgt)
    m.gt = "GT_" + mt.group(1)
    m.pts, m.conn = _regular_unit_mesh(nsubdiv, "simplex" if mt.group(1) == "PK" else "parallelepiped")
    m._dev = {}
    return m


class mesh_fem:
    def __init__(self, m, Qdim=1):
        self.linked_mesh = m
        self.Qdim = int(Qdim)
        self.K = None
        self._elem_dof = None
        self._ndof = None
        self._dev = {}

    def set_classical_finite_element(self, K):
        self.K = int(K)
        self._elem_dof = None
        self._ndof = None
        self._dev = {}

    def fem_kind(self):
        return "PK" if self.linked_mesh.gt == "GT_PK" else "QK"

    def nb_basic_dof_of_element(self):
        return fem_tables.nb_dof(self.fem_kind(), self.linked_mesh.dim(), self.K) * self.Qdim

    def set_dof_table(self, elem_dof, ndof):
        """Adopt a numbering produced elsewhere (e.g. read from the reference's mesh_fem)."""
        self._elem_dof = np.ascontiguousarray(elem_dof, np.int64)
        self._ndof = int(ndof)
        self._dev = {}

    def device(self, ctx):
        if ctx not in self._dev:
            m = self.linked_mesh
            nd = fem_tables.nb_dof(self.fem_kind(), m.dim(), self.K)
            kind = capi.FEM_PK if self.fem_kind() == "PK" else capi.FEM_QK
            self._dev[ctx] = capi.DeviceFem(ctx, m.device(ctx), kind, self.K, self.Qdim, nd, self._elem_dof,
                                            self._ndof or 0)
            if self._ndof is None:
                self._ndof = self._dev[ctx].ndof
        return self._dev[ctx]

    def nb_dof(self, ctx=None):
        if self._ndof is None:
            self.device(ctx or default_context())  # enumerate_dof on the device
        return self._ndof

    def ind_scalar_basic_dof_of_element(self, cv=None, ctx=None):
        """Element -> dof table (component 0 of each local node); all elements when cv is None."""
        if self._elem_dof is None:
            self._elem_dof = self.device(ctx or default_context()).elem_dof()
        return self._elem_dof if cv is None else self._elem_dof[cv]

    def basic_dof_nodes(self, ctx=None):
        """Physical coordinates of the node of every dof (mesh_fem::point_of_basic_dof)."""
        m = self.linked_mesh
        ed = self.ind_scalar_basic_dof_of_element(ctx=ctx)
        kind = self.fem_kind()
        X = fem_tables.ref_nodes(kind, m.dim(), self.K)
        shp, _ = fem_tables.lagrange_tables(kind, m.dim(), 1, X)  # [nd, ng]
        nodes = np.einsum("ig,egd->eid", shp, m.pts[m.conn])  # [ne, nd, dim]
        out = np.zeros((self.nb_dof(ctx), m.dim()))
        for q in range(self.Qdim):
            out[(ed + q).reshape(-1)] = nodes.reshape(-1, m.dim())
        return out


class mesh_im:
    def __init__(self, m):
        self.linked_mesh = m
        self.degree = None

    def set_integration_method(self, degree):
        self.degree = int(degree)


# ---------------------------------------------------------------- expression recognition
_ID = r"[A-Za-z_][A-Za-z_0-9]*"
_LAWS = {
    "Saint_Venant_Kirchhoff": "svk",
    "Compressible_Neo_Hookean_Ciarlet": "nh_ciarlet",
    "Compressible_Neo_Hookean_Bonet": "nh_bonet",
    "Compressible_Mooney_Rivlin": "mooney_rivlin",
    "Ciarlet_Geymonat": "ciarlet_geymonat",
    "Generalized_Blatz_Ko": "blatz_ko",
}


def _strip_outer(s):
    while s.startswith("(") and s.endswith(")"):
        depth = 0
        for i, ch in enumerate(s):
            depth += ch == "("
            depth -= ch == ")"
            if depth == 0 and i < len(s) - 1:
                return s
        s = s[1:-1]
    return s


def recognise(expr):
    """Maps a GWFL string to (family, variable, constant names).  Covers the strings the reference's
    bricks generate (getfem_models.cc:3943-3997, 6112-6113; getfem_nonlinear_elasticity.cc:2319-2320)
    and their usual hand-written forms."""
    s = _strip_outer(re.sub(r"\s+", "", expr))
    m = re.fullmatch(rf"(?:\(?({_ID})\)?\*)?\(?Grad_({_ID})[.:]Grad_Test_\2\)?", s)
    if m:
        return "laplace", m.group(2), [m.group(1)] if m.group(1) else []
    m = re.fullmatch(rf"\(?\(?({_ID})\)?\*Grad_({_ID})\)?[.:]Grad_Test_\2", s)
    if m:
        return "laplace", m.group(2), [m.group(1)]
    m = re.fullmatch(rf"(?:\(?({_ID})\)?\*)?\(?({_ID})[.:]Test_\2\)?", s)
    if m and not m.group(2).startswith("Grad_"):
        return "mass", m.group(2), [m.group(1)] if m.group(1) else []
    m = re.fullmatch(rf"\(Div_({_ID})\*\(\(?({_ID})\)?\*Id\(meshdim\)\)\+\(2\*\(?({_ID})\)?\)\*Sym\(Grad_\1\)\):Grad_Test_\1", s)
    if m:
        return "elast", m.group(1), [m.group(2), m.group(3)]
    m = re.fullmatch(rf"\(?({_ID})\)?\*\(?Div_({_ID})\*Div_Test_\2\)?\+\(?2\*\(?({_ID})\)?\)?\*\(?Sym\(Grad_\2\):Grad_Test_\2\)?", s)
    if m:
        return "elast", m.group(2), [m.group(1), m.group(3)]
    # normal source term (add_normal_source_term_brick, getfem_models.cc:4290-4299)
    m = re.fullmatch(rf"(-?)\(?\(\(?({_ID})\)?\.Normal\)\*Test_({_ID})\)?", s)
    if m:
        return ("nsource-" if m.group(1) else "nsource+"), m.group(3), [m.group(2)]
    m = re.fullmatch(rf"(-?)\(?\(Reshape\(({_ID}),qdim\(({_ID})\),meshdim\)\*Normal\)\.Test_\3\)?", s)
    if m:
        return ("nsource-" if m.group(1) else "nsource+"), m.group(3), [m.group(2)]
    m = re.fullmatch(rf"(-?)\(?({_ID})[.*]Test_({_ID})\)?", s)  # volumic source term: "-f*Test_u", "-(F.Test_u)", "F.Test_u"
    if m and not m.group(2).startswith("Grad_") and m.group(2) != m.group(3):
        return ("source-" if m.group(1) else "source+"), m.group(3), [m.group(2)]
    m = re.fullmatch(rf"\(\(Id\(meshdim\)\+Grad_({_ID})\)\*\(?({_ID})_PK2\(Grad_\1,({_ID})\)\)?\):Grad_Test_\1", s)
    if m and m.group(2) in _LAWS:
        return _LAWS[m.group(2)], m.group(1), [m.group(3)]
    raise capi.GfgpuError("expression not handled by the device path (no CPU fallback): " + expr)


class ga_workspace:
    """Device-backed ga_workspace for one fem variable and the recognised expression families."""

    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()
        self.variables = {}   # name -> (mesh_fem, values)
        self.constants = {}   # name -> values
        self.fem_constants = {}   # name -> (mesh_fem, nodal values): ga_workspace::add_fem_constant
        self.terms = []       # (family, variable, params, mim, DeviceTerm or None)
        self._K = None
        self._R = None
        self._Kdev = None

    # ---- declaration API (generic_assembly.h:453-470)
    def add_fem_variable(self, name, mf, I, V):
        self.variables[name] = (mf, V)

    def add_fixed_size_constant(self, name, V):
        self.constants[name] = np.atleast_1d(np.asarray(V, np.float64))

    def add_fem_constant(self, name, mf, V):
        """Fem data (generic_assembly.h:465): a coefficient or load given by nodal values on a mesh_fem."""
        V = np.ascontiguousarray(V, np.float64)
        self.fem_constants[name] = (mf, V)

    def add_expression(self, expr, mim, region=None, add_derivative_order=2):
        fam, var, cnames = recognise(expr)
        if var not in self.variables:
            raise capi.GfgpuError("unknown variable " + var)
        for c in cnames:
            if c not in self.constants and c not in self.fem_constants:
                raise capi.GfgpuError("unknown constant " + c)
        # fem-data coefficients: the leading parameters of the family given as fields on one data mesh_fem
        fields = None
        nfem = [c in self.fem_constants for c in cnames]
        if any(nfem):
            if fam.startswith("nsource") or fam not in ("laplace", "mass", "elast", "source-", "source+"):
                raise capi.GfgpuError("fem-data coefficients are not handled for this expression")
            if not nfem[0] or (len(nfem) > 1 and nfem[1] and not nfem[0]):
                raise capi.GfgpuError("a fem-data second coefficient needs a fem-data first coefficient")
            names = [c for c, f in zip(cnames, nfem) if f]
            mfd = self.fem_constants[names[0]][0]
            if any(self.fem_constants[c][0] is not mfd for c in names):
                raise capi.GfgpuError("the fem-data coefficients of one term must live on the same mesh_fem")
            if mfd.linked_mesh is not self.variables[var][0].linked_mesh:
                raise capi.GfgpuError("the data mesh_fem must be defined on the mesh of the variable")
            sign = -1.0 if fam == "source-" else 1.0
            fields = (mfd, [sign * self.fem_constants[c][1] for c in names])
            for c in names:  # placeholders: the parameter slots of the fields are ignored by the device term
                self.constants.setdefault(c, np.ones(self.variables[var][0].Qdim if fam.startswith("source") else 1))
        if fam.startswith("nsource"):
            mf = self.variables[var][0]
            g = self.constants[cnames[0]]
            if g.size != mf.Qdim * mf.linked_mesh.dim():
                raise capi.GfgpuError("the normal source term needs a constant of qdim x meshdim components")
            params = [(-1.0 if fam.endswith("-") else 1.0) * float(v) for v in g.reshape(-1)]
            fam = "nsource"
        elif fam.startswith("source"):
            mf = self.variables[var][0]
            f = self.constants[cnames[0]]
            if f.size != mf.Qdim:
                raise capi.GfgpuError("the source term needs a constant of qdim components")
            params = [(-1.0 if fam.endswith("-") else 1.0) * float(v) for v in f]
            fam = "source"
        elif fam in ("laplace", "mass"):
            params = [float(self.constants[cnames[0]][0])] if cnames else [1.0]
        elif fam == "elast":
            params = [float(self.constants[cnames[0]][0]), float(self.constants[cnames[1]][0])]
        else:
            p = self.constants[cnames[0]]
            if p.size != {"mooney_rivlin": 3, "ciarlet_geymonat": 3, "blatz_ko": 5}.get(fam, 2):
                raise capi.GfgpuError("wrong number of parameters for the hyperelastic law")
            params = [float(v) for v in p]
        if region is not None:
            if not isinstance(region, mesh_region) or not len(region):
                raise capi.GfgpuError("the region must be a non-empty mesh_region (or None for all convexes)")
            if not (region.is_only_faces() or region.is_only_convexes()):
                raise capi.GfgpuError("a region must hold either convexes or faces, not both")
        if fam == "nsource" and (region is None or not region.is_only_faces()):
            raise capi.GfgpuError("Normal is only defined on a region of faces")
        # ga_workspace::add_tree sums the expressions that share (mim, region, test variable) into ONE tree
        # (workspace.cc:472-493): their element matrices are thresholded together (C&E.cc:4889).  Two bilinear forms on
        # one region would therefore not reproduce the reference's pattern when assembled separately: refuse.
        if fam not in ("source", "nsource"):
            key = None if region is None else frozenset(region._items)
            for f2, v2, _, mim2, _, rg2, _ in self.terms:
                if f2 not in ("source", "nsource") and v2 == var and mim2 is mim and \
                        (None if rg2 is None else frozenset(rg2._items)) == key:
                    raise capi.GfgpuError("several bilinear forms on one region are thresholded together by the reference; "
                                          "not handled by the device path")
        if fields is not None:
            want = self.variables[var][0].Qdim if fam == "source" else 1
            if fields[0].Qdim != want:
                raise capi.GfgpuError("the data mesh_fem must have qdim %d for this term" % want)
        self.terms.append([fam, var, params, mim, None, region, fields])
        return len(self.terms) - 1

    def nb_trees(self):
        return len(self.terms)

    # ---- device objects
    def _term(self, k):
        fam, var, params, mim, dev, region, fields = self.terms[k]
        if dev is None:
            mf, _ = self.variables[var]
            m = mf.linked_mesh
            t = fem_tables.classical_tables(mf.fem_kind(), m.dim(), mf.K, mim.degree)
            tab = capi.DeviceTables(self.ctx, t["quad_w"], t["gt_grad"], t["phi"], t["gphi"])
            if region is not None and region.is_only_faces():
                ft = fem_tables.classical_face_tables(mf.fem_kind(), m.dim(), mf.K, mim.degree)
                tab.set_faces(ft["normals"], ft["quad_w"], ft["gt_grad"], ft["phi"], ft["gphi"])
            dev = capi.DeviceTerm(self.ctx, m.device(self.ctx), mf.device(self.ctx), tab, fam, params)
            if region is not None:
                cv, fc = region.items()
                dev.set_region(cv, fc if region.is_only_faces() else None)
            if fields is not None:
                mfd, vals = fields
                kind = mfd.fem_kind()
                phi = fem_tables.lagrange_tables(kind, m.dim(), mfd.K, t["quad_x"])[0]
                pf = None
                if region is not None and region.is_only_faces():
                    X = ft["quad_x"]
                    pf = fem_tables.lagrange_tables(kind, m.dim(), mfd.K, X.reshape(-1, m.dim()))[0].reshape(X.shape[0], X.shape[1], -1)
                dev.set_fields(mfd.device(self.ctx), phi, vals, pf)
            self.terms[k][4] = dev
        return dev

    def assembly(self, order):
        """order 1: residual vector; order 2: tangent matrix (workspace.cc:791-936).  Every term is assembled on the
        device; the tangents are accumulated there into one matrix (capi.DeviceMatrix) and exported once."""
        if order not in (1, 2):
            raise capi.GfgpuError("only assembly orders 1 and 2 are handled by the device path")
        if not self.terms:
            raise capi.GfgpuError("no expression")
        if order == 2:
            # every order-2 tree adds into ONE matrix (workspace.cc:791-936): accumulated on the device
            n = self._term(0).ndof
            if self._Kdev is None:
                self._Kdev = capi.DeviceMatrix(self.ctx, n)
            self._Kdev.clear(keep_pattern=False)
        else:
            vec = None
        for k in range(len(self.terms)):
            if order == 2 and self.terms[k][0] in ("source", "nsource"):
                continue  # an order-1 term has no order-2 tree (workspace.cc:545-600)
            dev = self._term(k)
            mf, V = self.variables[self.terms[k][1]]
            U = None if V is None else np.ascontiguousarray(V, np.float64)
            if order == 2:
                dev.assemble_host(U, capi.TANGENT, None, None)
                self._Kdev.add_term(dev, 1.0, 0, 0)
            else:
                R = np.empty(dev.ndof)
                dev.assemble_host(U, capi.RESIDUAL, None, R)
                vec = R if vec is None else vec + R
        if order == 2:
            self._K = self._Kdev.export_csc()
        else:
            self._R = vec

    def assembled_matrix_device(self):
        """The device-resident tangent of the last assembly(2) (capi.DeviceMatrix): csc_view(), mult(), ..."""
        return self._Kdev

    def assembled_matrix(self):
        """(jc, ir, pr): the layout of gmm::csc_matrix::init_with(K) (gmm_matrix.h:545-566)."""
        return self._K

    def assembled_vector(self):
        return self._R
