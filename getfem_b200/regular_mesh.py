"""Regular unit meshes with the point/convex numbering of getfem::regular_unit_mesh.

Restates (vectorised, numpy) what the reference does in
  src/getfem_regular_meshes.cc:55-77   (parallelepiped_regular_simplex_mesh_: cells in x-fastest
                                        odometer order, 2^N corners added per cell, simplices from the
                                        simplexification table, parity flip of the table when N != 3)
  src/getfem_regular_meshes.cc:121-137 (parallelepiped_regular_mesh_: one GT_QK(N,1) convex per cell)
  src/getfem_regular_meshes.cc:237-284 (regular_unit_mesh: the auxiliary mesh is re-added convex by
                                        convex with add_convex_by_points, so the final point ids are in
                                        FIRST-USE order of the (convex, local node) walk)
Simplexification of the reference square / cube: bgeot_convex_ref_simplexified.cc:28-39
(2 triangles {3,0,2},{3,0,1}; 6 Kuhn tetrahedra).  Checked bit-for-bit (ids and coordinates) against
meshes dumped from the reference (tests/golden).
"""
import numpy as np

# corner index = ix + 2*iy + 4*iz of the reference parallelepiped
_SIMPLEX_TABLE = {
    2: np.array([[3, 0, 2], [3, 0, 1]], np.int64),
    3: np.array([[3, 7, 0, 1], [7, 0, 5, 4], [7, 0, 1, 5], [3, 7, 0, 2], [6, 7, 0, 4], [6, 7, 0, 2]], np.int64),
}


def regular_unit_mesh(nsubdiv, kind):
    """nsubdiv: cells per direction (len 2 or 3); kind: 'simplex' (GT_PK(N,1)) or 'parallelepiped'
    (GT_QK(N,1)).  Returns (pts float64 [npts, N], conn int32 [ne, ng]) in the reference numbering."""
    ns = [int(v) for v in nsubdiv]
    N = len(ns)
    assert N in (2, 3) and all(v >= 1 for v in ns)
    ncell = int(np.prod(ns))
    cell = np.arange(ncell, dtype=np.int64)
    tab = np.empty((ncell, N), np.int64)  # odometer, first direction fastest
    r = cell.copy()
    for d in range(N):
        tab[:, d] = r % ns[d]
        r //= ns[d]
    nbpt = 1 << N
    corner_bits = np.array([[(c >> d) & 1 for d in range(N)] for c in range(nbpt)], np.int64)  # [nbpt, N]
    stride = np.ones(N, np.int64)
    for d in range(1, N):
        stride[d] = stride[d - 1] * (ns[d - 1] + 1)
    # lattice id of each corner of each cell
    lat = ((tab[:, None, :] + corner_bits[None, :, :]) * stride[None, None, :]).sum(-1)  # [ncell, nbpt]
    if kind == "simplex":
        tabl = _SIMPLEX_TABLE[N]
        nbs = tabl.shape[0]
        idx = np.broadcast_to(tabl[None, :, :], (ncell, nbs, N + 1)).copy()
        if N != 3:  # parity flip of the simplexification (getfem_regular_meshes.cc:66-67)
            odd = (tab.sum(1) & 1).astype(bool)
            idx[odd] = (idx[odd] + nbpt // 2) % nbpt
        conn_lat = np.take_along_axis(lat[:, None, :].repeat(nbs, 1), idx, axis=2).reshape(ncell * nbs, N + 1)
    elif kind == "parallelepiped":
        conn_lat = lat
    else:
        raise ValueError("kind must be 'simplex' or 'parallelepiped'")
    # first-use renumbering of the points (add_convex_by_points walk)
    flat = conn_lat.reshape(-1)
    uniq, first = np.unique(flat, return_index=True)
    order = np.argsort(first, kind="stable")
    newid = np.empty(uniq.shape[0], np.int64)
    newid[order] = np.arange(uniq.shape[0])
    conn = newid[np.searchsorted(uniq, flat)].reshape(conn_lat.shape).astype(np.int32)
    # coordinates as the reference computes them: fl(fl(h*tab) + h*bit) in the FIRST cell (odometer
    # order) touching the lattice point, i.e. tab = max(X-1, 0)
    lat_sorted = uniq[order]
    pts = np.empty((uniq.shape[0], N), np.float64)
    r = lat_sorted.copy()
    for d in range(N):
        X = r % (ns[d] + 1)
        r //= (ns[d] + 1)
        h = 1.0 / float(ns[d]) * 1.0
        t = np.maximum(X - 1, 0)
        pts[:, d] = (h * t.astype(np.float64)) + h * (X - t).astype(np.float64)
    return pts, conn
