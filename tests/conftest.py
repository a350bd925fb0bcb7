import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(bytes(d["meta"]).decode())
    args = dict(kv.split("=") for kv in d["meta"]["driver_args"].split())
    d["args"] = args
    fam = args["family"]
    d["family"] = "laplace" if fam == "laplace_vec" else fam
    d["Q"] = int(args["q"])
    d["gt_linear"] = args["gt"] == "pk"
    lam, mu, a = d["params"]
    if d["family"] == "source":  # "-f.Test_u" with f = a * (1..Q): the family integrates F.Test_u, F = -f
        d["fparams"] = -a * np.arange(1, d["Q"] + 1, dtype=np.float64)
    elif d["family"] == "nsource":  # "(Reshape(g,qdim,meshdim)*Normal).Test_u": A(b,n) = g[b + Q*n]
        d["fparams"] = d["gdata"].astype(np.float64)
    else:
        d["fparams"] = np.array([a]) if d["family"] in ("laplace", "mass") else np.array([lam, mu])
    # mesh regions: items in mr_visitor order; with faces the tables cover ALL integration points
    d["region"] = None
    d["tables"] = (d["quad_w"], d["gt_grad"], d["phi"], d["gphi"])
    if "items_cv" in d:
        d["region"] = {"items_cv": d["items_cv"], "items_f": d["items_f"]}
        if (d["items_f"] >= 0).any():
            d["region"].update(face_first=d["face_first"], face_nq=d["face_nq"], ref_normals=d["ref_normals"])
            d["tables"] = (d["all_w"], d["all_gt_grad"], d["all_phi"], d["all_gphi"])
    return d


def csc_to_scipy(jc, ir, pr, n):
    import scipy.sparse as sp
    return sp.csc_matrix((pr, ir, jc), shape=(n, n))


def make_region(m, name):
    """The regions of oracle/ref_driver.cc (region=outer|xmax|zmin|half) on the Python mirror of the mesh."""
    import getfem_b200 as gf
    if name in (None, "all"):
        return None
    rg = gf.mesh_region()
    if name == "half":  # convexes whose barycentre has x < 0.5
        for cv in np.nonzero(m.pts[m.conn][:, :, 0].mean(1) < 0.5)[0]:
            rg.add(cv)
        return rg
    cvs, fcs = gf.outer_faces_of_mesh(m).items()
    for cv, f in zip(cvs, fcs):
        P = m.points_of_face_of_convex(cv, f)
        if name == "outer" or (name == "xmax" and (abs(P[:, 0] - 1.0) < 1e-12).all()) or \
                (name == "zmin" and (abs(P[:, -1]) < 1e-12).all()):
            rg.add(cv, f)
    return rg
