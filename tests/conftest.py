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


def golden_names(oracle_only=False, mirror=False):
    names = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "*.npz")))
    # "o_*": the laws the oracle restated and pinned first (round 1); they are device families since round 2
    # "p4_*": P4 simplices reach the device with the REFERENCE's tables (C ABI, C++ shim); the Python mirror has no
    # degree-8 simplex cubature of its own, so the tests that regenerate tables in Python skip them
    if mirror:
        names = [n for n in names if not n.startswith("p4_")]
    return names


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
        d["fparams"] = (np.array([a]) if d["family"] in ("laplace", "mass") else
                        d["lawparams"].astype(np.float64) if "lawparams" in d else np.array([lam, mu]))
    # fem-data coefficients (coef=fem): the fields replace the leading parameters; the data fem's basis table covers ALL
    # integration points, so the all-point tables are used with it
    d["fields"] = None
    if args.get("coef") == "fem":
        vals = [d["d_vals0"].astype(np.float64)] + ([d["d_vals1"].astype(np.float64)] if "d_vals1" in d else [])
        if d["family"] == "source":
            vals = [-vals[0]]  # "-f.Test_u": the family integrates F = -f
        d["fields"] = {"d_elem_dof": d["d_elem_dof"], "d_phi": d["d_phi"], "vals": vals, "kd": int(args.get("kd", 1))}
    # further expressions of the same workspace (family2.., make_golden "m_*"): (family, parameters, region, region name)
    d["extra_terms"] = []
    for t in (2, 3, 4):
        ft = args.get("family%d" % t)
        if not ft:
            continue
        at = float(args.get("a%d" % t, 1.0))
        if ft == "source":
            fp = -at * np.arange(1, d["Q"] + 1, dtype=np.float64)
        elif ft == "nsource":
            fp = d["gdata%d" % t].astype(np.float64)
        elif ft == "elast":
            fp = np.array([float(args.get("lambda%d" % t, 1.0)), float(args.get("mu%d" % t, 1.0))])
        else:
            fp = np.array([at])
        rg = None
        if "items%d_cv" % t in d:
            rg = {"items_cv": d["items%d_cv" % t], "items_f": d["items%d_f" % t]}
            if (rg["items_f"] >= 0).any():
                rg.update(face_first=d["face_first"], face_nq=d["face_nq"], ref_normals=d["ref_normals"])
        d["extra_terms"].append((ft, fp, rg, args.get("region%d" % t, "all")))
    # mesh regions: items in mr_visitor order; with faces the tables cover ALL integration points
    d["region"] = None
    d["tables"] = (d["quad_w"], d["gt_grad"], d["phi"], d["gphi"])
    if "items_cv" in d:
        d["region"] = {"items_cv": d["items_cv"], "items_f": d["items_f"]}
        if (d["items_f"] >= 0).any():
            d["region"].update(face_first=d["face_first"], face_nq=d["face_nq"], ref_normals=d["ref_normals"])
            d["tables"] = (d["all_w"], d["all_gt_grad"], d["all_phi"], d["all_gphi"])
    if d["fields"] is not None or any(rg is not None and "face_first" in rg for _, _, rg, _ in d["extra_terms"]):
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


def csc_sum(mats, n):
    """Sum of CSC matrices (jc, ir, pr) with the UNION pattern: stored entries stay stored even when they are or become
    exactly zero (what adding into gmm::col_matrix<rsvector> does; scipy's + would drop them).  Terms are added in order."""
    keys = np.concatenate([np.repeat(np.arange(n, dtype=np.int64), np.diff(jc)) * n + np.asarray(ir, np.int64)
                           for jc, ir, _ in mats])
    vals = np.concatenate([pr for _, _, pr in mats])
    uk, inv = np.unique(keys, return_inverse=True)
    out = np.zeros(len(uk))
    np.add.at(out, inv, vals)
    cols = uk // n
    jc = np.zeros(n + 1, np.int64)
    np.add.at(jc, cols + 1, 1)
    return np.cumsum(jc), (uk % n).astype(np.int64), out
