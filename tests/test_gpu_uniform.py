"""GPU: the class-uniform tile kernel (csrc/recompute_uniform.cu) against the reference goldens, against the general tile
kernel on the same term, and its plan switches (sub-ranges of a column, image ring, task cutting).

GFGPU_UNIFORM=2 forces the kernel whatever the class sizes (the goldens are tiny meshes: most classes have a handful of
members, i.e. tiles with few active lanes); the default (1) engages it only when most columns have translated copies."""
import os

import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-12


class env:
    def __init__(self, **kw):
        self.kw = {k: str(v) for k, v in kw.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update(self.kw)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _recompute_goldens():
    out = []
    for n in golden_names():
        g = load_golden(n)
        faces = g["region"] is not None and "face_first" in g["region"]
        if (g["gt_linear"] and g["family"] in ("laplace", "elast", "mass") and not faces and g["fields"] is None and not g["extra_terms"]
                and int(g["args"]["k"]) <= 3):  # the per-nonzero kernels are instantiated up to P3
            out.append(n)
    return out


@pytest.mark.parametrize("name", _recompute_goldens())
def test_uniform_kernel_matches_reference_golden(name):
    from test_gpu_golden import device_assemble
    g = load_golden(name)
    with env(GFGPU_UNIFORM=2, GFGPU_COLS=0):
        jc, ir, pr, R, term = device_assemble(g, strategy=2)
        assert term.kernel_kind == 3
    assert np.array_equal(jc, g["K_jc"]) and np.array_equal(ir, g["K_ir"])
    rel = np.linalg.norm(pr - g["K_pr"]) / max(np.linalg.norm(g["K_pr"]), 1e-300)
    assert rel < TOL, rel
    relr = np.linalg.norm(R - g["R"]) / max(np.linalg.norm(g["R"]), 1e-300)
    assert relr < TOL, relr


@pytest.mark.parametrize("img,nbuf,taskcap", [(8192, 2, 16), (12288, 3, 4), (40960, 1, 1), (16384, 8, 64)])
def test_uniform_plan_switches(img, nbuf, taskcap):
    """small image buffers cut the columns into sub-ranges (three CSC pieces per lane), task caps cut the pair groups"""
    from test_gpu_golden import device_assemble
    g = load_golden("c3_elast3d_p2_n2")
    with env(GFGPU_UNIFORM=2, GFGPU_UT_IMG=img, GFGPU_UT_NBUF=nbuf, GFGPU_UT_TASKCAP=taskcap):
        jc, ir, pr, R, term = device_assemble(g, strategy=2)
        assert term.kernel_kind == 3
    assert np.array_equal(jc, g["K_jc"]) and np.array_equal(ir, g["K_ir"])
    assert np.linalg.norm(pr - g["K_pr"]) / np.linalg.norm(g["K_pr"]) < TOL


def _mirror_term(dim, n, k, Q, family, params, im):
    from test_gpu_workspace import build_ws
    ws, mf, m, _ = build_ws(dim, [n] * dim, "PK", k, Q, im, family, params)
    ws.assembly(2)
    dev = ws.terms[0][4]
    jc, ir, pr = dev.export_csc()
    return jc, ir, pr, dev, ws


@pytest.mark.parametrize("dim,n,k,Q,family,params,im", [
    (3, 9, 2, 3, "elast", [1.3, 0.7], 4),
    (3, 8, 2, 1, "laplace", [2.0], 4),
    (2, 24, 2, 2, "elast", [1.0, 1.0], 4),
    (3, 6, 2, 3, "mass", [1.5], 4),
    (3, 7, 1, 1, "laplace", [1.0], 2),
    (2, 16, 1, 2, "laplace", [0.5], 2),
])
def test_uniform_equals_general_tile_kernel(dim, n, k, Q, family, params, im):
    """same term through both per-nonzero kernels: identical pattern, values equal to round-off (same summation order)"""
    with env(GFGPU_UNIFORM=0, GFGPU_COLS=0):
        jc0, ir0, pr0, d0, ws0 = _mirror_term(dim, n, k, Q, family, params, im)
        assert d0.kernel_kind == 1
    with env(GFGPU_UNIFORM=2, GFGPU_COLS=0):
        jc1, ir1, pr1, d1, ws1 = _mirror_term(dim, n, k, Q, family, params, im)
        assert d1.kernel_kind == 3
    assert np.array_equal(jc0, jc1) and np.array_equal(ir0, ir1)
    assert np.linalg.norm(pr1 - pr0) <= 1e-14 * np.linalg.norm(pr0)
    # run-to-run bitwise determinism of the uniform kernel (task order is dynamic, the sums are not)
    with env(GFGPU_UNIFORM=2, GFGPU_COLS=0):
        ws1.assembly(2)
        pr2 = ws1.terms[0][4].export_csc()[2]
    assert np.array_equal(pr1, pr2)


def test_uniform_engages_on_a_regular_mesh_when_asked():
    """GFGPU_UNIFORM=1, n = 20: most columns sit in classes of >= 16 translated copies, the plan takes the uniform kernel;
    on the tiny goldens the same setting keeps the general kernel (coverage rule)"""
    with env(GFGPU_UNIFORM=1):
        jc, ir, pr, dev, ws = _mirror_term(3, 20, 2, 3, "elast", [1.0, 1.0], 4)
        assert dev.kernel_kind == 3
        assert _mirror_term(3, 3, 2, 3, "elast", [1.0, 1.0], 4)[3].kernel_kind == 1
    with env(GFGPU_UNIFORM=0):
        jc0, ir0, pr0, d0, _ = _mirror_term(3, 20, 2, 3, "elast", [1.0, 1.0], 4)
        assert d0.kernel_kind == 1
    assert np.array_equal(jc0, jc) and np.array_equal(ir0, ir)
    assert np.linalg.norm(pr - pr0) <= 1e-14 * np.linalg.norm(pr0)
    # size-independent properties: symmetry and K * (rigid translation) = 0
    import scipy.sparse as sp
    K = sp.csc_matrix((pr, ir, jc), shape=(len(jc) - 1,) * 2)
    assert abs(K - K.T).max() <= 1e-12 * abs(K).max()
    t = np.tile([1.0, -2.0, 0.5], (len(jc) - 1) // 3)
    assert np.linalg.norm(K @ t) <= 1e-10 * np.linalg.norm(pr)
