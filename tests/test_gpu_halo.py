"""GPU: the multi-GPU halo path of the C ABI (include/gfgpu.h "multi-GPU") on ONE device: several terms play the
ranks (disjoint element blocks in rank order), getfem_b200.halo.setup_local / exchange_local move the slices.
After the exchange every rank's owned column slab and residual slice must equal the single-term assembly:
pattern identical, values 1e-12."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [  # dim, nsub, gt, k, Q, im, family, params, strategy, nranks
    (3, [4, 3, 5], "PK", 2, 3, 4, "elast", [1.3, 0.7], 0, 2),
    (3, [4, 3, 5], "PK", 2, 3, 4, "elast", [1.3, 0.7], 0, 3),
    (3, [3, 3, 4], "PK", 1, 1, 2, "laplace", [2.0], 0, 2),
    (2, [7, 9], "PK", 2, 2, 4, "elast", [1.0, 2.0], 0, 3),
    (3, [4, 3, 5], "PK", 2, 3, 4, "elast", [1.3, 0.7], 1, 2),          # STAGED strategy
    (3, [2, 2, 4], "QK", 2, 3, 6, "svk", [1.0, 1.0], 0, 2),            # hexahedra, nonlinear family
]


def _setup(dim, nsub, gt, k, Q, im, family, params, strategy):
    import torch
    import getfem_b200 as gf
    from getfem_b200 import capi, fem_tables
    ctx = capi.Context(0)
    m = gf.mesh()
    gf.regular_unit_mesh(m, nsub, "GT_%s(%d,1)" % (gt, dim))
    mf = gf.mesh_fem(m, Q)
    mf.set_classical_finite_element(k)
    dmesh, dfem = m.device(ctx), mf.device(ctx)
    t = fem_tables.classical_tables(gt, dim, k, im)
    tab = capi.DeviceTables(ctx, t["quad_w"], t["gt_grad"], t["phi"], t["gphi"])
    ndof = dfem.ndof
    rng = np.random.default_rng(7)
    U = rng.uniform(-1, 1, ndof) * (0.02 if family in ("svk", "nh_ciarlet") else 1.0)
    U_dev = torch.from_numpy(U).cuda()
    mk = lambda: capi.DeviceTerm(ctx, dmesh, dfem, tab, family, params, 1.0, strategy)
    return ctx, m, mk, U_dev, ndof, (dmesh, dfem, tab)


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%s%d-q%d-s%d-r%d" % (c[6], c[2], c[3], c[4], c[8], c[9]))
def test_owned_slabs_match_single_term(case):
    import torch
    from getfem_b200 import capi, halo
    dim, nsub, gt, k, Q, im, family, params, strategy, nr = case
    ctx, m, mk, U_dev, ndof, keep = _setup(dim, nsub, gt, k, Q, im, family, params, strategy)
    ORDER = capi.TANGENT | capi.RESIDUAL
    full = mk()
    full.assemble_dev(U_dev.data_ptr(), ORDER)
    jc, ir, pr = full.export_csc()
    R = full.export_residual()
    ne = m.nb_convex()
    cuts = [round(r * ne / nr) for r in range(nr + 1)]
    terms = []
    for r in range(nr):
        t = mk()
        t.set_element_range(cuts[r], cuts[r + 1])
        terms.append(t)
    plans = halo.setup_local(terms, U_dev.data_ptr())
    assert plans[0].D[0] == 0 and plans[-1].D[-1] == ndof
    assert any(p.sends for p in plans[1:]), "the blocks share no dof: the case does not exercise the exchange"
    for rep in range(2):  # the second pass reuses pattern, plan and maps
        for t in terms:
            t.assemble_dev(U_dev.data_ptr(), ORDER)
        ctx.synchronize()
        halo.exchange_local(terms, plans, ORDER)
        torch.cuda.synchronize()
        for r, t in enumerate(terms):
            lo, hi = t.owned_range()
            assert (lo, hi) == plans[r].own
            if hi == lo:
                continue
            tjc, tir, tpr = t.export_csc()
            a, b = tjc[lo], tjc[hi]
            A, B = jc[lo], jc[hi]
            assert np.array_equal(tjc[lo:hi + 1] - a, jc[lo:hi + 1] - A), "column counts of the owned slab differ"
            assert np.array_equal(tir[a:b], ir[A:B]), "row indices of the owned slab differ"
            ref = pr[A:B]
            assert np.linalg.norm(tpr[a:b] - ref) <= 1e-12 * np.linalg.norm(ref)
            tR = t.export_residual()
            assert np.linalg.norm(tR[lo:hi] - R[lo:hi]) <= 1e-12 * max(np.linalg.norm(R[lo:hi]), 1e-300)


def test_exchange_is_deterministic():
    import torch
    from getfem_b200 import capi, halo
    ctx, m, mk, U_dev, ndof, keep = _setup(3, [4, 4, 4], "PK", 2, 3, 4, "elast", [1.0, 1.0], 0)
    ne = m.nb_convex()
    outs = []
    for _ in range(2):
        terms = []
        for r in range(2):
            t = mk()
            t.set_element_range(r * ne // 2, (r + 1) * ne // 2)
            terms.append(t)
        plans = halo.setup_local(terms, U_dev.data_ptr())
        for t in terms:
            t.assemble_dev(U_dev.data_ptr(), capi.TANGENT | capi.RESIDUAL)
        ctx.synchronize()
        halo.exchange_local(terms, plans, capi.TANGENT | capi.RESIDUAL)
        torch.cuda.synchronize()
        outs.append([(t.export_csc()[2], t.export_residual()) for t in terms])
    for (p0, r0), (p1, r1) in zip(*outs):
        assert np.array_equal(p0, p1) and np.array_equal(r0, r1)


def test_halo_with_the_uniform_tile_kernel(monkeypatch):
    """the class-uniform kernel under a halo: ghost columns are computed like owned ones, virtual pairs write zeros"""
    monkeypatch.setenv("GFGPU_UNIFORM", "2")
    monkeypatch.setenv("GFGPU_COLS", "0")
    test_owned_slabs_match_single_term((3, [4, 3, 6], "PK", 2, 3, 4, "elast", [1.3, 0.7], 0, 3))
    test_owned_slabs_match_single_term((3, [3, 3, 4], "PK", 1, 1, 2, "laplace", [2.0], 0, 2))


def test_halo_refuses_load_terms_loudly():
    """a source term has no tangent to announce; its interface contributions used to be dropped silently"""
    from getfem_b200 import capi
    ctx, m, mk, U_dev, ndof, keep = _setup(3, [3, 3, 4], "PK", 2, 3, 4, "elast", [1.0, 1.0], 0)
    dmesh, dfem, tab = keep
    t = capi.DeviceTerm(ctx, dmesh, dfem, tab, "source", [1.0, 2.0, 3.0], 1.0, 0)
    t.set_element_range(0, m.nb_convex() // 2)
    with pytest.raises(capi.GfgpuError, match="order-1-only"):
        t.halo_begin(U_dev.data_ptr())
