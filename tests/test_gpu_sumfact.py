"""GPU: the sum-factorised Laplace kernel for Q3/Q4 hexahedra (csrc/sumfact.cu) on DISTORTED meshes (a different
Jacobian at every Gauss point), against the generic element kernel (GFGPU_NO_SUMFACT=1) and against the CPU oracle.
Pattern identical, values and residual 1e-12; the reference-generated golden c5 (Q4) is covered by
test_gpu_golden.py / test_gpu_workspace.py (its tables carry the reference's 5e-10 round-off, are not an exact tensor
product, and therefore take the generic kernel: the factorisation check is part of the contract)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [  # k, im, nsub
    (4, 8, [2, 2, 3]),
    (4, 8, [1, 1, 1]),
    (3, 6, [2, 3, 2]),
]


def _setup(k, im, nsub, distort=0.15, Q=1, uscale=1.0):
    import getfem_b200 as gf
    from getfem_b200 import capi, fem_tables
    ctx = capi.Context(0)
    m = gf.mesh()
    gf.regular_unit_mesh(m, nsub, "GT_QK(3,1)")
    rng = np.random.default_rng(3)
    h = 1.0 / max(nsub)
    m.pts = m.pts + distort * h * rng.uniform(-1, 1, m.pts.shape)  # every hexahedron becomes genuinely trilinear
    m._dev = {}
    mf = gf.mesh_fem(m, Q)
    mf.set_classical_finite_element(k)
    dmesh, dfem = m.device(ctx), mf.device(ctx)
    t = fem_tables.classical_tables("QK", 3, k, im)
    tab = capi.DeviceTables(ctx, t["quad_w"], t["gt_grad"], t["phi"], t["gphi"])
    U = uscale * rng.uniform(-1, 1, dfem.ndof)
    return ctx, m, mf, dmesh, dfem, t, tab, U


def _assemble(ctx, dmesh, dfem, tab, U, strategy=0, family="laplace", params=(1.7,), alpha=0.5):
    from getfem_b200 import capi
    term = capi.DeviceTerm(ctx, dmesh, dfem, tab, family, list(params), alpha, strategy)
    R = np.empty(dfem.ndof)
    term.assemble_host(U, capi.TANGENT | capi.RESIDUAL, None, R)
    jc, ir, pr = term.export_csc()
    return jc, ir, pr, R, term.last_timings()


@pytest.mark.parametrize("case", CASES, ids=lambda c: "q%d-im%d-%s" % (c[0], c[1], "x".join(map(str, c[2]))))
def test_sumfact_matches_generic_kernel_and_oracle(case):
    from oracle import oracle
    k, im, nsub = case
    ctx, m, mf, dmesh, dfem, t, tab, U = _setup(k, im, nsub)
    os.environ.pop("GFGPU_NO_SUMFACT", None)
    jc, ir, pr, R, _ = _assemble(ctx, dmesh, dfem, tab, U)
    os.environ["GFGPU_NO_SUMFACT"] = "1"
    try:
        gjc, gir, gpr, gR, _ = _assemble(ctx, dmesh, dfem, tab, U)
    finally:
        os.environ.pop("GFGPU_NO_SUMFACT", None)
    assert np.array_equal(jc, gjc) and np.array_equal(ir, gir), "pattern differs from the generic kernel"
    assert np.linalg.norm(pr - gpr) <= 1e-12 * np.linalg.norm(gpr)
    assert np.linalg.norm(R - gR) <= 1e-12 * np.linalg.norm(gR)
    ed = dfem.elem_dof()
    ojc, oir, opr, oR = oracle.assemble(m.pts, m.conn, ed, dfem.ndof, 1, t["quad_w"], t["gt_grad"], t["phi"], t["gphi"],
                                        False, "laplace", [1.7], U)
    opr, oR = 0.5 * opr, 0.5 * oR  # the term's factor alpha = 0.5 (factor_of_variable)
    assert np.array_equal(jc, ojc) and np.array_equal(ir, oir), "pattern differs from the oracle"
    assert np.linalg.norm(pr - opr) <= 1e-12 * np.linalg.norm(opr)
    assert np.linalg.norm(R - oR) <= 1e-12 * np.linalg.norm(oR)


def test_sumfact_is_the_kernel_that_runs_and_is_faster():
    """The factorised kernel must actually be selected for exact tensor-product tables (no silent generic path)."""
    ctx, m, mf, dmesh, dfem, t, tab, U = _setup(4, 8, [4, 4, 4])
    os.environ.pop("GFGPU_NO_SUMFACT", None)
    _assemble(ctx, dmesh, dfem, tab, U)
    *_, tm = _assemble(ctx, dmesh, dfem, tab, U)
    os.environ["GFGPU_NO_SUMFACT"] = "1"
    try:
        *_, tg = _assemble(ctx, dmesh, dfem, tab, U)
    finally:
        os.environ.pop("GFGPU_NO_SUMFACT", None)
    assert tm["elem"] > 0 and tg["elem"] > 3 * tm["elem"], (tm, tg)


def test_noisy_tables_fall_back_to_the_generic_kernel():
    """Tables that are not a tensor product to 1e-13 (e.g. the reference's own Q4 tables) must not be factorised."""
    from getfem_b200 import capi
    ctx, m, mf, dmesh, dfem, t, tab, U = _setup(4, 8, [1, 1, 2])
    rng = np.random.default_rng(5)
    noisy = capi.DeviceTables(ctx, t["quad_w"], t["gt_grad"], t["phi"] * (1 + 1e-9 * rng.standard_normal(t["phi"].shape)),
                              t["gphi"])
    jc, ir, pr, R, _ = _assemble(ctx, dmesh, dfem, noisy, U)
    os.environ["GFGPU_NO_SUMFACT"] = "1"
    try:
        gjc, gir, gpr, gR, _ = _assemble(ctx, dmesh, dfem, noisy, U)
    finally:
        os.environ.pop("GFGPU_NO_SUMFACT", None)
    assert np.array_equal(pr, gpr) and np.array_equal(R, gR), "noisy tables must take the generic kernel (bitwise equal)"


HYPER = [  # family, im, nsub
    ("nh_ciarlet", 6, [2, 2, 2]),
    ("nh_bonet", 6, [2, 1, 2]),
    ("svk", 6, [2, 2, 1]),
    ("svk", 4, [1, 2, 2]),
]


@pytest.mark.parametrize("case", HYPER, ids=lambda c: "%s-im%d-%s" % (c[0], c[1], "x".join(map(str, c[2]))))
def test_hyperelastic_sumfact_matches_generic_kernel_and_oracle(case):
    """Q2 hexahedra, finite-strain laws (BASELINE config 4): factorised tangent + residual on a distorted mesh."""
    from oracle import oracle
    family, im, nsub = case
    ctx, m, mf, dmesh, dfem, t, tab, U = _setup(2, im, nsub, Q=3, uscale=0.03)
    par = (1.3, 0.8)
    os.environ.pop("GFGPU_NO_SUMFACT", None)
    jc, ir, pr, R, tm = _assemble(ctx, dmesh, dfem, tab, U, family=family, params=par, alpha=1.0)
    os.environ["GFGPU_NO_SUMFACT"] = "1"
    try:
        gjc, gir, gpr, gR, tg = _assemble(ctx, dmesh, dfem, tab, U, family=family, params=par, alpha=1.0)
    finally:
        os.environ.pop("GFGPU_NO_SUMFACT", None)
    assert np.array_equal(jc, gjc) and np.array_equal(ir, gir), "pattern differs from the generic kernel"
    assert np.linalg.norm(pr - gpr) <= 1e-12 * np.linalg.norm(gpr)
    assert np.linalg.norm(R - gR) <= 1e-12 * np.linalg.norm(gR)
    ed = dfem.elem_dof()
    ojc, oir, opr, oR = oracle.assemble(m.pts, m.conn, ed, dfem.ndof, 3, t["quad_w"], t["gt_grad"], t["phi"], t["gphi"],
                                        False, family, list(par), U)
    assert np.array_equal(jc, ojc) and np.array_equal(ir, oir), "pattern differs from the oracle"
    assert np.linalg.norm(pr - opr) <= 1e-12 * np.linalg.norm(opr)
    assert np.linalg.norm(R - oR) <= 1e-12 * np.linalg.norm(oR)


@pytest.mark.parametrize("case", [(4, 8, [2, 2, 3]), (3, 6, [2, 3, 2])], ids=["q4", "q3"])
def test_direct_mode_repeats_the_staged_pass_bit_for_bit(case):
    """From the second assembly on (fixed pattern) the scalar sum-factorised kernel writes its entries straight to their CSC
    slots and only the shared entries go through a compact stage (kernel_kind 4): same bits as the staged first pass, also
    after the state changes, and the same as with the direct mode switched off."""
    from getfem_b200 import capi
    k, im, nsub = case
    ctx, m, mf, dmesh, dfem, t, tab, U = _setup(k, im, nsub)
    os.environ.pop("GFGPU_NO_SUMFACT", None)
    order = capi.TANGENT | capi.RESIDUAL

    def passes(direct):
        if direct:
            os.environ.pop("GFGPU_NO_DIRECT", None)
        else:
            os.environ["GFGPU_NO_DIRECT"] = "1"
        try:
            term = capi.DeviceTerm(ctx, dmesh, dfem, tab, "laplace", [1.7], 0.5, 0)
            out = []
            for Uk in (U, U, 0.3 * U[::-1].copy(), None):
                R = np.empty(dfem.ndof)
                term.assemble_host(Uk, order, None, R)
                jc, ir, pr = term.export_csc()
                out.append((jc, ir, pr, R, term.kernel_kind))
            term.assemble_host(U, capi.TANGENT, None, None)  # tangent alone
            out.append(term.export_csc() + (None, term.kernel_kind))
            return out
        finally:
            os.environ.pop("GFGPU_NO_DIRECT", None)

    d, s = passes(True), passes(False)
    assert [o[4] for o in d] == [0, 4, 4, 4, 4] and [o[4] for o in s] == [0, 0, 0, 0, 0]
    for a, b in zip(d, s):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        assert a[3] is None or np.array_equal(a[3], b[3])
    assert np.array_equal(d[0][2], d[1][2]) and np.array_equal(d[0][3], d[1][3])
    assert not d[3][3].any()  # zero state: zero residual


def test_direct_mode_follows_a_new_region():
    """changing the element range invalidates the pattern: one staged pass, then direct again with the new slots"""
    from getfem_b200 import capi
    ctx, m, mf, dmesh, dfem, t, tab, U = _setup(4, 8, [2, 2, 3])
    term = capi.DeviceTerm(ctx, dmesh, dfem, tab, "laplace", [1.0], 1.0, 0)
    ref = capi.DeviceTerm(ctx, dmesh, dfem, tab, "laplace", [1.0], 1.0, 0)
    ne = m.nb_convex()
    for e0, e1 in ((0, ne), (2, ne - 3), (0, 5)):
        term.set_element_range(e0, e1)
        for _ in range(3):
            term.assemble_host(U, capi.TANGENT, None, None)
        assert term.kernel_kind == 4
        os.environ["GFGPU_NO_DIRECT"] = "1"
        try:
            ref.set_element_range(e0, e1)
            ref.assemble_host(U, capi.TANGENT, None, None)
        finally:
            os.environ.pop("GFGPU_NO_DIRECT", None)
        for x, y in zip(term.export_csc(), ref.export_csc()):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_kernel_variants_of_the_last_contraction_agree(variant):
    """GFGPU_SF_VARIANT: 1 = FMA pipe with one 15-warp CTA per SM, 2 / 3 = the last contraction on the fp64 tensor core
    (mma.sync.m8n8k4.f64, K 45 -> 48, tiles 25 -> 32).  Same pattern, values to 1e-13 of the default kernel."""
    ctx, m, mf, dmesh, dfem, t, tab, U = _setup(4, 8, [2, 2, 3])
    os.environ.pop("GFGPU_NO_SUMFACT", None)
    os.environ.pop("GFGPU_SF_VARIANT", None)
    jc, ir, pr, R, _ = _assemble(ctx, dmesh, dfem, tab, U)
    os.environ["GFGPU_SF_VARIANT"] = str(variant)
    try:
        vjc, vir, vpr, vR, _ = _assemble(ctx, dmesh, dfem, tab, U)
    finally:
        os.environ.pop("GFGPU_SF_VARIANT", None)
    assert np.array_equal(jc, vjc) and np.array_equal(ir, vir)
    assert np.linalg.norm(pr - vpr) <= 1e-13 * np.linalg.norm(pr)
    assert np.linalg.norm(R - vR) <= 1e-13 * np.linalg.norm(R)
