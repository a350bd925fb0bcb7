"""CPU: the closed forms behind the law operators of the NVRTC route (jit.cu: svk_pk2 / svk_dpk2, nh_pk2 / nh_dpk2 on N x N tensors)
restated in numpy and checked against the oracle's material point (asm_oracle.c gfo_hyper_law, itself pinned on the reference's
Saint-Venant Kirchhoff and Neo_Hookean_hyperelastic_law, getfem_nonlinear_elasticity.cc:503-540, 612-702): the 3D value and
directional derivative, and the PLANE STRAIN wrapper (plane_strain_hyperelastic_law, :906-945) -- the 2 x 2 formulas against the
2 x 2 block of the 3D law on the embedded gradient.  The device helpers themselves are compared with the reference in one
process by tests/test_gpu_dropin.py (model=finite_strain dim=2, the compound forms of JIT_VECTOR)."""
import numpy as np
import pytest


def nh_pk2(g, lam, mu, bonet):
    n = g.shape[0]
    f = g + np.eye(n)
    c = f.T @ f
    i3 = np.linalg.det(c)
    k = 0.5 * lam * np.log(i3) - mu if bonet else 0.5 * lam * (i3 - 1.0) - mu
    return k * np.linalg.inv(c) + mu * np.eye(n)


def nh_dpk2(g, lam, mu, bonet, h):
    n = g.shape[0]
    f = g + np.eye(n)
    c = f.T @ f
    ci = np.linalg.inv(c)
    dc = h.T @ f + f.T @ h
    i3 = np.linalg.det(c)
    k = 0.5 * lam * np.log(i3) - mu if bonet else 0.5 * lam * (i3 - 1.0) - mu
    di3 = i3 * np.trace(ci @ dc)
    dk = 0.5 * lam * di3 / i3 if bonet else 0.5 * lam * di3
    return dk * ci - k * (ci @ dc @ ci)


def svk_pk2(g, lam, mu):
    e = 0.5 * (g + g.T + g.T @ g)
    return lam * np.trace(e) * np.eye(g.shape[0]) + 2 * mu * e


def svk_dpk2(g, lam, mu, h):
    de = 0.5 * (h + h.T + g.T @ h + h.T @ g)
    return lam * np.trace(de) * np.eye(g.shape[0]) + 2 * mu * de


@pytest.mark.parametrize("family,bonet", [("svk", None), ("nh_ciarlet", 0), ("nh_bonet", 1)])
def test_law_helpers_match_the_pinned_material_point(family, bonet):
    from oracle import oracle
    oracle.build()
    rng = np.random.default_rng(1)
    par = [1.3, 0.7]
    for _ in range(25):
        g = 0.2 * rng.uniform(-1, 1, (3, 3))
        h = rng.uniform(-1, 1, (3, 3))
        for dim in (3, 2):
            gd, hd = g[:dim, :dim], h[:dim, :dim]
            g3, h3 = np.zeros((3, 3)), np.zeros((3, 3))
            g3[:dim, :dim], h3[:dim, :dim] = gd, hd
            S, dS = oracle.hyper_law(family, g3, par)
            dref = np.einsum("ijkl,kl->ij", dS, h3)
            mine = svk_pk2(gd, *par) if bonet is None else nh_pk2(gd, *par, bonet)
            dmine = svk_dpk2(gd, *par, hd) if bonet is None else nh_dpk2(gd, *par, bonet, hd)
            assert np.abs(S[:dim, :dim] - mine).max() <= 1e-13 * np.abs(S).max()
            assert np.abs(dref[:dim, :dim] - dmine).max() <= 1e-13 * np.abs(dref).max()


_WRAPPERS = r"""
static mat ld(const double *p) { mat r; for (int i = 0; i < GF_N; ++i) for (int j = 0; j < GF_N; ++j) r.m[i][j] = p[i * GF_N + j]; return r; }
static void st(mat a, double *p) { for (int i = 0; i < GF_N; ++i) for (int j = 0; j < GF_N; ++j) p[i * GF_N + j] = a.m[i][j]; }
extern "C" {
void t_law(int law, const double *g, double lam, double mu, const double *h, double *S, double *dS) {
  if (law == 0) { st(svk_pk2(ld(g), lam, mu), S); st(svk_dpk2(ld(g), lam, mu, ld(h)), dS); }
  else { st(nh_pk2(ld(g), lam, mu, law - 1), S); st(nh_dpk2(ld(g), lam, mu, law - 1, ld(h)), dS); }
}
void t_iso(int law, const double *g, const double *p, const double *h, double *S, double *dS) {
  st(iso_pk2(ld(g), law, p[0], p[1], p[2], p[3], p[4]), S);
  st(iso_dpk2(ld(g), law, p[0], p[1], p[2], p[3], p[4], ld(h)), dS);
}
void t_norm(const double *x, const double *h, const double *k, double *out) {
  // out: |x| of a matrix, dnorm(x):h, d2norm(x,h,k); then the same for the first row taken as a vector
  const mat X = ld(x), H = ld(h), K = ld(k);
  out[0] = gnorm(X); out[1] = ddot(dnorm(X), H); out[2] = d2norm(X, H, K);
  vec xv, hv, kv;
  for (int i = 0; i < GF_N; ++i) { xv.v[i] = x[i]; hv.v[i] = h[i]; kv.v[i] = k[i]; }
  out[3] = gnorm(xv); out[4] = dot(dnorm(xv), hv); out[5] = d2norm(xv, hv, kv);
  const mat Z = 0.0 * X;
  out[6] = ddot(dnorm(Z), H);  // 0 at x = 0, like the reference
  out[7] = DER_PDFUNC1_DER_PDFUNC1_POW(1.7, 2.5); out[8] = DER_PDFUNC2_DER_PDFUNC1_POW(1.7, 2.5);
  out[9] = DER_PDFUNC1_DER_PDFUNC2_POW(1.7, 2.5); out[10] = DER_PDFUNC2_DER_PDFUNC2_POW(1.7, 2.5);
  out[11] = DER_PDFUNC_DER_PDFUNC_LOG(1.7); out[12] = DER_PDFUNC_DER_PDFUNC_SQRT(1.7); out[13] = DER_PDFUNC_DER_PDFUNC_TANH(0.4);
  out[14] = DER_PDFUNC_DER_PDFUNC_ATAN(0.4); out[15] = DER_PDFUNC_DER_PDFUNC_COS(0.4); out[16] = DER_PDFUNC_DER_PDFUNC_TAN(0.4);
  out[17] = DER_PDFUNC_DER_PDFUNC_ASIN(0.4); out[18] = DER_PDFUNC_DER_PDFUNC_ACOS(0.4); out[19] = DER_PDFUNC_DER_PDFUNC_ASINH(0.4);
  out[20] = DER_PDFUNC_DER_PDFUNC_ATANH(0.4); out[21] = DER_PDFUNC_DER_PDFUNC_ACOSH(1.4); out[22] = DER_PDFUNC_DER_PDFUNC_LOG10(1.7);
  // the piecewise functions at their kinks, as the reference's C functions define them (functions_and_operators.cc:77-168)
  out[23] = sign(0.0); out[24] = Heaviside(0.0); out[25] = DER_PDFUNC_NEG_PART(0.0); out[26] = DER_PDFUNC1_MAX(1.0, 1.0);
  out[27] = DER_PDFUNC2_MAX(1.0, 1.0); out[28] = DER_PDFUNC1_MAX(0.0, 1.0); out[29] = sinc(1e-5); out[30] = sinc(0.7);
  { const vec c = cross(xv, hv); for (int i = 0; i < GF_N; ++i) out[35 + i] = c.v[i]; }
  out[31] = DER_PDFUNC_SINC(0.7); out[32] = DER2_PDFUNC_SINC(0.7); out[33] = DER_PDFUNC_SINC(1e-5); out[34] = DER2_PDFUNC_SINC(1e-5);
}
void t_ops(const double *a, const double *h, const double *k, double *sc, double *m) {
  // scalars: det, ddet, d2det, mat_i2, dmat_i2, d2mat_i2, then the same for mat_j1 and mat_j2 ; matrices: inv, dinv, d2inv, rcg, drcg, d2rcg, lcg, dlcg, d2lcg, glag, dglag, d2glag
  const mat A = ld(a), H = ld(h), K = ld(k);
  sc[0] = det(A); sc[1] = ddet(A, H); sc[2] = d2det(A, H, K); sc[3] = mat_i2(A); sc[4] = dmat_i2(A, H); sc[5] = d2mat_i2(A, H, K);
  sc[6] = mat_j1(A); sc[7] = dmat_j1(A, H); sc[8] = d2mat_j1(A, H, K); sc[9] = mat_j2(A); sc[10] = dmat_j2(A, H); sc[11] = d2mat_j2(A, H, K);
  const int s = GF_N * GF_N;
  st(inv(A), m); st(dinv(A, H), m + s); st(d2inv(A, H, K), m + 2 * s);
  st(rcg(A), m + 3 * s); st(drcg(A, H), m + 4 * s); st(d2rcg(A, H, K), m + 5 * s);
  st(lcg(A), m + 6 * s); st(dlcg(A, H), m + 7 * s); st(d2lcg(A, H, K), m + 8 * s);
  st(glag(A), m + 9 * s); st(dglag(A, H), m + 10 * s); st(d2glag(A, H, K), m + 11 * s);
}
}
"""


def _host_helpers(n, tmp_path):
    """the DEVICE helper text of the run-time kernel template (jit.cu, everything before the kernels) compiled for the host:
    __device__ / __forceinline__ defined away, so the very functions NVRTC compiles are exercised without a GPU"""
    import ctypes
    import os
    import subprocess
    from conftest import ROOT
    src = open(os.path.join(ROOT, "getfem_b200", "csrc", "jit.cu")).read()
    body = src[src.index('R"GFJIT(') + len('R"GFJIT('):]
    body = body[:body.index("#if GF_Q == 1")]
    cc = os.path.join(str(tmp_path), "jit_helpers_%d.cc" % n)
    so = os.path.join(str(tmp_path), "jit_helpers_%d.so" % n)
    with open(cc, "w") as f:
        f.write("#include <cmath>\nusing namespace std;\n#define __device__\n#define __forceinline__ inline\n#define GF_N %d\n" % n)
        f.write(body)
        f.write(_WRAPPERS)
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-o", so, cc])
    return ctypes.CDLL(so)


@pytest.mark.parametrize("n", [2, 3])
def test_device_helper_text_on_the_host(n, tmp_path):
    import ctypes as C
    from oracle import oracle
    oracle.build()
    L = _host_helpers(n, tmp_path)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    L.t_law.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    L.t_ops.argtypes = [C.c_void_p] * 5
    rng = np.random.default_rng(7)
    lam, mu = 1.3, 0.7
    for _ in range(10):
        g = np.ascontiguousarray(0.2 * rng.uniform(-1, 1, (n, n)))
        h = np.ascontiguousarray(rng.uniform(-1, 1, (n, n)))
        k = np.ascontiguousarray(rng.uniform(-1, 1, (n, n)))
        # the laws against the pinned material point (2D: the plane-strain embedding)
        g3, h3 = np.zeros((3, 3)), np.zeros((3, 3))
        g3[:n, :n], h3[:n, :n] = g, h
        for law, fam in enumerate(("svk", "nh_ciarlet", "nh_bonet")):
            S, dS = np.zeros((n, n)), np.zeros((n, n))
            L.t_law(law, P(g), lam, mu, P(h), P(S), P(dS))
            Sr, dSr = oracle.hyper_law(fam, g3, [lam, mu])
            dr = np.einsum("ijkl,kl->ij", dSr, h3)
            assert np.abs(S - Sr[:n, :n]).max() <= 1e-13 * np.abs(Sr).max(), fam
            assert np.abs(dS - dr[:n, :n]).max() <= 1e-13 * np.abs(dr).max(), fam
        # the invariant-based laws (compressible Mooney-Rivlin, Ciarlet-Geymonat, generalized Blatz-Ko) and their plane-strain wrappers
        L.t_iso.argtypes = [C.c_int] + [C.c_void_p] * 5
        for law, (fam, par) in enumerate((("mooney_rivlin", [0.8, 0.3, 2.0]), ("ciarlet_geymonat", [1.3, 0.7, 0.25]),
                                          ("blatz_ko", [1.0, 1.0, 1.5, -0.5, 1.5]))):
            p5 = np.zeros(5)
            p5[:len(par)] = par
            S, dS = np.zeros((n, n)), np.zeros((n, n))
            L.t_iso(law, P(g), P(p5), P(h), P(S), P(dS))
            Sr, dSr = oracle.hyper_law(fam, g3, par)
            dr = np.einsum("ijkl,kl->ij", dSr, h3)
            assert np.abs(S - Sr[:n, :n]).max() <= 1e-12 * np.abs(Sr).max(), (fam, n)
            assert np.abs(dS - dr[:n, :n]).max() <= 1e-12 * np.abs(dr).max(), (fam, n)
        # Norm and the second derivatives of the predefined functions against central differences
        L.t_norm.argtypes = [C.c_void_p] * 4
        out = np.zeros(38)
        L.t_norm(P(g), P(h), P(k), P(out))
        e = 1e-5
        for off, (xx, hh, kk) in ((0, (g, h, k)), (3, (g[0], h[0], k[0]))):
            nf = np.linalg.norm
            assert abs(out[off] - nf(xx)) <= 1e-14
            assert abs(out[off + 1] - (nf(xx + e * hh) - nf(xx - e * hh)) / (2 * e)) <= 1e-7
            d2 = (nf(xx + e * hh + e * kk) - nf(xx + e * hh - e * kk) - nf(xx - e * hh + e * kk) + nf(xx - e * hh - e * kk)) / (4 * e * e)
            assert abs(out[off + 2] - d2) <= 1e-4 * max(1.0, abs(d2))
        assert out[6] == 0.0
        import math
        second = lambda fn, t: (fn(t + 1e-4) - 2 * fn(t) + fn(t - 1e-4)) / 1e-8
        assert abs(out[7] - second(lambda t: t ** 2.5, 1.7)) <= 1e-5
        mixed = (lambda t, u: ((t + 1e-4) ** (u + 1e-4) - (t + 1e-4) ** (u - 1e-4) - (t - 1e-4) ** (u + 1e-4) + (t - 1e-4) ** (u - 1e-4)) / 4e-8)(1.7, 2.5)
        assert abs(out[8] - mixed) <= 1e-5 and abs(out[9] - mixed) <= 1e-5
        assert abs(out[10] - second(lambda u: 1.7 ** u, 2.5)) <= 1e-5
        for idx, (fn, t) in zip(range(11, 23), ((math.log, 1.7), (math.sqrt, 1.7), (math.tanh, 0.4), (math.atan, 0.4), (math.cos, 0.4),
                                                 (math.tan, 0.4), (math.asin, 0.4), (math.acos, 0.4), (math.asinh, 0.4), (math.atanh, 0.4),
                                                 (math.acosh, 1.4), (math.log10, 1.7))):
            want = second(fn, t)
            if fn is math.cos:  # the reference's table gives DER_PDFUNC_COS' = "-cos(t)": the second derivative of cos
                want = -math.cos(t)
            assert abs(out[idx] - want) <= 2e-5 * max(1.0, abs(want)), (idx, out[idx], want)
        assert list(out[23:29]) == [1.0, 1.0, 0.0, 1.0, 1.0, 0.0]
        sincf = lambda t: math.sin(t) / t
        assert abs(out[29] - 1.0) <= 1e-10 and abs(out[30] - sincf(0.7)) <= 1e-15
        assert abs(out[31] - (sincf(0.7 + 1e-6) - sincf(0.7 - 1e-6)) / 2e-6) <= 1e-8 and abs(out[32] - second(sincf, 0.7)) <= 1e-5
        assert abs(out[33] + 1e-5 / 3.0) <= 1e-12 and abs(out[34] + 1.0 / 3.0) <= 1e-9
        if n == 3:
            assert np.abs(out[35:38] - np.cross(g[0], h[0])).max() <= 1e-15
        # the matrix operators: values against numpy, derivatives against central differences of the values
        a = np.ascontiguousarray(np.eye(n) + g)
        sc, m = np.zeros(12), np.zeros((12, n, n))
        L.t_ops(P(a), P(h), P(k), P(sc), P(m))
        f = {"det": np.linalg.det, "i2": lambda x: 0.5 * (np.trace(x) ** 2 - np.trace(x @ x)), "inv": np.linalg.inv,
             "j1": lambda x: np.trace(x) * np.linalg.det(x) ** (-1.0 / 3.0),
             "j2": lambda x: 0.5 * (np.trace(x) ** 2 - np.trace(x @ x)) * np.linalg.det(x) ** (-2.0 / 3.0),
             "rcg": lambda x: x.T @ x, "lcg": lambda x: x @ x.T, "glag": lambda x: 0.5 * (x.T @ x - np.eye(n))}
        e = 1e-5
        d1 = lambda fn: (fn(a + e * h) - fn(a - e * h)) / (2 * e)
        d2 = lambda fn: (fn(a + e * h + e * k) - fn(a + e * h - e * k) - fn(a - e * h + e * k) + fn(a - e * h - e * k)) / (4 * e * e)
        for j, name in enumerate(("det", "i2", "j1", "j2")):
            assert abs(sc[3 * j] - f[name](a)) <= 1e-13 * max(1.0, abs(f[name](a)))
            assert abs(sc[3 * j + 1] - d1(f[name])) <= 1e-7 * max(1.0, abs(sc[3 * j + 1]))
            assert abs(sc[3 * j + 2] - d2(f[name])) <= 1e-4 * max(1.0, abs(sc[3 * j + 2]))
        for j, name in enumerate(("inv", "rcg", "lcg", "glag")):
            assert np.abs(m[3 * j] - f[name](a)).max() <= 1e-13 * max(1.0, np.abs(f[name](a)).max())
            assert np.abs(m[3 * j + 1] - d1(f[name])).max() <= 1e-7 * max(1.0, np.abs(m[3 * j + 1]).max())
            assert np.abs(m[3 * j + 2] - d2(f[name])).max() <= 1e-4 * max(1.0, np.abs(m[3 * j + 2]).max())
