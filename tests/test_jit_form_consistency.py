"""CPU: derivative consistency of what the NVRTC route compiles.  For every expression of the GPU lists the shim's dry run prints
the three C forms it translated from the reference's analysed trees (order 0 / 1 / 2).  Here those very strings are compiled for
the HOST together with the device helper text of the kernel template (jit.cu, `__device__` defined away) and checked at random
material points: the second form is the directional derivative of the first, and -- when the expression is a potential -- the
first form is the directional derivative of the order-0 form.  This pins the translation of the derivative nodes
(`Derivative_1_Op(A):H`, `(Derivative_1_1_Op(A):H2):H1`, `X:(Derivative_1_1_f(A):Y)`, `DER_PDFUNC_*`) and the helpers behind them
without a GPU; the values themselves are compared with the reference on the device (tests/test_gpu_dropin.py,
tests/test_gpu_zz_law_operators.py)."""
import ctypes as C
import importlib.util
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

BIN = os.path.join(ROOT, "oracle", "_ref", "model_test")

_WRAP_SCALAR = r"""
static vec lv(const double *p) { vec r; for (int i = 0; i < GF_N; ++i) r.v[i] = p[i]; return r; }
extern "C" {
double f0(const double *u, const double *g, const double *x, const double *n, const double *fld, const double *vf, const double *par) {
  const vec vfld = lv(vf), X = lv(x), Normal = lv(n), gu = lv(g); const double uu = u[0]; (void)vfld; (void)X; (void)Normal; (void)gu; (void)uu;
  { const double u = uu; return GF_FORM0; }
}
double f1(const double *u, const double *g, const double *x, const double *n, const double *fld, const double *vf, const double *par,
          const double *t, const double *tgp) {
  const vec vfld = lv(vf), X = lv(x), Normal = lv(n), gu = lv(g), tg = lv(tgp); const double uu = u[0], tv = t[0];
  { const double u = uu; return GF_FORM1; }
}
double f2(const double *u, const double *g, const double *x, const double *n, const double *fld, const double *vf, const double *par,
          const double *t, const double *tgp, const double *t2, const double *t2gp) {
  const vec vfld = lv(vf), X = lv(x), Normal = lv(n), gu = lv(g), tg = lv(tgp), t2g = lv(t2gp); const double uu = u[0], tv = t[0], t2v = t2[0];
  { const double u = uu; return GF_FORM2; }
}
}
"""

_WRAP_VECTOR = r"""
static vec lv(const double *p) { vec r; for (int i = 0; i < GF_N; ++i) r.v[i] = p[i]; return r; }
static mat lm(const double *p) { mat r; for (int i = 0; i < GF_N; ++i) for (int j = 0; j < GF_N; ++j) r.m[i][j] = p[i * GF_N + j]; return r; }
extern "C" {
double f0(const double *uu, const double *g, const double *x, const double *n, const double *fld, const double *vf, const double *par) {
  const vec vfld = lv(vf), X = lv(x), Normal = lv(n), u = lv(uu); const mat gu = lm(g); (void)vfld; (void)X; (void)Normal; (void)u; (void)gu;
  return GF_FORM0;
}
double f1(const double *uu, const double *g, const double *x, const double *n, const double *fld, const double *vf, const double *par,
          const double *t, const double *tgp) {
  const vec vfld = lv(vf), X = lv(x), Normal = lv(n), u = lv(uu), tv = lv(t); const mat gu = lm(g), tg = lm(tgp);
  return GF_FORM1;
}
double f2(const double *uu, const double *g, const double *x, const double *n, const double *fld, const double *vf, const double *par,
          const double *t, const double *tgp, const double *t2, const double *t2gp) {
  const vec vfld = lv(vf), X = lv(x), Normal = lv(n), u = lv(uu), tv = lv(t), t2v = lv(t2); const mat gu = lm(g), tg = lm(tgp), t2g = lm(t2gp);
  return GF_FORM2;
}
}
"""


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tests", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _cases():
    zz, dr = _load("test_gpu_zz_law_operators"), _load("test_gpu_dropin")
    ex = dr.JIT + dr.JIT_VECTOR + dr.JIT_X + zz.COMPOUND + zz.NORM_POTENTIALS + zz.CONSTANTS
    ex += [p for p in dr.POTENTIALS if "_potential(" not in p[1]]
    out = [(["model=expr"] + m.split() + ["expr=" + e]) for m, e in ex]
    return out + [c.split() for c in zz.MODELS] + ["model=finite_strain dim=2 n=3 gt=pk k=2".split()]


@pytest.mark.parametrize("args", _cases(), ids=lambda a: " ".join(a)[-70:])
def test_translated_forms_are_consistent_derivatives(args, tmp_path):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/model_test not built (needs the reference sources)")
    small = [("n=2" if a.startswith("n=") else a) for a in args]  # the forms do not depend on the mesh size
    out = subprocess.run([BIN] + small, capture_output=True, text=True, timeout=300, env=dict(os.environ, GFGPU_DRYRUN="1"))
    assert out.returncode == 0, out.stderr[-1500:]
    forms = {tuple(l.split("\t")) for l in out.stderr.splitlines() if l.startswith("[gfgpu dryrun jit]")}
    if not forms:
        pytest.skip("a closed-form family, not a run-time compiled term")
    src = open(os.path.join(ROOT, "getfem_b200", "csrc", "jit.cu")).read()
    body = src[src.index('R"GFJIT(') + len('R"GFJIT('):]
    body = body[:body.index("#if GF_Q == 1")]
    rng = np.random.default_rng(5)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    for k, (head, f1, f2, f0) in enumerate(sorted(forms)):
        m = re.search(r"dim=(\d) qdim=(\d)", head)
        n, q = int(m.group(1)), int(m.group(2))
        cc, so = os.path.join(str(tmp_path), "f%d.cc" % k), os.path.join(str(tmp_path), "f%d.so" % k)
        with open(cc, "w") as f:
            f.write("#include <cmath>\nusing namespace std;\n#define __device__\n#define __forceinline__ inline\n#define GF_N %d\n" % n)
            f.write("#define GF_FORM0 (%s)\n#define GF_FORM1 (%s)\n#define GF_FORM2 (%s)\n" % (f0 or "0.0", f1, f2))
            f.write(body)
            f.write(_WRAP_SCALAR if q == 1 else _WRAP_VECTOR)
        subprocess.check_call(["g++", "-O1", "-w", "-shared", "-fPIC", "-o", so, cc])
        L = C.CDLL(so)
        for fn, na in ((L.f0, 7), (L.f1, 9), (L.f2, 11)):
            fn.restype = C.c_double
            fn.argtypes = [C.c_void_p] * na
        su, sg = (1,), (n,)
        if q > 1:
            su, sg = (n,), (n, n)
        worst1 = worst0 = 0.0
        for _ in range(6):
            # a state with det(I + Grad_u) > 0 and away from the kinks of the piecewise functions
            u = np.ascontiguousarray(0.3 + 0.4 * rng.uniform(0, 1, su))
            g = np.ascontiguousarray(0.15 * rng.uniform(-1, 1, sg))
            x = np.ascontiguousarray(rng.uniform(0.2, 1, 3))
            nr = rng.uniform(-1, 1, 3)
            nr = np.ascontiguousarray(nr / np.linalg.norm(nr[:n]))
            fld = np.ascontiguousarray(rng.uniform(0.5, 2, 2))
            vf = np.ascontiguousarray(rng.uniform(-1, 1, 3))
            par = np.ascontiguousarray(np.array([1.3, 0.7, 0.25, 0.9, 1.5, 0.4, 0.6, 0.8, 1.1, 0.5, 1.0, 1.5]))
            tv, tg = np.ascontiguousarray(rng.uniform(-1, 1, su)), np.ascontiguousarray(rng.uniform(-1, 1, sg))
            t2v, t2g = np.ascontiguousarray(rng.uniform(-1, 1, su)), np.ascontiguousarray(rng.uniform(-1, 1, sg))
            e = 1e-6
            F1 = lambda uu, gg: L.f1(P(uu), P(gg), P(x), P(nr), P(fld), P(vf), P(par), P(tv), P(tg))
            a2 = L.f2(P(u), P(g), P(x), P(nr), P(fld), P(vf), P(par), P(tv), P(tg), P(t2v), P(t2g))
            up, um = np.ascontiguousarray(u + e * t2v), np.ascontiguousarray(u - e * t2v)
            gp, gm = np.ascontiguousarray(g + e * t2g), np.ascontiguousarray(g - e * t2g)
            d1 = (F1(up, gp) - F1(um, gm)) / (2 * e)
            worst1 = max(worst1, abs(a2 - d1) / max(1.0, abs(a2)))
            if f0:
                F0 = lambda uu, gg: L.f0(P(uu), P(gg), P(x), P(nr), P(fld), P(vf), P(par))
                up, um = np.ascontiguousarray(u + e * tv), np.ascontiguousarray(u - e * tv)
                gp, gm = np.ascontiguousarray(g + e * tg), np.ascontiguousarray(g - e * tg)
                d0 = (F0(up, gp) - F0(um, gm)) / (2 * e)
                worst0 = max(worst0, abs(F1(u, g) - d0) / max(1.0, abs(d0)))
        assert worst1 <= 2e-6, (f1[:120], f2[:120], worst1)
        assert worst0 <= 2e-6, (f0[:120], f1[:120], worst0)
