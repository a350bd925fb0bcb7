"""CPU: the NVRTC route compiles without a GPU (nvrtc targets sm_100a like nvcc does): the kernel template with valid
integrands compiles, an integrand with an error raises with the compiler's log, nothing falls back."""
import pytest


def test_valid_forms_compile_for_2d_and_3d():
    from getfem_b200 import capi
    f1 = "(1.0+sqr(u))*dot(gu,tg) + sin(u)*tv + par[0]*tv"
    f2 = "(2.0*u*t2v)*dot(gu,tg) + (1.0+sqr(u))*dot(t2g,tg) + cos(u)*t2v*tv"
    capi.jit_check(3, f1, f2)
    capi.jit_check(2, f1, f2)
    capi.jit_check(3, "dot(mkvec(1.0,2.0,3.0),tg)*gnorm(gu)+pos_part(u)*tv", "0.0*tv*t2v")
    # vector variable: u / tv are vec, gu / tg are mat
    lin1 = "par[0]*trace(gu)*trace(tg) + par[1]*ddot(gu+transp(gu),tg) + dot(u,tv)"
    lin2 = "par[0]*trace(t2g)*trace(tg) + par[1]*ddot(t2g+transp(t2g),tg) + dot(t2v,tv)"
    capi.jit_check(3, lin1, lin2, qdim=3)
    capi.jit_check(2, lin1, lin2, qdim=2)


def test_a_broken_form_raises_with_the_compiler_log():
    from getfem_b200 import capi
    with pytest.raises(capi.GfgpuError, match="does not compile"):
        capi.jit_check(3, "undefined_function(u)*tv", "0.0")
    with pytest.raises(capi.GfgpuError, match="does not compile"):
        capi.jit_check(3, "dot(gu,tg", "0.0")
    with pytest.raises(capi.GfgpuError):
        capi.jit_check(4, "tv", "tv*t2v")
