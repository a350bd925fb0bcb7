"""CPU: the C-ABI library loads and exports every symbol include/gfgpu.h declares (no compute)."""
import os
import re

from conftest import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "gfgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gfgpu_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from getfem_b200 import build, capi
    build.build()
    L = capi.lib()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), "libgfgpu.so does not export " + s
    assert sorted(capi.SIGNATURES) == syms, "capi.SIGNATURES is out of sync with include/gfgpu.h"
    assert L.gfgpu_version() >= 100


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the context creation must fail loudly (no silent CPU path)."""
    import pytest
    import torch
    from getfem_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.GfgpuError, match="no CUDA device"):
        capi.Context(0)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under getfem_b200/ may reference it."""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "getfem_b200")):
        if "_obj" in dp or "__pycache__" in dp:
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".cpp")):
                t = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"(^|\s)(from|import)\s+oracle\b|oracle/|asm_oracle|libgfo", t):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
