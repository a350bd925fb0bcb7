"""getfem_b200 -- B200-native generic weak-form assembly behind GetFEM's ga_workspace interface.

csrc/            hand-written sm_100a kernels + the C ABI of include/gfgpu.h  (libgfgpu.so)
capi.py          ctypes binding of that ABI
workspace.py     host mirror of mesh / mesh_fem / mesh_im / ga_workspace for the accelerated path
regular_mesh.py  regular_unit_mesh numbering;  fem_tables.py  PK/QK Lagrange + cubature tables
shim/            C++ drop-in for the real GetFEM (compiled against its headers)
"""
from . import capi, fem_tables  # noqa: F401
from .capi import GfgpuError  # noqa: F401
from .workspace import (default_context, ga_workspace, mesh, mesh_fem, mesh_im, mesh_region,  # noqa: F401
                        outer_faces_of_mesh, regular_unit_mesh)
