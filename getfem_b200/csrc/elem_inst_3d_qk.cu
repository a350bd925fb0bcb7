// elem_inst_3d_qk.cu -- explicit instantiations of the generic element kernel (3D hexahedra Q1/Q2).
#include "elem_kernel.cuh"
namespace gf {
bool launch_elem_inst_3d_qk(gfgpu_ctx *ctx, int dim, int Q, int nd, int fk, bool affine, const ElemArgs &a) {
  GF_ELEM_CASE(3, 1, 8, FK_LAPLACE, false)
  GF_ELEM_CASE(3, 3, 8, FK_LAPLACE, false)
  GF_ELEM_CASE(3, 1, 8, FK_MASS, false)
  GF_ELEM_CASE(3, 3, 8, FK_MASS, false)
  GF_ELEM_CASE(3, 3, 8, FK_ELAST, false)
  GF_ELEM_CASE(3, 3, 8, FK_HYPER, false)
  GF_ELEM_CASE(3, 1, 27, FK_LAPLACE, false)
  GF_ELEM_CASE(3, 3, 27, FK_LAPLACE, false)
  GF_ELEM_CASE(3, 1, 27, FK_MASS, false)
  GF_ELEM_CASE(3, 3, 27, FK_MASS, false)
  GF_ELEM_CASE(3, 3, 27, FK_ELAST, false)
  GF_ELEM_CASE(3, 3, 27, FK_HYPER, false)
  return false;
}
}  // namespace gf
