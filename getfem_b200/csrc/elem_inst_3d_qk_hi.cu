// elem_inst_3d_qk_hi.cu -- explicit instantiations of the generic element kernel (3D hexahedra Q3/Q4).
#include "elem_kernel.cuh"
namespace gf {
bool launch_elem_inst_3d_qk_hi(gfgpu_ctx *ctx, int dim, int Q, int nd, int fk, bool affine, const ElemArgs &a) {
  GF_ELEM_CASE(3, 1, 64, FK_LAPLACE, false)
  GF_ELEM_CASE(3, 1, 64, FK_MASS, false)
  GF_ELEM_CASE(3, 3, 64, FK_ELAST, false)
  GF_ELEM_CASE(3, 1, 125, FK_LAPLACE, false)
  GF_ELEM_CASE(3, 1, 125, FK_MASS, false)
  return false;
}
}  // namespace gf
