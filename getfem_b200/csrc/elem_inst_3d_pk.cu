// elem_inst_3d_pk.cu -- explicit instantiations of the generic element kernel (3D simplices, affine).
#include "elem_kernel.cuh"
namespace gf {
bool launch_elem_inst_3d_pk(gfgpu_ctx *ctx, int dim, int Q, int nd, int fk, bool affine, const ElemArgs &a) {
  GF_ELEM_CASE(3, 1, 4, FK_LAPLACE, true)
  GF_ELEM_CASE(3, 3, 4, FK_LAPLACE, true)
  GF_ELEM_CASE(3, 1, 4, FK_MASS, true)
  GF_ELEM_CASE(3, 3, 4, FK_MASS, true)
  GF_ELEM_CASE(3, 3, 4, FK_ELAST, true)
  GF_ELEM_CASE(3, 3, 4, FK_HYPER, true)
  GF_ELEM_CASE(3, 1, 10, FK_LAPLACE, true)
  GF_ELEM_CASE(3, 3, 10, FK_LAPLACE, true)
  GF_ELEM_CASE(3, 1, 10, FK_MASS, true)
  GF_ELEM_CASE(3, 3, 10, FK_MASS, true)
  GF_ELEM_CASE(3, 3, 10, FK_ELAST, true)
  GF_ELEM_CASE(3, 3, 10, FK_HYPER, true)
  GF_ELEM_CASE(3, 1, 20, FK_LAPLACE, true)
  GF_ELEM_CASE(3, 3, 20, FK_LAPLACE, true)
  GF_ELEM_CASE(3, 1, 20, FK_MASS, true)
  GF_ELEM_CASE(3, 3, 20, FK_MASS, true)
  GF_ELEM_CASE(3, 3, 20, FK_ELAST, true)
  GF_ELEM_CASE(3, 3, 20, FK_HYPER, true)
  return false;
}
}  // namespace gf
