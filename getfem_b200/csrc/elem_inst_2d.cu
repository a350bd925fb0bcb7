// elem_inst_2d.cu -- explicit instantiations of the generic element kernel (2D PK/QK).
#include "elem_kernel.cuh"
namespace gf {
bool launch_elem_inst_2d(gfgpu_ctx *ctx, int dim, int Q, int nd, int fk, bool affine, const ElemArgs &a) {
  GF_ELEM_CASE(2, 1, 3, FK_LAPLACE, true)
  GF_ELEM_CASE(2, 2, 3, FK_LAPLACE, true)
  GF_ELEM_CASE(2, 1, 3, FK_MASS, true)
  GF_ELEM_CASE(2, 2, 3, FK_MASS, true)
  GF_ELEM_CASE(2, 2, 3, FK_ELAST, true)
  GF_ELEM_CASE(2, 1, 6, FK_LAPLACE, true)
  GF_ELEM_CASE(2, 2, 6, FK_LAPLACE, true)
  GF_ELEM_CASE(2, 1, 6, FK_MASS, true)
  GF_ELEM_CASE(2, 2, 6, FK_MASS, true)
  GF_ELEM_CASE(2, 2, 6, FK_ELAST, true)
  GF_ELEM_CASE(2, 1, 10, FK_LAPLACE, true)
  GF_ELEM_CASE(2, 2, 10, FK_LAPLACE, true)
  GF_ELEM_CASE(2, 1, 10, FK_MASS, true)
  GF_ELEM_CASE(2, 2, 10, FK_MASS, true)
  GF_ELEM_CASE(2, 2, 10, FK_ELAST, true)
  GF_ELEM_CASE(2, 1, 4, FK_LAPLACE, false)
  GF_ELEM_CASE(2, 2, 4, FK_LAPLACE, false)
  GF_ELEM_CASE(2, 1, 4, FK_MASS, false)
  GF_ELEM_CASE(2, 2, 4, FK_MASS, false)
  GF_ELEM_CASE(2, 2, 4, FK_ELAST, false)
  GF_ELEM_CASE(2, 1, 9, FK_LAPLACE, false)
  GF_ELEM_CASE(2, 2, 9, FK_LAPLACE, false)
  GF_ELEM_CASE(2, 1, 9, FK_MASS, false)
  GF_ELEM_CASE(2, 2, 9, FK_MASS, false)
  GF_ELEM_CASE(2, 2, 9, FK_ELAST, false)
  GF_ELEM_CASE(2, 1, 16, FK_LAPLACE, false)
  GF_ELEM_CASE(2, 2, 16, FK_LAPLACE, false)
  GF_ELEM_CASE(2, 1, 16, FK_MASS, false)
  GF_ELEM_CASE(2, 2, 16, FK_MASS, false)
  GF_ELEM_CASE(2, 2, 16, FK_ELAST, false)
  return false;
}
}  // namespace gf
