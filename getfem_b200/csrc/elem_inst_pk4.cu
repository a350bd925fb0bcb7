// elem_inst_pk4.cu -- explicit instantiations of the generic element kernel for P4 simplices (15 / 35 local nodes):
// the degree of the reference's own published table (contrib/opt_assembly/opt_assembly.cc:704-712).
#include "elem_kernel.cuh"
namespace gf {
bool launch_elem_inst_pk4(gfgpu_ctx *ctx, int dim, int Q, int nd, int fk, bool affine, const ElemArgs &a) {
  GF_ELEM_CASE(2, 1, 15, FK_LAPLACE, true)
  GF_ELEM_CASE(2, 1, 15, FK_MASS, true)
  GF_ELEM_CASE(2, 2, 15, FK_MASS, true)
  GF_ELEM_CASE(2, 2, 15, FK_ELAST, true)
  GF_ELEM_CASE(3, 1, 35, FK_LAPLACE, true)
  GF_ELEM_CASE(3, 1, 35, FK_MASS, true)
  GF_ELEM_CASE(3, 3, 35, FK_MASS, true)
  GF_ELEM_CASE(3, 3, 35, FK_ELAST, true)
  GF_ELEM_CASE(3, 3, 35, FK_HYPER, true)
  return false;
}
}  // namespace gf
