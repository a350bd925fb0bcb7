// elem_dispatch.cu -- family -> kernel class mapping and dispatch over the instantiation files.
#include "elem_kernel.cuh"
namespace gf {
bool launch_elem_inst_2d(gfgpu_ctx *, int, int, int, int, bool, const ElemArgs &);
bool launch_elem_inst_3d_pk(gfgpu_ctx *, int, int, int, int, bool, const ElemArgs &);
bool launch_elem_inst_3d_qk(gfgpu_ctx *, int, int, int, int, bool, const ElemArgs &);
bool launch_elem_inst_3d_qk_hi(gfgpu_ctx *, int, int, int, int, bool, const ElemArgs &);
bool launch_elem_inst_pk4(gfgpu_ctx *, int, int, int, int, bool, const ElemArgs &);

bool launch_elem_kernel(gfgpu_ctx *ctx, int dim, int Q, int nd, bool affine, const ElemArgs &a) {
  int fk;
  switch (a.family) {
    case GFGPU_LAPLACE: fk = FK_LAPLACE; break;
    case GFGPU_MASS: case GFGPU_SOURCE: case GFGPU_NORMAL_SOURCE: fk = FK_MASS; break;  // source terms = the flux part of the mass kernel
    case GFGPU_ELASTICITY: fk = FK_ELAST; break;
    case GFGPU_SVK: case GFGPU_NEOHOOKEAN_CIARLET: case GFGPU_NEOHOOKEAN_BONET:
    case GFGPU_MOONEY_RIVLIN: case GFGPU_CIARLET_GEYMONAT: case GFGPU_BLATZ_KO: fk = FK_HYPER; break;
    default: return false;
  }
  return launch_elem_inst_2d(ctx, dim, Q, nd, fk, affine, a) ||
         launch_elem_inst_3d_pk(ctx, dim, Q, nd, fk, affine, a) ||
         launch_elem_inst_3d_qk(ctx, dim, Q, nd, fk, affine, a) ||
         launch_elem_inst_3d_qk_hi(ctx, dim, Q, nd, fk, affine, a) ||
         launch_elem_inst_pk4(ctx, dim, Q, nd, fk, affine, a);
}
}  // namespace gf
