// potential.cu -- order 0: the scalar a ga_workspace::assembly(0) accumulates into assembled_potential()
// (src/getfem_generic_assembly_workspace.cc:791-803, ga_instruction_scalar_assembly C&E.cc:4628-4640).
//
//   quadratic forms (Laplace, elasticity, mass)  : E = 1/2 u^T K u = 1/2 u . R with the device residual R = K u
//   linear forms (source, normal source)         : E = u . R
//   finite-strain laws                           : E = alpha * sum_e sum_q w_q J_q W(E(Grad_u(x_q))) with the strain energies of
//       getfem_nonlinear_elasticity.cc -- Saint-Venant-Kirchhoff (:1996-2003 via AHL_wrapper_potential :1841-1928), Neo-Hookean
//       Ciarlet / Bonet (:612-632), compressible Mooney-Rivlin (:503-527), Ciarlet-Geymonat (:817-836), generalized Blatz-Ko
//       (:706-721), 1e200 where det(Id + Grad_u) <= 0 like the reference.
// One thread per region item, per-item energies summed by a CUB reduction (fixed tree: reproducible).
#include <cub/cub.cuh>

#include "common.cuh"
#include "elem_kernel.cuh"

namespace gf {

__device__ inline double strain_energy(int law, const double *Gu, const double *par) {
  double E[9], F[9], C[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double s = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) s += Gu[k + 3 * i] * Gu[k + 3 * j];
      E[i + 3 * j] = 0.5 * (s + Gu[i + 3 * j] + Gu[j + 3 * i]);
      C[i + 3 * j] = 2 * E[i + 3 * j] + (i == j ? 1.0 : 0.0);
      F[i + 3 * j] = Gu[i + 3 * j] + (i == j ? 1.0 : 0.0);
    }
  if (law == GFGPU_SVK) {
    const double tr = E[0] + E[4] + E[8];
    double n2 = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) n2 += E[i] * E[i];
    return tr * tr * par[0] / 2 + n2 * par[1];
  }
  if (det3cm(F) <= 0) return 1e200;
  const double i1 = C[0] + C[4] + C[8], i3 = det3cm(C);
  double ff = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) ff += C[i + 3 * j] * C[j + 3 * i];
  const double i2 = (i1 * i1 - ff) / 2;
  if (law == GFGPU_NEOHOOKEAN_CIARLET || law == GFGPU_NEOHOOKEAN_BONET) {
    const double lg = log(i3);
    double W = par[1] / 2 * (i1 - 3.0 - lg);
    W += law == GFGPU_NEOHOOKEAN_BONET ? par[0] / 8 * lg * lg : par[0] / 4 * (i3 - 1.0 - lg);
    return W;
  }
  if (law == GFGPU_MOONEY_RIVLIN) {
    const double j1 = i1 * pow(fabs(i3), -1.0 / 3.0), j2 = i2 * pow(fabs(i3), -2.0 / 3.0), s = sqrt(fabs(i3)) - 1.0;
    return par[0] * (j1 - 3.0) + par[1] * (j2 - 3.0) + par[2] * s * s;
  }
  if (law == GFGPU_CIARLET_GEYMONAT) {
    const double a = par[2], b = par[1] / 2 - par[2], c = par[0] / 4 - par[1] / 2 + par[2], d = par[0] / 2 + par[1];
    const double e = -(3.0 * (a + b) + c);
    double n2 = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) n2 += C[i] * C[i];
    return a * i1 + b * (i1 * i1 - n2) / 2 + c * i3 - d * log(i3) / 2 + e;
  }
  return pow(par[0] * i1 + par[1] * sqrt(fabs(i3)) + par[2] * i2 / i3 + par[3], par[4]);  // generalized Blatz-Ko
}

struct PotArgs {
  const double *x, *y, *z;
  const int32_t *conn, *edof;
  const double *U, *w, *gt_grad, *gphi;
  int nq, ng, nd, law, affine;
  int64_t e0, ne;
  double par[GFGPU_MAX_PARAMS];
  double alpha;
  double *out;
};

__global__ void k_potential(const PotArgs a) {
  for (int64_t el = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; el < a.ne; el += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = a.e0 + el;
    double G[24];
    for (int i = 0; i < a.ng; ++i) {
      const int32_t p = a.conn[e * a.ng + i];
      G[3 * i] = a.x[p]; G[3 * i + 1] = a.y[p]; G[3 * i + 2] = a.z[p];
    }
    double geo[10], energy = 0;
    for (int q = 0; q < a.nq; ++q) {
      const double wq = a.w[q];
      if (wq == 0.0) continue;  // zero-weight points are skipped (C&E.cc:8852-8864)
      if (q == 0 || !a.affine) geometry<3>(G, a.gt_grad + (size_t)q * a.ng * 3, a.ng, geo);
      double Gu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int i = 0; i < a.nd; ++i) {
        const double *gh = a.gphi + ((size_t)q * a.nd + i) * 3;
        double Z[3];
#pragma unroll
        for (int n = 0; n < 3; ++n) Z[n] = gh[0] * geo[n] + gh[1] * geo[n + 3] + gh[2] * geo[n + 6];
        const int32_t d0 = a.edof[e * a.nd + i];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double u = a.U ? a.U[d0 + c] : 0.0;
#pragma unroll
          for (int n = 0; n < 3; ++n) Gu[c + 3 * n] += u * Z[n];
        }
      }
      energy += wq * geo[9] * strain_energy(a.law, Gu, a.par);
    }
    a.out[el] = a.alpha * energy;
  }
}

__global__ void k_dot_terms(const double *__restrict__ a, const double *__restrict__ b, int64_t n, double *__restrict__ out) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) out[k] = a[k] * b[k];
}

static double device_sum(gfgpu_ctx *ctx, const double *v, int64_t n) {
  DevBuf<double> out;
  out.alloc(ctx, 1);
  size_t tb = 0;
  GF_CUDA(cub::DeviceReduce::Sum(nullptr, tb, v, out.p, n, ctx->stream));
  void *tmp = cub_scratch(ctx, tb);
  GF_CUDA(cub::DeviceReduce::Sum(tmp, tb, v, out.p, n, ctx->stream));
  count_launch(2);
  double h = 0;
  out.download(&h);
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  return h;
}

void term_assemble_for_potential(gfgpu_term *t, const double *U_dev);  // api.cu: residual-only assembly

double term_potential(gfgpu_term *t, const double *U_dev) {
  gfgpu_ctx *ctx = t->ctx;
  const int fam = t->family;
  if (fam == GFGPU_JIT) {
    // run-time compiled term: the order-0 integrand is evaluated by the same kernel, next to the element residuals; the
    // elements' shares are summed in one reduction (ga_instruction_scalar_assembly, C&E.cc:4628-4640)
    GF_REQUIRE(!t->jit_form0.empty(), "this JIT term carries no order-0 form (gfgpu_term_set_jit_potential)");
    const int64_t ne = t->e1 - t->e0;
    if (ne <= 0) return 0.0;
    DevBuf<double> out;
    out.alloc(ctx, ne);
    out.zero();
    t->jit_epot = out.p;
    try { term_assemble_for_potential(t, U_dev); } catch (...) { t->jit_epot = nullptr; throw; }
    t->jit_epot = nullptr;
    return device_sum(ctx, out.p, ne);
  }
  const bool hyper = fam == GFGPU_SVK || fam == GFGPU_NEOHOOKEAN_CIARLET || fam == GFGPU_NEOHOOKEAN_BONET || fam >= GFGPU_MOONEY_RIVLIN;
  if (!hyper) {
    GF_REQUIRE(U_dev, "the potential of a linear or quadratic form needs the state vector");
    term_assemble_for_potential(t, U_dev);
    const int64_t n = t->fem->ndof;
    DevBuf<double> prod;
    prod.alloc(ctx, n);
    k_dot_terms<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, ctx->stream>>>(U_dev, t->R.p, n, prod.p);
    GF_LAUNCH_CHECK();
    const double d = device_sum(ctx, prod.p, n);
    return (fam == GFGPU_SOURCE || fam == GFGPU_NORMAL_SOURCE) ? d : 0.5 * d;
  }
  GF_REQUIRE(!t->region_faces, "the strain energy is a volume integral: the region must hold convexes");
  GF_REQUIRE(t->nfields == 0, "fem-data coefficients are not handled for the finite-strain potentials");
  const int64_t ne = t->e1 - t->e0;
  if (ne <= 0) return 0.0;
  PotArgs a;
  const int64_t np = t->mesh->npts;
  a.x = t->mesh->xyz.p; a.y = a.x + np; a.z = a.x + 2 * np;
  a.conn = t->conn_p(); a.edof = t->edof_p();
  a.U = U_dev; a.w = t->tab->w.p; a.gt_grad = t->tab->gt_grad.p; a.gphi = t->tab->gphi.p;
  a.nq = t->tab->nq; a.ng = t->mesh->ng; a.nd = t->fem->nd; a.law = fam;
  a.affine = t->mesh->gt_kind == GFGPU_GT_PK;
  a.e0 = t->e0; a.ne = ne;
  for (int k = 0; k < GFGPU_MAX_PARAMS; ++k) a.par[k] = t->par[k];
  a.alpha = t->alpha;
  GF_REQUIRE(a.ng <= 8, "geometric transformation with more than 8 nodes");
  DevBuf<double> out;
  out.alloc(ctx, ne);
  a.out = out.p;
  k_potential<<<(int)std::min<int64_t>((ne + 127) / 128, 148 * 16), 128, 0, ctx->stream>>>(a);
  GF_LAUNCH_CHECK();
  return device_sum(ctx, out.p, ne);
}

}  // namespace gf
