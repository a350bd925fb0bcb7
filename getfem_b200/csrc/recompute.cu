// recompute.cu -- strategy RECOMPUTE: tangent of an affine-geometry, constant-coefficient bilinear
// form assembled per NONZERO, with no element-matrix staging in HBM.
//
// For a simplex with an affine transformation, B = K^-T and J are constant over the element
// (bgeot_geometric_trans.h:462-468, C&E.cc:8836), so the quadrature loop of the reference
//     K_e(i a, j b) = sum_q J w_q  Z_q(i,n) D(a,n,b,l) Z_q(j,l),   Z_q(i,n) = sum_p ghat_q(i,p) B(n,p)
// factorises through element-independent reference tensors
//     M^{ij}(p,q) = sum_k w_k ghat_k(i,p) ghat_k(j,q)            (nd x nd x N x N, built once on the host
//                                                                  from the SAME tables the reference uses)
//     T = B M^{ij} B^T  (N x N)
//   Laplace    : K_e(i,j)       = a J tr(T)                      = <M^{ij}, a J B^T B>
//   elasticity : K_e(i a, j b)  = J [ lambda T(a,b) + mu T(b,a) + mu delta_ab tr(T) ]
//   mass       : K_e(i,j)       = a J sum_k w_k phi_k(i) phi_k(j)
// (getfem_models.cc:6112-6113).  Each (column node J, row node I) pair sums its
// (element, j, i) contributions in ascending element order and writes its kept entries straight into
// the CSC value array: HBM traffic = contribution list + per-element geometry + values, nothing else.
// The keep masks / CSC pattern come from the generic element kernel run once in mask-only mode, so the
// pattern is still the reference's per-element drop rule evaluated on the quadrature form.
#include "common.cuh"

namespace gf {

enum { RF_LAPLACE = 0, RF_ELAST = 1, RF_MASS = 2 };

template <int N>
__global__ void k_affine_geo(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
                             const int32_t *__restrict__ conn, int ng, const double *__restrict__ pc /* ng x N */,
                             int64_t e0, int64_t ne, int rf, double scale, double *__restrict__ eg, int egsz) {
  for (int64_t el = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; el < ne; el += (int64_t)gridDim.x * blockDim.x) {
    const int32_t *cv = conn + (e0 + el) * ng;
    double K[N * N];
#pragma unroll
    for (int k = 0; k < N * N; ++k) K[k] = 0.0;
    for (int i = 0; i < ng; ++i) {
      const int32_t p = cv[i];
      double g[3] = {x[p], y[p], N == 3 ? z[p] : 0.0};
#pragma unroll
      for (int c = 0; c < N; ++c)
#pragma unroll
        for (int r = 0; r < N; ++r) K[r + N * c] += g[r] * pc[i * N + c];
    }
    double B[N * N], J;
    if (N == 2) {
      double d = K[0] * K[3] - K[1] * K[2], id = 1.0 / d;
      B[0] = K[3] * id; B[2] = -K[1] * id; B[1] = -K[2] * id; B[3] = K[0] * id;
      J = fabs(d);
    } else {
#define K_(i, j) K[(i) + 3 * (j)]
      double c00 = K_(1, 1) * K_(2, 2) - K_(1, 2) * K_(2, 1);
      double c10 = K_(1, 2) * K_(2, 0) - K_(1, 0) * K_(2, 2);
      double c20 = K_(1, 0) * K_(2, 1) - K_(1, 1) * K_(2, 0);
      double d = K_(0, 0) * c00 + K_(0, 1) * c10 + K_(0, 2) * c20, id = 1.0 / d;
      B[0] = c00 * id; B[3] = c10 * id; B[6] = c20 * id;
      B[1] = (K_(0, 2) * K_(2, 1) - K_(0, 1) * K_(2, 2)) * id;
      B[4] = (K_(0, 0) * K_(2, 2) - K_(0, 2) * K_(2, 0)) * id;
      B[7] = (K_(0, 1) * K_(2, 0) - K_(0, 0) * K_(2, 1)) * id;
      B[2] = (K_(0, 1) * K_(1, 2) - K_(0, 2) * K_(1, 1)) * id;
      B[5] = (K_(0, 2) * K_(1, 0) - K_(0, 0) * K_(1, 2)) * id;
      B[8] = (K_(0, 0) * K_(1, 1) - K_(0, 1) * K_(1, 0)) * id;
#undef K_
      J = fabs(d);
    }
    double *o = eg + (size_t)el * egsz;
    if (rf == RF_ELAST) {
#pragma unroll
      for (int k = 0; k < N * N; ++k) o[k] = B[k];
      o[N * N] = scale * J;
    } else if (rf == RF_LAPLACE) {  // scale*J * B^T B, upper triangle row by row
      int k = 0;
#pragma unroll
      for (int p = 0; p < N; ++p)
#pragma unroll
        for (int q = p; q < N; ++q) {
          double s = 0;
#pragma unroll
          for (int n = 0; n < N; ++n) s += B[n + N * p] * B[n + N * q];
          o[k++] = scale * J * s;
        }
    } else {
      o[0] = scale * J;
    }
  }
}

template <int N, int RF>
struct RcCfg {
  static constexpr int EG = RF == RF_ELAST ? N * N + 1 : RF == RF_LAPLACE ? N * (N + 1) / 2 : 1;
  static constexpr int MT = RF == RF_ELAST ? N * N : RF == RF_LAPLACE ? N * (N + 1) / 2 : 1;  // table entries per (j,i)
};

// v1: one thread per node pair.
template <int N, int Q, int ND, int RF>
__global__ void __launch_bounds__(256)
k_recompute_pairs(const uint32_t *__restrict__ cstart, const uint32_t *__restrict__ csrc,
                  const int32_t *__restrict__ pJ, const uint16_t *__restrict__ pmask,
                  const uint32_t *__restrict__ prel, const int64_t *__restrict__ jc,
                  const double *__restrict__ eg, const double *__restrict__ Mtab, double lambda, double mu,
                  int64_t npairs, double *__restrict__ pr) {
  using C = RcCfg<N, RF>;
  constexpr int NB = ND * ND, MT = C::MT, EG = C::EG;
  __shared__ double sM[NB * MT];
  for (int k = threadIdx.x; k < NB * MT; k += blockDim.x) sM[k] = Mtab[k];
  __syncthreads();
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npairs; p += (int64_t)gridDim.x * blockDim.x) {
    double acc[RF == RF_ELAST ? Q * Q : 1];
#pragma unroll
    for (int m = 0; m < (RF == RF_ELAST ? Q * Q : 1); ++m) acc[m] = 0.0;
    for (uint32_t s = cstart[p], e = cstart[p + 1]; s < e; ++s) {
      const uint32_t c = csrc[s];
      const uint32_t el = c / NB, r = c - el * NB;
      const double *g = eg + (size_t)el * EG;
      const double *M = sM + r * MT;
      if (RF == RF_ELAST) {
        double B[N * N], W[N * N];
#pragma unroll
        for (int k = 0; k < N * N; ++k) B[k] = g[k];
        const double J = g[N * N];
        // W(a,q) = sum_p B(a,p) M(p,q) ; T(a,b) = sum_q W(a,q) B(b,q)
#pragma unroll
        for (int q = 0; q < N; ++q)
#pragma unroll
          for (int a = 0; a < N; ++a) {
            double s2 = 0;
#pragma unroll
            for (int pp = 0; pp < N; ++pp) s2 += B[a + N * pp] * M[pp * N + q];
            W[a + N * q] = s2;
          }
        double T[N * N], tr = 0;
#pragma unroll
        for (int b = 0; b < N; ++b)
#pragma unroll
          for (int a = 0; a < N; ++a) {
            double s2 = 0;
#pragma unroll
            for (int q = 0; q < N; ++q) s2 += W[a + N * q] * B[b + N * q];
            T[a + N * b] = s2;
            if (a == b) tr += s2;
          }
        const double jl = J * lambda, jm = J * mu;
#pragma unroll
        for (int b = 0; b < N; ++b)
#pragma unroll
          for (int a = 0; a < N; ++a)
            acc[b * Q + a] += jl * T[a + N * b] + jm * T[b + N * a] + (a == b ? jm * tr : 0.0);
      } else if (RF == RF_LAPLACE) {
        double s2 = 0;
#pragma unroll
        for (int k = 0; k < MT; ++k) s2 += M[k] * g[k];
        acc[0] += s2;
      } else {
        acc[0] += M[0] * g[0];
      }
    }
    const unsigned m = pmask[p];
    const int32_t J = pJ[p];
#pragma unroll
    for (int b = 0; b < Q; ++b) {
      int64_t pos = jc[J + b] + prel[(size_t)b * npairs + p];
#pragma unroll
      for (int a = 0; a < Q; ++a)
        if (m & (1u << (b * Q + a))) pr[pos++] = RF == RF_ELAST ? acc[RF == RF_ELAST ? b * Q + a : 0] : (a == b ? acc[0] : 0.0);
    }
  }
}

static int rf_of(int family) {
  return family == GFGPU_LAPLACE ? RF_LAPLACE : family == GFGPU_ELASTICITY ? RF_ELAST : family == GFGPU_MASS ? RF_MASS : -1;
}

bool recompute_supported(const gfgpu_term *t) {
  if (t->mesh->gt_kind != GFGPU_GT_PK) return false;
  if (rf_of(t->family) < 0) return false;
  const int N = t->mesh->dim, nd = t->fem->nd, Q = t->fem->qdim;
  const bool ndok = N == 2 ? (nd == 3 || nd == 6 || nd == 10) : (nd == 4 || nd == 10 || nd == 20);
  if (!ndok) return false;
  if (t->family == GFGPU_ELASTICITY) return Q == N;
  return Q == 1 || Q == N;
}

// Host: reference tensors from the staged tables; device: per-element geometry.
void recompute_prepare(gfgpu_term *t) {
  gfgpu_ctx *ctx = t->ctx;
  const int N = t->mesh->dim, nd = t->fem->nd, nq = t->tab->nq;
  const int rf = rf_of(t->family);
  const int MT = rf == RF_ELAST ? N * N : rf == RF_LAPLACE ? N * (N + 1) / 2 : 1;
  const int EG = rf == RF_ELAST ? N * N + 1 : rf == RF_LAPLACE ? N * (N + 1) / 2 : 1;
  const std::vector<double> &w = t->tab->h_w, &g = t->tab->h_gphi, &ph = t->tab->h_phi;
  std::vector<double> M((size_t)nd * nd * MT, 0.0);
  for (int j = 0; j < nd; ++j)
    for (int i = 0; i < nd; ++i) {
      double *o = M.data() + ((size_t)j * nd + i) * MT;
      if (rf == RF_MASS) {
        double s = 0;
        for (int k = 0; k < nq; ++k) s += w[k] * ph[(size_t)k * nd + i] * ph[(size_t)k * nd + j];
        o[0] = s;
        continue;
      }
      double full[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int k = 0; k < nq; ++k) {
        if (w[k] == 0.0) continue;
        const double *gi = g.data() + ((size_t)k * nd + i) * N, *gj = g.data() + ((size_t)k * nd + j) * N;
        for (int p = 0; p < N; ++p)
          for (int q = 0; q < N; ++q) full[p * N + q] += w[k] * gi[p] * gj[q];
      }
      if (rf == RF_ELAST) {
        for (int k = 0; k < N * N; ++k) o[k] = full[k];
      } else {  // symmetrised against G = a J B^T B (upper triangle, row by row)
        int k = 0;
        for (int p = 0; p < N; ++p)
          for (int q = p; q < N; ++q) o[k++] = p == q ? full[p * N + p] : full[p * N + q] + full[q * N + p];
      }
    }
  t->rc_M.alloc(ctx, M.size());
  t->rc_M.upload(M.data());
  const int64_t ne = t->e1 - t->e0;
  t->rc_eg.alloc(ctx, (size_t)ne * EG);
  const double scale = t->alpha * ((rf == RF_ELAST) ? 1.0 : t->par[0]);
  const int64_t np = t->mesh->npts;
  const double *x = t->mesh->xyz.p;
  if (ne) {
    int grid = (int)std::min<int64_t>((ne + 255) / 256, 148 * 16);
    if (N == 2)
      k_affine_geo<2><<<grid, 256, 0, ctx->stream>>>(x, x + np, x + 2 * np, t->mesh->conn.p, t->mesh->ng,
                                                     t->tab->gt_grad.p, t->e0, ne, rf, scale, t->rc_eg.p, EG);
    else
      k_affine_geo<3><<<grid, 256, 0, ctx->stream>>>(x, x + np, x + 2 * np, t->mesh->conn.p, t->mesh->ng,
                                                     t->tab->gt_grad.p, t->e0, ne, rf, scale, t->rc_eg.p, EG);
    GF_LAUNCH_CHECK();
  }
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  t->rc_ready = true;
}

template <int N, int Q, int ND, int RF>
static void launch_pairs(gfgpu_term *t) {
  Structure &st = t->st;
  int grid = (int)std::min<int64_t>((st.npairs + 255) / 256, 148 * 64);
  k_recompute_pairs<N, Q, ND, RF><<<grid, 256, 0, t->ctx->stream>>>(
      st.cstart.p, st.csrc.p, st.pJ.p, t->pmask.p, t->prel.p, t->jc.p, t->rc_eg.p, t->rc_M.p, t->par[0], t->par[1],
      st.npairs, t->pr.p);
  GF_LAUNCH_CHECK();
}

#define RC_CASE(NN, QQ, NDD, RFF)                                   \
  if (N == NN && Q == QQ && nd == NDD && rf == RFF) {               \
    launch_pairs<NN, QQ, NDD, RFF>(t);                              \
    return;                                                         \
  }

void recompute_tangent(gfgpu_term *t) {
  if (!t->st.npairs) return;
  const int N = t->mesh->dim, nd = t->fem->nd, Q = t->fem->qdim, rf = rf_of(t->family);
  RC_CASE(3, 3, 10, RF_ELAST) RC_CASE(3, 3, 4, RF_ELAST) RC_CASE(3, 3, 20, RF_ELAST)
  RC_CASE(3, 1, 10, RF_LAPLACE) RC_CASE(3, 1, 4, RF_LAPLACE) RC_CASE(3, 1, 20, RF_LAPLACE)
  RC_CASE(3, 3, 10, RF_LAPLACE) RC_CASE(3, 3, 4, RF_LAPLACE) RC_CASE(3, 3, 20, RF_LAPLACE)
  RC_CASE(3, 1, 10, RF_MASS) RC_CASE(3, 1, 4, RF_MASS) RC_CASE(3, 1, 20, RF_MASS)
  RC_CASE(3, 3, 10, RF_MASS) RC_CASE(3, 3, 4, RF_MASS) RC_CASE(3, 3, 20, RF_MASS)
  RC_CASE(2, 2, 3, RF_ELAST) RC_CASE(2, 2, 6, RF_ELAST) RC_CASE(2, 2, 10, RF_ELAST)
  RC_CASE(2, 1, 3, RF_LAPLACE) RC_CASE(2, 1, 6, RF_LAPLACE) RC_CASE(2, 1, 10, RF_LAPLACE)
  RC_CASE(2, 2, 3, RF_LAPLACE) RC_CASE(2, 2, 6, RF_LAPLACE) RC_CASE(2, 2, 10, RF_LAPLACE)
  RC_CASE(2, 1, 3, RF_MASS) RC_CASE(2, 1, 6, RF_MASS) RC_CASE(2, 1, 10, RF_MASS)
  RC_CASE(2, 2, 3, RF_MASS) RC_CASE(2, 2, 6, RF_MASS) RC_CASE(2, 2, 10, RF_MASS)
  GF_REQUIRE(false, "no recompute kernel for this combination");
}

}  // namespace gf
