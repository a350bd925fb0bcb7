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

// One warp per PACKET of whole column nodes (<= 32 element incidences, packed greedily on the host).
//   stage A  lane = (element, local column node j) incidence: copies the element geometry (B, J) into the
//            warp's shared memory once -- nd blocks reuse it; the packet's pair boundaries go to smem too;
//   stage B  the packet's contributions are in (row node, element) order; every lane takes an EQUAL run of
//            consecutive contributions (same work whatever the node valences).  A 16-bit descriptor
//            (incidence slot, j, i) selects geometry and reference tensor; the QxQ block is accumulated in
//            registers while the node pair stays the same and flushed when it changes: kept entries go to
//            their CSC slots, and (residual) the pair's part of K^T U to shared memory.  A pair cut by a
//            run boundary leaves a HEAD part in shared memory; the lane where the pair starts adds the
//            following head parts in lane order (= ascending element id, the reference's order);
//   stage C  lane = (column node, component): fixed-order sum of the pair parts -> R.
// No atomics, fixed summation order: bitwise reproducible.
struct ColArgs {
  const uint32_t *wcol;      // nw+1: packet w owns column nodes [wcol[w], wcol[w+1])
  const int64_t *wbase;      // nw: CSC position of the first entry of the packet
  const uint32_t *rstart, *rsrc;
  const int32_t *rdof;
  const uint32_t *colstart, *cstart;
  const uint16_t *cdesc;     // per contribution: slot << 10 | j << 5 | i
  const uint4 *prec;         // per pair: x = CSC offset of column component 0 (relative to wbase),
                             //   y = (offset of component 1 - x) | keep mask << 20, z = offset of component 2 - x, w = row dof
  const double *eg, *Mtab, *U;
  double lambda, mu;
  int64_t nw, npairs;
  int cap_inc, cap_pairs;
  double *pr, *R;
};

template <int N, int Q, int ND, int RF>
struct ColSmem {
  using C = RcCfg<N, RF>;
  static constexpr int ACC = RF == RF_ELAST ? Q * Q : 1;
  static constexpr int EGP = C::EG | 1;  // odd stride (doubles): conflict-free geometry rows
  static constexpr int HS = ACC | 1;     // odd stride of the head parts
  // doubles per warp: geometry, head parts, residual parts, then (as uint32) pair boundaries + head pair ids
  __host__ __device__ static size_t per_warp(int cap_inc, int cap_pairs, bool do_r) {
    size_t d = (size_t)cap_inc * EGP + 32 * HS + (do_r ? (size_t)cap_pairs * Q : 0);
    size_t u = (size_t)cap_pairs + 1 + 32;
    return d + (u + 1) / 2;
  }
};

template <int N, int Q, int ND, int RF, bool DO_T, bool DO_R>
__global__ void __launch_bounds__(256)
k_recompute_cols(const ColArgs a) {
  using C = RcCfg<N, RF>;
  using S = ColSmem<N, Q, ND, RF>;
  constexpr int NB = ND * ND, MT = C::MT, EG = C::EG, ACC = S::ACC, EGP = S::EGP, HS = S::HS;
  extern __shared__ double sm[];
  double *sM = sm;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, WPB = blockDim.x >> 5;
  double *sG = sM + NB * MT + (size_t)warp * S::per_warp(a.cap_inc, a.cap_pairs, DO_R);
  double *sH = sG + (size_t)a.cap_inc * EGP;          // head parts, one per lane
  double *rp = sH + 32 * HS;                          // residual parts, one per pair
  uint32_t *scs = (uint32_t *)(rp + (DO_R ? (size_t)a.cap_pairs * Q : 0));  // pair boundaries (relative)
  uint32_t *shid = scs + a.cap_pairs + 1;             // pair (relative) of each lane's head part, or ~0
  for (int k = threadIdx.x; k < NB * MT; k += blockDim.x) sM[k] = a.Mtab[k];
  __syncthreads();
  for (int64_t w = (int64_t)blockIdx.x * WPB + warp; w < a.nw; w += (int64_t)gridDim.x * WPB) {
    const uint32_t k0 = a.wcol[w], k1 = a.wcol[w + 1];
    const uint32_t r0 = a.rstart[k0], r1 = a.rstart[k1];
    const uint32_t p0 = a.colstart[k0], npk = a.colstart[k1] - p0;
    const uint32_t it0 = a.cstart[p0];
    const int64_t base = DO_T ? a.wbase[w] : 0;
    // ---- stage A (EG/2 lanes per geometry row when EG is even: few 128-byte lines per load instruction)
    if (EG % 2 == 0) {
      constexpr int PARTS = EG / 2 > 0 ? EG / 2 : 1;
      for (uint32_t idx = lane; idx < (r1 - r0) * PARTS; idx += 32) {
        const uint32_t row = idx / PARTS, part = idx - row * PARTS;
        const uint32_t el = a.rsrc[r0 + row] / ND;
        const double2 v = *reinterpret_cast<const double2 *>(a.eg + (size_t)el * EG + 2 * part);
        double *o = sG + (size_t)row * EGP + 2 * part;
        o[0] = v.x;
        o[1] = v.y;
      }
    } else {
      for (uint32_t ri = r0 + lane; ri < r1; ri += 32) {
        const uint32_t el = a.rsrc[ri] / ND;
        const double *g = a.eg + (size_t)el * EG;
        double *o = sG + (size_t)(ri - r0) * EGP;
#pragma unroll
        for (int k = 0; k < EG; ++k) o[k] = g[k];
      }
    }
    for (uint32_t q = lane; q <= npk; q += 32) scs[q] = a.cstart[p0 + q] - it0;
    shid[lane] = 0xffffffffu;
    __syncwarp();
    const uint32_t nit = scs[npk];
    // ---- stage B: equal runs of consecutive contributions
    const uint32_t L = (nit + 31) >> 5;
    const uint32_t tb = min(lane * L, nit), te = min(tb + L, nit);
    double acc[ACC];
#pragma unroll
    for (int m = 0; m < ACC; ++m) acc[m] = 0.0;
    // Pair metadata is fetched into registers when the lane STARTS a pair it owns, the U values one
    // contribution later (their address depends on the row node): by the time the pair is flushed the
    // loads have landed instead of stalling the flush.
    uint4 rec = make_uint4(0, 0, 0, 0);  // kept packed until the flush: unpacking would wait for the load
    double u[Q];
    bool need_u = false;
#pragma unroll
    for (int b = 0; b < Q; ++b) u[b] = 0.0;
    auto fetch_meta = [&](uint32_t qq) {
      rec = a.prec[p0 + qq];  // one 16-byte load
      need_u = DO_R;
    };
    auto fetch_u = [&]() {
#pragma unroll
      for (int aa = 0; aa < Q; ++aa) u[aa] = a.U ? a.U[(int32_t)rec.w + aa] : 0.0;
      need_u = false;
    };
    // flush of a COMPLETE pair sum held in acc
    auto flush = [&](uint32_t qq) {
      if (DO_T) {
        const uint32_t po[3] = {rec.x, rec.x + (rec.y & 0xfffffu), rec.x + rec.z};
        const unsigned pm = rec.y >> 20;
#pragma unroll
        for (int b = 0; b < Q; ++b) {
          double *dst = a.pr + (base + po[b]);
          const unsigned mb = (pm >> (b * Q)) & ((1u << Q) - 1);
#pragma unroll
          for (int aa = 0; aa < Q; ++aa)
            if (mb & (1u << aa))
              dst[__popc(mb & ((1u << aa) - 1))] =
                  RF == RF_ELAST ? acc[RF == RF_ELAST ? b * Q + aa : 0] : (aa == b ? acc[0] : 0.0);
        }
      }
      if (DO_R) {
        if (need_u) fetch_u();
#pragma unroll
        for (int b = 0; b < Q; ++b) {
          double s2 = 0;
          if (RF == RF_ELAST) {
#pragma unroll
            for (int aa = 0; aa < Q; ++aa) s2 += acc[RF == RF_ELAST ? b * Q + aa : 0] * u[aa];
          } else {
            s2 = acc[0] * u[b];
          }
          rp[(size_t)qq * Q + b] = s2;
        }
      }
    };
    uint32_t q = 0;                // current pair (relative)
    uint32_t tailq = 0xffffffffu;  // pair whose first part ends my run and continues in the next lanes
    unsigned dnext = 0;
    if (tb < te) {
      // pair containing contribution tb: largest q with scs[q] <= tb
      uint32_t lo = 0, hi = npk;
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (scs[mid] <= tb) lo = mid; else hi = mid;
      }
      q = lo;
      dnext = a.cdesc[(size_t)it0 + tb];
      if (scs[q] >= tb) fetch_meta(q);
    }
    uint32_t qend = (tb < te) ? scs[q + 1] : 0;
    for (uint32_t t = tb; t < te; ++t) {
      const unsigned d = dnext;
      if (t + 1 < te) dnext = a.cdesc[(size_t)it0 + t + 1];
      if (DO_R && need_u && t > tb) fetch_u();
      const int i = d & 31, j = (d >> 5) & 31;
      const double *G = sG + (size_t)(d >> 10) * EGP;
      const double *M = sM + (j * ND + i) * MT;
      if (RF == RF_ELAST) {
        double Bm[N * N], W[N * N];
#pragma unroll
        for (int k = 0; k < N * N; ++k) Bm[k] = G[k];
#pragma unroll
        for (int qq = 0; qq < N; ++qq)
#pragma unroll
          for (int aa = 0; aa < N; ++aa) {
            double s2 = 0;
#pragma unroll
            for (int pp = 0; pp < N; ++pp) s2 += Bm[aa + N * pp] * M[pp * N + qq];
            W[aa + N * qq] = s2;
          }
        double T[N * N], tr = 0;
#pragma unroll
        for (int b = 0; b < N; ++b)
#pragma unroll
          for (int aa = 0; aa < N; ++aa) {
            double s2 = 0;
#pragma unroll
            for (int qq = 0; qq < N; ++qq) s2 += W[aa + N * qq] * Bm[b + N * qq];
            T[aa + N * b] = s2;
            if (aa == b) tr += s2;
          }
        const double jl = G[N * N] * a.lambda, jm = G[N * N] * a.mu;
#pragma unroll
        for (int b = 0; b < N; ++b)
#pragma unroll
          for (int aa = 0; aa < N; ++aa)
            acc[RF == RF_ELAST ? b * Q + aa : 0] += jl * T[aa + N * b] + jm * T[b + N * aa] + (aa == b ? jm * tr : 0.0);
      } else if (RF == RF_LAPLACE) {
        double s2 = 0;
#pragma unroll
        for (int k = 0; k < MT; ++k) s2 += M[k] * G[k];
        acc[0] += s2;
      } else {
        acc[0] += M[0] * G[0];
      }
      if (t + 1 == qend || t + 1 == te) {  // the pair, or my run, ends here
        const bool started_here = scs[q] >= tb, ends_here = qend <= te;
        if (started_here && ends_here) {
          flush(q);
        } else if (!started_here) {  // head part of a pair that began in an earlier lane
          shid[lane] = q;
#pragma unroll
          for (int m = 0; m < ACC; ++m) sH[lane * HS + m] = acc[m];
        } else {  // began here, continues: I own it (its metadata stays in registers), the rest arrives as head parts
          tailq = q;
        }
        if (t + 1 < te) {
#pragma unroll
          for (int m = 0; m < ACC; ++m) acc[m] = 0.0;
          ++q;
          qend = scs[q + 1];
          fetch_meta(q);
        }
      }
    }
    __syncwarp();
    if (tailq != 0xffffffffu) {  // acc still holds my (first) part of the pair
      for (int l2 = lane + 1; l2 < 32 && shid[l2] == tailq; ++l2) {
#pragma unroll
        for (int m = 0; m < ACC; ++m) acc[m] += sH[l2 * HS + m];
      }
      flush(tailq);
    }
    __syncwarp();
    // ---- stage C
    if (DO_R) {
      for (uint32_t idx = lane; idx < (k1 - k0) * Q; idx += 32) {
        const uint32_t kc = k0 + idx / Q, b = idx % Q;
        double s2 = 0;
        for (uint32_t pp = a.colstart[kc] - p0, pe = a.colstart[kc + 1] - p0; pp < pe; ++pp) s2 += rp[(size_t)pp * Q + b];
        a.R[a.rdof[kc] + b] = s2;
      }
      __syncwarp();
    }
  }
}

// ---- one-time plan kernels
__global__ void k_invert_perm(const uint32_t *__restrict__ src, int64_t n, uint32_t *__restrict__ pos) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x)
    pos[src[s]] = (uint32_t)s;
}

__global__ void k_col_packet(const uint32_t *__restrict__ wcol, int64_t nw, uint32_t *__restrict__ colw) {
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < nw; w += (int64_t)gridDim.x * blockDim.x)
    for (uint32_t k = wcol[w]; k < wcol[w + 1]; ++k) colw[k] = (uint32_t)w;
}

// thread per column node: contribution descriptors and packet-relative CSC offsets of its pairs
template <int Q>
__global__ void k_col_plan(const uint32_t *__restrict__ wcol, const uint32_t *__restrict__ colw,
                           const uint32_t *__restrict__ colstart, const uint32_t *__restrict__ cstart,
                           const uint32_t *__restrict__ csrc, const uint32_t *__restrict__ rstart,
                           const uint32_t *__restrict__ rpos, const int32_t *__restrict__ pJ,
                           const int64_t *__restrict__ jc, const uint32_t *__restrict__ prel, int nd, int64_t ncol,
                           int64_t npairs, const int32_t *__restrict__ pI, const uint16_t *__restrict__ pmask,
                           uint16_t *__restrict__ cdesc, uint4 *__restrict__ prec, int64_t *__restrict__ wbase,
                           int *__restrict__ err) {
  const int nb = nd * nd;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < ncol; k += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t w = colw[k], kf = wcol[w];
    const uint32_t rbase = rstart[kf];
    const int64_t base = jc[pJ[colstart[kf]]];
    if (k == kf) wbase[w] = base;
    for (uint32_t p = colstart[k]; p < colstart[k + 1]; ++p) {
      const int32_t J = pJ[p];
      int64_t off[3] = {0, 0, 0};
      for (int b = 0; b < Q; ++b) off[b] = jc[J + b] + prel[(size_t)b * npairs + p] - base;
      const int64_t d1 = off[1] - off[0], d2 = off[2] - off[0];
      if (off[0] < 0 || off[0] >= (int64_t(1) << 32) || (Q > 1 && (d1 < 0 || d1 >= (1 << 20))) ||
          (Q > 2 && (d2 < 0 || d2 >= (int64_t(1) << 32))) || pmask[p] >= (1u << 12))
        *err = 1;
      prec[p] = make_uint4((uint32_t)off[0], (uint32_t)(Q > 1 ? d1 : 0) | ((uint32_t)pmask[p] << 20),
                           (uint32_t)(Q > 2 ? d2 : 0), (uint32_t)pI[p]);
      for (uint32_t s = cstart[p]; s < cstart[p + 1]; ++s) {
        const uint32_t c = csrc[s];
        const uint32_t el = c / nb, r = c - el * nb;
        const uint32_t j = r / nd, i = r - j * nd;
        const uint32_t slot = rpos[el * nd + j] - rbase;
        if (slot > 63u) *err = 2;
        cdesc[s] = (uint16_t)((slot << 10) | (j << 5) | i);
      }
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
  // ---- column packets: whole column nodes packed greedily up to 32 element incidences per warp
  Structure &st = t->st;
  GF_REQUIRE(st.ncolnodes == st.nrnodes, "column nodes and incidence nodes differ");
  GF_REQUIRE(nd <= 32, "too many local nodes for the 16-bit contribution descriptor");
  std::vector<uint32_t> rstart(st.nrnodes + 1), colstart(st.ncolnodes + 1);
  st.rstart.download(rstart.data());
  st.colstart.download(colstart.data());
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<uint32_t> wcol;
  wcol.reserve(st.ncolnodes / 2 + 2);
  int cap_inc = 1, cap_pairs = 1;
  for (int64_t k = 0; k < st.ncolnodes;) {
    wcol.push_back((uint32_t)k);
    int64_t k1 = k + 1;
    while (k1 < st.ncolnodes && rstart[k1 + 1] - rstart[k] <= 32) ++k1;
    cap_inc = std::max<int>(cap_inc, (int)(rstart[k1] - rstart[k]));
    cap_pairs = std::max<int>(cap_pairs, (int)(colstart[k1] - colstart[k]));
    k = k1;
  }
  wcol.push_back((uint32_t)st.ncolnodes);
  GF_REQUIRE(cap_inc <= 64, "node valence above 64 elements: use strategy STAGED");
  t->rc_nw = (int64_t)wcol.size() - 1;
  t->rc_cap_inc = cap_inc;
  t->rc_cap_pairs = cap_pairs;
  t->rc_wcol.alloc(ctx, wcol.size());
  t->rc_wcol.upload(wcol.data());
  t->rc_wbase.alloc(ctx, t->rc_nw);
  t->rc_cdesc.alloc(ctx, st.ncontrib);
  t->rc_prec.alloc(ctx, st.npairs);
  {
    DevBuf<uint32_t> colw, rpos;
    colw.alloc(ctx, st.ncolnodes);
    rpos.alloc(ctx, st.nrinc);
    t->flag.zero();
    const int B = 256;
    auto grid = [&](int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + B - 1) / B, 148 * 32)); };
    k_invert_perm<<<grid(st.nrinc), B, 0, ctx->stream>>>(st.rsrc.p, st.nrinc, rpos.p);
    GF_LAUNCH_CHECK();
    k_col_packet<<<grid(t->rc_nw), B, 0, ctx->stream>>>(t->rc_wcol.p, t->rc_nw, colw.p);
    GF_LAUNCH_CHECK();
#define GF_PLAN(QQ)                                                                                               \
  k_col_plan<QQ><<<grid(st.ncolnodes), B, 0, ctx->stream>>>(t->rc_wcol.p, colw.p, st.colstart.p, st.cstart.p,      \
                                                            st.csrc.p, st.rstart.p, rpos.p, st.pJ.p, t->jc.p,      \
                                                            t->prel.p, nd, st.ncolnodes, st.npairs, st.pI.p,       \
                                                            t->pmask.p, t->rc_cdesc.p, (uint4 *)t->rc_prec.p,      \
                                                            t->rc_wbase.p, (int *)t->flag.p)
    if (t->fem->qdim == 1) GF_PLAN(1);
    else if (t->fem->qdim == 2) GF_PLAN(2);
    else GF_PLAN(3);
#undef GF_PLAN
    GF_LAUNCH_CHECK();
    int32_t err = 0;
    t->flag.download(&err);
    GF_CUDA(cudaStreamSynchronize(ctx->stream));
    GF_REQUIRE(err == 0, "recompute plan: packet too large for the compact descriptors");
  }
  // the contribution list and the column-relative offsets are folded into cdesc / poff
  t->prel.release();
  t->rc_ready = true;
}

template <int N, int Q, int ND, int RF>
static void launch_cols(gfgpu_term *t, const double *U, bool do_t, bool do_r) {
  using C = RcCfg<N, RF>;
  Structure &st = t->st;
  ColArgs a;
  a.wcol = t->rc_wcol.p; a.wbase = t->rc_wbase.p; a.rstart = st.rstart.p; a.rsrc = st.rsrc.p; a.rdof = st.rdof.p;
  a.colstart = st.colstart.p; a.cstart = st.cstart.p; a.cdesc = t->rc_cdesc.p;
  a.prec = (const uint4 *)t->rc_prec.p;
  a.eg = t->rc_eg.p; a.Mtab = t->rc_M.p; a.U = U;
  a.lambda = t->par[0]; a.mu = t->par[1];
  a.nw = t->rc_nw; a.npairs = st.npairs;
  a.cap_inc = t->rc_cap_inc; a.cap_pairs = t->rc_cap_pairs;
  a.pr = t->pr.p; a.R = t->R.p;
  const int WPB = 8;
  const size_t smem = ((size_t)ND * ND * C::MT + (size_t)WPB * ColSmem<N, Q, ND, RF>::per_warp(a.cap_inc, a.cap_pairs, do_r)) * 8;
  GF_REQUIRE(smem <= 220 * 1024, "packet too large for shared memory; use strategy STAGED");
  auto launch = [&](auto kern) {
    GF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    GF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WPB * 32, smem));
    if (occ < 1) occ = 1;
    int64_t want = (a.nw + WPB - 1) / WPB;
    int grid = (int)std::min<int64_t>(want, (int64_t)t->ctx->sm_count * occ);
    kern<<<grid, WPB * 32, smem, t->ctx->stream>>>(a);
    GF_LAUNCH_CHECK();
  };
  if (do_t && do_r) launch(k_recompute_cols<N, Q, ND, RF, true, true>);
  else if (do_t) launch(k_recompute_cols<N, Q, ND, RF, true, false>);
  else launch(k_recompute_cols<N, Q, ND, RF, false, true>);
}

#define RC_CASE(NN, QQ, NDD, RFF)                                   \
  if (N == NN && Q == QQ && nd == NDD && rf == RFF) {               \
    launch_cols<NN, QQ, NDD, RFF>(t, U, do_t, do_r);                \
    return;                                                         \
  }

// tangent and/or residual (R = K^T U = K U, the handled forms are symmetric) in one kernel
void recompute_assemble(gfgpu_term *t, const double *U, bool do_t, bool do_r) {
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
