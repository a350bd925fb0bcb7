// elem_kernel.cuh -- generic per-element weak-form kernel (strategy STAGED).
//
// One element = one slot of TPE threads; a CTA carries EPB slots.  Per element:
//   A  gather node coordinates G (SoA -> smem) and local coefficients of u
//        (basic_mesh::points_of_convex bgeot_mesh.h:94; slice_vector_on_basic_dof_of_element
//         getfem_mesh_fem.h:662-689)
//   B1 K = G*pc, J = |det K|, B = K^-T at each Gauss point (once if the transformation is affine)
//        (bgeot_geometric_trans.cc:270-288, 321-355, 374-413)
//   B2 Z(i,n) = sum_p gradhat(i,p) B(n,p)                 (getfem_fem.cc:85-90, 160-168)
//   B3 Grad_u = sum_i u_i Z(i,.)                           (C&E.cc:692-747)
//   B4 material point: flux P(alpha,n), tangent D(alpha,n,beta,l), both scaled by alpha*J*w_q
//        (C&E.cc:8850-8851; getfem_models.cc:6112-6113; getfem_nonlinear_elasticity.cc:612-702,
//         1781-1827, 1945-1994)
//   C  K_e(i a, j b) += Z(i,n) D(a,n,b,l) Z(j,l)  in registers; r_e(i a) += P(a,n) Z(i,n)
//   D  ninf = max|K_e|, entries <= 1e-14*ninf are not inserted (C&E.cc:4889,4898; 5380-5402):
//        written as 0 to the stage and flagged 0 in the per-block keep mask.
// Boundary-face items (mesh regions of faces, C&E.cc:8827-8848): the tables are those of the face's points, the weight
// is J |B n_ref| w_q and the unit normal B n_ref / |B n_ref| (components below 1e-13 cleaned) is kept beside B and J.
// The Gauss points are processed in chunks of qc so that any (nd, nq) fits shared memory.
#pragma once
#include "common.cuh"

namespace gf {

enum { FK_LAPLACE = 0, FK_ELAST = 1, FK_HYPER = 2, FK_MASS = 3 };

template <int DIM, int Q, int ND, int FK, bool AFFINE>
struct ElemCfg {
  static constexpr int N = DIM, S1 = ND * Q, NB = ND * ND;
  static constexpr int TPE = NB <= 16 ? 16 : NB <= 32 ? 32 : NB <= 64 ? 64 : NB <= 128 ? 128 : 256;
  static constexpr int THREADS = TPE < 128 ? 128 : TPE;
  static constexpr int EPB = THREADS / TPE;
  static constexpr int BPT = (NB + TPE - 1) / TPE;
  static constexpr int RPT = (S1 + TPE - 1) / TPE;
  static constexpr bool SCALAR = (FK == FK_LAPLACE || FK == FK_MASS);
  static constexpr int DSZ = FK == FK_HYPER ? (Q * N * Q * N) : 1;
  static constexpr int ACC = SCALAR ? 1 : Q * Q;
  static constexpr int GEO = N * N + 1 + N;  // B, J, unit normal (faces)
  __host__ __device__ static constexpr int per_q() { return GEO + ND * N + Q * N + DSZ + Q * N + 2; }  // + (lambda, mu) at the point
  __host__ __device__ static int slot_doubles(int ng, int qc) { return N * ng + S1 + GEO + qc * per_q() + 8; }
};

// K col-major N x N -> B = K^{-T} (col-major), returns |det K|
template <int N>
__device__ __forceinline__ double inv_transpose(const double *K, double *B) {
  if (N == 2) {
    double d = K[0] * K[3] - K[1] * K[2];
    double id = 1.0 / d;
    // Ki = [K3 -K2; -K1 K0]/d (col-major Ki[0]=K3, Ki[1]=-K1, Ki[2]=-K2, Ki[3]=K0); B = Ki^T
    B[0] = K[3] * id; B[2] = -K[1] * id; B[1] = -K[2] * id; B[3] = K[0] * id;
    return fabs(d);
  } else {
#define K_(i, j) K[(i) + 3 * (j)]
    double c00 = K_(1, 1) * K_(2, 2) - K_(1, 2) * K_(2, 1);
    double c10 = K_(1, 2) * K_(2, 0) - K_(1, 0) * K_(2, 2);
    double c20 = K_(1, 0) * K_(2, 1) - K_(1, 1) * K_(2, 0);
    double d = K_(0, 0) * c00 + K_(0, 1) * c10 + K_(0, 2) * c20;
    double id = 1.0 / d;
    // Ki(i,j) = cof(j,i)/d ; B(i,j) = Ki(j,i) = cof(i,j)/d
    B[0 + 3 * 0] = c00 * id;
    B[0 + 3 * 1] = c10 * id;
    B[0 + 3 * 2] = c20 * id;
    B[1 + 3 * 0] = (K_(0, 2) * K_(2, 1) - K_(0, 1) * K_(2, 2)) * id;
    B[1 + 3 * 1] = (K_(0, 0) * K_(2, 2) - K_(0, 2) * K_(2, 0)) * id;
    B[1 + 3 * 2] = (K_(0, 1) * K_(2, 0) - K_(0, 0) * K_(2, 1)) * id;
    B[2 + 3 * 0] = (K_(0, 1) * K_(1, 2) - K_(0, 2) * K_(1, 1)) * id;
    B[2 + 3 * 1] = (K_(0, 2) * K_(1, 0) - K_(0, 0) * K_(1, 2)) * id;
    B[2 + 3 * 2] = (K_(0, 0) * K_(1, 1) - K_(0, 1) * K_(1, 0)) * id;
#undef K_
    return fabs(d);
  }
}

// geo[0..N*N) = B (col-major), geo[N*N] = J, from G (N x ng col-major) and pc (ng x N row-major).
// nref != nullptr (boundary face): Normal = B n_ref, J *= |Normal|, Normal /= |Normal|, gmm::clean(Normal, 1e-13)
// (C&E.cc:8836-8847) -> geo[N*N+1 ..).
template <int N>
__device__ __forceinline__ void geometry(const double *G, const double *pc, int ng, double *geo,
                                         const double *nref = nullptr) {
  double K[N * N];
#pragma unroll
  for (int k = 0; k < N * N; ++k) K[k] = 0.0;
  for (int i = 0; i < ng; ++i) {
#pragma unroll
    for (int c = 0; c < N; ++c) {
      double p = pc[i * N + c];
#pragma unroll
      for (int r = 0; r < N; ++r) K[r + N * c] += G[r + N * i] * p;
    }
  }
  double B[N * N];
  double J = inv_transpose<N>(K, B);
#pragma unroll
  for (int k = 0; k < N * N; ++k) geo[k] = B[k];
  if (nref) {
    double nr[N], nup = 0.0;
#pragma unroll
    for (int r = 0; r < N; ++r) {
      double s = 0;
#pragma unroll
      for (int c = 0; c < N; ++c) s += B[r + N * c] * nref[c];
      nr[r] = s;
      nup += s * s;
    }
    nup = sqrt(nup);
    J *= nup;
#pragma unroll
    for (int r = 0; r < N; ++r) {
      const double v = nr[r] / nup;
      geo[N * N + 1 + r] = fabs(v) < 1e-13 ? 0.0 : v;
    }
  }
  geo[N * N] = J;
}

__device__ __forceinline__ double det3cm(const double *A) {
  return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[3] * (A[1] * A[8] - A[2] * A[7]) + A[6] * (A[1] * A[5] - A[2] * A[4]);
}

// Hyperelastic material point (N = Q = 3), in two parts so that callers can spread the nine (l, n) slices of the
// tangent over threads (sumfact.cu) or run them in sequence (hyper_point below).  Gu(a,n) col-major.
//   S = PK2(E), dS(i,j,k,l) = dS_ij/dGu_kl ; P = F S ; D(a,n,b,l) = delta_ab S(l,n) + F(a,p) dS(p,n,b,l)
struct HyperPrep {
  double Gu[9], F[9], S[9], Ci[9], di3[9];
  double c1, c2, hd;
};

__device__ __forceinline__ void hyper_prep(int law, const double *Gu, double lambda, double mu, HyperPrep &h) {
  double E[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double s = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) s += Gu[k + 3 * i] * Gu[k + 3 * j];
      E[i + 3 * j] = 0.5 * (s + Gu[i + 3 * j] + Gu[j + 3 * i]);
      h.F[i + 3 * j] = Gu[i + 3 * j] + (i == j ? 1.0 : 0.0);
      h.Gu[i + 3 * j] = Gu[i + 3 * j];
    }
  h.c1 = h.c2 = h.hd = 0.0;
  if (law == GFGPU_SVK) {  // getfem_nonlinear_elasticity.cc:1945-1994
    const double trE = E[0] + E[4] + E[8];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        h.S[i + 3 * j] = 2 * mu * E[i + 3 * j] + (i == j ? lambda * trE : 0.0);
        h.Ci[i + 3 * j] = 0.0;
        h.di3[i + 3 * j] = 0.0;
      }
    // (the registered GWFL operator has no det F penalty, unlike the Neo-Hookean law below)
    return;
  }
  // Neo_Hookean_hyperelastic_law (:612-702) through AHL_wrapper_sigma (:1781-1827)
  const bool bonet = law == GFGPU_NEOHOOKEAN_BONET;
  const double detF = det3cm(h.F);
  double C[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) C[i] = 2 * E[i];
  C[0] += 1; C[4] += 1; C[8] += 1;
  const double i3 = det3cm(C);
  {
#define C_(i, j) C[(i) + 3 * (j)]
    h.Ci[0] = (C_(1, 1) * C_(2, 2) - C_(1, 2) * C_(2, 1)) / i3;
    h.Ci[1] = -(C_(1, 0) * C_(2, 2) - C_(1, 2) * C_(2, 0)) / i3;
    h.Ci[2] = (C_(1, 0) * C_(2, 1) - C_(1, 1) * C_(2, 0)) / i3;
    h.Ci[3] = -(C_(0, 1) * C_(2, 2) - C_(0, 2) * C_(2, 1)) / i3;
    h.Ci[4] = (C_(0, 0) * C_(2, 2) - C_(0, 2) * C_(2, 0)) / i3;
    h.Ci[5] = -(C_(0, 0) * C_(2, 1) - C_(0, 1) * C_(2, 0)) / i3;
    h.Ci[6] = (C_(0, 1) * C_(1, 2) - C_(0, 2) * C_(1, 1)) / i3;
    h.Ci[7] = -(C_(0, 0) * C_(1, 2) - C_(0, 2) * C_(1, 0)) / i3;
    h.Ci[8] = (C_(0, 0) * C_(1, 1) - C_(0, 1) * C_(1, 0)) / i3;
#undef C_
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) h.di3[i] = h.Ci[i] * i3;  // compute_di3 (:132-140)
  const double lg = bonet ? log(i3) : 0.0;
  const double cs = bonet ? (lambda / 2 * lg - mu) / i3 : lambda / 2 - lambda / (2 * i3) - mu / i3;
#pragma unroll
  for (int i = 0; i < 9; ++i) h.S[i] = cs * h.di3[i];
  h.S[0] += mu; h.S[4] += mu; h.S[8] += mu;
  if (detF <= 0) {  // :655-656
#pragma unroll
    for (int i = 0; i < 9; ++i) h.S[i] += 1e200 * C[i];
  }
  if (bonet) { h.c1 = (lambda * lg - 2 * mu) / i3; h.c2 = (lambda + 2 * mu - lambda * lg) / (i3 * i3); }
  else { h.c1 = lambda - (lambda + 2 * mu) / i3; h.c2 = (lambda + 2 * mu) / (i3 * i3); }
  h.hd = i3 / 2;
}

// the nine entries D(a,n,b,l), a,b = 0..2, of the slice (l, n), multiplied by coeff, written at their place in D[81]
__device__ __forceinline__ void hyper_slice(int law, const HyperPrep &h, int l, int n, double lambda, double mu,
                                            double coeff, double *D) {
  const double *F = h.F, *S = h.S, *Gu = h.Gu;
  if (law == GFGPU_SVK) {
    // D(a,n,b,l) = delta_ab S(l,n) + sum_p F(a,p) dS(p,n,b,l),
    // dS(p,n,b,l) = lambda(d_pn d_bl + d_pn Gu_bl) + mu(d_pb d_nl + d_pl d_nb + d_pl Gu_bn + d_ln Gu_bp)
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        double v = (a == b) ? S[l + 3 * n] : 0.0;
        const double fbl = Gu[b + 3 * l] + (b == l ? 1.0 : 0.0);  // delta_bl + Gu_bl
        v += lambda * F[a + 3 * n] * fbl;                         // p = n
        v += mu * ((n == l ? F[a + 3 * b] : 0.0) + (n == b ? F[a + 3 * l] : 0.0) + F[a + 3 * l] * Gu[b + 3 * n]);
        if (l == n) {
          double s = 0;
#pragma unroll
          for (int p = 0; p < 3; ++p) s += F[a + 3 * p] * Gu[b + 3 * p];
          v += mu * s;
        }
        D[a + 3 * (n + 3 * (b + 3 * l))] = coeff * v;
      }
    return;
  }
#define CI(i, j) h.Ci[(i) + 3 * (j)]
  // dS(p,n,b,l) = sum_m A(p,n,m,l) F(b,m);  A = c1*ddi3 + c2*di3 (x) di3  (:142-152, :674-692)
  // D(a,n,b,l) = delta_ab S(l,n) + sum_p F(a,p) sum_m A(p,n,m,l) F(b,m);  T(p,m) = A(p,n,m,l)
  double T[9];
#pragma unroll
  for (int p = 0; p < 3; ++p)
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      const double dd = h.hd * (CI(n, p) * CI(l, m) - CI(n, m) * CI(l, p) + CI(p, n) * CI(l, m) - CI(p, m) * CI(l, n));
      T[p + 3 * m] = h.c1 * dd + h.c2 * h.di3[p + 3 * n] * h.di3[m + 3 * l];
    }
#undef CI
  double FT[9];  // FT(a,m) = sum_p F(a,p) T(p,m)
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int m = 0; m < 3; ++m) FT[a + 3 * m] = F[a] * T[3 * m] + F[a + 3] * T[1 + 3 * m] + F[a + 6] * T[2 + 3 * m];
#pragma unroll
  for (int b = 0; b < 3; ++b)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      double v = FT[a] * F[b] + FT[a + 3] * F[b + 3] + FT[a + 6] * F[b + 6];
      if (a == b) v += S[l + 3 * n];
      D[a + 3 * (n + 3 * (b + 3 * l))] = coeff * v;
    }
}

// P = F S (flux), multiplied by coeff
__device__ __forceinline__ void hyper_flux(const HyperPrep &h, double coeff, double *P) {
#pragma unroll
  for (int n = 0; n < 3; ++n)
#pragma unroll
    for (int a = 0; a < 3; ++a)
      P[a + 3 * n] = coeff * (h.F[a] * h.S[3 * n] + h.F[a + 3] * h.S[1 + 3 * n] + h.F[a + 6] * h.S[2 + 3 * n]);
}

// Writes P (flux, 9) and D (81) to (shared) memory, both multiplied by `coeff`.
// law: GFGPU_SVK / NEOHOOKEAN_CIARLET / NEOHOOKEAN_BONET.
__device__ inline void hyper_point(int law, const double *Gu, double lambda, double mu, double coeff, double *P,
                                   double *D) {
  HyperPrep h;
  hyper_prep(law, Gu, lambda, mu, h);
  // (l, n) fully unrolled: every index into Ci / di3 / S is then static and the 3x3 temporaries stay in registers
#pragma unroll
  for (int l = 0; l < 3; ++l)
#pragma unroll
    for (int n = 0; n < 3; ++n) hyper_slice(law, h, l, n, lambda, mu, coeff, D);
  hyper_flux(h, coeff, P);
}

// The laws of add_finite_strain_elasticity_brick written on the invariants of C = F^T F (getfem_nonlinear_elasticity.cc:
// 2276-2290): compressible Mooney-Rivlin (:503-607, the law of the reference's tests/nonlinear_elastostatic.cc; invariants
// j1, j2, i3 and their derivatives from compute_invariants, :45-262), Ciarlet-Geymonat (:817-888), generalized Blatz-Ko
// (:706-815).  Their second derivative has one shape,
//   A(i,j,k,l) = alpha ddi3(i,j,k,l) + gamma ddi2(i,j,k,l) + sum_pq beta_pq dv_p(i,j) dv_q(k,l),   dv = (Id, di2, di3),
// with scalar coefficients per law; PK2 and the tangent dS(i,j,k,l) = sum_m A(i,j,m,l) F(k,m) then go through
// AHL_wrapper_sigma (:1781-1827) exactly like the Neo-Hookean laws above.  par: Mooney-Rivlin (C10, C01, D1),
// Ciarlet-Geymonat (lambda, mu, a), Blatz-Ko (a, b, c, d, n).  Writes P = coeff F S (9) and D (81) like hyper_point.
__device__ inline void inv_law_point(int law, const double *Gu, const double *par, double coeff, double *P, double *D) {
  double F[9], C[9], Ci[9], S[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double s = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) s += Gu[k + 3 * i] * Gu[k + 3 * j];
      C[i + 3 * j] = s + Gu[i + 3 * j] + Gu[j + 3 * i] + (i == j ? 1.0 : 0.0);  // 2 E + Id
      F[i + 3 * j] = Gu[i + 3 * j] + (i == j ? 1.0 : 0.0);
    }
  const double detF = det3cm(F), i3 = det3cm(C);
  {
#define C_(i, j) C[(i) + 3 * (j)]
    Ci[0] = (C_(1, 1) * C_(2, 2) - C_(1, 2) * C_(2, 1)) / i3;
    Ci[1] = -(C_(1, 0) * C_(2, 2) - C_(1, 2) * C_(2, 0)) / i3;
    Ci[2] = (C_(1, 0) * C_(2, 1) - C_(1, 1) * C_(2, 0)) / i3;
    Ci[3] = -(C_(0, 1) * C_(2, 2) - C_(0, 2) * C_(2, 1)) / i3;
    Ci[4] = (C_(0, 0) * C_(2, 2) - C_(0, 2) * C_(2, 0)) / i3;
    Ci[5] = -(C_(0, 0) * C_(2, 1) - C_(0, 1) * C_(2, 0)) / i3;
    Ci[6] = (C_(0, 1) * C_(1, 2) - C_(0, 2) * C_(1, 1)) / i3;
    Ci[7] = -(C_(0, 0) * C_(1, 2) - C_(0, 2) * C_(1, 0)) / i3;
    Ci[8] = (C_(0, 0) * C_(1, 1) - C_(0, 1) * C_(1, 0)) / i3;
#undef C_
  }
  const double i1 = C[0] + C[4] + C[8];
  double ff = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) ff += C[i + 3 * j] * C[j + 3 * i];
  const double i2 = (i1 * i1 - ff) / 2;
  // dv_1 = Id, dv_2 = di2 = i1 Id - C (compute_di2, :95-102), dv_3 = di3 = i3 C^-1 (compute_di3, :132-140)
  double alpha = 0, gamma = 0, b11 = 0, b12 = 0, b13 = 0, b22 = 0, b23 = 0, b33 = 0;
  double s1 = 0, s2 = 0, s3 = 0;  // S = 2 (s1 Id + s2 di2 + s3 di3)
  if (law == GFGPU_MOONEY_RIVLIN) {
    const double c10 = par[0], c01 = par[1], d1 = par[2];
    const double p13 = pow(fabs(i3), -1.0 / 3.0), p23 = pow(fabs(i3), -2.0 / 3.0);
    const double k1 = 1.0 / (3 * i3), k2 = 4 * k1 * k1 * i1;      // compute_ddj1 (:176-196)
    const double m1 = 2.0 / (3 * i3), m2 = 5 * m1 * m1 * i2 / 2;  // compute_ddj2 (:219-240)
    const double dw3 = d1 - d1 / sqrt(fabs(i3)), a22 = d1 / (2 * pow(fabs(i3), 1.5));
    s1 = c10 * p13;                                              // dj1 = (di1 - i1/(3 i3) di3) p13 (:169-174)
    s2 = c01 * p23;                                              // dj2 = (di2 - 2 i2/(3 i3) di3) p23 (:212-217)
    s3 = -c10 * p13 * i1 / (3 * i3) - c01 * p23 * 2 * i2 / (3 * i3) + dw3;
    alpha = -4 * c10 * p13 * i1 * k1 - 4 * c01 * p23 * i2 * m1 + 4 * dw3;
    gamma = 4 * c01 * p23;
    b13 = -4 * c10 * p13 * k1;
    b23 = -4 * c01 * p23 * m1;
    b33 = 4 * c10 * p13 * k2 + 4 * c01 * p23 * m2 + 4 * a22;
  } else if (law == GFGPU_CIARLET_GEYMONAT) {
    // W = a i1 + b i2 + c i3 - d/2 log i3 + e (:817-888): S = -2 b C + 2 (a + b tr C) Id + C^-1 (2 c i3 - d)
    const double a = par[2], b = par[1] / 2 - par[2], c = par[0] / 4 - par[1] / 2 + par[2], d = par[0] / 2 + par[1];
    s1 = a;  // 2 a Id + 2 b (i1 Id - C) = 2 (a + b tr C) Id - 2 b C
    s2 = b;
    s3 = detF <= 0 ? 0.0 : (2 * c * i3 - d) / (2 * i3);  // the penalty REPLACES this part when det F <= 0 (:848-851)
    alpha = -2 * (d - 2 * i3 * c) / i3;
    gamma = 4 * b;
    b33 = 2 * d / (i3 * i3);
  } else {  // generalized Blatz-Ko (:706-815): W = (a i1 + b sqrt|i3| + c i2 / i3 + d)^n
    const double a = par[0], b = par[1], c = par[2], d = par[3], n = par[4];
    const double z = a * i1 + b * sqrt(fabs(i3)) + c * i2 / i3 + d, nz = n * pow(z, n - 1.);
    const double y = b / (2. * sqrt(fabs(i3))) - c * i2 / (i3 * i3);
    s1 = nz * a; s2 = nz * c / i3; s3 = nz * y;
    const double nnz = n * (n - 1.) * pow(z, n - 2.);
    alpha = 4 * s3;
    gamma = 4 * s2;
    b11 = 4 * nnz * a * a;
    b12 = 4 * nnz * a * c / i3;
    b13 = 4 * nnz * a * y;
    b22 = 4 * nnz * c * c / (i3 * i3);
    b23 = 4 * (nnz * y * c / i3 - nz * c / (i3 * i3));
    b33 = 4 * (nnz * y * y + nz * (2. * c * i2 / (i3 * i3 * i3) - b / (4. * pow(i3, 1.5))));
  }
  double di2[9], di3[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    di2[i] = ((i % 4) == 0 ? i1 : 0.0) - C[i];
    di3[i] = Ci[i] * i3;
    S[i] = 2 * (((i % 4) == 0 ? s1 : 0.0) + s2 * di2[i] + s3 * di3[i]);
  }
  if (detF <= 0) {  // the reference's penalty on inverted elements
#pragma unroll
    for (int i = 0; i < 9; ++i) S[i] += 1e200 * C[i];
  }
  const double hd = i3 / 2;
#define CI(i, j) Ci[(i) + 3 * (j)]
#pragma unroll
  for (int l = 0; l < 3; ++l)
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      double T[9];  // T(p,m) = A(p,n,m,l)
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          const double dd3 = hd * (CI(n, p) * CI(l, m) - CI(n, m) * CI(l, p) + CI(p, n) * CI(l, m) - CI(p, m) * CI(l, n));
          const double dd2 = (p == n && m == l ? 1.0 : 0.0) - (n == m && p == l ? 0.5 : 0.0) - (p == m && n == l ? 0.5 : 0.0);
          const double u1 = p == n ? 1.0 : 0.0, u2 = di2[p + 3 * n], u3 = di3[p + 3 * n];
          const double v1 = m == l ? 1.0 : 0.0, v2 = di2[m + 3 * l], v3 = di3[m + 3 * l];
          T[p + 3 * m] = alpha * dd3 + gamma * dd2 + b11 * u1 * v1 + b12 * (u1 * v2 + u2 * v1) + b13 * (u1 * v3 + u3 * v1) +
                         b22 * u2 * v2 + b23 * (u2 * v3 + u3 * v2) + b33 * u3 * v3;
        }
      double FT[9];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int m = 0; m < 3; ++m) FT[a + 3 * m] = F[a] * T[3 * m] + F[a + 3] * T[1 + 3 * m] + F[a + 6] * T[2 + 3 * m];
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          double v = FT[a] * F[b] + FT[a + 3] * F[b + 3] + FT[a + 6] * F[b + 6];
          if (a == b) v += S[l + 3 * n];
          D[a + 3 * (n + 3 * (b + 3 * l))] = coeff * v;
        }
    }
#undef CI
#pragma unroll
  for (int n = 0; n < 3; ++n)
#pragma unroll
    for (int a = 0; a < 3; ++a) P[a + 3 * n] = coeff * (F[a] * S[3 * n] + F[a + 3] * S[1 + 3 * n] + F[a + 6] * S[2 + 3 * n]);
}

template <int DIM, int Q, int ND, int FK, bool AFFINE>
__global__ void __launch_bounds__(ElemCfg<DIM, Q, ND, FK, AFFINE>::THREADS)
elem_kernel(const ElemArgs a) {
  using C = ElemCfg<DIM, Q, ND, FK, AFFINE>;
  constexpr int N = C::N, S1 = C::S1, NB = C::NB, TPE = C::TPE, EPB = C::EPB, GEO = C::GEO;
  extern __shared__ double sm[];
  const int slot = threadIdx.x / TPE, lt = threadIdx.x % TPE;
  const int ng = a.ng, qc = a.qc, nq = a.face ? a.nqf : a.nq;
  double *sG = sm + (size_t)slot * C::slot_doubles(ng, qc);
  double *sU = sG + N * ng;
  double *sGeoA = sU + S1;
  double *sGeo = sGeoA + GEO;
  double *sZ = sGeo + qc * GEO;
  double *sGu = sZ + qc * ND * N;
  double *sD = sGu + qc * Q * N;
  double *sP = sD + qc * C::DSZ;
  double *sLM = sP + qc * Q * N;  // Lame coefficients at the Gauss points (constants, or fem-data fields)
  double *sRed = sLM + 2 * qc;
  const bool do_t = a.stage != nullptr || a.emask != nullptr;
  const bool do_r = a.rstage != nullptr;
  const bool need_gu = do_r || FK == FK_HYPER;
  const int64_t ne = a.e1 - a.e0;

  for (int64_t base = (int64_t)blockIdx.x * EPB; base < ne; base += (int64_t)gridDim.x * EPB) {
    const int64_t el = base + slot;
    const bool active = el < ne;
    const int64_t e = a.e0 + el;
    // tables of this item: the volume points, or the points of its face
    const int face = (a.face && active) ? a.face[e] : -1;
    const double *nref = face >= 0 ? a.fnormal + face * 3 : nullptr;
    const double *tw = face >= 0 ? a.fw + (size_t)face * nq : a.w;
    const double *tgt = face >= 0 ? a.fgt_grad + (size_t)face * nq * ng * N : a.gt_grad;
    const double *tphi = face >= 0 ? a.fphi + (size_t)face * nq * ND : a.phi;
    const double *tgphi = face >= 0 ? a.fgphi + (size_t)face * nq * ND * N : a.gphi;
    // ---- A: gather
    if (active) {
      for (int idx = lt; idx < N * ng; idx += TPE) {
        int i = idx / N, d = idx % N;
        int32_t p = a.conn[e * ng + i];
        sG[d + N * i] = (d == 0 ? a.x : d == 1 ? a.y : a.z)[p];
      }
      for (int idx = lt; idx < S1; idx += TPE) {
        int i = idx / Q, q = idx % Q;
        sU[idx] = a.U ? a.U[a.edof[e * ND + i] + q] : 0.0;
      }
    }
    __syncthreads();
    if (AFFINE && active && lt == 0) geometry<N>(sG, tgt, ng, sGeoA, nref);

    double acc[C::BPT][C::ACC];
    double racc[C::RPT];
#pragma unroll
    for (int k = 0; k < C::BPT; ++k)
#pragma unroll
      for (int m = 0; m < C::ACC; ++m) acc[k][m] = 0.0;
#pragma unroll
    for (int m = 0; m < C::RPT; ++m) racc[m] = 0.0;

    for (int q0 = 0; q0 < nq; q0 += qc) {
      const int qn = min(qc, nq - q0);
      // ---- B1: geometry per Gauss point
      if (!AFFINE && active)
        for (int q = lt; q < qn; q += TPE) geometry<N>(sG, tgt + (size_t)(q0 + q) * ng * N, ng, sGeo + q * GEO, nref);
      __syncthreads();
      // ---- B2: Z
      if (active)
        for (int idx = lt; idx < qn * ND; idx += TPE) {
          int q = idx / ND, i = idx % ND;
          const double *B = AFFINE ? sGeoA : sGeo + q * GEO;
          const double *g = tgphi + ((size_t)(q0 + q) * ND + i) * N;
          double gl[N];
#pragma unroll
          for (int p = 0; p < N; ++p) gl[p] = g[p];
#pragma unroll
          for (int n = 0; n < N; ++n) {
            double s = 0;
#pragma unroll
            for (int p = 0; p < N; ++p) s += gl[p] * B[n + N * p];
            sZ[(q * ND + i) * N + n] = s;
          }
        }
      __syncthreads();
      // ---- B3: Grad_u (or u_h for the mass family)
      if (need_gu && active) {
        if (FK == FK_MASS) {
          for (int idx = lt; idx < qn * Q; idx += TPE) {
            int q = idx / Q, c = idx % Q;
            const double *ph = tphi + (size_t)(q0 + q) * ND;
            double s = 0;
            for (int i = 0; i < ND; ++i) s += sU[i * Q + c] * ph[i];
            sGu[q * Q * N + c] = s;
          }
        } else {
          for (int idx = lt; idx < qn * Q * N; idx += TPE) {
            int q = idx / (Q * N), r = idx % (Q * N);
            int c = r % Q, n = r / Q;
            double s = 0;
            for (int i = 0; i < ND; ++i) s += sU[i * Q + c] * sZ[(q * ND + i) * N + n];
            sGu[q * Q * N + c + Q * n] = s;
          }
        }
      }
      __syncthreads();
      // ---- B4: material point
      if (active)
        for (int q = lt; q < qn; q += TPE) {
          const double wq = tw[q0 + q];
          const double J = AFFINE ? sGeoA[N * N] : sGeo[q * GEO + N * N];
          const double coeff = (wq == 0.0) ? 0.0 : a.alpha * J * wq;  // zero-weight points are skipped (C&E.cc:8852)
          const double *Gu = sGu + q * Q * N;
          double *P = sP + q * Q * N;
          // parameters at the point: constants, or fem-data fields (ga_instruction_val on the data fem, C&E.cc:636-690)
          double par0 = a.par[0], par1 = a.par[1], fsrc[Q];
#pragma unroll
          for (int c = 0; c < Q; ++c) fsrc[c] = a.par[c];
          if (a.nfields) {
            const double *dp = (face >= 0 ? a.dfphi + ((size_t)face * nq + q0 + q) * a.nd_d : a.dphi + (size_t)(q0 + q) * a.nd_d);
            const int32_t *dd = a.dedof + e * a.nd_d;
            if (a.family == GFGPU_SOURCE) {
#pragma unroll
              for (int c = 0; c < Q; ++c) fsrc[c] = 0.0;
              for (int i = 0; i < a.nd_d; ++i) {
                const double ph = dp[i];
                const double *v = a.dvals0 + dd[i];
#pragma unroll
                for (int c = 0; c < Q; ++c) fsrc[c] += v[c] * ph;
              }
            } else {
              double v0 = 0.0, v1 = 0.0;
              for (int i = 0; i < a.nd_d; ++i) {
                v0 += a.dvals0[dd[i]] * dp[i];
                if (a.nfields > 1) v1 += a.dvals1[dd[i]] * dp[i];
              }
              par0 = v0;
              if (a.nfields > 1) par1 = v1;
            }
          }
          const double lambda = par0, mu = par1;
          sLM[2 * q] = lambda;
          sLM[2 * q + 1] = mu;
          if (FK == FK_LAPLACE) {
            sD[q] = coeff * par0;
            if (need_gu)
              for (int r = 0; r < Q * N; ++r) P[r] = coeff * par0 * Gu[r];
          } else if (FK == FK_MASS) {
            if (a.family == GFGPU_SOURCE) {  // "F.Test_u": r_e(i b) = sum_q J w_q F_b phi_i (C&E.cc:437-461, 4669-4735)
              sD[q] = 0.0;
              if (need_gu)
                for (int c = 0; c < Q; ++c) P[c] = coeff * fsrc[c];
            } else if (a.family == GFGPU_NORMAL_SOURCE) {  // "(A*Normal).Test_u", A(b,n) = par[b + Q n] (getfem_models.cc:4290-4299)
              sD[q] = 0.0;
              const double *nrm = (AFFINE ? sGeoA : sGeo + q * GEO) + N * N + 1;
              if (need_gu)
                for (int c = 0; c < Q; ++c) {
                  double an = 0;
                  for (int n = 0; n < N; ++n) an += a.par[c + Q * n] * nrm[n];
                  P[c] = coeff * an;
                }
            } else {
              sD[q] = coeff * par0;
              if (need_gu)
                for (int c = 0; c < Q; ++c) P[c] = coeff * par0 * Gu[c];
            }
          } else if (FK == FK_ELAST) {
            sD[q] = coeff;
            if (need_gu) {
              double tr = 0;
              for (int c = 0; c < N; ++c) tr += Gu[c + Q * c];
              for (int n = 0; n < N; ++n)
                for (int c = 0; c < Q; ++c)
                  P[c + Q * n] = coeff * (mu * (Gu[c + Q * n] + Gu[n + Q * c]) + (c == n ? lambda * tr : 0.0));
            }
          } else {
            if (wq == 0.0) {
              for (int r = 0; r < C::DSZ; ++r) sD[q * C::DSZ + r] = 0.0;
              for (int r = 0; r < Q * N; ++r) P[r] = 0.0;
            } else {
              double gul[9];
              for (int r = 0; r < 9; ++r) gul[r] = Gu[r % (Q * N)];
              if (a.family >= GFGPU_MOONEY_RIVLIN) inv_law_point(a.family, gul, a.par, coeff, P, sD + q * C::DSZ);
              else hyper_point(a.family, gul, lambda, mu, coeff, P, sD + q * C::DSZ);
            }
          }
        }
      __syncthreads();
      // ---- C: accumulate the blocks owned by this thread
      if (active && do_t) {
#pragma unroll
        for (int k = 0; k < C::BPT; ++k) {
          const int p = lt + k * TPE;
          if (p < NB) {
            const int j = p / ND, i = p % ND;
            for (int q = 0; q < qn; ++q) {
              const double *Zi = sZ + (q * ND + i) * N, *Zj = sZ + (q * ND + j) * N;
              if (FK == FK_LAPLACE) {
                double s = 0;
#pragma unroll
                for (int n = 0; n < N; ++n) s += Zi[n] * Zj[n];
                acc[k][0] += sD[q] * s;
              } else if (FK == FK_MASS) {
                const double *ph = tphi + (size_t)(q0 + q) * ND;
                acc[k][0] += sD[q] * ph[i] * ph[j];
              } else if (FK == FK_ELAST) {
                double zi[N], zj[N], s = 0;
#pragma unroll
                for (int n = 0; n < N; ++n) { zi[n] = Zi[n]; zj[n] = Zj[n]; s += zi[n] * zj[n]; }
                const double c = sD[q], lambda = sLM[2 * q], mu = sLM[2 * q + 1];
#pragma unroll
                for (int bb = 0; bb < Q; ++bb)
#pragma unroll
                  for (int aa = 0; aa < Q; ++aa)
                    acc[k][bb * Q + aa] += c * (lambda * zi[aa] * zj[bb] + mu * zi[bb] * zj[aa] + (aa == bb ? mu * s : 0.0));
              } else {
                const double *D = sD + q * C::DSZ;
                double zi[N], zj[N];
#pragma unroll
                for (int n = 0; n < N; ++n) { zi[n] = Zi[n]; zj[n] = Zj[n]; }
#pragma unroll
                for (int bb = 0; bb < Q; ++bb)
#pragma unroll
                  for (int aa = 0; aa < Q; ++aa) {
                    double s = 0;
#pragma unroll
                    for (int l = 0; l < N; ++l)
#pragma unroll
                      for (int n = 0; n < N; ++n) s += zi[n] * D[aa + Q * (n + N * (bb + Q * l))] * zj[l];
                    acc[k][bb * Q + aa] += s;
                  }
              }
            }
          }
        }
      }
      if (active && do_r) {
#pragma unroll
        for (int m = 0; m < C::RPT; ++m) {
          const int idx = lt + m * TPE;
          if (idx < S1) {
            const int i = idx / Q, c = idx % Q;
            double s = 0;
            if (FK == FK_MASS) {
              for (int q = 0; q < qn; ++q) s += sP[q * Q * N + c] * tphi[(size_t)(q0 + q) * ND + i];
            } else {
              for (int q = 0; q < qn; ++q) {
                const double *Zi = sZ + (q * ND + i) * N;
#pragma unroll
                for (int n = 0; n < N; ++n) s += sP[q * Q * N + c + Q * n] * Zi[n];
              }
            }
            racc[m] += s;
          }
        }
      }
      __syncthreads();
    }

    // ---- D: drop rule and output
    if (do_t) {
      double vmax = 0.0;
#pragma unroll
      for (int k = 0; k < C::BPT; ++k)
#pragma unroll
        for (int m = 0; m < C::ACC; ++m) vmax = fmax(vmax, fabs(acc[k][m]));
#pragma unroll
      for (int off = (TPE < 32 ? TPE : 32) / 2; off > 0; off >>= 1)
        vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, off));
      if (TPE > 32) {
        if ((lt & 31) == 0) sRed[lt >> 5] = vmax;
        __syncthreads();
        vmax = 0.0;
#pragma unroll
        for (int wv = 0; wv < TPE / 32; ++wv) vmax = fmax(vmax, sRed[wv]);
      }
      // NaN-safe: a NaN/inf element keeps everything that compares greater
      const double thr = vmax * 1e-14;
      if (active) {
#pragma unroll
        for (int k = 0; k < C::BPT; ++k) {
          const int p = lt + k * TPE;
          if (p < NB) {
            const int j = p / ND, i = p % ND;
            unsigned mask = 0;
#pragma unroll
            for (int bb = 0; bb < Q; ++bb)
#pragma unroll
              for (int aa = 0; aa < Q; ++aa) {
                double v = C::SCALAR ? (aa == bb ? acc[k][0] : 0.0) : acc[k][(bb * Q + aa) % C::ACC];
                const bool keep = (vmax != 0.0) && (fabs(v) > thr);
                if (keep) mask |= 1u << (bb * Q + aa);
                if (a.stage) a.stage[(size_t)el * S1 * S1 + (size_t)(j * Q + bb) * S1 + i * Q + aa] = keep ? v : 0.0;
              }
            if (a.emask) a.emask[(size_t)el * NB + p] = (uint16_t)mask;
          }
        }
      }
    }
    if (do_r && active) {
#pragma unroll
      for (int m = 0; m < C::RPT; ++m) {
        const int idx = lt + m * TPE;
        if (idx < S1) a.rstage[(size_t)el * S1 + idx] = racc[m];
      }
    }
    __syncthreads();
  }
}

template <int DIM, int Q, int ND, int FK, bool AFFINE>
void launch_elem_t(gfgpu_ctx *ctx, ElemArgs a) {
  using C = ElemCfg<DIM, Q, ND, FK, AFFINE>;
  auto kern = elem_kernel<DIM, Q, ND, FK, AFFINE>;
  // Gauss-point chunk: as many points as fit ~100 KB of shared memory per CTA
  const size_t budget = 100 * 1024;
  const size_t fixed = (size_t)C::EPB * (C::N * a.ng + C::S1 + C::GEO + 8) * 8;
  GF_REQUIRE(fixed + (size_t)C::EPB * C::per_q() * 8 <= 200 * 1024, "element too large for shared memory");
  int qc = (int)((budget > fixed ? budget - fixed : 0) / ((size_t)C::EPB * C::per_q() * 8));
  if (qc < 1) qc = 1;
  const int nq_used = a.face ? a.nqf : a.nq;
  if (qc > nq_used) qc = nq_used;
  a.qc = qc;
  const size_t smem = (size_t)C::EPB * C::slot_doubles(a.ng, qc) * 8;
  GF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  GF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::THREADS, smem));
  if (occ < 1) occ = 1;
  const int64_t ne = a.e1 - a.e0;
  int64_t want = (ne + C::EPB - 1) / C::EPB;
  int64_t cap = (int64_t)ctx->sm_count * occ;
  int grid = (int)(want < cap ? want : cap);
  if (grid < 1) return;
  kern<<<grid, C::THREADS, smem, ctx->stream>>>(a);
  GF_LAUNCH_CHECK();
}

}  // namespace gf

#define GF_ELEM_CASE(DIM, QQ, NDD, FK, AFF)                                 \
  if (dim == DIM && Q == QQ && nd == NDD && fk == FK && affine == AFF) {    \
    gf::launch_elem_t<DIM, QQ, NDD, FK, AFF>(ctx, a);                       \
    return true;                                                            \
  }
