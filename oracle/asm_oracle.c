/* oracle/asm_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded CPU restatement of GetFEM's generic weak-form
 * assembly for the expression families this repository accelerates.  It is the
 * CHECKER for the CUDA path (tests/, __graft_entry__.smoke(), bench.py's
 * cpu_baseline leg); nothing in getfem_b200/ may link, import or call it.
 *
 * Parity pinning: this file is validated against the UNMODIFIED reference
 * (oracle/_ref/libgetfem.so driven by oracle/ref_driver.cc) and against the
 * golden fixtures in tests/golden/ generated from it (tests/test_oracle.py):
 * CSC pattern identical, values to 1e-13 relative Frobenius.
 *
 * What it follows (paths relative to /root/reference/src):
 *   element loop, Gauss loop, weight = J*w_q ....... getfem_generic_assembly_compile_and_exec.cc:8789-8866
 *   K = G*pc, J = |det K|, B = K^{-T} .............. bgeot_geometric_trans.cc:270-288, 321-355, 374-413
 *   grad(phi) = grad_ref(phi) * B^T ................ getfem_fem.cc:85-90, 160-168
 *   local coefficients coeff[node*Q+q] ............. getfem_mesh_fem.h:662-689
 *   Grad_u at the point ............................ compile_and_exec.cc:692-747
 *   integrands ................. getfem_models.cc:6112-6113,
 *                                                    getfem_nonlinear_elasticity.cc:612-702, 1781-1827, 1945-1994
 *   elem += coeff*t; finalize with ninf, threshold . compile_and_exec.cc:5359-5481
 *   add_elem_matrix: sorted column merge, drop
 *        |v| <= 1e-14*ninf ......................... compile_and_exec.cc:4853-4936
 *   vector assembly V[dof] += elem ................. compile_and_exec.cc:4669-4735
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { GFO_LAPLACE = 0, GFO_ELAST = 1, GFO_SVK = 2, GFO_NH_CIARLET = 3, GFO_NH_BONET = 4, GFO_MASS = 5, GFO_SOURCE = 6,
       GFO_NORMAL_SOURCE = 7, GFO_MOONEY_RIVLIN = 8 /* Compressible_Mooney_Rivlin_PK2, params (C10, C01, D1) */,
       GFO_CIARLET_GEYMONAT = 9 /* Ciarlet_Geymonat_PK2, params (lambda, mu, a) */,
       GFO_BLATZ_KO = 10 /* Generalized_Blatz_Ko_PK2, params (a, b, c, d, n) */ };

typedef struct { int64_t c; double e; } entry_t; /* gmm::elt_rsvector_ (gmm_vector.h:913-932) */
typedef struct { entry_t *v; int64_t n, cap; } col_t;

typedef struct gfo_result {
  int64_t ndof, nnz;
  col_t *cols;
  double *R;
} gfo_result;

static void col_insert(col_t *col, int64_t pos, entry_t ev) {
  if (col->n == col->cap) {
    col->cap = col->cap ? 2 * col->cap : 16;
    col->v = (entry_t *)realloc(col->v, (size_t)col->cap * sizeof(entry_t));
  }
  memmove(col->v + pos + 1, col->v + pos, (size_t)(col->n - pos) * sizeof(entry_t));
  col->v[pos] = ev;
  col->n++;
}

/* add_elem_matrix (compile_and_exec.cc:4853-4936): rows visited in ascending dof order,
   binary search from the last position, add in place or shifting insert. */
static void add_elem_matrix(col_t *K, int s1, int s2, const int64_t *dofs1, const int64_t *dofs2,
                            const double *elem, double threshold, int *sort) {
  for (int i = 0; i < s1; ++i) { /* insertion sort of the row dofs */
    int j = i;
    while (j > 0 && dofs1[i] < dofs1[sort[j - 1]]) { sort[j] = sort[j - 1]; j--; }
    sort[j] = i;
  }
  for (int c = 0; c < s2; ++c) {
    col_t *col = &K[dofs2[c]];
    const double *it = elem + (size_t)c * s1;
    int64_t ind = 0;
    for (int kk = 0; kk < s1; ++kk) {
      int k = sort[kk];
      entry_t ev; ev.e = it[k];
      if (fabs(ev.e) > threshold) {
        ev.c = dofs1[k];
        int64_t count = col->n - ind;
        while (count > 0) {
          int64_t step = count / 2, l = ind + step;
          if (col->v[l].c < ev.c) { ind = l + 1; count -= step + 1; } else count = step;
        }
        if (ind != col->n && col->v[ind].c == ev.c) col->v[ind].e += ev.e;
        else col_insert(col, ind, ev);
        ++ind;
      }
    }
  }
}

static double det3(const double *K, int N) { /* K column-major N x N */
  if (N == 1) return K[0];
  if (N == 2) return K[0] * K[3] - K[1] * K[2];
  return K[0] * (K[4] * K[8] - K[5] * K[7]) - K[3] * (K[1] * K[8] - K[2] * K[7]) +
         K[6] * (K[1] * K[5] - K[2] * K[4]);
}
/* Ainv = A^{-1}, column-major; returns det */
static double inv3(const double *A, double *Ai, int N) {
  double d = det3(A, N);
  if (N == 1) { Ai[0] = 1.0 / A[0]; return d; }
  if (N == 2) {
    Ai[0] = A[3] / d; Ai[1] = -A[1] / d; Ai[2] = -A[2] / d; Ai[3] = A[0] / d;
    return d;
  }
#define A_(i, j) A[(i) + 3 * (j)]
  Ai[0] = (A_(1, 1) * A_(2, 2) - A_(1, 2) * A_(2, 1)) / d;
  Ai[1] = -(A_(1, 0) * A_(2, 2) - A_(1, 2) * A_(2, 0)) / d;
  Ai[2] = (A_(1, 0) * A_(2, 1) - A_(1, 1) * A_(2, 0)) / d;
  Ai[3] = -(A_(0, 1) * A_(2, 2) - A_(0, 2) * A_(2, 1)) / d;
  Ai[4] = (A_(0, 0) * A_(2, 2) - A_(0, 2) * A_(2, 0)) / d;
  Ai[5] = -(A_(0, 0) * A_(2, 1) - A_(0, 1) * A_(2, 0)) / d;
  Ai[6] = (A_(0, 1) * A_(1, 2) - A_(0, 2) * A_(1, 1)) / d;
  Ai[7] = -(A_(0, 0) * A_(1, 2) - A_(0, 2) * A_(1, 0)) / d;
  Ai[8] = (A_(0, 0) * A_(1, 1) - A_(0, 1) * A_(1, 0)) / d;
#undef A_
  return d;
}

/* Hyperelastic PK2 stress S (N x N col-major) and dS(i,j,k,l) = dS_ij/d(Grad_u)_kl
   (first index fastest), as the registered GWFL operators compute them. */
static void hyper_law(int family, const double *Gu, const double *par, double *S, double *dS) {
  const int N = 3;
  double lambda = par[0], mu = par[1];
  double E[9], F[9];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      double s = 0;
      for (int k = 0; k < N; ++k) s += Gu[k + N * i] * Gu[k + N * j]; /* (Gu^T Gu)(i,j) */
      E[i + N * j] = 0.5 * (s + Gu[i + N * j] + Gu[j + N * i]);
      F[i + N * j] = Gu[i + N * j] + (i == j ? 1.0 : 0.0);
    }
  if (family == GFO_SVK) { /* getfem_nonlinear_elasticity.cc:1945-1994 */
    double trE = E[0] + E[4] + E[8];
    for (int j = 0; j < N; ++j)
      for (int i = 0; i < N; ++i) S[i + N * j] = 2 * mu * E[i + N * j] + (i == j ? lambda * trE : 0.0);
    double *it = dS;
    for (int l = 0; l < N; ++l)
      for (int k = 0; k < N; ++k)
        for (int j = 0; j < N; ++j)
          for (int i = 0; i < N; ++i, ++it) {
            double v = 0;
            if (i == j && k == l) v += lambda;
            if (i == j) v += lambda * Gu[k + N * l];
            if (i == k && j == l) v += mu;
            if (i == l && j == k) v += mu;
            if (i == l) v += mu * Gu[k + N * j];
            if (l == j) v += mu * Gu[k + N * i];
            *it = v;
          }
    return;
  }
  double detF = det3(F, N);
  double C[9], Ci[9];
  for (int i = 0; i < 9; ++i) C[i] = 2 * E[i];
  C[0] += 1; C[4] += 1; C[8] += 1;
  double i3 = inv3(C, Ci, N);
  double di3[9];
  for (int i = 0; i < 9; ++i) di3[i] = Ci[i] * i3; /* compute_di3 (cc:132-140) */
  double A4[81];
#define CI(i, j) Ci[(i) + 3 * (j)]
#define T4(T, i, j, k, l) T[(i) + 3 * (j) + 9 * (k) + 27 * (l)]
  if (family == GFO_MOONEY_RIVLIN) {
    /* Mooney_Rivlin_hyperelastic_law(compressible, !neohookean) (cc:503-607) on the invariants of C (compute_invariants,
       cc:45-262): W = C10 (j1 - 3) + C01 (j2 - 3) + D1 (sqrt|i3| - 1)^2, j1 = i1 i3^(-1/3), j2 = i2 i3^(-2/3) */
    const double c10 = par[0], c01 = par[1], d1 = par[2];
    double i1 = C[0] + C[4] + C[8], ff = 0;
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) ff += C[i + N * j] * C[j + N * i]; /* frobenius_product_trans(C, C) */
    const double i2 = (i1 * i1 - ff) / 2;
    double di1[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, di2[9], dd3[81], ddi2[81];
    for (int i = 0; i < 9; ++i) di2[i] = i1 * di1[i] - C[i]; /* compute_di2 (cc:95-102) */
    for (int i = 0; i < 81; ++i) ddi2[i] = 0;
    for (int i = 0; i < N; ++i)
      for (int k = 0; k < N; ++k) T4(ddi2, i, i, k, k) += 1.0; /* compute_ddi2 (cc:104-115) */
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) { T4(ddi2, i, j, j, i) -= 0.5; T4(ddi2, j, i, j, i) -= 0.5; }
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j)
        for (int k = 0; k < N; ++k)
          for (int l = 0; l < N; ++l) /* compute_ddi3 (cc:142-152) */
            T4(dd3, i, j, k, l) = i3 / 2 * (CI(j, i) * CI(l, k) - CI(j, k) * CI(l, i) + CI(i, j) * CI(l, k) - CI(i, k) * CI(l, j));
    const double p13 = pow(fabs(i3), -1.0 / 3.0), p23 = pow(fabs(i3), -2.0 / 3.0);
    const double k1 = 1.0 / (3 * i3), k2 = 4 * k1 * k1 * i1;          /* compute_ddj1 (cc:176-196) */
    const double m1 = 2.0 / (3 * i3), m2 = 5 * m1 * m1 * i2 / 2;      /* compute_ddj2 (cc:219-240) */
    const double dw3 = d1 - d1 / sqrt(fabs(i3)), a22 = d1 / (2 * pow(fabs(i3), 1.5));
    for (int i = 0; i < 9; ++i) {
      const double dj1 = (di1[i] - i1 / (3 * i3) * di3[i]) * p13;     /* compute_dj1 (cc:169-174) */
      const double dj2 = (di2[i] - 2 * i2 / (3 * i3) * di3[i]) * p23; /* compute_dj2 (cc:212-217) */
      S[i] = 2 * c10 * dj1 + 2 * c01 * dj2 + 2 * dw3 * di3[i];
    }
    if (detF <= 0) for (int i = 0; i < 9; ++i) S[i] += 1e200 * C[i];
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j)
        for (int k = 0; k < N; ++k)
          for (int l = 0; l < N; ++l) {
            const double d3ij = di3[i + 3 * j], d3kl = di3[k + 3 * l];
            const double ddj1 = (-i1 * k1 * T4(dd3, i, j, k, l) + d3ij * d3kl * k2 -
                                 (di1[i + 3 * j] * d3kl + di1[k + 3 * l] * d3ij) * k1) * p13;
            const double ddj2 = (T4(ddi2, i, j, k, l) - i2 * m1 * T4(dd3, i, j, k, l) + d3ij * d3kl * m2 -
                                 (di2[i + 3 * j] * d3kl + di2[k + 3 * l] * d3ij) * m1) * p23;
            T4(A4, i, j, k, l) = 4 * c10 * ddj1 + 4 * c01 * ddj2 + 4 * dw3 * T4(dd3, i, j, k, l) + 4 * a22 * d3ij * d3kl;
          }
  } else if (family == GFO_CIARLET_GEYMONAT) {
    /* Ciarlet_Geymonat_hyperelastic_law (cc:817-888): W = a i1(C) + b i2(C) + c i3(C) - d/2 log i3(C) + e */
    const double a = par[2], b = par[1] / 2 - par[2], c = par[0] / 4 - par[1] / 2 + par[2], d = par[0] / 2 + par[1];
    const double trC = C[0] + C[4] + C[8], b2 = 2 * b;
    for (int i = 0; i < 9; ++i) S[i] = -2 * b * C[i];
    for (int i = 0; i < N; ++i) S[i + N * i] += 2 * (a + b * trC);
    if (detF <= 0) for (int i = 0; i < 9; ++i) S[i] += 1e200 * C[i];
    else for (int i = 0; i < 9; ++i) S[i] += Ci[i] * (2 * c * i3 - d);
    for (int i = 0; i < 81; ++i) A4[i] = 0;
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) {
        T4(A4, i, i, j, j) += 2 * b2;
        T4(A4, i, j, i, j) -= b2;
        T4(A4, i, j, j, i) -= b2;
        for (int k = 0; k < N; ++k)
          for (int l = 0; l < N; ++l)
            T4(A4, i, j, k, l) += (CI(i, k) * CI(l, j) + CI(i, l) * CI(k, j)) * (d - 2 * i3 * c) + CI(i, j) * CI(k, l) * i3 * c * 4;
      }
  } else if (family == GFO_BLATZ_KO) {
    /* generalized_Blatz_Ko_hyperelastic_law (cc:706-815): W = (a i1 + b sqrt|i3| + c i2 / i3 + d)^n on the invariants of C */
    const double a = par[0], b = par[1], c = par[2], d = par[3], n = par[4];
    double i1 = C[0] + C[4] + C[8], ff = 0;
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) ff += C[i + N * j] * C[j + N * i];
    const double i2 = (i1 * i1 - ff) / 2;
    double di1[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, di2[9];
    for (int i = 0; i < 9; ++i) di2[i] = i1 * di1[i] - C[i];
    const double z = a * i1 + b * sqrt(fabs(i3)) + c * i2 / i3 + d, nz = n * pow(z, n - 1.);
    const double w1 = nz * a, w2 = nz * c / i3, y = b / (2. * sqrt(fabs(i3))) - c * i2 / (i3 * i3), w3 = nz * y;
    for (int i = 0; i < 9; ++i) S[i] = 2 * (w1 * di1[i] + w2 * di2[i] + w3 * di3[i]);
    if (detF <= 0) for (int i = 0; i < 9; ++i) S[i] += 1e200 * C[i];
    const double nnz = n * (n - 1.) * pow(z, n - 2.);
    double A[3][3];
    A[0][0] = nnz * a * a;
    A[1][0] = A[0][1] = nnz * a * c / i3;
    A[2][0] = A[0][2] = nnz * a * y;
    A[1][1] = nnz * c * c / (i3 * i3);
    A[2][1] = A[1][2] = nnz * y * c / i3 - nz * c / (i3 * i3);
    A[2][2] = nnz * y * y + nz * (2. * c * i2 / pow(i3, 3.) - b / (4. * pow(i3, 1.5)));
    const double *dv[3] = {di1, di2, di3};
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j)
        for (int k = 0; k < N; ++k)
          for (int l = 0; l < N; ++l) {
            double v = 0; /* 4 w2 ddi2 + 4 w3 ddi3 (ddi1 = 0), compute_ddi2 / compute_ddi3 (cc:104-115, 142-152) */
            double dd2 = (i == j && k == l ? 1.0 : 0.0) - (j == k && i == l ? 0.5 : 0.0) - (i == k && j == l ? 0.5 : 0.0);
            double dd3 = i3 / 2 * (CI(j, i) * CI(l, k) - CI(j, k) * CI(l, i) + CI(i, j) * CI(l, k) - CI(i, k) * CI(l, j));
            v = 4 * w2 * dd2 + 4 * w3 * dd3;
            for (int p = 0; p < 3; ++p)
              for (int q = 0; q < 3; ++q) v += 4. * A[p][q] * dv[p][i + 3 * j] * dv[q][k + 3 * l];
            T4(A4, i, j, k, l) = v;
          }
  } else {
  /* Neo_Hookean_hyperelastic_law (cc:612-702), through AHL_wrapper_sigma (cc:1781-1827) */
  int bonet = family == GFO_NH_BONET;
  double cs = bonet ? (lambda / 2 * log(i3) - mu) / i3 : lambda / 2 - lambda / (2 * i3) - mu / i3;
  for (int i = 0; i < 9; ++i) S[i] = cs * di3[i];
  S[0] += mu; S[4] += mu; S[8] += mu; /* mu * grad_i1 = mu * Id */
  if (detF <= 0) for (int i = 0; i < 9; ++i) S[i] += 1e200 * C[i];
  double c1, c2;
  if (bonet) { double lg = log(i3); c1 = (lambda * lg - 2 * mu) / i3; c2 = (lambda + 2 * mu - lambda * lg) / (i3 * i3); }
  else { c1 = lambda - (lambda + 2 * mu) / i3; c2 = (lambda + 2 * mu) / (i3 * i3); }
  double hd = i3 / 2;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j)
      for (int k = 0; k < N; ++k)
        for (int l = 0; l < N; ++l) { /* compute_ddi3 (cc:142-152) */
          double dd = hd * (CI(j, i) * CI(l, k) - CI(j, k) * CI(l, i) + CI(i, j) * CI(l, k) - CI(i, k) * CI(l, j));
          A4[i + 3 * j + 9 * k + 27 * l] = c1 * dd + c2 * di3[i + 3 * j] * di3[k + 3 * l];
        }
  }
#undef CI
#undef T4
  double *it = dS;
  for (int l = 0; l < N; ++l)
    for (int k = 0; k < N; ++k)
      for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i, ++it) {
          double v = 0;
          for (int m = 0; m < N; ++m) v += A4[i + 3 * j + 9 * m + 27 * l] * F[k + N * m];
          *it = v;
        }
}

/* Strain energy of the law at Grad_u ("<law>_potential(Grad_u,params)": AHL_wrapper_potential, cc:1841-1928, calls
   strain_energy(E, params, det(Id + Grad_u)): SVK cc:1996-2003, Neo-Hookean cc:612-632, Mooney-Rivlin cc:503-527,
   Ciarlet-Geymonat cc:817-836, Blatz-Ko cc:706-721) */
static double hyper_energy(int family, const double *Gu, const double *par) {
  const int N = 3;
  double E[9], F[9], C[9];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      double s = 0;
      for (int k = 0; k < N; ++k) s += Gu[k + N * i] * Gu[k + N * j];
      E[i + N * j] = 0.5 * (s + Gu[i + N * j] + Gu[j + N * i]);
      C[i + N * j] = 2 * E[i + N * j] + (i == j ? 1.0 : 0.0);
      F[i + N * j] = Gu[i + N * j] + (i == j ? 1.0 : 0.0);
    }
  if (family == GFO_SVK) {
    double tr = E[0] + E[4] + E[8], n2 = 0;
    for (int i = 0; i < 9; ++i) n2 += E[i] * E[i];
    return tr * tr * par[0] / 2 + n2 * par[1];
  }
  if (det3(F, N) <= 0) return 1e200;
  double i1 = C[0] + C[4] + C[8], i3 = det3(C, N), ff = 0, n2 = 0;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) { ff += C[i + N * j] * C[j + N * i]; n2 += C[i + N * j] * C[i + N * j]; }
  double i2 = (i1 * i1 - ff) / 2;
  if (family == GFO_NH_CIARLET || family == GFO_NH_BONET) {
    double lg = log(i3), W = par[1] / 2 * (i1 - 3.0 - lg);
    return W + (family == GFO_NH_BONET ? par[0] / 8 * lg * lg : par[0] / 4 * (i3 - 1.0 - lg));
  }
  if (family == GFO_MOONEY_RIVLIN) {
    double j1 = i1 * pow(fabs(i3), -1.0 / 3.0), j2 = i2 * pow(fabs(i3), -2.0 / 3.0), s = sqrt(fabs(i3)) - 1.0;
    return par[0] * (j1 - 3.0) + par[1] * (j2 - 3.0) + par[2] * s * s;
  }
  if (family == GFO_CIARLET_GEYMONAT) {
    double a = par[2], b = par[1] / 2 - par[2], c = par[0] / 4 - par[1] / 2 + par[2], d = par[0] / 2 + par[1];
    double e = -(3.0 * (a + b) + c);
    return a * i1 + b * (i1 * i1 - n2) / 2 + c * i3 - d * log(i3) / 2 + e;
  }
  return pow(par[0] * i1 + par[1] * sqrt(fabs(i3)) + par[2] * i2 / i3 + par[3], par[4]);
}

/* order 0 of the last gfo_assemble* call made with order_mask bit 2 (finite-strain families; the quadratic and linear
   families take 1/2 u.R and u.R from the residual, see oracle.py) */
static double g_last_potential = 0.0;
double gfo_last_potential(void) { return g_last_potential; }

/* pts: npts x dim (row-major), conn: ne x ng, elem_dof: ne x nd (dof of component 0),
   gt_grad: nq x ng x dim, phi: nq x nd, gphi: nq x nd x dim.  order_mask: bit0 residual, bit1 tangent.
   Region (mesh_region walked by mr_visitor, compile_and_exec.cc:8789): n_items items (item_cv[k], item_face[k]),
   face -1 = the whole convex; item_cv == NULL = all convexes.  With faces the tables hold ALL integration points
   (volume points first, then face by face: approx_integration::valid_method, getfem_integration.cc:353-368);
   the points of face f are [face_first[f], face_first[f] + face_nq[f]) and ref_normals[f] is pgt->normals()[f].
   On a face: Normal = B*n_ref, J *= |Normal|, Normal /= |Normal|, components below 1e-13 cleaned
   (compile_and_exec.cc:8836-8847). */
/* Fem-data coefficients (ga_workspace::add_fem_constant, generic_assembly.h:465; evaluated at the Gauss point by
   ga_instruction_val, compile_and_exec.cc:636-690): nfields > 0 replaces par[0 .. nfields) by the value at the point of a
   field on a data mesh_fem with nd_d local dofs (d_edof: ne x nd_d, dof of component 0; d_phi: its basis at every
   integration point, same point order as the other tables); the SOURCE family reads one field of Q components. */
gfo_result *gfo_assemble_fields(int dim, int64_t ne, int ng, const double *pts, const int32_t *conn, int nd, int Q,
                                const int64_t *elem_dof, int64_t ndof, int nq, const double *w, const double *gt_grad,
                                const double *phi, const double *gphi, int gt_linear, int family, const double *par_in,
                                const double *U, int order_mask, int64_t n_items, const int32_t *item_cv,
                                const int32_t *item_face, const int32_t *face_first, const int32_t *face_nq,
                                const double *ref_normals, int nfields, int nd_d, const int64_t *d_edof,
                                const double *d_phi, const double *d_vals0, const double *d_vals1) {
  double par[16];
  for (int k = 0; k < 16; ++k) par[k] = 0.0;
  { int np = family == GFO_SOURCE ? Q : family == GFO_NORMAL_SOURCE ? Q * dim
             : (family == GFO_LAPLACE || family == GFO_MASS) ? 1
             : (family == GFO_MOONEY_RIVLIN || family == GFO_CIARLET_GEYMONAT) ? 3 : family == GFO_BLATZ_KO ? 5 : 2;
    for (int k = 0; k < np; ++k) par[k] = par_in[k]; }
  const int N = dim, s1 = nd * Q;
  gfo_result *res = (gfo_result *)calloc(1, sizeof(gfo_result));
  res->ndof = ndof;
  res->cols = (col_t *)calloc((size_t)ndof, sizeof(col_t));
  res->R = (double *)calloc((size_t)ndof, sizeof(double));
  double *elem = (double *)malloc(sizeof(double) * s1 * s1), *t = (double *)malloc(sizeof(double) * s1 * s1);
  double *relem = (double *)malloc(sizeof(double) * s1), *Z = (double *)malloc(sizeof(double) * nd * N);
  double *G = (double *)malloc(sizeof(double) * N * ng), *ue = (double *)malloc(sizeof(double) * s1);
  int64_t *dofs = (int64_t *)malloc(sizeof(int64_t) * s1);
  int *sort = (int *)malloc(sizeof(int) * s1);
  double K[9], Ki[9], B[9], J = 0, D[81], P[9], Gu[9], S[9], dS[81], Nrm[3] = {0, 0, 0};
  const int nonlinear = family == GFO_SVK || family == GFO_NH_CIARLET || family == GFO_NH_BONET || family == GFO_MOONEY_RIVLIN ||
                        family == GFO_CIARLET_GEYMONAT || family == GFO_BLATZ_KO;
  if (!item_cv) n_items = ne;
  if (order_mask & 4) g_last_potential = 0.0;

  for (int64_t item = 0; item < n_items; ++item) {
    const int64_t cv = item_cv ? item_cv[item] : item;
    const int face = (item_cv && item_face) ? item_face[item] : -1;
    const int first_ind = face >= 0 ? face_first[face] : 0;
    const int nbpt = face >= 0 ? face_nq[face] : nq;
    for (int i = 0; i < ng; ++i) /* points_of_convex: G is N x ng column-major */
      for (int d = 0; d < N; ++d) G[d + N * i] = pts[(size_t)conn[cv * ng + i] * dim + d];
    for (int i = 0; i < nd; ++i)
      for (int q = 0; q < Q; ++q) {
        dofs[i * Q + q] = elem_dof[cv * nd + i] + q; /* populate_dofs_vector (cc:5008-5029) */
        ue[i * Q + q] = U ? U[dofs[i * Q + q]] : 0.0;
      }
    memset(elem, 0, sizeof(double) * s1 * s1);
    memset(relem, 0, sizeof(double) * s1);
    for (int ip = 0; ip < nbpt; ++ip) {
      const int ipt = first_ind + ip;
      if (ip == 0 || !gt_linear) {
        const double *pc = gt_grad + (size_t)ipt * ng * N; /* ng x P, row i = node */
        for (int r = 0; r < N; ++r)
          for (int c = 0; c < N; ++c) {
            double s = 0;
            for (int i = 0; i < ng; ++i) s += G[r + N * i] * pc[i * N + c];
            K[r + N * c] = s;
          }
        J = fabs(inv3(K, Ki, N));
        for (int r = 0; r < N; ++r)
          for (int c = 0; c < N; ++c) B[r + N * c] = Ki[c + N * r]; /* B = K^{-T} */
        if (face >= 0) { /* unit normal and surface Jacobian (cc:8836-8847) */
          double nup = 0;
          for (int r = 0; r < N; ++r) {
            double s = 0;
            for (int c = 0; c < N; ++c) s += B[r + N * c] * ref_normals[face * N + c];
            Nrm[r] = s;
            nup += s * s;
          }
          nup = sqrt(nup);
          J *= nup;
          for (int r = 0; r < N; ++r) {
            Nrm[r] /= nup;
            if (fabs(Nrm[r]) < 1e-13) Nrm[r] = 0.0; /* gmm::clean */
          }
        }
      }
      double coeff = J * w[ipt];
      if (w[ipt] == 0.0) continue; /* disabled points contribute coeff = 0 (cc:8852-8854) */
      if (nfields > 0) { /* coefficient fields at the point */
        const double *dp = d_phi + (size_t)ipt * nd_d;
        const int ncomp = family == GFO_SOURCE ? Q : 1;
        for (int k = 0; k < (family == GFO_SOURCE ? 1 : nfields); ++k) {
          const double *vals = k == 0 ? d_vals0 : d_vals1;
          for (int b = 0; b < ncomp; ++b) {
            double v = 0;
            for (int i = 0; i < nd_d; ++i) v += vals[d_edof[cv * nd_d + i] + b] * dp[i];
            par[k + b] = v;
          }
        }
      }
      const double *g = gphi + (size_t)ipt * nd * N;
      for (int i = 0; i < nd; ++i)
        for (int n = 0; n < N; ++n) {
          double s = 0;
          for (int p = 0; p < N; ++p) s += g[i * N + p] * B[n + N * p];
          Z[i * N + n] = s;
        }
      if (family == GFO_SOURCE) { /* "F.Test_u", F = par[0..Q): order 1 only (ga_instruction_vector_assembly_mf,
                                     cc:4669-4735, fed by val_base cc:437-461); no order-2 tree, the tangent is empty */
        const double *ph = phi + (size_t)ipt * nd;
        for (int i = 0; i < nd; ++i)
          for (int b = 0; b < Q; ++b) relem[i * Q + b] += coeff * par[b] * ph[i];
        continue;
      }
      if (family == GFO_NORMAL_SOURCE) { /* "(Reshape(A,qdim,meshdim)*Normal).Test_u" / "((g).Normal)*Test_u"
                                            (getfem_models.cc:4290-4299): A(b,n) = par[b + Q*n] */
        const double *ph = phi + (size_t)ipt * nd;
        for (int b = 0; b < Q; ++b) {
          double an = 0;
          for (int n = 0; n < N; ++n) an += par[b + Q * n] * Nrm[n];
          for (int i = 0; i < nd; ++i) relem[i * Q + b] += coeff * an * ph[i];
        }
        continue;
      }
      if (family == GFO_MASS) {
        const double *ph = phi + (size_t)ipt * nd;
        for (int j = 0; j < nd; ++j)
          for (int b = 0; b < Q; ++b)
            for (int i = 0; i < nd; ++i) elem[(i * Q + b) + s1 * (j * Q + b)] += coeff * par[0] * ph[i] * ph[j];
        for (int b = 0; b < Q; ++b) {
          double uh = 0;
          for (int i = 0; i < nd; ++i) uh += ue[i * Q + b] * ph[i];
          for (int i = 0; i < nd; ++i) relem[i * Q + b] += coeff * par[0] * uh * ph[i];
        }
        continue;
      }
      /* Grad_u(q, n) = sum_i u(i,q) Z(i,n)   (cc:692-747) */
      for (int q = 0; q < Q; ++q)
        for (int n = 0; n < N; ++n) {
          double s = 0;
          for (int i = 0; i < nd; ++i) s += ue[i * Q + q] * Z[i * N + n];
          Gu[q + Q * n] = s;
        }
      /* material tangent D(alpha,n,beta,l) and flux P(alpha,n) of the family */
      memset(D, 0, sizeof(D));
      memset(P, 0, sizeof(P));
#define D_(a, n, b, l) D[(a) + Q * ((n) + N * ((b) + Q * (l)))]
      if (family == GFO_LAPLACE) {
        for (int a = 0; a < Q; ++a)
          for (int n = 0; n < N; ++n) { D_(a, n, a, n) = par[0]; P[a + Q * n] = par[0] * Gu[a + Q * n]; }
      } else if (family == GFO_ELAST) {
        double lambda = par[0], mu = par[1], tr = 0;
        for (int a = 0; a < N; ++a) tr += Gu[a + Q * a];
        for (int a = 0; a < Q; ++a)
          for (int n = 0; n < N; ++n) {
            P[a + Q * n] = mu * (Gu[a + Q * n] + Gu[n + Q * a]) + (a == n ? lambda * tr : 0.0);
            for (int b = 0; b < Q; ++b)
              for (int l = 0; l < N; ++l)
                D_(a, n, b, l) = (a == n && b == l ? lambda : 0.0) + (a == l && b == n ? mu : 0.0) +
                                 (a == b && n == l ? mu : 0.0);
          }
      } else if (nonlinear) {
        if (order_mask & 4) g_last_potential += coeff * hyper_energy(family, Gu, par);
        hyper_law(family, Gu, par, S, dS);
        double F[9];
        for (int i = 0; i < 9; ++i) F[i] = Gu[i];
        F[0] += 1; F[4] += 1; F[8] += 1;
        for (int a = 0; a < 3; ++a)
          for (int n = 0; n < 3; ++n) {
            double s = 0;
            for (int m = 0; m < 3; ++m) s += F[a + 3 * m] * S[m + 3 * n];
            P[a + 3 * n] = s; /* (Id+Grad_u)*PK2 */
            for (int b = 0; b < 3; ++b)
              for (int l = 0; l < 3; ++l) {
                double v = (a == b) ? S[l + 3 * n] : 0.0; /* Grad_Test2_u*S */
                for (int p = 0; p < 3; ++p) v += F[a + 3 * p] * dS[p + 3 * n + 9 * b + 27 * l];
                D_(a, n, b, l) = v;
              }
          }
      }
      /* t(i alpha, j beta) = sum_{n,l} Z(i,n) D(alpha,n,beta,l) Z(j,l);  elem += coeff*t */
      if (order_mask & 2) {
        for (int j = 0; j < nd; ++j)
          for (int b = 0; b < Q; ++b)
            for (int i = 0; i < nd; ++i)
              for (int a = 0; a < Q; ++a) {
                double s = 0;
                for (int n = 0; n < N; ++n)
                  for (int l = 0; l < N; ++l) s += Z[i * N + n] * D_(a, n, b, l) * Z[j * N + l];
                t[(i * Q + a) + s1 * (j * Q + b)] = s;
              }
        for (int k = 0; k < s1 * s1; ++k) elem[k] += coeff * t[k];
      }
#undef D_
      if (order_mask & 1)
        for (int i = 0; i < nd; ++i)
          for (int a = 0; a < Q; ++a) {
            double s = 0;
            for (int n = 0; n < N; ++n) s += P[a + Q * n] * Z[i * N + n];
            relem[i * Q + a] += coeff * s;
          }
    }
    if (order_mask & 2) {
      double ninf = 0;
      for (int k = 0; k < s1 * s1; ++k) if (fabs(elem[k]) > ninf) ninf = fabs(elem[k]);
      if (ninf != 0.0) add_elem_matrix(res->cols, s1, s1, dofs, dofs, elem, ninf * 1e-14, sort);
    }
    if (order_mask & 1)
      for (int k = 0; k < s1; ++k) res->R[dofs[k]] += relem[k];
  }
  res->nnz = 0;
  for (int64_t j = 0; j < ndof; ++j) res->nnz += res->cols[j].n;
  free(elem); free(t); free(relem); free(Z); free(G); free(ue); free(dofs); free(sort);
  return res;
}

gfo_result *gfo_assemble_region(int dim, int64_t ne, int ng, const double *pts, const int32_t *conn, int nd, int Q,
                                const int64_t *elem_dof, int64_t ndof, int nq, const double *w, const double *gt_grad,
                                const double *phi, const double *gphi, int gt_linear, int family, const double *par,
                                const double *U, int order_mask, int64_t n_items, const int32_t *item_cv,
                                const int32_t *item_face, const int32_t *face_first, const int32_t *face_nq,
                                const double *ref_normals) {
  return gfo_assemble_fields(dim, ne, ng, pts, conn, nd, Q, elem_dof, ndof, nq, w, gt_grad, phi, gphi, gt_linear, family, par,
                             U, order_mask, n_items, item_cv, item_face, face_first, face_nq, ref_normals, 0, 0, NULL, NULL,
                             NULL, NULL);
}

gfo_result *gfo_assemble(int dim, int64_t ne, int ng, const double *pts, const int32_t *conn, int nd, int Q,
                         const int64_t *elem_dof, int64_t ndof, int nq, const double *w, const double *gt_grad,
                         const double *phi, const double *gphi, int gt_linear, int family, const double *par,
                         const double *U, int order_mask) {
  return gfo_assemble_region(dim, ne, ng, pts, conn, nd, Q, elem_dof, ndof, nq, w, gt_grad, phi, gphi, gt_linear, family,
                             par, U, order_mask, ne, NULL, NULL, NULL, NULL, NULL);
}

/* the material point alone, for the derivative check of tests/test_oracle.py (the restatement of
   abstract_hyperelastic_law::test_derivatives, getfem_nonlinear_elasticity.cc:298-347) */
void gfo_hyper_law(int family, const double *Gu, const double *par, double *S, double *dS) { hyper_law(family, Gu, par, S, dS); }

int64_t gfo_nnz(const gfo_result *r) { return r->nnz; }

/* gmm::csc_matrix::init_with_good_format (gmm_matrix.h:545-566) */
void gfo_get_csc(const gfo_result *r, int64_t *jc, int64_t *ir, double *pr) {
  int64_t k = 0;
  for (int64_t j = 0; j < r->ndof; ++j) {
    jc[j] = k;
    for (int64_t e = 0; e < r->cols[j].n; ++e, ++k) { ir[k] = r->cols[j].v[e].c; pr[k] = r->cols[j].v[e].e; }
  }
  jc[r->ndof] = k;
}
void gfo_get_residual(const gfo_result *r, double *R) { memcpy(R, r->R, sizeof(double) * (size_t)r->ndof); }
void gfo_free(gfo_result *r) {
  for (int64_t j = 0; j < r->ndof; ++j) free(r->cols[j].v);
  free(r->cols); free(r->R); free(r);
}
