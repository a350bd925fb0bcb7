// jit.cu -- the NVRTC route: an element kernel compiled at run time around a per-Gauss-point integrand given as SOURCE.
//
// The families of elem_kernel.cuh are closed forms picked by recognition.  A GWFL expression that is none of them still
// has, after the reference's own analysis and symbolic differentiation, an order-1 tree that is LINEAR in the test function
// and an order-2 tree that is BILINEAR in (Test, Test2) (C&E.cc:8750-9047 interprets exactly those trees).  For a scalar
// variable such a tree is, at a Gauss point, a function of (u, Grad u, constants) applied to tau = (Test_u, Grad_Test_u)
// [and tau2]: the host translates the tree into a C expression (getfem_b200/shim: jit_translate) and this module wraps it:
//     r(tau)        = form1(u, gu, par, tv, tg)                       linear in (tv, tg)
//     k(tau, tau2)  = form2(u, gu, par, tv, tg, t2v, t2g)             bilinear
// The kernel evaluates the forms with UNIT probes (tv = 1 or tg = e_k: literal constants, so the compiler folds each call
// into the coefficient it extracts), which gives the N+1 coefficients c_a and the (N+1)^2 coefficients C_ab of
//     r_e(i)    = sum_q w_q J sum_a c_a(q) T_a(phi_i)(q),        T_0 = value, T_k = d/dx_k
//     K_e(i,j)  = sum_q w_q J sum_ab C_ab(q) T_a(phi_i)(q) T_b(phi_j)(q)
// and writes the SAME stage / keep masks / element residuals as the generic kernel (drop rule included), so the pattern
// builder, the gather and everything downstream are unchanged.  NVRTC and the driver API are resolved with dlopen:
// libgfgpu.so has no link dependency on them, and a process that never creates a JIT term never loads them.
#include <dlfcn.h>

#include <cctype>
#include <mutex>
#include <sstream>

#include "common.cuh"

namespace gf {

// ---- minimal driver / NVRTC surface (types as the headers declare them: opaque pointers and ints)
typedef struct CUmod_st *CUmodule_t;
typedef struct CUfunc_st *CUfunction_t;
typedef struct _nvrtcProgram *nvrtcProgram_t;
struct JitApi {
  void *hn = nullptr, *hc = nullptr;
  int (*CreateProgram)(nvrtcProgram_t *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
  int (*CompileProgram)(nvrtcProgram_t, int, const char *const *) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram_t, size_t *) = nullptr;
  int (*GetProgramLog)(nvrtcProgram_t, char *) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram_t, size_t *) = nullptr;
  int (*GetCUBIN)(nvrtcProgram_t, char *) = nullptr;
  int (*DestroyProgram)(nvrtcProgram_t *) = nullptr;
  int (*ModuleLoadData)(CUmodule_t *, const void *) = nullptr;
  int (*ModuleGetFunction)(CUfunction_t *, CUmodule_t, const char *) = nullptr;
  int (*ModuleUnload)(CUmodule_t) = nullptr;
  int (*FuncSetAttribute)(CUfunction_t, int, int) = nullptr;
  int (*LaunchKernel)(CUfunction_t, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, cudaStream_t, void **,
                      void **) = nullptr;
};

static JitApi &jit_api(bool need_driver = true) {
  static JitApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char *n : {"libnvrtc.so.12", "libnvrtc.so"})
      if (!a.hn) a.hn = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    for (const char *n : {"libcuda.so.1", "libcuda.so"})
      if (!a.hc) a.hc = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
#define GF_SYM(h, field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name))
    if (a.hn) {
      GF_SYM(a.hn, CreateProgram, "nvrtcCreateProgram");
      GF_SYM(a.hn, CompileProgram, "nvrtcCompileProgram");
      GF_SYM(a.hn, GetProgramLogSize, "nvrtcGetProgramLogSize");
      GF_SYM(a.hn, GetProgramLog, "nvrtcGetProgramLog");
      GF_SYM(a.hn, GetCUBINSize, "nvrtcGetCUBINSize");
      GF_SYM(a.hn, GetCUBIN, "nvrtcGetCUBIN");
      GF_SYM(a.hn, DestroyProgram, "nvrtcDestroyProgram");
    }
    if (a.hc) {
      GF_SYM(a.hc, ModuleLoadData, "cuModuleLoadData");
      GF_SYM(a.hc, ModuleGetFunction, "cuModuleGetFunction");
      GF_SYM(a.hc, ModuleUnload, "cuModuleUnload");
      GF_SYM(a.hc, FuncSetAttribute, "cuFuncSetAttribute");
      GF_SYM(a.hc, LaunchKernel, "cuLaunchKernel");
    }
#undef GF_SYM
  });
  GF_REQUIRE(a.hn && a.CreateProgram && a.CompileProgram && a.GetCUBIN && a.GetCUBINSize && a.GetProgramLog && a.GetProgramLogSize &&
                 a.DestroyProgram,
             "NVRTC (libnvrtc.so.12) could not be loaded: JIT terms need it");
  GF_REQUIRE(!need_driver || (a.hc && a.ModuleLoadData && a.ModuleGetFunction && a.LaunchKernel && a.FuncSetAttribute && a.ModuleUnload),
             "the CUDA driver library (libcuda.so.1) could not be loaded: JIT terms need it");
  return a;
}

// ---- the kernel template.  GF_N (2 or 3), GF_FORM1, GF_FORM2 are defined on the NVRTC command line / prepended.
static const char *kJitSource = R"GFJIT(
struct vec { double v[GF_N]; };
__device__ __forceinline__ vec operator+(vec a, vec b) { vec r; for (int k = 0; k < GF_N; ++k) r.v[k] = a.v[k] + b.v[k]; return r; }
__device__ __forceinline__ vec operator-(vec a, vec b) { vec r; for (int k = 0; k < GF_N; ++k) r.v[k] = a.v[k] - b.v[k]; return r; }
__device__ __forceinline__ vec operator-(vec a) { vec r; for (int k = 0; k < GF_N; ++k) r.v[k] = -a.v[k]; return r; }
__device__ __forceinline__ vec operator*(double s, vec a) { vec r; for (int k = 0; k < GF_N; ++k) r.v[k] = s * a.v[k]; return r; }
__device__ __forceinline__ vec operator*(vec a, double s) { vec r; for (int k = 0; k < GF_N; ++k) r.v[k] = a.v[k] * s; return r; }
__device__ __forceinline__ vec operator/(vec a, double s) { vec r; for (int k = 0; k < GF_N; ++k) r.v[k] = a.v[k] / s; return r; }
__device__ __forceinline__ double dot(vec a, vec b) { double s = 0; for (int k = 0; k < GF_N; ++k) s += a.v[k] * b.v[k]; return s; }
__device__ __forceinline__ double dot(double a, double b) { return a * b; }
__device__ __forceinline__ vec dot(double a, vec b) { return a * b; }
__device__ __forceinline__ vec dot(vec a, double b) { return a * b; }
__device__ __forceinline__ double normsqr(vec a) { return dot(a, a); }
__device__ __forceinline__ double normsqr(double a) { return a * a; }
__device__ __forceinline__ double gnorm(vec a) { return sqrt(dot(a, a)); }
__device__ __forceinline__ double gnorm(double a) { return fabs(a); }
__device__ __forceinline__ vec cross(vec a, vec b) {  // Cross_product of two 3D vectors (bilinear: its derivatives are written inline)
  vec r;
  for (int i = 0; i < GF_N; ++i) r.v[i] = a.v[(i + 1) % GF_N] * b.v[(i + 2) % GF_N] - a.v[(i + 2) % GF_N] * b.v[(i + 1) % GF_N];
  return r;
}
__device__ __forceinline__ vec mkvec(double a, double b, double c) { vec r; r.v[0] = a; r.v[1] = b; if (GF_N > 2) r.v[GF_N - 1] = c; return r; }
__device__ __forceinline__ vec unit(int k) { vec r; for (int j = 0; j < GF_N; ++j) r.v[j] = j == k ? 1.0 : 0.0; return r; }
__device__ __forceinline__ double sqr(double x) { return x * x; }
__device__ __forceinline__ double pos_part(double x) { return x > 0 ? x : 0.0; }
__device__ __forceinline__ double neg_part(double x) { return x < 0 ? -x : 0.0; }
__device__ __forceinline__ double half_sqr_pos_part(double x) { return x > 0 ? 0.5 * x * x : 0.0; }
__device__ __forceinline__ double half_sqr_neg_part(double x) { return x < 0 ? 0.5 * x * x : 0.0; }
__device__ __forceinline__ double sqr_pos_part(double x) { return x > 0 ? x * x : 0.0; }
__device__ __forceinline__ double sqr_neg_part(double x) { return x < 0 ? x * x : 0.0; }
// first derivatives of the predefined scalar functions, under the names the reference's symbolic differentiation gives them
__device__ __forceinline__ double DER_PDFUNC_SQRT(double t) { return 0.5 / sqrt(t); }
__device__ __forceinline__ double DER_PDFUNC1_POW(double t, double e) { return e * pow(t, e - 1.0); }
__device__ __forceinline__ double DER_PDFUNC2_POW(double t, double e) { return pow(t, e) * log(fabs(t)); }
__device__ __forceinline__ double DER_PDFUNC_LOG(double t) { return 1.0 / t; }
__device__ __forceinline__ double DER_PDFUNC_LOG10(double t) { return 1.0 / (t * log(10.0)); }
__device__ __forceinline__ double DER_PDFUNC_TANH(double t) { const double h = tanh(t); return 1.0 - h * h; }
__device__ __forceinline__ double DER_PDFUNC_ASINH(double t) { return 1.0 / sqrt(t * t + 1.0); }
__device__ __forceinline__ double DER_PDFUNC_ACOSH(double t) { return 1.0 / sqrt(t * t - 1.0); }
__device__ __forceinline__ double DER_PDFUNC_ATANH(double t) { return 1.0 / (1.0 - t * t); }
__device__ __forceinline__ double DER_PDFUNC_COS(double t) { return -sin(t); }
__device__ __forceinline__ double DER_PDFUNC_TAN(double t) { const double h = tan(t); return 1.0 + h * h; }
__device__ __forceinline__ double DER_PDFUNC_ASIN(double t) { return 1.0 / sqrt(1.0 - t * t); }
__device__ __forceinline__ double DER_PDFUNC_ACOS(double t) { return -1.0 / sqrt(1.0 - t * t); }
__device__ __forceinline__ double DER_PDFUNC_ATAN(double t) { return 1.0 / (1.0 + t * t); }
__device__ __forceinline__ double DER_PDFUNC1_ATAN2(double t, double v) { return v / (t * t + v * v); }
__device__ __forceinline__ double DER_PDFUNC2_ATAN2(double t, double v) { return -t / (t * t + v * v); }
__device__ __forceinline__ double DER_PDFUNC_ERF(double t) { return exp(-t * t) * 1.1283791670955126; }
__device__ __forceinline__ double DER_PDFUNC_ERFC(double t) { return -exp(-t * t) * 1.1283791670955126; }
__device__ __forceinline__ double Heaviside(double x) { return x < 0 ? 0.0 : 1.0; }
// second derivatives, under the names the reference gives the derivative of a derivative that is defined by an expression
// (getfem_generic_assembly_functions_and_operators.cc:418-490: "-0.25/(t*sqrt(t))", "-1/sqr(t)", ...)
__device__ __forceinline__ double DER_PDFUNC_DER_PDFUNC_SQRT(double t) { return -0.25 / (t * sqrt(t)); }
__device__ __forceinline__ double DER_PDFUNC_DER_PDFUNC_LOG(double t) { return -1.0 / (t * t); }
__device__ __forceinline__ double DER_PDFUNC_DER_PDFUNC_LOG10(double t) { return -1.0 / (t * t * log(10.0)); }
__device__ __forceinline__ double DER_PDFUNC_DER_PDFUNC_TANH(double t) { const double h = tanh(t); return 2.0 * h * (h * h - 1.0); }
__device__ __forceinline__ double DER_PDFUNC_DER_PDFUNC_ASINH(double t) { return -t / pow(t * t + 1.0, 1.5); }
__device__ __forceinline__ double DER_PDFUNC_DER_PDFUNC_ACOSH(double t) { return -t / pow(t * t - 1.0, 1.5); }
__device__ __forceinline__ double DER_PDFUNC_DER_PDFUNC_ATANH(double t) { return 2.0 * t / ((1.0 - t * t) * (1.0 - t * t)); }
__device__ __forceinline__ double DER_PDFUNC_DER_PDFUNC_COS(double t) { return -cos(t); }
__device__ __forceinline__ double DER_PDFUNC_DER_PDFUNC_TAN(double t) { const double c = cos(t); return 2.0 * tan(t) / (c * c); }
__device__ __forceinline__ double DER_PDFUNC_DER_PDFUNC_ASIN(double t) { return t / pow(1.0 - t * t, 1.5); }
__device__ __forceinline__ double DER_PDFUNC_DER_PDFUNC_ACOS(double t) { return -t / pow(1.0 - t * t, 1.5); }
__device__ __forceinline__ double DER_PDFUNC_DER_PDFUNC_ATAN(double t) { return -2.0 * t / ((1.0 + t * t) * (1.0 + t * t)); }
__device__ __forceinline__ double DER_PDFUNC_NEG_PART(double t) { return t >= 0 ? 0.0 : -1.0; }  // ga_der_neg_part (:163)
__device__ __forceinline__ double DER_PDFUNC1_DER_PDFUNC1_POW(double t, double e) { return e * (e - 1.0) * pow(t, e - 2.0); }
__device__ __forceinline__ double DER_PDFUNC2_DER_PDFUNC1_POW(double t, double e) { return pow(t, e - 1.0) * (e * log(t) + 1.0); }
__device__ __forceinline__ double DER_PDFUNC1_DER_PDFUNC2_POW(double t, double e) { return pow(t, e - 1.0) * (e * log(t) + 1.0); }
__device__ __forceinline__ double DER_PDFUNC2_DER_PDFUNC2_POW(double t, double e) { const double l = log(t); return pow(t, e) * l * l; }
__device__ __forceinline__ double sign(double x) { return x >= 0 ? 1.0 : -1.0; }  // ga_sign: +1 at 0 (functions_and_operators.cc:110)
// sinc with the reference's series below 1e-4 (:96-127), max / min derivatives (:165-168: ties go to BOTH arguments)
__device__ __forceinline__ double sinc(double t) { const double t2 = t * t; return fabs(t) < 1e-4 ? 1.0 - t2 / 6.0 + t2 * t2 / 120.0 : sin(t) / t; }
__device__ __forceinline__ double DER_PDFUNC_SINC(double t) {
  const double t2 = t * t;
  return fabs(t) < 1e-4 ? -t / 3.0 + t * t2 / 30.0 - t * t2 * t2 / 840.0 : (t * cos(t) - sin(t)) / (t * t);
}
__device__ __forceinline__ double DER2_PDFUNC_SINC(double t) {
  const double t2 = t * t;
  return fabs(t) < 1e-4 ? -1.0 / 3.0 + t2 / 10.0 - t2 * t2 / 168.0 : ((2.0 - t * t) * sin(t) - 2.0 * t * cos(t)) / (t * t * t);
}
__device__ __forceinline__ double DER_PDFUNC1_MAX(double t, double u) { return t - u >= 0 ? 1.0 : 0.0; }
__device__ __forceinline__ double DER_PDFUNC2_MAX(double t, double u) { return u - t >= 0 ? 1.0 : 0.0; }

// matrices of the mesh dimension: m[c][k] (for Grad_u: component c, direction k, the GWFL convention)
struct mat { double m[GF_N][GF_N]; };
__device__ __forceinline__ mat operator+(mat a, mat b) { mat r; for (int i = 0; i < GF_N; ++i) for (int j = 0; j < GF_N; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r; }
__device__ __forceinline__ mat operator-(mat a, mat b) { mat r; for (int i = 0; i < GF_N; ++i) for (int j = 0; j < GF_N; ++j) r.m[i][j] = a.m[i][j] - b.m[i][j]; return r; }
__device__ __forceinline__ mat operator-(mat a) { mat r; for (int i = 0; i < GF_N; ++i) for (int j = 0; j < GF_N; ++j) r.m[i][j] = -a.m[i][j]; return r; }
__device__ __forceinline__ mat operator*(double s, mat a) { mat r; for (int i = 0; i < GF_N; ++i) for (int j = 0; j < GF_N; ++j) r.m[i][j] = s * a.m[i][j]; return r; }
__device__ __forceinline__ mat operator*(mat a, double s) { return s * a; }
__device__ __forceinline__ mat operator/(mat a, double s) { mat r; for (int i = 0; i < GF_N; ++i) for (int j = 0; j < GF_N; ++j) r.m[i][j] = a.m[i][j] / s; return r; }
__device__ __forceinline__ mat operator*(mat a, mat b) {
  mat r;
  for (int i = 0; i < GF_N; ++i) for (int j = 0; j < GF_N; ++j) { double t = 0; for (int k = 0; k < GF_N; ++k) t += a.m[i][k] * b.m[k][j]; r.m[i][j] = t; }
  return r;
}
__device__ __forceinline__ vec operator*(mat a, vec b) { vec r; for (int i = 0; i < GF_N; ++i) { double t = 0; for (int k = 0; k < GF_N; ++k) t += a.m[i][k] * b.v[k]; r.v[i] = t; } return r; }
__device__ __forceinline__ mat dot(mat a, mat b) { return a * b; }          // contraction of the last index with the first
__device__ __forceinline__ vec dot(mat a, vec b) { return a * b; }
__device__ __forceinline__ vec dot(vec a, mat b) { vec r; for (int j = 0; j < GF_N; ++j) { double t = 0; for (int k = 0; k < GF_N; ++k) t += a.v[k] * b.m[k][j]; r.v[j] = t; } return r; }
__device__ __forceinline__ mat dot(double a, mat b) { return a * b; }
__device__ __forceinline__ mat dot(mat a, double b) { return b * a; }
__device__ __forceinline__ double ddot(mat a, mat b) { double t = 0; for (int i = 0; i < GF_N; ++i) for (int j = 0; j < GF_N; ++j) t += a.m[i][j] * b.m[i][j]; return t; }
__device__ __forceinline__ double ddot(vec a, vec b) { return dot(a, b); }
__device__ __forceinline__ double ddot(double a, double b) { return a * b; }
__device__ __forceinline__ mat transp(mat a) { mat r; for (int i = 0; i < GF_N; ++i) for (int j = 0; j < GF_N; ++j) r.m[i][j] = a.m[j][i]; return r; }
__device__ __forceinline__ vec transp(vec a) { return a; }
__device__ __forceinline__ double transp(double a) { return a; }
__device__ __forceinline__ vec mcol(mat a, int j) { vec r; for (int i = 0; i < GF_N; ++i) r.v[i] = a.m[i][j]; return r; }  // M(:,j)
__device__ __forceinline__ vec mrow(mat a, int i) { vec r; for (int j = 0; j < GF_N; ++j) r.v[j] = a.m[i][j]; return r; }  // M(i,:)
__device__ __forceinline__ double trace(mat a) { double t = 0; for (int i = 0; i < GF_N; ++i) t += a.m[i][i]; return t; }
__device__ __forceinline__ mat sym(mat a) { return 0.5 * (a + transp(a)); }
__device__ __forceinline__ mat skew(mat a) { return 0.5 * (a - transp(a)); }
__device__ __forceinline__ mat deviator(mat a) { mat r = a; const double t = trace(a) / GF_N; for (int i = 0; i < GF_N; ++i) r.m[i][i] -= t; return r; }
__device__ __forceinline__ double normsqr(mat a) { return ddot(a, a); }
__device__ __forceinline__ double gnorm(mat a) { return sqrt(ddot(a, a)); }
__device__ __forceinline__ mat outer(vec a, vec b) { mat r; for (int i = 0; i < GF_N; ++i) for (int j = 0; j < GF_N; ++j) r.m[i][j] = a.v[i] * b.v[j]; return r; }
// mkmat takes the entries COLUMN-major, like GetFEM stores a constant matrix (first index fastest)
__device__ __forceinline__ mat mkmat(double a0, double a1, double a2, double a3, double a4, double a5, double a6, double a7, double a8) {
  const double a[9] = {a0, a1, a2, a3, a4, a5, a6, a7, a8};
  mat r;
  for (int j = 0; j < GF_N; ++j) for (int i = 0; i < GF_N; ++i) r.m[i][j] = a[i + GF_N * j];
  return r;
}

// Norm (norm_operator, getfem_generic_assembly_functions_and_operators.cc:181-218): derivative x/|x| (0 at x = 0), second
// derivative (|x|^2 Id - x (x) x)/|x|^3 in two directions (the reference clamps |x| at 1e-25 in the 1/|x| term only)
__device__ __forceinline__ vec dnorm(vec x) { const double n = gnorm(x); return n == 0.0 ? 0.0 * x : x / n; }
__device__ __forceinline__ mat dnorm(mat x) { const double n = gnorm(x); return n == 0.0 ? 0.0 * x : x / n; }
__device__ __forceinline__ double d2norm(vec x, vec h, vec k) {
  const double n = gnorm(x), n3 = n * n * n;
  return ddot(h, k) / (n < 1e-25 ? 1e-25 : n) - ddot(x, h) * ddot(x, k) / n3;
}
__device__ __forceinline__ double d2norm(mat x, mat h, mat k) {
  const double n = gnorm(x), n3 = n * n * n;
  return ddot(h, k) / (n < 1e-25 ? 1e-25 : n) - ddot(x, h) * ddot(x, k) / n3;
}

// nonlinear operators of one square matrix and their first derivatives in a direction h (the reference's Det, Inv,
// Right/Left_Cauchy_Green, Green_Lagrangian, Matrix_i2: getfem_generic_assembly_functions_and_operators.cc,
// getfem_nonlinear_elasticity.cc:1930-2040)
__device__ __forceinline__ double det(mat a) {
  if (GF_N == 2) return a.m[0][0] * a.m[1][1] - a.m[0][1] * a.m[1][0];
  return a.m[0][0] * (a.m[1][1] * a.m[GF_N - 1][GF_N - 1] - a.m[1][GF_N - 1] * a.m[GF_N - 1][1]) -
         a.m[0][1] * (a.m[1][0] * a.m[GF_N - 1][GF_N - 1] - a.m[1][GF_N - 1] * a.m[GF_N - 1][0]) +
         a.m[0][GF_N - 1] * (a.m[1][0] * a.m[GF_N - 1][1] - a.m[1][1] * a.m[GF_N - 1][0]);
}
__device__ __forceinline__ mat inv(mat a) {
  mat r;
  const double id = 1.0 / det(a);
  if (GF_N == 2) {
    r.m[0][0] = a.m[1][1] * id; r.m[0][1] = -a.m[0][1] * id; r.m[1][0] = -a.m[1][0] * id; r.m[1][1] = a.m[0][0] * id;
  } else {
    for (int i = 0; i < GF_N; ++i)
      for (int j = 0; j < GF_N; ++j) {  // r(i,j) = cofactor(j,i) / det, cyclic indices
        const int j1 = (j + 1) % GF_N, j2 = (j + 2) % GF_N, i1 = (i + 1) % GF_N, i2 = (i + 2) % GF_N;
        r.m[i][j] = (a.m[j1][i1] * a.m[j2][i2] - a.m[j1][i2] * a.m[j2][i1]) * id;
      }
  }
  return r;
}
__device__ __forceinline__ double ddet(mat a, mat h) { return det(a) * trace(inv(a) * h); }
__device__ __forceinline__ mat dinv(mat a, mat h) { const mat ia = inv(a); return -(ia * h * ia); }
__device__ __forceinline__ mat rcg(mat f) { return transp(f) * f; }
__device__ __forceinline__ mat drcg(mat f, mat h) { return transp(h) * f + transp(f) * h; }
__device__ __forceinline__ mat lcg(mat f) { return f * transp(f); }
__device__ __forceinline__ mat dlcg(mat f, mat h) { return h * transp(f) + f * transp(h); }
__device__ __forceinline__ mat glag(mat f) { mat r = 0.5 * (transp(f) * f); for (int i = 0; i < GF_N; ++i) r.m[i][i] -= 0.5; return r; }
__device__ __forceinline__ mat dglag(mat f, mat h) { return 0.5 * (transp(h) * f + transp(f) * h); }
__device__ __forceinline__ double mat_i2(mat a) { const double t = trace(a); return 0.5 * (t * t - trace(a * a)); }
__device__ __forceinline__ double dmat_i2(mat a, mat h) { return trace(a) * trace(h) - trace(a * h); }
__device__ __forceinline__ double d2mat_i2(mat a, mat h, mat k);
// Matrix_j1(A) = tr(A) det(A)^(-1/3), Matrix_j2(A) = i2(A) det(A)^(-2/3) (getfem_nonlinear_elasticity.cc:1351-1500; 1e200 where det <= 0)
__device__ __forceinline__ double mat_j1(mat a) { const double d = det(a); return d > 0.0 ? trace(a) * pow(d, -1.0 / 3.0) : 1e200; }
__device__ __forceinline__ double dmat_j1(mat a, mat h) {
  const double d = det(a);
  if (!(d > 0.0)) return 1e200;
  return pow(d, -1.0 / 3.0) * (trace(h) - trace(a) * trace(inv(a) * h) / 3.0);
}
__device__ __forceinline__ double d2mat_j1(mat a, mat h, mat k) {
  const mat ia = inv(a), ih = ia * h, ik = ia * k;
  const double r = pow(det(a), -1.0 / 3.0), i1 = trace(a), th = trace(ih), tk = trace(ik);
  const double t = trace(h) - i1 * th / 3.0, dt = -trace(k) * th / 3.0 + i1 * trace(ik * ih) / 3.0;
  return r * (dt - tk * t / 3.0);
}
__device__ __forceinline__ double mat_j2(mat a) { const double d = det(a); return d > 0.0 ? mat_i2(a) * pow(d, -2.0 / 3.0) : 1e200; }
__device__ __forceinline__ double dmat_j2(mat a, mat h) {
  const double d = det(a);
  if (!(d > 0.0)) return 1e200;
  return pow(d, -2.0 / 3.0) * (dmat_i2(a, h) - 2.0 * mat_i2(a) * trace(inv(a) * h) / 3.0);
}
__device__ __forceinline__ double d2mat_j2(mat a, mat h, mat k) {
  const mat ia = inv(a), ih = ia * h, ik = ia * k;
  const double r2 = pow(det(a), -2.0 / 3.0), i2 = mat_i2(a), th = trace(ih), tk = trace(ik);
  const double sh = dmat_i2(a, h) - 2.0 * i2 * th / 3.0;
  const double ds = d2mat_i2(a, h, k) - 2.0 * dmat_i2(a, k) * th / 3.0 + 2.0 * i2 * trace(ik * ih) / 3.0;
  return r2 * (ds - 2.0 * tk * sh / 3.0);
}
// second derivatives in two directions (symmetric in h, k): "(Derivative_1_1_Op(A):H2):H1" of an order-2 tree derived from a potential
__device__ __forceinline__ double d2det(mat a, mat h, mat k) {
  const mat ia = inv(a), ih = ia * h, ik = ia * k;
  return det(a) * (trace(ih) * trace(ik) - trace(ih * ik));
}
__device__ __forceinline__ mat d2inv(mat a, mat h, mat k) { const mat ia = inv(a), ih = ia * h, ik = ia * k; return ih * ik * ia + ik * ih * ia; }
__device__ __forceinline__ mat d2rcg(mat f, mat h, mat k) { return transp(h) * k + transp(k) * h; }
__device__ __forceinline__ mat d2lcg(mat f, mat h, mat k) { return h * transp(k) + k * transp(h); }
__device__ __forceinline__ mat d2glag(mat f, mat h, mat k) { return 0.5 * (transp(h) * k + transp(k) * h); }
__device__ __forceinline__ double d2mat_i2(mat a, mat h, mat k) { return trace(h) * trace(k) - trace(h * k); }

// Saint-Venant Kirchhoff, the law the reference defines in any dimension (getfem_nonlinear_elasticity.cc:503-540):
// E = (G + G' + G'G)/2, S = lambda tr(E) I + 2 mu E; dS[H] with dE = (H + H' + G'H + H'G)/2
__device__ __forceinline__ mat svk_pk2(mat g, double lam, double mu) {
  const mat gt = transp(g), E = 0.5 * (g + gt + gt * g);
  mat r = (2.0 * mu) * E;
  const double t = lam * trace(E);
  for (int i = 0; i < GF_N; ++i) r.m[i][i] += t;
  return r;
}
__device__ __forceinline__ mat svk_dpk2(mat g, double lam, double mu, mat h) {
  const mat gt = transp(g), ht = transp(h), dE = 0.5 * (h + ht + gt * h + ht * g);
  mat r = (2.0 * mu) * dE;
  const double t = lam * trace(dE);
  for (int i = 0; i < GF_N; ++i) r.m[i][i] += t;
  return r;
}

// compressible neo-Hookean laws (Neo_Hookean_hyperelastic_law::sigma / grad_sigma, getfem_nonlinear_elasticity.cc:612-702) on
// GF_N x GF_N tensors -- in 2D this is the plane-strain wrapper (:906-945: C33 = 1, so det C and the 2 x 2 block of C^-1 are
// those of the 2 x 2 tensor): S = mu I + k(i3) C^-1, i3 = det C, k = lambda/2 (i3 - 1) - mu (Ciarlet), lambda/2 log(i3) - mu
// (Bonet); + 1e200 C where det F <= 0 (the reference's penalty on inverted elements)
__device__ __forceinline__ mat nh_pk2(mat g, double lam, double mu, int bonet) {
  mat f = g;
  for (int i = 0; i < GF_N; ++i) f.m[i][i] += 1.0;
  const mat c = transp(f) * f, ci = inv(c);
  const double i3 = det(c), k = bonet ? 0.5 * lam * log(i3) - mu : 0.5 * lam * (i3 - 1.0) - mu;
  mat r = k * ci;
  for (int i = 0; i < GF_N; ++i) r.m[i][i] += mu;
  if (det(f) <= 0.0) r = r + 1e200 * c;
  return r;
}
__device__ __forceinline__ mat nh_dpk2(mat g, double lam, double mu, int bonet, mat h) {
  mat f = g;
  for (int i = 0; i < GF_N; ++i) f.m[i][i] += 1.0;
  const mat c = transp(f) * f, ci = inv(c), dc = transp(h) * f + transp(f) * h;
  const double i3 = det(c), k = bonet ? 0.5 * lam * log(i3) - mu : 0.5 * lam * (i3 - 1.0) - mu;
  const double di3 = i3 * trace(ci * dc), dk = bonet ? 0.5 * lam * di3 / i3 : 0.5 * lam * di3;
  return dk * ci - k * (ci * dc * ci);
}

// Isotropic laws given through the invariants of C = F'F (compute_invariants, getfem_nonlinear_elasticity.cc:45-262):
//   law 0  compressible Mooney-Rivlin (:503-607)  W = c1 (j1 - 3) + c2 (j2 - 3) + d1 (sqrt(i3) - 1)^2, j1 = i1 i3^(-1/3), j2 = i2 i3^(-2/3)
//   law 1  Ciarlet-Geymonat (:817-888)            W = a i1 + b i2 + c i3 - d/2 log(i3) + e  (a, b, c, d from lambda, mu, a)
//   law 2  generalized Blatz-Ko (:706-815)        W = z^n, z = a i1 + b sqrt(i3) + c i2 / i3 + d
// S = 2 sum_a W_a di_a, dS[H] = 2 sum_a [(sum_b W_ab (di_b : dC)) di_a + W_a d2i_a[dC]], dC = H'F + F'H, with
// di1 = I, di2 = i1 I - C, di3 = i3 C^-1; d2i2[dC] = tr(dC) I - dC, d2i3[dC] = i3 (tr(C^-1 dC) C^-1 - C^-1 dC C^-1).
// On 2 x 2 tensors this is the PLANE STRAIN wrapper (plane_strain_hyperelastic_law, :906-945: the strain embedded in a 3 x 3 one
// with zeros, C33 = 1): i1 and i2 are those of the embedded tensor, i3 and the 2 x 2 blocks of the gradients follow from C itself.
__device__ __forceinline__ void iso_law_w(int law, double i1, double i2, double i3, const double *p, double *Wa, double *Wab) {
  // Wab: 11, 12, 13, 22, 23, 33
  for (int k = 0; k < 6; ++k) Wab[k] = 0.0;
  if (law == 0) {
    const double c1 = p[0], c2 = p[1], d1 = p[2], r13 = pow(i3, -1.0 / 3.0), r23 = r13 * r13;
    Wa[0] = c1 * r13;
    Wa[1] = c2 * r23;
    Wa[2] = -c1 * i1 * r13 / (3.0 * i3) - 2.0 * c2 * i2 * r23 / (3.0 * i3) + d1 * (1.0 - 1.0 / sqrt(i3));
    Wab[2] = -c1 * r13 / (3.0 * i3);
    Wab[4] = -2.0 * c2 * r23 / (3.0 * i3);
    Wab[5] = 4.0 * c1 * i1 * r13 / (9.0 * i3 * i3) + 10.0 * c2 * i2 * r23 / (9.0 * i3 * i3) + 0.5 * d1 / (i3 * sqrt(i3));
  } else if (law == 1) {
    const double a = p[2], b = 0.5 * p[1] - p[2], c = 0.25 * p[0] - 0.5 * p[1] + p[2], d = 0.5 * p[0] + p[1];
    Wa[0] = a;
    Wa[1] = b;
    Wa[2] = c - 0.5 * d / i3;
    Wab[5] = 0.5 * d / (i3 * i3);
  } else {
    const double a = p[0], b = p[1], c = p[2], d = p[3], n = p[4], s3 = sqrt(i3);
    const double z = a * i1 + b * s3 + c * i2 / i3 + d;
    const double za[3] = {a, c / i3, 0.5 * b / s3 - c * i2 / (i3 * i3)};
    const double zab[6] = {0.0, 0.0, 0.0, 0.0, -c / (i3 * i3), -0.25 * b / (i3 * s3) + 2.0 * c * i2 / (i3 * i3 * i3)};
    const double w1 = n * pow(z, n - 1.0), w2 = n * (n - 1.0) * pow(z, n - 2.0);
    for (int k = 0; k < 3; ++k) Wa[k] = w1 * za[k];
    const int ia[6] = {0, 0, 0, 1, 1, 2}, ib[6] = {0, 1, 2, 1, 2, 2};
    for (int k = 0; k < 6; ++k) Wab[k] = w2 * za[ia[k]] * za[ib[k]] + w1 * zab[k];
  }
}
__device__ __forceinline__ mat iso_pk2(mat g, int law, double p0, double p1, double p2, double p3, double p4) {
  const double p[5] = {p0, p1, p2, p3, p4};
  mat f = g;
  for (int i = 0; i < GF_N; ++i) f.m[i][i] += 1.0;
  const mat c = transp(f) * f, ci = inv(c);
  const double emb = 3.0 - GF_N, i1 = trace(c) + emb, i2 = 0.5 * (i1 * i1 - trace(c * c) - emb), i3 = det(c);
  double Wa[3], Wab[6];
  iso_law_w(law, i1, i2, i3, p, Wa, Wab);
  mat r = (2.0 * Wa[2] * i3) * ci - (2.0 * Wa[1]) * c;
  for (int i = 0; i < GF_N; ++i) r.m[i][i] += 2.0 * (Wa[0] + Wa[1] * i1);
  if (det(f) <= 0.0) r = r + 1e200 * c;  // the reference's penalty on inverted elements
  return r;
}
__device__ __forceinline__ mat iso_dpk2(mat g, int law, double p0, double p1, double p2, double p3, double p4, mat h) {
  const double p[5] = {p0, p1, p2, p3, p4};
  mat f = g;
  for (int i = 0; i < GF_N; ++i) f.m[i][i] += 1.0;
  const mat c = transp(f) * f, ci = inv(c), dc = transp(h) * f + transp(f) * h, cidc = ci * dc;
  const double emb = 3.0 - GF_N, i1 = trace(c) + emb, i2 = 0.5 * (i1 * i1 - trace(c * c) - emb), i3 = det(c);
  double Wa[3], Wab[6];
  iso_law_w(law, i1, i2, i3, p, Wa, Wab);
  // di_b : dC
  const double t = trace(dc), q1 = t, q2 = i1 * t - trace(c * dc), q3 = i3 * trace(cidc);
  const double s1 = Wab[0] * q1 + Wab[1] * q2 + Wab[2] * q3, s2 = Wab[1] * q1 + Wab[3] * q2 + Wab[4] * q3,
               s3 = Wab[2] * q1 + Wab[4] * q2 + Wab[5] * q3;
  // sum_a s_a di_a + W_2 d2i2[dC] + W_3 d2i3[dC]
  mat r = (2.0 * (s3 * i3 + Wa[2] * q3)) * ci - (2.0 * s2) * c - (2.0 * Wa[1]) * dc - (2.0 * Wa[2] * i3) * (cidc * ci);
  for (int i = 0; i < GF_N; ++i) r.m[i][i] += 2.0 * (s1 + s2 * i1 + Wa[1] * t);
  return r;
}

#if GF_Q == 1
__device__ __forceinline__ double gf_form0(double u, vec gu, vec X, vec Normal, const double *fld, vec vfld, const double *par) { return GF_FORM0; }
__device__ __forceinline__ double gf_form1(double u, vec gu, vec X, vec Normal, const double *fld, vec vfld, const double *par, double tv, vec tg) { return GF_FORM1; }
__device__ __forceinline__ double gf_form2(double u, vec gu, vec X, vec Normal, const double *fld, vec vfld, const double *par, double tv, vec tg, double t2v, vec t2g) {
  return GF_FORM2;
}

extern "C" __global__ void __launch_bounds__(128)
gf_jit_elem(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z, const int *__restrict__ conn,
            const int *__restrict__ edof, const double *__restrict__ U, const double *__restrict__ w, const double *__restrict__ gt_grad,
            const double *__restrict__ phi, const double *__restrict__ gphi, const double *__restrict__ gt_val,
            const signed char *__restrict__ face, const double *__restrict__ fnormal,
            const int *__restrict__ dedof, const double *__restrict__ dphi, const double *__restrict__ dvals0,
            const double *__restrict__ dvals1, int nfields, int nd_d, int fqdim,
            const double *__restrict__ par, int ng, int nq, int nd,
            long long e0, long long ne, double alpha, double *__restrict__ stage, unsigned short *__restrict__ emask,
            double *__restrict__ rstage, double *__restrict__ epot) {
  constexpr int N = GF_N, NA = GF_N + 1;
  extern __shared__ double sm[];
  double *sG = sm;                       // N x ng
  double *sU = sG + N * ng;              // nd
  double *sT = sU + nd;                  // nq x nd x NA : value and physical gradient of every basis function
  double *sC1 = sT + (size_t)nq * nd * NA;   // nq x NA      (w J folded in)
  double *sC2 = sC1 + (size_t)nq * NA;       // nq x NA x NA
  double *sK = sC2 + (size_t)nq * NA * NA;   // nd x nd
  double *sP0 = sK + (size_t)nd * nd;        // nq: the order-0 integrand (w J folded in)
  __shared__ double sRed[4];
  const int tid = threadIdx.x;
  for (long long el = blockIdx.x; el < ne; el += gridDim.x) {
    const long long e = e0 + el;
    // tables of this item: the volume points, or (boundary faces, C&E.cc:8827-8848) the nq points of its face -- the face
    // tables are laid out face after face
    const int fc = face ? face[e] : 0;
    const double *wE = w + (size_t)fc * nq, *gtE = gt_grad + (size_t)fc * nq * ng * N, *phiE = phi + (size_t)fc * nq * nd;
    const double *gphiE = gphi + (size_t)fc * nq * nd * N, *gtvE = gt_val ? gt_val + (size_t)fc * nq * ng : nullptr;
    const double *nref = face ? fnormal + fc * 3 : nullptr;
    // fem-data coefficients fld[k] = sum_i vals_k[dof_i] psi_i(q) on the data fem (ga_instruction_val, C&E.cc:636-690)
    const double *dpE = nfields ? dphi + (size_t)fc * nq * nd_d : nullptr;
    const int *ddE = nfields ? dedof + e * nd_d : nullptr;
    for (int k = tid; k < N * ng; k += blockDim.x) {
      const int i = k / N, d = k % N;
      const int p = conn[e * ng + i];
      sG[d + N * i] = (d == 0 ? x : d == 1 ? y : z)[p];
    }
    for (int i = tid; i < nd; i += blockDim.x) sU[i] = U ? U[edof[e * nd + i]] : 0.0;
    __syncthreads();
    // per Gauss point: K = G pc(q), J = |det K|, B = K^-T; T_a(phi_i); the state; the coefficients of the two forms
    for (int q = tid; q < nq; q += blockDim.x) {
      double K[N * N], B[N * N];
      for (int k = 0; k < N * N; ++k) K[k] = 0.0;
      const double *pc = gtE + (size_t)q * ng * N;
      for (int i = 0; i < ng; ++i)
        for (int c = 0; c < N; ++c)
          for (int r = 0; r < N; ++r) K[r + N * c] += sG[r + N * i] * pc[i * N + c];
      double J;
      if (N == 2) {
        const double d = K[0] * K[3] - K[1] * K[2], id = 1.0 / d;
        B[0] = K[3] * id; B[2] = -K[1] * id; B[1] = -K[2] * id; B[3] = K[0] * id;
        J = fabs(d);
      } else {
        const double c00 = K[4] * K[8] - K[7] * K[5], c10 = K[7] * K[2] - K[1] * K[8], c20 = K[1] * K[5] - K[4] * K[2];
        const double d = K[0] * c00 + K[3] * c10 + K[6] * c20, id = 1.0 / d;
        B[0] = c00 * id; B[3] = c10 * id; B[6] = c20 * id;
        B[1] = (K[6] * K[5] - K[3] * K[8]) * id; B[4] = (K[0] * K[8] - K[6] * K[2]) * id; B[7] = (K[3] * K[2] - K[0] * K[5]) * id;
        B[2] = (K[3] * K[7] - K[6] * K[4]) * id; B[5] = (K[6] * K[1] - K[0] * K[7]) * id; B[8] = (K[0] * K[4] - K[3] * K[1]) * id;
        J = fabs(d);
      }
      vec Nq;
      for (int k = 0; k < N; ++k) Nq.v[k] = 0.0;
      if (nref) {  // unit normal = B n_ref / |B n_ref|, J *= |B n_ref|, gmm::clean(Normal, 1e-13) (C&E.cc:8836-8847)
        double nup = 0.0;
        for (int r = 0; r < N; ++r) {
          double t = 0.0;
          for (int c = 0; c < N; ++c) t += B[r + N * c] * nref[c];
          Nq.v[r] = t;
          nup += t * t;
        }
        nup = sqrt(nup);
        J *= nup;
        for (int r = 0; r < N; ++r) { const double t = Nq.v[r] / nup; Nq.v[r] = fabs(t) < 1e-13 ? 0.0 : t; }
      }
      double uq = 0.0;
      vec guq;
      for (int k = 0; k < N; ++k) guq.v[k] = 0.0;
      double *T = sT + (size_t)q * nd * NA;
      for (int i = 0; i < nd; ++i) {
        const double *g = gphiE + ((size_t)q * nd + i) * N;
        const double ph = phiE[(size_t)q * nd + i];
        T[i * NA] = ph;
        uq += sU[i] * ph;
        for (int k = 0; k < N; ++k) {
          double s = 0.0;
          for (int p = 0; p < N; ++p) s += B[k + N * p] * g[p];  // (B ghat)_k
          T[i * NA + 1 + k] = s;
          guq.v[k] += sU[i] * s;
        }
      }
      const double wq = wE[q];
      const double cw = wq == 0.0 ? 0.0 : alpha * J * wq;  // zero-weight points are skipped (C&E.cc:8852)
      vec zero, Xq;
      for (int k = 0; k < N; ++k) zero.v[k] = 0.0;
      Xq = zero;
      if (gtvE)
        for (int i = 0; i < ng; ++i)
          for (int k = 0; k < N; ++k) Xq.v[k] += sG[k + N * i] * gtvE[(size_t)q * ng + i];
      double fq[2] = {0.0, 0.0};
      vec vq = zero;  // ONE vector-valued field (a data fem of qdim = mesh dimension, e.g. an advection velocity) instead of scalars
      for (int i = 0; i < (nfields ? nd_d : 0); ++i) {
        const double ph = dpE[(size_t)q * nd_d + i];
        if (fqdim > 1) {
          for (int c = 0; c < N; ++c) vq.v[c] += dvals0[ddE[i] + c] * ph;
        } else {
          fq[0] += dvals0[ddE[i]] * ph;
          if (nfields > 1) fq[1] += dvals1[ddE[i]] * ph;
        }
      }
      if (epot) sP0[q] = cw == 0.0 ? 0.0 : cw * gf_form0(uq, guq, Xq, Nq, fq, vq, par);
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        const double tv = a == 0 ? 1.0 : 0.0;
        const vec tg = a == 0 ? zero : unit(a - 1);
        sC1[q * NA + a] = cw == 0.0 ? 0.0 : cw * gf_form1(uq, guq, Xq, Nq, fq, vq, par, tv, tg);
#pragma unroll
        for (int b = 0; b < NA; ++b) {
          const double t2v = b == 0 ? 1.0 : 0.0;
          const vec t2g = b == 0 ? zero : unit(b - 1);
          sC2[(q * NA + a) * NA + b] = cw == 0.0 ? 0.0 : cw * gf_form2(uq, guq, Xq, Nq, fq, vq, par, tv, tg, t2v, t2g);
        }
      }
    }
    __syncthreads();
    if (epot && tid == 0) {  // order 0: the element's share of the potential, Gauss points in order (ga_instruction_scalar_assembly)
      double s = 0.0;
      for (int q = 0; q < nq; ++q) s += sP0[q];
      epot[el] = s;
    }
    if (rstage)
      for (int i = tid; i < nd; i += blockDim.x) {
        double s = 0.0;
        for (int q = 0; q < nq; ++q) {
          const double *T = sT + ((size_t)q * nd + i) * NA;
          for (int a = 0; a < NA; ++a) s += sC1[q * NA + a] * T[a];
        }
        rstage[(size_t)el * nd + i] = s;
      }
    if (stage || emask) {
      double vmax = 0.0;
      for (int k = tid; k < nd * nd; k += blockDim.x) {  // k = i + nd * j (row i = Test, column j = Test2)
        const int i = k % nd, j = k / nd;
        double s = 0.0;
        for (int q = 0; q < nq; ++q) {
          const double *Ti = sT + ((size_t)q * nd + i) * NA, *Tj = sT + ((size_t)q * nd + j) * NA, *C = sC2 + (size_t)q * NA * NA;
          for (int a = 0; a < NA; ++a) {
            double t = 0.0;
            for (int b = 0; b < NA; ++b) t += C[a * NA + b] * Tj[b];
            s += Ti[a] * t;
          }
        }
        sK[k] = s;
        vmax = fmax(vmax, fabs(s));
      }
      for (int off = 16; off > 0; off >>= 1) vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, off));
      if ((tid & 31) == 0) sRed[tid >> 5] = vmax;
      __syncthreads();
      vmax = fmax(fmax(sRed[0], sRed[1]), fmax(sRed[2], sRed[3]));
      const double thr = vmax * 1e-14;  // the drop rule of add_elem_matrix (C&E.cc:4889,4898)
      for (int k = tid; k < nd * nd; k += blockDim.x) {
        const double v = sK[k];
        const bool keep = (vmax != 0.0) && (fabs(v) > thr);
        if (stage) stage[(size_t)el * nd * nd + k] = keep ? v : 0.0;
        if (emask) emask[(size_t)el * nd * nd + k] = keep ? 1 : 0;
      }
    }
    __syncthreads();
  }
}
#else  // ---------------------------------------------------------------- vector variable, qdim = mesh dimension
__device__ __forceinline__ double gf_form0(vec u, mat gu, vec X, vec Normal, const double *fld, vec vfld, const double *par) { return GF_FORM0; }
__device__ __forceinline__ double gf_form1(vec u, mat gu, vec X, vec Normal, const double *fld, vec vfld, const double *par, vec tv, mat tg) { return GF_FORM1; }
__device__ __forceinline__ double gf_form2(vec u, mat gu, vec X, vec Normal, const double *fld, vec vfld, const double *par, vec tv, mat tg, vec t2v, mat t2g) { return GF_FORM2; }

// probe slot s = c * (N+1) + a: a = 0 -> the value of component c, a = 1 + k -> d/dx_k of component c
__device__ __forceinline__ void gf_probe(int s, vec &tv, mat &tg) {
  const int c = s / (GF_N + 1), a = s % (GF_N + 1);
  for (int i = 0; i < GF_N; ++i) { tv.v[i] = 0.0; for (int j = 0; j < GF_N; ++j) tg.m[i][j] = 0.0; }
  if (a == 0) tv.v[c] = 1.0; else tg.m[c][a - 1] = 1.0;
}

extern "C" __global__ void __launch_bounds__(128)
gf_jit_elem(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z, const int *__restrict__ conn,
            const int *__restrict__ edof, const double *__restrict__ U, const double *__restrict__ w, const double *__restrict__ gt_grad,
            const double *__restrict__ phi, const double *__restrict__ gphi, const double *__restrict__ gt_val,
            const signed char *__restrict__ face, const double *__restrict__ fnormal,
            const int *__restrict__ dedof, const double *__restrict__ dphi, const double *__restrict__ dvals0,
            const double *__restrict__ dvals1, int nfields, int nd_d, int fqdim,
            const double *__restrict__ par, int ng, int nq, int nd,
            long long e0, long long ne, double alpha, double *__restrict__ stage, unsigned short *__restrict__ emask,
            double *__restrict__ rstage, double *__restrict__ epot) {
  constexpr int N = GF_N, NA = GF_N + 1, Q = GF_N, NS = Q * NA;
  extern __shared__ double sm[];
  const int s1 = nd * Q;
  double *sG = sm;                               // N x ng
  double *sU = sG + N * ng;                      // nd x Q
  double *sT = sU + s1;                          // nq x nd x NA
  constexpr int SS = Q + Q * N + 1 + 3 * N + 2;  // per point: u, Grad u, weight w J alpha, position X, unit normal, two scalar fields, one vector field
  double *sS = sT + (size_t)nq * nd * NA;        // nq x SS
  double *sC1 = sS + (size_t)nq * SS;            // nq x NS
  double *sC2 = sC1 + (size_t)nq * NS;           // nq x NS x NS
  double *sK = sC2 + (size_t)nq * NS * NS;       // s1 x s1
  double *sP0 = sK + (size_t)s1 * s1;            // nq: the order-0 integrand (w J folded in)
  __shared__ double sRed[4];
  const int tid = threadIdx.x;
  for (long long el = blockIdx.x; el < ne; el += gridDim.x) {
    const long long e = e0 + el;
    // tables of this item: the volume points, or (boundary faces, C&E.cc:8827-8848) the nq points of its face -- the face
    // tables are laid out face after face
    const int fc = face ? face[e] : 0;
    const double *wE = w + (size_t)fc * nq, *gtE = gt_grad + (size_t)fc * nq * ng * N, *phiE = phi + (size_t)fc * nq * nd;
    const double *gphiE = gphi + (size_t)fc * nq * nd * N, *gtvE = gt_val ? gt_val + (size_t)fc * nq * ng : nullptr;
    const double *nref = face ? fnormal + fc * 3 : nullptr;
    // fem-data coefficients fld[k] = sum_i vals_k[dof_i] psi_i(q) on the data fem (ga_instruction_val, C&E.cc:636-690)
    const double *dpE = nfields ? dphi + (size_t)fc * nq * nd_d : nullptr;
    const int *ddE = nfields ? dedof + e * nd_d : nullptr;
    for (int k = tid; k < N * ng; k += blockDim.x) {
      const int i = k / N, d = k % N;
      const int p = conn[e * ng + i];
      sG[d + N * i] = (d == 0 ? x : d == 1 ? y : z)[p];
    }
    for (int k = tid; k < s1; k += blockDim.x) sU[k] = U ? U[edof[e * nd + k / Q] + k % Q] : 0.0;
    __syncthreads();
    for (int q = tid; q < nq; q += blockDim.x) {
      double K[N * N], B[N * N];
      for (int k = 0; k < N * N; ++k) K[k] = 0.0;
      const double *pc = gtE + (size_t)q * ng * N;
      for (int i = 0; i < ng; ++i)
        for (int c = 0; c < N; ++c)
          for (int r = 0; r < N; ++r) K[r + N * c] += sG[r + N * i] * pc[i * N + c];
      double J;
      if (N == 2) {
        const double d = K[0] * K[3] - K[1] * K[2], id = 1.0 / d;
        B[0] = K[3] * id; B[2] = -K[1] * id; B[1] = -K[2] * id; B[3] = K[0] * id;
        J = fabs(d);
      } else {
        const double c00 = K[4] * K[8] - K[7] * K[5], c10 = K[7] * K[2] - K[1] * K[8], c20 = K[1] * K[5] - K[4] * K[2];
        const double d = K[0] * c00 + K[3] * c10 + K[6] * c20, id = 1.0 / d;
        B[0] = c00 * id; B[3] = c10 * id; B[6] = c20 * id;
        B[1] = (K[6] * K[5] - K[3] * K[8]) * id; B[4] = (K[0] * K[8] - K[6] * K[2]) * id; B[7] = (K[3] * K[2] - K[0] * K[5]) * id;
        B[2] = (K[3] * K[7] - K[6] * K[4]) * id; B[5] = (K[6] * K[1] - K[0] * K[7]) * id; B[8] = (K[0] * K[4] - K[3] * K[1]) * id;
        J = fabs(d);
      }
      vec Nq;
      for (int k = 0; k < N; ++k) Nq.v[k] = 0.0;
      if (nref) {  // unit normal = B n_ref / |B n_ref|, J *= |B n_ref|, gmm::clean(Normal, 1e-13) (C&E.cc:8836-8847)
        double nup = 0.0;
        for (int r = 0; r < N; ++r) {
          double t = 0.0;
          for (int c = 0; c < N; ++c) t += B[r + N * c] * nref[c];
          Nq.v[r] = t;
          nup += t * t;
        }
        nup = sqrt(nup);
        J *= nup;
        for (int r = 0; r < N; ++r) { const double t = Nq.v[r] / nup; Nq.v[r] = fabs(t) < 1e-13 ? 0.0 : t; }
      }
      double *S = sS + (size_t)q * SS;
      for (int k = 0; k < SS; ++k) S[k] = 0.0;
      if (gtvE)
        for (int i = 0; i < ng; ++i)
          for (int k = 0; k < N; ++k) S[Q + Q * N + 1 + k] += sG[k + N * i] * gtvE[(size_t)q * ng + i];
      for (int k = 0; k < N; ++k) S[Q + Q * N + 1 + N + k] = Nq.v[k];
      for (int i = 0; i < (nfields ? nd_d : 0); ++i) {
        const double ph = dpE[(size_t)q * nd_d + i];
        if (fqdim > 1) {
          for (int c = 0; c < N; ++c) S[Q + Q * N + 3 + 2 * N + c] += dvals0[ddE[i] + c] * ph;
        } else {
          S[Q + Q * N + 1 + 2 * N] += dvals0[ddE[i]] * ph;
          if (nfields > 1) S[Q + Q * N + 2 + 2 * N] += dvals1[ddE[i]] * ph;
        }
      }
      double *T = sT + (size_t)q * nd * NA;
      for (int i = 0; i < nd; ++i) {
        const double *g = gphiE + ((size_t)q * nd + i) * N;
        const double ph = phiE[(size_t)q * nd + i];
        T[i * NA] = ph;
        for (int c = 0; c < Q; ++c) S[c] += sU[i * Q + c] * ph;
        for (int k = 0; k < N; ++k) {
          double t = 0.0;
          for (int p = 0; p < N; ++p) t += B[k + N * p] * g[p];
          T[i * NA + 1 + k] = t;
          for (int c = 0; c < Q; ++c) S[Q + c * N + k] += sU[i * Q + c] * t;  // Grad_u(c, k)
        }
      }
      const double wq = wE[q];
      S[Q + Q * N] = wq == 0.0 ? 0.0 : alpha * J * wq;  // zero-weight points are skipped (C&E.cc:8852)
    }
    __syncthreads();
    // coefficients of the two forms: work item = (Gauss point, probe slot of Test)
    for (int it = tid; it < nq * NS; it += blockDim.x) {
      const int q = it / NS, s = it % NS;
      const double *S = sS + (size_t)q * SS;
      const double cw = S[Q + Q * N];
      vec uq, Xq, Nq; mat guq;
      for (int k = 0; k < N; ++k) { Xq.v[k] = S[Q + Q * N + 1 + k]; Nq.v[k] = S[Q + Q * N + 1 + N + k]; }
      const double *fq = S + Q + Q * N + 1 + 2 * N;
      vec vq;
      for (int k = 0; k < N; ++k) vq.v[k] = S[Q + Q * N + 3 + 2 * N + k];
      for (int c = 0; c < Q; ++c) { uq.v[c] = S[c]; for (int k = 0; k < N; ++k) guq.m[c][k] = S[Q + c * N + k]; }
      vec tv, t2v; mat tg, t2g;
      if (epot && s == 0) sP0[q] = cw == 0.0 ? 0.0 : cw * gf_form0(uq, guq, Xq, Nq, fq, vq, par);
      gf_probe(s, tv, tg);
      sC1[q * NS + s] = cw == 0.0 ? 0.0 : cw * gf_form1(uq, guq, Xq, Nq, fq, vq, par, tv, tg);
      for (int s2 = 0; s2 < NS; ++s2) {
        gf_probe(s2, t2v, t2g);
        sC2[((size_t)q * NS + s) * NS + s2] = cw == 0.0 ? 0.0 : cw * gf_form2(uq, guq, Xq, Nq, fq, vq, par, tv, tg, t2v, t2g);
      }
    }
    __syncthreads();
    if (epot && tid == 0) {
      double p0 = 0.0;
      for (int q = 0; q < nq; ++q) p0 += sP0[q];
      epot[el] = p0;
    }
    if (rstage)
      for (int k = tid; k < s1; k += blockDim.x) {
        const int i = k / Q, c = k % Q;
        double r = 0.0;
        for (int q = 0; q < nq; ++q) {
          const double *T = sT + ((size_t)q * nd + i) * NA;
          for (int a = 0; a < NA; ++a) r += sC1[q * NS + c * NA + a] * T[a];
        }
        rstage[(size_t)el * s1 + k] = r;
      }
    if (stage || emask) {
      double vmax = 0.0;
      for (int k = tid; k < s1 * s1; k += blockDim.x) {  // k = row + s1 * column, row = i*Q + c (Test), column = j*Q + d (Test2)
        const int row = k % s1, col = k / s1, i = row / Q, c = row % Q, j = col / Q, d = col % Q;
        double r = 0.0;
        for (int q = 0; q < nq; ++q) {
          const double *Ti = sT + ((size_t)q * nd + i) * NA, *Tj = sT + ((size_t)q * nd + j) * NA;
          const double *C = sC2 + ((size_t)q * NS + c * NA) * NS + d * NA;
          for (int a = 0; a < NA; ++a) {
            double t = 0.0;
            for (int b = 0; b < NA; ++b) t += C[a * NS + b] * Tj[b];
            r += Ti[a] * t;
          }
        }
        sK[k] = r;
        vmax = fmax(vmax, fabs(r));
      }
      for (int off = 16; off > 0; off >>= 1) vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, off));
      if ((tid & 31) == 0) sRed[tid >> 5] = vmax;
      __syncthreads();
      vmax = fmax(fmax(sRed[0], sRed[1]), fmax(sRed[2], sRed[3]));
      const double thr = vmax * 1e-14;  // the drop rule of add_elem_matrix (C&E.cc:4889,4898)
      if (stage)
        for (int k = tid; k < s1 * s1; k += blockDim.x) {
          const double v = sK[k];
          stage[(size_t)el * s1 * s1 + k] = ((vmax != 0.0) && (fabs(v) > thr)) ? v : 0.0;
        }
      if (emask)
        for (int p = tid; p < nd * nd; p += blockDim.x) {  // per node pair: bit (d*Q + c) = entry (row component c, column component d)
          const int j = p / nd, i = p % nd;
          unsigned mask = 0;
          for (int d = 0; d < Q; ++d)
            for (int c = 0; c < Q; ++c) {
              const double v = sK[(i * Q + c) + s1 * (j * Q + d)];
              if ((vmax != 0.0) && (fabs(v) > thr)) mask |= 1u << (d * Q + c);
            }
          emask[(size_t)el * nd * nd + p] = (unsigned short)mask;
        }
    }
    __syncthreads();
  }
}
#endif
)GFJIT";

struct JitKernel {
  CUmodule_t mod = nullptr;
  CUfunction_t fn = nullptr;
  std::string log;
  ~JitKernel() {
    if (mod) jit_api().ModuleUnload(mod);
  }
};

void jit_release(gfgpu_term *t) {
  delete static_cast<JitKernel *>(t->jit_kernel);
  t->jit_kernel = nullptr;
}

static JitKernel *jit_compile(gfgpu_term *t) {
  JitApi &api = jit_api();
  const int N = t->mesh->dim;
  std::string src = "#define GF_N " + std::to_string(N) + "\n#define GF_Q " + std::to_string(t->fem->qdim) + "\n#define GF_FORM0 (" +
                    (t->jit_form0.empty() ? std::string("0.0") : t->jit_form0) + ")\n#define GF_FORM1 (" +
                    t->jit_form1 + ")\n#define GF_FORM2 (" + t->jit_form2 + ")\n" + kJitSource;
  nvrtcProgram_t prog = nullptr;
  GF_REQUIRE(api.CreateProgram(&prog, src.c_str(), "gfgpu_jit.cu", 0, nullptr, nullptr) == 0, "nvrtcCreateProgram failed");
  const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo"};
  const int rc = api.CompileProgram(prog, 3, opts);
  size_t ls = 0;
  api.GetProgramLogSize(prog, &ls);
  std::string log(ls > 0 ? ls : 1, '\0');
  if (ls > 1) api.GetProgramLog(prog, &log[0]);
  if (rc != 0) {
    api.DestroyProgram(&prog);
    throw Error("the integrand does not compile (NVRTC):\n" + log + "\n--- form1: " + t->jit_form1 + "\n--- form2: " + t->jit_form2);
  }
  size_t cs = 0;
  GF_REQUIRE(api.GetCUBINSize(prog, &cs) == 0 && cs > 0, "nvrtcGetCUBINSize failed");
  std::vector<char> cubin(cs);
  GF_REQUIRE(api.GetCUBIN(prog, cubin.data()) == 0, "nvrtcGetCUBIN failed");
  api.DestroyProgram(&prog);
  std::unique_ptr<JitKernel> k(new JitKernel);
  k->log = log;
  GF_CUDA(cudaFree(0));  // the runtime's primary context is the driver's current context from here on
  int r = api.ModuleLoadData(&k->mod, cubin.data());
  GF_REQUIRE(r == 0, "cuModuleLoadData failed for the JIT kernel (error " + std::to_string(r) + ")");
  r = api.ModuleGetFunction(&k->fn, k->mod, "gf_jit_elem");
  GF_REQUIRE(r == 0, "cuModuleGetFunction failed for the JIT kernel");
  return k.release();
}

// compile-only check (no GPU needed): used by the CPU tests and by gfgpu_term_create_jit to fail early
std::string jit_check_source(int N, int Q, const std::string &form1, const std::string &form2) {
  JitApi &api = jit_api(false);
  std::string src = "#define GF_N " + std::to_string(N) + "\n#define GF_Q " + std::to_string(Q) + "\n#define GF_FORM0 (0.0)\n#define GF_FORM1 (" + form1 +
                    ")\n#define GF_FORM2 (" + form2 + ")\n" + kJitSource;
  nvrtcProgram_t prog = nullptr;
  if (api.CreateProgram(&prog, src.c_str(), "gfgpu_jit.cu", 0, nullptr, nullptr) != 0) return "nvrtcCreateProgram failed";
  const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17"};
  const int rc = api.CompileProgram(prog, 2, opts);
  size_t ls = 0;
  api.GetProgramLogSize(prog, &ls);
  std::string log(ls > 0 ? ls : 1, '\0');
  if (ls > 1) api.GetProgramLog(prog, &log[0]);
  api.DestroyProgram(&prog);
  return rc == 0 ? std::string() : log;
}

void launch_jit_kernel(gfgpu_term *t, const ElemArgs &a) {
  gfgpu_ctx *ctx = t->ctx;
  if (!t->jit_kernel) t->jit_kernel = jit_compile(t);
  JitKernel *k = static_cast<JitKernel *>(t->jit_kernel);
  JitApi &api = jit_api();
  const int N = t->mesh->dim, NA = N + 1, nd = t->fem->nd, nq = a.face ? a.nqf : a.nq, ng = a.ng;
  const int Q = t->fem->qdim;
  GF_REQUIRE(Q == 1 || Q == N, "JIT terms: scalar variables, or vector variables of the mesh dimension");
  const size_t NS = (size_t)Q * NA, s1 = (size_t)nd * Q;
  const size_t smem = Q == 1 ? ((size_t)N * ng + nd + (size_t)nq * nd * NA + (size_t)nq * NA + (size_t)nq * NA * NA + (size_t)nd * nd + nq + 2) * 8
                             : ((size_t)N * ng + s1 + (size_t)nq * nd * NA + (size_t)nq * (Q + Q * N + 1 + 3 * N + 2) + (size_t)nq * NS +
                                (size_t)nq * NS * NS + s1 * s1 + nq + 2) * 8;
  GF_REQUIRE(smem <= 220 * 1024, "JIT terms: element too large for the run-time kernel (nq x nd x (N+1) doubles of shared memory)");
  if (t->jit_par.n != (size_t)GFGPU_MAX_PARAMS) {
    t->jit_par.alloc(ctx, GFGPU_MAX_PARAMS);
  }
  t->jit_par.upload(t->par);
  GF_REQUIRE(api.FuncSetAttribute(k->fn, 8 /* CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES */, (int)smem) == 0,
             "cuFuncSetAttribute failed for the JIT kernel");
  long long e0 = a.e0, ne = a.e1 - a.e0;
  if (ne <= 0) return;
  int ing = ng, inq = nq, ind = nd;
  double alpha = a.alpha;
  const double *par = t->jit_par.p;
  const double *x = a.x, *y = a.y, *z = a.z, *U = a.U, *w = a.w, *gt = a.gt_grad, *phi = a.phi, *gphi = a.gphi;
  const double *gtv = t->tab->gt_val.n ? t->tab->gt_val.p : nullptr;
  const signed char *face = reinterpret_cast<const signed char *>(a.face);
  const double *fnormal = a.fnormal;
  const int32_t *dedof = a.nfields ? a.dedof : nullptr;
  const double *dphi = a.nfields ? a.dphi : nullptr, *dv0 = a.dvals0, *dv1 = a.dvals1;
  int nfields = a.nfields, nd_d = a.nd_d, fqdim = (a.nfields && t->dfem) ? t->dfem->qdim : 1;
  if (face) {  // boundary faces: the kernel indexes the face tables by the item's face
    w = a.fw; gt = a.fgt_grad; phi = a.fphi; gphi = a.fgphi;
    gtv = t->tab->fgt_val.n ? t->tab->fgt_val.p : nullptr;
    if (a.nfields) {
      GF_REQUIRE(a.dfphi, "fem-data coefficients on a region of faces need the data fem's basis at the face points");
      dphi = a.dfphi;
    }
  }
  auto mentions_X = [](const std::string &f) {
    for (size_t k = 0; k < f.size(); ++k)
      if (f[k] == 'X' && (k == 0 || !(isalnum((unsigned char)f[k - 1]) || f[k - 1] == '_')) &&
          (k + 1 == f.size() || !(isalnum((unsigned char)f[k + 1]) || f[k + 1] == '_')))
        return true;
    return false;
  };
  auto mentions = [](const std::string &f, const std::string &id) {
    for (size_t k = f.find(id); k != std::string::npos; k = f.find(id, k + 1))
      if ((k == 0 || !(isalnum((unsigned char)f[k - 1]) || f[k - 1] == '_')) &&
          (k + id.size() == f.size() || !(isalnum((unsigned char)f[k + id.size()]) || f[k + id.size()] == '_')))
        return true;
    return false;
  };
  GF_REQUIRE(gtv || !(mentions_X(t->jit_form1) || mentions_X(t->jit_form2)),
             "the integrand mentions the position X: give the tables the geometric transformation's values (gfgpu_tables_set_gt_values)");
  GF_REQUIRE(face || !(mentions(t->jit_form1, "Normal") || mentions(t->jit_form2, "Normal")),
             "the integrand mentions the unit normal: the term must be restricted to a region of faces (gfgpu_term_set_region)");
  GF_REQUIRE(nfields || !(mentions(t->jit_form1, "fld") || mentions(t->jit_form2, "fld")),
             "the integrand mentions a fem-data coefficient fld[k]: give the term its fields (gfgpu_term_set_fields)");
  const int32_t *conn = a.conn, *edof = a.edof;
  double *stage = a.stage, *rstage = a.rstage;
  uint16_t *emask = a.emask;
  double *epot = t->jit_epot;
  void *params[] = {&x, &y, &z, &conn, &edof, &U, &w, &gt, &phi, &gphi, &gtv, &face, &fnormal, &dedof, &dphi, &dv0, &dv1, &nfields, &nd_d, &fqdim, &par, &ing, &inq, &ind, &e0, &ne, &alpha, &stage, &emask, &rstage, &epot};
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(ne, (long long)ctx->sm_count * 4));
  const int r = api.LaunchKernel(k->fn, grid, 1, 1, 128, 1, 1, (unsigned)smem, ctx->stream, params, nullptr);
  GF_REQUIRE(r == 0, "cuLaunchKernel failed for the JIT kernel (error " + std::to_string(r) + ")");
  count_launch(1);
}

}  // namespace gf
