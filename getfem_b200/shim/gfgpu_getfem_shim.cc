// gfgpu_getfem_shim.cc -- see gfgpu_getfem_shim.h.
#include "gfgpu_getfem_shim.h"
#ifdef _OPENMP
#include <omp.h>
#endif

#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <regex>
#include <set>
#include <sstream>
#include <tuple>

#include "getfem/getfem_fem.h"
#include "getfem/getfem_generic_assembly_tree.h"
#include "getfem/getfem_im_data.h"
#include "getfem/getfem_integration.h"
#include "getfem/getfem_mesh_fem.h"
#include "getfem/getfem_mesh_im.h"
#include "getfem/getfem_omp.h"

namespace getfem_b200 {

using getfem::size_type;

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

#define GFGPU_CALL(call) GMM_ASSERT1((call) == 0, "gfgpu: " << gfgpu_last_error())

static std::string strip(const std::string &s) {
  std::string r;
  for (char c : s)
    if (c != ' ' && c != '\t' && c != '\n') r.push_back(c);
  return r;
}

static bool recognise_string(const getfem::ga_workspace &ws, const std::string &v, const std::string &s, recognised_term &out);
static bool recognise_coupled(const getfem::ga_workspace &ws, const std::string &s0, int order, const std::string &test1,
                              const std::string &test2, recognised_term &out);

bool recognise_tree(const getfem::ga_workspace &ws, size_type itree, recognised_term &out) {
  const getfem::ga_workspace::tree_description &td = ws.tree_info(itree);
  if (td.order != 1 || td.operation != getfem::ga_workspace::ASSEMBLY) return false;
  return recognise_string(ws, td.name_test1, strip(getfem::ga_tree_to_string(*td.ptree)), out);
}

// removes parentheses that enclose the whole string
static std::string strip_outer(std::string s) {
  for (;;) {
    if (s.size() < 2 || s.front() != '(' || s.back() != ')') return s;
    int depth = 0;
    bool encloses = true;
    for (size_t i = 0; i < s.size(); ++i) {
      if (s[i] == '(') ++depth;
      else if (s[i] == ')') --depth;
      if (depth == 0 && i + 1 < s.size()) { encloses = false; break; }
    }
    if (!encloses) return s;
    s = s.substr(1, s.size() - 2);
  }
}

static bool recognise_sum(const getfem::ga_workspace &ws, const std::string &v, const std::string &s0,
                          std::vector<recognised_term> &out) {
  const std::string s = strip_outer(s0);
  recognised_term rt;
  if (recognise_string(ws, v, s, rt) || recognise_string(ws, v, s0, rt)) { out.push_back(rt); return true; }
  if (recognise_coupled(ws, s, 1, v, "", rt) || recognise_coupled(ws, s0, 1, v, "", rt)) { out.push_back(rt); return true; }
  int depth = 0;  // LAST top-level '+' or binary '-': add_tree builds ((A)+(B))+(C), a written "A - B" prints as (A)-(B)
  size_t at = std::string::npos;
  for (size_t i = 0; i < s.size(); ++i) {
    if (s[i] == '(' || s[i] == '[') ++depth;
    else if (s[i] == ')' || s[i] == ']') --depth;
    else if (s[i] == '+' && depth == 0 && i > 0) at = i;
    else if (s[i] == '-' && depth == 0 && i > 0 && (s[i - 1] == ')' || s[i - 1] == ']' || std::isalnum((unsigned char)s[i - 1]) || s[i - 1] == '_'))
      at = i;
  }
  if (at == std::string::npos) return false;
  if (s[at] == '-')  // X - Y = X + (-(Y)): the right part is recognised in its negated spelling
    return recognise_sum(ws, v, s.substr(0, at), out) && recognise_sum(ws, v, "-(" + strip_outer(s.substr(at + 1)) + ")", out);
  return recognise_sum(ws, v, s.substr(0, at), out) && recognise_sum(ws, v, s.substr(at + 1), out);
}

// number of top-level summands of a printed tree: "(A)+(B)" -> 2
static size_t count_top_level_summands(const std::string &s0) {
  const std::string s = strip_outer(s0);
  int depth = 0;
  size_t n = 1;
  for (size_t i = 0; i < s.size(); ++i) {
    if (s[i] == '(' || s[i] == '[') ++depth;
    else if (s[i] == ')' || s[i] == ']') --depth;
    else if (s[i] == '+' && depth == 0 && i > 0) ++n;
    else if (s[i] == '-' && depth == 0 && i > 0 && (s[i - 1] == ')' || s[i - 1] == ']' || std::isalnum((unsigned char)s[i - 1]) || s[i - 1] == '_'))
      ++n;  // binary minus
  }
  return n;
}

// ORDER-2 trees written directly with Test_ / Test2_ (no order-1 tree): the legacy asm_* wrappers of getfem_assembling.h --
// "Test_u:Test2_u" (asm_mass_matrix :695-720), "Grad_Test_u:Grad_Test2_u" (asm_stiffness_matrix_for_homogeneous_laplacian
// :1108-1131), "((lambda*Div_Test_u)*Id(meshdim)+(2*mu)*Sym(Grad_Test_u)):Grad_Test2_u" (asm_stiffness_matrix_for_*linear_
// elasticity :972-1048).  The handled forms are symmetric: Test_ and Test2_ may come in either order.
static bool recognise_order2_string(const getfem::ga_workspace &ws, const std::string &v, const std::string &s0,
                                    recognised_term &out) {
  const std::string s = strip_outer(s0);
  const std::string ID = "([A-Za-z_][A-Za-z_0-9]*)", T = "Test2?_" + v;
  const std::string I3 = "\\[\\[1,0,0\\],\\[0,1,0\\],\\[0,0,1\\]\\]", I2 = "\\[\\[1,0\\],\\[0,1\\]\\]";
  const std::string Idm = "(?:" + I3 + "|" + I2 + ")";
  std::smatch m;
  out.varname = v;
  out.params.clear();
  out.field_names.clear();
  out.field_sign = 1.0;
  auto coef = [&](const std::string &name) {
    GMM_ASSERT1(ws.is_constant(name), "gfgpu: '" << name << "' must be a constant");
    if (ws.associated_mf(name) || ws.associated_im_data(name)) { out.field_names.push_back(name); return 1.0; }  // fem data / im data: a field
    GMM_ASSERT1(ws.value(name).size() == 1, "gfgpu: '" << name << "' must be a scalar constant");
    return ws.value(name)[0];
  };
  auto both_tests = [&](const std::string &str) {  // exactly one Test_v and one Test2_v
    return str.find("Test_" + v) != std::string::npos && str.find("Test2_" + v) != std::string::npos;
  };
  if (!both_tests(s)) return false;
  if (std::regex_match(s, std::regex(T + "[.:]" + T))) { out.family = GFGPU_MASS; out.params = {1.0}; return true; }
  if (std::regex_match(s, m, std::regex("\\(" + ID + "\\*" + T + "\\)[.:]" + T))) {
    out.family = GFGPU_MASS; out.params = {coef(m[1])}; return true;
  }
  if (std::regex_match(s, std::regex("Grad_" + T + "[.:]Grad_" + T))) { out.family = GFGPU_LAPLACE; out.params = {1.0}; return true; }
  if (std::regex_match(s, m, std::regex("\\(" + ID + "\\*Grad_" + T + "\\)[.:]Grad_" + T))) {
    out.family = GFGPU_LAPLACE; out.params = {coef(m[1])}; return true;
  }
  if (std::regex_match(s, m, std::regex("\\(\\(\\(" + ID + "\\*Div_" + T + "\\)\\*" + Idm + "\\)\\+\\(\\(2\\*" + ID +
                                        "\\)\\*\\(Sym\\(Grad_" + T + "\\)\\)\\)\\):Grad_" + T))) {
    GMM_ASSERT1((ws.associated_mf(m[1]) || ws.associated_im_data(m[1])) || !(ws.associated_mf(m[2]) || ws.associated_im_data(m[2])),
                "gfgpu: a fem-data mu needs a fem-data lambda (fields replace the LEADING parameters)");
    out.family = GFGPU_ELASTICITY; out.params = {coef(m[1]), coef(m[2])}; return true;
  }
  return false;
}

// ---------------------------------------------------------------- recognition by probe
// The reference's own tests write ONE bilinear form in many algebraically equivalent ways (tests/test_assembly.cc:777-866:
// "lambda*Div_Test_u*Div_Test2_u + mu*(Grad_Test_u'+Grad_Test_u):Grad_Test2_u", "lambda*Trace(..)*Trace(..) + ...", ...); their
// analysed trees all print differently.  When no printed normal form matches a directly written order-2 tree, the form is
// identified NUMERICALLY: the tree's expression and the normal forms of the families are assembled by the reference's own
// interpreter on the first TWO convexes of the region (a symbolic-phase cost of microseconds, not a fallback: every other element
// of the region is assembled on the device), and the element matrices are fitted:  K = a M (mass), a L (Laplace),
// lambda D + mu S (elasticity).  A fit to 1e-11 of |K| is a recognition; anything else (non-constant coefficient, another
// form, a nonlinear tangent) is not.  The probe runs at every assembly, so constants may change between calls.
static void (*g_reference_assembly)(getfem::ga_workspace &, size_type) = nullptr;
void set_reference_assembly(void (*f)(getfem::ga_workspace &, size_type)) { g_reference_assembly = f; }

static bool probe_matrix(const getfem::ga_workspace &ws, const std::string &v, const getfem::mesh_fem &mf, const getfem::mesh_im &mim,
                         const getfem::mesh_region &rg, const std::string &expr, std::vector<double> &dense,
                         const std::vector<size_type> &dofs, const getfem::model_real_plain_vector *private_state = nullptr) {
  try {
    getfem::ga_workspace w2(ws, getfem::ga_workspace::inherit::ALL);  // sees the variables and constants of the caller
    // a PRIVATE state of the unknown shadows the caller's (own variables are looked up first, workspace.cc:328-339)
    if (private_state) w2.add_fem_variable(v, mf, ws.interval_of_variable(v), *private_state);
    w2.add_expression(expr, mim, rg, 2);
    getfem::model_real_sparse_matrix K(ws.nb_primary_dof() ? ws.nb_primary_dof() : mf.nb_dof(),
                                       ws.nb_primary_dof() ? ws.nb_primary_dof() : mf.nb_dof());
    w2.set_assembled_matrix(K);
    if (g_reference_assembly) g_reference_assembly(w2, 2); else w2.assembly(2);
    const size_type off = ws.interval_of_variable(v).first();
    dense.assign(dofs.size() * dofs.size(), 0.0);
    for (size_t c = 0; c < dofs.size(); ++c)
      for (size_t r = 0; r < dofs.size(); ++r) dense[r + dofs.size() * c] = K(off + dofs[r], off + dofs[c]);
    return true;
  } catch (const std::exception &) { return false; }
}

// a fitted coefficient carries the round-off of the fit (1e-14 relative): 12 significant digits, and exact zeros
static double snap_coefficient(double x, double scale) {
  if (std::fabs(x) <= 1e-12 * scale) return 0.0;
  char b[40];
  std::snprintf(b, sizeof b, "%.12g", x);
  return std::strtod(b, nullptr);
}

// Two probe items say nothing about the REST of the region unless the form is the same everywhere: the probe is only allowed on
// expressions whose leaves are the variable itself and FIXED-SIZE constants.  Fem data (a piecewise material!), im_data, other
// variables, the position X, the normal, element sizes, interpolate / elementary transformations ... are left to the printed
// normal forms (which handle fem-data coefficients and normal loads explicitly) or refused.
// `unknown_allowed` = false also refuses the variable itself WITHOUT a Test_ / Test2_ prefix: a bilinear form or a load that
// contains the unknown depends on the state, and a numerical fit at ONE state (zero on the probe convexes, say, at the first
// Newton step) would take a nonlinear form such as (1+u*u)*Grad_u.Grad_Test_u for a Laplacian.  Only the order-1 tree of a
// derived pair may contain the unknown: its derivative -- checked with unknown_allowed = false -- is then state independent, so
// the tree is affine in u, and the r = K u check at the current state leaves no room for a constant part.
static bool probe_is_safe(const getfem::ga_workspace &ws, const std::string &v, const std::string &printed,
                          bool unknown_allowed) {
  static const char *forbidden[] = {"X", "Normal", "element_size", "element_K", "element_B", "Interpolate", "Interpolate_filter",
                                    "Interpolate_derivative", "Elementary", "Elementary_transformation", "Secondary_domain",
                                    "Secondary_Domain", "Xfem_plus", "Xfem_minus", "Cross_product", "Print"};
  static const char *prefixes[] = {"Test2_", "Test_", "Grad_", "Hess_", "Div_", "Previous_", "Old_"};
  static const std::regex ident("[A-Za-z_][A-Za-z_0-9]*");
  for (auto it = std::sregex_iterator(printed.begin(), printed.end(), ident); it != std::sregex_iterator(); ++it) {
    std::string tok = it->str();
    for (const char *f : forbidden)
      if (tok == f) return false;
    bool is_test = false;
    for (bool again = true; again;) {
      again = false;
      for (const char *p : prefixes) {
        const size_t n = std::strlen(p);
        if (tok.size() > n && tok.compare(0, n, p) == 0) {
          is_test = is_test || p[0] == 'T';
          tok = tok.substr(n);
          again = true;
        }
      }
    }
    if (tok == v) {
      if (!is_test && !unknown_allowed) return false;
      continue;
    }
    if (ws.variable_group_exists(tok)) return false;
    if (!ws.variable_exists(tok)) continue;  // a function or operator name
    if (!ws.is_constant(tok) || ws.associated_mf(tok) || ws.associated_im_data(tok)) return false;
  }
  return true;
}

// the order-1 tree `expr` (a linear form in Test_v at the current state) assembled on the probe region -> values on `dofs`
static bool probe_vector(const getfem::ga_workspace &ws, const std::string &v, const getfem::mesh_fem &mf, const getfem::mesh_im &mim,
                         const getfem::mesh_region &rg, const std::string &expr, std::vector<double> &vec,
                         const std::vector<size_type> &dofs) {
  try {
    getfem::ga_workspace w2(ws, getfem::ga_workspace::inherit::ALL);
    w2.add_expression(expr, mim, rg, 1);
    getfem::base_vector V(ws.nb_primary_dof() ? ws.nb_primary_dof() : mf.nb_dof());
    w2.set_assembled_vector(V);
    if (g_reference_assembly) g_reference_assembly(w2, 1); else w2.assembly(1);
    const size_type off = ws.interval_of_variable(v).first();
    vec.resize(dofs.size());
    for (size_t r = 0; r < dofs.size(); ++r) vec[r] = V[off + dofs[r]];
    return true;
  } catch (const std::exception &) { return false; }
}

// An order-1 tree without any derivative tree does not depend on the unknowns: a load.  Whatever its spelling, a CONSTANT load is
// r = sum_b F_b int Test_v(b): F is fitted on the first two items (convexes or faces) of the region and the tree becomes the
// SOURCE family; a load that varies in space, or with the normal of non-coplanar faces, does not fit.
static bool recognise_load_by_probe(const getfem::ga_workspace &ws, size_type itree, recognised_term &out) {
  const getfem::ga_workspace::tree_description &td = ws.tree_info(itree);
  const std::string v = td.name_test1;
  for (size_type j = 0; j < ws.nb_trees(); ++j) {  // a coupling tangent (Test_v, Test2_w) means the tree depends on w: not a load
    const auto &t2 = ws.tree_info(j);
    if (t2.order == 2 && t2.mim == td.mim && t2.rg == td.rg && (t2.name_test1 == v || t2.name_test2 == v)) return false;
  }
  const getfem::mesh_fem *pmf = ws.associated_mf(v);
  if (!pmf || pmf->is_reduced() || !td.mim || !td.rg) return false;
  if (!probe_is_safe(ws, v, getfem::ga_tree_to_string(*td.ptree), false)) return false;
  const getfem::mesh &m = pmf->linked_mesh();
  const size_type Q = pmf->get_qdim();
  getfem::mesh_region rg2;
  std::vector<size_type> dofs;
  size_type nit = 0;
  for (getfem::mr_visitor it(*td.rg, m); !it.finished() && nit < 2; ++it, ++nit) {
    if (it.f() != getfem::short_type(-1)) rg2.add(it.cv(), it.f()); else rg2.add(it.cv());
    for (size_type d : pmf->ind_basic_dof_of_element(it.cv())) dofs.push_back(d);
  }
  if (!nit) return false;
  std::sort(dofs.begin(), dofs.end());
  dofs.erase(std::unique(dofs.begin(), dofs.end()), dofs.end());
  std::vector<double> r, t, fit(dofs.size(), 0.0), F(Q, 0.0);
  if (!probe_vector(ws, v, *pmf, *td.mim, rg2, getfem::ga_tree_to_string(*td.ptree), r, dofs)) return false;
  double nr = 0;
  for (double x : r) nr += x * x;
  if (nr == 0) return false;
  for (size_type b = 0; b < Q; ++b) {  // the component templates have disjoint supports: independent one-parameter fits
    const std::string e = Q == 1 ? "Test_" + v : "Test_" + v + "(" + std::to_string(b + 1) + ")";
    if (!probe_vector(ws, v, *pmf, *td.mim, rg2, e, t, dofs)) return false;
    double tt = 0, tr = 0;
    for (size_t k = 0; k < t.size(); ++k) { tt += t[k] * t[k]; tr += t[k] * r[k]; }
    if (tt == 0) return false;
    F[b] = tr / tt;
    for (size_t k = 0; k < t.size(); ++k) fit[k] += F[b] * t[k];
  }
  double dr = 0, sc = 0;
  for (size_t k = 0; k < r.size(); ++k) dr += (r[k] - fit[k]) * (r[k] - fit[k]);
  if (dr > 1e-22 * nr) return false;
  for (double f : F) sc = std::max(sc, std::fabs(f));
  out.varname = v;
  out.field_names.clear();
  out.field_sign = 1.0;
  out.by_probe = true;
  out.family = GFGPU_SOURCE;
  out.params.clear();
  for (double f : F) out.params.push_back(snap_coefficient(f, sc));
  return true;
}

static bool recognise_by_probe(const getfem::ga_workspace &ws, size_type itree, recognised_term &out) {
  const getfem::ga_workspace::tree_description &td0 = ws.tree_info(itree);
  // an order-1 tree is identified through its derivative (the order-2 tree of the same mim / region / variable, added by
  // add_expression with the default derivative order): the bilinear form K is probed there, and the order-1 tree itself must
  // then be the linear form K u at the current state, nothing else (no load mixed in)
  size_type i2 = itree;
  if (td0.order == 1) {
    i2 = size_type(-1);
    for (size_type j = 0; j < ws.nb_trees(); ++j) {
      const auto &t2 = ws.tree_info(j);
      if (t2.order == 2 && t2.mim == td0.mim && t2.rg == td0.rg && t2.name_test1 == td0.name_test1 &&
          t2.name_test2 == td0.name_test1)
        i2 = j;
    }
    if (i2 == size_type(-1)) return recognise_load_by_probe(ws, itree, out);
  }
  const getfem::ga_workspace::tree_description &td = ws.tree_info(i2);
  const std::string v = td.name_test1;
  const getfem::mesh_fem *pmf = ws.associated_mf(v);
  if (!pmf || pmf->is_reduced() || !td.mim || !td.rg) return false;
  // the tangent tree may mention the unknown and still be state independent (sqr(Norm(Grad_u))/2 differentiates into
  // Norm(Grad_u)*Derivative_1_Norm(..)): such a tree is then probed at TWO private random states, never at the caller's
  const bool k_mentions_u = !probe_is_safe(ws, v, getfem::ga_tree_to_string(*td.ptree), false);
  if (!probe_is_safe(ws, v, getfem::ga_tree_to_string(*td.ptree), true) ||
      !probe_is_safe(ws, v, getfem::ga_tree_to_string(*td0.ptree), true))
    return false;
  if (k_mentions_u && td0.order != 1) {
    // asked about the tangent tree itself: fine when it is the derivative of an order-1 tree (same two-state check below);
    // a DIRECTLY written "bilinear" form that mentions the unknown has a state-dependent coefficient
    bool derived = false;
    for (size_type j = 0; j < ws.nb_trees() && !derived; ++j) {
      const auto &t1 = ws.tree_info(j);
      derived = t1.order == 1 && t1.mim == td.mim && t1.rg == td.rg && t1.name_test1 == v;
    }
    if (!derived) return false;
  }
  const getfem::mesh &m = pmf->linked_mesh();
  const size_type Q = pmf->get_qdim(), N = m.dim();
  getfem::mesh_region rg2;
  std::vector<size_type> dofs;
  size_type ncv = 0;
  for (getfem::mr_visitor it(*td.rg, m); !it.finished() && ncv < 2; ++it) {
    if (it.f() != getfem::short_type(-1)) return false;  // face regions: printed forms only
    rg2.add(it.cv());
    for (size_type d : pmf->ind_basic_dof_of_element(it.cv())) dofs.push_back(d);
    ++ncv;
  }
  if (!ncv) return false;
  std::sort(dofs.begin(), dofs.end());
  dofs.erase(std::unique(dofs.begin(), dofs.end()), dofs.end());
  std::vector<double> K, A, B;
  if (k_mentions_u) {
    getfem::model_real_plain_vector Ua(pmf->nb_dof()), Ub(pmf->nb_dof());
    uint64_t sd = 0x9E3779B97F4A7C15ull;
    auto rnd = [&sd]() {  // splitmix64 -> (-1, 1)
      sd += 0x9E3779B97F4A7C15ull;
      uint64_t z = sd;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z ^= z >> 31;
      return double(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
    };
    for (double &x : Ua) x = rnd();
    for (double &x : Ub) x = 3.0 * rnd();
    std::vector<double> Kb;
    if (!probe_matrix(ws, v, *pmf, *td.mim, rg2, getfem::ga_tree_to_string(*td.ptree), K, dofs, &Ua) ||
        !probe_matrix(ws, v, *pmf, *td.mim, rg2, getfem::ga_tree_to_string(*td.ptree), Kb, dofs, &Ub))
      return false;
    double na = 0, dab = 0;
    for (size_t k = 0; k < K.size(); ++k) { na += K[k] * K[k]; dab += (K[k] - Kb[k]) * (K[k] - Kb[k]); }
    if (!(dab <= 1e-22 * na) || !std::isfinite(na) || na > 1e150) return false;  // the tangent moves with the state (or is not finite): a nonlinear form
  } else if (!probe_matrix(ws, v, *pmf, *td.mim, rg2, getfem::ga_tree_to_string(*td.ptree), K, dofs)) {
    return false;
  }
  double nK = 0;
  for (double x : K) nK += x * x;
  if (!(nK > 0) || !std::isfinite(nK) || nK > 1e150) return false;  // (a law's 1e200 penalty on an inverted probe state is no fit)
  if (td0.order == 1) {
    std::vector<double> r;
    if (!probe_vector(ws, v, *pmf, *td.mim, rg2, getfem::ga_tree_to_string(*td0.ptree), r, dofs)) return false;
    const getfem::model_real_plain_vector &Uv = ws.value(v);
    double nr = 0, dr = 0, nu = 0;
    for (size_t a = 0; a < dofs.size(); ++a) {
      double ku = 0;
      for (size_t b = 0; b < dofs.size(); ++b) ku += K[a + dofs.size() * b] * Uv[dofs[b]];
      nr += r[a] * r[a];
      dr += (r[a] - ku) * (r[a] - ku);
      nu += Uv[dofs[a]] * Uv[dofs[a]];
    }
    if (dr > 1e-22 * std::max(nr, nK * nu)) return false;
  }
  const std::string T1 = "Test_" + v, T2 = "Test2_" + v;
  auto fit1 = [&](const std::vector<double> &X, double &a) {  // K = a X ?
    double xx = 0, xk = 0;
    for (size_t k = 0; k < K.size(); ++k) { xx += X[k] * X[k]; xk += X[k] * K[k]; }
    if (xx == 0) return false;
    a = xk / xx;
    double r = 0;
    for (size_t k = 0; k < K.size(); ++k) r += (K[k] - a * X[k]) * (K[k] - a * X[k]);
    return r <= 1e-22 * nK;
  };
  out.varname = v;
  out.field_names.clear();
  out.field_sign = 1.0;
  out.by_probe = true;
  double a = 0;
  if (probe_matrix(ws, v, *pmf, *td.mim, rg2, Q == 1 ? T1 + "*" + T2 : T1 + "." + T2, A, dofs) && fit1(A, a)) {
    out.family = GFGPU_MASS; out.params = {snap_coefficient(a, std::fabs(a))}; return true;
  }
  if (probe_matrix(ws, v, *pmf, *td.mim, rg2, Q == 1 ? "Grad_" + T1 + ".Grad_" + T2 : "Grad_" + T1 + ":Grad_" + T2, A, dofs) && fit1(A, a)) {
    out.family = GFGPU_LAPLACE; out.params = {snap_coefficient(a, std::fabs(a))}; return true;
  }
  if (Q == N && Q > 1 && probe_matrix(ws, v, *pmf, *td.mim, rg2, "Div_" + T1 + "*Div_" + T2, A, dofs) &&
      probe_matrix(ws, v, *pmf, *td.mim, rg2, "(Grad_" + T1 + "'+Grad_" + T1 + "):Grad_" + T2, B, dofs)) {
    double aa = 0, ab = 0, bb = 0, ak = 0, bk = 0;  // normal equations of K = lambda A + mu B
    for (size_t k = 0; k < K.size(); ++k) { aa += A[k] * A[k]; ab += A[k] * B[k]; bb += B[k] * B[k]; ak += A[k] * K[k]; bk += B[k] * K[k]; }
    const double det = aa * bb - ab * ab;
    if (det > 1e-12 * aa * bb) {
      const double lam = (ak * bb - bk * ab) / det, mu = (bk * aa - ak * ab) / det;
      double r = 0;
      for (size_t k = 0; k < K.size(); ++k) { const double d = K[k] - lam * A[k] - mu * B[k]; r += d * d; }
      if (r <= 1e-22 * nK) {
        const double sc = std::max(std::fabs(lam), std::fabs(mu));
        out.family = GFGPU_ELASTICITY;
        out.params = {snap_coefficient(lam, sc), snap_coefficient(mu, sc)};
        return true;
      }
    }
  }
  return false;
}

// ---------------------------------------------------------------- the NVRTC route: tree -> C expression
// For a SCALAR fem variable v an analysed tree is, at a Gauss point, an expression in u = v, gu = Grad_v, scalar constants
// and the test functions; Test_v / Grad_Test_v (and Test2) become the probe arguments tv / tg (t2v / t2g) of
// gfgpu_term_create_jit.  Every value is a scalar (rank 0) or a vector of the mesh dimension (rank 1); anything else --
// matrices, X, Normal, other variables, fem or im data, interpolate transformations -- is refused, and the caller then
// reports the expression as not handled.  (ga_exec evaluates the same tree with tensors whose leading dimensions index
// the test functions, C&E.cc:2769-3760; for a scalar variable the contraction pattern reduces to these ranks.)
struct jit_value { std::string code; int rank; };  // rank 0 scalar, 1 vector of the mesh dimension, 2 matrix N x N

static bool jit_emit(const getfem::ga_workspace &ws, const getfem::pga_tree_node &n, const std::string &v, int N, int Q,
                     std::vector<std::string> &params, std::vector<std::string> &fields, jit_value &out) {
  using namespace getfem;
  if (!n) return false;
  auto num = [](double x) { char b[48]; std::snprintf(b, sizeof b, "(%.17g)", x); std::string r(b);
                            if (r.find_first_of(".eEn") == std::string::npos) r.insert(r.size() - 1, ".0"); return r; };
  auto child = [&](size_t k, jit_value &o) { return k < n->children.size() && jit_emit(ws, n->children[k], v, N, Q, params, fields, o); };
  const int rv = Q == 1 ? 0 : 1;  // rank of the variable's value; its gradient has one more
  auto tensor_const = [&](const base_tensor &t, jit_value &o) {
    if (t.size() == 1) { o = {num(t[0]), 0}; return true; }
    if (t.sizes().size() == 1 && t.size() == size_t(N)) {
      o = {"mkvec(" + num(t[0]) + "," + num(t[1]) + "," + (N > 2 ? num(t[2]) : std::string("0.0")) + ")", 1};
      return true;
    }
    if (t.sizes().size() == 2 && t.sizes()[0] == size_type(N) && t.sizes()[1] == size_type(N)) {  // first index fastest
      std::string c = "mkmat(";
      for (int k = 0; k < 9; ++k) c += (k ? "," : "") + (k < N * N ? num(t[k]) : std::string("0.0"));
      o = {c + ")", 2};
      return true;
    }
    return false;
  };
  // "Saint_Venant_Kirchhoff_PK2(G, params)" (AHL_wrapper_sigma with the SVK law, getfem_nonlinear_elasticity.cc:503-540, 1781-1827:
  // E = (G + G' + G'G)/2, S = lambda tr(E) I + 2 mu E -- the one law the reference defines in ANY dimension, so this is also its
  // 2D finite strain) and its first derivative contracted with a direction H: args = "G,lambda,mu"
  // The same for the compressible neo-Hookean laws (Neo_Hookean_hyperelastic_law, :612-702: S = mu I + k(det C) C^-1,
  // k = lambda/2 (i3 - 1) - mu (Ciarlet) or lambda/2 log(i3) - mu (Bonet), + 1e200 C where det F <= 0), which the reference
  // defines on 3 x 3 tensors, and for the PLANE STRAIN wrappers of the three laws (plane_strain_hyperelastic_law, :906-945: the
  // 2 x 2 strain embedded in a 3 x 3 one with zeros; with C33 = 1 the 2 x 2 block of S is the SAME formula on 2 x 2 tensors).
  // `fn` receives the device helper: svk_pk2 / nh_pk2 (+ a bonet flag folded into `args`).
  auto svk_args = [&](const pga_tree_node &pn, size_type der1, std::string &args, std::string *fn = nullptr) {
    if (pn->node_type != GA_NODE_PARAMS || pn->children.size() != 3) return false;
    const pga_tree_node &f = pn->children[0];
    if (f->node_type != GA_NODE_OPERATOR || f->der1 != der1 || f->der2 != 0) return false;
    std::string lawfn, tail, head;
    size_t np = 2;  // parameters of the law
    const std::string ps = N == 2 ? "Plane_Strain_" : "";
    if (f->name == "Saint_Venant_Kirchhoff_PK2" || (N == 2 && f->name == "Plane_Strain_Saint_Venant_Kirchhoff_PK2")) lawfn = "svk";
    else if (f->name == ps + "Compressible_Neo_Hookean_Ciarlet_PK2") { lawfn = "nh"; tail = ",0"; }
    else if (f->name == ps + "Compressible_Neo_Hookean_Bonet_PK2") { lawfn = "nh"; tail = ",1"; }
    // the laws given through the invariants of C (iso_pk2 / iso_dpk2: law id first, five parameter slots)
    else if (f->name == ps + "Compressible_Mooney_Rivlin_PK2") { lawfn = "iso"; head = "0,"; np = 3; }
    else if (f->name == ps + "Ciarlet_Geymonat_PK2") { lawfn = "iso"; head = "1,"; np = 3; }
    else if (f->name == ps + "Generalized_Blatz_Ko_PK2") { lawfn = "iso"; head = "2,"; np = 5; }
    else return false;
    if (fn) *fn = lawfn;
    jit_value g;
    if (!jit_emit(ws, pn->children[1], v, N, Q, params, fields, g) || g.rank != 2) return false;
    const pga_tree_node &pp = pn->children[2];
    std::vector<std::string> pv(np);
    if (pp->node_type == GA_NODE_CONSTANT && pp->tensor().size() == np) {
      for (size_t c = 0; c < np; ++c) pv[c] = num(pp->tensor()[c]);
    } else if (pp->node_type == GA_NODE_VAL && ws.variable_exists(pp->name) && ws.is_constant(pp->name) && !ws.associated_mf(pp->name) &&
               !ws.associated_im_data(pp->name) && ws.value(pp->name).size() == np) {
      for (size_t c = 0; c < np; ++c) {  // component c of a fixed-size constant vector: the parameter "name#c"
        const std::string key = pp->name + "#" + std::to_string(c);
        size_t k = 0;
        while (k < params.size() && params[k] != key) ++k;
        if (k == params.size()) params.push_back(key);
        if (params.size() > size_t(GFGPU_MAX_PARAMS)) return false;
        pv[c] = "par[" + std::to_string(k) + "]";
      }
    } else return false;
    args = g.code + "," + head;
    for (size_t c = 0; c < (lawfn == "iso" ? 5 : np); ++c) args += (c ? "," : "") + (c < np ? pv[c] : std::string("0.0"));
    args += tail;
    return true;
  };
  // nonlinear operators of ONE square-matrix argument (getfem_generic_assembly_functions_and_operators.cc, the large-strain
  // helpers of getfem_nonlinear_elasticity.cc:1930-2040): value, and first derivative contracted with a direction H --
  // "Derivative_1_Op(A):H", the only way the reference's symbolic differentiation uses them in an order-2 tree
  struct mat_operator { const char *name, *value; int rank; const char *deriv, *deriv2; };
  static const mat_operator mat_ops[] = {
      {"Det", "det", 0, "ddet", "d2det"},                         // d det[H] = det(A) tr(A^-1 H)
      {"Inv", "inv", 2, "dinv", "d2inv"},                         // d inv[H] = -A^-1 H A^-1
      {"Right_Cauchy_Green", "rcg", 2, "drcg", "d2rcg"},          // F'F ; H'F + F'H
      {"Left_Cauchy_Green", "lcg", 2, "dlcg", "d2lcg"},           // FF' ; HF' + FH'
      {"Green_Lagrangian", "glag", 2, "dglag", "d2glag"},         // (F'F - I)/2 ; (H'F + F'H)/2
      {"Matrix_i2", "mat_i2", 0, "dmat_i2", "d2mat_i2"},          // ((tr A)^2 - tr(A^2))/2 ; tr A tr H - tr(A H)
      {"Matrix_j1", "mat_j1", 0, "dmat_j1", "d2mat_j1"},          // tr(A) det(A)^(-1/3)
      {"Matrix_j2", "mat_j2", 0, "dmat_j2", "d2mat_j2"},          // i2(A) det(A)^(-2/3)
  };
  // der = 0: the value; 1: Derivative_1_Op; 2: Derivative_1_1_Op (an order-2 tree derived from a potential)
  auto mat_op_args = [&](const pga_tree_node &pn, size_type der, const mat_operator *&op, std::string &args) {
    if (pn->node_type != GA_NODE_PARAMS || pn->children.size() != 2) return false;
    const pga_tree_node &f = pn->children[0];
    if (f->node_type != GA_NODE_OPERATOR || f->der1 != (der ? 1 : 0) || f->der2 != (der == 2 ? 1 : 0)) return false;
    op = nullptr;
    for (const mat_operator &o : mat_ops) if (f->name == o.name) op = &o;
    if (!op) return false;
    jit_value g;
    if (!jit_emit(ws, pn->children[1], v, N, Q, params, fields, g) || g.rank != 2) return false;
    args = g.code;
    return true;
  };
  // "(Derivative_1_1_Norm(x) op H2) op H1" with op = "." (vector x) or ":" (matrix x): the second derivative of the norm in two
  // directions; fills x and H2 when node `nd` has that shape
  auto norm_second = [&](const pga_tree_node &nd, int op, jit_value &x, jit_value &h2) {
    if (nd->children.size() != 2) return false;
    const pga_tree_node &l = nd->children[0];
    if (l->node_type != GA_NODE_OP || int(l->op_type) != op || l->children.size() != 2) return false;
    const pga_tree_node &pn = l->children[0];
    if (pn->node_type != GA_NODE_PARAMS || pn->children.size() != 2) return false;
    const pga_tree_node &f = pn->children[0];
    if (f->node_type != GA_NODE_OPERATOR || f->name != "Norm" || f->der1 != 1 || f->der2 != 1) return false;
    return jit_emit(ws, pn->children[1], v, N, Q, params, fields, x) && jit_emit(ws, l->children[1], v, N, Q, params, fields, h2);
  };
  switch (n->node_type) {
    case GA_NODE_ZERO: {
      if (n->test_function_type != 0 && n->test_function_type != size_type(-1)) return false;
      base_tensor z = n->tensor();
      for (auto &x : z) x = 0.0;
      return tensor_const(z, out);
    }
    case GA_NODE_CONSTANT:
      if (n->test_function_type != 0 && n->test_function_type != size_type(-1)) return false;
      return tensor_const(n->tensor(), out);
    case GA_NODE_VAL: {
      if (n->name == v) { out = {"u", rv}; return true; }
      if (ws.variable_exists(n->name) && ws.is_constant(n->name) && !ws.variable_group_exists(n->name) &&
          (ws.associated_mf(n->name) || ws.associated_im_data(n->name))) {
        // a SCALAR fem-data / im-data coefficient (add_fem_constant, add_im_data): fld[k], evaluated at the Gauss point on its
        // own fem (ga_instruction_val, C&E.cc:636-690); at most two, all on one data mesh_fem (checked when the term is built)
        const getfem::mesh_fem *pmd = ws.associated_mf(n->name);
        const getfem::im_data *pid = ws.associated_im_data(n->name);
        if (pmd && int(pmd->get_qdim()) == N && N > 1) {  // ONE vector-valued fem-data field (an advection velocity, a body force direction)
          if (!fields.empty() && fields[0] != n->name) return false;
          if (fields.empty()) fields.push_back(n->name);
          out = {"vfld", 1};
          return true;
        }
        if (pmd ? pmd->get_qdim() != 1 : pid->nb_tensor_elem() != 1) return false;
        if (!fields.empty() && ws.associated_mf(fields[0]) && int(ws.associated_mf(fields[0])->get_qdim()) != 1) return false;
        size_t k = 0;
        while (k < fields.size() && fields[k] != n->name) ++k;
        if (k == fields.size()) fields.push_back(n->name);
        if (fields.size() > 2) return false;
        if (pmd ? ws.associated_mf(fields[0]) != pmd : ws.associated_im_data(fields[0]) != pid) return false;
        out = {"fld[" + std::to_string(k) + "]", 0};
        return true;
      }
      if (!ws.variable_exists(n->name) || !ws.is_constant(n->name) || ws.associated_mf(n->name) || ws.associated_im_data(n->name)) return false;
      if (ws.value(n->name).size() != 1) {
        // a fixed-size VECTOR (N entries) or MATRIX (N x N, first index fastest like every GetFEM tensor) constant -- a load
        // direction, an anisotropic diffusion tensor "(A*Grad_u).Grad_Test_u": its components are the parameters "name#c"
        const bgeot::multi_index &sz = n->tensor().sizes();
        const bool isvec = sz.size() == 1 && sz[0] == size_type(N), ismat = sz.size() == 2 && sz[0] == size_type(N) && sz[1] == size_type(N);
        if ((!isvec && !ismat) || ws.value(n->name).size() != (isvec ? size_t(N) : size_t(N * N))) return false;
        std::vector<std::string> pv;
        for (size_t c = 0; c < ws.value(n->name).size(); ++c) {
          const std::string key = n->name + "#" + std::to_string(c);
          size_t k = 0;
          while (k < params.size() && params[k] != key) ++k;
          if (k == params.size()) params.push_back(key);
          if (params.size() > size_t(GFGPU_MAX_PARAMS)) return false;
          pv.push_back("par[" + std::to_string(k) + "]");
        }
        if (isvec) {
          out = {"mkvec(" + pv[0] + "," + pv[1] + "," + (N > 2 ? pv[2] : std::string("0.0")) + ")", 1};
        } else {
          std::string c = "mkmat(";
          for (int k = 0; k < 9; ++k) c += (k ? "," : "") + (k < N * N ? pv[k] : std::string("0.0"));
          out = {c + ")", 2};
        }
        return true;
      }
      size_t k = 0;
      while (k < params.size() && params[k] != n->name) ++k;
      if (k == params.size()) params.push_back(n->name);
      if (params.size() > size_t(GFGPU_MAX_PARAMS)) return false;
      out = {"par[" + std::to_string(k) + "]", 0};
      return true;
    }
    case GA_NODE_C_MATRIX: {
      // an explicit vector "[a; b; c]" or matrix "[a, b; c, d]" of scalar expressions (test functions included: they are values
      // here): children in tensor storage order, first index fastest (ga_instruction_c_matrix_with_tests) -- mkmat's convention
      const size_t nc = n->children.size();
      const bgeot::multi_index &sz = n->tensor().sizes();
      const size_t nb = size_t(n->nbc1);
      if (nb < 1 || nb > 2 || sz.size() < nb) return false;
      bool shape_ok = true;
      for (size_t k = 0; k < nb; ++k) shape_ok = shape_ok && sz[sz.size() - nb + k] == size_type(N);
      if (!shape_ok || nc != (nb == 1 ? size_t(N) : size_t(N * N))) return false;
      std::vector<jit_value> ch(nc);
      for (size_t k = 0; k < nc; ++k)
        if (!child(k, ch[k]) || ch[k].rank != 0) return false;
      if (nb == 1) {
        out = {"mkvec(" + ch[0].code + "," + ch[1].code + "," + (N > 2 ? ch[2].code : std::string("0.0")) + ")", 1};
      } else {
        std::string c = "mkmat(";
        for (int k = 0; k < 9; ++k) c += (k ? "," : "") + (k < N * N ? ch[k].code : std::string("0.0"));
        out = {c + ")", 2};
      }
      return true;
    }
    case GA_NODE_X:  // the position: the whole vector (nbc1 == 0) or one coordinate
      if (n->nbc1 == 0) { out = {"X", 1}; return true; }
      if (int(n->nbc1) > N) return false;
      out = {"X.v[" + std::to_string(int(n->nbc1) - 1) + "]", 0};
      return true;
    case GA_NODE_NORMAL:  // the unit outward normal of a boundary face (the term refuses a region of convexes)
      out = {"Normal", 1};
      return true;
    case GA_NODE_GRAD:
      if (n->name != v) return false;
      out = {"gu", rv + 1};
      return true;
    case GA_NODE_DIVERG:
      if (n->name != v || Q == 1) return false;
      out = {"trace(gu)", 0};
      return true;
    case GA_NODE_VAL_TEST:
      if (n->name != v) return false;
      out = {n->test_function_type == 2 ? "t2v" : "tv", rv};
      return n->test_function_type == 1 || n->test_function_type == 2;
    case GA_NODE_GRAD_TEST:
      if (n->name != v) return false;
      out = {n->test_function_type == 2 ? "t2g" : "tg", rv + 1};
      return n->test_function_type == 1 || n->test_function_type == 2;
    case GA_NODE_DIVERG_TEST:
      if (n->name != v || Q == 1) return false;
      out = {n->test_function_type == 2 ? "trace(t2g)" : "trace(tg)", 0};
      return n->test_function_type == 1 || n->test_function_type == 2;
    case GA_NODE_OP: {
      jit_value a, b;
      switch (n->op_type) {
        case GA_UNARY_MINUS:
          if (!child(0, a)) return false;
          out = {"(-" + a.code + ")", a.rank};
          return true;
        case GA_QUOTE:
          if (!child(0, a)) return false;
          out = {"transp(" + a.code + ")", a.rank};
          return true;
        case GA_TRACE:
          if (!child(0, a) || a.rank != 2) return false;
          out = {"trace(" + a.code + ")", 0};
          return true;
        case GA_SYM: case GA_SKEW: case GA_DEVIATOR:
          if (!child(0, a) || a.rank != 2) return false;
          out = {std::string(n->op_type == GA_SYM ? "sym(" : n->op_type == GA_SKEW ? "skew(" : "deviator(") + a.code + ")", 2};
          return true;
        case GA_PLUS: case GA_MINUS:
          if (!child(0, a) || !child(1, b) || a.rank != b.rank) return false;
          out = {"(" + a.code + (n->op_type == GA_PLUS ? "+" : "-") + b.code + ")", a.rank};
          return true;
        case GA_MULT:
          if (!child(0, a) || !child(1, b)) return false;
          if (a.rank == 0 || b.rank == 0) { out = {"(" + a.code + "*" + b.code + ")", a.rank + b.rank}; return true; }
          if (a.rank == 2 && b.rank == 1) { out = {"(" + a.code + "*" + b.code + ")", 1}; return true; }
          if (a.rank == 2 && b.rank == 2) { out = {"(" + a.code + "*" + b.code + ")", 2}; return true; }
          return false;
        case GA_DOTMULT:  // componentwise: a product as soon as one side is a scalar (the differentiation writes f'(u).*Test2_u)
          if (!child(0, a) || !child(1, b) || (a.rank && b.rank)) return false;
          out = {"(" + a.code + "*" + b.code + ")", a.rank + b.rank};
          return true;
        case GA_DIV: case GA_DOTDIV:
          if (!child(0, a) || !child(1, b) || b.rank) return false;
          out = {"(" + a.code + "/" + b.code + ")", a.rank};
          return true;
        case GA_DOT:  // contraction of the last index of a with the first of b
          {
            jit_value x, h2;
            if (norm_second(n, GA_DOT, x, h2)) {
              if (!child(1, b) || b.rank != 1 || x.rank != 1 || h2.rank != 1) return false;
              out = {"d2norm(" + x.code + "," + h2.code + "," + b.code + ")", 0};
              return true;
            }
          }
          if (!child(0, a) || !child(1, b)) return false;
          out = {"dot(" + a.code + "," + b.code + ")", (a.rank && b.rank) ? a.rank + b.rank - 2 : a.rank + b.rank};
          return true;
        case GA_COLON: {
          {
            jit_value x, h2;
            if (norm_second(n, GA_COLON, x, h2)) {
              if (!child(1, b) || b.rank != 2 || x.rank != 2 || h2.rank != 2) return false;
              out = {"d2norm(" + x.code + "," + h2.code + "," + b.code + ")", 0};
              return true;
            }
          }
          std::string sa;  // Derivative_1_Saint_Venant_Kirchhoff_PK2(G, params):H = the directional derivative dS[H]
          std::string lawfn;
          if (n->children.size() == 2 && svk_args(n->children[0], 1, sa, &lawfn)) {
            if (!child(1, b) || b.rank != 2) return false;
            out = {lawfn + "_dpk2(" + sa + "," + b.code + ")", 2};
            return true;
          }
          const mat_operator *mop = nullptr;
          if (n->children.size() == 2 && mat_op_args(n->children[0], 1, mop, sa)) {
            if (!child(1, b) || b.rank != 2) return false;
            out = {std::string(mop->deriv) + "(" + sa + "," + b.code + ")", mop->rank};
            return true;
          }
          // "X:(Derivative_1_1_Op(A):Y)" with a SCALAR-valued operator (the chain rule of f(G(F)) puts the second derivative of f
          // between the two first derivatives of G): D2f[X, Y], symmetric
          if (n->children.size() == 2 && n->children[1]->node_type == GA_NODE_OP && n->children[1]->op_type == GA_COLON &&
              n->children[1]->children.size() == 2 && mat_op_args(n->children[1]->children[0], 2, mop, sa) && mop->rank == 0) {
            jit_value y;
            if (!jit_emit(ws, n->children[1]->children[1], v, N, Q, params, fields, y) || y.rank != 2) return false;
            if (!child(0, a) || a.rank != 2) return false;
            out = {std::string(mop->deriv2) + "(" + sa + "," + a.code + "," + y.code + ")", 0};
            return true;
          }
          // "(Derivative_1_1_Op(A):H2):H1": the second derivative in two directions
          if (n->children.size() == 2 && n->children[0]->node_type == GA_NODE_OP && n->children[0]->op_type == GA_COLON &&
              n->children[0]->children.size() == 2 && mat_op_args(n->children[0]->children[0], 2, mop, sa)) {
            jit_value h2;
            if (!jit_emit(ws, n->children[0]->children[1], v, N, Q, params, fields, h2) || h2.rank != 2) return false;
            if (!child(1, b) || b.rank != 2) return false;
            out = {std::string(mop->deriv2) + "(" + sa + "," + h2.code + "," + b.code + ")", mop->rank};
            return true;
          }
        }
          if (!child(0, a) || !child(1, b) || a.rank != b.rank) return false;
          out = {"ddot(" + a.code + "," + b.code + ")", 0};
          return true;
        case GA_TMULT:
          if (!child(0, a) || !child(1, b) || a.rank != 1 || b.rank != 1) return false;
          out = {"outer(" + a.code + "," + b.code + ")", 2};
          return true;
        default: return false;
      }
    }
    case GA_NODE_PARAMS: {
      if (n->children.empty()) return false;
      const pga_tree_node &f = n->children[0];
      if (f->node_type == GA_NODE_PREDEF_FUNC) {
        static const std::map<std::string, std::string> fn = {
            {"sqrt", "sqrt"}, {"sqr", "sqr"}, {"pow", "pow"}, {"exp", "exp"}, {"log", "log"}, {"log10", "log10"}, {"sinh", "sinh"},
            {"cosh", "cosh"}, {"tanh", "tanh"}, {"asinh", "asinh"}, {"acosh", "acosh"}, {"atanh", "atanh"}, {"sin", "sin"},
            {"cos", "cos"}, {"tan", "tan"}, {"asin", "asin"}, {"acos", "acos"}, {"atan", "atan"}, {"atan2", "atan2"}, {"erf", "erf"},
            {"erfc", "erfc"}, {"Heaviside", "Heaviside"}, {"sign", "sign"}, {"abs", "fabs"}, {"pos_part", "pos_part"},
            {"neg_part", "neg_part"}, {"sqr_pos_part", "sqr_pos_part"}, {"sqr_neg_part", "sqr_neg_part"},
            {"half_sqr_pos_part", "half_sqr_pos_part"}, {"half_sqr_neg_part", "half_sqr_neg_part"}, {"max", "fmax"}, {"min", "fmin"},
            {"DER_PDFUNC_SQRT", "DER_PDFUNC_SQRT"}, {"DER_PDFUNC1_POW", "DER_PDFUNC1_POW"}, {"DER_PDFUNC2_POW", "DER_PDFUNC2_POW"},
            {"DER_PDFUNC_LOG", "DER_PDFUNC_LOG"}, {"DER_PDFUNC_LOG10", "DER_PDFUNC_LOG10"}, {"DER_PDFUNC_TANH", "DER_PDFUNC_TANH"},
            {"DER_PDFUNC_ASINH", "DER_PDFUNC_ASINH"}, {"DER_PDFUNC_ACOSH", "DER_PDFUNC_ACOSH"}, {"DER_PDFUNC_ATANH", "DER_PDFUNC_ATANH"},
            {"DER_PDFUNC_COS", "DER_PDFUNC_COS"}, {"DER_PDFUNC_TAN", "DER_PDFUNC_TAN"}, {"DER_PDFUNC_ASIN", "DER_PDFUNC_ASIN"},
            {"DER_PDFUNC_ACOS", "DER_PDFUNC_ACOS"}, {"DER_PDFUNC_ATAN", "DER_PDFUNC_ATAN"}, {"DER_PDFUNC1_ATAN2", "DER_PDFUNC1_ATAN2"},
            {"DER_PDFUNC2_ATAN2", "DER_PDFUNC2_ATAN2"}, {"DER_PDFUNC_ERF", "DER_PDFUNC_ERF"}, {"DER_PDFUNC_ERFC", "DER_PDFUNC_ERFC"},
            {"DER_PDFUNC_NEG_PART", "DER_PDFUNC_NEG_PART"},
            {"sinc", "sinc"}, {"DER_PDFUNC_SINC", "DER_PDFUNC_SINC"}, {"DER2_PDFUNC_SINC", "DER2_PDFUNC_SINC"},
            {"DER_PDFUNC1_MAX", "DER_PDFUNC1_MAX"}, {"DER_PDFUNC2_MAX", "DER_PDFUNC2_MAX"},
            {"DER_PDFUNC1_DER_PDFUNC1_POW", "DER_PDFUNC1_DER_PDFUNC1_POW"}, {"DER_PDFUNC2_DER_PDFUNC1_POW", "DER_PDFUNC2_DER_PDFUNC1_POW"},
            {"DER_PDFUNC1_DER_PDFUNC2_POW", "DER_PDFUNC1_DER_PDFUNC2_POW"}, {"DER_PDFUNC2_DER_PDFUNC2_POW", "DER_PDFUNC2_DER_PDFUNC2_POW"},
            // second derivatives (the derivative of a derivative the reference defines by an expression)
            {"DER_PDFUNC_DER_PDFUNC_SQRT", "DER_PDFUNC_DER_PDFUNC_SQRT"}, {"DER_PDFUNC_DER_PDFUNC_LOG", "DER_PDFUNC_DER_PDFUNC_LOG"},
            {"DER_PDFUNC_DER_PDFUNC_LOG10", "DER_PDFUNC_DER_PDFUNC_LOG10"}, {"DER_PDFUNC_DER_PDFUNC_TANH", "DER_PDFUNC_DER_PDFUNC_TANH"},
            {"DER_PDFUNC_DER_PDFUNC_ASINH", "DER_PDFUNC_DER_PDFUNC_ASINH"}, {"DER_PDFUNC_DER_PDFUNC_ACOSH", "DER_PDFUNC_DER_PDFUNC_ACOSH"},
            {"DER_PDFUNC_DER_PDFUNC_ATANH", "DER_PDFUNC_DER_PDFUNC_ATANH"}, {"DER_PDFUNC_DER_PDFUNC_COS", "DER_PDFUNC_DER_PDFUNC_COS"},
            {"DER_PDFUNC_DER_PDFUNC_TAN", "DER_PDFUNC_DER_PDFUNC_TAN"}, {"DER_PDFUNC_DER_PDFUNC_ASIN", "DER_PDFUNC_DER_PDFUNC_ASIN"},
            {"DER_PDFUNC_DER_PDFUNC_ACOS", "DER_PDFUNC_DER_PDFUNC_ACOS"}, {"DER_PDFUNC_DER_PDFUNC_ATAN", "DER_PDFUNC_DER_PDFUNC_ATAN"}};
        auto it = fn.find(f->name);
        if (it == fn.end() || n->children.size() < 2 || n->children.size() > 3) return false;
        std::string code = it->second + "(";
        for (size_t k = 1; k < n->children.size(); ++k) {
          jit_value a;
          if (!child(k, a) || a.rank) return false;
          code += (k > 1 ? "," : "") + a.code;
        }
        out = {code + ")", 0};
        return true;
      }
      if (f->node_type == GA_NODE_RESHAPE && n->children.size() == 4 && n->children[1]->node_type == GA_NODE_VAL) {
        // "Reshape(A, N, N)" of a fixed-size constant with N x N entries (first index fastest): a matrix of parameters
        const pga_tree_node &pv0 = n->children[1];
        auto cst = [&](size_t k) {
          return n->children[k]->node_type == GA_NODE_CONSTANT && n->children[k]->tensor().size() == 1 &&
                 int(n->children[k]->tensor()[0]) == N;
        };
        if (!cst(2) || !cst(3) || !ws.variable_exists(pv0->name) || !ws.is_constant(pv0->name) || ws.associated_mf(pv0->name) ||
            ws.associated_im_data(pv0->name) || ws.value(pv0->name).size() != size_t(N * N))
          return false;
        std::string c = "mkmat(";
        for (int k = 0; k < 9; ++k) {
          std::string e = "0.0";
          if (k < N * N) {
            const std::string key = pv0->name + "#" + std::to_string(k);
            size_t q = 0;
            while (q < params.size() && params[q] != key) ++q;
            if (q == params.size()) params.push_back(key);
            if (params.size() > size_t(GFGPU_MAX_PARAMS)) return false;
            e = "par[" + std::to_string(q) + "]";
          }
          c += (k ? "," : "") + e;
        }
        out = {c + ")", 2};
        return true;
      }
      {
        std::string sa;
        std::string lawfn;
        if (svk_args(n, 0, sa, &lawfn)) { out = {lawfn + "_pk2(" + sa + ")", 2}; return true; }
        const mat_operator *mop = nullptr;
        if (mat_op_args(n, 0, mop, sa)) { out = {std::string(mop->value) + "(" + sa + ")", mop->rank}; return true; }
      }
      if (f->node_type == GA_NODE_CROSS_PRODUCT && n->children.size() == 3 && N == 3) {
        jit_value a, b;  // bilinear: the reference's differentiation writes Cross_product(Test2_u, b) inline
        if (!child(1, a) || !child(2, b) || a.rank != 1 || b.rank != 1) return false;
        out = {"cross(" + a.code + "," + b.code + ")", 1};
        return true;
      }
      if (f->node_type == GA_NODE_OPERATOR && (f->name == "Norm_sqr" || f->name == "Norm") && n->children.size() == 2 &&
          f->der1 == 1 && f->der2 == 0) {  // Derivative_1_Norm(x) = x/|x| (0 at 0), Derivative_1_Norm_sqr(x) = 2 x: same rank as x
        jit_value a;
        if (!child(1, a) || a.rank < 1) return false;
        out = {f->name == "Norm" ? "dnorm(" + a.code + ")" : "((2.0)*" + a.code + ")", a.rank};
        return true;
      }
      if (f->node_type == GA_NODE_OPERATOR && (f->name == "Norm_sqr" || f->name == "Norm") && n->children.size() == 2 &&
          f->der1 == 0 && f->der2 == 0) {
        jit_value a;
        if (!child(1, a)) return false;
        out = {(f->name == "Norm" ? "gnorm(" : "normsqr(") + a.code + ")", 0};
        return true;
      }
      // component access: v(i) of a vector, M(i,j) of a matrix, with constant indices
      auto index = [&](size_t k, int &o) {
        if (k >= n->children.size() || n->children[k]->node_type != GA_NODE_CONSTANT || n->children[k]->tensor().size() != 1) return false;
        o = int(n->children[k]->tensor()[0]);
        return o >= 1 && o <= N;
      };
      jit_value a;
      int i1 = 0, i2 = 0;
      if (n->children.size() == 2 && index(1, i1) && child(0, a) && a.rank == 1) {
        out = {"(" + a.code + ").v[" + std::to_string(i1 - 1) + "]", 0};
        return true;
      }
      // slices of a matrix: M(:,j) (column j) and M(i,:) (row i)
      auto all_idx = [&](size_t k) { return k < n->children.size() && n->children[k]->node_type == GA_NODE_ALLINDICES; };
      if (n->children.size() == 3 && all_idx(1) && index(2, i2) && child(0, a) && a.rank == 2) {
        out = {"mcol(" + a.code + "," + std::to_string(i2 - 1) + ")", 1};
        return true;
      }
      if (n->children.size() == 3 && index(1, i1) && all_idx(2) && child(0, a) && a.rank == 2) {
        out = {"mrow(" + a.code + "," + std::to_string(i1 - 1) + ")", 1};
        return true;
      }
      if (n->children.size() == 3 && index(1, i1) && index(2, i2) && child(0, a) && a.rank == 2) {
        out = {"(" + a.code + ").m[" + std::to_string(i1 - 1) + "][" + std::to_string(i2 - 1) + "]", 0};
        return true;
      }
      return false;
    }
    default: return false;
  }
}

// Translates the order-1 tree `i1` of the scalar variable v (and the order-2 tree on (v, v) with the same integration method and
// region, if there is one) into a JIT term.  Returns false when something in the trees is outside the translator's language.
static bool recognise_jit(const getfem::ga_workspace &ws, size_type i1, recognised_term &out) {
  const auto &td = ws.tree_info(i1);
  if (td.order != 1 || td.operation != getfem::ga_workspace::ASSEMBLY) return false;
  const std::string &v = td.name_test1;
  const getfem::mesh_fem *pmf = ws.associated_mf(v);
  if (!pmf || ws.is_constant(v)) return false;
  const int N = int(pmf->linked_mesh().dim()), Q = int(pmf->get_qdim());
  if ((N != 2 && N != 3) || (Q != 1 && Q != N)) return false;
  std::vector<std::string> params, fields;
  jit_value f1, f2{"(0.0)", 0};
  if (!td.ptree || !jit_emit(ws, td.ptree->root, v, N, Q, params, fields, f1) || f1.rank != 0) return false;
  for (size_type j = 0; j < ws.nb_trees(); ++j) {
    const auto &t2 = ws.tree_info(j);
    if (t2.order == 2 && t2.mim == td.mim && t2.rg == td.rg && t2.name_test1 == v && t2.name_test2 == v) {
      if (!t2.ptree || !jit_emit(ws, t2.ptree->root, v, N, Q, params, fields, f2) || f2.rank != 0) return false;
    } else if (t2.order == 2 && t2.mim == td.mim && t2.rg == td.rg && (t2.name_test1 == v) != (t2.name_test2 == v)) {
      return false;  // coupled to another variable: not this route
    }
  }
  // the order-0 tree the two were derived from (add_expression of a potential): same integration method and region; it must
  // translate too, without test functions -- then assembly(0) runs on the device as well
  jit_value f0{"", 0};
  for (size_type j = 0; j < ws.nb_trees(); ++j) {
    const auto &t0 = ws.tree_info(j);
    if (t0.order != 0 || t0.mim != td.mim || t0.rg != td.rg || t0.operation != getfem::ga_workspace::ASSEMBLY) continue;
    jit_value g;
    if (!f0.code.empty() || !t0.ptree || !jit_emit(ws, t0.ptree->root, v, N, Q, params, fields, g) || g.rank != 0) { f0.code.clear(); break; }
    f0 = g;
  }
  out = recognised_term();
  out.family = GFGPU_JIT;
  out.varname = v;
  out.jit_form0 = f0.code;
  out.jit_form1 = f1.code;
  out.jit_form2 = f2.code;
  out.jit_params = params;
  out.field_names = fields;
  out.field_sign = 1.0;
  for (const std::string &pn : params) {  // "name" = a scalar constant, "name#c" = component c of a fixed-size vector constant
    const size_t h = pn.find('#');
    out.params.push_back(h == std::string::npos ? ws.value(pn)[0] : ws.value(pn.substr(0, h))[std::stoul(pn.substr(h + 1))]);
  }
  return true;
}

// COUPLED div-pressure parts, as the reference prints them after its semantic analysis (the incompressibility bricks add
// "-p*Div_Test_u - Test_p*Div_u", getfem_models.cc add_linear_incompressibility; the test driver ref_coupled.cc prints the trees):
//   order 2 (Test_u, Test2_p)  "(-Test2_p)*Div_Test_u"     order 2 (Test_p, Test2_u)  "-(Test_p*Div_Test2_u)"
//   order 1  Test_u            "(-p)*Div_Test_u"           order 1  Test_p            "-(Test_p*Div_u)"
// and the same without the minus signs.  u must be a vector fem variable of the mesh dimension, p a scalar one.
static bool recognise_coupled(const getfem::ga_workspace &ws, const std::string &s0, int order, const std::string &test1,
                              const std::string &test2, recognised_term &out) {
  const std::string s = strip_outer(s0), ID = "([A-Za-z_][A-Za-z_0-9]*)";
  std::smatch m;
  // capture 1 = the scalar variable (A), capture 2 = the vector variable (B)
  std::string A, B;
  bool found = false;
  if (order == 2) {  // coupled mass: "Test_a:Test2_b" (asm_mass_matrix on two fems), also with '.', '*' and the factors swapped
    std::string ta, tb;
    if (std::regex_match(s, m, std::regex("Test_" + ID + "[.:*]Test2_" + ID))) { ta = m[1]; tb = m[2]; }
    else if (std::regex_match(s, m, std::regex("Test2_" + ID + "[.:*]Test_" + ID))) { ta = m[2]; tb = m[1]; }
    if (!ta.empty() && ta == test1 && tb == test2 && ta != tb && ws.variable_exists(ta) && ws.variable_exists(tb)) {
      const getfem::mesh_fem *mfa = ws.associated_mf(ta), *mfb = ws.associated_mf(tb);
      if (mfa && mfb && &mfa->linked_mesh() == &mfb->linked_mesh() && mfa->get_qdim() == mfb->get_qdim()) {
        out.family = GFGPU_SHIM_COUPLED_MASS;
        out.varname = test1;
        out.varname_u = ta;  // rows
        out.varname_p = tb;  // columns
        out.transposed = false;
        out.sign = 1.0;
        out.params.clear();
        out.field_names.clear();
        return true;
      }
    }
  }
  if (order == 2) {
    const std::vector<std::pair<std::string, double>> fu = {{"\\(-Test2_" + ID + "\\)\\*Div_Test_" + ID, -1.0},
                                                            {"-\\(Test2_" + ID + "\\*Div_Test_" + ID + "\\)", -1.0},
                                                            {"Test2_" + ID + "\\*Div_Test_" + ID, 1.0}};
    const std::vector<std::pair<std::string, double>> fp = {{"-\\(Test_" + ID + "\\*Div_Test2_" + ID + "\\)", -1.0},
                                                            {"\\(-Test_" + ID + "\\)\\*Div_Test2_" + ID, -1.0},
                                                            {"Test_" + ID + "\\*Div_Test2_" + ID, 1.0}};
    for (const auto &f : fu)
      if (!found && std::regex_match(s, m, std::regex(f.first))) { A = m[1]; B = m[2]; out.sign = f.second; out.transposed = false; found = true; }
    for (const auto &f : fp)
      if (!found && std::regex_match(s, m, std::regex(f.first))) { A = m[1]; B = m[2]; out.sign = f.second; out.transposed = true; found = true; }
    if (!found) return false;
    if (out.transposed ? (test1 != A || test2 != B) : (test1 != B || test2 != A)) return false;
  } else {
    const std::vector<std::pair<std::string, double>> fu = {{"\\(-" + ID + "\\)\\*Div_Test_" + ID, -1.0},
                                                            {"-\\(" + ID + "\\*Div_Test_" + ID + "\\)", -1.0},
                                                            {ID + "\\*Div_Test_" + ID, 1.0}};
    const std::vector<std::pair<std::string, double>> fp = {{"-\\(Test_" + ID + "\\*Div_" + ID + "\\)", -1.0},
                                                            {"\\(-Test_" + ID + "\\)\\*Div_" + ID, -1.0},
                                                            {"Test_" + ID + "\\*Div_" + ID, 1.0}};
    for (const auto &f : fu)
      if (!found && std::regex_match(s, m, std::regex(f.first))) { A = m[1]; B = m[2]; out.sign = f.second; out.transposed = false; found = true; }
    for (const auto &f : fp)
      if (!found && std::regex_match(s, m, std::regex(f.first))) { A = m[1]; B = m[2]; out.sign = f.second; out.transposed = true; found = true; }
    if (!found) return false;
    if (test1 != (out.transposed ? A : B)) return false;
  }
  if (A == B || !ws.variable_exists(A) || !ws.variable_exists(B) || ws.is_constant(A) || ws.is_constant(B)) return false;
  const getfem::mesh_fem *mfp = ws.associated_mf(A), *mfu = ws.associated_mf(B);
  if (!mfp || !mfu || &mfp->linked_mesh() != &mfu->linked_mesh()) return false;
  if (mfp->get_qdim() != 1 || mfu->get_qdim() != mfu->linked_mesh().dim()) return false;
  out.family = GFGPU_SHIM_COUPLED_DIV;
  out.varname = test1;
  out.varname_u = B;
  out.varname_p = A;
  out.params.clear();
  out.field_names.clear();
  return true;
}

bool recognise_tree_sum(const getfem::ga_workspace &ws, size_type itree, std::vector<recognised_term> &out) {
  const getfem::ga_workspace::tree_description &td = ws.tree_info(itree);
  if (td.order == 2 && td.operation == getfem::ga_workspace::ASSEMBLY && td.name_test1 != td.name_test2) {
    recognised_term rt;
    out.clear();
    if (!recognise_coupled(ws, strip(getfem::ga_tree_to_string(*td.ptree)), 2, td.name_test1, td.name_test2, rt)) return false;
    out.push_back(rt);
    return true;
  }
  if (td.order == 2 && td.operation == getfem::ga_workspace::ASSEMBLY && td.name_test1 == td.name_test2) {
    recognised_term rt;
    out.clear();
    if (!recognise_order2_string(ws, td.name_test1, strip(getfem::ga_tree_to_string(*td.ptree)), rt) &&
        !recognise_by_probe(ws, itree, rt))
      return false;
    out.push_back(rt);
    return true;
  }
  if (td.order != 1 || td.operation != getfem::ga_workspace::ASSEMBLY) return false;
  out.clear();
  if (!recognise_sum(ws, td.name_test1, strip(getfem::ga_tree_to_string(*td.ptree)), out)) {
    recognised_term rt;
    out.clear();
    bool probed = false;
    try { probed = recognise_by_probe(ws, itree, rt); } catch (const gmm::gmm_error &) { probed = false; }
    if (!probed && !recognise_jit(ws, itree, rt)) return false;  // last resort: the NVRTC route (scalar variables)
    out.push_back(rt);
    return true;
  }
  size_t bilinear = 0;
  for (const recognised_term &rt : out)
    bilinear += rt.family != GFGPU_SOURCE && rt.family != GFGPU_NORMAL_SOURCE && rt.family != GFGPU_SHIM_COUPLED_DIV;
  if (bilinear > 1) {  // the reference thresholds the element matrix of the SUM: one run-time compiled term does exactly that
    recognised_term rt;
    if (recognise_jit(ws, itree, rt)) {
      out.clear();
      out.push_back(rt);
      return true;
    }
  }
  GMM_ASSERT1(bilinear <= 1, "gfgpu: several bilinear forms summed on one region are thresholded together by the reference "
                             "(C&E.cc:4889); give them distinct regions or one expression family");
  return true;
}

static bool recognise_string(const getfem::ga_workspace &ws, const std::string &v, const std::string &s, recognised_term &out) {
  const std::string ID = "([A-Za-z_][A-Za-z_0-9]*)";
  const std::string I3 = "\\[\\[1,0,0\\],\\[0,1,0\\],\\[0,0,1\\]\\]", I2 = "\\[\\[1,0\\],\\[0,1\\]\\]";
  const std::string Idm = "(?:" + I3 + "|" + I2 + ")";
  std::smatch m;
  out.varname = v;
  out.params.clear();
  out.field_names.clear();
  out.field_sign = 1.0;
  auto scalar = [&](const std::string &name) {
    GMM_ASSERT1(ws.is_constant(name), "gfgpu: '" << name << "' must be a constant");
    if (ws.associated_mf(name) || ws.associated_im_data(name)) {  // fem data (add_fem_constant) or im data (add_im_data): a field
      out.field_names.push_back(name);
      return 1.0;
    }
    GMM_ASSERT1(ws.value(name).size() == 1, "gfgpu: '" << name << "' must be a scalar constant");
    return ws.value(name)[0];
  };
  if (std::regex_match(s, m, std::regex("\\(" + ID + "\\*Grad_" + v + "\\)[.:]Grad_Test_" + v))) {
    out.family = GFGPU_LAPLACE; out.params = {scalar(m[1])}; return true;
  }
  if (std::regex_match(s, std::regex("Grad_" + v + "[.:]Grad_Test_" + v))) {
    out.family = GFGPU_LAPLACE; out.params = {1.0}; return true;
  }
  if (std::regex_match(s, m, std::regex("\\(" + ID + "\\*" + v + "\\)\\.Test_" + v))) {
    out.family = GFGPU_MASS; out.params = {scalar(m[1])}; return true;
  }
  if (std::regex_match(s, std::regex(v + "\\.Test_" + v))) {
    out.family = GFGPU_MASS; out.params = {1.0}; return true;
  }
  {  // volumic source term (add_source_term_brick, getfem_models.cc:4124-): "(-f)*Test_u", "(-f).Test_u", "-(f.Test_u)", "f.Test_u"
    double sign = 0;
    std::string name;
    if (std::regex_match(s, m, std::regex("\\(-" + ID + "\\)[.*:]Test_" + v))) { sign = -1; name = m[1]; }
    else if (std::regex_match(s, m, std::regex("-\\(" + ID + "[.*:]Test_" + v + "\\)"))) { sign = -1; name = m[1]; }
    else if (std::regex_match(s, m, std::regex(ID + "[.*:]Test_" + v))) { sign = 1; name = m[1]; }  // (":" = asm_source_term)
    if (sign != 0 && name != v && ws.is_constant(name) && !ws.variable_group_exists(name)) {
      const getfem::mesh_fem *pmf = ws.associated_mf(v);
      out.family = GFGPU_SOURCE;
      if (const getfem::mesh_fem *pmd = ws.associated_mf(name)) {  // distributed load: fem data of the variable's qdim
        GMM_ASSERT1(pmf && pmd->get_qdim() == pmf->get_qdim(), "gfgpu: the load's mesh_fem must have the qdim of the variable");
        out.field_names.push_back(name);
        out.field_sign = sign;
        out.params.assign(pmf->get_qdim(), 0.0);
        return true;
      }
      if (const getfem::im_data *pid = ws.associated_im_data(name)) {  // the load stored per Gauss point
        GMM_ASSERT1(pmf && pid->nb_tensor_elem() == pmf->get_qdim(), "gfgpu: the load's im_data must have qdim components");
        out.field_names.push_back(name);
        out.field_sign = sign;
        out.params.assign(pmf->get_qdim(), 0.0);
        return true;
      }
      GMM_ASSERT1(pmf && ws.value(name).size() == pmf->get_qdim(),
                  "gfgpu: the source term needs a fixed-size constant with qdim components");
      for (size_type k = 0; k < ws.value(name).size(); ++k) out.params.push_back(sign * ws.value(name)[k]);
      return true;
    }
  }
  {  // normal source term (add_normal_source_term_brick, getfem_models.cc:4290-4299), on regions of faces:
     // "(g.Normal)*Test_u" (scalar u, g of meshdim components) / "(Reshape(g,Q,N)*Normal).Test_u" (A(b,n) = g[b + Q n])
    double sign = 0;
    std::string name;
    if (std::regex_match(s, m, std::regex("\\(" + ID + "\\.Normal\\)\\*Test_" + v))) { sign = 1; name = m[1]; }
    else if (std::regex_match(s, m, std::regex("\\(-\\(" + ID + "\\.Normal\\)\\)\\*Test_" + v))) { sign = -1; name = m[1]; }
    else if (std::regex_match(s, m, std::regex("\\(Reshape\\(" + ID + ",\\d+,\\d+\\)\\*Normal\\)\\.Test_" + v))) { sign = 1; name = m[1]; }
    else if (std::regex_match(s, m, std::regex("\\(-\\(Reshape\\(" + ID + ",\\d+,\\d+\\)\\*Normal\\)\\)\\.Test_" + v))) { sign = -1; name = m[1]; }
    if (sign != 0 && ws.is_constant(name)) {
      const getfem::mesh_fem *pmf = ws.associated_mf(v);
      GMM_ASSERT1(pmf && ws.value(name).size() == size_type(pmf->get_qdim()) * pmf->linked_mesh().dim(),
                  "gfgpu: the normal source term needs a fixed-size constant with qdim x meshdim components");
      out.family = GFGPU_NORMAL_SOURCE;
      for (size_type k = 0; k < ws.value(name).size(); ++k) out.params.push_back(sign * ws.value(name)[k]);
      return true;
    }
  }
  if (std::regex_match(s, m, std::regex("\\(\\(Div_" + v + "\\*\\(" + ID + "\\*" + Idm + "\\)\\)\\+\\(\\(2\\*" + ID +
                                        "\\)\\*\\(Sym\\(Grad_" + v + "\\)\\)\\)\\):Grad_Test_" + v))) {
    GMM_ASSERT1((ws.associated_mf(m[1]) || ws.associated_im_data(m[1])) || !(ws.associated_mf(m[2]) || ws.associated_im_data(m[2])),
                "gfgpu: a fem-data mu needs a fem-data lambda (fields replace the LEADING parameters)");
    out.family = GFGPU_ELASTICITY; out.params = {scalar(m[1]), scalar(m[2])}; return true;
  }
  if (std::regex_match(s, m, std::regex("\\(\\(" + ID + "\\*Div_" + v + "\\)\\*Div_Test_" + v + "\\)\\+\\(\\(\\(2\\*" + ID +
                                        "\\)\\*\\(Sym\\(Grad_" + v + "\\)\\)\\):Grad_Test_" + v + "\\)"))) {
    GMM_ASSERT1((ws.associated_mf(m[1]) || ws.associated_im_data(m[1])) || !(ws.associated_mf(m[2]) || ws.associated_im_data(m[2])),
                "gfgpu: a fem-data mu needs a fem-data lambda (fields replace the LEADING parameters)");
    out.family = GFGPU_ELASTICITY; out.params = {scalar(m[1]), scalar(m[2])}; return true;
  }
  // "lambda*Div_u*Div_Test_u + mu*(Grad_u+Grad_u'):Grad_Test_u" (the spelling of the reference's tests, tests/test_assembly.cc:812):
  // mu (grad u + grad u^T) : grad v = 2 mu eps(u) : grad v
  if (std::regex_match(s, m, std::regex("\\(\\(" + ID + "\\*Div_" + v + "\\)\\*Div_Test_" + v + "\\)\\+\\(\\(" + ID + "\\*\\(Grad_" + v +
                                        "\\+\\(Grad_" + v + "'\\)\\)\\):Grad_Test_" + v + "\\)"))) {
    GMM_ASSERT1((ws.associated_mf(m[1]) || ws.associated_im_data(m[1])) || !(ws.associated_mf(m[2]) || ws.associated_im_data(m[2])),
                "gfgpu: a fem-data mu needs a fem-data lambda (fields replace the LEADING parameters)");
    out.family = GFGPU_ELASTICITY; out.params = {scalar(m[1]), scalar(m[2])}; return true;
  }
  // the brick's form (Id+Grad_u)*<law>_PK2(Grad_u,params):Grad_Test_u, or the first variation of the law's registered
  // potential, Derivative_1_<law>_potential(Grad_u,params):Grad_Test_u -- the same linear form (dW/dF = F S)
  if (std::regex_match(s, m, std::regex("\\(\\(" + I3 + "\\+Grad_" + v + "\\)\\*" + ID + "_PK2\\(Grad_" + v + "," + ID +
                                        "\\)\\):Grad_Test_" + v)) ||
      std::regex_match(s, m, std::regex("\\(?Derivative_1_" + ID + "_potential\\(Grad_" + v + "," + ID + "\\)\\)?:Grad_Test_" + v))) {
    const std::string law = m[1];
    if (law == "Saint_Venant_Kirchhoff") out.family = GFGPU_SVK;
    else if (law == "Compressible_Neo_Hookean_Ciarlet") out.family = GFGPU_NEOHOOKEAN_CIARLET;
    else if (law == "Compressible_Neo_Hookean_Bonet") out.family = GFGPU_NEOHOOKEAN_BONET;
    else if (law == "Compressible_Mooney_Rivlin") out.family = GFGPU_MOONEY_RIVLIN;
    else if (law == "Ciarlet_Geymonat") out.family = GFGPU_CIARLET_GEYMONAT;
    else if (law == "Generalized_Blatz_Ko") out.family = GFGPU_BLATZ_KO;
    else return false;
    const std::string pn = m[2];
    const size_type np = out.family == GFGPU_BLATZ_KO ? 5 : out.family >= GFGPU_MOONEY_RIVLIN ? 3 : 2;
    GMM_ASSERT1(ws.is_constant(pn) && ws.value(pn).size() == np, "gfgpu: wrong parameters for " << law);
    out.params.assign(ws.value(pn).begin(), ws.value(pn).end());
    return true;
  }
  return false;
}

// A cached device term mirrors a getfem::mesh / mesh_fem / mesh_im.  It registers as a context dependent of the three
// (getfem_context.h): mesh::translation / transformation / refinement, set_finite_element, set_integration_method ... all
// touch() their dependents, and the destruction of any of them invalidates the context -- either way the entry is dropped at
// the next lookup, so moved nodes or a new object at a recycled address never meet stale device data.
struct context_watcher : public getfem::context_dependencies {
  mutable bool stale = false;
  void update_from_context() const override { stale = true; }
  bool still_valid() const {
    if (!is_context_valid()) return false;
    context_check();
    return !stale && is_context_valid();
  }
};

// the convexes of (mesh, mesh_fem, mesh_im) grouped by (fem, geometric transformation, integration method): an O(ne) walk
// of GetFEM's per-convex tables, done once and kept until one of the three objects says it changed
struct device_assembler::grouping {
  context_watcher watch;
  std::vector<std::vector<size_type>> groups;
};

// device side of a coupled term: one mesh, the two fems and their tables at the integration method's points, the block
struct device_assembler::rect_entry {
  context_watcher watch;
  gfgpu_mesh *mesh = nullptr;
  gfgpu_fem *fu = nullptr, *fp = nullptr;
  gfgpu_tables *tu = nullptr, *tp = nullptr;
  gfgpu_rect *rect = nullptr;
  bool assembled = false;
  size_type nu = 0, np = 0;  // BASIC dofs of the two fems (what the device block is indexed by)
  ~rect_entry() {
    gfgpu_rect_destroy(rect);
    gfgpu_tables_destroy(tu);
    gfgpu_tables_destroy(tp);
    gfgpu_fem_destroy(fu);
    gfgpu_fem_destroy(fp);
    gfgpu_mesh_destroy(mesh);
  }
};

// extension matrix of a reduced mesh_fem on the device (mesh_fem::extension_matrix(), a gmm::csr_matrix)
struct device_assembler::reduction_entry {
  context_watcher watch;
  gfgpu_reduction *E = nullptr;
  ~reduction_entry() { gfgpu_reduction_destroy(E); }
};

gfgpu_reduction *device_assembler::reduction_of(const getfem::mesh_fem &mf) {
  if (!mf.is_reduced()) return nullptr;
  std::unique_ptr<reduction_entry> &pe = reductions_[(const void *)&mf];
  if (pe && !pe->watch.still_valid()) pe.reset();
  if (pe) return pe->E;
  if (reductions_.size() > 16)
    for (auto jt = reductions_.begin(); jt != reductions_.end();) jt = (&jt->second == &pe) ? std::next(jt) : reductions_.erase(jt);
  pe.reset(new reduction_entry);
  pe->watch.add_dependency(mf);
  const size_type nb = mf.nb_basic_dof(), nr = mf.nb_dof();
  const auto &E = mf.extension_matrix();  // nb x nr, rows = basic dofs
  GMM_ASSERT1(gmm::mat_nrows(E) == nb && gmm::mat_ncols(E) == nr, "gfgpu: unexpected size of the extension matrix");
  std::vector<int64_t> rp(nb + 1, 0);
  std::vector<int32_t> col;
  std::vector<double> val;
  for (size_type j = 0; j < nb; ++j) {
    std::vector<std::pair<size_type, double>> row;
    auto r = gmm::mat_const_row(E, j);
    for (auto it = gmm::vect_const_begin(r); it != gmm::vect_const_end(r); ++it) row.emplace_back(it.index(), *it);
    std::sort(row.begin(), row.end());
    for (const auto &cv : row) { col.push_back(int32_t(cv.first)); val.push_back(cv.second); }
    rp[j + 1] = int64_t(col.size());
  }
  GFGPU_CALL(gfgpu_reduction_create(ctx_, int64_t(nb), int64_t(nr), rp.data(), col.data(), val.data(), &pe->E));
  return pe->E;
}

struct device_assembler::entry {
  context_watcher watch;
  uint64_t last_use = 0;
  gfgpu_mesh *mesh = nullptr;
  gfgpu_fem *fem = nullptr;
  gfgpu_fem *dfem = nullptr;  // data fem of the fem-data coefficients
  bool used = false;
  gfgpu_tables *tab = nullptr;
  gfgpu_term *term = nullptr;
  size_type ndof = 0;
  std::vector<size_type> imd_map;  // im_data coefficients: index in the im_data of (element of the group, Gauss point)
  size_type imd_qd = 0;
  ~entry() {
    gfgpu_term_destroy(term);
    gfgpu_tables_destroy(tab);
    gfgpu_fem_destroy(dfem);
    gfgpu_fem_destroy(fem);
    gfgpu_mesh_destroy(mesh);
  }
};

struct device_assembler::tangent_cache {
  size_type n = 0;
  gfgpu_matrix *m = nullptr;
  std::string sig;          // the terms and intervals accumulated last time
  int64_t gen = -1;         // pattern generation of the host copy below
  std::vector<int64_t> jc;
  std::vector<int32_t> ir;
  std::vector<double> pr;
  ~tangent_cache() { gfgpu_matrix_destroy(m); }
};

device_assembler::device_assembler(int device) { GFGPU_CALL(gfgpu_ctx_create(device, nullptr, &ctx_)); }
device_assembler::~device_assembler() {
  tangent_.reset();
  cache_.clear();
  rect_cache_.clear();
  reductions_.clear();
  gfgpu_ctx_destroy(ctx_);
}

// parses "FEM_PK(3,2)" / "FEM_QK(3,2)" / "GT_PK(3,1)"; returns false on anything else
static bool parse_kind(const std::string &name, const char *prefix, bool &qk, int &dim, int &deg) {
  std::smatch m;
  if (!std::regex_match(name, m, std::regex(std::string(prefix) + "_(PK|QK)\\((\\d+),(\\d+)\\)"))) return false;
  qk = m[1] == "QK";
  dim = std::stoi(m[2]);
  deg = std::stoi(m[3]);
  return true;
}

// the order-0 expression `expr` assembled by the reference's interpreter on the probe region at a PRIVATE state
static bool probe_scalar(const getfem::ga_workspace &ws, const std::string &v, const getfem::mesh_fem &mf,
                         const getfem::mesh_im &mim, const getfem::mesh_region &rg, const std::string &expr,
                         const getfem::model_real_plain_vector &state, double &E) {
  try {
    getfem::ga_workspace w2(ws, getfem::ga_workspace::inherit::ALL);
    w2.add_fem_variable(v, mf, ws.interval_of_variable(v), state);
    w2.add_expression(expr, mim, rg, 0);
    if (g_reference_assembly) g_reference_assembly(w2, 0); else w2.assembly(0);
    E = w2.assembled_potential();
    return true;
  } catch (const std::exception &) { return false; }
}

// Order-0 tree number `itree`: a law's registered potential "<law>_potential(Grad_u,params)" (getfem_nonlinear_elasticity.cc:
// 2050-2150), or a potential whose first variation -- the order-1 tree add_expression derived from it -- is a recognised
// quadratic or linear family; then P(u) = 1/2 u^T K u + c (or F.u + c) and c = P(0) is checked to vanish on the first two
// convexes with the reference's interpreter at a private zero state.
static bool recognise_potential(const getfem::ga_workspace &ws, size_type itree,
                                const std::function<bool(size_type, std::vector<recognised_term> &)> &recognise_order1,
                                recognised_term &out) {
  const getfem::ga_workspace::tree_description &td = ws.tree_info(itree);
  const std::string printed = strip(getfem::ga_tree_to_string(*td.ptree));
  std::smatch m;
  static const std::string ID = "([A-Za-z_][A-Za-z_0-9]*)";
  if (std::regex_match(printed, m, std::regex("\\(?" + ID + "_potential\\(Grad_" + ID + "," + ID + "\\)\\)?"))) {
    const std::string law = m[1], pn = m[3];
    out.varname = m[2];
    if (law == "Saint_Venant_Kirchhoff") out.family = GFGPU_SVK;
    else if (law == "Compressible_Neo_Hookean_Ciarlet") out.family = GFGPU_NEOHOOKEAN_CIARLET;
    else if (law == "Compressible_Neo_Hookean_Bonet") out.family = GFGPU_NEOHOOKEAN_BONET;
    else if (law == "Compressible_Mooney_Rivlin") out.family = GFGPU_MOONEY_RIVLIN;
    else if (law == "Ciarlet_Geymonat") out.family = GFGPU_CIARLET_GEYMONAT;
    else if (law == "Generalized_Blatz_Ko") out.family = GFGPU_BLATZ_KO;
    else return false;
    const size_type np = out.family == GFGPU_BLATZ_KO ? 5 : out.family >= GFGPU_MOONEY_RIVLIN ? 3 : 2;
    if (!ws.variable_exists(pn) || !ws.is_constant(pn) || ws.value(pn).size() != np) return false;
    out.params.assign(ws.value(pn).begin(), ws.value(pn).end());
    return true;
  }
  for (size_type j = 0; j < ws.nb_trees(); ++j) {
    const auto &t1 = ws.tree_info(j);
    if (t1.order != 1 || t1.mim != td.mim || t1.rg != td.rg) continue;
    std::vector<recognised_term> r1;
    if (!recognise_order1(j, r1) || r1.size() != 1) continue;
    const int f = r1[0].family;
    if (f != GFGPU_LAPLACE && f != GFGPU_ELASTICITY && f != GFGPU_MASS && f != GFGPU_SOURCE && f != GFGPU_NORMAL_SOURCE) continue;
    const std::string v = t1.name_test1;
    const getfem::mesh_fem *pmf = ws.associated_mf(v);
    if (!pmf || pmf->is_reduced() || !td.mim || !td.rg) continue;
    const getfem::mesh &msh = pmf->linked_mesh();
    getfem::mesh_region rg2;
    size_type nit = 0;
    for (getfem::mr_visitor it(*td.rg, msh); !it.finished() && nit < 2; ++it, ++nit) {
      if (it.f() != getfem::short_type(-1)) rg2.add(it.cv(), it.f()); else rg2.add(it.cv());
    }
    if (!nit) continue;
    getfem::model_real_plain_vector zero(pmf->nb_dof(), 0.0);
    double c = 1;
    if (!probe_scalar(ws, v, *pmf, *td.mim, rg2, getfem::ga_tree_to_string(*td.ptree), zero, c) || c != 0.0) continue;
    out = r1[0];
    return true;
  }
  return false;
}

// Device objects of a coupled div-pressure term between the vector variable `vu` and the scalar variable `vp` (same mesh,
// one classical Lagrange fem each on every convex, one approximate integration method), created once and watched through
// GetFEM's context_dependencies like every other cached entry.
device_assembler::rect_entry &device_assembler::coupled_entry(getfem::ga_workspace &ws, const getfem::mesh_im &mim,
                                                             const std::string &vu, const std::string &vp, int family,
                                                             const std::vector<int32_t> *rg_cv, const std::vector<int32_t> *rg_f) {
  const getfem::mesh_fem *pmu = ws.associated_mf(vu), *pmp = ws.associated_mf(vp);
  GMM_ASSERT1(pmu && pmp, "gfgpu: coupled terms need two fem variables");
  // (a reduced mesh_fem -- the multiplier space of a Dirichlet brick -- is mirrored on its BASIC dofs; the block is projected
  //  with the extension matrices when it is added to the tangent, workspace.cc:861-935)
  GMM_ASSERT1(family == GFGPU_RECT_MASS || (!pmu->is_reduced() && !pmp->is_reduced()),
              "gfgpu: div-pressure terms need two non-reduced fem variables");
  const getfem::mesh_fem &mfu = *pmu, &mfp = *pmp;
  const getfem::mesh &m = mfu.linked_mesh();
  GMM_ASSERT1(&mfp.linked_mesh() == &m && &mim.linked_mesh() == &m, "gfgpu: coupled variables must share the mesh");
  std::ostringstream key;
  key << &m << "/" << &mfu << "/" << &mfp << "/" << &mim << "/" << mfu.nb_basic_dof() << "/" << mfp.nb_basic_dof() << "/f" << family;
  size_type rg_faces = 0;
  if (rg_cv) {  // the region's content is part of the key (FNV-1a over the items)
    uint64_t h = 1469598103934665603ull;
    for (size_t k = 0; k < rg_cv->size(); ++k) {
      h = (h ^ uint64_t(uint32_t((*rg_cv)[k]))) * 1099511628211ull;
      h = (h ^ uint64_t(uint32_t((*rg_f)[k]))) * 1099511628211ull;
      rg_faces += (*rg_f)[k] >= 0;
    }
    key << "/rg" << rg_cv->size() << ":" << h;
  }
  std::unique_ptr<rect_entry> &pe = rect_cache_[key.str()];
  if (pe && !pe->watch.still_valid()) pe.reset();
  if (pe) return *pe;
  if (rect_cache_.size() > 8) {
    for (auto jt = rect_cache_.begin(); jt != rect_cache_.end();) jt = (&jt->second == &pe) ? std::next(jt) : rect_cache_.erase(jt);
  }
  pe.reset(new rect_entry);
  rect_entry &e = *pe;
  e.watch.add_dependency(m);
  e.watch.add_dependency(mfu);
  e.watch.add_dependency(mfp);
  e.watch.add_dependency(mim);
  const size_type ne = m.convex_index().card();
  GMM_ASSERT1(ne > 0 && ne == m.convex_index().last_true() + 1, "gfgpu: convex ids must be contiguous (call mesh::optimize_structure)");
  const size_type cv0 = 0;
  getfem::pfem pfu = mfu.fem_of_element(cv0), pfp = mfp.fem_of_element(cv0);
  bgeot::pgeometric_trans pgt = m.trans_of_convex(cv0);
  getfem::pintegration_method pim = mim.int_method_of_element(cv0);
  for (size_type cv = 0; cv < ne; ++cv)
    GMM_ASSERT1(mfu.convex_index().is_in(cv) && mfp.convex_index().is_in(cv) && mim.convex_index().is_in(cv) &&
                    mfu.fem_of_element(cv) == pfu && mfp.fem_of_element(cv) == pfp && m.trans_of_convex(cv) == pgt &&
                    mim.int_method_of_element(cv) == pim,
                "gfgpu: coupled terms are handled on uniform meshes (one fem per variable, one integration method)");
  GMM_ASSERT1(pim->type() == getfem::IM_APPROX, "gfgpu: exact integration methods are not handled");
  bool uqk, pqk, gqk;
  int udim, udeg, pdim, pdeg, gdim, gdeg;
  GMM_ASSERT1(parse_kind(getfem::name_of_fem(pfu), "FEM", uqk, udim, udeg) && parse_kind(getfem::name_of_fem(pfp), "FEM", pqk, pdim, pdeg),
              "gfgpu: fem not handled: " << getfem::name_of_fem(pfu) << " / " << getfem::name_of_fem(pfp));
  GMM_ASSERT1(parse_kind(bgeot::name_of_geometric_trans(pgt), "GT", gqk, gdim, gdeg) && gdeg == 1 && gqk == uqk && gqk == pqk,
              "gfgpu: geometric transformation not handled: " << bgeot::name_of_geometric_trans(pgt));
  const int dim = int(m.dim());
  const size_type ndu = pfu->nb_dof(cv0), ndp = pfp->nb_dof(cv0), ng = pgt->nb_points();
  getfem::papprox_integration pai = pim->approx_method();
  const size_type nq = pai->nb_points_on_convex();
  e.nu = mfu.nb_basic_dof();
  e.np = mfp.nb_basic_dof();
  const size_type npts = m.points_index().last_true() + 1;
  std::vector<double> pts(npts * dim, 0.0);
  for (dal::bv_visitor p(m.points_index()); !p.finished(); ++p)
    for (int d = 0; d < dim; ++d) pts[p * dim + d] = m.points()[p][d];
  std::vector<int32_t> conn(ne * ng);
  std::vector<int64_t> edu(ne * ndu), edp(ne * ndp);
  for (size_type cv = 0; cv < ne; ++cv) {
    for (size_type i = 0; i < ng; ++i) conn[cv * ng + i] = int32_t(m.ind_points_of_convex(cv)[i]);
    const auto &cu = mfu.ind_scalar_basic_dof_of_element(cv);
    for (size_type i = 0; i < ndu; ++i) edu[cv * ndu + i] = int64_t(cu[i]);
    const auto &cp = mfp.ind_scalar_basic_dof_of_element(cv);
    for (size_type i = 0; i < ndp; ++i) edp[cv * ndp + i] = int64_t(cp[i]);
  }
  bgeot::pstored_point_tab pspt = pai->pintegration_points();
  bgeot::pgeotrans_precomp pgp = bgeot::geotrans_precomp(pgt, pspt, 0);
  std::vector<double> w(nq), gtg(nq * ng * dim);
  for (size_type q = 0; q < nq; ++q) {
    w[q] = pai->coeff(q);
    const bgeot::base_matrix &pc = pgp->grad(q);
    for (size_type i = 0; i < ng; ++i)
      for (int d = 0; d < dim; ++d) gtg[(q * ng + i) * dim + d] = pc(i, d);
  }
  auto tables = [&](getfem::pfem pf, size_type nd, gfgpu_tables **out) {
    getfem::pfem_precomp pfp2 = getfem::fem_precomp(pf, pspt, 0);
    std::vector<double> phi(nq * nd), gphi(nq * nd * dim);
    for (size_type q = 0; q < nq; ++q) {
      const bgeot::base_tensor &bv = pfp2->val(q), &bg = pfp2->grad(q);
      for (size_type i = 0; i < nd; ++i) {
        phi[q * nd + i] = bv[i];
        for (int d = 0; d < dim; ++d) gphi[(q * nd + i) * dim + d] = bg[i + nd * d];
      }
    }
    GFGPU_CALL(gfgpu_tables_create(ctx_, dim, int(nq), int(ng), int(nd), w.data(), gtg.data(), phi.data(), gphi.data(), out));
    if (rg_faces) {  // tables at the face points, face after face (approx_integration::valid_method, getfem_integration.cc:353-368)
      const size_type nf = pgt->structure()->nb_faces(), nqf = pai->nb_points_on_face(0);
      std::vector<double> fn(nf * dim), fw(nf * nqf), fgtg(nf * nqf * ng * dim), fphi(nf * nqf * nd), fgphi(nf * nqf * nd * dim);
      for (size_type f = 0; f < nf; ++f) {
        GMM_ASSERT1(pai->nb_points_on_face(getfem::short_type(f)) == nqf,
                    "gfgpu: faces with different numbers of integration points are not handled");
        for (int d = 0; d < dim; ++d) fn[f * dim + d] = pgt->normals()[f][d];
        for (size_type q = 0; q < nqf; ++q) {
          const size_type ip = pai->ind_first_point_on_face(getfem::short_type(f)) + q, o = f * nqf + q;
          fw[o] = pai->coeff(ip);
          const bgeot::base_matrix &pc = pgp->grad(ip);
          for (size_type i = 0; i < ng; ++i)
            for (int d = 0; d < dim; ++d) fgtg[(o * ng + i) * dim + d] = pc(i, d);
          const bgeot::base_tensor &bv = pfp2->val(ip), &bg = pfp2->grad(ip);
          for (size_type i = 0; i < nd; ++i) {
            fphi[o * nd + i] = bv[i];
            for (int d = 0; d < dim; ++d) fgphi[(o * nd + i) * dim + d] = bg[i + nd * d];
          }
        }
      }
      GFGPU_CALL(gfgpu_tables_set_faces(*out, int(nf), int(nqf), fn.data(), fw.data(), fgtg.data(), fphi.data(), fgphi.data()));
    }
  };
  GFGPU_CALL(gfgpu_mesh_create(ctx_, dim, int64_t(npts), pts.data(), int64_t(ne), int(ng), conn.data(), gqk ? GFGPU_GT_QK : GFGPU_GT_PK,
                               &e.mesh));
  GFGPU_CALL(gfgpu_fem_create(ctx_, e.mesh, uqk ? GFGPU_FEM_QK : GFGPU_FEM_PK, udeg, int(mfu.get_qdim()), int(ndu), edu.data(),
                              int64_t(e.nu), &e.fu));
  GFGPU_CALL(gfgpu_fem_create(ctx_, e.mesh, pqk ? GFGPU_FEM_QK : GFGPU_FEM_PK, pdeg, int(mfp.get_qdim()), int(ndp), edp.data(),
                              int64_t(e.np), &e.fp));
  tables(pfu, ndu, &e.tu);
  tables(pfp, ndp, &e.tp);
  GFGPU_CALL(gfgpu_rect_create(ctx_, e.mesh, e.fu, e.tu, e.fp, e.tp, family, 1.0, 1.0, &e.rect));
  if (rg_cv) GFGPU_CALL(gfgpu_rect_set_region(e.rect, int64_t(rg_cv->size()), rg_cv->data(), rg_faces ? rg_f->data() : nullptr));
  return e;
}

// What a recognition result depends on, as a string: the printed tree, where it is integrated, and for every name the
// tree mentions that the workspace knows -- a variable: its mesh_fem and interval; a fixed-size constant: its values bit
// for bit; fem / im data: the object it lives on (the values travel at every call anyway).
static std::string recognition_key(const getfem::ga_workspace &ws, size_type i) {
  const auto &td = ws.tree_info(i);
  const std::string ts = getfem::ga_tree_to_string(*td.ptree);
  std::ostringstream k;
  k << ts << "|" << (const void *)td.mim << "|" << (td.rg ? long(td.rg->id()) : -2L) << "|" << td.name_test1 << "|" << td.name_test2
    << "|" << int(td.order) << "|" << int(td.operation);
  if (td.mim) k << "|dim" << int(td.mim->linked_mesh().dim()) << "|ne" << td.mim->linked_mesh().convex_index().card();
  static const std::regex ident("[A-Za-z_][A-Za-z0-9_]*");
  static const char *prefixes[] = {"", "Grad_", "Hess_", "Div_", "Test_", "Test2_", "Grad_Test_", "Grad_Test2_", "Hess_Test_",
                                   "Hess_Test2_", "Div_Test_", "Div_Test2_"};
  std::set<std::string> seen;
  for (std::sregex_iterator it(ts.begin(), ts.end(), ident), end; it != end; ++it) {
    const std::string tok = it->str();
    for (const char *pre : prefixes) {
      const size_t lp = std::strlen(pre);
      if (tok.size() <= lp || tok.compare(0, lp, pre) != 0) continue;
      const std::string name = tok.substr(lp);
      if (!ws.variable_exists(name) || !seen.insert(name).second) continue;
      k << "|" << name << ":";
      if (const getfem::mesh_fem *pmf = ws.associated_mf(name)) {
        k << "mf" << (const void *)pmf << "q" << int(pmf->get_qdim()) << "n" << pmf->nb_dof();
        if (!ws.is_constant(name)) k << "@" << ws.interval_of_variable(name).first();
      } else if (ws.associated_im_data(name)) {
        k << "imd" << (const void *)ws.associated_im_data(name);
      } else {
        char hb[40];
        for (double v : ws.value(name)) { std::snprintf(hb, sizeof hb, "%a,", v); k << hb; }
      }
    }
  }
  return k.str();
}

void device_assembler::assembly(getfem::ga_workspace &ws, size_type order) {
  GMM_ASSERT1(order <= 2, "gfgpu: assembly orders 0, 1 and 2 run on the device");
  double t0 = now_s();
  // ---- every order-1 tree must be a recognised family (the order-2 trees are their derivatives)
  std::vector<std::pair<size_type, recognised_term>> terms;
  // recognition (with its probe assemblies) runs once per tree and call, whoever asks
  // ... and, across calls, once per (printed tree, integration method, region, variables, VALUES of the constants it names):
  // a Newton or time loop re-assembles the same workspace, and a probe costs two small reference assemblies whose set-up
  // walks the whole mesh (0.4 s per call on a 660 000-element mesh before this memo).  Anything the recognised
  // parameters were read from is part of the key, so a constant that changes is recognised again.
  std::map<size_type, std::pair<bool, std::vector<recognised_term>>> memo;
  auto recognise_memo = [&](size_type i, std::vector<recognised_term> &out) {
    auto it = memo.find(i);
    if (it == memo.end()) {
      const std::string key = recognition_key(ws, i);
      auto pit = recognised_.find(key);
      if (pit == recognised_.end()) {
        std::vector<recognised_term> r;
        const bool ok = recognise_tree_sum(ws, i, r);
        if (recognised_.size() >= 64) recognised_.clear();
        pit = recognised_.emplace(key, std::make_pair(ok, std::move(r))).first;
      }
      it = memo.emplace(i, pit->second).first;
    }
    out = it->second.second;
    return it->second.first;
  };
  for (size_type i = 0; i < ws.nb_trees(); ++i) {
    const auto &td = ws.tree_info(i);
    if (order == 0) {  // only the order-0 trees contribute to assembled_potential() (workspace.cc:791-803)
      if (td.order != 0) continue;
      recognised_term rt;
      bool ok0 = recognise_potential(ws, i, recognise_memo, rt);
      if (!ok0) {  // the NVRTC route: the order-1 tree derived from this potential is a run-time compiled term that carries it
        for (size_type j = 0; j < ws.nb_trees() && !ok0; ++j) {
          const auto &t1 = ws.tree_info(j);
          std::vector<recognised_term> r1;
          if (t1.order != 1 || t1.mim != td.mim || t1.rg != td.rg || !recognise_memo(j, r1)) continue;
          if (r1.size() == 1 && r1[0].family == GFGPU_JIT && !r1[0].jit_form0.empty()) { rt = r1[0]; ok0 = true; }
        }
      }
      GMM_ASSERT1(ok0, "gfgpu: potential not handled by the device path (no CPU fallback): " << getfem::ga_tree_to_string(*td.ptree));
      terms.emplace_back(i, rt);
      continue;
    }
    if (td.order == 0) continue;
    GMM_ASSERT1(td.operation == getfem::ga_workspace::ASSEMBLY, "gfgpu: assignments are not handled");
    if (td.order == 2 && td.name_test1 != td.name_test2) {
      // a coupled tree: the block (interval of Test's variable) x (interval of Test2's variable)
      if (order != 2) continue;
      std::vector<recognised_term> rc;
      GMM_ASSERT1(recognise_memo(i, rc), "gfgpu: coupled term (" << td.name_test1 << ", " << td.name_test2
                                             << ") not handled by the device path (no CPU fallback): "
                                             << getfem::ga_tree_to_string(*td.ptree));
      for (const recognised_term &rt : rc) terms.emplace_back(i, rt);
      continue;
    }
    if (td.order == 2) {
      // the derivative of an order-1 tree of the same (mim, region, variable) -- or a bilinear form written directly
      // with Test_ / Test2_ (the asm_* wrappers)
      size_type i1 = size_type(-1);
      for (size_type j = 0; j < ws.nb_trees(); ++j) {
        const auto &t1 = ws.tree_info(j);
        if (t1.order == 1 && t1.mim == td.mim && t1.rg == td.rg && t1.name_test1 == td.name_test1) i1 = j;
      }
      if (i1 != size_type(-1)) {
        // a direct order-2 expression on the same region would have been SUMMED into this tree by add_tree: the number of
        // top-level summands must be that of the order-1 tree minus its source terms
        std::vector<recognised_term> r1;
        GMM_ASSERT1(recognise_memo(i1, r1), "gfgpu: expression not handled by the device path (no CPU fallback): "
                                                        << getfem::ga_tree_to_string(*ws.tree_info(i1).ptree));
        if (r1.size() == 1 && r1[0].by_probe) continue;  // the probe checked K against THIS tree and r = K u against the order-1 tree
        if (r1.size() == 1 && r1[0].family == GFGPU_JIT) continue;  // the JIT term carries this very tree as its second form
        size_t nsrc = 0;
        for (const recognised_term &rt : r1)  // (a coupled residual part has no derivative with respect to its own test variable)
          nsrc += rt.family == GFGPU_SOURCE || rt.family == GFGPU_NORMAL_SOURCE || rt.family == GFGPU_SHIM_COUPLED_DIV;
        const size_t n1 = count_top_level_summands(strip(getfem::ga_tree_to_string(*ws.tree_info(i1).ptree)));
        const size_t n2 = count_top_level_summands(strip(getfem::ga_tree_to_string(*td.ptree)));
        GMM_ASSERT1(n2 + nsrc == n1, "gfgpu: a tangent tree that mixes derived and directly written terms is not handled: "
                                         << getfem::ga_tree_to_string(*td.ptree));
        continue;
      }
      if (order != 2) continue;  // a directly written bilinear form contributes to the matrix only
      std::vector<recognised_term> r2;
      GMM_ASSERT1(recognise_memo(i, r2), "gfgpu: expression not handled by the device path (no CPU fallback): "
                                                     << getfem::ga_tree_to_string(*td.ptree));
      for (const recognised_term &rt : r2) terms.emplace_back(i, rt);
      continue;
    }
    std::vector<recognised_term> rts;
    GMM_ASSERT1(recognise_memo(i, rts), "gfgpu: expression not handled by the device path (no CPU fallback): "
                                            << getfem::ga_tree_to_string(*td.ptree));
    bool has_derivative = false;  // only the order-2 trees that exist are assembled (workspace.cc:791-936)
    for (size_type j = 0; j < ws.nb_trees() && !has_derivative; ++j) {
      const auto &t2 = ws.tree_info(j);
      has_derivative = t2.order == 2 && t2.mim == td.mim && t2.rg == td.rg && t2.name_test1 == td.name_test1;
    }
    for (recognised_term &rt : rts) {
      // the coupled part of a residual has its tangent in the order-2 trees of the OTHER variable pair (handled above)
      if (rt.family == GFGPU_SHIM_COUPLED_DIV && order == 2) continue;
      rt.no_tangent = order == 2 && !has_derivative;
      terms.emplace_back(i, rt);
    }
  }
  t_extract = t_device = t_fill = 0;
  static const bool trace_t = std::getenv("GFGPU_TRACE_TIMING") != nullptr;
  auto mark = [&](const char *what) {
    if (trace_t) std::fprintf(stderr, "[gfgpu timing] order %d %-28s %.4f s since the call began\n", int(order), what, now_s() - t0);
  };
  mark("trees recognised");
  const size_type nprim = ws.nb_primary_dof() ? ws.nb_primary_dof() : 0;
  if (terms.empty() && order == 0) return;  // nothing adds to the potential
  if (terms.empty()) {
    // no tree of this order (assembly(1) of a workspace that only holds directly written bilinear forms, say): the
    // reference runs an empty instruction list and only sizes its result (workspace.cc:805-826)
    if (order == 1) {
      getfem::base_vector &V = ws.assembled_vector();
      if (V.size() < nprim) V.resize(nprim, 0.0);
    } else {
      getfem::model_real_sparse_matrix &K = ws.assembled_matrix();
      if (gmm::mat_nrows(K) < nprim || gmm::mat_ncols(K) < nprim) gmm::resize(K, nprim, nprim);
    }
    return;
  }
  // order 2: every tree adds into ONE tangent (workspace.cc:791-936) -- accumulated on the device at the variables'
  // intervals (gfgpu_matrix_*), downloaded once
  size_type need_all = nprim;
  for (auto &it : terms) {
    for (const std::string &vn : {it.second.varname, it.second.varname_u, it.second.varname_p}) {
      if (vn.empty()) continue;
      const getfem::mesh_fem *pmf = ws.associated_mf(vn);
      GMM_ASSERT1(pmf, "gfgpu: the variable must be a fem variable");
      need_all = std::max<size_type>(need_all, ws.interval_of_variable(vn).first() + pmf->nb_dof());
    }
  }
  // The workspace tangent lives on the device across calls (a Newton loop, a time loop): same size -> the matrix is
  // zeroed with its pattern kept, the terms are added again, and as long as gfgpu_matrix_pattern_generation does not move
  // only the VALUES come back to the host (jc / ir, 8.8 GB for BASELINE config 3, are downloaded once).
  struct { gfgpu_matrix *m = nullptr; } dK;
  std::string terms_sig;
  struct pending_add {
    gfgpu_term *term; double alpha; int64_t off; gfgpu_rect *rect = nullptr; int transposed = 0; int64_t coff = 0;
    gfgpu_reduction *Er = nullptr, *Ec = nullptr;  // reduced mesh_fems: the block is projected, E_rows^T S E_cols
  };
  std::vector<pending_add> pending_adds;
  if (order == 2) {
    if (tangent_ && tangent_->n != need_all) tangent_.reset();
    if (!tangent_) {
      tangent_.reset(new tangent_cache);
      tangent_->n = need_all;
      GFGPU_CALL(gfgpu_matrix_create(ctx_, int64_t(need_all), int64_t(need_all), &tangent_->m));
    }
    dK.m = tangent_->m;
  }
  size_type n_added = 0;  // tangents accumulated into dK
  // integration region in mr_visitor order (the order ga_exec walks it, C&E.cc:8789): all convexes (returns true), a set of
  // convexes, or a set of faces
  auto walk_region = [&](const getfem::ga_workspace::tree_description &td, const getfem::mesh &m, std::vector<int32_t> &rg_cv,
                         std::vector<int32_t> &rg_f, size_type &rg_faces) -> bool {
    GMM_ASSERT1(td.rg, "gfgpu: no region");
    const bool all_cv = td.rg->id() == getfem::mesh_region::all_convexes().id();
    rg_cv.clear(); rg_f.clear();
    rg_faces = 0;
    if (!all_cv) {
      // Inside GETFEM_OMP_PARALLEL with several partitions mr_visitor walks only the calling thread's slice
      // (getfem_mesh_region.cc:186-200, 503-540), and a copy of a mesh-owned region SHARES its implementation (operator=,
      // getfem_mesh_region.cc:100-106), partition caches included.  The device assembles the WHOLE region in one call (made by
      // the thread of partition 0 alone, see the dispatch patch): there the item list is rebuilt from the membership test,
      // which ignores the partition -- ascending convexes, whole convex first, then faces: the visitor's order.
      if (getfem::me_is_multithreaded_now() && getfem::partition_master::get().get_nb_partitions() > 1) {
        for (dal::bv_visitor cv(m.convex_index()); !cv.finished(); ++cv) {
          if (td.rg->is_in(cv, getfem::short_type(-1), m)) { rg_cv.push_back(int32_t(cv)); rg_f.push_back(-1); }
          const getfem::short_type nf = m.structure_of_convex(cv)->nb_faces();
          for (getfem::short_type f = 0; f < nf; ++f)
            if (td.rg->is_in(cv, f, m)) { rg_cv.push_back(int32_t(cv)); rg_f.push_back(int32_t(f)); ++rg_faces; }
        }
      } else {
        for (getfem::mr_visitor v(*td.rg, m); !v.finished(); ++v) {
          rg_cv.push_back(int32_t(v.cv()));
          const bool isf = v.f() != getfem::short_type(-1);
          rg_f.push_back(isf ? int32_t(v.f()) : -1);
          rg_faces += isf;
        }
      }
      GMM_ASSERT1(rg_faces == 0 || rg_faces == rg_cv.size(), "gfgpu: a region must hold either convexes or faces");
    }
    return all_cv;
  };
  for (auto &it : terms) {
    const auto &td = ws.tree_info(it.first);
    const recognised_term &rt = it.second;
    if (rt.family == GFGPU_SHIM_COUPLED_MASS) {
      // "Test_a:Test2_b" on two fems (asm_mass_matrix(M, mim, mf1, mf2, rg); the constraint matrix of the Dirichlet bricks with
      // multipliers, getfem_models.cc:4386-4421): a directly written bilinear form, so order 2 only; rows = a, columns = b
      if (order != 2) continue;
      GMM_ASSERT1(!(getfem::me_is_multithreaded_now() && getfem::partition_master::get().get_nb_partitions() > 1) ||
                      getfem::partition_master::get().get_current_partition() == 0, "gfgpu: internal error (partition)");
      const getfem::mesh_fem &mfa = *ws.associated_mf(rt.varname_u), &mfb = *ws.associated_mf(rt.varname_p);
      std::vector<int32_t> rg_cv, rg_f;
      size_type rg_faces = 0;
      const bool all_cv = walk_region(td, mfa.linked_mesh(), rg_cv, rg_f, rg_faces);
      if (!all_cv && rg_cv.empty()) continue;
      rect_entry &re = coupled_entry(ws, *td.mim, rt.varname_u, rt.varname_p, GFGPU_RECT_MASS, all_cv ? nullptr : &rg_cv,
                                     all_cv ? nullptr : &rg_f);
      const gmm::sub_interval &Ia = ws.interval_of_variable(rt.varname_u), &Ib = ws.interval_of_variable(rt.varname_p);
      double t1 = now_s();
      t_extract += t1 - t0;
      if (!re.assembled) {
        GFGPU_CALL(gfgpu_rect_assemble_dev(re.rect));
        re.assembled = true;
      }
      pending_add pa{nullptr, rt.sign * ws.factor_of_variable(rt.varname_u) * ws.factor_of_variable(rt.varname_p), int64_t(Ia.first())};
      pa.rect = re.rect; pa.transposed = 0; pa.coff = int64_t(Ib.first());
      pa.Er = reduction_of(mfa); pa.Ec = reduction_of(mfb);
      std::ostringstream sg;
      sg << "rectM" << (const void *)re.rect << "@" << pa.off << "," << pa.coff << "/" << (const void *)pa.Er << "/" << (const void *)pa.Ec << ";";
      terms_sig += sg.str();
      pending_adds.push_back(pa);
      ++n_added;
      t_device += now_s() - t1;
      t0 = now_s();
      continue;
    }
    if (rt.family == GFGPU_SHIM_COUPLED_DIV) {
      GMM_ASSERT1(td.rg && td.rg->id() == getfem::mesh_region::all_convexes().id(),
                  "gfgpu: coupled terms are handled on the whole mesh (no region)");
      GMM_ASSERT1(!(getfem::me_is_multithreaded_now() && getfem::partition_master::get().get_nb_partitions() > 1) ||
                      getfem::partition_master::get().get_current_partition() == 0, "gfgpu: internal error (partition)");
      rect_entry &re = coupled_entry(ws, *td.mim, rt.varname_u, rt.varname_p);
      const gmm::sub_interval &Iu = ws.interval_of_variable(rt.varname_u), &Ip = ws.interval_of_variable(rt.varname_p);
      double t1 = now_s();
      t_extract += t1 - t0;
      if (!re.assembled) {  // a constant-coefficient linear block: its values never change
        GFGPU_CALL(gfgpu_rect_assemble_dev(re.rect));
        re.assembled = true;
      }
      if (order == 2) {
        // rows = Test's variable, columns = Test2's: (u, p) is the block, (p, u) its transpose
        const int64_t roff = int64_t(rt.transposed ? Ip.first() : Iu.first()), coff = int64_t(rt.transposed ? Iu.first() : Ip.first());
        pending_add pa{nullptr, rt.sign * ws.factor_of_variable(rt.varname_u) * ws.factor_of_variable(rt.varname_p), roff};
        pa.rect = re.rect; pa.transposed = rt.transposed ? 1 : 0; pa.coff = coff;
        std::ostringstream sg;
        sg << "rect" << (const void *)re.rect << (rt.transposed ? "T" : "N") << "@" << roff << "," << coff << "*" << rt.sign << ";";
        terms_sig += sg.str();
        pending_adds.push_back(pa);
        ++n_added;
        t_device += now_s() - t1;
      } else if (order == 1) {
        // residual parts: sign * B p into u's interval, sign * B^T u into p's
        const getfem::model_real_plain_vector &X = ws.value(rt.transposed ? rt.varname_u : rt.varname_p);
        const size_type nout = rt.transposed ? re.np : re.nu;
        GMM_ASSERT1(X.size() == (rt.transposed ? re.nu : re.np), "gfgpu: bad size of a coupled variable's value vector");
        std::vector<double> y(nout, 0.0);
        GFGPU_CALL(gfgpu_rect_mult_host(re.rect, rt.transposed ? 1 : 0, rt.sign, X.data(), 0.0, y.data()));
        double t2 = now_s();
        t_device += t2 - t1;
        const size_type off = rt.transposed ? Ip.first() : Iu.first();
        getfem::base_vector &V = ws.assembled_vector();
        if (V.size() < off + nout) V.resize(std::max<size_type>(nprim, off + nout), 0.0);
        for (size_type d = 0; d < nout; ++d) V[off + d] += y[d];
        t_fill += now_s() - t2;
      }
      t0 = now_s();
      continue;
    }
    const getfem::mesh_fem *pmf = ws.associated_mf(rt.varname);
    GMM_ASSERT1(pmf, "gfgpu: the variable must be a fem variable");
    const getfem::mesh_fem &mf = *pmf;
    // A REDUCED mesh_fem (partial_mesh_fem: multiplier spaces; periodic / enriched spaces) is assembled on its basic dofs and
    // projected with its extension matrix: K(I, I) += E^T K_basic E, V(I) += E^T V_basic, state U_basic = E U
    // (workspace.cc:861-935).  `ndof` below counts the BASIC dofs (what the device term is indexed by), `nred` the variable's.
    gfgpu_reduction *Ered = reduction_of(mf);
    const getfem::mesh_im &mim = *td.mim;
    const getfem::mesh &m = mf.linked_mesh();
    // integration region in mr_visitor order (the order ga_exec walks it, C&E.cc:8789): all convexes, a set of
    // convexes, or a set of faces
    std::vector<int32_t> rg_cv, rg_f;
    size_type rg_faces = 0;
    const bool all_cv = walk_region(td, m, rg_cv, rg_f, rg_faces);
    if (!all_cv && rg_cv.empty()) continue;  // an empty region assembles nothing: ga_exec walks zero elements (C&E.cc:8789-8866)
    mark("region walked");
    const gmm::sub_interval &I = ws.interval_of_variable(rt.varname);
    const size_type nred = mf.nb_dof();  // triggers enumerate_dof
    const size_type ndof = mf.nb_basic_dof();
    mark("nb_dof");
    GMM_ASSERT1(m.convex_index().card() > 0 && m.convex_index().card() == m.convex_index().last_true() + 1,
                "gfgpu: convex ids must be contiguous (call mesh::optimize_structure)");
    const size_type ne_mesh = m.convex_index().card();
    // Non-uniform meshes (C&E.cc:5902-5936, the is_uniform() == false path): the convexes are grouped by (fem, geometric
    // transformation, integration method); every group is a device term of its own -- own compact mesh, dof rows and
    // tables -- and the groups' tangents meet in the workspace matrix (gfgpu_matrix_add_term: union pattern, summed values),
    // their residuals in V.  A uniform mesh is one group and takes exactly the path it always took.
    std::ostringstream gkey;
    gkey << &m << "/" << &mf << "/" << &mim;
    std::unique_ptr<grouping> &pg = groupings_[gkey.str()];
    if (pg && !pg->watch.still_valid()) pg.reset();
    if (!pg) {
      if (groupings_.size() > 16) {  // bounded: drop everything but the slot being filled
        for (auto jt = groupings_.begin(); jt != groupings_.end();) jt = (&jt->second == &pg) ? std::next(jt) : groupings_.erase(jt);
      }
      pg.reset(new grouping);
      pg->watch.add_dependency(m);
      pg->watch.add_dependency(mf);
      pg->watch.add_dependency(mim);
      std::map<std::tuple<const void *, const void *, const void *>, size_t> gid;
      for (size_type cv = 0; cv < ne_mesh; ++cv) {
        if (!mf.convex_index().is_in(cv) || !mim.convex_index().is_in(cv)) continue;  // no fem / no im: not assembled
        const auto k = std::make_tuple((const void *)mf.fem_of_element(cv).get(), (const void *)m.trans_of_convex(cv).get(),
                                       (const void *)mim.int_method_of_element(cv).get());
        auto it = gid.find(k);
        if (it == gid.end()) { it = gid.emplace(k, pg->groups.size()).first; pg->groups.emplace_back(); }
        pg->groups[it->second].push_back(cv);
      }
    }
    mark("convex groups");
    const std::vector<std::vector<size_type>> &groups = pg->groups;
    std::vector<int32_t> local_of(groups.size() == 1 && groups[0].size() == ne_mesh ? 0 : ne_mesh, -1);
    for (size_t ig = 0; ig < groups.size(); ++ig) {
    const std::vector<size_type> &gcv = groups[ig];
    const bool whole_mesh = groups.size() == 1 && gcv.size() == ne_mesh;
    if (!whole_mesh) for (size_t k = 0; k < gcv.size(); ++k) local_of[gcv[k]] = int32_t(k);
    // the region restricted to this group, in the group's local numbering (visitor order kept)
    std::vector<int32_t> grg_cv, grg_f;
    if (!all_cv) {
      for (size_t k = 0; k < rg_cv.size(); ++k) {
        const size_type cv = size_type(rg_cv[k]);
        if (whole_mesh) {
          if (cv < ne_mesh) { grg_cv.push_back(int32_t(cv)); grg_f.push_back(rg_f[k]); }
        } else if (cv < ne_mesh && local_of[cv] >= 0 && std::binary_search(gcv.begin(), gcv.end(), cv)) {
          grg_cv.push_back(local_of[cv]);
          grg_f.push_back(rg_f[k]);
        }
      }
      if (grg_cv.empty()) continue;
    }
    const bool use_region = !all_cv;
    const size_type ne = gcv.size(), cv0 = gcv[0];
    // within a group: one classical Lagrange fem / degree-1 geometric transformation / approximate im
    getfem::pfem pf = mf.fem_of_element(cv0);
    bgeot::pgeometric_trans pgt = m.trans_of_convex(cv0);
    getfem::pintegration_method pim = mim.int_method_of_element(cv0);
    GMM_ASSERT1(pim->type() == getfem::IM_APPROX, "gfgpu: exact integration methods are not handled");
    bool fqk, gqk;
    int fdim, fdeg, gdim, gdeg;
    GMM_ASSERT1(parse_kind(getfem::name_of_fem(pf), "FEM", fqk, fdim, fdeg),
                "gfgpu: fem not handled: " << getfem::name_of_fem(pf));
    GMM_ASSERT1(parse_kind(bgeot::name_of_geometric_trans(pgt), "GT", gqk, gdim, gdeg) && gdeg == 1 && gqk == fqk,
                "gfgpu: geometric transformation not handled: " << bgeot::name_of_geometric_trans(pgt));
    const int dim = int(m.dim()), Q = int(mf.get_qdim());
    const size_type nd = pf->nb_dof(cv0), ng = pgt->nb_points();
    getfem::papprox_integration pai = pim->approx_method();
    const size_type nq = pai->nb_points_on_convex();

    std::ostringstream key;
    key << &m << "/" << &mf << "/" << &mim << "/" << rt.family << "/" << ne << "/" << ndof << "/" << fdeg << "/"
        << getfem::name_of_int_method(pim) << "/g" << ig << "of" << groups.size() << "@" << cv0;
    {  // parameters bit for bit (a load that changes by 1e-9 is another term), and the family's own scale stays 1:
       // factor_of_variable enters at gfgpu_matrix_add_term for order 2 only, like the reference (C&E.cc:5359-5418 vs 4669-4735)
      char hb[40];
      if (rt.family == GFGPU_JIT) key << "/jit:" << rt.jit_form0 << "|" << rt.jit_form1 << "|" << rt.jit_form2;  // parameters are refreshed at every call
      else for (double p : rt.params) { std::snprintf(hb, sizeof hb, "/%a", p); key << hb; }
    }
    for (const std::string &fn : rt.field_names)
      key << "/field:" << fn << "@" << (const void *)ws.associated_mf(fn) << "/" << (const void *)ws.associated_im_data(fn);
    if (use_region) {  // the region's content is part of the key (FNV-1a over the items)
      uint64_t h = 1469598103934665603ull;
      for (size_t k = 0; k < grg_cv.size(); ++k) {
        h = (h ^ uint64_t(uint32_t(grg_cv[k]))) * 1099511628211ull;
        h = (h ^ uint64_t(uint32_t(grg_f[k]))) * 1099511628211ull;
      }
      key << "/rg" << grg_cv.size() << ":" << h;
    }
    if (!whole_mesh) {  // ... and so is the group's convex list
      uint64_t h = 1469598103934665603ull;
      for (size_type cv : gcv) h = (h ^ uint64_t(cv)) * 1099511628211ull;
      key << "/cv" << h;
    }
    {  // drop what the context says is stale, and the least recently used entries beyond the cache bound
      auto it = cache_.find(key.str());
      if (it != cache_.end() && !it->second->watch.still_valid()) cache_.erase(it);
      const size_t cache_max = 12;
      while (cache_.size() >= cache_max && cache_.find(key.str()) == cache_.end()) {
        auto lru = cache_.begin();
        for (auto jt = cache_.begin(); jt != cache_.end(); ++jt)
          if (jt->second->last_use < lru->second->last_use) lru = jt;
        cache_.erase(lru);
      }
    }
    mark("cache looked up");
    std::unique_ptr<entry> &pe = cache_[key.str()];
    if (!pe) {
      pe.reset(new entry);
      entry &e = *pe;
      e.ndof = ndof;
      e.watch.add_dependency(m);
      e.watch.add_dependency(mf);
      e.watch.add_dependency(mim);
      // mesh (basic_mesh::points_of_convex / ind_points_of_convex)
      const size_type npts = m.points_index().last_true() + 1;
      std::vector<double> pts(npts * dim, 0.0);
      for (dal::bv_visitor p(m.points_index()); !p.finished(); ++p)
        for (int d = 0; d < dim; ++d) pts[p * dim + d] = m.points()[p][d];
      std::vector<int32_t> conn(ne * ng);
      std::vector<int64_t> edof(ne * nd);
      for (size_type k = 0; k < ne; ++k) {  // the group's convexes, in ascending order
        const size_type cv = gcv[k];
        for (size_type i = 0; i < ng; ++i) conn[k * ng + i] = int32_t(m.ind_points_of_convex(cv)[i]);
        const auto &ct = mf.ind_scalar_basic_dof_of_element(cv);
        for (size_type i = 0; i < nd; ++i) edof[k * nd + i] = int64_t(ct[i]);
      }
      // reference tables at the volume quadrature points (geotrans_precomp_ / fem_precomp_)
      bgeot::pstored_point_tab pspt = pai->pintegration_points();
      getfem::pfem_precomp pfp = getfem::fem_precomp(pf, pspt, 0);
      bgeot::pgeotrans_precomp pgp = bgeot::geotrans_precomp(pgt, pspt, 0);
      std::vector<double> w(nq), gtg(nq * ng * dim), phi(nq * nd), gphi(nq * nd * dim);
      for (size_type q = 0; q < nq; ++q) {
        w[q] = pai->coeff(q);
        const bgeot::base_matrix &pc = pgp->grad(q);
        for (size_type i = 0; i < ng; ++i)
          for (int d = 0; d < dim; ++d) gtg[(q * ng + i) * dim + d] = pc(i, d);
        const bgeot::base_tensor &bv = pfp->val(q), &bg = pfp->grad(q);
        for (size_type i = 0; i < nd; ++i) {
          phi[q * nd + i] = bv[i];
          for (int d = 0; d < dim; ++d) gphi[(q * nd + i) * dim + d] = bg[i + nd * d];
        }
      }
      GFGPU_CALL(gfgpu_mesh_create(ctx_, dim, int64_t(npts), pts.data(), int64_t(ne), int(ng), conn.data(),
                                   gqk ? GFGPU_GT_QK : GFGPU_GT_PK, &e.mesh));
      GFGPU_CALL(gfgpu_fem_create(ctx_, e.mesh, fqk ? GFGPU_FEM_QK : GFGPU_FEM_PK, fdeg, Q, int(nd), edof.data(),
                                  int64_t(ndof), &e.fem));
      GFGPU_CALL(gfgpu_tables_create(ctx_, dim, int(nq), int(ng), int(nd), w.data(), gtg.data(), phi.data(), gphi.data(),
                                     &e.tab));
      // shape values of the geometric transformation at the Gauss points: the position X of run-time compiled integrands
      std::vector<double> gtv(nq * ng);
      for (size_type q = 0; q < nq; ++q)
        for (size_type i = 0; i < ng; ++i) gtv[q * ng + i] = pgp->val(q)[i];
      GFGPU_CALL(gfgpu_tables_set_gt_values(e.tab, gtv.data(), nullptr));
      if (rg_faces) {
        // tables at the face points (they follow the volume points in the method's point table, face after face:
        // approx_integration::valid_method, getfem_integration.cc:353-368) and the reference normals
        const size_type nf = pgt->structure()->nb_faces(), nqf = pai->nb_points_on_face(0);
        std::vector<double> fn(nf * dim), fw(nf * nqf), fgtg(nf * nqf * ng * dim), fphi(nf * nqf * nd), fgphi(nf * nqf * nd * dim);
        for (size_type f = 0; f < nf; ++f) {
          GMM_ASSERT1(pai->nb_points_on_face(getfem::short_type(f)) == nqf,
                      "gfgpu: faces with different numbers of integration points are not handled");
          for (int d = 0; d < dim; ++d) fn[f * dim + d] = pgt->normals()[f][d];
          for (size_type q = 0; q < nqf; ++q) {
            const size_type ip = pai->ind_first_point_on_face(getfem::short_type(f)) + q, o = f * nqf + q;
            fw[o] = pai->coeff(ip);
            const bgeot::base_matrix &pc = pgp->grad(ip);
            for (size_type i = 0; i < ng; ++i)
              for (int d = 0; d < dim; ++d) fgtg[(o * ng + i) * dim + d] = pc(i, d);
            const bgeot::base_tensor &bv = pfp->val(ip), &bg = pfp->grad(ip);
            for (size_type i = 0; i < nd; ++i) {
              fphi[o * nd + i] = bv[i];
              for (int d = 0; d < dim; ++d) fgphi[(o * nd + i) * dim + d] = bg[i + nd * d];
            }
          }
        }
        GFGPU_CALL(gfgpu_tables_set_faces(e.tab, int(nf), int(nqf), fn.data(), fw.data(), fgtg.data(), fphi.data(),
                                          fgphi.data()));
        std::vector<double> fgtv(nf * nqf * ng);
        for (size_type f = 0; f < nf; ++f)
          for (size_type q = 0; q < nqf; ++q) {
            const size_type ip = pai->ind_first_point_on_face(getfem::short_type(f)) + q;
            for (size_type i = 0; i < ng; ++i) fgtv[((f * nqf + q) * ng) + i] = pgp->val(ip)[i];
          }
        GFGPU_CALL(gfgpu_tables_set_gt_values(e.tab, gtv.data(), fgtv.data()));
      }
      if (rt.family == GFGPU_JIT) {
        const bool vdep = rt.jit_form2.find("u") != std::string::npos;  // "u" or "gu" in the tangent: its pattern may move
        GFGPU_CALL(gfgpu_term_create_jit(ctx_, e.mesh, e.fem, e.tab, rt.jit_form1.c_str(), rt.jit_form2.c_str(), rt.params.data(),
                                         int(rt.params.size()), 1.0, vdep ? 1 : 0, &e.term));
        if (!rt.jit_form0.empty()) GFGPU_CALL(gfgpu_term_set_jit_potential(e.term, rt.jit_form0.c_str()));
      } else {
        GFGPU_CALL(gfgpu_term_create(ctx_, e.mesh, e.fem, e.tab, rt.family, rt.params.data(), int(rt.params.size()), 1.0,
                                     GFGPU_STRATEGY_AUTO, &e.term));
      }
      if (use_region)
        GFGPU_CALL(gfgpu_term_set_region(e.term, int64_t(grg_cv.size()), grg_cv.data(), rg_faces ? grg_f.data() : nullptr));
      if (!rt.field_names.empty() && ws.associated_im_data(rt.field_names[0])) {
        // im data (ga_workspace::add_im_data): one value per Gauss point, read by ga_exec as U[filtered_index_of_point(cv, q)]
        // (C&E.cc:134-165).  On the device it is a field on a synthetic discontinuous "fem" with one dof per Gauss point and the
        // identity as basis table: the kernels' field evaluation sum_i vals[dof_i] phi_i(q) then returns the stored value.
        const getfem::im_data *pid = ws.associated_im_data(rt.field_names[0]);
        for (const std::string &fn : rt.field_names)
          GMM_ASSERT1(ws.associated_im_data(fn) == pid, "gfgpu: the im_data coefficients of one term must share their im_data object");
        GMM_ASSERT1(&pid->linked_mesh_im() == &mim, "gfgpu: im data have to be used on their original integration method");
        GMM_ASSERT1(!rg_faces, "gfgpu: im_data coefficients are handled in volume terms");
        const size_type qd = pid->nb_tensor_elem();
        GMM_ASSERT1(qd >= 1 && qd <= 3, "gfgpu: scalar or vector im_data only");
        std::vector<int64_t> ded(ne * nq);
        e.imd_map.resize(ne * nq);
        for (size_type k = 0; k < ne; ++k)
          for (size_type q = 0; q < nq; ++q) {
            ded[k * nq + q] = int64_t((k * nq + q) * qd);
            const size_type ip = pid->filtered_index_of_point(gcv[k], q);
            GMM_ASSERT1(ip != size_type(-1), "gfgpu: im data with no data on an integration point of the region");
            e.imd_map[k * nq + q] = ip;
          }
        e.imd_qd = qd;
        GFGPU_CALL(gfgpu_fem_create(ctx_, e.mesh, fqk ? GFGPU_FEM_QK : GFGPU_FEM_PK, 0, int(qd), int(nq), ded.data(),
                                    int64_t(ne * nq * qd), &e.dfem));
        std::vector<double> ident(nq * nq, 0.0);
        for (size_type q = 0; q < nq; ++q) ident[q * nq + q] = 1.0;
        std::vector<std::vector<double>> vals;
        for (const std::string &fn : rt.field_names) {
          const getfem::model_real_plain_vector &src = ws.value(fn);
          vals.emplace_back(ne * nq * qd);
          for (size_type k = 0; k < ne * nq; ++k)
            for (size_type c = 0; c < qd; ++c) vals.back()[k * qd + c] = rt.field_sign * src[e.imd_map[k] * qd + c];
        }
        GFGPU_CALL(gfgpu_term_set_fields(e.term, int(vals.size()), e.dfem, ident.data(), nullptr, vals[0].data(),
                                         vals.size() > 1 ? vals[1].data() : nullptr));
      } else if (!rt.field_names.empty()) {
        // fem-data coefficients: the data mesh_fem's dof table and its basis at the same points (fem_precomp_::val)
        const getfem::mesh_fem *pmd = ws.associated_mf(rt.field_names[0]);
        for (const std::string &fn : rt.field_names)
          GMM_ASSERT1(ws.associated_mf(fn) == pmd, "gfgpu: the fem-data coefficients of one term must share their mesh_fem");
        GMM_ASSERT1(&pmd->linked_mesh() == &m && !pmd->is_reduced(), "gfgpu: the data mesh_fem must be a non-reduced fem of the same mesh");
        getfem::pfem pfd = pmd->fem_of_element(cv0);
        for (size_type cv : gcv)
          GMM_ASSERT1(pmd->fem_of_element(cv) == pfd, "gfgpu: the data fem must be uniform on a group of like convexes");
        bool dqk; int ddim, ddeg;
        GMM_ASSERT1(parse_kind(getfem::name_of_fem(pfd), "FEM", dqk, ddim, ddeg) && dqk == fqk,
                    "gfgpu: data fem not handled: " << getfem::name_of_fem(pfd));
        const size_type ndd = pfd->nb_dof(cv0);
        std::vector<int64_t> ded(ne * ndd);
        for (size_type k = 0; k < ne; ++k) {
          const auto &ct = pmd->ind_scalar_basic_dof_of_element(gcv[k]);
          for (size_type i = 0; i < ndd; ++i) ded[k * ndd + i] = int64_t(ct[i]);
        }
        GFGPU_CALL(gfgpu_fem_create(ctx_, e.mesh, dqk ? GFGPU_FEM_QK : GFGPU_FEM_PK, ddeg, int(pmd->get_qdim()), int(ndd),
                                    ded.data(), int64_t(pmd->nb_dof()), &e.dfem));
        getfem::pfem_precomp pfpd = getfem::fem_precomp(pfd, pspt, 0);
        std::vector<double> dphi(nq * ndd), dfphi;
        for (size_type q = 0; q < nq; ++q)
          for (size_type i = 0; i < ndd; ++i) dphi[q * ndd + i] = pfpd->val(q)[i];
        if (rg_faces) {
          const size_type nf = pgt->structure()->nb_faces(), nqf = pai->nb_points_on_face(0);
          dfphi.resize(nf * nqf * ndd);
          for (size_type f = 0; f < nf; ++f)
            for (size_type q = 0; q < nqf; ++q)
              for (size_type i = 0; i < ndd; ++i)
                dfphi[((f * nqf) + q) * ndd + i] = pfpd->val(pai->ind_first_point_on_face(getfem::short_type(f)) + q)[i];
        }
        std::vector<std::vector<double>> vals;
        for (const std::string &fn : rt.field_names) {
          vals.emplace_back(ws.value(fn).begin(), ws.value(fn).end());
          for (double &x : vals.back()) x *= rt.field_sign;
        }
        GFGPU_CALL(gfgpu_term_set_fields(e.term, int(vals.size()), e.dfem, dphi.data(), rg_faces ? dfphi.data() : nullptr,
                                         vals[0].data(), vals.size() > 1 ? vals[1].data() : nullptr));
      }
    }
    entry &e = *pe;
    if (!rt.field_names.empty() && e.used) {  // cached device term: the data may have changed since the last assembly
      for (size_t k = 0; k < rt.field_names.size(); ++k) {
        std::vector<double> vals;
        const getfem::model_real_plain_vector &src = ws.value(rt.field_names[k]);
        if (!e.imd_map.empty()) {  // im data: into the device order (element of the group, Gauss point)
          vals.resize(e.imd_map.size() * e.imd_qd);
          for (size_t j = 0; j < e.imd_map.size(); ++j)
            for (size_type c = 0; c < e.imd_qd; ++c) vals[j * e.imd_qd + c] = rt.field_sign * src[e.imd_map[j] * e.imd_qd + c];
        } else {
          vals.assign(src.begin(), src.end());
          for (double &x : vals) x *= rt.field_sign;
        }
        GFGPU_CALL(gfgpu_term_update_field(e.term, int(k), vals.data()));
      }
    }
    if (rt.family == GFGPU_JIT && !rt.params.empty())  // the constants of the expression as they are NOW
      GFGPU_CALL(gfgpu_term_set_params(e.term, rt.params.data(), int(rt.params.size())));
    e.used = true;
    e.last_use = ++use_clock_;
    // the variable's values, in the fem's own numbering (the workspace interval only offsets the result)
    const getfem::model_real_plain_vector &Ured = ws.value(rt.varname);
    GMM_ASSERT1(Ured.size() == nred, "gfgpu: bad size of the variable's value vector");
    std::vector<double> Ubasic;
    if (Ered) {
      Ubasic.resize(ndof);
      GFGPU_CALL(gfgpu_reduction_extend_host(Ered, Ured.data(), Ubasic.data()));
    }
    const double *Udata = Ered ? Ubasic.data() : Ured.data();
    struct { const double *p; const double *data() const { return p; } } U{Udata};
    mark("entry ready");
    double t1 = now_s();
    t_extract += t1 - t0;

    if (order == 0) {
      double E = 0;
      GFGPU_CALL(gfgpu_term_potential_host(e.term, U.data(), &E));
      ws.assembled_potential() += E;
      t_device += now_s() - t1;
    } else if (order == 1) {
      std::vector<double> R(ndof);
      GFGPU_CALL(gfgpu_term_assemble_host(e.term, U.data(), GFGPU_RESIDUAL, nullptr, R.data()));
      double t2 = now_s();
      t_device += t2 - t1;
      getfem::base_vector &V = ws.assembled_vector();
      if (V.size() < I.first() + nred) V.resize(std::max<size_type>(nprim, I.first() + nred), 0.0);
      if (Ered) {
        std::vector<double> Vr(nred, 0.0);
        GFGPU_CALL(gfgpu_reduction_restrict_add_host(Ered, 1.0, R.data(), Vr.data()));
        for (size_type d = 0; d < nred; ++d) V[I.first() + d] += Vr[d];
      } else {
        for (size_type d = 0; d < ndof; ++d) V[I.first() + d] += R[d];
      }
      t_fill += now_s() - t2;
    } else if (rt.family == GFGPU_SOURCE || rt.family == GFGPU_NORMAL_SOURCE || rt.no_tangent ||
               (rt.family == GFGPU_JIT && rt.jit_form2 == "(0.0)")) {  // (a run-time compiled LOAD: no order-2 tree exists)
      // an order-1 term contributes nothing to the tangent; K only gets its size (workspace.cc:805-812)
      getfem::model_real_sparse_matrix &K = ws.assembled_matrix();
      const size_type need = std::max<size_type>(nprim, I.first() + nred);
      if (gmm::mat_nrows(K) < need || gmm::mat_ncols(K) < need) gmm::resize(K, need, need);
    } else {
      GFGPU_CALL(gfgpu_term_assemble_host(e.term, U.data(), GFGPU_TANGENT, nullptr, nullptr));
      const double alpha = ws.factor_of_variable(rt.varname);  // alpha1 * alpha2 of the matrix assembly instructions
      terms_sig += key.str() + "@" + std::to_string(I.first()) + ";";
      pending_adds.push_back(pending_add{e.term, alpha * alpha, int64_t(I.first())});
      pending_adds.back().Er = pending_adds.back().Ec = Ered;
      ++n_added;
      t_device += now_s() - t1;
    }
    t0 = now_s();
    }  // groups of like convexes
  }
  if (order == 2 && n_added == 0) {  // only order-1 terms / empty regions: K just gets its size (workspace.cc:805-812)
    getfem::model_real_sparse_matrix &K = ws.assembled_matrix();
    if (gmm::mat_nrows(K) == gmm::mat_ncols(K) && gmm::mat_nrows(K) < need_all) gmm::resize(K, need_all, need_all);
    return;
  }
  if (order == 2) {
    double t1 = now_s();
    tangent_cache &tc = *tangent_;
    // same terms at the same places as last time: zero the values, keep the pattern (and the host copy of jc / ir)
    GFGPU_CALL(gfgpu_matrix_clear(tc.m, tc.sig == terms_sig ? 1 : 0));
    if (tc.sig != terms_sig) { tc.sig = terms_sig; tc.gen = -1; }
    for (const pending_add &pa : pending_adds) {
      if (pa.rect && (pa.Er || pa.Ec)) {
        GFGPU_CALL(gfgpu_matrix_add_rect_reduced(tc.m, pa.rect, pa.transposed, pa.Er, pa.Ec, pa.alpha, pa.off, pa.coff));
      } else if (pa.rect) {
        GFGPU_CALL(gfgpu_matrix_add_rect(tc.m, pa.rect, pa.transposed, pa.alpha, pa.off, pa.coff));
      } else if (pa.Er) {
        GFGPU_CALL(gfgpu_matrix_add_term_reduced(tc.m, pa.term, pa.Er, pa.alpha, pa.off, pa.off));
      } else {
        GFGPU_CALL(gfgpu_matrix_add_term(tc.m, pa.term, pa.alpha, pa.off, pa.off));
      }
    }
    const int64_t nnz = gfgpu_matrix_nnz(tc.m);
    tc.pr.resize((size_t)nnz);
    if (tc.gen != gfgpu_matrix_pattern_generation(tc.m) || tc.ir.size() != (size_t)nnz) {
      tc.jc.resize(need_all + 1);
      tc.ir.resize((size_t)nnz);
      GFGPU_CALL(gfgpu_matrix_export_csc_host(tc.m, tc.jc.data(), tc.ir.data(), tc.pr.data()));
      tc.gen = gfgpu_matrix_pattern_generation(tc.m);
      ++pattern_downloads;
    } else {
      GFGPU_CALL(gfgpu_matrix_export_csc_host(tc.m, nullptr, nullptr, tc.pr.data()));
    }
    double t2 = now_s();
    t_device += t2 - t1;
    getfem::model_real_sparse_matrix &K = ws.assembled_matrix();
    if (gmm::mat_nrows(K) != gmm::mat_ncols(K)) {
      // a RECTANGULAR matrix given by the caller -- asm_mass_matrix(B, mim, mf_mult, mf_u, rg) sets overlapping intervals
      // (0, n_mult) and (0, n_u) and a n_mult x n_u matrix (getfem_assembling.h:743-755): the reference trusts the caller with
      // the size (workspace.cc:807-811); every entry must fit
      for (size_type j = gmm::mat_ncols(K); j < need_all; ++j)
        GMM_ASSERT1(tc.jc[j] == tc.jc[j + 1], "gfgpu: the assembled matrix is too small for the terms (columns)");
      for (int32_t r : tc.ir) GMM_ASSERT1(size_type(r) < gmm::mat_nrows(K), "gfgpu: the assembled matrix is too small for the terms (rows)");
    } else if (gmm::mat_nrows(K) < need_all || gmm::mat_ncols(K) < need_all) gmm::resize(K, need_all, need_all);
    fill_col_matrix(K, std::min<size_type>(need_all, gmm::mat_ncols(K)), tc.jc.data(), tc.ir.data(), tc.pr.data());
    t_fill += now_s() - t2;
  }
}

// CSC -> gmm::col_matrix<rsvector>: each column is a row-sorted vector of (index, value) (gmm_vector.h:913-1030).  Empty columns
// take the device column as is; otherwise the values are added (accumulate-into-aliased-K, workspace.cc:805-812).  Columns are
// independent objects, so the loop runs on the host cores when OpenMP is on and we are not already inside a parallel region
// (the bricks' GETFEM_OMP_PARALLEL blocks): at benchmark size this fill, not the device, is what a drop-in caller waits for.
void fill_col_matrix(getfem::model_real_sparse_matrix &K, size_type ncols, const int64_t *jc, const int32_t *ir, const double *pr) {
  const long n = long(ncols);
  int nt = 1;
#ifdef _OPENMP
  if (!omp_in_parallel() && n >= 4096) nt = std::max(1, std::min(omp_get_num_procs(), 32));
#endif
#pragma omp parallel for schedule(static, 512) num_threads(nt) if (nt > 1)
  for (long j = 0; j < n; ++j) {
    const int64_t b = jc[j], en = jc[j + 1];
    if (b == en) continue;
    gmm::rsvector<double> &col = K[size_type(j)];
    if (col.nb_stored() == 0) {
      col.base_resize(size_type(en - b));
      auto itc = col.begin();
      for (int64_t k = b; k < en; ++k, ++itc) { itc->c = size_type(ir[size_t(k)]); itc->e = pr[size_t(k)]; }
    } else {
      for (int64_t k = b; k < en; ++k) col.w(size_type(ir[size_t(k)]), col.r(size_type(ir[size_t(k)])) + pr[size_t(k)]);
    }
  }
}

void assembly(getfem::ga_workspace &ws, size_type order, int device) {
  device_assembler a(device);
  a.assembly(ws, order);
}

}  // namespace getfem_b200
