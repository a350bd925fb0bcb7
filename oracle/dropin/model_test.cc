// oracle/dropin/model_test.cc -- TEST INFRASTRUCTURE.  Drop-in at the MODEL level: a getfem::model built from the
// reference's own bricks (add_isotropic_linearized_elasticity_brick, add_generic_elliptic_brick, add_source_term_brick on
// a volume and on a Neumann boundary, add_linear_term for a Robin condition, add_finite_strain_elasticity_brick) is
// assembled twice by model::assembly(BUILD_ALL) in ONE process, against ONE library (oracle/_ref/libgetfem_gfgpu.so =
// the unmodified reference objects + the dispatch patch of INTEGRATION.md section 2): once through the reference's
// ga_exec, once through the device path.  The model's tangent matrix and right-hand side must agree: CSC pattern
// identical, values / rhs to 1e-12 (tests/test_gpu_dropin.py).
#include <chrono>
#include <cstdio>
#include <map>
#include <random>

#include "getfem/getfem_assembling.h"
#include "getfem/getfem_models.h"
#include "getfem/getfem_nonlinear_elasticity.h"
#include "getfem/getfem_omp.h"
#include "getfem/getfem_partial_mesh_fem.h"
#include "getfem/getfem_regular_meshes.h"
#include "gmm/gmm_kernel.h"

namespace getfem_b200 {
void gfgpu_enable(bool on);
long gfgpu_device_calls();
long gfgpu_reference_calls();
void gfgpu_last_timings(double *t3);
long gfgpu_pattern_downloads();
}  // namespace getfem_b200

using getfem::size_type;

int main(int argc, char **argv) {
  std::map<std::string, std::string> a;
  for (int i = 1; i < argc; ++i) {
    std::string s(argv[i]);
    size_t e = s.find('=');
    if (e == std::string::npos) return 2;
    a[s.substr(0, e)] = s.substr(e + 1);
  }
  auto geti = [&](const char *k, long d) { return a.count(k) ? std::stol(a[k]) : d; };
  auto gets = [&](const char *k, const char *d) { return a.count(k) ? a[k] : std::string(d); };
  const std::string kind = gets("model", "elasticity");
  const int dim = (int)geti("dim", 3), n = (int)geti("n", 4), K = (int)geti("k", 2);
  const bool qk = gets("gt", "pk") == "qk";
  // threads > 1: the bricks' GETFEM_OMP_PARALLEL blocks slice the regions; the patch leaves that regime on the reference path
  if (geti("threads", 1) > 1) getfem::set_num_threads(int(geti("threads", 1)));
  const int Q = kind == "expr" ? (int)geti("q", dim) : kind == "timing" ? dim : (kind == "poisson" || kind == "asm_laplacian") ? 1 : dim;

  getfem::mesh m;
  std::vector<size_type> ns(dim, size_type(n));
  getfem::regular_unit_mesh(m, ns, qk ? bgeot::parallelepiped_geotrans(dim, 1) : bgeot::simplex_geotrans(dim, 1));
  // boundary regions: 1 = Neumann (x = 1), 2 = Robin (the other outer faces)
  getfem::mesh_region outer;
  getfem::outer_faces_of_mesh(m, outer);
  for (getfem::mr_visitor v(outer); !v.finished(); ++v) {
    bgeot::base_node un = m.normal_of_face_of_convex(v.cv(), v.f());
    un /= gmm::vect_norm2(un);
    m.region(un[0] > 0.999 ? 1 : 2).add(v.cv(), v.f());
  }
  getfem::mesh_fem mf(m, getfem::dim_type(Q));
  mf.set_classical_finite_element(getfem::dim_type(K));
  if (geti("reduced", 0) == 1) {
    // a REDUCED mesh_fem by selection of basic dofs (what partial_mesh_fem / the multiplier filters do)
    dal::bit_vector kept;
    for (size_type d = 0; d < mf.nb_basic_dof(); ++d) if (d % 5 != 1) kept.add(d);
    mf.reduce_to_basic_dof(kept);
  } else if (geti("reduced", 0) == 2) {
    // a GENERAL extension matrix (periodic / constrained spaces, mesh_fem::set_reduction_matrices): the last quarter of the
    // basic dofs are combinations of two reduced dofs each, so E^T K E sums several products per entry
    const size_type nb = mf.nb_basic_dof(), nr = nb - nb / 4;
    gmm::row_matrix<gmm::rsvector<double>> R(nr, nb), E(nb, nr);
    for (size_type j = 0; j < nr; ++j) { R(j, j) = 1.0; E(j, j) = 1.0; }
    for (size_type j = nr; j < nb; ++j) { E(j, j - nr) = 0.5; E(j, (7 * (j - nr) + 3) % nr) = -0.25; }
    mf.set_reduction_matrices(R, E);
  }
  getfem::mesh_im mim(m);
  mim.set_integration_method(getfem::dim_type(qk ? 2 * K + 2 : 2 * K));
  getfem::mesh_fem mf_d(m, 1);  // fem data: heterogeneous coefficient
  mf_d.set_classical_finite_element(1);
  getfem::mesh_fem mf_p(m, 1);  // pressure of the mixed formulation
  mf_p.set_classical_finite_element(getfem::dim_type(K > 1 ? K - 1 : 1));

  if (kind == "expr") {
    // one bilinear form written directly with Test_ / Test2_, in any of the algebraically equivalent spellings of the
    // reference's own tests (tests/test_assembly.cc:777-866, lambda = 3, mu = 2 there): reference path against device path
    const std::string expr = gets("expr", "Grad_Test_u:Grad_Test2_u");
    std::vector<double> U(mf.nb_dof(), 0.0), LAMBDA(1, 3.0), MU(1, 2.0), A(1, 1.7);
    {
      std::mt19937_64 rng(99);
      std::uniform_real_distribution<double> dist(-1.0, 1.0);
      for (auto &x : U) x = dist(rng);
    }
    if (a.count("uscale")) for (auto &x : U) x *= std::stod(a["uscale"]);
    if (a.count("uzero")) {  // state that vanishes on the first dofs (the probe convexes) or everywhere (uzero=-1)
      const long nz = geti("uzero", -1);
      for (size_t k = 0; k < U.size(); ++k)
        if (nz < 0 || long(k) < nz) U[k] = 0.0;
    }
    std::vector<double> Vr, Vg, C0(mf_d.nb_dof(), 1.0), PARAMS{1.3, 0.7};
    std::vector<double> DVEC(dim), AMAT(dim * dim);
    for (int k = 0; k < dim; ++k) DVEC[k] = 0.5 * (k + 1);
    for (int k = 0; k < dim * dim; ++k) AMAT[k] = (k % (dim + 1) == 0 ? 2.0 : 0.0) + 0.1 * k;
    const bool with_p = geti("pvar", 0) != 0;
    std::vector<double> Pv(mf_p.nb_dof());
    {
      std::mt19937_64 rng(7);
      std::uniform_real_distribution<double> dist(-1.0, 1.0);
      for (auto &x : Pv) x = dist(rng);
    }
    double Er = 0, Eg = 0;
    C0.back() = 5.0;
    getfem::mesh_fem mf_dv(m, getfem::dim_type(dim));
    mf_dv.set_classical_finite_element(1);
    std::vector<double> W0(mf_dv.nb_dof());
    for (size_type d = 0; d < mf_dv.nb_dof(); ++d) {
      bgeot::base_node P = mf_dv.point_of_basic_dof(d);
      W0[d] = (d % dim == 0 ? 1.0 : -0.5) + 0.3 * std::sin(2.0 * P[0] + 1.3 * P[dim - 1] + double(d % dim));
    }
    auto run = [&](bool device, gmm::csc_matrix<double> &C) {
      getfem_b200::gfgpu_enable(device);
      getfem::ga_workspace ws;
      ws.add_fem_variable("u", mf, gmm::sub_interval(0, mf.nb_dof()), U);
      if (with_p) ws.add_fem_variable("p", mf_p, gmm::sub_interval(mf.nb_dof(), mf_p.nb_dof()), Pv);  // mixed forms
      ws.add_fixed_size_constant("lambda", LAMBDA);
      ws.add_fixed_size_constant("mu", MU);
      ws.add_fixed_size_constant("a", A);
      ws.add_fixed_size_constant("params", PARAMS);
      ws.add_fixed_size_constant("dvec", DVEC);  // a vector constant of the mesh dimension
      ws.add_fixed_size_constant("amat", AMAT);  // N x N entries (an anisotropic diffusion tensor through Reshape(amat,N,N)), not symmetric
      ws.add_fem_constant("c0", mf_d, C0);  // a material that is 1 on the first convexes and 5 in a far corner
      ws.add_fem_constant("w0", mf_dv, W0);  // a vector-valued field (an advection velocity)
      if (a.count("region")) ws.add_expression(expr, mim, m.region(size_type(geti("region", 1))));
      else ws.add_expression(expr, mim);
      const size_type ntot = mf.nb_dof() + (with_p ? mf_p.nb_dof() : 0);
      getfem::model_real_sparse_matrix M(ntot, ntot);
      ws.set_assembled_matrix(M);
      ws.assembly(2);
      C.init_with(M);
      ws.assembly(1);  // an expression written in u (not in Test2_u) also has a residual
      (device ? Vg : Vr).assign(ws.assembled_vector().begin(), ws.assembled_vector().end());
      bool has0 = false;  // the potential, when the expression has an order-0 tree
      for (size_type i = 0; i < ws.nb_trees(); ++i) has0 = has0 || ws.tree_info(i).order == 0;
      if (has0) {
        ws.assembly(0);
        (device ? Eg : Er) = ws.assembled_potential();
      }
      getfem_b200::gfgpu_enable(false);
    };
    gmm::csc_matrix<double> Cr, Cg;
    run(false, Cr);
    run(true, Cg);
    double nV = 0, dV = 0;
    for (size_t k = 0; k < Vr.size() && k < Vg.size(); ++k) { nV += Vr[k] * Vr[k]; dV += (Vr[k] - Vg[k]) * (Vr[k] - Vg[k]); }
    bool pattern_ok = Cr.jc.size() == Cg.jc.size() && Cr.ir.size() == Cg.ir.size();
    for (size_t k = 0; pattern_ok && k < Cr.jc.size(); ++k) pattern_ok = Cr.jc[k] == Cg.jc[k];
    for (size_t k = 0; pattern_ok && k < Cr.ir.size(); ++k) pattern_ok = Cr.ir[k] == Cg.ir[k];
    double nK = 0, dK = 0;
    size_t n_only = 0;
    double max_only_rel = 0;
    if (geti("reduced", 0) && Cr.jc.size() == Cg.jc.size()) {
      // Reduced mesh_fems: the reference's E^T K E goes through gmm's sparse products and K(i,j) += m, which REMOVE an entry
      // whose running sum is exactly 0.0 (rsvector::w).  An entry of K_basic that cancels to 1e-17 on one path and to 0.0 on the
      // other is therefore stored on one side only: the comparison runs over the union of the two patterns, the one-sided
      // entries are counted and must be round-off (same criterion as the model-level comparison below).
      double maxK = 0;
      for (size_t j = 0; j + 1 < Cr.jc.size(); ++j) {
        size_t a0 = Cr.jc[j], a1 = Cr.jc[j + 1], b0 = Cg.jc[j], b1 = Cg.jc[j + 1];
        while (a0 < a1 || b0 < b1) {
          const size_t ra = a0 < a1 ? Cr.ir[a0] : size_t(-1), rb = b0 < b1 ? Cg.ir[b0] : size_t(-1);
          double va = 0, vb = 0;
          if (ra <= rb) va = Cr.pr[a0++];
          if (rb <= ra) vb = Cg.pr[b0++];
          if (ra != rb) { ++n_only; max_only_rel = std::max(max_only_rel, std::max(std::fabs(va), std::fabs(vb))); }
          nK += va * va; dK += (va - vb) * (va - vb);
          maxK = std::max(maxK, std::fabs(va));
        }
      }
      max_only_rel = maxK > 0 ? max_only_rel / maxK : 0.0;
      pattern_ok = max_only_rel < 1e-14;
    } else if (pattern_ok)
      for (size_t k = 0; k < Cr.pr.size(); ++k) { nK += Cr.pr[k] * Cr.pr[k]; dK += (Cr.pr[k] - Cg.pr[k]) * (Cr.pr[k] - Cg.pr[k]); }
    std::printf("{\"model\": \"expr\", \"ndof\": %zu, \"nnz_ref\": %zu, \"nnz_gpu\": %zu, \"pattern_ok\": %s, \"rel_K\": %.3e, "
                "\"rel_V\": %.3e, \"norm_V\": %.3e, \"E_ref\": %.17g, \"E_gpu\": %.17g, \"entries_on_one_side_only\": %zu, "
                "\"max_one_sided_rel\": %.3e, \"device_workspace_calls\": %ld}\n",
                size_t(mf.nb_dof()), Cr.pr.size(), Cg.pr.size(), pattern_ok ? "true" : "false",
                pattern_ok && nK > 0 ? std::sqrt(dK / nK) : -1.0, nV > 0 ? std::sqrt(dV / nV) : 0.0, std::sqrt(nV), Er, Eg, n_only,
                max_only_rel, getfem_b200::gfgpu_device_calls());
    return pattern_ok ? 0 : 1;
  }
  if (kind == "timing") {
    // The REAL drop-in, timed: ga_workspace::assembly(2) (+ assembly(1)) of the C3 form through libgetfem_gfgpu.so, wall
    // clock around the reference's own call, host containers in and out (model_real_sparse_matrix = col_matrix<rsvector>).
    // reps device calls on one workspace (a Newton / time loop: the pattern is downloaded once), then, with ref=1, the
    // reference's own ga_exec on the same workspace -- the same-configuration ratio of INTEGRATION.md.
    const int reps = (int)geti("reps", 3);
    std::vector<double> U(mf.nb_dof()), LAMBDA(1, 1.0), MU(1, 1.0);
    {
      std::mt19937_64 rng(99);
      std::uniform_real_distribution<double> dist(-1.0, 1.0);
      for (auto &x : U) x = dist(rng);
    }
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    getfem::ga_workspace ws;
    ws.add_fem_variable("u", mf, gmm::sub_interval(0, mf.nb_dof()), U);
    ws.add_fixed_size_constant("lambda", LAMBDA);
    ws.add_fixed_size_constant("mu", MU);
    ws.add_expression(gets("expr", "lambda*Div_u*Div_Test_u + mu*(Grad_u+Grad_u'):Grad_Test_u"), mim);
    std::vector<double> t2(reps), t1(reps), tx(reps), td(reps), tf(reps);
    double chk = 0, chkv = 0;
    size_t nnz = 0;
    getfem_b200::gfgpu_enable(true);
    {
      getfem::model_real_sparse_matrix M(mf.nb_dof(), mf.nb_dof());
      ws.set_assembled_matrix(M);
      for (int r = 0; r < reps; ++r) {
        gmm::clear(M);
        double a0 = now();
        ws.assembly(2);
        t2[r] = now() - a0;
        double t3[3];
        getfem_b200::gfgpu_last_timings(t3);
        tx[r] = t3[0]; td[r] = t3[1]; tf[r] = t3[2];
        a0 = now();
        ws.assembly(1);
        t1[r] = now() - a0;
      }
      nnz = gmm::nnz(M);
      for (size_type j = 0; j < gmm::mat_ncols(M); ++j)
        for (auto it = M[j].begin(); it != M[j].end(); ++it) chk += it->e * it->e;
      for (double x : ws.assembled_vector()) chkv += x * x;
    }
    const long downloads = getfem_b200::gfgpu_pattern_downloads();
    getfem_b200::gfgpu_enable(false);
    double tr2 = -1, tr1 = -1, chkr = 0, chkrv = 0;
    if (geti("ref", 0)) {
      getfem::model_real_sparse_matrix M(mf.nb_dof(), mf.nb_dof());
      ws.set_assembled_matrix(M);
      double a0 = now();
      ws.assembly(2);
      tr2 = now() - a0;
      a0 = now();
      ws.assembly(1);
      tr1 = now() - a0;
      for (size_type j = 0; j < gmm::mat_ncols(M); ++j)
        for (auto it = M[j].begin(); it != M[j].end(); ++it) chkr += it->e * it->e;
      for (double x : ws.assembled_vector()) chkrv += x * x;
    }
    auto lst = [&](const std::vector<double> &v) {
      std::string o = "[";
      for (size_t k = 0; k < v.size(); ++k) { char b[32]; std::snprintf(b, sizeof b, "%s%.4f", k ? ", " : "", v[k]); o += b; }
      return o + "]";
    };
    std::printf("{\"model\": \"timing\", \"n\": %d, \"elements\": %zu, \"ndof\": %zu, \"nnz\": %zu, \"assembly2_s\": %s, "
                "\"assembly1_s\": %s, \"extract_s\": %s, \"device_s\": %s, \"fill_s\": %s, \"pattern_downloads\": %ld, "
                "\"norm_K\": %.15e, \"norm_V\": %.15e, \"ref_assembly2_s\": %.4f, \"ref_assembly1_s\": %.4f, "
                "\"ref_norm_K\": %.15e, \"ref_norm_V\": %.15e, \"device_workspace_calls\": %ld}\n",
                n, size_t(m.convex_index().card()), size_t(mf.nb_dof()), nnz, lst(t2).c_str(), lst(t1).c_str(), lst(tx).c_str(),
                lst(td).c_str(), lst(tf).c_str(), downloads, std::sqrt(chk), std::sqrt(chkv), tr2, tr1, std::sqrt(chkr),
                std::sqrt(chkrv), getfem_b200::gfgpu_device_calls());
    return 0;
  }
  if (kind.rfind("asm_", 0) == 0) {
    // the legacy asm_* wrappers (getfem_assembling.h): thin layers over ga_workspace, written with Test / Test2 directly
    std::vector<double> LAMBDA(mf_d.nb_dof()), MU(mf_d.nb_dof());
    for (size_type d = 0; d < mf_d.nb_dof(); ++d) {
      bgeot::base_node P = mf_d.point_of_basic_dof(d);
      LAMBDA[d] = 1.3 * (1.0 + 0.25 * std::cos(1.1 * P[0] - 0.7 * P[1]));
      MU[d] = 0.7 * (1.0 + 0.2 * std::cos(1.3 * P[0] - P[dim - 1]));
    }
    // a multiplier space: the fem of one degree less, whole (asm_mass_rect) or restricted to the dofs of the Robin boundary
    // through a partial_mesh_fem (asm_mass_rect_partial: a REDUCED mesh_fem, what model::add_multiplier builds)
    getfem::mesh_fem mf_l(m, getfem::dim_type(Q));
    mf_l.set_classical_finite_element(getfem::dim_type(K > 1 ? K - 1 : 1));
    getfem::partial_mesh_fem mf_lp(mf_l);
    mf_lp.adapt(mf_l.basic_dof_on_region(m.region(2)));
    getfem::partial_mesh_fem mf_up(mf);  // the unknown itself restricted to the dofs of the Neumann boundary
    mf_up.adapt(mf.basic_dof_on_region(m.region(1)));
    auto run = [&](bool device, gmm::csc_matrix<double> &C) {
      getfem_b200::gfgpu_enable(device);
      getfem::model_real_sparse_matrix M(mf.nb_dof(), mf.nb_dof());
      if (kind == "asm_mass_rect") { gmm::resize(M, mf_l.nb_dof(), mf.nb_dof()); getfem::asm_mass_matrix(M, mim, mf_l, mf, m.region(2)); }
      else if (kind == "asm_mass_rect_volume") { gmm::resize(M, mf_l.nb_dof(), mf.nb_dof()); getfem::asm_mass_matrix(M, mim, mf_l, mf); }
      else if (kind == "asm_mass_rect_partial") { gmm::resize(M, mf_lp.nb_dof(), mf.nb_dof()); getfem::asm_mass_matrix(M, mim, mf_lp, mf, m.region(2)); }
      else if (kind == "asm_mass_partial") { gmm::resize(M, mf_up.nb_dof(), mf_up.nb_dof()); getfem::asm_mass_matrix(M, mim, mf_up, m.region(1)); }
      else if (kind == "asm_mass_partial_both") { gmm::resize(M, mf_lp.nb_dof(), mf_up.nb_dof()); getfem::asm_mass_matrix(M, mim, mf_lp, mf_up); }
      else if (kind == "asm_mass") getfem::asm_mass_matrix(M, mim, mf);
      else if (kind == "asm_laplacian") getfem::asm_stiffness_matrix_for_homogeneous_laplacian(M, mim, mf);
      else if (kind == "asm_elasticity") getfem::asm_stiffness_matrix_for_linear_elasticity(M, mim, mf, mf_d, LAMBDA, MU);
      else if (kind == "asm_mass_boundary") getfem::asm_mass_matrix(M, mim, mf, m.region(2));
      else GMM_ASSERT1(false, "unknown asm_ case");
      C.init_with(M);
      getfem_b200::gfgpu_enable(false);
    };
    gmm::csc_matrix<double> Cr, Cg;
    run(false, Cr);
    run(true, Cg);
    bool pattern_ok = Cr.jc.size() == Cg.jc.size() && Cr.ir.size() == Cg.ir.size();
    for (size_t k = 0; pattern_ok && k < Cr.jc.size(); ++k) pattern_ok = Cr.jc[k] == Cg.jc[k];
    for (size_t k = 0; pattern_ok && k < Cr.ir.size(); ++k) pattern_ok = Cr.ir[k] == Cg.ir[k];
    double nK = 0, dK = 0;
    if (pattern_ok)
      for (size_t k = 0; k < Cr.pr.size(); ++k) { nK += Cr.pr[k] * Cr.pr[k]; dK += (Cr.pr[k] - Cg.pr[k]) * (Cr.pr[k] - Cg.pr[k]); }
    std::printf("{\"model\": \"%s\", \"ndof\": %zu, \"nnz_ref\": %zu, \"nnz_gpu\": %zu, \"pattern_ok\": %s, \"rel_K\": %.3e, "
                "\"device_workspace_calls\": %ld}\n",
                kind.c_str(), size_t(mf.nb_dof()), Cr.pr.size(), Cg.pr.size(), pattern_ok ? "true" : "false",
                pattern_ok && nK > 0 ? std::sqrt(dK / nK) : -1.0, getfem_b200::gfgpu_device_calls());
    return pattern_ok ? 0 : 1;
  }
  auto assemble = [&](bool device, gmm::csc_matrix<double> &C, std::vector<double> &rhs) {
  // a FRESH model per path: model::assembly caches the matrices of linear bricks
  getfem::model md;
  md.add_fem_variable("u", mf);
  const size_type ndof = mf.nb_dof();
  std::vector<double> U(ndof);
  if (kind == "finite_strain") {
    for (size_type d = 0; d < ndof; ++d) {
      bgeot::base_node P = mf.point_of_basic_dof(d);
      int k = int(d % Q);
      U[d] = 0.03 * std::sin(2 * M_PI * P[(k + 1) % dim]) * std::cos(M_PI * P[k % dim]);
    }
  } else {
    std::mt19937_64 rng(4321);
    std::uniform_real_distribution<double> dist(-1.0, 1.0);
    for (auto &v : U) v = dist(rng);
  }
  gmm::copy(U, md.set_real_variable("u"));

  std::vector<double> F(Q), G(Q);
  for (int k = 0; k < Q; ++k) { F[k] = 0.5 * (k + 1); G[k] = -0.3 * (k + 2); }
  md.add_initialized_fixed_size_data("F", F);
  md.add_initialized_fixed_size_data("G", G);
  md.add_initialized_scalar_data("robin", 7.5);
  if (kind == "elasticity" || kind == "incompressible") {
    md.add_initialized_scalar_data("lambda", kind == "incompressible" ? 0.0 : 1.3);
    md.add_initialized_scalar_data("mu", 0.7);
    getfem::add_isotropic_linearized_elasticity_brick(md, mim, "u", "lambda", "mu");
    if (kind == "incompressible") {
      // mixed formulation: a pressure on a scalar fem of one degree less, coupled trees (Test_u, Test2_p) and (Test_p, Test2_u)
      // (add_linear_incompressibility, getfem_models.cc:6373-6409: "-p*Div_Test_u-Test_p*Div_u")
      md.add_fem_variable("p", mf_p);
      std::vector<double> P(mf_p.nb_dof());
      std::mt19937_64 rng(99);
      std::uniform_real_distribution<double> dist(-1.0, 1.0);
      for (auto &v : P) v = dist(rng);
      gmm::copy(P, md.set_real_variable("p"));
      getfem::add_linear_incompressibility(md, mim, "u", "p");
    }
  } else if (kind == "poisson") {
    std::vector<double> A(mf_d.nb_dof());
    for (size_type d = 0; d < mf_d.nb_dof(); ++d) {
      bgeot::base_node P = mf_d.point_of_basic_dof(d);
      A[d] = 1.0 + 0.4 * std::sin(2.0 * P[0] + P[1]);
    }
    md.add_initialized_fem_data("a", mf_d, A);
    getfem::add_generic_elliptic_brick(md, mim, "u", "a");
  } else {
    std::vector<double> params{1.3, 0.7};
    const std::string lawn = gets("law", "Saint_Venant_Kirchhoff");
    if (lawn == "Compressible_Mooney_Rivlin") params = {0.8, 0.3, 2.0};
    else if (lawn == "Ciarlet_Geymonat") params = {1.3, 0.7, 0.25};
    else if (lawn == "Generalized_Blatz_Ko") params = {1.0, 1.0, 1.5, -0.5, 1.5};
    md.add_initialized_fixed_size_data("params", params);
    getfem::add_finite_strain_elasticity_brick(md, mim, gets("law", "Saint_Venant_Kirchhoff"), "u", "params");
  }
  getfem::add_source_term_brick(md, mim, "u", "F");          // volumic load
  getfem::add_source_term_brick(md, mim, "u", "G", 1);       // Neumann load on x = 1
  getfem::add_linear_term(md, mim, "robin*u.Test_u", 2);     // Robin condition on the rest of the boundary
  if (gets("dirichlet", "") == "mult") {
    // the standard Dirichlet brick: a multiplier on a REDUCED mesh_fem (model::add_multiplier filters the dofs of the region
    // through a partial_mesh_fem), constraint matrix = asm_mass_matrix(B, mim, mf_mult, mf_u, region), right-hand side
    // asm_source_term on the multiplier space (getfem_models.cc:4386-4450)
    std::vector<double> D(Q);
    for (int k = 0; k < Q; ++k) D[k] = 0.25 * (k + 1);
    md.add_initialized_fixed_size_data("Dd", D);
    getfem::add_Dirichlet_condition_with_multipliers(md, mim, "u", getfem::dim_type(K > 1 ? K - 1 : 1), 1, "Dd");
  } else if (gets("dirichlet", "") == "penal") {
    std::vector<double> D(Q);
    for (int k = 0; k < Q; ++k) D[k] = 0.25 * (k + 1);
    md.add_initialized_fixed_size_data("Dd", D);
    getfem::add_Dirichlet_condition_with_penalization(md, mim, "u", 1e6, 1, "Dd");
  }
  if (geti("empty_region", 0)) {  // bricks on regions without any element: ga_exec walks nothing, the device path must do the same
    getfem::add_source_term_brick(md, mim, "u", "G", 77);
    getfem::add_linear_term(md, mim, "robin*u.Test_u", 78);
  }

    getfem_b200::gfgpu_enable(device);
    md.assembly(getfem::model::BUILD_ALL);
    C.init_with(md.real_tangent_matrix());
    rhs.assign(md.real_rhs().begin(), md.real_rhs().end());
    getfem_b200::gfgpu_enable(false);
  };
  gmm::csc_matrix<double> Cr, Cg;
  std::vector<double> Rr, Rg;
  assemble(false, Cr, Rr);
  const long ref_calls = getfem_b200::gfgpu_reference_calls();
  assemble(true, Cg, Rg);
  const long dev_calls = getfem_b200::gfgpu_device_calls();

  // Model level: the bricks copy the workspace matrices with gmm::copy, which drops entries that are EXACTLY zero
  // (gmm_blas.h copy of sparse vectors).  An entry that cancels to round-off (1e-17) in one path and to 0.0 in the other is
  // therefore stored by one model tangent only: the comparison is made over the union of the two patterns, and the entries
  // present on one side only are reported with their magnitude (they must be round-off).
  const size_t nc = Cr.jc.size() - 1;
  double nK = 0, dK = 0, nR = 0, dR = 0, maxK = 0, max_only = 0;
  size_t n_only = 0;
  bool sizes_ok = Cr.jc.size() == Cg.jc.size();
  for (size_t j = 0; sizes_ok && j < nc; ++j) {
    size_t a0 = Cr.jc[j], a1 = Cr.jc[j + 1], b0 = Cg.jc[j], b1 = Cg.jc[j + 1];
    while (a0 < a1 || b0 < b1) {
      const size_t ra = a0 < a1 ? Cr.ir[a0] : size_t(-1), rb = b0 < b1 ? Cg.ir[b0] : size_t(-1);
      double va = 0, vb = 0;
      if (ra <= rb) va = Cr.pr[a0++];
      if (rb <= ra) vb = Cg.pr[b0++];
      if (ra != rb) { ++n_only; max_only = std::max(max_only, std::max(std::fabs(va), std::fabs(vb))); }
      nK += va * va; dK += (va - vb) * (va - vb);
      maxK = std::max(maxK, std::fabs(va));
    }
  }
  for (size_t d = 0; d < Rr.size(); ++d) { nR += Rr[d] * Rr[d]; dR += (Rr[d] - Rg[d]) * (Rr[d] - Rg[d]); }
  std::printf("{\"model\": \"%s\", \"ne\": %zu, \"ndof\": %zu, \"nnz_ref\": %zu, \"nnz_gpu\": %zu, \"entries_on_one_side_only\": %zu, "
              "\"max_one_sided_rel\": %.3e, \"rel_K\": %.3e, \"rel_rhs\": %.3e, \"reference_workspace_calls\": %ld, "
              "\"device_workspace_calls\": %ld}\n",
              kind.c_str(), m.convex_index().card(), size_t(mf.nb_dof()), Cr.pr.size(), Cg.pr.size(), n_only,
              maxK > 0 ? max_only / maxK : 0.0, sizes_ok && nK > 0 ? std::sqrt(dK / nK) : -1.0, nR > 0 ? std::sqrt(dR / nR) : 0.0,
              ref_calls, dev_calls);
  return sizes_ok ? 0 : 1;
}
