// oracle/shim_test.cc -- TEST INFRASTRUCTURE.  In ONE process: the UNMODIFIED reference assembles with
// ga_workspace::assembly() on the CPU, the drop-in shim (getfem_b200/shim) assembles the SAME workspace
// description on the GPU through the C ABI, and the two gmm matrices / vectors are compared:
// CSC pattern identical, relative Frobenius / l2 differences printed as JSON (tests/test_gpu_shim.py).
#include <chrono>
#include <cstdio>
#include <map>
#include <random>

#include "getfem/getfem_generic_assembly.h"
#include "getfem/getfem_mesh_fem.h"
#include "getfem/getfem_mesh_im.h"
#include "getfem/getfem_im_data.h"
#include "getfem/getfem_regular_meshes.h"
#include "gfgpu_getfem_shim.h"
#include "gmm/gmm_kernel.h"

using getfem::size_type;

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

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
  const int dim = (int)geti("dim", 3), n = (int)geti("n", 4), K = (int)geti("k", 2), Q = (int)geti("q", 3);
  const int imdeg = (int)geti("im", 4);
  const std::string gt = gets("gt", "pk"), family = gets("family", "elast"), rgname = gets("region", "all");
  const double lambda = 1.3, mu = 0.7, acoef = 2.5;

  getfem::mesh m;
  std::vector<size_type> ns(dim, size_type(n));
  bgeot::pgeometric_trans pgt = gt == "pk" ? bgeot::simplex_geotrans(dim, 1) : bgeot::parallelepiped_geotrans(dim, 1);
  getfem::regular_unit_mesh(m, ns, pgt);
  getfem::mesh_fem mf(m, getfem::dim_type(Q));
  mf.set_classical_finite_element(getfem::dim_type(K));
  getfem::mesh_im mim(m);
  mim.set_integration_method(getfem::dim_type(imdeg));
  if (geti("mixed", 0)) {
    // a NON-UNIFORM mesh_fem / mesh_im (C&E.cc:5902-5936): the convexes whose barycentre has x < 0.5 get degree K-1 and a
    // lower integration method; the device path assembles one term per group of like convexes
    dal::bit_vector low;
    for (dal::bv_visitor cv(m.convex_index()); !cv.finished(); ++cv) {
      double bx = 0;
      for (size_type i = 0; i < pgt->nb_points(); ++i) bx += m.points_of_convex(cv)[i][0];
      if (bx / double(pgt->nb_points()) < 0.5) low.add(cv);
    }
    mf.set_classical_finite_element(low, getfem::dim_type(K - 1));
    mim.set_integration_method(low, getfem::dim_type(imdeg - 2));
  }
  const size_type ndof = mf.nb_dof();
  std::vector<double> U(ndof);
  if (family == "elast" || family == "laplace" || family == "mass" || family == "source" || family == "nsource") {
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> d(-1.0, 1.0);
    for (auto &v : U) v = d(rng);
  } else {
    for (size_type d = 0; d < ndof; ++d) {
      bgeot::base_node P = mf.point_of_basic_dof(d);
      int k = int(d % Q);
      U[d] = 0.03 * std::sin(2 * M_PI * P[(k + 1) % dim]) * std::cos(M_PI * P[k % dim]);
    }
  }
  std::string expr;
  if (family == "laplace") expr = "a*Grad_u:Grad_Test_u";
  else if (family == "mass") expr = "a*u.Test_u";
  else if (family == "source") expr = Q == 1 ? "-f*Test_u" : "-(f.Test_u)";
  else if (family == "nsource") expr = Q == 1 ? "((g).Normal)*Test_u" : "-(Reshape(g,qdim(u),meshdim)*Normal).Test_u";
  else if (family == "elast") expr = "(Div_u*((lambda)*Id(meshdim))+(2*(mu))*Sym(Grad_u)):Grad_Test_u";
  else {
    std::string law = family == "svk" ? "Saint_Venant_Kirchhoff"
                      : family == "nh_ciarlet" ? "Compressible_Neo_Hookean_Ciarlet"
                      : family == "mooney_rivlin" ? "Compressible_Mooney_Rivlin"
                      : family == "ciarlet_geymonat" ? "Ciarlet_Geymonat"
                      : family == "blatz_ko" ? "Generalized_Blatz_Ko" : "Compressible_Neo_Hookean_Bonet";
    expr = "((Id(meshdim)+Grad_u)*(" + law + "_PK2(Grad_u,params))):Grad_Test_u";
  }
  const std::vector<double> c_a{acoef}, c_l{lambda}, c_m{mu};
  const std::vector<double> c_p = family == "mooney_rivlin" ? std::vector<double>{0.8, 0.3, 2.0}             // C10, C01, D1
                                  : family == "ciarlet_geymonat" ? std::vector<double>{lambda, mu, 0.25}     // lambda, mu, a
                                  : family == "blatz_ko" ? std::vector<double>{1.0, 1.0, 1.5, -0.5, 1.5}      // a, b, c, d, n
                                                         : std::vector<double>{lambda, mu};
  std::vector<double> c_f(Q);
  for (int k = 0; k < Q; ++k) c_f[k] = 0.75 * (k + 1);
  std::vector<double> c_g(size_t(Q) * dim);
  for (size_t k = 0; k < c_g.size(); ++k) c_g[k] = 0.4 + 0.3 * double(k) * ((k & 1) ? -1.0 : 1.0);
  // region: all convexes, the outer faces, the faces on x = 1, or the convexes with barycentre x < 0.5
  getfem::mesh_region rg_sel;
  if (rgname == "half") {
    for (dal::bv_visitor cv(m.convex_index()); !cv.finished(); ++cv) {
      double bx = 0;
      for (size_type i = 0; i < pgt->nb_points(); ++i) bx += m.points_of_convex(cv)[i][0];
      if (bx / double(pgt->nb_points()) < 0.5) rg_sel.add(cv);
    }
  } else if (rgname != "all") {
    getfem::mesh_region outer;
    getfem::outer_faces_of_mesh(m, outer);
    for (getfem::mr_visitor v(outer); !v.finished(); ++v) {
      bgeot::base_node un = m.normal_of_face_of_convex(v.cv(), v.f());
      un /= gmm::vect_norm2(un);
      if (rgname == "outer" || (rgname == "xmax" && un[0] > 0.999)) rg_sel.add(v.cv(), v.f());
    }
  }
  getfem::mesh_region rg_outer;
  getfem::outer_faces_of_mesh(m, rg_outer);
  const getfem::mesh_region rg_all = getfem::mesh_region::all_convexes();
  const getfem::mesh_region &rg = rgname == "all" ? rg_all : rg_sel;
  // coef=fem: the coefficients / the load are fem data on classical mesh_fems of degree kd (add_fem_constant)
  const bool coef_fem = gets("coef", "const") == "fem";
  getfem::mesh_fem mf_d(m, 1), mf_dq(m, getfem::dim_type(Q));
  mf_d.set_classical_finite_element(getfem::dim_type(geti("kd", 1)));
  mf_dq.set_classical_finite_element(getfem::dim_type(geti("kd", 1)));
  std::vector<double> d_a(mf_d.nb_dof()), d_l(mf_d.nb_dof()), d_m(mf_d.nb_dof()), d_f(mf_dq.nb_dof());
  for (size_type d = 0; d < mf_d.nb_dof(); ++d) {
    bgeot::base_node P = mf_d.point_of_basic_dof(d);
    d_a[d] = acoef * (1.0 + 0.3 * std::sin(1.7 * P[0] + 0.9 * P[1]));
    d_l[d] = lambda * (1.0 + 0.25 * std::cos(1.1 * P[0] - 0.7 * P[1]));
    d_m[d] = mu * (1.0 + 0.2 * std::cos(1.3 * P[0] - P[dim - 1]));
  }
  for (size_type d = 0; d < mf_dq.nb_dof(); ++d) {
    bgeot::base_node P = mf_dq.point_of_basic_dof(d);
    d_f[d] = 0.75 * double(d % Q + 1) * (1.0 + 0.5 * P[0] - 0.25 * P[dim - 1]);
  }
  // coef=imd: the same coefficient fields stored per Gauss point in im_data objects (ga_workspace::add_im_data)
  const bool coef_imd = gets("coef", "const") == "imd";
  getfem::im_data imd_s(mim), imd_v(mim);
  {
    bgeot::multi_index ts(1);  // (multi_index(n) makes n ZERO entries)
    ts[0] = size_type(Q);
    imd_v.set_tensor_size(ts);
  }
  std::vector<double> i_a, i_l, i_m, i_f;
  if (coef_imd) {
    i_a.resize(imd_s.nb_index()); i_l.resize(imd_s.nb_index()); i_m.resize(imd_s.nb_index());
    i_f.resize(imd_v.nb_index() * Q);
    for (dal::bv_visitor cv(m.convex_index()); !cv.finished(); ++cv) {
      getfem::papprox_integration pai2 = mim.int_method_of_element(cv)->approx_method();
      bgeot::pgeometric_trans pgt2 = m.trans_of_convex(cv);
      for (size_type q = 0; q < pai2->nb_points_on_convex(); ++q) {
        bgeot::base_node P = pgt2->transform(pai2->point(q), m.points_of_convex(cv));
        const size_type is = imd_s.index_of_point(cv, q), iv = imd_v.index_of_point(cv, q);
        i_a[is] = acoef * (1.0 + 0.3 * std::sin(1.7 * P[0] + 0.9 * P[1]));
        i_l[is] = lambda * (1.0 + 0.25 * std::cos(1.1 * P[0] - 0.7 * P[1]));
        i_m[is] = mu * (1.0 + 0.2 * std::cos(1.3 * P[0] - P[dim - 1]));
        for (int c = 0; c < Q; ++c) i_f[iv * Q + c] = 0.75 * double(c + 1) * (1.0 + 0.5 * P[0] - 0.25 * P[dim - 1]);
      }
    }
  }
  auto setup = [&](getfem::ga_workspace &ws) {
    ws.add_fem_variable("u", mf, gmm::sub_interval(0, ndof), U);
    if (coef_imd) {
      ws.add_im_data("a", imd_s, i_a);
      ws.add_im_data("lambda", imd_s, i_l);
      ws.add_im_data("mu", imd_s, i_m);
      ws.add_im_data("f", imd_v, i_f);
    } else if (coef_fem) {
      ws.add_fem_constant("a", mf_d, d_a);
      ws.add_fem_constant("lambda", mf_d, d_l);
      ws.add_fem_constant("mu", mf_d, d_m);
      ws.add_fem_constant("f", mf_dq, d_f);
    } else {
      ws.add_fixed_size_constant("a", c_a);
      ws.add_fixed_size_constant("lambda", c_l);
      ws.add_fixed_size_constant("mu", c_m);
      ws.add_fixed_size_constant("f", c_f);
    }
    ws.add_fixed_size_constant("params", c_p);
    ws.add_fixed_size_constant("g", c_g);
    ws.add_expression(expr, mim, rg);
    if (a.count("model")) {  // a model-like workspace: + Robin mass on the outer faces + Neumann load + volumic source
      ws.add_expression("a*u.Test_u", mim, rg_outer);
      ws.add_expression(Q == 1 ? "-((g).Normal)*Test_u" : "-(Reshape(g,qdim(u),meshdim)*Normal).Test_u", mim, rg_outer);
      ws.add_expression(Q == 1 ? "-f*Test_u" : "-(f.Test_u)", mim);
    }
  };
  // ---- reference on the CPU
  getfem::ga_workspace wr;
  setup(wr);
  double t0 = now_s();
  wr.assembly(2);
  double t_ref2 = now_s() - t0;
  gmm::csc_matrix<double> Cr;
  Cr.init_with(wr.assembled_matrix());
  t0 = now_s();
  wr.assembly(1);
  double t_ref1 = now_s() - t0;
  std::vector<double> Rr(wr.assembled_vector().begin(), wr.assembled_vector().end());
  // ---- drop-in on the GPU (twice: the second call reuses the cached device state)
  getfem::ga_workspace wg;
  setup(wg);
  getfem_b200::device_assembler dev(0);
  t0 = now_s();
  dev.assembly(wg, 2);
  double t_gpu2_first = now_s() - t0;
  gmm::csc_matrix<double> Cg;
  Cg.init_with(wg.assembled_matrix());
  gmm::clear(wg.assembled_matrix());
  t0 = now_s();
  dev.assembly(wg, 2);
  double t_gpu2 = now_s() - t0;
  // the second call found the tangent resident on the device: same pattern generation, only the values travelled
  gmm::csc_matrix<double> Cg2;
  Cg2.init_with(wg.assembled_matrix());
  bool second_same = Cg2.jc.size() == Cg.jc.size() && Cg2.pr.size() == Cg.pr.size();
  for (size_t k = 0; second_same && k < Cg.pr.size(); ++k) second_same = Cg2.pr[k] == Cg.pr[k] && Cg2.ir[k] == Cg.ir[k];
  const double te = dev.t_extract, td = dev.t_device, tf = dev.t_fill;
  dev.assembly(wg, 1);
  std::vector<double> Rg(wg.assembled_vector().begin(), wg.assembled_vector().end());

  bool pattern_ok = Cr.nrows() == Cg.nrows() && Cr.jc.size() == Cg.jc.size() && Cr.ir.size() == Cg.ir.size();
  if (pattern_ok)
    for (size_t k = 0; k < Cr.jc.size() && pattern_ok; ++k) pattern_ok = Cr.jc[k] == Cg.jc[k];
  if (pattern_ok)
    for (size_t k = 0; k < Cr.ir.size() && pattern_ok; ++k) pattern_ok = Cr.ir[k] == Cg.ir[k];
  double nK = 0, dK = 0, nR = 0, dR = 0;
  if (pattern_ok)
    for (size_t k = 0; k < Cr.pr.size(); ++k) { nK += Cr.pr[k] * Cr.pr[k]; dK += (Cr.pr[k] - Cg.pr[k]) * (Cr.pr[k] - Cg.pr[k]); }
  for (size_type d = 0; d < ndof; ++d) { nR += Rr[d] * Rr[d]; dR += (Rr[d] - Rg[d]) * (Rr[d] - Rg[d]); }
  std::printf("{\"family\": \"%s\", \"ne\": %zu, \"ndof\": %zu, \"nnz_ref\": %zu, \"nnz_gpu\": %zu, \"pattern_ok\": %s, "
              "\"rel_K\": %.3e, \"rel_R\": %.3e, \"t_ref_asm2\": %.4f, \"t_ref_asm1\": %.4f, \"t_gpu_asm2_first\": %.4f, "
              "\"t_gpu_asm2\": %.4f, \"t_extract\": %.4f, \"t_device\": %.4f, \"t_fill\": %.4f, \"pattern_downloads\": %ld, "
              "\"second_call_same\": %s}\n",
              family.c_str(), m.convex_index().card(), ndof, Cr.pr.size(), Cg.pr.size(), pattern_ok ? "true" : "false",
              pattern_ok ? (nK > 0 ? std::sqrt(dK / nK) : 0.0) : -1.0, std::sqrt(dR / nR), t_ref2, t_ref1, t_gpu2_first, t_gpu2, te, td, tf,
              dev.pattern_downloads, second_same ? "true" : "false");
  return pattern_ok ? 0 : 1;
}
