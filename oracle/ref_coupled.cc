// oracle/ref_coupled.cc -- TEST INFRASTRUCTURE.  Golden-vector generator for COUPLED trees: the unmodified reference
// (oracle/_ref/libgetfem.so) assembles "-p*Div_Test_u - Test_p*Div_u" (the mixed part of the incompressibility bricks,
// getfem_models.cc add_linear_incompressibility) on a mesh with two Lagrange fems (u: vector degree k, p: scalar degree kp)
// and dumps the two off-diagonal blocks of the workspace matrix (CSC), the two residual parts, and everything the device
// path is given: mesh, dof tables, reference tables of both fems at the quadrature points.
//   gf_ref_coupled dim=3 n=2 gt=pk k=2 kp=1 im=4 noise=0.15 out=<dir>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <map>
#include <random>
#include <string>
#include <sys/stat.h>

#include "getfem/getfem_generic_assembly.h"
#include "getfem/getfem_generic_assembly_tree.h"
#include "getfem/getfem_mesh_fem.h"
#include "getfem/getfem_mesh_im.h"
#include "getfem/getfem_regular_meshes.h"
#include "gmm/gmm_kernel.h"

using getfem::size_type;

static void write_npy(const std::string &path, const char *descr, size_t itemsize, const std::vector<size_t> &shape, const void *data) {
  std::string hdr = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': (";
  size_t tot = 1;
  for (size_t i = 0; i < shape.size(); ++i) {
    hdr += std::to_string(shape[i]) + (shape.size() == 1 || i + 1 < shape.size() ? "," : "") + (i + 1 < shape.size() ? " " : "");
    tot *= shape[i];
  }
  hdr += "), }";
  hdr.append((64 - (10 + hdr.size() + 1) % 64) % 64, ' ');
  hdr += "\n";
  std::ofstream f(path, std::ios::binary);
  const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
  f.write((const char *)magic, 8);
  uint16_t hl = (uint16_t)hdr.size();
  f.write((const char *)&hl, 2);
  f.write(hdr.data(), hdr.size());
  f.write((const char *)data, tot * itemsize);
}
static void f64(const std::string &p, const std::vector<size_t> &s, const std::vector<double> &v) { write_npy(p, "<f8", 8, s, v.data()); }
static void i64(const std::string &p, const std::vector<size_t> &s, const std::vector<int64_t> &v) { write_npy(p, "<i8", 8, s, v.data()); }

int main(int argc, char **argv) {
  std::map<std::string, std::string> a;
  for (int i = 1; i < argc; ++i) {
    std::string s(argv[i]);
    size_t e = s.find('=');
    if (e == std::string::npos) return 2;
    a[s.substr(0, e)] = s.substr(e + 1);
  }
  auto geti = [&](const char *k, long d) { return a.count(k) ? std::stol(a[k]) : d; };
  const int dim = (int)geti("dim", 3), n = (int)geti("n", 2), K = (int)geti("k", 2), KP = (int)geti("kp", 1), imdeg = (int)geti("im", 4);
  const bool qk = a.count("gt") && a["gt"] == "qk";
  const double noise = a.count("noise") ? std::stod(a["noise"]) : 0.0;
  const std::string out = a.count("out") ? a["out"] : "";
  bgeot::pgeometric_trans pgt = qk ? bgeot::parallelepiped_geotrans(dim, 1) : bgeot::simplex_geotrans(dim, 1);
  getfem::mesh m;
  std::vector<size_type> ns(dim, size_type(n));
  if (a.count("nz")) ns[dim - 1] = size_type(geti("nz", n));
  getfem::regular_unit_mesh(m, ns, pgt);
  if (noise > 0) {  // same scheme as ref_driver.cc: nodes moved, mesh rebuilt convex by convex, numbering unchanged
    std::vector<bgeot::base_node> np(m.points_index().last_true() + 1);
    std::mt19937_64 rng(777);
    std::uniform_real_distribution<double> dist(-1.0, 1.0);
    for (dal::bv_visitor i(m.points_index()); !i.finished(); ++i) {
      np[i] = m.points()[i];
      for (int d = 0; d < dim; ++d) np[i][d] += noise * dist(rng) / double(ns[d]);
    }
    getfem::mesh m2;
    for (dal::bv_visitor cv(m.convex_index()); !cv.finished(); ++cv) {
      std::vector<bgeot::base_node> cp;
      for (size_type i : m.ind_points_of_convex(cv)) cp.push_back(np[i]);
      m2.add_convex_by_points(m.trans_of_convex(cv), cp.begin());
    }
    m.clear();
    m.copy_from(m2);
  }
  getfem::mesh_fem mfu(m, getfem::dim_type(dim)), mfp(m, 1);
  mfu.set_classical_finite_element(getfem::dim_type(K));
  mfp.set_classical_finite_element(getfem::dim_type(KP));
  getfem::mesh_im mim(m);
  mim.set_integration_method(getfem::dim_type(imdeg));
  const size_type nu = mfu.nb_dof(), np_ = mfp.nb_dof(), ne = m.convex_index().card();
  std::vector<double> U(nu), P(np_);
  {
    std::mt19937_64 rng(4242);
    std::uniform_real_distribution<double> dist(-1.0, 1.0);
    for (auto &x : U) x = dist(rng);
    for (auto &x : P) x = dist(rng);
  }
  getfem::ga_workspace ws;
  ws.add_fem_variable("u", mfu, gmm::sub_interval(0, nu), U);
  ws.add_fem_variable("p", mfp, gmm::sub_interval(nu, np_), P);
  const std::string expr = a.count("expr") ? a["expr"] : "-p*Div_Test_u - Test_p*Div_u";
  ws.add_expression(expr, mim);
  getfem::model_real_sparse_matrix Kmat(nu + np_, nu + np_);
  ws.set_assembled_matrix(Kmat);
  ws.assembly(2);
  ws.assembly(1);
  std::vector<double> R(ws.assembled_vector().begin(), ws.assembled_vector().end());
  std::string trees;
  for (size_type i = 0; i < ws.nb_trees(); ++i) {
    const auto &td = ws.tree_info(i);
    trees += (i ? "; " : "") + std::to_string(int(td.order)) + ":" + td.name_test1 + "," + td.name_test2 + ":" + getfem::ga_tree_to_string(*td.ptree);
  }
  // blocks of the workspace matrix, entry by entry as the rsvector columns store them (gmm::copy would drop the entries
  // whose contributions cancelled to an exact 0.0 -- they ARE stored by add_elem_matrix, C&E.cc:4853-4936)
  struct csc { std::vector<int64_t> jc, ir; std::vector<double> pr; };
  auto block = [&](size_type r0, size_type nr, size_type c0, size_type nc) {
    csc C;
    C.jc.push_back(0);
    for (size_type j = c0; j < c0 + nc; ++j) {
      for (auto it = Kmat[j].begin(); it != Kmat[j].end(); ++it)
        if (it->c >= r0 && it->c < r0 + nr) { C.ir.push_back(int64_t(it->c - r0)); C.pr.push_back(it->e); }
      C.jc.push_back(int64_t(C.ir.size()));
    }
    return C;
  };
  const csc Cup = block(0, nu, nu, np_), Cpu = block(nu, np_, 0, nu), Cuu = block(0, nu, 0, nu), Cpp = block(nu, np_, nu, np_);
  const size_type cv0 = m.convex_index().first_true();
  getfem::pfem pfu = mfu.fem_of_element(cv0), pfp = mfp.fem_of_element(cv0);
  getfem::papprox_integration pai = mim.int_method_of_element(cv0)->approx_method();
  const size_type nq = pai->nb_points_on_convex(), ng = pgt->nb_points(), ndu = pfu->nb_dof(cv0), ndp = pfp->nb_dof(cv0);
  std::printf("{\"dim\": %d, \"ne\": %zu, \"ndof_u\": %zu, \"ndof_p\": %zu, \"nnz_up\": %zu, \"nnz_pu\": %zu, \"nnz_uu\": %zu, "
              "\"nnz_pp\": %zu, \"fem_u\": \"%s\", \"fem_p\": \"%s\", \"im\": \"%s\", \"nq\": %zu, \"expr\": \"%s\", \"trees\": \"%s\"}\n",
              dim, size_t(ne), size_t(nu), size_t(np_), Cup.ir.size(), Cpu.ir.size(), Cuu.ir.size(), Cpp.ir.size(),
              getfem::name_of_fem(pfu).c_str(), getfem::name_of_fem(pfp).c_str(),
              getfem::name_of_int_method(mim.int_method_of_element(cv0)).c_str(), size_t(nq), expr.c_str(), trees.c_str());
  if (out.empty()) return 0;
  mkdir(out.c_str(), 0755);
  const size_type npts = m.points_index().last_true() + 1;
  std::vector<double> pts(npts * dim);
  for (size_type p = 0; p < npts; ++p)
    for (int d = 0; d < dim; ++d) pts[p * dim + d] = m.points()[p][d];
  std::vector<int64_t> conn(ne * ng), edu(ne * ndu), edp(ne * ndp);
  for (size_type cv = 0; cv < ne; ++cv) {
    for (size_type i = 0; i < ng; ++i) conn[cv * ng + i] = int64_t(m.ind_points_of_convex(cv)[i]);
    const auto &cu = mfu.ind_scalar_basic_dof_of_element(cv);
    for (size_type i = 0; i < ndu; ++i) edu[cv * ndu + i] = int64_t(cu[i]);
    const auto &cp = mfp.ind_scalar_basic_dof_of_element(cv);
    for (size_type i = 0; i < ndp; ++i) edp[cv * ndp + i] = int64_t(cp[i]);
  }
  f64(out + "/pts.npy", {npts, size_t(dim)}, pts);
  i64(out + "/conn.npy", {ne, ng}, conn);
  i64(out + "/elem_dof_u.npy", {ne, ndu}, edu);
  i64(out + "/elem_dof_p.npy", {ne, ndp}, edp);
  bgeot::pstored_point_tab pspt = pai->pintegration_points();
  bgeot::pgeotrans_precomp pgp = bgeot::geotrans_precomp(pgt, pspt, 0);
  std::vector<double> w(nq), gtg(nq * ng * dim);
  for (size_type q = 0; q < nq; ++q) {
    w[q] = pai->coeff(q);
    const bgeot::base_matrix &pc = pgp->grad(q);
    for (size_type i = 0; i < ng; ++i)
      for (int d = 0; d < dim; ++d) gtg[(q * ng + i) * dim + d] = pc(i, d);
  }
  f64(out + "/quad_w.npy", {nq}, w);
  f64(out + "/gt_grad.npy", {nq, ng, size_t(dim)}, gtg);
  auto tables = [&](getfem::pfem pf, size_type nd, const std::string &sfx) {
    getfem::pfem_precomp pfp2 = getfem::fem_precomp(pf, pspt, 0);
    std::vector<double> phi(nq * nd), gphi(nq * nd * dim);
    for (size_type q = 0; q < nq; ++q) {
      const bgeot::base_tensor &v = pfp2->val(q), &g = pfp2->grad(q);
      for (size_type i = 0; i < nd; ++i) {
        phi[q * nd + i] = v[i];
        for (int d = 0; d < dim; ++d) gphi[(q * nd + i) * dim + d] = g[i + nd * d];
      }
    }
    f64(out + "/phi_" + sfx + ".npy", {nq, nd}, phi);
    f64(out + "/gphi_" + sfx + ".npy", {nq, nd, size_t(dim)}, gphi);
  };
  tables(pfu, ndu, "u");
  tables(pfp, ndp, "p");
  auto dump = [&](const csc &C, size_type ncols, const std::string &name) {
    i64(out + "/" + name + "_jc.npy", {ncols + 1}, C.jc);
    i64(out + "/" + name + "_ir.npy", {C.ir.size()}, C.ir);
    f64(out + "/" + name + "_pr.npy", {C.pr.size()}, C.pr);
  };
  dump(Cup, np_, "Kup");
  dump(Cpu, nu, "Kpu");
  f64(out + "/U.npy", {nu}, U);
  f64(out + "/P.npy", {np_}, P);
  f64(out + "/R.npy", {nu + np_}, R);
  return 0;
}
