// oracle/ref_driver.cc -- TEST INFRASTRUCTURE (never linked into the product).
//
// Drives the UNMODIFIED GetFEM reference (oracle/_ref/libgetfem.so, built from
// /root/reference/src by oracle/Makefile) through its own public API
//   getfem::regular_unit_mesh            (src/getfem_regular_meshes.cc:237-284)
//   getfem::mesh_fem / mesh_im           (src/getfem_mesh_fem.cc, getfem_mesh_im.cc)
//   getfem::ga_workspace::add_expression / assembly
//                                        (src/getfem_generic_assembly_workspace.cc:545-600, 791-936)
//   getfem::model::assembly + bricks     (src/getfem_models.cc:2330, 6102-6136;
//                                         src/getfem_nonlinear_elasticity.cc:2301-2325)
// and dumps, as .npy files, (a) every INPUT the device path needs (node
// coordinates, connectivity, dof table, reference tables at the quadrature
// points) read through the reference's accessors and (b) the reference RESULT
// (tangent as CSC jc/ir/pr via gmm::csc_matrix::init_with, residual vector).
// It is also the CPU baseline timer (mode=time).
//
// usage: gf_ref_driver key=value ...
//   dim=3 n=4 gt=pk|qk k=2 q=3 im=4 | imname="IM_TETRAHEDRON(5)"
//   family=laplace|elast|svk|nh_ciarlet|nh_bonet|mooney_rivlin|mass|source  lambda=1 mu=1 a=1   (mooney_rivlin: C10=lambda C01=mu D1=a)
//   family=nsource (normal source term, getfem_models.cc:4290-4299; boundary regions only)
//   region=all|outer|xmax|zmin|half   (outer faces / faces on x=1 / on z=0 (last coord) / convexes with barycentre x<0.5)
//   family2=.. region2=.. a2=..  family3=.. region3=.. a3=..   further expressions of the SAME workspace (linear
//       families laplace|mass|source|nsource|elast with constants a2/f2/g2/lambda2/mu2 ...): every tree adds into the one
//       K / V, as a model's bricks do
//   coef=fem kd=1   the coefficients (a | lambda, mu | f) are fem data on a classical mesh_fem of degree kd
//                   (ws.add_fem_constant) instead of fixed-size constants: heterogeneous material / distributed load
//   u=smooth|random|zero  out=DIR  mode=dump|time|model|crosscheck|lawcheck  threads=T reps=R
//   mode=crosscheck  the reference's own old-vs-new comparison (tests/test_assembly.cc:279-349, 773-881): the tangent of the
//                    GWFL expression against the LEGACY generic_assembly string of the same form (an independent code path,
//                    getfem_assembling_tensors.cc), tolerance 1e-10 there; families laplace | elast | mass, constant coefficients
//   mode=lawcheck    abstract_hyperelastic_law::test_derivatives (getfem_nonlinear_elasticity.cc:298-347) for the laws behind
//                    the families svk | nh_ciarlet | nh_bonet
#include "getfem/getfem_regular_meshes.h"
#include "getfem/getfem_mesh_fem.h"
#include "getfem/getfem_mesh_im.h"
#include "getfem/getfem_generic_assembly.h"
#include "getfem/getfem_assembling.h"
#include "getfem/getfem_models.h"
#include "getfem/getfem_nonlinear_elasticity.h"
#include "getfem/getfem_omp.h"
#include "getfem/getfem_accumulated_distro.h"
#include "gmm/gmm_kernel.h"
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <fstream>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <sys/stat.h>

using getfem::size_type;
using getfem::scalar_type;
using bgeot::base_node;

// ---------------------------------------------------------------- npy writer
static void write_npy(const std::string &path, const char *descr, size_t itemsize,
                      const std::vector<size_t> &shape, const void *data) {
  std::string hdr = "{'descr': '";
  hdr += descr;
  hdr += "', 'fortran_order': False, 'shape': (";
  size_t tot = 1;
  for (size_t i = 0; i < shape.size(); ++i) {
    hdr += std::to_string(shape[i]);
    if (shape.size() == 1 || i + 1 < shape.size()) hdr += ",";
    if (i + 1 < shape.size()) hdr += " ";
    tot *= shape[i];
  }
  hdr += "), }";
  size_t len = 10 + hdr.size() + 1;
  size_t pad = (64 - len % 64) % 64;
  hdr.append(pad, ' ');
  hdr += "\n";
  std::ofstream f(path, std::ios::binary);
  const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
  f.write((const char *)magic, 8);
  uint16_t hl = (uint16_t)hdr.size();
  f.write((const char *)&hl, 2);
  f.write(hdr.data(), hdr.size());
  f.write((const char *)data, tot * itemsize);
}
static void npy_f64(const std::string &p, const std::vector<size_t> &s, const std::vector<double> &v) {
  write_npy(p, "<f8", 8, s, v.data());
}
static void npy_i64(const std::string &p, const std::vector<size_t> &s, const std::vector<int64_t> &v) {
  write_npy(p, "<i8", 8, s, v.data());
}
static void npy_i32(const std::string &p, const std::vector<size_t> &s, const std::vector<int32_t> &v) {
  write_npy(p, "<i4", 4, s, v.data());
}

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
  std::map<std::string, std::string> a;
  for (int i = 1; i < argc; ++i) {
    std::string s(argv[i]);
    size_t e = s.find('=');
    if (e == std::string::npos) { std::fprintf(stderr, "bad arg %s\n", argv[i]); return 2; }
    a[s.substr(0, e)] = s.substr(e + 1);
  }
  auto geti = [&](const char *k, long d) { return a.count(k) ? std::stol(a[k]) : d; };
  auto getd = [&](const char *k, double d) { return a.count(k) ? std::stod(a[k]) : d; };
  auto gets = [&](const char *k, const char *d) { return a.count(k) ? a[k] : std::string(d); };

  const int dim = (int)geti("dim", 3), n = (int)geti("n", 2), K = (int)geti("k", 1);
  const int Q = (int)geti("q", 1), imdeg = (int)geti("im", 2);
  const int nx = (int)geti("nx", n), ny = (int)geti("ny", n), nz = (int)geti("nz", n);
  const std::string gt = gets("gt", "pk"), family = gets("family", "laplace");
  const std::string umode = gets("u", "random"), out = gets("out", ""), mode = gets("mode", "dump");
  const double lambda = getd("lambda", 1.0), mu = getd("mu", 1.0), acoef = getd("a", 1.0);
  const int threads = (int)geti("threads", 1), reps = (int)geti("reps", 3), warm = (int)geti("warm", 1);
  const double uamp = getd("uamp", 0.02);

  // the thread partition must exist before any per-thread singleton is touched (getfem_omp.h:164)
  if (mode == "omp" || mode == "model") getfem::set_num_threads(threads);

  // ---- mesh / fem / im through the reference's own constructors
  getfem::mesh m;
  std::vector<size_type> ns;
  ns.push_back(nx); if (dim > 1) ns.push_back(ny); if (dim > 2) ns.push_back(nz);
  bgeot::pgeometric_trans pgt =
      gt == "pk" ? bgeot::simplex_geotrans(dim, 1) : bgeot::parallelepiped_geotrans(dim, 1);
  double t0 = now_s();
  getfem::regular_unit_mesh(m, ns, pgt);
  double t_mesh = now_s() - t0;
  // noise=<amp>: every node moved by amp * h * uniform(-1, 1) per direction (seeded), the mesh rebuilt convex by convex with
  // add_convex_by_points like regular_unit_mesh does (src/getfem_regular_meshes.cc:237-284), so the numbering is unchanged:
  // distorted simplices (every element its own K, B, J) and non-affine GT_QK cells (geometry at every Gauss point)
  const double noise = getd("noise", 0.0);
  if (noise > 0) {
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
  getfem::mesh_fem mf(m, getfem::dim_type(Q));
  mf.set_classical_finite_element(getfem::dim_type(K));
  getfem::mesh_im mim(m);
  if (a.count("imname")) mim.set_integration_method(getfem::int_method_descriptor(a["imname"]));
  else mim.set_integration_method(getfem::dim_type(imdeg));
  t0 = now_s();
  const size_type ndof = mf.nb_dof();  // triggers enumerate_dof (src/getfem_mesh_fem.cc:320-446)
  double t_enum = now_s() - t0;
  const size_type ne = m.convex_index().card();
  const size_type cv0 = m.convex_index().first_true();
  getfem::pfem pf = mf.fem_of_element(cv0);
  getfem::pintegration_method pim = mim.int_method_of_element(cv0);
  getfem::papprox_integration pai = pim->approx_method();
  const size_type nq = pai->nb_points_on_convex();
  const size_type nd = pf->nb_dof(cv0), ng = pgt->nb_points();

  // ---- integration region (mesh_region of convexes or of faces, getfem_mesh_region.h)
  const std::string rgname = gets("region", "all");
  getfem::mesh_region rg_all = getfem::mesh_region::all_convexes();
  getfem::mesh_region rg_sel;
  auto make_region = [&](const std::string &rgname, getfem::mesh_region &rg_sel) {
    if (rgname == "half") {
      for (dal::bv_visitor cv(m.convex_index()); !cv.finished(); ++cv) {
        double bx = 0;
        for (size_type i = 0; i < ng; ++i) bx += m.points_of_convex(cv)[i][0];
        if (bx / double(ng) < 0.5) rg_sel.add(cv);
      }
    } else {
      getfem::mesh_region outer;
      getfem::outer_faces_of_mesh(m, outer);
      for (getfem::mr_visitor v(outer); !v.finished(); ++v) {
        base_node un = m.normal_of_face_of_convex(v.cv(), v.f());
        un /= gmm::vect_norm2(un);
        bool keep = rgname == "outer" || (rgname == "xmax" && un[0] > 0.999) ||
                    (rgname == "zmin" && un[dim - 1] < -0.999);
        if (keep) rg_sel.add(v.cv(), v.f());
      }
    }
  };
  if (rgname != "all") make_region(rgname, rg_sel);
  const getfem::mesh_region &rg_use = rgname == "all" ? rg_all : rg_sel;
  // further expressions of the same workspace
  struct extra_term { std::string family, region, expr, sfx; getfem::mesh_region rg; std::vector<double> c_a, c_f, c_g, c_l, c_m; };
  std::vector<std::unique_ptr<extra_term>> extras;
  for (int t = 2; t <= 4; ++t) {
    const std::string sfx = std::to_string(t);
    if (!a.count("family" + sfx)) continue;
    std::unique_ptr<extra_term> e(new extra_term);
    e->family = a["family" + sfx];
    e->region = a.count("region" + sfx) ? a["region" + sfx] : "all";
    e->sfx = sfx;
    if (e->region != "all") make_region(e->region, e->rg);
    const double ac = a.count("a" + sfx) ? std::stod(a["a" + sfx]) : 1.0;
    e->c_a = {ac};
    e->c_l = {a.count("lambda" + sfx) ? std::stod(a["lambda" + sfx]) : 1.0};
    e->c_m = {a.count("mu" + sfx) ? std::stod(a["mu" + sfx]) : 1.0};
    for (int k = 0; k < Q; ++k) e->c_f.push_back(ac * double(k + 1));
    for (int k = 0; k < Q * dim; ++k) e->c_g.push_back(ac * (0.5 + 0.37 * double(k)) * ((k % 3) == 1 ? -1.0 : 1.0));
    if (e->family == "laplace") e->expr = "a" + sfx + "*Grad_u:Grad_Test_u";
    else if (e->family == "mass") e->expr = "a" + sfx + "*u.Test_u";
    else if (e->family == "source") e->expr = Q == 1 ? "-f" + sfx + "*Test_u" : "-f" + sfx + ".Test_u";
    else if (e->family == "nsource")
      e->expr = Q == 1 ? "((g" + sfx + ").Normal)*Test_u" : "(Reshape(g" + sfx + ",qdim(u),meshdim)*Normal).Test_u";
    else if (e->family == "elast")
      e->expr = "(Div_u*((lambda" + sfx + ")*Id(meshdim))+(2*(mu" + sfx + "))*Sym(Grad_u)):Grad_Test_u";
    else { std::fprintf(stderr, "family%s=%s not handled\n", sfx.c_str(), e->family.c_str()); return 2; }
    extras.push_back(std::move(e));
  }

  // ---- expression of the family (the brick strings of the reference)
  std::string expr;
  if (family == "laplace") expr = "a*Grad_u.Grad_Test_u";  // generic elliptic, scalar a
  else if (family == "laplace_vec") expr = "a*Grad_u:Grad_Test_u";
  else if (family == "mass") expr = "a*u.Test_u";
  // volumic source term, the string of add_source_term_brick (getfem_models.cc:4124-, "-(A.Test_u)" shape):
  // order 1 only, no tangent.  f = a * (1, 2, .., Q)
  else if (family == "source") expr = Q == 1 ? "-f*Test_u" : "-f.Test_u";
  // normal source term brick (getfem_models.cc:4290-4299): data g of meshdim (scalar u) or Q x meshdim (vector u)
  else if (family == "nsource")
    expr = Q == 1 ? "((g).Normal)*Test_u" : "(Reshape(g,qdim(u),meshdim)*Normal).Test_u";
  else if (family == "elast")  // src/getfem_models.cc:6112-6113
    expr = "(Div_u*((lambda)*Id(meshdim))+(2*(mu))*Sym(Grad_u)):Grad_Test_u";
  else {
    std::string law = family == "svk" ? "Saint_Venant_Kirchhoff"
                    : family == "nh_ciarlet" ? "Compressible_Neo_Hookean_Ciarlet"
                    : family == "nh_bonet" ? "Compressible_Neo_Hookean_Bonet"
                    : family == "mooney_rivlin" ? "Compressible_Mooney_Rivlin"
                    : family == "ciarlet_geymonat" ? "Ciarlet_Geymonat"
                    : family == "blatz_ko" ? "Generalized_Blatz_Ko" : "";
    if (law.empty()) { std::fprintf(stderr, "unknown family %s\n", family.c_str()); return 2; }
    // src/getfem_nonlinear_elasticity.cc:2319-2320
    expr = "((Id(meshdim)+Grad_u)*(" + law + "_PK2(Grad_u,params))):Grad_Test_u";
  }
  if (a.count("expr")) expr = a["expr"];

  // ---- state vector
  std::vector<double> U(ndof, 0.0);
  if (umode == "random") {
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> d(-1.0, 1.0);
    for (auto &v : U) v = d(rng);
  } else if (umode == "smooth") {
    // u_k(x) = amp * sin(2 pi x_{(k+1) mod dim}) * cos(pi x_k)   
    for (size_type d = 0; d < ndof; ++d) {
      base_node P = mf.point_of_basic_dof(d);
      int k = int(d % Q);
      double xk = P[k % dim], xn = P[(k + 1) % dim];
      U[d] = uamp * std::sin(2 * M_PI * xn) * std::cos(M_PI * xk);
    }
  }
  if (a.count("uscale")) for (auto &v : U) v *= getd("uscale", 1.0);

  // constants are BORROWED by the workspace (generic_assembly.h:277): keep them alive
  const std::vector<double> c_a{acoef}, c_lambda{lambda}, c_mu{mu};
  // the laws' parameter vector: (lambda, mu); Compressible_Mooney_Rivlin takes (C10, C01, D1), passed as lambda= mu= a=
  std::vector<double> c_params = (family == "mooney_rivlin" || family == "ciarlet_geymonat") ? std::vector<double>{lambda, mu, acoef}
                                                                                             : std::vector<double>{lambda, mu};
  if (a.count("params")) {  // params=p0,p1,...: the law's whole parameter vector (Generalized_Blatz_Ko takes five)
    c_params.clear();
    std::stringstream ss(a["params"]);
    for (std::string tok; std::getline(ss, tok, ',');) c_params.push_back(std::stod(tok));
  }
  std::vector<double> c_f(Q);
  for (size_type k = 0; k < size_type(Q); ++k) c_f[k] = acoef * double(k + 1);
  // fem-data coefficients (coef=fem): scalar fields on mf_d, the source term's field on mf_dq (qdim Q)
  const bool coef_fem = gets("coef", "const") == "fem";
  const int kd = (int)geti("kd", 1);
  getfem::mesh_fem mf_d(m, 1), mf_dq(m, getfem::dim_type(Q));
  mf_d.set_classical_finite_element(getfem::dim_type(kd));
  mf_dq.set_classical_finite_element(getfem::dim_type(kd));
  std::vector<double> d_a, d_l, d_m, d_f;
  if (coef_fem) {
    d_a.resize(mf_d.nb_dof()); d_l.resize(mf_d.nb_dof()); d_m.resize(mf_d.nb_dof()); d_f.resize(mf_dq.nb_dof());
    for (size_type d = 0; d < mf_d.nb_dof(); ++d) {
      base_node P = mf_d.point_of_basic_dof(d);
      const double x0 = P[0], x1 = P[1], x2 = dim > 2 ? P[2] : 0.0;
      d_a[d] = acoef * (1.0 + 0.3 * std::sin(1.7 * x0 + 0.9 * x1 + 0.4 * x2));
      d_l[d] = lambda * (1.0 + 0.25 * std::cos(1.1 * x0 - 0.7 * x1 + 0.3 * x2));
      d_m[d] = mu * (1.0 + 0.2 * std::cos(1.3 * x0 - x1));
    }
    for (size_type d = 0; d < mf_dq.nb_dof(); ++d) {
      base_node P = mf_dq.point_of_basic_dof(d);
      d_f[d] = acoef * double(d % Q + 1) * (1.0 + 0.5 * P[0] - 0.25 * P[dim - 1]);
    }
  }
  std::vector<double> c_g(size_t(Q) * dim);
  for (size_t k = 0; k < c_g.size(); ++k) c_g[k] = acoef * (0.5 + 0.37 * double(k)) * ((k % 3) == 1 ? -1.0 : 1.0);
  auto setup_ws = [&](getfem::ga_workspace &ws, const getfem::mesh_region &rg) {
    ws.add_fem_variable("u", mf, gmm::sub_interval(0, ndof), U);
    if (coef_fem && (family == "laplace" || family == "laplace_vec" || family == "mass"))
      ws.add_fem_constant("a", mf_d, d_a);
    else if (coef_fem && family == "source")
      ws.add_fem_constant("f", mf_dq, d_f);
    else if (coef_fem && family == "elast") {
      ws.add_fem_constant("lambda", mf_d, d_l);
      ws.add_fem_constant("mu", mf_d, d_m);
    } else if (family == "laplace" || family == "laplace_vec" || family == "mass")
      ws.add_fixed_size_constant("a", c_a);
    else if (family == "source")
      ws.add_fixed_size_constant("f", c_f);
    else if (family == "nsource")
      ws.add_fixed_size_constant("g", c_g);
    else if (family == "elast") {
      ws.add_fixed_size_constant("lambda", c_lambda);
      ws.add_fixed_size_constant("mu", c_mu);
    } else
      ws.add_fixed_size_constant("params", c_params);
    ws.add_expression(expr, mim, rg);
    for (auto &e : extras) {
      ws.add_fixed_size_constant("a" + e->sfx, e->c_a);
      ws.add_fixed_size_constant("f" + e->sfx, e->c_f);
      ws.add_fixed_size_constant("g" + e->sfx, e->c_g);
      ws.add_fixed_size_constant("lambda" + e->sfx, e->c_l);
      ws.add_fixed_size_constant("mu" + e->sfx, e->c_m);
      if (e->region == "all") ws.add_expression(e->expr, mim, getfem::mesh_region::all_convexes());
      else ws.add_expression(e->expr, mim, e->rg);
    }
  };

  std::printf("{\"dim\": %d, \"ne\": %zu, \"ndof\": %zu, \"nd\": %zu, \"ng\": %zu, \"nq\": %zu, "
              "\"fem\": \"%s\", \"im\": \"%s\", \"t_mesh\": %.4f, \"t_enum\": %.4f",
              dim, ne, ndof, nd, ng, nq, getfem::name_of_fem(pf).c_str(),
              getfem::name_of_int_method(pim).c_str(), t_mesh, t_enum);

  if (mode == "potential") {
    // order 0: ga_workspace::assembly(0) of the POTENTIAL whose first variation is the family's order-1 form
    // (workspace.cc:791-803; the laws' potentials are the registered "<law>_potential" operators, nonlinear_elasticity.cc:2050-2150)
    std::string pexpr;
    if (family == "laplace" || family == "laplace_vec") pexpr = "a*Norm_sqr(Grad_u)/2";
    else if (family == "mass") pexpr = "a*Norm_sqr(u)/2";
    else if (family == "elast") pexpr = "lambda*sqr(Div_u)/2 + mu*Norm_sqr(Sym(Grad_u))";
    else if (family == "source") pexpr = Q == 1 ? "-f*u" : "-f.u";
    else if (family == "nsource") pexpr = Q == 1 ? "((g).Normal)*u" : "(Reshape(g,qdim(u),meshdim)*Normal).u";
    else pexpr = expr.substr(expr.find("*(") + 2, expr.find("_PK2") - expr.find("*(") - 2) + "_potential(Grad_u,params)";
    const std::string keep = expr;
    expr = pexpr;
    extras.clear();
    getfem::ga_workspace ws0;
    setup_ws(ws0, rgname == "all" ? rg_all : rg_sel);
    ws0.assembly(0);
    expr = keep;
    std::printf(", \"potential_expr\": \"%s\", \"potential\": %.17g}\n", pexpr.c_str(), ws0.assembled_potential());
    return 0;
  }
  if (mode == "lawcheck") {
    bgeot::base_vector bp(2);
    bp[0] = lambda; bp[1] = mu;
    bool ok = true;
    std::string what;
    const double hfd = getd("h", 1e-8);  // step of the finite difference (no call site in the reference fixes one)
    try {
      if (family == "svk") getfem::SaintVenant_Kirchhoff_hyperelastic_law().test_derivatives(3, hfd, bp);
      else if (family == "nh_ciarlet") getfem::Neo_Hookean_hyperelastic_law(false).test_derivatives(3, hfd, bp);
      else if (family == "nh_bonet") getfem::Neo_Hookean_hyperelastic_law(true).test_derivatives(3, hfd, bp);
      else if (family == "mooney_rivlin") {
        bgeot::base_vector b3(3);
        b3[0] = lambda; b3[1] = mu; b3[2] = acoef;
        getfem::Mooney_Rivlin_hyperelastic_law(true, false).test_derivatives(3, hfd, b3);
      } else if (family == "ciarlet_geymonat") {
        bgeot::base_vector b3(3);
        b3[0] = lambda; b3[1] = mu; b3[2] = acoef;
        getfem::Ciarlet_Geymonat_hyperelastic_law().test_derivatives(3, hfd, b3);
      } else if (family == "blatz_ko") {
        bgeot::base_vector b5(5);
        b5[0] = 1.0; b5[1] = 1.0; b5[2] = 1.5; b5[3] = -0.5; b5[4] = 1.5;  // the law's own default set (cc:810-815)
        getfem::generalized_Blatz_Ko_hyperelastic_law().test_derivatives(3, hfd, b5);
      }
      else { ok = false; what = "no law for this family"; }
    } catch (const std::exception &ex) { ok = false; what = ex.what(); }
    std::printf(", \"lawcheck\": %s, \"what\": \"%s\"}\n", ok ? "true" : "false", what.c_str());
    return ok ? 0 : 1;
  }
  if (mode == "crosscheck") {
    getfem::ga_workspace ws;
    setup_ws(ws, rg_use);
    getfem::model_real_sparse_matrix Knew(ndof, ndof), Kold(ndof, ndof);
    ws.set_assembled_matrix(Knew);
    ws.assembly(2);
    std::string old;
    std::vector<double> LAMBDA(1, lambda), MU(1, mu);
    if (family == "laplace" || family == "laplace_vec")
      old = Q == 1 ? "M$1(#1,#1)+=sym(comp(Grad(#1).Grad(#1))(:,i,:,i))" : "M$1(#1,#1)+=sym(comp(vGrad(#1).vGrad(#1))(:,i,j,:,i,j))";
    else if (family == "mass")
      old = Q == 1 ? "M$1(#1,#1)+=sym(comp(Base(#1).Base(#1)))" : "M$1(#1,#1)+=sym(comp(vBase(#1).vBase(#1))(:,i,:,i))";
    else if (family == "elast")  // old_asm_stiffness_matrix_for_homogeneous_linear_elasticity, tests/test_assembly.cc:319-336
      old = "lambda=data$1(1); mu=data$2(1); t=comp(vGrad(#1).vGrad(#1));"
            "M(#1,#1)+= sym(t(:,i,j,:,i,j).mu(1)+ t(:,j,i,:,i,j).mu(1)+ t(:,i,i,:,j,j).lambda(1))";
    GMM_ASSERT1(!old.empty(), "crosscheck: families laplace | mass | elast");
    getfem::generic_assembly assem(old);
    assem.push_mi(mim);
    assem.push_mf(mf);
    if (family == "elast") { assem.push_data(LAMBDA); assem.push_data(MU); }
    assem.push_mat(Kold);
    assem.assembly(rg_use);
    if (family != "elast") gmm::scale(Kold, acoef);  // the legacy strings carry no coefficient
    const double nn = gmm::mat_euclidean_norm(Knew);
    gmm::add(gmm::scaled(Knew, -1.0), Kold);
    const double dd = gmm::mat_euclidean_norm(Kold);
    std::printf(", \"cross_rel\": %.3e, \"norm\": %.6e}\n", dd / nn, nn);
    return 0;
  }
  if (mode == "time") {
    // direct ga_workspace path, 1 thread (ws.assembly is single-threaded by design)
    double best2 = 1e300, best1 = 1e300;
    size_type nnz = 0;
    for (int r = 0; r < reps + 1; ++r) {   // first pass = warm-up (precomp caches)
      getfem::ga_workspace ws;
      setup_ws(ws, rg_use);
      getfem::model_real_sparse_matrix Kmat(ndof, ndof);
      ws.set_assembled_matrix(Kmat);
      t0 = now_s(); ws.assembly(2); double t2 = now_s() - t0;
      t0 = now_s(); ws.assembly(1); double t1 = now_s() - t0;
      if (r > 0) { best2 = std::min(best2, t2); best1 = std::min(best1, t1); }
      nnz = gmm::nnz(Kmat);
    }
    std::printf(", \"t_asm2\": %.6f, \"t_asm1\": %.6f, \"nnz\": %zu, \"threads\": 1}\n", best2, best1, nnz);
    return 0;
  }
  if (mode == "omp") {
    // The reference's own OpenMP scheme (src/getfem/getfem_accumulated_distro.h:157-224,
    // src/getfem_models.cc:2686-2722): per-thread workspaces on the thread's slice of the
    // region, per-thread matrix/vector copies, summed afterwards.
    double best = 1e300, tot = 0; size_type nnz = 0;
    for (int r = 0; r < reps + warm; ++r) {
      getfem::model_real_sparse_matrix Kmat(ndof, ndof);
      std::vector<double> R(ndof, 0.0);
      t0 = now_s();
      {
        getfem::accumulated_distro<getfem::model_real_sparse_matrix> Kd(Kmat);
        getfem::accumulated_distro<std::vector<double>> Rd(R);
        GETFEM_OMP_PARALLEL(
          getfem::ga_workspace ws;
          setup_ws(ws, rg_use);
          ws.set_assembled_matrix(Kd);
          ws.assembly(2);
          ws.set_assembled_vector(Rd);
          ws.assembly(1);
        )
      }
      double t = now_s() - t0;
      if (r >= warm) { best = std::min(best, t); tot += t; }
      nnz = gmm::nnz(Kmat);
    }
    std::printf(", \"t_asm21\": %.6f, \"t_asm21_mean\": %.6f, \"nnz\": %zu, \"threads\": %d}\n", best,
                tot / reps, nnz, threads);
    return 0;
  }

  // ---- dump mode: reference result
  getfem::ga_workspace ws;
  setup_ws(ws, rg_use);
  getfem::model_real_sparse_matrix Kmat(ndof, ndof);
  ws.set_assembled_matrix(Kmat);
  t0 = now_s(); ws.assembly(2); double t2 = now_s() - t0;
  t0 = now_s(); ws.assembly(1); double t1 = now_s() - t0;
  std::vector<double> R(ws.assembled_vector().begin(), ws.assembled_vector().end());
  gmm::csc_matrix<double> C;
  C.init_with(Kmat);
  const size_type nnz = C.jc[ndof];
  std::printf(", \"t_asm2\": %.6f, \"t_asm1\": %.6f, \"nnz\": %zu, \"expr\": \"%s\"}\n", t2, t1, nnz, expr.c_str());
  if (out.empty()) return 0;
  mkdir(out.c_str(), 0755);

  // mesh: points in point-id order, connectivity in convex order
  GMM_ASSERT1(m.convex_index().card() == m.convex_index().last_true() + 1, "non-contiguous convex ids");
  const size_type npts = m.points_index().last_true() + 1;
  std::vector<double> pts(npts * dim);
  for (size_type p = 0; p < npts; ++p)
    for (int d = 0; d < dim; ++d) pts[p * dim + d] = m.points()[p][d];
  std::vector<int32_t> conn(ne * ng);
  std::vector<int64_t> edof(ne * nd);
  for (size_type cv = 0; cv < ne; ++cv) {
    for (size_type i = 0; i < ng; ++i) conn[cv * ng + i] = int32_t(m.ind_points_of_convex(cv)[i]);
    const auto &ct = mf.ind_scalar_basic_dof_of_element(cv);
    for (size_type i = 0; i < nd; ++i) edof[cv * nd + i] = int64_t(ct[i]);
  }
  npy_f64(out + "/pts.npy", {npts, size_t(dim)}, pts);
  npy_i32(out + "/conn.npy", {ne, ng}, conn);
  npy_i64(out + "/elem_dof.npy", {ne, nd}, edof);
  std::vector<double> dxyz(ndof * dim);
  for (size_type d = 0; d < ndof; ++d) {
    base_node P = mf.point_of_basic_dof(d);
    for (int c = 0; c < dim; ++c) dxyz[d * dim + c] = P[c];
  }
  npy_f64(out + "/dof_xyz.npy", {ndof, size_t(dim)}, dxyz);

  // reference tables at the volume quadrature points
  bgeot::pstored_point_tab pspt = pai->pintegration_points();
  getfem::pfem_precomp pfp = getfem::fem_precomp(pf, pspt, 0);
  bgeot::pgeotrans_precomp pgp = bgeot::geotrans_precomp(pgt, pspt, 0);
  std::vector<double> w(nq), xq(nq * dim), gtg(nq * ng * dim), phi(nq * nd), gphi(nq * nd * dim);
  for (size_type q = 0; q < nq; ++q) {
    w[q] = pai->coeff(q);
    for (int d = 0; d < dim; ++d) xq[q * dim + d] = (*pspt)[q][d];
    const bgeot::base_matrix &pc = pgp->grad(q);  // ng x P
    for (size_type i = 0; i < ng; ++i)
      for (int d = 0; d < dim; ++d) gtg[(q * ng + i) * dim + d] = pc(i, d);
    const bgeot::base_tensor &v = pfp->val(q);    // (nd, target_dim)
    const bgeot::base_tensor &g = pfp->grad(q);   // (nd, target_dim, P) first index fastest
    for (size_type i = 0; i < nd; ++i) {
      phi[q * nd + i] = v[i];
      for (int d = 0; d < dim; ++d) gphi[(q * nd + i) * dim + d] = g[i + nd * d];
    }
  }
  npy_f64(out + "/quad_w.npy", {nq}, w);
  npy_f64(out + "/quad_x.npy", {nq, size_t(dim)}, xq);
  npy_f64(out + "/gt_grad.npy", {nq, ng, size_t(dim)}, gtg);
  npy_f64(out + "/phi.npy", {nq, nd}, phi);
  npy_f64(out + "/gphi.npy", {nq, nd, size_t(dim)}, gphi);
  // reference-element node coordinates of the fem
  std::vector<double> rnodes(nd * dim);
  for (size_type i = 0; i < nd; ++i)
    for (int d = 0; d < dim; ++d) rnodes[i * dim + d] = pf->node_of_dof(cv0, i)[d];
  npy_f64(out + "/ref_nodes.npy", {nd, size_t(dim)}, rnodes);

  // region items in mr_visitor order (the order ga_exec walks them, C&E.cc:8789): face -1 = the whole convex
  if (rgname != "all") {
    std::vector<int32_t> icv, ifc;
    for (getfem::mr_visitor v(rg_use, m); !v.finished(); ++v) {
      icv.push_back(int32_t(v.cv()));
      ifc.push_back(v.f() == getfem::short_type(-1) ? -1 : int32_t(v.f()));
    }
    npy_i32(out + "/items_cv.npy", {icv.size()}, icv);
    npy_i32(out + "/items_f.npy", {ifc.size()}, ifc);
  }
  if (coef_fem) {  // the data fems: dof tables, basis values at ALL integration points, nodal values
    const getfem::mesh_fem &mfx = family == "source" ? mf_dq : mf_d;
    getfem::pfem pfd = mfx.fem_of_element(cv0);
    const size_type ndd = pfd->nb_dof(cv0), nqa = pai->nb_points();
    std::vector<int64_t> ded(ne * ndd);
    for (size_type cv = 0; cv < ne; ++cv) {
      const auto &ct = mfx.ind_scalar_basic_dof_of_element(cv);
      for (size_type i = 0; i < ndd; ++i) ded[cv * ndd + i] = int64_t(ct[i]);
    }
    getfem::pfem_precomp pfpd = getfem::fem_precomp(pfd, pai->pintegration_points(), 0);
    std::vector<double> dphi(nqa * ndd);
    for (size_type q = 0; q < nqa; ++q)
      for (size_type i = 0; i < ndd; ++i) dphi[q * ndd + i] = pfpd->val(q)[i];
    npy_i64(out + "/d_elem_dof.npy", {ne, ndd}, ded);
    npy_f64(out + "/d_phi.npy", {nqa, ndd}, dphi);
    if (family == "source") npy_f64(out + "/d_vals0.npy", {d_f.size()}, d_f);
    else if (family == "elast") { npy_f64(out + "/d_vals0.npy", {d_l.size()}, d_l); npy_f64(out + "/d_vals1.npy", {d_m.size()}, d_m); }
    else npy_f64(out + "/d_vals0.npy", {d_a.size()}, d_a);
  }
  for (auto &e : extras) {
    if (e->region != "all") {
      std::vector<int32_t> icv, ifc;
      for (getfem::mr_visitor v(e->rg, m); !v.finished(); ++v) {
        icv.push_back(int32_t(v.cv()));
        ifc.push_back(v.f() == getfem::short_type(-1) ? -1 : int32_t(v.f()));
      }
      npy_i32(out + "/items" + e->sfx + "_cv.npy", {icv.size()}, icv);
      npy_i32(out + "/items" + e->sfx + "_f.npy", {ifc.size()}, ifc);
    }
    npy_f64(out + "/gdata" + e->sfx + ".npy", {e->c_g.size()}, e->c_g);
  }
  { // tables at ALL integration points (volume points first, then the points of face 0, 1, ...:
    // approx_integration::valid_method, getfem_integration.cc:353-368) + reference normals of the faces
    const size_type nqa = pai->nb_points(), nf = pgt->structure()->nb_faces();
    std::vector<double> aw(nqa), ax(nqa * dim), agtg(nqa * ng * dim), aphi(nqa * nd), agphi(nqa * nd * dim);
    for (size_type q = 0; q < nqa; ++q) {
      aw[q] = pai->coeff(q);
      for (int d = 0; d < dim; ++d) ax[q * dim + d] = (*pspt)[q][d];
      const bgeot::base_matrix &pc = pgp->grad(q);
      for (size_type i = 0; i < ng; ++i)
        for (int d = 0; d < dim; ++d) agtg[(q * ng + i) * dim + d] = pc(i, d);
      const bgeot::base_tensor &v = pfp->val(q);
      const bgeot::base_tensor &g = pfp->grad(q);
      for (size_type i = 0; i < nd; ++i) {
        aphi[q * nd + i] = v[i];
        for (int d = 0; d < dim; ++d) agphi[(q * nd + i) * dim + d] = g[i + nd * d];
      }
    }
    std::vector<int32_t> ff(nf), fn(nf);
    std::vector<double> rn(nf * dim);
    for (size_type f = 0; f < nf; ++f) {
      ff[f] = int32_t(pai->ind_first_point_on_face(getfem::short_type(f)));
      fn[f] = int32_t(pai->nb_points_on_face(getfem::short_type(f)));
      for (int d = 0; d < dim; ++d) rn[f * dim + d] = pgt->normals()[f][d];
    }
    npy_f64(out + "/all_w.npy", {nqa}, aw);
    npy_f64(out + "/all_x.npy", {nqa, size_t(dim)}, ax);
    npy_f64(out + "/all_gt_grad.npy", {nqa, ng, size_t(dim)}, agtg);
    npy_f64(out + "/all_phi.npy", {nqa, nd}, aphi);
    npy_f64(out + "/all_gphi.npy", {nqa, nd, size_t(dim)}, agphi);
    npy_i32(out + "/face_first.npy", {nf}, ff);
    npy_i32(out + "/face_nq.npy", {nf}, fn);
    npy_f64(out + "/ref_normals.npy", {nf, size_t(dim)}, rn);
  }

  npy_f64(out + "/U.npy", {ndof}, U);
  std::vector<int64_t> jc(ndof + 1), ir(nnz);
  std::vector<double> pr(nnz);
  for (size_type j = 0; j <= ndof; ++j) jc[j] = C.jc[j];
  for (size_type k = 0; k < nnz; ++k) { ir[k] = C.ir[k]; pr[k] = C.pr[k]; }
  npy_i64(out + "/K_jc.npy", {ndof + 1}, jc);
  npy_i64(out + "/K_ir.npy", {nnz}, ir);
  npy_f64(out + "/K_pr.npy", {nnz}, pr);
  npy_f64(out + "/R.npy", {ndof}, R);
  std::vector<double> par = {lambda, mu, acoef};
  npy_f64(out + "/params.npy", {3}, par);
  npy_f64(out + "/lawparams.npy", {c_params.size()}, c_params);
  npy_f64(out + "/gdata.npy", {c_g.size()}, c_g);
  return 0;
}
