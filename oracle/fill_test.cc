// oracle/fill_test.cc -- TEST INFRASTRUCTURE (CPU only, no device call).  getfem_b200::fill_col_matrix, the host end of the
// drop-in (device CSC -> gmm::col_matrix<rsvector>, OpenMP over the columns), against the reference's own containers: the
// filled matrix must be what gmm::add of the two CSC operands gives, for a fresh K and for a K that already holds entries
// (ga_workspace::assembly ACCUMULATES into an aliased matrix, getfem_generic_assembly_workspace.cc:805-812).
#include <chrono>
#include <cstdio>
#include <random>

#include "getfem/getfem_generic_assembly.h"
#include "gfgpu_getfem_shim.h"
#include "gmm/gmm_kernel.h"

using getfem::size_type;

struct csc { std::vector<int64_t> jc; std::vector<int32_t> ir; std::vector<double> pr; };

static csc random_csc(size_type n, unsigned seed, int maxcol) {
  std::mt19937_64 rng(seed);
  csc c;
  c.jc.assign(n + 1, 0);
  for (size_type j = 0; j < n; ++j) {
    const int cnt = int(rng() % (maxcol + 1));  // empty columns included
    std::vector<int32_t> rows;
    for (int k = 0; k < cnt; ++k) rows.push_back(int32_t(rng() % n));
    std::sort(rows.begin(), rows.end());
    rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
    for (int32_t r : rows) { c.ir.push_back(r); c.pr.push_back((double(int64_t(rng() % 2001) - 1000) + 0.5) / 64.0); }  // never 0.0: rsvector::w drops zeros
    c.jc[j + 1] = int64_t(c.ir.size());
  }
  return c;
}

static void add_serial(getfem::model_real_sparse_matrix &K, const csc &c) {  // the plain gmm way
  for (size_type j = 0; j + 1 < c.jc.size(); ++j)
    for (int64_t k = c.jc[j]; k < c.jc[j + 1]; ++k) K(size_type(c.ir[size_t(k)]), j) += c.pr[size_t(k)];
}

int main(int argc, char **argv) {
  const size_type n = argc > 1 ? size_type(std::atol(argv[1])) : 20000;
  const csc a = random_csc(n, 1, 40), b = random_csc(n, 2, 25);
  getfem::model_real_sparse_matrix K(n, n), R(n, n);
  const auto c0 = std::chrono::steady_clock::now();
  getfem_b200::fill_col_matrix(K, n, a.jc.data(), a.ir.data(), a.pr.data());  // fresh columns
  const double t_fresh = std::chrono::duration<double>(std::chrono::steady_clock::now() - c0).count();
  getfem_b200::fill_col_matrix(K, n, b.jc.data(), b.ir.data(), b.pr.data());  // accumulation into stored columns
  add_serial(R, a);
  add_serial(R, b);
  // identical structure is not required where an accumulated value cancels to 0.0 (rsvector keeps it on both sides here:
  // w() stores zeros), so compare entry by entry through CSC copies
  gmm::csc_matrix<double> CK, CR;
  CK.init_with(K);
  CR.init_with(R);
  bool same = CK.jc.size() == CR.jc.size() && CK.ir.size() == CR.ir.size();
  for (size_t k = 0; same && k < CK.jc.size(); ++k) same = CK.jc[k] == CR.jc[k];
  for (size_t k = 0; same && k < CK.ir.size(); ++k) same = CK.ir[k] == CR.ir[k] && CK.pr[k] == CR.pr[k];
  bool sorted = true;  // rsvector invariant: rows strictly ascending inside every column
  for (size_type j = 0; j < n && sorted; ++j) {
    const gmm::rsvector<double> &col = K[j];
    for (auto it = col.begin(); it != col.end() && sorted; ++it)
      if (it + 1 != col.end()) sorted = it->c < (it + 1)->c;
  }
  std::printf("{\"n\": %zu, \"nnz\": %zu, \"same\": %s, \"sorted\": %s, \"fresh_fill_s\": %.4f, \"fresh_nnz\": %zu}\n", size_t(n),
              CK.ir.size(), same ? "true" : "false", sorted ? "true" : "false", t_fresh, a.ir.size());
  return same && sorted ? 0 : 1;
}
