// gfgpu_getfem_shim.h -- drop-in device path for getfem::ga_workspace::assembly().
//
// Compiled against the UNMODIFIED GetFEM headers (it is the code a GetFEM maintainer would add next
// to src/getfem_generic_assembly_workspace.cc).  Everything it needs from GetFEM
// goes through public accessors; everything it asks of the GPU goes through the C ABI of
// include/gfgpu.h.  No CPU fallback: an expression, fem or integration method the device path does
// not cover raises gmm::gmm_error (GMM_ASSERT1), exactly like any other unsupported GWFL construct.
#ifndef GFGPU_GETFEM_SHIM_H
#define GFGPU_GETFEM_SHIM_H

#include <map>
#include <memory>
#include <string>

#include "getfem/getfem_generic_assembly.h"
#include "gfgpu.h"

namespace getfem_b200 {

struct recognised_term {
  int family;             // GFGPU_LAPLACE ...
  std::string varname;    // the fem variable (Test and Test2)
  std::vector<double> params;
  // fem-data coefficients (ga_workspace::add_fem_constant): names of the leading parameters that are fields on a data
  // mesh_fem; `field_sign` multiplies their nodal values (the source term's "-f")
  std::vector<std::string> field_names;
  double field_sign = 1.0;
  bool by_probe = false;  // identified numerically (recognise_by_probe), not from a printed normal form
  // assembly(2) on a workspace whose expression was added WITHOUT derivative trees (add_expression(.., order 1)): the reference
  // assembles the order-2 trees that exist and nothing else, so this order-1 term must not contribute its family's tangent
  bool no_tangent = false;
  // COUPLED div-pressure parts (family GFGPU_SHIM_COUPLED_DIV, a shim-level tag): the block B(row dof of the vector variable
  // `varname_u`, column dof of the scalar variable `varname_p`) = int psi_p div(phi_u), with factor `sign`.
  //   order 2: `transposed` = false for the tree (Test_u, Test2_p), true for (Test_p, Test2_u);
  //   order 1: `transposed` = false for the residual part of u, sign * B p; true for the part of p, sign * B^T u.
  std::string varname_u, varname_p;
  bool transposed = false;
  double sign = 1.0;
  // JIT terms (family GFGPU_JIT, the NVRTC route): the order-1 / order-2 trees translated into C expressions
  // (include/gfgpu.h, gfgpu_term_create_jit) and the names of the scalar constants behind par[k]
  std::string jit_form1, jit_form2;
  std::string jit_form0;  // the order-0 tree on the same (integration method, region), when there is one and it translates
  std::vector<std::string> jit_params;
};
// GFGPU_SHIM_COUPLED_MASS: "Test_a:Test2_b" on two variables of the same qdim (asm_mass_matrix(M, mim, mf1, mf2, rg), the
// constraint matrix of the Dirichlet bricks with multipliers): rows = varname_u (Test), columns = varname_p (Test2), factor sign
enum { GFGPU_SHIM_COUPLED_DIV = 1000, GFGPU_SHIM_COUPLED_MASS = 1001 };

// Matches the ORDER-1 tree of `ws` number `itree` (as printed by ga_tree_to_string after the
// reference's semantic analysis) against the families of include/gfgpu.h.  Returns false if unknown.
bool recognise_tree(const getfem::ga_workspace &ws, getfem::size_type itree, recognised_term &out);
// ga_workspace::add_tree SUMS the expressions that share (mim, region, test variables) into one tree
// (getfem_generic_assembly_workspace.cc:472-493): this splits such a tree "(A)+(B)" into its recognised summands.
// At most one summand may carry a tangent: the reference thresholds the element matrix of the SUM
// (C&E.cc:4889,4898), which two separately assembled bilinear forms would not reproduce.
bool recognise_tree_sum(const getfem::ga_workspace &ws, getfem::size_type itree, std::vector<recognised_term> &out);

// Recognition by probe (see the .cc) assembles a few normal forms on two convexes with the reference's interpreter.  Inside a
// library whose ga_workspace::assembly IS the dispatch patch, give it the body of the reference here (the dispatch patch of INTEGRATION.md section 2 does).
void set_reference_assembly(void (*f)(getfem::ga_workspace &, getfem::size_type));

// Device-side state of one (mesh, mesh_fem, mesh_im, term); reusable across Newton iterations.
class device_assembler {
 public:
  explicit device_assembler(int device = 0);
  ~device_assembler();
  device_assembler(const device_assembler &) = delete;
  device_assembler &operator=(const device_assembler &) = delete;

  // Same contract as getfem::ga_workspace::assembly(order) for order 1 and 2
  // (src/getfem_generic_assembly_workspace.cc:791-936): results are ADDED into
  // ws.assembled_vector() / ws.assembled_matrix(), which are sized first if needed.
  void assembly(getfem::ga_workspace &ws, getfem::size_type order);

  // seconds spent in the last call: extraction of GetFEM data, device work, fill of the gmm containers
  double t_extract = 0, t_device = 0, t_fill = 0;
  // how many times the pattern (jc, ir) of the workspace tangent had to be downloaded: 1 for a Newton / time loop whose
  // pattern does not move (the values alone travel afterwards)
  long pattern_downloads = 0;

 private:
  struct entry;
  struct rect_entry;
  struct grouping;
  struct tangent_cache;
  std::unique_ptr<tangent_cache> tangent_;  // the workspace tangent, resident on the device between calls
  gfgpu_ctx *ctx_ = nullptr;
  std::map<std::string, std::unique_ptr<entry>> cache_;  // bounded (LRU); entries die with the getfem objects they mirror
  std::map<std::string, std::unique_ptr<rect_entry>> rect_cache_;  // coupled terms: two fems on one mesh
  rect_entry &coupled_entry(getfem::ga_workspace &ws, const getfem::mesh_im &mim, const std::string &vu, const std::string &vp,
                            int family = 0, const std::vector<int32_t> *rg_cv = nullptr, const std::vector<int32_t> *rg_f = nullptr);
  struct reduction_entry;
  std::map<const void *, std::unique_ptr<reduction_entry>> reductions_;  // extension matrices of reduced mesh_fems
  gfgpu_reduction *reduction_of(const getfem::mesh_fem &mf);  // nullptr for a non-reduced mesh_fem
  std::map<std::string, std::pair<bool, std::vector<recognised_term>>> recognised_;  // recognition results across calls
  std::map<std::string, std::unique_ptr<grouping>> groupings_;  // convex groups per (mesh, mesh_fem, mesh_im)
  uint64_t use_clock_ = 0;
};

// Adds a CSC matrix (jc[ncols+1], ir, pr: gmm::csc_matrix layout, rows ascending inside a column) into K, column by column
// (OpenMP over the columns when available); K must already have at least ncols columns.
void fill_col_matrix(getfem::model_real_sparse_matrix &K, getfem::size_type ncols, const int64_t *jc, const int32_t *ir,
                     const double *pr);

// One-shot convenience: getfem_b200::assembly(ws, 2) instead of ws.assembly(2).
void assembly(getfem::ga_workspace &ws, getfem::size_type order, int device = 0);

}  // namespace getfem_b200
#endif
