// oracle/dropin/ws_dispatch.cc -- TEST INFRASTRUCTURE: the reference-side patch of INTEGRATION.md section 2, as real code.
//
// Defines ga_workspace::assembly (getfem_generic_assembly_workspace.cc:791 in the reference): when the device path is
// enabled it prepares K / V exactly as the reference does before ga_exec (workspace.cc:803-848: an OWNED matrix / vector is
// cleared and resized to nb_prim_dof, an aliased one is trusted and accumulated into) and hands the workspace to the shim;
// otherwise it calls the reference's own implementation (renamed by ws_reference_renamed.cc).  With this library in place
// of libgetfem.so, model::assembly, the bricks and every asm_* wrapper run on the GPU unchanged.
#include <atomic>
#include <cstdio>
#include <cstdlib>

#include "getfem/getfem_generic_assembly.h"
#include "getfem/getfem_generic_assembly_tree.h"
#include "getfem/getfem_models.h"
#include "getfem/getfem_omp.h"
#include "gfgpu_getfem_shim.h"

namespace getfem_b200 {
void reference_assembly(getfem::ga_workspace &ws, getfem::size_type order, bool condensation);
static bool g_enabled = false;
static long g_device_calls = 0, g_reference_calls = 0;
static std::atomic<long> g_skipped_calls{0};  // threads other than 0 of a sliced assembly (they contribute zeros)
static void reference_order(getfem::ga_workspace &ws, getfem::size_type order) { reference_assembly(ws, order, false); }
void gfgpu_enable(bool on) {
  g_enabled = on;
  set_reference_assembly(&reference_order);  // the probe of the shim's recogniser must not re-enter the dispatch
}
long gfgpu_device_calls() { return g_device_calls; }
long gfgpu_reference_calls() { return g_reference_calls; }
long gfgpu_skipped_calls() { return g_skipped_calls.load(); }
static device_assembler &thread_assembler() {
  static thread_local device_assembler dev(0);
  return dev;
}
// seconds of the calling thread's last device call: extraction of the GetFEM data, device work, fill of the gmm containers;
// and how many times it had to download the pattern of the workspace tangent
void gfgpu_last_timings(double *t3) {
  device_assembler &d = thread_assembler();
  t3[0] = d.t_extract; t3[1] = d.t_device; t3[2] = d.t_fill;
}
long gfgpu_pattern_downloads() { return thread_assembler().pattern_downloads; }
}  // namespace getfem_b200

namespace getfem {
void ga_workspace::assembly(size_type order, bool condensation) {
  // inside GETFEM_OMP_PARALLEL with SEVERAL partitions every thread assembles its slice of the region into a private
  // copy (getfem_accumulated_distro.h:157-224).  With one partition (the library default, partition_master's constructor calls set_num_threads(1), getfem_omp.cc:236)
  // the slice is the whole region although me_is_multithreaded_now() is true inside the bricks' parallel blocks.
  const bool sliced = getfem::me_is_multithreaded_now() && getfem::partition_master::get().get_nb_partitions() > 1;
  if (!getfem_b200::g_enabled || condensation || order > 2) {
    ++getfem_b200::g_reference_calls;
    getfem_b200::reference_assembly(*this, order, condensation);
    return;
  }
  if (std::getenv("GFGPU_DRYRUN")) {  // CPU debugging aid: what would be sent to the device, then the reference path
    for (size_type i = 0; i < nb_trees(); ++i) {
      const tree_description &td = tree_info(i);
      std::vector<getfem_b200::recognised_term> rts;
      bool ok = false;
      std::string why;
      try { ok = td.order >= 1 ? getfem_b200::recognise_tree_sum(*this, i, rts) : true; } catch (const std::exception &ex) { why = ex.what(); }
      std::string fam;
      for (const auto &rt : rts) {
        fam += " family " + std::to_string(rt.family) + " (";
        for (double p : rt.params) { char b[40]; std::snprintf(b, sizeof b, " %.12g", p); fam += b; }
        fam += " )";
      }
      std::fprintf(stderr, "[gfgpu dryrun] order %d (assembly order %d) region %ld: %s -> %s%s %s\n", int(td.order), int(order),
                   long(td.rg->id()), ga_tree_to_string(*td.ptree).c_str(), ok ? "recognised" : "NOT recognised", fam.c_str(),
                   why.c_str());
      for (const auto &rt : rts)  // the translated forms of a run-time compiled term: tests/test_shim_probe.py compiles them (NVRTC, no GPU)
        if (ok && order == 2 && rt.family == GFGPU_JIT && associated_mf(rt.varname))
          std::fprintf(stderr, "[gfgpu dryrun jit] dim=%d qdim=%d\t%s\t%s\t%s\n", int(associated_mf(rt.varname)->linked_mesh().dim()),
                       int(associated_mf(rt.varname)->get_qdim()), rt.jit_form1.c_str(), rt.jit_form2.c_str(), rt.jit_form0.c_str());
    }
    getfem_b200::reference_assembly(*this, order, condensation);
    return;
  }
  const ga_workspace *w = this;
  while (w->parent_workspace) w = w->parent_workspace;
  if (w->md) w->md->nb_dof();  // actualize_sizes, as the reference does first
  if (order == 2 && K.use_count()) {
    gmm::clear(*K);
    gmm::resize(*K, nb_prim_dof, nb_prim_dof);
  }
  if (order == 0) {  // the scalar result: a one-entry tensor set to zero (C&E.cc:8067, 8733)
    assemb_t.adjust_sizes(1);
    assemb_t[0] = 0.0;
  }
  if (order == 1) {
    if (V.use_count()) {
      gmm::clear(*V);
      gmm::resize(*V, nb_prim_dof);
    } else
      GMM_ASSERT1(V->size() == nb_prim_dof, "Wrong size of assembled vector in workspace");
  }
  if (sliced && getfem::partition_master::get().get_current_partition() != 0) {
    // under several OpenMP partitions every thread of a brick's GETFEM_OMP_PARALLEL block holds a private,
    // zero-initialised copy of the result (accumulated_distro::get(), getfem_accumulated_distro.h:157-224) that the block
    // sums afterwards.  Thread 0 assembles the whole region on the device (the shim walks the region with partitioning
    // prohibited); the other threads contribute their zeros.
    ++getfem_b200::g_skipped_calls;
    if (std::getenv("GFGPU_TRACE"))
      std::fprintf(stderr, "[gfgpu trace] skipped order %d partition %d\n", int(order),
                   int(getfem::partition_master::get().get_current_partition()));
    return;
  }
  getfem_b200::device_assembler &dev = getfem_b200::thread_assembler();
  ++getfem_b200::g_device_calls;
  dev.assembly(*this, order);  // throws gmm::gmm_error when the workspace is not covered: no silent CPU fallback
  if (std::getenv("GFGPU_TRACE")) {
    double nk = 0, nv = 0;
    if (order == 2)
      for (size_type j = 0; j < gmm::mat_ncols(*K); ++j)
        for (auto it = (*K)[j].begin(); it != (*K)[j].end(); ++it) nk += it->e * it->e;
    if (order == 1) for (double x : *V) nv += x * x;
    std::fprintf(stderr, "[gfgpu trace] device call order %d partition %d/%d trees %d |K|^2 %.17g |V|^2 %.17g owned K %d V %d\n",
                 int(order), int(getfem::partition_master::get().get_current_partition()),
                 int(getfem::partition_master::get().get_nb_partitions()), int(nb_trees()), nk, nv, int(K.use_count()),
                 int(V.use_count()));
  }
}
}  // namespace getfem
