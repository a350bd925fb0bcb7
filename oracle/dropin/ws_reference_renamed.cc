// oracle/dropin/ws_reference_renamed.cc -- TEST INFRASTRUCTURE (drop-in demonstration, never part of the product).
//
// Compiles the UNMODIFIED reference translation unit src/getfem_generic_assembly_workspace.cc from where it lies, with
// ONE token renamed by the preprocessor: ga_workspace::assembly becomes ga_workspace::assembly_reference.  Nothing of the
// reference is copied or edited.  The symbol ga_workspace::assembly itself is then provided by ws_dispatch.cc, which is
// the patch INTEGRATION.md section 2 describes (dispatch to the device path, else the reference's own code).
#define assembly assembly_reference
#include "getfem_generic_assembly_workspace.cc"  // found through -I$(REF)/src
#undef assembly

namespace getfem_b200 {
void reference_assembly(getfem::ga_workspace &ws, getfem::size_type order, bool condensation) {
  ws.assembly_reference(order, condensation);
}
}  // namespace getfem_b200
