// common.cuh -- shared declarations of libgfgpu (sm_100a only; no CPU fallback).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gfgpu.h"

namespace gf {

struct Error : std::runtime_error {
  explicit Error(const std::string &s) : std::runtime_error(s) {}
};

#define GF_STR2(x) #x
#define GF_STR(x) GF_STR2(x)
#define GF_REQUIRE(cond, msg)                                                        \
  do {                                                                               \
    if (!(cond)) throw gf::Error(std::string(__FILE__ ":" GF_STR(__LINE__) ": ") + (msg)); \
  } while (0)
#define GF_CUDA(call)                                                                \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess)                                                           \
      throw gf::Error(std::string(__FILE__ ":" GF_STR(__LINE__) ": " #call ": ") + cudaGetErrorString(e_)); \
  } while (0)

extern std::atomic<int64_t> g_launches;  // kernels launched by this process (ours + CUB passes we call)
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
#define GF_LAUNCH_CHECK()             \
  do {                                \
    gf::count_launch();               \
    GF_CUDA(cudaGetLastError());      \
  } while (0)

}  // namespace gf

struct gfgpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  int64_t bytes = 0;
  void *cub_tmp = nullptr;  // grow-only scratch for CUB calls
  size_t cub_tmp_bytes = 0;
  // side stream: the residual path of strategy RECOMPUTE runs next to the tile kernel (fork / join through these events)
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

namespace gf {

// Device buffer tied to a context (cudaMalloc; sizes are large and long-lived, no pooling needed).
template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  gfgpu_ctx *ctx = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void alloc(gfgpu_ctx *c, size_t count) {
    release();
    ctx = c;
    n = count;
    if (count) {
      GF_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
      ctx->bytes += (int64_t)(count * sizeof(T));
    }
  }
  void release() {
    if (p) {
      cudaFree(p);
      if (ctx) ctx->bytes -= (int64_t)(n * sizeof(T));
    }
    p = nullptr;
    n = 0;
  }
  void zero() {
    if (n) GF_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), ctx->stream));
  }
  void upload(const T *h) {
    if (n) GF_CUDA(cudaMemcpyAsync(p, h, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  }
  void download(T *h) const {
    if (n) GF_CUDA(cudaMemcpyAsync(h, p, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
  }
};

void *cub_scratch(gfgpu_ctx *ctx, size_t bytes);

}  // namespace gf

struct gfgpu_mesh {
  gfgpu_ctx *ctx;
  int dim, ng, gt_kind;
  int64_t npts, ne;
  gf::DevBuf<double> xyz;    // SoA: x[npts] y[npts] z[npts]
  gf::DevBuf<int32_t> conn;  // ne x ng
};

namespace gf {
// Structural (value-independent) part of the scatter: node pairs and their contributions.
struct Structure {
  int64_t e0 = 0, e1 = 0;  // element range it was built for
  int64_t ncontrib = 0;    // (e1-e0) * nd * nd : LOCAL contributions; ids >= ncontrib in csrc are virtual (halo)
  int64_t nvirt = 0;       // virtual contributions: pairs announced by other ranks for columns owned here
  int64_t npairs = 0;
  int64_t ncolnodes = 0;
  DevBuf<int32_t> pI, pJ;      // dof0 of row / column node of each pair (pairs sorted by (J, I))
  DevBuf<uint32_t> cstart;     // npairs+1: contributions of pair p are csrc[cstart[p] .. cstart[p+1])
  DevBuf<uint32_t> csrc;       // contribution id = (e-e0)*nd*nd + j*nd + i, ascending inside a pair
  DevBuf<uint32_t> colstart;   // ncolnodes+1: pairs of column node k are [colstart[k], colstart[k+1])
  // residual: node -> (element, local node) incidences
  int64_t nrnodes = 0, nrinc = 0;
  DevBuf<int32_t> rdof;        // dof0 of each touched node
  DevBuf<uint32_t> rstart;     // nrnodes+1
  DevBuf<uint32_t> rsrc;       // incidence id = (e-e0)*nd + i
};
}  // namespace gf

struct gfgpu_fem {
  gfgpu_ctx *ctx;
  gfgpu_mesh *mesh;
  int fem_kind, degree, qdim, nd;
  int64_t ndof;
  gf::DevBuf<int32_t> edof;  // ne x nd : dof of component 0
};

struct gfgpu_tables {
  gfgpu_ctx *ctx;
  int dim, nq, ng, nd;
  gf::DevBuf<double> w, gt_grad, phi, gphi;
  gf::DevBuf<double> gt_val;  // nq x ng shape values of the geometric transformation (gfgpu_tables_set_gt_values): the position X
  gf::DevBuf<double> fgt_val;  // nf x nqf x ng: the same at the face points
  std::vector<double> h_w, h_gt_grad, h_phi, h_gphi;
  // face points (gfgpu_tables_set_faces): nf faces with nqf points each, tables laid out face after face
  int nf = 0, nqf = 0;
  gf::DevBuf<double> fw, fgt_grad, fphi, fgphi, fnormal;  // fnormal: nf x 3
};

// A COUPLED bilinear term: test functions on one fem (rows), trial functions on another (columns), same mesh and
// quadrature -- the off-diagonal blocks of mixed formulations (rect.cu).
// extension matrix E of a reduced mesh_fem (nb_basic_dof x nb_dof), CSR, and its transpose (matrix.cu)
struct gfgpu_reduction {
  gfgpu_ctx *ctx;
  int64_t n_basic, n_dof, nnz;
  gf::DevBuf<int64_t> rp, trp;
  gf::DevBuf<int32_t> col, tcol;
  gf::DevBuf<double> val, tval;
};

struct gfgpu_rect {
  gfgpu_ctx *ctx;
  gfgpu_mesh *mesh;
  gfgpu_fem *fr, *fc;       // row / column fem
  gfgpu_tables *tr, *tc;    // their tables at the SAME quadrature points
  int family;
  double coef, alpha;
  int64_t nrows, ncols, ne;
  int sr, sc;               // local rows / columns of the element matrix: nd * qdim
  bool pat_valid = false;
  int64_t nnz = 0, generation = 0, nkept = 0;
  gf::DevBuf<double> stage;      // ne x sr x sc element matrices, column-major, thresholded
  gf::DevBuf<uint8_t> keep;      // ... and which entries the drop rule keeps
  gf::DevBuf<uint32_t> perm;     // kept contributions sorted by (column, row), ascending element inside an entry
  gf::DevBuf<uint32_t> seg;      // nnz + 1: entry -> range in perm
  gf::DevBuf<int64_t> jc, jct;   // CSC of the block (ncols + 1) and of its transpose (nrows + 1)
  gf::DevBuf<int32_t> ir, irt;
  gf::DevBuf<double> pr, prt;
  gf::DevBuf<uint32_t> tperm;    // entry k of the transpose = entry tperm[k] of the block
  // mesh region (gfgpu_rect_set_region): region-ordered copies of the connectivity and of the two dof tables, face of every item
  bool region = false, region_faces = false;
  gf::DevBuf<int32_t> r_conn, r_edr, r_edc;
  gf::DevBuf<int8_t> r_face;
};

struct gfgpu_term {
  gfgpu_ctx *ctx;
  gfgpu_mesh *mesh;
  gfgpu_fem *fem;
  gfgpu_tables *tab;
  int family, strategy, strategy_asked = 0;
  double par[GFGPU_MAX_PARAMS];
  double alpha;
  int64_t e0, e1;
  // mesh region (gfgpu_term_set_region): the term then runs over `n_items` items (convex, face) through region-ordered
  // copies of the connectivity and dof tables, so that every kernel keeps addressing rows [e0, e1)
  bool region = false, region_faces = false;
  int64_t n_items = 0;
  gf::DevBuf<int32_t> r_conn, r_edof;
  gf::DevBuf<int8_t> r_face;
  std::vector<int32_t> h_items_cv;
  // fem-data coefficients (gfgpu_term_set_fields): the leading `nfields` parameters are fields on `dfem`
  int nfields = 0;
  gfgpu_fem *dfem = nullptr;
  gf::DevBuf<double> dphi, dfphi, dvals[2];   // basis of the data fem at the volume / face points; nodal values
  gf::DevBuf<int32_t> r_dedof;                 // region-ordered copy of the data fem's dof rows
  bool r_dedof_valid = false;
  const int32_t *conn_p() const { return region ? r_conn.p : mesh->conn.p; }
  const int32_t *edof_p() const { return region ? r_edof.p : fem->edof.p; }
  int64_t nb_items() const { return region ? n_items : mesh->ne; }
  // structure
  gf::Structure st;
  bool st_valid = false;
  // pattern (value dependent)
  bool pat_valid = false;
  int64_t generation = 0;
  int64_t nnz = 0;
  gf::DevBuf<uint16_t> pmask;   // npairs
  gf::DevBuf<uint32_t> prel;    // Q x npairs: offset of the pair's first kept entry inside column QJ+beta
  gf::DevBuf<int64_t> jc;       // ndof+1
  gf::DevBuf<int64_t> ctot;     // ndof (scratch)
  gf::DevBuf<int32_t> ir;       // nnz
  gf::DevBuf<double> pr;        // nnz
  gf::DevBuf<double> R;         // ndof
  // staging
  gf::DevBuf<double> stage;     // ne_loc x s1 x s1
  gf::DevBuf<uint16_t> emask;   // ne_loc x nd x nd
  gf::DevBuf<double> rstage;    // ne_loc x s1
  // Q = 1 entry-wise gather (scatter.cu): stage index of every single-contribution entry, list of the other pairs
  gf::DevBuf<uint32_t> g1_src, g1_multi;
  int64_t g1_nmulti = 0, g1_generation = -1;
  // Q = 1 direct mode: slot of every local contribution; compact stage + offsets of the multi-contribution entries
  gf::DevBuf<uint32_t> dslot, moff, dmpos;
  int64_t d_nmulti = 0;
  gf::DevBuf<uint16_t> dkloc;   // local index of the entry that goes to dslot[.] (slots sorted ascending inside an element)
  gf::DevBuf<double> mstage;
  int64_t d_generation = -1;
  int direct_ok = 0;  // 0 unknown, 1 yes, -1 no
  gf::DevBuf<double> Ubuf;      // ndof (host path)
  // JIT family (jit.cu): the two forms as C expressions, the compiled kernel, the parameters on the device
  std::string jit_form1, jit_form2, jit_form0;  // jit_form0: the order-0 integrand (gfgpu_term_set_jit_potential), may be empty
  double *jit_epot = nullptr;                    // set while a potential is computed: per-element shares, ne doubles
  void *jit_kernel = nullptr;
  gf::DevBuf<double> jit_par;
  bool jit_value_dependent = true;
  gf::DevBuf<int32_t> flag;     // pattern-changed flag
  bool flag_pending = false;    // a value-dependent gather raised (or not) the flag; nobody has read it yet (api.cu term_settle)
  // per-phase events of the last assemble: [0,1] element kernel, [2,3] gather, [4,5] residual gather, [6,7] pattern
  // [2k, 2k+1], k = 0 element kernel, 1 gather, 2 residual gather, 3 pattern, 4 recompute kernel
  cudaEvent_t ev[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool ev_used[5] = {false, false, false, false, false};
  // multi-GPU halo: this rank owns the columns [own_lo, own_hi); columns below own_lo are ghosts
  // whose partial values are sent to their owners, sources are the ranks that send their parts of OUR columns
  struct HaloSource {
    int rank = 0;
    int64_t n = 0, nvals = 0;            // announced pairs; values per assembly
    int64_t r_lo = 0, r_hi = 0;          // residual range received from this source
    gf::DevBuf<int32_t> J, I;
    gf::DevBuf<uint16_t> mask;
    gf::DevBuf<uint32_t> soff;           // Q x n: offset of the pair's first kept entry of component b in the source's segment
    gf::DevBuf<int64_t> map;             // nvals: position in pr of every received value
    gf::DevBuf<double> recv, rrecv;      // tangent values, residual slice
  };
  bool halo = false;
  int64_t own_lo = 0, own_hi = 0;
  // what this rank sends at every exchange (gfgpu_term_halo_add_send): ghost columns [dof_lo, dof_hi) of rank `owner`,
  // residual slice [r_lo, dof_hi); pr_off / pr_cnt = that slice of pr under the current pattern
  struct HaloSend {
    int owner = 0;
    int64_t dof_lo = 0, dof_hi = 0, r_lo = 0, pr_off = 0, pr_cnt = 0;
  };
  std::vector<HaloSend> halo_sends;
  int64_t halo_sends_generation = -1;
  std::vector<std::unique_ptr<HaloSource>> halo_src;
  gf::DevBuf<int32_t> vJ, vI;            // all announced pairs, concatenated in source order
  gf::DevBuf<uint16_t> vmask;
  // strategy RECOMPUTE
  bool rc_ready = false;
  gf::DevBuf<double> rc_M;      // reference tensors per (j,i)
  gf::DevBuf<double> rc_eg;     // per-element geometry
  gf::DevBuf<double> rc_tgeo;   // per tile: the geometry rows of its distinct elements, in slot order (one bulk copy per tile)
  int rc_geo_rows = 0;
  gf::DevBuf<double> rc_L;      // low-rank factor of the reference Gram matrix (residual kernel)
  int rc_rank = 0;
  struct alignas(16) PairRec { uint32_t x, y, z, w; };
  gf::DevBuf<uint16_t> rc_dblob; // per long task (> 2 steps): descriptors of steps 0..9, [step][lane]
  gf::DevBuf<PairRec> rc_prec;   // per (task, lane): packed CSC offsets (relative to the tile base), keep mask, row dof
  gf::DevBuf<uint8_t> rc_hdr;    // TileHdr per tile (recompute_tiles.cu)
  gf::DevBuf<uint32_t> rc_els;   // per tile: sorted distinct local element ids (stride rc_cap_inc)
  int64_t rc_nt = 0, rc_ntask = 0;
  int rc_cap_inc = 0, rc_cap_len = 0, rc_cap_pairs = 0, rc_cap_slots = 0, rc_cap_tasks = 0, rc_cap_long = 0;
  int rc_res_block = 128;         // threads per block of k_affine_residual (64 when it has to fit next to the tile kernel)
  // column kernel (recompute_cols.cu): low-order scalar-coefficient forms, thread per column node
  bool rc_cols = false;
  int rc_cmaxq = 0;               // most pairs in one column
  gf::DevBuf<int64_t> rc_cgoff;   // per group of 32 column nodes: first ELL row
  gf::DevBuf<double> rc_cell;     // warp-transposed incidence records: geometry row, j and pair slots
  int64_t rc_cngroups = 0;
  // class-uniform tile kernel (recompute_uniform.cu): columns with translated copies run in lock step, one per lane
  bool rc_uni = false;
  gf::DevBuf<uint8_t> ru_tiles;    // UTile per tile (+ one sentinel)
  gf::DevBuf<uint32_t> ru_cta;     // first tile of every CTA
  gf::DevBuf<uint32_t> ru_prog;    // per (class, sub-range) programs
  gf::DevBuf<uint32_t> ru_ld;      // per chunk of 32 class members: CSC base and element strip positions of every lane
  gf::DevBuf<double> ru_eg;        // per-element geometry, component-major, strip order
  int64_t ru_nepad = 0, ru_ntiles = 0, ru_ntasks = 0;
  int ru_grid = 0, ru_nbuf = 0, ru_imgcap = 0, ru_kg = 3, ru_tw = 4;
};

namespace gf {

// ---- element kernels (elem_*.cu)
struct ElemArgs {
  const double *x, *y, *z;
  const int32_t *conn;
  const int32_t *edof;
  const double *U;
  const double *w, *gt_grad, *phi, *gphi;
  int nq, ng, qc;
  int64_t e0, e1;
  double par[GFGPU_MAX_PARAMS];
  double alpha;
  int family;
  // boundary faces: face of every item (nullptr = volume integration), tables at the face points, reference normals
  const int8_t *face;
  const double *fw, *fgt_grad, *fphi, *fgphi, *fnormal;  // fnormal: nf x 3
  int nqf;
  // fem-data coefficients: values at the Gauss point = sum_i vals[dedof[e][i] (+ component)] * dphi[q][i]
  int nfields, nd_d;
  const int32_t *dedof;
  const double *dphi, *dfphi, *dvals0, *dvals1;
  double *stage;
  uint16_t *emask;
  double *rstage;
  // direct mode (sum-factorised scalar kernels): every element entry goes straight to its CSC slot (slot < nnz32), to its
  // place in the compact stage of the multi-contribution entries (slot - nnz32), or nowhere (0xffffffff: not in the pattern)
  const uint32_t *slot = nullptr;
  const uint16_t *kloc = nullptr;
  double *pr = nullptr, *mstage = nullptr;
  uint32_t nnz32 = 0;
};
// returns false when the (dim, Q, nd, family, affine) combination has no instantiation
bool launch_elem_kernel(gfgpu_ctx *ctx, int dim, int Q, int nd, bool affine, const ElemArgs &a);

// sum-factorised kernel for the scalar Laplace form on Q3/Q4 hexahedra (sumfact.cu); false = not handled, use the generic one
bool launch_sumfact_kernel(gfgpu_ctx *ctx, const gfgpu_tables *tab, int dim, int Q, int nd, bool affine, const ElemArgs &a);
// 0: not handled by sumfact.cu, 1: scalar Laplace kernel (supports the direct mode), 2: hyperelastic kernel
int sumfact_kind(const gfgpu_tables *tab, int dim, int Q, int nd, bool affine, const ElemArgs &a);

// ---- first-touch dof numbering (dof_enum.cu); returns ndof
int64_t enumerate_dof(gfgpu_ctx *ctx, const int32_t *conn, int64_t ne, int ng, int N, bool qk, int k, int Q, int nd,
                      const int8_t *lat_host, int32_t *edof_dev);

// ---- scatter structure / pattern / gather (scatter.cu)
void build_structure(gfgpu_ctx *ctx, const int32_t *edof, int nd, int64_t e0, int64_t e1, int64_t ndof,
                     Structure &st, const int32_t *vJ = nullptr, const int32_t *vI = nullptr, int64_t nvirt = 0);
void halo_build_maps(gfgpu_term *t);
void halo_accumulate(gfgpu_term *t, bool do_t, bool do_r);
void build_pattern(gfgpu_term *t);
void gather_tangent(gfgpu_term *t, bool check);
// direct mode (Q = 1): slots of every local contribution, built once per pattern; false = not applicable (index range)
bool direct_prepare(gfgpu_term *t);
void direct_finish(gfgpu_term *t);  // ordered sums of the multi-contribution entries
void gather_residual(gfgpu_term *t);

// ---- strategy RECOMPUTE (recompute.cu)
bool recompute_supported(const gfgpu_term *t);
void recompute_prepare(gfgpu_term *t);
void recompute_assemble(gfgpu_term *t, const double *U, bool do_t, bool do_r);
// column kernel for low-order Laplace / mass forms (recompute_cols.cu)
bool recompute_cols_wanted(const gfgpu_term *t);
bool recompute_cols_prepare(gfgpu_term *t, const std::vector<uint32_t> &colstart, const std::vector<uint32_t> &rstart);
void recompute_cols_tangent(gfgpu_term *t, const double *U, bool with_r);
// class-uniform tile kernel (recompute_uniform.cu); prepare returns false when the term keeps the general tile kernel
bool uniform_prepare(gfgpu_term *t);
void launch_jit_kernel(gfgpu_term *t, const ElemArgs &a);  // jit.cu
void jit_release(gfgpu_term *t);
std::string jit_check_source(int N, int Q, const std::string &form1, const std::string &form2);
void term_settle_pending(gfgpu_term *t);  // api.cu: deferred pattern check of a value-dependent tangent
void uniform_tangent(gfgpu_term *t);

}  // namespace gf
