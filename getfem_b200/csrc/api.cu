// api.cu -- the extern "C" boundary declared in include/gfgpu.h.
#include <algorithm>
#include <climits>
#include <cstring>
#include <memory>

#include "common.cuh"

namespace gf {
std::atomic<int64_t> g_launches{0};
static thread_local std::string g_err;
void set_last_error(const std::string &s) { g_err = s; }
}  // namespace gf

#define GF_API_BEGIN try {
#define GF_API_END                      \
  return 0;                             \
  }                                     \
  catch (const std::exception &ex) {    \
    gf::g_err = ex.what();              \
    return 1;                           \
  }                                     \
  catch (...) {                         \
    gf::g_err = "unknown error";        \
    return 1;                           \
  }

using gf::DevBuf;

namespace gf {
// fp64 FMA throughput probe: 8 independent accumulator chains per thread, operands in registers
__global__ void __launch_bounds__(256) k_dfma_probe(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[0] = s;  // never true: keeps the chains alive
}
// fp64 tensor-core throughput probe: independent chains of mma.sync m8n8k4 f64 (DMMA), operands in registers
__global__ void __launch_bounds__(256) k_dmma_probe(double *out, int iters, double a, double b) {
  double c0[2] = {threadIdx.x * 1e-9, 1.0}, c1[2] = {2.0, 3.0}, c2[2] = {4.0, 5.0}, c3[2] = {6.0, 7.0};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
    }
  }
  const double s = ((c0[0] + c0[1]) + (c1[0] + c1[1])) + ((c2[0] + c2[1]) + (c3[0] + c3[1]));
  if (s == 123.456) out[0] = s;
}
}  // namespace gf

extern "C" {

const char *gfgpu_last_error(void) { return gf::g_err.c_str(); }
int gfgpu_version(void) { return 100; }
int64_t gfgpu_launch_count(void) { return gf::g_launches.load(); }

int gfgpu_ctx_create(int device, void *stream, gfgpu_ctx **out) {
  GF_API_BEGIN
  GF_REQUIRE(out, "null output");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  GF_REQUIRE(e == cudaSuccess && ndev > 0,
             std::string("no CUDA device: this library has no CPU fallback (") + cudaGetErrorString(e) + ")");
  GF_REQUIRE(device >= 0 && device < ndev, "bad device index");
  GF_CUDA(cudaSetDevice(device));
  std::unique_ptr<gfgpu_ctx> c(new gfgpu_ctx);
  c->device = device;
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    GF_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  GF_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
  GF_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
  GF_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  GF_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  *out = c.release();
  GF_API_END
}

int gfgpu_ctx_destroy(gfgpu_ctx *ctx) {
  GF_API_BEGIN
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->cub_tmp) cudaFree(ctx->cub_tmp);
  if (ctx->stream2) { cudaStreamSynchronize(ctx->stream2); cudaStreamDestroy(ctx->stream2); }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  GF_API_END
}

int gfgpu_ctx_synchronize(gfgpu_ctx *ctx) {
  GF_API_BEGIN
  GF_REQUIRE(ctx, "null context");
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  GF_API_END
}

int64_t gfgpu_ctx_bytes_in_use(gfgpu_ctx *ctx) { return ctx ? ctx->bytes : 0; }

int gfgpu_ctx_measure_fp64_peak(gfgpu_ctx *ctx, double *tflops) {
  GF_API_BEGIN
  GF_REQUIRE(ctx && tflops, "null argument");
  GF_CUDA(cudaSetDevice(ctx->device));
  DevBuf<double> out;
  out.alloc(ctx, 1);
  cudaEvent_t e0, e1;
  GF_CUDA(cudaEventCreate(&e0));
  GF_CUDA(cudaEventCreate(&e1));
  const int grid = ctx->sm_count * 8, iters = 4096;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {  // first pass = warm-up
    GF_CUDA(cudaEventRecord(e0, ctx->stream));
    gf::k_dfma_probe<<<grid, 256, 0, ctx->stream>>>(out.p, iters, 0.999999, 1e-9);
    GF_LAUNCH_CHECK();
    GF_CUDA(cudaEventRecord(e1, ctx->stream));
    GF_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    GF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 2.0 * 64.0 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
    if (rep > 0) best = std::max(best, tf);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  GF_API_END
}

int gfgpu_ctx_measure_dmma_peak(gfgpu_ctx *ctx, double *tflops) {
  GF_API_BEGIN
  GF_REQUIRE(ctx && tflops, "null argument");
  GF_CUDA(cudaSetDevice(ctx->device));
  DevBuf<double> out;
  out.alloc(ctx, 1);
  cudaEvent_t e0, e1;
  GF_CUDA(cudaEventCreate(&e0));
  GF_CUDA(cudaEventCreate(&e1));
  const int grid = ctx->sm_count * 8, iters = 2048;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    GF_CUDA(cudaEventRecord(e0, ctx->stream));
    gf::k_dmma_probe<<<grid, 256, 0, ctx->stream>>>(out.p, iters, 0.999999, 1e-9);
    GF_LAUNCH_CHECK();
    GF_CUDA(cudaEventRecord(e1, ctx->stream));
    GF_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    GF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    // one m8n8k4 = 8*8*4 FMA per warp; 32 per iteration and warp; 8 warps per CTA
    const double tf = 2.0 * 256.0 * 32.0 * iters * 8.0 * grid / (ms * 1e-3) / 1e12;
    if (rep > 0) best = std::max(best, tf);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  GF_API_END
}

int gfgpu_mesh_create(gfgpu_ctx *ctx, int dim, int64_t npts, const double *pts, int64_t ne, int ng,
                      const int32_t *conn, int gt_kind, gfgpu_mesh **out) {
  GF_API_BEGIN
  GF_REQUIRE(ctx && out, "null argument");
  GF_REQUIRE(dim == 2 || dim == 3, "mesh dimension must be 2 or 3");
  GF_REQUIRE(npts >= 0 && ne >= 0 && ng >= dim + 1, "bad mesh sizes");
  GF_REQUIRE(gt_kind == GFGPU_GT_PK || gt_kind == GFGPU_GT_QK, "unknown geometric transformation kind");
  GF_REQUIRE((pts || !npts) && (conn || !ne), "null mesh arrays");
  GF_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<gfgpu_mesh> m(new gfgpu_mesh);
  m->ctx = ctx; m->dim = dim; m->ng = ng; m->gt_kind = gt_kind; m->npts = npts; m->ne = ne;
  for (int64_t k = 0; k < ne * ng; ++k)
    GF_REQUIRE(conn[k] >= 0 && conn[k] < npts, "connectivity refers to a point outside [0, npts)");
  // AoS -> SoA on the host, then one upload
  std::vector<double> soa((size_t)3 * npts, 0.0);
  for (int64_t p = 0; p < npts; ++p)
    for (int d = 0; d < dim; ++d) soa[(size_t)d * npts + p] = pts[p * dim + d];
  m->xyz.alloc(ctx, soa.size());
  m->xyz.upload(soa.data());
  m->conn.alloc(ctx, (size_t)ne * ng);
  m->conn.upload(conn);
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = m.release();
  GF_API_END
}

int gfgpu_mesh_destroy(gfgpu_mesh *m) {
  GF_API_BEGIN
  delete m;
  GF_API_END
}

int gfgpu_fem_create(gfgpu_ctx *ctx, gfgpu_mesh *mesh, int fem_kind, int degree, int qdim, int nd,
                     const int64_t *elem_dof, int64_t ndof, gfgpu_fem **out) {
  GF_API_BEGIN
  GF_REQUIRE(ctx && mesh && out, "null argument");
  GF_REQUIRE(qdim >= 1 && qdim <= 3, "qdim must be 1, 2 or 3");
  GF_REQUIRE(nd >= 1, "bad number of local dofs");
  GF_REQUIRE(fem_kind == GFGPU_FEM_PK || fem_kind == GFGPU_FEM_QK, "unknown fem kind");
  GF_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<gfgpu_fem> f(new gfgpu_fem);
  f->ctx = ctx; f->mesh = mesh; f->fem_kind = fem_kind; f->degree = degree; f->qdim = qdim; f->nd = nd;
  if (elem_dof) {
    GF_REQUIRE(ndof > 0 && ndof < (int64_t(1) << 31) - 4, "ndof out of range");
    f->ndof = ndof;
    std::vector<int32_t> ed((size_t)mesh->ne * nd);
    for (size_t k = 0; k < ed.size(); ++k) {
      GF_REQUIRE(elem_dof[k] >= 0 && elem_dof[k] + qdim <= ndof, "element dof outside [0, ndof)");
      ed[k] = (int32_t)elem_dof[k];
    }
    f->edof.alloc(ctx, ed.size());
    f->edof.upload(ed.data());
    GF_CUDA(cudaStreamSynchronize(ctx->stream));
  } else {
    // local lattice of the classical Lagrange element (getfem_fem.cc:760-763, 816-825)
    const int N = mesh->dim;
    GF_REQUIRE(degree >= 1 && degree <= 8, "fem degree must be in 1..8 for device enumeration");
    const bool qk = fem_kind == GFGPU_FEM_QK;
    GF_REQUIRE(qk == (mesh->gt_kind == GFGPU_GT_QK), "fem kind does not match the geometric transformation");
    GF_REQUIRE(mesh->ng == (qk ? (1 << N) : N + 1), "device enumeration needs a degree-1 geometric transformation");
    std::vector<int8_t> lat;
    int cnt = 0;
    if (qk) {
      const int n1 = degree + 1;
      int tot = 1;
      for (int d = 0; d < N; ++d) tot *= n1;
      for (int i = 0; i < tot; ++i) {
        int r = i;
        int8_t l[4] = {0, 0, 0, 0};
        for (int d = 0; d < N; ++d) { l[d] = (int8_t)(r % n1); r /= n1; }
        lat.insert(lat.end(), l, l + 4);
        ++cnt;
      }
    } else {
      for (int c = 0; c <= (N == 3 ? degree : 0); ++c)
        for (int b = 0; b <= degree; ++b)
          for (int a = 0; a <= degree; ++a)
            if (a + b + c <= degree) {
              int8_t l[4] = {(int8_t)(degree - a - b - c), (int8_t)a, (int8_t)b, (int8_t)c};
              lat.insert(lat.end(), l, l + 4);
              ++cnt;
            }
    }
    GF_REQUIRE(cnt == nd, "nd does not match the classical Lagrange element of this kind/degree");
    f->edof.alloc(ctx, (size_t)mesh->ne * nd);
    f->ndof = gf::enumerate_dof(ctx, mesh->conn.p, mesh->ne, mesh->ng, N, qk, degree, qdim, nd, lat.data(), f->edof.p);
    GF_REQUIRE(f->ndof < (int64_t(1) << 31) - 4, "ndof out of range");
  }
  *out = f.release();
  GF_API_END
}

int64_t gfgpu_fem_nb_dof(gfgpu_fem *f) { return f ? f->ndof : -1; }

int gfgpu_fem_get_elem_dof(gfgpu_fem *f, int64_t *out) {
  GF_API_BEGIN
  GF_REQUIRE(f && out, "null argument");
  GF_CUDA(cudaSetDevice(f->ctx->device));
  std::vector<int32_t> ed(f->edof.n);
  f->edof.download(ed.data());
  GF_CUDA(cudaStreamSynchronize(f->ctx->stream));
  for (size_t k = 0; k < ed.size(); ++k) out[k] = ed[k];
  GF_API_END
}

int gfgpu_fem_destroy(gfgpu_fem *f) {
  GF_API_BEGIN
  delete f;
  GF_API_END
}

int gfgpu_tables_create(gfgpu_ctx *ctx, int dim, int nq, int ng, int nd, const double *w, const double *gt_grad,
                        const double *phi, const double *gphi, gfgpu_tables **out) {
  GF_API_BEGIN
  GF_REQUIRE(ctx && out && w && gt_grad && phi && gphi, "null argument");
  GF_REQUIRE(nq >= 1 && ng >= 1 && nd >= 1, "bad table sizes");
  GF_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<gfgpu_tables> t(new gfgpu_tables);
  t->ctx = ctx; t->dim = dim; t->nq = nq; t->ng = ng; t->nd = nd;
  t->h_w.assign(w, w + nq);
  t->h_gt_grad.assign(gt_grad, gt_grad + (size_t)nq * ng * dim);
  t->h_phi.assign(phi, phi + (size_t)nq * nd);
  t->h_gphi.assign(gphi, gphi + (size_t)nq * nd * dim);
  t->w.alloc(ctx, t->h_w.size()); t->w.upload(t->h_w.data());
  t->gt_grad.alloc(ctx, t->h_gt_grad.size()); t->gt_grad.upload(t->h_gt_grad.data());
  t->phi.alloc(ctx, t->h_phi.size()); t->phi.upload(t->h_phi.data());
  t->gphi.alloc(ctx, t->h_gphi.size()); t->gphi.upload(t->h_gphi.data());
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = t.release();
  GF_API_END
}

int gfgpu_tables_set_gt_values(gfgpu_tables *t, const double *gt_val, const double *face_gt_val) {
  GF_API_BEGIN
  GF_REQUIRE(t && gt_val, "null argument");
  GF_CUDA(cudaSetDevice(t->ctx->device));
  t->gt_val.alloc(t->ctx, (size_t)t->nq * t->ng);
  t->gt_val.upload(gt_val);
  if (face_gt_val) {
    GF_REQUIRE(t->nf > 0, "face values need the face tables first (gfgpu_tables_set_faces)");
    t->fgt_val.alloc(t->ctx, (size_t)t->nf * t->nqf * t->ng);
    t->fgt_val.upload(face_gt_val);
  }
  GF_CUDA(cudaStreamSynchronize(t->ctx->stream));
  GF_API_END
}

int gfgpu_tables_set_faces(gfgpu_tables *t, int nf, int nqf, const double *normals, const double *w,
                           const double *gt_grad, const double *phi, const double *gphi) {
  GF_API_BEGIN
  GF_REQUIRE(t && normals && w && gt_grad && phi && gphi, "null argument");
  GF_REQUIRE(nf >= 1 && nf <= GFGPU_MAX_FACES && nqf >= 1, "bad face table sizes");
  gfgpu_ctx *ctx = t->ctx;
  GF_CUDA(cudaSetDevice(ctx->device));
  const size_t np = (size_t)nf * nqf;
  t->nf = nf; t->nqf = nqf;
  std::vector<double> hn((size_t)nf * 3, 0.0);
  for (int f = 0; f < nf; ++f)
    for (int d = 0; d < t->dim; ++d) hn[f * 3 + d] = normals[f * t->dim + d];
  t->fnormal.alloc(ctx, hn.size()); t->fnormal.upload(hn.data());
  t->fw.alloc(ctx, np); t->fw.upload(w);
  t->fgt_grad.alloc(ctx, np * t->ng * t->dim); t->fgt_grad.upload(gt_grad);
  t->fphi.alloc(ctx, np * t->nd); t->fphi.upload(phi);
  t->fgphi.alloc(ctx, np * t->nd * t->dim); t->fgphi.upload(gphi);
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  GF_API_END
}

int gfgpu_tables_destroy(gfgpu_tables *t) {
  GF_API_BEGIN
  delete t;
  GF_API_END
}

int gfgpu_term_create(gfgpu_ctx *ctx, gfgpu_mesh *mesh, gfgpu_fem *fem, gfgpu_tables *tab, int family,
                      const double *params, int nparams, double alpha, int strategy, gfgpu_term **out) {
  GF_API_BEGIN
  GF_REQUIRE(ctx && mesh && fem && tab && out, "null argument");
  GF_REQUIRE(fem->mesh == mesh, "the fem was built on another mesh");
  GF_REQUIRE(tab->dim == mesh->dim && tab->ng == mesh->ng && tab->nd == fem->nd, "tables do not match mesh/fem");
  GF_REQUIRE(family >= GFGPU_LAPLACE && family <= GFGPU_BLATZ_KO, "unknown expression family");
  const int need = family == GFGPU_SOURCE ? fem->qdim
                   : family == GFGPU_NORMAL_SOURCE ? fem->qdim * mesh->dim
                   : (family == GFGPU_LAPLACE || family == GFGPU_MASS) ? 1
                   : (family == GFGPU_MOONEY_RIVLIN || family == GFGPU_CIARLET_GEYMONAT) ? 3
                   : family == GFGPU_BLATZ_KO ? 5 : 2;
  GF_REQUIRE(params && nparams >= need, "missing parameters for this family");
  if (family == GFGPU_ELASTICITY) GF_REQUIRE(fem->qdim == mesh->dim, "elasticity needs qdim == mesh dimension");
  if (family == GFGPU_SVK || family == GFGPU_NEOHOOKEAN_CIARLET || family == GFGPU_NEOHOOKEAN_BONET || family >= GFGPU_MOONEY_RIVLIN)
    GF_REQUIRE(fem->qdim == 3 && mesh->dim == 3, "finite-strain families need a 3D vector field");
  GF_REQUIRE(strategy >= GFGPU_STRATEGY_AUTO && strategy <= GFGPU_STRATEGY_RECOMPUTE, "unknown strategy");
  std::unique_ptr<gfgpu_term> t(new gfgpu_term);
  t->ctx = ctx; t->mesh = mesh; t->fem = fem; t->tab = tab; t->family = family;
  const bool rc_ok = gf::recompute_supported(t.get());
  GF_REQUIRE(strategy != GFGPU_STRATEGY_RECOMPUTE || rc_ok,
             "strategy RECOMPUTE needs an affine (simplex) mesh and a Laplace / elasticity / mass term");
  t->strategy = strategy == GFGPU_STRATEGY_AUTO ? (rc_ok ? GFGPU_STRATEGY_RECOMPUTE : GFGPU_STRATEGY_STAGED) : strategy;
  t->strategy_asked = strategy;
  for (int k = 0; k < GFGPU_MAX_PARAMS; ++k) t->par[k] = k < nparams ? params[k] : 0.0;
  t->alpha = alpha;
  t->e0 = 0; t->e1 = mesh->ne;
  GF_CUDA(cudaSetDevice(ctx->device));
  t->flag.alloc(ctx, 1);
  t->flag.zero();
  for (int k = 0; k < 10; ++k) GF_CUDA(cudaEventCreate(&t->ev[k]));
  *out = t.release();
  GF_API_END
}

int gfgpu_term_create_jit(gfgpu_ctx *ctx, gfgpu_mesh *mesh, gfgpu_fem *fem, gfgpu_tables *tab, const char *form1, const char *form2,
                          const double *params, int nparams, double alpha, int value_dependent, gfgpu_term **out) {
  GF_API_BEGIN
  GF_REQUIRE(ctx && mesh && fem && tab && form1 && form2 && out, "null argument");
  GF_REQUIRE(fem->mesh == mesh, "the fem was built on another mesh");
  GF_REQUIRE(tab->dim == mesh->dim && tab->ng == mesh->ng && tab->nd == fem->nd, "tables do not match mesh/fem");
  GF_REQUIRE(mesh->dim == 2 || mesh->dim == 3, "JIT terms: 2D and 3D meshes");
  GF_REQUIRE(fem->qdim == 1 || fem->qdim == mesh->dim, "JIT terms: scalar variables, or vector variables of the mesh dimension");
  GF_REQUIRE(nparams >= 0 && nparams <= GFGPU_MAX_PARAMS && (nparams == 0 || params), "bad parameters");
  std::unique_ptr<gfgpu_term> t(new gfgpu_term);
  t->ctx = ctx; t->mesh = mesh; t->fem = fem; t->tab = tab; t->family = GFGPU_JIT;
  t->strategy = GFGPU_STRATEGY_STAGED;
  t->strategy_asked = GFGPU_STRATEGY_STAGED;
  for (int k = 0; k < GFGPU_MAX_PARAMS; ++k) t->par[k] = k < nparams ? params[k] : 0.0;
  t->alpha = alpha;
  t->e0 = 0; t->e1 = mesh->ne;
  t->jit_form1 = form1; t->jit_form2 = form2;
  t->jit_value_dependent = value_dependent != 0;
  GF_CUDA(cudaSetDevice(ctx->device));
  t->flag.alloc(ctx, 1);
  t->flag.zero();
  for (int k = 0; k < 10; ++k) GF_CUDA(cudaEventCreate(&t->ev[k]));
  *out = t.release();
  GF_API_END
}

int gfgpu_term_set_params(gfgpu_term *t, const double *params, int nparams) {
  GF_API_BEGIN
  GF_REQUIRE(t && nparams >= 0 && nparams <= GFGPU_MAX_PARAMS && (nparams == 0 || params), "bad parameters");
  GF_REQUIRE(t->family == GFGPU_JIT, "parameters can be replaced on JIT terms (the closed-form families fold them into their plans)");
  for (int k = 0; k < nparams; ++k) t->par[k] = params[k];
  GF_API_END
}

int gfgpu_term_set_jit_potential(gfgpu_term *t, const char *form0) {
  GF_API_BEGIN
  GF_REQUIRE(t && form0, "null argument");
  GF_REQUIRE(t->family == GFGPU_JIT, "only JIT terms take an order-0 form");
  if (t->jit_form0 != form0) {
    GF_CUDA(cudaSetDevice(t->ctx->device));
    GF_CUDA(cudaStreamSynchronize(t->ctx->stream));
    gf::jit_release(t);  // compiled again, with the potential, at the next launch
    t->jit_form0 = form0;
  }
  GF_API_END
}

int gfgpu_jit_check(int dim, int qdim, const char *form1, const char *form2) {
  GF_API_BEGIN
  GF_REQUIRE(form1 && form2 && (dim == 2 || dim == 3) && (qdim == 1 || qdim == dim), "bad argument");
  const std::string log = gf::jit_check_source(dim, qdim, form1, form2);
  GF_REQUIRE(log.empty(), "the integrand does not compile (NVRTC):\n" + log);
  GF_API_END
}

int gfgpu_term_destroy(gfgpu_term *t) {
  GF_API_BEGIN
  if (t) {
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    if (t->jit_kernel) gf::jit_release(t);
    for (int k = 0; k < 10; ++k)
      if (t->ev[k]) cudaEventDestroy(t->ev[k]);
  }
  delete t;
  GF_API_END
}

static void term_forget_symbolic(gfgpu_term *t) {
  t->st_valid = false;
  t->pat_valid = false;
  t->rc_ready = false;
  t->halo = false;
  t->halo_src.clear();
  t->vJ.release(); t->vI.release(); t->vmask.release();
}

int gfgpu_term_set_fields(gfgpu_term *t, int nfields, gfgpu_fem *dfem, const double *phi, const double *phi_faces,
                          const double *vals0, const double *vals1) {
  GF_API_BEGIN
  GF_REQUIRE(t, "null term");
  gfgpu_ctx *ctx = t->ctx;
  GF_CUDA(cudaSetDevice(ctx->device));
  term_forget_symbolic(t);
  t->stage.release(); t->emask.release(); t->rstage.release();
  t->r_dedof_valid = false;
  if (nfields == 0) {
    t->nfields = 0; t->dfem = nullptr;
    t->dphi.release(); t->dfphi.release(); t->dvals[0].release(); t->dvals[1].release(); t->r_dedof.release();
    if (!t->region_faces)
      t->strategy = t->strategy_asked == GFGPU_STRATEGY_AUTO
                        ? (gf::recompute_supported(t) ? GFGPU_STRATEGY_RECOMPUTE : GFGPU_STRATEGY_STAGED)
                        : t->strategy_asked;
    return 0;
  }
  GF_REQUIRE(dfem && phi && vals0, "null argument");
  GF_REQUIRE(dfem->mesh == t->mesh, "the data fem was built on another mesh");
  const int fam = t->family;
  const int maxf = (fam == GFGPU_LAPLACE || fam == GFGPU_MASS || fam == GFGPU_SOURCE) ? 1
                   : (fam == GFGPU_ELASTICITY || fam == GFGPU_JIT) ? 2 : 0;  // JIT terms: fld[0], fld[1] in the integrand
  GF_REQUIRE(maxf > 0, "fem-data coefficients are handled for the Laplace, mass, elasticity, source and JIT families");
  GF_REQUIRE(nfields >= 1 && nfields <= maxf, "wrong number of coefficient fields for this family");
  GF_REQUIRE(nfields < 2 || vals1, "null argument");
  if (fam == GFGPU_JIT && dfem->qdim != 1)  // JIT terms: scalar fields fld[k], or ONE vector field vfld of the mesh dimension
    GF_REQUIRE(dfem->qdim == t->mesh->dim && nfields == 1, "JIT terms: 1-2 scalar fields, or one vector field of the mesh dimension");
  else
    GF_REQUIRE(dfem->qdim == (fam == GFGPU_SOURCE ? t->fem->qdim : 1),
               "the data fem must be scalar (the source term's: qdim of the variable)");
  GF_REQUIRE(t->strategy_asked != GFGPU_STRATEGY_RECOMPUTE, "fem-data coefficients use strategy STAGED");
  const int ndd = dfem->nd, nq = t->tab->nq;
  t->dphi.alloc(ctx, (size_t)nq * ndd); t->dphi.upload(phi);
  if (phi_faces) {
    GF_REQUIRE(t->tab->nf > 0, "face tables of the data fem need gfgpu_tables_set_faces first");
    t->dfphi.alloc(ctx, (size_t)t->tab->nf * t->tab->nqf * ndd); t->dfphi.upload(phi_faces);
  } else t->dfphi.release();
  t->dvals[0].alloc(ctx, dfem->ndof); t->dvals[0].upload(vals0);
  if (nfields > 1) { t->dvals[1].alloc(ctx, dfem->ndof); t->dvals[1].upload(vals1); } else t->dvals[1].release();
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  t->nfields = nfields; t->dfem = dfem;
  t->strategy = GFGPU_STRATEGY_STAGED;
  GF_API_END
}

int gfgpu_term_update_field(gfgpu_term *t, int k, const double *vals) {
  GF_API_BEGIN
  GF_REQUIRE(t && vals && k >= 0 && k < t->nfields, "no such coefficient field");
  GF_CUDA(cudaSetDevice(t->ctx->device));
  t->dvals[k].upload(vals);
  GF_CUDA(cudaStreamSynchronize(t->ctx->stream));
  t->pat_valid = false;  // the drop rule looks at the values
  GF_API_END
}

int gfgpu_term_set_element_range(gfgpu_term *t, int64_t e0, int64_t e1) {
  GF_API_BEGIN
  GF_REQUIRE(t, "null term");
  GF_REQUIRE(0 <= e0 && e0 <= e1 && e1 <= t->nb_items(), "bad element range");
  if (e0 != t->e0 || e1 != t->e1) {
    t->e0 = e0; t->e1 = e1;
    term_forget_symbolic(t);
  }
  GF_API_END
}

int gfgpu_term_set_region(gfgpu_term *t, int64_t n_items, const int32_t *cv, const int32_t *face) {
  GF_API_BEGIN
  GF_REQUIRE(t, "null term");
  GF_REQUIRE(n_items >= 0 && (cv || n_items == 0), "bad region");
  gfgpu_ctx *ctx = t->ctx;
  GF_CUDA(cudaSetDevice(ctx->device));
  term_forget_symbolic(t);
  t->stage.release(); t->emask.release(); t->rstage.release();
  if (!cv) {  // back to all convexes
    t->region = t->region_faces = false;
    t->n_items = 0;
    t->h_items_cv.clear();
    t->r_dedof_valid = false;
    t->r_conn.release(); t->r_edof.release(); t->r_face.release();
    if (!t->nfields)
      t->strategy = t->strategy_asked == GFGPU_STRATEGY_AUTO
                        ? (gf::recompute_supported(t) ? GFGPU_STRATEGY_RECOMPUTE : GFGPU_STRATEGY_STAGED)
                        : t->strategy_asked;
    t->e0 = 0; t->e1 = t->mesh->ne;
    return 0;
  }
  const int ng = t->mesh->ng, nd = t->fem->nd;
  int nfaces = 0;
  for (int64_t k = 0; k < n_items; ++k) {
    GF_REQUIRE(cv[k] >= 0 && cv[k] < t->mesh->ne, "region refers to a convex outside the mesh");
    GF_REQUIRE(k == 0 || cv[k] > cv[k - 1] || (cv[k] == cv[k - 1] && face && face[k] > face[k - 1]),
               "region items must come in mr_visitor order (ascending convex, then face)");
    if (face && face[k] >= 0) ++nfaces;
  }
  GF_REQUIRE(nfaces == 0 || nfaces == n_items, "a region must hold either convexes or faces, not both");
  if (nfaces) {
    GF_REQUIRE(t->tab->nf > 0, "a region of faces needs gfgpu_tables_set_faces");
    GF_REQUIRE(t->strategy_asked != GFGPU_STRATEGY_RECOMPUTE, "regions of faces use strategy STAGED");
    for (int64_t k = 0; k < n_items; ++k) GF_REQUIRE(face[k] < t->tab->nf, "face number outside the reference element");
  } else {
    GF_REQUIRE(t->family != GFGPU_NORMAL_SOURCE, "the normal source term needs a region of faces");
  }
  // region-ordered copies of the connectivity and dof rows (built on the host: a region is set once)
  std::vector<int32_t> hc((size_t)t->mesh->ne * ng), hd((size_t)t->mesh->ne * nd);
  t->mesh->conn.download(hc.data());
  t->fem->edof.download(hd.data());
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<int32_t> rc((size_t)n_items * ng), rd((size_t)n_items * nd);
  std::vector<int8_t> rf((size_t)n_items);
  for (int64_t k = 0; k < n_items; ++k) {
    std::copy(hc.begin() + (size_t)cv[k] * ng, hc.begin() + (size_t)(cv[k] + 1) * ng, rc.begin() + (size_t)k * ng);
    std::copy(hd.begin() + (size_t)cv[k] * nd, hd.begin() + (size_t)(cv[k] + 1) * nd, rd.begin() + (size_t)k * nd);
    rf[k] = (int8_t)(nfaces ? face[k] : -1);
  }
  t->r_conn.alloc(ctx, rc.size()); t->r_conn.upload(rc.data());
  t->r_edof.alloc(ctx, rd.size()); t->r_edof.upload(rd.data());
  if (nfaces) { t->r_face.alloc(ctx, rf.size()); t->r_face.upload(rf.data()); } else t->r_face.release();
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  t->region = true;
  t->region_faces = nfaces > 0;
  t->n_items = n_items;
  t->h_items_cv.assign(cv, cv + n_items);
  t->r_dedof_valid = false;
  if (nfaces) t->strategy = GFGPU_STRATEGY_STAGED;
  t->e0 = 0; t->e1 = n_items;
  GF_API_END
}

// The pattern check of a value-dependent tangent, deferred: if the last gather found entries outside the pattern, the
// pattern is rebuilt from the masks of that same pass and the (still staged) element matrices are gathered again.
static void term_settle(gfgpu_term *t) {
  if (!t || !t->flag_pending) return;
  gfgpu_ctx *ctx = t->ctx;
  GF_CUDA(cudaSetDevice(ctx->device));
  t->flag_pending = false;
  int32_t changed = 0;
  t->flag.download(&changed);
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  if (changed) {
    GF_REQUIRE(!t->halo, "the pattern moved under a multi-GPU halo: announce the pairs again (halo_begin .. halo_commit)");
    gf::build_pattern(t);
    gf::gather_tangent(t, false);
  }
}

static void term_assemble(gfgpu_term *t, const double *U_dev, int order_mask) {
  gfgpu_ctx *ctx = t->ctx;
  GF_CUDA(cudaSetDevice(ctx->device));
  term_settle(t);  // before the stage of the previous pass is overwritten
  GF_REQUIRE(order_mask & (GFGPU_RESIDUAL | GFGPU_TANGENT), "order_mask selects nothing");
  bool do_t = order_mask & GFGPU_TANGENT;
  const bool do_r = order_mask & GFGPU_RESIDUAL;
  const int nd = t->fem->nd, Q = t->fem->qdim, s1 = nd * Q;
  const int64_t ne = t->e1 - t->e0;
  const bool order1_only = t->family == GFGPU_SOURCE || t->family == GFGPU_NORMAL_SOURCE;
  GF_REQUIRE(t->family != GFGPU_NORMAL_SOURCE || t->region_faces, "the normal source term needs a region of faces");
  if (order1_only && do_t) {
    // an order-1 term has no order-2 tree (workspace.cc:545-600): the tangent is structurally empty
    if (t->jc.n != (size_t)t->fem->ndof + 1) t->jc.alloc(ctx, t->fem->ndof + 1);
    t->jc.zero();
    t->nnz = 0;
    t->pat_valid = true;
    do_t = false;
    if (!do_r) return;
  }
  if (!t->st_valid) {
    gf::build_structure(ctx, t->edof_p(), nd, t->e0, t->e1, t->fem->ndof, t->st, t->vJ.p, t->vI.p, (int64_t)t->vJ.n);
    t->st_valid = true;
    t->pat_valid = false;
  }
  if (order1_only && t->jc.n) t->pat_valid = true;  // the (empty) pattern of an order-1 term never moves
  const bool recompute = t->strategy == GFGPU_STRATEGY_RECOMPUTE;
  // what the generic element kernel has to produce in this call.  RECOMPUTE needs it only once, for
  // the keep masks of the pattern; its residual is K^T U inside the per-nonzero kernel.
  // linear families: the keep masks do not depend on U, so a valid pattern needs no masks (and none are written)
  const bool value_dependent = t->family == GFGPU_JIT ? t->jit_value_dependent
                               : (t->family == GFGPU_SVK || t->family == GFGPU_NEOHOOKEAN_CIARLET ||
                                  t->family == GFGPU_NEOHOOKEAN_BONET || t->family >= GFGPU_MOONEY_RIVLIN);
  const bool need_masks = recompute ? !t->pat_valid : (do_t && (!t->pat_valid || value_dependent));
  // direct mode: a scalar sum-factorised kernel under a fixed pattern writes its entries straight to their CSC slots
  // (scatter.cu direct_prepare): no element matrix in HBM, no gather -- only the ordered sums of the shared entries
  bool direct = false;
  if (!recompute && do_t && Q == 1 && t->pat_valid && !value_dependent && !t->region_faces && !t->nfields && ne > 0 &&
      !getenv("GFGPU_NO_DIRECT")) {
    gf::ElemArgs probe;
    probe.family = t->family; probe.ng = t->mesh->ng; probe.nq = t->tab->nq;
    if (gf::sumfact_kind(t->tab, t->mesh->dim, Q, nd, t->mesh->gt_kind == GFGPU_GT_PK, probe) == 1) direct = gf::direct_prepare(t);
  }
  const bool need_stage = do_t && !recompute && !direct;
  const bool need_rstage = do_r && !recompute;
  if (need_stage && t->stage.n != (size_t)ne * s1 * s1) t->stage.alloc(ctx, (size_t)ne * s1 * s1);
  if (need_masks && t->emask.n != (size_t)ne * nd * nd) t->emask.alloc(ctx, (size_t)ne * nd * nd);
  if (need_rstage && t->rstage.n != (size_t)ne * s1) t->rstage.alloc(ctx, (size_t)ne * s1);
  if (do_r && t->R.n != (size_t)t->fem->ndof) t->R.alloc(ctx, t->fem->ndof);
  gf::ElemArgs a;
  const int64_t np = t->mesh->npts;
  a.x = t->mesh->xyz.p; a.y = a.x + np; a.z = a.y + np;
  a.conn = t->conn_p();
  a.edof = t->edof_p();
  a.U = U_dev;
  a.w = t->tab->w.p; a.gt_grad = t->tab->gt_grad.p; a.phi = t->tab->phi.p; a.gphi = t->tab->gphi.p;
  a.nq = t->tab->nq; a.ng = t->mesh->ng; a.qc = 0;
  a.e0 = t->e0; a.e1 = t->e1;
  for (int k = 0; k < GFGPU_MAX_PARAMS; ++k) a.par[k] = t->par[k];
  a.alpha = t->alpha;
  a.family = t->family;
  a.stage = need_stage ? t->stage.p : nullptr;
  a.emask = need_masks ? t->emask.p : nullptr;
  a.rstage = need_rstage ? t->rstage.p : nullptr;
  if (direct) { a.slot = t->dslot.p; a.kloc = t->dkloc.p; a.pr = t->pr.p; a.mstage = t->mstage.p; a.nnz32 = (uint32_t)t->nnz; }
  a.face = t->region_faces ? t->r_face.p : nullptr;
  a.fw = t->tab->fw.p; a.fgt_grad = t->tab->fgt_grad.p; a.fphi = t->tab->fphi.p; a.fgphi = t->tab->fgphi.p;
  a.fnormal = t->tab->fnormal.p;
  a.nqf = t->tab->nqf;
  a.nfields = t->nfields; a.nd_d = 0;
  a.dedof = nullptr; a.dphi = a.dfphi = a.dvals0 = a.dvals1 = nullptr;
  if (t->nfields) {
    GF_REQUIRE(!t->region_faces || t->dfphi.n, "a region of faces needs the data fem's basis at the face points");
    if (t->region && !t->r_dedof_valid) {  // region-ordered copy of the data fem's dof rows
      const int ndd = t->dfem->nd;
      std::vector<int32_t> hd((size_t)t->mesh->ne * ndd), rd((size_t)t->n_items * ndd);
      t->dfem->edof.download(hd.data());
      GF_CUDA(cudaStreamSynchronize(ctx->stream));
      for (int64_t k = 0; k < t->n_items; ++k)
        std::copy(hd.begin() + (size_t)t->h_items_cv[k] * ndd, hd.begin() + (size_t)(t->h_items_cv[k] + 1) * ndd,
                  rd.begin() + (size_t)k * ndd);
      t->r_dedof.alloc(ctx, rd.size());
      t->r_dedof.upload(rd.data());
      GF_CUDA(cudaStreamSynchronize(ctx->stream));
      t->r_dedof_valid = true;
    }
    a.nd_d = t->dfem->nd;
    a.dedof = t->region ? t->r_dedof.p : t->dfem->edof.p;
    a.dphi = t->dphi.p; a.dfphi = t->dfphi.p; a.dvals0 = t->dvals[0].p; a.dvals1 = t->dvals[1].p;
  }
  for (int k = 0; k < 5; ++k) t->ev_used[k] = false;
  auto tic = [&](int k) { GF_CUDA(cudaEventRecord(t->ev[2 * k], ctx->stream)); };
  auto toc = [&](int k) { GF_CUDA(cudaEventRecord(t->ev[2 * k + 1], ctx->stream)); t->ev_used[k] = true; };
  if (ne > 0 && (need_stage || need_masks || need_rstage || direct)) {
    tic(0);
    const bool affine = t->mesh->gt_kind == GFGPU_GT_PK;
    if (t->family == GFGPU_JIT) {
      gf::launch_jit_kernel(t, a);
    }
    bool ok = t->family == GFGPU_JIT ||
              (!t->region_faces && !t->nfields && gf::launch_sumfact_kernel(ctx, t->tab, t->mesh->dim, Q, nd, affine, a)) ||
              (!direct && gf::launch_elem_kernel(ctx, t->mesh->dim, Q, nd, affine, a));  // only sumfact.cu knows the direct mode
    GF_REQUIRE(ok, "no device kernel for this (dimension, qdim, local dofs, family) combination");
    toc(0);
  }
  if (recompute) {
    if (!t->pat_valid) {
      tic(3); gf::build_pattern(t); toc(3);
      t->emask.release();  // constant-coefficient linear form: the pattern cannot move any more
      if (t->halo) gf::halo_build_maps(t);
    }
    if (!t->rc_ready && t->st.npairs) gf::recompute_prepare(t);  // nothing to plan for an empty element range
    if (do_r && do_t && t->rc_cols) {  // column kernel: tangent and R = K U in one pass
      tic(4); gf::recompute_assemble(t, U_dev, true, true); toc(4);
      return;
    }
    static const bool no_overlap = getenv("GFGPU_NO_OVERLAP") != nullptr;
    if (do_r && do_t && ctx->stream2 && !no_overlap) {
      // The tile kernel is a persistent 1-CTA-per-SM kernel bound by the shared-memory pipe; the residual path (a latency
      // bound per-element kernel + an HBM bound gather) runs NEXT to it on the side stream: fork after everything queued so
      // far, tile kernel first (its CTAs take their SMs), residual kernels into the registers / thread slots that are left
      // (64-thread blocks: 238 registers x 64 fit beside 106 x 384), join before anything that follows.
      GF_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
      GF_CUDA(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
      tic(4); gf::recompute_assemble(t, U_dev, true, false); toc(4);
      cudaStream_t s0 = ctx->stream;
      ctx->stream = ctx->stream2;
      t->rc_res_block = 64;
      try {
        tic(2); gf::recompute_assemble(t, U_dev, false, true); toc(2);
      } catch (...) {
        ctx->stream = s0;
        t->rc_res_block = 128;
        throw;
      }
      ctx->stream = s0;
      t->rc_res_block = 128;
      GF_CUDA(cudaEventRecord(ctx->ev_join, ctx->stream2));
      GF_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
      return;
    }
    if (do_r) { tic(2); gf::recompute_assemble(t, U_dev, false, true); toc(2); }
    if (do_t) { tic(4); gf::recompute_assemble(t, U_dev, true, false); toc(4); }
    return;
  }
  if (do_t) {
    if (!t->pat_valid) {
      tic(3); gf::build_pattern(t); toc(3);
      if (t->halo) gf::halo_build_maps(t);
      tic(1); gf::gather_tangent(t, false); toc(1);
    } else if (direct) {
      tic(1); gf::direct_finish(t); toc(1);
    } else if (!value_dependent) {
      tic(1); gf::gather_tangent(t, false); toc(1);
    } else {
      // value-dependent keep masks: the gather compares them with the pattern and raises a DEVICE flag.  No host round
      // trip here: the flag is read lazily (term_settle) by whoever consumes the tangent next, or by the next assembly.
      t->flag.zero();
      tic(1); gf::gather_tangent(t, true); toc(1);
      t->flag_pending = true;
      if (t->halo) term_settle(t);  // the exchange that follows reads pr at once
    }
  }
  if (do_r) { tic(2); gf::gather_residual(t); toc(2); }
}

}  // extern "C"
namespace gf {
double term_potential(gfgpu_term *t, const double *U_dev);  // potential.cu
void term_assemble_for_potential(gfgpu_term *t, const double *U_dev) { term_assemble(t, U_dev, GFGPU_RESIDUAL); }
void term_settle_pending(gfgpu_term *t) { term_settle(t); }
}  // namespace gf
extern "C" {

int gfgpu_term_potential_dev(gfgpu_term *t, const double *U_dev, double *E_host) {
  GF_API_BEGIN
  GF_REQUIRE(t && E_host, "null argument");
  GF_CUDA(cudaSetDevice(t->ctx->device));
  *E_host = gf::term_potential(t, U_dev);
  GF_API_END
}

int gfgpu_term_potential_host(gfgpu_term *t, const double *U_host, double *E_host) {
  GF_API_BEGIN
  GF_REQUIRE(t && E_host, "null argument");
  gfgpu_ctx *ctx = t->ctx;
  GF_CUDA(cudaSetDevice(ctx->device));
  const double *U_dev = nullptr;
  if (U_host) {
    if (t->Ubuf.n != (size_t)t->fem->ndof) t->Ubuf.alloc(ctx, t->fem->ndof);
    t->Ubuf.upload(U_host);
    U_dev = t->Ubuf.p;
  }
  *E_host = gf::term_potential(t, U_dev);
  GF_API_END
}

int gfgpu_term_assemble_dev(gfgpu_term *t, const double *U_dev, int order_mask) {
  GF_API_BEGIN
  GF_REQUIRE(t, "null term");
  term_assemble(t, U_dev, order_mask);
  GF_API_END
}

int gfgpu_term_assemble_host(gfgpu_term *t, const double *U_host, int order_mask, double *pr_host, double *R_host) {
  GF_API_BEGIN
  GF_REQUIRE(t, "null term");
  gfgpu_ctx *ctx = t->ctx;
  GF_CUDA(cudaSetDevice(ctx->device));
  const double *U_dev = nullptr;
  if (U_host) {
    if (t->Ubuf.n != (size_t)t->fem->ndof) t->Ubuf.alloc(ctx, t->fem->ndof);
    t->Ubuf.upload(U_host);
    U_dev = t->Ubuf.p;
  }
  term_assemble(t, U_dev, order_mask);
  term_settle(t);
  if ((order_mask & GFGPU_TANGENT) && pr_host) t->pr.download(pr_host);
  if ((order_mask & GFGPU_RESIDUAL) && R_host) t->R.download(R_host);
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  GF_API_END
}

int gfgpu_term_last_timings(gfgpu_term *t, float *out8) {
  GF_API_BEGIN
  GF_REQUIRE(t && out8, "null argument");
  GF_CUDA(cudaSetDevice(t->ctx->device));
  GF_CUDA(cudaStreamSynchronize(t->ctx->stream));
  for (int k = 0; k < 8; ++k) out8[k] = 0.f;
  for (int k = 0; k < 5; ++k)
    if (t->ev_used[k]) GF_CUDA(cudaEventElapsedTime(&out8[k], t->ev[2 * k], t->ev[2 * k + 1]));
  GF_API_END
}

int gfgpu_term_strategy(gfgpu_term *t) { return t ? t->strategy : -1; }
int gfgpu_term_kernel_kind(gfgpu_term *t) {
  if (!t) return -1;
  if (t->strategy != GFGPU_STRATEGY_RECOMPUTE && t->direct_ok == 1 && t->d_generation == t->generation) return 4;
  if (t->strategy != GFGPU_STRATEGY_RECOMPUTE || !t->rc_ready) return 0;
  return t->rc_cols ? 2 : t->rc_uni ? 3 : 1;
}

static bool settle_quiet(gfgpu_term *t) {
  try { term_settle(t); } catch (const std::exception &e) { gf::set_last_error(e.what()); return false; }
  return true;
}
int64_t gfgpu_term_nnz(gfgpu_term *t) { return (t && settle_quiet(t) && t->pat_valid) ? t->nnz : -1; }
int64_t gfgpu_term_nb_dof(gfgpu_term *t) { return t ? t->fem->ndof : -1; }
int64_t gfgpu_term_pattern_generation(gfgpu_term *t) { return (t && settle_quiet(t)) ? t->generation : -1; }

int gfgpu_term_csc_view(gfgpu_term *t, const int64_t **jc, const int32_t **ir, const double **pr) {
  GF_API_BEGIN
  GF_REQUIRE(t && t->pat_valid, "no assembled tangent");
  term_settle(t);
  if (jc) *jc = t->jc.p;
  if (ir) *ir = t->ir.p;
  if (pr) *pr = t->pr.p;
  GF_API_END
}

int gfgpu_term_residual_view(gfgpu_term *t, const double **R) {
  GF_API_BEGIN
  GF_REQUIRE(t && t->R.n, "no assembled residual");
  if (R) *R = t->R.p;
  GF_API_END
}

int gfgpu_term_export_csc_host(gfgpu_term *t, int64_t *jc, int32_t *ir, double *pr) {
  GF_API_BEGIN
  GF_REQUIRE(t && t->pat_valid, "no assembled tangent");
  GF_CUDA(cudaSetDevice(t->ctx->device));
  term_settle(t);
  if (jc) t->jc.download(jc);
  if (ir) t->ir.download(ir);
  if (pr) t->pr.download(pr);
  GF_CUDA(cudaStreamSynchronize(t->ctx->stream));
  GF_API_END
}

int gfgpu_term_export_residual_host(gfgpu_term *t, double *R) {
  GF_API_BEGIN
  GF_REQUIRE(t && t->R.n, "no assembled residual");
  GF_CUDA(cudaSetDevice(t->ctx->device));
  if (R) t->R.download(R);
  GF_CUDA(cudaStreamSynchronize(t->ctx->stream));
  GF_API_END
}


/* ------------------------------------------------------------------ multi-GPU halo */
int gfgpu_term_halo_begin(gfgpu_term *t, const double *U_dev, int64_t *touched_lo, int64_t *touched_hi) {
  GF_API_BEGIN
  GF_REQUIRE(t, "null term");
  // the halo is driven by the tangent's ghost columns; a load has none, and its interface contributions would be lost silently
  GF_REQUIRE(t->family != GFGPU_SOURCE && t->family != GFGPU_NORMAL_SOURCE,
             "halo: order-1-only terms (source / normal source) have no tangent to announce -- assemble the load on every rank's "
             "element range and sum the residual slices of the interface dofs yourself, or add it on one rank over the whole region");
  GF_CUDA(cudaSetDevice(t->ctx->device));
  // forget a previous halo, then build the LOCAL structure + pattern (one tangent pass)
  t->halo = false;
  t->halo_src.clear();
  t->halo_sends.clear();
  t->vJ.release(); t->vI.release(); t->vmask.release();
  t->st_valid = false; t->pat_valid = false; t->rc_ready = false;
  term_assemble(t, U_dev, GFGPU_TANGENT);
  // the pair records of RECOMPUTE dropped prel; the announcement below only needs pJ / pI / pmask
  int64_t lo = 0, hi = 0;
  if (t->st.nrnodes) {
    int32_t first = 0, last = 0;
    GF_CUDA(cudaMemcpyAsync(&first, t->st.rdof.p, sizeof(int32_t), cudaMemcpyDeviceToHost, t->ctx->stream));
    GF_CUDA(cudaMemcpyAsync(&last, t->st.rdof.p + t->st.nrnodes - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, t->ctx->stream));
    GF_CUDA(cudaStreamSynchronize(t->ctx->stream));
    lo = first; hi = (int64_t)last + t->fem->qdim;
  }
  if (touched_lo) *touched_lo = lo;
  if (touched_hi) *touched_hi = hi;
  GF_API_END
}

int gfgpu_term_halo_ghost_pairs(gfgpu_term *t, int64_t dof_lo, int64_t dof_hi, int64_t *n, int32_t *J_host,
                                int32_t *I_host, uint16_t *mask_host) {
  GF_API_BEGIN
  GF_REQUIRE(t && n && t->pat_valid, "halo_begin first");
  GF_CUDA(cudaSetDevice(t->ctx->device));
  gf::Structure &st = t->st;
  // pairs are sorted by (J, I): the columns [dof_lo, dof_hi) are a contiguous range of pairs
  std::vector<int32_t> pJ(st.npairs);
  st.pJ.download(pJ.data());
  GF_CUDA(cudaStreamSynchronize(t->ctx->stream));
  const int64_t a = std::lower_bound(pJ.begin(), pJ.end(), (int32_t)std::min<int64_t>(dof_lo, INT32_MAX)) - pJ.begin();
  const int64_t b = std::lower_bound(pJ.begin(), pJ.end(), (int32_t)std::min<int64_t>(dof_hi, INT32_MAX)) - pJ.begin();
  *n = b - a;
  if (J_host && I_host && mask_host && b > a) {
    std::copy(pJ.begin() + a, pJ.begin() + b, J_host);
    GF_CUDA(cudaMemcpyAsync(I_host, st.pI.p + a, (b - a) * sizeof(int32_t), cudaMemcpyDeviceToHost, t->ctx->stream));
    GF_CUDA(cudaMemcpyAsync(mask_host, t->pmask.p + a, (b - a) * sizeof(uint16_t), cudaMemcpyDeviceToHost, t->ctx->stream));
    GF_CUDA(cudaStreamSynchronize(t->ctx->stream));
  }
  GF_API_END
}

int gfgpu_term_halo_add_source(gfgpu_term *t, int src_rank, int64_t n, const int32_t *J_host, const int32_t *I_host,
                               const uint16_t *mask_host, int64_t r_lo, int64_t r_hi) {
  GF_API_BEGIN
  GF_REQUIRE(t && (n == 0 || (J_host && I_host && mask_host)), "null argument");
  GF_REQUIRE(t->halo_src.empty() || t->halo_src.back()->rank < src_rank, "sources must be added in ascending rank order");
  GF_REQUIRE(0 <= r_lo && r_lo <= r_hi && r_hi <= t->fem->ndof, "bad residual range");
  GF_CUDA(cudaSetDevice(t->ctx->device));
  const int Q = t->fem->qdim;
  std::unique_ptr<gfgpu_term::HaloSource> hs(new gfgpu_term::HaloSource);
  hs->rank = src_rank; hs->n = n; hs->r_lo = r_lo; hs->r_hi = r_hi;
  // layout of the source's segment = its CSC restricted to these columns: columns ascending, component by
  // component, rows ascending.  The announced pairs are sorted by (J, I).
  std::vector<uint32_t> soff((size_t)Q * std::max<int64_t>(n, 1), 0);
  int64_t run = 0;
  for (int64_t k = 0; k < n;) {
    int64_t k1 = k;
    while (k1 < n && J_host[k1] == J_host[k]) ++k1;
    for (int b = 0; b < Q; ++b)
      for (int64_t q = k; q < k1; ++q) {
        GF_REQUIRE(q == k || I_host[q] > I_host[q - 1], "announced pairs are not sorted by (J, I)");
        soff[(size_t)b * n + q] = (uint32_t)run;
        run += __builtin_popcount((mask_host[q] >> (b * Q)) & ((1u << Q) - 1));
      }
    GF_REQUIRE(k1 == n || J_host[k1] > J_host[k], "announced pairs are not sorted by (J, I)");
    k = k1;
  }
  GF_REQUIRE(run < (int64_t(1) << 32), "halo segment too large");
  hs->nvals = run;
  hs->J.alloc(t->ctx, n); hs->I.alloc(t->ctx, n); hs->mask.alloc(t->ctx, n); hs->soff.alloc(t->ctx, (size_t)Q * n);
  hs->J.upload(J_host); hs->I.upload(I_host); hs->mask.upload(mask_host); hs->soff.upload(soff.data());
  hs->rrecv.alloc(t->ctx, std::max<int64_t>(r_hi - r_lo, 1));
  GF_CUDA(cudaStreamSynchronize(t->ctx->stream));
  t->halo_src.push_back(std::move(hs));
  GF_API_END
}

int gfgpu_term_halo_commit(gfgpu_term *t, int64_t own_lo, int64_t own_hi) {
  GF_API_BEGIN
  GF_REQUIRE(t, "null term");
  GF_REQUIRE(0 <= own_lo && own_lo <= own_hi && own_hi <= t->fem->ndof, "bad owned range");
  GF_CUDA(cudaSetDevice(t->ctx->device));
  int64_t nv = 0;
  for (auto &hs : t->halo_src) nv += hs->n;
  t->vJ.alloc(t->ctx, nv); t->vI.alloc(t->ctx, nv); t->vmask.alloc(t->ctx, nv);
  int64_t at = 0;
  for (auto &hs : t->halo_src) {
    if (!hs->n) continue;
    GF_CUDA(cudaMemcpyAsync(t->vJ.p + at, hs->J.p, hs->n * sizeof(int32_t), cudaMemcpyDeviceToDevice, t->ctx->stream));
    GF_CUDA(cudaMemcpyAsync(t->vI.p + at, hs->I.p, hs->n * sizeof(int32_t), cudaMemcpyDeviceToDevice, t->ctx->stream));
    GF_CUDA(cudaMemcpyAsync(t->vmask.p + at, hs->mask.p, hs->n * sizeof(uint16_t), cudaMemcpyDeviceToDevice, t->ctx->stream));
    at += hs->n;
  }
  t->halo = true;
  t->own_lo = own_lo; t->own_hi = own_hi;
  t->st_valid = false; t->pat_valid = false; t->rc_ready = false;  // rebuilt with the virtual pairs at the next assemble
  GF_API_END
}

int gfgpu_term_halo_send_view(gfgpu_term *t, int64_t dof_lo, int64_t dof_hi, const double **pr_dev, int64_t *count,
                              const double **R_dev) {
  GF_API_BEGIN
  GF_REQUIRE(t && t->halo && t->pat_valid, "assemble after halo_commit first");
  GF_REQUIRE(0 <= dof_lo && dof_lo <= dof_hi && dof_hi <= t->fem->ndof, "bad dof range");
  GF_CUDA(cudaSetDevice(t->ctx->device));
  int64_t a = 0, b = 0;
  GF_CUDA(cudaMemcpyAsync(&a, t->jc.p + dof_lo, sizeof(int64_t), cudaMemcpyDeviceToHost, t->ctx->stream));
  GF_CUDA(cudaMemcpyAsync(&b, t->jc.p + dof_hi, sizeof(int64_t), cudaMemcpyDeviceToHost, t->ctx->stream));
  GF_CUDA(cudaStreamSynchronize(t->ctx->stream));
  if (pr_dev) *pr_dev = t->pr.p + a;
  if (count) *count = b - a;
  if (R_dev) *R_dev = t->R.n ? t->R.p + dof_lo : nullptr;
  GF_API_END
}

int gfgpu_term_halo_recv_view(gfgpu_term *t, int src_rank, double **pr_recv_dev, int64_t *count, double **R_recv_dev,
                              int64_t *r_count) {
  GF_API_BEGIN
  GF_REQUIRE(t && t->halo && t->pat_valid, "assemble after halo_commit first");
  for (auto &hs : t->halo_src)
    if (hs->rank == src_rank) {
      if (pr_recv_dev) *pr_recv_dev = hs->recv.p;
      if (count) *count = hs->nvals;
      if (R_recv_dev) *R_recv_dev = hs->rrecv.p;
      if (r_count) *r_count = hs->r_hi - hs->r_lo;
      return 0;
    }
  GF_REQUIRE(false, "unknown halo source rank");
  GF_API_END
}

int gfgpu_term_halo_accumulate(gfgpu_term *t, int order_mask) {
  GF_API_BEGIN
  GF_REQUIRE(t && t->halo && t->pat_valid, "assemble after halo_commit first");
  GF_CUDA(cudaSetDevice(t->ctx->device));
  gf::halo_accumulate(t, order_mask & GFGPU_TANGENT, order_mask & GFGPU_RESIDUAL);
  GF_API_END
}

int gfgpu_term_owned_range(gfgpu_term *t, int64_t *own_lo, int64_t *own_hi) {
  GF_API_BEGIN
  GF_REQUIRE(t, "null term");
  if (own_lo) *own_lo = t->halo ? t->own_lo : 0;
  if (own_hi) *own_hi = t->halo ? t->own_hi : t->fem->ndof;
  GF_API_END
}

}  // extern "C"
