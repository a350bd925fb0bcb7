// comm.cu -- multi-GPU halo exchange INSIDE the C ABI: one NCCL communicator per process (one process per GPU), the
// exchange of SURVEY 8(e) as ONE ncclGroup of ncclSend / ncclRecv on the context's stream followed by the owner's
// fixed-order accumulation.  The C++ drop-in (getfem_b200/shim) can therefore use several GPUs without Python.
//
// Replaces the reference's MPI_SUM_SPARSE_MATRIX / MPI_SUM_VECTOR of whole matrices / vectors
// (src/getfem_generic_assembly_workspace.cc:855-858, src/getfem/getfem_config.h:214-341) by point-to-point slices:
// a rank sends its partial values of the ghost columns (a contiguous slice of pr) and the matching residual slice to
// their owner; nobody reduces full-size objects.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2): libgfgpu.so has no link dependency on it, a process that
// already carries an NCCL (torch's) shares that copy, and single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "common.cuh"

namespace gf {

struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi &nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.h) break;
    }
    if (!api.h) return;
#define GF_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.h, name))
    GF_SYM(GetUniqueId, "ncclGetUniqueId");
    GF_SYM(CommInitRank, "ncclCommInitRank");
    GF_SYM(CommDestroy, "ncclCommDestroy");
    GF_SYM(GroupStart, "ncclGroupStart");
    GF_SYM(GroupEnd, "ncclGroupEnd");
    GF_SYM(Send, "ncclSend");
    GF_SYM(Recv, "ncclRecv");
    GF_SYM(GetErrorString, "ncclGetErrorString");
#undef GF_SYM
  });
  GF_REQUIRE(api.h && api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send &&
                 api.Recv && api.GetErrorString,
             "NCCL (libnccl.so.2) could not be loaded: the multi-GPU exchange needs it");
  return api;
}

#define GF_NCCL(call)                                                                                       \
  do {                                                                                                      \
    ncclResult_t r_ = (call);                                                                               \
    if (r_ != ncclSuccess)                                                                                  \
      throw gf::Error(std::string(__FILE__ ":" GF_STR(__LINE__) ": " #call ": ") + gf::nccl().GetErrorString(r_)); \
  } while (0)

}  // namespace gf

struct gfgpu_comm {
  gfgpu_ctx *ctx = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 0, rank = 0;
};

namespace gf {
void set_last_error(const std::string &s);  // api.cu
}

#define GF_API_BEGIN try {
#define GF_API_END                    \
  return 0;                           \
  }                                   \
  catch (const std::exception &e) {   \
    gf::set_last_error(e.what());     \
    return 1;                         \
  }                                   \
  catch (...) {                       \
    gf::set_last_error("unknown error"); \
    return 1;                         \
  }

extern "C" {

int gfgpu_comm_unique_id(char *id_out, int capacity) {
  GF_API_BEGIN
  GF_REQUIRE(id_out && capacity >= (int)sizeof(ncclUniqueId), "the id buffer needs GFGPU_COMM_ID_BYTES bytes");
  ncclUniqueId id;
  GF_NCCL(gf::nccl().GetUniqueId(&id));
  memcpy(id_out, &id, sizeof id);
  GF_API_END
}

int gfgpu_comm_create(gfgpu_ctx *ctx, int nranks, int rank, const char *id, gfgpu_comm **out) {
  GF_API_BEGIN
  GF_REQUIRE(ctx && id && out && nranks >= 1 && rank >= 0 && rank < nranks, "bad communicator arguments");
  GF_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<gfgpu_comm> c(new gfgpu_comm);
  c->ctx = ctx; c->nranks = nranks; c->rank = rank;
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof uid);
  GF_NCCL(gf::nccl().CommInitRank(&c->comm, nranks, uid, rank));
  *out = c.release();
  GF_API_END
}

int gfgpu_comm_destroy(gfgpu_comm *c) {
  if (!c) return 0;
  if (c->comm) gf::nccl().CommDestroy(c->comm);
  delete c;
  return 0;
}

int gfgpu_comm_rank(gfgpu_comm *c) { return c ? c->rank : -1; }
int gfgpu_comm_size(gfgpu_comm *c) { return c ? c->nranks : -1; }

int gfgpu_term_halo_add_send(gfgpu_term *t, int owner_rank, int64_t dof_lo, int64_t dof_hi, int64_t r_lo) {
  GF_API_BEGIN
  GF_REQUIRE(t && t->halo, "halo_commit first");
  GF_REQUIRE(owner_rank >= 0 && 0 <= dof_lo && dof_lo <= r_lo && r_lo <= dof_hi && dof_hi <= t->fem->ndof, "bad send range");
  gfgpu_term::HaloSend s;
  s.owner = owner_rank; s.dof_lo = dof_lo; s.dof_hi = dof_hi; s.r_lo = r_lo;
  t->halo_sends.push_back(s);
  t->halo_sends_generation = -1;
  GF_API_END
}

// One exchange step, after gfgpu_term_assemble_dev on the same stream: my ghost slices go to their owners, the slices of
// my sources land in their receive buffers (one NCCL group: the transfers of all peers overlap), then the owner adds the
// received parts in ascending source rank (halo_accumulate).  Asynchronous: everything is queued on the context's stream.
int gfgpu_term_halo_exchange(gfgpu_term *t, gfgpu_comm *c, int order_mask) {
  GF_API_BEGIN
  GF_REQUIRE(t && c && t->halo && t->pat_valid, "assemble after halo_commit first");
  GF_REQUIRE(c->ctx == t->ctx, "the communicator belongs to another context");
  gfgpu_ctx *ctx = t->ctx;
  GF_CUDA(cudaSetDevice(ctx->device));
  const bool do_t = order_mask & GFGPU_TANGENT, do_r = order_mask & GFGPU_RESIDUAL;
  if (t->halo_sends_generation != t->generation) {  // positions of the ghost slices in pr: once per pattern
    for (auto &s : t->halo_sends) {
      int64_t a = 0, b = 0;
      GF_CUDA(cudaMemcpyAsync(&a, t->jc.p + s.dof_lo, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
      GF_CUDA(cudaMemcpyAsync(&b, t->jc.p + s.dof_hi, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
      GF_CUDA(cudaStreamSynchronize(ctx->stream));
      s.pr_off = a;
      s.pr_cnt = b - a;
    }
    t->halo_sends_generation = t->generation;
  }
  gf::NcclApi &n = gf::nccl();
  GF_NCCL(n.GroupStart());
  for (const auto &s : t->halo_sends) {
    if (do_t && s.pr_cnt) GF_NCCL(n.Send(t->pr.p + s.pr_off, (size_t)s.pr_cnt, ncclDouble, s.owner, c->comm, ctx->stream));
    if (do_r && s.dof_hi > s.r_lo && t->R.n)
      GF_NCCL(n.Send(t->R.p + s.r_lo, (size_t)(s.dof_hi - s.r_lo), ncclDouble, s.owner, c->comm, ctx->stream));
  }
  for (auto &hs : t->halo_src) {
    if (do_t && hs->nvals) GF_NCCL(n.Recv(hs->recv.p, (size_t)hs->nvals, ncclDouble, hs->rank, c->comm, ctx->stream));
    if (do_r && hs->r_hi > hs->r_lo)
      GF_NCCL(n.Recv(hs->rrecv.p, (size_t)(hs->r_hi - hs->r_lo), ncclDouble, hs->rank, c->comm, ctx->stream));
  }
  GF_NCCL(n.GroupEnd());
  gf::halo_accumulate(t, do_t, do_r);
  GF_API_END
}

}  // extern "C"
