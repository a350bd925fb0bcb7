// recompute_cols.cu -- strategy RECOMPUTE, column kernel: tangent of a SCALAR-coefficient affine form (Laplace, mass) on
// LOW-ORDER elements, one thread per column node.
//
// The tile kernel (recompute_tiles.cu) gives a lane to a node PAIR; that fits P2 vector elasticity (100 pairs per
// element, ~3 contributions per pair, 54 DFMA each) but not P1: a P1 node has few pairs (7 on the regular tetrahedral
// mesh) with MANY contributions each (24 on the diagonal), most pairs become one-task "wide" pairs and the lanes idle
// (BASELINE config 2: 2.5 M tasks for 12 M entries).  Here the loop runs over the INCIDENCES of a column node J
// (element e, local index j), ascending element: the element's geometry row G_e (alpha a J B^T B, 6 doubles, from
// k_tile_geo) is loaded once and gives the nd contributions  G_e . M^{ji}  to the pairs (J, I_i), whose accumulators
// live in shared memory (slot-major, conflict free).  A one-byte slot per (incidence, i), found once at plan time, says
// which pair of the column the contribution belongs to.  Every pair still sums its contributions in ascending element
// order with the same FMA chain as the tile kernel: bitwise reproducible.
//
// Layout: a first version walked rsrc -> el -> rc_eg per lane; every load instruction then touched 32 different
// sectors and the kernel was bound by the L1 tag stage (1.51 ms on config 2, no faster than the tile kernel).  The
// incidence records (geometry row + j + slots: WPI 8-byte words) are therefore laid out once, at plan time, warp
// transposed (ELL per group of 32 consecutive column nodes, word-major, lane-minor): every load of the assembly
// kernel is one coalesced 256-byte access, with no indirection left.  Like the tile kernel's per-tile geometry blobs
// this replicates the geometry row of an element once per incidence (nd copies).
//
// Replaces, like the tile kernel, ga_exec + add_elem_matrix for these forms (C&E.cc:8750-8870, 4853-4936); the keep
// mask / CSC positions come from build_pattern (scatter.cu), the residual from k_affine_residual.
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace gf {

constexpr int CL_THREADS = 128;

// thread per column node: slot of every local contribution of its pairs -> desc[(incidence) * DS + i]
__global__ void k_cols_plan(const uint32_t *__restrict__ colstart, const uint32_t *__restrict__ cstart,
                            const uint32_t *__restrict__ csrc, const uint32_t *__restrict__ rstart,
                            const uint32_t *__restrict__ rsrc, int nd, int ds, uint32_t nlocal, int64_t nnodes,
                            uint8_t *__restrict__ desc, int *__restrict__ err) {
  const uint32_t nb = (uint32_t)(nd * nd);
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < nnodes; u += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t p0 = colstart[u], p1 = colstart[u + 1], r0 = rstart[u], r1 = rstart[u + 1];
    for (uint32_t p = p0; p < p1; ++p)
      for (uint32_t s = cstart[p], e = cstart[p + 1]; s < e; ++s) {
        const uint32_t c = csrc[s];
        if (c >= nlocal) continue;  // virtual (halo) contribution: arrives through halo_accumulate
        const uint32_t el = c / nb, r = c % nb, j = r / nd, i = r % nd, key = el * nd + j;
        uint32_t lo = r0, hi = r1;  // the incidence (el, j) of this node: rsrc is ascending
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if (rsrc[mid] < key) lo = mid + 1; else hi = mid;
        }
        if (lo >= r1 || rsrc[lo] != key) { atomicExch(err, 1); continue; }
        desc[(size_t)lo * ds + i] = (uint8_t)(p - p0);
      }
  }
}

struct ColArgs {
  const uint32_t *colstart;
  const int64_t *goff;   // per group of 32 column nodes: first ELL row; goff[g+1]-goff[g] rows (the longest incidence list)
  const double *ell;     // [row][word][lane], WPI words per row
  const int32_t *rdof;
  const uint16_t *pmask;
  const uint32_t *prel;
  const int64_t *jc;
  const int32_t *pI;   // row node dof0 of every pair (fused residual)
  const double *U;     // state vector (fused residual)
  double *R;
  const double *Mtab;
  int64_t nnodes, npairs, ngroups;
  double *pr;
};

template <int MT, int ND>
struct ColCfg {
  static constexpr int DW = (ND + 1 + 7) / 8;  // descriptor words: byte 0 = j, bytes 1..ND = pair slots (0xff = none)
  static constexpr int WPI = MT + DW;          // 8-byte words per incidence record
};

// thread per column node: its incidence records -> the warp-transposed ELL block of its group (padding rows: no slot)
template <int MT, int ND>
__global__ void k_cols_fill(const uint32_t *__restrict__ rstart, const uint32_t *__restrict__ rsrc,
                            const uint8_t *__restrict__ desc, int ds, const double *__restrict__ eg,
                            const int64_t *__restrict__ goff, int64_t nnodes, int64_t ngroups, double *__restrict__ ell) {
  using C = ColCfg<MT, ND>;
  const int lane = threadIdx.x & 31;
  for (int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; g < ngroups; g += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const int64_t u = g * 32 + lane;
    const uint32_t r0 = u < nnodes ? rstart[u] : 0u, cnt = u < nnodes ? rstart[u + 1] - r0 : 0u;
    const int64_t row0 = goff[g], len = goff[g + 1] - row0;
    double *base = ell + (size_t)row0 * C::WPI * 32 + lane;
    for (int64_t k = 0; k < len; ++k) {
      double *o = base + (size_t)k * C::WPI * 32;
      unsigned long long dw[C::DW];
#pragma unroll
      for (int d = 0; d < C::DW; ++d) dw[d] = ~0ull;
      if (k < cnt) {
        const uint32_t r = r0 + (uint32_t)k, elj = rsrc[r], el = elj / ND, j = elj % ND;
#pragma unroll
        for (int c = 0; c < MT; ++c) o[c * 32] = eg[(size_t)el * MT + c];
        dw[0] = (dw[0] & ~0xffull) | j;
#pragma unroll
        for (int i = 0; i < ND; ++i) {
          const unsigned long long q = desc[(size_t)r * ds + i];
          const int bi = i + 1;
          dw[bi >> 3] = (dw[bi >> 3] & ~(0xffull << (8 * (bi & 7)))) | (q << (8 * (bi & 7)));
        }
      } else {
#pragma unroll
        for (int c = 0; c < MT; ++c) o[c * 32] = 0.0;
        dw[0] &= ~0xffull;  // j = 0, every slot 0xff
      }
#pragma unroll
      for (int d = 0; d < C::DW; ++d) o[(MT + d) * 32] = __longlong_as_double((long long)dw[d]);
    }
  }
}

// WITH_R: the handled forms are symmetric, so the column J of K also gives R_J = sum_I K(I,J) U_I (= (K U)_J); the
// unmasked sums are used (the drop rule of add_elem_matrix does not apply to the vector assembly, C&E.cc:4669-4735)
template <int MT, int Q, int ND, bool WITH_R>
__global__ void __launch_bounds__(CL_THREADS)
k_cols(const ColArgs a) {
  using C = ColCfg<MT, ND>;
  extern __shared__ __align__(16) double cl_sm[];
  double *sM = cl_sm;                                // ND*ND x MT
  double *sAcc = cl_sm + ((ND * ND * MT + 1) & ~1);  // maxq x CL_THREADS, slot-major
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k < ND * ND * MT; k += CL_THREADS) sM[k] = a.Mtab[k];
  __syncthreads();
  double *acc = sAcc + tid;
  for (int64_t g = blockIdx.x * (int64_t)(CL_THREADS / 32) + warp; g < a.ngroups; g += (int64_t)gridDim.x * (CL_THREADS / 32)) {
    const int64_t u = g * 32 + lane;
    const bool live = u < a.nnodes;
    const uint32_t p0 = live ? a.colstart[u] : 0u, np = live ? a.colstart[u + 1] - p0 : 0u;
    for (uint32_t q = 0; q < np; ++q) acc[q * CL_THREADS] = 0.0;
    const int64_t row0 = a.goff[g], len = a.goff[g + 1] - row0;
    const double *base = a.ell + (size_t)row0 * C::WPI * 32 + lane;
    // the record of the next incidence is in flight while the current one is accumulated (the accumulation is a chain of
    // dependent shared-memory updates: ncu showed the warps waiting on the loads and on that chain in turn)
    double nw[C::WPI];
    if (len > 0) {
#pragma unroll
      for (int c = 0; c < C::WPI; ++c) nw[c] = __ldg(base + c * 32);
    }
    for (int64_t k = 0; k < len; ++k) {
      double G[MT];
#pragma unroll
      for (int c = 0; c < MT; ++c) G[c] = nw[c];
      unsigned long long dw[C::DW];
#pragma unroll
      for (int d = 0; d < C::DW; ++d) dw[d] = (unsigned long long)__double_as_longlong(nw[MT + d]);
      if (k + 1 < len) {
        const double *rec = base + (size_t)(k + 1) * C::WPI * 32;
#pragma unroll
        for (int c = 0; c < C::WPI; ++c) nw[c] = __ldg(rec + c * 32);
      }
      const double *Mj = sM + (size_t)(dw[0] & 0xffu) * ND * MT;
#pragma unroll
      for (int i = 0; i < ND; ++i) {
        const int bi = i + 1;
        const uint32_t q = (uint32_t)(dw[bi >> 3] >> (8 * (bi & 7))) & 0xffu;
        if (q == 0xffu) continue;  // padding row (or no local pair behind this contribution)
        double s = acc[q * CL_THREADS];
#pragma unroll
        for (int c = 0; c < MT; ++c) s += Mj[i * MT + c] * G[c];
        acc[q * CL_THREADS] = s;
      }
    }
    if (!live) continue;
    const int32_t J = a.rdof[u];
    double rs[Q];
#pragma unroll
    for (int b = 0; b < Q; ++b) rs[b] = 0.0;
    for (uint32_t q = 0; q < np; ++q) {
      const uint32_t p = p0 + q;
      const unsigned m = a.pmask[p];
      const double v = acc[q * CL_THREADS];
      if (WITH_R) {
        const int32_t I = a.pI[p];
#pragma unroll
        for (int b = 0; b < Q; ++b) rs[b] += v * __ldg(a.U + I + b);
      }
#pragma unroll
      for (int b = 0; b < Q; ++b) {
        int64_t pos = a.jc[J + b] + a.prel[(size_t)b * a.npairs + p];
#pragma unroll
        for (int c = 0; c < Q; ++c)
          if (m & (1u << (b * Q + c))) a.pr[pos++] = c == b ? v : 0.0;
      }
    }
    if (WITH_R) {
#pragma unroll
      for (int b = 0; b < Q; ++b) a.R[J + b] = rs[b];
    }
  }
}

static int cols_env() {
  const char *s = getenv("GFGPU_COLS");
  return s && *s ? atoi(s) : -1;
}

// GFGPU_COLS=0 never, 1 (default) for nd <= 4, 2 for every supported nd
bool recompute_cols_wanted(const gfgpu_term *t) {
  if (t->family != GFGPU_LAPLACE && t->family != GFGPU_MASS) return false;
  const int nd = t->fem->nd, mode = cols_env();
  if (mode == 0) return false;
  if (nd != 3 && nd != 4 && nd != 6 && nd != 10) return false;
  return nd <= 4 || mode == 2;
}

template <int MT, int ND>
static void fill_cols(gfgpu_term *t, const DevBuf<uint8_t> &desc, int ds, int64_t ngroups) {
  Structure &st = t->st;
  k_cols_fill<MT, ND><<<(unsigned)std::max<int64_t>(1, std::min<int64_t>((ngroups + 3) / 4, 148 * 32)), 128, 0, t->ctx->stream>>>(
      st.rstart.p, st.rsrc.p, desc.p, ds, t->rc_eg.p, t->rc_cgoff.p, st.ncolnodes, ngroups, t->rc_cell.p);
  GF_LAUNCH_CHECK();
}

// builds the incidence records; false = this structure does not fit (too many pairs in a column): use the tile kernel
bool recompute_cols_prepare(gfgpu_term *t, const std::vector<uint32_t> &colstart, const std::vector<uint32_t> &rstart) {
  Structure &st = t->st;
  gfgpu_ctx *ctx = t->ctx;
  const int nd = t->fem->nd, N = t->mesh->dim;
  const int MT = t->family == GFGPU_MASS ? 1 : N * (N + 1) / 2;
  uint32_t maxq = 0;
  for (int64_t k = 0; k < st.ncolnodes; ++k) maxq = std::max(maxq, colstart[k + 1] - colstart[k]);
  const size_t smem = ((size_t)((nd * nd * MT + 1) & ~1) + (size_t)maxq * CL_THREADS) * 8;
  if (maxq > 254 || smem > 160 * 1024) return false;
  if (st.ncontrib >= (int64_t(1) << 32) - 1 || st.nrinc >= (int64_t(1) << 32) - 1) return false;
  // ---- slot of every (incidence, local row node)
  const int ds = (nd + 3) & ~3;
  DevBuf<uint8_t> desc;
  desc.alloc(ctx, (size_t)st.nrinc * ds);
  GF_CUDA(cudaMemsetAsync(desc.p, 0xff, desc.n, ctx->stream));
  t->flag.zero();
  if (st.ncolnodes) {
    k_cols_plan<<<(unsigned)std::min<int64_t>((st.ncolnodes + 127) / 128, 148 * 32), 128, 0, ctx->stream>>>(
        st.colstart.p, st.cstart.p, st.csrc.p, st.rstart.p, st.rsrc.p, nd, ds, (uint32_t)st.ncontrib, st.ncolnodes,
        desc.p, (int *)t->flag.p);
    GF_LAUNCH_CHECK();
  }
  int32_t err = 0;
  t->flag.download(&err);
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  GF_REQUIRE(err == 0, "column kernel plan: a contribution has no incidence in its column node");
  // ---- ELL rows per group of 32 column nodes
  const int64_t ngroups = (st.ncolnodes + 31) / 32;
  std::vector<int64_t> goff(ngroups + 1, 0);
  for (int64_t g = 0; g < ngroups; ++g) {
    uint32_t len = 0;
    for (int64_t u = g * 32; u < std::min<int64_t>(st.ncolnodes, g * 32 + 32); ++u) len = std::max(len, rstart[u + 1] - rstart[u]);
    goff[g + 1] = goff[g] + len;
  }
  const int wpi = MT + (nd + 1 + 7) / 8;
  t->rc_cgoff.alloc(ctx, goff.size());
  t->rc_cgoff.upload(goff.data());
  t->rc_cell.alloc(ctx, (size_t)goff[ngroups] * wpi * 32);
  t->rc_cngroups = ngroups;
  if (ngroups) {
#define CF_CASE(MM, NDD) if (MT == MM && nd == NDD) fill_cols<MM, NDD>(t, desc, ds, ngroups);
#define CF_ND(MM) CF_CASE(MM, 3) CF_CASE(MM, 4) CF_CASE(MM, 6) CF_CASE(MM, 10)
    CF_ND(1) CF_ND(3) CF_ND(6)
#undef CF_ND
#undef CF_CASE
  }
  GF_CUDA(cudaStreamSynchronize(ctx->stream));  // goff (host) and desc (local) are read by the stream until here
  t->rc_cmaxq = (int)maxq;
  t->rc_cols = true;
  if (getenv("GFGPU_DEBUG"))
    fprintf(stderr, "[gfgpu] column kernel: %lld column nodes, %lld incidences in %lld ELL rows x 32 lanes x %d words, <= %u pairs per "
            "column, %zu B shared memory\n",
            (long long)st.ncolnodes, (long long)st.nrinc, (long long)goff[ngroups], wpi, maxq, smem);
  return true;
}

template <int MT, int Q, int ND>
static void launch_cols(gfgpu_term *t, const double *U, bool with_r) {
  Structure &st = t->st;
  ColArgs a;
  a.colstart = st.colstart.p;
  a.goff = t->rc_cgoff.p; a.ell = t->rc_cell.p;
  a.rdof = st.rdof.p;
  a.pmask = t->pmask.p; a.prel = t->prel.p; a.jc = t->jc.p;
  a.Mtab = t->rc_M.p;
  a.nnodes = st.ncolnodes; a.npairs = st.npairs; a.ngroups = t->rc_cngroups;
  a.pr = t->pr.p;
  a.pI = st.pI.p; a.U = U; a.R = t->R.p;
  if (with_r) t->R.zero();  // dofs outside the region; a NULL state leaves R = 0
  const size_t smem = ((size_t)((ND * ND * MT + 1) & ~1) + (size_t)t->rc_cmaxq * CL_THREADS) * 8;
  auto kern = with_r && U ? k_cols<MT, Q, ND, true> : k_cols<MT, Q, ND, false>;
  GF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = (a.ngroups + CL_THREADS / 32 - 1) / (CL_THREADS / 32);
  kern<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)t->ctx->sm_count * 64)), CL_THREADS, smem,
         t->ctx->stream>>>(a);
  GF_LAUNCH_CHECK();
}

// tangent, and with `with_r` the residual R = K U from the same pass
void recompute_cols_tangent(gfgpu_term *t, const double *U, bool with_r) {
  const int N = t->mesh->dim, nd = t->fem->nd, Q = t->fem->qdim;
  const int MT = t->family == GFGPU_MASS ? 1 : N * (N + 1) / 2;
#define CL_CASE(MM, QQ, NDD)                 \
  if (MT == MM && Q == QQ && nd == NDD) {    \
    launch_cols<MM, QQ, NDD>(t, U, with_r);  \
    return;                                  \
  }
#define CL_ND(MM, QQ) CL_CASE(MM, QQ, 3) CL_CASE(MM, QQ, 4) CL_CASE(MM, QQ, 6) CL_CASE(MM, QQ, 10)
  CL_ND(1, 1) CL_ND(1, 2) CL_ND(1, 3) CL_ND(3, 1) CL_ND(3, 2) CL_ND(6, 1) CL_ND(6, 3)
#undef CL_ND
#undef CL_CASE
  GF_REQUIRE(false, "no column kernel for this combination");
}

}  // namespace gf
