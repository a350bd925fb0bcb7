// scatter.cu -- replaces add_elem_matrix / add_to_coeff scatter into gmm::col_matrix<rsvector>
// (C&E.cc:4853-4936) and ga_instruction_vector_assembly_mf (C&E.cc:4669-4735) by a precomputed
// map + deterministic gather-sum:
//   structure : node pairs (J,I) present in the mesh graph and, per pair, the ascending list of
//               (element, local j, local i) contributions          [value independent, per fem]
//   pattern   : per pair a QxQ keep mask = OR over elements of "|K_e(r,c)| > 1e-14*max|K_e|"
//               (the reference's drop rule) -> CSC jc/ir identical to gmm::csc_matrix::init_with
//   gather    : pr[slot] = sum over the pair's contributions, ascending element id (the order the
//               reference's sequential ga_exec adds them), no atomics.
// CUB (radix sort / scan / select) is used for the symbolic phase only.
#include <cub/cub.cuh>

#include "common.cuh"

namespace gf {

void *cub_scratch(gfgpu_ctx *ctx, size_t bytes) {
  if (bytes > ctx->cub_tmp_bytes) {
    if (ctx->cub_tmp) {
      GF_CUDA(cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->cub_tmp);
      ctx->bytes -= (int64_t)ctx->cub_tmp_bytes;
    }
    size_t nb = bytes + bytes / 8 + 256;
    GF_CUDA(cudaMalloc(&ctx->cub_tmp, nb));
    ctx->cub_tmp_bytes = nb;
    ctx->bytes += (int64_t)nb;
  }
  return ctx->cub_tmp;
}

static inline int nbits(int64_t v) {
  int b = 1;
  while (b < 63 && (int64_t(1) << b) <= v) ++b;
  return b;
}

static inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

// ------------------------------------------------------------------ structure
__global__ void k_pair_keys(const int32_t *__restrict__ edof, int nd, int64_t e0, int64_t ncontrib, int bI,
                            uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  const int nb = nd * nd;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncontrib; c += (int64_t)gridDim.x * blockDim.x) {
    int64_t el = c / nb;
    int r = (int)(c % nb);
    int j = r / nd, i = r % nd;
    const int32_t *ed = edof + (e0 + el) * nd;
    keys[c] = ((uint64_t)(uint32_t)ed[j] << bI) | (uint64_t)(uint32_t)ed[i];
    vals[c] = (uint32_t)c;
  }
}

__global__ void k_node_keys(const int32_t *__restrict__ edof, int nd, int64_t e0, int64_t ninc,
                            uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ninc; c += (int64_t)gridDim.x * blockDim.x) {
    keys[c] = (uint32_t)edof[e0 * nd + c];
    vals[c] = (uint32_t)c;
  }
}

template <class K>
__global__ void k_head_flags(const K *__restrict__ keys, int64_t n, uint8_t *__restrict__ flags) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x)
    flags[s] = (s == 0 || keys[s] != keys[s - 1]) ? 1 : 0;
}

__global__ void k_pair_ids(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ cstart, int64_t npairs,
                           int bI, int32_t *__restrict__ pI, int32_t *__restrict__ pJ, uint8_t *__restrict__ colflag) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npairs; p += (int64_t)gridDim.x * blockDim.x) {
    uint64_t k = keys[cstart[p]];
    int32_t J = (int32_t)(k >> bI);
    pI[p] = (int32_t)(k & ((uint64_t(1) << bI) - 1));
    pJ[p] = J;
    bool head = p == 0;
    if (!head) head = (int32_t)(keys[cstart[p - 1]] >> bI) != J;
    colflag[p] = head ? 1 : 0;
  }
}

__global__ void k_node_ids(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ rstart, int64_t n,
                           int32_t *__restrict__ rdof) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x)
    rdof[p] = (int32_t)keys[rstart[p]];
}

__global__ void k_set_u32(uint32_t *p, uint32_t v) { *p = v; }

// positions of set flags -> out[0..count), count returned on host
static int64_t select_heads(gfgpu_ctx *ctx, const uint8_t *flags, int64_t n, uint32_t *out) {
  GF_REQUIRE(n < (int64_t(1) << 32), "too many contributions for 32-bit positions");
  DevBuf<int64_t> dcount;
  dcount.alloc(ctx, 1);
  cub::CountingInputIterator<uint32_t> it(0);
  size_t tb = 0;
  GF_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, it, flags, out, dcount.p, n, ctx->stream));
  void *tmp = cub_scratch(ctx, tb);
  GF_CUDA(cub::DeviceSelect::Flagged(tmp, tb, it, flags, out, dcount.p, n, ctx->stream));
  count_launch(2);
  int64_t h = 0;
  dcount.download(&h);
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  return h;
}

__global__ void k_virtual_keys(const int32_t *__restrict__ vJ, const int32_t *__restrict__ vI, int64_t nvirt, int bI,
                               int64_t first, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nvirt; k += (int64_t)gridDim.x * blockDim.x) {
    keys[first + k] = ((uint64_t)(uint32_t)vJ[k] << bI) | (uint64_t)(uint32_t)vI[k];
    vals[first + k] = (uint32_t)(first + k);
  }
}

void build_structure(gfgpu_ctx *ctx, const int32_t *edof, int nd, int64_t e0, int64_t e1, int64_t ndof, Structure &st,
                     const int32_t *vJ, const int32_t *vI, int64_t nvirt) {
  const int64_t ne = e1 - e0;
  const int64_t nb = (int64_t)nd * nd;
  st.e0 = e0;
  st.e1 = e1;
  st.ncontrib = ne * nb;
  st.nvirt = nvirt;
  GF_REQUIRE(st.ncontrib + nvirt < (int64_t(1) << 32), "element block too large: (ne*nd*nd) must stay below 2^32");
  const int bI = nbits(ndof);
  GF_REQUIRE(2 * bI <= 64, "ndof too large");
  cudaStream_t s = ctx->stream;
  const int B = 256;
  {  // ---- tangent: sort contributions by (J, I); the virtual ones (halo) go after the local ones of their pair
    const int64_t n = st.ncontrib + nvirt;
    DevBuf<uint64_t> k0, k1;
    DevBuf<uint32_t> v0;
    k0.alloc(ctx, n);
    k1.alloc(ctx, n);
    v0.alloc(ctx, n);
    st.csrc.alloc(ctx, n);
    if (n) {
      if (st.ncontrib) {
        k_pair_keys<<<min(grid_for(st.ncontrib, B), 148 * 16), B, 0, s>>>(edof, nd, e0, st.ncontrib, bI, k0.p, v0.p);
        GF_LAUNCH_CHECK();
      }
      if (nvirt) {
        k_virtual_keys<<<min(grid_for(nvirt, B), 148 * 16), B, 0, s>>>(vJ, vI, nvirt, bI, st.ncontrib, k0.p, v0.p);
        GF_LAUNCH_CHECK();
      }
      // explicit in/out buffers: the sorted values land in st.csrc
      size_t tb = 0;
      GF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k0.p, k1.p, v0.p, st.csrc.p, n, 0, 2 * bI, s));
      void *tmp = cub_scratch(ctx, tb);
      GF_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, k0.p, k1.p, v0.p, st.csrc.p, n, 0, 2 * bI, s));
      count_launch(2 * ((2 * bI + 7) / 8) + 1);
    }
    v0.release();
    k0.release();
    DevBuf<uint8_t> flags;
    flags.alloc(ctx, n);
    DevBuf<uint32_t> heads;
    heads.alloc(ctx, n + 1);
    if (n) {
      k_head_flags<uint64_t><<<min(grid_for(n, B), 148 * 16), B, 0, s>>>(k1.p, n, flags.p);
      GF_LAUNCH_CHECK();
    }
    st.npairs = n ? select_heads(ctx, flags.p, n, heads.p) : 0;
    st.cstart.alloc(ctx, st.npairs + 1);
    GF_CUDA(cudaMemcpyAsync(st.cstart.p, heads.p, st.npairs * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    k_set_u32<<<1, 1, 0, s>>>(st.cstart.p + st.npairs, (uint32_t)n);
    GF_LAUNCH_CHECK();
    st.pI.alloc(ctx, st.npairs);
    st.pJ.alloc(ctx, st.npairs);
    flags.alloc(ctx, st.npairs);
    if (st.npairs) {
      k_pair_ids<<<min(grid_for(st.npairs, B), 148 * 16), B, 0, s>>>(k1.p, st.cstart.p, st.npairs, bI, st.pI.p, st.pJ.p,
                                                                   flags.p);
      GF_LAUNCH_CHECK();
    }
    k1.release();
    heads.alloc(ctx, st.npairs + 1);
    st.ncolnodes = st.npairs ? select_heads(ctx, flags.p, st.npairs, heads.p) : 0;
    st.colstart.alloc(ctx, st.ncolnodes + 1);
    GF_CUDA(cudaMemcpyAsync(st.colstart.p, heads.p, st.ncolnodes * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    k_set_u32<<<1, 1, 0, s>>>(st.colstart.p + st.ncolnodes, (uint32_t)st.npairs);
    GF_LAUNCH_CHECK();
    GF_CUDA(cudaStreamSynchronize(s));
  }
  {  // ---- residual: sort (element, local node) incidences by node
    const int64_t n = ne * nd;
    st.nrinc = n;
    DevBuf<uint32_t> k0, k1, v0;
    k0.alloc(ctx, n);
    k1.alloc(ctx, n);
    v0.alloc(ctx, n);
    st.rsrc.alloc(ctx, n);
    if (n) {
      k_node_keys<<<min(grid_for(n, B), 148 * 16), B, 0, s>>>(edof, nd, e0, n, k0.p, v0.p);
      GF_LAUNCH_CHECK();
      size_t tb = 0;
      GF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k0.p, k1.p, v0.p, st.rsrc.p, n, 0, bI, s));
      void *tmp = cub_scratch(ctx, tb);
      GF_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, k0.p, k1.p, v0.p, st.rsrc.p, n, 0, bI, s));
      count_launch(2 * ((bI + 7) / 8) + 1);
    }
    DevBuf<uint8_t> flags;
    flags.alloc(ctx, n);
    DevBuf<uint32_t> heads;
    heads.alloc(ctx, n + 1);
    if (n) {
      k_head_flags<uint32_t><<<min(grid_for(n, B), 148 * 16), B, 0, s>>>(k1.p, n, flags.p);
      GF_LAUNCH_CHECK();
    }
    st.nrnodes = n ? select_heads(ctx, flags.p, n, heads.p) : 0;
    st.rstart.alloc(ctx, st.nrnodes + 1);
    GF_CUDA(cudaMemcpyAsync(st.rstart.p, heads.p, st.nrnodes * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    k_set_u32<<<1, 1, 0, s>>>(st.rstart.p + st.nrnodes, (uint32_t)n);
    GF_LAUNCH_CHECK();
    st.rdof.alloc(ctx, st.nrnodes);
    if (st.nrnodes) {
      k_node_ids<<<min(grid_for(st.nrnodes, B), 148 * 16), B, 0, s>>>(k1.p, st.rstart.p, st.nrnodes, st.rdof.p);
      GF_LAUNCH_CHECK();
    }
    GF_CUDA(cudaStreamSynchronize(s));
  }
}

// ------------------------------------------------------------------ pattern
__global__ void k_pair_masks(const uint32_t *__restrict__ cstart, const uint32_t *__restrict__ csrc,
                             const uint16_t *__restrict__ emask, int64_t npairs, uint32_t nlocal,
                             const uint16_t *__restrict__ vmask, uint16_t *__restrict__ pmask) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npairs; p += (int64_t)gridDim.x * blockDim.x) {
    unsigned m = 0;
    for (uint32_t s = cstart[p], e = cstart[p + 1]; s < e; ++s) {
      const uint32_t c = csrc[s];
      m |= c < nlocal ? emask[c] : vmask[c - nlocal];  // virtual contribution: the mask announced by its rank
    }
    pmask[p] = (uint16_t)m;
  }
}

// one thread per column node: running counts of kept entries per component column
template <int Q>
__global__ void k_column_scan(const uint32_t *__restrict__ colstart, const int32_t *__restrict__ pJ,
                              const uint16_t *__restrict__ pmask, int64_t ncol, int64_t npairs,
                              uint32_t *__restrict__ prel, int64_t *__restrict__ ctot) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < ncol; k += (int64_t)gridDim.x * blockDim.x) {
    uint32_t cnt[Q];
#pragma unroll
    for (int b = 0; b < Q; ++b) cnt[b] = 0;
    const uint32_t p0 = colstart[k], p1 = colstart[k + 1];
    for (uint32_t p = p0; p < p1; ++p) {
      unsigned m = pmask[p];
#pragma unroll
      for (int b = 0; b < Q; ++b) {
        prel[(size_t)b * npairs + p] = cnt[b];
        cnt[b] += __popc((m >> (b * Q)) & ((1u << Q) - 1));
      }
    }
    const int32_t J = pJ[p0];
#pragma unroll
    for (int b = 0; b < Q; ++b) ctot[J + b] = cnt[b];
  }
}

template <int Q>
__global__ void k_fill_ir(const int32_t *__restrict__ pI, const int32_t *__restrict__ pJ,
                          const uint16_t *__restrict__ pmask, const uint32_t *__restrict__ prel,
                          const int64_t *__restrict__ jc, int64_t npairs, int32_t *__restrict__ ir) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npairs; p += (int64_t)gridDim.x * blockDim.x) {
    const unsigned m = pmask[p];
    const int32_t I = pI[p], J = pJ[p];
#pragma unroll
    for (int b = 0; b < Q; ++b) {
      int64_t pos = jc[J + b] + prel[(size_t)b * npairs + p];
#pragma unroll
      for (int a = 0; a < Q; ++a)
        if (m & (1u << (b * Q + a))) ir[pos++] = I + a;
    }
  }
}

template <int Q>
static void build_pattern_t(gfgpu_term *t) {
  gfgpu_ctx *ctx = t->ctx;
  Structure &st = t->st;
  cudaStream_t s = ctx->stream;
  const int B = 256;
  const int64_t ndof = t->fem->ndof;
  t->pmask.alloc(ctx, st.npairs);
  t->prel.alloc(ctx, (size_t)Q * st.npairs);
  if (t->ctot.n != (size_t)ndof + 1) t->ctot.alloc(ctx, ndof + 1);
  if (t->jc.n != (size_t)ndof + 1) t->jc.alloc(ctx, ndof + 1);
  t->ctot.zero();
  if (st.npairs) {
    k_pair_masks<<<min(grid_for(st.npairs, B), 148 * 32), B, 0, s>>>(st.cstart.p, st.csrc.p, t->emask.p, st.npairs,
                                                                   (uint32_t)st.ncontrib, t->vmask.p, t->pmask.p);
    GF_LAUNCH_CHECK();
    k_column_scan<Q><<<min(grid_for(st.ncolnodes, B), 148 * 32), B, 0, s>>>(st.colstart.p, st.pJ.p, t->pmask.p,
                                                                          st.ncolnodes, st.npairs, t->prel.p, t->ctot.p);
    GF_LAUNCH_CHECK();
  }
  size_t tb = 0;
  GF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, t->ctot.p, t->jc.p, ndof + 1, s));
  void *tmp = cub_scratch(ctx, tb);
  GF_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, t->ctot.p, t->jc.p, ndof + 1, s));
  count_launch(2);
  int64_t nnz = 0;
  GF_CUDA(cudaMemcpyAsync(&nnz, t->jc.p + ndof, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  GF_CUDA(cudaStreamSynchronize(s));
  t->nnz = nnz;
  t->ir.alloc(ctx, nnz);
  t->pr.alloc(ctx, nnz);
  if (st.npairs) {
    k_fill_ir<Q><<<min(grid_for(st.npairs, B), 148 * 32), B, 0, s>>>(st.pI.p, st.pJ.p, t->pmask.p, t->prel.p, t->jc.p,
                                                                   st.npairs, t->ir.p);
    GF_LAUNCH_CHECK();
  }
  t->pat_valid = true;
  t->generation++;
}

void build_pattern(gfgpu_term *t) {
  switch (t->fem->qdim) {
    case 1: build_pattern_t<1>(t); break;
    case 2: build_pattern_t<2>(t); break;
    case 3: build_pattern_t<3>(t); break;
    default: GF_REQUIRE(false, "qdim must be 1, 2 or 3");
  }
}

// ------------------------------------------------------------------ halo maps (multi-GPU)
// thread per announced pair of one source: position in OUR pr of each value the source will send
template <int Q>
__global__ void k_halo_map(const int32_t *__restrict__ sJ, const int32_t *__restrict__ sI,
                           const uint16_t *__restrict__ smask, const uint32_t *__restrict__ soff, int64_t n,
                           const int32_t *__restrict__ pJ, const int32_t *__restrict__ pI,
                           const uint16_t *__restrict__ pmask, const uint32_t *__restrict__ prel,
                           const int64_t *__restrict__ jc, int64_t npairs, int64_t *__restrict__ map,
                           int *__restrict__ err) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const int32_t J = sJ[k], I = sI[k];
    int64_t lo = 0, hi = npairs;  // first pair >= (J, I)
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      const bool less = pJ[mid] < J || (pJ[mid] == J && pI[mid] < I);
      if (less) lo = mid + 1; else hi = mid;
    }
    if (lo >= npairs || pJ[lo] != J || pI[lo] != I) { *err = 1; continue; }
    const unsigned ms = smask[k], mm = pmask[lo];
    if (ms & ~mm) { *err = 2; continue; }
#pragma unroll
    for (int b = 0; b < Q; ++b) {
      const unsigned sb = (ms >> (b * Q)) & ((1u << Q) - 1), mb = (mm >> (b * Q)) & ((1u << Q) - 1);
      int64_t src = soff[(size_t)b * n + k];
      const int64_t dst = jc[J + b] + prel[(size_t)b * npairs + lo];
#pragma unroll
      for (int a = 0; a < Q; ++a)
        if (sb & (1u << a)) map[src++] = dst + __popc(mb & ((1u << a) - 1));
    }
  }
}

__global__ void k_halo_add(const double *__restrict__ recv, const int64_t *__restrict__ map, int64_t n,
                           double *__restrict__ pr) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    pr[map[k]] += recv[k];
}

__global__ void k_vec_add(const double *__restrict__ src, int64_t n, double *__restrict__ dst) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    dst[k] += src[k];
}

void halo_build_maps(gfgpu_term *t) {
  Structure &st = t->st;
  const int Q = t->fem->qdim;
  cudaStream_t s = t->ctx->stream;
  t->flag.zero();
  for (auto &hs : t->halo_src) {
    hs->map.alloc(t->ctx, std::max<int64_t>(hs->nvals, 1));
    hs->recv.alloc(t->ctx, std::max<int64_t>(hs->nvals, 1));
    if (!hs->n) continue;
    const int grid = min(grid_for(hs->n, 256), 148 * 32);
#define GF_MAP(QQ)                                                                                                     \
  k_halo_map<QQ><<<grid, 256, 0, s>>>(hs->J.p, hs->I.p, hs->mask.p, hs->soff.p, hs->n, st.pJ.p, st.pI.p, t->pmask.p,   \
                                      t->prel.p, t->jc.p, st.npairs, hs->map.p, (int *)t->flag.p)
    if (Q == 1) GF_MAP(1);
    else if (Q == 2) GF_MAP(2);
    else GF_MAP(3);
#undef GF_MAP
    GF_LAUNCH_CHECK();
  }
  int32_t err = 0;
  t->flag.download(&err);
  GF_CUDA(cudaStreamSynchronize(s));
  GF_REQUIRE(err == 0, "halo: an announced pair is missing from the merged pattern (code " + std::to_string(err) + ")");
}

void halo_accumulate(gfgpu_term *t, bool do_t, bool do_r) {
  cudaStream_t s = t->ctx->stream;
  for (auto &hs : t->halo_src) {  // ascending source rank: fixed summation order
    if (do_t && hs->nvals) {
      k_halo_add<<<min(grid_for(hs->nvals, 256), 148 * 32), 256, 0, s>>>(hs->recv.p, hs->map.p, hs->nvals, t->pr.p);
      GF_LAUNCH_CHECK();
    }
    if (do_r && hs->r_hi > hs->r_lo) {
      const int64_t n = hs->r_hi - hs->r_lo;
      k_vec_add<<<min(grid_for(n, 256), 148 * 32), 256, 0, s>>>(hs->rrecv.p, n, t->R.p + hs->r_lo);
      GF_LAUNCH_CHECK();
    }
  }
}

// ------------------------------------------------------------------ gather
// One thread per node pair: sums the QxQ blocks of its contributions from the stage (ascending
// element id) and writes the kept entries to their CSC slots.  With CHECK it also re-derives
// the pair's keep mask from the fresh element masks and raises `flag` when the pattern moved.
template <int Q, bool CHECK>
__global__ void __launch_bounds__(256)
k_gather(const uint32_t *__restrict__ cstart, const uint32_t *__restrict__ csrc, const int32_t *__restrict__ pJ,
         const uint16_t *__restrict__ pmask, const uint32_t *__restrict__ prel, const int64_t *__restrict__ jc,
         const double *__restrict__ stage, const uint16_t *__restrict__ emask, int nd, int64_t npairs,
         uint32_t nlocal, const uint16_t *__restrict__ vmask, double *__restrict__ pr, int *__restrict__ flag) {
  const int nb = nd * nd, s1 = nd * Q;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npairs; p += (int64_t)gridDim.x * blockDim.x) {
    double acc[Q * Q];
#pragma unroll
    for (int m = 0; m < Q * Q; ++m) acc[m] = 0.0;
    unsigned mnew = 0;
    for (uint32_t s = cstart[p], e = cstart[p + 1]; s < e; ++s) {
      const uint32_t c = csrc[s];
      if (c >= nlocal) {  // virtual (halo) contribution: its values arrive through gfgpu_term_halo_accumulate
        if (CHECK) mnew |= vmask[c - nlocal];
        continue;
      }
      const uint32_t el = c / nb, r = c % nb;
      const int j = r / nd, i = r % nd;
      const double *src = stage + (size_t)el * s1 * s1 + (size_t)(j * Q) * s1 + i * Q;
#pragma unroll
      for (int b = 0; b < Q; ++b)
#pragma unroll
        for (int a = 0; a < Q; ++a) acc[b * Q + a] += src[b * s1 + a];
      if (CHECK) mnew |= emask[c];
    }
    const unsigned m = pmask[p];
    if (CHECK && mnew != m) atomicExch(flag, 1);
    const int32_t J = pJ[p];
#pragma unroll
    for (int b = 0; b < Q; ++b) {
      int64_t pos = jc[J + b] + prel[(size_t)b * npairs + p];
#pragma unroll
      for (int a = 0; a < Q; ++a)
        if (m & (1u << (b * Q + a))) pr[pos++] = acc[b * Q + a];
    }
  }
}

// ---- scalar fems (Q = 1), linear forms: most node pairs of a high-order element have ONE contribution (both nodes interior
// to the same element), and for Q = 1 the contribution id IS the index of the value in the stage (nb = nd^2 = s1^2).  Such
// entries are copied entry-wise through a 4-byte source index (20 B of traffic per entry instead of ~34 B of pair
// metadata); the pairs with several contributions (or halo parts) keep the ordered per-pair sum.
__global__ void k_g1_build(const uint32_t *__restrict__ cstart, const uint32_t *__restrict__ csrc, const int32_t *__restrict__ pJ,
                           const uint16_t *__restrict__ pmask, const uint32_t *__restrict__ prel, const int64_t *__restrict__ jc,
                           int64_t npairs, uint32_t nlocal, uint32_t *__restrict__ src, uint32_t *__restrict__ multi,
                           unsigned long long *__restrict__ nmulti) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npairs; p += (int64_t)gridDim.x * blockDim.x) {
    if (!(pmask[p] & 1u)) continue;  // dropped entry: no slot
    const int64_t pos = jc[pJ[p]] + prel[p];
    const uint32_t s0 = cstart[p], cnt = cstart[p + 1] - s0;
    if (cnt == 1 && csrc[s0] < nlocal) {
      src[pos] = csrc[s0];
    } else {
      src[pos] = 0xffffffffu;
      multi[atomicAdd(nmulti, 1ull)] = (uint32_t)p;  // any order: every listed pair is summed on its own
    }
  }
}

__global__ void k_g1_copy(const uint32_t *__restrict__ src, const double *__restrict__ stage, int64_t nnz, double *__restrict__ pr) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t s = src[k];
    if (s != 0xffffffffu) pr[k] = stage[s];
  }
}

__global__ void k_g1_multi(const uint32_t *__restrict__ multi, int64_t nmulti, const uint32_t *__restrict__ cstart,
                           const uint32_t *__restrict__ csrc, const int32_t *__restrict__ pJ, const uint32_t *__restrict__ prel,
                           const int64_t *__restrict__ jc, const double *__restrict__ stage, uint32_t nlocal,
                           double *__restrict__ pr) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nmulti; k += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t p = multi[k];
    double acc = 0.0;
    for (uint32_t s = cstart[p], e = cstart[p + 1]; s < e; ++s) {
      const uint32_t c = csrc[s];
      if (c < nlocal) acc += stage[c];  // ascending element id; virtual (halo) parts arrive through halo_accumulate
    }
    pr[jc[pJ[p]] + prel[p]] = acc;
  }
}

static void g1_prepare(gfgpu_term *t);
static bool gather_tangent_q1_fast(gfgpu_term *t) {
  Structure &st = t->st;
  gfgpu_ctx *ctx = t->ctx;
  if (t->nnz >= (int64_t(1) << 32) - 1 || st.ncontrib >= (int64_t(1) << 32) - 1 || st.npairs >= (int64_t(1) << 32) - 1) return false;
  if (t->fem->nd < 27) return false;  // low-order elements: few single-contribution pairs, the pair kernel is as good
  const int B = 256;
  g1_prepare(t);
  k_g1_copy<<<min(grid_for(t->nnz, B), 148 * 64), B, 0, ctx->stream>>>(t->g1_src.p, t->stage.p, t->nnz, t->pr.p);
  GF_LAUNCH_CHECK();
  if (t->g1_nmulti) {
    k_g1_multi<<<min(grid_for(t->g1_nmulti, B), 148 * 64), B, 0, ctx->stream>>>(t->g1_multi.p, t->g1_nmulti, st.cstart.p, st.csrc.p,
                                                                              st.pJ.p, t->prel.p, t->jc.p, t->stage.p,
                                                                              (uint32_t)st.ncontrib, t->pr.p);
    GF_LAUNCH_CHECK();
  }
  return true;
}

// single-contribution sources and the list of the other pairs, once per pattern
static void g1_prepare(gfgpu_term *t) {
  Structure &st = t->st;
  gfgpu_ctx *ctx = t->ctx;
  const int B = 256;
  if (t->g1_generation != t->generation) {
    t->g1_src.alloc(ctx, t->nnz);
    t->g1_multi.alloc(ctx, st.npairs);
    DevBuf<int64_t> cnt;
    cnt.alloc(ctx, 1);
    cnt.zero();
    k_g1_build<<<min(grid_for(st.npairs, B), 148 * 64), B, 0, ctx->stream>>>(st.cstart.p, st.csrc.p, st.pJ.p, t->pmask.p, t->prel.p,
                                                                          t->jc.p, st.npairs, (uint32_t)st.ncontrib, t->g1_src.p,
                                                                          t->g1_multi.p, (unsigned long long *)cnt.p);
    GF_LAUNCH_CHECK();
    cnt.download(&t->g1_nmulti);
    GF_CUDA(cudaStreamSynchronize(ctx->stream));
    if (t->g1_nmulti > 1) {  // the list was filled through an atomic counter: back to pair (= CSC) order, for the locality of k_g1_multi
      DevBuf<uint32_t> sorted;
      sorted.alloc(ctx, t->g1_nmulti);
      size_t tb = 0;
      GF_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, t->g1_multi.p, sorted.p, (int64_t)t->g1_nmulti, 0, 32, ctx->stream));
      void *tmp = cub_scratch(ctx, tb);
      GF_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tb, t->g1_multi.p, sorted.p, (int64_t)t->g1_nmulti, 0, 32, ctx->stream));
      count_launch(4);
      GF_CUDA(cudaStreamSynchronize(ctx->stream));
      t->g1_multi.release();
      std::swap(t->g1_multi.p, sorted.p);
      std::swap(t->g1_multi.n, sorted.n);
      std::swap(t->g1_multi.ctx, sorted.ctx);
    } else if (!t->g1_nmulti) {
      t->g1_multi.release();
    }
    t->g1_generation = t->generation;
  }
}

// ---- direct mode (Q = 1, sum-factorised element kernels): the element kernel writes every entry straight to its CSC slot.
// A fixed pattern fixes the destination of every local contribution c (= its index in the stage, Q = 1):
//   single-contribution pair  -> slot[c] = position in pr (a plain store from the element kernel, no stage, no gather);
//   pair with several contributions (or halo parts) -> slot[c] = nnz + place in a compact stage; the ordered sum
//     (ascending element id, as everywhere) is a small second pass over those pairs only;
//   pair outside the pattern -> 0xffffffff.
__global__ void k_slot_single(const uint32_t *__restrict__ cstart, const uint32_t *__restrict__ csrc, const int32_t *__restrict__ pJ,
                              const uint16_t *__restrict__ pmask, const uint32_t *__restrict__ prel, const int64_t *__restrict__ jc,
                              int64_t npairs, uint32_t nlocal, uint32_t *__restrict__ slot) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npairs; p += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t s0 = cstart[p], s1 = cstart[p + 1];
    if (!(pmask[p] & 1u)) {
      for (uint32_t s = s0; s < s1; ++s)
        if (csrc[s] < nlocal) slot[csrc[s]] = 0xffffffffu;
    } else if (s1 - s0 == 1 && csrc[s0] < nlocal) {
      slot[csrc[s0]] = (uint32_t)(jc[pJ[p]] + prel[p]);
    }
  }
}
__global__ void k_slot_multi_count(const uint32_t *__restrict__ multi, int64_t nmulti, const uint32_t *__restrict__ cstart,
                                   const uint32_t *__restrict__ csrc, uint32_t nlocal, uint32_t *__restrict__ cnt) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k <= nmulti; k += (int64_t)gridDim.x * blockDim.x) {
    uint32_t n = 0;
    if (k < nmulti) {
      const uint32_t p = multi[k];
      for (uint32_t s = cstart[p], e = cstart[p + 1]; s < e; ++s) n += csrc[s] < nlocal;
    }
    cnt[k] = n;
  }
}
__global__ void k_slot_multi(const uint32_t *__restrict__ multi, int64_t nmulti, const uint32_t *__restrict__ moff,
                             const uint32_t *__restrict__ cstart, const uint32_t *__restrict__ csrc, uint32_t nlocal, uint32_t nnz,
                             uint32_t *__restrict__ slot) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nmulti; k += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t p = multi[k];
    uint32_t o = nnz + moff[k];
    for (uint32_t s = cstart[p], e = cstart[p + 1]; s < e; ++s)
      if (csrc[s] < nlocal) slot[csrc[s]] = o++;
  }
}
// (the pair id is replaced by the CSC position once the slots are built: the sum then needs no pair metadata at all)
__global__ void k_multi_pos(uint32_t *__restrict__ multi, int64_t nmulti, const int32_t *__restrict__ pJ,
                            const uint32_t *__restrict__ prel, const int64_t *__restrict__ jc) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nmulti; k += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t p = multi[k];
    multi[k] = (uint32_t)(jc[pJ[p]] + prel[p]);
  }
}
__global__ void k_multi_sum(const uint32_t *__restrict__ mpos, int64_t nmulti, const uint32_t *__restrict__ moff,
                            const double *__restrict__ mstage, double *__restrict__ pr) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nmulti; k += (int64_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (uint32_t o = moff[k], e = moff[k + 1]; o < e; ++o) acc += mstage[o];  // ascending element id
    pr[mpos[k]] = acc;
  }
}

// CTA per element: its nb slots sorted ascending, with the local index k of each (uint16).  Scattered 8-byte stores in
// LOCAL order cost the element kernel 2x its whole run time (profiles/round2_c5_direct_experiments.txt: the rows a column
// receives from one element are runs of 4 in local order, unaligned to the 32-byte sectors); in DESTINATION order the same
// stores are long contiguous runs -- an interior column of a Q4 element is one 1000-byte segment of pr.
template <int IPT>
__global__ void __launch_bounds__(512) k_slot_sort(uint32_t *__restrict__ slot, uint16_t *__restrict__ kloc, int nb) {
  using Sort = cub::BlockRadixSort<uint32_t, 512, IPT, uint32_t>;
  extern __shared__ __align__(16) unsigned char sort_raw[];
  typename Sort::TempStorage &tmp = *reinterpret_cast<typename Sort::TempStorage *>(sort_raw);
  uint32_t keys[IPT], vals[IPT];
  uint32_t *sl = slot + (size_t)blockIdx.x * nb;
  uint16_t *kl = kloc + (size_t)blockIdx.x * nb;
#pragma unroll
  for (int r = 0; r < IPT; ++r) {
    const int k = threadIdx.x * IPT + r;
    keys[r] = k < nb ? sl[k] : 0xffffffffu;
    vals[r] = (uint32_t)k;
  }
  Sort(tmp).Sort(keys, vals);
#pragma unroll
  for (int r = 0; r < IPT; ++r) {
    const int k = threadIdx.x * IPT + r;
    if (k < nb) { sl[k] = keys[r]; kl[k] = (uint16_t)vals[r]; }
  }
}

bool direct_prepare(gfgpu_term *t) {
  Structure &st = t->st;
  gfgpu_ctx *ctx = t->ctx;
  if (t->d_generation == t->generation) return t->direct_ok == 1;
  t->d_generation = t->generation;
  t->direct_ok = -1;
  if (t->fem->qdim != 1 || !st.npairs) return false;
  if (t->nnz >= (int64_t(1) << 32) - 1 || st.ncontrib >= (int64_t(1) << 32) - 1 || st.npairs >= (int64_t(1) << 32) - 1) return false;
  const int B = 256;
  g1_prepare(t);
  // places of the multi-contribution entries in the compact stage
  int64_t total = 0;
  if (t->g1_nmulti) {
    DevBuf<uint32_t> cnt;
    cnt.alloc(ctx, t->g1_nmulti + 1);
    t->moff.alloc(ctx, t->g1_nmulti + 1);
    k_slot_multi_count<<<min(grid_for(t->g1_nmulti + 1, B), 148 * 64), B, 0, ctx->stream>>>(t->g1_multi.p, t->g1_nmulti, st.cstart.p,
                                                                                         st.csrc.p, (uint32_t)st.ncontrib, cnt.p);
    GF_LAUNCH_CHECK();
    size_t tb = 0;
    GF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, t->moff.p, (int64_t)t->g1_nmulti + 1, ctx->stream));
    void *tmp = cub_scratch(ctx, tb);
    GF_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt.p, t->moff.p, (int64_t)t->g1_nmulti + 1, ctx->stream));
    count_launch(2);
    uint32_t tot32 = 0;
    GF_CUDA(cudaMemcpyAsync(&tot32, t->moff.p + t->g1_nmulti, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    GF_CUDA(cudaStreamSynchronize(ctx->stream));
    total = tot32;
  }
  if (t->nnz + total >= (int64_t(1) << 32) - 1) return false;
  t->dslot.alloc(ctx, st.ncontrib);
  GF_CUDA(cudaMemsetAsync(t->dslot.p, 0xff, (size_t)st.ncontrib * sizeof(uint32_t), ctx->stream));
  k_slot_single<<<min(grid_for(st.npairs, B), 148 * 64), B, 0, ctx->stream>>>(st.cstart.p, st.csrc.p, st.pJ.p, t->pmask.p, t->prel.p,
                                                                           t->jc.p, st.npairs, (uint32_t)st.ncontrib, t->dslot.p);
  GF_LAUNCH_CHECK();
  if (t->g1_nmulti) {
    k_slot_multi<<<min(grid_for(t->g1_nmulti, B), 148 * 64), B, 0, ctx->stream>>>(t->g1_multi.p, t->g1_nmulti, t->moff.p, st.cstart.p,
                                                                               st.csrc.p, (uint32_t)st.ncontrib, (uint32_t)t->nnz,
                                                                               t->dslot.p);
    GF_LAUNCH_CHECK();
  }
  {  // destination order inside every element
    const int nb = t->fem->nd * t->fem->nd;
    const int64_t nel = st.ncontrib / nb;
    GF_REQUIRE(nb <= 65536, "direct mode: element matrices above 65536 entries are not handled");
    t->dkloc.alloc(ctx, st.ncontrib);
    auto launch = [&](auto kern, size_t smem) {
      GF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<(unsigned)nel, 512, smem, ctx->stream>>>(t->dslot.p, t->dkloc.p, nb);
    };
    if (nb <= 512 * 8) launch(k_slot_sort<8>, sizeof(cub::BlockRadixSort<uint32_t, 512, 8, uint32_t>::TempStorage));
    else if (nb <= 512 * 31) launch(k_slot_sort<31>, sizeof(cub::BlockRadixSort<uint32_t, 512, 31, uint32_t>::TempStorage));
    else return false;
    GF_LAUNCH_CHECK();
  }
  t->mstage.alloc(ctx, std::max<int64_t>(total, 1));
  // the full stage and the entry-wise sources are not needed while the pattern stands; the list of the shared pairs
  // becomes the list of their CSC positions (dmpos)
  t->dmpos.release();
  if (t->g1_nmulti) {
    k_multi_pos<<<min(grid_for(t->g1_nmulti, B), 148 * 64), B, 0, ctx->stream>>>(t->g1_multi.p, t->g1_nmulti, st.pJ.p, t->prel.p, t->jc.p);
    GF_LAUNCH_CHECK();
    std::swap(t->dmpos.p, t->g1_multi.p);
    std::swap(t->dmpos.n, t->g1_multi.n);
    std::swap(t->dmpos.ctx, t->g1_multi.ctx);
  }
  t->d_nmulti = t->g1_nmulti;
  t->g1_src.release();
  t->g1_multi.release();
  t->g1_generation = -1;
  t->stage.release();
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  t->direct_ok = 1;
  return true;
}

void direct_finish(gfgpu_term *t) {
  if (!t->d_nmulti) return;
  const int B = 256;
  k_multi_sum<<<min(grid_for(t->d_nmulti, B), 148 * 64), B, 0, t->ctx->stream>>>(t->dmpos.p, t->d_nmulti, t->moff.p, t->mstage.p, t->pr.p);
  GF_LAUNCH_CHECK();
}

template <int Q>
static void gather_tangent_t(gfgpu_term *t, bool check) {
  Structure &st = t->st;
  if (!st.npairs) return;
  if (Q == 1 && !check && gather_tangent_q1_fast(t)) return;
  const int B = 256;
  int grid = min(grid_for(st.npairs, B), 148 * 64);
  if (check)
    k_gather<Q, true><<<grid, B, 0, t->ctx->stream>>>(st.cstart.p, st.csrc.p, st.pJ.p, t->pmask.p, t->prel.p, t->jc.p,
                                                      t->stage.p, t->emask.p, t->fem->nd, st.npairs,
                                                      (uint32_t)st.ncontrib, t->vmask.p, t->pr.p, (int *)t->flag.p);
  else
    k_gather<Q, false><<<grid, B, 0, t->ctx->stream>>>(st.cstart.p, st.csrc.p, st.pJ.p, t->pmask.p, t->prel.p, t->jc.p,
                                                       t->stage.p, t->emask.p, t->fem->nd, st.npairs,
                                                       (uint32_t)st.ncontrib, t->vmask.p, t->pr.p, (int *)t->flag.p);
  GF_LAUNCH_CHECK();
}

void gather_tangent(gfgpu_term *t, bool check) {
  switch (t->fem->qdim) {
    case 1: gather_tangent_t<1>(t, check); break;
    case 2: gather_tangent_t<2>(t, check); break;
    case 3: gather_tangent_t<3>(t, check); break;
    default: GF_REQUIRE(false, "qdim must be 1, 2 or 3");
  }
}

// one thread per (node, component): R[dof0 + a] = sum of staged element residuals, ascending element id
__global__ void k_gather_residual(const uint32_t *__restrict__ rstart, const uint32_t *__restrict__ rsrc,
                                  const int32_t *__restrict__ rdof, const double *__restrict__ rstage, int Q,
                                  int64_t nrnodes, double *__restrict__ R) {
  const int64_t n = nrnodes * Q;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = idx / Q;
    const int a = (int)(idx % Q);
    double s = 0.0;
    for (uint32_t k = rstart[u], e = rstart[u + 1]; k < e; ++k) s += rstage[(size_t)rsrc[k] * Q + a];
    R[rdof[u] + a] = s;
  }
}

void gather_residual(gfgpu_term *t) {
  Structure &st = t->st;
  t->R.zero();
  if (!st.nrnodes) return;
  const int B = 256;
  const int Q = t->fem->qdim;
  k_gather_residual<<<min(grid_for(st.nrnodes * Q, B), 148 * 64), B, 0, t->ctx->stream>>>(
      st.rstart.p, st.rsrc.p, st.rdof.p, t->rstage.p, Q, st.nrnodes, t->R.p);
  GF_LAUNCH_CHECK();
}

}  // namespace gf
