// recompute_tiles.cu -- strategy RECOMPUTE, tile kernel: tangent (and fused residual) of an
// affine-geometry, constant-coefficient bilinear form assembled per NONZERO, no element matrix in HBM.
//
// Same mathematics as recompute.cu (reference tensors M^{ij}; SURVEY appendix B, getfem_models.cc:6112-6113,
// bgeot_geometric_trans.h:462-468), different work decomposition:
//
//   tile  = a run of consecutive column nodes (<= cap_inc element incidences, <= cap_pairs node pairs).
//           The tile's DISTINCT elements are staged once in shared memory (elasticity: sqrt(|alpha| J) B,
//           9 doubles; Laplace: alpha a J B^T B, 6 doubles; mass: alpha a J) together with its pair records,
//           by a producer warp (TMA bulk copies + LDGSTS gather on an mbarrier), double buffered.
//   task  = 32 node pairs of the tile with (almost) the same number of contributions: the tile's pairs are
//           counting-sorted by contribution count, so that one warp runs `cnt` steps with every lane busy and
//           all lanes flush together (no divergence).  Lane = pair: the Q x Q block is accumulated in registers,
//           one step = one (element, j, i) contribution = T += B~ M^{ji} B~^T (54 DFMA), operands from shared
//           memory through a 16-bit descriptor (slot | j*nd+i) carried in the pair record.  Short pairs of a
//           task are padded with a descriptor that points at an all-zero geometry slot (adds +0.0).
//   flush = the elasticity combination lambda T + mu T^T + mu tr(T) I is applied ONCE per pair (it is linear);
//           kept entries are written to an image of the tile's CSC segment in shared memory (the segment of a run of
//           column nodes is contiguous), and the finished image leaves with ONE bulk async store (TMA) per tile:
//           full 32-byte sectors only.  Scattered 8-byte global stores made L2 fetch every partially written
//           sector from DRAM (measured: profiles/round1_ncu_tiles_v1_c3_n110.txt and the symmetric-store experiment).
//   The residual is a separate per-element kernel (r_e = K_e u_e through the same reference tensors) followed
//   by the fixed-order per-node gather of scatter.cu.
//
// Every pair sums its contributions in ascending element order; results are bitwise reproducible.
#include <cub/cub.cuh>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>

#include "common.cuh"
#include "tile_common.cuh"

namespace gf {

// ---------------------------------------------------------------- per-element geometry
template <int N>
__global__ void k_tile_geo(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
                           const int32_t *__restrict__ conn, int ng, const double *__restrict__ pc /* ng x N */,
                           int64_t e0, int64_t ne, int rf, double scale, double *__restrict__ eg, int gsz) {
  for (int64_t el = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; el < ne; el += (int64_t)gridDim.x * blockDim.x) {
    const int32_t *cv = conn + (e0 + el) * ng;
    double K[N * N];
#pragma unroll
    for (int k = 0; k < N * N; ++k) K[k] = 0.0;
    for (int i = 0; i < ng; ++i) {
      const int32_t p = cv[i];
      double g[3] = {x[p], y[p], N == 3 ? z[p] : 0.0};
#pragma unroll
      for (int c = 0; c < N; ++c)
#pragma unroll
        for (int r = 0; r < N; ++r) K[r + N * c] += g[r] * pc[i * N + c];
    }
    double B[N * N], J;  // B = K^-T, B(n,p) at n + N*p
    if (N == 2) {
      double d = K[0] * K[3] - K[1] * K[2], id = 1.0 / d;
      B[0] = K[3] * id; B[2] = -K[1] * id; B[1] = -K[2] * id; B[3] = K[0] * id;
      J = fabs(d);
    } else {
#define K_(i, j) K[(i) + 3 * (j)]
      double c00 = K_(1, 1) * K_(2, 2) - K_(1, 2) * K_(2, 1);
      double c10 = K_(1, 2) * K_(2, 0) - K_(1, 0) * K_(2, 2);
      double c20 = K_(1, 0) * K_(2, 1) - K_(1, 1) * K_(2, 0);
      double d = K_(0, 0) * c00 + K_(0, 1) * c10 + K_(0, 2) * c20, id = 1.0 / d;
      B[0] = c00 * id; B[3] = c10 * id; B[6] = c20 * id;
      B[1] = (K_(0, 2) * K_(2, 1) - K_(0, 1) * K_(2, 2)) * id;
      B[4] = (K_(0, 0) * K_(2, 2) - K_(0, 2) * K_(2, 0)) * id;
      B[7] = (K_(0, 1) * K_(2, 0) - K_(0, 0) * K_(2, 1)) * id;
      B[2] = (K_(0, 1) * K_(1, 2) - K_(0, 2) * K_(1, 1)) * id;
      B[5] = (K_(0, 2) * K_(1, 0) - K_(0, 0) * K_(1, 2)) * id;
      B[8] = (K_(0, 0) * K_(1, 1) - K_(0, 1) * K_(1, 0)) * id;
#undef K_
      J = fabs(d);
    }
    double *o = eg + (size_t)el * gsz;
    if (rf == TF_ELAST) {  // sqrt(|alpha| J) B : T~ = B~ M B~^T = |alpha| J B M B^T, the sign of alpha is applied at the flush
      const double s = sqrt(fabs(scale) * J);
#pragma unroll
      for (int k = 0; k < N * N; ++k) o[k] = s * B[k];
    } else if (rf == TF_LAPLACE) {  // scale*J * B^T B, upper triangle row by row
      int k = 0;
#pragma unroll
      for (int p = 0; p < N; ++p)
#pragma unroll
        for (int q = p; q < N; ++q) {
          double s = 0;
#pragma unroll
          for (int n = 0; n < N; ++n) s += B[n + N * p] * B[n + N * q];
          o[k++] = scale * J * s;
        }
    } else {
      o[0] = scale * J;
    }
  }
}

// ---------------------------------------------------------------- plan
struct alignas(16) TileHdr {
  int64_t base;              // CSC position of the tile's first entry
  uint32_t pair0, npairs;    // pair range (Structure order = CSC order)
  uint32_t node0, nnodes;    // column-node range
  uint32_t task0, ntasks;    // task range
  uint32_t nel, n_long;      // distinct elements staged; tasks that need a descriptor blob (they come first)
  uint32_t n_wide, r2_0;     // pairs with more than TL_INREC contributions (one task each); first blob (task units)
  uint32_t len;              // entries of the tile's CSC segment [base, base + len)
  uint32_t nitems;           // lane items of the normal tasks (an item = up to TL_KG pairs sharing column node + element list)
  uint32_t pad2, pad3;
};
static_assert(sizeof(TileHdr) == 64, "TileHdr layout");

__global__ void k_tile_base(TileHdr *__restrict__ hdr, int64_t nt, const int32_t *__restrict__ pJ,
                            const int32_t *__restrict__ rdof, int Q, const int64_t *__restrict__ jc) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nt; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = jc[pJ[hdr[k].pair0]];
    hdr[k].base = b;
    hdr[k].len = (uint32_t)(jc[rdof[hdr[k].node0 + hdr[k].nnodes - 1] + Q] - b);
  }
}

// CTA per tile: sorted distinct local element ids of the tile's incidences -> els[tile*stride ..], hdr.nel
__global__ void __launch_bounds__(128)
k_tile_elements(TileHdr *__restrict__ hdr, const uint32_t *__restrict__ rstart, const uint32_t *__restrict__ rsrc, int nd,
                int stride, uint32_t *__restrict__ els) {
  extern __shared__ uint32_t sv[];  // P values + P flags
  const int64_t tile = blockIdx.x;
  const uint32_t n0 = hdr[tile].node0, n1 = n0 + hdr[tile].nnodes;
  const uint32_t r0 = rstart[n0], n = rstart[n1] - r0;
  int P = 32;
  while (P < (int)n) P <<= 1;
  uint32_t *flag = sv + P;
  for (int k = threadIdx.x; k < P; k += blockDim.x) sv[k] = k < (int)n ? rsrc[r0 + k] / (uint32_t)nd : 0xffffffffu;
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1)
    for (int stridek = size >> 1; stridek > 0; stridek >>= 1) {
      for (int k = threadIdx.x; k < P; k += blockDim.x) {
        const int partner = k ^ stridek;
        if (partner > k) {
          const bool up = (k & size) == 0;
          const uint32_t a = sv[k], b = sv[partner];
          if ((a > b) == up) { sv[k] = b; sv[partner] = a; }
        }
      }
      __syncthreads();
    }
  for (int k = threadIdx.x; k < P; k += blockDim.x)
    flag[k] = (sv[k] != 0xffffffffu && (k == 0 || sv[k] != sv[k - 1])) ? 1u : 0u;
  __syncthreads();
  // exclusive scan of the flags: each thread owns a contiguous chunk
  __shared__ uint32_t part[128];
  const int chunk = P / blockDim.x > 0 ? P / blockDim.x : 1;
  const int c0 = threadIdx.x * chunk;
  uint32_t s = 0;
  if (c0 < P)
    for (int k = c0; k < c0 + chunk; ++k) s += flag[k];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int k = 0; k < (int)blockDim.x; ++k) { const uint32_t v = part[k]; part[k] = run; run += v; }
    hdr[tile].nel = run;
  }
  __syncthreads();
  if (c0 < P) {
    uint32_t pos = part[threadIdx.x];
    for (int k = c0; k < c0 + chunk; ++k)
      if (flag[k]) els[(size_t)tile * stride + pos++] = sv[k];
  }
}

// CTA per tile: the geometry rows of the tile's distinct elements, copied once into a contiguous per-tile blob so that
// the tangent kernel stages them with ONE bulk copy (the LDGSTS gather of 8-byte pieces kept the producer warp busy for
// 1.4 us per tile and cost ~700 shared-memory wavefronts: profiles/round1_tile_trace_producer.txt)
__global__ void __launch_bounds__(128)
k_tile_gather_geo(const TileHdr *__restrict__ hdr, const uint32_t *__restrict__ els, int els_stride,
                  const double *__restrict__ eg, int gsz, int gsp /* padded row length, as in shared memory */,
                  int rows_stride, double *__restrict__ tgeo) {
  const int64_t tile = blockIdx.x;
  const uint32_t nel = hdr[tile].nel;
  const uint32_t *te = els + (size_t)tile * els_stride;
  double *o = tgeo + (size_t)tile * rows_stride * gsp;
  for (uint32_t k = threadIdx.x; k < nel * (uint32_t)gsz; k += blockDim.x)
    o[(k / gsz) * gsp + k % gsz] = eg[(size_t)te[k / gsz] * gsz + k % gsz];
}

// Tasks of a tile, longest first:
//   wide tasks   : one per pair with more than TL_INREC contributions (vertex-diagonal pairs): the warp spreads the
//                  contributions over its lanes (lane l takes l, l+32, ...) and tree-reduces the accumulators;
//   normal tasks : 32 pairs each of the remaining pairs sorted by contribution count (descending), lane = pair.
constexpr int TL_INREC = 10;  // largest step count of a normal task = rows of a descriptor blob

// CTA per tile: sort of the tile's pairs by (contribution count descending, signature, pair id) -> sp_pair[pair0 ..],
// number of wide pairs and of tasks that need a descriptor blob.  The signature hashes the pair's sequence of local
// node couples (j, i): pairs that are translated copies of each other (structured parts of a mesh) become neighbours,
// so the lanes of a task mostly read the SAME reference tensor M^{ji} at every step (a shared-memory broadcast instead
// of a 32-way gather).  Any order inside a count class is correct (the output goes through the shared-memory image,
// so the lane order has no effect on the global stores); this one only changes bank conflicts.
constexpr int TL_MAXPAIRS = 4096;
#ifndef GF_TL_KG
#define GF_TL_KG 1
#endif
constexpr int TL_KG = GF_TL_KG;  // pairs per lane item

// element identity of contribution `ctr` (virtual halo contributions are unique: they never make two pairs equal)
__device__ __forceinline__ uint32_t tl_elem_of(uint32_t ctr, uint32_t nb, uint32_t nlocal) {
  return ctr < nlocal ? ctr / nb : (0x80000000u | ctr);
}

// CTA per tile.  Phase 1 sorts the tile's pairs by (contribution count descending, hash of (column node, element list), pair
// id): pairs of ONE column node that are fed by the SAME elements become neighbours and are cut into ITEMS of up to TL_KG
// pairs -- a lane item: the geometry row of a step is loaded once and serves all its pairs.  Phase 2 sorts the items by
// (count descending, members descending, joint signature of their (j, i) sequences, position): items that are translated copies of each other
// (structured parts of a mesh) become neighbours, so the lanes of a task mostly read the SAME reference tensor at a step.
// Outputs: sp_pair (phase-1 order), it_head (per item: members << 16 | position of its first pair in sp_pair), the number
// of wide pairs (more than TL_INREC contributions: one task each, they come first), of items, of long tasks.
__global__ void __launch_bounds__(256)
k_tile_sort_pairs(TileHdr *__restrict__ hdr, const uint32_t *__restrict__ cstart, const uint32_t *__restrict__ csrc,
                  const int32_t *__restrict__ pJ, uint32_t nb, uint32_t nlocal, int use_sig, uint32_t *__restrict__ sp_pair,
                  uint32_t *__restrict__ it_head, int *__restrict__ err) {
  __shared__ uint64_t key[TL_MAXPAIRS];
  __shared__ uint8_t same[TL_MAXPAIRS];
  __shared__ uint32_t s_nwide, s_nitems, s_nit_gt2;
  const int64_t tile = blockIdx.x;
  const TileHdr h = hdr[tile];
  const uint32_t n = h.npairs;
  if (n > (uint32_t)TL_MAXPAIRS) {
    if (threadIdx.x == 0) *err = 8;
    return;
  }
  uint32_t P = 32;
  while (P < n) P <<= 1;
  if (threadIdx.x == 0) { s_nwide = 0; s_nitems = 0; s_nit_gt2 = 0; }
  __syncthreads();
  auto bitonic = [&]() {
    for (uint32_t size = 2; size <= P; size <<= 1)
      for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
        for (uint32_t k = threadIdx.x; k < P; k += blockDim.x) {
          const uint32_t partner = k ^ stride;
          if (partner > k) {
            const bool up = (k & size) == 0;
            const uint64_t a = key[k], b = key[partner];
            if ((a > b) == up) { key[k] = b; key[partner] = a; }
          }
        }
        __syncthreads();
      }
  };
  // ---- phase 1
  for (uint32_t k = threadIdx.x; k < P; k += blockDim.x) {
    uint64_t v = ~0ull;
    if (k < n) {
      const uint32_t p = h.pair0 + k, s0 = cstart[p], cnt = cstart[p + 1] - s0;
      if (cnt > 255u) *err = 3;
      uint32_t hh = (uint32_t)pJ[p] * 0x85EBCA6Bu;
      if (cnt <= (uint32_t)TL_INREC)
        for (uint32_t c = 0; c < cnt; ++c) {
          hh = (hh ^ tl_elem_of(csrc[s0 + c], nb, nlocal)) * 0x9E3779B1u + 0x7F4A7C15u;
          hh ^= hh >> 15;
        }
      if (cnt > (uint32_t)TL_INREC) atomicAdd(&s_nwide, 1u);
      v = ((uint64_t)(255u - min(cnt, 255u)) << 56) | ((uint64_t)(hh & 0xffffffu) << 32) | k;
    }
    key[k] = v;
  }
  __syncthreads();
  bitonic();
  const uint32_t nwide = s_nwide;
  // same[pos]: the pair at pos continues the run of its predecessor (same column node, same element list)
  for (uint32_t pos = threadIdx.x; pos < n; pos += blockDim.x) {
    bool sm = false;
    if (TL_KG > 1 && pos > nwide) {
      const uint32_t p = h.pair0 + (uint32_t)(key[pos] & 0xffffffffu), q = h.pair0 + (uint32_t)(key[pos - 1] & 0xffffffffu);
      const uint32_t sp = cstart[p], sq = cstart[q], cnt = cstart[p + 1] - sp;
      sm = cnt == cstart[q + 1] - sq && pJ[p] == pJ[q];
      for (uint32_t c = 0; sm && c < cnt; ++c)
        sm = tl_elem_of(csrc[sp + c], nb, nlocal) == tl_elem_of(csrc[sq + c], nb, nlocal);
    }
    same[pos] = sm ? 1 : 0;
    sp_pair[h.pair0 + pos] = h.pair0 + (uint32_t)(key[pos] & 0xffffffffu);
  }
  __syncthreads();
  // ---- phase 2: item heads
  uint64_t mykey[TL_MAXPAIRS / 256];
#pragma unroll
  for (int r = 0; r < TL_MAXPAIRS / 256; ++r) {
    const uint32_t pos = threadIdx.x + r * 256;
    uint64_t v = ~0ull;
    if (pos < n && pos >= nwide) {
      uint32_t back = 0;
      while (same[pos - back]) ++back;  // same[nwide] == 0 stops the walk
      if (back % TL_KG == 0) {
        uint32_t m = 1;
        while (m < (uint32_t)TL_KG && pos + m < n && same[pos + m]) ++m;
        const uint32_t p0 = h.pair0 + (uint32_t)(key[pos] & 0xffffffffu);
        const uint32_t cnt = cstart[p0 + 1] - cstart[p0];
        uint32_t sig = m;
        if (use_sig)
          for (uint32_t g = 0; g < m; ++g) {
            const uint32_t p = h.pair0 + (uint32_t)(key[pos + g] & 0xffffffffu), s0 = cstart[p];
            for (uint32_t c = 0; c < cnt; ++c) {
              const uint32_t ctr = csrc[s0 + c];
              const uint32_t rr = ctr < nlocal ? ctr % nb : 0xffffu;
              sig = (sig ^ rr) * 0x9E3779B1u + 0x7F4A7C15u;
              sig ^= sig >> 15;
            }
          }
        atomicAdd(&s_nitems, 1u);
        if (cnt > 2u) atomicAdd(&s_nit_gt2, 1u);
        // items with the same number of members share tasks: a task then runs exactly the members its lanes have
        v = ((uint64_t)(255u - min(cnt, 255u)) << 56) | ((uint64_t)((uint32_t)TL_KG - m) << 52) |
            ((uint64_t)(sig & 0xfffffu) << 32) | (m << 16) | pos;
      }
    }
    mykey[r] = v;
  }
  __syncthreads();  // every thread has read the phase-1 keys it needs
#pragma unroll
  for (int r = 0; r < TL_MAXPAIRS / 256; ++r) {
    const uint32_t pos = threadIdx.x + r * 256;
    if (pos < P) key[pos] = mykey[r];
  }
  __syncthreads();
  bitonic();
  const uint32_t nitems = s_nitems;
  for (uint32_t k = threadIdx.x; k < nitems; k += blockDim.x) it_head[h.pair0 + k] = (uint32_t)(key[k] & 0xffffffffu);
  if (threadIdx.x == 0) {
    hdr[tile].n_wide = nwide;
    hdr[tile].nitems = nitems;
    hdr[tile].n_long = nwide + (s_nit_gt2 + 31) / 32;
  }
}

// thread per task: owning tile
__global__ void k_tile_task_owner(const TileHdr *__restrict__ hdr, int64_t nt, uint32_t *__restrict__ tk_tile) {
  for (int64_t tile = blockIdx.x; tile < nt; tile += gridDim.x)
    for (uint32_t k = threadIdx.x; k < hdr[tile].ntasks; k += blockDim.x) tk_tile[hdr[tile].task0 + k] = (uint32_t)tile;
}

// thread per (task, lane): the TL_KG pair records of the lane's item + descriptors.
//   rec[(task*TL_KG + g)*32 + lane], g = member of the item:
//   rec.x  = CSC offset of the column-component-0 piece (relative to the tile base, 20 bits)
//   rec.y  = offset of the component-1 piece relative to x (20 bits) | keep mask << 20 | wide << 30 | valid << 31
//   rec.z  = offset of the component-2 piece relative to x (20 bits) | steps of the task << 20
//   rec.w  = descriptors of steps 0 and 1 (16 bits each): element slot << cbits | (j*nd + i)
//   dblob  = tasks with more than 2 steps and wide tasks: descriptors as [member][step][lane] uint16, TL_INREC rows per
//            member.  Wide task: lane l, row r = contribution r*32 + l of the pair; only lane 0 / member 0 is valid.
//   The members of an item share their element list: at every step their descriptors carry the same slot.
template <int Q>
__global__ void k_tile_fill(const TileHdr *__restrict__ hdr, const uint32_t *__restrict__ tk_tile,
                            const uint32_t *__restrict__ sp_pair, const uint32_t *__restrict__ it_head,
                            const uint32_t *__restrict__ els, int stride,
                            const uint32_t *__restrict__ cstart, const uint32_t *__restrict__ csrc,
                            const int32_t *__restrict__ pJ, const uint16_t *__restrict__ pmask,
                            const uint32_t *__restrict__ prel, const int64_t *__restrict__ jc, int64_t npairs, int nd,
                            int cbits, uint32_t zslot, int64_t ntask, uint32_t nlocal, uint4 *__restrict__ rec,
                            uint16_t *__restrict__ dblob, int *__restrict__ err) {
  const uint32_t nb = nd * nd;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < ntask * 32;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t task = idx >> 5;
    const uint32_t lane = (uint32_t)(idx & 31);
    const uint32_t tile = tk_tile[task];
    const TileHdr h = hdr[tile];
    const uint32_t tkl = (uint32_t)(task - h.task0);
    const bool wide = tkl < h.n_wide;
    // position (in sp_pair) of the first pair of my item and of the task's first (longest) item
    uint32_t pos = tkl, members = 1, posf = tkl;
    bool exists = true;
    if (!wide) {
      const uint32_t t0 = (tkl - h.n_wide) * 32, tl = t0 + lane;
      posf = it_head[h.pair0 + t0] & 0xffffu;
      exists = tl < h.nitems;
      if (exists) {
        const uint32_t ih = it_head[h.pair0 + tl];
        pos = ih & 0xffffu;
        members = (ih >> 16) & 0xffu;
      }
    }
    const uint32_t pf = sp_pair[h.pair0 + posf];
    const uint32_t cntf = cstart[pf + 1] - cstart[pf];
    const uint32_t steps = wide ? (cntf + 31) / 32 : cntf;
    if (steps > (uint32_t)TL_INREC) *err = 4;
    if (members > (uint32_t)TL_KG) *err = 6;
    const uint32_t zdesc = (zslot << cbits) & 0xffffu;
    const uint32_t *te = els + (size_t)tile * stride;
    for (int g = 0; g < TL_KG; ++g) {
      uint32_t w0 = zdesc | (zdesc << 16);
      uint16_t *bl = tkl < h.n_long ? dblob + (((size_t)h.r2_0 + tkl) * TL_KG + g) * (TL_INREC * 32) + lane : nullptr;
      if (!bl && steps > 2) *err = 5;
      uint4 r = make_uint4(0, wide ? 0x40000000u : 0u, steps << 20, 0);
      uint32_t cnt = 0, s0 = 0;
      if (exists && (uint32_t)g < members && pos + g < h.npairs) {
        const uint32_t p = sp_pair[h.pair0 + pos + g];
        const int32_t J = pJ[p];
        int64_t off[3] = {0, 0, 0};
        for (int b = 0; b < Q; ++b) off[b] = jc[J + b] + prel[(size_t)b * npairs + p] - h.base;
        const int64_t d1 = off[1] - off[0], d2 = off[2] - off[0];
        const uint32_t m = pmask[p];
        if (off[0] < 0 || off[0] >= (1 << 20) || (Q > 1 && (d1 < 0 || d1 >= (1 << 20))) ||
            (Q > 2 && (d2 < 0 || d2 >= (1 << 20))) || m >= (1u << 9))
          *err = 1;
        r.x = (uint32_t)off[0];
        r.y |= (uint32_t)(Q > 1 ? d1 : 0) | (m << 20) | ((!wide || lane == 0) ? 0x80000000u : 0u);
        r.z |= (uint32_t)(Q > 2 ? d2 : 0);
        s0 = cstart[p];
        cnt = cstart[p + 1] - s0;
      }
      for (uint32_t c = 0; c < (bl ? (uint32_t)TL_INREC : steps); ++c) {
        const uint32_t ci = wide ? c * 32 + lane : c;  // contribution handled at step c
        uint32_t d = zdesc;
        if (ci < cnt && c < steps && csrc[s0 + ci] < nlocal) {  // virtual (halo) contributions add nothing here
          const uint32_t ctr = csrc[s0 + ci];
          const uint32_t el = ctr / nb, rr = ctr - el * nb;  // rr = j*nd + i
          uint32_t lo = 0, hi = h.nel;                       // first position with te[pos] >= el
          while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (te[mid] < el) lo = mid + 1; else hi = mid;
          }
          if (lo >= h.nel || te[lo] != el) *err = 2;
          d = (lo << cbits) | rr;
        }
        if (c < 2) w0 = (w0 & ~(0xffffu << (16 * c))) | (d << (16 * c));
        if (bl) bl[c * 32] = (uint16_t)d;
      }
      r.w = w0;
      rec[((size_t)task * TL_KG + g) * 32 + lane] = r;
    }
  }
}

// ---------------------------------------------------------------- the tangent kernel
// Warp-specialised CTA, one per SM: TL_CW consumer warps + one producer warp, two tile buffers in shared memory.
//   producer : per tile, one bulk async copy (TMA, mbarrier complete_tx) of the tile's pair records, one of its
//              long-task descriptors, and an LDGSTS gather of the element geometry rows, all arriving on full[b];
//   consumer : waits full[b], takes tasks (first one static, then a shared-memory counter, longest first), lane =
//              pair, operands from shared memory only, kept entries stored straight to their CSC slots, then
//              arrives on empty[b].  No __syncthreads in the steady state.
struct TileArgs {
  const TileHdr *hdr;
  const double *tgeo;   // per tile: geo_rows rows of GSP doubles (the tile's distinct elements, slot order; zero padded)
  int geo_rows;
  const uint4 *rec;
  const uint16_t *dblob;
  const double *Mtab;
  double sl, smu;  // sign(alpha) * lambda, sign(alpha) * mu (elasticity)
  int64_t nt;
  int zslot, cap_tasks, cap_long, cap_len;
  double *pr;
  unsigned long long *trace;  // optional (GFGPU_TILE_TRACE): per-tile timestamps of CTA 0
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v));
  return v;
}

#ifndef GF_TL_CW
#define GF_TL_CW 11
#endif
constexpr int TL_CW = GF_TL_CW;                  // consumer warps
constexpr int TL_THREADS = (TL_CW + 1) * 32;     // + the producer warp
constexpr int TL_BLOB = TL_INREC * 32 * 2;       // bytes of a long task's descriptor blob

template <int N, int RF>
struct TlSmem {  // byte offsets inside one tile buffer: pair records | long-task descriptor blobs | geometry rows
  using C = TlCfg<N, RF>;
  static constexpr int GSP = C::GSZ | 1;
  __host__ __device__ static size_t blob_off(int cap_tasks) { return (size_t)cap_tasks * 512 * TL_KG; }
  __host__ __device__ static size_t geo_off(int cap_tasks, int cap_long) {
    return (size_t)cap_tasks * 512 * TL_KG + (size_t)cap_long * TL_BLOB * TL_KG;
  }
  __host__ __device__ static size_t bytes(int cap_tasks, int cap_long, int zslot) {
    return (geo_off(cap_tasks, cap_long) + (size_t)(zslot + 1) * GSP * 8 + 127) / 128 * 128;
  }
  // image of the tile's CSC segment: entry e lives at index (base & 1) + e so that shared and global addresses agree
  // modulo 16 (bulk copies need 16-byte alignment on both sides)
  __host__ __device__ static size_t out_bytes(int cap_len) { return ((size_t)(cap_len + 2) * 8 + 127) / 128 * 128; }
};

template <int N, int Q, int ND, int RF>
__global__ void __launch_bounds__(TL_THREADS, 1)
k_tiles(const TileArgs a) {
  using C = TlCfg<N, RF>;
  using L = TlSmem<N, RF>;
  constexpr int NB = ND * ND, MT = C::MT, GSZ = C::GSZ, GSP = GSZ | 1, ACC = C::ACC, CB = tl_clog2(NB);
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ uint64_t full[2], empty[2];
  __shared__ TileHdr s_hdr[2];
  __shared__ unsigned s_next[2];
  const size_t bufsz = L::bytes(a.cap_tasks, a.cap_long, a.zslot);
  const size_t outsz = L::out_bytes(a.cap_len);
  unsigned char *outraw = smraw + 2 * bufsz;
  double *sM = reinterpret_cast<double *>(outraw + 2 * outsz);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k < NB * MT; k += TL_THREADS) sM[k] = a.Mtab[k];
  if (tid < 2 * GSP) {  // the all-zero geometry slot of both buffers
    double *g = reinterpret_cast<double *>(smraw + (tid / GSP) * bufsz + L::geo_off(a.cap_tasks, a.cap_long));
    g[a.zslot * GSP + tid % GSP] = 0.0;
  }
  if (tid == 0) {
    mbar_init(&full[0], 2); mbar_init(&full[1], 2);        // expect_tx (three bulk copies) + "output image free"
    mbar_init(&empty[0], TL_CW); mbar_init(&empty[1], TL_CW);
  }
  __syncthreads();
  if (warp == TL_CW) {
    // ================= producer =================
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    // the header and the element ids of the NEXT tile are loaded (into registers) before waiting for its buffer
    TileHdr h;
    auto fetch = [&](int64_t tile) { h = a.hdr[tile]; };
    if ((int64_t)blockIdx.x < a.nt) fetch(blockIdx.x);
    int4 hprev[2] = {make_int4(0, 0, 0, 0), make_int4(0, 0, 0, 0)};  // base (lo, hi), len of the tile in each buffer
    auto drain = [&](int b, int4 hp) {
      const int64_t base = (int64_t)(uint32_t)hp.x | ((int64_t)hp.y << 32);
      const int64_t len = (uint32_t)hp.z;
      if (lane == 0 && len) {
        fence_async_smem();
        const double *img = reinterpret_cast<const double *>(outraw + b * outsz);
        const int64_t odd = base & 1, gs = base + odd, ge = (base + len) & ~int64_t(1);
        if (odd) a.pr[base] = img[1];
        if (ge > gs) bulk_s2g(a.pr + gs, img + 2 * odd, (uint32_t)((ge - gs) * 8));
        if (((base + len) & 1) && base + len - 1 >= gs) a.pr[base + len - 1] = img[odd + len - 1];
        bulk_commit();
        bulk_wait_read0();
      }
      __syncwarp();
    };
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < a.nt; tile += gridDim.x, ++it) {
      const int b = it & 1;
      const unsigned long long tp0 = a.trace ? gtimer() : 0;
      if (it >= 2) mbar_wait(&empty[b], ((it >> 1) - 1) & 1);
      const unsigned long long tp1 = a.trace ? gtimer() : 0;
      unsigned char *buf = smraw + b * bufsz;
      if (lane == 0) {
        s_hdr[b] = h;
        s_next[b] = 2 * TL_CW;
        const uint32_t nb1 = h.ntasks * 512u * TL_KG, nb2 = h.n_long * (uint32_t)(TL_BLOB * TL_KG);
        // geometry rows: a whole number of 16-byte units (the blob is zero padded, a spare row may land on the zero slot)
        const uint32_t nb3 = (h.nel * (uint32_t)(GSP * 8) + 15u) & ~15u;
        mbar_arrive_expect_tx(&full[b], nb1 + nb2 + nb3);
        bulk_g2s(buf, a.rec + (size_t)h.task0 * 32 * TL_KG, nb1, &full[b], pol);
        if (nb2) bulk_g2s(buf + L::blob_off(a.cap_tasks), a.dblob + (size_t)h.r2_0 * (TL_BLOB / 2) * TL_KG, nb2, &full[b], pol);
        if (nb3) bulk_g2s(buf + L::geo_off(a.cap_tasks, a.cap_long), a.tgeo + (size_t)tile * a.geo_rows * GSP, nb3, &full[b], pol);
      }
      const unsigned long long tpa = a.trace ? gtimer() : 0;
      const unsigned long long tpb = tpa;
      // the image of the tile that used this buffer two iterations ago is complete (empty[b] fired): send it
      if (it >= 2) drain(b, hprev[b]);
      if (lane == 0) mbar_arrive(&full[b]);  // output image b is free again
      hprev[b] = make_int4((int)(h.base & 0xffffffffll), (int)(h.base >> 32), (int)h.len, 0);
      const unsigned long long tp2 = a.trace ? gtimer() : 0;
      if (tile + gridDim.x < a.nt) fetch(tile + gridDim.x);
      if (a.trace && blockIdx.x == 0 && lane == 0 && it < 256) {
        a.trace[it * 12 + 0] = tp0; a.trace[it * 12 + 1] = tp1; a.trace[it * 12 + 2] = tp2; a.trace[it * 12 + 3] = gtimer() + (h.nel & 0);
        a.trace[it * 12 + 8] = tpa; a.trace[it * 12 + 9] = tpb; a.trace[it * 12 + 10] = h.nel; a.trace[it * 12 + 11] = h.len;
      }
    }
    // the last two tiles of this CTA: `it` is now the number of tiles it processed
    for (int k = it >= 2 ? it - 2 : 0; k < it; ++k) {
      const int b = k & 1;
      mbar_wait(&empty[b], (k >> 1) & 1);
      drain(b, hprev[b]);
    }
    if (lane == 0) bulk_wait0();
    return;
  }
  // ================= consumers =================
  int it = 0;
  for (int64_t tile = blockIdx.x; tile < a.nt; tile += gridDim.x, ++it) {
    const int b = it & 1;
    const unsigned long long tc0 = a.trace ? gtimer() : 0;
    mbar_wait(&full[b], (it >> 1) & 1);
    const unsigned long long tc1 = a.trace ? gtimer() : 0;
    const unsigned char *buf = smraw + b * bufsz;
    const uint4 *sRec = reinterpret_cast<const uint4 *>(buf) + lane;
    const uint16_t *sBlob = reinterpret_cast<const uint16_t *>(buf + L::blob_off(a.cap_tasks)) + lane;
    const double *sG = reinterpret_cast<const double *>(buf + L::geo_off(a.cap_tasks, a.cap_long));
    const unsigned ntasks = s_hdr[b].ntasks;
    double *prb = reinterpret_cast<double *>(outraw + b * outsz) + (s_hdr[b].base & 1);  // image of the CSC segment
    // the next task index and its first record are fetched while the current task runs
    unsigned tk = warp, tkB = warp + TL_CW;
    uint4 rec0 = make_uint4(0, 0, 0, 0);
    if (tk < ntasks) rec0 = sRec[tk * (32 * TL_KG)];
    while (tk < ntasks) {
      unsigned nxt = 0;
      if (lane == 0) nxt = atomicAdd(&s_next[b], 1u);
      uint4 recB = make_uint4(0, 0, 0, 0);
      if (tkB < ntasks) recB = sRec[tkB * (32 * TL_KG)];
      uint4 rec[TL_KG];
      rec[0] = rec0;
#pragma unroll
      for (int g = 1; g < TL_KG; ++g) rec[g] = sRec[(tk * TL_KG + g) * 32];
      // members in use by some lane of this task (warp uniform): the members of an item are contiguous from 0
      int kmax = 1;
#pragma unroll
      for (int g = 1; g < TL_KG; ++g)
        if (__any_sync(0xffffffffu, rec[g].y >> 31)) kmax = g + 1;
      const int steps = (int)(rec0.z >> 20);
      double acc[TL_KG][ACC];
#pragma unroll
      for (int g = 0; g < TL_KG; ++g)
#pragma unroll
        for (int m = 0; m < ACC; ++m) acc[g][m] = 0.0;
      // one step: the geometry row (slot of member 0's descriptor: the members share their element list) is loaded once
      auto step = [&](const unsigned (&d)[TL_KG]) {
        const double *G = sG + (d[0] >> CB) * GSP;
        if (RF == TF_ELAST) {
          double Bm[N * N];
#pragma unroll
          for (int k = 0; k < N * N; ++k) Bm[k] = G[k];
#pragma unroll
          for (int g = 0; g < TL_KG; ++g) {
            if (g < kmax) {
              const double *M = sM + (d[g] & ((1u << CB) - 1u)) * MT;
#pragma unroll
              for (int qq = 0; qq < N; ++qq) {
                double Wq[N];  // column qq of W = B M
#pragma unroll
                for (int aa = 0; aa < N; ++aa) {
                  double s2 = 0;
#pragma unroll
                  for (int pp = 0; pp < N; ++pp) s2 += Bm[aa + N * pp] * M[pp * N + qq];
                  Wq[aa] = s2;
                }
#pragma unroll
                for (int b2 = 0; b2 < N; ++b2)
#pragma unroll
                  for (int aa = 0; aa < N; ++aa) acc[g][RF == TF_ELAST ? aa + N * b2 : 0] += Wq[aa] * Bm[b2 + N * qq];
              }
            }
          }
        } else {
          double Gv[MT];
#pragma unroll
          for (int k = 0; k < MT; ++k) Gv[k] = G[k];
#pragma unroll
          for (int g = 0; g < TL_KG; ++g) {
            if (g < kmax) {
              const double *M = sM + (d[g] & ((1u << CB) - 1u)) * MT;
              double s2 = acc[g][0];
#pragma unroll
              for (int k = 0; k < MT; ++k) s2 += M[k] * Gv[k];
              acc[g][0] = s2;
            }
          }
        }
      };
      const bool wide = rec0.y & 0x40000000u;
      if (steps <= 2 && !wide) {  // both descriptors travel in the records
        unsigned d[TL_KG];
#pragma unroll
        for (int g = 0; g < TL_KG; ++g) d[g] = rec[g].w & 0xffffu;
        step(d);
        if (steps == 2) {
#pragma unroll
          for (int g = 0; g < TL_KG; ++g) d[g] = rec[g].w >> 16;
          step(d);
        }
      } else {
        const uint16_t *bl = sBlob + tk * (TL_KG * (TL_BLOB / 2));
        unsigned dn[TL_KG];
#pragma unroll
        for (int g = 0; g < TL_KG; ++g) dn[g] = g < kmax ? bl[g * (TL_BLOB / 2)] : 0u;
        for (int c = 0; c < steps; ++c) {
          unsigned d[TL_KG];
#pragma unroll
          for (int g = 0; g < TL_KG; ++g) d[g] = dn[g];
          if (c + 1 < steps) {
#pragma unroll
            for (int g = 0; g < TL_KG; ++g)
              if (g < kmax) dn[g] = bl[g * (TL_BLOB / 2) + (c + 1) * 32];
          }
          step(d);
        }
        if (wide) {  // wide task: the lanes hold parts of ONE pair; fixed-order tree sum
#pragma unroll
          for (int off = 16; off > 0; off >>= 1)
#pragma unroll
            for (int m = 0; m < ACC; ++m) acc[0][m] += __shfl_xor_sync(0xffffffffu, acc[0][m], off);
        }
      }
      // ---- flush: all lanes of the task together; per column component the kept entries are compacted
#pragma unroll
      for (int g = 0; g < TL_KG; ++g) {
        if (g < kmax && (rec[g].y >> 31)) {
          double kv[Q * Q];  // kv[b*Q + aa] = K(row component aa, column component b)
          if (RF == TF_ELAST) {
            double tr = 0;
#pragma unroll
            for (int n = 0; n < N; ++n) tr += acc[g][RF == TF_ELAST ? n + N * n : 0];
#pragma unroll
            for (int b2 = 0; b2 < Q; ++b2)
#pragma unroll
              for (int aa = 0; aa < Q; ++aa)
                kv[b2 * Q + aa] = a.sl * acc[g][RF == TF_ELAST ? aa + N * b2 : 0] + a.smu * acc[g][RF == TF_ELAST ? b2 + N * aa : 0] +
                                  (aa == b2 ? a.smu * tr : 0.0);
          } else {
#pragma unroll
            for (int b2 = 0; b2 < Q; ++b2)
#pragma unroll
              for (int aa = 0; aa < Q; ++aa) kv[b2 * Q + aa] = aa == b2 ? acc[g][0] : 0.0;
          }
          double *p0 = prb + (rec[g].x & 0xfffffu);
          const unsigned pm = (rec[g].y >> 20) & 0x1ffu;
#pragma unroll
          for (int b2 = 0; b2 < Q; ++b2) {
            double *dst = p0 + (b2 == 0 ? 0u : b2 == 1 ? (rec[g].y & 0xfffffu) : (rec[g].z & 0xfffffu));
            const unsigned mb = (pm >> (b2 * Q)) & ((1u << Q) - 1);
            if (Q == 3) {
              const double v0 = kv[b2 * Q], v1 = kv[b2 * Q + (Q > 1 ? 1 : 0)], v2 = kv[b2 * Q + (Q > 2 ? 2 : 0)];
              const int n = __popc(mb);
              const double x0 = (mb & 1u) ? v0 : ((mb & 2u) ? v1 : v2);
              const double x1 = ((mb & 3u) == 3u) ? v1 : v2;
              if (n >= 1) dst[0] = x0;
              if (n >= 2) dst[1] = x1;
              if (n >= 3) dst[2] = v2;
            } else {
#pragma unroll
              for (int aa = 0; aa < Q; ++aa)
                if (mb & (1u << aa)) dst[__popc(mb & ((1u << aa) - 1))] = kv[b2 * Q + aa];
            }
          }
        }
      }
      tk = tkB;
      rec0 = recB;
      tkB = __shfl_sync(0xffffffffu, nxt, 0);
    }
    fence_async_smem();  // my image writes (generic proxy) before the producer's bulk store (async proxy)
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[b]);
    if (a.trace && blockIdx.x == 0 && tid == 0 && it < 256) {
      a.trace[it * 12 + 4] = tc0; a.trace[it * 12 + 5] = tc1; a.trace[it * 12 + 6] = gtimer(); a.trace[it * 12 + 7] = ntasks;
    }
  }
}

// ---------------------------------------------------------------- the residual kernel
// Thread per element: r_e = K_e u_e without forming K_e.  The reference tensors M_ij(p,q) = sum_k w_k ghat_k(i,p)
// ghat_k(j,q) form a Gram matrix of low rank (the gradients of a degree-k element span P_{k-1}: rank 4 for P2
// whatever the number of quadrature points), factorised once on the host, M = sum_k s_k l_k l_k^T.  Per generalised
// point k (uniform shared-memory reads, every thread of the warp at the same entry):
//   elasticity  Uh = sum_j u_j (x) l_k(j),  grad = Uh B~^T,  sigma = sl tr(grad) I + smu (grad + grad^T),
//               r_i += s_k (sigma B~) l_k(i)
//   Laplace     r_i += s_k (Uh G) l_k(i)          mass   r_i += s_k G l_k(i) sum_j l_k(j) u_j
// into the per-element stage; scatter.cu::gather_residual sums them per node in ascending element order
// (replaces ga_instruction_vector_assembly_mf, C&E.cc:4669-4735).
struct ResArgs {
  const int32_t *edof;
  const double *eg, *Ltab, *U;  // Ltab: rank x (sign, then nd x N (or nd) entries)
  double sl, smu;
  int64_t e0, ne;
  int rank;
  double *rstage;
};

template <int N, int Q, int ND, int RF>
__global__ void __launch_bounds__(128, 1)
k_affine_residual(const ResArgs a) {
  using C = TlCfg<N, RF>;
  constexpr int GSZ = C::GSZ, LW = RF == TF_MASS ? ND : ND * N, LS = LW + 1;
  extern __shared__ double sL[];
  for (int k = threadIdx.x; k < a.rank * LS; k += blockDim.x) sL[k] = a.Ltab[k];
  __syncthreads();
  for (int64_t el = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; el < a.ne; el += (int64_t)gridDim.x * blockDim.x) {
    double G[GSZ];
#pragma unroll
    for (int k = 0; k < GSZ; ++k) G[k] = a.eg[(size_t)el * GSZ + k];
    const int32_t *ed = a.edof + (a.e0 + el) * ND;
    double u[ND][Q], r[ND][Q];
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      const int32_t dj = ed[j];
#pragma unroll
      for (int b = 0; b < Q; ++b) {
        u[j][b] = a.U ? a.U[dj + b] : 0.0;
        r[j][b] = 0.0;
      }
    }
#pragma unroll 1
    for (int k = 0; k < a.rank; ++k) {
      const double *L = sL + k * LS;
      const double sk = L[0];
      ++L;
      if (RF == TF_MASS) {
        double c[Q];
#pragma unroll
        for (int b = 0; b < Q; ++b) c[b] = 0.0;
#pragma unroll
        for (int j = 0; j < ND; ++j)
#pragma unroll
          for (int b = 0; b < Q; ++b) c[b] += L[j] * u[j][b];
        const double g = sk * G[0];
#pragma unroll
        for (int i = 0; i < ND; ++i)
#pragma unroll
          for (int b = 0; b < Q; ++b) r[i][b] += g * L[i] * c[b];
      } else {
        double Uh[Q][N];  // reference gradient of u at the generalised point
#pragma unroll
        for (int b = 0; b < Q; ++b)
#pragma unroll
          for (int q = 0; q < N; ++q) Uh[b][q] = 0.0;
#pragma unroll
        for (int j = 0; j < ND; ++j)
#pragma unroll
          for (int q = 0; q < N; ++q) {
            const double l = L[RF == TF_MASS ? 0 : j * N + q];
#pragma unroll
            for (int b = 0; b < Q; ++b) Uh[b][q] += u[j][b] * l;
          }
        double P[Q][N];
        if (RF == TF_ELAST) {
          double gr[N][N], tr = 0;  // (scaled) physical gradient, Q == N
#pragma unroll
          for (int aa = 0; aa < N; ++aa)
#pragma unroll
            for (int bb = 0; bb < N; ++bb) {
              double s2 = 0;
#pragma unroll
              for (int q = 0; q < N; ++q) s2 += Uh[aa % Q][q] * G[RF == TF_ELAST ? bb + N * q : 0];
              gr[aa][bb] = s2;
              if (aa == bb) tr += s2;
            }
          double sg[N][N];
#pragma unroll
          for (int aa = 0; aa < N; ++aa)
#pragma unroll
            for (int bb = 0; bb < N; ++bb) sg[aa][bb] = a.smu * (gr[aa][bb] + gr[bb][aa]) + (aa == bb ? a.sl * tr : 0.0);
#pragma unroll
          for (int aa = 0; aa < N; ++aa)
#pragma unroll
            for (int q = 0; q < N; ++q) {
              double s2 = 0;
#pragma unroll
              for (int bb = 0; bb < N; ++bb) s2 += sg[aa][bb] * G[RF == TF_ELAST ? bb + N * q : 0];
              P[aa % Q][q] = sk * s2;
            }
        } else {  // Laplace: G = upper triangle of the symmetric N x N matrix alpha a J B^T B, row by row
          double Gs[N][N];
          {
            int kk = 0;
#pragma unroll
            for (int p = 0; p < N; ++p)
#pragma unroll
              for (int q = p; q < N; ++q) {
                Gs[p][q] = Gs[q][p] = G[RF == TF_LAPLACE ? kk : 0];
                ++kk;
              }
          }
#pragma unroll
          for (int b = 0; b < Q; ++b)
#pragma unroll
            for (int p = 0; p < N; ++p) {
              double s2 = 0;
#pragma unroll
              for (int q = 0; q < N; ++q) s2 += Gs[p][q] * Uh[b][q];
              P[b][p] = sk * s2;
            }
        }
#pragma unroll
        for (int i = 0; i < ND; ++i)
#pragma unroll
          for (int q = 0; q < N; ++q) {
            const double l = L[RF == TF_MASS ? 0 : i * N + q];
#pragma unroll
            for (int b = 0; b < Q; ++b) r[i][b] += P[b][q] * l;
          }
      }
    }
    double *out = a.rstage + (size_t)el * ND * Q;
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
      for (int b = 0; b < Q; ++b) out[i * Q + b] = r[i][b];
  }
}

// Elasticity (Q == N), one LANE PER (element, component): lane a holds u_a(j), r_a(i) (2 x ND doubles instead of 2 x ND x N:
// ~100 registers instead of 238, so that 4x more warps are resident -- the thread-per-element kernel is latency bound at 12 %
// occupancy -- and several blocks fit NEXT to the tile kernel on the side stream).  The three lanes of an element sit in
// one warp (32 / N elements per warp, the last 32 % N lanes idle) and exchange the rows of the scaled gradient with
// shuffles (sigma couples the components); the loads of U and the stores of r become N-lane contiguous pieces.
template <int N, int ND>
__global__ void __launch_bounds__(128, (ND <= 10 ? 3 : 1))
k_affine_residual_split(const ResArgs a) {
  constexpr int LW = ND * N, LS = LW + 1, EPW = 32 / N;  // elements per warp
  extern __shared__ double sL[];
  for (int k = threadIdx.x; k < a.rank * LS; k += blockDim.x) sL[k] = a.Ltab[k];
  __syncthreads();
  const int lane = threadIdx.x & 31, le = lane / N, c = lane % N;  // local element, component
  const bool lane_ok = le < EPW;
  const int base = le * N;  // first lane of my element
  const int64_t wpb = blockDim.x >> 5, nwarps = (int64_t)gridDim.x * wpb;
  const int64_t ngroups = (a.ne + EPW - 1) / EPW;
  for (int64_t g = blockIdx.x * wpb + (threadIdx.x >> 5); g < ngroups; g += nwarps) {
    const int64_t el = g * EPW + le;
    const bool live = lane_ok && el < a.ne;  // idle lanes run along on element 0 of the group's range (shuffles need them converged)
    const int64_t e = live ? el : g * EPW;
    double G[N * N];
#pragma unroll
    for (int k = 0; k < N * N; ++k) G[k] = a.eg[(size_t)e * (N * N) + k];
    const int32_t *ed = a.edof + (a.e0 + e) * ND;
    double u[ND], r[ND];
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      u[j] = a.U ? a.U[ed[j] + c] : 0.0;
      r[j] = 0.0;
    }
#pragma unroll 1
    for (int k = 0; k < a.rank; ++k) {
      const double *L = sL + k * LS;
      const double sk = L[0];
      ++L;
      double Uh[N];  // row c of the reference gradient of u at the generalised point
#pragma unroll
      for (int q = 0; q < N; ++q) Uh[q] = 0.0;
#pragma unroll
      for (int j = 0; j < ND; ++j)
#pragma unroll
        for (int q = 0; q < N; ++q) Uh[q] += u[j] * L[j * N + q];
      double grc[N];  // row c of the (scaled) physical gradient
#pragma unroll
      for (int bb = 0; bb < N; ++bb) {
        double s2 = 0;
#pragma unroll
        for (int q = 0; q < N; ++q) s2 += Uh[q] * G[bb + N * q];
        grc[bb] = s2;
      }
      // column c of the gradient (gr[bb][c]) and the trace come from the other lanes of the element
      double tr = 0, sg[N];
#pragma unroll
      for (int bb = 0; bb < N; ++bb) {
        double col = 0, dg = 0;  // gr[bb][c], gr[bb][bb]
#pragma unroll
        for (int x = 0; x < N; ++x) {
          const double v = __shfl_sync(0xffffffffu, grc[x], base + bb);  // gr[bb][x]
          if (x == c) col = v;
          if (x == bb) dg = v;
        }
        tr += dg;
        sg[bb] = a.smu * (grc[bb] + col);
      }
#pragma unroll
      for (int bb = 0; bb < N; ++bb)
        if (bb == c) sg[bb] += a.sl * tr;
      double P[N];
#pragma unroll
      for (int q = 0; q < N; ++q) {
        double s2 = 0;
#pragma unroll
        for (int bb = 0; bb < N; ++bb) s2 += sg[bb] * G[bb + N * q];
        P[q] = sk * s2;
      }
#pragma unroll
      for (int i = 0; i < ND; ++i)
#pragma unroll
        for (int q = 0; q < N; ++q) r[i] += P[q] * L[i * N + q];
    }
    if (live) {
      double *out = a.rstage + (size_t)el * ND * N + c;
#pragma unroll
      for (int i = 0; i < ND; ++i) out[i * N] = r[i];
    }
  }
}

// symmetric eigen-decomposition (cyclic Jacobi), A (m x m, row-major) -> eigenvalues in d, eigenvectors in the
// COLUMNS of V.  m <= 60, host only.
static void jacobi_eig(std::vector<double> &A, int m, std::vector<double> &d, std::vector<double> &V) {
  V.assign((size_t)m * m, 0.0);
  for (int i = 0; i < m; ++i) V[(size_t)i * m + i] = 1.0;
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0, diag = 0;
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) (i == j ? diag : off) += A[(size_t)i * m + j] * A[(size_t)i * m + j];
    if (off <= 1e-60 || off <= 1e-34 * diag) break;
    for (int p = 0; p < m; ++p)
      for (int q = p + 1; q < m; ++q) {
        const double apq = A[(size_t)p * m + q];
        if (apq == 0.0) continue;
        const double theta = (A[(size_t)q * m + q] - A[(size_t)p * m + p]) / (2.0 * apq);
        const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(tt * tt + 1.0), sn = tt * c;
        for (int k = 0; k < m; ++k) {
          const double akp = A[(size_t)k * m + p], akq = A[(size_t)k * m + q];
          A[(size_t)k * m + p] = c * akp - sn * akq;
          A[(size_t)k * m + q] = sn * akp + c * akq;
        }
        for (int k = 0; k < m; ++k) {
          const double apk = A[(size_t)p * m + k], aqk = A[(size_t)q * m + k];
          A[(size_t)p * m + k] = c * apk - sn * aqk;
          A[(size_t)q * m + k] = sn * apk + c * aqk;
        }
        for (int k = 0; k < m; ++k) {
          const double vkp = V[(size_t)k * m + p], vkq = V[(size_t)k * m + q];
          V[(size_t)k * m + p] = c * vkp - sn * vkq;
          V[(size_t)k * m + q] = sn * vkp + c * vkq;
        }
      }
  }
  d.resize(m);
  for (int i = 0; i < m; ++i) d[i] = A[(size_t)i * m + i];
}

// ---------------------------------------------------------------- host side
bool recompute_supported(const gfgpu_term *t) {
  if (t->mesh->gt_kind != GFGPU_GT_PK) return false;
  if (tf_of(t->family) < 0) return false;
  const int N = t->mesh->dim, nd = t->fem->nd, Q = t->fem->qdim;
  const bool ndok = N == 2 ? (nd == 3 || nd == 6 || nd == 10) : (nd == 4 || nd == 10 || nd == 20);
  if (!ndok) return false;
  if (t->family == GFGPU_ELASTICITY) return Q == N;
  return Q == 1 || Q == N;
}

static int env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return s && *s ? atoi(s) : dflt;
}

void recompute_prepare(gfgpu_term *t) {
  gfgpu_ctx *ctx = t->ctx;
  cudaStream_t s = ctx->stream;
  const int N = t->mesh->dim, nd = t->fem->nd, nq = t->tab->nq, Q = t->fem->qdim;
  const int rf = tf_of(t->family);
  const int MT = rf == TF_ELAST ? N * N : rf == TF_LAPLACE ? N * (N + 1) / 2 : 1;
  const int GSZ = MT;
  // ---- reference tensors M^{ji} from the staged tables (the same tables the reference integrates with)
  const std::vector<double> &w = t->tab->h_w, &g = t->tab->h_gphi, &ph = t->tab->h_phi;
  std::vector<double> M((size_t)nd * nd * MT, 0.0);
  for (int j = 0; j < nd; ++j)
    for (int i = 0; i < nd; ++i) {
      double *o = M.data() + ((size_t)j * nd + i) * MT;
      if (rf == TF_MASS) {
        double sum = 0;
        for (int k = 0; k < nq; ++k) sum += w[k] * ph[(size_t)k * nd + i] * ph[(size_t)k * nd + j];
        o[0] = sum;
        continue;
      }
      double full[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int k = 0; k < nq; ++k) {
        if (w[k] == 0.0) continue;
        const double *gi = g.data() + ((size_t)k * nd + i) * N, *gj = g.data() + ((size_t)k * nd + j) * N;
        for (int p = 0; p < N; ++p)
          for (int q = 0; q < N; ++q) full[p * N + q] += w[k] * gi[p] * gj[q];
      }
      if (rf == TF_ELAST) {
        for (int k = 0; k < N * N; ++k) o[k] = full[k];
      } else {  // symmetrised against G = a J B^T B (upper triangle, row by row)
        int k = 0;
        for (int p = 0; p < N; ++p)
          for (int q = p; q < N; ++q) o[k++] = p == q ? full[p * N + p] : full[p * N + q] + full[q * N + p];
      }
    }
  t->rc_M.alloc(ctx, M.size());
  t->rc_M.upload(M.data());
  {  // ---- low-rank factor of the Gram matrix of the reference (gradient) functions, for the residual kernel
    const int m = rf == TF_MASS ? nd : nd * N;
    std::vector<double> A((size_t)m * m, 0.0), ev, V;
    for (int k = 0; k < nq; ++k) {
      if (w[k] == 0.0) continue;
      const double *f = rf == TF_MASS ? ph.data() + (size_t)k * nd : g.data() + (size_t)k * nd * N;
      for (int r = 0; r < m; ++r)
        for (int c = 0; c < m; ++c) A[(size_t)r * m + c] += w[k] * f[r] * f[c];
    }
    jacobi_eig(A, m, ev, V);
    double emax = 0;
    for (double e : ev) emax = std::max(emax, fabs(e));
    std::vector<double> L;
    int rank = 0;
    for (int k = 0; k < m; ++k) {
      if (fabs(ev[k]) <= 1e-13 * emax) continue;
      L.push_back(ev[k] < 0 ? -1.0 : 1.0);
      const double sc = sqrt(fabs(ev[k]));
      for (int r = 0; r < m; ++r) L.push_back(sc * V[(size_t)r * m + k]);
      ++rank;
    }
    t->rc_rank = rank;
    t->rc_L.alloc(ctx, std::max<size_t>(L.size(), 1));
    t->rc_L.upload(L.data());
    GF_CUDA(cudaStreamSynchronize(s));
  }
  // ---- per-element geometry
  const int64_t ne = t->e1 - t->e0;
  t->rc_eg.alloc(ctx, (size_t)ne * GSZ);
  const double scale = t->alpha * ((rf == TF_ELAST) ? 1.0 : t->par[0]);
  const int64_t np = t->mesh->npts;
  const double *x = t->mesh->xyz.p;
  if (ne) {
    int grid = (int)std::min<int64_t>((ne + 255) / 256, 148 * 16);
    if (N == 2)
      k_tile_geo<2><<<grid, 256, 0, s>>>(x, x + np, x + 2 * np, t->conn_p(), t->mesh->ng, t->tab->gt_grad.p, t->e0, ne,
                                        rf, scale, t->rc_eg.p, GSZ);
    else
      k_tile_geo<3><<<grid, 256, 0, s>>>(x, x + np, x + 2 * np, t->conn_p(), t->mesh->ng, t->tab->gt_grad.p, t->e0, ne,
                                        rf, scale, t->rc_eg.p, GSZ);
    GF_LAUNCH_CHECK();
  }
  // ---- tiles: consecutive column nodes, greedily, bounded by incidences and pairs
  Structure &st = t->st;
  GF_REQUIRE(st.ncolnodes == st.nrnodes, "column nodes and incidence nodes differ");
  const int cbits = tl_clog2(nd * nd);
  GF_REQUIRE(cbits <= 10, "too many local nodes for the 16-bit contribution descriptor");
  const int slot_max = (1 << (16 - cbits)) - 1;  // the top slot index is the all-zero geometry
  const int cap_inc_want = std::max(1, std::min(std::min(env_int("GFGPU_TILE_INC", 510), slot_max), 1024));
  std::vector<uint32_t> rstart(st.nrnodes + 1), colstart(st.ncolnodes + 1);
  st.rstart.download(rstart.data());
  st.colstart.download(colstart.data());
  GF_CUDA(cudaStreamSynchronize(s));
  t->rc_cols = false;
  if (recompute_cols_wanted(t) && recompute_cols_prepare(t, colstart, rstart)) {  // low-order scalar forms: no tile plan at all
    t->rc_ready = true;
    return;
  }
  if (uniform_prepare(t)) {  // translated structure: class-uniform tiles, no per-pair records at all
    t->rc_ready = true;
    return;
  }
  // The pair capacity of a tile is lowered until two input buffers + two output images fit in shared memory.
  const int GSPh = GSZ | 1;
  const size_t smem_limit = 224 * 1024;
  std::vector<TileHdr> hdr;
  DevBuf<uint32_t> sp_pair, it_head, tk_tile;
  sp_pair.alloc(ctx, st.npairs);
  it_head.alloc(ctx, st.npairs);
  TileHdr *dh = nullptr;
  int cap_inc = 1, cap_slots = 1, cap_tasks = 1, cap_long = 1, cap_len = 1;
  int64_t ntask = 0, nlong = 0, nt = 0;
  int cap_pairs_want = std::max(32, std::min(env_int("GFGPU_TILE_PAIRS", 1280), 4095));
  for (;;) {
    hdr.clear();
    hdr.reserve(st.ncolnodes / 8 + 2);
    cap_inc = 1;
    for (int64_t k = 0; k < st.ncolnodes;) {
      int64_t k1 = k + 1;
      while (k1 < st.ncolnodes && rstart[k1 + 1] - rstart[k] <= (uint32_t)cap_inc_want &&
             colstart[k1 + 1] - colstart[k] <= (uint32_t)cap_pairs_want)
        ++k1;
      TileHdr h;
      memset(&h, 0, sizeof h);
      h.pair0 = colstart[k];
      h.npairs = colstart[k1] - colstart[k];
      h.node0 = (uint32_t)k;
      h.nnodes = (uint32_t)(k1 - k);
      cap_inc = std::max<int>(cap_inc, (int)(rstart[k1] - rstart[k]));
      hdr.push_back(h);
      k = k1;
    }
    GF_REQUIRE(cap_inc <= std::min(slot_max, 1024), "node valence too high for the compact tile descriptors: use strategy STAGED");
    nt = (int64_t)hdr.size();
    t->rc_hdr.alloc(ctx, hdr.size() * sizeof(TileHdr));
    GF_CUDA(cudaMemcpyAsync(t->rc_hdr.p, hdr.data(), hdr.size() * sizeof(TileHdr), cudaMemcpyHostToDevice, s));
    dh = (TileHdr *)t->rc_hdr.p;
    t->rc_els.alloc(ctx, (size_t)nt * cap_inc);
    t->flag.zero();
    {
      int P = 32;
      while (P < cap_inc) P <<= 1;
      k_tile_elements<<<(unsigned)nt, 128, 2 * P * sizeof(uint32_t), s>>>(dh, st.rstart.p, st.rsrc.p, nd, cap_inc, t->rc_els.p);
      GF_LAUNCH_CHECK();
    }
    k_tile_sort_pairs<<<(unsigned)nt, 256, 0, s>>>(dh, st.cstart.p, st.csrc.p, st.pJ.p, (uint32_t)(nd * nd),
                                                   (uint32_t)st.ncontrib, env_int("GFGPU_TILE_SIGSORT", 1), sp_pair.p,
                                                   it_head.p, (int *)t->flag.p);
    GF_LAUNCH_CHECK();
    k_tile_base<<<(unsigned)std::min<int64_t>((nt + 255) / 256, 148 * 8), 256, 0, s>>>(dh, nt, st.pJ.p, st.rdof.p, Q, t->jc.p);
    GF_LAUNCH_CHECK();
    // task ranges, blob ranges and the shared-memory capacities (host scan over the tiles)
    GF_CUDA(cudaMemcpyAsync(hdr.data(), dh, hdr.size() * sizeof(TileHdr), cudaMemcpyDeviceToHost, s));
    GF_CUDA(cudaStreamSynchronize(s));
    cap_slots = cap_tasks = cap_long = cap_len = 1;
    ntask = nlong = 0;
    for (TileHdr &h : hdr) {
      h.ntasks = h.n_wide + (h.nitems + 31) / 32;
      h.task0 = (uint32_t)ntask;
      h.r2_0 = (uint32_t)nlong;
      ntask += h.ntasks;
      nlong += h.n_long;
      cap_slots = std::max<int>(cap_slots, (int)h.nel);
      cap_tasks = std::max<int>(cap_tasks, (int)h.ntasks);
      cap_long = std::max<int>(cap_long, (int)h.n_long);
      cap_len = std::max<int>(cap_len, (int)h.len);
    }
    const size_t in_b = ((size_t)cap_tasks * 512 * TL_KG + (size_t)cap_long * TL_BLOB * TL_KG + (size_t)(cap_slots + 1) * GSPh * 8 + 127) / 128 * 128;
    const size_t out_b = ((size_t)(cap_len + 2) * 8 + 127) / 128 * 128;
    const size_t need = 2 * in_b + 2 * out_b + (size_t)nd * nd * MT * 8;
    if (need <= smem_limit) break;
    GF_REQUIRE(cap_pairs_want > 32, "a single column node does not fit the tile kernel's shared memory: use strategy STAGED");
    cap_pairs_want = std::max(32, cap_pairs_want * 3 / 4);
  }
  t->rc_nt = nt;
  t->rc_cap_inc = cap_inc;
  t->rc_cap_len = cap_len;
  if (getenv("GFGPU_DEBUG"))
    fprintf(stderr, "[gfgpu] tiles: %lld (pairs/tile <= %d), tasks %lld (long %lld), caps: incidences %d slots %d tasks %d long %d len %d\n",
            (long long)nt, cap_pairs_want, (long long)ntask, (long long)nlong, cap_inc, cap_slots, cap_tasks, cap_long, cap_len);
  GF_REQUIRE(ntask < (int64_t(1) << 31) && nlong < (int64_t(1) << 31), "too many tasks");
  GF_CUDA(cudaMemcpyAsync(dh, hdr.data(), hdr.size() * sizeof(TileHdr), cudaMemcpyHostToDevice, s));
  t->rc_ntask = ntask;
  t->rc_cap_slots = cap_slots;
  t->rc_cap_tasks = cap_tasks;
  t->rc_cap_long = cap_long;
  tk_tile.alloc(ctx, ntask);
  k_tile_task_owner<<<(unsigned)std::min<int64_t>(nt, 148 * 32), 128, 0, s>>>(dh, nt, tk_tile.p);
  GF_LAUNCH_CHECK();
  t->rc_prec.alloc(ctx, (size_t)ntask * 32 * TL_KG);
  t->rc_dblob.alloc(ctx, std::max<size_t>((size_t)nlong * TL_INREC * 32 * TL_KG, 1));
  {
    const int B = 256;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((ntask * 32 + B - 1) / B, 148 * 64));
#define GF_FILL(QQ)                                                                                                  \
  k_tile_fill<QQ><<<grid, B, 0, s>>>(dh, tk_tile.p, sp_pair.p, it_head.p, t->rc_els.p, cap_inc, st.cstart.p, st.csrc.p, \
                                     st.pJ.p, t->pmask.p, t->prel.p, t->jc.p, st.npairs, nd, cbits,                 \
                                     (uint32_t)cap_slots, ntask, (uint32_t)st.ncontrib, (uint4 *)t->rc_prec.p,      \
                                     t->rc_dblob.p,                                                                 \
                                     (int *)t->flag.p)
    if (Q == 1) GF_FILL(1);
    else if (Q == 2) GF_FILL(2);
    else GF_FILL(3);
#undef GF_FILL
    GF_LAUNCH_CHECK();
  }
  int32_t err = 0;
  t->flag.download(&err);
  GF_CUDA(cudaStreamSynchronize(s));
  GF_REQUIRE(err == 0, "recompute plan failed (code " + std::to_string(err) + "): tile too large for the compact records");
  if (!t->halo) t->prel.release();  // folded into the pair records
  // ---- per-tile geometry blobs (one bulk copy per tile in the tangent kernel); the element lists are then done with
  {
    int rows = cap_slots + 1;
    if ((rows * GSPh) & 1) ++rows;  // every tile's blob starts on a 16-byte boundary (bulk copy source alignment)
    t->rc_geo_rows = rows;
    t->rc_tgeo.alloc(ctx, (size_t)nt * rows * GSPh);
    t->rc_tgeo.zero();
    if (nt) {
      k_tile_gather_geo<<<(unsigned)nt, 128, 0, s>>>(dh, t->rc_els.p, cap_inc, t->rc_eg.p, GSZ, GSPh, rows, t->rc_tgeo.p);
      GF_LAUNCH_CHECK();
    }
    GF_CUDA(cudaStreamSynchronize(s));
    t->rc_els.release();
  }
  t->rc_ready = true;
}

template <int N, int Q, int ND, int RF>
static void launch_tiles(gfgpu_term *t, const double *U, bool do_t, bool do_r) {
  using C = TlCfg<N, RF>;
  const double sign = t->alpha < 0 ? -1.0 : 1.0;
  const bool fused_r = do_r && do_t && t->rc_cols;  // the column kernel forms R = K U next to the tangent
  if (do_r && !fused_r) {  // per-element residual -> stage -> fixed-order gather per node
    const int64_t ne = t->e1 - t->e0;
    if (t->rstage.n != (size_t)ne * ND * Q) t->rstage.alloc(t->ctx, (size_t)ne * ND * Q);
    ResArgs r;
    r.edof = t->edof_p(); r.eg = t->rc_eg.p; r.Ltab = t->rc_L.p; r.U = U;
    r.rank = t->rc_rank;
    r.sl = sign * t->par[0]; r.smu = sign * t->par[1];
    r.e0 = t->e0; r.ne = ne; r.rstage = t->rstage.p;
    const size_t smem = (size_t)t->rc_rank * ((RF == TF_MASS ? ND : ND * N) + 1) * 8;
    auto kern = k_affine_residual<N, Q, ND, RF>;
    GF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int rb = t->rc_res_block;
    static const bool no_split = getenv("GFGPU_NO_RSPLIT") != nullptr;
    bool launched = false;
    if constexpr (RF == TF_ELAST) {
      if (!no_split) {  // lane per (element, component)
        auto ks = k_affine_residual_split<N, ND>;
        GF_CUDA(cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int64_t groups = (ne + 32 / N - 1) / (32 / N), wpb = rb / 32;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((groups + wpb - 1) / wpb, (int64_t)t->ctx->sm_count * 16 * (128 / rb)));
        ks<<<grid, rb, smem, t->ctx->stream>>>(r);
        GF_LAUNCH_CHECK();
        launched = true;
      }
    }
    if (!launched) {
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((ne + rb - 1) / rb, (int64_t)t->ctx->sm_count * 8 * (128 / rb)));
      kern<<<grid, rb, smem, t->ctx->stream>>>(r);
      GF_LAUNCH_CHECK();
    }
    gather_residual(t);
  }
  if (!do_t) return;
  if (t->rc_cols) {
    recompute_cols_tangent(t, U, fused_r);
    return;
  }
  if (t->rc_uni) {
    uniform_tangent(t);
    return;
  }
  using L = TlSmem<N, RF>;
  TileArgs a;
  a.hdr = (const TileHdr *)t->rc_hdr.p;
  a.tgeo = t->rc_tgeo.p;
  a.geo_rows = t->rc_geo_rows;
  a.rec = (const uint4 *)t->rc_prec.p;
  a.dblob = t->rc_dblob.p;
  a.Mtab = t->rc_M.p;
  a.sl = sign * t->par[0]; a.smu = sign * t->par[1];
  a.nt = t->rc_nt;
  a.zslot = t->rc_cap_slots; a.cap_tasks = t->rc_cap_tasks; a.cap_long = t->rc_cap_long; a.cap_len = t->rc_cap_len;
  a.pr = t->pr.p;
  a.trace = nullptr;
  DevBuf<unsigned long long> trace;
  if (getenv("GFGPU_TILE_TRACE")) {
    trace.alloc(t->ctx, 256 * 12);
    trace.zero();
    a.trace = trace.p;
  }
  const size_t smem = 2 * L::bytes(a.cap_tasks, a.cap_long, a.zslot) + 2 * L::out_bytes(a.cap_len) + (size_t)ND * ND * C::MT * 8;
  GF_REQUIRE(smem <= 226 * 1024, "tile too large for shared memory; use strategy STAGED");
  auto kern = k_tiles<N, Q, ND, RF>;
  GF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<int64_t>(a.nt, (int64_t)t->ctx->sm_count);
  kern<<<grid, TL_THREADS, smem, t->ctx->stream>>>(a);
  GF_LAUNCH_CHECK();
  if (a.trace) {
    std::vector<unsigned long long> h(256 * 12);
    trace.download(h.data());
    GF_CUDA(cudaStreamSynchronize(t->ctx->stream));
    FILE *f = fopen(getenv("GFGPU_TILE_TRACE"), "w");
    if (f) {
      fprintf(f, "it p_wait_start p_empty_ok p_issued p_fetched c_wait_start c_full_ok c_done ntasks p_tma_issued p_ldgsts_issued nel len (ns, relative)\n");
      const unsigned long long t0 = h[0];
      for (int k = 0; k < 256; ++k) {
        fprintf(f, "%d", k);
        for (int j = 0; j < 7; ++j) fprintf(f, " %lld", (long long)(h[k * 12 + j] - t0));
        fprintf(f, " %llu %lld %lld %llu %llu\n", h[k * 12 + 7], (long long)(h[k * 12 + 8] - t0), (long long)(h[k * 12 + 9] - t0),
                h[k * 12 + 10], h[k * 12 + 11]);
      }
      fclose(f);
    }
  }
}

#define TL_CASE(NN, QQ, NDD, RFF)                                  \
  if (N == NN && Q == QQ && nd == NDD && rf == RFF) {              \
    launch_tiles<NN, QQ, NDD, RFF>(t, U, do_t, do_r);              \
    return;                                                        \
  }

// tangent and/or residual (R = K^T U = K U, the handled forms are symmetric) in one kernel
void recompute_assemble(gfgpu_term *t, const double *U, bool do_t, bool do_r) {
  if (!t->st.npairs) {  // no element in the range (empty region / a rank without elements): empty tangent, ZERO residual
    if (do_r) t->R.zero();
    return;
  }
  const int N = t->mesh->dim, nd = t->fem->nd, Q = t->fem->qdim, rf = tf_of(t->family);
  TL_CASE(3, 3, 10, TF_ELAST) TL_CASE(3, 3, 4, TF_ELAST) TL_CASE(3, 3, 20, TF_ELAST)
  TL_CASE(3, 1, 10, TF_LAPLACE) TL_CASE(3, 1, 4, TF_LAPLACE) TL_CASE(3, 1, 20, TF_LAPLACE)
  TL_CASE(3, 3, 10, TF_LAPLACE) TL_CASE(3, 3, 4, TF_LAPLACE) TL_CASE(3, 3, 20, TF_LAPLACE)
  TL_CASE(3, 1, 10, TF_MASS) TL_CASE(3, 1, 4, TF_MASS) TL_CASE(3, 1, 20, TF_MASS)
  TL_CASE(3, 3, 10, TF_MASS) TL_CASE(3, 3, 4, TF_MASS) TL_CASE(3, 3, 20, TF_MASS)
  TL_CASE(2, 2, 3, TF_ELAST) TL_CASE(2, 2, 6, TF_ELAST) TL_CASE(2, 2, 10, TF_ELAST)
  TL_CASE(2, 1, 3, TF_LAPLACE) TL_CASE(2, 1, 6, TF_LAPLACE) TL_CASE(2, 1, 10, TF_LAPLACE)
  TL_CASE(2, 2, 3, TF_LAPLACE) TL_CASE(2, 2, 6, TF_LAPLACE) TL_CASE(2, 2, 10, TF_LAPLACE)
  TL_CASE(2, 1, 3, TF_MASS) TL_CASE(2, 1, 6, TF_MASS) TL_CASE(2, 1, 10, TF_MASS)
  TL_CASE(2, 2, 3, TF_MASS) TL_CASE(2, 2, 6, TF_MASS) TL_CASE(2, 2, 10, TF_MASS)
  GF_REQUIRE(false, "no recompute kernel for this combination");
}

}  // namespace gf
