// recompute_uniform.cu -- strategy RECOMPUTE, class-uniform tile kernel: the per-nonzero tangent of an affine-geometry,
// constant-coefficient bilinear form on meshes with translated structure.
//
// Same mathematics and the same CSC pattern as recompute_tiles.cu (reference tensors M^{ji}, T = B~ M^{ji} B~^T per
// contribution; replaces ga_exec + add_elem_matrix, C&E.cc:8750-8870 / 4853-4936); what changes is WHO does a contribution:
//
//   lane  = a column node.  A tile is up to 32 column nodes of one CLASS (uniform_plan.h: word-for-word equal descriptors,
//           i.e. translated copies of one node): at every step the 32 lanes handle the SAME local couple (j, i) of 32
//           DIFFERENT elements, so the reference tensor M^{ji} is one shared-memory broadcast (16-byte loads, 5 per step)
//           and the program -- which pairs, which steps, where the results go -- is stored once per class instead of once
//           per (pair, lane): the 16-byte pair records and descriptor blobs of the general kernel (8 GB on BASELINE
//           config 3) shrink to 4 bytes per (column, incident element) + 8 bytes per column.
//   task  = up to 3 node pairs of the column that are fed by the same elements (a P2 edge midside node and the two vertices
//           of its edge; a vertex and the midside node towards it, ...): the geometry row B~_e of a step is loaded ONCE into
//           registers and applied to all of them -- register-level operand reuse.
//   B~_e  = read straight from global memory, coalesced: the geometry table is stored component-major in STRIP ORDER
//           (elements sorted by (rank in the incidence list of their first node, class position of that node)), so the
//           elements the 32 lanes need at a step are consecutive: one 256-byte line pair per component, no shared-memory
//           staging, no bank conflicts.  The table is read from L2 (0.6 GB for config 3), L1 keeps the rows of the tiles in
//           flight.
//   image = per tile, the CSC segments of its columns in shared memory (one row per lane, stride = 2 mod 16 doubles);
//           when the last task of a tile has stored its results the finishing warp sends every row with one bulk async
//           store (TMA, full 16-byte units; the odd first / last entries go out as single stores).  Several tiles are in
//           flight per CTA (ring of image buffers), tasks are handed out through one shared-memory counter, longest first.
//
// Columns whose class is small simply run with few active lanes; terms whose columns mostly have no translated copies
// (unstructured meshes) keep the general tile kernel (recompute_tiles.cu).  Every pair still sums its contributions in a
// fixed order (ascending element): bitwise reproducible, no atomics on data.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "common.cuh"
#include "tile_common.cuh"
#include "uniform_plan.h"

namespace gf {

static inline int ugrid(int64_t n, int block, int cap = 148 * 16) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + block - 1) / block, cap));
}
static int uenv_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return s && *s ? atoi(s) : dflt;
}

// ---------------------------------------------------------------- column descriptors
struct UIn {
  const uint32_t *colstart, *cstart, *csrc, *rstart, *rsrc;
  const int32_t *rdof;
  const uint16_t *pmask;
  const uint32_t *prel;
  const int64_t *jc;
  int64_t npairs;
  uint32_t nlocal;
  int nd, Q;
};

// the descriptor words of column node k (uniform_plan.h), in order; false when an incidence is missing (corrupt structure)
template <class F>
__device__ __forceinline__ bool ut_col_words(const UIn &in, int64_t k, F &&f) {
  const uint32_t p0 = in.colstart[k], p1 = in.colstart[k + 1], r0 = in.rstart[k], r1 = in.rstart[k + 1];
  const int32_t J = in.rdof[k];
  const int64_t j0 = in.jc[J];
  f(p1 - p0);
  f(r1 - r0);
  for (int b = 0; b < in.Q; ++b) f((uint32_t)(in.jc[J + b + 1] - j0));
  const uint32_t nb = (uint32_t)(in.nd * in.nd);
  bool ok = true;
  for (uint32_t p = p0; p < p1; ++p) {
    const uint32_t s0 = in.cstart[p], s1 = in.cstart[p + 1];
    uint32_t cntl = 0;
    while (s0 + cntl < s1 && in.csrc[s0 + cntl] < in.nlocal) ++cntl;  // the virtual (halo) contributions come last
    f((uint32_t)in.pmask[p] | (cntl << 16));
    for (int b = 0; b < in.Q; ++b) f(in.prel[(size_t)b * in.npairs + p]);
    for (uint32_t s = s0; s < s0 + cntl; ++s) {
      const uint32_t c = in.csrc[s], el = c / nb, rr = c - el * nb, key = el * (uint32_t)in.nd + rr / (uint32_t)in.nd;
      uint32_t lo = r0, hi = r1;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (in.rsrc[mid] < key) lo = mid + 1; else hi = mid;
      }
      if (lo >= r1 || in.rsrc[lo] != key) ok = false;
      f(((lo - r0) << 16) | rr);
    }
  }
  return ok;
}

__global__ void k_ut_hash(const UIn in, int64_t ncol, uint64_t *__restrict__ h1, uint64_t *__restrict__ h2,
                          uint32_t *__restrict__ dlen, uint32_t *__restrict__ ids, int *__restrict__ err) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < ncol; k += (int64_t)gridDim.x * blockDim.x) {
    uint64_t a = 0x9E3779B97F4A7C15ull, b = 0xC2B2AE3D27D4EB4Full;
    uint32_t n = 0;
    const bool ok = ut_col_words(in, k, [&](uint32_t w) {
      a = (a ^ w) * 0x100000001B3ull;
      a ^= a >> 29;
      b = (b + w) * 0xD6E8FEB86659FD93ull;
      b ^= b >> 32;
      ++n;
    });
    if (!ok) atomicExch(err, 21);
    const uint32_t r = in.rstart[k + 1] - in.rstart[k];
    if (r > 0xffffu || in.colstart[k + 1] - in.colstart[k] > 0xffffu) atomicExch(err, 22);
    h1[k] = a;
    h2[k] = b;
    dlen[k] = n;
    ids[k] = (uint32_t)k;
  }
}

// sorted position s starts a class when (h1, h2) differ from the position before
__global__ void k_ut_class_flags(const uint64_t *__restrict__ h1s, const uint64_t *__restrict__ h2,
                                 const uint32_t *__restrict__ scol, int64_t ncol, uint8_t *__restrict__ flags) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < ncol; s += (int64_t)gridDim.x * blockDim.x)
    flags[s] = (s == 0 || h1s[s] != h1s[s - 1] || h2[scol[s]] != h2[scol[s - 1]]) ? 1 : 0;
}

__global__ void k_ut_leader_len(const uint32_t *__restrict__ cls_start, const uint32_t *__restrict__ scol,
                                const uint32_t *__restrict__ dlen, int64_t ncls, uint32_t *__restrict__ out) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncls; c += (int64_t)gridDim.x * blockDim.x)
    out[c] = dlen[scol[cls_start[c]]];
}

__global__ void k_ut_leader_desc(const UIn in, const uint32_t *__restrict__ cls_start, const uint32_t *__restrict__ scol,
                                 int64_t ncls, uint32_t dw, uint32_t *__restrict__ desc) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncls; c += (int64_t)gridDim.x * blockDim.x) {
    uint32_t *o = desc + (size_t)c * dw;
    uint32_t n = 0;
    ut_col_words(in, scol[cls_start[c]], [&](uint32_t w) {
      if (n < dw) o[n] = w;
      ++n;
    });
  }
}

__global__ void k_ut_inverse(const uint32_t *__restrict__ scol, int64_t ncol, uint32_t *__restrict__ sp) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < ncol; s += (int64_t)gridDim.x * blockDim.x)
    sp[scol[s]] = (uint32_t)s;
}

// strip order of the elements: key = (rank of the element in the incidence list of its local node 0, class position of
// that node).  Translated copies of an element then sit next to each other, in the order of the lanes that use them.
__global__ void k_ut_elem_key(const int32_t *__restrict__ edof, int nd, int64_t e0, int64_t ne,
                              const int32_t *__restrict__ rdof, int64_t ncol, const uint32_t *__restrict__ rstart,
                              const uint32_t *__restrict__ rsrc, const uint32_t *__restrict__ sp,
                              uint64_t *__restrict__ keys, uint32_t *__restrict__ ids, int *__restrict__ err) {
  for (int64_t el = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; el < ne; el += (int64_t)gridDim.x * blockDim.x) {
    const int32_t dof = edof[(e0 + el) * nd];
    int64_t lo = 0, hi = ncol;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (rdof[mid] < dof) lo = mid + 1; else hi = mid;
    }
    uint64_t key = ~0ull;
    if (lo < ncol && rdof[lo] == dof) {
      const uint32_t r0 = rstart[lo], r1 = rstart[lo + 1], want = (uint32_t)el * (uint32_t)nd;
      uint32_t a = r0, b = r1;
      while (a < b) {
        const uint32_t mid = (a + b) >> 1;
        if (rsrc[mid] < want) a = mid + 1; else b = mid;
      }
      if (a < r1 && rsrc[a] == want) key = ((uint64_t)(a - r0) << 32) | sp[lo];
    }
    if (key == ~0ull) atomicExch(err, 23);
    keys[el] = key;
    ids[el] = (uint32_t)el;
  }
}

__global__ void k_ut_scatter_pos(const uint32_t *__restrict__ order, int64_t ne, uint32_t *__restrict__ epos) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < ne; i += (int64_t)gridDim.x * blockDim.x)
    epos[order[i]] = (uint32_t)i;
}

__global__ void k_ut_eg_blocked(const double *__restrict__ eg, int gsz, int64_t ne, const uint32_t *__restrict__ epos,
                                double *__restrict__ out) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < ne * gsz; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t el = idx / gsz;
    const int c = (int)(idx - el * gsz);
    const uint32_t pos = epos[el];
    out[(size_t)(pos >> 5) * (gsz * 32) + c * 32 + (pos & 31u)] = eg[idx];
  }
}

// thread per (chunk, lane): the lane data of the chunk's tiles -- CSC base of the lane's column (two 32-bit halves) and the
// strip position of every element of its incidence list -- after checking that the column's descriptor IS the leader's
struct UChunk {
  uint32_t pos0, cls, ld, nmem;
};
__global__ void k_ut_lane_data(const UIn in, const UChunk *__restrict__ chunks, int64_t nchunks,
                               const uint32_t *__restrict__ scol, const uint32_t *__restrict__ desc, uint32_t dw,
                               const uint32_t *__restrict__ epos, int gsz, uint32_t *__restrict__ ld, int *__restrict__ err) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < nchunks * 32; idx += (int64_t)gridDim.x * blockDim.x) {
    const UChunk ch = chunks[idx >> 5];
    const uint32_t lane = (uint32_t)(idx & 31);
    uint32_t *o = ld + ch.ld;
    const uint32_t m = desc[(size_t)ch.cls * dw + 1];
    if (lane >= ch.nmem) {
      o[lane] = 0; o[32 + lane] = 0;
      for (uint32_t r = 0; r < m; ++r) o[64 + r * 32 + lane] = 0;
      continue;
    }
    const int64_t k = scol[ch.pos0 + lane];
    const uint32_t *L = desc + (size_t)ch.cls * dw;
    uint32_t n = 0;
    bool same = true;
    const bool ok = ut_col_words(in, k, [&](uint32_t w) {
      if (n >= dw || L[n] != w) same = false;
      ++n;
    });
    if (!ok || !same) atomicExch(err, 24);
    const int64_t jc = in.jc[in.rdof[k]];
    o[lane] = (uint32_t)(jc & 0xffffffffll);
    o[32 + lane] = (uint32_t)(jc >> 32);
    const uint32_t r0 = in.rstart[k];
    for (uint32_t r = 0; r < m; ++r) {  // byte offset of the element's entry in the blocked geometry table
      const uint32_t pos = epos[in.rsrc[r0 + r] / (uint32_t)in.nd];
      o[64 + r * 32 + lane] = ((pos >> 5) * (uint32_t)(gsz * 32) + (pos & 31u)) * 8u;
    }
  }
}

// ---------------------------------------------------------------- the tangent kernel
struct alignas(16) UTile {
  uint32_t prog, ld;  // program (8-byte units), lane data (words)
  uint32_t nmem, pad;
};
static_assert(sizeof(UTile) == 16, "UTile layout");

struct UArgs {
  const UTile *tiles;
  const uint32_t *cta_t0;  // grid + 1: tiles of CTA c are [cta_t0[c], cta_t0[c+1])
  const uint2 *prog;
  const uint32_t *ld;
  const double *eg;        // strip order, blocked: [strip position / 32][GSZ][strip position % 32]
  const double *Mtab;
  double sl, smu;
  double *pr;
  int imgcap;              // doubles per image buffer (one per team)
  int dbg;                 // experiments (GFGPU_UT_DBG): 1 = no flush of the images, 2 = no arithmetic, 4 = no team barriers
};

// A CTA = UT_TEAMS teams of UT_TW warps (one per SM sub-partition); a team runs one tile at a time: every warp its task (the
// plan packs the tile's groups into UT_TW tasks), team barrier, every warp sends UT_ROWS rows of the image, team barrier.
// No spinning, no atomics on data; the teams are independent, so one team's flush overlaps the others' arithmetic.


__device__ __forceinline__ void team_barrier(int team, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(nthreads) : "memory");
}

// flush of a tile image: warp wq sends rows [wq*UT_ROWS, (wq+1)*UT_ROWS) -- the CSC segments of these columns.  Four lanes per
// row (8 rows side by side), 16-byte units: every store instruction writes 8 runs of 64 contiguous bytes (full sectors except at
// the two ends of a row); the odd first / last entry of a row goes out as a single 8-byte store of lane 0 / 1 of its quad.
// Kept out of line: its address registers must not weigh on the allocation of the arithmetic loop.
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
// program words are shared by every tile of a class: ask L1 to keep them (the geometry rows stream through the same cache)
__device__ __forceinline__ uint2 ldg_keep_u2(const uint2 *p) {
  uint2 v;
#ifdef GF_UT_NO_EVICT_LAST
  v = __ldg(p);
#else
  asm volatile("ld.global.nc.L1::evict_last.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
#endif
  return v;
}
template <int ROWS>  // rows per warp (32 / warps of a team): 32 / ROWS lanes per row
__device__ __noinline__ void ut_flush_rows(uint32_t img_s /* shared address of the image */, double *pr, uint32_t rowstride,
                                           uint32_t nmem, int wq, int lane, uint32_t jclo, uint32_t jchi, uint32_t npieces,
                                           uint2 pc0, uint2 pc1, uint2 pc2) {
  constexpr int LPR = 32 / ROWS;  // lanes per row
  const uint32_t k = (uint32_t)(wq * ROWS + lane / LPR), sub = (uint32_t)(lane % LPR);
  const uint32_t lo = __shfl_sync(0xffffffffu, jclo, k), hi = __shfl_sync(0xffffffffu, jchi, k);
  const int64_t jc = (int64_t)(((uint64_t)hi << 32) | lo);
  if (k >= nmem) return;
  for (uint32_t pc = 0; pc < npieces; ++pc) {
    const uint2 pw = pc == 0 ? pc0 : pc == 1 ? pc1 : pc2;
    const uint32_t len = pw.y & 0xffffu;
    if (!len) continue;
    const int64_t gstart = jc + pw.x;
    const uint32_t odd = (uint32_t)(gstart & 1), tail = (odd + len) & 1u;
    const uint32_t src = img_s + (k * rowstride + (pw.y >> 16) + odd) * 8u;  // entry e of the piece sits at src + 8 e
    double *g = pr + gstart;
    if (sub == 0 && odd) g[0] = lds_f64(src);
    if (sub == 1 && tail && len - 1u >= odd) g[len - 1u] = lds_f64(src + (len - 1u) * 8u);
    const int nun = (int)(len - odd - tail) >> 1;  // aligned 16-byte units, the first one at entry `odd`
    double2 *d2 = reinterpret_cast<double2 *>(g + odd);
    const uint32_t s2 = src + odd * 8u;
    int u = (int)sub;
    for (; u + 3 * LPR < nun; u += 4 * LPR) {
      const double2 v0 = lds_f64x2(s2 + u * 16u), v1 = lds_f64x2(s2 + (u + LPR) * 16u), v2 = lds_f64x2(s2 + (u + 2 * LPR) * 16u),
                    v3 = lds_f64x2(s2 + (u + 3 * LPR) * 16u);
      d2[u] = v0; d2[u + LPR] = v1; d2[u + 2 * LPR] = v2; d2[u + 3 * LPR] = v3;
    }
    for (; u < nun; u += LPR) d2[u] = lds_f64x2(s2 + u * 16u);
  }
}

// asynchronous copies global -> shared without registers (LDGSTS): the warp's next instruction stream and, a few steps
// ahead, the strip positions of the elements it is going to need
__device__ __forceinline__ void cp_async16(uint32_t dst_s, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_s), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst_s, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_s), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NLEFT>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NLEFT) : "memory"); }
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

constexpr int UT_SB = 64;  // units of a task's instruction stream (one 16-byte asynchronous copy per lane); the plan enforces it
constexpr int UT_PR = 8;   // slots of a warp's ring of strip positions
constexpr int UT_PD = 6;   // a step's strip positions are requested UT_PD instructions ahead
constexpr int UT_WS = 2 * UT_SB * 8 + UT_PR * 128;  // bytes of shared memory per warp: two stream buffers + the ring

// asynchronous flush: lane r < ROWS of warp wq sends row wq*ROWS + r with one bulk async store (TMA) per piece; nobody waits
// here -- the copies read the image while the team already computes the first group of its next tile
template <int ROWS>
__device__ __noinline__ void ut_flush_rows_bulk(const double *img, double *pr, uint32_t rowstride, uint32_t nmem, int wq, int lane,
                                                uint32_t jclo, uint32_t jchi, uint32_t npieces, uint2 pc0, uint2 pc1, uint2 pc2) {
  const uint32_t k = (uint32_t)(wq * ROWS + (lane % ROWS));
  const uint32_t lo = __shfl_sync(0xffffffffu, jclo, k), hi = __shfl_sync(0xffffffffu, jchi, k);
  const int64_t jc = (int64_t)(((uint64_t)hi << 32) | lo);
  fence_async_smem();
  if (lane < ROWS && k < nmem) {
    for (uint32_t pc = 0; pc < npieces; ++pc) {
      const uint2 pw = pc == 0 ? pc0 : pc == 1 ? pc1 : pc2;
      const int64_t len = pw.y & 0xffffu;
      if (!len) continue;
      const int64_t gstart = jc + pw.x;
      const int64_t odd = gstart & 1, gs = gstart + odd, ge = (gstart + len) & ~int64_t(1);
      const double *src = img + (size_t)k * rowstride + (pw.y >> 16) + odd;  // entry e of the piece sits at src[e]
      if (odd) pr[gstart] = src[0];
      if (ge > gs) bulk_s2g(pr + gs, src + odd, (uint32_t)((ge - gs) * 8));
      if (((gstart + len) & 1) && gstart + len - 1 >= gs) pr[gstart + len - 1] = src[len - 1];
    }
    bulk_commit();
  }
}

template <int N, int Q, int ND, int RF, int KG, int UT_TW, int TEAMS>
__global__ void __launch_bounds__(TEAMS * UT_TW * 32, 1)
k_utiles(const UArgs a) {
  constexpr int UT_TEAMS = TEAMS, UT_THREADS = TEAMS * UT_TW * 32, UT_ROWS = 32 / UT_TW;
  using C = TlCfg<N, RF>;
  constexpr int NB = ND * ND, MT = C::MT, MTP = (MT + 1) & ~1, GSZ = C::GSZ, ACC = C::ACC;
  extern __shared__ __align__(128) unsigned char smraw[];
  double *sM = reinterpret_cast<double *>(smraw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, team = warp / UT_TW, wq = warp % UT_TW;
  double *img = sM + ((NB * MTP + 15) & ~15) + (size_t)team * a.imgcap;
  // per-warp scratch behind the images: instruction streams of the current / next tile, ring of strip positions
  const uint32_t ws_s = smem_u32(sM + ((NB * MTP + 15) & ~15) + (size_t)UT_TEAMS * a.imgcap) + (uint32_t)warp * UT_WS;
  const uint32_t ring_s = ws_s + 2 * UT_SB * 8 + (uint32_t)lane * 4u;
  for (int k = tid; k < NB * MTP; k += UT_THREADS) sM[k] = (k % MTP) < MT ? a.Mtab[(k / MTP) * MT + k % MTP] : 0.0;
  __syncthreads();
  const uint32_t t0 = a.cta_t0[blockIdx.x], ntl = a.cta_t0[blockIdx.x + 1] - t0;
  const uint32_t sM_s = smem_u32(sM);
  // Tiles go to the teams round robin, so the tiles a team will run are known from the start: the record of tile k+2, the
  // program header / task range / CSC base of tile k+1 are loaded while tile k runs, and the instruction stream of tile k+1
  // is copied to shared memory during the flush of tile k (a cold chain of dependent loads per tile, then per instruction,
  // kept every warp waiting for 85 % of its time: profiles/round2_ncu_utiles_v5_flush_4_lanes_per_row_c3_n110.txt)
  struct Meta {
    uint4 tw;
    uint2 h0, tr, pc0, pc1, pc2;
    uint32_t jclo, jchi;
  };
  auto load_tw = [&](uint32_t lt) { return __ldg(reinterpret_cast<const uint4 *>(a.tiles + t0 + min(lt, ntl - 1u))); };
  auto load_meta = [&](Meta &m) {
    const uint2 *P = a.prog + m.tw.x;
    m.h0 = __ldg(P);
    m.tr = __ldg(P + uplan::HDR_UNITS + wq);
    m.pc0 = __ldg(P + 1); m.pc1 = __ldg(P + 2); m.pc2 = __ldg(P + 3);
    const uint32_t le = min((uint32_t)lane, m.tw.z - 1u);
    m.jclo = __ldg(a.ld + m.tw.y + le);
    m.jchi = __ldg(a.ld + m.tw.y + 32 + le);
  };
  auto stream_copy = [&](const Meta &m, int buf) {  // the task's units [tr.x, tr.x + UT_SB): tr.x is even (16-byte source)
    cp_async16(ws_s + (uint32_t)buf * (UT_SB * 8) + (uint32_t)lane * 16u, a.prog + m.tw.x + m.tr.x + 2 * lane);
    cp_async_commit();
  };
  if (ntl == 0 || (uint32_t)team >= ntl) return;
  Meta cur;
  cur.tw = load_tw((uint32_t)team);
  load_meta(cur);
  stream_copy(cur, 0);
  uint4 tw1 = load_tw((uint32_t)team + UT_TEAMS);
  int it = 0;
  for (uint32_t lt = (uint32_t)team; lt < ntl; lt += UT_TEAMS, ++it) {
    Meta nxt;
    nxt.tw = tw1;
    load_meta(nxt);                       // tile lt + TEAMS: header, task range, CSC base (in flight during this tile)
    tw1 = load_tw(lt + 2 * UT_TEAMS);     // record of tile lt + 2 TEAMS
    const uint4 tw = cur.tw;
    const uint32_t nmem = tw.z;
    const uint2 tr = cur.tr, pc0 = cur.pc0, pc1 = cur.pc1, pc2 = cur.pc2;
    const uint32_t rowstride = cur.h0.x, npieces = cur.h0.y & 0xffu;
    const uint32_t le = min((uint32_t)lane, nmem - 1u);
    const char *lpos = reinterpret_cast<const char *>(a.ld + tw.y + 64 + le);
    const uint32_t jclo = cur.jclo, jchi = cur.jchi;
    const uint32_t row_s = smem_u32(img) + (uint32_t)lane * rowstride * 8u;  // shared address of the lane's image row
    const uint32_t par0 = (jclo + pc0.x) & 1u, par1 = (jclo + pc1.x) & 1u, par2 = (jclo + pc2.x) & 1u;
    const uint32_t sb_s = ws_s + (uint32_t)(it & 1) * (UT_SB * 8);  // my instruction stream, unit u at sb_s + 8 u
    cp_async_wait<0>();  // the stream of this tile (requested during the previous one) has landed
    __syncwarp();

    const uint32_t nu = (a.dbg & 2) ? 0u : tr.y - tr.x;  // units of my task
    // bulk flush (a.dbg & 8): the image may be overwritten once every warp's bulk stores of the previous tile have read it:
    // checked as late as possible, before the first store of the task
    bool img_ok = !(a.dbg & 8);
    auto image_ready = [&]() {
      if (!img_ok) {
        bulk_wait_read0();
        team_barrier(team, UT_TW * 32);
        img_ok = true;
      }
    };
    if (nu) {
      double acc[KG][ACC];
#pragma unroll
      for (int p = 0; p < KG; ++p)
#pragma unroll
        for (int m = 0; m < ACC; ++m) acc[p][m] = 0.0;
      auto unit = [&](uint32_t u) { return lds_u2(sb_s + min(u, nu - 1u) * 8u); };  // past the end: the final FLUSH again
      auto is_step = [](const uint2 &I) { return (int)I.x >= 0; };
      // strip position of unit u's element -> ring slot u % UT_PR (one group per unit, also for a FLUSH: the count stays aligned)
      auto request_pos = [&](uint32_t u) {
        const uint2 I = unit(u);
        if (u < nu && is_step(I)) cp_async4(ring_s + (u % UT_PR) * 128u, lpos + (I.x & 0x7ff80u));
        cp_async_commit();
      };
      auto load_g = [&](double (&G)[GSZ], uint32_t u) {  // geometry row of unit u: its strip position has landed in the ring
        const uint32_t off = lds_u32(ring_s + (u % UT_PR) * 128u);
        const double *gp = reinterpret_cast<const double *>(reinterpret_cast<const char *>(a.eg) + off);
#pragma unroll
        for (int c = 0; c < GSZ; ++c) G[c] = __ldg(gp + c * 32);
      };
#pragma unroll
      for (int d = 0; d < UT_PD; ++d) request_pos((uint32_t)d);
      double GA[GSZ], GB[GSZ];
#pragma unroll
      for (int c = 0; c < GSZ; ++c) GA[c] = GB[c] = 0.0;
      uint2 I0 = unit(0), I1 = unit(1);
      cp_async_wait<UT_PD - 1>();  // position of unit 0
      if (is_step(I0)) load_g(GA, 0);
      uint32_t ip = 0;

      auto flush_pair = [&](const double (&A)[ACC], const uint2 &I) {
        if ((uint32_t)lane >= nmem) return;
        double kv[Q * Q];  // kv[b2*Q + aa] = K(row component aa, column component b2)
        if (RF == TF_ELAST) {
          double trc = A[0];
#pragma unroll
          for (int n = 1; n < N; ++n) trc += A[RF == TF_ELAST ? n + N * n : 0];
          trc *= a.smu;
#pragma unroll
          for (int b2 = 0; b2 < Q; ++b2)
#pragma unroll
            for (int aa = 0; aa < Q; ++aa) {
              const double v = a.sl * A[RF == TF_ELAST ? aa + N * b2 : 0] + a.smu * A[RF == TF_ELAST ? b2 + N * aa : 0];
              kv[b2 * Q + aa] = aa == b2 ? v + trc : v;
            }
        } else {
#pragma unroll
          for (int b2 = 0; b2 < Q; ++b2)
#pragma unroll
            for (int aa = 0; aa < Q; ++aa) kv[b2 * Q + aa] = aa == b2 ? A[0] : 0.0;
        }
        if (I.x & (1u << 17)) {  // no local contribution (a pair announced by another rank): zeros
#pragma unroll
          for (int k = 0; k < Q * Q; ++k) kv[k] = 0.0;
        }
#pragma unroll
        for (int b2 = 0; b2 < Q; ++b2) {
          const uint32_t piece = (I.x >> (9 + 2 * b2)) & 3u;
          const uint32_t dst = row_s + ((((I.y >> (10 * b2)) & 0x3ffu) + (piece == 0 ? par0 : piece == 1 ? par1 : par2)) << 3);
          const unsigned mb = (I.x >> (b2 * Q)) & ((1u << Q) - 1);
          if (mb == (1u << Q) - 1) {  // the mask is the same for every lane: the common case takes no selects
#pragma unroll
            for (int aa = 0; aa < Q; ++aa) sts_f64(dst + 8 * aa, kv[b2 * Q + aa]);
          } else {
#pragma unroll
            for (int aa = 0; aa < Q; ++aa)
              if (mb & (1u << aa)) sts_f64(dst + 8 * __popc(mb & ((1u << aa) - 1)), kv[b2 * Q + aa]);
          }
        }
      };

      // one instruction: while unit ip runs, the geometry row of ip+1 is loaded and the strip position of ip+UT_PD requested
      auto phase = [&](double (&Gc)[GSZ], double (&Gn)[GSZ]) {
        request_pos(ip + UT_PD);
        cp_async_wait<UT_PD - 1>();  // position of unit ip + 1
        if (ip + 1 < nu && is_step(I1)) load_g(Gn, ip + 1);
        if (is_step(I0)) {
          const int npairs = (int)((I0.x >> 29) & 3u);
          const uint32_t code[3] = {(I0.x >> 19) & 0x3ffu, I0.y & 0x3ffu, (I0.y >> 10) & 0x3ffu};
          if (I0.x & 1u) {  // first step of a group: its accumulators start from zero (a flush leaves them alone)
#pragma unroll
            for (int p = 0; p < KG; ++p)
#pragma unroll
              for (int m = 0; m < ACC; ++m) acc[p][m] = 0.0;
          }
#pragma unroll
          for (int p = 0; p < KG; ++p) {
            if (p < npairs) {
              double M[MTP];
              const uint32_t ma = sM_s + code[p] * (MTP * 8);
#pragma unroll
              for (int q = 0; q < MTP / 2; ++q) {
                const double2 v = lds_f64x2(ma + q * 16);
                M[2 * q] = v.x;
                M[2 * q + 1] = v.y;
              }
              if (RF == TF_ELAST) {
#pragma unroll
                for (int qq = 0; qq < N; ++qq) {
                  double Wq[N];  // column qq of W = B~ M
#pragma unroll
                  for (int aa = 0; aa < N; ++aa) {
                    double s2 = Gc[RF == TF_ELAST ? aa : 0] * M[RF == TF_ELAST ? qq : 0];
#pragma unroll
                    for (int pp = 1; pp < N; ++pp) s2 += Gc[RF == TF_ELAST ? aa + N * pp : 0] * M[RF == TF_ELAST ? pp * N + qq : 0];
                    Wq[aa] = s2;
                  }
#pragma unroll
                  for (int b2 = 0; b2 < N; ++b2)
#pragma unroll
                    for (int aa = 0; aa < N; ++aa)
                      acc[p][RF == TF_ELAST ? aa + N * b2 : 0] += Wq[aa] * Gc[RF == TF_ELAST ? b2 + N * qq : 0];
                }
              } else {
                double s2 = acc[p][0];
#pragma unroll
                for (int k = 0; k < MT; ++k) s2 += M[k] * Gc[k];
                acc[p][0] = s2;
              }
            }
          }
        } else {
          image_ready();
          const uint32_t slot = (I0.x >> 15) & 3u;
          if (KG == 1 || slot == 0) flush_pair(acc[0], I0);
          else if (KG == 2 || slot == 1) flush_pair(acc[KG > 1 ? 1 : 0], I0);
          else flush_pair(acc[KG > 2 ? 2 : 0], I0);
        }
        I0 = I1;
        I1 = unit(ip + 2);
      };
      // (a geometry row TWO instructions ahead, three register buffers in rotation, was tried: slower -- the extra 18
      // registers cost more than the longer prefetch distance brought: profiles/round2_utiles_experiments.txt)
      for (;;) {
        phase(GA, GB);
        if (++ip >= nu) break;
        phase(GB, GA);
        if (++ip >= nu) break;
      }
    }
    image_ready();  // (a task without a flush)
    if (!(a.dbg & 4)) team_barrier(team, UT_TW * 32);  // the image of the tile is complete
    cp_async_wait<0>();              // (my ring requests past the end of the task)
    stream_copy(nxt, (it + 1) & 1);  // lands during the flush
    if (a.dbg & 8) {
      ut_flush_rows_bulk<UT_ROWS>(img, a.pr, rowstride, nmem, wq, lane, jclo, jchi, npieces, pc0, pc1, pc2);
    } else {
      if (!(a.dbg & 1)) ut_flush_rows<UT_ROWS>(smem_u32(img), a.pr, rowstride, nmem, wq, lane, jclo, jchi, npieces, pc0, pc1, pc2);
      if (!(a.dbg & 4)) team_barrier(team, UT_TW * 32);  // the image may be overwritten
    }
    cur = nxt;
  }
  cp_async_wait<0>();
  bulk_wait0();
}

// ---------------------------------------------------------------- host side
static int64_t ut_select_heads(gfgpu_ctx *ctx, const uint8_t *flags, int64_t n, uint32_t *out) {
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

// Builds the class-uniform plan of term t.  false = the term keeps the general tile kernel (GFGPU_UNIFORM=0, or too few of
// its columns have translated copies).  GFGPU_UNIFORM=2 forces the uniform kernel whatever the class sizes (tests).
bool uniform_prepare(gfgpu_term *t) {
  // 0 (default): general tile kernel.  1: this kernel when most columns have translated copies.  2: always (tests).
  // Measured on BASELINE config 3 (profiles/round2_utiles_experiments.txt): 10.4 ms against 8.6 ms for the general kernel --
  // 27 % less DRAM traffic and 9 GB less device memory, but the per-tile interpretation (team barriers, image flush, short
  // instruction streams) leaves the fp64 pipe idle 74 % of the time; it stays an opt-in until that is solved.
  const int mode = uenv_int("GFGPU_UNIFORM", 0);
  t->rc_uni = false;
  if (!mode) return false;
  gfgpu_ctx *ctx = t->ctx;
  cudaStream_t s = ctx->stream;
  Structure &st = t->st;
  const int64_t ncol = st.ncolnodes, ne = t->e1 - t->e0;
  const int nd = t->fem->nd, Q = t->fem->qdim, N = t->mesh->dim;
  const int rf = tf_of(t->family);
  const int GSZ = rf == TF_ELAST ? N * N : rf == TF_LAPLACE ? N * (N + 1) / 2 : 1;
  if (!ncol || !ne || ncol >= (int64_t(1) << 32) || ne >= (int64_t(1) << 32)) return false;
  const int B = 256;
  UIn in;
  in.colstart = st.colstart.p; in.cstart = st.cstart.p; in.csrc = st.csrc.p; in.rstart = st.rstart.p; in.rsrc = st.rsrc.p;
  in.rdof = st.rdof.p; in.pmask = t->pmask.p; in.prel = t->prel.p; in.jc = t->jc.p;
  in.npairs = st.npairs; in.nlocal = (uint32_t)st.ncontrib; in.nd = nd; in.Q = Q;
  GF_REQUIRE(t->prel.n == (size_t)Q * st.npairs, "uniform plan: the pattern offsets are gone");
  t->flag.zero();
  // ---- classes: hash of the descriptor, sort, run heads
  DevBuf<uint64_t> h1, h2, h1s;
  DevBuf<uint32_t> dlen, ids, scol;
  h1.alloc(ctx, ncol); h2.alloc(ctx, ncol); h1s.alloc(ctx, ncol);
  dlen.alloc(ctx, ncol); ids.alloc(ctx, ncol); scol.alloc(ctx, ncol);
  k_ut_hash<<<ugrid(ncol, B), B, 0, s>>>(in, ncol, h1.p, h2.p, dlen.p, ids.p, (int *)t->flag.p);
  GF_LAUNCH_CHECK();
  {
    size_t tb = 0;
    GF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, h1.p, h1s.p, ids.p, scol.p, ncol, 0, 64, s));
    void *tmp = cub_scratch(ctx, tb);
    GF_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, h1.p, h1s.p, ids.p, scol.p, ncol, 0, 64, s));
    count_launch(17);
  }
  DevBuf<uint8_t> flags;
  flags.alloc(ctx, ncol);
  k_ut_class_flags<<<ugrid(ncol, B), B, 0, s>>>(h1s.p, h2.p, scol.p, ncol, flags.p);
  GF_LAUNCH_CHECK();
  DevBuf<uint32_t> cls_start;
  cls_start.alloc(ctx, ncol + 1);
  const int64_t ncls = ut_select_heads(ctx, flags.p, ncol, cls_start.p);
  flags.release(); h1.release(); h1s.release(); h2.release(); ids.release();
  {
    int32_t err = 0;
    t->flag.download(&err);
    GF_CUDA(cudaStreamSynchronize(s));
    GF_REQUIRE(err == 0, "uniform plan: corrupt structure (code " + std::to_string(err) + ")");
  }
  const int64_t max_classes = (int64_t)uenv_int("GFGPU_UT_MAXCLASSES", 60000);
  if (mode != 2 && (ncls > max_classes || ncls * 8 > ncol)) return false;
  std::vector<uint32_t> h_cls(ncls + 1);
  GF_CUDA(cudaMemcpyAsync(h_cls.data(), cls_start.p, ncls * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  GF_CUDA(cudaStreamSynchronize(s));
  h_cls[ncls] = (uint32_t)ncol;
  if (mode != 2) {  // columns in classes of at least 16 members
    int64_t covered = 0;
    for (int64_t c = 0; c < ncls; ++c)
      if (h_cls[c + 1] - h_cls[c] >= 16u) covered += h_cls[c + 1] - h_cls[c];
    if (covered * 10 < ncol * 9) return false;
  }
  // ---- leaders' descriptors -> host -> programs
  DevBuf<uint32_t> llen;
  llen.alloc(ctx, ncls);
  GF_CUDA(cudaMemcpyAsync(cls_start.p + ncls, &h_cls[ncls], sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  k_ut_leader_len<<<ugrid(ncls, B), B, 0, s>>>(cls_start.p, scol.p, dlen.p, ncls, llen.p);
  GF_LAUNCH_CHECK();
  std::vector<uint32_t> h_llen(ncls);
  llen.download(h_llen.data());
  GF_CUDA(cudaStreamSynchronize(s));
  uint32_t dw = 0;
  for (uint32_t v : h_llen) dw = std::max(dw, v);
  GF_REQUIRE((size_t)ncls * dw < (size_t(1) << 31), "uniform plan: descriptors too large");
  DevBuf<uint32_t> desc;
  desc.alloc(ctx, (size_t)ncls * dw);
  k_ut_leader_desc<<<ugrid(ncls, 64), 64, 0, s>>>(in, cls_start.p, scol.p, ncls, dw, desc.p);
  GF_LAUNCH_CHECK();
  std::vector<uint32_t> h_desc((size_t)ncls * dw), h_scol(ncol);
  desc.download(h_desc.data());
  scol.download(h_scol.data());
  GF_CUDA(cudaStreamSynchronize(s));
  dlen.release(); llen.release();

  // kernel variant: pairs per group (accumulator sets per lane) / warps per team / teams per CTA
  int kg = 3, tw = 4, teams = 3;
  if (const char *v = getenv("GFGPU_UT_VARIANT")) {
    if (sscanf(v, "%d,%d,%d", &kg, &tw, &teams) != 3) kg = 3, tw = 4, teams = 3;
  }
  GF_REQUIRE(kg >= 1 && kg <= 3 && (tw == 4 || tw == 8) && teams >= 2 && teams <= 4, "GFGPU_UT_VARIANT: kg,tw,teams out of range");
  const int img_bytes = std::max(8192, std::min(uenv_int("GFGPU_UT_IMG", teams == 4 ? 45056 : 51200), 200 * 1024));
  const uint32_t row_cap = (uint32_t)(img_bytes / 8 / 32);
  const int group_cap = std::max(1, uenv_int("GFGPU_UT_TASKCAP", 24));  // contributions per group of pairs
  std::vector<uint32_t> prog;
  std::vector<uplan::ClassPlan> cplan(ncls);
  uint32_t max_stride = 2;
  for (int64_t c = 0; c < ncls; ++c) {
    std::string err;
    const bool ok = uplan::build_class(h_desc.data() + (size_t)c * dw, h_llen[c], Q, nd, row_cap, group_cap, tw, kg, UT_SB, prog,
                                       cplan[c], err);
    if (!ok) {  // a column the uniform kernel's formats cannot hold: the term keeps the general tile kernel
      if (getenv("GFGPU_DEBUG")) fprintf(stderr, "[gfgpu] uniform tiles: not used (%s)\n", err.c_str());
      GF_REQUIRE(mode != 2, "uniform plan: " + err);
      return false;
    }
    for (const uplan::Sub &sb : cplan[c].subs) max_stride = std::max(max_stride, sb.rowstride);
    GF_REQUIRE(prog.size() < (size_t(1) << 31), "uniform plan: programs too large");
    GF_REQUIRE(cplan[c].m <= 4096, "uniform plan: column valence beyond the instruction format");
  }
  // ---- chunks of 32 class members, in the order of their first column; tiles = chunk x sub-range
  struct HChunk { uint32_t pos0, cls, nmem, first; };
  std::vector<HChunk> hch;
  hch.reserve(ncol / 32 + ncls + 1);
  for (int64_t c = 0; c < ncls; ++c)
    for (uint32_t p = h_cls[c]; p < h_cls[c + 1]; p += 32)
      hch.push_back({p, (uint32_t)c, std::min<uint32_t>(32u, h_cls[c + 1] - p), h_scol[p]});
  std::sort(hch.begin(), hch.end(), [](const HChunk &x, const HChunk &y) { return x.first < y.first; });
  std::vector<UChunk> chunks(hch.size());
  std::vector<UTile> tiles;
  std::vector<uint64_t> wsum;
  tiles.reserve(hch.size() * 2 + 1);
  uint64_t ldw = 0, wtot = 0, ntask = 0;
  for (size_t q = 0; q < hch.size(); ++q) {
    const uplan::ClassPlan &cp = cplan[hch[q].cls];
    GF_REQUIRE(ldw < (uint64_t(1) << 32), "uniform plan: lane data too large");
    chunks[q] = {hch[q].pos0, hch[q].cls, (uint32_t)ldw, hch[q].nmem};
    for (const uplan::Sub &sb : cp.subs) {
      UTile tl;
      tl.prog = sb.prog; tl.ld = (uint32_t)ldw;
      tl.nmem = hch[q].nmem; tl.pad = 0;
      tiles.push_back(tl);
      wtot += sb.weight;
      wsum.push_back(wtot);
      ntask += sb.ntasks;
    }
    ldw += 64 + 32 * (uint64_t)cp.m;
  }
  GF_REQUIRE(ntask < (uint64_t(1) << 32) && tiles.size() < (size_t(1) << 31), "uniform plan: too many tasks");
  const int64_t ntiles = (int64_t)tiles.size();
  const int grid = (int)std::min<int64_t>(ntiles, (int64_t)ctx->sm_count * std::max(1, uenv_int("GFGPU_UT_CTAS_PER_SM", 1)));
  std::vector<uint32_t> cta_t0(grid + 1, 0);
  {
    size_t tpos = 0;
    for (int c = 1; c < grid; ++c) {
      const uint64_t target = wtot * (uint64_t)c / (uint64_t)grid;
      while (tpos < (size_t)ntiles && wsum[tpos] <= target) ++tpos;
      cta_t0[c] = (uint32_t)std::max<size_t>(tpos, cta_t0[c - 1]);
    }
    cta_t0[grid] = (uint32_t)ntiles;
  }
  t->ru_tiles.alloc(ctx, tiles.size() * sizeof(UTile));
  GF_CUDA(cudaMemcpyAsync(t->ru_tiles.p, tiles.data(), tiles.size() * sizeof(UTile), cudaMemcpyHostToDevice, s));
  t->ru_cta.alloc(ctx, cta_t0.size());
  t->ru_cta.upload(cta_t0.data());
  prog.resize(prog.size() + 2 * UT_SB + 4, 0u);  // a stream copy always reads UT_SB units
  t->ru_prog.alloc(ctx, std::max<size_t>(prog.size(), 1));
  GF_CUDA(cudaMemcpyAsync(t->ru_prog.p, prog.data(), prog.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  DevBuf<uint8_t> dchunks;
  dchunks.alloc(ctx, chunks.size() * sizeof(UChunk));
  GF_CUDA(cudaMemcpyAsync(dchunks.p, chunks.data(), chunks.size() * sizeof(UChunk), cudaMemcpyHostToDevice, s));
  // ---- strip order of the elements, geometry table component-major in that order
  DevBuf<uint32_t> sp, epos;
  sp.alloc(ctx, ncol);
  epos.alloc(ctx, ne);
  k_ut_inverse<<<ugrid(ncol, B), B, 0, s>>>(scol.p, ncol, sp.p);
  GF_LAUNCH_CHECK();
  {
    DevBuf<uint64_t> k0, k1;
    DevBuf<uint32_t> v0, v1;
    k0.alloc(ctx, ne); k1.alloc(ctx, ne); v0.alloc(ctx, ne); v1.alloc(ctx, ne);
    k_ut_elem_key<<<ugrid(ne, B), B, 0, s>>>(t->edof_p(), nd, t->e0, ne, st.rdof.p, ncol, st.rstart.p, st.rsrc.p, sp.p,
                                           k0.p, v0.p, (int *)t->flag.p);
    GF_LAUNCH_CHECK();
    size_t tb = 0;
    GF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k0.p, k1.p, v0.p, v1.p, ne, 0, 64, s));
    void *tmp = cub_scratch(ctx, tb);
    GF_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, k0.p, k1.p, v0.p, v1.p, ne, 0, 64, s));
    count_launch(17);
    k_ut_scatter_pos<<<ugrid(ne, B), B, 0, s>>>(v1.p, ne, epos.p);
    GF_LAUNCH_CHECK();
    GF_CUDA(cudaStreamSynchronize(s));
  }
  sp.release();
  const int64_t nepad = (ne + 31) / 32 * 32;
  t->ru_nepad = nepad;
  GF_REQUIRE((size_t)nepad * GSZ * 8 < (size_t(1) << 32), "uniform plan: geometry table beyond 32-bit byte offsets");
  t->ru_eg.alloc(ctx, (size_t)nepad * GSZ);
  t->ru_eg.zero();
  k_ut_eg_blocked<<<ugrid(ne * GSZ, B), B, 0, s>>>(t->rc_eg.p, GSZ, ne, epos.p, t->ru_eg.p);
  GF_LAUNCH_CHECK();
  // ---- lane data (with the check that every member really has its leader's descriptor)
  t->ru_ld.alloc(ctx, std::max<uint64_t>(ldw, 1));
  k_ut_lane_data<<<ugrid((int64_t)chunks.size() * 32, B), B, 0, s>>>(in, (const UChunk *)dchunks.p, (int64_t)chunks.size(),
                                                                  scol.p, desc.p, dw, epos.p, GSZ, t->ru_ld.p, (int *)t->flag.p);
  GF_LAUNCH_CHECK();
  {
    int32_t err = 0;
    t->flag.download(&err);
    GF_CUDA(cudaStreamSynchronize(s));
    GF_REQUIRE(err == 0, "uniform plan failed (code " + std::to_string(err) + ")");
  }
  t->ru_grid = grid;
  t->ru_nbuf = teams;
  t->ru_kg = kg;
  t->ru_tw = tw;
  t->ru_imgcap = (int)(32 * max_stride);
  t->ru_ntiles = ntiles;
  t->ru_ntasks = (int64_t)ntask;
  if (getenv("GFGPU_DEBUG")) {
    int64_t big = 0;
    for (int64_t c = 0; c < ncls; ++c)
      if (h_cls[c + 1] - h_cls[c] >= 32u) ++big;
    fprintf(stderr,
            "[gfgpu] uniform tiles: %lld columns in %lld classes (%lld with >= 32 members), %lld chunks, %lld tiles, %llu tasks, "
            "programs %.1f KB, lane data %.1f MB, image %d B x %d, grid %d\n",
            (long long)ncol, (long long)ncls, (long long)big, (long long)chunks.size(), (long long)ntiles,
            (unsigned long long)ntask, prog.size() * 4 / 1024.0, ldw * 4 / 1048576.0, t->ru_imgcap * 8, teams, grid);
    fprintf(stderr, "[gfgpu] uniform tiles: variant kg %d, %d warps per team, %d teams\n", kg, tw, teams);
  }
  if (!t->halo) t->prel.release();
  t->rc_uni = true;
  return true;
}

template <int N, int Q, int ND, int RF, int KG, int TW, int TEAMS>
static void launch_utiles_t(gfgpu_term *t) {
  using C = TlCfg<N, RF>;
  constexpr int MTP = (C::MT + 1) & ~1;
  const double sign = t->alpha < 0 ? -1.0 : 1.0;
  UArgs a;
  a.tiles = (const UTile *)t->ru_tiles.p;
  a.cta_t0 = t->ru_cta.p;
  a.prog = (const uint2 *)t->ru_prog.p;
  a.ld = t->ru_ld.p;
  a.eg = t->ru_eg.p;
  a.Mtab = t->rc_M.p;
  a.sl = sign * t->par[0]; a.smu = sign * t->par[1];
  a.pr = t->pr.p;
  a.imgcap = t->ru_imgcap;
  a.dbg = uenv_int("GFGPU_UT_DBG", 0);
  const size_t smem = (size_t)((ND * ND * MTP + 15) & ~15) * 8 + (size_t)TEAMS * a.imgcap * 8 + (size_t)TEAMS * TW * UT_WS;
  GF_REQUIRE(smem <= 226 * 1024, "uniform tiles: image buffers too large for shared memory (GFGPU_UT_IMG / GFGPU_UT_VARIANT)");
  auto kern = k_utiles<N, Q, ND, RF, KG, TW, TEAMS>;
  GF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<t->ru_grid, TEAMS * TW * 32, smem, t->ctx->stream>>>(a);
  GF_LAUNCH_CHECK();
}

// kernel variants (GFGPU_UT_VARIANT=kg,tw,teams): the default everywhere; the experimental ones only for BASELINE config 3
template <int N, int Q, int ND, int RF>
static void launch_utiles(gfgpu_term *t) {
  const int kg = t->ru_kg, tw = t->ru_tw, teams = t->ru_nbuf;
  if (kg == 3 && tw == 4 && teams == 3) return launch_utiles_t<N, Q, ND, RF, 3, 4, 3>(t);
  if constexpr (N == 3 && Q == 3 && ND == 10 && RF == TF_ELAST) {
#define UT_VAR(K, W, T) if (kg == K && tw == W && teams == T) return launch_utiles_t<N, Q, ND, RF, K, W, T>(t);
    UT_VAR(1, 8, 3) UT_VAR(2, 8, 2) UT_VAR(1, 8, 2) UT_VAR(2, 4, 3) UT_VAR(1, 4, 3) UT_VAR(2, 8, 3) UT_VAR(3, 8, 2)
#undef UT_VAR
  }
  GF_REQUIRE(false, "uniform tiles: this GFGPU_UT_VARIANT is not compiled");
}

#define UT_CASE(NN, QQ, NDD, RFF)                     \
  if (N == NN && Q == QQ && nd == NDD && rf == RFF) { \
    launch_utiles<NN, QQ, NDD, RFF>(t);               \
    return;                                           \
  }

void uniform_tangent(gfgpu_term *t) {
  const int N = t->mesh->dim, nd = t->fem->nd, Q = t->fem->qdim, rf = tf_of(t->family);
  UT_CASE(3, 3, 10, TF_ELAST) UT_CASE(3, 3, 4, TF_ELAST) UT_CASE(3, 3, 20, TF_ELAST)
  UT_CASE(3, 1, 10, TF_LAPLACE) UT_CASE(3, 1, 4, TF_LAPLACE) UT_CASE(3, 1, 20, TF_LAPLACE)
  UT_CASE(3, 3, 10, TF_LAPLACE) UT_CASE(3, 3, 4, TF_LAPLACE) UT_CASE(3, 3, 20, TF_LAPLACE)
  UT_CASE(3, 1, 10, TF_MASS) UT_CASE(3, 1, 4, TF_MASS) UT_CASE(3, 1, 20, TF_MASS)
  UT_CASE(3, 3, 10, TF_MASS) UT_CASE(3, 3, 4, TF_MASS) UT_CASE(3, 3, 20, TF_MASS)
  UT_CASE(2, 2, 3, TF_ELAST) UT_CASE(2, 2, 6, TF_ELAST) UT_CASE(2, 2, 10, TF_ELAST)
  UT_CASE(2, 1, 3, TF_LAPLACE) UT_CASE(2, 1, 6, TF_LAPLACE) UT_CASE(2, 1, 10, TF_LAPLACE)
  UT_CASE(2, 2, 3, TF_LAPLACE) UT_CASE(2, 2, 6, TF_LAPLACE) UT_CASE(2, 2, 10, TF_LAPLACE)
  UT_CASE(2, 1, 3, TF_MASS) UT_CASE(2, 1, 6, TF_MASS) UT_CASE(2, 1, 10, TF_MASS)
  UT_CASE(2, 2, 3, TF_MASS) UT_CASE(2, 2, 6, TF_MASS) UT_CASE(2, 2, 10, TF_MASS)
  GF_REQUIRE(false, "no uniform tile kernel for this combination");
}

}  // namespace gf
