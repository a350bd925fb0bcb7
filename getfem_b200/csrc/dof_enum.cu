// dof_enum.cu -- device re-derivation of mesh_fem::enumerate_dof (getfem_mesh_fem.cc:320-446) for
// classical Lagrange PK / QK elements.
//
// The reference walks the elements in ascending convex id and, inside an element, the local dofs in
// ascending order; a Lagrange node seen for the first time receives the next number (advanced by Qdim,
// :397,403,418), a node already seen through a neighbour keeps its number (the reference identifies them
// geometrically with a kd-tree, :410-427).  Here coincident nodes are identified TOPOLOGICALLY: every
// local node gets a canonical key made of the global vertex ids of the mesh entity (vertex / edge /
// face) that carries it plus its lattice position on that entity, expressed in an element-independent
// frame; interior nodes are unique to their element.  Keys are grouped with a stable radix sort, the
// first (element, local) incidence of each group is its first touch, and ranking the first touches in
// walk order reproduces the sequential numbering.  dof = Qdim * rank.
#include <cub/cub.cuh>

#include "common.cuh"

namespace gf {

struct DofKey { uint64_t hi, lo; };

// lat[i*4 + d]: PK: barycentric multi-index (a0..aN) of local node i; QK: lattice index (ix,iy,iz,0)
template <bool QK>
__global__ void k_dof_keys(const int32_t *__restrict__ conn, int ng, int N, int k, int nd,
                           const int8_t *__restrict__ lat, int64_t ninc, uint64_t *__restrict__ khi,
                           uint64_t *__restrict__ klo, uint32_t *__restrict__ vals) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ninc; c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = c / nd;
    const int i = (int)(c % nd);
    const int32_t *cv = conn + e * ng;
    uint32_t v[3] = {0, 0, 0};
    uint32_t t[3] = {0, 0, 0};
    int cnt = 0;
    bool interior = false;
    if (!QK) {
      // support of the barycentric multi-index, sorted by global vertex id
      for (int d = 0; d <= N; ++d) {
        int a = lat[i * 4 + d];
        if (a > 0) {
          if (cnt == 3) { interior = true; break; }
          v[cnt] = (uint32_t)cv[d];
          t[cnt] = (uint32_t)a;
          ++cnt;
        }
      }
      if (cnt == N + 1) interior = true;
      if (!interior) {  // insertion sort by vertex id
        for (int a = 1; a < cnt; ++a)
          for (int b = a; b > 0 && v[b] < v[b - 1]; --b) {
            uint32_t x = v[b]; v[b] = v[b - 1]; v[b - 1] = x;
            x = t[b]; t[b] = t[b - 1]; t[b - 1] = x;
          }
      }
    } else {
      int idx[3] = {0, 0, 0}, fax[3];
      int nf = 0;
      for (int d = 0; d < N; ++d) {
        idx[d] = lat[i * 4 + d];
        if (idx[d] > 0 && idx[d] < k) fax[nf++] = d;
      }
      if (nf == N) interior = true;
      else {
        int base = 0;  // corner with the free axes at their 0 end
        for (int d = 0; d < N; ++d)
          if (idx[d] == k) base |= 1 << d;
        // origin = entity corner with the smallest global vertex id
        int best = base;
        uint32_t bestv = (uint32_t)cv[base];
        for (int m = 1; m < (1 << nf); ++m) {
          int cidx = base;
          for (int f = 0; f < nf; ++f)
            if (m & (1 << f)) cidx |= 1 << fax[f];
          uint32_t vv = (uint32_t)cv[cidx];
          if (vv < bestv) { bestv = vv; best = cidx; }
        }
        v[0] = bestv;
        cnt = 1;
        for (int f = 0; f < nf; ++f) {
          int d = fax[f];
          int nb = best ^ (1 << d);  // neighbour corner of the origin along this axis
          v[cnt] = (uint32_t)cv[nb];
          t[cnt] = (uint32_t)((best & (1 << d)) ? k - idx[d] : idx[d]);
          ++cnt;
        }
        if (cnt == 3 && v[2] < v[1]) {
          uint32_t x = v[1]; v[1] = v[2]; v[2] = x;
          x = t[1]; t[1] = t[2]; t[2] = x;
        }
      }
    }
    uint64_t hi, lo;
    if (interior) {
      hi = (uint64_t(3) << 62) | (uint64_t)e;
      lo = (uint64_t)i;
    } else {
      hi = ((uint64_t)(cnt - 1) << 62) | ((uint64_t)v[0] << 31) | (uint64_t)v[1];
      lo = ((uint64_t)v[2] << 32) | ((uint64_t)t[0] << 16) | ((uint64_t)t[1] << 8) | (uint64_t)t[2];
    }
    khi[c] = hi;
    klo[c] = lo;
    vals[c] = (uint32_t)c;
  }
}

__global__ void k_gather_u64(const uint64_t *__restrict__ src, const uint32_t *__restrict__ idx, int64_t n,
                             uint64_t *__restrict__ dst) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x)
    dst[s] = src[idx[s]];
}

// head of a group in sorted order; a head marks its incidence as a first touch
__global__ void k_dof_heads(const uint64_t *__restrict__ shi, const uint64_t *__restrict__ slo,
                            const uint32_t *__restrict__ svals, int64_t n, uint32_t *__restrict__ headflag,
                            uint32_t *__restrict__ isfirst) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
    bool head = s == 0 || shi[s] != shi[s - 1] || slo[s] != slo[s - 1];
    headflag[s] = head ? 1u : 0u;
    if (head) isfirst[svals[s]] = 1u;
  }
}

// group index g(s) = inclusive-scan(headflag)-1; firstc[g] = svals at the head
__global__ void k_dof_group_first(const uint32_t *__restrict__ headflag, const uint32_t *__restrict__ gidx,
                                  const uint32_t *__restrict__ svals, int64_t n, uint32_t *__restrict__ firstc) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x)
    if (headflag[s]) firstc[gidx[s] - 1] = svals[s];
}

__global__ void k_dof_assign(const uint32_t *__restrict__ gidx, const uint32_t *__restrict__ svals,
                             const uint32_t *__restrict__ firstc, const uint32_t *__restrict__ rank, int Q, int64_t n,
                             int32_t *__restrict__ edof) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x)
    edof[svals[s]] = (int32_t)(Q * rank[firstc[gidx[s] - 1]]);
}

static int grid_for(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  return (int)(g < 148 * 16 ? (g < 1 ? 1 : g) : 148 * 16);
}

// returns ndof
int64_t enumerate_dof(gfgpu_ctx *ctx, const int32_t *conn, int64_t ne, int ng, int N, bool qk, int k, int Q, int nd,
                      const int8_t *lat_host, int32_t *edof_dev) {
  const int64_t n = ne * nd;
  GF_REQUIRE(n < (int64_t(1) << 32), "too many (element, local dof) incidences");
  GF_REQUIRE(k >= 1 && k < 128, "bad fem degree");
  cudaStream_t s = ctx->stream;
  const int B = 256;
  DevBuf<int8_t> lat;
  lat.alloc(ctx, (size_t)nd * 4);
  lat.upload(lat_host);
  DevBuf<uint64_t> khi, klo, tmp64a, tmp64b;
  DevBuf<uint32_t> v0, v1, v2;
  khi.alloc(ctx, n); klo.alloc(ctx, n); tmp64a.alloc(ctx, n); tmp64b.alloc(ctx, n);
  v0.alloc(ctx, n); v1.alloc(ctx, n); v2.alloc(ctx, n);
  if (qk) k_dof_keys<true><<<grid_for(n, B), B, 0, s>>>(conn, ng, N, k, nd, lat.p, n, khi.p, klo.p, v0.p);
  else k_dof_keys<false><<<grid_for(n, B), B, 0, s>>>(conn, ng, N, k, nd, lat.p, n, khi.p, klo.p, v0.p);
  GF_LAUNCH_CHECK();
  // LSD: stable sort by lo, then by hi
  size_t tb = 0;
  GF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, klo.p, tmp64a.p, v0.p, v1.p, n, 0, 64, s));
  void *tmp = cub_scratch(ctx, tb);
  GF_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, klo.p, tmp64a.p, v0.p, v1.p, n, 0, 64, s));  // tmp64a = sorted lo
  k_gather_u64<<<grid_for(n, B), B, 0, s>>>(khi.p, v1.p, n, tmp64b.p);                          // hi in lo-order
  GF_LAUNCH_CHECK();
  GF_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, tmp64b.p, khi.p, v1.p, v2.p, n, 0, 64, s));  // khi = sorted hi, v2 = c
  k_gather_u64<<<grid_for(n, B), B, 0, s>>>(klo.p, v2.p, n, tmp64a.p);                          // lo in final order
  GF_LAUNCH_CHECK();
  count_launch(18);
  // heads / first touches
  DevBuf<uint32_t> headflag, isfirst, gidx, rank, firstc;
  headflag.alloc(ctx, n); isfirst.alloc(ctx, n); gidx.alloc(ctx, n); rank.alloc(ctx, n);
  isfirst.zero();
  k_dof_heads<<<grid_for(n, B), B, 0, s>>>(khi.p, tmp64a.p, v2.p, n, headflag.p, isfirst.p);
  GF_LAUNCH_CHECK();
  size_t tb2 = 0;
  GF_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb2, headflag.p, gidx.p, n, s));
  tmp = cub_scratch(ctx, tb2);
  GF_CUDA(cub::DeviceScan::InclusiveSum(tmp, tb2, headflag.p, gidx.p, n, s));
  GF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb2, isfirst.p, rank.p, n, s));
  tmp = cub_scratch(ctx, tb2);
  GF_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb2, isfirst.p, rank.p, n, s));
  count_launch(4);
  uint32_t ngroups = 0;
  GF_CUDA(cudaMemcpyAsync(&ngroups, gidx.p + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  GF_CUDA(cudaStreamSynchronize(s));
  firstc.alloc(ctx, ngroups);
  k_dof_group_first<<<grid_for(n, B), B, 0, s>>>(headflag.p, gidx.p, v2.p, n, firstc.p);
  GF_LAUNCH_CHECK();
  k_dof_assign<<<grid_for(n, B), B, 0, s>>>(gidx.p, v2.p, firstc.p, rank.p, Q, n, edof_dev);
  GF_LAUNCH_CHECK();
  GF_CUDA(cudaStreamSynchronize(s));
  return (int64_t)ngroups * Q;
}

}  // namespace gf
