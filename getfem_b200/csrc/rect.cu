// rect.cu -- coupled bilinear terms (Test on one fem, Test2 on another): the off-diagonal blocks of mixed formulations.
//
// GetFEM assembles an order-2 tree whose two test functions belong to DIFFERENT variables into the block
// (interval of Test's variable) x (interval of Test2's variable) of the workspace matrix, element matrix by element matrix
// through the same add_elem_matrix / drop rule as every other tree (C&E.cc:4853-4936, 5380-5402; the incompressibility
// bricks add "-p*Div_Test_u - Test_p*Div_u", getfem_models.cc add_linear_incompressibility).  Here:
//   element kernel : one CTA walks elements; geometry per Gauss point (affine or multilinear), the sr x sc element matrix in
//                    shared memory, the drop rule |v| > 1e-14 max|E| on the whole element matrix, thresholded values and
//                    keep flags to the stage.  Family DIV_PRESSURE: E[(i,a), j] = coef * sum_q w_q J psi_j (B ghat_i)_a.
//   pattern        : the kept contributions radix-sorted (CUB, stable) by (column dof, row dof): an entry exists iff one of
//                    its contributions is kept, like rsvector; inside an entry the contributions stay in ascending element
//                    order.  The same entries sorted by (row, column) give the CSC of the TRANSPOSED block (the tree with
//                    the test functions swapped).
//   values         : thread per entry, ordered sum of its contributions; bitwise reproducible, no atomics.
//   products       : y = B x and y = B^T x (residuals of the two variables: R_u = B p, R_p = B^T u), warp per output row.
#include <cub/cub.cuh>

#include "elem_kernel.cuh"

namespace gf {
void set_last_error(const std::string &s);

struct RectArgs {
  const double *x, *y, *z;
  const int32_t *conn, *edr, *edc;
  const double *w, *gt_grad, *gphi_r, *phi_r, *phi_c;
  const int8_t *face;      // face of every item (region of faces), or nullptr; the tables are then the face tables, face after face
  const double *fnormal;   // nf x 3 reference normals
  int family;
  int ng, nq, ndr, ndc, qr, qc;
  int64_t ne;
  double coef;
  double *stage;
  uint8_t *keep;
};

template <int N>
__global__ void __launch_bounds__(128) k_rect_div(const RectArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int sr = a.ndr * a.qr, sc = a.ndc * a.qc;
  double *sG = sm;                    // N x ng coordinates
  double *sGeo = sG + N * a.ng;       // nq x (N*N + 1 + N)
  double *sE = sGeo + a.nq * (N * N + 1 + N);  // sr x sc
  __shared__ double sRed[4];
  constexpr int GEO = N * N + 1 + N;
  const int tid = threadIdx.x;
  for (int64_t e = blockIdx.x; e < a.ne; e += gridDim.x) {
    // tables of this item: the volume points, or the nq points of its face (C&E.cc:8827-8848)
    const int fc = a.face ? a.face[e] : 0;
    const double *tw = a.w + (size_t)fc * a.nq, *tgt = a.gt_grad + (size_t)fc * a.nq * a.ng * N;
    const double *tgr = a.gphi_r + (size_t)fc * a.nq * a.ndr * N, *tpr = a.phi_r + (size_t)fc * a.nq * a.ndr;
    const double *tpc = a.phi_c + (size_t)fc * a.nq * a.ndc;
    const double *nref = a.face ? a.fnormal + fc * 3 : nullptr;
    for (int k = tid; k < N * a.ng; k += blockDim.x) {
      const int i = k / N, d = k % N;
      const int32_t p = a.conn[e * a.ng + i];
      sG[d + N * i] = (d == 0 ? a.x : d == 1 ? a.y : a.z)[p];
    }
    __syncthreads();
    for (int q = tid; q < a.nq; q += blockDim.x) geometry<N>(sG, tgt + (size_t)q * a.ng * N, a.ng, sGeo + q * GEO, nref);
    __syncthreads();
    double vmax = 0.0;
    for (int k = tid; k < sr * sc; k += blockDim.x) {  // k = row + sr * column, row = i * qr + a
      const int row = k % sr, col = k / sr, i = row / a.qr, c = row % a.qr, j = col / a.qc, d = col % a.qc;
      double s = 0.0;
      if (a.family == GFGPU_RECT_MASS) {  // E[(i,c), (j,d)] = delta_cd sum_q w_q J phi_i psi_j
        if (c == d)
          for (int q = 0; q < a.nq; ++q) {
            const double wq = tw[q];
            if (wq == 0.0) continue;
            s += (wq * sGeo[q * GEO + N * N]) * tpr[(size_t)q * a.ndr + i] * tpc[(size_t)q * a.ndc + j];
          }
      } else {
        for (int q = 0; q < a.nq; ++q) {
          const double wq = tw[q];
          if (wq == 0.0) continue;  // zero-weight points are skipped (C&E.cc:8852)
          const double *geo = sGeo + q * GEO;
          const double *g = tgr + ((size_t)q * a.ndr + i) * N;
          double dv = 0.0;  // (B ghat_i)_c = d phi_i / d x_c
#pragma unroll
          for (int p = 0; p < N; ++p) dv += geo[c + N * p] * g[p];
          s += (wq * geo[N * N]) * tpc[(size_t)q * a.ndc + j] * dv;
        }
      }
      s *= a.coef;
      sE[k] = s;
      vmax = fmax(vmax, fabs(s));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, off));
    if ((tid & 31) == 0) sRed[tid >> 5] = vmax;
    __syncthreads();
    vmax = fmax(fmax(sRed[0], sRed[1]), fmax(sRed[2], sRed[3]));
    const double thr = vmax * 1e-14;
    for (int k = tid; k < sr * sc; k += blockDim.x) {
      const double v = sE[k];
      const bool kp = (vmax != 0.0) && (fabs(v) > thr);
      a.stage[(size_t)e * sr * sc + k] = kp ? v : 0.0;
      a.keep[(size_t)e * sr * sc + k] = kp ? 1 : 0;
    }
    __syncthreads();
  }
}

// key of contribution c = (element, row, column): column dof * nrows + row dof, or ~0 when the drop rule discards it
__global__ void k_rect_keys(const int32_t *__restrict__ edr, const int32_t *__restrict__ edc, const uint8_t *__restrict__ keep,
                            int ndr, int qr, int ndc, int qc, int64_t nct, int64_t nrows, unsigned long long *__restrict__ key,
                            uint32_t *__restrict__ val) {
  const int sr = ndr * qr, sc = ndc * qc;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nct; c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = c / (sr * sc);
    const int k = (int)(c - e * (sr * sc)), row = k % sr, col = k / sr;
    const int64_t rd = edr[e * ndr + row / qr] + row % qr, cd = edc[e * ndc + col / qc] + col % qc;
    key[c] = keep[c] ? (unsigned long long)cd * (unsigned long long)nrows + (unsigned long long)rd : ~0ull;
    val[c] = (uint32_t)c;
  }
}
__global__ void k_rect_heads(const unsigned long long *__restrict__ key, int64_t n, uint32_t *__restrict__ head) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    head[k] = (key[k] != ~0ull && (k == 0 || key[k] != key[k - 1])) ? 1u : 0u;
}
// entry id of every sorted contribution = exclusive scan of the heads (+ head - 1); heads record the segment starts
__global__ void k_rect_entries(const unsigned long long *__restrict__ key, const uint32_t *__restrict__ head,
                               const uint32_t *__restrict__ scan, int64_t n, int64_t nrows, uint32_t *__restrict__ seg,
                               int32_t *__restrict__ ir, unsigned long long *__restrict__ colcnt,
                               unsigned long long *__restrict__ rowcnt) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    if (!head[k]) continue;
    const uint32_t id = scan[k];
    seg[id] = (uint32_t)k;
    const unsigned long long cd = key[k] / (unsigned long long)nrows, rd = key[k] - cd * (unsigned long long)nrows;
    ir[id] = (int32_t)rd;
    atomicAdd(&colcnt[cd], 1ull);  // integer counts: order independent
    atomicAdd(&rowcnt[rd], 1ull);
  }
}
__global__ void k_rect_tkeys(const int64_t *__restrict__ jc, const int32_t *__restrict__ ir, int64_t ncols, int64_t ncols_mult,
                             unsigned long long *__restrict__ key, uint32_t *__restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = w0; c < ncols; c += nw)
    for (int64_t k = jc[c] + lane; k < jc[c + 1]; k += 32) {
      key[k] = (unsigned long long)ir[k] * (unsigned long long)ncols_mult + (unsigned long long)c;
      val[k] = (uint32_t)k;
    }
}
__global__ void k_rect_tcols(const unsigned long long *__restrict__ key, int64_t nnz, int64_t ncols_mult, int32_t *__restrict__ irt) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x)
    irt[k] = (int32_t)(key[k] % (unsigned long long)ncols_mult);
}
__global__ void k_rect_gather(const uint32_t *__restrict__ seg, const uint32_t *__restrict__ perm, const double *__restrict__ stage,
                              int64_t nnz, double alpha, double *__restrict__ pr) {
  for (int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; id < nnz; id += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (uint32_t k = seg[id], e = seg[id + 1]; k < e; ++k) s += stage[perm[k]];  // ascending element id
    pr[id] = alpha * s;
  }
}
__global__ void k_rect_tvals(const uint32_t *__restrict__ tperm, const double *__restrict__ pr, int64_t nnz, double *__restrict__ prt) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) prt[k] = pr[tperm[k]];
}
// y[c] = beta y[c] + alpha sum_k pr[k] x[ir[k]] over column c, warp per column, fixed lane tree
__global__ void k_rect_colmult(const int64_t *__restrict__ jc, const int32_t *__restrict__ ir, const double *__restrict__ pr,
                               int64_t ncols, const double *__restrict__ x, double alpha, double beta, double *__restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = w0; c < ncols; c += nw) {
    double s = 0.0;
    for (int64_t k = jc[c] + lane; k < jc[c + 1]; k += 32) s += pr[k] * x[ir[k]];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[c] = (beta == 0.0 ? 0.0 : beta * y[c]) + alpha * s;
  }
}

struct IsKept {
  __host__ __device__ uint32_t operator()(unsigned long long k) const { return k != ~0ull ? 1u : 0u; }
};

static inline int rgrid(int64_t n, int block) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + block - 1) / block, 148 * 32));
}

static void scan_u32(gfgpu_ctx *ctx, const uint32_t *in, uint32_t *out, int64_t n) {
  size_t tb = 0;
  GF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, ctx->stream));
  void *tmp = cub_scratch(ctx, tb);
  GF_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, in, out, n, ctx->stream));
  count_launch(2);
}
static void scan_i64(gfgpu_ctx *ctx, const int64_t *in, int64_t *out, int64_t n) {
  size_t tb = 0;
  GF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, ctx->stream));
  void *tmp = cub_scratch(ctx, tb);
  GF_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, in, out, n, ctx->stream));
  count_launch(2);
}

static void rect_elements(gfgpu_rect *r) {
  gfgpu_ctx *ctx = r->ctx;
  RectArgs a;
  const int64_t np = r->mesh->npts;
  a.x = r->mesh->xyz.p; a.y = a.x + np; a.z = a.y + np;
  a.conn = r->region ? r->r_conn.p : r->mesh->conn.p;
  a.edr = r->region ? r->r_edr.p : r->fr->edof.p;
  a.edc = r->region ? r->r_edc.p : r->fc->edof.p;
  a.family = r->family;
  if (r->region_faces) {
    a.w = r->tr->fw.p; a.gt_grad = r->tr->fgt_grad.p; a.gphi_r = r->tr->fgphi.p; a.phi_r = r->tr->fphi.p; a.phi_c = r->tc->fphi.p;
    a.face = r->r_face.p; a.fnormal = r->tr->fnormal.p;
  } else {
    a.w = r->tr->w.p; a.gt_grad = r->tr->gt_grad.p; a.gphi_r = r->tr->gphi.p; a.phi_r = r->tr->phi.p; a.phi_c = r->tc->phi.p;
    a.face = nullptr; a.fnormal = nullptr;
  }
  a.ng = r->mesh->ng; a.nq = r->region_faces ? r->tr->nqf : r->tr->nq; a.ndr = r->fr->nd; a.ndc = r->fc->nd; a.qr = r->fr->qdim; a.qc = r->fc->qdim;
  a.ne = r->ne; a.coef = r->coef;
  a.stage = r->stage.p; a.keep = r->keep.p;
  const int N = r->mesh->dim;
  const size_t smem = ((size_t)N * a.ng + (size_t)a.nq * (N * N + 1 + N) + (size_t)r->sr * r->sc + 2) * 8;  // (a.nq: volume or face points)
  GF_REQUIRE(smem <= 200 * 1024, "coupled term: element matrix too large for shared memory");
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(r->ne, (int64_t)ctx->sm_count * 8));
  if (N == 2) {
    GF_CUDA(cudaFuncSetAttribute(k_rect_div<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_rect_div<2><<<grid, 128, smem, ctx->stream>>>(a);
  } else {
    GF_CUDA(cudaFuncSetAttribute(k_rect_div<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_rect_div<3><<<grid, 128, smem, ctx->stream>>>(a);
  }
  GF_LAUNCH_CHECK();
}

static void rect_pattern(gfgpu_rect *r) {
  gfgpu_ctx *ctx = r->ctx;
  cudaStream_t s = ctx->stream;
  const int B = 256;
  const int64_t nct = r->ne * r->sr * r->sc;
  DevBuf<unsigned long long> key0, key1;
  DevBuf<uint32_t> val0, head, scan;
  key0.alloc(ctx, nct); key1.alloc(ctx, nct); val0.alloc(ctx, nct); head.alloc(ctx, nct); scan.alloc(ctx, nct);
  r->perm.alloc(ctx, nct);
  k_rect_keys<<<rgrid(nct, B), B, 0, s>>>(r->region ? r->r_edr.p : r->fr->edof.p, r->region ? r->r_edc.p : r->fc->edof.p, r->keep.p, r->fr->nd, r->fr->qdim, r->fc->nd, r->fc->qdim, nct,
                                        r->nrows, key0.p, val0.p);
  GF_LAUNCH_CHECK();
  size_t tb = 0;
  GF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, key0.p, key1.p, val0.p, r->perm.p, nct, 0, 64, s));
  void *tmp = cub_scratch(ctx, tb);
  GF_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, key0.p, key1.p, val0.p, r->perm.p, nct, 0, 64, s));
  count_launch(2);
  k_rect_heads<<<rgrid(nct, B), B, 0, s>>>(key1.p, nct, head.p);
  GF_LAUNCH_CHECK();
  scan_u32(ctx, head.p, scan.p, nct);
  uint32_t last_scan = 0, last_head = 0;
  GF_CUDA(cudaMemcpyAsync(&last_scan, scan.p + nct - 1, 4, cudaMemcpyDeviceToHost, s));
  GF_CUDA(cudaMemcpyAsync(&last_head, head.p + nct - 1, 4, cudaMemcpyDeviceToHost, s));
  GF_CUDA(cudaStreamSynchronize(s));
  r->nnz = (int64_t)last_scan + last_head;
  r->seg.alloc(ctx, r->nnz + 1);
  r->ir.alloc(ctx, std::max<int64_t>(r->nnz, 1));
  r->pr.alloc(ctx, std::max<int64_t>(r->nnz, 1));
  DevBuf<int64_t> colcnt, rowcnt;
  colcnt.alloc(ctx, r->ncols + 1); rowcnt.alloc(ctx, r->nrows + 1);
  colcnt.zero(); rowcnt.zero();
  k_rect_entries<<<rgrid(nct, B), B, 0, s>>>(key1.p, head.p, scan.p, nct, r->nrows, r->seg.p, r->ir.p,
                                           (unsigned long long *)colcnt.p, (unsigned long long *)rowcnt.p);
  GF_LAUNCH_CHECK();
  // the kept contributions are the first nkept of the sorted list (discarded ones carry the largest key)
  {
    cub::TransformInputIterator<uint32_t, IsKept, unsigned long long *> it(key1.p, IsKept());
    size_t tb2 = 0;
    DevBuf<uint32_t> total;
    total.alloc(ctx, 1);
    GF_CUDA(cub::DeviceReduce::Sum(nullptr, tb2, it, total.p, nct, s));
    void *tmp2 = cub_scratch(ctx, tb2);
    GF_CUDA(cub::DeviceReduce::Sum(tmp2, tb2, it, total.p, nct, s));
    count_launch(1);
    uint32_t nk = 0;
    total.download(&nk);
    GF_CUDA(cudaStreamSynchronize(s));
    r->nkept = nk;
    GF_CUDA(cudaMemcpyAsync(r->seg.p + r->nnz, &nk, 4, cudaMemcpyHostToDevice, s));
    GF_CUDA(cudaStreamSynchronize(s));
  }
  r->jc.alloc(ctx, r->ncols + 1); r->jct.alloc(ctx, r->nrows + 1);
  scan_i64(ctx, colcnt.p, r->jc.p, r->ncols + 1);
  scan_i64(ctx, rowcnt.p, r->jct.p, r->nrows + 1);
  // transposed block: the same entries sorted by (row, column)
  if (r->nnz) {
    DevBuf<unsigned long long> tk0, tk1;
    DevBuf<uint32_t> tv0;
    tk0.alloc(ctx, r->nnz); tk1.alloc(ctx, r->nnz); tv0.alloc(ctx, r->nnz);
    r->tperm.alloc(ctx, r->nnz); r->irt.alloc(ctx, r->nnz); r->prt.alloc(ctx, r->nnz);
    k_rect_tkeys<<<rgrid(r->ncols * 32, B), B, 0, s>>>(r->jc.p, r->ir.p, r->ncols, r->ncols, tk0.p, tv0.p);
    GF_LAUNCH_CHECK();
    size_t tb3 = 0;
    GF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb3, tk0.p, tk1.p, tv0.p, r->tperm.p, r->nnz, 0, 64, s));
    void *tmp3 = cub_scratch(ctx, tb3);
    GF_CUDA(cub::DeviceRadixSort::SortPairs(tmp3, tb3, tk0.p, tk1.p, tv0.p, r->tperm.p, r->nnz, 0, 64, s));
    count_launch(2);
    k_rect_tcols<<<rgrid(r->nnz, B), B, 0, s>>>(tk1.p, r->nnz, r->ncols, r->irt.p);
    GF_LAUNCH_CHECK();
  }
  GF_CUDA(cudaStreamSynchronize(s));
  r->pat_valid = true;
  r->generation++;
}

void rect_assemble(gfgpu_rect *r) {
  gfgpu_ctx *ctx = r->ctx;
  GF_CUDA(cudaSetDevice(ctx->device));
  const int64_t nct = r->ne * r->sr * r->sc;
  GF_REQUIRE(nct > 0 && nct < (int64_t(1) << 32) - 1, "coupled term: empty, or more than 2^32 element entries");
  if (r->stage.n != (size_t)nct) { r->stage.alloc(ctx, nct); r->keep.alloc(ctx, nct); }
  rect_elements(r);
  if (!r->pat_valid) rect_pattern(r);  // a linear constant-coefficient form: the keep flags cannot move afterwards
  if (!r->nnz) return;
  const int B = 256;
  k_rect_gather<<<rgrid(r->nnz, B), B, 0, ctx->stream>>>(r->seg.p, r->perm.p, r->stage.p, r->nnz, r->alpha, r->pr.p);
  GF_LAUNCH_CHECK();
  k_rect_tvals<<<rgrid(r->nnz, B), B, 0, ctx->stream>>>(r->tperm.p, r->pr.p, r->nnz, r->prt.p);
  GF_LAUNCH_CHECK();
}

}  // namespace gf

#define GFR_BEGIN try {
#define GFR_END                                  \
  return 0;                                      \
  }                                              \
  catch (const std::exception &ex) {             \
    gf::set_last_error(ex.what());               \
    return 1;                                    \
  }                                              \
  catch (...) {                                  \
    gf::set_last_error("unknown error");         \
    return 1;                                    \
  }

extern "C" {

int gfgpu_rect_create(gfgpu_ctx *ctx, gfgpu_mesh *mesh, gfgpu_fem *fem_rows, gfgpu_tables *tab_rows, gfgpu_fem *fem_cols,
                      gfgpu_tables *tab_cols, int family, double coef, double alpha, gfgpu_rect **out) {
  GFR_BEGIN
  GF_REQUIRE(ctx && mesh && fem_rows && fem_cols && tab_rows && tab_cols && out, "null argument");
  GF_REQUIRE(fem_rows->mesh == mesh && fem_cols->mesh == mesh, "both fems must live on the term's mesh");
  GF_REQUIRE(tab_rows->nq == tab_cols->nq && tab_rows->ng == tab_cols->ng && tab_rows->dim == mesh->dim,
             "the two table sets must share the quadrature points");
  GF_REQUIRE(tab_rows->nd == fem_rows->nd && tab_cols->nd == fem_cols->nd, "tables do not match the fems");
  GF_REQUIRE(family == GFGPU_RECT_DIV_PRESSURE || family == GFGPU_RECT_MASS, "unknown coupled family");
  if (family == GFGPU_RECT_DIV_PRESSURE)
    GF_REQUIRE(fem_rows->qdim == mesh->dim && fem_cols->qdim == 1, "div-pressure coupling: vector rows (qdim = mesh dimension), scalar columns");
  else
    GF_REQUIRE(fem_rows->qdim == fem_cols->qdim, "coupled mass term: both fems must have the same qdim");
  GF_REQUIRE(mesh->dim == 2 || mesh->dim == 3, "coupled terms: 2D and 3D meshes");
  std::unique_ptr<gfgpu_rect> r(new gfgpu_rect);
  r->ctx = ctx; r->mesh = mesh; r->fr = fem_rows; r->fc = fem_cols; r->tr = tab_rows; r->tc = tab_cols;
  r->family = family; r->coef = coef; r->alpha = alpha;
  r->nrows = fem_rows->ndof; r->ncols = fem_cols->ndof; r->ne = mesh->ne;
  r->sr = fem_rows->nd * fem_rows->qdim; r->sc = fem_cols->nd * fem_cols->qdim;
  *out = r.release();
  GFR_END
}

int gfgpu_rect_set_region(gfgpu_rect *r, int64_t n_items, const int32_t *cv, const int32_t *face) {
  GFR_BEGIN
  GF_REQUIRE(r, "null term");
  GF_REQUIRE(n_items >= 0 && (cv || n_items == 0), "bad region");
  gfgpu_ctx *ctx = r->ctx;
  GF_CUDA(cudaSetDevice(ctx->device));
  r->pat_valid = false;  // another set of elements: another pattern
  r->stage.release(); r->keep.release();
  if (!cv) {
    r->region = r->region_faces = false;
    r->r_conn.release(); r->r_edr.release(); r->r_edc.release(); r->r_face.release();
    r->ne = r->mesh->ne;
    return 0;
  }
  const int ng = r->mesh->ng, ndr = r->fr->nd, ndc = r->fc->nd;
  int64_t nfaces = 0;
  for (int64_t k = 0; k < n_items; ++k) {
    GF_REQUIRE(cv[k] >= 0 && cv[k] < r->mesh->ne, "region refers to a convex outside the mesh");
    GF_REQUIRE(k == 0 || cv[k] > cv[k - 1] || (cv[k] == cv[k - 1] && face && face[k] > face[k - 1]),
               "region items must come in mr_visitor order (ascending convex, then face)");
    if (face && face[k] >= 0) ++nfaces;
  }
  GF_REQUIRE(nfaces == 0 || nfaces == n_items, "a region must hold either convexes or faces, not both");
  if (nfaces) {
    GF_REQUIRE(r->tr->nf > 0 && r->tc->nf == r->tr->nf && r->tc->nqf == r->tr->nqf, "a region of faces needs gfgpu_tables_set_faces on both table sets");
    for (int64_t k = 0; k < n_items; ++k) GF_REQUIRE(face[k] < r->tr->nf, "face number outside the reference element");
  }
  std::vector<int32_t> hc((size_t)r->mesh->ne * ng), hr((size_t)r->mesh->ne * ndr), hcol((size_t)r->mesh->ne * ndc);
  r->mesh->conn.download(hc.data());
  r->fr->edof.download(hr.data());
  r->fc->edof.download(hcol.data());
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<int32_t> rc((size_t)n_items * ng), rr((size_t)n_items * ndr), rcc((size_t)n_items * ndc);
  std::vector<int8_t> rf((size_t)n_items);
  for (int64_t k = 0; k < n_items; ++k) {
    std::copy(hc.begin() + (size_t)cv[k] * ng, hc.begin() + (size_t)(cv[k] + 1) * ng, rc.begin() + (size_t)k * ng);
    std::copy(hr.begin() + (size_t)cv[k] * ndr, hr.begin() + (size_t)(cv[k] + 1) * ndr, rr.begin() + (size_t)k * ndr);
    std::copy(hcol.begin() + (size_t)cv[k] * ndc, hcol.begin() + (size_t)(cv[k] + 1) * ndc, rcc.begin() + (size_t)k * ndc);
    rf[k] = (int8_t)(nfaces ? face[k] : -1);
  }
  r->r_conn.alloc(ctx, rc.size()); r->r_conn.upload(rc.data());
  r->r_edr.alloc(ctx, rr.size()); r->r_edr.upload(rr.data());
  r->r_edc.alloc(ctx, rcc.size()); r->r_edc.upload(rcc.data());
  if (nfaces) { r->r_face.alloc(ctx, rf.size()); r->r_face.upload(rf.data()); } else r->r_face.release();
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  r->region = true;
  r->region_faces = nfaces > 0;
  r->ne = n_items;
  GFR_END
}

int gfgpu_rect_destroy(gfgpu_rect *r) {
  GFR_BEGIN
  if (r) {
    cudaSetDevice(r->ctx->device);
    cudaStreamSynchronize(r->ctx->stream);
  }
  delete r;
  GFR_END
}

int gfgpu_rect_assemble_dev(gfgpu_rect *r) {
  GFR_BEGIN
  GF_REQUIRE(r, "null term");
  gf::rect_assemble(r);
  GFR_END
}

int64_t gfgpu_rect_nnz(gfgpu_rect *r) { return (r && r->pat_valid) ? r->nnz : -1; }

int gfgpu_rect_export_csc_host(gfgpu_rect *r, int transposed, int64_t *jc, int32_t *ir, double *pr) {
  GFR_BEGIN
  GF_REQUIRE(r && r->pat_valid, "no assembled coupled term");
  GF_CUDA(cudaSetDevice(r->ctx->device));
  if (jc) (transposed ? r->jct : r->jc).download(jc);
  if (r->nnz) {
    if (ir) (transposed ? r->irt : r->ir).download(ir);
    if (pr) (transposed ? r->prt : r->pr).download(pr);
  }
  GF_CUDA(cudaStreamSynchronize(r->ctx->stream));
  GFR_END
}

int gfgpu_rect_mult_dev(gfgpu_rect *r, int transposed, double alpha, const double *x_dev, double beta, double *y_dev) {
  GFR_BEGIN
  GF_REQUIRE(r && r->pat_valid && x_dev && y_dev, "no assembled coupled term");
  GF_CUDA(cudaSetDevice(r->ctx->device));
  const int B = 256;
  // y = B x: the rows of B are the columns of its transpose; y = B^T x: the columns of B
  const int64_t nout = transposed ? r->ncols : r->nrows;
  if (!r->nnz) {
    if (beta == 0.0) GF_CUDA(cudaMemsetAsync(y_dev, 0, nout * sizeof(double), r->ctx->stream));
    return 0;
  }
  if (transposed)
    gf::k_rect_colmult<<<gf::rgrid(nout * 32, B), B, 0, r->ctx->stream>>>(r->jc.p, r->ir.p, r->pr.p, nout, x_dev, alpha, beta, y_dev);
  else
    gf::k_rect_colmult<<<gf::rgrid(nout * 32, B), B, 0, r->ctx->stream>>>(r->jct.p, r->irt.p, r->prt.p, nout, x_dev, alpha, beta, y_dev);
  GF_LAUNCH_CHECK();
  GFR_END
}

int gfgpu_rect_mult_host(gfgpu_rect *r, int transposed, double alpha, const double *x_host, double beta, double *y_host) {
  GFR_BEGIN
  GF_REQUIRE(r && x_host && y_host, "null argument");
  gfgpu_ctx *ctx = r->ctx;
  GF_CUDA(cudaSetDevice(ctx->device));
  const int64_t nin = transposed ? r->nrows : r->ncols, nout = transposed ? r->ncols : r->nrows;
  gf::DevBuf<double> x, y;
  x.alloc(ctx, nin); y.alloc(ctx, nout);
  x.upload(x_host);
  if (beta != 0.0) y.upload(y_host); else y.zero();
  GF_REQUIRE(gfgpu_rect_mult_dev(r, transposed, alpha, x.p, beta, y.p) == 0, gfgpu_last_error());
  y.download(y_host);
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  GFR_END
}

}  // extern "C"
