// matrix.cu -- the workspace-level tangent kept on the device.
//
// ga_workspace::assembly(2) adds EVERY order-2 tree into one gmm::col_matrix<rsvector> (workspace.cc:791-936), a model
// then sums brick matrices into its tangent and forms residuals of linear bricks as K*u (getfem_models.cc:2536-2620,
// 2753-2900).  Here the terms' CSC slabs are accumulated into ONE device CSC:
//   pattern  = union of the terms' patterns at their (row, column) offsets -- an entry an element matrix inserted stays
//              stored even when a later term cancels it, like rsvector::w / add_elem_matrix (C&E.cc:4853-4936);
//   values   = sum over the terms in the order they were added; inside one term every (row, column) occurs once, so a
//              term is added by one kernel without atomics: results are bitwise reproducible;
//   products = y = K^T x (gather per column) and y = K x (gather per row through a row-sorted permutation built once per
//              pattern), fixed summation order, no atomics.
#include <algorithm>
#include <cmath>
#include <cub/cub.cuh>

#include "common.cuh"

struct gfgpu_matrix {
  gfgpu_ctx *ctx = nullptr;
  int64_t nrows = 0, ncols = 0, nnz = 0;
  int64_t generation = 0;       // pattern generation
  gf::DevBuf<int64_t> jc;       // ncols + 1
  gf::DevBuf<int32_t> ir;       // nnz, ascending inside a column
  gf::DevBuf<double> pr;        // nnz
  gf::DevBuf<int64_t> cnt;      // ncols + 1 scratch
  gf::DevBuf<int32_t> flag;
  // row-major view for y = K x: entries sorted by (row, column)
  int64_t csr_generation = -1;
  gf::DevBuf<int64_t> rp;       // nrows + 1
  gf::DevBuf<uint32_t> rperm;   // nnz: entry index
  gf::DevBuf<int32_t> rcol;     // nnz: column of that entry
};

namespace gf {

static inline int mgrid(int64_t n, int block, int cap = 148 * 32) {
  int64_t g = (n + block - 1) / block;
  return (int)std::max<int64_t>(1, std::min<int64_t>(g, cap));
}

// first position in [lo, hi) with ir[pos] >= row
__device__ __forceinline__ int64_t lower_row(const int32_t *__restrict__ ir, int64_t lo, int64_t hi, int32_t row) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (ir[mid] < row) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// warp per term column: how many of its rows are missing from the matrix column
__global__ void k_mat_missing(const int64_t *__restrict__ tjc, const int32_t *__restrict__ tir, int64_t tn, int64_t row_off,
                              int64_t col_off, const int64_t *__restrict__ jc, const int32_t *__restrict__ ir,
                              int64_t *__restrict__ cnt /* per matrix column: stored + missing */, int32_t *__restrict__ any) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t j = w0; j < tn; j += nw) {
    const int64_t a = tjc[j], b = tjc[j + 1], c = col_off + j;
    const int64_t lo = jc ? jc[c] : 0, hi = jc ? jc[c + 1] : 0;
    int miss = 0;
    for (int64_t k = a + lane; k < b; k += 32) {
      const int32_t r = (int32_t)(row_off + tir[k]);
      const int64_t p = lower_row(ir, lo, hi, r);
      if (p >= hi || ir[p] != r) ++miss;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) miss += __shfl_xor_sync(0xffffffffu, miss, o);
    if (lane == 0 && miss) {
      cnt[c] = (hi - lo) + miss;
      *any = 1;
    }
  }
}

__global__ void k_mat_counts(const int64_t *__restrict__ jc, int64_t n, int64_t *__restrict__ cnt) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
    cnt[j] = jc ? jc[j + 1] - jc[j] : 0;
}

// thread per matrix column: merge the old column with the term column (both row-sorted) into the new layout
__global__ void k_mat_merge(const int64_t *__restrict__ ojc, const int32_t *__restrict__ oir, const double *__restrict__ opr,
                            const int64_t *__restrict__ tjc, const int32_t *__restrict__ tir, int64_t tn, int64_t row_off,
                            int64_t col_off, const int64_t *__restrict__ njc, int64_t ncols, int32_t *__restrict__ nir,
                            double *__restrict__ npr) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncols; c += (int64_t)gridDim.x * blockDim.x) {
    int64_t a = ojc ? ojc[c] : 0;
    const int64_t ae = ojc ? ojc[c + 1] : 0;
    const int64_t j = c - col_off;
    int64_t b = (j >= 0 && j < tn) ? tjc[j] : 0;
    const int64_t be = (j >= 0 && j < tn) ? tjc[j + 1] : 0;
    int64_t o = njc[c];
    while (a < ae || b < be) {
      const int32_t ra = a < ae ? oir[a] : INT32_MAX;
      const int32_t rb = b < be ? (int32_t)(row_off + tir[b]) : INT32_MAX;
      if (ra <= rb) {
        nir[o] = ra; npr[o] = opr[a]; ++a;
        if (ra == rb) ++b;
      } else {
        nir[o] = rb; npr[o] = 0.0; ++b;
      }
      ++o;
    }
  }
}

// warp per term column: pr[pos(row, column)] += alpha * value.  Every (row, column) occurs once per term: no conflicts.
__global__ void k_mat_add(const int64_t *__restrict__ tjc, const int32_t *__restrict__ tir, const double *__restrict__ tpr,
                          int64_t tn, double alpha, int64_t row_off, int64_t col_off, const int64_t *__restrict__ jc,
                          const int32_t *__restrict__ ir, double *__restrict__ pr, int32_t *__restrict__ err) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t j = w0; j < tn; j += nw) {
    const int64_t a = tjc[j], b = tjc[j + 1], c = col_off + j;
    const int64_t lo = jc[c], hi = jc[c + 1];
    // the term's rows ascend with k: the search restarts from the previous hit of this lane
    int64_t from = lo;
    for (int64_t k = a + lane; k < b; k += 32) {
      const int32_t r = (int32_t)(row_off + tir[k]);
      const int64_t p = lower_row(ir, from, hi, r);
      if (p >= hi || ir[p] != r) { *err = 1; continue; }
      pr[p] += alpha * tpr[k];
      from = p + 1;
    }
  }
}

// y[col] = beta*y[col] + alpha * sum_k pr[k] x[ir[k]]   (K^T x), warp per column, fixed lane tree
__global__ void k_mat_tmult(const int64_t *__restrict__ jc, const int32_t *__restrict__ ir, const double *__restrict__ pr,
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

// y[row] = beta*y[row] + alpha * sum over the row's entries (ascending column) pr[e] x[col(e)]
__global__ void k_mat_mult(const int64_t *__restrict__ rp, const uint32_t *__restrict__ rperm, const int32_t *__restrict__ rcol,
                           const double *__restrict__ pr, int64_t nrows, const double *__restrict__ x, double alpha,
                           double beta, double *__restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = w0; r < nrows; r += nw) {
    double s = 0.0;
    for (int64_t k = rp[r] + lane; k < rp[r + 1]; k += 32) s += pr[rperm[k]] * x[rcol[k]];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[r] = (beta == 0.0 ? 0.0 : beta * y[r]) + alpha * s;
  }
}

__global__ void k_mat_entry_cols(const int64_t *__restrict__ jc, int64_t ncols, int32_t *__restrict__ col, uint32_t *__restrict__ idx) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = w0; c < ncols; c += nw)
    for (int64_t k = jc[c] + lane; k < jc[c + 1]; k += 32) { col[k] = (int32_t)c; idx[k] = (uint32_t)k; }
}

__global__ void k_mat_row_hist(const int32_t *__restrict__ ir, int64_t nnz, unsigned long long *__restrict__ cnt) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&cnt[ir[k]], 1ull);  // integer counts: the order of the additions does not matter
}

static void scan_counts(gfgpu_ctx *ctx, int64_t *cnt, int64_t *out, int64_t n) {
  size_t tb = 0;
  GF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, out, n, ctx->stream));
  void *tmp = cub_scratch(ctx, tb);
  GF_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, out, n, ctx->stream));
  count_launch(2);
}

static void mat_add(gfgpu_matrix *m, const int64_t *tjc, const int32_t *tir, const double *tpr, int64_t tn, int64_t tnnz,
                    double alpha, int64_t row_off, int64_t col_off) {
  gfgpu_ctx *ctx = m->ctx;
  cudaStream_t s = ctx->stream;
  if (!tn || !tnnz) return;
  const int B = 256;
  // ---- symbolic: does the matrix already store every (row, column) of the term?
  if (m->cnt.n != (size_t)m->ncols + 1) m->cnt.alloc(ctx, m->ncols + 1);
  m->flag.zero();
  k_mat_counts<<<mgrid(m->ncols, B), B, 0, s>>>(m->jc.n ? m->jc.p : nullptr, m->ncols, m->cnt.p);
  GF_LAUNCH_CHECK();
  GF_CUDA(cudaMemsetAsync(m->cnt.p + m->ncols, 0, sizeof(int64_t), s));
  k_mat_missing<<<mgrid(tn * 32, B), B, 0, s>>>(tjc, tir, tn, row_off, col_off, m->jc.n ? m->jc.p : nullptr, m->ir.p, m->cnt.p,
                                               m->flag.p);
  GF_LAUNCH_CHECK();
  int32_t grow = 0;
  m->flag.download(&grow);
  GF_CUDA(cudaStreamSynchronize(s));
  if (grow) {
    DevBuf<int64_t> njc;
    njc.alloc(ctx, m->ncols + 1);
    scan_counts(ctx, m->cnt.p, njc.p, m->ncols + 1);
    int64_t nnz = 0;
    GF_CUDA(cudaMemcpyAsync(&nnz, njc.p + m->ncols, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    GF_CUDA(cudaStreamSynchronize(s));
    DevBuf<int32_t> nir;
    DevBuf<double> npr;
    nir.alloc(ctx, nnz);
    npr.alloc(ctx, nnz);
    k_mat_merge<<<mgrid(m->ncols, 128), 128, 0, s>>>(m->jc.n ? m->jc.p : nullptr, m->ir.p, m->pr.p, tjc, tir, tn, row_off, col_off,
                                                    njc.p, m->ncols, nir.p, npr.p);
    GF_LAUNCH_CHECK();
    GF_CUDA(cudaStreamSynchronize(s));
    std::swap(m->jc.p, njc.p); std::swap(m->jc.n, njc.n); std::swap(m->jc.ctx, njc.ctx);
    std::swap(m->ir.p, nir.p); std::swap(m->ir.n, nir.n); std::swap(m->ir.ctx, nir.ctx);
    std::swap(m->pr.p, npr.p); std::swap(m->pr.n, npr.n); std::swap(m->pr.ctx, npr.ctx);
    m->nnz = nnz;
    m->generation++;
  }
  // ---- numeric
  m->flag.zero();
  k_mat_add<<<mgrid(tn * 32, B), B, 0, s>>>(tjc, tir, tpr, tn, alpha, row_off, col_off, m->jc.p, m->ir.p, m->pr.p, m->flag.p);
  GF_LAUNCH_CHECK();
  int32_t err = 0;
  m->flag.download(&err);
  GF_CUDA(cudaStreamSynchronize(s));
  GF_REQUIRE(err == 0, "internal error: a term entry has no slot in the matrix pattern");
}

// ---------------------------------------------------------------- model-level algebra on the resident tangent
// Dirichlet conditions with simplification (model::add_Dirichlet_condition_with_simplification ->
// real_dof_constraints, getfem_models.cc:2806-2871): rows (and, for a symmetric model, columns) of the constrained dofs
// are cleared and their diagonal set to 1.  mark[dof] = 1 + position in the constraint list, 0 = free.
__global__ void k_mat_mark(const int64_t *__restrict__ dof, int64_t n, int32_t *__restrict__ mark) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    mark[dof[k]] = (int32_t)k + 1;
}

// warp per column: entries in constrained rows -> 0; a constrained column of a symmetric model -> 0; diagonal -> 1
__global__ void k_mat_constrain(const int64_t *__restrict__ jc, const int32_t *__restrict__ ir, double *__restrict__ pr,
                                int64_t ncols, const int32_t *__restrict__ mark, int symmetric, int32_t *__restrict__ flag) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = w0; c < ncols; c += nw) {
    const bool cc = mark[c] != 0;
    bool diag = false;
    for (int64_t k = jc[c] + lane; k < jc[c + 1]; k += 32) {
      const int32_t r = ir[k];
      if (cc && r == c) { pr[k] = 1.0; diag = true; }
      else if (mark[r] || (cc && symmetric)) pr[k] = 0.0;
    }
    if (cc && !__any_sync(0xffffffffu, diag) && lane == 0) atomicExch(flag, 1);
  }
}

// rhs[dof[k]] = go[k] (linear model) or += go[k] - pr[k] (nonlinear: the Newton increment goes to the prescribed value)
__global__ void k_rhs_constrain(const int64_t *__restrict__ dof, const double *__restrict__ go, const double *__restrict__ prv,
                                int64_t n, int linear, double *__restrict__ rhs) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    if (linear) rhs[dof[k]] = go[k];
    else rhs[dof[k]] += go[k] - prv[k];
  }
}

__global__ void k_axpy_off(const double *__restrict__ x, int64_t n, double alpha, double *__restrict__ y) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) y[k] += alpha * x[k];
}

__global__ void k_scatter_vals(const int64_t *__restrict__ dof, const double *__restrict__ v, int64_t n, double *__restrict__ out) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) out[dof[k]] = v[k];
}

// CSR hand-off: values (and int32 row pointers when they fit) in row-major order for a device solver
__global__ void k_csr_vals(const uint32_t *__restrict__ rperm, const double *__restrict__ pr, int64_t nnz, double *__restrict__ val) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) val[k] = pr[rperm[k]];
}

// ---- Jacobi-preconditioned conjugate gradient on the resident matrix (symmetric positive definite tangent).  The point is
// the hand-off: K, the residual and the solution never leave the device.  Scalars live in device memory (sc[]):
// 0 rz, 1 pq, 2 rz_new, 3 rr; reductions are two-stage with a fixed order (bitwise reproducible).
constexpr int CG_PARTS = 148 * 4;
__global__ void __launch_bounds__(256) k_dot_part(const double *__restrict__ x, const double *__restrict__ y, int64_t n, double *__restrict__ part) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) s += x[k] * y[k];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) k_dot_final(const double *__restrict__ part, int np, double *__restrict__ out) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int k = threadIdx.x; k < np; k += 256) s += part[k];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}
__global__ void k_mat_diag_inv(const int64_t *__restrict__ jc, const int32_t *__restrict__ ir, const double *__restrict__ pr,
                               int64_t ncols, double *__restrict__ dinv) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncols; c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = lower_row(ir, jc[c], jc[c + 1], (int32_t)c);
    const double d = (k < jc[c + 1] && ir[k] == c) ? pr[k] : 0.0;
    dinv[c] = d != 0.0 ? 1.0 / d : 1.0;
  }
}
__global__ void k_cg_init(const double *__restrict__ b, const double *__restrict__ dinv, int64_t n, double *__restrict__ r,
                          double *__restrict__ z, double *__restrict__ p) {  // r = b - q (q = K x0 already in r's place: r holds K x0)
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const double rk = b[k] - r[k];
    r[k] = rk;
    z[k] = dinv[k] * rk;
    p[k] = z[k];
  }
}
__global__ void k_cg_step1(const double *__restrict__ sc, const double *__restrict__ p, const double *__restrict__ q,
                           const double *__restrict__ dinv, int64_t n, double *__restrict__ x, double *__restrict__ r,
                           double *__restrict__ z) {
  const double alpha = sc[1] != 0.0 ? sc[0] / sc[1] : 0.0;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    x[k] += alpha * p[k];
    const double rk = r[k] - alpha * q[k];
    r[k] = rk;
    z[k] = dinv[k] * rk;
  }
}
__global__ void k_cg_step2(double *__restrict__ sc, const double *__restrict__ z, int64_t n, double *__restrict__ p) {
  const double beta = sc[0] != 0.0 ? sc[2] / sc[0] : 0.0;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) p[k] = z[k] + beta * p[k];
}
__global__ void k_cg_roll(double *sc) { sc[0] = sc[2]; }

static void dot_dev(gfgpu_ctx *ctx, const double *x, const double *y, int64_t n, double *part, double *out) {
  k_dot_part<<<CG_PARTS, 256, 0, ctx->stream>>>(x, y, n, part);
  GF_LAUNCH_CHECK();
  k_dot_final<<<1, 256, 0, ctx->stream>>>(part, CG_PARTS, out);
  GF_LAUNCH_CHECK();
}

static void mat_build_csr(gfgpu_matrix *m) {
  if (m->csr_generation == m->generation) return;
  gfgpu_ctx *ctx = m->ctx;
  cudaStream_t s = ctx->stream;
  GF_REQUIRE(m->nnz < (int64_t(1) << 31) - 1, "y = K x needs nnz < 2^31 (use the transposed product, or split the matrix)");
  const int B = 256;
  const int64_t nnz = m->nnz;
  DevBuf<int32_t> col0, key1;
  DevBuf<uint32_t> idx0;
  col0.alloc(ctx, nnz); idx0.alloc(ctx, nnz); key1.alloc(ctx, nnz);
  m->rperm.alloc(ctx, nnz); m->rcol.alloc(ctx, nnz); m->rp.alloc(ctx, m->nrows + 1);
  k_mat_entry_cols<<<mgrid(m->ncols * 32, B), B, 0, s>>>(m->jc.p, m->ncols, col0.p, idx0.p);
  GF_LAUNCH_CHECK();
  int bits = 1;
  while (bits < 31 && (int64_t(1) << bits) < m->nrows) ++bits;
  // stable sort by row: inside a row the entries keep their CSC order = ascending column
  size_t tb = 0;
  GF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, m->ir.p, key1.p, idx0.p, m->rperm.p, (int)nnz, 0, bits, s));
  void *tmp = cub_scratch(ctx, tb);
  GF_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, m->ir.p, key1.p, idx0.p, m->rperm.p, (int)nnz, 0, bits, s));
  GF_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, m->ir.p, key1.p, col0.p, m->rcol.p, (int)nnz, 0, bits, s));
  count_launch(4);
  DevBuf<int64_t> cnt;
  cnt.alloc(ctx, m->nrows + 1);
  cnt.zero();
  k_mat_row_hist<<<mgrid(nnz, B), B, 0, s>>>(m->ir.p, nnz, (unsigned long long *)cnt.p);
  GF_LAUNCH_CHECK();
  scan_counts(ctx, cnt.p, m->rp.p, m->nrows + 1);
  GF_CUDA(cudaStreamSynchronize(s));
  m->csr_generation = m->generation;
}


// ---------------------------------------------------------------- reduced mesh_fem: K += E^T K_basic E, V += E^T V_basic
// A reduced mesh_fem (mesh_fem::is_reduced(): partial_mesh_fem -- the multiplier spaces of the Dirichlet bricks --, periodic or
// enriched spaces) is assembled by the reference on its BASIC dofs into unreduced matrices and then projected with the
// extension matrix E (nb_basic_dof x nb_dof): K(I1, I2) += E1^T K_basic E2, V(I1) += E1^T V_basic (workspace.cc:861-935,
// gmm::mult of sparse matrices).  Here each side is one expand-sort-compress pass on the device:
//   rows pass  A = E^T S : every stored (j, k, v) of S and every stored (r, a) of row j of E gives (r, k, a v)
//   cols pass  M = A E   : every stored (r, k, v) of A and every stored (c, b) of row k of E gives (r, c, v b)
// products are expanded in the CSC order of the source, radix-sorted (CUB, stable) by (column, row) and summed per entry in
// that order -- ascending inner index, the order of gmm's column / rank-one products -- by one thread per entry; sums that are
// exactly 0.0 are not stored (rsvector::w removes them).  No atomics on values, bitwise reproducible.
__global__ void k_red_count(const int64_t *__restrict__ jc, const int32_t *__restrict__ ir, int64_t ncols, const int64_t *__restrict__ erp,
                            int side, int64_t *__restrict__ cnt) {
  // one warp per source column
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (int64_t k = warp; k < ncols; k += nw) {
    const int64_t lenk = side ? erp[k + 1] - erp[k] : 0;
    for (int64_t e = jc[k] + lane; e < jc[k + 1]; e += 32) cnt[e] = side ? lenk : erp[ir[e] + 1] - erp[ir[e]];
  }
}

__global__ void k_red_expand(const int64_t *__restrict__ jc, const int32_t *__restrict__ ir, const double *__restrict__ pr, int64_t ncols,
                             const int64_t *__restrict__ erp, const int32_t *__restrict__ ecol, const double *__restrict__ eval,
                             int side, int64_t out_nrows, const int64_t *__restrict__ off, unsigned long long *__restrict__ key,
                             double *__restrict__ val) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (int64_t k = warp; k < ncols; k += nw) {
    for (int64_t e = jc[k] + lane; e < jc[k + 1]; e += 32) {
      const int64_t j = ir[e], o = off[e];
      const double v = pr[e];
      if (!side) {  // rows pass: (r, k) for every (r, a) of E's row j
        for (int64_t q = erp[j]; q < erp[j + 1]; ++q) {
          key[o + q - erp[j]] = (unsigned long long)k * (unsigned long long)out_nrows + (unsigned long long)ecol[q];
          val[o + q - erp[j]] = eval[q] * v;
        }
      } else {      // cols pass: (j, c) for every (c, b) of E's row k
        for (int64_t q = erp[k]; q < erp[k + 1]; ++q) {
          key[o + q - erp[k]] = (unsigned long long)ecol[q] * (unsigned long long)out_nrows + (unsigned long long)j;
          val[o + q - erp[k]] = v * eval[q];
        }
      }
    }
  }
}

// thread per sorted product: the first product of an entry sums its run in order and flags a nonzero result
__global__ void k_red_sum(const unsigned long long *__restrict__ key, const double *__restrict__ val, int64_t n,
                          double *__restrict__ sum, int64_t *__restrict__ keep) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t kp = 0;
    if (i == 0 || key[i] != key[i - 1]) {
      double s = val[i];
      for (int64_t q = i + 1; q < n && key[q] == key[i]; ++q) s += val[q];
      sum[i] = s;
      kp = s != 0.0 ? 1 : 0;
    }
    keep[i] = kp;
  }
}

__global__ void k_red_compact(const unsigned long long *__restrict__ key, const double *__restrict__ sum, const int64_t *__restrict__ keep,
                              const int64_t *__restrict__ pos, int64_t n, int64_t out_nrows, int32_t *__restrict__ ir,
                              double *__restrict__ pr, unsigned long long *__restrict__ colcnt) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (!keep[i]) continue;
    const unsigned long long c = key[i] / (unsigned long long)out_nrows;
    ir[pos[i]] = (int32_t)(key[i] % (unsigned long long)out_nrows);
    pr[pos[i]] = sum[i];
    atomicAdd(colcnt + c, 1ull);  // a count, not a value: order does not matter
  }
}

__global__ void k_red_extend(const int64_t *__restrict__ erp, const int32_t *__restrict__ ecol, const double *__restrict__ eval, int64_t n,
                             const double *__restrict__ x, double *__restrict__ y) {  // y = E x, thread per basic dof
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int64_t q = erp[j]; q < erp[j + 1]; ++q) s += eval[q] * x[ecol[q]];
    y[j] = s;
  }
}

__global__ void k_red_restrict(const int64_t *__restrict__ trp, const int32_t *__restrict__ tcol, const double *__restrict__ tval, int64_t n,
                               double alpha, const double *__restrict__ x, double *__restrict__ y) {  // y += alpha E^T x, thread per dof
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int64_t q = trp[r]; q < trp[r + 1]; ++q) s += tval[q] * x[tcol[q]];
    y[r] += alpha * s;
  }
}

struct CscBuf {
  DevBuf<int64_t> jc;
  DevBuf<int32_t> ir;
  DevBuf<double> pr;
  int64_t nrows = 0, ncols = 0, nnz = 0;
};

// one pass: out = E^T S (side 0) or S E (side 1); S is nrows_in x ncols_in in CSC
static void red_pass(gfgpu_ctx *ctx, const int64_t *jc, const int32_t *ir, const double *pr, int64_t nrows_in, int64_t ncols_in,
                     int64_t nnz_in, const gfgpu_reduction *E, int side, CscBuf &out) {
  cudaStream_t s = ctx->stream;
  GF_REQUIRE((side ? ncols_in : nrows_in) == E->n_basic, "the extension matrix does not match the block");
  out.nrows = side ? nrows_in : E->n_dof;
  out.ncols = side ? E->n_dof : ncols_in;
  out.jc.alloc(ctx, out.ncols + 1);
  out.jc.zero();
  out.nnz = 0;
  if (!nnz_in) return;
  const int B = 256;
  DevBuf<int64_t> cnt, off;
  cnt.alloc(ctx, nnz_in + 1);
  off.alloc(ctx, nnz_in + 1);
  GF_CUDA(cudaMemsetAsync(cnt.p + nnz_in, 0, sizeof(int64_t), s));
  k_red_count<<<mgrid(ncols_in * 32, B), B, 0, s>>>(jc, ir, ncols_in, E->rp.p, side, cnt.p);
  GF_LAUNCH_CHECK();
  scan_counts(ctx, cnt.p, off.p, nnz_in + 1);
  int64_t T = 0;
  GF_CUDA(cudaMemcpyAsync(&T, off.p + nnz_in, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  GF_CUDA(cudaStreamSynchronize(s));
  cnt.release();
  if (!T) return;
  DevBuf<unsigned long long> key, key2;
  DevBuf<double> val, val2;
  key.alloc(ctx, T); key2.alloc(ctx, T); val.alloc(ctx, T); val2.alloc(ctx, T);
  k_red_expand<<<mgrid(ncols_in * 32, B), B, 0, s>>>(jc, ir, pr, ncols_in, E->rp.p, E->col.p, E->val.p, side, out.nrows, off.p, key.p,
                                                    val.p);
  GF_LAUNCH_CHECK();
  off.release();
  int bits = 1;
  while (bits < 64 && ((unsigned long long)out.nrows * (unsigned long long)out.ncols) >> bits) ++bits;
  size_t tb = 0;
  GF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, key.p, key2.p, val.p, val2.p, T, 0, bits, s));
  void *tmp = cub_scratch(ctx, tb);
  GF_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, key.p, key2.p, val.p, val2.p, T, 0, bits, s));  // stable: products stay in order
  count_launch(3);
  key.release();
  DevBuf<int64_t> keep, pos;
  keep.alloc(ctx, T + 1);
  pos.alloc(ctx, T + 1);
  GF_CUDA(cudaMemsetAsync(keep.p + T, 0, sizeof(int64_t), s));
  k_red_sum<<<mgrid(T, B), B, 0, s>>>(key2.p, val2.p, T, val.p, keep.p);
  GF_LAUNCH_CHECK();
  scan_counts(ctx, keep.p, pos.p, T + 1);
  GF_CUDA(cudaMemcpyAsync(&out.nnz, pos.p + T, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  GF_CUDA(cudaStreamSynchronize(s));
  if (!out.nnz) return;
  out.ir.alloc(ctx, out.nnz);
  out.pr.alloc(ctx, out.nnz);
  DevBuf<int64_t> colcnt;
  colcnt.alloc(ctx, out.ncols + 1);
  colcnt.zero();
  k_red_compact<<<mgrid(T, B), B, 0, s>>>(key2.p, val.p, keep.p, pos.p, T, out.nrows, out.ir.p, out.pr.p,
                                                   reinterpret_cast<unsigned long long *>(colcnt.p));
  GF_LAUNCH_CHECK();
  scan_counts(ctx, colcnt.p, out.jc.p, out.ncols + 1);
  GF_CUDA(cudaStreamSynchronize(s));
}

// m(row_off.., col_off..) += alpha * Er^T S Ec  (a null extension matrix = that side is not reduced)
static void mat_add_projected(gfgpu_matrix *m, const int64_t *jc, const int32_t *ir, const double *pr, int64_t nrows, int64_t ncols,
                              int64_t nnz, const gfgpu_reduction *Er, const gfgpu_reduction *Ec, double alpha, int64_t row_off,
                              int64_t col_off) {
  gfgpu_ctx *ctx = m->ctx;
  CscBuf a, b;
  if (Er) {
    red_pass(ctx, jc, ir, pr, nrows, ncols, nnz, Er, 0, a);
    jc = a.jc.p; ir = a.ir.p; pr = a.pr.p; nrows = a.nrows; ncols = a.ncols; nnz = a.nnz;
  }
  if (Ec) {
    red_pass(ctx, jc, ir, pr, nrows, ncols, nnz, Ec, 1, b);
    jc = b.jc.p; ir = b.ir.p; pr = b.pr.p; nrows = b.nrows; ncols = b.ncols; nnz = b.nnz;
  }
  GF_REQUIRE(row_off >= 0 && col_off >= 0 && row_off + nrows <= m->nrows && col_off + ncols <= m->ncols,
             "the projected block does not fit the matrix");
  mat_add(m, jc, ir, pr, ncols, nnz, alpha, row_off, col_off);
}

}  // namespace gf

namespace gf { void set_last_error(const std::string &); }  // api.cu

#define GFM_BEGIN try {
#define GFM_END                                  \
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

int gfgpu_matrix_create(gfgpu_ctx *ctx, int64_t nrows, int64_t ncols, gfgpu_matrix **out) {
  GFM_BEGIN
  GF_REQUIRE(ctx && out, "null argument");
  GF_REQUIRE(nrows >= 0 && ncols >= 0 && nrows < (int64_t(1) << 31) - 4 && ncols < (int64_t(1) << 31) - 4, "bad matrix sizes");
  GF_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<gfgpu_matrix> m(new gfgpu_matrix);
  m->ctx = ctx; m->nrows = nrows; m->ncols = ncols;
  m->flag.alloc(ctx, 1);
  *out = m.release();
  GFM_END
}

int gfgpu_matrix_destroy(gfgpu_matrix *m) {
  GFM_BEGIN
  if (m) {
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
  }
  delete m;
  GFM_END
}

int gfgpu_matrix_clear(gfgpu_matrix *m, int keep_pattern) {
  GFM_BEGIN
  GF_REQUIRE(m, "null matrix");
  GF_CUDA(cudaSetDevice(m->ctx->device));
  if (keep_pattern) {
    m->pr.zero();
  } else {
    m->jc.release(); m->ir.release(); m->pr.release();
    m->nnz = 0;
    m->generation++;
  }
  GFM_END
}

int gfgpu_matrix_add_term(gfgpu_matrix *m, gfgpu_term *t, double alpha, int64_t row_off, int64_t col_off) {
  GFM_BEGIN
  GF_REQUIRE(m && t, "null argument");
  GF_REQUIRE(m->ctx == t->ctx, "matrix and term live on different contexts");
  GF_REQUIRE(t->pat_valid, "the term has no assembled tangent");
  gf::term_settle_pending(t);
  const int64_t n = t->fem->ndof;
  GF_REQUIRE(row_off >= 0 && col_off >= 0 && row_off + n <= m->nrows && col_off + n <= m->ncols, "the term does not fit the matrix");
  GF_CUDA(cudaSetDevice(m->ctx->device));
  gf::mat_add(m, t->jc.p, t->ir.p, t->pr.p, n, t->nnz, alpha, row_off, col_off);
  GFM_END
}

int64_t gfgpu_matrix_nnz(gfgpu_matrix *m) { return m ? m->nnz : -1; }
int64_t gfgpu_matrix_pattern_generation(gfgpu_matrix *m) { return m ? m->generation : -1; }

int gfgpu_matrix_csc_view(gfgpu_matrix *m, const int64_t **jc, const int32_t **ir, const double **pr) {
  GFM_BEGIN
  GF_REQUIRE(m, "null matrix");
  if (jc) *jc = m->jc.p;
  if (ir) *ir = m->ir.p;
  if (pr) *pr = m->pr.p;
  GFM_END
}

int gfgpu_matrix_export_csc_host(gfgpu_matrix *m, int64_t *jc, int32_t *ir, double *pr) {
  GFM_BEGIN
  GF_REQUIRE(m, "null matrix");
  GF_CUDA(cudaSetDevice(m->ctx->device));
  if (jc) {
    if (m->jc.n) m->jc.download(jc);
    else for (int64_t k = 0; k <= m->ncols; ++k) jc[k] = 0;
  }
  if (ir) m->ir.download(ir);
  if (pr) m->pr.download(pr);
  GF_CUDA(cudaStreamSynchronize(m->ctx->stream));
  GFM_END
}

int gfgpu_matrix_mult_dev(gfgpu_matrix *m, int transposed, double alpha, const double *x_dev, double beta, double *y_dev) {
  GFM_BEGIN
  GF_REQUIRE(m && x_dev && y_dev, "null argument");
  GF_CUDA(cudaSetDevice(m->ctx->device));
  cudaStream_t s = m->ctx->stream;
  const int B = 256;
  const int64_t nout = transposed ? m->ncols : m->nrows;
  if (!m->nnz) {  // empty matrix: y = beta * y
    if (beta == 0.0) GF_CUDA(cudaMemsetAsync(y_dev, 0, nout * sizeof(double), s));
    else GF_REQUIRE(beta == 1.0, "empty matrix: only beta = 0 or 1 is handled");
    return 0;
  }
  if (transposed) {
    gf::k_mat_tmult<<<gf::mgrid(m->ncols * 32, B), B, 0, s>>>(m->jc.p, m->ir.p, m->pr.p, m->ncols, x_dev, alpha, beta, y_dev);
    GF_LAUNCH_CHECK();
  } else {
    gf::mat_build_csr(m);
    gf::k_mat_mult<<<gf::mgrid(m->nrows * 32, B), B, 0, s>>>(m->rp.p, m->rperm.p, m->rcol.p, m->pr.p, m->nrows, x_dev, alpha, beta, y_dev);
    GF_LAUNCH_CHECK();
  }
  GFM_END
}

int gfgpu_matrix_mult_host(gfgpu_matrix *m, int transposed, double alpha, const double *x_host, double beta, double *y_host) {
  GFM_BEGIN
  GF_REQUIRE(m && x_host && y_host, "null argument");
  gfgpu_ctx *ctx = m->ctx;
  GF_CUDA(cudaSetDevice(ctx->device));
  const int64_t nin = transposed ? m->nrows : m->ncols, nout = transposed ? m->ncols : m->nrows;
  gf::DevBuf<double> x, y;
  x.alloc(ctx, nin); y.alloc(ctx, nout);
  x.upload(x_host);
  if (beta != 0.0) y.upload(y_host); else y.zero();
  GF_REQUIRE(gfgpu_matrix_mult_dev(m, transposed, alpha, x.p, beta, y.p) == 0, gfgpu_last_error());
  y.download(y_host);
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  GFM_END
}

int gfgpu_matrix_apply_dof_constraints(gfgpu_matrix *m, int64_t n, const int64_t *dof_host, const double *go_host,
                                       const double *pr_host, double *rhs_dev, int flags) {
  GFM_BEGIN
  GF_REQUIRE(m && m->nrows == m->ncols, "dof constraints need a square matrix");
  GF_REQUIRE(n >= 0 && (n == 0 || (dof_host && go_host)), "null argument");
  const bool linear = flags & GFGPU_MODEL_LINEAR, symmetric = flags & GFGPU_MODEL_SYMMETRIC, do_m = flags & GFGPU_BUILD_MATRIX;
  GF_REQUIRE(linear || !rhs_dev || pr_host, "a nonlinear model needs the present values of the constrained dofs");
  if (!n) return 0;
  gfgpu_ctx *ctx = m->ctx;
  GF_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const int B = 256;
  for (int64_t k = 0; k < n; ++k) GF_REQUIRE(dof_host[k] >= 0 && dof_host[k] < m->nrows, "constrained dof out of range");
  gf::DevBuf<int64_t> dof;
  gf::DevBuf<double> go, prv;
  dof.alloc(ctx, n); go.alloc(ctx, n);
  dof.upload(dof_host); go.upload(go_host);
  if (pr_host) { prv.alloc(ctx, n); prv.upload(pr_host); }
  if (rhs_dev) {
    if (linear && symmetric) {  // rhs -= K(:, SI) go, with the matrix as it is BEFORE the rows and columns are cleared
      double nrm = 0;
      for (int64_t k = 0; k < n; ++k) nrm += go_host[k] * go_host[k];
      if (nrm > 0) {
        GF_REQUIRE(do_m, "Rhs only for a symmetric linear problem with dof constraint not allowed");  // models.cc:2843-2845
        gf::DevBuf<double> full;
        full.alloc(ctx, m->ncols);
        full.zero();
        gf::k_scatter_vals<<<gf::mgrid(n, B), B, 0, s>>>(dof.p, go.p, n, full.p);
        GF_LAUNCH_CHECK();
        GF_REQUIRE(gfgpu_matrix_mult_dev(m, 0, -1.0, full.p, 1.0, rhs_dev) == 0, gfgpu_last_error());
        GF_CUDA(cudaStreamSynchronize(s));
      }
    }
    gf::k_rhs_constrain<<<gf::mgrid(n, B), B, 0, s>>>(dof.p, go.p, prv.p, n, linear ? 1 : 0, rhs_dev);
    GF_LAUNCH_CHECK();
  }
  if (do_m) {
    // the diagonal slots of the constrained dofs must exist: add them to the pattern (alpha = 0) if some are missing
    {
      std::vector<int64_t> sd(dof_host, dof_host + n);
      std::sort(sd.begin(), sd.end());
      sd.erase(std::unique(sd.begin(), sd.end()), sd.end());
      std::vector<int64_t> tjc(m->ncols + 1, 0);
      std::vector<int32_t> tir(sd.size());
      std::vector<double> tpr(sd.size(), 0.0);
      for (size_t k = 0; k < sd.size(); ++k) { tjc[sd[k] + 1] = 1; tir[k] = (int32_t)sd[k]; }
      for (int64_t c = 0; c < m->ncols; ++c) tjc[c + 1] += tjc[c];
      gf::DevBuf<int64_t> djc; gf::DevBuf<int32_t> dir; gf::DevBuf<double> dpr;
      djc.alloc(ctx, tjc.size()); dir.alloc(ctx, tir.size()); dpr.alloc(ctx, tpr.size());
      djc.upload(tjc.data()); dir.upload(tir.data()); dpr.upload(tpr.data());
      gf::mat_add(m, djc.p, dir.p, dpr.p, m->ncols, (int64_t)sd.size(), 0.0, 0, 0);
    }
    gf::DevBuf<int32_t> mark;
    mark.alloc(ctx, m->nrows);
    mark.zero();
    gf::k_mat_mark<<<gf::mgrid(n, B), B, 0, s>>>(dof.p, n, mark.p);
    GF_LAUNCH_CHECK();
    m->flag.zero();
    gf::k_mat_constrain<<<gf::mgrid(m->ncols * 32, B), B, 0, s>>>(m->jc.p, m->ir.p, m->pr.p, m->ncols, mark.p, symmetric ? 1 : 0,
                                                                 m->flag.p);
    GF_LAUNCH_CHECK();
    int32_t err = 0;
    m->flag.download(&err);
    GF_CUDA(cudaStreamSynchronize(s));
    GF_REQUIRE(err == 0, "internal error: a constrained dof has no diagonal slot");
  }
  GF_CUDA(cudaStreamSynchronize(s));
  GFM_END
}

int gfgpu_matrix_export_csr_dev(gfgpu_matrix *m, int64_t *rowptr_dev, int32_t *col_dev, double *val_dev) {
  GFM_BEGIN
  GF_REQUIRE(m, "null matrix");
  GF_CUDA(cudaSetDevice(m->ctx->device));
  cudaStream_t s = m->ctx->stream;
  if (!m->nnz) {
    if (rowptr_dev) GF_CUDA(cudaMemsetAsync(rowptr_dev, 0, (m->nrows + 1) * sizeof(int64_t), s));
    return 0;
  }
  gf::mat_build_csr(m);
  if (rowptr_dev) GF_CUDA(cudaMemcpyAsync(rowptr_dev, m->rp.p, (m->nrows + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
  if (col_dev) GF_CUDA(cudaMemcpyAsync(col_dev, m->rcol.p, m->nnz * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  if (val_dev) {
    gf::k_csr_vals<<<gf::mgrid(m->nnz, 256), 256, 0, s>>>(m->rperm.p, m->pr.p, m->nnz, val_dev);
    GF_LAUNCH_CHECK();
  }
  GFM_END
}

int gfgpu_matrix_cg_dev(gfgpu_matrix *m, const double *b_dev, double *x_dev, double rtol, int max_iter, int *iters_out,
                        double *relres_out) {
  GFM_BEGIN
  GF_REQUIRE(m && b_dev && x_dev && m->nrows == m->ncols && m->nnz, "cg needs a non-empty square matrix and device vectors");
  gfgpu_ctx *ctx = m->ctx;
  GF_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const int64_t n = m->nrows;
  const int B = 256, G = gf::mgrid(n, B);
  gf::DevBuf<double> r, z, p, q, dinv, part, sc;
  r.alloc(ctx, n); z.alloc(ctx, n); p.alloc(ctx, n); q.alloc(ctx, n); dinv.alloc(ctx, n);
  part.alloc(ctx, gf::CG_PARTS); sc.alloc(ctx, 4);
  // K is symmetric here: K x is formed as K^T x (a gather per column, no row-major copy of the matrix)
  auto apply = [&](const double *x, double *y) {
    gf::k_mat_tmult<<<gf::mgrid(n * 32, B), B, 0, s>>>(m->jc.p, m->ir.p, m->pr.p, n, x, 1.0, 0.0, y);
    GF_LAUNCH_CHECK();
  };
  gf::k_mat_diag_inv<<<G, B, 0, s>>>(m->jc.p, m->ir.p, m->pr.p, n, dinv.p);
  GF_LAUNCH_CHECK();
  apply(x_dev, r.p);
  gf::k_cg_init<<<G, B, 0, s>>>(b_dev, dinv.p, n, r.p, z.p, p.p);
  GF_LAUNCH_CHECK();
  gf::dot_dev(ctx, r.p, z.p, n, part.p, sc.p + 0);
  gf::dot_dev(ctx, b_dev, b_dev, n, part.p, sc.p + 3);
  double h[4];
  sc.download(h);
  GF_CUDA(cudaStreamSynchronize(s));
  const double bb = h[3] > 0 ? h[3] : 1.0;
  int it = 0;
  double rr = bb;
  const int check_every = 8;
  while (it < max_iter) {
    apply(p.p, q.p);
    gf::dot_dev(ctx, p.p, q.p, n, part.p, sc.p + 1);
    gf::k_cg_step1<<<G, B, 0, s>>>(sc.p, p.p, q.p, dinv.p, n, x_dev, r.p, z.p);
    GF_LAUNCH_CHECK();
    gf::dot_dev(ctx, r.p, z.p, n, part.p, sc.p + 2);
    gf::k_cg_step2<<<G, B, 0, s>>>(sc.p, z.p, n, p.p);
    GF_LAUNCH_CHECK();
    gf::k_cg_roll<<<1, 1, 0, s>>>(sc.p);
    GF_LAUNCH_CHECK();
    ++it;
    if (it % check_every == 0 || it == max_iter) {  // the only host round trip: one scalar every few iterations
      gf::dot_dev(ctx, r.p, r.p, n, part.p, sc.p + 3);
      sc.download(h);
      GF_CUDA(cudaStreamSynchronize(s));
      rr = h[3];
      if (!(rr == rr)) break;
      if (rr <= rtol * rtol * bb) break;
    }
  }
  if (iters_out) *iters_out = it;
  if (relres_out) *relres_out = std::sqrt(rr / bb);
  GFM_END
}

int gfgpu_term_residual_add_dev(gfgpu_term *t, double alpha, double *rhs_dev, int64_t row_off) {
  GFM_BEGIN
  GF_REQUIRE(t && rhs_dev && t->R.n, "no assembled residual");
  GF_REQUIRE(row_off >= 0, "bad offset");
  GF_CUDA(cudaSetDevice(t->ctx->device));
  gf::k_axpy_off<<<gf::mgrid((int64_t)t->R.n, 256), 256, 0, t->ctx->stream>>>(t->R.p, (int64_t)t->R.n, alpha, rhs_dev + row_off);
  GF_LAUNCH_CHECK();
  GFM_END
}

int gfgpu_term_tmult_dev(gfgpu_term *t, double alpha, const double *x_dev, double beta, double *y_dev) {
  GFM_BEGIN
  GF_REQUIRE(t && x_dev && y_dev, "null argument");
  GF_REQUIRE(t->pat_valid, "the term has no assembled tangent");
  gf::term_settle_pending(t);
  GF_CUDA(cudaSetDevice(t->ctx->device));
  const int64_t n = t->fem->ndof;
  if (n) gf::k_mat_tmult<<<gf::mgrid(n * 32, 256), 256, 0, t->ctx->stream>>>(t->jc.p, t->ir.p, t->pr.p, n, x_dev, alpha, beta, y_dev);
  GF_LAUNCH_CHECK();
  GFM_END
}

int gfgpu_matrix_add_rect(gfgpu_matrix *m, gfgpu_rect *r, int transposed, double alpha, int64_t row_off, int64_t col_off) {
  GFM_BEGIN
  GF_REQUIRE(m && r, "null argument");
  GF_REQUIRE(m->ctx == r->ctx, "matrix and term live on different contexts");
  GF_REQUIRE(r->pat_valid, "the coupled term has no assembled block");
  const int64_t nr = transposed ? r->ncols : r->nrows, nc = transposed ? r->nrows : r->ncols;
  GF_REQUIRE(row_off >= 0 && col_off >= 0 && row_off + nr <= m->nrows && col_off + nc <= m->ncols, "the block does not fit the matrix");
  GF_CUDA(cudaSetDevice(m->ctx->device));
  if (transposed) gf::mat_add(m, r->jct.p, r->irt.p, r->prt.p, nc, r->nnz, alpha, row_off, col_off);
  else gf::mat_add(m, r->jc.p, r->ir.p, r->pr.p, nc, r->nnz, alpha, row_off, col_off);
  GFM_END
}


int gfgpu_reduction_create(gfgpu_ctx *ctx, int64_t n_basic, int64_t n_dof, const int64_t *rowptr, const int32_t *col, const double *val,
                           gfgpu_reduction **out) {
  GFM_BEGIN
  GF_REQUIRE(ctx && rowptr && out && n_basic >= 0 && n_dof >= 0, "bad argument");
  const int64_t nnz = rowptr[n_basic];
  GF_REQUIRE(nnz == 0 || (col && val), "null argument");
  for (int64_t j = 0; j < n_basic; ++j) {
    GF_REQUIRE(rowptr[j] <= rowptr[j + 1], "row pointers must ascend");
    for (int64_t q = rowptr[j]; q < rowptr[j + 1]; ++q)
      GF_REQUIRE(col[q] >= 0 && col[q] < n_dof && (q == rowptr[j] || col[q - 1] < col[q]), "columns out of range or not ascending");
  }
  GF_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<gfgpu_reduction> E(new gfgpu_reduction);
  E->ctx = ctx; E->n_basic = n_basic; E->n_dof = n_dof; E->nnz = nnz;
  E->rp.alloc(ctx, n_basic + 1); E->rp.upload(rowptr);
  E->col.alloc(ctx, nnz); E->col.upload(col);
  E->val.alloc(ctx, nnz); E->val.upload(val);
  // the transpose (rows = dofs, ascending basic dof inside a row): V += E^T V_basic as an ordered gather
  std::vector<int64_t> trp(n_dof + 1, 0);
  for (int64_t q = 0; q < nnz; ++q) ++trp[col[q] + 1];
  for (int64_t r = 0; r < n_dof; ++r) trp[r + 1] += trp[r];
  std::vector<int32_t> tcol(nnz);
  std::vector<double> tval(nnz);
  std::vector<int64_t> fill(trp.begin(), trp.end() - 1);
  for (int64_t j = 0; j < n_basic; ++j)
    for (int64_t q = rowptr[j]; q < rowptr[j + 1]; ++q) { tcol[fill[col[q]]] = (int32_t)j; tval[fill[col[q]]++] = val[q]; }
  E->trp.alloc(ctx, n_dof + 1); E->trp.upload(trp.data());
  E->tcol.alloc(ctx, nnz); E->tcol.upload(tcol.data());
  E->tval.alloc(ctx, nnz); E->tval.upload(tval.data());
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = E.release();
  GFM_END
}

int gfgpu_reduction_destroy(gfgpu_reduction *E) {
  GFM_BEGIN
  if (E) { cudaSetDevice(E->ctx->device); delete E; }
  GFM_END
}

int gfgpu_reduction_extend_host(gfgpu_reduction *E, const double *x_host, double *y_host) {
  GFM_BEGIN
  GF_REQUIRE(E && x_host && y_host, "null argument");
  gfgpu_ctx *ctx = E->ctx;
  GF_CUDA(cudaSetDevice(ctx->device));
  gf::DevBuf<double> x, y;
  x.alloc(ctx, E->n_dof); y.alloc(ctx, E->n_basic);
  x.upload(x_host);
  if (E->n_basic) gf::k_red_extend<<<gf::mgrid(E->n_basic, 256), 256, 0, ctx->stream>>>(E->rp.p, E->col.p, E->val.p, E->n_basic, x.p, y.p);
  GF_LAUNCH_CHECK();
  y.download(y_host);
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  GFM_END
}

int gfgpu_reduction_restrict_add_host(gfgpu_reduction *E, double alpha, const double *x_host, double *y_host) {
  GFM_BEGIN
  GF_REQUIRE(E && x_host && y_host, "null argument");
  gfgpu_ctx *ctx = E->ctx;
  GF_CUDA(cudaSetDevice(ctx->device));
  gf::DevBuf<double> x, y;
  x.alloc(ctx, E->n_basic); y.alloc(ctx, E->n_dof);
  x.upload(x_host); y.upload(y_host);
  if (E->n_dof) gf::k_red_restrict<<<gf::mgrid(E->n_dof, 256), 256, 0, ctx->stream>>>(E->trp.p, E->tcol.p, E->tval.p, E->n_dof, alpha, x.p, y.p);
  GF_LAUNCH_CHECK();
  y.download(y_host);
  GF_CUDA(cudaStreamSynchronize(ctx->stream));
  GFM_END
}

int gfgpu_matrix_add_term_reduced(gfgpu_matrix *m, gfgpu_term *t, gfgpu_reduction *E, double alpha, int64_t row_off, int64_t col_off) {
  GFM_BEGIN
  GF_REQUIRE(m && t && E, "null argument");
  GF_REQUIRE(m->ctx == t->ctx && m->ctx == E->ctx, "matrix, term and extension matrix live on different contexts");
  GF_REQUIRE(t->pat_valid, "the term has no assembled tangent");
  gf::term_settle_pending(t);
  GF_CUDA(cudaSetDevice(m->ctx->device));
  gf::mat_add_projected(m, t->jc.p, t->ir.p, t->pr.p, t->fem->ndof, t->fem->ndof, t->nnz, E, E, alpha, row_off, col_off);
  GFM_END
}

int gfgpu_matrix_add_rect_reduced(gfgpu_matrix *m, gfgpu_rect *r, int transposed, gfgpu_reduction *E_rows, gfgpu_reduction *E_cols,
                                  double alpha, int64_t row_off, int64_t col_off) {
  GFM_BEGIN
  GF_REQUIRE(m && r, "null argument");
  GF_REQUIRE(m->ctx == r->ctx && (!E_rows || E_rows->ctx == m->ctx) && (!E_cols || E_cols->ctx == m->ctx),
             "matrix, term and extension matrices live on different contexts");
  GF_REQUIRE(r->pat_valid, "the coupled term has no assembled block");
  const int64_t nr = transposed ? r->ncols : r->nrows, nc = transposed ? r->nrows : r->ncols;
  GF_CUDA(cudaSetDevice(m->ctx->device));
  if (transposed) gf::mat_add_projected(m, r->jct.p, r->irt.p, r->prt.p, nr, nc, r->nnz, E_rows, E_cols, alpha, row_off, col_off);
  else gf::mat_add_projected(m, r->jc.p, r->ir.p, r->pr.p, nr, nc, r->nnz, E_rows, E_cols, alpha, row_off, col_off);
  GFM_END
}

}  // extern "C"
