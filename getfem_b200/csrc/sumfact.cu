// sumfact.cu -- sum-factorised element kernel for the scalar Laplace form on high-order hexahedra (BASELINE config 5:
// FEM_QK(3,4), IM_GAUSS_PARALLELEPIPED(3,8), GT_QK(3,1)).  Strategy STAGED: same outputs as the generic element kernel
// (elem_kernel.cuh: element matrix column-major in `stage`, per-block keep masks, element residual), 13x fewer flops.
//
// Replaces, for this family, rows 6-15 of SURVEY 8(a): K(), B() per Gauss point (bgeot_geometric_trans.cc:270-413),
// pfp_grad_base_value (getfem_fem.cc:160-199), the per-point tensor algebra of ga_exec (C&E.cc:2769-3760) and the
// accumulation elem += J w_q t (C&E.cc:5047-5057), which cost the reference 2 nd^2 nq N = 11.7 Mflop per Q4 element.
//
// A tensor-product basis phi_i = l_{i1}(x) l_{i2}(y) l_{i3}(z) on a tensor-product rule q = (q1, q2, q3) gives
//   K_e(i, j) = sum_{a,b} sum_{q1 q2 q3} C_ab(q) X^a_{i1}(q1) X^b_{j1}(q1) Y^a_{i2}(q2) Y^b_{j2}(q2) Z^a_{i3}(q3) Z^b_{j3}(q3)
// with C(q) = alpha a w_q J(q) B(q)^T B(q) (3x3 symmetric) and X^a = l' if a == 0 else l (Y: a == 1, Z: a == 2).
// The three sums are done one direction at a time:
//   1  S1[ab][q1 q2]      = sum_q3 C_ab(q) Z^a_{i3} Z^b_{j3}                       per slice (i3, j3)
//   2  s2[ab][q1](i2, j2) = sum_q2 S1[ab][q1 q2] Y^a_{i2} Y^b_{j2}                 lane = (i2, j2), in registers
//   3  K(i1 i2 i3, j1 j2 j3) += P_x[ab][q1][(i1, j1)] * s2                         25 accumulators per lane
// Step 3 carries 80 % of the flops; its left operand P_x = X^a_{i1}(q1) X^b_{j1}(q1) is the same for every element,
// lane and slice, so it travels as a KERNEL PARAMETER: with the loops unrolled every DFMA takes it straight from the
// constant bank (c[0x0][imm]) -- no register, no shared-memory traffic, the fp64 pipe is the only unit that works.
// (On B200 the fp64 tensor-core peak equals the fp64 FMA peak; DMMA m8n8k4 would need the K = 45 contraction padded to
// 48 and 25 x 25 tiles padded to 32 x 32, i.e. 1.7x the flops of this formulation for the same peak.)
//
// The 1D tables are recovered from the full tables handed over the C ABI (partition of unity: summing phi over the other
// two directions leaves l_{i1}(q1)) and the factorisation is VERIFIED entry by entry; when the tables are not an exact
// tensor product (1e-13) the generic kernel is used instead -- so the result always reproduces the given tables.
#include <cmath>
#include <cstring>

#include "elem_kernel.cuh"

namespace gf {

template <int ND1, int NQ1>
struct SfTables {
  static constexpr int NN = ND1 * ND1;
  double px[9][NQ1][NN];  // [a*3+b][q1][i1 + ND1*j1] = X^a_{i1}(q1) X^b_{j1}(q1)
  double py[4][NN][NQ1];  // [(a==1)*2 + (b==1)][i2 + ND1*j2][q2]
  double pz[4][NN][NQ1];  // [(a==2)*2 + (b==2)][i3 + ND1*j3][q3]
  double l1[ND1][NQ1];    // the 1D basis and its derivative at the 1D points
  double d1[ND1][NQ1];
};

struct SfArgs {
  const double *x, *y, *z;
  const int32_t *conn, *edof;
  const double *U, *w, *gt_grad;
  int64_t e0, ne;
  double coef;  // alpha * a
  const double *gphi;         // hyperelastic kernel: full gradient table [q][i][3] (Grad_u and the residual)
  double lambda, mu, alpha;   // hyperelastic kernel
  int law;                    // GFGPU_SVK / GFGPU_NEOHOOKEAN_*
  double *stage;
  uint16_t *emask;
  double *rstage;
  const uint32_t *slot;  // direct mode (common.cuh ElemArgs): destinations of the element's entries, ascending
  const uint16_t *kloc;  // ... and which entry goes there
  double *pr, *mstage;
  uint32_t nnz32;
};

// K_e is symmetric (C is): only the slices (i3 <= j3) are contracted and only they are STORED -- the shared-memory image
// of K_e is its upper block triangle (ND1 (ND1+1)/2 blocks of NN x NN: 75 KB instead of 125 KB for Q4), which lets TWO CTAs
// share an SM: while one is in its output phase (HBM latency) the other contracts.  kt(i, j) addresses K(i, j) in that image.
template <int ND1>
__device__ __forceinline__ int sf_tri(int a3, int b3) { return a3 * ND1 - (a3 * (a3 - 1)) / 2 + (b3 - a3); }  // a3 <= b3
template <int ND1>
__device__ __forceinline__ int sf_kt(int i, int j) {
  constexpr int NN = ND1 * ND1;
  const int i3 = i / NN, i12 = i - i3 * NN, j3 = j / NN, j12 = j - j3 * NN;
  return i3 <= j3 ? sf_tri<ND1>(i3, j3) * NN * NN + i12 + NN * j12 : sf_tri<ND1>(j3, i3) * NN * NN + j12 + NN * i12;
}

// DMMA: the last contraction on the fp64 tensor core instead (north_star asks for the A/B): per slice
//   acc(m, n) = sum_k px[k][m] s2[k][n],  m = (i1, j1) 25 -> 32,  n = (i2, j2) 25 -> 32,  k = (ab, q1) 45 -> 48,
// as 4 x 4 x 12 mma.sync.m8n8k4.f64; the A fragments (the constant table px, zero padded) sit in shared memory in fragment
// order, the B fragment s2[k][n] is formed by the lane that owns it, accumulators in registers.  1.75x the multiply-adds
// of the FMA formulation for a pipe that peaks 9 % higher (api.cu fp64 probes): measured in profiles/round2_c5_dmma_ab.txt.
template <int ND1, int NQ1, int NW, int OCC, bool DMMA>
__global__ void __launch_bounds__(NW * 32, OCC)
k_sumfact_laplace(const SfArgs a, const __grid_constant__ SfTables<ND1, NQ1> T) {
  constexpr int NN = ND1 * ND1, ND = ND1 * ND1 * ND1, NQ2 = NQ1 * NQ1, NQ = NQ1 * NQ1 * NQ1, NT = NW * 32;
  constexpr int NSL = ND1 * (ND1 + 1) / 2, BL = NN * NN;
  constexpr int KT = (9 * NQ1 + 3) / 4, MT = (NN + 7) / 8;  // k-steps and m / n tiles of the DMMA variant
  extern __shared__ __align__(16) double sm[];
  double *sK = sm;                       // NSL blocks (i3 <= j3) of NN x NN: row (i1,i2) + NN * column (j1,j2)
  double *sC = sK + NSL * BL;            // 6 x NQ : C00 C01 C02 C11 C12 C22
  double *sS1 = sC + 6 * NQ;             // per warp 9 x NQ2
  double *sG = sS1 + NW * 9 * NQ2;       // 3 x 8 node coordinates
  double *sU = sG + 24;                  // ND coefficients
  double *sRed = sU + ND;                // NW
  double *sAf = sRed + NW + (NW & 1);    // DMMA: KT x MT x 32 A fragments
  double *sPy = sAf + (DMMA ? KT * MT * 32 : 0);  // DMMA: 4 x NN x NQ1
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (DMMA) {
    for (int idx = tid; idx < KT * MT * 32; idx += NT) {
      const int ln = idx & 31, mt = (idx >> 5) % MT, kt = idx / (32 * MT);
      const int k = 4 * kt + (ln & 3), m = 8 * mt + (ln >> 2);
      sAf[idx] = (k < 9 * NQ1 && m < NN) ? T.px[k / NQ1][k % NQ1][m] : 0.0;
    }
    for (int idx = tid; idx < 4 * NN * NQ1; idx += NT) sPy[idx] = T.py[idx / (NN * NQ1)][(idx / NQ1) % NN][idx % NQ1];
    __syncthreads();
  }
  // lane = (i2, j2): its 1D factors for the y direction
  double yy[4][NQ1];
  if (lane < NN) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int q = 0; q < NQ1; ++q) yy[c][q] = T.py[c][lane][q];
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int q = 0; q < NQ1; ++q) yy[c][q] = 0.0;
  }
  for (int64_t el = blockIdx.x; el < a.ne; el += gridDim.x) {
    const int64_t e = a.e0 + el;
    // ---- A: gather
    if (tid < 24) {
      const int i = tid / 3, d = tid % 3;
      const int32_t p = a.conn[e * 8 + i];
      sG[d + 3 * i] = (d == 0 ? a.x : d == 1 ? a.y : a.z)[p];
    }
    if (a.rstage)
      for (int i = tid; i < ND; i += NT) sU[i] = a.U ? a.U[a.edof[e * ND + i]] : 0.0;
    __syncthreads();
    // ---- B: metric at the Gauss points
    for (int q = tid; q < NQ; q += NT) {
      double geo[10];
      geometry<3>(sG, a.gt_grad + (size_t)q * 24, 8, geo);
      const double wq = a.w[q];
      const double c = (wq == 0.0) ? 0.0 : a.coef * geo[9] * wq;  // zero-weight points are skipped (C&E.cc:8852)
      int k = 0;
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int r = p; r < 3; ++r) {
          double s2 = 0;
#pragma unroll
          for (int n = 0; n < 3; ++n) s2 += geo[n + 3 * p] * geo[n + 3 * r];
          sC[(k++) * NQ + q] = c * s2;
        }
    }
    __syncthreads();
    // ---- C: slices (i3 <= j3), one per warp and round
    double *S1 = sS1 + warp * 9 * NQ2;
    for (int task = warp; task < NSL; task += NW) {
      int i3 = 0, rem = task;  // task -> (i3 <= j3), rows of the upper triangle
      while (rem >= ND1 - i3) { rem -= ND1 - i3; ++i3; }
      const int j3 = i3 + rem, sl = i3 + ND1 * j3;
      // 1: contract the third direction
      for (int idx = lane; idx < 9 * NQ2; idx += 32) {
        const int ab = idx / NQ2, q12 = idx - ab * NQ2;
        const int aa = ab / 3, bb = ab - aa * 3;
        const int lo = aa < bb ? aa : bb, hi = aa < bb ? bb : aa;
        const double *C = sC + (lo == 0 ? hi : lo == 1 ? 2 + hi : 5) * NQ + q12;
        const double *pz = T.pz[(aa == 2) * 2 + (bb == 2)][sl];
        double s2 = 0;
#pragma unroll
        for (int q3 = 0; q3 < NQ1; ++q3) s2 += C[q3 * NQ2] * pz[q3];
        S1[idx] = s2;
      }
      __syncwarp();
      if (DMMA) {
        double cf[MT][MT][2];  // [m tile][n tile]: C(row = lane >> 2, column = 2 (lane & 3) + {0, 1})
#pragma unroll
        for (int x = 0; x < MT; ++x)
#pragma unroll
          for (int y = 0; y < MT; ++y) cf[x][y][0] = cf[x][y][1] = 0.0;
#pragma unroll 1
        for (int kt = 0; kt < KT; ++kt) {
          const int k = 4 * kt + (lane & 3);
          const bool kok = k < 9 * NQ1;
          const int ab = kok ? k / NQ1 : 0, q1 = kok ? k % NQ1 : 0;
          const int yc = ((ab / 3) == 1) * 2 + ((ab % 3) == 1);
          double s1v[NQ1];
#pragma unroll
          for (int q2 = 0; q2 < NQ1; ++q2) s1v[q2] = S1[ab * NQ2 + q1 + NQ1 * q2];
          double af[MT];
#pragma unroll
          for (int x = 0; x < MT; ++x) af[x] = sAf[(kt * MT + x) * 32 + lane];
#pragma unroll
          for (int y = 0; y < MT; ++y) {
            const int n = 8 * y + (lane >> 2);
            double bf = 0.0;  // B(row = lane & 3, column = lane >> 2) = s2[k][n]
            if (kok && n < NN) {
              const double *pyv = sPy + (yc * NN + n) * NQ1;
#pragma unroll
              for (int q2 = 0; q2 < NQ1; ++q2) bf += s1v[q2] * pyv[q2];
            }
#pragma unroll
            for (int x = 0; x < MT; ++x)
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                           : "+d"(cf[x][y][0]), "+d"(cf[x][y][1])
                           : "d"(af[x]), "d"(bf));
          }
        }
        double *o = sK + task * BL;
#pragma unroll
        for (int x = 0; x < MT; ++x)
#pragma unroll
          for (int y = 0; y < MT; ++y)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int m = 8 * x + (lane >> 2), n = 8 * y + 2 * (lane & 3) + h;
              if (m < NN && n < NN) o[(m % ND1) + ND1 * (n % ND1) + NN * ((m / ND1) + ND1 * (n / ND1))] = cf[x][y][h];
            }
      } else {
      // 2 + 3: contract the second direction in registers, then the first against the constant-bank operand
      double acc[NN];
#pragma unroll
      for (int m = 0; m < NN; ++m) acc[m] = 0.0;
#pragma unroll
      for (int ab = 0; ab < 9; ++ab) {
        const int yc = ((ab / 3) == 1) * 2 + ((ab % 3) == 1);
#pragma unroll
        for (int q1 = 0; q1 < NQ1; ++q1) {
          double s2 = 0;
#pragma unroll
          for (int q2 = 0; q2 < NQ1; ++q2) s2 += S1[ab * NQ2 + q1 + NQ1 * q2] * yy[yc][q2];
#pragma unroll
          for (int m = 0; m < NN; ++m) acc[m] += T.px[ab][q1][m] * s2;
        }
      }
      if (lane < NN) {
        const int i2 = lane % ND1, j2 = lane / ND1;
        double *o = sK + task * BL + ND1 * i2 + NN * (ND1 * j2);
#pragma unroll
        for (int m = 0; m < NN; ++m) o[(m % ND1) + NN * (m / ND1)] = acc[m];
      }
      }
      __syncwarp();
    }
    __syncthreads();
    // ---- element residual r = K_e u_e (before the drop rule, like the quadrature form of the generic kernel)
    if (a.rstage)
      for (int i = tid; i < ND; i += NT) {
        const int i3 = i / NN, i12 = i - i3 * NN;
        double s2 = 0;
        for (int j3 = 0; j3 < ND1; ++j3) {
          const double *Kb = i3 <= j3 ? sK + sf_tri<ND1>(i3, j3) * BL + i12 : sK + sf_tri<ND1>(j3, i3) * BL + NN * i12;
          const int st = i3 <= j3 ? NN : 1;
          for (int j12 = 0; j12 < NN; ++j12) s2 += Kb[st * j12] * sU[j12 + NN * j3];
        }
        a.rstage[(size_t)el * ND + i] = s2;
      }
    // ---- D: drop rule and output (C&E.cc:4889,4898; 5380-5402)
    if (a.stage || a.emask || a.slot) {
      double vmax = 0.0;
      for (int k = tid; k < NSL * BL; k += NT) vmax = fmax(vmax, fabs(sK[k]));
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, off));
      if (lane == 0) sRed[warp] = vmax;
      __syncthreads();
      vmax = 0.0;
#pragma unroll
      for (int wv = 0; wv < NW; ++wv) vmax = fmax(vmax, sRed[wv]);
      const double thr = vmax * 1e-14;
      double *st = a.stage ? a.stage + (size_t)el * ND * ND : nullptr;
      uint16_t *em = a.emask ? a.emask + (size_t)el * ND * ND : nullptr;
      if (a.slot) {
        // direct mode: the pattern is fixed, every entry has its destination -- a CSC slot of pr (single-contribution
        // entries: 77 % on Q4), or a place in the compact stage of the shared entries; no element matrix goes to HBM.
        // Destination order: consecutive threads, consecutive addresses (entries outside the pattern sort last).
        const uint32_t *sl = a.slot + (size_t)el * ND * ND;
        const uint16_t *kl = a.kloc + (size_t)el * ND * ND;
        for (int k0 = tid; k0 < ND * ND; k0 += 8 * NT) {
          uint32_t u[8];
          uint16_t kk[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int k = k0 + r * NT;
            u[r] = k < ND * ND ? __ldcs(sl + k) : 0xffffffffu;
            kk[r] = k < ND * ND ? __ldcs(kl + k) : (uint16_t)0;
          }
#pragma unroll
          for (int r = 0; r < 8; ++r)
            if (u[r] != 0xffffffffu) {
              const int kq = kk[r], j = kq / ND, i = kq - j * ND;
              const double v = sK[sf_kt<ND1>(i, j)];
              const bool keep = (vmax != 0.0) && (fabs(v) > thr);
              if (u[r] < a.nnz32) a.pr[u[r]] = keep ? v : 0.0;
              else a.mstage[u[r] - a.nnz32] = keep ? v : 0.0;
            }
        }
      } else {
        for (int k = tid; k < ND * ND; k += NT) {  // k = i + ND*j: the stage index and the mask index p = j*ND + i
          const int j = k / ND, i = k - j * ND;
          const double v = sK[sf_kt<ND1>(i, j)];
          const bool keep = (vmax != 0.0) && (fabs(v) > thr);
          if (st) st[k] = keep ? v : 0.0;
          if (em) em[k] = keep ? 1 : 0;
        }
      }
    }
    __syncthreads();
  }
}


// ---------------------------------------------------------------- hyperelastic (vector, Q = 3) variant
// K_e(i al, j be) = sum_q sum_{p,r} ghat_i^p(q) A^(al,p,be,r)(q) ghat_j^r(q) with the tangent pulled back to reference
// coordinates, A^(al,p,be,r) = sum_{n,l} B(n,p) D(al,n,be,l) B(l,r) (D from hyper_point, elem_kernel.cuh): nine
// scalar systems of the Laplace shape, one per (al, be), with a full 3x3 metric each.  Lane = (i2, j2, al); the lane
// runs be = 0..2 in sequence (9 accumulators (i1, j1) per be).  BASELINE config 4: FEM_QK(3,2), 64 Gauss points.
template <int ND1, int NQ1, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
k_sumfact_hyper(const SfArgs a, const __grid_constant__ SfTables<ND1, NQ1> T) {
  constexpr int NN = ND1 * ND1, ND = ND1 * ND1 * ND1, S1 = 3 * ND, NQ2 = NQ1 * NQ1, NQ = NQ1 * NQ1 * NQ1, NT = NW * 32;
  constexpr int DS = 82;  // row stride of the per-point tangent
  static_assert(3 * NN <= 32 && NN <= NW * 32, "lane mapping");
  extern __shared__ __align__(16) double sm[];
  double *sK = sm;                    // S1 x S1 column-major: (i*3 + al) + S1 * (j*3 + be)
  double *sD = sK + S1 * S1;          // NQ x DS : D(al,n,be,l) per point
  double *sA = sD + NQ * DS;          // 81 x NQ : A^[(al + 3 be) * 9 + p*3 + r][q]
  double *sS1 = sA + 81 * NQ;         // per warp 3 x 9 x NQ2
  double *sP = sS1 + NW * 27 * NQ2;   // NQ x 9 : P^(al,p) = sum_n P(al,n) B(n,p)
  double *sGeo = sP + 9 * NQ;         // NQ x 10 : B, J per point
  double *sG = sGeo + 10 * NQ;
  double *sU = sG + 24;
  double *sRed = sU + S1;             // NW, then 3 x S1 partial residuals
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ln = lane % NN, lal = lane / NN;  // lane = (i2 + ND1*j2) + NN * al ; lanes >= 3*NN idle
  const bool lact = lane < 3 * NN;
  double yy[4][NQ1];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int q = 0; q < NQ1; ++q) yy[c][q] = lact ? T.py[c][ln][q] : 0.0;
  for (int64_t el = blockIdx.x; el < a.ne; el += gridDim.x) {
    const int64_t e = a.e0 + el;
    if (tid < 24) {
      const int i = tid / 3, d = tid % 3;
      const int32_t p = a.conn[e * 8 + i];
      sG[d + 3 * i] = (d == 0 ? a.x : d == 1 ? a.y : a.z)[p];
    }
    for (int i = tid; i < S1; i += NT) sU[i] = a.U ? a.U[a.edof[e * ND + i / 3] + i % 3] : 0.0;
    __syncthreads();
    // ---- B: material point at every Gauss point (thread = point).  Spreading the nine (l, n) slices of the tangent
    // over threads (hyper_prep / hyper_slice) was measured slower: the shared prologue is most of the work.
    for (int q = tid; q < NQ; q += NT) {
      double *geo = sGeo + q * 10;
      geometry<3>(sG, a.gt_grad + (size_t)q * 24, 8, geo);
      double Gh[9];  // reference gradient of u: Gh(al,p) = sum_i u_i(al) ghat_i^p
#pragma unroll
      for (int k = 0; k < 9; ++k) Gh[k] = 0.0;
      const double *g = a.gphi + (size_t)q * ND * 3;
      for (int i = 0; i < ND; ++i) {
        const double g0 = g[i * 3], g1 = g[i * 3 + 1], g2 = g[i * 3 + 2];
#pragma unroll
        for (int al = 0; al < 3; ++al) {
          const double u = sU[i * 3 + al];
          Gh[al] += u * g0; Gh[al + 3] += u * g1; Gh[al + 6] += u * g2;
        }
      }
      double Gu[9];  // Gu(al,n) = sum_p Gh(al,p) B(n,p)
#pragma unroll
      for (int n = 0; n < 3; ++n)
#pragma unroll
        for (int al = 0; al < 3; ++al) Gu[al + 3 * n] = Gh[al] * geo[n] + Gh[al + 3] * geo[n + 3] + Gh[al + 6] * geo[n + 6];
      const double wq = a.w[q];
      const double coeff = (wq == 0.0) ? 0.0 : a.alpha * geo[9] * wq;  // zero-weight points are skipped (C&E.cc:8852)
      double P[9];
      hyper_point(a.law, Gu, a.lambda, a.mu, coeff, P, sD + (size_t)q * DS);
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int al = 0; al < 3; ++al)
          sP[q * 9 + al + 3 * p] = P[al] * geo[3 * p] + P[al + 3] * geo[1 + 3 * p] + P[al + 6] * geo[2 + 3 * p];
    }
    __syncthreads();
    // ---- B2: pull-back of the tangent: work item = (point, be, r)
    for (int w = tid; w < NQ * 9; w += NT) {
      const int q = w / 9, br = w % 9, be = br / 3, r = br % 3;
      const double *geo = sGeo + q * 10;
      const double *D = sD + (size_t)q * DS;
      double t[9];  // t(al,n) = sum_l D(al,n,be,l) B(l,r)
#pragma unroll
      for (int n = 0; n < 3; ++n)
#pragma unroll
        for (int al = 0; al < 3; ++al) {
          double s = 0;
#pragma unroll
          for (int l = 0; l < 3; ++l) s += D[al + 3 * (n + 3 * (be + 3 * l))] * geo[l + 3 * r];
          t[al + 3 * n] = s;
        }
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int al = 0; al < 3; ++al) {
          const double v = geo[3 * p] * t[al] + geo[1 + 3 * p] * t[al + 3] + geo[2 + 3 * p] * t[al + 6];
          sA[((al + 3 * be) * 9 + p * 3 + r) * NQ + q] = v;
        }
    }
    __syncthreads();
    // ---- C: slices (i3, j3), one per warp
    double *S1w = sS1 + warp * 27 * NQ2;
    for (int sl = warp; sl < NN; sl += NW) {
      double acc[3][NN];
#pragma unroll
      for (int be = 0; be < 3; ++be)
#pragma unroll
        for (int m = 0; m < NN; ++m) acc[be][m] = 0.0;
#pragma unroll
      for (int be = 0; be < 3; ++be) {
        for (int idx = lane; idx < 27 * NQ2; idx += 32) {  // S1[al'][ab][q1 q2] for this be
          const int al2 = idx / (9 * NQ2), rem = idx - al2 * 9 * NQ2, ab = rem / NQ2, q12 = rem - ab * NQ2;
          const double *A = sA + ((al2 + 3 * be) * 9 + ab) * NQ + q12;
          const double *pz = T.pz[((ab / 3) == 2) * 2 + ((ab % 3) == 2)][sl];
          double s = 0;
#pragma unroll
          for (int q3 = 0; q3 < NQ1; ++q3) s += A[q3 * NQ2] * pz[q3];
          S1w[idx] = s;
        }
        __syncwarp();
        const double *S1 = S1w + (lact ? lal : 0) * 9 * NQ2;
#pragma unroll
        for (int ab = 0; ab < 9; ++ab) {
          const int yc = ((ab / 3) == 1) * 2 + ((ab % 3) == 1);
#pragma unroll
          for (int q1 = 0; q1 < NQ1; ++q1) {
            double s2 = 0;
#pragma unroll
            for (int q2 = 0; q2 < NQ1; ++q2) s2 += S1[ab * NQ2 + q1 + NQ1 * q2] * yy[yc][q2];
#pragma unroll
            for (int m = 0; m < NN; ++m) acc[be][m] += T.px[ab][q1][m] * s2;
          }
        }
        __syncwarp();
      }
      if (lact) {
        const int i2 = ln % ND1, j2 = ln / ND1, i3 = sl % ND1, j3 = sl / ND1;
#pragma unroll
        for (int be = 0; be < 3; ++be)
#pragma unroll
          for (int m = 0; m < NN; ++m) {
            const int i = (m % ND1) + ND1 * i2 + NN * i3, j = (m / ND1) + ND1 * j2 + NN * j3;
            sK[(i * 3 + lal) + S1 * (j * 3 + be)] = acc[be][m];
          }
      }
    }
    __syncthreads();
    // ---- element residual r(i al) = sum_q sum_p ghat_i^p(q) P^(al,p)(q)        (C&E.cc:4669-4735)
    // three threads per entry, each a third of the points; the partial sums are added in a fixed order
    if (a.rstage) {
      double *sR = sRed + NW;
      for (int w = tid; w < 3 * S1; w += NT) {
        const int part = w / S1, k = w - part * S1, i = k / 3, al = k % 3;
        const int q0 = part * NQ / 3, q1 = (part + 1) * NQ / 3;
        double s = 0;
        for (int q = q0; q < q1; ++q) {
          const double *g = a.gphi + ((size_t)q * ND + i) * 3;
          s += g[0] * sP[q * 9 + al] + g[1] * sP[q * 9 + al + 3] + g[2] * sP[q * 9 + al + 6];
        }
        sR[w] = s;
      }
      __syncthreads();
      for (int k = tid; k < S1; k += NT) a.rstage[(size_t)el * S1 + k] = (sR[k] + sR[S1 + k]) + sR[2 * S1 + k];
    }
    // ---- D: drop rule and output
    if (a.stage || a.emask) {
      double vmax = 0.0;
      for (int k = tid; k < S1 * S1; k += NT) vmax = fmax(vmax, fabs(sK[k]));
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, off));
      if (lane == 0) sRed[warp] = vmax;
      __syncthreads();
      vmax = 0.0;
#pragma unroll
      for (int wv = 0; wv < NW; ++wv) vmax = fmax(vmax, sRed[wv]);
      const double thr = vmax * 1e-14;
      if (a.stage) {
        double *st = a.stage + (size_t)el * S1 * S1;
        for (int k = tid; k < S1 * S1; k += NT) {
          const double v = sK[k];
          st[k] = ((vmax != 0.0) && (fabs(v) > thr)) ? v : 0.0;
        }
      }
      if (a.emask) {
        uint16_t *em = a.emask + (size_t)el * ND * ND;
        for (int p = tid; p < ND * ND; p += NT) {
          const int j = p / ND, i = p % ND;
          unsigned mask = 0;
#pragma unroll
          for (int bb = 0; bb < 3; ++bb)
#pragma unroll
            for (int aa = 0; aa < 3; ++aa) {
              const double v = sK[(i * 3 + aa) + S1 * (j * 3 + bb)];
              if ((vmax != 0.0) && (fabs(v) > thr)) mask |= 1u << (bb * 3 + aa);
            }
          em[p] = (uint16_t)mask;
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- hyperelastic variant, two elements in flight
// Same mathematics and outputs as k_sumfact_hyper.  The profile of that kernel (profiles/round2_ncu_sumfact_hyper_v1_c4.txt)
// shows one CTA per SM walking through serial phases: 15 % of the warp samples sit at the barrier behind the material-point
// phase (64 threads of 288 busy), the contraction phase is latency bound at 9 warps.  Here the CTA is WARP SPECIALISED:
//   material group (3 warps)  : element n+1 -- gather, material point per Gauss point, pull-back of the tangent into
//                               sA[(n+1) & 1], flux into sP[(n+1) & 1], element residual;
//   contraction group (9 warps): element n  -- slices of the sum-factorised contraction from sA[n & 1], drop rule, output.
// The two groups meet only at named barriers (full[b] / empty[b], bar.arrive + bar.sync, no __syncthreads).
// K_e is symmetric for a hyperelastic law (major symmetry of D): only the slices (i3 <= j3) are contracted, the others
// are their mirror images -- 18 (slice, beta) tasks instead of 27, two per warp.
// Grad_u and the residual use the 1D tables (products of three 1D values) instead of the full gradient table in global
// memory: with 222 KB of shared memory the L1 that served that table is gone.
template <int ND1, int NQ1>
__global__ void __launch_bounds__(384, 1)
k_sumfact_hyper_ws(const SfArgs a, const __grid_constant__ SfTables<ND1, NQ1> T) {
  constexpr int NN = ND1 * ND1, ND = ND1 * ND1 * ND1, S1 = 3 * ND, NQ2 = NQ1 * NQ1, NQ = NQ1 * NQ1 * NQ1;
  constexpr int NWC = 9, NTC = NWC * 32, NTM = 96;
  // strides chosen for the shared-memory banks: lanes = Gauss points read / write sD and sGeo (odd strides), the three
  // components al' of S1 sit 2 doubles apart modulo 16 (they are read by the three lane groups of a warp at once)
  constexpr int DS = 81, GS = 11, SS = 9 * NQ2 + 2;
  constexpr int NSL = ND1 * (ND1 + 1) / 2, NTASK = NSL * 3;
  static_assert(ND1 == 3 && 3 * NN <= 32 && NQ2 <= 32, "lane mapping");
  extern __shared__ __align__(16) double sm[];
  double *sK = sm;                          // S1 x S1 column-major
  double *sA = sK + S1 * S1;                // 2 x 81 x NQ
  double *sD = sA + 2 * 81 * NQ;            // NQ x DS
  double *sS1 = sD + NQ * DS;               // per contraction warp 3 x SS
  double *sP = sS1 + NWC * 3 * SS;          // 2 x NQ x 9
  double *sGeo = sP + 2 * 9 * NQ;           // NQ x GS
  double *sG = sGeo + GS * NQ;              // 24
  double *sU = sG + 24;                     // S1
  double *sRed = sU + S1;                   // NWC
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // named barriers: 1 material group, 2 contraction group, 3 + b full[b], 5 + b empty[b]
  auto bar_sync = [](int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); };
  auto bar_arrive = [](int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); };

  if (warp >= NWC) {
    // ================= material group =================
    const int mt = tid - NTC;
    int it = 0;
    for (int64_t el = blockIdx.x; el < a.ne; el += gridDim.x, ++it) {
      const int b = it & 1;
      const int64_t e = a.e0 + el;
      double *sAb = sA + b * 81 * NQ, *sPb = sP + b * 9 * NQ;
      if (mt < 24) {
        const int i = mt / 3, d = mt % 3;
        const int32_t p = a.conn[e * 8 + i];
        sG[d + 3 * i] = (d == 0 ? a.x : d == 1 ? a.y : a.z)[p];
      }
      for (int i = mt; i < S1; i += NTM) sU[i] = a.U ? a.U[a.edof[e * ND + i / 3] + i % 3] : 0.0;
      bar_sync(5 + b, 384);  // the contraction group has finished with sA[b], sP[b] (element it - 2)
      bar_sync(1, NTM);
      for (int q = mt; q < NQ; q += NTM) {
        double *geo = sGeo + q * GS;
        geometry<3>(sG, a.gt_grad + (size_t)q * 24, 8, geo);
        const int q1 = q % NQ1, q2 = (q / NQ1) % NQ1, q3 = q / NQ2;
        double lx[ND1], dx[ND1];
#pragma unroll
        for (int i = 0; i < ND1; ++i) { lx[i] = T.l1[i][q1]; dx[i] = T.d1[i][q1]; }
        double Gh[9];  // reference gradient of u: Gh(al,p) = sum_i u_i(al) ghat_i^p
#pragma unroll
        for (int k = 0; k < 9; ++k) Gh[k] = 0.0;
#pragma unroll
        for (int i3 = 0; i3 < ND1; ++i3) {
          const double lz = T.l1[i3][q3], dz = T.d1[i3][q3];
#pragma unroll
          for (int i2 = 0; i2 < ND1; ++i2) {
            const double ly = T.l1[i2][q2], dy = T.d1[i2][q2];
            const double w0 = ly * lz, w1 = dy * lz, w2 = ly * dz;
            const double *u = sU + (ND1 * i2 + NN * i3) * 3;
#pragma unroll
            for (int al = 0; al < 3; ++al) {
              double t0 = 0, t1 = 0;
#pragma unroll
              for (int i1 = 0; i1 < ND1; ++i1) { t0 += u[i1 * 3 + al] * dx[i1]; t1 += u[i1 * 3 + al] * lx[i1]; }
              Gh[al] += t0 * w0; Gh[al + 3] += t1 * w1; Gh[al + 6] += t1 * w2;
            }
          }
        }
        double Gu[9];
#pragma unroll
        for (int n = 0; n < 3; ++n)
#pragma unroll
          for (int al = 0; al < 3; ++al) Gu[al + 3 * n] = Gh[al] * geo[n] + Gh[al + 3] * geo[n + 3] + Gh[al + 6] * geo[n + 6];
        const double wq = a.w[q];
        const double coeff = (wq == 0.0) ? 0.0 : a.alpha * geo[9] * wq;  // zero-weight points are skipped (C&E.cc:8852)
        double P[9];
        hyper_point(a.law, Gu, a.lambda, a.mu, coeff, P, sD + (size_t)q * DS);
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
          for (int al = 0; al < 3; ++al)
            sPb[q * 9 + al + 3 * p] = P[al] * geo[3 * p] + P[al + 3] * geo[1 + 3 * p] + P[al + 6] * geo[2 + 3 * p];
      }
      bar_sync(1, NTM);
      for (int w = mt; w < NQ * 9; w += NTM) {  // pull-back of the tangent: work item = (point, be, r)
        const int q = w % NQ, br = w / NQ, be = br / 3, r = br % 3;  // lanes = consecutive points
        const double *geo = sGeo + q * GS;
        const double *D = sD + (size_t)q * DS;
        double t[9];
#pragma unroll
        for (int n = 0; n < 3; ++n)
#pragma unroll
          for (int al = 0; al < 3; ++al) {
            double s2 = 0;
#pragma unroll
            for (int l = 0; l < 3; ++l) s2 += D[al + 3 * (n + 3 * (be + 3 * l))] * geo[l + 3 * r];
            t[al + 3 * n] = s2;
          }
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
          for (int al = 0; al < 3; ++al) {
            const double v = geo[3 * p] * t[al] + geo[1 + 3 * p] * t[al + 3] + geo[2 + 3 * p] * t[al + 6];
            sAb[((al + 3 * be) * 9 + p * 3 + r) * NQ + q] = v;
          }
      }
      __threadfence_block();
      bar_arrive(3 + b, 384);  // sA[b] is complete (sP[b] was complete at the barrier above)
      // element residual r(i al) = sum_q sum_p ghat_i^p(q) P^(al,p)(q)   (C&E.cc:4669-4735), thread per entry
      if (a.rstage)
        for (int k = mt; k < S1; k += NTM) {
          const int i = k / 3, al = k % 3, i1 = i % ND1, i2 = (i / ND1) % ND1, i3 = i / NN;
          double acc0 = 0, acc1 = 0;  // two interleaved chains (even / odd planes), added in a fixed order
          for (int q3 = 0; q3 < NQ1; ++q3) {
            const double lz = T.l1[i3][q3], dz = T.d1[i3][q3];
            double sq = 0;
            for (int q2 = 0; q2 < NQ1; ++q2) {
              const double ly = T.l1[i2][q2], dy = T.d1[i2][q2];
              const double w0 = ly * lz, w1 = dy * lz, w2 = ly * dz;
              const double *Pq = sPb + (size_t)(NQ1 * q2 + NQ2 * q3) * 9 + al;
#pragma unroll
              for (int q1 = 0; q1 < NQ1; ++q1)
                sq += (T.d1[i1][q1] * w0) * Pq[q1 * 9] + (T.l1[i1][q1] * w1) * Pq[q1 * 9 + 3] + (T.l1[i1][q1] * w2) * Pq[q1 * 9 + 6];
            }
            if (q3 & 1) acc1 += sq; else acc0 += sq;
          }
          a.rstage[(size_t)el * S1 + k] = acc0 + acc1;
        }
    }
    // the contraction group arrived once more than this group waited on each empty barrier: settle the balance
    bar_sync(5 + (it & 1), 384);
    bar_sync(5 + ((it + 1) & 1), 384);
    return;
  }

  // ================= contraction group =================
  const int ln = lane % NN, lal = lane / NN;  // lane = (i2 + ND1*j2) + NN * al ; lanes >= 3*NN idle
  const bool lact = lane < 3 * NN;
  double yy[4][NQ1];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int q = 0; q < NQ1; ++q) yy[c][q] = lact ? T.py[c][ln][q] : 0.0;
  bar_arrive(5, 384);  // both buffers start empty
  bar_arrive(6, 384);
  double *S1w = sS1 + warp * 3 * SS;
  int it = 0;
  for (int64_t el = blockIdx.x; el < a.ne; el += gridDim.x, ++it) {
    const int b = it & 1;
    const double *sAb = sA + b * 81 * NQ;
    bar_sync(3 + b, 384);
    for (int task = warp; task < NTASK; task += NWC) {
      const int be = task % 3, sidx = task / 3;
      // slice (i3 <= j3): sidx 0..5 -> (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
      const int i3 = sidx < 3 ? 0 : sidx < 5 ? 1 : 2, j3 = sidx < 3 ? sidx : sidx < 5 ? sidx - 2 : 2;
      const int sl = i3 + ND1 * j3;
      double pzr[4][NQ1];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int q = 0; q < NQ1; ++q) pzr[c][q] = T.pz[c][sl][q];
      // 1: contract the third direction; lane = (q1, q2), the 27 (al', ab) combinations unrolled
      if (lane < NQ2) {
#pragma unroll
        for (int al2 = 0; al2 < 3; ++al2)
#pragma unroll
          for (int ab = 0; ab < 9; ++ab) {
            const double *A = sAb + ((al2 + 3 * be) * 9 + ab) * NQ + lane;
            const int cz = ((ab / 3) == 2) * 2 + ((ab % 3) == 2);
            double s2 = 0;
#pragma unroll
            for (int q3 = 0; q3 < NQ1; ++q3) s2 += A[q3 * NQ2] * pzr[cz][q3];
            S1w[al2 * SS + ab * NQ2 + lane] = s2;
          }
      }
      __syncwarp();
      // 2 + 3: second direction in registers, first against the constant-bank operand
      double acc[NN];
#pragma unroll
      for (int m = 0; m < NN; ++m) acc[m] = 0.0;
      const double *S1l = S1w + (lact ? lal : 0) * SS;
#pragma unroll
      for (int ab = 0; ab < 9; ++ab) {
        const int yc = ((ab / 3) == 1) * 2 + ((ab % 3) == 1);
#pragma unroll
        for (int q1 = 0; q1 < NQ1; ++q1) {
          double s2 = 0;
#pragma unroll
          for (int q2 = 0; q2 < NQ1; ++q2) s2 += S1l[ab * NQ2 + q1 + NQ1 * q2] * yy[yc][q2];
#pragma unroll
          for (int m = 0; m < NN; ++m) acc[m] += T.px[ab][q1][m] * s2;
        }
      }
      __syncwarp();
      if (lact) {
        const int i2 = ln % ND1, j2 = ln / ND1;
#pragma unroll
        for (int m = 0; m < NN; ++m) {
          const int i = (m % ND1) + ND1 * i2 + NN * i3, j = (m / ND1) + ND1 * j2 + NN * j3;
          sK[(i * 3 + lal) + S1 * (j * 3 + be)] = acc[m];
          if (i3 != j3) sK[(j * 3 + be) + S1 * (i * 3 + lal)] = acc[m];  // the mirror slice (j3, i3)
        }
      }
    }
    __threadfence_block();
    bar_arrive(5 + b, 384);  // sA[b] may be refilled
    bar_sync(2, NTC);
    // ---- drop rule and output (C&E.cc:4889,4898; 5380-5402)
    if (a.stage || a.emask) {
      double vmax = 0.0;
      for (int k = tid; k < S1 * S1; k += NTC) vmax = fmax(vmax, fabs(sK[k]));
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, off));
      if (lane == 0) sRed[warp] = vmax;
      bar_sync(2, NTC);
      vmax = 0.0;
#pragma unroll
      for (int wv = 0; wv < NWC; ++wv) vmax = fmax(vmax, sRed[wv]);
      const double thr = vmax * 1e-14;
      if (a.stage) {
        double *st = a.stage + (size_t)el * S1 * S1;
        for (int k = tid; k < S1 * S1; k += NTC) {
          const double v = sK[k];
          st[k] = ((vmax != 0.0) && (fabs(v) > thr)) ? v : 0.0;
        }
      }
      if (a.emask) {
        uint16_t *em = a.emask + (size_t)el * ND * ND;
        for (int p = tid; p < ND * ND; p += NTC) {
          const int j = p / ND, i = p % ND;
          unsigned mask = 0;
#pragma unroll
          for (int bb = 0; bb < 3; ++bb)
#pragma unroll
            for (int aa = 0; aa < 3; ++aa) {
              const double v = sK[(i * 3 + aa) + S1 * (j * 3 + bb)];
              if ((vmax != 0.0) && (fabs(v) > thr)) mask |= 1u << (bb * 3 + aa);
            }
          em[p] = (uint16_t)mask;
        }
      }
    }
    bar_sync(2, NTC);  // sK and sRed are free for the next element
  }
}

// ---------------------------------------------------------------- host: factorise and verify the tables
struct SfHost {
  int nd1 = 0, nq1 = 0;
  bool ok = false;
  std::vector<double> l, d;  // nd1 x nq1
};

static int icbrt(int v) {
  int r = (int)std::lround(std::cbrt((double)v));
  return r * r * r == v ? r : 0;
}

static SfHost factorise(const gfgpu_tables *tab) {
  SfHost h;
  if (tab->dim != 3 || tab->ng != 8) return h;
  const int nd1 = icbrt(tab->nd), nq1 = icbrt(tab->nq);
  if (nd1 < 2 || nq1 < 1) return h;
  const int nd = tab->nd, nq = tab->nq, nn = nd1 * nd1;
  const std::vector<double> &phi = tab->h_phi, &g = tab->h_gphi;
  h.nd1 = nd1; h.nq1 = nq1;
  h.l.assign((size_t)nd1 * nq1, 0.0);
  h.d.assign((size_t)nd1 * nq1, 0.0);
  for (int q1 = 0; q1 < nq1; ++q1)  // q = (q1, 0, 0): sum over (i2, i3) by partition of unity
    for (int i1 = 0; i1 < nd1; ++i1) {
      double sl = 0, sd = 0;
      for (int r = 0; r < nn; ++r) {
        const int i = i1 + nd1 * r;
        sl += phi[(size_t)q1 * nd + i];
        sd += g[((size_t)q1 * nd + i) * 3 + 0];
      }
      h.l[(size_t)i1 * nq1 + q1] = sl;
      h.d[(size_t)i1 * nq1 + q1] = sd;
    }
  double err = 0, ref = 0;
  for (int q = 0; q < nq; ++q) {
    const int q1 = q % nq1, q2 = (q / nq1) % nq1, q3 = q / (nq1 * nq1);
    for (int i = 0; i < nd; ++i) {
      const int i1 = i % nd1, i2 = (i / nd1) % nd1, i3 = i / nn;
      const double l1 = h.l[(size_t)i1 * nq1 + q1], l2 = h.l[(size_t)i2 * nq1 + q2], l3 = h.l[(size_t)i3 * nq1 + q3];
      const double d1 = h.d[(size_t)i1 * nq1 + q1], d2 = h.d[(size_t)i2 * nq1 + q2], d3 = h.d[(size_t)i3 * nq1 + q3];
      const double v[4] = {l1 * l2 * l3, d1 * l2 * l3, l1 * d2 * l3, l1 * l2 * d3};
      const double t[4] = {phi[(size_t)q * nd + i], g[((size_t)q * nd + i) * 3], g[((size_t)q * nd + i) * 3 + 1],
                           g[((size_t)q * nd + i) * 3 + 2]};
      for (int k = 0; k < 4; ++k) {
        err = std::max(err, std::fabs(v[k] - t[k]));
        ref = std::max(ref, std::fabs(t[k]));
      }
    }
  }
  h.ok = ref > 0 && err <= 1e-13 * ref;
  return h;
}

template <int ND1, int NQ1>
static void fill_tables(const SfHost &h, SfTables<ND1, NQ1> &T) {
  constexpr int NN = ND1 * ND1;
  auto L = [&](int i, int q) { return h.l[(size_t)i * NQ1 + q]; };
  auto D = [&](int i, int q) { return h.d[(size_t)i * NQ1 + q]; };
  for (int aa = 0; aa < 3; ++aa)
    for (int bb = 0; bb < 3; ++bb)
      for (int q = 0; q < NQ1; ++q)
        for (int m = 0; m < NN; ++m) {
          const int i1 = m % ND1, j1 = m / ND1;
          T.px[aa * 3 + bb][q][m] = (aa == 0 ? D(i1, q) : L(i1, q)) * (bb == 0 ? D(j1, q) : L(j1, q));
        }
  for (int c = 0; c < 4; ++c)
    for (int m = 0; m < NN; ++m)
      for (int q = 0; q < NQ1; ++q) {
        const int i = m % ND1, j = m / ND1;
        const double v = ((c & 2) ? D(i, q) : L(i, q)) * ((c & 1) ? D(j, q) : L(j, q));
        T.py[c][m][q] = v;
        T.pz[c][m][q] = v;
      }
  for (int i = 0; i < ND1; ++i)
    for (int q = 0; q < NQ1; ++q) { T.l1[i][q] = L(i, q); T.d1[i][q] = D(i, q); }
}

template <int ND1, int NQ1, int NW, int OCC, bool DMMA>
static void launch_sf(gfgpu_ctx *ctx, const SfHost &h, const SfArgs &a) {
  constexpr int ND = ND1 * ND1 * ND1, NN = ND1 * ND1, NQ2 = NQ1 * NQ1, NQ = NQ2 * NQ1;
  constexpr int KT = (9 * NQ1 + 3) / 4, MT = (NN + 7) / 8;
  std::unique_ptr<SfTables<ND1, NQ1>> Tp(new SfTables<ND1, NQ1>);  // filled per call (host side, cheap)
  fill_tables<ND1, NQ1>(h, *Tp);
  static_assert(sizeof(SfTables<ND1, NQ1>) + sizeof(SfArgs) <= 32000, "tables exceed the kernel parameter space");
  const size_t smem = ((size_t)(ND1 * (ND1 + 1) / 2) * NN * NN + 6 * NQ + (size_t)NW * 9 * NQ2 + 24 + ND + NW + 2 +
                       (DMMA ? KT * MT * 32 + 4 * NN * NQ1 : 0)) * 8;
  auto kern = k_sumfact_laplace<ND1, NQ1, NW, OCC, DMMA>;
  GF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(a.ne, OCC * (int64_t)ctx->sm_count));
  kern<<<grid, NW * 32, smem, ctx->stream>>>(a, *Tp);
  GF_LAUNCH_CHECK();
}

template <int ND1, int NQ1, int NW>
static void launch_sf_hyper(gfgpu_ctx *ctx, const SfHost &h, const SfArgs &a) {
  constexpr int ND = ND1 * ND1 * ND1, S1 = 3 * ND, NQ2 = NQ1 * NQ1, NQ = NQ2 * NQ1;
  std::unique_ptr<SfTables<ND1, NQ1>> Tp(new SfTables<ND1, NQ1>);
  fill_tables<ND1, NQ1>(h, *Tp);
  static const bool v1 = getenv("GFGPU_SF_HYPER_V1") != nullptr;  // the serial-phase kernel, kept for the A/B in profiles/
  if (!v1) {
    const size_t smem2 = ((size_t)S1 * S1 + 2 * 81 * (size_t)NQ + (size_t)NQ * 81 + 9 * 3 * (9 * (size_t)NQ2 + 2) + 2 * 9 * (size_t)NQ +
                          11 * (size_t)NQ + 24 + S1 + 9 + 2) * 8;
    auto kern2 = k_sumfact_hyper_ws<ND1, NQ1>;
    GF_CUDA(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    const int grid2 = (int)std::max<int64_t>(1, std::min<int64_t>(a.ne, ctx->sm_count));
    kern2<<<grid2, 384, smem2, ctx->stream>>>(a, *Tp);
    GF_LAUNCH_CHECK();
    return;
  }
  const size_t smem = ((size_t)S1 * S1 + (size_t)NQ * 82 + 81 * (size_t)NQ + (size_t)NW * 27 * NQ2 + 19 * (size_t)NQ + 24 + S1 + NW +
                       3 * S1 + 2) * 8;
  auto kern = k_sumfact_hyper<ND1, NQ1, NW>;
  GF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(a.ne, 2 * (int64_t)ctx->sm_count));  // two CTAs per SM
  kern<<<grid, NW * 32, smem, ctx->stream>>>(a, *Tp);
  GF_LAUNCH_CHECK();
}

static int sumfact_shape(int dim, int Q, int nd, bool affine, const ElemArgs &ea) {
  if (dim != 3 || affine || ea.ng != 8 || getenv("GFGPU_NO_SUMFACT")) return 0;
  const bool lap = ea.family == GFGPU_LAPLACE && Q == 1 && ((nd == 125 && ea.nq == 125) || (nd == 64 && ea.nq == 64));
  const bool hyp = (ea.family == GFGPU_SVK || ea.family == GFGPU_NEOHOOKEAN_CIARLET || ea.family == GFGPU_NEOHOOKEAN_BONET) &&
                   Q == 3 && nd == 27 && (ea.nq == 64 || ea.nq == 27);
  return lap ? 1 : hyp ? 2 : 0;
}

int sumfact_kind(const gfgpu_tables *tab, int dim, int Q, int nd, bool affine, const ElemArgs &ea) {
  const int k = sumfact_shape(dim, Q, nd, affine, ea);
  if (!k) return 0;
  return factorise(tab).ok ? k : 0;
}

// returns false when this family / element / table set is not handled here (the caller then uses the generic kernel)
bool launch_sumfact_kernel(gfgpu_ctx *ctx, const gfgpu_tables *tab, int dim, int Q, int nd, bool affine, const ElemArgs &ea) {
  const int kind = sumfact_shape(dim, Q, nd, affine, ea);
  if (!kind) return false;
  const bool lap = kind == 1, hyp = kind == 2;
  const SfHost h = factorise(tab);
  if (!h.ok) return false;
  SfArgs a;
  a.x = ea.x; a.y = ea.y; a.z = ea.z; a.conn = ea.conn; a.edof = ea.edof; a.U = ea.U; a.w = ea.w; a.gt_grad = ea.gt_grad;
  a.e0 = ea.e0; a.ne = ea.e1 - ea.e0;
  a.coef = ea.alpha * ea.par[0];
  a.gphi = ea.gphi; a.lambda = ea.par[0]; a.mu = ea.par[1]; a.alpha = ea.alpha; a.law = ea.family;
  a.stage = ea.stage; a.emask = ea.emask; a.rstage = ea.rstage;
  a.slot = lap ? ea.slot : nullptr; a.kloc = ea.kloc; a.pr = ea.pr; a.mstage = ea.mstage; a.nnz32 = ea.nnz32;
  GF_REQUIRE(!ea.slot || lap, "the direct mode is a feature of the scalar sum-factorised kernel");
  if (a.ne <= 0) return true;
  // GFGPU_SF_VARIANT (A/B of profiles/round2_c5_dmma_ab.txt): 0 = FMA pipe, two CTAs of 8 warps per SM (default);
  // 1 = FMA pipe, one CTA of 15 warps; 2 = fp64 tensor core (mma.sync.m8n8k4), one CTA of 15 warps; 3 = DMMA, 2 x 8 warps
  const int variant = getenv("GFGPU_SF_VARIANT") ? atoi(getenv("GFGPU_SF_VARIANT")) : 0;
  if (lap && h.nd1 == 5 && h.nq1 == 5) {  // 15 slices (i3 <= j3)
    if (variant == 1) launch_sf<5, 5, 15, 1, false>(ctx, h, a);
    else if (variant == 2) launch_sf<5, 5, 15, 1, true>(ctx, h, a);
    else if (variant == 3) launch_sf<5, 5, 8, 2, true>(ctx, h, a);
    else launch_sf<5, 5, 8, 2, false>(ctx, h, a);
  } else if (lap && h.nd1 == 4 && h.nq1 == 4) launch_sf<4, 4, 5, 2, false>(ctx, h, a);
  else if (hyp && h.nd1 == 3 && h.nq1 == 4) launch_sf_hyper<3, 4, 9>(ctx, h, a);
  else if (hyp && h.nd1 == 3 && h.nq1 == 3) launch_sf_hyper<3, 3, 9>(ctx, h, a);
  else return false;
  return true;
}

}  // namespace gf
