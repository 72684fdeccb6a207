// Two-stage tridiagonalisation for the GCV fit (fields::Tps -> Krig.engine.default -> eigen(Q2' K Q2), V73:722, 751;
// SURVEY.md 3.2) - the BLAS-3 alternative to the one-stage k_sytrd of sytrd.cu, selected with
// mb_set_param("sytrd_mode", 3).  tools/proto_two_stage.py is the numpy statement of the same algorithm, block for block.
//
// Stage 1  dense -> band with kBw = 32 sub-diagonals.  Per panel of 32 columns:
//   k_sbr_qr   Householder QR of the block below the band (rows in shared memory, 512 per CTA, ONE fused grid-wide
//              reduction per column: |x|^2, v'P and the V'V column that builds T), compact WY factors V, T
//   k_sbr_av   Z0 = A22 V               (register-tiled FP64, split over k)
//   k_sbr_vtz  G0 = V'Z0, gz = V'z      (per-CTA partials, fixed summation order)
//   k_sbr_w    S = T'G0 T,  W = Z0 T - 1/2 V S,  z <- z - V T' gz
//   k_sbr_r2k  A22 -= V W' + W V'       (register-tiled FP64, 128 x 128 tiles)
//   Both triangles of the trailing matrix are kept up to date (first version: plain products, no symmetric tile logic).
// Stage 2  band -> tridiagonal by bulge chasing (Lang's scheme, one reflector per sweep and step).  k_sbr_chase: one warp per
//   sweep, sweeps handed out by a ticket counter; sweep s runs step k once sweep s-1 has published step k+1 (release /
//   acquire flag per sweep).  Blocks live in shared memory, lane = row or lane = column, so every access is conflict-free.
// No atomics on data, fixed summation orders: the result is deterministic.
#include "common.cuh"
#include "internal.h"

#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>

namespace mb {

namespace cg = cooperative_groups;

namespace {

constexpr int kBw = 32;            // sub-diagonals of the band
constexpr int kLdb = 2 * kBw;      // rows of the band storage: B[off + kLdb * j] = A[j + off, j], off < 2 kBw (room for the bulge)
constexpr int kQrRows = 512;       // rows (= threads) per CTA of the panel QR
constexpr int kPad = 33;

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// grid barrier (same protocol as sytrd.cu): arrivals on a counter, release through a flag on its own line
__device__ __forceinline__ void sbr_grid_sync(unsigned* bar, unsigned& gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++gen;
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(bar) : "memory");
    if (old == gen * gridDim.x - 1) {
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 32), "r"(gen) : "memory");
    } else {
      unsigned seen;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar + 32) : "memory");
      } while ((int)(seen - gen) < 0);
    }
  }
  __syncthreads();
}

// dlarfg: x = [alpha; x1], |x1|^2 = xn2  ->  H x = beta e1, v = [1; x1 * scale]
// (r3i: reciprocal square root / reciprocal from the MUFU seeds with one third-order correction instead of sqrt() and the two
// divisions (~50 instructions fewer), together with the 32-term dot products of the bulge chase as trees instead of four 8-deep
// FMA chains (~36 more), changed NOTHING - chase 33.9 against 33.7 ms, profiles/r3i_bench_tpsonly_fast_house.json: a step of
// those kernels is bound by the instruction count of its busiest warp and by the hand-over, not by the FP64 latency of these
// chains.  Removed; the seeds alone remain a candidate)
__device__ __forceinline__ void make_house(double alpha, double xn2, double& beta, double& tau, double& scale) {
  beta = alpha; tau = 0.0; scale = 0.0;
  if (xn2 != 0.0) {
    beta = -copysign(sqrt(fma(alpha, alpha, xn2)), alpha);
    tau = (beta - alpha) / beta;
    scale = 1.0 / (alpha - beta);
  }
}

// ---------------------------------------------------------------------------------------------
// lower -> upper mirror (the products below read the trailing matrix as a plain square)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sbr_mirror(double* __restrict__ A, int ld, int m) {
  __shared__ double t[32][kPad];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bi < bj) return;
  const int r0 = bi * 32, c0 = bj * 32, x = threadIdx.x & 31, y = threadIdx.x >> 5;
  for (int q = y; q < 32; q += 8) {
    const int i = r0 + x, j = c0 + q;
    t[q][x] = (i < m && j < m) ? A[i + (size_t)j * ld] : 0.0;
  }
  __syncthreads();
  for (int q = y; q < 32; q += 8) {
    const int ii = r0 + q, jj = c0 + x;          // lower element (ii, jj) = t[x][q] goes to (jj, ii)
    if (ii < m && jj < m && ii > jj) A[jj + (size_t)ii * ld] = t[x][q];
  }
}

// ---------------------------------------------------------------------------------------------
// panel QR
// ---------------------------------------------------------------------------------------------
struct QrArgs {
  double* A; int ld; int m; int j0;
  double* V; int ldv;              // ldv x 32, column-major; row 0 = matrix row j0 + kBw; rows >= r are written as zero
  double* T;                       // 32 x 32, row-major, upper triangular
  double* slots;                   // grid version: 2 x (G + 1) x 32 per-CTA partial sums (+ the pivot row), by column parity
  unsigned* bar;                   // grid version: barrier words
};
// X | red | s | piv | T | part (cluster version: [parity][partial sums 32 | pivot row 32], read by the peers over DSMEM)
constexpr size_t kQrSmem = sizeof(double) * (kQrRows * kPad + 16 * 32 + 32 + 32 + 32 * kPad + 2 * 64);

// kCluster: the CTAs form ONE thread-block cluster; the per-column reduction goes through distributed shared memory and the
// hardware cluster barrier.  Otherwise (more than 16 x 512 rows, or no cluster of that size schedulable): global slots and the
// software grid barrier (cooperative launch).
template <bool kCluster>
__global__ void __launch_bounds__(kQrRows, 1) k_sbr_qr(QrArgs a) {
  extern __shared__ double sm_qr[];
  double* X = sm_qr;                      // [512][33]   row t = this thread's row of the panel
  double* red = X + kQrRows * kPad;       // [16][32]
  double* s_sh = red + 16 * 32;           // [32]
  double* piv_sh = s_sh + 32;             // [32]
  double* Tsh = piv_sh + 32;              // [32][33]
  double* part = Tsh + 32 * kPad;         // [2][64]
  cg::cluster_group cluster = cg::this_cluster();
  const int t = threadIdx.x, blk = blockIdx.x, G = gridDim.x;
  const int r = a.m - a.j0 - kBw;
  const int gi = blk * kQrRows + t;
  const size_t row = (size_t)(a.j0 + kBw) + gi;
#pragma unroll 8
  for (int c = 0; c < 32; ++c) X[t * kPad + c] = gi < r ? a.A[row + (size_t)(a.j0 + c) * a.ld] : 0.0;
  for (int i = t; i < 32 * kPad; i += kQrRows) Tsh[i] = 0.0;
  __syncthreads();
  const int nref = min(kBw, r - 1);
  unsigned gen = 0;
  const int c = t & 31, seg = t >> 5;
  for (int j = 0; j < nref; ++j) {
    // ---- s_c = sum over rows i > j of x_i[j] x_i[c] ----------------------------------------------------
    double acc0 = 0.0, acc1 = 0.0;
    const int rbase = seg * 32;
    const int first = j + 1 - blk * kQrRows;        // local rows >= first take part
#pragma unroll 8
    for (int rr = 0; rr < 32; rr += 2) {
      const int rw = rbase + rr;
      const double a0 = X[rw * kPad + j] * X[rw * kPad + c], a1 = X[(rw + 1) * kPad + j] * X[(rw + 1) * kPad + c];
      if (rw >= first) acc0 += a0;
      if (rw + 1 >= first) acc1 += a1;
    }
    red[seg * 32 + c] = acc0 + acc1;
    __syncthreads();
    double s = 0.0, piv = 0.0;
    if (kCluster) {
      double* mine = part + (j & 1) * 64;
      if (t < 32) {
        double p = 0.0;
#pragma unroll
        for (int q = 0; q < 16; ++q) p += red[q * 32 + t];
        mine[t] = p;
        if (blk == 0) mine[32 + t] = X[j * kPad + t];          // the pivot row
      }
      cluster.sync();
      if (t < 32) {
        double pr[16];                                         // all remote loads in flight before the first add
#pragma unroll
        for (int q = 0; q < 16; ++q) pr[q] = q < G ? cluster.map_shared_rank(mine, q)[t] : 0.0;
        piv = cluster.map_shared_rank(mine, 0)[32 + t];
#pragma unroll
        for (int q = 0; q < 16; ++q) s += pr[q];
      }
    } else {
      double* slot = a.slots + (size_t)(j & 1) * (G + 1) * 32;
      if (t < 32) {
        double p = 0.0;
#pragma unroll
        for (int q = 0; q < 16; ++q) p += red[q * 32 + t];
        slot[blk * 32 + t] = p;
        if (blk == 0) slot[G * 32 + t] = X[j * kPad + t];
      }
      sbr_grid_sync(a.bar, gen);
      if (t < 32) {
        for (int q = 0; q < G; ++q) s += __ldcg(slot + q * 32 + t);
        piv = __ldcg(slot + G * 32 + t);
      }
    }
    if (t < 32) {
      // warp 0: the reflector, w_c = v'P[:, c] = piv_c + scale s_c (for c < j the same expression is V[:, c]'v_j)
      double beta, tau, scale;
      make_house(__shfl_sync(0xffffffffu, piv, j), __shfl_sync(0xffffffffu, s, j), beta, tau, scale);
      s_sh[t] = fma(scale, s, piv);
      if (t == 0) { piv_sh[0] = beta; piv_sh[1] = tau; piv_sh[2] = scale; }
    }
    __syncthreads();
    const double beta = piv_sh[0], tau = piv_sh[1], scale = piv_sh[2];
    // ---- column j of T (CTA 0, warp 1): T[:j, j] = -tau T[:j, :j] (V[:, :j]' v_j) --------------------------
    if (blk == 0 && t >= 32 && t < 64) {
      const int q = t - 32;
      if (q < j) {
        double sum = 0.0;
        for (int p = q; p < j; ++p) sum = fma(Tsh[q * kPad + p], s_sh[p], sum);
        Tsh[q * kPad + j] = -tau * sum;
      } else if (q == j) {
        Tsh[j * kPad + j] = tau;
      }
    }
    // ---- apply H_j to the remaining columns of this thread's row ------------------------------------------
    if (gi > j) {
      const double vj = X[t * kPad + j] * scale;
      const double f = tau * vj;
      for (int cc = j + 1; cc < 32; ++cc) X[t * kPad + cc] = fma(-f, s_sh[cc], X[t * kPad + cc]);
      X[t * kPad + j] = vj;
    } else if (gi == j) {
      for (int cc = j + 1; cc < 32; ++cc) X[t * kPad + cc] = fma(-tau, s_sh[cc], X[t * kPad + cc]);
      X[t * kPad + j] = beta;
    }
    __syncthreads();
  }
  // ---- V with explicit unit diagonal; rows r .. ldv-1 zero ------------------------------------------------------
  if (gi < a.ldv) {
#pragma unroll 8
    for (int q = 0; q < 32; ++q) {
      double v = 0.0;
      if (gi < r && q < nref) v = gi > q ? X[t * kPad + q] : (gi == q ? 1.0 : 0.0);
      a.V[gi + (size_t)q * a.ldv] = v;
    }
  }
  // ---- R: the part of the panel that stays inside the band --------------------------------------------------------
  if (blk == 0 && t < 32 && gi < r)
    for (int cc = t; cc < 32; ++cc) a.A[row + (size_t)(a.j0 + cc) * a.ld] = X[t * kPad + cc];
  if (blk == 0)
    for (int i = t; i < 1024; i += kQrRows) a.T[i] = Tsh[(i >> 5) * kPad + (i & 31)];
  if (kCluster) cluster.sync();           // nobody leaves while a peer may still read its partial sums
}

// ---------------------------------------------------------------------------------------------
// panel QR, rows in REGISTERS (used for clusters of up to 4 CTAs; the kernel above keeps them in shared memory and is bound
// by its bandwidth: 3 wavefronts per row and column in the reduction, 5 in the update).  Thread t holds row t of the panel as 32 doubles; the
// 32 column sums of one step are reduced inside the warp by a reduce-scatter butterfly (16 + 16 shuffles: after the exchange
// over lane bits 4..1 lane L holds the sum of column L >> 1, one more exchange over bit 0 completes it), across the 16 warps
// through 4 KB of shared memory, across the CTAs as above.  The update H_j x is 31 - j register FMAs.
// ---------------------------------------------------------------------------------------------
constexpr size_t kQrRegSmem = sizeof(double) * (16 * 32 + 32 + 4 + 32 * kPad + 2 * 64);

template <bool kCluster>
__global__ void __launch_bounds__(kQrRows, 1) k_sbr_qr_reg(QrArgs a) {
  extern __shared__ double sm_qr[];
  double* red = sm_qr;                    // [16][32]
  double* s_sh = red + 16 * 32;           // [32]
  double* par = s_sh + 32;                // beta, tau, scale
  double* Tsh = par + 4;                  // [32][33]
  double* part = Tsh + 32 * kPad;         // [2][64]
  cg::cluster_group cluster = cg::this_cluster();
  const int t = threadIdx.x, blk = blockIdx.x, G = gridDim.x, lane = t & 31, wid = t >> 5;
  const int r = a.m - a.j0 - kBw;
  const int gi = blk * kQrRows + t;
  const size_t row = (size_t)(a.j0 + kBw) + gi;
  double x[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) x[c] = gi < r ? a.A[row + (size_t)(a.j0 + c) * a.ld] : 0.0;
  for (int i = t; i < 32 * kPad; i += kQrRows) Tsh[i] = 0.0;
  __syncthreads();
  const int nref = min(kBw, r - 1);
  unsigned gen = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < nref) {                                            // uniform over the grid
      // ---- s_c = sum over rows i > j of x_i[j] x_i[c] ----------------------------------------------------
      const double xj = gi > j ? x[j] : 0.0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = xj * x[16 * h + i];
#pragma unroll
        for (int o = 16, n = 8; n >= 1; o >>= 1, n >>= 1) {
          const bool up = (lane & o) != 0;
#pragma unroll
          for (int i = 0; i < n; ++i) {
            const double send = up ? v[i] : v[i + n];
            const double keep = up ? v[i + n] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
          }
        }
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
        if ((lane & 1) == 0) red[wid * 32 + 16 * h + (lane >> 1)] = v[0];
      }
      __syncthreads();
      double s = 0.0, piv = 0.0;
      double* mine = kCluster ? part + (j & 1) * 64 : a.slots + (size_t)(j & 1) * (G + 1) * 32;
      if (t < 32) {
        double p = 0.0;
#pragma unroll
        for (int q = 0; q < 16; ++q) p += red[q * 32 + t];
        if (kCluster) mine[t] = p; else mine[blk * 32 + t] = p;
      }
      if (gi == j) {                                           // the pivot row: thread j of CTA 0
        double* pr = kCluster ? mine + 32 : mine + G * 32;
#pragma unroll
        for (int c = 0; c < 32; ++c) pr[c] = x[c];
      }
      if (kCluster) {
        cluster.sync();
        if (t < 32) {
          double pr[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) pr[q] = q < G ? cluster.map_shared_rank(mine, q)[t] : 0.0;
          piv = cluster.map_shared_rank(mine, 0)[32 + t];
#pragma unroll
          for (int q = 0; q < 16; ++q) s += pr[q];
        }
      } else {
        sbr_grid_sync(a.bar, gen);
        if (t < 32) {
          for (int q = 0; q < G; ++q) s += __ldcg(mine + q * 32 + t);
          piv = __ldcg(mine + G * 32 + t);
        }
      }
      if (t < 32) {
        double beta, tau, scale;
        make_house(__shfl_sync(0xffffffffu, piv, j), __shfl_sync(0xffffffffu, s, j), beta, tau, scale);
        s_sh[t] = fma(scale, s, piv);                          // w_c = v'P[:, c]; for c < j it is V[:, c]'v_j
        if (t == 0) { par[0] = beta; par[1] = tau; par[2] = scale; }
      }
      __syncthreads();
      const double beta = par[0], tau = par[1], scale = par[2];
      if (blk == 0 && t >= 32 && t < 64) {                     // column j of T
        const int q = t - 32;
        if (q < j) {
          double sum = 0.0;
          for (int p = q; p < j; ++p) sum = fma(Tsh[q * kPad + p], s_sh[p], sum);
          Tsh[q * kPad + j] = -tau * sum;
        } else if (q == j) {
          Tsh[j * kPad + j] = tau;
        }
      }
      if (gi > j) {
        const double vj = x[j] * scale;
        const double f = tau * vj;
#pragma unroll
        for (int cc = j + 1; cc < 32; ++cc) x[cc] = fma(-f, s_sh[cc], x[cc]);
        x[j] = vj;
      } else if (gi == j) {
#pragma unroll
        for (int cc = j + 1; cc < 32; ++cc) x[cc] = fma(-tau, s_sh[cc], x[cc]);
        x[j] = beta;
      }
    }
  }
  __syncthreads();
  if (gi < a.ldv) {
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      double v = 0.0;
      if (gi < r && q < nref) v = gi > q ? x[q] : (gi == q ? 1.0 : 0.0);
      a.V[gi + (size_t)q * a.ldv] = v;
    }
  }
  if (blk == 0 && t < 32 && gi < r) {
#pragma unroll
    for (int cc = 0; cc < 32; ++cc)
      if (cc >= t) a.A[row + (size_t)(a.j0 + cc) * a.ld] = x[cc];
  }
  if (blk == 0)
    for (int i = t; i < 1024; i += kQrRows) a.T[i] = Tsh[(i >> 5) * kPad + (i & 31)];
  if (kCluster) cluster.sync();
}

// ---------------------------------------------------------------------------------------------
// Z0 = A22 V : CTA = 128 rows x 32 columns over the k range [sp * chunk, (sp + 1) * chunk); thread = 8 x 4 outputs;
// the next 16-column slab of A22 and V travels in registers while the current one is multiplied
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_sbr_av(const double* __restrict__ A22, int ld, int r, const double* __restrict__ V,
                                                int ldv, double* __restrict__ Zp, int chunk) {
  __shared__ __align__(16) double As[16][128];
  __shared__ __align__(16) double Vs[16][32];
  const int t = threadIdx.x, I = blockIdx.x * 128, sp = blockIdx.y;
  const int k0 = sp * chunk, k1 = min(r, k0 + chunk);
  const int tr = t & 15, tc = t >> 4;           // rows tr*2 + 32*a + {0, 1}, columns tc*4 .. +3
  double acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  const int gi = I + t;
  double pa[16], pv[4];
  auto fetch = [&](int kb) {
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const int k = kb + kk;
      pa[kk] = (gi < r && k < k1) ? __ldg(A22 + gi + (size_t)k * ld) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = t + 128 * q, kk = idx & 15, cc = idx >> 4;
      const int k = kb + kk;
      pv[q] = k < k1 ? __ldg(V + k + (size_t)cc * ldv) : 0.0;
    }
  };
  fetch(k0);
  for (int kb = k0; kb < k1; kb += 16) {
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) As[kk][t] = pa[kk];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = t + 128 * q;
      Vs[idx & 15][idx >> 4] = pv[q];
    }
    __syncthreads();
    if (kb + 16 < k1) fetch(kb + 16);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double av[8], vv[4];
#pragma unroll
      for (int a4 = 0; a4 < 4; ++a4) {
        const double2 p = *reinterpret_cast<const double2*>(&As[kk][tr * 2 + 32 * a4]);
        av[2 * a4] = p.x; av[2 * a4 + 1] = p.y;
      }
      const double2 v01 = *reinterpret_cast<const double2*>(&Vs[kk][tc * 4]);
      const double2 v23 = *reinterpret_cast<const double2*>(&Vs[kk][tc * 4 + 2]);
      vv[0] = v01.x; vv[1] = v01.y; vv[2] = v23.x; vv[3] = v23.y;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], vv[j], acc[i][j]);
    }
    __syncthreads();
  }
  double* out = Zp + (size_t)sp * ldv * 32;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int a4 = 0; a4 < 4; ++a4) {
      const int rr = I + tr * 2 + 32 * a4;
      *reinterpret_cast<double2*>(out + rr + (size_t)(tc * 4 + j) * ldv) = make_double2(acc[2 * a4][j], acc[2 * a4 + 1][j]);
    }
}

// ---------------------------------------------------------------------------------------------
// Z0 = sum of the k-split partials; per-CTA partial of G0 = V'Z0 (32 x 32) and gz = V'z (32 x L).  CTA = 128 rows in two
// halves of 64; thread = (row of G, two columns).
// ---------------------------------------------------------------------------------------------
constexpr size_t kVtzSmem = sizeof(double) * 3 * 64 * kPad;

__global__ void __launch_bounds__(512) k_sbr_vtz(const double* __restrict__ V, int ldv, int r, const double* __restrict__ Zp,
                                                 int nsplit, double* __restrict__ Z0, const double* __restrict__ z, int ldz,
                                                 int zrow0, int L, double* __restrict__ Gp) {
  extern __shared__ double sm_vtz[];
  double (*Vs)[kPad] = reinterpret_cast<double (*)[kPad]>(sm_vtz);
  double (*Zs)[kPad] = reinterpret_cast<double (*)[kPad]>(sm_vtz + 64 * kPad);
  double (*zs)[kPad] = reinterpret_cast<double (*)[kPad]>(sm_vtz + 2 * 64 * kPad);
  const int t = threadIdx.x, I = blockIdx.x * 128;
  const int a = t & 31, bq = t >> 5;            // outputs G[a][bq*2 + x]
  double g[2] = {0.0, 0.0}, gz[2] = {0.0, 0.0};
  for (int h = 0; h < 2; ++h) {
    const int R0 = I + h * 64;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = t + 512 * q, rw = idx & 63, col = idx >> 6;
      const int gi = R0 + rw;
      const size_t o = gi + (size_t)col * ldv;
      double s = 0.0;
      for (int sp = 0; sp < nsplit; ++sp) s += Zp[(size_t)sp * ldv * 32 + o];
      Z0[o] = s;
      Vs[rw][col] = V[o];
      Zs[rw][col] = s;
      zs[rw][col] = (col < L && gi < r) ? z[(size_t)zrow0 + gi + (size_t)col * ldz] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int rw = 0; rw < 64; ++rw) {
      const double va = Vs[rw][a];
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        g[x] = fma(va, Zs[rw][bq * 2 + x], g[x]);
        gz[x] = fma(va, zs[rw][bq * 2 + x], gz[x]);
      }
    }
    __syncthreads();
  }
  double* out = Gp + (size_t)blockIdx.x * 2048;
#pragma unroll
  for (int x = 0; x < 2; ++x) {
    out[a * 32 + bq * 2 + x] = g[x];
    out[1024 + a * 32 + bq * 2 + x] = gz[x];
  }
}

// ---------------------------------------------------------------------------------------------
// One CTA: G0, gz = fixed-order sums of the partials;  ST[0:1024] = S = T' G0 T,  ST[1024:2048] = T' gz  (row-major 32 x 32)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_sbr_st(const double* __restrict__ Gp, int nblk, const double* __restrict__ T,
                                                 double* __restrict__ ST) {
  __shared__ double Ts[32][kPad], Gs[32][kPad], Xs[32][kPad], gzs[32][kPad];
  const int t = threadIdx.x, aa = t >> 5, cc = t & 31;
  double g0 = 0.0, g1 = 0.0, z0 = 0.0, z1 = 0.0;
  int bl = 0;
  for (; bl + 1 < nblk; bl += 2) {
    g0 += Gp[(size_t)bl * 2048 + t];
    z0 += Gp[(size_t)bl * 2048 + 1024 + t];
    g1 += Gp[(size_t)(bl + 1) * 2048 + t];
    z1 += Gp[(size_t)(bl + 1) * 2048 + 1024 + t];
  }
  if (bl < nblk) { g0 += Gp[(size_t)bl * 2048 + t]; z0 += Gp[(size_t)bl * 2048 + 1024 + t]; }
  Ts[aa][cc] = T[t];
  Gs[aa][cc] = g0 + g1;
  gzs[aa][cc] = z0 + z1;
  __syncthreads();
  double s = 0.0;
#pragma unroll 8
  for (int p = 0; p < 32; ++p) s = fma(Gs[aa][p], Ts[p][cc], s);      // X = G0 T
  Xs[aa][cc] = s;
  __syncthreads();
  double u = 0.0;
  s = 0.0;
#pragma unroll 8
  for (int p = 0; p < 32; ++p) {
    s = fma(Ts[p][aa], Xs[p][cc], s);                                  // S = T' X
    u = fma(Ts[p][aa], gzs[p][cc], u);                                 // T' gz
  }
  ST[t] = s;
  ST[1024 + t] = u;
}

// ---------------------------------------------------------------------------------------------
// W = Z0 T - 1/2 V S;  z <- z - V (T' gz).  Thread = one row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_sbr_w(const double* __restrict__ V, int ldv, int r, const double* __restrict__ Z0,
                                               const double* __restrict__ T, const double* __restrict__ ST,
                                               double* __restrict__ W, double* __restrict__ z, int ldz, int zrow0, int L) {
  __shared__ double Ts[32][kPad], Ss[32][kPad], tzs[32][kPad];
  const int t = threadIdx.x;
  const int i = blockIdx.x * 128 + t;           // < ldv
  double zr[32], vr[32];
#pragma unroll
  for (int q = 0; q < 32; ++q) {
    zr[q] = Z0[i + (size_t)q * ldv];
    vr[q] = V[i + (size_t)q * ldv];
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int idx = t + 128 * q, aa = idx >> 5, cc = idx & 31;
    Ts[aa][cc] = T[idx];
    Ss[aa][cc] = ST[idx];
    tzs[aa][cc] = ST[1024 + idx];
  }
  __syncthreads();
  for (int cc = 0; cc < 32; ++cc) {
    double w1 = 0.0, w2 = 0.0, w3 = 0.0, w4 = 0.0;
#pragma unroll
    for (int q = 0; q < 32; q += 2) {
      w1 = fma(zr[q], Ts[q][cc], w1);
      w2 = fma(vr[q], Ss[q][cc], w2);
      w3 = fma(zr[q + 1], Ts[q + 1][cc], w3);
      w4 = fma(vr[q + 1], Ss[q + 1][cc], w4);
    }
    W[i + (size_t)cc * ldv] = (w1 + w3) - 0.5 * (w2 + w4);
  }
  if (i < r)
    for (int l = 0; l < L; ++l) {
      double u = 0.0;
#pragma unroll
      for (int q = 0; q < 32; ++q) u = fma(vr[q], tzs[q][l], u);
      z[(size_t)zrow0 + i + (size_t)l * ldz] -= u;
    }
}

// ---------------------------------------------------------------------------------------------
// C -= V W' + W V' on the 128 x 128 tiles of the lower triangle; thread = 8 x 8 outputs, the operands of the whole tile in
// shared memory.  Rows >= 32 of the first 32 columns - the next panel - are left to k_sbr_pu, so that the QR of the next
// panel runs beside this kernel (look-ahead).  An off-diagonal tile is written twice: in place and - transposed through shared memory, so that both
// stores are coalesced - into the upper triangle, which k_sbr_av reads as a plain square.
// ---------------------------------------------------------------------------------------------
constexpr int kR2kLds = 130;       // row stride of the transposed staging tile (16-byte aligned rows)
constexpr size_t kR2kSmem = sizeof(double) * 128 * kR2kLds;     // >= 4 x 32 x 128 operands

__global__ void __launch_bounds__(256, 1) k_sbr_r2k(double* __restrict__ C, int ld, int r, const double* __restrict__ V,
                                                    const double* __restrict__ W, int ldv) {
  extern __shared__ __align__(16) double sm_r2k[];
  double (*Vi)[128] = reinterpret_cast<double (*)[128]>(sm_r2k);
  double (*Wi)[128] = reinterpret_cast<double (*)[128]>(sm_r2k + 32 * 128);
  double (*Vj)[128] = reinterpret_cast<double (*)[128]>(sm_r2k + 2 * 32 * 128);
  double (*Wj)[128] = reinterpret_cast<double (*)[128]>(sm_r2k + 3 * 32 * 128);
  // lower-triangle tile index -> (bi >= bj)
  int bi = (int)((sqrt(8.0 * blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((bi + 1) * (bi + 2) / 2 <= (int)blockIdx.x) ++bi;
  while (bi * (bi + 1) / 2 > (int)blockIdx.x) --bi;
  const int bj = blockIdx.x - bi * (bi + 1) / 2;
  const int t = threadIdx.x, I = bi * 128, J = bj * 128;
  const int tx = t & 15, ty = t >> 4;           // rows tx*2 + 32*a + {0, 1}, columns ty*8 .. +7
  double acc[8][8];
  // the tile of C rides in the accumulators: its loads overlap the operand staging
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = J + ty * 8 + j;
#pragma unroll
    for (int a4 = 0; a4 < 4; ++a4) {
      const int rr = I + tx * 2 + 32 * a4;
      double2 cv = make_double2(0.0, 0.0);
      if (col < r && !(col < 32 && rr >= 32)) {          // rows >= 32 of the first 32 columns = the next panel (k_sbr_pu)
        if (rr + 1 < r) cv = *reinterpret_cast<const double2*>(C + rr + (size_t)col * ld);
        else if (rr < r) cv.x = C[rr + (size_t)col * ld];
      }
      acc[2 * a4][j] = cv.x; acc[2 * a4 + 1][j] = cv.y;
    }
  }
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int idx = t + 256 * q, kk = idx >> 7, rw = idx & 127;
    Vi[kk][rw] = V[I + rw + (size_t)kk * ldv];
    Wi[kk][rw] = W[I + rw + (size_t)kk * ldv];
    Vj[kk][rw] = V[J + rw + (size_t)kk * ldv];
    Wj[kk][rw] = W[J + rw + (size_t)kk * ldv];
  }
  __syncthreads();
#pragma unroll 2
  for (int kk = 0; kk < 32; ++kk) {
    double vi[8], wi[8], vj[8], wj[8];
#pragma unroll
    for (int a4 = 0; a4 < 4; ++a4) {
      const double2 p = *reinterpret_cast<const double2*>(&Vi[kk][tx * 2 + 32 * a4]);
      const double2 q2 = *reinterpret_cast<const double2*>(&Wi[kk][tx * 2 + 32 * a4]);
      vi[2 * a4] = -p.x; vi[2 * a4 + 1] = -p.y;
      wi[2 * a4] = -q2.x; wi[2 * a4 + 1] = -q2.y;
      const double2 p3 = *reinterpret_cast<const double2*>(&Vj[kk][ty * 8 + 2 * a4]);
      const double2 p4 = *reinterpret_cast<const double2*>(&Wj[kk][ty * 8 + 2 * a4]);
      vj[2 * a4] = p3.x; vj[2 * a4 + 1] = p3.y;
      wj[2 * a4] = p4.x; wj[2 * a4 + 1] = p4.y;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[i][j] = fma(vi[i], wj[j], acc[i][j]);
        acc[i][j] = fma(wi[i], vj[j], acc[i][j]);
      }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = J + ty * 8 + j;
    if (col >= r) continue;
#pragma unroll
    for (int a4 = 0; a4 < 4; ++a4) {
      const int rr = I + tx * 2 + 32 * a4;
      if (col < 32 && rr >= 32) continue;
      if (rr + 1 < r) *reinterpret_cast<double2*>(C + rr + (size_t)col * ld) = make_double2(acc[2 * a4][j], acc[2 * a4 + 1][j]);
      else if (rr < r) C[rr + (size_t)col * ld] = acc[2 * a4][j];
    }
  }
  if (bi == bj) return;
  // ---- mirror: S[row][col] = new C[I + row, J + col]  ->  C[J + col, I + row] -------------------------------------------
  __syncthreads();                               // the operands are dead
  double* S = sm_r2k;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rw = tx * 2 + 32 * (i >> 1) + (i & 1);
#pragma unroll
    for (int j = 0; j < 8; ++j) S[rw * kR2kLds + ty * 8 + j] = acc[i][j];
  }
  __syncthreads();
  const int c2 = (t & 63) * 2;                   // J + c2 + 1 <= J + 127 < I < r: always inside
  if (J == 0 && c2 < 32) return;                 // mirror image of the next panel: never read again
  for (int rw = t >> 6; rw < 128; rw += 4) {
    if (I + rw >= r) break;
    *reinterpret_cast<double2*>(C + (size_t)(J + c2) + (size_t)(I + rw) * ld) =
        *reinterpret_cast<const double2*>(S + rw * kR2kLds + c2);
  }
}

// ---------------------------------------------------------------------------------------------
// look-ahead: the same update restricted to the next panel, C[32:, 0:32] -= V W[0:32]' + W V[0:32]'.  Thread = one row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_sbr_pu(double* __restrict__ C, int ld, int r, const double* __restrict__ V,
                                                const double* __restrict__ W, int ldv) {
  __shared__ double Vt[32][kPad], Wt[32][kPad];          // [k][c] = V[c, k], W[c, k] of the first 32 rows
  const int t = threadIdx.x;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int idx = t + 128 * q, kk = idx >> 5, cc = idx & 31;
    Vt[kk][cc] = V[cc + (size_t)kk * ldv];
    Wt[kk][cc] = W[cc + (size_t)kk * ldv];
  }
  const int i = 32 + blockIdx.x * 128 + t;
  double vr[32], wr[32];
#pragma unroll
  for (int q = 0; q < 32; ++q) {
    vr[q] = i < r ? V[i + (size_t)q * ldv] : 0.0;
    wr[q] = i < r ? W[i + (size_t)q * ldv] : 0.0;
  }
  __syncthreads();
  if (i >= r) return;
  for (int cc = 0; cc < 32; ++cc) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      s0 = fma(vr[q], Wt[q][cc], s0);
      s1 = fma(wr[q], Vt[q][cc], s1);
    }
    C[i + (size_t)cc * ld] -= s0 + s1;
  }
}

// ---------------------------------------------------------------------------------------------
// band storage
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sbr_band(const double* __restrict__ A, int ld, int m, double* __restrict__ Bd, int ncolb) {
  const int off = threadIdx.x & 63;
  const int j = blockIdx.x * 4 + (threadIdx.x >> 6);
  if (j >= ncolb) return;
  double v = 0.0;
  if (off <= kBw && j + off < m) v = A[(size_t)(j + off) + (size_t)j * ld];
  Bd[off + (size_t)j * kLdb] = v;
}

__global__ void k_sbr_diag(const double* __restrict__ Bd, int m, double* __restrict__ d, double* __restrict__ e) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) d[i] = Bd[(size_t)i * kLdb];
  if (i < m - 1) e[i] = Bd[1 + (size_t)i * kLdb];
}

// ---------------------------------------------------------------------------------------------
// bulge chasing
// ---------------------------------------------------------------------------------------------
struct ChaseArgs {
  double* Bd; int m;
  double* z; int L;                // m x L, ld = m
  int* prog;                       // [m] steps finished per sweep;  prog[m] = ticket counter, prog[m + 1] = error flag
  unsigned sleep_ns;               // back-off of the spinning lanes (watcher, publisher, thread 0): the chase shares its SMs with the
                                   // per-cell ensemble kernels, and three hot spin loops per CTA cost those kernels issue slots
};
__device__ __forceinline__ void chase_backoff(unsigned ns) {
  if (ns) __nanosleep(ns);
}

// One CTA (three warps) per sweep.  Per step: thread 0 waits for the predecessor sweep; warp 0 owns the off-diagonal block
// (row in registers for the right-application and the new reflector, shared memory for the column pass), warp 1 the
// diagonal block, warp 2 the right-hand sides - their loads are issued together, right after the wait, and the diagonal /
// right-hand-side updates run beside the left-application of warp 0.  A wait that does not end within ~2^24 polls
// (seconds; a step takes microseconds) raises the error flag instead of hanging the device.
//
// kDec ("sbr_chase_impl" = 1, experimental - DESIGN.md section 9 option (a)): two more warps take the global hand-shakes off
// the step time.  A WATCHER warp polls the predecessor's progress word (ld.relaxed.gpu) into shared memory, so the compute
// warps check a shared-memory word; a PUBLISHER warp turns "step k stored" (shared memory) into fence + st.relaxed.gpu, so the
// compute warps go on with step k + 1 at once.  The compute warps then synchronise on named barrier 1 (96 threads).
constexpr int kChaseThreads = 96;
constexpr int kChaseThreadsDec = 160;

template <bool kDec>
__device__ __forceinline__ void chase_csync() {
  if (kDec) asm volatile("bar.sync 1, 96;" ::: "memory");
  else __syncthreads();
}

template <bool kDec>
__global__ void __launch_bounds__(kDec ? kChaseThreadsDec : kChaseThreads) k_sbr_chase_t(ChaseArgs a) {
  __shared__ double Bs[32][kPad], Ds[32][kPad], zs[32][kPad];
  __shared__ double vs[32], ws[32], vps[32], ts[32];
  __shared__ double sh_tau;
  __shared__ int sh_s;
  __shared__ int sh_seen, sh_done;   // kDec: predecessor progress seen by the watcher; steps of this sweep stored
  const int tid = threadIdx.x, l = tid & 31, wid = tid >> 5, m = a.m;
  constexpr int b = kBw;
  for (;;) {
    __syncthreads();
    if (tid == 0) {
      sh_s = atomicAdd(a.prog + m, 1);
      if (kDec) { sh_seen = 0; sh_done = 0; }
    }
    __syncthreads();
    const int s = sh_s;
    if (s >= m - 2) break;
    const int totp = s > 0 ? (m - s + b - 1) / b : 0;     // steps of sweep s - 1
    const int tot = (m - s - 1 + b - 1) / b;              // steps of this sweep: row blocks below the diagonal of column s
    if (kDec && wid == 3) {                               // watcher
      if (l == 0 && s > 0) {
        const int* p = a.prog + (s - 1);
        int v = 0, last = 0;
        long long spins = 0;
        // RELAXED poll: ld.acquire.gpu is LDG.STRONG.GPU + CCTL.IVALL in SASS - an invalidation of the SM's whole L1 per poll, i.e.
        // continuously on 80 SMs for the 45 ms of the chase, which the per-cell ensemble kernels sharing those SMs pay for
        // (+20 ms on whichever of them runs beside the chase, profiles/r2h_*).  No acquire is needed: every load of band data
        // below is ld.global.cg (LDG.STRONG.GPU, served by L2), issued after the flag value has travelled through shared memory.
        while (v < totp && ++spins < (1ll << 28)) {
          asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
          if (v > last) { *(volatile int*)&sh_seen = v; last = v; }
          else chase_backoff(a.sleep_ns);
        }
        if (v < totp) atomicExch(a.prog + m + 1, 1);
      }
      continue;
    }
    if (kDec && wid == 4) {                               // publisher
      if (l == 0) {
        int pub = 0;
        long long spins = 0;
        while (pub < tot && ++spins < (1ll << 32)) {
          const int dn = *(volatile int*)&sh_done;
          if (dn <= pub) { chase_backoff(a.sleep_ns); continue; }
          {
            // ONE gpu-scope fence (cumulative: the compute warps' stores, seen through sh_done) + a relaxed store; st.release
            // on top of __threadfence() was a second MEMBAR + CCTL.IVALL per step
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(a.prog + s), "r"(dn) : "memory");
            pub = dn;
          }
        }
        if (pub < tot) atomicExch(a.prog + m + 1, 1);
      }
      continue;
    }
    double taup = 0.0;
    for (int k = 0; k < tot; ++k) {
      if (s > 0 && tid == 0) {
        const int need = min(k + 2, totp);
        if (kDec) {
          int spins = 0;
          while (*(volatile int*)&sh_seen < need && ++spins < (1 << 28)) chase_backoff(a.sleep_ns);
          if (*(volatile int*)&sh_seen < need) atomicExch(a.prog + m + 1, 1);
          __threadfence_block();                          // the block loads below stay behind the flag
        } else {
          const int* p = a.prog + (s - 1);
          int v, spins = 0;
          for (;;) {                                      // relaxed: see the watcher warp above
            asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
            if (v >= need || ++spins >= (1 << 24)) break;
            chase_backoff(a.sleep_ns);
          }
          if (v < need) atomicExch(a.prog + m + 1, 1);
        }
      }
      chase_csync<kDec>();
      // block geometry (uniform)
      int st = 0, lp = 0, r0, ln;
      if (k == 0) {
        ln = min(b, m - 1 - s);
        r0 = s + 1;
      } else {
        st = s + 1 + (k - 1) * b;
        lp = min(b, m - st);
        r0 = st + lp;
        ln = min(b, m - r0);
      }
      if (wid == 0) {
        double beta, tau, scale, x;
        if (k == 0) {
          // ---- type 1: reflector from column s, rows s+1 .. s+ln ---------------------------------------------
          x = l < ln ? __ldcg(a.Bd + (1 + l) + (size_t)s * kLdb) : 0.0;
          const double xn2 = wsum((l >= 1 && l < ln) ? x * x : 0.0);
          make_house(__shfl_sync(0xffffffffu, x, 0), xn2, beta, tau, scale);
          vs[l] = l == 0 ? 1.0 : ((l < ln && tau != 0.0) ? x * scale : 0.0);
          if (l == 0) sh_tau = tau;
          if (l < ln) a.Bd[(1 + l) + (size_t)s * kLdb] = l == 0 ? beta : 0.0;
        } else {
          // ---- type 2: block below the previous diagonal block: right-apply H_prev, new reflector ---------------
          double br[32];
#pragma unroll
          for (int q = 0; q < 32; ++q)
            br[q] = (l < ln && q < lp) ? __ldcg(a.Bd + (lp + l - q) + (size_t)(st + q) * kLdb) : 0.0;
          double u0 = 0.0, u1 = 0.0, u2 = 0.0, u3 = 0.0;
#pragma unroll
          for (int q = 0; q < 32; q += 4) {
            u0 = fma(br[q], vps[q], u0);
            u1 = fma(br[q + 1], vps[q + 1], u1);
            u2 = fma(br[q + 2], vps[q + 2], u2);
            u3 = fma(br[q + 3], vps[q + 3], u3);
          }
          const double u = taup * ((u0 + u1) + (u2 + u3));
#pragma unroll
          for (int q = 0; q < 32; ++q) br[q] = fma(-u, vps[q], br[q]);
          x = br[0];
          const double xn2 = wsum((l >= 1 && l < ln) ? x * x : 0.0);
          make_house(__shfl_sync(0xffffffffu, x, 0), xn2, beta, tau, scale);
          vs[l] = l == 0 ? 1.0 : ((l < ln && tau != 0.0) ? x * scale : 0.0);
          if (l == 0) sh_tau = tau;
          br[0] = l == 0 ? beta : 0.0;
#pragma unroll
          for (int q = 0; q < 32; ++q) Bs[l][q] = br[q];
        }
      } else if (wid == 1) {
        // ---- diagonal block r0 .. r0+ln-1, both triangles in shared memory ----------------------------------------
        double dr[32];
#pragma unroll
        for (int q = 0; q < 32; ++q)
          dr[q] = (q <= l && l < ln) ? __ldcg(a.Bd + (l - q) + (size_t)(r0 + q) * kLdb) : 0.0;
#pragma unroll
        for (int q = 0; q < 32; ++q) Ds[l][q] = dr[q];
        __syncwarp();
#pragma unroll 8
        for (int q = 0; q < 32; ++q)
          if (q > l) Ds[l][q] = Ds[q][l];
      } else {
        for (int cc = 0; cc < a.L; ++cc) zs[l][cc] = l < ln ? __ldcg(a.z + (size_t)(r0 + l) + (size_t)cc * m) : 0.0;
      }
      chase_csync<kDec>();
      const double tau = sh_tau;
      const double vl = vs[l];
      if (wid == 0) {
        if (k > 0) {
          // ---- left-apply the new reflector to columns 1 .. lp-1 (lane = column), then store the block by rows ----
          if (l >= 1 && l < lp) {
            double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              y0 = fma(vs[i], Bs[i][l], y0);
              y1 = fma(vs[i + 1], Bs[i + 1][l], y1);
              y2 = fma(vs[i + 2], Bs[i + 2][l], y2);
              y3 = fma(vs[i + 3], Bs[i + 3][l], y3);
            }
            const double y = tau * ((y0 + y1) + (y2 + y3));
#pragma unroll
            for (int i = 0; i < 32; ++i) Bs[i][l] = fma(-vs[i], y, Bs[i][l]);
          }
          __syncwarp();
          if (l < ln)
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (q < lp) a.Bd[(lp + l - q) + (size_t)(st + q) * kLdb] = Bs[l][q];
        }
        vps[l] = vl;                      // the reflector the next step applies from the right
      } else if (wid == 1) {
        // ---- type 3: D <- H D H ---------------------------------------------------------------------------------
        double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          p0 = fma(Ds[l][q], vs[q], p0);
          p1 = fma(Ds[l][q + 1], vs[q + 1], p1);
          p2 = fma(Ds[l][q + 2], vs[q + 2], p2);
          p3 = fma(Ds[l][q + 3], vs[q + 3], p3);
        }
        const double p = tau * ((p0 + p1) + (p2 + p3));
        const double w = fma(-0.5 * tau * wsum(p * vl), vl, p);
        ws[l] = w;
        __syncwarp();
        if (l < ln)
#pragma unroll
          for (int q = 0; q < 32; ++q)
            if (q <= l) a.Bd[(l - q) + (size_t)(r0 + q) * kLdb] = Ds[l][q] - vl * ws[q] - w * vs[q];
      } else if (a.L > 0) {
        // ---- right-hand sides: z[r0 .. r0+ln-1, :] <- H z -----------------------------------------------------------
        if (l < a.L) {
          double t0 = 0.0, t1 = 0.0;
#pragma unroll 8
          for (int i = 0; i < 32; i += 2) {
            t0 = fma(vs[i], zs[i][l], t0);
            t1 = fma(vs[i + 1], zs[i + 1][l], t1);
          }
          ts[l] = tau * (t0 + t1);
        }
        __syncwarp();
        if (l < ln)
          for (int cc = 0; cc < a.L; ++cc) a.z[(size_t)(r0 + l) + (size_t)cc * m] = fma(-vl, ts[cc], zs[l][cc]);
      }
      taup = tau;
      // publish the step: the CTA barrier orders every thread's stores before thread 0's release store (cumulativity), the
      // pattern of a cooperative-groups grid barrier - one fence on the critical path instead of two
      chase_csync<kDec>();
      if (tid == 0) {
        if (kDec) { __threadfence_block(); *(volatile int*)&sh_done = k + 1; }
        else asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.prog + s), "r"(k + 1) : "memory");
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Bulge chase without flags and without fences ("sbr_chase_impl" = 3): the LL protocol of collective libraries applied to the
// band.  Every stored double travels as two 8-byte words, each carrying 32 bits of the value and a 32-bit TAG = 1 + the number of
// the sweep that wrote it (0 = the band as stage 1 left it); 8-byte accesses are single-copy atomic, so a reader that finds the
// expected tag in both words holds the value that sweep wrote - no flag, no fence, no release / acquire pair, and a step is
// handed to the successor element by element, as soon as each store lands in L2, instead of "two steps later, after a fence".
// Why one expected tag per sweep suffices: sweep s - 1 partitions rows s .. m-1 into blocks R''_k of 32 and rewrites every band
// element whose row and column blocks differ by at most one (diagonal blocks, sub-diagonal blocks incl. their bulge), each
// exactly once; sweep s's blocks are those shifted by one row and column, so everything it reads at step k was written by
// sweep s - 1 at step k or k + 1 and carries tag s - EXCEPT the last row of its sub-diagonal block left of that block's last
// column (row block k + 1 against column block k - 1 of the predecessor): structurally zero (outside band and bulge), never
// written, and not read here.  The arithmetic is the one of k_sbr_chase_t, bit for bit.
// A wait that does not end (or the error flag raised by another sweep) aborts the sweep instead of hanging the device.
// ---------------------------------------------------------------------------------------------
struct ChaseLLArgs {
  ulonglong2* Bt; int m;           // tagged band, element (off, col) at off + col * kLdb
  ulonglong2* zt; int L;           // tagged right-hand sides, m x L
  int* ctl;                        // ctl[0] = ticket counter, ctl[1] = error flag
};
__device__ __forceinline__ void ll_store(ulonglong2* p, double v, unsigned tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v), t = (unsigned long long)tag << 32;
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"((b & 0xffffffffull) | t), "l"((b >> 32) | t) : "memory");
}
// load and check are separate so that a lane can have a batch of loads in flight before it looks at the first tag (a check
// right behind every load would serialise 32 L2 round trips per step); volatile, no memory clobber: the loads keep their order
// among themselves and are re-issued in every pass of a spin loop, but do not fence the surrounding code
__device__ __forceinline__ ulonglong2 ll_ld(const ulonglong2* p) {
  ulonglong2 w;
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w.x), "=l"(w.y) : "l"(p));
  return w;
}
__device__ __forceinline__ bool ll_ok(const ulonglong2 w, unsigned tag, double& v) {
  if ((unsigned)(w.x >> 32) != tag || (unsigned)(w.y >> 32) != tag) return false;
  v = __longlong_as_double((long long)((w.x & 0xffffffffull) | (w.y << 32)));
  return true;
}
__device__ __forceinline__ bool ll_try(const ulonglong2* p, unsigned tag, double& v) { return ll_ok(ll_ld(p), tag, v); }
constexpr int kLLBatch = 16;
// true = give up (own time-out or somebody else's error)
__device__ __forceinline__ bool ll_giveup(int* ctl, unsigned& spins) {
  if ((++spins & 1023u) != 0) return false;
  int e;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(e) : "l"(ctl + 1) : "memory");
  if (e) return true;
  if (spins >= (1u << 23)) { atomicExch(ctl + 1, 1); return true; }
  return false;
}

// CTA = 3 compute warps (the roles of k_sbr_chase_t) + 4 loader warps.  The loaders fetch the blocks of step k + 1 - every
// thread ONE batch of <= 16 tagged elements, so a block costs one L2 round trip however often it has to be polled - into the
// other half of a double-buffered shared-memory stage while the compute warps work on step k: blocks of different steps of a
// sweep are disjoint, and what step k + 1 reads comes from the predecessor sweep, not from step k.  Producer / consumer
// hand-over through named barriers (full[2] = 1, 2; empty[2] = 3, 4; compute-only = 5).
constexpr int kLLCompute = 96, kLLLoaders = 128, kLLThreads = kLLCompute + kLLLoaders;
__device__ __forceinline__ void nb_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nb_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// kMinCtas = 2 caps the registers at 146 per thread (some spills): half the footprint beside a co-resident per-cell kernel
template <int kMinCtas>
__global__ void __launch_bounds__(kLLThreads, kMinCtas) k_sbr_chase_ll(ChaseLLArgs a) {
  __shared__ double Bs[2][32][kPad], Ds[2][32][kPad], zs[32][kPad];
  __shared__ __align__(16) double vs[32], ws[32], vps[32], ts[32];   // 16-byte aligned: broadcast reads of neighbours as one LDS.128
  __shared__ double sh_tau;
  __shared__ int sh_s;
  const int tid = threadIdx.x, l = tid & 31, wid = tid >> 5, m = a.m;
  constexpr int b = kBw;
  for (;;) {
    __syncthreads();
    if (tid == 0) sh_s = atomicAdd(a.ctl, 1);
    __syncthreads();
    const int s = sh_s;
    if (s >= m - 2) break;
    const unsigned rtag = (unsigned)s, wtag = (unsigned)s + 1u;   // written by sweep s - 1 (or initial) / by this sweep
    const int tot = (m - s - 1 + b - 1) / b;
    unsigned spins = 0;
    if (wid >= 3) {
      // =================================== loaders: warps 3, 4 sub-diagonal block, 5, 6 diagonal block ===================
      const int lw = wid - 3, half = lw & 1;                        // half: columns q in [16 half, 16 half + 16)
      for (int k = 0; k < tot; ++k) {
        int st = 0, lp = 0, r0, ln;
        if (k == 0) { ln = min(b, m - 1 - s); r0 = s + 1; }
        else { st = s + 1 + (k - 1) * b; lp = min(b, m - st); r0 = st + lp; ln = min(b, m - r0); }
        const int buf = k & 1;
        if (k >= 2) nb_sync(3 + buf, kLLThreads);                   // the compute warps are done with step k - 2
        double v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.0;
        unsigned pend = 0;
        const ulonglong2* base;
        if (lw < 2) {
          if (k == 0) {                                             // column s, rows s+1 .. s+ln -> Bs[.][l][0]
            base = a.Bt + (1 + l) + (size_t)s * kLdb;
            pend = (half == 0 && l < ln) ? 1u : 0u;
          } else {
            // (l == 31, q <= 30): row block k + 1 against column block k - 1 of the predecessor - structurally zero, never written
            const unsigned rowmask = l < ln ? ((l == 31 ? 0x80000000u : 0xffffffffu) & (lp >= 32 ? 0xffffffffu : ((1u << lp) - 1u))) : 0u;
            pend = (rowmask >> (16 * half)) & 0xffffu;
            base = a.Bt + (lp + l) + (size_t)st * kLdb + (size_t)(16 * half) * (kLdb - 1);   // element q at + q (kLdb - 1)
          }
        } else {
          const unsigned rowmask = l < ln ? (l >= 31 ? 0xffffffffu : ((2u << l) - 1u)) : 0u;  // q <= l
          pend = (rowmask >> (16 * half)) & 0xffffu;
          base = a.Bt + l + (size_t)r0 * kLdb + (size_t)(16 * half) * (kLdb - 1);
        }
        while (pend) {
          ulonglong2 raw[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if ((pend >> j) & 1u) raw[j] = ll_ld(base + (size_t)j * (kLdb - 1));
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (((pend >> j) & 1u) && ll_ok(raw[j], rtag, v[j])) pend &= ~(1u << j);
          if (pend && ll_giveup(a.ctl, spins)) break;
        }
        double (*dst)[kPad] = lw < 2 ? Bs[buf] : Ds[buf];
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[l][16 * half + j] = v[j];
        nb_arrive(1 + buf, kLLThreads);                             // stage `buf` holds step k
      }
      continue;
    }
    // ======================================= compute warps ============================================================
    double taup = 0.0;
    for (int k = 0; k < tot; ++k) {
      int st = 0, lp = 0, r0, ln;
      if (k == 0) { ln = min(b, m - 1 - s); r0 = s + 1; }
      else { st = s + 1 + (k - 1) * b; lp = min(b, m - st); r0 = st + lp; ln = min(b, m - r0); }
      const int buf = k & 1;
      double (*B)[kPad] = Bs[buf];
      double (*D)[kPad] = Ds[buf];
      if (wid == 2) {                                               // right-hand sides: few elements, fetched here
        for (int c0 = 0; c0 < a.L; c0 += 16) {
          const int nc = min(16, a.L - c0);
          double v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.0;
          unsigned pend = l < ln ? ((1u << nc) - 1u) : 0u;
          const ulonglong2* base = a.zt + (size_t)(r0 + l) + (size_t)c0 * m;
          while (pend) {
            ulonglong2 raw[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if ((pend >> j) & 1u) raw[j] = ll_ld(base + (size_t)j * m);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (((pend >> j) & 1u) && ll_ok(raw[j], rtag, v[j])) pend &= ~(1u << j);
            if (pend && ll_giveup(a.ctl, spins)) break;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < nc) zs[l][c0 + j] = v[j];
        }
      }
      nb_sync(1 + buf, kLLThreads);                                 // the loaders have filled stage `buf`
      if (wid == 0) {
        double beta, tau, scale, x;
        if (k == 0) {
          // ---- type 1: reflector from column s, rows s+1 .. s+ln ---------------------------------------------
          x = l < ln ? B[l][0] : 0.0;
          const double xn2 = wsum((l >= 1 && l < ln) ? x * x : 0.0);
          make_house(__shfl_sync(0xffffffffu, x, 0), xn2, beta, tau, scale);
          vs[l] = l == 0 ? 1.0 : ((l < ln && tau != 0.0) ? x * scale : 0.0);
          if (l == 0) sh_tau = tau;
          if (l < ln) ll_store(a.Bt + (1 + l) + (size_t)s * kLdb, l == 0 ? beta : 0.0, wtag);
        } else {
          // ---- type 2: block below the previous diagonal block: right-apply H_prev, new reflector ---------------
          double br[32];
#pragma unroll
          for (int q = 0; q < 32; ++q) br[q] = B[l][q];
          double u0 = 0.0, u1 = 0.0, u2 = 0.0, u3 = 0.0;
#pragma unroll
          for (int q = 0; q < 32; q += 4) {
            u0 = fma(br[q], vps[q], u0);
            u1 = fma(br[q + 1], vps[q + 1], u1);
            u2 = fma(br[q + 2], vps[q + 2], u2);
            u3 = fma(br[q + 3], vps[q + 3], u3);
          }
          const double u = taup * ((u0 + u1) + (u2 + u3));
#pragma unroll
          for (int q = 0; q < 32; ++q) br[q] = fma(-u, vps[q], br[q]);
          x = br[0];
          const double xn2 = wsum((l >= 1 && l < ln) ? x * x : 0.0);
          make_house(__shfl_sync(0xffffffffu, x, 0), xn2, beta, tau, scale);
          vs[l] = l == 0 ? 1.0 : ((l < ln && tau != 0.0) ? x * scale : 0.0);
          if (l == 0) sh_tau = tau;
          br[0] = l == 0 ? beta : 0.0;
#pragma unroll
          for (int q = 0; q < 32; ++q) B[l][q] = br[q];
        }
      } else if (wid == 1) {
        // ---- diagonal block r0 .. r0+ln-1: mirror the lower triangle the loaders brought ----------------------------
#pragma unroll 8
        for (int q = 0; q < 32; ++q)
          if (q > l) D[l][q] = D[q][l];
      }
      nb_sync(5, kLLCompute);
      const double tau = sh_tau;
      const double vl = vs[l];
      if (wid == 0) {
        if (k > 0) {
          // ---- left-apply the new reflector to columns 1 .. lp-1 (lane = column), then store the block by rows ----
          // A step is ~900 instructions of THIS warp and its time is that count (two warps per scheduler at most: every
          // instruction costs its 4 - 6 cycles of dependent issue), so the data movement is kept minimal: the column is read
          // from shared memory once for the dot product and the update, and the 32 stores of a lane go to base + q (kLdb - 1)
          // with the offsets as immediates (the full block - the common case - without a predicate per element).  Same
          // operations in the same order as k_sbr_chase_t: the results stay bit-identical.
          if (l >= 1 && l < lp) {
            double bc[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) bc[i] = B[i][l];
            double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              y0 = fma(vs[i], bc[i], y0);
              y1 = fma(vs[i + 1], bc[i + 1], y1);
              y2 = fma(vs[i + 2], bc[i + 2], y2);
              y3 = fma(vs[i + 3], bc[i + 3], y3);
            }
            const double y = tau * ((y0 + y1) + (y2 + y3));
#pragma unroll
            for (int i = 0; i < 32; ++i) B[i][l] = fma(-vs[i], y, bc[i]);
          }
          __syncwarp();
          if (l < ln) {
            ulonglong2* const base = a.Bt + (size_t)(lp + l) + (size_t)st * kLdb;   // (lp + l - q) + (st + q) kLdb = base + q (kLdb - 1)
            if (lp == 32) {
#pragma unroll
              for (int q = 0; q < 32; ++q) ll_store(base + q * (kLdb - 1), B[l][q], wtag);
            } else {
#pragma unroll
              for (int q = 0; q < 32; ++q)
                if (q < lp) ll_store(base + q * (kLdb - 1), B[l][q], wtag);
            }
          }
        }
        vps[l] = vl;                      // the reflector the next step applies from the right
      } else if (wid == 1) {
        // ---- type 3: D <- H D H ---------------------------------------------------------------------------------
        double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          p0 = fma(D[l][q], vs[q], p0);
          p1 = fma(D[l][q + 1], vs[q + 1], p1);
          p2 = fma(D[l][q + 2], vs[q + 2], p2);
          p3 = fma(D[l][q + 3], vs[q + 3], p3);
        }
        const double p = tau * ((p0 + p1) + (p2 + p3));
        const double w = fma(-0.5 * tau * wsum(p * vl), vl, p);
        ws[l] = w;
        __syncwarp();
        if (l < ln) {
          ulonglong2* const base = a.Bt + (size_t)l + (size_t)r0 * kLdb;             // (l - q) + (r0 + q) kLdb = base + q (kLdb - 1)
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const double val = D[l][q] - vl * ws[q] - w * vs[q];
            if (q <= l) ll_store(base + q * (kLdb - 1), val, wtag);
          }
        }
      } else if (a.L > 0) {
        // ---- right-hand sides: z[r0 .. r0+ln-1, :] <- H z -----------------------------------------------------------
        if (l < a.L) {
          double t0 = 0.0, t1 = 0.0;
#pragma unroll 8
          for (int i = 0; i < 32; i += 2) {
            t0 = fma(vs[i], zs[i][l], t0);
            t1 = fma(vs[i + 1], zs[i + 1][l], t1);
          }
          ts[l] = tau * (t0 + t1);
        }
        __syncwarp();
        if (l < ln)
          for (int cc = 0; cc < a.L; ++cc) ll_store(a.zt + (size_t)(r0 + l) + (size_t)cc * m, fma(-vl, ts[cc], zs[l][cc]), wtag);
      }
      taup = tau;
      nb_sync(5, kLLCompute);             // vs / zs / ts are rewritten by the next step
      if (k + 2 < tot) nb_arrive(3 + buf, kLLThreads);   // stage `buf` may be refilled (step k + 2); nobody waits after that
    }
  }
}

// plain <-> tagged storage (tag 0 = "as stage 1 left it")
__global__ void k_sbr_tag(const double* __restrict__ src, ulonglong2* __restrict__ dst, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ll_store(dst + i, src[i], 0u);
}
__global__ void k_sbr_untag(const ulonglong2* __restrict__ src, double* __restrict__ dst, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const ulonglong2 w = src[i];
  dst[i] = __longlong_as_double((long long)((w.x & 0xffffffffull) | (w.y << 32)));
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace

// ---------------------------------------------------------------------------------------------
// Coefficients from the band form ("coef_impl" = 1; experimental, see DESIGN.md section 9 and tools/proto_two_stage.py
// coefficients_from_band):  (M + lam I)^-1 z = Q1 (B + lam I)^-1 Q1'z.
// k_band_solve: ONE CTA.  Block Cholesky of B + lam I with 32 x 32 blocks (diagonal block potrf, its inverse,
// L_{j+1,j} = A_{j+1,j} L_jj^-T, next diagonal block -= L_{j+1,j} L_{j+1,j}'), factors kept in global memory, then forward and
// backward substitution of one right-hand side as 32 x 32 matrix-vector products.  O(m b^2) flops, m / 32 dependent steps.
// ---------------------------------------------------------------------------------------------
struct BandSolveArgs {
  const double* Bd; int m; double lam;
  double* fac;                     // [nb][2][1024]: Linv_j (row-major, lower), Ls_j = L_{j+1,j} (row-major)
  double* x;                       // nb * 32 doubles: right-hand side in, solution out (rows >= m are zero)
  int* err;                        // set to 1 if a pivot is not positive
};

// One CTA of 8 warps.  The sequential part of a block step - potrf of the 32 x 32 diagonal block and the inverse of its factor -
// stays on warp 0 (fully unrolled, right-looking, operands in registers: see the comments there); the two 32 x 32 x 32 products (L_{j+1,j} = A_{j+1,j} L_jj^-T, next diagonal block -= L_{j+1,j} L_{j+1,j}') and all
// block loads / stores are spread over the 256 threads (thread = row x 4 columns).  Every output element is ONE fma chain in
// ascending k (deterministic; 9.4 ms at 5 000 knots on one warp, 6.9 ms with the CTA-wide products, r3: right-looking warp part).
constexpr int kBandThreads = 256;
__global__ void __launch_bounds__(kBandThreads) k_band_solve(BandSolveArgs a) {
  __shared__ double B0[32][kPad], B1[32][kPad], B2[32][kPad];
  __shared__ double xs[32], tv[32], xn[32];
  __shared__ __align__(16) double colb[2][32];
  const int tid = threadIdx.x, l = tid & 31, wq = (tid >> 5) * 4, m = a.m;      // this thread: row l, columns wq .. wq + 3
  constexpr int b = 32;
  const int nb = (m + b - 1) / b;
  double (*D)[kPad] = B0;          // current diagonal block / its Cholesky factor, then L_{j+1,j}
  double (*S)[kPad] = B1;          // sub-diagonal block, then the next diagonal block
  double (*Li)[kPad] = B2;         // inverse of the diagonal factor
  // diagonal block J of B + lam I, lower part, identity padding past m (this thread's 4 elements)
  auto load_diag = [&](double (*dst)[kPad], int J) {
    const int i = J * b + l;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
      const int q = wq + qq, j = J * b + q;
      double v = 0.0;
      if (q <= l) {
        if (i < m) v = __ldcg(a.Bd + (i - j) + (size_t)j * kLdb) + (q == l ? a.lam : 0.0);
        else v = q == l ? 1.0 : 0.0;
      }
      dst[l][q] = v;
    }
  };
  load_diag(D, 0);
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    if (tid < 32) {
      // ---- D = L L' (lower), lane = row: the row lives in registers, a column travels through a 32-double buffer ---------
      // Right-looking and fully unrolled.  Per column the dependent chain is shuffle (pivot) -> rsqrt -> one multiply -> buffer ->
      // one FMA; sqrt(d) = d rsqrt(d) and 1 / sqrt(d) = rsqrt(d) replace the square root and 32 divisions of the first version,
      // and the 31 - c updates of a lane are independent register FMAs (upper-triangle entries ride along unpredicated: they
      // start as zeros and are never read by another lane).
      double r[32], rs[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) r[q] = D[l][q];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const double dcc = __shfl_sync(0xffffffffu, r[c], c);
        if (!(dcc > 0.0) && l == 0) *a.err = 1;
        rs[c] = rsqrt(dcc > 0.0 ? dcc : 1.0);
        const double lic = r[c] * rs[c];              // lane c: the diagonal entry sqrt(d); lanes > c: L[l][c]
        r[c] = lic;
        colb[c & 1][l] = lic;
        __syncwarp();
#pragma unroll
        for (int q = c + 1; q < 32; ++q) r[q] = fma(-lic, colb[c & 1][q], r[q]);
      }
#pragma unroll
      for (int q = 0; q < 32; ++q) D[l][q] = r[q];
      __syncwarp();
      // ---- Li = L^-1, lane = column: right-looking too - once Li[k][l] is known every later row takes its term, so the
      // chain per row is one FMA + one multiply by rsqrt(pivot) instead of a dot product of up to 31 dependent FMAs ---------
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = i == l ? 1.0 : 0.0;      // r = pending right-hand side of column l
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const double xk = r[k] * rs[k];               // zero above the diagonal (k < l) without a predicate
        Li[k][l] = xk;
#pragma unroll
        for (int i = k + 1; i < 32; ++i) r[i] = fma(-D[i][k], xk, r[i]);
      }
    }
    __syncthreads();
    double* fj = a.fac + (size_t)j * 2048;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) fj[l * 32 + wq + qq] = Li[l][wq + qq];
    if (j + 1 < nb) {
      // ---- S = A_{j+1,j}: rows (j+1) b + l, columns j b + q, inside the band for l <= q -------------------------
      {
        const int i = (j + 1) * b + l;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const int q = wq + qq, jj = j * b + q;
          S[l][q] = (l <= q && i < m) ? __ldcg(a.Bd + (i - jj) + (size_t)jj * kLdb) : 0.0;
        }
      }
      __syncthreads();
      // ---- Ls = S Li' (into D: the factor itself is not needed any more) -----------------------------------------
      {
        double acc4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const double sv = S[l][k];
#pragma unroll
          for (int qq = 0; qq < 4; ++qq)
            if (k <= wq + qq) acc4[qq] = fma(sv, Li[wq + qq][k], acc4[qq]);
        }
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) D[l][wq + qq] = acc4[qq];
      }
      __syncthreads();
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) fj[1024 + l * 32 + wq + qq] = D[l][wq + qq];
      // ---- next diagonal block = A_{j+1,j+1} + lam I - Ls Ls' (into S) ---------------------------------------------
      load_diag(S, j + 1);
      {
        double acc4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const double dv = D[l][k];
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) acc4[qq] = fma(dv, D[wq + qq][k], acc4[qq]);
        }
#pragma unroll
        for (int qq = 0; qq < 4; ++qq)
          if (wq + qq <= l) S[l][wq + qq] -= acc4[qq];
      }
      __syncthreads();
      double (*tmp)[kPad] = D; D = S; S = tmp;
    }
  }
  __syncthreads();
  __threadfence_block();
  // ---- forward substitution: t = Linv_j x_j ; x_{j+1} -= Ls_j t ---------------------------------------------------
  for (int j = 0; j < nb; ++j) {
    const double* fj = a.fac + (size_t)j * 2048;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
      Li[l][wq + qq] = fj[l * 32 + wq + qq];
      if (j + 1 < nb) S[l][wq + qq] = fj[1024 + l * 32 + wq + qq];
    }
    if (tid < 32) xs[l] = a.x[j * b + l];
    __syncthreads();
    if (tid < 32) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k <= l) t = fma(Li[l][k], xs[k], t);
      tv[l] = t;
      a.x[j * b + l] = t;
      __syncwarp();
      if (j + 1 < nb) {
        double u = 0.0;
#pragma unroll
        for (int k = 0; k < 32; ++k) u = fma(S[l][k], tv[k], u);
        a.x[(j + 1) * b + l] -= u;
      }
    }
    __syncthreads();
  }
  // ---- backward substitution: x_j -= Ls_j' x_{j+1} ; x_j = Linv_j' x_j ------------------------------------------------
  for (int j = nb - 1; j >= 0; --j) {
    const double* fj = a.fac + (size_t)j * 2048;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
      Li[l][wq + qq] = fj[l * 32 + wq + qq];
      if (j + 1 < nb) S[l][wq + qq] = fj[1024 + l * 32 + wq + qq];
    }
    if (tid < 32) {
      xs[l] = a.x[j * b + l];
      xn[l] = j + 1 < nb ? a.x[(j + 1) * b + l] : 0.0;
    }
    __syncthreads();
    if (tid < 32) {
      double v = xs[l];
      if (j + 1 < nb) {
#pragma unroll
        for (int k = 0; k < 32; ++k) v = fma(-S[k][l], xn[k], v);
      }
      tv[l] = v;
      __syncwarp();
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k >= l) t = fma(Li[k][l], tv[k], t);
      a.x[j * b + l] = t;
    }
    __syncthreads();
  }
}

// k_apply_q1: y <- Q1 y = H_1 ... H_K y, H_k = I - V_k T_k V_k', last panel first.  One thread-block cluster, thread = row of
// the panel (as in the panel QR): V_k'y is reduced inside the warp by the reduce-scatter butterfly, across warps through shared
// memory, across the CTAs through distributed shared memory; u = T_k (V_k'y); y -= V_k u.
struct ApplyQ1Args {
  const double* Vall; const double* Tall;
  const unsigned long long* voff; const int* rk; const int* ldvk; const int* r0k;   // device arrays, one entry per panel
  int npanel;
  double* y;                       // m doubles
};

__global__ void __launch_bounds__(kQrRows, 1) k_apply_q1(ApplyQ1Args a) {
  __shared__ double red[16 * 32], Tsh[32 * kPad], tsh[32], ush[32], part[2 * 32];
  cg::cluster_group cluster = cg::this_cluster();
  const int t = threadIdx.x, blk = blockIdx.x, G = gridDim.x, lane = t & 31, wid = t >> 5;
  const int gi = blk * kQrRows + t;
  for (int k = a.npanel - 1; k >= 0; --k) {
    const int r = a.rk[k], ldv = a.ldvk[k], r0 = a.r0k[k];
    const double* V = a.Vall + a.voff[k];
    const double* T = a.Tall + (size_t)1024 * k;
    const bool live = gi < r;
    double vr[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) vr[q] = live ? V[gi + (size_t)q * ldv] : 0.0;
    for (int i = t; i < 1024; i += kQrRows) Tsh[(i >> 5) * kPad + (i & 31)] = T[i];
    const double yv = live ? __ldcg(a.y + r0 + gi) : 0.0;
    // ---- t_q = sum over rows of V[row][q] y[row] ----------------------------------------------------------------
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = vr[16 * h + i] * yv;
#pragma unroll
      for (int o = 16, n = 8; n >= 1; o >>= 1, n >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
          const double send = up ? v[i] : v[i + n];
          const double keepv = up ? v[i + n] : v[i];
          v[i] = keepv + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
      if ((lane & 1) == 0) red[wid * 32 + 16 * h + (lane >> 1)] = v[0];
    }
    __syncthreads();
    double* mine = part + (k & 1) * 32;
    if (t < 32) {
      double p = 0.0;
#pragma unroll
      for (int q = 0; q < 16; ++q) p += red[q * 32 + t];
      mine[t] = p;
    }
    cluster.sync();
    if (t < 32) {
      double pr[16], s = 0.0;
#pragma unroll
      for (int q = 0; q < 16; ++q) pr[q] = q < G ? cluster.map_shared_rank(mine, q)[t] : 0.0;
#pragma unroll
      for (int q = 0; q < 16; ++q) s += pr[q];
      tsh[t] = s;
    }
    __syncthreads();
    if (t < 32) {
      double u = 0.0;
      for (int q = 0; q < 32; ++q) u = fma(Tsh[t * kPad + q], tsh[q], u);
      ush[t] = u;
    }
    __syncthreads();
    if (live) {
      double acc = 0.0;
#pragma unroll
      for (int q = 0; q < 32; ++q) acc = fma(vr[q], ush[q], acc);
      a.y[r0 + gi] = yv - acc;
    }
    __threadfence();
    cluster.sync();                 // the next panel reads rows that other CTAs have just written
  }
}

bool band_coefficients(mb_ctx* ctx, double lambda, int rhs, double* out_dev, cudaStream_t st) {
  mb_band_form& bf = ctx->band_form;
  if (!bf.valid || rhs < 0 || rhs >= bf.L) return false;
  const int m = bf.m, nb = ceil_div(m, 32);
  // the back-transformation runs as one cluster: the largest panel must fit into 16 x 512 rows
  int cs = 1;
  while (cs * kQrRows < std::max(bf.rmax, 1)) cs *= 2;
  if (bf.npanel > 0 && cs > 16) return false;
  Arena& ar = ctx->arena;
  double* fac = ar.take_n<double>((size_t)nb * 2048);
  double* x = ar.take_n<double>((size_t)nb * 32);
  int* err = ar.take_n<int>(1);
  MB_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * nb * 32, st));
  MB_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
  MB_CUDA(cudaMemcpyAsync(x, bf.z1 + (size_t)rhs * m, sizeof(double) * m, cudaMemcpyDeviceToDevice, st));
  BandSolveArgs sa{bf.band, m, lambda, fac, x, err};
  MB_LAUNCH(ctx, "k_band_solve", st) k_band_solve<<<1, kBandThreads, 0, st>>>(sa);
  if (bf.npanel > 0) {
    std::vector<unsigned long long> voff(bf.voff.begin(), bf.voff.end());
    ApplyQ1Args qa{bf.Vall, bf.Tall, ar.upload(voff.data(), voff.size(), st), ar.upload(bf.r.data(), bf.r.size(), st),
                   ar.upload(bf.ldv.data(), bf.ldv.size(), st), ar.upload(bf.r0.data(), bf.r0.size(), st), bf.npanel, x};
    static thread_local bool attr = false;
    if (!attr) { (void)cudaFuncSetAttribute(k_apply_q1, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); attr = true; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs); cfg.blockDim = dim3(kQrRows); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    MB_LAUNCH(ctx, "k_apply_q1", st) MB_CUDA(cudaLaunchKernelEx(&cfg, k_apply_q1, qa));
  }
  MB_CUDA(cudaMemcpyAsync(out_dev, x, sizeof(double) * m, cudaMemcpyDeviceToDevice, st));
  int herr = 0;
  MB_CUDA(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));      // also keeps the uploaded panel tables alive until the kernels are done
  return herr == 0;     // a non-positive pivot (B + lambda I numerically indefinite): the caller falls back to the dense Cholesky
}

// Reduces the symmetric m x m matrix at A (column-major, leading dimension ld >= m, LOWER triangle valid on entry, both
// triangles destroyed) to tridiagonal form: d (m) and e (m - 1) on the device; the L right-hand sides z (m x L, ld = m)
// become Q'z.  Synchronises st.
void sym_band_tridiag(mb_ctx* ctx, double* A, int ld, int m, double* z, int L, double* d, double* e, cudaStream_t st) {
  MB_REQUIRE(m >= 3 && ld >= m && ld % 2 == 0, "two-stage tridiagonalisation: bad matrix shape");
  MB_REQUIRE(L >= 0 && L <= 32, "at most 32 right-hand sides per tridiagonalisation");
  Arena& ar = ctx->arena;
  ctx->band_form = mb_band_form();
  // cluster size the panel QR may use: 16 (non-portable) if such a cluster can be scheduled, else 8, else none
  static thread_local int max_cluster = -1;
  if (max_cluster < 0) {
    MB_CUDA(cudaFuncSetAttribute(k_sbr_qr<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kQrSmem));
    MB_CUDA(cudaFuncSetAttribute(k_sbr_qr<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kQrSmem));
    (void)cudaFuncSetAttribute(k_sbr_qr_reg<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    MB_CUDA(cudaFuncSetAttribute(k_sbr_r2k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kR2kSmem));
    MB_CUDA(cudaFuncSetAttribute(k_sbr_vtz, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kVtzSmem));
    max_cluster = 0;
    const bool np_ok = cudaFuncSetAttribute(k_sbr_qr<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    for (int cs : {16, 8}) {
      if (cs > 8 && !np_ok) continue;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs); cfg.blockDim = dim3(kQrRows); cfg.dynamicSmemBytes = kQrSmem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, k_sbr_qr<true>, &cfg) == cudaSuccess && nclusters >= 1) {
        max_cluster = cs;
        break;
      }
    }
    (void)cudaGetLastError();
    if (ctx->sbr_debug) std::fprintf(stderr, "[sbr] panel QR cluster size limit: %d\n", max_cluster);
  }
  const int rmax = m - kBw;
  if (rmax >= 2) {
    const int nt = ceil_div(m, 32);
    MB_LAUNCH(ctx, "k_sbr_mirror", st) k_sbr_mirror<<<dim3(nt, nt), 256, 0, st>>>(A, ld, m);
    const int rpadmax = ceil_div(rmax, 128) * 128;
    const int gmax = ceil_div(rmax, kQrRows);
    int npanel = 0;
    for (int j0 = 0; m - j0 - kBw >= 2; j0 += kBw) ++npanel;
    const bool keep = ctx->coef_impl != 2;
    mb_band_form& bf = ctx->band_form;
    if (keep) {
      size_t tot = 0;
      for (int j0 = 0; m - j0 - kBw >= 2; j0 += kBw) {
        const int r = m - j0 - kBw, ldv = ceil_div(r, 128) * 128;
        bf.voff.push_back(tot); bf.r.push_back(r); bf.ldv.push_back(ldv); bf.r0.push_back(j0 + kBw);
        tot += (size_t)ldv * 32;
      }
      bf.npanel = npanel;
      bf.rmax = rmax;
      bf.Vall = ar.take_n<double>(tot);
      bf.Tall = ar.take_n<double>((size_t)1024 * npanel);
    }
    double* Vb[2] = {nullptr, nullptr};
    if (!keep) { Vb[0] = ar.take_n<double>((size_t)rpadmax * 32); Vb[1] = ar.take_n<double>((size_t)rpadmax * 32); }
    double* Wb[2] = {ar.take_n<double>((size_t)rpadmax * 32), ar.take_n<double>((size_t)rpadmax * 32)};
    double* Z0 = ar.take_n<double>((size_t)rpadmax * 32);
    double* Zp = ar.take_n<double>((size_t)8 * rpadmax * 32);
    double* T = ar.take_n<double>(1024);
    double* ST = ar.take_n<double>(2048);
    double* Gp = ar.take_n<double>((size_t)(rpadmax / 128) * 2048);
    double* slots = ar.take_n<double>((size_t)2 * (gmax + 1) * 32);
    unsigned* bars = ar.take_n<unsigned>((size_t)64 * npanel);
    MB_CUDA(cudaMemsetAsync(bars, 0, sizeof(unsigned) * 64 * npanel, st));
    // look-ahead: the trailing update of panel k runs on a second stream beside the QR of panel k + 1
    if (!ctx->sbr_aux) {
      int prio_lo = 0, prio_hi = 0;
      MB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      MB_CUDA(cudaStreamCreateWithPriority(&ctx->sbr_aux, cudaStreamNonBlocking, prio_hi));
      for (cudaEvent_t& ev : ctx->sbr_ev) MB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
    cudaStream_t aux = ctx->sbr_aux;
    // inside mb_mltps_predict* with SM partitions (greenctx.cu): stage 1 runs on the streams of the FIT partition
    cudaStream_t s1 = st;
    const bool gc = ctx->gc_stage1 && ctx->gc_ok;
    int cluster_limit = max_cluster;
    if (gc) {
      MB_CUDA(cudaEventRecord(ctx->gc_ev[0], st));
      MB_CUDA(cudaStreamWaitEvent(ctx->gc_fit_stream, ctx->gc_ev[0], 0));
      s1 = ctx->gc_fit_stream;
      aux = ctx->gc_fit_aux;
      ar.also_used(s1);
      ar.also_used(aux);
      static thread_local int max_cluster_gc = -1;         // what the partition can co-schedule (<= the device-wide limit)
      if (max_cluster_gc < 0) {
        max_cluster_gc = 0;
        for (int cs : {16, 8, 4, 2}) {
          if (cs > max_cluster && cs > 4) continue;
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(cs); cfg.blockDim = dim3(kQrRows); cfg.dynamicSmemBytes = kQrSmem; cfg.stream = s1;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension;
          at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          int nclusters = 0;
          if (cudaOccupancyMaxActiveClusters(&nclusters, k_sbr_qr<true>, &cfg) == cudaSuccess && nclusters >= 1) { max_cluster_gc = cs; break; }
        }
        (void)cudaGetLastError();
        if (ctx->sbr_debug) std::fprintf(stderr, "[sbr] panel QR cluster size limit on the fit partition (%d SMs): %d\n", ctx->gc_fit_sms, max_cluster_gc);
      }
      cluster_limit = std::min(max_cluster, max_cluster_gc);
    }
    int pk = 0;
    for (int j0 = 0; m - j0 - kBw >= 2; j0 += kBw, ++pk) {
      const int r = m - j0 - kBw;
      const int ldv = ceil_div(r, 128) * 128;
      const int rblocks = ldv / 128;
      double* V = keep ? bf.Vall + bf.voff[pk] : Vb[pk & 1];
      double* W = Wb[pk & 1];
      double* Tk = keep ? bf.Tall + (size_t)1024 * pk : T;
      QrArgs qa{A, ld, m, j0, V, ldv, Tk, slots, bars + (size_t)64 * pk};
      const int gq = ceil_div(r, kQrRows);
      if (gq <= cluster_limit && ctx->sbr_qr_grid == 0) {
        int cs = 1;
        while (cs < gq) cs *= 2;
        cudaLaunchConfig_t cfg = {};
        // measured (profiles/r1v_two_stage_check.txt): rows in registers win up to clusters of 4 CTAs (2.6 vs 3.3 us per
        // column), rows in shared memory at 8 and 16 (3.4 vs 4.0 us)
        const bool reg = ctx->sbr_qr_impl == 2 || (ctx->sbr_qr_impl == 0 && cs <= 4);
        cfg.gridDim = dim3(cs); cfg.blockDim = dim3(kQrRows); cfg.dynamicSmemBytes = reg ? kQrRegSmem : kQrSmem; cfg.stream = s1;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (reg) {
          MB_LAUNCH(ctx, "k_sbr_qr", s1) MB_CUDA(cudaLaunchKernelEx(&cfg, k_sbr_qr_reg<true>, qa));
        } else {
          MB_LAUNCH(ctx, "k_sbr_qr_smem", s1) MB_CUDA(cudaLaunchKernelEx(&cfg, k_sbr_qr<true>, qa));
        }
      } else {
        void* params[] = {&qa};
        if (ctx->sbr_qr_impl != 1) {
          MB_LAUNCH(ctx, "k_sbr_qr_grid", s1)
            MB_CUDA(cudaLaunchCooperativeKernel((const void*)k_sbr_qr_reg<false>, dim3(gq), dim3(kQrRows), params, kQrRegSmem, s1));
        } else {
          MB_LAUNCH(ctx, "k_sbr_qr_smem_grid", s1)
            MB_CUDA(cudaLaunchCooperativeKernel((const void*)k_sbr_qr<false>, dim3(gq), dim3(kQrRows), params, kQrSmem, s1));
        }
      }
      double* A22 = A + (size_t)(j0 + kBw) * ((size_t)ld + 1);
      int nsplit = std::max(1, std::min(8, ceil_div(2 * ctx->sm_count, rblocks)));
      const int chunk = ceil_div(ceil_div(r, nsplit), 16) * 16;
      nsplit = ceil_div(r, chunk);
      if (pk > 0) MB_CUDA(cudaStreamWaitEvent(s1, ctx->sbr_ev[2 + ((pk - 1) & 1)], 0));      // trailing update of panel k - 1
      MB_LAUNCH(ctx, "k_sbr_av", s1) k_sbr_av<<<dim3(rblocks, nsplit), 128, 0, s1>>>(A22, ld, r, V, ldv, Zp, chunk);
      // (vtz + st and w + pu as two fused launches were measured in r2q: 36 + 36 us per panel against 24 + 13 + 14 + 24 -
      //  the last-CTA epilogue and the recomputed W rows cost what the two launches save; deleted)
      MB_LAUNCH(ctx, "k_sbr_vtz", s1)
        k_sbr_vtz<<<rblocks, 512, kVtzSmem, s1>>>(V, ldv, r, Zp, nsplit, Z0, z, m, j0 + kBw, L, Gp);
      MB_LAUNCH(ctx, "k_sbr_st", s1) k_sbr_st<<<1, 1024, 0, s1>>>(Gp, rblocks, Tk, ST);
      MB_LAUNCH(ctx, "k_sbr_w", s1) k_sbr_w<<<rblocks, 128, 0, s1>>>(V, ldv, r, Z0, Tk, ST, W, z, m, j0 + kBw, L);
      if (r > 32) MB_LAUNCH(ctx, "k_sbr_pu", s1) k_sbr_pu<<<ceil_div(r - 32, 128), 128, 0, s1>>>(A22, ld, r, V, W, ldv);
      MB_CUDA(cudaEventRecord(ctx->sbr_ev[pk & 1], s1));
      MB_CUDA(cudaStreamWaitEvent(aux, ctx->sbr_ev[pk & 1], 0));
      MB_LAUNCH(ctx, "k_sbr_r2k", aux)
        k_sbr_r2k<<<rblocks * (rblocks + 1) / 2, 256, kR2kSmem, aux>>>(A22, ld, r, V, W, ldv);
      MB_CUDA(cudaEventRecord(ctx->sbr_ev[2 + (pk & 1)], aux));
    }
    if (pk > 0) MB_CUDA(cudaStreamWaitEvent(s1, ctx->sbr_ev[2 + ((pk - 1) & 1)], 0));
    if (gc) {                                      // back to the caller's stream (all SMs) for stage 2 and everything after it
      MB_CUDA(cudaEventRecord(ctx->gc_ev[1], s1));
      MB_CUDA(cudaStreamWaitEvent(st, ctx->gc_ev[1], 0));
    }
    MB_CUDA(cudaGetLastError());
  }
  if (ctx->after_stage1) ctx->after_stage1();
  // ---- stage 2 ------------------------------------------------------------------------------------------------------
  const int ncolb = m + kLdb;
  double* Bd = ar.take_n<double>((size_t)kLdb * ncolb);
  int* prog = ar.take_n<int>((size_t)m + 2);
  MB_CUDA(cudaMemsetAsync(prog, 0, sizeof(int) * ((size_t)m + 2), st));
  MB_LAUNCH(ctx, "k_sbr_band", st) k_sbr_band<<<ceil_div(ncolb, 4), 256, 0, st>>>(A, ld, m, Bd, ncolb);
  if (ctx->coef_impl != 2 && L > 0) {
    // band form for the coefficient solve: B and Q1'z as they are now (the chase below overwrites both)
    mb_band_form& bf = ctx->band_form;
    bf.m = m; bf.L = L;
    bf.band = ar.take_n<double>((size_t)kLdb * ncolb);
    bf.z1 = ar.take_n<double>((size_t)m * L);
    MB_CUDA(cudaMemcpyAsync(bf.band, Bd, sizeof(double) * kLdb * ncolb, cudaMemcpyDeviceToDevice, st));
    MB_CUDA(cudaMemcpyAsync(bf.z1, z, sizeof(double) * (size_t)m * L, cudaMemcpyDeviceToDevice, st));
    bf.valid = true;
  }
  if (ctx->sbr_debug) {
    ctx->dbg_band.resize((size_t)kLdb * m);
    MB_CUDA(cudaMemcpyAsync(ctx->dbg_band.data(), Bd, sizeof(double) * kLdb * m, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
  }
  ChaseArgs ca{Bd, m, z, L, prog, (unsigned)(ctx->sbr_chase_sleep < 0 ? 0 : ctx->sbr_chase_sleep)};
  // one CTA per sweep in flight: at most m / (2 b) sweeps can overlap; "sbr_chase_ctas" caps the grid (fewer SMs held by the chase
  // beside the ensemble kernels, at the price of a shallower pipeline at the start)
  int G = std::max(1, std::min(ctx->sm_count, ceil_div(m, 2 * kBw) + 2));
  if (ctx->sbr_chase_ctas > 0) G = std::max(1, std::min(G, ctx->sbr_chase_ctas));
  // Register footprint decides (profiles/r2l_*, r2m_*): the chase sits on 80 SMs for ~50 ms; a CTA with the watcher / publisher
  // warps holds 160 x 192 registers, which costs a co-resident per-cell kernel (ksvm: 4 CTAs of 16 K registers = the whole file)
  // TWO of its four CTAs on that SM (k_ens_svm_tma 50 -> 69 ms), the three-warp CTA (96 x 160) one (-> 59 ms).  Spinning is not
  // the cost: back-off in the poll loops ("sbr_chase_sleep") changes nothing.
  // A fit that has the GPU to itself takes the tagged-element chase (34 against 41 ms at 5 000 knots, faster at every size,
  // profiles/r2o3_chase_check.txt); its CTA (224 threads x 255 registers) would leave a co-resident per-cell kernel nothing,
  // so beside the ensemble the three-warp CTA stays.
  const bool dec = ctx->sbr_chase_impl == 1;
  if (ctx->sbr_chase_impl == 3 || ctx->sbr_chase_impl == 4 || (ctx->sbr_chase_impl == 0 && (!ctx->fit_shares_gpu || (ctx->gc_stage1 && ctx->gc_ok)))) {
    // tagged copies of the band and the right-hand sides, chase, plain copies back (2 x 5 MB at 5 000 knots: ~10 us each)
    const size_t nb = (size_t)kLdb * ncolb, nz = (size_t)m * std::max(L, 0);
    ulonglong2* Bt = ar.take_n<ulonglong2>(nb);
    ulonglong2* zt = ar.take_n<ulonglong2>(std::max<size_t>(nz, 1));
    k_sbr_tag<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(Bd, Bt, nb);
    if (nz) k_sbr_tag<<<(unsigned)((nz + 255) / 256), 256, 0, st>>>(z, zt, nz);
    ChaseLLArgs la{Bt, m, zt, L, prog + m};
    if (ctx->sbr_chase_impl == 4) {
      MB_LAUNCH(ctx, "k_sbr_chase_ll2", st) k_sbr_chase_ll<2><<<G, kLLThreads, 0, st>>>(la);
    } else {
      MB_LAUNCH(ctx, "k_sbr_chase_ll", st) k_sbr_chase_ll<1><<<G, kLLThreads, 0, st>>>(la);
    }
    k_sbr_untag<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(Bt, Bd, nb);
    if (nz) k_sbr_untag<<<(unsigned)((nz + 255) / 256), 256, 0, st>>>(zt, z, nz);
  } else if (dec) {
    MB_LAUNCH(ctx, "k_sbr_chase_dec", st) k_sbr_chase_t<true><<<G, kChaseThreadsDec, 0, st>>>(ca);
  } else {
    MB_LAUNCH(ctx, "k_sbr_chase", st) k_sbr_chase_t<false><<<G, kChaseThreads, 0, st>>>(ca);
  }
  MB_LAUNCH(ctx, "k_sbr_diag", st) k_sbr_diag<<<ceil_div(m, 256), 256, 0, st>>>(Bd, m, d, e);
  MB_CUDA(cudaGetLastError());
  int chase_err = 0;
  MB_CUDA(cudaMemcpyAsync(&chase_err, prog + m + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
  if (chase_err) throw Error(MB_E_NUMERIC, "two-stage tridiagonalisation: a bulge-chasing sweep waited too long for its predecessor");
}

}  // namespace mb
