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

#include <algorithm>
#include <cmath>

namespace mb {

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
  double* slots;                   // 2 x (G + 1) x 32: per-CTA partial sums (+ the pivot row), double-buffered by column parity
  unsigned* bar;
};
constexpr size_t kQrSmem = sizeof(double) * (kQrRows * kPad + 16 * 32 + 32 + 32 + 32 * kPad);

__global__ void __launch_bounds__(kQrRows, 1) k_sbr_qr(QrArgs a) {
  extern __shared__ double sm_qr[];
  double* X = sm_qr;                      // [512][33]   row t = this thread's row of the panel
  double* red = X + kQrRows * kPad;       // [16][32]
  double* s_sh = red + 16 * 32;           // [32]
  double* piv_sh = s_sh + 32;             // [32]
  double* Tsh = piv_sh + 32;              // [32][33]
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
    double acc = 0.0;
    const int rbase = seg * 32;
#pragma unroll 8
    for (int rr = 0; rr < 32; ++rr) {
      const int rw = rbase + rr;
      if (blk * kQrRows + rw > j) acc = fma(X[rw * kPad + j], X[rw * kPad + c], acc);
    }
    red[seg * 32 + c] = acc;
    __syncthreads();
    double* slot = a.slots + (size_t)(j & 1) * (G + 1) * 32;
    if (t < 32) {
      double p = 0.0;
#pragma unroll
      for (int q = 0; q < 16; ++q) p += red[q * 32 + t];
      slot[blk * 32 + t] = p;
      if (blk == 0) slot[G * 32 + t] = X[j * kPad + t];     // the pivot row
    }
    sbr_grid_sync(a.bar, gen);
    if (t < 32) {
      double s = 0.0;
      for (int q = 0; q < G; ++q) s += __ldcg(slot + q * 32 + t);
      s_sh[t] = s;
      piv_sh[t] = __ldcg(slot + G * 32 + t);
    }
    __syncthreads();
    const double alpha = piv_sh[j];
    double beta, tau, scale;
    make_house(alpha, s_sh[j], beta, tau, scale);
    // ---- column j of T (CTA 0, warp 1): T[:j, j] = -tau T[:j, :j] (V[:, :j]' v_j) --------------------------
    if (blk == 0 && t >= 32 && t < 64) {
      const int q = t - 32;
      if (q < j) {
        double sum = 0.0;
        for (int p = q; p < j; ++p) sum = fma(Tsh[q * kPad + p], fma(scale, s_sh[p], piv_sh[p]), sum);
        Tsh[q * kPad + j] = -tau * sum;
      } else if (q == j) {
        Tsh[j * kPad + j] = tau;
      }
    }
    // ---- apply H_j to the remaining columns of this thread's row ------------------------------------------
    if (gi > j) {
      const double vj = X[t * kPad + j] * scale;
      const double f = tau * vj;
      for (int cc = j + 1; cc < 32; ++cc) X[t * kPad + cc] -= f * fma(scale, s_sh[cc], piv_sh[cc]);
      X[t * kPad + j] = vj;
    } else if (gi == j) {
      for (int cc = j + 1; cc < 32; ++cc) X[t * kPad + cc] -= tau * fma(scale, s_sh[cc], piv_sh[cc]);
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
}

// ---------------------------------------------------------------------------------------------
// Z0 = A22 V : CTA = 128 rows x 32 columns over the k range [sp * chunk, (sp + 1) * chunk); thread = 8 x 4 outputs
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
  for (int kb = k0; kb < k1; kb += 16) {
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const int k = kb + kk;
      As[kk][t] = (gi < r && k < k1) ? __ldg(A22 + gi + (size_t)k * ld) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = t + 128 * q, kk = idx & 15, cc = idx >> 4;
      const int k = kb + kk;
      Vs[kk][cc] = k < k1 ? __ldg(V + k + (size_t)cc * ldv) : 0.0;
    }
    __syncthreads();
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
// Z0 = sum of the k-split partials; per-CTA partial of G0 = V'Z0 (32 x 32) and gz = V'z (32 x L)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sbr_vtz(const double* __restrict__ V, int ldv, int r, const double* __restrict__ Zp,
                                                 int nsplit, double* __restrict__ Z0, const double* __restrict__ z, int ldz,
                                                 int zrow0, int L, double* __restrict__ Gp) {
  __shared__ double Vs[32][kPad], Zs[32][kPad], zs[32][kPad];
  const int t = threadIdx.x, I = blockIdx.x * 128;
  const int a = t & 31, bq = t >> 5;            // outputs G[a][bq*4 + x]
  double g[4] = {0.0, 0.0, 0.0, 0.0}, gz[4] = {0.0, 0.0, 0.0, 0.0};
  for (int h = 0; h < 4; ++h) {
    const int R0 = I + h * 32;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = t + 256 * q, rw = idx & 31, col = idx >> 5;
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
    for (int rw = 0; rw < 32; ++rw) {
      const double va = Vs[rw][a];
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        g[x] = fma(va, Zs[rw][bq * 4 + x], g[x]);
        gz[x] = fma(va, zs[rw][bq * 4 + x], gz[x]);
      }
    }
    __syncthreads();
  }
  double* out = Gp + (size_t)blockIdx.x * 2048;
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    out[a * 32 + bq * 4 + x] = g[x];
    out[1024 + a * 32 + bq * 4 + x] = gz[x];
  }
}

// ---------------------------------------------------------------------------------------------
// W = Z0 T - 1/2 V (T' G0 T);  z <- z - V T' gz.  Thread = one row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_sbr_w(const double* __restrict__ V, int ldv, int r, const double* __restrict__ Z0,
                                               const double* __restrict__ T, const double* __restrict__ Gp, int nblk,
                                               double* __restrict__ W, double* __restrict__ z, int ldz, int zrow0, int L) {
  __shared__ double Ts[32][kPad], Gs[32][kPad], Xs[32][kPad], gzs[32][kPad], tzs[32][kPad];
  const int t = threadIdx.x;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int idx = t + 128 * q, aa = idx >> 5, cc = idx & 31;
    double g = 0.0, gz = 0.0;
    for (int bl = 0; bl < nblk; ++bl) {
      g += Gp[(size_t)bl * 2048 + idx];
      gz += Gp[(size_t)bl * 2048 + 1024 + idx];
    }
    Ts[aa][cc] = T[idx];
    Gs[aa][cc] = g;
    gzs[aa][cc] = gz;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 8; ++q) {                 // X = G0 T
    const int idx = t + 128 * q, aa = idx >> 5, cc = idx & 31;
    double s = 0.0;
    for (int p = 0; p < 32; ++p) s = fma(Gs[aa][p], Ts[p][cc], s);
    Xs[aa][cc] = s;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 8; ++q) {                 // S = T' X (into Gs), tz = T' gz
    const int idx = t + 128 * q, aa = idx >> 5, cc = idx & 31;
    double s = 0.0, u = 0.0;
    for (int p = 0; p < 32; ++p) {
      s = fma(Ts[p][aa], Xs[p][cc], s);
      u = fma(Ts[p][aa], gzs[p][cc], u);
    }
    Gs[aa][cc] = s;                             // every thread has finished reading G0 (barrier above)
    tzs[aa][cc] = u;
  }
  __syncthreads();
  const int i = blockIdx.x * 128 + t;           // < ldv
  double zr[32], vr[32];
#pragma unroll
  for (int q = 0; q < 32; ++q) {
    zr[q] = Z0[i + (size_t)q * ldv];
    vr[q] = V[i + (size_t)q * ldv];
  }
  for (int cc = 0; cc < 32; ++cc) {
    double w1 = 0.0, w2 = 0.0;
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      w1 = fma(zr[q], Ts[q][cc], w1);
      w2 = fma(vr[q], Gs[q][cc], w2);
    }
    W[i + (size_t)cc * ldv] = w1 - 0.5 * w2;
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
// C -= V W' + W V' on 128 x 128 tiles (full square); thread = 8 x 8 outputs, the operands of the whole tile in shared memory
// ---------------------------------------------------------------------------------------------
constexpr size_t kR2kSmem = sizeof(double) * 4 * 32 * 128;

__global__ void __launch_bounds__(256, 1) k_sbr_r2k(double* __restrict__ C, int ld, int r, const double* __restrict__ V,
                                                    const double* __restrict__ W, int ldv) {
  extern __shared__ __align__(16) double sm_r2k[];
  double (*Vi)[128] = reinterpret_cast<double (*)[128]>(sm_r2k);
  double (*Wi)[128] = reinterpret_cast<double (*)[128]>(sm_r2k + 32 * 128);
  double (*Vj)[128] = reinterpret_cast<double (*)[128]>(sm_r2k + 2 * 32 * 128);
  double (*Wj)[128] = reinterpret_cast<double (*)[128]>(sm_r2k + 3 * 32 * 128);
  const int t = threadIdx.x, I = blockIdx.x * 128, J = blockIdx.y * 128;
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
      if (col < r) {
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
      if (rr + 1 < r) *reinterpret_cast<double2*>(C + rr + (size_t)col * ld) = make_double2(acc[2 * a4][j], acc[2 * a4 + 1][j]);
      else if (rr < r) C[rr + (size_t)col * ld] = acc[2 * a4][j];
    }
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
};

// A wait that does not end within ~2^24 polls (seconds; a step takes microseconds) raises the error flag instead of
// hanging the device: the host turns it into MB_E_NUMERIC.
__device__ __forceinline__ void chase_wait(const int* p, int need, int* err) {
  if ((threadIdx.x & 31) == 0) {
    int v, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    } while (v < need && ++spins < (1 << 24));
    if (v < need) atomicExch(err, 1);
  }
  __syncwarp();
}
__device__ __forceinline__ void chase_post(int* p, int val) {
  __threadfence();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(val) : "memory");
}

__global__ void __launch_bounds__(32) k_sbr_chase(ChaseArgs a) {
  __shared__ double Bs[32][kPad], Ds[32][kPad], zs[32][kPad];
  __shared__ double vs[32], ws[32], vps[32], ts[32];
  const int l = threadIdx.x, m = a.m;
  constexpr int b = kBw;
  for (;;) {
    int s = 0;
    if (l == 0) s = atomicAdd(a.prog + m, 1);
    s = __shfl_sync(0xffffffffu, s, 0);
    if (s >= m - 2) break;
    const int totp = s > 0 ? (m - s + b - 1) / b : 0;     // steps of sweep s - 1
    const int tot = (m - s - 1 + b - 1) / b;              // steps of this sweep: row blocks below the diagonal of column s
    double tau = 0.0;
    for (int k = 0; k < tot; ++k) {
      if (s > 0) chase_wait(a.prog + (s - 1), min(k + 2, totp), a.prog + m + 1);
      int r0, ln;
      double beta, scale;
      if (k == 0) {
        // ---- type 1: reflector from column s, rows s+1 .. s+ln ---------------------------------------------
        ln = min(b, m - 1 - s);
        r0 = s + 1;
        const double x = l < ln ? __ldcg(a.Bd + (1 + l) + (size_t)s * kLdb) : 0.0;
        const double xn2 = wsum((l >= 1 && l < ln) ? x * x : 0.0);
        make_house(__shfl_sync(0xffffffffu, x, 0), xn2, beta, tau, scale);
        vs[l] = l == 0 ? 1.0 : ((l < ln && tau != 0.0) ? x * scale : 0.0);
        if (l < ln) a.Bd[(1 + l) + (size_t)s * kLdb] = l == 0 ? beta : 0.0;
        __syncwarp();
      } else {
        // ---- type 2: block below the previous diagonal block: right-apply H_prev, new reflector, left-apply ----
        const int st = s + 1 + (k - 1) * b;
        const int lp = min(b, m - st);
        const int j1 = st + lp;
        ln = min(b, m - j1);
        r0 = j1;
        const double taup = tau;
        vps[l] = vs[l];
        __syncwarp();
#pragma unroll 8
        for (int q = 0; q < 32; ++q)
          Bs[l][q] = (l < ln && q < lp) ? __ldcg(a.Bd + (lp + l - q) + (size_t)(st + q) * kLdb) : 0.0;
        double u = 0.0;
#pragma unroll 8
        for (int q = 0; q < 32; ++q) u = fma(Bs[l][q], vps[q], u);
        u *= taup;
#pragma unroll 8
        for (int q = 0; q < 32; ++q) Bs[l][q] -= u * vps[q];
        const double x = Bs[l][0];
        const double xn2 = wsum((l >= 1 && l < ln) ? x * x : 0.0);
        make_house(__shfl_sync(0xffffffffu, x, 0), xn2, beta, tau, scale);
        vs[l] = l == 0 ? 1.0 : ((l < ln && tau != 0.0) ? x * scale : 0.0);
        Bs[l][0] = l == 0 ? beta : 0.0;
        __syncwarp();
        if (l >= 1 && l < lp) {                 // lane = column
          double y = 0.0;
#pragma unroll 8
          for (int i = 0; i < 32; ++i) y = fma(vs[i], Bs[i][l], y);
          y *= tau;
#pragma unroll 8
          for (int i = 0; i < 32; ++i) Bs[i][l] -= vs[i] * y;
        }
        __syncwarp();
        if (l < ln)
#pragma unroll 8
          for (int q = 0; q < 32; ++q)
            if (q < lp) a.Bd[(lp + l - q) + (size_t)(st + q) * kLdb] = Bs[l][q];
      }
      // ---- type 3: two-sided update of the diagonal block r0 .. r0+ln-1 ---------------------------------------
#pragma unroll 8
      for (int q = 0; q < 32; ++q)
        Ds[l][q] = (q <= l && l < ln) ? __ldcg(a.Bd + (l - q) + (size_t)(r0 + q) * kLdb) : 0.0;
      __syncwarp();
#pragma unroll 8
      for (int q = 0; q < 32; ++q)
        if (q > l) Ds[l][q] = Ds[q][l];
      __syncwarp();
      const double vl = vs[l];
      double p = 0.0;
#pragma unroll 8
      for (int q = 0; q < 32; ++q) p = fma(Ds[l][q], vs[q], p);
      p *= tau;
      const double w = fma(-0.5 * tau * wsum(p * vl), vl, p);
      ws[l] = w;
      __syncwarp();
      if (l < ln)
#pragma unroll 8
        for (int q = 0; q < 32; ++q)
          if (q <= l) a.Bd[(l - q) + (size_t)(r0 + q) * kLdb] = Ds[l][q] - vl * ws[q] - w * vs[q];
      // ---- right-hand sides: z[r0 .. r0+ln-1, :] <- H z ------------------------------------------------------------
      if (a.L > 0) {
        for (int cc = 0; cc < a.L; ++cc) zs[l][cc] = l < ln ? __ldcg(a.z + (size_t)(r0 + l) + (size_t)cc * m) : 0.0;
        __syncwarp();
        if (l < a.L) {
          double tt = 0.0;
#pragma unroll 8
          for (int i = 0; i < 32; ++i) tt = fma(vs[i], zs[i][l], tt);
          ts[l] = tau * tt;
        }
        __syncwarp();
        if (l < ln)
          for (int cc = 0; cc < a.L; ++cc) a.z[(size_t)(r0 + l) + (size_t)cc * m] = zs[l][cc] - vl * ts[cc];
      }
      chase_post(a.prog + s, k + 1);
    }
  }
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace

// Reduces the symmetric m x m matrix at A (column-major, leading dimension ld >= m, LOWER triangle valid on entry, both
// triangles destroyed) to tridiagonal form: d (m) and e (m - 1) on the device; the L right-hand sides z (m x L, ld = m)
// become Q'z.  Synchronises st.
void sym_band_tridiag(mb_ctx* ctx, double* A, int ld, int m, double* z, int L, double* d, double* e, cudaStream_t st) {
  MB_REQUIRE(m >= 3 && ld >= m && ld % 2 == 0, "two-stage tridiagonalisation: bad matrix shape");
  MB_REQUIRE(L >= 0 && L <= 32, "at most 32 right-hand sides per tridiagonalisation");
  Arena& ar = ctx->arena;
  static thread_local bool attr = false;
  if (!attr) {
    MB_CUDA(cudaFuncSetAttribute(k_sbr_qr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kQrSmem));
    MB_CUDA(cudaFuncSetAttribute(k_sbr_r2k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kR2kSmem));
    attr = true;
  }
  const int rmax = m - kBw;
  if (rmax >= 2) {
    const int nt = ceil_div(m, 32);
    MB_LAUNCH(ctx, "k_sbr_mirror", st) k_sbr_mirror<<<dim3(nt, nt), 256, 0, st>>>(A, ld, m);
    const int rpadmax = ceil_div(rmax, 128) * 128;
    const int gmax = ceil_div(rmax, kQrRows);
    int npanel = 0;
    for (int j0 = 0; m - j0 - kBw >= 2; j0 += kBw) ++npanel;
    double* V = ar.take_n<double>((size_t)rpadmax * 32);
    double* W = ar.take_n<double>((size_t)rpadmax * 32);
    double* Z0 = ar.take_n<double>((size_t)rpadmax * 32);
    double* Zp = ar.take_n<double>((size_t)16 * rpadmax * 32);
    double* T = ar.take_n<double>(1024);
    double* Gp = ar.take_n<double>((size_t)(rpadmax / 128) * 2048);
    double* slots = ar.take_n<double>((size_t)2 * (gmax + 1) * 32);
    unsigned* bars = ar.take_n<unsigned>((size_t)64 * npanel);
    MB_CUDA(cudaMemsetAsync(bars, 0, sizeof(unsigned) * 64 * npanel, st));
    int pk = 0;
    for (int j0 = 0; m - j0 - kBw >= 2; j0 += kBw, ++pk) {
      const int r = m - j0 - kBw;
      const int ldv = ceil_div(r, 128) * 128;
      const int rblocks = ldv / 128;
      QrArgs qa{A, ld, m, j0, V, ldv, T, slots, bars + (size_t)64 * pk};
      void* params[] = {&qa};
      MB_LAUNCH(ctx, "k_sbr_qr", st)
        MB_CUDA(cudaLaunchCooperativeKernel((const void*)k_sbr_qr, dim3(ceil_div(r, kQrRows)), dim3(kQrRows), params, kQrSmem, st));
      double* A22 = A + (size_t)(j0 + kBw) * ((size_t)ld + 1);
      int nsplit = std::max(1, std::min(16, ceil_div(3 * ctx->sm_count, rblocks)));
      const int chunk = ceil_div(ceil_div(r, nsplit), 16) * 16;
      nsplit = ceil_div(r, chunk);
      MB_LAUNCH(ctx, "k_sbr_av", st) k_sbr_av<<<dim3(rblocks, nsplit), 128, 0, st>>>(A22, ld, r, V, ldv, Zp, chunk);
      MB_LAUNCH(ctx, "k_sbr_vtz", st)
        k_sbr_vtz<<<rblocks, 256, 0, st>>>(V, ldv, r, Zp, nsplit, Z0, z, m, j0 + kBw, L, Gp);
      MB_LAUNCH(ctx, "k_sbr_w", st) k_sbr_w<<<rblocks, 128, 0, st>>>(V, ldv, r, Z0, T, Gp, rblocks, W, z, m, j0 + kBw, L);
      MB_LAUNCH(ctx, "k_sbr_r2k", st)
        k_sbr_r2k<<<dim3(rblocks, rblocks), 256, kR2kSmem, st>>>(A22, ld, r, V, W, ldv);
    }
    MB_CUDA(cudaGetLastError());
  }
  // ---- stage 2 ------------------------------------------------------------------------------------------------------
  const int ncolb = m + kLdb;
  double* Bd = ar.take_n<double>((size_t)kLdb * ncolb);
  int* prog = ar.take_n<int>((size_t)m + 2);
  MB_CUDA(cudaMemsetAsync(prog, 0, sizeof(int) * ((size_t)m + 2), st));
  MB_LAUNCH(ctx, "k_sbr_band", st) k_sbr_band<<<ceil_div(ncolb, 4), 256, 0, st>>>(A, ld, m, Bd, ncolb);
  if (ctx->sbr_debug) {
    ctx->dbg_band.resize((size_t)kLdb * m);
    MB_CUDA(cudaMemcpyAsync(ctx->dbg_band.data(), Bd, sizeof(double) * kLdb * m, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
  }
  ChaseArgs ca{Bd, m, z, L, prog};
  const int G = std::max(1, std::min(ctx->sm_count, ceil_div(m, 2 * kBw) + 2));
  MB_LAUNCH(ctx, "k_sbr_chase", st) k_sbr_chase<<<G, 32, 0, st>>>(ca);
  MB_LAUNCH(ctx, "k_sbr_diag", st) k_sbr_diag<<<ceil_div(m, 256), 256, 0, st>>>(Bd, m, d, e);
  MB_CUDA(cudaGetLastError());
  int chase_err = 0;
  MB_CUDA(cudaMemcpyAsync(&chase_err, prog + m + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
  if (chase_err) throw Error(MB_E_NUMERIC, "two-stage tridiagonalisation: a bulge-chasing sweep waited too long for its predecessor");
}

}  // namespace mb
