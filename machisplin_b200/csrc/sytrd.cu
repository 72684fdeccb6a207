// In-house symmetric tridiagonalisation + tridiagonal eigenvalues for the GCV fit (fields::Tps -> Krig.engine.default
// -> eigen(Q2' K Q2), V73:722, 751).  Replaces the cuSOLVER Dsyevd stand-in of the first slice.
//
// fields needs eigen(M) only through two functionals of lambda - trA(lambda) = 3 + sum eta/(eta + lambda) and
// RSS(lambda) = lambda^2 |(M + lambda I)^-1 z|^2 - so no eigenvector is ever formed: M = Q T Q' (Householder),
// z^ = Q' z is carried along, eta = eig(T) by Sturm bisection, RSS through O(m) tridiagonal solves on the host,
// and the coefficients at the selected lambda come from the tensor-core Cholesky of M + lambda I (tps_fit.cu).
//
// k_sytrd: ONE persistent kernel, blocked right-looking reduction on the LOWER triangle only (the 100 MB lower
// half of a 5 000 x 5 000 float64 matrix stays in the 126 MB L2; the full matrix would not).  Per column j:
//   P1 rows   x = A[j+1:, j] - V W[j,:]' - W V[j,:]'   (deferred rank-2k update of the panel), d_j, partial norms / dots
//   P2 tiles  y = A v on 64 x 64 lower tiles: each tile is read once and used for both y_I += A_IJ v_J and
//             y_J += A_IJ' v_I (second pass from shared memory); partial results go to fixed slots (no atomics:
//             the summation order, hence T, is deterministic)
//   P3 rows   w = tau (y - V W'v - W V'v) - 1/2 tau (p'v) v ; z^ <- H_j z^
// separated by a software grid barrier (release/acquire on one counter); every 32 columns the trailing matrix
// takes the rank-64 update A -= V W' + W V' (lower tiles).  The same phase functions are also exposed as separate
// kernels (mb_set_param "sytrd_mode" = 2): kernel boundaries replace the grid barrier.
#include "common.cuh"
#include "internal.h"

#include <algorithm>
#include <cmath>

namespace mb {

constexpr int kTs = 64;            // tile edge
constexpr int kTsPad = kTs + 1;    // shared-memory column stride (conflict-free transposed pass)
constexpr int kNbMax = 32;         // panel width
constexpr int kSW = 128;           // doubles per CTA in the P1 partial block: [0] |x|^2, [1..32] z dots, [32+q] W'x, [64+q] V'x
constexpr int kSyThreads = 256;
constexpr int kChunk = 32;         // rows per ownership chunk

struct SytrdArgs {
  double* A; int ld; int m;
  double* V; double* W;            // m x kNbMax, column-major, ld = m
  double* x;                       // m
  double* Pb;                      // nt x m partial products
  double* S;                       // G x kSW      P1 partials
  double* S2;                      // G            P2 partials (v'Av)
  double* fin;                     // 4 + 2 kNbMax + 32: tau, scale, beta, alpha | g1 | g2 | z dots
  double* z; int L;                // m x L right-hand sides, transformed in place
  double* d; double* e;            // diagonal (m), off-diagonal (m - 1)
  unsigned* bar;
  int nt;                          // tiles per dimension
};

// ---- grid barrier -------------------------------------------------------------------------------
__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    unsigned seen;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
    } while ((int)(seen - target) < 0);
  }
  __syncthreads();
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// fixed-order sum of n doubles spaced `stride` apart, by one warp (all lanes return the sum)
__device__ __forceinline__ double warp_strided_sum(const double* p, int n, int stride, int lane) {
  double acc = 0.0;
  for (int i = lane; i < n; i += 32) acc += __ldcg(p + (size_t)i * stride);
  return warp_sum(acc);
}

struct Scal { double tau, scale, beta, alpha; };

__device__ __forceinline__ Scal householder_scalars(double alpha, double xnorm2) {
  Scal s;
  s.alpha = alpha;
  if (xnorm2 == 0.0) { s.tau = 0.0; s.scale = 0.0; s.beta = alpha; return s; }
  const double beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
  s.beta = beta;
  s.tau = (beta - alpha) / beta;
  s.scale = 1.0 / (alpha - beta);
  return s;
}

// ---- P1: true column j (rows >= j), partial norms / dots ------------------------------------------
__device__ void sytrd_p1(const SytrdArgs& a, int j, int jj, double* smem) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = a.m;
  double* s_red = smem;                       // [8][32]
  double* s_part = smem + 8 * 32;             // [kSW] accumulated over this CTA's chunks
  for (int i = tid; i < kSW; i += kSyThreads) s_part[i] = 0.0;
  __syncthreads();
  const int first_chunk = j / kChunk;
  const int nchunk = (m + kChunk - 1) / kChunk;
  // chunk c is owned by CTA c % gridDim.x
  int c = first_chunk + (((int)blockIdx.x - first_chunk) % (int)gridDim.x + (int)gridDim.x) % (int)gridDim.x;
  for (; c < nchunk; c += gridDim.x) {
    const int i = c * kChunk + lane;
    const bool act = i >= j && i < m;
    // slice of the deferred update: q = warp, warp + 8, ...
    double acc = 0.0;
    if (act)
      for (int q = warp; q < jj; q += 8)
        acc += __ldcg(&a.V[(size_t)q * m + i]) * __ldcg(&a.W[(size_t)q * m + j]) +
               __ldcg(&a.W[(size_t)q * m + i]) * __ldcg(&a.V[(size_t)q * m + j]);
    s_red[warp * 32 + lane] = acc;
    __syncthreads();
    double xi = 0.0;
    if (warp == 0) {
      double corr = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) corr += s_red[k * 32 + lane];
      if (act) {
        xi = __ldcg(&a.A[(size_t)j * a.ld + i]) - corr;
        if (i == j) a.d[j] = xi;
        else a.x[i] = xi;
      }
      s_red[lane] = (act && i > j) ? xi : 0.0;     // x of this chunk for the other warps
    }
    __syncthreads();
    xi = s_red[lane];
    const bool tail = act && i > j + 1;            // rows below the pivot row j + 1
    // warp 0: |x|^2 ; warps: z dots, W'x, V'x slices (x of the FULL column incl. the pivot row for W'x / V'x)
    if (warp == 0) {
      const double n2 = warp_sum(tail ? xi * xi : 0.0);
      if (lane == 0) s_part[0] += n2;
    }
    for (int r = warp; r < a.L; r += 8) {
      const double t = warp_sum(tail ? xi * __ldcg(&a.z[(size_t)r * m + i]) : 0.0);
      if (lane == 0) s_part[1 + r] += t;
    }
    for (int q = warp; q < jj; q += 8) {
      const double t1 = warp_sum(tail ? xi * __ldcg(&a.W[(size_t)q * m + i]) : 0.0);
      const double t2 = warp_sum(tail ? xi * __ldcg(&a.V[(size_t)q * m + i]) : 0.0);
      if (lane == 0) { s_part[33 + q] += t1; s_part[65 + q] += t2; }
    }
    __syncthreads();
  }
  __syncthreads();
  for (int i = tid; i < kSW; i += kSyThreads) a.S[(size_t)blockIdx.x * kSW + i] = s_part[i];
}

// ---- P2: y = A v on the lower tiles; reduction of the P1 partials by the last CTA -------------------
__device__ void sytrd_p2(const SytrdArgs& a, int j, int jj, double* smem) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = a.m, G = gridDim.x;
  double* s_tile = smem;                        // [64][65]
  double* s_vI = s_tile + kTs * kTsPad;         // 64
  double* s_vJ = s_vI + kTs;                    // 64
  double* s_red = s_vJ + kTs;                   // [4][64]
  double* s_sc = s_red + 4 * kTs;               // scalars: 0 xnorm2, 1 vAv accumulator
  if (warp == 0) {
    const double n2 = warp_strided_sum(a.S, G, kSW, lane);
    if (lane == 0) { s_sc[0] = n2; s_sc[1] = 0.0; }
  }
  __syncthreads();
  const double alpha = __ldcg(&a.x[j + 1]);
  const Scal sc = householder_scalars(alpha, s_sc[0]);
  // the last CTA publishes the scalars and the reduced dots for P3
  if (blockIdx.x == G - 1) {
    if (tid == 0) {
      a.fin[0] = sc.tau; a.fin[1] = sc.scale; a.fin[2] = sc.beta; a.fin[3] = sc.alpha;
      a.e[j] = sc.beta;
    }
    // v'z_r = z[j+1] + scale sum_{i>j+1} x_i z_i ;  g1 = W'v, g2 = V'v likewise (pivot row enters with v = 1)
    for (int k = warp; k < a.L + 2 * jj; k += 8) {
      int col; const double* piv;
      if (k < a.L) { col = 1 + k; piv = &a.z[(size_t)k * m + j + 1]; }
      else if (k < a.L + jj) { col = 33 + (k - a.L); piv = &a.W[(size_t)(k - a.L) * m + j + 1]; }
      else { col = 65 + (k - a.L - jj); piv = &a.V[(size_t)(k - a.L - jj) * m + j + 1]; }
      const double t = warp_strided_sum(a.S + col, G, kSW, lane);
      if (lane == 0) a.fin[4 + (col - 1)] = __ldcg(piv) + sc.scale * t;   // fin[4 + r] z dots, fin[36 + q] g1, fin[68 + q] g2
    }
  }
  const int Ib0 = (j + 1) / kTs;
  const int na = a.nt - Ib0;
  const int ntiles = na * (na + 1) / 2;
  const int r = tid & 63, cg = tid >> 6;
  double vav = 0.0;
  for (int t = blockIdx.x; t < ntiles; t += G) {
    int Ii = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while (Ii * (Ii + 1) / 2 > t) --Ii;
    while ((Ii + 1) * (Ii + 2) / 2 <= t) ++Ii;
    const int Ji = t - Ii * (Ii + 1) / 2;
    const int I = Ib0 + Ii, J = Ib0 + Ji;
    const int R0 = I * kTs, C0 = J * kTs;
    __syncthreads();                                        // previous tile done with shared memory
    if (tid < kTs) {
      const int i = R0 + tid;
      double v = 0.0;
      if (i == j + 1) v = 1.0; else if (i > j + 1 && i < m) v = __ldcg(&a.x[i]) * sc.scale;
      s_vI[tid] = v;
    } else if (tid < 2 * kTs) {
      const int i = C0 + tid - kTs;
      double v = 0.0;
      if (i == j + 1) v = 1.0; else if (i > j + 1 && i < m) v = __ldcg(&a.x[i]) * sc.scale;
      s_vJ[tid - kTs] = v;
    }
    __syncthreads();
    // pass A: rows along lanes (coalesced), 16 columns per thread
    const int rg = R0 + r;
    double tacc = 0.0;
#pragma unroll 4
    for (int cc = 0; cc < 16; ++cc) {
      const int c = cg * 16 + cc;
      const int cgl = C0 + c;
      double v = 0.0;
      if (rg < m && rg > j && cgl > j && (I != J || r >= c)) v = __ldcg(&a.A[(size_t)cgl * a.ld + rg]);
      tacc = fma(v, s_vJ[c], tacc);
      s_tile[c * kTsPad + r] = (I == J && r == c) ? 0.0 : v;   // the diagonal is used once (pass A)
    }
    s_red[cg * kTs + r] = tacc;
    __syncthreads();
    // pass B: columns along lanes, 16 rows per thread (transposed product from shared memory)
    const int c2 = tid & 63, rq = tid >> 6;
    double uacc = 0.0;
#pragma unroll 4
    for (int rr = 0; rr < 16; ++rr) uacc = fma(s_tile[c2 * kTsPad + rq * 16 + rr], s_vI[rq * 16 + rr], uacc);
    double tr = 0.0;
    if (tid < kTs) tr = s_red[r] + s_red[kTs + r] + s_red[2 * kTs + r] + s_red[3 * kTs + r];
    __syncthreads();
    s_red[rq * kTs + c2] = uacc;
    __syncthreads();
    if (tid < kTs) {
      const double uc = s_red[tid] + s_red[kTs + tid] + s_red[2 * kTs + tid] + s_red[3 * kTs + tid];
      if (I == J) {
        if (R0 + tid < m) a.Pb[(size_t)I * m + R0 + tid] = tr + uc;
        vav += s_vI[tid] * (tr + uc);
      } else {
        if (R0 + tid < m) a.Pb[(size_t)J * m + R0 + tid] = tr;
        if (C0 + tid < m) a.Pb[(size_t)I * m + C0 + tid] = uc;
        vav += 2.0 * s_vI[tid] * tr;
      }
    }
  }
  __syncthreads();
  // v'Av partial of this CTA (threads 0..63 hold the pieces): fixed-order reduction
  if (tid < kTs) s_red[tid] = vav;
  __syncthreads();
  if (warp == 0) {
    const double t = warp_sum(s_red[lane] + s_red[32 + lane]);
    if (lane == 0) a.S2[blockIdx.x] = t;
  }
}

// ---- P3: w column of the panel, z update ------------------------------------------------------------
__device__ void sytrd_p3(const SytrdArgs& a, int j, int jj, double* smem) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = a.m, G = gridDim.x;
  double* s_red = smem;                 // [8][32] y slices, [8][32] corr slices
  double* s_g = smem + 2 * 8 * 32;      // fin copy (4 + 96)
  double* s_sc = s_g + 128;
  for (int i = tid; i < 4 + 32 + 2 * kNbMax; i += kSyThreads) s_g[i] = __ldcg(&a.fin[i]);
  if (warp == 0) {
    const double t = warp_strided_sum(a.S2, G, 1, lane);
    if (lane == 0) s_sc[0] = t;
  }
  __syncthreads();
  const double tau = s_g[0], scale = s_g[1];
  const double* g1 = s_g + 36;          // W'v
  const double* g2 = s_g + 68;          // V'v
  double g12 = 0.0;
  for (int q = 0; q < jj; ++q) g12 += g1[q] * g2[q];
  const double ptv = tau * (s_sc[0] - 2.0 * g12);
  const int Ib0 = (j + 1) / kTs;
  const int first_chunk = (j + 1) / kChunk;
  const int nchunk = (m + kChunk - 1) / kChunk;
  int c = first_chunk + (((int)blockIdx.x - first_chunk) % G + G) % G;
  for (; c < nchunk; c += G) {
    const int i = c * kChunk + lane;
    const bool act = i > j && i < m;
    double ys = 0.0, cs = 0.0;
    if (act) {
      for (int t = Ib0 + warp; t < a.nt; t += 8) ys += __ldcg(&a.Pb[(size_t)t * m + i]);
      for (int q = warp; q < jj; q += 8)
        cs += __ldcg(&a.V[(size_t)q * m + i]) * g1[q] + __ldcg(&a.W[(size_t)q * m + i]) * g2[q];
    }
    __syncthreads();
    s_red[warp * 32 + lane] = ys;
    s_red[256 + warp * 32 + lane] = cs;
    __syncthreads();
    if (warp == 0 && act) {
      double y = 0.0, corr = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) { y += s_red[k * 32 + lane]; corr += s_red[256 + k * 32 + lane]; }
      const double vi = (i == j + 1) ? 1.0 : __ldcg(&a.x[i]) * scale;
      const double p = tau * (y - corr);
      a.V[(size_t)jj * m + i] = vi;
      a.W[(size_t)jj * m + i] = p - 0.5 * tau * ptv * vi;
      for (int r = 0; r < a.L; ++r) {
        double* zp = &a.z[(size_t)r * m + i];
        *zp = __ldcg(zp) - tau * s_g[4 + r] * vi;
      }
    }
  }
}

// ---- panel end: A -= V W' + W V' on the lower tiles of the trailing matrix (rows / cols >= jn) -----------
__device__ void sytrd_update(const SytrdArgs& a, int jn, int nbp, double* smem) {
  const int tid = threadIdx.x;
  const int m = a.m, G = gridDim.x;
  double* s_vr = smem;                       // [nbp][64] V rows of the tile's row block
  double* s_wr = s_vr + kNbMax * kTs;
  double* s_vc = s_wr + kNbMax * kTs;        // ... of the column block
  double* s_wc = s_vc + kNbMax * kTs;
  const int Ib0 = jn / kTs;
  const int na = a.nt - Ib0;
  const int ntiles = na * (na + 1) / 2;
  const int tr = (tid & 15) * 4, tc = (tid >> 4) * 4;     // 4 x 4 outputs per thread
  for (int t = blockIdx.x; t < ntiles; t += G) {
    int Ii = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while (Ii * (Ii + 1) / 2 > t) --Ii;
    while ((Ii + 1) * (Ii + 2) / 2 <= t) ++Ii;
    const int Ji = t - Ii * (Ii + 1) / 2;
    const int R0 = (Ib0 + Ii) * kTs, C0 = (Ib0 + Ji) * kTs;
    __syncthreads();
    for (int idx = tid; idx < nbp * kTs; idx += kSyThreads) {
      const int q = idx / kTs, k = idx % kTs;
      const int ir = R0 + k, ic = C0 + k;
      const bool okr = ir >= jn && ir < m, okc = ic >= jn && ic < m;
      s_vr[idx] = okr ? __ldcg(&a.V[(size_t)q * m + ir]) : 0.0;
      s_wr[idx] = okr ? __ldcg(&a.W[(size_t)q * m + ir]) : 0.0;
      s_vc[idx] = okc ? __ldcg(&a.V[(size_t)q * m + ic]) : 0.0;
      s_wc[idx] = okc ? __ldcg(&a.W[(size_t)q * m + ic]) : 0.0;
    }
    __syncthreads();
    double acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
    for (int q = 0; q < nbp; ++q) {
      double vr[4], wr[4], vc[4], wc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        vr[u] = s_vr[q * kTs + tr + u]; wr[u] = s_wr[q * kTs + tr + u];
        vc[u] = s_vc[q * kTs + tc + u]; wc[u] = s_wc[q * kTs + tc + u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fma(vr[u], wc[v], fma(wr[u], vc[v], acc[u][v]));
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int ic = C0 + tc + v;
      if (ic < jn || ic >= m) continue;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ir = R0 + tr + u;
        if (ir >= ic && ir < m) {
          double* p = &a.A[(size_t)ic * a.ld + ir];
          *p = __ldcg(p) - acc[u][v];
        }
      }
    }
  }
}

__device__ void sytrd_tail(const SytrdArgs& a) {
  const int m = a.m;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.d[m - 2] = __ldcg(&a.A[(size_t)(m - 2) * a.ld + m - 2]);
    a.d[m - 1] = __ldcg(&a.A[(size_t)(m - 1) * a.ld + m - 1]);
    a.e[m - 2] = __ldcg(&a.A[(size_t)(m - 2) * a.ld + m - 1]);
  }
}

constexpr size_t kSytrdSmemP2 = sizeof(double) * (kTs * kTsPad + 2 * kTs + 4 * kTs + 8);
constexpr size_t kSytrdSmemUpd = sizeof(double) * 4 * kNbMax * kTs;
constexpr size_t kSytrdSmem = kSytrdSmemP2 > kSytrdSmemUpd ? kSytrdSmemP2 : kSytrdSmemUpd;

__global__ void __launch_bounds__(kSyThreads) k_sytrd(SytrdArgs a) {
  extern __shared__ __align__(16) double sy_smem[];
  unsigned target = 0;
  const int m = a.m;
  for (int j0 = 0; j0 < m - 2; j0 += kNbMax) {
    const int nbp = min(kNbMax, m - 2 - j0);
    for (int jj = 0; jj < nbp; ++jj) {
      const int j = j0 + jj;
      sytrd_p1(a, j, jj, sy_smem);
      grid_sync(a.bar, target);
      sytrd_p2(a, j, jj, sy_smem);
      grid_sync(a.bar, target);
      sytrd_p3(a, j, jj, sy_smem);
      grid_sync(a.bar, target);
    }
    sytrd_update(a, j0 + nbp, nbp, sy_smem);
    grid_sync(a.bar, target);
  }
  sytrd_tail(a);
}

// the same phases as separate kernels (kernel boundaries instead of the grid barrier)
__global__ void __launch_bounds__(kSyThreads) k_sytrd_p1(SytrdArgs a, int j, int jj) {
  extern __shared__ __align__(16) double sy_smem[];
  sytrd_p1(a, j, jj, sy_smem);
}
__global__ void __launch_bounds__(kSyThreads) k_sytrd_p2(SytrdArgs a, int j, int jj) {
  extern __shared__ __align__(16) double sy_smem[];
  sytrd_p2(a, j, jj, sy_smem);
}
__global__ void __launch_bounds__(kSyThreads) k_sytrd_p3(SytrdArgs a, int j, int jj) {
  extern __shared__ __align__(16) double sy_smem[];
  sytrd_p3(a, j, jj, sy_smem);
}
__global__ void __launch_bounds__(kSyThreads) k_sytrd_update(SytrdArgs a, int jn, int nbp, int last) {
  extern __shared__ __align__(16) double sy_smem[];
  sytrd_update(a, jn, nbp, sy_smem);
  (void)last;
}
__global__ void k_sytrd_tail(SytrdArgs a) { sytrd_tail(a); }

// ---------------------------------------------------------------------------------------------
// Eigenvalues of the symmetric tridiagonal (d, e) by bisection on the Sturm count.  One thread per eigenvalue
// index (ascending); the start interval is the Gershgorin hull.  The count uses the three-term recurrence
// p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2} (two dependent FP64 operations per step instead of a division),
// rescaled by an exact power of two every 4 steps; #{eigenvalues < x} = #{i : sign p_i != sign p_{i-1}}, an
// exact zero counting as a change.  Every eigenvalue is resolved until the midpoint stops moving.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) k_tri_eig(const double* __restrict__ d, const double* __restrict__ e2, int m,
                                                double lo, double hi, double* __restrict__ eta) {
  extern __shared__ double s_tri[];           // d | e2
  double* s_d = s_tri;
  double* s_e2 = s_tri + m;
  for (int i = threadIdx.x; i < m; i += blockDim.x) { s_d[i] = d[i]; s_e2[i] = i < m - 1 ? e2[i] : 0.0; }
  __syncthreads();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;     // k-th smallest eigenvalue
  if (k >= m) return;
  double a = lo, b = hi;
  for (int it = 0; it < 100; ++it) {
    const double mid = 0.5 * (a + b);
    if (!(mid > a && mid < b)) break;
    double p0 = 1.0, p1 = s_d[0] - mid;
    bool neg1 = p1 < 0.0 || p1 == 0.0;        // p_0 = 1 > 0: a zero counts as a sign change
    int cnt = neg1;
    for (int i = 1; i < m; ++i) {
      const double p2 = fma(s_d[i] - mid, p1, -s_e2[i - 1] * p0);
      const bool neg2 = p2 < 0.0 || (p2 == 0.0 && !neg1);
      cnt += neg2 != neg1;
      p0 = p1; p1 = p2; neg1 = neg2;
      if ((i & 3) == 0) {
        int ex = (__double2hiint(p1) >> 20) & 0x7ff;
        if (ex == 0) ex = (__double2hiint(p0) >> 20) & 0x7ff;
        if (ex > 0 && ex < 2046) {
          const double sc = __hiloint2double((2046 - ex) << 20, 0);   // 2^(1023 - ex)
          p0 *= sc; p1 *= sc;
        }
      }
    }
    if (cnt > k) b = mid; else a = mid;      // cnt = #eigenvalues < mid
  }
  eta[k] = 0.5 * (a + b);
}

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
// Tridiagonalises the symmetric m x m matrix at A (column-major, ld; the LOWER triangle is read and destroyed),
// transforms the L right-hand sides z (m x L, ld = m) to Q'z in place and returns the eigenvalues of T
// (ascending) together with T itself.  All outputs are host vectors; synchronises st.
void sym_tridiag_eig(mb_ctx* ctx, double* A, int ld, int m, double* z_dev, int L, std::vector<double>& diag,
                     std::vector<double>& off, std::vector<double>& eta, cudaStream_t st) {
  MB_REQUIRE(m >= 3, "tridiagonalisation needs m >= 3");
  MB_REQUIRE(L >= 0 && L <= 32, "at most 32 right-hand sides per tridiagonalisation");
  Arena& ar = ctx->arena;
  SytrdArgs a{};
  a.A = A; a.ld = ld; a.m = m;
  a.nt = (m + kTs - 1) / kTs;
  a.V = ar.take_n<double>((size_t)m * kNbMax);
  a.W = ar.take_n<double>((size_t)m * kNbMax);
  a.x = ar.take_n<double>(m);
  a.Pb = ar.take_n<double>((size_t)a.nt * m);
  a.z = z_dev; a.L = L;
  a.d = ar.take_n<double>(m);
  a.e = ar.take_n<double>(m);
  a.fin = ar.take_n<double>(128);
  a.bar = ar.take_n<unsigned>(4);
  static thread_local bool attr = false;
  if (!attr) {
    MB_CUDA(cudaFuncSetAttribute(k_sytrd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSytrdSmem));
    MB_CUDA(cudaFuncSetAttribute(k_sytrd_p1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSytrdSmem));
    MB_CUDA(cudaFuncSetAttribute(k_sytrd_p2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSytrdSmem));
    MB_CUDA(cudaFuncSetAttribute(k_sytrd_p3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSytrdSmem));
    MB_CUDA(cudaFuncSetAttribute(k_sytrd_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSytrdSmem));
    attr = true;
  }
  int occ = 0;
  MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sytrd, kSyThreads, kSytrdSmem));
  if (occ < 1) throw Error(MB_E_UNSUPPORTED, "k_sytrd does not fit on an SM");
  // the grid barrier needs every CTA resident: at most two per SM, and never more than the device can hold
  const int G = std::max(1, std::min(occ, ctx->sytrd_ctas_per_sm > 0 ? ctx->sytrd_ctas_per_sm : 2)) * ctx->sm_count;
  a.S = ar.take_n<double>((size_t)G * kSW);
  a.S2 = ar.take_n<double>(G);
  MB_CUDA(cudaMemsetAsync(a.bar, 0, 4 * sizeof(unsigned), st));
  if (ctx->sytrd_mode != 2) {
    void* params[] = {&a};
    MB_LAUNCH(ctx, "k_sytrd", st)
      MB_CUDA(cudaLaunchCooperativeKernel((const void*)k_sytrd, dim3(G), dim3(kSyThreads), params, kSytrdSmem, st));
  } else {
    for (int j0 = 0; j0 < m - 2; j0 += kNbMax) {
      const int nbp = std::min(kNbMax, m - 2 - j0);
      for (int jj = 0; jj < nbp; ++jj) {
        k_sytrd_p1<<<G, kSyThreads, kSytrdSmem, st>>>(a, j0 + jj, jj);
        k_sytrd_p2<<<G, kSyThreads, kSytrdSmem, st>>>(a, j0 + jj, jj);
        k_sytrd_p3<<<G, kSyThreads, kSytrdSmem, st>>>(a, j0 + jj, jj);
      }
      k_sytrd_update<<<G, kSyThreads, kSytrdSmem, st>>>(a, j0 + nbp, nbp, 0);
    }
    k_sytrd_tail<<<1, 32, 0, st>>>(a);
    ctx->launches += 1;
  }
  MB_CUDA(cudaGetLastError());
  diag.resize(m);
  off.resize(m - 1);
  MB_CUDA(cudaMemcpyAsync(diag.data(), a.d, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaMemcpyAsync(off.data(), a.e, sizeof(double) * (m - 1), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
  // Gershgorin hull and the dlaebz pivot guard
  double lo = 1e300, hi = -1e300, emax2 = 0.0;
  std::vector<double> e2(m, 0.0);
  for (int i = 0; i < m; ++i) {
    const double l = i > 0 ? std::fabs(off[i - 1]) : 0.0, r = i < m - 1 ? std::fabs(off[i]) : 0.0;
    lo = std::min(lo, diag[i] - l - r);
    hi = std::max(hi, diag[i] + l + r);
    if (i < m - 1) { e2[i] = off[i] * off[i]; emax2 = std::max(emax2, e2[i]); }
    if (!std::isfinite(diag[i]) || !std::isfinite(l)) throw Error(MB_E_NUMERIC, "tridiagonalisation produced a non-finite entry");
  }
  const double span = std::max(hi - lo, 1e-300);
  lo -= 1e-12 * span + 2.2e-308;
  hi += 1e-12 * span + 2.2e-308;
  double* d_e2 = ar.upload(e2.data(), e2.size(), st);
  double* d_eta = ar.take_n<double>(m);
  const size_t smem = sizeof(double) * 2 * (size_t)m;
  if (smem > 200 * 1024) throw Error(MB_E_UNSUPPORTED, "tridiagonal eigenvalue kernel: matrix too large for shared memory");
  static thread_local bool attr2 = false;
  if (!attr2) {
    MB_CUDA(cudaFuncSetAttribute(k_tri_eig, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr2 = true;
  }
  MB_LAUNCH(ctx, "k_tri_eig", st) k_tri_eig<<<(m + 63) / 64, 64, smem, st>>>(a.d, d_e2, m, lo, hi, d_eta);
  MB_CUDA(cudaGetLastError());
  eta.resize(m);
  MB_CUDA(cudaMemcpyAsync(eta.data(), d_eta, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
}

}  // namespace mb
