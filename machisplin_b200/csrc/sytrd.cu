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
//   P2  tiles  y = A v on 64 x 64 lower tiles: each tile is read once and used for both y_I += A_IJ v_J and
//              y_J += A_IJ' v_I (second pass from shared memory); partial results go to fixed slots (no atomics:
//              the summation order, hence T, is deterministic); a few CTAs also reduce the panel dots W'v, V'v, v'z
//   P31 rows   w = tau (y - V W'v - W V'v) - 1/2 tau (p'v) v ; z^ <- H_j z^ ; then the next column
//              x = A[j+2:, j+1] - V W[j+1,:]' - W V[j+1,:]' (deferred rank-2k update of the panel), d_{j+1}, |x|^2
// separated by a software grid barrier (two per column; arrivals on a counter, release through a flag on its own
// cache line); every 32 columns the trailing matrix
// takes the rank-64 update A -= V W' + W V' (lower tiles).  The same phase functions are also exposed as separate
// kernels (mb_set_param "sytrd_mode" = 2): kernel boundaries replace the grid barrier.
#include "common.cuh"
#include "internal.h"

#include <algorithm>
#include <cmath>

namespace mb {

constexpr int kTs = 64;            // tile edge
constexpr int kSeg = 4;            // tiles per strip segment
constexpr int kNbMax = 32;         // panel width
constexpr int kSW = 128;           // doubles per row chunk in the P1 partial block: [0] |x|^2, [1..32] z dots, [33+q] W'x, [65+q] V'x
constexpr int kSyThreads = 256;
constexpr int kChunk = 32;         // rows per ownership chunk

struct SytrdArgs {
  double* A; int ld; int m;
  double* V; double* W;            // m x kNbMax, column-major, ld = m
  double* x;                       // m
  double* Pb;                      // nt x m partial products (slot a, rows of block b)
  double* S;                       // nchunk x kSW  P1 partials, one block per row chunk
  double* S2;                      // G             P2 partials (v'Av)
  double* fin;                     // [0..3] tau, scale, beta, alpha | [4 + r] v'z_r | [36 + q] W'v | [68 + q] V'v
  double* z; int L;                // m x L right-hand sides, transformed in place
  double* d; double* e;            // diagonal (m), off-diagonal (m - 1)
  unsigned* bar;                   // [0] arrival counter, [32] release flag (separate 128-byte lines)
  int nt;                          // tiles per dimension
  unsigned long long* prof;        // [8] nanoseconds per phase seen by CTA 0 (P1, sync, P2, sync, P3, sync, update, sync)
};

// ---- grid barrier: arrivals on a counter, the last arriver publishes the generation on a separate line that the
// others poll, so polling never delays an arrival ------------------------------------------------------------
__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned& gen) {
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

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// fixed-order sum of n doubles spaced `stride` apart, by one warp (all lanes return the sum)
__device__ __forceinline__ double warp_strided_sum(const double* p, int n, int stride, int lane) {
  double acc = 0.0;
  for (int base = 0; base < n; base += 256) {       // 8 independent loads per lane in flight
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * 32 + lane;
      v[u] = i < n ? __ldcg(p + (size_t)i * stride) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u];
  }
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

// ---- P1 (first column of a panel; later columns are formed by P31): true column j (rows >= j), |x|^2 per chunk
__device__ void sytrd_p1(const SytrdArgs& a, int j, double* smem) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = a.m, G = gridDim.x;
  const int first_chunk = j / kChunk;
  const int nchunk = (m + kChunk - 1) / kChunk;
  // chunk c is owned by CTA c % G; jj = 0: no deferred update to apply
  int c = first_chunk + (((int)blockIdx.x - first_chunk) % G + G) % G;
  for (; c < nchunk; c += G) {
    if (warp != 0) continue;
    const int i = c * kChunk + lane;
    const bool act = i >= j && i < m;
    double xi = 0.0;
    if (act) {
      xi = __ldcg(&a.A[(size_t)j * a.ld + i]);
      if (i == j) a.d[j] = xi;
      else a.x[i] = xi;
    }
    const double n2 = warp_sum((act && i > j + 1) ? xi * xi : 0.0);
    if (lane == 0) a.S[(size_t)c * kSW] = n2;
  }
  (void)smem;
}

// segments of block row r (relative to the first active block): r / kSeg + 1; rows 4a .. 4a+3 before them hold
// 2a(a+1) + b(a+1) segments in total
__device__ __forceinline__ void segment_of(int s, int& Ii, int& seg) {
  int aq = (int)((sqrt(1.0 + 2.0 * s) - 1.0) * 0.5);
  while (2 * aq * (aq + 1) > s) --aq;
  while (2 * (aq + 1) * (aq + 2) <= s) ++aq;
  const int rem = s - 2 * aq * (aq + 1);
  Ii = kSeg * aq + rem / (aq + 1);
  seg = rem % (aq + 1);
}
__device__ __forceinline__ int segment_count(int na) {
  const int aq = na / kSeg, bq = na % kSeg;
  return (aq + 1) * (2 * aq + bq);
}

// ---- P2: y = A v on the lower tiles ------------------------------------------------------------------------
// Work unit = segment of <= 4 consecutive tiles of one block row I.  No shared memory and no barrier per tile:
// lane = (rgrp, cgrp), thread = 8 rows x 2 columns of the tile in registers (128-bit loads, 512 contiguous bytes
// per column across the 8 row groups; the next tile is prefetched while the current one is used).  Each element
// feeds both products: t_r += A[r,c] v_c accumulates over the whole segment in registers, u_c = sum_r A[r,c] v_r is
// reduced over the 8 row groups by three shuffles.  v is zero outside (j, m), so stale rows / columns <= j and the
// zero padding of the work matrix need no predicate; only the diagonal tile masks its upper triangle.
// Slots: Pb[a][rows of block b] holds, for a <= b, the row product of the segment of block row b that STARTS at
// tile column a (zero for the other columns of the segment) and, for a > b, the transposed product of tile (a, b).
__device__ void sytrd_p2(const SytrdArgs& a, int j, int jj, double* smem) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = a.m, G = gridDim.x;
  double* s_red = smem;                               // [8][64]: row products per warp
  double* s_ud = s_red + 8 * kTs;                    // 64: transposed product of the diagonal tile
  double* s_vI = s_ud + kTs;                          // 64
  const int first_chunk = j / kChunk;
  const int nchunk = (m + kChunk - 1) / kChunk;
  const int Ib0 = (j + 1) / kTs;
  const int na = a.nt - Ib0;
  const int nseg = segment_count(na);
  const int rgrp = lane & 7, cgrp = lane >> 3;
  const int r0 = rgrp * 8;                            // rows r0 .. r0 + 7 of the tile
  const int c0 = warp * 8 + cgrp * 2;                 // columns c0, c0 + 1 of the tile
  // Everything that does not depend on the Householder scalars is requested first: the first tile of this CTA's
  // first segment and the raw x entries it needs travel while |x|^2 is being reduced.
  double pre[16], pvj[2], vI[8];
  auto load_tile = [&](int I, int J) {
    const double* base = a.A + (size_t)(J * kTs + c0) * a.ld + I * kTs + r0;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc)
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const double2 v2 = __ldcg(reinterpret_cast<const double2*>(base + (size_t)cc * a.ld) + h);
        pre[cc * 8 + 2 * h] = v2.x;
        pre[cc * 8 + 2 * h + 1] = v2.y;
      }
  };
  auto xraw = [&](int i) -> double { return (i > j + 1 && i < m) ? __ldcg(&a.x[i]) : 0.0; };
  int sidx = blockIdx.x, Ii = 0, seg = 0;
  if (sidx < nseg) {
    segment_of(sidx, Ii, seg);
    load_tile(Ib0 + Ii, Ib0 + seg * kSeg);
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) vI[rr] = xraw((Ib0 + Ii) * kTs + r0 + rr);
    pvj[0] = xraw((Ib0 + seg * kSeg) * kTs + c0);
    pvj[1] = xraw((Ib0 + seg * kSeg) * kTs + c0 + 1);
  }
  const double alpha = __ldcg(&a.x[j + 1]);
  // every warp reduces the |x|^2 partials itself (same order, same result): no barrier on the critical path
  const double xn2 = warp_strided_sum(a.S + (size_t)first_chunk * kSW, nchunk - first_chunk, kSW, lane);
  const Scal sc = householder_scalars(alpha, xn2);
  if (blockIdx.x == 0 && tid == 0) {
    a.fin[0] = sc.tau; a.fin[1] = sc.scale; a.fin[2] = sc.beta; a.fin[3] = sc.alpha;
    a.e[j] = sc.beta;
  }
  // v_i from the raw x entry (0 outside (j + 1, m)); the pivot row carries v = 1, rows <= j and >= m stay 0
  auto vfix = [&](double raw, int i) -> double { return i == j + 1 ? 1.0 : raw * sc.scale; };
  // dots for P31, one per CTA (from the back of the grid: those CTAs own the fewest segments), straight from x:
  // v'z_r = z[j+1] + scale sum_{i>j+1} x_i z_i ;  g1 = W'v, g2 = V'v likewise (the pivot row enters with v = 1)
  for (int k = G - 1 - (int)blockIdx.x; k < a.L + 2 * jj; k += G) {
    int slot; const double* colp;
    if (k < a.L) { slot = 4 + k; colp = a.z + (size_t)k * m; }
    else if (k < a.L + jj) { slot = 36 + (k - a.L); colp = a.W + (size_t)(k - a.L) * m; }
    else { slot = 68 + (k - a.L - jj); colp = a.V + (size_t)(k - a.L - jj) * m; }
    double acc = 0.0;
    for (int i0 = j + 2 + tid; i0 < m; i0 += 4 * kSyThreads) {
      double xv[4], cv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kSyThreads;
        xv[u] = i < m ? __ldcg(&a.x[i]) : 0.0;
        cv[u] = i < m ? __ldcg(&colp[i]) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc = fma(xv[u], cv[u], acc);
    }
    acc = warp_sum(acc);
    __syncthreads();
    if (lane == 0) s_red[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
#pragma unroll
      for (int q = 0; q < kSyThreads / 32; ++q) t += s_red[q];
      a.fin[slot] = __ldcg(&colp[j + 1]) + sc.scale * t;
    }
  }
  double vav = 0.0;
  bool first = true;
  for (; sidx < nseg; sidx += G) {
    if (!first) segment_of(sidx, Ii, seg);
    const int I = Ib0 + Ii;
    const int Ja = Ib0 + seg * kSeg, Jb = min(I + 1, Ja + kSeg);     // tile columns [Ja, Jb)
    const int R0 = I * kTs;
    if (!first) {
      load_tile(I, Ja);
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) vI[rr] = xraw(R0 + r0 + rr);
      pvj[0] = xraw(Ja * kTs + c0);
      pvj[1] = xraw(Ja * kTs + c0 + 1);
    }
    first = false;
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) vI[rr] = vfix(vI[rr], R0 + r0 + rr);
    double tsum[8];                                    // row products of the whole segment (off-diagonal + diagonal tile)
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) tsum[rr] = 0.0;
    for (int J = Ja; J < Jb; ++J) {
      const double vj0 = vfix(pvj[0], J * kTs + c0), vj1 = vfix(pvj[1], J * kTs + c0 + 1);
      double u0a = 0.0, u0b = 0.0, u1a = 0.0, u1b = 0.0;
      if (I != J) {
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          tsum[rr] = fma(pre[rr], vj0, fma(pre[8 + rr], vj1, tsum[rr]));
          if (rr & 1) { u0b = fma(pre[rr], vI[rr], u0b); u1b = fma(pre[8 + rr], vI[rr], u1b); }
          else { u0a = fma(pre[rr], vI[rr], u0a); u1a = fma(pre[8 + rr], vI[rr], u1a); }
        }
      } else {
        // diagonal tile: lower triangle only; the diagonal itself belongs to the row product.  Its row product
        // enters v'Av once, not twice like the off-diagonal tiles: remember q = v_I' t_diag and take it off below.
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int r = r0 + rr;
          const double a0 = r >= c0 ? pre[rr] : 0.0, a1 = r >= c0 + 1 ? pre[8 + rr] : 0.0;
          const double td = fma(a0, vj0, a1 * vj1);
          tsum[rr] += td;
          vav -= vI[rr] * td;
          u0a = fma(r > c0 ? a0 : 0.0, vI[rr], u0a);
          u1a = fma(r > c0 + 1 ? a1 : 0.0, vI[rr], u1a);
        }
      }
      if (J + 1 < Jb) {                               // in flight during the reduction below and the other warps' work
        load_tile(I, J + 1);
        pvj[0] = xraw((J + 1) * kTs + c0);
        pvj[1] = xraw((J + 1) * kTs + c0 + 1);
      }
      double u0 = u0a + u0b, u1 = u1a + u1b;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        u0 += __shfl_xor_sync(0xffffffffu, u0, o);
        u1 += __shfl_xor_sync(0xffffffffu, u1, o);
      }
      if (rgrp == 0) {
        if (I == J) { s_ud[c0] = u0; s_ud[c0 + 1] = u1; }
        else {
          if (J * kTs + c0 < m) a.Pb[(size_t)I * m + J * kTs + c0] = u0;
          if (J * kTs + c0 + 1 < m) a.Pb[(size_t)I * m + J * kTs + c0 + 1] = u1;
        }
      }
    }
    // segment end: the row products of the 4 column groups of a warp, then of the 8 warps (fixed order)
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      tsum[rr] += __shfl_xor_sync(0xffffffffu, tsum[rr], 8);
      tsum[rr] += __shfl_xor_sync(0xffffffffu, tsum[rr], 16);
    }
    if (cgrp == 0) {
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) s_red[(size_t)warp * kTs + r0 + rr] = tsum[rr];
    }
    if (warp == 0 && cgrp == 1) {
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) s_vI[r0 + rr] = vI[rr];
    }
    __syncthreads();
    if (tid < kTs) {
      const bool has_diag = Jb == I + 1;
      double t = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += s_red[(size_t)q * kTs + tid];
      const double ud = has_diag ? s_ud[tid] : 0.0;
      if (R0 + tid < m) {
        a.Pb[(size_t)Ja * m + R0 + tid] = t + ud;
        for (int J = Ja + 1; J < Jb; ++J) a.Pb[(size_t)J * m + R0 + tid] = 0.0;
      }
      // v'Av: off-diagonal tiles count twice (tile and its mirror), the diagonal tile once (t_diag + u_diag):
      // 2 t + u_diag here, minus the v' t_diag every thread took off above
      vav += s_vI[tid] * (2.0 * t + ud);
    }
    __syncthreads();                        // s_red / s_ud / s_vI are free for the next segment
  }
  // v'Av partial of this CTA: fixed-order reduction over all threads
  vav = warp_sum(vav);
  __syncthreads();
  if (lane == 0) s_red[warp] = vav;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < kSyThreads / 32; ++q) t += s_red[q];
    a.S2[blockIdx.x] = t;
  }
}

// ---- P31: w column of the panel and z update for column j, then (with_next) the true column j + 1 ---------
// Every CTA first forms row j + 1 of the new panel column redundantly (w_{j+1}; v_{j+1} = 1), which is all the
// next column needs from the other CTAs, so P3 of column j and P1 of column j + 1 share one barrier interval.
__device__ void sytrd_p31(const SytrdArgs& a, int j, int jj, bool with_next, double* smem) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = a.m, G = gridDim.x;
  double* s_red = smem;                 // [3][8][32]: y, corr, next-column update slices
  const int Ib0 = (j + 1) / kTs;
  const int first_chunk = (j + 1) / kChunk;
  const int nchunk = (m + kChunk - 1) / kChunk;
  int c = first_chunk + (((int)blockIdx.x - first_chunk) % G + G) % G;
  if (c >= nchunk) return;              // no rows of this CTA left: nothing to do in this phase
  // per-warp constants of the panel: g1 = W'v, g2 = V'v and row j + 1 of W, V for q = warp + 8 u
  double g1q[kNbMax / 8], g2q[kNbMax / 8], rwq[kNbMax / 8], rvq[kNbMax / 8];
#pragma unroll
  for (int u = 0; u < kNbMax / 8; ++u) {
    const int q = warp + 8 * u;
    const bool on = q < jj;
    g1q[u] = on ? __ldcg(&a.fin[36 + q]) : 0.0;
    g2q[u] = on ? __ldcg(&a.fin[68 + q]) : 0.0;
    rwq[u] = on ? __ldcg(&a.W[(size_t)q * m + j + 1]) : 0.0;
    rvq[u] = on ? __ldcg(&a.V[(size_t)q * m + j + 1]) : 0.0;
  }
  // warp 0 owns the scalars: tau, scale, p'v and w_{j+1}; lane q carries column q of the panel
  double tau = 0.0, scale = 0.0, ptv = 0.0, wpiv = 0.0;
  bool have_scalars = false;
  for (; c < nchunk; c += G) {
    const int i = c * kChunk + lane;
    const bool act = i > j && i < m;
    // ---- requests first --------------------------------------------------------------------------------
    double pb[12];
#pragma unroll
    for (int u = 0; u < 12; ++u) {
      const int t = Ib0 + warp + 8 * u;
      pb[u] = (act && t < a.nt) ? __ldcg(&a.Pb[(size_t)t * m + i]) : 0.0;
    }
    double vq[kNbMax / 8], wq[kNbMax / 8];
#pragma unroll
    for (int u = 0; u < kNbMax / 8; ++u) {
      const int q = warp + 8 * u;
      vq[u] = (act && q < jj) ? __ldcg(&a.V[(size_t)q * m + i]) : 0.0;
      wq[u] = (act && q < jj) ? __ldcg(&a.W[(size_t)q * m + i]) : 0.0;
    }
    double xi = 0.0, anext = 0.0;
    if (warp == 0 && act) {
      xi = __ldcg(&a.x[i]);
      if (with_next) anext = __ldcg(&a.A[(size_t)(j + 1) * a.ld + i]);
    }
    if (warp == 0 && !have_scalars) {
      const double s2 = warp_strided_sum(a.S2, G, 1, lane);
      const double ypiv = warp_strided_sum(a.Pb + (size_t)Ib0 * m + j + 1, a.nt - Ib0, m, lane);
      tau = __ldcg(&a.fin[0]);
      scale = __ldcg(&a.fin[1]);
      const bool on = lane < jj;
      const double g1 = on ? __ldcg(&a.fin[36 + lane]) : 0.0, g2 = on ? __ldcg(&a.fin[68 + lane]) : 0.0;
      const double rw = on ? __ldcg(&a.W[(size_t)lane * m + j + 1]) : 0.0, rv = on ? __ldcg(&a.V[(size_t)lane * m + j + 1]) : 0.0;
      const double g12 = warp_sum(g1 * g2);
      const double cpiv = warp_sum(rv * g1 + rw * g2);
      ptv = tau * (s2 - 2.0 * g12);
      wpiv = tau * (ypiv - cpiv) - 0.5 * tau * ptv;      // w_{j+1}
      have_scalars = true;
    }
    // ---- slices ----------------------------------------------------------------------------------------
    double ys = 0.0, cs = 0.0, us = 0.0;
#pragma unroll
    for (int u = 0; u < 12; u += 4) ys += (pb[u] + pb[u + 1]) + (pb[u + 2] + pb[u + 3]);
    if (Ib0 + warp + 96 < a.nt)                                 // more than 96 tile rows (m > 6144): remaining slots
      for (int t = Ib0 + warp + 96; t < a.nt; t += 8) ys += act ? __ldcg(&a.Pb[(size_t)t * m + i]) : 0.0;
#pragma unroll
    for (int u = 0; u < kNbMax / 8; ++u) {
      cs += vq[u] * g1q[u] + wq[u] * g2q[u];
      us += vq[u] * rwq[u] + wq[u] * rvq[u];                    // deferred update of the next column
    }
    __syncthreads();                                            // previous chunk done with s_red
    s_red[warp * 32 + lane] = ys;
    s_red[256 + warp * 32 + lane] = cs;
    s_red[512 + warp * 32 + lane] = us;
    __syncthreads();
    if (warp == 0) {
      double xn = 0.0;
      if (act) {
        double y = 0.0, corr = 0.0, upd = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { y += s_red[k * 32 + lane]; corr += s_red[256 + k * 32 + lane]; upd += s_red[512 + k * 32 + lane]; }
        const double vi = (i == j + 1) ? 1.0 : xi * scale;
        const double wi = tau * (y - corr) - 0.5 * tau * ptv * vi;
        a.V[(size_t)jj * m + i] = vi;
        a.W[(size_t)jj * m + i] = wi;
        for (int r = 0; r < a.L; ++r) {
          double* zp = &a.z[(size_t)r * m + i];
          *zp = __ldcg(zp) - tau * __ldcg(&a.fin[4 + r]) * vi;
        }
        if (with_next) {
          xn = anext - upd - (vi * wpiv + wi);
          if (i == j + 1) a.d[j + 1] = xn;
          else a.x[i] = xn;
        }
      }
      if (with_next) {
        const double n2 = warp_sum((act && i > j + 2) ? xn * xn : 0.0);
        if (lane == 0) a.S[(size_t)c * kSW] = n2;
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

constexpr size_t kSytrdSmemP2 = sizeof(double) * (8 * kTs + 2 * kTs + 8);
constexpr size_t kSytrdSmemUpd = sizeof(double) * 4 * kNbMax * kTs;
constexpr size_t kSytrdSmem = kSytrdSmemP2 > kSytrdSmemUpd ? kSytrdSmemP2 : kSytrdSmemUpd;

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(kSyThreads, 2) k_sytrd(SytrdArgs a) {
  extern __shared__ __align__(16) double sy_smem[];
  unsigned target = 0;
  const int m = a.m;
  const bool prof = a.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  unsigned long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t0 = prof ? gtimer() : 0;
#define MB_PROF(slot) if (prof) { const unsigned long long t1 = gtimer(); acc[slot] += t1 - t0; t0 = t1; }
  for (int j0 = 0; j0 < m - 2; j0 += kNbMax) {
    const int nbp = min(kNbMax, m - 2 - j0);
    sytrd_p1(a, j0, sy_smem);
    MB_PROF(0)
    grid_sync(a.bar, target);
    MB_PROF(1)
    for (int jj = 0; jj < nbp; ++jj) {
      const int j = j0 + jj;
      sytrd_p2(a, j, jj, sy_smem);
      MB_PROF(2)
      grid_sync(a.bar, target);
      MB_PROF(3)
      sytrd_p31(a, j, jj, jj + 1 < nbp, sy_smem);
      MB_PROF(4)
      grid_sync(a.bar, target);
      MB_PROF(5)
    }
    sytrd_update(a, j0 + nbp, nbp, sy_smem);
    MB_PROF(6)
    grid_sync(a.bar, target);
    MB_PROF(7)
  }
#undef MB_PROF
  sytrd_tail(a);
  if (prof)
    for (int k = 0; k < 8; ++k) a.prof[k] = acc[k];
}

// the same phases as separate kernels (kernel boundaries instead of the grid barrier)
__global__ void __launch_bounds__(kSyThreads) k_sytrd_p1(SytrdArgs a, int j) {
  extern __shared__ __align__(16) double sy_smem[];
  sytrd_p1(a, j, sy_smem);
}
__global__ void __launch_bounds__(kSyThreads, 2) k_sytrd_p2(SytrdArgs a, int j, int jj) {
  extern __shared__ __align__(16) double sy_smem[];
  sytrd_p2(a, j, jj, sy_smem);
}
__global__ void __launch_bounds__(kSyThreads) k_sytrd_p31(SytrdArgs a, int j, int jj, int with_next) {
  extern __shared__ __align__(16) double sy_smem[];
  sytrd_p31(a, j, jj, with_next != 0, sy_smem);
}
__global__ void __launch_bounds__(kSyThreads) k_sytrd_update(SytrdArgs a, int jn, int nbp, int last) {
  extern __shared__ __align__(16) double sy_smem[];
  sytrd_update(a, jn, nbp, sy_smem);
  (void)last;
}
__global__ void k_sytrd_tail(SytrdArgs a) { sytrd_tail(a); }

// ---------------------------------------------------------------------------------------------
// Eigenvalues of the symmetric tridiagonal (d, e) by multisection on the Sturm count.  Eight threads share one
// eigenvalue index k (ascending): each evaluates the count at one of 8 interior points of the current bracket,
// the bracket shrinks 9-fold per round (3.17 bits instead of 1), start = Gershgorin hull.  The count uses the
// three-term recurrence p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2} (one FP64 FMA on the critical path instead of
// a division), rescaled by an exact power of two every 4 steps (off the dependent chain, see the loop); #{eigenvalues < x} = #{i : sign p_i != sign
// p_{i-1}} (sign bits; an exact zero and its successor contribute one change).  Rounds stop when the bracket is below abstol (a fixed fraction
// of ulp(|T|)) or 2 ulp of its own magnitude - deterministic, independent of the launch geometry.
// ---------------------------------------------------------------------------------------------
// Geometry: every CTA keeps its own copy of (d, e^2) in shared memory (16 m bytes: two CTAs per SM at 5 000 knots), and the run
// time of a CTA is the length of the dependent recurrence, not the number of eigenvalues it owns - so the grid is ONE wave:
// eigenvalues per CTA = ceil(m / (2 SMs)) rounded up to whole warps (round 2: 313 CTAs of 16 eigenvalues on 296 slots ran as two
// waves, 5.8 ms).  Signs are read from the exponent word with integer instructions: the FP64 pipe (2 issue cycles per warp
// instruction) keeps only the three operations of the recurrence.
__global__ void __launch_bounds__(1024) k_tri_eig(const double* __restrict__ d, const double* __restrict__ e2,
                                                  int m, int per_cta, double lo, double hi, double abstol,
                                                  double* __restrict__ eta) {
  extern __shared__ double s_tri[];           // d | e2
  double* s_d = s_tri;
  double* s_e2 = s_tri + m;
  for (int i = threadIdx.x; i < m; i += blockDim.x) { s_d[i] = d[i]; s_e2[i] = i < m - 1 ? e2[i] : 0.0; }
  __syncthreads();
  const int sec = threadIdx.x & 7;
  int k = blockIdx.x * per_cta + (threadIdx.x >> 3);   // k-th smallest eigenvalue
  const bool live = k < m && (int)(threadIdx.x >> 3) < per_cta;
  if (!live) k = m - 1;                       // keep the whole warp in the shuffles
  double a = lo, b = hi;
  constexpr int kPts = 1;                       // section points per thread: 8 kPts per eigenvalue, bracket / (8 kPts + 1) per round
  constexpr double kInv = 1.0 / (8 * kPts + 1);
  for (int it = 0; it < 40; ++it) {
    const double w = b - a;
    if (!(w > abstol && w > 4.4e-16 * fmax(fabs(a), fabs(b)))) break;   // uniform within the group of 8
    // The kPts recurrences of a thread are independent chains of one DFMA per step.  Two per thread (17-fold shrink per round, 16
    // rounds instead of 20 at 5 000 knots) measured SLOWER: 5.2 against 4.0 ms (profiles/r3z_bench_tpsonly.json against
    // r3f_bench_tpsonly.json) - with the rescaling off the chain the kernel is bound by the FP64 pipe as much as by its latency.
    double x[kPts], p0[kPts], p1[kPts], sc[kPts];
    int hprev[kPts];                              // exponent word of the previous term (its sign bit is what matters)
    int cnt[kPts];
#pragma unroll
    for (int u = 0; u < kPts; ++u) {
      x[u] = a + w * ((kPts * sec + u + 1) * kInv);
      p0[u] = 1.0; p1[u] = s_d[0] - x[u];
      const bool neg1 = p1[u] < 0.0 || p1[u] == 0.0;   // p_0 = 1 > 0: a zero first term counts as a sign change
      cnt[u] = neg1;
      hprev[u] = neg1 ? (int)0x80000000 : 0;
      sc[u] = 1.0;
    }
    // A sign change is the XOR of the sign bits of consecutive terms: three integer instructions per step (the kernel is bound by
    // its instruction count - 58 % issue-active with two warps per scheduler, profiles/r3x_ncu_full_tri_eig.md - and the
    // "p2 < 0 || (p2 == 0 && !neg1)" logic was half of it).  An exact zero takes the sign of its sign bit; the term after it is
    // -e^2 p_{i-1}, so the pair (zero, successor) still contributes exactly one change - the count is the one of the old rule.
#define MB_STURM_COUNT(u, p2)                                                            \
    {                                                                                    \
      const int h2 = __double2hiint(p2);                                                 \
      cnt[u] += (int)((unsigned)(h2 ^ hprev[u]) >> 31);                                  \
      hprev[u] = h2;                                                                     \
    }
    // The power of two that brings a pair back to unit magnitude is taken from the values at the end of a group of four steps
    // and folded into the COEFFICIENTS of the first step of the next group (the recurrence is linear in (p0, p1) and powers of
    // two are exact, so every sign is the one of the unscaled sequence): the only dependent operation per step is one DFMA -
    // with "p0 *= sc; p1 *= sc" the exponent extraction and a DMUL sat on the chain of a kernel that is nothing but that chain.
    int i = 1;
    for (; i + 3 < m; i += 4) {
      {
        const double di = s_d[i], ei = s_e2[i - 1];
#pragma unroll
        for (int u = 0; u < kPts; ++u) {
          const double dx = (di - x[u]) * sc[u], ee = ei * sc[u];
          const double p2 = fma(dx, p1[u], -ee * p0[u]);
          p0[u] = p1[u] * sc[u]; p1[u] = p2;
          MB_STURM_COUNT(u, p2)
        }
      }
#pragma unroll
      for (int v = 1; v < 4; ++v) {
        const double di = s_d[i + v], ei = s_e2[i + v - 1];
#pragma unroll
        for (int u = 0; u < kPts; ++u) {
          const double p2 = fma(di - x[u], p1[u], -ei * p0[u]);
          p0[u] = p1[u]; p1[u] = p2;
          MB_STURM_COUNT(u, p2)
        }
      }
#pragma unroll
      for (int u = 0; u < kPts; ++u) {
        int ex = (__double2hiint(p1[u]) >> 20) & 0x7ff;
        if (ex == 0) ex = (__double2hiint(p0[u]) >> 20) & 0x7ff;
        sc[u] = (ex > 0 && ex < 2046) ? __hiloint2double((2046 - ex) << 20, 0) : 1.0;   // 2^(1023 - ex)
      }
    }
    for (; i < m; ++i) {
      const double di = s_d[i], ei = s_e2[i - 1];
#pragma unroll
      for (int u = 0; u < kPts; ++u) {
        const double dx = (di - x[u]) * sc[u], ee = ei * sc[u];
        const double p2 = fma(dx, p1[u], -ee * p0[u]);
        p0[u] = p1[u] * sc[u]; p1[u] = p2;
        sc[u] = 1.0;
        MB_STURM_COUNT(u, p2)
      }
    }
#undef MB_STURM_COUNT
    // cnt = #eigenvalues < x, non-decreasing in the point index j = kPts sec + u: the new bracket is
    // [largest x with cnt <= k, smallest x with cnt > k]
    const unsigned grp = 0xffu << ((threadIdx.x & 31) & ~7);
    int first = 8 * kPts;                         // first section point whose count exceeds k
#pragma unroll
    for (int u = 0; u < kPts; ++u) {
      const unsigned above = __ballot_sync(0xffffffffu, cnt[u] > k) & grp;
      if (above) first = min(first, kPts * ((__ffs(above) - 1) & 7) + u);
    }
    const double na = first == 0 ? a : a + w * (first * kInv);
    const double nb = first == 8 * kPts ? b : a + w * ((first + 1) * kInv);
    a = na; b = nb;
  }
  if (live && sec == 0) eta[k] = 0.5 * (a + b);
}

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
// Tridiagonalises the symmetric m x m matrix at A (column-major, ld = 64 ceil(m / 64) rows AND columns allocated, zero
// outside m x m: the tile loads carry no bounds predicate; the LOWER triangle is read and destroyed),
// transforms the L right-hand sides z (m x L, ld = m) to Q'z in place and returns the eigenvalues of T
// (ascending) together with T itself.  All outputs are host vectors; synchronises st.
void sym_tridiag_eig(mb_ctx* ctx, double* A, int ld, int m, double* z_dev, int L, std::vector<double>& diag,
                     std::vector<double>& off, std::vector<double>& eta, cudaStream_t st) {
  MB_REQUIRE(m >= 3, "tridiagonalisation needs m >= 3");
  MB_REQUIRE(ld >= (m + kTs - 1) / kTs * kTs && ld % 2 == 0,
             "tridiagonalisation: the work matrix must be zero-padded to a multiple of 64 rows and columns");
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
  a.bar = ar.take_n<unsigned>(64);
  a.prof = ctx->timing ? ar.take_n<unsigned long long>(8) : nullptr;
  static thread_local bool attr = false;
  if (!attr) {
    MB_CUDA(cudaFuncSetAttribute(k_sytrd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSytrdSmem));
    MB_CUDA(cudaFuncSetAttribute(k_sytrd_p1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSytrdSmem));
    MB_CUDA(cudaFuncSetAttribute(k_sytrd_p2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSytrdSmem));
    MB_CUDA(cudaFuncSetAttribute(k_sytrd_p31, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSytrdSmem));
    MB_CUDA(cudaFuncSetAttribute(k_sytrd_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSytrdSmem));
    attr = true;
  }
  int occ = 0;
  MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sytrd, kSyThreads, kSytrdSmem));
  if (occ < 1) throw Error(MB_E_UNSUPPORTED, "k_sytrd does not fit on an SM");
  // the grid barrier needs every CTA resident: at most two per SM, and never more than the device can hold
  int G = std::max(1, std::min(occ, ctx->sytrd_ctas_per_sm > 0 ? ctx->sytrd_ctas_per_sm : 2)) * ctx->sm_count;
  {
    // small matrices (the 1 500-px tiles of mltps carry a few hundred knots): no more CTAs than there is work for -
    // row chunks, strip segments, panel dot products - so that the barrier and the partial reductions stay cheap
    const int nchunk = (m + kChunk - 1) / kChunk;
    const int na = a.nt, nseg = (na / kSeg + 1) * (2 * (na / kSeg) + na % kSeg);
    G = std::min(G, std::max(std::max(nchunk, nseg), 8));
  }
  a.S = ar.take_n<double>((size_t)((m + kChunk - 1) / kChunk) * kSW);
  a.S2 = ar.take_n<double>(G);
  MB_CUDA(cudaMemsetAsync(a.bar, 0, 64 * sizeof(unsigned), st));
  if (ctx->sytrd_mode == 0 || ctx->sytrd_mode == 3) {
    // two-stage reduction (sbr.cu): dense -> band by compact-WY panels, band -> tridiagonal by bulge chasing
    sym_band_tridiag(ctx, A, ld, m, z_dev, L, a.d, a.e, st);
  } else if (ctx->sytrd_mode != 2) {
    void* params[] = {&a};
    MB_LAUNCH(ctx, "k_sytrd", st)
      MB_CUDA(cudaLaunchCooperativeKernel((const void*)k_sytrd, dim3(G), dim3(kSyThreads), params, kSytrdSmem, st));
  } else {
    for (int j0 = 0; j0 < m - 2; j0 += kNbMax) {
      const int nbp = std::min(kNbMax, m - 2 - j0);
      k_sytrd_p1<<<G, kSyThreads, kSytrdSmem, st>>>(a, j0);
      for (int jj = 0; jj < nbp; ++jj) {
        k_sytrd_p2<<<G, kSyThreads, kSytrdSmem, st>>>(a, j0 + jj, jj);
        k_sytrd_p31<<<G, kSyThreads, kSytrdSmem, st>>>(a, j0 + jj, jj, jj + 1 < nbp ? 1 : 0);
      }
      k_sytrd_update<<<G, kSyThreads, kSytrdSmem, st>>>(a, j0 + nbp, nbp, 0);
    }
    k_sytrd_tail<<<1, 32, 0, st>>>(a);
    ctx->launches += 1;
  }
  MB_CUDA(cudaGetLastError());
  diag.resize(m);
  off.resize(m - 1);
  unsigned long long prof_ns[8] = {0};
  if (a.prof && ctx->sytrd_mode == 1)
    MB_CUDA(cudaMemcpyAsync(prof_ns, a.prof, sizeof(prof_ns), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaMemcpyAsync(diag.data(), a.d, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaMemcpyAsync(off.data(), a.e, sizeof(double) * (m - 1), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
  for (int k = 0; k < 8; ++k) ctx->sytrd_prof_ms[k] = prof_ns[k] * 1e-6;
  // Gershgorin hull
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
  double tnorm = 0.0;
  for (int i = 0; i < m; ++i) tnorm = std::max(tnorm, std::fabs(diag[i]) + (i > 0 ? std::fabs(off[i - 1]) : 0.0) + (i < m - 1 ? std::fabs(off[i]) : 0.0));
  const double abstol = 1e-3 * 2.220446049250313e-16 * tnorm;   // far below what the reduction itself resolves
  // one wave: as many CTAs as fit beside each other (shared memory decides), whole warps of 4 eigenvalues x 8 section points
  const int per_sm = std::max(1, std::min(8, (int)((220u * 1024u) / (smem + 1024))));
  const int slots = std::max(1, per_sm * ctx->sm_count);
  const int per_cta = std::min(128, ((m + slots - 1) / slots + 3) & ~3);
  MB_LAUNCH(ctx, "k_tri_eig", st)
    k_tri_eig<<<(m + per_cta - 1) / per_cta, 8 * per_cta, smem, st>>>(a.d, d_e2, m, per_cta, lo, hi, abstol, d_eta);
  MB_CUDA(cudaGetLastError());
  eta.resize(m);
  MB_CUDA(cudaMemcpyAsync(eta.data(), d_eta, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
}

}  // namespace mb
