// TPS surface evaluation on a raster window - replaces terra::interpolate(rast(template), Tps)
// -> predict.Krig -> Fortran multrb at V73:726 and V73:753.
//
//   f(cell) = d0 + d1 sx + d2 sy + (1/16pi) sum_i c_i d2_i log(d2_i),   d2_i = max(|s - s_i|^2, 1e-20)
//
// Two device paths:
//   tps_eval_direct  K4: O(cells x knots) float64 pair sum with libm log - the parity kernel.
//   tps_eval_fast    K5: the roofline kernel.  The far field of every leaf box (32 x bh cells) is a
//                    P x P Chebyshev expansion built by a kernel-independent treecode (P2L from the
//                    interaction list of each quadtree level + exact polynomial re-expansion
//                    parent -> child), the near field (knots in the 3 x 3 box neighbourhood) is summed
//                    directly.  Everything is float64: the coefficients c cancel by up to 1e8
//                    (T'c = 0), so float32 pair terms are not accurate enough.
#include "common.cuh"
#include "async_copy.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>

namespace mb {

// =========================================================================================
// float64 table-driven log: x = 2^e m, m in [1,2); m_k = centre of the 1/256-wide mantissa bin;
// log x = e ln2 - log(inv_k) + log1p(m inv_k - 1), |m inv_k - 1| <= 2^-9, series to t^5
// (truncation 2^-54/6).  ~10 FP64 ops instead of ~45 for libm log.
// =========================================================================================
void init_logtab(mb_ctx* ctx) {
  std::vector<double2> tab(256);
  for (int k = 0; k < 256; ++k) {
    long double mk = 1.0L + (k + 0.5L) / 256.0L;
    double inv = (double)(1.0L / mk);
    tab[k].x = inv;
    tab[k].y = (double)(-logl((long double)inv));
  }
  ctx->logtab.upload(tab.data(), tab.size(), ctx->stream);
  MB_CUDA(cudaStreamSynchronize(ctx->stream));
}

__device__ __forceinline__ double4 ldg4(const double4* p) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(p));
  const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ double tlog(double x, const double2* __restrict__ tab) {
  int hi = __double2hiint(x);
  int lo = __double2loint(x);
  int e = (hi >> 20) - 1023;
  int idx = (hi >> 12) & 0xFF;
  double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);
  double2 t2 = tab[idx];
  double t = fma(m, t2.x, -1.0);
  double p = fma(t, 0.2, -0.25);
  p = fma(t, p, 0.33333333333333333);
  p = fma(t, p, -0.5);
  p = fma(t * t, p, t);
  double ed = __hiloint2double(0x43300000, e ^ 0x80000000) - 4503601774854144.0;  // (double)e
  return fma(ed, 0.69314718055994531, t2.y + p);
}

// =========================================================================================
// K4: direct evaluation
// =========================================================================================
constexpr int kDirectTile = 256;

__global__ void __launch_bounds__(256) k_tps_eval_direct(
    const double* __restrict__ ksx, const double* __restrict__ ksy, const double* __restrict__ kc, int np,
    double d0, double d1, double d2c, GridAffine a, mb_window w, double* __restrict__ out, int64_t stride) {
  __shared__ double s_x[kDirectTile], s_y[kDirectTile], s_c[kDirectTile];
  const int col = w.c0 + blockIdx.x * 32 + (threadIdx.x & 31);
  const int row = w.r0 + blockIdx.y * 8 + (threadIdx.x >> 5);
  // terra xFromCol / yFromRow, then Krig's transformx - same operation order as the oracle
  const double x = a.xmin + (col + 0.5) * a.rx;
  const double y = a.ymax - (row + 0.5) * a.ry;
  const double sx = (x - a.cx) / a.scx;
  const double sy = (y - a.cy) / a.scy;
  double acc = 0.0;
  for (int base = 0; base < np; base += kDirectTile) {
    const int n = min(kDirectTile, np - base);
    __syncthreads();
    if (threadIdx.x < n) {
      s_x[threadIdx.x] = ksx[base + threadIdx.x];
      s_y[threadIdx.x] = ksy[base + threadIdx.x];
      s_c[threadIdx.x] = kc[base + threadIdx.x];
    }
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
      const double dx = sx - s_x[i];
      const double dy = sy - s_y[i];
      double r2 = dx * dx + dy * dy;
      r2 = fmax(r2, kD2Clamp);
      acc += s_c[i] * (0.5 * log(r2) * r2);
    }
  }
  if (col < w.c1 && row < w.r1)
    out[(int64_t)(row - w.r0) * stride + (col - w.c0)] = d0 + d1 * sx + d2c * sy + kRbfConst * acc;
}

void tps_eval_direct(mb_ctx* ctx, const mb_spline* s, const mb_grid& g, const mb_window& w, double* out,
                     int64_t stride, cudaStream_t st) {
  GridAffine a = make_affine(g, *s);
  dim3 grid((w.c1 - w.c0 + 31) / 32, (w.r1 - w.r0 + 7) / 8);
  MB_LAUNCH(ctx, "k_tps_eval_direct", st) k_tps_eval_direct<<<grid, 256, 0, st>>>(s->d_sx.p, s->d_sy.p, s->d_c.p, s->np, s->d[0], s->d[1], s->d[2], a, w,
                                          out, stride);
  MB_CUDA(cudaGetLastError());
}

__global__ void __launch_bounds__(256) k_tps_points(
    const double* __restrict__ ksx, const double* __restrict__ ksy, const double* __restrict__ kc, int np,
    double d0, double d1, double d2c, double cx, double cy, double scx, double scy,
    const double* __restrict__ px, const double* __restrict__ py, int n, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double sx = (px[i] - cx) / scx;
  const double sy = (py[i] - cy) / scy;
  double acc = 0.0;
  for (int k = 0; k < np; ++k) {
    const double dx = sx - ksx[k];
    const double dy = sy - ksy[k];
    double r2 = fmax(dx * dx + dy * dy, kD2Clamp);
    acc += kc[k] * (0.5 * log(r2) * r2);
  }
  out[i] = d0 + d1 * sx + d2c * sy + kRbfConst * acc;
}

void tps_predict_points_dev(mb_ctx* ctx, const mb_spline* s, const double* x_dev, const double* y_dev, int n,
                            double* out_dev, cudaStream_t st) {
  if (n <= 0) return;
  MB_LAUNCH(ctx, "k_tps_points", st) k_tps_points<<<(n + 255) / 256, 256, 0, st>>>(s->d_sx.p, s->d_sy.p, s->d_c.p, s->np, s->d[0], s->d[1], s->d[2],
                                                s->center[0], s->center[1], s->scale[0], s->scale[1], x_dev, y_dev,
                                                n, out_dev);
  MB_CUDA(cudaGetLastError());
}

void spline_finalize(mb_ctx* ctx, mb_spline* s, double fscale_known) {
  s->ctx = ctx;
  s->device = ctx->device;
  cudaStream_t st = ctx->stream;
  s->d_sx.upload(s->sx, st);
  s->d_sy.upload(s->sy, st);
  s->d_c.upload(s->c, st);
  s->sum_abs_c = 0;
  for (double v : s->c) s->sum_abs_c += std::fabs(v);
  // fscale = max |f(knot)|: sets the accuracy target of the fast evaluator
  if (fscale_known >= 0) { s->fscale = fscale_known; return; }
  ABuf<double> dx(ctx->arena, s->np), dy(ctx->arena, s->np), df(ctx->arena, s->np);
  dx.upload(s->kx, st);
  dy.upload(s->ky, st);
  tps_predict_points_dev(ctx, s, dx.p, dy.p, s->np, df.p, st);
  std::vector<double> f(s->np);
  MB_CUDA(cudaMemcpyAsync(f.data(), df.p, sizeof(double) * s->np, cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
  s->fscale = 0;
  for (double v : f) s->fscale = std::max(s->fscale, std::fabs(v));
}

// =========================================================================================
// K5: fast evaluation
// =========================================================================================
struct ChebTables {
  int P = 0;
  std::vector<double> xn, C, R0, R1;     // nodes, values->coefficients, child re-expansions (row-major P x P)
  DevBuf<double> d_tab;                  // [xn | C | R0 | R1]
};

static void cheb_T(int P, long double x, long double* T) {
  T[0] = 1.0L;
  if (P > 1) T[1] = x;
  for (int k = 2; k < P; ++k) T[k] = 2.0L * x * T[k - 1] - T[k - 2];
}

static ChebTables* get_tables(mb_ctx* ctx, int P) {
  // one table set per (device, P) for the whole process; the tile lanes of tiles_tps call this concurrently
  static std::map<std::pair<int, int>, std::unique_ptr<ChebTables>> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(ctx->device, P);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second.get();
  auto t = std::make_unique<ChebTables>();
  t->P = P;
  const long double pi = 3.14159265358979323846264338327950288L;
  std::vector<long double> xn(P), C(P * P), T(P);
  for (int k = 0; k < P; ++k) xn[k] = cosl((2 * k + 1) * pi / (2.0L * P));
  for (int i = 0; i < P; ++i) {
    cheb_T(P, xn[i], T.data());
    for (int j = 0; j < P; ++j) C[j * P + i] = (j == 0 ? 1.0L : 2.0L) / P * T[j];
  }
  t->xn.resize(P); t->C.resize(P * P); t->R0.resize(P * P); t->R1.resize(P * P);
  for (int k = 0; k < P; ++k) t->xn[k] = (double)xn[k];
  for (int i = 0; i < P * P; ++i) t->C[i] = (double)C[i];
  for (int c = 0; c < 2; ++c) {
    std::vector<long double> R(P * P, 0.0L);
    for (int i = 0; i < P; ++i) {
      cheb_T(P, (xn[i] + (2 * c - 1)) / 2.0L, T.data());   // child node i in parent coordinates
      for (int j = 0; j < P; ++j)
        for (int k = 0; k < P; ++k) R[j * P + k] += C[j * P + i] * T[k];
    }
    auto& dst = c == 0 ? t->R0 : t->R1;
    for (int i = 0; i < P * P; ++i) dst[i] = (double)R[i];
  }
  std::vector<double> all;
  all.insert(all.end(), t->xn.begin(), t->xn.end());
  all.insert(all.end(), t->C.begin(), t->C.end());
  all.insert(all.end(), t->R0.begin(), t->R0.end());
  all.insert(all.end(), t->R1.begin(), t->R1.end());
  t->d_tab.upload(all, ctx->stream);
  MB_CUDA(cudaStreamSynchronize(ctx->stream));
  ChebTables* raw = t.get();
  cache[key] = std::move(t);
  return raw;
}

__host__ __device__ __forceinline__ uint32_t part1by1(uint32_t x) {
  x &= 0x0000ffff;
  x = (x ^ (x << 8)) & 0x00ff00ff;
  x = (x ^ (x << 4)) & 0x0f0f0f0f;
  x = (x ^ (x << 2)) & 0x33333333;
  x = (x ^ (x << 1)) & 0x55555555;
  return x;
}
__host__ __device__ __forceinline__ uint32_t morton(uint32_t i, uint32_t j) { return part1by1(i) | (part1by1(j) << 1); }

struct Lattice {
  // s_x = sxo + hx * U, s_y = syo + hy * V with (U, V) in leaf-box units; hy < 0 (rows run south)
  double sxo, hx, syo, hy;
  int L;            // leaf level: 2^L x 2^L boxes
  int offx, offy;   // leaf box of the window's first cell
  int nbx, nby;     // leaf boxes covering the window
  int bh;           // rows per leaf box (columns are always 32)
};

struct LevelInfo {
  int lvl, sh;      // level, L - level
  int I0, J0, nI, nJ;
  int nsplit;
  size_t coef_off;  // offset (in doubles) of this level's coefficient block
  size_t part_off;  // offset of the partial-sum block
};

// -----------------------------------------------------------------------------------------
// P2L: node values of the far field produced by the interaction list of each active box.
// grid = (box, split); thread = Chebyshev node (jy, ix).  Candidate source boxes are the 6 x 6
// children of the parent's 3 x 3 neighbourhood minus the box's own 3 x 3 neighbourhood.
// -----------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(((P * P + 31) / 32) * 32) k_far_p2l(
    Lattice lat, LevelInfo lv, const double* __restrict__ tab, const int* __restrict__ start,
    const double4* __restrict__ knots, const double2* __restrict__ logtab, double* __restrict__ partial) {
  __shared__ double2 s_log[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_log[i] = logtab[i];
  __syncthreads();
  const int box = blockIdx.x;
  const int split = blockIdx.y;
  const int I = lv.I0 + box % lv.nI;
  const int J = lv.J0 + box / lv.nI;
  const int node = threadIdx.x;
  const bool live = node < P * P;
  const int jy = live ? node / P : 0, ix = live ? node % P : 0;
  const double scale = (double)(1 << lv.sh);
  const double sxn = lat.sxo + lat.hx * scale * (I + 0.5 * (1.0 + tab[ix]));
  const double syn = lat.syo + lat.hy * scale * (J + 0.5 * (1.0 + tab[jy]));
  const int nside = 1 << lv.lvl;
  const int bi0 = ((I >> 1) - 1) * 2, bj0 = ((J >> 1) - 1) * 2;
  double acc = 0.0;
  for (int cand = split; cand < 36; cand += lv.nsplit) {
    const int i = bi0 + cand % 6, j = bj0 + cand / 6;
    if (i < 0 || j < 0 || i >= nside || j >= nside) continue;
    if (abs(i - I) <= 1 && abs(j - J) <= 1) continue;
    const uint32_t z = morton(i, j);
    const int a = start[(size_t)z << (2 * lv.sh)];
    const int b = start[((size_t)z + 1) << (2 * lv.sh)];
    for (int k = a; k < b; ++k) {
      const double4 kn = ldg4(&knots[k]);
      const double dx = sxn - kn.x, dy = syn - kn.y;
      double r2 = fmax(fma(dx, dx, dy * dy), kD2Clamp);
      acc = fma(kn.z * r2, tlog(r2, s_log), acc);
    }
  }
  if (live) partial[lv.part_off + ((size_t)box * lv.nsplit + split) * (P * P) + node] = acc;
}

// -----------------------------------------------------------------------------------------
// values -> Chebyshev coefficients, plus the parent's expansion re-expanded on this child
// (exact for polynomials), plus - at the first level - the affine part d0 + d1 sx + d2 sy.
// -----------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(((P * P + 31) / 32) * 32) k_far_transform(
    Lattice lat, LevelInfo lv, LevelInfo parent, int has_parent, const double* __restrict__ tab,
    const double* __restrict__ partial, double* __restrict__ coef, double d0, double d1, double d2c) {
  __shared__ double s_a[P * P], s_b[P * P], s_p[P * P];
  const double* Cm = tab + P;
  const int box = blockIdx.x;
  const int I = lv.I0 + box % lv.nI;
  const int J = lv.J0 + box / lv.nI;
  const int t = threadIdx.x;
  const bool live = t < P * P;
  const int r = live ? t / P : 0, q = live ? t % P : 0;
  if (live) {
    double v = 0.0;
    for (int s = 0; s < lv.nsplit; ++s) v += partial[lv.part_off + ((size_t)box * lv.nsplit + s) * (P * P) + t];
    s_a[t] = v;                                   // vals[jy][ix]
    if (has_parent) {
      const int pb = ((J >> 1) - parent.J0) * parent.nI + ((I >> 1) - parent.I0);
      s_p[t] = coef[parent.coef_off + (size_t)pb * (P * P) + t];
    }
  }
  __syncthreads();
  double tmp = 0.0, tmp2 = 0.0;
  if (live) {
    // tmp[j=r][ix=q] = sum_jy C[r][jy] vals[jy][q]
    for (int m = 0; m < P; ++m) tmp = fma(Cm[r * P + m], s_a[m * P + q], tmp);
    if (has_parent) {
      const double* Ry = tab + P + P * P + ((J & 1) ? P * P : 0);
      for (int m = 0; m < P; ++m) tmp2 = fma(Ry[r * P + m], s_p[m * P + q], tmp2);
    }
  }
  __syncthreads();
  if (live) { s_b[t] = tmp; s_a[t] = tmp2; }
  __syncthreads();
  if (live) {
    double v = 0.0;
    // coef[j=r][k=q] = sum_ix tmp[r][ix] C[q][ix]
    for (int m = 0; m < P; ++m) v = fma(s_b[r * P + m], Cm[q * P + m], v);
    v *= 0.5 * kRbfConst;                         // E/2 factored out of the pair terms
    if (has_parent) {
      const double* Rx = tab + P + P * P + ((I & 1) ? P * P : 0);
      double v2 = 0.0;
      for (int m = 0; m < P; ++m) v2 = fma(s_a[r * P + m], Rx[q * P + m], v2);
      v += v2;
    } else {
      const double scale = (double)(1 << lv.sh);
      const double sxc = lat.sxo + lat.hx * scale * (I + 0.5);
      const double syc = lat.syo + lat.hy * scale * (J + 0.5);
      if (r == 0 && q == 0) v += d0 + d1 * sxc + d2c * syc;
      if (r == 0 && q == 1) v += d1 * lat.hx * scale * 0.5;
      if (r == 1 && q == 0) v += d2c * lat.hy * scale * 0.5;
    }
    coef[lv.coef_off + (size_t)box * (P * P) + t] = v;
  }
}

// -----------------------------------------------------------------------------------------
// Leaf preparation (one warp per leaf box):
//  (1) accuracy estimate of the mixed-precision leaf path, max over boxes:
//        far : the k >= 1 columns of the expansion are collapsed AND evaluated in float32 -> 2 (P + 2) 2^-24 sum_{(j,k) != (0,0)} |A_jk|
//        near: pair terms in float32 with box-local coordinates      -> 4 2^-24 phi(r_max) sum_{3x3} |c_i E/2|
//      The leaf kernels compare the estimate with their threshold on the device (no host round trip):
//      k_leaf_stream runs when the mixed path is accurate enough, k_leaf_f64 otherwise.
//  (2) the near list of the box - the knots of its 3 x 3 neighbourhood in box-local float32 coordinates
//      (x - ox, y - oy, c E/2 ln2) - as one 256-byte block {count, overflow offset, 15 inline entries}
//      that k_leaf_stream fetches with a single bulk copy; longer lists go to an overflow array.
// -----------------------------------------------------------------------------------------
constexpr int kNearInline = 15;
struct __align__(16) NearBlk {
  int cnt, off, pad0, pad1;
  float4 e[kNearInline];
};
static_assert(sizeof(NearBlk) == 256, "NearBlk must be one 256-byte record");
// Leaf record of one box, fetched by k_leaf_stream with ONE bulk copy:
//   [0, 256)            NearBlk
//   [256, 256 + 8 P)    a0[j] = A[j][0]      float64: the column that carries the cancellation
//   then                af[j][k - 1] = (float) A[j][k], k >= 1, row stride GFS (zero padded): the variation of the far
//                       field inside the box, which the mixed path evaluates in float32 anyway
__host__ __device__ constexpr int leaf_gfs(int P) { return ((P - 1 + 3) / 4) * 4; }
__host__ __device__ constexpr int leaf_rec_bytes(int P) { return 256 + 8 * P + 4 * P * leaf_gfs(P); }

template <int P>
__global__ void __launch_bounds__(256) k_leaf_prep(Lattice lat, LevelInfo leaf, const double* __restrict__ coef,
                                                   const int* __restrict__ start, const double4* __restrict__ knots,
                                                   double phi_max, unsigned long long* __restrict__ est_bits,
                                                   unsigned char* __restrict__ recs, float4* __restrict__ near_over,
                                                   int* __restrict__ over_cursor) {
  constexpr int GFS = leaf_gfs(P);
  const int lane = threadIdx.x & 31;
  const int box = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (box >= leaf.nI * leaf.nJ) return;
  const double* A = coef + leaf.coef_off + (size_t)box * (P * P);
  unsigned char* rec = recs + (size_t)box * leaf_rec_bytes(P);
  double* a0 = reinterpret_cast<double*>(rec + 256);
  float* af = reinterpret_cast<float*>(rec + 256 + 8 * P);
  double v = 0.0;
  for (int i = lane; i < P * GFS; i += 32) {          // (j, k - 1) incl. the zero padding of every row
    const int jr = i / GFS, k = i % GFS + 1;
    const double aij = k < P ? A[jr * P + k] : 0.0;
    af[i] = (float)aij;
    v += fabs(aij);
  }
  for (int jr = lane; jr < P; jr += 32) {
    const double aj0 = A[jr * P];
    a0[jr] = aj0;
    if (jr > 0) v += fabs(aj0);
  }
  const int I = lat.offx + box % leaf.nI, J = lat.offy + box / leaf.nI;
  int ra = 0, rlen = 0;
  double wsum = 0.0;
  if (lane < 9) {
    const int i = I + lane % 3 - 1, j = J + lane / 3 - 1;
    const int nside = 1 << lat.L;
    if (i >= 0 && j >= 0 && i < nside && j < nside) {
      const uint32_t z = morton(i, j);
      ra = start[z];
      rlen = start[z + 1] - ra;
      for (int k = 0; k < rlen; ++k) wsum += fabs(knots[ra + k].w);
    }
  }
  int tot = rlen;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    v += __shfl_xor_sync(0xffffffffu, v, o);
    wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    tot += __shfl_xor_sync(0xffffffffu, tot, o);
  }
  if (lane == 0) {
    const double est = 5.9604644775390625e-08 * (2 * (P + 2) * v + 4.0 * phi_max * wsum);
    atomicMax(est_bits, (unsigned long long)__double_as_longlong(est));   // est >= 0: bit order = value order
  }
  int off = 0;
  if (tot > kNearInline) {
    if (lane == 0) off = atomicAdd(over_cursor, tot);
    off = __shfl_sync(0xffffffffu, off, 0);
  }
  NearBlk* nb = reinterpret_cast<NearBlk*>(rec);
  if (lane == 0) { nb->cnt = tot; nb->off = off; nb->pad0 = 0; nb->pad1 = 0; }
  const double ox = lat.sxo + lat.hx * I, oy = lat.syo + lat.hy * J;
  for (int base = 0; base < tot; base += 32) {
    int f = base + lane, idx = -1;
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int la = __shfl_sync(0xffffffffu, ra, r), ll = __shfl_sync(0xffffffffu, rlen, r);
      if (f >= 0 && f < ll) idx = la + f;
      f -= ll;
    }
    if (idx >= 0) {
      const double4 kn = ldg4(&knots[idx]);
      const float4 ent = make_float4((float)(kn.x - ox), (float)(kn.y - oy), (float)(kn.w * 0.69314718055994531), 0.f);
      if (tot <= kNearInline) nb->e[base + lane] = ent;
      else near_over[off + base + lane] = ent;
    }
  }
}

// -----------------------------------------------------------------------------------------
// async-copy plumbing of k_leaf_stream: mbarrier + cp.async.bulk (the TMA unit's 1-D bulk copy, UBLKCP)
// -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// -----------------------------------------------------------------------------------------
// k_leaf_stream - the grid-evaluation kernel (mixed-precision path).
// Persistent CTAs; CTA c owns leaf boxes c, c + G, c + 2G ... (row-major over the window: concurrently active
// CTAs write neighbouring boxes).  Per box the inputs are three bulk copies into a 3-stage shared-memory ring
// (the P x P expansion block, the 256-byte near block and - with an accumulator - the box's 32 x bh tile of
// it, one 256-byte copy per row), completion tracked by one mbarrier per stage; nothing is staged through
// registers, so the copies of the next boxes are in flight while the current one is computed.
//   collapse(i+1)  G[lr][k] = sum_j T_j(ty_lr) A[j][k]   float64; column 0 kept in float64, columns k >= 1
//                  (the variation of the far field inside the box) rounded to float32
//   rows(i)        lane = column: far = G0[lr] + sum_{k>=1} Gf[lr][k] T_k(tx_lane)  (float32 FMA, T_k in registers)
//                  near = sum over the near list, float32 with box-local coordinates and MUFU.LG2
//                  out  = far + near (+ acc * inv_w : mltps part 5, NaN in acc = NA cell)
// collapse(i+1) and rows(i) share one __syncthreads per box (G is double-buffered).  The large, cancelling part
// of the sum never leaves float64, so the result keeps ~1e-9 relative accuracy while the per-cell work is
// P - 1 FFMA + one conversion + one float64 add: the kernel is an HBM stream (8 B written per cell, + 8 B read
// with an accumulator).
// -----------------------------------------------------------------------------------------
constexpr int kLeafThreads = 256;
constexpr int kLeafStages = 4;
__host__ __device__ inline int leaf_stage_bytes(int P, int bh, bool acc) {
  return (leaf_rec_bytes(P) + (acc ? bh * 256 : 0) + 127) / 128 * 128;
}

// r2 is clamped to >= 1e-20 (a normal float32): the plain MUFU.LG2 without the denormal pre-scaling of __log2f
__device__ __forceinline__ float lg2_fast(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 3 CTAs per SM (<= 85 registers): at 4 (64 registers) the compiler rematerialised the per-box address and shared-memory
// arithmetic in every iteration of a kernel that is issue-bound - fused 0.785 -> 0.825 of the measured HBM bandwidth, TPS only
// 0.547 -> 0.622; 2 CTAs (110 registers) fall back to 0.54 (profiles/r2y_leaf_launch_bounds.txt).
template <int P, bool kAcc>
__global__ void __launch_bounds__(kLeafThreads, 3) k_leaf_stream(
    const __grid_constant__ CUtensorMap acc_map, int acc_tma,
    Lattice lat, mb_window w, const unsigned char* __restrict__ recs, const float4* __restrict__ near_over,
    const unsigned long long* __restrict__ est_bits, double mixed_threshold, AccFuse fz, double* __restrict__ out,
    int64_t stride) {
  static_assert(P % 2 == 0, "P must be even");
  static_assert((kLeafStages & (kLeafStages - 1)) == 0, "stage count must be a power of two");
  if (!(__longlong_as_double((long long)__ldg(est_bits)) <= mixed_threshold)) return;   // k_leaf_f64 runs instead
  constexpr int GFS = leaf_gfs(P);
  constexpr int NQ = GFS / 2;              // float2 column pairs (k = 2q+1, 2q+2) of a G row
  constexpr uint32_t kRecBytes = leaf_rec_bytes(P);
  extern __shared__ __align__(128) unsigned char leaf_smem[];
  __shared__ __align__(8) uint64_t s_full[kLeafStages];
  const int bh = lat.bh;
  const int stage_bytes = leaf_stage_bytes(P, bh, kAcc);
  double* s_G0 = reinterpret_cast<double*>(leaf_smem + kLeafStages * stage_bytes);   // [2][bh]
  float* s_Gf = reinterpret_cast<float*>(s_G0 + 2 * bh);                             // [2][bh][GFS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nboxes = lat.nbx * lat.nby;
  const int G = gridDim.x;
  const int nmine = (nboxes - (int)blockIdx.x + G - 1) / G;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kLeafStages; ++s) mbar_init(&s_full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // box walk without divisions: box b = blockIdx.x + i G has lattice position (bi, bj); every step adds (dI, dJ)
  const int dI = G % lat.nbx, dJ = G / lat.nbx;
  int pbi = blockIdx.x % lat.nbx, pbj = blockIdx.x / lat.nbx;      // producer cursor (warp 0)
  int ibox = 0;                                                     // producer: next box to issue
  // Stage layout: [accumulator tile, bh rows of 256 bytes (kAcc)] [record].  The tile comes first because the destination of a
  // tensor copy must be 128-byte aligned (stage_bytes is a multiple of 128), the record's bulk copy needs 16 bytes only.
  const uint32_t acc_bytes = kAcc ? (uint32_t)bh * 256u : 0u;
  auto issue = [&]() {   // warp 0: copies of this CTA's next box into stage ibox % kLeafStages
    const int s = ibox & (kLeafStages - 1);
    unsigned char* sp = leaf_smem + s * stage_bytes;
    if (lane == 0) {
      mbar_expect_tx(&s_full[s], kRecBytes + acc_bytes);
      bulk_g2s(sp + acc_bytes, recs + (size_t)(blockIdx.x + ibox * G) * kRecBytes, kRecBytes, &s_full[s]);
      // the box's 32 x bh tile of the accumulator: ONE 2-D tensor copy (UTMALDG) issued by one lane.  The 32 row copies it
      // replaces were serialised by the uniform datapath (ELECT + R2UR + UBLKCP per row: ~250 instructions of warp 0 per box,
      // as many as a box's whole row work - every other warp waited for them at the box barrier: 51 % of all stall samples,
      // profiles/r2h_ncu_full_leaf_before_tma.md)
      if (kAcc && acc_tma) ac_tma_load_2d(sp, &acc_map, pbi * 32, pbj * bh, &s_full[s]);
    }
    if (kAcc) {
      if (!acc_tma) {       // no tensor map (driver entry point missing): one 256-byte bulk copy per row
        __syncwarp();
        const double* src = fz.acc + (int64_t)pbj * bh * fz.stride + pbi * 32;
        for (int r = lane; r < bh; r += 32) bulk_g2s(sp + r * 256, src + (int64_t)r * fz.stride, 256, &s_full[s]);
      }
      pbi += dI; pbj += dJ;
      if (pbi >= lat.nbx) { pbi -= lat.nbx; ++pbj; }
    }
    ++ibox;
  };
  if (warp == 0)
    for (int i = 0; i < min(kLeafStages, nmine); ++i) issue();

  // lane = (half, l): half selects one row of a row pair, l the column pair (l, 31 - l): tx_{31-l} = -tx_l, so the
  // even and the odd part of the expansion serve both columns.  (odd, even) terms are packed: one FFMA2 each.
  const int half = lane >> 4, l = lane & 15;
  float2 Tp[NQ];                           // (T_1,T_2), (T_3,T_4), ... ; entries with k >= P are 0 (G is zero padded too)
  {
    float T[2 * NQ + 1];
    const float tx = (2.0f * l + 1.0f) / 32.0f - 1.0f;   // exact in float32
    T[0] = 1.0f;
    T[1] = tx;
#pragma unroll
    for (int k = 2; k <= 2 * NQ; ++k) T[k] = k < P ? 2.0f * tx * T[k - 1] - T[k - 2] : 0.0f;
#pragma unroll
    for (int q = 0; q < NQ; ++q) Tp[q] = make_float2(T[2 * q + 1], T[2 * q + 2]);
  }
  const float cxa = (float)(lat.hx * ((l + 0.5) / 32.0)), cxb = (float)(lat.hx * ((31 - l + 0.5) / 32.0));
  const double hyb = lat.hy / bh, inv_bh = 1.0 / bh;
  const float inv_bh_f = (float)inv_bh;
  const int npair = (bh + 1) >> 1;

  // Collapse y of box i into G buffer i & 1.  Rows lr and bh - 1 - lr have ty of opposite sign: the even and the odd
  // part in ty serve both.  Column 0 in float64 (one thread per row pair), columns k >= 1 in float32, two columns per
  // thread (packed FMA): G[lr][k] = sum_j T_j(ty_lr) A[j][k].
  auto collapse = [&](int i) {
    const unsigned char* rec = leaf_smem + (i & (kLeafStages - 1)) * stage_bytes + acc_bytes;
    const double* a0 = reinterpret_cast<const double*>(rec + 256);
    const float2* af = reinterpret_cast<const float2*>(rec + 256 + 8 * P);     // [P][NQ]
    double* g0buf = s_G0 + (i & 1) * bh;
    float* gfbuf = s_Gf + (i & 1) * bh * GFS;
    const int n32 = npair * NQ;
    for (int o = tid; o < n32 + npair; o += kLeafThreads) {
      if (o < n32) {
        const int lr = o / NQ, q = o % NQ;
        const float ty = fmaf((float)(2 * lr + 1), inv_bh_f, -1.0f);   // exact for bh = 32
        float t0 = 1.0f, t1 = ty;
        float2 ev = af[q], od = af[NQ + q];
        od.x *= ty; od.y *= ty;
#pragma unroll
        for (int j = 2; j < P; ++j) {
          const float t2 = 2.0f * ty * t1 - t0;
          if (j & 1) od = __ffma2_rn(make_float2(t2, t2), af[j * NQ + q], od);
          else ev = __ffma2_rn(make_float2(t2, t2), af[j * NQ + q], ev);
          t0 = t1; t1 = t2;
        }
        const int lm = bh - 1 - lr;
        *reinterpret_cast<float2*>(gfbuf + lr * GFS + 2 * q) = make_float2(ev.x + od.x, ev.y + od.y);
        *reinterpret_cast<float2*>(gfbuf + lm * GFS + 2 * q) = make_float2(ev.x - od.x, ev.y - od.y);
      } else {
        const int lr = o - n32;
        const double ty = (2.0 * lr + 1.0) * inv_bh - 1.0;
        double t0 = 1.0, t1 = ty, ev = a0[0], od = ty * a0[1];
#pragma unroll
        for (int j = 2; j < P; ++j) {
          const double t2 = 2.0 * ty * t1 - t0;
          if (j & 1) od = fma(t2, a0[j], od);
          else ev = fma(t2, a0[j], ev);
          t0 = t1; t1 = t2;
        }
        g0buf[lr] = ev + od;
        g0buf[bh - 1 - lr] = ev - od;
      }
    }
  };

  // one row of the box for this lane's two columns: far field from G, near field from the near list
  auto row_pass = [&](int lr, const float* gfbuf, const double* g0buf, const NearBlk* nb, int cnt, const double* accT,
                      double* dst, bool okA, bool okB) {
    const float4* g4 = reinterpret_cast<const float4*>(gfbuf + lr * GFS);
    float2 eo = make_float2(0.f, 0.f);                  // (odd, even) parts
#pragma unroll
    for (int q = 0; q < GFS / 4; ++q) {
      const float4 g = g4[q];
      eo = __ffma2_rn(make_float2(g.x, g.y), Tp[2 * q], eo);
      eo = __ffma2_rn(make_float2(g.z, g.w), Tp[2 * q + 1], eo);
    }
    float fa = eo.y + eo.x, fb = eo.y - eo.x;
    if (cnt > 0) {
      const float cyl = (float)(hyb * (lr + 0.5));
      auto pair_term = [&](const float4 kn) {
        const float dy = cyl - kn.y, dy2 = dy * dy;
        const float dxa = cxa - kn.x, dxb = cxb - kn.x;
        const float r2a = fmaxf(fmaf(dxa, dxa, dy2), 1e-20f), r2b = fmaxf(fmaf(dxb, dxb, dy2), 1e-20f);
        fa = fmaf(kn.z * r2a, lg2_fast(r2a), fa);
        fb = fmaf(kn.z * r2b, lg2_fast(r2b), fb);
      };
      if (cnt <= kNearInline) {
#pragma unroll 1
        for (int q = 0; q < cnt; ++q) pair_term(nb->e[q]);
      } else {
        const float4* src = near_over + nb->off;
#pragma unroll 1
        for (int q = 0; q < cnt; ++q) pair_term(__ldg(&src[q]));
      }
    }
    const double g0 = g0buf[lr];
    double va = g0 + (double)fa, vb = g0 + (double)fb;
    if (kAcc) {
      va = fma(accT[lr * 32 + l], fz.inv_w, va);
      vb = fma(accT[lr * 32 + 31 - l], fz.inv_w, vb);
    }
    if (okA) __stcs(dst, va);
    if (okB) __stcs(dst + (31 - 2 * l), vb);
  };

  const int n_items = npair * (NQ + 1);
  mbar_wait(&s_full[0], 0);
  collapse(0);
  __syncthreads();
  // consumer cursor: lattice position of the current box and the address of this lane's first cell in it
  int cbi = blockIdx.x % lat.nbx, cbj = blockIdx.x / lat.nbx;
  // Full 32-row boxes: the collapse of the next box occupies warps 0-3, so those take one of the 16 row pairs each
  // and warps 4-7 three each (equal instruction counts per warp between two barriers).  Other boxes: row pairs
  // warp, warp + 8, ...
  const int pfirst = warp < 4 ? warp : 4 + (warp - 4) * 3;
  const int lr0 = 2 * pfirst + half, lrg = 2 * warp + half;
  // element offset of this lane's first cell of the current box (64-bit only where the address is formed)
  int64_t e0 = ((int64_t)cbj * bh + lr0) * stride + (cbi * 32 + l);
  const int64_t gen_off = (int64_t)(lrg - lr0) * stride;
  const int64_t box_step = (int64_t)dJ * bh * stride + dI * 32, wrap_step = (int64_t)bh * stride - (int64_t)lat.nbx * 32;
  const int64_t dstep = 16 * stride;
  const int wcols = w.c1 - w.c0, wrows = w.r1 - w.r0;
  const int full_bi = wcols >> 5, full_bj = bh == 32 ? (wrows >> 5) : 0;   // boxes below these indices are complete 32 x 32
  for (int i = 0; i < nmine; ++i) {
    if (i + 1 < nmine && warp * 32 < n_items) {   // warps without collapse items are ordered behind these by the box barrier
      mbar_wait(&s_full[(i + 1) & (kLeafStages - 1)], ((i + 1) / kLeafStages) & 1);
      collapse(i + 1);
    }
    // ---- rows of box i: warp -> row pairs warp, warp + 8, ... ; lane half -> row of the pair -------------------
    {
      const unsigned char* sp = leaf_smem + (i & (kLeafStages - 1)) * stage_bytes;
      const NearBlk* nb = reinterpret_cast<const NearBlk*>(sp + acc_bytes);
      const double* accT = reinterpret_cast<const double*>(sp);
      const int cnt = nb->cnt;
      const double* g0buf = s_G0 + (i & 1) * bh;
      const float* gfbuf = s_Gf + (i & 1) * bh * GFS;
      double* dst0 = out + e0;
      if (cbi < full_bi && cbj < full_bj) {   // complete box: no predicates
        row_pass(lr0, gfbuf, g0buf, nb, cnt, accT, dst0, true, true);
        if (warp >= 4) {
          row_pass(lr0 + 2, gfbuf, g0buf, nb, cnt, accT, dst0 + 2 * stride, true, true);
          row_pass(lr0 + 4, gfbuf, g0buf, nb, cnt, accT, dst0 + 4 * stride, true, true);
        }
      } else {
        const int rows_here = min(bh, wrows - cbj * bh);
        const bool okA = cbi * 32 + l < wcols, okB = cbi * 32 + 31 - l < wcols;
        double* dst = dst0 + gen_off;
        for (int lr = lrg; lr < rows_here; lr += 16, dst += dstep) row_pass(lr, gfbuf, g0buf, nb, cnt, accT, dst, okA, okB);
      }
    }
    cbi += dI; cbj += dJ; e0 += box_step;
    if (cbi >= lat.nbx) { cbi -= lat.nbx; ++cbj; e0 += wrap_step; }
    __syncthreads();   // stage i % kLeafStages and G buffer i & 1 are free again
    if (warp == 0 && ibox < nmine) issue();
  }
}

// -----------------------------------------------------------------------------------------
// k_leaf_warp - the TPS-only grid evaluation (no accumulator) with ONE WARP PER BOX and no CTA barrier at all.
// k_leaf_stream<P, false> is issue-bound: of its ~1 800 warp-instructions per 1 024-cell box, ~640 are the per-box cursor,
// barrier and staging arithmetic that each of the 8 warps of the CTA repeats, and the collapse of box i + 1 occupies 3 warps
// while 5 wait for the box barrier.  Here a warp owns whole boxes: its own 4-stage record ring (one bulk copy + mbarrier per
// box, issued by lane 0), its own G buffers, collapse (3 items per lane) -> __syncwarp -> 16 row pairs -> __syncwarp.  The per-box
// overhead is paid once instead of eight times, nobody waits for anybody, and 24 warps per SM interleave.  Same arithmetic
// per cell as k_leaf_stream: the output is bit-identical (tools/leaf_check.py).
// -----------------------------------------------------------------------------------------
constexpr int kLeafWarpStages = 4;
__host__ __device__ inline int leaf_warp_bytes(int P, int bh) {      // shared memory of one warp
  return kLeafWarpStages * ((leaf_rec_bytes(P) + 127) / 128 * 128) + ((((bh + 1) & ~1) * 8 + bh * leaf_gfs(P) * 4 + 127) / 128 * 128);
}

template <int P>
__global__ void __launch_bounds__(kLeafThreads, 4) k_leaf_warp(
    Lattice lat, mb_window w, const unsigned char* __restrict__ recs, const float4* __restrict__ near_over,
    const unsigned long long* __restrict__ est_bits, double mixed_threshold, double* __restrict__ out, int64_t stride) {
  static_assert(P % 2 == 0, "P must be even");
  if (!(__longlong_as_double((long long)__ldg(est_bits)) <= mixed_threshold)) return;   // k_leaf_f64 runs instead
  constexpr int GFS = leaf_gfs(P);
  constexpr int NQ = GFS / 2;
  constexpr uint32_t kRecBytes = leaf_rec_bytes(P);
  constexpr int kRecStride = (leaf_rec_bytes(P) + 127) / 128 * 128;
  extern __shared__ __align__(128) unsigned char leaf_smem[];
  __shared__ __align__(8) uint64_t s_full[kLeafThreads / 32][kLeafWarpStages];
  const int bh = lat.bh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* mine = leaf_smem + (size_t)warp * leaf_warp_bytes(P, bh);
  double* g0buf = reinterpret_cast<double*>(mine + kLeafWarpStages * kRecStride);    // [bh]
  float* gfbuf = reinterpret_cast<float*>(g0buf + ((bh + 1) & ~1));                   // [bh][GFS], 16-byte aligned rows
  const int nboxes = lat.nbx * lat.nby;
  const int W = gridDim.x * (kLeafThreads / 32);
  const int wg = blockIdx.x * (kLeafThreads / 32) + warp;
  const int nmine = wg < nboxes ? (nboxes - wg + W - 1) / W : 0;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kLeafWarpStages; ++s) mbar_init(&s_full[warp][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < min(kLeafWarpStages, nmine); ++i) {
      mbar_expect_tx(&s_full[warp][i], kRecBytes);
      bulk_g2s(mine + i * kRecStride, recs + (size_t)(wg + i * W) * kRecBytes, kRecBytes, &s_full[warp][i]);
    }
  }
  __syncwarp();
  const int half = lane >> 4, l = lane & 15;
  float2 Tp[NQ];
  {
    float T[2 * NQ + 1];
    const float tx = (2.0f * l + 1.0f) / 32.0f - 1.0f;
    T[0] = 1.0f;
    T[1] = tx;
#pragma unroll
    for (int k = 2; k <= 2 * NQ; ++k) T[k] = k < P ? 2.0f * tx * T[k - 1] - T[k - 2] : 0.0f;
#pragma unroll
    for (int q = 0; q < NQ; ++q) Tp[q] = make_float2(T[2 * q + 1], T[2 * q + 2]);
  }
  const float cxa = (float)(lat.hx * ((l + 0.5) / 32.0)), cxb = (float)(lat.hx * ((31 - l + 0.5) / 32.0));
  const double hyb = lat.hy / bh, inv_bh = 1.0 / bh;
  const float inv_bh_f = (float)inv_bh;
  const int npair = (bh + 1) >> 1;
  const int n32 = npair * NQ;
  const int wcols = w.c1 - w.c0, wrows = w.r1 - w.r0;
  const int full_bi = wcols >> 5, full_bj = bh == 32 ? (wrows >> 5) : 0;
  const int dI = W % lat.nbx, dJ = W / lat.nbx;
  int cbi = wg % lat.nbx, cbj = wg / lat.nbx;
  for (int i = 0; i < nmine; ++i) {
    const int st = i & (kLeafWarpStages - 1);
    const unsigned char* rec = mine + st * kRecStride;
    mbar_wait(&s_full[warp][st], (i / kLeafWarpStages) & 1);
    // ---- collapse y: G[lr][k] = sum_j T_j(ty_lr) A[j][k] (same statements as k_leaf_stream::collapse) ----------------------
    {
      const double* a0 = reinterpret_cast<const double*>(rec + 256);
      const float2* af = reinterpret_cast<const float2*>(rec + 256 + 8 * P);
      for (int o = lane; o < n32 + npair; o += 32) {
        if (o < n32) {
          const int lr = o / NQ, q = o % NQ;
          const float ty = fmaf((float)(2 * lr + 1), inv_bh_f, -1.0f);
          float t0 = 1.0f, t1 = ty;
          float2 ev = af[q], od = af[NQ + q];
          od.x *= ty; od.y *= ty;
#pragma unroll
          for (int j = 2; j < P; ++j) {
            const float t2 = 2.0f * ty * t1 - t0;
            if (j & 1) od = __ffma2_rn(make_float2(t2, t2), af[j * NQ + q], od);
            else ev = __ffma2_rn(make_float2(t2, t2), af[j * NQ + q], ev);
            t0 = t1; t1 = t2;
          }
          const int lm = bh - 1 - lr;
          *reinterpret_cast<float2*>(gfbuf + lr * GFS + 2 * q) = make_float2(ev.x + od.x, ev.y + od.y);
          *reinterpret_cast<float2*>(gfbuf + lm * GFS + 2 * q) = make_float2(ev.x - od.x, ev.y - od.y);
        } else {
          const int lr = o - n32;
          const double ty = (2.0 * lr + 1.0) * inv_bh - 1.0;
          double t0 = 1.0, t1 = ty, ev = a0[0], od = ty * a0[1];
#pragma unroll
          for (int j = 2; j < P; ++j) {
            const double t2 = 2.0 * ty * t1 - t0;
            if (j & 1) od = fma(t2, a0[j], od);
            else ev = fma(t2, a0[j], ev);
            t0 = t1; t1 = t2;
          }
          g0buf[lr] = ev + od;
          g0buf[bh - 1 - lr] = ev - od;
        }
      }
    }
    __syncwarp();
    // ---- rows: lane = (row of a pair, column pair l / 31 - l) -----------------------------------------------------------------
    {
      const NearBlk* nb = reinterpret_cast<const NearBlk*>(rec);
      const int cnt = nb->cnt;
      const bool full = cbi < full_bi && cbj < full_bj;
      const int rows_here = full ? bh : min(bh, wrows - cbj * bh);
      const bool okA = full || cbi * 32 + l < wcols, okB = full || cbi * 32 + 31 - l < wcols;
      double* dst = out + ((int64_t)cbj * bh + half) * stride + (cbi * 32 + l);
      const int64_t dstep = 2 * stride;
      for (int lr = half; lr < rows_here; lr += 2, dst += dstep) {
        const float4* g4 = reinterpret_cast<const float4*>(gfbuf + lr * GFS);
        float2 eo = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < GFS / 4; ++q) {
          const float4 g = g4[q];
          eo = __ffma2_rn(make_float2(g.x, g.y), Tp[2 * q], eo);
          eo = __ffma2_rn(make_float2(g.z, g.w), Tp[2 * q + 1], eo);
        }
        float fa = eo.y + eo.x, fb = eo.y - eo.x;
        if (cnt > 0) {
          const float cyl = (float)(hyb * (lr + 0.5));
          auto pair_term = [&](const float4 kn) {
            const float dy = cyl - kn.y, dy2 = dy * dy;
            const float dxa = cxa - kn.x, dxb = cxb - kn.x;
            const float r2a = fmaxf(fmaf(dxa, dxa, dy2), 1e-20f), r2b = fmaxf(fmaf(dxb, dxb, dy2), 1e-20f);
            fa = fmaf(kn.z * r2a, lg2_fast(r2a), fa);
            fb = fmaf(kn.z * r2b, lg2_fast(r2b), fb);
          };
          if (cnt <= kNearInline) {
#pragma unroll 1
            for (int q = 0; q < cnt; ++q) pair_term(nb->e[q]);
          } else {
            const float4* src = near_over + nb->off;
#pragma unroll 1
            for (int q = 0; q < cnt; ++q) pair_term(__ldg(&src[q]));
          }
        }
        const double g0 = g0buf[lr];
        const double va = g0 + (double)fa, vb = g0 + (double)fb;
        if (okA) __stcs(dst, va);
        if (okB) __stcs(dst + (31 - 2 * l), vb);
      }
    }
    cbi += dI; cbj += dJ;
    if (cbi >= lat.nbx) { cbi -= lat.nbx; ++cbj; }
    __syncwarp();                                  // the stage and the G buffers are free
    if (lane == 0 && i + kLeafWarpStages < nmine) {
      mbar_expect_tx(&s_full[warp][st], kRecBytes);
      bulk_g2s(mine + st * kRecStride, recs + (size_t)(wg + (i + kLeafWarpStages) * W) * kRecBytes, kRecBytes, &s_full[warp][st]);
    }
  }
}

// -----------------------------------------------------------------------------------------
// k_leaf_f64 - the same evaluation entirely in float64 with the table-driven log (always valid; runs when the
// device-side estimate says the mixed path would not be accurate enough, or on request).  One leaf box per loop
// iteration of a persistent CTA, lane = column:
//   G[lr][k] = sum_j T_j(ty_lr) A[j][k]      (collapse y once per box row)
//   far      = sum_k G[lr][k] T_k(tx_lane)   (T_k(tx_lane) lives in registers)
//   near     = sum over the knots of the 3 x 3 box neighbourhood, streamed through shared memory
// -----------------------------------------------------------------------------------------
constexpr int kNearCap = 256;

template <int P, bool kAcc>
__global__ void __launch_bounds__(kLeafThreads) k_leaf_f64(
    Lattice lat, LevelInfo leaf, mb_window w, const double* __restrict__ coef, const int* __restrict__ start,
    const double4* __restrict__ knots, const double2* __restrict__ logtab,
    const unsigned long long* __restrict__ est_bits, double mixed_threshold, AccFuse fz,
    double* __restrict__ out, int64_t stride) {
  if (__longlong_as_double((long long)__ldg(est_bits)) <= mixed_threshold) return;   // k_leaf_stream ran instead
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* s_log = reinterpret_cast<double2*>(smem_raw);              // 256
  double4* s_near = reinterpret_cast<double4*>(s_log + 256);          // kNearCap
  double* s_A = reinterpret_cast<double*>(s_near + kNearCap);         // P*P
  double* s_G = s_A + P * P;                                          // bh*P
  __shared__ int s_rng[9][2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kLeafThreads / 32;
  for (int i = tid; i < 256; i += kLeafThreads) s_log[i] = logtab[i];
  double Tx[P];
  {
    const double tx = (2.0 * lane + 1.0) / 32.0 - 1.0;
    Tx[0] = 1.0;
    if (P > 1) Tx[1] = tx;
#pragma unroll
    for (int k = 2; k < P; ++k) Tx[k] = 2.0 * tx * Tx[k - 1] - Tx[k - 2];
  }
  const int nboxes = lat.nbx * lat.nby;
  for (int box = blockIdx.x; box < nboxes; box += gridDim.x) {
    const int bi = box % lat.nbx, bj = box / lat.nbx;
    const int I = lat.offx + bi, J = lat.offy + bj;
    __syncthreads();   // previous box done with s_A / s_G / s_near / s_rng
    for (int i = tid; i < P * P; i += kLeafThreads) s_A[i] = coef[leaf.coef_off + (size_t)box * (P * P) + i];
    if (tid < 9) {
      const int i = I + tid % 3 - 1, j = J + tid / 3 - 1;
      const int nside = 1 << lat.L;
      int a = 0, b = 0;
      if (i >= 0 && j >= 0 && i < nside && j < nside) {
        const uint32_t z = morton(i, j);
        a = start[z];
        b = start[z + 1];
      }
      s_rng[tid][0] = a;
      s_rng[tid][1] = b;
    }
    __syncthreads();
    // collapse y: thread -> (lr, k)
    for (int o = tid; o < lat.bh * P; o += kLeafThreads) {
      const int lr = o / P, k = o % P;
      const double ty = (2.0 * lr + 1.0) / lat.bh - 1.0;
      double t0 = 1.0, t1 = ty, g = s_A[k];
      if (P > 1) g = fma(t1, s_A[P + k], g);
#pragma unroll
      for (int j = 2; j < P; ++j) {
        const double t2 = 2.0 * ty * t1 - t0;
        g = fma(t2, s_A[j * P + k], g);
        t0 = t1; t1 = t2;
      }
      s_G[o] = g;
    }
    int tot = 0;
#pragma unroll
    for (int r = 0; r < 9; ++r) tot += s_rng[r][1] - s_rng[r][0];
    const int col = w.c0 + bi * 32 + lane;
    const int row_base = w.r0 + bj * lat.bh;
    const double sx = lat.sxo + lat.hx * (I + (lane + 0.5) / 32.0);
    // knots of the 3 x 3 neighbourhood are streamed through shared memory kNearCap at a time;
    // one chunk is the common case (about one knot per box).
    int base = 0;
    do {
      const int n = min(kNearCap, tot - base);
      __syncthreads();
      for (int i = tid; i < n; i += kLeafThreads) {
        int f = base + i, idx = 0;
#pragma unroll
        for (int r = 0; r < 9; ++r) {
          const int len = s_rng[r][1] - s_rng[r][0];
          if (f >= 0 && f < len) idx = s_rng[r][0] + f;
          f -= len;
        }
        s_near[i] = ldg4(&knots[idx]);
      }
      __syncthreads();
      for (int lr = warp; lr < lat.bh; lr += kWarps) {
        const int row = row_base + lr;
        if (row >= w.r1) break;
        const double sy = lat.syo + lat.hy * (J + (lr + 0.5) / lat.bh);
        double* dst = out + (int64_t)(row - w.r0) * stride + (col - w.c0);
        double a = 0.0;
        if (base == 0) {
#pragma unroll
          for (int k = 0; k < P; ++k) a = fma(s_G[lr * P + k], Tx[k], a);
        } else if (col < w.c1) {
          a = *dst;
        }
#pragma unroll 2
        for (int i = 0; i < n; ++i) {
          const double4 kn = s_near[i];
          const double dx = sx - kn.x, dy = sy - kn.y;
          const double r2 = fmax(fma(dx, dx, dy * dy), kD2Clamp);
          a = fma(kn.w * r2, tlog(r2, s_log), a);
        }
        if (col < w.c1) {
          if (kAcc && base + n >= tot) a = fma(fz.acc[(int64_t)(row - w.r0) * fz.stride + (col - w.c0)], fz.inv_w, a);
          *dst = a;
        }
      }
      base += kNearCap;
    } while (base < tot);
  }
}

// -----------------------------------------------------------------------------------------
// host side: plan + launches
// -----------------------------------------------------------------------------------------
static int choose_p(const mb_ctx* ctx, const mb_spline* s) {
  if (ctx->cheb_p) return ctx->cheb_p;
  // measured interpolation error per unit sum|c| (3 x 3 near field), x3 margin
  static const int ps[5] = {8, 10, 12, 14, 16};
  static const double err[5] = {2.5e-11, 5e-13, 2e-14, 6e-16, 2e-17};
  const double target = 2e-7 * std::max(s->fscale, 1e-300);
  for (int i = 0; i < 5; ++i)
    if (err[i] * s->sum_abs_c <= target) return ps[i];
  return 16;
}


template <class K>
static int leaf_grid(mb_ctx* ctx, K kernel, size_t smem, int nboxes) {
  int occ = 0;
  // a fixed, generous opt-in limit: the attribute is global per kernel, and the tile lanes of tiles_tps launch the
  // same kernel concurrently with different leaf-box heights (a smaller value set by one lane would make the
  // launch of another fail with "invalid argument")
  constexpr size_t kLeafSmemLimit = 200 * 1024;
  if (smem > kLeafSmemLimit) throw Error(MB_E_UNSUPPORTED, "leaf kernel: leaf box too tall for shared memory");
  MB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLeafSmemLimit));
  MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kLeafThreads, smem));
  if (occ < 1) throw Error(MB_E_UNSUPPORTED, "leaf kernel does not fit on an SM (leaf box too tall)");
  return std::max(1, std::min(nboxes, occ * ctx->sm_count));
}

template <int P>
static void run_fast(mb_ctx* ctx, const mb_spline* s, const Lattice& lat, const std::vector<LevelInfo>& levels,
                     size_t coef_total, size_t part_total, const std::vector<int>& start,
                     const std::vector<double4>& sorted, const mb_window& w, double* out, int64_t stride,
                     const AccFuse* fuse, cudaStream_t st) {
  ChebTables* tabs = get_tables(ctx, P);
  Arena& ar = ctx->arena;
  const int* d_start = ar.upload(start.data(), start.size(), st);
  const double4* d_knots = ar.upload(sorted.data(), sorted.size(), st);
  double* d_coef = ar.take_n<double>(coef_total);
  double* d_part = ar.take_n<double>(part_total);
  // {estimate bits, overflow cursor}
  unsigned long long* d_est = ar.take_n<unsigned long long>(2);
  int* d_cursor = reinterpret_cast<int*>(d_est + 1);
  const int nboxes = lat.nbx * lat.nby;
  unsigned char* d_recs = ar.take_n<unsigned char>((size_t)nboxes * leaf_rec_bytes(P));
  float4* d_over = ar.take_n<float4>(9 * sorted.size() + 1);
  constexpr int threads = ((P * P + 31) / 32) * 32;
  for (size_t li = 0; li < levels.size(); ++li) {
    const LevelInfo& lv = levels[li];
    dim3 grid(lv.nI * lv.nJ, lv.nsplit);
    MB_LAUNCH(ctx, "k_far_p2l", st) k_far_p2l<P><<<grid, threads, 0, st>>>(lat, lv, tabs->d_tab.p, d_start, d_knots, ctx->logtab.p, d_part);
    const LevelInfo& par = li ? levels[li - 1] : lv;
    MB_LAUNCH(ctx, "k_far_transform", st) k_far_transform<P><<<lv.nI * lv.nJ, threads, 0, st>>>(lat, lv, par, li ? 1 : 0, tabs->d_tab.p, d_part,
                                                           d_coef, s->d[0], s->d[1], s->d[2]);
  }
  MB_CUDA(cudaGetLastError());
  const LevelInfo& leaf = levels.back();
  // mixed-precision decision (device side).  Explicit cheb_p / eval_precision = 1 force float64.
  double thr = 2e-8 * std::max(s->fscale, 1e-300);
  const double r2max = 4.0 * (lat.hx * lat.hx + lat.hy * lat.hy);
  if (ctx->eval_precision == 1 || (ctx->eval_precision == 0 && ctx->cheb_p != 0) || r2max >= 0.3) thr = -1.0;
  if (ctx->eval_precision == 2) thr = 1e300;
  MB_CUDA(cudaMemsetAsync(d_est, 0, 2 * sizeof(unsigned long long), st));
  if (thr > 0) {
    const double phi_max = r2max * std::fabs(std::log(r2max));
    MB_LAUNCH(ctx, "k_leaf_prep", st) k_leaf_prep<P><<<(nboxes + 7) / 8, 256, 0, st>>>(lat, leaf, d_coef, d_start, d_knots, phi_max, d_est, d_recs, d_over, d_cursor);
    MB_CUDA(cudaGetLastError());
  }
  const AccFuse fz = fuse ? *fuse : AccFuse{nullptr, 0, 0.0};
  // Everything above depends on the spline alone; only the kernels below read the ensemble accumulator.  mb_mltps_predict* hands
  // the join of its ensemble stream over here (instead of waiting for it before the far-field set-up), so that the ~1.5 ms of
  // k_far_p2l / k_far_transform / k_leaf_prep run beside the last ensemble kernel.
  if (ctx->leaf_wait) {
    MB_CUDA(cudaStreamWaitEvent(st, ctx->leaf_wait, 0));
    ctx->leaf_wait = nullptr;
  }
  if (thr > 0) {
    const size_t smem = (size_t)kLeafStages * leaf_stage_bytes(P, lat.bh, fuse != nullptr) +
                        2 * (size_t)lat.bh * (sizeof(double) + leaf_gfs(P) * sizeof(float));
    CUtensorMap amap;
    std::memset(&amap, 0, sizeof amap);
    if (fuse) {
      // all boxes of the lattice lie inside the padded accumulator (acc_rows(w) = rows + 128, stride = columns rounded up to 32)
      const int acc_tma = (ctx->leaf_tma != 2 &&
                           make_f64_matrix_tensor_map(&amap, fz.acc, fz.stride, (int64_t)lat.nby * lat.bh, 32, lat.bh)) ? 1 : 0;
      const int grid = leaf_grid(ctx, k_leaf_stream<P, true>, smem, nboxes);
      MB_LAUNCH(ctx, "k_leaf_fused", st) k_leaf_stream<P, true><<<grid, kLeafThreads, smem, st>>>(amap, acc_tma, lat, w, d_recs, d_over, d_est, thr, fz, out, stride);
    } else if (ctx->leaf_impl != 2) {            // TPS only: one warp per box ("leaf_impl" = 2: the CTA-per-box kernel)
      const size_t wsm = (size_t)(kLeafThreads / 32) * leaf_warp_bytes(P, lat.bh);
      const int grid = leaf_grid(ctx, k_leaf_warp<P>, wsm, (nboxes + kLeafThreads / 32 - 1) / (kLeafThreads / 32));
      MB_LAUNCH(ctx, "k_leaf", st) k_leaf_warp<P><<<grid, kLeafThreads, wsm, st>>>(lat, w, d_recs, d_over, d_est, thr, out, stride);
    } else {
      const int grid = leaf_grid(ctx, k_leaf_stream<P, false>, smem, nboxes);
      MB_LAUNCH(ctx, "k_leaf", st) k_leaf_stream<P, false><<<grid, kLeafThreads, smem, st>>>(amap, 0, lat, w, d_recs, d_over, d_est, thr, fz, out, stride);
    }
    MB_CUDA(cudaGetLastError());
  }
  if (thr < 1e300) {
    const size_t smem = 256 * sizeof(double2) + kNearCap * sizeof(double4) + sizeof(double) * (P * P + (size_t)lat.bh * P);
    if (fuse) {
      const int grid = leaf_grid(ctx, k_leaf_f64<P, true>, smem, nboxes);
      MB_LAUNCH(ctx, "k_leaf_f64_fused", st) k_leaf_f64<P, true><<<grid, kLeafThreads, smem, st>>>(lat, leaf, w, d_coef, d_start, d_knots, ctx->logtab.p, d_est, thr, fz, out, stride);
    } else {
      const int grid = leaf_grid(ctx, k_leaf_f64<P, false>, smem, nboxes);
      MB_LAUNCH(ctx, "k_leaf_f64", st) k_leaf_f64<P, false><<<grid, kLeafThreads, smem, st>>>(lat, leaf, w, d_coef, d_start, d_knots, ctx->logtab.p, d_est, thr, fz, out, stride);
    }
    MB_CUDA(cudaGetLastError());
  }
}

void tps_eval_fast(mb_ctx* ctx, const mb_spline* s, const mb_grid& g, const mb_window& w, double* out,
                   int64_t stride, cudaStream_t st, const AccFuse* fuse) {
  const GridAffine a = make_affine(g, *s);
  const int P = choose_p(ctx, s);
  const double ax = a.rx / a.scx, ay = -a.ry / a.scy;
  int bh = ctx->leaf_rows;
  if (bh <= 0) bh = (int)std::lround(32.0 * ax / std::fabs(ay));
  bh = std::max(4, std::min(128, bh));
  Lattice lat;
  lat.bh = bh;
  lat.hx = ax * 32.0;
  lat.hy = ay * bh;
  lat.nbx = (w.c1 - w.c0 + 31) / 32;
  lat.nby = (w.r1 - w.r0 + bh - 1) / bh;
  // s at the window's west / north cell edge
  const double sxe = ((a.xmin + (w.c0 + 0.5) * a.rx) - a.cx) / a.scx - 0.5 * ax;
  const double sye = ((a.ymax - (w.r0 + 0.5) * a.ry) - a.cy) / a.scy - 0.5 * ay;
  // knot lattice coordinates relative to the window origin
  const int np = s->np;
  std::vector<int> ki(np), kj(np);
  int imin = 0, jmin = 0, imax = lat.nbx - 1, jmax = lat.nby - 1;
  for (int k = 0; k < np; ++k) {
    ki[k] = (int)std::floor((s->sx[k] - sxe) / lat.hx);
    kj[k] = (int)std::floor((s->sy[k] - sye) / lat.hy);
    imin = std::min(imin, ki[k]); imax = std::max(imax, ki[k]);
    jmin = std::min(jmin, kj[k]); jmax = std::max(jmax, kj[k]);
  }
  lat.offx = -imin;
  lat.offy = -jmin;
  const int span = std::max(imax - imin + 1, jmax - jmin + 1);
  int L = 2;
  while ((1 << L) < span) ++L;
  if (L > 12) throw Error(MB_E_UNSUPPORTED, "fast evaluator: knot cloud too far outside the window (lattice > 4096^2 boxes)");
  lat.L = L;
  lat.sxo = sxe - lat.hx * lat.offx;
  lat.syo = sye - lat.hy * lat.offy;
  // Morton-sorted knots + CSR over leaf boxes; c is pre-multiplied only in the transform (E/2)
  const size_t nleaf = (size_t)1 << (2 * L);
  std::vector<int> start(nleaf + 1, 0);
  std::vector<uint32_t> kz(np);
  for (int k = 0; k < np; ++k) {
    kz[k] = morton(ki[k] + lat.offx, kj[k] + lat.offy);
    start[kz[k] + 1]++;
  }
  for (size_t z = 0; z < nleaf; ++z) start[z + 1] += start[z];
  std::vector<int> cursor(start.begin(), start.end() - 1);
  std::vector<double4> sorted(np);
  for (int k = 0; k < np; ++k) {
    const int dst = cursor[kz[k]]++;
    // .z = c for the far field (E/2 applied after the transform), .w = c E/2 for the near field
    sorted[dst] = make_double4(s->sx[k], s->sy[k], s->c[k], s->c[k] * 0.5 * kRbfConst);
  }
  // levels 2..L
  std::vector<LevelInfo> levels;
  size_t coef_total = 0, part_total = 0;
  for (int lvl = 2; lvl <= L; ++lvl) {
    LevelInfo lv;
    lv.lvl = lvl;
    lv.sh = L - lvl;
    lv.I0 = lat.offx >> lv.sh;
    lv.J0 = lat.offy >> lv.sh;
    lv.nI = ((lat.offx + lat.nbx - 1) >> lv.sh) - lv.I0 + 1;
    lv.nJ = ((lat.offy + lat.nby - 1) >> lv.sh) - lv.J0 + 1;
    const int nb = lv.nI * lv.nJ;
    lv.nsplit = std::max(1, std::min(36, (4 * ctx->sm_count + nb - 1) / nb));
    lv.coef_off = coef_total;
    lv.part_off = part_total;
    coef_total += (size_t)nb * P * P;
    part_total += (size_t)nb * lv.nsplit * P * P;
    levels.push_back(lv);
  }
  switch (P) {
    case 8:  run_fast<8>(ctx, s, lat, levels, coef_total, part_total, start, sorted, w, out, stride, fuse, st); break;
    case 10: run_fast<10>(ctx, s, lat, levels, coef_total, part_total, start, sorted, w, out, stride, fuse, st); break;
    case 12: run_fast<12>(ctx, s, lat, levels, coef_total, part_total, start, sorted, w, out, stride, fuse, st); break;
    case 14: run_fast<14>(ctx, s, lat, levels, coef_total, part_total, start, sorted, w, out, stride, fuse, st); break;
    case 16: run_fast<16>(ctx, s, lat, levels, coef_total, part_total, start, sorted, w, out, stride, fuse, st); break;
    default: throw Error(MB_E_ARG, "cheb_p must be one of 8, 10, 12, 14, 16");
  }
}

}  // namespace mb
