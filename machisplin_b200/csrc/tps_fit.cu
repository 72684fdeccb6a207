// Thin-plate smoothing-spline fit on the device - replaces fields::Tps(x, Y) at V73:722 and V73:751
// (Tps -> Krig -> Krig.engine.default "WBW" decomposition -> gcv.Krig -> Krig.coef; SURVEY.md 3.2,
// Appendix A).  m = 2, d = 2, scale.type = "range", method = "GCV" unless a lambda is given.
//
//   K1 k_assemble        W^1/2 K W^1/2,  K_ij = (1/8pi) 1/2 d2 log d2        (N x N float64, coalesced)
//   K1b reflectors       M = Q2' K Q2 by three two-sided Householder updates  (gemv + rank-2)
//   K2 Cholesky          fixed lambda: blocked right-looking LL' of M + lambda I; the trailing SYRK
//                        runs on the FP64 tensor pipe (mma.sync.m8n8k4.f64 -> DMMA)
//   K3 GCV               M = Q T Q' by the in-house persistent Householder kernel (sytrd.cu), eta = eig(T) by Sturm
//                        bisection, z^ = Q'Q2'y; fields' 200-point df grid on the device (k_gcv_grid), golden
//                        section on the host with RSS(lambda) = lambda^2 |(T + lambda I)^-1 z^|^2 (O(m) solves);
//                        coefficients at the selected lambda through the Cholesky of K2.  No eigenvector is
//                        formed.  (mb_set_param "eigen_impl" = 1 selects cuSOLVER Dsyevd - validation only.)
#include "common.cuh"
#include "internal.h"

#include <cusolverDn.h>   // types only: the validation path binds libcusolver at run time, the library does not link it
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <map>
#include <numeric>

namespace mb {

// ---------------------------------------------------------------------------------------------
// library handles kept per context
// ---------------------------------------------------------------------------------------------
// cuSOLVER serves ONE purpose: "eigen_impl" = 1, the validation of the in-house GCV fit against Dsyevd (tests, tools/fit_check.py).
// It is resolved with dlopen when that parameter is first used, so the product library carries no dependency on it.
struct CusolverApi {
  void* handle = nullptr;
  cusolverStatus_t (*Create)(cusolverDnHandle_t*) = nullptr;
  cusolverStatus_t (*Destroy)(cusolverDnHandle_t) = nullptr;
  cusolverStatus_t (*SetStream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
  cusolverStatus_t (*DsyevdBufferSize)(cusolverDnHandle_t, cusolverEigMode_t, cublasFillMode_t, int, const double*, int,
                                       const double*, int*) = nullptr;
  cusolverStatus_t (*Dsyevd)(cusolverDnHandle_t, cusolverEigMode_t, cublasFillMode_t, int, double*, int, double*, double*, int,
                             int*) = nullptr;
};
static CusolverApi& cusolver() {
  static CusolverApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    for (const char* n : {"libcusolver.so.11", "libcusolver.so.12", "libcusolver.so"}) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (api.handle) break;
    }
    if (api.handle) {
      api.Create = reinterpret_cast<decltype(api.Create)>(dlsym(api.handle, "cusolverDnCreate"));
      api.Destroy = reinterpret_cast<decltype(api.Destroy)>(dlsym(api.handle, "cusolverDnDestroy"));
      api.SetStream = reinterpret_cast<decltype(api.SetStream)>(dlsym(api.handle, "cusolverDnSetStream"));
      api.DsyevdBufferSize = reinterpret_cast<decltype(api.DsyevdBufferSize)>(dlsym(api.handle, "cusolverDnDsyevd_bufferSize"));
      api.Dsyevd = reinterpret_cast<decltype(api.Dsyevd)>(dlsym(api.handle, "cusolverDnDsyevd"));
    }
  }
  if (!api.handle || !api.Create || !api.Destroy || !api.SetStream || !api.DsyevdBufferSize || !api.Dsyevd)
    throw Error(MB_E_UNSUPPORTED, "eigen_impl = 1 (validation against cuSOLVER Dsyevd) needs libcusolver, which could not be loaded");
  return api;
}
struct FitLibs {
  cusolverDnHandle_t solver = nullptr;
};
static std::map<mb_ctx*, FitLibs>& libs_map() {
  static std::map<mb_ctx*, FitLibs> m;
  return m;
}
static FitLibs& libs(mb_ctx* ctx) {
  FitLibs& l = libs_map()[ctx];
  if (!l.solver) {
    if (cusolver().Create(&l.solver) != CUSOLVER_STATUS_SUCCESS) throw Error(MB_E_CUDA, "cusolverDnCreate failed");
    cusolver().SetStream(l.solver, ctx->stream);
  }
  return l;
}
void fit_release(mb_ctx* ctx) {
  auto it = libs_map().find(ctx);
  if (it == libs_map().end()) return;
  if (it->second.solver) cusolver().Destroy(it->second.solver);
  libs_map().erase(it);
}

// ---------------------------------------------------------------------------------------------
// K1: assemble sqrt(w_i) K_ij sqrt(w_j), column-major with leading dimension ld
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_assemble(const double* __restrict__ sx, const double* __restrict__ sy,
                                                  const double* __restrict__ w2, int np, double* __restrict__ K,
                                                  int ld) {
  const int i = blockIdx.x * 32 + (threadIdx.x & 31);
  const int j0 = blockIdx.y * 32 + (threadIdx.x >> 5) * 4;
  if (i >= np) return;
  const double xi = sx[i], yi = sy[i], wi = w2[i];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int j = j0 + q;
    if (j < np) {
      const double dx = xi - sx[j], dy = yi - sy[j];
      const double r2 = fmax(dx * dx + dy * dy, kD2Clamp);
      K[(size_t)j * ld + i] = wi * w2[j] * (kRbfConst * (0.5 * log(r2) * r2));
    }
  }
}

// y_partial[chunk][i] = sum_{j in chunk} A[i, j] x[j]   (column-major, rows coalesced)
constexpr int kGemvChunks = 32;
__global__ void __launch_bounds__(128) k_gemv_n_partial(const double* __restrict__ A, int ld, int m, int n,
                                                        const double* __restrict__ x, double* __restrict__ part) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  const int chunk = blockIdx.y;
  const int per = (n + kGemvChunks - 1) / kGemvChunks;
  const int j0 = chunk * per, j1 = min(n, j0 + per);
  if (i >= m) return;
  double acc = 0.0;
  for (int j = j0; j < j1; ++j) acc = fma(A[(size_t)j * ld + i], x[j], acc);
  part[(size_t)chunk * m + i] = acc;
}
__global__ void k_gemv_n_reduce(const double* __restrict__ part, int m, double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double acc = 0.0;
  for (int c = 0; c < kGemvChunks; ++c) acc += part[(size_t)c * m + i];
  y[i] = acc;
}
// y[j] = sum_i A[i, j] x[i] : one warp per column
__global__ void __launch_bounds__(256) k_gemv_t(const double* __restrict__ A, int ld, int m, int n,
                                                const double* __restrict__ x, double* __restrict__ y) {
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= n) return;
  double acc = 0.0;
  for (int i = lane; i < m; i += 32) acc = fma(A[(size_t)j * ld + i], x[i], acc);
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[j] = acc;
}

// q = 2 p - 2 (v'p) v   (single CTA; H K H = K - v q' - q v')
__global__ void __launch_bounds__(1024) k_make_q(const double* __restrict__ v, const double* __restrict__ p, int n,
                                                 double* __restrict__ q) {
  __shared__ double s_red[32];
  __shared__ double s_alpha;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc = fma(v[i], p[i], acc);
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) a += s_red[i];
    s_alpha = a;
  }
  __syncthreads();
  const double alpha = s_alpha;
  for (int i = threadIdx.x; i < n; i += blockDim.x) q[i] = 2.0 * p[i] - 2.0 * alpha * v[i];
}
__global__ void __launch_bounds__(256) k_rank2(double* __restrict__ K, int ld, int n, const double* __restrict__ v,
                                               const double* __restrict__ q) {
  const int i = blockIdx.x * 32 + (threadIdx.x & 31);
  const int j0 = blockIdx.y * 32 + (threadIdx.x >> 5) * 4;
  if (i >= n) return;
  const double vi = v[i], qi = q[i];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int j = j0 + t;
    if (j < n) K[(size_t)j * ld + i] -= vi * q[j] + qi * v[j];
  }
}

// ---------------------------------------------------------------------------------------------
// K2: blocked Cholesky (lower) of the m x m matrix at A (column-major, ld), NB = 64
// ---------------------------------------------------------------------------------------------
constexpr int kNB = 64;

__global__ void k_add_diag(double* A, int ld, int m, double lam) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) A[(size_t)i * ld + i] += lam;
}

// Factor the kb x kb diagonal block in shared memory (256 threads: thread = (row, column slice) of the rank-1
// update) and leave its inverse in `inv` (64 x 64, row-major, zero above the diagonal): the panel solve and the
// substitution sweeps then are plain matrix products.  info = 1 + failing column if not SPD.
__global__ void __launch_bounds__(256) k_potrf_diag(double* __restrict__ A, int ld, int kb, int col0,
                                                    double* __restrict__ inv, int* __restrict__ info) {
  extern __shared__ __align__(16) double potrf_smem[];
  double (*s)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(potrf_smem);
  double (*sx)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(potrf_smem + kNB * (kNB + 1));   // inverse, built row by row
  double (*sp)[kNB] = reinterpret_cast<double (*)[kNB]>(potrf_smem + 2 * kNB * (kNB + 1));
  const int t = threadIdx.x & 63, cq = threadIdx.x >> 6;
  for (int j = cq; j < kNB; j += 4) {
    s[t][j] = (t < kb && j < kb) ? A[(size_t)j * ld + t] : (t == j ? 1.0 : 0.0);   // identity padding
    sx[t][j] = 0.0;
  }
  __syncthreads();
  for (int j = 0; j < kb; ++j) {
    const double djj = s[j][j];
    if (!(djj > 0.0)) {
      if (threadIdx.x == 0 && *info == 0) *info = col0 + j + 1;
      return;                                   // uniform: every thread reads the same s[j][j]
    }
    const double rl = rsqrt(djj);
    __syncthreads();                            // everyone has read the pivot
    if (cq == 0) {
      if (t == j) s[j][j] = djj * rl;
      else if (t > j && t < kb) s[t][j] *= rl;
    }
    __syncthreads();
    if (t > j && t < kb) {
      const double ltj = s[t][j];
      int c = j + 1 + cq;
      for (; c + 12 <= t; c += 16) {            // four independent updates in flight
        const double u0 = s[c][j], u1 = s[c + 4][j], u2 = s[c + 8][j], u3 = s[c + 12][j];
        const double v0 = s[t][c], v1 = s[t][c + 4], v2 = s[t][c + 8], v3 = s[t][c + 12];
        s[t][c] = fma(-ltj, u0, v0); s[t][c + 4] = fma(-ltj, u1, v1);
        s[t][c + 8] = fma(-ltj, u2, v2); s[t][c + 12] = fma(-ltj, u3, v3);
      }
      for (; c <= t; c += 4) s[t][c] = fma(-ltj, s[c][j], s[t][c]);
    }
    __syncthreads();                            // the next pivot is final
  }
  // inverse: column c of X = L^-1 by forward substitution, the sum over k split over the 4 slices of a column
  // X[r][c] = (delta_rc - sum_{k=c}^{r-1} L[r][k] X[k][c]) / L[r][r]; rows are processed in lock step (one barrier each)
  for (int r = 0; r < kNB; ++r) {
    double part = 0.0;
    if (t <= r)                                   // t = column c of X
      for (int k = t + cq; k < r; k += 4) part = fma(s[r][k], sx[k][t], part);
    sp[cq][t] = part;
    __syncthreads();
    if (cq == 0 && t <= r) {
      const double sum = sp[0][t] + sp[1][t] + sp[2][t] + sp[3][t];
      sx[r][t] = ((t == r ? 1.0 : 0.0) - sum) / s[r][r];
    }
    __syncthreads();
  }
  for (int j = cq; j < kb; j += 4)
    if (t < kb) A[(size_t)j * ld + t] = (t >= j) ? s[t][j] : 0.0;
  for (int j = cq; j < kNB; j += 4) inv[(size_t)t * kNB + j] = sx[t][j];
}

// rows below the diagonal block: X = A21 L11^-T = A21 inv', one thread per row with the row in registers and four
// independent dot products in flight; right-hand sides ride along as extra rows.
__global__ void __launch_bounds__(128) k_trsm_panel(const double* __restrict__ inv, double* __restrict__ A21, int ld,
                                                    int kb, int nrows) {
  __shared__ double s[kNB][kNB + 1];            // s[j][c] = inv[j][c]
  for (int idx = threadIdx.x; idx < kNB * kNB; idx += blockDim.x) s[idx / kNB][idx % kNB] = inv[idx];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  double a[kNB];
#pragma unroll
  for (int c = 0; c < kNB; ++c) a[c] = (c < kb) ? A21[(size_t)c * ld + i] : 0.0;
#pragma unroll
  for (int j0 = 0; j0 < kNB; j0 += 4) {
    double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
#pragma unroll
    for (int c = 0; c < j0 + 4; ++c) {          // inv is lower triangular: columns c <= j
      x0 = fma(a[c], s[j0][c], x0);
      x1 = fma(a[c], s[j0 + 1][c], x1);
      x2 = fma(a[c], s[j0 + 2][c], x2);
      x3 = fma(a[c], s[j0 + 3][c], x3);
    }
    if (j0 < kb) A21[(size_t)j0 * ld + i] = x0;
    if (j0 + 1 < kb) A21[(size_t)(j0 + 1) * ld + i] = x1;
    if (j0 + 2 < kb) A21[(size_t)(j0 + 2) * ld + i] = x2;
    if (j0 + 3 < kb) A21[(size_t)(j0 + 3) * ld + i] = x3;
  }
}

// trailing update on the FP64 tensor pipe:  C(lower tiles) -= P P',  P = panel (n x kb, column-major ld)
// CTA = 64 x 64 tile of C, 8 warps; warp w owns tile rows [8w, 8w+8) and all 8 column blocks.
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256) k_syrk_dmma(const double* __restrict__ Pn, double* __restrict__ C, int ld,
                                                   int n, int kb) {
  const int ti = blockIdx.x, tj = blockIdx.y;
  if (tj > ti) return;
  constexpr int S = kNB + 4;               // padded stride: conflict-free 8 x 4 fragment reads
  extern __shared__ __align__(16) double s_syrk[];
  double* sA = s_syrk;                     // sA[row * S + k]
  double* sB = s_syrk + kNB * S;
  const int i0 = ti * kNB, j0 = tj * kNB;
  for (int idx = threadIdx.x; idx < kNB * kNB; idx += 256) {
    const int r = idx % kNB, k = idx / kNB;  // coalesced along rows of the column-major panel
    sA[r * S + k] = (i0 + r < n && k < kb) ? Pn[(size_t)k * ld + i0 + r] : 0.0;
    sB[r * S + k] = (j0 + r < n && k < kb) ? Pn[(size_t)k * ld + j0 + r] : 0.0;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int fr = lane >> 2, fk = lane & 3;
  double acc[8][2];
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) acc[nb][0] = acc[nb][1] = 0.0;
#pragma unroll 4
  for (int k0 = 0; k0 < kNB; k0 += 4) {
    const double a = sA[(warp * 8 + fr) * S + k0 + fk];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const double b = sB[(nb * 8 + fr) * S + k0 + fk];
      dmma_m8n8k4(acc[nb][0], acc[nb][1], a, b);
    }
  }
  const int gi = i0 + warp * 8 + fr;
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int gj = j0 + nb * 8 + fk * 2 + e;
      if (gi < n && gj < n && gj <= gi) C[(size_t)gj * ld + gi] -= acc[nb][e];
    }
  }
}

// Forward (L y = b) / backward (L' x = y) substitution with the Cholesky factor, one launch per 64-column block,
// right-looking in both directions: every CTA solves the 64 x 64 triangular block redundantly (warp 0, operands
// in shared memory) and then eliminates it from its share of the remaining unknowns - rows below the block
// (forward; coalesced over rows) or columns left of it (backward; each thread streams 64 contiguous doubles).
// The solved block goes to `out`, the remaining right-hand side is updated in place: no two CTAs touch the same
// element within a launch.
template <bool kBackward>
__global__ void __launch_bounds__(256) k_chol_sweep(const double* __restrict__ Lm, int ld, int m, int k0, int kb,
                                                    const double* __restrict__ inv, double* __restrict__ B,
                                                    double* __restrict__ out, int ldb, int nrhs) {
  __shared__ double s_I[kNB][kNB + 1];          // inverse of the diagonal block (row-major, lower triangular)
  __shared__ double s_b[kNB], s_x[kNB];
  for (int idx = threadIdx.x; idx < kNB * kNB; idx += blockDim.x) s_I[idx / kNB][idx % kNB] = inv[idx];
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int r = 0; r < nrhs; ++r) {
    double* b = B + (size_t)r * ldb;
    __syncthreads();
    if (threadIdx.x < kNB) s_b[threadIdx.x] = threadIdx.x < kb ? b[k0 + threadIdx.x] : 0.0;
    __syncthreads();
    if (threadIdx.x < kNB) {
      const int j = threadIdx.x;
      double a0 = 0.0, a1 = 0.0;
      if (!kBackward) {                         // y = L^-1 b
        for (int c = 0; c + 1 <= j; c += 2) { a0 = fma(s_I[j][c], s_b[c], a0); a1 = fma(s_I[j][c + 1], s_b[c + 1], a1); }
        if (!(j & 1)) a0 = fma(s_I[j][j], s_b[j], a0);
      } else {                                  // x = L^-T y
        for (int c = j; c + 1 < kNB; c += 2) { a0 = fma(s_I[c][j], s_b[c], a0); a1 = fma(s_I[c + 1][j], s_b[c + 1], a1); }
        if ((kNB - j) & 1) a0 = fma(s_I[kNB - 1][j], s_b[kNB - 1], a0);
      }
      s_x[j] = a0 + a1;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x < kb) out[(size_t)r * ldb + k0 + threadIdx.x] = s_x[threadIdx.x];
    if (!kBackward) {
      const int i = k0 + kb + gid;
      if (i < m) {
        double acc = b[i];
#pragma unroll 8
        for (int j = 0; j < kb; ++j) acc = fma(-Lm[(size_t)(k0 + j) * ld + i], s_x[j], acc);
        b[i] = acc;
      }
    } else {
      const int c = gid;
      if (c < k0) {
        const double* col = Lm + (size_t)c * ld + k0;
        double acc = b[c];
#pragma unroll 8
        for (int j = 0; j < kb; ++j) acc = fma(-col[j], s_x[j], acc);
        b[c] = acc;
      }
    }
  }
}

// x <- (L L')^-1 b for nrhs right-hand sides: B (m x nrhs, ldb) is overwritten with the solution; tmp is
// scratch of the same shape.
static void cholesky_solve(mb_ctx* ctx, const double* Lm, int ld, int m, const double* inv, double* B, double* tmp,
                           int ldb, int nrhs, cudaStream_t st) {
  for (int k0 = 0; k0 < m; k0 += kNB) {          // L y = b : y -> tmp
    const int kb = std::min(kNB, m - k0);
    const int rest = m - k0 - kb;
    MB_LAUNCH(ctx, "k_chol_sweep", st) k_chol_sweep<false><<<std::max(1, (rest + 255) / 256), 256, 0, st>>>(Lm, ld, m, k0, kb, inv + (size_t)(k0 / kNB) * kNB * kNB, B, tmp, ldb, nrhs);
  }
  for (int k1 = m; k1 > 0;) {                    // L' x = y : x -> B
    const int k0 = ((k1 - 1) / kNB) * kNB;
    const int kb = k1 - k0;
    MB_LAUNCH(ctx, "k_chol_sweep", st) k_chol_sweep<true><<<std::max(1, (k0 + 255) / 256), 256, 0, st>>>(Lm, ld, m, k0, kb, inv + (size_t)(k0 / kNB) * kNB * kNB, tmp, B, ldb, nrhs);
    k1 = k0;
  }
  MB_CUDA(cudaGetLastError());
}

// returns the inverses of the 64 x 64 diagonal blocks of L (block k at + k 64 64), for cholesky_solve
static const double* cholesky_lower(mb_ctx* ctx, double* A, int ld, int m, cudaStream_t st) {
  constexpr size_t kSyrkSmem = 2 * kNB * (kNB + 4) * sizeof(double);
  constexpr size_t kPotrfSmem = (2 * kNB * (kNB + 1) + 4 * kNB) * sizeof(double);
  MB_CUDA(cudaFuncSetAttribute(k_syrk_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSyrkSmem));
  MB_CUDA(cudaFuncSetAttribute(k_potrf_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPotrfSmem));
  ABuf<int> d_info(ctx->arena, 1);
  double* d_inv = ctx->arena.take_n<double>((size_t)((m + kNB - 1) / kNB) * kNB * kNB);
  MB_CUDA(cudaMemsetAsync(d_info.p, 0, sizeof(int), st));
  for (int k = 0; k < m; k += kNB) {
    const int kb = std::min(kNB, m - k);
    double* Akk = A + (size_t)k * ld + k;
    double* invk = d_inv + (size_t)(k / kNB) * kNB * kNB;
    MB_LAUNCH(ctx, "k_potrf_diag", st) k_potrf_diag<<<1, 256, kPotrfSmem, st>>>(Akk, ld, kb, k, invk, d_info.p);
    const int rest = m - k - kb;
    if (rest > 0) {
      MB_LAUNCH(ctx, "k_trsm_panel", st) k_trsm_panel<<<(rest + 127) / 128, 128, 0, st>>>(invk, Akk + kb, ld, kb, rest);
      const int nt = (rest + kNB - 1) / kNB;
      MB_LAUNCH(ctx, "k_syrk_dmma", st) k_syrk_dmma<<<dim3(nt, nt), 256, kSyrkSmem, st>>>(Akk + kb, A + (size_t)(k + kb) * ld + (k + kb), ld, rest, kb);
    }
  }
  int info = 0;
  MB_CUDA(cudaMemcpyAsync(&info, d_info.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
  MB_CUDA(cudaGetLastError());
  if (info != 0)
    throw Error(MB_E_NUMERIC, "Cholesky: Q2'KQ2 + lambda I is not positive definite at column " + std::to_string(info));
  return d_inv;
}

// ---------------------------------------------------------------------------------------------
// host: Householder QR of sqrt(w) [1, s1, s2], replicate pooling, fields' GCV search
// ---------------------------------------------------------------------------------------------
struct QRT {
  int n = 0;
  std::vector<double> v[3];   // unit Householder vectors, v[j][i] = 0 for i < j
  double R[3][3] = {{0}};
  explicit QRT(const std::vector<double>* T, int n_) : n(n_) {
    std::vector<double> A[3] = {T[0], T[1], T[2]};
    for (int j = 0; j < 3; ++j) {
      double nrm = 0;
      for (int i = j; i < n; ++i) nrm += A[j][i] * A[j][i];
      nrm = std::sqrt(nrm);
      const double alpha = -std::copysign(nrm, A[j][j] != 0 ? A[j][j] : 1.0);
      v[j].assign(n, 0.0);
      for (int i = j; i < n; ++i) v[j][i] = A[j][i];
      v[j][j] -= alpha;
      double vn = 0;
      for (int i = j; i < n; ++i) vn += v[j][i] * v[j][i];
      vn = std::sqrt(vn);
      if (!(vn > 0)) throw Error(MB_E_NUMERIC, "Regression matrix for fixed part of model is colinear");
      for (int i = j; i < n; ++i) v[j][i] /= vn;
      for (int c = j; c < 3; ++c) {
        double dot = 0;
        for (int i = j; i < n; ++i) dot += v[j][i] * A[c][i];
        for (int i = j; i < n; ++i) A[c][i] -= 2.0 * dot * v[j][i];
      }
    }
    double dmax = 0, dmin = 1e300;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) R[r][c] = (c >= r) ? A[c][r] : 0.0;
    for (int r = 0; r < 3; ++r) { dmax = std::max(dmax, std::fabs(R[r][r])); dmin = std::min(dmin, std::fabs(R[r][r])); }
    if (dmin < 1e-12 * dmax) throw Error(MB_E_NUMERIC, "Regression matrix for fixed part of model is colinear");
  }
  void qty(std::vector<double>& b) const {
    for (int j = 0; j < 3; ++j) {
      double dot = 0;
      for (int i = j; i < n; ++i) dot += v[j][i] * b[i];
      for (int i = j; i < n; ++i) b[i] -= 2.0 * dot * v[j][i];
    }
  }
  void qy(std::vector<double>& b) const {
    for (int j = 2; j >= 0; --j) {
      double dot = 0;
      for (int i = j; i < n; ++i) dot += v[j][i] * b[i];
      for (int i = j; i < n; ++i) b[i] -= 2.0 * dot * v[j][i];
    }
  }
  void coef(std::vector<double> b, double* d) const {   // R^-1 Q1' b
    qty(b);
    for (int r = 2; r >= 0; --r) {
      double s = b[r];
      for (int c = r + 1; c < 3; ++c) s -= R[r][c] * d[c];
      d[r] = s / R[r][r];
    }
  }
};

namespace gcv {
static double tr_a(double lam, const std::vector<double>& D) {
  double s = 0;
  for (double d : D) s += 1.0 / (1.0 + lam * d);
  return s;
}
// RSS(lambda) = sum_k (u_k lambda D_k / (1 + lambda D_k))^2, from the rotated data u (eigenbasis) ...
struct RssEigen {
  const std::vector<double>& D;
  const std::vector<double>& u;
  double operator()(double lam) const {
    double rss = 0;
    for (size_t k = 0; k < D.size(); ++k) {
      const double lD = D[k] * lam;
      const double t = (u[k] * lD) / (1.0 + lD);
      rss += t * t;
    }
    return rss;
  }
};
// ... or, equivalently, lambda^2 |(T + lambda I)^-1 z^|^2 from the tridiagonal form (no eigenvectors): one
// symmetric positive definite tridiagonal solve (Thomas) per evaluation
struct RssTridiag {
  const std::vector<double>& dg;
  const std::vector<double>& of;
  const std::vector<double>& zh;
  mutable std::vector<double> piv, rhs;
  double operator()(double lam) const {
    const size_t m = dg.size();
    piv.resize(m); rhs.resize(m);
    piv[0] = dg[0] + lam; rhs[0] = zh[0];
    for (size_t i = 1; i < m; ++i) {
      const double w = of[i - 1] / piv[i - 1];
      piv[i] = dg[i] + lam - w * of[i - 1];
      rhs[i] = zh[i] - w * rhs[i - 1];
    }
    double x = rhs[m - 1] / piv[m - 1], rss = x * x;
    for (size_t i = m - 1; i-- > 0;) {
      x = (rhs[i] - of[i] * x) / piv[i];
      rss += x * x;
    }
    return lam * lam * rss;
  }
};
template <class R>
static double value(double lam, const std::vector<double>& D, const R& rssfun, int n_obs, double pure_ss) {
  const int np = (int)D.size();
  const double rss = rssfun(lam);
  double tra = 0;
  for (int k = 0; k < np; ++k) tra += 1.0 / (1.0 + D[k] * lam);
  double mse = rss / np;
  if (n_obs - np > 0) mse += pure_ss / (n_obs - np);
  const double den = 1.0 - ((tra - 3.0) + 3.0) / np;   // cost = 1, offset = 0, nt = 3
  return den > 0 ? mse / (den * den) : NAN;
}
static double df_to_lambda(double df, const std::vector<double>& D) {
  double l1 = 1.0;
  for (int k = 0; k < 25; ++k) { if (tr_a(l1, D) <= df) break; l1 *= 4.0; }
  double l2 = 1.0;
  for (int k = 0; k < 25; ++k) { if (tr_a(l2, D) >= df) break; l2 /= 4.0; }
  double x1 = std::log(l1), x2 = std::log(l2);
  double f1 = tr_a(std::exp(x1), D) - df, f2 = tr_a(std::exp(x2), D) - df;
  if (f1 > f2) throw Error(MB_E_NUMERIC, "bisection.search: f1 must be < f2");
  for (int k = 0; k < 25; ++k) {
    const double xm = (x1 + x2) / 2.0;
    const double fm = tr_a(std::exp(xm), D) - df;
    if (fm < 0) { x1 = xm; f1 = fm; } else { x2 = xm; f2 = fm; }
    if (std::fabs(fm) < 1e-5) break;
  }
  return std::exp((x1 + x2) / 2.0);
}
static std::vector<double> lambda_grid(const std::vector<double>& D) {
  const int np = (int)D.size(), nstep = 200;
  std::vector<double> g(nstep);
  for (int k = 0; k < nstep; ++k) {
    double df = 3.0 + (0.95 * np - 3.0) * k / (nstep - 1.0);
    if (k == 0) df += 0.001;
    g[k] = df_to_lambda(df, D);
  }
  std::sort(g.begin(), g.end());
  return g;
}
template <class F>
static double golden(double ax, double bx, double cx, F f, double tol) {
  const double r = 0.61803399, con = 1.0 - r;
  double x0 = ax, x3 = cx, x1, x2;
  if (std::fabs(cx - bx) > std::fabs(bx - ax)) { x1 = bx; x2 = bx + con * (bx - ax); }
  else { x2 = bx; x1 = bx - con * (bx - ax); }
  double f1 = f(x1), f2 = f(x2);
  for (int k = 0; k < 25; ++k) {
    if (f2 < f1) { x0 = x1; x1 = x2; x2 = r * x1 + con * x3; f1 = f2; f2 = f(x2); }
    else { x3 = x2; x2 = x1; x1 = r * x2 + con * x0; f2 = f1; f1 = f(x1); }
    if (std::fabs(f2 - f1) < tol) break;
  }
  return f1 < f2 ? x1 : x2;
}
template <class R>
static double search(const std::vector<double>& D, const R& u, const std::vector<double>& grid,
                     int n_obs, double pure_ss, double* gcv_min) {
  std::vector<double> lg, gg;
  for (double l : grid) {
    const double g = value(l, D, u, n_obs, pure_ss);
    if (!std::isnan(g)) { lg.push_back(l); gg.push_back(g); }
  }
  if (lg.empty()) throw Error(MB_E_NUMERIC, "GCV is undefined on the whole lambda grid");
  const int il = (int)(std::min_element(gg.begin(), gg.end()) - gg.begin());
  double lam = lg[il];
  if (il > 0 && il < (int)lg.size() - 1) {
    auto f = [&](double l) { return value(l, D, u, n_obs, pure_ss); };
    lam = golden(lg[il - 1], lg[il], lg[il + 1], f, 1e-5 * gg[il]);
  }  // else: fields warns "GCV search gives a minimum at the endpoints of the grid search" and keeps the endpoint
  if (gcv_min) *gcv_min = value(lam, D, u, n_obs, pure_ss);
  return lam;
}
}  // namespace gcv

// ---------------------------------------------------------------------------------------------
// K3: fields' lambda grid on the device.  One CTA per grid point runs Krig.df.to.lambda for its
// target df (x4 bracketing, then bisection.search on log lambda, <= 25 steps, |f| < 1e-5) with
// trA(lambda) = sum_k 1 / (1 + lambda D_k) reduced in a fixed order (deterministic).  The 200
// searches are independent, which is all the parallelism there is: ~15 000 O(np) trace evaluations
// that cost ~35 ms of scalar host time at np = 5000 take < 1 ms here.
// ---------------------------------------------------------------------------------------------
constexpr int kGcvThreads = 256;

__device__ __forceinline__ double block_tr_a(double lam, const double* __restrict__ s_D, int np, double* s_red) {
  double acc = 0.0;
  for (int k = threadIdx.x; k < np; k += kGcvThreads) acc += 1.0 / (1.0 + lam * s_D[k]);
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __syncthreads();                       // s_red may still be read by the previous evaluation
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int q = 0; q < kGcvThreads / 32; ++q) tot += s_red[q];
  return tot;
}

__global__ void __launch_bounds__(kGcvThreads) k_gcv_grid(const double* __restrict__ D, int np, int nstep,
                                                         double* __restrict__ grid, int* __restrict__ err) {
  extern __shared__ double s_D[];
  __shared__ double s_red[kGcvThreads / 32];
  for (int k = threadIdx.x; k < np; k += kGcvThreads) s_D[k] = D[k];
  __syncthreads();
  const int g = blockIdx.x;
  double df = 3.0 + (0.95 * np - 3.0) * g / (nstep - 1.0);
  if (g == 0) df += 0.001;
  double l1 = 1.0;
  for (int k = 0; k < 25; ++k) { if (block_tr_a(l1, s_D, np, s_red) <= df) break; l1 *= 4.0; }
  double l2 = 1.0;
  for (int k = 0; k < 25; ++k) { if (block_tr_a(l2, s_D, np, s_red) >= df) break; l2 /= 4.0; }
  double x1 = log(l1), x2 = log(l2);
  const double f1 = block_tr_a(exp(x1), s_D, np, s_red) - df, f2 = block_tr_a(exp(x2), s_D, np, s_red) - df;
  if (f1 > f2) { if (threadIdx.x == 0) *err = 1; return; }     // bisection.search: "f1 must be < f2"
  for (int k = 0; k < 25; ++k) {
    const double xm = (x1 + x2) / 2.0;
    const double fm = block_tr_a(exp(xm), s_D, np, s_red) - df;
    if (fm < 0) x1 = xm; else x2 = xm;
    if (fabs(fm) < 1e-5) break;
  }
  if (threadIdx.x == 0) grid[g] = exp((x1 + x2) / 2.0);
}

// ---------------------------------------------------------------------------------------------
// the fit
// ---------------------------------------------------------------------------------------------
static void gemv_n(mb_ctx* ctx, const double* A, int ld, int m, int n, const double* x, double* y, double* part,
                   cudaStream_t st) {
  MB_LAUNCH(ctx, "k_gemv_n_partial", st) k_gemv_n_partial<<<dim3((m + 127) / 128, kGemvChunks), 128, 0, st>>>(A, ld, m, n, x, part);
  MB_LAUNCH(ctx, "k_gemv_n_reduce", st) k_gemv_n_reduce<<<(m + 255) / 256, 256, 0, st>>>(part, m, y);
}

void tps_fit(mb_ctx* ctx, const double* xy, const double* y, int n, int L, double lambda, mb_spline** out) {
  cudaStream_t st = ctx->stream;
  for (int r = 0; r < L; ++r) out[r] = nullptr;
  ctx->band_form = mb_band_form();   // arena pointers of an earlier fit: only the two-stage reduction of THIS call may set it
  // ---- transformx (scale.type = "range") + Krig.replicates -------------------------------------
  double cmin[2] = {1e300, 1e300}, cmax[2] = {-1e300, -1e300};
  for (int i = 0; i < n; ++i)
    for (int a = 0; a < 2; ++a) {
      const double v = xy[(size_t)a * n + i];
      MB_REQUIRE(std::isfinite(v), "non-finite coordinate in xy");
      cmin[a] = std::min(cmin[a], v);
      cmax[a] = std::max(cmax[a], v);
    }
  const double scale[2] = {cmax[0] - cmin[0], cmax[1] - cmin[1]};
  if (!(scale[0] > 0 && scale[1] > 0)) throw Error(MB_E_NUMERIC, "degenerate knot cloud (zero range)");
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    if (xy[a] != xy[b]) return xy[a] < xy[b];
    return xy[(size_t)n + a] < xy[(size_t)n + b];
  });
  std::vector<int> group(n);
  int np = 0;
  for (int k = 0; k < n; ++k) {
    if (k == 0 || xy[order[k]] != xy[order[k - 1]] || xy[(size_t)n + order[k]] != xy[(size_t)n + order[k - 1]]) ++np;
    group[order[k]] = np - 1;
  }
  const bool replicated = np < n;
  if (!replicated)
    for (int i = 0; i < n; ++i) group[i] = i;     // keep the caller's order when locations are distinct
  if (np <= 3) throw Error(MB_E_NUMERIC, "need more than 3 unique locations");
  std::vector<double> kx(np), ky(np), wM(np, 0.0);
  std::vector<std::vector<double>> yM(L, std::vector<double>(np, 0.0));
  std::vector<double> pure_ss(L, 0.0);
  for (int i = 0; i < n; ++i) {
    const int gidx = group[i];
    kx[gidx] = xy[i];
    ky[gidx] = xy[(size_t)n + i];
    wM[gidx] += 1.0;
    for (int r = 0; r < L; ++r) yM[r][gidx] += y[(size_t)r * n + i];
  }
  for (int r = 0; r < L; ++r) {
    for (int k = 0; k < np; ++k) yM[r][k] /= wM[k];
    if (replicated)
      for (int i = 0; i < n; ++i) {
        const double e = y[(size_t)r * n + i] - yM[r][group[i]];
        pure_ss[r] += e * e;
      }
  }
  std::vector<double> sx(np), sy(np), w2(np);
  for (int k = 0; k < np; ++k) {
    sx[k] = (kx[k] - cmin[0]) / scale[0];
    sy[k] = (ky[k] - cmin[1]) / scale[1];
    w2[k] = std::sqrt(wM[k]);
  }
  // ---- qr(sqrt(w) T) ------------------------------------------------------------------------------
  std::vector<double> T[3] = {std::vector<double>(np), std::vector<double>(np), std::vector<double>(np)};
  for (int k = 0; k < np; ++k) { T[0][k] = w2[k]; T[1][k] = w2[k] * sx[k]; T[2][k] = w2[k] * sy[k]; }
  QRT qr(T, np);
  const int m = np - 3;
  // ---- device: K, projection ------------------------------------------------------------------------
  Arena& ar = ctx->arena;   // no cudaMalloc / cudaFree here: the fit overlaps the ensemble kernels of another stream
  ABuf<double> d_sx(ar), d_sy(ar), d_w2(ar), d_K(ar, (size_t)np * np), d_v(ar, np), d_p(ar, np), d_q(ar, np),
      d_part(ar, (size_t)kGemvChunks * np);
  d_sx.upload(sx, st);
  d_sy.upload(sy, st);
  d_w2.upload(w2, st);
  const dim3 g2((np + 31) / 32, (np + 31) / 32);
  MB_LAUNCH(ctx, "k_assemble", st) k_assemble<<<g2, 256, 0, st>>>(d_sx.p, d_sy.p, d_w2.p, np, d_K.p, np);
  for (int j = 0; j < 3; ++j) {
    d_v.upload(qr.v[j], st);
    gemv_n(ctx, d_K.p, np, np, np, d_v.p, d_p.p, d_part.p, st);
    MB_LAUNCH(ctx, "k_make_q", st) k_make_q<<<1, 1024, 0, st>>>(d_v.p, d_p.p, np, d_q.p);
    MB_LAUNCH(ctx, "k_rank2", st) k_rank2<<<g2, 256, 0, st>>>(d_K.p, np, np, d_v.p, d_q.p);
    MB_CUDA(cudaStreamSynchronize(st));   // d_v is re-uploaded from a host vector next iteration
  }
  MB_CUDA(cudaGetLastError());
  // M = Q2' W^1/2 K W^1/2 Q2 = rows/cols 3.. of the reflected matrix, compacted to a 16-byte aligned m x m block
  ABuf<double> d_M(ar, (size_t)m * m);
  MB_CUDA(cudaMemcpy2DAsync(d_M.p, sizeof(double) * m, d_K.p + (size_t)3 * np + 3, sizeof(double) * np,
                            sizeof(double) * m, m, cudaMemcpyDeviceToDevice, st));
  double* M = d_M.p;

  // z_r = Q2' sqrt(w) yM_r
  std::vector<std::vector<double>> z(L);
  for (int r = 0; r < L; ++r) {
    std::vector<double> b(np);
    for (int k = 0; k < np; ++k) b[k] = w2[k] * yM[r][k];
    qr.qty(b);
    z[r].assign(b.begin() + 3, b.end());
  }

  std::vector<double> lam(L, lambda), edf(L, -1.0), gcvv(L, -1.0);
  std::vector<std::vector<double>> beta(L, std::vector<double>(m));
  std::vector<double> eta_desc, tri_diag, tri_off;
  std::vector<std::vector<double>> zhat(L);

  if (lambda < 0) {
    MB_REQUIRE(L <= 32, "GCV fit: at most 32 responses per call");
    std::vector<double> eta;            // ascending eigenvalues of M
    std::vector<double> D(np, 0.0);
    const int nstep = 200;
    auto make_grid = [&]() {
      if (!(eta[0] > 0)) throw Error(MB_E_NUMERIC, "Q2'KQ2 is not positive definite (smallest eigenvalue <= 0)");
      // D = c(0,0,0, 1/eta) in R's decreasing-eigenvalue order
      for (int k = 0; k < m; ++k) D[3 + k] = 1.0 / eta[m - 1 - k];
      eta_desc.assign(eta.rbegin(), eta.rend());
      // lambda grid: device kernel when D fits in shared memory, host loop otherwise
      std::vector<double> grid;
      if ((size_t)np * sizeof(double) <= 200 * 1024) {
        ABuf<double> d_D(ar), d_grid(ar, nstep);
        ABuf<int> d_err(ar, 1);
        d_D.upload(D, st);
        MB_CUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int), st));
        static thread_local bool attr = false;
        if (!attr) {
          MB_CUDA(cudaFuncSetAttribute(k_gcv_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
          attr = true;
        }
        MB_LAUNCH(ctx, "k_gcv_grid", st) k_gcv_grid<<<nstep, kGcvThreads, (size_t)np * sizeof(double), st>>>(d_D.p, np, nstep, d_grid.p, d_err.p);
        MB_CUDA(cudaGetLastError());
        grid.resize(nstep);
        int gerr = 0;
        MB_CUDA(cudaMemcpyAsync(grid.data(), d_grid.p, sizeof(double) * nstep, cudaMemcpyDeviceToHost, st));
        MB_CUDA(cudaMemcpyAsync(&gerr, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        MB_CUDA(cudaStreamSynchronize(st));
        if (gerr) throw Error(MB_E_NUMERIC, "bisection.search: f1 must be < f2");
        std::sort(grid.begin(), grid.end());
      } else {
        grid = gcv::lambda_grid(D);
      }
      return grid;
    };
    if (ctx->eigen_impl != 1) {
      // ---- in-house: M = Q T Q' on a work copy (M itself is kept for the Cholesky below) ----------------
      const int ldT = (m + 63) / 64 * 64;      // zero-padded to whole 64 x 64 tiles (no bounds predicates in k_sytrd)
      ABuf<double> d_T(ar, (size_t)ldT * ldT), d_z(ar, (size_t)m * L);
      MB_CUDA(cudaMemsetAsync(d_T.p, 0, sizeof(double) * (size_t)ldT * ldT, st));
      MB_CUDA(cudaMemcpy2DAsync(d_T.p, sizeof(double) * ldT, M, sizeof(double) * m, sizeof(double) * m, m,
                                cudaMemcpyDeviceToDevice, st));
      for (int r = 0; r < L; ++r)
        MB_CUDA(cudaMemcpyAsync(d_z.p + (size_t)r * m, z[r].data(), sizeof(double) * m, cudaMemcpyHostToDevice, st));
      std::vector<double> tdiag, toff;
      sym_tridiag_eig(ctx, d_T.p, ldT, m, d_z.p, L, tdiag, toff, eta, st);
      std::vector<double> zh((size_t)m * L);
      MB_CUDA(cudaMemcpyAsync(zh.data(), d_z.p, sizeof(double) * m * L, cudaMemcpyDeviceToHost, st));
      MB_CUDA(cudaStreamSynchronize(st));
      const std::vector<double> grid = make_grid();
      tri_diag = tdiag; tri_off = toff;
      for (int r = 0; r < L; ++r) {
        zhat[r].assign(zh.begin() + (size_t)r * m, zh.begin() + (size_t)(r + 1) * m);
        gcv::RssTridiag rss{tdiag, toff, zhat[r], {}, {}};
        lam[r] = gcv::search(D, rss, grid, n, pure_ss[r], &gcvv[r]);
        edf[r] = gcv::tr_a(lam[r], D);
      }
      // ---- Krig.coef at the selected lambda: (M + lambda I) beta = z by the tensor-core Cholesky ----------
      ABuf<double> d_B(ar, m), d_tmp(ar, m);
      double eta_max = 0.0, eta_min = 0.0;
      if (!eta.empty()) { eta_max = *std::max_element(eta.begin(), eta.end()); eta_min = *std::min_element(eta.begin(), eta.end()); }
      for (int r = 0; r < L; ++r) {
        // From the band form of the two-stage reduction (no dense factorisation, two launches instead of ~300) when the system is
        // well conditioned: B = Q1'MQ1 carries the rounding of stage 1 (~1e-16 |M|), so the band solve is good to
        // cond(M + lambda I) * 1e-16 - the known spectrum gives cond exactly.  Near-interpolating fits (GCV minimum at the small
        // end of fields' grid, cond >= 1e8) keep the dense Cholesky of M + lambda I itself, as do non-positive band pivots.
        const double cond = (eta_max + lam[r]) / std::max(eta_min + lam[r], 1e-300);
        if (ctx->coef_impl != 2 && (cond <= 1e8 || ctx->coef_impl == 1) && band_coefficients(ctx, lam[r], r, d_B.p, st)) {
          MB_CUDA(cudaMemcpyAsync(beta[r].data(), d_B.p, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
          MB_CUDA(cudaStreamSynchronize(st));
          continue;
        }
        MB_CUDA(cudaMemcpyAsync(d_T.p, M, sizeof(double) * (size_t)m * m, cudaMemcpyDeviceToDevice, st));
        MB_LAUNCH(ctx, "k_add_diag", st) k_add_diag<<<(m + 255) / 256, 256, 0, st>>>(d_T.p, m, m, lam[r]);
        const double* d_inv = cholesky_lower(ctx, d_T.p, m, m, st);
        MB_CUDA(cudaMemcpyAsync(d_B.p, z[r].data(), sizeof(double) * m, cudaMemcpyHostToDevice, st));
        cholesky_solve(ctx, d_T.p, m, m, d_inv, d_B.p, d_tmp.p, m, 1, st);
        MB_CUDA(cudaMemcpyAsync(beta[r].data(), d_B.p, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
        MB_CUDA(cudaStreamSynchronize(st));
      }
    } else {
      // ---- validation path: eigen(M) by cuSOLVER Dsyevd (library call) ---------------------------------
      FitLibs& lb = libs(ctx);
      ABuf<double> d_eta(ar, m);
      ABuf<int> d_info(ar, 1);
      int lwork = 0;
      if (cusolver().DsyevdBufferSize(lb.solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, m, M, m, d_eta.p,
                                      &lwork) != CUSOLVER_STATUS_SUCCESS)
        throw Error(MB_E_CUDA, "cusolverDnDsyevd_bufferSize failed");
      ABuf<double> d_work(ar, (size_t)lwork);
      cusolverStatus_t cs = CUSOLVER_STATUS_SUCCESS;
      MB_LAUNCH(ctx, "cusolverDnDsyevd", st)
        cs = cusolver().Dsyevd(lb.solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, m, M, m, d_eta.p, d_work.p,
                              lwork, d_info.p);
      if (cs != CUSOLVER_STATUS_SUCCESS) throw Error(MB_E_CUDA, "cusolverDnDsyevd failed");
      int info = 0;
      eta.resize(m);
      MB_CUDA(cudaMemcpyAsync(&info, d_info.p, sizeof(int), cudaMemcpyDeviceToHost, st));
      MB_CUDA(cudaMemcpyAsync(eta.data(), d_eta.p, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
      MB_CUDA(cudaStreamSynchronize(st));
      if (info != 0) throw Error(MB_E_NUMERIC, "symmetric eigensolver did not converge (info=" + std::to_string(info) + ")");
      const std::vector<double> grid = make_grid();
      ABuf<double> d_z(ar, m), d_u(ar, m), d_g(ar, m), d_beta(ar, m);
      for (int r = 0; r < L; ++r) {
        d_z.upload(z[r], st);
        MB_LAUNCH(ctx, "k_gemv_t", st) k_gemv_t<<<(m + 7) / 8, 256, 0, st>>>(M, m, m, m, d_z.p, d_u.p);   // u = V' z (ascending order)
        std::vector<double> u_asc(m);
        MB_CUDA(cudaMemcpyAsync(u_asc.data(), d_u.p, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
        MB_CUDA(cudaStreamSynchronize(st));
        std::vector<double> u(np, 0.0);
        for (int k = 0; k < m; ++k) u[3 + k] = u_asc[m - 1 - k];
        gcv::RssEigen rss{D, u};
        lam[r] = gcv::search(D, rss, grid, n, pure_ss[r], &gcvv[r]);
        edf[r] = gcv::tr_a(lam[r], D);
        std::vector<double> gvec(m);
        for (int k = 0; k < m; ++k) gvec[k] = u_asc[k] / (eta[k] + lam[r]);
        d_g.upload(gvec, st);
        gemv_n(ctx, M, m, m, m, d_g.p, d_beta.p, d_part.p, st);              // beta = V diag(1/(eta+lambda)) u
        MB_CUDA(cudaMemcpyAsync(beta[r].data(), d_beta.p, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
        MB_CUDA(cudaStreamSynchronize(st));
      }
    }
  } else {
    // ---- fixed lambda: (M + lambda I) beta = z by tensor-core Cholesky ------------------------------
    MB_LAUNCH(ctx, "k_add_diag", st) k_add_diag<<<(m + 255) / 256, 256, 0, st>>>(M, m, m, lambda);
    const double* d_inv = cholesky_lower(ctx, M, m, m, st);
    ABuf<double> d_B(ar, (size_t)m * L);
    for (int r = 0; r < L; ++r)
      MB_CUDA(cudaMemcpyAsync(d_B.p + (size_t)r * m, z[r].data(), sizeof(double) * m, cudaMemcpyHostToDevice, st));
    ABuf<double> d_tmp(ar, (size_t)m * L);
    cholesky_solve(ctx, M, m, m, d_inv, d_B.p, d_tmp.p, m, L, st);
    for (int r = 0; r < L; ++r)
      MB_CUDA(cudaMemcpyAsync(beta[r].data(), d_B.p + (size_t)r * m, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    MB_CUDA(cudaGetLastError());
  }

  // ---- Krig.coef: c = sqrt(w) Q (0; beta), d = qr.coef(sqrt(w) (yM - K c)) -----------------------------
  std::vector<std::unique_ptr<mb_spline>> res(L);
  for (int r = 0; r < L; ++r) {
    std::vector<double> cfull(np, 0.0);
    std::copy(beta[r].begin(), beta[r].end(), cfull.begin() + 3);
    qr.qy(cfull);
    for (int k = 0; k < np; ++k) cfull[k] *= w2[k];
    auto s = std::make_unique<mb_spline>();
    s->np = np;
    s->kx = kx; s->ky = ky; s->sx = sx; s->sy = sy;
    s->c = cfull;
    s->center[0] = cmin[0]; s->center[1] = cmin[1];
    s->scale[0] = scale[0]; s->scale[1] = scale[1];
    s->d[0] = s->d[1] = s->d[2] = 0.0;
    s->lambda = lam[r]; s->eff_df = edf[r]; s->gcv = gcvv[r];
    if (lambda < 0) { s->eta = eta_desc; s->tri_diag = tri_diag; s->tri_off = tri_off; s->zhat = zhat[r]; }
    // K c through the evaluation kernel with d = 0 (K was overwritten by the projection)
    s->ctx = ctx;
    s->device = ctx->device;
    s->d_sx.upload(s->sx, st); s->d_sy.upload(s->sy, st); s->d_c.upload(s->c, st);
    ABuf<double> d_kx(ar), d_ky(ar), d_kc(ar, np);
    d_kx.upload(kx, st); d_ky.upload(ky, st);
    tps_predict_points_dev(ctx, s.get(), d_kx.p, d_ky.p, np, d_kc.p, st);
    std::vector<double> Kc(np);
    MB_CUDA(cudaMemcpyAsync(Kc.data(), d_kc.p, sizeof(double) * np, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    std::vector<double> rhs(np);
    for (int k = 0; k < np; ++k) rhs[k] = w2[k] * (yM[r][k] - Kc[k]);
    qr.coef(rhs, s->d);
    // f(x_k) = yM_k - lambda c_k / w_k at the knots (first block row of the Krig system): no evaluation needed
    double fs = 0.0;
    for (int k = 0; k < np; ++k) fs = std::max(fs, std::fabs(yM[r][k] - lam[r] * s->c[k] / wM[k]));
    spline_finalize(ctx, s.get(), fs);
    res[r] = std::move(s);
  }
  for (int r = 0; r < L; ++r) out[r] = res[r].release();
}

}  // namespace mb
