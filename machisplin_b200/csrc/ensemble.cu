// Per-cell evaluation of the weighted model ensemble - replaces terra::predict(rast_stack, model_k)
// for the kept models and the accumulation / normalisation of V73:468-475, 497-499, 521-523, 543-545,
// 582-584, 604-606, 619, plus the NA-propagating sum with the TPS surface of part 5 (V73:906-907).
//
//   pred = sum_{k kept} round(w_k, 2) f_k(cov_1..cov_C, LONG, LAT) / sum_all w_k ;  final = pred + TPS
//
// Feature order is xnam (V73:194): the C covariate planes, then LONG, LAT (cell-centre coordinates,
// V73:127-133).  Kernels:
//   k_ens_trees   randomForest + gbm: trees re-laid out breadth-first into 8-byte nodes
//                 {float threshold | leaf value, feature/leaf/child}; children adjacent.  Thresholds are
//                 converted so that every comparison is bit-identical to the double comparison of the
//                 reference on float-valued rasters: x <= t  <=>  x <= fl_down(t);  x < t  <=>
//                 x <= pred(fl_up(t)); splits on LONG / LAT become splits on the integer column / row.
//                 One CTA owns a 32 x (8 R)-cell tile: it reduces the per-feature [min, max] of the tile,
//                 walks every tree ONCE with that interval (a split whose threshold lies outside the
//                 interval sends all cells the same way) and keeps only the trees whose path forks inside
//                 the tile, from the forking node down; collapsed trees add one tile constant.  Rasters
//                 are spatially coherent, so >90 % of the node visits of the per-cell walk disappear while
//                 every surviving comparison is the reference's own (results are identical).
//   k_ens_svm     ksvm rbfdot: exp2(a + b_i + x . sv'_i), float32 dot product, float64 accumulation.
//   smooth models gam + nnet + earth in float64: folded into k_ens_svm (whose FP64 pipe is idle) when ksvm is
//                 kept, else k_ens_smooth.  The last kernel of the chain also encodes the NA rule: the
//                 accumulator holds NaN where any covariate is NA.
//   k_ens_final   acc / total weight + precomputed TPS surface (the fused path is k_leaf_stream, tps_eval.cu).
// The accumulator uses the padded layout of AccFuse (common.cuh): row stride acc_stride(w).
#include "common.cuh"
#include "internal.h"
#include "ens_device.cuh"
#include "async_copy.cuh"

#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <exception>
#include <queue>
#include <thread>

struct PackedForest {
  int ntrees = 0;
  double offset = 0.0;               // subtracted from leaf values before the float conversion
};

struct mb_ensemble {
  mb_ctx* ctx = nullptr;             // not dereferenced by ensemble_free (a handle may outlive its context)
  int device = 0;
  mb_grid g{};
  int P = 0, C = 0;
  double w[6] = {0, 0, 0, 0, 0, 0};  // b g n m r v
  double w_total = 1.0;
  bool has[6] = {false, false, false, false, false, false};
  bool only_gbm = false;             // gbm alone returns numbers on NA cells (MissingNode)
  // smooth models (float64)
  mb::DevBuf<double> gam, nn;
  int nn_H = 0;
  double nn_max2 = 1, nn_min = 0;
  int mars_T = 0;
  mb::DevBuf<double> mars_coef, mars_cut;   // per factor
  mb::DevBuf<int> mars_off, mars_var, mars_dir;
  // svm
  int svm_S = 0;
  int svm_pairs = 0;                 // ceil(S / 2)
  mb::DevBuf<float4> svm_svp;        // [pair][NQ]: (sv_2j[2q], sv_2j+1[2q], sv_2j[2q+1], sv_2j+1[2q+1]) x 2 sigma log2e
  mb::DevBuf<float4> svm_bap;        // [pair]: (b_2j, b_2j+1, alpha_2j, alpha_2j+1), b = -sigma |sv|^2 log2e
  // tensor-core variant (k_ens_svm_mma, P <= 8, the default there; not built with "svm_impl" = 2):
  int svm_oct = 0;                   // ceil(S / 8)
  mb::DevBuf<float4> svm_bq;         // [octet][lane]: B fragments of mma.m16n8k8.tf32, (hi f=tig, hi f=tig+4, lo f=tig, lo f=tig+4) of SV 8o + (lane >> 2)
  mb::DevBuf<float4> svm_baq;        // svm_bap padded with zeros to 4 pairs per octet
  mb::DevBuf<uint4> svm_bh;          // [octet][lane]: B fragments of mma.m16n8k16.f16 (svm_impl 3), half2 pairs of features 2 tig, 2 tig + 1: (hi, hi, lo, lo)
  double svm_hs = 1.0;               // power of two that moves part of 2 sigma log2(e) from the support vectors to x in that layout
  mb::DevBuf<double> svm_xc, svm_xis; // centre, 1/scale
  double svm_bias = 0, svm_sigma = 0, svm_yc = 0, svm_ys = 1;
  // trees
  PackedForest rf, gbm;
  // both forests in one node array for the tile-pruned kernel: rf trees first, then gbm;
  // node zero_leaf is a leaf holding 0.0f (padding entry of the residual lists)
  mb::DevBuf<int2> forest_nodes;
  mb::DevBuf<int> forest_roots;
  int forest_zero_leaf = 0;
  // reference-layout copies for the point path (float64 thresholds, arbitrary coordinates)
  mb::DevBuf<int> rfp_left, rfp_right, rfp_var;
  mb::DevBuf<signed char> rfp_status;
  mb::DevBuf<double> rfp_split, rfp_pred;
  int rf_ntree = 0, rf_nrnodes = 0;
  mb::DevBuf<int> gbp_off, gbp_var, gbp_left, gbp_right, gbp_miss;
  mb::DevBuf<double> gbp_code;
  int gb_ntrees = 0;
  double gb_initF = 0;
  mb::DevBuf<double> svp_sv, svp_alpha;   // float64 copies for the point path
};

namespace mb {

enum { MB_B = 0, MB_G = 1, MB_N = 2, MB_M = 3, MB_R = 4, MB_V = 5 };
static int letter_index(char c) {
  switch (c) {
    case 'b': return MB_B; case 'g': return MB_G; case 'n': return MB_N;
    case 'm': return MB_M; case 'r': return MB_R; case 'v': return MB_V;
  }
  throw Error(MB_E_ARG, std::string("unknown model letter '") + c + "' (expected b g n m r v)");
}

// ---------------------------------------------------------------------------------------------
// host: threshold conversion and tree re-layout
// ---------------------------------------------------------------------------------------------
static float fl_down(double t) {     // largest float <= t
  float f = (float)t;
  if ((double)f > t) f = std::nextafterf(f, -INFINITY);
  return f;
}
static float fl_up(double t) {       // smallest float >= t
  float f = (float)t;
  if ((double)f < t) f = std::nextafterf(f, INFINITY);
  return f;
}
// largest index i in [0, n) with coord(i) <= t (strict: < t), coord monotone; -1 if none.
template <class F>
static int last_index_le(int n, F coord, double t, bool strict, bool increasing) {
  // returns count of indices satisfying the predicate from the "low-coordinate" end
  int lo = 0, hi = n;   // predicate true on a prefix (increasing) or suffix (decreasing)
  auto pred = [&](int i) { const double v = coord(i); return strict ? v < t : v <= t; };
  if (increasing) {
    while (lo < hi) { int mid = (lo + hi) / 2; if (pred(mid)) lo = mid + 1; else hi = mid; }
    return lo - 1;      // indices [0, lo) satisfy
  }
  while (lo < hi) { int mid = (lo + hi) / 2; if (pred(mid)) hi = mid; else lo = mid + 1; }
  return lo;            // indices [lo, n) satisfy
}

constexpr int kMetaLeaf = 1 << 4;
struct HostNode { float val; int meta; };

// Builds the "go to first child iff feature <= thr" node for a reference split.
//   strict = false: reference goes LEFT iff x <= t (randomForest);  strict = true: LEFT iff x < t (gbm).
// Returns the threshold and whether the children must be swapped (LAT decreases with the row index).
static void convert_split(const mb_grid& g, int C, int var, double t, bool strict, float* thr, bool* swap) {
  *swap = false;
  if (var < C) {
    *thr = strict ? std::nextafterf(fl_up(t), -INFINITY) : fl_down(t);
    return;
  }
  const double rx = (g.xmax - g.xmin) / g.ncol, ry = (g.ymax - g.ymin) / g.nrow;
  if (var == C) {       // LONG = xmin + (col + 0.5) rx, increasing in col
    const int c = last_index_le(g.ncol, [&](int i) { return g.xmin + (i + 0.5) * rx; }, t, strict, true);
    *thr = (float)c;    // col <= c
  } else {              // LAT = ymax - (row + 0.5) ry, decreasing in row: LEFT iff row >= r0
    const int r0 = last_index_le(g.nrow, [&](int i) { return g.ymax - (i + 0.5) * ry; }, t, strict, false);
    *thr = (float)(r0 - 1);   // row <= r0 - 1  ->  RIGHT
    *swap = true;
  }
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------

// ---------------------------------------------------------------------------------------------
// k_ens_trees_plain: one thread per cell, every tree walked from its root.  Only used when gbm is the
// single kept model (NA cells then follow MissingNode, V73:493-501); all other cases run k_ens_trees.
// ---------------------------------------------------------------------------------------------
constexpr int kTreeThreads = 256;

template <bool kMissing>
__device__ __forceinline__ float walk_tree(const int2* __restrict__ nodes, int idx, const float* s_feat) {
  for (;;) {
    const int2 nd = __ldg(&nodes[idx]);
    if (nd.y & kMetaLeaf) return __int_as_float(nd.x);
    const float x = s_feat[(nd.y & 15) * kTreeThreads];
    int next = (nd.y >> 5) + (x <= __int_as_float(nd.x) ? 0 : 1);
    if (kMissing) { if (x != x) next = (nd.y >> 5) + 2; }
    idx = next;
  }
}

__global__ void __launch_bounds__(kTreeThreads) k_ens_trees_plain(
    const float* __restrict__ cov, int C, int64_t plane, EnsGeom eg, mb_window w,
    const int2* __restrict__ rf_nodes, const int* __restrict__ rf_root, int rf_n, double rf_w, double rf_off,
    const int2* __restrict__ gb_nodes, const int* __restrict__ gb_root, int gb_n, double gb_w, double gb_init,
    int gbm_missing, int64_t acc_stride, double* __restrict__ acc) {
  __shared__ float s_feat[16 * kTreeThreads];
  const int wc = w.c1 - w.c0;
  const int64_t cell = (int64_t)blockIdx.x * kTreeThreads + threadIdx.x;
  const int64_t ncell = (int64_t)(w.r1 - w.r0) * wc;
  const bool live = cell < ncell;
  const int row = w.r0 + (int)(live ? cell / wc : 0);
  const int col = w.c0 + (int)(live ? cell % wc : 0);
  bool anynan = false;
  for (int f = 0; f < C; ++f) {
    const float v = live ? cov[f * plane + (int64_t)row * eg.ncol + col] : 0.f;
    anynan |= (v != v);
    s_feat[f * kTreeThreads + threadIdx.x] = v;
  }
  s_feat[C * kTreeThreads + threadIdx.x] = (float)col;
  s_feat[(C + 1) * kTreeThreads + threadIdx.x] = (float)row;
  if (!live) return;
  const float* sf = s_feat + threadIdx.x;
  double out = 0.0;
  if (rf_n > 0 && !anynan) {
    double s = 0.0;
    for (int t = 0; t < rf_n; ++t) s += (double)walk_tree<false>(rf_nodes, __ldg(&rf_root[t]), sf);
    out += rf_w * (s / rf_n + rf_off);
  }
  if (gb_n > 0 && (!anynan || gbm_missing)) {
    double s = 0.0;
    if (gbm_missing) for (int t = 0; t < gb_n; ++t) s += (double)walk_tree<true>(gb_nodes, __ldg(&gb_root[t]), sf);
    else for (int t = 0; t < gb_n; ++t) s += (double)walk_tree<false>(gb_nodes, __ldg(&gb_root[t]), sf);
    out += gb_w * (s + gb_init);
  }
  acc[(int64_t)(row - w.r0) * acc_stride + (col - w.c0)] += out;
}

// ---------------------------------------------------------------------------------------------
// k_ens_trees: tile-pruned forest evaluation (see the file header).
//   phase 1  features of the tile -> shared memory, per-feature [lo, hi] over the evaluated cells
//   phase 2  warp w walks the 32-tree blocks w, w + 8, ... with the interval: collapsed trees add to a
//            per-thread constant, forking trees append (fork node | kind << 30) to the warp's list
//            (ballot-ordered -> the summation order, hence the result, is deterministic)
//   phase 3  every cell walks the surviving subtrees, kTreeIlp at a time
// ---------------------------------------------------------------------------------------------
constexpr int kTreeChunk = 2048;     // trees pruned per pass (bounds the residual lists)
constexpr int kTreeSeg = kTreeChunk / 8;
// tuning constants (-D overrides: tools/build_variants.sh builds A/B copies of the library)
#ifndef MB_TREE_ILP
#define MB_TREE_ILP 4
#endif
#ifndef MB_TREE_CTAS
#define MB_TREE_CTAS 5
#endif
#ifndef MB_PRUNE_ILP
#define MB_PRUNE_ILP 3
#endif
#ifndef MB_WPRUNE_ILP
#define MB_WPRUNE_ILP 2
#endif
constexpr int kTreeIlp = MB_TREE_ILP;        // subtrees a cell walks at once (independent dependent-load chains: the walk is bound by L2 latency)
constexpr int kPruneIlp = MB_PRUNE_ILP;      // trees a lane prunes at once in phase 2
constexpr int kWarpPruneIlp = MB_WPRUNE_ILP; // fork-list entries a lane prunes at once in phase 2b
constexpr int kKindShift = 30;
constexpr int kChainDone = (int)0x80000000, kChainFork = 1 << 29, kChainNode = (1 << 26) - 1;   // state word of a pruning chain

// Two-level variant (the default): the CTA tile of 32 x 8 cells is cut into eight 8 x 4 blocks, one per warp.  After the CTA-level
// prune (phase 2, as above) every warp prunes the CTA's fork list once more with the feature intervals of ITS 32 cells, from the
// fork nodes down (phase 2b): most surviving trees collapse to a warp constant there, and the cells walk only what forks inside
// their own 8 x 4 block.  Measured on config 3 (tools/forest_prune_sim.c): 41 + 698 node visits per cell with one level,
// 41 + 17 + 186 with two.  Every comparison is still the reference's own; the summation order is fixed (ballot order at both
// levels), so results are deterministic.  On rough rasters (the reference's slope / TWI) most trees fork at every level and the
// kernel degrades to the plain per-cell walk plus two cheap interval passes.
constexpr int kWarpList = 256;       // fork-list entries a warp prunes per pass (bounds its residual list)

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

template <int R>
__global__ void __launch_bounds__(kTreeThreads, MB_TREE_CTAS) k_ens_trees(
    const float* __restrict__ cov, int C, int64_t plane, EnsGeom eg, mb_window w,
    const int2* __restrict__ nodes, const int* __restrict__ roots, int n_rf, int n_gb, int zero_leaf,
    double rf_scale, double gb_scale, double base, int accumulate, int64_t acc_stride, double* __restrict__ acc, int levels) {
  static_assert(R == 1, "the two-level kernel owns one cell per thread");
  extern __shared__ __align__(16) unsigned char tree_smem[];
  float* s_feat = reinterpret_cast<float*>(tree_smem);                    // [(C + 2)][256]
  int* s_list = reinterpret_cast<int*>(s_feat + (C + 2) * kTreeThreads);  // [8][kTreeSeg + kTreeIlp]   CTA-level fork list
  int* s_wlist = s_list + 8 * (kTreeSeg + kTreeIlp);                      // [8][kWarpList + kTreeIlp]  warp-level fork lists
  __shared__ float2 s_wiv[16][8], s_iv[16];        // (lo, hi) of feature f over warp q's block / over the tile: one LDS.64 per split
  __shared__ double s_wsum[8];
  __shared__ int s_cnt[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // warp = 8 x 4 block (bx, by) of the tile, lane = cell (dx, dy) of the block
  const int col = w.c0 + blockIdx.x * 32 + 8 * (warp & 3) + (lane & 7);
  const int row = w.r0 + blockIdx.y * 8 + 4 * (warp >> 2) + (lane >> 3);
  const int64_t wc = acc_stride;
  const bool inside = col < w.c1 && row < w.r1;
  // ---- phase 1: features -> shared memory, intervals of the warp's block and of the tile -------------------
  bool eval = inside;
  for (int f = 0; f < C; ++f) {
    float v = 0.f, lo = INFINITY, hi = -INFINITY;
    if (inside) {
      v = __ldg(&cov[f * plane + (int64_t)row * eg.ncol + col]);
      if (v != v) eval = false;
      lo = hi = v;                    // NaN: dropped by fminf / fmaxf below
    }
    s_feat[f * kTreeThreads + tid] = v;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) s_wiv[f][warp] = make_float2(lo, hi);
  }
  s_feat[C * kTreeThreads + tid] = (float)col;
  s_feat[(C + 1) * kTreeThreads + tid] = (float)row;
  if (lane == 0) {
    const int c_lo = w.c0 + blockIdx.x * 32 + 8 * (warp & 3), r_lo = w.r0 + blockIdx.y * 8 + 4 * (warp >> 2);
    s_wiv[C][warp] = make_float2((float)c_lo, (float)min(w.c1 - 1, c_lo + 7));
    s_wiv[C + 1][warp] = make_float2((float)r_lo, (float)min(w.r1 - 1, r_lo + 3));
  }
  const int any_eval = __syncthreads_or(eval);
  if (!any_eval) {                                  // sea tile: nothing to evaluate
    if (inside && !accumulate) acc[(int64_t)(row - w.r0) * wc + (col - w.c0)] = 0.0;   // accumulate 1 / 2: nothing to add
    return;
  }
  if (tid < C + 2) {
    float lo = INFINITY, hi = -INFINITY;
#pragma unroll
    for (int q = 0; q < 8; ++q) { lo = fminf(lo, s_wiv[tid][q].x); hi = fmaxf(hi, s_wiv[tid][q].y); }
    s_iv[tid] = make_float2(lo, hi);
  }
  __syncthreads();
  const int ntrees = n_rf + n_gb;
  // 32-bit shared-memory addresses for the loads of the hot loops: through generic pointers the compiler rebuilt the shared window
  // (S2UR SR_CgaCtaId + UMOV + UIADD3 + ULEA) at every node visit - 7 of the 18 instructions of a visit
  const uint32_t sf_addr = (uint32_t)__cvta_generic_to_shared(s_feat) + 4u * tid;
  const uint32_t iv_addr = (uint32_t)__cvta_generic_to_shared(&s_iv[0]);
  const uint32_t wiv_addr = (uint32_t)__cvta_generic_to_shared(&s_wiv[0][warp]);
  double cell_sum = 0.0;
  double wsum = 0.0;                                // trees that collapse on this warp's block, walked by this lane
  int* my_list = s_list + warp * (kTreeSeg + kTreeIlp);
  int* my_wlist = s_wlist + warp * (kWarpList + kTreeIlp);
  const bool warp_eval = __any_sync(0xffffffffu, eval);
  for (int t0 = 0; t0 < ntrees; t0 += kTreeChunk) {
    // ---- phase 2: CTA-level prune, warp q owns trees [q per, (q + 1) per) of the chunk ----------------------
    // Blocks of 32 consecutive trees are dealt round-robin to the warps (block b -> warp b mod 8), so that every warp gets the
    // same mix of deep randomForest and shallow gbm trees (rf trees come first in the chunk): with contiguous shares the three
    // warps that owned the rf trees needed 40 dependent rounds while the other five waited at the barrier after 8 (14 % of
    // all stall samples, profiles/r2z_ncu_full_trees.md).  A block is homogeneous, so the side-by-side chains of a round
    // have similar depths.
    const int nchunk = min(kTreeChunk, ntrees - t0);
    const int nblk = (nchunk + 31) >> 5;
    int cnt = 0;
    double csum = 0.0;                              // trees that collapse on the whole tile, walked by this thread
    for (int b0 = warp; b0 < nblk; b0 += 8 * kPruneIlp) {
      // kPruneIlp trees per lane, walked side by side: every step is a dependent L2 access, the chains overlap.  The whole state
      // of a chain is one integer (node | fork << 29 | kind << 30 | done << 31): flag arrays made ptxas spill predicates.
      int st[kPruneIlp];
#pragma unroll
      for (int u = 0; u < kPruneIlp; ++u) {
        const int tl = 32 * (b0 + 8 * u) + lane;     // tree of this lane within the chunk
        st[u] = kChainDone;
        if (tl < nchunk) st[u] = __ldg(&roots[t0 + tl]) | (((t0 + tl) >= n_rf ? 1 : 0) << kKindShift);
      }
      for (;;) {
        int live = -1;
        int2 nd[kPruneIlp];
#pragma unroll
        for (int u = 0; u < kPruneIlp; ++u) {
          if (st[u] >= 0) nd[u] = __ldg(&nodes[st[u] & kChainNode]);
          live &= st[u];
        }
        if (live < 0) break;                         // every chain done
#pragma unroll
        for (int u = 0; u < kPruneIlp; ++u) {
          if (st[u] < 0) continue;
          if (nd[u].y & kMetaLeaf) {                 // added in program order: fixed by the data, not by timing
            csum += ((st[u] >> kKindShift) ? gb_scale : rf_scale) * (double)__int_as_float(nd[u].x);
            st[u] = kChainDone;
            continue;
          }
          const int f = nd[u].y & 15;
          const float thr = __int_as_float(nd[u].x);
          const int kindbit = st[u] & (1 << kKindShift);
          const float2 iv = lds_f32x2(iv_addr + 8u * f);
          if (iv.y <= thr) st[u] = (nd[u].y >> 5) | kindbit;
          else if (iv.x > thr) st[u] = ((nd[u].y >> 5) + 1) | kindbit;
          else st[u] |= kChainDone | kChainFork;
        }
      }
#pragma unroll
      for (int u = 0; u < kPruneIlp; ++u) {          // fixed order: the list does not depend on who finished first
        const bool fork = (st[u] & kChainFork) != 0;
        const unsigned m = __ballot_sync(0xffffffffu, fork);
        if (fork) my_list[cnt + __popc(m & ((1u << lane) - 1u))] = st[u] & (kChainNode | (1 << kKindShift));
        cnt += __popc(m);
      }
    }
    // the tile constant of this chunk goes to shared memory now (fixed-order reduction), so that nothing of phase 2 stays live
#pragma unroll
    for (int o = 16; o; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
    if (lane == 0) { s_cnt[warp] = cnt; s_wsum[warp] = (t0 ? s_wsum[warp] : 0.0) + csum; }
    __syncthreads();
    // ---- phase 2b + 3, warp-local: prune the CTA list with the block's intervals, then walk what is left -----
    if (warp_eval) {
      // The survivors of all eight CTA lists collect in ONE warp list that is walked when it is nearly full and at the end:
      // walked per CTA list (r2) most batches of kTreeIlp subtrees were padding (about 5 survivors per list on config 3).
      int wcnt = 0;
      auto walk_list = [&]() {
        if (lane < kTreeIlp) my_wlist[wcnt + lane] = zero_leaf;   // pad to a whole ILP batch
        __syncwarp();
        if (eval) {
          double s = 0.0;
          for (int i = 0; i < wcnt; i += kTreeIlp) {
            int e[kTreeIlp], idx[kTreeIlp];
            int2 nd[kTreeIlp];
#pragma unroll
            for (int u = 0; u < kTreeIlp; ++u) { e[u] = my_wlist[i + u]; idx[u] = e[u] & ((1 << kKindShift) - 1); }
            for (;;) {
              int leafs = kMetaLeaf;
#pragma unroll
              for (int u = 0; u < kTreeIlp; ++u) { nd[u] = __ldg(&nodes[idx[u]]); leafs &= nd[u].y; }
              if (leafs) break;
#pragma unroll
              for (int u = 0; u < kTreeIlp; ++u) {     // branch-free: a chain that sits on a leaf stays there
                const float x = lds_f32(sf_addr + ((nd[u].y & 15) << 10));
                const int nxt = (nd[u].y >> 5) + (x <= __int_as_float(nd[u].x) ? 0 : 1);
                idx[u] = (nd[u].y & kMetaLeaf) ? idx[u] : nxt;
              }
            }
#pragma unroll
            for (int u = 0; u < kTreeIlp; ++u)
              s += ((e[u] >> kKindShift) ? gb_scale : rf_scale) * (double)__int_as_float(nd[u].x);
          }
          cell_sum += s;
        }
        __syncwarp();                                // the warp list is rewritten by the next pass
        wcnt = 0;
      };
      for (int q = 0; q < 8; ++q) {
        const int n = s_cnt[q];
        const int* lst = s_list + q * (kTreeSeg + kTreeIlp);
        for (int i = 0; i < n; i += 32 * kWarpPruneIlp) {
          // kWarpPruneIlp entries per lane side by side (every step is a dependent L2 access); the list keeps the entry order
          int st[kWarpPruneIlp];                     // chain state as in phase 2; the list entries are node | kind << 30
#pragma unroll
          for (int u = 0; u < kWarpPruneIlp; ++u) {
            const int j = i + 32 * u + lane;
            st[u] = kChainDone;
            if (j < n) st[u] = lst[j] | (levels < 2 ? kChainDone | kChainFork : 0);   // "tree_levels" = 1 (A/B): every entry stays
          }
          for (;;) {
            int live = -1;
            int2 nd[kWarpPruneIlp];
#pragma unroll
            for (int u = 0; u < kWarpPruneIlp; ++u) {
              if (st[u] >= 0) nd[u] = __ldg(&nodes[st[u] & kChainNode]);
              live &= st[u];
            }
            if (live < 0) break;
#pragma unroll
            for (int u = 0; u < kWarpPruneIlp; ++u) {
              if (st[u] < 0) continue;
              if (nd[u].y & kMetaLeaf) {
                wsum += ((st[u] >> kKindShift) ? gb_scale : rf_scale) * (double)__int_as_float(nd[u].x);
                st[u] = kChainDone;
                continue;
              }
              const int f = nd[u].y & 15;
              const float thr = __int_as_float(nd[u].x);
              const int kindbit = st[u] & (1 << kKindShift);
              const float2 iv = lds_f32x2(wiv_addr + 64u * f);
              if (iv.y <= thr) st[u] = (nd[u].y >> 5) | kindbit;
              else if (iv.x > thr) st[u] = ((nd[u].y >> 5) + 1) | kindbit;
              else st[u] |= kChainDone | kChainFork;
            }
          }
#pragma unroll
          for (int u = 0; u < kWarpPruneIlp; ++u) {
            const bool fork = (st[u] & kChainFork) != 0;
            const unsigned m = __ballot_sync(0xffffffffu, fork);
            if (fork) my_wlist[wcnt + __popc(m & ((1u << lane) - 1u))] = st[u] & (kChainNode | (1 << kKindShift));
            wcnt += __popc(m);
          }
          if (wcnt > kWarpList - 32 * kWarpPruneIlp) walk_list();    // no room for another round of entries
        }
      }
      if (wcnt > 0) walk_list();
    }
    __syncthreads();                                 // the CTA list is rewritten by the next chunk
  }
  // ---- constants: fixed-order reductions of the per-thread sums (tile) and per-lane sums (block) ----------
#pragma unroll
  for (int o = 16; o; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
  double tile_const = base;                         // s_wsum is complete since the barrier that ended the last chunk's phase 2
#pragma unroll
  for (int q = 0; q < 8; ++q) tile_const += s_wsum[q];
  if (inside) {
    double* dst = acc + (int64_t)(row - w.r0) * wc + (col - w.c0);
    const double v = eval ? (tile_const + wsum) + cell_sum : 0.0;
    if (accumulate == 2) atomicAdd(dst, v);          // beside k_ens_svm_tma on a zeroed accumulator: two commutative adds per cell
    else *dst = accumulate ? *dst + v : v;
  }
}

// ---------------------------------------------------------------------------------------------
// k_ens_svm: sum_i alpha_i exp(-sigma |x - sv_i|^2) as sum_i alpha_i 2^(a0 + b_i + x . sv'_i).
// Support vectors are streamed through shared memory in PAIRS, feature-interleaved, so that one
// packed FFMA2 (fma.rn.f32x2, sm_100) advances the exponents of two support vectors and one LDS.128
// feeds two features of both.  Every thread owns two cells (rows r and r + 8 of a 32 x 16 tile) to reuse
// each load twice.  Issue slots per (2 SV x 2 cells): NQ + 1 LDS.128, 2 (2 NQ + 2) packed FP32, 4 MUFU
// -> the kernel is balanced between the issue port and the MUFU.EX2 pipe (16 / clk / SM on sm_100).
// ---------------------------------------------------------------------------------------------
constexpr int kSvmThreads = 256;
constexpr int kSvmPairs = 64;        // support-vector pairs per shared-memory stage

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// the float64 smooth models of one cell, kept out of line so that the register allocation of the
// support-vector loop is not affected (called once per cell before the loop)
__device__ __noinline__ double smooth_cell(const float* __restrict__ cov, int C, int64_t plane, EnsGeom eg, int row,
                                           int col, SmoothParams sp) {
  double x[16];
  for (int f = 0; f < C; ++f) x[f] = (double)__ldg(&cov[f * plane + (int64_t)row * eg.ncol + col]);
  x[C] = eg.xmin + (col + 0.5) * eg.rx;
  x[C + 1] = eg.ymax - (row + 0.5) * eg.ry;
  return smooth_models(x, C + 2, sp);
}

template <int NQ>   // NQ = ceil(P / 2) feature pairs
__global__ void __launch_bounds__(kSvmThreads, NQ <= 4 ? 5 : 3) k_ens_svm(
    const float* __restrict__ cov, int C, int P, int64_t plane, EnsGeom eg, mb_window w,
    const float4* __restrict__ svp, const float4* __restrict__ bap, int npairs,
    const double* __restrict__ xc, const double* __restrict__ xis, double sigma, double bias, double ys, double yc,
    double wv, int accumulate, SmoothParams sp, int64_t acc_stride, double* __restrict__ acc) {
  __shared__ float4 s_sv[kSvmPairs * NQ];
  __shared__ float4 s_ba[kSvmPairs];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int col = w.c0 + blockIdx.x * 32 + lane;
  const int row0 = w.r0 + blockIdx.y * 16 + warp;
  const int64_t wc = acc_stride;
  // gam / nnet / earth ride along (this kernel is MUFU / issue bound and leaves the FP64 pipe idle); evaluated
  // first, while almost nothing is live across the out-of-line call
  double smooth[2] = {0.0, 0.0};
  if (sp.gam || sp.nn || sp.mars_T > 0) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int row = row0 + 8 * c;
      if (col < w.c1 && row < w.r1) smooth[c] = smooth_cell(cov, C, plane, eg, row, col, sp);
    }
  }
  float2 xd[2][2 * NQ];      // features of both cells, each duplicated into the two halves of a pair
  float2 a0d[2];
  bool ok[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int row = row0 + 8 * c;
    const bool live = col < w.c1 && row < w.r1;
    double n2 = 0.0;
    bool anynan = false;
#pragma unroll
    for (int f = 0; f < 2 * NQ; ++f) {
      double xs = 0.0;
      if (f < P) {
        double v;
        if (f < C) v = live ? (double)__ldg(&cov[f * plane + (int64_t)row * eg.ncol + col]) : 0.0;
        else if (f == C) v = eg.xmin + (col + 0.5) * eg.rx;
        else v = eg.ymax - (row + 0.5) * eg.ry;
        anynan |= (v != v);
        xs = (v - xc[f]) * xis[f];
      }
      n2 += xs * xs;
      xd[c][f] = make_float2((float)xs, (float)xs);
    }
    const float a0 = (float)(-sigma * n2 * 1.4426950408889634);
    a0d[c] = make_float2(a0, a0);
    ok[c] = live && !anynan;
  }
  double total[2] = {0.0, 0.0};
  for (int base = 0; base < npairs; base += kSvmPairs) {
    const int n = min(kSvmPairs, npairs - base);
    __syncthreads();
    for (int i = tid; i < n * NQ; i += kSvmThreads) s_sv[i] = __ldg(&svp[(size_t)base * NQ + i]);
    for (int i = tid; i < n; i += kSvmThreads) s_ba[i] = __ldg(&bap[base + i]);
    __syncthreads();
    float2 part[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll 2
    for (int i = 0; i < n; ++i) {
      const float4 ba = s_ba[i];
      float2 e[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) e[c] = __fadd2_rn(make_float2(ba.x, ba.y), a0d[c]);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const float4 v = s_sv[i * NQ + q];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          e[c] = __ffma2_rn(xd[c][2 * q], make_float2(v.x, v.y), e[c]);
          e[c] = __ffma2_rn(xd[c][2 * q + 1], make_float2(v.z, v.w), e[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < 2; ++c)
        part[c] = __ffma2_rn(make_float2(ba.z, ba.w), make_float2(ex2_approx(e[c].x), ex2_approx(e[c].y)), part[c]);
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) total[c] += (double)part[c].x + (double)part[c].y;
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int row = row0 + 8 * c;
    if (col < w.c1 && row < w.r1) {
      double* dst = acc + (int64_t)(row - w.r0) * wc + (col - w.c0);
      // last kernel of the chain: NA rule (V73: terra::predict returns NA where any layer is NA)
      double v = __longlong_as_double(0x7ff8000000000000LL);
      if (ok[c]) v = (accumulate ? *dst : 0.0) + wv * ((total[c] - bias) * ys + yc) + smooth[c];
      *dst = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k_ens_smooth: gam / nnet / earth in float64 when no ksvm kernel carries them; closes the chain (NA rule).
// k_ens_final : acc / total weight + precomputed TPS surface (tiled mode, tps = FALSE).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ens_smooth(const float* __restrict__ cov, int C, int64_t plane, EnsGeom eg,
                                                    mb_window w, SmoothParams sp, int accumulate, int64_t acc_stride,
                                                    double* __restrict__ acc) {
  const int col = w.c0 + blockIdx.x * 32 + (threadIdx.x & 31);
  const int row = w.r0 + blockIdx.y * 8 + (threadIdx.x >> 5);
  if (col >= w.c1 || row >= w.r1) return;
  double x[16];
  bool anynan = false;
  for (int f = 0; f < C; ++f) {
    const float v = cov[f * plane + (int64_t)row * eg.ncol + col];
    anynan |= (v != v);
    x[f] = (double)v;
  }
  x[C] = eg.xmin + (col + 0.5) * eg.rx;
  x[C + 1] = eg.ymax - (row + 0.5) * eg.ry;
  double* dst = acc + (int64_t)(row - w.r0) * acc_stride + (col - w.c0);
  double s = accumulate ? *dst : 0.0;
  if (!anynan) s += smooth_models(x, C + 2, sp);
  if (anynan && !sp.only_gbm) s = __longlong_as_double(0x7ff8000000000000LL);
  *dst = s;
}

__global__ void __launch_bounds__(256) k_ens_final(mb_window w, const double* __restrict__ acc, int64_t acc_stride,
                                                   double inv_w, const double* __restrict__ tps,
                                                   double* __restrict__ out) {
  const int col = w.c0 + blockIdx.x * 32 + (threadIdx.x & 31);
  const int row = w.r0 + blockIdx.y * 8 + (threadIdx.x >> 5);
  if (col >= w.c1 || row >= w.r1) return;
  const int64_t o = (int64_t)(row - w.r0) * (w.c1 - w.c0) + (col - w.c0);
  double r = acc[(int64_t)(row - w.r0) * acc_stride + (col - w.c0)] * inv_w;
  if (tps) r += tps[o];
  out[o] = r;
}

// ---------------------------------------------------------------------------------------------
// point path (residual side, V73:477-611): float64 features given explicitly, reference tree layout
// ---------------------------------------------------------------------------------------------
struct PointModels {
  SmoothParams sp;
  int P;
  // rf
  int rf_ntree, rf_nrnodes; const int* rf_left; const int* rf_right; const signed char* rf_status;
  const int* rf_var; const double* rf_split; const double* rf_pred; double w_r;
  // gbm
  int gb_ntrees; double gb_initF; const int* gb_off; const int* gb_var; const double* gb_code;
  const int* gb_left; const int* gb_right; const int* gb_miss; double w_b;
  // svm
  int S; const double* sv; const double* alpha; const double* xc; const double* xis;
  double sigma, bias, ys, yc, w_v;
};

__global__ void __launch_bounds__(128) k_ens_points(const double* __restrict__ X, int n, PointModels pm,
                                                    double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[16];
  bool anynan = false;
  for (int f = 0; f < pm.P; ++f) { x[f] = X[(size_t)f * n + i]; anynan |= (x[f] != x[f]); }
  const SmoothParams& sp = pm.sp;
  double s = 0.0;
  if (!anynan) {
    s = smooth_models(x, pm.P, sp);
    if (pm.rf_ntree > 0) {
      double a = 0.0;
      for (int t = 0; t < pm.rf_ntree; ++t) {
        const size_t o = (size_t)t * pm.rf_nrnodes;
        int k = 0;
        while (pm.rf_status[o + k] != -1)
          k = (x[pm.rf_var[o + k] - 1] <= pm.rf_split[o + k] ? pm.rf_left[o + k] : pm.rf_right[o + k]) - 1;
        a += pm.rf_pred[o + k];
      }
      s = fma(pm.w_r, a / pm.rf_ntree, s);
    }
    if (pm.S > 0) {
      double a = 0.0;
      for (int q = 0; q < pm.S; ++q) {
        double d2 = 0.0;
        for (int f = 0; f < pm.P; ++f) {
          const double d = (x[f] - pm.xc[f]) * pm.xis[f] - pm.sv[(size_t)q * pm.P + f];
          d2 = fma(d, d, d2);
        }
        a = fma(pm.alpha[q], exp(-pm.sigma * d2), a);
      }
      s = fma(pm.w_v, (a - pm.bias) * pm.ys + pm.yc, s);
    }
  }
  if (pm.gb_ntrees > 0 && (!anynan || sp.only_gbm)) {
    double a = pm.gb_initF;
    for (int t = 0; t < pm.gb_ntrees; ++t) {
      const int o = pm.gb_off[t];
      int k = 0;
      while (pm.gb_var[o + k] != -1) {
        const double xv = x[pm.gb_var[o + k]];
        k = (xv != xv) ? pm.gb_miss[o + k] : (xv < pm.gb_code[o + k] ? pm.gb_left[o + k] : pm.gb_right[o + k]);
      }
      a += pm.gb_code[o + k];
    }
    s = fma(pm.w_b, a, s);
  }
  double r = s / sp.w_total;
  if (anynan && !sp.only_gbm) r = __longlong_as_double(0x7ff8000000000000LL);
  out[i] = r;
}

// ---------------------------------------------------------------------------------------------
// host: create / evaluate
// ---------------------------------------------------------------------------------------------
mb_ensemble* ensemble_create(mb_ctx* ctx, const mb_grid& g, const mb_models& m, const char* kept, const double* w,
                             double w_total) {
  cudaStream_t st = ctx->stream;
  auto e = std::make_unique<mb_ensemble>();
  e->ctx = ctx;
  e->device = ctx->device;
  e->g = g;
  e->P = m.P;
  e->C = m.P - 2;
  MB_REQUIRE(m.P >= 2 && m.P <= 16, "P = C + 2 must be in [2, 16]");
  MB_REQUIRE(w_total > 0, "w_total must be positive");
  const int nk = (int)std::strlen(kept);
  for (int i = 0; i < nk; ++i) {
    const int k = letter_index(kept[i]);
    MB_REQUIRE(!e->has[k], "model letter repeated in `kept`");
    e->has[k] = true;
    e->w[k] = w[i];
  }
  e->w_total = w_total;
  e->only_gbm = e->has[MB_B] && nk == 1;
  const int P = m.P, C = e->C;

  if (e->has[MB_G]) {
    MB_REQUIRE(m.gam_coef, "gam kept but gam_coef is NULL");
    e->gam.upload(m.gam_coef, P + 1, st);
  }
  if (e->has[MB_N]) {
    MB_REQUIRE(m.nn_wts && m.nn_H >= 1, "nnet kept but descriptor is empty");
    e->nn_H = m.nn_H; e->nn_max2 = m.nn_max2; e->nn_min = m.nn_min;
    e->nn.upload(m.nn_wts, (size_t)(P + 1) * m.nn_H + m.nn_H + 1, st);
  }
  if (e->has[MB_M]) {
    MB_REQUIRE(m.mars_T >= 1 && m.mars_dirs && m.mars_cuts && m.mars_coef, "earth kept but descriptor is empty");
    std::vector<int> off(1, 0), var, dir;
    std::vector<double> cut;
    for (int t = 0; t < m.mars_T; ++t) {
      for (int v = 0; v < P; ++v) {
        const int dd = m.mars_dirs[(size_t)t * P + v];
        if (dd != 0) {
          MB_REQUIRE(dd == 1 || dd == -1 || dd == 2, "earth dirs must be 0, +-1 or 2");
          var.push_back(v); dir.push_back(dd); cut.push_back(m.mars_cuts[(size_t)t * P + v]);
        }
      }
      off.push_back((int)var.size());
    }
    e->mars_T = m.mars_T;
    e->mars_coef.upload(m.mars_coef, m.mars_T, st);
    e->mars_off.upload(off, st);
    if (var.empty()) { var.push_back(0); dir.push_back(1); cut.push_back(0.0); }
    e->mars_var.upload(var, st); e->mars_dir.upload(dir, st); e->mars_cut.upload(cut, st);
  }
  if (e->has[MB_V]) {
    MB_REQUIRE(m.svm_S >= 1 && m.svm_sv && m.svm_alpha && m.svm_x_center && m.svm_x_scale, "ksvm kept but descriptor is empty");
    const int S = m.svm_S;
    const double l2e = 1.4426950408889634;
    const int NQ = (P + 1) / 2, npairs = (S + 1) / 2;
    std::vector<float4> svp((size_t)npairs * NQ, make_float4(0.f, 0.f, 0.f, 0.f)), bap(npairs, make_float4(0.f, 0.f, 0.f, 0.f));
    std::vector<double> xis(P);
    for (int f = 0; f < P; ++f) {
      MB_REQUIRE(m.svm_x_scale[f] != 0, "ksvm x.scale has a zero entry");
      xis[f] = 1.0 / m.svm_x_scale[f];
    }
    for (int i = 0; i < S; ++i) {
      double n2 = 0;
      float* ba = reinterpret_cast<float*>(&bap[i / 2]);
      for (int f = 0; f < P; ++f) {
        const double v = m.svm_sv[(size_t)i * P + f];
        n2 += v * v;
        float* q4 = reinterpret_cast<float*>(&svp[(size_t)(i / 2) * NQ + f / 2]);
        q4[(f & 1) * 2 + (i & 1)] = (float)(2.0 * m.svm_sigma * l2e * v);
      }
      ba[i & 1] = (float)(-m.svm_sigma * n2 * l2e);
      ba[2 + (i & 1)] = (float)m.svm_alpha[i];
    }
    e->svm_S = S;
    e->svm_pairs = npairs;
    e->svm_svp.upload(svp, st); e->svm_bap.upload(bap, st);
    e->svm_xc.upload(m.svm_x_center, P, st);
    e->svm_xis.upload(xis, st);
    e->svm_bias = m.svm_b; e->svm_sigma = m.svm_sigma; e->svm_yc = m.svm_y_center; e->svm_ys = m.svm_y_scale;
    e->svp_sv.upload(m.svm_sv, (size_t)S * P, st);
    e->svp_alpha.upload(m.svm_alpha, S, st);
    if (ctx->svm_impl != 2 && P <= 8) {
      // 3 x TF32 operands: value = hi + lo with hi, lo exactly representable in TF32 (low 13 mantissa bits zero)
      auto tf32 = [](float x) { uint32_t u; std::memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xffffe000u; float y; std::memcpy(&y, &u, 4); return y; };
      const int noct = (S + 7) / 8;
      std::vector<float4> bq((size_t)noct * 32, make_float4(0.f, 0.f, 0.f, 0.f)), baq((size_t)noct * 4, make_float4(0.f, 0.f, 0.f, 0.f));
      for (int o = 0; o < noct; ++o)
        for (int ln = 0; ln < 32; ++ln) {
          const int i = 8 * o + (ln >> 2), tig = ln & 3;
          if (i >= S) continue;
          float v[2];
          for (int fh = 0; fh < 2; ++fh) {
            const int f = tig + 4 * fh;
            v[fh] = f < P ? (float)(2.0 * m.svm_sigma * l2e * m.svm_sv[(size_t)i * P + f]) : 0.f;
          }
          const float h0 = tf32(v[0]), h1 = tf32(v[1]);
          bq[(size_t)o * 32 + ln] = make_float4(h0, h1, tf32(v[0] - h0), tf32(v[1] - h1));
        }
      for (int pr = 0; pr < npairs; ++pr) baq[pr] = bap[pr];
      e->svm_oct = noct;
      e->svm_bq.upload(bq, st); e->svm_baq.upload(baq, st);
      // FP16 split operands for mma.m16n8k16 (svm_impl 3): value = hi + lo, both halves; the factor 2 sigma log2(e) is shared out
      // between x and the support vectors by a power of two so that both stay in the middle of the half range
      const double fac = 2.0 * m.svm_sigma * l2e;
      const double hs = std::exp2(std::round(0.5 * std::log2(fac)));
      std::vector<uint4> bh((size_t)noct * 32, make_uint4(0u, 0u, 0u, 0u));
      bool in_range = std::isfinite(hs) && hs > 0;
      for (int o = 0; o < noct && in_range; ++o)
        for (int ln = 0; ln < 32; ++ln) {
          const int i = 8 * o + (ln >> 2), tig = ln & 3;
          if (i >= S) continue;
          unsigned short hi[2], lo[2];
          for (int k = 0; k < 2; ++k) {
            const int f = 2 * tig + k;
            const float v = f < P ? (float)(fac / hs * m.svm_sv[(size_t)i * P + f]) : 0.f;
            if (!(std::fabs(v) < 16384.f)) in_range = false;
            const __half h = __float2half_rn(v);
            hi[k] = __half_as_ushort(h);
            lo[k] = __half_as_ushort(__float2half_rn(v - __half2float(h)));
          }
          // b0 = b1 (K rows 2t, 2t+1 and 2t+8, 2t+9 are the same two features): stored twice so that one LDS.128 delivers both register pairs
          const unsigned h2 = (unsigned)hi[0] | ((unsigned)hi[1] << 16), l2 = (unsigned)lo[0] | ((unsigned)lo[1] << 16);
          bh[(size_t)o * 32 + ln] = make_uint4(h2, h2, l2, l2);
        }
      if (in_range) { e->svm_hs = hs; e->svm_bh.upload(bh, st); }
    }
  }
  std::vector<int2> nodes;          // both forests, rf first (child indices are absolute)
  std::vector<int> froots;
  // Trees are re-laid out independently (breadth-first, children adjacent, child indices local to the tree) on all
  // host threads, then concatenated; the child index of every split node is shifted by the tree's offset.
  auto pack_forest = [&](int nt, auto&& pack_one) {
    std::vector<std::vector<int2>> packs(nt);
    std::vector<std::exception_ptr> errs;
    const int nthr = std::max(1, std::min<int>(nt, (int)std::thread::hardware_concurrency()));
    errs.resize(nthr);
    std::vector<std::thread> thr;
    for (int w = 0; w < nthr; ++w)
      thr.emplace_back([&, w] {
        try {
          for (int t = w; t < nt; t += nthr) pack_one(t, packs[t]);
        } catch (...) { errs[w] = std::current_exception(); }
      });
    for (auto& th : thr) th.join();
    for (auto& er : errs) if (er) std::rethrow_exception(er);
    for (int t = 0; t < nt; ++t) {
      const size_t off = nodes.size();
      MB_REQUIRE(off + packs[t].size() < ((size_t)1 << 26), "forest too large for the packed node format");
      froots.push_back((int)off);
      for (const int2& nd : packs[t])
        nodes.push_back((nd.y & kMetaLeaf) ? nd : make_int2(nd.x, nd.y + (int)(off << 5)));
    }
  };
  if (e->has[MB_R]) {
    MB_REQUIRE(m.rf_ntree >= 1 && m.rf_nrnodes >= 1 && m.rf_left && m.rf_right && m.rf_status && m.rf_bestvar &&
                   m.rf_split && m.rf_nodepred, "randomForest kept but descriptor is empty");
    const int nt = m.rf_ntree, nn = m.rf_nrnodes;
    // leaf offset: the mean leaf value of the first tree is a good centre for the float32 leaf storage
    double off = 0; int cnt = 0;
    for (int k = 0; k < nn; ++k) if (m.rf_status[k] == -1) { off += m.rf_nodepred[k]; ++cnt; }
    off = cnt ? off / cnt : 0.0;
    pack_forest(nt, [&](int t, std::vector<int2>& out) {
      const size_t o = (size_t)t * nn;
      std::queue<std::pair<int, size_t>> q;   // (reference node, packed position)
      out.reserve(256);
      out.push_back(make_int2(0, 0));
      q.push({0, 0});
      while (!q.empty()) {
        auto [k, pos] = q.front();
        q.pop();
        MB_REQUIRE(k >= 0 && k < nn, "randomForest: daughter index out of range");
        if (m.rf_status[o + k] == -1) {
          const float v = (float)(m.rf_nodepred[o + k] - off);
          out[pos] = make_int2(__builtin_bit_cast(int, v), kMetaLeaf);
          continue;
        }
        const int var = m.rf_bestvar[o + k] - 1;
        MB_REQUIRE(var >= 0 && var < P, "randomForest: bestvar out of range");
        float thr; bool swap;
        convert_split(g, C, var, m.rf_split[o + k], false, &thr, &swap);
        const size_t child = out.size();
        out.push_back(make_int2(0, 0));
        out.push_back(make_int2(0, 0));
        out[pos] = make_int2(__builtin_bit_cast(int, thr), (int)(child << 5) | var);
        const int l = m.rf_left[o + k] - 1, r = m.rf_right[o + k] - 1;
        q.push({swap ? r : l, child});
        q.push({swap ? l : r, child + 1});
      }
    });
    e->rf.ntrees = nt; e->rf.offset = off;
    e->rf_ntree = nt; e->rf_nrnodes = nn;
    const size_t tot = (size_t)nt * nn;
    e->rfp_left.upload(m.rf_left, tot, st); e->rfp_right.upload(m.rf_right, tot, st);
    e->rfp_status.upload((const signed char*)m.rf_status, tot, st); e->rfp_var.upload(m.rf_bestvar, tot, st);
    e->rfp_split.upload(m.rf_split, tot, st); e->rfp_pred.upload(m.rf_nodepred, tot, st);
  }
  if (e->has[MB_B]) {
    MB_REQUIRE(m.gbm_ntrees >= 1 && m.gbm_tree_off && m.gbm_splitvar && m.gbm_splitcode && m.gbm_left &&
                   m.gbm_right && m.gbm_missing, "gbm kept but descriptor is empty");
    const int nt = m.gbm_ntrees;
    pack_forest(nt, [&](int t, std::vector<int2>& out) {
      const int o = m.gbm_tree_off[t], cntn = m.gbm_tree_off[t + 1] - o;
      std::queue<std::pair<int, size_t>> q;
      out.reserve(64);
      out.push_back(make_int2(0, 0));
      q.push({0, 0});
      while (!q.empty()) {
        auto [k, pos] = q.front();
        q.pop();
        MB_REQUIRE(k >= 0 && k < cntn, "gbm: child index out of range");
        const int var = m.gbm_splitvar[o + k];
        if (var == -1) {
          const float v = (float)m.gbm_splitcode[o + k];
          out[pos] = make_int2(__builtin_bit_cast(int, v), kMetaLeaf);
          continue;
        }
        MB_REQUIRE(var >= 0 && var < P, "gbm: SplitVar out of range");
        float thr; bool swap;
        convert_split(g, C, var, m.gbm_splitcode[o + k], true, &thr, &swap);
        const size_t child = out.size();
        for (int c3 = 0; c3 < 3; ++c3) out.push_back(make_int2(0, 0));
        out[pos] = make_int2(__builtin_bit_cast(int, thr), (int)(child << 5) | var);
        const int l = m.gbm_left[o + k], r = m.gbm_right[o + k];
        q.push({swap ? r : l, child});
        q.push({swap ? l : r, child + 1});
        q.push({m.gbm_missing[o + k], child + 2});
      }
    });
    e->gbm.ntrees = nt;
    e->gb_ntrees = nt; e->gb_initF = m.gbm_initF;
    const size_t tot = m.gbm_tree_off[nt];
    e->gbp_off.upload(m.gbm_tree_off, nt + 1, st); e->gbp_var.upload(m.gbm_splitvar, tot, st);
    e->gbp_code.upload(m.gbm_splitcode, tot, st); e->gbp_left.upload(m.gbm_left, tot, st);
    e->gbp_right.upload(m.gbm_right, tot, st); e->gbp_miss.upload(m.gbm_missing, tot, st);
  }
  if (!froots.empty()) {
    e->forest_zero_leaf = (int)nodes.size();
    nodes.push_back(make_int2(0, kMetaLeaf));       // leaf holding 0.0f: padding entry of the residual lists
    e->forest_nodes.upload(nodes, st);
    e->forest_roots.upload(froots, st);
  }
  MB_CUDA(cudaStreamSynchronize(st));
  return e.release();
}

void ensemble_free(mb_ensemble* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  delete e;
}

mb_grid ensemble_grid(const mb_ensemble* e) { return e->g; }
bool ensemble_has_forest_kernel(const mb_ensemble* e) { return (e->has[MB_R] || e->has[MB_B]) && !e->only_gbm; }
int ensemble_ncov(const mb_ensemble* e) { return e->C; }

static SmoothParams smooth_params(const mb_ensemble* e) {
  SmoothParams sp{};
  sp.gam = e->has[MB_G] ? e->gam.p : nullptr; sp.w_g = e->w[MB_G];
  sp.nn = e->has[MB_N] ? e->nn.p : nullptr; sp.nn_H = e->nn_H; sp.nn_max2 = e->nn_max2; sp.nn_min = e->nn_min;
  sp.w_n = e->w[MB_N];
  sp.mars_T = e->has[MB_M] ? e->mars_T : 0; sp.mars_coef = e->mars_coef.p; sp.mars_off = e->mars_off.p;
  sp.mars_var = e->mars_var.p; sp.mars_dir = e->mars_dir.p; sp.mars_cut = e->mars_cut.p; sp.w_m = e->w[MB_M];
  sp.w_total = e->w_total;
  sp.only_gbm = e->only_gbm ? 1 : 0;
  return sp;
}

static void launch_trees(mb_ctx* ctx, const mb_ensemble* e, const float* cov, int C, int64_t plane, const EnsGeom& eg,
                         const mb_window& w, const int* roots, int n_rf, int n_gb, double* acc, int accumulate, cudaStream_t st) {
  const double rf_scale = n_rf ? e->w[MB_R] / n_rf : 0.0, gb_scale = e->w[MB_B];
  const double base = (n_rf ? e->w[MB_R] * e->rf.offset : 0.0) + (n_gb ? e->w[MB_B] * e->gb_initF : 0.0);
  const int R = 1;                                          // one cell per thread (R = 2 / 4 measured slower in round 1, profiles/r1r_tree_rows.txt)
  const size_t smem = sizeof(float) * (size_t)(C + 2) * kTreeThreads +
                      sizeof(int) * 8 * ((kTreeSeg + kTreeIlp) + (kWarpList + kTreeIlp));
  dim3 grid((w.c1 - w.c0 + 31) / 32, (w.r1 - w.r0 + 7) / 8);
#define MB_TREES_CASE(RR)                                                                                         \
  case RR: {                                                                                                      \
    static thread_local bool attr = false;                                                                        \
    if (!attr) {                                                                                                  \
      MB_CUDA(cudaFuncSetAttribute(k_ens_trees<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));    \
      attr = true;                                                                                                \
    }                                                                                                             \
    MB_LAUNCH(ctx, "k_ens_trees", st) k_ens_trees<RR><<<grid, kTreeThreads, smem, st>>>(                          \
        cov, C, plane, eg, w, e->forest_nodes.p, roots, n_rf, n_gb, e->forest_zero_leaf, rf_scale, gb_scale, base, \
        accumulate, acc_stride(w), acc, ctx->tree_levels == 1 ? 1 : 2);                                                    \
  } break;
  switch (R) {
    MB_TREES_CASE(1)
    default: throw Error(MB_E_ARG, "tree tile rows-per-thread must be 1");
  }
#undef MB_TREES_CASE
}

static SmoothParams smooth_params(const mb_ensemble* e);


// ---------------------------------------------------------------------------------------------
// ksvm on the tensor pipe (k_ens_svm_tma below; the default for P <= 8; measured 46.1 against 52.7 ms for k_ens_svm on config 3,
// profiles/r2d_ens_check.txt): the 8-feature dot products of
// k_ens_svm on the tensor pipe.  ncu (r1s) shows k_ens_svm bound by the FMA pipe (67 % busy: 16 of the 20 packed FP32
// instructions per 2 SV x 2 cells are the dot products), with MUFU.EX2 at 53 %; moving the dot products to
// mma.sync.m16n8k8.tf32 leaves 2 FADD2 + 2 FFMA2 + 4 MUFU per 16 cells x 8 SVs per M-tile on the other pipes.
// Accuracy: 3 x TF32 (x = x_hi + x_lo, sv = sv_hi + sv_lo, products hi.hi + hi.lo + lo.hi, FP32 accumulation), i.e. a
// relative error of ~2^-21 per product instead of TF32's 2^-11 - the exponent needs ~1e-6 absolute.
// A warp owns 32 cells of one row = two 16-cell M-tiles; the A fragments (cells x features) stay in registers for the whole
// support-vector loop, the B fragments (features x 8 SVs) come pre-shuffled from the host, one LDS.128 per octet.
// Fragment layouts (PTX ISA, m16n8k8 .tf32; g = lane >> 2, t = lane & 3):  A a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);
// B b0 (k = t, n = g) b1 (k = t+4, n = g);  C/D c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1).
// ---------------------------------------------------------------------------------------------
constexpr int kSvmOct = 32;          // octets (256 support vectors) per shared-memory stage

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const float (&a)[4], float b0, float b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                 "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

// ---------------------------------------------------------------------------------------------
// k_ens_svm_tma: k_ens_svm_mma with the covariate tile staged by TMA (the default ksvm kernel for P <= 8).
//   * a CTA of 8 warps walks tiles of 32 x 8 cells (warp = one row of 32 cells = two 16-cell M-tiles); the tile's C covariate
//     planes arrive by ONE 3-D tensor copy (cp.async.bulk.tensor, box 32 x 8 x C, UTMALDG.3D; cells outside the raster are
//     filled with NaN = NA) into a two-stage shared-memory ring: the copy of the next tile is in flight while the current one is
//     evaluated (mbarrier complete_tx).  A raster whose row stride is not a multiple of 16 bytes cannot be described by a tensor
//     map: the same ring is then filled with plain loads.
//   * two schedules.  PERSISTENT (mode 0; used when forests are kept too): 2 CTAs per SM - enough warps to saturate MUFU.EX2,
//     which bounds this kernel - loop over all tiles while k_ens_trees, launched on a second stream, fills the rest of every SM
//     with its latency-bound node walks: the two kernels saturate different resources and run side by side (each adds its sum
//     to the zeroed accumulator with one RED.ADD.F64 per cell - two commutative adds, so the result does not depend on who
//     comes first).  CHUNKED (mode 1; no forest): a CTA owns 8 consecutive tiles and retires, so that the small kernels of the
//     TPS fit on the other stream keep finding free SMs.
//   A single fused kernel (forest warps + ksvm warps per CTA, one TMA-staged tile, accumulator written once) was built and
//   measured first (profiles/r2d_*): DRAM traffic 1.007 x the algorithmic bytes, but 97 ms against 85 for the two kernels one
//   after the other - the forest walk is bound by L2 latency and needs every warp slot of the SM, which a fused CTA cannot give it.
// ---------------------------------------------------------------------------------------------
constexpr int kTileCells = 256;                            // 32 x 8
constexpr int kSvmChunkTiles = 8;
struct __align__(16) SvmTmaScratch {
  float4 bq[kSvmOct * 32];
  float4 ba[kSvmOct * 4];
  unsigned long long bar[2];
};
struct SvmTmaArgs {
  const float* cov; int C, P; int64_t plane; EnsGeom eg; mb_window w;
  const float4* bq; const float4* baq; int noct; const double* xc; const double* xis;
  const uint4* bh; double hs;      // FP16 layout (F16 instantiations)
  double sigma, bias, ys, yc, wv;
  SmoothParams sp;
  int64_t acc_stride; double* acc;
  int epilogue;                    // 0 store, 1 acc += (plain read-modify-write, after the forest kernel), 2 RED.ADD (beside it)
  int use_tma, tiles_x, ntiles, chunked;
  int cbase;                       // column of the first tile: w.c0 snapped down to a multiple of 4 (16-byte aligned tile rows)
};

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void svm_octet(const float4 bf, const float4 ba, const float (&ah)[2][4], const float (&al)[2][4],
                                          const float (&a0c)[2][2], float2 (&part)[2][2]) {
  // the two M-tiles advance in lock-step (the mma statements are volatile, so this IS the issue order): the three dependent
  // HMMAs of one tile hide behind those of the other instead of stalling the warp
  float d[2][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const float2 d01 = __fadd2_rn(make_float2(a0c[mt][0], a0c[mt][0]), make_float2(ba.x, ba.y));
    const float2 d23 = __fadd2_rn(make_float2(a0c[mt][1], a0c[mt][1]), make_float2(ba.x, ba.y));
    d[mt][0] = d01.x; d[mt][1] = d01.y; d[mt][2] = d23.x; d[mt][3] = d23.y;
  }
  mma_tf32(d[0], al[0], bf.x, bf.y);          // x_lo . sv_hi
  mma_tf32(d[1], al[1], bf.x, bf.y);
  mma_tf32(d[0], ah[0], bf.z, bf.w);          // x_hi . sv_lo
  mma_tf32(d[1], ah[1], bf.z, bf.w);
  mma_tf32(d[0], ah[0], bf.x, bf.y);          // x_hi . sv_hi
  mma_tf32(d[1], ah[1], bf.x, bf.y);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    part[mt][0] = __ffma2_rn(make_float2(ba.z, ba.w), make_float2(ex2_approx(d[mt][0]), ex2_approx(d[mt][1])), part[mt][0]);
    part[mt][1] = __ffma2_rn(make_float2(ba.z, ba.w), make_float2(ex2_approx(d[mt][2]), ex2_approx(d[mt][3])), part[mt][1]);
  }
}

// The same octet with FP16 split operands (svm_impl 3): K = 16 holds the 8 features twice - A = [x_hi | x_lo], B = [sv_hi ; sv_hi] and
// then [sv_lo ; sv_lo] - so TWO HMMA.16816 give x_hi.sv_hi + x_lo.sv_hi + x_hi.sv_lo + x_lo.sv_lo (products of halves are exact in
// the FP32 accumulator: 22 bits per operand like 3 x TF32, plus the lo.lo term that scheme drops), the A fragments are the same
// four registers for both.
__device__ __forceinline__ void svm_octet16(const uint4 bf, const float4 ba, const uint32_t (&ax)[2][4], const float (&a0c)[2][2],
                                            float2 (&part)[2][2]) {
  float d[2][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const float2 d01 = __fadd2_rn(make_float2(a0c[mt][0], a0c[mt][0]), make_float2(ba.x, ba.y));
    const float2 d23 = __fadd2_rn(make_float2(a0c[mt][1], a0c[mt][1]), make_float2(ba.x, ba.y));
    d[mt][0] = d01.x; d[mt][1] = d01.y; d[mt][2] = d23.x; d[mt][3] = d23.y;
  }
  mma_f16(d[0], ax[0], bf.z, bf.w);           // (x_hi + x_lo) . sv_lo
  mma_f16(d[1], ax[1], bf.z, bf.w);
  mma_f16(d[0], ax[0], bf.x, bf.y);           // (x_hi + x_lo) . sv_hi
  mma_f16(d[1], ax[1], bf.x, bf.y);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    part[mt][0] = __ffma2_rn(make_float2(ba.z, ba.w), make_float2(ex2_approx(d[mt][0]), ex2_approx(d[mt][1])), part[mt][0]);
    part[mt][1] = __ffma2_rn(make_float2(ba.z, ba.w), make_float2(ex2_approx(d[mt][2]), ex2_approx(d[mt][3])), part[mt][1]);
  }
}

template <bool F16>
__global__ void __launch_bounds__(kSvmThreads, 4) k_ens_svm_tma(const __grid_constant__ CUtensorMap tmap, SvmTmaArgs a) {
  extern __shared__ __align__(128) unsigned char svm_smem[];
  const int C = a.C;
  const int stage_floats = (C * kTileCells + 31) & ~31;
  float* stage0 = reinterpret_cast<float*>(svm_smem);
  float* stage1 = stage0 + stage_floats;
  SvmTmaScratch& S = *reinterpret_cast<SvmTmaScratch*>(stage1 + stage_floats);
  uint64_t* bars = reinterpret_cast<uint64_t*>(S.bar);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const mb_window w = a.w;
  const uint32_t tile_bytes = (uint32_t)C * kTileCells * 4u;
  const bool tma = a.use_tma != 0 && C > 0;
  const int first = a.chunked ? blockIdx.x * kSvmChunkTiles : blockIdx.x;
  const int step = a.chunked ? 1 : gridDim.x;
  const int last = a.chunked ? min(a.ntiles, first + kSvmChunkTiles) : a.ntiles;
  if (tid == 0 && tma) {
    ac_mbar_init(&bars[0], 1);
    ac_mbar_init(&bars[1], 1);
    ac_fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0 && tma && first < last) {
    ac_mbar_expect_tx(&bars[0], tile_bytes);
    ac_tma_load_3d(stage0, &tmap, a.cbase + (first % a.tiles_x) * 32, w.r0 + (first / a.tiles_x) * 8, 0, &bars[0]);
  }
  int it = 0;
  for (int tile = first; tile < last; tile += step, ++it) {
    float* feat = (it & 1) ? stage1 : stage0;
    const int col0 = a.cbase + (tile % a.tiles_x) * 32, row0 = w.r0 + (tile / a.tiles_x) * 8;
    if (tma) {
      if (tid == 0 && tile + step < last) {               // the other stage was released by the barrier that ended the last tile
        const int nt = tile + step;
        ac_mbar_expect_tx(&bars[(it + 1) & 1], tile_bytes);
        ac_tma_load_3d((it & 1) ? stage0 : stage1, &tmap, a.cbase + (nt % a.tiles_x) * 32, w.r0 + (nt / a.tiles_x) * 8, 0,
                       &bars[(it + 1) & 1]);
      }
      ac_mbar_wait(&bars[it & 1], (uint32_t)((it >> 1) & 1));
    } else {
      for (int i = tid; i < C * kTileCells; i += kSvmThreads) {
        const int f = i >> 8, cell = i & 255, r = row0 + (cell >> 5), c = col0 + (cell & 31);
        feat[i] = (r < a.eg.nrow && c < a.eg.ncol) ? __ldg(&a.cov[f * a.plane + (int64_t)r * a.eg.ncol + c]) : __int_as_float(0x7fc00000);
      }
      __syncthreads();
    }
    const int row = row0 + warp;
    const float* frow = feat + warp * 32;
    float ah[2][4], al[2][4], a0c[2][2];
    uint32_t ax[2][4];                                 // F16: a0 / a1 = (hi, hi) of features 2t, 2t+1 of rows g / g + 8, a2 / a3 = (lo, lo)
    unsigned nanbits = 0;                              // bit 2 mt + hr: an NA layer in cell (mt, hr) of this lane's group
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        const int cl = 16 * mt + g + 8 * hr;
        const int col = col0 + cl;
        const bool live = col >= w.c0 && col < w.c1 && row < w.r1;
        double n2 = 0.0;
        int anynan = 0;
        float xf[2];
#pragma unroll
        for (int fh = 0; fh < 2; ++fh) {
          const int f = F16 ? 2 * t + fh : t + 4 * fh;
          double xs = 0.0;
          if (f < a.P) {
            double v;
            if (f < C) v = live ? (double)frow[f * kTileCells + cl] : 0.0;
            else if (f == C) v = a.eg.xmin + (col + 0.5) * a.eg.rx;
            else v = a.eg.ymax - (row + 0.5) * a.eg.ry;
            anynan |= (v != v);
            xs = (v - a.xc[f]) * a.xis[f];
          }
          n2 += xs * xs;
          if constexpr (F16) {
            xf[fh] = fminf(fmaxf((float)(xs * a.hs), -32768.f), 32768.f);   // beyond that exp(-sigma |x - sv|^2) is 0 in any precision
          } else {
            xf[fh] = (float)xs;
            const float hi = to_tf32(xf[fh]);
            ah[mt][hr + 2 * fh] = hi;
            al[mt][hr + 2 * fh] = to_tf32(xf[fh] - hi);
          }
        }
        if constexpr (F16) {
          const __half2 hi = __floats2half2_rn(xf[0], xf[1]);
          const float2 hf = __half22float2(hi);
          const __half2 lo = __floats2half2_rn(xf[0] - hf.x, xf[1] - hf.y);
          ax[mt][hr] = *reinterpret_cast<const uint32_t*>(&hi);
          ax[mt][2 + hr] = *reinterpret_cast<const uint32_t*>(&lo);
        }
        n2 += __shfl_xor_sync(0xffffffffu, n2, 1);
        n2 += __shfl_xor_sync(0xffffffffu, n2, 2);
        anynan |= __shfl_xor_sync(0xffffffffu, anynan, 1);
        anynan |= __shfl_xor_sync(0xffffffffu, anynan, 2);
        a0c[mt][hr] = (float)(-a.sigma * n2 * 1.4426950408889634);
        nanbits |= (anynan ? 1u : 0u) << (2 * mt + hr);
      }
    float2 part[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) part[mt][hr] = make_float2(0.f, 0.f);
    for (int base = 0; base < a.noct; base += kSvmOct) {
      const int n = min(kSvmOct, a.noct - base);
      __syncthreads();
      uint4* const sbh = reinterpret_cast<uint4*>(S.bq);
      if constexpr (F16) {
        for (int i = tid; i < n * 32; i += kSvmThreads) sbh[i] = __ldg(&a.bh[(size_t)base * 32 + i]);
      } else {
        for (int i = tid; i < n * 32; i += kSvmThreads) S.bq[i] = __ldg(&a.bq[(size_t)base * 32 + i]);
      }
      for (int i = tid; i < n * 4; i += kSvmThreads) S.ba[i] = __ldg(&a.baq[(size_t)base * 4 + i]);
      __syncthreads();
      // (b, b', alpha, alpha') in S.ba are those of support vectors 2t, 2t+1 of the octet
      if constexpr (F16) {
#pragma unroll 2
        for (int o = 0; o < n; ++o) svm_octet16(sbh[o * 32 + lane], S.ba[o * 4 + t], ax, a0c, part);
      } else {
#pragma unroll 2
        for (int o = 0; o < n; ++o) svm_octet(S.bq[o * 32 + lane], S.ba[o * 4 + t], ah, al, a0c, part);
      }
    }
    // per-cell totals: the four lanes of a group hold the 8 support-vector columns; then to the lane that owns the cell
    double mine = 0.0;
    const unsigned gbits = __shfl_sync(0xffffffffu, nanbits, 4 * (lane & 7));
    const int mynan = (gbits >> (lane >> 3)) & 1u;           // lane = 16 mt + 8 hr + g
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        double tot = (double)part[mt][hr].x + (double)part[mt][hr].y;
        tot += __shfl_xor_sync(0xffffffffu, tot, 1);
        tot += __shfl_xor_sync(0xffffffffu, tot, 2);
        const double v = __shfl_sync(0xffffffffu, tot, 4 * (lane & 7));
        if ((lane >> 4) == mt && ((lane >> 3) & 1) == hr) mine = v;
      }
    // everything that is per OWNED cell (cell `lane` of the row) comes after the support-vector loop, so that nothing of it
    // is live across it: the loop needs the registers for two independent accumulator sets
    const int ocol = col0 + lane;
    if (ocol >= w.c0 && ocol < w.c1 && row < w.r1) {
      double* const dst = a.acc + (int64_t)(row - w.r0) * a.acc_stride + (ocol - w.c0);
      double prev = 0.0;
      if (a.epilogue == 1) prev = __ldcs(dst);               // the forest kernel's sum of this cell
      double smooth = 0.0;                                   // the float64 smooth models
      if ((a.sp.gam || a.sp.nn || a.sp.mars_T > 0) && !mynan) {
        double x[16];
        for (int f = 0; f < C; ++f) x[f] = (double)frow[f * kTileCells + lane];
        x[C] = a.eg.xmin + (ocol + 0.5) * a.eg.rx;
        x[C + 1] = a.eg.ymax - (row + 0.5) * a.eg.ry;
        smooth = smooth_models(x, C + 2, a.sp);
      }
      // last link of the chain: NA rule (V73: terra::predict returns NA where any layer is NA)
      const double v = mynan ? __longlong_as_double(0x7ff8000000000000LL) : a.wv * ((mine - a.bias) * a.ys + a.yc) + smooth;
      if (a.epilogue == 2) atomicAdd(dst, v);
      else *dst = prev + v;
    }
    __syncthreads();                                         // this stage and the support-vector buffers may be overwritten
  }
}

static void launch_svm_tma(mb_ctx* ctx, const mb_ensemble* e, const float* cov, int64_t plane, const EnsGeom& eg,
                           const mb_window& w, double* acc, int epilogue, bool chunked, cudaStream_t st) {
  const mb_grid& g = e->g;
  SvmTmaArgs sa{};
  sa.cov = cov; sa.C = e->C; sa.P = e->P; sa.plane = plane; sa.eg = eg; sa.w = w;
  sa.bq = e->svm_bq.p; sa.baq = e->svm_baq.p; sa.noct = e->svm_oct; sa.xc = e->svm_xc.p; sa.xis = e->svm_xis.p;
  sa.sigma = e->svm_sigma; sa.bias = e->svm_bias; sa.ys = e->svm_ys; sa.yc = e->svm_yc; sa.wv = e->w[MB_V];
  sa.sp = smooth_params(e);
  sa.acc_stride = acc_stride(w); sa.acc = acc; sa.epilogue = epilogue;
  sa.cbase = w.c0 & ~3;      // tile rows start on 16-byte boundaries of the raster row (probed: tools/tma_probe.cu)
  sa.tiles_x = (w.c1 - sa.cbase + 31) / 32;
  sa.ntiles = sa.tiles_x * ((w.r1 - w.r0 + 7) / 8);
  sa.chunked = chunked ? 1 : 0;
  CUtensorMap tmap;
  std::memset(&tmap, 0, sizeof tmap);
  sa.use_tma = (ctx->ens_tma != 2 && make_plane_tensor_map(&tmap, cov, g.ncol, g.nrow, e->C, plane, 32, 8)) ? 1 : 0;
  const size_t stage = ((size_t)e->C * kTileCells + 31) / 32 * 32 * sizeof(float);
  const size_t smem = 2 * stage + sizeof(SvmTmaScratch);
  const int per_sm = ctx->svm_ctas_per_sm > 0 ? ctx->svm_ctas_per_sm : 2;
  const int grid = chunked ? (sa.ntiles + kSvmChunkTiles - 1) / kSvmChunkTiles : std::min(sa.ntiles, per_sm * ctx->sm_count);
  const bool f16 = ctx->svm_impl != 1 && e->svm_bh.p != nullptr;     // svm_impl 1 = 3 x TF32 (A/B), 0 / 3 = FP16 split operands
  sa.bh = e->svm_bh.p; sa.hs = e->svm_hs;
  static thread_local bool attr = false;
  if (!attr) {
    MB_CUDA(cudaFuncSetAttribute(k_ens_svm_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    MB_CUDA(cudaFuncSetAttribute(k_ens_svm_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr = true;
  }
  if (f16) MB_LAUNCH(ctx, "k_ens_svm_tma", st) k_ens_svm_tma<true><<<grid, kSvmThreads, smem, st>>>(tmap, sa);
  else MB_LAUNCH(ctx, "k_ens_svm_tma", st) k_ens_svm_tma<false><<<grid, kSvmThreads, smem, st>>>(tmap, sa);
}

template <int NQ>
static void launch_svm(const mb_ensemble* e, const float* cov, int64_t plane, const EnsGeom& eg, const mb_window& w,
                       double* acc, int accumulate, cudaStream_t st) {
  mb_ctx* ctx = e->ctx;
  dim3 grid((w.c1 - w.c0 + 31) / 32, (w.r1 - w.r0 + 15) / 16);
  MB_LAUNCH(ctx, "k_ens_svm", st) k_ens_svm<NQ><<<grid, kSvmThreads, 0, st>>>(
      cov, e->C, e->P, plane, eg, w, e->svm_svp.p, e->svm_bap.p, e->svm_pairs, e->svm_xc.p, e->svm_xis.p,
      e->svm_sigma, e->svm_bias, e->svm_ys, e->svm_yc, e->w[MB_V], accumulate, smooth_params(e), acc_stride(w), acc);
}

static EnsGeom ens_geom(const mb_grid& g) {
  return EnsGeom{g.xmin, g.ymax, (g.xmax - g.xmin) / g.ncol, (g.ymax - g.ymin) / g.nrow, g.nrow, g.ncol};
}

// acc <- sum_k round(w_k, 2) f_k(cell) over EVERY kept model for window w, in the padded accumulator layout
// (acc_stride(w) x acc_rows(w) doubles); NaN where a covariate is NA (except a gbm-only ensemble, which
// follows MissingNode).  Chain: trees -> svm (+ smooth models) | smooth; the last kernel applies the NA rule.
// part: 0 = the whole chain; 1 = the forest kernel only (it opens the chain: stores); 2 = what follows the forest kernel (adds).
// mb_mltps_predict* with SM partitions runs part 1 on the ensemble partition beside stage 1 of the fit and part 2 afterwards.
void ensemble_accumulate(mb_ctx* ctx, const mb_ensemble* e, const float* cov, int C, const mb_window& w, double* acc,
                         cudaStream_t st, int part) {
  const mb_grid& g = e->g;
  const int64_t plane = (int64_t)g.nrow * g.ncol;
  const int64_t astride = acc_stride(w);
  const EnsGeom eg = ens_geom(g);
  bool started = false;
  if ((e->has[MB_R] || e->has[MB_B]) && part == 2) {
    if (e->only_gbm) return;
    started = true;                           // the forest kernel has run (part 1)
  } else if (e->has[MB_R] || e->has[MB_B]) {
    const int n_rf = e->has[MB_R] ? e->rf.ntrees : 0, n_gb = e->has[MB_B] ? e->gbm.ntrees : 0;
    const int* roots = e->forest_roots.p;   // rf trees first (if kept), then gbm
    if (e->only_gbm) {
      const int64_t ncell = (int64_t)(w.r1 - w.r0) * (w.c1 - w.c0);
      MB_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * astride * (w.r1 - w.r0), st));   // k_ens_trees_plain accumulates
      MB_LAUNCH(ctx, "k_ens_trees_plain", st) k_ens_trees_plain<<<(unsigned)((ncell + kTreeThreads - 1) / kTreeThreads), kTreeThreads, 0, st>>>(
          cov, C, plane, eg, w, e->forest_nodes.p, roots, 0, 0.0, 0.0, e->forest_nodes.p, roots, n_gb, e->w[MB_B],
          e->gb_initF, 1, astride, acc);
      MB_CUDA(cudaGetLastError());
      return;                                // gbm alone: numbers on NA cells too, nothing else to add
    }
    // "ens_overlap" = 1 (A/B measurements; NOT the default - measured 139 ms against 83 ms one after the other on config 3,
    // profiles/r2e_ens_check.txt: the persistent ksvm CTAs need 4 per SM to hide the MUFU latency, which leaves the forest kernel
    // nothing): the two kernels side by side, each adding its sum to the zeroed accumulator with RED.ADD.F64
    const bool overlap = e->has[MB_V] && e->svm_oct > 0 && ctx->ens_overlap == 1 && part == 0;
    if (overlap) {
      if (!ctx->ens_aux) {
        int prio_lo = 0, prio_hi = 0;
        MB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        MB_CUDA(cudaStreamCreateWithPriority(&ctx->ens_aux, cudaStreamNonBlocking, prio_hi));
        for (cudaEvent_t& ev : ctx->ev_ens) MB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      }
      MB_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * astride * (w.r1 - w.r0), st));
      MB_CUDA(cudaEventRecord(ctx->ev_ens[0], st));
      MB_CUDA(cudaStreamWaitEvent(ctx->ens_aux, ctx->ev_ens[0], 0));
      launch_svm_tma(ctx, e, cov, plane, eg, w, acc, /*epilogue=*/2, /*chunked=*/false, ctx->ens_aux);
      launch_trees(ctx, e, cov, C, plane, eg, w, roots, n_rf, n_gb, acc, 2, st);
      MB_CUDA(cudaEventRecord(ctx->ev_ens[1], ctx->ens_aux));
      MB_CUDA(cudaStreamWaitEvent(st, ctx->ev_ens[1], 0));
      MB_CUDA(cudaGetLastError());
      return;
    }
    // "ens_order" = 2: the ksvm kernel first, the forest kernel second (default: the other way round).  Inside mb_mltps_predict the
    // first kernel runs beside the bulge chase of the fit, whose CTAs take registers from it on 80 SMs (sbr.cu); measured on config 3
    // with the small-footprint chase: forests first 140.2 ms / step, ksvm first 142.4 (profiles/r2n_*).
    if (e->has[MB_V] && e->svm_oct > 0 && ctx->ens_order == 2 && part == 0) {
      launch_svm_tma(ctx, e, cov, plane, eg, w, acc, /*epilogue=*/0, /*chunked=*/true, st);
      launch_trees(ctx, e, cov, C, plane, eg, w, roots, n_rf, n_gb, acc, 1, st);
      MB_CUDA(cudaGetLastError());
      return;
    }
    launch_trees(ctx, e, cov, C, plane, eg, w, roots, n_rf, n_gb, acc, 0, st);
    started = true;
  }
  if (part == 1) {
    MB_CUDA(cudaGetLastError());
    return;
  }
  if (e->has[MB_V] && e->svm_oct > 0) {           // the handle was created for the tensor-pipe kernel (P <= 8, svm_impl != 2)
    launch_svm_tma(ctx, e, cov, plane, eg, w, acc, started ? 1 : 0, /*chunked=*/true, st);
  } else if (e->has[MB_V]) {
    switch ((e->P + 1) / 2) {
#define MB_SVM_CASE(n) case n: launch_svm<n>(e, cov, plane, eg, w, acc, started ? 1 : 0, st); break;
      MB_SVM_CASE(1) MB_SVM_CASE(2) MB_SVM_CASE(3) MB_SVM_CASE(4) MB_SVM_CASE(5) MB_SVM_CASE(6) MB_SVM_CASE(7)
      MB_SVM_CASE(8)
#undef MB_SVM_CASE
    }
  } else {
    dim3 grid((w.c1 - w.c0 + 31) / 32, (w.r1 - w.r0 + 7) / 8);
    MB_LAUNCH(ctx, "k_ens_smooth", st) k_ens_smooth<<<grid, 256, 0, st>>>(cov, C, plane, eg, w, smooth_params(e), started ? 1 : 0, astride, acc);
  }
  MB_CUDA(cudaGetLastError());
}

// acc / total weight + TPS (spline evaluated in the same pass by the grid-evaluation kernel, or a precomputed
// surface, or none) -> out  (V73:619, 906-907)
void ensemble_finish(mb_ctx* ctx, const mb_ensemble* e, const mb_spline* spline, const double* tps_surface,
                     const mb_window& w, const double* acc, double* out, cudaStream_t st) {
  const double inv_w = 1.0 / e->w_total;
  if (spline) {
    MB_REQUIRE(!tps_surface, "pass either a spline or a precomputed TPS surface, not both");
    const AccFuse fz{acc, acc_stride(w), inv_w};
    tps_eval_fast(ctx, spline, e->g, w, out, w.c1 - w.c0, st, &fz);
    return;
  }
  dim3 grid((w.c1 - w.c0 + 31) / 32, (w.r1 - w.r0 + 7) / 8);
  MB_LAUNCH(ctx, "k_ens_final", st) k_ens_final<<<grid, 256, 0, st>>>(w, acc, acc_stride(w), inv_w, tps_surface, out);
  MB_CUDA(cudaGetLastError());
}

void ensemble_eval(mb_ctx* ctx, const mb_ensemble* e, const float* cov, int C, const mb_spline* spline,
                   const double* tps_surface, const mb_window* wp, double* out, cudaStream_t st) {
  check_window(&e->g, wp);
  const mb_window w = *wp;
  MB_REQUIRE(C == e->C, "number of covariate planes does not match the model descriptors (P = C + 2)");
  double* acc = ctx->arena.take_n<double>((size_t)(acc_stride(w) * acc_rows(w)));
  ensemble_accumulate(ctx, e, cov, C, w, acc, st);
  ensemble_finish(ctx, e, spline, tps_surface, w, acc, out, st);
}

void ensemble_predict_points(mb_ctx* ctx, const mb_ensemble* e, const double* X, int n, double* out_host) {
  if (n <= 0) return;
  cudaStream_t st = ctx->stream;
  ABuf<double> dX(ctx->arena, (size_t)n * e->P), dO(ctx->arena, n);
  dX.upload(X, (size_t)n * e->P, st);
  PointModels pm{};
  pm.sp = smooth_params(e);
  pm.P = e->P;
  pm.rf_ntree = e->has[MB_R] ? e->rf_ntree : 0; pm.rf_nrnodes = e->rf_nrnodes;
  pm.rf_left = e->rfp_left.p; pm.rf_right = e->rfp_right.p; pm.rf_status = e->rfp_status.p;
  pm.rf_var = e->rfp_var.p; pm.rf_split = e->rfp_split.p; pm.rf_pred = e->rfp_pred.p; pm.w_r = e->w[MB_R];
  pm.gb_ntrees = e->has[MB_B] ? e->gb_ntrees : 0; pm.gb_initF = e->gb_initF; pm.gb_off = e->gbp_off.p;
  pm.gb_var = e->gbp_var.p; pm.gb_code = e->gbp_code.p; pm.gb_left = e->gbp_left.p; pm.gb_right = e->gbp_right.p;
  pm.gb_miss = e->gbp_miss.p; pm.w_b = e->w[MB_B];
  pm.S = e->has[MB_V] ? e->svm_S : 0; pm.sv = e->svp_sv.p; pm.alpha = e->svp_alpha.p; pm.xc = e->svm_xc.p;
  pm.xis = e->svm_xis.p; pm.sigma = e->svm_sigma; pm.bias = e->svm_bias; pm.ys = e->svm_ys; pm.yc = e->svm_yc;
  pm.w_v = e->w[MB_V];
  MB_LAUNCH(ctx, "k_ens_points", st) k_ens_points<<<(n + 127) / 128, 128, 0, st>>>(dX.p, n, pm, dO.p);
  MB_CUDA(cudaGetLastError());
  MB_CUDA(cudaMemcpyAsync(out_host, dO.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
}

}  // namespace mb
