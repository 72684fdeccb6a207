// mltps part 3 / part 4 on the device (V73:649-895) and machisplin.tiles.merge (V73:1392-1548):
//   tiles_tps    ceil(n/1500)-px tiling, per-tile knot selection + Tps fit + evaluation on the keep
//                window, then the merge below.
//   tiles_merge  K7: NA-ignoring mean mosaic of the tiles, linear cross-fade on every seam strip
//                (strip = bounding box of the cells where both neighbours are non-NA, fade between the
//                first and last cell centre of that box), mean of all strip blends covering a cell, strips
//                take priority over the tile mean.  One pass over the output raster: each cell reads the
//                <= 4 tiles that cover it; no full-extent per-tile rasters are ever materialised
//                (the reference extends every tile to the full extent, V73:729).
//   gram         K8: G = R'R of the cross-validation residual matrix (V73:329-331 objective).
//   gather_cells part 5 point extraction (V73:910).
#include "common.cuh"

#include <exception>
#include <thread>
#include "internal.h"

#include <algorithm>
#include <cmath>

namespace mb {

static int cround(double x) { return x >= 0 ? (int)std::floor(x + 0.5) : -(int)std::floor(-x + 0.5); }

// terra::crop(rast, ext) as an index window: snap = "near" on the cell edges, clipped to the raster
static mb_window crop_window(const mb_grid& g, double exmin, double exmax, double eymin, double eymax) {
  const double rx = (g.xmax - g.xmin) / g.ncol, ry = (g.ymax - g.ymin) / g.nrow;
  int c0 = cround((exmin - g.xmin) / rx), c1 = cround((exmax - g.xmin) / rx);
  int b0 = cround((eymin - g.ymin) / ry), b1 = cround((eymax - g.ymin) / ry);
  c0 = std::max(0, std::min(g.ncol, c0)); c1 = std::max(0, std::min(g.ncol, c1));
  b0 = std::max(0, std::min(g.nrow, b0)); b1 = std::max(0, std::min(g.nrow, b1));
  return mb_window{g.nrow - b1, g.nrow - b0, c0, c1};
}

// ---------------------------------------------------------------------------------------------
// K7 tile blend
// ---------------------------------------------------------------------------------------------
struct TileDesc {
  const double* p;
  int r0, r1, c0, c1;
};
struct SeamDesc {
  int a, b;        // tile indices: a = west / south tile ("tile 1" of the reference), b = east / north
  int axis;        // 0 = vertical seam (fade in x), 1 = horizontal seam (fade in y)
  int r0, r1, c0, c1;   // window intersection
};

__global__ void __launch_bounds__(256) k_seam_bbox(const TileDesc* __restrict__ tiles,
                                                   const SeamDesc* __restrict__ seams, int* __restrict__ bbox) {
  const SeamDesc s = seams[blockIdx.y];   // an empty window (r0 = r1): nothing to scan, the box keeps its initial value
  const int wc = s.c1 - s.c0, wr = s.r1 - s.r0;
  const int64_t n = (int64_t)wc * wr;
  int rmin = INT_MAX, rmax = INT_MIN, cmin = INT_MAX, cmax = INT_MIN;
  const TileDesc ta = tiles[s.a], tb = tiles[s.b];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = s.r0 + (int)(i / wc), c = s.c0 + (int)(i % wc);
    const double va = ta.p[(int64_t)(r - ta.r0) * (ta.c1 - ta.c0) + (c - ta.c0)];
    const double vb = tb.p[(int64_t)(r - tb.r0) * (tb.c1 - tb.c0) + (c - tb.c0)];
    if (va == va && vb == vb) {
      rmin = min(rmin, r); rmax = max(rmax, r); cmin = min(cmin, c); cmax = max(cmax, c);
    }
  }
  rmin = __reduce_min_sync(0xffffffffu, rmin); rmax = __reduce_max_sync(0xffffffffu, rmax);
  cmin = __reduce_min_sync(0xffffffffu, cmin); cmax = __reduce_max_sync(0xffffffffu, cmax);
  if ((threadIdx.x & 31) == 0 && rmin <= rmax) {
    int* bb = bbox + 4 * blockIdx.y;
    atomicMin(bb + 0, rmin); atomicMax(bb + 1, rmax); atomicMin(bb + 2, cmin); atomicMax(bb + 3, cmax);
  }
}

__global__ void __launch_bounds__(256) k_tile_blend(
    const TileDesc* __restrict__ tiles, int nC, int nR, const int2* __restrict__ colcand,
    const int2* __restrict__ rowcand, const int* __restrict__ vbbox, const int* __restrict__ hbbox,
    double xmin, double ymax, double rx, double ry, mb_window ow, double* __restrict__ out) {
  // ow: the window of the raster this launch writes (the whole raster, or the cells a rank owns); out is ow-shaped
  const int col = ow.c0 + blockIdx.x * 32 + (threadIdx.x & 31);
  const int row = ow.r0 + blockIdx.y * 8 + (threadIdx.x >> 5);
  if (col >= ow.c1 || row >= ow.r1) return;
  const int2 hc = colcand[col];     // tile columns h in [hc.x, hc.y]
  const int2 jc = rowcand[row];     // tile rows    j in [jc.x, jc.y]  (j counts from the south)
  double v[3][3];
  bool ok[3][3];
  double tsum = 0.0;
  int tcnt = 0;
#pragma unroll
  for (int dj = 0; dj < 3; ++dj)
#pragma unroll
    for (int dh = 0; dh < 3; ++dh) {
      ok[dj][dh] = false;
      v[dj][dh] = 0.0;
      const int j = jc.x + dj, h = hc.x + dh;
      if (j <= jc.y && h <= hc.y) {
        const TileDesc t = tiles[j * nC + h];
        if (row >= t.r0 && row < t.r1 && col >= t.c0 && col < t.c1) {
          const double x = t.p[(int64_t)(row - t.r0) * (t.c1 - t.c0) + (col - t.c0)];
          if (x == x) { ok[dj][dh] = true; v[dj][dh] = x; tsum += x; ++tcnt; }
        }
      }
    }
  double ssum = 0.0;
  int scnt = 0;
  // vertical seams (tile (j,h) | (j,h+1)), V73:764-806
#pragma unroll
  for (int dj = 0; dj < 3; ++dj)
#pragma unroll
    for (int dh = 0; dh < 2; ++dh) {
      const int j = jc.x + dj, h = hc.x + dh;
      if (j <= jc.y && h + 1 <= hc.y && ok[dj][dh] && ok[dj][dh + 1]) {
        const int* bb = vbbox + 4 * (j * (nC - 1) + h);
        if (row >= bb[0] && row <= bb[1] && col >= bb[2] && col <= bb[3]) {
          const double x = xmin + (col + 0.5) * rx;
          const double x0 = xmin + (bb[2] + 0.5) * rx, x1 = xmin + (bb[3] + 0.5) * rx;
          const double t = (x - x0) / (x1 - x0);
          const double f = v[dj][dh + 1] * t + v[dj][dh] * (1.0 - t);
          if (f == f) { ssum += f; ++scnt; }
        }
      }
    }
  // horizontal seams (tile (j,h) south | (j+1,h) north), V73:815-877
#pragma unroll
  for (int dj = 0; dj < 2; ++dj)
#pragma unroll
    for (int dh = 0; dh < 3; ++dh) {
      const int j = jc.x + dj, h = hc.x + dh;
      if (j + 1 <= jc.y && h <= hc.y && ok[dj][dh] && ok[dj + 1][dh]) {
        const int* bb = hbbox + 4 * (j * nC + h);
        if (row >= bb[0] && row <= bb[1] && col >= bb[2] && col <= bb[3]) {
          const double y = ymax - (row + 0.5) * ry;
          const double y0 = ymax - (bb[1] + 0.5) * ry, y1 = ymax - (bb[0] + 0.5) * ry;   // min / max latitude
          const double t = (y - y0) / (y1 - y0);
          const double f = v[dj + 1][dh] * t + v[dj][dh] * (1.0 - t);
          if (f == f) { ssum += f; ++scnt; }
        }
      }
    }
  double r;
  if (scnt > 0) r = ssum / scnt;
  else if (tcnt > 0) r = tsum / tcnt;
  else r = __longlong_as_double(0x7ff8000000000000LL);
  out[(int64_t)(row - ow.r0) * (ow.c1 - ow.c0) + (col - ow.c0)] = r;
}

// Host side of the blend: which tile columns / rows can cover a raster column / row, and the seams of the lattice
struct MergeLattice {
  std::vector<int2> colcand, rowcand;
  std::vector<SeamDesc> seams;      // nR (nC - 1) vertical seams, then (nR - 1) nC horizontal ones
  int nv = 0;
};
static MergeLattice merge_lattice(const mb_grid& g, int nC, int nR, const mb_window* wins) {
  MergeLattice L;
  const int nt = nC * nR;
  for (int t = 0; t < nt; ++t) check_window(&g, &wins[t]);
  // candidate tile columns / rows of every raster column / row (windows are monotone along the lattice)
  L.colcand.assign(g.ncol, make_int2(0, -1));
  L.rowcand.assign(g.nrow, make_int2(0, -1));
  for (int h = 0; h < nC; ++h) {
    int lo = g.ncol, hi = 0;
    for (int j = 0; j < nR; ++j) { lo = std::min(lo, wins[j * nC + h].c0); hi = std::max(hi, wins[j * nC + h].c1); }
    for (int c = lo; c < hi; ++c) {
      if (L.colcand[c].y < L.colcand[c].x) L.colcand[c] = make_int2(h, h);
      else L.colcand[c].y = h;
    }
  }
  for (int j = 0; j < nR; ++j) {
    int lo = g.nrow, hi = 0;
    for (int h = 0; h < nC; ++h) { lo = std::min(lo, wins[j * nC + h].r0); hi = std::max(hi, wins[j * nC + h].r1); }
    for (int r = lo; r < hi; ++r) {
      if (L.rowcand[r].y < L.rowcand[r].x) L.rowcand[r] = make_int2(j, j);
      else L.rowcand[r].y = j;
    }
  }
  for (auto& c : L.colcand) MB_REQUIRE(c.y - c.x <= 2, "more than three tile columns overlap on one raster column");
  for (auto& r : L.rowcand) MB_REQUIRE(r.y - r.x <= 2, "more than three tile rows overlap on one raster row");
  auto add_seam = [&](int a, int b, int axis) {
    SeamDesc s{a, b, axis, std::max(wins[a].r0, wins[b].r0), std::min(wins[a].r1, wins[b].r1),
               std::max(wins[a].c0, wins[b].c0), std::min(wins[a].c1, wins[b].c1)};
    if (s.r1 <= s.r0 || s.c1 <= s.c0) { s.r0 = s.r1 = s.c0 = s.c1 = 0; }
    L.seams.push_back(s);
  };
  L.nv = nR * (nC - 1);
  for (int j = 0; j < nR; ++j)
    for (int h = 0; h + 1 < nC; ++h) add_seam(j * nC + h, j * nC + h + 1, 0);
  for (int j = 0; j + 1 < nR; ++j)
    for (int h = 0; h < nC; ++h) add_seam(j * nC + h, (j + 1) * nC + h, 1);
  return L;
}
static std::vector<int> bbox_initial(size_t nseams) {
  std::vector<int> b(4 * std::max<size_t>(1, nseams));
  for (size_t s = 0; s < b.size() / 4; ++s) {
    b[4 * s + 0] = INT_MAX; b[4 * s + 1] = INT_MIN;
    b[4 * s + 2] = INT_MAX; b[4 * s + 3] = INT_MIN;
  }
  return b;
}

void tiles_merge(mb_ctx* ctx, const mb_grid& g, int nC, int nR, const mb_window* wins,
                 const double* const* tiles_dev, double* out_dev, cudaStream_t st) {
  const int nt = nC * nR;
  const MergeLattice L = merge_lattice(g, nC, nR, wins);
  std::vector<TileDesc> td(nt);
  for (int t = 0; t < nt; ++t) td[t] = TileDesc{tiles_dev[t], wins[t].r0, wins[t].r1, wins[t].c0, wins[t].c1};
  Arena& ar = ctx->arena;
  ABuf<TileDesc> d_tiles(ar);
  ABuf<SeamDesc> d_seams(ar);
  ABuf<int2> d_col(ar), d_row(ar);
  ABuf<int> d_bbox(ar, (size_t)4 * std::max<size_t>(1, L.seams.size()));
  d_tiles.upload(td, st);
  d_col.upload(L.colcand, st);
  d_row.upload(L.rowcand, st);
  const std::vector<int> bbox_init = bbox_initial(L.seams.size());
  d_bbox.upload(bbox_init, st);
  if (!L.seams.empty()) {
    d_seams.upload(L.seams, st);
    MB_LAUNCH(ctx, "k_seam_bbox", st) k_seam_bbox<<<dim3(64, (unsigned)L.seams.size()), 256, 0, st>>>(d_tiles.p, d_seams.p, d_bbox.p);
  }
  const double rx = (g.xmax - g.xmin) / g.ncol, ry = (g.ymax - g.ymin) / g.nrow;
  dim3 grid((g.ncol + 31) / 32, (g.nrow + 7) / 8);
  MB_LAUNCH(ctx, "k_tile_blend", st) k_tile_blend<<<grid, 256, 0, st>>>(d_tiles.p, nC, nR, d_col.p, d_row.p, d_bbox.p, d_bbox.p + 4 * (size_t)L.nv,
                                     g.xmin, g.ymax, rx, ry, mb_window{0, g.nrow, 0, g.ncol}, out_dev);
  MB_CUDA(cudaGetLastError());
  MB_CUDA(cudaStreamSynchronize(st));   // descriptor uploads come from host vectors that die here
}

// ---------------------------------------------------------------------------------------------
// machisplin.tiles.merge across GPUs (the tile-border blend of the multi-GPU path).
// Tile t of the nC x nR lattice lives on rank t % size.  The raster is PARTITIONED into one owned window per tile - the
// tile's share, cut through the middle of every overlap zone - and every rank blends the cells its tiles own, so the merged
// raster stays distributed and no rank ever sees a whole neighbour tile:
//   1. for each pair (source tile a, destination tile b) on different ranks with W_a intersecting own_b, rank(a) packs that
//      rectangle (cudaMemcpy2DAsync) and sends it to rank(b) - ncclSend / ncclRecv inside one group; with the reference's
//      feather.d = 50 these are strips 25 cells wide (a few MB per seam, against GBs for the tiles themselves);
//   2. every rank scans the seams inside its owned windows for the cells where both neighbours are non-NA; the bounding boxes
//      of the seam strips (V73:772-779) are the min / max over all ranks: ONE ncclAllReduce(min) of 4 integers per seam;
//   3. k_tile_blend runs on the owned windows with the received strips standing in for the neighbour tiles.
// Every cell sees the same tile values and the same boxes as in tiles_merge, so the result is bit-identical to it.
// ---------------------------------------------------------------------------------------------
mb_window tiles_owned_window(const mb_grid& g, int nC, int nR, const mb_window* wins, int t) {
  MB_REQUIRE(nC >= 1 && nR >= 1 && t >= 0 && t < nC * nR, "tile index out of range");
  const int h = t % nC, j = t / nC;
  for (int jj = 0; jj < nR; ++jj)
    MB_REQUIRE(wins[jj * nC + h].c0 == wins[t].c0 && wins[jj * nC + h].c1 == wins[t].c1,
               "sharded merge needs a regular tile lattice: the tiles of one column share their column range");
  for (int hh = 0; hh < nC; ++hh)
    MB_REQUIRE(wins[j * nC + hh].r0 == wins[t].r0 && wins[j * nC + hh].r1 == wins[t].r1,
               "sharded merge needs a regular tile lattice: the tiles of one row share their row range");
  mb_window o;
  o.c0 = h == 0 ? 0 : (wins[t - 1].c1 + wins[t].c0) / 2;
  o.c1 = h == nC - 1 ? g.ncol : (wins[t].c1 + wins[t + 1].c0) / 2;
  o.r1 = j == 0 ? g.nrow : (wins[t - nC].r0 + wins[t].r1) / 2;          // tile rows count from the south
  o.r0 = j == nR - 1 ? 0 : (wins[t].r0 + wins[t + nC].r1) / 2;
  MB_REQUIRE(o.c0 < o.c1 && o.r0 < o.r1, "a tile owns no cell: the windows do not form a lattice");
  return o;
}

static mb_window intersect(const mb_window& a, const mb_window& b) {
  mb_window w{std::max(a.r0, b.r0), std::min(a.r1, b.r1), std::max(a.c0, b.c0), std::min(a.c1, b.c1)};
  if (w.r1 <= w.r0 || w.c1 <= w.c0) w = mb_window{0, 0, 0, 0};
  return w;
}
static bool empty(const mb_window& w) { return w.r1 <= w.r0 || w.c1 <= w.c0; }

__global__ void k_bbox_flip(int* bbox, int n4) {   // (min, max, min, max) <-> (min, -max, min, -max) so that ONE min-reduction serves both
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4 || (i & 1) == 0) return;
  const int v = bbox[i];
  bbox[i] = v == INT_MIN ? INT_MAX : (v == INT_MAX ? INT_MIN : -v);
}

void tiles_merge_shard(mb_ctx* ctx, const mb_grid& g, int nC, int nR, const mb_window* wins, const double* const* my_tiles_dev,
                       double* const* out_dev, cudaStream_t st) {
  const int nt = nC * nR;
  const int size = ctx->comm ? ctx->comm_size : 1, me = ctx->comm ? ctx->comm_rank : 0;
  const MergeLattice L = merge_lattice(g, nC, nR, wins);
  std::vector<mb_window> own(nt);
  for (int t = 0; t < nt; ++t) own[t] = tiles_owned_window(g, nC, nR, wins, t);
  auto rank_of = [&](int t) { return t % size; };
  for (int t = 0; t < nt; ++t)
    if (rank_of(t) == me) MB_REQUIRE(my_tiles_dev[t] && out_dev[t], "a tile of this rank (t % size == rank) has no raster or no output buffer");
  Arena& ar = ctx->arena;
  // ---- 1. strips: (source tile a -> owned window of tile b), same enumeration on every rank --------------------------
  std::vector<std::vector<TileDesc>> view(nt);          // view[b][a]: what the blend of own_b reads for tile a (my tiles b only)
  std::vector<CommMsg> msgs;
  for (int b = 0; b < nt; ++b)
    if (rank_of(b) == me) view[b].assign(nt, TileDesc{nullptr, 0, 0, 0, 0});
  for (int a = 0; a < nt; ++a)
    for (int b = 0; b < nt; ++b) {
      const mb_window I = intersect(wins[a], own[b]);
      if (empty(I)) continue;
      const bool mine_a = rank_of(a) == me, mine_b = rank_of(b) == me;
      if (mine_a && mine_b) {                            // both here: the tile itself serves
        view[b][a] = TileDesc{my_tiles_dev[a], wins[a].r0, wins[a].r1, wins[a].c0, wins[a].c1};
        continue;
      }
      if (!mine_a && !mine_b) continue;
      const size_t cnt = (size_t)(I.r1 - I.r0) * (I.c1 - I.c0);
      double* buf = ar.take_n<double>(cnt);
      if (mine_a) {
        const int wa = wins[a].c1 - wins[a].c0;
        MB_CUDA(cudaMemcpy2DAsync(buf, sizeof(double) * (I.c1 - I.c0),
                                  my_tiles_dev[a] + (size_t)(I.r0 - wins[a].r0) * wa + (I.c0 - wins[a].c0), sizeof(double) * wa,
                                  sizeof(double) * (I.c1 - I.c0), (size_t)(I.r1 - I.r0), cudaMemcpyDeviceToDevice, st));
        msgs.push_back(CommMsg{rank_of(b), buf, cnt, true});
      } else {
        view[b][a] = TileDesc{buf, I.r0, I.r1, I.c0, I.c1};
        msgs.push_back(CommMsg{rank_of(a), buf, cnt, false});
      }
    }
  comm_exchange_f64(ctx, msgs, st);
  // ---- 2. seam boxes: partial over the owned windows, then min / max over the ranks ------------------------------------
  ABuf<int2> d_col(ar), d_row(ar);
  d_col.upload(L.colcand, st);
  d_row.upload(L.rowcand, st);
  const size_t ns = L.seams.size();
  ABuf<int> d_bbox(ar, (size_t)4 * std::max<size_t>(1, ns));
  const std::vector<int> bbox_init = bbox_initial(ns);
  d_bbox.upload(bbox_init, st);
  std::vector<std::vector<SeamDesc>> clipped(nt);       // kept alive until the stream has consumed the uploads
  std::vector<TileDesc*> d_views(nt, nullptr);
  for (int b = 0; b < nt; ++b) {
    if (rank_of(b) != me) continue;
    d_views[b] = ar.upload(view[b].data(), view[b].size(), st);
    if (ns == 0) continue;
    clipped[b] = L.seams;
    for (SeamDesc& s : clipped[b]) {
      const mb_window I = intersect(mb_window{s.r0, s.r1, s.c0, s.c1}, own[b]);
      // both tiles of a seam that reaches into own_b have a strip here (their windows contain the seam window)
      s.r0 = I.r0; s.r1 = I.r1; s.c0 = I.c0; s.c1 = I.c1;
    }
    SeamDesc* d_seams = ar.upload(clipped[b].data(), clipped[b].size(), st);
    MB_LAUNCH(ctx, "k_seam_bbox", st) k_seam_bbox<<<dim3(64, (unsigned)ns), 256, 0, st>>>(d_views[b], d_seams, d_bbox.p);
  }
  if (ns > 0 && ctx->comm) {
    const int n4 = (int)(4 * ns);
    k_bbox_flip<<<(n4 + 255) / 256, 256, 0, st>>>(d_bbox.p, n4);
    comm_allreduce_i32_min(ctx, d_bbox.p, n4, st);
    k_bbox_flip<<<(n4 + 255) / 256, 256, 0, st>>>(d_bbox.p, n4);
  }
  // ---- 3. blend the owned windows ------------------------------------------------------------------------------------------
  const double rx = (g.xmax - g.xmin) / g.ncol, ry = (g.ymax - g.ymin) / g.nrow;
  for (int b = 0; b < nt; ++b) {
    if (rank_of(b) != me) continue;
    const mb_window o = own[b];
    dim3 grid((o.c1 - o.c0 + 31) / 32, (o.r1 - o.r0 + 7) / 8);
    MB_LAUNCH(ctx, "k_tile_blend", st) k_tile_blend<<<grid, 256, 0, st>>>(d_views[b], nC, nR, d_col.p, d_row.p, d_bbox.p, d_bbox.p + 4 * (size_t)L.nv,
                                       g.xmin, g.ymax, rx, ry, o, out_dev[b]);
  }
  MB_CUDA(cudaGetLastError());
  MB_CUDA(cudaStreamSynchronize(st));   // descriptor uploads come from host vectors that die here
}

// ---------------------------------------------------------------------------------------------
// mltps part 3: tiling + per-tile fit + evaluation
// ---------------------------------------------------------------------------------------------
__global__ void k_fill(double* p, int64_t n, double v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

void tiles_tps(mb_ctx* ctx, const mb_grid& g, const double* knots_xy, const double* resid, int n, int tile_px,
               double fit_halo, double keep_halo, int min_pts, double lambda, int method, double* out_dev,
               cudaStream_t st) {
  const int nRx = (g.nrow + tile_px - 1) / tile_px;       // V73:656-657
  const int nCx = (g.ncol + tile_px - 1) / tile_px;       // V73:660-661
  const mb_window full{0, g.nrow, 0, g.ncol};
  auto eval = [&](const mb_spline* s, const mb_window& w, double* dst) {
    if (method == MB_EVAL_DIRECT) tps_eval_direct(ctx, s, g, w, dst, w.c1 - w.c0, st);
    else tps_eval_fast(ctx, s, g, w, dst, w.c1 - w.c0, st);
  };
  if (nRx * nCx == 1) {                                    // V73:748-753
    mb_spline* sp = nullptr;
    tps_fit(ctx, knots_xy, resid, n, 1, lambda, &sp);
    std::unique_ptr<mb_spline> hold(sp);
    eval(sp, full, out_dev);
    MB_CUDA(cudaStreamSynchronize(st));
    return;
  }
  const double longDist = (g.xmax - g.xmin) / nCx, latDist = (g.ymax - g.ymin) / nRx;
  const double rx = (g.xmax - g.xmin) / g.ncol, ry = (g.ymax - g.ymin) / g.nrow;
  // cell of every knot (terra::extract: the containing cell)
  std::vector<int> krow(n), kcol(n);
  for (int i = 0; i < n; ++i) {
    const double x = knots_xy[i], y = knots_xy[(size_t)n + i];
    int c = (int)std::floor((x - g.xmin) / rx), r = (int)std::floor((g.ymax - y) / ry);
    if (x == g.xmax) c = g.ncol - 1;
    if (y == g.ymin) r = g.nrow - 1;
    if (x < g.xmin || x > g.xmax || y < g.ymin || y > g.ymax) c = r = -1;
    kcol[i] = c; krow[i] = r;
  }
  const int nt = nRx * nCx;
  std::vector<mb_window> keep(nt);
  std::vector<double*> bufs(nt);
  std::vector<const double*> ptrs(nt);
  // ---- per tile: windows, knots, output buffer (host arithmetic only) ---------------------------------
  struct TileJob { std::vector<double> xy, y; int m = 0; };
  std::vector<TileJob> jobs(nt);
  int m_max = 0;
  for (int j = 1; j <= nRx; ++j)
    for (int h = 1; h <= nCx; ++h) {
      const int t = (j - 1) * nCx + (h - 1);
      const mb_window fw = crop_window(g, g.xmin + ((longDist * (h - 1)) - (longDist * fit_halo)),
                                       g.xmin + ((longDist * h) + (longDist * fit_halo)),
                                       g.ymin + ((latDist * (j - 1))) - (latDist * fit_halo),
                                       g.ymin + ((latDist * j)) + (latDist * fit_halo));      // V73:673
      mb_window kw = crop_window(g, g.xmin + ((longDist * (h - 1)) - (longDist * keep_halo)),
                                 g.xmin + ((longDist * h) + (longDist * keep_halo)),
                                 g.ymin + ((latDist * (j - 1))) - (latDist * keep_halo),
                                 g.ymin + ((latDist * j)) + (latDist * keep_halo));           // V73:680
      kw.r0 = std::max(kw.r0, fw.r0); kw.r1 = std::min(kw.r1, fw.r1);
      kw.c0 = std::max(kw.c0, fw.c0); kw.c1 = std::min(kw.c1, fw.c1);
      MB_REQUIRE(kw.r1 > kw.r0 && kw.c1 > kw.c0, "empty tile window");
      keep[t] = kw;
      std::vector<double> tx, tyy, ty;
      for (int i = 0; i < n; ++i)
        if (krow[i] >= fw.r0 && krow[i] < fw.r1 && kcol[i] >= fw.c0 && kcol[i] < fw.c1) {   // V73:699-706
          tx.push_back(knots_xy[i]); tyy.push_back(knots_xy[(size_t)n + i]); ty.push_back(resid[i]);
        }
      TileJob& job = jobs[t];
      job.m = (int)ty.size();
      m_max = std::max(m_max, job.m);
      job.xy.resize((size_t)2 * job.m);
      std::copy(tx.begin(), tx.end(), job.xy.begin());
      std::copy(tyy.begin(), tyy.end(), job.xy.begin() + job.m);
      job.y = std::move(ty);
      const size_t cells = (size_t)(kw.r1 - kw.r0) * (kw.c1 - kw.c0);
      bufs[t] = ctx->arena.take_n<double>(cells);
      ptrs[t] = bufs[t];
    }
  // ---- fit + evaluate every tile (V73:710-728) -----------------------------------------------------------
  auto run_tile = [&](mb_ctx* c, cudaStream_t cs, int t) {
    const TileJob& job = jobs[t];
    const mb_window& kw = keep[t];
    const size_t cells = (size_t)(kw.r1 - kw.r0) * (kw.c1 - kw.c0);
    if (job.m < min_pts) {                                                                   // V73:710-721
      MB_LAUNCH(c, "k_fill", cs) k_fill<<<256, 256, 0, cs>>>(bufs[t], (int64_t)cells, 0.0);
      return;
    }
    mb_spline* sp = nullptr;
    tps_fit(c, job.xy.data(), job.y.data(), job.m, 1, lambda, &sp);                           // V73:722
    std::unique_ptr<mb_spline> hold(sp);
    if (method == MB_EVAL_DIRECT) tps_eval_direct(c, sp, g, kw, bufs[t], kw.c1 - kw.c0, cs);   // V73:726-728
    else tps_eval_fast(c, sp, g, kw, bufs[t], kw.c1 - kw.c0, cs);
    MB_CUDA(cudaStreamSynchronize(cs));
  };
  // Tiles are independent: up to eight worker lanes (own stream + scratch arena) keep the device busy while one lane
  // sits in a host synchronisation of its fit.  Only for small fits: their persistent kernels use small grids, so
  // several are co-resident; a large fit takes the whole device and runs alone.
  // lanes: as many as keep every fit's persistent grid co-resident (the grid of k_sytrd follows the matrix size:
  // row chunks of 32, strip segments of 4 tiles; 2 CTAs per SM fit on the device)
  const int nchunk_max = (m_max + 31) / 32, ntile_max = (m_max + 63) / 64;
  const int g_est = std::max(8, std::max(nchunk_max, (ntile_max / 4 + 1) * (2 * (ntile_max / 4) + ntile_max % 4)));
  const int kLanes = std::max(1, std::min(std::min(8, nt), (2 * ctx->sm_count) / g_est));
  const bool concurrent = kLanes >= 2 && ctx->eigen_impl == 0 && st == ctx->stream;
  if (!concurrent) {
    for (int t = 0; t < nt; ++t) run_tile(ctx, st, t);
  } else {
    while ((int)ctx->lanes.size() < kLanes) {
      auto w = std::make_unique<mb_ctx>();
      w->device = ctx->device;
      w->sm_count = ctx->sm_count;
      int prio_lo = 0, prio_hi = 0;
      MB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      MB_CUDA(cudaStreamCreateWithPriority(&w->stream, cudaStreamNonBlocking, prio_hi));
      ctx->lanes.push_back(std::move(w));
    }
    std::vector<std::exception_ptr> errs(kLanes);
    std::vector<std::thread> threads;
    for (int l = 0; l < kLanes; ++l) {
      mb_ctx* w = ctx->lanes[l].get();
      w->timing = ctx->timing;
      w->cheb_p = ctx->cheb_p; w->leaf_cols = ctx->leaf_cols; w->leaf_rows = ctx->leaf_rows;
      w->eval_precision = ctx->eval_precision; w->sytrd_mode = ctx->sytrd_mode; w->sytrd_ctas_per_sm = ctx->sytrd_ctas_per_sm;
      w->logtab.p = ctx->logtab.p; w->logtab.n = ctx->logtab.n;        // borrowed (read-only table)
      threads.emplace_back([&, l, w] {
        try {
          MB_CUDA(cudaSetDevice(w->device));
          for (int t = l; t < nt; t += kLanes) {
            w->arena.begin(w->stream);
            run_tile(w, w->stream, t);
          }
          MB_CUDA(cudaStreamSynchronize(w->stream));
        } catch (...) {
          errs[l] = std::current_exception();
        }
      });
    }
    for (auto& th : threads) th.join();
    for (int l = 0; l < kLanes; ++l) {
      mb_ctx* w = ctx->lanes[l].get();
      w->logtab.p = nullptr; w->logtab.n = 0;                          // give the borrowed table back
      ctx->launches += w->launches; w->launches = 0;
      ctx->timed.insert(ctx->timed.end(), w->timed.begin(), w->timed.end());
      w->timed.clear();
    }
    for (auto& e : errs) if (e) std::rethrow_exception(e);
  }
  tiles_merge(ctx, g, nCx, nRx, keep.data(), ptrs.data(), out_dev, st);                       // V73:739-895
}

// ---------------------------------------------------------------------------------------------
// K8 gram: G = R'R, R n x K column-major, K <= 8
// ---------------------------------------------------------------------------------------------
constexpr int kGramBlocks = 296;

__global__ void __launch_bounds__(256) k_gram_partial(const double* __restrict__ R, int n, int K,
                                                      double* __restrict__ part) {
  double acc[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) acc[i] = 0.0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    double x[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) x[a] = a < K ? R[(size_t)a * n + r] : 0.0;
    int q = 0;
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b) { acc[q] = fma(x[a], x[b], acc[q]); ++q; }
  }
  __shared__ double s[8][36];
#pragma unroll
  for (int i = 0; i < 36; ++i) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 36) {
    double v = 0.0;
    for (int wq = 0; wq < 8; ++wq) v += s[wq][threadIdx.x];
    part[(size_t)blockIdx.x * 36 + threadIdx.x] = v;
  }
}
__global__ void k_gram_final(const double* __restrict__ part, int nblocks, int K, double* __restrict__ G) {
  const int t = threadIdx.x;
  if (t >= 36) return;
  double v = 0.0;
  for (int b = 0; b < nblocks; ++b) v += part[(size_t)b * 36 + t];
  int a = 0, q = t;
  while (q > a) { q -= a + 1; ++a; }   // t = a(a+1)/2 + b
  const int b = q;
  if (a < K && b < K) { G[(size_t)b * K + a] = v; G[(size_t)a * K + b] = v; }
}

void gram(mb_ctx* ctx, const double* R_dev, int n, int K, double* G_dev, cudaStream_t st) {
  ABuf<double> part(ctx->arena, (size_t)kGramBlocks * 36);
  const int blocks = std::min(kGramBlocks, (n + 255) / 256);
  MB_LAUNCH(ctx, "k_gram_partial", st) k_gram_partial<<<blocks, 256, 0, st>>>(R_dev, n, K, part.p);
  MB_LAUNCH(ctx, "k_gram_final", st) k_gram_final<<<1, 64, 0, st>>>(part.p, blocks, K, G_dev);
  MB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// part 5: f.actual <- extract(final, points)
// ---------------------------------------------------------------------------------------------
__global__ void k_gather(const double* __restrict__ ras, int64_t stride, int nrow, int ncol, const int* __restrict__ row,
                         const int* __restrict__ col, int n, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = row[i], c = col[i];
  // terra::extract gives NA for a point outside the raster
  out[i] = ((unsigned)r < (unsigned)nrow && (unsigned)c < (unsigned)ncol) ? ras[(int64_t)r * stride + c]
                                                                          : __longlong_as_double(0x7ff8000000000000LL);
}

void gather_cells(mb_ctx* ctx, const double* raster_dev, int64_t row_stride, int nrow, int ncol, const int32_t* row,
                  const int32_t* col, int n, double* out_host, cudaStream_t st) {
  if (n <= 0) return;
  ABuf<int> dr(ctx->arena, n), dc(ctx->arena, n);
  ABuf<double> dout(ctx->arena, n);
  dr.upload(row, n, st);
  dc.upload(col, n, st);
  MB_LAUNCH(ctx, "k_gather", st) k_gather<<<(n + 255) / 256, 256, 0, st>>>(raster_dev, row_stride, nrow, ncol, dr.p, dc.p, n, dout.p);
  MB_CUDA(cudaGetLastError());
  MB_CUDA(cudaMemcpyAsync(out_host, dout.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
}

}  // namespace mb
