// Multi-GPU plumbing behind the C ABI: one process per GPU, one NCCL communicator per context.
//
// The path shards by rows / tiles with no raster exchange: the only data that crosses NVLink is
//   - the K x K Gram of the cross-validation residuals (V73:329-333) - ncclAllReduce of <= 64 doubles,
//   - the descriptor of a fitted spline (knots, c, d, centre / scale: 24 bytes per knot) - ncclBroadcast from the rank that
//     ran fields::Tps to the ranks that evaluate it on their own rows,
//   - per-tile spline descriptors in tiled mode - ncclAllGather.
// NCCL is bound at run time (dlopen): the host process decides which libnccl is in the address space (an R session: the
// system library; a torchrun-launched Python host: the copy torch already mapped), and a single-GPU host needs none at all.
#include "common.cuh"
#include "internal.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <mutex>

namespace mb {
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string source;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    const char* names[4] = {std::getenv("MB_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
    // first: a copy that is already mapped into the process (never two NCCLs in one address space)
    for (int pass = 0; pass < 2 && !api.handle; ++pass)
      for (const char* n : names) {
        if (!n || !*n) continue;
        api.handle = dlopen(n, pass == 0 ? (RTLD_NOW | RTLD_NOLOAD) : (RTLD_NOW | RTLD_LOCAL));
        if (api.handle) { api.source = std::string(n) + (pass == 0 ? " (already loaded by the host process)" : ""); break; }
      }
    if (!api.handle) { err = std::string("libnccl.so.2 not found (set MB_NCCL_LIB): ") + (dlerror() ? dlerror() : ""); return; }
    auto sym = [&](const char* s) {
      void* p = dlsym(api.handle, s);
      if (!p) err += std::string(" missing symbol ") + s;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  });
  if (!err.empty()) throw Error(MB_E_UNSUPPORTED, "NCCL: " + err);
  return api;
}

#define MB_NCCL(expr)                                                                                            \
  do {                                                                                                           \
    ncclResult_t _r = (expr);                                                                                    \
    if (_r != ncclSuccess)                                                                                       \
      throw mb::Error(MB_E_CUDA, std::string(#expr) + ": " + nccl().GetErrorString(_r));                         \
  } while (0)

ncclComm_t comm_of(mb_ctx* ctx) {
  MB_REQUIRE(ctx->comm != nullptr, "no communicator: call mb_comm_init first");
  return static_cast<ncclComm_t>(ctx->comm);
}

}  // namespace

void comm_release(mb_ctx* ctx) {
  if (ctx->comm) {
    nccl().CommDestroy(static_cast<ncclComm_t>(ctx->comm));
    ctx->comm = nullptr;
    ctx->comm_rank = 0;
    ctx->comm_size = 1;
  }
}

void comm_allreduce_f64(mb_ctx* ctx, double* dev, int n, int op, cudaStream_t st) {
  if (!ctx->comm) return;                     // no communicator = one process; a 1-rank communicator still goes through NCCL
  ctx->launches++;
  MB_NCCL(nccl().AllReduce(dev, dev, (size_t)n, ncclDouble, op == 1 ? ncclMax : ncclSum, comm_of(ctx), st));
}

void comm_allgather_f64(mb_ctx* ctx, const double* send_dev, double* recv_dev, size_t count_per_rank, cudaStream_t st) {
  if (!ctx->comm) {
    if (send_dev != recv_dev)
      MB_CUDA(cudaMemcpyAsync(recv_dev, send_dev, count_per_rank * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return;
  }
  ctx->launches++;
  MB_NCCL(nccl().AllGather(send_dev, recv_dev, count_per_rank, ncclDouble, comm_of(ctx), st));
}

void comm_allreduce_i32_min(mb_ctx* ctx, int* dev, int n, cudaStream_t st) {
  if (!ctx->comm) return;
  ctx->launches++;
  MB_NCCL(nccl().AllReduce(dev, dev, (size_t)n, ncclInt32, ncclMin, comm_of(ctx), st));
}

// Point-to-point exchange of float64 blocks (the seam strips of machisplin.tiles.merge): every message of the list is posted
// inside ONE NCCL group, so the call cannot deadlock whatever the order; messages between the same pair of ranks are matched
// in list order, which both sides derive from the same (source tile, destination tile) enumeration.
void comm_exchange_f64(mb_ctx* ctx, const std::vector<CommMsg>& msgs, cudaStream_t st) {
  if (msgs.empty()) return;
  MB_REQUIRE(ctx->comm != nullptr, "no communicator: call mb_comm_init first");
  MB_NCCL(nccl().GroupStart());
  for (const CommMsg& m : msgs) {
    ctx->launches++;
    ncclResult_t r = m.send ? nccl().Send(m.ptr, m.count, ncclDouble, m.peer, comm_of(ctx), st)
                            : nccl().Recv(m.ptr, m.count, ncclDouble, m.peer, comm_of(ctx), st);
    if (r != ncclSuccess) {
      nccl().GroupEnd();
      throw Error(MB_E_CUDA, std::string("ncclSend / ncclRecv: ") + nccl().GetErrorString(r));
    }
  }
  MB_NCCL(nccl().GroupEnd());
}

// Spline descriptor on the wire: 16 header doubles + 3 * cap payload doubles (kx | ky | c, each cap long, np <= cap used).
constexpr int kHdr = 16;
size_t spline_wire_doubles(int cap) { return kHdr + 3 * (size_t)cap; }

void spline_pack(const mb_spline* s, int cap, double* w) {
  MB_REQUIRE(s->np <= cap, "spline has more knots than the broadcast capacity");
  std::fill(w, w + spline_wire_doubles(cap), 0.0);
  w[0] = s->np; w[1] = s->lambda; w[2] = s->eff_df; w[3] = s->gcv;
  w[4] = s->center[0]; w[5] = s->center[1]; w[6] = s->scale[0]; w[7] = s->scale[1];
  w[8] = s->d[0]; w[9] = s->d[1]; w[10] = s->d[2]; w[11] = s->fscale; w[12] = 1.0 /* valid */;
  std::copy(s->kx.begin(), s->kx.end(), w + kHdr);
  std::copy(s->ky.begin(), s->ky.end(), w + kHdr + cap);
  std::copy(s->c.begin(), s->c.end(), w + kHdr + 2 * (size_t)cap);
}

mb_spline* spline_unpack(mb_ctx* ctx, const double* w, int cap) {
  if (w[12] != 1.0) return nullptr;                    // "no spline" marker (e.g. a tile with < min_pts knots)
  const int np = (int)w[0];
  MB_REQUIRE(np >= 1 && np <= cap, "corrupt spline descriptor");
  auto s = std::make_unique<mb_spline>();
  s->np = np;
  s->lambda = w[1]; s->eff_df = w[2]; s->gcv = w[3];
  s->center[0] = w[4]; s->center[1] = w[5]; s->scale[0] = w[6]; s->scale[1] = w[7];
  s->d[0] = w[8]; s->d[1] = w[9]; s->d[2] = w[10];
  s->kx.assign(w + kHdr, w + kHdr + np);
  s->ky.assign(w + kHdr + cap, w + kHdr + cap + np);
  s->c.assign(w + kHdr + 2 * (size_t)cap, w + kHdr + 2 * (size_t)cap + np);
  s->sx.resize(np); s->sy.resize(np);
  for (int i = 0; i < np; ++i) {
    s->sx[i] = (s->kx[i] - s->center[0]) / s->scale[0];
    s->sy[i] = (s->ky[i] - s->center[1]) / s->scale[1];
  }
  spline_finalize(ctx, s.get(), w[11]);
  return s.release();
}

// Root sends *s (may be NULL on the other ranks); every other rank receives a new handle.  cap >= np on every rank.
// Runs on ctx->stream and synchronises it (the receiver needs the descriptor on the host to build its handle).
mb_spline* spline_bcast(mb_ctx* ctx, const mb_spline* s_root, int cap, int root) {
  MB_REQUIRE(root >= 0 && root < ctx->comm_size, "broadcast root out of range");
  const size_t nd = spline_wire_doubles(cap);
  double* d_w = ctx->arena.take_n<double>(nd);
  std::vector<double>& h = ctx->comm_host;
  h.resize(nd);
  const bool is_root = ctx->comm_rank == root;
  if (is_root) {
    MB_REQUIRE(s_root != nullptr, "broadcast root has no spline");
    spline_pack(s_root, cap, h.data());
    MB_CUDA(cudaMemcpyAsync(d_w, h.data(), nd * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  ctx->launches++;
  MB_NCCL(nccl().Broadcast(d_w, d_w, nd, ncclDouble, root, comm_of(ctx), ctx->stream));
  if (is_root) return nullptr;
  MB_CUDA(cudaMemcpyAsync(h.data(), d_w, nd * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  MB_CUDA(cudaStreamSynchronize(ctx->stream));
  return spline_unpack(ctx, h.data(), cap);
}

}  // namespace mb

using namespace mb;

extern "C" {

int mb_comm_unique_id(void* id_out) {
  return guarded([&] {
    MB_REQUIRE(id_out != nullptr, "id_out is NULL");
    static_assert(sizeof(ncclUniqueId) == MB_COMM_ID_BYTES, "MB_COMM_ID_BYTES must match ncclUniqueId");
    ncclUniqueId id;
    MB_NCCL(nccl().GetUniqueId(&id));
    std::memcpy(id_out, &id, sizeof id);
  });
}

int mb_comm_init(mb_ctx* ctx, int nranks, int rank, const void* id) {
  return guarded([&] {
    MB_REQUIRE(ctx && id, "NULL argument");
    MB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "rank / nranks");
    MB_REQUIRE(ctx->comm == nullptr, "the context already has a communicator");
    MB_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof uid);
    ncclComm_t c = nullptr;
    MB_NCCL(nccl().CommInitRank(&c, nranks, uid, rank));
    ctx->comm = c;
    ctx->comm_rank = rank;
    ctx->comm_size = nranks;
  });
}

int mb_comm_destroy(mb_ctx* ctx) {
  return guarded([&] {
    MB_REQUIRE(ctx, "ctx is NULL");
    MB_CUDA(cudaSetDevice(ctx->device));
    MB_CUDA(cudaDeviceSynchronize());
    comm_release(ctx);
  });
}

int mb_comm_rank(const mb_ctx* ctx) { return ctx ? ctx->comm_rank : MB_E_ARG; }
int mb_comm_size(const mb_ctx* ctx) { return ctx ? ctx->comm_size : MB_E_ARG; }

const char* mb_comm_backend(void) {
  static thread_local std::string s;
  try { s = "NCCL via " + nccl().source; } catch (const Error& e) { s = e.what(); }
  return s.c_str();
}

int mb_comm_allreduce_f64(mb_ctx* ctx, double* values_host, int n, int op) {
  return guarded([&] {
    MB_REQUIRE(ctx && values_host && n >= 1, "bad argument");
    MB_REQUIRE(op == MB_SUM || op == MB_MAX, "op must be MB_SUM or MB_MAX");
    if (!ctx->comm) return;
    MB_CUDA(cudaSetDevice(ctx->device));
    ctx->arena.begin(ctx->stream);
    double* d = ctx->arena.upload(values_host, (size_t)n, ctx->stream);
    comm_allreduce_f64(ctx, d, n, op, ctx->stream);
    MB_CUDA(cudaMemcpyAsync(values_host, d, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int mb_gram_allreduce(mb_ctx* ctx, const double* R_host, int n_local, int K, double* G_host) {
  return guarded([&] {
    MB_REQUIRE(ctx && G_host, "NULL argument");
    MB_REQUIRE(n_local >= 0 && K >= 1 && K <= 8, "need n_local >= 0 and 1 <= K <= 8");
    MB_REQUIRE(n_local == 0 || R_host, "R is NULL");
    MB_CUDA(cudaSetDevice(ctx->device));
    ctx->arena.begin(ctx->stream);
    ABuf<double> dG(ctx->arena, (size_t)K * K);
    if (n_local > 0) {
      ABuf<double> dR(ctx->arena, (size_t)n_local * K);
      dR.upload(R_host, (size_t)n_local * K, ctx->stream);
      gram(ctx, dR.p, n_local, K, dG.p, ctx->stream);
    } else {
      MB_CUDA(cudaMemsetAsync(dG.p, 0, sizeof(double) * K * K, ctx->stream));   // a rank that owns no residual rows
    }
    comm_allreduce_f64(ctx, dG.p, K * K, MB_SUM, ctx->stream);
    MB_CUDA(cudaMemcpyAsync(G_host, dG.p, sizeof(double) * K * K, cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int mb_spline_bcast(mb_ctx* ctx, mb_spline** s, int max_knots, int root) {
  return guarded([&] {
    MB_REQUIRE(ctx && s, "NULL argument");
    MB_REQUIRE(max_knots >= 1, "max_knots must be positive");
    if (!ctx->comm) return;
    MB_CUDA(cudaSetDevice(ctx->device));
    ctx->arena.begin(ctx->stream);
    mb_spline* got = spline_bcast(ctx, ctx->comm_rank == root ? *s : nullptr, max_knots, root);
    if (ctx->comm_rank != root) *s = got;
  });
}

}  // extern "C"
