// SM partitions for the one phase of mb_mltps_predict* in which two kinds of work cannot share the GPU the ordinary way.
//
// Stage 1 of the GCV fit's tridiagonalisation is ~1 100 small dependent launches (cluster QR, panel kernels) between full-GPU
// FP64 products; the per-cell ensemble kernels are grids of 10^4 .. 10^5 CTAs that fill every SM.  Side by side on ordinary
// streams the panel kernels queue behind ensemble CTAs (224 instead of 171 ms per step in round 1), so round 1 DEFERRED the
// ensemble until stage 1 was over - 41 ms during which most SMs idle.  CUDA green contexts give the third option: the device's SMs
// are split into a FIT partition and an ENSEMBLE partition; stage 1 runs on streams of the first, the forest kernel on a stream of
// the second from t = 0, neither can take the other's SMs, and everything after stage 1 goes back to the ordinary streams
// (all SMs).  The driver entry points are resolved at run time (the library does not link libcuda); if anything is missing or
// refused the context simply keeps the deferred schedule.
#include "common.cuh"
#include "internal.h"

#include <cuda.h>

#include <cstdlib>

namespace mb {
namespace {

template <class F>
bool entry(const char* name, F& fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    (void)cudaGetLastError();
    return false;
  }
  fn = reinterpret_cast<F>(p);
  return true;
}

}  // namespace

// Creates the two partitions and their streams once per context.  fit_sms = SMs wanted for the fit partition (rounded up by the
// driver to its granularity, 8 on sm_90+).  Returns false (and leaves ctx->gc_ok false) when green contexts are not available.
bool greenctx_setup(mb_ctx* ctx, int fit_sms) {
  if (ctx->gc_tried) return ctx->gc_ok;
  ctx->gc_tried = true;
  // Under a tool that injects into the process (Nsight Compute, compute-sanitizer) the context keeps the deferred schedule: ncu
  // 2025.2 dies without a message (exit code 9) at the first launch on a green-context stream when it profiles EVERY kernel
  // (`--metrics ... -c N` launch lists; `-k regex` captures survive) - measured in r2z, profiles/README.md.
  if (std::getenv("CUDA_INJECTION64_PATH") || std::getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || std::getenv("MB_NO_SM_PARTITIONS"))
    return false;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*GetDevResource)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
  CUresult (*SplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int) = nullptr;
  CUresult (*GenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
  CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
  CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
  if (!entry("cuDeviceGet", DeviceGet) || !entry("cuDeviceGetDevResource", GetDevResource) ||
      !entry("cuDevSmResourceSplitByCount", SplitByCount) || !entry("cuDevResourceGenerateDesc", GenerateDesc) ||
      !entry("cuGreenCtxCreate", GreenCtxCreate) || !entry("cuGreenCtxStreamCreate", GreenCtxStreamCreate))
    return false;
  CUdevice dev;
  if (DeviceGet(&dev, ctx->device) != CUDA_SUCCESS) return false;
  CUdevResource all, grp, rest;
  if (GetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
  const int total = (int)all.sm.smCount;
  if (fit_sms < 8 || fit_sms > total - 16) return false;
  unsigned int ngroups = 1;
  std::memset(&grp, 0, sizeof grp);
  std::memset(&rest, 0, sizeof rest);
  if (SplitByCount(&grp, &ngroups, &all, &rest, 0, (unsigned int)fit_sms) != CUDA_SUCCESS || ngroups < 1) return false;
  if (grp.sm.smCount < 8 || rest.sm.smCount < 16) return false;
  CUdevResourceDesc da = nullptr, db = nullptr;
  CUgreenCtx ga = nullptr, gb = nullptr;
  if (GenerateDesc(&da, &grp, 1) != CUDA_SUCCESS || GreenCtxCreate(&ga, da, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
  if (GenerateDesc(&db, &rest, 1) != CUDA_SUCCESS || GreenCtxCreate(&gb, db, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
  int prio_lo = 0, prio_hi = 0;
  if (cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess) return false;
  CUstream sa = nullptr, sx = nullptr, sb = nullptr;
  if (GreenCtxStreamCreate(&sa, ga, CU_STREAM_NON_BLOCKING, prio_hi) != CUDA_SUCCESS ||
      GreenCtxStreamCreate(&sx, ga, CU_STREAM_NON_BLOCKING, prio_hi) != CUDA_SUCCESS ||
      GreenCtxStreamCreate(&sb, gb, CU_STREAM_NON_BLOCKING, prio_lo) != CUDA_SUCCESS)
    return false;
  for (cudaEvent_t& ev : ctx->gc_ev)
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return false;
  ctx->gc_fit = ga; ctx->gc_ens = gb;
  ctx->gc_fit_stream = (cudaStream_t)sa; ctx->gc_fit_aux = (cudaStream_t)sx; ctx->gc_ens_stream = (cudaStream_t)sb;
  ctx->gc_fit_sms = (int)grp.sm.smCount; ctx->gc_ens_sms = (int)rest.sm.smCount;
  ctx->gc_ok = true;
  return true;
}

void greenctx_release(mb_ctx* ctx) {
  if (!ctx->gc_ok) return;
  CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
  for (cudaStream_t s : {ctx->gc_fit_stream, ctx->gc_fit_aux, ctx->gc_ens_stream})
    if (s) cudaStreamDestroy(s);
  for (cudaEvent_t ev : ctx->gc_ev)
    if (ev) cudaEventDestroy(ev);
  if (entry("cuGreenCtxDestroy", GreenCtxDestroy)) {
    if (ctx->gc_fit) GreenCtxDestroy((CUgreenCtx)ctx->gc_fit);
    if (ctx->gc_ens) GreenCtxDestroy((CUgreenCtx)ctx->gc_ens);
  }
  ctx->gc_ok = false;
}

}  // namespace mb
