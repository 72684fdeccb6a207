// mbarrier + TMA plumbing shared by the kernels that stage tiles through shared memory: the 3-D tensor copy
// (cp.async.bulk.tensor, UTMALDG in SASS) that brings a covariate tile [C planes][rows][cols] in with one instruction, and the
// host side that encodes its CUtensorMap (driver entry point resolved at run time - the library does not link libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace mb {

__device__ __forceinline__ uint32_t ac_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ac_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ac_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ac_fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void ac_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void ac_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ac_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ac_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(ac_smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// box of the 3-D tensor described by *tmap at element coordinates (c0 = column, c1 = row, c2 = plane) -> dst (128-byte aligned);
// c0 * element size must be a multiple of 16 bytes - any other start column faults with "illegal instruction" (tools/tma_probe.cu);
// out-of-bounds elements are filled as the map says (NaN here: an NA cell, which is what a cell outside the raster is)
__device__ __forceinline__ void ac_tma_load_3d(void* dst, const CUtensorMap* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          ac_smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(ac_smem_u32(bar))
      : "memory");
}

// box of the 2-D tensor described by *tmap at element coordinates (c0 = column, c1 = row) -> dst (128-byte aligned)
__device__ __forceinline__ void ac_tma_load_2d(void* dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          ac_smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(ac_smem_u32(bar))
      : "memory");
}

using TensorMapEncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    (void)cudaGetLastError();
    return reinterpret_cast<TensorMapEncodeFn>(p);
  }();
  return fn;
}

// Host: tensor map of a row-major float64 matrix [rows][row_stride] (the padded ensemble accumulator) with a box of
// box_cols x box_rows.  The box start column must be a multiple of 2 (16 bytes; probed: tools/tma_probe.cu).
inline bool make_f64_matrix_tensor_map(CUtensorMap* out, const double* base, int64_t row_stride, int64_t rows, int box_cols,
                                       int box_rows) {
  TensorMapEncodeFn fn = tensor_map_encoder();
  if (!fn || rows <= 0 || box_rows <= 0 || box_rows > 256 || box_cols > 256 || (row_stride % 2) != 0 ||
      (reinterpret_cast<uintptr_t>(base) & 15u) != 0)
    return false;
  const cuuint64_t dims[2] = {(cuuint64_t)row_stride, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)row_stride * 8};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  // FLOAT64 elements are moved as bits: no conversion, NaN payloads (NA cells) survive
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Host: tensor map of C float32 planes [C][nrow][ncol] (plane stride `plane` elements) with a box of box_cols x box_rows x C.
// Returns false when the layout cannot be described (row stride or base not 16-byte aligned, C = 0, no driver entry point):
// the caller then stages the tile with plain loads.
inline bool make_plane_tensor_map(CUtensorMap* out, const float* base, int ncol, int nrow, int C, int64_t plane, int box_cols,
                                  int box_rows) {
  if (C <= 0 || (ncol % 4) != 0 || (plane % 4) != 0 || (reinterpret_cast<uintptr_t>(base) & 15u) != 0) return false;
  TensorMapEncodeFn fn = tensor_map_encoder();
  if (!fn) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)ncol, (cuuint64_t)nrow, (cuuint64_t)C};
  const cuuint64_t strides[2] = {(cuuint64_t)ncol * 4, (cuuint64_t)plane * 4};
  const cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, (cuuint32_t)C};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA) == CUDA_SUCCESS;
}

}  // namespace mb
