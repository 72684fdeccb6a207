// placeholder until the tiling kernels land (next commit)
#include "common.cuh"
#include "internal.h"
namespace mb {
void tiles_tps(mb_ctx*, const mb_grid&, const double*, const double*, int, int, double, double, int, double, int, double*, cudaStream_t) { throw Error(MB_E_UNSUPPORTED, "tiles: not built yet"); }
void tiles_merge(mb_ctx*, const mb_grid&, int, int, const mb_window*, const double* const*, double*, cudaStream_t) { throw Error(MB_E_UNSUPPORTED, "tiles: not built yet"); }
void gram(mb_ctx*, const double*, int, int, double*, cudaStream_t) { throw Error(MB_E_UNSUPPORTED, "gram: not built yet"); }
void gather_cells(mb_ctx*, const double*, int64_t, const int32_t*, const int32_t*, int, double*) { throw Error(MB_E_UNSUPPORTED, "gather: not built yet"); }
}
