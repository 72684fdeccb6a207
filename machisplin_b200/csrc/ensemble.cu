// placeholder until the ensemble kernels land (next commit)
#include "common.cuh"
#include "internal.h"
struct mb_ensemble { mb_grid g; };
namespace mb {
mb_ensemble* ensemble_create(mb_ctx*, const mb_grid&, const mb_models&, const char*, const double*, double) { throw Error(MB_E_UNSUPPORTED, "ensemble: not built yet"); }
void ensemble_free(mb_ensemble* e) { delete e; }
mb_grid ensemble_grid(const mb_ensemble* e) { return e->g; }
void ensemble_eval(mb_ctx*, const mb_ensemble*, const float*, int, const mb_spline*, const double*, const mb_window*, double*, cudaStream_t) { throw Error(MB_E_UNSUPPORTED, "ensemble: not built yet"); }
void ensemble_predict_points(mb_ctx*, const mb_ensemble*, const double*, int, double*) { throw Error(MB_E_UNSUPPORTED, "ensemble: not built yet"); }
}
