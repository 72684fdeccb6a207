// Device-side pieces of ensemble.cu (k_ens_svm, k_ens_smooth, k_ens_points): grid geometry and the float64
// "smooth" models g / n / m.  (Part 5 is fused into the grid-evaluation kernel through AccFuse, common.cuh.)
#pragma once
#include "common.cuh"

namespace mb {

struct EnsGeom {
  double xmin, ymax, rx, ry;
  int nrow, ncol;
};

// gam (V73:604-608), nnet (V73:463-479), earth (V73:539-549) - Appendix B of SURVEY.md
struct SmoothParams {
  const double* gam; double w_g;
  const double* nn; int nn_H; double nn_max2, nn_min, w_n;
  int mars_T; const double* mars_coef; const int* mars_off; const int* mars_var; const int* mars_dir;
  const double* mars_cut; double w_m;
  double w_total;
  int only_gbm;
};

// sum_k w_k f_k(x) over the kept smooth models, x[0..P) in float64
__device__ __forceinline__ double smooth_models(const double* x, int P, const SmoothParams& sp) {
  double s = 0.0;
  if (sp.gam) {
    double v = sp.gam[0];
    for (int f = 0; f < P; ++f) v = fma(sp.gam[1 + f], x[f], v);
    s = fma(sp.w_g, v, s);
  }
  if (sp.nn) {
    const double* wo = sp.nn + (P + 1) * sp.nn_H;
    double v = wo[0];
    for (int h = 0; h < sp.nn_H; ++h) {
      const double* wh = sp.nn + h * (P + 1);
      double z = wh[0];
      for (int f = 0; f < P; ++f) z = fma(wh[1 + f], x[f], z);
      v = fma(wo[1 + h], 1.0 / (1.0 + exp(-z)), v);
    }
    s = fma(sp.w_n, v * sp.nn_max2 + sp.nn_min, s);
  }
  if (sp.mars_T > 0) {
    double v = 0.0;
    for (int t = 0; t < sp.mars_T; ++t) {
      double b = sp.mars_coef[t];
      for (int q = sp.mars_off[t]; q < sp.mars_off[t + 1]; ++q) {
        const double xv = x[sp.mars_var[q]];
        const int dir = sp.mars_dir[q];
        b *= (dir == 2) ? xv : fmax(0.0, dir * (xv - sp.mars_cut[q]));
      }
      v += b;
    }
    s = fma(sp.w_m, v, s);
  }
  return s;
}

}  // namespace mb
