/*
 * CPU restatement (plain C + OpenMP) of the O(cells x knots) and O(cells x model) loops of the hot
 * path.  TEST INFRASTRUCTURE / CPU BASELINE ONLY - never linked into the product library.
 * PARITY UNPINNED (see oracle/__init__.py).
 *
 *   orc_tps_eval       predict.Krig -> Fortran multrb as reached from terra::interpolate
 *                      (V73:726, V73:753): one log per (cell, knot), float64, radfun clamp 1e-20.
 *   orc_ensemble_eval  terra::predict(rast_stack, model_k) for the kept models + weighted sum / total
 *                      weight (V73:468-619) + the NA-propagating add of the TPS surface (V73:906-907).
 * threads <= 0 uses every core OpenMP reports; threads = 1 is the faithful stand-in for the reference,
 * which is single-threaded R (V73:116-117).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <omp.h>

#define RBF_CONST 0.039788735772973836 /* 1/(8 pi) */

int orc_max_threads(void) { return omp_get_max_threads(); }

void orc_tps_eval(const double* ksx, const double* ksy, const double* c, int np, const double* d,
                  const double* center, const double* scale, double xmin, double ymax, double rx, double ry,
                  int r0, int r1, int c0, int c1, double* out, int threads) {
  const int wc = c1 - c0;
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int row = r0; row < r1; ++row) {
    const double y = ymax - (row + 0.5) * ry;
    const double sy = (y - center[1]) / scale[1];
    for (int col = c0; col < c1; ++col) {
      const double x = xmin + (col + 0.5) * rx;
      const double sx = (x - center[0]) / scale[0];
      double acc = 0.0;
      for (int k = 0; k < np; ++k) {
        const double dx = sx - ksx[k], dy = sy - ksy[k];
        double d2 = dx * dx + dy * dy;
        if (d2 < 1e-20) d2 = 1e-20;
        acc += c[k] * (0.5 * log(d2) * d2);
      }
      out[(size_t)(row - r0) * wc + (col - c0)] = d[0] + d[1] * sx + d[2] * sy + RBF_CONST * acc;
    }
  }
}

typedef struct {
  int P;
  const double* gam_coef;
  const double* nn_wts; int nn_H; double nn_max2, nn_min;
  int mars_T; const int8_t* mars_dirs; const double* mars_cuts; const double* mars_coef;
  int svm_S; const double* svm_sv; const double* svm_alpha; double svm_b, svm_sigma;
  const double* svm_x_center; const double* svm_x_scale; double svm_y_center, svm_y_scale;
  int rf_ntree, rf_nrnodes;
  const int32_t* rf_left; const int32_t* rf_right; const int8_t* rf_status; const int32_t* rf_bestvar;
  const double* rf_split; const double* rf_nodepred;
  int gbm_ntrees; double gbm_initF; const int32_t* gbm_tree_off;
  const int32_t* gbm_splitvar; const double* gbm_splitcode;
  const int32_t* gbm_left; const int32_t* gbm_right; const int32_t* gbm_missing;
} orc_models;

static double f_gam(const orc_models* m, const double* x) {
  double v = m->gam_coef[0];
  for (int f = 0; f < m->P; ++f) v += m->gam_coef[1 + f] * x[f];
  return v;
}
static double f_nnet(const orc_models* m, const double* x) {
  const int P = m->P, H = m->nn_H;
  const double* wo = m->nn_wts + (size_t)(P + 1) * H;
  double v = wo[0];
  for (int h = 0; h < H; ++h) {
    const double* wh = m->nn_wts + (size_t)h * (P + 1);
    double z = wh[0];
    for (int f = 0; f < P; ++f) z += wh[1 + f] * x[f];
    v += wo[1 + h] / (1.0 + exp(-z));
  }
  return v * m->nn_max2 + m->nn_min;
}
static double f_mars(const orc_models* m, const double* x) {
  double v = 0.0;
  for (int t = 0; t < m->mars_T; ++t) {
    double b = m->mars_coef[t];
    for (int f = 0; f < m->P; ++f) {
      const int dir = m->mars_dirs[(size_t)t * m->P + f];
      if (dir == 2) b *= x[f];
      else if (dir != 0) { const double h = dir * (x[f] - m->mars_cuts[(size_t)t * m->P + f]); b *= h > 0 ? h : 0.0; }
    }
    v += b;
  }
  return v;
}
static double f_svm(const orc_models* m, const double* x) {
  double xs[16];
  for (int f = 0; f < m->P; ++f) xs[f] = (x[f] - m->svm_x_center[f]) / m->svm_x_scale[f];
  double a = 0.0;
  for (int i = 0; i < m->svm_S; ++i) {
    double d2 = 0.0;
    for (int f = 0; f < m->P; ++f) { const double dd = xs[f] - m->svm_sv[(size_t)i * m->P + f]; d2 += dd * dd; }
    a += m->svm_alpha[i] * exp(-m->svm_sigma * d2);
  }
  return (a - m->svm_b) * m->svm_y_scale + m->svm_y_center;
}
static double f_rf(const orc_models* m, const double* x) {
  double a = 0.0;
  for (int t = 0; t < m->rf_ntree; ++t) {
    const size_t o = (size_t)t * m->rf_nrnodes;
    int k = 0;
    while (m->rf_status[o + k] != -1)
      k = (x[m->rf_bestvar[o + k] - 1] <= m->rf_split[o + k] ? m->rf_left[o + k] : m->rf_right[o + k]) - 1;
    a += m->rf_nodepred[o + k];
  }
  return a / m->rf_ntree;
}
static double f_gbm(const orc_models* m, const double* x) {
  double a = m->gbm_initF;
  for (int t = 0; t < m->gbm_ntrees; ++t) {
    const int o = m->gbm_tree_off[t];
    int k = 0;
    while (m->gbm_splitvar[o + k] != -1) {
      const double xv = x[m->gbm_splitvar[o + k]];
      k = (xv != xv) ? m->gbm_missing[o + k] : (xv < m->gbm_splitcode[o + k] ? m->gbm_left[o + k] : m->gbm_right[o + k]);
    }
    a += m->gbm_splitcode[o + k];
  }
  return a;
}

/* cov: C planes of the full grid (float32); kept: letters; w: weights; tps: window-sized surface or NULL */
void orc_ensemble_eval(const orc_models* m, const char* kept, const double* w, double w_total, const float* cov,
                       int C, int nrow, int ncol, double xmin, double ymax, double rx, double ry, int r0, int r1,
                       int c0, int c1, const double* tps, double* out, int threads) {
  const int wc = c1 - c0;
  const size_t plane = (size_t)nrow * ncol;
  if (threads <= 0) threads = omp_get_max_threads();
  int only_gbm = (kept[0] == 'b' && kept[1] == 0);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int row = r0; row < r1; ++row) {
    double x[16];
    for (int col = c0; col < c1; ++col) {
      int anynan = 0;
      for (int f = 0; f < C; ++f) {
        const float v = cov[f * plane + (size_t)row * ncol + col];
        x[f] = (double)v;
        anynan |= (v != v);
      }
      x[C] = xmin + (col + 0.5) * rx;
      x[C + 1] = ymax - (row + 0.5) * ry;
      double s = 0.0;
      for (int q = 0; kept[q]; ++q) {
        double v = 0.0;
        if (anynan && kept[q] != 'b') continue;
        switch (kept[q]) {
          case 'b': v = f_gbm(m, x); break;
          case 'g': v = f_gam(m, x); break;
          case 'n': v = f_nnet(m, x); break;
          case 'm': v = f_mars(m, x); break;
          case 'r': v = f_rf(m, x); break;
          case 'v': v = f_svm(m, x); break;
        }
        s += v * w[q];
      }
      double r = s / w_total;
      if (anynan && !only_gbm) r = NAN;
      const size_t o = (size_t)(row - r0) * wc + (col - c0);
      if (tps) r += tps[o];
      out[o] = r;
    }
  }
}
