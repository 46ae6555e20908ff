/*
 * phovo_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT; see phovo_oracle.h).
 *
 * Double-precision, single-threaded C restatement of the reference hot path.  Citations:
 *   AN = phovo/include/CPhotoconsistencyOdometryAnalytic.h
 *   CE = phovo/include/CPhotoconsistencyOdometryCeres.h
 *   SA = third_party/sample.h   JE = third_party/jet_extras.h
 *   BASE = phovo/include/CPhotoconsistencyOdometry.h
 * Third-party arithmetic (OpenCV imgproc, Eigen 6x6 inverse, Ceres LM) is restated from the
 * libraries' documented algorithms and pinned against python cv2 in tests/.
 *
 * Build: gcc -O3 -mtune=native -std=c11 -fPIC -shared -pthread (the reference's own flags are
 * "-O3 -mtune=native", CMakeLists.txt:58-60; no -march, hence no FMA contraction on x86-64).
 */
#define _POSIX_C_SOURCE 200809L
#include "phovo_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define MAXL PHOVO_MAX_LEVELS

/* ======================================================================================== */
/* image arithmetic                                                                          */
/* ======================================================================================== */

/* AN:471 / AN:484: intensityImage.convertTo(aux, CV_64F, 1./255) -> saturate_cast<double>(v*alpha) */
void pho_convert_u8(const uint8_t* src, size_t step, int rows, int cols, double* dst) {
  const double alpha = 1. / 255;
  for (int r = 0; r < rows; ++r) {
    const uint8_t* s = src + (size_t)r * step;
    for (int c = 0; c < cols; ++c) dst[(size_t)r * cols + c] = (double)s[c] * alpha;
  }
}

/* cvRound: round-half-to-even (lrint under the default rounding mode) */
static int cv_round(double v) { return (int)lrint(v); }

/* AN:132: cv::resize(img, aux, Size(0,0), factor, factor) -> dsize = cvRound(ssize * factor) */
void pho_level_size(int rows, int cols, int level, int* out_rows, int* out_cols) {
  double factor = 1.;
  for (int l = 0; l < level; ++l) factor = factor / 2; /* AN:159 */
  *out_rows = level == 0 ? rows : cv_round(rows * factor);
  *out_cols = level == 0 ? cols : cv_round(cols * factor);
}

/* One axis of OpenCV's INTER_LINEAR coordinate table (imgproc/resize.cpp, bilinear branch):
 *   f = (float)((d + 0.5) * scale - 0.5); s = floor(f); f -= s; clamp at both borders. */
static void linear_axis(int d, double scale, int ssize, int* s0, float* w1, int is_x) {
  float f = (float)((d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (s < 0) { f = 0.f; s = 0; }
  if (is_x) {
    if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
  }
  *s0 = s;
  *w1 = f;
}

/* cv::resize by EXACTLY one half: imgproc/resize.cpp switches INTER_LINEAR to INTER_AREA when both
 * integer scale factors are 2 ("if( interpolation == INTER_LINEAR && is_area_fast && iscale_x == 2 &&
 * iscale_y == 2 ) interpolation = INTER_AREA") and runs resizeAreaFast_Invoker<double, double>:
 *   whole 2x2 cells : sum = ((S[0,0] + S[0,1]) + S[1,0]) + S[1,1], D = sum * 0.25f
 *   cells cut by the right / bottom edge (output size is cvRound(size / 2), so sizes = 3 mod 4 have
 *   them): the taps inside the image are summed in raster order and D = (float)sum / count -- in
 *   SINGLE precision.
 * Pinned bit for bit against cv2 with cv2.setUseOptimized(False) (plain C++ path; the IPP / SIMD
 * dispatch of a particular build rounds differently), tests/test_oracle_pins.py. */
static void resize_half_area(const double* src, int rows, int cols, int orows, int ocols, double* dst) {
  const int dwidth1 = cols / 2;
  for (int dy = 0; dy < orows; ++dy) {
    double* D = dst + (size_t)dy * ocols;
    const int sy0 = dy * 2;
    const int w = sy0 + 2 <= rows ? dwidth1 : 0;
    if (sy0 >= rows) { for (int dx = 0; dx < ocols; ++dx) D[dx] = 0; continue; }
    int dx = 0;
    for (; dx < w; ++dx) {
      const double* S = src + (size_t)sy0 * cols + 2 * dx;
      double sum = 0;
      sum += S[0] + S[1] + S[cols] + S[cols + 1];
      D[dx] = sum * 0.25f;
    }
    for (; dx < ocols; ++dx) {
      double sum = 0; int count = 0; const int sx0 = 2 * dx;
      for (int sy = 0; sy < 2; ++sy) {
        if (sy0 + sy >= rows) break;
        const double* S = src + (size_t)(sy0 + sy) * cols + sx0;
        for (int sx = 0; sx < 2; ++sx) {
          if (sx0 + sx >= cols) break;
          sum += S[sx]; count++;
        }
      }
      D[dx] = (double)((float)sum / count);
    }
  }
}

/* AN:132, default interpolation INTER_LINEAR, always from the ORIGINAL image (not iterative).
 * For the exact factor 2^-level, level >= 2, the taps are the central 2x2 of each 2^level cell,
 * weights .5; level 1 takes OpenCV's area path (above). */
void pho_resize_level(const double* src, int rows, int cols, int level, double* dst) {
  int orows, ocols;
  pho_level_size(rows, cols, level, &orows, &ocols);
  if (level == 0) {
    memcpy(dst, src, sizeof(double) * (size_t)rows * cols);
    return;
  }
  if (level == 1) { resize_half_area(src, rows, cols, orows, ocols, dst); return; }
  double scale = 1.;
  for (int l = 0; l < level; ++l) scale *= 2.; /* 1 / inv_scale, inv_scale = factor exactly */
  int* xofs = (int*)malloc(sizeof(int) * ocols);
  float* xw = (float*)malloc(sizeof(float) * ocols);
  for (int x = 0; x < ocols; ++x) linear_axis(x, scale, cols, &xofs[x], &xw[x], 1);
  double* row0 = (double*)malloc(sizeof(double) * ocols);
  double* row1 = (double*)malloc(sizeof(double) * ocols);
  for (int y = 0; y < orows; ++y) {
    int sy; float fy;
    linear_axis(y, scale, rows, &sy, &fy, 0);
    int y0 = sy < 0 ? 0 : (sy > rows - 1 ? rows - 1 : sy);
    int y1 = sy + 1 < 0 ? 0 : (sy + 1 > rows - 1 ? rows - 1 : sy + 1);
    const double* S0 = src + (size_t)y0 * cols;
    const double* S1 = src + (size_t)y1 * cols;
    /* horizontal pass first (HResizeLinear), then vertical (VResizeLinear) */
    for (int x = 0; x < ocols; ++x) {
      int sx = xofs[x];
      if (sx + 1 < cols) {
        float a0 = 1.f - xw[x], a1 = xw[x];
        row0[x] = S0[sx] * a0 + S0[sx + 1] * a1;
        row1[x] = S1[sx] * a0 + S1[sx + 1] * a1;
      } else { /* dx >= xmax: single tap with weight ONE */
        row0[x] = S0[sx];
        row1[x] = S1[sx];
      }
    }
    float b0 = 1.f - fy, b1 = fy;
    double* D = dst + (size_t)y * ocols;
    for (int x = 0; x < ocols; ++x) D[x] = row0[x] * b0 + row1[x] * b1;
  }
  free(xofs); free(xw); free(row0); free(row1);
}

/* BORDER_REFLECT_101 (cv::BORDER_DEFAULT): gfedcb|abcdefgh|gfedcba */
static int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * (len - 1) - p;
  }
  return p;
}

/* cv::sepFilter2D for CV_64F: generic RowFilter (taps accumulated left to right) followed by a
 * SymmColumnFilter (centre tap first, then symmetric / antisymmetric pairs).
 * col_sym: +1 symmetric, -1 antisymmetric. */
static void sep_filter(const double* src, int rows, int cols, const double* kx, int kxn,
                       const double* ky, int kyn, int col_sym, double* dst) {
  double* buf = (double*)malloc(sizeof(double) * (size_t)rows * cols);
  int ax = kxn / 2, ay = kyn / 2;
  for (int r = 0; r < rows; ++r) {
    const double* S = src + (size_t)r * cols;
    double* B = buf + (size_t)r * cols;
    for (int c = 0; c < cols; ++c) {
      double s = kx[0] * S[reflect101(c - ax, cols)];
      for (int k = 1; k < kxn; ++k) s += kx[k] * S[reflect101(c - ax + k, cols)];
      B[c] = s;
    }
  }
  for (int r = 0; r < rows; ++r) {
    double* D = dst + (size_t)r * cols;
    for (int c = 0; c < cols; ++c) {
      double s;
      if (col_sym > 0) {
        s = ky[ay] * buf[(size_t)r * cols + c] + 0.0; /* + delta */
        for (int k = 1; k <= ay; ++k) {
          double a = buf[(size_t)reflect101(r + k, rows) * cols + c];
          double b = buf[(size_t)reflect101(r - k, rows) * cols + c];
          s += ky[ay + k] * (a + b);
        }
      } else {
        s = 0.0; /* delta */
        for (int k = 1; k <= ay; ++k) {
          double a = buf[(size_t)reflect101(r + k, rows) * cols + c];
          double b = buf[(size_t)reflect101(r - k, rows) * cols + c];
          s += ky[ay + k] * (a - b);
        }
      }
      D[c] = s;
    }
  }
  free(buf);
}

/* AN:181-187: cv::Scharr(src, dst, CV_64F, dx, dy, scale, 0, BORDER_DEFAULT).
 * getScharrKernels gives the derivative kernel [-1 0 1] and the smoothing kernel [3 10 3];
 * cv::Scharr multiplies the SMOOTHING kernel by `scale` (kx if dx==0, else ky). */
void pho_scharr(const double* src, int rows, int cols, int dx, int dy, double scale, double* dst) {
  double kd[3] = {-1., 0., 1.};
  double ks[3] = {3., 10., 3.};
  if (scale != 1.) { ks[0] *= scale; ks[1] *= scale; ks[2] *= scale; }
  if (dx == 1 && dy == 0) sep_filter(src, rows, cols, kd, 3, ks, 3, +1, dst);
  else                    sep_filter(src, rows, cols, ks, 3, kd, 3, -1, dst);
}

/* AN:146-147: cv::GaussianBlur(img, img, Size(k,k), 3).  getGaussianKernel(k, sigma, CV_64F). */
void pho_gaussian_blur(double* img, int rows, int cols, int ksize, double sigma) {
  if (ksize <= 0) return;
  if (ksize == 1) return; /* 1x1 kernel is the identity */
  double* k = (double*)malloc(sizeof(double) * ksize);
  double sigmaX = sigma > 0 ? sigma : ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
  double scale2X = -0.5 / (sigmaX * sigmaX);
  double sum = 0;
  for (int i = 0; i < ksize; ++i) {
    double x = i - (ksize - 1) * 0.5;
    double t = exp(scale2X * x * x);
    k[i] = t;
    sum += t;
  }
  sum = 1. / sum;
  for (int i = 0; i < ksize; ++i) k[i] *= sum;
  double* out = (double*)malloc(sizeof(double) * (size_t)rows * cols);
  sep_filter(img, rows, cols, k, ksize, k, ksize, +1, out);
  memcpy(img, out, sizeof(double) * (size_t)rows * cols);
  free(out); free(k);
}

/* ======================================================================================== */
/* solver object                                                                              */
/* ======================================================================================== */

struct pho_oracle {
  phovo_config cfg;
  int storage_f32, lean;
  double K[9];
  int have_src, have_tgt;
  int rows[MAXL], cols[MAXL];
  double* I0[MAXL]; double* D0[MAXL]; double* I1[MAXL]; double* Gx[MAXL]; double* Gy[MAXL];
  double state[6];
  phovo_iter_stats* log; int nlog, caplog;
  int iters_per_level[MAXL];
  double optimize_seconds;
};

static double now_sec(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

pho_oracle* pho_create(void) {
  pho_oracle* o = (pho_oracle*)calloc(1, sizeof(pho_oracle));
  /* AN:430-443 constructor defaults */
  o->cfg.mode = PHOVO_MODE_ANALYTIC_REF;
  o->cfg.num_levels = 5;
  for (int l = 0; l < MAXL; ++l) {
    o->cfg.grad_scale[l] = 0.0625; o->cfg.lambda_step[l] = 1.; o->cfg.min_gradient_norm[l] = 300.;
  }
  o->cfg.max_num_iterations[2] = 5; o->cfg.max_num_iterations[3] = 20; o->cfg.max_num_iterations[4] = 50;
  o->cfg.min_depth = 0.3; o->cfg.max_depth = 5.0;
  return o;
}

static void free_levels(double** p) {
  for (int l = 0; l < MAXL; ++l) { free(p[l]); p[l] = NULL; }
}

void pho_destroy(pho_oracle* o) {
  if (!o) return;
  free_levels(o->I0); free_levels(o->D0); free_levels(o->I1); free_levels(o->Gx); free_levels(o->Gy);
  free(o->log);
  free(o);
}

void pho_set_config(pho_oracle* o, const phovo_config* cfg) { o->cfg = *cfg; }
void pho_set_options(pho_oracle* o, int storage_f32, int lean) { o->storage_f32 = storage_f32; o->lean = lean; }
void pho_set_intrinsics(pho_oracle* o, const double K[9]) { memcpy(o->K, K, sizeof(double) * 9); }
void pho_set_initial_state(pho_oracle* o, const double s[6]) { memcpy(o->state, s, sizeof(double) * 6); }
void pho_get_state(const pho_oracle* o, double s[6]) { memcpy(s, o->state, sizeof(double) * 6); }

static void round_to_f32(double* p, size_t n) {
  for (size_t i = 0; i < n; ++i) p[i] = (double)(float)p[i];
}

/* AN:115-163 BuildPyramid: every level is resized from the ORIGINAL image; optional blur twice */
static void build_pyramid(pho_oracle* o, const double* img, int rows, int cols, double** pyr, int apply_blur) {
  for (int l = 0; l < o->cfg.num_levels; ++l) {
    int r, c;
    pho_level_size(rows, cols, l, &r, &c);
    o->rows[l] = r; o->cols[l] = c;
    free(pyr[l]);
    pyr[l] = (double*)malloc(sizeof(double) * (size_t)r * c);
    pho_resize_level(img, rows, cols, l, pyr[l]);
    if (apply_blur) {
      int k = o->cfg.blur_filter_size[l];
      if (k > 0) { /* AN:144-148, ENABLE_GAUSSIAN_BLUR 1 */
        pho_gaussian_blur(pyr[l], r, c, k, 3);
        pho_gaussian_blur(pyr[l], r, c, k, 3);
      }
    }
  }
}

/* AN:466-476 */
void pho_set_source(pho_oracle* o, const uint8_t* gray, size_t gray_step, const double* depth,
                    size_t depth_step, int rows, int cols) {
  size_t n = (size_t)rows * cols;
  double* aux = (double*)malloc(sizeof(double) * n);
  pho_convert_u8(gray, gray_step, rows, cols, aux);
  build_pyramid(o, aux, rows, cols, o->I0, 1);
  for (int r = 0; r < rows; ++r)
    memcpy(aux + (size_t)r * cols, (const char*)depth + (size_t)r * depth_step, sizeof(double) * cols);
  build_pyramid(o, aux, rows, cols, o->D0, 0);
  free(aux);
  if (o->storage_f32)
    for (int l = 0; l < o->cfg.num_levels; ++l) {
      round_to_f32(o->I0[l], (size_t)o->rows[l] * o->cols[l]);
      if (o->storage_f32 == 1) round_to_f32(o->D0[l], (size_t)o->rows[l] * o->cols[l]);
    }
  o->have_src = 1;
}

/* AN:479-491 + BuildDerivativesPyramids AN:165-189 */
void pho_set_target(pho_oracle* o, const uint8_t* gray, size_t gray_step, int rows, int cols) {
  size_t n = (size_t)rows * cols;
  double* aux = (double*)malloc(sizeof(double) * n);
  pho_convert_u8(gray, gray_step, rows, cols, aux);
  build_pyramid(o, aux, rows, cols, o->I1, 1);
  free(aux);
  for (int l = 0; l < o->cfg.num_levels; ++l) {
    size_t m = (size_t)o->rows[l] * o->cols[l];
    free(o->Gx[l]); free(o->Gy[l]);
    o->Gx[l] = (double*)malloc(sizeof(double) * m);
    o->Gy[l] = (double*)malloc(sizeof(double) * m);
    /* gradients are taken from the double-precision level image (before any f32 rounding) */
    pho_scharr(o->I1[l], o->rows[l], o->cols[l], 1, 0, o->cfg.grad_scale[l], o->Gx[l]);
    pho_scharr(o->I1[l], o->rows[l], o->cols[l], 0, 1, o->cfg.grad_scale[l], o->Gy[l]);
    if (o->storage_f32) { round_to_f32(o->I1[l], m); round_to_f32(o->Gx[l], m); round_to_f32(o->Gy[l], m); }
  }
  o->have_tgt = 1;
}

const double* pho_level_image(const pho_oracle* o, int which, int level, int* rows, int* cols) {
  if (level < 0 || level >= o->cfg.num_levels) return NULL;
  *rows = o->rows[level]; *cols = o->cols[level];
  switch (which) {
    case 0: return o->I0[level];
    case 1: return o->D0[level];
    case 2: return o->I1[level];
    case 3: return o->Gx[level];
    case 4: return o->Gy[level];
  }
  return NULL;
}

/* BASE:47-71 eigenPose */
void pho_state_to_rt(const double s[6], double P[16]) {
  double x = s[0], y = s[1], z = s[2], yaw = s[3], pitch = s[4], roll = s[5];
  P[0] = cos(yaw) * cos(pitch);
  P[1] = cos(yaw) * sin(pitch) * sin(roll) - sin(yaw) * cos(roll);
  P[2] = cos(yaw) * sin(pitch) * cos(roll) + sin(yaw) * sin(roll);
  P[3] = x;
  P[4] = sin(yaw) * cos(pitch);
  P[5] = sin(yaw) * sin(pitch) * sin(roll) + cos(yaw) * cos(roll);
  P[6] = sin(yaw) * sin(pitch) * cos(roll) - cos(yaw) * sin(roll);
  P[7] = y;
  P[8] = -sin(pitch);
  P[9] = cos(pitch) * sin(roll);
  P[10] = cos(pitch) * cos(roll);
  P[11] = z;
  P[12] = 0; P[13] = 0; P[14] = 0; P[15] = 1;
}
void pho_get_rt(const pho_oracle* o, double rt[16]) { pho_state_to_rt(o->state, rt); }

/* ---------------------------------------------------------------------------------------- */
/* AN:191-367 ComputeResidualsAndJacobians.                                                  */
/* residuals: N (target-indexed, last writer wins); jac: N x 6 column-major as in the        */
/* reference (Matrix.h:114-135), source-indexed.  Both must be zeroed by the caller          */
/* (AN:519-524).  winners (optional): per target slot the last source index, else -1.        */
/* Returns the number of source pixels that were valid and in bounds.                        */
/* ---------------------------------------------------------------------------------------- */
static int analytic_residuals_jacobians(const pho_oracle* o, int level, const double st[6],
                                        double* residuals, double* jac, int32_t* winners) {
  const int nRows = o->rows[level], nCols = o->cols[level];
  const size_t N = (size_t)nRows * nCols;
  const double* I0 = o->I0[level]; const double* D0 = o->D0[level];
  const double* I1 = o->I1[level]; const double* Gx = o->Gx[level]; const double* Gy = o->Gy[level];
  const int fixed = o->cfg.mode == PHOVO_MODE_ANALYTIC_FIXED;
  const double minD = o->cfg.min_depth, maxD = o->cfg.max_depth;

  /* AN:203-209 */
  double scaleFactor = 1.0 / pow(2, level);
  double fx = o->K[0] * scaleFactor, fy = o->K[4] * scaleFactor;
  double ox = o->K[2] * scaleFactor, oy = o->K[5] * scaleFactor;
  double inv_fx = 1.f / fx, inv_fy = 1.f / fy;

  double x = st[0], y = st[1], z = st[2], yaw = st[3], pitch = st[4], roll = st[5];
  /* AN:219-241 */
  double sy = sin(yaw), cy = cos(yaw), sp = sin(pitch), cp = cos(pitch), sr = sin(roll), cr = cos(roll);
  double R00 = cy * cp, R01 = cy * sp * sr - sy * cr, R02 = cy * sp * cr + sy * sr;
  double R10 = sy * cp, R11 = sy * sp * sr + cy * cr, R12 = sy * sp * cr - cy * sr;
  double R20 = -sp, R21 = cp * sr, R22 = cp * cr;
  /* AN:243-266 */
  double t1 = cp * sr, t2 = cp * cr, t3 = sp;
  double t4 = (sr * sy + sp * cr * cy);
  double t5 = (sp * sr * cy - cr * sy);
  double t6 = (sp * sr * sy + cr * cy);
  double t7 = (-sp * sr * sy - cr * cy);
  double t8 = (sr * cy - sp * cr * sy);
  double t9 = (sp * cr * sy - sr * cy);
  double t10 = cp * sr * cy;
  double t11 = cp * cy + x; /* AN:253 -- the reference's slip; the Maxima derivation has no "+x" inside */
  double t12 = cp * cr * cy;
  double t13 = sp * cy;
  double t14 = cp * sy;
  double t15 = cp * cy;
  double t16 = sp * sr;
  double t17 = sp * cr;
  double t18 = cp * sr * sy;
  double t19 = cp * cr * sy;
  double t20 = sp * sy;
  double t21 = (cr * sy - sp * sr * cy);
  double t22 = cp * cr;
  double t23 = cp * sr;
  double t24 = cp;

  if (winners) for (size_t i = 0; i < N; ++i) winners[i] = -1;
  int count = 0;
  for (int r = 0; r < nRows; ++r) {
    for (int c = 0; c < nCols; ++c) {
      const size_t i = (size_t)nCols * r + c;
      double pz = D0[i];
      if (!(minD < pz && pz < maxD)) continue; /* AN:280 strict */
      double px = (c - ox) * pz * inv_fx;      /* AN:282 */
      double py = (r - oy) * pz * inv_fy;      /* AN:283 */
      /* AN:291 Rt*point3D (4x4 . 4x1, w = 1) */
      double X = R00 * px + R01 * py + R02 * pz + x * 1.0;
      double Y = R10 * px + R11 * py + R12 * pz + y * 1.0;
      double Z = R20 * px + R21 * py + R22 * pz + z * 1.0;
      double invZ = 1.0 / Z;                   /* AN:294 */
      double tc = (X * fx) * invZ + ox;        /* AN:295 */
      double tr = (Y * fy) * invZ + oy;        /* AN:296 */
      double rr = round(tr), rc = round(tc);   /* AN:297-298: C round(), half away from zero */
      /* AN:302-303; a non-finite coordinate is UB in the reference (cast of NaN/inf to int);
       * the oracle defines it as out of bounds. */
      if (!(rr >= 0. && rr < (double)nRows && rc >= 0. && rc < (double)nCols)) continue;
      int ti = (int)rr, tj = (int)rc;
      double pixel1 = I0[i];
      double pixel2 = I1[(size_t)nCols * ti + tj];

      /* AN:312-342 */
      double t25 = 1.0 / (z + py * t1 + pz * t2 - px * t3);
      double t26 = t25 * t25;
      double A = fixed ? (pz * t4 + py * t5 + px * t15 + x) : (pz * t4 + py * t5 + px * t11);
      double B = (py * t6 + pz * t9 + px * t14 + y);
      double J00 = fx * t25, J10 = 0.0;
      double J01 = 0.0, J11 = fy * t25;
      double J02 = -fx * A * t26;
      double J12 = -fy * B * t26;
      double J03 = fx * (py * t7 + pz * t8 - px * t14) * t25;
      double J13 = fy * (pz * t4 + py * t5 + px * t15) * t25;
      double J04 = fx * (py * t10 + pz * t12 - px * t13) * t25 - fx * (-py * t16 - pz * t17 - px * t24) * A * t26;
      double J14 = fy * (py * t18 + pz * t19 - px * t20) * t25 - fy * (-py * t16 - pz * t17 - px * t24) * B * t26;
      double J05 = fx * (py * t4 + pz * t21) * t25 - fx * (py * t22 - pz * t23) * A * t26;
      double J15 = fy * (pz * t7 + py * t9) * t25 - fy * (py * t22 - pz * t23) * B * t26;

      /* AN:345-356: gradients of I1 are read at the SOURCE index i */
      double gx = Gx[i], gy = Gy[i];
      jac[i + 0 * N] = gx * J00 + gy * J10;
      jac[i + 1 * N] = gx * J01 + gy * J11;
      jac[i + 2 * N] = gx * J02 + gy * J12;
      jac[i + 3 * N] = gx * J03 + gy * J13;
      jac[i + 4 * N] = gx * J04 + gy * J14;
      jac[i + 5 * N] = gx * J05 + gy * J15;
      /* AN:358: residual goes to the TARGET index */
      residuals[(size_t)nCols * ti + tj] = pixel2 - pixel1;
      if (winners) winners[(size_t)nCols * ti + tj] = (int32_t)i;
      ++count;
    }
  }
  return count;
}

/* ---------------------------------------------------------------------------------------- */
/* CE:156-269 residual functor evaluated on T = double and on T = Jet<double,6>.             */
/* The Jet derivative of (tc, tr) is the exact derivative of the projection (what autodiff   */
/* computes), i.e. the Maxima expressions without the AN:253 slip; SampleWithDerivative      */
/* (SA:104-123) chains the bilinearly sampled Gx/Gy into it (JE:87-109).                     */
/* residuals N, jac N x 6 ROW-major (Ceres layout), both target-indexed, zero-initialised    */
/* here (CE:206-212).                                                                         */
/* jet_arith: Ceres' AutoDiffCostFunction runs the functor on T = double when only the cost   */
/* is wanted and on T = Jet<double,6> when the Jacobian is; the scalar part of a Jet quotient */
/* is f.a * (1 / g.a) (ceres/jet.h operator/), not f.a / g.a, so CE:241-242 round differently */
/* in the two instantiations.  It only matters where a projected coordinate sits on an        */
/* integer (CE:250-251 truncates) -- which is EVERY pixel at the identity state the apps      */
/* start from.  Pinned against the reference functor compiled on the Jet stand-in             */
/* (oracle/shim/ceres/jet.h) in tests/test_reference_ceres_functor.py.                        */
/* ---------------------------------------------------------------------------------------- */
static void linear_init_axis(double x, int size, int* x1, int* x2, double* dx) { /* SA:36-50 */
  const int ix = (int)x;
  if (ix < 0) { *x1 = 0; *x2 = 0; *dx = 1.0; }
  else if (ix > size - 2) { *x1 = size - 1; *x2 = size - 1; *dx = 1.0; }
  else { *x1 = ix; *x2 = ix + 1; *dx = *x2 - x; }
}

static int ceres_residuals_jacobians(const pho_oracle* o, int level, const double st[6],
                                     double* residuals, double* jac, int32_t* winners, int jet_arith) {
  const int nRows = o->rows[level], nCols = o->cols[level];
  const size_t N = (size_t)nRows * nCols;
  const double* I0 = o->I0[level]; const double* D0 = o->D0[level];
  const double* I1 = o->I1[level]; const double* Gx = o->Gx[level]; const double* Gy = o->Gy[level];
  const double minD = o->cfg.min_depth, maxD = o->cfg.max_depth;
  /* CE:163-168 */
  double fx = o->K[0] / pow(2, (double)level), fy = o->K[4] / pow(2, (double)level);
  double inv_fx = 1. / fx, inv_fy = 1. / fy;
  double ox = o->K[2] / pow(2, (double)level), oy = o->K[5] / pow(2, (double)level);
  double x = st[0], y = st[1], z = st[2], yaw = st[3], pitch = st[4], roll = st[5];
  double sy = sin(yaw), cy = cos(yaw), sp = sin(pitch), cp = cos(pitch), sr = sin(roll), cr = cos(roll);
  double R00 = cy * cp, R01 = cy * sp * sr - sy * cr, R02 = cy * sp * cr + sy * sr;
  double R10 = sy * cp, R11 = sy * sp * sr + cy * cr, R12 = sy * sp * cr - cy * sr;
  double R20 = -sp, R21 = cp * sr, R22 = cp * cr;

  for (size_t i = 0; i < N; ++i) residuals[i] = 0.;
  if (jac) memset(jac, 0, sizeof(double) * N * 6);
  if (winners) for (size_t i = 0; i < N; ++i) winners[i] = -1;
  int count = 0;
  for (int r = 0; r < nRows; ++r) {
    for (int c = 0; c < nCols; ++c) {
      const size_t i = (size_t)nCols * r + c;
      double d = D0[i];
      if (!(minD < d && d < maxD)) continue;           /* CE:226 */
      double pz = d;
      double px = ((double)c - ox) * pz * inv_fx;      /* CE:230 */
      double py = ((double)r - oy) * pz * inv_fy;      /* CE:231 */
      double X = R00 * px + R01 * py + R02 * pz + x * 1.0; /* CE:235-238 */
      double Y = R10 * px + R11 * py + R12 * pz + y * 1.0;
      double Z = R20 * px + R21 * py + R22 * pz + z * 1.0;
      double tc, tr;
      if (jet_arith) {                                 /* CE:241-242 on Jets: quotient = f.a * (1 / g.a) */
        const double z_inverse = 1.0 / Z;
        tc = ((X * fx) * z_inverse) + ox;
        tr = ((Y * fy) * z_inverse) + oy;
      } else {
        tc = ((X * fx) / Z) + ox;                      /* CE:241 */
        tr = ((Y * fy) / Z) + oy;                      /* CE:242 */
      }
      if (!(tr >= 0. && tr < (double)nRows && tc >= 0. && tc < (double)nCols)) continue; /* CE:246-247 */
      int tri = (int)tr, tci = (int)tc;                /* CE:250-251 truncation */
      /* SA:53-99 SampleLinear at (x = tc, y = tr) with the -0.5 shift */
      double sxp = tc - 0.5, syp = tr - 0.5;
      int x1, x2, y1, y2; double dx, dy;
      linear_init_axis(syp, nRows, &y1, &y2, &dy);
      linear_init_axis(sxp, nCols, &x1, &x2, &dx);
#define BIL(IMG) (dy * (dx * IMG[(size_t)y1 * nCols + x1] + (1.0 - dx) * IMG[(size_t)y1 * nCols + x2]) + \
                  (1 - dy) * (dx * IMG[(size_t)y2 * nCols + x1] + (1.0 - dx) * IMG[(size_t)y2 * nCols + x2]))
      double s0 = BIL(I1), s1 = BIL(Gx), s2 = BIL(Gy);
#undef BIL
      size_t t = (size_t)nCols * tri + tci;
      residuals[t] = 1. * (s0 - I0[i]);                /* CE:253-254 */
      if (winners) winners[t] = (int32_t)i;
      if (jac) {
        /* exact d(tc,tr)/d(state): quotient rule on X*fx/Z */
        double iz = 1. / Z, iz2 = iz * iz;
        double q0 = X - x, q1 = Y - y, q2 = Z - z;
        double Zp = -(sp * sr * py + sp * cr * pz + cp * px);
        double Zr = R22 * py - R21 * pz;
        double Ju[6], Jv[6];
        Ju[0] = fx * iz;            Jv[0] = 0.;
        Ju[1] = 0.;                 Jv[1] = fy * iz;
        Ju[2] = -fx * X * iz2;      Jv[2] = -fy * Y * iz2;
        Ju[3] = -fx * q1 * iz;      Jv[3] = fy * q0 * iz;
        Ju[4] = fx * (cy * q2 * iz - Zp * X * iz2);
        Jv[4] = fy * (sy * q2 * iz - Zp * Y * iz2);
        Ju[5] = fx * ((R02 * py - R01 * pz) * iz - Zr * X * iz2);
        Jv[5] = fy * ((R12 * py - R11 * pz) * iz - Zr * Y * iz2);
        for (int k = 0; k < 6; ++k) jac[t * 6 + k] = s1 * Ju[k] + s2 * Jv[k]; /* JE:87-109 */
      }
      ++count;
    }
  }
  return count;
}

/* J^T J (upper 21) and J^T r, plain sequential double sums (Eigen's blocking differs at 1e-13) */
static void normal_equations_colmajor(const double* jac, const double* res, size_t N, double H[21], double g[6], double* cost) {
  int k = 0;
  for (int a = 0; a < 6; ++a)
    for (int b = a; b < 6; ++b) {
      double s = 0;
      const double* ja = jac + a * N; const double* jb = jac + b * N;
      for (size_t i = 0; i < N; ++i) s += ja[i] * jb[i];
      H[k++] = s;
    }
  for (int a = 0; a < 6; ++a) {
    double s = 0; const double* ja = jac + a * N;
    for (size_t i = 0; i < N; ++i) s += ja[i] * res[i];
    g[a] = s;
  }
  double s = 0;
  for (size_t i = 0; i < N; ++i) s += res[i] * res[i];
  *cost = 0.5 * s;
}

static void normal_equations_rowmajor(const double* jac, const double* res, size_t N, double H[21], double g[6], double* cost) {
  for (int k = 0; k < 21; ++k) H[k] = 0;
  for (int k = 0; k < 6; ++k) g[k] = 0;
  double s = 0;
  for (size_t i = 0; i < N; ++i) {
    const double* j = jac + i * 6;
    int k = 0;
    for (int a = 0; a < 6; ++a) for (int b = a; b < 6; ++b) H[k++] += j[a] * j[b];
    for (int a = 0; a < 6; ++a) g[a] += j[a] * res[i];
    s += res[i] * res[i];
  }
  *cost = 0.5 * s;
}

static void expand_sym(const double H[21], double M[36]) {
  int k = 0;
  for (int a = 0; a < 6; ++a) for (int b = a; b < 6; ++b) { M[a * 6 + b] = H[k]; M[b * 6 + a] = H[k]; ++k; }
}

/* Eigen: Matrix<double,6,6>::inverse() -> PartialPivLU then solve against the identity.
 * Returns the inverse in inv (row-major).  No singularity check (the reference has none). */
static void inverse6_partial_piv_lu(const double M[36], double inv[36]) {
  double A[36]; int perm[6];
  memcpy(A, M, sizeof(A));
  for (int i = 0; i < 6; ++i) perm[i] = i;
  for (int k = 0; k < 6; ++k) {
    int p = k; double best = fabs(A[k * 6 + k]);
    for (int i = k + 1; i < 6; ++i) if (fabs(A[i * 6 + k]) > best) { best = fabs(A[i * 6 + k]); p = i; }
    if (p != k) {
      for (int j = 0; j < 6; ++j) { double t = A[k * 6 + j]; A[k * 6 + j] = A[p * 6 + j]; A[p * 6 + j] = t; }
      int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
    }
    for (int i = k + 1; i < 6; ++i) {
      A[i * 6 + k] /= A[k * 6 + k];
      for (int j = k + 1; j < 6; ++j) A[i * 6 + j] -= A[i * 6 + k] * A[k * 6 + j];
    }
  }
  for (int c = 0; c < 6; ++c) {
    double yv[6];
    for (int i = 0; i < 6; ++i) {
      double s = (perm[i] == c) ? 1. : 0.;
      for (int j = 0; j < i; ++j) s -= A[i * 6 + j] * yv[j];
      yv[i] = s;
    }
    for (int i = 5; i >= 0; --i) {
      double s = yv[i];
      for (int j = i + 1; j < 6; ++j) s -= A[i * 6 + j] * inv[j * 6 + c];
      inv[i * 6 + c] = s / A[i * 6 + i];
    }
  }
}

static phovo_iter_stats* push_log(pho_oracle* o) {
  if (o->nlog == o->caplog) {
    o->caplog = o->caplog ? 2 * o->caplog : 64;
    o->log = (phovo_iter_stats*)realloc(o->log, sizeof(phovo_iter_stats) * o->caplog);
  }
  phovo_iter_stats* s = &o->log[o->nlog++];
  memset(s, 0, sizeof(*s));
  return s;
}

void pho_eval(pho_oracle* o, int level, const double state[6], phovo_iter_stats* out,
              double* residuals, double* jacobian) {
  const size_t N = (size_t)o->rows[level] * o->cols[level];
  double* res = (double*)calloc(N, sizeof(double));
  double* jac = (double*)calloc(N * 6, sizeof(double));
  memset(out, 0, sizeof(*out));
  out->level = level;
  memcpy(out->state_in, state, sizeof(double) * 6);
  memcpy(out->state_out, state, sizeof(double) * 6);
  if (o->cfg.mode == PHOVO_MODE_CERES) {
    out->num_valid = ceres_residuals_jacobians(o, level, state, res, jac, NULL, 1);
    normal_equations_rowmajor(jac, res, N, out->H, out->g, &out->cost);
    if (jacobian) memcpy(jacobian, jac, sizeof(double) * N * 6);
  } else {
    out->num_valid = analytic_residuals_jacobians(o, level, state, res, jac, NULL);
    normal_equations_colmajor(jac, res, N, out->H, out->g, &out->cost);
    if (jacobian) /* hand back row-major for a uniform interface */
      for (size_t i = 0; i < N; ++i) for (int k = 0; k < 6; ++k) jacobian[i * 6 + k] = jac[i + k * N];
  }
  double s = 0; for (int k = 0; k < 6; ++k) s += out->g[k] * out->g[k];
  out->grad_norm = sqrt(s);
  if (residuals) memcpy(residuals, res, sizeof(double) * N);
  free(res); free(jac);
}

void pho_winner_map(pho_oracle* o, int level, const double state[6], int32_t* out) {
  const size_t N = (size_t)o->rows[level] * o->cols[level];
  double* res = (double*)calloc(N, sizeof(double));
  double* jac = (double*)calloc(N * 6, sizeof(double));
  if (o->cfg.mode == PHOVO_MODE_CERES) ceres_residuals_jacobians(o, level, state, res, NULL, out, 1);
  else analytic_residuals_jacobians(o, level, state, res, jac, out);
  free(res); free(jac);
}

/* AN:500-563 Optimize + AN:376-426 TestTerminationCriteria */
static void optimize_analytic(pho_oracle* o) {
  for (int level = o->cfg.num_levels - 1; level >= 0; --level) {
    const size_t N = (size_t)o->rows[level] * o->cols[level];
    int iteration = 0;
    double g[6] = {0, 0, 0, 0, 0, 0}; /* m_Gradients persists across levels in the reference; it is
                                         only read after being written unless max_iters==0, where
                                         the iteration test fires first (AN:383) */
    double* res = NULL; double* jac = NULL;
    if (o->lean) { res = (double*)malloc(sizeof(double) * N); jac = (double*)malloc(sizeof(double) * N * 6); }
    while (1) {
      if (!o->lean) { /* AN:519-524: fresh, zeroed N x 1 and N x 6 every pass, even when nothing is computed */
        res = (double*)malloc(sizeof(double) * N); jac = (double*)malloc(sizeof(double) * N * 6);
        memset(res, 0, sizeof(double) * N); memset(jac, 0, sizeof(double) * N * 6);
      }
      if (o->cfg.max_num_iterations[level] > 0) { /* AN:526 */
        if (o->lean) { memset(res, 0, sizeof(double) * N); memset(jac, 0, sizeof(double) * N * 6); }
        phovo_iter_stats* s = push_log(o);
        s->level = level; s->iteration = iteration; s->accepted = 1;
        memcpy(s->state_in, o->state, sizeof(double) * 6);
        s->num_valid = analytic_residuals_jacobians(o, level, o->state, res, jac, NULL);
        normal_equations_colmajor(jac, res, N, s->H, s->g, &s->cost); /* AN:538-539 */
        memcpy(g, s->g, sizeof(g));
        double M[36], Minv[36], step[6];
        expand_sym(s->H, M);
        inverse6_partial_piv_lu(M, Minv);
        for (int a = 0; a < 6; ++a) { double t = 0; for (int b = 0; b < 6; ++b) t += Minv[a * 6 + b] * g[b]; step[a] = t; }
        for (int a = 0; a < 6; ++a) o->state[a] = o->state[a] - o->cfg.lambda_step[level] * step[a]; /* AN:539-540 */
        memcpy(s->state_out, o->state, sizeof(double) * 6);
        double n2 = 0; for (int a = 0; a < 6; ++a) n2 += g[a] * g[a];
        s->grad_norm = sqrt(n2);
      }
      if (!o->lean) { free(res); free(jac); res = jac = NULL; }
      iteration++; /* AN:547 */
      /* AN:376-392 */
      double n2 = 0; for (int a = 0; a < 6; ++a) n2 += g[a] * g[a];
      if (iteration >= o->cfg.max_num_iterations[level]) break;
      else if (sqrt(n2) < o->cfg.min_gradient_norm[level]) break;
    }
    if (o->lean) { free(res); free(jac); }
    o->iters_per_level[level] = o->cfg.max_num_iterations[level] > 0 ? iteration : 0;
  }
}

/* ---------------------------------------------------------------------------------------- */
/* Ceres-mode driver: CE:433-500 builds one ceres::Problem per level and calls ceres::Solve.  */
/* Ceres is third-party and absent; below is a restatement of its trust-region minimiser with */
/* the Levenberg-Marquardt strategy under the options the reference sets (CE:464-477),        */
/* defaults otherwise (jacobi_scaling on, min/max LM diagonal 1e-6/1e32, monotonic steps).    */
/* PARITY UNPINNED for the trajectory; the residual/Jacobian above are pinned.                */
/* ---------------------------------------------------------------------------------------- */
static int chol_solve6(const double M[36], const double b[6], double xout[6]) {
  double L[36]; memset(L, 0, sizeof(L));
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = M[i * 6 + j];
      for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k];
      if (i == j) { if (!(s > 0)) return 0; L[i * 6 + i] = sqrt(s); }
      else L[i * 6 + j] = s / L[j * 6 + j];
    }
  double yv[6];
  for (int i = 0; i < 6; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * yv[k]; yv[i] = s / L[i * 6 + i]; }
  for (int i = 5; i >= 0; --i) { double s = yv[i]; for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * xout[k]; xout[i] = s / L[i * 6 + i]; }
  return 1;
}

/* The restated trust-region loop, written against an evaluation callback so that oracle/_ref can
 * run the SAME loop on the reference's own functor (oracle/shim/ref_driver_ceres.cpp).
 * eval(user, x, want_jac, H, g, cost, count): H/g are only written when want_jac != 0.
 * new_entry(user) returns a zeroed log record for the iteration that starts.  Returns iterations. */
int pho_lm_minimize(const pho_lm_options* opt, pho_lm_eval_fn eval, void* eval_user,
                    pho_lm_entry_fn new_entry, void* entry_user, int level, double x[6]) {
  const int max_it = opt->max_num_iterations;
  double radius = opt->initial_trust_region_radius;
  const double max_radius = opt->max_trust_region_radius;
  const double min_radius = opt->min_trust_region_radius;
  const double eta = opt->min_relative_decrease;
  double decrease_factor = 2.0;
  double H[21], g[6], cost; int count;
  eval(eval_user, x, 1, H, g, &cost, &count);
  /* jacobi scaling: 1 / (1 + column norm), fixed from the first Jacobian */
  double scale[6]; { double Mfull[36]; expand_sym(H, Mfull);
    for (int a = 0; a < 6; ++a) scale[a] = 1.0 / (1.0 + sqrt(Mfull[a * 6 + a])); }
  double gmax = 0; for (int a = 0; a < 6; ++a) if (fabs(g[a]) > gmax) gmax = fabs(g[a]);
  int iteration = 0;
  if (!(gmax <= opt->gradient_tolerance)) {
    while (1) {
      if (iteration >= max_it) break;
      ++iteration;
      phovo_iter_stats* s = new_entry(entry_user);
      s->level = level; s->iteration = iteration - 1; s->num_valid = count; s->cost = cost; s->radius = radius;
      memcpy(s->H, H, sizeof(H)); memcpy(s->g, g, sizeof(g)); memcpy(s->state_in, x, sizeof(double) * 6);
      memcpy(s->state_out, x, sizeof(double) * 6);
      { double n2 = 0; for (int a = 0; a < 6; ++a) n2 += g[a] * g[a]; s->grad_norm = sqrt(n2); }
      /* scaled system */
      double M[36], Ms[36], gs[6];
      expand_sym(H, M);
      for (int a = 0; a < 6; ++a) { gs[a] = g[a] * scale[a]; for (int b = 0; b < 6; ++b) Ms[a * 6 + b] = M[a * 6 + b] * scale[a] * scale[b]; }
      double A[36]; memcpy(A, Ms, sizeof(A));
      for (int a = 0; a < 6; ++a) {
        double d = Ms[a * 6 + a]; if (d < 1e-6) d = 1e-6; if (d > 1e32) d = 1e32;
        A[a * 6 + a] += d / radius;
      }
      double step[6]; int ok = chol_solve6(A, gs, step);
      for (int a = 0; a < 6; ++a) step[a] = -step[a];
      double model_cost_change = 0;
      if (ok) {
        /* -(J d)^T (r + J d / 2) = -(d^T g + 0.5 d^T M d) */
        double dg = 0, dMd = 0;
        for (int a = 0; a < 6; ++a) { dg += step[a] * gs[a]; double t = 0; for (int b = 0; b < 6; ++b) t += Ms[a * 6 + b] * step[b]; dMd += step[a] * t; }
        model_cost_change = -(dg + 0.5 * dMd);
        for (int a = 0; a < 6; ++a) if (!isfinite(step[a])) ok = 0;
      }
      if (!ok || !(model_cost_change > 0)) {
        /* max_num_consecutive_invalid_steps = 0 (CE:477): the first invalid step terminates */
        s->accepted = 0;
        break;
      }
      double delta[6], xn[6], step_norm = 0, x_norm = 0;
      for (int a = 0; a < 6; ++a) { delta[a] = step[a] * scale[a]; xn[a] = x[a] + delta[a]; step_norm += delta[a] * delta[a]; x_norm += x[a] * x[a]; }
      step_norm = sqrt(step_norm); x_norm = sqrt(x_norm);
      double new_cost; int ncount; double Hd[21], gd[6];
      eval(eval_user, xn, 0, Hd, gd, &new_cost, &ncount);
      const double ptol = opt->parameter_tolerance;
      if (step_norm <= ptol * (x_norm + ptol)) { s->accepted = 0; break; }
      double cost_change = cost - new_cost;
      if (fabs(cost_change) < opt->function_tolerance * cost) { s->accepted = 0; break; }
      double rho = cost_change / model_cost_change;
      if (rho > eta) {
        memcpy(x, xn, sizeof(double) * 6);
        s->accepted = 1; memcpy(s->state_out, x, sizeof(double) * 6);
        eval(eval_user, x, 1, H, g, &cost, &count);
        gmax = 0; for (int a = 0; a < 6; ++a) if (fabs(g[a]) > gmax) gmax = fabs(g[a]);
        if (gmax <= opt->gradient_tolerance) break;
        double t = 2.0 * rho - 1.0;
        double f = 1.0 - t * t * t; if (f < 1.0 / 3.0) f = 1.0 / 3.0;
        radius = radius / f; if (radius > max_radius) radius = max_radius;
        decrease_factor = 2.0;
      } else {
        s->accepted = 0;
        radius = radius / decrease_factor; decrease_factor *= 2.0;
      }
      if (radius < min_radius) break;
    }
  }
  return iteration;
}

/* J^T J (upper 21), J^T r and the cost from a row-major N x 6 Jacobian: the sums the loop above needs */
void pho_normal_equations_rowmajor(const double* jac, const double* res, size_t n, double H[21], double g[6], double* cost) {
  normal_equations_rowmajor(jac, res, n, H, g, cost);
}

typedef struct { pho_oracle* o; int level; double* res; double* jac; } ceres_level_ctx;

static void ceres_level_eval(void* user, const double st[6], int want_jac, double H[21], double g[6], double* cost, int* count) {
  ceres_level_ctx* c = (ceres_level_ctx*)user;
  const size_t N = (size_t)c->o->rows[c->level] * c->o->cols[c->level];
  *count = ceres_residuals_jacobians(c->o, c->level, st, c->res, want_jac ? c->jac : NULL, NULL, want_jac);
  if (want_jac) normal_equations_rowmajor(c->jac, c->res, N, H, g, cost);
  else { double s = 0; for (size_t i = 0; i < N; ++i) s += c->res[i] * c->res[i]; *cost = 0.5 * s; }
}

static phovo_iter_stats* ceres_level_entry(void* user) { return push_log((pho_oracle*)user); }

static void optimize_ceres(pho_oracle* o) {
  for (int level = o->cfg.num_levels - 1; level >= 0; --level) {
    o->iters_per_level[level] = 0;
    if (!(o->cfg.max_num_iterations[level] > 0)) continue; /* CE:437 */
    const size_t N = (size_t)o->rows[level] * o->cols[level];
    ceres_level_ctx ctx = { o, level, (double*)malloc(sizeof(double) * N), (double*)malloc(sizeof(double) * N * 6) };
    pho_lm_options opt;                                    /* CE:464-477 */
    opt.max_num_iterations = o->cfg.max_num_iterations[level];
    opt.function_tolerance = o->cfg.function_tolerance[level];
    opt.gradient_tolerance = o->cfg.gradient_tolerance[level];
    opt.parameter_tolerance = o->cfg.parameter_tolerance[level];
    opt.initial_trust_region_radius = o->cfg.initial_trust_region_radius[level];
    opt.max_trust_region_radius = o->cfg.max_trust_region_radius[level];
    opt.min_trust_region_radius = o->cfg.min_trust_region_radius[level];
    opt.min_relative_decrease = o->cfg.min_relative_decrease[level];
    o->iters_per_level[level] = pho_lm_minimize(&opt, ceres_level_eval, &ctx, ceres_level_entry, o, level, o->state);
    free(ctx.res); free(ctx.jac);
  }
}

void pho_optimize(pho_oracle* o) {
  o->nlog = 0;
  double t0 = now_sec();
  if (o->cfg.mode == PHOVO_MODE_CERES) optimize_ceres(o);
  else optimize_analytic(o);
  o->optimize_seconds = now_sec() - t0;
}

int pho_num_iter_stats(const pho_oracle* o) { return o->nlog; }
int pho_get_iter_stats(const pho_oracle* o, int index, phovo_iter_stats* out) {
  if (index < 0 || index >= o->nlog) return -1;
  *out = o->log[index];
  return 0;
}

/* ---------------------------------------------------------------------------------------- */
/* batch helper for the CPU baseline                                                         */
/* ---------------------------------------------------------------------------------------- */
typedef struct {
  const phovo_config* cfg; const double* K; int num_pairs, rows, cols;
  const uint8_t* gray0; const double* depth0; const uint8_t* gray1;
  int lean; double* states; int32_t* iterations;
  int tid, nthreads; double opt_seconds;
} batch_job;

static void* batch_worker(void* arg) {
  batch_job* j = (batch_job*)arg;
  pho_oracle* o = pho_create();
  pho_set_config(o, j->cfg);
  pho_set_options(o, 0, j->lean);
  pho_set_intrinsics(o, j->K);
  const size_t n = (size_t)j->rows * j->cols;
  const double zero[6] = {0, 0, 0, 0, 0, 0};
  for (int p = j->tid; p < j->num_pairs; p += j->nthreads) {
    pho_set_source(o, j->gray0 + p * n, j->cols, j->depth0 + p * n, sizeof(double) * j->cols, j->rows, j->cols);
    pho_set_target(o, j->gray1 + p * n, j->cols, j->rows, j->cols);
    pho_set_initial_state(o, zero);
    pho_optimize(o);
    j->opt_seconds += o->optimize_seconds;
    if (j->states) pho_get_state(o, j->states + 6 * (size_t)p);
    if (j->iterations)
      for (int l = 0; l < j->cfg->num_levels; ++l) j->iterations[(size_t)p * PHOVO_MAX_LEVELS + l] = o->iters_per_level[l];
  }
  pho_destroy(o);
  return NULL;
}

double pho_align_batch(const phovo_config* cfg, const double K[9], int num_pairs, int rows, int cols,
                       const uint8_t* gray0, const double* depth0, const uint8_t* gray1,
                       int num_threads, int lean, double* states, int32_t* iterations,
                       double* optimize_seconds) {
  if (num_threads < 1) num_threads = 1;
  if (num_threads > 256) num_threads = 256;
  batch_job jobs[256]; pthread_t th[256];
  double t0 = now_sec();
  for (int t = 0; t < num_threads; ++t) {
    batch_job j = {cfg, K, num_pairs, rows, cols, gray0, depth0, gray1, lean, states, iterations, t, num_threads, 0.0};
    jobs[t] = j;
    pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
  }
  double opt = 0;
  for (int t = 0; t < num_threads; ++t) { pthread_join(th[t], NULL); opt += jobs[t].opt_seconds; }
  if (optimize_seconds) *optimize_seconds = opt;
  return now_sec() - t0;
}
