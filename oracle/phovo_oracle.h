/*
 * phovo_oracle.h -- CPU oracle for the photoconsistency alignment hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  The product (libphovo_b200.so) never
 * links, loads or calls it.
 *
 * It is a single-threaded, double-precision C restatement of the reference's algorithm
 * (paths relative to the reference tree):
 *   AN = phovo/include/CPhotoconsistencyOdometryAnalytic.h
 *   CE = phovo/include/CPhotoconsistencyOdometryCeres.h
 *   SA = third_party/sample.h          JE = third_party/jet_extras.h
 *   BASE = phovo/include/CPhotoconsistencyOdometry.h
 *
 * PARITY PINNING.  The reference ships no tests, golden vectors or sample data, and its build
 * needs OpenCV, Eigen, Ceres and Boost, none of which is installed here.  The oracle is pinned
 * three ways, all in tests/ (-m "not gpu"):
 *   1. AGAINST THE REFERENCE'S OWN SOURCE (analytic path): oracle/_ref/libphovo_ref.so is
 *      CPhotoconsistencyOdometryAnalytic.h compiled UNMODIFIED from /root/reference against minimal
 *      stand-ins for the cv:: / Eigen:: types it touches (oracle/shim/, oracle/Makefile).  Run on
 *      the same inputs the oracle agrees with it to the last bit on the tested pairs (per-iteration
 *      J^T J / J^T r, iteration counts, final state; tests/test_reference_source_pins.py), and
 *      tests/golden/ref_*.npz are outputs of that library (tests/golden/make_reference_golden.py).
 *      The stand-ins replace third-party code only; behind them the image arithmetic is (2).
 *   2. the third-party image arithmetic (cv::resize, cv::Scharr, cv::GaussianBlur, convertTo) is
 *      checked against OpenCV 4.13 itself through python cv2, live and via tests/golden/cv2_ops.npz;
 *   3. an independently written numpy restatement (oracle/np_restatement.py) and, for the
 *      Jacobian, a symbolic re-derivation of phovo/Maxima/derivatives_photoconsistency.wxm.
 * Ceres mode: the residual functor (CE:156-269, sample.h, jet_extras.h) needs Ceres' Jet type and
 * cannot be compiled here; its restatement is pinned by (2)/(3) and by finite differences only,
 * and the Ceres solver itself (trust-region LM) is third-party and absent: the Ceres-mode LM
 * trajectory is "parity unpinned".
 */
#ifndef PHOVO_ORACLE_H_
#define PHOVO_ORACLE_H_

#include <stddef.h>
#include <stdint.h>
#include "../include/phovo_b200.h" /* phovo_config, phovo_iter_stats (POD only) */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pho_oracle pho_oracle;

/* ---- stand-alone image arithmetic (each follows the OpenCV call the reference makes) ---- */
/* Mat::convertTo(CV_64F, 1/255)  AN:471,484 */
void pho_convert_u8(const uint8_t* src, size_t step, int rows, int cols, double* dst);
/* output size of cv::resize(img, Size(0,0), f, f) with f = 2^-level  AN:132 */
void pho_level_size(int rows, int cols, int level, int* out_rows, int* out_cols);
/* cv::resize(..., INTER_LINEAR) by 2^-level from the ORIGINAL image  AN:132 */
void pho_resize_level(const double* src, int rows, int cols, int level, double* dst);
/* cv::GaussianBlur(img, img, Size(k,k), sigma) BORDER_REFLECT_101  AN:146-147 (applied once) */
void pho_gaussian_blur(double* img, int rows, int cols, int ksize, double sigma);
/* cv::Scharr(src, dst, CV_64F, dx, dy, scale, 0, BORDER_DEFAULT)  AN:181-187 */
void pho_scharr(const double* src, int rows, int cols, int dx, int dy, double scale, double* dst);

/* ---- the solver object, mirroring the reference call sequence ---- */
pho_oracle* pho_create(void);
void pho_destroy(pho_oracle* o);
void pho_set_config(pho_oracle* o, const phovo_config* cfg);
/* oracle-only switches:
 *  storage_f32  : 1 = round every pyramid/gradient level to float after building it (emulates the
 *                 device's fp32 image storage so integer decisions can be compared exactly);
 *                 2 = the same but the depth pyramid stays double (the device's Ceres-mode layout)
 *  lean         : skip the reference's per-pass allocation + setZero of the N x 7 doubles and the
 *                 materialised Jacobian (AN:519-524); identical results, used only for timing   */
void pho_set_options(pho_oracle* o, int storage_f32, int lean);
void pho_set_intrinsics(pho_oracle* o, const double K[9]);
/* AN:466-476; depth in metres, double, row stride in BYTES */
void pho_set_source(pho_oracle* o, const uint8_t* gray, size_t gray_step,
                    const double* depth, size_t depth_step, int rows, int cols);
/* AN:479-491 */
void pho_set_target(pho_oracle* o, const uint8_t* gray, size_t gray_step, int rows, int cols);
void pho_set_initial_state(pho_oracle* o, const double state[6]);
/* AN:500-563 (analytic modes) or the restated LM of CE:433-500 (Ceres mode) */
void pho_optimize(pho_oracle* o);
void pho_get_state(const pho_oracle* o, double state[6]);
void pho_get_rt(const pho_oracle* o, double rt[16]);          /* BASE:47-71 */
int  pho_num_iter_stats(const pho_oracle* o);
int  pho_get_iter_stats(const pho_oracle* o, int index, phovo_iter_stats* out);
/* which: 0 I0, 1 D0, 2 I1, 3 Gx1, 4 Gy1; returns pointer into the oracle (double, rows*cols) */
const double* pho_level_image(const pho_oracle* o, int which, int level, int* rows, int* cols);

/* one evaluation at `state` on `level` without stepping.
 * analytic modes: AN:191-367 + AN:538-540 products.  Ceres mode: CE:156-269 residual and Jacobian.
 * residuals (N) / jacobian (N x 6 row-major) may be NULL. */
void pho_eval(pho_oracle* o, int level, const double state[6], phovo_iter_stats* out,
              double* residuals, double* jacobian);
/* winner map of the residual scatter at `state` (AN:358 / CE:261): for each target slot the
 * source index that wrote last, or -1.  out has rows*cols ints. */
void pho_winner_map(pho_oracle* o, int level, const double state[6], int32_t* out);

/* ---- batch helper for the CPU baseline: aligns `num_pairs` pairs with `num_threads` worker
 * threads (one oracle object per thread; each alignment itself is single-threaded like the
 * reference).  Layout as phovo_batch_align, depth is double.  Returns wall seconds; if
 * optimize_seconds is not NULL it receives the summed time spent inside pho_optimize only
 * (what the reference apps' TickMeter brackets). */
double pho_align_batch(const phovo_config* cfg, const double K[9], int num_pairs, int rows, int cols,
                       const uint8_t* gray0, const double* depth0, const uint8_t* gray1,
                       int num_threads, int lean, double* states, int32_t* iterations,
                       double* optimize_seconds);

void pho_state_to_rt(const double s[6], double rt[16]);

/* ---- the restated Levenberg-Marquardt loop of Ceres mode (CE:433-500 hands each level to
 * ceres::Solve; Ceres is third-party and absent, see the header comment), exposed so that
 * oracle/_ref can run the same loop on the reference's OWN residual functor. ---- */
typedef struct {
  int max_num_iterations;                                    /* CE:465 */
  double function_tolerance, gradient_tolerance, parameter_tolerance;          /* CE:468-470 */
  double initial_trust_region_radius, max_trust_region_radius, min_trust_region_radius; /* CE:471-473 */
  double min_relative_decrease;                              /* CE:474 */
} pho_lm_options;
typedef void (*pho_lm_eval_fn)(void* user, const double x[6], int want_jac, double H[21], double g[6], double* cost, int* count);
typedef phovo_iter_stats* (*pho_lm_entry_fn)(void* user);
int pho_lm_minimize(const pho_lm_options* opt, pho_lm_eval_fn eval, void* eval_user,
                    pho_lm_entry_fn new_entry, void* entry_user, int level, double x[6]);
void pho_normal_equations_rowmajor(const double* jac, const double* res, size_t n, double H[21], double g[6], double* cost);

#ifdef __cplusplus
}
#endif
#endif
