// kernels_pyramid.cu -- frame setup for the general (any resolution, any config) path:
// K1 pyramid level from the full-resolution source, K2b Gaussian blur, K2 Scharr.  All level images are fp64, bit-identical
// to the reference's cv::Mat_<double> pyramids as OpenCV's portable C++ code computes them (same
// operation order, no FMA contraction; pinned against cv2 with its IPP / SIMD dispatch switched off).
// Replaces CPhotoconsistencyOdometryAnalytic.h:115-189 (BuildPyramid / BuildDerivativesPyramids).
// The batched path has its own fused, shared-memory version (kernels_batch.cu).
#include <math.h>

#include "phovo_kernels.h"

namespace phovo {
namespace {

__device__ __forceinline__ int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
  return p;
}

template <typename T>
__device__ __forceinline__ double load_src(const void* base, size_t step_bytes, int r, int c, double scale);
template <>
__device__ __forceinline__ double load_src<uint8_t>(const void* base, size_t step, int r, int c, double scale) {
  return __dmul_rn((double)__ldg((const uint8_t*)base + (size_t)r * step + c), scale);
}
template <>
__device__ __forceinline__ double load_src<uint16_t>(const void* base, size_t step, int r, int c, double scale) {
  return __dmul_rn((double)__ldg((const uint16_t*)((const char*)base + (size_t)r * step) + c), scale);
}
template <>
__device__ __forceinline__ double load_src<float>(const void* base, size_t step, int r, int c, double) {
  return (double)__ldg((const float*)((const char*)base + (size_t)r * step) + c);
}
template <>
__device__ __forceinline__ double load_src<double>(const void* base, size_t step, int r, int c, double) {
  return __ldg((const double*)((const char*)base + (size_t)r * step) + c);
}

// One axis of OpenCV's INTER_LINEAR table: f = (float)((d+0.5)*scale - 0.5); s = floor(f); f -= s.
__device__ __forceinline__ void linear_axis(int d, double scale, int ssize, bool is_x, int& s0, float& w1) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (s < 0) { f = 0.f; s = 0; }
  if (is_x && s >= ssize - 1) { f = 0.f; s = ssize - 1; }
  s0 = s; w1 = f;
}

// K1.  One thread per output pixel.  For the exact factor 2^-level, level >= 2, the four taps are the
// central 2x2 of the pixel's 2^level cell with weights 1/2 (horizontal pair first, then vertical) --
// cv::resize INTER_LINEAR from the ORIGINAL image, AN:132; level 1 is OpenCV's area path (below).  A warp reads one contiguous span of
// each of two source rows, so every 32-byte sector fetched is used by the warp.
// value of level pixel (y, x): the conversion of the source pixel at level 0, else cv::resize's four taps
template <typename T>
__device__ __forceinline__ double level_value(const void* __restrict__ src, size_t step, double src_scale,
                                              int rows, int cols, int level, double scale, int y, int x) {
  if (level == 0) return load_src<T>(src, step, y, x, src_scale);
  if (level == 1) {
    // cv::resize by exactly 1/2 is NOT bilinear: resize.cpp switches INTER_LINEAR to INTER_AREA when
    // both integer scale factors are 2 (resizeAreaFast_Invoker<double,double>): a whole 2x2 cell is
    // ((S00 + S01) + S10) + S11 times 0.25f; a cell cut by the right / bottom edge (sizes = 3 mod 4:
    // the output size is cvRound(size / 2)) sums its taps inside the image in raster order and
    // divides by their count IN SINGLE PRECISION ((float)sum / count).
    const int sy0 = 2 * y, sx0 = 2 * x;
    if (sy0 + 2 <= rows && x < cols / 2) {
      const double a = load_src<T>(src, step, sy0, sx0, src_scale), b = load_src<T>(src, step, sy0, sx0 + 1, src_scale);
      const double c = load_src<T>(src, step, sy0 + 1, sx0, src_scale), d = load_src<T>(src, step, sy0 + 1, sx0 + 1, src_scale);
      return __dmul_rn(__dadd_rn(__dadd_rn(__dadd_rn(a, b), c), d), 0.25);
    }
    double sum = 0.; int count = 0;
    for (int sy = 0; sy < 2 && sy0 + sy < rows; ++sy)
      for (int sx = 0; sx < 2 && sx0 + sx < cols; ++sx) { sum = __dadd_rn(sum, load_src<T>(src, step, sy0 + sy, sx0 + sx, src_scale)); ++count; }
    return (double)__fdiv_rn(__double2float_rn(sum), (float)count);
  }
  int sx, sy; float fx, fy;
  linear_axis(x, scale, cols, true, sx, fx);
  linear_axis(y, scale, rows, false, sy, fy);
  const int y0 = min(max(sy, 0), rows - 1), y1 = min(max(sy + 1, 0), rows - 1);
  double r0, r1;
  if (sx + 1 < cols) {
    const double a0 = (double)(1.f - fx), a1 = (double)fx;
    r0 = __dadd_rn(__dmul_rn(load_src<T>(src, step, y0, sx, src_scale), a0), __dmul_rn(load_src<T>(src, step, y0, sx + 1, src_scale), a1));
    r1 = __dadd_rn(__dmul_rn(load_src<T>(src, step, y1, sx, src_scale), a0), __dmul_rn(load_src<T>(src, step, y1, sx + 1, src_scale), a1));
  } else {
    r0 = load_src<T>(src, step, y0, sx, src_scale);
    r1 = load_src<T>(src, step, y1, sx, src_scale);
  }
  const double b0 = (double)(1.f - fy), b1 = (double)fy;
  return __dadd_rn(__dmul_rn(r0, b0), __dmul_rn(r1, b1));
}

template <typename T>
__global__ void __launch_bounds__(256) k_build_level(const void* __restrict__ src, size_t step, double src_scale,
                                                     int rows, int cols, int level, double scale,
                                                     double* __restrict__ dst, int orows, int ocols) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= ocols || y >= orows) return;
  dst[(size_t)y * ocols + x] = level_value<T>(src, step, src_scale, rows, cols, level, scale, y, x);
}

// K1+K2 fused over ALL active levels of one frame: one thread per level pixel (levels laid end to end
// in the thread index).  Every level is resized from the ORIGINAL image (AN:132), so the levels do not
// depend on each other; with GRAD the thread also evaluates its eight level neighbours (the same
// four-tap function, so the same bits as the stored image; the taps come from L1/L2) and applies
// cv::Scharr's arithmetic (AN:181-187) -- one launch instead of two per level, which is what a
// VO frame's set-up time consists of.
template <typename T, bool GRAD>
__global__ void __launch_bounds__(256) k_build_levels(const void* __restrict__ src, size_t step, double src_scale,
                                                      int rows, int cols, const __grid_constant__ PyramidLevels P) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= P.px_offset[P.num]) return;
  int a = 0;
  while (j >= P.px_offset[a + 1]) ++a;
  const int q = j - P.px_offset[a];
  const int oc = P.ocols[a], orr = P.orows[a], level = P.level[a];
  const int y = q / oc, x = q - y * oc;
  const double scale = (double)(1 << level);
  if (!GRAD) {
    P.dst[a][q] = level_value<T>(src, step, src_scale, rows, cols, level, scale, y, x);
    return;
  }
  const int ym = reflect101(y - 1, orr), yp = reflect101(y + 1, orr), xm = reflect101(x - 1, oc), xp = reflect101(x + 1, oc);
  double t[3][3];
  const int ys[3] = {ym, y, yp}, xs[3] = {xm, x, xp};
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) t[dy][dx] = level_value<T>(src, step, src_scale, rows, cols, level, scale, ys[dy], xs[dx]);
  const double ks0 = P.ks0[a], ks1 = P.ks1[a];
  // dx: row filter [-1 0 1] on three rows, then column filter centre-first
  const double dm = __dsub_rn(t[0][2], t[0][0]);
  const double d0 = __dsub_rn(t[1][2], t[1][0]);
  const double dp = __dsub_rn(t[2][2], t[2][0]);
  const double gx = __dadd_rn(__dmul_rn(ks1, d0), __dmul_rn(ks0, __dadd_rn(dp, dm)));
  // dy: row filter [3 10 3]*scale left-to-right on rows y-1 and y+1, then column [-1 0 1]
  const double sm = __dadd_rn(__dadd_rn(__dmul_rn(ks0, t[0][0]), __dmul_rn(ks1, t[0][1])), __dmul_rn(ks0, t[0][2]));
  const double sp = __dadd_rn(__dadd_rn(__dmul_rn(ks0, t[2][0]), __dmul_rn(ks1, t[2][1])), __dmul_rn(ks0, t[2][2]));
  P.dst[a][q] = t[1][1];
  P.gx[a][q] = gx;
  P.gy[a][q] = __dsub_rn(sp, sm);
}

struct GaussTaps { double k[32]; int n; };

// K2b row pass: generic cv RowFilter order (taps accumulated left to right).
__global__ void __launch_bounds__(256) k_gauss_rows(const double* __restrict__ src, double* __restrict__ dst, int rows, int cols, GaussTaps t) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= cols) return;
  const double* S = src + (size_t)r * cols;
  const int a = t.n / 2;
  double s = __dmul_rn(t.k[0], S[reflect101(c - a, cols)]);
  for (int k = 1; k < t.n; ++k) s = __dadd_rn(s, __dmul_rn(t.k[k], S[reflect101(c - a + k, cols)]));
  dst[(size_t)r * cols + c] = s;
}
// K2b column pass: cv SymmColumnFilter order (centre tap, then symmetric pairs).
__global__ void __launch_bounds__(256) k_gauss_cols(const double* __restrict__ src, double* __restrict__ dst, int rows, int cols, GaussTaps t) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= cols) return;
  const int a = t.n / 2;
  double s = __dmul_rn(t.k[a], src[(size_t)r * cols + c]);
  for (int k = 1; k <= a; ++k) {
    const double p = src[(size_t)reflect101(r + k, rows) * cols + c], q = src[(size_t)reflect101(r - k, rows) * cols + c];
    s = __dadd_rn(s, __dmul_rn(t.k[a + k], __dadd_rn(p, q)));
  }
  dst[(size_t)r * cols + c] = s;
}

// K2: Scharr in x and y with cv::Scharr's separable evaluation order (derivative kernel [-1 0 1],
// smoothing kernel [3 10 3]*scale).  32x8 tiles with a
// 1-pixel halo staged in shared memory so every level pixel is read from global memory once
// (plus halo), then 18 shared-memory reads per pixel.
#define SCH_TW 32
#define SCH_TH 8
__global__ void __launch_bounds__(SCH_TW * SCH_TH) k_scharr_store(const double* __restrict__ img, int rows, int cols,
                                                                   double ks0, double ks1,
                                                                   double* __restrict__ Gx, double* __restrict__ Gy) {
  __shared__ double tile[SCH_TH + 2][SCH_TW + 2];
  const int x0 = blockIdx.x * SCH_TW, y0 = blockIdx.y * SCH_TH;
  for (int i = threadIdx.y * SCH_TW + threadIdx.x; i < (SCH_TH + 2) * (SCH_TW + 2); i += SCH_TW * SCH_TH) {
    const int ty = i / (SCH_TW + 2), tx = i % (SCH_TW + 2);
    const int gy = reflect101(min(y0 + ty - 1, rows), rows), gx = reflect101(min(x0 + tx - 1, cols), cols);
    tile[ty][tx] = img[(size_t)gy * cols + gx];
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= cols || y >= rows) return;
  const int tx = threadIdx.x + 1, ty = threadIdx.y + 1;
  // dx: row filter [-1 0 1] on three rows, then column filter centre-first
  const double dm = __dsub_rn(tile[ty - 1][tx + 1], tile[ty - 1][tx - 1]);
  const double d0 = __dsub_rn(tile[ty][tx + 1], tile[ty][tx - 1]);
  const double dp = __dsub_rn(tile[ty + 1][tx + 1], tile[ty + 1][tx - 1]);
  const double gx = __dadd_rn(__dmul_rn(ks1, d0), __dmul_rn(ks0, __dadd_rn(dp, dm)));
  // dy: row filter [3 10 3]*scale left-to-right on rows y-1 and y+1, then column [-1 0 1]
  const double sm = __dadd_rn(__dadd_rn(__dmul_rn(ks0, tile[ty - 1][tx - 1]), __dmul_rn(ks1, tile[ty - 1][tx])), __dmul_rn(ks0, tile[ty - 1][tx + 1]));
  const double sp = __dadd_rn(__dadd_rn(__dmul_rn(ks0, tile[ty + 1][tx - 1]), __dmul_rn(ks1, tile[ty + 1][tx])), __dmul_rn(ks0, tile[ty + 1][tx + 1]));
  const double gy = __dsub_rn(sp, sm);
  const size_t o = (size_t)y * cols + x;

  Gx[o] = gx;
  Gy[o] = gy;
}

}  // namespace

int launch_build_level(cudaStream_t stream, const void* src, int src_type, size_t step, double src_scale,
                       int rows, int cols, int level, double* dst, int orows, int ocols) {
  dim3 block(64, 4), grid((ocols + 63) / 64, (orows + 3) / 4);
  const double scale = ldexp(1.0, level);
  switch (src_type) {
    case SRC_U8:  k_build_level<uint8_t><<<grid, block, 0, stream>>>(src, step, src_scale, rows, cols, level, scale, dst, orows, ocols); break;
    case SRC_U16: k_build_level<uint16_t><<<grid, block, 0, stream>>>(src, step, src_scale, rows, cols, level, scale, dst, orows, ocols); break;
    case SRC_F32: k_build_level<float><<<grid, block, 0, stream>>>(src, step, src_scale, rows, cols, level, scale, dst, orows, ocols); break;
    default:      k_build_level<double><<<grid, block, 0, stream>>>(src, step, src_scale, rows, cols, level, scale, dst, orows, ocols); break;
  }
  return 1;
}

int launch_build_levels(cudaStream_t stream, const void* src, int src_type, size_t step, double src_scale,
                        int rows, int cols, const PyramidLevels& P, bool gradients) {
  const int total = P.px_offset[P.num];
  if (total <= 0) return 0;
  const int grid = (total + 255) / 256;
#define PHOVO_BUILD_LEVELS(T)                                                                                   \
  if (gradients) k_build_levels<T, true><<<grid, 256, 0, stream>>>(src, step, src_scale, rows, cols, P);        \
  else k_build_levels<T, false><<<grid, 256, 0, stream>>>(src, step, src_scale, rows, cols, P)
  switch (src_type) {
    case SRC_U8:  PHOVO_BUILD_LEVELS(uint8_t); break;
    case SRC_U16: PHOVO_BUILD_LEVELS(uint16_t); break;
    case SRC_F32: PHOVO_BUILD_LEVELS(float); break;
    default:      PHOVO_BUILD_LEVELS(double); break;
  }
#undef PHOVO_BUILD_LEVELS
  return 1;
}

int launch_gaussian_blur(cudaStream_t stream, double* img, double* tmp, int rows, int cols, int ksize, double sigma) {
  if (ksize <= 1) return 0;
  GaussTaps t; t.n = ksize;
  // cv::getGaussianKernel(ksize, sigma, CV_64F)
  const double sigmaX = sigma > 0 ? sigma : ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
  const double scale2X = -0.5 / (sigmaX * sigmaX);
  double sum = 0;
  for (int i = 0; i < ksize; ++i) { const double x = i - (ksize - 1) * 0.5; t.k[i] = exp(scale2X * x * x); sum += t.k[i]; }
  sum = 1. / sum;
  for (int i = 0; i < ksize; ++i) t.k[i] *= sum;
  dim3 block(256), grid((cols + 255) / 256, rows);
  k_gauss_rows<<<grid, block, 0, stream>>>(img, tmp, rows, cols, t);
  k_gauss_cols<<<grid, block, 0, stream>>>(tmp, img, rows, cols, t);
  return 2;
}

int launch_scharr_store(cudaStream_t stream, const double* img, int rows, int cols, double scale,
                        double* Gx, double* Gy) {
  double ks0 = 3., ks1 = 10.;
  if (scale != 1.) { ks0 *= scale; ks1 *= scale; }
  dim3 block(SCH_TW, SCH_TH), grid((cols + SCH_TW - 1) / SCH_TW, (rows + SCH_TH - 1) / SCH_TH);
  k_scharr_store<<<grid, block, 0, stream>>>(img, rows, cols, ks0, ks1, Gx, Gy);
  return 1;
}


// ---------------------------------------------------------------------------------------------
// helpers of the photometric + depth solver (BiObjective.h:213-239, 299)
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) k_scale(const double* __restrict__ src, double alpha, double* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) dst[i] = __dmul_rn(src[i], alpha);
}
// one CTA, fixed order: thread t sums elements t, t+1024, ...; then a fixed tree over the threads
__global__ void __launch_bounds__(1024) k_mean_ratio(const double* __restrict__ a, const double* __restrict__ b, size_t n, double* out) {
  __shared__ double sa[1024], sb[1024];
  double x = 0., y = 0.;
  for (size_t i = threadIdx.x; i < n; i += 1024) { x += a[i]; y += b[i]; }
  sa[threadIdx.x] = x; sb[threadIdx.x] = y;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sb[threadIdx.x] += sb[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (sa[0] / (double)n) / (sb[0] / (double)n);   // cv::mean(I).val[0] / cv::mean(D).val[0]
}
}  // namespace

int launch_scale(cudaStream_t stream, const double* src, double alpha, double* dst, size_t n) {
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  k_scale<<<blocks > 0 ? blocks : 1, 256, 0, stream>>>(src, alpha, dst, n);
  return 1;
}
int launch_mean_ratio(cudaStream_t stream, const double* a, const double* b, size_t n, double* out) {
  k_mean_ratio<<<1, 1024, 0, stream>>>(a, b, n, out);
  return 1;
}

// ---------------------------------------------------------------------------------------------
// phovo::warpImage, CPhotoconsistencyOdometry.h:73-134 (post-hoc visualisation of both apps)
// ---------------------------------------------------------------------------------------------
namespace {
struct WarpImgParams { double R[12]; double fx, fy, ox, oy, inv_fx, inv_fy; };

template <typename DT>
__global__ void __launch_bounds__(256) k_warp_splat(const uint8_t* __restrict__ gray, size_t gray_step, const DT* __restrict__ depth,
                                                    size_t depth_step, double depth_scale, int rows, int cols,
                                                    const __grid_constant__ WarpImgParams W, unsigned long long* __restrict__ keys) {
  const size_t n = (size_t)rows * cols;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const int r = (int)(i / cols), c = (int)(i - (size_t)r * cols);
    const DT* drow = (const DT*)((const char*)depth + (size_t)r * depth_step);
    double d = (double)drow[c];
    if (sizeof(DT) == 2) d = __dmul_rn(d, depth_scale);
    if (!(d > 0.)) continue;                                                           // BASE:106
    // BASE:109-112, fp64, the reference's operation order, no FMA contraction
    const double px = __dmul_rn(__dmul_rn(__dsub_rn((double)c, W.ox), d), W.inv_fx);
    const double py = __dmul_rn(__dmul_rn(__dsub_rn((double)r, W.oy), d), W.inv_fy);
    // Eigen 4x4 * 4x1 (BASE:115): sum over k = 0..3 in order, the homogeneous 1 multiplies the translation
    const double X = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(W.R[0], px), __dmul_rn(W.R[1], py)), __dmul_rn(W.R[2], d)), W.R[3]);
    const double Y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(W.R[4], px), __dmul_rn(W.R[5], py)), __dmul_rn(W.R[6], d)), W.R[7]);
    const double Z = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(W.R[8], px), __dmul_rn(W.R[9], py)), __dmul_rn(W.R[10], d)), W.R[11]);
    const double tc = __dadd_rn(__ddiv_rn(__dmul_rn(X, W.fx), Z), W.ox);               // BASE:118-121: true division
    const double tr = __dadd_rn(__ddiv_rn(__dmul_rn(Y, W.fy), Z), W.oy);
    // static_cast<int>: truncation toward zero; NaN / out-of-range (undefined in the reference) -> skipped
    if (!(tc > -1. && tc < (double)cols && tr > -1. && tr < (double)rows)) continue;
    const int tj = (int)tc, ti = (int)tr;
    const unsigned long long key = ((unsigned long long)(i + 1) << 8) | gray[(size_t)r * gray_step + c];
    atomicMax(keys + (size_t)ti * cols + tj, key);                                     // raster-order last writer wins
  }
}

__global__ void __launch_bounds__(256) k_warp_resolve(const unsigned long long* __restrict__ keys, int rows, int cols,
                                                      uint8_t* __restrict__ warped, size_t warped_step,
                                                      const uint8_t* __restrict__ target, size_t target_step,
                                                      uint8_t* __restrict__ diff, size_t diff_step) {
  const size_t n = (size_t)rows * cols;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const int r = (int)(i / cols), c = (int)(i - (size_t)r * cols);
    const unsigned long long k = keys[i];
    const int v = k ? (int)(k & 0xffull) : 0;                                          // BASE:98 zeros()
    warped[(size_t)r * warped_step + c] = (uint8_t)v;
    if (diff) { const int t = target[(size_t)r * target_step + c]; diff[(size_t)r * diff_step + c] = (uint8_t)(t > v ? t - v : v - t); }
  }
}
}  // namespace

int launch_warp_image(cudaStream_t stream, const uint8_t* gray, size_t gray_step, const void* depth, int depth_type,
                      size_t depth_step, double depth_scale, int rows, int cols, const double rt[16],
                      double fx, double fy, double ox, double oy, unsigned long long* keys,
                      uint8_t* warped, size_t warped_step, const uint8_t* target, size_t target_step,
                      uint8_t* diff, size_t diff_step) {
  WarpImgParams W;
  for (int k = 0; k < 12; ++k) W.R[k] = rt[k];
  W.fx = fx; W.fy = fy; W.ox = ox; W.oy = oy;
  W.inv_fx = 1.f / fx; W.inv_fy = 1.f / fy;          // BASE:91-92
  const size_t n = (size_t)rows * cols;
  cudaMemsetAsync(keys, 0, n * sizeof(unsigned long long), stream);
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  switch (depth_type) {
    case SRC_F32: k_warp_splat<float><<<blocks, 256, 0, stream>>>(gray, gray_step, (const float*)depth, depth_step, depth_scale, rows, cols, W, keys); break;
    case SRC_U16: k_warp_splat<uint16_t><<<blocks, 256, 0, stream>>>(gray, gray_step, (const uint16_t*)depth, depth_step, depth_scale, rows, cols, W, keys); break;
    default:      k_warp_splat<double><<<blocks, 256, 0, stream>>>(gray, gray_step, (const double*)depth, depth_step, depth_scale, rows, cols, W, keys); break;
  }
  k_warp_resolve<<<blocks, 256, 0, stream>>>(keys, rows, cols, warped, warped_step, target, target_step, diff, diff_step);
  return 2;
}

}  // namespace phovo
