// kernels_batch.cu -- the batched path (BASELINE config "N independent 640x480 pairs"):
//
//   K1b k_batch_pyramid : one streaming pass over the full-resolution inputs of every pair that
//        writes, for each ACTIVE pyramid level only, I0 and I1 as EXACT 10-bit tap sums (u16,
//        value = sum/1020) and D0 as fp64 (the reference's own double average).  HBM-bound.
//        (AN:115-163 with blur 0.)
//   K3-batch k_batch_align : persistent CTAs; a CTA takes one pair at a time, keeps the level's
//        I0/I1 tap sums, the 16-bit winner map and the 16-bit target map resident in shared
//        memory (8 B/px), streams D0 (fp64) from L2, recomputes the Scharr gradients exactly from
//        the resident I1 sums (integer arithmetic), and runs the complete coarse-to-fine
//        Gauss-Newton loop (AN:500-563) without leaving the SM: no grid-wide synchronisation, no
//        atomics on floating-point data.  The thread->pixel mapping is fixed, so results are
//        bitwise reproducible and independent of which SM, CTA or GPU processes the pair.
//
// Why exact storage: the residual scatter (AN:358) makes the cost piecewise constant in the
// warp; when a level does not converge (it often runs to max_num_iterations) any 1e-7
// perturbation (e.g. fp32 images) flips a rounding somewhere along the 50-iteration trajectory
// and the final pose moves by >1e-4 in ~15% of pairs.  Integer tap sums + fp64 depth keep the
// device within ~1e-15 of the reference's doubles, so trajectories stay together.
#include "phovo_batch.h"
#include "phovo_device.cuh"
#include "phovo_kernels.h"

namespace phovo {
namespace {

constexpr int BT = kBatchThreads;

__device__ __forceinline__ void linear_axis(int d, double scale, int ssize, bool is_x, int& s0, float& w1) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (s < 0) { f = 0.f; s = 0; }
  if (is_x && s >= ssize - 1) { f = 0.f; s = ssize - 1; }
  s0 = s; w1 = f;
}

template <typename DT> __device__ __forceinline__ double depth_at(const DT* p, double scale);
template <> __device__ __forceinline__ double depth_at<double>(const double* p, double) { return __ldg(p); }
template <> __device__ __forceinline__ double depth_at<float>(const float* p, double) { return (double)__ldg(p); }
template <> __device__ __forceinline__ double depth_at<uint16_t>(const uint16_t* p, double scale) { return __dmul_rn((double)__ldg(p), scale); }

// K1b.  grid = (ceil(pixels of all active levels / 256), pairs).  One thread per level pixel:
// 4 taps from each of gray0, gray1 and depth0 (the central 2x2 of the pixel's 2^level cell).
// Consecutive threads read consecutive cells of the same two source rows, so a warp touches one
// contiguous span per row and every fetched sector is consumed by the warp.
template <typename DT>
__global__ void __launch_bounds__(256) k_batch_pyramid(const __grid_constant__ BatchParams bp,
                                                       const uint8_t* __restrict__ gray0, const DT* __restrict__ depth0,
                                                       double depth_scale, const uint8_t* __restrict__ gray1,
                                                       uint8_t* __restrict__ store) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= bp.px_offset[bp.num_active]) return;
  int a = 0;
  while (j >= bp.px_offset[a + 1]) ++a;
  const int q = j - bp.px_offset[a];
  const int lc = bp.lcols[a];
  const int y = q / lc, x = q - y * lc;
  const int level = bp.level[a];
  const double scale = (double)(1 << level);
  const size_t pair = blockIdx.y;
  const size_t frame = (size_t)bp.rows * bp.cols;
  const uint8_t* G0 = gray0 + pair * frame;
  const uint8_t* G1 = gray1 + pair * frame;
  const DT* D = depth0 + pair * frame;
  int sx, sy; float fx, fy;
  linear_axis(x, scale, bp.cols, true, sx, fx);
  linear_axis(y, scale, bp.rows, false, sy, fy);
  const int y0 = min(max(sy, 0), bp.rows - 1), y1 = min(max(sy + 1, 0), bp.rows - 1);
  const bool two = sx + 1 < bp.cols;
  const size_t o00 = (size_t)y0 * bp.cols + sx, o10 = (size_t)y1 * bp.cols + sx;
  unsigned s0, s1;
  double dv;
  if (two) {
    s0 = (unsigned)__ldg(G0 + o00) + __ldg(G0 + o00 + 1) + __ldg(G0 + o10) + __ldg(G0 + o10 + 1);
    s1 = (unsigned)__ldg(G1 + o00) + __ldg(G1 + o00 + 1) + __ldg(G1 + o10) + __ldg(G1 + o10 + 1);
    const double a0 = (double)(1.f - fx), a1 = (double)fx;
    const double r0 = __dadd_rn(__dmul_rn(depth_at<DT>(D + o00, depth_scale), a0), __dmul_rn(depth_at<DT>(D + o00 + 1, depth_scale), a1));
    const double r1 = __dadd_rn(__dmul_rn(depth_at<DT>(D + o10, depth_scale), a0), __dmul_rn(depth_at<DT>(D + o10 + 1, depth_scale), a1));
    dv = __dadd_rn(__dmul_rn(r0, (double)(1.f - fy)), __dmul_rn(r1, (double)fy));
  } else {  // right border of an odd-sized image: a single horizontal tap with weight one
    s0 = 2u * ((unsigned)__ldg(G0 + o00) + __ldg(G0 + o10));
    s1 = 2u * ((unsigned)__ldg(G1 + o00) + __ldg(G1 + o10));
    dv = __dadd_rn(__dmul_rn(depth_at<DT>(D + o00, depth_scale), (double)(1.f - fy)), __dmul_rn(depth_at<DT>(D + o10, depth_scale), (double)fy));
  }
  uint8_t* rec = store + pair * bp.record_bytes;
  ((uint16_t*)(rec + bp.off_I0[a]))[q] = (uint16_t)s0;
  ((uint16_t*)(rec + bp.off_I1[a]))[q] = (uint16_t)s1;
  ((double*)(rec + bp.off_D0[a]))[q] = dv;
}

// 16-bit "max" into shared memory: 0 = empty, otherwise source index + 1.
__device__ __forceinline__ void smem_max_u16(unsigned short* addr, unsigned short v) {
  unsigned short old = *addr;
  while (v > old) {
    const unsigned short prev = atomicCAS(addr, old, v);
    if (prev == old) break;
    old = prev;
  }
}

struct BatchShared {
  PoseDev pose;          // state + rotation / trig of the current iterate
  double totals[32];
  int done;
  int iteration;
};

constexpr unsigned short kNoTarget = 0xFFFFu;

template <int MODE>
__global__ void __launch_bounds__(BT, 1) k_batch_align(const __grid_constant__ BatchParams bp, const uint8_t* __restrict__ store,
                                                       const double* __restrict__ init_states, double* __restrict__ states,
                                                       int32_t* __restrict__ iters, phovo_iter_stats* __restrict__ log,
                                                       int32_t* __restrict__ log_counts, int nmax) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: I0 u16[nmax] | I1 u16[nmax] | win u16[nmax] | tgt u16[nmax] | reduction scratch | BatchShared
  unsigned short* sI0 = (unsigned short*)smem_raw;
  unsigned short* sI1 = sI0 + nmax;
  unsigned short* sWin = sI1 + nmax;
  unsigned short* sTgt = sWin + nmax;
  double* sRed = (double*)(sTgt + nmax);
  BatchShared* sh = (BatchShared*)(sRed + (BT / 32) * PHOVO_ACC_STRIDE);
  const int tid = threadIdx.x;

  for (int pair = blockIdx.x; pair < bp.num_pairs; pair += gridDim.x) {
    __syncthreads();   // the previous pair's outputs have been read from shared memory
    if (tid == 0) {
      double s[6];
      for (int k = 0; k < 6; ++k) s[k] = init_states ? init_states[(size_t)pair * 6 + k] : 0.;
      Pose P;
      pose_from_state(s, P);
      for (int k = 0; k < 6; ++k) sh->pose.state[k] = s[k];
      pose_store(P, &sh->pose);
      sh->pose.log_count = 0;
    }
    const uint8_t* rec = store + (size_t)pair * bp.record_bytes;
    for (int a = 0; a < bp.num_active; ++a) {
      const int rows = bp.lrows[a], cols = bp.lcols[a], n = rows * cols;
      const double* __restrict__ gD0 = (const double*)(rec + bp.off_D0[a]);
      __syncthreads();   // previous level / pair fully consumed before the buffers are overwritten
      {
        // HBM -> shared memory, 16-byte vectors (record offsets are 16-byte aligned)
        const uint4* gI0 = (const uint4*)(rec + bp.off_I0[a]);
        const uint4* gI1 = (const uint4*)(rec + bp.off_I1[a]);
        uint4* i04 = (uint4*)sI0; uint4* i14 = (uint4*)sI1; uint4* w4 = (uint4*)sWin;
        const int ni = (n * 2 + 15) / 16;
        for (int k = tid; k < ni; k += BT) { i04[k] = __ldg(gI0 + k); i14[k] = __ldg(gI1 + k); w4[k] = make_uint4(0, 0, 0, 0); }
      }
      if (tid == 0) { sh->done = 0; sh->iteration = 0; }
      __syncthreads();

      LevelParams L;
      L.fx = bp.fx[a]; L.fy = bp.fy[a]; L.ox = bp.ox[a]; L.oy = bp.oy[a]; L.inv_fx = bp.inv_fx[a]; L.inv_fy = bp.inv_fy[a];
      L.min_depth = bp.min_depth; L.max_depth = bp.max_depth; L.rows = rows; L.cols = cols;
      L.lambda = bp.lambda[a]; L.min_grad_norm = bp.min_grad[a]; L.max_iters = bp.max_iters[a]; L.level = bp.level[a];
      const double gk = bp.grad_k[a];
      // pixel i = tid + k*BT: (r, c) advance by a fixed (dr, dc) per step -- no division in the loops
      const int r_first = tid / cols, c_first = tid - r_first * cols;
      const int dr = BT / cols, dc = BT - dr * cols;

      for (int it = 0; it < L.max_iters; ++it) {
        Pose T;
        pose_load(&sh->pose, T);
        // ---- phase A1: warp every source pixel; record its target slot; plain (racy) store of
        //      the candidate winner -- colliding writers are resolved in A2 ----
        {
          int r = r_first, c = c_first;
          for (int i = tid; i < n; i += BT) {
            Warped w;
            unsigned short t = kNoTarget;
            if (warp_pixel<false>(L, T, r, c, __ldg(gD0 + i), w)) { t = (unsigned short)w.t; sWin[w.t] = (unsigned short)(i + 1); }
            sTgt[i] = t;
            c += dc; r += dr;
            if (c >= cols) { c -= cols; ++r; }
          }
        }
        __syncthreads();
        // ---- phase A2: a source that should have won its slot but lost the store race fixes it
        //      (AN:358 last-writer-wins in raster order == largest source index); ~1% of pixels ----
        for (int i = tid; i < n; i += BT) {
          const unsigned short t = sTgt[i];
          if (t != kNoTarget && sWin[t] < (unsigned short)(i + 1)) smem_max_u16(sWin + t, (unsigned short)(i + 1));
        }
        __syncthreads();
        // ---- phase B: residual + Jacobian + normal equations (AN:271-366, 538-539) ----
        double acc[PHOVO_NACC];
#pragma unroll
        for (int v = 0; v < PHOVO_NACC; ++v) acc[v] = 0.;
        {
          int r = r_first, c = c_first;
          for (int i = tid; i < n; i += BT) {
            const unsigned short win = sWin[i];
            sWin[i] = 0;
            double res = 0.;
            if (win) {
              // I = sum/1020: one rounding, within 2 ulp of the reference's convertTo + resize doubles
              res = (double)((int)sI1[i] - (int)sI0[win - 1]) * (1.0 / 1020.0);
              acc[27] = fma(res, res, acc[27]);
            }
            if (sTgt[i] != kNoTarget) {
              const double d = __ldg(gD0 + i);
              Warped w;
              w.px = __dmul_rn(__dmul_rn(__dsub_rn((double)c, L.ox), d), L.inv_fx);
              w.py = __dmul_rn(__dmul_rn(__dsub_rn((double)r, L.oy), d), L.inv_fy);
              w.q0 = __dadd_rn(__dadd_rn(__dmul_rn(T.R00, w.px), __dmul_rn(T.R01, w.py)), __dmul_rn(T.R02, d));
              w.q1 = __dadd_rn(__dadd_rn(__dmul_rn(T.R10, w.px), __dmul_rn(T.R11, w.py)), __dmul_rn(T.R12, d));
              w.q2 = __dadd_rn(__dadd_rn(__dmul_rn(T.R20, w.px), __dmul_rn(T.R21, w.py)), __dmul_rn(T.R22, d));
              w.X = __dadd_rn(w.q0, T.x); w.Y = __dadd_rn(w.q1, T.y); w.Z = __dadd_rn(w.q2, T.z);
              w.iz = __ddiv_rn(1.0, w.Z);
              // Scharr of I1 at the SOURCE index (AN:346-347), reflect-101, exact integer numerators
              const int rm = r > 0 ? r - 1 : (rows > 1 ? 1 : 0), rp = r < rows - 1 ? r + 1 : (rows > 1 ? rows - 2 : 0);
              const int cm = c > 0 ? c - 1 : (cols > 1 ? 1 : 0), cp = c < cols - 1 ? c + 1 : (cols > 1 ? cols - 2 : 0);
              const unsigned short* Rm = sI1 + rm * cols; const unsigned short* R0 = sI1 + r * cols; const unsigned short* Rp = sI1 + rp * cols;
              const int a00 = Rm[cm], a01 = Rm[c], a02 = Rm[cp], a10 = R0[cm], a12 = R0[cp], a20 = Rp[cm], a21 = Rp[c], a22 = Rp[cp];
              const int gxn = 10 * (a12 - a10) + 3 * ((a22 - a20) + (a02 - a00));
              const int gyn = (3 * a20 + 10 * a21 + 3 * a22) - (3 * a00 + 10 * a01 + 3 * a02);
              const double gx = (double)gxn * gk, gy = (double)gyn * gk;
              double Ju[6], Jv[6], J[6];
              projection_jacobian<MODE == 0>(L, T, w, d, Ju, Jv);
#pragma unroll
              for (int k = 0; k < 6; ++k) J[k] = gx * Ju[k] + gy * Jv[k];
              accumulate_row(acc, J, res);
              acc[28] += 1.;
            }
            c += dc; r += dr;
            if (c >= cols) { c -= cols; ++r; }
          }
        }
        const double total = block_reduce<BT>(acc, sRed);
        if (tid < PHOVO_NACC) sh->totals[tid] = total;
        __syncthreads();
        // ---- Gauss-Newton step + termination test (AN:538-549, 376-392), one thread ----
        if (tid == 0) {
          const double* t = sh->totals;
          double g[6], step[6], s_in[6], s_out[6], n2 = 0.;
          for (int k = 0; k < 6; ++k) { g[k] = t[21 + k]; n2 = fma(g[k], g[k], n2); s_in[k] = sh->pose.state[k]; }
          solve6_lu(t, g, step);
          for (int k = 0; k < 6; ++k) s_out[k] = s_in[k] - L.lambda * step[k];
          const double gnorm = sqrt(n2);
          const int done = (it + 1 >= L.max_iters) || (gnorm < L.min_grad_norm);
          if (log && sh->pose.log_count < bp.log_cap) {
            phovo_iter_stats* e = log + (size_t)pair * bp.log_cap + sh->pose.log_count;
            e->level = L.level; e->iteration = it; e->num_valid = (int)t[28]; e->accepted = 1;
            for (int k = 0; k < 21; ++k) e->H[k] = t[k];
            for (int k = 0; k < 6; ++k) { e->g[k] = g[k]; e->state_in[k] = s_in[k]; e->state_out[k] = s_out[k]; }
            e->grad_norm = gnorm; e->cost = 0.5 * t[27]; e->radius = 0.;
          }
          sh->pose.log_count += 1;
          Pose P;
          pose_from_state(s_out, P);
          for (int k = 0; k < 6; ++k) sh->pose.state[k] = s_out[k];
          pose_store(P, &sh->pose);
          sh->done = done;
          sh->iteration = it + 1;
        }
        __syncthreads();
        if (sh->done) break;
      }
      if (tid == 0 && iters) iters[(size_t)pair * PHOVO_MAX_LEVELS + L.level] = sh->iteration;
    }
    __syncthreads();
    if (tid < 6) states[(size_t)pair * 6 + tid] = sh->pose.state[tid];
    if (tid == 0 && log_counts) log_counts[pair] = sh->pose.log_count;
  }
}

}  // namespace

size_t batch_align_smem_bytes(int nmax) {
  nmax = (nmax + 7) & ~7;
  return (size_t)nmax * 8 + (size_t)(BT / 32) * PHOVO_ACC_STRIDE * sizeof(double) + sizeof(BatchShared) + 64;
}

cudaError_t batch_align_prepare(size_t smem_bytes) {
  cudaError_t e = cudaFuncSetAttribute(k_batch_align<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_batch_align<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
}

int launch_batch_pyramid(cudaStream_t stream, const BatchParams& bp, const uint8_t* gray0, const void* depth0,
                         int depth_type, double depth_scale, const uint8_t* gray1, uint8_t* store) {
  dim3 grid((bp.px_offset[bp.num_active] + 255) / 256, bp.num_pairs);
  switch (depth_type) {
    case SRC_F32: k_batch_pyramid<float><<<grid, 256, 0, stream>>>(bp, gray0, (const float*)depth0, depth_scale, gray1, store); break;
    case SRC_U16: k_batch_pyramid<uint16_t><<<grid, 256, 0, stream>>>(bp, gray0, (const uint16_t*)depth0, depth_scale, gray1, store); break;
    default:      k_batch_pyramid<double><<<grid, 256, 0, stream>>>(bp, gray0, (const double*)depth0, depth_scale, gray1, store); break;
  }
  return 1;
}

int launch_batch_align(cudaStream_t stream, const BatchParams& bp, int grid, size_t smem_bytes, const uint8_t* store,
                       const double* init_states, double* states, int32_t* iters, phovo_iter_stats* log, int32_t* log_counts) {
  int nmax = 0;
  for (int a = 0; a < bp.num_active; ++a) nmax = max(nmax, bp.lrows[a] * bp.lcols[a]);
  nmax = (nmax + 7) & ~7;
  if (bp.mode == PHOVO_MODE_ANALYTIC_FIXED)
    k_batch_align<1><<<grid, BT, smem_bytes, stream>>>(bp, store, init_states, states, iters, log, log_counts, nmax);
  else
    k_batch_align<0><<<grid, BT, smem_bytes, stream>>>(bp, store, init_states, states, iters, log, log_counts, nmax);
  return 1;
}

}  // namespace phovo
