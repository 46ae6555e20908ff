// phovo_batch.h -- parameter block and launchers of the batched (one CTA per pair) path.
#ifndef PHOVO_BATCH_H_
#define PHOVO_BATCH_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include "phovo_internal.h"

namespace phovo {

constexpr int kBatchThreads = 480;          // 15 warps, 128 registers per thread; 480 is a multiple of the level widths 160 / 80 / 40 / 20 (640x480) and 240 / 120 / 60 (8K, 960-wide)
constexpr int kBatchThreadsSmall = 160;      // small levels: 3 CTAs of 5 warps per SM (160 is a multiple of 80 / 40 / 20)
constexpr int kBatchSmallLevelPixels = 6400; // largest level that runs 3 CTAs per SM
constexpr int kBatchMaxLevelPixels = 22528; // 10 B/px of shared memory + tables + scratch must fit 227 KB (checked exactly
                                            // by the host); also <= 64 * kBatchThreads (validity mask) and < 65535

// The level kernel prefetches D0 / D32 / I0 up to 8 * kBatchThreads pixels past the end of a level without
// bounds guards (the values are masked); the allocation behind the last record must cover that.
constexpr size_t kBatchStoreSlackBytes = 8 * (size_t)kBatchThreads * sizeof(double) + 1024;

// Everything the two batch kernels need, passed by value (__grid_constant__).
struct BatchParams {
  int num_pairs, rows, cols;
  int num_active;                       // active levels, coarse -> fine
  int mode;
  int log_cap;                          // per-pair stats slots (0: do not record)
  int exact_always;                     // test hook: every pixel takes the exact warp (no estimate shortcut)
  int force_generic;                    // test hook: generic (r, c) bookkeeping even when BT % cols == 0
  int src_period, src_keep_begin, src_keep;   // row compaction of staged inputs (src_period 0: dense frames)
  int level[PHOVO_MAX_LEVELS];          // pyramid level index of active level a
  int lrows[PHOVO_MAX_LEVELS], lcols[PHOVO_MAX_LEVELS];
  int max_iters[PHOVO_MAX_LEVELS];
  int px_offset[PHOVO_MAX_LEVELS + 1];  // prefix sum of level pixel counts (pyramid kernel indexing)
  unsigned long long off_I0[PHOVO_MAX_LEVELS], off_I1[PHOVO_MAX_LEVELS], off_D0[PHOVO_MAX_LEVELS];
  unsigned long long off_D32[PHOVO_MAX_LEVELS];   // fp32 copy of D0, 0 where the depth fails the range test (phase A's estimate)
  unsigned long long record_bytes;      // bytes of one pair's packed level record in HBM
  double fx[PHOVO_MAX_LEVELS], fy[PHOVO_MAX_LEVELS], ox[PHOVO_MAX_LEVELS], oy[PHOVO_MAX_LEVELS];
  double inv_fx[PHOVO_MAX_LEVELS], inv_fy[PHOVO_MAX_LEVELS];
  double lambda[PHOVO_MAX_LEVELS], min_grad[PHOVO_MAX_LEVELS];
  double grad_k[PHOVO_MAX_LEVELS];      // imageGradientsScalingFactor / 1020
  double min_depth, max_depth;
};

// K1b: all active pyramid levels of I0, I1 (as exact u16 tap sums) and D0 (fp64) for every pair,
// one pass over the full-resolution inputs.  depth_type: SRC_F64 / SRC_F32 / SRC_U16.
int launch_batch_pyramid(cudaStream_t stream, const BatchParams& bp, const uint8_t* gray0, const void* depth0,
                         int depth_type, double depth_scale, const uint8_t* gray1, uint8_t* store);
// K3-batch: persistent CTAs, one pair at a time per CTA, the level's images resident in shared
// memory, the whole Gauss-Newton loop of the level on chip; one launch per active level.
// Returns the number of launches.  `next_pair`: PHOVO_MAX_LEVELS zeroed device counters.
int launch_batch_align(cudaStream_t stream, const BatchParams& bp, int sm_count, const uint8_t* store,
                       const double* init_states, double* states, int32_t* iters, phovo_iter_stats* log,
                       int32_t* log_counts, unsigned int* next_pair);
size_t batch_level_smem_bytes(int rows, int cols);
bool batch_level_is_small(int rows, int cols);
cudaError_t batch_align_prepare();

}  // namespace phovo
#endif
