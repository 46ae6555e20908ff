// phovo_internal.h -- declarations shared by the host translation units of libphovo_b200.so.
#ifndef PHOVO_INTERNAL_H_
#define PHOVO_INTERNAL_H_

#include <stdint.h>
#include <string>

#include "../../include/phovo_b200.h"

int  phovo_internal_parse_yaml(const char* path, phovo_config* cfg, std::string* err);
void phovo_internal_default_config(phovo_config* cfg);

// Per-level pinhole parameters, computed on the host exactly as the reference does
// (CPhotoconsistencyOdometryAnalytic.h:203-209 for the analytic modes,
//  CPhotoconsistencyOdometryCeres.h:163-168 for Ceres mode) and passed to kernels by value.
struct LevelParams {
  double fx, fy, ox, oy, inv_fx, inv_fy;
  double min_depth, max_depth;
  double lambda, min_grad_norm;
  int rows, cols;
  int max_iters;
  int mode;       // PHOVO_MODE_*
  int level;
  int row_begin, row_end;  // source rows whose J/H/g this launch accumulates (row sharding); winner map always covers all rows
};

// Device-resident solver state: written by the solve kernel, read by the per-pixel kernels.
// Everything the iteration loop needs lives here so the loop runs without host involvement.
struct PoseDev {
  double state[6];
  double R[9];
  double sy, cy, sp, cp, sr, cr;
  int iteration;     // iterations executed on the current level
  int done;          // termination flag of the current level (AN:376-392)
  int log_count;     // entries written to the stats log so far
  int log_capacity;
  int iters_per_level[PHOVO_MAX_LEVELS];
  int pad[2];
};

#define PHOVO_NACC 29  // 21 H + 6 g + cost + count
#define PHOVO_ACC_STRIDE 32

#endif
