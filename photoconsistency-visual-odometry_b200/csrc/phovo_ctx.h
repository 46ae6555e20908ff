// phovo_ctx.h -- the context object behind the C ABI (one CUDA device + one stream).
#ifndef PHOVO_CTX_H_
#define PHOVO_CTX_H_

#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "phovo_internal.h"
#include "phovo_kernels.h"

struct phovo_batch_state;  // phovo_batch.cu

struct phovo_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err, graph_error;

  phovo_config cfg;
  double K[9] = {0};
  bool have_K = false;
  bool build_all_levels = false;
  bool have_src = false, have_tgt = false;

  // geometry of the current frames
  int rows = 0, cols = 0;
  int lrows[PHOVO_MAX_LEVELS] = {0}, lcols[PHOVO_MAX_LEVELS] = {0};

  // HBM layout (general path): per active level five dense row-major fp64 images -- bit-identical to the
  // reference's cv::Mat_<double> levels (fp32 storage makes ~15% of trajectories diverge)
  double* I0[PHOVO_MAX_LEVELS] = {nullptr};
  double* D0[PHOVO_MAX_LEVELS] = {nullptr};
  double* I1[PHOVO_MAX_LEVELS] = {nullptr};
  double* Gx[PHOVO_MAX_LEVELS] = {nullptr};
  double* Gy[PHOVO_MAX_LEVELS] = {nullptr};
  size_t lcap[PHOVO_MAX_LEVELS][5] = {{0}};
  // photometric + depth solver: target depth pyramid, Scharr of depth / max_depth, gain per level
  double* D1[PHOVO_MAX_LEVELS] = {nullptr};
  double* GxD[PHOVO_MAX_LEVELS] = {nullptr};
  double* GyD[PHOVO_MAX_LEVELS] = {nullptr};
  size_t bcap[PHOVO_MAX_LEVELS][3] = {{0}};
  double* d_gain = nullptr;
  bool have_tgt_depth = false;
  int* winner = nullptr; size_t winner_cap = 0;           // one int per pixel of the largest active level
  unsigned char* valid = nullptr; size_t valid_cap = 0;   // one flag per pixel (K3a -> K3b)
  double* scratch64[2] = {nullptr, nullptr}; size_t scratch_cap[2] = {0, 0};
  double* partials = nullptr; size_t partials_cap = 0;    // [grid][32] per-block normal-equation partials
  char* stage_gray[2] = {nullptr, nullptr}; size_t stage_gray_cap[2] = {0, 0};
  char* stage_depth = nullptr; size_t stage_depth_cap = 0;
  double* dump_res = nullptr; size_t dump_res_cap = 0;
  double* dump_jac = nullptr; size_t dump_jac_cap = 0;
  // phovo_warp_image scratch
  unsigned long long* warp_keys = nullptr; size_t warp_keys_cap = 0;
  char* warp_io[3] = {nullptr, nullptr, nullptr}; size_t warp_io_cap[3] = {0, 0, 0};   // device copies of warped / target / diff for host callers

  // solver state
  double state[6] = {0};
  PoseDev* d_pose = nullptr; PoseDev* h_pose = nullptr;
  phovo_iter_stats* d_log = nullptr; phovo_iter_stats* h_log = nullptr; int log_cap = 0;
  phovo_iter_stats* d_eval = nullptr; phovo_iter_stats* h_eval = nullptr;
  double* d_state_in = nullptr; double* h_state_in = nullptr;
  std::vector<phovo_iter_stats> log;

  // CUDA graph of the whole Optimize()
  bool use_graph = true, graph_broken = false;
  // 2 (default): persistent cooperative kernel per level; 1: CUDA graph with WHILE nodes; 0: stream launches + host polling
  int execution = 2; bool coop_broken = false; bool cluster_broken = false; int sm_count = 0; int last_path = 0;
  cudaGraph_t graph = nullptr; cudaGraphExec_t graph_exec = nullptr;
  int graph_launches_fixed = 0, graph_launches_per_iter = 0;
  int last_used_graph = 0;

  // row sharding
  int shard_rank = 0, shard_world = 1, shard_level = -1;
  double* d_shard = nullptr;
  double* d_level_dmin = nullptr;          // [PHOVO_MAX_LEVELS] smallest valid depth per level of the current source frame
  bool level_dmin_valid = false;
  // fused peer-store exchange: own area (cudaMalloc, IPC-exported), peers' areas (IPC-opened)
  phovo::ShardExchange* xchg_own = nullptr;
  phovo::ShardExchange* xchg_peer[8] = {nullptr};
  bool xchg_opened[8] = {false};
  phovo::ShardExchange** xchg_peers_dev = nullptr;   // device copy of xchg_peer[]
  bool xchg_table_dirty = true;
  unsigned long long xchg_epoch = 0;

  // bookkeeping
  cudaEvent_t ev_copy = nullptr; cudaEvent_t ev_time[4] = {nullptr, nullptr, nullptr, nullptr};
  bool h2d_pending = false, copy_event_armed = false, setup_timed = false;
  bool defer_device_input_drain = false; // batch slots: the owner of the context keeps device inputs alive and waits itself
  bool device_input_in_flight = false;   // a Set*Frame call was given a device pointer and its kernels are queued
  phovo::LaunchState launch_state;       // per-DEVICE function attributes / occupancy of the persistent kernels
  int64_t launches = 0;

  phovo_batch_state* batch = nullptr;

  // Batch slots (phovo_batch.cu, wave path): a context used for its device side only.  Its level images, winner map,
  // scratch and PoseDev are carved out of ONE allocation owned by the batch state (several hundred slots otherwise
  // cost tens of cudaMalloc / cudaMallocHost calls each: seconds); no stream of its own, no pinned host buffers.
  bool is_slot = false;
  char* arena_base = nullptr; size_t arena_bytes = 0, arena_used = 0;
  void arena_reset(char* base, size_t bytes);   // forget every arena-backed buffer; base == nullptr: leave the arena

  int fail(int code, const std::string& what);
  int cuda_fail(const char* what, cudaError_t e);
  bool level_active(int l) const;
  LevelParams level_params(int level) const;
  phovo::LevelPtrs level_ptrs(int level) const;
  void invalidate_graph();
};

void phovo_batch_release(phovo_ctx* ctx);
// a slot context on `device` (see phovo_ctx::is_slot): set the stream and the arena before the first frame
int phovo_internal_create_slot(int device, phovo_ctx** out);
// device bytes a slot needs for rows x cols frames under its current config (upper bound, incl. alignment)
size_t phovo_internal_slot_bytes(const phovo_ctx* ctx, int rows, int cols);

#endif
