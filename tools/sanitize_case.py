#!/usr/bin/env python
"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): one 120x160 pair through the
general path (all three drivers) and a 6-pair batch through the per-level batch kernels
(small-level and large-level variants, column-fixed and generic bookkeeping)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
phovo = importlib.import_module("photoconsistency-visual-odometry_b200")
K = np.array([[131.25, 0., 79.5], [0., 131.25, 59.5], [0., 0., 1.]])
g0, d0, g1, _ = phovo.synth.make_pair(120, 160, K=K, seed=1)
cfg = phovo.default_config()
cfg.num_levels = 3
for l, m in enumerate((2, 3, 4)):
    cfg.max_num_iterations[l] = m
    cfg.min_gradient_norm[l] = 1.
odo = phovo.CPhotoconsistencyOdometryCuda()
odo.SetConfig(cfg); odo.SetIntrinsicMatrix(K)
for path in (2, 1, 0):
    odo.SetExecution(path)
    odo.SetSourceFrame(g0, d0); odo.SetTargetFrame(g1); odo.SetInitialStateVector(np.zeros(6)); odo.Optimize()
    print("path", path, odo.LastPath(), odo.GetOptimalStateVector())
G0, D0, G1, _ = phovo.synth.make_batch(6, 120, 160, K=K, seed0=10)
for flags in (0, 2):
    odo.BatchSetDebugFlags(flags)
    st, it = odo.BatchAlign(G0, D0.astype(np.float32), G1)
    print("batch flags", flags, it[:, :3].tolist())
G0, D0, G1, _ = phovo.synth.make_batch(2, 93, 141, K=K, seed0=20)
odo.BatchSetDebugFlags(0)
st, it = odo.BatchAlign(G0, D0.astype(np.float32), G1)
print("batch odd", it[:, :3].tolist())
