#!/usr/bin/env python
"""Extreme image sizes (a few pixels per level, one-row / one-column strips, single level) through the batch
kernels and the general path against the CPU oracle (GPU box).  Prints one line per case."""
import importlib, sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
phovo = importlib.import_module("photoconsistency-visual-odometry_b200"); phovo.build()
import oracle_py; oracle_py.build()
bad = 0
for (rows, cols, levels, P) in [(8, 8, 3, 3), (9, 13, 2, 2), (16, 24, 4, 5), (33, 47, 3, 4), (5, 200, 2, 3), (200, 5, 2, 3), (64, 80, 1, 2), (120, 160, 1, 1), (4, 4, 2, 1)]:
    f = 0.9 * cols
    K = np.array([[f, 0, (cols - 1) / 2.], [0, f, (rows - 1) / 2.], [0, 0, 1.]])
    g0, d0, g1, _ = phovo.synth.make_batch(P, rows, cols, K=K, seed0=900 + rows)
    cfg = phovo.default_config(); cfg.num_levels = levels
    for l in range(levels):
        cfg.max_num_iterations[l] = 6; cfg.min_gradient_norm[l] = 1.0
    odo = phovo.CPhotoconsistencyOdometryCuda(); odo.SetConfig(cfg); odo.SetIntrinsicMatrix(K)
    try:
        st, it = odo.BatchAlign(g0, d0, g1)
    except phovo.PhovoError as e:
        print(rows, cols, levels, "batch refused:", e); continue
    ocfg = oracle_py.Config.from_buffer_copy(bytes(cfg))
    ost, oit, _, _ = oracle_py.align_batch(ocfg, K, g0, d0, g1, num_threads=2, lean=True)
    same_it = np.array_equal(it, oit)
    with np.errstate(invalid="ignore"):
        diff = np.nanmax(np.abs(st - ost)) if np.isfinite(ost).any() else 0.0
    nan_same = np.array_equal(np.isnan(st), np.isnan(ost))
    # general path too
    gdiff = 0.0
    for p in range(P):
        odo.SetSourceFrame(g0[p], d0[p]); odo.SetTargetFrame(g1[p]); odo.SetInitialStateVector(np.zeros(6))
        try:
            odo.Optimize()
        except phovo.PhovoError as e:     # the general path reports a non-finite state as an error (the reference returns NaNs)
            if np.isfinite(ost[p]).all(): gdiff = float("inf")
            continue
        s = odo.GetOptimalStateVector()
        with np.errstate(invalid="ignore"):
            if np.isfinite(ost[p]).all(): gdiff = max(gdiff, float(np.max(np.abs(s - ost[p]))))
            else: gdiff = float("inf")        # oracle went non-finite but the device did not
    ok = same_it and nan_same and (diff < 1e-7) and gdiff < 1e-7
    bad += not ok
    if not same_it:
        print("   device iterations", it.tolist(), "oracle", oit.tolist(), "device states", st.tolist()[:2], "oracle", ost.tolist()[:2])
    print(rows, cols, levels, P, "iters equal", same_it, "nan pattern equal", nan_same, "max diff batch", diff, "general", gdiff, "OK" if ok else "MISMATCH")
print("bad", bad)
