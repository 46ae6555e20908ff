#!/usr/bin/env python
"""Corner cases against the CPU oracle (GPU box), one line per case:
  * extreme image sizes (a few pixels per level, one-row / one-column strips, a single level), where the
    normal equations go singular and the state turns NaN: iteration counts and NaN patterns must match;
  * unusual VALUES at an ordinary size (NaN / inf / negative / zero / huge depths, depth ranges (0, 1e300),
    (-2, 5), (0.3, inf), large initial states, a scene lying on the target camera plane) through the batch
    kernels and through the general path under the cooperative (estimated warp) and graph (exact warp) drivers:
    iteration counts, per-iteration valid-pixel counts and poses must match."""
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

# ---- unusual VALUES at an ordinary size: NaN / inf / negative / zero / huge depths, exotic depth ranges, large initial states ----
print("--- unusual values")
rows, cols, P = 96, 128, 4
K = np.array([[110., 0, 63.5], [0, 110., 47.5], [0, 0, 1.]])
g0, d0, g1, _ = phovo.synth.make_batch(P, rows, cols, K=K, seed0=1234)
rng = np.random.default_rng(5)
bad2 = 0
cases = []
d = d0.copy(); m = rng.random(d.shape); d[m < 0.02] = np.nan; d[(m > 0.02) & (m < 0.03)] = np.inf; d[(m > 0.03) & (m < 0.04)] = -1.5; d[(m > 0.04) & (m < 0.05)] = 0.; d[(m > 0.05) & (m < 0.06)] = 1e30
cases.append(("nan/inf/negative/zero/huge depths", d, (0.3, 5.0), None))
cases.append(("min_depth 0, max_depth 1e300", d, (0.0, 1e300), None))
cases.append(("negative min_depth", d, (-2.0, 5.0), None))
cases.append(("max_depth inf", d0, (0.3, float("inf")), None))
big = np.zeros((P, 6)); big[:, 0] = [0.5, -2., 10., 0.]; big[:, 3] = [0.3, 1.5, -3.0, 0.]; big[:, 5] = [0., 0.7, 0.1, 3.1]
cases.append(("large initial states", d0, (0.3, 5.0), big))
tiny = np.zeros((P, 6)); tiny[:, 2] = -d0.reshape(P, -1).mean(axis=1)
cases.append(("scene on the camera plane", d0, (0.05, 50.0), tiny))
for name, dd, (mn, mx), init in cases:
    cfg = phovo.default_config(); cfg.num_levels = 3
    for l in range(3):
        cfg.max_num_iterations[l] = (0, 5, 8)[l]; cfg.min_gradient_norm[l] = 1e-2
    cfg.min_depth, cfg.max_depth = mn, mx
    odo = phovo.CPhotoconsistencyOdometryCuda(); odo.SetConfig(cfg); odo.SetIntrinsicMatrix(K)
    odo.BatchSetRecordStats(True)
    st, it = odo.BatchAlign(g0, dd, g1, initial_states=init)
    ocfg = oracle_py.Config.from_buffer_copy(bytes(cfg))
    worst, same_it, same_valid = 0.0, True, True
    for p in range(P):
        o = oracle_py.Oracle(ocfg, K)
        o.set_source(g0[p], dd[p]); o.set_target(g1[p]); o.set_initial_state(np.zeros(6) if init is None else init[p]); o.optimize()
        olog, glog = o.iter_stats(), odo.BatchIterationStats(p)
        same_it &= len(olog) == len(glog)
        for a, b in zip(glog, olog):
            same_valid &= a["num_valid"] == b["num_valid"]
        os_ = o.state()
        if np.isfinite(os_).all() and np.isfinite(st[p]).all():
            worst = max(worst, float(np.max(np.abs(st[p] - os_))))
        else:
            same_valid &= bool(np.array_equal(np.isnan(st[p]), np.isnan(os_)))
    ok = same_it and same_valid and worst < 1e-6
    bad2 += not ok
    print(name, "| iterations equal", same_it, "| valid counts equal", same_valid, "| worst pose diff", worst, "OK" if ok else "MISMATCH")
print("bad", bad2)

# ---- the same unusual values through the general path (cooperative driver: estimated warp; graph driver: exact warp) ----
print("--- unusual values, general path")
bad3 = 0
for name, dd, (mn, mx), init in cases:
    cfg = phovo.default_config(); cfg.num_levels = 3
    for l in range(3):
        cfg.max_num_iterations[l] = (0, 5, 8)[l]; cfg.min_gradient_norm[l] = 1e-2
    cfg.min_depth, cfg.max_depth = mn, mx
    ocfg = oracle_py.Config.from_buffer_copy(bytes(cfg))
    for path in (2, 1):
        odo = phovo.CPhotoconsistencyOdometryCuda(); odo.SetConfig(cfg); odo.SetIntrinsicMatrix(K); odo.SetExecution(path)
        worst, same = 0.0, True
        for p in range(P):
            o = oracle_py.Oracle(ocfg, K)
            s0 = np.zeros(6) if init is None else init[p]
            o.set_source(g0[p], dd[p]); o.set_target(g1[p]); o.set_initial_state(s0); o.optimize()
            odo.SetSourceFrame(g0[p], dd[p]); odo.SetTargetFrame(g1[p]); odo.SetInitialStateVector(s0)
            try:
                odo.Optimize()
            except phovo.PhovoError:
                same &= not np.isfinite(o.state()).all()
                continue
            glog, olog = odo.IterationStats(), o.iter_stats()
            same &= len(glog) == len(olog) and all(a["num_valid"] == b["num_valid"] for a, b in zip(glog, olog))
            if np.isfinite(o.state()).all():
                worst = max(worst, float(np.max(np.abs(odo.GetOptimalStateVector() - o.state()))))
        ok = same and worst < 1e-6
        bad3 += not ok
        print(name, "| driver", path, "| iterations and valid counts equal", same, "| worst pose diff", worst, "OK" if ok else "MISMATCH")
print("bad", bad3)
