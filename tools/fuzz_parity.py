#!/usr/bin/env python
"""Randomised parity sweep (GPU box): many small random pairs -- random sizes (odd and even, widths that do and
do not divide the CTA widths), intrinsics, depth ranges, motions (small and large), configs (every level active
or not, lambda != 1) -- through the batch kernels, the general path (cooperative and graph drivers) and the CPU
oracle.  Counts pairs whose iteration counts differ or whose pose differs by more than the north-star bar, and
reports the worst deviations.  Prints one JSON line."""
import argparse, importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np


def wave_sweep(args):
    phovo = importlib.import_module("photoconsistency-visual-odometry_b200")
    phovo.build()
    import oracle_py
    oracle_py.build()
    rng = np.random.default_rng(args.seed)
    try:
        import ref_py
        if not ref_py.available(): ref_py = None
    except Exception:
        ref_py = None
    import tempfile
    tmp = tempfile.mkdtemp()
    stats = dict(reference_checked=0, iter_mismatch_vs_reference=0, pose_over_bar_vs_reference=0, worst_vs_reference=0.,
                 pairs=0, groups=0, by_solver={}, iter_mismatch_vs_pool=0, worst_vs_pool=0., oracle_checked=0, iter_mismatch_vs_oracle=0,
                 pose_over_bar_vs_oracle=0, worst_trans_vs_oracle=0., worst_rot_vs_oracle=0., nonfinite_mismatch=0)
    odo = phovo.CPhotoconsistencyOdometryCuda()
    t0 = time.time()
    for gi in range(args.groups):
        rows = int(rng.integers(40, 150)); cols = int(rng.choice([80, 160, 96, 120, 64, int(rng.integers(50, 200))]))
        f = float(rng.uniform(0.7, 1.4)) * cols
        K = np.array([[f, 0, (cols - 1) / 2 + rng.uniform(-3, 3)], [0, f * rng.uniform(0.97, 1.03), (rows - 1) / 2 + rng.uniform(-3, 3)], [0, 0, 1.]])
        levels = int(rng.integers(2, 4))
        solver = str(rng.choice(["blur", "ceres", "bi"]))
        if args.solver: solver = args.solver
        if solver == "ceres":
            cfg = phovo.configs.to_config("config_5_level_optimization_ceres", phovo.capi)
        else:
            cfg = phovo.default_config()
            cfg.mode = 3 if solver == "bi" else int(rng.integers(0, 2))
        cfg.num_levels = levels
        for l in range(levels):
            cfg.max_num_iterations[l] = int(rng.integers(0, 12)) if l > 0 or rng.random() < 0.5 else 0
            if solver != "ceres":
                cfg.min_gradient_norm[l] = float(rng.choice([1., 30., 300.]))
                cfg.lambda_step[l] = float(rng.choice([1., 1., 0.8]))
            cfg.blur_filter_size[l] = int(rng.choice([0, 3, 5])) if solver == "blur" else 0
        if sum(cfg.max_num_iterations[l] for l in range(levels)) == 0:
            cfg.max_num_iterations[levels - 1] = 5
        if solver == "blur" and not any(cfg.blur_filter_size[l] > 1 and cfg.max_num_iterations[l] > 0 for l in range(levels)):
            l = max(l for l in range(levels) if cfg.max_num_iterations[l] > 0); cfg.blur_filter_size[l] = 3
        cfg.min_depth, cfg.max_depth = float(rng.uniform(0.2, 1.0)), float(rng.uniform(2.1, 6.5))
        if ref_py is not None: cfg.min_depth, cfg.max_depth = 0.3, 5.0  # the reference classes' own range (checked against them below)
        P = args.pairs
        scale = float(rng.choice([1., 1., 3.]))
        g0 = np.empty((P + 1, rows, cols), np.uint8); g1 = np.empty_like(g0); d0 = np.empty((P + 1, rows, cols))
        for p in range(P + 1):
            xi = phovo.synth.random_motion(int(rng.integers(0, 1 << 30))) * scale
            g0[p], d0[p], g1[p], _ = phovo.synth.make_pair(rows, cols, K, xi, int(rng.integers(0, 1 << 30)))
        kw = {"depth1": d0[1:P + 1].copy()} if solver == "bi" else {}
        g0, d0, g1 = g0[:P], d0[:P], g1[:P]
        init = np.zeros((P, 6)); init[:, :3] = rng.uniform(-2e-3, 2e-3, (P, 3))
        odo.SetConfig(cfg); odo.SetIntrinsicMatrix(K)
        if os.environ.get("PHOVO_FUZZ_EXEC"): odo.SetExecution(int(os.environ["PHOVO_FUZZ_EXEC"]))
        try:
            odo.BatchSetDebugFlags(4 if os.environ.get("PHOVO_FUZZ_POOL_ONLY") else 8)
            st, it = odo.BatchAlign(g0, d0, g1, initial_states=init, **kw)
            assert odo.BatchLastPath() == (2 if os.environ.get("PHOVO_FUZZ_POOL_ONLY") else 3), odo.BatchLastPath()
            odo.BatchSetDebugFlags(4)
            pst, pit = odo.BatchAlign(g0, d0, g1, initial_states=init, **kw)
        except Exception:
            print("FAILED in group", gi, solver, rows, cols, levels, [cfg.max_num_iterations[l] for l in range(levels)],
                  [cfg.blur_filter_size[l] for l in range(levels)], "after", odo.BatchLastPath(), file=sys.stderr)
            raise
        assert odo.BatchLastPath() == 2
        odo.BatchSetDebugFlags(0)
        stats["groups"] += 1; stats["pairs"] += P; stats["by_solver"][solver] = stats["by_solver"].get(solver, 0) + P
        for p in range(P):
            fin, pfin = np.isfinite(st[p]).all(), np.isfinite(pst[p]).all()
            if fin != pfin:
                stats["nonfinite_mismatch"] += 1
                continue
            if not np.array_equal(it[p], pit[p]):
                stats["iter_mismatch_vs_pool"] += 1
            elif fin:
                stats["worst_vs_pool"] = max(stats["worst_vs_pool"], float(np.max(np.abs(st[p] - pst[p]))))
        if solver == "bi" and ref_py is not None:            # the REFERENCE'S OWN solver (oracle/_ref: its header, compiled unmodified)
            C = phovo.configs
            yml = os.path.join(tmp, "fuzz_bi.yml")
            with open(yml, "w") as f:
                f.write(C.to_yaml({C.K_LEVELS: levels, C.K_BLUR: [0] * levels, C.K_GRAD: [float(cfg.grad_scale[l]) for l in range(levels)],
                                   C.K_LAMBDA: [float(cfg.lambda_step[l]) for l in range(levels)],
                                   C.K_ITERS: [int(cfg.max_num_iterations[l]) for l in range(levels)],
                                   C.K_MINGRAD: [float(cfg.min_gradient_norm[l]) for l in range(levels)], "visualizeIterations": 0}))
            ref = ref_py.ReferenceBiObjective(yml, K)
            for p in range(0, P, 4):
                rs, riters = ref.align(g0[p], d0[p], g1[p], kw["depth1"][p], state0=init[p])
                stats["reference_checked"] += 1
                if not (np.isfinite(rs).all() and np.isfinite(st[p]).all()):
                    stats["nonfinite_mismatch"] += int(np.isfinite(rs).all() != np.isfinite(st[p]).all())
                    continue
                if len(riters) != int(it[p].sum()):
                    stats["iter_mismatch_vs_reference"] += 1
                    continue
                stats["worst_vs_reference"] = max(stats["worst_vs_reference"], float(np.max(np.abs(st[p] - rs))))
                if np.max(np.abs(st[p, :3] - rs[:3])) >= 1e-4 or np.max(np.abs(st[p, 3:] - rs[3:])) >= 1e-5:
                    stats["pose_over_bar_vs_reference"] += 1
        if solver != "bi" and ref_py is not None and (solver == "ceres" or cfg.mode == 0):
            # analytic (bug-compatible mode = the reference's Jacobian) and Ceres-mode: the reference's own headers again
            C = phovo.configs
            vals = {C.K_LEVELS: levels, C.K_BLUR: [int(cfg.blur_filter_size[l]) for l in range(levels)],
                    C.K_GRAD: [float(cfg.grad_scale[l]) for l in range(levels)],
                    C.K_ITERS: [int(cfg.max_num_iterations[l]) for l in range(levels)], "visualizeIterations": 0}
            if solver == "ceres":
                for key, arr in ((C.K_FTOL, cfg.function_tolerance), (C.K_GTOL, cfg.gradient_tolerance), (C.K_PTOL, cfg.parameter_tolerance),
                                 (C.K_R0, cfg.initial_trust_region_radius), (C.K_RMAX, cfg.max_trust_region_radius),
                                 (C.K_RMIN, cfg.min_trust_region_radius), (C.K_ETA, cfg.min_relative_decrease)):
                    vals[key] = [float(arr[l]) for l in range(levels)]
                vals.update({"num_threads": 1, "num_linear_solver_threads": 1, "minimizer_progress_to_stdout": 0})
            else:
                vals[C.K_LAMBDA] = [float(cfg.lambda_step[l]) for l in range(levels)]
                vals[C.K_MINGRAD] = [float(cfg.min_gradient_norm[l]) for l in range(levels)]
            yml = os.path.join(tmp, "fuzz_%s.yml" % solver)
            with open(yml, "w") as f:
                f.write(C.to_yaml(vals))
            ref = ref_py.ReferenceCeres(yml, K) if solver == "ceres" else ref_py.Reference(yml, K)
            for p in range(1, P, 4):
                if solver == "ceres":
                    ref.set_frames(g0[p], d0[p], g1[p])
                    rs, rlog = ref.optimize(init[p]); rn = len(rlog)
                else:
                    rs, _, riters = ref.align(g0[p], d0[p], g1[p], state0=init[p]); rn = len(riters)
                stats["reference_checked"] += 1
                if not (np.isfinite(rs).all() and np.isfinite(st[p]).all()):
                    stats["nonfinite_mismatch"] += int(np.isfinite(rs).all() != np.isfinite(st[p]).all())
                    continue
                if rn != int(it[p].sum()):
                    stats["iter_mismatch_vs_reference"] += 1
                    continue
                stats["worst_vs_reference"] = max(stats["worst_vs_reference"], float(np.max(np.abs(st[p] - rs))))
                if np.max(np.abs(st[p, :3] - rs[:3])) >= 1e-4 or np.max(np.abs(st[p, 3:] - rs[3:])) >= 1e-5:
                    stats["pose_over_bar_vs_reference"] += 1
        if solver != "bi":                                   # the oracle has the analytic and the Ceres-mode solver
            ocfg = oracle_py.Config.from_buffer_copy(bytes(cfg))
            for p in range(0, P, 3):
                o = oracle_py.Oracle(ocfg, K)
                o.set_source(g0[p], d0[p]); o.set_target(g1[p]); o.set_initial_state(init[p]); o.optimize()
                os_ = o.state()
                stats["oracle_checked"] += 1
                if not (np.isfinite(os_).all() and np.isfinite(st[p]).all()):
                    stats["nonfinite_mismatch"] += int(np.isfinite(os_).all() != np.isfinite(st[p]).all())
                    continue
                if len(o.iter_stats()) != int(it[p].sum()):
                    stats["iter_mismatch_vs_oracle"] += 1
                    continue
                dt, dr = float(np.max(np.abs(st[p, :3] - os_[:3]))), float(np.max(np.abs(st[p, 3:] - os_[3:])))
                stats["worst_trans_vs_oracle"] = max(stats["worst_trans_vs_oracle"], dt); stats["worst_rot_vs_oracle"] = max(stats["worst_rot_vs_oracle"], dr)
                if dt >= 1e-4 or dr >= 1e-5:
                    stats["pose_over_bar_vs_oracle"] += 1
    stats["seconds"] = time.time() - t0
    print(json.dumps(stats))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--groups", type=int, default=40)
    ap.add_argument("--pairs", type=int, default=24)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--solver", default="", help="--wave: only this kind of group (blur, ceres, bi)")
    ap.add_argument("--wave", action="store_true", help="the wave path of the batch entry (Ceres-mode, photometric + depth solver, blurred "
                    "analytic levels) against the pool of per-pair contexts and, where the oracle has the solver, the CPU oracle")
    args = ap.parse_args()
    if args.wave:
        return wave_sweep(args)
    phovo = importlib.import_module("photoconsistency-visual-odometry_b200")
    phovo.build()
    import oracle_py
    oracle_py.build()
    rng = np.random.default_rng(args.seed)
    try:
        import ref_py
        if not ref_py.available(): ref_py = None
    except Exception:
        ref_py = None
    import tempfile
    tmp = tempfile.mkdtemp()
    stats = dict(pairs=0, groups=0, iter_mismatch=0, pose_over_bar=0, worst_trans=0., worst_rot=0., general_checked=0,
                 worst_general_vs_batch=0., nonfinite=0, reference_checked=0, iter_mismatch_vs_reference=0, pose_over_bar_vs_reference=0,
                 worst_vs_reference=0., sizes=[])
    odo = phovo.CPhotoconsistencyOdometryCuda()
    t0 = time.time()
    for gi in range(args.groups):
        rows = int(rng.integers(40, 150)); cols = int(rng.choice([80, 160, 96, 120, 64, int(rng.integers(50, 200))]))
        f = float(rng.uniform(0.7, 1.4)) * cols
        K = np.array([[f, 0, (cols - 1) / 2 + rng.uniform(-3, 3)], [0, f * rng.uniform(0.97, 1.03), (rows - 1) / 2 + rng.uniform(-3, 3)], [0, 0, 1.]])
        levels = int(rng.integers(2, 4))
        cfg = phovo.default_config()
        cfg.num_levels = levels
        for l in range(levels):
            cfg.max_num_iterations[l] = int(rng.integers(0, 12)) if l > 0 or rng.random() < 0.5 else 0
            cfg.min_gradient_norm[l] = float(rng.choice([1., 30., 300.]))
            cfg.lambda_step[l] = float(rng.choice([1., 1., 0.8]))
        if sum(cfg.max_num_iterations[l] for l in range(levels)) == 0:
            cfg.max_num_iterations[levels - 1] = 5
        cfg.min_depth, cfg.max_depth = float(rng.uniform(0.2, 1.0)), float(rng.uniform(2.1, 6.5))
        cfg.mode = int(rng.integers(0, 2))
        check_ref = ref_py is not None and cfg.mode == 0      # mode 0 IS the reference's Jacobian: its own header is the judge
        if check_ref: cfg.min_depth, cfg.max_depth = 0.3, 5.0 # (the reference class's range)
        P = args.pairs
        scale = float(rng.choice([1., 1., 3.]))            # some groups with large motions
        g0 = np.empty((P, rows, cols), np.uint8); g1 = np.empty_like(g0); d0 = np.empty((P, rows, cols))
        for p in range(P):
            xi = phovo.synth.random_motion(int(rng.integers(0, 1 << 30))) * scale
            g0[p], d0[p], g1[p], _ = phovo.synth.make_pair(rows, cols, K, xi, int(rng.integers(0, 1 << 30)))
        odo.SetConfig(cfg); odo.SetIntrinsicMatrix(K)
        try:
            st, it = odo.BatchAlign(g0, d0, g1)
        except phovo.PhovoError as e:
            if e.code == phovo.capi.E_UNSUPPORTED:
                continue
            raise
        ocfg = oracle_py.Config.from_buffer_copy(bytes(cfg))
        ost, oit, _, _ = oracle_py.align_batch(ocfg, K, g0, d0, g1, num_threads=os.cpu_count(), lean=True)
        stats["groups"] += 1; stats["pairs"] += P; stats["sizes"].append([rows, cols, levels])
        for p in range(P):
            if not (np.isfinite(st[p]).all() and np.isfinite(ost[p]).all()):
                stats["nonfinite"] += int(np.isfinite(st[p]).all() != np.isfinite(ost[p]).all())
                continue
            if not np.array_equal(it[p], oit[p]):
                stats["iter_mismatch"] += 1
            dt, dr = float(np.max(np.abs(st[p, :3] - ost[p, :3]))), float(np.max(np.abs(st[p, 3:] - ost[p, 3:])))
            stats["worst_trans"] = max(stats["worst_trans"], dt); stats["worst_rot"] = max(stats["worst_rot"], dr)
            if dt >= 1e-4 or dr >= 1e-5:
                stats["pose_over_bar"] += 1
        if check_ref:                                        # oracle/_ref: CPhotoconsistencyOdometryAnalytic.h compiled unmodified
            C = phovo.configs
            yml = os.path.join(tmp, "fuzz_analytic.yml")
            with open(yml, "w") as f:
                f.write(C.to_yaml({C.K_LEVELS: levels, C.K_BLUR: [0] * levels, C.K_GRAD: [float(cfg.grad_scale[l]) for l in range(levels)],
                                   C.K_LAMBDA: [float(cfg.lambda_step[l]) for l in range(levels)],
                                   C.K_ITERS: [int(cfg.max_num_iterations[l]) for l in range(levels)],
                                   C.K_MINGRAD: [float(cfg.min_gradient_norm[l]) for l in range(levels)], "visualizeIterations": 0}))
            ref = ref_py.Reference(yml, K)
            for p in range(0, P, 6):
                rs, _, riters = ref.align(g0[p], d0[p], g1[p])
                stats["reference_checked"] += 1
                if not (np.isfinite(rs).all() and np.isfinite(st[p]).all()):
                    continue
                if len(riters) != int(it[p].sum()):
                    stats["iter_mismatch_vs_reference"] += 1
                    continue
                stats["worst_vs_reference"] = max(stats["worst_vs_reference"], float(np.max(np.abs(st[p] - rs))))
                if np.max(np.abs(st[p, :3] - rs[:3])) >= 1e-4 or np.max(np.abs(st[p, 3:] - rs[3:])) >= 1e-5:
                    stats["pose_over_bar_vs_reference"] += 1
        for path in (2, 1):                                  # the general path on two pairs of the group
            odo.SetExecution(path)
            for p in (0, P - 1):
                odo.SetSourceFrame(g0[p], d0[p]); odo.SetTargetFrame(g1[p]); odo.SetInitialStateVector(np.zeros(6))
                try:
                    odo.Optimize()
                except phovo.PhovoError as e:
                    if e.code == phovo.capi.E_NUMERIC:
                        continue
                    raise
                if np.isfinite(st[p]).all():
                    stats["worst_general_vs_batch"] = max(stats["worst_general_vs_batch"], float(np.max(np.abs(odo.GetOptimalStateVector() - st[p]))))
                    stats["general_checked"] += 1
        odo.SetExecution(2)
    stats["seconds"] = time.time() - t0
    stats["sizes"] = stats["sizes"][:8]
    print(json.dumps(stats))


if __name__ == "__main__":
    main()
