"""Batched persistent kernel (one CTA per pair, level images resident in shared memory) against
the oracle and against the per-pair general path.  Needs a GPU."""
import numpy as np
import pytest

from helpers import assert_logs_match, assert_pose_close
from test_gpu_parity import conv_cfg, make_odo, run_gpu

pytestmark = pytest.mark.gpu


def torch_cuda(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def oracle_batch(oracle, cfg, K, g0, d0, g1, threads=8):
    st, it, _, _ = oracle.align_batch(conv_cfg(oracle, cfg), K, g0, d0.astype(np.float64), g1, num_threads=threads, lean=True)
    return st, it


@pytest.mark.parametrize("cfg_name,K_name,mode", [
    ("config_4_level_optimization_analytic", "K_FRAME_ALIGNMENT", 0),      # BASELINE config 4 (small batch)
    ("config_5_level_optimization_analytic", "K_VISUAL_ODOMETRY", 0),
    ("config_4_level_optimization_analytic", "K_FRAME_ALIGNMENT", 1),
])
def test_batch_matches_oracle_640x480(phovo, oracle, cfg_name, K_name, mode):
    K = getattr(phovo.synth, K_name)
    P = 6
    g0, d0, g1, xis = phovo.synth.make_batch(P, 480, 640, K=K, seed0=100)
    cfg = phovo.configs.to_config(cfg_name, phovo.capi, mode=mode)
    odo = make_odo(phovo, cfg, K)
    odo.BatchSetRecordStats(True)
    st, it = odo.BatchAlign(g0, d0.astype(np.float32), g1)
    ost, oit = oracle_batch(oracle, cfg, K, g0, d0, g1)
    assert np.array_equal(it, oit), (it, oit)           # executed iterations per level, every pair
    for p in range(P):
        assert_pose_close(st[p], ost[p], "pair %d" % p)
        o = oracle.Oracle(conv_cfg(oracle, cfg), K)
        o.set_source(g0[p], d0[p])
        o.set_target(g1[p])
        o.set_initial_state(np.zeros(6))
        o.optimize()
        assert_logs_match(odo.BatchIterationStats(p), o.iter_stats(), what="pair %d" % p)
    # the per-pair general path gives the same answers (different summation order only)
    for p in (0, P - 1):
        s, log = run_gpu(odo, g0[p], d0[p], g1[p])
        assert np.max(np.abs(s - st[p])) < 1e-10
        assert len(log) == int(it[p].sum())
    # f64 depth input and device-resident inputs give bitwise the same result
    st64, it64 = odo.BatchAlign(g0, d0, g1)
    assert np.array_equal(st64, st) and np.array_equal(it64, it)
    # so does the kernel variant without the iteration log (it does not accumulate sum r^2)
    odo.BatchSetRecordStats(False)
    stq, itq = odo.BatchAlign(g0, d0.astype(np.float32), g1)
    assert np.array_equal(stq, st) and np.array_equal(itq, it)
    import torch
    tg0, td0, tg1 = (torch.from_numpy(a).cuda() for a in (g0, d0.astype(np.float32), g1))
    stdev, itdev = odo.BatchAlign(tg0, td0, tg1)
    assert np.array_equal(stdev, st) and np.array_equal(itdev, it)
    out_s = torch.zeros((P, 6), dtype=torch.float64, device="cuda")
    out_i = torch.zeros((P, phovo.MAXL), dtype=torch.int32, device="cuda")
    odo.BatchAlignDevice(tg0, td0, tg1, out_s, out_i)
    odo.Synchronize()
    assert np.array_equal(out_s.cpu().numpy(), st) and np.array_equal(out_i.cpu().numpy(), it)


def test_batch_is_order_and_grid_independent(phovo):
    """Pair p's result does not depend on the batch it travels in (fixed thread->pixel mapping):
    this is what makes 1-GPU and N-GPU sharded runs bitwise identical."""
    K = phovo.synth.K_FRAME_ALIGNMENT
    P = 5
    g0, d0, g1, _ = phovo.synth.make_batch(P, 240, 320, K=K, seed0=200)
    d0 = d0.astype(np.float32)
    cfg = phovo.configs.to_config("config_4_level_optimization_analytic", phovo.capi)
    cfg.num_levels = 3
    for l, m in enumerate((0, 20, 50)):
        cfg.max_num_iterations[l] = m
    odo = make_odo(phovo, cfg, K)
    st, it = odo.BatchAlign(g0, d0, g1)
    perm = np.array([3, 0, 4, 1, 2])
    st2, it2 = odo.BatchAlign(g0[perm], d0[perm], g1[perm])
    assert np.array_equal(st2, st[perm]) and np.array_equal(it2, it[perm])
    st3, it3 = odo.BatchAlign(g0[2:3], d0[2:3], g1[2:3])
    assert np.array_equal(st3[0], st[2])
    # initial states are honoured per pair
    init = np.tile(np.array([1e-3, -1e-3, 2e-3, 1e-3, 0, -1e-3]), (P, 1))
    st4, _ = odo.BatchAlign(g0, d0, g1, initial_states=init)
    assert not np.array_equal(st4, st)


def test_batch_odd_size_and_u16_depth(phovo, oracle):
    K = np.array([[120., 0, 70.], [0, 118., 46.], [0, 0, 1]])
    P = 3
    g0, d0, g1, _ = phovo.synth.make_batch(P, 93, 141, K=K, seed0=300)
    cfg = phovo.default_config()
    cfg.num_levels = 3
    for l, (m, thr) in enumerate(((0, 300.), (8, 100.), (12, 100.))):
        cfg.max_num_iterations[l] = m
        cfg.min_gradient_norm[l] = thr
    odo = make_odo(phovo, cfg, K)
    raw = np.clip(np.rint(d0 * 5000.), 0, 65535).astype(np.uint16)
    st, it = odo.BatchAlign(g0, raw, g1, depth_scale=1. / 5000.)
    ost, oit = oracle_batch(oracle, cfg, K, g0, raw.astype(np.float64) * (1. / 5000.), g1, threads=3)
    assert np.array_equal(it, oit)
    for p in range(P):
        assert_pose_close(st[p], ost[p])


@pytest.mark.parametrize("shape,cfg_name", [((480, 640), "config_4_level_optimization_analytic"),
                                            ((93, 141), "test_3_level_all_active"),
                                            ((120, 160), "test_3_level_all_active")])
def test_batch_shortcuts_do_not_change_results(phovo, shape, cfg_name):
    """The estimate-then-verify warp and the column-fixed bookkeeping are pure optimisations: with
    either switched off (exact reference warp for EVERY pixel / generic bookkeeping) states,
    iteration counts and every logged normal equation are bitwise the same."""
    K = phovo.synth.K_FRAME_ALIGNMENT.copy()
    K[:2] *= shape[1] / 640.
    P = 4
    g0, d0, g1, _ = phovo.synth.make_batch(P, shape[0], shape[1], K=K, seed0=400)
    cfg = phovo.configs.to_config(cfg_name, phovo.capi)
    if shape == (120, 160):
        cfg.max_num_iterations[0] = 0        # 120x160 = 19 200 px fits; keep level 0 out to stay small
    odo = make_odo(phovo, cfg, K)
    odo.BatchSetRecordStats(True)
    runs = []
    for flags in (0, 1, 2, 3):
        odo.BatchSetDebugFlags(flags)
        st, it = odo.BatchAlign(g0, d0.astype(np.float32), g1)
        if odo.BatchLastPath() != 1:
            pytest.skip("a level does not fit the shared-memory-resident kernels: the pool path ran (no shortcuts to switch off)")
        logs = [odo.BatchIterationStats(p) for p in range(P)]
        runs.append((st, it, logs))
    odo.BatchSetDebugFlags(0)
    st0, it0, logs0 = runs[0]
    assert it0.sum() > 0
    for st, it, logs in runs[1:]:
        assert np.array_equal(st, st0) and np.array_equal(it, it0)
        for la, lb in zip(logs, logs0):
            assert len(la) == len(lb)
            for a, b in zip(la, lb):
                assert a["num_valid"] == b["num_valid"]
                assert np.array_equal(a["H"], b["H"]) and np.array_equal(a["g"], b["g"])


def test_batch_pixels_near_the_camera_plane(phovo, oracle):
    """Initial states that put the scene ON the camera plane of the target frame (Z' = d + z crosses
    zero inside the image): the estimated warp loses its accuracy to cancellation there, the kernel
    has to notice (|Z'| < zmin) and take the exact path.  First-iteration winner counts and normal
    equations must equal the oracle's for every pair, whatever happens to the trajectory later."""
    K = np.array([[130., 0, 79.5], [0, 130., 59.5], [0, 0, 1]])
    P = 4
    g0, d0, g1, _ = phovo.synth.make_batch(P, 120, 160, K=K, seed0=700)
    cfg = phovo.default_config()
    cfg.num_levels = 2
    cfg.max_num_iterations[0] = 3
    cfg.max_num_iterations[1] = 3
    cfg.min_gradient_norm[0] = cfg.min_gradient_norm[1] = 1e-3
    cfg.min_depth, cfg.max_depth = 0.05, 20.
    odo = make_odo(phovo, cfg, K)
    odo.BatchSetRecordStats(True)
    valid = d0[d0 > 0]
    zs = [-float(np.median(valid)), -float(np.percentile(valid, 30)) + 1e-9, -float(valid.min()) - 1e-7, -float(np.percentile(valid, 70))]
    init = np.zeros((P, 6))
    for p in range(P):
        init[p] = [1e-3, -2e-3, zs[p], 2e-3, -1e-3, 1e-3]
    st, it = odo.BatchAlign(g0, d0, g1, initial_states=init)
    for p in range(P):
        o = oracle.Oracle(conv_cfg(oracle, cfg), K)
        o.set_source(g0[p], d0[p])
        o.set_target(g1[p])
        o.set_initial_state(init[p])
        o.optimize()
        glog, olog = odo.BatchIterationStats(p), o.iter_stats()
        assert len(glog) >= 1 and len(olog) >= 1
        a, b = glog[0], olog[0]
        assert a["num_valid"] == b["num_valid"], (p, a["num_valid"], b["num_valid"])
        scale = max(np.max(np.abs(b["H"])), 1e-300)
        assert np.max(np.abs(a["H"] - b["H"])) <= 1e-9 * scale, p
        assert np.max(np.abs(a["g"] - b["g"])) <= 1e-9 * max(np.max(np.abs(b["g"])), 1e-300), p


@pytest.mark.parametrize("rows,cols,levels", [(8, 8, 3), (5, 200, 2), (16, 24, 4), (9, 13, 2)])
def test_batch_degenerate_levels_behave_like_the_reference(phovo, oracle, rows, cols, levels):
    """Levels of a few pixels: the normal equations go singular, the state turns NaN, every pixel is
    invalid from then on and the reference's zero Jacobian rows give a zero gradient, which ends each
    later level after one iteration (AN:388).  Iteration counts and the NaN pattern must be the
    reference's; where the alignment stays finite (9x13) the poses must agree as usual."""
    f = 0.9 * cols
    K = np.array([[f, 0, (cols - 1) / 2.], [0, f, (rows - 1) / 2.], [0, 0, 1.]])
    P = 3
    g0, d0, g1, _ = phovo.synth.make_batch(P, rows, cols, K=K, seed0=900 + rows)
    cfg = phovo.default_config()
    cfg.num_levels = levels
    for l in range(levels):
        cfg.max_num_iterations[l] = 6
        cfg.min_gradient_norm[l] = 1.0
    odo = make_odo(phovo, cfg, K)
    st, it = odo.BatchAlign(g0, d0, g1)
    ost, oit = oracle_batch(oracle, cfg, K, g0, d0, g1, threads=2)
    assert np.array_equal(it, oit), (it.tolist(), oit.tolist())
    assert np.array_equal(np.isnan(st), np.isnan(ost))
    for p in range(P):
        if np.isfinite(ost[p]).all():
            assert_pose_close(st[p], ost[p])


def test_batch_unusual_depth_values_and_ranges(phovo, oracle):
    """NaN / inf / negative / zero / huge depths, depth ranges (0, 1e300), (-2, 5), (0.3, inf) and large
    initial states: iteration counts, per-iteration valid-pixel counts and poses must be the reference's.
    (An unbounded depth range switches the estimated warp off: every pixel takes the exact path.)"""
    rows, cols, P = 96, 128, 4
    K = np.array([[110., 0, 63.5], [0, 110., 47.5], [0, 0, 1.]])
    g0, d0, g1, _ = phovo.synth.make_batch(P, rows, cols, K=K, seed0=1234)
    rng = np.random.default_rng(5)
    d = d0.copy()
    m = rng.random(d.shape)
    d[m < 0.02] = np.nan
    d[(m > 0.02) & (m < 0.03)] = np.inf
    d[(m > 0.03) & (m < 0.04)] = -1.5
    d[(m > 0.04) & (m < 0.05)] = 0.
    d[(m > 0.05) & (m < 0.06)] = 1e30
    big = np.zeros((P, 6))
    big[:, 0] = [0.5, -2., 10., 0.]
    big[:, 3] = [0.3, 1.5, -3.0, 0.]
    big[:, 5] = [0., 0.7, 0.1, 3.1]
    for dd, (mn, mx), init in ((d, (0.3, 5.0), None), (d, (0.0, 1e300), None), (d, (-2.0, 5.0), None),
                               (d0, (0.3, float("inf")), None), (d0, (0.3, 5.0), big)):
        cfg = phovo.default_config()
        cfg.num_levels = 3
        for l in range(3):
            cfg.max_num_iterations[l] = (0, 5, 8)[l]
            cfg.min_gradient_norm[l] = 1e-2
        cfg.min_depth, cfg.max_depth = mn, mx
        odo = make_odo(phovo, cfg, K)
        odo.BatchSetRecordStats(True)
        st, _ = odo.BatchAlign(g0, dd, g1, initial_states=init)
        for p in range(P):
            o = oracle.Oracle(conv_cfg(oracle, cfg), K)
            o.set_source(g0[p], dd[p])
            o.set_target(g1[p])
            o.set_initial_state(np.zeros(6) if init is None else init[p])
            o.optimize()
            olog, glog = o.iter_stats(), odo.BatchIterationStats(p)
            assert len(olog) == len(glog), (mn, mx, p)
            assert [a["num_valid"] for a in glog] == [b["num_valid"] for b in olog], (mn, mx, p)
            if np.isfinite(o.state()).all():
                assert_pose_close(st[p], o.state(), "range (%g, %g) pair %d" % (mn, mx, p))
            else:
                assert np.array_equal(np.isnan(st[p]), np.isnan(o.state()))


def test_batch_configurations_beyond_the_resident_kernels_take_the_wave_path(phovo):
    """phovo_batch_align never refuses a configuration: a level that does not fit in shared memory, blurred levels and the
    Ceres-mode solver run in waves of per-pair slots -- pyramids by the general path's kernels, then one CTA per pair
    through every level (k_align_slots).  Same iteration counts as the per-pair API and its states to the last bits
    (one CTA sums what a grid summed); the pool of per-pair contexts (debug flag 4) gives them bitwise."""
    K = phovo.synth.K_FRAME_ALIGNMENT
    P = 6
    g0, d0, g1, _ = phovo.synth.make_batch(P, 480, 640, K=K, seed0=1)
    d0 = d0.astype(np.float32)
    cases = []
    cfg = phovo.configs.to_config("config_4_level_optimization_analytic", phovo.capi)
    cfg.max_num_iterations[1] = 3          # level 1 = 240x320 = 76 800 px does not fit in shared memory
    cases.append(cfg)
    cfg = phovo.configs.to_config("config_4_level_optimization_analytic", phovo.capi)
    cfg.blur_filter_size[3] = 3
    cases.append(cfg)
    cases.append(phovo.configs.to_config("config_5_level_optimization_ceres", phovo.capi))
    odo = make_odo(phovo, cases[0], K)
    init = np.zeros((P, 6)); init[:, 0] = 1e-3 * np.arange(P)
    for cfg in cases:
        odo.SetConfig(cfg)
        single = make_odo(phovo, cfg, K)
        ref_states, ref_iters = [], []
        for p in range(P):
            single.SetSourceFrame(g0[p], d0[p]); single.SetTargetFrame(g1[p]); single.SetInitialStateVector(init[p])
            single.Optimize()
            ref_states.append(single.GetOptimalStateVector()); ref_iters.append(len(single.IterationStats()))
        st, it = odo.BatchAlign(g0, d0, g1, initial_states=init)
        assert odo.BatchLastPath() == 2                       # a small batch: the pool by default
        for flags, path in ((8, 3), (4, 2)):
            odo.BatchSetDebugFlags(flags)
            for inputs in ((g0, d0, g1), tuple(torch_cuda(a) for a in (g0, d0, g1))):   # host and device-resident batches
                st, it = odo.BatchAlign(*inputs, initial_states=init)
                assert odo.BatchLastPath() == path
                for p in range(P):
                    if path == 2: assert np.array_equal(st[p], ref_states[p]), (cfg.mode, p)
                    else: assert np.max(np.abs(st[p] - ref_states[p])) < 1e-10, (cfg.mode, p, st[p] - ref_states[p])
                    assert int(it[p].sum()) == ref_iters[p] > 0, (cfg.mode, path, p)
        odo.BatchSetDebugFlags(0)
        # per-pair per-iteration stats: the per-pair API's own log, entry by entry
        odo.BatchSetRecordStats(True)
        st, it = odo.BatchAlign(g0, d0, g1, initial_states=init)
        assert odo.BatchLastPath() == 2
        single.SetSourceFrame(g0[P - 1], d0[P - 1]); single.SetTargetFrame(g1[P - 1]); single.SetInitialStateVector(init[P - 1])
        single.Optimize()
        blog, slog = odo.BatchIterationStats(P - 1), single.IterationStats()
        assert len(blog) == len(slog) == ref_iters[P - 1]
        for a, b in zip(blog, slog):
            assert a["level"] == b["level"] and a["num_valid"] == b["num_valid"] and a["accepted"] == b["accepted"]
            assert np.array_equal(a["H"], b["H"]) and np.array_equal(a["g"], b["g"]) and np.array_equal(a["state_out"], b["state_out"])
        odo.BatchSetRecordStats(False)
    # back on the resident kernels
    odo.SetConfig(phovo.configs.to_config("config_4_level_optimization_analytic", phovo.capi))
    odo.BatchAlign(g0, d0, g1)
    assert odo.BatchLastPath() == 1
    # all levels inactive: states pass through
    cfg = phovo.configs.to_config("config_4_level_optimization_analytic", phovo.capi)
    for l in range(4):
        cfg.max_num_iterations[l] = 0
    odo.SetConfig(cfg)
    st, it = odo.BatchAlign(g0, d0, g1, initial_states=np.full((P, 6), 0.01))
    assert np.array_equal(st, np.full((P, 6), 0.01)) and not it.any()
    # the same through the wave path (Ceres mode) and the pool
    cfg = phovo.configs.to_config("config_5_level_optimization_ceres", phovo.capi)
    for l in range(5):
        cfg.max_num_iterations[l] = 0
    odo.SetConfig(cfg)
    for flags, path in ((8, 3), (4, 2)):
        odo.BatchSetDebugFlags(flags)
        st, it = odo.BatchAlign(g0, d0, g1, initial_states=np.full((P, 6), 0.02))
        assert odo.BatchLastPath() == path
        assert np.array_equal(st, np.full((P, 6), 0.02)) and not it.any()
    odo.BatchSetDebugFlags(0)


def test_batch_at_scale_matches_general_path(phovo):
    """Size-independent property at a BASELINE-sized workload: every pair of a 296-pair 640x480
    batch (two full waves of persistent CTAs, dynamic scheduling, both kernel variants) gets the
    pose and iteration counts the per-pair general path computes for it, and a second run is
    bitwise identical."""
    import torch
    K = phovo.synth.K_FRAME_ALIGNMENT
    P = 296
    g0, d0, g1, _ = phovo.synth.render_batch_torch(P, 480, 640, K, torch.device("cuda", 0), seed0=900)
    cfg = phovo.configs.to_config("config_4_level_optimization_analytic", phovo.capi)
    odo = make_odo(phovo, cfg, K)
    st, it = odo.BatchAlign(g0, d0, g1)
    st2, it2 = odo.BatchAlign(g0, d0, g1)
    assert np.array_equal(st, st2) and np.array_equal(it, it2)
    assert np.isfinite(st).all() and it[:, 2].min() >= 1 and it[:, 3].min() >= 1 and it[:, 3].max() <= 50
    worst = 0.
    for p in range(0, P, 7):
        odo.SetSourceFrame(g0[p], d0[p])
        odo.SetTargetFrame(g1[p])
        odo.SetInitialStateVector(np.zeros(6))
        odo.Optimize()
        log = odo.IterationStats()
        assert [sum(1 for e in log if e["level"] == l) for l in (2, 3)] == [int(it[p, 2]), int(it[p, 3])], p
        worst = max(worst, float(np.max(np.abs(odo.GetOptimalStateVector() - st[p]))))
    assert worst < 1e-9, worst


def test_wave_path_over_several_waves_matches_the_pool_and_the_oracle(phovo, oracle):
    """More pairs than slots: the two halves of slots alternate (a half is harvested before it is set up again), host
    and device inputs, caller's initial states, both other solvers and a blurred analytic configuration.  Every pair
    gets the iteration counts of the per-pair API (the pool, bitwise that API) and its state to 1e-10; a sample is
    checked against the CPU oracle as well."""
    import torch
    K = phovo.synth.K_FRAME_ALIGNMENT.copy(); K[:2] *= 80 / 640.
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    P = 4 * sms + 37                                   # two full waves + a partial third: half 0 is reused
    g0, d0, g1, _ = phovo.synth.make_batch(P + 1, 60, 80, K=K, seed0=5)
    d1 = d0[1:P + 1].copy(); g0, d0, g1 = g0[:P], d0[:P], g1[:P]
    init = np.zeros((P, 6)); init[:, 1] = 1e-4 * (np.arange(P) % 11)
    cases = []
    cfg = phovo.configs.to_config("test_3_level_all_active", phovo.capi); cfg.blur_filter_size[1] = 3
    cases.append(("blurred analytic", cfg, {}))
    cases.append(("photometric + depth", phovo.configs.to_config("test_3_level_all_active", phovo.capi, mode=phovo.MODE_BIOBJECTIVE), {"depth1": d1}))
    cfg = phovo.configs.to_config("config_5_level_optimization_ceres", phovo.capi); cfg.num_levels = 3
    cases.append(("ceres-mode", cfg, {}))
    for name, cfg, kw in cases:
        odo = make_odo(phovo, cfg, K)
        odo.BatchSetDebugFlags(4)
        st_pool, it_pool = odo.BatchAlign(g0, d0, g1, initial_states=init, **kw)
        assert odo.BatchLastPath() == 2
        odo.BatchSetDebugFlags(0)
        dev_kw = {k: torch_cuda(v) for k, v in kw.items()}
        for inputs, kws in (((g0, d0, g1), kw), (tuple(torch_cuda(a) for a in (g0, d0, g1)), dev_kw)):
            st, it = odo.BatchAlign(*inputs, initial_states=init, **kws)
            assert odo.BatchLastPath() == 3
            assert np.array_equal(it, it_pool), (name, np.nonzero((it != it_pool).any(axis=1))[0][:10])
            fin = np.isfinite(st_pool).all(axis=1)
            assert np.array_equal(np.isfinite(st).all(axis=1), fin), name
            assert np.max(np.abs(st[fin] - st_pool[fin])) < 1e-10, (name, float(np.max(np.abs(st[fin] - st_pool[fin]))))
        if name == "blurred analytic":                 # the oracle on a sample (analytic modes run in its batch entry)
            sel = np.arange(0, P, 97)
            for p in sel:
                o = oracle.Oracle(conv_cfg(oracle, cfg), K)
                o.set_source(g0[p], d0[p].astype(np.float64)); o.set_target(g1[p]); o.set_initial_state(init[p]); o.optimize()
                assert len(o.iter_stats()) == int(it[p].sum()), (name, p)
                if np.isfinite(o.state()).all():
                    assert_pose_close(st[p], o.state(), "%s pair %d" % (name, p))
        odo.close()


def test_wave_path_with_a_small_memory_budget_runs_more_smaller_waves(phovo, monkeypatch):
    """The slot arena takes at most half of the free device memory: with little of it the waves get smaller, the results
    do not change (PHOVO_WAVE_BUDGET_MB is the test hook for 'little')."""
    K = phovo.synth.K_FRAME_ALIGNMENT.copy(); K[:2] *= 160 / 640.
    P = 23
    g0, d0, g1, _ = phovo.synth.make_batch(P, 120, 160, K=K, seed0=3)
    cfg = phovo.configs.to_config("config_5_level_optimization_ceres", phovo.capi); cfg.num_levels = 3
    odo = make_odo(phovo, cfg, K)
    odo.BatchSetDebugFlags(8)                              # waves whatever the batch size
    st_ref, it_ref = odo.BatchAlign(g0, d0, g1)
    assert odo.BatchLastPath() == 3
    monkeypatch.setenv("PHOVO_WAVE_BUDGET_MB", "8")       # a slot of this configuration is ~1.3 MB: 3 slots per half, 8 waves
    st, it = odo.BatchAlign(g0, d0, g1)
    assert odo.BatchLastPath() == 3
    assert np.array_equal(it, it_ref) and np.array_equal(st, st_ref)   # same kernel, same order of sums: bitwise
    monkeypatch.delenv("PHOVO_WAVE_BUDGET_MB")
    st, it = odo.BatchAlign(g0, d0, g1)
    assert np.array_equal(it, it_ref) and np.array_equal(st, st_ref)
    # CTAs per pair: a small wave gives every pair a thread-block cluster (here 8 by default); every size gives the same
    # iterations and the same states up to the grouping of the partial sums
    for c in (1, 2, 4, 8):
        monkeypatch.setenv("PHOVO_WAVE_CLUSTER", str(c))
        st, it = odo.BatchAlign(g0, d0, g1)
        assert odo.BatchLastPath() == 3
        assert np.array_equal(it, it_ref), c
        assert np.max(np.abs(st - st_ref)) < 1e-11, (c, float(np.max(np.abs(st - st_ref))))
        if c == 8: assert np.array_equal(st, st_ref)
    monkeypatch.delenv("PHOVO_WAVE_CLUSTER")
    # the arena, the slots and the pool can be given back; the next call builds them again (flags back to defaults)
    import torch
    free0 = torch.cuda.mem_get_info()[0]
    odo.BatchReleaseMemory()
    assert torch.cuda.mem_get_info()[0] > free0
    odo.BatchSetDebugFlags(8)
    st, it = odo.BatchAlign(g0, d0, g1)
    assert odo.BatchLastPath() == 3 and np.array_equal(it, it_ref) and np.array_equal(st, st_ref)
    odo.close()


def test_growing_frame_sizes_keep_the_winner_map_initialised(phovo, oracle):
    """Regression: after a frame-size change the winner map is reallocated; cudaFree + cudaMalloc can return the SAME address
    with a larger extent, whose tail holds garbage indices unless the map is refilled (it used to be refilled only when
    the pointer changed -> sporadic illegal addresses in phase B, found by the randomised sweep on the pool of 8 contexts).
    Alternating small and large frames through the pool and the per-pair API, checked against the CPU oracle."""
    rng = np.random.default_rng(5)
    odo = phovo.CPhotoconsistencyOdometryCuda()
    odo.BatchSetDebugFlags(4)
    for rep in range(10):
        for rows, cols in ((40 + 4 * rep, 56), (128, 64 + 8 * rep)):
            f = 1.1 * cols
            K = np.array([[f, 0, (cols - 1) / 2], [0, f, (rows - 1) / 2], [0, 0, 1.]])
            cfg = phovo.default_config(); cfg.num_levels = 3
            for l, (it, bl) in enumerate(zip((9, 0, 8), (5, 5, 0))):
                cfg.max_num_iterations[l] = it; cfg.blur_filter_size[l] = bl
            P = 12
            g0, d0, g1, _ = phovo.synth.make_batch(P, rows, cols, K=K, seed0=int(rng.integers(0, 1 << 20)))
            odo.SetConfig(cfg); odo.SetIntrinsicMatrix(K)
            st, it = odo.BatchAlign(g0, d0, g1)
            assert odo.BatchLastPath() == 2
            for p in (0, P - 1):
                o = oracle.Oracle(conv_cfg(oracle, cfg), K)
                o.set_source(g0[p], d0[p].astype(np.float64)); o.set_target(g1[p]); o.set_initial_state(np.zeros(6)); o.optimize()
                assert len(o.iter_stats()) == int(it[p].sum())
                if np.isfinite(o.state()).all():
                    assert_pose_close(st[p], o.state(), "rep %d %dx%d pair %d" % (rep, rows, cols, p))
    odo.close()


def test_randomised_parity_sweep(phovo, oracle):
    """A small slice of tools/fuzz_parity.py as a regression test: random sizes / intrinsics / configs."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_parity.py"), "--groups", "30", "--pairs", "8", "--seed", "3"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    st = json.loads(out.stdout.strip().splitlines()[-1])
    assert st["pairs"] >= 100 and st["iter_mismatch"] == 0 and st["pose_over_bar"] == 0 and st["nonfinite"] == 0, st
    assert st["worst_trans"] < 1e-9 and st["worst_rot"] < 1e-9 and st["worst_general_vs_batch"] < 1e-9, st
    # where the prebuilt oracle/_ref travelled: the bug-compatible groups are also judged by the reference's own header
    assert st["iter_mismatch_vs_reference"] == 0 and st["pose_over_bar_vs_reference"] == 0 and st["worst_vs_reference"] < 1e-9, st


def test_randomised_parity_sweep_of_the_wave_path(phovo, oracle):
    """The same for what the resident kernels do not take (Ceres-mode, photometric + depth solver, blurred levels): slot
    waves against the pool of per-pair contexts, the CPU oracle and -- when oracle/_ref is present -- the reference's own
    three solver headers."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_parity.py"), "--wave", "--groups", "45", "--pairs", "12", "--seed", "3"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    st = json.loads(out.stdout.strip().splitlines()[-1])
    assert st["pairs"] >= 500 and len(st["by_solver"]) == 3, st
    assert st["iter_mismatch_vs_pool"] == 0 and st["worst_vs_pool"] < 1e-10 and st["nonfinite_mismatch"] == 0, st
    assert st["oracle_checked"] > 50 and st["iter_mismatch_vs_oracle"] == 0 and st["pose_over_bar_vs_oracle"] == 0, st
    assert st["iter_mismatch_vs_reference"] == 0 and st["pose_over_bar_vs_reference"] == 0 and st["worst_vs_reference"] < 1e-9, st


def test_sequence_as_one_batch_matches_the_vo_loop(phovo, oracle):
    """AlignSequence: the VO app's per-frame alignments (zero initial state every frame) computed as one
    batch over overlapping views of the uploaded sequence == the sequential loop, frame by frame."""
    K = phovo.synth.K_VISUAL_ODOMETRY
    n = 12
    frames = [phovo.synth.make_sequence_frame(k, 480, 640, K=K) for k in range(n)]
    gray = np.stack([f[0] for f in frames])
    depth = np.stack([f[1] for f in frames]).astype(np.float32)
    cfg = phovo.configs.to_config("config_5_level_optimization_analytic", phovo.capi)
    odo = make_odo(phovo, cfg, K)
    st, it, poses = odo.AlignSequence(gray, depth)
    assert st.shape == (n - 1, 6) and poses.shape == (n, 4, 4)
    seq = make_odo(phovo, cfg, K)
    pose = np.eye(4)
    seq.SetSourceFrame(*frames[0])
    for k in range(1, n):
        if k > 1:
            seq.PromoteTargetToSource(frames[k - 1][1])
        seq.SetTargetFrame(frames[k][0])
        seq.SetInitialStateVector(np.zeros(6))
        seq.Optimize()
        assert np.max(np.abs(seq.GetOptimalStateVector() - st[k - 1])) < 1e-9, k
        assert len(seq.IterationStats()) == int(it[k - 1].sum())
        pose = pose @ np.linalg.inv(seq.GetOptimalRigidTransformationMatrix())
        assert np.max(np.abs(pose - poses[k])) < 1e-8
    # against the oracle for the first frames
    for k in (1, 2):
        o = oracle.Oracle(conv_cfg(oracle, cfg), K)
        o.set_source(*frames[k - 1]); o.set_target(frames[k][0]); o.set_initial_state(np.zeros(6)); o.optimize()
        assert_pose_close(st[k - 1], o.state(), "frame %d" % k)
