"""Pins for the CPU oracle (run without a GPU).

1. third-party image arithmetic vs the real OpenCV (python cv2): convertTo, resize, Scharr, GaussianBlur;
2. per-pixel + Gauss-Newton restatement vs the independent numpy/cv2 restatement;
3. the Jacobian vs a symbolic re-derivation of phovo/Maxima/derivatives_photoconsistency.wxm;
4. semantics pins from SURVEY section 4: collision winner order, round-half-away, strict depth bounds,
   zero-iteration levels, step-then-stop.
"""
import importlib

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
synth = importlib.import_module("photoconsistency-visual-odometry_b200.synth")


def ulps(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.spacing(np.maximum(np.abs(a), np.abs(b))), 1e-300))


@pytest.fixture
def plain_opencv():
    """OpenCV's plain C++ code path: cv2.setUseOptimized(False) switches off the IPP / SIMD dispatch of
    this particular build (opencv-python 4.13 ships IPP), whose kernels round differently in the last
    bit.  The reference asks for "OpenCV >= 2.4.5" (CMakeLists.txt:29): the portable code is the pin."""
    cv2.setUseOptimized(False)
    yield
    cv2.setUseOptimized(True)


@pytest.mark.parametrize("shape", [(480, 640), (135, 240), (135, 243), (101, 77), (33, 47)])
def test_image_ops_match_opencv_bit_for_bit(oracle, plain_opencv, shape):
    rng = np.random.default_rng(7)
    img8 = rng.integers(0, 256, shape).astype(np.uint8)
    a = oracle.convert_u8(img8)
    assert np.array_equal(a, img8.astype(np.float64) * (1. / 255))          # Mat::convertTo(CV_64F, 1/255)
    for lvl in range(1, 6):
        f = 0.5 ** lvl
        if min(shape) * f < 1:
            break
        ref = cv2.resize(a, (0, 0), fx=f, fy=f)                             # AN:132
        mine = oracle.resize_level(a, lvl)
        assert ref.shape == mine.shape == oracle.level_size(shape[0], shape[1], lvl)
        # level 1 is OpenCV's INTER_AREA fast path (sequential 4-tap sum, single precision on cells cut by
        # the border: 135 and 243 are 3 mod 4), levels >= 2 the bilinear table
        assert np.array_equal(ref, mine), lvl
    for dx, dy in ((1, 0), (0, 1)):
        for sc in (0.0625, 1.0, 0.005, 2.0):
            ref = cv2.Scharr(a, cv2.CV_64F, dx, dy, scale=sc)               # AN:181-187
            assert np.array_equal(ref, oracle.scharr(a, dx, dy, sc))
    for k in (3, 5, 7):
        # the filter arithmetic (generic row filter, symmetric column filter, reflect-101) bit for bit, on
        # the kernel of OpenCV 2.4 / 3.x getGaussianKernel: exp() per tap, sequential sum, times 1/sum ...
        t = np.array([np.exp((-0.5 / 9.) * (i - (k - 1) * 0.5) ** 2) for i in range(k)])
        total = 0.
        for v in t:
            total += v
        kern = t * (1. / total)
        assert np.array_equal(cv2.sepFilter2D(a, cv2.CV_64F, kern, kern), oracle.gaussian_blur(a, k, 3.0))
        # ... OpenCV 4.x builds the same kernel with soft-float and a folded sum (taps differ in the last
        # bit), hence not bit-identical to cv2 4.13's GaussianBlur                               AN:146
        assert np.max(np.abs(cv2.GaussianBlur(a, (k, k), 3) - oracle.gaussian_blur(a, k, 3.0))) < 1e-15


def test_image_ops_vs_this_builds_optimised_opencv(oracle):
    """With the build's IPP / AVX2 dispatch on (cv2's default) the same calls agree to the last ulps."""
    rng = np.random.default_rng(8)
    a = oracle.convert_u8(rng.integers(0, 256, (136, 240)).astype(np.uint8))   # no cells cut by the border (IPP does not go through single precision there)
    for lvl in (1, 2, 3):
        assert ulps(cv2.resize(a, (0, 0), fx=0.5 ** lvl, fy=0.5 ** lvl), oracle.resize_level(a, lvl)) <= 2
    assert np.max(np.abs(cv2.Scharr(a, cv2.CV_64F, 1, 0, scale=0.0625) - oracle.scharr(a, 1, 0, 0.0625))) <= 4e-15


def test_level_sizes_round_half_even(oracle):
    assert oracle.level_size(135, 240, 1) == (68, 120)      # cvRound(67.5) = 68
    assert oracle.level_size(45, 45, 1) == (22, 22)         # cvRound(22.5) = 22
    assert oracle.level_size(4320, 7680, 5) == (135, 240)
    assert oracle.level_size(480, 640, 3) == (60, 80)


def _setup(oracle, mode=0, levels=4, iters=(0, 0, 20, 50), rows=480, cols=640, seed=0, **kw):
    g0, d0, g1, _ = synth.make_pair(rows, cols, seed=seed)
    cfg = oracle.make_config(mode=mode, num_levels=levels, max_iters=iters, **kw)
    o = oracle.Oracle(cfg, synth.K_FRAME_ALIGNMENT)
    o.set_source(g0, d0)
    o.set_target(g1)
    return o, (g0, d0, g1)


@pytest.mark.parametrize("mode", [0, 1])
def test_oracle_matches_numpy_restatement(oracle, nr, mode):
    o, (g0, d0, g1) = _setup(oracle, mode)
    o.set_initial_state(np.zeros(6))
    o.optimize()
    I0, D0, I1, Gx, Gy = nr.build_pyramids(g0, d0, g1, 4, [0] * 4, [0.0625] * 4)
    st, log = nr.analytic_optimize(I0, D0, I1, Gx, Gy, synth.K_FRAME_ALIGNMENT, 4, [0, 0, 20, 50], [1] * 4,
                                   [300] * 4, np.zeros(6), fixed=(mode == 1))
    ol = o.iter_stats()
    assert len(ol) == len(log) == 9                      # 4 @ L3 + 5 @ L2 (SURVEY 3.5 probe)
    assert [s["level"] for s in ol] == [3] * 4 + [2] * 5
    for a, b in zip(ol, log):
        assert a["num_valid"] == b["num_valid"]
        Hn = nr.pack_upper(b["H"])
        assert np.max(np.abs(a["H"] - Hn) / np.abs(Hn)) < 1e-10
        assert np.max(np.abs(a["g"] - b["g"])) < 1e-9 * np.max(np.abs(b["g"]))
    assert np.max(np.abs(o.state() - st)) < 1e-12


def test_collision_winner_is_last_in_raster_order(oracle, nr):
    o, (g0, d0, g1) = _setup(oracle)
    st = np.array([0.01, -0.005, 0.02, 0.004, -0.003, 0.002])
    w = o.winner_map(3, st)
    I0, D0, I1, Gx, Gy = nr.build_pyramids(g0, d0, g1, 4, [0] * 4, [0.0625] * 4)
    _, _, cnt, _, wn = nr.analytic_eval(I0[3], D0[3], I1[3], Gx[3], Gy[3], synth.K_FRAME_ALIGNMENT, 3, st)
    assert np.array_equal(w, wn)
    assert cnt > (w >= 0).sum()          # the scene does produce collisions


def test_jacobian_matches_symbolic_derivation(oracle):
    """Maxima spec (derivatives_photoconsistency.wxm:5-37) re-derived with sympy: the corrected mode
    equals it everywhere, the bug-compatible mode differs only in row 0 of d/dz, d/dpitch, d/droll."""
    sp = pytest.importorskip("sympy")
    x, y, z, yaw, pitch, roll, px, py, pz, fx, fy, ox, oy = sp.symbols("x y z yaw pitch roll px py pz fx fy ox oy")
    cy, sy, cp, spi, cr, sr = sp.cos(yaw), sp.sin(yaw), sp.cos(pitch), sp.sin(pitch), sp.cos(roll), sp.sin(roll)
    Rt = sp.Matrix([[cy * cp, cy * spi * sr - sy * cr, cy * spi * cr + sy * sr, x],
                    [sy * cp, sy * spi * sr + cy * cr, sy * spi * cr - cy * sr, y],
                    [-spi, cp * sr, cp * cr, z], [0, 0, 0, 1]])
    P = Rt * sp.Matrix([px, py, pz, 1])
    uv = sp.Matrix([P[0] * fx / P[2] + ox, P[1] * fy / P[2] + oy])
    Jsym = uv.jacobian([x, y, z, yaw, pitch, roll])
    f = sp.lambdify([x, y, z, yaw, pitch, roll, px, py, pz, fx, fy, ox, oy], Jsym, "numpy")
    rows, cols = 12, 16
    rng = np.random.default_rng(3)
    K = np.array([[20., 0, 7.5], [0, 21., 5.5], [0, 0, 1]])
    g0 = rng.integers(0, 256, (rows, cols)).astype(np.uint8)
    g1 = rng.integers(0, 256, (rows, cols)).astype(np.uint8)
    d0 = rng.uniform(1.0, 3.0, (rows, cols))
    st = np.array([0.03, -0.02, 0.05, 0.02, -0.03, 0.01])
    for mode in (1, 0):
        cfg = oracle.make_config(mode=mode, num_levels=1, max_iters=(1,), grad_scale=1.0)
        o = oracle.Oracle(cfg, K)
        o.set_source(g0, d0)
        o.set_target(g1)
        e = o.eval(0, st, want_residuals=True, want_jacobian=True)
        Gx, Gy = o.level_image(3, 0), o.level_image(4, 0)
        worst_fixed = 0.0
        diff_cols = set()
        for r in range(rows):
            for c in range(cols):
                d = d0[r, c]
                p = np.array([(c - K[0, 2]) * d / K[0, 0], (r - K[1, 2]) * d / K[1, 1], d])
                J2 = np.array(f(*st, *p, K[0, 0], K[1, 1], K[0, 2], K[1, 2]), dtype=float)
                want = Gx[r, c] * J2[0] + Gy[r, c] * J2[1]
                got = e["jacobian"][r * cols + c]
                if not np.any(got):
                    continue      # projected out of bounds
                if mode == 1:
                    worst_fixed = max(worst_fixed, np.max(np.abs(got - want) / (np.abs(want) + 1e-9)))
                else:
                    for k in range(6):
                        if abs(got[k] - want[k]) > 1e-9 * (abs(want[k]) + 1e-9):
                            diff_cols.add(k)
        if mode == 1:
            assert worst_fixed < 1e-10
        else:
            assert diff_cols == {2, 4, 5}     # z, pitch, roll (AN:253 slip)


def test_semantics_round_half_away_and_strict_depth(oracle):
    # 1x4 image, identity pose except a pure x-translation chosen so tc lands exactly on .5
    rows, cols = 4, 8
    K = np.array([[8., 0, 3.5], [0, 8., 1.5], [0, 0, 1]])
    g = (np.arange(rows * cols).reshape(rows, cols) * 3 % 251).astype(np.uint8)
    d = np.full((rows, cols), 2.0)
    d[0, 0] = 0.3          # == min depth -> invalid (strict, AN:280)
    d[0, 1] = 5.0          # == max depth -> invalid
    cfg = oracle.make_config(num_levels=1, max_iters=(1,))
    o = oracle.Oracle(cfg, K)
    o.set_source(g, d)
    o.set_target(g)
    # tc = (px + tx) * fx / z + ox ; with tx = 0.125 m, z = 2, fx = 8 -> shift of +0.5 px exactly
    w = o.winner_map(0, np.array([0.125, 0, 0, 0, 0, 0]))
    w = w.reshape(rows, cols)
    # c + 0.5 rounds half away from zero -> c + 1; last column falls out of bounds
    assert w[1, 0] == -1 and w[1, 1] == cols * 1 + 0 and w[1, cols - 1] == cols * 1 + cols - 2
    assert w[0, 1] == -1 and w[0, 2] == -1      # sources (0,0) and (0,1) are depth-invalid
    e = o.eval(0, np.array([0.125, 0, 0, 0, 0, 0]))
    assert e["num_valid"] == rows * (cols - 1) - 2
    # -0.5 shift: tc = c - 0.5 rounds AWAY from zero: back to c for c >= 1, and -0.5 -> -1 (out) for c = 0
    w = o.winner_map(0, np.array([-0.125, 0, 0, 0, 0, 0])).reshape(rows, cols)
    assert w[1, 0] == -1 and w[1, 1] == cols + 1 and w[1, cols - 1] == 2 * cols - 1
    # tc in (-0.5, 0) rounds to -0 and is accepted as column 0
    w = o.winner_map(0, np.array([-0.0625, 0, 0, 0, 0, 0])).reshape(rows, cols)
    assert w[1, 0] == cols + 0


def test_zero_iteration_levels_and_step_then_stop(oracle):
    o, _ = _setup(oracle, levels=4, iters=(0, 0, 0, 50), min_grad_norm=1e9)
    o.set_initial_state(np.zeros(6))
    o.optimize()
    log = o.iter_stats()
    # threshold huge -> the first gradient is already "small": one step is still applied, then stop
    assert len(log) == 1 and log[0]["level"] == 3
    assert np.any(o.state() != 0) and np.array_equal(o.state(), log[0]["state_out"])
    o2, _ = _setup(oracle, levels=4, iters=(0, 0, 0, 0))
    o2.set_initial_state(np.array([1e-3] * 6))
    o2.optimize()
    assert len(o2.iter_stats()) == 0 and np.array_equal(o2.state(), np.array([1e-3] * 6))
    o3, _ = _setup(oracle, levels=4, iters=(0, 0, 0, 3), min_grad_norm=0.0)
    o3.set_initial_state(np.zeros(6))
    o3.optimize()
    assert len(o3.iter_stats()) == 3          # max-iterations stop


def test_ceres_residual_matches_numpy(oracle, nr):
    rows, cols = 24, 32
    g0, d0, g1, _ = synth.make_pair(rows, cols, K=np.array([[30., 0, 15.5], [0, 30., 11.5], [0, 0, 1]]), seed=5)
    K = np.array([[30., 0, 15.5], [0, 30., 11.5], [0, 0, 1]])
    cfg = oracle.make_config(mode=2, num_levels=2, max_iters=(5, 5))
    o = oracle.Oracle(cfg, K)
    o.set_source(g0, d0)
    o.set_target(g1)
    st = np.array([0.01, -0.02, 0.015, 0.01, -0.008, 0.006])
    I0, D0, I1, Gx, Gy = nr.build_pyramids(g0, d0, g1, 2, [0, 0], [0.0625] * 2)
    for lvl in (0, 1):
        e = o.eval(lvl, st, want_residuals=True, want_jacobian=True)
        r, J = nr.ceres_eval(I0[lvl], D0[lvl], I1[lvl], Gx[lvl], Gy[lvl], K, lvl, st)
        assert np.max(np.abs(e["residuals"] - r)) < 1e-13
        assert np.max(np.abs(e["jacobian"] - J)) < 1e-9 * max(1.0, np.max(np.abs(J)))
        assert abs(e["cost"] - 0.5 * r @ r) < 1e-12


def test_ceres_jacobian_is_derivative_of_residual_away_from_pixel_borders(oracle):
    """Finite differences of the oracle's Ceres-mode residual reproduce its Jacobian on rows whose
    sample stays inside the same bilinear cell and the same target slot."""
    rows, cols = 24, 32
    K = np.array([[30., 0, 15.5], [0, 30., 11.5], [0, 0, 1]])
    g0, d0, g1, _ = synth.make_pair(rows, cols, K=K, seed=6, holes=False)
    cfg = oracle.make_config(mode=2, num_levels=1, max_iters=(5,), grad_scale=1 / 32.)   # true derivative scale
    o = oracle.Oracle(cfg, K)
    o.set_source(g0, d0)
    o.set_target(g1)
    st = np.array([0.004, -0.003, 0.002, 0.003, -0.002, 0.001])
    e = o.eval(0, st, want_residuals=True, want_jacobian=True)
    # the Jacobian chains the *Scharr* gradient (a smoothed central difference), not the derivative of
    # the bilinear interpolant, so only a loose agreement is expected: same sign and magnitude on average
    h = 1e-6
    num = np.zeros_like(e["jacobian"])
    for k in range(6):
        sp_, sm_ = st.copy(), st.copy()
        sp_[k] += h
        sm_[k] -= h
        rp = o.eval(0, sp_, want_residuals=True)["residuals"]
        rm = o.eval(0, sm_, want_residuals=True)["residuals"]
        num[:, k] = (rp - rm) / (2 * h)
    ok = np.all(np.abs(num) < 1e3, axis=1) & (e["residuals"] != 0)
    c = np.corrcoef(num[ok].ravel(), e["jacobian"][ok].ravel())[0, 1]
    assert c > 0.8
