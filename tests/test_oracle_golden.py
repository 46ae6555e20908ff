"""The C oracle against the committed golden vectors (tests/golden/*.npz, minted by
tests/golden/make_golden.py from real OpenCV + the numpy restatement).  CPU only."""
import numpy as np
import pytest

from helpers import assert_logs_match, golden_log, load_golden


def ulps(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.spacing(np.maximum(np.abs(a), np.abs(b))), 1e-300))


def test_image_ops_against_opencv_golden(oracle):
    gd = load_golden("cv2_ops")
    for tag in ("a", "b"):
        a = oracle.convert_u8(gd[tag + "_u8"])
        for lvl in (1, 2, 3):
            ref = gd["%s_resize%d" % (tag, lvl)]
            mine = oracle.resize_level(a, lvl)
            assert mine.shape == ref.shape and np.array_equal(mine, ref)      # bit for bit (incl. the area path of level 1)
        assert np.array_equal(oracle.scharr(a, 1, 0, 0.0625), gd[tag + "_scharr_x"])
        assert np.array_equal(oracle.scharr(a, 0, 1, 0.0625), gd[tag + "_scharr_y"])
        assert np.max(np.abs(oracle.gaussian_blur(a, 3) - gd[tag + "_blur3"])) < 1e-15
        assert np.max(np.abs(oracle.gaussian_blur(a, 5) - gd[tag + "_blur5"])) < 1e-15


@pytest.mark.parametrize("name", ["pair_96x128_ref", "pair_96x128_fixed", "pair_90x135_ref"])
def test_analytic_against_golden(oracle, name):
    gd = load_golden(name)
    levels = int(gd["levels"])
    cfg = oracle.make_config(mode=int(gd["fixed"]), num_levels=levels, max_iters=tuple(int(v) for v in gd["iters"]),
                             min_grad_norm=float(gd["min_grad"]))
    o = oracle.Oracle(cfg, gd["K"])
    o.set_source(gd["gray0"], gd["depth0"].astype(np.float64))
    o.set_target(gd["gray1"])
    o.set_initial_state(np.zeros(6))
    o.optimize()
    assert_logs_match(o.iter_stats(), golden_log(gd), rel=1e-9, what=name)
    assert np.max(np.abs(o.state() - gd["final_state"])) < 1e-11
    for i in range(3):
        lvl, st = int(gd["eval%d_level" % i]), gd["eval%d_state" % i]
        e = o.eval(lvl, st, want_residuals=True)
        assert e["num_valid"] == int(gd["eval%d_count" % i])
        assert np.array_equal(o.winner_map(lvl, st), gd["eval%d_winner" % i])
        assert np.max(np.abs(e["residuals"] - gd["eval%d_res" % i])) < 1e-14
        assert np.max(np.abs(e["H"] - gd["eval%d_H" % i]) / np.abs(gd["eval%d_H" % i]).max()) < 1e-12
        assert np.max(np.abs(e["g"] - gd["eval%d_g" % i])) < 1e-10 * np.abs(gd["eval%d_g" % i]).max()


def test_ceres_residual_against_golden(oracle):
    gd = load_golden("ceres_24x32")
    cfg = oracle.make_config(mode=2, num_levels=2, max_iters=(5, 5))
    o = oracle.Oracle(cfg, gd["K"])
    o.set_source(gd["gray0"], gd["depth0"].astype(np.float64))
    o.set_target(gd["gray1"])
    for lvl in (0, 1):
        e = o.eval(lvl, gd["state"], want_residuals=True, want_jacobian=True)
        assert np.max(np.abs(e["residuals"] - gd["res%d" % lvl])) < 1e-13
        assert np.max(np.abs(e["jacobian"] - gd["jac%d" % lvl])) < 1e-9 * max(1.0, np.abs(gd["jac%d" % lvl]).max())
