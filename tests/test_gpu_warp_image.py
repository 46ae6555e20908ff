"""phovo::warpImage + absdiff (CPhotoconsistencyOdometry.h:73-134; the apps' post-Optimize display)
on the GPU against outputs of the REFERENCE'S OWN function (tests/golden/ref_warp_image_120x160.npz,
minted from oracle/_ref by tests/golden/make_reference_golden.py) and, when the prebuilt reference
library is present, live on other inputs.  u8 images: bit-exact."""
import numpy as np
import pytest

from helpers import load_golden

pytestmark = pytest.mark.gpu


def test_warp_image_matches_reference_golden(phovo):
    gd = load_golden("ref_warp_image_120x160")
    odo = phovo.CPhotoconsistencyOdometryCuda()
    d0 = gd["depth0"].astype(np.float64)
    for level in (0, 1):
        K = gd["K"].copy()
        w = odo.WarpImage(gd["gray0"], d0, gd["rt"], K, level=level)
        assert np.array_equal(w, gd["warped_l%d" % level]), level
    # with the target: |target - warped| as the apps show it
    w, diff = odo.WarpImage(gd["gray0"], d0, gd["rt"], gd["K"], targetImage=gd["gray1"])
    assert np.array_equal(w, gd["warped_l0"])
    assert np.array_equal(diff, np.abs(gd["gray1"].astype(np.int16) - w.astype(np.int16)).astype(np.uint8))
    # f32 depth (f32-representable here) and strided inputs give the same image
    assert np.array_equal(odo.WarpImage(gd["gray0"], gd["depth0"], gd["rt"], gd["K"]), gd["warped_l0"])
    gp = np.zeros((120, 192), np.uint8); gp[:, :160] = gd["gray0"]
    dp = np.zeros((120, 176)); dp[:, :160] = d0
    assert np.array_equal(odo.WarpImage(gp[:, :160], dp[:, :160], gd["rt"], gd["K"]), gd["warped_l0"])


def test_warp_image_matches_reference_source_live(phovo):
    import ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref library not present")
    rng = np.random.default_rng(5)
    K = phovo.synth.K_FRAME_ALIGNMENT
    g0, d0, g1, _ = phovo.synth.make_pair(480, 640, K=K, seed=36)
    odo = phovo.CPhotoconsistencyOdometryCuda()
    for _ in range(3):
        st = np.concatenate([rng.uniform(-0.2, 0.2, 3), rng.uniform(-0.1, 0.1, 3)])
        rt = phovo.state_to_rt(st)
        assert np.array_equal(odo.WarpImage(g0, d0, rt, K), ref_py.warp_image(g0, d0, rt, K))
    # raw u16 depth x 1/5000 (VisualOdometry.cpp:163) == the doubles the app would hand over
    raw = np.clip(np.rint(d0 * 5000.), 0, 65535).astype(np.uint16)
    rt = phovo.state_to_rt(np.array([0.05, 0.02, -0.03, 0.02, 0.01, -0.02]))
    assert np.array_equal(odo.WarpImage(g0, raw, rt, K, depth_scale=1. / 5000.),
                          ref_py.warp_image(g0, raw.astype(np.float64) * (1. / 5000.), rt, K))


def test_warp_image_error_paths(phovo):
    odo = phovo.CPhotoconsistencyOdometryCuda()
    g = np.zeros((4, 4), np.uint8)
    with pytest.raises(phovo.PhovoError):
        odo.WarpImage(g, np.ones((4, 4)), np.eye(4), np.eye(3), level=-1)
