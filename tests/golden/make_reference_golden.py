"""Mints tests/golden/ref_*.npz from the REFERENCE ITSELF: the reference's own
CPhotoconsistencyOdometryAnalytic.h compiled unmodified into oracle/_ref/libphovo_ref.so
(oracle/Makefile; OpenCV/Eigen replaced by oracle/shim) and run on the synthetic pairs below.
Run in the build container (needs /root/reference):

    python tests/golden/make_reference_golden.py

Each file holds the inputs (gray0, depth0, gray1, K, config name), the reference's final state and
Rt, and per executed iteration the level size n, J^T J (6x6) and J^T r it formed (AN:538-540).
"""
import importlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_py  # noqa: E402

phovo = importlib.import_module("photoconsistency-visual-odometry_b200")

CASES = [
    # name, rows, cols, K, config, seed
    ("ref_pair_120x160_cfg4", 120, 160, np.array([[131.25, 0, 79.5], [0, 131.25, 59.5], [0, 0, 1.]]), "config_4_level_optimization_analytic", 31),
    ("ref_pair_240x320_cfg4", 240, 320, np.array([[262.5, 0, 159.5], [0, 262.5, 119.5], [0, 0, 1.]]), "config_4_level_optimization_analytic", 33),
    ("ref_pair_96x128_all_levels", 96, 128, np.array([[105., 0, 63.5], [0, 105., 47.5], [0, 0, 1.]]), "test_3_level_all_active", 34),
    ("ref_pair_135x241_cfg5", 135, 241, np.array([[190., 0, 121.3], [0, 188., 66.1], [0, 0, 1.]]), "config_5_level_optimization_analytic", 32),
]


def main():
    tmp = tempfile.mkdtemp()
    for name, rows, cols, K, cfg, seed in CASES:
        g0, d0, g1, xi = phovo.synth.make_pair(rows, cols, K=K, seed=seed)
        yml = phovo.configs.write_yaml(cfg, tmp)
        s, rt, iters = ref_py.Reference(yml, K).align(g0, d0, g1)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), gray0=g0, depth0=d0.astype(np.float32), gray1=g1, K=K,
                            config=np.array(cfg), state=s, rt=rt, n=np.array([it["n"] for it in iters]),
                            H=np.array([it["H"] for it in iters]), g=np.array([it["g"] for it in iters]))
        print(name, len(iters), "iterations", s)


def warp_golden():
    """phovo::warpImage (CPhotoconsistencyOdometry.h:73-134) run from the reference source."""
    K = np.array([[131.25, 0., 79.5], [0., 131.25, 59.5], [0., 0., 1.]])
    g0, d0, g1, xi = phovo.synth.make_pair(120, 160, K=K, seed=35)
    rt = phovo.state_to_rt(np.array([0.03, -0.02, 0.05, 0.02, -0.015, 0.01]))
    out = {"gray0": g0, "depth0": d0.astype(np.float32), "gray1": g1, "K": K, "rt": rt}
    for level in (0, 1):
        Kl = K.copy(); Kl[:2] /= 2 ** level      # the reference divides K by 2^level inside warpImage; level 0 here, scaled K for level 1
        out["warped_l%d" % level] = ref_py.warp_image(g0, d0, rt, Kl)
    np.savez_compressed(os.path.join(HERE, "ref_warp_image_120x160.npz"), **out)
    print("ref_warp_image_120x160", int((out["warped_l0"] > 0).sum()), "pixels written")


def biobjective_golden():
    """CPhotoconsistencyOdometryBiObjective.h run from the reference source (photometric + depth rows)."""
    tmp = tempfile.mkdtemp()
    cases = [("ref_bi_120x160_all_levels", 120, 160, np.array([[131.25, 0, 79.5], [0, 131.25, 59.5], [0, 0, 1.]]), "test_3_level_all_active", 37),
             ("ref_bi_240x320_cfg4", 240, 320, np.array([[262.5, 0, 159.5], [0, 262.5, 119.5], [0, 0, 1.]]), "config_4_level_optimization_analytic", 38)]
    for name, rows, cols, K, cfg, seed in cases:
        g0, d0, g1, d1 = phovo.synth.make_pair(rows, cols, K=K, seed=seed)
        yml = phovo.configs.write_yaml(cfg, tmp)
        s, iters = ref_py.ReferenceBiObjective(yml, K).align(g0, d0, g1, d1)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), gray0=g0, depth0=d0.astype(np.float32), gray1=g1, depth1=d1.astype(np.float32),
                            K=K, config=np.array(cfg), state=s, n=np.array([it["n"] for it in iters]),
                            H=np.array([it["H"] for it in iters]), g=np.array([it["g"] for it in iters]))
        print(name, len(iters), "iterations", s)


if __name__ == "__main__":
    warp_golden()
    biobjective_golden()
    main()
