"""Generates the committed golden fixtures in tests/golden/ (run in the build container).

The reference ships no test vectors, so these are minted from the two independent restatements:
third-party image arithmetic comes from the REAL OpenCV (cv2), the solver quantities from the
numpy restatement (oracle/np_restatement.py, which itself calls cv2 for the pyramids).  The C
oracle is NOT used to produce them -- tests check the C oracle (CPU) and the CUDA path (GPU)
against these files.

    python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cv2  # noqa: E402
import np_restatement as nr  # noqa: E402

synth = importlib.import_module("photoconsistency-visual-odometry_b200.synth")


def golden_cv2_ops():
    # OpenCV's portable C++ code path (what "OpenCV >= 2.4.5" pins); this build's IPP kernels round differently
    cv2.setUseOptimized(False)
    rng = np.random.default_rng(11)
    out = {}
    for tag, shape in (("a", (48, 64)), ("b", (45, 77))):
        img8 = rng.integers(0, 256, shape).astype(np.uint8)
        a = img8.astype(np.float64) * (1. / 255)
        out[tag + "_u8"] = img8
        for lvl in (1, 2, 3):
            out["%s_resize%d" % (tag, lvl)] = cv2.resize(a, (0, 0), fx=0.5 ** lvl, fy=0.5 ** lvl)
        out[tag + "_scharr_x"] = cv2.Scharr(a, cv2.CV_64F, 1, 0, scale=0.0625)
        out[tag + "_scharr_y"] = cv2.Scharr(a, cv2.CV_64F, 0, 1, scale=0.0625)
        out[tag + "_blur3"] = cv2.GaussianBlur(a, (3, 3), 3)
        out[tag + "_blur5"] = cv2.GaussianBlur(a, (5, 5), 3)
    np.savez_compressed(os.path.join(HERE, "cv2_ops.npz"), **out)
    cv2.setUseOptimized(True)


def golden_pair(name, rows, cols, K, levels, iters, min_grad, xi, seed, fixed):
    g0, d0, g1, _ = synth.make_pair(rows, cols, K=K, xi=xi, seed=seed)
    I0, D0, I1, Gx, Gy = nr.build_pyramids(g0, d0, g1, levels, [0] * levels, [0.0625] * levels)
    st, log = nr.analytic_optimize(I0, D0, I1, Gx, Gy, K, levels, iters, [1.] * levels, [min_grad] * levels,
                                   np.zeros(6), fixed=fixed)
    out = dict(gray0=g0, depth0=d0.astype(np.float32), gray1=g1, K=K, levels=levels, iters=np.array(iters),
               min_grad=min_grad, fixed=int(fixed), final_state=st,
               log_level=np.array([e["level"] for e in log]), log_iter=np.array([e["iteration"] for e in log]),
               log_num_valid=np.array([e["num_valid"] for e in log]),
               log_H=np.array([nr.pack_upper(e["H"]) for e in log]), log_g=np.array([e["g"] for e in log]),
               log_state_in=np.array([e["state_in"] for e in log]), log_state_out=np.array([e["state_out"] for e in log]))
    # a few fixed-state evaluations incl. the winner map (collision semantics)
    states = [np.zeros(6), np.array(xi) * 0.5, np.array([0.03, -0.02, 0.04, 0.02, -0.015, 0.01])]
    for i, s in enumerate(states):
        lvl = levels - 1 - (i % 2)
        H, g, cnt, res, win = nr.analytic_eval(I0[lvl], D0[lvl], I1[lvl], Gx[lvl], Gy[lvl], K, lvl, s, fixed=fixed)
        out["eval%d_level" % i] = lvl
        out["eval%d_state" % i] = s
        out["eval%d_H" % i] = nr.pack_upper(H)
        out["eval%d_g" % i] = g
        out["eval%d_count" % i] = cnt
        out["eval%d_winner" % i] = win.astype(np.int32)
        out["eval%d_res" % i] = res
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "iterations", len(log), "final", st)


def golden_ceres(name, rows, cols, K, seed):
    g0, d0, g1, _ = synth.make_pair(rows, cols, K=K, seed=seed)
    I0, D0, I1, Gx, Gy = nr.build_pyramids(g0, d0, g1, 2, [0, 0], [0.0625] * 2)
    st = np.array([0.01, -0.02, 0.015, 0.01, -0.008, 0.006])
    out = dict(gray0=g0, depth0=d0.astype(np.float32), gray1=g1, K=K, state=st)
    for lvl in (0, 1):
        r, J = nr.ceres_eval(I0[lvl], D0[lvl], I1[lvl], Gx[lvl], Gy[lvl], K, lvl, st)
        out["res%d" % lvl] = r
        out["jac%d" % lvl] = J
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    golden_cv2_ops()
    Ks = np.array([[105., 0, 63.5], [0, 105., 47.5], [0, 0, 1]])
    golden_pair("pair_96x128_ref", 96, 128, Ks, 3, [0, 10, 20], 100., synth.XI_CONFIG1, 21, False)
    golden_pair("pair_96x128_fixed", 96, 128, Ks, 3, [0, 10, 20], 100., synth.XI_CONFIG1, 21, True)
    golden_pair("pair_90x135_ref", 90, 135, np.array([[110., 0, 67.], [0, 108., 44.5], [0, 0, 1]]), 3, [4, 6, 8], 150.,
                [0.01, 0.008, -0.012, -0.006, 0.007, 0.004], 22, False)
    golden_ceres("ceres_24x32", 24, 32, np.array([[30., 0, 15.5], [0, 30., 11.5], [0, 0, 1]]), 5)
