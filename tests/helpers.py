"""Shared comparison helpers for the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north-star tolerances (BASELINE.json): per-level J^T J / J^T r within 1e-5 relative,
# final pose within 1e-4 m translation and 1e-5 rad rotation, identical iteration counts.
REL_NORMAL_EQ = 1e-5
TOL_TRANS_M = 1e-4
TOL_ROT_RAD = 1e-5


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def h_rel_err(H, Href):
    """Relative error of the 21 packed J^T J entries, each measured against the geometric mean of
    its two diagonal entries (entries of a Gram matrix are bounded by it), so near-zero
    off-diagonals are not compared against themselves."""
    idx = [(a, b) for a in range(6) for b in range(a, 6)]
    diag = {}
    for k, (a, b) in enumerate(idx):
        if a == b:
            diag[a] = abs(Href[k])
    scale = np.array([np.sqrt(diag[a] * diag[b]) for a, b in idx])
    return float(np.max(np.abs(np.asarray(H) - np.asarray(Href)) / scale))


def g_rel_err(g, gref):
    """Relative error of J^T r measured against its largest entry."""
    return float(np.max(np.abs(np.asarray(g) - np.asarray(gref))) / np.max(np.abs(gref)))


def assert_logs_match(log, ref, rel=REL_NORMAL_EQ, what=""):
    assert len(log) == len(ref), "%s executed %d iterations, expected %d" % (what, len(log), len(ref))
    for a, b in zip(log, ref):
        assert (a["level"], a["iteration"]) == (b["level"], b["iteration"])
        assert a["num_valid"] == b["num_valid"], (what, a["level"], a["iteration"], a["num_valid"], b["num_valid"])
        eh, eg = h_rel_err(a["H"], b["H"]), g_rel_err(a["g"], b["g"])
        assert eh < rel, (what, a["level"], a["iteration"], "H", eh)
        assert eg < rel, (what, a["level"], a["iteration"], "g", eg)


def assert_pose_close(s, sref, what=""):
    s, sref = np.asarray(s), np.asarray(sref)
    assert np.max(np.abs(s[:3] - sref[:3])) < TOL_TRANS_M, (what, s, sref)
    assert np.max(np.abs(s[3:] - sref[3:])) < TOL_ROT_RAD, (what, s, sref)


def golden_log(gd):
    return [dict(level=int(gd["log_level"][i]), iteration=int(gd["log_iter"][i]), num_valid=int(gd["log_num_valid"][i]),
                 H=gd["log_H"][i], g=gd["log_g"][i], state_in=gd["log_state_in"][i], state_out=gd["log_state_out"][i])
            for i in range(len(gd["log_level"]))]
