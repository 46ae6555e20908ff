"""End to end on a recorded sequence (SURVEY 8(f) rows 1 + 4): TUM-style directory -> background decode
into pinned buffers -> CUDA solver with the target pyramid promoted on the device -> TUM trajectory,
against the CPU oracle run frame by frame like PhotoconsistencyVisualOdometry.cpp:196-262."""
import importlib

import numpy as np
import pytest

from test_dataset_reader import write_sequence
from test_gpu_parity import conv_cfg

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def phovo():
    m = importlib.import_module("photoconsistency-visual-odometry_b200")
    m.build()
    return m


@pytest.fixture(scope="module")
def oracle():
    import oracle_py
    oracle_py.build()
    return oracle_py


def test_visual_odometry_over_a_recorded_sequence(phovo, oracle, tmp_path):
    ds = phovo.dataset
    n = 6
    K, frames = write_sequence(phovo, str(tmp_path), n=n, rows=120, cols=160)
    cfg = phovo.default_config()
    cfg.num_levels = 3
    for l, m in enumerate((4, 8, 12)):
        cfg.max_num_iterations[l] = m
        cfg.min_gradient_norm[l] = 30.
    odo = phovo.CPhotoconsistencyOdometryCuda()
    odo.SetConfig(cfg)
    odo.SetIntrinsicMatrix(K)
    traj = str(tmp_path / "trajectory.txt")
    seen = []
    poses = ds.run_visual_odometry(odo, ds.PrefetchingSource(ds.open_rgbd_dataset(str(tmp_path)), ahead=3), traj,
                                   on_frame=lambda k, item, Rt: seen.append(k))
    assert seen == list(range(1, n)) and len(poses) == n - 1
    # the app's loop on the CPU oracle: depth = double(raw) * 1/5000 (:208, :220), zero state every frame (:224)
    pose = np.eye(4)
    got = np.loadtxt(traj)
    assert got.shape == (n - 1, 8)
    for k in range(1, n):
        o = oracle.Oracle(conv_cfg(oracle, cfg), K)
        o.set_source(frames[k - 1][1], frames[k - 1][2].astype(np.float64) * (1. / 5000.))
        o.set_target(frames[k][1])
        o.set_initial_state(np.zeros(6))
        o.optimize()
        pose = pose @ np.linalg.inv(o.rt())
        ts, P = poses[k - 1]
        assert ts == pytest.approx(frames[k][0], abs=1e-6)
        assert np.max(np.abs(P - pose)) < 1e-9
        assert got[k - 1, 0] == pytest.approx(frames[k][0], abs=1e-6)
        assert np.max(np.abs(got[k - 1, 1:4] - pose[:3, 3])) < 1e-9
        assert np.max(np.abs(got[k - 1, 4:8] - ds.quaternion_of(pose[:3, :3]))) < 1e-9
    # rebuilding the source pyramid every frame (what the app does) gives bitwise the same trajectory
    poses2 = ds.run_visual_odometry(odo, ds.open_rgbd_dataset(str(tmp_path)), promote=False)
    assert all(np.array_equal(a[1], b[1]) for a, b in zip(poses, poses2))
