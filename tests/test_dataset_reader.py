"""Recorded-sequence reader: the host-side mirror of CCameraRecord /
CMultiSensorDataSource (phovo/include/CCameraRecord.h:63-108, CMultiSensorDataSource.h:73-92) and of
the trajectory line of the VO app (PhotoconsistencyVisualOdometry.cpp:234-243).  No GPU needed."""
import importlib
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def phovo():
    return importlib.import_module("photoconsistency-visual-odometry_b200")


def write_sequence(phovo, directory, n=4, rows=48, cols=64, colour=True):
    """A TUM-style directory: rgb.txt / depth.txt with comment lines, 8-bit (colour) PNGs, 16-bit depth PNGs."""
    K = np.array([[60., 0, 31.5], [0, 60., 23.5], [0, 0, 1]])
    os.makedirs(os.path.join(directory, "rgb"))
    os.makedirs(os.path.join(directory, "depth"))
    frames = []
    with open(os.path.join(directory, "rgb.txt"), "w") as fr, open(os.path.join(directory, "depth.txt"), "w") as fd:
        fr.write("# color images\n# file: 'synthetic'\n# timestamp filename\n")
        fd.write("# depth maps\n# timestamp filename\n\n")
        for k in range(n):
            g, d = phovo.synth.make_sequence_frame(k, rows, cols, K=K)
            raw = np.clip(np.rint(d * 5000.), 0, 65535).astype(np.uint16)
            ts = 1305031102.175304 + 0.033 * k
            name = "%.6f.png" % ts
            img = np.dstack([g, g, g]) if colour else g          # grey as BGR: imread(file, 0) gives g back exactly
            assert cv2.imwrite(os.path.join(directory, "rgb", name), img)
            assert cv2.imwrite(os.path.join(directory, "depth", name), raw)
            fr.write("%.6f rgb/%s\n" % (ts, name))
            fd.write("%.6f depth/%s\n" % (ts + 0.011, name))
            frames.append((ts, g, raw))
    return K, frames


def test_camera_record_reads_lines_in_order_and_decodes_like_the_reference(phovo, tmp_path):
    ds = phovo.dataset
    K, frames = write_sequence(phovo, str(tmp_path))
    rec = ds.CCameraRecord(True)
    rec.SetFileName(str(tmp_path / "rgb.txt"))
    rec.Start()
    for ts, g, _ in frames:
        sd = rec.GetSensorData()
        assert sd.GetTimeStamp() == pytest.approx(ts, abs=1e-6)
        assert sd.GetData().dtype == np.uint8 and sd.GetData().ndim == 2 and np.array_equal(sd.GetData(), g)
    assert rec.GetSensorData() is None and rec.GetSensorData() is None      # exhausted: keeps returning nothing
    rec.Stop()
    dep = ds.CCameraRecord(False)
    dep.SetFileName(str(tmp_path / "depth.txt"))
    dep.Start()
    sd = dep.GetSensorData()
    assert sd.GetData().dtype == np.uint16 and np.array_equal(sd.GetData(), frames[0][2])   # raw, unchanged (imread -1)
    dep.Stop()


def test_missing_files_fail_loudly(phovo, tmp_path):
    ds = phovo.dataset
    rec = ds.CCameraRecord(True)
    rec.SetFileName(str(tmp_path / "nope.txt"))
    with pytest.raises(RuntimeError, match="Unable to open camera record file"):
        rec.Start()
    with pytest.raises(RuntimeError, match="does not exist"):
        ds.open_rgbd_dataset(str(tmp_path))
    (tmp_path / "rgb.txt").write_text("1.0 rgb/missing.png\n")
    rec.SetFileName(str(tmp_path / "rgb.txt"))
    rec.Start()
    with pytest.raises(RuntimeError, match="Unable to read image"):
        rec.GetSensorData()


def test_multi_sensor_source_pairs_by_order_and_stops_with_the_shorter_record(phovo, tmp_path):
    ds = phovo.dataset
    _, frames = write_sequence(phovo, str(tmp_path), n=4)
    # drop the last depth line: the pair stream ends one frame early (CMultiSensorDataSource.h:81-86)
    lines = (tmp_path / "depth.txt").read_text().splitlines()
    (tmp_path / "depth.txt").write_text("\n".join(lines[:-1]) + "\n")
    src = ds.open_rgbd_dataset(str(tmp_path))
    src.Start()
    got = []
    while True:
        item = src.GetMultiSensorData()
        if item is None:
            break
        got.append(item)
    src.Stop()
    assert len(got) == 3
    for (ts, g, raw), item in zip(frames, got):
        assert np.array_equal(item[ds.IntensityCameraIdentifier].GetData(), g)
        assert np.array_equal(item[ds.DepthCameraIdentifier].GetData(), raw)
        assert item[ds.DepthCameraIdentifier].GetTimeStamp() == pytest.approx(ts + 0.011, abs=1e-6)   # own timestamps, no association


@pytest.mark.parametrize("workers", [1, 3])
def test_prefetching_source_yields_the_same_frames(phovo, tmp_path, workers):
    ds = phovo.dataset
    _, frames = write_sequence(phovo, str(tmp_path), n=7 if workers == 1 else 17)
    import time
    held = []
    for k, item in enumerate(ds.PrefetchingSource(ds.open_rgbd_dataset(str(tmp_path)), ahead=2, pin=False, workers=workers)):
        held.append(item)
        time.sleep(0.05)          # let the worker run as far ahead as it can: it must not recycle a slot still in use
        # the consumer keeps the previous frame while it works on the current one: both must still be intact
        for j in (k - 1, k):
            if j >= 0:
                assert np.array_equal(held[j][ds.IntensityCameraIdentifier].GetData(), frames[j][1])
                assert np.array_equal(held[j][ds.DepthCameraIdentifier].GetData(), frames[j][2])
    assert len(held) == len(frames)
    # errors of the worker thread surface in the consumer
    (tmp_path / "rgb.txt").write_text("1.0 rgb/missing.png\n")
    with pytest.raises(RuntimeError, match="Unable to read image"):
        list(ds.PrefetchingSource(ds.open_rgbd_dataset(str(tmp_path)), pin=False, workers=workers))


def test_quaternion_and_trajectory_line(phovo):
    ds = phovo.dataset
    rng = np.random.default_rng(0)
    for _ in range(50):
        s = rng.uniform(-3.1, 3.1, 6)
        Rt = phovo.state_to_rt(s)
        q = ds.quaternion_of(Rt[:3, :3])
        x, y, z, w = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert abs(np.linalg.norm(q) - 1) < 1e-12 and np.max(np.abs(R - Rt[:3, :3])) < 1e-12
    line = ds.trajectory_line(1305031102.175304, np.eye(4))
    f = line.split()
    assert len(f) == 8 and f[0] == "1305031102.175304" and [float(v) for v in f[1:]] == [0, 0, 0, 0, 0, 0, 1]
