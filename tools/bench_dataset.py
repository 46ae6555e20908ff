#!/usr/bin/env python
"""End-to-end visual odometry over a recorded sequence ON DISK: writes a synthetic
640x480 TUM-style directory (colour PNGs + 16-bit depth PNGs, rgb.txt / depth.txt), then runs the VO
app's loop (dataset.run_visual_odometry) with the PNG decode pool at several widths.  PNG decoding is
the slowest stage by far, so frames/s scales with the decode workers until the solver's ~0.2 ms per frame
is reached.  Prints one JSON line."""
import argparse, importlib, json, os, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=240)
    ap.add_argument("--workers", type=int, nargs="*", default=[1, 2, 4, 8, 16])
    args = ap.parse_args()
    import cv2
    phovo = importlib.import_module("photoconsistency-visual-odometry_b200")
    phovo.build()
    ds = phovo.dataset
    K = phovo.synth.K_VISUAL_ODOMETRY
    d = tempfile.mkdtemp(prefix="phovo_seq_")
    try:
        os.makedirs(os.path.join(d, "rgb")); os.makedirs(os.path.join(d, "depth"))
        with open(os.path.join(d, "rgb.txt"), "w") as fr, open(os.path.join(d, "depth.txt"), "w") as fd:
            fr.write("# color images\n"); fd.write("# depth maps\n")
            for k in range(args.frames):
                g, z = phovo.synth.make_sequence_frame(k, 480, 640, K=K)
                ts = 1305031102.175304 + k / 30.
                cv2.imwrite(os.path.join(d, "rgb", "%06d.png" % k), np.dstack([g, g, g]))
                cv2.imwrite(os.path.join(d, "depth", "%06d.png" % k), np.clip(np.rint(z * 5000.), 0, 65535).astype(np.uint16))
                fr.write("%.6f rgb/%06d.png\n" % (ts, k)); fd.write("%.6f depth/%06d.png\n" % (ts, k))
        png_bytes = sum(os.path.getsize(os.path.join(d, s, f)) for s in ("rgb", "depth") for f in os.listdir(os.path.join(d, s)))
        cfg = phovo.configs.to_config("config_5_level_optimization_analytic", phovo.capi)
        odo = phovo.CPhotoconsistencyOdometryCuda(); odo.SetConfig(cfg); odo.SetIntrinsicMatrix(K)
        results, ref = {}, None
        for w in args.workers:
            t0 = time.perf_counter()
            poses = ds.run_visual_odometry(odo, ds.PrefetchingSource(ds.open_rgbd_dataset(d), ahead=8, workers=w), os.path.join(d, "traj_%d.txt" % w))
            dt = time.perf_counter() - t0
            results[str(w)] = {"frames_per_s": (args.frames - 1) / dt, "seconds": dt}
            last = poses[-1][1]
            if ref is None:
                ref = last
            assert np.array_equal(ref, last), "trajectory depends on the number of decode workers"
        print(json.dumps({"workload": "VO over %d recorded 640x480 frames on disk (PNG), config_5_level_optimization_analytic" % args.frames,
                          "png_bytes_per_frame": png_bytes / args.frames, "host_cpus": len(os.sched_getaffinity(0)),
                          "decode_workers": results, "final_translation": ref[:3, 3].tolist()}))
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
