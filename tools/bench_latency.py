#!/usr/bin/env python
"""Per-pair latency of the general (non-batched) path -- BASELINE configs[0] and configs[1]:

  single : PhotoconsistencyFrameAlignment on one 640x480 pair, config_4_level_optimization_analytic:
           Optimize() alone (what the app's TickMeter brackets, FrameAlignment.cpp:99-101) and
           SetSourceFrame + SetTargetFrame + Optimize from host buffers;
  vo     : PhotoconsistencyVisualOdometry loop over a synthetic sequence, config_5_level_optimization_analytic,
           sequential, 1 GPU, per-frame latency (host frame in -> pose out), target pyramid promoted to
           source between frames.

Each mode prints one JSON line; the CPU oracle is timed beside it on the same inputs (1 thread, like the
reference).  Not the driver's bench (that is bench.py); results are kept under profiles/."""
import argparse, importlib, json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="single", choices=["single", "vo", "ceres"])
    ap.add_argument("--frames", type=int, default=100)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--path", type=int, default=2, help="iteration-loop driver: 2 cooperative kernel, 1 CUDA graph, 0 stream")
    args = ap.parse_args()
    phovo = importlib.import_module("photoconsistency-visual-odometry_b200")
    phovo.build()
    import oracle_py
    import torch
    oracle_py.build()

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t

    if args.mode == "ceres":
        # BASELINE configs[2]: config_5_level_optimization_ceres (levels 0 and 1 ARE optimised): GPU residual /
        # Jacobian evaluation + restated LM against the oracle's restated LM on the CPU
        name, K = "config_5_level_optimization_ceres", phovo.synth.K_FRAME_ALIGNMENT
        cfg = phovo.configs.to_config(name, phovo.capi)
        g0, d0, g1, _ = phovo.synth.make_pair(480, 640, K=K, seed=0)
        odo = phovo.CPhotoconsistencyOdometryCuda()
        odo.SetConfig(cfg); odo.SetIntrinsicMatrix(K)
        wall = []
        for rep in range(args.reps // 5 + 3):
            odo.SetSourceFrame(g0, d0); odo.SetTargetFrame(g1); odo.SetInitialStateVector(np.zeros(6))
            t0 = time.perf_counter(); odo.Optimize(); wall.append((time.perf_counter() - t0) * 1e3)
        s, log = odo.GetOptimalStateVector(), odo.IterationStats()
        o = oracle_py.Oracle(oracle_py.Config.from_buffer_copy(bytes(cfg)), K)
        o.set_source(g0, d0); o.set_target(g1); o.set_initial_state(np.zeros(6))
        t0 = time.perf_counter(); o.optimize(); cpu = (time.perf_counter() - t0) * 1e3
        print(json.dumps({"mode": "ceres", "config": name, "lm_iterations": len(log), "accepted": int(sum(e["accepted"] for e in log)),
                          "gpu_optimize_ms_wall_median": float(np.median(wall[2:])), "cpu_oracle_optimize_ms": cpu,
                          "pose_abs_diff_vs_oracle": float(np.max(np.abs(s - o.state()))),
                          "note": "LM loop restated from Ceres' documented algorithm on both sides; parity with real Ceres unpinned"}))
        return
    if args.mode == "single":
        name, K = "config_4_level_optimization_analytic", phovo.synth.K_FRAME_ALIGNMENT
        cfg = phovo.configs.to_config(name, phovo.capi)
        g0, d0, g1, _ = phovo.synth.make_pair(480, 640, K=K, seed=0)
        odo = phovo.CPhotoconsistencyOdometryCuda()
        odo.SetConfig(cfg); odo.SetIntrinsicMatrix(K); odo.SetExecution(args.path)
        hg0, hd0, hg1 = pinned(g0), pinned(d0), pinned(g1)
        opt_ms, setup_ms, wall_ms = [], [], []
        for rep in range(args.reps + 5):
            t0 = time.perf_counter()
            odo.SetSourceFrame(hg0, hd0)
            odo.SetTargetFrame(hg1)
            odo.SetInitialStateVector(np.zeros(6))
            odo.Optimize()
            s = odo.GetOptimalStateVector()
            t1 = time.perf_counter()
            a, b = odo.Timings()
            if rep >= 5:
                setup_ms.append(a); opt_ms.append(b); wall_ms.append((t1 - t0) * 1e3)
        log = odo.IterationStats()
        o = oracle_py.Oracle(oracle_py.Config.from_buffer_copy(bytes(cfg)), K)
        cpu_opt, cpu_all = [], []
        for rep in range(5):
            t0 = time.perf_counter()
            o.set_source(g0, d0); o.set_target(g1); o.set_initial_state(np.zeros(6))
            t1 = time.perf_counter()
            o.optimize()
            t2 = time.perf_counter()
            cpu_opt.append((t2 - t1) * 1e3); cpu_all.append((t2 - t0) * 1e3)
        print(json.dumps({"mode": "single", "config": name, "iterations": len(log), "driver": odo.LastPath(),
                          "gpu_optimize_ms_device_median": float(np.median(opt_ms)), "gpu_setup_ms_device_median": float(np.median(setup_ms)),
                          "gpu_host_in_pose_out_ms_wall_median": float(np.median(wall_ms)),
                          "kernel_launches_per_optimize": None,
                          "cpu_oracle_optimize_ms": float(np.median(cpu_opt)), "cpu_oracle_setframes_plus_optimize_ms": float(np.median(cpu_all)),
                          "pose_abs_diff_vs_oracle": float(np.max(np.abs(s - o.state())))}))
    else:
        name, K = "config_5_level_optimization_analytic", phovo.synth.K_VISUAL_ODOMETRY
        cfg = phovo.configs.to_config(name, phovo.capi)
        n = args.frames
        frames = [phovo.synth.make_sequence_frame(k, 480, 640, K=K) for k in range(n)]
        hf = [(pinned(g), pinned(d)) for g, d in frames]
        odo = phovo.CPhotoconsistencyOdometryCuda()
        odo.SetConfig(cfg); odo.SetIntrinsicMatrix(K)
        lat, states, iters = [], [], []
        for rep in range(2):          # first pass warms up (graph build, allocations)
            lat, states, iters = [], [], []
            odo.SetSourceFrame(*hf[0])
            for k in range(1, n):
                t0 = time.perf_counter()
                if k > 1:
                    odo.PromoteTargetToSource(hf[k - 1][1])
                odo.SetTargetFrame(hf[k][0])
                odo.SetInitialStateVector(np.zeros(6))
                odo.Optimize()
                states.append(odo.GetOptimalStateVector())
                lat.append((time.perf_counter() - t0) * 1e3)
                iters.append(len(odo.IterationStats()))
        o = oracle_py.Oracle(oracle_py.Config.from_buffer_copy(bytes(cfg)), K)
        cpu, err = [], 0.
        for k in range(1, min(n, 11)):
            t0 = time.perf_counter()
            o.set_source(*frames[k - 1]); o.set_target(frames[k][0]); o.set_initial_state(np.zeros(6)); o.optimize()
            cpu.append((time.perf_counter() - t0) * 1e3)
            err = max(err, float(np.max(np.abs(o.state() - states[k - 1]))))
        # the same sequence as ONE batch (frames are independent: the app re-zeroes the state every frame)
        gray = np.stack([f[0] for f in frames]); depth = np.stack([f[1] for f in frames]).astype(np.float32)
        tg, td = pinned(gray), pinned(depth)
        seq_ms = []
        for rep in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            st_b, it_b, poses_b = odo.AlignSequence(tg, td)
            seq_ms.append((time.perf_counter() - t0) * 1e3)
        batch_err = float(np.max(np.abs(st_b - np.array(states))))
        print(json.dumps({"mode": "vo", "config": name, "frames": n,
                          "sequence_as_one_batch_ms_wall_median": float(np.median(seq_ms[1:])), "sequence_as_one_batch_frames_per_s": (n - 1) / (float(np.median(seq_ms[1:])) * 1e-3),
                          "sequence_as_one_batch_max_abs_state_diff_vs_sequential": batch_err, "mean_iterations_per_frame": float(np.mean(iters)),
                          "gpu_ms_per_frame_wall_median": float(np.median(lat)), "gpu_ms_per_frame_wall_p99": float(np.percentile(lat, 99)),
                          "gpu_frames_per_s": 1e3 / float(np.mean(lat)),
                          "cpu_oracle_ms_per_frame": float(np.median(cpu)), "pose_abs_diff_vs_oracle_first_10": err}))


if __name__ == "__main__":
    main()
