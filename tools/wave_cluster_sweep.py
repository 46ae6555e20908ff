#!/usr/bin/env python
"""Slot waves with 1 / 2 / 4 / 8 CTAs per pair (PHOVO_WAVE_CLUSTER) and the pool, by batch size.
usage (GPU box): python tools/wave_cluster_sweep.py"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
phovo = importlib.import_module("photoconsistency-visual-odometry_b200")
K = phovo.synth.K_FRAME_ALIGNMENT
g0, d0, g1, _ = phovo.synth.render_batch_torch(1185, 480, 640, K, device="cuda")
for key, name, mode in (("photometric_plus_depth", "config_4_level_optimization_analytic", phovo.MODE_BIOBJECTIVE), ("ceres_mode", "config_5_level_optimization_ceres", None)):
    odo = phovo.CPhotoconsistencyOdometryCuda(device=0)
    odo.SetConfig(phovo.configs.to_config(name, phovo.capi, mode=mode) if mode is not None else phovo.configs.to_config(name, phovo.capi))
    odo.SetIntrinsicMatrix(K)
    for P in [int(a) for a in sys.argv[1:]] or (8, 16, 32, 64, 128, 148, 296, 592, 1184):
        row = {"solver": key, "pairs": P}
        kw = {"depth1": d0[1:P + 1].contiguous()} if mode is not None else {}
        ref = None
        for label, flag, cl in (("pool_ms", 4, None), ("auto_ms", 0, None), ("c1_ms", 8, 1), ("c2_ms", 8, 2), ("c4_ms", 8, 4), ("c8_ms", 8, 8)):
            if label == "pool_ms" and P > 600:
                continue
            odo.BatchSetDebugFlags(flag)
            if cl is None: os.environ.pop("PHOVO_WAVE_CLUSTER", None)
            else: os.environ["PHOVO_WAVE_CLUSTER"] = str(cl)
            odo.BatchAlign(g0[:P], d0[:P], g1[:P], **kw)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            st, it = odo.BatchAlign(g0[:P], d0[:P], g1[:P], **kw)
            row[label] = round(1e3 * (time.perf_counter() - t0), 3)
            if ref is None: ref = (st, it)
            elif label != "c1_ms" or True:
                row[label.replace("_ms", "_iters_equal")] = bool(np.array_equal(it, ref[1]))
                row[label.replace("_ms", "_max_diff")] = float(np.max(np.abs(st - ref[0])))
        row["auto_path"] = None
        odo.BatchSetDebugFlags(0); os.environ.pop("PHOVO_WAVE_CLUSTER", None)
        odo.BatchAlign(g0[:P], d0[:P], g1[:P], **kw); row["auto_path"] = odo.BatchLastPath()
        print(json.dumps(row))
    odo.close()
