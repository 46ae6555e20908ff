#!/usr/bin/env python
"""Pool of per-pair contexts (debug flag 4) vs slot waves (flag 8) by batch size: where the wave path starts to win.
usage (GPU box): python tools/wave_crossover.py"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
phovo = importlib.import_module("photoconsistency-visual-odometry_b200")
K = phovo.synth.K_FRAME_ALIGNMENT
g0, d0, g1, _ = phovo.synth.render_batch_torch(513, 480, 640, K, device="cuda")
for key, name, mode in (("photometric_plus_depth", "config_4_level_optimization_analytic", phovo.MODE_BIOBJECTIVE), ("ceres_mode", "config_5_level_optimization_ceres", None)):
    odo = phovo.CPhotoconsistencyOdometryCuda(device=0)
    odo.SetConfig(phovo.configs.to_config(name, phovo.capi, mode=mode) if mode is not None else phovo.configs.to_config(name, phovo.capi))
    odo.SetIntrinsicMatrix(K)
    for P in (8, 16, 32, 64, 96, 128, 192, 256, 512):
        row = {"solver": key, "pairs": P}
        for flag, label in ((4, "pool_ms"), (8, "waves_ms")):
            odo.BatchSetDebugFlags(flag)
            kw = {"depth1": d0[1:P + 1].contiguous()} if mode is not None else {}
            odo.BatchAlign(g0[:P], d0[:P], g1[:P], **kw)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            odo.BatchAlign(g0[:P], d0[:P], g1[:P], **kw)
            row[label] = round(1e3 * (time.perf_counter() - t0), 3)
        print(json.dumps(row))
    odo.close()
