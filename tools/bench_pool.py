#!/usr/bin/env python
"""Throughput of phovo_batch_align for what the shared-memory-resident kernels do not take (Ceres mode, photometric +
depth solver): the wave path (path 3), device-resident 640x480 pairs.  PHOVO_WAVE_TRACE=1 prints per-wave timings.
usage (GPU box): python tools/bench_pool.py <ignored> [pairs]     (first argument kept for old command lines)"""
import importlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(pool, pairs):
    os.environ["PHOVO_POOL_CONTEXTS"] = str(pool)
    import numpy as np
    import torch
    phovo = importlib.import_module("photoconsistency-visual-odometry_b200")
    K = phovo.synth.K_FRAME_ALIGNMENT
    g0, d0, g1, _ = phovo.synth.render_batch_torch(pairs + 1, 480, 640, K, device="cuda")
    out = {"pairs": pairs}
    for key, name, mode in (("biobjective", "config_4_level_optimization_analytic", phovo.MODE_BIOBJECTIVE),
                                         ("ceres", "config_5_level_optimization_ceres", None)):
        odo = phovo.CPhotoconsistencyOdometryCuda(device=0)
        odo.SetConfig(phovo.configs.to_config(name, phovo.capi, mode=mode) if mode is not None else phovo.configs.to_config(name, phovo.capi))
        odo.SetIntrinsicMatrix(K)
        kw = {"depth1": d0[1:pairs + 1].contiguous()} if mode is not None else {}
        a, b, c = g0[:pairs], d0[:pairs], g1[:pairs]
        odo.BatchAlign(a, b, c, **kw)   # warm-up: slots, arena, pinned buffers
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st, it = odo.BatchAlign(a, b, c, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert odo.BatchLastPath() == 3
        out[key] = {"pairs_per_s": round(pairs / dt, 1), "ms_per_pair": round(1e3 * dt / pairs, 4), "mean_iterations": float(it.sum()) / pairs,
                    "finite": bool(np.isfinite(st).all())}
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 2:
        one(int(sys.argv[1]), int(sys.argv[2]))
    else:
        one(4, int(sys.argv[1]) if len(sys.argv) > 1 else 2048)
