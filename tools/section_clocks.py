#!/usr/bin/env python
"""Where the cycles of k_batch_level go: runs the bench workload on a library built with
-DPHOVO_SECTION_CLOCKS (tools/build_variant.sh secclk -- -DPHOVO_SECTION_CLOCKS) and prints, per kernel
flavour (small level: 3 CTAs x 160 threads per SM; large level: 1 CTA x 480), thread 0's clocks per section.
usage (GPU box): python tools/section_clocks.py build/variants/secclk.so [pairs]"""
import ctypes
import importlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib_path = os.path.join(ROOT, "photoconsistency-visual-odometry_b200", "libphovo_b200.so")
variant = sys.argv[1]
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
backup = lib_path + ".product"
shutil.copy(lib_path, backup)
shutil.copy(variant, lib_path)
try:
    import torch
    phovo = importlib.import_module("photoconsistency-visual-odometry_b200")
    cfg = phovo.configs.to_config("config_4_level_optimization_analytic", phovo.capi)
    K = phovo.synth.K_FRAME_ALIGNMENT
    g0, d0, g1, _ = phovo.synth.render_batch_torch(pairs, 480, 640, K, device="cuda", depth_dtype=torch.uint16)
    odo = phovo.CPhotoconsistencyOdometryCuda(device=0)
    odo.SetConfig(cfg)
    odo.SetIntrinsicMatrix(K)
    lib = phovo.capi.lib()
    lib.phovo_debug_section_clocks.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
    odo.BatchAlign(g0, d0, g1)            # warm-up
    torch.cuda.synchronize()
    lib.phovo_debug_section_clocks(None, 1)
    st, it = odo.BatchAlign(g0, d0, g1)
    torch.cuda.synchronize()
    out = (ctypes.c_ulonglong * 16)()
    assert lib.phovo_debug_section_clocks(out, 0) == 0
    names = ["tables+barrier", "phase A (thread 0)", "barrier after A", "phase B (thread 0)", "reduce+barrier", "solve (warp 0)", "exit", "kernel total"]
    res = {}
    for f, label in enumerate(["small level (3 x 160)", "large level (1 x 480)"]):
        v = [int(out[8 * f + k]) for k in range(8)]
        tot = v[7]
        inside = sum(v[:7])
        res[label] = {n: round(x / tot, 4) for n, x in zip(names[:7], v[:7])}
        res[label]["per-pair setup + waits"] = round((tot - inside) / tot, 4)
        res[label]["clocks_total"] = tot
    res["iterations"] = {"mean_per_pair": float(it.sum()) / pairs}
    print(json.dumps(res, indent=1))
finally:
    shutil.copy(backup, lib_path)
    os.remove(backup)
