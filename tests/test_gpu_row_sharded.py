"""BASELINE configs[4] / SURVEY 8(e): the row-sharded Optimize() of one large pair.

  * one GPU: phovo_shard_optimize with a world of one is the persistent cooperative Optimize();
  * >= 2 GPUs visible: tests/row_sharded_worker.py under torch.distributed (NCCL, one process per GPU) -- the host-driven
    loop with the NCCL all-reduce / all-gather / fused peer-store exchanges and the loop inside the persistent kernel
    (in-kernel NVLink exchange) all reproduce the single-GPU result, with equal iteration counts, bitwise identical on
    every rank;
  * the full 7680 x 4320 / config_6_level pair on the GPU against the CPU oracle.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import REL_NORMAL_EQ, assert_pose_close, g_rel_err, h_rel_err
from test_gpu_parity import conv_cfg, make_odo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_optimize_world_of_one_is_optimize(phovo):
    K = phovo.synth.K_FRAME_ALIGNMENT
    g0, d0, g1, _ = phovo.synth.make_pair(480, 640, K=K, seed=9)
    cfg = phovo.configs.to_config("config_5_level_optimization_analytic", phovo.capi)
    a, b = make_odo(phovo, cfg, K), make_odo(phovo, cfg, K)
    for odo in (a, b):
        odo.SetSourceFrame(g0, d0); odo.SetTargetFrame(g1); odo.SetInitialStateVector(np.zeros(6))
    a.Optimize()
    b.ShardConfigure(0, 1)
    b.ShardOptimize(0)
    assert np.array_equal(a.GetOptimalStateVector(), b.GetOptimalStateVector())
    assert len(a.IterationStats()) == len(b.IterationStats()) > 0


def test_row_sharded_matches_single_gpu_under_torch_distributed():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "row_sharded_worker.py"), "1080", "1920"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["world"] == world and line["iterations"] > 0
    for k in ("host_loop_allreduce", "host_loop_allgather", "host_loop_peer", "fused_all_levels", "fused_large_levels"):
        assert line[k] < 1e-10, (k, line[k])


def test_full_size_8k_pair_matches_oracle(phovo, oracle):
    """7680 x 4320, config_6_level_optimization_analytic (levels 2-5 active, 2 Mpx at the largest): GPU vs CPU oracle,
    per executed iteration the normal equations, equal iteration counts, final pose."""
    import torch
    K = phovo.synth.K_8K
    dev = torch.device("cuda", 0)
    g0, d0, g1, _ = phovo.synth.render_batch_torch(1, 4320, 7680, K, dev, seed0=11, chunk=1, xis=phovo.synth.XI_CONFIG1[None])
    g0, d0, g1 = g0[0].cpu().numpy(), d0[0].cpu().numpy().astype(np.float64), g1[0].cpu().numpy()
    torch.cuda.empty_cache()
    cfg = phovo.configs.to_config("config_6_level_optimization_analytic", phovo.capi)
    odo = make_odo(phovo, cfg, K)
    odo.SetSourceFrame(g0, d0); odo.SetTargetFrame(g1); odo.SetInitialStateVector(np.zeros(6))
    odo.Optimize()
    log = odo.IterationStats()
    o = oracle.Oracle(conv_cfg(oracle, cfg), K)
    o.set_source(g0, d0); o.set_target(g1); o.set_initial_state(np.zeros(6))
    o.optimize()
    olog = o.iter_stats()
    assert len(log) == len(olog) > 0
    for a, b in zip(log, olog):
        assert (a["level"], a["iteration"], a["num_valid"]) == (b["level"], b["iteration"], b["num_valid"])
        assert h_rel_err(a["H"], b["H"]) < REL_NORMAL_EQ and g_rel_err(a["g"], b["g"]) < REL_NORMAL_EQ
        assert h_rel_err(a["H"], b["H"]) < 1e-10 and g_rel_err(a["g"], b["g"]) < 1e-9
    assert_pose_close(odo.GetOptimalStateVector(), o.state(), "8K pair")
    assert np.max(np.abs(odo.GetOptimalStateVector() - o.state())) < 1e-10
    # BASELINE configs[4] at its stated size against the REFERENCE'S OWN header (oracle/_ref, when the prebuilt library
    # travelled): iteration counts, normal equations of every executed iteration, final state
    import ref_py
    if ref_py.available():
        import tempfile
        ref = ref_py.Reference(phovo.configs.write_yaml("config_6_level_optimization_analytic", tempfile.mkdtemp()), K)
        rs, _, riters = ref.align(g0, d0, g1)
        assert len(riters) == len(log)
        idx = [(a, b) for a in range(6) for b in range(a, 6)]
        for e, it in zip(log, riters):
            Hp = np.array([it["H"][a, b] for a, b in idx])
            assert h_rel_err(e["H"], Hp) < 1e-10 and g_rel_err(e["g"], it["g"]) < 1e-9
        assert np.max(np.abs(odo.GetOptimalStateVector() - rs)) < 1e-10
