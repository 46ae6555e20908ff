"""Worker of tests/test_gpu_row_sharded.py: one process per GPU (torch.distributed, NCCL).  Every rank renders the same
pair, computes the single-GPU result and then runs the row-sharded Optimize() in every flavour; asserts on every rank
and prints one JSON line on rank 0."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rows, cols = int(sys.argv[1]), int(sys.argv[2])
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    phovo = importlib.import_module("photoconsistency-visual-odometry_b200")
    phovo.build()
    s = cols / 7680.
    K = np.array([[6300. * s, 0, (cols - 1) / 2.], [0, 6300. * s, (rows - 1) / 2.], [0, 0, 1.]])
    g0, d0, g1, _ = phovo.synth.render_batch_torch(1, rows, cols, K, dev, seed0=3, chunk=1, xis=phovo.synth.XI_CONFIG1[None])
    g0, d0, g1 = g0[0].contiguous(), d0[0].contiguous(), g1[0].contiguous()
    cfg = phovo.configs.to_config("config_6_level_optimization_analytic", phovo.capi)

    def new_odo():
        odo = phovo.CPhotoconsistencyOdometryCuda(device=local)
        odo.SetConfig(cfg); odo.SetIntrinsicMatrix(K)
        odo.SetSourceFrame(g0, d0); odo.SetTargetFrame(g1)        # device tensors: the wrapper follows torch's current stream
        return odo

    ref = new_odo()
    ref.SetInitialStateVector(np.zeros(6)); ref.Optimize()
    s_ref, log_ref = ref.GetOptimalStateVector(), ref.IterationStats()
    per_level = {}
    for e in log_ref:
        per_level[e["level"]] = per_level.get(e["level"], 0) + 1
    out = {"world": world, "iterations": len(log_ref)}
    # host-driven loop, every exchange -- the class binds the context to torch's stream itself (no SetStream here)
    for exchange in ("allreduce", "allgather", "peer"):
        ra = phovo.sharded.RowShardedAlignment(new_odo(), rank, world, local, exchange=exchange)
        st, executed = ra.optimize()
        err = float(np.max(np.abs(st - s_ref)))
        assert executed == per_level, (exchange, executed, per_level)
        assert err < 1e-10, (exchange, err)
        out["host_loop_" + exchange] = err
    # the loop inside the persistent kernel: every level sharded, and the policy that leaves small levels unsharded
    ra = phovo.sharded.RowShardedAlignment(new_odo(), rank, world, local, exchange="peer")
    for tag, min_px in (("all_levels", 0), ("large_levels", 65536)):
        for rep in range(3):                                            # repeated calls: epochs keep increasing
            st, executed = ra.optimize_fused(min_shard_pixels=min_px)
            err = float(np.max(np.abs(st - s_ref)))
            assert executed == per_level, (tag, executed, per_level)
            assert err < 1e-10, (tag, err)
        mine = torch.tensor(st, dtype=torch.float64, device=dev)
        gathered = torch.zeros((world, 6), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(gathered, mine)
        assert bool((gathered == gathered[0]).all().item()), "ranks disagree bitwise"
        out["fused_" + tag] = err
    torch.cuda.synchronize(dev)
    dist.barrier()
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
