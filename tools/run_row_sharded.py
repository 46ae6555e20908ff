#!/usr/bin/env python
"""BASELINE configs[4]: one synthetic 7680x4320 pair, config_6_level_optimization_analytic, the per-pixel
reduction row-sharded across the ranks with a 27(+5)-value exchange per Gauss-Newton iteration.

Launch (one process per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/run_row_sharded.py
Every rank renders the same pair, runs the single-GPU Optimize() as the reference result, then the
sharded Optimize with each exchange ("allreduce" = NCCL sum, "allgather" = NCCL all_gather + fixed-order
sum, "peer" = fused NVLink peer-store kernel, "none" = no exchange, timing floor only) and checks state and
iteration counts.  Rank 0 prints one JSON line."""
import argparse, importlib, json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4320)
    ap.add_argument("--cols", type=int, default=7680)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    phovo = importlib.import_module("photoconsistency-visual-odometry_b200")
    phovo.build()
    from importlib import import_module
    sharded = import_module("photoconsistency-visual-odometry_b200.sharded")
    s = args.cols / 7680.
    K = np.array([[6300. * s, 0, (args.cols - 1) / 2.], [0, 6300. * s, (args.rows - 1) / 2.], [0, 0, 1.]])
    g0, d0, g1, xis = phovo.synth.render_batch_torch(1, args.rows, args.cols, K, dev, seed0=0, chunk=1)
    g0, d0, g1 = g0[0].contiguous(), d0[0].contiguous(), g1[0].contiguous()
    cfg = phovo.configs.to_config("config_6_level_optimization_analytic", phovo.capi)
    stream = torch.cuda.current_stream(dev)

    def new_odo():
        odo = phovo.CPhotoconsistencyOdometryCuda(device=local)
        odo.SetConfig(cfg); odo.SetIntrinsicMatrix(K); odo.SetStream(stream.cuda_stream)
        odo.SetSourceFrame(g0, d0); odo.SetTargetFrame(g1)
        return odo

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # single-GPU result (CUDA-graph Optimize) on every rank
    ref = new_odo()
    times = []
    for _ in range(args.reps):
        ref.SetInitialStateVector(np.zeros(6))
        sync(); t0 = time.perf_counter()
        ref.Optimize()
        times.append((time.perf_counter() - t0) * 1e3)
    s_ref, log_ref = ref.GetOptimalStateVector(), ref.IterationStats()
    setup_ms, opt_ms = ref.Timings()
    n_iter = len(log_ref)
    out = {"workload": "single %dx%d pair, config_6_level_optimization_analytic, rows sharded x%d" % (args.cols, args.rows, world),
           "world": world, "iterations": n_iter, "iterations_per_level": {str(l): sum(1 for e in log_ref if e["level"] == l) for l in range(cfg.num_levels) if cfg.max_num_iterations[l] > 0},
           "single_gpu_optimize_ms_device": opt_ms, "single_gpu_optimize_ms_wall_median": float(np.median(times)), "single_gpu_setup_ms_device": setup_ms,
           "state": s_ref.tolist(), "exchanges": {}}
    for exchange in (["none", "allreduce", "allgather", "peer"] if world > 1 else ["none"]):
        odo = new_odo()
        ra = sharded.RowShardedAlignment(odo, rank, world, local, exchange=exchange if exchange != "none" else "allgather")
        if exchange == "none":
            ra.world = 1          # timing floor: local partials only, no collective (result is NOT the alignment)
        ws = []
        for _ in range(args.reps):
            sync(); t0 = time.perf_counter()
            st, executed = ra.optimize()
            torch.cuda.synchronize(dev)
            ws.append((time.perf_counter() - t0) * 1e3)
        t = torch.tensor([float(np.median(ws))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        iters = sum(executed.values())
        rec = {"optimize_ms_wall_median_max_over_ranks": float(t.item()), "iterations": iters,
               "ms_per_iteration": float(t.item()) / max(iters, 1)}
        if exchange != "none":
            err = float(np.max(np.abs(st - s_ref)))
            e = torch.tensor([err], dtype=torch.float64, device=dev)
            gathered = torch.zeros((world, 6), dtype=torch.float64, device=dev)
            mine = torch.tensor(st, dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(e, op=dist.ReduceOp.MAX)
                dist.all_gather_into_tensor(gathered, mine)
            else:
                gathered[0] = mine
            rec.update({"max_abs_state_diff_vs_single_gpu": float(e.item()), "iterations_equal": iters == n_iter,
                        "ranks_bitwise_identical": bool((gathered == gathered[0]).all().item())})
            assert iters == n_iter and float(e.item()) < 1e-9, (exchange, iters, n_iter, float(e.item()))
        out["exchanges"][exchange] = rec
    if world > 1:
        base = out["exchanges"]["none"]["ms_per_iteration"]
        for k in ("allreduce", "allgather", "peer"):
            out["exchanges"][k]["exchange_us_per_iteration_over_floor"] = (out["exchanges"][k]["ms_per_iteration"] - base) * 1e3
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
