"""Host-side logic of the multi-GPU paths on CPU: world_size 2, gloo, 127.0.0.1.

The CUDA kernels cannot run here, so each rank's compute is done by the CPU oracle (test
infrastructure); what is under test is the product's sharding module: the pair partition and the
final pose gather (BASELINE config 4), and the row-band split + fixed-order exchange of the 32
normal-equation sums (BASELINE config 5), including that every rank ends with bitwise the same sums.
"""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "photoconsistency-visual-odometry_b200"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, what, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import oracle_py
    phovo = importlib.import_module(PKG)
    sharded = phovo.sharded
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        K = np.array([[131.25, 0., 79.5], [0., 131.25, 59.5], [0., 0., 1.]])
        cfg = phovo.default_config()
        cfg.num_levels = 3
        for l, m in enumerate((0, 6, 10)):
            cfg.max_num_iterations[l] = m
            cfg.min_gradient_norm[l] = 50.
        ocfg = oracle_py.Config.from_buffer_copy(bytes(cfg))
        if what == "pairs":
            P = 5
            g0, d0, g1, _ = phovo.synth.make_batch(P, 120, 160, K=K, seed0=700)
            b, e = sharded.shard_pairs(P, rank, world)
            st, it, _, _ = oracle_py.align_batch(ocfg, K, g0[b:e], d0[b:e], g1[b:e], num_threads=1, lean=True)
            mine = torch.zeros((3, 6), dtype=torch.float64)           # padded to the largest shard
            mine[: e - b] = torch.from_numpy(st)
            gathered = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(gathered, mine)                            # the final pose gather
            poses = np.concatenate([gathered[r][: np.diff(sharded.shard_pairs(P, r, world))[0]].numpy() for r in range(world)])
            full, _, _, _ = oracle_py.align_batch(ocfg, K, g0, d0, g1, num_threads=1, lean=True)
            q.put((rank, bool(np.array_equal(poses, full)), (b, e)))
        else:
            g0, d0, g1, _ = phovo.synth.make_pair(120, 160, K=K, seed=701)
            o = oracle_py.Oracle(ocfg, K)
            o.set_source(g0, d0)
            o.set_target(g1)
            ok, bands = True, []
            state = np.zeros(6)
            for level in (2, 1):
                rows, cols = o.level_image(0, level).shape
                r0, r1 = sharded.row_band(rows, rank, world)
                bands.append((level, r0, r1, rows))
                ev = o.eval(level, state, want_residuals=True, want_jacobian=True)
                J = ev["jacobian"].reshape(rows, cols, 6)[r0:r1].reshape(-1, 6)
                r = ev["residuals"].reshape(rows, cols)[r0:r1].reshape(-1)
                H = J.T @ J
                buf = torch.zeros(32, dtype=torch.float64)
                buf[:21] = torch.from_numpy(np.array([H[a, b] for a in range(6) for b in range(a, 6)]))
                buf[21:27] = torch.from_numpy(J.T @ r)
                before = buf.clone()
                sharded.exchange_fixed_order(buf)
                # fixed order: rank 0 + rank 1, bitwise identical on every rank
                both = [torch.zeros(32, dtype=torch.float64) for _ in range(world)]
                dist.all_gather(both, before)
                ok &= bool(torch.equal(buf, both[0] + both[1]))
                mirror = [torch.zeros(32, dtype=torch.float64) for _ in range(world)]
                dist.all_gather(mirror, buf)
                ok &= bool(torch.equal(mirror[0], mirror[1]))
                scale = np.max(np.abs(ev["H"]))
                ok &= bool(np.max(np.abs(buf[:21].numpy() - ev["H"])) < 1e-11 * scale)
                ok &= bool(np.max(np.abs(buf[21:27].numpy() - ev["g"])) < 1e-11 * np.max(np.abs(ev["g"])))
                state = state + 1e-3                                   # a second, non-trivial state
            q.put((rank, ok, bands))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("what", ["pairs", "rows"])
def test_two_rank_sharding_over_gloo(phovo, oracle, what):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, what, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in results), results
    if what == "pairs":
        spans = sorted(s for _, _, s in results)
        assert spans == [(0, 3), (3, 5)]
    else:
        by_rank = dict((r, b) for r, _, b in results)
        for (lvl, a0, a1, rows), (_, b0, b1, _) in zip(by_rank[0], by_rank[1]):
            assert a0 == 0 and a1 == b0 and b1 == rows                 # bands tile the level exactly


def test_partitions_are_exact():
    sharded = importlib.import_module(PKG).sharded
    for n in (1, 7, 4096):
        for world in (1, 2, 3, 8):
            spans = [sharded.shard_pairs(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1
    for rows in (1, 135, 1080):
        for world in (1, 2, 4, 8):
            bands = [sharded.row_band(rows, r, world) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == rows
            assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
