"""Multi-GPU use of the path, one process per GPU over torch.distributed (BASELINE configs 4 and 5).

Two shardings, both cut where the work divides naturally:

* `shard_pairs` -- a batch of independent frame pairs is partitioned by pair; ranks never
  communicate on the data path, only the final poses are gathered.
* `RowShardedAlignment` -- ONE large pair: every rank holds the level pyramids, evaluates the
  normal equations (AN:538-540) over its band of source rows and the 27 sums (+ cost, count; 32
  doubles) are all-reduced every Gauss-Newton iteration, after which every rank takes the same
  step redundantly.  Three exchanges:
     "peer"      fused: the reduction kernel stores its sums into every peer's memory over NVLink
                 and sums the slots in rank order (one kernel, no collective launch; needs CUDA IPC)
     "allgather" NCCL all_gather of 32 doubles + sum in rank order (bitwise equal on all ranks)
     "allreduce" NCCL all_reduce(SUM) (order chosen by NCCL)
  The host-side protocol (band split, lock-step termination) is independent of the device and is
  exercised on CPU with gloo in tests/test_sharding_gloo.py.
"""
import numpy as np


def shard_pairs(num_pairs, rank, world):
    """Contiguous block of pairs owned by `rank`: [begin, end)."""
    base, extra = divmod(num_pairs, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def row_band(rows, rank, world):
    """Source rows [begin, end) of a level that `rank` accumulates -- the split libphovo_b200 uses
    (phovo_api.cu level_params: rows * rank / world)."""
    return (rows * rank) // world, (rows * (rank + 1)) // world


def exchange_fixed_order(buf, group=None):
    """All-reduce of a small tensor with a FIXED summation order (rank 0 + rank 1 + ...): every
    rank ends with bitwise the same values, run to run.  Works on any backend (NCCL, gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf.contiguous(), group=group)
    total = gathered[0].clone()
    for r in range(1, world):
        total += gathered[r]
    buf.copy_(total)
    return buf


class _DeviceBuffer:
    """torch view of the library's 32-double reduction buffer (no copy)."""

    def __init__(self, ptr, device_index):
        self.__cuda_array_interface__ = {"shape": (32,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}
        self.device_index = device_index


class RowShardedAlignment:
    """Optimize() of one pair with the per-pixel reduction split by rows across the ranks of
    `group`.  `odo` is this rank's CPhotoconsistencyOdometryCuda with config, intrinsics and the
    SAME frames set on every rank."""

    def __init__(self, odo, rank, world, device_index, group=None, exchange="peer", poll=4):
        import torch
        import torch.distributed as dist
        self.odo, self.rank, self.world, self.group, self.exchange, self.poll = odo, rank, world, group, exchange, poll
        # The collectives of the "allreduce" / "allgather" exchanges read and write the library's buffer on
        # torch's current stream (NCCL orders itself after that stream): the context must run on the SAME
        # stream, or ShardPartial / the collective / ShardStep are unordered with respect to each other.
        odo.SetStream(torch.cuda.current_stream(torch.device("cuda", device_index)).cuda_stream)
        odo.ShardConfigure(rank, world)
        self.buf = torch.as_tensor(_DeviceBuffer(odo.ShardBuffer(), device_index), device=torch.device("cuda", device_index))
        if exchange == "peer" and world > 1:
            mine = odo.ShardPeerExport()
            handles = [None] * world
            dist.all_gather_object(handles, mine, group=group)
            for r, h in enumerate(handles):
                odo.ShardPeerImport(r, h)
            dist.barrier(group=group)

    def optimize_fused(self, initial_state=None, min_shard_pixels=262144):
        """Optimize() with the sharded loop INSIDE the persistent kernel (exchange "peer" only): one cooperative launch
        per active level on every rank, the 29 sums exchanged in the kernel over NVLink peer memory, no host in the
        loop.  Levels below `min_shard_pixels` run unsharded on every rank (an exchange costs more than it saves
        there).  Returns (state, executed iterations per level dict)."""
        if self.world > 1 and self.exchange != "peer":
            raise ValueError('the fused loop needs exchange="peer"')
        odo = self.odo
        odo.SetInitialStateVector(np.zeros(6) if initial_state is None else initial_state)
        odo.ShardOptimize(min_shard_pixels)
        executed = {}
        for e in odo.IterationStats():
            executed[e["level"]] = executed.get(e["level"], 0) + 1
        return odo.GetOptimalStateVector(), executed

    def optimize(self, initial_state=None):
        """Returns (state, executed iterations per level dict).  Lock step: every rank executes the
        same number of exchanges because every rank holds the same state."""
        import torch.distributed as dist
        odo = self.odo
        cfg = odo.GetConfig()
        odo.SetInitialStateVector(np.zeros(6) if initial_state is None else initial_state)
        odo.ShardBegin()
        executed = {}
        for level in range(cfg.num_levels - 1, -1, -1):          # AN:502-503
            M = cfg.max_num_iterations[level]
            if M <= 0:
                continue
            odo.ShardBeginLevel(level)
            it, done = 0, False
            while it < M and not done:
                chunk = min(self.poll, M - it)
                for k in range(chunk):
                    last = k == chunk - 1
                    if self.exchange == "peer" and self.world > 1:
                        odo.ShardPartialExchange()
                    else:
                        odo.ShardPartial()
                        if self.world > 1:
                            if self.exchange == "allreduce":
                                dist.all_reduce(self.buf, group=self.group)
                            else:
                                exchange_fixed_order(self.buf, self.group)
                    # iterations enqueued after convergence are no-ops on the device (pose->done)
                    done = odo.ShardStep(want_done=last)
                it += chunk
            executed[level] = None
        odo.ShardFinish()
        log = odo.IterationStats()
        for level in executed:
            executed[level] = sum(1 for e in log if e["level"] == level)
        return odo.GetOptimalStateVector(), executed
