"""Batch sharding for the multi-GPU mode: independent QP instances are partitioned over ranks, one
process per GPU.  The only collective of the data path is ONE broadcast of the problem descriptor at setup; the
interior-point loop itself never communicates.  Timing / throughput reductions (max over ranks of the device time, sum of
the instances solved) and the optional gather of per-instance results are bookkeeping around the path.

Works with any torch.distributed backend: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
import numpy as np


def shard_bounds(global_batch, world, rank):
    """contiguous shard [lo, hi) of rank `rank`; the first (global_batch % world) ranks own one extra instance"""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world: %r/%r" % (rank, world))
    if global_batch < 0:
        raise ValueError("negative batch")
    base, rem = divmod(int(global_batch), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def lpt_assign(costs, world):
    """Heterogeneous suites (BASELINE config 5): longest-processing-time-first assignment of work items to ranks.
    Items are taken by decreasing cost (ties: by index) and given to the currently least-loaded rank (ties: lowest rank), so
    every rank computes the same assignment from the same cost vector without communicating.  Returns owner[item]."""
    if world <= 0:
        raise ValueError("bad world size")
    costs = [float(c) for c in costs]
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    owner = [0] * len(costs)
    for i in order:
        r = min(range(world), key=lambda q: (load[q], q))
        owner[i] = r
        load[r] += costs[i]
    return owner


class ShardedBatch:
    """Descriptor of a global batch of independent QPs split over the ranks of a process group.

    Rank 0's (global_batch, seed0) win: they are broadcast once (the setup broadcast of the north star); instance g of the
    global batch is generated from seed0 + g on whichever rank owns it, so the union of all shards is independent of the
    number of ranks (the property the world_size-2 tests check).
    """

    def __init__(self, global_batch, seed0=42, dist=None, device="cpu"):
        import torch
        self.dist = dist if (dist is not None and dist.is_initialized()) else None
        self.device = device
        self.rank = self.dist.get_rank() if self.dist else 0
        self.world = self.dist.get_world_size() if self.dist else 1
        desc = torch.tensor([int(global_batch), int(seed0)], dtype=torch.int64, device=device)
        if self.dist:
            self.dist.broadcast(desc, src=0)
        self.global_batch, self.seed0 = [int(v) for v in desc.tolist()]
        self.lo, self.hi = shard_bounds(self.global_batch, self.world, self.rank)

    @property
    def local_batch(self):
        return self.hi - self.lo

    def local_seeds(self):
        return [self.seed0 + g for g in range(self.lo, self.hi)]

    def barrier(self):
        if self.dist:
            self.dist.barrier()

    def reduce_max(self, value):
        """max over ranks (device time of the timed region)"""
        import torch
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device)
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(self, values):
        """element-wise sum over ranks (instances solved, flops, iterations)"""
        import torch
        t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=self.device)
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def gather_rows(self, local):
        """all ranks receive the (global_batch, k) array whose rows [lo, hi) are this rank's `local` (local_batch, k)"""
        import torch
        local = np.ascontiguousarray(np.asarray(local, dtype=np.float64).reshape(self.local_batch, -1))
        if not self.dist:
            return local
        k = local.shape[1]
        cap = max(shard_bounds(self.global_batch, self.world, r)[1] - shard_bounds(self.global_batch, self.world, r)[0] for r in range(self.world))
        pad = torch.zeros((cap, k), dtype=torch.float64, device=self.device)
        pad[: self.local_batch] = torch.from_numpy(local).to(self.device)
        out = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(out, pad)
        rows = []
        for r in range(self.world):
            lo, hi = shard_bounds(self.global_batch, self.world, r)
            rows.append(out[r][: hi - lo].cpu().numpy())
        return np.concatenate(rows, axis=0)
