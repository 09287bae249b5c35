"""world_size-2 `gloo` tests of the multi-GPU host logic on CPU: descriptor broadcast, shard bounds,
seed-by-global-index generation, max/sum reductions and the result gather.  The per-instance solves in here are done by the
CPU oracle (this is a test of the sharding layer, not of the CUDA path)."""
import os
import socket
import sys

import numpy as np
import pytest

from piqp_b200.distributed import ShardedBatch, shard_bounds

HERE = os.path.dirname(os.path.abspath(__file__))


def test_shard_bounds_cover_the_batch_exactly():
    for B in (0, 1, 2, 7, 256, 1024, 1025):
        for W in (1, 2, 3, 4, 8):
            spans = [shard_bounds(B, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[r][1] == spans[r + 1][0] for r in range(W - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _solve_shard(seeds):
    """rank-local work: generate the instance of each global seed and solve it with the CPU oracle"""
    sys.path.insert(0, os.path.join(HERE, ".."))
    from oracle import pyoracle as oracle
    from piqp_b200.synth import dense_strongly_convex_qp
    rows = []
    for seed in seeds:
        q = dense_strongly_convex_qp(12, 4, 6, seed=seed)
        s = oracle.DenseSolver(); s.setup(q["P"], q["c"], q["A"], q["b"], q["G"], q["h_l"], q["h_u"], q["x_l"], q["x_u"])
        status = s.solve(); r = s.result()
        rows.append(np.concatenate([[status, r.info.iter], r.x]))
    return np.array(rows).reshape(len(seeds), -1)


def _worker(rank, world, port, global_batch, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # only rank 0 knows the real descriptor; the others pass garbage and must receive rank 0's
        sb = ShardedBatch(global_batch if rank == 0 else 999, seed0=42 if rank == 0 else 7, dist=dist)
        assert (sb.global_batch, sb.seed0, sb.world, sb.rank) == (global_batch, 42, world, rank)
        assert (sb.lo, sb.hi) == shard_bounds(global_batch, world, rank)
        local = _solve_shard(sb.local_seeds())
        sb.barrier()
        tmax = sb.reduce_max(10.0 + rank)                       # "device ms" of this rank
        tot = sb.reduce_sum([sb.local_batch, float(local[:, 1].sum())])
        full = sb.gather_rows(local)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), tmax=tmax, tot=np.array(tot), full=full, lo=sb.lo, hi=sb.hi)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("global_batch", [5, 6])
def test_two_rank_gloo_sharding_matches_single_process(tmp_path, oracle, global_batch):
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, global_batch, str(tmp_path)), nprocs=world, join=True)
    single = ShardedBatch(global_batch, seed0=42)
    assert (single.lo, single.hi, single.world) == (0, global_batch, 1)
    ref = _solve_shard(single.local_seeds())
    assert np.all(ref[:, 0] == 1)
    outs = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    for r, o in enumerate(outs):
        assert float(o["tmax"]) == 11.0                                         # max over ranks
        assert o["tot"][0] == global_batch and o["tot"][1] == ref[:, 1].sum()   # sum over ranks
        assert np.array_equal(o["full"], ref)         # union of shards == single-process batch, bit for bit
    assert outs[0]["hi"] == outs[1]["lo"] and outs[1]["hi"] == global_batch


def test_lpt_assignment_is_balanced_and_deterministic():
    from piqp_b200.distributed import lpt_assign
    rng = np.random.default_rng(0)
    costs = np.concatenate([rng.uniform(1, 10, 60), [200.0, 150.0, 90.0]])
    for world in (1, 2, 4, 8):
        owner = lpt_assign(costs, world)
        assert owner == lpt_assign(list(costs), world) and set(owner) <= set(range(world))
        load = np.bincount(owner, weights=costs, minlength=world)
        assert load.max() <= max(costs.max(), 4.0 / 3.0 * costs.sum() / world + 1e-9)      # LPT bound: 4/3 OPT, OPT >= max(largest item, mean load)
    with pytest.raises(ValueError):
        lpt_assign([1.0], 0)


def _suite_worker(rank, world, port, out_dir):
    """config-5 style heterogeneous suite: every rank solves the QPs LPT assigns to it (oracle on the CPU here, the CUDA
    solver on the GPU box: tools/mm_suite.py), one all-reduce collects the totals"""
    import torch.distributed as dist
    from piqp_b200.distributed import lpt_assign
    from oracle import pyoracle
    from helpers import load_mm_small
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        probs, gold = load_mm_small()
        names = sorted(probs)[:16]
        owner = lpt_assign([gold[nm]["n"] + gold[nm]["p"] + gold[nm]["m"] for nm in names], world)
        solved = iters = 0
        for nm, r in zip(names, owner):
            if r != rank:
                continue
            s = pyoracle.SparseSolver(pyoracle.default_settings(kkt_solver="sparse_ldlt")); s.setup(*probs[nm])
            solved += int(s.solve() == 1); iters += s.result().info.iter
        sb = ShardedBatch(len(names), dist=dist)
        tot = sb.reduce_sum([solved, iters, sum(1 for r in owner if r == rank)])
        np.savez(os.path.join(out_dir, "suite%d.npz" % rank), tot=np.array(tot))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_heterogeneous_suite(tmp_path, oracle):
    import torch.multiprocessing as mp
    from helpers import load_mm_small
    _, gold = load_mm_small()
    names = sorted(gold)[:16]
    port = _free_port()
    mp.spawn(_suite_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    outs = [np.load(os.path.join(str(tmp_path), "suite%d.npz" % r))["tot"] for r in range(2)]
    assert np.array_equal(outs[0], outs[1])
    assert outs[0][0] == 16 and outs[0][2] == 16 and outs[0][1] == sum(gold[nm]["iter"] for nm in names)
