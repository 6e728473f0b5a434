"""CPU, world_size 2, gloo: the sharding / gathering logic of the multi-GPU ranking path.
The compute stand-in is the oracle (the product has no CPU compute path; on GPUs the same
DistributedRanker drives FitnessEngine over NCCL)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GP = dict(coverage_aging=80, coverage_constr_0=0.3, coverage_constr_f=0.95)


class OracleEngine(object):
    """maximin / arena_order with the oracle's arithmetic, honouring row_range."""

    def maximin(self, obj, constr, row_range=None):
        import fitness_oracle as fo
        o = obj.numpy()
        n = len(o)
        lo, hi = row_range or (0, n)
        sc = o.copy()
        sc[:, 1] = np.where(o[:, 0] > constr, o[:, 1] + 1000.0, o[:, 1])
        fit = np.full(n, np.nan)
        for i in range(lo, hi):
            m = np.minimum(sc[i, 0] - sc[:, 0], sc[i, 1] - sc[:, 1])
            m[i] = -np.inf
            fit[i] = m.max()
        return torch.from_numpy(fit)

    def arena_order(self, fit):
        order = np.argsort(fit.numpy(), kind='stable')
        return torch.from_numpy(order.astype(np.int32)), torch.from_numpy((fit.numpy() < 0))


def _worker(rank, world, port, objs, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from genetic_coverage_planning_b200.distributed import DistributedRanker, shard_range
    lo, hi = shard_range(len(objs), rank, world)
    ranker = DistributedRanker(OracleEngine())
    glad, fit, order, nondom = ranker.rank(torch.from_numpy(objs[lo:hi]), -0.3)
    q.put((rank, glad.numpy(), fit.numpy(), order.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range():
    from genetic_coverage_planning_b200.distributed import shard_range
    for n in (0, 1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_ranking_matches_single():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fitness_oracle as fo
    rs = np.random.RandomState(4)
    objs = np.stack([-rs.randint(0, 50, 96) / 50.0, rs.rand(96) * 2], 1)
    objs[rs.rand(96) < 0.4, 0] = 0.0
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, objs, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_fit = fo.rack_n_stack(objs, 0, GP)
    want_order = fo.arena_order(want_fit)
    for rank, glad, fit, order in res:
        assert np.array_equal(glad, objs)
        assert np.array_equal(fit, want_fit)
        assert np.array_equal(order, want_order)


def test_numa_binding_is_a_noop_without_topology():
    """bind_to_gpu_numa_node must never raise or change affinity when there is no GPU / sysfs node."""
    import os
    from genetic_coverage_planning_b200.distributed import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before
