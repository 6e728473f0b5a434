"""Population sharding over the GPUs of one box.

Organisms are independent in evaluation, so each rank evaluates its own contiguous shard
with replicated tables and NO collective.  Ranking is global: the per-organism objective
records (16 B each) are all-gathered over NCCL/NVLink, every rank computes the maximin rows
of its own shard against all columns, the fits are all-gathered, and every rank derives
the identical stable order locally.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous balanced shard [lo, hi) of n items for `rank` of `world`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class DistributedRanker(object):
    """Global maximin + arena order over shards.  `engine` needs maximin(obj, constr,
    row_range=) and arena_order(fit) (FitnessEngine, or a stand-in in CPU tests)."""

    def __init__(self, engine, group=None):
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank_id = dist.get_rank(group) if dist.is_initialized() else 0

    def gather_objectives(self, local_obj):
        """[P_local, 2] per rank (equal P_local) -> [world * P_local, 2], rank-major."""
        if self.world == 1:
            return local_obj
        out = local_obj.new_empty((self.world * local_obj.shape[0], local_obj.shape[1]))
        dist.all_gather_into_tensor(out, local_obj.contiguous(), group=self.group)
        return out

    def rank(self, local_obj, coverage_constr):
        """Returns (glad_obj [N,2], fit [N], order [N], nondom [N]) replicated on all ranks."""
        glad = self.gather_objectives(local_obj)
        n = glad.shape[0]
        lo, hi = shard_range(n, self.rank_id, self.world)
        fit_full = self.engine.maximin(glad, coverage_constr, row_range=(lo, hi))
        if self.world > 1:
            per = n // self.world
            assert per * self.world == n, "equal shards expected"
            fit = fit_full.new_empty(n)
            dist.all_gather_into_tensor(fit, fit_full[lo:hi].contiguous(), group=self.group)
        else:
            fit = fit_full
        order, nondom = self.engine.arena_order(fit)
        return glad, fit, order, nondom
