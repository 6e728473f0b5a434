"""Population sharding over the GPUs of one box.

Organisms are independent in evaluation, so each rank evaluates its own contiguous shard
with replicated tables and NO collective.  Ranking is global: the per-organism objective
records (16 B each) are all-gathered over NCCL/NVLink, every rank computes the maximin rows
of its own shard against all columns, the fits are all-gathered, and every rank derives
the identical stable order locally.
"""
import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(device=0):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (Linux sysfs), so that the
    pinned staging buffers allocated afterwards are node-local: with one process per GPU on a
    two-socket host, host->device copies otherwise cross the socket interconnect.  Returns the node
    id, or None when the topology is not exposed (then nothing is changed)."""
    import os
    import torch
    try:
        p = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return node if cpus else None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:                 # no sysfs topology, restricted cpuset, old torch: leave the process alone
        return None


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


def _all_gather(t, group=None):
    """all_gather_into_tensor for device tensors: NCCL directly, other backends (gloo in the
    single-GPU / CPU tests) through host memory."""
    world = dist.get_world_size(group)
    out = t.new_empty((world * t.shape[0],) + tuple(t.shape[1:]))
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
        return out
    host = t.contiguous().cpu()
    buf = host.new_empty(out.shape)
    dist.all_gather_into_tensor(buf, host, group=group)
    return buf.to(t.device)


class DistributedEvolution(object):
    """One generation of geneticalgorithm.py:55-76 with the CHILDREN sharded over the ranks.

    Parents (DNA, objectives, fitness) are replicated.  Every rank resamples the same parents
    (lowVarSample, tiny), builds and evaluates only its own contiguous range of child pairs, then
    the children's DNA and objective records are all-gathered (NCCL over NVLink) and every rank
    derives the identical maximin fitness, stable order and survivors.  Philox streams are keyed
    by GLOBAL organism ids (`vary_org_offset`), so the result is bit-identical to the single-GPU
    generation with the same seed, for any number of ranks."""

    def __init__(self, engine, gen_params, group=None):
        self.eng = engine
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank_id = dist.get_rank(group) if dist.is_initialized() else 0
        self.gamma = gen_params['gamma']
        self.c0 = -gen_params['coverage_constr_0']
        self.cf = -gen_params['coverage_constr_f']
        self.aging = gen_params['coverage_aging']

    def constraint(self, gen):
        alpha = min(1, gen / self.aging)                       # geneticalgorithm.py:98-100
        return (1 - alpha) * self.c0 + alpha * self.cf

    def step(self, par_dna, par_obj, par_fit, seed, gen):
        """-> (new parents dna [P,A,L], obj [P,2], fit [P]) replicated on all ranks."""
        eng, P = self.eng, par_dna.shape[0]
        half = P // 2
        if half % self.world:
            raise ValueError("pairs (%d) must divide evenly over %d ranks" % (half, self.world))
        m = half // self.world
        k0 = self.rank_id * m
        r0 = float(eng.lib.gcp_low_var_r0(int(seed), int(gen), P))
        strong = eng.low_var_select(par_fit, self.gamma, r0)
        local = torch.cat([strong[k0:k0 + m], strong[half + k0:half + k0 + m]]).contiguous()
        eng.set_option("vary_org_offset", 2 * k0)
        try:
            kids = eng.crossover(par_dna, local, seed, gen)    # [2m, A, L]: children 2k0 .. 2(k0+m)-1
            eng.mutate(kids, seed, gen, salt=0)
        finally:
            eng.set_option("vary_org_offset", 0)
        kid_obj = eng.evaluate(kids, detail=False).obj
        if self.world > 1:
            if eng.pw.N <= 32767:
                # waypoint ids (and the -1 padding) fit 16 bits: half the bytes on the wire for the
                # one large exchange of a generation (2.5 GB of int32 at 8 x 65,536 organisms)
                wire = kids.to(torch.int16).contiguous().view(torch.uint8)      # neither NCCL nor gloo moves int16
                kids_all = _all_gather(wire, self.group).view(torch.int16).to(torch.int32)
            else:
                kids_all = _all_gather(kids, self.group)
            kid_obj = _all_gather(kid_obj, self.group)
        else:
            kids_all = kids
        glad_obj = torch.cat([par_obj, kid_obj])
        fit = eng.maximin(glad_obj, self.constraint(gen))
        order, _ = eng.arena_order(fit)
        return eng.gather_survivors(order, par_dna, kids_all, glad_obj, fit)
