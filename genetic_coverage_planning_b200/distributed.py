"""Population sharding over the GPUs of one box.

Organisms are independent in evaluation, so each rank evaluates its own contiguous shard
with replicated tables and NO collective.  Ranking is global: the per-organism objective
records (16 B each) are all-gathered over NCCL/NVLink and every rank derives the identical
maximin fitness and stable order locally.

Whole generations (`ShardedEvolution`, geneticalgorithm.py:55-84): ONE population of P organisms,
rank r holds parents [r P/W, (r+1) P/W) and builds, mutates and evaluates the children of the same
range.  Per generation the ranks exchange exactly one NCCL all-gather of the children's objective
records (16 B / organism) - chromosomes are never all-gathered: a crossover parent or a survivor
that lives on another GPU is read by the kernel that needs it straight from the owner's memory
over NVLink (shards are mapped into every rank with CUDA IPC, `gcp_shard_open`).  The same class
with W = 1 is the single-GPU generation loop, so there is one code path.
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


class ShardedEvolution(object):
    """One population of `pop_size` organisms sharded over the ranks of `group` (or one GPU).

    State per rank: parents shard (double-buffered) and children shard, each [P/W, A, L] genes
    (int16 when the waypoint ids fit and the fused evaluation kernel is in use, else int32), in
    memory the peers can map; objective records [P, 2] and fitness [P] of the parents replicated.

    One generation (`step`) on rank r - the reference lines are geneticalgorithm.py:59-84:
      1. lowVarSample over the replicated fitness (tiny, identical on every rank)         :61
      2. crossover of pairs [r P/2W, (r+1) P/2W): parents read through the shard table     :64-68
         (peer loads over NVLink), children written to the local children shard
      3. constructor variation + evaluation of the local children                          :156-171
      4. NCCL all-gather of the children's objective records (16 B / organism)             :72
      5. maximin over the 2P gladiators + stable order, replicated                         :86-123
      6. truncation: survivors [r P/W, (r+1) P/W) are fetched from wherever they live       :83-84
         (own or a peer's parents / children shard) into the second parents buffer
      7. rendezvous of the ranks on the device (`gcp_peer_barrier`, flags in peer memory)
    Philox streams are keyed by GLOBAL organism ids: the result is bit-identical for any W."""

    def __init__(self, engine, gen_params, pop_size, group=None, gene_dtype=None, barrier="peer",
                 shard=True):
        """shard=False: keep the whole population on this GPU even under torch.distributed (every
        rank then evolves the same population redundantly; same results)."""
        self.eng = engine
        self.group = group
        on = shard and dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if on else 1
        self.rank_id = dist.get_rank(group) if on else 0
        self.P = int(pop_size)
        if self.P % (2 * self.world):
            raise ValueError("population %d must be a multiple of 2 x %d ranks" % (self.P, self.world))
        if self.world > 16:
            raise ValueError("at most 16 shards (one box)")
        self.per = self.P // self.world
        self.gamma = gen_params['gamma']
        self.c0 = -gen_params['coverage_constr_0']
        self.cf = -gen_params['coverage_constr_f']
        self.aging = gen_params['coverage_aging']
        if gene_dtype is None:
            gene_dtype = torch.int16 if engine.gene16_available() else torch.int32
        self.gene_dtype = gene_dtype
        self.gb = 2 if gene_dtype == torch.int16 else 4
        A, L, dev = engine.A, engine.L, engine.device
        self.shape = (self.per, A, L)
        nbytes = self.per * A * L * self.gb
        if self.world == 1:
            bufs = [torch.empty(self.shape, dtype=gene_dtype, device=dev) for _ in range(3)]
            self._par = bufs[:2]
            self._kid = bufs[2]
            self._par_tab = [engine.shard_table([b.data_ptr()], self.per) for b in self._par]
            self._kid_tab = engine.shard_table([self._kid.data_ptr()], self.per)
            self._flag_tab = None
        else:
            own = [engine.shard_alloc(nbytes) for _ in range(3)] + [engine.shard_alloc(256)]
            self._own = own
            own[3].tensor.zero_()
            torch.cuda.synchronize(dev)
            mine = torch.tensor(list(b"".join(b.handle for b in own)), dtype=torch.uint8)
            handles = _all_gather(mine.to(dev).view(1, -1), group).cpu().numpy()      # [W, 4 * 64]
            ptrs = [[], [], [], []]
            for w in range(self.world):
                for k in range(4):
                    ptrs[k].append(own[k].ptr if w == self.rank_id else
                                   engine.shard_open(bytes(handles[w, 64 * k:64 * k + 64])))
            self._par = [own[0].view(gene_dtype, self.shape), own[1].view(gene_dtype, self.shape)]
            self._kid = own[2].view(gene_dtype, self.shape)
            self._par_tab = [engine.shard_table(ptrs[0], self.per), engine.shard_table(ptrs[1], self.per)]
            self._kid_tab = engine.shard_table(ptrs[2], self.per)
            self._flag_tab = engine.shard_table(ptrs[3], 1)
        self._cur = 0
        # gladiator records: [0, P) parents, [P, 2P) children; double-buffered so that the
        # truncation writes the next generation's parent records while it reads this one's
        self._glad = [torch.zeros((2 * self.P, 2), dtype=torch.float64, device=dev) for _ in range(2)]
        self._fit = [torch.zeros(self.P, dtype=torch.float64, device=dev) for _ in range(2)]
        self._kid_obj = torch.empty((self.per, 2), dtype=torch.float64, device=dev)
        self._kid_flags = torch.zeros((self.per, 4), dtype=torch.uint8, device=dev)
        self._timed_out = torch.zeros(1, dtype=torch.int32, device=dev)
        self._epoch = 0
        # two processes that share one GPU (single-GPU test boxes) time-slice it: a spinning
        # barrier kernel would only burn the peer's slice, so they meet through the collective
        self.barrier_mode = barrier if self.world > 1 else "none"
        if self.world > 1 and dist.get_backend(group) != "nccl":
            self.barrier_mode = "collective"
        self._sync_buf = torch.zeros(1, device=dev)
        self._sync()

    # ---- state -------------------------------------------------------------------------
    @property
    def parents(self):
        """this rank's parents shard [P/W, A, L] (gene dtype)."""
        return self._par[self._cur]

    @property
    def obj(self):
        return self._glad[self._cur][:self.P]

    @property
    def fit(self):
        return self._fit[self._cur]

    def constraint(self, gen):
        alpha = min(1, gen / self.aging)                       # geneticalgorithm.py:98-100
        return (1 - alpha) * self.c0 + alpha * self.cf

    def load(self, dna, gen=0):
        """Install parents.  dna: [P, A, L] (every rank passes the same full population) or this
        rank's [P/W, A, L] shard; any integer dtype.  Evaluates and ranks them."""
        eng = self.eng
        dna = torch.as_tensor(dna, device=eng.device)
        if dna.shape[0] == self.P and self.world > 1:
            dna = dna[self.rank_id * self.per:(self.rank_id + 1) * self.per]
        if tuple(dna.shape) != self.shape:
            raise ValueError("parents must be [%d or %d, %d, %d]" % ((self.P,) + self.shape))
        self._par[self._cur].copy_(dna.to(self.gene_dtype))
        obj = eng.evaluate_genes(self._par[self._cur]).obj
        self._glad[self._cur][:self.P].copy_(_all_gather(obj, self.group) if self.world > 1 else obj)
        self._fit[self._cur].copy_(eng.maximin(self.obj, self.constraint(gen)))
        self._sync()

    def gather_parents(self):
        """Full population [P, A, L] int32 on every rank (host views, pickling - not the hot loop)."""
        p = self.parents
        if self.world > 1:
            p = _all_gather(p.contiguous().view(torch.uint8).view(self.per, -1), self.group)
            p = p.view(self.gene_dtype).view(self.P, self.shape[1], self.shape[2])
        return p.to(torch.int32)

    def _sync(self):
        """Rendezvous: peers may touch this rank's shards only between two of these."""
        if self.world == 1:
            return
        if self.barrier_mode == "peer":
            self._epoch += 1
            self.eng.peer_barrier(self._flag_tab, self.rank_id, self._epoch, self._timed_out)
        elif dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self._sync_buf, group=self.group)          # stream-ordered
        else:
            torch.cuda.current_stream(self.eng.device).synchronize()
            dist.barrier(group=self.group)

    def check(self):
        """Raises if a device rendezvous gave up waiting for a peer (host sync; call off the hot loop)."""
        if int(self._timed_out.item()):
            raise RuntimeError("gcp_peer_barrier timed out: a rank of the sharded population is gone")

    # ---- one generation ------------------------------------------------------------------
    def step(self, seed, gen):
        eng, P, per, r = self.eng, self.P, self.per, self.rank_id
        cur, nxt = self._cur, self._cur ^ 1
        glad, glad_next = self._glad[cur], self._glad[nxt]
        r0 = float(eng.lib.gcp_low_var_r0(int(seed), int(gen), P))
        strong = eng.low_var_select(self._fit[cur], self.gamma, r0)
        eng.crossover_shards(self._par_tab[cur], self.gb, strong, P, r * (per // 2), per // 2,
                             seed, gen, self._kid)
        eng.mutate_genes(self._kid, r * per, seed, gen, salt=0)
        if self.world > 1:
            eng.evaluate_genes(self._kid, self._kid_obj, self._kid_flags)
            if dist.get_backend(self.group) == "nccl":
                dist.all_gather_into_tensor(glad[P:], self._kid_obj, group=self.group)
            else:
                glad[P:].copy_(_all_gather(self._kid_obj, self.group))
        else:
            eng.evaluate_genes(self._kid, glad[P:], self._kid_flags)
        fit = eng.maximin(glad, self.constraint(gen))
        order, _ = eng.arena_order(fit)
        eng.gather_shards(order, P, r * per, per, self._par_tab[cur], self._kid_tab, self.gb,
                          glad, fit, self._par[nxt], glad_next, self._fit[nxt])
        self._cur = nxt
        self._sync()
