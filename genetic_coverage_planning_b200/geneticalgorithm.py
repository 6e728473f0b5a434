"""Drop-in for the reference's geneticalgorithm.py: same class names, constructor
signatures, attributes and return values, with the population living on the GPU.

    from genetic_coverage_planning_b200.geneticalgorithm import GeneticalGorithm, Organism
    population = GeneticalGorithm(mappy, pather, gen_params)      # pather_driver.py:235
    pareto_hist = population.runEvolution(100)                    # pather_driver.py:240

`mappy` / `pather` are the reference's own Mappy / PathMaker objects (or anything with the
same attributes: _img, _traverse_frustum/_angles/_dists, _graph, _XY, ...).  They are packed
once (packing.pack_from_reference) and uploaded; after that each generation is a fixed
sequence of kernels (select -> crossover -> mutate -> evaluate -> maximin -> order ->
gather) with no host round trip except the per-generation objective list the reference
also returns.

Differences from the reference, all deliberate:
  * random draws come from Philox streams keyed by (seed, generation, organism, agent),
    not numpy's global MT19937: runs are reproducible for a seed but not draw-for-draw equal;
  * survivors are ordered by the stable (fit, index) rule;
  * parents are never deep-copied; `findCrossoverCand()` is not required to have been called.
"""
import copy

import numpy as np
import torch

from . import packing
from .engine import FitnessEngine

_ENGINES = {}
_MAX_ENGINES = 4        # each holds the packed tables of one world on the GPU


def _engine_for(mappy, pather, org_params, num_agents, device=0):
    """One engine (packed tables + CUDA context) per (mappy, pather, A, L, device).  The scalar
    parameters (constraints, variation probabilities, coverage blend, ...) are refreshed on every
    call, so two populations on the same world with different parameters do not share them."""
    key = (id(mappy), id(pather), int(num_agents), int(org_params['max_dna_len']), int(device))
    ent = _ENGINES.get(key)
    path_params = {'path_memory': getattr(pather, '_path_memory', 10),
                   'log_scale_weight': getattr(pather, '_log_scale_weight', 0.7)}
    if ent is None or ent[1] is not mappy or ent[2] is not pather:
        pw = packing.pack_from_reference(mappy, pather)
        eng = FitnessEngine(pw, org_params, num_agents=num_agents, device=device)
        ent = (eng, mappy, pather)
        while len(_ENGINES) >= _MAX_ENGINES:              # oldest first; its tables are freed with it
            _ENGINES.pop(next(iter(_ENGINES)))
        _ENGINES[key] = ent
    eng = ent[0]
    eng.set_params(org_params, solo_sep_thresh=getattr(mappy, '_solo_sep_thresh', None),
                   blend=getattr(mappy, '_coverage_blend', None))
    eng.set_vary_params(org_params, path_params)
    return eng


def _pad_rows(dna_list, L):
    d = -np.ones((len(dna_list), L), dtype=np.int32)
    for a, row in enumerate(dna_list):
        row = np.asarray(row, dtype=np.int64)
        row = row[row != -1][:L]
        d[a, :len(row)] = row
    return d


class Organism(object):
    """geneticalgorithm.py:126-357.  Constructed from a list of per-agent waypoint arrays;
    variation (mutation / muterpolate / U-turn pruning) and `calcObj` run on the GPU."""

    def __init__(self, dna_list, mappy, pather, org_params, _precomputed=None, _seed=None):
        self._max_dna_len = org_params['max_dna_len']
        self._min_dna_len = org_params['min_dna_len']
        self._xover_probability = org_params['crossover_prob']
        self._time_thresh = org_params['crossover_time_thresh']
        self._mutate_probability = org_params['mutation_prob']
        self._muterpolate_probability = org_params['muterpolate_prob']
        self._num_muterpolations = org_params['num_muterpolations']
        self._srch_dist = org_params['muterpolation_srch_dist']
        self._P_muterpolate_each = org_params['muterpolation_sub_prob']
        self._min_solo_lcs = org_params['min_solo_lcs']
        self._min_comb_lcs = org_params['min_comb_lcs']
        self._ft_scale = org_params['flight_time_scale']
        self._org_params = org_params
        self._pather = pather
        self._mappy = mappy
        self._scale = getattr(mappy, '_scale', None)
        self._narrowest_hall = getattr(mappy, '_hall_width', None)
        self._safety_buffer = getattr(mappy, '_safety_buffer', None)
        self._num_agents = len(dna_list)
        self._obj_val_sc = [None, None]
        if _precomputed is not None:                       # view of a device-side organism
            self._dna = np.asarray(_precomputed[0]).astype(int)
            self._obj_val = [float(_precomputed[1][0]), float(_precomputed[1][1])]
            if len(_precomputed) > 2 and _precomputed[2] is not None:
                self._obj_val_sc = [float(_precomputed[2][0]), float(_precomputed[2][1])]
        else:
            eng = _engine_for(mappy, pather, org_params, self._num_agents)
            dna = torch.from_numpy(_pad_rows(dna_list, self._max_dna_len)[None]).to(eng.device)
            seed = int(np.random.randint(0, 2 ** 31 - 1)) if _seed is None else int(_seed)
            eng.mutate(dna, seed, 0, salt=7)               # constructor variation :158-168
            self._dna = dna[0].cpu().numpy().astype(int)
            self._obj_val = self.calcObj()
        self._dna_list = [row[row != -1] for row in self._dna]
        # reference quirk F12: _len_dna ends up as the padded length
        self._len_dna = (np.ones(self._num_agents) * self._max_dna_len).astype(int)

    def calcObj(self):
        """geneticalgorithm.py:177-201 on the GPU."""
        eng = _engine_for(self._mappy, self._pather, self._org_params, self._num_agents)
        r = eng.evaluate(self._dna[None].astype(np.int32), detail=False)
        eng.raise_on_error(r)
        o = r.obj[0].cpu().numpy()
        return [float(o[0]), float(o[1])]

    def crossover(self, mate):
        """geneticalgorithm.py:211-249 -> [kid, kid]."""
        eng = _engine_for(self._mappy, self._pather, self._org_params, self._num_agents)
        parents = torch.from_numpy(np.stack([self._dna, mate._dna]).astype(np.int32)).to(eng.device)
        strong = torch.tensor([0, 1], dtype=torch.int32, device=eng.device)
        seed = int(np.random.randint(0, 2 ** 31 - 1))
        kids = eng.crossover(parents, strong, seed, 0)
        eng.mutate(kids, seed, 0, salt=0)
        objs = eng.evaluate(kids, detail=False).obj.cpu().numpy()
        kids = kids.cpu().numpy()
        return [Organism([None] * self._num_agents, self._mappy, self._pather, self._org_params,
                         _precomputed=(kids[k], objs[k])) for k in range(2)]


class GeneticalGorithm(object):
    """geneticalgorithm.py:15-123.  The population lives on the GPU(s): under torch.distributed
    (one process per GPU) it is ONE population sharded over the ranks (distributed.ShardedEvolution);
    every rank must construct it with the same arguments - the seed is taken from rank 0."""

    def __init__(self, mappy, pather, gen_params, device=0, seed=None, initial_dna=None, gen_num=0):
        self._G_sz = gen_params['gen_size']
        self._starting_path_len = gen_params['starting_path_len']
        self._num_agents = gen_params['num_agents']
        self._gamma = gen_params['gamma']
        self._cov_constr_0 = -gen_params['coverage_constr_0']
        self._cov_constr_f = -gen_params['coverage_constr_f']
        self._cov_aging = gen_params['coverage_aging']
        self._org_params = gen_params['org_params']
        self._mappy, self._pather = mappy, pather
        self._device = device
        self._seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        self._seed = _same_on_all_ranks(self._seed, device)
        self._gen_num = int(gen_num)
        self._views = None
        self._dna_cache = None
        if self._G_sz % 2:
            raise ValueError("gen_size must be even (pather_driver.py:107)")
        self._eng = _engine_for(mappy, pather, self._org_params, self._num_agents, device)
        self._evo = self._new_evo()
        if initial_dna is not None:
            dna = self._eng.to_device_dna(initial_dna)
        else:
            dna = self.firstGeneration(mappy, pather, self._org_params)
        self._evo.load(dna, self._gen_num)
        self._obj_sc = self._scaled()

    def _new_evo(self):
        import torch.distributed as dist
        from .distributed import ShardedEvolution
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        # a population that does not split into whole pairs per rank stays replicated: every rank
        # evolves all of it (same seed, same result) instead of failing
        return ShardedEvolution(self._eng, {'gamma': self._gamma,
                                            'coverage_constr_0': -self._cov_constr_0,
                                            'coverage_constr_f': -self._cov_constr_f,
                                            'coverage_aging': self._cov_aging}, self._G_sz,
                                shard=(self._G_sz % (2 * world) == 0))

    def _scaled(self):
        return self._eng.maximin(self._evo.obj, self._coverage_constr(), return_scaled=True)[1]

    @classmethod
    def from_population(cls, population, gen_params=None, device=0, seed=None):
        """Continue a population that was evolved (or unpickled, gori_tools.py:255-257) with the
        REFERENCE's GeneticalGorithm - or anything that looks like one: `_gen_parent` (objects
        with `_dna`, `_mappy`, `_pather`), `_gen_num`, and the scalar attributes of
        geneticalgorithm.py:16-35.  The parents are re-evaluated on the GPU (their stored
        objectives are not trusted), ranked, and evolution continues from `_gen_num`."""
        first = population._gen_parent[0]
        if gen_params is None:
            gen_params = {'gen_size': population._G_sz,
                          'starting_path_len': population._starting_path_len,
                          'num_agents': population._num_agents, 'gamma': population._gamma,
                          'coverage_constr_0': -population._cov_constr_0,
                          'coverage_constr_f': -population._cov_constr_f,
                          'coverage_aging': population._cov_aging,
                          # the reference keeps org_params on its organisms only (geneticalgorithm.py:25,130)
                          'org_params': getattr(population, '_org_params', None) or first._org_params}
        L = gen_params['org_params']['max_dna_len']
        dna = np.stack([_pad_rows(list(o._dna), L) for o in population._gen_parent]).astype(np.int32)
        pop = cls(first._mappy, first._pather, gen_params, device=device, seed=seed,
                  initial_dna=dna, gen_num=int(getattr(population, '_gen_num', 0)))
        return pop

    # -- reference: geneticalgorithm.py:37-53 ------------------------------------------
    def firstGeneration(self, mappy, pather, org_params, max_batches=10000):
        """Random paths + constructor variation + evaluation in batches, keeping the
        candidates that meet the initial coverage constraint (:50), in generation order.
        Returns the [P, A, L] int32 chromosomes (identical on every rank: same seed)."""
        eng, P = self._eng, self._G_sz
        start = org_params['start_idx']
        got_dna, n = [], 0
        batch_sz = int(min(max(2 * P, 256), 1 << 16))
        for b in range(max_batches):
            dna = eng.random_walks(batch_sz, start, self._starting_path_len, self._seed, batch=b)
            eng.mutate(dna, self._seed, 0xFFFF0000 + b, salt=3)
            r = eng.evaluate(dna, detail=False)
            eng.raise_on_error(r)
            ok = r.obj[:, 0] <= self._cov_constr_0
            k = int(ok.sum().item())
            if k:
                got_dna.append(dna[ok])
                n += k
            if n >= P:
                break
        if n < P:
            raise RuntimeError("firstGeneration: only %d of %d organisms met coverage >= %g"
                               % (n, P, -self._cov_constr_0))
        return torch.cat(got_dna)[:P].contiguous()

    def _coverage_constr(self):
        alpha = min(1, (self._gen_num / self._cov_aging))                    # :98
        return (1 - alpha) * self._cov_constr_0 + alpha * self._cov_constr_f  # :99-100

    # -- reference: geneticalgorithm.py:55-76 ------------------------------------------
    def runEvolution(self, num_generations):
        evo = self._evo
        hist_dev = []
        for _ in range(num_generations):
            evo.step(self._seed, self._gen_num)          # select -> crossover -> variation -> evaluate -> arena
            self._gen_num += 1
            hist_dev.append(evo.obj.clone())
        self._views = None
        self._dna_cache = None
        if not hist_dev:
            return []
        self._obj_sc = self._scaled()
        evo.check()
        self._eng.raise_on_error(flags=evo._kid_flags)    # chromosome / coverage-pass errors of the last children
        flat = torch.stack(hist_dev).cpu().numpy()                            # one D2H at the end
        return [[[float(o[0]), float(o[1])] for o in gen] for gen in flat]

    RunEvolution = runEvolution          # README.md:17 spells it this way

    # -- reference: geneticalgorithm.py:78-123 -----------------------------------------
    def rackNstack(self, gladiators):
        """Maximin fitness of a list of Organism-likes; also fills their _obj_val_sc."""
        objs = np.array([[float(g._obj_val[0]), float(g._obj_val[1])] for g in gladiators])
        fit, sc = self._eng.maximin(torch.from_numpy(objs), self._coverage_constr(),
                                    return_scaled=True)
        sc = sc.cpu().numpy()
        for g, s in zip(gladiators, sc):
            g._obj_val_sc = [float(s[0]), float(s[1])]
        return [float(f) for f in fit.cpu().numpy()]

    def arena(self, gladiators):
        fits = self.rackNstack(gladiators)
        order, _ = self._eng.arena_order(torch.tensor(fits, dtype=torch.float64))
        order = order.cpu().numpy()[:self._G_sz]
        keep = [gladiators[i] for i in order]
        self._evo.load(np.stack([g._dna for g in keep]).astype(np.int32), self._gen_num)
        self._obj_sc = self._scaled()
        self._views = None
        self._dna_cache = None

    # -- device state ------------------------------------------------------------------------
    _LAZY = ('_eng', '_evo', '_obj_sc')

    def __getattr__(self, name):
        # after unpickling, the CUDA context and the device tensors are created on first use, so a
        # pickled population opens on a machine without a GPU (plotting / export only read host copies)
        if name in GeneticalGorithm._LAZY and '_host_state' in self.__dict__:
            self._materialise()
            return self.__dict__[name]
        raise AttributeError(name)

    def _materialise(self):
        dna, obj, fit, sc = self.__dict__.pop('_host_state')
        self._eng = _engine_for(self._mappy, self._pather, self._org_params, self._num_agents,
                                self._device)
        self._evo = self._new_evo()
        self._evo.load(dna, self._gen_num)
        self._obj_sc = self._scaled()

    @property
    def _dna(self):
        """[P, A, L] int32 device tensor of all parents (gathered from the shards on first use)."""
        if self.__dict__.get('_dna_cache') is None:
            self._dna_cache = self._evo.gather_parents()
        return self._dna_cache

    @property
    def _obj(self):
        return self._evo.obj

    @property
    def _fit(self):
        return self._evo.fit

    def _host_arrays(self):
        if '_host_state' in self.__dict__:
            return self.__dict__['_host_state']
        return (self._dna.cpu().numpy(), self._obj.cpu().numpy(), self._fit.cpu().numpy(),
                self._obj_sc.cpu().numpy() if self._obj_sc is not None else None)

    # -- attributes the reference's plotting / export code reads -------------------------
    @property
    def _gen_parent(self):
        """List of Organism views (gori_tools.py:23,248-250; pointselector.py:52), built on
        first access after each change; no table copies."""
        if self._views is None:
            dna, obj, _, sc = self._host_arrays()
            if sc is None:
                sc = [None] * len(dna)
            self._views = [Organism([None] * self._num_agents, self._mappy, self._pather,
                                    self._org_params, _precomputed=(dna[k], obj[k], sc[k]))
                           for k in range(len(dna))]
        return self._views

    @property
    def _gen_parent_fit(self):
        return [float(f) for f in self._host_arrays()[2]]

    # -- pickle (gori_tools.py:252-257): device state round-trips through numpy ----------
    def __getstate__(self):
        st = {k: v for k, v in self.__dict__.items() if k not in ('_eng', '_evo', '_views', '_dna_cache',
                                                                   '_obj_sc', '_host_state')}
        dna, obj, fit, sc = self._host_arrays()
        st['_dna_np'], st['_obj_np'], st['_fit_np'], st['_obj_sc_np'] = dna, obj, fit, sc
        return st

    def __setstate__(self, st):
        st = dict(st)
        host = (st.pop('_dna_np'), st.pop('_obj_np'), st.pop('_fit_np'), st.pop('_obj_sc_np'))
        self.__dict__.update(st)
        self._views = None
        self._dna_cache = None
        self.__dict__['_host_state'] = host        # device side is created on first use (_materialise)


def _same_on_all_ranks(seed, device):
    """Under torch.distributed every rank must draw the same population: rank 0's seed wins."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return seed
    dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([seed], dtype=torch.int64, device=dev)
    dist.broadcast(t, src=0)
    return int(t.item())


def getObjValsList(population):
    """gori_tools.py:248-250 without materialising Organism views."""
    return [[float(o[0]), float(o[1])] for o in population._host_arrays()[1]]
