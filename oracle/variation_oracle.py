"""CPU restatement (numpy) of the reference's selection / variation path and generation loop.

TEST INFRASTRUCTURE.  This module is part of the ORACLE: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it; the product package never
does.

Parity status: PINNED.  Every function draws from numpy's global MT19937 stream through the SAME
numpy calls in the SAME order as the reference (np.random.rand / randint / choice / uniform), so
after `np.random.seed(s)` it reproduces the reference's paths, children and whole generations
exactly.  tests/test_oracle_variation.py checks that against fixtures made by the unmodified
reference (tests/golden/make_golden_variation_rng.py -> variation_rng.npz) and, when
/root/reference is present, against the live reference.  Canonicalisation as everywhere in this
repo: the reference is run with np.argsort forced to kind='stable'.

Restated (citations into /root/reference/src):
  pathmaker.py:175-186   PathMaker.assignWeights      -> assign_weights
  pathmaker.py:188-213   PathMaker.makeMeAPath        -> make_me_a_path
  geneticalgorithm.py:127-175  Organism.__init__      -> Organism
  geneticalgorithm.py:203-209  padMinDNA              -> Organism.pad_min_dna
  geneticalgorithm.py:211-249  crossover              -> Organism.crossover
  geneticalgorithm.py:251-268  mutation               -> Organism.mutation
  geneticalgorithm.py:270-309  muterpolate            -> Organism.muterpolate
  geneticalgorithm.py:311-333  matchWaypt             -> Organism.match_waypt
  geneticalgorithm.py:335-344  addTelomere            -> Organism.add_telomere
  geneticalgorithm.py:346-357  pruneUTurns            -> Organism.prune_uturns
  geneticalgorithm.py:37-84    firstGeneration / runEvolution / arena -> Population
  gori_tools.py:207-229        lowVarSample           -> low_var_sample

Third-party arithmetic the reference calls here: scipy.stats.norm.logpdf (pathmaker.py:182),
restated below from scipy's definition (checked equal to scipy in the tests).
"""
import numpy as np

import fitness_oracle as fo

_NORM_PDF_LOGC = np.log(np.sqrt(2 * np.pi))      # scipy.stats._continuous_distns._norm_pdf_logC


def norm_logpdf(x, scale):
    """scipy.stats.norm.logpdf(x, scale=scale): rv_continuous.logpdf standardises
    y = (x - 0) / scale, evaluates _norm_logpdf(y) = -y**2 / 2.0 - _norm_pdf_logC and
    subtracts log(scale)."""
    y = np.asarray((np.asarray(x) - 0) / scale, dtype=np.float64)
    return (-y ** 2 / 2.0 - _NORM_PDF_LOGC) - np.log(scale)


class Graph(object):
    """Traversability queries on the oracle World's CSR graph with the reference's dense-matrix
    semantics (`_graph[a, b]`, `_graph[a]`; negative indices wrap like numpy's)."""

    def __init__(self, world):
        self.N = world.N
        self.row_ptr = world.row_ptr
        self.col = world.col
        self.xy = world.xy
        self._sets = None

    def out(self, a):
        """np.where(self._graph[a])[0]: ascending out-neighbours."""
        a = int(a)
        if a < 0:
            a += self.N
        c = self.col[self.row_ptr[a]:self.row_ptr[a + 1]]
        return np.sort(c).astype(np.int64)

    def is_edge(self, a, b):
        b = int(b)
        if b < 0:
            b += self.N
        return bool((self.out(a) == b).any())

    def common(self, a, b):
        """np.where(graph[a] * graph[b])[0]"""
        return np.intersect1d(self.out(a), self.out(b))


def assign_weights(g, current_heading, current_idx, choices, log_scale_weight):
    """pathmaker.py:175-186"""
    xy = g.xy
    displacements = np.array([xy[choices, 0] - xy[current_idx, 0],
                              xy[choices, 1] - xy[current_idx, 1]])
    angles = np.arctan2(displacements[1], displacements[0]) - current_heading
    angles = fo.rad_wrap_pi(angles)
    w = norm_logpdf(angles, log_scale_weight)
    w -= np.max(w)
    w = np.exp(w)
    w = w / np.sum(w)
    return w


def make_me_a_path(g, path_len, start_idx, prev_path=None, path_memory=10, log_scale_weight=0.7):
    """pathmaker.py:188-213 (draws: one rand() for the heading, one choice(p=w) per step)."""
    current_idx = start_idx
    current_heading = np.random.rand() * 2 * np.pi - np.pi
    if prev_path is not None:
        path_idx = prev_path[-path_memory:]
    else:
        path_idx = np.array([start_idx])
    lop_len = len(path_idx) - 1
    for _ in range(path_len):
        choices = g.out(current_idx)
        choices_comb = np.setdiff1d(choices, path_idx[-path_memory:])
        if len(choices_comb) > 0:
            w = assign_weights(g, current_heading, current_idx, choices_comb, log_scale_weight)
            new_idx = np.random.choice(choices_comb, p=w)
        else:
            w = assign_weights(g, current_heading, current_idx, choices, log_scale_weight)
            new_idx = np.random.choice(choices, p=w)
        path_idx = np.append(path_idx, new_idx)
        current_idx = new_idx
    return path_idx[lop_len:].astype(int)


class Organism(object):
    """geneticalgorithm.py:126-357 on the oracle World.  `world.path_params` (path_memory,
    log_scale_weight) default to the driver's values (pather_driver.py:173-177)."""

    def __init__(self, dna_list, world, org_params, graph=None, evaluate=True):
        self.world = world
        self.g = graph if graph is not None else Graph(world)
        op = org_params
        self.op = op
        pp = getattr(world, 'path_params', None) or {}
        self.path_memory = pp.get('path_memory', 10)
        self.log_scale_weight = pp.get('log_scale_weight', 0.7)
        self.L = op['max_dna_len']
        self.dna_list = list(dna_list)
        self.A = len(self.dna_list)
        self.len_dna = np.ones(self.A).astype(int)
        for a in range(self.A):
            self.len_dna[a] = len(self.dna_list[a])
        self.add_telomere()                                        # :156
        do_it = np.random.rand()                                   # :159
        if do_it < op['mutation_prob']:
            self.mutation()
        do_it = np.random.rand()                                   # :163
        if do_it < op['muterpolate_prob']:
            self.muterpolate()
        self.prune_uturns()                                        # :168
        self.obj_val = self.calc_obj() if evaluate else [None, None]   # :171
        self.obj_val_sc = [None, None]
        for a in range(self.A):
            self.len_dna[a] = len(self.dna_list[a])                # :173-174 (F12: the padded length)
        self.add_telomere()

    def _path(self, n, start, prev=None):
        return make_me_a_path(self.g, n, start, prev, self.path_memory, self.log_scale_weight)

    def calc_obj(self):
        return list(fo.calc_obj(self.world, self.dna, self.op))

    def add_telomere(self):                                        # :335-344
        for a in range(self.A):
            if self.len_dna[a] > self.L:
                self.dna_list[a] = self.dna_list[a][0:self.L]
                self.len_dna[a] = self.L
            elif self.len_dna[a] < self.L:
                self.dna_list[a] = np.append(self.dna_list[a],
                                             np.ones(self.L - self.len_dna[a]) * -1).astype(int)
        self.dna = np.vstack(self.dna_list)

    def pad_min_dna(self, dna):                                    # :203-209
        dna = dna[dna != -1]
        if len(dna) < self.op['min_dna_len']:
            len_tail = self.op['min_dna_len'] - len(dna)
            dna_tail = self._path(len_tail, dna[-1])
            dna = np.append(dna, dna_tail[1:])
        return dna

    def match_waypt(self, dna_mate, len_dna_mate):                 # :311-333
        keepout_idx = 1
        cand = self.world.crossover_cand()
        xover_pts = []
        for a in range(self.dna.shape[0]):
            dna_poss = self.dna[a][self.dna[a] != -1]
            mate_poss = dna_mate[a][dna_mate[a] != -1]
            poss_mask = np.isin(dna_poss, cand)
            dna_poss[~poss_mask] = -1
            poss_mask = np.isin(mate_poss, cand)
            mate_poss[~poss_mask] = -10
            brd = np.abs(dna_poss[keepout_idx:self.len_dna[a] + 1, None] -
                         mate_poss[None, keepout_idx:len_dna_mate[a] + 1])
            un_pruned = np.array(np.where(brd == 0)).T + keepout_idx
            tf = np.abs(un_pruned[:, 0] - un_pruned[:, 1]) < self.op['crossover_time_thresh']
            xover_pts.append(un_pruned[tf])
        return xover_pts

    def crossover(self, mate):                                     # :211-249
        uniform_num = np.random.rand(self.A)
        xover_pts = self.match_waypt(mate.dna, mate.len_dna)
        new1, new2 = [], []
        for a in range(self.A):
            if uniform_num[a] < self.op['crossover_prob'] and len(xover_pts[a]) > 0:
                xover_idx = np.arange(len(xover_pts[a]))
                single_idx = np.random.choice(xover_idx)
                pt = xover_pts[a][single_idx]
                dna1 = self.dna[a][0:pt[0]]
                dna2 = mate.dna[a][0:pt[1]]
                dna1 = np.append(dna1, mate.dna[a][pt[1]:])
                dna2 = np.append(dna2, self.dna[a][pt[0]:])
                dna1 = self.pad_min_dna(dna1)
                dna2 = self.pad_min_dna(dna2)
            else:
                dna1 = self.dna[a]
                dna2 = mate.dna[a]
            dna1 = dna1[dna1 != -1]
            dna2 = dna2[dna2 != -1]
            new1.append(dna1)
            new2.append(dna2)
        kid1 = Organism(new1, self.world, self.op, self.g)
        kid2 = Organism(new2, self.world, self.op, self.g)
        return [kid1, kid2]

    def mutation(self):                                            # :251-268
        for a in range(self.A):
            idx = np.random.randint(1, self.len_dna[a] - 1)
            len_tail = self.len_dna[a] - idx
            if (self.len_dna[a] + 1) < self.L:
                len_tail += np.random.randint(self.L - self.len_dna[a])
            dna_tail = self._path(len_tail, self.dna[a, idx], self.dna[a, 0:idx + 1])
            dna_head = self.dna_list[a][0:idx]
            self.dna_list[a] = np.append(dna_head, dna_tail)
            self.len_dna[a] = len(self.dna_list[a])
        self.add_telomere()

    def muterpolate(self):                                         # :270-309
        op = self.op
        srch = op['muterpolation_srch_dist']
        for a in range(self.A):
            try:
                mut_points = np.random.choice(np.arange(0, self.len_dna[a] - (srch + 3), 1),
                                              op['num_muterpolations'])
                mut_points = np.sort(mut_points)[::-1]
            except ValueError:                                     # "DNA is getting too short" :279-281
                mut_points = []
            for idx1 in mut_points:
                for idx2 in range(idx1 + srch + 1, idx1 + 1, -1):
                    do_it = np.random.rand()
                    if do_it < op['muterpolation_sub_prob']:
                        try:
                            pt1 = self.dna_list[a][idx1]
                            pt2 = self.dna_list[a][idx2]
                        except IndexError:
                            break
                        if self.g.is_edge(pt1, pt2):
                            break                                   # np.delete result discarded :293
                        options = self.g.common(pt1, pt2)
                        if len(options) > 0:
                            step = np.random.choice(options)
                            self.dna_list[a][idx1 + 1] = step
                            self.dna_list[a] = np.delete(self.dna_list[a],
                                                         np.arange(idx1 + 2, idx2 - 1, 1))
                            break
            self.dna_list[a] = self.dna_list[a][self.dna_list[a] != -1]
            self.len_dna[a] = len(self.dna_list[a])
        self.add_telomere()

    def prune_uturns(self):                                        # :346-357
        for a in range(self.A):
            _, dups = fo.get_duplicate_wps(self.dna[a][self.dna[a] != -1])
            for dup in dups:
                dup = np.sort(dup)
                diffs = np.diff(dup)
                for ii, diff in enumerate(diffs):
                    if diff <= 2:
                        self.dna[a, dup[ii]:dup[ii + 1]] = -1
            self.dna_list[a] = self.dna[a][self.dna[a] != -1]
            self.len_dna[a] = len(self.dna_list[a])
        self.add_telomere()


def low_var_sample(n, fitnesses, pressure):
    """gori_tools.py:207-229 -> indices of the resampled parents (one uniform draw)."""
    r = np.random.uniform(0, 1 / n)
    return fo.low_var_sample_indices(fitnesses, pressure, r)


class Population(object):
    """geneticalgorithm.py:15-84 on the oracle World (stable arena order)."""

    def __init__(self, world, gen_params, first_generation=True):
        self.world = world
        self.gp = gen_params
        self.op = gen_params['org_params']
        self.G = gen_params['gen_size']
        self.g = Graph(world)
        self.gen_num = 0
        self.parents = []
        self.fit = []
        if first_generation:
            self.first_generation()

    def first_generation(self):                                    # :37-53
        start = self.op['start_idx']
        pp = getattr(self.world, 'path_params', None) or {}
        while len(self.parents) < self.G:
            paths = [make_me_a_path(self.g, self.gp['starting_path_len'], start[a], None,
                                    pp.get('path_memory', 10), pp.get('log_scale_weight', 0.7))
                     for a in range(self.gp['num_agents'])]
            cand = Organism(paths, self.world, self.op, self.g)
            if cand.obj_val[0] <= -self.gp['coverage_constr_0']:
                self.parents.append(cand)
        self.fit = self.rack_n_stack(self.parents)

    def rack_n_stack(self, gladiators):                            # :86-123
        objs = np.array([g.obj_val for g in gladiators], dtype=np.float64)
        sc = fo.scaled_objectives(objs, self.gen_num, self.gp)
        for g, s in zip(gladiators, sc):
            g.obj_val_sc = [s[0], s[1]]
        return list(fo.rack_n_stack(objs, self.gen_num, self.gp))

    def arena(self, gladiators):                                   # :78-84
        fit = self.rack_n_stack(gladiators)
        order = fo.arena_order(np.array(fit))
        self.parents = [gladiators[i] for i in order[:self.G]]
        self.fit = [fit[i] for i in order[:self.G]]

    def run_evolution(self, n):                                    # :55-76
        hist = []
        for _ in range(n):
            idx = low_var_sample(len(self.parents), self.fit, self.gp['gamma'])
            strong = [self.parents[i] for i in idx]
            kids = []
            for mommy, daddy in zip(strong[:self.G // 2], strong[self.G // 2:]):
                kids += mommy.crossover(daddy)
            self.arena(self.parents + kids)
            self.gen_num += 1
            hist.append([list(p.obj_val) for p in self.parents])
        return hist
