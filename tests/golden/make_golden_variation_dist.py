#!/usr/bin/env python
"""Reference DISTRIBUTIONS of the random variation operators, sampled from the unmodified
reference on config C1 (data/map_scaled.png + wilk_3 graph, 6 agents): the Philox kernels of
the CUDA path cannot be equal draw for draw to numpy's MT19937, so `-m gpu` tests compare their
output distributions with these samples (two-sample chi-square, tests/test_gpu_variation.py).

Statistics are functions of the OUTPUT rows only, so both sides compute them the same way
(`row_stats` below is imported by the test):
  walks        PathMaker.makeMeAPath(150, start): node after steps 1, 3, 10; number of distinct nodes
  mutation     Organism constructor with mutation_prob = 1: common prefix with the input row, length
  muterpolate  Organism constructor with muterpolate_prob = 1: length removed, common prefix
  crossover    Organism.crossover of one fixed pair (variation off): per agent the common prefix of
               child 1 with the mother and its length (functions of the cut pair (i, j))

Run from the repo root: python tests/golden/make_golden_variation_dist.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

N_WALK, N_MUT, N_MUTERP, N_XO = 700, 1500, 1500, 2500


def valid_len(rows):
    """rows [..., L] -1 padded -> valid length"""
    return (np.asarray(rows) != -1).sum(-1)


def common_prefix(rows, ref_row):
    """length of the common prefix of every row with ref_row (both -1 padded to the same L)"""
    rows = np.asarray(rows)
    ne = rows != np.asarray(ref_row)[None, :]
    return np.where(ne.any(-1), ne.argmax(-1), rows.shape[-1])


def n_distinct(rows):
    return np.array([len(set(r[r != -1].tolist())) for r in np.asarray(rows)])


def main():
    import ref_shim
    ga, mappy_m, pm_m, got = ref_shim.load()
    m, p, (mp, pp, op, gp) = ref_shim.build_world("small", 6)
    L, A = op['max_dna_len'], 6
    out = {}
    with ref_shim.stable_argsort():
        # ---- walks ------------------------------------------------------------------------
        np.random.seed(100)
        walks = np.array([[p.makeMeAPath(150, s) for s in op['start_idx']] for _ in range(N_WALK)])
        out["walks"] = walks.astype(np.int16)                         # [N_WALK, A, 151]
        # ---- fixed input organism (no variation) ------------------------------------------
        o0 = dict(op)
        o0['mutation_prob'] = o0['muterpolate_prob'] = 0.0
        np.random.seed(101)
        base = ga.Organism([p.makeMeAPath(120, s) for s in op['start_idx']], m, p, o0)
        mate = ga.Organism([p.makeMeAPath(150, s) for s in op['start_idx']], m, p, o0)
        out["base"], out["mate"] = base._dna.astype(np.int16), mate._dna.astype(np.int16)
        rows = [r[r != -1] for r in base._dna]
        # ---- mutation ------------------------------------------------------------------------
        o1 = dict(op)
        o1['mutation_prob'], o1['muterpolate_prob'] = 1.0, 0.0
        np.random.seed(102)
        out["mutation"] = np.array([ga.Organism([r.copy() for r in rows], m, p, o1)._dna
                                    for _ in range(N_MUT)]).astype(np.int16)
        # ---- muterpolate ---------------------------------------------------------------------
        o2 = dict(op)
        o2['mutation_prob'], o2['muterpolate_prob'] = 0.0, 1.0
        np.random.seed(103)
        out["muterpolate"] = np.array([ga.Organism([r.copy() for r in rows], m, p, o2)._dna
                                       for _ in range(N_MUTERP)]).astype(np.int16)
        # ---- crossover (children's constructor variation off; crossover_prob as shipped) ----
        np.random.seed(104)
        kids = []
        for _ in range(N_XO):
            k = base.crossover(mate)
            kids.append([k[0]._dna, k[1]._dna])
        out["crossover"] = np.array(kids).astype(np.int16)            # [N_XO, 2, A, L]
    np.savez_compressed(os.path.join(HERE, "variation_dist.npz"), **out)
    print("variation_dist golden:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
