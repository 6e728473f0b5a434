#!/usr/bin/env python
"""Golden vectors for the deterministic parts of the variation path, produced by the
unmodified reference: Organism.pruneUTurns + addTelomere (geneticalgorithm.py:335-357).
Run from the repo root: python tests/golden/make_golden_variation.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402


def main():
    ga, mappy_m, pm_m, got = ref_shim.load()
    m, p, (mp, pp, op, gp) = ref_shim.build_world("big", 6)
    g = p._graph
    N = g.shape[0]
    A, L = 6, op['max_dna_len']
    rng = np.random.RandomState(21)
    before, after = [], []
    for case in range(60):
        rows = []
        for a in range(A):
            n = int(rng.randint(1, L + 1))
            w = int(rng.randint(N))
            row = []
            while len(row) < n:
                row.append(w)
                r = rng.rand()
                nb = np.nonzero(g[w])[0]
                if r < 0.15 and len(row) >= 2:
                    w = row[-2]                      # A-B-A u-turn
                elif r < 0.2:
                    pass                             # A-A stall
                elif r < 0.25 and len(row) >= 3:
                    w = row[-3]                      # A-B-C-A (gap 3: must survive)
                else:
                    w = int(rng.choice(nb))
            rows.append(np.array(row[:L]))
        org = ga.Organism.__new__(ga.Organism)
        org._mappy = m
        org._num_agents = A
        org._max_dna_len = L
        org._dna_list = [r.copy() for r in rows]
        org._len_dna = np.array([len(r) for r in rows])
        org.addTelomere()
        before.append(org._dna.copy())
        org.pruneUTurns()
        after.append(org._dna.copy())
    np.savez_compressed(os.path.join(HERE, "prune.npz"), before=np.array(before).astype(np.int16),
                        after=np.array(after).astype(np.int16))
    print("prune golden:", len(before), "organisms; rows changed:",
          int((np.array(before) != np.array(after)).any(axis=2).sum()))


if __name__ == "__main__":
    main()
