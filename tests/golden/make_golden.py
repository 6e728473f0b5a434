#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

Run from the repo root:  python tests/golden/make_golden.py
Needs /root/reference (absent on the GPU box, which only reads the committed .npz).
Pins: python 3.12.3, numpy 2.3.5, cv2 4.13.0 (see each file's `versions` entry).

What is captured (all produced by reference code, never by this repo's oracle):
  world_{big,small}.npz   the inputs of the path: 8-bit map, waypoints, CSR graph and
                          the per-edge polygon / heading / length tables computed by
                          Mappy.computeFrustums (mappy.py:274-328)
  eval_{big,small}.npz    seeded populations -> Mappy.getCoverage, Mappy.getLoopClosures
                          (stable argsort = canonical, and as shipped), Organism.calcObj
  edge_{big}.npz          hand-made DNA edge cases (ragged / empty rows)
  rank.npz                GeneticalGorithm.rackNstack + got.sortBy (stable) + lowVarSample
"""
import os
import sys
import copy

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402
from world import World  # noqa: E402

import cv2  # noqa: E402

VERSIONS = np.array("numpy %s cv2 %s" % (np.__version__, cv2.__version__))


class stable_argsort(object):
    """SURVEY F9/F15 canonicalisation: np.argsort -> kind='stable' while active."""

    def __enter__(self):
        self._orig = np.argsort

        def _stable(a, *args, **kw):
            kw['kind'] = 'stable'
            return self._orig(a, *args, **kw)
        np.argsort = _stable

    def __exit__(self, *exc):
        np.argsort = self._orig


def eval_records(ga, m, p, op, dnas):
    """Run the reference on each dna [A,L] and collect every intermediate."""
    rec = dict(neg_cov=[], dist=[], turn=[], sum_draw=[], sum_draw_mask=[], lc=[],
               lc_shipped=[], obj=[], ncomp=[], n_painted=[], n_painted_mask=[])
    buffer_mask = np.logical_and(m._safety_img < 0.3, m._safety_img > 0)
    for dna in dnas:
        dna = np.asarray(dna).astype(int)
        c, d, t, draw = m.getCoverage(dna, return_map=True)
        rec['neg_cov'].append(c)
        rec['dist'].append(d)
        rec['turn'].append(t)
        rec['sum_draw'].append(np.sum(draw))
        rec['sum_draw_mask'].append(np.sum(draw[buffer_mask]))
        # painted-pixel counts: run the reference on an all-zero map of the same shape
        saved = m._img
        m._img = np.zeros_like(saved)
        _, _, _, zdraw = m.getCoverage(dna, return_map=True)
        m._img = saved
        rec['n_painted'].append(int(zdraw.sum()))
        rec['n_painted_mask'].append(int(zdraw[buffer_mask].sum()))
        rec['lc_shipped'].append(m.getLoopClosures(dna))
        with stable_argsort():
            rec['lc'].append(m.getLoopClosures(dna))
            # calcObj through a bare Organism shell (no mutation side effects)
            org = ga.Organism.__new__(ga.Organism)
            org._mappy = m
            org._dna = dna
            org._num_agents = dna.shape[0]
            org._min_solo_lcs = op['min_solo_lcs']
            org._min_comb_lcs = op['min_comb_lcs']
            org._ft_scale = op['flight_time_scale']
            rec['obj'].append([float(v) for v in org.calcObj()])
        lc = rec['lc'][-1]
        solo = np.diagonal(lc)
        combo = lc - np.diag(solo)
        adj = ((combo + combo.T) >= 1).astype(int)
        eig, _ = np.linalg.eig(np.diag(adj.sum(0)) - adj)
        rec['ncomp'].append(int(dna.shape[0] - np.sum(eig > 1e-6)))
    return {k: np.array(v) for k, v in rec.items()}


def seeded_population(ga, m, p, op, n_org, walk_len, seed):
    np.random.seed(seed)
    dnas = []
    for _ in range(n_org):
        paths = [p.makeMeAPath(walk_len, s) for s in op['start_idx']]
        org = ga.Organism(paths, m, p, op)      # mutation / muterpolate / prune shapes
        dnas.append(org._dna.copy())
    return np.array(dnas)


def edge_case_dnas(m, p, op):
    """Hand-made [A,L] rows hitting the documented quirks."""
    N = p._graph.shape[0]
    L = op['max_dna_len']
    A = len(op['start_idx'])
    g = p._graph
    rng = np.random.RandomState(7)

    def walk(start, n):
        out = [start]
        for _ in range(n - 1):
            nb = np.nonzero(g[out[-1]])[0]
            out.append(int(rng.choice(nb)))
        return out

    def pad(rows):
        d = -np.ones((A, L), dtype=int)
        for a, r in enumerate(rows):
            d[a, :len(r)] = r[:L]
        return d
    cases = []
    starts = op['start_idx']
    # full-length rows: no sentinel step (F3)
    cases.append(pad([walk(s, L) for s in starts]))
    # rows of 1 and 2 genes, one empty row
    rows = [walk(s, 40) for s in starts]
    rows[0] = rows[0][:1]
    rows[1] = rows[1][:2]
    if A > 2:
        rows[2] = []
    cases.append(pad(rows))
    # all rows empty
    cases.append(pad([[] for _ in starts]))
    # last valid waypoint has an edge into N-1 (sentinel reads a real frustum, F3)
    into_last = np.nonzero(g[:, N - 1])[0]
    rows = [walk(s, 30) for s in starts]
    if len(into_last):
        rows[0] = walk(starts[0], 20) + [int(into_last[0])]
    cases.append(pad(rows))
    # non-edge steps and out-of-order jumps (F4)
    rows = [walk(s, 50) for s in starts]
    rows[0][10] = (rows[0][9] + 17) % N
    rows[1][5] = rows[1][3]
    cases.append(pad(rows))
    # heavy repetition: every agent runs the same loop (solo + combo + >=3-agent groups)
    loop = walk(starts[0], 12)
    cases.append(pad([(loop * 8)[:90] for _ in starts]))
    # two agents share a stretch far apart in time, others disjoint; interior -1 hole
    rows = [walk(s, 120) for s in starts]
    rows[1][60:80] = rows[0][20:40]
    d = pad(rows)
    d[A - 1, 30] = -1       # hole: coverage stops there, loop closures keep later genes
    cases.append(d)
    # random walks with random lengths
    for _ in range(12):
        cases.append(pad([walk(s, int(rng.randint(1, L + 1))) for s in starts]))
    return np.array(cases)


def main():
    ga, mappy_m, pm_m, got = ref_shim.load()
    out = HERE

    for cfg, n_org, walk in (("big", 48, 150), ("small", 32, 120)):
        m, p, (mp, pp, op, gp) = ref_shim.build_world(cfg, 6)
        w = World.from_reference(m, p, op, name=cfg)
        w.save_npz(os.path.join(out, "world_%s.npz" % cfg))
        dnas = seeded_population(ga, m, p, op, n_org, walk, seed=0)
        rec = eval_records(ga, m, p, op, dnas)
        np.savez_compressed(os.path.join(out, "eval_%s.npz" % cfg), dna=dnas.astype(np.int16),
                            versions=VERSIONS, **rec)
        print(cfg, "eval:", len(dnas), "organisms; shipped-vs-stable lc mismatches:",
              int(np.sum(np.any(rec['lc'] != rec['lc_shipped'], axis=(1, 2)))))
        if cfg == "big":
            ed = edge_case_dnas(m, p, op)
            rec = eval_records(ga, m, p, op, ed)
            np.savez_compressed(os.path.join(out, "edge_big.npz"), dna=ed.astype(np.int16),
                                versions=VERSIONS, **rec)
            # blend 0.1 (paper value) and 1/3-agent variants reuse the same world tables
            m._coverage_blend = 0.1
            rec = eval_records(ga, m, p, op, dnas[:12])
            np.savez_compressed(os.path.join(out, "eval_big_blend01.npz"),
                                dna=dnas[:12].astype(np.int16), versions=VERSIONS, **rec)
            m._coverage_blend = 1.0
            for A in (1, 3):
                _, _, opA, _ = ref_shim.driver_params(A)
                opA['start_idx'] = op['start_idx'][:A]
                m._solo_sep_thresh = ref_shim.driver_params(A)[0]['solo_sep_thresh']
                dA = seeded_population(ga, m, p, opA, 16, 120, seed=A)
                rec = eval_records(ga, m, p, opA, dA)
                np.savez_compressed(
                    os.path.join(out, "eval_big_A%d.npz" % A), dna=dA.astype(np.int16),
                    versions=VERSIONS, solo_sep_thresh=m._solo_sep_thresh,
                    min_solo_lcs=opA['min_solo_lcs'], min_comb_lcs=opA['min_comb_lcs'],
                    max_dna_len=opA['max_dna_len'], **rec)
            m._solo_sep_thresh = mp['solo_sep_thresh']

            # ---- ranking: rackNstack / sortBy / lowVarSample on real organisms ----
            np.random.seed(3)
            orgs = []
            for _ in range(60):
                paths = [p.makeMeAPath(150, s) for s in op['start_idx']]
                orgs.append(ga.Organism(paths, m, p, op))
            pop = ga.GeneticalGorithm.__new__(ga.GeneticalGorithm)
            pop._cov_constr_0 = -gp['coverage_constr_0']
            pop._cov_constr_f = -gp['coverage_constr_f']
            pop._cov_aging = gp['coverage_aging']
            rank = dict(objs=np.array([[float(v) for v in o._obj_val] for o in orgs]))
            for gen in (0, 17, 80, 200):
                pop._gen_num = gen
                fit = np.array(pop.rackNstack(orgs))
                rank['fit_gen%d' % gen] = fit
                rank['sc_gen%d' % gen] = np.array([[float(v) for v in o._obj_val_sc] for o in orgs])
                with stable_argsort():
                    _, sfit = got.sortBy(list(range(len(orgs))), list(fit))
                    order = np.argsort(fit)
                rank['order_gen%d' % gen] = order
            # synthetic objective sets with exact ties / duplicates (F15)
            rs = np.random.RandomState(11)
            syn = np.stack([-rs.randint(0, 40, 200) / 40.0, rs.randint(1, 30, 200) * 0.125], 1)
            syn[rs.rand(200) < 0.4, 0] = 0.0

            class _O(object):
                pass
            sorgs = []
            for o in syn:
                x = _O()
                x._obj_val = [o[0], o[1]]
                x._obj_val_sc = [None, None]
                sorgs.append(x)
            pop._gen_num = 40
            rank['syn_objs'] = syn
            rank['syn_fit_gen40'] = np.array(pop.rackNstack(sorgs))
            # lowVarSample: capture chosen indices by sampling a list of ints
            saved = got.copy.deepcopy
            pop._gen_num = 17
            fit = rank['fit_gen17']
            lv = []
            for seed in range(6):
                np.random.seed(100 + seed)
                lv.append(got.lowVarSample(list(range(len(fit))), list(fit), gp['gamma']))
                np.random.seed(100 + seed)
                rank.setdefault('lowvar_r', []).append(np.random.uniform(0, 1 / len(fit)))
            got.copy.deepcopy = saved
            rank['lowvar_idx'] = np.array(lv)
            rank['lowvar_r'] = np.array(rank['lowvar_r'])
            rank['gamma'] = gp['gamma']
            np.savez_compressed(os.path.join(out, "rank.npz"), versions=VERSIONS, **rank)
            print("rank fixtures written")


if __name__ == "__main__":
    main()
