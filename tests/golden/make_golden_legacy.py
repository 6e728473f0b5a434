#!/usr/bin/env python
"""A population evolved and pickled by the UNMODIFIED reference (got.savePopulation,
gori_tools.py:252-253) -> tests/golden/legacy_population.pkl.gz, plus the objectives the reference
itself computed for its parents (legacy_population.npz).  Also records the graph the reference's
PathMaker.computeTraversableGraph (pathmaker.py:119-155) builds for the cropped map (N4 golden).

The world is a crop of the reference's small map (data/map_scaled.png) with the shipped waypoints
that fall inside it, so that the pickle (which holds one dense N x N x 17 x 2 frustum table per
deep-copied parent) stays small.

  python tests/golden/make_golden_legacy.py      (this container only: needs /root/reference)
"""
import gzip
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

CROP = (40, 200, 20, 200)      # r0, r1, c0, c1 of map_scaled.png
GEN_SIZE, N_GEN, AGENTS, L = 8, 2, 3, 60


def main():
    import cv2
    ga, mappy_m, pm_m, got = ref_shim.load()
    mp, pp, op, gp = ref_shim.driver_params(AGENTS)
    img = cv2.imread(os.path.join(ref_shim.REF_DATA, "map_scaled.png"), cv2.IMREAD_GRAYSCALE) / 255
    r0, r1, c0, c1 = CROP
    img = np.ascontiguousarray(img[r0:r1, c0:c1])
    xy = np.load(os.path.join(ref_shim.REF_DATA, "wilk_3_wpts.npy"))
    keep = (xy[:, 0] >= r0 + 8) & (xy[:, 0] < r1 - 8) & (xy[:, 1] >= c0 + 8) & (xy[:, 1] < c1 - 8)
    xy = xy[keep] - np.array([r0, c0])
    m = mappy_m.Mappy(img, mp)
    p = pm_m.PathMaker(m, pp)
    p._XY = xy.astype(int)
    p.computeTraversableGraph()                      # reference graph construction (+ findCrossoverCand)
    m.setWaypoints(p.waypoint_locs)
    m.computeFrustums(p._graph)
    deg = p._graph.sum(1)
    starts = [int(v) for v in np.argsort(-deg, kind='stable')[:AGENTS]]
    op = dict(op)
    op.update(start_idx=starts, max_dna_len=L, min_dna_len=10, min_solo_lcs=0, min_comb_lcs=0,
              num_muterpolations=5, crossover_time_thresh=30)
    gp = dict(gp)
    gp.update(gen_size=GEN_SIZE, starting_path_len=40, num_agents=AGENTS, coverage_constr_0=0.0,
              org_params=op)
    np.random.seed(0)
    pop = ga.GeneticalGorithm(m, p, gp)
    pop.runEvolution(N_GEN)
    raw = pickle.dumps(pop)                          # == got.savePopulation(pop, file)
    with gzip.open(os.path.join(HERE, "legacy_population.pkl.gz"), "wb", compresslevel=9) as f:
        f.write(raw)
    dna = np.array([o._dna for o in pop._gen_parent])
    obj = np.array([o._obj_val for o in pop._gen_parent], dtype=float)
    np.savez_compressed(os.path.join(HERE, "legacy_population.npz"), dna=dna, obj=obj,
                        fit=np.array(pop._gen_parent_fit, dtype=float), gen_num=pop._gen_num,
                        img=img, xy=p._XY, graph=p._graph.astype(np.uint8),
                        max_traverse_dist=pp['max_traverse_dist'], scale=mp['scale'],
                        narrowest_hall=mp['narrowest_hall'])
    print("N=%d E=%d raw pickle %.1f MB, gz %.2f MB, gen_num %d" %
          (len(xy), int(p._graph.sum()), len(raw) / 1e6,
           os.path.getsize(os.path.join(HERE, "legacy_population.pkl.gz")) / 1e6, pop._gen_num))
    print("obj", obj)


if __name__ == "__main__":
    main()
