#!/usr/bin/env python
"""Graphs built by the UNMODIFIED reference PathMaker.computeTraversableGraph (pathmaker.py:119-155)
for the shipped waypoints of its two maps -> tests/golden/graph_ref.npz (N4 golden; the maps and
waypoints themselves are in world_small.npz / world_big.npz).  Also records whether the graph files
shipped in data/ equal what the shipped code produces.

  python tests/golden/make_golden_graph.py      (this container only: needs /root/reference)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402


def main():
    import cv2
    ga, mappy_m, pm_m, got = ref_shim.load()
    out = {}
    for name, img_f, g_f, w_f in (("small", "map_scaled.png", "wilk_3_graph.npy", "wilk_3_wpts.npy"),
                                  ("big", "big_map_scaled.png", "wilk_3_graph_big.npy", "wilk_3_wpts_big.npy")):
        mp, pp, op, gp = ref_shim.driver_params(6)
        img = cv2.imread(os.path.join(ref_shim.REF_DATA, img_f), cv2.IMREAD_GRAYSCALE) / 255
        m = mappy_m.Mappy(img, mp)
        p = pm_m.PathMaker(m, pp)
        p.loadWptsXY(os.path.join(ref_shim.REF_DATA, w_f))
        t0 = time.time()
        p.computeTraversableGraph()
        g = p._graph
        shipped = np.load(os.path.join(ref_shim.REF_DATA, g_f))
        src, dst = np.nonzero(g == 1)
        row_ptr = np.zeros(len(g) + 1, dtype=np.int64)
        np.add.at(row_ptr, src + 1, 1)
        out[name + "_row_ptr"] = np.cumsum(row_ptr)
        out[name + "_col"] = dst
        out[name + "_equals_shipped"] = bool(np.array_equal(g, shipped))
        out[name + "_max_traverse_dist"] = pp['max_traverse_dist']
        print("%s: N=%d E=%d in %.1fs, equals shipped graph: %s (shipped E=%d)" %
              (name, len(g), len(dst), time.time() - t0, out[name + "_equals_shipped"], int(shipped.sum())))
    np.savez_compressed(os.path.join(HERE, "graph_ref.npz"), **out)


if __name__ == "__main__":
    main()
