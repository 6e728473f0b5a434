#!/usr/bin/env python
"""Mappy.computeFrustums on the device vs the windowed numpy restatement.
  python profiles/run_frustums.py 4096 [16384]     (on a GPU box; one JSON line per map size)
Builds the synthetic map + graph, runs gcp_compute_frustums over all edges (kernel time from CUDA
events, call time by wall clock incl. host tables and copies), runs the host restatement on the
first `--host-nodes` source waypoints (all of them when it is 0) and compares those polygons."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    from genetic_coverage_planning_b200 import synth, frustums
    sizes = [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [4096]
    host_nodes = 20000
    for a in sys.argv[1:]:
        if a.startswith("--host-nodes="):
            host_nodes = int(a.split("=")[1])
    for size in sizes:
        mp, pp = dict(synth.DEFAULT_MAP_PARAMS), dict(synth.DEFAULT_PATH_PARAMS)
        u8 = synth.make_map(size, 0)
        img = u8.astype(np.float64)
        xy = synth.place_waypoints(u8, mp, None)
        row_ptr, col = synth.build_graph(u8, xy, mp, pp)
        N, E = len(xy), len(col)
        frustums.compute_frustums_device(img[:256, :256], xy[:0], np.zeros(1, np.int64), col[:0], mp)  # context warm-up
        t0 = time.time()
        poly, theta, dist, info = frustums.compute_frustums_device(img, xy, row_ptr, col, mp, return_info=True)
        t_call = time.time() - t0
        nodes = np.arange(N) if host_nodes == 0 or host_nodes >= N else np.arange(host_nodes)
        t0 = time.time()
        hp, ht, hd = synth.compute_frustums(img, xy, row_ptr, col, mp, only_nodes=nodes)
        t_host = time.time() - t0
        e_hi = int(row_ptr[nodes[-1] + 1])
        same = bool(np.array_equal(hp[:e_hi], poly[:e_hi]) and np.array_equal(ht[:e_hi], theta[:e_hi]) and
                    np.array_equal(hd[:e_hi], dist[:e_hi]))
        print(json.dumps({"map": "%d^2 synthetic" % size, "N": N, "E": E, "kernel_ms": info["kernel_ms"],
                          "device_call_s": t_call, "ambiguous_edges": info["ambiguous_edges"],
                          "edges_per_s_kernel": E / (info["kernel_ms"] * 1e-3),
                          "host_restatement_s": t_host, "host_edges": e_hi,
                          "host_edges_per_s_1core": e_hi / t_host,
                          "bit_identical_on_host_edges": same}), flush=True)
        # the reference's graph rule (nearest waypoint per 45-degree sector + line
        # collision check against the dilated map) on the same map and waypoints
        from genetic_coverage_planning_b200 import graphs
        hall, maxd = mp['narrowest_hall'], pp['max_traverse_dist']
        graphs.compute_traversable_graph_device(img[:256, :256], xy[:1], mp['scale'], maxd, hall)   # warm-up
        t0 = time.time()
        rp, cl, ginfo = graphs.compute_traversable_graph_device(img, xy, mp['scale'], maxd, hall, return_info=True)
        t_call = time.time() - t0
        nd = int(((hall / 2) / 3) / mp['scale'])
        t0 = time.time()
        obstacles = np.array(np.nonzero(graphs._dilated(img, nd))).T
        rows = np.arange(0, N, max(1, N // 64))[:64]
        ok = True
        for i in rows:
            got = graphs._row_host(int(i), xy, mp['scale'], maxd, obstacles)
            ok = ok and sorted(int(v) for v in got if v >= 0) == [int(v) for v in cl[rp[i]:rp[i + 1]]]
        t_host = time.time() - t0
        print(json.dumps({"map": "%d^2 synthetic" % size, "graph": "reference rule (pathmaker.py:119-155)", "N": N,
                          "E": int(len(cl)), "kernels_ms": ginfo["kernel_ms"], "device_call_s": t_call,
                          "ambiguous_rows": ginfo["ambiguous_rows"], "host_rows_checked": int(len(rows)),
                          "host_restatement_s_per_row": t_host / len(rows),
                          "host_restatement_s_extrapolated": t_host / len(rows) * N,
                          "rows_identical": bool(ok)}), flush=True)


if __name__ == "__main__":
    main()
